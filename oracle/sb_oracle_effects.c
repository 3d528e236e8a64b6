/* TEST INFRASTRUCTURE -- see sb_oracle.h.  The 112 card effects (cards/XXXX.py activate_ability), one
 * case per card, each citing the reference file it restates.  p[] = the card's numeric ability_*
 * attributes as exported by tools/gen_card_table.py (names listed beside each row of
 * sb_card_table.inc).
 */
#include <string.h>
#include "sb_oracle.h"

#define CUR(g) ((g)->current_order)
#define PT_OFFBOARD (-2)

static Target T(int kind, int side) { Target t; memset(&t, 0, sizeof t); t.kind = kind; t.side = side; return t; }
static Target Tt(int kind, int side, int types) { Target t = T(kind, side); t.types = types; return t; }

/* board.at(point) that the card code dereferences without a None check -> AttributeError (Q11) */
static int need(Game *g, int pt) {
  int id = o_at_pt(g, pt);
  if (id < 0) ERR(g, SB_ERR_NONE_TARGET);
  return id;
}
static int choice(Game *g, const int *l, int n) {
  if (n <= 0) { ERR(g, SB_ERR_EMPTY_CHOICE); return PT_NONE; }
  return l[o_rng_below(g, n)];
}
/* list.sort(key=lambda t: (k1(t), random.random()), reverse=desc): one random() per element in list
 * order, then a stable sort (cards/b002.py:20, b008.py:24, b009.py:20, b104.py:19, s101.py:21) */
static void keyed_sort(Game *g, int *pts, const int *k1, int n, int desc) {
  double r[24];
  int k[24];
  for (int i = 0; i < n; i++) { r[i] = o_rng_random(g); k[i] = k1[i]; }
  for (int i = 1; i < n; i++) {
    int p = pts[i], kk = k[i];
    double rr = r[i];
    int j = i - 1;
    while (j >= 0) {
      int less = (k[j] < kk) || (k[j] == kk && r[j] < rr);      /* element j sorts before i ascending */
      int greater = (k[j] > kk) || (k[j] == kk && r[j] > rr);
      if (desc ? less : greater) { pts[j + 1] = pts[j]; k[j + 1] = k[j]; r[j + 1] = r[j]; j--; } else break;
    }
    pts[j + 1] = p; k[j + 1] = kk; r[j + 1] = rr;
  }
}
static int count_types_friendly(Game *g) { /* cards/up02.py:13-19, up03.py:14-20 */
  Target t = T(TK_UNIT, TS_FRIENDLY);
  int pts[24], n = o_get_targets(g, CUR(g), &t, PT_NONE, pts), m = 0, c = 0;
  for (int i = 0; i < n; i++) { int id = need(g, pts[i]); if (id >= 0) m |= g->e[id].types; }
  for (int i = 0; i < 16; i++) c += (m >> i) & 1;
  return c;
}

/* ---- Temple of Time memories (cards/b005.py:13,24-33): a forest of deep copies, see Mem in sb_oracle.h */
static int mem_push_entity(Game *g, int temple, int parent, int pos, const Ent *s) {
  if (g->n_mem >= NMEM_W) { ERR(g, SB_ERR_OVERFLOW); return -1; }
  Mem *m = &g->mem[g->n_mem];
  m->b005 = temple; m->parent = parent; m->pos = pos; m->card = s->card; m->owner = s->owner; m->is_struct = s->is_struct;
  m->fixed = s->fixed; m->strength = s->strength; m->detached = 0;
  for (int k = 0; k < 5; k++) m->st[k] = s->st[k];
  return g->n_mem++;
}
static int mem_copy_subtree(Game *g, int src, int new_parent, int limit) {
  if (g->n_mem >= NMEM_W) { ERR(g, SB_ERR_OVERFLOW); return -1; }
  int me = g->n_mem++;
  g->mem[me] = g->mem[src];
  g->mem[me].b005 = -1; g->mem[me].parent = new_parent; g->mem[me].detached = 1;
  for (int q = src + 1; q < limit; q++)
    if (g->mem[q].parent == src && mem_copy_subtree(g, q, me, limit) < 0) return -1;
  return me;
}
static void mem_delete_temple(Game *g, int temple) { /* self.ability_remembered = [] */
  int keep[NMEM_W], newidx[NMEM_W], w = 0;
  for (int i = 0; i < g->n_mem; i++) {
    const Mem *m = &g->mem[i];
    keep[i] = m->parent < 0 ? (m->b005 != temple) : keep[m->parent];
    newidx[i] = keep[i] ? w++ : -1;
  }
  for (int i = 0; i < g->n_mem; i++) if (keep[i]) {
    Mem m = g->mem[i];
    if (m.parent >= 0) m.parent = newidx[m.parent];
    g->mem[newidx[i]] = m;
  }
  g->n_mem = w;
}

void o_effect(Game *g, int id, int pos_pt, int has_source) {
  Ent *e = &g->e[id];
  const int *p = OCARDS[e->card].p;
  const int me = e->owner;
  int pts[24], n, tid;
  Target t;
  switch (e->card) {
  /* ------------------------------------------------------------ structures */
  case SBC_B002: { /* cards/b002.py:13-21 */
    t = T(TK_ANY, TS_ENEMY);
    n = o_get_targets(g, CUR(g), &t, PT_NONE, pts);
    if (n > 0) {
      int ky[24];
      for (int i = 0; i < n; i++) ky[i] = PTY(pts[i]);
      keyed_sort(g, pts, ky, n, 1);
      o_deal_damage_pt(g, pts[0], p[0], 1);
    }
    break; }
  case SBC_B004: { /* cards/b004.py:13-22 */
    t = T(TK_ANY, TS_ENEMY); t.base = 1;
    n = o_get_targets(g, CUR(g), &t, PT_NONE, pts);
    for (int i = 0; i < n; i++) { o_deal_damage_pt(g, pts[i], p[0], 1); if (g->err) return; }
    o_destroy(g, id, 1);
    break; }
  case SBC_B005: { /* cards/b005.py:15-33, including the memories of remembered temple copies (deepcopy) */
    t = T(TK_ANY, TS_FRIENDLY);
    n = o_surrounding(g, e->x, e->y, CUR(g), &t, pts);
    int mine = 0;
    for (int i = 0; i < g->n_mem; i++) if (g->mem[i].parent < 0 && g->mem[i].b005 == id) mine++;
    if (mine == 0) {
      for (int i = 0; i < n; i++) {
        tid = need(g, pts[i]);
        if (tid < 0) return;
        int r = mem_push_entity(g, id, -1, pts[i], &g->e[tid]);
        if (r < 0) return;
        if (g->e[tid].card == SBC_B005) { /* the copy carries a deep copy of that temple's own memories */
          int nm0 = g->n_mem;
          for (int q = 0; q < nm0; q++)
            if (g->mem[q].parent < 0 && g->mem[q].b005 == tid && mem_copy_subtree(g, q, r, nm0) < 0) return;
        }
      }
    } else {
      int count = 0, nm0 = g->n_mem;
      for (int i = 0; i < nm0 && count < p[0]; i++) {
        Mem *m = &g->mem[i];
        if (m->parent >= 0 || m->b005 != id) continue;
        int occ = o_at_pt(g, m->pos);
        if (occ < 0 || (g->e[occ].is_struct == m->is_struct && g->e[occ].card == m->card && g->e[occ].owner == m->owner)) {
          if (m->detached) { ERR(g, SB_ERR_UNSUPPORTED); return; } /* would live on a deep-copied board */
          int c = o_new_ent(g, m->card, m->owner, m->strength);
          g->e[c].fixed = m->fixed;
          for (int k = 0; k < 5; k++) g->e[c].st[k] = m->st[k];
          o_set(g, PTX(m->pos), PTY(m->pos), c);
          for (int q = i + 1; q < nm0; q++) /* the restored object keeps its own ability_remembered */
            if (g->mem[q].parent == i) { g->mem[q].parent = -1; g->mem[q].b005 = c; }
          count++;
        }
      }
      mem_delete_temple(g, id);
    }
    break; }
  case SBC_B006: { /* cards/b006.py:14-39 */
    t = T(TK_UNIT, TS_FRIENDLY);
    n = o_get_targets(g, CUR(g), &t, PT_NONE, pts);
    int sel[24], ns = 0;
    for (int i = 0; i < n; i++) { tid = need(g, pts[i]); if (tid >= 0 && g->e[tid].st[SB_ST_VITALIZED] == 0) sel[ns++] = pts[i]; }
    o_shuffle(g, sel, ns);
    for (int i = 0; i < ns && i < p[1]; i++) { tid = need(g, sel[i]); if (tid < 0) return; o_vitalize(g, tid); }
    int tiles[2], nt = 0, fr[5], bh[5];
    int nf = o_front(g, e->x, e->y, CUR(g), NULL, fr);
    int nb = o_behind(g, e->x, e->y, CUR(g), NULL, bh);
    if (nf > 0 && o_at_pt(g, fr[0]) < 0 && o_is_within_front_line(g, me, PTY(fr[0]))) tiles[nt++] = fr[0];
    if (nb > 0 && o_at_pt(g, bh[0]) < 0) tiles[nt++] = bh[0];
    if (nt > 0) {
      int c = o_new_ent(g, e->card, me, p[0]);
      int where = choice(g, tiles, nt);
      o_struct_play(g, c, PTX(where), PTY(where));
    }
    break; }
  case SBC_B007: { /* cards/b007.py:12-19 */
    int opp = o_opponent(g, me);
    if (g->pl[me].base == g->pl[opp].base) return;
    int stronger = g->pl[me].base > g->pl[opp].base ? me : opp;
    o_player_damage(g, stronger, p[0]);
    g->pl[o_opponent(g, stronger)].base += p[0];
    break; }
  case SBC_B008: { /* cards/b008.py:14-26 */
    Ply *pl = &g->pl[me];
    if (pl->n_hand > 0 && OCARDS[pl->hand[0].card].kind == KIND_UNIT) pl->hand[0].flags ^= SB_CF_FIXED;
    t = T(TK_UNIT, TS_ANY); t.status = 1 << SB_ST_CONFUSED;
    n = o_get_targets(g, CUR(g), &t, PT_NONE, pts);
    if (n > 0) {
      int ks[24];
      for (int i = 0; i < n; i++) ks[i] = g->e[o_at_pt(g, pts[i])].strength;
      keyed_sort(g, pts, ks, n, 0);
      tid = need(g, pts[0]);
      if (tid >= 0) o_destroy(g, tid, 1);
    }
    break; }
  case SBC_B009: { /* cards/b009.py:13-24 */
    t = T(TK_UNIT, TS_ENEMY);
    n = o_get_targets(g, CUR(g), &t, PT_NONE, pts);
    if (n > 0) {
      int ky[24];
      for (int i = 0; i < n; i++) ky[i] = PTY(pts[i]);
      keyed_sort(g, pts, ky, n, 1);
      if (n > p[0]) n = p[0];
      o_shuffle(g, pts, n);
      for (int i = 0; i < n; i++) { tid = need(g, pts[i]); if (tid < 0) return; o_confuse(g, tid); }
    }
    break; }
  case SBC_B104: { /* cards/b104.py:12-20 */
    t = T(TK_UNIT, TS_ENEMY);
    n = o_get_targets(g, CUR(g), &t, PT_NONE, pts);
    if (n > 0) {
      int ky[24];
      for (int i = 0; i < n; i++) ky[i] = PTY(pts[i]);
      keyed_sort(g, pts, ky, n, 1);
      tid = need(g, pts[0]);
      if (tid >= 0) o_freeze(g, tid);
    }
    break; }
  case SBC_B203: { /* cards/b203.py:12-21 */
    t = T(TK_UNIT, TS_FRIENDLY);
    n = o_front(g, e->x, e->y, me, &t, pts);
    for (int i = 0; i < n; i++) {
      tid = need(g, pts[i]);
      if (tid < 0) return;
      if (g->e[tid].is_struct) { ERR(g, SB_ERR_NONE_TARGET); return; }
      if (g->e[tid].st[SB_ST_CONFUSED] > 0) o_st_remove(g, tid, SB_ST_CONFUSED);
      o_command(g, tid);
    }
    break; }
  case SBC_B304: /* cards/b304.py:12-13 */
    o_struct_deal_damage(g, id, p[0], 0, 1);
    break;
  case SBC_B305: { /* cards/b305.py:16-45 */
    t = T(TK_STRUCTURE, TS_FRIENDLY);
    n = o_get_targets(g, CUR(g), &t, PT(e->x, e->y), pts);
    for (int i = 0; i < n && i < p[0]; i++) {
      tid = need(g, pts[i]);
      if (tid < 0) return;
      if (g->e[tid].card == e->card) {
        Target tu = T(TK_UNIT, TS_ANY);
        int sp[24], tx = PTX(pts[i]), ty = PTY(pts[i]);
        int ns = o_surrounding(g, tx, ty, CUR(g), &tu, sp);
        for (int k = 0; k < ns; k++) {
          int nx = PTX(sp[k]) - tx + e->x, ny = PTY(sp[k]) - ty + e->y;
          if (valid_xy(nx, ny)) { int u = need(g, sp[k]); if (u < 0) return; o_teleport(g, u, nx, ny); }
        }
        o_destroy(g, tid, 1);
        Ply *pl = &g->pl[me];
        if (pl->n_deck == 0) { ERR(g, SB_ERR_INDEX); return; }
        pl->deck[pl->n_deck - 1].cost = p[2];
        return;
      }
    }
    /* no other temple: the BOARD INSTANCE itself goes to the hand with cost 2 (a live link) */
    {
      Ply *pl = &g->pl[me];
      if (!e->single_use) { if (pl->n_deck == 0) { ERR(g, SB_ERR_INDEX); return; } pl->n_deck--; }
      if (pl->n_hand >= HAND_W) { ERR(g, SB_ERR_OVERFLOW); return; }
      CardRec r = {e->card, p[1], (e->single_use ? SB_CF_SINGLE_USE : 0) | SB_CF_OBJ, 0, 0, id};
      pl->hand[pl->n_hand++] = r;
    }
    break; }
  /* ------------------------------------------------------------ units */
  case SBC_U007: { /* cards/u007.py:13-21 */
    t = T(TK_UNIT, TS_ENEMY);
    n = o_surrounding(g, e->x, e->y, me, &t, pts);
    if (n > 0) { tid = need(g, choice(g, pts, n)); if (tid < 0) return; o_heal(g, tid, p[0]); o_vitalize(g, tid); }
    break; }
  case SBC_U017: { /* cards/u017.py:18-34 */
    Ply *pl = &g->pl[me];
    int cand[8], nc = 0;
    for (int i = 0; i < pl->n_hand; i++) if (OCARDS[pl->hand[i].card].kind == KIND_SPELL && pl->hand[i].cost <= 8) cand[nc++] = i;
    if (nc > 0) {
      o_shuffle(g, cand, nc);
      int remaining = 8, chosen[8], nch = 0;
      for (int i = 0; i < nc; i++) if (pl->hand[cand[i]].cost <= remaining) { chosen[nch++] = cand[i]; remaining -= pl->hand[cand[i]].cost; }
      for (int i = 0; i < nch; i++) {
        const OCard *c = &OCARDS[pl->hand[chosen[i]].card];
        int where = PT_NONE;
        if (c->has_target) {
          Target rt = {c->t_kind, c->t_side, c->t_types, c->t_xtypes, c->t_status, c->t_xstatus, c->t_limit >= 0, c->t_limit, c->t_nonhero, c->t_base};
          n = o_get_targets(g, CUR(g), &rt, PT_NONE, pts);
          where = choice(g, pts, n);
          if (g->err) return;
        }
        int idx = chosen[i];
        o_player_play(g, me, idx, where);
        if (g->err) return;
        for (int k = i + 1; k < nch; k++) if (chosen[k] > idx) chosen[k]--; /* hand.index(card) after the removal */
      }
    }
    break; }
  case SBC_U018: { /* cards/u018.py:13-26 */
    t = T(TK_UNIT, TS_ANY);
    n = o_surrounding(g, e->x, e->y, CUR(g), &t, pts);
    int m = 0, cnt = 0;
    for (int i = 0; i < n; i++) { tid = need(g, pts[i]); if (tid >= 0) m |= 1 << OCARDS[g->e[tid].card].first_type; }
    for (int i = 0; i < 16; i++) cnt += (m >> i) & 1;
    for (int k = 0; k < cnt; k++) {
      Target tb = T(TK_ANY, TS_ENEMY); tb.base = 1;
      n = o_get_targets(g, CUR(g), &tb, PT_NONE, pts);
      int where = choice(g, pts, n);
      if (g->err) return;
      o_deal_damage_pt(g, where, p[0], 1);
      if (g->err) return;
    }
    break; }
  case SBC_U021: { /* cards/u021.py:13-19 */
    t = T(TK_UNIT, TS_FRIENDLY);
    n = o_get_targets(g, CUR(g), &t, PT(e->x, e->y), pts);
    if (n > 0) { tid = need(g, choice(g, pts, n)); if (tid >= 0) o_heal(g, tid, p[0]); }
    break; }
  case SBC_U026: { /* cards/u026.py:13-16 */
    t = T(TK_ANY, TS_ENEMY);
    n = o_behind(g, e->x, e->y, CUR(g), &t, pts);
    if (n > 0) o_deal_damage_pt(g, pts[0], p[0], 1);
    break; }
  case SBC_U036: /* cards/u036.py:12-14 */
    if (g->pl[me].n_hand == 0) { tid = need(g, PT(e->x, e->y)); if (tid >= 0) o_heal(g, tid, p[0]); }
    break;
  case SBC_U040: /* cards/u040.py:13-22 (the print is dropped) */
    if (has_source) {
      n = o_surrounding(g, e->x, e->y, me, NULL, pts);
      if (n > 0) {
        int where = choice(g, pts, n);
        int c = o_new_ent(g, e->card, me, p[0]); /* unit.py:384-402 respawn: fresh instance of the class */
        o_set(g, PTX(where), PTY(where), c);
      }
    }
    break;
  case SBC_U050: /* cards/u050.py:13-15 */
    if (e->y == 4) o_gain_speed(g, id, p[0]);
    break;
  case SBC_U051: /* cards/u051.py:14-21 */
    t = T(TK_UNIT, TS_ANY);
    if (o_bordering(g, e->x, e->y, CUR(g), &t, pts) == 0) o_gain_speed(g, id, p[0]);
    else o_heal(g, id, p[1]);
    break;
  case SBC_U053: /* cards/u053.py:14-24: ability_amount=1, ability_movement=2 */
    t = T(TK_UNIT, TS_ANY);
    if (o_surrounding(g, e->x, e->y, CUR(g), &t, pts) == 0) o_gain_speed(g, id, p[1]);
    else if (o_bordering(g, e->x, e->y, CUR(g), &t, pts) == 0) o_gain_speed(g, id, p[0]);
    break;
  case SBC_U055: /* cards/u055.py:12-19 */
    t = T(TK_UNIT, TS_ENEMY); t.xstatus = 1 << SB_ST_CONFUSED;
    n = o_front(g, e->x, e->y, CUR(g), &t, pts);
    for (int i = 0; i < n; i++) { tid = need(g, pts[i]); if (tid < 0) return; o_confuse(g, tid); }
    break;
  case SBC_U061: /* cards/u061.py:12-23 */
    o_confuse(g, id);
    t = T(TK_UNIT, TS_FRIENDLY);
    n = o_get_targets(g, CUR(g), &t, PT(e->x, e->y), pts);
    if (n > 0) { tid = need(g, choice(g, pts, n)); if (tid < 0) return; o_confuse(g, tid); }
    o_gain_speed(g, id, 2);
    break;
  case SBC_U071: { /* cards/u071.py:12-27 */
    t = T(TK_UNIT, TS_ENEMY);
    n = o_bordering(g, e->x, e->y, me, &t, pts);
    int nc[8], nn = 0;
    for (int i = 0; i < n; i++) { tid = need(g, pts[i]); if (tid < 0) return; if (g->e[tid].st[SB_ST_CONFUSED] == 0) nc[nn++] = pts[i]; }
    if (nn > 0) {
      tid = need(g, choice(g, nc, nn));
      if (tid < 0) return;
      o_confuse(g, tid);
      int fr[5];
      int nf = o_front(g, e->x, e->y, CUR(g), NULL, fr);
      if (nf > 0 && o_at_pt(g, fr[0]) < 0) o_teleport(g, id, PTX(fr[0]), PTY(fr[0]));
    }
    break; }
  case SBC_U074: /* cards/u074.py:12-19 */
    t = T(TK_UNIT, TS_ENEMY);
    n = o_front(g, e->x, e->y, me, &t, pts);
    if (n > 0) o_force_attack(g, id, PTX(pts[0]), PTY(pts[0]));
    break;
  case SBC_U076: { /* cards/u076.py:14-26: ability_damage, ability_strength */
    t = T(TK_UNIT, TS_ANY); t.xtypes = 1 << UT_DRAGON;
    n = o_surrounding(g, e->x, e->y, CUR(g), &t, pts);
    if (n > 0) {
      tid = need(g, choice(g, pts, n));
      if (tid < 0) return;
      o_unit_deal_damage(g, tid, p[0], 0, 1);
      if (g->e[tid].strength <= 0) o_spawn_token_unit(g, me, PT(g->e[tid].x, g->e[tid].y), p[1], UT_DRAGON);
    }
    break; }
  case SBC_U101: { /* cards/u101.py:13-26 */
    if (pos_pt < 0 || pos_pt >= 20) return;
    tid = o_at_pt(g, pos_pt);
    if (tid < 0 || g->e[tid].is_struct || g->e[tid].st[SB_ST_FROZEN] == 0) return;
    t = T(TK_UNIT, TS_ENEMY); t.status = 1 << SB_ST_FROZEN;
    n = o_surrounding(g, e->x, e->y, me, &t, pts);
    for (int i = 0; i < n; i++) { o_deal_damage_pt(g, pts[i], p[0], 1); if (g->err) return; }
    break; }
  case SBC_U103: /* cards/u103.py:12-18 */
    t = T(TK_UNIT, TS_ENEMY);
    n = o_bordering(g, e->x, e->y, CUR(g), &t, pts);
    for (int i = 0; i < n; i++) { tid = need(g, pts[i]); if (tid < 0) return; o_freeze(g, tid); }
    break;
  case SBC_U106: /* cards/u106.py:13-18 */
    t = T(TK_STRUCTURE, TS_FRIENDLY);
    if (o_bordering(g, e->x, e->y, CUR(g), &t, pts) > 0 || e->y == 4) o_heal(g, id, p[0]);
    break;
  case SBC_U111: /* cards/u111.py:13-21 */
    for (int k = 0; k < p[0]; k++) {
      t = T(TK_UNIT, TS_FRIENDLY);
      n = o_surrounding(g, e->x, e->y, me, &t, pts);
      if (n > 0) { tid = need(g, choice(g, pts, n)); if (tid < 0) return; o_heal(g, tid, 1); }
    }
    break;
  case SBC_U117: o_freeze(g, id); break; /* cards/u117.py:11-12 */
  case SBC_U206: o_player_damage(g, me, p[0]); break; /* cards/u206.py:12-13 */
  case SBC_U211: { /* cards/u211.py:14-19: ability_max_strength, ability_min_strength */
    n = o_front(g, e->x, e->y, CUR(g), NULL, pts);
    if (n > 0 && o_at_pt(g, pts[0]) < 0) {
      int s = p[1] + o_rng_below(g, p[0] + 1 - p[1]);
      o_spawn_token_unit(g, me, pts[0], s, UT_SATYR);
    }
    break; }
  case SBC_U216: o_player_damage(g, me, p[0]); break; /* cards/u216.py:12-13 */
  case SBC_U217: { /* cards/u217.py:13-17 */
    int row[4], nr = 0;
    for (int x = 0; x < 4; x++) if (g->board[4][x] < 0) row[nr++] = PT(x, 4);
    if (nr > 0) o_spawn_token_unit(g, me, choice(g, row, nr), p[0], UT_SATYR);
    break; }
  case SBC_U302: { /* cards/u302.py:12-23 */
    if (pos_pt < 0 || pos_pt >= 20) return;
    tid = o_at_pt(g, pos_pt);
    if (tid < 0 || g->e[tid].is_struct) return;
    if (g->e[tid].strength > e->strength) {
      o_unit_deal_damage(g, tid, p[0], 0, 1);
      if (g->e[tid].strength > 0) o_push(g, tid, e->x, e->y);
    }
    break; }
  case SBC_U305: /* cards/u305.py:13-21 */
    t = Tt(TK_UNIT, TS_FRIENDLY, 1 << UT_CONSTRUCT);
    n = o_bordering(g, e->x, e->y, CUR(g), &t, pts);
    if (n > 0) { tid = need(g, choice(g, pts, n)); if (tid < 0) return; o_heal(g, tid, p[0]); o_heal(g, id, p[0]); }
    break;
  case SBC_U306: /* cards/u306.py:13-20 */
    t = T(TK_ANY, TS_FRIENDLY);
    n = o_get_targets(g, CUR(g), &t, PT(e->x, e->y), pts);
    if (n > 0) o_deal_damage_pt(g, choice(g, pts, n), p[0], 1);
    break;
  case SBC_U310: { /* cards/u310.py:12-41 */
    t = T(TK_UNIT, TS_ENEMY);
    n = o_bordering(g, e->x, e->y, CUR(g), &t, pts);
    int behind = PT_NONE, right = PT_NONE, left = PT_NONE, front = PT_NONE;
    for (int i = n - 1; i >= 0; i--) {
      if (PTY(pts[i]) == e->y + 1) behind = pts[i];
      if (PTX(pts[i]) == e->x + 1) right = pts[i];
      if (PTX(pts[i]) == e->x - 1) left = pts[i];
      if (PTY(pts[i]) == e->y - 1) front = pts[i];
    }
    int target = PT_NONE;
    if (behind != PT_NONE && PTY(behind) < 4 && o_at(g, PTX(behind), PTY(behind) + 1) < 0) target = behind;
    else if (left != PT_NONE && PTX(left) > 0 && o_at(g, PTX(left) - 1, PTY(left)) < 0) target = left;
    else if (right != PT_NONE && PTX(right) < 3 && o_at(g, PTX(right) + 1, PTY(right)) < 0) target = right;
    else if (front != PT_NONE && PTY(front) > 0 && o_at(g, PTX(front), PTY(front) - 1) < 0) target = front;
    if (target == PT_NONE) { ERR(g, SB_ERR_INDEX); return; } /* UnboundLocalError */
    tid = need(g, target);
    if (tid >= 0) o_push(g, tid, e->x, e->y);
    break; }
  case SBC_U313: /* cards/u313.py:13-21 */
    t = T(TK_UNIT, TS_FRIENDLY);
    n = o_surrounding(g, e->x, e->y, CUR(g), &t, pts);
    if (n > 0) { tid = need(g, choice(g, pts, n)); if (tid < 0) return; o_heal(g, tid, p[0]); }
    o_heal(g, id, p[0]);
    break;
  case SBC_U314: /* cards/u314.py:11-17 */
    t = T(TK_UNIT, TS_FRIENDLY);
    n = o_front(g, e->x, e->y, CUR(g), &t, pts);
    if (n > 0) { tid = need(g, pts[0]); if (tid >= 0) o_push(g, tid, e->x, e->y); }
    break;
  case SBC_U316: /* cards/u316.py:13-23 */
    t = T(TK_UNIT, TS_FRIENDLY);
    n = o_surrounding(g, e->x, e->y, CUR(g), &t, pts);
    if (n > 0) {
      o_shuffle(g, pts, n);
      for (int i = 0; i < n && i < p[0]; i++) { tid = need(g, pts[i]); if (tid < 0) return; o_vitalize(g, tid); }
    }
    o_vitalize(g, id);
    break;
  case SBC_U320: /* cards/u320.py:13-19 */
    t = T(TK_UNIT, TS_FRIENDLY);
    n = o_surrounding(g, e->x, e->y, me, &t, pts);
    if (n > 0) { tid = need(g, choice(g, pts, n)); if (tid >= 0) o_heal(g, tid, p[0]); }
    break;
  case SBC_U401: /* cards/u401.py:13-22: `damage` */
    t = T(TK_UNIT, TS_ANY);
    n = o_bordering(g, e->x, e->y, CUR(g), &t, pts);
    for (int i = 0; i < n; i++) {
      tid = o_at_pt(g, pts[i]);
      if (tid >= 0) {
        if (g->e[tid].is_struct) o_struct_deal_damage(g, tid, p[0], 0, 1); else o_unit_deal_damage(g, tid, p[0], 0, 1);
        if (g->e[tid].is_struct) { ERR(g, SB_ERR_NONE_TARGET); return; }
        o_poison(g, tid);
      }
    }
    break;
  case SBC_U403: { /* cards/u403.py:14-24: ability_amount, ability_strength */
    t = T(TK_UNIT, TS_ANY); t.status = 1 << SB_ST_POISONED;
    n = o_surrounding(g, e->x, e->y, CUR(g), &t, pts);
    for (int i = 0; i < n; i++) {
      int bt[4], em[4], ne = 0;
      int nb = o_bordering(g, PTX(pts[i]), PTY(pts[i]), CUR(g), NULL, bt);
      for (int k = 0; k < nb; k++) if (o_at_pt(g, bt[k]) < 0) em[ne++] = bt[k];
      o_shuffle(g, em, ne);
      for (int k = 0; k < ne && k < p[0]; k++) o_spawn_token_unit(g, me, em[k], p[1], UT_TOAD);
    }
    break; }
  case SBC_U405: /* cards/u405.py:13-21 */
    t = T(TK_UNIT, TS_ANY);
    n = o_bordering(g, e->x, e->y, CUR(g), &t, pts);
    for (int i = 0; i < n; i++) {
      int dealt = o_deal_damage_pt(g, pts[i], p[0], 1);
      if (g->err) return;
      o_heal(g, id, dealt);
    }
    break;
  case SBC_U406: { /* cards/u406.py:13-20 */
    int bt[4], em[4], ne = 0;
    int nb = o_bordering(g, e->x, e->y, CUR(g), NULL, bt);
    for (int k = 0; k < nb; k++) if (o_at_pt(g, bt[k]) < 0) em[ne++] = bt[k];
    if (ne > 0) { int where = choice(g, em, ne); o_spawn_token_unit(g, o_opponent(g, me), where, p[0], UT_RAVEN); }
    break; }
  case SBC_U411: /* cards/u411.py:13-22 */
    t = T(TK_UNIT, TS_ENEMY);
    n = o_get_targets(g, CUR(g), &t, PT_NONE, pts);
    if (n > 0) {
      o_shuffle(g, pts, n);
      for (int i = 0; i < n && i < p[0]; i++) { tid = need(g, pts[i]); if (tid < 0) return; o_poison(g, tid); }
    }
    break;
  case SBC_UA03: { /* cards/ua03.py:12-15 */
    int row[5], nr = 0;
    for (int x = 0; x < 4; x++) if (g->board[e->y][x] < 0) row[nr++] = PT(x, e->y);
    row[nr++] = PT(e->x, e->y);
    int where = choice(g, row, nr);
    o_teleport(g, id, PTX(where), PTY(where));
    break; }
  case SBC_UA04: { /* cards/ua04.py:12-33 */
    int lp[24], rp[24], lu[24], ru[24];
    Target tf = T(TK_UNIT, TS_FRIENDLY), te = T(TK_UNIT, TS_ENEMY);
    int nl = o_get_targets(g, CUR(g), &tf, PT_NONE, lp);
    for (int i = 0; i < nl; i++) lu[i] = o_at_pt(g, lp[i]);
    int nr = o_get_targets(g, CUR(g), &te, PT_NONE, rp);
    for (int i = 0; i < nr; i++) ru[i] = o_at_pt(g, rp[i]);
    int *src = NULL, ns = 0, sel[24], nsel = 0;
    if (nl > nr) { src = lu; ns = nl; } else if (nr > nl) { src = ru; ns = nr; }
    if (src) {
      int mn = g->e[src[0]].strength;
      for (int i = 1; i < ns; i++) if (g->e[src[i]].strength < mn) mn = g->e[src[i]].strength;
      for (int i = 0; i < ns; i++) if (g->e[src[i]].strength == mn) sel[nsel++] = src[i];
    }
    if (nsel > 0) o_destroy(g, sel[o_rng_below(g, nsel)], 1);
    break; }
  case SBC_UA05: { /* cards/ua05.py:13-19 */
    int sd[2], em[2], ne = 0;
    int ns = o_side(g, e->x, e->y, CUR(g), NULL, sd);
    for (int k = 0; k < ns; k++) if (o_at_pt(g, sd[k]) < 0) em[ne++] = sd[k];
    for (int k = 0; k < ne; k++) o_spawn_token_unit(g, me, em[k], p[0], UT_ANCIENT);
    break; }
  case SBC_UA07: /* cards/ua07.py:11-22 */
    switch (o_rng_below(g, 5)) {
      case 0: o_freeze(g, id); break;
      case 1: o_poison(g, id); break;
      case 2: o_vitalize(g, id); break;
      case 3: o_confuse(g, id); break;
      case 4: o_disable(g, id); break;
    }
    break;
  case SBC_UA20: { /* cards/ua20.py:21-32 */
    static const int cand[4] = {SBC_B005, SBC_B006, SBC_B203, SBC_B305};
    t = T(TK_UNIT, TS_ENEMY);
    if (o_front(g, e->x, e->y, me, &t, pts) == 0) {
      Ply *pl = &g->pl[me];
      int c = cand[o_rng_below(g, 4)];
      if (pl->n_deck >= DECK_W) { ERR(g, SB_ERR_OVERFLOW); return; }
      CardRec r = {c, p[0], SB_CF_SINGLE_USE, 0, 0, -1};
      pl->deck[pl->n_deck++] = r;
    }
    break; }
  case SBC_UD01: /* cards/ud01.py:13-19 */
    t = Tt(TK_UNIT, TS_FRIENDLY, 1 << UT_DRAGON);
    n = o_get_targets(g, me, &t, PT_NONE, pts);
    if (n > 0) { tid = need(g, choice(g, pts, n)); if (tid >= 0) o_heal(g, tid, p[0]); }
    break;
  case SBC_UD02: /* cards/ud02.py:13-19 */
    t = T(TK_UNIT, TS_ANY); t.xtypes = 1 << UT_DRAGON;
    n = o_front(g, e->x, e->y, CUR(g), &t, pts);
    for (int i = 0; i < n; i++) { o_deal_damage_pt(g, pts[i], p[0], 1); if (g->err) return; }
    break;
  case SBC_UD31: /* cards/ud31.py:13-24 */
    if (pos_pt < 0 || pos_pt >= 20) return;
    tid = o_at_pt(g, pos_pt);
    if (tid < 0 || g->e[tid].is_struct) return;
    t = Tt(TK_UNIT, TS_FRIENDLY, 1 << UT_DRAGON);
    n = o_surrounding(g, e->x, e->y, me, &t, pts);
    if (n > 0) { tid = need(g, choice(g, pts, n)); if (tid < 0) return; o_heal(g, tid, p[0]); }
    o_heal(g, id, p[0]);
    break;
  case SBC_UE01: { /* cards/ue01.py:11-19 */
    int times = e->damage_taken;
    for (int k = 0; k < times; k++) {
      t = T(TK_UNIT, TS_ENEMY);
      n = o_get_targets(g, me, &t, PT_NONE, pts);
      if (n > 0) { o_deal_damage_pt(g, choice(g, pts, n), 1, 1); if (g->err) return; }
    }
    break; }
  case SBC_UE03: { /* cards/ue03.py:11-17 */
    int bt[4], em[4], ne = 0;
    int nb = o_bordering(g, e->x, e->y, CUR(g), NULL, bt);
    for (int k = 0; k < nb; k++) if (o_at_pt(g, bt[k]) < 0) em[ne++] = bt[k];
    if (ne > 0) o_spawn_token_unit(g, me, choice(g, em, ne), e->strength, UT_ELDER);
    break; }
  case SBC_UE04: { /* cards/ue04.py:12-18 */
    t = T(TK_UNIT, TS_ENEMY);
    n = o_get_targets(g, me, &t, PT_NONE, pts);
    int c = 0;
    for (int i = 0; i < n; i++) { tid = need(g, pts[i]); if (tid >= 0 && g->e[tid].strength > e->strength) c++; }
    o_heal(g, id, c * p[0]);
    break; }
  case SBC_UE05: /* cards/ue05.py:12-19 */
    t = T(TK_UNIT, TS_FRIENDLY); t.has_limit = 1; t.limit = e->strength - 1;
    n = o_get_targets(g, me, &t, PT(e->x, e->y), pts);
    for (int i = 0; i < n && i < p[0]; i++) { tid = need(g, pts[i]); if (tid < 0) return; g->e[tid].strength = e->strength; }
    break;
  case SBC_UE11: o_heal(g, id, p[0]); break; /* cards/ue11.py:12-13 */
  case SBC_UE12: /* cards/ue12.py:11-18 */
    t = T(TK_UNIT, TS_ENEMY);
    n = o_front(g, e->x, e->y, me, &t, pts);
    for (int i = 0; i < n; i++) { tid = need(g, pts[i]); if (tid < 0) return; o_destroy(g, tid, 1); }
    break;
  case SBC_UE21: /* cards/ue21.py:11-18 */
    t = T(TK_UNIT, TS_FRIENDLY); t.has_limit = 1; t.limit = e->strength;
    n = o_get_targets(g, me, &t, PT(e->x, e->y), pts);
    for (int i = 0; i < n; i++) {
      tid = need(g, pts[i]);
      if (tid < 0) return;
      if (g->e[tid].is_struct) { ERR(g, SB_ERR_NONE_TARGET); return; }
      o_command(g, tid);
    }
    break;
  case SBC_UE22: /* cards/ue22.py:12-22 */
    t = T(TK_UNIT, TS_FRIENDLY);
    n = o_get_targets(g, me, &t, PT(e->x, e->y), pts);
    if (n > 0) {
      o_shuffle(g, pts, n);
      for (int i = 0; i < n && i < p[0]; i++) { tid = need(g, pts[i]); if (tid < 0) return; o_heal(g, tid, e->damage_taken); }
    }
    break;
  case SBC_UE31: e->strength = p[0]; break; /* cards/ue31.py:12-13 */
  case SBC_UE32: { /* cards/ue32.py:12-22 */
    int damage = e->strength < 6 ? e->strength : 6;
    t = T(TK_ANY, TS_ENEMY);
    n = o_front(g, e->x, e->y, me, &t, pts);
    if (n > 0) o_deal_damage_pt(g, pts[0], damage, 1);
    else o_player_damage(g, o_opponent(g, me), damage);
    break; }
  case SBC_UE41: o_convert(g, id); break; /* cards/ue41.py:10-11 */
  case SBC_UE42: { /* cards/ue42.py:12-15 */
    int amount = e->damage_taken < p[0] ? e->damage_taken : p[0];
    o_heal(g, id, o_player_damage(g, o_opponent(g, me), amount));
    break; }
  case SBC_UP02: o_heal(g, id, p[0] * count_types_friendly(g)); break; /* cards/up02.py:12-20 */
  case SBC_UP03: { /* cards/up03.py:12-28 */
    int c = count_types_friendly(g);
    t = T(TK_UNIT, TS_ENEMY);
    n = o_surrounding(g, e->x, e->y, CUR(g), &t, pts);
    for (int i = 0; i < n; i++) {
      tid = need(g, pts[i]);
      if (tid < 0) return;
      int s = g->e[tid].strength - p[0] * c;
      g->e[tid].strength = s > 1 ? s : 1; /* unit.py:233-234 reduce */
    }
    break; }
  default: break;
  }
  (void)pos_pt;
}

void o_spell_effect(Game *g, int card, int caster, int pos_pt) {
  const int *p = OCARDS[card].p;
  int pts[24], n, tid;
  Target t;
  switch (card) {
  case SBC_S001: o_deal_damage_pt(g, pos_pt, p[0], 1); break; /* cards/s001.py:13-14 */
  case SBC_S003: /* cards/s003.py:14-19: ability_max_damage, ability_min_damage */
    t = T(TK_ANY, TS_ENEMY);
    n = o_get_targets(g, CUR(g), &t, PT_NONE, pts);
    for (int i = 0; i < n; i++) {
      if (need(g, pts[i]) < 0) return;
      o_deal_damage_pt(g, pts[i], p[1] + o_rng_below(g, p[0] + 1 - p[1]), 1);
      if (g->err) return;
    }
    break;
  case SBC_S004: { /* cards/s004.py:14-23: ability_max_amount, ability_min_amount */
    int tl[24], em[24], ne = 0;
    int nt = o_get_within_front_line(g, caster, tl);
    for (int i = 0; i < nt; i++) if (o_at_pt(g, tl[i]) < 0) em[ne++] = tl[i];
    if (ne > 0) {
      o_shuffle(g, em, ne);
      int amount = p[1] + o_rng_below(g, p[0] + 1 - p[1]);
      for (int i = 0; i < ne && i < amount; i++) o_spawn_token_unit(g, caster, em[i], 1, UT_TOAD);
    }
    break; }
  case SBC_S007: /* cards/s007.py:13-16 */
    tid = need(g, pos_pt);
    if (tid < 0) return;
    o_heal(g, tid, p[0]); o_vitalize(g, tid);
    break;
  case SBC_S012: { /* cards/s012.py:13-18 */
    int tl[24], em[24], ne = 0;
    int nt = o_get_within_front_line(g, caster, tl);
    for (int i = 0; i < nt; i++) if (o_at_pt(g, tl[i]) < 0) em[ne++] = tl[i];
    if (ne > 0) o_spawn_token_unit(g, caster, choice(g, em, ne), p[0], UT_KNIGHT);
    break; }
  case SBC_S013: { /* cards/s013.py:13-27 */
    int chosen[16], nc = 0;
    for (int ut = 0; ut < 16; ut++) {
      t = Tt(TK_UNIT, TS_ANY, 1 << ut);
      n = o_get_targets(g, CUR(g), &t, PT_NONE, pts);
      int units[24], nu = 0;
      for (int i = 0; i < n; i++) {
        int dup = 0;
        for (int k = 0; k < nc; k++) if (chosen[k] == pts[i]) dup = 1;
        if (!dup) units[nu++] = pts[i];
      }
      if (nu > 0) chosen[nc++] = choice(g, units, nu);
    }
    for (int i = 0; i < nc; i++) { o_deal_damage_pt(g, chosen[i], p[0], 1); if (g->err) return; }
    break; }
  case SBC_S021: { /* cards/s021.py:13-25 */
    tid = need(g, pos_pt);
    if (tid < 0) return;
    o_confuse(g, tid);
    t = Tt(TK_UNIT, TS_FRIENDLY, 1 << UT_FELINE);
    n = o_get_targets(g, CUR(g), &t, PT_NONE, pts);
    if (n > 0) {
      int mn = 1 << 30, wk[24], nw = 0;
      for (int i = 0; i < n; i++) { int s = g->e[o_at_pt(g, pts[i])].strength; if (s < mn) mn = s; }
      for (int i = 0; i < n; i++) if (g->e[o_at_pt(g, pts[i])].strength == mn) wk[nw++] = pts[i];
      tid = need(g, choice(g, wk, nw));
      if (tid >= 0) o_heal(g, tid, p[0]);
    }
    break; }
  case SBC_S101: { /* cards/s101.py:14-22: ability_mana, ability_strength */
    g->pl[caster].mana += p[0];
    t = T(TK_UNIT, TS_FRIENDLY);
    n = o_get_targets(g, CUR(g), &t, PT_NONE, pts);
    if (n == 0) { ERR(g, SB_ERR_INDEX); return; }
    int ks[24];
    for (int i = 0; i < n; i++) ks[i] = g->e[o_at_pt(g, pts[i])].strength;
    keyed_sort(g, pts, ks, n, 0);
    tid = need(g, pts[0]);
    if (tid >= 0) o_heal(g, tid, p[1]);
    break; }
  case SBC_S104: /* cards/s104.py:13-19 */
    tid = need(g, pos_pt);
    if (tid < 0) return;
    if (g->e[tid].st[SB_ST_FROZEN] > 0) o_unit_deal_damage(g, tid, p[0], 0, 1); else o_freeze(g, tid);
    break;
  case SBC_S105: /* cards/s105.py:13-14 */
    tid = need(g, pos_pt);
    if (tid >= 0) o_heal(g, tid, p[0]);
    break;
  case SBC_S203: { /* cards/s203.py:14-30.  list(set(...)) order is str-hash dependent in the reference (Q14):
                    canonical first-occurrence order here (documented deviation). */
    int fr[24], all[64], na = 0;
    t = T(TK_UNIT, TS_FRIENDLY);
    int nf = o_get_targets(g, CUR(g), &t, PT_NONE, fr);
    for (int i = 0; i < nf; i++) {
      Target te = T(TK_ANY, TS_ENEMY); te.base = 1;
      n = o_surrounding(g, PTX(fr[i]), PTY(fr[i]), CUR(g), &te, pts);
      for (int k = 0; k < n; k++) {
        int dup = 0;
        for (int q = 0; q < na; q++) if (all[q] == pts[k]) dup = 1;
        if (!dup) all[na++] = pts[k];
      }
    }
    for (int i = 0; i < na; i++) { o_deal_damage_pt(g, all[i], p[0], 1); if (g->err) return; }
    break; }
  case SBC_S302: /* cards/s302.py:14-22: ability_damage, ability_targets */
    t = T(TK_ANY, TS_ENEMY); t.base = 1;
    n = o_get_targets(g, CUR(g), &t, PT_NONE, pts);
    o_shuffle(g, pts, n);
    for (int i = 0; i < n && i < p[1]; i++) { o_deal_damage_pt(g, pts[i], p[0], 1); if (g->err) return; }
    break;
  case SBC_S403: /* cards/s403.py:13-22 */
    t = T(TK_UNIT, TS_FRIENDLY); t.status = 1 << SB_ST_POISONED;
    n = o_get_targets(g, CUR(g), &t, PT_NONE, pts);
    for (int i = 0; i < n; i++) { tid = need(g, pts[i]); if (tid < 0) return; o_heal(g, tid, p[0]); o_vitalize(g, tid); }
    break;
  default: break;
  }
}
