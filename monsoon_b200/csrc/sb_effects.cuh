// sb_effects.cuh -- the 112 card effects (reference cards/<id>.py activate_ability) as one dispatch
// per card over the primitives of sb_engine.cuh.  p[] = the card's ability_* attributes exported by
// tools/gen_card_table.py (names are listed beside each row of sb_card_table.inc).
#pragma once
#include "sb_engine.cuh"

#define CUR(g) ((g).current_order)

SBD_FI int need(G& g, int pt) {  // board.at(pt) dereferenced without a None check -> AttributeError (Q11)
  int id = at_pt(g, pt);
  if (id < 0) GERR(g, SB_ERR_NONE_TARGET);
  return id;
}
SBD_FI int choice_pt(G& g, const i8* l, int n) {
  if (n <= 0) { GERR(g, SB_ERR_EMPTY_CHOICE); return PT_NONE; }
  return l[rng_below(g, n)];
}
// list.sort(key=lambda t: (k1(t), random.random()), reverse=desc): one random() per element in list order,
// then a stable sort (cards/b002.py:20, b008.py:24, b009.py:20, b104.py:19, s101.py:21)
SBD_NI void keyed_sort(G& g, i8* pts, const i16* k1, int n, bool desc) {
  G_LOCAL(g);
  P_LOCAL(pts);
  P_LOCAL(k1);
  double r[22];
  i16 k[22];
  #pragma unroll 1
  for (int i = 0; i < n; i++) { r[i] = rng_random(g); k[i] = k1[i]; }
  #pragma unroll 1
  for (int i = 1; i < n; i++) {
    i8 p = pts[i]; i16 kk = k[i]; double rr = r[i];
    int j = i - 1;
    #pragma unroll 1
    while (j >= 0) {
      bool less = (k[j] < kk) || (k[j] == kk && r[j] < rr);
      bool greater = (k[j] > kk) || (k[j] == kk && r[j] > rr);
      if (desc ? less : greater) { pts[j + 1] = pts[j]; k[j + 1] = k[j]; r[j + 1] = r[j]; j--; } else break;
    }
    pts[j + 1] = p; k[j + 1] = kk; r[j + 1] = rr;
  }
}
SBD_NI int count_types_friendly(G& g) {  // cards/up02.py:13-19, up03.py:14-20
  G_LOCAL(g);
  Target t = mkT(TK_UNIT, TS_FRIENDLY);
  i8 pts[22];
  int n = get_targets(g, CUR(g), t, PT_NONE, pts);
  u32 m = 0;
  #pragma unroll 1
  for (int i = 0; i < n; i++) m |= CARD(g, g.e[at_pt(g, pts[i])].card).types;
  return __popc(m);
}
SBD_FI int empty_of(const G& g, const i8* in, int n, i8* out) {
  int k = 0;
  #pragma unroll 1
  for (int i = 0; i < n; i++) if (at_pt(g, in[i]) < 0) out[k++] = in[i];
  return k;
}
// "frontmost enemy" family: get_targets -> keyed sort on y desc -> take `take`
SBD_NI int frontmost(G& g, const Target& t, i8* pts) {
  G_LOCAL(g);
  P_LOCAL(&t);
  P_LOCAL(pts);
  int n = get_targets(g, CUR(g), t, PT_NONE, pts);
  if (n > 0) {
    i16 ky[22];
    #pragma unroll 1
    for (int i = 0; i < n; i++) ky[i] = (i16)PTY(pts[i]);
    keyed_sort(g, pts, ky, n, true);
  }
  return n;
}

// ---- Temple of Time memories (cards/b005.py:13,24-33): a forest of deep copies, see Mem in sb_engine.cuh
SBD_NI int mem_push_entity(G& g, int temple, int parent, int pos, const Ent& s) {
  G_LOCAL(g);
  if (g.n_mem >= NMEM) { GERR(g, SB_ERR_OVERFLOW); return -1; }
  Mem& m = g.mem[g.n_mem];
  m.b005 = (i8)temple; m.parent = (i8)parent; m.pos = (u8)pos; m.card = s.card; m.fl = s.fl & (EF_OWNER | EF_STRUCT | EF_FIXED);
  m.strength = s.strength;
  #pragma unroll 1
  for (int k = 0; k < 5; k++) m.st[k] = s.st[k];
  return g.n_mem++;
}
// deep copy of the subtree rooted at mem[src] under new_parent (iterative: parents precede children in the array)
SBD_NI int mem_copy_subtree(G& g, int src, int new_parent, int limit) {
  G_LOCAL(g);
  i8 map[NMEM];
  #pragma unroll 1
  for (int q = 0; q < NMEM; q++) map[q] = -1;
  int root = -1;
  #pragma unroll 1
  for (int q = src; q < limit; q++) {
    const int par = g.mem[q].parent;
    const bool is_root = (q == src);
    if (!is_root && (par < 0 || map[par] < 0)) continue;
    if (g.n_mem >= NMEM) { GERR(g, SB_ERR_OVERFLOW); return -1; }
    const int me = g.n_mem++;
    g.mem[me] = g.mem[q];
    g.mem[me].b005 = -1;
    g.mem[me].fl |= EF_SINGLE;  // detached: deep-copied below another temple's copy (lives on a cloned board once restored)
    g.mem[me].parent = (i8)(is_root ? new_parent : map[par]);
    map[q] = (i8)me;
    if (is_root) root = me;
  }
  return root;
}
SBD_NI void mem_delete_temple(G& g, int temple) {  // self.ability_remembered = []
  G_LOCAL(g);
  u8 keep[NMEM], nidx[NMEM];
  int w = 0;
  #pragma unroll 1
  for (int i = 0; i < g.n_mem; i++) {
    const Mem& m = g.mem[i];
    bool k = m.parent < 0 ? (m.b005 != temple) : (keep[m.parent] != 0);
    keep[i] = k; nidx[i] = k ? (u8)w++ : (u8)0xFF;
  }
  #pragma unroll 1
  for (int i = 0; i < g.n_mem; i++) if (keep[i]) {
    Mem m = g.mem[i];
    if (m.parent >= 0) m.parent = (i8)nidx[m.parent];
    g.mem[nidx[i]] = m;
  }
  g.n_mem = (u8)w;
}

SBD_NI void effect(G& g, int id, int pos_pt, int has_source) {
  G_LOCAL(g);
  Ent& e = g.e[id];
  const DCard& cd = CARD(g, e.card);
  const i8* p = cd.p;
  const int me = ent_owner(e);
  const int ex = e.x, ey = e.y;
  i8 pts[22];
  int n, tid;
  Target t;
  switch (e.card) {
    // ------------------------------------------------------------ structures
    case SBC_B002:  // cards/b002.py:13-21
      t = mkT(TK_ANY, TS_ENEMY);
      if (frontmost(g, t, pts) > 0) deal_damage_pt(g, pts[0], p[0], 1);
      break;
    case SBC_B004:  // cards/b004.py:13-22
      t = mkT(TK_ANY, TS_ENEMY); t.base = 1;
      n = get_targets(g, CUR(g), t, PT_NONE, pts);
      #pragma unroll 1
      for (int i = 0; i < n; i++) { deal_damage_pt(g, pts[i], p[0], 1); if (g.err) return; }
      destroy(g, id, 1);
      break;
    case SBC_B005: {  // cards/b005.py:15-33, including the memories of remembered temple copies (deepcopy)
      t = mkT(TK_ANY, TS_FRIENDLY);
      n = surrounding(g, ex, ey, CUR(g), &t, pts);
      int mine = 0;
      #pragma unroll 1
      for (int i = 0; i < g.n_mem; i++) if (g.mem[i].parent < 0 && g.mem[i].b005 == id) mine++;
      if (mine == 0) {
        #pragma unroll 1
        for (int i = 0; i < n; i++) {
          tid = need(g, pts[i]);
          if (tid < 0) return;
          int r = mem_push_entity(g, id, -1, pts[i], g.e[tid]);
          if (r < 0) return;
          if (g.e[tid].card == SBC_B005) {  // the copy carries a deep copy of that temple's own memories
            const int nm0 = g.n_mem;
            #pragma unroll 1
            for (int q = 0; q < nm0; q++)
              if (g.mem[q].parent < 0 && g.mem[q].b005 == tid && mem_copy_subtree(g, q, r, nm0) < 0) return;
          }
        }
      } else {
        int count = 0;
        const int nm0 = g.n_mem;
        #pragma unroll 1
        for (int i = 0; i < nm0 && count < p[0]; i++) {
          const Mem m = g.mem[i];
          if (m.parent >= 0 || m.b005 != id) continue;
          int occ = at_pt(g, m.pos);
          if (occ < 0 || (g.e[occ].card == m.card && ((g.e[occ].fl ^ m.fl) & (EF_OWNER | EF_STRUCT)) == 0)) {
            if (m.fl & EF_SINGLE) { GERR(g, SB_ERR_UNSUPPORTED); return; }  // detached copy: not modelled
            int c = new_ent(g, m.card, m.fl & EF_OWNER, m.strength);
            g.e[c].fl = (u8)((g.e[c].fl & ~EF_FIXED) | (m.fl & EF_FIXED));
            #pragma unroll 1
            for (int k = 0; k < 5; k++) g.e[c].st[k] = m.st[k];
            set_xy(g, PTX(m.pos), PTY(m.pos), c);
            #pragma unroll 1
            for (int q = i + 1; q < nm0; q++)  // the restored object keeps its own ability_remembered
              if (g.mem[q].parent == i) { g.mem[q].parent = -1; g.mem[q].b005 = (i8)c; }
            count++;
          }
        }
        mem_delete_temple(g, id);
      }
      break; }
    case SBC_B006: {  // cards/b006.py:14-39: ability_strength, ability_targets
      t = mkT(TK_UNIT, TS_FRIENDLY);
      n = get_targets(g, CUR(g), t, PT_NONE, pts);
      i8 sel[22]; int ns = 0;
      #pragma unroll 1
      for (int i = 0; i < n; i++) if (!g.e[at_pt(g, pts[i])].st[SB_ST_VITALIZED]) sel[ns++] = pts[i];
      shuffle(g, sel, ns);
      #pragma unroll 1
      for (int i = 0; i < ns && i < p[1]; i++) { tid = need(g, sel[i]); if (tid < 0) return; v_vitalize(g, tid); }
      i8 tiles[2], fr[5], bh[5]; int nt = 0;
      int nf = column_tiles(g, ex, ey, CUR(g), nullptr, true, fr);
      int nb = column_tiles(g, ex, ey, CUR(g), nullptr, false, bh);
      if (nf > 0 && at_pt(g, fr[0]) < 0 && within_front_line(g, me, PTY(fr[0]))) tiles[nt++] = fr[0];
      if (nb > 0 && at_pt(g, bh[0]) < 0) tiles[nt++] = bh[0];
      if (nt > 0) {
        int c = new_ent(g, e.card, me, p[0]);
        int where = choice_pt(g, tiles, nt);
        struct_play(g, c, PTX(where), PTY(where));
      }
      break; }
    case SBC_B007: {  // cards/b007.py:12-19
      int opp = opponent_of(g, me);
      if (g.pl[me].base == g.pl[opp].base) return;
      int stronger = g.pl[me].base > g.pl[opp].base ? me : opp;
      player_damage(g, stronger, p[0]);
      g.pl[opponent_of(g, stronger)].base += p[0];
      break; }
    case SBC_B008: {  // cards/b008.py:14-26
      Ply& pl = g.pl[me];
      if (pl.n_hand > 0 && CARD(g, pl.hand[0].card).kind == KIND_UNIT) pl.hand[0].flags ^= SB_CF_FIXED;
      t = mkT(TK_UNIT, TS_ANY); t.status = 1 << SB_ST_CONFUSED;
      n = get_targets(g, CUR(g), t, PT_NONE, pts);
      if (n > 0) {
        i16 ks[22];
        #pragma unroll 1
        for (int i = 0; i < n; i++) ks[i] = g.e[at_pt(g, pts[i])].strength;
        keyed_sort(g, pts, ks, n, false);
        tid = need(g, pts[0]);
        if (tid >= 0) destroy(g, tid, 1);
      }
      break; }
    case SBC_B009:  // cards/b009.py:13-24
      t = mkT(TK_UNIT, TS_ENEMY);
      n = frontmost(g, t, pts);
      if (n > 0) {
        if (n > p[0]) n = p[0];
        shuffle(g, pts, n);
        #pragma unroll 1
        for (int i = 0; i < n; i++) { tid = need(g, pts[i]); if (tid < 0) return; v_confuse(g, tid); }
      }
      break;
    case SBC_B104:  // cards/b104.py:12-20
      t = mkT(TK_UNIT, TS_ENEMY);
      if (frontmost(g, t, pts) > 0) { tid = need(g, pts[0]); if (tid >= 0) v_freeze(g, tid); }
      break;
    case SBC_B203:  // cards/b203.py:12-21
      t = mkT(TK_UNIT, TS_FRIENDLY);
      n = column_tiles(g, ex, ey, me, &t, true, pts);
      #pragma unroll 1
      for (int i = 0; i < n; i++) {
        tid = need(g, pts[i]);
        if (tid < 0) return;
        if (ent_struct(g.e[tid])) { GERR(g, SB_ERR_NONE_TARGET); return; }
        if (g.e[tid].st[SB_ST_CONFUSED]) st_remove(g, tid, SB_ST_CONFUSED);
        v_command(g, tid);
      }
      break;
    case SBC_B304:  // cards/b304.py:12-13
      deal_damage(g, id, p[0], 0, 1);
      break;
    case SBC_B305: {  // cards/b305.py:16-45: ability_amount, ability_mana, original_cost
      t = mkT(TK_STRUCTURE, TS_FRIENDLY);
      n = get_targets(g, CUR(g), t, PT(ex, ey), pts);
      #pragma unroll 1
      for (int i = 0; i < n && i < p[0]; i++) {
        tid = need(g, pts[i]);
        if (tid < 0) return;
        if (g.e[tid].card == e.card) {
          Target tu = mkT(TK_UNIT, TS_ANY);
          i8 sp[22];
          int tx = PTX(pts[i]), ty = PTY(pts[i]);
          int ns = surrounding(g, tx, ty, CUR(g), &tu, sp);
          #pragma unroll 1
          for (int k = 0; k < ns; k++) {
            int nx = PTX(sp[k]) - tx + ex, ny = PTY(sp[k]) - ty + ey;
            if (valid_xy(nx, ny)) { int u = need(g, sp[k]); if (u < 0) return; v_teleport(g, u, nx, ny); }
          }
          destroy(g, tid, 1);
          Ply& pl = g.pl[me];
          if (pl.n_deck == 0) { GERR(g, SB_ERR_INDEX); return; }
          pl.deck[pl.n_deck - 1].cost = p[2];
          return;
        }
      }
      {  // no other temple: the BOARD INSTANCE itself goes to the hand with cost 2 (a live link)
        Ply& pl = g.pl[me];
        bool single = (e.fl & EF_SINGLE) != 0;
        if (!single) { if (pl.n_deck == 0) { GERR(g, SB_ERR_INDEX); return; } pl.n_deck--; }
        if (pl.n_hand >= HAND_W) { GERR(g, SB_ERR_OVERFLOW); return; }
        CardRec r; r.card = e.card; r.cost = p[1]; r.flags = (u8)((single ? SB_CF_SINGLE_USE : 0) | SB_CF_OBJ); r.link = (i8)id; r.wn = 0; r.xstr = 0;
        pl.hand[pl.n_hand++] = r;
        g.n_obj++;
      }
      break; }
    // ------------------------------------------------------------ units
    case SBC_U007:  // cards/u007.py:13-21
      t = mkT(TK_UNIT, TS_ENEMY);
      n = surrounding(g, ex, ey, me, &t, pts);
      if (n > 0) { tid = need(g, choice_pt(g, pts, n)); if (tid < 0) return; v_heal(g, tid, p[0]); v_vitalize(g, tid); }
      break;
    case SBC_U017: {  // cards/u017.py:18-34
      Ply& pl = g.pl[me];
      i8 cand[HAND_W]; int nc = 0;
      #pragma unroll 1
      for (int i = 0; i < pl.n_hand; i++) if (CARD(g, pl.hand[i].card).kind == KIND_SPELL && pl.hand[i].cost <= 8) cand[nc++] = (i8)i;
      if (nc > 0) {
        shuffle(g, cand, nc);
        int remaining = 8, nch = 0;
        i8 chosen[HAND_W];
        #pragma unroll 1
        for (int i = 0; i < nc; i++) if (pl.hand[cand[i]].cost <= remaining) { chosen[nch++] = cand[i]; remaining -= pl.hand[cand[i]].cost; }
        #pragma unroll 1
        for (int i = 0; i < nch; i++) {
          const DCard& c = CARD(g, pl.hand[chosen[i]].card);
          int where = PT_NONE;
          if (c.flags & DCF_TARGET) {
            Target rt = card_target(c);
            n = get_targets(g, CUR(g), rt, PT_NONE, pts);
            where = choice_pt(g, pts, n);
            if (g.err) return;
          }
          int idx = chosen[i];
          player_play(g, me, idx, where);
          if (g.err) return;
          #pragma unroll 1
          for (int k = i + 1; k < nch; k++) if (chosen[k] > idx) chosen[k]--;
        }
      }
      break; }
    case SBC_U018: {  // cards/u018.py:13-26
      t = mkT(TK_UNIT, TS_ANY);
      n = surrounding(g, ex, ey, CUR(g), &t, pts);
      u32 m = 0;
      #pragma unroll 1
      for (int i = 0; i < n; i++) m |= 1u << CARD(g, g.e[at_pt(g, pts[i])].card).first_type;
      int cnt = __popc(m);
      #pragma unroll 1
      for (int k = 0; k < cnt; k++) {
        Target tb = mkT(TK_ANY, TS_ENEMY); tb.base = 1;
        n = get_targets(g, CUR(g), tb, PT_NONE, pts);
        int where = choice_pt(g, pts, n);
        if (g.err) return;
        deal_damage_pt(g, where, p[0], 1);
        if (g.err) return;
      }
      break; }
    case SBC_U021:  // cards/u021.py:13-19
      t = mkT(TK_UNIT, TS_FRIENDLY);
      n = get_targets(g, CUR(g), t, PT(ex, ey), pts);
      if (n > 0) { tid = need(g, choice_pt(g, pts, n)); if (tid >= 0) v_heal(g, tid, p[0]); }
      break;
    case SBC_U026:  // cards/u026.py:13-16
      t = mkT(TK_ANY, TS_ENEMY);
      n = column_tiles(g, ex, ey, CUR(g), &t, false, pts);
      if (n > 0) deal_damage_pt(g, pts[0], p[0], 1);
      break;
    case SBC_U036:  // cards/u036.py:12-14
      if (g.pl[me].n_hand == 0) { tid = need(g, PT(ex, ey)); if (tid >= 0) v_heal(g, tid, p[0]); }
      break;
    case SBC_U040:  // cards/u040.py:13-22 (the print is dropped); respawn unit.py:384-402
      if (has_source) {
        n = surrounding(g, ex, ey, me, nullptr, pts);
        if (n > 0) {
          int where = choice_pt(g, pts, n);
          int c = new_ent(g, e.card, me, p[0]);
          set_xy(g, PTX(where), PTY(where), c);
        }
      }
      break;
    case SBC_U050:  // cards/u050.py:13-15
      if (ey == 4) gain_speed(g, id, p[0]);
      break;
    case SBC_U051:  // cards/u051.py:14-21: ability_movement, ability_strength
      t = mkT(TK_UNIT, TS_ANY);
      if (bordering(g, ex, ey, CUR(g), &t, pts) == 0) gain_speed(g, id, p[0]);
      else v_heal(g, id, p[1]);
      break;
    case SBC_U053:  // cards/u053.py:14-24: ability_amount=1, ability_movement=2
      t = mkT(TK_UNIT, TS_ANY);
      if (surrounding(g, ex, ey, CUR(g), &t, pts) == 0) gain_speed(g, id, p[1]);
      else if (bordering(g, ex, ey, CUR(g), &t, pts) == 0) gain_speed(g, id, p[0]);
      break;
    case SBC_U055:  // cards/u055.py:12-19
      t = mkT(TK_UNIT, TS_ENEMY); t.xstatus = 1 << SB_ST_CONFUSED;
      n = column_tiles(g, ex, ey, CUR(g), &t, true, pts);
      #pragma unroll 1
      for (int i = 0; i < n; i++) { tid = need(g, pts[i]); if (tid < 0) return; v_confuse(g, tid); }
      break;
    case SBC_U061:  // cards/u061.py:12-23
      v_confuse(g, id);
      t = mkT(TK_UNIT, TS_FRIENDLY);
      n = get_targets(g, CUR(g), t, PT(ex, ey), pts);
      if (n > 0) { tid = need(g, choice_pt(g, pts, n)); if (tid < 0) return; v_confuse(g, tid); }
      gain_speed(g, id, 2);
      break;
    case SBC_U071: {  // cards/u071.py:12-27
      t = mkT(TK_UNIT, TS_ENEMY);
      n = bordering(g, ex, ey, me, &t, pts);
      i8 nc[4]; int nn = 0;
      #pragma unroll 1
      for (int i = 0; i < n; i++) if (!g.e[at_pt(g, pts[i])].st[SB_ST_CONFUSED]) nc[nn++] = pts[i];
      if (nn > 0) {
        tid = need(g, choice_pt(g, nc, nn));
        if (tid < 0) return;
        v_confuse(g, tid);
        i8 fr[5];
        int nf = column_tiles(g, ex, ey, CUR(g), nullptr, true, fr);
        if (nf > 0 && at_pt(g, fr[0]) < 0) v_teleport(g, id, PTX(fr[0]), PTY(fr[0]));
      }
      break; }
    case SBC_U074:  // cards/u074.py:12-19
      t = mkT(TK_UNIT, TS_ENEMY);
      n = column_tiles(g, ex, ey, me, &t, true, pts);
      if (n > 0) v_force_attack(g, id, PTX(pts[0]), PTY(pts[0]));
      break;
    case SBC_U076:  // cards/u076.py:14-26: ability_damage, ability_strength
      t = mkT(TK_UNIT, TS_ANY); t.xtypes = 1 << UT_DRAGON;
      n = surrounding(g, ex, ey, CUR(g), &t, pts);
      if (n > 0) {
        tid = need(g, choice_pt(g, pts, n));
        if (tid < 0) return;
        deal_damage(g, tid, p[0], 0, 1);
        if (g.e[tid].strength <= 0) spawn_token_unit(g, me, PT(g.e[tid].x, g.e[tid].y), p[1], UT_DRAGON);
      }
      break;
    case SBC_U101:  // cards/u101.py:13-26
      if (pos_pt < 0 || pos_pt >= 20) return;
      tid = at_pt(g, pos_pt);
      if (tid < 0 || ent_struct(g.e[tid]) || !g.e[tid].st[SB_ST_FROZEN]) return;
      t = mkT(TK_UNIT, TS_ENEMY); t.status = 1 << SB_ST_FROZEN;
      n = surrounding(g, ex, ey, me, &t, pts);
      #pragma unroll 1
      for (int i = 0; i < n; i++) { deal_damage_pt(g, pts[i], p[0], 1); if (g.err) return; }
      break;
    case SBC_U103:  // cards/u103.py:12-18
      t = mkT(TK_UNIT, TS_ENEMY);
      n = bordering(g, ex, ey, CUR(g), &t, pts);
      #pragma unroll 1
      for (int i = 0; i < n; i++) { tid = need(g, pts[i]); if (tid < 0) return; v_freeze(g, tid); }
      break;
    case SBC_U106:  // cards/u106.py:13-18
      t = mkT(TK_STRUCTURE, TS_FRIENDLY);
      if (bordering(g, ex, ey, CUR(g), &t, pts) > 0 || ey == 4) v_heal(g, id, p[0]);
      break;
    case SBC_U111:  // cards/u111.py:13-21
      #pragma unroll 1
      for (int k = 0; k < p[0]; k++) {
        t = mkT(TK_UNIT, TS_FRIENDLY);
        n = surrounding(g, ex, ey, me, &t, pts);
        if (n > 0) { tid = need(g, choice_pt(g, pts, n)); if (tid < 0) return; v_heal(g, tid, 1); }
      }
      break;
    case SBC_U117: v_freeze(g, id); break;          // cards/u117.py:11-12
    case SBC_U206: player_damage(g, me, p[0]); break;  // cards/u206.py:12-13
    case SBC_U211:  // cards/u211.py:14-19: ability_max_strength, ability_min_strength
      n = column_tiles(g, ex, ey, CUR(g), nullptr, true, pts);
      if (n > 0 && at_pt(g, pts[0]) < 0) {
        int s = p[1] + rng_below(g, p[0] + 1 - p[1]);
        spawn_token_unit(g, me, pts[0], s, UT_SATYR);
      }
      break;
    case SBC_U216: player_damage(g, me, p[0]); break;  // cards/u216.py:12-13
    case SBC_U217: {  // cards/u217.py:13-17
      i8 row[4]; int nr = 0;
      #pragma unroll 1
      for (int x = 0; x < 4; x++) if (g.board[16 + x] < 0) row[nr++] = (i8)(16 + x);
      if (nr > 0) spawn_token_unit(g, me, choice_pt(g, row, nr), p[0], UT_SATYR);
      break; }
    case SBC_U302:  // cards/u302.py:12-23
      if (pos_pt < 0 || pos_pt >= 20) return;
      tid = at_pt(g, pos_pt);
      if (tid < 0 || ent_struct(g.e[tid])) return;
      if (g.e[tid].strength > e.strength) {
        deal_damage(g, tid, p[0], 0, 1);
        if (g.e[tid].strength > 0) v_push(g, tid, e.x, e.y);
      }
      break;
    case SBC_U305:  // cards/u305.py:13-21
      t = mkT(TK_UNIT, TS_FRIENDLY); t.types = 1 << UT_CONSTRUCT;
      n = bordering(g, ex, ey, CUR(g), &t, pts);
      if (n > 0) { tid = need(g, choice_pt(g, pts, n)); if (tid < 0) return; v_heal(g, tid, p[0]); v_heal(g, id, p[0]); }
      break;
    case SBC_U306:  // cards/u306.py:13-20
      t = mkT(TK_ANY, TS_FRIENDLY);
      n = get_targets(g, CUR(g), t, PT(ex, ey), pts);
      if (n > 0) deal_damage_pt(g, choice_pt(g, pts, n), p[0], 1);
      break;
    case SBC_U310: {  // cards/u310.py:12-41
      t = mkT(TK_UNIT, TS_ENEMY);
      n = bordering(g, ex, ey, CUR(g), &t, pts);
      int behind = PT_NONE, right = PT_NONE, left = PT_NONE, front = PT_NONE;
      #pragma unroll 1
      for (int i = n - 1; i >= 0; i--) {
        if (PTY(pts[i]) == ey + 1) behind = pts[i];
        if (PTX(pts[i]) == ex + 1) right = pts[i];
        if (PTX(pts[i]) == ex - 1) left = pts[i];
        if (PTY(pts[i]) == ey - 1) front = pts[i];
      }
      int target = PT_NONE;
      if (behind != PT_NONE && PTY(behind) < 4 && at_xy(g, PTX(behind), PTY(behind) + 1) < 0) target = behind;
      else if (left != PT_NONE && PTX(left) > 0 && at_xy(g, PTX(left) - 1, PTY(left)) < 0) target = left;
      else if (right != PT_NONE && PTX(right) < 3 && at_xy(g, PTX(right) + 1, PTY(right)) < 0) target = right;
      else if (front != PT_NONE && PTY(front) > 0 && at_xy(g, PTX(front), PTY(front) - 1) < 0) target = front;
      if (target == PT_NONE) { GERR(g, SB_ERR_INDEX); return; }  // UnboundLocalError
      tid = need(g, target);
      if (tid >= 0) v_push(g, tid, ex, ey);
      break; }
    case SBC_U313:  // cards/u313.py:13-21
      t = mkT(TK_UNIT, TS_FRIENDLY);
      n = surrounding(g, ex, ey, CUR(g), &t, pts);
      if (n > 0) { tid = need(g, choice_pt(g, pts, n)); if (tid < 0) return; v_heal(g, tid, p[0]); }
      v_heal(g, id, p[0]);
      break;
    case SBC_U314:  // cards/u314.py:11-17
      t = mkT(TK_UNIT, TS_FRIENDLY);
      n = column_tiles(g, ex, ey, CUR(g), &t, true, pts);
      if (n > 0) { tid = need(g, pts[0]); if (tid >= 0) v_push(g, tid, ex, ey); }
      break;
    case SBC_U316:  // cards/u316.py:13-23
      t = mkT(TK_UNIT, TS_FRIENDLY);
      n = surrounding(g, ex, ey, CUR(g), &t, pts);
      if (n > 0) {
        shuffle(g, pts, n);
        #pragma unroll 1
        for (int i = 0; i < n && i < p[0]; i++) { tid = need(g, pts[i]); if (tid < 0) return; v_vitalize(g, tid); }
      }
      v_vitalize(g, id);
      break;
    case SBC_U320:  // cards/u320.py:13-19
      t = mkT(TK_UNIT, TS_FRIENDLY);
      n = surrounding(g, ex, ey, me, &t, pts);
      if (n > 0) { tid = need(g, choice_pt(g, pts, n)); if (tid >= 0) v_heal(g, tid, p[0]); }
      break;
    case SBC_U401:  // cards/u401.py:13-22: `damage`
      t = mkT(TK_UNIT, TS_ANY);
      n = bordering(g, ex, ey, CUR(g), &t, pts);
      #pragma unroll 1
      for (int i = 0; i < n; i++) {
        tid = at_pt(g, pts[i]);
        if (tid >= 0) {
          deal_damage(g, tid, p[0], 0, 1);
          if (ent_struct(g.e[tid])) { GERR(g, SB_ERR_NONE_TARGET); return; }
          v_poison(g, tid);
        }
      }
      break;
    case SBC_U403:  // cards/u403.py:14-24: ability_amount, ability_strength
      t = mkT(TK_UNIT, TS_ANY); t.status = 1 << SB_ST_POISONED;
      n = surrounding(g, ex, ey, CUR(g), &t, pts);
      #pragma unroll 1
      for (int i = 0; i < n; i++) {
        i8 bt[4], em[4];
        int nb = bordering(g, PTX(pts[i]), PTY(pts[i]), CUR(g), nullptr, bt);
        int ne = empty_of(g, bt, nb, em);
        shuffle(g, em, ne);
        #pragma unroll 1
        for (int k = 0; k < ne && k < p[0]; k++) spawn_token_unit(g, me, em[k], p[1], UT_TOAD);
      }
      break;
    case SBC_U405:  // cards/u405.py:13-21
      t = mkT(TK_UNIT, TS_ANY);
      n = bordering(g, ex, ey, CUR(g), &t, pts);
      #pragma unroll 1
      for (int i = 0; i < n; i++) {
        int dealt = deal_damage_pt(g, pts[i], p[0], 1);
        if (g.err) return;
        v_heal(g, id, dealt);
      }
      break;
    case SBC_U406: {  // cards/u406.py:13-20
      i8 bt[4], em[4];
      int nb = bordering(g, ex, ey, CUR(g), nullptr, bt);
      int ne = empty_of(g, bt, nb, em);
      if (ne > 0) { int where = choice_pt(g, em, ne); spawn_token_unit(g, opponent_of(g, me), where, p[0], UT_RAVEN); }
      break; }
    case SBC_U411:  // cards/u411.py:13-22
      t = mkT(TK_UNIT, TS_ENEMY);
      n = get_targets(g, CUR(g), t, PT_NONE, pts);
      if (n > 0) {
        shuffle(g, pts, n);
        #pragma unroll 1
        for (int i = 0; i < n && i < p[0]; i++) { tid = need(g, pts[i]); if (tid < 0) return; v_poison(g, tid); }
      }
      break;
    case SBC_UA03: {  // cards/ua03.py:12-15
      i8 row[5]; int nr = 0;
      #pragma unroll 1
      for (int x = 0; x < 4; x++) if (g.board[ey * 4 + x] < 0) row[nr++] = (i8)PT(x, ey);
      row[nr++] = (i8)PT(ex, ey);
      int where = choice_pt(g, row, nr);
      v_teleport(g, id, PTX(where), PTY(where));
      break; }
    case SBC_UA04: {  // cards/ua04.py:12-33
      i8 lp[22], rp[22];
      Target tf = mkT(TK_UNIT, TS_FRIENDLY), te = mkT(TK_UNIT, TS_ENEMY);
      int nl = get_targets(g, CUR(g), tf, PT_NONE, lp);
      int nr = get_targets(g, CUR(g), te, PT_NONE, rp);
      const i8* src = nullptr; int ns = 0;
      if (nl > nr) { src = lp; ns = nl; } else if (nr > nl) { src = rp; ns = nr; }
      if (src) {
        i8 sel[22]; int nsel = 0;
        int mn = 32767;
        #pragma unroll 1
        for (int i = 0; i < ns; i++) { int s = g.e[at_pt(g, src[i])].strength; if (s < mn) mn = s; }
        #pragma unroll 1
        for (int i = 0; i < ns; i++) if (g.e[at_pt(g, src[i])].strength == mn) sel[nsel++] = src[i];
        if (nsel > 0) destroy(g, at_pt(g, sel[rng_below(g, nsel)]), 1);
      }
      break; }
    case SBC_UA05: {  // cards/ua05.py:13-19
      i8 sd[2], em[2];
      int ns = side_list(ex, ey, sd);
      int ne = empty_of(g, sd, ns, em);
      #pragma unroll 1
      for (int k = 0; k < ne; k++) spawn_token_unit(g, me, em[k], p[0], UT_ANCIENT);
      break; }
    case SBC_UA07:  // cards/ua07.py:11-22
      switch (rng_below(g, 5)) {
        case 0: v_freeze(g, id); break;
        case 1: v_poison(g, id); break;
        case 2: v_vitalize(g, id); break;
        case 3: v_confuse(g, id); break;
        case 4: v_disable(g, id); break;
      }
      break;
    case SBC_UA20:  // cards/ua20.py:21-32: ability_cost, ability_level
      t = mkT(TK_UNIT, TS_ENEMY);
      if (column_tiles(g, ex, ey, me, &t, true, pts) == 0) {
        Ply& pl = g.pl[me];
        int k = rng_below(g, 4);
        int c = k == 0 ? SBC_B005 : k == 1 ? SBC_B006 : k == 2 ? SBC_B203 : SBC_B305;
        if (pl.n_deck >= DECK_W) { GERR(g, SB_ERR_OVERFLOW); return; }
        CardRec r; r.card = (u8)c; r.cost = p[0]; r.flags = SB_CF_SINGLE_USE; r.link = -1; r.wn = 0; r.xstr = 0;
        pl.deck[pl.n_deck++] = r;
      }
      break;
    case SBC_UD01:  // cards/ud01.py:13-19
      t = mkT(TK_UNIT, TS_FRIENDLY); t.types = 1 << UT_DRAGON;
      n = get_targets(g, me, t, PT_NONE, pts);
      if (n > 0) { tid = need(g, choice_pt(g, pts, n)); if (tid >= 0) v_heal(g, tid, p[0]); }
      break;
    case SBC_UD02:  // cards/ud02.py:13-19
      t = mkT(TK_UNIT, TS_ANY); t.xtypes = 1 << UT_DRAGON;
      n = column_tiles(g, ex, ey, CUR(g), &t, true, pts);
      #pragma unroll 1
      for (int i = 0; i < n; i++) { deal_damage_pt(g, pts[i], p[0], 1); if (g.err) return; }
      break;
    case SBC_UD31:  // cards/ud31.py:13-24
      if (pos_pt < 0 || pos_pt >= 20) return;
      tid = at_pt(g, pos_pt);
      if (tid < 0 || ent_struct(g.e[tid])) return;
      t = mkT(TK_UNIT, TS_FRIENDLY); t.types = 1 << UT_DRAGON;
      n = surrounding(g, ex, ey, me, &t, pts);
      if (n > 0) { tid = need(g, choice_pt(g, pts, n)); if (tid < 0) return; v_heal(g, tid, p[0]); }
      v_heal(g, id, p[0]);
      break;
    case SBC_UE01: {  // cards/ue01.py:11-19
      int times = e.dmg;
      #pragma unroll 1
      for (int k = 0; k < times; k++) {
        t = mkT(TK_UNIT, TS_ENEMY);
        n = get_targets(g, me, t, PT_NONE, pts);
        if (n > 0) { deal_damage_pt(g, choice_pt(g, pts, n), 1, 1); if (g.err) return; }
      }
      break; }
    case SBC_UE03: {  // cards/ue03.py:11-17
      i8 bt[4], em[4];
      int nb = bordering(g, ex, ey, CUR(g), nullptr, bt);
      int ne = empty_of(g, bt, nb, em);
      if (ne > 0) spawn_token_unit(g, me, choice_pt(g, em, ne), e.strength, UT_ELDER);
      break; }
    case SBC_UE04: {  // cards/ue04.py:12-18
      t = mkT(TK_UNIT, TS_ENEMY);
      n = get_targets(g, me, t, PT_NONE, pts);
      int c = 0;
      #pragma unroll 1
      for (int i = 0; i < n; i++) if (g.e[at_pt(g, pts[i])].strength > e.strength) c++;
      v_heal(g, id, c * p[0]);
      break; }
    case SBC_UE05:  // cards/ue05.py:12-19
      t = mkT(TK_UNIT, TS_FRIENDLY); t.has_limit = 1; t.limit = (i16)(e.strength - 1);
      n = get_targets(g, me, t, PT(ex, ey), pts);
      #pragma unroll 1
      for (int i = 0; i < n && i < p[0]; i++) g.e[at_pt(g, pts[i])].strength = e.strength;
      break;
    case SBC_UE11: v_heal(g, id, p[0]); break;  // cards/ue11.py:12-13
    case SBC_UE12:  // cards/ue12.py:11-18
      t = mkT(TK_UNIT, TS_ENEMY);
      n = column_tiles(g, ex, ey, me, &t, true, pts);
      #pragma unroll 1
      for (int i = 0; i < n; i++) { tid = need(g, pts[i]); if (tid < 0) return; destroy(g, tid, 1); }
      break;
    case SBC_UE21:  // cards/ue21.py:11-18
      t = mkT(TK_UNIT, TS_FRIENDLY); t.has_limit = 1; t.limit = e.strength;
      n = get_targets(g, me, t, PT(ex, ey), pts);
      #pragma unroll 1
      for (int i = 0; i < n; i++) {
        tid = need(g, pts[i]);
        if (tid < 0) return;
        if (ent_struct(g.e[tid])) { GERR(g, SB_ERR_NONE_TARGET); return; }
        v_command(g, tid);
      }
      break;
    case SBC_UE22:  // cards/ue22.py:12-22
      t = mkT(TK_UNIT, TS_FRIENDLY);
      n = get_targets(g, me, t, PT(ex, ey), pts);
      if (n > 0) {
        shuffle(g, pts, n);
        #pragma unroll 1
        for (int i = 0; i < n && i < p[0]; i++) { tid = need(g, pts[i]); if (tid < 0) return; v_heal(g, tid, e.dmg); }
      }
      break;
    case SBC_UE31: e.strength = p[0]; break;  // cards/ue31.py:12-13
    case SBC_UE32: {  // cards/ue32.py:12-22
      int damage = e.strength < 6 ? e.strength : 6;
      t = mkT(TK_ANY, TS_ENEMY);
      n = column_tiles(g, ex, ey, me, &t, true, pts);
      if (n > 0) deal_damage_pt(g, pts[0], damage, 1);
      else player_damage(g, opponent_of(g, me), damage);
      break; }
    case SBC_UE41: v_convert(g, id); break;  // cards/ue41.py:10-11
    case SBC_UE42: {  // cards/ue42.py:12-15
      int amount = e.dmg < p[0] ? e.dmg : p[0];
      v_heal(g, id, player_damage(g, opponent_of(g, me), amount));
      break; }
    case SBC_UP02: v_heal(g, id, p[0] * count_types_friendly(g)); break;  // cards/up02.py:12-20
    case SBC_UP03: {  // cards/up03.py:12-28
      int c = count_types_friendly(g);
      t = mkT(TK_UNIT, TS_ENEMY);
      n = surrounding(g, ex, ey, CUR(g), &t, pts);
      #pragma unroll 1
      for (int i = 0; i < n; i++) {
        tid = need(g, pts[i]);
        if (tid < 0) return;
        int s = g.e[tid].strength - p[0] * c;
        g.e[tid].strength = (i16)(s > 1 ? s : 1);  // unit.py:233-234 reduce
      }
      break; }
    default: break;
  }
}

SBD_NI void spell_effect(G& g, int card, int caster, int pos_pt) {
  G_LOCAL(g);
  const i8* p = CARD(g, card).p;
  i8 pts[22];
  int n, tid;
  Target t;
  switch (card) {
    case SBC_S001: deal_damage_pt(g, pos_pt, p[0], 1); break;  // cards/s001.py:13-14
    case SBC_S003:  // cards/s003.py:14-19: ability_max_damage, ability_min_damage
      t = mkT(TK_ANY, TS_ENEMY);
      n = get_targets(g, CUR(g), t, PT_NONE, pts);
      #pragma unroll 1
      for (int i = 0; i < n; i++) {
        if (need(g, pts[i]) < 0) return;
        deal_damage_pt(g, pts[i], p[1] + rng_below(g, p[0] + 1 - p[1]), 1);
        if (g.err) return;
      }
      break;
    case SBC_S004: {  // cards/s004.py:14-23: ability_max_amount, ability_min_amount
      i8 tl[20], em[20];
      int nt = within_front_line_tiles(g, caster, tl);
      int ne = empty_of(g, tl, nt, em);
      if (ne > 0) {
        shuffle(g, em, ne);
        int amount = p[1] + rng_below(g, p[0] + 1 - p[1]);
        #pragma unroll 1
        for (int i = 0; i < ne && i < amount; i++) spawn_token_unit(g, caster, em[i], 1, UT_TOAD);
      }
      break; }
    case SBC_S007:  // cards/s007.py:13-16
      tid = need(g, pos_pt);
      if (tid < 0) return;
      v_heal(g, tid, p[0]); v_vitalize(g, tid);
      break;
    case SBC_S012: {  // cards/s012.py:13-18
      i8 tl[20], em[20];
      int nt = within_front_line_tiles(g, caster, tl);
      int ne = empty_of(g, tl, nt, em);
      if (ne > 0) spawn_token_unit(g, caster, choice_pt(g, em, ne), p[0], UT_KNIGHT);
      break; }
    case SBC_S013: {  // cards/s013.py:13-27
      i8 chosen[16]; int nc = 0;
      u32 taken = 0;
      #pragma unroll 1
      for (int ut = 0; ut < 16; ut++) {
        t = mkT(TK_UNIT, TS_ANY); t.types = (u16)(1u << ut);
        n = get_targets(g, CUR(g), t, PT_NONE, pts);
        i8 units[22]; int nu = 0;
        #pragma unroll 1
        for (int i = 0; i < n; i++) if (!(taken >> pts[i] & 1)) units[nu++] = pts[i];
        if (nu > 0) { int c = choice_pt(g, units, nu); chosen[nc++] = (i8)c; taken |= 1u << c; }
      }
      #pragma unroll 1
      for (int i = 0; i < nc; i++) { deal_damage_pt(g, chosen[i], p[0], 1); if (g.err) return; }
      break; }
    case SBC_S021:  // cards/s021.py:13-25
      tid = need(g, pos_pt);
      if (tid < 0) return;
      v_confuse(g, tid);
      t = mkT(TK_UNIT, TS_FRIENDLY); t.types = 1 << UT_FELINE;
      n = get_targets(g, CUR(g), t, PT_NONE, pts);
      if (n > 0) {
        int mn = 32767; i8 wk[22]; int nw = 0;
        #pragma unroll 1
        for (int i = 0; i < n; i++) { int s = g.e[at_pt(g, pts[i])].strength; if (s < mn) mn = s; }
        #pragma unroll 1
        for (int i = 0; i < n; i++) if (g.e[at_pt(g, pts[i])].strength == mn) wk[nw++] = pts[i];
        tid = need(g, choice_pt(g, wk, nw));
        if (tid >= 0) v_heal(g, tid, p[0]);
      }
      break;
    case SBC_S101: {  // cards/s101.py:14-22: ability_mana, ability_strength
      g.pl[caster].mana = (i16)(g.pl[caster].mana + p[0]);
      t = mkT(TK_UNIT, TS_FRIENDLY);
      n = get_targets(g, CUR(g), t, PT_NONE, pts);
      if (n == 0) { GERR(g, SB_ERR_INDEX); return; }
      i16 ks[22];
      #pragma unroll 1
      for (int i = 0; i < n; i++) ks[i] = g.e[at_pt(g, pts[i])].strength;
      keyed_sort(g, pts, ks, n, false);
      tid = need(g, pts[0]);
      if (tid >= 0) v_heal(g, tid, p[1]);
      break; }
    case SBC_S104:  // cards/s104.py:13-19
      tid = need(g, pos_pt);
      if (tid < 0) return;
      if (g.e[tid].st[SB_ST_FROZEN]) deal_damage(g, tid, p[0], 0, 1); else v_freeze(g, tid);
      break;
    case SBC_S105:  // cards/s105.py:13-14
      tid = need(g, pos_pt);
      if (tid >= 0) v_heal(g, tid, p[0]);
      break;
    case SBC_S203: {  // cards/s203.py:14-30; list(set(...)) order is str-hash dependent upstream (Q14): canonical first-occurrence order
      i8 fr[22], all[24]; int na = 0;
      u32 seen = 0;
      t = mkT(TK_UNIT, TS_FRIENDLY);
      int nf = get_targets(g, CUR(g), t, PT_NONE, fr);
      #pragma unroll 1
      for (int i = 0; i < nf; i++) {
        Target te = mkT(TK_ANY, TS_ENEMY); te.base = 1;
        n = surrounding(g, PTX(fr[i]), PTY(fr[i]), CUR(g), &te, pts);
        #pragma unroll 1
        for (int k = 0; k < n; k++) if (!(seen >> pts[k] & 1)) { seen |= 1u << pts[k]; all[na++] = pts[k]; }
      }
      #pragma unroll 1
      for (int i = 0; i < na; i++) { deal_damage_pt(g, all[i], p[0], 1); if (g.err) return; }
      break; }
    case SBC_S302:  // cards/s302.py:14-22: ability_damage, ability_targets
      t = mkT(TK_ANY, TS_ENEMY); t.base = 1;
      n = get_targets(g, CUR(g), t, PT_NONE, pts);
      shuffle(g, pts, n);
      #pragma unroll 1
      for (int i = 0; i < n && i < p[1]; i++) { deal_damage_pt(g, pts[i], p[0], 1); if (g.err) return; }
      break;
    case SBC_S403:  // cards/s403.py:13-22
      t = mkT(TK_UNIT, TS_FRIENDLY); t.status = 1 << SB_ST_POISONED;
      n = get_targets(g, CUR(g), t, PT_NONE, pts);
      #pragma unroll 1
      for (int i = 0; i < n; i++) { tid = need(g, pts[i]); if (tid < 0) return; v_heal(g, tid, p[0]); v_vitalize(g, tid); }
      break;
    default: break;
  }
}
