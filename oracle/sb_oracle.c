/* TEST INFRASTRUCTURE -- see sb_oracle.h.  Rules engine restated from the reference, function by
 * function; every function cites the reference lines it follows (paths relative to /root/reference).
 * Compile with -ffp-contract=off: the weighted draw must round exactly like CPython floats.
 */
#include <string.h>
#include <stdlib.h>
#include "sb_oracle.h"

const OCard OCARDS[SBC_COUNT] = {
#include "sb_card_table.inc"
};

/* ------------------------------------------------------------------ weights (player.py:32,59) */
static double WT[1024];
static int wt_ready = 0;
static void wt_init(void) {
  if (wt_ready) return;
  double w = 1.0;
  for (int i = 0; i < 1024; i++) { WT[i] = w; w = w * 1.6 + 100.0; }
  wt_ready = 1;
}

/* ------------------------------------------------------------------ Philox stream (ref_harness.PhiloxRandomState) */
static void philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
static void rng_block(Game *g, uint32_t w[4]) {
  philox(g->draw, g->turn, 0, 0, g->seed_lo, g->seed_hi, w);
  g->draw++;
}
int o_rng_below(Game *g, int n) {
  if (n <= 0) { ERR(g, SB_ERR_EMPTY_CHOICE); return 0; }
  uint32_t w[4];
  rng_block(g, w);
  return (int)(((uint64_t)w[0] * (uint64_t)n) >> 32);
}
double o_rng_random(Game *g) {
  uint32_t w[4];
  rng_block(g, w);
  return ((double)(w[0] >> 5) * 67108864.0 + (double)(w[1] >> 6)) / 9007199254740992.0;
}
void o_shuffle(Game *g, int *a, int n) {
  for (int i = n - 1; i > 0; i--) {
    int j = o_rng_below(g, i + 1);
    int t = a[i]; a[i] = a[j]; a[j] = t;
  }
}
uint32_t sbo_agent_pick(uint64_t seed, uint32_t step, uint32_t n) {
  uint32_t w[4];
  philox(step, 0, 0xA6E7u, 0, (uint32_t)seed, (uint32_t)(seed >> 32), w);
  return (uint32_t)(((uint64_t)w[0] * n) >> 32);
}

/* ------------------------------------------------------------------ board access (board.py:58-71) */
Ply *o_local(Game *g) { return &g->pl[g->local_order]; }
Ply *o_remote(Game *g) { return &g->pl[1 - g->local_order]; }
/* player.py:42-44: `board.remote if order == FIRST else board.local` -- returns SELF after an odd number of flips (Q3) */
int o_opponent(const Game *g, int order) { return order == 0 ? 1 - g->local_order : g->local_order; }
int o_at(const Game *g, int x, int y) { return valid_xy(x, y) ? g->board[y][x] : -1; }
int o_at_pt(const Game *g, int pt) { return (pt >= 0 && pt < 20) ? g->board[pt >> 2][pt & 3] : -1; }
void o_set(Game *g, int x, int y, int id) {
  g->board[y][x] = id;
  if (id >= 0) { g->e[id].x = x; g->e[id].y = y; }
}
/* board.py:78-92 */
void o_calc_front_line(Game *g, int order) {
  if (order == g->local_order) {
    Ply *p = o_local(g);
    p->front_line = 4;
    for (int y = 0; y < 5; y++) {
      int any = 0;
      for (int x = 0; x < 4; x++) { int id = g->board[y][x]; if (id >= 0 && g->e[id].owner == order) any = 1; }
      if (any) { p->front_line = y > 1 ? y : 1; break; }
    }
  } else {
    Ply *p = o_remote(g);
    p->front_line = 0;
    for (int y = 4; y >= 0; y--) {
      int any = 0;
      for (int x = 0; x < 4; x++) { int id = g->board[y][x]; if (id >= 0 && g->e[id].owner == order) any = 1; }
      if (any) { p->front_line = y < 3 ? y : 3; break; }
    }
  }
}

/* ------------------------------------------------------------------ target queries (board.py:147-296) */
int o_get_targets(Game *g, int pov, const Target *t, int exclude_pt, int *out) {
  int n = 0;
  int pov_local = (pov == g->local_order);
  for (int yi = 0; yi < 5; yi++) {
    int y = pov_local ? yi : 4 - yi;
    for (int xi = 0; xi < 4; xi++) {
      int x = pov_local ? xi : 3 - xi;
      int id = g->board[y][x];
      if (id < 0) continue;
      const Ent *e = &g->e[id];
      if (e->strength <= 0) continue;
      int strength_ok = !t->has_limit || e->strength <= t->limit;
      int unit_ok = 0, struct_ok = 0;
      if (!e->is_struct) {
        int ok = 1;
        if (t->types && !(e->types & t->types)) ok = 0;
        if (t->xtypes && (e->types & t->xtypes)) ok = 0;
        if (t->nonhero && (e->types & (1 << UT_HERO))) ok = 0;
        if (t->status) {
          int any = 0;
          for (int s = 0; s < 5; s++) if ((t->status >> s & 1) && e->st[s] > 0) any = 1;
          if (!any) ok = 0;
        }
        if (t->xstatus) {
          for (int s = 0; s < 5; s++) if ((t->xstatus >> s & 1) && e->st[s] > 0) ok = 0;
        }
        unit_ok = ok && strength_ok;
      } else {
        struct_ok = strength_ok;
      }
      int kind_ok = (t->kind == TK_ANY && (unit_ok || struct_ok)) || (t->kind == TK_UNIT && unit_ok) ||
                    (t->kind == TK_STRUCTURE && struct_ok);
      int side_ok = t->side == TS_ANY || (t->side == TS_FRIENDLY && e->owner == pov) ||
                    (t->side == TS_ENEMY && e->owner != pov);
      if (kind_ok && side_ok) out[n++] = PT(x, y);
    }
  }
  if (t->base) {
    int friendly = pov_local ? PT_BASE_LOCAL : PT_BASE_REMOTE;
    int enemy = pov_local ? PT_BASE_REMOTE : PT_BASE_LOCAL;
    if (t->side == TS_FRIENDLY || t->side == TS_ANY) out[n++] = friendly;
    if (t->side == TS_ENEMY || t->side == TS_ANY) out[n++] = enemy;
  }
  if (exclude_pt != PT_NONE) {
    for (int i = 0; i < n; i++)
      if (out[i] == exclude_pt) { memmove(out + i, out + i + 1, (n - i - 1) * sizeof(int)); n--; break; }
  }
  return n;
}

static int in_list(const int *l, int n, int v) { for (int i = 0; i < n; i++) if (l[i] == v) return 1; return 0; }

/* keep `targets` order, membership in region or (base && include_base): board.py:215,230,262,276,294 */
static int filter_region(Game *g, int pov, const Target *t, const int *region, int nr, int base_passes, int *out) {
  int tg[24];
  int nt = o_get_targets(g, pov, t, PT_NONE, tg);
  int n = 0;
  for (int i = 0; i < nt; i++)
    if (in_list(region, nr, tg[i]) || (base_passes && is_base_pt(tg[i]) && t->base)) out[n++] = tg[i];
  return n;
}
static void sort_by_y(int *a, int n, int desc) { /* stable insertion sort on Point.y (board.py:217,232) */
  for (int i = 1; i < n; i++) {
    int v = a[i], j = i - 1;
    while (j >= 0 && (desc ? PTY(a[j]) < PTY(v) : PTY(a[j]) > PTY(v))) { a[j + 1] = a[j]; j--; }
    a[j + 1] = v;
  }
}
/* board.py:206-219 */
int o_front(Game *g, int x, int y, int pov, const Target *t, int *out) {
  int region[5], nr = 0, n;
  int pov_local = (pov == g->local_order);
  if (pov_local) for (int i = y - 1; i >= 0; i--) region[nr++] = PT(x, i);
  else for (int i = y + 1; i < 5; i++) region[nr++] = PT(x, i);
  if (t) n = filter_region(g, pov, t, region, nr, 1, out);
  else { memcpy(out, region, nr * sizeof(int)); n = nr; }
  sort_by_y(out, n, pov_local);
  return n;
}
/* board.py:221-234 */
int o_behind(Game *g, int x, int y, int pov, const Target *t, int *out) {
  int region[5], nr = 0, n;
  int pov_local = (pov == g->local_order);
  if (pov_local) for (int i = y + 1; i < 5; i++) region[nr++] = PT(x, i);
  else for (int i = y - 1; i >= 0; i--) region[nr++] = PT(x, i);
  if (t) n = filter_region(g, pov, t, region, nr, 1, out);
  else { memcpy(out, region, nr * sizeof(int)); n = nr; }
  sort_by_y(out, n, !pov_local); /* reverse = (pov == remote) */
  return n;
}
static int region_generic(Game *g, const int (*d)[2], int nd, int x, int y, int pov, const Target *t, int base_passes, int *out) {
  int region[8], nr = 0;
  for (int i = 0; i < nd; i++) {
    int xx = x + d[i][0], yy = y + d[i][1];
    if (valid_xy(xx, yy)) region[nr++] = PT(xx, yy);
  }
  if (t) return filter_region(g, pov, t, region, nr, base_passes, out);
  memcpy(out, region, nr * sizeof(int));
  return nr;
}
/* board.py:236-246 */
int o_side(Game *g, int x, int y, int pov, const Target *t, int *out) {
  static const int d[2][2] = {{-1, 0}, {1, 0}};
  return region_generic(g, d, 2, x, y, pov, t, 0, out);
}
/* board.py:248-255 */
int o_row(Game *g, int x, int y, int pov, const Target *t, int *out) {
  int region[4];
  (void)x;
  for (int i = 0; i < 4; i++) region[i] = PT(i, y);
  if (t) return filter_region(g, pov, t, region, 4, 0, out);
  memcpy(out, region, sizeof region);
  return 4;
}
/* board.py:266-278 */
int o_bordering(Game *g, int x, int y, int pov, const Target *t, int *out) {
  static const int d[4][2] = {{-1, 0}, {1, 0}, {0, -1}, {0, 1}};
  return region_generic(g, d, 4, x, y, pov, t, 1, out);
}
/* board.py:280-296 */
int o_surrounding(Game *g, int x, int y, int pov, const Target *t, int *out) {
  static const int d[8][2] = {{-1, 0}, {-1, -1}, {-1, 1}, {1, 0}, {1, -1}, {1, 1}, {0, -1}, {0, 1}};
  return region_generic(g, d, 8, x, y, pov, t, 1, out);
}

/* player.py:96-111 -- keyed on ORDER, not local/remote (Q22) */
int o_is_within_front_line(const Game *g, int order, int y) {
  return order == 0 ? y >= g->pl[order].front_line : y <= g->pl[order].front_line;
}
int o_get_within_front_line(const Game *g, int order, int *out) {
  int n = 0, fl = g->pl[order].front_line;
  if (order == 0) { for (int y = fl; y < 5; y++) for (int x = 0; x < 4; x++) out[n++] = PT(x, y); }
  else { for (int y = fl; y >= 0; y--) for (int x = 3; x >= 0; x--) out[n++] = PT(x, y); }
  return n;
}

/* ------------------------------------------------------------------ entities */
int o_new_ent(Game *g, int card, int owner, int strength) {
  if (g->n_ent >= MAXE) { ERR(g, SB_ERR_OVERFLOW); return MAXE - 1; }
  int id = g->n_ent++;
  Ent *e = &g->e[id];
  const OCard *c = &OCARDS[card];
  memset(e, 0, sizeof *e);
  e->card = card; e->owner = owner; e->is_struct = (c->kind == KIND_STRUCTURE);
  e->fixed = c->fixed; e->strength = strength; e->movement = c->movement;
  e->trigger = c->trigger; e->types = c->types; e->has_ability = c->has_ability;
  e->x = 0; e->y = 0;
  return id;
}
/* board.py:298-311 (all callers pass exactly one type) */
int o_spawn_token_unit(Game *g, int owner, int pt, int strength, int type) {
  int id = o_new_ent(g, SBC_TOKEN_UNIT0 + type, owner, strength);
  o_set(g, PTX(pt), PTY(pt), id);
  o_calc_front_line(g, owner);
  return id;
}

/* ------------------------------------------------------------------ trigger stack (board.py:46-56, card.py:48-62) */
static void push_trigger(Game *g, int id, int has_source) {
  if (g->n_trig >= MAXTRIG) { ERR(g, SB_ERR_OVERFLOW); return; }
  g->trig_ent[g->n_trig] = id; g->trig_src[g->n_trig] = has_source; g->n_trig++;
}
static void pop_trigger(Game *g) {
  if (g->n_trig == 0 || g->resolving) return;
  g->n_trig--;
  o_ability(g, g->trig_ent[g->n_trig], PT_NONE, g->trig_src[g->n_trig]);
}
/* the wrapper card.py:54-60 exists only on classes that override activate_ability */
void o_ability(Game *g, int id, int pos_pt, int has_source) {
  if (!g->e[id].has_ability) return;
  if (g->depth > MAXDEPTH) { ERR(g, SB_ERR_DEPTH); return; }
  g->depth++;
  g->resolving = 1;
  o_effect(g, id, pos_pt, has_source);
  g->resolving = 0;
  pop_trigger(g);
  g->depth--;
}
void o_spell_ability(Game *g, int card, int caster, int pos_pt) {
  if (g->depth > MAXDEPTH) { ERR(g, SB_ERR_DEPTH); return; }
  g->depth++;
  g->resolving = 1;
  o_spell_effect(g, card, caster, pos_pt);
  g->resolving = 0;
  pop_trigger(g);
  g->depth--;
}

/* ------------------------------------------------------------------ status verbs (unit.py:239-275) */
void o_st_add(Game *g, int id, int st) { /* 6-bit packed counters: the 64th copy of one status flags the game */
  if (g->e[id].st[st] < 63) g->e[id].st[st]++; else ERR(g, SB_ERR_OVERFLOW);
}
void o_st_remove(Game *g, int id, int st) { /* list.remove raises ValueError if absent */
  if (g->e[id].st[st] > 0) g->e[id].st[st]--; else ERR(g, SB_ERR_INDEX);
}
void o_freeze(Game *g, int id) { o_st_add(g, id, SB_ST_FROZEN); }
void o_poison(Game *g, int id) {
  if (g->e[id].st[SB_ST_VITALIZED] > 0) o_st_remove(g, id, SB_ST_VITALIZED);
  o_st_add(g, id, SB_ST_POISONED);
}
void o_vitalize(Game *g, int id) {
  if (g->e[id].st[SB_ST_POISONED] > 0) o_st_remove(g, id, SB_ST_POISONED);
  o_st_add(g, id, SB_ST_VITALIZED);
}
void o_confuse(Game *g, int id) { o_st_add(g, id, SB_ST_CONFUSED); }
void o_disable(Game *g, int id) { if (g->e[id].has_ability) o_st_add(g, id, SB_ST_DISABLED); } /* unit.py:269-271 */
void o_heal(Game *g, int id, int amount) { g->e[id].strength += amount; }

/* ------------------------------------------------------------------ damage (unit.py:205-231, structure.py:52-69, player.py:83-88) */
int o_player_damage(Game *g, int order, int amount) { g->pl[order].base -= amount; return amount; }

void o_destroy(Game *g, int id, int has_source) {
  Ent *e = &g->e[id];
  if (e->is_struct) { /* structure.py:65-69 */
    e->damage_taken = e->strength;
    g->board[e->y][e->x] = -1;
    o_calc_front_line(g, o_opponent(g, g->current_order));
    return;
  }
  g->board[e->y][e->x] = -1; /* stale position on purpose (Q21) */
  e->path_len = 0;
  e->damage_taken = e->strength;
  if (e->trigger == TR_ON_DEATH) { push_trigger(g, id, has_source); pop_trigger(g); }
  o_calc_front_line(g, o_opponent(g, g->current_order));
}
int o_unit_deal_damage(Game *g, int id, int amount, int pending, int has_source) {
  Ent *e = &g->e[id];
  if (e->strength - amount < 0) amount = e->strength;
  e->damage_taken = amount;
  e->strength -= amount;
  if (!pending && e->strength <= 0) o_destroy(g, id, has_source);
  else if (e->trigger == TR_AFTER_SURVIVING && e->strength > 0) { push_trigger(g, id, has_source); pop_trigger(g); }
  return amount;
}
int o_struct_deal_damage(Game *g, int id, int amount, int pending, int has_source) {
  Ent *e = &g->e[id];
  if (e->strength - amount < 0) amount = e->strength;
  e->damage_taken = amount;
  e->strength -= amount;
  if (!pending && e->strength <= 0) o_destroy(g, id, has_source);
  return amount;
}
/* `board.at(point).deal_damage(amount, source=self)` for a point that may be a base (board.py:58-60) */
int o_deal_damage_pt(Game *g, int pt, int amount, int has_source) {
  if (pt == PT_BASE_LOCAL) return o_player_damage(g, g->local_order, amount);
  if (pt == PT_BASE_REMOTE) return o_player_damage(g, 1 - g->local_order, amount);
  int id = o_at_pt(g, pt);
  if (id < 0) { ERR(g, SB_ERR_NONE_TARGET); return 0; }
  return g->e[id].is_struct ? o_struct_deal_damage(g, id, amount, 0, has_source)
                            : o_unit_deal_damage(g, id, amount, 0, has_source);
}

/* ------------------------------------------------------------------ movement (unit.py:66-203) */
static int xy_in(const XY *l, int n, int x, int y) { for (int i = 0; i < n; i++) if (l[i].x == x && l[i].y == y) return 1; return 0; }

void o_set_path(Game *g, int id, int on_play) { /* unit.py:78-122 */
  Ent *e = &g->e[id];
  XY dest[MAXPATH];
  int nd = 0;
  int px = e->x, py = e->y;
  int confused_cached = e->st[SB_ST_CONFUSED];
  int is_local = (e->owner == g->local_order);
  int steps = on_play ? e->movement : 1;
  if (steps > MAXPATH) { ERR(g, SB_ERR_OVERFLOW); steps = MAXPATH; }
  for (int i = 0; i < steps; i++) {
    int dx = px, dy = py + (is_local ? -1 : 1);
    int nxt = o_at(g, dx, dy);
    if (confused_cached > 0) {
      int delta;
      if (px == 0) { o_rng_below(g, 1); delta = 1; }
      else if (px == 3) { o_rng_below(g, 1); delta = -1; }
      else delta = o_rng_below(g, 2) == 0 ? -1 : 1;
      dx = px + delta; dy = py;
      confused_cached--;
    } else if (on_play && !e->fixed && dy != (is_local ? -1 : 5) && (nxt < 0 || g->e[nxt].owner == e->owner)) {
      int left = px > 0 ? o_at(g, px - 1, py) : -1;
      int right = px < 3 ? o_at(g, px + 1, py) : -1;
      int left_ok = left >= 0 && g->e[left].owner != e->owner && !xy_in(dest, nd, px - 1, py);
      int right_ok = right >= 0 && g->e[right].owner != e->owner && !xy_in(dest, nd, px + 1, py);
      if (px <= 1) {
        if (right_ok) { dx = px + 1; dy = py; }
        else if (left_ok) { dx = px - 1; dy = py; }
      } else {
        if (left_ok) { dx = px - 1; dy = py; }
        else if (right_ok) { dx = px + 1; dy = py; }
      }
    }
    dest[nd].x = (int8_t)dx; dest[nd].y = (int8_t)dy; nd++;
    px = dx; py = dy;
  }
  memcpy(e->path, dest, sizeof(XY) * nd);
  e->path_len = nd;
}

void o_move(Game *g, int id) { /* unit.py:124-203 */
  Ent *e = &g->e[id];
  if (g->depth > MAXDEPTH) { ERR(g, SB_ERR_DEPTH); return; }
  g->depth++;
  e->move_id++;
  int current_id = e->move_id;
  if (g->phase == PH_TURN_START) {
    if (e->st[SB_ST_POISONED] > 0) o_unit_deal_damage(g, id, 1, 0, 0);
    else if (e->st[SB_ST_VITALIZED] > 0) o_heal(g, id, 1);
    if (e->st[SB_ST_FROZEN] > 0) { o_st_remove(g, id, SB_ST_FROZEN); g->depth--; return; }
  }
  if (e->path_len == 0) { g->depth--; return; }
  if (e->trigger == TR_BEFORE_MOVING && e->st[SB_ST_DISABLED] == 0) o_ability(g, id, PT_NONE, 1);
  if (e->st[SB_ST_FROZEN] > 0) { g->depth--; return; }
  XY path[MAXPATH];
  int np = e->path_len;
  memcpy(path, e->path, sizeof(XY) * np); /* `for destination in self.path` iterates the list object bound now */
  for (int i = 0; i < np; i++) {
    int dx = path[i].x, dy = path[i].y;
    if (dy < 0 || dy > 4) { /* to base */
      if (e->trigger == TR_BEFORE_ATTACKING && e->st[SB_ST_DISABLED] == 0) o_ability(g, id, -2 - (dy < 0 ? 0 : 1), 1);
      int target = dy < 0 ? 1 - g->local_order : g->local_order;
      o_player_damage(g, target, e->strength);
      if (g->pl[target].base > 0) o_destroy(g, id, 0);
      g->depth--;
      return;
    }
    int tid = o_at(g, dx, dy);
    int attacked = 0;
    if (tid >= 0 && g->e[tid].owner == e->owner && dx == e->x) { g->depth--; return; }
    if (tid >= 0 && (e->st[SB_ST_CONFUSED] > 0 || g->e[tid].owner != e->owner)) {
      if (e->trigger == TR_BEFORE_ATTACKING && e->st[SB_ST_DISABLED] == 0) o_ability(g, id, PT(dx, dy), 1);
      tid = o_at(g, dx, dy);
      if (tid >= 0) {
        Ent *t = &g->e[tid];
        int tstr = t->strength;
        int t_pending = !t->is_struct && t->trigger == TR_ON_DEATH && t->st[SB_ST_DISABLED] == 0;
        int l_pending = e->trigger == TR_ON_DEATH && e->st[SB_ST_DISABLED] == 0;
        if (t->is_struct) o_struct_deal_damage(g, tid, e->strength, t_pending, 0);
        else o_unit_deal_damage(g, tid, e->strength, t_pending, 0);
        o_unit_deal_damage(g, id, tstr, l_pending, 0);
        if (t->strength <= 0 && t_pending) o_destroy(g, tid, 0);
        if (e->strength <= 0 && l_pending) o_destroy(g, id, 0);
        attacked = 1;
      }
    }
    if (current_id != e->move_id) { g->depth--; return; }
    if (o_at(g, dx, dy) < 0 && e->strength > 0) {
      g->board[e->y][e->x] = -1;
      o_set(g, dx, dy, id);
      Ply *p = &g->pl[e->owner];
      if (p->front_line > dy) p->front_line = dy > 1 ? dy : 1;
      if (attacked && e->trigger == TR_AFTER_ATTACKING && e->st[SB_ST_DISABLED] == 0) o_ability(g, id, PT_NONE, 1);
      if (e->st[SB_ST_CONFUSED] > 0) o_st_remove(g, id, SB_ST_CONFUSED);
    }
  }
  g->depth--;
}

static void unit_play(Game *g, int id, int x, int y) { /* unit.py:66-76 */
  Ent *e = &g->e[id];
  e->resolving_play = 1;
  o_set(g, x, y, id);
  o_set_path(g, id, 1);
  if (e->trigger == TR_ON_PLAY) o_ability(g, id, PT_NONE, 1);
  o_move(g, id);
  e->resolving_play = 0;
}
void o_struct_play(Game *g, int id, int x, int y) { /* structure.py:45-50 */
  o_set(g, x, y, id);
  if (g->e[id].trigger == TR_ON_PLAY) o_ability(g, id, PT_NONE, 1);
}
void o_gain_speed(Game *g, int id, int amount) { /* unit.py:277-280 */
  Ent *e = &g->e[id];
  e->movement += amount;
  o_set_path(g, id, e->resolving_play);
  e->movement -= amount;
}
void o_command(Game *g, int id) { /* unit.py:282-289 */
  Ent *e = &g->e[id];
  int cache = e->fixed;
  e->fixed = 1;
  o_set_path(g, id, 0);
  o_move(g, id);
  e->fixed = cache;
}
void o_convert(Game *g, int id) { /* unit.py:291-293 */
  Ent *e = &g->e[id];
  e->owner = o_opponent(g, e->owner);
  o_set_path(g, id, e->resolving_play);
}
void o_push(Game *g, int id, int fx, int fy) { /* unit.py:318-339: push this unit away from (fx,fy) */
  Ent *e = &g->e[id];
  int dx = 0, dy = 0;
  if (fy < e->y) dy = 1; else if (fy > e->y) dy = -1; else if (fx < e->x) dx = 1; else if (fx > e->x) dx = -1;
  if (dx || dy) {
    for (;;) {
      int nx = e->x + dx, ny = e->y + dy;
      if (!valid_xy(nx, ny)) break;
      if (g->board[ny][nx] >= 0) return; /* `return` skips the front-line update too */
      g->board[e->y][e->x] = -1;
      o_set(g, nx, ny, id);
    }
  }
  Ply *p = &g->pl[e->owner];
  if (p->front_line > e->y) p->front_line = e->y > 1 ? e->y : 1;
}
void o_force_attack(Game *g, int id, int tx, int ty) { /* unit.py:341-371 */
  Ent *e = &g->e[id];
  if ((tx != e->x && ty != e->y) || o_at(g, tx, ty) < 0) return;
  XY dest[MAXPATH];
  int nd = 0;
  int vertical = (tx == e->x);
  int fixed = vertical ? e->x : e->y, start = vertical ? e->y : e->x, end = vertical ? ty : tx;
  int delta = end > start ? 1 : -1;
  for (int i = start + delta; i != end + delta; i += delta) {
    int x = vertical ? fixed : i, y = vertical ? i : fixed;
    if (i != end && o_at(g, x, y) >= 0) return;
    dest[nd].x = (int8_t)x; dest[nd].y = (int8_t)y; nd++;
  }
  if (nd > 0) {
    memcpy(e->path, dest, sizeof(XY) * nd);
    e->path_len = nd;
    o_move(g, id);
  }
}
void o_teleport(Game *g, int id, int dx, int dy) { /* unit.py:373-382 */
  Ent *e = &g->e[id];
  if (o_at(g, dx, dy) < 0) {
    g->board[e->y][e->x] = -1;
    o_set(g, dx, dy, id);
    Ply *p = &g->pl[e->owner];
    if (p->front_line > dy) p->front_line = dy > 1 ? dy : 1;
    o_set_path(g, id, e->resolving_play);
  }
}

/* ------------------------------------------------------------------ hand / deck (player.py:46-81) */
/* list.remove(target): index of the first element that `is` target or == target.
 * Unit/Structure __eq__: card_id, player, position (unit.py:25-26, structure.py:18-19); Spell: uuid
 * (card.py:22-23).  A board instance of B305 (SB_CF_OBJ) carries a Point position, a pristine card
 * None: comparing the two evaluates Point.__eq__(None) -> AttributeError (point.py:7). */
static int first_equal(Game *g, const CardRec *l, int n, int idx) {
  const CardRec *t = &l[idx];
  if (OCARDS[t->card].kind == KIND_SPELL) return idx;
  for (int i = 0; i < idx && i < n; i++) {
    if (l[i].card != t->card) continue;
    int oi = l[i].flags & SB_CF_OBJ, ot = t->flags & SB_CF_OBJ;
    if (oi != ot) { ERR(g, SB_ERR_NONE_TARGET); return idx; }
    if (!oi) return i; /* both pristine: None == None */
  }
  return idx;
}
static void player_draw(Game *g, int order, int amount) { /* player.py:46-52 */
  Ply *p = &g->pl[order];
  for (int k = 0; k < amount; k++) {
    int n = p->n_deck;
    if (n <= 0) { ERR(g, SB_ERR_EMPTY_CHOICE); return; }
    double sum = 0.0, cdf[DECK_W], acc = 0.0;
    for (int i = 0; i < n; i++) sum = sum + WT[p->deck[i].wn];
    for (int i = 0; i < n; i++) { acc = acc + WT[p->deck[i].wn] / sum; cdf[i] = acc; }
    double last = cdf[n - 1];
    double u = o_rng_random(g);
    int idx = 0;
    for (int i = 0; i < n; i++) if (cdf[i] / last <= u) idx++;
    if (idx > n - 1) idx = n - 1;
    CardRec c = p->deck[idx];
    c.wn = 0;
    if (p->n_hand >= HAND_W) { ERR(g, SB_ERR_OVERFLOW); return; }
    p->hand[p->n_hand++] = c;
    int j = first_equal(g, p->deck, n, idx);
    if (j != idx) p->deck[idx].wn = 0; /* the drawn object stays in the deck (weight 1); an equal one leaves */
    memmove(&p->deck[j], &p->deck[j + 1], sizeof(CardRec) * (n - j - 1));
    p->n_deck--;
  }
}
static void player_fill_hand(Game *g, int order) { player_draw(g, order, 4 - g->pl[order].n_hand); }
static void player_discard(Game *g, int order, int index) { /* player.py:57-66 */
  Ply *p = &g->pl[order];
  for (int i = 0; i < p->n_deck; i++) {
    if (p->deck[i].wn >= 1023) ERR(g, SB_ERR_OVERFLOW); else p->deck[i].wn++;
  }
  CardRec target = p->hand[index];
  int j = first_equal(g, p->hand, p->n_hand, index);
  memmove(&p->hand[j], &p->hand[j + 1], sizeof(CardRec) * (p->n_hand - j - 1));
  p->n_hand--;
  if (!(target.flags & SB_CF_SINGLE_USE)) {
    if (p->n_deck >= DECK_W) { ERR(g, SB_ERR_OVERFLOW); return; }
    target.wn = 0;
    p->deck[p->n_deck++] = target;
  }
}
void o_player_play(Game *g, int order, int index, int pos_pt) { /* player.py:68-77 */
  Ply *p = &g->pl[order];
  if (index < 0 || index >= p->n_hand) { ERR(g, SB_ERR_INDEX); return; }
  CardRec target = p->hand[index];
  if (g->hist_n < 4) { g->hist_card[g->hist_n] = target.card; g->hist_owner[g->hist_n] = order; g->hist_n++; }
  else {
    for (int i = 0; i < 3; i++) { g->hist_card[i] = g->hist_card[i + 1]; g->hist_owner[i] = g->hist_owner[i + 1]; }
    g->hist_card[3] = target.card; g->hist_owner[3] = order;
  }
  player_discard(g, order, index);
  const OCard *c = &OCARDS[target.card];
  if (c->kind == KIND_SPELL) { /* spell.py:22-24 */
    int ok = 1;
    if (c->has_target) {
      Target t = {c->t_kind, c->t_side, c->t_types, c->t_xtypes, c->t_status, c->t_xstatus, c->t_limit >= 0, c->t_limit, c->t_nonhero, c->t_base};
      int tg[24];
      int n = o_get_targets(g, g->current_order, &t, PT_NONE, tg);
      ok = in_list(tg, n, pos_pt);
    }
    if (ok) o_spell_ability(g, target.card, order, pos_pt);
    return;
  }
  if (pos_pt < 0 || pos_pt >= 20) { ERR(g, SB_ERR_INDEX); return; }
  int strength = c->strength;
  if (target.flags & SB_CF_OBJ) strength = target.link >= 0 ? g->e[target.link].strength : target.xstr;
  int id = o_new_ent(g, target.card, order, strength); /* target.copy(), player.py:74 */
  g->e[id].fixed = (target.flags & SB_CF_FIXED) ? 1 : 0;
  g->e[id].single_use = (target.flags & SB_CF_SINGLE_USE) ? 1 : 0;
  if (c->kind == KIND_UNIT) unit_play(g, id, PTX(pos_pt), PTY(pos_pt));
  else o_struct_play(g, id, PTX(pos_pt), PTY(pos_pt));
}
static void player_cycle(Game *g, int order, int index) { /* player.py:79-81 */
  player_discard(g, order, index);
  player_draw(g, order, 1);
}

/* ------------------------------------------------------------------ turn pipeline (board.py:94-145) */
static void board_flip(Game *g) { /* board.py:94-115 */
  g->local_order ^= 1;
  g->pl[0].front_line = 4 - g->pl[0].front_line;
  g->pl[1].front_line = 4 - g->pl[1].front_line;
  int nb[5][4];
  for (int y = 0; y < 5; y++) for (int x = 0; x < 4; x++) nb[y][x] = g->board[4 - y][3 - x];
  memcpy(g->board, nb, sizeof nb);
  for (int y = 0; y < 5; y++) for (int x = 0; x < 4; x++) {
    int id = g->board[y][x];
    if (id >= 0) { g->e[id].x = x; g->e[id].y = y; }
  }
}
static void to_next_turn(Game *g) { /* board.py:117-145 */
  static const Target T_STRUCT_F = {TK_STRUCTURE, TS_FRIENDLY, 0, 0, 0, 0, 0, 0, 0, 0};
  static const Target T_UNIT_F = {TK_UNIT, TS_FRIENDLY, 0, 0, 0, 0, 0, 0, 0, 0};
  int pts[24], ids[24], n;
  g->phase = PH_TURN_END;
  player_fill_hand(g, g->current_order);
  n = o_get_targets(g, g->current_order, &T_STRUCT_F, PT_NONE, pts);
  for (int i = 0; i < n; i++) ids[i] = o_at_pt(g, pts[i]);
  for (int i = 0; i < n; i++)
    if (g->e[ids[i]].trigger == TR_TURN_END) o_ability(g, ids[i], PT(g->e[ids[i]].x, g->e[ids[i]].y), 1);
  o_calc_front_line(g, g->local_order);
  o_calc_front_line(g, 1 - g->local_order);
  g->pl[g->current_order].max_mana += 1;
  g->pl[0].mana = g->pl[0].max_mana;
  g->pl[1].mana = g->pl[1].max_mana;
  g->phase = PH_TURN_START;
  g->current_order = (g->current_order == g->local_order) ? 1 - g->local_order : g->local_order;
  g->pl[g->current_order].replacable = 1;
  g->pl[g->current_order].leftmost = 1;
  n = o_get_targets(g, g->current_order, &T_STRUCT_F, PT_NONE, pts);
  for (int i = 0; i < n; i++) ids[i] = o_at_pt(g, pts[i]);
  for (int i = 0; i < n; i++)
    if (g->e[ids[i]].trigger == TR_TURN_START) o_ability(g, ids[i], PT(g->e[ids[i]].x, g->e[ids[i]].y), 1);
  n = o_get_targets(g, g->current_order, &T_UNIT_F, PT_NONE, pts);
  for (int i = 0; i < n; i++) ids[i] = o_at_pt(g, pts[i]);
  for (int i = 0; i < n; i++) { o_set_path(g, ids[i], 0); o_move(g, ids[i]); } /* ghosts included (Q21) */
  g->phase = PH_PLAY;
}

/* ------------------------------------------------------------------ legal actions / step (games/stormbound.py:318-373,528-561) */
static void mask_set(uint32_t *m, int a) { m[a >> 5] |= 1u << (a & 31); }
static int legal_mask(Game *g, uint32_t m[SB_MASK_WORDS]) {
  Ply *p = o_local(g);
  int n_play = 0, n = 0;
  memset(m, 0, sizeof(uint32_t) * SB_MASK_WORDS);
  for (int ci = 0; ci < p->n_hand && ci < 4; ci++) {
    const OCard *c = &OCARDS[p->hand[ci].card];
    if (p->hand[ci].cost > p->mana) continue;
    if (c->kind != KIND_SPELL) {
      for (int y = 4; y >= p->front_line && y >= 1; y--)
        for (int x = 0; x < 4; x++)
          if (g->board[y][x] < 0) { mask_set(m, 16 * ci + (4 - y) * 4 + x); n_play++; }
    } else if (!c->has_target) {
      mask_set(m, 64 + 21 * ci); n_play++;
    } else {
      Target t = {c->t_kind, c->t_side, c->t_types, c->t_xtypes, c->t_status, c->t_xstatus, c->t_limit >= 0, c->t_limit, c->t_nonhero, c->t_base};
      int tg[24];
      int nt = o_get_targets(g, g->current_order, &t, PT_NONE, tg);
      for (int i = 0; i < nt; i++) { /* Action.to_int: 65 + 21*card + ordinal over y=4..0, x=0..3 */
        if (is_base_pt(tg[i])) continue; /* to_int falls through to 155 for a base point */
        mask_set(m, 65 + 21 * ci + (4 - PTY(tg[i])) * 4 + PTX(tg[i])); n_play++;
      }
    }
  }
  n = n_play;
  if (p->replacable) for (int ci = 0; ci < p->n_hand && ci < 4; ci++) { mask_set(m, 148 + ci); n++; }
  if (n_play == 0) { mask_set(m, 155); n++; }
  return n;
}
static int have_winner(Game *g) { return g->pl[0].base < 0 || g->pl[1].base < 0; }

/* ------------------------------------------------------------------ scripted opponent (games/stormbound.py:563-637) */
static int mask_has_range(const uint32_t *m, int lo, int hi) {
  for (int a = lo; a <= hi; a++) if (m[a >> 5] >> (a & 31) & 1) return 1;
  return 0;
}
static int place_action(int ci, int pt) { /* Action.to_int PLACE (games/stormbound.py:261-270); row 0 is not encodable -> 155 */
  int x = PTX(pt), y = PTY(pt);
  return (y >= 1 && y <= 4) ? 16 * ci + (4 - y) * 4 + x : SB_ACTION_PASS;
}
static int expert_action(Game *g) {
  uint32_t m[SB_MASK_WORDS];
  Ply *p = o_local(g);
  int pts[24], n;
  legal_mask(g, m);
  if (mask_has_range(m, 148, 151)) {
    int max_cost = -1000, sel[SB_HAND_MAX + 2], ns = 0;
    if (p->n_hand == 0) { ERR(g, SB_ERR_EMPTY_CHOICE); return SB_ACTION_PASS; } /* max([]) */
    for (int i = 0; i < p->n_hand; i++) if (p->hand[i].cost > max_cost) max_cost = p->hand[i].cost;
    if (max_cost > p->mana) {
      for (int i = 0; i < p->n_hand; i++) if (p->hand[i].cost == max_cost) sel[ns++] = i;
      return 148 + sel[o_rng_below(g, ns)];
    }
  }
  int playable[4], np = 0;
  for (int i = 0; i < 4; i++) if (mask_has_range(m, 16 * i, 16 * i + 15) || mask_has_range(m, 21 * i + 64, 21 * i + 84)) playable[np++] = i;
  if (np > 0) {
    int sel[4], ns = 0, any_eq = 0, min_cost = 1 << 30;
    for (int k = 0; k < np; k++) { int c = p->hand[playable[k]].cost; if (c == p->mana) any_eq = 1; if (c < min_cost) min_cost = c; }
    for (int k = 0; k < np; k++) if (p->hand[playable[k]].cost == (any_eq ? p->mana : min_cost)) sel[ns++] = k;
    int index = playable[sel[o_rng_below(g, ns)]];
    const OCard *c = &OCARDS[p->hand[index].card];
    static const Target T_UNIT_E = {TK_UNIT, TS_ENEMY, 0, 0, 0, 0, 0, 0, 0, 0};
    int bbe[24], nb = 0;
    n = o_get_targets(g, g->current_order, &T_UNIT_E, PT_NONE, pts);
    for (int i = 0; i < n; i++) if (PTY(pts[i]) == 4) bbe[nb++] = pts[i];
    if (c->kind == KIND_SPELL) {
      if (!c->has_target) return 64 + 21 * index;
      Target t = {c->t_kind, c->t_side, c->t_types, c->t_xtypes, c->t_status, c->t_xstatus, c->t_limit >= 0, c->t_limit, c->t_nonhero, c->t_base};
      int tg[24], nt = o_get_targets(g, g->current_order, &t, PT_NONE, tg);
      if (nt == 0) { ERR(g, SB_ERR_EMPTY_CHOICE); return SB_ACTION_PASS; }
      int where = tg[o_rng_below(g, nt)];
      if (is_base_pt(where)) return SB_ACTION_PASS;
      return 65 + 21 * index + (4 - PTY(where)) * 4 + PTX(where);
    } else if (c->kind == KIND_UNIT && nb > 0) {
      int cand[24], nc = 0;
      for (int i = 0; i < nb; i++) {
        int x = PTX(bbe[i]), y = PTY(bbe[i]);
        if (x > 0 && o_at(g, x - 1, y) < 0) cand[nc++] = PT(x - 1, y);
        else if (x < 3 && o_at(g, x + 1, y) < 0) cand[nc++] = PT(x + 1, y);
      }
      if (nc > 0) return place_action(index, cand[o_rng_below(g, nc)]);
    } else {
      int cand[48], nc = 0, fl = p->front_line;
      for (int x = 0; x < 4; x++) if (o_at(g, x, fl) < 0 && valid_xy(x, fl)) cand[nc++] = PT(x, fl);
      for (int i = 0; i < n; i++) {
        int x = PTX(pts[i]), y = PTY(pts[i]);
        if (x > 0 && y >= fl && o_at(g, x - 1, y) < 0) cand[nc++] = PT(x - 1, y);
        else if (x < 3 && y >= fl && o_at(g, x + 1, y) < 0) cand[nc++] = PT(x + 1, y);
        else if (y < 4 && y + 1 >= fl && o_at(g, x, y + 1) < 0) cand[nc++] = PT(x, y + 1);
      }
      if (nc > 0) return place_action(index, cand[o_rng_below(g, nc)]);
    }
  }
  return SB_ACTION_PASS;
}

static void game_step(Game *g, int action) {
  Ply *p = o_local(g);
  if (action < 64) {
    int ci = action / 16, idx = action % 16;
    if (ci >= p->n_hand) { ERR(g, SB_ERR_INDEX); }
    else { p->mana -= p->hand[ci].cost; o_player_play(g, g->local_order, ci, PT(idx % 4, 4 - idx / 4)); }
  } else if (action < 148) {
    int ci = (action - 64) / 21, idx = (action - 64) % 21;
    if (idx < 20) { /* index 20 never matches a tile: complete no-op (Q5) */
      if (ci >= p->n_hand) { ERR(g, SB_ERR_INDEX); }
      else {
        const OCard *c = &OCARDS[p->hand[ci].card];
        if (c->kind != KIND_SPELL) ERR(g, SB_ERR_INDEX); /* .required_targets AttributeError on a Unit */
        else { p->mana -= p->hand[ci].cost; o_player_play(g, g->local_order, ci, c->has_target ? PT(idx % 4, 4 - idx / 4) : PT_NONE); }
      }
    }
  } else if (action < 152) {
    int ci = action - 148;
    if (ci >= p->n_hand) ERR(g, SB_ERR_INDEX);
    else { player_cycle(g, g->local_order, ci); p->replacable = 0; }
  } else if (action < 155) {
    int ci = action - 151;
    if (ci >= p->n_hand) ERR(g, SB_ERR_INDEX);
    else { CardRec t = p->hand[ci]; p->hand[ci] = p->hand[0]; p->hand[0] = t; p->leftmost = 0; }
  }
  uint32_t m[SB_MASK_WORDS];
  int done = have_winner(g) || legal_mask(g, m) == 0;
  int reward = o_remote(g)->base <= 0;
  g->done = (done ? SB_DONE : 0) | (reward ? SB_REWARD : 0);
  if (action == 155) {
    g->turn++; g->draw = 0; /* harness hook: the stream is keyed by (turn, draw) */
    g->player_sign = -g->player_sign;
    board_flip(g);
    to_next_turn(g);
  }
  g->steps++;
}

/* ------------------------------------------------------------------ pack / unpack */
void o_unpack(Game *g, const SbState *s) {
  wt_init();
  memset(g, 0, sizeof *g);
  g->seed_lo = s->seed_lo; g->seed_hi = s->seed_hi; g->turn = s->turn; g->draw = s->draw; g->steps = s->steps;
  g->local_order = s->local_order; g->current_order = s->current_order; g->player_sign = s->player_sign;
  g->phase = s->phase; g->err = s->err; g->done = s->done;
  g->hist_n = s->hist_n;
  for (int i = 0; i < 4; i++) { g->hist_card[i] = s->hist_card[i]; g->hist_owner[i] = s->hist_owner[i]; }
  for (int o = 0; o < 2; o++) {
    const SbPlayer *sp = &s->pl[o];
    Ply *p = &g->pl[o];
    p->base = sp->base; p->max_mana = sp->max_mana; p->mana = sp->mana; p->front_line = sp->front_line;
    p->replacable = !!(sp->flags & SB_PF_REPLACABLE); p->leftmost = !!(sp->flags & SB_PF_LEFTMOST);
    p->faction = sp->faction; p->n_hand = sp->n_hand; p->n_deck = sp->n_deck;
    for (int i = 0; i < SB_HAND_MAX; i++) { p->hand[i].card = sp->hand_card[i]; p->hand[i].cost = sp->hand_cost[i]; p->hand[i].flags = sp->hand_flags[i]; p->hand[i].wn = 0; p->hand[i].xstr = 0; p->hand[i].link = -1; }
    for (int i = 0; i < SB_DECK_MAX; i++) { p->deck[i].card = sp->deck_card[i]; p->deck[i].cost = sp->deck_cost[i]; p->deck[i].flags = sp->deck_flags[i]; p->deck[i].wn = sp->deck_wn[i]; p->deck[i].xstr = 0; p->deck[i].link = -1; }
  }
  for (int y = 0; y < 5; y++) for (int x = 0; x < 4; x++) {
    const SbTile *t = &s->tile[y * 4 + x];
    g->board[y][x] = -1;
    if (!t->card) continue;
    int id = o_new_ent(g, t->card, (t->flags & SB_TF_OWNER) ? 1 : 0, t->strength);
    Ent *e = &g->e[id];
    e->fixed = !!(t->flags & SB_TF_FIXED);
    for (int k = 0; k < 5; k++) e->st[k] = (t->status >> (SB_ST_BITS * k)) & ((1 << SB_ST_BITS) - 1);
    o_set(g, x, y, id);
  }
  /* ext: B005 memories keyed by the B005's current tile, then B305 board-instance card records */
  const uint8_t *x = s->ext;
  int nm = x[0];
  for (int i = 0; i < nm && i < NMEM_PACKED; i++) {
    const uint8_t *r = x + 1 + 10 * i;
    Mem *m = &g->mem[g->n_mem++];
    if (r[0] & 0x80) { m->parent = r[0] & 0x7F; m->b005 = -1; }   /* memory of the remembered temple copy #parent */
    else { m->parent = -1; m->b005 = o_at_pt(g, r[0]); }
    m->pos = r[1]; m->card = r[2];
    m->owner = (r[3] & SB_TF_OWNER) ? 1 : 0; m->is_struct = !!(r[3] & SB_TF_STRUCTURE); m->fixed = !!(r[3] & SB_TF_FIXED);
    m->detached = !!(r[3] & 8);
    m->strength = (int16_t)(r[4] | (r[5] << 8));
    uint32_t w = r[6] | (r[7] << 8) | (r[8] << 16) | ((uint32_t)r[9] << 24);
    for (int k = 0; k < 5; k++) m->st[k] = (w >> (SB_ST_BITS * k)) & 63;
  }
  int no = x[91];
  for (int i = 0; i < no && i < NOBJ_PACKED; i++) {
    const uint8_t *r = x + 92 + 4 * i;
    Ply *p = &g->pl[r[0] >> 7];
    CardRec *c = (r[0] & 64) ? &p->deck[r[0] & 63] : &p->hand[r[0] & 63];
    if (r[1] != 0xFF) c->link = o_at_pt(g, r[1]);
    else { c->link = -1; c->xstr = (int16_t)(r[2] | (r[3] << 8)); }
  }
}
static void pack_mem(const Game *g, SbState *s, int i, int key, int *nm) {
  if (*nm >= NMEM_PACKED) { if (!s->err) s->err = SB_ERR_OVERFLOW; return; }
  const Mem *m = &g->mem[i];
  const int me = (*nm)++;
  uint8_t *r = s->ext + 1 + 10 * me;
  r[0] = (uint8_t)key; r[1] = (uint8_t)m->pos; r[2] = (uint8_t)m->card;
  r[3] = (m->owner ? SB_TF_OWNER : 0) | (m->is_struct ? SB_TF_STRUCTURE : 0) | (m->fixed ? SB_TF_FIXED : 0) | (m->detached ? 8 : 0);
  r[4] = (uint8_t)(m->strength & 255); r[5] = (uint8_t)((m->strength >> 8) & 255);
  uint32_t w = 0;
  if (!m->is_struct) for (int k = 0; k < 5; k++) w |= (uint32_t)(m->st[k] > 63 ? 63 : m->st[k]) << (SB_ST_BITS * k);
  r[6] = w & 255; r[7] = (w >> 8) & 255; r[8] = (w >> 16) & 255; r[9] = (w >> 24) & 255;
  for (int q = i + 1; q < g->n_mem; q++) if (g->mem[q].parent == i) pack_mem(g, s, q, 0x80 | me, nm);
}
void o_pack(const Game *g, SbState *s) {
  memset(s, 0, sizeof *s);
  s->seed_lo = g->seed_lo; s->seed_hi = g->seed_hi; s->turn = (uint16_t)g->turn; s->draw = (uint16_t)g->draw;
  s->steps = (uint16_t)g->steps;
  s->local_order = (uint8_t)g->local_order; s->current_order = (uint8_t)g->current_order;
  s->player_sign = (int8_t)g->player_sign; s->phase = (uint8_t)g->phase; s->err = (uint8_t)g->err; s->done = (uint8_t)g->done;
  s->hist_n = (uint8_t)g->hist_n;
  for (int i = 0; i < 4; i++) { s->hist_card[i] = (uint8_t)g->hist_card[i]; s->hist_owner[i] = (uint8_t)g->hist_owner[i]; }
  for (int o = 0; o < 2; o++) {
    SbPlayer *sp = &s->pl[o];
    const Ply *p = &g->pl[o];
    sp->base = (int16_t)p->base; sp->max_mana = (int16_t)p->max_mana; sp->mana = (int16_t)p->mana;
    sp->front_line = (int8_t)p->front_line;
    sp->flags = (p->replacable ? SB_PF_REPLACABLE : 0) | (p->leftmost ? SB_PF_LEFTMOST : 0);
    sp->faction = (uint8_t)p->faction; sp->n_hand = (uint8_t)p->n_hand; sp->n_deck = (uint8_t)p->n_deck;
    if (p->n_hand > SB_HAND_MAX || p->n_deck > SB_DECK_MAX) { /* more than the packed layout holds */
      if (!s->err) s->err = SB_ERR_OVERFLOW;
      if (p->n_hand > SB_HAND_MAX) sp->n_hand = SB_HAND_MAX;
      if (p->n_deck > SB_DECK_MAX) sp->n_deck = SB_DECK_MAX;
    }
    for (int i = 0; i < p->n_hand && i < SB_HAND_MAX; i++) { sp->hand_card[i] = (uint8_t)p->hand[i].card; sp->hand_cost[i] = (int8_t)p->hand[i].cost; sp->hand_flags[i] = (uint8_t)p->hand[i].flags; }
    for (int i = 0; i < p->n_deck && i < SB_DECK_MAX; i++) { sp->deck_card[i] = (uint8_t)p->deck[i].card; sp->deck_cost[i] = (int8_t)p->deck[i].cost; sp->deck_flags[i] = (uint8_t)p->deck[i].flags; sp->deck_wn[i] = (uint16_t)p->deck[i].wn; }
  }
  for (int y = 0; y < 5; y++) for (int x = 0; x < 4; x++) {
    int id = g->board[y][x];
    if (id < 0) continue;
    const Ent *e = &g->e[id];
    SbTile *t = &s->tile[y * 4 + x];
    t->card = (uint8_t)e->card;
    t->flags = (e->owner ? SB_TF_OWNER : 0) | (e->is_struct ? SB_TF_STRUCTURE : 0) | (e->fixed ? SB_TF_FIXED : 0);
    t->strength = (int16_t)e->strength;
    uint32_t w = 0;
    if (!e->is_struct) for (int k = 0; k < 5; k++) {
      int c = e->st[k] > 63 ? 63 : e->st[k];
      w |= (uint32_t)c << (SB_ST_BITS * k);
    }
    t->status = w;
  }
  uint8_t *x = s->ext;
  int nm = 0;
  for (int tile = 0; tile < 20; tile++) { /* canonical order: temples in tile order, each memory followed by its own subtree */
    int bid = g->board[tile >> 2][tile & 3];
    if (bid < 0 || g->e[bid].card != SBC_B005) continue;
    for (int i = 0; i < g->n_mem; i++)
      if (g->mem[i].parent < 0 && g->mem[i].b005 == bid) pack_mem(g, s, i, tile, &nm);
  }
  x[0] = (uint8_t)nm;
  int no = 0;
  for (int o = 0; o < 2; o++) for (int where = 0; where < 2; where++) {
    const Ply *p = &g->pl[o];
    int cnt = where ? p->n_deck : p->n_hand;
    for (int i = 0; i < cnt; i++) {
      const CardRec *c = where ? &p->deck[i] : &p->hand[i];
      if (!(c->flags & SB_CF_OBJ)) continue;
      if (no >= NOBJ_PACKED) { s->err = s->err ? s->err : SB_ERR_OVERFLOW; break; }
      uint8_t *r = x + 92 + 4 * no++;
      r[0] = (uint8_t)((o << 7) | (where << 6) | i);
      int on_board = c->link >= 0 && g->board[g->e[c->link].y][g->e[c->link].x] == c->link;
      int str = c->link >= 0 ? g->e[c->link].strength : c->xstr;
      r[1] = on_board ? (uint8_t)PT(g->e[c->link].x, g->e[c->link].y) : 0xFF;
      r[2] = on_board ? 0 : (uint8_t)(str & 255); r[3] = on_board ? 0 : (uint8_t)((str >> 8) & 255);
    }
  }
  x[91] = (uint8_t)no;
}

/* ------------------------------------------------------------------ public (ctypes) entry points */
int sbo_state_bytes(void) { return (int)sizeof(SbState); }

/* games/stormbound.py:293-304 + player.py:13-37 */
void sbo_new_game(SbState *out, uint64_t seed, const uint8_t *deck0, const uint8_t *deck1, int n_deck, int faction0, int faction1) {
  Game *g = (Game *)malloc(sizeof(Game));
  wt_init();
  memset(g, 0, sizeof *g);
  g->seed_lo = (uint32_t)seed; g->seed_hi = (uint32_t)(seed >> 32);
  g->local_order = 0; g->current_order = 0; g->player_sign = 1; g->phase = PH_PLAY;
  for (int y = 0; y < 5; y++) for (int x = 0; x < 4; x++) g->board[y][x] = -1;
  for (int o = 0; o < 2; o++) {
    Ply *p = &g->pl[o];
    const uint8_t *deck = o == 0 ? deck0 : deck1;
    int order[SB_DECK_MAX];
    p->max_mana = o == 0 ? 3 : 4; p->mana = p->max_mana; p->base = 20; p->front_line = o == 0 ? 4 : 0;
    p->replacable = 1; p->leftmost = 1; p->faction = o == 0 ? faction0 : faction1;
    for (int i = 0; i < n_deck; i++) order[i] = deck[i];
    o_shuffle(g, order, n_deck);
    p->n_deck = n_deck;
    for (int i = 0; i < n_deck; i++) {
      p->deck[i].card = order[i]; p->deck[i].cost = OCARDS[order[i]].cost;
      p->deck[i].flags = OCARDS[order[i]].fixed ? SB_CF_FIXED : 0; p->deck[i].wn = i;
      p->deck[i].xstr = 0; p->deck[i].link = -1;
    }
    player_fill_hand(g, o);
  }
  o_pack(g, out);
  free(g);
}
void sbo_legal_mask(const SbState *s, uint32_t *mask) {
  Game *g = (Game *)malloc(sizeof(Game));
  o_unpack(g, s);
  legal_mask(g, mask);
  free(g);
}
void sbo_step(SbState *s, int action) {
  Game *g = (Game *)malloc(sizeof(Game));
  o_unpack(g, s);
  game_step(g, action);
  o_pack(g, s);
  free(g);
}
uint64_t sbo_digest(const SbState *s) {
  const uint8_t *b = (const uint8_t *)s;
  uint64_t h = 0xCBF29CE484222325ull;
  for (int i = 0; i < (int)sizeof(SbState); i++) { h ^= b[i]; h *= 0x100000001B3ull; }
  return h;
}
/* Stormbound.expert_action: returns the action, advances the game's random stream in *s (it draws from it). */
int sbo_expert_action(SbState *s) {
  Game *g = (Game *)malloc(sizeof(Game));
  o_unpack(g, s);
  int a = expert_action(g);
  o_pack(g, s);
  free(g);
  return a;
}
/* Uniform-random agent rollout of one game (config 2).  digests/actions may be NULL.  Returns steps. */
int sbo_rollout_random(SbState *s, int max_steps, uint8_t *actions, uint64_t *digests, uint32_t *masks) {
  Game *g = (Game *)malloc(sizeof(Game));
  uint64_t seed = ((uint64_t)s->seed_hi << 32) | s->seed_lo;
  int step = s->steps;
  int k = 0;
  while (!(s->done & SB_DONE) && !s->err && k < max_steps) {
    uint32_t m[SB_MASK_WORDS];
    int legal[SB_N_ACTIONS], n = 0;
    o_unpack(g, s);
    legal_mask(g, m);
    for (int a = 0; a < SB_N_ACTIONS; a++) if (m[a >> 5] >> (a & 31) & 1) legal[n++] = a;
    int a = legal[sbo_agent_pick(seed, (uint32_t)step, (uint32_t)n)];
    game_step(g, a);
    o_pack(g, s);
    if (actions) actions[k] = (uint8_t)a;
    if (masks) memcpy(masks + SB_MASK_WORDS * k, m, sizeof m);
    if (digests) digests[k] = sbo_digest(s);
    k++; step++;
  }
  free(g);
  return k;
}

/* ------------------------------------------------------------------ deck generation (utils.py:26-241)
 * generate_random_deck / DeckEvolutionConfig.get_deck_configuration with the injected stream
 *   block = philox(counter=(draw, generation, 0xDEC4, 0), key=game seed)
 * random.sample (CPython 3.12 Lib/random.py:359-452 call shape): pool method when n <= setsize, else the
 * rejection-set method; _randbelow(n) = (w0*n)>>32, random() = 53-bit as everywhere else.
 * mode 0 exploit (archetypes verbatim), 1 explore (n_preserve cards sampled from the archetype + random rest),
 * 2 balance (two random() < q draws first, then archetype or a fully random deck per side),
 * 3 fully random decks of the given factions.  decks: card ids [2][12]. */
typedef struct { uint32_t lo, hi, gen, draw; } DeckRng;
static uint32_t deck_w(DeckRng *r, uint32_t *w1) {
  uint32_t w[4];
  philox(r->draw++, r->gen, 0xDEC4u, 0, r->lo, r->hi, w);
  if (w1) *w1 = w[1];
  return w[0];
}
static int deck_below(DeckRng *r, int n) { return (int)(((uint64_t)deck_w(r, 0) * (uint64_t)n) >> 32); }
static double deck_random(DeckRng *r) {
  uint32_t w1, w0 = deck_w(r, &w1);
  return ((double)(w0 >> 5) * 67108864.0 + (double)(w1 >> 6)) / 9007199254740992.0;
}
static void py_sample(DeckRng *r, const uint8_t *pop, int n, int k, uint8_t *out) {
  int setsize = 21;
  if (k > 5) { int x = 3 * k, p = 1; while (p < x) p *= 4; setsize += p; }  /* 4 ** ceil(log(3k, 4)) */
  if (n <= setsize) {
    uint8_t pool[128];
    memcpy(pool, pop, (size_t)n);
    for (int i = 0; i < k; i++) { int j = deck_below(r, n - i); out[i] = pool[j]; pool[j] = pool[n - i - 1]; }
  } else {
    uint8_t taken[128];
    memset(taken, 0, sizeof taken);
    for (int i = 0; i < k; i++) {
      int j = deck_below(r, n);
      while (taken[j]) j = deck_below(r, n);
      taken[j] = 1; out[i] = pop[j];
    }
  }
}
static int faction_pool(int faction, uint8_t *pool) { /* dir(cards) order == card index order; own faction + NEUTRAL */
  int n = 0;
  for (int c = 1; c <= 112; c++) if (OCARDS[c].faction == faction || OCARDS[c].faction == 0) pool[n++] = (uint8_t)c;
  return n;
}
static void random_deck(DeckRng *r, int faction, const uint8_t *original, int n_preserve, uint8_t *deck) {
  uint8_t pool[128];
  if (n_preserve >= 12) { memcpy(deck, original, 12); return; }  /* preserve_ratio == 1.0 */
  if (n_preserve > 0) py_sample(r, original, 12, n_preserve, deck); else n_preserve = 0;
  int n = faction_pool(faction, pool);
  py_sample(r, pool, n, 12 - n_preserve, deck + n_preserve);
}
void sbo_generate_decks(uint64_t seed, uint32_t generation, int mode, int n_preserve, double q, const uint8_t *archetypes,
                        const uint8_t *factions, uint8_t *decks) {
  DeckRng r = {(uint32_t)seed, (uint32_t)(seed >> 32), generation, 0};
  if (mode == 0) { memcpy(decks, archetypes, 24); return; }
  if (mode == 1) { for (int o = 0; o < 2; o++) random_deck(&r, factions[o], archetypes + 12 * o, n_preserve, decks + 12 * o); return; }
  if (mode == 2) {
    int keep[2];
    keep[0] = deck_random(&r) < q;
    keep[1] = deck_random(&r) < q;
    for (int o = 0; o < 2; o++) {
      if (keep[o]) memcpy(decks + 12 * o, archetypes + 12 * o, 12);
      else random_deck(&r, factions[o], archetypes + 12 * o, 0, decks + 12 * o);
    }
    return;
  }
  for (int o = 0; o < 2; o++) random_deck(&r, factions[o], archetypes, 0, decks + 12 * o);
}
