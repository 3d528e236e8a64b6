// sbw_core.cuh -- warp-per-game Stormbound rules engine for sm_100a.
//
// One game per WARP; the working set (struct WG, ~2.3 KB) lives in SHARED memory and all 32 lanes walk the rules together
// (vocabulary: sbw_warp.cuh).  Control flow is uniform across the warp, so there is no divergence inside a game; the scans
// that the thread-per-game engine (sb_engine.cuh) runs as loops are lane-parallel here:
//   target queries      side / kind by tile-mask algebra, strength / tribe / status filters by one lane-per-entity pass + REDUX
//   weighted draw       lane per deck card: weight lookup, FP64 prefix scan by shuffles, ballot against u
//   discard / reweight  lane per deck card
//   board flip, pool compaction, pack / unpack   lane per tile
// Entities keep an identity beyond tile occupancy (pool of 48 slots + tile -> slot map: "ghost" units), stored
// as structure-of-arrays so that lane l reads slot l without bank conflicts.  Target lists are (tile mask, scan direction,
// base bits) in registers instead of arrays.  The reference's Python recursion (ability -> damage -> death trigger -> ...)
// is real device recursion with uniform control flow; frames hold a handful of registers.
//
// Semantics are those of sb_engine.cuh function by function (which is pinned to the reference through the oracle and
// the golden fixtures); reference lines are cited per function (paths relative to the reference checkout).
#pragma once
#include "sb_defs.h"
#include "sbw_warp.cuh"

// entity flags (EF_ONB: board[pos] == this slot, kept by the three board writers)
#define WEF_OWNER 1
#define WEF_STRUCT 2
#define WEF_FIXED 4
#define WEF_SINGLE 8
#define WEF_RPLAY 16
#define WEF_ONB 32

struct __align__(8) WCard { u8 card; i8 cost; u8 flags; i8 link; u16 wn; i16 xstr; };  // hand / deck record, one 64-bit move
struct __align__(8) WMem { i8 b005; u8 pos, card, fl; i16 strength; u8 st[5]; i8 parent; u8 pad[4]; };  // cards/b005.py memories
struct WPly {  // player.py:13-37
  i16 base, max_mana, mana;
  i8 front_line;
  u8 replacable, leftmost, n_hand, n_deck, faction;
  u8 pad[4];
  WCard hand[8];   // HAND_W = 6 used
  WCard deck[24];  // DECK_W = 20 used
};
struct __align__(16) WG {
  // entity pool, structure of arrays (unit.py:8-23 / structure.py:8-16; statics live in DCard)
  i16 e_str[MAXE];
  i16 e_dmg[MAXE];
  u32 e_st[MAXE];  // five 6-bit counters, StatusEffect order (a multiset: unit.py:239-275)
  u8 e_card[MAXE], e_fl[MAXE], e_pos[MAXE], e_mid[MAXE], e_plen[MAXE];
  __align__(8) u8 e_path[MAXE][MAXPATH];  // (y+1)*4 + x, y in -1..5
  __align__(8) i8 board[24];              // tile -> slot or -1 (20 used)
  WPly pl[2];
  u32 seed_lo, seed_hi;
  u32 occ, own1, strc;  // occupied tiles / tiles of order 1 / structure tiles (bits of empty tiles: don't care)
  u16 turn, draw, steps;
  u8 local_order, current_order, phase, err, done;
  i8 player_sign;
  u8 hist_n, hist_card[4], hist_owner[4];
  u8 n_ent, n_trig, resolving, depth, n_mem, n_obj, maybe_badobs;
  u8 trig[MAXTRIG];  // entity id | has_source << 7
  __align__(8) WMem mem[NMEM];
  // ---- everything above is the game (w_copy_game moves it); below: scratch, never live across a call that may recurse
  u32 lm[SB_MASK_WORDS];
  i8 scr[24];
  i8 tn_ids[24];  // to_next_turn's snapshot of entity ids
  u8 remap[MAXE];
  __align__(8) double scr_d[DECK_W];
  double feat[SB_N_FEATURES];  // w_features() of this game (heuristic agent)
  const DCard* cards;  // shared-memory copy of the card table
  const double* wt;    // f^n(1) table in global memory
};

#define WERR(wg, code) do { if (!(wg)->err) (wg)->err = (code); } while (0)
SBW_FI const DCard& WCARD(const WG* wg, int card) { const DCard* c = wg->cards; W_SHARED(c); return c[card]; }
SBW_FI int wpt_x(int pt) { return pt >= 20 ? -1 : (pt & 3); }
SBW_FI int wpt_y(int pt) { return pt == PT_BASE_REMOTE ? -1 : pt == PT_BASE_LOCAL ? 5 : (pt >> 2); }
SBW_FI bool w_valid_xy(int x, int y) { return (unsigned)x <= 3u && (unsigned)y <= 4u; }
SBW_FI int w_owner(const WG* wg, int id) { return wg->e_fl[id] & WEF_OWNER; }
SBW_FI bool w_is_struct(const WG* wg, int id) { return (wg->e_fl[id] & WEF_STRUCT) != 0; }
SBW_FI int w_st(const WG* wg, int id, int s) { return (int)((wg->e_st[id] >> (SB_ST_BITS * s)) & 63u); }
SBW_FI int w_ex(const WG* wg, int id) { return wg->e_pos[id] & 3; }
SBW_FI int w_ey(const WG* wg, int id) { return wg->e_pos[id] >> 2; }

// ---------------------------------------------------------------- Philox4x32-10 counter stream
SBW_FI void w_philox(u32 c0, u32 c1, u32 c2, u32 c3, u32 k0, u32 k1, u32& o0, u32& o1) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    u32 h0 = w_mulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
    u32 h1 = w_mulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
    u32 n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
    c0 = n0; c1 = l1; c2 = n2; c3 = l0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  o0 = c0; o1 = c1;
}
SBW_NI int w_rng_below(WG* wg, int n) {
  W_SHARED(wg);
  if (n <= 0) { WERR(wg, SB_ERR_EMPTY_CHOICE); return 0; }
  u32 w0, w1;
  w_philox(wg->draw, wg->turn, 0, 0, wg->seed_lo, wg->seed_hi, w0, w1);
  wg->draw++;
  return (int)w_mulhi(w0, (u32)n);
}
SBW_NI double w_rng_random(WG* wg) {
  W_SHARED(wg);
  u32 w0, w1;
  w_philox(wg->draw, wg->turn, 0, 0, wg->seed_lo, wg->seed_hi, w0, w1);
  wg->draw++;
  return d_mul(d_add(d_mul((double)(w0 >> 5), 67108864.0), (double)(w1 >> 6)), 0x1.0p-53);
}
SBW_FI u32 w_agent_pick(u32 seed_lo, u32 seed_hi, u32 step, u32 n) {
  u32 w0, w1;
  w_philox(step, 0, 0xA6E7u, 0, seed_lo, seed_hi, w0, w1);
  return w_mulhi(w0, n);
}
// in-place Fisher-Yates from the end over a scratch list in shared memory (numpy RandomState.shuffle call shape)
SBW void w_shuffle(WG* wg, i8* a, int n) {
#pragma unroll 1
  for (int i = n - 1; i > 0; i--) {
    int j = w_rng_below(wg, i + 1);
    i8 t = a[i]; a[i] = a[j]; a[j] = t;
  }
}

// ---------------------------------------------------------------- board access (board.py:58-92)
SBW_FI int w_opponent_of(const WG* wg, int order) { return order == 0 ? 1 - wg->local_order : wg->local_order; }  // player.py:42-44 (Q3)
SBW_FI int w_at_xy(const WG* wg, int x, int y) { return w_valid_xy(x, y) ? wg->board[y * 4 + x] : -1; }
SBW_FI int w_at_pt(const WG* wg, int pt) { return (unsigned)pt < 20u ? wg->board[pt] : -1; }
SBW_FI void w_set_xy(WG* wg, int x, int y, int id) {
  const int t = y * 4 + x;
  const u32 b = 1u << t;
  const int prev = wg->board[t];
  if (prev >= 0 && prev != id) wg->e_fl[prev] &= (u8)~WEF_ONB;  // an overwritten occupant is no longer on the board (Q7, Q13)
  wg->board[t] = (i8)id;
  if (id >= 0) {
    const u8 fl = wg->e_fl[id] | WEF_ONB;
    wg->e_fl[id] = fl;
    wg->e_pos[id] = (u8)t;
    wg->occ |= b;
    wg->own1 = (fl & WEF_OWNER) ? (wg->own1 | b) : (wg->own1 & ~b);
    wg->strc = (fl & WEF_STRUCT) ? (wg->strc | b) : (wg->strc & ~b);
  } else {
    wg->occ &= ~b;
  }
}
// board.set(self.position, None): by the entity's (possibly stale) position, whoever stands there (Q21)
SBW_FI void w_clear_at(WG* wg, int id) {
  const int t = wg->e_pos[id];
  const int victim = wg->board[t];
  if (victim >= 0) wg->e_fl[victim] &= (u8)~WEF_ONB;
  wg->board[t] = -1;
  wg->occ &= ~(1u << t);
}
SBW_FI int w_next_tile(u32& m, bool ascending) {
  int t = ascending ? w_ffs(m) - 1 : 31 - w_clz(m);
  m &= ~(1u << t);
  return t;
}
SBW_FI void w_calc_front_line(WG* wg, int order) {  // board.py:78-92
  const bool local = order == wg->local_order;
  int fl = local ? 4 : 0;
  const u32 m = wg->occ & (order ? wg->own1 : ~wg->own1);
  if (m) {
    const int y = (local ? w_ffs(m) - 1 : 31 - w_clz(m)) >> 2;
    fl = local ? (y > 1 ? y : 1) : (y < 3 ? y : 3);
  }
  wg->pl[order].front_line = (i8)fl;
}

// ---------------------------------------------------------------- target queries (board.py:147-296)
// A query (target.py Target) in two registers.
struct TQ { u32 a, b; };  // a: kind 0-1 | side 2-3 | status 4-8 | xstatus 9-13 | has_limit 14 | base 15 | types 16-31;  b: xtypes 0-15 | limit 16-31
SBW_FI TQ w_mkT(int kind, int side) { TQ t; t.a = (u32)kind | ((u32)side << 2); t.b = 0; return t; }
SBW_FI TQ tq_types(TQ t, u32 types) { t.a |= types << 16; return t; }
SBW_FI TQ tq_xtypes(TQ t, u32 x) { t.b |= x & 0xFFFFu; return t; }
SBW_FI TQ tq_status(TQ t, u32 s) { t.a |= (s & 31u) << 4; return t; }
SBW_FI TQ tq_xstatus(TQ t, u32 s) { t.a |= (s & 31u) << 9; return t; }
SBW_FI TQ tq_limit(TQ t, int limit) { t.a |= 1u << 14; t.b = (t.b & 0xFFFFu) | ((u32)(u16)(i16)limit << 16); return t; }
SBW_FI TQ tq_base(TQ t) { t.a |= 1u << 15; return t; }
SBW_FI TQ w_card_target(const DCard& c) {
  TQ t = w_mkT(c.t_ks & 3, c.t_ks >> 2);
  t = tq_types(t, c.t_types);
  t = tq_xtypes(t, (u32)c.t_xtypes | ((c.flags & DCF_TNONHERO) ? (1u << UT_HERO) : 0u));
  t = tq_status(t, c.t_status);
  t = tq_xstatus(t, c.t_xstatus);
  if (c.t_limit >= 0) t = tq_limit(t, c.t_limit);
  if (c.flags & DCF_TBASE) t = tq_base(t);
  return t;
}
// A target list: tiles in scan order (ascending when pov == local: y=0..4, x=0..3; descending otherwise, board.py:157-158),
// then the friendly base, then the enemy base (board.py:191-199).
struct TL { u32 m; u32 meta; };  // meta: bit 0 ascending | bit 1 friendly base present | bit 2 enemy base present | bit 3 pov is local
SBW_FI int tl_n(TL t) { return w_popc(t.m) + (int)((t.meta >> 1) & 1u) + (int)((t.meta >> 2) & 1u); }
SBW_FI int tl_friendly_base(TL t) { return (t.meta & 8u) ? PT_BASE_LOCAL : PT_BASE_REMOTE; }
SBW_FI int tl_enemy_base(TL t) { return (t.meta & 8u) ? PT_BASE_REMOTE : PT_BASE_LOCAL; }
SBW_FI int tl_pop(TL& t) {  // first remaining point in list order (PT_NONE when empty)
  if (t.m) return w_next_tile(t.m, (t.meta & 1u) != 0);
  if (t.meta & 2u) { t.meta &= ~2u; return tl_friendly_base(t); }
  if (t.meta & 4u) { t.meta &= ~4u; return tl_enemy_base(t); }
  return PT_NONE;
}
SBW_FI int tl_nth(TL t, int k) {  // k-th point in list order
#pragma unroll 1
  for (int q = 0; q < k; q++) tl_pop(t);
  return tl_pop(t);
}
SBW_FI bool tl_has(TL t, int pt) {
  if ((unsigned)pt < 20u) return (t.m >> pt) & 1u;
  return ((t.meta & 2u) && pt == tl_friendly_base(t)) || ((t.meta & 4u) && pt == tl_enemy_base(t));
}
SBW_FI TL tl_tiles(u32 m, bool ascending) { TL t; t.m = m; t.meta = ascending ? 1u : 0u; return t; }

// tile mask of the on-board entities whose slot satisfies pred(slot): one lane per entity, one REDUX per 32 slots
template <class F> SBW_FI u32 w_tilemask(const WG* wg, F pred) {
  u32 m = 0;
  const int ne = wg->n_ent;
#pragma unroll 1
  for (int base = 0; base < ne; base += 32)
    m |= w_or([&](int l) -> u32 {
      const int s = base + l;
      return (s < ne && (wg->e_fl[s] & WEF_ONB) && pred(s)) ? (1u << wg->e_pos[s]) : 0u;
    });
  return m;
}
// region = tile bitmask a target must belong to (0xFFFFF = whole board).  base_passes: base points survive the region
// filter when include_base (board.py:215,262,276,294).
SBW_NI TL w_targets_region(const WG* wg, int pov, TQ t, int exclude_pt, u32 region, bool base_passes) {
  W_SHARED(wg);
  const bool pov_local = (pov == wg->local_order);
  const int kind = t.a & 3, side = (t.a >> 2) & 3;
  u32 m = wg->occ & region;
  if ((unsigned)exclude_pt < 20u) m &= ~(1u << exclude_pt);
  if (side != TS_ANY) {
    const u32 mine = pov ? wg->own1 : ~wg->own1;
    m &= (side == TS_FRIENDLY) ? mine : ~mine;
  }
  if (kind == TK_UNIT) m &= ~wg->strc; else if (kind == TK_STRUCTURE) m &= wg->strc;
  if (m) {
    const bool has_limit = (t.a >> 14) & 1u;
    const int limit = (i16)(t.b >> 16);
    const u32 want_types = t.a >> 16, bad_types = t.b & 0xFFFFu;
    const u32 want_status = (t.a >> 4) & 31u, bad_status = (t.a >> 9) & 31u;
    const DCard* cards = wg->cards;
    W_SHARED(cards);
    m &= w_tilemask(wg, [&](int s) -> bool {
      const int str = wg->e_str[s];
      if (str <= 0) return false;  // board.py:164
      if (has_limit && str > limit) return false;
      if (!(wg->e_fl[s] & WEF_STRUCT)) {  // structures only honour strength_limit (board.py:179)
        if (want_types | bad_types) {
          const u32 types = cards[wg->e_card[s]].types;
          if ((want_types && !(types & want_types)) || (types & bad_types)) return false;
        }
        if (want_status | bad_status) {
          const u32 w = wg->e_st[s];
          u32 have = 0;
#pragma unroll
          for (int k = 0; k < 5; k++) have |= (((w >> (SB_ST_BITS * k)) & 63u) ? 1u : 0u) << k;
          if ((want_status && !(have & want_status)) || (have & bad_status)) return false;
        }
      }
      return true;
    });
  }
  TL out;
  out.m = m;
  out.meta = (pov_local ? 9u : 0u);
  if (((t.a >> 15) & 1u) && base_passes) {
    const int friendly = pov_local ? PT_BASE_LOCAL : PT_BASE_REMOTE;
    const int enemy = pov_local ? PT_BASE_REMOTE : PT_BASE_LOCAL;
    if (side != TS_ENEMY && friendly != exclude_pt) out.meta |= 2u;
    if (side != TS_FRIENDLY && enemy != exclude_pt) out.meta |= 4u;
  }
  return out;
}
SBW_FI TL w_targets(const WG* wg, int pov, TQ t, int exclude_pt) { return w_targets_region(wg, pov, t, exclude_pt, 0xFFFFFu, true); }
SBW_FI u32 w_border_mask(int x, int y) {
  u32 m = 0;
  if (x > 0) m |= 1u << (y * 4 + x - 1);
  if (x < 3) m |= 1u << (y * 4 + x + 1);
  if (y > 0) m |= 1u << (y * 4 + x - 4);
  if (y < 4) m |= 1u << (y * 4 + x + 4);
  return m;
}
SBW_FI u32 w_surround_mask(int x, int y) {
  u32 row = (x > 0 ? 1u << (x - 1) : 0u) | (1u << x) | (x < 3 ? 1u << (x + 1) : 0u);
  u32 m = (row & ~(1u << x)) << (y * 4);
  if (y > 0) m |= row << (y * 4 - 4);
  if (y < 4) m |= row << (y * 4 + 4);
  return m;
}
// Unfiltered neighbourhood lists keep the reference's own order (board.py:236-296), which is not tile order: a small
// packed list, one byte per point.  side [L,R]; bordering [L,R,y-1,y+1]; surrounding [L, L y-1, L y+1, R, R y-1, R y+1, y-1, y+1].
struct PL { u64 v; int n; };
SBW_FI void pl_push(PL& p, int pt) { p.v |= (u64)(u8)pt << (8 * p.n); p.n++; }
SBW_FI int pl_get(PL p, int k) { return (int)((p.v >> (8 * k)) & 0xFFu); }
SBW_FI PL w_side_list(int x, int y) {
  PL p; p.v = 0; p.n = 0;
  if (x > 0) pl_push(p, y * 4 + x - 1);
  if (x < 3) pl_push(p, y * 4 + x + 1);
  return p;
}
SBW_FI PL w_border_list(int x, int y) {
  PL p = w_side_list(x, y);
  if (y > 0) pl_push(p, y * 4 + x - 4);
  if (y < 4) pl_push(p, y * 4 + x + 4);
  return p;
}
SBW_FI PL w_surround_list(int x, int y) {
  PL p; p.v = 0; p.n = 0;
#pragma unroll 1
  for (int dx = -1; dx <= 1; dx += 2) {
    int xx = x + dx;
    if (xx < 0 || xx > 3) continue;
    pl_push(p, y * 4 + xx);
    if (y > 0) pl_push(p, y * 4 + xx - 4);
    if (y < 4) pl_push(p, y * 4 + xx + 4);
  }
  if (y > 0) pl_push(p, y * 4 + x - 4);
  if (y < 4) pl_push(p, y * 4 + x + 4);
  return p;
}
SBW_FI PL pl_empty_of(const WG* wg, PL in) {
  PL o; o.v = 0; o.n = 0;
#pragma unroll 1
  for (int i = 0; i < in.n; i++) { int pt = pl_get(in, i); if (w_at_pt(wg, pt) < 0) pl_push(o, pt); }
  return o;
}
// column toward the enemy of pov (front) or toward pov's own base (behind), nearest first (board.py:206-234)
SBW_FI u32 w_column_region(const WG* wg, int x, int y, int pov, bool front, bool& up) {
  const bool pov_local = (pov == wg->local_order);
  up = (pov_local == front);  // decreasing y
  const u32 col = 0x11111u << x;
  return up ? (col & ((1u << (y * 4)) - 1u)) : (col & ~((1u << (y * 4 + 4)) - 1u) & 0xFFFFFu);
}
// filtered: get_targets order restricted to the column, then a stable sort on y -- within one column that is simply the
// scan direction "nearest first" (no card asks a column query for bases)
SBW_FI TL w_column_targets(const WG* wg, int x, int y, int pov, TQ t, bool front) {
  bool up;
  const u32 region = w_column_region(wg, x, y, pov, front, up);
  TL r = w_targets_region(wg, pov, t, PT_NONE, region, true);
  r.meta = (r.meta & ~1u) | (up ? 0u : 1u);  // up: nearest = highest tile index first
  return r;
}
SBW_FI int w_column_first_tile(const WG* wg, int x, int y, int pov, bool front) {  // first plain tile of the column list, or PT_NONE
  const bool pov_local = (pov == wg->local_order);
  const bool up = (pov_local == front);
  const int ny = up ? y - 1 : y + 1;
  return (ny >= 0 && ny <= 4) ? ny * 4 + x : PT_NONE;
}
SBW_FI TL w_bordering_t(const WG* wg, int x, int y, int pov, TQ t) { return w_targets_region(wg, pov, t, PT_NONE, w_border_mask(x, y), true); }
SBW_FI TL w_surrounding_t(const WG* wg, int x, int y, int pov, TQ t) { return w_targets_region(wg, pov, t, PT_NONE, w_surround_mask(x, y), true); }
SBW_FI bool w_within_front_line(const WG* wg, int order, int y) {  // player.py:96-100 (Q22)
  return order == 0 ? y >= wg->pl[order].front_line : y <= wg->pl[order].front_line;
}
// player.py:102-111 as a tile list: FIRST rows front_line..4 ascending, SECOND rows front_line..0 with x descending
SBW_FI TL w_within_front_line_tiles(const WG* wg, int order) {
  const int fl = wg->pl[order].front_line;
  u32 m;
  if (order == 0) m = fl > 4 ? 0u : (fl < 0 ? 0xFFFFFu : (0xFFFFFu & ~((1u << (fl * 4)) - 1u)));
  else m = fl < 0 ? 0u : (fl > 4 ? 0xFFFFFu : ((1u << (fl * 4 + 4)) - 1u));
  return tl_tiles(m, order == 0);
}

// ---------------------------------------------------------------- entities
SBW_NI int w_new_ent(WG* wg, int card, int owner, int strength) {
  W_SHARED(wg);
  if (wg->n_ent >= MAXE) { WERR(wg, SB_ERR_OVERFLOW); return MAXE - 1; }
  const int id = wg->n_ent;
  wg->n_ent = (u8)(id + 1);
  const DCard& c = WCARD(wg, card);
  wg->e_card[id] = (u8)card;
  wg->e_fl[id] = (u8)((owner ? WEF_OWNER : 0) | (c.kind == KIND_STRUCTURE ? WEF_STRUCT : 0) | ((c.flags & DCF_FIXED) ? WEF_FIXED : 0));
  wg->e_str[id] = (i16)strength; wg->e_dmg[id] = 0;
  wg->e_st[id] = 0;
  wg->e_mid[id] = 0; wg->e_pos[id] = 0; wg->e_plen[id] = 0;
  return id;
}
SBW_NI int w_spawn_token_unit(WG* wg, int owner, int pt, int strength, int type) {  // board.py:298-311
  W_SHARED(wg);
  int id = w_new_ent(wg, SBC_TOKEN_UNIT0 + type, owner, strength);
  w_set_xy(wg, wpt_x(pt), wpt_y(pt), id);
  w_calc_front_line(wg, owner);
  return id;
}

// forward declarations of the mutually recursive core
SBW_NI void w_ability(WG* wg, int id, int pos_pt, int has_source);
SBW_NI void w_effect(WG* wg, int id, int pos_pt, int has_source);
SBW_NI void w_spell_effect(WG* wg, int card, int caster, int pos_pt);
SBW_NI void w_unit_move(WG* wg, int id);
SBW_NI void w_player_play(WG* wg, int order, int index, int pos_pt);

// ---------------------------------------------------------------- trigger stack (board.py:46-56, card.py:48-62)
SBW_FI void w_push_trigger(WG* wg, int id, int has_source) {
  if (wg->n_trig >= MAXTRIG) { WERR(wg, SB_ERR_OVERFLOW); return; }
  const int n = wg->n_trig;
  wg->trig[n] = (u8)(id | (has_source ? 0x80 : 0));
  wg->n_trig = (u8)(n + 1);
}
SBW_FI void w_pop_trigger(WG* wg) {
  if (wg->n_trig == 0 || wg->resolving) return;
  const int n = wg->n_trig - 1;
  wg->n_trig = (u8)n;
  const u8 t = wg->trig[n];
  w_ability(wg, t & 0x7F, PT_NONE, t >> 7);
}
SBW_NI void w_ability(WG* wg, int id, int pos_pt, int has_source) {
  W_SHARED(wg);
  if (!(WCARD(wg, wg->e_card[id]).flags & DCF_ABILITY)) return;  // un-overridden Card.activate_ability: no wrapper
  if (wg->depth > MAXDEPTH) { WERR(wg, SB_ERR_DEPTH); return; }
  wg->depth++;
  wg->resolving = 1;
  w_effect(wg, id, pos_pt, has_source);
  wg->resolving = 0;
  w_pop_trigger(wg);
  wg->depth--;
}
SBW_FI void w_spell_ability(WG* wg, int card, int caster, int pos_pt) {
  if (wg->depth > MAXDEPTH) { WERR(wg, SB_ERR_DEPTH); return; }
  wg->depth++;
  wg->resolving = 1;
  w_spell_effect(wg, card, caster, pos_pt);
  wg->resolving = 0;
  w_pop_trigger(wg);
  wg->depth--;
}

// ---------------------------------------------------------------- status verbs (unit.py:239-275)
SBW_FI void w_st_add(WG* wg, int id, int s) {  // 6-bit packed counters: the 64th copy of one status flags the game
  if (w_st(wg, id, s) < 63) wg->e_st[id] += 1u << (SB_ST_BITS * s); else WERR(wg, SB_ERR_OVERFLOW);
}
SBW_FI void w_st_remove(WG* wg, int id, int s) {
  if (w_st(wg, id, s)) wg->e_st[id] -= 1u << (SB_ST_BITS * s); else WERR(wg, SB_ERR_INDEX);
}
SBW_FI void wv_freeze(WG* wg, int id) { w_st_add(wg, id, SB_ST_FROZEN); }
SBW_FI void wv_poison(WG* wg, int id) { if (w_st(wg, id, SB_ST_VITALIZED)) w_st_remove(wg, id, SB_ST_VITALIZED); w_st_add(wg, id, SB_ST_POISONED); }
SBW_FI void wv_vitalize(WG* wg, int id) { if (w_st(wg, id, SB_ST_POISONED)) w_st_remove(wg, id, SB_ST_POISONED); w_st_add(wg, id, SB_ST_VITALIZED); }
SBW_FI void wv_confuse(WG* wg, int id) { w_st_add(wg, id, SB_ST_CONFUSED); }
SBW_FI void wv_disable(WG* wg, int id) { if (WCARD(wg, wg->e_card[id]).flags & DCF_ABILITY) w_st_add(wg, id, SB_ST_DISABLED); }
SBW_FI void wv_heal(WG* wg, int id, int amount) { wg->e_str[id] = (i16)(wg->e_str[id] + amount); }

// ---------------------------------------------------------------- damage (unit.py:205-231, structure.py:52-69, player.py:83-88)
SBW_FI int w_player_damage(WG* wg, int order, int amount) { wg->pl[order].base = (i16)(wg->pl[order].base - amount); return amount; }
SBW_NI void w_destroy(WG* wg, int id, int has_source) {
  W_SHARED(wg);
  w_clear_at(wg, id);  // by (possibly stale) position, like board.set(self.position, None) (Q21)
  wg->e_dmg[id] = wg->e_str[id];
  if (!w_is_struct(wg, id)) {
    wg->e_plen[id] = 0;
    if (WCARD(wg, wg->e_card[id]).trigger == TR_ON_DEATH) { w_push_trigger(wg, id, has_source); w_pop_trigger(wg); }
  }
  w_calc_front_line(wg, w_opponent_of(wg, wg->current_order));
}
SBW_NI int w_deal_damage(WG* wg, int id, int amount, int pending, int has_source) {
  W_SHARED(wg);
  const int s0 = wg->e_str[id];
  if (s0 - amount < 0) amount = s0;
  wg->e_dmg[id] = (i16)amount;
  const int s1 = s0 - amount;
  wg->e_str[id] = (i16)s1;
  if (!pending && s1 <= 0) w_destroy(wg, id, has_source);
  else if (!w_is_struct(wg, id) && s1 > 0 && WCARD(wg, wg->e_card[id]).trigger == TR_AFTER_SURVIVING) {
    w_push_trigger(wg, id, has_source);
    w_pop_trigger(wg);
  }
  return amount;
}
SBW_FI int w_deal_damage_pt(WG* wg, int pt, int amount, int has_source) {  // board.at(point).deal_damage(...)
  if (pt == PT_BASE_LOCAL) return w_player_damage(wg, wg->local_order, amount);
  if (pt == PT_BASE_REMOTE) return w_player_damage(wg, 1 - wg->local_order, amount);
  int id = w_at_pt(wg, pt);
  if (id < 0) { WERR(wg, SB_ERR_NONE_TARGET); return 0; }
  return w_deal_damage(wg, id, amount, 0, has_source);
}

// ---------------------------------------------------------------- movement (unit.py:66-203, 277-382)
SBW_FI u8 w_enc_xy(int x, int y) { return (u8)((y + 1) * 4 + x); }
SBW_NI void w_set_path(WG* wg, int id, int on_play, int extra_movement) {  // unit.py:78-122
  W_SHARED(wg);
  u64 dest = 0;  // up to MAXPATH encoded destinations, one byte each
  int nd = 0;
  int px = w_ex(wg, id), py = w_ey(wg, id);
  int confused_cached = w_st(wg, id, SB_ST_CONFUSED);
  const int fl = wg->e_fl[id];
  const int owner = fl & WEF_OWNER;
  const bool is_local = owner == wg->local_order;
  int steps = on_play ? WCARD(wg, wg->e_card[id]).movement + extra_movement : 1;
  if (steps > MAXPATH) { WERR(wg, SB_ERR_OVERFLOW); steps = MAXPATH; }
  const u32 mine = owner ? wg->own1 : ~wg->own1;
#pragma unroll 1
  for (int i = 0; i < steps; i++) {
    int dx = px, dy = py + (is_local ? -1 : 1);
    if (confused_cached > 0) {
      int delta;
      if (px == 0) { w_rng_below(wg, 1); delta = 1; }
      else if (px == 3) { w_rng_below(wg, 1); delta = -1; }
      else delta = w_rng_below(wg, 2) == 0 ? -1 : 1;
      dx = px + delta; dy = py;
      confused_cached--;
    } else if (on_play && !(fl & WEF_FIXED) && dy != (is_local ? -1 : 5)) {
      const u32 occ = wg->occ;
      const bool ahead_free_or_own = !w_valid_xy(dx, dy) || !((occ >> (dy * 4 + dx)) & 1u) || ((mine >> (dy * 4 + dx)) & 1u);
      if (ahead_free_or_own) {
        const int t = py * 4 + px;
        bool left_ok = px > 0 && w_valid_xy(px - 1, py) && ((occ & ~mine) >> (t - 1) & 1u);
        bool right_ok = px < 3 && w_valid_xy(px + 1, py) && ((occ & ~mine) >> (t + 1) & 1u);
        const u8 lenc = w_enc_xy(px - 1, py), renc = w_enc_xy(px + 1, py);
#pragma unroll 1
        for (int k = 0; k < nd; k++) {
          const u8 d = (u8)(dest >> (8 * k));
          if (d == lenc) left_ok = false;
          if (d == renc) right_ok = false;
        }
        if (px <= 1) { if (right_ok) { dx = px + 1; dy = py; } else if (left_ok) { dx = px - 1; dy = py; } }
        else { if (left_ok) { dx = px - 1; dy = py; } else if (right_ok) { dx = px + 1; dy = py; } }
      }
    }
    dest |= (u64)w_enc_xy(dx, dy) << (8 * nd);
    nd++;
    px = dx; py = dy;
  }
  *reinterpret_cast<u64*>(wg->e_path[id]) = dest;
  wg->e_plen[id] = (u8)nd;
}
// the common case of unit.py:78-122 inline: one step straight ahead (turn start / command, unit not confused)
SBW_FI void w_set_path_step(WG* wg, int id) {
  if (w_st(wg, id, SB_ST_CONFUSED)) { w_set_path(wg, id, 0, 0); return; }
  const int t = wg->e_pos[id];
  const bool is_local = w_owner(wg, id) == wg->local_order;
  wg->e_path[id][0] = (u8)(t + 4 + (is_local ? -4 : 4));  // enc = (y + 1) * 4 + x of the tile one row ahead
  wg->e_plen[id] = 1;
}
SBW_NI void w_unit_move(WG* wg, int id) {  // unit.py:124-203
  W_SHARED(wg);
  if (wg->depth > MAXDEPTH) { WERR(wg, SB_ERR_DEPTH); return; }
  wg->depth++;
  const int trig = WCARD(wg, wg->e_card[id]).trigger;
  const u8 current_id = (u8)(wg->e_mid[id] + 1);
  wg->e_mid[id] = current_id;
  if (wg->phase == PH_TURN_START) {
    if (w_st(wg, id, SB_ST_POISONED)) w_deal_damage(wg, id, 1, 0, 0);
    else if (w_st(wg, id, SB_ST_VITALIZED)) wv_heal(wg, id, 1);
    if (w_st(wg, id, SB_ST_FROZEN)) { w_st_remove(wg, id, SB_ST_FROZEN); wg->depth--; return; }
  }
  if (wg->e_plen[id] == 0) { wg->depth--; return; }
  if (trig == TR_BEFORE_MOVING && !w_st(wg, id, SB_ST_DISABLED)) w_ability(wg, id, PT_NONE, 1);
  if (w_st(wg, id, SB_ST_FROZEN)) { wg->depth--; return; }
  const int np = wg->e_plen[id];  // `for destination in self.path` iterates the list object bound now
  const u64 path = *reinterpret_cast<const u64*>(wg->e_path[id]);
#pragma unroll 1
  for (int i = 0; i < np; i++) {
    const int enc = (int)((path >> (8 * i)) & 0xFFu);
    const int dx = enc & 3, dy = (enc >> 2) - 1;
    const int owner = w_owner(wg, id);
    if (dy < 0 || dy > 4) {  // to base
      if (trig == TR_BEFORE_ATTACKING && !w_st(wg, id, SB_ST_DISABLED)) w_ability(wg, id, -2, 1);
      const int target = dy < 0 ? 1 - wg->local_order : wg->local_order;
      w_player_damage(wg, target, wg->e_str[id]);
      if (wg->pl[target].base > 0) w_destroy(wg, id, 0);
      wg->depth--;
      return;
    }
    int tid = w_at_xy(wg, dx, dy);
    bool attacked = false;
    if (tid >= 0 && w_owner(wg, tid) == owner && dx == w_ex(wg, id)) { wg->depth--; return; }
    if (tid >= 0 && (w_st(wg, id, SB_ST_CONFUSED) || w_owner(wg, tid) != owner)) {
      if (trig == TR_BEFORE_ATTACKING && !w_st(wg, id, SB_ST_DISABLED)) w_ability(wg, id, dy * 4 + dx, 1);
      tid = w_at_xy(wg, dx, dy);
      if (tid >= 0) {
        const int tstr = wg->e_str[tid];
        const int t_pending = !w_is_struct(wg, tid) && WCARD(wg, wg->e_card[tid]).trigger == TR_ON_DEATH && !w_st(wg, tid, SB_ST_DISABLED);
        const int l_pending = trig == TR_ON_DEATH && !w_st(wg, id, SB_ST_DISABLED);
        w_deal_damage(wg, tid, wg->e_str[id], t_pending, 0);
        w_deal_damage(wg, id, tstr, l_pending, 0);
        if (wg->e_str[tid] <= 0 && t_pending) w_destroy(wg, tid, 0);
        if (wg->e_str[id] <= 0 && l_pending) w_destroy(wg, id, 0);
        attacked = true;
      }
    }
    if (current_id != wg->e_mid[id]) { wg->depth--; return; }
    if (w_at_xy(wg, dx, dy) < 0 && wg->e_str[id] > 0) {
      w_clear_at(wg, id);
      w_set_xy(wg, dx, dy, id);
      WPly& p = wg->pl[w_owner(wg, id)];
      if (p.front_line > dy) p.front_line = (i8)(dy > 1 ? dy : 1);
      if (attacked && trig == TR_AFTER_ATTACKING && !w_st(wg, id, SB_ST_DISABLED)) w_ability(wg, id, PT_NONE, 1);
      if (w_st(wg, id, SB_ST_CONFUSED)) w_st_remove(wg, id, SB_ST_CONFUSED);
    }
  }
  wg->depth--;
}
SBW_FI void w_unit_play(WG* wg, int id, int x, int y) {  // unit.py:66-76
  wg->e_fl[id] |= WEF_RPLAY;
  w_set_xy(wg, x, y, id);
  w_set_path(wg, id, 1, 0);
  if (WCARD(wg, wg->e_card[id]).trigger == TR_ON_PLAY) w_ability(wg, id, PT_NONE, 1);
  w_unit_move(wg, id);
  wg->e_fl[id] &= (u8)~WEF_RPLAY;
}
SBW_NI void w_struct_play(WG* wg, int id, int x, int y) {  // structure.py:45-50
  W_SHARED(wg);
  w_set_xy(wg, x, y, id);
  if (WCARD(wg, wg->e_card[id]).trigger == TR_ON_PLAY) w_ability(wg, id, PT_NONE, 1);
}
SBW_FI void w_gain_speed(WG* wg, int id, int amount) { w_set_path(wg, id, (wg->e_fl[id] & WEF_RPLAY) != 0, amount); }  // unit.py:277-280
SBW_NI void wv_command(WG* wg, int id) {  // unit.py:282-289
  W_SHARED(wg);
  const u8 cache = wg->e_fl[id] & WEF_FIXED;
  wg->e_fl[id] |= WEF_FIXED;
  w_set_path_step(wg, id);
  w_unit_move(wg, id);
  wg->e_fl[id] = (u8)((wg->e_fl[id] & ~WEF_FIXED) | cache);
}
SBW_NI void wv_convert(WG* wg, int id) {  // unit.py:291-293
  W_SHARED(wg);
  const int o = w_opponent_of(wg, w_owner(wg, id));
  wg->e_fl[id] = (u8)((wg->e_fl[id] & ~WEF_OWNER) | (o ? WEF_OWNER : 0));
  const int t = wg->e_pos[id];
  if (wg->board[t] == id) {  // keep the owner mask in step (a converted ghost is not on the board)
    const u32 b = 1u << t;
    wg->own1 = o ? (wg->own1 | b) : (wg->own1 & ~b);
  }
  w_set_path(wg, id, (wg->e_fl[id] & WEF_RPLAY) != 0, 0);
}
SBW_NI void wv_push(WG* wg, int id, int fx, int fy) {  // unit.py:318-339
  W_SHARED(wg);
  int dx = 0, dy = 0;
  const int ex = w_ex(wg, id), ey = w_ey(wg, id);
  if (fy < ey) dy = 1; else if (fy > ey) dy = -1; else if (fx < ex) dx = 1; else if (fx > ex) dx = -1;
  if (dx || dy) {
#pragma unroll 1
    for (;;) {
      const int nx = w_ex(wg, id) + dx, ny = w_ey(wg, id) + dy;
      if (!w_valid_xy(nx, ny)) break;
      if (wg->board[ny * 4 + nx] >= 0) return;
      w_clear_at(wg, id);
      w_set_xy(wg, nx, ny, id);
    }
  }
  WPly& p = wg->pl[w_owner(wg, id)];
  const int y = w_ey(wg, id);
  if (p.front_line > y) p.front_line = (i8)(y > 1 ? y : 1);
}
SBW_NI void wv_force_attack(WG* wg, int id, int tx, int ty) {  // unit.py:341-371
  W_SHARED(wg);
  const int ex = w_ex(wg, id), ey = w_ey(wg, id);
  if ((tx != ex && ty != ey) || w_at_xy(wg, tx, ty) < 0) return;
  u64 dest = 0;
  int nd = 0;
  const bool vertical = (tx == ex);
  const int fixed = vertical ? ex : ey, start = vertical ? ey : ex, end = vertical ? ty : tx;
  const int delta = end > start ? 1 : -1;
#pragma unroll 1
  for (int i = start + delta; i != end + delta; i += delta) {
    const int x = vertical ? fixed : i, y = vertical ? i : fixed;
    if (i != end && w_at_xy(wg, x, y) >= 0) return;
    dest |= (u64)w_enc_xy(x, y) << (8 * nd);
    nd++;
  }
  if (nd > 0) {
    *reinterpret_cast<u64*>(wg->e_path[id]) = dest;
    wg->e_plen[id] = (u8)nd;
    w_unit_move(wg, id);
  }
}
SBW_NI void wv_teleport(WG* wg, int id, int dx, int dy) {  // unit.py:373-382
  W_SHARED(wg);
  if (w_at_xy(wg, dx, dy) < 0) {
    w_clear_at(wg, id);
    w_set_xy(wg, dx, dy, id);
    WPly& p = wg->pl[w_owner(wg, id)];
    if (p.front_line > dy) p.front_line = (i8)(dy > 1 ? dy : 1);
    w_set_path(wg, id, (wg->e_fl[id] & WEF_RPLAY) != 0, 0);
  }
}

// ---------------------------------------------------------------- hand / deck (player.py:46-81)
// list.remove(target): index of the first element that `is` target or == target (see first_equal in sb_engine.cuh)
SBW_FI int w_first_equal(WG* wg, const WCard* l, int n, int idx) {
  const WCard t = l[idx];
  if (WCARD(wg, t.card).kind == KIND_SPELL) return idx;
  const int lim = idx < n ? idx : n;
  u32 same = w_ballot([&](int q) -> bool { return q < lim && l[q].card == t.card; });  // hands and decks hold at most 24 records
#pragma unroll 1
  while (same) {
    const int i = w_ffs(same) - 1;
    same &= same - 1;
    const int oi = l[i].flags & SB_CF_OBJ, ot = t.flags & SB_CF_OBJ;
    if (oi != ot) { WERR(wg, SB_ERR_NONE_TARGET); return idx; }
    if (!oi) return i;
  }
  return idx;
}
SBW_NI void w_player_draw(WG* wg, int order, int amount) {  // player.py:46-52 + numpy choice(p=) semantics
  W_SHARED(wg);
  WPly& p = wg->pl[order];
  const double* wt = wg->wt;
#pragma unroll 1
  for (int k = 0; k < amount; k++) {
    const int n = p.n_deck;
    if (n <= 0) { WERR(wg, SB_ERR_EMPTY_CHOICE); return; }
    // numpy: p_i = fl(w_i / sum), cdf = cumsum(p), cdf /= cdf[-1], idx = searchsorted(cdf, u, side='right')
    //      = #{i : fl(cdf_i / last) <= u}.
    // Fast path: lane i holds a_i = (w_0 + .. + w_i) / sum from a shuffle prefix scan.  a_i differs from the exactly
    // rounded fl(cdf_i / last) by less than 1e-14 (n <= 20 terms of a few ulp each), so the comparison with u is decided
    // whenever |a_i - u| > 1e-13 (the partial sums only grow).  Only a draw that lands closer than that to a boundary
    // (about one in 1e12) takes the exact path below, which forms every quotient like numpy does, in deck order.
    LV<double> c;
    FOR_LANES(l) LVAL(c, l) = l < n ? wt[p.deck[l].wn] : 0.0; END_LANES
    LV<double> w = c;
    w_scan_add(c);
    const double total = w_bcast(c, 31);
    const double u = w_rng_random(wg);
    const double inv = d_div(1.0, total);
    const u32 below = w_ballot([&](int l) -> bool { return l < n && LVAL(c, l) * inv - u < -1e-13; });
    u32 close_call = w_ballot([&](int l) -> bool { const double d = LVAL(c, l) * inv - u; return l < n && d >= -1e-13 && d <= 1e-13; });
    int idx = w_popc(below);
#ifdef SB_FORCE_EXACT_DRAW  // test builds: always take the exact path
    close_call = 1;
#endif
    if (close_call) {
      FOR_LANES(l) if (l < n) wg->scr_d[l] = LVAL(w, l); END_LANES
      double sum = 0.0;
#pragma unroll 1
      for (int i = 0; i < n; i++) sum = d_add(sum, wg->scr_d[i]);
      double acc = 0.0;
#pragma unroll 1
      for (int i = 0; i < n; i++) { acc = d_add(acc, d_div(wg->scr_d[i], sum)); wg->scr_d[i] = acc; }
      const double last = wg->scr_d[n - 1];
      idx = 0;
#pragma unroll 1
      for (int i = 0; i < n; i++) idx += d_div(wg->scr_d[i], last) <= u;
    }
    if (idx > n - 1) idx = n - 1;
    WCard card = p.deck[idx];
    card.wn = 0;
    if (p.n_hand >= HAND_W) { WERR(wg, SB_ERR_OVERFLOW); return; }
    const int nh = p.n_hand;
    p.hand[nh] = card;
    p.n_hand = (u8)(nh + 1);
    const int j = w_first_equal(wg, p.deck, n, idx);
    if (j != idx) p.deck[idx].wn = 0;
    u64* d64 = reinterpret_cast<u64*>(p.deck);
    LV<u64> tmp;
    FOR_LANES(l) LVAL(tmp, l) = (l >= j && l < n - 1) ? d64[l + 1] : (l < 24 ? d64[l] : 0ull); END_LANES
    FOR_LANES(l) if (l >= j && l < n - 1) d64[l] = LVAL(tmp, l); END_LANES
    p.n_deck = (u8)(n - 1);
  }
}
SBW_FI void w_player_fill_hand(WG* wg, int order) { w_player_draw(wg, order, 4 - wg->pl[order].n_hand); }
SBW_NI void w_player_discard(WG* wg, int order, int index) {  // player.py:57-66
  W_SHARED(wg);
  WPly& p = wg->pl[order];
  const int nd = p.n_deck;
  if (w_ballot([&](int l) -> bool { return l < nd && p.deck[l].wn >= WT_N - 1; })) WERR(wg, SB_ERR_OVERFLOW);
  FOR_LANES(l) if (l < nd && p.deck[l].wn < WT_N - 1) p.deck[l].wn++; END_LANES
  WCard target = p.hand[index];
  const int j = w_first_equal(wg, p.hand, p.n_hand, index);
  const int nh = p.n_hand;
#pragma unroll 1
  for (int i = j; i < nh - 1; i++) p.hand[i] = p.hand[i + 1];
  p.n_hand = (u8)(nh - 1);
  if (!(target.flags & SB_CF_SINGLE_USE)) {
    if (nd >= DECK_W) { WERR(wg, SB_ERR_OVERFLOW); return; }
    target.wn = 0;
    p.deck[nd] = target;
    p.n_deck = (u8)(nd + 1);
  }
}
SBW_NI void w_player_play(WG* wg, int order, int index, int pos_pt) {  // player.py:68-77
  W_SHARED(wg);
  WPly& p = wg->pl[order];
  if (index < 0 || index >= p.n_hand) { WERR(wg, SB_ERR_INDEX); return; }
  const WCard target = p.hand[index];
  if (wg->hist_n < 4) { const int h = wg->hist_n; wg->hist_card[h] = target.card; wg->hist_owner[h] = (u8)order; wg->hist_n = (u8)(h + 1); }
  else {
    const u8 c1 = wg->hist_card[1], c2 = wg->hist_card[2], c3 = wg->hist_card[3];
    const u8 o1 = wg->hist_owner[1], o2 = wg->hist_owner[2], o3 = wg->hist_owner[3];
    wg->hist_card[0] = c1; wg->hist_card[1] = c2; wg->hist_card[2] = c3; wg->hist_card[3] = target.card;
    wg->hist_owner[0] = o1; wg->hist_owner[1] = o2; wg->hist_owner[2] = o3; wg->hist_owner[3] = (u8)order;
  }
  w_player_discard(wg, order, index);
  const DCard& c = WCARD(wg, target.card);
  if (c.kind == KIND_SPELL) {  // spell.py:22-24
    bool ok = true;
    if (c.flags & DCF_TARGET) ok = tl_has(w_targets(wg, wg->current_order, w_card_target(c), PT_NONE), pos_pt) && pos_pt != PT_NONE;
    if (ok) w_spell_ability(wg, target.card, order, pos_pt);
    return;
  }
  if (pos_pt < 0 || pos_pt >= 20) { WERR(wg, SB_ERR_INDEX); return; }
  int strength = c.strength;
  if (target.flags & SB_CF_OBJ) strength = target.link >= 0 ? wg->e_str[target.link] : target.xstr;
  const int id = w_new_ent(wg, target.card, order, strength);  // target.copy(), player.py:74
  wg->e_fl[id] = (u8)((wg->e_fl[id] & ~(WEF_FIXED | WEF_SINGLE)) | ((target.flags & SB_CF_FIXED) ? WEF_FIXED : 0) |
                      ((target.flags & SB_CF_SINGLE_USE) ? WEF_SINGLE : 0));
  if (c.kind == KIND_UNIT) w_unit_play(wg, id, wpt_x(pos_pt), wpt_y(pos_pt));
  else w_struct_play(wg, id, wpt_x(pos_pt), wpt_y(pos_pt));
}
SBW_FI void w_player_cycle(WG* wg, int order, int index) { w_player_discard(wg, order, index); w_player_draw(wg, order, 1); }  // player.py:79-81

// ---------------------------------------------------------------- turn pipeline (board.py:94-145)
SBW_NI void w_board_flip(WG* wg) {  // board.py:94-115
  W_SHARED(wg);
  wg->local_order ^= 1;
  wg->pl[0].front_line = (i8)(4 - wg->pl[0].front_line);
  wg->pl[1].front_line = (i8)(4 - wg->pl[1].front_line);
  LV<int> b;
  FOR_LANES(l) LVAL(b, l) = l < 20 ? wg->board[19 - l] : -1; END_LANES
  FOR_LANES(l) if (l < 20) wg->board[l] = (i8)LVAL(b, l); END_LANES
  wg->occ = w_brev(wg->occ) >> 12;  // tile t -> 19 - t
  wg->own1 = w_brev(wg->own1) >> 12;
  wg->strc = w_brev(wg->strc) >> 12;
  const int ne = wg->n_ent;
#pragma unroll 1
  for (int base = 0; base < ne; base += 32) {  // only entities on the board move with it; ghosts keep their stale position
    FOR_LANES(l) { const int s = base + l; if (s < ne && (wg->e_fl[s] & WEF_ONB)) wg->e_pos[s] = (u8)(19 - wg->e_pos[s]); } END_LANES
  }
}
// entity ids of a tile list, in list order, into wg->tn_ids (the list-of-objects snapshots of board.py:137,141)
SBW_FI int w_snapshot_ids(WG* wg, TL t) {
  int n = 0;
#pragma unroll 1
  while (t.m) { const int tile = w_next_tile(t.m, (t.meta & 1u) != 0); wg->tn_ids[n++] = wg->board[tile]; }
  return n;
}
SBW_NI void w_to_next_turn(WG* wg) {  // board.py:117-145
  W_SHARED(wg);
  wg->phase = PH_TURN_END;
  w_player_fill_hand(wg, wg->current_order);
  // TURN_END structures: none of the 12 structures has that trigger (structure.py:36-38); list building has no side effect
  w_calc_front_line(wg, wg->local_order);
  w_calc_front_line(wg, 1 - wg->local_order);
  wg->pl[wg->current_order].max_mana++;
  wg->pl[0].mana = wg->pl[0].max_mana;
  wg->pl[1].mana = wg->pl[1].max_mana;
  wg->phase = PH_TURN_START;
  const int cur = (wg->current_order == wg->local_order) ? 1 - wg->local_order : wg->local_order;
  wg->current_order = (u8)cur;
  wg->pl[cur].replacable = 1;
  wg->pl[cur].leftmost = 1;
  int n = w_snapshot_ids(wg, w_targets(wg, cur, w_mkT(TK_STRUCTURE, TS_FRIENDLY), PT_NONE));
#pragma unroll 1
  for (int i = 0; i < n; i++) {
    const int id = wg->tn_ids[i];
    if (WCARD(wg, wg->e_card[id]).trigger == TR_TURN_START) w_ability(wg, id, wg->e_pos[id], 1);
  }
  n = w_snapshot_ids(wg, w_targets(wg, cur, w_mkT(TK_UNIT, TS_FRIENDLY), PT_NONE));
#pragma unroll 1
  for (int i = 0; i < n; i++) { const int id = wg->tn_ids[i]; w_set_path_step(wg, id); w_unit_move(wg, id); }  // snapshot incl. ghosts (Q21)
  wg->phase = PH_PLAY;
}

// ---------------------------------------------------------------- legal actions / step (games/stormbound.py:318-373,528-561)
SBW_FI void w_mask_set(u32* m, int a) { m[a >> 5] |= 1u << (a & 31); }
SBW_NI int w_legal_mask(WG* wg) {  // result in wg->lm; returns the number of legal actions
  W_SHARED(wg);
  u32* m = wg->lm;
  const WPly& p = wg->pl[wg->local_order];
  int n_play = 0;
#pragma unroll
  for (int i = 0; i < SB_MASK_WORDS; i++) m[i] = 0;
  // PLACE ordinals over y=4..1, x=0..3 that are empty and within the front line: ordinal = (4-y)*4 + x
  const u32 fr = ~wg->occ;
  u32 empty16 = ((fr >> 16) & 0xFu) | (((fr >> 12) & 0xFu) << 4) | (((fr >> 8) & 0xFu) << 8) | (((fr >> 4) & 0xFu) << 12);
  const int fl = p.front_line < 1 ? 1 : p.front_line;
  empty16 &= fl > 4 ? 0u : (0xFFFFu >> ((fl - 1) * 4));
  const int n_empty = w_popc(empty16);
  const int nh = p.n_hand < SB_HAND_MAX ? p.n_hand : SB_HAND_MAX;
#pragma unroll 1
  for (int ci = 0; ci < nh; ci++) {
    const DCard& c = WCARD(wg, p.hand[ci].card);
    if (p.hand[ci].cost > p.mana) continue;
    if (c.kind != KIND_SPELL) {
      const int a0 = 16 * ci;  // 16-bit field at bit a0 (never straddles two words)
      m[a0 >> 5] |= empty16 << (a0 & 31);
      n_play += n_empty;
    } else if (!(c.flags & DCF_TARGET)) {
      w_mask_set(m, 64 + 21 * ci); n_play++;
    } else {
      TL t = w_targets(wg, wg->current_order, w_card_target(c), PT_NONE);
      u32 tm = t.m;  // base points are not encodable
#pragma unroll 1
      while (tm) {
        const int tile = w_ffs(tm) - 1;
        tm &= tm - 1;
        w_mask_set(m, 65 + 21 * ci + (4 - (tile >> 2)) * 4 + (tile & 3)); n_play++;
      }
    }
  }
  int n = n_play;
  if (p.replacable) for (int ci = 0; ci < nh; ci++) { w_mask_set(m, 148 + ci); n++; }
  if (n_play == 0) { w_mask_set(m, 155); n++; }
  return n;
}
SBW_FI bool w_have_winner(const WG* wg) { return wg->pl[0].base < 0 || wg->pl[1].base < 0; }
SBW_FI int w_nth_action(const u32* m, int pick) {  // pick-th set bit of the 156-bit mask
#pragma unroll 1
  for (int w = 0; w < SB_MASK_WORDS; w++) {
    const int c = w_popc(m[w]);
    if (pick < c) { u32 v = m[w]; for (int q = 0; q < pick; q++) v &= v - 1; return w * 32 + w_ffs(v) - 1; }
    pick -= c;
  }
  return SB_ACTION_PASS;
}
SBW_FI int w_pick_action(WG* wg) {  // uniform-random legal agent
  const int nl = w_legal_mask(wg);
  return w_nth_action(wg->lm, (int)w_agent_pick(wg->seed_lo, wg->seed_hi, wg->steps, (u32)nl));
}

SBW_NI void w_game_step(WG* wg, int action) {
  W_SHARED(wg);
  const int lo = wg->local_order;
  WPly& p = wg->pl[lo];
  if (action < 64) {
    const int ci = action >> 4, idx = action & 15;
    if (ci >= p.n_hand) WERR(wg, SB_ERR_INDEX);
    else { p.mana = (i16)(p.mana - p.hand[ci].cost); w_player_play(wg, lo, ci, (4 - (idx >> 2)) * 4 + (idx & 3)); }
  } else if (action < 148) {
    const int ci = (action - 64) / 21, idx = (action - 64) % 21;
    if (idx < 20) {  // index 20 matches no tile: complete no-op (Q5)
      if (ci >= p.n_hand) WERR(wg, SB_ERR_INDEX);
      else {
        const DCard& c = WCARD(wg, p.hand[ci].card);
        if (c.kind != KIND_SPELL) WERR(wg, SB_ERR_INDEX);
        else {
          p.mana = (i16)(p.mana - p.hand[ci].cost);
          w_player_play(wg, lo, ci, (c.flags & DCF_TARGET) ? (4 - (idx >> 2)) * 4 + (idx & 3) : PT_NONE);
        }
      }
    }
  } else if (action < 152) {
    const int ci = action - 148;
    if (ci >= p.n_hand) WERR(wg, SB_ERR_INDEX);
    else { w_player_cycle(wg, lo, ci); p.replacable = 0; }
  } else if (action < 155) {
    const int ci = action - 151;
    if (ci >= p.n_hand) WERR(wg, SB_ERR_INDEX);
    else { const WCard t = p.hand[ci]; const WCard h0 = p.hand[0]; p.hand[ci] = h0; p.hand[0] = t; p.leftmost = 0; }
  }
  const bool done = w_have_winner(wg);  // legal_actions() is never empty (PASS is added when nothing is playable)
  const bool reward = wg->pl[1 - wg->local_order].base <= 0;
  wg->done = (u8)((done ? SB_DONE : 0) | (reward ? SB_REWARD : 0));
  if (action == 155) {
    wg->turn++; wg->draw = 0;  // stream key (turn, draw)
    wg->player_sign = (i8)-wg->player_sign;
    w_board_flip(wg);
    w_to_next_turn(wg);
  }
  wg->steps++;
}

// Between steps only on-board entities matter: rebuild the pool in tile order (what pack + unpack would do).
SBW_NI void w_compact(WG* wg) {
  W_SHARED(wg);
  const u32 occ = wg->occ;
  FOR_LANES(l) { wg->remap[l] = 0xFF; if (l + 32 < MAXE) wg->remap[l + 32] = 0xFF; } END_LANES
  FOR_LANES(l) { if (l < 20) { const int s = wg->board[l]; if (s >= 0) wg->remap[s] = (u8)w_popc(occ & ((1u << l) - 1u)); } } END_LANES
  // frozen strength of board-instance card records whose object left the board (n_obj is an upper bound: 0 = none)
#pragma unroll 1
  for (int o = 0; o < 2 && wg->n_obj; o++) {
    WPly& p = wg->pl[o];
#pragma unroll 1
    for (int i = 0; i < p.n_hand; i++) if (p.hand[i].link >= 0) {
      const int lk = p.hand[i].link;
      const u8 r = wg->remap[lk];
      if (r == 0xFF) { p.hand[i].xstr = wg->e_str[lk]; p.hand[i].link = -1; } else p.hand[i].link = (i8)r;
    }
#pragma unroll 1
    for (int i = 0; i < p.n_deck; i++) if (p.deck[i].link >= 0) {
      const int lk = p.deck[i].link;
      const u8 r = wg->remap[lk];
      if (r == 0xFF) { p.deck[i].xstr = wg->e_str[lk]; p.deck[i].link = -1; } else p.deck[i].link = (i8)r;
    }
  }
  int w = 0;
  if (wg->n_mem) {  // a memory survives iff the temple at the root of its tree is still on the board
    i8* keep = wg->scr;
    i8* nidx = wg->scr + 12;
    const int nm = wg->n_mem;
#pragma unroll 1
    for (int i = 0; i < nm; i++) {
      const WMem& m = wg->mem[i];
      const bool k = m.parent < 0 ? (m.b005 >= 0 && wg->remap[m.b005] != 0xFF) : (keep[m.parent] != 0);
      keep[i] = k; nidx[i] = k ? (i8)w++ : (i8)-1;
    }
#pragma unroll 1
    for (int i = 0; i < nm; i++) if (keep[i]) {
      WMem m = wg->mem[i];
      if (m.parent >= 0) m.parent = nidx[m.parent]; else m.b005 = (i8)wg->remap[m.b005];
      wg->mem[nidx[i]] = m;
    }
  }
  wg->n_mem = (u8)w;
  // what would not fit the packed layout is an overflow there too (keeps rollouts == step-per-launch)
  int nobj = 0;
#pragma unroll 1
  for (int o = 0; o < 2 && wg->n_obj; o++) {
#pragma unroll 1
    for (int i = 0; i < wg->pl[o].n_hand; i++) nobj += (wg->pl[o].hand[i].flags & SB_CF_OBJ) != 0;
#pragma unroll 1
    for (int i = 0; i < wg->pl[o].n_deck; i++) nobj += (wg->pl[o].deck[i].flags & SB_CF_OBJ) != 0;
  }
  if (w > NMEM_PACKED || nobj > NOBJ_PACKED) WERR(wg, SB_ERR_OVERFLOW);
  wg->n_obj = (u8)(nobj > 255 ? 255 : nobj);
  // lane t carries the entity of tile t to slot rank(t): read everything, then write
  LV<int> id, str, fl, card;
  LV<u32> st;
  FOR_LANES(l) {
    const int s = (l < 20) ? wg->board[l] : -1;
    LVAL(id, l) = s;
    LVAL(str, l) = s >= 0 ? wg->e_str[s] : 0; LVAL(fl, l) = s >= 0 ? wg->e_fl[s] : 0;
    LVAL(card, l) = s >= 0 ? wg->e_card[s] : 0; LVAL(st, l) = s >= 0 ? wg->e_st[s] : 0u;
  } END_LANES
  FOR_LANES(l) {
    if (LVAL(id, l) >= 0) {
      const int r = w_popc(occ & ((1u << l) - 1u));
      wg->e_str[r] = (i16)LVAL(str, l); wg->e_dmg[r] = 0;
      wg->e_fl[r] = (u8)(LVAL(fl, l) & ~(WEF_RPLAY | WEF_SINGLE));
      wg->e_card[r] = (u8)LVAL(card, l); wg->e_st[r] = LVAL(st, l);
      wg->e_pos[r] = (u8)l; wg->e_mid[r] = 0; wg->e_plen[r] = 0;
      wg->board[l] = (i8)r;
    }
  } END_LANES
  wg->n_ent = (u8)w_popc(occ);
  wg->n_trig = 0; wg->resolving = 0; wg->depth = 0;
}
// Fast path between two steps of an in-kernel rollout: nothing references an off-board entity any more (the trigger stack
// is empty, no B005 memory, no B305 board-instance records), so the garbage can stay until the pool no longer guarantees
// the 28 free slots a fresh unpack would give (same policy as end_of_step in sb_engine.cuh).
SBW_FI void w_end_of_step(WG* wg) {
  if (wg->pl[0].n_hand > SB_HAND_MAX || wg->pl[1].n_hand > SB_HAND_MAX || wg->pl[0].n_deck > SB_DECK_MAX || wg->pl[1].n_deck > SB_DECK_MAX)
    WERR(wg, SB_ERR_OVERFLOW);
  if (wg->n_ent > SB_N_TILES || wg->n_mem || wg->n_obj) w_compact(wg);
  else { wg->n_trig = 0; wg->resolving = 0; wg->depth = 0; }
}
