// sb_engine.cuh -- device-side Stormbound rules engine for sm_100a (one game per thread).
//
// Working set: the 512-byte packed state (include/sb_state.h) is expanded into a compact thread-private
// struct (byte-wide entity pool + tile->entity map + byte-wide hands/decks); card statics come from a
// 24-byte-per-card table staged in SHARED memory by every CTA.  Everything the reference does by
// Python recursion (ability -> damage -> death trigger -> ability -> command -> move ...) is done with
// real device recursion on the per-thread stack (cudaLimitStackSize is raised by sb_create).
//
// Reference lines are cited per function (paths relative to the reference checkout).  The semantic
// twin on the CPU is oracle/sb_oracle*.c -- a separate, object-per-entity restatement used only by
// the tests; the two share nothing but the generated card DATA table and include/sb_state.h.
#pragma once
#include <stdint.h>
#include "../../include/sb_state.h"
#include "sb_card_ids.h"

#include "sb_defs.h"

#define SBD __device__
#define SBD_NI __device__ __noinline__
#define SBD_FI __device__ __forceinline__

struct Target { u8 kind, side, status, xstatus, has_limit, nonhero, base, pad; u16 types, xtypes; i16 limit; };

// entity flags
#define EF_OWNER 1
#define EF_STRUCT 2
#define EF_FIXED 4
#define EF_SINGLE 8
#define EF_RPLAY 16
struct __align__(8) Ent {  // unit.py:8-23 / structure.py:8-16 (statics live in DCard); 8-byte aligned: copies are 3 x 64-bit moves
  u8 card, fl;
  i16 strength, dmg;
  u8 st[5];
  u8 move_id, x, y, path_len;
  u8 path[MAXPATH];  // (y+1)*4 + x, y in -1..5
};
struct __align__(8) CardRec { u8 card; i8 cost; u8 flags; i8 link; u16 wn; i16 xstr; };  // one 64-bit move per record
struct Ply {  // player.py:13-37
  i16 base, max_mana, mana;
  i8 front_line;
  u8 replacable, leftmost, n_hand, n_deck, faction;
  CardRec hand[HAND_W];  // working capacity > packed capacity: a cycle holds 17 deck cards for a moment
  CardRec deck[DECK_W];
};
// cards/b005.py remembered deep copies.  parent < 0: a memory of the live temple entity `b005`; parent >= 0: a
// memory held BY the remembered temple copy mem[parent] (its own ability_remembered); parent index < own index.
struct __align__(8) Mem { i8 b005; u8 pos, card, fl; i16 strength; u8 st[5]; i8 parent; };  // 16 bytes

struct G {
  Ent e[MAXE];
  i8 board[SB_N_TILES];
  Ply pl[2];
  u32 seed_lo, seed_hi;
  u16 turn, draw, steps;
  u8 local_order, current_order, phase, err, done;
  i8 player_sign;
  u8 hist_n, hist_card[4], hist_owner[4];
  u8 n_ent, n_trig, resolving, depth, n_mem;
  u8 n_obj;   // card records that are board instances of B305 (SB_CF_OBJ); 0 on the fast path
  u8 maybe_badobs;  // 0 = no card of this game has an unencodable observation id (UP01-03, Q12): features() skips those scans
  u32 occ;    // occupied-tile bitmask, mirrors board[]
  u32 own1;   // occupied tiles whose entity belongs to order 1 (bits of empty tiles are don't-care)
  u32 strc;   // occupied tiles holding a structure (bits of empty tiles are don't-care)
  __align__(8) u8 trig[MAXTRIG];  // entity id | has_source << 7   (8-byte aligned: end of the block copy_g moves)
  Mem mem[NMEM];
  const DCard* cards;   // shared-memory copy
  const double* wt;     // f^n(1) table in global memory
};

#define GERR(g, code) do { if (!(g).err) (g).err = (code); } while (0)
// Every G is a thread-private local of its kernel.  Telling the compiler so at the top of the out-of-line functions
// turns the generic 64-bit loads/stores through `G&` into local-space ones (no per-access descriptor moves).
#ifdef __CUDA_ARCH__
#define G_LOCAL(g) __builtin_assume(__isLocal(&(g)))
#define P_LOCAL(p) __builtin_assume(__isLocal(p))  // same for Target objects and result lists (always caller locals)
#define P_SHARED(p) __builtin_assume(__isShared(p))  // the card table every CTA stages in shared memory
#else
#define G_LOCAL(g) ((void)0)
#define P_LOCAL(p) ((void)0)
#define P_SHARED(p) ((void)0)
#endif

SBD_FI int PTX(int pt) { return pt >= 20 ? -1 : (pt & 3); }
SBD_FI int PTY(int pt) { return pt == PT_BASE_REMOTE ? -1 : pt == PT_BASE_LOCAL ? 5 : (pt >> 2); }
SBD_FI int PT(int x, int y) { return y * 4 + x; }
SBD_FI bool valid_xy(int x, int y) { return (unsigned)x <= 3u && (unsigned)y <= 4u; }
SBD_FI bool is_base_pt(int pt) { return pt >= 20; }
SBD_FI const DCard& CARD(const G& g, int card) { const DCard* c = g.cards; P_SHARED(c); return c[card]; }
SBD_FI int ent_owner(const Ent& e) { return e.fl & EF_OWNER; }
SBD_FI bool ent_struct(const Ent& e) { return e.fl & EF_STRUCT; }

// ---------------------------------------------------------------- Philox4x32-10 counter stream
SBD_FI void philox(u32 c0, u32 c1, u32 c2, u32 c3, u32 k0, u32 k1, u32& o0, u32& o1) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    u32 h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
    u32 h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
    u32 n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
    c0 = n0; c1 = l1; c2 = n2; c3 = l0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  o0 = c0; o1 = c1;
}
// one out-of-line IEEE FP64 division (each inline expansion is ~30 instructions of Newton iteration; the
// rollout kernels are instruction-fetch bound, so code size matters more than a call)
SBD_NI double ddiv(double a, double b) { return __ddiv_rn(a, b); }
SBD_NI int rng_below(G& g, int n) {
  G_LOCAL(g);
  if (n <= 0) { GERR(g, SB_ERR_EMPTY_CHOICE); return 0; }
  u32 w0, w1;
  philox(g.draw, g.turn, 0, 0, g.seed_lo, g.seed_hi, w0, w1);
  g.draw++;
  return (int)__umulhi(w0, (u32)n);
}
SBD_NI double rng_random(G& g) {
  G_LOCAL(g);
  u32 w0, w1;
  philox(g.draw, g.turn, 0, 0, g.seed_lo, g.seed_hi, w0, w1);
  g.draw++;
  // (a * 2^26 + b) / 2^53: a power-of-two divisor, so the multiplication by 2^-53 is bit-identical
  return __dmul_rn(__dadd_rn(__dmul_rn((double)(w0 >> 5), 67108864.0), (double)(w1 >> 6)), 0x1.0p-53);
}
SBD void shuffle(G& g, i8* a, int n) {
  #pragma unroll 1
  for (int i = n - 1; i > 0; i--) {
    int j = rng_below(g, i + 1);
    i8 t = a[i]; a[i] = a[j]; a[j] = t;
  }
}
SBD_FI u32 agent_pick(u32 seed_lo, u32 seed_hi, u32 step, u32 n) {
  u32 w0, w1;
  philox(step, 0, 0xA6E7u, 0, seed_lo, seed_hi, w0, w1);
  return __umulhi(w0, n);
}

// ---------------------------------------------------------------- board access (board.py:58-92)
SBD_FI int opponent_of(const G& g, int order) { return order == 0 ? 1 - g.local_order : g.local_order; }  // player.py:42-44 (Q3)
SBD_FI int at_xy(const G& g, int x, int y) { return valid_xy(x, y) ? g.board[y * 4 + x] : -1; }
SBD_FI int at_pt(const G& g, int pt) { return (unsigned)pt < 20u ? g.board[pt] : -1; }
// g.occ = bitmask of occupied tiles, kept in step with board[] by the three writers below (+ unpack);
// every board scan walks the set bits only (a board holds ~6 entities, not 20).
SBD_FI void set_xy(G& g, int x, int y, int id) {
  const int t = y * 4 + x;
  const u32 b = 1u << t;
  g.board[t] = (i8)id;
  if (id >= 0) {
    Ent& e = g.e[id];
    e.x = (u8)x; e.y = (u8)y;
    g.occ |= b;
    g.own1 = (e.fl & EF_OWNER) ? (g.own1 | b) : (g.own1 & ~b);
    g.strc = (e.fl & EF_STRUCT) ? (g.strc | b) : (g.strc & ~b);
  } else {
    g.occ &= ~b;
  }
}
SBD_FI void clear_at(G& g, const Ent& e) { const int t = e.y * 4 + e.x; g.board[t] = -1; g.occ &= ~(1u << t); }
// next tile of mask m in scan order: ascending (pov == local: y=0..4, x=0..3) or descending (board.py:157-158)
SBD_FI int next_tile(u32& m, bool ascending) {
  int t = ascending ? __ffs(m) - 1 : 31 - __clz(m);
  m &= ~(1u << t);
  return t;
}
SBD_FI void calc_front_line(G& g, int order) {  // board.py:78-92
  const bool local = order == g.local_order;
  int fl = local ? 4 : 0;
  const u32 m = g.occ & (order ? g.own1 : ~g.own1);  // this side's entities: first (local) / last (remote) occupied row
  if (m) {
    const int y = (local ? __ffs(m) - 1 : 31 - __clz(m)) >> 2;
    fl = local ? (y > 1 ? y : 1) : (y < 3 ? y : 3);
  }
  g.pl[order].front_line = (i8)fl;
}

// ---------------------------------------------------------------- target queries (board.py:147-296)
SBD_FI Target mkT(int kind, int side) {
  Target t; t.kind = (u8)kind; t.side = (u8)side; t.status = 0; t.xstatus = 0; t.has_limit = 0; t.nonhero = 0; t.base = 0;
  t.pad = 0; t.types = 0; t.xtypes = 0; t.limit = 0; return t;
}
SBD_FI Target card_target(const DCard& c) {
  Target t = mkT(c.t_ks & 3, c.t_ks >> 2);
  t.types = c.t_types; t.xtypes = c.t_xtypes; t.status = c.t_status; t.xstatus = c.t_xstatus;
  t.has_limit = c.t_limit >= 0; t.limit = c.t_limit; t.nonhero = (c.flags & DCF_TNONHERO) != 0; t.base = (c.flags & DCF_TBASE) != 0;
  return t;
}
// region = bitmask over tiles (bit t) that a tile must belong to; 0xFFFFF = whole board.
// base_passes: base points survive the region filter when include_base (board.py:215,262,276,294).
SBD_NI int get_targets_region(const G& g, int pov, const Target& t, int exclude_pt, u32 region, bool base_passes, i8* out) {
  G_LOCAL(g);
  P_LOCAL(&t);
  P_LOCAL(out);
  int n = 0;
  const bool pov_local = (pov == g.local_order);
  // side and kind are decided by bitmask algebra (own1 = tiles whose entity belongs to order 1, strc =
  // structure tiles); the per-entity loop only sees real candidates and usually checks strength > 0 alone
  u32 m = g.occ & region;
  if ((unsigned)exclude_pt < 20u) m &= ~(1u << exclude_pt);
  if (t.side != TS_ANY) {
    const u32 mine = pov ? g.own1 : ~g.own1;
    m &= (t.side == TS_FRIENDLY) ? mine : ~mine;
  }
  if (t.kind == TK_UNIT) m &= ~g.strc; else if (t.kind == TK_STRUCTURE) m &= g.strc;
  // the rare filters of board.py:170-179, read once (the stores to `out` would otherwise force a reload per candidate)
  const bool has_limit = t.has_limit != 0;
  const int limit = t.limit;
  const u32 want_types = t.types, bad_types = (u32)t.xtypes | (t.nonhero ? (1u << UT_HERO) : 0u);
  const u32 want_status = t.status, bad_status = t.xstatus;
  const bool filters = has_limit | (want_types != 0) | (bad_types != 0) | (want_status != 0) | (bad_status != 0);
  const DCard* cards = g.cards;
  P_SHARED(cards);
  #pragma unroll 1
  while (m) {
    const int tile = next_tile(m, pov_local);
    const Ent& e = g.e[g.board[tile]];
    const int str = e.strength;
    if (str <= 0) continue;  // board.py:164
    if (filters) {
      if (has_limit && str > limit) continue;
      if (!ent_struct(e)) {  // structures only honour strength_limit (board.py:179)
        if (want_types | bad_types) {
          const u32 types = cards[e.card].types;
          if ((want_types && !(types & want_types)) || (types & bad_types)) continue;
        }
        if (want_status | bad_status) {
          u32 have = 0;
#pragma unroll
          for (int k = 0; k < 5; k++) have |= (e.st[k] ? 1u : 0u) << k;
          if ((want_status && !(have & want_status)) || (have & bad_status)) continue;
        }
      }
    }
    out[n++] = (i8)tile;
  }
  if (t.base && base_passes) {
    int friendly = pov_local ? PT_BASE_LOCAL : PT_BASE_REMOTE;
    int enemy = pov_local ? PT_BASE_REMOTE : PT_BASE_LOCAL;
    if (t.side != TS_ENEMY && friendly != exclude_pt) out[n++] = (i8)friendly;
    if (t.side != TS_FRIENDLY && enemy != exclude_pt) out[n++] = (i8)enemy;
  }
  return n;
}
SBD_FI int get_targets(const G& g, int pov, const Target& t, int exclude_pt, i8* out) {
  return get_targets_region(g, pov, t, exclude_pt, 0xFFFFFu, true, out);
}
// Neighbourhoods as arithmetic (no lookup tables): masks for the filtered queries, ordered lists for the
// unfiltered ones.  Orders follow board.py:236-296: side [L,R]; bordering [L,R,y-1,y+1];
// surrounding [L, L y-1, L y+1, R, R y-1, R y+1, y-1, y+1].
SBD_FI u32 border_mask(int x, int y) {
  u32 m = 0;
  if (x > 0) m |= 1u << (y * 4 + x - 1);
  if (x < 3) m |= 1u << (y * 4 + x + 1);
  if (y > 0) m |= 1u << (y * 4 + x - 4);
  if (y < 4) m |= 1u << (y * 4 + x + 4);
  return m;
}
SBD_FI u32 surround_mask(int x, int y) {
  u32 row = (x > 0 ? 1u << (x - 1) : 0u) | (1u << x) | (x < 3 ? 1u << (x + 1) : 0u);  // columns x-1..x+1
  u32 m = (row & ~(1u << x)) << (y * 4);
  if (y > 0) m |= row << (y * 4 - 4);
  if (y < 4) m |= row << (y * 4 + 4);
  return m;
}
SBD_FI int side_list(int x, int y, i8* out) {
  int n = 0;
  if (x > 0) out[n++] = (i8)(y * 4 + x - 1);
  if (x < 3) out[n++] = (i8)(y * 4 + x + 1);
  return n;
}
SBD_FI int border_list(int x, int y, i8* out) {
  int n = side_list(x, y, out);
  if (y > 0) out[n++] = (i8)(y * 4 + x - 4);
  if (y < 4) out[n++] = (i8)(y * 4 + x + 4);
  return n;
}
SBD_FI int surround_list(int x, int y, i8* out) {
  int n = 0;
  #pragma unroll 1
  for (int dx = -1; dx <= 1; dx += 2) {
    int xx = x + dx;
    if (xx < 0 || xx > 3) continue;
    out[n++] = (i8)(y * 4 + xx);
    if (y > 0) out[n++] = (i8)(y * 4 + xx - 4);
    if (y < 4) out[n++] = (i8)(y * 4 + xx + 4);
  }
  if (y > 0) out[n++] = (i8)(y * 4 + x - 4);
  if (y < 4) out[n++] = (i8)(y * 4 + x + 4);
  return n;
}

SBD_FI void sort_pts_by_y(i8* a, int n, bool desc) {  // stable insertion sort on Point.y (board.py:217,232)
  #pragma unroll 1
  for (int i = 1; i < n; i++) {
    i8 v = a[i];
    int j = i - 1;
    #pragma unroll 1
    while (j >= 0 && (desc ? PTY(a[j]) < PTY(v) : PTY(a[j]) > PTY(v))) { a[j + 1] = a[j]; j--; }
    a[j + 1] = v;
  }
}
// board.py:206-234.  toward_enemy: front tiles; else behind tiles.  t == nullptr: the plain tile list.
SBD_NI int column_tiles(const G& g, int x, int y, int pov, const Target* t, bool front, i8* out) {
  G_LOCAL(g);
  P_LOCAL(out);
  bool pov_local = (pov == g.local_order);
  bool up = (pov_local == front);  // decreasing y
  int n = 0;
  if (t) {
    u32 region = 0;
    if (up) for (int i = y - 1; i >= 0; i--) region |= 1u << PT(x, i);
    else for (int i = y + 1; i < 5; i++) region |= 1u << PT(x, i);
    n = get_targets_region(g, pov, *t, PT_NONE, region, true, out);
    sort_pts_by_y(out, n, up);
  } else {
    if (up) for (int i = y - 1; i >= 0; i--) out[n++] = (i8)PT(x, i);
    else for (int i = y + 1; i < 5; i++) out[n++] = (i8)PT(x, i);
  }
  return n;
}
SBD_FI int bordering(const G& g, int x, int y, int pov, const Target* t, i8* out) {  // board.py:266-278
  if (t) return get_targets_region(g, pov, *t, PT_NONE, border_mask(x, y), true, out);
  return border_list(x, y, out);
}
SBD_FI int surrounding(const G& g, int x, int y, int pov, const Target* t, i8* out) {  // board.py:280-296
  if (t) return get_targets_region(g, pov, *t, PT_NONE, surround_mask(x, y), true, out);
  return surround_list(x, y, out);
}
SBD_FI bool within_front_line(const G& g, int order, int y) {  // player.py:96-100 (Q22)
  return order == 0 ? y >= g.pl[order].front_line : y <= g.pl[order].front_line;
}
SBD_FI int within_front_line_tiles(const G& g, int order, i8* out) {  // player.py:102-111
  int n = 0, fl = g.pl[order].front_line;
  if (order == 0) { for (int y = fl; y < 5; y++) for (int x = 0; x < 4; x++) out[n++] = (i8)PT(x, y); }
  else { for (int y = fl; y >= 0; y--) for (int x = 3; x >= 0; x--) out[n++] = (i8)PT(x, y); }
  return n;
}

// ---------------------------------------------------------------- entities
SBD_NI int new_ent(G& g, int card, int owner, int strength) {
  G_LOCAL(g);
  if (g.n_ent >= MAXE) { GERR(g, SB_ERR_OVERFLOW); return MAXE - 1; }
  int id = g.n_ent++;
  Ent& e = g.e[id];
  const DCard& c = CARD(g, card);
  e.card = (u8)card;
  e.fl = (u8)((owner ? EF_OWNER : 0) | (c.kind == KIND_STRUCTURE ? EF_STRUCT : 0) | ((c.flags & DCF_FIXED) ? EF_FIXED : 0));
  e.strength = (i16)strength; e.dmg = 0;
#pragma unroll
  for (int k = 0; k < 5; k++) e.st[k] = 0;
  e.move_id = 0; e.x = 0; e.y = 0; e.path_len = 0;
  return id;
}
SBD_NI int spawn_token_unit(G& g, int owner, int pt, int strength, int type) {  // board.py:298-311
  G_LOCAL(g);
  int id = new_ent(g, SBC_TOKEN_UNIT0 + type, owner, strength);
  set_xy(g, PTX(pt), PTY(pt), id);
  calc_front_line(g, owner);
  return id;
}

// forward declarations of the mutually recursive core
SBD_NI void ability(G& g, int id, int pos_pt, int has_source);
SBD_NI void effect(G& g, int id, int pos_pt, int has_source);
SBD_NI void spell_effect(G& g, int card, int caster, int pos_pt);
SBD_NI void unit_move(G& g, int id);
SBD_NI void player_play(G& g, int order, int index, int pos_pt);

// ---------------------------------------------------------------- trigger stack (board.py:46-56, card.py:48-62)
SBD_FI void push_trigger(G& g, int id, int has_source) {
  if (g.n_trig >= MAXTRIG) { GERR(g, SB_ERR_OVERFLOW); return; }
  g.trig[g.n_trig++] = (u8)(id | (has_source ? 0x80 : 0));
}
SBD_FI void pop_trigger(G& g) {
  if (g.n_trig == 0 || g.resolving) return;
  u8 t = g.trig[--g.n_trig];
  ability(g, t & 0x7F, PT_NONE, t >> 7);
}
SBD_NI void ability(G& g, int id, int pos_pt, int has_source) {
  G_LOCAL(g);
  if (!(CARD(g, g.e[id].card).flags & DCF_ABILITY)) return;  // un-overridden Card.activate_ability: no wrapper
  if (g.depth > MAXDEPTH) { GERR(g, SB_ERR_DEPTH); return; }
  g.depth++;
  g.resolving = 1;
  effect(g, id, pos_pt, has_source);
  g.resolving = 0;
  pop_trigger(g);
  g.depth--;
}
SBD_FI void spell_ability(G& g, int card, int caster, int pos_pt) {
  if (g.depth > MAXDEPTH) { GERR(g, SB_ERR_DEPTH); return; }
  g.depth++;
  g.resolving = 1;
  spell_effect(g, card, caster, pos_pt);
  g.resolving = 0;
  pop_trigger(g);
  g.depth--;
}

// ---------------------------------------------------------------- status verbs (unit.py:239-275)
SBD_FI void st_add(G& g, int id, int s) {  // 6-bit packed counters: the 64th copy of one status flags the game
  if (g.e[id].st[s] < 63) g.e[id].st[s]++; else GERR(g, SB_ERR_OVERFLOW);
}
SBD_FI void st_remove(G& g, int id, int s) { if (g.e[id].st[s]) g.e[id].st[s]--; else GERR(g, SB_ERR_INDEX); }
SBD_FI void v_freeze(G& g, int id) { st_add(g, id, SB_ST_FROZEN); }
SBD_FI void v_poison(G& g, int id) { if (g.e[id].st[SB_ST_VITALIZED]) st_remove(g, id, SB_ST_VITALIZED); st_add(g, id, SB_ST_POISONED); }
SBD_FI void v_vitalize(G& g, int id) { if (g.e[id].st[SB_ST_POISONED]) st_remove(g, id, SB_ST_POISONED); st_add(g, id, SB_ST_VITALIZED); }
SBD_FI void v_confuse(G& g, int id) { st_add(g, id, SB_ST_CONFUSED); }
SBD_FI void v_disable(G& g, int id) { if (CARD(g, g.e[id].card).flags & DCF_ABILITY) st_add(g, id, SB_ST_DISABLED); }
SBD_FI void v_heal(G& g, int id, int amount) { g.e[id].strength = (i16)(g.e[id].strength + amount); }

// ---------------------------------------------------------------- damage (unit.py:205-231, structure.py:52-69, player.py:83-88)
SBD_FI int player_damage(G& g, int order, int amount) { g.pl[order].base = (i16)(g.pl[order].base - amount); return amount; }
SBD_NI void destroy(G& g, int id, int has_source) {
  G_LOCAL(g);
  Ent& e = g.e[id];
  clear_at(g, e);  // by (possibly stale) position, like board.set(self.position, None) (Q21)
  e.dmg = e.strength;
  if (!ent_struct(e)) {
    e.path_len = 0;
    if (CARD(g, e.card).trigger == TR_ON_DEATH) { push_trigger(g, id, has_source); pop_trigger(g); }
  }
  calc_front_line(g, opponent_of(g, g.current_order));
}
SBD_NI int deal_damage(G& g, int id, int amount, int pending, int has_source) {
  G_LOCAL(g);
  Ent& e = g.e[id];
  if (e.strength - amount < 0) amount = e.strength;
  e.dmg = (i16)amount;
  e.strength = (i16)(e.strength - amount);
  if (!pending && e.strength <= 0) destroy(g, id, has_source);
  else if (!ent_struct(e) && e.strength > 0 && CARD(g, e.card).trigger == TR_AFTER_SURVIVING) { push_trigger(g, id, has_source); pop_trigger(g); }
  return amount;
}
SBD_FI int deal_damage_pt(G& g, int pt, int amount, int has_source) {  // board.at(point).deal_damage(...)
  if (pt == PT_BASE_LOCAL) return player_damage(g, g.local_order, amount);
  if (pt == PT_BASE_REMOTE) return player_damage(g, 1 - g.local_order, amount);
  int id = at_pt(g, pt);
  if (id < 0) { GERR(g, SB_ERR_NONE_TARGET); return 0; }
  return deal_damage(g, id, amount, 0, has_source);
}

// ---------------------------------------------------------------- movement (unit.py:66-203, 277-382)
SBD_FI u8 enc_xy(int x, int y) { return (u8)((y + 1) * 4 + x); }
SBD_FI void set_path_inl(G& g, int id, int on_play, int extra_movement) {  // unit.py:78-122; inlined where a move follows at once
  Ent& e = g.e[id];
  u8 dest[MAXPATH];
  int nd = 0;
  int px = e.x, py = e.y;
  int confused_cached = e.st[SB_ST_CONFUSED];
  int owner = ent_owner(e);
  bool is_local = owner == g.local_order;
  int steps = on_play ? CARD(g, e.card).movement + extra_movement : 1;
  if (steps > MAXPATH) { GERR(g, SB_ERR_OVERFLOW); steps = MAXPATH; }
  #pragma unroll 1
  for (int i = 0; i < steps; i++) {
    int dx = px, dy = py + (is_local ? -1 : 1);
    if (confused_cached > 0) {
      int delta;
      if (px == 0) { rng_below(g, 1); delta = 1; }
      else if (px == 3) { rng_below(g, 1); delta = -1; }
      else delta = rng_below(g, 2) == 0 ? -1 : 1;
      dx = px + delta; dy = py;
      confused_cached--;
    } else if (on_play && !(e.fl & EF_FIXED) && dy != (is_local ? -1 : 5)) {
      int nxt = at_xy(g, dx, dy);
      if (nxt < 0 || ent_owner(g.e[nxt]) == owner) {
        int left = px > 0 ? at_xy(g, px - 1, py) : -1;
        int right = px < 3 ? at_xy(g, px + 1, py) : -1;
        u8 lenc = enc_xy(px - 1, py), renc = enc_xy(px + 1, py);
        bool left_ok = left >= 0 && ent_owner(g.e[left]) != owner;
        bool right_ok = right >= 0 && ent_owner(g.e[right]) != owner;
        #pragma unroll 1
        for (int k = 0; k < nd; k++) { if (dest[k] == lenc) left_ok = false; if (dest[k] == renc) right_ok = false; }
        if (px <= 1) { if (right_ok) { dx = px + 1; dy = py; } else if (left_ok) { dx = px - 1; dy = py; } }
        else { if (left_ok) { dx = px - 1; dy = py; } else if (right_ok) { dx = px + 1; dy = py; } }
      }
    }
    dest[nd++] = enc_xy(dx, dy);
    px = dx; py = dy;
  }
  #pragma unroll 1
  for (int i = 0; i < nd; i++) e.path[i] = dest[i];
  e.path_len = (u8)nd;
}

SBD_NI void set_path(G& g, int id, int on_play, int extra_movement) {  // out-of-line copy for the rare callers
  G_LOCAL(g);
  set_path_inl(g, id, on_play, extra_movement);
}
SBD_NI void unit_move(G& g, int id) {  // unit.py:124-203
  G_LOCAL(g);
  Ent& e = g.e[id];
  if (g.depth > MAXDEPTH) { GERR(g, SB_ERR_DEPTH); return; }
  g.depth++;
  const int trig = CARD(g, e.card).trigger;
  const u8 current_id = ++e.move_id;
  if (g.phase == PH_TURN_START) {
    if (e.st[SB_ST_POISONED]) deal_damage(g, id, 1, 0, 0);
    else if (e.st[SB_ST_VITALIZED]) v_heal(g, id, 1);
    if (e.st[SB_ST_FROZEN]) { st_remove(g, id, SB_ST_FROZEN); g.depth--; return; }
  }
  if (e.path_len == 0) { g.depth--; return; }
  if (trig == TR_BEFORE_MOVING && !e.st[SB_ST_DISABLED]) ability(g, id, PT_NONE, 1);
  if (e.st[SB_ST_FROZEN]) { g.depth--; return; }
  u8 path[MAXPATH];
  const int np = e.path_len;  // `for destination in self.path` iterates the list object bound now
  #pragma unroll 1
  for (int i = 0; i < np; i++) path[i] = e.path[i];
  #pragma unroll 1
  for (int i = 0; i < np; i++) {
    const int dx = path[i] & 3, dy = (path[i] >> 2) - 1;
    const int owner = ent_owner(e);
    if (dy < 0 || dy > 4) {  // to base
      if (trig == TR_BEFORE_ATTACKING && !e.st[SB_ST_DISABLED]) ability(g, id, -2, 1);
      int target = dy < 0 ? 1 - g.local_order : g.local_order;
      player_damage(g, target, e.strength);
      if (g.pl[target].base > 0) destroy(g, id, 0);
      g.depth--;
      return;
    }
    int tid = at_xy(g, dx, dy);
    bool attacked = false;
    if (tid >= 0 && ent_owner(g.e[tid]) == owner && dx == e.x) { g.depth--; return; }
    if (tid >= 0 && (e.st[SB_ST_CONFUSED] || ent_owner(g.e[tid]) != owner)) {
      if (trig == TR_BEFORE_ATTACKING && !e.st[SB_ST_DISABLED]) ability(g, id, PT(dx, dy), 1);
      tid = at_xy(g, dx, dy);
      if (tid >= 0) {
        Ent& t = g.e[tid];
        int tstr = t.strength;
        int t_pending = !ent_struct(t) && CARD(g, t.card).trigger == TR_ON_DEATH && !t.st[SB_ST_DISABLED];
        int l_pending = trig == TR_ON_DEATH && !e.st[SB_ST_DISABLED];
        deal_damage(g, tid, e.strength, t_pending, 0);
        deal_damage(g, id, tstr, l_pending, 0);
        if (t.strength <= 0 && t_pending) destroy(g, tid, 0);
        if (e.strength <= 0 && l_pending) destroy(g, id, 0);
        attacked = true;
      }
    }
    if (current_id != e.move_id) { g.depth--; return; }
    if (at_xy(g, dx, dy) < 0 && e.strength > 0) {
      clear_at(g, e);
      set_xy(g, dx, dy, id);
      Ply& p = g.pl[ent_owner(e)];
      if (p.front_line > dy) p.front_line = (i8)(dy > 1 ? dy : 1);
      if (attacked && trig == TR_AFTER_ATTACKING && !e.st[SB_ST_DISABLED]) ability(g, id, PT_NONE, 1);
      if (e.st[SB_ST_CONFUSED]) st_remove(g, id, SB_ST_CONFUSED);
    }
  }
  g.depth--;
}
SBD_FI void unit_play(G& g, int id, int x, int y) {  // unit.py:66-76
  g.e[id].fl |= EF_RPLAY;
  set_xy(g, x, y, id);
  set_path_inl(g, id, 1, 0);
  if (CARD(g, g.e[id].card).trigger == TR_ON_PLAY) ability(g, id, PT_NONE, 1);
  unit_move(g, id);
  g.e[id].fl &= ~EF_RPLAY;
}
SBD_NI void struct_play(G& g, int id, int x, int y) {  // structure.py:45-50
  G_LOCAL(g);
  set_xy(g, x, y, id);
  if (CARD(g, g.e[id].card).trigger == TR_ON_PLAY) ability(g, id, PT_NONE, 1);
}
SBD_FI void gain_speed(G& g, int id, int amount) { set_path(g, id, (g.e[id].fl & EF_RPLAY) != 0, amount); }  // unit.py:277-280
SBD_NI void v_command(G& g, int id) {  // unit.py:282-289
  G_LOCAL(g);
  u8 cache = g.e[id].fl & EF_FIXED;
  g.e[id].fl |= EF_FIXED;
  set_path(g, id, 0, 0);
  unit_move(g, id);
  g.e[id].fl = (u8)((g.e[id].fl & ~EF_FIXED) | cache);
}
SBD_NI void v_convert(G& g, int id) {  // unit.py:291-293
  G_LOCAL(g);
  Ent& e = g.e[id];
  int o = opponent_of(g, ent_owner(e));
  e.fl = (u8)((e.fl & ~EF_OWNER) | (o ? EF_OWNER : 0));
  if (g.board[e.y * 4 + e.x] == id) {  // keep the owner mask in step (a converted ghost is not on the board)
    const u32 b = 1u << (e.y * 4 + e.x);
    g.own1 = o ? (g.own1 | b) : (g.own1 & ~b);
  }
  set_path(g, id, (e.fl & EF_RPLAY) != 0, 0);
}
SBD_NI void v_push(G& g, int id, int fx, int fy) {  // unit.py:318-339
  G_LOCAL(g);
  Ent& e = g.e[id];
  int dx = 0, dy = 0;
  if (fy < e.y) dy = 1; else if (fy > e.y) dy = -1; else if (fx < e.x) dx = 1; else if (fx > e.x) dx = -1;
  if (dx || dy) {
    #pragma unroll 1
    for (;;) {
      int nx = e.x + dx, ny = e.y + dy;
      if (!valid_xy(nx, ny)) break;
      if (g.board[ny * 4 + nx] >= 0) return;
      clear_at(g, e);
      set_xy(g, nx, ny, id);
    }
  }
  Ply& p = g.pl[ent_owner(e)];
  if (p.front_line > e.y) p.front_line = (i8)(e.y > 1 ? e.y : 1);
}
SBD_NI void v_force_attack(G& g, int id, int tx, int ty) {  // unit.py:341-371
  G_LOCAL(g);
  Ent& e = g.e[id];
  if ((tx != e.x && ty != e.y) || at_xy(g, tx, ty) < 0) return;
  u8 dest[MAXPATH];
  int nd = 0;
  bool vertical = (tx == e.x);
  int fixed = vertical ? e.x : e.y, start = vertical ? e.y : e.x, end = vertical ? ty : tx;
  int delta = end > start ? 1 : -1;
  #pragma unroll 1
  for (int i = start + delta; i != end + delta; i += delta) {
    int x = vertical ? fixed : i, y = vertical ? i : fixed;
    if (i != end && at_xy(g, x, y) >= 0) return;
    dest[nd++] = enc_xy(x, y);
  }
  if (nd > 0) {
    #pragma unroll 1
    for (int i = 0; i < nd; i++) e.path[i] = dest[i];
    e.path_len = (u8)nd;
    unit_move(g, id);
  }
}
SBD_NI void v_teleport(G& g, int id, int dx, int dy) {  // unit.py:373-382
  G_LOCAL(g);
  Ent& e = g.e[id];
  if (at_xy(g, dx, dy) < 0) {
    clear_at(g, e);
    set_xy(g, dx, dy, id);
    Ply& p = g.pl[ent_owner(e)];
    if (p.front_line > dy) p.front_line = (i8)(dy > 1 ? dy : 1);
    set_path(g, id, (e.fl & EF_RPLAY) != 0, 0);
  }
}

// ---------------------------------------------------------------- hand / deck (player.py:46-81)
// list.remove(target): index of the first element that `is` target or == target.  Unit/Structure __eq__:
// card_id, player, position (unit.py:25-26, structure.py:18-19); Spell: uuid (card.py:22-23).  A board
// instance of B305 (SB_CF_OBJ) has a Point position, a pristine card None: comparing the two evaluates
// Point.__eq__(None) -> AttributeError (point.py:7).
SBD_FI int first_equal(G& g, const CardRec* l, int n, int idx) {
  const CardRec t = l[idx];
  if (CARD(g, t.card).kind == KIND_SPELL) return idx;
  #pragma unroll 1
  for (int i = 0; i < idx && i < n; i++) {
    if (l[i].card != t.card) continue;
    int oi = l[i].flags & SB_CF_OBJ, ot = t.flags & SB_CF_OBJ;
    if (oi != ot) { GERR(g, SB_ERR_NONE_TARGET); return idx; }
    if (!oi) return i;
  }
  return idx;
}
SBD_NI void player_draw(G& g, int order, int amount) {  // player.py:46-52 + numpy choice(p=) semantics
  G_LOCAL(g);
  Ply& p = g.pl[order];
  #pragma unroll 1
  for (int k = 0; k < amount; k++) {
    int n = p.n_deck;
    if (n <= 0) { GERR(g, SB_ERR_EMPTY_CHOICE); return; }
    double wv[DECK_W];
    double sum = 0.0;
    #pragma unroll 1
    for (int i = 0; i < n; i++) { wv[i] = __ldg(&g.wt[p.deck[i].wn]); sum = __dadd_rn(sum, wv[i]); }
    const double u = rng_random(g);
    // numpy: p_i = fl(w_i / sum), cdf = cumsum(p), cdf /= cdf[-1], idx = searchsorted(cdf, u, side='right')
    //      = #{i : fl(cdf_i / last) <= u}.
    // Fast path: one division.  a_i = sum_j w_j * (1/sum) differs from the exactly rounded fl(cdf_i / last) by less
    // than 1e-14 (n <= 20 terms, each a few ulp of a value <= 1; last = 1 +- n ulp), so the comparison with u is
    // already decided whenever |a_i - u| > 1e-13.  Only a draw that lands closer than that to a boundary (about one
    // in 1e12) takes the exact path below, which forms every quotient like numpy does.
    const double inv = ddiv(1.0, sum);
    int idx = 0;
    bool close_call = false;
    {
      double acc = 0.0;
      #pragma unroll 1
      for (int i = 0; i < n; i++) {  // the partial sums only grow (weights >= 1): stop at the first one safely above u
        acc += wv[i] * inv;
        const double d = acc - u;
        if (d > 1e-13) break;
        idx += d < -1e-13;
        close_call |= d >= -1e-13;
      }
    }
#ifdef SB_FORCE_EXACT_DRAW  // test builds: always take the exact path
    close_call = true;
#endif
    if (close_call) {
      double cdf[DECK_W];
      double acc = 0.0;
      #pragma unroll 1
      for (int i = 0; i < n; i++) { acc = __dadd_rn(acc, ddiv(wv[i], sum)); cdf[i] = acc; }
      const double last = cdf[n - 1];
      idx = 0;
      #pragma unroll 1
      for (int i = 0; i < n; i++) idx += ddiv(cdf[i], last) <= u;
    }
    if (idx > n - 1) idx = n - 1;
    CardRec c = p.deck[idx];
    c.wn = 0;
    if (p.n_hand >= HAND_W) { GERR(g, SB_ERR_OVERFLOW); return; }
    p.hand[p.n_hand++] = c;
    int j = first_equal(g, p.deck, n, idx);
    if (j != idx) p.deck[idx].wn = 0;
    #pragma unroll 1
    for (int i = j; i < n - 1; i++) p.deck[i] = p.deck[i + 1];
    p.n_deck--;
  }
}
SBD_FI void player_fill_hand(G& g, int order) { player_draw(g, order, 4 - g.pl[order].n_hand); }
SBD_NI void player_discard(G& g, int order, int index) {  // player.py:57-66
  G_LOCAL(g);
  Ply& p = g.pl[order];
  #pragma unroll 1
  for (int i = 0; i < p.n_deck; i++) { if (p.deck[i].wn >= WT_N - 1) GERR(g, SB_ERR_OVERFLOW); else p.deck[i].wn++; }
  CardRec target = p.hand[index];
  int j = first_equal(g, p.hand, p.n_hand, index);
  #pragma unroll 1
  for (int i = j; i < p.n_hand - 1; i++) p.hand[i] = p.hand[i + 1];
  p.n_hand--;
  if (!(target.flags & SB_CF_SINGLE_USE)) {
    if (p.n_deck >= DECK_W) { GERR(g, SB_ERR_OVERFLOW); return; }
    target.wn = 0;
    p.deck[p.n_deck++] = target;
  }
}
SBD_NI void player_play(G& g, int order, int index, int pos_pt) {  // player.py:68-77
  G_LOCAL(g);
  Ply& p = g.pl[order];
  if (index < 0 || index >= p.n_hand) { GERR(g, SB_ERR_INDEX); return; }
  CardRec target = p.hand[index];
  if (g.hist_n < 4) { g.hist_card[g.hist_n] = target.card; g.hist_owner[g.hist_n] = (u8)order; g.hist_n++; }
  else {
    #pragma unroll 1
    for (int i = 0; i < 3; i++) { g.hist_card[i] = g.hist_card[i + 1]; g.hist_owner[i] = g.hist_owner[i + 1]; }
    g.hist_card[3] = target.card; g.hist_owner[3] = (u8)order;
  }
  player_discard(g, order, index);
  const DCard& c = CARD(g, target.card);
  if (c.kind == KIND_SPELL) {  // spell.py:22-24
    bool ok = true;
    if (c.flags & DCF_TARGET) {
      i8 tg[24];
      Target t = card_target(c);
      int n = get_targets(g, g.current_order, t, PT_NONE, tg);
      ok = false;
      #pragma unroll 1
      for (int i = 0; i < n; i++) if (tg[i] == pos_pt) ok = true;
    }
    if (ok) spell_ability(g, target.card, order, pos_pt);
    return;
  }
  if (pos_pt < 0 || pos_pt >= 20) { GERR(g, SB_ERR_INDEX); return; }
  int strength = c.strength;
  if (target.flags & SB_CF_OBJ) strength = target.link >= 0 ? g.e[target.link].strength : target.xstr;
  int id = new_ent(g, target.card, order, strength);  // target.copy(), player.py:74
  g.e[id].fl = (u8)((g.e[id].fl & ~(EF_FIXED | EF_SINGLE)) | ((target.flags & SB_CF_FIXED) ? EF_FIXED : 0) |
                    ((target.flags & SB_CF_SINGLE_USE) ? EF_SINGLE : 0));
  if (c.kind == KIND_UNIT) unit_play(g, id, PTX(pos_pt), PTY(pos_pt));
  else struct_play(g, id, PTX(pos_pt), PTY(pos_pt));
}
SBD_FI void player_cycle(G& g, int order, int index) { player_discard(g, order, index); player_draw(g, order, 1); }  // player.py:79-81

// ---------------------------------------------------------------- turn pipeline (board.py:94-145)
SBD_NI void board_flip(G& g) {  // board.py:94-115
  G_LOCAL(g);
  g.local_order ^= 1;
  g.pl[0].front_line = (i8)(4 - g.pl[0].front_line);
  g.pl[1].front_line = (i8)(4 - g.pl[1].front_line);
  #pragma unroll 1
  for (int t = 0; t < 10; t++) { i8 a = g.board[t]; g.board[t] = g.board[19 - t]; g.board[19 - t] = a; }
  g.occ = (__brev(g.occ) >> 12);  // tile t -> 19 - t
  g.own1 = (__brev(g.own1) >> 12);
  g.strc = (__brev(g.strc) >> 12);
  u32 m = g.occ;
  #pragma unroll 1
  while (m) { int t = next_tile(m, true); int id = g.board[t]; g.e[id].x = (u8)(t & 3); g.e[id].y = (u8)(t >> 2); }
}
SBD_NI void to_next_turn(G& g) {  // board.py:117-145
  G_LOCAL(g);
  i8 pts[24], ids[24];
  int n;
  g.phase = PH_TURN_END;
  player_fill_hand(g, g.current_order);
  // TURN_END structures: none of the 12 structures has that trigger (structure.py:36-38); list building has no side effect
  calc_front_line(g, g.local_order);
  calc_front_line(g, 1 - g.local_order);
  g.pl[g.current_order].max_mana++;
  g.pl[0].mana = g.pl[0].max_mana;
  g.pl[1].mana = g.pl[1].max_mana;
  g.phase = PH_TURN_START;
  g.current_order = (g.current_order == g.local_order) ? 1 - g.local_order : g.local_order;
  g.pl[g.current_order].replacable = 1;
  g.pl[g.current_order].leftmost = 1;
  Target ts = mkT(TK_STRUCTURE, TS_FRIENDLY);
  n = get_targets(g, g.current_order, ts, PT_NONE, pts);
  #pragma unroll 1
  for (int i = 0; i < n; i++) ids[i] = (i8)at_pt(g, pts[i]);
  #pragma unroll 1
  for (int i = 0; i < n; i++)
    if (CARD(g, g.e[ids[i]].card).trigger == TR_TURN_START) ability(g, ids[i], PT(g.e[ids[i]].x, g.e[ids[i]].y), 1);
  Target tu = mkT(TK_UNIT, TS_FRIENDLY);
  n = get_targets(g, g.current_order, tu, PT_NONE, pts);
  #pragma unroll 1
  for (int i = 0; i < n; i++) ids[i] = (i8)at_pt(g, pts[i]);
  #pragma unroll 1
  for (int i = 0; i < n; i++) { set_path_inl(g, ids[i], 0, 0); unit_move(g, ids[i]); }  // snapshot incl. ghosts (Q21)
  g.phase = PH_PLAY;
}

// ---------------------------------------------------------------- legal actions / step (games/stormbound.py:318-373,528-561)
SBD_FI void mask_set(u32* m, int a) { m[a >> 5] |= 1u << (a & 31); }
SBD_NI int legal_mask(const G& g, u32* m) {
  G_LOCAL(g);
  P_LOCAL(m);
  const Ply& p = g.pl[g.local_order];
  int n_play = 0;
#pragma unroll
  for (int i = 0; i < SB_MASK_WORDS; i++) m[i] = 0;
  // PLACE ordinals over y=4..1, x=0..3 that are empty and within the front line: ordinal = (4-y)*4 + x
  const u32 fr = ~g.occ;
  u32 empty16 = ((fr >> 16) & 0xFu) | (((fr >> 12) & 0xFu) << 4) | (((fr >> 8) & 0xFu) << 8) | (((fr >> 4) & 0xFu) << 12);
  const int fl = p.front_line < 1 ? 1 : p.front_line;
  empty16 &= fl > 4 ? 0u : (0xFFFFu >> ((fl - 1) * 4));
  int n_empty = __popc(empty16);
  #pragma unroll 1
  for (int ci = 0; ci < p.n_hand && ci < SB_HAND_MAX; ci++) {
    const DCard& c = CARD(g, p.hand[ci].card);
    if (p.hand[ci].cost > p.mana) continue;
    if (c.kind != KIND_SPELL) {
      int a0 = 16 * ci;  // 16-bit field at bit a0 (never straddles more than two words)
      m[a0 >> 5] |= empty16 << (a0 & 31);
      n_play += n_empty;
    } else if (!(c.flags & DCF_TARGET)) {
      mask_set(m, 64 + 21 * ci); n_play++;
    } else {
      i8 tg[24];
      Target t = card_target(c);
      int nt = get_targets(g, g.current_order, t, PT_NONE, tg);
      #pragma unroll 1
      for (int i = 0; i < nt; i++) {
        if (is_base_pt(tg[i])) continue;
        mask_set(m, 65 + 21 * ci + (4 - PTY(tg[i])) * 4 + PTX(tg[i])); n_play++;
      }
    }
  }
  int n = n_play;
  if (p.replacable) for (int ci = 0; ci < p.n_hand && ci < SB_HAND_MAX; ci++) { mask_set(m, 148 + ci); n++; }
  if (n_play == 0) { mask_set(m, 155); n++; }
  return n;
}
SBD_FI bool have_winner(const G& g) { return g.pl[0].base < 0 || g.pl[1].base < 0; }

// ---------------------------------------------------------------- scripted opponent (games/stormbound.py:563-637)
// Draws its choices from the GAME's stream (self.random), so it advances g.draw.
SBD_FI bool mask_any(const u32* m, int lo, int hi) {  // any legal action in [lo, hi]
  #pragma unroll 1
  for (int a = lo; a <= hi; a++) if (m[a >> 5] >> (a & 31) & 1u) return true;
  return false;
}
SBD_FI int place_action(int ci, int pt) {  // Action.to_int PLACE (games/stormbound.py:261-270): row 0 is not encodable
  const int y = PTY(pt);
  return (y >= 1 && y <= 4) ? 16 * ci + (4 - y) * 4 + PTX(pt) : SB_ACTION_PASS;
}
SBD_NI int expert_action(G& g) {
  G_LOCAL(g);
  u32 m[SB_MASK_WORDS];
  const Ply& p = g.pl[g.local_order];
  legal_mask(g, m);
  if (mask_any(m, 148, 151)) {
    if (p.n_hand == 0) { GERR(g, SB_ERR_EMPTY_CHOICE); return SB_ACTION_PASS; }  // max([])
    int max_cost = -1000, ns = 0;
    i8 sel[HAND_W];
    #pragma unroll 1
    for (int i = 0; i < p.n_hand; i++) if (p.hand[i].cost > max_cost) max_cost = p.hand[i].cost;
    if (max_cost > p.mana) {
      #pragma unroll 1
      for (int i = 0; i < p.n_hand; i++) if (p.hand[i].cost == max_cost) sel[ns++] = (i8)i;
      return 148 + sel[rng_below(g, ns)];
    }
  }
  i8 playable[4];
  int np = 0;
  #pragma unroll 1
  for (int i = 0; i < 4; i++) if (mask_any(m, 16 * i, 16 * i + 15) || mask_any(m, 21 * i + 64, 21 * i + 84)) playable[np++] = (i8)i;
  if (np == 0) return SB_ACTION_PASS;
  bool any_eq = false;
  int min_cost = 1 << 20, ns = 0;
  i8 sel[4];
  #pragma unroll 1
  for (int k = 0; k < np; k++) { const int c = p.hand[playable[k]].cost; any_eq |= c == p.mana; min_cost = c < min_cost ? c : min_cost; }
  const int want = any_eq ? (int)p.mana : min_cost;
  #pragma unroll 1
  for (int k = 0; k < np; k++) if (p.hand[playable[k]].cost == want) sel[ns++] = (i8)k;
  const int index = playable[sel[rng_below(g, ns)]];
  const DCard& c = CARD(g, p.hand[index].card);
  i8 en[24], cand[48];
  const int n = get_targets(g, g.current_order, mkT(TK_UNIT, TS_ENEMY), PT_NONE, en);
  int nb = 0, nc = 0;
  #pragma unroll 1
  for (int i = 0; i < n; i++) nb += PTY(en[i]) == 4;
  if (c.kind == KIND_SPELL) {
    if (!(c.flags & DCF_TARGET)) return 64 + 21 * index;
    i8 tg[24];
    const int nt = get_targets(g, g.current_order, card_target(c), PT_NONE, tg);
    if (nt == 0) { GERR(g, SB_ERR_EMPTY_CHOICE); return SB_ACTION_PASS; }
    const int where = tg[rng_below(g, nt)];
    return is_base_pt(where) ? SB_ACTION_PASS : 65 + 21 * index + (4 - PTY(where)) * 4 + PTX(where);
  }
  if (c.kind == KIND_UNIT && nb > 0) {  // an enemy unit stands next to the base: block beside it
    #pragma unroll 1
    for (int i = 0; i < n; i++) {
      const int x = PTX(en[i]), y = PTY(en[i]);
      if (y != 4) continue;
      if (x > 0 && at_xy(g, x - 1, y) < 0) cand[nc++] = (i8)PT(x - 1, y);
      else if (x < 3 && at_xy(g, x + 1, y) < 0) cand[nc++] = (i8)PT(x + 1, y);
    }
  } else {
    const int fl = p.front_line;
    #pragma unroll 1
    for (int x = 0; x < 4; x++) if (valid_xy(x, fl) && at_xy(g, x, fl) < 0) cand[nc++] = (i8)PT(x, fl);
    #pragma unroll 1
    for (int i = 0; i < n; i++) {
      const int x = PTX(en[i]), y = PTY(en[i]);
      if (x > 0 && y >= fl && at_xy(g, x - 1, y) < 0) cand[nc++] = (i8)PT(x - 1, y);
      else if (x < 3 && y >= fl && at_xy(g, x + 1, y) < 0) cand[nc++] = (i8)PT(x + 1, y);
      else if (y < 4 && y + 1 >= fl && at_xy(g, x, y + 1) < 0) cand[nc++] = (i8)PT(x, y + 1);
    }
  }
  return nc > 0 ? place_action(index, cand[rng_below(g, nc)]) : SB_ACTION_PASS;
}

SBD_NI void game_step(G& g, int action) {
  G_LOCAL(g);
  Ply& p = g.pl[g.local_order];
  if (action < 64) {
    int ci = action >> 4, idx = action & 15;
    if (ci >= p.n_hand) GERR(g, SB_ERR_INDEX);
    else { p.mana = (i16)(p.mana - p.hand[ci].cost); player_play(g, g.local_order, ci, PT(idx & 3, 4 - (idx >> 2))); }
  } else if (action < 148) {
    int ci = (action - 64) / 21, idx = (action - 64) % 21;
    if (idx < 20) {  // index 20 matches no tile: complete no-op (Q5)
      if (ci >= p.n_hand) GERR(g, SB_ERR_INDEX);
      else {
        const DCard& c = CARD(g, p.hand[ci].card);
        if (c.kind != KIND_SPELL) GERR(g, SB_ERR_INDEX);
        else {
          p.mana = (i16)(p.mana - p.hand[ci].cost);
          player_play(g, g.local_order, ci, (c.flags & DCF_TARGET) ? PT(idx & 3, 4 - (idx >> 2)) : PT_NONE);
        }
      }
    }
  } else if (action < 152) {
    int ci = action - 148;
    if (ci >= p.n_hand) GERR(g, SB_ERR_INDEX);
    else { player_cycle(g, g.local_order, ci); p.replacable = 0; }
  } else if (action < 155) {
    int ci = action - 151;
    if (ci >= p.n_hand) GERR(g, SB_ERR_INDEX);
    else { CardRec t = p.hand[ci]; p.hand[ci] = p.hand[0]; p.hand[0] = t; p.leftmost = 0; }
  }
  bool done = have_winner(g);  // legal_actions() is never empty (PASS is added when nothing is playable)
  bool reward = g.pl[1 - g.local_order].base <= 0;
  g.done = (u8)((done ? SB_DONE : 0) | (reward ? SB_REWARD : 0));
  if (action == 155) {
    g.turn++; g.draw = 0;  // stream key (turn, draw)
    g.player_sign = (i8)-g.player_sign;
    board_flip(g);
    to_next_turn(g);
  }
  g.steps++;
}

// Between steps only on-board entities matter: rebuild the pool in tile order (what pack+unpack would do).
SBD_NI void compact(G& g);
// Fast path between two steps of an in-kernel rollout: nothing references an off-board entity any more
// (the trigger stack is empty, no B005 memory, no B305 board-instance records), so the garbage can stay
// until the pool no longer guarantees the 28 free slots a fresh unpack would give.
SBD_FI void end_of_step(G& g) {
  // what the packed layout cannot hold is an overflow here too (keeps in-kernel rollouts == step-per-launch)
  if (g.pl[0].n_hand > SB_HAND_MAX || g.pl[1].n_hand > SB_HAND_MAX || g.pl[0].n_deck > SB_DECK_MAX || g.pl[1].n_deck > SB_DECK_MAX)
    GERR(g, SB_ERR_OVERFLOW);
  if (g.n_ent > SB_N_TILES || g.n_mem || g.n_obj) compact(g);
  else { g.n_trig = 0; g.resolving = 0; g.depth = 0; }
}
SBD_NI void compact(G& g) {
  G_LOCAL(g);
  u8 remap[MAXE];
  #pragma unroll 1
  for (int i = 0; i < g.n_ent; i++) remap[i] = 0xFF;
  Ent tmp[SB_N_TILES];
  int n = 0;
  #pragma unroll 1
  for (int t = 0; t < 20; t++) {
    int id = g.board[t];
    if (id < 0) continue;
    remap[id] = (u8)n;
    tmp[n] = g.e[id];
    tmp[n].path_len = 0; tmp[n].move_id = 0; tmp[n].dmg = 0; tmp[n].fl &= ~(EF_RPLAY | EF_SINGLE);
    #pragma unroll 1
    for (int k = 0; k < 5; k++) if (tmp[n].st[k] > 63) tmp[n].st[k] = 63;
    g.board[t] = (i8)n;
    n++;
  }
  // frozen strength of board-instance card records whose object left the board (n_obj is an upper bound: 0 = none)
  #pragma unroll 1
  for (int o = 0; o < 2 && g.n_obj; o++) {
    Ply& p = g.pl[o];
    #pragma unroll 1
    for (int i = 0; i < p.n_hand; i++) if (p.hand[i].link >= 0) {
      u8 r = remap[p.hand[i].link];
      if (r == 0xFF) { p.hand[i].xstr = g.e[p.hand[i].link].strength; p.hand[i].link = -1; } else p.hand[i].link = (i8)r;
    }
    #pragma unroll 1
    for (int i = 0; i < p.n_deck; i++) if (p.deck[i].link >= 0) {
      u8 r = remap[p.deck[i].link];
      if (r == 0xFF) { p.deck[i].xstr = g.e[p.deck[i].link].strength; p.deck[i].link = -1; } else p.deck[i].link = (i8)r;
    }
  }
  int w = 0;
  {  // a memory survives iff the temple at the root of its tree is still on the board
    u8 keep[NMEM], nidx[NMEM];
    #pragma unroll 1
    for (int i = 0; i < g.n_mem; i++) {
      const Mem& m = g.mem[i];
      bool k = m.parent < 0 ? (m.b005 >= 0 && remap[m.b005] != 0xFF) : (keep[m.parent] != 0);
      keep[i] = k; nidx[i] = k ? (u8)w++ : (u8)0xFF;
    }
    #pragma unroll 1
    for (int i = 0; i < g.n_mem; i++) if (keep[i]) {
      Mem m = g.mem[i];
      if (m.parent >= 0) m.parent = (i8)nidx[m.parent]; else m.b005 = (i8)remap[m.b005];
      g.mem[nidx[i]] = m;
    }
  }
  g.n_mem = (u8)w;
  // what would not fit the packed layout is an overflow there too (keeps rollouts == step-per-launch)
  int nobj = 0;
  #pragma unroll 1
  for (int o = 0; o < 2 && g.n_obj; o++) {
    #pragma unroll 1
    for (int i = 0; i < g.pl[o].n_hand; i++) nobj += (g.pl[o].hand[i].flags & SB_CF_OBJ) != 0;
    #pragma unroll 1
    for (int i = 0; i < g.pl[o].n_deck; i++) nobj += (g.pl[o].deck[i].flags & SB_CF_OBJ) != 0;
  }
  if (w > NMEM_PACKED || nobj > NOBJ_PACKED) GERR(g, SB_ERR_OVERFLOW);
  g.n_obj = (u8)(nobj > 255 ? 255 : nobj);
  #pragma unroll 1
  for (int i = 0; i < n; i++) g.e[i] = tmp[i];
  g.n_ent = (u8)n;
  g.n_trig = 0; g.resolving = 0; g.depth = 0;
}
