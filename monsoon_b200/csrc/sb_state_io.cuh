// sb_state_io.cuh -- packed 512-byte state <-> thread-private working set, plus the observation
// (games/stormbound.py:400-526) and the ten StateFeatures (evo/features.py:12-342) computed straight
// from the working set.
#pragma once
#include "sb_engine.cuh"

// 128-bit vectorised copy of one packed state (global or shared <-> local)
SBD_FI void load_state(SbState& dst, const void* src) {
  const uint4* s = reinterpret_cast<const uint4*>(src);
  uint4* d = reinterpret_cast<uint4*>(&dst);
#pragma unroll 8
  for (int i = 0; i < SB_STATE_BYTES / 16; i++) d[i] = s[i];
}
SBD_FI void store_state(void* dst, const SbState& src) {
  uint4* d = reinterpret_cast<uint4*>(dst);
  const uint4* s = reinterpret_cast<const uint4*>(&src);
#pragma unroll 8
  for (int i = 0; i < SB_STATE_BYTES / 16; i++) d[i] = s[i];
}

SBD_NI void unpack(G& g, const SbState& s) {
  G_LOCAL(g);
  g.seed_lo = s.seed_lo; g.seed_hi = s.seed_hi; g.turn = s.turn; g.draw = s.draw; g.steps = s.steps;
  g.local_order = s.local_order; g.current_order = s.current_order; g.player_sign = s.player_sign;
  g.phase = s.phase; g.err = s.err; g.done = s.done; g.hist_n = s.hist_n;
#pragma unroll
  for (int i = 0; i < 4; i++) { g.hist_card[i] = s.hist_card[i]; g.hist_owner[i] = s.hist_owner[i]; }
  g.n_ent = 0; g.n_trig = 0; g.resolving = 0; g.depth = 0; g.n_mem = 0; g.n_obj = 0; g.occ = 0; g.own1 = 0; g.strc = 0;
  g.maybe_badobs = 1;  // conservative until scan_badobs() has looked
  #pragma unroll 1
  for (int o = 0; o < 2; o++) {
    const SbPlayer& sp = s.pl[o];
    Ply& p = g.pl[o];
    p.base = sp.base; p.max_mana = sp.max_mana; p.mana = sp.mana; p.front_line = sp.front_line;
    p.replacable = (sp.flags & SB_PF_REPLACABLE) != 0; p.leftmost = (sp.flags & SB_PF_LEFTMOST) != 0;
    p.n_hand = sp.n_hand; p.n_deck = sp.n_deck; p.faction = sp.faction;
    const int nh = sp.n_hand < SB_HAND_MAX ? sp.n_hand : SB_HAND_MAX, nd = sp.n_deck < SB_DECK_MAX ? sp.n_deck : SB_DECK_MAX;
    #pragma unroll 1
    for (int i = 0; i < nh; i++) {  // records beyond n_hand / n_deck are never read
      CardRec& c = p.hand[i];
      c.card = sp.hand_card[i]; c.cost = sp.hand_cost[i]; c.flags = sp.hand_flags[i]; c.link = -1; c.wn = 0; c.xstr = 0;
    }
    #pragma unroll 1
    for (int i = 0; i < nd; i++) {
      CardRec& c = p.deck[i];
      c.card = sp.deck_card[i]; c.cost = sp.deck_cost[i]; c.flags = sp.deck_flags[i]; c.link = -1; c.wn = sp.deck_wn[i]; c.xstr = 0;
    }
  }
  #pragma unroll 1
  for (int t = 0; t < SB_N_TILES; t++) {
    const SbTile& st = s.tile[t];
    g.board[t] = -1;
    if (!st.card) continue;
    int id = g.n_ent++;
    Ent& e = g.e[id];
    e.card = st.card;
    e.fl = (u8)(((st.flags & SB_TF_OWNER) ? EF_OWNER : 0) | ((st.flags & SB_TF_STRUCTURE) ? EF_STRUCT : 0) |
                ((st.flags & SB_TF_FIXED) ? EF_FIXED : 0));
    e.strength = st.strength; e.dmg = 0;
#pragma unroll
    for (int k = 0; k < 5; k++) e.st[k] = (u8)((st.status >> (SB_ST_BITS * k)) & 63);
    e.move_id = 0; e.x = (u8)(t & 3); e.y = (u8)(t >> 2); e.path_len = 0;
    g.board[t] = (i8)id;
    g.occ |= 1u << t;
    if (e.fl & EF_OWNER) g.own1 |= 1u << t;
    if (e.fl & EF_STRUCT) g.strc |= 1u << t;
  }
  const u8* x = s.ext;
  int nm = x[0];
  #pragma unroll 1
  for (int i = 0; i < nm && i < NMEM_PACKED; i++) {
    const u8* r = x + 1 + 10 * i;
    Mem& m = g.mem[g.n_mem++];
    if (r[0] & 0x80) { m.parent = (i8)(r[0] & 0x7F); m.b005 = -1; }  // memory of the remembered temple copy #parent
    else { m.parent = -1; m.b005 = (i8)at_pt(g, r[0]); }
    m.pos = r[1]; m.card = r[2];
    m.fl = (u8)(((r[3] & SB_TF_OWNER) ? EF_OWNER : 0) | ((r[3] & SB_TF_STRUCTURE) ? EF_STRUCT : 0) | ((r[3] & SB_TF_FIXED) ? EF_FIXED : 0) |
                ((r[3] & 8) ? EF_SINGLE : 0));  // bit 3 = detached copy
    m.strength = (i16)(r[4] | (r[5] << 8));
    u32 w = r[6] | (r[7] << 8) | (r[8] << 16) | ((u32)r[9] << 24);
#pragma unroll
    for (int k = 0; k < 5; k++) m.st[k] = (u8)((w >> (SB_ST_BITS * k)) & 63);
  }
  int no = x[91];
  g.n_obj = (u8)no;
  #pragma unroll 1
  for (int i = 0; i < no && i < NOBJ_PACKED; i++) {
    const u8* r = x + 92 + 4 * i;
    Ply& p = g.pl[r[0] >> 7];
    int idx = r[0] & 63;
    if ((r[0] & 64) ? idx >= SB_DECK_MAX : idx >= SB_HAND_MAX) continue;
    CardRec& c = (r[0] & 64) ? p.deck[idx] : p.hand[idx];
    if (r[1] != 0xFF) c.link = (i8)at_pt(g, r[1]);
    else { c.link = -1; c.xstr = (i16)(r[2] | (r[3] << 8)); }
  }
}

// one memory tree in pre-order (explicit stack; key = owning temple tile, or 0x80 | packed index of the parent copy)
SBD_NI void pack_mem(const G& g, SbState& s, int root, int root_key, int& nm) {
  G_LOCAL(g);
  i8 st_idx[NMEM];
  u8 st_key[NMEM];
  int sp = 0;
  st_idx[sp] = (i8)root; st_key[sp] = (u8)root_key; sp++;
  #pragma unroll 1
  while (sp > 0) {
    sp--;
    const int i = st_idx[sp];
    const int key = st_key[sp];
    if (nm >= NMEM_PACKED) { if (!s.err) s.err = SB_ERR_OVERFLOW; return; }
    const Mem& m = g.mem[i];
    const int me = nm++;
    u8* r = s.ext + 1 + 10 * me;
    r[0] = (u8)key; r[1] = m.pos; r[2] = m.card;
    r[3] = (u8)(((m.fl & EF_OWNER) ? SB_TF_OWNER : 0) | ((m.fl & EF_STRUCT) ? SB_TF_STRUCTURE : 0) | ((m.fl & EF_FIXED) ? SB_TF_FIXED : 0) |
               ((m.fl & EF_SINGLE) ? 8 : 0));
    r[4] = (u8)(m.strength & 255); r[5] = (u8)((m.strength >> 8) & 255);
    u32 w = 0;
    if (!(m.fl & EF_STRUCT)) {
      #pragma unroll 1
      for (int k = 0; k < 5; k++) w |= (u32)(m.st[k] > 63 ? 63 : m.st[k]) << (SB_ST_BITS * k);
    }
    r[6] = (u8)(w & 255); r[7] = (u8)((w >> 8) & 255); r[8] = (u8)((w >> 16) & 255); r[9] = (u8)((w >> 24) & 255);
    #pragma unroll 1
    for (int q = g.n_mem - 1; q > i; q--)  // children pushed in reverse so the lowest index pops first
      if (g.mem[q].parent == i && sp < NMEM) { st_idx[sp] = (i8)q; st_key[sp] = (u8)(0x80 | me); sp++; }
  }
}
SBD_NI void pack(const G& g, SbState& s) {
  G_LOCAL(g);
  uint4* z = reinterpret_cast<uint4*>(&s);
#pragma unroll 8
  for (int i = 0; i < SB_STATE_BYTES / 16; i++) z[i] = make_uint4(0, 0, 0, 0);
  s.seed_lo = g.seed_lo; s.seed_hi = g.seed_hi; s.turn = g.turn; s.draw = g.draw; s.steps = g.steps;
  s.local_order = g.local_order; s.current_order = g.current_order; s.player_sign = g.player_sign;
  s.phase = g.phase; s.err = g.err; s.done = g.done; s.hist_n = g.hist_n;
#pragma unroll
  for (int i = 0; i < 4; i++) { s.hist_card[i] = g.hist_card[i]; s.hist_owner[i] = g.hist_owner[i]; }
  #pragma unroll 1
  for (int o = 0; o < 2; o++) {
    SbPlayer& sp = s.pl[o];
    const Ply& p = g.pl[o];
    sp.base = p.base; sp.max_mana = p.max_mana; sp.mana = p.mana; sp.front_line = p.front_line;
    sp.flags = (u8)((p.replacable ? SB_PF_REPLACABLE : 0) | (p.leftmost ? SB_PF_LEFTMOST : 0));
    sp.n_hand = p.n_hand; sp.n_deck = p.n_deck; sp.faction = p.faction;
    if (p.n_hand > SB_HAND_MAX || p.n_deck > SB_DECK_MAX) {  // more than the packed layout holds
      if (!s.err) s.err = SB_ERR_OVERFLOW;
      if (p.n_hand > SB_HAND_MAX) sp.n_hand = SB_HAND_MAX;
      if (p.n_deck > SB_DECK_MAX) sp.n_deck = SB_DECK_MAX;
    }
    #pragma unroll 1
    for (int i = 0; i < p.n_hand && i < SB_HAND_MAX; i++) { sp.hand_card[i] = p.hand[i].card; sp.hand_cost[i] = p.hand[i].cost; sp.hand_flags[i] = p.hand[i].flags; }
    #pragma unroll 1
    for (int i = 0; i < p.n_deck && i < SB_DECK_MAX; i++) { sp.deck_card[i] = p.deck[i].card; sp.deck_cost[i] = p.deck[i].cost; sp.deck_flags[i] = p.deck[i].flags; sp.deck_wn[i] = p.deck[i].wn; }
  }
  #pragma unroll 1
  for (int t = 0; t < SB_N_TILES; t++) {
    int id = g.board[t];
    if (id < 0) continue;
    const Ent& e = g.e[id];
    SbTile& st = s.tile[t];
    st.card = e.card;
    st.flags = (u8)(((e.fl & EF_OWNER) ? SB_TF_OWNER : 0) | ((e.fl & EF_STRUCT) ? SB_TF_STRUCTURE : 0) | ((e.fl & EF_FIXED) ? SB_TF_FIXED : 0));
    st.strength = e.strength;
    u32 w = 0;
    if (!(e.fl & EF_STRUCT)) {
#pragma unroll
      for (int k = 0; k < 5; k++) w |= (u32)(e.st[k] > 63 ? 63 : e.st[k]) << (SB_ST_BITS * k);
    }
    st.status = w;
  }
  u8* x = s.ext;
  int nm = 0;
  #pragma unroll 1
  for (int tile = 0; tile < SB_N_TILES && g.n_mem; tile++) {  // canonical order: temples in tile order, each memory followed by its subtree
    int bid = g.board[tile];
    if (bid < 0 || g.e[bid].card != SBC_B005) continue;
    #pragma unroll 1
    for (int i = 0; i < g.n_mem; i++)
      if (g.mem[i].parent < 0 && g.mem[i].b005 == bid) pack_mem(g, s, i, tile, nm);
  }
  x[0] = (u8)nm;
  int no = 0;
  #pragma unroll 1
  for (int o = 0; o < 2 && g.n_obj; o++) for (int where = 0; where < 2; where++) {  // n_obj is an upper bound: 0 = no such record
    const Ply& p = g.pl[o];
    int cnt = where ? p.n_deck : p.n_hand;
    #pragma unroll 1
    for (int i = 0; i < cnt; i++) {
      const CardRec& c = where ? p.deck[i] : p.hand[i];
      if (!(c.flags & SB_CF_OBJ)) continue;
      if (no >= NOBJ_PACKED) { if (!s.err) s.err = SB_ERR_OVERFLOW; break; }
      u8* r = x + 92 + 4 * no++;
      r[0] = (u8)((o << 7) | (where << 6) | i);
      bool on_board = c.link >= 0 && g.board[g.e[c.link].y * 4 + g.e[c.link].x] == c.link;
      int str = c.link >= 0 ? g.e[c.link].strength : c.xstr;
      r[1] = on_board ? (u8)PT(g.e[c.link].x, g.e[c.link].y) : (u8)0xFF;
      r[2] = on_board ? (u8)0 : (u8)(str & 255); r[3] = on_board ? (u8)0 : (u8)((str >> 8) & 255);
    }
  }
  x[91] = (u8)no;
}

SBD_FI unsigned long long digest_state(const SbState& s) {  // FNV-1a 64 over the 512 bytes
  const u8* b = reinterpret_cast<const u8*>(&s);
  unsigned long long h = 0xCBF29CE484222325ull;
  #pragma unroll 1
  for (int i = 0; i < SB_STATE_BYTES; i++) { h ^= b[i]; h *= 0x100000001B3ull; }
  return h;
}

// ---------------------------------------------------------------- features (evo/features.py) without the 27x5x4 detour
// Every value is what StateFeatures would read from get_observation() of this state; FP64 with explicit
// round-to-nearest ops in the reference's accumulation order (no FMA contraction).
SBD_FI int card_strength_of(const G& g, const CardRec& c) {
  if (c.flags & SB_CF_OBJ) return c.link >= 0 ? g.e[c.link].strength : c.xstr;
  return CARD(g, c.card).strength;
}
SBD_FI double clip01(double v) { return v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v); }
// k / 5.0 for k = 1..5 (row distance weights, evo/features.py): the correctly rounded quotients as constants instead of
// one FP64 division per entity on the board
// (assembled from the bit patterns with integer selects: the ternary chain on doubles compiled to a jump table, and
// the lanes of a warp ask for different rows)
SBD_FI double fifths(int k) {
  static_assert(sizeof(double) == 8, "IEEE double");
  // 1/5 = 0x3FC999999999999A, 2/5 = 0x3FD999999999999A, 3/5 = 0x3FE3333333333333, 4/5 = 0x3FE999999999999A, 1 = 0x3FF0000000000000
  u32 hi = 0x3FF00000u, lo = 0u;
  hi = k == 4 ? 0x3FE99999u : hi;  lo = k == 4 ? 0x9999999Au : lo;
  hi = k == 3 ? 0x3FE33333u : hi;  lo = k == 3 ? 0x33333333u : lo;
  hi = k == 2 ? 0x3FD99999u : hi;  lo = k == 2 ? 0x9999999Au : lo;
  hi = k == 1 ? 0x3FC99999u : hi;  lo = k == 1 ? 0x9999999Au : lo;
  return __hiloint2double((int)hi, (int)lo);
}
// The card ids of a game are a closed set (every card in play descends from the two decks; copies keep the id), so
// one look at everything a freshly unpacked state holds tells whether ANY later state can contain a card without an
// observation id.  Almost always none can, and features() skips its per-call id scans.
SBD_NI void scan_badobs(G& g) {
  G_LOCAL(g);
  bool any = false;
  #pragma unroll 1
  for (int o = 0; o < 2; o++) {
    const Ply& p = g.pl[o];
    #pragma unroll 1
    for (int i = 0; i < p.n_hand; i++) any |= CARD(g, p.hand[i].card).obs_id == -32768;
    #pragma unroll 1
    for (int i = 0; i < p.n_deck; i++) any |= CARD(g, p.deck[i].card).obs_id == -32768;
  }
  #pragma unroll 1
  for (int i = 0; i < g.n_ent; i++) any |= CARD(g, g.e[i].card).obs_id == -32768;
  #pragma unroll 1
  for (int i = 0; i < g.n_mem; i++) any |= CARD(g, g.mem[i].card).obs_id == -32768;
  #pragma unroll 1
  for (int i = 0; i < g.hist_n; i++) any |= CARD(g, g.hist_card[i]).obs_id == -32768;
  g.maybe_badobs = any ? 1 : 0;
}
// returns 0 or SB_ERR_OBS_ID (int(card) raises for UP01-03 anywhere on board, in hand, deck or history: Q12)
SBD_NI int features(const G& g, double* f) {
  G_LOCAL(g);
  P_LOCAL(f);
  int err = 0;
  const int lo = g.local_order;
  const Ply& L = g.pl[lo];
  const Ply& R = g.pl[1 - lo];
  double m = L.mana != -1 ? (double)L.mana : 0.0;
  double hl = L.base != -1 ? (double)L.base : 20.0;
  double hr = R.base != -1 ? (double)R.base : 20.0;
  double est = __dadd_rn(m, 2.0);
  if (est < 3.0) est = 3.0;
  if (est > 10.0) est = 10.0;
  f[0] = clip01(__dsub_rn(1.0, ddiv(m, est)));
  f[1] = __dsub_rn(hl, hr);
  // The forks of one decision share most tiles, so the owner / kind branches below are mostly uniform across a warp
  // (ncu source view: ~53 warp instructions per tile against 83 for a select-only form).  Sums of i16 strengths fit an int.
  int sl = 0, sr = 0;
  int nl = 0, nr = 0, nsl = 0, nsr = 0, minl = 99, maxr = -1;
  double threat = 0.0, prot = 0.0;
  const bool check_ids = g.maybe_badobs != 0;
  u32 occ = g.occ;
  #pragma unroll 1
  while (occ) {  // occupied tiles in ascending order (the accumulation order of the reference's plane scan)
    const int t = __ffs(occ) - 1;
    occ &= occ - 1;
    const Ent& e = g.e[g.board[t]];
    const int y = t >> 2;
    if (check_ids && CARD(g, e.card).obs_id == -32768) err = SB_ERR_OBS_ID;
    // the observation uses -1 as "empty": an entity whose strength is exactly -1 would vanish; strengths are >= 0
    const bool counted = e.strength != -1;
    if (ent_owner(e) == lo) {
      if (!ent_struct(e)) { nl++; if (y < minl) minl = y; } else nsl++;
      if (counted) { sl += e.strength; prot = __dadd_rn(prot, __dmul_rn((double)e.strength, fifths(5 - y))); }
    } else {
      if (!ent_struct(e)) {
        nr++; if (y > maxr) maxr = y;
        if (counted) threat = __dadd_rn(threat, __dmul_rn((double)e.strength, fifths(y + 1)));
      } else nsr++;
      if (counted) sr += e.strength;
    }
  }
  const int tot = sl + sr;
  f[2] = tot == 0 ? 0.0 : ddiv((double)(sl - sr), (double)tot);
  f[3] = (nl == 0 && nr == 0) ? 0.0 : __dmul_rn((double)((nr ? maxr : 0) - (nl ? minl : 4)), 0.25);  // /4: exact scaling
  f[4] = (double)(sl - sr);
  f[5] = (double)(nl - nr);
  f[6] = (double)(nsl - nsr);
  f[7] = threat;
  f[8] = prot;
  int playable = 0, valid = 0;
  double total = 0.0;
  #pragma unroll 1
  for (int i = 0; i < L.n_hand && i < 4; i++) {
    const DCard& c = CARD(g, L.hand[i].card);
    if (c.obs_id == -32768) err = SB_ERR_OBS_ID;
    if (c.obs_id == -1 || c.obs_id == 32767) continue;
    int cost = L.hand[i].cost;
    int str = c.kind == KIND_SPELL ? 0 : card_strength_of(g, L.hand[i]);
    if (str == -1) str = 0;
    valid++;
    if (cost > 0) {
      total = __dadd_rn(total, ddiv((double)str, (double)cost));
      if ((double)cost <= m) playable++;
    }
  }
  if (valid == 0) f[9] = 0.0;
  else {
    // valid is 1..4: divisions by 1, 2 and 4 are exact scalings, only /3 needs the divider
    double playability, avg;
    if (valid == 3) { playability = ddiv((double)playable, 3.0); avg = ddiv(total, 3.0); }
    else { const double inv = valid == 1 ? 1.0 : valid == 2 ? 0.5 : 0.25; playability = __dmul_rn((double)playable, inv); avg = __dmul_rn(total, inv); }
    f[9] = __dmul_rn(__dadd_rn(playability, clip01(ddiv(avg, 3.0))), 0.5);  // /2: exact scaling
  }
  if (check_ids) {
    #pragma unroll 1
    for (int i = 0; i < L.n_deck; i++) if (CARD(g, L.deck[i].card).obs_id == -32768) err = SB_ERR_OBS_ID;
    #pragma unroll 1
    for (int i = 0; i < g.hist_n; i++) if (CARD(g, g.hist_card[i]).obs_id == -32768) err = SB_ERR_OBS_ID;
  }
  return err;
}

// ---------------------------------------------------------------- observation (games/stormbound.py:400-526)
#define OBSI(l, r, c) obs[((l) * 5 + (r)) * 4 + (c)]
SBD_FI void obs_card_row(const G& g, int* obs, int layer, int row, const CardRec& c, int& err) {
  const DCard& d = CARD(g, c.card);
  if (d.obs_id == -32768) err = SB_ERR_OBS_ID;
  OBSI(layer, row, 0) = d.obs_id;
  OBSI(layer, row, 1) = c.cost;
  OBSI(layer, row, 2) = d.kind == KIND_SPELL ? -1 : card_strength_of(g, c);
  OBSI(layer, row, 3) = d.kind == KIND_UNIT ? d.movement : -1;
}
SBD_NI int observe(const G& g, int* obs) {
  G_LOCAL(g);
  int err = 0;
  #pragma unroll 1
  for (int i = 0; i < SB_OBS_INTS; i++) obs[i] = -1;
  const int lo = g.local_order;
  #pragma unroll 1
  for (int t = 0; t < SB_N_TILES; t++) {
    int id = g.board[t];
    if (id < 0) continue;
    const Ent& e = g.e[id];
    const DCard& d = CARD(g, e.card);
    if (d.obs_id == -32768) err = SB_ERR_OBS_ID;
    const int base = ent_owner(e) == lo ? 0 : 16, y = t >> 2, x = t & 3;
    if (!ent_struct(e)) {
      OBSI(base + 0, y, x) = d.obs_id;
      OBSI(base + 1, y, x) = e.strength;
      OBSI(base + 2, y, x) = d.movement;
      OBSI(base + 3, y, x) = (e.st[SB_ST_VITALIZED] ? 1 : 0) | (e.st[SB_ST_POISONED] ? 2 : 0) | (e.st[SB_ST_CONFUSED] ? 4 : 0) |
                             (e.st[SB_ST_FROZEN] ? 8 : 0) | (e.st[SB_ST_DISABLED] ? 16 : 0);
    } else {
      OBSI(base + 4, y, x) = d.obs_id;
      OBSI(base + 5, y, x) = e.strength;
    }
  }
  const Ply& L = g.pl[lo];
  const Ply& R = g.pl[1 - lo];
  #pragma unroll 1
  for (int i = 0; i < L.n_hand && i < 4; i++) obs_card_row(g, obs, 6, i, L.hand[i], err);
  #pragma unroll 1
  for (int c = 0; c < 4; c++) OBSI(6, 4, c) = 32767;
  u8 idx[SB_DECK_MAX];
  #pragma unroll 1
  for (int i = 0; i < L.n_deck; i++) idx[i] = (u8)i;
  #pragma unroll 1
  for (int i = 1; i < L.n_deck; i++) {  // sorted(deck, key=(cost, card_id)), stable
    u8 v = idx[i];
    int j = i - 1;
    #pragma unroll 1
    while (j >= 0 && (L.deck[idx[j]].cost > L.deck[v].cost ||
                      (L.deck[idx[j]].cost == L.deck[v].cost && L.deck[idx[j]].card > L.deck[v].card))) { idx[j + 1] = idx[j]; j--; }
    idx[j + 1] = v;
  }
  #pragma unroll 1
  for (int layer = 0; layer < 6; layer++) {
    #pragma unroll 1
    for (int k = 0; k < 4; k++) { int d = layer * 4 + k; if (d < L.n_deck) obs_card_row(g, obs, 7 + layer, k, L.deck[idx[d]], err); }
    #pragma unroll 1
    for (int c = 0; c < 4; c++) OBSI(7 + layer, 4, c) = 32768;
  }
  #pragma unroll 1
  for (int r = 0; r < 5; r++) for (int c = 0; c < 4; c++) {
    OBSI(13, r, c) = L.mana; OBSI(14, r, c) = L.base; OBSI(15, r, c) = L.faction;
    OBSI(22, r, c) = R.mana; OBSI(23, r, c) = R.base; OBSI(24, r, c) = R.faction;
    OBSI(25, r, c) = g.player_sign * 99999;
  }
  #pragma unroll 1
  for (int i = 0; i < 4; i++) {
    int h = i - (4 - g.hist_n);
    if (h >= 0) {
      OBSI(26, i, 0) = g.hist_owner[h] ? -99999 : 99999;
      OBSI(26, i, 1) = CARD(g, g.hist_card[h]).obs_id;
      if (CARD(g, g.hist_card[h]).obs_id == -32768) err = SB_ERR_OBS_ID;
    }
  }
  #pragma unroll 1
  for (int c = 0; c < 4; c++) OBSI(26, 4, c) = 32769;
  return err;
}
