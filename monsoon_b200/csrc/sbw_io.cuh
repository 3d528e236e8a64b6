// sbw_io.cuh -- warp-per-game engine: packed 512-byte state <-> shared-memory working set (one lane per tile / deck slot),
// the observation (games/stormbound.py:400-526), the ten StateFeatures (evo/features.py:12-342) and the scripted
// opponent (games/stormbound.py:563-637).  The packed record sits in shared (GPU) or host (tests) memory.
#pragma once
#include "sbw_effects.cuh"

SBW_NI void w_unpack(WG* wg, const SbState* s) {
  W_SHARED(wg);
  wg->seed_lo = s->seed_lo; wg->seed_hi = s->seed_hi; wg->turn = s->turn; wg->draw = s->draw; wg->steps = s->steps;
  wg->local_order = s->local_order; wg->current_order = s->current_order; wg->player_sign = s->player_sign;
  wg->phase = s->phase; wg->err = s->err; wg->done = s->done; wg->hist_n = s->hist_n;
#pragma unroll
  for (int i = 0; i < 4; i++) { wg->hist_card[i] = s->hist_card[i]; wg->hist_owner[i] = s->hist_owner[i]; }
  wg->n_trig = 0; wg->resolving = 0; wg->depth = 0; wg->n_mem = 0; wg->n_obj = 0;
  wg->maybe_badobs = 1;  // conservative until w_scan_badobs() has looked
#pragma unroll 1
  for (int o = 0; o < 2; o++) {
    const SbPlayer& sp = s->pl[o];
    WPly& p = wg->pl[o];
    p.base = sp.base; p.max_mana = sp.max_mana; p.mana = sp.mana; p.front_line = sp.front_line;
    p.replacable = (sp.flags & SB_PF_REPLACABLE) != 0; p.leftmost = (sp.flags & SB_PF_LEFTMOST) != 0;
    p.n_hand = sp.n_hand; p.n_deck = sp.n_deck; p.faction = sp.faction;
    const int nh = sp.n_hand < SB_HAND_MAX ? sp.n_hand : SB_HAND_MAX, nd = sp.n_deck < SB_DECK_MAX ? sp.n_deck : SB_DECK_MAX;
    FOR_LANES(l) {  // records beyond n_hand / n_deck are never read
      if (l < nd) { WCard c; c.card = sp.deck_card[l]; c.cost = sp.deck_cost[l]; c.flags = sp.deck_flags[l]; c.link = -1; c.wn = sp.deck_wn[l]; c.xstr = 0; p.deck[l] = c; }
      else if (l >= 24 && l - 24 < nh) { const int i = l - 24; WCard c; c.card = sp.hand_card[i]; c.cost = sp.hand_cost[i]; c.flags = sp.hand_flags[i]; c.link = -1; c.wn = 0; c.xstr = 0; p.hand[i] = c; }
    } END_LANES
  }
  // lane t owns tile t; entities are numbered in tile order
  const u32 occ = w_ballot([&](int l) -> bool { return l < SB_N_TILES && s->tile[l].card != 0; });
  wg->occ = occ;
  wg->own1 = w_ballot([&](int l) -> bool { return l < SB_N_TILES && s->tile[l].card != 0 && (s->tile[l].flags & SB_TF_OWNER); });
  wg->strc = w_ballot([&](int l) -> bool { return l < SB_N_TILES && s->tile[l].card != 0 && (s->tile[l].flags & SB_TF_STRUCTURE); });
  FOR_LANES(l) {
    if (l < SB_N_TILES) {
      const SbTile st = s->tile[l];
      if (st.card) {
        const int id = w_popc(occ & ((1u << l) - 1u));
        wg->e_card[id] = st.card;
        wg->e_fl[id] = (u8)(((st.flags & SB_TF_OWNER) ? WEF_OWNER : 0) | ((st.flags & SB_TF_STRUCTURE) ? WEF_STRUCT : 0) |
                            ((st.flags & SB_TF_FIXED) ? WEF_FIXED : 0) | WEF_ONB);
        wg->e_str[id] = st.strength; wg->e_dmg[id] = 0;
        wg->e_st[id] = st.status & 0x3FFFFFFFu;
        wg->e_mid[id] = 0; wg->e_pos[id] = (u8)l; wg->e_plen[id] = 0;
        wg->board[l] = (i8)id;
      } else wg->board[l] = -1;
    }
  } END_LANES
  wg->n_ent = (u8)w_popc(occ);
  const u8* x = s->ext;
  const int nm = x[0];
#pragma unroll 1
  for (int i = 0; i < nm && i < NMEM_PACKED; i++) {
    const u8* r = x + 1 + 10 * i;
    WMem m;
    if (r[0] & 0x80) { m.parent = (i8)(r[0] & 0x7F); m.b005 = -1; }  // memory of the remembered temple copy #parent
    else { m.parent = -1; m.b005 = (i8)w_at_pt(wg, r[0]); }
    m.pos = r[1]; m.card = r[2];
    m.fl = (u8)(((r[3] & SB_TF_OWNER) ? WEF_OWNER : 0) | ((r[3] & SB_TF_STRUCTURE) ? WEF_STRUCT : 0) | ((r[3] & SB_TF_FIXED) ? WEF_FIXED : 0) |
                ((r[3] & 8) ? WEF_SINGLE : 0));  // bit 3 = detached copy
    m.strength = (i16)(r[4] | (r[5] << 8));
    const u32 w = r[6] | (r[7] << 8) | (r[8] << 16) | ((u32)r[9] << 24);
#pragma unroll
    for (int k = 0; k < 5; k++) m.st[k] = (u8)((w >> (SB_ST_BITS * k)) & 63);
#pragma unroll
    for (int k = 0; k < 4; k++) m.pad[k] = 0;
    wg->mem[i] = m;
    wg->n_mem = (u8)(i + 1);
  }
  const int no = x[91];
  wg->n_obj = (u8)no;
#pragma unroll 1
  for (int i = 0; i < no && i < NOBJ_PACKED; i++) {
    const u8* r = x + 92 + 4 * i;
    WPly& p = wg->pl[r[0] >> 7];
    const int idx = r[0] & 63;
    if ((r[0] & 64) ? idx >= SB_DECK_MAX : idx >= SB_HAND_MAX) continue;
    WCard& c = (r[0] & 64) ? p.deck[idx] : p.hand[idx];
    if (r[1] != 0xFF) c.link = (i8)w_at_pt(wg, r[1]);
    else { c.link = -1; c.xstr = (i16)(r[2] | (r[3] << 8)); }
  }
}

// Stormbound.legal_actions (games/stormbound.py:528-557) STRAIGHT FROM THE PACKED RECORD, one lane per tile: the streaming
// form of w_unpack + w_legal_mask for the record-in / mask-out API (sb_legal_mask), ~100 warp instructions per record instead
// of ~600, so that the kernel is bound by HBM, not by issue.  m: 5 words, uniform result.  Returns the number of legal actions.
SBW_FI int w_legal_mask_packed(const SbState* s, const DCard* cards, u32* m) {
  W_SHARED(cards);
  const int lo = s->local_order;
  const SbPlayer& p = s->pl[lo];
  const u32 occ = w_ballot([&](int l) -> bool { return l < SB_N_TILES && s->tile[l].card != 0; });
#pragma unroll
  for (int i = 0; i < SB_MASK_WORDS; i++) m[i] = 0;
  const u32 fr = ~occ;
  u32 empty16 = ((fr >> 16) & 0xFu) | (((fr >> 12) & 0xFu) << 4) | (((fr >> 8) & 0xFu) << 8) | (((fr >> 4) & 0xFu) << 12);
  const int fl = p.front_line < 1 ? 1 : p.front_line;
  empty16 &= fl > 4 ? 0u : (0xFFFFu >> ((fl - 1) * 4));
  const int n_empty = w_popc(empty16);
  const int nh = p.n_hand < SB_HAND_MAX ? p.n_hand : SB_HAND_MAX;
  const int mana = p.mana;
  int n_play = 0;
#pragma unroll 1
  for (int ci = 0; ci < nh; ci++) {
    const DCard& c = cards[p.hand_card[ci]];
    if (p.hand_cost[ci] > mana) continue;
    if (c.kind != KIND_SPELL) {
      const int a0 = 16 * ci;
      m[a0 >> 5] |= empty16 << (a0 & 31);
      n_play += n_empty;
    } else if (!(c.flags & DCF_TARGET)) {
      w_mask_set(m, 64 + 21 * ci); n_play++;
    } else {  // board.get_targets(None, required_targets) on the packed tiles (board.py:147-204); pov = board.current_player
      const int pov = s->current_order;
      const int kind = c.t_ks & 3, side = c.t_ks >> 2;
      const bool has_limit = c.t_limit >= 0;
      const int limit = c.t_limit;
      const u32 want_types = c.t_types, bad_types = (u32)c.t_xtypes | ((c.flags & DCF_TNONHERO) ? (1u << UT_HERO) : 0u);
      const u32 want_status = c.t_status, bad_status = c.t_xstatus;
      u32 tm = w_ballot([&](int l) -> bool {
        if (l >= SB_N_TILES) return false;
        const SbTile t = s->tile[l];
        if (!t.card || t.strength <= 0) return false;
        const bool is_struct = (t.flags & SB_TF_STRUCTURE) != 0;
        if (kind == TK_UNIT ? is_struct : (kind == TK_STRUCTURE ? !is_struct : false)) return false;
        const bool mine = ((t.flags & SB_TF_OWNER) ? 1 : 0) == pov;
        if (side == TS_FRIENDLY ? !mine : (side == TS_ENEMY ? mine : false)) return false;
        if (has_limit && t.strength > limit) return false;
        if (!is_struct) {
          if (want_types | bad_types) {
            const u32 types = cards[t.card].types;
            if ((want_types && !(types & want_types)) || (types & bad_types)) return false;
          }
          if (want_status | bad_status) {
            u32 have = 0;
#pragma unroll
            for (int k = 0; k < 5; k++) have |= (((t.status >> (SB_ST_BITS * k)) & 63u) ? 1u : 0u) << k;
            if ((want_status && !(have & want_status)) || (have & bad_status)) return false;
          }
        }
        return true;
      });
#pragma unroll 1
      while (tm) {
        const int tile = w_ffs(tm) - 1;
        tm &= tm - 1;
        w_mask_set(m, 65 + 21 * ci + (4 - (tile >> 2)) * 4 + (tile & 3)); n_play++;
      }
    }
  }
  int n = n_play;
  if (p.flags & SB_PF_REPLACABLE) for (int ci = 0; ci < nh; ci++) { w_mask_set(m, 148 + ci); n++; }
  if (n_play == 0) { w_mask_set(m, 155); n++; }
  return n;
}

// Stormbound.get_observation (games/stormbound.py:400-526) STRAIGHT FROM THE PACKED RECORD into a 540-int buffer `ob`
// (shared memory on the GPU: the kernel then streams it out with 128-bit stores).  One lane per tile / hand card / deck card /
// history entry; the deck's sorted(key=(cost, card_id)) is a rank computed by every deck lane against all deck cards.
// Returns 0 or SB_ERR_OBS_ID (Q12).  Pure data movement: ~350 warp instructions for 512 B in, 2,160 B out.
SBW_FI int w_packed_card_strength(const SbState* s, const DCard* cards, int order, int in_deck, int idx, int card, int flags) {
  if (!(flags & SB_CF_OBJ)) return cards[card].strength;
  const u8* x = s->ext;  // board-instance card record (cards/b305.py:41-45): live link to a tile, or the frozen strength
  const int no = x[91] < NOBJ_PACKED ? x[91] : NOBJ_PACKED;
  const int key = (order << 7) | (in_deck << 6) | idx;
#pragma unroll 1
  for (int i = 0; i < no; i++) {
    const u8* r = x + 92 + 4 * i;
    if (r[0] != key) continue;
    if (r[1] != 0xFF) return s->tile[r[1] < SB_N_TILES ? r[1] : 0].strength;
    return (i16)(r[2] | (r[3] << 8));
  }
  return 0;  // w_unpack leaves link = -1, xstr = 0 for a record that did not fit ext
}
SBW_FI int w_observe_packed(const SbState* s, const DCard* cards, int* ob) {
  W_SHARED(cards);
  const int lo = s->local_order;
  const SbPlayer& L = s->pl[lo];
  const SbPlayer& R = s->pl[1 - lo];
  const int sign = s->player_sign * 99999;
  FOR_LANES(l) {
#pragma unroll 1
    for (int i = l; i < SB_OBS_INTS; i += 32) {
      const int layer = i / 20, row = (i % 20) >> 2;
      int v = -1;
      if (layer == 13) v = L.mana; else if (layer == 14) v = L.base; else if (layer == 15) v = L.faction;
      else if (layer == 22) v = R.mana; else if (layer == 23) v = R.base; else if (layer == 24) v = R.faction;
      else if (layer == 25) v = sign;
      else if (row == 4) { if (layer == 6) v = 32767; else if (layer >= 7 && layer <= 12) v = 32768; else if (layer == 26) v = 32769; }
      ob[i] = v;
    }
  } END_LANES
  const int nh = L.n_hand < SB_HAND_MAX ? L.n_hand : SB_HAND_MAX, nd = L.n_deck < SB_DECK_MAX ? L.n_deck : SB_DECK_MAX;
  const int hn = s->hist_n < 4 ? s->hist_n : 4;
  u32 bad = 0;
  FOR_LANES(l) {
    if (l < SB_N_TILES) {  // board layers 0-5 (own) / 16-21 (enemy)
      const SbTile t = s->tile[l];
      if (t.card) {
        const DCard& d = cards[t.card];
        const int base = (((t.flags & SB_TF_OWNER) ? 1 : 0) == lo ? 0 : 16) * 20 + l;
        if (!(t.flags & SB_TF_STRUCTURE)) {
          const u32 w = t.status;
          ob[base] = d.obs_id; ob[base + 20] = t.strength; ob[base + 40] = d.movement;
          ob[base + 60] = (((w >> (SB_ST_BITS * SB_ST_VITALIZED)) & 63u) ? 1 : 0) | (((w >> (SB_ST_BITS * SB_ST_POISONED)) & 63u) ? 2 : 0) |
                          (((w >> (SB_ST_BITS * SB_ST_CONFUSED)) & 63u) ? 4 : 0) | (((w >> (SB_ST_BITS * SB_ST_FROZEN)) & 63u) ? 8 : 0) |
                          (((w >> (SB_ST_BITS * SB_ST_DISABLED)) & 63u) ? 16 : 0);
        } else { ob[base + 80] = d.obs_id; ob[base + 100] = t.strength; }
      }
    } else if (l < SB_N_TILES + 4) {  // hand, layer 6
      const int i = l - SB_N_TILES;
      if (i < nh) {
        const int card = L.hand_card[i];
        const DCard& d = cards[card];
        int* o = ob + 6 * 20 + i * 4;
        o[0] = d.obs_id; o[1] = L.hand_cost[i];
        o[2] = d.kind == KIND_SPELL ? -1 : w_packed_card_strength(s, cards, lo, 0, i, card, L.hand_flags[i]);
        o[3] = d.kind == KIND_UNIT ? d.movement : -1;
      }
    } else if (l < SB_N_TILES + 8) {  // history, layer 26: the last four plays, oldest first, right-aligned
      const int i = l - SB_N_TILES - 4;
      const int h = i - (4 - hn);
      if (h >= 0) { ob[26 * 20 + i * 4] = s->hist_owner[h] ? -99999 : 99999; ob[26 * 20 + i * 4 + 1] = cards[s->hist_card[h]].obs_id; }
    }
  } END_LANES
  FOR_LANES(l) {  // deck, layers 7-12: position = rank under the stable order (cost, card id)
    if (l < nd) {
      const int card = L.deck_card[l], cost = L.deck_cost[l];
      int rank = 0;
#pragma unroll 1
      for (int j = 0; j < nd; j++) {
        const int cj = L.deck_cost[j], kj = L.deck_card[j];
        rank += (cj < cost || (cj == cost && (kj < card || (kj == card && j < l)))) ? 1 : 0;
      }
      const DCard& d = cards[card];
      int* o = ob + (7 + (rank >> 2)) * 20 + (rank & 3) * 4;
      o[0] = d.obs_id; o[1] = cost;
      o[2] = d.kind == KIND_SPELL ? -1 : w_packed_card_strength(s, cards, lo, 1, l, card, L.deck_flags[l]);
      o[3] = d.kind == KIND_UNIT ? d.movement : -1;
    }
  } END_LANES
  bad = w_ballot([&](int l) -> bool {  // int(card) raises for UP01-03 on the board, in the hand, the deck or the history (Q12)
    bool b = false;
    if (l < SB_N_TILES) { const int c = s->tile[l].card; b = c && cards[c].obs_id == -32768; }
    if (l < nd) b = b || cards[L.deck_card[l]].obs_id == -32768;
    if (l < nh) b = b || cards[L.hand_card[l]].obs_id == -32768;
    if (l < hn) b = b || cards[s->hist_card[l]].obs_id == -32768;
    return b;
  });
  return bad ? SB_ERR_OBS_ID : 0;
}

// one memory tree in pre-order (explicit stack in scratch; key = owning temple tile, or 0x80 | packed index of the parent copy)
SBW_NI void w_pack_mem(WG* wg, SbState* s, int root, int root_key, int& nm) {
  W_SHARED(wg);
  i8* st_idx = wg->scr;       // NMEM entries
  i8* st_key = wg->scr + 12;  // NMEM entries (stored as i8, read back as u8)
  int sp = 0;
  st_idx[sp] = (i8)root; st_key[sp] = (i8)root_key; sp++;
#pragma unroll 1
  while (sp > 0) {
    sp--;
    const int i = st_idx[sp];
    const int key = (u8)st_key[sp];
    if (nm >= NMEM_PACKED) { if (!s->err) s->err = SB_ERR_OVERFLOW; return; }
    const WMem m = wg->mem[i];
    const int me = nm++;
    u8* r = s->ext + 1 + 10 * me;
    r[0] = (u8)key; r[1] = m.pos; r[2] = m.card;
    r[3] = (u8)(((m.fl & WEF_OWNER) ? SB_TF_OWNER : 0) | ((m.fl & WEF_STRUCT) ? SB_TF_STRUCTURE : 0) | ((m.fl & WEF_FIXED) ? SB_TF_FIXED : 0) |
               ((m.fl & WEF_SINGLE) ? 8 : 0));
    r[4] = (u8)(m.strength & 255); r[5] = (u8)((m.strength >> 8) & 255);
    u32 w = 0;
    if (!(m.fl & WEF_STRUCT)) {
#pragma unroll 1
      for (int k = 0; k < 5; k++) w |= (u32)(m.st[k] > 63 ? 63 : m.st[k]) << (SB_ST_BITS * k);
    }
    r[6] = (u8)(w & 255); r[7] = (u8)((w >> 8) & 255); r[8] = (u8)((w >> 16) & 255); r[9] = (u8)((w >> 24) & 255);
#pragma unroll 1
    for (int q = wg->n_mem - 1; q > i; q--)  // children pushed in reverse so the lowest index pops first
      if (wg->mem[q].parent == i && sp < NMEM) { st_idx[sp] = (i8)q; st_key[sp] = (i8)(0x80 | me); sp++; }
  }
}
SBW_NI void w_pack(WG* wg, SbState* s) {
  W_SHARED(wg);
  u32* z = reinterpret_cast<u32*>(s);
  FOR_LANES(l) {
#pragma unroll
    for (int i = 0; i < SB_STATE_BYTES / 4 / 32; i++) z[i * 32 + l] = 0u;
  } END_LANES
  s->seed_lo = wg->seed_lo; s->seed_hi = wg->seed_hi; s->turn = wg->turn; s->draw = wg->draw; s->steps = wg->steps;
  s->local_order = wg->local_order; s->current_order = wg->current_order; s->player_sign = wg->player_sign;
  s->phase = wg->phase; s->err = wg->err; s->done = wg->done; s->hist_n = wg->hist_n;
#pragma unroll
  for (int i = 0; i < 4; i++) { s->hist_card[i] = wg->hist_card[i]; s->hist_owner[i] = wg->hist_owner[i]; }
#pragma unroll 1
  for (int o = 0; o < 2; o++) {
    SbPlayer& sp = s->pl[o];
    const WPly& p = wg->pl[o];
    sp.base = p.base; sp.max_mana = p.max_mana; sp.mana = p.mana; sp.front_line = p.front_line;
    sp.flags = (u8)((p.replacable ? SB_PF_REPLACABLE : 0) | (p.leftmost ? SB_PF_LEFTMOST : 0));
    sp.n_hand = p.n_hand; sp.n_deck = p.n_deck; sp.faction = p.faction;
    if (p.n_hand > SB_HAND_MAX || p.n_deck > SB_DECK_MAX) {  // more than the packed layout holds
      if (!s->err) s->err = SB_ERR_OVERFLOW;
      if (p.n_hand > SB_HAND_MAX) sp.n_hand = SB_HAND_MAX;
      if (p.n_deck > SB_DECK_MAX) sp.n_deck = SB_DECK_MAX;
    }
    const int nh = p.n_hand < SB_HAND_MAX ? p.n_hand : SB_HAND_MAX, nd = p.n_deck < SB_DECK_MAX ? p.n_deck : SB_DECK_MAX;
    FOR_LANES(l) {
      if (l < nd) { const WCard c = p.deck[l]; sp.deck_card[l] = c.card; sp.deck_cost[l] = c.cost; sp.deck_flags[l] = c.flags; sp.deck_wn[l] = c.wn; }
      else if (l >= 24 && l - 24 < nh) { const int i = l - 24; const WCard c = p.hand[i]; sp.hand_card[i] = c.card; sp.hand_cost[i] = c.cost; sp.hand_flags[i] = c.flags; }
    } END_LANES
  }
  FOR_LANES(l) {
    if (l < SB_N_TILES) {
      const int id = wg->board[l];
      if (id >= 0) {
        const u8 fl = wg->e_fl[id];
        SbTile st;
        st.card = wg->e_card[id];
        st.flags = (u8)(((fl & WEF_OWNER) ? SB_TF_OWNER : 0) | ((fl & WEF_STRUCT) ? SB_TF_STRUCTURE : 0) | ((fl & WEF_FIXED) ? SB_TF_FIXED : 0));
        st.strength = wg->e_str[id];
        st.status = (fl & WEF_STRUCT) ? 0u : wg->e_st[id];
        s->tile[l] = st;
      }
    }
  } END_LANES
  u8* x = s->ext;
  int nm = 0;
#pragma unroll 1
  for (int tile = 0; tile < SB_N_TILES && wg->n_mem; tile++) {  // canonical order: temples in tile order, each memory followed by its subtree
    const int bid = wg->board[tile];
    if (bid < 0 || wg->e_card[bid] != SBC_B005) continue;
#pragma unroll 1
    for (int i = 0; i < wg->n_mem; i++)
      if (wg->mem[i].parent < 0 && wg->mem[i].b005 == bid) w_pack_mem(wg, s, i, tile, nm);
  }
  x[0] = (u8)nm;
  int no = 0;
#pragma unroll 1
  for (int o = 0; o < 2 && wg->n_obj; o++) for (int where = 0; where < 2; where++) {  // n_obj is an upper bound: 0 = no such record
    const WPly& p = wg->pl[o];
    const int cnt = where ? p.n_deck : p.n_hand;
#pragma unroll 1
    for (int i = 0; i < cnt; i++) {
      const WCard c = where ? p.deck[i] : p.hand[i];
      if (!(c.flags & SB_CF_OBJ)) continue;
      if (no >= NOBJ_PACKED) { if (!s->err) s->err = SB_ERR_OVERFLOW; break; }
      u8* r = x + 92 + 4 * no++;
      r[0] = (u8)((o << 7) | (where << 6) | i);
      const bool on_board = c.link >= 0 && wg->board[wg->e_pos[c.link]] == c.link;
      const int str = c.link >= 0 ? (int)wg->e_str[c.link] : (int)c.xstr;
      r[1] = on_board ? wg->e_pos[c.link] : (u8)0xFF;
      r[2] = on_board ? (u8)0 : (u8)(str & 255); r[3] = on_board ? (u8)0 : (u8)((str >> 8) & 255);
    }
  }
  x[91] = (u8)no;
}

SBW_NI u64 w_digest_state(const SbState* s) {  // FNV-1a 64 over the 512 bytes (test mode only: inherently sequential)
  const u8* b = reinterpret_cast<const u8*>(s);
  u64 h = 0xCBF29CE484222325ull;
#pragma unroll 1
  for (int i = 0; i < SB_STATE_BYTES; i++) { h ^= b[i]; h *= 0x100000001B3ull; }
  return h;
}

// ---------------------------------------------------------------- features (evo/features.py) without the 27x5x4 detour
SBW_FI int w_card_strength_of(const WG* wg, const WCard& c) {
  if (c.flags & SB_CF_OBJ) return c.link >= 0 ? (int)wg->e_str[c.link] : (int)c.xstr;
  return WCARD(wg, c.card).strength;
}
SBW_FI double w_clip01(double v) { return v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v); }
// k / 5.0 for k = 1..5: the correctly rounded quotients as bit patterns
SBW_FI double w_fifths(int k) {
  u32 hi = 0x3FF00000u, lo = 0u;
  hi = k == 4 ? 0x3FE99999u : hi;  lo = k == 4 ? 0x9999999Au : lo;
  hi = k == 3 ? 0x3FE33333u : hi;  lo = k == 3 ? 0x33333333u : lo;
  hi = k == 2 ? 0x3FD99999u : hi;  lo = k == 2 ? 0x9999999Au : lo;
  hi = k == 1 ? 0x3FC99999u : hi;  lo = k == 1 ? 0x9999999Au : lo;
  return d_hilo(hi, lo);
}
// The card ids of a game are a closed set, so one look at a freshly unpacked state tells whether ANY later state can hold
// a card without an observation id (UP01-03, Q12); almost always none can, and w_features() skips its id scans.
SBW_NI void w_scan_badobs(WG* wg) {
  W_SHARED(wg);
  const DCard* cards = wg->cards;
  W_SHARED(cards);
  u32 any = 0;
#pragma unroll 1
  for (int o = 0; o < 2; o++) {
    const WPly& p = wg->pl[o];
    any |= w_ballot([&](int l) -> bool { return (l < p.n_deck && l < 24 && cards[p.deck[l].card].obs_id == -32768) ||
                                                (l >= 24 && l - 24 < p.n_hand && cards[p.hand[l - 24].card].obs_id == -32768); });
  }
  const int ne = wg->n_ent;
#pragma unroll 1
  for (int base = 0; base < ne; base += 32) any |= w_ballot([&](int l) -> bool { return base + l < ne && cards[wg->e_card[base + l]].obs_id == -32768; });
#pragma unroll 1
  for (int i = 0; i < wg->n_mem; i++) any |= cards[wg->mem[i].card].obs_id == -32768;
#pragma unroll 1
  for (int i = 0; i < wg->hist_n; i++) any |= cards[wg->hist_card[i]].obs_id == -32768;
  wg->maybe_badobs = any ? 1 : 0;
}
// returns 0 or SB_ERR_OBS_ID (int(card) raises for UP01-03 anywhere on board, in hand, deck or history: Q12).
// Integer parts by warp reductions (one lane per entity); the two FP64 sums keep the reference's accumulation order
// (ascending tiles) because their terms are not exactly representable.
SBW_NI int w_features(const WG* wg, double* f) {
  W_SHARED(wg);
  int err = 0;
  const int lo = wg->local_order;
  const WPly& L = wg->pl[lo];
  const WPly& R = wg->pl[1 - lo];
  const double m = L.mana != -1 ? (double)L.mana : 0.0;
  const double hl = L.base != -1 ? (double)L.base : 20.0;
  const double hr = R.base != -1 ? (double)R.base : 20.0;
  double est = d_add(m, 2.0);
  if (est < 3.0) est = 3.0;
  if (est > 10.0) est = 10.0;
  f[0] = w_clip01(d_sub(1.0, d_div(m, est)));
  f[1] = d_sub(hl, hr);
  int sl = 0, sr = 0;
  int nl = 0, nr = 0, nsl = 0, nsr = 0, minl = 99, maxr = -1;
  double threat = 0.0, prot = 0.0;
  const bool check_ids = wg->maybe_badobs != 0;
  u32 occ = wg->occ;
  const u32 mine = lo ? wg->own1 : ~wg->own1;
  const u32 strc = wg->strc;
#pragma unroll 1
  while (occ) {  // occupied tiles in ascending order (the accumulation order of the reference's plane scan)
    const int t = w_ffs(occ) - 1;
    occ &= occ - 1;
    const int id = wg->board[t];
    const int y = t >> 2;
    const int str = wg->e_str[id];
    if (check_ids && WCARD(wg, wg->e_card[id]).obs_id == -32768) err = SB_ERR_OBS_ID;
    // the observation uses -1 as "empty": an entity whose strength is exactly -1 would vanish; strengths are >= 0
    const bool counted = str != -1;
    const bool is_struct = (strc >> t) & 1u;
    if ((mine >> t) & 1u) {
      if (!is_struct) { nl++; if (y < minl) minl = y; } else nsl++;
      if (counted) { sl += str; prot = d_add(prot, d_mul((double)str, w_fifths(5 - y))); }
    } else {
      if (!is_struct) {
        nr++; if (y > maxr) maxr = y;
        if (counted) threat = d_add(threat, d_mul((double)str, w_fifths(y + 1)));
      } else nsr++;
      if (counted) sr += str;
    }
  }
  const int tot = sl + sr;
  f[2] = tot == 0 ? 0.0 : d_div((double)(sl - sr), (double)tot);
  f[3] = (nl == 0 && nr == 0) ? 0.0 : d_mul((double)((nr ? maxr : 0) - (nl ? minl : 4)), 0.25);  // /4: exact scaling
  f[4] = (double)(sl - sr);
  f[5] = (double)(nl - nr);
  f[6] = (double)(nsl - nsr);
  f[7] = threat;
  f[8] = prot;
  int playable = 0, valid = 0;
  double total = 0.0;
#pragma unroll 1
  for (int i = 0; i < L.n_hand && i < 4; i++) {
    const DCard& c = WCARD(wg, L.hand[i].card);
    if (c.obs_id == -32768) err = SB_ERR_OBS_ID;
    if (c.obs_id == -1 || c.obs_id == 32767) continue;
    const int cost = L.hand[i].cost;
    int str = c.kind == KIND_SPELL ? 0 : w_card_strength_of(wg, L.hand[i]);
    if (str == -1) str = 0;
    valid++;
    if (cost > 0) {
      total = d_add(total, d_div((double)str, (double)cost));
      if ((double)cost <= m) playable++;
    }
  }
  if (valid == 0) f[9] = 0.0;
  else {
    // valid is 1..4: divisions by 1, 2 and 4 are exact scalings, only /3 needs the divider
    double playability, avg;
    if (valid == 3) { playability = d_div((double)playable, 3.0); avg = d_div(total, 3.0); }
    else { const double inv = valid == 1 ? 1.0 : valid == 2 ? 0.5 : 0.25; playability = d_mul((double)playable, inv); avg = d_mul(total, inv); }
    f[9] = d_mul(d_add(playability, w_clip01(d_div(avg, 3.0))), 0.5);  // /2: exact scaling
  }
  if (check_ids) {
#pragma unroll 1
    for (int i = 0; i < L.n_deck; i++) if (WCARD(wg, L.deck[i].card).obs_id == -32768) err = SB_ERR_OBS_ID;
#pragma unroll 1
    for (int i = 0; i < wg->hist_n; i++) if (WCARD(wg, wg->hist_card[i]).obs_id == -32768) err = SB_ERR_OBS_ID;
  }
  return err;
}
SBW_FI double w_score_delta(const double* w, const double* fc, const double* fn) {  // evo/heuristic_agent.py:23-51,82-122
  double d = 0.0;
#pragma unroll
  // np.dot at n = 10 is OpenBLAS's scalar tail loop with FMA contraction: sequential fused multiply-add
  for (int i = 0; i < SB_N_FEATURES; i++) d = d_fma(w[i], d_sub(fn[i], fc[i]), d);
  const double eff = d_sub(fn[0], fc[0]);
  const double rp = eff < -0.3 ? d_mul(eff < 0.0 ? -eff : eff, 0.2) : 0.0;
  return d_sub(d_sub(-d, d), rp);
}

// The ten StateFeatures STRAIGHT FROM THE PACKED RECORD (streaming form of w_unpack + w_features for sb_features): the same
// arithmetic in the same order (ascending tiles for the two FP64 sums), reading the 8-byte tile records of the staging image.
SBW_FI int w_features_packed(const SbState* s, const DCard* cards, double* f) {
  W_SHARED(cards);
  const int lo = s->local_order;
  const SbPlayer& L = s->pl[lo];
  const SbPlayer& R = s->pl[1 - lo];
  const double m = L.mana != -1 ? (double)L.mana : 0.0;
  const double hl = L.base != -1 ? (double)L.base : 20.0;
  const double hr = R.base != -1 ? (double)R.base : 20.0;
  double est = d_add(m, 2.0);
  if (est < 3.0) est = 3.0;
  if (est > 10.0) est = 10.0;
  f[0] = w_clip01(d_sub(1.0, d_div(m, est)));
  f[1] = d_sub(hl, hr);
  int sl = 0, sr = 0, nl = 0, nr = 0, nsl = 0, nsr = 0, minl = 99, maxr = -1;
  double threat = 0.0, prot = 0.0;
  u32 occ = w_ballot([&](int l) -> bool { return l < SB_N_TILES && s->tile[l].card != 0; });
#pragma unroll 1
  while (occ) {
    const int t = w_ffs(occ) - 1;
    occ &= occ - 1;
    const SbTile tl = s->tile[t];
    const int y = t >> 2, str = tl.strength;
    const bool counted = str != -1;
    const bool is_struct = (tl.flags & SB_TF_STRUCTURE) != 0;
    if (((tl.flags & SB_TF_OWNER) ? 1 : 0) == lo) {
      if (!is_struct) { nl++; if (y < minl) minl = y; } else nsl++;
      if (counted) { sl += str; prot = d_add(prot, d_mul((double)str, w_fifths(5 - y))); }
    } else {
      if (!is_struct) {
        nr++; if (y > maxr) maxr = y;
        if (counted) threat = d_add(threat, d_mul((double)str, w_fifths(y + 1)));
      } else nsr++;
      if (counted) sr += str;
    }
  }
  const int tot = sl + sr;
  f[2] = tot == 0 ? 0.0 : d_div((double)(sl - sr), (double)tot);
  f[3] = (nl == 0 && nr == 0) ? 0.0 : d_mul((double)((nr ? maxr : 0) - (nl ? minl : 4)), 0.25);
  f[4] = (double)(sl - sr);
  f[5] = (double)(nl - nr);
  f[6] = (double)(nsl - nsr);
  f[7] = threat;
  f[8] = prot;
  const int nh = L.n_hand < SB_HAND_MAX ? L.n_hand : SB_HAND_MAX, nd = L.n_deck < SB_DECK_MAX ? L.n_deck : SB_DECK_MAX;
  const int hn = s->hist_n < 4 ? s->hist_n : 4;
  int playable = 0, valid = 0;
  double total = 0.0;
#pragma unroll 1
  for (int i = 0; i < nh; i++) {
    const int card = L.hand_card[i];
    const DCard& c = cards[card];
    if (c.obs_id == -1 || c.obs_id == 32767) continue;
    const int cost = L.hand_cost[i];
    int str = c.kind == KIND_SPELL ? 0 : w_packed_card_strength(s, cards, lo, 0, i, card, L.hand_flags[i]);
    if (str == -1) str = 0;
    valid++;
    if (cost > 0) {
      total = d_add(total, d_div((double)str, (double)cost));
      if ((double)cost <= m) playable++;
    }
  }
  if (valid == 0) f[9] = 0.0;
  else {
    double playability, avg;
    if (valid == 3) { playability = d_div((double)playable, 3.0); avg = d_div(total, 3.0); }
    else { const double inv = valid == 1 ? 1.0 : valid == 2 ? 0.5 : 0.25; playability = d_mul((double)playable, inv); avg = d_mul(total, inv); }
    f[9] = d_mul(d_add(playability, w_clip01(d_div(avg, 3.0))), 0.5);
  }
  const u32 bad = w_ballot([&](int l) -> bool {
    bool b = false;
    if (l < SB_N_TILES) { const int c = s->tile[l].card; b = c && cards[c].obs_id == -32768; }
    if (l < nd) b = b || cards[L.deck_card[l]].obs_id == -32768;
    if (l < nh) b = b || cards[L.hand_card[l]].obs_id == -32768;
    if (l < hn) b = b || cards[s->hist_card[l]].obs_id == -32768;
    return b;
  });
  return bad ? SB_ERR_OBS_ID : 0;
}

// ---------------------------------------------------------------- observation (games/stormbound.py:400-526)
// obs: 540 int32 in global (GPU) or host memory.  Uniform stores: every lane writes the same value to the same address.
#define WOBSI(l, r, c) obs[((l) * 5 + (r)) * 4 + (c)]
SBW_FI void w_obs_card_row(const WG* wg, int* obs, int layer, int row, const WCard& c, int& err) {
  const DCard& d = WCARD(wg, c.card);
  if (d.obs_id == -32768) err = SB_ERR_OBS_ID;
  WOBSI(layer, row, 0) = d.obs_id;
  WOBSI(layer, row, 1) = c.cost;
  WOBSI(layer, row, 2) = d.kind == KIND_SPELL ? -1 : w_card_strength_of(wg, c);
  WOBSI(layer, row, 3) = d.kind == KIND_UNIT ? d.movement : -1;
}
SBW_NI int w_observe(WG* wg, int* obs) {
  W_SHARED(wg);
  int err = 0;
  FOR_LANES(l) { for (int i = l; i < SB_OBS_INTS; i += 32) obs[i] = -1; } END_LANES
  const int lo = wg->local_order;
#pragma unroll 1
  for (int t = 0; t < SB_N_TILES; t++) {
    const int id = wg->board[t];
    if (id < 0) continue;
    const DCard& d = WCARD(wg, wg->e_card[id]);
    if (d.obs_id == -32768) err = SB_ERR_OBS_ID;
    const int base = w_owner(wg, id) == lo ? 0 : 16, y = t >> 2, x = t & 3;
    if (!w_is_struct(wg, id)) {
      WOBSI(base + 0, y, x) = d.obs_id;
      WOBSI(base + 1, y, x) = wg->e_str[id];
      WOBSI(base + 2, y, x) = d.movement;
      WOBSI(base + 3, y, x) = (w_st(wg, id, SB_ST_VITALIZED) ? 1 : 0) | (w_st(wg, id, SB_ST_POISONED) ? 2 : 0) | (w_st(wg, id, SB_ST_CONFUSED) ? 4 : 0) |
                              (w_st(wg, id, SB_ST_FROZEN) ? 8 : 0) | (w_st(wg, id, SB_ST_DISABLED) ? 16 : 0);
    } else {
      WOBSI(base + 4, y, x) = d.obs_id;
      WOBSI(base + 5, y, x) = wg->e_str[id];
    }
  }
  const WPly& L = wg->pl[lo];
  const WPly& R = wg->pl[1 - lo];
#pragma unroll 1
  for (int i = 0; i < L.n_hand && i < 4; i++) w_obs_card_row(wg, obs, 6, i, L.hand[i], err);
#pragma unroll 1
  for (int c = 0; c < 4; c++) WOBSI(6, 4, c) = 32767;
  i8* idx = wg->scr;
  const int nd = L.n_deck < 24 ? L.n_deck : 24;
#pragma unroll 1
  for (int i = 0; i < nd; i++) idx[i] = (i8)i;
#pragma unroll 1
  for (int i = 1; i < nd; i++) {  // sorted(deck, key=(cost, card_id)), stable
    const i8 v = idx[i];
    int j = i - 1;
#pragma unroll 1
    while (j >= 0 && (L.deck[idx[j]].cost > L.deck[v].cost ||
                      (L.deck[idx[j]].cost == L.deck[v].cost && L.deck[idx[j]].card > L.deck[v].card))) { idx[j + 1] = idx[j]; j--; }
    idx[j + 1] = v;
  }
#pragma unroll 1
  for (int layer = 0; layer < 6; layer++) {
#pragma unroll 1
    for (int k = 0; k < 4; k++) { const int d = layer * 4 + k; if (d < nd) w_obs_card_row(wg, obs, 7 + layer, k, L.deck[idx[d]], err); }
#pragma unroll 1
    for (int c = 0; c < 4; c++) WOBSI(7 + layer, 4, c) = 32768;
  }
#pragma unroll 1
  for (int r = 0; r < 5; r++) for (int c = 0; c < 4; c++) {
    WOBSI(13, r, c) = L.mana; WOBSI(14, r, c) = L.base; WOBSI(15, r, c) = L.faction;
    WOBSI(22, r, c) = R.mana; WOBSI(23, r, c) = R.base; WOBSI(24, r, c) = R.faction;
    WOBSI(25, r, c) = wg->player_sign * 99999;
  }
#pragma unroll 1
  for (int i = 0; i < 4; i++) {
    const int h = i - (4 - wg->hist_n);
    if (h >= 0) {
      WOBSI(26, i, 0) = wg->hist_owner[h] ? -99999 : 99999;
      WOBSI(26, i, 1) = WCARD(wg, wg->hist_card[h]).obs_id;
      if (WCARD(wg, wg->hist_card[h]).obs_id == -32768) err = SB_ERR_OBS_ID;
    }
  }
#pragma unroll 1
  for (int c = 0; c < 4; c++) WOBSI(26, 4, c) = 32769;
  return err;
}

// ---------------------------------------------------------------- scripted opponent (games/stormbound.py:563-637)
// Draws its choices from the GAME's stream (self.random), so it advances wg->draw.
SBW_FI bool w_mask_any(const u32* m, int lo, int hi) {  // any legal action in [lo, hi]
#pragma unroll 1
  for (int a = lo; a <= hi; a++) if (m[a >> 5] >> (a & 31) & 1u) return true;
  return false;
}
SBW_FI int w_place_action(int ci, int pt) {  // Action.to_int PLACE (games/stormbound.py:261-270): row 0 is not encodable
  const int y = wpt_y(pt);
  return (y >= 1 && y <= 4) ? 16 * ci + (4 - y) * 4 + wpt_x(pt) : SB_ACTION_PASS;
}
SBW_NI int w_expert_action(WG* wg) {
  W_SHARED(wg);
  const WPly& p = wg->pl[wg->local_order];
  w_legal_mask(wg);
  const u32* m = wg->lm;
  if (w_mask_any(m, 148, 151)) {
    if (p.n_hand == 0) { WERR(wg, SB_ERR_EMPTY_CHOICE); return SB_ACTION_PASS; }  // max([])
    int max_cost = -1000;
    PL sel; sel.v = 0; sel.n = 0;
#pragma unroll 1
    for (int i = 0; i < p.n_hand; i++) if (p.hand[i].cost > max_cost) max_cost = p.hand[i].cost;
    if (max_cost > p.mana) {
#pragma unroll 1
      for (int i = 0; i < p.n_hand; i++) if (p.hand[i].cost == max_cost) pl_push(sel, i);
      return 148 + pl_get(sel, w_rng_below(wg, sel.n));
    }
  }
  PL playable; playable.v = 0; playable.n = 0;
#pragma unroll 1
  for (int i = 0; i < 4; i++) if (w_mask_any(m, 16 * i, 16 * i + 15) || w_mask_any(m, 21 * i + 64, 21 * i + 84)) pl_push(playable, i);
  if (playable.n == 0) return SB_ACTION_PASS;
  bool any_eq = false;
  int min_cost = 1 << 20;
#pragma unroll 1
  for (int k = 0; k < playable.n; k++) { const int c = p.hand[pl_get(playable, k)].cost; any_eq |= c == p.mana; min_cost = c < min_cost ? c : min_cost; }
  const int want = any_eq ? (int)p.mana : min_cost;
  PL sel; sel.v = 0; sel.n = 0;
#pragma unroll 1
  for (int k = 0; k < playable.n; k++) if (p.hand[pl_get(playable, k)].cost == want) pl_push(sel, k);
  const int index = pl_get(playable, pl_get(sel, w_rng_below(wg, sel.n)));
  const DCard& c = WCARD(wg, p.hand[index].card);
  const TL en = w_targets(wg, wg->current_order, w_mkT(TK_UNIT, TS_ENEMY), PT_NONE);
  const int nb = w_popc(en.m & 0xF0000u);
  if (c.kind == KIND_SPELL) {
    if (!(c.flags & DCF_TARGET)) return 64 + 21 * index;
    const TL tg = w_targets(wg, wg->current_order, w_card_target(c), PT_NONE);
    const int nt = tl_n(tg);
    if (nt == 0) { WERR(wg, SB_ERR_EMPTY_CHOICE); return SB_ACTION_PASS; }
    const int where = tl_nth(tg, w_rng_below(wg, nt));
    return where >= 20 ? SB_ACTION_PASS : 65 + 21 * index + (4 - wpt_y(where)) * 4 + wpt_x(where);
  }
  i8* cand = wg->scr;
  int nc = 0;
  TL it = en;
  if (c.kind == KIND_UNIT && nb > 0) {  // an enemy unit stands next to the base: block beside it
#pragma unroll 1
    for (;;) {
      const int pt = tl_pop(it);
      if (pt == PT_NONE) break;
      const int x = wpt_x(pt), y = wpt_y(pt);
      if (y != 4) continue;
      if (x > 0 && w_at_xy(wg, x - 1, y) < 0) cand[nc++] = (i8)(y * 4 + x - 1);
      else if (x < 3 && w_at_xy(wg, x + 1, y) < 0) cand[nc++] = (i8)(y * 4 + x + 1);
    }
  } else {
    const int fl = p.front_line;
#pragma unroll 1
    for (int x = 0; x < 4; x++) if (w_valid_xy(x, fl) && w_at_xy(wg, x, fl) < 0) cand[nc++] = (i8)(fl * 4 + x);
#pragma unroll 1
    for (;;) {
      const int pt = tl_pop(it);
      if (pt == PT_NONE) break;
      const int x = wpt_x(pt), y = wpt_y(pt);
      if (x > 0 && y >= fl && w_at_xy(wg, x - 1, y) < 0) cand[nc++] = (i8)(y * 4 + x - 1);
      else if (x < 3 && y >= fl && w_at_xy(wg, x + 1, y) < 0) cand[nc++] = (i8)(y * 4 + x + 1);
      else if (y < 4 && y + 1 >= fl && w_at_xy(wg, x, y + 1) < 0) cand[nc++] = (i8)((y + 1) * 4 + x);
    }
  }
  return nc > 0 ? w_place_action(index, cand[w_rng_below(wg, nc)]) : SB_ACTION_PASS;
}
