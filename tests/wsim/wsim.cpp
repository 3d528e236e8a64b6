// wsim.cpp -- HOST build of the warp-per-game engine (monsoon_b200/csrc/sbw_*.cuh) for the tests.
//
// The engine is single-source: on the GPU the 32 lanes of a warp walk one game together; here the same code is compiled
// with FOR_LANES as a loop (sbw_warp.cuh).  This library is TEST INFRASTRUCTURE: it lets the CPU suite compare the warp
// engine's rules with the oracle step by step without a GPU.  Nothing in the product path loads it.
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#include "../../monsoon_b200/csrc/sb_card_table_host.h"
#include "../../monsoon_b200/csrc/sbw_agent.cuh"

static DCard g_cards[SBC_COUNT];
static double g_wt[WT_N];
static int g_init = 0;
static void init_once() {
  if (g_init) return;
  sb_build_dcards(g_cards);
  sb_build_weights(g_wt);
  g_init = 1;
}
static void wg_init(WG* wg) { memset(wg, 0, sizeof(WG)); wg->cards = g_cards; wg->wt = g_wt; }

extern "C" {
int wsim_sizeof_wg(void) { return (int)sizeof(WG); }
int wsim_live_bytes(void) { return (int)offsetof(WG, lm); }
void wsim_legal_mask(const uint8_t* state, uint32_t* mask) {
  init_once();
  WG wg; wg_init(&wg);
  w_unpack(&wg, (const SbState*)state);
  w_legal_mask(&wg);
  for (int i = 0; i < SB_MASK_WORDS; i++) mask[i] = wg.lm[i];
}
void wsim_legal_mask_packed(const uint8_t* state, uint32_t* mask) {  // the streaming form: no unpack
  init_once();
  w_legal_mask_packed((const SbState*)state, g_cards, mask);
}
void wsim_step(uint8_t* state, int action) {
  init_once();
  WG wg; wg_init(&wg);
  w_unpack(&wg, (const SbState*)state);
  w_game_step(&wg, action);
  w_pack(&wg, (SbState*)state);
}
// whole uniform-random game like k_rollout_random: returns env steps; digests[k] = FNV-1a of the packed state after step k
int wsim_rollout_random(uint8_t* state, int max_steps, uint64_t* digests, uint8_t* actions) {
  init_once();
  WG wg; wg_init(&wg);
  SbState* s = (SbState*)state;
  w_unpack(&wg, s);
  int k = 0;
  bool alive = !(wg.done & SB_DONE) && !wg.err && max_steps > 0;
  while (alive) {
    const int a = w_pick_action(&wg);
    w_game_step(&wg, a);
    w_end_of_step(&wg);
    if (actions) actions[k] = (uint8_t)a;
    if (digests) { SbState t; w_pack(&wg, &t); digests[k] = w_digest_state(&t); }
    k++;
    alive = !(wg.done & SB_DONE) && !wg.err && k < max_steps;
  }
  w_pack(&wg, s);
  return k;
}
int wsim_features(const uint8_t* state, double* f) {
  init_once();
  WG wg; wg_init(&wg);
  w_unpack(&wg, (const SbState*)state);
  w_scan_badobs(&wg);
  return w_features(&wg, f);
}
int wsim_features_packed(const uint8_t* state, double* f) {  // the streaming form: no unpack
  init_once();
  return w_features_packed((const SbState*)state, g_cards, f);
}
int wsim_observe(const uint8_t* state, int32_t* obs) {
  init_once();
  WG wg; wg_init(&wg);
  w_unpack(&wg, (const SbState*)state);
  return w_observe(&wg, obs);
}
int wsim_observe_packed(const uint8_t* state, int32_t* obs) {  // the streaming form: no unpack
  init_once();
  return w_observe_packed((const SbState*)state, g_cards, obs);
}
int wsim_expert_action(uint8_t* state) {
  init_once();
  WG wg; wg_init(&wg);
  SbState* s = (SbState*)state;
  w_unpack(&wg, s);
  const int a = w_expert_action(&wg);
  s->draw = wg.draw;
  s->err = wg.err;
  return a;
}
int wsim_select_action(const uint8_t* state, const double* w, double* scores) {
  init_once();
  static WG base, work;
  wg_init(&base); wg_init(&work);
  w_unpack(&base, (const SbState*)state);
  w_scan_badobs(&base);
  return w_decide(&base, &work, w, scores, false);
}
// whole game like k_rollout_heuristic; a NULL weight vector hands that seat to the scripted opponent.
// returns the result (0 FIRST wins, 1 SECOND wins, -1 draw / step limit, -2 engine status); *steps_out = env steps
int wsim_play_heuristic(uint8_t* state, const double* w_first, const double* w_second, int max_steps, int* steps_out, uint8_t* actions) {
  init_once();
  static WG base, work;
  wg_init(&base); wg_init(&work);
  SbState* s = (SbState*)state;
  w_unpack(&base, s);
  w_scan_badobs(&base);
  int k = 0, res = -1;
  for (;;) {
    if (k >= max_steps || base.pl[0].base < 0 || base.pl[1].base < 0) break;
    const bool first_to_move = base.player_sign == 1;
    const double* w = first_to_move ? w_first : w_second;
    int a;
    if (w) a = w_decide(&base, &work, w, nullptr, true);
    else { a = w_expert_action(&base); w_game_step(&base, a); w_end_of_step(&base); }
    if (actions) actions[k] = (uint8_t)a;
    k++;
    if (base.err) { res = -2; break; }
  }
  if (res != -2) {
    const bool l0 = base.pl[0].base < 0, l1 = base.pl[1].base < 0;
    res = (l1 && !l0) ? 0 : (l0 && !l1) ? 1 : -1;
  }
  w_pack(&base, s);
  if (steps_out) *steps_out = k;
  return res;
}
}  // extern "C"
