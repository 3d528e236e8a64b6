// sbw_agent.cuh -- HeuristicAgent.select_action on the warp-per-game engine (evo/heuristic_agent.py:23-80,
// evo/game_adapter.py:280-337): every legal action is applied to a fork of the game, the fork's ten features are scored
// against the current ones with the agent's weights, the first maximum wins.
//
// The game (`base`) and the fork (`work`) are two working sets in shared memory; a fork is a lane-parallel copy of the
// live part (w_copy_game: five 64-bit moves per lane), never a trip through local or global memory.
#pragma once
#include "sbw_io.cuh"

SBW_FI void w_copy_game(WG* dst, const WG* src) {
  const u64* s8 = reinterpret_cast<const u64*>(src);
  u64* d8 = reinterpret_cast<u64*>(dst);
  const int nw = (int)(offsetof(WG, lm) / 8);
  static_assert(offsetof(WG, lm) % 8 == 0, "w_copy_game layout");
  FOR_LANES(l) {
#pragma unroll 1
    for (int i = l; i < nw; i += 32) d8[i] = s8[i];
  } END_LANES
}
// One decision.  w: the mover's ten weights (shared or host memory).  scores_out (optional): f64[156], NaN-initialised by
// the caller, receives the score of every legal action.  commit: `base` becomes the post-action state.
SBW_NI int w_decide(WG* base, WG* work, const double* w, double* scores_out, bool commit) {
  W_SHARED(base);
  W_SHARED(work);
  const int n_legal = w_legal_mask(base);
  u32 m[SB_MASK_WORDS];
#pragma unroll
  for (int i = 0; i < SB_MASK_WORDS; i++) m[i] = base->lm[i];
  int best_action = -1;
  double best_score = 0.0;
  if (n_legal == 1 && !scores_out) best_action = w_nth_action(m, 0);  // forced move: argmax of one
  else {
    const int cur_err = w_features(base, base->feat);
#pragma unroll 1
    for (int wd = 0; wd < SB_MASK_WORDS; wd++) {
      u32 bits = m[wd];
#pragma unroll 1
      while (bits) {
        const int a = wd * 32 + w_ffs(bits) - 1;
        bits &= bits - 1;
        w_copy_game(work, base);
        w_game_step(work, a);
        w_end_of_step(work);  // a candidate that leaves more than the packed layout holds is an engine status, like for the committed state
        double sc = 0.0;
        int nerr = work->err;
        if (!nerr) nerr = w_features(work, work->feat);
        if (!nerr && !cur_err) sc = w_score_delta(w, base->feat, work->feat);
        if (scores_out) scores_out[a] = sc;
        if (best_action < 0 || sc > best_score) { best_score = sc; best_action = a; }  // np.argmax: first maximum
      }
    }
  }
  const int action = best_action < 0 ? SB_ACTION_PASS : best_action;
  if (commit) {
    w_game_step(base, action);
    w_end_of_step(base);
  }
  return action;
}
