// sbw_kernels.cu -- sm_100a kernels of the warp-per-game engine (sbw_*.cuh) behind the C ABI of sb_kernels.cu.
//
// Every kernel maps ONE GAME TO ONE WARP.  A warp owns a slot of dynamic shared memory: the working set WG (2.2 KB,
// structure-of-arrays entity pool + hands / decks) and a 512-byte staging image of the packed record.  Records move between
// HBM and the staging image as one coalesced 128-bit access per lane (32 lanes x 16 B = one 512-byte record), the rules run
// on shared memory with warp-uniform control flow (ballots / REDUX / shuffles for the scans), nothing lives in local memory
// but call frames.  The card table (130 x 24 B) is staged once per CTA.
//   kw_rollout_random     whole uniform-random game per launch; persistent grid, a warp takes the next game from a counter
//   kw_step               one env step: load, unpack, step, pack, store, fused next legal mask
//   kw_legal_mask / kw_observe / kw_features / kw_expert_action
//   kw_select_action      one HeuristicAgent decision (fork = shared-memory copy, sbw_agent.cuh)
//   kw_rollout_heuristic  whole heuristic-agent game per launch, fitness result in the epilogue
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include "sbw_agent.cuh"
#include "sbw_launch.h"

#define W_SLOT_BYTES ((int)((sizeof(WG) + sizeof(SbState) + 15) & ~15u))

__device__ __forceinline__ void kw_stage_cards(DCard* s_cards, const DCard* cards) {
  const u32* src = reinterpret_cast<const u32*>(cards);
  u32* dst = reinterpret_cast<u32*>(s_cards);
  for (int i = threadIdx.x; i < (int)(SBC_COUNT * sizeof(DCard) / 4); i += blockDim.x) dst[i] = __ldg(src + i);
  __syncthreads();
}
__device__ __forceinline__ WG* kw_slot(unsigned char* dyn, int warp, int slot_bytes) { return reinterpret_cast<WG*>(dyn + (size_t)warp * slot_bytes); }
__device__ __forceinline__ SbState* kw_image(WG* wg) { return reinterpret_cast<SbState*>(reinterpret_cast<unsigned char*>(wg) + sizeof(WG)); }
// one packed record <-> the staging image: 32 lanes x 128 bits, fully coalesced
__device__ __forceinline__ void kw_load_record(SbState* img, const u8* states, int gi) {
  const int lane = threadIdx.x & 31;
  reinterpret_cast<uint4*>(img)[lane] = __ldg(reinterpret_cast<const uint4*>(states + (size_t)gi * SB_STATE_BYTES) + lane);
  __syncwarp();
}
__device__ __forceinline__ void kw_store_record(u8* states, int gi, const SbState* img) {
  const int lane = threadIdx.x & 31;
  __syncwarp();
  reinterpret_cast<uint4*>(states + (size_t)gi * SB_STATE_BYTES)[lane] = reinterpret_cast<const uint4*>(img)[lane];
}
__device__ __forceinline__ int kw_next_game(int* queue, int fallback) {  // persistent grid: next game index from a global counter
  if (!queue) return fallback;
  int gi = 0;
  if ((threadIdx.x & 31) == 0) gi = atomicAdd(queue, 1);
  return __shfl_sync(0xFFFFFFFFu, gi, 0);
}

// SYNC: turn-synchronous CTA schedule.  A game alternates "a few non-PASS actions, then PASS"; with SYNC all warps of the
// CTA play their non-PASS actions round by round (CTA-wide vote per round) and then their PASS together, so the warps of
// an SM walk the same functions at the same time and share the instruction-cache lines they fetch (the unsynchronised
// kernel spends 70-85 % of its stall samples waiting for instructions: profiles/r2_summary.md).
template <int WPC, int MINB, int SYNC>
__global__ void __launch_bounds__(WPC * 32, MINB) kw_rollout_random(int n, u8* states, int max_steps, int* steps_out, unsigned long long* chain,
                                                                     const DCard* cards, const double* wt, int* queue) {
  __shared__ DCard s_cards[SBC_COUNT];
  extern __shared__ __align__(16) unsigned char s_dyn[];
  kw_stage_cards(s_cards, cards);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WG* wg = kw_slot(s_dyn, warp, W_SLOT_BYTES);
  SbState* img = kw_image(wg);
  wg->cards = s_cards; wg->wt = wt;
  const bool digest = chain != nullptr;
  int gi = blockIdx.x * WPC + warp;
  if (SYNC == 0) {
#pragma unroll 1
    for (;;) {
      gi = kw_next_game(queue, gi);
      if (gi >= n) break;
      kw_load_record(img, states, gi);
      w_unpack(wg, img);
      unsigned long long ch = digest ? chain[gi] : 0ull;
      int k = 0;
      bool alive = !(wg->done & SB_DONE) && !wg->err && max_steps > 0;
#pragma unroll 1
      while (alive) {
        w_game_step(wg, w_pick_action(wg));
        w_end_of_step(wg);
        if (digest) { w_pack(wg, img); __syncwarp(); ch = (ch ^ w_digest_state(img)) * 0x100000001B3ull; }
        k++;
        alive = !(wg->done & SB_DONE) && !wg->err && k < max_steps;
      }
      w_pack(wg, img);
      kw_store_record(states, gi, img);
      if (lane == 0) {
        if (steps_out) steps_out[gi] = k;
        if (digest) chain[gi] = ch;
      }
      if (!queue) break;
      __syncwarp();
    }
  } else {
    bool has_game = false, alive = false, at_pass = false, dry = false;
    unsigned long long ch = 0ull;
    int k = 0;
    bool first = true;
#pragma unroll 1
    for (;;) {
      if (!alive && !dry) {  // retire the finished game, take the next one
        if (has_game) {
          w_pack(wg, img);
          kw_store_record(states, gi, img);
          if (lane == 0) { if (steps_out) steps_out[gi] = k; if (digest) chain[gi] = ch; }
          __syncwarp();
          has_game = false;
        }
        if (queue) gi = kw_next_game(queue, gi); else if (!first) gi = n;
        first = false;
        if (gi < n) {
          has_game = true;
          kw_load_record(img, states, gi);
          w_unpack(wg, img);
          ch = digest ? chain[gi] : 0ull;
          k = 0; at_pass = false;
          alive = !(wg->done & SB_DONE) && !wg->err && max_steps > 0;
        } else dry = true;
      }
      if (!__syncthreads_or(alive)) break;
#pragma unroll 1
      for (;;) {  // phase A: non-PASS actions, one round at a time
        int a = -1;
        if (alive && !at_pass) {
          a = w_pick_action(wg);
          if (a == SB_ACTION_PASS) { at_pass = true; a = -1; }
        }
        if (SYNC == 2) {  // two barriers per turn: every warp plays ALL its non-PASS actions of the turn, then everybody's PASS
          if (a < 0) { __syncthreads(); break; }
        } else if (!__syncthreads_or(a >= 0)) break;
        if (a >= 0) {
          w_game_step(wg, a);
          w_end_of_step(wg);
          if (digest) { w_pack(wg, img); __syncwarp(); ch = (ch ^ w_digest_state(img)) * 0x100000001B3ull; }
          k++;
          alive = !(wg->done & SB_DONE) && !wg->err && k < max_steps;
        }
      }
      if (alive && at_pass) {  // phase B: everybody's PASS
        w_game_step(wg, SB_ACTION_PASS);
        w_end_of_step(wg);
        if (digest) { w_pack(wg, img); __syncwarp(); ch = (ch ^ w_digest_state(img)) * 0x100000001B3ull; }
        k++;
        at_pass = false;
        alive = !(wg->done & SB_DONE) && !wg->err && k < max_steps;
      }
    }
    if (has_game) {
      w_pack(wg, img);
      kw_store_record(states, gi, img);
      if (lane == 0) { if (steps_out) steps_out[gi] = k; if (digest) chain[gi] = ch; }
    }
  }
}

template <int WPC>
__global__ void __launch_bounds__(WPC * 32) kw_step(int n, u8* states, const u8* actions, i8* reward, u8* done, u8* err, u32* next_masks,
                                                    const DCard* cards, const double* wt) {
  __shared__ DCard s_cards[SBC_COUNT];
  extern __shared__ __align__(16) unsigned char s_dyn[];
  kw_stage_cards(s_cards, cards);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WG* wg = kw_slot(s_dyn, warp, W_SLOT_BYTES);
  SbState* img = kw_image(wg);
  wg->cards = s_cards; wg->wt = wt;
#pragma unroll 1
  for (int gi = blockIdx.x * WPC + warp; gi < n; gi += gridDim.x * WPC) {
    kw_load_record(img, states, gi);
    w_unpack(wg, img);
    w_game_step(wg, actions[gi]);
    w_pack(wg, img);
    kw_store_record(states, gi, img);
    if (lane == 0) {
      if (reward) reward[gi] = (wg->done & SB_REWARD) ? 1 : 0;
      if (done) done[gi] = (wg->done & SB_DONE) ? 1 : 0;
      if (err) err[gi] = img->err;
    }
    if (next_masks) {
      w_legal_mask(wg);
      if (lane < SB_MASK_WORDS) next_masks[(size_t)gi * SB_MASK_WORDS + lane] = wg->lm[lane];
    }
    __syncwarp();
  }
}

// The record-in / result-out queries as STREAMING kernels: legal mask (MODE 0), observation (1) and features (2) are computed
// straight from the packed record (w_*_packed, one lane per tile / card), 250-450 warp instructions per record instead of the
// 600-1,400 of unpack + query, so that the kernels run on the record stream instead of the issue slots.  Every warp keeps the
// NEXT record's 512 bytes in flight (one 128-bit load per lane, issued before the current record is processed); the
// observation is assembled in shared memory and leaves as 135 coalesced 128-bit stores.  Persistent grid, 2,048 threads per SM.
// Algorithmic bytes per record: 512 in + 20 / 2,160 / 80 out.
template <int WPC, int MODE>
__global__ void __launch_bounds__(WPC * 32, 2048 / (WPC * 32)) kw_stream(int n, const u8* states, void* out, u8* err, const DCard* cards) {
  __shared__ DCard s_cards[SBC_COUNT];
  __shared__ __align__(16) SbState s_img[WPC];
  __shared__ __align__(16) int s_out[MODE == 1 ? WPC * SB_OBS_INTS : (MODE == 2 ? WPC * 2 * SB_N_FEATURES : 4)];
  kw_stage_cards(s_cards, cards);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int stride = gridDim.x * WPC;
  int gi = blockIdx.x * WPC + warp;
  uint4 nxt = make_uint4(0, 0, 0, 0);
  if (gi < n) nxt = __ldg(reinterpret_cast<const uint4*>(states + (size_t)gi * SB_STATE_BYTES) + lane);
#pragma unroll 1
  for (; gi < n; gi += stride) {
    const uint4 cur = nxt;
    if (gi + stride < n) nxt = __ldg(reinterpret_cast<const uint4*>(states + (size_t)(gi + stride) * SB_STATE_BYTES) + lane);
    reinterpret_cast<uint4*>(&s_img[warp])[lane] = cur;
    __syncwarp();
    if (MODE == 0) {
      u32 m[SB_MASK_WORDS];
      w_legal_mask_packed(&s_img[warp], s_cards, m);
      const u32 v = lane == 0 ? m[0] : lane == 1 ? m[1] : lane == 2 ? m[2] : lane == 3 ? m[3] : m[4];
      if (lane < SB_MASK_WORDS) reinterpret_cast<u32*>(out)[(size_t)gi * SB_MASK_WORDS + lane] = v;
    } else if (MODE == 1) {
      int* ob = s_out + warp * SB_OBS_INTS;
      const int e = w_observe_packed(&s_img[warp], s_cards, ob);
      uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<int*>(out) + (size_t)gi * SB_OBS_INTS);
      const uint4* src = reinterpret_cast<const uint4*>(ob);
#pragma unroll
      for (int i = lane; i < SB_OBS_INTS / 4; i += 32) dst[i] = src[i];
      if (err && lane == 0) err[gi] = (u8)e;
    } else {
      double* f = reinterpret_cast<double*>(s_out) + warp * SB_N_FEATURES;
      const int e = w_features_packed(&s_img[warp], s_cards, f);
      __syncwarp();
      if (lane < SB_N_FEATURES) reinterpret_cast<double*>(out)[(size_t)gi * SB_N_FEATURES + lane] = f[lane];
      if (err && lane == 0) err[gi] = (u8)e;
    }
    __syncwarp();
  }
}

// mode 0 legal mask, 1 observation, 2 features, 3 expert action
template <int WPC, int MODE>
__global__ void __launch_bounds__(WPC * 32) kw_query(int n, u8* states, void* out, u8* err, const DCard* cards, const double* wt) {
  __shared__ DCard s_cards[SBC_COUNT];
  extern __shared__ __align__(16) unsigned char s_dyn[];
  kw_stage_cards(s_cards, cards);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WG* wg = kw_slot(s_dyn, warp, W_SLOT_BYTES);
  SbState* img = kw_image(wg);
  wg->cards = s_cards; wg->wt = wt;
#pragma unroll 1
  for (int gi = blockIdx.x * WPC + warp; gi < n; gi += gridDim.x * WPC) {
    kw_load_record(img, states, gi);
    w_unpack(wg, img);
    if (MODE == 0) {
      w_legal_mask(wg);
      if (lane < SB_MASK_WORDS) reinterpret_cast<u32*>(out)[(size_t)gi * SB_MASK_WORDS + lane] = wg->lm[lane];
    } else if (MODE == 1) {
      const int e = w_observe(wg, reinterpret_cast<int*>(out) + (size_t)gi * SB_OBS_INTS);
      if (err && lane == 0) err[gi] = (u8)e;
    } else if (MODE == 2) {
      w_scan_badobs(wg);
      const int e = w_features(wg, wg->feat);
      if (lane < SB_N_FEATURES) reinterpret_cast<double*>(out)[(size_t)gi * SB_N_FEATURES + lane] = wg->feat[lane];
      if (err && lane == 0) err[gi] = (u8)e;
    } else {
      const int a = w_expert_action(wg);
      if (lane == 0) {  // only the stream position (and a possible error code) change in the record
        reinterpret_cast<u8*>(out)[gi] = (u8)a;
        SbState* rec = reinterpret_cast<SbState*>(states + (size_t)gi * SB_STATE_BYTES);
        rec->draw = wg->draw;
        rec->err = wg->err;
      }
    }
    __syncwarp();
  }
}

// a warp's slot for the heuristic agent: the game, the fork, the two seats' weights, the staging image
#define W_HSLOT_BYTES ((int)((2 * sizeof(WG) + sizeof(SbState) + 2 * SB_N_FEATURES * sizeof(double) + 15) & ~15u))
__device__ __forceinline__ WG* kw_hfork(WG* base) { return base + 1; }
__device__ __forceinline__ SbState* kw_himage(WG* base) { return reinterpret_cast<SbState*>(base + 2); }
__device__ __forceinline__ double* kw_hweights(WG* base) { return reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(base + 2) + sizeof(SbState)); }

template <int WPC>
__global__ void __launch_bounds__(WPC * 32) kw_select_action(int n, const u8* states, const double* weights, u8* actions, double* scores,
                                                             const DCard* cards, const double* wt) {
  __shared__ DCard s_cards[SBC_COUNT];
  extern __shared__ __align__(16) unsigned char s_dyn[];
  kw_stage_cards(s_cards, cards);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WG* base = kw_slot(s_dyn, warp, W_HSLOT_BYTES);
  WG* work = kw_hfork(base);
  SbState* img = kw_himage(base);
  double* w = kw_hweights(base);
  base->cards = s_cards; base->wt = wt; work->cards = s_cards; work->wt = wt;
#pragma unroll 1
  for (int gi = blockIdx.x * WPC + warp; gi < n; gi += gridDim.x * WPC) {
    kw_load_record(img, states, gi);
    w_unpack(base, img);
    w_scan_badobs(base);
    if (lane < SB_N_FEATURES) w[lane] = weights[(size_t)gi * SB_N_FEATURES + lane];
    if (scores) for (int a = lane; a < SB_N_ACTIONS; a += 32) scores[(size_t)gi * SB_N_ACTIONS + a] = __longlong_as_double(0x7FF8000000000000ll);
    __syncwarp();
    const int a = w_decide(base, work, w, scores ? scores + (size_t)gi * SB_N_ACTIONS : nullptr, false);
    if (lane == 0) actions[gi] = (u8)a;
    __syncwarp();
  }
}

// Whole heuristic-agent games.  A seat without a weight table is played by the scripted opponent.  The fitness result
// (0 FIRST wins, 1 SECOND wins, -1 draw or step limit, -2 engine status) is written in the epilogue; counts (optional,
// i32[P,4]: win / draw / loss / aborted of the FIRST seat's individual) are accumulated with one atomicAdd per game.
template <int WPC, int MINB>
__global__ void __launch_bounds__(WPC * 32, MINB) kw_rollout_heuristic(int n, u8* states, const double* w_first, const double* w_second, const int* idx_first,
                                                                        const int* idx_second, int max_steps, i8* result, int* steps_out,
                                                                        const DCard* cards, const double* wt, int* queue) {
  __shared__ DCard s_cards[SBC_COUNT];
  extern __shared__ __align__(16) unsigned char s_dyn[];
  kw_stage_cards(s_cards, cards);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WG* base = kw_slot(s_dyn, warp, W_HSLOT_BYTES);
  WG* work = kw_hfork(base);
  SbState* img = kw_himage(base);
  double* wf = kw_hweights(base);
  double* ws = wf + SB_N_FEATURES;
  base->cards = s_cards; base->wt = wt; work->cards = s_cards; work->wt = wt;
  int gi = blockIdx.x * WPC + warp;
#pragma unroll 1
  for (;;) {
    gi = kw_next_game(queue, gi);
    if (gi >= n) break;
    kw_load_record(img, states, gi);
    w_unpack(base, img);
    w_scan_badobs(base);
    if (lane < SB_N_FEATURES) {
      if (w_first) wf[lane] = w_first[(size_t)(idx_first ? idx_first[gi] : gi) * SB_N_FEATURES + lane];
      if (w_second) ws[lane] = w_second[(size_t)(idx_second ? idx_second[gi] : gi) * SB_N_FEATURES + lane];
    }
    __syncwarp();
    int k = 0, res = -1;
#pragma unroll 1
    for (;;) {
      if (k >= max_steps || base->pl[0].base < 0 || base->pl[1].base < 0) break;
      const bool first_to_move = base->player_sign == 1;
      if (first_to_move ? w_first != nullptr : w_second != nullptr) w_decide(base, work, first_to_move ? wf : ws, nullptr, true);
      else {  // expert_action draws from the game's own stream, then the action is stepped (games/stormbound.py:563-637)
        const int a = w_expert_action(base);
        w_game_step(base, a);
        w_end_of_step(base);
      }
      k++;
      if (base->err) { res = -2; break; }
    }
    if (res != -2) {
      const bool l0 = base->pl[0].base < 0, l1 = base->pl[1].base < 0;
      res = (l1 && !l0) ? 0 : (l0 && !l1) ? 1 : -1;
    }
    w_pack(base, img);
    kw_store_record(states, gi, img);
    if (lane == 0) {
      if (result) result[gi] = (i8)res;
      if (steps_out) steps_out[gi] = k;
    }
    if (!queue) break;
    __syncwarp();
  }
}

// ================================================================ launchers (called by the C ABI in sb_kernels.cu)
static inline int grid_for(int n, int per) { return (n + per - 1) / per; }
static int g_carveout = -1;  // percent of the unified L1 / shared memory given to shared memory; -1 = driver default
template <class K> static cudaError_t set_smem(K kernel, int bytes) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess || g_carveout < 0) return e;
  return cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, g_carveout);
}
#define W_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return e_; } while (0)

int sbw_wg_bytes(void) { return (int)sizeof(WG); }

cudaError_t sbw_init(void) {
  { const char* e = getenv("SBW_CARVEOUT"); if (e) g_carveout = atoi(e); }
  W_TRY((set_smem(kw_rollout_random<4, 8, 0>, 4 * W_SLOT_BYTES)));
  W_TRY((set_smem(kw_rollout_random<8, 4, 0>, 8 * W_SLOT_BYTES)));
  W_TRY((set_smem(kw_rollout_random<8, 8, 0>, 8 * W_SLOT_BYTES)));
  W_TRY((set_smem(kw_rollout_random<8, 4, 1>, 8 * W_SLOT_BYTES)));
  W_TRY((set_smem(kw_rollout_random<16, 2, 1>, 16 * W_SLOT_BYTES)));
  W_TRY((set_smem(kw_rollout_random<32, 1, 1>, 32 * W_SLOT_BYTES)));
  W_TRY((set_smem(kw_rollout_random<32, 2, 1>, 32 * W_SLOT_BYTES)));
  W_TRY((set_smem(kw_rollout_random<28, 1, 1>, 28 * W_SLOT_BYTES)));
  W_TRY((set_smem(kw_rollout_random<14, 2, 1>, 14 * W_SLOT_BYTES)));
  W_TRY((set_smem(kw_rollout_random<32, 1, 2>, 32 * W_SLOT_BYTES)));
  W_TRY((set_smem(kw_rollout_random<16, 2, 2>, 16 * W_SLOT_BYTES)));
  W_TRY(set_smem(kw_step<8>, 8 * W_SLOT_BYTES));
  W_TRY(set_smem(kw_query<8, 3>, 8 * W_SLOT_BYTES));
  W_TRY(set_smem(kw_select_action<4>, 4 * W_HSLOT_BYTES));
  W_TRY(set_smem(kw_rollout_heuristic<4, 8>, 4 * W_HSLOT_BYTES));
  W_TRY(set_smem(kw_rollout_heuristic<4, 9>, 4 * W_HSLOT_BYTES));
  return cudaSuccess;
}

// shape: 0 = 8-warp CTAs x 4 per SM (1,024 threads, 64 registers), 1 = 4-warp CTAs x 8 per SM (same occupancy, finer CTAs),
//        2 = 8-warp CTAs x 8 per SM (2,048 threads, 32 registers); turn-synchronous CTAs: 3 = 8 warps x 4, 4 = 16 warps x 2,
//        5 = 32 warps x 1, 6 = 32 warps x 2 (32 registers), 7 = 28 warps x 1 (4,096 games = 147 CTAs: every SM busy), 8 = 14 warps x 2;
//        two barriers per turn instead of one vote per action: 9 = 32 warps x 1, 10 = 16 warps x 2
void sbw_rollout_random(const SbwCtx* c, int n, uint8_t* states, int max_steps, int32_t* steps, uint64_t* chain, int shape, int grid_override,
                        cudaStream_t st) {
  static const int WPCS[11] = {8, 4, 8, 8, 16, 32, 32, 28, 14, 32, 16}, PER_SM[11] = {4, 8, 8, 4, 2, 1, 2, 1, 2, 1, 2};
  if (shape < 0 || shape > 10) shape = 0;
  const int wpc = WPCS[shape], per_sm = PER_SM[shape];
  const int resident = c->sm_count * per_sm;  // CTAs in flight
  const int full = grid_for(n, wpc);
  int* q = nullptr;
  int grid = full;
  if (grid_override > 0 || full > resident) {  // persistent grid: warps take games from a counter
    q = c->d_queue;
    cudaMemsetAsync(q, 0, sizeof(int), st);
    grid = grid_override > 0 ? grid_override : resident;
  }
  unsigned long long* ch = (unsigned long long*)chain;
#define RR(W, B, S) kw_rollout_random<W, B, S><<<grid, W * 32, W * W_SLOT_BYTES, st>>>(n, states, max_steps, steps, ch, c->d_cards, c->d_wt, q)
  switch (shape) {
    case 0: RR(8, 4, 0); break;
    case 1: RR(4, 8, 0); break;
    case 2: RR(8, 8, 0); break;
    case 3: RR(8, 4, 1); break;
    case 4: RR(16, 2, 1); break;
    case 5: RR(32, 1, 1); break;
    case 6: RR(32, 2, 1); break;
    case 7: RR(28, 1, 1); break;
    case 8: RR(14, 2, 1); break;
    case 9: RR(32, 1, 2); break;
    default: RR(16, 2, 2); break;
  }
#undef RR
}
static inline int query_grid(const SbwCtx* c, int n, int wpc) {
  const int full = grid_for(n, wpc), cap = c->sm_count * 32;
  return full < cap ? full : cap;
}
void sbw_step(const SbwCtx* c, int n, uint8_t* states, const uint8_t* actions, int8_t* reward, uint8_t* done, uint8_t* err, uint32_t* next_masks,
              cudaStream_t st) {
  kw_step<8><<<query_grid(c, n, 8), 256, 8 * W_SLOT_BYTES, st>>>(n, states, actions, (i8*)reward, done, err, next_masks, c->d_cards, c->d_wt);
}
void sbw_legal_mask(const SbwCtx* c, int n, const uint8_t* states, uint32_t* masks, cudaStream_t st) {
  const int full = grid_for(n, 8), cap = c->sm_count * 8;  // persistent grid: 8 CTAs of 8 warps per SM
  kw_stream<8, 0><<<full < cap ? full : cap, 256, 0, st>>>(n, states, masks, nullptr, c->d_cards);
}
void sbw_observe(const SbwCtx* c, int n, const uint8_t* states, int32_t* obs, uint8_t* err, cudaStream_t st) {
  const int full = grid_for(n, 8), cap = c->sm_count * 8;
  kw_stream<8, 1><<<full < cap ? full : cap, 256, 0, st>>>(n, states, obs, err, c->d_cards);
}
void sbw_features(const SbwCtx* c, int n, const uint8_t* states, double* feat, uint8_t* err, cudaStream_t st) {
  const int full = grid_for(n, 8), cap = c->sm_count * 8;
  kw_stream<8, 2><<<full < cap ? full : cap, 256, 0, st>>>(n, states, feat, err, c->d_cards);
}
void sbw_expert_action(const SbwCtx* c, int n, uint8_t* states, uint8_t* actions, cudaStream_t st) {
  kw_query<8, 3><<<query_grid(c, n, 8), 256, 8 * W_SLOT_BYTES, st>>>(n, states, actions, nullptr, c->d_cards, c->d_wt);
}
void sbw_select_action(const SbwCtx* c, int n, const uint8_t* states, const double* weights, uint8_t* actions, double* scores, cudaStream_t st) {
  kw_select_action<4><<<query_grid(c, n, 4), 128, 4 * W_HSLOT_BYTES, st>>>(n, states, weights, actions, scores, c->d_cards, c->d_wt);
}
// shape 0: 8 CTAs of 4 warps per SM (1,024 threads, 64 registers); shape 1: 9 CTAs per SM (1,152 threads, 56 registers; 212 KB of shared memory)
void sbw_rollout_heuristic(const SbwCtx* c, int n, uint8_t* states, const double* w_first, const double* w_second, const int32_t* idx_first,
                           const int32_t* idx_second, int max_steps, int8_t* result, int32_t* steps, int shape, int grid_override, cudaStream_t st) {
  const int wpc = 4, per_sm = shape == 1 ? 9 : 8;
  const int resident = c->sm_count * per_sm, full = grid_for(n, wpc);
  int* q = nullptr;
  int grid = full;
  if (grid_override > 0 || full > resident) {
    q = c->d_queue + 1;
    cudaMemsetAsync(q, 0, sizeof(int), st);
    grid = grid_override > 0 ? grid_override : resident;
  }
  if (shape == 1)
    kw_rollout_heuristic<4, 9><<<grid, 128, 4 * W_HSLOT_BYTES, st>>>(n, states, w_first, w_second, idx_first, idx_second, max_steps, (i8*)result, steps,
                                                                    c->d_cards, c->d_wt, q);
  else
    kw_rollout_heuristic<4, 8><<<grid, 128, 4 * W_HSLOT_BYTES, st>>>(n, states, w_first, w_second, idx_first, idx_second, max_steps, (i8*)result, steps,
                                                                   c->d_cards, c->d_wt, q);
}
