// sbw_launch.h -- launch interface of the warp-per-game kernels (sbw_kernels.cu), used by the C ABI in sb_kernels.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "sb_defs.h"

struct SbwCtx {
  const DCard* d_cards;  // card table in HBM (staged into shared memory by every CTA)
  const double* d_wt;    // draw weights f^n(1)
  int* d_queue;          // [0] random rollout, [1] heuristic rollout: next-game counters of the persistent grids
  int sm_count;
};
cudaError_t sbw_init(void);
int sbw_wg_bytes(void);
void sbw_rollout_random(const SbwCtx* c, int n, uint8_t* states, int max_steps, int32_t* steps, uint64_t* chain, int shape, int grid_override, cudaStream_t st);
void sbw_step(const SbwCtx* c, int n, uint8_t* states, const uint8_t* actions, int8_t* reward, uint8_t* done, uint8_t* err, uint32_t* next_masks, cudaStream_t st);
void sbw_legal_mask(const SbwCtx* c, int n, const uint8_t* states, uint32_t* masks, cudaStream_t st);
void sbw_observe(const SbwCtx* c, int n, const uint8_t* states, int32_t* obs, uint8_t* err, cudaStream_t st);
void sbw_features(const SbwCtx* c, int n, const uint8_t* states, double* feat, uint8_t* err, cudaStream_t st);
void sbw_expert_action(const SbwCtx* c, int n, uint8_t* states, uint8_t* actions, cudaStream_t st);
void sbw_select_action(const SbwCtx* c, int n, const uint8_t* states, const double* weights, uint8_t* actions, double* scores, cudaStream_t st);
void sbw_rollout_heuristic(const SbwCtx* c, int n, uint8_t* states, const double* w_first, const double* w_second, const int32_t* idx_first,
                           const int32_t* idx_second, int max_steps, int8_t* result, int32_t* steps, int shape, int grid_override, cudaStream_t st);
