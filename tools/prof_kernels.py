"""GPU box, under `ncu --profile-from-start off --set full`: ONE measured launch of every kernel behind the C ABI except the
whole-game rollouts (which have their own captures).  Every API is called once untimed first (warm-up, outside the profiled
range), then once between cudaProfilerStart / Stop.
  thread-per-game: k_reset, k_generate_decks, k_legal_mask, k_observe, k_features, k_expert_action, k_step (1 M games),
                   k_select_action, k_accumulate_fitness, k_count_aborted, k_eval_schedule, k_es_offspring / select / reset_sigmas / inject_diversity
  warp-per-game:   kw_stream<legal mask / observation / features> (streaming, straight from the packed record), kw_query<expert action>,
                   kw_step, kw_select_action"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np, torch
from monsoon_b200.engine import Engine
eng = Engine(0); dev = eng.device
N_BIG, N_MID = 1 << 20, 1 << 16
seeds_big = torch.arange(N_BIG, dtype=torch.int64, device=dev)
eng.set_option("engine", 0)
st_big = eng.reset(seeds_big); eng.rollout_random(st_big, max_steps=30)   # mid-game records
st_mid = st_big[:N_MID].clone()
w_mid = torch.from_numpy(np.random.RandomState(1).uniform(0, 1, (16384, 10))).to(dev)
masks = eng.legal_mask(st_big)
mm = masks.to(torch.int64) & 0xFFFFFFFF
acts = torch.full((N_BIG,), 155, dtype=torch.int64, device=dev)
for k in range(4, -1, -1):  # lowest legal action
    v = mm[:, k]; low = v & -v
    idx = torch.log2(low.clamp(min=1).to(torch.float64)).to(torch.int64) + 32 * k
    acts = torch.where(v != 0, idx, acts)
acts = acts.to(torch.uint8)
mu = 512
esw = torch.rand((2 * mu, 10), dtype=torch.float64, device=dev); ess = torch.full((2 * mu, 10), 0.1, dtype=torch.float64, device=dev)
fit = torch.rand(2 * mu, dtype=torch.float64, device=dev)
res = torch.randint(-2, 2, (N_MID,), device=dev).to(torch.int8); idx = (torch.arange(N_MID, device=dev) % 256).to(torch.int32)
counts = torch.zeros((256, 3), dtype=torch.int32, device=dev)
nm = torch.empty_like(masks)


def thread_calls():
    eng.set_option("engine", 0)
    eng.reset(seeds_big[:N_MID])
    eng.generate_decks(seeds_big, 5, 3, factions=torch.ones((N_BIG, 2), dtype=torch.uint8, device=dev))
    eng.legal_mask(st_big)
    eng.observe(st_mid)
    eng.features(st_big)
    eng.expert_action(st_mid.clone())
    eng.step(st_big.clone(), acts, next_masks=nm)
    eng.select_action(st_mid[:16384], w_mid)
    eng.accumulate_fitness(res, idx, counts)
    eng.count_aborted(st_mid, res)
    eng.eval_schedule("round_robin", 256, 261, 4, 1, 0, 0, N_MID)
    eng.es_offspring(1, 1, mu, mu, 0.1, 0.01, 1e-5, esw, ess)
    eng.es_select(mu, fit, esw, ess)
    eng.es_reset_sigmas(1, 1, 0.1, ess[:mu].contiguous())
    eng.es_inject_diversity(1, 1, 0.1, 0.01, 1e-5, 0.1, esw[:mu].contiguous(), ess[:mu].contiguous())


def warp_calls():
    eng.set_option("engine", 1)
    eng.legal_mask(st_big)
    eng.observe(st_mid)
    eng.features(st_big)
    eng.expert_action(st_mid.clone())
    eng.step(st_big[:4096].clone(), acts[:4096], next_masks=nm[:4096])
    eng.step(st_big.clone(), acts, next_masks=nm)
    eng.select_action(st_mid[:4096], w_mid[:4096])


thread_calls(); warp_calls(); torch.cuda.synchronize()
torch.cuda.profiler.start()
thread_calls(); warp_calls(); torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
