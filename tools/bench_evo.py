"""GPU box: BASELINE config 3 on ONE GPU -- population P x G games/individual as FIRST vs one fixed
baseline vector, heuristic agents on both sides, default decks, max 400 env steps."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np, torch
from monsoon_b200.engine import Engine
eng = Engine(0)
dev = eng.device
P = int(sys.argv[1]) if len(sys.argv) > 1 else 256
GPI = int(sys.argv[2]) if len(sys.argv) > 2 else 256
n = P * GPI
w = np.concatenate([np.random.RandomState(42).uniform(0, 1, (P, 10)), np.random.RandomState(7).uniform(0, 1, (1, 10))])
w = torch.from_numpy(w).to(dev)
idx_first = (torch.arange(n, device=dev) // GPI).to(torch.int32)
idx_second = torch.full((n,), P, dtype=torch.int32, device=dev)
seeds = torch.arange(n, dtype=torch.int64, device=dev)
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    st = eng.reset(seeds)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res, steps = eng.rollout_heuristic(st, w, w, idx_first, idx_second, max_steps=400)
    e1.record()
    counts = torch.zeros((P + 1, 3), dtype=torch.int32, device=dev)
    eng.accumulate_fitness(res, idx_first, counts)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    ms = e0.elapsed_time(e1)
    tot = int(steps.sum())
    r = res.cpu().numpy()
    print("pop %d x %d games = %d games: kernel %.1f ms, wall %.3f s, %.0f games/s, %.2f M env-steps/s, mean steps %.1f, first wins %.3f draws %.3f aborted %d"
          % (P, GPI, n, ms, dt, n / (ms * 1e-3), tot / ms / 1e3, tot / n, (r == 0).mean(), (r == -1).mean(), (r == -2).sum()), flush=True)
