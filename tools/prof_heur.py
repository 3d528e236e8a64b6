"""GPU box: one heuristic rollout launch (target for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np, torch
from monsoon_b200.engine import Engine
eng = Engine(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
w = torch.from_numpy(np.random.RandomState(42).uniform(0, 1, (n, 10))).to(eng.device)
w2 = torch.from_numpy(np.random.RandomState(7).uniform(0, 1, (n, 10))).to(eng.device)
seeds = torch.arange(n, dtype=torch.int64, device=eng.device)
for rep in range(2):
    st = eng.reset(seeds)
    res, steps = eng.rollout_heuristic(st, w, w2, max_steps=400)
    torch.cuda.synchronize()
print("steps", int(steps.sum()))
