"""GPU box (target for ncu): one heuristic rollout launch of 16,384 games, K games per warp (argv[1], 0 = one game per warp)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np, torch
from monsoon_b200.engine import Engine
eng = Engine(0); dev = eng.device
eng.set_option("engine", 0); eng.set_option("heur_pack", int(sys.argv[1]) if len(sys.argv) > 1 else 4)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
P = 256; GPI = n // P
w = torch.from_numpy(np.concatenate([np.random.RandomState(42).uniform(0, 1, (P, 10)), np.random.RandomState(7).uniform(0, 1, (1, 10))])).to(dev)
i1 = (torch.arange(n, device=dev) // GPI).to(torch.int32); i2 = torch.full((n,), P, dtype=torch.int32, device=dev)
st = eng.reset(torch.arange(n, dtype=torch.int64, device=dev))
res, steps = eng.rollout_heuristic(st, w, w, i1, i2, max_steps=400)
torch.cuda.synchronize()
print("steps", int(steps.sum()))
