"""GPU box: one reset + one rollout launch at a given batch size (target for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
from monsoon_b200.engine import Engine
eng = Engine(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
seeds = torch.arange(n, dtype=torch.int64, device=eng.device) + 777
for rep in range(2):
    st = eng.reset(seeds)
    steps = eng.rollout_random(st, 400)
    torch.cuda.synchronize()
print("steps", int(steps.sum()))
