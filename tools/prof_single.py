"""GPU box (target for ncu): the longest default-deck game (seed 1195, 200 env steps) alone on the chip (independent-warp shape),
then 32 copies of it in one turn-synchronous CTA -- the latency chain of one warp, and the same chain under a full SM."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
from monsoon_b200.engine import Engine
eng = Engine(0); dev = eng.device
eng.set_option("engine", 1)
for shape, n in ((1, 1), (5, 32)):
    eng.set_option("w_shape", shape)
    st = eng.reset(torch.full((n,), 1195, dtype=torch.int64, device=dev))
    steps = eng.rollout_random(st, 400)
    torch.cuda.synchronize()
    print("shape", shape, "games", n, "steps", int(steps.max()))
