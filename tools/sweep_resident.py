"""GPU box: refill mode, 1,024-thread CTAs: one vs two resident CTAs per SM (the kernel must be compiled for it)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
from monsoon_b200.engine import Engine
eng = Engine(0)
dev = eng.device
for n in (131072, 163840, 196608, 229376, 262144, 524288):
    seeds = torch.arange(n, dtype=torch.int64, device=dev) + 777
    for refill, dense in ((-1, 0), (1, 0), (1, 1)):
        bs, ctas = 1024, 1 + dense
        eng.set_option("block_sync", 1024)
        eng.set_option("refill", refill)
        eng.set_option("dense", dense)
        best = 1e9
        for rep in range(3):
            st = eng.reset(seeds)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            steps = eng.rollout_random(st, 400)
            e1.record()
            torch.cuda.synchronize()
            if rep:
                best = min(best, e0.elapsed_time(e1))
        print("n %7d refill %2d dense %d: %8.2f ms  %7.1f M env-steps/s" % (n, refill, dense, best, int(steps.sum()) / best / 1e3), flush=True)
