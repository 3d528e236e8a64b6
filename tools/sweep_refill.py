"""GPU box: random-rollout throughput with and without lane refill at several batch sizes; checks identical results."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
from monsoon_b200.engine import Engine
eng = Engine(0)
dev = eng.device
for n in (131072, 196608, 262144, 524288, 1048576):
    seeds = torch.arange(n, dtype=torch.int64, device=dev) + 777
    ref = None
    for refill, bs in ((0, -1), (1, 1024), (1, 512)):
        eng.set_option("refill", refill)
        eng.set_option("block_sync", bs)
        best = 1e9
        for rep in range(3):
            st = eng.reset(seeds)
            chain = torch.zeros(n, dtype=torch.int64, device=dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            steps = eng.rollout_random(st, 400, chain=chain if rep == 0 else None)
            e1.record()
            torch.cuda.synchronize()
            if rep:
                best = min(best, e0.elapsed_time(e1))
            else:
                sig = (int(steps.sum()), int(chain.sum()), int(st.to(torch.int64).sum()))
        if ref is None:
            ref = sig
        tot = sig[0]
        print("n %8d refill %d bs %5d: %8.2f ms  %7.1f M env-steps/s  same=%s" % (n, refill, bs, best, tot / best / 1e3, sig == ref), flush=True)
