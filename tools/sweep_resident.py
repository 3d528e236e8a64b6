"""GPU box: refill mode, resident threads per SM vs throughput (does the thread-private working set fit L2?)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
from monsoon_b200.engine import Engine
eng = Engine(0)
dev = eng.device
n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
seeds = torch.arange(n, dtype=torch.int64, device=dev) + 777
eng.set_option("refill", 1)
for bs, ctas in ((1024, 1), (768, 2)):
    eng.set_option("block_sync", bs)
    eng.set_option("refill_ctas", ctas)
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
    print("n %d cta %4d x %d/SM (%4d thr/SM): %8.2f ms  %7.1f M env-steps/s" % (n, bs, ctas, bs * ctas, best, int(steps.sum()) / best / 1e3), flush=True)
