import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
from monsoon_b200.engine import Engine
eng = Engine(0)
sizes = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "4096,65536,262144").split(",")]
for n in sizes:
    seeds = torch.arange(n, dtype=torch.int64, device=eng.device) + 12345
    ref = None
    for bs in (0, 128, 256, 512, 1024, -1):
        eng.lib.sb_set_option(eng.h, b"block_sync", bs)
        best = 1e9
        for rep in range(3):
            st = eng.reset(seeds)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); steps = eng.rollout_random(st, 400); e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        tot = int(steps.sum()); h = hash(st.cpu().numpy().tobytes()); ref = ref or h
        print("games %7d block_sync %3d  %8.2f ms  %7.2f M steps/s %s" % (n, bs, best, tot / best / 1e3, "ok" if h == ref else "MISMATCH"), flush=True)
