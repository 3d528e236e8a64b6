"""GPU box: turn-synchronous vs lock-step warp schedule, by batch size and games per warp."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
from monsoon_b200.engine import Engine
eng = Engine(0)
sizes = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "4096,16384,65536,262144").split(",")]
for n in sizes:
    seeds = torch.arange(n, dtype=torch.int64, device=eng.device) + 12345
    ref = None
    for sync, gpw in [(0, 32), (1, 32), (1, 16), (1, 8), (0, 0), (1, 0)]:
        eng.lib.sb_set_option(eng.h, b"turn_sync", sync)
        eng.lib.sb_set_option(eng.h, b"games_per_warp", gpw)
        best = 1e9
        for rep in range(3):
            st = eng.reset(seeds)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            steps = eng.rollout_random(st, 400)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        tot = int(steps.sum())
        h = hash(st.cpu().numpy().tobytes())
        ref = ref or h
        print("games %7d turn_sync %d gpw %2d  %8.2f ms  %7.2f M steps/s %s" % (n, sync, gpw, best, tot / best / 1e3, "ok" if h == ref else "MISMATCH"), flush=True)
