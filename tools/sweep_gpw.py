"""GPU box: sweep games-per-warp for the thread-per-game rollout kernel against batch size."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
from monsoon_b200.engine import Engine
eng = Engine(0)
gpws = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "32,16,8,4,2,1,0").split(",")]
sizes = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "4096,16384,65536,262144").split(",")]
eng.lib.sb_set_option(eng.h, b"lanes_per_game", 1)
for n in sizes:
    seeds = torch.arange(n, dtype=torch.int64, device=eng.device) + 12345
    ref = None
    for gpw in gpws:
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
        print("games %7d gpw %2d  %8.2f ms  %7.2f M steps/s  max_steps %d %s" % (n, gpw, best, tot / best / 1e3, int(steps.max()), "ok" if h == ref else "MISMATCH"), flush=True)
