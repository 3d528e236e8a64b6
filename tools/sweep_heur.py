import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np, torch
from monsoon_b200.engine import Engine
eng = Engine(0); dev = eng.device
for n in (4096, 16384, 65536):
    P = 256; GPI = n // P
    w = torch.from_numpy(np.concatenate([np.random.RandomState(42).uniform(0, 1, (P, 10)), np.random.RandomState(7).uniform(0, 1, (1, 10))])).to(dev)
    i1 = (torch.arange(n, device=dev) // GPI).to(torch.int32); i2 = torch.full((n,), P, dtype=torch.int32, device=dev)
    seeds = torch.arange(n, dtype=torch.int64, device=dev)
    ref = None
    for wpc in (4, 8, 16, 32):
        eng.lib.sb_set_option(eng.h, b"heur_wpc", wpc)
        best = 1e9
        for rep in range(2):
            st = eng.reset(seeds)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); res, steps = eng.rollout_heuristic(st, w, w, i1, i2, max_steps=400); e1.record()
            torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
        h = hash(st.cpu().numpy().tobytes() + res.cpu().numpy().tobytes()); ref = ref or h
        print("games %6d warps/CTA %2d  %8.1f ms  %7.0f games/s  %6.2f M env-steps/s %s" % (n, wpc, best, n / best * 1e3, int(steps.sum()) / best / 1e3, "ok" if h == ref else "MISMATCH"), flush=True)
