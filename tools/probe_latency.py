"""GPU box: per-step latency of ONE game's chain as a function of how many games share the chip (instruction-cache probe).
time / max(steps) of a batch = time per env step along the longest game."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
from monsoon_b200.engine import Engine
eng = Engine(0); dev = eng.device
for eng_id, shape in ((1, 1), (1, 5), (0, 0)):
    eng.set_option("engine", eng_id); eng.set_option("w_shape", shape)
    for n in (148, 592, 1184, 2368, 4096, 8192):
        seeds = torch.arange(n, dtype=torch.int64, device=dev)
        best = 1e9
        for rep in range(4):
            st = eng.reset(seeds)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); steps = eng.rollout_random(st, 400); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        mx = int(steps.max()); tot = int(steps.sum())
        print("engine %d shape %d  games %5d  %7.3f ms  max steps %3d  -> %6.2f us per step of the longest game   %6.1f M env-steps/s" %
              (eng_id, shape, n, best, mx, best * 1e3 / mx, tot / best / 1e3), flush=True)
