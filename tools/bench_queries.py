"""GPU box: the state-streaming query kernels (legal mask, observation, features, one heuristic decision, expert action)
under both engines: time per launch and achieved HBM bytes (records in + results out)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np, torch
from monsoon_b200.engine import Engine
eng = Engine(0); dev = eng.device
def timed(f, reps=8):
    ts = []
    for i in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        if i >= 2: ts.append(e0.elapsed_time(e1))
    return min(ts)
for n in (4096, 65536, 1048576):
    seeds = torch.arange(n, dtype=torch.int64, device=dev)
    eng.set_option("engine", 0)
    st = eng.reset(seeds); eng.rollout_random(st, max_steps=30)
    w = torch.from_numpy(np.random.RandomState(1).uniform(0, 1, (n, 10))).to(dev)
    for engine_id in (0, 1):
        eng.set_option("engine", engine_id)
        name = "thread" if engine_id == 0 else "warp  "
        rows = [("legal_mask", lambda: eng.legal_mask(st), 512 + 20), ("features", lambda: eng.features(st), 512 + 80),
                ("expert_action", lambda: eng.expert_action(st.clone()), 512 + 4)]
        rows += [("observe", lambda: eng.observe(st), 512 + 2160)]
        if n <= 65536:
            rows += [("select_action", lambda: eng.select_action(st, w), 512 + 81)]
        for what, f, bytes_per in rows:
            ms = timed(f)
            print("%s %-14s n=%8d  %8.3f ms  %8.1f M games/s  %7.1f GB/s" % (name, what, n, ms, n / ms / 1e3, bytes_per * n / ms / 1e6), flush=True)
