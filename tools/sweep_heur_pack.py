"""GPU box: heuristic rollout, one game per warp (k_rollout_heuristic) against K games per warp (k_rollout_heuristic_packed).
  python tools/sweep_heur_pack.py [sizes] [packs]"""
import hashlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np, torch
from monsoon_b200.engine import Engine
eng = Engine(0); dev = eng.device
sizes = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "4096,16384,65536").split(",")]
packs = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "0,2,4,8").split(",")]
eng.set_option("engine", 0)
for n in sizes:
    P = 256; GPI = max(n // P, 1)
    w = torch.from_numpy(np.concatenate([np.random.RandomState(42).uniform(0, 1, (P, 10)), np.random.RandomState(7).uniform(0, 1, (1, 10))])).to(dev)
    i1 = (torch.arange(n, device=dev) // GPI).clamp(max=P - 1).to(torch.int32); i2 = torch.full((n,), P, dtype=torch.int32, device=dev)
    seeds = torch.arange(n, dtype=torch.int64, device=dev)
    for pack in packs:
        eng.set_option("heur_pack", pack)
        ts = []
        for rep in range(3):
            st = eng.reset(seeds)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); res, steps = eng.rollout_heuristic(st, w, w, i1, i2, max_steps=400); e1.record()
            torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        h = hashlib.sha1(st.cpu().numpy().tobytes() + res.cpu().numpy().tobytes() + steps.cpu().numpy().tobytes()).hexdigest()[:10]
        best = min(ts)
        print("heuristic %7d games  pack %d  %8.1f ms  %8.0f games/s  %6.2f M env-steps/s  sha1 %s" % (n, pack, best, n / best * 1e3, int(steps.sum()) / best / 1e3, h), flush=True)
