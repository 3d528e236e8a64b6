"""GPU box: time the heuristic rollout (config-3 shape: 256 individuals vs one baseline) at a few batch sizes; prints a hash
of the final states + results so that two builds can be compared for identical output."""
import hashlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np, torch
from monsoon_b200 import _lib
if os.environ.get('SB_LIB'):
    _lib.SO_PATH = os.path.abspath(os.environ['SB_LIB'])  # A/B: time another build of the library
from monsoon_b200.engine import Engine
eng = Engine(0); dev = eng.device
for kv in os.environ.get('SB_OPTS', '').split(','):  # e.g. SB_OPTS=heur_iw=8,heur_refill=1
    if kv:
        k, v = kv.split('=')
        assert eng.lib.sb_set_option(eng.h, k.encode(), int(v)) == 0, kv
sizes = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "4096,16384,65536").split(",")]
for n in sizes:
    P = 256; GPI = max(n // P, 1)
    w = torch.from_numpy(np.concatenate([np.random.RandomState(42).uniform(0, 1, (P, 10)), np.random.RandomState(7).uniform(0, 1, (1, 10))])).to(dev)
    i1 = (torch.arange(n, device=dev) // GPI).clamp(max=P - 1).to(torch.int32); i2 = torch.full((n,), P, dtype=torch.int32, device=dev)
    seeds = torch.arange(n, dtype=torch.int64, device=dev)
    ts = []
    for rep in range(3):
        st = eng.reset(seeds)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); res, steps = eng.rollout_heuristic(st, w, w, i1, i2, max_steps=400); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    h = hashlib.sha1(st.cpu().numpy().tobytes() + res.cpu().numpy().tobytes()).hexdigest()[:12]
    best = min(ts)
    print("games %6d  %8.1f ms  %7.0f games/s  %6.2f M env-steps/s  sha1 %s" % (n, best, n / best * 1e3, int(steps.sum()) / best / 1e3, h), flush=True)
