"""GPU box: random rollout (reset + whole games) and heuristic rollout under both engines and the warp engine's CTA shapes.
Prints env-steps/s and a hash of the final states so that the engines can be compared for identical output.
  python tools/sweep_engines.py [random sizes] [heuristic sizes]"""
import hashlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np, torch
from monsoon_b200.engine import Engine
eng = Engine(0); dev = eng.device
rs = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "4096,16384,65536,262144").split(",") if x]
hs = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "4096,16384,65536").split(",") if x]
CFG = [("thread", dict(engine=0))] + [("warp s%d" % k, dict(engine=1, w_shape=k)) for k in [int(x) for x in os.environ.get("W_SHAPES", "0,1,2,3,4,5,6").split(",")]]
flush = torch.empty(160 << 20, dtype=torch.uint8, device=dev)
for n in rs:
    seeds = torch.arange(n, dtype=torch.int64, device=dev)
    for name, opts in CFG:
        for k, v in opts.items(): eng.set_option(k, v)
        ts = []
        for rep in range(4):
            flush.fill_(rep)
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record(); st = eng.reset(seeds); e1.record(); steps = eng.rollout_random(st, 400); e2.record()
            torch.cuda.synchronize(); ts.append((e1.elapsed_time(e2), e0.elapsed_time(e1)))
        best, rst = min(ts)
        h = hashlib.sha1(st.cpu().numpy().tobytes()).hexdigest()[:10]
        tot = int(steps.sum())
        print("random %7d games  %-8s rollout %8.3f ms (reset %.3f)  %8.1f M env-steps/s  sha1 %s" % (n, name, best, rst, tot / best / 1e3, h), flush=True)
HCFG = [("thread", dict(engine=0)), ("warp h0", dict(engine=1, w_hshape=0)), ("warp h1", dict(engine=1, w_hshape=1))]
for n in hs:
    P = 256; GPI = max(n // P, 1)
    w = torch.from_numpy(np.concatenate([np.random.RandomState(42).uniform(0, 1, (P, 10)), np.random.RandomState(7).uniform(0, 1, (1, 10))])).to(dev)
    i1 = (torch.arange(n, device=dev) // GPI).clamp(max=P - 1).to(torch.int32); i2 = torch.full((n,), P, dtype=torch.int32, device=dev)
    seeds = torch.arange(n, dtype=torch.int64, device=dev)
    for name, opts in HCFG:
        for k, v in opts.items(): eng.set_option(k, v)
        ts = []
        for rep in range(3):
            st = eng.reset(seeds)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); res, steps = eng.rollout_heuristic(st, w, w, i1, i2, max_steps=400); e1.record()
            torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        h = hashlib.sha1(st.cpu().numpy().tobytes() + res.cpu().numpy().tobytes()).hexdigest()[:10]
        best = min(ts)
        print("heuristic %7d games  %-8s %8.1f ms  %8.0f games/s  %6.2f M env-steps/s  sha1 %s" % (n, name, best, n / best * 1e3, int(steps.sum()) / best / 1e3, h), flush=True)
