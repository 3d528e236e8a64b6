"""GPU box: the turn-synchronous warp-per-game rollout at the named size as a PERSISTENT grid -- fewer resident games per SM than the
batch holds, a warp takes the next game when its game ends.   python tools/sweep_wgrid.py [n]"""
import hashlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
from monsoon_b200.engine import Engine
eng = Engine(0); dev = eng.device
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
seeds = torch.arange(n, dtype=torch.int64, device=dev)
flush = torch.empty(160 << 20, dtype=torch.uint8, device=dev)
eng.set_option("engine", 1)
CFG = [(5, 0), (7, 0)] + [(s, g) for s in (4, 3, 5, 7, 9, 10) for g in (148, 296)]
for shape, grid in CFG:
    eng.set_option("w_shape", shape); eng.set_option("w_grid", grid)
    ts = []
    for rep in range(4):
        flush.fill_(rep)
        st = eng.reset(seeds)
        e1, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e1.record(); steps = eng.rollout_random(st, 400); e2.record()
        torch.cuda.synchronize(); ts.append(e1.elapsed_time(e2))
    h = hashlib.sha1(st.cpu().numpy().tobytes()).hexdigest()[:10]
    print("random %6d games  warp s%-2d grid %4d  %7.3f ms  %6.1f M env-steps/s  sha1 %s" % (n, shape, grid, min(ts), int(steps.sum()) / min(ts) / 1e3, h), flush=True)
