"""GPU box: what makes a turn-synchronous 32-warp CTA slow -- the spread of step costs between its games, or sharing the SM?
Same game in every warp of a CTA (no spread, perfectly aligned) against 32 different games, for the warp engine's shapes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
from monsoon_b200.engine import Engine
eng = Engine(0); dev = eng.device
eng.set_option("engine", 1)
# the longest game among seeds 0..4095 has 200 steps: find it
seeds = torch.arange(4096, dtype=torch.int64, device=dev)
st = eng.reset(seeds); steps = eng.rollout_random(st, 400)
longest = int(steps.argmax()); print("longest game: seed", longest, "steps", int(steps.max()), flush=True)
def run(name, seeds, shape):
    eng.set_option("w_shape", shape)
    best = 1e9
    for rep in range(4):
        st = eng.reset(seeds)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); steps = eng.rollout_random(st, 400); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    mx = int(steps.max())
    print("%-44s shape %d  %7.3f ms  max steps %3d -> %6.2f us per step of the longest game" % (name, shape, best, mx, best * 1e3 / mx), flush=True)
QUICK = os.environ.get("PROBE_QUICK")
for shape in ((1, 5) if QUICK else (1, 5, 9)):
    if QUICK:
        run("1 game", torch.full((1,), longest, dtype=torch.int64, device=dev), shape)
        run("32 copies of the longest game (1 CTA)", torch.full((32,), longest, dtype=torch.int64, device=dev), shape)
        run("longest + 31 others (1 CTA)", torch.cat([torch.tensor([longest], device=dev), torch.arange(31, device=dev)]).to(torch.int64), shape)
        run("4096 different games", seeds, shape)
        continue
    run("1 game", torch.full((1,), longest, dtype=torch.int64, device=dev), shape)
    run("32 copies of the longest game (1 CTA)", torch.full((32,), longest, dtype=torch.int64, device=dev), shape)
    run("4096 copies of the longest game", torch.full((4096,), longest, dtype=torch.int64, device=dev), shape)
    run("longest + 31 others (1 CTA)", torch.cat([torch.tensor([longest], device=dev), torch.arange(31, device=dev)]).to(torch.int64), shape)
    run("4096 different games", seeds, shape)
    for k in (2, 4, 8, 16):
        run("%d copies of the longest game" % k, torch.full((k,), longest, dtype=torch.int64, device=dev), shape)
