"""GPU box: the random-deck heuristic games of tests/parity_at_scale.py that differ from the oracle -- which engine status, at which step,
under both heuristic kernels."""
import os, sys, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np, torch
import sb_oracle as oracle
from monsoon_b200.engine import Engine
from monsoon_b200._card_table import CARDS
EXCLUDE = ()
eng = Engine(0); dev = eng.device
def rdecks(seed):
    rng = random.Random(seed); decks = []; f = []
    for _ in range(2):
        fa = rng.choice([1, 2, 3, 4])
        pool = [i for i, c in enumerate(CARDS[:113]) if i > 0 and c["faction"] in (0, fa) and c["name"] not in EXCLUDE]
        decks.append(rng.sample(pool, 12)); f.append(fa)
    return decks, f
NR = 10000
seeds = np.arange(NR, dtype=np.int64) + 1300000
w1 = np.random.RandomState(11).uniform(0, 1, (20000, 10)); w2 = np.random.RandomState(12).uniform(0, 1, (20000, 10))
dd = [rdecks(int(s)) for s in seeds]
decks = torch.tensor([d for d, _f in dd], dtype=torch.uint8, device=dev); fac = torch.tensor([f for _d, f in dd], dtype=torch.uint8, device=dev)
out = {}
for pack in (0, 1):
    eng.set_option("heur_pack", pack)
    st = eng.reset(torch.from_numpy(seeds).to(dev), decks, fac)
    res, steps = eng.rollout_heuristic(st, torch.from_numpy(w1[:NR]).to(dev), torch.from_numpy(w2[:NR]).to(dev), max_steps=400)
    out[pack] = (res.cpu().numpy(), steps.cpu().numpy(), st.cpu().numpy())
print("pack 0 == pack 1:", all(np.array_equal(a, b) for a, b in zip(out[0], out[1])))
bad = 0
for i in range(NR):
    d, f = dd[i]
    s = oracle.new_game(int(seeds[i]), d[0], d[1], f[0], f[1])
    r, acts = oracle.play_heuristic(s, w1[i], w2[i], 400)
    g = out[1]
    if r != int(g[0][i]) or len(acts) != int(g[1][i]) or s.tobytes() != g[2][i].tobytes():
        bad += 1
        names = [[CARDS[c]["name"] for c in dk] for dk in d]
        print("seed", int(seeds[i]), "oracle res/steps/err", r, len(acts), int(s[18]), "| gpu", int(g[0][i]), int(g[1][i]), int(g[2][i][18]), "| pack0", int(out[0][0][i]), int(out[0][1][i]), int(out[0][2][i][18]))
        print("   decks", names)
print("mismatches", bad)
