"""TEST TOOL (GPU box; lives under tests/ because it uses the oracle): rollout kernel vs CPU oracle on many more games
than the committed fixtures hold -- N default-deck and N random-deck games (all 112 cards minus UP01-03 / S203): digest
chain of every step, step count and final record.   python tests/parity_at_scale.py 250000
(tests/test_gpu_scale.py runs the same comparison at a size that fits the GPU test suite.)"""
import os, sys, random, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # repo root
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np, torch
import sb_oracle as oracle
from monsoon_b200.engine import Engine, DEFAULT_DECKS, DEFAULT_FACTIONS, deck_indices
from monsoon_b200._card_table import CARDS
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
# SB_SCALE_ALL_CARDS=1: also the cards the reference cannot observe / orders by hash (UP01-03, S203): engine vs oracle only
EXCLUDE = () if os.environ.get("SB_SCALE_ALL_CARDS") else ("UP01", "UP02", "UP03", "S203")
eng = Engine(0); dev = eng.device
if os.environ.get("SB_ENGINE"):  # force one engine (0 thread-per-game, 1 warp-per-game) instead of the per-kernel policy
    eng.set_option("engine", int(os.environ["SB_ENGINE"]))
FNV = 0x100000001B3; M64 = (1 << 64) - 1


def chain_of(d):
    c = 0
    for x in d:
        c = ((c ^ int(x)) * FNV) & M64
    return c


def rdecks(seed):
    rng = random.Random(seed); decks = []; f = []
    for _ in range(2):
        fa = rng.choice([1, 2, 3, 4])
        pool = [i for i, c in enumerate(CARDS[:113]) if i > 0 and c["faction"] in (0, fa) and c["name"] not in EXCLUDE]
        decks.append(rng.sample(pool, 12)); f.append(fa)
    return decks, f


for label, use_random in (("default decks", False), ("random decks", True)):
    seeds = np.arange(N, dtype=np.int64) + (700000 if use_random else 400000)
    if use_random:
        dd = [rdecks(int(s)) for s in seeds]
        decks = torch.tensor([d for d, _f in dd], dtype=torch.uint8, device=dev); fac = torch.tensor([f for _d, f in dd], dtype=torch.uint8, device=dev)
        st = eng.reset(torch.from_numpy(seeds).to(dev), decks, fac)
    else:
        d0, d1 = (deck_indices(d) for d in DEFAULT_DECKS)
        st = eng.reset(torch.from_numpy(seeds).to(dev))
    chain = torch.zeros(N, dtype=torch.int64, device=dev)
    steps = eng.rollout_random(st, 400, chain=chain)
    g_steps, g_chain, g_host = steps.cpu().numpy(), chain.cpu().numpy().view(np.uint64), st.cpu().numpy()
    bad = flagged = tot = 0
    t0 = time.time()
    for i in range(N):
        if use_random:
            d, f = dd[i]; s = oracle.new_game(int(seeds[i]), d[0], d[1], f[0], f[1])
        else:
            s = oracle.new_game(int(seeds[i]), d0, d1, *DEFAULT_FACTIONS)
        _a, dig, _m = oracle.rollout_random(s, 400)
        tot += len(dig)
        flagged += s[18] != 0
        if len(dig) != g_steps[i] or chain_of(dig) != int(g_chain[i]) or s.tobytes() != g_host[i].tobytes():
            bad += 1
            if bad < 5:
                print("MISMATCH", label, "seed", int(seeds[i]), len(dig), int(g_steps[i]), int(s[18]), int(g_host[i][18]))
    print("%s: %d games, %d env steps, %d flagged by both engines, %d mismatches (oracle %.0f s)" % (label, N, tot, flagged, bad, time.time() - t0), flush=True)

# heuristic agents on both seats (config 1 shape): winner, length and final record of whole games
NH = min(max(N // 5, 200), 20000)
seeds = np.arange(NH, dtype=np.int64) + 900000
w1 = np.random.RandomState(11).uniform(0, 1, (NH, 10))
w2 = np.random.RandomState(12).uniform(0, 1, (NH, 10))
d0, d1 = (deck_indices(d) for d in DEFAULT_DECKS)
st = eng.reset(torch.from_numpy(seeds).to(dev))
res, steps = eng.rollout_heuristic(st, torch.from_numpy(w1).to(dev), torch.from_numpy(w2).to(dev), max_steps=400)
g_res, g_steps, g_host = res.cpu().numpy(), steps.cpu().numpy(), st.cpu().numpy()
bad = tot = 0
t0 = time.time()
for i in range(NH):
    s = oracle.new_game(int(seeds[i]), d0, d1, *DEFAULT_FACTIONS)
    r, acts = oracle.play_heuristic(s, w1[i], w2[i], 400)
    tot += len(acts)
    if r != int(g_res[i]) or len(acts) != int(g_steps[i]) or s.tobytes() != g_host[i].tobytes():
        bad += 1
        if bad < 5:
            print("MISMATCH heuristic seed", int(seeds[i]), r, int(g_res[i]), len(acts), int(g_steps[i]))
print("heuristic agents: %d games, %d env steps, %d mismatches (oracle %.0f s)" % (NH, tot, bad, time.time() - t0), flush=True)


# heuristic agents with RANDOM decks (all cards of the pool above): the forks of the evaluation kernel reach every card effect
# every second game holds UA20 (its copies grow hands and decks beyond the packed layout: the capacity status must come up at the
# same step in both engines, also when it is a CANDIDATE that overflows)
NR = max(NH // 2, 100)
seeds = np.arange(NR, dtype=np.int64) + 1300000
dd = [rdecks(int(s)) for s in seeds]
UA20 = [i for i, c in enumerate(CARDS) if c["name"] == "UA20"][0]
for i in range(1, NR, 2):
    if UA20 not in dd[i][0][1] and CARDS[UA20]["faction"] in (0, dd[i][1][1]):
        dd[i][0][1][0] = UA20
decks = torch.tensor([d for d, _f in dd], dtype=torch.uint8, device=dev); fac = torch.tensor([f for _d, f in dd], dtype=torch.uint8, device=dev)
st = eng.reset(torch.from_numpy(seeds).to(dev), decks, fac)
res, steps = eng.rollout_heuristic(st, torch.from_numpy(w1[:NR]).to(dev), torch.from_numpy(w2[:NR]).to(dev), max_steps=400)
g_res, g_steps, g_host = res.cpu().numpy(), steps.cpu().numpy(), st.cpu().numpy()
bad = tot = flagged = 0
t0 = time.time()
for i in range(NR):
    d, f = dd[i]
    s = oracle.new_game(int(seeds[i]), d[0], d[1], f[0], f[1])
    r, acts = oracle.play_heuristic(s, w1[i], w2[i], 400)
    tot += len(acts)
    flagged += r == -2
    if r != int(g_res[i]) or len(acts) != int(g_steps[i]) or s.tobytes() != g_host[i].tobytes():
        bad += 1
        if bad < 5:
            print("MISMATCH heuristic random decks seed", int(seeds[i]), r, int(g_res[i]), len(acts), int(g_steps[i]))
print("heuristic agents, random decks: %d games, %d env steps, %d stopped by an engine status in both, %d mismatches (oracle %.0f s)" % (NR, tot, flagged, bad, time.time() - t0), flush=True)
