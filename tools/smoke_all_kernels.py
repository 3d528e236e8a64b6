"""GPU box (plain run; compute-sanitizer is closed on this pool): one small call of every round-2 kernel (heuristic rollout in the owner / holder structure for
K = 1, 2, 4, 8 with random decks from all cards, device-side schedule, streaming queries, warp-engine rollouts and step)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np, torch
from monsoon_b200.engine import Engine
eng = Engine(0); dev = eng.device
n = 192
seeds = torch.arange(n, dtype=torch.int64, device=dev) + 5
fac = torch.from_numpy(np.random.RandomState(9).randint(1, 5, (n, 2)).astype(np.uint8)).to(dev)
decks, factions = eng.generate_decks(seeds, 0, 3, factions=fac)
w = torch.from_numpy(np.random.RandomState(5).uniform(0, 1, (n, 10))).to(dev)
for pack in (1, 2, 4, 8, 0):
    eng.set_option("heur_pack", pack)
    for d in (None, (decks, factions)):
        st = eng.reset(seeds) if d is None else eng.reset(seeds, *d)
        res, steps = eng.rollout_heuristic(st, w, w, max_steps=400)
        st = eng.reset(seeds) if d is None else eng.reset(seeds, *d)
        res, steps = eng.rollout_heuristic(st, w, None, max_steps=400)   # expert second seat
    print("pack", pack, "ok", int(steps.sum()), flush=True)
eng.set_option("heur_pack", -1)
counts, ab = eng.eval_population("round_robin", 6, w[:8].contiguous(), 3, 7, 1, 0, 6 * 7 * 3, chunk_games=50)
counts, ab = eng.eval_population("versus", 5, w[:8].contiguous(), 3, 7, 1, 2, 5 * 3 * 3 - 1)
print("eval_population ok", counts.sum().item(), flush=True)
for engine in (1, 0):
    eng.set_option("engine", engine)
    st = eng.reset(seeds, decks, factions)
    eng.rollout_random(st, max_steps=25)
    m = eng.legal_mask(st); o = eng.observe(st); f = eng.features(st)
    a = eng.expert_action(st.clone())
    act = torch.full((n,), 155, dtype=torch.uint8, device=dev)
    eng.step(st.clone(), act)
    eng.select_action(st, w)
    st2 = eng.reset(seeds, decks, factions); eng.rollout_random(st2, max_steps=400)
    print("engine", engine, "ok", flush=True)
torch.cuda.synchronize()
print("done")
