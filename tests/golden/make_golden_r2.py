#!/usr/bin/env python3
"""Round-2 golden fixtures FROM THE LIVE REFERENCE (build container only; companion of make_golden.py).

  python tests/golden/make_golden_r2.py [--procs 8] [--only obs,focus,coverage]

  obs_features.npz   states sampled from reference games (default decks + random decks with the uniform-random agent,
                     HeuristicAgent-vs-HeuristicAgent games; both seats, i.e. both flip parities) with
                       obs   = Stormbound.get_observation()            int32[27,5,4]  (games/stormbound.py:400-526, card.py:25-46)
                       feat  = StateFeatures(obs, to_play).get_feature_vector()  f64[10]  (evo/features.py:327-342)
                     beside the packed state -- the reference pin of SURVEY rows a11 / a12.
  card_focus.npz     for every one of the 112 cards, games whose two decks both contain it (the other 11 cards sampled from
                     its faction + NEUTRAL), uniform-random agent: decks, steps, per-step digest chain, final digest, flags.
                     Pins the rare branches the plain random-deck fixture seldom reaches, S203 (with the first-occurrence
                     dedupe order, Q14) and UP01-03 (Q12 observation bypass) included.
  card_coverage.json activate_ability calls per card class over: the games of randdeck_chain.npz, 2,000 games of
                     default_chain_10k.npz, the games of obs_features.npz and card_focus.npz.
"""
import argparse
import collections
import json
import multiprocessing as mp
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "oracle"), ROOT, HERE]

from make_golden import chain_of, random_decks  # noqa: E402

N_FOCUS = 8  # games per card


def _drain_counts(h):
    c = dict(h.ref().activations)
    h.ref().activations.clear()
    return c


def focus_decks(ci, j):
    """Both decks hold card `ci`; the rest is sampled from its faction (a random one for NEUTRAL cards) + NEUTRAL."""
    import random
    import ref_harness as h
    r = h.ref()
    rng = random.Random(1000003 * ci + j)
    name = r.table[ci]["name"]
    decks, factions = [], []
    for _ in range(2):
        f = r.table[ci]["faction"] or rng.choice([1, 2, 3, 4])
        pool = [c["name"] for c in r.table[1:113] if c["faction"] in (0, f) and c["name"] != name]
        decks.append([name] + rng.sample(pool, 11))
        factions.append(f)
    return decks, factions


def work_focus(job):
    ci, j = job
    import ref_harness as h
    r = h.ref()
    seed = 300000 + 64 * ci + j
    decks, factions = focus_decks(ci, j)
    t = h.play_random_game(seed, decks, factions, record=False)
    fin = 0 if t["final"] is None else h.fnv1a64(t["final"].tobytes())
    return (seed, ci, [[r.index[n] for n in d] for d in decks], factions, t["n_steps"], chain_of(t["digests"]), fin, t["err"],
            int(t["done"]), _drain_counts(h))


def work_count_rand(seed):
    import ref_harness as h
    decks, factions = random_decks(seed)
    h.play_random_game(seed, decks, factions, record=False)
    return _drain_counts(h)


def work_count_default(seed):
    import ref_harness as h
    h.play_random_game(seed, record=False)
    return _drain_counts(h)


def _sample(h, game, steps, done, SF):
    env = game.env
    obs = env.get_observation()
    feat = SF(obs, game.to_play()).get_feature_vector()
    st = h.pack_reference(game, steps=steps, done=done)
    return np.frombuffer(st.tobytes(), dtype=np.uint8).copy(), np.asarray(obs, dtype=np.int32).copy(), np.asarray(feat, dtype=np.float64)


def work_obs(job):
    """kind 0: default decks, random agent; 1: random decks, random agent; 2: HeuristicAgent vs HeuristicAgent."""
    seed, kind = job
    import ref_harness as h
    h.ref()
    os.chdir(h.REF)
    from evo.features import StateFeatures as SF
    out = []
    if kind == 2:
        from evo.game_adapter import StormboundAdapter
        from evo.heuristic_agent import HeuristicAgent
        from evo.weights import WeightVector
        wv = []
        for k in (1000, 2000):
            x = WeightVector(10)
            x.weights = np.random.RandomState(k + seed).uniform(0, 1, 10)
            wv.append(x)
        agents = [HeuristicAgent(wv[0], 0), HeuristicAgent(wv[1], 1)]
        adapter = StormboundAdapter(h.make_game(seed))
        steps = 0
        with h.quiet():
            while not adapter.game.env.have_winner() and steps < 400:
                a = agents[adapter.get_current_player()].select_action(adapter)
                adapter = adapter.apply_action(a)
                steps += 1
                if steps % 4 == seed % 4:
                    out.append(_sample(h, adapter.game, steps, 0, SF))
        return kind, out, _drain_counts(h)
    decks, factions = (None, None) if kind == 0 else random_decks(seed)
    game = h.make_game(seed, decks, factions)
    step, done = 0, False
    with h.quiet():
        while not done and step < 400:
            legal = game.legal_actions()
            a = legal[h.agent_pick(seed, step, len(legal))]
            try:
                _o, reward, done = game.step(a)
            except Exception:  # noqa: BLE001 -- the reference raised (Q11): stop sampling this game
                break
            step += 1
            if step % 5 == seed % 5:
                try:
                    out.append(_sample(h, game, step, (1 if done else 0) | (2 if reward else 0), SF))
                except h.PackOverflow:
                    break
    return kind, out, _drain_counts(h)


def add(total, c):
    for k, v in c.items():
        total[k] += v


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--procs", type=int, default=os.cpu_count())
    ap.add_argument("--only", default="obs,focus,coverage")
    args = ap.parse_args()
    only = set(args.only.split(","))
    pool = mp.Pool(args.procs)
    cov = collections.OrderedDict()
    if "obs" in only:
        jobs = [(s, 0) for s in range(20000, 20080)] + [(s, 1) for s in range(120000, 120110)] + [(s, 2) for s in range(7000, 7008)]
        res = pool.map(work_obs, jobs, chunksize=1)
        tot = collections.Counter()
        S, O, F, K = [], [], [], []
        for kind, out, c in res:
            add(tot, c)
            for st, obs, feat in out:
                S.append(st); O.append(obs); F.append(feat); K.append(kind)
        np.savez_compressed(os.path.join(HERE, "obs_features.npz"), states=np.stack(S), obs=np.stack(O), feat=np.stack(F),
                            kind=np.array(K, dtype=np.uint8))
        cov["obs_features"] = dict(tot)
        S = np.stack(S)
        print("obs_features.npz", len(S), "states; by kind", np.bincount(K).tolist(), "flip parities", np.bincount(S[:, 14]).tolist())
    if "focus" in only:
        import ref_harness as h
        n_cards = 112
        jobs = [(ci, j) for ci in range(1, n_cards + 1) for j in range(N_FOCUS)]
        res = pool.map(work_focus, jobs, chunksize=2)
        tot = collections.Counter()
        for r in res:
            add(tot, r[9])
        np.savez_compressed(os.path.join(HERE, "card_focus.npz"),
                            seeds=np.array([r[0] for r in res], dtype=np.uint64), card=np.array([r[1] for r in res], dtype=np.uint8),
                            decks=np.array([r[2] for r in res], dtype=np.uint8), factions=np.array([r[3] for r in res], dtype=np.uint8),
                            steps=np.array([r[4] for r in res], dtype=np.int32), chain=np.array([r[5] for r in res], dtype=np.uint64),
                            final=np.array([r[6] for r in res], dtype=np.uint64), err=np.array([r[7] for r in res], dtype=np.uint8),
                            done=np.array([r[8] for r in res], dtype=np.uint8))
        cov["card_focus"] = dict(tot)
        print("card_focus.npz", len(res), "ref exceptions", sum(1 for r in res if r[7] == 1), "overflow", sum(1 for r in res if r[7] == 2))
    if "coverage" in only:
        tot = collections.Counter()
        for c in pool.map(work_count_rand, range(100000, 103000), chunksize=8):
            add(tot, c)
        cov["randdeck_chain"] = dict(tot)
        tot = collections.Counter()
        for c in pool.map(work_count_default, range(2000), chunksize=8):
            add(tot, c)
        cov["default_chain_first_2000"] = dict(tot)
    path = os.path.join(HERE, "card_coverage.json")
    old = json.load(open(path)) if os.path.exists(path) else {}
    old.update(cov)
    names = sorted({k for v in old.values() if isinstance(v, dict) for k in v})
    old["total"] = {n: sum(v.get(n, 0) for k, v in old.items() if k != "total" and isinstance(v, dict)) for n in names}
    json.dump(old, open(path, "w"), indent=1, sort_keys=True)
    low = {n: c for n, c in old["total"].items() if c < 20}
    print("card_coverage.json: %d card classes with an ability; fewer than 20 activations: %s" % (len(names), low))


if __name__ == "__main__":
    main()
