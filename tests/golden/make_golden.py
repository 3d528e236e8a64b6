#!/usr/bin/env python3
"""Generate the committed golden fixtures FROM THE LIVE REFERENCE (build container only).

  python tests/golden/make_golden.py [--procs 8] [--only tapes,chain,rand,heur]

Every fixture is produced by the UNMODIFIED reference engine (/root/reference) driven through
oracle/ref_harness.py (injected Philox stream, uniform-random agent stream or the reference's own
HeuristicAgent), then serialised with pack_reference.  Files:

  default_tapes.npz     48 default-deck games, full per-step record (actions, legal masks, state digests,
                        initial and final packed states)
  default_chain_10k.npz seeds 0..9999, default decks: steps, chained per-step digest, final digest, flags
  randdeck_chain.npz    3000 random 12-card faction decks (generate_random_deck semantics; UP01-03 and S203
                        excluded): decks, factions, steps, chain, final digest, outcome kind
  heuristic_decisions.npz  states sampled from reference HeuristicAgent-vs-HeuristicAgent games with the
                        reference's per-action scores, legal set and chosen action, plus whole-game results
  deck_generation.npz   decks returned by the reference's DeckEvolutionConfig.get_deck_configuration /
                        generate_random_deck (utils.py) drawing from the injected per-game stream: three schedules x
                        every generation x 24 seeds, plus 400 fully random faction decks
  es_operators.npz      the reference's Population.generate_offspring / select_from_combined and
                        WeightVector.mutate driven with the per-row streams (oracle/ref_harness.EsStream): three
                        scenarios (plain, sigma reset, diversity injection) x 4 generations
  ref_population.pkl    a checkpoint written by the reference's Population.save_population
  ref_training_log.csv  two rows written by the reference's EvolutionEngine._save_generation_log
  heuristic_games.npz   BASELINE config 1 at scale: 1,024 whole reference games HeuristicAgent vs HeuristicAgent (seeds 5000.., weights
                        RandomState(1000+s) / (2000+s)): all actions, winner, final digest
  heuristic_vs_expert.npz  reference games HeuristicAgent vs Stormbound.expert_action (the match of play_vs_expert.py:65-94
                        with the intended loop): 16 games agent FIRST + 16 games agent SECOND, all actions, result, final digest
  expert_tapes.npz      both seats play the reference's Stormbound.expert_action (it draws from the game's stream):
                        160 default-deck + 240 random-deck games, per-step actions and state digests
"""
import argparse
import multiprocessing as mp
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "oracle"), ROOT]

FNV_PRIME = 0x100000001B3
M64 = 0xFFFFFFFFFFFFFFFF


def chain_of(digests):
    ch = 0
    for d in digests:
        ch = ((ch ^ int(d)) * FNV_PRIME) & M64
    return ch


def random_decks(seed, exclude=("UP01", "UP02", "UP03", "S203")):
    import random
    import ref_harness as h
    r = h.ref()
    rng = random.Random(seed)
    decks, factions = [], []
    for _ in range(2):
        f = rng.choice([1, 2, 3, 4])
        pool = [c["name"] for c in r.table[1:113] if c["faction"] in (0, f) and c["name"] not in exclude]
        decks.append(rng.sample(pool, 12))
        factions.append(f)
    return decks, factions


def work_chain(seed):
    import ref_harness as h
    t = h.play_random_game(seed, record=False)
    fin = 0 if t["final"] is None else h.fnv1a64(t["final"].tobytes())
    return seed, t["n_steps"], chain_of(t["digests"]), fin, t["err"], int(t["done"])


def work_rand(seed):
    import ref_harness as h
    decks, factions = random_decks(seed)
    r = h.ref()
    t = h.play_random_game(seed, decks, factions, record=False)
    fin = 0 if t["final"] is None else h.fnv1a64(t["final"].tobytes())
    return (seed, [[r.index[n] for n in d] for d in decks], factions, t["n_steps"], chain_of(t["digests"]), fin, t["err"],
            int(t["done"]), int(t["actions"][-1]) if len(t["actions"]) else 255)


def work_tape(seed):
    import ref_harness as h
    t = h.play_random_game(seed, record=True)
    return seed, t["init"], t["final"], t["actions"], t["masks"], t["digests"]


def work_heur(seed):
    """One reference game HeuristicAgent vs HeuristicAgent through StormboundAdapter (intended loop, Q15 bypassed)."""
    import ref_harness as h
    r = h.ref()
    os.chdir(h.REF)
    from evo.game_adapter import StormboundAdapter
    from evo.heuristic_agent import HeuristicAgent
    from evo.weights import WeightVector
    w1 = np.random.RandomState(1000 + seed).uniform(0, 1, 10)
    w2 = np.random.RandomState(2000 + seed).uniform(0, 1, 10)
    wv1, wv2 = WeightVector(10), WeightVector(10)
    wv1.weights, wv2.weights = w1.copy(), w2.copy()
    agents = [HeuristicAgent(wv1, 0), HeuristicAgent(wv2, 1)]
    game = h.make_game(seed)
    adapter = StormboundAdapter(game)
    samples, actions = [], []
    steps = 0
    with h.quiet():
        while not adapter.game.env.have_winner() and steps < 400:
            cur = adapter.get_current_player()
            agent = agents[cur]
            legal = adapter.get_legal_actions()
            if steps % 7 == 3:  # sample: state, weights, per-action reference scores
                st = h.pack_reference(adapter.game, steps=steps, done=0)
                scores = np.full(156, np.nan)
                for a in legal:
                    scores[a] = agent.score_action(adapter, a)
                samples.append((st, (w1, w2)[cur].copy(), h.legal_mask(legal), scores))
            a = agent.select_action(adapter)
            if samples and steps % 7 == 3:
                samples[-1] = samples[-1] + (a,)
            adapter = adapter.apply_action(a)
            actions.append(a)
            steps += 1
    b = adapter.game.env.board
    first = b.local if int(b.local.order) == 0 else b.remote
    second = b.remote if int(b.local.order) == 0 else b.local
    result = 0 if (second.strength < 0 and first.strength >= 0) else 1 if (first.strength < 0 and second.strength >= 0) else -1
    final = h.pack_reference(adapter.game, steps=steps, done=0)
    return seed, w1, w2, np.array(actions, dtype=np.uint8), result, h.fnv1a64(final.tobytes()), samples


def work_heur_game(seed):
    """work_heur without the per-decision samples (whole-game record only)."""
    import ref_harness as h
    h.ref()
    os.chdir(h.REF)
    from evo.game_adapter import StormboundAdapter
    from evo.heuristic_agent import HeuristicAgent
    from evo.weights import WeightVector
    w1 = np.random.RandomState(1000 + seed).uniform(0, 1, 10)
    w2 = np.random.RandomState(2000 + seed).uniform(0, 1, 10)
    wv1, wv2 = WeightVector(10), WeightVector(10)
    wv1.weights, wv2.weights = w1.copy(), w2.copy()
    agents = [HeuristicAgent(wv1, 0), HeuristicAgent(wv2, 1)]
    adapter = StormboundAdapter(h.make_game(seed))
    actions, steps, err = [], 0, 0
    with h.quiet():
        while not adapter.game.env.have_winner() and steps < 400:
            try:
                a = agents[adapter.get_current_player()].select_action(adapter)
                adapter = adapter.apply_action(a)
            except Exception:  # noqa: BLE001
                err = 1
                break
            actions.append(a)
            steps += 1
    b = adapter.game.env.board
    first = b.local if int(b.local.order) == 0 else b.remote
    second = b.remote if int(b.local.order) == 0 else b.local
    result = -2 if err else 0 if (second.strength < 0 and first.strength >= 0) else 1 if (first.strength < 0 and second.strength >= 0) else -1
    final = 0 if err else h.fnv1a64(h.pack_reference(adapter.game, steps=steps, done=0).tobytes())
    return seed, w1, w2, np.array(actions, dtype=np.uint8), result, final


def work_hve(job):
    """One reference game: HeuristicAgent on `seat`, Stormbound.expert_action on the other seat."""
    seed, seat = job
    import ref_harness as h
    h.ref()
    os.chdir(h.REF)
    from evo.game_adapter import StormboundAdapter
    from evo.heuristic_agent import HeuristicAgent
    from evo.weights import WeightVector
    w = np.random.RandomState(3000 + seed).uniform(0, 1, 10)
    wv = WeightVector(10)
    wv.weights = w.copy()
    agent = HeuristicAgent(wv, seat)
    adapter = StormboundAdapter(h.make_game(seed))
    actions, steps, err = [], 0, 0
    with h.quiet():
        while not adapter.game.env.have_winner() and steps < 400:
            try:
                if adapter.get_current_player() == seat:
                    a = agent.select_action(adapter)
                else:
                    a = adapter.game.env.expert_action()
                adapter = adapter.apply_action(a)
            except Exception:  # noqa: BLE001 -- the reference raised inside expert_action / step
                err = 1
                break
            actions.append(a)
            steps += 1
    b = adapter.game.env.board
    first = b.local if int(b.local.order) == 0 else b.remote
    second = b.remote if int(b.local.order) == 0 else b.local
    result = -2 if err else 0 if (second.strength < 0 and first.strength >= 0) else 1 if (first.strength < 0 and second.strength >= 0) else -1
    final = 0 if err else h.fnv1a64(h.pack_reference(adapter.game, steps=steps, done=0).tobytes())
    return seed, seat, w, np.array(actions, dtype=np.uint8), result, final


def work_expert(seed):
    import ref_harness as h
    r = h.ref()
    if seed < 200000:
        decks, factions = h.DEFAULT_DECKS, h.DEFAULT_FACTIONS
    else:
        decks, factions = random_decks(seed)
    t = h.play_expert_game(seed, decks, factions, record=False)
    return (seed, [[r.index[n] for n in d] for d in decks], list(factions), np.frombuffer(t["init"].tobytes(), dtype=np.uint8),
            t["actions"], t["digests"], t["n_steps"], t["err"], int(t["done"]))


DECK_SCHEDULES = (dict(), dict(exploit_generations=2, explore_generations=13, max_random_ratio=1.0, balance_archetype_ratio=0.4),
                  dict(exploit_generations=0, explore_generations=7, max_random_ratio=0.8))


def make_decks():
    import ref_harness as h
    from monsoon_b200.evo import DeckEvolutionConfig as Mirror
    r = h.ref()
    import utils
    from enums import Faction
    rows = []  # seed, generation, mode, n_preserve, q, factions[2], archetypes[24], decks[24]
    a1, a2 = h.DEFAULT_DECKS
    for kw in DECK_SCHEDULES:
        ref_cfg = utils.DeckEvolutionConfig([getattr(r.cards, n)() for n in a1], [getattr(r.cards, n)() for n in a2], **kw)
        mir = Mirror(a1, a2, **kw)
        for gen in range(ref_cfg.exploit_generations + ref_cfg.explore_generations + 3):
            assert ref_cfg.get_phase_info(gen) == mir.get_phase_info(gen)
            mode, k, q = mir.phase_parameters(gen)
            for j in range(24):
                seed = j * 7919 + gen
                d = h.reference_decks(seed, gen, ref_cfg)
                rows.append((seed, gen, mode, k, q, [mir.player1_faction, mir.player2_faction],
                             mir.player1_archetype + mir.player2_archetype, d[0] + d[1]))
    for seed in range(400):
        f = [1 + seed % 4, 1 + (seed // 4) % 4]
        saved, utils.random = utils.random, h.PhiloxPyRandom(seed, 5)
        try:
            d = [[r.index[type(c).__name__] for c in utils.generate_random_deck(Faction(x))] for x in f]
        finally:
            utils.random = saved
        rows.append((seed, 5, 3, 0, 0.0, f, [0] * 24, d[0] + d[1]))
    np.savez_compressed(os.path.join(HERE, "deck_generation.npz"),
                        seeds=np.array([x[0] for x in rows], dtype=np.uint64), generation=np.array([x[1] for x in rows], dtype=np.uint32),
                        mode=np.array([x[2] for x in rows], dtype=np.uint8), n_preserve=np.array([x[3] for x in rows], dtype=np.uint8),
                        q=np.array([x[4] for x in rows], dtype=np.float64), factions=np.array([x[5] for x in rows], dtype=np.uint8),
                        archetypes=np.array([x[6] for x in rows], dtype=np.uint8), decks=np.array([x[7] for x in rows], dtype=np.uint8))
    print("deck_generation.npz", len(rows))


ES_CFG = dict(mu=12, lambda_=20, tau=0.1, tau_prime=0.01, min_sigma=1e-5, initial_sigma=0.1)


def make_es():
    import contextlib
    import io
    import ref_harness as h
    import validate_vs_reference as v
    out = {}
    for si, scenario in enumerate(("normal", "reset", "inject")):
        seed = 21 + si
        pop, stream, WV = h.reference_population(ES_CFG, seed)
        rs = np.random.RandomState(seed)
        mu, lam = ES_CFG["mu"], ES_CFG["lambda_"]
        w = rs.uniform(0, 1, (mu, 10))
        s = rs.uniform(0.05, 0.2, (mu, 10)) if scenario != "reset" else rs.uniform(1e-5, 5e-5, (mu, 10))
        pop.individuals = []
        for i in range(mu):
            x = WV(10)
            x.set_weights(w[i].copy())
            x.set_sigmas(s[i].copy())
            pop.individuals.append(x)
        pop.fitness_scores, pop.generation = [0.0] * mu, 1
        out["%s_seed" % scenario] = np.int64(seed)
        out["%s_w0" % scenario], out["%s_s0" % scenario] = w, s
        fits, ow, os_, sw, ss, sf = [], [], [], [], [], []
        for g in range(4):
            vals = np.full(mu + lam, 0.5) if (scenario == "inject" and g % 2 == 1) else np.round(np.random.RandomState(100 * seed + g).uniform(0, 1, mu + lam), 1)
            (cw, cs), _f = v.es_reference_generation(pop, stream, lambda n, vals=vals: vals[:n])
            fits.append(vals)
            ow.append(cw)
            os_.append(cs)
            sw.append(np.array([c.weights for c in pop.individuals]))
            ss.append(np.array([c.sigmas for c in pop.individuals]))
            sf.append(np.array(pop.fitness_scores))
        for k, a in (("fitness", fits), ("off_w", ow), ("off_s", os_), ("sur_w", sw), ("sur_s", ss), ("sur_f", sf)):
            out["%s_%s" % (scenario, k)] = np.stack(a)
        if scenario == "normal":  # f4: files written by the reference itself
            with contextlib.redirect_stdout(io.StringIO()):
                pop.save_population(os.path.join(HERE, "ref_population.pkl"))
                import evo.evolution as re_
                eng = re_.EvolutionEngine.__new__(re_.EvolutionEngine)
                eng.results_dir = HERE
                log = os.path.join(HERE, "training_log.csv")
                if os.path.exists(log):
                    os.remove(log)
                st = pop.get_population_stats()
                eng._save_generation_log(st, {"games_per_second": 1234.56}, 7.891)
                eng._save_generation_log(dict(st, generation=st["generation"] + 1, best_fitness=0.75), {"games_per_second": 99.0}, 0.004)
                os.replace(log, os.path.join(HERE, "ref_training_log.csv"))
            out["log_stats"] = np.array([st[k] for k in ("generation", "best_fitness", "mean_fitness", "std_fitness", "diversity", "avg_mutation_strength")])
    np.savez_compressed(os.path.join(HERE, "es_operators.npz"), **out)
    print("es_operators.npz", len(out), "arrays; ref_population.pkl; ref_training_log.csv")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--procs", type=int, default=os.cpu_count())
    ap.add_argument("--only", default="tapes,chain,rand,heur,expert,decks,es,hve,heurgames")
    ap.add_argument("--n-chain", type=int, default=10000)
    ap.add_argument("--n-rand", type=int, default=3000)
    ap.add_argument("--n-heur", type=int, default=24)
    ap.add_argument("--n-heur-games", type=int, default=1024)
    args = ap.parse_args()
    only = set(args.only.split(","))
    pool = mp.Pool(args.procs)
    if "tapes" in only:
        res = pool.map(work_tape, range(48))
        np.savez_compressed(os.path.join(HERE, "default_tapes.npz"),
                            seeds=np.array([r[0] for r in res], dtype=np.uint64),
                            init=np.stack([np.frombuffer(r[1].tobytes(), dtype=np.uint8) for r in res]),
                            final=np.stack([np.frombuffer(r[2].tobytes(), dtype=np.uint8) for r in res]),
                            lengths=np.array([len(r[3]) for r in res], dtype=np.int32),
                            actions=np.concatenate([r[3] for r in res]),
                            masks=np.concatenate([r[4] for r in res]),
                            digests=np.concatenate([r[5] for r in res]))
        print("default_tapes.npz", len(res))
    if "chain" in only:
        res = pool.map(work_chain, range(args.n_chain), chunksize=16)
        a = np.array(res, dtype=object)
        np.savez_compressed(os.path.join(HERE, "default_chain_10k.npz"),
                            seeds=np.array([r[0] for r in res], dtype=np.uint64), steps=np.array([r[1] for r in res], dtype=np.int32),
                            chain=np.array([r[2] for r in res], dtype=np.uint64), final=np.array([r[3] for r in res], dtype=np.uint64),
                            err=np.array([r[4] for r in res], dtype=np.uint8), done=np.array([r[5] for r in res], dtype=np.uint8))
        print("default_chain_10k.npz", len(res), "errs", sum(1 for r in res if r[4]))
    if "rand" in only:
        res = pool.map(work_rand, range(100000, 100000 + args.n_rand), chunksize=8)
        np.savez_compressed(os.path.join(HERE, "randdeck_chain.npz"),
                            seeds=np.array([r[0] for r in res], dtype=np.uint64), decks=np.array([r[1] for r in res], dtype=np.uint8),
                            factions=np.array([r[2] for r in res], dtype=np.uint8), steps=np.array([r[3] for r in res], dtype=np.int32),
                            chain=np.array([r[4] for r in res], dtype=np.uint64), final=np.array([r[5] for r in res], dtype=np.uint64),
                            err=np.array([r[6] for r in res], dtype=np.uint8), done=np.array([r[7] for r in res], dtype=np.uint8),
                            last_action=np.array([r[8] for r in res], dtype=np.uint8))
        print("randdeck_chain.npz", len(res), "ref exceptions", sum(1 for r in res if r[6] == 1), "overflow", sum(1 for r in res if r[6] == 2))
    if "heur" in only:
        res = pool.map(work_heur, range(args.n_heur), chunksize=1)
        samples = [s for r in res for s in r[6] if len(s) == 5]
        np.savez_compressed(os.path.join(HERE, "heuristic_decisions.npz"),
                            game_seeds=np.array([r[0] for r in res], dtype=np.uint64),
                            w_first=np.stack([r[1] for r in res]), w_second=np.stack([r[2] for r in res]),
                            game_lengths=np.array([len(r[3]) for r in res], dtype=np.int32),
                            game_actions=np.concatenate([r[3] for r in res]),
                            game_result=np.array([r[4] for r in res], dtype=np.int8),
                            game_final=np.array([r[5] for r in res], dtype=np.uint64),
                            states=np.stack([np.frombuffer(s[0].tobytes(), dtype=np.uint8) for s in samples]),
                            weights=np.stack([s[1] for s in samples]), masks=np.stack([s[2] for s in samples]),
                            scores=np.stack([s[3] for s in samples]), chosen=np.array([s[4] for s in samples], dtype=np.uint8))
        print("heuristic_decisions.npz games", len(res), "samples", len(samples))
    if "heurgames" in only:
        res = pool.map(work_heur_game, range(5000, 5000 + args.n_heur_games), chunksize=2)
        np.savez_compressed(os.path.join(HERE, "heuristic_games.npz"),
                            seeds=np.array([r[0] for r in res], dtype=np.uint64), w_first=np.stack([r[1] for r in res]),
                            w_second=np.stack([r[2] for r in res]), lengths=np.array([len(r[3]) for r in res], dtype=np.int32),
                            actions=np.concatenate([r[3] for r in res]), result=np.array([r[4] for r in res], dtype=np.int8),
                            final=np.array([r[5] for r in res], dtype=np.uint64))
        print("heuristic_games.npz", len(res), "aborted", sum(1 for r in res if r[4] == -2))
    if "hve" in only:
        res = pool.map(work_hve, [(400 + i, i % 2) for i in range(32)], chunksize=1)
        np.savez_compressed(os.path.join(HERE, "heuristic_vs_expert.npz"),
                            seeds=np.array([r[0] for r in res], dtype=np.uint64), seat=np.array([r[1] for r in res], dtype=np.uint8),
                            weights=np.stack([r[2] for r in res]), lengths=np.array([len(r[3]) for r in res], dtype=np.int32),
                            actions=np.concatenate([r[3] for r in res]), result=np.array([r[4] for r in res], dtype=np.int8),
                            final=np.array([r[5] for r in res], dtype=np.uint64))
        print("heuristic_vs_expert.npz", len(res), "results", [int(r[4]) for r in res])
    if "decks" in only:
        make_decks()
    if "es" in only:
        make_es()
    if "expert" in only:
        res = pool.map(work_expert, list(range(160)) + list(range(200000, 200240)), chunksize=4)
        np.savez_compressed(os.path.join(HERE, "expert_tapes.npz"),
                            seeds=np.array([r[0] for r in res], dtype=np.uint64), decks=np.array([r[1] for r in res], dtype=np.uint8),
                            factions=np.array([r[2] for r in res], dtype=np.uint8), init=np.stack([r[3] for r in res]),
                            lengths=np.array([len(r[4]) for r in res], dtype=np.int32), actions=np.concatenate([r[4] for r in res]),
                            steps=np.array([r[6] for r in res], dtype=np.int32), digests=np.concatenate([r[5] for r in res]),
                            err=np.array([r[7] for r in res], dtype=np.uint8), done=np.array([r[8] for r in res], dtype=np.uint8))
        print("expert_tapes.npz", len(res), "ref exceptions", sum(1 for r in res if r[7] in (1, 3)), "overflow", sum(1 for r in res if r[7] == 2))


if __name__ == "__main__":
    main()
