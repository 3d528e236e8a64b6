"""CPU: pins the C oracle against the committed golden fixtures that the UNMODIFIED reference produced
(tests/golden/make_golden.py).  This is what makes the oracle a trustworthy checker for the -m gpu tests."""
import os

import numpy as np
import pytest

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FNV_PRIME, M64 = 0x100000001B3, 0xFFFFFFFFFFFFFFFF


def chain_of(digests):
    ch = 0
    for d in digests:
        ch = ((ch ^ int(d)) * FNV_PRIME) & M64
    return ch


def final_digest(oracle, st):
    """the fixtures' `final` state carries the done bit only (ref_harness.play_random_game)"""
    f = st.copy()
    f[19] &= 1
    return oracle.digest(f)


def load(name):
    p = os.path.join(G, name)
    if not os.path.exists(p):
        pytest.skip("fixture %s not generated" % name)
    return np.load(p)


def default_decks():
    from monsoon_b200.engine import DEFAULT_DECKS, deck_indices
    return [deck_indices(d) for d in DEFAULT_DECKS]


def test_full_tapes_every_step(oracle):
    z = load("default_tapes.npz")
    d0, d1 = default_decks()
    off = 0
    for g, seed in enumerate(z["seeds"]):
        n = int(z["lengths"][g])
        st = oracle.new_game(int(seed), d0, d1, 3, 2)
        assert st.tobytes() == z["init"][g].tobytes()
        for k in range(n):
            assert np.array_equal(oracle.legal_mask(st), z["masks"][off + k]), (g, k)
            oracle.step(st, int(z["actions"][off + k]))
            assert oracle.digest(st) == int(z["digests"][off + k]), (g, k)
        fin = st.copy()
        fin[19] &= 1
        assert fin.tobytes() == z["final"][g].tobytes()
        off += n


def test_10k_default_games_bit_exact(oracle):
    """BASELINE target: trajectories, final boards and winners bit-exact on 10k seeded games."""
    z = load("default_chain_10k.npz")
    d0, d1 = default_decks()
    bad = []
    for seed, steps, chain, final, err in zip(z["seeds"], z["steps"], z["chain"], z["final"], z["err"]):
        st = oracle.new_game(int(seed), d0, d1, 3, 2)
        _a, dig, _m = oracle.rollout_random(st, 400)
        if err == 0:
            ok = len(dig) == steps and chain_of(dig) == int(chain) and final_digest(oracle, st) == int(final) and st[18] == 0
        else:  # the reference raised inside step `steps`: the oracle must flag the same step
            ok = st[18] != 0 and len(dig) == steps + 1 and chain_of(dig[:-1]) == int(chain)
        if not ok:
            bad.append(int(seed))
    assert not bad, bad[:10]


def test_random_deck_games(oracle):
    """All implemented cards (minus UP01-03 / S203) in faction decks; the only tolerated
    difference is a game the engine flags SB_ERR_UNSUPPORTED (nested Temple-of-Time memories)."""
    z = load("randdeck_chain.npz")
    bad, unsupported, flagged = [], 0, 0
    for i in range(len(z["seeds"])):
        d, f = z["decks"][i], z["factions"][i]
        st = oracle.new_game(int(z["seeds"][i]), d[0], d[1], int(f[0]), int(f[1]))
        _a, dig, _m = oracle.rollout_random(st, 400)
        kind = int(z["err"][i])
        if st[18] == 5 or (st[18] == 6 and (kind != 2 or len(dig) <= z["steps"][i])):
            unsupported += 1  # documented capacity / modelling limits; capacity may trip a few steps early
            continue
        if kind == 0:
            ok = len(dig) == z["steps"][i] and chain_of(dig) == int(z["chain"][i]) and final_digest(oracle, st) == int(z["final"][i])
        else:
            flagged += 1
            ok = st[18] != 0 and len(dig) == z["steps"][i] + 1 and chain_of(dig[:-1]) == int(z["chain"][i])
            if kind == 2:
                ok = ok and st[18] == 6
        if not ok:
            bad.append(int(z["seeds"][i]))
    assert not bad, bad[:10]
    assert unsupported <= len(z["seeds"]) * 3 // 200, unsupported  # <= 1.5 %: nested Temple-of-Time memories + ext capacity


def test_card_focus_games(oracle):
    """Every one of the 112 cards in games whose two decks both hold it (tests/golden/make_golden_r2.py): pins the rare
    branches, S203 (first-occurrence dedupe order, Q14) and UP01-03 (Q12 observation bypass) to the reference."""
    z = load("card_focus.npz")
    assert set(z["card"].tolist()) == set(range(1, 113))
    bad, unsupported = [], 0
    for i in range(len(z["seeds"])):
        d, f = z["decks"][i], z["factions"][i]
        st = oracle.new_game(int(z["seeds"][i]), d[0], d[1], int(f[0]), int(f[1]))
        _a, dig, _m = oracle.rollout_random(st, 400)
        kind = int(z["err"][i])
        if st[18] == 5 or (st[18] == 6 and (kind != 2 or len(dig) <= z["steps"][i])):
            unsupported += 1
            continue
        if kind == 0:
            ok = len(dig) == z["steps"][i] and chain_of(dig) == int(z["chain"][i]) and final_digest(oracle, st) == int(z["final"][i])
        else:
            ok = st[18] != 0 and len(dig) == z["steps"][i] + 1 and chain_of(dig[:-1]) == int(z["chain"][i])
            if kind == 2:
                ok = ok and st[18] == 6
        if not ok:
            bad.append((int(z["seeds"][i]), int(z["card"][i])))
    assert not bad, bad[:10]
    assert unsupported <= len(z["seeds"]) * 3 // 200, unsupported


def test_observation_and_features_vs_reference(oracle):
    """SURVEY rows a11 / a12 pinned: Stormbound.get_observation (games/stormbound.py:400-526, ids card.py:25-46) and
    StateFeatures.get_feature_vector (evo/features.py:327-342) of the reference on 3,000+ sampled states (default decks,
    random decks, heuristic games; both flip parities) -- every int32 of the 27x5x4 observation and every feature bit."""
    z = load("obs_features.npz")
    assert len(z["states"]) >= 2000 and set(np.unique(z["states"][:, 14]).tolist()) == {0, 1}
    for i in range(len(z["states"])):
        st = z["states"][i].copy()
        obs, err = oracle.observe(st)
        assert err == 0 and np.array_equal(obs, z["obs"][i]), i
        f, err = oracle.features(st)
        assert err == 0 and np.array_equal(f, z["feat"][i]), i


def test_card_coverage_table():
    """every card class with an ability fired at least 20 times in the reference games behind the fixtures"""
    import json
    p = os.path.join(G, "card_coverage.json")
    if not os.path.exists(p):
        pytest.skip("coverage table not generated")
    total = json.load(open(p))["total"]
    assert len(total) >= 83 and min(total.values()) >= 20, {k: v for k, v in total.items() if v < 20}


def test_expert_games(oracle):
    """Both seats play Stormbound.expert_action (games/stormbound.py:563-637): every action and every per-step
    state digest of the reference's tapes, default and random decks."""
    z = load("expert_tapes.npz")
    aoff = doff = 0  # `actions` holds one more entry than `digests` for a game the reference aborted inside step
    bad, tolerated = [], 0
    for i in range(len(z["seeds"])):
        d, f, steps, kind = z["decks"][i], z["factions"][i], int(z["steps"][i]), int(z["err"][i])
        st = oracle.new_game(int(z["seeds"][i]), d[0], d[1], int(f[0]), int(f[1]))
        assert st.tobytes() == z["init"][i].tobytes()
        ok = True
        for k in range(steps):
            a = oracle.expert_action(st)
            oracle.step(st, a)
            if st[18] in (5, 6):  # documented capacity / modelling limits
                break
            if a != z["actions"][aoff + k] or oracle.digest(st) != int(z["digests"][doff + k]):
                ok = False
                break
        if st[18] in (5, 6):
            tolerated += 1
        elif ok and kind:  # the reference raised (1: inside step, 3: inside expert_action, 2: no longer packable)
            a = oracle.expert_action(st)
            if kind != 3:
                ok = a == z["actions"][aoff + steps]
                oracle.step(st, a)
            ok = ok and st[18] != 0
        if not ok:
            bad.append(int(z["seeds"][i]))
        aoff += int(z["lengths"][i])
        doff += steps
    assert not bad, bad[:10]
    assert tolerated <= len(z["seeds"]) * 3 // 200, tolerated


def test_deck_generation(oracle):
    """utils.py generate_random_deck / DeckEvolutionConfig decks recorded from the reference (injected stream)."""
    z = load("deck_generation.npz")
    for i in range(len(z["seeds"])):
        got = oracle.generate_decks(int(z["seeds"][i]), int(z["generation"][i]), int(z["mode"][i]), int(z["n_preserve"][i]),
                                    float(z["q"][i]), z["archetypes"][i], z["factions"][i])
        assert np.array_equal(got.reshape(24), z["decks"][i]), i


ES_CFG = dict(mu=12, lambda_=20, tau=0.1, tau_prime=0.01, min_sigma=1e-5, initial_sigma=0.1)


def es_generation(oracle, seed, generation, w, s, fitness, cfg=ES_CFG):
    """One (mu + lambda) generation with the oracle's operators and the conditions of evo/population.py:92-176."""
    mu, lam = cfg["mu"], cfg["lambda_"]
    W = np.concatenate([w, np.zeros((lam, w.shape[1]))])
    S = np.concatenate([s, np.zeros((lam, w.shape[1]))])
    oracle.es_offspring(seed, generation, mu, lam, cfg["tau"], cfg["tau_prime"], cfg["min_sigma"], W, S)
    w2, s2, f2, _order = oracle.es_select(mu, fitness, W, S)
    if np.mean(s2) < cfg["min_sigma"] * 10:
        oracle.es_reset_sigmas(seed, generation + 1, cfg["initial_sigma"], s2)
    if np.std(f2) == 0.0 and len(set(f2.tolist())) == 1:
        oracle.es_inject_diversity(seed, generation + 1, cfg["tau"], cfg["tau_prime"], cfg["min_sigma"], cfg["initial_sigma"], w2, s2)
    return W[mu:], S[mu:], w2, s2, f2


def test_es_operators(oracle):
    """generate_offspring / mutate / select_from_combined recorded from the reference (per-row streams injected):
    selection order exact, weights and sigmas within 1e-12 (the reference uses numpy's exp)."""
    z = load("es_operators.npz")
    for scenario in ("normal", "reset", "inject"):
        seed = int(z[scenario + "_seed"])
        w, s = z[scenario + "_w0"].copy(), z[scenario + "_s0"].copy()
        for g in range(z[scenario + "_fitness"].shape[0]):
            ow, os_, w2, s2, f2 = es_generation(oracle, seed, 1 + g, w, s, z[scenario + "_fitness"][g])
            assert np.allclose(ow, z[scenario + "_off_w"][g], rtol=0, atol=1e-12), (scenario, g)
            assert np.allclose(os_, z[scenario + "_off_s"][g], rtol=1e-12, atol=0), (scenario, g)
            assert np.array_equal(f2, z[scenario + "_sur_f"][g]), (scenario, g)
            assert np.allclose(w2, z[scenario + "_sur_w"][g], rtol=0, atol=1e-12), (scenario, g)
            assert np.allclose(s2, z[scenario + "_sur_s"][g], rtol=1e-12, atol=0), (scenario, g)
            w, s = z[scenario + "_sur_w"][g].copy(), z[scenario + "_sur_s"][g].copy()


def test_heuristic_scores_and_choices(oracle):
    """Float action scores within 1e-5 relative of the reference's; identical choices where the top-two gap
    exceeds that tolerance (BASELINE north star)."""
    z = load("heuristic_decisions.npz")
    n_cmp = 0
    for i in range(len(z["states"])):
        a, s, m = oracle.select_action(z["states"][i].copy(), z["weights"][i])
        ref = z["scores"][i]
        legal = ~np.isnan(ref)
        assert np.array_equal(m, z["masks"][i]) and np.array_equal(legal, ~np.isnan(s))
        np.testing.assert_allclose(s[legal], ref[legal], rtol=1e-5, atol=1e-9)
        top = np.sort(ref[legal])[::-1]
        if len(top) == 1 or top[0] - top[1] > 1e-5 * max(1.0, abs(top[0])):
            assert a == int(z["chosen"][i]), i
            n_cmp += 1
    assert n_cmp > len(z["states"]) // 2


def test_heuristic_whole_games(oracle):
    z = load("heuristic_decisions.npz")
    d0, d1 = default_decks()
    off, same = 0, 0
    for g, seed in enumerate(z["game_seeds"]):
        n = int(z["game_lengths"][g])
        st = oracle.new_game(int(seed), d0, d1, 3, 2)
        r, acts = oracle.play_heuristic(st, z["w_first"][g], z["w_second"][g], 400)
        if len(acts) == n and np.array_equal(acts, z["game_actions"][off:off + n]):
            fin = st.copy()
            fin[19] = 0  # the fixture's final state was packed without the done/reward byte
            assert r == int(z["game_result"][g]) and oracle.digest(fin) == int(z["game_final"][g])
            same += 1
        off += n
    # a near-tie broken differently by BLAS summation order would fork a game; none is expected at 1e-5
    assert same == len(z["game_seeds"])


def test_heuristic_vs_expert_games(oracle):
    """HeuristicAgent against Stormbound.expert_action (play_vs_expert.py:65-94, intended loop), both seatings:
    every action, the winner and the final state of the reference's games."""
    z = load("heuristic_vs_expert.npz")
    d0, d1 = default_decks()
    off = 0
    for i in range(len(z["seeds"])):
        n, seat = int(z["lengths"][i]), int(z["seat"][i])
        st = oracle.new_game(int(z["seeds"][i]), d0, d1, 3, 2)
        w = z["weights"][i]
        r, acts = oracle.play_heuristic(st, w if seat == 0 else None, w if seat == 1 else None, 400)
        assert r == int(z["result"][i]), i
        if r != -2:
            assert np.array_equal(acts, z["actions"][off:off + n]), i
            fin = st.copy()
            fin[19] = 0  # the fixture's final state was packed without the done/reward byte
            assert oracle.digest(fin) == int(z["final"][i]), i
        off += n


def test_heuristic_games_at_scale(oracle):
    """BASELINE config 1: 1,024 whole reference games HeuristicAgent vs HeuristicAgent -- every action, winner, final state."""
    z = load("heuristic_games.npz")
    d0, d1 = default_decks()
    off, bad = 0, []
    for i in range(len(z["seeds"])):
        n = int(z["lengths"][i])
        st = oracle.new_game(int(z["seeds"][i]), d0, d1, 3, 2)
        r, acts = oracle.play_heuristic(st, z["w_first"][i], z["w_second"][i], 400)
        ok = r == int(z["result"][i])
        if ok and r != -2:
            fin = st.copy()
            fin[19] = 0
            ok = np.array_equal(acts, z["actions"][off:off + n]) and oracle.digest(fin) == int(z["final"][i])
        if not ok:
            bad.append(int(z["seeds"][i]))
        off += n
    assert not bad, bad[:10]

