"""CPU: the warp-per-game engine (monsoon_b200/csrc/sbw_*.cuh) compiled for the HOST (tests/wsim: lane loops instead of
lanes) against the reference fixtures and the oracle.  The same source is what the sm_100a kernels run, so these tests pin
the RULES of the warp engine without a GPU; the -m gpu tests then pin the kernels (scheduling, shared-memory layout)."""
import os

import numpy as np
import pytest

import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_oracle_golden import G, chain_of, default_decks, final_digest, load  # noqa: E402,F401


@pytest.fixture(scope="module")
def wsim():
    import wsim as w
    w.lib()
    return w


def _games(z, oracle):
    for i in range(len(z["seeds"])):
        if "decks" in z.files:
            d, f = z["decks"][i], z["factions"][i]
            yield i, oracle.new_game(int(z["seeds"][i]), d[0], d[1], int(f[0]), int(f[1]))
        else:
            d0, d1 = default_decks()
            yield i, oracle.new_game(int(z["seeds"][i]), d0, d1, 3, 2)


@pytest.mark.parametrize("name,limit", [("default_chain_10k.npz", 1500), ("randdeck_chain.npz", None), ("card_focus.npz", None)])
def test_random_games_vs_reference_fixtures(oracle, wsim, name, limit):
    """same acceptance rule as the oracle's own fixture tests (test_oracle_golden.py)"""
    z = load(name)
    bad, unsupported = [], 0
    for i, st in _games(z, oracle):
        if limit is not None and i >= limit:
            break
        _a, dig = wsim.rollout_random(st, 400)
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
            bad.append(int(z["seeds"][i]))
    assert not bad, bad[:10]
    assert unsupported <= max(1, len(z["seeds"]) * 3 // 200)


def test_every_step_equals_oracle(oracle, wsim):
    """legal mask and packed state after every step, default and random decks, incl. flagged games"""
    z = load("randdeck_chain.npz")
    import ref_harness as h
    for i, st in _games(z, oracle):
        if i >= 400:
            break
        sw, seed = st.copy(), int(z["seeds"][i])
        for k in range(400):
            m = oracle.legal_mask(st)
            assert np.array_equal(m, wsim.legal_mask(sw)), (seed, k)
            assert np.array_equal(m, wsim.legal_mask_packed(sw)), (seed, k)  # the streaming form (no unpack)
            legal = [a for a in range(156) if m[a >> 5] >> (a & 31) & 1]
            a = legal[h.agent_pick(seed, k, len(legal))]
            oracle.step(st, a)
            wsim.step(sw, a)
            assert st.tobytes() == sw.tobytes(), (seed, k, a)
            if st[19] & 1 or st[18]:
                break


def test_exact_draw_path(oracle, wsim):
    """the numpy-shaped slow path of the weighted draw (taken about once in 1e12 draws) forced on every draw"""
    L = wsim.lib(exact_draw=True)
    z = load("default_chain_10k.npz")
    for i, st in _games(z, oracle):
        if i >= 150:
            break
        _a, dig = wsim.rollout_random(st, 400, L=L)
        if z["err"][i] == 0:
            assert len(dig) == z["steps"][i] and chain_of(dig) == int(z["chain"][i])


def test_observation_and_features_vs_reference(wsim):
    z = load("obs_features.npz")
    for i in range(len(z["states"])):
        st = z["states"][i].copy()
        obs, err = wsim.observe(st)
        assert err == 0 and np.array_equal(obs, z["obs"][i]), i
        f, err = wsim.features(st)
        assert err == 0 and np.array_equal(f, z["feat"][i]), i
        obs, err = wsim.observe_packed(st)  # the streaming forms (no unpack) behind sb_observe / sb_features
        assert err == 0 and np.array_equal(obs, z["obs"][i]), i
        f, err = wsim.features_packed(st)
        assert err == 0 and np.array_equal(f, z["feat"][i]), i


def test_decisions_and_heuristic_games(oracle, wsim):
    z = load("heuristic_decisions.npz")
    for i in range(len(z["states"])):
        a, sc = wsim.select_action(z["states"][i].copy(), z["weights"][i])
        ao, sco, _m = oracle.select_action(z["states"][i].copy(), z["weights"][i])
        assert a == ao and np.array_equal(np.nan_to_num(sc, nan=-7.0), np.nan_to_num(sco, nan=-7.0)), i
        ref = z["scores"][i]
        legal = ~np.isnan(ref)
        assert np.allclose(sc[legal], ref[legal], rtol=1e-5, atol=1e-9)  # north star: scores within 1e-5 relative
    z = load("heuristic_games.npz")
    d0, d1 = default_decks()
    off = 0
    for i in range(len(z["seeds"])):
        n = int(z["lengths"][i])
        if i < 200 and int(z["result"][i]) != -2:
            st = oracle.new_game(int(z["seeds"][i]), d0, d1, 3, 2)
            res, steps, act = wsim.play_heuristic(st, z["w_first"][i], z["w_second"][i], 400)
            st[19] = 0
            assert res == int(z["result"][i]) and np.array_equal(act, z["actions"][off:off + n]) and oracle.digest(st) == int(z["final"][i]), i
        off += n


def test_expert_games(oracle, wsim):
    z = load("expert_tapes.npz")
    for i, st in _games(z, oracle):
        so = st.copy()
        for _k in range(int(z["steps"][i])):
            a, ao = wsim.expert_action(st), oracle.expert_action(so)
            wsim.step(st, a)
            oracle.step(so, ao)
            assert a == ao and st.tobytes() == so.tobytes(), i
            if st[18]:
                break
    z = load("heuristic_vs_expert.npz")
    d0, d1 = default_decks()
    for i in range(len(z["seeds"])):
        st = oracle.new_game(int(z["seeds"][i]), d0, d1, 3, 2)
        w = z["weights"][i]
        res, steps, _act = wsim.play_heuristic(st, w if z["seat"][i] == 0 else None, None if z["seat"][i] == 0 else w, 400)
        st[19] = 0
        assert res == int(z["result"][i]) and (res == -2 or oracle.digest(st) == int(z["final"][i])), i
