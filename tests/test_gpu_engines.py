"""GPU: both engines behind the C ABI -- one game per thread (sb_engine.cuh) and one game per warp (sbw_*.cuh) -- against the
reference fixtures, forced explicitly (the default policy picks one of them by kernel and batch size), plus the CTA shapes
and the persistent grids of the warp kernels.  Identical results are the contract that makes the measured policy safe."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    p = os.path.join(G, name)
    if not os.path.exists(p):
        pytest.skip("fixture %s not generated" % name)
    return np.load(p)


@pytest.fixture(params=[0, 1], ids=["thread-per-game", "warp-per-game"])
def eng(request, engine):
    engine.set_option("engine", request.param)
    yield engine
    for k, v in (("engine", -1), ("w_shape", -1), ("w_grid", 0), ("w_hshape", -1)):
        engine.set_option(k, v)


def _check_chain(engine, z):
    dev = engine.device
    seeds = torch.from_numpy(z["seeds"].astype(np.int64)).to(dev)
    if "decks" in z.files:
        st = engine.reset(seeds, torch.from_numpy(z["decks"]).to(dev), torch.from_numpy(z["factions"]).to(dev))
    else:
        st = engine.reset(seeds)
    chain = torch.zeros(len(seeds), dtype=torch.int64, device=dev)
    steps = engine.rollout_random(st, 400, chain=chain)
    host = st.cpu().numpy()
    steps, chain = steps.cpu().numpy(), chain.cpu().numpy().view(np.uint64)
    unsupported = (host[:, 18] == 5) | ((host[:, 18] == 6) & ((z["err"] != 2) | (steps <= z["steps"])))
    clean = (z["err"] == 0) & ~unsupported
    assert np.array_equal(steps[clean], z["steps"][clean]) and np.array_equal(chain[clean], z["chain"][clean])
    raised = (z["err"] != 0) & ~unsupported
    assert (host[raised][:, 18] != 0).all() and np.array_equal(steps[raised], z["steps"][raised] + 1)
    assert unsupported.sum() <= max(1, len(unsupported) * 3 // 200)
    return host


@pytest.mark.parametrize("name", ["default_chain_10k.npz", "randdeck_chain.npz", "card_focus.npz"])
def test_reference_games(eng, name):
    _check_chain(eng, load(name))


@pytest.mark.parametrize("shape,grid", [(0, 0), (1, 0), (2, 0), (3, 0), (4, 0), (5, 0), (6, 0), (7, 0), (8, 0), (1, 5), (2, 3), (5, 2), (7, 3)])
def test_warp_kernel_shapes(engine, shape, grid):
    """every CTA shape of kw_rollout_random (independent warps, turn-synchronous CTAs, 32-register builds) and persistent
    grids far smaller than the batch (warps take games from the counter): results do not depend on the schedule"""
    try:
        engine.set_option("engine", 1)
        engine.set_option("w_shape", shape)
        engine.set_option("w_grid", grid)
        _check_chain(engine, load("randdeck_chain.npz"))
    finally:
        for k, v in (("engine", -1), ("w_shape", -1), ("w_grid", 0)):
            engine.set_option(k, v)


def test_engines_agree_step_by_step(engine, oracle):
    """sb_step / sb_legal_mask / sb_features / sb_observe / sb_expert_action / sb_select_action of the two engines on the same
    mid-game states (random decks, all cards), against each other and the oracle"""
    z = load("randdeck_chain.npz")
    dev = engine.device
    n = 512
    seeds = torch.from_numpy(z["seeds"][:n].astype(np.int64)).to(dev)
    engine.set_option("engine", 0)
    st0 = engine.reset(seeds, torch.from_numpy(z["decks"][:n]).to(dev), torch.from_numpy(z["factions"][:n]).to(dev))
    try:
        for rnd in range(12):
            engine.rollout_random(st0, 5)  # advance a few steps
            w = torch.from_numpy(np.random.RandomState(rnd).uniform(0, 1, (n, 10))).to(dev)
            out = {}
            for e in (0, 1):
                engine.set_option("engine", e)
                s = st0.clone()
                m = engine.legal_mask(s)
                f, fe = engine.features(s)
                o, oe = engine.observe(s)
                a, sc = engine.select_action(s, w, want_scores=True)
                x = engine.expert_action(s)
                # lowest legal action for everybody, then one step
                mm = m.cpu().numpy().view(np.uint32)
                act = np.array([next(k for k in range(156) if mm[i, k >> 5] >> (k & 31) & 1) for i in range(n)], dtype=np.uint8)
                nm = torch.empty_like(m)
                r, d, er = engine.step(s, torch.from_numpy(act).to(dev), next_masks=nm)
                out[e] = [t.cpu().numpy() for t in (m, f, fe, o, oe, a, sc, x, s, r, d, er, nm)]
            for u, v in zip(out[0], out[1]):
                assert np.array_equal(u, v, equal_nan=True)
        host = st0.cpu().numpy()
        for i in range(0, n, 37):  # spot check against the oracle
            f, _e = oracle.features(host[i].copy())
            engine.set_option("engine", 1)
            fg, _ = engine.features(st0[i:i + 1])
            assert np.array_equal(f, fg.cpu().numpy()[0])
    finally:
        engine.set_option("engine", -1)


def test_heuristic_games_both_engines(eng):
    """whole reference HeuristicAgent games (heuristic_games.npz): winner, length, final state; the warp kernels also through
    a persistent grid of 3 CTAs"""
    z = load("heuristic_games.npz")
    dev = eng.device
    n = 256
    seeds = torch.from_numpy(z["seeds"][:n].astype(np.int64)).to(dev)
    for grid in (0, 3):
        eng.set_option("w_grid", grid)
        st = eng.reset(seeds)
        res, steps = eng.rollout_heuristic(st, torch.from_numpy(z["w_first"][:n]).to(dev), torch.from_numpy(z["w_second"][:n]).to(dev))
        res, steps, host = res.cpu().numpy(), steps.cpu().numpy(), st.cpu().numpy()
        ok = z["result"][:n] != -2
        assert np.array_equal(res[ok], z["result"][:n][ok]) and np.array_equal(steps[ok], z["lengths"][:n][ok])
        host[:, 19] = 0
        import sb_oracle
        fin = np.array([sb_oracle.digest(host[i]) for i in range(n)], dtype=np.uint64)
        assert np.array_equal(fin[ok], z["final"][:n][ok])
    eng.set_option("w_grid", 0)


def test_observation_and_features_both_engines(eng):
    z = load("obs_features.npz")
    st = torch.from_numpy(z["states"]).to(eng.device)
    obs, err = eng.observe(st)
    assert int(err.max()) == 0 and np.array_equal(obs.cpu().numpy(), z["obs"])
    f, err = eng.features(st)
    assert int(err.max()) == 0 and np.array_equal(f.cpu().numpy(), z["feat"])
