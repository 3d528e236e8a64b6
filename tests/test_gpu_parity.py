"""GPU parity tests: every call goes through the C ABI (libsb_b200.so) and is compared bit for bit
with the CPU oracle on the same seeded inputs.  Integer/byte state: exact.  FP64 features/scores:
exact op order is reproduced, the asserted tolerance is the north star's 1e-5 relative."""
import random

import numpy as np
import pytest
import torch

from monsoon_b200.engine import DEFAULT_DECKS, DEFAULT_FACTIONS, deck_indices
from monsoon_b200._card_table import CARDS

pytestmark = pytest.mark.gpu

D0, D1 = (deck_indices(d) for d in DEFAULT_DECKS)


def random_decks(seed, exclude=("UP01", "UP02", "UP03")):
    rng = random.Random(seed)
    decks, factions = [], []
    for _ in range(2):
        f = rng.choice([1, 2, 3, 4])
        pool = [i for i, c in enumerate(CARDS[:113]) if i > 0 and c["faction"] in (0, f) and c["name"] not in exclude]
        decks.append(rng.sample(pool, 12))
        factions.append(f)
    return decks, factions


def oracle_states(oracle, seeds, decks=None):
    out = []
    for i, s in enumerate(seeds):
        if decks is None:
            out.append(oracle.new_game(int(s), D0, D1, *DEFAULT_FACTIONS))
        else:
            d, f = decks[i]
            out.append(oracle.new_game(int(s), d[0], d[1], f[0], f[1]))
    return np.stack(out)


def test_reset_matches_oracle(engine, oracle):
    seeds = np.arange(512, dtype=np.int64) * 7919 + 3
    st = engine.reset(torch.from_numpy(seeds).to(engine.device))
    assert np.array_equal(st.cpu().numpy(), oracle_states(oracle, seeds))


def test_reset_per_game_decks(engine, oracle):
    n = 256
    seeds = np.arange(n, dtype=np.int64) + 1000
    decks = [random_decks(1000 + i) for i in range(n)]
    dk = torch.tensor([d for d, _ in decks], dtype=torch.uint8, device=engine.device)
    fc = torch.tensor([f for _, f in decks], dtype=torch.uint8, device=engine.device)
    st = engine.reset(torch.from_numpy(seeds).to(engine.device), dk, fc)
    assert np.array_equal(st.cpu().numpy(), oracle_states(oracle, seeds, decks))


@pytest.mark.parametrize("use_random_decks", [False, True])
def test_step_per_launch_trajectories(engine, oracle, use_random_decks):
    """sb_legal_mask + sb_step, one launch per env step, compared with the oracle after EVERY step."""
    n, max_steps = 192, 140
    seeds = np.arange(n, dtype=np.int64) + (50000 if use_random_decks else 0)
    decks = [random_decks(int(s), exclude=("UP01", "UP02", "UP03", "S203")) for s in seeds] if use_random_decks else None
    host = oracle_states(oracle, seeds, decks)
    dev = torch.from_numpy(host.copy()).to(engine.device)
    alive = np.ones(n, dtype=bool)
    for step in range(max_steps):
        masks = engine.legal_mask(dev).cpu().numpy().view(np.uint32)
        actions = np.full(n, 155, dtype=np.uint8)
        for i in range(n):
            if not alive[i]:
                continue
            om = oracle.legal_mask(host[i])
            assert np.array_equal(om, masks[i]), (i, step)
            legal = [a for a in range(156) if om[a >> 5] >> (a & 31) & 1]
            actions[i] = legal[oracle.lib().sbo_agent_pick(int(seeds[i]), step, len(legal))]
        frozen = host.copy()
        reward, done, err = engine.step(dev, torch.from_numpy(actions).to(engine.device))
        got = dev.cpu().numpy()
        for i in range(n):
            if not alive[i]:
                continue
            oracle.step(host[i], int(actions[i]))
            assert host[i].tobytes() == got[i].tobytes(), "game %d step %d action %d" % (i, step, actions[i])
            assert int(done[i]) == (host[i][19] & 1) and int(reward[i]) == ((host[i][19] >> 1) & 1)
            assert int(err[i]) == host[i][18]
            if host[i][19] & 1 or host[i][18]:
                alive[i] = False
        # finished games keep being stepped on the device with PASS; restore them so both sides stay equal
        if not alive.all():
            dead = np.where(~alive)[0]
            for i in dead:
                host[i] = got[i]
        if not alive.any():
            break
        del frozen


@pytest.mark.parametrize("use_random_decks", [False, True])
def test_rollout_kernel_matches_oracle(engine, oracle, use_random_decks):
    """Whole-game rollout in ONE launch (state stays on the SM) == the oracle's step-by-step rollout:
    final state bytes, step count and the chained per-step digests (i.e. the whole trajectory)."""
    n = 768
    seeds = np.arange(n, dtype=np.int64) * 31 + (90000 if use_random_decks else 17)
    decks = [random_decks(int(s), exclude=("UP01", "UP02", "UP03", "S203")) for s in seeds] if use_random_decks else None
    host = oracle_states(oracle, seeds, decks)
    dev = torch.from_numpy(host.copy()).to(engine.device)
    chain = torch.zeros(n, dtype=torch.int64, device=engine.device)
    steps = engine.rollout_random(dev, max_steps=400, chain=chain)
    got, steps, chain = dev.cpu().numpy(), steps.cpu().numpy(), chain.cpu().numpy().view(np.uint64)
    for i in range(n):
        _a, digests, _m = oracle.rollout_random(host[i], 400)
        ch = 0
        for d in digests:
            ch = ((ch ^ int(d)) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
        assert len(digests) == steps[i], i
        assert host[i].tobytes() == got[i].tobytes(), i
        assert ch == int(chain[i]), i


def _midgame_states(engine, n, steps, seed0=0):
    seeds = torch.arange(n, dtype=torch.int64, device=engine.device) + seed0
    st = engine.reset(seeds)
    engine.rollout_random(st, max_steps=steps)
    return st


def test_observation_and_features(engine, oracle):
    st = _midgame_states(engine, 512, 37, seed0=4000)
    obs, oerr = engine.observe(st)
    feat, ferr = engine.features(st)
    host = st.cpu().numpy()
    obs, feat = obs.cpu().numpy(), feat.cpu().numpy()
    for i in range(host.shape[0]):
        o_obs, e1 = oracle.observe(host[i])
        o_f, e2 = oracle.features(host[i])
        assert np.array_equal(o_obs, obs[i]), i
        assert e1 == int(oerr[i]) and e2 == int(ferr[i])
        assert np.array_equal(o_f, feat[i]), (i, o_f, feat[i])  # same op order => identical doubles


def test_select_action_scores(engine, oracle):
    n = 384
    st = _midgame_states(engine, n, 29, seed0=7000)
    w = torch.from_numpy(np.random.RandomState(5).uniform(0, 1, (n, 10))).to(engine.device)
    actions, scores = engine.select_action(st, w, want_scores=True)
    host, wn = st.cpu().numpy(), w.cpu().numpy()
    actions, scores = actions.cpu().numpy(), scores.cpu().numpy()
    for i in range(n):
        a, s, m = oracle.select_action(host[i].copy(), wn[i])
        legal = ~np.isnan(s)
        assert np.array_equal(legal, ~np.isnan(scores[i])), i
        np.testing.assert_allclose(scores[i][legal], s[legal], rtol=1e-5, atol=1e-12)
        assert a == int(actions[i]), (i, a, int(actions[i]))


def test_heuristic_rollout(engine, oracle):
    n = 48
    seeds = np.arange(n, dtype=np.int64) + 123
    host = oracle_states(oracle, seeds)
    dev = torch.from_numpy(host.copy()).to(engine.device)
    rs = np.random.RandomState(11)
    wf, ws = rs.uniform(0, 1, (n, 10)), rs.uniform(0, 1, (n, 10))
    result, steps = engine.rollout_heuristic(dev, torch.from_numpy(wf).to(engine.device), torch.from_numpy(ws).to(engine.device),
                                             max_steps=400)
    got, result, steps = dev.cpu().numpy(), result.cpu().numpy(), steps.cpu().numpy()
    for i in range(n):
        r, acts = oracle.play_heuristic(host[i], wf[i], ws[i], 400)
        assert (r, len(acts)) == (int(result[i]), int(steps[i])), i
        assert host[i].tobytes() == got[i].tobytes(), i


def test_features_and_decisions_with_unencodable_cards(engine, oracle):
    """UP01-03 have no observation id: `int(card)` raises in the reference as soon as one is on the board, in the
    mover's hand or deck, or in the play history.  The engine reports SB_ERR_OBS_ID for exactly those
    states (features, observation) and scores such candidates 0.0; games without these cards skip the id scans."""
    names = {c["name"]: i for i, c in enumerate(CARDS)}
    ups = [names["UP01"], names["UP02"], names["UP03"]]
    n = 96
    decks, factions = [], []
    for s in range(n):
        d, f = random_decks(900 + s, exclude=("UP01", "UP02", "UP03", "S203"))
        if s % 3 != 2:  # two thirds of the games carry one or two of the cards, in either deck
            d[s % 2][s % 12] = ups[s % 3]
            if s % 5 == 0:
                d[1 - s % 2][(s + 4) % 12] = ups[(s + 1) % 3]
        decks.append(d)
        factions.append(f)
    dev = engine.device
    seeds = torch.arange(n, dtype=torch.int64, device=dev) + 31000
    st = engine.reset(seeds, torch.tensor(decks, dtype=torch.uint8, device=dev), torch.tensor(factions, dtype=torch.uint8, device=dev))
    flagged = 0
    for depth in (0, 3, 9):
        if depth:
            engine.rollout_random(st, 3 if depth == 3 else 6)
        host = st.cpu().numpy()
        feat, ferr = engine.features(st)
        obs, oerr = engine.observe(st)
        w = torch.from_numpy(np.random.RandomState(depth).uniform(0, 1, (n, 10))).to(dev)
        acts, scores = engine.select_action(st, w, want_scores=True)
        feat, ferr, oerr, acts, scores = feat.cpu().numpy(), ferr.cpu().numpy(), oerr.cpu().numpy(), acts.cpu().numpy(), scores.cpu().numpy()
        for i in range(n):
            if host[i][18] or host[i][19] & 1:
                continue
            o_f, e2 = oracle.features(host[i])
            _o, e1 = oracle.observe(host[i])
            assert e2 == int(ferr[i]) and e1 == int(oerr[i]), (depth, i)
            flagged += e2 != 0
            if not e2:
                assert np.array_equal(o_f, feat[i]), (depth, i)
            a, sc, _m = oracle.select_action(host[i], w[i].cpu().numpy())
            assert a == int(acts[i]), (depth, i)
            legal = ~np.isnan(sc)
            assert np.array_equal(legal, ~np.isnan(scores[i])) and np.allclose(sc[legal], scores[i][legal], rtol=1e-5, atol=0), (depth, i)
    assert flagged >= 20  # the error path really was exercised
