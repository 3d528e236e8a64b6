"""TEST INFRASTRUCTURE -- ctypes binding of the C oracle (oracle/libsb_oracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this.  The product package (monsoon_b200) never does.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libsb_oracle.so")
S = 512

_lib = None


def build(force=False):
    srcs = ["sb_oracle.c", "sb_oracle_effects.c", "sb_oracle_agent.c", "sb_oracle_es.c", "sb_oracle.h", "sb_card_table.inc",
            "sb_card_ids.h", os.path.join("..", "include", "sb_state.h")]
    newest = max(os.path.getmtime(os.path.join(HERE, s)) for s in srcs)
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < newest:
        subprocess.check_call(["make", "-C", HERE, "-s"])
    return SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            build()
        L = ctypes.CDLL(SO)
        vp, u8p = ctypes.c_void_p, ctypes.c_void_p
        L.sbo_state_bytes.restype = ctypes.c_int
        L.sbo_new_game.argtypes = [vp, ctypes.c_uint64, u8p, u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.sbo_legal_mask.argtypes = [vp, vp]
        L.sbo_step.argtypes = [vp, ctypes.c_int]
        L.sbo_digest.argtypes = [vp]
        L.sbo_digest.restype = ctypes.c_uint64
        L.sbo_rollout_random.argtypes = [vp, ctypes.c_int, vp, vp, vp]
        L.sbo_rollout_random.restype = ctypes.c_int
        L.sbo_observe.argtypes = [vp, vp]
        L.sbo_observe.restype = ctypes.c_int
        L.sbo_features.argtypes = [vp, vp]
        L.sbo_features.restype = ctypes.c_int
        L.sbo_select_action.argtypes = [vp, vp, vp, vp]
        L.sbo_select_action.restype = ctypes.c_int
        L.sbo_play_heuristic.argtypes = [vp, vp, vp, ctypes.c_int, vp, vp]
        L.sbo_play_heuristic.restype = ctypes.c_int
        L.sbo_batch_random.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp]
        L.sbo_batch_random.restype = ctypes.c_long
        L.sbo_batch_heuristic.argtypes = [vp, ctypes.c_int, vp, vp, vp, vp, ctypes.c_int, ctypes.c_int, vp, vp]
        L.sbo_batch_heuristic.restype = ctypes.c_long
        L.sbo_expert_action.argtypes = [vp]
        L.sbo_expert_action.restype = ctypes.c_int
        L.sbo_generate_decks.argtypes = [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int, ctypes.c_int, ctypes.c_double, vp, vp, vp]
        L.sbo_generate_decks.restype = None
        dbl, u64, u32, i32 = ctypes.c_double, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int
        L.sbo_det_log.argtypes = [dbl]
        L.sbo_det_log.restype = dbl
        L.sbo_det_exp.argtypes = [dbl]
        L.sbo_det_exp.restype = dbl
        L.sbo_es_offspring.argtypes = [u64, u32, i32, i32, i32, dbl, dbl, dbl, vp, vp, vp]
        L.sbo_es_select.argtypes = [i32, i32, i32, vp, vp, vp, vp, vp, vp, vp]
        L.sbo_es_reset_sigmas.argtypes = [u64, u32, i32, i32, dbl, vp]
        L.sbo_es_inject_diversity.argtypes = [u64, u32, i32, i32, dbl, dbl, dbl, dbl, vp, vp, vp]
        for f in (L.sbo_es_offspring, L.sbo_es_select, L.sbo_es_reset_sigmas, L.sbo_es_inject_diversity):
            f.restype = None
        L.sbo_agent_pick.argtypes = [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32]
        L.sbo_agent_pick.restype = ctypes.c_uint32
        assert L.sbo_state_bytes() == S
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data if a is not None else None


def new_game(seed, deck0, deck1, faction0=3, faction1=2):
    st = np.zeros(S, dtype=np.uint8)
    d0 = np.ascontiguousarray(deck0, dtype=np.uint8)
    d1 = np.ascontiguousarray(deck1, dtype=np.uint8)
    assert len(d0) == len(d1)
    lib().sbo_new_game(_p(st), int(seed), _p(d0), _p(d1), len(d0), faction0, faction1)
    return st


def legal_mask(st):
    m = np.zeros(5, dtype=np.uint32)
    lib().sbo_legal_mask(_p(st), _p(m))
    return m


def step(st, action):
    lib().sbo_step(_p(st), int(action))
    return st


def expert_action(st):
    """Stormbound.expert_action (games/stormbound.py:563-637); advances the random stream stored in st."""
    return int(lib().sbo_expert_action(_p(st)))


def generate_decks(seed, generation, mode, n_preserve, q, archetypes, factions):
    """DeckEvolutionConfig.get_deck_configuration (utils.py:121-241) for ONE game -> u8[2,12] card ids."""
    arch = np.ascontiguousarray(np.asarray(archetypes, dtype=np.uint8).reshape(24))
    fac = np.ascontiguousarray(np.asarray(factions, dtype=np.uint8).reshape(2))
    out = np.zeros((2, 12), dtype=np.uint8)
    lib().sbo_generate_decks(int(seed), int(generation), int(mode), int(n_preserve), float(q), _p(arch), _p(fac), _p(out))
    return out


# ---- evolution-strategy operators (sb_oracle_es.c); w, s: float64 [rows, nf] C-contiguous, modified in place
def es_offspring(seed, generation, mu, lam, tau, tau_prime, min_sigma, w, s):
    parents = np.zeros(lam, dtype=np.int32)
    lib().sbo_es_offspring(int(seed), int(generation), mu, lam, w.shape[1], tau, tau_prime, min_sigma, _p(w), _p(s), _p(parents))
    return parents


def es_select(mu, fitness, w, s):
    total, nf = w.shape
    fitness = np.ascontiguousarray(fitness, dtype=np.float64)
    wo, so, fo = np.zeros((mu, nf)), np.zeros((mu, nf)), np.zeros(mu)
    order = np.zeros(mu, dtype=np.int32)
    lib().sbo_es_select(total, mu, nf, _p(fitness), _p(w), _p(s), _p(wo), _p(so), _p(fo), _p(order))
    return wo, so, fo, order


def es_reset_sigmas(seed, generation, initial_sigma, s):
    lib().sbo_es_reset_sigmas(int(seed), int(generation), s.shape[0], s.shape[1], initial_sigma, _p(s))


def es_inject_diversity(seed, generation, tau, tau_prime, min_sigma, initial_sigma, w, s):
    chosen = np.zeros(max(1, w.shape[0] // 2), dtype=np.int32)
    lib().sbo_es_inject_diversity(int(seed), int(generation), w.shape[0], w.shape[1], tau, tau_prime, min_sigma, initial_sigma,
                                  _p(w), _p(s), _p(chosen))
    return chosen


def digest(st):
    return int(lib().sbo_digest(_p(st)))


def rollout_random(st, max_steps=400):
    actions = np.zeros(max_steps, dtype=np.uint8)
    digests = np.zeros(max_steps, dtype=np.uint64)
    masks = np.zeros((max_steps, 5), dtype=np.uint32)
    n = lib().sbo_rollout_random(_p(st), max_steps, _p(actions), _p(digests), _p(masks))
    return actions[:n], digests[:n], masks[:n]


def observe(st):
    obs = np.zeros((27, 5, 4), dtype=np.int32)
    err = lib().sbo_observe(_p(st), _p(obs))
    return obs, err


def features(st):
    f = np.zeros(10, dtype=np.float64)
    err = lib().sbo_features(_p(st), _p(f))
    return f, err


def select_action(st, w):
    w = np.ascontiguousarray(w, dtype=np.float64)
    scores = np.full(156, np.nan, dtype=np.float64)
    mask = np.zeros(5, dtype=np.uint32)
    a = lib().sbo_select_action(_p(st), _p(w), _p(scores), _p(mask))
    return a, scores, mask


def play_heuristic(st, w_first, w_second, max_steps=400):
    wf = None if w_first is None else np.ascontiguousarray(w_first, dtype=np.float64)   # None: that seat plays expert_action
    ws = None if w_second is None else np.ascontiguousarray(w_second, dtype=np.float64)
    actions = np.zeros(max_steps, dtype=np.uint8)
    n = ctypes.c_int(0)
    r = lib().sbo_play_heuristic(_p(st), _p(wf), _p(ws), max_steps, _p(actions), ctypes.addressof(n))
    return r, actions[:n.value]


def batch_random(states, max_steps=400, nthreads=1):
    states = np.ascontiguousarray(states)
    steps = np.zeros(len(states), dtype=np.int32)
    tot = lib().sbo_batch_random(_p(states), len(states), max_steps, nthreads, _p(steps))
    return tot, steps


def batch_heuristic(states, w_first, w_second, idx_first=None, idx_second=None, max_steps=400, nthreads=1):
    states = np.ascontiguousarray(states)
    wf = np.ascontiguousarray(w_first, dtype=np.float64).reshape(-1, 10)
    ws = np.ascontiguousarray(w_second, dtype=np.float64).reshape(-1, 10)
    i1 = None if idx_first is None else np.ascontiguousarray(idx_first, dtype=np.int32)
    i2 = None if idx_second is None else np.ascontiguousarray(idx_second, dtype=np.int32)
    res = np.zeros(len(states), dtype=np.int32)
    steps = np.zeros(len(states), dtype=np.int32)
    tot = lib().sbo_batch_heuristic(_p(states), len(states), _p(wf), _p(ws), _p(i1), _p(i2), max_steps, nthreads,
                                    _p(res), _p(steps))
    return tot, res, steps
