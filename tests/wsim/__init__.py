"""Host build of the warp-per-game engine (monsoon_b200/csrc/sbw_*.cuh) -- TEST INFRASTRUCTURE ONLY.

The engine source is single-source SPMD (sbw_warp.cuh): this compiles it with g++ (lane loops instead of lanes) so that the
CPU suite can check the rules of the warp engine against the oracle and the reference fixtures without a GPU.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libsb_wsim.so")
CSRC = os.path.join(os.path.dirname(os.path.dirname(HERE)), "monsoon_b200", "csrc")
_lib = None


def build(force=False, exact_draw=False):
    so = SO.replace(".so", "_exact.so") if exact_draw else SO
    deps = [os.path.join(HERE, "wsim.cpp")] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.startswith("sbw_") or f.startswith("sb_")]
    if not force and os.path.exists(so) and all(os.path.getmtime(d) <= os.path.getmtime(so) for d in deps):
        return so
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-strict-aliasing", "-Wall", "-Wno-unused-function", "-Wno-unknown-pragmas",
           "-x", "c++", os.path.join(HERE, "wsim.cpp"), "-o", so]
    if exact_draw:
        cmd.insert(1, "-DSB_FORCE_EXACT_DRAW")
    subprocess.check_call(cmd)
    return so


def lib(exact_draw=False):
    global _lib
    if _lib is not None and not exact_draw:
        return _lib
    L = ctypes.CDLL(build(exact_draw=exact_draw))
    L.wsim_rollout_random.restype = ctypes.c_int
    L.wsim_play_heuristic.restype = ctypes.c_int
    if not exact_draw:
        _lib = L
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def legal_mask(st, L=None):
    m = np.zeros(5, dtype=np.uint32)
    (L or lib()).wsim_legal_mask(_p(st), _p(m))
    return m


def legal_mask_packed(st):
    m = np.zeros(5, dtype=np.uint32)
    lib().wsim_legal_mask_packed(_p(st), _p(m))
    return m


def step(st, action, L=None):
    (L or lib()).wsim_step(_p(st), int(action))


def rollout_random(st, max_steps=400, L=None):
    dig = np.zeros(max_steps, dtype=np.uint64)
    act = np.zeros(max_steps, dtype=np.uint8)
    n = (L or lib()).wsim_rollout_random(_p(st), int(max_steps), _p(dig), _p(act))
    return act[:n], dig[:n]


def features(st):
    f = np.zeros(10, dtype=np.float64)
    return f, lib().wsim_features(_p(st), _p(f))


def features_packed(st):
    f = np.zeros(10, dtype=np.float64)
    return f, lib().wsim_features_packed(_p(st), _p(f))


def observe(st):
    obs = np.zeros((27, 5, 4), dtype=np.int32)
    return obs, lib().wsim_observe(_p(st), _p(obs))


def observe_packed(st):
    obs = np.zeros((27, 5, 4), dtype=np.int32)
    return obs, lib().wsim_observe_packed(_p(st), _p(obs))


def expert_action(st):
    return lib().wsim_expert_action(_p(st))


def select_action(st, w):
    w = np.ascontiguousarray(w, dtype=np.float64)
    scores = np.full(156, np.nan, dtype=np.float64)
    a = lib().wsim_select_action(_p(st), _p(w), _p(scores))
    return a, scores


def play_heuristic(st, w_first, w_second, max_steps=400):
    wf = None if w_first is None else np.ascontiguousarray(w_first, dtype=np.float64)
    ws = None if w_second is None else np.ascontiguousarray(w_second, dtype=np.float64)
    steps = ctypes.c_int(0)
    act = np.zeros(max_steps, dtype=np.uint8)
    res = lib().wsim_play_heuristic(_p(st), _p(wf), _p(ws), int(max_steps), ctypes.byref(steps), _p(act))
    return res, steps.value, act[:steps.value]
