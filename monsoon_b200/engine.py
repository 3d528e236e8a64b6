"""Batched Stormbound on one B200: device-resident packed states + the kernels of libsb_b200.so.

torch is plumbing only (device memory, streams, torch.distributed); every rule, draw and score is
computed by the CUDA kernels behind the C ABI (include/sb_b200.h).  No CPU path exists.
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib
from ._card_table import CARDS

STATE_BYTES = 512
N_ACTIONS = 156
MASK_WORDS = 5
PASS = 155

CARD_INDEX = {c["name"]: i for i, c in enumerate(CARDS)}
# games/stormbound.py:295-302
DEFAULT_DECKS = (
    ["UA07", "U007", "U306", "U061", "B304", "U305", "U320", "U302", "U313", "UA02", "UT32", "U316"],
    ["UA07", "U007", "U001", "U053", "UE01", "U211", "U206", "U071", "U020", "S013", "B001", "U061"],
)
DEFAULT_FACTIONS = (3, 2)  # Faction.IRONCLAD, Faction.SWARM


def deck_indices(names):
    return [CARD_INDEX[n.upper()] for n in names]


class Engine:
    """One handle per device (sb_create).  All tensors passed in must live on that device."""

    def __init__(self, device=0):
        if not torch.cuda.is_available():
            raise _lib.SbError("no CUDA device: the B200 simulator has no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
        h = ctypes.c_void_p()
        rc = self.lib.sb_create(self.device.index, ctypes.byref(h))
        if rc != 0:
            msg = self.lib.sb_last_error(h).decode() if h else "sb_create failed"
            raise _lib.SbError("sb_create(%d) -> %d: %s" % (self.device.index, rc, msg))
        self.h = h  # the caller's current device is left alone: every sb_* entry switches to the handle's device and back
        self._default_decks = None
        self.sm_count = self.lib.sb_sm_count(h)

    def close(self):
        if getattr(self, "h", None):
            self.lib.sb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    # ------------------------------------------------------------------ helpers
    def _check(self, rc, what):
        if rc != 0:
            raise _lib.SbError("%s -> %d: %s" % (what, rc, self.lib.sb_last_error(self.h).decode()))

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @staticmethod
    def _p(t):
        return ctypes.c_void_p(t.data_ptr()) if t is not None else None

    def _dev(self, t, dtype):
        assert t.device == self.device and t.dtype == dtype and t.is_contiguous(), (t.device, t.dtype)
        return t

    @property
    def launches(self):
        return int(self.lib.sb_launch_count(self.h))

    def set_option(self, key, value):
        """sb_set_option: tuning knobs of the rollout kernels (include/sb_b200.h)"""
        self._check(self.lib.sb_set_option(self.h, key.encode(), int(value)), "sb_set_option(%s)" % key)

    def empty_states(self, n):
        return torch.empty((n, STATE_BYTES), dtype=torch.uint8, device=self.device)

    # ------------------------------------------------------------------ C ABI wrappers (device tensors)
    def reset(self, seeds, decks=None, factions=None, out=None):
        """seeds: int64/uint64 tensor [n] on device.  decks: None (default decks), [2,k] shared or [n,2,k]."""
        n = seeds.numel()
        seeds = self._dev(seeds.view(torch.int64) if seeds.dtype != torch.int64 else seeds, torch.int64)
        if decks is None:
            if self._default_decks is None:  # built once: two pageable H2D copies per call otherwise
                self._default_decks = (torch.tensor([deck_indices(d) for d in DEFAULT_DECKS], dtype=torch.uint8, device=self.device),
                                       torch.tensor(DEFAULT_FACTIONS, dtype=torch.uint8, device=self.device))
            decks, factions = self._default_decks
        decks = self._dev(decks, torch.uint8)
        shared = 1 if decks.dim() == 2 else 0
        n_deck = decks.shape[-1]
        if factions is None:
            factions = torch.zeros((2,) if shared else (n, 2), dtype=torch.uint8, device=self.device)
        factions = self._dev(factions, torch.uint8)
        assert (factions.dim() == 1) == bool(shared)
        states = out if out is not None else self.empty_states(n)
        self._check(self.lib.sb_reset(self.h, n, self._p(seeds), self._p(decks), n_deck, shared, self._p(factions),
                                      self._p(states), self._stream()), "sb_reset")
        return states

    def generate_decks(self, seeds, generation, mode, n_preserve=0, q=0.0, archetypes=None, arch_factions=None, factions=None):
        """sb_generate_decks: per-game deck pairs of DeckEvolutionConfig.get_deck_configuration (utils.py:121-241).
        Returns (decks u8[n,2,12], factions u8[n,2]) on the device, ready for reset()."""
        seeds = torch.as_tensor(np.asarray(seeds, dtype=np.int64) if not torch.is_tensor(seeds) else seeds).to(self.device, torch.int64).contiguous()
        n = seeds.numel()
        arch = None if archetypes is None else np.ascontiguousarray(np.asarray(archetypes, dtype=np.uint8).reshape(24))
        af = None if arch_factions is None else np.ascontiguousarray(np.asarray(arch_factions, dtype=np.uint8).reshape(2))
        if factions is not None:
            factions = torch.as_tensor(factions).to(self.device, torch.uint8).contiguous()
            assert factions.shape == (n, 2)
        decks = torch.empty((n, 2, 12), dtype=torch.uint8, device=self.device)
        fout = torch.empty((n, 2), dtype=torch.uint8, device=self.device)
        self._check(self.lib.sb_generate_decks(self.h, n, self._p(seeds), int(generation), int(mode), int(n_preserve), float(q),
                                               None if arch is None else arch.ctypes.data, None if af is None else af.ctypes.data,
                                               self._p(factions), self._p(decks), self._p(fout), self._stream()), "sb_generate_decks")
        return decks, fout

    # ---- evolution-strategy operators on a resident population (include/sb_b200.h, sb_es_*)
    def es_offspring(self, seed, generation, mu, lam, tau, tau_prime, min_sigma, w, s, parents=None):
        assert w.shape == s.shape and w.shape[0] >= mu + lam and w.dtype == torch.float64 and w.is_contiguous() and s.is_contiguous()
        self._check(self.lib.sb_es_offspring(self.h, int(seed) & 0xFFFFFFFFFFFFFFFF, int(generation), mu, lam, w.shape[1], float(tau),
                                             float(tau_prime), float(min_sigma), self._p(w), self._p(s), self._p(parents), self._stream()),
                    "sb_es_offspring")

    def es_select(self, mu, fitness, w, s, want_order=False):
        total, nf = w.shape
        fitness = torch.as_tensor(fitness, dtype=torch.float64).to(self.device).contiguous()
        assert fitness.numel() == total
        wo = torch.empty((mu, nf), dtype=torch.float64, device=self.device)
        so = torch.empty_like(wo)
        fo = torch.empty(mu, dtype=torch.float64, device=self.device)
        order = torch.empty(mu, dtype=torch.int32, device=self.device) if want_order else None
        self._check(self.lib.sb_es_select(self.h, total, mu, nf, self._p(fitness), self._p(w), self._p(s), self._p(wo), self._p(so),
                                          self._p(fo), self._p(order), self._stream()), "sb_es_select")
        return wo, so, fo, order

    def es_reset_sigmas(self, seed, generation, initial_sigma, s):
        self._check(self.lib.sb_es_reset_sigmas(self.h, int(seed) & 0xFFFFFFFFFFFFFFFF, int(generation), s.shape[0], s.shape[1],
                                                float(initial_sigma), self._p(s), self._stream()), "sb_es_reset_sigmas")

    def es_inject_diversity(self, seed, generation, tau, tau_prime, min_sigma, initial_sigma, w, s, chosen=None):
        self._check(self.lib.sb_es_inject_diversity(self.h, int(seed) & 0xFFFFFFFFFFFFFFFF, int(generation), w.shape[0], w.shape[1], float(tau),
                                                    float(tau_prime), float(min_sigma), float(initial_sigma), self._p(w), self._p(s),
                                                    self._p(chosen), self._stream()), "sb_es_inject_diversity")

    def legal_mask(self, states, out=None):
        n = states.shape[0]
        masks = out if out is not None else torch.empty((n, MASK_WORDS), dtype=torch.int32, device=self.device)
        self._check(self.lib.sb_legal_mask(self.h, n, self._p(states), self._p(masks), self._stream()), "sb_legal_mask")
        return masks

    def step(self, states, actions, reward=None, done=None, err=None, next_masks=None):
        n = states.shape[0]
        actions = self._dev(actions, torch.uint8)
        reward = reward if reward is not None else torch.empty(n, dtype=torch.int8, device=self.device)
        done = done if done is not None else torch.empty(n, dtype=torch.uint8, device=self.device)
        err = err if err is not None else torch.empty(n, dtype=torch.uint8, device=self.device)
        self._check(self.lib.sb_step(self.h, n, self._p(states), self._p(actions), self._p(reward), self._p(done),
                                     self._p(err), self._p(next_masks), self._stream()), "sb_step")
        return reward, done, err

    def observe(self, states):
        n = states.shape[0]
        obs = torch.empty((n, 27, 5, 4), dtype=torch.int32, device=self.device)
        err = torch.empty(n, dtype=torch.uint8, device=self.device)
        self._check(self.lib.sb_observe(self.h, n, self._p(states), self._p(obs), self._p(err), self._stream()), "sb_observe")
        return obs, err

    def features(self, states):
        n = states.shape[0]
        f = torch.empty((n, 10), dtype=torch.float64, device=self.device)
        err = torch.empty(n, dtype=torch.uint8, device=self.device)
        self._check(self.lib.sb_features(self.h, n, self._p(states), self._p(f), self._p(err), self._stream()), "sb_features")
        return f, err

    def select_action(self, states, weights, want_scores=False):
        n = states.shape[0]
        weights = self._dev(weights, torch.float64)
        assert weights.shape == (n, 10)
        actions = torch.empty(n, dtype=torch.uint8, device=self.device)
        scores = torch.empty((n, N_ACTIONS), dtype=torch.float64, device=self.device) if want_scores else None
        self._check(self.lib.sb_select_action(self.h, n, self._p(states), self._p(weights), self._p(actions),
                                              self._p(scores), self._stream()), "sb_select_action")
        return actions, scores

    def expert_action(self, states):
        """Stormbound.expert_action per game; advances each state's random-stream draw counter in place."""
        n = states.shape[0]
        actions = torch.empty(n, dtype=torch.uint8, device=self.device)
        self._check(self.lib.sb_expert_action(self.h, n, self._p(states), self._p(actions), self._stream()), "sb_expert_action")
        return actions

    def rollout_random(self, states, max_steps=400, chain=None):
        n = states.shape[0]
        steps = torch.empty(n, dtype=torch.int32, device=self.device)
        self._check(self.lib.sb_rollout_random(self.h, n, self._p(states), max_steps, self._p(steps), self._p(chain),
                                               self._stream()), "sb_rollout_random")
        return steps

    def rollout_heuristic(self, states, w_first, w_second, idx_first=None, idx_second=None, max_steps=400):
        """Whole heuristic games; w_first / w_second = None hands that seat to the scripted expert opponent."""
        n = states.shape[0]
        if w_first is None and w_second is None:
            raise ValueError("at least one seat needs a weight table (expert-vs-expert games: expert_action + step)")
        w_first = None if w_first is None else self._dev(w_first, torch.float64)
        w_second = None if w_second is None else self._dev(w_second, torch.float64)
        for w, idx in ((w_first, idx_first), (w_second, idx_second)):  # the kernels index these tables unchecked
            if w is not None:
                assert w.dim() == 2 and w.shape[1] == 10, "weight tables are f64[P, 10] (SB_N_FEATURES)"
                if idx is None:
                    assert w.shape[0] >= n
                else:
                    self._dev(idx, torch.int32)
                    assert idx.numel() == n
        result = torch.empty(n, dtype=torch.int8, device=self.device)
        steps = torch.empty(n, dtype=torch.int32, device=self.device)
        self._check(self.lib.sb_rollout_heuristic(self.h, n, self._p(states), self._p(w_first), self._p(w_second),
                                                  self._p(idx_first), self._p(idx_second), max_steps, self._p(result),
                                                  self._p(steps), self._stream()), "sb_rollout_heuristic")
        return result, steps

    def accumulate_fitness(self, result, idx_first, counts):
        n = result.shape[0]
        self._check(self.lib.sb_accumulate_fitness(self.h, n, self._p(result), self._p(idx_first), self._p(counts),
                                                   self._stream()), "sb_accumulate_fitness")
        return counts

    def count_aborted(self, states, result, out=None):
        """i32[2] += {aborted like the reference would (err 1..4), aborted by a limit of this engine (err 5..7)}"""
        out = out if out is not None else torch.zeros(2, dtype=torch.int32, device=self.device)
        self._check(self.lib.sb_count_aborted(self.h, result.shape[0], self._p(states), self._p(result), self._p(out), self._stream()),
                    "sb_count_aborted")
        return out

    SCHEDULES = {"round_robin": 0, "versus": 1, "solo": 2}

    def eval_schedule(self, mode, n_ind, n_total, games_per_pair, base_seed, generation, game_lo, n):
        """sb_eval_schedule: (idx_first i32[n], idx_second i32[n], seeds i64[n]) of the games game_lo .. game_lo + n - 1"""
        i1 = torch.empty(n, dtype=torch.int32, device=self.device)
        i2 = torch.empty(n, dtype=torch.int32, device=self.device)
        seeds = torch.empty(n, dtype=torch.int64, device=self.device)
        self._check(self.lib.sb_eval_schedule(self.h, self.SCHEDULES[mode], n_ind, n_total, games_per_pair, int(base_seed) & 0xFFFFFFFFFFFFFFFF,
                                              int(generation), int(game_lo), n, self._p(i1), self._p(i2), self._p(seeds), self._stream()),
                    "sb_eval_schedule")
        return i1, i2, seeds

    def eval_population(self, mode, n_ind, weights, games_per_pair, base_seed, generation, game_lo, game_hi, max_steps=400,
                        counts=None, aborted=None, chunk_games=0):
        """sb_eval_population over the default decks: counts i32[n_ind,3] += {wins, draws, losses}, aborted i32[2] += ...; asynchronous."""
        weights = self._dev(weights, torch.float64)
        assert weights.dim() == 2 and weights.shape[1] == 10 and weights.shape[0] >= n_ind, "weight tables are f64[P, 10] (SB_N_FEATURES)"
        if self._default_decks is None:
            self._default_decks = (torch.tensor([deck_indices(d) for d in DEFAULT_DECKS], dtype=torch.uint8, device=self.device),
                                   torch.tensor(DEFAULT_FACTIONS, dtype=torch.uint8, device=self.device))
        decks, factions = self._default_decks
        counts = counts if counts is not None else torch.zeros((n_ind, 3), dtype=torch.int32, device=self.device)
        aborted = aborted if aborted is not None else torch.zeros(2, dtype=torch.int32, device=self.device)
        self._dev(counts, torch.int32)
        assert counts.numel() >= 3 * n_ind
        self._check(self.lib.sb_eval_population(self.h, self.SCHEDULES[mode], n_ind, weights.shape[0], games_per_pair,
                                                int(base_seed) & 0xFFFFFFFFFFFFFFFF, int(generation), int(game_lo), int(game_hi), self._p(weights),
                                                self._p(decks), decks.shape[-1], self._p(factions), max_steps, int(chunk_games), self._p(counts),
                                                self._p(aborted), self._stream()), "sb_eval_population")
        return counts, aborted

    # ------------------------------------------------------------------ host-buffer (e2e) entry points
    def step_host(self, states_np, actions_np, want_masks=True):
        """numpy in, numpy out; H2D + kernel + D2H inside the call (sb_step_host)."""
        n = states_np.shape[0]
        reward = np.empty(n, dtype=np.int8)
        done = np.empty(n, dtype=np.uint8)
        err = np.empty(n, dtype=np.uint8)
        masks = np.empty((n, MASK_WORDS), dtype=np.uint32) if want_masks else None
        self._check(self.lib.sb_step_host(self.h, n, states_np.ctypes.data, actions_np.ctypes.data, reward.ctypes.data,
                                          done.ctypes.data, err.ctypes.data, masks.ctypes.data if want_masks else None),
                    "sb_step_host")
        return reward, done, err, masks

    def rollout_random_host(self, seeds_np, decks_np=None, factions_np=None, max_steps=400, want_states=True, want_chain=False,
                            out_states=None, out_steps=None):
        """Host-buffer variant (sb_rollout_random_host).  out_states / out_steps: caller-owned result buffers, e.g. pinned."""
        n = seeds_np.shape[0]
        if decks_np is None:
            decks_np = np.array([deck_indices(d) for d in DEFAULT_DECKS], dtype=np.uint8)
            factions_np = np.array(DEFAULT_FACTIONS, dtype=np.uint8)
        seeds_np = np.ascontiguousarray(seeds_np, dtype=np.uint64)
        states = (out_states if out_states is not None else np.empty((n, STATE_BYTES), dtype=np.uint8)) if want_states else None
        steps = out_steps if out_steps is not None else np.empty(n, dtype=np.int32)
        assert steps.dtype == np.int32 and steps.shape == (n,) and (states is None or (states.dtype == np.uint8 and states.shape == (n, STATE_BYTES)))
        chain = np.zeros(n, dtype=np.uint64) if want_chain else None
        self._check(self.lib.sb_rollout_random_host(self.h, n, seeds_np.ctypes.data, decks_np.ctypes.data, decks_np.shape[-1],
                                                    factions_np.ctypes.data, max_steps,
                                                    states.ctypes.data if want_states else None, steps.ctypes.data,
                                                    chain.ctypes.data if want_chain else None), "sb_rollout_random_host")
        return states, steps, chain


_engines = {}


def get_engine(device=None):
    """One engine per device; None = the caller's current CUDA device (LOCAL_RANK under torchrun once the launcher has
    called torch.cuda.set_device), never a hard-wired device 0."""
    if device is None:
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        lr = os.environ.get("LOCAL_RANK")
        if device == 0 and lr is not None and torch.cuda.is_available():  # torchrun rank that never called set_device
            device = int(lr) % torch.cuda.device_count()
    if isinstance(device, torch.device):
        device = device.index or 0
    if device not in _engines:
        _engines[device] = Engine(device)
    return _engines[device]
