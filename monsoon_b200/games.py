"""Drop-in for the reference's game layer (games/abstract_game.py:4-101, games/stormbound.py:121-250):
same method names, argument meaning and return conventions, backed by the CUDA kernels.

`Game` is the single-game wrapper the reference's callers use (evo/game_adapter.py:316-324,
play_vs_*.py); `BatchedGames` is the same contract over n games at once (the shape the GPU wants).
Where the reference would raise a Python exception inside step `EngineError` is raised.
"""
import numpy as np
import torch

from .engine import DEFAULT_DECKS, DEFAULT_FACTIONS, MASK_WORDS, N_ACTIONS, STATE_BYTES, deck_indices, get_engine

_ERR_TEXT = {1: "AttributeError: board.at(point) is None", 2: "ValueError: choice on an empty sequence",
             3: "IndexError/UnboundLocalError", 4: "ValueError: int(card) for UP01-03", 5: "construct outside the modelled subset",
             6: "packed-state overflow", 7: "recursion guard"}


class EngineError(RuntimeError):
    """The reference engine raises at this point (the code says which exception, include/sb_state.h)."""

    def __init__(self, code):
        super().__init__("engine error %d: %s" % (code, _ERR_TEXT.get(code, "?")))
        self.code = code


def mask_to_actions(mask_words):
    """u32[5] -> sorted list of legal action ints (games/stormbound.py:557)."""
    out = []
    for w in range(MASK_WORDS):
        v = int(mask_words[w]) & 0xFFFFFFFF
        while v:
            b = (v & -v).bit_length() - 1
            out.append(w * 32 + b)
            v &= v - 1
    return out


class BatchedGames:
    """n independent games resident on one GPU."""

    def __init__(self, seeds, decks=None, factions=None, device=None, engine=None):
        self.eng = engine or get_engine(device)
        dev = self.eng.device
        self.seeds = torch.as_tensor(np.asarray(seeds, dtype=np.int64)).to(dev)
        self.n = self.seeds.numel()
        self.decks = None if decks is None else torch.as_tensor(np.asarray(decks, dtype=np.uint8)).to(dev)
        self.factions = None if factions is None else torch.as_tensor(np.asarray(factions, dtype=np.uint8)).to(dev)
        self.states = self.eng.reset(self.seeds, self.decks, self.factions)
        self.masks = torch.empty((self.n, MASK_WORDS), dtype=torch.int32, device=dev)

    def legal_masks(self):
        return self.eng.legal_mask(self.states, out=self.masks)

    def step(self, actions):
        a = torch.as_tensor(actions, dtype=torch.uint8, device=self.eng.device)
        return self.eng.step(self.states, a, next_masks=self.masks)  # reward, done, err (+ fused next legal masks)

    def observe(self):
        return self.eng.observe(self.states)

    def expert_actions(self):
        """Stormbound.expert_action for every game (u8[n]); consumes draws of each game's stream like the reference."""
        return self.eng.expert_action(self.states)

    def to_play(self):
        # Stormbound.to_play: 0 if player == 1 else 1 (games/stormbound.py:312-313); player_sign is byte 16
        return (self.states[:, 16].view(torch.int8) != 1).to(torch.int64)


class _PlayerView:
    def __init__(self, game, which):
        self._g, self._which = game, which

    def _order(self):
        lo = int(self._g._host()[14])
        return lo if self._which == "local" else 1 - lo

    @property
    def strength(self):
        st = self._g._host()
        off = 32 + 104 * self._order()
        return int(np.frombuffer(st[off:off + 2].tobytes(), dtype="<i2")[0])

    @property
    def current_mana(self):
        st = self._g._host()
        off = 32 + 104 * self._order() + 4
        return int(np.frombuffer(st[off:off + 2].tobytes(), dtype="<i2")[0])


class _BoardView:
    def __init__(self, game):
        self.local = _PlayerView(game, "local")
        self.remote = _PlayerView(game, "remote")


class _Env:
    """`Game.env` duck type the reference's callers touch: get_observation, have_winner,
    board.local/remote.strength, actions."""

    def __init__(self, game):
        self._g = game
        self.board = _BoardView(game)
        self.actions = game._action_names

    def get_observation(self):
        return self._g._observe()

    def have_winner(self):  # games/stormbound.py:560-561: strictly negative base
        return self.board.local.strength < 0 or self.board.remote.strength < 0

    def legal_actions(self):
        return self._g.legal_actions()

    def expert_action(self):
        return self._g.expert_agent()

    def to_play(self):
        return self._g.to_play()

    def step(self, action):
        obs, reward, done = self._g.step(action)
        return obs, reward // 10, done


class Game:
    """games/stormbound.py:121-250 `Game` (an AbstractGame): one game, state on the GPU."""

    def __init__(self, seed=None, decks=None, factions=None, device=None, engine=None, _state=None):
        self.eng = engine or get_engine(device)
        if seed is None:  # the reference seeds RandomState from OS entropy in this case (Q16)
            seed = int(np.random.SeedSequence().entropy & 0x7FFFFFFFFFFFFFFF)
        self.seed = int(seed)
        self._action_names = _ACTION_NAMES
        if _state is not None:
            self.state = _state
        else:
            dev = self.eng.device
            d = None if decks is None else torch.tensor([deck_indices(x) if isinstance(x[0], str) else list(x) for x in decks],
                                                        dtype=torch.uint8, device=dev)
            f = None if decks is None else torch.tensor(list(factions or (0, 0)), dtype=torch.uint8, device=dev)
            self.state = self.eng.reset(torch.tensor([self.seed], dtype=torch.int64, device=dev), d, f)
        self._host_cache = None
        self.env = _Env(self)

    # -- helpers
    def _host(self):
        if self._host_cache is None:
            self._host_cache = self.state[0].cpu().numpy()
        return self._host_cache

    def _observe(self):
        obs, err = self.eng.observe(self.state)
        if int(err[0]):
            raise EngineError(int(err[0]))
        return obs[0].cpu().numpy()

    def clone(self):
        """copy.deepcopy(game) of evo/game_adapter.py:284: an independent state INCLUDING the stream position."""
        g = Game.__new__(Game)
        g.eng, g.seed, g._action_names = self.eng, self.seed, self._action_names
        g.state = self.state.clone()
        g._host_cache = None
        g.env = _Env(g)
        return g

    # -- AbstractGame contract
    def step(self, action):
        a = torch.tensor([int(action)], dtype=torch.uint8, device=self.eng.device)
        reward, done, err = self.eng.step(self.state, a)
        self._host_cache = None
        if int(err[0]):
            raise EngineError(int(err[0]))
        return self._observe(), int(reward[0]) * 10, bool(done[0])

    def to_play(self):
        return 0 if int(np.int8(self._host()[16])) == 1 else 1

    def legal_actions(self):
        return mask_to_actions(self.eng.legal_mask(self.state)[0].cpu().numpy().view(np.uint32))

    def reset(self):  # Stormbound.reset does not re-deal (games/stormbound.py:315-316)
        return self._observe()

    def render(self):
        print(self._observe()[[0, 1, 16, 17]])

    def close(self):
        pass

    def human_to_action(self):
        return int(input("action (0-155): "))

    def expert_agent(self):
        """games/stormbound.py:201-209 -> Stormbound.expert_action (:563-637); draws from the game's own stream."""
        a = self.eng.expert_action(self.state)
        self._host_cache = None
        err = int(self._host()[18])
        if err:
            raise EngineError(err)
        return int(a[0])

    def action_to_string(self, action_number):
        return self._action_names[action_number]


class EvolutionaryStormbound:
    """games/evolutionary_stormbound.py:21-232: the same engine with the decks of a DeckEvolutionConfig schedule.

    Mirrors the reference class member for member, including what it does NOT have: there is no `.env` attribute,
    so callers that reach for `game.env` (evo/game_adapter.py:334-371, evo/fitness.py:216) get the AttributeError the reference
    gives them -- FitnessEvaluator's faithful mode turns that into a draw exactly like evo/fitness.py:223-225.  `reset()`
    re-deals (:129-152), unlike Stormbound.reset; `step` returns the raw 0/1 reward (Stormbound.step, no x10 of Game.step).
    The reference keeps dealing from one RandomState; here deal number k of a game uses the stream key seed + k * 2^32.
    """

    def __init__(self, seed=None, generation=0, deck_config=None, device=None, engine=None):
        from .evo import DeckEvolutionConfig
        self.eng = engine or get_engine(device)
        if seed is None:
            seed = int(np.random.SeedSequence().entropy & 0x7FFFFFFF)
        self.seed = int(seed)
        self.generation = generation
        if deck_config is None:
            deck_config = self._create_default_deck_config()
        self.deck_config = deck_config
        self.actions = _ACTION_NAMES
        self.player = 1
        self._deals = 0
        self._host_rng = np.random.RandomState(self.seed & 0xFFFFFFFF)  # expert_agent placeholder only (:221-227)
        self._deal()

    @staticmethod
    def _create_default_deck_config():  # :77-122
        from .evo import DeckEvolutionConfig
        return DeckEvolutionConfig(DEFAULT_DECKS[0], DEFAULT_DECKS[1], exploit_generations=30, explore_generations=30,
                                   max_random_ratio=0.5, balance_archetype_ratio=0.7)

    def _deal(self):
        dev = self.eng.device
        key = (self.seed + (self._deals << 32)) & 0x7FFFFFFFFFFFFFFF
        self._deals += 1
        seeds = torch.tensor([key], dtype=torch.int64, device=dev)
        decks, factions = self.deck_config.generate_batch(self.eng, seeds, self.generation)
        names = self.deck_config._names
        d = decks[0].cpu().numpy()
        self.player1_deck, self.player2_deck = [names[c] for c in d[0]], [names[c] for c in d[1]]
        self.state = self.eng.reset(seeds, decks, factions)
        self.player = 1
        self._host_cache = None

    def _host(self):
        if self._host_cache is None:
            self._host_cache = self.state[0].cpu().numpy()
        return self._host_cache

    def to_play(self):
        return 0 if self.player == 1 else 1

    def reset(self):
        self._deal()
        return self.get_observation()

    def set_generation(self, generation):
        self.generation = generation  # applied by the next reset(), like the reference

    def get_phase_info(self):
        return self.deck_config.get_phase_info(self.generation)

    def step(self, action):
        a = torch.tensor([int(action)], dtype=torch.uint8, device=self.eng.device)
        reward, done, err = self.eng.step(self.state, a)
        self._host_cache = None
        if int(err[0]):
            raise EngineError(int(err[0]))
        self.player = int(np.int8(self._host()[16]))
        return self.get_observation(), int(reward[0]), bool(done[0])

    def legal_actions(self):
        return mask_to_actions(self.eng.legal_mask(self.state)[0].cpu().numpy().view(np.uint32))

    def get_observation(self):
        obs, err = self.eng.observe(self.state)
        if int(err[0]):
            raise EngineError(int(err[0]))
        return obs[0].cpu().numpy()

    def have_winner(self):
        st = self._host()
        bases = [int(np.frombuffer(st[32 + 104 * o:34 + 104 * o].tobytes(), dtype="<i2")[0]) for o in (0, 1)]
        return bases[0] < 0 or bases[1] < 0

    def clone(self):
        g = EvolutionaryStormbound.__new__(EvolutionaryStormbound)
        g.__dict__.update(self.__dict__)
        g.state = self.state.clone()
        g._host_cache = None
        return g

    def render(self):
        info = self.get_phase_info()
        print("=== Evolutionary Stormbound (Generation %d) ===" % self.generation)
        print("Phase: %s | Random Ratio: %.2f" % (info["phase"], info["random_ratio"]))
        print(self.get_observation()[[0, 1, 16, 17]])

    def close(self):
        pass

    def expert_agent(self):  # :221-227 "placeholder": a uniformly random legal action, else PASS
        legal = self.legal_actions()
        return int(self._host_rng.choice(legal)) if legal else 155

    def action_to_string(self, action_number):
        if action_number < len(self.actions):
            return self.actions[action_number]
        return None

    def human_to_action(self):
        return int(input("action (0-155): "))


def _make_action_names():
    """actions.txt equivalent, generated from the comment block enums.py:10-36 / :85-105."""
    names = []
    for card in range(4):
        for y in range(4, 0, -1):
            for x in range(4):
                names.append("Place unit or structure card at index %d of hand at (%d, %d)" % (card, x, y))
    for card in range(4):
        names.append("Use spell card at index %d of hand with no target" % card)
        for y in range(4, -1, -1):
            for x in range(4):
                names.append("Use spell card at index %d of hand at (%d, %d)" % (card, x, y))
    for card in range(4):
        names.append("Replace card at index %d of hand" % card)
    for card in range(1, 4):
        names.append("Move card at index %d of hand to leftmost" % card)
    names.append("Pass the turn")
    assert len(names) == N_ACTIONS
    return names


_ACTION_NAMES = _make_action_names()
__all__ = ["Game", "EvolutionaryStormbound", "BatchedGames", "EngineError", "mask_to_actions", "DEFAULT_DECKS", "DEFAULT_FACTIONS", "STATE_BYTES"]
