"""TEST INFRASTRUCTURE -- drives the UNMODIFIED reference engine (/root/reference) with an injected
counter-based Philox stream and serialises its object graph into the packed state of sb_layout.py.

Only usable in the build container (the reference does not travel to the GPU box).  Used by
tests/golden/make_golden.py (fixture generation) and oracle/validate_vs_reference.py (live check of
the C oracle).  Nothing in the product package imports this module.

Reference entry points exercised: games/stormbound.py:293-373 (Stormbound ctor/step),
:528-557 (legal_actions), player.py:13-37, board.py:16-28.
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SB_REFERENCE", "/root/reference")

from sb_layout import (STATE_DTYPE, CF_FIXED, CF_OBJ, CF_SINGLE_USE, PF_LEFTMOST_MOVABLE, PF_REPLACABLE, ST_BITS,
                       TF_FIXED, TF_OWNER, TF_STRUCTURE, weight_table, fnv1a64)  # noqa: E402

M32 = 0xFFFFFFFF


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox-4x32-10 (Salmon et al., SC'11), the published round function and Weyl constants."""
    for _ in range(10):
        p0 = 0xD2511F53 * c0
        p1 = 0xCD9E8D57 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & M32, p1 & M32, ((p0 >> 32) ^ c3 ^ k1) & M32, p0 & M32
        k0 = (k0 + 0x9E3779B9) & M32
        k1 = (k1 + 0xBB67AE85) & M32
    return c0, c1, c2, c3


class PhiloxRandomState:
    """Drop-in for the 5 call shapes the rules engine uses on np.random.RandomState.

    Stream: word block = philox(counter=(draw, turn, 0, 0), key=(seed_lo, seed_hi)); every call
    consumes exactly one block (shuffle of n: n-1 blocks).  Plain attributes -> copy.deepcopy clones the
    stream position, as evo/game_adapter.py:284 relies on.
    """

    def __init__(self, seed):
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.turn = 0
        self.draw = 0
        self.calls = 0

    def _block(self):
        out = philox4x32_10(self.draw & M32, self.turn & M32, 0, 0, self.seed & M32, self.seed >> 32)
        self.draw += 1
        self.calls += 1
        return out

    def _below(self, n):
        if n <= 0:
            raise ValueError("empty range")
        return (self._block()[0] * n) >> 32

    def random(self):
        w = self._block()
        return ((w[0] >> 5) * 67108864 + (w[1] >> 6)) / 9007199254740992.0

    def randint(self, low, high=None):
        if high is None:
            low, high = 0, low
        return low + self._below(high - low)

    def shuffle(self, seq):
        for i in range(len(seq) - 1, 0, -1):
            j = self._below(i + 1)
            seq[i], seq[j] = seq[j], seq[i]

    def choice(self, seq, size=None, p=None):
        seq = list(seq)
        if len(seq) == 0:
            raise ValueError("'a' cannot be empty unless no samples are taken")
        if p is None:
            assert size is None
            return seq[self._below(len(seq))]
        # numpy legacy choice(p=): cdf = cumsum(p); cdf /= cdf[-1]; searchsorted(cdf, u, side="right")
        assert size == 1
        cdf = []
        acc = 0.0
        for q in p:
            acc = acc + q
            cdf.append(acc)
        last = cdf[-1]
        u = self.random()
        idx = 0
        for c in cdf:
            if c / last <= u:
                idx += 1
        return [seq[min(idx, len(seq) - 1)]]


def agent_pick(seed, step, n):
    """Uniform-random agent stream: philox(counter=(step,0,0xA6E7,0), key=seed)."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    w = philox4x32_10(step & M32, 0, 0xA6E7, 0, seed & M32, seed >> 32)
    return (w[0] * n) >> 32


_ref = None


def ref():
    """Import the reference once (flat module names, cwd-relative opens: games/stormbound.py:306-310)."""
    global _ref
    if _ref is not None:
        return _ref
    if not os.path.isdir(REF):
        raise RuntimeError("reference not available at %s" % REF)
    for p in (REF, os.path.join(HERE, "refshim")):
        if p not in sys.path:
            sys.path.insert(0, p)
    cwd = os.getcwd()
    os.chdir(REF)
    try:
        import cards as c
        import games.stormbound as gs
        import board, player, unit, structure, spell, enums, point  # noqa: E401
    finally:
        os.chdir(cwd)

    class NS:
        pass

    ns = NS()
    ns.cards, ns.gs, ns.board, ns.player, ns.unit = c, gs, board, player, unit
    ns.structure, ns.spell, ns.enums, ns.point = structure, spell, enums, point
    sys.path.insert(0, os.path.dirname(HERE))
    from monsoon_b200._card_table import CARDS
    ns.table = CARDS
    ns.index = {r["name"]: i for i, r in enumerate(CARDS)}
    with open(os.path.join(REF, "actions.txt")) as f:
        ns.actions = f.read().splitlines()
    ns.wn = {float(w): n for n, w in enumerate(weight_table())}
    # Q14: cards/s203.py:27 `list(set(tiles))` iterates in str-hash order (PYTHONHASHSEED).  The module-level name `set`
    # is shadowed INSIDE cards.s203 only (the reference source is untouched, same technique as the RNG injection) so the
    # dedupe keeps first-occurrence order -- the canonical order of the oracle and the CUDA engine.
    c.s203.set = lambda it: dict.fromkeys(it)
    ns.activations = _install_activation_counter(c, ns.index)
    _ref = ns
    return ns


def _install_activation_counter(cards_mod, index):
    """Count every activate_ability call per card class (coverage table of tests/golden/make_golden_r2.py).  The wrapper sits
    OUTSIDE the trigger bookkeeping of card.py:48-62 and only increments a counter."""
    import collections
    import functools
    counts = collections.Counter()
    for name in index:
        cls = getattr(cards_mod, name, None)
        if cls is None or "activate_ability" not in cls.__dict__:
            continue
        inner = cls.activate_ability

        def make(inner, name):
            @functools.wraps(inner)
            def counted(self, *a, **k):
                counts[name] += 1
                return inner(self, *a, **k)
            return counted
        cls.activate_ability = make(inner, name)
    return counts


DEFAULT_DECKS = (
    ["UA07", "U007", "U306", "U061", "B304", "U305", "U320", "U302", "U313", "UA02", "UT32", "U316"],
    ["UA07", "U007", "U001", "U053", "UE01", "U211", "U206", "U071", "U020", "S013", "B001", "U061"],
)
DEFAULT_FACTIONS = (3, 2)  # IRONCLAD, SWARM (games/stormbound.py:295,299)


def make_game(seed, decks=None, factions=None):
    """Repeat games/stormbound.py:293-310 with the Philox stream in place BEFORE the Player
    constructors run (they shuffle and draw, player.py:28,35).  Returns a reference `Game`."""
    r = ref()
    gs = r.gs
    decks = decks or DEFAULT_DECKS
    factions = factions or DEFAULT_FACTIONS
    rnd = PhiloxRandomState(seed)

    class TurnCountingStormbound(gs.Stormbound):
        def step(self, action):  # the reference has no turn counter; the stream is keyed by it
            if action == 155:
                self.random.turn += 1
                self.random.draw = 0
            return super().step(action)

        def get_observation(self):  # Q12: int(card) raises for UP01-03 (card.py:46) AFTER the step has been applied;
            try:                     # the bypass lets fixtures pin the effects of those three cards through step()
                return super().get_observation()
            except ValueError:
                return None

    env = TurnCountingStormbound.__new__(TurnCountingStormbound)
    env.random = rnd
    mk = lambda names: [getattr(r.cards, n)() for n in names]  # noqa: E731
    local = r.player.Player(r.enums.Faction(factions[0]), mk(decks[0]), r.enums.PlayerOrder.FIRST, rnd)
    remote = r.player.Player(r.enums.Faction(factions[1]), mk(decks[1]), r.enums.PlayerOrder.SECOND, rnd)
    env.board = r.board.Board(local, remote, rnd)
    env.player = 1
    env.actions = r.actions
    env.cards = []
    env.n_steps = 0
    game = gs.Game.__new__(gs.Game)
    game.env = env
    return game


def card_index(card):
    r = ref()
    cid = card.card_id
    if cid[0] == "f":
        return 113 + int(cid[1:])
    name = type(card).__name__.upper()
    if name in r.index:
        return r.index[name]
    if isinstance(card, r.structure.Structure):
        return 129
    raise KeyError(cid)


def _status_word(unit):
    w = 0
    counts = [0] * 5
    for s in unit.status_effects:
        counts[int(s)] += 1
        w += 1 << (ST_BITS * int(s))
    if max(counts) > 63:
        raise PackOverflow("status counter")  # 64 copies of one status do not fit the 6-bit field
    return w


class PackOverflow(Exception):
    """The reference state no longer fits the packed layout (deck > 16 cards, hand > 4, ...)."""


def pack_reference(game, steps=0, done=0, err=0):
    """Serialise the reference object graph into one packed 512-byte state."""
    r = ref()
    env = game.env
    b = env.board
    for pl in (b.local, b.remote):
        if len(pl.deck) > 16 or len(pl.hand) > 4:
            raise PackOverflow()
    s = np.zeros((), dtype=STATE_DTYPE)
    s["seed_lo"] = env.random.seed & M32
    s["seed_hi"] = env.random.seed >> 32
    s["turn"] = env.random.turn
    s["draw"] = env.random.draw
    s["steps"] = steps
    s["local_order"] = int(b.local.order)
    s["current_order"] = int(b.current_player.order)
    s["player_sign"] = env.player
    s["phase"] = int(b.phase)
    s["err"] = err
    s["done"] = done
    hist = b.history[-4:]
    s["hist_n"] = len(hist)
    for i, c in enumerate(hist):
        s["hist_card"][i] = card_index(c)
        s["hist_owner"][i] = int(c.player.order)
    objs = []

    def card_flags(c):
        f = (CF_FIXED if getattr(c, "fixedly_forward", False) else 0) | (CF_SINGLE_USE if c.is_single_use else 0)
        if getattr(c, "position", None) is not None:
            f |= CF_OBJ
        return f

    def note_obj(c, order, in_deck, i):
        if getattr(c, "position", None) is None:
            return
        on_board = c.position.is_valid and b.board[c.position.y][c.position.x] is c
        objs.append(((order << 7) | (in_deck << 6) | i, c.position.y * 4 + c.position.x if on_board else 0xFF,
                     0 if on_board else c.strength))

    for pl in (b.local, b.remote):
        p = s["pl"][int(pl.order)]
        p["base"], p["max_mana"], p["mana"] = pl.strength, pl.max_mana, pl.current_mana
        p["front_line"] = pl.front_line
        p["flags"] = (PF_REPLACABLE if pl.replacable else 0) | (PF_LEFTMOST_MOVABLE if pl.leftmost_movable else 0)
        p["n_hand"], p["n_deck"], p["faction"] = len(pl.hand), len(pl.deck), int(pl.faction)
        for i, c in enumerate(pl.hand):
            p["hand_card"][i] = card_index(c)
            p["hand_cost"][i] = c.cost
            p["hand_flags"][i] = card_flags(c)
        for i, c in enumerate(pl.deck):
            p["deck_card"][i] = card_index(c)
            p["deck_cost"][i] = c.cost
            p["deck_flags"][i] = card_flags(c)
            p["deck_wn"][i] = r.wn[float(c.weight)]
    for o in (0, 1):  # ext serialisation order = hand then deck, FIRST then SECOND (o_pack)
        pl = b.local if int(b.local.order) == o else b.remote
        for i, c in enumerate(pl.hand):
            note_obj(c, o, 0, i)
        for i, c in enumerate(pl.deck):
            note_obj(c, o, 1, i)
    ext = s["ext"]
    ext[91] = len(objs)
    for k, (loc, tile, strength) in enumerate(objs[:4]):
        ext[92 + 4 * k] = loc
        ext[93 + 4 * k] = tile
        ext[94 + 4 * k] = strength & 255
        ext[95 + 4 * k] = (strength >> 8) & 255
    nmem = [0]

    def emit_mem(m, key):
        if nmem[0] >= 9:
            raise PackOverflow()
        me = nmem[0]
        nmem[0] += 1
        is_struct = isinstance(m, r.structure.Structure)
        base = 1 + 10 * me
        ext[base] = key
        ext[base + 1] = m.position.y * 4 + m.position.x
        ext[base + 2] = card_index(m)
        ext[base + 3] = (TF_OWNER if int(m.player.order) else 0) | (TF_STRUCTURE if is_struct else 0) | \
                        (TF_FIXED if getattr(m, "fixedly_forward", False) else 0) | \
                        (0 if (m.player is b.local or m.player is b.remote) else 8)  # detached: deep-copied Player/Board
        ext[base + 4] = m.strength & 255
        ext[base + 5] = (m.strength >> 8) & 255
        w = 0 if is_struct else _status_word(m)
        for q in range(4):
            ext[base + 6 + q] = (w >> (8 * q)) & 255
        for c in getattr(m, "ability_remembered", None) or []:  # a remembered temple copy keeps its own memories
            emit_mem(c, 0x80 | me)

    for y in range(5):
        for x in range(4):
            e = b.board[y][x]
            if e is None or type(e).__name__ != "B005":
                continue
            for m in e.ability_remembered:
                emit_mem(m, y * 4 + x)
    nmem = nmem[0]
    ext[0] = nmem
    for y in range(5):
        for x in range(4):
            e = b.board[y][x]
            if e is None:
                continue
            t = s["tile"][y * 4 + x]
            t["card"] = card_index(e)
            is_struct = isinstance(e, r.structure.Structure)
            t["flags"] = (TF_OWNER if int(e.player.order) else 0) | (TF_STRUCTURE if is_struct else 0) | \
                         (TF_FIXED if getattr(e, "fixedly_forward", False) else 0)
            t["strength"] = e.strength
            t["status"] = 0 if is_struct else _status_word(e)
    return s


def legal_mask(actions):
    m = np.zeros(5, dtype=np.uint32)
    for a in actions:
        m[a >> 5] |= np.uint32(1 << (a & 31))
    return m


@contextlib.contextmanager
def quiet():
    """cards/u040.py:14 prints inside the ability."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield


def play_random_game(seed, decks=None, factions=None, max_steps=400, record=True):
    """Uniform-random legal agent (agent_pick stream) until `done` or max_steps.

    Returns dict(actions u8[n], masks u32[n,5] (legal set BEFORE each action), digests u64[n]
    (fnv1a64 of the packed state AFTER each action), states (list of packed states if record),
    init packed state, err flag (an exception escaped step), final packed state).
    """
    game = make_game(seed, decks, factions)
    env = game.env
    init = pack_reference(game)
    actions, masks, digests, states, rewards, dones = [], [], [], [], [], []
    err = 0
    done = False
    step = 0
    with quiet():
        while not done and step < max_steps:
            legal = game.legal_actions()
            a = legal[agent_pick(seed, step, len(legal))]
            try:
                _obs, reward, done = game.step(a)
            except Exception as e:  # noqa: BLE001 -- the reference's callers swallow these (Q11)
                err = 1
                actions.append(a)
                masks.append(legal_mask(legal))
                break
            step += 1
            try:
                st = pack_reference(game, steps=step, done=(1 if done else 0) | (2 if reward else 0))
            except PackOverflow:
                err = 2
                step -= 1
                actions.append(a)
                masks.append(legal_mask(legal))
                break
            actions.append(a)
            masks.append(legal_mask(legal))
            digests.append(fnv1a64(st.tobytes()))
            rewards.append(reward)
            dones.append(done)
            if record:
                states.append(st)
    # NOTE: `final` carries the done bit only (no reward bit); the golden tests mask byte 19 accordingly
    final = None if err == 2 else pack_reference(game, steps=step, done=(1 if done else 0))
    return dict(seed=seed, init=init, actions=np.array(actions, dtype=np.uint8),
                masks=np.array(masks, dtype=np.uint32).reshape(-1, 5),
                digests=np.array(digests, dtype=np.uint64), states=states, err=err, final=final,
                n_steps=step, done=bool(done), game=game)


def play_expert_game(seed, decks=None, factions=None, max_steps=400, record=True):
    """Both seats play Stormbound.expert_action (games/stormbound.py:563-637), which draws its choices from
    the GAME's stream.  Same tape layout as play_random_game, without masks."""
    game = make_game(seed, decks, factions)
    init = pack_reference(game)
    actions, digests, states = [], [], []
    err = 0
    done = False
    step = 0
    with quiet():
        while not done and step < max_steps:
            try:
                a = game.env.expert_action()
            except Exception:  # noqa: BLE001 -- choice([]) / max([]) inside expert_action itself
                err = 3
                break
            try:
                _obs, reward, done = game.step(a)
            except Exception:  # noqa: BLE001
                err = 1
                actions.append(a)
                break
            step += 1
            try:
                st = pack_reference(game, steps=step, done=(1 if done else 0) | (2 if reward else 0))
            except PackOverflow:
                err = 2
                step -= 1
                actions.append(a)
                break
            actions.append(a)
            digests.append(fnv1a64(st.tobytes()))
            if record:
                states.append(st)
    return dict(seed=seed, init=init, actions=np.array(actions, dtype=np.uint8),
                digests=np.array(digests, dtype=np.uint64), states=states, err=err,
                n_steps=step, done=bool(done), game=game)


# ---------------------------------------------------------------- deck generation (utils.py:26-241)
import random as _pyrandom


class PhiloxPyRandom(_pyrandom.Random):
    """Injected stand-in for the `random` module inside utils.py (sample / choices / random call shapes):
    block = philox(counter=(draw, generation, 0xDEC4, 0), key=game seed).  random.sample itself is CPython's."""

    def __init__(self, seed, generation):
        super().__init__(0)
        self._k = int(seed) & 0xFFFFFFFFFFFFFFFF
        self._gen = int(generation) & M32
        self._draw = 0

    def _block(self):
        out = philox4x32_10(self._draw & M32, self._gen, 0xDEC4, 0, self._k & M32, self._k >> 32)
        self._draw += 1
        return out

    def random(self):
        w = self._block()
        return ((w[0] >> 5) * 67108864 + (w[1] >> 6)) / 9007199254740992.0

    def _randbelow(self, n):
        return (self._block()[0] * n) >> 32


def reference_decks(seed, generation, deck_config):
    """(deck1, deck2) card indices that the UNMODIFIED DeckEvolutionConfig.get_deck_configuration(generation)
    returns when utils.py draws from the injected stream of (seed, generation)."""
    r = ref()
    import utils
    saved = utils.random
    utils.random = PhiloxPyRandom(seed, generation)
    try:
        d1, d2 = deck_config.get_deck_configuration(generation)
    finally:
        utils.random = saved
    return [[r.index[type(c).__name__] for c in d] for d in (d1, d2)]


# ---------------------------------------------------------------- evolution-strategy operators (evo/weights.py, evo/population.py)
import math as _math
import struct as _struct


def det_log(x):
    """sbo_det_log restated in Python floats (IEEE double, no fused operations): identical bits."""
    u = _struct.unpack("<Q", _struct.pack("<d", x))[0]
    e = ((u >> 52) & 0x7FF) - 1022
    m = _struct.unpack("<d", _struct.pack("<Q", (u & 0x000FFFFFFFFFFFFF) | 0x3FE0000000000000))[0]
    if m < 0.70710678118654752440:
        m = m * 2.0
        e -= 1
    z = (m - 1.0) / (m + 1.0)
    z2 = z * z
    p = 1.0 / 27.0
    for k in range(25, 0, -2):
        p = p * z2 + 1.0 / float(k)
    return float(e) * 0.693147180559945309417232 + 2.0 * z * p


class EsStream:
    """Stand-in for `np.random` inside evo/weights.py and evo/population.py: the call shapes of generate_offspring
    (randint, normal(), normal(size), normal(0, sigmas)), of the sigma reset (uniform(lo, hi, n)) and of the
    diversity injection (choice(n, k, replace=False), then 9 normal calls per chosen row) mapped onto per-row
    counter streams philox(counter=(draw, row, tag, generation), key=seed)."""

    def __init__(self, seed):
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.generation = 0
        self.tag, self.row, self.draw = 0xE5, 0, 0
        self.children = 0
        self.resets = 0
        self.queue, self.normal_calls = [], 0

    def begin(self, generation):
        self.generation = int(generation)
        self.children = self.resets = 0
        self.queue, self.normal_calls = [], 0

    def _block(self):
        out = philox4x32_10(self.draw & M32, self.row & M32, self.tag, self.generation & M32, self.seed & M32, self.seed >> 32)
        self.draw += 1
        return out

    @staticmethod
    def _u53(a, b):
        return ((a >> 5) * 67108864 + (b >> 6)) / 9007199254740992.0

    def _gauss(self):
        while True:
            w = self._block()
            u = 2.0 * self._u53(w[0], w[1]) - 1.0
            v = 2.0 * self._u53(w[2], w[3]) - 1.0
            s = u * u + v * v
            if s >= 1.0 or s == 0.0:
                continue
            return u * _math.sqrt(-2.0 * det_log(s) / s)

    def randint(self, low, high):
        self.tag, self.row, self.draw = 0xE5, self.children, 0
        self.children += 1
        return low + ((self._block()[0] * (high - low)) >> 32)

    def normal(self, loc=0.0, scale=1.0, size=None):
        if self.queue:  # diversity injection: 3 mutate calls x 3 normal calls per chosen row
            if self.normal_calls % 9 == 0:
                self.tag, self.row, self.draw = 0xE7, self.queue[self.normal_calls // 9], 0
            self.normal_calls += 1
        if size is None and np.ndim(scale) == 0:
            return loc + scale * self._gauss()
        n = size if size is not None else len(scale)
        z = np.array([self._gauss() for _ in range(n)])
        return loc + np.asarray(scale, dtype=np.float64) * z

    def uniform(self, low, high, size):
        if low == 0 and high == 1:  # WeightVector.__init__ inside parent.copy(): overwritten at once, draws nothing here
            return np.zeros(size)
        self.tag, self.row, self.draw = 0xE6, self.resets, 0
        self.resets += 1
        out = np.empty(size)
        for i in range(size):
            w = self._block()
            out[i] = low + (high - low) * self._u53(w[0], w[1])
        return out

    def choice(self, n, k, replace=False):
        assert not replace
        self.tag, self.row, self.draw = 0xE7, 0xFFFFFFFF, 0
        perm = list(range(n))
        for i in range(n - 1, 0, -1):
            j = (self._block()[0] * (i + 1)) >> 32
            perm[i], perm[j] = perm[j], perm[i]
        self.queue, self.normal_calls = perm[:k], 0
        return np.array(perm[:k])

    def seed_(self, *_a):
        pass


class NumpyWithStream:
    """`np` as seen by the reference's evo modules: numpy, except `.random`."""

    def __init__(self, stream):
        self.random = stream

    def __getattr__(self, name):
        return getattr(np, name)


def reference_population(config_kwargs, seed):
    """(population object, stream): the reference's Population with evo.weights / evo.population drawing from an
    EsStream.  Initial individuals are set by the caller (set_weights / set_sigmas)."""
    ref()
    import contextlib
    import io
    import evo.population as rp
    import evo.weights as rw
    from evo.config import EvolutionaryConfig
    stream = EsStream(seed)
    shim = NumpyWithStream(stream)
    rp.np = shim
    rw.np = shim
    cfg = EvolutionaryConfig(**config_kwargs)
    cfg.seed = None  # Population.__init__ would call np.random.seed
    with contextlib.redirect_stdout(io.StringIO()):
        pop = rp.Population(cfg)
    return pop, stream, rw.WeightVector

