"""Host bridge between the 512-byte packed game record (include/sb_state.h) and readable Python values -- the
`sb_pack / sb_unpack` debug bridge of SURVEY 8(b).

The record replaces the reference's object graph (board.py:16-28 Board, player.py:13-37 Player, unit.py:8-23 Unit,
structure.py:8-16, games/stormbound.py:293-304 Stormbound); `unpack_state` turns one record into a dict with the
reference's attribute names (players by ORDER, `board[y][x]` in the current orientation like Board.board), `pack_state`
is its inverse.  Pure numpy on host bytes: it moves no game forward, so it needs neither the GPU nor the library."""
import numpy as np

from ._card_table import CARDS

STATE_BYTES, N_TILES, HAND_MAX, DECK_MAX, EXT_BYTES = 512, 20, 4, 16, 112
STATUS_NAMES = ("FROZEN", "POISONED", "CONFUSED", "DISABLED", "VITALIZED")  # enums.py:77-82, 6-bit counters each
ERR_NAMES = ("", "NONE_TARGET", "EMPTY_CHOICE", "INDEX", "OBS_ID", "UNSUPPORTED", "OVERFLOW", "DEPTH")

PLAYER_DTYPE = np.dtype([
    ("base", "<i2"), ("max_mana", "<i2"), ("mana", "<i2"), ("front_line", "i1"), ("flags", "u1"),
    ("n_hand", "u1"), ("n_deck", "u1"), ("faction", "u1"), ("pad", "u1"),
    ("hand_card", "u1", (HAND_MAX,)), ("hand_cost", "i1", (HAND_MAX,)), ("hand_flags", "u1", (HAND_MAX,)),
    ("deck_card", "u1", (DECK_MAX,)), ("deck_cost", "i1", (DECK_MAX,)), ("deck_flags", "u1", (DECK_MAX,)),
    ("deck_wn", "<u2", (DECK_MAX,)),
])
TILE_DTYPE = np.dtype([("card", "u1"), ("flags", "u1"), ("strength", "<i2"), ("status", "<u4")])
STATE_DTYPE = np.dtype([
    ("seed_lo", "<u4"), ("seed_hi", "<u4"), ("turn", "<u2"), ("draw", "<u2"), ("steps", "<u2"),
    ("local_order", "u1"), ("current_order", "u1"), ("player_sign", "i1"), ("phase", "u1"),
    ("err", "u1"), ("done", "u1"), ("hist_n", "u1"), ("hist_card", "u1", (4,)), ("hist_owner", "u1", (4,)),
    ("pad", "u1", (3,)),
    ("pl", PLAYER_DTYPE, (2,)),
    ("tile", TILE_DTYPE, (N_TILES,)),
    ("ext", "u1", (EXT_BYTES,)),
])
assert PLAYER_DTYPE.itemsize == 104 and TILE_DTYPE.itemsize == 8 and STATE_DTYPE.itemsize == STATE_BYTES

_INDEX = {r["name"]: i for i, r in enumerate(CARDS)}


def as_records(states):
    """u8[n,512] (numpy or a CPU/GPU torch tensor) -> structured array of n records (a view for numpy input)"""
    if hasattr(states, "detach"):
        states = states.detach().cpu().numpy()
    a = np.ascontiguousarray(states, dtype=np.uint8).reshape(-1, STATE_BYTES)
    return a.view(STATE_DTYPE).reshape(-1)


def _card_name(i):
    return CARDS[int(i)]["name"] if 0 < int(i) < len(CARDS) else None


def _card_records(cards, costs, flags, n, weights=None):
    out = []
    for k in range(int(n)):
        c = {"card": _card_name(cards[k]), "cost": int(costs[k]), "fixedly_forward": bool(flags[k] & 1), "is_single_use": bool(flags[k] & 2),
             "board_instance": bool(flags[k] & 4)}
        if weights is not None:
            c["weight_exponent"] = int(weights[k])  # weight = f^n(1), f(w) = w * 1.6 + 100 (player.py:32,59)
        out.append(c)
    return out


def unpack_state(record):
    """One record (512 bytes, a row of the states tensor, or an element of as_records()) -> dict."""
    r = record if isinstance(record, np.void) else as_records(record)[0]
    players = []
    for o in (0, 1):
        p = r["pl"][o]
        players.append({"order": o, "strength": int(p["base"]), "max_mana": int(p["max_mana"]), "current_mana": int(p["mana"]),
                        "front_line": int(p["front_line"]), "replacable": bool(p["flags"] & 1), "leftmost_movable": bool(p["flags"] & 2),
                        "faction": int(p["faction"]),
                        "hand": _card_records(p["hand_card"], p["hand_cost"], p["hand_flags"], min(int(p["n_hand"]), HAND_MAX)),
                        "deck": _card_records(p["deck_card"], p["deck_cost"], p["deck_flags"], min(int(p["n_deck"]), DECK_MAX), p["deck_wn"])})
    board = [[None] * 4 for _ in range(5)]
    for t in range(N_TILES):
        tl = r["tile"][t]
        if tl["card"]:
            st = int(tl["status"])
            board[t >> 2][t & 3] = {"card": _card_name(tl["card"]), "owner": int(tl["flags"] & 1), "is_structure": bool(tl["flags"] & 2),
                                    "fixedly_forward": bool(tl["flags"] & 4), "strength": int(tl["strength"]),
                                    "status_effects": {n: (st >> (6 * k)) & 63 for k, n in enumerate(STATUS_NAMES) if (st >> (6 * k)) & 63}}
    return {"seed": int(r["seed_lo"]) | (int(r["seed_hi"]) << 32), "turn": int(r["turn"]), "draw": int(r["draw"]), "steps": int(r["steps"]),
            "local_order": int(r["local_order"]), "current_order": int(r["current_order"]), "player_sign": int(r["player_sign"]),
            "phase": int(r["phase"]), "err": int(r["err"]), "err_name": ERR_NAMES[int(r["err"])] if int(r["err"]) < len(ERR_NAMES) else "?",
            "done": bool(r["done"] & 1), "reward": bool(r["done"] & 2),
            "history": [{"card": _card_name(r["hist_card"][k]), "owner": int(r["hist_owner"][k])} for k in range(min(int(r["hist_n"]), 4))],
            "players": players, "board": board, "ext": bytes(r["ext"])}


def pack_state(d):
    """Inverse of unpack_state: dict -> u8[512]."""
    r = np.zeros((), dtype=STATE_DTYPE)
    r["seed_lo"], r["seed_hi"] = d["seed"] & 0xFFFFFFFF, d["seed"] >> 32
    for k in ("turn", "draw", "steps", "local_order", "current_order", "player_sign", "phase", "err"):
        r[k] = d[k]
    r["done"] = (1 if d["done"] else 0) | (2 if d["reward"] else 0)
    r["hist_n"] = len(d["history"])
    for k, h in enumerate(d["history"]):
        r["hist_card"][k], r["hist_owner"][k] = _INDEX[h["card"]], h["owner"]

    def flags_of(c):
        return (1 if c["fixedly_forward"] else 0) | (2 if c["is_single_use"] else 0) | (4 if c["board_instance"] else 0)
    for o, p in enumerate(d["players"]):
        q = r["pl"][o]
        q["base"], q["max_mana"], q["mana"], q["front_line"] = p["strength"], p["max_mana"], p["current_mana"], p["front_line"]
        q["flags"] = (1 if p["replacable"] else 0) | (2 if p["leftmost_movable"] else 0)
        q["faction"], q["n_hand"], q["n_deck"] = p["faction"], len(p["hand"]), len(p["deck"])
        for k, c in enumerate(p["hand"]):
            q["hand_card"][k], q["hand_cost"][k], q["hand_flags"][k] = _INDEX[c["card"]], c["cost"], flags_of(c)
        for k, c in enumerate(p["deck"]):
            q["deck_card"][k], q["deck_cost"][k], q["deck_flags"][k], q["deck_wn"][k] = _INDEX[c["card"]], c["cost"], flags_of(c), c["weight_exponent"]
    for y in range(5):
        for x in range(4):
            e = d["board"][y][x]
            if e is None:
                continue
            tl = r["tile"][y * 4 + x]
            tl["card"], tl["strength"] = _INDEX[e["card"]], e["strength"]
            tl["flags"] = e["owner"] | (2 if e["is_structure"] else 0) | (4 if e["fixedly_forward"] else 0)
            st = 0
            for k, n in enumerate(STATUS_NAMES):
                st |= (e["status_effects"].get(n, 0) & 63) << (6 * k)
            tl["status"] = st
    r["ext"] = np.frombuffer(d["ext"], dtype=np.uint8)
    return np.frombuffer(r.tobytes(), dtype=np.uint8).copy()


def render_state(record):
    """A few lines of text: the board from the local player's point of view, bases, mana, hands (debugging aid)."""
    d = unpack_state(record)
    lines = ["turn %d step %d  to play: %s  local order %d%s" % (d["turn"], d["steps"], "FIRST" if d["player_sign"] == 1 else "SECOND", d["local_order"],
                                                               ("  ERR " + d["err_name"]) if d["err"] else "")]
    for y in range(5):
        row = []
        for x in range(4):
            e = d["board"][y][x]
            row.append("   .    " if e is None else "%s%s%-3d " % (e["card"][:4], "+" if e["owner"] == d["local_order"] else "-", e["strength"]))
        lines.append(" ".join(row))
    for p in d["players"]:
        lines.append("order %d: base %d mana %d/%d hand %s deck %d" % (p["order"], p["strength"], p["current_mana"], p["max_mana"],
                                                                      [c["card"] for c in p["hand"]], len(p["deck"])))
    return "\n".join(lines)


__all__ = ["STATE_DTYPE", "PLAYER_DTYPE", "TILE_DTYPE", "as_records", "unpack_state", "pack_state", "render_state"]
