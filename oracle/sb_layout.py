"""TEST INFRASTRUCTURE (oracle side) -- packed game-state layout shared by the harness, the C oracle
and the CUDA engine.  The authoritative C declaration is include/sb_state.h; this numpy dtype must
mirror it byte for byte (tests/test_layout.py checks sizeof/offsets through the C-ABI).

One game = 512 bytes, little endian.  Players are indexed by ORDER (0 = FIRST, 1 = SECOND), never by
local/remote; tiles are stored in the CURRENT board orientation (index = y*4 + x, board.py:17), i.e.
the array is rotated by 180 degrees on every PASS exactly like Board.flip (board.py:94-115).
"""
import numpy as np

SB_STATE_BYTES = 512
N_TILES = 20
HAND_MAX = 4
DECK_MAX = 16
EXT_BYTES = 112

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
assert PLAYER_DTYPE.itemsize == 104, PLAYER_DTYPE.itemsize
assert TILE_DTYPE.itemsize == 8
assert STATE_DTYPE.itemsize == SB_STATE_BYTES, STATE_DTYPE.itemsize

# player flags
PF_REPLACABLE = 1
PF_LEFTMOST_MOVABLE = 2
# hand / deck card flags
CF_FIXED = 1        # Unit.fixedly_forward of the card object (B008 toggles it, cards/b008.py:15-16)
CF_SINGLE_USE = 2   # Card.is_single_use (cards/ua20.py:31)
CF_OBJ = 4          # record is a (former) board instance of B305 (cards/b305.py:41-45)
# tile flags
TF_OWNER = 1        # order of entity.player
TF_STRUCTURE = 2
TF_FIXED = 4
# status field: 6-bit counters (status_effects is a multiset, unit.py:17,239-275), StatusEffect order
ST_BITS = 6
ST_FROZEN, ST_POISONED, ST_CONFUSED, ST_DISABLED, ST_VITALIZED = range(5)

# error codes (state.err); anything != 0 means "the reference raised a Python exception here"
ERR_NONE = 0
ERR_NONE_TARGET = 1     # board.at(p) was None where the card code dereferences it (AttributeError, Q11)
ERR_EMPTY_CHOICE = 2    # RandomState.choice on an empty sequence (ValueError, cards/u017.py:32)
ERR_INDEX = 3           # IndexError / UnboundLocalError (cards/s101.py:22, cards/u310.py:40, hand index)
ERR_OBS_ID = 4          # int(card) ValueError for UP01-03 (card.py:46, Q12)
ERR_UNSUPPORTED = 5     # construct outside the modelled subset (documented deviations)
ERR_OVERFLOW = 6        # a fixed-size pool of the packed/working state overflowed
ERR_DEPTH = 7           # recursion guard


def weight_table(n=1024):
    """w_0 = 1, w_{k+1} = w_k * 1.6 + 100 in IEEE double, exactly as player.py:32,59 computes it."""
    t = np.empty(n, dtype=np.float64)
    w = 1.0
    for i in range(n):
        t[i] = w
        w = w * 1.6 + 100
    return t


def fnv1a64(buf: bytes) -> int:
    h = 0xCBF29CE484222325
    for b in buf:
        h ^= b
        h = (h * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h
