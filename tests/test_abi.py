"""CPU: the C-ABI library loads and exports every symbol include/sb_b200.h declares; layout constants agree."""
import ctypes
import os
import re
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "sb_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sb_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from monsoon_b200 import _lib
    lib = _lib.load()
    names = declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libsb_b200.so does not export %s" % n
        assert n in _lib.SIGNATURES, "ctypes prototype missing for %s" % n
    assert set(_lib.SIGNATURES) == set(names)


def test_introspection_without_gpu():
    from monsoon_b200 import _lib
    from monsoon_b200._card_table import CARDS
    lib = _lib.load()
    assert lib.sb_state_bytes() == 512 and lib.sb_abi_version() >= 1
    assert lib.sb_card_count() == len(CARDS) == 130
    out = (ctypes.c_int32 * 12)()
    for i, c in enumerate(CARDS):
        assert lib.sb_card_info(i, out) == 0
        got = list(out)
        want = [c[k] for k in ("kind", "faction", "cost", "strength", "movement", "trigger", "fixed", "has_ability",
                               "first_type", "types", "obs_id", "has_target")]
        assert got == want, (c["name"], got, want)


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        return
    from monsoon_b200 import _lib
    import pytest
    from monsoon_b200.engine import Engine
    with pytest.raises(_lib.SbError):
        Engine(0)
    h = ctypes.c_void_p()
    assert _lib.load().sb_create(0, ctypes.byref(h)) != 0  # no CPU fallback behind the ABI either


def test_numpy_layout_matches_c_header(tmp_path):
    """oracle/sb_layout.py (numpy dtype) == include/sb_state.h (C struct), field by field."""
    import sb_layout
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "sb_state.h"
#define P(f) printf(#f " %zu\n", offsetof(SbState, f))
#define Q(f) printf("pl." #f " %zu\n", offsetof(SbPlayer, f))
int main(void){ P(seed_lo);P(seed_hi);P(turn);P(draw);P(steps);P(local_order);P(current_order);P(player_sign);P(phase);P(err);P(done);
P(hist_n);P(hist_card);P(hist_owner);P(pl);P(tile);P(ext);
Q(base);Q(max_mana);Q(mana);Q(front_line);Q(flags);Q(n_hand);Q(n_deck);Q(faction);Q(hand_card);Q(hand_cost);Q(hand_flags);Q(deck_card);Q(deck_cost);Q(deck_flags);Q(deck_wn);
printf("tile.card %zu\ntile.flags %zu\ntile.strength %zu\ntile.status %zu\n", offsetof(SbTile,card),offsetof(SbTile,flags),offsetof(SbTile,strength),offsetof(SbTile,status));
printf("sizeof %zu %zu %zu\n", sizeof(SbState), sizeof(SbPlayer), sizeof(SbTile)); return 0; }
'''
    c = tmp_path / "off.c"
    c.write_text(prog)
    exe = tmp_path / "off"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(c)])
    lines = subprocess.check_output([str(exe)], text=True).split("\n")
    offs = {l.split()[0]: int(l.split()[1]) for l in lines if l and not l.startswith("sizeof")}
    for name in sb_layout.STATE_DTYPE.names:
        if name == "pad":
            continue
        assert sb_layout.STATE_DTYPE.fields[name][1] == offs[name], name
    for name in sb_layout.PLAYER_DTYPE.names:
        if name == "pad":
            continue
        assert sb_layout.PLAYER_DTYPE.fields[name][1] == offs["pl." + name], name
    for name in sb_layout.TILE_DTYPE.names:
        assert sb_layout.TILE_DTYPE.fields[name][1] == offs["tile." + name], name
    assert "sizeof 512 104 8" in lines
