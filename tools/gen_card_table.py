#!/usr/bin/env python3
"""Generate the card table from the reference's card CLASSES (not cards.json, which is display
metadata and disagrees with the classes on 22 cards -- SURVEY.md fact 5).

Runs only in the build container (needs /root/reference).  Instantiates every class exported by
/root/reference/cards/__init__.py:5-13, reads the stats set by its ``super().__init__`` call
(unit.py:8-23, structure.py:8-16, spell.py:8-12), the numeric ``ability_*`` attributes and the
observation id ``int(card)`` (card.py:25-46), and writes three *data* files that are committed:

  monsoon_b200/_card_table.py          python dicts used by the host layer
  monsoon_b200/csrc/sb_card_table.inc  C initialiser list included by the CUDA engine
  oracle/sb_card_table.inc             identical initialiser list included by the C oracle

Card index: 0 = empty, 1..112 = classes in sorted name order, 113..128 = unit tokens of
UnitType 0..15 (board.py:298-311), 129 = token structure (board.py:313-322).
"""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path[:0] = [os.path.join(REPO, "oracle", "refshim"), REF]
os.chdir(REF)

import cards as ref_cards  # noqa: E402
from card import Card  # noqa: E402
from unit import Unit  # noqa: E402
from structure import Structure  # noqa: E402
from spell import Spell  # noqa: E402
from enums import TriggerType, UnitType  # noqa: E402

KIND_UNIT, KIND_STRUCTURE, KIND_SPELL = 0, 1, 2
TRIG_NONE = 255
N_PARAMS = 4


def obs_id(card):
    try:
        return int(card)
    except ValueError:  # UP01-03: "2p01" is not hex (Q12)
        return -32768


def mask(values):
    m = 0
    for v in values or []:
        m |= 1 << int(v)
    return m


def target_fields(t):
    if t is None:
        return dict(has_target=0, t_kind=0, t_side=0, t_types=0, t_xtypes=0, t_status=0, t_xstatus=0,
                    t_limit=-1, t_nonhero=0, t_base=0)
    return dict(has_target=1, t_kind=int(t.kind), t_side=int(t.side), t_types=mask(t.unit_types),
                t_xtypes=mask(t.exclude_unit_types), t_status=mask(t.status_effects),
                t_xstatus=mask(t.exclude_status_effects),
                t_limit=-1 if t.strength_limit is None else int(t.strength_limit),
                t_nonhero=int(bool(t.non_hero)), t_base=int(bool(t.include_base)))


def collect():
    names = sorted(
        n for n in dir(ref_cards)
        if len(n) == 4 and isinstance(getattr(ref_cards, n), type)
        and issubclass(getattr(ref_cards, n), Card)
        and getattr(ref_cards, n) not in (Unit, Structure, Spell, Card))
    rows = [dict(name="NONE", kind=0, faction=0, cost=0, strength=0, movement=0, trigger=TRIG_NONE,
                 fixed=0, types=0, first_type=0, has_ability=0, obs_id=-1, params=[0] * N_PARAMS,
                 param_names=[], **target_fields(None))]
    for n in names:
        cls = getattr(ref_cards, n)
        c = cls()
        has_ability = int("activate_ability" in cls.__dict__)
        pnames = sorted(k for k, v in vars(c).items()
                        if (k.startswith("ability_") or k in ("damage", "original_cost"))
                        and isinstance(v, int) and not isinstance(v, bool))
        params = [int(getattr(c, k)) for k in pnames][:N_PARAMS]
        assert len(pnames) <= N_PARAMS, (n, pnames)
        params += [0] * (N_PARAMS - len(params))
        if isinstance(c, Unit):
            row = dict(kind=KIND_UNIT, strength=c.strength, movement=c.movement,
                       trigger=TRIG_NONE if c.trigger is None else int(c.trigger),
                       fixed=int(c.fixedly_forward), types=mask(c.unit_types),
                       first_type=int(c.unit_types[0]), **target_fields(None))
        elif isinstance(c, Structure):
            assert len(c.triggers) == 1
            row = dict(kind=KIND_STRUCTURE, strength=c.strength, movement=0,
                       trigger=int(c.triggers[0]), fixed=0, types=0, first_type=0,
                       **target_fields(None))
        else:
            row = dict(kind=KIND_SPELL, strength=0, movement=0, trigger=TRIG_NONE, fixed=0, types=0,
                       first_type=0, **target_fields(c.required_targets))
        row.update(name=n, faction=int(c.faction), cost=int(c.cost), has_ability=has_ability,
                   obs_id=obs_id(c), params=params, param_names=pnames)
        assert c.card_id == n.lower()
        rows.append(row)
    assert len(rows) == 113
    for t in UnitType:  # tokens: Unit(NEUTRAL, [t], 0, strength, 1), card_id "f" + str(t).zfill(3)
        cid = "f" + str(int(t)).zfill(3)
        rows.append(dict(name="TOK%02d" % int(t), kind=KIND_UNIT, faction=0, cost=0, strength=0, movement=1,
                         trigger=TRIG_NONE, fixed=0, types=1 << int(t), first_type=int(t), has_ability=0,
                         obs_id=int("4" + cid[1] + cid[2:], 16), params=[0] * N_PARAMS, param_names=[],
                         **target_fields(None)))
    # token structure: Structure(NEUTRAL, 0, strength) with default triggers [TURN_START], card_id "b001"
    rows.append(dict(name="TOKB", kind=KIND_STRUCTURE, faction=0, cost=0, strength=0, movement=0,
                     trigger=int(TriggerType.TURN_START), fixed=0, types=0, first_type=0, has_ability=0,
                     obs_id=int("0001", 16), params=[0] * N_PARAMS, param_names=[], **target_fields(None)))
    return rows


C_FIELDS = ["kind", "faction", "cost", "strength", "movement", "trigger", "fixed", "has_ability",
            "first_type", "types", "obs_id", "has_target", "t_kind", "t_side", "t_types", "t_xtypes",
            "t_status", "t_xstatus", "t_limit", "t_nonhero", "t_base"]


def emit_c(rows, path, who):
    with open(path, "w") as f:
        f.write("/* GENERATED by tools/gen_card_table.py from the reference card classes -- do not edit.\n"
                " * %s\n * fields: %s, params[4]\n */\n" % (who, ", ".join(C_FIELDS)))
        for i, r in enumerate(rows):
            vals = ", ".join(str(r[k]) for k in C_FIELDS)
            f.write("/* %3d %-5s %-40s */ { %s, { %s } },\n" % (
                i, r["name"], " ".join(r["param_names"]), vals, ", ".join(str(p) for p in r["params"])))


def emit_enum(rows, path):
    with open(path, "w") as f:
        f.write("/* GENERATED by tools/gen_card_table.py -- card indices. */\n")
        for i, r in enumerate(rows):
            f.write("#define SBC_%s %d\n" % (r["name"], i))
        f.write("#define SBC_COUNT %d\n#define SBC_TOKEN_UNIT0 113\n#define SBC_TOKEN_STRUCTURE 129\n" % len(rows))


def main():
    rows = collect()
    with open(os.path.join(REPO, "monsoon_b200", "_card_table.py"), "w") as f:
        f.write('"""GENERATED by tools/gen_card_table.py from the reference card classes -- do not edit."""\n')
        f.write("CARD_FIELDS = %r\n" % (C_FIELDS + ["params"],))
        f.write("CARDS = ")
        json.dump(rows, f, indent=0)
        f.write("\n")
    for d, who in (("monsoon_b200/csrc", "CUDA engine copy"), ("oracle", "C oracle copy")):
        emit_c(rows, os.path.join(REPO, d, "sb_card_table.inc"), who)
        emit_enum(rows, os.path.join(REPO, d, "sb_card_ids.h"))
    print("wrote", len(rows), "cards")


if __name__ == "__main__":
    main()
