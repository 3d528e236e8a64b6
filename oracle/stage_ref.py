"""TEST / BASELINE INFRASTRUCTURE -- stage the reference's own Python engine for the GPU box.

The reference (dvrp0/Monsoon) is pure Python: nothing to compile, and /root/reference does not exist on the GPU box.  BASELINE.md
section 3 asks for its engine, unmodified, multiprocessed over the box's host cores as the CPU baseline beside the GPU numbers.
This recipe copies the files of the hot path (rules engine, cards, games/, evo/) from the read-only checkout into
oracle/_ref/monsoon/ -- git-ignored, so no reference source ever enters the history, but not gpurun-ignored, so the copy
travels with the snapshot like a built .so.  Run by __graft_entry__.build() whenever /root/reference is present.
Only bench.py's `cpu_baseline_reference` leg (oracle/ref_python_baseline.py) executes the staged copy."""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("SB_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref", "monsoon")
FLAT = ["board.py", "card.py", "const.py", "enums.py", "player.py", "point.py", "spell.py", "structure.py", "target.py", "test.py",
        "unit.py", "utils.py", "actions.txt", "cards.json", "LICENSE"]
DIRS = ["cards", "games", "evo"]


def stage(force=False):
    if not os.path.isdir(SRC):
        return None
    stamp = os.path.join(DST, ".staged")
    if os.path.exists(stamp) and not force:
        return DST
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    for f in FLAT:
        if os.path.exists(os.path.join(SRC, f)):
            shutil.copy2(os.path.join(SRC, f), os.path.join(DST, f))
    for d in DIRS:
        shutil.copytree(os.path.join(SRC, d), os.path.join(DST, d),
                        ignore=lambda _p, names: [n for n in names if not (n.endswith(".py") or os.path.isdir(os.path.join(_p, n))) or n == "__pycache__"])
    with open(stamp, "w") as f:
        f.write("staged from %s (unmodified files; see oracle/stage_ref.py)\n" % SRC)
    return DST


if __name__ == "__main__":
    print(stage(force="--force" in sys.argv))
