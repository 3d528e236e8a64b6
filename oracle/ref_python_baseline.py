"""BASELINE INFRASTRUCTURE -- the reference's own Python engine, unmodified, multiprocessed over the host cores
(SURVEY 8d "CPU baseline beside it").

Plays whole default-deck games with the uniform-random legal agent of config 2 through games/stormbound.py of the staged copy
(oracle/_ref/monsoon, see stage_ref.py; /root/reference itself in the build container), one game per task over
multiprocessing.Pool(os.cpu_count()), and reports env steps per second.  The injected Philox stream of ref_harness.py is used so
that the games are the SAME games the GPU plays (seed for seed).  Reported beside the GPU numbers, never optimised against."""
import multiprocessing as mp
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))


def reference_dir():
    staged = os.path.join(HERE, "_ref", "monsoon")
    if os.path.exists(os.path.join(staged, ".staged")):
        return staged
    if os.path.isdir("/root/reference"):
        return "/root/reference"
    return None


def _init(ref):
    os.environ["SB_REFERENCE"] = ref
    sys.path[:0] = [HERE, os.path.dirname(HERE)]
    import ref_harness as h
    h.ref()


def _play(seed):
    import ref_harness as h
    t = h.play_random_game(seed, record=False)
    return t["n_steps"]


def run(n_games=256, procs=None, seed0=30_000_000):
    """returns dict(value = env steps / s over all processes, cores, games, steps, seconds) or None without a reference"""
    ref = reference_dir()
    if ref is None:
        return None
    procs = procs or os.cpu_count() or 1
    with mp.get_context("spawn").Pool(procs, initializer=_init, initargs=(ref,)) as pool:
        pool.map(_play, range(seed0, seed0 + procs))  # warm-up: imports, card classes
        t0 = time.perf_counter()
        steps = pool.map(_play, range(seed0 + 1000, seed0 + 1000 + n_games), chunksize=max(1, n_games // (procs * 4)))
        dt = time.perf_counter() - t0
    total = int(sum(steps))
    return {"value": total / dt, "unit": "env_steps/s", "cores": procs, "kind": "reference", "games": n_games, "steps": total, "seconds": dt,
            "sample": "%d default-deck games (%d env steps) of the reference's Python engine (games/stormbound.py, uniform-random legal "
                      "agent, injected Philox stream), multiprocessing.Pool(%d), %.1f s wall" % (n_games, total, procs, dt)}


if __name__ == "__main__":
    print(run(int(sys.argv[1]) if len(sys.argv) > 1 else 64))
