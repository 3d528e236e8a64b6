"""GPU box: BASELINE config 5 -- 1 M games per generation, every game with its own random 12-card decks
(generate_random_deck semantics, factions round-robin), drawn on the device.  Reports the deck kernel, the
random-agent rollout and the heuristic evaluation (P individuals as FIRST vs one baseline vector), plus the
engine status of the finished games (the reference's own exceptions for some cards)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np, torch
from monsoon_b200.engine import Engine
eng = Engine(0)
dev = eng.device
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
P = 1024
CH = 262144
gen = 7
seeds = torch.arange(N, dtype=torch.int64, device=dev)
fac = torch.stack([1 + seeds % 4, 1 + (seeds // 4) % 4], dim=1).to(torch.uint8).contiguous()


def timed(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize(); return out, time.perf_counter() - t0


(decks, fout), t = timed(lambda: eng.generate_decks(seeds, gen, 3, factions=fac))
(decks, fout), t = timed(lambda: eng.generate_decks(seeds, gen, 3, factions=fac))
print("deck generation: %d games x 2 decks in %.2f ms (%.1f M decks/s)" % (N, t * 1e3, 2 * N / t / 1e6))

# random agents
def random_pass():
    tot = 0
    errs = np.zeros(8, dtype=np.int64)
    for c0 in range(0, N, CH):
        st = eng.reset(seeds[c0:c0 + CH], decks[c0:c0 + CH], fout[c0:c0 + CH])
        steps = eng.rollout_random(st, 400)
        tot += int(steps.sum())
        errs += np.bincount(st[:, 18].cpu().numpy(), minlength=8)[:8]
    return tot, errs
random_pass()
(tot, errs), t = timed(random_pass)
print("random agents: %.2f s  %.1f M env-steps/s  %.2f M games/s  status counts (code 0..7) %s" % (t, tot / t / 1e6, N / t / 1e6, errs.tolist()))

# heuristic evaluation
w = torch.from_numpy(np.concatenate([np.random.RandomState(42).uniform(0, 1, (P, 10)), np.random.RandomState(7).uniform(0, 1, (1, 10))])).to(dev)
def heur_pass(n):
    counts = torch.zeros((P + 1, 3), dtype=torch.int32, device=dev)
    tot = 0
    for c0 in range(0, n, CH):
        c1 = min(n, c0 + CH)
        i1 = (seeds[c0:c1] % P).to(torch.int32)
        i2 = torch.full((c1 - c0,), P, dtype=torch.int32, device=dev)
        st = eng.reset(seeds[c0:c1], decks[c0:c1], fout[c0:c1])
        res, steps = eng.rollout_heuristic(st, w, w, i1, i2, max_steps=400)
        eng.accumulate_fitness(res, i1, counts)
        tot += int(steps.sum())
    return tot, counts.cpu().numpy()
heur_pass(CH)
(tot, counts), t = timed(lambda: heur_pass(N))
print("heuristic evaluation: %.2f s  %.1f k games/s  %.2f M env-steps/s  W/D/L of the population %s" % (
    t, N / t / 1e3, tot / t / 1e6, counts[:P].sum(axis=0).tolist()))
