"""torchrun target (or plain python for world size 1): BASELINE config 3 through the drop-in FitnessEvaluator -- population
256 x 256 games per individual as FIRST against one baseline vector (65,536 heuristic-agent games, sharded over the ranks
by game index, weights broadcast and the i32[P,3] win / draw / loss counts all-reduced over NCCL).  Rank 0 writes the counts;
tests/test_gpu_multi.py compares them for world sizes 1, 2, 4, 8: they must be IDENTICAL.
Also builds the training driver WITHOUT an explicit engine on every rank and checks that it lands on the rank's own GPU.
  python tools/nccl_parity.py out.npy [P] [games]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np, torch, torch.distributed as dist
from monsoon_b200.evo import FitnessEvaluator, WeightVector

out = sys.argv[1]
P = int(sys.argv[2]) if len(sys.argv) > 2 else 256
G = int(sys.argv[3]) if len(sys.argv) > 3 else 256
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))


class Cfg:
    games_per_pairing = G; max_turns = 400; seed = 1; num_workers = 4


def vec(w):
    v = WeightVector(10); v.weights = np.asarray(w, dtype=np.float64); return v


pop = [vec(w) for w in np.random.RandomState(42).uniform(0, 1, (P, 10))]
opp = [vec(np.random.RandomState(7).uniform(0, 1, 10))]
ev = FitnessEvaluator(Cfg())  # no device, no engine: must pick this rank's GPU
fit = ev.evaluate_vs(pop, opp, generation=0, games_per_opponent=G)
assert ev._engine().device.index == lr, (ev._engine().device, lr)
# the training driver without an explicit engine lands on the rank's GPU as well
from monsoon_b200.training import EvolutionaryConfig, EvolutionEngine
cfg = EvolutionaryConfig()
cfg.mu, cfg.lambda_, cfg.results_dir, cfg.save_logs = 8, 8, os.path.join(os.path.dirname(os.path.abspath(out)), "nccl_parity_run_%d" % rank), False
drv = EvolutionEngine(cfg)
drv.initialize()
assert drv.population.eng.device.index == lr and drv.population.w.device.index == lr, (drv.population.eng.device, lr)
if rank == 0:
    np.save(out, ev.last_counts)
    print(json.dumps({"world": world, "P": P, "games_per_individual": G, "mean_fitness": float(np.mean(fit)),
                      "counts_sum": ev.last_counts.sum(axis=0).tolist(), "aborted": list(ev.last_aborted)}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
