"""torchrun target: wall time of one evo fitness evaluation through the drop-in FitnessEvaluator on N GPUs.
   config 3: pop 256 x 256 games/individual vs one heuristic baseline (65,536 games)
   config 4: pop 1,024 x 16 games vs 5 hall-of-fame vectors + 1 baseline (98,304 games)"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np, torch, torch.distributed as dist
from monsoon_b200.evo import FitnessEvaluator, WeightVector

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))

class Cfg: games_per_pairing = 256; max_turns = 400; seed = 1; num_workers = 4

def vec(w):
    v = WeightVector(10); v.weights = np.asarray(w, dtype=np.float64); return v

def run(name, P, opponents, gpo):
    pop = [vec(w) for w in np.random.RandomState(42).uniform(0, 1, (P, 10))]
    opp = [vec(w) for w in np.random.RandomState(7).uniform(0, 1, (opponents, 10))]
    ev = FitnessEvaluator(Cfg(), device=lr)
    if os.environ.get("EVO_RESIDENT", "1") == "1":  # the weight table of training.Population: resident f64[P, 10]
        pop = torch.from_numpy(np.stack([v.weights for v in pop])).to(torch.device("cuda", lr))
    out = []
    for gen in range(4):
        torch.cuda.synchronize()
        if world > 1: dist.barrier()
        t0 = time.perf_counter()
        fit = ev.evaluate_vs(pop, opp, generation=gen, games_per_opponent=gpo)
        torch.cuda.synchronize()
        if world > 1: dist.barrier()
        out.append(time.perf_counter() - t0)
    if rank == 0:
        games = P * opponents * gpo
        print(json.dumps({"config": name, "n_gpus": world, "games": games, "generation_wall_s": [round(x, 4) for x in out],
                          "games_per_sec": games / min(out), "mean_fitness": float(np.mean(fit)), "fitness_head": [round(x, 4) for x in fit[:4]]}), flush=True)

run("config3 pop256x256 vs baseline", 256, 1, 256)
run("config4 pop1024x16 vs 6 opponents", 1024, 6, 16)
if world > 1:
    dist.destroy_process_group()
