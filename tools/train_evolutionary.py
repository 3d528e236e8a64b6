#!/usr/bin/env python3
"""train_evolutionary.py of the reference on the B200 engine: same flags (--config, --resume, --save-config,
--generations, --population, --workers is accepted and ignored), same artefacts in config.results_dir
(training_log.csv, checkpoint_gen*.pkl, final_population.pkl, best_weights.pkl, config.json).

  python tools/train_evolutionary.py --population 512 --generations 5 --games 16 --evaluation vs_hall_of_fame --seed 1
  torchrun --nproc-per-node 8 tools/train_evolutionary.py ...      # games sharded over the ranks (needs --seed)
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
from monsoon_b200.training import EvolutionaryConfig, EvolutionEngine


def main():
    ap = argparse.ArgumentParser(description="Train Stormbound heuristic weights with a (mu + lambda) evolution strategy on the GPU")
    ap.add_argument("--config", help="JSON configuration (nested sections as in the reference's evo_config.json)")
    ap.add_argument("--resume", help="population checkpoint (.pkl, ours or the reference's)")
    ap.add_argument("--save-config", help="write the default configuration to this file and exit")
    ap.add_argument("--workers", type=int, help="accepted for compatibility; the GPU engine has no worker pool")
    ap.add_argument("--generations", type=int)
    ap.add_argument("--population", type=int, help="mu and lambda")
    ap.add_argument("--games", type=int, help="games per pairing")
    ap.add_argument("--max-turns", type=int)
    ap.add_argument("--seed", type=int)
    ap.add_argument("--results-dir")
    ap.add_argument("--evaluation", default="round_robin", choices=["round_robin", "vs_hall_of_fame"])
    args = ap.parse_args()
    if args.save_config:
        with open(args.save_config, "w") as f:
            json.dump(EvolutionaryConfig().to_dict(), f, indent=2)
        return
    cfg = EvolutionaryConfig.from_json(args.config) if args.config else EvolutionaryConfig()
    for k, v in (("generations", args.generations), ("games_per_pairing", args.games), ("max_turns", args.max_turns), ("seed", args.seed),
                 ("results_dir", args.results_dir)):
        if v is not None:
            setattr(cfg, k, v)
    if args.population:
        cfg.mu = cfg.lambda_ = args.population
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        if cfg.seed is None:
            raise SystemExit("--seed is required with several ranks (every rank must build the same population)")
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
        if rank != 0:  # one writer
            cfg.save_logs = False
            cfg.checkpoint_interval = 10 ** 9
            cfg.results_dir = os.path.join(cfg.results_dir, "rank%d" % rank)
    from monsoon_b200.engine import get_engine
    engine = EvolutionEngine(cfg, engine=get_engine(int(os.environ.get("LOCAL_RANK", "0"))), evaluation=args.evaluation)
    if args.resume:
        engine.load_checkpoint(args.resume)
    else:
        engine.initialize()
    if rank == 0:
        with open(os.path.join(cfg.results_dir, "config.json"), "w") as f:
            json.dump(engine.config.to_dict(), f, indent=2)
    t0 = time.time()
    results = engine.run()
    if rank == 0:
        gens = max(results["final_generation"], 1)
        print(json.dumps({"best_fitness": results["best_fitness"], "generations": results["final_generation"], "total_s": round(time.time() - t0, 3),
                          "s_per_generation": round(results["total_time"] / gens, 4), "games": results["eval_stats"]["total_games"],
                          "games_per_sec": round(results["eval_stats"]["games_per_second"], 1), "n_gpus": world, "results_dir": cfg.results_dir}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
