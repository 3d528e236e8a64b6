"""Host-side mirror of the reference's evolution driver for the rows SURVEY 8(f) marks next:

  f1  (mu + lambda) operators -- Population.generate_offspring / select_from_combined and WeightVector.mutate
      (evo/population.py:75-176, evo/weights.py:20-40) run as CUDA kernels on a population that stays resident
      in HBM (sb_es_* in include/sb_b200.h);
  f4  checkpoint / log writers -- population pickles interchangeable with the reference's
      Population.save_population / load_population (evo/population.py:281-310) and the training_log.csv rows of
      EvolutionEngine._save_generation_log (evo/evolution.py:131-144).

Same class and method names as the reference (EvolutionaryConfig, Population, EvolutionEngine).  Randomness of the
operators comes from per-row counter streams keyed by config.seed (see sb_es.cuh); population initialisation is
host numpy with the reference's draw order, so `np.random.seed(config.seed)` gives the reference's initial population.
"""
import dataclasses
import io
import json
import os
import pickle
import sys
import time
import types
from datetime import datetime
from typing import Optional

import numpy as np
import torch

from .engine import get_engine
from .evo import FEATURE_NAMES, FitnessEvaluator, WeightVector


@dataclasses.dataclass
class EvolutionaryConfig:
    """evo/config.py:10-146: same fields and defaults (they are part of the checkpoint format)."""
    mu: int = 10
    lambda_: int = 10
    generations: int = 100
    games_per_pairing: int = 20
    deck_configs: int = 3
    tau: float = 0.1
    tau_prime: float = 0.01
    min_sigma: float = 1e-5
    initial_sigma: float = 0.1
    sigma_reset_threshold: float = 1e-4
    sigma_boost_factor: float = 2.0
    fitness_stagnation_gens: int = 10
    max_turns: int = 100
    num_workers: int = 128
    timeout_seconds: int = 30
    checkpoint_interval: int = 10
    save_best_n: int = 5
    log_level: str = "INFO"
    save_logs: bool = True
    save_generation_details: bool = True
    track_weight_evolution: bool = True
    results_dir: str = "results/evolutionary2"
    min_generations: int = 50
    fitness_plateau_threshold: float = 0.001
    plateau_generations: int = 25
    seed: Optional[int] = None

    _JSON_SECTIONS = {
        "population": ("mu", "lambda_"), "evolution": ("generations",), "evaluation": ("games_per_pairing", "deck_configs"),
        "mutation": ("tau", "tau_prime", "min_sigma", "initial_sigma", "sigma_reset_threshold", "sigma_boost_factor",
                     "fitness_stagnation_gens"),
        "simulation": ("max_turns", "num_workers", "timeout_seconds"),
        "checkpointing": ("checkpoint_interval", "save_best_n", "results_dir"),
        "logging": ("log_level", "save_generation_details", "track_weight_evolution"),
        "convergence_criteria": ("min_generations", "fitness_plateau_threshold", "plateau_generations"),
    }

    def __post_init__(self):
        for name, text in (("mu", "Parent population size (mu)"), ("lambda_", "Offspring size (lambda)"),
                           ("generations", "Number of generations"), ("games_per_pairing", "Games per pairing"),
                           ("min_sigma", "Minimum sigma"), ("num_workers", "Number of workers")):
            if getattr(self, name) <= 0:
                raise ValueError("%s must be positive" % text)
        if self.tau <= 0 or self.tau_prime <= 0:
            raise ValueError("Mutation parameters (tau, tau_prime) must be positive")

    @classmethod
    def from_json(cls, json_path):  # nested sections of evo/config.py:74-138
        with open(json_path) as f:
            data = json.load(f)
        flat = {}
        for section, keys in cls._JSON_SECTIONS.items():
            if section in data:
                for k in keys:
                    flat[k] = data[section][k]
        return cls(**flat)

    @classmethod
    def from_dict(cls, config_dict):
        return cls(**config_dict)

    def to_dict(self):
        return dataclasses.asdict(self)


# ---------------------------------------------------------------- checkpoint format (evo/population.py:281-310)
_REF_CLASSES = {("evo.weights", "WeightVector"): WeightVector, ("evo.config", "EvolutionaryConfig"): EvolutionaryConfig}


class _Unpickler(pickle.Unpickler):
    """Reads pickles written by the reference: its classes resolve to the mirrors here."""

    def find_class(self, module, name):
        return _REF_CLASSES.get((module, name)) or super().find_class(module, name)


def load_checkpoint(path_or_file):
    f = open(path_or_file, "rb") if isinstance(path_or_file, (str, os.PathLike)) else path_or_file
    try:
        return _Unpickler(f).load()
    finally:
        if f is not path_or_file:
            f.close()


def dump_checkpoint(obj, path_or_file):
    """Writes `obj` (dict / list / WeightVector / EvolutionaryConfig graph) so that the REFERENCE can pickle.load it:
    mirror instances are emitted as evo.weights.WeightVector / evo.config.EvolutionaryConfig objects."""
    stubs = {}
    for (module, name), mirror in _REF_CLASSES.items():
        stubs[(module, name)] = type(name, (), {"__module__": module, "__qualname__": name})

    def convert(x):
        for key, mirror in _REF_CLASSES.items():
            if type(x) is mirror:
                y = stubs[key].__new__(stubs[key])
                y.__dict__.update({k: convert(v) for k, v in vars(x).items()})
                return y
        if isinstance(x, dict):
            return {k: convert(v) for k, v in x.items()}
        if isinstance(x, (list, tuple)):
            return type(x)(convert(v) for v in x)
        return x

    saved = {m: sys.modules.get(m) for m in ("evo", "evo.weights", "evo.config")}
    try:  # pickle stores classes by reference and checks that module.name resolves to the very class object
        pkg = types.ModuleType("evo")
        pkg.__path__ = []
        sys.modules["evo"] = pkg
        for (module, name), cls in stubs.items():
            mod = types.ModuleType(module)
            setattr(mod, name, cls)
            sys.modules[module] = mod
            setattr(pkg, module.split(".")[1], mod)
        data = pickle.dumps(convert(obj), protocol=4)
    finally:
        for m, old in saved.items():
            if old is None:
                sys.modules.pop(m, None)
            else:
                sys.modules[m] = old
    if isinstance(path_or_file, (str, os.PathLike)):
        with open(path_or_file, "wb") as f:
            f.write(data)
    else:
        path_or_file.write(data)


LOG_HEADER = "generation,time,best_fitness,mean_fitness,std_fitness,diversity,avg_sigma,games_per_sec\n"


def append_generation_log(log_file, stats, eval_stats, generation_time):
    """One training_log.csv row in the reference's format (evo/evolution.py:131-144)."""
    if not os.path.exists(log_file):
        with open(log_file, "w") as f:
            f.write(LOG_HEADER)
    with open(log_file, "a") as f:
        f.write("%d,%.2f,%.6f,%.6f,%.6f,%.6f,%.6f,%.1f\n" % (
            stats["generation"], generation_time, stats["best_fitness"], stats["mean_fitness"], stats["std_fitness"],
            stats["diversity"], stats["avg_mutation_strength"], eval_stats["games_per_second"]))


# ---------------------------------------------------------------- population resident on the device
class Population:
    """evo/population.py:15-323.  weights / sigmas live in two f64 [mu + lambda, n] device tensors; rows [0, mu)
    are the parents, rows [mu, mu + lambda) the latest offspring.  `individuals` materialises WeightVector objects
    (host copies) for the callers that want them (FitnessEvaluator, checkpoints)."""

    def __init__(self, config, engine=None, device=None):
        self.config = config
        self.eng = engine or get_engine(device)  # None = the current CUDA device (LOCAL_RANK under torchrun)
        self.fitness_scores = []
        self.generation = 0
        self.w = self.s = None
        self.feature_count = 0
        self.last_events = []
        if config.seed is not None:
            np.random.seed(config.seed)

    # -- state <-> objects
    def _alloc(self, n):
        rows = self.config.mu + self.config.lambda_
        self.feature_count = n
        self.w = torch.zeros((rows, n), dtype=torch.float64, device=self.eng.device)
        self.s = torch.zeros((rows, n), dtype=torch.float64, device=self.eng.device)

    @staticmethod
    def _vectors(w, s):
        out = []
        for wi, si in zip(w, s):
            v = WeightVector.__new__(WeightVector)
            v.weights, v.sigmas, v.size = wi.copy(), si.copy(), len(wi)
            out.append(v)
        return out

    @property
    def individuals(self):
        mu = self.config.mu
        return self._vectors(self.w[:mu].cpu().numpy(), self.s[:mu].cpu().numpy())

    @individuals.setter
    def individuals(self, vectors):
        vectors = list(vectors)
        if self.w is None or self.feature_count != len(vectors[0].weights):
            self._alloc(len(vectors[0].weights))
        n = min(len(vectors), self.config.mu)
        self.w[:n] = torch.from_numpy(np.stack([np.asarray(v.weights, dtype=np.float64) for v in vectors[:n]])).to(self.eng.device)
        self.s[:n] = torch.from_numpy(np.stack([np.asarray(v.sigmas, dtype=np.float64) for v in vectors[:n]])).to(self.eng.device)

    def initialize_population(self, feature_count):
        """evo/population.py:28-68: three diversity groups; numpy draws in the reference's order."""
        mu, third = self.config.mu, self.config.mu // 3
        w = np.empty((mu, feature_count))
        s = np.empty((mu, feature_count))
        for i in range(mu):
            np.random.uniform(0, 1, feature_count)  # the WeightVector constructor's draw
            if i < third:
                row = np.random.uniform(0.2, 0.8, feature_count)
            elif i < 2 * third:
                row = np.random.choice([0.0, 1.0], feature_count, p=[0.3, 0.7])
                row = np.clip(row + np.random.normal(0, 0.1, feature_count), 0, 1)
            else:
                row = np.random.uniform(0.0, 1.0, feature_count)
            w[i] = np.clip(row, 0, 1)
            spread = np.random.uniform(0.5, 2.0)
            s[i] = np.maximum(np.full(feature_count, self.config.initial_sigma * spread) * np.random.uniform(0.8, 1.2, feature_count), 1e-10)
        self._alloc(feature_count)
        self.w[:mu] = torch.from_numpy(w).to(self.eng.device)
        self.s[:mu] = torch.from_numpy(s).to(self.eng.device)
        self.fitness_scores = [0.0] * mu

    def get_parents(self):
        return self.individuals

    def _seed(self):
        return int(self.config.seed or 0)

    def generate_offspring(self, materialize=True):
        """materialize=False: the offspring stay rows mu.. of the resident tables (the generation loop uses them there)"""
        c = self.config
        self.eng.es_offspring(self._seed(), self.generation, c.mu, c.lambda_, c.tau, c.tau_prime, c.min_sigma, self.w, self.s)
        return self._vectors(self.w[c.mu:].cpu().numpy(), self.s[c.mu:].cpu().numpy()) if materialize else None

    def select_from_combined(self, all_individuals, fitness_scores):
        """Top-mu of parents + offspring, then the two repair steps of evo/population.py:128-170.  `all_individuals`
        must be get_parents() + generate_offspring() of this generation (they are the resident rows)."""
        c = self.config
        if len(all_individuals) != len(fitness_scores):
            raise ValueError("Individuals (%d) must match fitness scores (%d)" % (len(all_individuals), len(fitness_scores)))
        if len(fitness_scores) != c.mu + c.lambda_:
            raise ValueError("Expected %d fitness scores for mu+lambda selection, got %d" % (c.mu + c.lambda_, len(fitness_scores)))
        w2, s2, f2, _ = self.eng.es_select(c.mu, np.asarray(fitness_scores, dtype=np.float64), self.w, self.s)
        self.generation += 1
        self.last_events = []
        fit = f2.cpu().numpy()
        if float(np.mean(s2.cpu().numpy())) < c.min_sigma * 10:
            self.eng.es_reset_sigmas(self._seed(), self.generation, c.initial_sigma, s2)
            self.last_events.append("sigma_reset")
        if float(np.std(fit)) == 0.0 and len(set(fit.tolist())) == 1:
            self.eng.es_inject_diversity(self._seed(), self.generation, c.tau, c.tau_prime, c.min_sigma, c.initial_sigma, w2, s2)
            self.last_events.append("diversity_injection")
        self.w[:c.mu] = w2
        self.s[:c.mu] = s2
        self.fitness_scores = fit.tolist()

    def get_best_individual(self):
        if not self.fitness_scores:
            raise ValueError("No fitness scores available")
        i = int(np.argmax(self.fitness_scores))
        return self.individuals[i], self.fitness_scores[i]

    def get_population_stats(self):
        if not self.fitness_scores:
            return {"error": "No fitness scores available"}
        mu = self.config.mu
        f = np.array(self.fitness_scores)
        w, s = self.w[:mu].cpu().numpy(), self.s[:mu].cpu().numpy()
        return {"generation": self.generation, "population_size": mu, "best_fitness": float(f.max()), "worst_fitness": float(f.min()),
                "mean_fitness": float(f.mean()), "std_fitness": float(f.std()), "diversity": float(np.mean(np.std(w, axis=0))),
                "avg_mutation_strength": float(np.mean(s))}

    def should_terminate(self):  # evo/population.py:247-279
        c = self.config
        if self.generation >= c.generations:
            return True
        if self.generation > max(10, c.generations // 10) and len(self.fitness_scores) > 1:
            f = np.asarray(self.fitness_scores)
            if (np.std(f) < 1e-10 and abs(np.mean(f)) > 1e-3 and len(set(np.round(f, 10))) == 1 and self.generation > c.generations // 2):
                return True
        return False

    def save_population(self, filepath):
        dump_checkpoint({"generation": self.generation, "individuals": self.individuals, "fitness_scores": list(self.fitness_scores),
                         "config": self.config}, filepath)

    def load_population(self, filepath):
        data = load_checkpoint(filepath)
        self.config = data["config"]
        self.generation = data["generation"]
        self.fitness_scores = list(data["fitness_scores"])
        self.w = None
        self.individuals = data["individuals"]

    def __len__(self):
        return self.config.mu


class EvolutionEngine:
    """evo/evolution.py:18-276: the generation loop over the device-resident population and the batched evaluator."""

    def __init__(self, config, deck_config=None, engine=None, evaluation="round_robin"):
        """evaluation: "round_robin" = the reference's schedule (everyone against everyone and the hall of fame,
        evo/fitness.py:32-118); "vs_hall_of_fame" = every individual plays games_per_pairing games as FIRST against the
        hall of fame and one fixed baseline vector (BASELINE config 4: work linear in the population)."""
        self.config = config
        self.deck_config = deck_config
        self.eng = engine
        self.evaluation = evaluation
        self.population = None
        self.fitness_evaluator = None
        self.results_dir = config.results_dir
        self.start_time = None
        self.baseline = WeightVector.__new__(WeightVector)
        self.baseline.weights = np.random.RandomState(7).uniform(0, 1, len(FEATURE_NAMES))
        self.baseline.sigmas, self.baseline.size = np.full(len(FEATURE_NAMES), 0.1), len(FEATURE_NAMES)

    def initialize(self):
        os.makedirs(self.results_dir, exist_ok=True)
        self.population = Population(self.config, engine=self.eng)
        self.population.initialize_population(len(FEATURE_NAMES))
        self.fitness_evaluator = FitnessEvaluator(self.config, self.deck_config, engine=self.population.eng)

    def load_checkpoint(self, checkpoint_path):
        """Resume from a population checkpoint (ours or the reference's, evo/evolution.py:243-254)."""
        os.makedirs(self.results_dir, exist_ok=True)
        self.population = Population(self.config, engine=self.eng)
        self.population.load_population(checkpoint_path)
        self.config = self.population.config
        self.fitness_evaluator = FitnessEvaluator(self.config, self.deck_config, engine=self.population.eng)

    def _evaluate(self, everyone, generation):
        ev = self.fitness_evaluator
        if self.evaluation == "round_robin":
            return ev.evaluate_population(everyone, generation)
        fitness = ev.evaluate_vs(everyone, list(ev.hall_of_fame) + [self.baseline], generation)
        ev._update_hall_of_fame(everyone, fitness)
        return fitness

    def step(self):
        """One pass of the loop body of evo/evolution.py:77-110; returns the generation wall time."""
        t0 = time.time()
        pop = self.population
        # parents (+ offspring) as rows of the resident weight table: weights never leave the device inside the loop
        everyone = pop.w[:self.config.mu]
        if pop.generation > 0:
            pop.generate_offspring(materialize=False)
            everyone = pop.w[:self.config.mu + self.config.lambda_]
        fitness = self._evaluate(everyone, pop.generation)
        if pop.generation == 0:
            pop.fitness_scores = fitness
            pop.generation += 1
        else:
            pop.select_from_combined(everyone, fitness)
        dt = time.time() - t0
        if self.config.save_logs:
            append_generation_log(os.path.join(self.results_dir, "training_log.csv"), pop.get_population_stats(),
                                  self.fitness_evaluator.get_stats(), dt)
        if pop.generation % self.config.checkpoint_interval == 0:
            stamp = datetime.now().strftime("%Y%m%d_%H%M%S")
            pop.save_population(os.path.join(self.results_dir, "checkpoint_gen%d_%s.pkl" % (pop.generation, stamp)))
        return dt

    def run(self):
        if self.population is None:
            raise ValueError("Engine not initialized. Call initialize() first.")
        self.start_time = time.time()
        while not self.population.should_terminate():
            self.step()
        total = time.time() - self.start_time
        best, best_fitness = self.population.get_best_individual()
        dump_checkpoint(best, os.path.join(self.results_dir, "best_weights.pkl"))
        self.population.save_population(os.path.join(self.results_dir, "final_population.pkl"))
        return {"config": self.config, "total_time": total, "final_generation": self.population.generation,
                "best_fitness": best_fitness, "best_weights": best, "final_stats": self.population.get_population_stats(),
                "eval_stats": self.fitness_evaluator.get_stats()}


__all__ = ["EvolutionaryConfig", "Population", "EvolutionEngine", "load_checkpoint", "dump_checkpoint", "append_generation_log", "LOG_HEADER"]
