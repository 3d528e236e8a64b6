"""GPU: rows f1 / f4 of SURVEY 8(f) -- the evolution-strategy operators on the resident population against the CPU
oracle (bit-exact) and against what the reference's Population / WeightVector produced (fixture), and the driver loop."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def dev(a, engine):
    return torch.from_numpy(np.ascontiguousarray(a)).to(engine.device)


@pytest.mark.parametrize("mu,lam,nf", [(12, 20, 10), (512, 512, 10), (100, 37, 33)])
def test_es_kernels_bit_exact(engine, oracle, mu, lam, nf):
    rs = np.random.RandomState(mu + lam)
    W = np.concatenate([rs.uniform(0, 1, (mu, nf)), np.zeros((lam, nf))])
    S = np.concatenate([rs.uniform(1e-4, 0.3, (mu, nf)), np.zeros((lam, nf))])
    w, s = dev(W, engine), dev(S, engine)
    parents = torch.empty(lam, dtype=torch.int32, device=engine.device)
    engine.es_offspring(77, 3, mu, lam, 0.1, 0.01, 1e-5, w, s, parents)
    want_parents = oracle.es_offspring(77, 3, mu, lam, 0.1, 0.01, 1e-5, W, S)
    assert np.array_equal(parents.cpu().numpy(), want_parents)
    assert np.array_equal(w.cpu().numpy(), W) and np.array_equal(s.cpu().numpy(), S)  # same bits, exp/log/normal included
    fitness = np.round(rs.uniform(0, 1, mu + lam), 1)  # ties: stable order matters
    w2, s2, f2, order = engine.es_select(mu, fitness, w, s, want_order=True)
    ow, os_, of, oorder = oracle.es_select(mu, fitness, W, S)
    assert np.array_equal(order.cpu().numpy(), oorder) and np.array_equal(f2.cpu().numpy(), of)
    assert np.array_equal(w2.cpu().numpy(), ow) and np.array_equal(s2.cpu().numpy(), os_)
    engine.es_reset_sigmas(77, 4, 0.1, s2)
    oracle.es_reset_sigmas(77, 4, 0.1, os_)
    assert np.array_equal(s2.cpu().numpy(), os_)
    chosen = torch.empty(max(1, mu // 2), dtype=torch.int32, device=engine.device)
    engine.es_inject_diversity(77, 4, 0.1, 0.01, 1e-5, 0.1, w2, s2, chosen)
    ochosen = oracle.es_inject_diversity(77, 4, 0.1, 0.01, 1e-5, 0.1, ow, os_)
    assert np.array_equal(chosen.cpu().numpy(), ochosen)
    assert np.array_equal(w2.cpu().numpy(), ow) and np.array_equal(s2.cpu().numpy(), os_)


def test_population_follows_reference_fixture(engine):
    """monsoon_b200.training.Population through the same generations the reference's Population went through."""
    from monsoon_b200.evo import WeightVector
    from monsoon_b200.training import EvolutionaryConfig, Population
    z = np.load(os.path.join(GOLDEN, "es_operators.npz"))
    for scenario, event in (("normal", None), ("reset", "sigma_reset"), ("inject", "diversity_injection")):
        cfg = EvolutionaryConfig(mu=12, lambda_=20, tau=0.1, tau_prime=0.01, min_sigma=1e-5, initial_sigma=0.1, seed=int(z[scenario + "_seed"]))
        pop = Population(cfg, engine=engine)
        vecs = []
        for wi, si in zip(z[scenario + "_w0"], z[scenario + "_s0"]):
            v = WeightVector.__new__(WeightVector)
            v.weights, v.sigmas, v.size = wi.copy(), si.copy(), 10
            vecs.append(v)
        pop.individuals = vecs
        pop.fitness_scores, pop.generation = [0.0] * 12, 1
        seen = set()
        for g in range(z[scenario + "_fitness"].shape[0]):
            everyone = pop.get_parents() + pop.generate_offspring()
            assert np.allclose(np.array([v.weights for v in everyone[12:]]), z[scenario + "_off_w"][g], rtol=0, atol=1e-12)
            assert np.allclose(np.array([v.sigmas for v in everyone[12:]]), z[scenario + "_off_s"][g], rtol=1e-12, atol=0)
            pop.select_from_combined(everyone, z[scenario + "_fitness"][g].tolist())
            seen.update(pop.last_events)
            assert pop.generation == 2 + g and pop.fitness_scores == z[scenario + "_sur_f"][g].tolist()
            assert np.allclose(np.array([v.weights for v in pop.individuals]), z[scenario + "_sur_w"][g], rtol=0, atol=1e-12)
            assert np.allclose(np.array([v.sigmas for v in pop.individuals]), z[scenario + "_sur_s"][g], rtol=1e-12, atol=0)
            # continue from the reference's state so that 1e-16 differences of exp() cannot accumulate
            pop.individuals = [_vec(a, b) for a, b in zip(z[scenario + "_sur_w"][g], z[scenario + "_sur_s"][g])]
        assert event is None or event in seen
        st = pop.get_population_stats()
        assert st["generation"] == pop.generation and st["population_size"] == 12


def _vec(w, s):
    from monsoon_b200.evo import WeightVector
    v = WeightVector.__new__(WeightVector)
    v.weights, v.sigmas, v.size = w.copy(), s.copy(), len(w)
    return v


def test_evolution_engine_loop(engine, tmp_path):
    """evo/evolution.py:60-110 end to end: three generations, the log and the checkpoint in the reference's formats."""
    from monsoon_b200.training import EvolutionaryConfig, EvolutionEngine, LOG_HEADER, load_checkpoint
    cfg = EvolutionaryConfig(mu=4, lambda_=4, generations=3, games_per_pairing=2, max_turns=400, seed=5, checkpoint_interval=2,
                             results_dir=str(tmp_path))
    ee = EvolutionEngine(cfg, engine=engine)
    ee.initialize()
    first = np.array([v.weights for v in ee.population.individuals])
    out = ee.run()
    assert out["final_generation"] == 3 and 0.0 <= out["best_fitness"] <= 1.0
    lines = open(os.path.join(str(tmp_path), "training_log.csv")).read().splitlines()
    assert lines[0] + "\n" == LOG_HEADER and [int(x.split(",")[0]) for x in lines[1:]] == [1, 2, 3]
    final = load_checkpoint(os.path.join(str(tmp_path), "final_population.pkl"))
    assert final["generation"] == 3 and len(final["individuals"]) == 4 and sorted(final["fitness_scores"], reverse=True) == final["fitness_scores"]
    assert any(f.startswith("checkpoint_gen2_") for f in os.listdir(str(tmp_path)))
    # same seed -> same run (per-row counter streams, integer win counts)
    cfg2 = EvolutionaryConfig(**dict(cfg.to_dict(), results_dir=str(tmp_path / "again")))
    ee2 = EvolutionEngine(cfg2, engine=engine)
    ee2.initialize()
    assert np.array_equal(first, np.array([v.weights for v in ee2.population.individuals]))
    out2 = ee2.run()
    assert out2["best_fitness"] == out["best_fitness"] and np.array_equal(out2["best_weights"].weights, out["best_weights"].weights)


def test_population_checkpoint_round_trip(engine, tmp_path):
    """Population.save_population / load_population (evo/population.py:281-310) through the reference-compatible
    pickle, and resuming from the reference's own checkpoint file."""
    from monsoon_b200.training import EvolutionaryConfig, Population
    cfg = EvolutionaryConfig(mu=6, lambda_=6, seed=3)
    pop = Population(cfg, engine=engine)
    pop.initialize_population(10)
    pop.fitness_scores, pop.generation = [0.6, 0.5, 0.4, 0.3, 0.2, 0.1], 4
    path = str(tmp_path / "pop.pkl")
    pop.save_population(path)
    other = Population(EvolutionaryConfig(mu=2, lambda_=2), engine=engine)
    other.load_population(path)
    assert other.generation == 4 and other.fitness_scores == pop.fitness_scores and other.config == cfg
    assert np.array_equal(np.array([v.weights for v in other.individuals]), np.array([v.weights for v in pop.individuals]))
    assert np.array_equal(np.array([v.sigmas for v in other.individuals]), np.array([v.sigmas for v in pop.individuals]))
    children = other.generate_offspring()  # the resumed population is usable: rows for the offspring exist
    assert len(children) == 6 and len(other) == 6
    ref = Population(EvolutionaryConfig(), engine=engine)
    ref.load_population(os.path.join(GOLDEN, "ref_population.pkl"))  # written by the reference
    assert ref.generation == 5 and ref.config.mu == 12 and len(ref.individuals) == 12
    assert len(ref.generate_offspring()) == 20
    best, fit = ref.get_best_individual()
    assert fit == max(ref.fitness_scores) and best.size == 10


def test_evolution_engine_vs_hall_of_fame_and_resume(engine, tmp_path):
    """The linear schedule of BASELINE config 4 (hall of fame + baseline opponents) and resuming from a checkpoint."""
    from monsoon_b200.training import EvolutionaryConfig, EvolutionEngine
    cfg = EvolutionaryConfig(mu=6, lambda_=6, generations=2, games_per_pairing=2, max_turns=400, seed=9, checkpoint_interval=1,
                             results_dir=str(tmp_path / "a"))
    ee = EvolutionEngine(cfg, engine=engine, evaluation="vs_hall_of_fame")
    ee.initialize()
    out = ee.run()
    assert out["final_generation"] == 2 and len(ee.fitness_evaluator.hall_of_fame) == 5
    assert out["eval_stats"]["total_games"] == 6 * 1 * 2 + 12 * 6 * 2  # generation 0: baseline only; then 5 hall-of-fame + baseline
    ckpt = sorted(f for f in os.listdir(cfg.results_dir) if f.startswith("checkpoint_gen1_"))[0]
    cfg2 = EvolutionaryConfig(**dict(cfg.to_dict(), results_dir=str(tmp_path / "b"), generations=3))
    ee2 = EvolutionEngine(cfg2, engine=engine, evaluation="vs_hall_of_fame")
    ee2.load_checkpoint(os.path.join(cfg.results_dir, ckpt))
    assert ee2.population.generation == 1 and ee2.config.generations == 2  # the checkpoint carries its configuration
    ee2.config.generations = 3
    out2 = ee2.run()
    assert out2["final_generation"] == 3
