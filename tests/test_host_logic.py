"""CPU: host-side logic of the drop-in layer (no kernels): pairing schedule, sharding, seeds, the
reference-faithful evaluator, hall of fame, and the N>1 reduction path over gloo (world_size 2)."""
import os
import socket
import sys

import numpy as np
import pytest


class Cfg:
    games_per_pairing = 3
    max_turns = 60
    seed = 7
    num_workers = 4


def test_pairings_match_reference_schedule():
    from monsoon_b200.evo import FitnessEvaluator
    p = FitnessEvaluator.pairings(3, 5)  # 3 individuals + 2 hall-of-fame opponents
    assert p == [(0, 1), (0, 2), (0, 3), (0, 4), (1, 0), (1, 2), (1, 3), (1, 4), (2, 0), (2, 1), (2, 3), (2, 4)]


def test_shard_is_a_partition():
    from monsoon_b200.evo import FitnessEvaluator
    for n in (0, 1, 7, 65536, 1000003):
        for world in (1, 2, 3, 8):
            blocks = [FitnessEvaluator.shard(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_game_seed_is_deterministic_and_spread():
    from monsoon_b200.evo import game_seed
    s = {game_seed(7, g, i, j, k) for g in range(3) for i in range(8) for j in range(8) for k in range(4)}
    assert len(s) == 3 * 8 * 8 * 4
    assert game_seed(7, 1, 2, 3, 4) == game_seed(7, 1, 2, 3, 4) and all(0 <= x < 2 ** 63 for x in s)


def test_mask_to_actions_sorted():
    from monsoon_b200.games import mask_to_actions
    m = np.zeros(5, dtype=np.uint32)
    for a in (0, 31, 32, 64, 148, 155):
        m[a >> 5] |= np.uint32(1 << (a & 31))
    assert mask_to_actions(m) == [0, 31, 32, 64, 148, 155]


def test_reference_faithful_evaluator_reproduces_degenerate_fitness():
    """the reference's evaluate_population returns 1.0 for everybody (no game is stepped)."""
    from monsoon_b200.evo import FitnessEvaluator, WeightVector
    np.random.seed(0)
    pop = [WeightVector(10) for _ in range(4)]
    ev = FitnessEvaluator(Cfg(), reference_faithful=True)
    assert ev.evaluate_population(pop, 0) == [1.0] * 4
    assert len(ev.hall_of_fame) == 4  # top-5 copies of a population of 4
    assert ev.evaluate_population(pop, 1) == [1.0] * 4  # now with hall-of-fame opponents: still (n_total-1)*g / ((n_total-1)*g)
    st = ev.get_stats()
    assert st["total_games"] == (4 * 3 + 4 * 7) * 3 and set(st) == {"total_games", "total_time", "avg_time_per_game", "games_per_second"}
    assert ev.hall_of_fame[0] is not pop[0] and np.array_equal(ev.hall_of_fame[0].weights, pop[0].weights)


class OracleEngine:
    """Test double with the Engine surface FitnessEvaluator._play uses, computing on the CPU with the oracle.
    Lets the sharding + collective logic run under gloo without a GPU."""

    def __init__(self):
        import torch
        import sb_oracle
        self.o, self.device = sb_oracle, torch.device("cpu")
        from monsoon_b200.engine import DEFAULT_DECKS, deck_indices
        self.d = [deck_indices(x) for x in DEFAULT_DECKS]

    def reset(self, seeds, decks=None, factions=None):
        import torch
        return torch.from_numpy(np.stack([self.o.new_game(int(s), self.d[0], self.d[1], 3, 2) for s in seeds.tolist()]))

    def rollout_heuristic(self, states, w_first, w_second, idx_first, idx_second, max_steps=400):
        import torch
        st = states.numpy()
        _tot, res, steps = self.o.batch_heuristic(st, w_first.numpy(), w_second.numpy(), idx_first.numpy(), idx_second.numpy(),
                                                  max_steps, 2)
        return torch.from_numpy(res.astype(np.int8)), torch.from_numpy(steps)

    def accumulate_fitness(self, result, idx_first, counts):
        for r, i in zip(result.tolist(), idx_first.tolist()):
            counts[i, 0 if r == 0 else 2 if r == 1 else 1] += 1
        return counts


def _rank_main(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")]
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from monsoon_b200.evo import FitnessEvaluator, WeightVector
    np.random.seed(3)
    pop = [WeightVector(10) for _ in range(3)]
    if rank != 0:  # only rank 0's weights count: the broadcast must overwrite these
        for v in pop:
            v.weights = np.zeros(10)
    ev = FitnessEvaluator(Cfg(), engine=OracleEngine(), chunk_games=5)
    fit = ev.evaluate_population(pop, 0)
    q.put((rank, fit, ev.last_counts.tolist()))
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_rank():
    import torch.multiprocessing as mp
    from monsoon_b200.evo import FitnessEvaluator, WeightVector
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    np.random.seed(3)
    pop = [WeightVector(10) for _ in range(3)]
    single = FitnessEvaluator(Cfg(), engine=OracleEngine(), chunk_games=1000)
    want = single.evaluate_population(pop, 0)
    assert single.last_counts.sum() == 3 * 2 * 3
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(60)
    for _rank, fit, counts in got:
        assert fit == want and counts == single.last_counts.tolist()
