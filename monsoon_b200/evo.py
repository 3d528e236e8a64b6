"""Drop-in for the reference's evo hot path: evo/game_adapter.py (StormboundAdapter),
evo/heuristic_agent.py (HeuristicAgent), evo/weights.py (WeightVector) and the evaluation entry point
evo/fitness.py (FitnessEvaluator.evaluate_population) -- same names, signatures and return shapes,
computed by the CUDA kernels.  `train_evolutionary.py` / evo/evolution.py run unchanged on top
(INTEGRATION.md).

Two semantics are offered for the evaluator:
  * default: the INTENDED loop of evo/fitness.py:193-206 (play until have_winner or max_turns env steps);
  * reference_faithful=True: reproduces what the reference actually returns (is_terminal is always True,
    no game is stepped, the row player is always credited a win -> every fitness is 1.0).
"""
import time

import numpy as np
import torch

from .engine import get_engine
from .games import EngineError, Game, mask_to_actions

FEATURE_NAMES = ["mana_efficiency", "health_advantage", "board_control", "front_line_advantage", "total_strength",
                 "unit_count", "structure_count", "threatened_base", "protection_value", "hand_quality"]


class WeightVector:
    """evo/weights.py:12-123 (host-side numpy; the ES operators are row f1 of SURVEY 8(f))."""

    def __init__(self, size):
        self.weights = np.random.uniform(0, 1, size)
        self.sigmas = np.full(size, 0.1)
        self.size = size

    def mutate(self, tau, tau_prime, min_sigma):  # evo/weights.py:20-40
        global_noise = np.random.normal(0, 1)
        individual_noise = np.random.normal(0, 1, len(self.sigmas))
        self.sigmas = np.maximum(self.sigmas * np.exp(tau_prime * global_noise + tau * individual_noise), min_sigma)
        self.weights = np.clip(self.weights + np.random.normal(0, self.sigmas), 0, 1)

    def copy(self):
        v = WeightVector(len(self.weights))
        v.weights, v.sigmas, v.size = self.weights.copy(), self.sigmas.copy(), self.size
        return v

    def dot_product(self, features):
        if len(features) != self.size:
            raise ValueError("Feature vector size %d doesn't match weight vector size %d" % (len(features), self.size))
        return np.dot(self.weights, features)

    def get_weights(self):
        return self.weights.copy()

    def get_sigmas(self):
        return self.sigmas.copy()

    def set_weights(self, weights):
        if len(weights) != self.size:
            raise ValueError("Weight array size %d doesn't match expected size %d" % (len(weights), self.size))
        self.weights = np.clip(weights, 0, 1)

    def set_sigmas(self, sigmas):
        if len(sigmas) != self.size:
            raise ValueError("Sigma array size %d doesn't match expected size %d" % (len(sigmas), self.size))
        self.sigmas = np.maximum(sigmas, 1e-10)

    def distance_to(self, other):
        return np.linalg.norm(self.weights - other.weights)


class StateFeatures:
    """evo/features.py:9-362 result object: the ten features as attributes + get_feature_vector()."""

    def __init__(self, vector, current_player):
        self._v = np.asarray(vector, dtype=np.float64)
        self.current_player = current_player
        for name, val in zip(FEATURE_NAMES, self._v):
            setattr(self, name, float(val))

    def get_feature_vector(self):
        return self._v.copy()

    @staticmethod
    def get_feature_count():
        return 10

    @staticmethod
    def get_feature_names():
        return list(FEATURE_NAMES)


class StormboundAdapter:
    """evo/game_adapter.py:272-372.  Value semantics: apply_action returns an independent state (incl. the
    random stream position) and leaves this one untouched."""

    def __init__(self, game, reference_faithful=False):
        self.game = game
        self.reference_faithful = reference_faithful
        self.initial_observation = None
        self._cached_observation = None

    def clone_state(self):
        a = StormboundAdapter(self.game.clone(), self.reference_faithful)
        a._cached_observation = self._cached_observation
        return a

    def get_legal_actions(self):
        return self.game.legal_actions()

    def apply_action(self, action):
        new_state = self.clone_state()
        observation, _reward, _done = new_state.game.step(action)
        new_state._cached_observation = observation
        return new_state

    def extract_features(self):
        f, err = self.game.eng.features(self.game.state)
        if int(err[0]):
            raise EngineError(int(err[0]))
        return StateFeatures(f[0].cpu().numpy(), self.game.to_play())

    def is_terminal(self):
        if self.reference_faithful:
            return True  # `have_winner() is not None` is always True (evo/game_adapter.py:342, Q15)
        return self.game.env.have_winner()

    def get_result(self):
        if self.reference_faithful:
            return 1 if self.game.env.have_winner() else 0  # bool == 0 / == 1 (evo/game_adapter.py:351-357)
        first, second = self._bases()
        return 0 if (second < 0 <= first) else 1 if (first < 0 <= second) else -1

    def _bases(self):
        lo = int(self.game._host()[14])
        l, r = self.game.env.board.local.strength, self.game.env.board.remote.strength
        return (l, r) if lo == 0 else (r, l)

    def get_current_player(self):
        return self.game.to_play()

    def get_observation(self):
        if self._cached_observation is None:
            self._cached_observation = self.game.env.get_observation()
        return self._cached_observation


class HeuristicAgent:
    """evo/heuristic_agent.py:14-127; weights is anything with a `.weights` float64[10] (the reference's
    WeightVector works unchanged)."""

    def __init__(self, weights, player_idx):
        self.weights = weights
        self.player_idx = player_idx
        self.action_count = 0
        self.game_count = 0

    def _w(self, eng):
        return torch.as_tensor(np.asarray(self.weights.weights, dtype=np.float64).reshape(1, 10)).to(eng.device)

    def score_action(self, state, action):
        eng = state.game.eng
        _a, scores = eng.select_action(state.game.state, self._w(eng), want_scores=True)
        s = float(scores[0, int(action)])
        return 0.0 if np.isnan(s) else s  # evo/heuristic_agent.py:48-51: any failure scores 0.0

    def select_action(self, state):
        eng = state.game.eng
        a, _ = eng.select_action(state.game.state, self._w(eng))
        self.action_count += 1
        return int(a[0])

    def reset_for_new_game(self):
        self.action_count = 0
        self.game_count += 1

    def get_weights(self):
        return self.weights

    def set_weights(self, weights):
        self.weights = weights

    def get_player_idx(self):
        return self.player_idx


def game_seed(base_seed, generation, i, j, k):
    """Deterministic per-game seed (the reference uses OS entropy here, Q16): partition-invariant, so any
    sharding of the games over ranks gives the same fitness."""
    x = (int(base_seed) * 0x9E3779B97F4A7C15 + generation * 0xBF58476D1CE4E5B9 + i * 0x94D049BB133111EB + j * 0xD6E8FEB86659FD93 + k) & 0x7FFFFFFFFFFFFFFF
    x ^= x >> 31
    return (x * 0x2545F4914F6CDD1D) & 0x7FFFFFFFFFFFFFFF


def game_seed_array(base_seed, generation, i, j, k):
    """Vectorised game_seed (uint64 wrap-around arithmetic == the & 0x7FFF... masks of the scalar version)."""
    M = np.uint64(0x7FFFFFFFFFFFFFFF)
    with np.errstate(over="ignore"):
        x = (np.uint64(int(base_seed) & 0xFFFFFFFFFFFFFFFF) * np.uint64(0x9E3779B97F4A7C15)
             + np.uint64(generation) * np.uint64(0xBF58476D1CE4E5B9)
             + i.astype(np.uint64) * np.uint64(0x94D049BB133111EB)
             + j.astype(np.uint64) * np.uint64(0xD6E8FEB86659FD93) + k.astype(np.uint64)) & M
        x ^= x >> np.uint64(31)
        x = (x * np.uint64(0x2545F4914F6CDD1D)) & M
    return x.astype(np.int64)


class DeckEvolutionConfig:
    """utils.py:121-241: exploit -> explore -> balance schedule of the decks each game is dealt.

    Archetypes are 12 card names ("UA07"), card-table indices or reference card objects.  The decks of a whole
    batch are drawn on the device (`generate_batch` -> sb_generate_decks): game g uses its own counter-based
    stream keyed by the game seed where the reference draws from the process-global `random` module.
    """

    def __init__(self, player1_archetype, player2_archetype, exploit_generations=30, explore_generations=30,
                 max_random_ratio=0.5, balance_archetype_ratio=0.7):
        from ._card_table import CARDS
        index = {r["name"]: i for i, r in enumerate(CARDS)}

        def ids(deck):
            out = [c if isinstance(c, (int, np.integer)) else index[c if isinstance(c, str) else type(c).__name__] for c in deck]
            if len(out) != 12:
                raise NotImplementedError("archetypes of 12 cards only (the packed deck layout; utils.py pads/truncates)")
            return [int(c) for c in out]

        self.player1_archetype = ids(player1_archetype)
        self.player2_archetype = ids(player2_archetype)
        self.exploit_generations = exploit_generations
        self.explore_generations = explore_generations
        self.max_random_ratio = max_random_ratio
        self.balance_archetype_ratio = balance_archetype_ratio
        self.player1_faction = int(CARDS[self.player1_archetype[0]]["faction"])  # utils.py:152-153
        self.player2_faction = int(CARDS[self.player2_archetype[0]]["faction"])
        self._names = [r["name"] for r in CARDS]

    def phase_parameters(self, generation):
        """(mode, n_preserve, q) of sb_generate_decks for this generation (utils.py:155-218, :35-58)."""
        if generation < self.exploit_generations:
            return 0, 12, 0.0
        if generation < self.exploit_generations + self.explore_generations:
            progress = (generation - self.exploit_generations) / self.explore_generations
            preserve_ratio = max(0.0, min(1.0, 1.0 - progress * self.max_random_ratio))
            if preserve_ratio == 1.0:
                return 1, 12, 0.0  # generate_random_deck returns original[:12]
            return 1, (min(int(12 * preserve_ratio), 12) if preserve_ratio > 0.0 else 0), 0.0
        return 2, 0, float(self.balance_archetype_ratio)

    def generate_batch(self, engine, seeds, generation):
        mode, n_preserve, q = self.phase_parameters(generation)
        return engine.generate_decks(seeds, generation, mode, n_preserve, q, [self.player1_archetype, self.player2_archetype],
                                     [self.player1_faction, self.player2_faction])

    def get_deck_configuration(self, generation, seed=0, engine=None):
        """One game's (player1_deck, player2_deck) as card names."""
        decks, _f = self.generate_batch(engine or get_engine(), np.asarray([seed], dtype=np.int64), generation)
        d = decks[0].cpu().numpy()
        return [self._names[c] for c in d[0]], [self._names[c] for c in d[1]]

    def get_phase_info(self, generation):  # utils.py:220-241
        if generation < self.exploit_generations:
            phase, random_ratio = "Exploit", 0.0
        elif generation < self.exploit_generations + self.explore_generations:
            phase = "Explore"
            random_ratio = (generation - self.exploit_generations) / self.explore_generations * self.max_random_ratio
        else:
            phase, random_ratio = "Balance", 1.0 - self.balance_archetype_ratio
        return {"phase": phase, "generation": generation, "random_ratio": random_ratio,
                "exploit_complete": generation >= self.exploit_generations,
                "explore_complete": generation >= self.exploit_generations + self.explore_generations}


class FitnessEvaluator:
    """evo/fitness.py:18-259.  evaluate_population(population, generation) -> List[float] in [0, 1].

    All games of a generation are played by sb_reset + sb_rollout_heuristic launches; win/draw/loss counts
    are reduced on the device (sb_accumulate_fitness) and, when torch.distributed is initialised, summed
    over ranks with one int32 all-reduce after the weights were broadcast from rank 0.
    `config` needs .games_per_pairing, .max_turns, .seed (num_workers is ignored: there are no threads)."""

    def __init__(self, config, deck_config=None, reference_faithful=False, device=None, engine=None, chunk_games=262144):
        self.config = config
        self.deck_config = deck_config
        self.reference_faithful = reference_faithful
        self.total_games = 0
        self.total_time = 0.0
        self.hall_of_fame = []
        self.hall_of_fame_size = 5
        self.use_hall_of_fame = True
        self.chunk_games = chunk_games
        self._eng = engine
        self._device = device
        self.last_counts = None
        self.last_aborted = (0, 0)  # games of the last evaluation stopped (like the reference would / by a limit of this engine)
        self.total_aborted = [0, 0]

    # -- plumbing
    def _engine(self):
        if self._eng is None:
            dev = self._device
            if dev is None:
                dev = torch.cuda.current_device() if torch.cuda.is_available() else 0
            self._eng = get_engine(dev)
        return self._eng

    @staticmethod
    def pairings(n_individuals, n_total):
        """evo/fitness.py:52-59: ordered pairs, the row player is FIRST, no self-play inside the population."""
        return [(i, j) for i in range(n_individuals) for j in range(n_total) if not (j < n_individuals and i == j)]

    @staticmethod
    def shard(n_games, rank, world):
        """contiguous block of the game index space for this rank."""
        lo = n_games * rank // world
        hi = n_games * (rank + 1) // world
        return lo, hi

    def evaluate_population(self, population, generation=0):
        n = len(population)
        hof = list(self.hall_of_fame) if self.use_hall_of_fame else []
        if torch.is_tensor(population):  # resident rows (training.Population.w): the table never visits the host
            all_opponents = torch.cat([population, self._weight_table(hof, population.device)]) if hof else population
        else:
            all_opponents = list(population) + hof
        n_total = len(all_opponents)
        g = int(self.config.games_per_pairing)
        pairs = self.pairings(n, n_total) if (self.reference_faithful or n_total < 2) else None
        n_pairs = n * (n_total - 1)
        start = time.time()
        if self.reference_faithful or not n_pairs:
            # evo/fitness.py:193,217: the loop never runs and `False == 0` credits the row player (Q15)
            counts = np.zeros((n, 3), dtype=np.int64)
            for i, _j in pairs or []:
                counts[i, 0] += g
        else:
            counts = self._play(population, all_opponents, None, g, generation, schedule="round_robin")
        self.last_counts = counts
        scores = counts[:, 0] * 1.0 + counts[:, 1] * 0.5
        per_individual = (n_total - 1) * g  # evo/fitness.py:112 (Q20 normalisation kept)
        fitness = [float(s) / per_individual if per_individual else 0.0 for s in scores]
        self.total_games += n_pairs * g
        self.total_time += time.time() - start
        self._update_hall_of_fame(population, fitness)
        return fitness

    def evaluate_vs(self, population, opponents, generation=0, games_per_opponent=None):
        """Added schedule: every individual plays `games_per_opponent` games as FIRST
        against each fixed opponent (a baseline vector, hall of fame ...).  Same kernels, same reduction;
        returns (wins + 0.5 draws) / games per individual."""
        n = len(population)
        g = int(games_per_opponent or self.config.games_per_pairing)
        start = time.time()
        if torch.is_tensor(population):  # resident weight tables (training.Population.w): no host detour
            table = torch.cat([population, self._weight_table(opponents, population.device)])
        else:
            table = list(population) + list(opponents)
        counts = self._play(population, table, None, g, generation, schedule="versus")
        self.last_counts = counts
        self.total_games += n * len(opponents) * g
        self.total_time += time.time() - start
        per = len(opponents) * g
        return [float(c[0] + 0.5 * c[1]) / per for c in counts]

    def evaluate_vs_expert(self, population, generation=0, games=None):
        """Every individual plays `games` games as FIRST against the scripted opponent Stormbound.expert_action
        (the agent-vs-expert match of play_vs_expert.py:65-94, batched); returns (wins + 0.5 draws) / games."""
        n = len(population)
        g = int(games or self.config.games_per_pairing)
        start = time.time()
        counts = self._play(population, population if torch.is_tensor(population) else list(population), None, g, generation, expert_second=True,
                            schedule="solo")
        self.last_counts = counts
        self.total_games += n * g
        self.total_time += time.time() - start
        return [float(c[0] + 0.5 * c[1]) / g for c in counts]

    @staticmethod
    def _weight_table(vectors, dev):
        """rows of a weight table: a list of WeightVectors, or a resident f64[n, 10] tensor (training.Population.w) used as it is"""
        if torch.is_tensor(vectors):
            return vectors.to(device=dev, dtype=torch.float64).contiguous()
        return torch.as_tensor(np.stack([np.asarray(v.weights, dtype=np.float64) for v in vectors])).to(dev)

    def _play(self, population, all_opponents, pairs, g, generation, expert_second=False, schedule=None):
        """schedule: "round_robin" / "versus" / "solo" when `pairs` is that standard enumeration -- the device derives pair indices
        and seeds itself (sb_eval_population) and nothing but the weight table crosses the bus; None: `pairs` as given."""
        import torch.distributed as dist
        eng = self._engine()
        dev = eng.device
        dist_on = dist.is_available() and dist.is_initialized()
        rank, world = (dist.get_rank(), dist.get_world_size()) if dist_on else (0, 1)
        w = self._weight_table(all_opponents, dev)
        if dist_on:
            dist.broadcast(w, src=0)  # population weights from rank 0
        n_ind = len(population)
        n_pairs = len(pairs) if pairs is not None else {"round_robin": n_ind * (w.shape[0] - 1), "versus": n_ind * (w.shape[0] - n_ind),
                                                       "solo": n_ind}[schedule]
        n_games = n_pairs * g
        lo, hi = self.shard(n_games, rank, world)
        base_seed = int(getattr(self.config, "seed", 0) or 0)
        max_steps = int(self.config.max_turns)
        counts = torch.zeros((n_ind, 3), dtype=torch.int32, device=dev)
        aborted = torch.zeros(2, dtype=torch.int32, device=dev)
        if schedule is not None and self.deck_config is None:
            eng.eval_population(schedule, n_ind, w, g, base_seed, generation, lo, hi, max_steps=max_steps, counts=counts, aborted=aborted,
                                chunk_games=self.chunk_games)
        else:
            pair_arr = np.asarray(pairs, dtype=np.int64) if pairs is not None else None
            for c0 in range(lo, hi, self.chunk_games):
                c1 = min(hi, c0 + self.chunk_games)
                if schedule is not None:
                    idx_first, idx_second, seeds_d = eng.eval_schedule(schedule, n_ind, w.shape[0], g, base_seed, generation, c0, c1 - c0)
                else:
                    gi = np.arange(c0, c1, dtype=np.int64)
                    pi, k = gi // g, gi % g
                    i_idx, j_idx = pair_arr[pi, 0], pair_arr[pi, 1]
                    seeds = game_seed_array(base_seed, generation, i_idx, j_idx, k)
                    idx_first = torch.as_tensor(i_idx.astype(np.int32)).to(dev)
                    idx_second = torch.as_tensor(j_idx.astype(np.int32)).to(dev)
                    seeds_d = torch.as_tensor(seeds).to(dev)
                if self.deck_config is not None:  # evo/fitness.py:136-141: generation-aware decks, drawn per game
                    decks, factions = self.deck_config.generate_batch(eng, seeds_d, generation)
                    states = eng.reset(seeds_d, decks, factions)
                else:
                    states = eng.reset(seeds_d)
                result, _steps = eng.rollout_heuristic(states, w, None if expert_second else w, idx_first, None if expert_second else idx_second,
                                                       max_steps=max_steps)
                eng.accumulate_fitness(result, idx_first, counts)
                eng.count_aborted(states, result, aborted)
        if dist_on:
            both = torch.cat([counts.view(-1), aborted])  # one collective: integer counts are order-independent, bit-exact
            dist.all_reduce(both, op=dist.ReduceOp.SUM)
            counts, aborted = both[:-2].view(n_ind, 3), both[-2:]
        host = torch.cat([counts.view(-1), aborted]).cpu().numpy()  # the one device-to-host read of an evaluation
        ab = host[-2:]
        self.last_aborted = (int(ab[0]), int(ab[1]))
        self.total_aborted[0] += int(ab[0])
        self.total_aborted[1] += int(ab[1])
        if ab[1] * 100 > max(n_games, 1):  # scored as draws although the reference would have kept playing: say so
            import warnings
            warnings.warn("%d of %d games stopped at a capacity limit of the engine (SB_ERR_UNSUPPORTED/OVERFLOW/DEPTH) and were "
                          "scored as draws" % (int(ab[1]), n_games))
        return host[:-2].reshape(n_ind, 3).astype(np.int64)

    def get_stats(self):
        return {"total_games": self.total_games, "total_time": self.total_time,
                "avg_time_per_game": self.total_time / max(self.total_games, 1),
                "games_per_second": self.total_games / max(self.total_time, 1e-6)}  # exactly the reference's keys (evo/fitness.py:230-240)

    def get_aborted(self):
        """games stopped early since the last reset_stats(): by an exception the reference raises too (scored as a draw there as
        well, evo/fitness.py:208-210) / by a capacity limit of this engine (the reference would have kept playing)"""
        return {"aborted_like_reference": self.total_aborted[0], "aborted_by_engine_limit": self.total_aborted[1]}

    def reset_stats(self):
        self.total_games = 0
        self.total_time = 0.0
        self.total_aborted = [0, 0]

    def _update_hall_of_fame(self, population, fitness):  # evo/fitness.py:247-259
        pairs = sorted(zip(fitness, range(len(population))), key=lambda x: x[0], reverse=True)
        top = [i for _f, i in pairs[:self.hall_of_fame_size]]
        if torch.is_tensor(population):  # resident rows: only the winners come to the host (sigmas are not part of an opponent)
            rows = population[torch.as_tensor(top, device=population.device)].cpu().numpy() if top else []
            self.hall_of_fame = []
            for r in rows:
                v = WeightVector.__new__(WeightVector)
                v.weights, v.sigmas, v.size = r.copy(), np.zeros_like(r), len(r)
                self.hall_of_fame.append(v)
        else:
            self.hall_of_fame = [population[i].copy() for i in top]


def play_game(adapter, agent1, agent2, max_turns=400):
    """The intended loop of evo/fitness.py:178-228 through the shim objects (API-level use; the batched
    evaluator does the same thing inside one kernel)."""
    turn_count = 0
    while not adapter.game.env.have_winner() and turn_count < max_turns:
        agent = agent1 if adapter.get_current_player() == 0 else agent2
        adapter = adapter.apply_action(agent.select_action(adapter))
        turn_count += 1
    return adapter, turn_count


__all__ = ["WeightVector", "StateFeatures", "StormboundAdapter", "HeuristicAgent", "FitnessEvaluator", "DeckEvolutionConfig", "play_game",
           "game_seed", "mask_to_actions", "Game"]
