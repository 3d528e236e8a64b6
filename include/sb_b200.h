/* sb_b200.h -- C ABI of libsb_b200.so, the B200 (sm_100a) batched Stormbound simulator.
 *
 * The reference has no FFI: its boundary is three Python class contracts.  Every entry
 * point below names the reference interface it replaces; the Python shims in monsoon_b200/ (Game,
 * StormboundAdapter, HeuristicAgent, FitnessEvaluator) bind these with ctypes and keep the reference's
 * signatures.  INTEGRATION.md shows the binding a maintainer of the reference would add.
 *
 * Conventions: plain C types only; every `*_d` pointer is a DEVICE pointer owned by the caller (e.g.
 * torch tensor .data_ptr()); `stream` is a cudaStream_t passed as void* (NULL = default stream); calls
 * are asynchronous on that stream unless named *_host; return 0 on success, a negative CUDA error
 * code otherwise (sb_last_error gives the text); no hidden allocation after sb_create except the
 * *_host helpers' staging buffers; one handle per device, not thread-safe.
 * There is NO CPU fallback: without a CUDA device sb_create fails.
 */
#ifndef SB_B200_H
#define SB_B200_H
#include <stdint.h>
#include "sb_state.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct SbHandle SbHandle;

/* library / layout introspection (no GPU needed) */
int sb_abi_version(void);
int sb_state_bytes(void);                 /* == SB_STATE_BYTES */
int sb_card_count(void);                  /* rows of the card table (130) */
int sb_card_info(int card, int32_t out[12]); /* kind,faction,cost,strength,movement,trigger,fixed,has_ability,first_type,types,obs_id,has_target */

/* One handle per device and per stream of use: the handle owns small device scratch (refill counter, staged archetypes,
 * host-variant staging buffer), so two calls through the SAME handle must not run concurrently on different streams. */
int sb_create(int device, SbHandle **out);
int sb_destroy(SbHandle *h);
const char *sb_last_error(SbHandle *h);
int sb_device(SbHandle *h);
int sb_sm_count(SbHandle *h);
/* tuning knobs for the rollout kernels (all default to "auto by batch size"): "games_per_warp" = 0 (auto) or
 * 1..32 consecutive lanes of every warp that carry a game; "turn_sync" 0/1; "block_sync" = -1 auto, 0, or the
 * CTA size (128..1024) whose warps change phase together; "refill" = -1 auto, 0/1: finished lanes of the random
 * rollout take the next game from a counter ("refill_ctas" persistent CTAs per SM, "refill_grid" CTAs in total,
 * 0 = auto); "dense" = -1 auto, 0/1: the 32-register builds with 2,048 resident lanes per SM (random rollout, k_step);
 * "heur_wpc" = -1 / 4: independent warps (default), 8 / 16 / 32: warps per CTA deciding in step; "heur_refill" = -1 auto
 * (batches beyond one wave of resident warps), 0/1: a warp of the heuristic rollout whose game ended takes the next game
 * from a counter; "heur_iw" = -1 auto, 4 / 8 / 16 independent warps per CTA; "heur_grid" = persistent CTAs of that shape
 * (0 = one wave; tests use tiny grids); "lanes_per_game" is accepted and ignored (retired shape).
 * Round 2: "engine" = -1 auto (measured per-kernel policy), 0 thread-per-game kernels, 1 warp-per-game kernels; "w_shape" /
 * "w_hshape" / "w_grid" = CTA shape and persistent grid of the warp-per-game rollouts (-1 / 0 = auto); "heur_pack" = -1 auto
 * (= 1), 0 the round-1 heuristic rollout, 1 one game per warp in the owner / holder structure, 2 / 3 / 4 / 8 several games per
 * warp (candidates of K games dealt to the 32 lanes), 5 / 6 / 7 other CTA shapes of K = 1.  Every setting gives identical results. */
int sb_set_option(SbHandle *h, const char *key, int value);
/* kernels launched through this handle so far (bench.py's gpu_launches) */
uint64_t sb_launch_count(SbHandle *h);

/* Stormbound.__init__ + Player.__init__ + Board.__init__ (games/stormbound.py:293-304, player.py:13-37,
 * board.py:16-28): shuffle both 12-card decks, weights, draw 4, base 20, mana 3/4.
 * decks_d: u8[n,2,n_deck] card indices, or u8[2,n_deck] when decks_shared; factions_d: u8[n,2] or u8[2]. */
int sb_reset(SbHandle *h, int n, const uint64_t *seeds_d, const uint8_t *decks_d, int n_deck, int decks_shared,
             const uint8_t *factions_d, uint8_t *states_d, void *stream);

/* Stormbound.legal_actions (games/stormbound.py:528-557) -> 156-bit mask per game, u32[n,5] */
int sb_legal_mask(SbHandle *h, int n, const uint8_t *states_d, uint32_t *masks_d, void *stream);

/* Stormbound.step (games/stormbound.py:318-373): apply actions_d[n]; reward {0,1} (Game.step multiplies
 * by 10, :140), done, err (non-zero = the reference raises here); next_masks_d (nullable) = legal set
 * of the resulting state, fused into the same pass. */
int sb_step(SbHandle *h, int n, uint8_t *states_d, const uint8_t *actions_d, int8_t *reward_d, uint8_t *done_d,
            uint8_t *err_d, uint32_t *next_masks_d, void *stream);

/* Stormbound.get_observation (games/stormbound.py:400-526): i32[n,27,5,4] */
int sb_observe(SbHandle *h, int n, const uint8_t *states_d, int32_t *obs_d, uint8_t *err_d, void *stream);

/* StateFeatures(obs).get_feature_vector() (evo/features.py:12-342): f64[n,10] */
int sb_features(SbHandle *h, int n, const uint8_t *states_d, double *feat_d, uint8_t *err_d, void *stream);

/* HeuristicAgent.select_action / score_action (evo/heuristic_agent.py:23-80) on StormboundAdapter forks
 * (evo/game_adapter.py:280-324): weights_d f64[n,10]; actions_d u8[n]; scores_d (nullable) f64[n,156],
 * NaN for illegal actions.  One warp per game, one lane per candidate action. */
int sb_select_action(SbHandle *h, int n, const uint8_t *states_d, const double *weights_d, uint8_t *actions_d,
                     double *scores_d, void *stream);

/* DeckEvolutionConfig.get_deck_configuration / generate_random_deck (utils.py:26-119,121-241) for n games: game g
 * draws from philox(counter=(draw, generation, 0xDEC4, 0), key=seeds_d[g]) with CPython's random.sample call
 * shape.  mode 0 exploit (archetypes verbatim), 1 explore (n_preserve = min(int(12*preserve_ratio), 12) cards
 * sampled from the archetype, the rest from own faction + NEUTRAL), 2 balance (random() < q keeps the archetype,
 * drawn for both seats first), 3 fully random decks for per-game factions_d u8[n,2] (BASELINE config 5).
 * archetypes u8[2][12] card ids and arch_factions u8[2] are HOST pointers (the config object's fields);
 * factions_d nullable (= arch_factions for every game); decks_d u8[n,2,12]; factions_out_d nullable u8[n,2].
 * The outputs feed sb_reset(decks_d, 12, 0, factions). */
int sb_generate_decks(SbHandle *h, int n, const uint64_t *seeds_d, uint32_t generation, int mode, int n_preserve, double q,
                      const uint8_t *archetypes, const uint8_t *arch_factions, const uint8_t *factions_d, uint8_t *decks_d,
                      uint8_t *factions_out_d, void *stream);

/* ---- (mu + lambda) evolution-strategy operators on a device-resident population.
 * weights / sigmas: f64 [rows][n_features] row-major, n_features <= 64, mu <= 4096.  Row r draws from
 * philox(counter=(draw, r, tag, generation), key=seed) with the reference's call shapes; exp/log use correctly
 * rounded basic operations only, so results are bit-identical to oracle/sb_oracle_es.c.
 *
 * sb_es_offspring: Population.generate_offspring (evo/population.py:75-90) + WeightVector.mutate
 * (evo/weights.py:20-40): child c = mutate(copy(parent randint(0, mu))) written to row mu + c; parents_d nullable i32[lambda].
 * sb_es_select: the top-mu part of Population.select_from_combined (evo/population.py:99-107): fitness descending,
 * ties in input order; order_d nullable i32[mu] = source row of each survivor.
 * sb_es_reset_sigmas: evo/population.py:128-139.  sb_es_inject_diversity: evo/population.py:146-170; chosen_d
 * nullable i32[max(1, mu/2)].  The caller evaluates the two trigger conditions from the survivors' statistics. */
int sb_es_offspring(SbHandle *h, uint64_t seed, uint32_t generation, int mu, int lambda, int n_features, double tau, double tau_prime,
                    double min_sigma, double *w_d, double *s_d, int32_t *parents_d, void *stream);
int sb_es_select(SbHandle *h, int total, int mu, int n_features, const double *fitness_d, const double *w_d, const double *s_d,
                 double *w_out_d, double *s_out_d, double *fit_out_d, int32_t *order_d, void *stream);
int sb_es_reset_sigmas(SbHandle *h, uint64_t seed, uint32_t generation, int mu, int n_features, double initial_sigma, double *s_d, void *stream);
int sb_es_inject_diversity(SbHandle *h, uint64_t seed, uint32_t generation, int mu, int n_features, double tau, double tau_prime,
                           double min_sigma, double initial_sigma, double *w_d, double *s_d, int32_t *chosen_d, void *stream);

/* Stormbound.expert_action (games/stormbound.py:563-637) for n games: the scripted opponent behind
 * Game.expert_agent (games/stormbound.py:201-209).  It draws its choices from the GAME's stream, so each
 * state's draw counter (and err byte, for the reference's choice([]) / max([]) exceptions) is updated in
 * place; actions_d u8[n]. */
int sb_expert_action(SbHandle *h, int n, uint8_t *states_d, uint8_t *actions_d, void *stream);

/* Uniform-random legal agent until done/err or max_steps more steps (the agent
 * stream is philox(counter=(step,0,0xA6E7,0), key=seed)).  steps_d i32[n] = steps taken by this call;
 * chain_d (nullable) u64[n] = running hash of the per-step state digests (parity evidence). */
int sb_rollout_random(SbHandle *h, int n, uint8_t *states_d, int max_steps, int32_t *steps_d, uint64_t *chain_d,
                      void *stream);

/* FitnessEvaluator._play_game (evo/fitness.py:178-228) with the intended loop (until have_winner or
 * max_steps env steps): FIRST plays w_first_d[idx_first_d[g]], SECOND w_second_d[idx_second_d[g]]
 * (idx arrays nullable = row g... row 0 when the weight table has one row is expressed by idx).
 * A NULL weight table hands that seat to Stormbound.expert_action (games/stormbound.py:563-637): the agent-vs-expert
 * match of play_vs_expert.py:65-94, batched.
 * result_d i8[n]: 0 FIRST won, 1 SECOND won, -1 draw/timeout, -2 aborted by an engine exception. */
int sb_rollout_heuristic(SbHandle *h, int n, uint8_t *states_d, const double *w_first_d, const double *w_second_d,
                         const int32_t *idx_first_d, const int32_t *idx_second_d, int max_steps, int8_t *result_d,
                         int32_t *steps_d, void *stream);

/* FitnessEvaluator.evaluate_population inner reduction (evo/fitness.py:95,160-166): counts_d i32[P,3]
 * += {wins, draws, losses} of individual idx_first_d[g] over the n finished games. */
int sb_accumulate_fitness(SbHandle *h, int n, const int8_t *result_d, const int32_t *idx_first_d, int32_t *counts_d,
                          void *stream);

/* The evaluation schedule on the device (evo/fitness.py:52-59 pairings, :100-121 games per pairing): game index
 * game_lo + t -> idx_first_d[t], idx_second_d[t] (rows of the weight table) and seeds_d[t] for t < n.  A pairing plays
 * games_per_pair consecutive games.  mode 0 = the reference's round robin: ordered pairs (i, j), i < n_ind, j < n_total,
 * j != i (rows n_ind.. are the hall of fame); mode 1 = every individual as FIRST against each of the n_total - n_ind fixed
 * opponents; mode 2 = (i, i), for a SECOND seat played by expert_action.  seeds are the partition-invariant hash of
 * (base_seed, generation, i, j, replicate) (the reference seeds every game from OS entropy). */
int sb_eval_schedule(SbHandle *h, int mode, int n_ind, int n_total, int games_per_pair, uint64_t base_seed, uint32_t generation,
                     int64_t game_lo, int n, int32_t *idx_first_d, int32_t *idx_second_d, uint64_t *seeds_d, void *stream);

/* FitnessEvaluator.evaluate_population (evo/fitness.py:32-121) for the games [game_lo, game_hi) of a schedule (a rank's shard),
 * entirely on the device: sb_eval_schedule -> sb_reset (decks_d u8[2,n_deck] shared by all games, factions_d u8[2]) ->
 * sb_rollout_heuristic over weights_d f64[n_total,10] -> sb_accumulate_fitness into counts_d i32[n_ind,3] (+=) and
 * sb_count_aborted into aborted_d i32[2] (+=, nullable), chunk_games at a time (<= 0: 262,144) in a workspace owned by the
 * handle.  Asynchronous on `stream`; nothing is copied to the host. */
int sb_eval_population(SbHandle *h, int mode, int n_ind, int n_total, int games_per_pair, uint64_t base_seed, uint32_t generation,
                       int64_t game_lo, int64_t game_hi, const double *weights_d, const uint8_t *decks_d, int n_deck,
                       const uint8_t *factions_d, int max_steps, int chunk_games, int32_t *counts_d, int32_t *aborted_d, void *stream);

/* Games a rollout aborted (result -2), split by cause: out_d i32[2] += {games stopped by an exception the reference raises
 * too (state.err 1..4; evo/fitness.py:208-210 scores them as a draw as well), games stopped by a limit of this engine
 * (state.err 5..7: SB_ERR_UNSUPPORTED / OVERFLOW / DEPTH -- the reference would have kept playing)}.  The evaluator
 * reports both and warns when the second exceeds 1 % of the games. */
int sb_count_aborted(SbHandle *h, int n, const uint8_t *states_d, const int8_t *result_d, int32_t *out_d, void *stream);

/* Host-buffer variants (pinned or pageable host memory; H2D + kernel + D2H + sync inside): the e2e path. */
int sb_step_host(SbHandle *h, int n, uint8_t *states, const uint8_t *actions, int8_t *reward, uint8_t *done,
                 uint8_t *err, uint32_t *next_masks);
int sb_rollout_random_host(SbHandle *h, int n, const uint64_t *seeds, const uint8_t *decks, int n_deck,
                           const uint8_t *factions, int max_steps, uint8_t *states_out, int32_t *steps_out,
                           uint64_t *chain_out);

#ifdef __cplusplus
}
#endif
#endif
