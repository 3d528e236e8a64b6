/* TEST INFRASTRUCTURE -- see sb_oracle.h.  Observation (games/stormbound.py:400-526), the ten
 * StateFeatures (evo/features.py:12-342), HeuristicAgent scoring / argmax
 * (evo/heuristic_agent.py:23-122) and the intended game loop of evo/fitness.py:178-228, on the CPU.
 */
#include <string.h>
#include <stdlib.h>
#include <pthread.h>
#include <math.h>
#include "sb_oracle.h"

void sbo_step(SbState *s, int action);
void sbo_legal_mask(const SbState *s, uint32_t *mask);
uint32_t sbo_agent_pick(uint64_t seed, uint32_t step, uint32_t n);
uint64_t sbo_digest(const SbState *s);

#define OBS(l, r, c) obs[((l) * 5 + (r)) * 4 + (c)]

/* strength of a (former) board instance of B305 sitting in hand/deck: see CardRec in sb_oracle.h */
static int obj_strength(const SbState *s, int order, int in_deck, int idx, int dflt) {
  const uint8_t *x = s->ext;
  for (int i = 0; i < x[91] && i < NOBJ_PACKED; i++) {
    const uint8_t *r = x + 92 + 4 * i;
    if (r[0] == (uint8_t)((order << 7) | (in_deck << 6) | idx))
      return r[1] != 0xFF ? s->tile[r[1]].strength : (int16_t)(r[2] | (r[3] << 8));
  }
  return dflt;
}
static void card_row(int32_t *row, int card, int cost, int *err) {
  const OCard *c = &OCARDS[card];
  if (c->obs_id == -32768) *err = SB_ERR_OBS_ID; /* card.py:46 ValueError (Q12) */
  row[0] = c->obs_id;
  row[1] = cost;
  row[2] = c->kind == KIND_SPELL ? -1 : c->strength;
  row[3] = c->kind == KIND_UNIT ? c->movement : -1;
}

/* games/stormbound.py:400-526.  Returns 0 or SB_ERR_OBS_ID. */
int sbo_observe(const SbState *s, int32_t *obs) {
  int err = 0;
  int lo = s->local_order, ro = 1 - lo;
  for (int i = 0; i < SB_OBS_INTS; i++) obs[i] = -1;
  for (int side = 0; side < 2; side++) {
    int order = side == 0 ? lo : ro;
    int base = side == 0 ? 0 : 16;
    for (int y = 0; y < 5; y++) for (int x = 0; x < 4; x++) {
      const SbTile *t = &s->tile[y * 4 + x];
      if (!t->card || ((t->flags & SB_TF_OWNER) ? 1 : 0) != order) continue;
      const OCard *c = &OCARDS[t->card];
      if (c->obs_id == -32768) err = SB_ERR_OBS_ID;
      if (!(t->flags & SB_TF_STRUCTURE)) {
        OBS(base + 0, y, x) = c->obs_id;
        OBS(base + 1, y, x) = t->strength;
        OBS(base + 2, y, x) = c->movement;
        int st = 0;
        if ((t->status >> (SB_ST_BITS * SB_ST_VITALIZED)) & 63) st |= 1;
        if ((t->status >> (SB_ST_BITS * SB_ST_POISONED)) & 63) st |= 2;
        if ((t->status >> (SB_ST_BITS * SB_ST_CONFUSED)) & 63) st |= 4;
        if ((t->status >> (SB_ST_BITS * SB_ST_FROZEN)) & 63) st |= 8;
        if ((t->status >> (SB_ST_BITS * SB_ST_DISABLED)) & 63) st |= 16;
        OBS(base + 3, y, x) = st;
      } else {
        OBS(base + 4, y, x) = c->obs_id;
        OBS(base + 5, y, x) = t->strength;
      }
    }
  }
  const SbPlayer *L = &s->pl[lo], *R = &s->pl[ro];
  for (int i = 0; i < 4; i++) if (i < L->n_hand) {
    card_row(&OBS(6, i, 0), L->hand_card[i], L->hand_cost[i], &err);
    if (L->hand_flags[i] & SB_CF_OBJ) OBS(6, i, 2) = obj_strength(s, lo, 0, i, OBS(6, i, 2));
  }
  for (int c = 0; c < 4; c++) OBS(6, 4, c) = 32767;
  /* deck sorted by (cost, card_id), stable; card index order == card_id order */
  int idx[SB_DECK_MAX];
  for (int i = 0; i < L->n_deck; i++) idx[i] = i;
  for (int i = 1; i < L->n_deck; i++) {
    int v = idx[i], j = i - 1;
    while (j >= 0 && (L->deck_cost[idx[j]] > L->deck_cost[v] ||
                      (L->deck_cost[idx[j]] == L->deck_cost[v] && L->deck_card[idx[j]] > L->deck_card[v]))) { idx[j + 1] = idx[j]; j--; }
    idx[j + 1] = v;
  }
  for (int layer = 0; layer < 6; layer++) {
    for (int k = 0; k < 4; k++) {
      int d = layer * 4 + k;
      if (d < L->n_deck) {
        card_row(&OBS(7 + layer, k, 0), L->deck_card[idx[d]], L->deck_cost[idx[d]], &err);
        if (L->deck_flags[idx[d]] & SB_CF_OBJ) OBS(7 + layer, k, 2) = obj_strength(s, lo, 1, idx[d], OBS(7 + layer, k, 2));
      }
    }
    for (int c = 0; c < 4; c++) OBS(7 + layer, 4, c) = 32768;
  }
  for (int y = 0; y < 5; y++) for (int x = 0; x < 4; x++) {
    OBS(13, y, x) = L->mana; OBS(14, y, x) = L->base; OBS(15, y, x) = L->faction;
    OBS(22, y, x) = R->mana; OBS(23, y, x) = R->base; OBS(24, y, x) = R->faction;
    OBS(25, y, x) = s->player_sign * 99999;
  }
  for (int i = 0; i < 4; i++) { /* ([None]*4 + history)[-4:] */
    int h = i - (4 - s->hist_n);
    if (h >= 0) {
      OBS(26, i, 0) = s->hist_owner[h] ? -99999 : 99999;
      OBS(26, i, 1) = OCARDS[s->hist_card[h]].obs_id;
      if (OCARDS[s->hist_card[h]].obs_id == -32768) err = SB_ERR_OBS_ID;
    }
  }
  for (int c = 0; c < 4; c++) OBS(26, 4, c) = 32769;
  return err;
}

static double clip01(double v) { return v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v); }

/* evo/features.py:12-342 from the observation layers it reads */
void sbo_features_from_obs(const int32_t *obs, double *f) {
  double m = OBS(13, 0, 0) != -1 ? (double)OBS(13, 0, 0) : 0.0;
  double hl = OBS(14, 0, 0) != -1 ? (double)OBS(14, 0, 0) : 20.0;
  double hr = OBS(23, 0, 0) != -1 ? (double)OBS(23, 0, 0) : 20.0;
  double est = m + 2.0; if (est < 3.0) est = 3.0; if (est > 10.0) est = 10.0;
  f[0] = clip01(1.0 - (m / est));
  f[1] = hl - hr;
  long sl = 0, sr = 0;
  int nl = 0, nr = 0, nsl = 0, nsr = 0, minl = 99, maxr = -1;
  double threat = 0.0, prot = 0.0;
  for (int y = 0; y < 5; y++) for (int x = 0; x < 4; x++) {
    if (OBS(1, y, x) != -1) sl += OBS(1, y, x);
    if (OBS(5, y, x) != -1) sl += OBS(5, y, x);
    if (OBS(17, y, x) != -1) sr += OBS(17, y, x);
    if (OBS(21, y, x) != -1) sr += OBS(21, y, x);
    if (OBS(0, y, x) != -1) { nl++; if (y < minl) minl = y; }
    if (OBS(16, y, x) != -1) { nr++; if (y > maxr) maxr = y; }
    if (OBS(4, y, x) != -1) nsl++;
    if (OBS(20, y, x) != -1) nsr++;
    if (OBS(17, y, x) != -1) threat += (double)OBS(17, y, x) * ((double)(y + 1) / 5.0);
    double dw = (double)(5 - y) / 5.0;
    if (OBS(1, y, x) != -1) prot += (double)OBS(1, y, x) * dw;
    if (OBS(5, y, x) != -1) prot += (double)OBS(5, y, x) * dw;
  }
  long tot = sl + sr;
  f[2] = tot == 0 ? 0.0 : (double)(sl - sr) / (double)tot;
  if (nl == 0 && nr == 0) f[3] = 0.0;
  else f[3] = (double)((nr ? maxr : 0) - (nl ? minl : 4)) / 4.0;
  f[4] = (double)(sl - sr);
  f[5] = (double)(nl - nr);
  f[6] = (double)(nsl - nsr);
  f[7] = threat;
  f[8] = prot;
  int playable = 0, valid = 0;
  double total = 0.0;
  for (int i = 0; i < 4; i++) {
    int cid = OBS(6, i, 0);
    if (cid != -1 && cid != 32767) {
      int cost = OBS(6, i, 1);
      int str = OBS(6, i, 2) != -1 ? OBS(6, i, 2) : 0;
      valid++;
      if (cost > 0) {
        total += (double)str / (double)cost;
        if ((double)cost <= m) playable++;
      }
    }
  }
  if (valid == 0) f[9] = 0.0;
  else {
    double playability = (double)playable / (double)valid;
    double avg = total / (double)valid;
    f[9] = (playability + clip01(avg / 3.0)) / 2.0;
  }
}
int sbo_features(const SbState *s, double *f) {
  int32_t obs[SB_OBS_INTS];
  int err = sbo_observe(s, obs);
  sbo_features_from_obs(obs, f);
  return err;
}

/* evo/heuristic_agent.py:23-51: score = w.(-d) - w.d - rp for every legal action; exception -> 0.0.
 * scores[a] is written for legal a only.  Returns the argmax action (first maximum, :67-68). */
int sbo_select_action(const SbState *s, const double *w, double *scores, uint32_t *mask_out) {
  uint32_t m[SB_MASK_WORDS];
  double fc[SB_N_FEATURES], fn[SB_N_FEATURES];
  int cur_err = sbo_features(s, fc);
  sbo_legal_mask(s, m);
  if (mask_out) memcpy(mask_out, m, sizeof m);
  int best = -1;
  double best_score = 0.0;
  for (int a = 0; a < SB_N_ACTIONS; a++) {
    if (!(m[a >> 5] >> (a & 31) & 1)) continue;
    SbState nx = *s;
    sbo_step(&nx, a);
    double sc = 0.0;
    int nerr = nx.err;
    if (!nerr) nerr = sbo_features(&nx, fn);
    if (!nerr && !cur_err) {
      double d = 0.0;
      /* np.dot for n = 10 (OpenBLAS ddot, n < 32: scalar tail loop compiled with FMA contraction) ==
       * sequential fused multiply-add; verified bit-exact against numpy 2.3.5 / OpenBLAS 0.3.30 */
      for (int i = 0; i < SB_N_FEATURES; i++) d = fma(w[i], fn[i] - fc[i], d);
      double eff = fn[0] - fc[0];
      double rp = eff < -0.3 ? (eff < 0 ? -eff : eff) * 0.2 : 0.0;
      sc = (-d) - d - rp;
    }
    if (scores) scores[a] = sc;
    if (best < 0 || sc > best_score) { best = a; best_score = sc; }
  }
  return best < 0 ? SB_ACTION_PASS : best;
}

int sbo_expert_action(SbState *s);
/* Intended loop of evo/fitness.py:193-206 (the is_terminal bug Q15 bypassed): until have_winner or
 * max_steps.  Returns: 0 FIRST wins, 1 SECOND wins, -1 draw/timeout, -2 aborted by an engine exception. */
int sbo_play_heuristic(SbState *s, const double *w_first, const double *w_second, int max_steps,
                       uint8_t *actions, int *n_steps) {
  int k = 0;
  while (k < max_steps) {
    if (s->pl[0].base < 0 || s->pl[1].base < 0) break;
    int to_play = s->player_sign == 1 ? 0 : 1;
    const double *w = to_play == 0 ? w_first : w_second;
    /* a seat without weights is played by Stormbound.expert_action (games/stormbound.py:563-637; the agent-vs-expert
     * match of play_vs_expert.py:65-94): it draws from the game's own stream before the step */
    int a = w ? sbo_select_action(s, w, NULL, NULL) : sbo_expert_action(s);
    sbo_step(s, a);
    if (actions) actions[k] = (uint8_t)a;
    k++;
    if (s->err) { if (n_steps) *n_steps = k; return -2; }
  }
  if (n_steps) *n_steps = k;
  if (s->pl[0].base < 0 && s->pl[1].base >= 0) return 1;
  if (s->pl[1].base < 0 && s->pl[0].base >= 0) return 0;
  if (s->pl[0].base < 0 && s->pl[1].base < 0) return -1;
  return -1;
}

/* ------------------------------------------------------------------ threaded batch drivers (CPU baseline) */
typedef struct {
  SbState *states; int n, tid, nthreads, max_steps, mode;
  const double *w_first, *w_second; const int32_t *idx_first, *idx_second;
  int32_t *result, *steps; long total_steps;
} Job;
int sbo_rollout_random(SbState *s, int max_steps, uint8_t *actions, uint64_t *digests, uint32_t *masks);
static void *job_main(void *arg) {
  Job *j = (Job *)arg;
  long tot = 0;
  for (int i = j->tid; i < j->n; i += j->nthreads) {
    int k = 0;
    if (j->mode == 0) {
      k = sbo_rollout_random(&j->states[i], j->max_steps, NULL, NULL, NULL);
      if (j->result) j->result[i] = 0;
    } else {
      const double *wf = j->w_first + 10 * (j->idx_first ? j->idx_first[i] : 0);
      const double *ws = j->w_second + 10 * (j->idx_second ? j->idx_second[i] : 0);
      int r = sbo_play_heuristic(&j->states[i], wf, ws, j->max_steps, NULL, &k);
      if (j->result) j->result[i] = r;
    }
    if (j->steps) j->steps[i] = k;
    tot += k;
  }
  j->total_steps = tot;
  return NULL;
}
static long run_jobs(Job proto, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
  Job *jobs = (Job *)malloc(sizeof(Job) * nthreads);
  long tot = 0;
  for (int t = 0; t < nthreads; t++) { jobs[t] = proto; jobs[t].tid = t; jobs[t].nthreads = nthreads; pthread_create(&th[t], NULL, job_main, &jobs[t]); }
  for (int t = 0; t < nthreads; t++) { pthread_join(th[t], NULL); tot += jobs[t].total_steps; }
  free(th); free(jobs);
  return tot;
}
long sbo_batch_random(SbState *states, int n, int max_steps, int nthreads, int32_t *steps) {
  Job j; memset(&j, 0, sizeof j);
  j.states = states; j.n = n; j.max_steps = max_steps; j.mode = 0; j.steps = steps;
  return run_jobs(j, nthreads);
}
long sbo_batch_heuristic(SbState *states, int n, const double *w_first, const double *w_second,
                         const int32_t *idx_first, const int32_t *idx_second, int max_steps, int nthreads,
                         int32_t *result, int32_t *steps) {
  Job j; memset(&j, 0, sizeof j);
  j.states = states; j.n = n; j.max_steps = max_steps; j.mode = 1; j.steps = steps; j.result = result;
  j.w_first = w_first; j.w_second = w_second; j.idx_first = idx_first; j.idx_second = idx_second;
  return run_jobs(j, nthreads);
}
