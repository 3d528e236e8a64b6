/* TEST INFRASTRUCTURE -- CPU restatement (plain C) of the reference's rules engine, observation,
 * features and heuristic scoring.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this; the product path never does.
 *
 * Parity status: PINNED against the live reference (oracle/validate_vs_reference.py, run in the build
 * container) and against the committed golden tapes under tests/golden/ that the reference produced
 * with the injected Philox stream (tests/golden/make_golden.py).
 *
 * Structure follows the reference one-to-one (object-per-entity, real recursion), NOT the CUDA
 * engine's packed working set: it exists to be obviously equal to the Python.
 */
#ifndef SB_ORACLE_H
#define SB_ORACLE_H
#include <stdint.h>
#include "../include/sb_state.h"
#include "sb_card_ids.h"

#define MAXE 48      /* entity pool per step (same bound as the CUDA working set) */
#define MAXTRIG 32
#define MAXPATH 8
#define HAND_W 6
#define DECK_W 20
#define MAXDEPTH 60

enum { KIND_UNIT = 0, KIND_STRUCTURE = 1, KIND_SPELL = 2 };
enum { TR_ON_PLAY = 0, TR_ON_DEATH, TR_BEFORE_ATTACKING, TR_AFTER_ATTACKING, TR_AFTER_SURVIVING,
       TR_BEFORE_MOVING, TR_TURN_START, TR_TURN_END, TR_NONE = 255 };
enum { PH_TURN_START = 0, PH_PLAY = 1, PH_TURN_END = 2 };
enum { TK_UNIT = 0, TK_STRUCTURE = 1, TK_ANY = 2 };
enum { TS_FRIENDLY = 0, TS_ENEMY = 1, TS_ANY = 2 };
enum { UT_CONSTRUCT = 0, UT_FLAKE, UT_KNIGHT, UT_PIRATE, UT_RAVEN, UT_RODENT, UT_SATYR, UT_TOAD, UT_UNDEAD,
       UT_VIKING, UT_HERO, UT_DRAGON, UT_ELDER, UT_FELINE, UT_ANCIENT, UT_PRIMAL };

/* point ids: 0..19 = tile y*4+x; the two base points of point.py:20-21 */
#define PT_BASE_REMOTE 20   /* Point(-1,-1) */
#define PT_BASE_LOCAL 21    /* Point(-1, 5) */
#define PT_NONE (-1)

typedef struct {
  int kind, faction, cost, strength, movement, trigger, fixed, has_ability, first_type, types, obs_id;
  int has_target, t_kind, t_side, t_types, t_xtypes, t_status, t_xstatus, t_limit, t_nonhero, t_base;
  int p[4];
} OCard;
extern const OCard OCARDS[SBC_COUNT];

typedef struct {  /* target.py:13-24 */
  int kind, side, types, xtypes, status, xstatus, has_limit, limit, nonhero, base;
} Target;

typedef struct { int8_t x, y; } XY;

typedef struct {  /* unit.py:8-23 / structure.py:8-16 */
  int card, owner, is_struct, fixed, single_use;
  int strength, movement, trigger, types, has_ability;
  int st[5];
  int x, y;
  XY path[MAXPATH];
  int path_len;
  int damage_taken;
  int move_id, resolving_play;
} Ent;

/* xstr/link: only for records that are (former) BOARD INSTANCES of B305 (cards/b305.py:41-45 appends
 * the board object itself to the hand): link = entity id while that object exists this step, else xstr
 * is the object's frozen strength.  flags carries SB_CF_OBJ for them. */
typedef struct { int card, cost, flags, wn, xstr, link; } CardRec;
/* B005 per-instance memory (cards/b005.py:13,24-33): deep copies of neighbouring friendly entities */
/* parent < 0: a memory of the live temple entity `b005`; parent >= 0: a memory held BY the remembered temple
 * copy mem[parent] (deepcopy keeps the copy's own ability_remembered list); parent index < own index always */
typedef struct { int b005, parent, pos, card, owner, is_struct, fixed, detached, strength, st[5]; } Mem;
/* detached: the record was deep-copied as part of ANOTHER temple's copy.  Card.copy re-points only the top
 * object's .player (card.py:71-75); everything below keeps the deep-copied Player/Board, so such an entity
 * would act on a cloned board once restored.  Restoring one is not modelled: SB_ERR_UNSUPPORTED. */
#define NMEM_W 12
#define NMEM_PACKED 9
#define NOBJ_PACKED 4

typedef struct {  /* player.py:13-37 */
  int base, max_mana, mana, front_line, replacable, leftmost, faction;
  int n_hand, n_deck;
  CardRec hand[HAND_W]; /* working capacity > packed capacity (a cycle holds 17 deck cards for a moment) */
  CardRec deck[DECK_W];
} Ply;

typedef struct {
  Ent e[MAXE];
  int n_ent;
  int board[5][4]; /* entity id or -1 */
  Ply pl[2];       /* by order */
  int local_order, current_order, player_sign, phase, err, done, steps;
  int trig_ent[MAXTRIG], trig_src[MAXTRIG], n_trig, resolving;
  uint32_t seed_lo, seed_hi, turn, draw;
  int hist_n, hist_card[4], hist_owner[4];
  int depth;
  Mem mem[NMEM_W];
  int n_mem;
} Game;

#define ERR(g, code) do { if (!(g)->err) (g)->err = (code); } while (0)

/* engine (sb_oracle.c) */
void o_unpack(Game *g, const SbState *s);
void o_pack(const Game *g, SbState *s);
int o_rng_below(Game *g, int n);
double o_rng_random(Game *g);
void o_shuffle(Game *g, int *a, int n);
int o_at(const Game *g, int x, int y);
int o_at_pt(const Game *g, int pt);
void o_set(Game *g, int x, int y, int id);
Ply *o_local(Game *g);
Ply *o_remote(Game *g);
int o_opponent(const Game *g, int order);
void o_calc_front_line(Game *g, int order);
int o_get_targets(Game *g, int pov, const Target *t, int exclude_pt, int *out);
int o_front(Game *g, int x, int y, int pov, const Target *t, int *out);
int o_behind(Game *g, int x, int y, int pov, const Target *t, int *out);
int o_side(Game *g, int x, int y, int pov, const Target *t, int *out);
int o_row(Game *g, int x, int y, int pov, const Target *t, int *out);
int o_bordering(Game *g, int x, int y, int pov, const Target *t, int *out);
int o_surrounding(Game *g, int x, int y, int pov, const Target *t, int *out);
int o_new_ent(Game *g, int card, int owner, int strength);
int o_spawn_token_unit(Game *g, int owner, int pt, int strength, int type);
void o_ability(Game *g, int id, int pos_pt, int has_source);
void o_spell_ability(Game *g, int card, int caster, int pos_pt);
int o_unit_deal_damage(Game *g, int id, int amount, int pending, int has_source);
int o_struct_deal_damage(Game *g, int id, int amount, int pending, int has_source);
int o_deal_damage_pt(Game *g, int pt, int amount, int has_source);
int o_player_damage(Game *g, int order, int amount);
void o_destroy(Game *g, int id, int has_source);
void o_heal(Game *g, int id, int amount);
void o_st_add(Game *g, int id, int st);
void o_st_remove(Game *g, int id, int st);
void o_freeze(Game *g, int id);
void o_poison(Game *g, int id);
void o_vitalize(Game *g, int id);
void o_confuse(Game *g, int id);
void o_disable(Game *g, int id);
void o_set_path(Game *g, int id, int on_play);
void o_move(Game *g, int id);
void o_gain_speed(Game *g, int id, int amount);
void o_command(Game *g, int id);
void o_convert(Game *g, int id);
void o_push(Game *g, int id, int fx, int fy);
void o_force_attack(Game *g, int id, int dx, int dy);
void o_teleport(Game *g, int id, int dx, int dy);
void o_player_play(Game *g, int order, int index, int pos_pt);
void o_struct_play(Game *g, int id, int x, int y);
int o_is_within_front_line(const Game *g, int order, int y);
int o_get_within_front_line(const Game *g, int order, int *out);
/* effects (sb_oracle_effects.c) */
void o_effect(Game *g, int id, int pos_pt, int has_source);
void o_spell_effect(Game *g, int card, int caster, int pos_pt);

static inline int PTX(int pt) { return pt == PT_BASE_REMOTE || pt == PT_BASE_LOCAL ? -1 : (pt & 3); }
static inline int PTY(int pt) { return pt == PT_BASE_REMOTE ? -1 : pt == PT_BASE_LOCAL ? 5 : (pt >> 2); }
static inline int PT(int x, int y) { return y * 4 + x; }
static inline int valid_xy(int x, int y) { return x >= 0 && x <= 3 && y >= 0 && y <= 4; }
static inline int is_base_pt(int pt) { return pt == PT_BASE_REMOTE || pt == PT_BASE_LOCAL; }
#endif
