/* sb_state.h -- packed Stormbound game state (512 bytes per game), the data format on both sides of
 * the drop-in boundary.  Replaces the reference's object graph:
 *   board.py:16-28 (Board), player.py:13-37 (Player), unit.py:8-23 (Unit), structure.py:8-16,
 *   card.py:13-20 (Card), games/stormbound.py:293-304 (Stormbound).
 * Players are indexed by ORDER (0 FIRST, 1 SECOND; enums.py:3-5).  Tiles are stored in the current
 * board orientation, index = y*4+x (board.py:17); a PASS rotates the array like Board.flip
 * (board.py:94-115).  Plain C, no CUDA or torch types.  oracle/sb_layout.py mirrors this file.
 */
#ifndef SB_STATE_H
#define SB_STATE_H
#include <stdint.h>

#define SB_STATE_BYTES 512
#define SB_N_TILES 20
#define SB_HAND_MAX 4
#define SB_DECK_MAX 16
#define SB_EXT_BYTES 112
#define SB_N_ACTIONS 156        /* enums.py:10-36, actions.txt */
#define SB_MASK_WORDS 5         /* 156-bit legal-action mask */
#define SB_OBS_INTS 540         /* 27 x 5 x 4 int32, games/stormbound.py:400-526 */
#define SB_N_FEATURES 10        /* evo/features.py:327-342 */
#define SB_ACTION_PASS 155

/* player flags */
#define SB_PF_REPLACABLE 1
#define SB_PF_LEFTMOST 2
/* hand / deck card flags */
#define SB_CF_FIXED 1           /* Unit.fixedly_forward of the card object (cards/b008.py:15-16) */
#define SB_CF_SINGLE_USE 2      /* Card.is_single_use (cards/ua20.py:31) */
#define SB_CF_OBJ 4             /* the record IS a (former) board instance of B305 (cards/b305.py:41-45); see ext */
/* tile flags */
#define SB_TF_OWNER 1
#define SB_TF_STRUCTURE 2
#define SB_TF_FIXED 4
/* tile status: five 6-bit counters in StatusEffect order (enums.py:77-82); a multiset, unit.py:239-275 */
#define SB_ST_BITS 6
#define SB_ST_FROZEN 0
#define SB_ST_POISONED 1
#define SB_ST_CONFUSED 2
#define SB_ST_DISABLED 3
#define SB_ST_VITALIZED 4
/* state.done bits */
#define SB_DONE 1               /* games/stormbound.py:365 */
#define SB_REWARD 2             /* games/stormbound.py:366 */
/* state.err: non-zero = the reference raises a Python exception at this point */
#define SB_ERR_NONE 0
#define SB_ERR_NONE_TARGET 1
#define SB_ERR_EMPTY_CHOICE 2
#define SB_ERR_INDEX 3
#define SB_ERR_OBS_ID 4
#define SB_ERR_UNSUPPORTED 5
#define SB_ERR_OVERFLOW 6
#define SB_ERR_DEPTH 7

typedef struct SbPlayer {
  int16_t base;                 /* Player.strength */
  int16_t max_mana;
  int16_t mana;                 /* Player.current_mana */
  int8_t front_line;
  uint8_t flags;
  uint8_t n_hand;
  uint8_t n_deck;
  uint8_t faction;
  uint8_t pad;
  uint8_t hand_card[SB_HAND_MAX];
  int8_t hand_cost[SB_HAND_MAX];
  uint8_t hand_flags[SB_HAND_MAX];
  uint8_t deck_card[SB_DECK_MAX];
  int8_t deck_cost[SB_DECK_MAX];
  uint8_t deck_flags[SB_DECK_MAX];
  uint16_t deck_wn[SB_DECK_MAX];  /* weight = f^n(1), f(w) = w*1.6+100 (player.py:32,59) */
} SbPlayer;                       /* 104 bytes */

typedef struct SbTile {
  uint8_t card;                 /* 0 = empty; index into the card table */
  uint8_t flags;
  int16_t strength;
  uint32_t status;
} SbTile;                       /* 8 bytes */

typedef struct SbState {
  uint32_t seed_lo, seed_hi;    /* Philox key */
  uint16_t turn;                /* PASS actions so far = Philox counter word 1 */
  uint16_t draw;                /* draws this turn      = Philox counter word 0 */
  uint16_t steps;               /* env steps taken */
  uint8_t local_order;          /* order of Board.local */
  uint8_t current_order;        /* order of Board.current_player */
  int8_t player_sign;           /* Stormbound.player, games/stormbound.py:304 */
  uint8_t phase;
  uint8_t err;
  uint8_t done;
  uint8_t hist_n;               /* Board.history[-4:], oldest first */
  uint8_t hist_card[4];
  uint8_t hist_owner[4];
  uint8_t pad[3];
  SbPlayer pl[2];
  SbTile tile[SB_N_TILES];
  /* ext[0] = n B005 memories, ext[1..90] = 9 x {key, pos, card, tile flags, strength i16, status u32}
   * (cards/b005.py:13,24-33); key = tile of the owning temple, or 0x80 | index of the remembered temple COPY
   * (an earlier record) whose own memory this is -- deep copies keep their lists; pre-order, temples by tile; ext[91] = n board-instance card records, ext[92..107] = 4 x {order<<7 |
   * in_deck<<6 | index, tile or 0xFF, frozen strength i16} (cards/b305.py:41-45). */
  uint8_t ext[SB_EXT_BYTES];
} SbState;

#ifdef __cplusplus
static_assert(sizeof(SbPlayer) == 104, "SbPlayer layout");
static_assert(sizeof(SbTile) == 8, "SbTile layout");
static_assert(sizeof(SbState) == SB_STATE_BYTES, "SbState layout");
#else
_Static_assert(sizeof(SbPlayer) == 104, "SbPlayer layout");
_Static_assert(sizeof(SbTile) == 8, "SbTile layout");
_Static_assert(sizeof(SbState) == SB_STATE_BYTES, "SbState layout");
#endif
#endif
