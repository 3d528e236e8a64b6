// sb_defs.h -- constants, enums and the card-statics record shared by the thread-per-game engine (sb_engine.cuh), the
// warp-per-game engine (sbw_core.cuh) and its host build for the tests (tests/wsim).  Plain C++, no CUDA types.
#pragma once
#include <stdint.h>
#include "../../include/sb_state.h"
#include "sb_card_ids.h"

typedef signed char i8;
typedef unsigned char u8;
typedef short i16;
typedef unsigned short u16;
typedef unsigned int u32;


#define MAXE 48
#define MAXTRIG 32
#define MAXPATH 8
#define MAXDEPTH 60
#define NMEM 12
#define NMEM_PACKED 9
#define NOBJ_PACKED 4
#define WT_N 1024
#define HAND_W 6
#define DECK_W 20

enum { KIND_UNIT = 0, KIND_STRUCTURE = 1, KIND_SPELL = 2 };
enum { TR_ON_PLAY = 0, TR_ON_DEATH, TR_BEFORE_ATTACKING, TR_AFTER_ATTACKING, TR_AFTER_SURVIVING,
       TR_BEFORE_MOVING, TR_TURN_START, TR_TURN_END, TR_NONE = 255 };
enum { PH_TURN_START = 0, PH_PLAY = 1, PH_TURN_END = 2 };
enum { TK_UNIT = 0, TK_STRUCTURE = 1, TK_ANY = 2 };
enum { TS_FRIENDLY = 0, TS_ENEMY = 1, TS_ANY = 2 };
enum { UT_CONSTRUCT = 0, UT_FLAKE, UT_KNIGHT, UT_PIRATE, UT_RAVEN, UT_RODENT, UT_SATYR, UT_TOAD, UT_UNDEAD,
       UT_VIKING, UT_HERO, UT_DRAGON, UT_ELDER, UT_FELINE, UT_ANCIENT, UT_PRIMAL };

#define PT_BASE_REMOTE 20  // Point(-1,-1)
#define PT_BASE_LOCAL 21   // Point(-1, 5)
#define PT_NONE (-1)

// card statics, 24 bytes (host builds it from sb_card_table.inc in sb_host.cu)
#define DCF_FIXED 1
#define DCF_ABILITY 2
#define DCF_TARGET 4
#define DCF_TBASE 8
#define DCF_TNONHERO 16
struct DCard {
  u8 kind; i8 cost; i8 strength; u8 movement; u8 trigger; u8 flags; u8 first_type; u8 t_ks;
  u16 types; i16 obs_id; u16 t_types; u16 t_xtypes; u8 t_status; u8 t_xstatus; i8 t_limit; i8 p[4]; u8 pad;
};
static_assert(sizeof(DCard) == 24, "DCard");

