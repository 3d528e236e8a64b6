// sb_card_table_host.h -- the generated card table (sb_card_table.inc, tools/gen_card_table.py) as host data, and its
// conversion to the 24-byte device record.  Shared by the C ABI (sb_kernels.cu) and the host build of the warp engine
// that the tests use (tests/wsim).
#pragma once
#include <string.h>
#include "sb_defs.h"

struct HostCard {
  int kind, faction, cost, strength, movement, trigger, fixed, has_ability, first_type, types, obs_id;
  int has_target, t_kind, t_side, t_types, t_xtypes, t_status, t_xstatus, t_limit, t_nonhero, t_base;
  int p[4];
};
static const HostCard HOST_CARDS[SBC_COUNT] = {
#include "sb_card_table.inc"
};

static inline void sb_build_dcards(DCard* tab) {
  memset(tab, 0, sizeof(DCard) * SBC_COUNT);
  for (int i = 0; i < SBC_COUNT; i++) {
    const HostCard& c = HOST_CARDS[i];
    DCard& d = tab[i];
    d.kind = (u8)c.kind; d.cost = (i8)c.cost; d.strength = (i8)c.strength; d.movement = (u8)c.movement; d.trigger = (u8)c.trigger;
    d.flags = (u8)((c.fixed ? DCF_FIXED : 0) | (c.has_ability ? DCF_ABILITY : 0) | (c.has_target ? DCF_TARGET : 0) |
                   (c.t_base ? DCF_TBASE : 0) | (c.t_nonhero ? DCF_TNONHERO : 0));
    d.first_type = (u8)c.first_type; d.t_ks = (u8)(c.t_kind | (c.t_side << 2));
    d.types = (u16)c.types; d.obs_id = (i16)c.obs_id; d.t_types = (u16)c.t_types; d.t_xtypes = (u16)c.t_xtypes;
    d.t_status = (u8)c.t_status; d.t_xstatus = (u8)c.t_xstatus; d.t_limit = (i8)c.t_limit;
    for (int k = 0; k < 4; k++) d.p[k] = (i8)c.p[k];
  }
}
// player.py:32,59: w*1.6+100 with two roundings (volatile blocks FMA contraction)
static inline void sb_build_weights(double* wt) {
  volatile double w = 1.0;
  for (int i = 0; i < WT_N; i++) { wt[i] = w; volatile double m = w * 1.6; w = m + 100.0; }
}
