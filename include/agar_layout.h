/*
 * agar_layout.h — host arithmetic that turns an AgarConfig into the env-record layout.
 * Header-only so that the GPU library and the CPU oracle compute byte-identical layouts
 * without linking one another.
 *
 * Sizes follow src/model/field.py:57-66 and src/model/networkParameters.py:75-102.
 */
#ifndef AGAR_LAYOUT_H
#define AGAR_LAYOUT_H

#include <math.h>
#include <string.h>
#include "agar_b200.h"

static inline uint64_t agar__align(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }

static inline int agar_layout_compute(const AgarConfig* c, AgarLayout* L) {
    memset(L, 0, sizeof(*L));
    int K = c->n_players;
    if (K < 1 || K > AGAR_MAX_PLAYERS) return AGAR_E_INVALID;
    if (c->grid_squares < 1 || c->grid_squares > 84) return AGAR_E_INVALID;
    if (c->frame_skip < 0) return AGAR_E_INVALID;
    if (c->enable_greedy_split) return AGAR_E_UNSUPPORTED;
    if (c->enable_eject && !c->enable_split) return AGAR_E_UNSUPPORTED; /* reference raises TypeError, bot.py:568 */
    /* VIRUS_GRID without VIRUS_SPAWN: the reference assigns gsVirus = None into the grid, a channel of NaN
     * (bot.py:380-382, 481); networkParameters.py derives VIRUS_GRID from VIRUS_SPAWN, so only a hand-edited file gets here */
    if (c->virus_grid && !c->virus_enabled) return AGAR_E_UNSUPPORTED;
    /* ALL_PLAYER_GRID replaces the self / enemy grids (networkParameters.py:89-91); with both, the reference stores None
     * (NaN) channels and None history grids (bot.py:372-378, 483-494) */
    if (c->all_player_grid && (c->self_grid || c->enemy_grid || c->self_grid_lf || c->self_grid_slf || c->enemy_grid_lf ||
                               c->enemy_grid_slf))
        return AGAR_E_UNSUPPORTED;
    int n_agents = 0;
    for (int k = 0; k < K; ++k) {
        int t = c->bot_type[k];
        if (t < AGAR_BOT_NN || t > AGAR_BOT_RANDOM) return AGAR_E_INVALID;
        if (t == AGAR_BOT_NN) {
            if (k != n_agents) return AGAR_E_INVALID; /* NN bots come first, aigar.py:778-780 */
            ++n_agents;
        }
    }
    /* field.py:58  size = int(SIZE_INCREASE_PER_PLAYER * math.sqrt(len(players))) */
    int S = (int)(75.0 * sqrt((double)K));
    if (S > 1023) return AGAR_E_RANGE;
    L->field_size = S;
    L->n_players = K;
    L->n_agents = n_agents;
    L->cell_cap = (c->enable_split || c->virus_enabled) ? AGAR_MAX_CELLS : 1;
    /* field.py:65-66 */
    L->max_pellets = c->pellet_spawn ? (double)(S * S) * 0.015 : 0.0;
    L->max_viruses = (double)(S * S) * 0.00005;
    int P = 0;
    while ((double)P < L->max_pellets) ++P; /* while len(pellets) < maxCollectibleCount: spawn */
    if (c->pellet_cap > P) P = c->pellet_cap; /* extra slots never get refilled: spawning stops at max_pellets */
    L->pellet_cap = P;
    int V0 = 0;
    while ((double)V0 < L->max_viruses) ++V0;
    L->virus_cap = c->virus_enabled ? (c->virus_cap > 0 ? c->virus_cap : 2 * V0 + 6) : 0;
    L->blob_cap = c->enable_eject ? (c->blob_cap > 0 ? c->blob_cap : 8 * K + 8) : 0;
    L->fat_cap = c->enable_eject ? (c->fat_cap > 0 ? c->fat_cap : 16 * K + 16) : 0;
    L->event_cap = c->event_cap > 0 ? c->event_cap : 0;
    L->grid_squares = c->grid_squares;
    /* networkParameters.py:98-102 */
    L->n_grids = (c->pellet_grid != 0) + (c->self_grid != 0) + (c->wall_grid != 0) + (c->virus_grid != 0) +
                 (c->enemy_grid != 0) + (c->self_grid_lf != 0) + (c->self_grid_slf != 0) + (c->enemy_grid_lf != 0) +
                 (c->enemy_grid_slf != 0) + (c->all_player_grid != 0);
    L->n_extra = (c->use_fovsize != 0) + (c->use_totalmass != 0) + 4 * (c->use_last_action != 0) +
                 4 * (c->use_second_last_action != 0) + (c->use_last_fovsize != 0);
    L->state_len = c->grid_squares * c->grid_squares * L->n_grids + L->n_extra;
    if (c->simple_state) L->n_grids = 0, L->n_extra = 0, L->state_len = AGAR_SIMPLE_STATE_LEN; /* bot.py:511-548 */
    L->action_len = 2 + (c->enable_split != 0) + (c->enable_eject != 0);
    /* history grids (bot.py:479-495) only exist when a last-frame channel is enabled */
    L->n_hist = (!c->simple_state && (c->self_grid_lf || c->self_grid_slf || c->enemy_grid_lf || c->enemy_grid_slf)) ? 4 : 0;
    uint64_t off = 0;
    L->off_header = off;
    off = agar__align(off + sizeof(AgarEnvHeader), 16);
    L->off_players = off;
    off = agar__align(off + (uint64_t)K * sizeof(AgarPlayer), 16);
    L->off_cells = off;
    off = agar__align(off + (uint64_t)K * L->cell_cap * sizeof(AgarCell), 16);
    L->off_viruses = off;
    off = agar__align(off + (uint64_t)L->virus_cap * sizeof(AgarMote), 16);
    L->off_blobs = off;
    off = agar__align(off + (uint64_t)L->blob_cap * sizeof(AgarMote), 16);
    L->off_fat = off;
    off = agar__align(off + (uint64_t)L->fat_cap * sizeof(AgarFatPellet), 16);
    L->off_pellets = off;
    off = agar__align(off + (uint64_t)P * sizeof(uint32_t), 16);
    L->off_hist = off;
    off = agar__align(off + (uint64_t)n_agents * L->n_hist * c->grid_squares * c->grid_squares * sizeof(float), 16);
    L->off_events = off;
    off = off + (uint64_t)L->event_cap * sizeof(AgarEvent);
    L->record_bytes = agar__align(off, 128);
    return AGAR_OK;
}

#endif /* AGAR_LAYOUT_H */
