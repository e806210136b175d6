/*
 * agar_b200.cu — kernels + the C ABI (include/agar_b200.h) of the B200-native batched agar.io step.
 * Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -shared -Xcompiler -fPIC
 * No torch, no CPU fallback: every entry point runs CUDA kernels on the handle's device or fails.
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/agar_b200.h"
#include "../../include/agar_layout.h"
#include "agar_bots.cuh"
#include "agar_dev.cuh"
#include "agar_simple.cuh"

extern __shared__ __align__(16) uint8_t g_smem[];

/* ------------------------------------------------------------------ staging: HBM <-> shared memory, whole CTA */
__device__ __forceinline__ void stage_in(const DevParams& P, const uint8_t* state, int env0, int n_here, bool zero) {
    const int chunks = (int)(P.L.record_bytes / 8);
    const int total = n_here * chunks;
    const uint2* src = (const uint2*)(state + (size_t)env0 * P.L.record_bytes);
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        int e = i / chunks, w = i - e * chunks;
        uint2 v = zero ? make_uint2(0u, 0u) : src[i];
        ((uint2*)(g_smem + (size_t)e * P.rec_stride))[w] = v;
    }
    __syncthreads();
}
__device__ __forceinline__ void stage_out(const DevParams& P, uint8_t* state, int env0, int n_here) {
    __syncthreads();
    const int chunks = (int)(P.L.record_bytes / 8);
    const int total = n_here * chunks;
    uint2* dst = (uint2*)(state + (size_t)env0 * P.L.record_bytes);
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        int e = i / chunks, w = i - e * chunks;
        dst[i] = ((const uint2*)(g_smem + (size_t)e * P.rec_stride))[w];
    }
}
/* stage >= 2 ("hot sections"): big records stay in HBM / L2; what every scan and every lane-0 step reads is cached in
 * shared memory — the pellet pool, and with hot_a > 0 also bytes [0, hot_a) = header, players, cells, viruses */
__device__ __forceinline__ int hot_a_bytes(const DevParams& P) { return (P.hot_a + 15) / 16 * 16; }
__device__ __forceinline__ int pel_cache_bytes(const DevParams& P) { return hot_a_bytes(P) + (P.L.pellet_cap * 4 + 15) / 16 * 16; }
__device__ __forceinline__ void pel_cache_copy(const DevParams& P, uint8_t* state, int env0, int n_here, int tiles, bool in) {
    const int wa = P.hot_a / 4, words = wa + P.L.pellet_cap;
    for (int i = threadIdx.x; i < n_here * words; i += blockDim.x) {
        int e = i / words, w = i - e * words;
        uint8_t* rec = state + (size_t)(env0 + e) * P.L.record_bytes;
        uint8_t* hot = g_smem + (size_t)tiles * P.scratch_bytes + (size_t)e * pel_cache_bytes(P);
        uint32_t* g = w < wa ? (uint32_t*)rec + w : (uint32_t*)(rec + P.L.off_pellets) + (w - wa);
        uint32_t* s = w < wa ? (uint32_t*)hot + w : (uint32_t*)(hot + hot_a_bytes(P)) + (w - wa);
        if (in) *s = *g; else *g = *s;
    }
    __syncthreads();
}
template <int W>
__device__ __forceinline__ void bind_ctx(Ctx<W>& c, const DevParams& P, int tile_id, int tiles, int env, uint8_t* state) {
    c.lane = c.t.thread_rank();
    c.env_id = (uint32_t)(P.first_env + (uint64_t)env);
    c.rec = P.stage == 1 ? g_smem + (size_t)tile_id * P.rec_stride : state + (size_t)env * P.L.record_bytes;
    c.h = (AgarEnvHeader*)(c.rec + P.L.off_header);
    c.pl = (AgarPlayer*)(c.rec + P.L.off_players);
    c.cells = (AgarCell*)(c.rec + P.L.off_cells);
    c.vir = (AgarMote*)(c.rec + P.L.off_viruses);
    c.blob = (AgarMote*)(c.rec + P.L.off_blobs);
    c.fat = (AgarFatPellet*)(c.rec + P.L.off_fat);
    c.pel = (uint32_t*)(c.rec + P.L.off_pellets);
    c.hist = (float*)(c.rec + P.L.off_hist);
    c.ev = (AgarEvent*)(c.rec + P.L.off_events);
    c.scratch = g_smem + (P.stage == 1 ? (size_t)tiles * P.rec_stride : 0) + (size_t)tile_id * P.scratch_bytes;
    if (P.stage >= 2) {
        uint8_t* hot = g_smem + (size_t)tiles * P.scratch_bytes + (size_t)tile_id * pel_cache_bytes(P);
        c.pel = (uint32_t*)(hot + hot_a_bytes(P));
        /* header, players, cells, viruses are contiguous from byte 0 of the record: whichever of them end inside the cached
         * prefix [0, hot_a) are served from shared memory (stage 4 caches header + players only: the bot bookkeeping and the
         * command / fov fields every lane-0 bot turn walks through) */
        if (P.hot_a >= (int)P.L.off_cells) {
            c.h = (AgarEnvHeader*)(hot + P.L.off_header);
            c.pl = (AgarPlayer*)(hot + P.L.off_players);
        }
        if (P.hot_a >= (int)P.L.off_viruses) c.cells = (AgarCell*)(hot + P.L.off_cells);
        if (P.hot_a >= (int)P.L.off_blobs) c.vir = (AgarMote*)(hot + P.L.off_viruses);
    }
}

#ifdef AGAR_PHASE_CLOCKS
/* debug build only (tools/phase_clocks.py): cycles each env's tile spent per frame phase, summed over the launch */

#define CLK_MARK(slot)                                                      \
    do {                                                                    \
        if (active) clk_mark(c, slot);                                      \
    } while (0)
extern "C" int agar_debug_read_clocks(unsigned long long* host, int n_envs, int reset) {
    if (cudaMemcpyFromSymbol(host, g_clk, (size_t)n_envs * 16 * 8) != cudaSuccess) return -1;
    if (reset) {
        void* p = nullptr;
        cudaGetSymbolAddress(&p, g_clk);
        cudaMemset(p, 0, sizeof(unsigned long long) * 32768 * 16);
    }
    return 0;
}
#else
#define CLK_MARK(slot) do { } while (0)
#endif

/* ------------------------------------------------------------------ the step kernel */
template <int W, bool FULL, int MAXT>
/* minBlocks = 1 matters for the multi-agent kernels: with maxThreads alone ptxas budgets for two 512-thread CTAs per SM
 * (64 registers, 540 bytes of spills); one CTA per SM is what their shared memory allows anyway -> 112-116 registers, no
 * spills (measured +11 % config 3, +7 % config 4).  The single-cell general kernel runs many small CTAs per SM and keeps
 * the 64-register budget (measured -13 % at 65536 envs with 125 registers).  MAXT = 1024: 32 envs per CTA when their hot
 * sections fit (1-vs-greedy config): twice the warps walking the frame body in lock-step is worth more (+26 %) than the
 * registers (64, with spills). */
__global__ void __launch_bounds__(MAXT, (FULL || MAXT > 512) ? 1 : 2)
k_main(const __grid_constant__ DevParams P, uint8_t* __restrict__ state, const float* __restrict__ actions,
       float* __restrict__ obs, int n_frames, int n_dec, int flags, uint32_t dec_base) {
    const int tiles = blockDim.x / W;
    const int env0 = blockIdx.x * tiles;
    const int n_here = min(tiles, P.n_envs - env0);
    if (P.stage == 1) stage_in(P, state, env0, n_here, false);
    if (P.stage >= 2) pel_cache_copy(P, state, env0, n_here, tiles, true);
    const int tile_id = threadIdx.x / W;
    const bool active = tile_id < n_here;
    {
        Ctx<W> c(cg::tiled_partition<W>(cg::this_thread_block()));
        const int env = env0 + (active ? tile_id : 0);
        bind_ctx(c, P, active ? tile_id : 0, tiles, env, state);
        const int A = P.L.n_agents, K = P.L.n_players, SL = P.L.state_len;
        float* obs_env = obs ? obs + (size_t)env * A * SL : nullptr;
        const bool psync = P.phase_sync != 0;             /* barriers at frame start and before the field update */
        const bool psync_field = P.phase_sync == 2 || P.phase_sync >= 4; /* ... between the field phases */
        const bool psync_bots = P.phase_sync >= 3;        /* ... between bot turns */
        /* One loop with ONE call site per helper: the multi-agent frame is ~22 k SASS instructions and instruction
         * fetch is its top stall (profiles/r01_k_main_cfg3_*), so nothing big may be inlined twice.  Iteration `it` is
         * frame f of decision d; the trailing observation (KF_OBS_AFTER) is one extra iteration without a frame.
         * With phase_sync, CTA barriers between the phases keep all warps of the CTA in the same code region, so the
         * instruction cache serves them together (idle tiles of a partly filled CTA only hit the barriers). */
        const int total = n_dec * n_frames;
        for (int it = 0; it <= total; ++it) {
            const bool last = it == total;
            if (last && !(flags & KF_OBS_AFTER)) break;
#ifdef AGAR_PHASE_CLOCKS
            c.clk_t = clock64(), c.clk_env = env;
#endif
            if (psync) __syncthreads();
            CLK_MARK(0); /* (ptxas hoists this clock read above the barrier: the wait lands in the NEXT slot, the fov pass) */
            const int f = n_frames > 0 ? it % n_frames : 0, d = n_frames > 0 ? it / n_frames : 0;
            const bool emit = last || (f == 0 && (flags & KF_OBS_BEFORE)); /* observations leave the kernel here */
            if (active && !last && c.lane == 0) c.h->n_events = 0;
            if (FULL && K > 1 && active && !last) { /* every alive player's turn below would compute its own, on lane 0 */
                update_all_fovs(c, P);
                c.fov_done = true;
                CLK_MARK(8); /* fov pass */
                if (K * P.L.cell_cap > W) build_live_cells(c, P); /* one iteration covers all slots otherwise */
                CLK_MARK(11); /* live-cell list */
            }
            for (int k = 0; k < K; ++k) {
                if (psync_bots && k > 0) __syncthreads(); /* one bot turn per barrier interval */
                if (active) {
                    if (!FULL || P.cfg.bot_type[k] == AGAR_BOT_NN) {
                        nn_turn_begin<W, FULL>(c, P, k, (emit && obs_env) ? obs_env + (size_t)k * SL : nullptr);
                        if (last) continue;
                        if (c.lane == 0) {
                            float act[4] = {0.f, 0.f, 0.f, 0.f};
                            if (c.pl[k].bot.need_action) {
                                if (flags & KF_RANDOM_ACTIONS) { /* random-action driver */
                                    uint32_t w[4];
                                    philox(dec_base + (uint32_t)d, 7u, c.env_id, (uint32_t)k, (uint32_t)P.seed,
                                           (uint32_t)(P.seed >> 32), w);
                                    for (int i = 0; i < 4; ++i) act[i] = (float)(w[i] >> 8) * (1.0f / 16777216.0f);
                                } else {
                                    const float* ap = actions + ((size_t)env * A + k) * 4;
                                    for (int i = 0; i < 4; ++i) act[i] = ap[i];
                                }
                            }
                            nn_turn_end(c, P, k, act);
                        }
                        c.t.sync();
                    } else if (FULL && !last) {
                        CLK_MARK(1); /* NN turns so far */
                        scripted_turn(c, P, k);
                        CLK_MARK(7); /* scripted (greedy / random) turns */
                    }
                }
            }
            if (last) break;
            CLK_MARK(1); /* bot turns */
            c.fov_done = false, c.n_live = -1;
            for (int ph = 0; ph < AGAR_FIELD_PHASES; ++ph) {
                if (ph == 0 ? psync : psync_field) __syncthreads();
                if (ph == 0) CLK_MARK(2); /* wait at the field barrier */
                if (active) field_update_phase<W, FULL>(c, P, ph);
                CLK_MARK(3 + ph); /* 3: viruses / blobs / players  4: merge, virus overlaps  5: pellets  6: blobs, player-player, spawn */
            }
            if (active) {
                if (c.lane == 0) c.h->frame += 1;
                c.t.sync();
            }
            CLK_MARK(12); /* frame tail */
        }
        if (active && c.lane == 0 && (P.turn_reward || P.turn_done))
            for (int a = 0; a < A; ++a) {
                if (P.turn_reward) P.turn_reward[(size_t)env * A + a] = (float)c.pl[a].bot.last_reward;
                if (P.turn_done) P.turn_done[(size_t)env * A + a] = (uint8_t)c.pl[a].bot.exp_done;
            }
    }
    if (P.stage == 1) stage_out(P, state, env0, n_here);
    if (P.stage >= 2) {
        __syncthreads();
        pel_cache_copy(P, state, env0, n_here, tiles, false);
    }
    export_tail(P, obs, env0, n_here);
}


/* ------------------------------------------------------------------ register-resident kernel (agar_simple.cuh) */
struct SimplePlan {
    int frame_sync; /* CTA barrier at every frame: the warps of a CTA fetch the same instructions together */
    int strideB; /* bytes between the staged [pellets .. end of record] regions of consecutive envs */
    int grid_off; /* byte offset of the env's G x G observation accumulator inside its region; 0: accumulate in the caller's row */
};
/* one lane per env is throughput-bound: cap registers at 128 for 8 CTAs / SM (measured +10..50 % at >= 64k envs);
 * wider tiles are latency-bound at small env counts and prefer the uncapped allocation (measured) */
template <int W>
__global__ void __launch_bounds__(256, W == 1 ? 2 : 1)
k_simple(const __grid_constant__ DevParams P, const SimplePlan SP, uint8_t* __restrict__ state, const float* __restrict__ actions,
         float* __restrict__ obs, int n_frames, int n_dec, int flags, uint32_t dec_base) {
    const int per = blockDim.x / W; /* envs per CTA */
    const int env0 = blockIdx.x * per;
    const int n_here = min(per, P.n_envs - env0);
    const int rec_words = (int)(P.L.record_bytes / 4), pel_word = (int)(P.L.off_pellets / 4);
    const int tail_words = rec_words - pel_word;
    {   /* stage the pellet pool (+ event ring) in: coalesced 4-byte words; spare slots replicate the last env so that
         * every lane of a partly filled warp runs well-defined work */
        for (int i = threadIdx.x; i < per * tail_words; i += blockDim.x) {
            int e = i / tail_words, w = i - e * tail_words;
            int es = e < n_here ? e : n_here - 1;
            const uint32_t* src = (const uint32_t*)(state + (size_t)(env0 + es) * P.L.record_bytes) + pel_word;
            *(uint32_t*)(g_smem + (size_t)e * SP.strideB + w * 4) = src[w];
        }
    }
    __syncthreads();
    const int slot = threadIdx.x / W, sub = threadIdx.x % W;
    const bool valid = slot < n_here;
    const int env = env0 + (valid ? slot : n_here - 1);
    const uint32_t env_id = (uint32_t)(P.first_env + (uint64_t)env);
    uint8_t* rec = state + (size_t)env * P.L.record_bytes;
    uint8_t* b = g_smem + (size_t)slot * SP.strideB;
    SPtr q;
    q.h = (AgarEnvHeader*)(rec + P.L.off_header);
    q.p = (AgarPlayer*)(rec + P.L.off_players);
    q.c = (AgarCell*)(rec + P.L.off_cells);
    q.pel = (uint32_t*)b;
    q.ev = (AgarEvent*)(b + (P.L.off_events - P.L.off_pellets));
    q.grid = SP.grid_off ? (uint32_t*)(b + SP.grid_off) : nullptr;
    SReg r;
    s_load(r, q);
    float* row = (obs != nullptr && valid) ? obs + (size_t)env * P.L.state_len : nullptr;
    for (int d = 0; d < n_dec; ++d) {
        if (flags & KF_OBS_BEFORE) {
            int d_o = s_turn_begin(r, P);
            s_observe<W>(r, q, P, d_o != 0, row, sub);
        }
        for (int f = 0; f < n_frames; ++f) {
            if (SP.frame_sync == 1 || SP.frame_sync == 4) __syncthreads();
            if (SP.frame_sync == 2) __syncwarp();
            r.n_events = 0;
            int d_o = s_turn_begin(r, P);
            s_observe<W>(r, q, P, d_o != 0, nullptr, sub);
            float act[4] = {0.f, 0.f, 0.f, 0.f};
            if (r.need_action) {
                if (flags & KF_RANDOM_ACTIONS) {
                    uint32_t w[4];
                    philox(dec_base + (uint32_t)d, 7u, env_id, 0u, (uint32_t)P.seed, (uint32_t)(P.seed >> 32), w);
                    for (int i = 0; i < 4; ++i) act[i] = (float)(w[i] >> 8) * (1.0f / 16777216.0f);
                } else {
                    const float* ap = actions + (size_t)env * 4;
                    for (int i = 0; i < 4; ++i) act[i] = ap[i];
                }
            }
            double speed_pow;
            s_turn_end<W>(r, P, act, sub, speed_pow);
            if (SP.frame_sync == 4) __syncthreads();
            s_field_update<W>(r, q, P, env_id, sub, speed_pow);
        }
    }
    if (flags & KF_OBS_AFTER) {
        int d_o = s_turn_begin(r, P);
        s_observe<W>(r, q, P, d_o != 0, row, sub);
    }
    if (valid && sub == 0) {
        s_store(r, q);
        if (P.turn_reward) P.turn_reward[env] = (float)r.last_reward; /* Bot.getLastReward / done of this turn, fused */
        if (P.turn_done) P.turn_done[env] = (uint8_t)r.exp_done;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_here * tail_words; i += blockDim.x) { /* stage the pools out */
        int e = i / tail_words, w = i - e * tail_words;
        uint32_t* dst = (uint32_t*)(state + (size_t)(env0 + e) * P.L.record_bytes) + pel_word;
        dst[w] = *(const uint32_t*)(g_smem + (size_t)e * SP.strideB + w * 4);
    }
    export_tail(P, obs, env0, n_here);
}

/* mode 0: Model(...) + createBot*K + Model.initialize (model.py:51,154-162,90-94; field.py:57-67)
 * mode 1: Model.resetModel -> Field.reset (field.py:69-83)      mode 2: Model.resetBots (bot.py:125-164) */
template <int W, bool FULL>
__global__ void __launch_bounds__(1024)
k_init(const __grid_constant__ DevParams P, uint8_t* __restrict__ state, const uint8_t* __restrict__ mask, int mode) {
    const int tiles = blockDim.x / W;
    const int env0 = blockIdx.x * tiles;
    const int n_here = min(tiles, P.n_envs - env0);
    if (P.stage == 1)
        stage_in(P, state, env0, n_here, mode == 0);
    else if (mode == 0) { /* records stay in HBM: clear them in place */
        const int chunks = (int)(P.L.record_bytes / 8);
        uint2* dst = (uint2*)(state + (size_t)env0 * P.L.record_bytes);
        for (int i = threadIdx.x; i < n_here * chunks; i += blockDim.x) dst[i] = make_uint2(0u, 0u);
        __syncthreads();
    }
    if (P.stage >= 2) pel_cache_copy(P, state, env0, n_here, tiles, true);
    const int tile_id = threadIdx.x / W;
    if (tile_id < n_here && (mask == nullptr || mode == 0 || mask[env0 + tile_id])) {
        Ctx<W> c(cg::tiled_partition<W>(cg::this_thread_block()));
        bind_ctx(c, P, tile_id, tiles, env0 + tile_id, state);
        const int K = P.L.n_players;
        if (mode == 0 || mode == 1) {
            if (mode == 1) { /* clear pools cooperatively: fresh lists and hash tables */
                for (int i = c.lane; i < P.L.pellet_cap; i += W) c.pel[i] = 0;
                AgarFatPellet zf = {};
                for (int i = c.lane; i < P.L.fat_cap; i += W) c.fat[i] = zf;
                AgarMote zm = {};
                for (int i = c.lane; i < P.L.blob_cap; i += W) c.blob[i] = zm;
                for (int i = c.lane; i < P.L.virus_cap; i += W) c.vir[i] = zm;
                c.t.sync();
            }
            if (c.lane == 0) {
                if (mode == 0) {
                    for (int k = 0; k < K; ++k) {
                        AgarPlayer* p = &c.pl[k];
                        p->alive = 1;
                        p->cmd_x = p->cmd_y = -1;
                        p->bot.type = P.cfg.bot_type[k];
                    }
                } else {
                    c.h->n_events = 0;
                    c.h->n_pellets = c.h->n_fat = c.h->n_blobs = c.h->n_viruses = 0;
                    c.h->n_dead = 0;
                    for (int i = 0; i < AGAR_MAX_PLAYERS; ++i) c.h->dead_order[i] = 0;
                    for (int k = 0; k < K; ++k)
                        for (int i = 0; i < c.pl[k].n_cells; ++i) CELLP(c, P, k, i)->flags &= ~AGAR_CF_INHASH;
                    c.h->frame = 0;
                }
                c.h->event_hash = 0xCBF29CE484222325ULL;
                for (int k = 0; k < K; ++k) initialize_player(c, P, k);
            }
            c.t.sync();
            spawn_pellets(c, P);
            if (c.lane == 0) {
                if (P.cfg.virus_enabled) spawn_viruses(c, P);
                if (c.h->n_dead) spawn_players(c, P);
            }
            c.t.sync();
        }
        if (mode == 0 || mode == 2) {
            if (c.lane == 0)
                for (int k = 0; k < K; ++k) bot_reset(c, P, k);
            const int nh = P.L.n_agents * P.L.n_hist * P.L.grid_squares * P.L.grid_squares;
            for (int i = c.lane; i < nh; i += W) c.hist[i] = 0.f;
            c.t.sync();
        }
    }
    if (P.stage == 1) stage_out(P, state, env0, n_here);
    if (P.stage >= 2) {
        __syncthreads();
        pel_cache_copy(P, state, env0, n_here, tiles, false);
    }
}

/* agar_get: one thread per (env, agent) or per env, straight from HBM */
__global__ void k_get(const __grid_constant__ DevParams P, const uint8_t* __restrict__ state, int which, void* out) {
    const int A = P.L.n_agents;
    const int per_env = (which == AGAR_GET_OVERFLOW || which == AGAR_GET_EVENT_HASH);
    const int n = per_env ? P.n_envs : P.n_envs * A;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int env = per_env ? i : i / A, a = per_env ? 0 : i - env * A;
    const uint8_t* rec = state + (size_t)env * P.L.record_bytes;
    const AgarEnvHeader* h = (const AgarEnvHeader*)(rec + P.L.off_header);
    const AgarPlayer* p = (const AgarPlayer*)(rec + P.L.off_players) + a;
    const AgarCell* cells = (const AgarCell*)(rec + P.L.off_cells) + (size_t)a * P.L.cell_cap;
    switch (which) {
    case AGAR_GET_REWARD: ((float*)out)[i] = (float)p->bot.last_reward; break;
    case AGAR_GET_DONE: ((uint8_t*)out)[i] = (uint8_t)p->bot.exp_done; break;
    case AGAR_GET_VALID: ((uint8_t*)out)[i] = (uint8_t)p->bot.exp_valid; break;
    case AGAR_GET_NEED_ACTION: ((uint8_t*)out)[i] = (uint8_t)p->bot.need_action; break;
    case AGAR_GET_MASS: {
        int nc = p->n_cells;
        ((float*)out)[i] = nc ? (float)np_sum([&](int j) { return cells[j].mass; }, nc) : 0.f;
        break;
    }
    case AGAR_GET_FOV: ((float*)out)[i] = (float)p->fov_size; break;
    case AGAR_GET_NCELLS: ((int32_t*)out)[i] = p->n_cells; break;
    case AGAR_GET_ALIVE: ((uint8_t*)out)[i] = (uint8_t)p->alive; break;
    case AGAR_GET_STATS: {
        double* o = (double*)out + (size_t)i * 4;
        o[0] = p->bot.stat_mass_sum, o[1] = p->bot.stat_mass_max, o[2] = p->bot.stat_frames, o[3] = p->bot.stat_deaths;
        break;
    }
    case AGAR_GET_OVERFLOW: ((uint32_t*)out)[i] = h->overflow; break;
    case AGAR_GET_EVENT_HASH: ((uint64_t*)out)[i] = h->event_hash; break;
    }
}

/* ------------------------------------------------------------------ host side */
struct AgarEnv {
    AgarConfig cfg;
    AgarLayout L;
    DevParams P;
    uint8_t* state;
    double* deg_tab;
    int n_envs, device, W, full, threads, tiles;
    size_t smem_bytes;
    int init_W, init_threads, init_tiles;
    size_t init_smem;
    SimplePlan sp;
    int simple_W, simple_threads;
    size_t simple_smem;
    int64_t launches;
    char err[256];
    /* step_host staging */
    float *d_actions, *d_obs, *d_reward;
    uint8_t* d_done;
    void* h_turn; /* pinned: float reward[EA] | uint8 done[EA], then (64-byte aligned) the completion flag word */
    volatile uint32_t* h_flag; /* inside h_turn's allocation; the last CTA of a zero-copy step launch writes flag_seq there */
    unsigned int* d_export_count;
    uint32_t flag_seq;
    int host_polling; /* the pending step raises h_flag (zero-copy path) */
    int host_pending; /* agar_step_host_begin issued, _end not yet */
    int host_zerocopy; /* pinned caller buffers are read / written in place (AGAR_HOST_ZEROCOPY=0 disables) */
};
static char g_create_err[256] = "";

static int fail(AgarEnv* e, int code, const char* fmt, const char* detail) {
    char* dst = e ? e->err : g_create_err;
    snprintf(dst, 256, fmt, detail);
    return code;
}
#define CU(call)                                                                 \
    do {                                                                         \
        cudaError_t _e = (call);                                                 \
        if (_e != cudaSuccess) return fail(env, AGAR_E_CUDA, "CUDA error: %s", cudaGetErrorString(_e)); \
    } while (0)

extern "C" int agar_layout_for_config(const AgarConfig* cfg, AgarLayout* out) {
    if (!cfg || !out) return AGAR_E_INVALID;
    return agar_layout_compute(cfg, out);
}

static bool config_is_simple(const AgarConfig& c, const AgarLayout& L) {
    return c.n_players == 1 && c.bot_type[0] == AGAR_BOT_NN && !c.virus_enabled && !c.enable_split && !c.enable_eject &&
           L.cell_cap == 1 && !c.self_grid && !c.wall_grid && !c.enemy_grid && !c.virus_grid && !c.self_grid_lf &&
           !c.self_grid_slf && !c.enemy_grid_lf && !c.enemy_grid_slf && !c.use_last_action && !c.use_second_last_action &&
           !c.use_last_fovsize && L.n_hist == 0 && !c.all_player_grid && !c.simple_state && /* the 12-value representation lives in the general kernel */
           L.field_size < 128; /* the float32 candidate filter of k_simple's eat chain is analysed for coordinates below 128 (one player: 75) */
}

/* pick the launch shape for tile width W; returns false if one record does not fit in shared memory */
static bool plan_launch(AgarEnv* e, int W) {
    const size_t budget = 200 * 1024;
    size_t per_tile = (e->P.stage == 1 ? (size_t)e->P.rec_stride : 0) + (size_t)e->P.scratch_bytes +
                      (e->P.stage >= 2 ? (size_t)((e->P.hot_a + 15) / 16 * 16) + (size_t)((e->L.pellet_cap * 4 + 15) / 16 * 16) : 0);
    const int max_threads = (e->full && W == 32) ? 1024 : 512; /* k_main<32, true, 1024> exists for 32-lane tiles only */
    int tiles = e->full ? max_threads / W : 128 / W; /* as many envs per CTA as fit: the barriers then align more warps */
    if (getenv("AGAR_MAX_TILES")) tiles = atoi(getenv("AGAR_MAX_TILES"));
    if (tiles * W > max_threads) tiles = max_threads / W;
    while (tiles > 1 && per_tile * tiles > budget) tiles -= 1;
    if (W == 32 && tiles > 4) tiles -= tiles % 4; /* warps spread evenly over the four schedulers of an SM */
    if (per_tile * tiles > budget) return false;
    if (e->full && W == 32 && tiles > 8 && !getenv("AGAR_MAX_TILES")) {
        /* one CTA per SM and all CTAs take about as long: fewer envs per CTA can fill the last wave better.  Cost model
         * waves x envs-per-CTA (measured, config 3 at 16384 envs: 32 -> 1.01e8, 28 -> 1.05e8, 24 -> 0.97e8, 20 -> 0.89e8) */
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, e->device);
        long long best_cost = -1;
        int best = tiles;
        for (int t = tiles; t >= 8 && t >= tiles - 4; t -= 4) { /* one step: smaller CTAs lose lock-step efficiency */
            long long ctas = (e->n_envs + t - 1) / t, waves = (ctas + sms - 1) / sms;
            long long cost = waves * t;
            if (best_cost < 0 || cost < best_cost) best_cost = cost, best = t;
        }
        tiles = best;
    }
    e->W = W;
    e->tiles = tiles;
    e->threads = tiles * W;
    e->smem_bytes = per_tile * tiles;
    return true;
}

/* cudaFuncSetAttribute is per kernel FUNCTION and process-wide, not per handle: two handles with different shared-memory
 * needs (a training batch and an evaluation batch of the same config) must not lower each other's limit.  Every
 * instantiation is therefore raised ONCE per device to the device's opt-in maximum (the launch itself still requests only
 * what it needs, so occupancy is unaffected).  `done` is a per-instantiation static; setting twice is harmless. */
template <typename K>
static cudaError_t ensure_max_smem(K kernel, int device, unsigned char* done) {
    if (device >= 0 && device < 64 && done[device]) return cudaSuccess;
    int optin = 0;
    cudaError_t err = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    if (err != cudaSuccess) return err;
    err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
    if (err == cudaSuccess && device >= 0 && device < 64) done[device] = 1;
    return err;
}

template <int W, bool FULL, int MAXT>
static cudaError_t launch_main_k(AgarEnv* e, const float* actions, float* obs, int n_frames, int n_dec, int flags,
                                 uint32_t dec_base, cudaStream_t s) {
    static unsigned char done[64];
    cudaError_t err = ensure_max_smem(k_main<W, FULL, MAXT>, e->device, done);
    if (err != cudaSuccess) return err;
    int blocks = (e->n_envs + e->tiles - 1) / e->tiles;
    k_main<W, FULL, MAXT><<<blocks, e->threads, e->smem_bytes, s>>>(e->P, e->state, actions, obs, n_frames, n_dec, flags, dec_base);
    return cudaGetLastError();
}
template <int W, bool FULL>
static cudaError_t launch_main_t(AgarEnv* e, const float* actions, float* obs, int n_frames, int n_dec, int flags,
                                 uint32_t dec_base, cudaStream_t s) {
    if (W == 32 && FULL && e->threads > 768) return launch_main_k<32, true, 1024>(e, actions, obs, n_frames, n_dec, flags, dec_base, s);
    if (W == 32 && FULL && e->threads > 512) return launch_main_k<32, true, 768>(e, actions, obs, n_frames, n_dec, flags, dec_base, s);
    return launch_main_k<W, FULL, 512>(e, actions, obs, n_frames, n_dec, flags, dec_base, s);
}
template <int W, bool FULL>
static cudaError_t launch_init_t(AgarEnv* e, const uint8_t* mask, int mode, cudaStream_t s) {
    static unsigned char done[64];
    cudaError_t err = ensure_max_smem(k_init<W, FULL>, e->device, done);
    if (err != cudaSuccess) return err;
    int blocks = (e->n_envs + e->init_tiles - 1) / e->init_tiles;
    k_init<W, FULL><<<blocks, e->init_threads, e->init_smem, s>>>(e->P, e->state, mask, mode);
    return cudaGetLastError();
}
#define DISPATCH(fn, ...)                                                 \
    (e->full ? (e->W == 32  ? fn<32, true>(__VA_ARGS__)                   \
                : e->W == 16 ? fn<16, true>(__VA_ARGS__)                  \
                : e->W == 8  ? fn<8, true>(__VA_ARGS__)                   \
                             : fn<4, true>(__VA_ARGS__))                  \
             : (e->W == 32  ? fn<32, false>(__VA_ARGS__)                  \
                             : fn<32, false>(__VA_ARGS__)))

static int launch_main(AgarEnv* e, const float* actions, float* obs, int n_frames, int n_dec, int flags, uint32_t dec_base,
                       void* stream) {
    AgarEnv* env = e;
    CU(cudaSetDevice(e->device));
    cudaError_t err;
    if (e->simple_W) {
        int blocks = (e->n_envs * e->simple_W + e->simple_threads - 1) / e->simple_threads;
#define LAUNCH_SIMPLE(WW)                                                                                                     \
    do {                                                                                                                      \
        static unsigned char done_##WW[64];                                                                                   \
        err = ensure_max_smem(k_simple<WW>, e->device, done_##WW);                                                            \
        if (err == cudaSuccess) {                                                                                             \
            k_simple<WW><<<blocks, e->simple_threads, e->simple_smem, (cudaStream_t)stream>>>(e->P, e->sp, e->state, actions, obs, \
                                                                                              n_frames, n_dec, flags, dec_base);  \
            err = cudaGetLastError();                                                                                         \
        }                                                                                                                     \
    } while (0)
        if (e->simple_W == 1) LAUNCH_SIMPLE(1);
        else if (e->simple_W == 2) LAUNCH_SIMPLE(2);
        else if (e->simple_W == 4) LAUNCH_SIMPLE(4);
        else if (e->simple_W == 8) LAUNCH_SIMPLE(8);
        else if (e->simple_W == 16) LAUNCH_SIMPLE(16);
        else LAUNCH_SIMPLE(32);
#undef LAUNCH_SIMPLE
    } else
        err = DISPATCH(launch_main_t, e, actions, obs, n_frames, n_dec, flags, dec_base, (cudaStream_t)stream);
    if (err != cudaSuccess) return fail(e, AGAR_E_CUDA, "kernel launch failed: %s", cudaGetErrorString(err));
    e->launches += 1;
    return AGAR_OK;
}
static int launch_init(AgarEnv* e, const uint8_t* mask, int mode, void* stream) {
    AgarEnv* env = e;
    CU(cudaSetDevice(e->device));
    cudaError_t err = e->full ? (e->init_W == 32 ? launch_init_t<32, true>(e, mask, mode, (cudaStream_t)stream)
                                                  : launch_init_t<8, true>(e, mask, mode, (cudaStream_t)stream))
                              : launch_init_t<32, false>(e, mask, mode, (cudaStream_t)stream);
    if (err != cudaSuccess) return fail(e, AGAR_E_CUDA, "kernel launch failed: %s", cudaGetErrorString(err));
    e->launches += 1;
    return AGAR_OK;
}

/* Lanes per env.  Single-cell pellet-collection configs: 1, 2, 4, 8 select the register-resident kernel
 * (agar_simple.cuh), 16 / 32 the general kernel; every other config: 4, 8, 16, 32 (general kernel). */
extern "C" int agar_set_tile_width(AgarEnv* e, int W) {
    if (!e) return AGAR_E_INVALID;
    /* the register-resident kernel bins a pellet into at most two squares per axis: 2 * radius < fov / G needs G <= 16 */
    if (!e->full && e->L.grid_squares <= 16 && (W == 1 || W == 2 || W == 4 || W == 8 || W == 16 || W == 32) &&
        !getenv("AGAR_GENERAL_KERNEL")) {
        int tail_words = (int)((e->L.record_bytes - e->L.off_pellets) / 4);
        e->sp.strideB = (tail_words | 1) * 4;
        e->sp.grid_off = 0;
        /* the observation's G x G sums are accumulated in shared memory and leave once, as consecutive floats of the row
         * (measured: 4096 envs W=8 +1.2 %, 65536 envs W=2 +8 %).  Not at one lane per env: 484 more bytes per env halve the
         * resident envs per SM there (1M envs: 3.28e9 -> 1.52e9), and a lone lane's stores are strided either way. */
        const char* genv = getenv("AGAR_SIMPLE_OBS_SMEM");
        if (genv ? atoi(genv) != 0 : W >= 2) {
            const int gg = e->L.grid_squares * e->L.grid_squares;
            e->sp.grid_off = tail_words * 4;
            e->sp.strideB = ((tail_words + gg) | 1) * 4;
        }
        if (e->L.pellet_cap > (W >= 4 ? 64 : 128) * W) /* two mask words per lane: SMask<W>, agar_simple.cuh */
            return fail(e, AGAR_E_UNSUPPORTED, "pellet pool too large for this tile width%s", "");
        int threads = 128; /* measured (round 2, exact libm arithmetic): four warps per CTA beat two at every batch size (4096 envs: 7.7e8 vs 7.0e8) */
        const char* tenv = getenv("AGAR_SIMPLE_THREADS");
        if (tenv && atoi(tenv) >= 32) threads = atoi(tenv) / 32 * 32;
        if (threads > 256) threads = 256; /* __launch_bounds__(256, ...) */
        /* measured: one CTA barrier per frame is worth 1.9x at 1M envs and 2.3x at 4096 (a warp-level barrier is
         * worth nothing): the warps of a CTA then walk the ~5000-instruction frame body together */
        e->sp.frame_sync = getenv("AGAR_SIMPLE_SYNC") ? atoi(getenv("AGAR_SIMPLE_SYNC")) : 1;
        size_t per_env = (size_t)e->sp.strideB;
        while (threads > 32 && per_env * (threads / W) > 200 * 1024) threads -= 32;
        if (per_env * (threads / W) > 200 * 1024)
            return fail(e, AGAR_E_NOMEM, "pellet pool + event ring of %s do not fit in shared memory at this tile width", "the envs of one CTA");
        e->simple_W = W;
        e->simple_threads = threads;
        e->simple_smem = per_env * (threads / W);
        e->W = W;
        return AGAR_OK;
    }
    bool ok = W == 32 || (e->full && (W == 4 || W == 8 || W == 16));
    if (!ok) return fail(e, AGAR_E_INVALID, "unsupported tile width%s", "");
    if (!plan_launch(e, W)) return fail(e, AGAR_E_NOMEM, "env record does not fit in shared memory%s", "");
    e->simple_W = 0;
    return AGAR_OK;
}
extern "C" int agar_get_tile_width(const AgarEnv* e) { return e ? e->W : 0; }

extern "C" int agar_create(const AgarConfig* cfg, int n_envs, int device, uint64_t seed, uint64_t first_env_id, void* stream,
                           AgarEnv** out) {
    AgarEnv* env = nullptr;
    if (!cfg || !out || n_envs < 1) return fail(nullptr, AGAR_E_INVALID, "bad arguments to agar_create%s", "");
    AgarLayout L;
    int rc = agar_layout_compute(cfg, &L);
    if (rc != AGAR_OK) return fail(nullptr, rc, "config rejected by agar_layout_compute%s", "");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
        return fail(nullptr, AGAR_E_CUDA, "no such CUDA device%s", "");
    CU(cudaSetDevice(device));
    AgarEnv* e = (AgarEnv*)calloc(1, sizeof(AgarEnv));
    env = e;
    e->cfg = *cfg;
    e->L = L;
    e->n_envs = n_envs;
    e->device = device;
    e->full = config_is_simple(*cfg, L) ? 0 : 1;
    DevParams& P = e->P;
    memset(&P, 0, sizeof P);
    P.cfg = *cfg;
    P.L = L;
    P.S = L.field_size;
    P.nb = (int)ceil((double)P.S / AG_BUCKET);
    P.n_envs = n_envs;
    P.rec_stride = (int)L.record_bytes + 8;
    P.full = e->full;
    static_assert(AGAR_MAX_CELLS == 16, "cap_shift assumes cell_cap in {1, 16} (agar_layout.h)");
    P.cap_shift = L.cell_cap == AGAR_MAX_CELLS ? 4 : 0;
    P.g_magic = L.grid_squares > 1 ? (uint32_t)((1ull << 32) / (unsigned)L.grid_squares) + 1u : 0u;
    int vel_bytes = e->full ? L.n_players * L.cell_cap * 2 * 8 : 0;
    int obs_bytes = obs_scratch_bytes(L.grid_squares, e->full != 0);
    /* the pellet index (agar_dev.cuh) lives in the same scratch: it is built after the players have moved (velocities dead)
     * and before the next bot phase (observation tables dead); pools of <= 8 chunks are scanned directly */
    P.pel_index = e->full && L.pellet_cap > (getenv("AGAR_PEL_INDEX_MIN") ? atoi(getenv("AGAR_PEL_INDEX_MIN")) : 256) && !(getenv("AGAR_PEL_INDEX") && atoi(getenv("AGAR_PEL_INDEX")) == 0);
    const int gb_idx = (P.S + 9) / 10; /* AG_IDX_CELL */
    int idx_bytes = P.pel_index ? ((gb_idx * gb_idx + 3) & ~1) * 2 + L.pellet_cap * 2 + 128 + 64 + 16 : 0; /* counters, entries, sort list, ex-blob list */
    if (idx_bytes > vel_bytes) vel_bytes = idx_bytes;
    P.scratch_bytes = ((vel_bytes > obs_bytes ? vel_bytes : obs_bytes) + 15) / 16 * 16 + 8;
    P.live_off = P.scratch_bytes; /* live-cell list: uint16 per cell slot, after the observation / velocity scratch */
    if (e->full) P.scratch_bytes += (L.n_players * L.cell_cap * 2 + 15) / 16 * 16;
    /* Where a record lives during a launch.  1: whole record staged in shared memory (single-cell general kernel).
     * Multi-agent configs keep the record in HBM / L2 and cache on chip what every scan and lane-0 step reads:
     * 3 = header + players + cells + viruses + pellet pool if at least 12 envs per CTA still fit, else 2 = the pellet
     * pool only (the 16-player arena: 23 KB of cells).  Measured when the policy was chosen (16 envs per CTA): config 3
     * 6.3e7 / 7.3e7 / 7.7e7 env-steps/s for 1 / 2 / 3; config 4 0.75e6 (1, 3 warps per SM) / 3.3e6 (2) / 1.6e6 (3);
     * with 32 envs per CTA (profiles/r01_sweep.txt) config 3 does 9.1e7 (2) / 9.6e7 (3). */
    if (!e->full)
        P.stage = 1;
    else {
        size_t hot3 = (size_t)((L.off_blobs + 15) / 16 * 16) + (size_t)((L.pellet_cap * 4 + 15) / 16 * 16) + P.scratch_bytes;
        P.stage = hot3 * 12 <= 200 * 1024 ? 3 : 2;
    }
    if (getenv("AGAR_STAGE")) P.stage = atoi(getenv("AGAR_STAGE"));
    P.hot_a = P.stage == 3 ? (int)L.off_blobs : (P.stage == 4 ? (int)L.off_cells : 0);
    /* multi-agent kernels are instruction-fetch bound: lock-step the CTA's warps phase by phase (measured 2x) */
    P.phase_sync = getenv("AGAR_PHASE_SYNC") ? atoi(getenv("AGAR_PHASE_SYNC")) : e->full;
    double speed_modifier = 1.0 / 30;
    P.move_speed = 90 * speed_modifier;
    P.decay_rate = 1 - (0.01 * speed_modifier);
    P.blob_mass = 18 * 0.8;
    P.virus_split_mass = 100.0 + 7 * 18 * 0.8;
    P.start_radius = sqrt(10.0 / M_PI);
    P.virus_radius = sqrt(100.0 / M_PI);
    for (int m = 0; m < 4; ++m) P.pellet_r[m] = m ? sqrt((double)m / M_PI) : 0.0;
    for (int n = 1; n <= 16; ++n) P.pow_n[n] = pow((double)n, 0.32);
    P.seed = seed;
    P.first_env = first_env_id;
    double tab[720];
    for (int d = 0; d < 360; ++d) {
        double a = d * (M_PI / 180.0);
        tab[d] = cos(a), tab[360 + d] = sin(a);
    }
    if (cudaMalloc(&e->deg_tab, sizeof tab) != cudaSuccess || cudaMalloc(&e->state, (size_t)n_envs * L.record_bytes) != cudaSuccess) {
        cudaGetLastError();
        if (e->deg_tab) cudaFree(e->deg_tab);
        free(e);
        return fail(nullptr, AGAR_E_NOMEM, "cudaMalloc of the env state failed%s", "");
    }
    CU(cudaMemcpyAsync(e->deg_tab, tab, sizeof tab, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    CU(cudaStreamSynchronize((cudaStream_t)stream)); /* tab is a stack buffer */
    P.deg_tab = e->deg_tab;
    {   /* k_init always runs with 32-lane tiles (8 if a record is too big for four tiles... it is not hot) */
        if (!plan_launch(e, 32)) {
            cudaFree(e->state), cudaFree(e->deg_tab), free(e);
            return fail(nullptr, AGAR_E_NOMEM, "env record does not fit in shared memory%s", "");
        }
        e->init_W = 32, e->init_threads = e->threads, e->init_tiles = e->tiles, e->init_smem = e->smem_bytes;
    }
    int W = e->full ? 32 : (n_envs <= 8192 ? 8 : (n_envs <= 32768 ? 2 : 1)); /* tools/sweep.py, profiles/r01_sweep.txt */
    const char* wenv = getenv("AGAR_TILE_W");
    if (wenv && atoi(wenv) > 0) W = atoi(wenv);
    if (agar_set_tile_width(e, W) != AGAR_OK && agar_set_tile_width(e, 8) != AGAR_OK && agar_set_tile_width(e, 32) != AGAR_OK) {
        snprintf(g_create_err, sizeof g_create_err, "%s", e->err);
        cudaFree(e->state), cudaFree(e->deg_tab), free(e);
        return AGAR_E_NOMEM;
    }
    rc = launch_init(e, nullptr, 0, stream);
    if (rc != AGAR_OK) {
        snprintf(g_create_err, sizeof g_create_err, "%s", e->err);
        cudaFree(e->state), cudaFree(e->deg_tab), free(e);
        return rc;
    }
    *out = e;
    return AGAR_OK;
}
extern "C" int agar_destroy(AgarEnv* e) {
    if (!e) return AGAR_E_INVALID;
    cudaSetDevice(e->device);
    cudaFree(e->state);
    cudaFree(e->deg_tab);
    if (e->d_actions) cudaFree(e->d_actions);
    if (e->d_obs) cudaFree(e->d_obs);
    if (e->d_reward) cudaFree(e->d_reward);
    if (e->h_turn) cudaFreeHost(e->h_turn);
    if (e->d_export_count) cudaFree(e->d_export_count);
    free(e);
    return AGAR_OK;
}
extern "C" const char* agar_last_error(const AgarEnv* e) { return e ? e->err : g_create_err; }
extern "C" int agar_get_layout(const AgarEnv* e, AgarLayout* out) {
    if (!e || !out) return AGAR_E_INVALID;
    *out = e->L;
    return AGAR_OK;
}
extern "C" int agar_num_envs(const AgarEnv* e) { return e ? e->n_envs : AGAR_E_INVALID; }
extern "C" int64_t agar_launch_count(const AgarEnv* e) { return e ? e->launches : 0; }
extern "C" void* agar_state_ptr(const AgarEnv* e) { return e ? e->state : nullptr; }

extern "C" int agar_reset(AgarEnv* e, const uint8_t* env_mask_dev, void* stream) {
    if (!e) return AGAR_E_INVALID;
    return launch_init(e, env_mask_dev, 1, stream);
}
extern "C" int agar_reset_bots(AgarEnv* e, const uint8_t* env_mask_dev, void* stream) {
    if (!e) return AGAR_E_INVALID;
    return launch_init(e, env_mask_dev, 2, stream);
}
extern "C" int agar_observe(AgarEnv* e, float* obs_dev, void* stream) {
    if (!e) return AGAR_E_INVALID;
    return launch_main(e, nullptr, obs_dev, 0, 0, KF_OBS_AFTER, 0, stream);
}
extern "C" int agar_step(AgarEnv* e, const float* actions_dev, int n_frames, void* stream) {
    if (!e || n_frames < 0) return AGAR_E_INVALID;
    if (!actions_dev && e->L.n_agents > 0) return fail(e, AGAR_E_INVALID, "actions_dev is NULL%s", "");
    return launch_main(e, actions_dev, nullptr, n_frames, 1, 0, 0, stream);
}
extern "C" int agar_step_observe(AgarEnv* e, const float* actions_dev, int n_frames, float* obs_dev, void* stream) {
    if (!e || n_frames < 0) return AGAR_E_INVALID;
    if (!actions_dev && e->L.n_agents > 0) return fail(e, AGAR_E_INVALID, "actions_dev is NULL%s", "");
    return launch_main(e, actions_dev, obs_dev, n_frames, 1, KF_OBS_AFTER, 0, stream);
}
/* n_decisions x (observe -> uniform random action from Philox stream 7 -> n_frames frames), one launch.
 * The random-action driver of BASELINE config 2; obs_dev (nullable) receives every decision's observation. */
extern "C" int agar_rollout_random(AgarEnv* e, int n_decisions, int n_frames, uint32_t decision_base, float* obs_dev,
                                   void* stream) {
    if (!e || n_decisions < 0 || n_frames < 0) return AGAR_E_INVALID;
    return launch_main(e, nullptr, obs_dev, n_frames, n_decisions, KF_OBS_BEFORE | KF_RANDOM_ACTIONS, decision_base, stream);
}
extern "C" int agar_get(AgarEnv* e, AgarField which, void* out_dev, void* stream) {
    AgarEnv* env = e;
    if (!e || !out_dev || (int)which < 0 || (int)which > AGAR_GET_EVENT_HASH) return AGAR_E_INVALID;
    CU(cudaSetDevice(e->device));
    int per_env = (which == AGAR_GET_OVERFLOW || which == AGAR_GET_EVENT_HASH);
    int n = per_env ? e->n_envs : e->n_envs * e->L.n_agents;
    if (n == 0) return AGAR_OK;
    k_get<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(e->P, e->state, (int)which, out_dev);
    CU(cudaGetLastError());
    e->launches += 1;
    return AGAR_OK;
}
extern "C" int agar_debug_dump(AgarEnv* e, int env_index, void* record_host, size_t bytes, void* stream) {
    AgarEnv* env = e;
    if (!e || !record_host || env_index < 0 || env_index >= e->n_envs || bytes != e->L.record_bytes) return AGAR_E_INVALID;
    CU(cudaSetDevice(e->device));
    CU(cudaMemcpyAsync(record_host, e->state + (size_t)env_index * e->L.record_bytes, bytes, cudaMemcpyDeviceToHost,
                       (cudaStream_t)stream));
    CU(cudaStreamSynchronize((cudaStream_t)stream));
    return AGAR_OK;
}
extern "C" int agar_debug_load(AgarEnv* e, int env_index, const void* record_host, size_t bytes, void* stream) {
    AgarEnv* env = e;
    if (!e || !record_host || env_index < 0 || env_index >= e->n_envs || bytes != e->L.record_bytes) return AGAR_E_INVALID;
    CU(cudaSetDevice(e->device));
    CU(cudaMemcpyAsync(e->state + (size_t)env_index * e->L.record_bytes, record_host, bytes, cudaMemcpyHostToDevice,
                       (cudaStream_t)stream));
    CU(cudaStreamSynchronize((cudaStream_t)stream));
    return AGAR_OK;
}
/* device-visible address of a pinned host buffer, or nullptr for pageable memory */
static void* pinned_dev_ptr(const void* p) {
    if (!p) return nullptr;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return at.type == cudaMemoryTypeHost ? at.devicePointer : nullptr;
}

extern "C" int agar_step_host_begin(AgarEnv* e, const float* actions_host, int n_frames, float* obs_host, void* stream) {
    AgarEnv* env = e;
    if (!e || !actions_host || n_frames < 0) return AGAR_E_INVALID;
    CU(cudaSetDevice(e->device));
    cudaStream_t s = (cudaStream_t)stream;
    const size_t EA = (size_t)e->n_envs * (e->L.n_agents ? e->L.n_agents : 1);
    const size_t turn_words = (EA * 5 + 3) / 4;
    if (!e->d_actions) {
        CU(cudaMalloc(&e->d_actions, EA * 4 * sizeof(float)));
        CU(cudaMalloc(&e->d_obs, EA * e->L.state_len * sizeof(float)));
        CU(cudaMalloc(&e->d_reward, turn_words * 4)); /* packed: float reward[EA] | uint8 done[EA] -> one copy back */
        const size_t flag_off = (turn_words * 4 + 63) / 64 * 64;
        CU(cudaMallocHost(&e->h_turn, flag_off + 64));
        e->h_flag = (volatile uint32_t*)((uint8_t*)e->h_turn + flag_off);
        *e->h_flag = 0;
        e->flag_seq = 0;
        CU(cudaMalloc(&e->d_export_count, sizeof(unsigned int)));
        CU(cudaMemsetAsync(e->d_export_count, 0, sizeof(unsigned int), s));
        e->d_done = (uint8_t*)e->d_reward + EA * 4;
        CU(cudaMemsetAsync(e->d_obs, 0, EA * e->L.state_len * sizeof(float), s));
        CU(cudaMemsetAsync(e->d_reward, 0, turn_words * 4, s));
        const char* zc = getenv("AGAR_HOST_ZEROCOPY");
        e->host_zerocopy = zc ? atoi(zc) : 1;
    }
    /* Pinned caller buffers (cudaHostAlloc / cudaHostRegister / torch pin_memory) are visible to the device: the step
     * kernel reads the actions in place and its CTAs store the results straight into them (export_tail).  Pageable buffers take the
     * copy-engine path. */
    const float* act_dev = e->host_zerocopy ? (const float*)pinned_dev_ptr(actions_host) : nullptr;
    float* obs_dev_host = (e->host_zerocopy && obs_host) ? (float*)pinned_dev_ptr(obs_host) : nullptr;
    const bool zero_copy = act_dev && (!obs_host || (obs_dev_host && ((uintptr_t)obs_dev_host & 15) == 0));
    if (!zero_copy) CU(cudaMemcpyAsync(e->d_actions, actions_host, EA * 4 * sizeof(float), cudaMemcpyHostToDevice, s));
    e->P.turn_reward = e->d_reward, e->P.turn_done = e->d_done; /* reward / done written by the step kernel itself */
    e->host_polling = 0;
    if (zero_copy) { /* ONE launch: each CTA exports its own rows over PCIe and the last one raises the flag (export_tail) */
        void* turn_dev_host = nullptr;
        void* flag_dev_host = nullptr;
        CU(cudaHostGetDevicePointer(&turn_dev_host, e->h_turn, 0));
        CU(cudaHostGetDevicePointer(&flag_dev_host, (void*)e->h_flag, 0));
        e->flag_seq += 1;
        if (e->flag_seq == 0) e->flag_seq = 1;
        e->P.host_obs = obs_dev_host, e->P.host_turn = (uint32_t*)turn_dev_host, e->P.export_count = e->d_export_count;
        e->P.host_flag = (volatile uint32_t*)flag_dev_host, e->P.flag_value = e->flag_seq;
        e->host_polling = 1;
    }
    int rc = launch_main(e, zero_copy ? act_dev : e->d_actions, e->d_obs, n_frames, 1, KF_OBS_AFTER, 0, s);
    e->P.turn_reward = nullptr, e->P.turn_done = nullptr;
    e->P.host_obs = nullptr, e->P.host_turn = nullptr, e->P.export_count = nullptr, e->P.host_flag = nullptr;
    if (rc != AGAR_OK) {
        e->host_polling = 0;
        return rc;
    }
    if (!zero_copy) {
        const size_t obs_bytes = EA * e->L.state_len * sizeof(float);
        if (obs_host) CU(cudaMemcpyAsync(obs_host, e->d_obs, obs_bytes, cudaMemcpyDeviceToHost, s));
        CU(cudaMemcpyAsync(e->h_turn, e->d_reward, EA * 5, cudaMemcpyDeviceToHost, s));
    }
    e->host_pending = 1;
    return AGAR_OK;
}

extern "C" int agar_step_host_end(AgarEnv* e, float* reward_host, uint8_t* done_host, void* stream) {
    AgarEnv* env = e;
    if (!e) return AGAR_E_INVALID;
    if (!e->host_pending) return fail(env, AGAR_E_INVALID, "%s", "agar_step_host_end without agar_step_host_begin");
    CU(cudaSetDevice(e->device));
    if (e->host_polling) {
        /* the launch's last CTA writes flag_seq into pinned memory after every row has been made visible system-wide: poll it
         * (a few hundred ns of latency instead of a driver wake-up); cudaStreamQuery now and then catches a failed launch */
        const uint32_t want = e->flag_seq;
        unsigned spins = 0;
        while (*e->h_flag != want) {
            if ((++spins & 0xfff) == 0) {
                cudaError_t q = cudaStreamQuery((cudaStream_t)stream);
                if (q == cudaSuccess) break;
                if (q != cudaErrorNotReady) return fail(env, AGAR_E_CUDA, "CUDA error: %s", cudaGetErrorString(q));
            }
#if defined(__x86_64__) || defined(__i386__)
            __builtin_ia32_pause();
#endif
        }
        if (*e->h_flag != want) CU(cudaStreamSynchronize((cudaStream_t)stream)); /* the stream drained without the flag: surface the error */
        e->host_polling = 0;
    } else {
        CU(cudaStreamSynchronize((cudaStream_t)stream));
    }
    e->host_pending = 0;
    const size_t EA = (size_t)e->n_envs * (e->L.n_agents ? e->L.n_agents : 1);
    if (reward_host) memcpy(reward_host, e->h_turn, EA * 4);
    if (done_host) memcpy(done_host, (uint8_t*)e->h_turn + EA * 4, EA);
    return AGAR_OK;
}

extern "C" int agar_step_host(AgarEnv* e, const float* actions_host, int n_frames, float* obs_host, float* reward_host,
                              uint8_t* done_host, void* stream) {
    int rc = agar_step_host_begin(e, actions_host, n_frames, obs_host, stream);
    return rc != AGAR_OK ? rc : agar_step_host_end(e, reward_host, done_host, stream);
}
