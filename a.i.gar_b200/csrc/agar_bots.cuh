/*
 * agar_bots.cuh — the agent half of the frame: grid-vision observation, reward / done / frame-skip
 * bookkeeping, action -> command point, and the scripted (greedy / random) opponents.
 * Replaces src/model/bot.py (file:line cited per function).  Same execution model as agar_dev.cuh.
 */
#pragma once
#include "agar_dev.cuh"

/* scratch layout used while observing (aliases the velocity scratch of update_players).  The multi-agent encoder
 * builds its channels ONE AT A TIME through a single 8-byte table, so an env needs (G+1)^2 * 12 bytes of table scratch
 * instead of (G+1)^2 * 44: 32 envs of the 1-vs-greedy config then share a CTA (measured +26 %). */
struct ObsScratch {
    int* pel_i;       /* [(G+1)^2] integer pellet mass per FOV bucket; later: 1 + index of the bucket's virus */
    double* mid;      /* [2][G]    gsMidPoint sequences (x then y)                 */
    double* tab;      /* FULL: [(G+1)^2] the channel being built (pellet sums incl. float pellets, then biggest own /
                       * enemy cell mass as u64 bit patterns, then the best virus radius) */
};
__host__ __device__ inline int obs_scratch_bytes(int G, bool full) {
    int nb = (G + 1) * (G + 1);
    int bytes = ((nb * 4 + 7) / 8) * 8 + 2 * G * 8;
    if (full) bytes += nb * 8;
    return bytes;
}
DEV ObsScratch obs_scratch(uint8_t* base, int G, bool full) {
    ObsScratch s;
    int nb = (G + 1) * (G + 1);
    s.pel_i = (int*)base;
    base += ((nb * 4 + 7) / 8) * 8;
    s.mid = (double*)base;
    base += 2 * G * 8;
    s.tab = full ? (double*)base : nullptr;
    return s;
}

/* spatialHashTable.py:91-108 getIdsForAreaFloatingPoint, one axis: the set { int(x / gs) : x = bucketLeft,
 * bucketLeft + gs, ... <= limit } as a bitmask, bit-identical to the reference's float arithmetic but without its
 * divisions and fmod:
 *   bucketLeft = cl - cl % gs   rounds the real number q*gs, q = floor(cl / gs)  ->  the rounded product q * gs;
 *   int(x / gs) for x within a few ulp of k*gs is k, unless the ROUNDED quotient falls below k: with the exact
 *   residual rho = k*gs - x (one fma), that happens iff rho > half_gap_below(k) * gs (a tie rounds to k: even).
 * inv = 1 / gs is only used for estimates that the exact residuals then settle.  Checked against the plain
 * formula on 6e7 adversarial inputs on the host and end to end by the GPU parity tests. */
/* mathematically exact floor(v / gs), v >= 0 (estimate from the reciprocal, settled by one exact fma residual) */
DEV int exact_floor_div(double v, double gs, double inv) {
    double q = floor(v * inv);
    double res = fma(-q, gs, v);
    q = res < 0 ? q - 1 : (res >= gs ? q + 1 : q);
    return (int)q;
}
/* columns as a 64-bit mask: G <= 63, cols <= 64 (agar_layout.h) */
DEV unsigned long long axis_buckets(double p, double radius, double fov, double gs, double inv, bool canon) {
    if (canon) { /* AGAR_OBS_CANONICAL: buckets floor(lo / gs) .. floor(hi / gs) with exact floors */
        double lo = py_max0(p - radius), hi = (p + radius < fov - 1) ? p + radius : fov - 1;
        if (hi < 0) return 0ull;
        int b0 = exact_floor_div(lo, gs, inv), b1 = exact_floor_div(hi, gs, inv);
        if (b1 < b0) return 0ull;
        return (b1 >= 63 ? ~0ull : ((2ull << b1) - 1)) & ~((1ull << b0) - 1);
    }
    double cl = py_max0(p - radius);
    double q = floor(cl * inv);
    double res = fma(-q, gs, cl);
    if (res < 0)
        q -= 1;
    else if (res >= gs)
        q += 1;
    double x = q * gs;
    const double lim = (p + radius < fov - 1) ? p + radius : fov - 1; /* min(size - 1, pos + radius) */
    unsigned long long m = 0;
    while (x <= lim) {
        double k = rint(x * inv);
        int col = (int)k;
        if (col > 0) {
            double rho = fma(k, gs, -x);
            if (rho > 0) {
                uint64_t kb = agar_double_to_bits(k);
                int e = (int)((kb >> 52) & 0x7ff);                         /* biased exponent of k          */
                int pow2 = (kb & 0x000fffffffffffffULL) == 0;              /* gap below a power of two is half */
                double thr = gs * agar_bits_to_double((uint64_t)(e - 53 - pow2) << 52);
                if (rho > thr) col -= 1;
            }
        }
        m |= 1ull << col;
        x += gs;
    }
    return m;
}
/* int(x / gs) for x within a few ulp of a multiple of gs — the exact-residual test of axis_buckets, branch-light */
DEV int bucket_of_edge(double x, double gs, double inv) {
    double k = rint(x * inv);
    int col = (int)k;
    double rho = fma(k, gs, -x);
    uint64_t kb = agar_double_to_bits(k);
    int e = (int)((kb >> 52) & 0x7ff);
    int pow2 = (kb & 0x000fffffffffffffULL) == 0;
    double thr = gs * agar_bits_to_double((uint64_t)(e - 53 - pow2) << 52);
    return (col > 0 && rho > thr) ? col - 1 : col;
}
/* axis_buckets for objects with 2 * radius < gs (every integer pellet: radius < 1, gs > 2.3 for G <= 16): the
 * reference loop visits x0 = bucketLeft and at most x1 = x0 + gs.  Straight-line, so a warp stays converged.
 * Returns the two bucket indices (or -1) instead of a mask. */
DEV void axis_buckets2(double p, double radius, double fov, double gs, double inv, bool canon, int& b0, int& b1) {
    if (canon) {
        double lo = py_max0(p - radius), hi = (p + radius < fov - 1) ? p + radius : fov - 1;
        int c0 = exact_floor_div(lo, gs, inv), c1 = hi >= 0 ? exact_floor_div(py_max0(hi), gs, inv) : -1;
        b0 = c1 >= c0 ? c0 : -1;
        b1 = c1 > c0 ? c1 : -1; /* 2 * radius < gs: at most two buckets */
        return;
    }
    double cl = py_max0(p - radius);
    double q = floor(cl * inv);
    double res = fma(-q, gs, cl);
    q = res < 0 ? q - 1 : (res >= gs ? q + 1 : q);
    const double x0 = q * gs, x1 = x0 + gs;
    const double lim = (p + radius < fov - 1) ? p + radius : fov - 1;
    int c0 = bucket_of_edge(x0, gs, inv), c1 = bucket_of_edge(x1, gs, inv);
    b0 = x0 <= lim ? c0 : -1;
    b1 = (x1 <= lim && c1 != c0) ? c1 : -1; /* ids is a set */
}
/* axis_buckets2 with the edge lookups from a per-observation bit table: for x0 = q * gs, int(x0 / gs) is q or q - 1, and
 * for x1 = x0 + gs, int(x1 / gs) is q + 1 or q — both depend on q only.  bit q of `edges`: x0 falls to q - 1;
 * bit 16 + q: x1 falls to q.  (built by edge_bits() once per observation) */
DEV unsigned edge_bits(double gs, double inv, int G) {
    unsigned m = 0;
    for (int k = 0; k <= G; ++k) {
        double x0 = (double)k * gs, x1 = x0 + gs;
        if (bucket_of_edge(x0, gs, inv) != k) m |= 1u << k;
        if (bucket_of_edge(x1, gs, inv) != k + 1) m |= 1u << (16 + k);
    }
    return m;
}
DEV void axis_buckets2_bits(double p, double radius, double fov, double gs, double inv, unsigned edges, int& b0, int& b1) {
    double cl = py_max0(p - radius);
    double q = floor(cl * inv);
    double res = fma(-q, gs, cl);
    q = res < 0 ? q - 1 : (res >= gs ? q + 1 : q);
    const double x0 = q * gs, x1 = x0 + gs;
    const double lim = (p + radius < fov - 1) ? p + radius : fov - 1;
    const int iq = (int)q;
    const int c0 = iq - (int)((edges >> iq) & 1u), c1 = iq + 1 - (int)((edges >> (16 + iq)) & 1u);
    b0 = x0 <= lim ? c0 : -1;
    b1 = (x1 <= lim && c1 != c0) ? c1 : -1;
}
/* axis_buckets for grids of more than 64 columns (CNN_INPUT_DIM_2 = 84): the same arithmetic, two mask words */
DEV void axis_buckets_wide(double p, double radius, double fov, double gs, double inv, bool canon, unsigned long long& lo,
                           unsigned long long& hi) {
    lo = hi = 0ull;
    auto set = [&](int col) {
        if (col < 64) lo |= 1ull << col;
        else hi |= 1ull << (col - 64);
    };
    if (canon) {
        double l = py_max0(p - radius), h = (p + radius < fov - 1) ? p + radius : fov - 1;
        if (h < 0) return;
        int b0 = exact_floor_div(l, gs, inv), b1 = exact_floor_div(h, gs, inv);
        for (int b = b0; b <= b1 && b < 128; ++b) set(b);
        return;
    }
    double cl = py_max0(p - radius);
    double q = floor(cl * inv);
    double res = fma(-q, gs, cl);
    q = res < 0 ? q - 1 : (res >= gs ? q + 1 : q);
    double x = q * gs;
    const double lim = (p + radius < fov - 1) ? p + radius : fov - 1;
    while (x <= lim) {
        set(bucket_of_edge(x, gs, inv));
        x += gs;
    }
}
/* calls f(id) once per distinct bucket of the object (ids is a set in the reference) */
template <class F>
DEV void for_each_fov_bucket(double ox, double oy, double radius, double left, double top, double fov, double gs,
                             double inv, int cols, bool canon, F f) {
    if (cols > 64) { /* uniform per observation */
        unsigned long long mx[2], my[2];
        axis_buckets_wide(ox - left, radius, fov, gs, inv, canon, mx[0], mx[1]);
        axis_buckets_wide(oy - top, radius, fov, gs, inv, canon, my[0], my[1]);
        for (int hx = 0; hx < 2; ++hx)
            for (unsigned long long a = mx[hx]; a; a &= a - 1) {
                int col = 64 * hx + __ffsll((long long)a) - 1;
                for (int hy = 0; hy < 2; ++hy)
                    for (unsigned long long t = my[hy]; t; t &= t - 1) f(col + (64 * hy + __ffsll((long long)t) - 1) * cols);
            }
        return;
    }
    unsigned long long mx = axis_buckets(ox - left, radius, fov, gs, inv, canon);
    const unsigned long long my = axis_buckets(oy - top, radius, fov, gs, inv, canon);
    while (mx) {
        int col = __ffsll((long long)mx) - 1;
        mx &= mx - 1;
        unsigned long long t = my;
        while (t) {
            int row = __ffsll((long long)t) - 1;
            t &= t - 1;
            f(col + row * cols);
        }
    }
}

/* Pellet slots of this lane whose integer position lies in the window [wx0, wx1] x [wy0, wy1]: f(slot, packed).
 * Two phases, so that the expensive per-candidate body runs with many lanes: a converged scan marks the candidates in a
 * per-lane bit mask (few of a big pool are in view: ~6 % in the 16-player arena, where the one-phase loop ran its body
 * with 1.5 lanes of 32), then every lane walks its own candidates.  No collectives inside f. */
template <int W, class F>
DEV void for_each_window_pellet(const Ctx<W>& c, const DevParams& P, int wx0, int wx1, int wy0, int wy1, F f) {
    const int cap = P.L.pellet_cap;
    /* worth it when the window filters: with most of the field in view (single-cell configs) one phase is cheaper */
    const bool filters = (long long)(wx1 - wx0 + 1) * (wy1 - wy0 + 1) * 2 < (long long)P.S * P.S;
    if (cap <= 64 * W && filters) {
        unsigned long long cand = 0;
        int j = 0;
        for (int s = c.lane; s < cap; s += W, ++j) {
            uint32_t pk = c.pel[s];
            int px = AGAR_PELLET_X(pk), py = AGAR_PELLET_Y(pk);
            bool in = px >= wx0 && px <= wx1 && py >= wy0 && py <= wy1 && pk != 0;
            cand |= (unsigned long long)in << j;
        }
        while (cand) {
            int jj = __ffsll((long long)cand) - 1;
            cand &= cand - 1;
            int s = c.lane + W * jj;
            f(s, c.pel[s]);
        }
    } else {
        for (int s = c.lane; s < cap; s += W) {
            uint32_t pk = c.pel[s];
            int px = AGAR_PELLET_X(pk), py = AGAR_PELLET_Y(pk);
            if (px < wx0 || px > wx1 || py < wy0 || py > wy1 || !pk) continue;
            f(s, pk);
        }
    }
}

/* cooperative; bot.py:326-497 + :302-323 + :272-299.  obs: this agent's row of the caller's buffer or nullptr. */
template <int W, bool FULL>
DEV void observe_agent(Ctx<W>& c, const DevParams& P, int k, int agent, float* obs) {
    const AgarConfig& cf = P.cfg;
    AgarPlayer* p = &c.pl[k];
    const int G = P.L.grid_squares, GG = G * G;
    const double S = (double)P.S;
    if (c.lane == 0 && !c.fov_done) update_fov(c, P, k);
    c.t.sync();
    const double fov = p->fov_size, fx = p->fov_x, fy = p->fov_y;
    const double left = fx - fov / 2, top = fy - fov / 2;
    const double gs = fov / G, inv = 1.0 / gs;
    const bool canon = cf.obs_mode == AGAR_OBS_CANONICAL;
    const int cols = canon ? G : (int)ceil(fov / gs); /* spatialHashTable.py:19 */
    const int nbk = cols * cols;
    ObsScratch sc = obs_scratch(c.scratch, G, FULL);
    unsigned long long* tabu = (unsigned long long*)sc.tab;
    for (int i = c.lane; i < nbk; i += W) sc.pel_i[i] = 0;
    if (c.lane == 0) { /* gsMidPoint sequences: mid += gsSize, accumulated exactly like the reference loop */
        double mx = left + gs / 2, my = top + gs / 2;
        for (int i = 0; i < G; ++i) {
            sc.mid[i] = mx;
            sc.mid[G + i] = my;
            mx += gs;
            my += gs;
        }
    }
    c.t.sync();
    const Rect ra = rect_of_call(P.S, fx, fy, fov / 2);
    /* channel slots in bot.py:458-495 order: PELLET, SELF, WALL, ENEMY, ALL_PLAYER, VIRUS, SELF_SLF, SELF_LF, ENEMY_SLF, ENEMY_LF */
    int ch = 0;
    const int ch_pel = cf.pellet_grid ? ch++ : -1;
    const int ch_self = (FULL && cf.self_grid) ? ch++ : -1;
    const int ch_wall = (FULL && cf.wall_grid) ? ch++ : -1;
    const int ch_enemy = (FULL && cf.enemy_grid) ? ch++ : -1;
    const int ch_all = (FULL && cf.all_player_grid) ? ch++ : -1;
    const int ch_virus = (FULL && cf.virus_grid) ? ch++ : -1;
    const int ch_self_slf = (FULL && cf.self_grid_slf) ? ch++ : -1;
    const int ch_self_lf = (FULL && cf.self_grid_lf) ? ch++ : -1;
    const int ch_enemy_slf = (FULL && cf.enemy_grid_slf) ? ch++ : -1;
    const int ch_enemy_lf = (FULL && cf.enemy_grid_lf) ? ch++ : -1;
    float* hist = (FULL && P.L.n_hist) ? c.hist + (size_t)agent * P.L.n_hist * GG : nullptr;
    /* square idx of the G x G view (bucket index uses G, not cols: the reference's shear when cols == G + 1): does it
     * show anything?  squares entirely outside the field do not (bot.py:392-393) */
    auto inside = [&](int idx) {
        int cc = G > 1 ? (int)__umulhi((unsigned)idx, P.g_magic) : idx, r = idx - cc * G;
        double midx = sc.mid[r], midy = sc.mid[G + cc];
        return !(midx + gs / 2 < 0 || midx - gs / 2 > S || midy + gs / 2 < 0 || midy - gs / 2 > S);
    };
    /* ---- PELLET: integer pellets (field.getPelletsInFov -> insertAllFloatingPointObjects): integer adds commute */
    /* integer window that contains every pellet in_fov() can accept (radius < 1) */
    const int wx0 = (int)floor(fx - fov / 2) - 1, wx1 = (int)ceil(fx + fov / 2) + 1;
    const int wy0 = (int)floor(fy - fov / 2) - 1, wy1 = (int)ceil(fy + fov / 2) + 1;
    for_each_window_pellet(c, P, wx0, wx1, wy0, wy1, [&](int, uint32_t pk) {
        int px = AGAR_PELLET_X(pk), py = AGAR_PELLET_Y(pk), pm = AGAR_PELLET_M(pk);
        double pr = P.pellet_r[pm & 3];
        if (!rect_hit(ra, pellet_rect(px, py)) || !in_fov((double)px, (double)py, pr, fx, fy, fov)) return;
        for_each_fov_bucket((double)px, (double)py, pr, left, top, fov, gs, inv, cols, canon,
                            [&](int id) { atomicAdd(&sc.pel_i[id], pm); });
    });
    c.t.sync();
    if (!FULL) {
        for (int idx = c.lane; idx < GG; idx += W)
            if (obs && ch_pel >= 0) obs[ch_pel * GG + idx] = inside(idx) ? (float)(double)sc.pel_i[idx] : 0.f;
    } else {
        for (int i = c.lane; i < nbk; i += W) sc.tab[i] = (double)sc.pel_i[i];
        c.t.sync();
        /* float pellets are added after the integer ones, in slot order (canonical candidate order; the float64 sums depend on
         * it).  The slots are tested W at a time — one memory latency per round instead of one per slot on a lone lane — and
         * only the few that are in view add their mass, one after the other in ascending slot order. */
        if (c.h->n_fat)
            for (int base = 0; base < P.L.fat_cap; base += W) {
                const int sl = base + c.lane;
                AgarFatPellet f = {};
                if (sl < P.L.fat_cap) f = c.fat[sl];
                const bool hit = f.mass != 0 && rect_hit(ra, rect_of_call(P.S, f.x, f.y, f.radius)) && in_fov(f.x, f.y, f.radius, fx, fy, fov);
                unsigned hits = c.t.ballot(hit);
                while (hits) {
                    const int b = __ffs((int)hits) - 1;
                    hits &= hits - 1;
                    if (c.lane == b) {
                        const double fm = f.mass;
                        for_each_fov_bucket(f.x, f.y, f.radius, left, top, fov, gs, inv, cols, canon,
                                            [&](int id) { sc.tab[id] = sc.tab[id] + fm; });
                    }
                    c.t.sync();
                }
            }
        c.t.sync();
        for (int idx = c.lane; idx < GG; idx += W) {
            if (obs && ch_pel >= 0) obs[ch_pel * GG + idx] = inside(idx) ? (float)sc.tab[idx] : 0.f;
            if (ch_wall >= 0) { /* bot.py:443-450 */
                int cc = G > 1 ? (int)__umulhi((unsigned)idx, P.g_magic) : idx, r = idx - cc * G;
                double midx = sc.mid[r], midy = sc.mid[G + cc];
                double lb = py_minS(S, py_max0(midx - gs / 2)), tb = py_minS(S, py_max0(midy - gs / 2));
                double rb = py_max0(py_minS(S, midx + gs / 2)), bb = py_max0(py_minS(S, midy + gs / 2));
                double free_area = (rb - lb) * (bb - tb);
                double w = agar_round_dec(1 - (free_area / agar_pow(gs, 2.0)), 1e3); /* gsSize ** 2 == C pow */
                if (obs) obs[ch_wall * GG + idx] = (float)w;
            }
        }
        /* ---- SELF then ENEMY: biggest cell mass per square; max commutes and positive doubles order like their bit
         * patterns.  Own cells need no hash lookup (bot.py:344), enemies come through the player table (field.py:437-439). */
        const int K = P.L.n_players, cap = P.L.cell_cap;
        double biggest = 0.0; /* NORMALIZE_GRID_BY_MAX_MASS (bot.py:364-367): mass of the biggest player cell in view, own or enemy */
        if (cf.normalize_grid_by_max_mass) {
            const uint16_t* live = live_cells(c, P);
            const int n_it = c.n_live >= 0 ? c.n_live : K * cap;
            for (int t = c.lane; t < n_it; t += W) {
                int idx = c.n_live >= 0 ? (int)live[t] : t;
                int k2 = idx >> P.cap_shift, j = idx - (k2 << P.cap_shift);
                if (j >= c.pl[k2].n_cells) continue;
                const AgarCell* o = CELLP(c, P, k2, j);
                if (k2 == k) {
                    if (!in_fov(o->x, o->y, o->radius, fx, fy, fov)) continue;
                } else if (!(o->flags & AGAR_CF_INHASH) || !rect_hit(ra, rect_of_call(P.S, o->x, o->y, o->radius)) ||
                           !in_fov(o->x, o->y, o->radius, fx, fy, fov))
                    continue;
                if (o->mass > biggest) biggest = o->mass;
            }
            for (int off = W / 2; off > 0; off >>= 1) { /* max over the tile: positive doubles, any order */
                const double other = c.t.shfl_xor(biggest, off);
                if (other > biggest) biggest = other;
            }
        }
        for (int pass = 0; pass < 3; ++pass) { /* 0: own cells, 1: enemy cells, 2: ALL_PLAYER_GRID (both, bot.py:348-351) */
            const bool own = pass == 0, all = pass == 2;
            const int ch_now = all ? ch_all : (own ? ch_self : ch_enemy), ch_slf = all ? -1 : (own ? ch_self_slf : ch_enemy_slf),
                      ch_lf = all ? -1 : (own ? ch_self_lf : ch_enemy_lf);
            if (ch_now < 0 && ch_slf < 0 && ch_lf < 0) continue;
            c.t.sync();
            for (int i = c.lane; i < nbk; i += W) tabu[i] = 0ull;
            c.t.sync();
            const uint16_t* live = live_cells(c, P);
            const int n_it = c.n_live >= 0 ? c.n_live : K * cap;
            for (int t = c.lane; t < n_it; t += W) {
                int idx = c.n_live >= 0 ? (int)live[t] : t;
                int k2 = idx >> P.cap_shift, j = idx - (k2 << P.cap_shift);
                if ((!all && (k2 == k) != own) || j >= c.pl[k2].n_cells) continue;
                const AgarCell* o = CELLP(c, P, k2, j);
                if (k2 == k) {
                    if (!in_fov(o->x, o->y, o->radius, fx, fy, fov)) continue;
                } else if (!(o->flags & AGAR_CF_INHASH) || !rect_hit(ra, rect_of_call(P.S, o->x, o->y, o->radius)) ||
                           !in_fov(o->x, o->y, o->radius, fx, fy, fov))
                    continue;
                unsigned long long bits = (unsigned long long)__double_as_longlong(o->mass);
                for_each_fov_bucket(o->x, o->y, o->radius, left, top, fov, gs, inv, cols, canon,
                                    [&](int id) { atomicMax(&tabu[id], bits); });
            }
            c.t.sync();
            float* h_lf = hist ? hist + (own ? 0 : 2) * GG : nullptr;  /* last observation's grid */
            float* h_slf = hist ? hist + (own ? 1 : 3) * GG : nullptr; /* the one before */
            for (int idx = c.lane; idx < GG; idx += W) {
                double v = inside(idx) ? sc.tab[idx] : 0.0; /* 0 bits == 0.0: empty square */
                if (cf.normalize_grid_by_max_mass && v != 0.0) v = v / biggest; /* bot.py:412,422,430 */
                if (obs && ch_now >= 0) obs[ch_now * GG + idx] = (float)v;
                if (ch_slf >= 0) {
                    if (obs) obs[ch_slf * GG + idx] = h_slf[idx];
                    h_slf[idx] = h_lf[idx];
                }
                if (ch_lf >= 0) {
                    if (obs) obs[ch_lf * GG + idx] = h_lf[idx];
                    h_lf[idx] = (float)v;
                }
            }
        }
        /* ---- VIRUS: mass of the first virus with the largest radius (bot.py:436-441); pel_i now holds 1 + its index */
        if (ch_virus >= 0) {
            c.t.sync();
            for (int i = c.lane; i < nbk; i += W) sc.tab[i] = 0.0, sc.pel_i[i] = 0;
            c.t.sync();
            if (c.lane == 0 && cf.virus_enabled)
                for (int v = 0; v < c.h->n_viruses; ++v) {
                    const AgarMote* o = &c.vir[v];
                    if (!(o->aux & AGAR_CF_INHASH) || !rect_hit(ra, rect_of_call(P.S, o->x, o->y, o->radius)) ||
                        !in_fov(o->x, o->y, o->radius, fx, fy, fov))
                        continue;
                    double orr = o->radius;
                    for_each_fov_bucket(o->x, o->y, o->radius, left, top, fov, gs, inv, cols, canon, [&](int id) {
                        if (orr > sc.tab[id]) sc.tab[id] = orr, sc.pel_i[id] = v + 1;
                    });
                }
            c.t.sync();
            for (int idx = c.lane; idx < GG; idx += W) {
                int vi = sc.pel_i[idx];
                if (obs) obs[ch_virus * GG + idx] = (inside(idx) && vi) ? (float)c.vir[vi - 1].mass : 0.f;
            }
        }
    }
    if (c.lane == 0) { /* getAdditionalFeatures bot.py:302-323 */
        AgarBot* B = &p->bot;
        int n = GG * P.L.n_grids;
        if (cf.use_last_fovsize) {
            B->last_fov_size_feat = B->fov_size_feat;
            if (obs) obs[n] = (float)B->last_fov_size_feat;
            ++n;
        }
        if (cf.use_fovsize) {
            B->fov_size_feat = p->fov_size;
            if (obs) obs[n] = (float)B->fov_size_feat;
            ++n;
        }
        if (cf.use_totalmass) {
            if (obs) obs[n] = (float)total_mass(c, P, k);
            ++n;
        }
        if (cf.use_last_action)
            for (int i = 0; i < 4; ++i, ++n)
                if (obs) obs[n] = (float)(B->has_action ? B->cur_action[i] : 0.0);
        if (cf.use_second_last_action)
            for (int i = 0; i < 4; ++i, ++n)
                if (obs) obs[n] = (float)(B->has_last_action ? B->last_action[i] : 0.0);
    }
    c.t.sync();
}

/* lane 0; getReward bot.py:654-667 */
template <int W>
DEV double bot_reward(const Ctx<W>& c, const DevParams& P, int k) {
    const AgarConfig& cf = P.cfg;
    const AgarPlayer* p = &c.pl[k];
    if (cf.mass_as_reward) return p->alive ? total_mass(c, P, k) - cf.reward_term : cf.death_term - cf.reward_term;
    double reward;
    if (!p->alive)
        reward = -1 * p->bot.last_mass * cf.death_factor + cf.death_term;
    else
        reward = total_mass(c, P, k) - p->bot.last_mass;
    return reward * cf.reward_scale - cf.reward_term;
}
/* lane 0; set_command_point bot.py:550-577 + Player.setCommands */
template <int W>
DEV void set_command_point(Ctx<W>& c, const DevParams& P, int k, double a0, double a1, double a2, double a3, int len,
                           bool fov_fresh) {
    AgarPlayer* p = &c.pl[k];
    if (!fov_fresh && !c.fov_done) update_fov(c, P, k); /* fresh: computed a moment ago in this bot turn from the same cells */
    int x = (int)p->fov_x, y = (int)p->fov_y;
    int left = x - (int)(p->fov_size / 2), top = y - (int)(p->fov_size / 2);
    int size = (int)p->fov_size;
    p->cmd_x = left + a0 * size;
    p->cmd_y = top + a1 * size;
    int split = 0, eject = 0;
    if (len == 3)
        split = a2 > 0.5;
    else if (len == 4) {
        split = a2 > 0.5;
        eject = a3 > 0.5;
    }
    p->do_split = split;
    p->do_eject = eject;
}

/* cooperative; Bot.getSimpleStateRepresentation (bot.py:511-548; GRID_VIEW_ENABLED = False, networkParameters.py:119): first own cell,
 * closest enemy cell in the field of view, closest pellet in the INTEGER field of view (relative positions / radii rounded to 5
 * decimals), distances to the visible field edges.  `min(..., key=squaredDistance)` keeps the first minimal candidate of the list:
 * every lane keeps its own first minimum over a strided share, the tile reduces (distance, canonical position) lexicographically. */
template <int W>
DEV void simple_closest_reduce(Ctx<W>& c, double& d, int& ord) {
    for (int off = W / 2; off > 0; off >>= 1) {
        const double od = c.t.shfl_xor(d, off);
        const int oo = c.t.shfl_xor(ord, off);
        if (oo >= 0 && (ord < 0 || od < d || (od == d && oo < ord))) d = od, ord = oo;
    }
}
template <int W>
DEV void simple_state_agent(Ctx<W>& c, const DevParams& P, int k, float* obs) {
    AgarPlayer* p = &c.pl[k];
    if (c.lane == 0 && !c.fov_done) update_fov(c, P, k);
    c.t.sync();
    if (obs == nullptr) return; /* a decision inside a multi-frame step: only the field-of-view caches advance */
    const double fov = p->fov_size, fx = p->fov_x, fy = p->fov_y;
    const int x = (int)fx, y = (int)fy;
    const int left = x - (int)(fov / 2), top = y - (int)(fov / 2), size = (int)fov;
    const AgarCell* first = CELLP(c, P, k, 0);
    const double ox = first->x, oy = first->y;
    const int K = P.L.n_players, cap = P.L.cell_cap;
    /* closest enemy cell, field.py:434-436 through the player table (rectangle + INHASH) */
    double ed = 0.0;
    int eo = -1;
    {
        const Rect ra = rect_of_call(P.S, fx, fy, fov / 2);
        for (int idx = c.lane; idx < K * cap; idx += W) {
            const int k2 = idx >> P.cap_shift, j = idx - (k2 << P.cap_shift);
            if (k2 == k || j >= c.pl[k2].n_cells) continue;
            const AgarCell* o = CELLP(c, P, k2, j);
            if (!(o->flags & AGAR_CF_INHASH) || !rect_hit(ra, rect_of_call(P.S, o->x, o->y, o->radius)) || !in_fov(o->x, o->y, o->radius, fx, fy, fov))
                continue;
            const double d = (o->x - ox) * (o->x - ox) + (o->y - oy) * (o->y - oy); /* cell.py:158-160 */
            if (eo < 0 || d < ed) ed = d, eo = idx;
        }
        simple_closest_reduce<W>(c, ed, eo);
    }
    /* closest pellet, getPelletsInFov(midPoint, int(size)) field.py:442-444: integer pellets by slot, then the ex-blob pellets */
    double pd = 0.0;
    int po = -1;
    {
        const double isz = (double)size;
        const Rect ra = rect_of_call(P.S, fx, fy, isz / 2);
        for (int s = c.lane; s < P.L.pellet_cap; s += W) {
            const uint32_t pk = c.pel[s];
            if (!pk) continue;
            const int px = AGAR_PELLET_X(pk), py = AGAR_PELLET_Y(pk), pm = AGAR_PELLET_M(pk);
            if (!rect_hit(ra, pellet_rect(px, py)) || !in_fov((double)px, (double)py, P.pellet_r[pm & 3], fx, fy, isz)) continue;
            const double d = ((double)px - ox) * ((double)px - ox) + ((double)py - oy) * ((double)py - oy);
            if (po < 0 || d < pd) pd = d, po = s;
        }
        for (int s = c.lane; s < P.L.fat_cap; s += W) {
            const AgarFatPellet* f = &c.fat[s];
            if (f->mass == 0) continue;
            if (!rect_hit(ra, rect_of_call(P.S, f->x, f->y, f->radius)) || !in_fov(f->x, f->y, f->radius, fx, fy, isz)) continue;
            const double d = (f->x - ox) * (f->x - ox) + (f->y - oy) * (f->y - oy);
            if (po < 0 || d < pd) pd = d, po = P.L.pellet_cap + s; /* a lane meets its integer pellets first: ascending canonical position */
        }
        simple_closest_reduce<W>(c, pd, po);
    }
    if (c.lane == 0) {
        const double dsz = (double)size, S = (double)P.S;
        auto rel = [&](double v, int base) { return (float)agar_round_dec((v - base) / dsz, 1e5); }; /* getRelativeCellPos :16-20 */
        auto rad = [&](double r) { return (float)agar_round_dec(r <= dsz ? r / dsz : 1.0, 1e5); };  /* isRelativeCellData :635-637 */
        obs[0] = rel(ox, left), obs[1] = rel(oy, top), obs[2] = rad(first->radius);
        if (eo >= 0) {
            const AgarCell* o = c.cells + eo;
            obs[3] = rel(o->x, left), obs[4] = rel(o->y, top), obs[5] = rad(o->radius);
        } else
            obs[3] = obs[4] = obs[5] = 0.f;
        if (po >= P.L.pellet_cap) {
            const AgarFatPellet* f = &c.fat[po - P.L.pellet_cap];
            obs[6] = rel(f->x, left), obs[7] = rel(f->y, top);
        } else if (po >= 0) {
            const uint32_t pk = c.pel[po];
            obs[6] = rel((double)AGAR_PELLET_X(pk), left), obs[7] = rel((double)AGAR_PELLET_Y(pk), top);
        } else
            obs[6] = obs[7] = 0.f;
        obs[8] = left <= 0 ? (float)((double)x / dsz) : 1.f; /* :541-547 */
        obs[9] = left + size >= P.S ? (float)((S - x) / dsz) : 1.f;
        obs[10] = top <= 0 ? (float)((double)y / dsz) : 1.f;
        obs[11] = top + size >= P.S ? (float)((S - y) / dsz) : 1.f;
    }
    c.t.sync();
}

/* cooperative; first half of move_NN (bot.py:195-217) + makeMove :253 */
template <int W, bool FULL>
DEV void nn_turn_begin(Ctx<W>& c, const DevParams& P, int k, float* obs) {
    AgarPlayer* p = &c.pl[k];
    AgarBot* B = &p->bot;
    int do_obs = 0;
    if (c.lane == 0 && !B->turn_begun) {
        double tm = total_mass(c, P, k);
        B->stat_mass_sum += tm;
        if (tm > B->stat_mass_max) B->stat_mass_max = tm;
        B->stat_frames += 1;
        B->skipping = 0;
        B->need_action = B->exp_valid = B->exp_done = 0;
        if (B->has_action) {
            if (B->has_last_mass && B->last_mass != 0) B->cum_reward += bot_reward(c, P, k);
            B->last_reward = B->cum_reward;
            if (B->skip_frames > 0) {
                B->skip_frames -= 1;
                if (p->alive) B->skipping = 1;
            }
        }
        if (!B->skipping) {
            if (B->has_old_state) {
                B->time += 1;
                B->exp_valid = 1;
                B->exp_done = !p->alive;
            }
            B->need_action = p->alive;
            do_obs = p->alive;
        }
        B->turn_begun = 1;
    }
    do_obs = c.t.shfl(do_obs, 0);
    CLK_IN(c, 9); /* reward / frame-skip bookkeeping */
    if (do_obs) {
        if (FULL && P.cfg.simple_state)
            simple_state_agent<W>(c, P, k, obs);
        else
            observe_agent<W, FULL>(c, P, k, k, obs);
    }
    CLK_IN(c, 10); /* observation */
}
/* lane 0; second half of move_NN (bot.py:223-232) + tail of makeMove (:256-270) */
template <int W>
DEV void nn_turn_end(Ctx<W>& c, const DevParams& P, int k, const float* action) {
    AgarPlayer* p = &c.pl[k];
    AgarBot* B = &p->bot;
    if (B->need_action) {
        B->cum_reward = 0;
        B->skip_frames = P.cfg.frame_skip;
        B->has_old_state = 1;
        for (int i = 0; i < 4; ++i) B->last_action[i] = B->cur_action[i];
        B->has_last_action = B->has_action;
        for (int i = 0; i < 4; ++i) B->cur_action[i] = i < P.L.action_len ? (double)action[i] : 0.0;
        B->has_action = 1;
    }
    if (!B->skipping && p->alive) {
        B->last_mass = total_mass(c, P, k);
        B->has_last_mass = 1;
    }
    B->turn_begun = 0;
    if (!p->alive) return;
    int len = P.L.action_len;
    double a2 = B->cur_action[2], a3 = B->cur_action[3];
    if (B->skipping) a2 = a3 = 0, len = 4;
    set_command_point(c, P, k, B->cur_action[0], B->cur_action[1], a2, a3, len, B->need_action != 0 /* observed this turn */);
}

/* cooperative; make_greedy_bot_move bot.py:579-633 / make_random_bot_move :243-249 */
template <int W>
DEV void scripted_turn(Ctx<W>& c, const DevParams& P, int k) {
    AgarPlayer* p = &c.pl[k];
    AgarBot* B = &p->bot;
    if (c.lane == 0) {
        double tm = total_mass(c, P, k);
        B->stat_mass_sum += tm;
        if (tm > B->stat_mass_max) B->stat_mass_max = tm;
        B->stat_frames += 1;
        if (p->alive && !c.fov_done) update_fov(c, P, k);
    }
    c.t.sync();
    if (!p->alive) return; /* uniform: nothing below changes alive */
    if (B->type == AGAR_BOT_GREEDY) {
        const double fx = p->fov_x, fy = p->fov_y, fov = p->fov_size;
        int big = 0;
        const AgarCell* base = CELLP(c, P, k, 0);
        for (int i = 1; i < p->n_cells; ++i)
            if (base[i].mass > base[big].mass) big = i;
        const double bxx = base[big].x, byy = base[big].y, bm = base[big].mass;
        const Rect ra = rect_of_call(P.S, fx, fy, fov / 2);
        /* argmax of mass / d^2 with first-max-wins in canonical order: order index = position in the candidate list */
        double best = -1.0, bestx = 0, besty = 0;
        int best_ord = 0x7fffffff;
        auto consider = [&](double ox, double oy, double om, int ord) {
            double d2 = (ox - bxx) * (ox - bxx) + (oy - byy) * (oy - byy);
            double key = om / (d2 != 0 ? d2 : 1);
            if (key > best || (key == best && ord < best_ord)) best = key, bestx = ox, besty = oy, best_ord = ord;
        };
        int ord0 = 0;
        const int wx0 = (int)floor(fx - fov / 2) - 1, wx1 = (int)ceil(fx + fov / 2) + 1;
        const int wy0 = (int)floor(fy - fov / 2) - 1, wy1 = (int)ceil(fy + fov / 2) + 1;
        for_each_window_pellet(c, P, wx0, wx1, wy0, wy1, [&](int s, uint32_t pk) { /* integer FOV window first */
            int px = AGAR_PELLET_X(pk), py = AGAR_PELLET_Y(pk), pm = AGAR_PELLET_M(pk);
            double pr = P.pellet_r[pm & 3];
            if (!rect_hit(ra, pellet_rect(px, py)) || !in_fov((double)px, (double)py, pr, fx, fy, fov)) return;
            consider((double)px, (double)py, (double)pm, ord0 + s);
        });
        ord0 += P.L.pellet_cap;
        if (c.h->n_fat) /* mostly free slots: four independent mass loads in flight per lane, then the rare live ones */
            for (int base = c.lane; base < P.L.fat_cap; base += 4 * W) {
                double fm[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) fm[i] = base + i * W < P.L.fat_cap ? c.fat[base + i * W].mass : 0.0;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (fm[i] == 0) continue;
                    const int s = base + i * W;
                    const AgarFatPellet* f = &c.fat[s];
                    if (!rect_hit(ra, rect_of_call(P.S, f->x, f->y, f->radius)) || !in_fov(f->x, f->y, f->radius, fx, fy, fov)) continue;
                    consider(f->x, f->y, fm[i], ord0 + s);
                }
            }
        ord0 += P.L.fat_cap;
        const int K = P.L.n_players, cap = P.L.cell_cap;
        const uint16_t* live = live_cells(c, P);
        const int n_it = c.n_live >= 0 ? c.n_live : K * cap;
        for (int t = c.lane; t < n_it; t += W) {
            int idx = c.n_live >= 0 ? (int)live[t] : t;
            int k2 = idx >> P.cap_shift, j = idx - (k2 << P.cap_shift);
            if (k2 == k || j >= c.pl[k2].n_cells) continue;
            const AgarCell* o = CELLP(c, P, k2, j);
            if (!(o->flags & AGAR_CF_INHASH) || !rect_hit(ra, rect_of_call(P.S, o->x, o->y, o->radius)) ||
                !in_fov(o->x, o->y, o->radius, fx, fy, fov))
                continue;
            if (bm > 1.25 * o->mass) consider(o->x, o->y, o->mass, ord0 + idx);
        }
        ord0 += K * cap;
        if (P.cfg.virus_enabled)
            for (int v = c.lane; v < c.h->n_viruses; v += W) {
                const AgarMote* o = &c.vir[v];
                if (!(o->aux & AGAR_CF_INHASH) || !rect_hit(ra, rect_of_call(P.S, o->x, o->y, o->radius)) ||
                    !in_fov(o->x, o->y, o->radius, fx, fy, fov))
                    continue;
                if (bm > 1.25 * o->mass) consider(o->x, o->y, o->mass, ord0 + v);
            }
        /* tile argmax (key desc, order asc) */
        for (int off = W / 2; off > 0; off >>= 1) {
            double ob = c.t.shfl_xor(best, off), ox = c.t.shfl_xor(bestx, off), oy = c.t.shfl_xor(besty, off);
            int oo = c.t.shfl_xor(best_ord, off);
            if (ob > best || (ob == best && oo < best_ord)) best = ob, bestx = ox, besty = oy, best_ord = oo;
        }
        if (c.lane == 0) {
            if (best_ord != 0x7fffffff) { /* getRelativeCellPos bot.py:16-20: relative to the FLOAT fov size */
                int x = (int)fx, y = (int)fy;
                int left = x - (int)(fov / 2), top = y - (int)(fov / 2);
                B->cur_action[0] = agar_round_dec((bestx - left) / fov, 1e5);
                B->cur_action[1] = agar_round_dec((besty - top) / fov, 1e5);
            } else {
                B->cur_action[0] = draw_random(c, P, 1);
                B->cur_action[1] = draw_random(c, P, 1);
            }
            B->cur_action[2] = B->cur_action[3] = 0;
        }
    } else if (c.lane == 0) {
        if (P.cfg.frame_skip == 0 || B->time % P.cfg.frame_skip == 0) {
            B->cur_action[0] = draw_random(c, P, 1);
            B->cur_action[1] = draw_random(c, P, 1);
            B->cur_action[2] = P.cfg.enable_split ? draw_random(c, P, 1) : 0;
            B->cur_action[3] = P.cfg.enable_eject ? draw_random(c, P, 1) : 0;
        }
        B->time += 1;
    }
    if (c.lane == 0) {
        B->has_action = 1;
        set_command_point(c, P, k, B->cur_action[0], B->cur_action[1], B->cur_action[2], B->cur_action[3], 4, true);
    }
    c.t.sync();
}

/* lane 0; Bot.reset bot.py:125-164 (history grids are cleared by the caller cooperatively) */
template <int W>
DEV void bot_reset(Ctx<W>& c, const DevParams& P, int k) {
    AgarBot* B = &c.pl[k].bot;
    B->has_last_mass = 0, B->last_mass = 0;
    B->has_old_state = 0;
    B->skip_frames = 0;
    B->cum_reward = 0, B->last_reward = 0;
    B->skipping = 0;
    B->turn_begun = B->need_action = B->exp_valid = B->exp_done = 0;
    for (int i = 0; i < 4; ++i) B->cur_action[i] = 0;
    if (B->type == AGAR_BOT_NN) {
        B->has_action = 0;
        B->fov_size_feat = 0, B->last_fov_size_feat = 0;
    } else
        B->has_action = 1;
}
