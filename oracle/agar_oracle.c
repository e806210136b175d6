/*
 * agar_oracle.c — CPU restatement of the agar.io env step of NILOIDE/A.I.gar.
 *
 * TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; the product (a.i.gar_b200/) never does.
 *
 * One env, sequential, double precision, operating on the env RECORD of include/agar_b200.h.  Every
 * function cites the reference lines (relative to /root/reference/src/model/) it restates.  Parity pin:
 * oracle/ref_harness.py runs the unmodified reference Python with injected Philox draws and a canonical
 * candidate order; tests/test_oracle_vs_reference.py requires this file (default libm build) to reproduce
 * the reference's records BIT FOR BIT over rollouts of every config, and tests/golden/ holds records the
 * reference produced so that the same check runs where /root/reference is absent.
 *
 * Two builds (oracle/Makefile):
 *   libagar_oracle.so     libm transcendental functions, exactly as CPython calls them
 *   libagar_oracle_pm.so  -DAGAR_PORTABLE_MATH: include/agar_math.h (the arithmetic the CUDA kernels use; since round 2 it is
 *                         bit-identical to this image's libm, so the two builds give identical results)
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/agar_b200.h"
#include "../include/agar_layout.h"
#include "../include/agar_math.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------ constants (parameters.py:10-35) */
#define FPS 30
#define BUCKET 20
#define START_MASS 10.0
#define VIRUS_BASE_SIZE 100.0
#define VIRUS_EAT_FACTOR 0.5
#define VIRUS_EXPLOSION_PROPORTION 0.6
#define EJECT_BASE_MASS 18
#define MAX_MASS 22500.0
#define BASE_MERGE_TIME 25
#define MERGE_TIME_MASS_FACTOR 0.0233
#define MERGE_TIME_VIRUS_FACTOR 0.85

typedef struct OracleEnv {
    AgarConfig cfg;
    AgarLayout L;
    uint64_t seed, env_id;
    uint8_t* rec;
    AgarEnvHeader* h;
    AgarPlayer* pl;
    AgarCell* cells;
    AgarMote* vir;
    AgarMote* blob;
    AgarFatPellet* fat;
    uint32_t* pel;
    float* hist;
    AgarEvent* ev;
    int S, nb; /* field size; world hash cols == rows (spatialHashTable.py:19) */
    double speed_modifier, move_speed, decay_rate, blob_mass, virus_split_mass, start_radius, virus_radius;
    double cos_deg[360], sin_deg[360], pow_n[17];
    int owns_rec;
} OracleEnv;

#define CELLP(e, k, i) (&(e)->cells[(size_t)(k) * (e)->L.cell_cap + (i)])

/* ------------------------------------------------------------------ math selection */
#ifdef AGAR_PORTABLE_MATH
static inline void dir_of(double dy, double dx, double* c, double* s) { agar_dir(dy, dx, c, s); }
static inline double m_pow(double x, double y) { return agar_pow(x, y); }
static inline double m_sq(double x) { return agar_pow(x, 2.0); } /* gsSize ** 2 == C pow(x, 2.0), bot.py:449 */
static inline double round_dec(double x, int nd) { return agar_round_dec(x, nd == 5 ? 1e5 : 1e3); }
#else
/* cell.py:55-57 / 49-51: angle = math.atan2(yDiff, xDiff); math.cos(angle), math.sin(angle) */
static inline void dir_of(double dy, double dx, double* c, double* s) {
    double a = atan2(dy, dx);
    *c = cos(a);
    *s = sin(a);
}
static inline double m_pow(double x, double y) { return pow(x, y); }
static inline double m_sq(double x) { return pow(x, 2.0); } /* gsSize ** 2, bot.py:449 */
/* Python round(x, nd): correctly rounded decimal of the exact binary value, then back to double */
static inline double round_dec(double x, int nd) {
    char buf[64];
    snprintf(buf, sizeof buf, "%.*f", nd, x);
    return strtod(buf, NULL);
}
#endif

static inline double py_max0(double v) { return v > 0 ? v : 0.0; }              /* max(0, v)  */
static inline double py_minS(double S, double v) { return v < S ? v : S; }       /* min(S, v)  */
static inline double clampS(double v, double S) { return py_minS(S, py_max0(v)); } /* min(S, max(0, v)) */
static inline double radius_of(double mass) { return mass > 0 ? sqrt(mass / M_PI) : 0.0; } /* cell.py:210-212 */

/* numpy.sum of a short list == numpy's pairwise summation kernel (verified against numpy 2.3 for n<=20) */
static double np_sum(const double* a, int n) {
    if (n < 8) {
        double r = 0.0;
        for (int i = 0; i < n; ++i) r += a[i];
        return r;
    }
    double r[8];
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i = 8;
    for (; i < n - (n % 8); i += 8)
        for (int j = 0; j < 8; ++j) r[j] += a[i + j];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
}

/* ------------------------------------------------------------------ Philox4x32-10 (oracle/philox.py) */
static void philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1,
                 n3 = (uint32_t)p0;
        c0 = n0, c1 = n1, c2 = n2, c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0, out[1] = c1, out[2] = c2, out[3] = c3;
}
static void draw_words(OracleEnv* e, int stream, uint32_t out[4]) {
    uint32_t* serial = stream == 0 ? &e->h->rng_field : &e->h->rng_bot;
    philox(*serial, (uint32_t)stream, (uint32_t)e->env_id, 0, (uint32_t)e->seed, (uint32_t)(e->seed >> 32), out);
    *serial += 1;
}
/* numpy.random.randint(lo, hi) with float bounds truncated toward zero */
static int64_t draw_randint(OracleEnv* e, int stream, double lo, double hi) {
    int64_t l = (int64_t)lo, h = (int64_t)hi;
    uint32_t w[4];
    draw_words(e, stream, w);
    return l + (int64_t)(((uint64_t)w[0] * (uint64_t)(h - l)) >> 32);
}
static double draw_random(OracleEnv* e, int stream) {
    uint32_t w[4];
    draw_words(e, stream, w);
    return ((double)(w[0] >> 5) * 67108864.0 + (double)(w[1] >> 6)) / 9007199254740992.0;
}

/* ------------------------------------------------------------------ event log */
static void log_ev(OracleEnv* e, int type, int a, int b, int c, int d) {
    int v[5] = {type, a, b, c, d};
    if (e->h->n_events < e->L.event_cap) {
        AgarEvent* ev = &e->ev[e->h->n_events];
        ev->type = type, ev->a = a, ev->b = b, ev->c = c, ev->d = d;
    } /* a full ring just stops recording: n_events keeps counting and the running hash covers everything */
    e->h->n_events += 1;
    /* COLLIDE is logged but not hashed: the re-test of a pair that was just pushed apart to distance == r1 + r2
     * is decided by the last bit of the positions, and it changes no discrete state */
    if (type == AGAR_EV_COLLIDE) return;
    uint64_t hh = e->h->event_hash;
    for (int i = 0; i < 5; ++i) hh = (hh ^ (uint64_t)(uint32_t)v[i]) * 0x100000001B3ULL;
    e->h->event_hash = hh;
}

/* ------------------------------------------------------------------ world hash (spatialHashTable.py:70-83)
 * An object lives in every bucket of the rectangle getIdsForArea(pos, radius) returns; two objects are
 * "nearby" iff their rectangles share a bucket.  Rect = inclusive column / row ranges, empty if x1 < x0. */
typedef struct Rect { int x0, x1, y0, y1; } Rect;
static void axis_range(double p, double radius, int S, int* b0, int* b1) {
    double cl = py_max0(p - radius);
    int bucket_left = (int)(cl - fmod(cl, (double)BUCKET));
    int limit = (int)py_minS((double)S, p + radius + 1);
    *b0 = bucket_left / BUCKET;
    *b1 = limit > bucket_left ? *b0 + (limit - bucket_left - 1) / BUCKET : *b0 - 1;
}
static Rect rect_of(const OracleEnv* e, double x, double y, double radius) {
    Rect r;
    axis_range(x, radius, e->S, &r.x0, &r.x1);
    axis_range(y, radius, e->S, &r.y0, &r.y1);
    return r;
}
static int rect_hit(Rect a, Rect b) {
    if (a.x1 < a.x0 || a.y1 < a.y0 || b.x1 < b.x0 || b.y1 < b.y0) return 0;
    return a.x0 <= b.x1 && b.x0 <= a.x1 && a.y0 <= b.y1 && b.y0 <= a.y1;
}
static inline double pellet_radius(int m) { return sqrt((double)m / M_PI); }

/* cell.py:143-152 overlap(self=a, cell=b): the bigger one by mass (tie -> b) must cover the other's centre */
static int overlap(double ax, double ay, double am, double ar, double bx, double by, double bm, double br) {
    double bigx, bigy, bigr, smx, smy;
    if (am > bm)
        bigx = ax, bigy = ay, bigr = ar, smx = bx, smy = by;
    else
        bigx = bx, bigy = by, bigr = br, smx = ax, smy = ay;
    double d2 = (bigx - smx) * (bigx - smx) + (bigy - smy) * (bigy - smy);
    return d2 * 1.1 < bigr * bigr;
}
/* cell.py:119-121 grow + field.py:14-17 adjustCellSize (the hash delete / insert is implicit in rect_of) */
static void grow(AgarCell* c, double food) {
    double nm = c->mass + food;
    if (!(nm < MAX_MASS)) nm = MAX_MASS; /* min(MAX, mass + food) */
    c->mass = nm;
    c->radius = radius_of(nm);
}
static void grow_mote(AgarMote* c, double food) {
    double nm = c->mass + food;
    if (!(nm < MAX_MASS)) nm = MAX_MASS;
    c->mass = nm;
    c->radius = radius_of(nm);
}
/* cell.py:154-155 */
static double merge_time_for(double factor, double mass) {
    return factor * (BASE_MERGE_TIME + mass * MERGE_TIME_MASS_FACTOR) * FPS / 2 / 1;
}

/* ------------------------------------------------------------------ player helpers (player.py:129-167) */
static double total_mass(const OracleEnv* e, int k) {
    double m[AGAR_MAX_CELLS];
    int n = e->pl[k].n_cells;
    if (n == 0) return 0.0;
    for (int i = 0; i < n; ++i) m[i] = CELLP(e, k, i)->mass;
    return np_sum(m, n);
}
static void update_fov_pos(OracleEnv* e, int k) { /* getFovPos :156-161 */
    AgarPlayer* p = &e->pl[k];
    if (!p->alive) return;
    double tm = total_mass(e, k);
    if (tm == 0) return;
    double ax[AGAR_MAX_CELLS], ay[AGAR_MAX_CELLS];
    for (int i = 0; i < p->n_cells; ++i) {
        const AgarCell* c = CELLP(e, k, i);
        ax[i] = c->x * c->mass;
        ay[i] = c->y * c->mass;
    }
    p->fov_x = np_sum(ax, p->n_cells) / tm;
    p->fov_y = np_sum(ay, p->n_cells) / tm;
    p->fov_valid = 1;
}
static void update_fov_size(OracleEnv* e, int k) { /* getFovSize :163-167 */
    AgarPlayer* p = &e->pl[k];
    if (!p->alive) return;
    double rmax = CELLP(e, k, 0)->radius;
    for (int i = 1; i < p->n_cells; ++i)
        if (CELLP(e, k, i)->radius > rmax) rmax = CELLP(e, k, i)->radius;
#ifdef AGAR_PORTABLE_MATH
    double pn = e->pow_n[p->n_cells];
#else
    double pn = pow((double)p->n_cells, 0.32);
#endif
    p->fov_size = m_pow(rmax, 0.475) * pn * 35;
}
static void cell_remove(OracleEnv* e, int k, int i) {
    AgarPlayer* p = &e->pl[k];
    for (int j = i; j + 1 < p->n_cells; ++j) *CELLP(e, k, j) = *CELLP(e, k, j + 1);
    p->n_cells -= 1;
    memset(CELLP(e, k, p->n_cells), 0, sizeof(AgarCell));
}
static AgarCell* cell_append(OracleEnv* e, int k, double x, double y, double mass) { /* Cell(x, y, mass, player) */
    AgarPlayer* p = &e->pl[k];
    AgarCell* c = CELLP(e, k, p->n_cells);
    memset(c, 0, sizeof *c);
    c->x = x, c->y = y, c->mass = mass, c->radius = radius_of(mass);
    c->uid = e->h->next_uid++;
    p->n_cells += 1;
    return c;
}
/* field.py:382-388 deletePlayerCell */
static void delete_player_cell(OracleEnv* e, int k, int i) {
    cell_remove(e, k, i);
    AgarPlayer* p = &e->pl[k];
    if (p->n_cells == 0) {
        e->h->dead_order[e->h->n_dead++] = k;
        p->alive = 0;
        p->respawn_time = 1; /* player.py:6,110-112 */
        log_ev(e, AGAR_EV_PLAYER_DIED, k, 0, 0, 0);
        p->bot.stat_deaths += 1;
    }
}

/* ------------------------------------------------------------------ momentum / movement (cell.py:96-141) */
static void add_momentum(OracleEnv* e, double x, double y, double px, double py, double orig_radius, double* svx,
                         double* svy, int32_t* counter) {
    double cx = py_max0(py_minS((double)e->S, px)), cy = py_max0(py_minS((double)e->S, py));
    double c, s;
    dir_of(cy - y, cx - x, &c, &s);
    double speed = 2 + orig_radius * 0.05;
    *svx = c * speed;
    *svy = s * speed;
    *counter = 15;
}
static void update_momentum(double* svx, double* svy, int32_t* counter) { /* :105-116 */
    if (*counter == -1) return;
    if (*counter > 0) {
        *counter -= 1;
        double ratio = (double)*counter / 15;
        if (ratio < 0.1) {
            *svx *= (1 - ratio);
            *svy *= (1 - ratio);
        }
    } else {
        *svx = 0, *svy = 0;
        *counter = -1;
    }
}
static void update_pos(double* x, double* y, double vx, double vy, double* svx, double* svy, int counter, double S) {
    double xs = vx + *svx, ys = vy + *svy; /* :132-141 */
    *x = clampS(*x + xs, S);
    *y = clampS(*y + ys, S);
    if ((counter && *x == S) || *x == 0) *svx *= -1;
    if ((counter && *y == S) || *y == 0) *svy *= -1;
}

/* ------------------------------------------------------------------ spawning (field.py:262-313) */
static int player_bucket_occupied(const OracleEnv* e, int bx, int by) {
    for (int k = 0; k < e->L.n_players; ++k)
        for (int i = 0; i < e->pl[k].n_cells; ++i) {
            const AgarCell* c = CELLP(e, k, i);
            if (!(c->flags & AGAR_CF_INHASH)) continue;
            Rect r = rect_of(e, c->x, c->y, c->radius);
            if (r.x1 < r.x0 || r.y1 < r.y0) continue;
            if (bx >= r.x0 && bx <= r.x1 && by >= r.y0 && by <= r.y1) return 1;
        }
    return 0;
}
static void get_spawn_pos(OracleEnv* e, double radius, double* ox, double* oy) { /* :283-301 */
    int cols = e->nb, total = cols * cols;
    int b = (int)draw_randint(e, 0, 0, total), count = 0;
    while (count < total && player_bucket_occupied(e, b % cols, b / cols)) {
        b = (b + 1) % total;
        count++;
    }
    /* the reference tests `buckets[b] and count < total` — same loop, bucket test first; equivalent */
    if (count == total) {
        *ox = (double)draw_randint(e, 0, 0, e->S);
        *oy = (double)draw_randint(e, 0, 0, e->S);
    } else {
        int x = b % cols;
        double y = (double)(b - x) / cols;
        double left = (double)((x - 1) * BUCKET), top = y * BUCKET;
        *ox = (double)draw_randint(e, 0, left + radius, left + BUCKET - radius);
        *oy = (double)draw_randint(e, 0, top + radius, top + BUCKET - radius);
    }
}
static void initialize_player(OracleEnv* e, int k) { /* :49-55 */
    AgarPlayer* p = &e->pl[k];
    for (int i = 0; i < e->L.cell_cap; ++i) memset(CELLP(e, k, i), 0, sizeof(AgarCell));
    p->n_cells = 0;
    double x, y;
    get_spawn_pos(e, e->start_radius, &x, &y);
    AgarCell* c = cell_append(e, k, x, y, START_MASS);
    p->alive = 1;
    p->respawn_time = 0;
    log_ev(e, AGAR_EV_SPAWN_PLAYER, k, (int)c->uid, (int)x, (int)y);
}
static void spawn_pellets(OracleEnv* e) { /* :303-313, :20-26 */
    while ((double)(e->h->n_pellets + e->h->n_fat) < e->L.max_pellets) {
        int x = (int)draw_randint(e, 0, 0, e->S), y = (int)draw_randint(e, 0, 0, e->S);
        int v = (int)draw_randint(e, 0, 0, 50);
        int m = v > 46 ? 50 - v : 1;
        int slot = 0;
        while (slot < e->L.pellet_cap && e->pel[slot]) ++slot;
        if (slot == e->L.pellet_cap) abort(); /* cannot happen: n_pellets < max <= cap */
        log_ev(e, AGAR_EV_SPAWN_PELLET, slot, x, y, m);
        e->pel[slot] = AGAR_PELLET_PACK(x, y, m);
        e->h->n_pellets += 1;
    }
}
static void spawn_viruses(OracleEnv* e) { /* :262-275 */
    while ((double)e->h->n_viruses < e->L.max_viruses) {
        if (e->h->n_viruses >= e->L.virus_cap) {
            e->h->overflow |= AGAR_OVF_VIRUS;
            break;
        }
        double x, y;
        get_spawn_pos(e, e->virus_radius, &x, &y);
        double acc = BUCKET - e->virus_radius;
        x += (double)draw_randint(e, 0, (-1) * acc / 2, acc / 2);
        y += (double)draw_randint(e, 0, (-1) * acc / 2, acc / 2);
        AgarMote* v = &e->vir[e->h->n_viruses];
        memset(v, 0, sizeof *v);
        v->x = x, v->y = y, v->mass = VIRUS_BASE_SIZE, v->radius = radius_of(VIRUS_BASE_SIZE);
        log_ev(e, AGAR_EV_SPAWN_VIRUS, e->h->n_viruses, (int)x, (int)y, 0);
        e->h->n_viruses += 1;
    }
}
static void spawn_players(OracleEnv* e) { /* :277-281 */
    int n = e->h->n_dead, w = 0;
    int order[AGAR_MAX_PLAYERS];
    memcpy(order, e->h->dead_order, sizeof order);
    for (int i = 0; i < n; ++i) {
        int k = order[i];
        if (e->pl[k].respawn_time == 0)
            initialize_player(e, k);
        else
            e->h->dead_order[w++] = k;
    }
    e->h->n_dead = w;
    for (int i = w; i < AGAR_MAX_PLAYERS; ++i) e->h->dead_order[i] = 0;
}
static void spawn_stuff(OracleEnv* e) { /* :256-260 */
    spawn_pellets(e);
    if (e->cfg.virus_enabled) spawn_viruses(e);
    spawn_players(e);
}

/* ------------------------------------------------------------------ Field.update phases */
static void update_viruses(OracleEnv* e) { /* :94-97 */
    for (int i = 0; i < e->h->n_viruses; ++i) {
        AgarMote* v = &e->vir[i];
        update_momentum(&v->svx, &v->svy, &v->counter);
        update_pos(&v->x, &v->y, 0, 0, &v->svx, &v->svy, v->counter, (double)e->S);
    }
}
static void blob_remove(OracleEnv* e, int i) {
    for (int j = i; j + 1 < e->h->n_blobs; ++j) e->blob[j] = e->blob[j + 1];
    e->h->n_blobs -= 1;
    memset(&e->blob[e->h->n_blobs], 0, sizeof(AgarMote));
}
static void virus_remove(OracleEnv* e, int i) {
    for (int j = i; j + 1 < e->h->n_viruses; ++j) e->vir[j] = e->vir[j + 1];
    e->h->n_viruses -= 1;
    memset(&e->vir[e->h->n_viruses], 0, sizeof(AgarMote));
}
static void update_blobs(OracleEnv* e) { /* :99-110 */
    int still[1024], ns = 0;
    for (int i = 0; i < e->h->n_blobs; ++i) {
        AgarMote* b = &e->blob[i];
        if (b->counter == 0) {
            still[ns++] = i;
            continue;
        }
        update_momentum(&b->svx, &b->svy, &b->counter);
        update_pos(&b->x, &b->y, 0, 0, &b->svx, &b->svy, b->counter, (double)e->S);
    }
    for (int j = 0; j < ns; ++j) {
        int i = still[j] - j; /* earlier removals shifted the list */
        AgarMote b = e->blob[i];
        blob_remove(e, i);
        int slot = 0;
        while (slot < e->L.fat_cap && e->fat[slot].mass != 0) ++slot;
        if (slot == e->L.fat_cap) {
            e->h->overflow |= AGAR_OVF_FAT; /* pool full: the blob is dropped (reported, never UB) */
            continue;
        }
        e->fat[slot].x = b.x, e->fat[slot].y = b.y, e->fat[slot].mass = b.mass, e->fat[slot].radius = b.radius;
        e->h->n_fat += 1;
        log_ev(e, AGAR_EV_BLOB_TO_PELLET, slot, 0, 0, 0);
    }
}
static void player_update(OracleEnv* e, int k) { /* player.py:30-36 */
    AgarPlayer* p = &e->pl[k];
    double S = (double)e->S;
    double vx[AGAR_MAX_CELLS], vy[AGAR_MAX_CELLS];
    memset(vx, 0, sizeof vx);
    memset(vy, 0, sizeof vy);
    for (int i = 0; i < p->n_cells; ++i) { /* decayMass, cell.py:123-126 */
        AgarCell* c = CELLP(e, k, i);
        if (c->mass >= 4) {
            c->mass = c->mass * e->decay_rate;
            c->radius = radius_of(c->mass);
        }
    }
    for (int i = 0; i < p->n_cells; ++i) { /* updateCellProperties */
        AgarCell* c = CELLP(e, k, i);
        update_momentum(&c->svx, &c->svy, &c->counter);
        if (c->merge_time > 0) c->merge_time -= 1;
        /* setMoveDirection, cell.py:47-57 */
        double xd = p->cmd_x - c->x, yd = p->cmd_y - c->y;
        double h2 = xd * xd + yd * yd, r2 = c->radius * c->radius;
        double sm = (h2 < r2 ? h2 : r2) / r2; /* min(h2, r2) / r2 */
        double cs, sn;
        dir_of(yd, xd, &cs, &sn);
        double rs = e->move_speed * m_pow(c->mass, -0.35);
        vx[i] = rs * sm * cs;
        vy[i] = rs * sm * sn;
    }
    if (p->do_split) { /* player.py:53-61 */
        /* stable sort by mass, descending (persistently reorders the list) */
        for (int i = 1; i < p->n_cells; ++i) {
            AgarCell t = *CELLP(e, k, i);
            double tvx = vx[i], tvy = vy[i];
            int j = i - 1;
            while (j >= 0 && CELLP(e, k, j)->mass < t.mass) {
                *CELLP(e, k, j + 1) = *CELLP(e, k, j);
                vx[j + 1] = vx[j], vy[j + 1] = vy[j];
                --j;
            }
            *CELLP(e, k, j + 1) = t;
            vx[j + 1] = tvx, vy[j + 1] = tvy;
        }
        int n0 = p->n_cells;
        for (int i = 0; i < n0; ++i) {
            AgarCell* c = CELLP(e, k, i);
            if (c->mass > 36 && p->n_cells < 16) { /* cell.py:72-85 */
                double parent_radius = c->radius;
                AgarCell* nc = cell_append(e, k, c->x, c->y, c->mass / 2);
                c = CELLP(e, k, i);
                double cs, sn;
                dir_of(p->cmd_y - nc->y, p->cmd_x - nc->x, &cs, &sn);
                double xp = cs * nc->radius * 4.5 + c->x, yp = sn * nc->radius * 4.5 + c->y;
                add_momentum(e, nc->x, nc->y, xp, yp, parent_radius, &nc->svx, &nc->svy, &nc->counter);
                nc->merge_time = merge_time_for(1, nc->mass);
                c->mass = c->mass / 2;
                c->radius = radius_of(c->mass);
                log_ev(e, AGAR_EV_SPLIT, k, (int)c->uid, (int)nc->uid, 0);
            }
        }
    }
    if (p->do_eject) /* player.py:63-68 */
        for (int i = 0; i < p->n_cells; ++i)
            if (CELLP(e, k, i)->mass >= 35) CELLP(e, k, i)->flags |= AGAR_CF_EJECT;
    for (int i = 0; i < p->n_cells; ++i) { /* updateCellsMovement */
        AgarCell* c = CELLP(e, k, i);
        update_pos(&c->x, &c->y, vx[i], vy[i], &c->svx, &c->svy, c->counter, S);
    }
}
static void perform_ejections(OracleEnv* e, int k) { /* field.py:134-146, cell.py:90-94 */
    AgarPlayer* p = &e->pl[k];
    for (int i = 0; i < p->n_cells; ++i) {
        AgarCell* c = CELLP(e, k, i);
        if (!(c->flags & AGAR_CF_EJECT)) continue;
        c->mass -= EJECT_BASE_MASS; /* radius deliberately left stale */
        c->flags &= ~AGAR_CF_EJECT;
        if (e->h->n_blobs >= e->L.blob_cap) {
            e->h->overflow |= AGAR_OVF_BLOB;
            continue;
        }
        AgarMote* b = &e->blob[e->h->n_blobs];
        memset(b, 0, sizeof *b);
        b->x = c->x, b->y = c->y, b->mass = e->blob_mass, b->radius = radius_of(e->blob_mass);
        add_momentum(e, b->x, b->y, p->cmd_x, p->cmd_y, c->radius, &b->svx, &b->svy, &b->counter);
        b->aux = c->uid;
        log_ev(e, AGAR_EV_EJECT, k, (int)c->uid, e->h->n_blobs, 0);
        e->h->n_blobs += 1;
    }
}
static void handle_player_collisions(OracleEnv* e, int k) { /* field.py:149-181 */
    AgarPlayer* p = &e->pl[k];
    double S = (double)e->S;
    for (int i = 0; i < p->n_cells; ++i) {
        AgarCell* a = CELLP(e, k, i);
        if (a->counter > 0) continue;
        for (int j = 0; j < p->n_cells; ++j) {
            AgarCell* b = CELLP(e, k, j);
            if (i == j || b->counter > 0 || (a->merge_time <= 0 && b->merge_time <= 0)) continue;
            double d2 = (a->x - b->x) * (a->x - b->x) + (a->y - b->y) * (a->y - b->y);
            double dist = sqrt(d2), sum = a->radius + b->radius;
            if (dist < sum && dist != 0) {
                log_ev(e, AGAR_EV_COLLIDE, k, (int)a->uid, (int)b->uid, 0);
                AgarCell *big, *sm;
                if (a->mass > b->mass)
                    big = a, sm = b;
                else
                    big = b, sm = a;
                double ds = (sum - dist) / dist, q = sm->mass / big->mass;
                double xs = (big->x - sm->x) * ds, ys = (big->y - sm->y) * ds;
                double nbx = big->x + xs * q, nby = big->y + ys * q;
                double nsx = sm->x - xs * (1 - q), nsy = sm->y - ys * (1 - q);
                big->x = clampS(nbx, S), big->y = clampS(nby, S);
                sm->x = clampS(nsx, S), sm->y = clampS(nsy, S);
            }
        }
    }
}
static void update_players(OracleEnv* e) { /* :112-119 */
    for (int k = 0; k < e->L.n_players; ++k) {
        if (e->pl[k].alive) {
            player_update(e, k);
            perform_ejections(e, k);
            handle_player_collisions(e, k);
        } else
            e->pl[k].respawn_time -= 1;
    }
}
static void update_hash_tables(OracleEnv* e) { /* :121-132 */
    for (int k = 0; k < e->L.n_players; ++k)
        for (int i = 0; i < e->pl[k].n_cells; ++i) CELLP(e, k, i)->flags |= AGAR_CF_INHASH;
    for (int i = 0; i < e->h->n_viruses; ++i) e->vir[i].aux = AGAR_CF_INHASH;
}
static void merge_player_cells(OracleEnv* e) { /* :183-198, :372-380 */
    for (int k = 0; k < e->L.n_players; ++k) {
        AgarPlayer* p = &e->pl[k];
        if (!p->alive) continue;
        uint32_t uid[AGAR_MAX_CELLS];
        int n = 0;
        for (int i = 0; i < p->n_cells; ++i)
            if (CELLP(e, k, i)->merge_time <= 0) uid[n++] = CELLP(e, k, i)->uid;
        if (n <= 1) continue;
        /* stable sort of the mergeable list by mass, descending */
        double mass[AGAR_MAX_CELLS];
        for (int a = 0; a < n; ++a)
            for (int i = 0; i < p->n_cells; ++i)
                if (CELLP(e, k, i)->uid == uid[a]) mass[a] = CELLP(e, k, i)->mass;
        for (int a = 1; a < n; ++a) {
            uint32_t tu = uid[a];
            double tm = mass[a];
            int j = a - 1;
            while (j >= 0 && mass[j] < tm) {
                uid[j + 1] = uid[j], mass[j + 1] = mass[j];
                --j;
            }
            uid[j + 1] = tu, mass[j + 1] = tm;
        }
        int alive[AGAR_MAX_CELLS];
        for (int a = 0; a < n; ++a) alive[a] = 1;
#define FIND(u, out)                                  \
    do {                                              \
        out = -1;                                     \
        for (int _i = 0; _i < p->n_cells; ++_i)       \
            if (CELLP(e, k, _i)->uid == (u)) out = _i; \
    } while (0)
        for (int a = 0; a < n; ++a) {
            if (!alive[a]) continue;
            for (int b = 0; b < n; ++b) {
                if (!alive[b] || b == a) continue;
                int ia, ib;
                FIND(uid[a], ia);
                FIND(uid[b], ib);
                AgarCell *c1 = CELLP(e, k, ia), *c2 = CELLP(e, k, ib);
                if (overlap(c1->x, c1->y, c1->mass, c1->radius, c2->x, c2->y, c2->mass, c2->radius)) {
                    int big = c1->mass > c2->mass ? a : b, sm = big == a ? b : a;
                    AgarCell *cb = big == a ? c1 : c2, *cs = big == a ? c2 : c1;
                    log_ev(e, AGAR_EV_MERGE, k, (int)cb->uid, (int)cs->uid, 0);
                    grow(cb, cs->mass);
                    alive[sm] = 0;
                    delete_player_cell(e, k, sm == a ? ia : ib);
                    if (!alive[a]) break;
                }
            }
        }
#undef FIND
    }
}

static void virus_blob_overlap(OracleEnv* e) { /* :246-253, :316-325 */
    for (int vi = 0; vi < e->h->n_viruses; ++vi) { /* the list may grow while iterated */
        AgarMote* v = &e->vir[vi];
        Rect rv = rect_of(e, v->x, v->y, v->radius);
        int cand[1024], nc = 0;
        for (int b = 0; b < e->h->n_blobs; ++b)
            if (rect_hit(rv, rect_of(e, e->blob[b].x, e->blob[b].y, e->blob[b].radius))) cand[nc++] = b;
        for (int c = 0; c < nc; ++c) {
            int b = cand[c];
            AgarMote* bl = &e->blob[b];
            if (!overlap(v->x, v->y, v->mass, v->radius, bl->x, bl->y, bl->mass, bl->radius)) continue;
            double bx = bl->x, by = bl->y;
            grow_mote(v, bl->mass);
            blob_remove(e, b);
            for (int c2 = c + 1; c2 < nc; ++c2) cand[c2] -= 1; /* later candidates sit after b */
            int split = 0;
            if (v->mass >= e->virus_split_mass) {
                if (e->h->n_viruses >= e->L.virus_cap)
                    e->h->overflow |= AGAR_OVF_VIRUS;
                else {
                    /* Cell.split on a virus: cell.py:72-85 */
                    double ox = 2 * v->x - bx, oy = 2 * v->y - by;
                    AgarMote* nv = &e->vir[e->h->n_viruses];
                    memset(nv, 0, sizeof *nv);
                    nv->x = v->x, nv->y = v->y, nv->mass = v->mass / 2, nv->radius = radius_of(nv->mass);
                    double cs, sn;
                    dir_of(oy - nv->y, ox - nv->x, &cs, &sn);
                    double xp = cs * nv->radius * 4.5 + v->x, yp = sn * nv->radius * 4.5 + v->y;
                    add_momentum(e, nv->x, nv->y, xp, yp, v->radius, &nv->svx, &nv->svy, &nv->counter);
                    v->mass = v->mass / 2;
                    v->radius = radius_of(v->mass);
                    e->h->n_viruses += 1;
                    split = 1;
                }
            }
            log_ev(e, AGAR_EV_VIRUS_EAT_BLOB, vi, b, split, 0);
        }
    }
}
static void player_cell_ate_virus(OracleEnv* e, int k, int ci) { /* :350-370 */
    AgarPlayer* p = &e->pl[k];
    int n_new = 16 - p->n_cells;
    if (n_new == 0) return;
    AgarCell* c = CELLP(e, k, ci);
    double distributed = c->mass * VIRUS_EXPLOSION_PROPORTION;
    double per = distributed / n_new;
    c->merge_time = merge_time_for(MERGE_TIME_VIRUS_FACTOR, c->mass);
    grow(c, -1 * per * n_new);
    for (int j = 0; j < n_new; ++j) {
        AgarCell* nc = cell_append(e, k, c->x, c->y, per);
        int deg = (int)draw_randint(e, 0, 0, 360);
        double cs = e->cos_deg[deg], sn = e->sin_deg[deg];
        double xp = cs * c->radius * 12 + c->x, yp = sn * c->radius * 12 + c->y;
        add_momentum(e, nc->x, nc->y, xp, yp, c->radius, &nc->svx, &nc->svy, &nc->counter);
        nc->merge_time = merge_time_for(0.8, nc->mass);
        nc->flags |= AGAR_CF_INHASH; /* addPlayerCell inserts into the player table */
    }
}
static void player_virus_overlap(OracleEnv* e) { /* :225-231, :333-335 */
    for (int k = 0; k < e->L.n_players; ++k) {
        AgarPlayer* p = &e->pl[k];
        if (!p->alive) continue;
        for (int ci = 0; ci < p->n_cells; ++ci) { /* exploded cells are appended and visited */
            AgarCell* c = CELLP(e, k, ci);
            Rect rc = rect_of(e, c->x, c->y, c->radius);
            int cand[256], nc = 0;
            for (int v = 0; v < e->h->n_viruses; ++v)
                if ((e->vir[v].aux & AGAR_CF_INHASH) &&
                    rect_hit(rc, rect_of(e, e->vir[v].x, e->vir[v].y, e->vir[v].radius)))
                    cand[nc++] = v;
            for (int q = 0; q < nc; ++q) {
                AgarMote* v = &e->vir[cand[q]];
                if (overlap(c->x, c->y, c->mass, c->radius, v->x, v->y, v->mass, v->radius) && c->mass > 1.25 * v->mass) {
                    log_ev(e, AGAR_EV_EAT_VIRUS, k, (int)c->uid, cand[q], 16 - p->n_cells);
                    grow(c, v->mass * VIRUS_EAT_FACTOR);
                    virus_remove(e, cand[q]);
                    for (int q2 = q + 1; q2 < nc; ++q2) cand[q2] -= 1;
                    player_cell_ate_virus(e, k, ci);
                }
            }
        }
    }
}
static void player_pellet_overlap(OracleEnv* e) { /* :207-213, :327-344 */
    for (int k = 0; k < e->L.n_players; ++k) {
        AgarPlayer* p = &e->pl[k];
        if (!p->alive) continue;
        for (int ci = 0; ci < p->n_cells; ++ci) {
            AgarCell* c = CELLP(e, k, ci);
            Rect rc = rect_of(e, c->x, c->y, c->radius); /* candidates are fixed before the cell grows */
            for (int s = 0; s < e->L.pellet_cap; ++s) {
                uint32_t pk = e->pel[s];
                if (!pk) continue;
                int px = AGAR_PELLET_X(pk), py = AGAR_PELLET_Y(pk), pm = AGAR_PELLET_M(pk);
                double pr = pellet_radius(pm);
                if (!rect_hit(rc, rect_of(e, px, py, pr))) continue;
                if (overlap(c->x, c->y, c->mass, c->radius, px, py, pm, pr) && c->mass > 1.25 * pm) {
                    log_ev(e, AGAR_EV_EAT_PELLET, k, (int)c->uid, s, 0);
                    grow(c, pm);
                    e->pel[s] = 0;
                    e->h->n_pellets -= 1;
                }
            }
            for (int s = 0; s < e->L.fat_cap; ++s) {
                AgarFatPellet* f = &e->fat[s];
                if (f->mass == 0) continue;
                if (!rect_hit(rc, rect_of(e, f->x, f->y, f->radius))) continue;
                if (overlap(c->x, c->y, c->mass, c->radius, f->x, f->y, f->mass, f->radius) && c->mass > 1.25 * f->mass) {
                    log_ev(e, AGAR_EV_EAT_PELLET, k, (int)c->uid, s | 0x10000, 0);
                    grow(c, f->mass);
                    memset(f, 0, sizeof *f);
                    e->h->n_fat -= 1;
                }
            }
        }
    }
}
static void player_blob_overlap(OracleEnv* e) { /* :215-222 */
    for (int k = 0; k < e->L.n_players; ++k) {
        AgarPlayer* p = &e->pl[k];
        if (!p->alive) continue;
        for (int ci = 0; ci < p->n_cells; ++ci) {
            AgarCell* c = CELLP(e, k, ci);
            Rect rc = rect_of(e, c->x, c->y, c->radius);
            int cand[1024], nc = 0;
            for (int b = 0; b < e->h->n_blobs; ++b)
                if (rect_hit(rc, rect_of(e, e->blob[b].x, e->blob[b].y, e->blob[b].radius))) cand[nc++] = b;
            for (int q = 0; q < nc; ++q) {
                AgarMote* b = &e->blob[cand[q]];
                if (overlap(c->x, c->y, c->mass, c->radius, b->x, b->y, b->mass, b->radius) && b->aux != c->uid &&
                    c->mass > 1.25 * b->mass) {
                    log_ev(e, AGAR_EV_EAT_BLOB, k, (int)c->uid, cand[q], (int)b->aux);
                    grow(c, b->mass);
                    blob_remove(e, cand[q]);
                    for (int q2 = q + 1; q2 < nc; ++q2) cand[q2] -= 1;
                }
            }
        }
    }
}
static void player_player_overlap(OracleEnv* e) { /* :233-244, :346-348 */
    for (int k = 0; k < e->L.n_players; ++k) {
        AgarPlayer* p = &e->pl[k];
        if (!p->alive) continue;
        for (int ci = 0; ci < p->n_cells; ++ci) { /* Python list iterator: a removed current cell skips the next */
            AgarCell* c = CELLP(e, k, ci);
            uint32_t my_uid = c->uid;
            Rect rc = rect_of(e, c->x, c->y, c->radius);
            uint32_t cand_uid[AGAR_MAX_PLAYERS * AGAR_MAX_CELLS];
            int cand_k[AGAR_MAX_PLAYERS * AGAR_MAX_CELLS], nc = 0;
            for (int k2 = 0; k2 < e->L.n_players; ++k2) {
                if (k2 == k) continue;
                for (int j = 0; j < e->pl[k2].n_cells; ++j) {
                    AgarCell* o = CELLP(e, k2, j);
                    if ((o->flags & AGAR_CF_INHASH) && rect_hit(rc, rect_of(e, o->x, o->y, o->radius))) {
                        cand_uid[nc] = o->uid;
                        cand_k[nc++] = k2;
                    }
                }
            }
            for (int q = 0; q < nc; ++q) {
                int k2 = cand_k[q], j = -1;
                for (int t = 0; t < e->pl[k2].n_cells; ++t)
                    if (CELLP(e, k2, t)->uid == cand_uid[q]) j = t;
                if (j < 0) abort(); /* candidates of one cell are only removed by that cell */
                AgarCell* o = CELLP(e, k2, j);
                c = CELLP(e, k, ci);
                if (!overlap(c->x, c->y, c->mass, c->radius, o->x, o->y, o->mass, o->radius)) continue;
                if (c->mass > 1.25 * o->mass) {
                    log_ev(e, AGAR_EV_EAT_CELL, k, (int)c->uid, k2, (int)o->uid);
                    grow(c, o->mass);
                    delete_player_cell(e, k2, j);
                } else if (o->mass > 1.25 * c->mass) {
                    log_ev(e, AGAR_EV_EAT_CELL, k2, (int)o->uid, k, (int)my_uid);
                    grow(o, c->mass);
                    delete_player_cell(e, k, ci);
                    break;
                }
            }
        }
    }
}
static void field_update(OracleEnv* e) { /* field.py:85-92 */
    update_viruses(e);
    update_blobs(e);
    update_players(e);
    update_hash_tables(e);
    merge_player_cells(e);
    virus_blob_overlap(e);
    player_virus_overlap(e);
    player_pellet_overlap(e);
    player_blob_overlap(e);
    player_player_overlap(e);
    spawn_stuff(e);
}

/* ------------------------------------------------------------------ field-of-view queries (field.py:414-456) */
static int in_fov(double x, double y, double r, double fx, double fy, double fov) { /* cell.py:169-177 */
    double h = fov / 2;
    double xmin = fx - h, xmax = fx + h, ymin = fy - h, ymax = fy + h;
    if (x + r < xmin || x - r > xmax || y + r < ymin || y - r > ymax) return 0;
    return 1;
}

/* ------------------------------------------------------------------ grid vision (bot.py:326-497, spatialHashTable.py:85-112) */
#define AGAR_GRID_NB (85 * 85) /* (G + 1)^2 buckets, G <= 84 (agar_layout.h) */
typedef struct GridTables {
    int cols, canonical;
    double pel_sum[AGAR_GRID_NB];
    double own_max[AGAR_GRID_NB], enemy_max[AGAR_GRID_NB], vir_best_r[AGAR_GRID_NB], vir_mass[AGAR_GRID_NB];
    uint8_t pel_has[AGAR_GRID_NB], own_has[AGAR_GRID_NB], enemy_has[AGAR_GRID_NB], vir_has[AGAR_GRID_NB];
    int stamp[AGAR_GRID_NB];
    int serial;
} GridTables;
enum { T_PELLET, T_OWN, T_ENEMY, T_VIRUS };
/* mathematically exact floor(v / gs) for v >= 0 (AGAR_OBS_CANONICAL binning) */
static int exact_floor_div(double v, double gs) {
    double q = floor(v / gs);
    double res = fma(-q, gs, v);
    if (res < 0)
        q -= 1;
    else if (res >= gs)
        q += 1;
    return (int)q;
}
static void grid_insert(GridTables* g, int table, double ox, double oy, double radius, double mass, double left,
                        double top, double fov, double gs) {
    /* getIdsForAreaFloatingPoint */
    double px = ox - left, py = oy - top;
    double cl = py_max0(px - radius), ct = py_max0(py - radius);
    double bl = cl - fmod(cl, gs), bt = ct - fmod(ct, gs);
    double lx = (px + radius < fov - 1) ? px + radius : fov - 1; /* min(size - 1, pos + radius) */
    double ly = (py + radius < fov - 1) ? py + radius : fov - 1;
    g->serial += 1;
    int canon = g->canonical;
    int cx0 = 0, cx1 = -1, cy0 = 0, cy1 = -1;
    if (canon) { /* robust mode: the buckets floor(lo / gs) .. floor(hi / gs), exact floors */
        cx0 = exact_floor_div(cl, gs), cx1 = lx >= 0 ? exact_floor_div(lx, gs) : -1;
        cy0 = exact_floor_div(ct, gs), cy1 = ly >= 0 ? exact_floor_div(ly, gs) : -1;
        bl = bt = 0; /* the float loops below run once; ids come from the integer ranges */
        lx = ly = 0;
    }
    for (double x = bl; x <= lx; x += gs)
        for (double y = bt; y <= ly; y += gs)
          for (int bx = canon ? cx0 : 0; bx <= (canon ? cx1 : 0); ++bx)
            for (int by = canon ? cy0 : 0; by <= (canon ? cy1 : 0); ++by) {
            int id = canon ? bx + by * g->cols : (int)(x / gs) + (int)(y / gs) * g->cols;
            if (id < 0 || id >= g->cols * g->cols) abort(); /* KeyError in the reference */
            if (g->stamp[id] == g->serial) continue;       /* ids is a set */
            g->stamp[id] = g->serial;
            switch (table) {
            case T_PELLET:
                g->pel_sum[id] = g->pel_has[id] ? g->pel_sum[id] + mass : 0 + mass;
                g->pel_has[id] = 1;
                break;
            case T_OWN:
                if (!g->own_has[id] || mass > g->own_max[id]) g->own_max[id] = mass;
                g->own_has[id] = 1;
                break;
            case T_ENEMY:
                if (!g->enemy_has[id] || mass > g->enemy_max[id]) g->enemy_max[id] = mass;
                g->enemy_has[id] = 1;
                break;
            case T_VIRUS:
                if (!g->vir_has[id] || radius > g->vir_best_r[id]) g->vir_best_r[id] = radius, g->vir_mass[id] = mass;
                g->vir_has[id] = 1;
                break;
            }
        }
}
static void observe_agent(OracleEnv* e, int k, int agent, float* obs, double* obs64) {
    const AgarConfig* cf = &e->cfg;
    AgarPlayer* p = &e->pl[k];
    AgarBot* B = &p->bot;
    int G = e->L.grid_squares, GG = G * G;
    double S = (double)e->S;
    /* getGridStateRepresentation */
    update_fov_size(e, k);
    update_fov_pos(e, k);
    double fov = p->fov_size, fx = p->fov_x, fy = p->fov_y;
    double left = fx - fov / 2, top = fy - fov / 2;
    double gs = fov / G;
    static __thread GridTables g;
    {   /* clear what this observation can touch: (G + 1)^2 buckets */
        size_t nb = (size_t)(G + 1) * (G + 1);
        memset(g.pel_has, 0, nb), memset(g.own_has, 0, nb), memset(g.enemy_has, 0, nb), memset(g.vir_has, 0, nb);
        memset(g.stamp, 0, nb * sizeof(int));
        g.serial = 0;
    }
    g.cols = cf->obs_mode == AGAR_OBS_CANONICAL ? G : (int)ceil(fov / gs); /* spatialHashTable.py:19 */
    g.canonical = cf->obs_mode == AGAR_OBS_CANONICAL;
    Rect ra = rect_of(e, fx, fy, fov / 2);
    for (int s = 0; s < e->L.pellet_cap; ++s) { /* getPelletsInFov */
        uint32_t pk = e->pel[s];
        if (!pk) continue;
        int px = AGAR_PELLET_X(pk), py = AGAR_PELLET_Y(pk), pm = AGAR_PELLET_M(pk);
        double pr = pellet_radius(pm);
        if (!rect_hit(ra, rect_of(e, px, py, pr)) || !in_fov(px, py, pr, fx, fy, fov)) continue;
        grid_insert(&g, T_PELLET, px, py, pr, pm, left, top, fov, gs);
    }
    for (int s = 0; s < e->L.fat_cap; ++s) {
        const AgarFatPellet* f = &e->fat[s];
        if (f->mass == 0) continue;
        if (!rect_hit(ra, rect_of(e, f->x, f->y, f->radius)) || !in_fov(f->x, f->y, f->radius, fx, fy, fov)) continue;
        grid_insert(&g, T_PELLET, f->x, f->y, f->radius, f->mass, left, top, fov, gs);
    }
    double biggest = 0.0; /* bot.py:364-367: mass of the biggest player cell in view (own or enemy), NORMALIZE_GRID_BY_MAX_MASS */
    for (int k2 = 0; k2 < e->L.n_players; ++k2) /* getEnemyPlayerCellsInFov (through the player table) */
        for (int j = 0; j < e->pl[k2].n_cells; ++j) {
            const AgarCell* o = CELLP(e, k2, j);
            if (k2 == k) continue;
            if (!(o->flags & AGAR_CF_INHASH) || !rect_hit(ra, rect_of(e, o->x, o->y, o->radius)) ||
                !in_fov(o->x, o->y, o->radius, fx, fy, fov))
                continue;
            grid_insert(&g, T_ENEMY, o->x, o->y, o->radius, o->mass, left, top, fov, gs);
            if (o->mass > biggest) biggest = o->mass;
        }
    for (int j = 0; j < p->n_cells; ++j) { /* own cells: getPortionOfCellsInFov(player.getCells()) */
        const AgarCell* o = CELLP(e, k, j);
        if (!in_fov(o->x, o->y, o->radius, fx, fy, fov)) continue;
        grid_insert(&g, T_OWN, o->x, o->y, o->radius, o->mass, left, top, fov, gs);
        if (o->mass > biggest) biggest = o->mass;
    }
    if (cf->virus_enabled)
        for (int v = 0; v < e->h->n_viruses; ++v) {
            const AgarMote* o = &e->vir[v];
            if (!(o->aux & AGAR_CF_INHASH) || !rect_hit(ra, rect_of(e, o->x, o->y, o->radius)) ||
                !in_fov(o->x, o->y, o->radius, fx, fy, fov))
                continue;
            grid_insert(&g, T_VIRUS, o->x, o->y, o->radius, o->mass, left, top, fov, gs);
        }
    static __thread double pel[AGAR_GRID_NB], own[AGAR_GRID_NB], enemy[AGAR_GRID_NB], vir[AGAR_GRID_NB], wall[AGAR_GRID_NB];
    memset(pel, 0, GG * sizeof(double)), memset(own, 0, GG * sizeof(double)), memset(enemy, 0, GG * sizeof(double));
    memset(vir, 0, GG * sizeof(double)), memset(wall, 0, GG * sizeof(double));
    double midx = left + gs / 2, midy = top + gs / 2;
    for (int c = 0; c < G; ++c) {
        for (int r = 0; r < G; ++r) {
            int count = r + c * G;
            if (!(midx + gs / 2 < 0 || midx - gs / 2 > S || midy + gs / 2 < 0 || midy - gs / 2 > S)) {
                if (g.pel_has[count]) pel[count] = g.pel_sum[count];
                if (g.enemy_has[count]) enemy[count] = cf->normalize_grid_by_max_mass ? g.enemy_max[count] / biggest : g.enemy_max[count];
                if (g.own_has[count]) own[count] = cf->normalize_grid_by_max_mass ? g.own_max[count] / biggest : g.own_max[count];
                if (cf->virus_enabled && g.vir_has[count]) vir[count] = g.vir_mass[count];
            }
            double lb = py_minS(S, py_max0(midx - gs / 2)), tb = py_minS(S, py_max0(midy - gs / 2));
            double rb = py_max0(py_minS(S, midx + gs / 2)), bb = py_max0(py_minS(S, midy + gs / 2));
            double free_area = (rb - lb) * (bb - tb);
            wall[count] = round_dec(1 - (free_area / m_sq(gs)), 3);
            midx += gs;
        }
        midx = left + gs / 2;
        midy += gs;
    }
    /* channel order bot.py:458-495 */
    static __thread double out[11 * 84 * 84 + 16];
    int n = 0;
    float* hist = e->L.n_hist ? e->hist + (size_t)agent * e->L.n_hist * GG : NULL;
#define EMIT(src)                                      \
    do {                                               \
        for (int i = 0; i < GG; ++i) out[n + i] = (src)[i]; \
        n += GG;                                       \
    } while (0)
#define EMITF(src)                                     \
    do {                                               \
        for (int i = 0; i < GG; ++i) out[n + i] = (double)(src)[i]; \
        n += GG;                                       \
    } while (0)
    if (cf->pellet_grid) EMIT(pel);
    if (cf->self_grid) EMIT(own);
    if (cf->wall_grid) EMIT(wall);
    if (cf->enemy_grid) EMIT(enemy);
    if (cf->all_player_grid) { /* bot.py:348-351, 407-413, 472-474: biggest cell of any player in the square */
        for (int i = 0; i < GG; ++i) out[n + i] = own[i] > enemy[i] ? own[i] : enemy[i];
        n += GG;
    }
    if (cf->virus_grid) EMIT(vir);
    if (cf->self_grid_slf) {
        EMITF(hist + 1 * GG);
        memcpy(hist + 1 * GG, hist + 0 * GG, GG * sizeof(float));
    }
    if (cf->self_grid_lf) {
        EMITF(hist + 0 * GG);
        for (int i = 0; i < GG; ++i) hist[0 * GG + i] = (float)own[i];
    }
    if (cf->enemy_grid_slf) {
        EMITF(hist + 3 * GG);
        memcpy(hist + 3 * GG, hist + 2 * GG, GG * sizeof(float));
    }
    if (cf->enemy_grid_lf) {
        EMITF(hist + 2 * GG);
        for (int i = 0; i < GG; ++i) hist[2 * GG + i] = (float)enemy[i];
    }
#undef EMIT
#undef EMITF
    /* getAdditionalFeatures bot.py:302-323 */
    if (cf->use_last_fovsize) {
        B->last_fov_size_feat = B->fov_size_feat;
        out[n++] = B->last_fov_size_feat;
    }
    if (cf->use_fovsize) {
        update_fov_size(e, k);
        B->fov_size_feat = p->fov_size;
        out[n++] = B->fov_size_feat;
    }
    if (cf->use_totalmass) out[n++] = total_mass(e, k);
    if (cf->use_last_action)
        for (int i = 0; i < 4; ++i) out[n++] = B->has_action ? B->cur_action[i] : 0.0;
    if (cf->use_second_last_action)
        for (int i = 0; i < 4; ++i) out[n++] = B->has_last_action ? B->last_action[i] : 0.0;
    if (n != e->L.state_len) abort();
    if (obs)
        for (int i = 0; i < n; ++i) obs[i] = (float)out[i];
    if (obs64)
        for (int i = 0; i < n; ++i) obs64[i] = out[i];
}

/* Bot.getSimpleStateRepresentation (bot.py:511-548; GRID_VIEW_ENABLED = False, networkParameters.py:119): 12 values.
 * Candidates come from the hash tables like everywhere else (rectangle + INHASH flag); `min(..., key=squaredDistance)` keeps the
 * FIRST minimal candidate of the list, which the harness orders canonically (int pellets by slot, then ex-blob pellets; players
 * by index, cells by list position) — the reference's own order is a Python set's. */
static void rel_cell_pos(double cx, double cy, int left, int top, int size, double* o) { /* getRelativeCellPos :16-20 */
    o[0] = round_dec((cx - left) / size, 5);
    o[1] = round_dec((cy - top) / size, 5);
}
static void simple_state_agent(OracleEnv* e, int k, float* obs, double* obs64) {
    AgarPlayer* p = &e->pl[k];
    double S = (double)e->S;
    update_fov_size(e, k);
    update_fov_pos(e, k);
    const double fov = p->fov_size, fx = p->fov_x, fy = p->fov_y;
    const int x = (int)fx, y = (int)fy;
    const int left = x - (int)(fov / 2), top = y - (int)(fov / 2), size = (int)fov;
    const AgarCell* first = CELLP(e, k, 0);
    double out[AGAR_SIMPLE_STATE_LEN];
    memset(out, 0, sizeof(out));
    /* getCellDataOwnPlayer :648-652 -> isRelativeCellData :635-637 of the first cell */
    rel_cell_pos(first->x, first->y, left, top, size, out);
    out[2] = round_dec(first->radius <= size ? first->radius / size : 1.0, 5);
    /* closest enemy cell in the (float) field of view: field.py:434-436 */
    {
        Rect ra = rect_of(e, fx, fy, fov / 2);
        const AgarCell* best = NULL;
        double bd = 0;
        for (int k2 = 0; k2 < e->L.n_players; ++k2) {
            if (k2 == k) continue;
            for (int j = 0; j < e->pl[k2].n_cells; ++j) {
                const AgarCell* o = CELLP(e, k2, j);
                if (!(o->flags & AGAR_CF_INHASH) || !rect_hit(ra, rect_of(e, o->x, o->y, o->radius)) ||
                    !in_fov(o->x, o->y, o->radius, fx, fy, fov))
                    continue;
                double d = (o->x - first->x) * (o->x - first->x) + (o->y - first->y) * (o->y - first->y); /* cell.py:158-160 */
                if (!best || d < bd) best = o, bd = d;
            }
        }
        if (best) {
            rel_cell_pos(best->x, best->y, left, top, size, out + 3);
            out[5] = round_dec(best->radius <= size ? best->radius / size : 1.0, 5);
        }
    }
    /* closest pellet in the INTEGER field of view: getPelletsInFov(midPoint, int(size)), field.py:442-444 */
    {
        const double isz = (double)size;
        Rect ra = rect_of(e, fx, fy, isz / 2);
        int have = 0;
        double bd = 0, bx = 0, by = 0;
        for (int s = 0; s < e->L.pellet_cap; ++s) {
            uint32_t pk = e->pel[s];
            if (!pk) continue;
            int px = AGAR_PELLET_X(pk), py = AGAR_PELLET_Y(pk), pm = AGAR_PELLET_M(pk);
            double pr = pellet_radius(pm);
            if (!rect_hit(ra, rect_of(e, px, py, pr)) || !in_fov(px, py, pr, fx, fy, isz)) continue;
            double d = (px - first->x) * (px - first->x) + (py - first->y) * (py - first->y);
            if (!have || d < bd) have = 1, bd = d, bx = px, by = py;
        }
        for (int s = 0; s < e->L.fat_cap; ++s) {
            const AgarFatPellet* f = &e->fat[s];
            if (f->mass == 0) continue;
            if (!rect_hit(ra, rect_of(e, f->x, f->y, f->radius)) || !in_fov(f->x, f->y, f->radius, fx, fy, isz)) continue;
            double d = (f->x - first->x) * (f->x - first->x) + (f->y - first->y) * (f->y - first->y);
            if (!have || d < bd) have = 1, bd = d, bx = f->x, by = f->y;
        }
        if (have) rel_cell_pos(bx, by, left, top, size, out + 6);
    }
    /* distances to the visible field edges :541-547 (ints divided by an int: true division) */
    out[8] = left <= 0 ? (double)x / size : 1.0;
    out[9] = left + size >= e->S ? (S - x) / size : 1.0;
    out[10] = top <= 0 ? (double)y / size : 1.0;
    out[11] = top + size >= e->S ? (S - y) / size : 1.0;
    if (obs)
        for (int i = 0; i < AGAR_SIMPLE_STATE_LEN; ++i) obs[i] = (float)out[i];
    if (obs64)
        for (int i = 0; i < AGAR_SIMPLE_STATE_LEN; ++i) obs64[i] = out[i];
}

/* ------------------------------------------------------------------ bots (bot.py) */
static double bot_reward(OracleEnv* e, int k) { /* getReward :654-667 */
    const AgarConfig* cf = &e->cfg;
    AgarPlayer* p = &e->pl[k];
    if (cf->mass_as_reward) return p->alive ? total_mass(e, k) - cf->reward_term : cf->death_term - cf->reward_term;
    double reward;
    if (!p->alive)
        reward = -1 * p->bot.last_mass * cf->death_factor + cf->death_term;
    else
        reward = total_mass(e, k) - p->bot.last_mass;
    return reward * cf->reward_scale - cf->reward_term;
}
static void set_command_point(OracleEnv* e, int k, const double* a, int len) { /* :550-577 */
    AgarPlayer* p = &e->pl[k];
    update_fov_pos(e, k);
    update_fov_size(e, k);
    int x = (int)p->fov_x, y = (int)p->fov_y;
    int left = x - (int)(p->fov_size / 2), top = y - (int)(p->fov_size / 2);
    int size = (int)p->fov_size;
    p->cmd_x = left + a[0] * size;
    p->cmd_y = top + a[1] * size;
    int split = 0, eject = 0;
    if (len == 3)
        split = a[2] > 0.5; /* ENABLE_SPLIT branch; eject-only raises in the reference and is rejected by the layout */
    else if (len == 4) {
        split = a[2] > 0.5;
        eject = a[3] > 0.5;
    }
    p->do_split = split;
    p->do_eject = eject;
}
/* move_NN first half (bot.py:195-217) */
static void nn_turn_begin(OracleEnv* e, int k, int agent, float* obs, double* obs64) {
    AgarPlayer* p = &e->pl[k];
    AgarBot* B = &p->bot;
    if (B->turn_begun) return;
    /* makeMove :253 totalMasses.append */
    double tm = total_mass(e, k);
    B->stat_mass_sum += tm;
    if (tm > B->stat_mass_max) B->stat_mass_max = tm;
    B->stat_frames += 1;
    B->skipping = 0;
    B->need_action = B->exp_valid = B->exp_done = 0;
    if (B->has_action) {
        if (B->has_last_mass && B->last_mass != 0) B->cum_reward += bot_reward(e, k); /* updateRewards */
        B->last_reward = B->cum_reward;
        if (B->skip_frames > 0) { /* updateFrameSkip */
            B->skip_frames -= 1;
            if (p->alive) B->skipping = 1;
        }
    }
    if (!B->skipping) {
        if (p->alive) {
            if (e->cfg.simple_state)
                simple_state_agent(e, k, obs, obs64);
            else
                observe_agent(e, k, agent, obs, obs64);
        }
        if (B->has_old_state) {
            B->time += 1;
            B->exp_valid = 1;
            B->exp_done = !p->alive;
        }
        B->need_action = p->alive;
    }
    B->turn_begun = 1;
}
/* move_NN second half (bot.py:223-232) + the tail of makeMove (:256-270) */
static void nn_turn_end(OracleEnv* e, int k, const float* action) {
    AgarPlayer* p = &e->pl[k];
    AgarBot* B = &p->bot;
    if (B->need_action) { /* decideMove + updateValues :180-192 */
        B->cum_reward = 0;
        B->skip_frames = e->cfg.frame_skip;
        B->has_old_state = 1;
        memcpy(B->last_action, B->cur_action, sizeof B->cur_action);
        B->has_last_action = B->has_action;
        for (int i = 0; i < 4; ++i) B->cur_action[i] = i < e->L.action_len ? (double)action[i] : 0.0;
        B->has_action = 1;
    }
    if (!B->skipping && p->alive) {
        B->last_mass = total_mass(e, k);
        B->has_last_mass = 1;
    }
    B->turn_begun = 0;
    if (!p->alive) return;
    double a[4];
    memcpy(a, B->cur_action, sizeof a);
    int len = e->L.action_len;
    if (B->skipping) {
        a[2] = a[3] = 0;
        len = 4;
    }
    set_command_point(e, k, a, len);
}
static void scripted_turn(OracleEnv* e, int k) {
    AgarPlayer* p = &e->pl[k];
    AgarBot* B = &p->bot;
    double tm = total_mass(e, k);
    B->stat_mass_sum += tm;
    if (tm > B->stat_mass_max) B->stat_mass_max = tm;
    B->stat_frames += 1;
    if (!p->alive) return;
    if (B->type == AGAR_BOT_GREEDY) { /* make_greedy_bot_move :579-633 */
        update_fov_pos(e, k);
        update_fov_size(e, k);
        double fx = p->fov_x, fy = p->fov_y, fov = p->fov_size;
        int x = (int)fx, y = (int)fy;
        int left = x - (int)(fov / 2), top = y - (int)(fov / 2);
        int big = 0;
        for (int i = 1; i < p->n_cells; ++i)
            if (CELLP(e, k, i)->mass > CELLP(e, k, big)->mass) big = i;
        const AgarCell* bc = CELLP(e, k, big);
        Rect ra = rect_of(e, fx, fy, fov / 2);
        int have = 0;
        double best = 0, bestx = 0, besty = 0;
#define CONSIDER(ox, oy, om)                                                            \
    do {                                                                                \
        double _d2 = ((ox)-bc->x) * ((ox)-bc->x) + ((oy)-bc->y) * ((oy)-bc->y);         \
        double _key = (om) / (_d2 != 0 ? _d2 : 1);                                      \
        if (!have || _key > best) have = 1, best = _key, bestx = (ox), besty = (oy);    \
    } while (0)
        for (int s = 0; s < e->L.pellet_cap; ++s) {
            uint32_t pk = e->pel[s];
            if (!pk) continue;
            int px = AGAR_PELLET_X(pk), py = AGAR_PELLET_Y(pk), pm = AGAR_PELLET_M(pk);
            double pr = pellet_radius(pm);
            if (!rect_hit(ra, rect_of(e, px, py, pr)) || !in_fov(px, py, pr, fx, fy, fov)) continue;
            CONSIDER((double)px, (double)py, (double)pm);
        }
        for (int s = 0; s < e->L.fat_cap; ++s) {
            const AgarFatPellet* f = &e->fat[s];
            if (f->mass == 0) continue;
            if (!rect_hit(ra, rect_of(e, f->x, f->y, f->radius)) || !in_fov(f->x, f->y, f->radius, fx, fy, fov)) continue;
            CONSIDER(f->x, f->y, f->mass);
        }
        for (int k2 = 0; k2 < e->L.n_players; ++k2)
            for (int j = 0; j < e->pl[k2].n_cells; ++j) {
                const AgarCell* o = CELLP(e, k2, j);
                if (k2 == k || !(o->flags & AGAR_CF_INHASH) || !rect_hit(ra, rect_of(e, o->x, o->y, o->radius)) ||
                    !in_fov(o->x, o->y, o->radius, fx, fy, fov))
                    continue;
                if (bc->mass > 1.25 * o->mass) CONSIDER(o->x, o->y, o->mass);
            }
        if (e->cfg.virus_enabled)
            for (int v = 0; v < e->h->n_viruses; ++v) {
                const AgarMote* o = &e->vir[v];
                if (!(o->aux & AGAR_CF_INHASH) || !rect_hit(ra, rect_of(e, o->x, o->y, o->radius)) ||
                    !in_fov(o->x, o->y, o->radius, fx, fy, fov))
                    continue;
                if (bc->mass > 1.25 * o->mass) CONSIDER(o->x, o->y, o->mass);
            }
#undef CONSIDER
        if (have) { /* getRelativeCellPos :16-20 — relative to the FLOAT fov size */
            B->cur_action[0] = round_dec((bestx - left) / fov, 5);
            B->cur_action[1] = round_dec((besty - top) / fov, 5);
        } else {
            B->cur_action[0] = draw_random(e, 1);
            B->cur_action[1] = draw_random(e, 1);
        }
        B->cur_action[2] = B->cur_action[3] = 0;
    } else { /* make_random_bot_move :243-249 */
        if (e->cfg.frame_skip == 0 || B->time % e->cfg.frame_skip == 0) {
            B->cur_action[0] = draw_random(e, 1);
            B->cur_action[1] = draw_random(e, 1);
            B->cur_action[2] = e->cfg.enable_split ? draw_random(e, 1) : 0;
            B->cur_action[3] = e->cfg.enable_eject ? draw_random(e, 1) : 0;
        }
        B->time += 1;
    }
    B->has_action = 1;
    set_command_point(e, k, B->cur_action, 4);
}
static void bot_reset(OracleEnv* e, int k, int agent) { /* Bot.reset :125-164 */
    AgarBot* B = &e->pl[k].bot;
    B->has_last_mass = 0, B->last_mass = 0;
    B->has_old_state = 0;
    B->skip_frames = 0;
    B->cum_reward = 0, B->last_reward = 0;
    B->skipping = 0;
    B->turn_begun = B->need_action = B->exp_valid = B->exp_done = 0;
    if (B->type == AGAR_BOT_NN) {
        B->has_action = 0;
        memset(B->cur_action, 0, sizeof B->cur_action);
        if (e->L.n_hist)
            memset(e->hist + (size_t)agent * e->L.n_hist * e->L.grid_squares * e->L.grid_squares, 0,
                   sizeof(float) * e->L.n_hist * e->L.grid_squares * e->L.grid_squares);
        B->fov_size_feat = 0, B->last_fov_size_feat = 0;
    } else {
        B->has_action = 1;
        memset(B->cur_action, 0, sizeof B->cur_action);
    }
}

/* ------------------------------------------------------------------ public API (ctypes) */
static void bind(OracleEnv* e, void* rec) {
    e->rec = (uint8_t*)rec;
    e->h = (AgarEnvHeader*)(e->rec + e->L.off_header);
    e->pl = (AgarPlayer*)(e->rec + e->L.off_players);
    e->cells = (AgarCell*)(e->rec + e->L.off_cells);
    e->vir = (AgarMote*)(e->rec + e->L.off_viruses);
    e->blob = (AgarMote*)(e->rec + e->L.off_blobs);
    e->fat = (AgarFatPellet*)(e->rec + e->L.off_fat);
    e->pel = (uint32_t*)(e->rec + e->L.off_pellets);
    e->hist = (float*)(e->rec + e->L.off_hist);
    e->ev = (AgarEvent*)(e->rec + e->L.off_events);
}
int oracle_layout(const AgarConfig* cfg, AgarLayout* out) { return agar_layout_compute(cfg, out); }

OracleEnv* oracle_create(const AgarConfig* cfg, uint64_t seed, uint64_t env_id) {
    OracleEnv* e = (OracleEnv*)calloc(1, sizeof *e);
    e->cfg = *cfg;
    if (agar_layout_compute(cfg, &e->L) != AGAR_OK) {
        free(e);
        return NULL;
    }
    e->seed = seed, e->env_id = env_id;
    e->S = e->L.field_size;
    e->nb = (int)ceil((double)e->S / BUCKET);
    e->speed_modifier = 1.0 / FPS;
    e->move_speed = 90 * e->speed_modifier;
    e->decay_rate = 1 - (0.01 * e->speed_modifier);
    e->blob_mass = EJECT_BASE_MASS * 0.8;
    e->virus_split_mass = VIRUS_BASE_SIZE + 7 * EJECT_BASE_MASS * 0.8;
    e->start_radius = sqrt(START_MASS / M_PI);
    e->virus_radius = sqrt(VIRUS_BASE_SIZE / M_PI);
    for (int d = 0; d < 360; ++d) { /* numpy.deg2rad(int) == d * (pi / 180) (verified) */
        double a = d * (M_PI / 180.0);
        e->cos_deg[d] = cos(a), e->sin_deg[d] = sin(a);
    }
    for (int n = 1; n <= 16; ++n) e->pow_n[n] = pow((double)n, 0.32);
    bind(e, calloc(1, e->L.record_bytes));
    e->owns_rec = 1;
    /* Model(...) + createBot * K (model.py:51,154-162; player.py:11-28) */
    for (int k = 0; k < e->L.n_players; ++k) {
        AgarPlayer* p = &e->pl[k];
        p->alive = 1;
        p->cmd_x = p->cmd_y = -1;
        p->bot.type = cfg->bot_type[k];
    }
    e->h->event_hash = 0xCBF29CE484222325ULL;
    /* Model.initialize: Field.initialize (field.py:57-67) + resetBots */
    for (int k = 0; k < e->L.n_players; ++k) initialize_player(e, k);
    spawn_stuff(e);
    int agent = 0;
    for (int k = 0; k < e->L.n_players; ++k) {
        bot_reset(e, k, agent);
        if (cfg->bot_type[k] == AGAR_BOT_NN) ++agent;
    }
    return e;
}
void oracle_destroy(OracleEnv* e) {
    if (!e) return;
    if (e->owns_rec) free(e->rec);
    free(e);
}
void* oracle_record(OracleEnv* e) { return e->rec; }
uint64_t oracle_record_bytes(const OracleEnv* e) { return e->L.record_bytes; }
void oracle_load_record(OracleEnv* e, const void* rec) { memcpy(e->rec, rec, e->L.record_bytes); }
void oracle_set_key(OracleEnv* e, uint64_t seed, uint64_t env_id) { e->seed = seed, e->env_id = env_id; }

/* Model.resetModel (model.py:96-98) -> Field.reset (field.py:69-83) */
void oracle_reset(OracleEnv* e) {
    e->h->n_events = 0;
    e->h->event_hash = 0xCBF29CE484222325ULL;
    memset(e->pel, 0, sizeof(uint32_t) * e->L.pellet_cap);
    memset(e->fat, 0, sizeof(AgarFatPellet) * e->L.fat_cap);
    memset(e->blob, 0, sizeof(AgarMote) * e->L.blob_cap);
    memset(e->vir, 0, sizeof(AgarMote) * e->L.virus_cap);
    e->h->n_pellets = e->h->n_fat = e->h->n_blobs = e->h->n_viruses = 0;
    e->h->n_dead = 0;
    memset(e->h->dead_order, 0, sizeof e->h->dead_order);
    /* fresh hash tables: nothing is in the player table until the next frame rebuilds it */
    for (int k = 0; k < e->L.n_players; ++k)
        for (int i = 0; i < e->pl[k].n_cells; ++i) CELLP(e, k, i)->flags &= ~AGAR_CF_INHASH;
    for (int k = 0; k < e->L.n_players; ++k) initialize_player(e, k);
    spawn_stuff(e);
    e->h->frame = 0;
}
void oracle_reset_bots(OracleEnv* e) {
    int agent = 0;
    for (int k = 0; k < e->L.n_players; ++k) {
        bot_reset(e, k, agent);
        if (e->cfg.bot_type[k] == AGAR_BOT_NN) ++agent;
    }
}
/* first half of every NN bot's turn; obs: float[A][L] (rows of skipping / dead agents untouched) */
void oracle_observe(OracleEnv* e, float* obs, double* obs64) {
    for (int a = 0; a < e->L.n_agents; ++a)
        nn_turn_begin(e, a, a, obs ? obs + (size_t)a * e->L.state_len : NULL,
                      obs64 ? obs64 + (size_t)a * e->L.state_len : NULL);
}
/* n_frames x Model.update (model.py:100-112); actions: float[A][4] */
void oracle_step(OracleEnv* e, const float* actions, int n_frames) {
    for (int f = 0; f < n_frames; ++f) {
        e->h->n_events = 0; /* events are per frame; the running hash is not */
        for (int k = 0; k < e->L.n_players; ++k) {
            if (e->cfg.bot_type[k] == AGAR_BOT_NN) {
                nn_turn_begin(e, k, k, NULL, NULL);
                nn_turn_end(e, k, actions + (size_t)k * 4);
            } else
                scripted_turn(e, k);
        }
        field_update(e);
        e->h->frame += 1;
    }
}
/* per-agent turn results, AgarField order of agar_get */
void oracle_get_turn(const OracleEnv* e, float* reward, uint8_t* done, uint8_t* valid, uint8_t* need_action) {
    for (int a = 0; a < e->L.n_agents; ++a) {
        const AgarBot* B = &e->pl[a].bot;
        if (reward) reward[a] = (float)B->last_reward;
        if (done) done[a] = (uint8_t)B->exp_done;
        if (valid) valid[a] = (uint8_t)B->exp_valid;
        if (need_action) need_action[a] = (uint8_t)B->need_action;
    }
}

/* n_decisions x (observe -> uniform random action from Philox stream 7 -> n_frames frames): the single-env twin
 * of agar_rollout_random (random-action driver of BASELINE config 2).  obs (nullable): float[A][L], overwritten
 * at every decision. */
void oracle_rollout_random(OracleEnv* e, int n_decisions, int n_frames, uint32_t decision_base, float* obs) {
    float act[AGAR_MAX_PLAYERS * 4];
    for (int d = 0; d < n_decisions; ++d) {
        oracle_observe(e, obs, NULL);
        for (int a = 0; a < e->L.n_agents; ++a) {
            uint32_t w[4];
            philox(decision_base + (uint32_t)d, 7u, (uint32_t)e->env_id, (uint32_t)a, (uint32_t)e->seed,
                   (uint32_t)(e->seed >> 32), w);
            for (int q = 0; q < 4; ++q) act[a * 4 + q] = (float)(w[q] >> 8) * (1.0f / 16777216.0f);
        }
        oracle_step(e, act, n_frames);
    }
}

/* ---- batched multi-thread driver for bench.py's CPU baseline: E independent envs, T decision periods of
 * (observe; step frame_skip+1 frames) with uniform random actions (SURVEY §8d config 2 driver).  Returns
 * env-steps executed.  One pthread per requested thread, envs dealt round-robin. */
#include <pthread.h>
typedef struct BatchJob {
    const AgarConfig* cfg;
    int n_envs, n_decisions, tid, n_threads;
    uint64_t seed, first_env, steps;
    double mass_sum;
} BatchJob;
static void* batch_worker(void* arg) {
    BatchJob* j = (BatchJob*)arg;
    int period = j->cfg->frame_skip + 1;
    for (int i = j->tid; i < j->n_envs; i += j->n_threads) {
        OracleEnv* e = oracle_create(j->cfg, j->seed, j->first_env + (uint64_t)i);
        float* obs = (float*)malloc(sizeof(float) * e->L.state_len * (e->L.n_agents ? e->L.n_agents : 1));
        oracle_rollout_random(e, j->n_decisions, period, 0, obs);
        j->steps += (uint64_t)period * (uint64_t)j->n_decisions;
        for (int k = 0; k < e->L.n_players; ++k) j->mass_sum += total_mass(e, k);
        free(obs);
        oracle_destroy(e);
    }
    return NULL;
}
uint64_t oracle_rollout_batch(const AgarConfig* cfg, int n_envs, uint64_t seed, uint64_t first_env, int n_decisions,
                              int n_threads, double* mass_sum_out) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    pthread_t th[256];
    BatchJob jobs[256];
    for (int t = 0; t < n_threads; ++t) {
        jobs[t] = (BatchJob){cfg, n_envs, n_decisions, t, n_threads, seed, first_env, 0, 0.0};
        pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
    }
    uint64_t steps = 0;
    double mass = 0;
    for (int t = 0; t < n_threads; ++t) {
        pthread_join(th[t], NULL);
        steps += jobs[t].steps;
        mass += jobs[t].mass_sum;
    }
    if (mass_sum_out) *mass_sum_out = mass;
    return steps;
}

/* ---- thin exports of include/agar_math.h so that tests can bound the portable math against libm / Python */
double oracle_pm_pow(double x, double y) { return agar_pow(x, y); }
double oracle_pm_log(double x) { return agar_log(x); }
double oracle_pm_exp(double x) { return agar_exp(x); }
/* agar_pow against the C library's pow (what CPython's math.pow calls) on the path's domain: n pseudo-random inputs
 * (radii, masses, a wide log-uniform band) x the exponents the path and the replay buffer use, plus every mass on a
 * 1/64 grid up to the 22500 cap and its radius.  Returns the number of results whose BITS differ. */
uint64_t oracle_pm_pow_mismatches(uint64_t n, uint64_t seed, double* first_x, double* first_y) {
    static const double ys[8] = {0.475, -0.35, 0.32, 2.0, 0.6, -0.4, -1.0, 0.5};
    uint64_t s = seed * 0x9E3779B97F4A7C15ULL + 88172645463325252ULL, bad = 0;
    for (uint64_t i = 0; i < n; ++i) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        double u = (double)(s >> 11) * (1.0 / 9007199254740992.0), x;
        switch (i & 3) {
            case 0: x = 0.5 + u * 90; break;
            case 1: x = 1 + u * 22500; break;
            case 2: x = exp((u - 0.5) * 40); break;
            default: x = 4 + u * 200; break;
        }
        double y = ys[(i >> 2) & 7], a = agar_pow(x, y), b = pow(x, y);
        if (memcmp(&a, &b, 8)) {
            if (!bad && first_x) { *first_x = x; *first_y = y; }
            ++bad;
        }
    }
    for (int j = 64; j <= 22500 * 64; ++j) {
        double m = j / 64.0, r = sqrt(m / M_PI);
        double a = agar_pow(m, -0.35), b = pow(m, -0.35), c = agar_pow(r, 0.475), d = pow(r, 0.475);
        if (memcmp(&a, &b, 8)) { if (!bad && first_x) { *first_x = m; *first_y = -0.35; } ++bad; }
        if (memcmp(&c, &d, 8)) { if (!bad && first_x) { *first_x = r; *first_y = 0.475; } ++bad; }
    }
    return bad;
}
/* agar_atan2 / agar_sin / agar_cos against the C library on n pseudo-random inputs: direction vectors of every magnitude
 * mix the game produces (field-scale, near-axis, tiny, integer lattice, near-diagonal, 35 decades log-uniform), and angles
 * both uniform in [-pi, pi] and exactly as atan2 returns them.  Returns the number of results whose BITS differ. */
uint64_t oracle_pm_trig_mismatches(uint64_t n, uint64_t seed) {
    uint64_t s = seed * 0x9E3779B97F4A7C15ULL + 88172645463325252ULL, bad = 0;
#define U01() (s ^= s << 13, s ^= s >> 7, s ^= s << 17, (double)(s >> 11) * (1.0 / 9007199254740992.0))
    for (uint64_t i = 0; i < n; ++i) {
        double x, y;
        switch (i & 7) {
            case 0: x = (U01() - 0.5) * 600; y = (U01() - 0.5) * 600; break;
            case 1: x = (U01() - 0.5) * 2; y = (U01() - 0.5) * 600; break;
            case 2: x = (U01() - 0.5) * 600; y = (U01() - 0.5) * 2; break;
            case 3: x = (U01() - 0.5) * 1e-6; y = (U01() - 0.5) * 100; break;
            case 4: x = (U01() - 0.5) * 100; y = (U01() - 0.5) * 1e-9; break;
            case 5: x = floor((U01() - 0.5) * 100); y = floor((U01() - 0.5) * 100); break;
            case 6: {
                double m0 = exp((U01() - 0.5) * 80), s0 = U01(), m1 = exp((U01() - 0.5) * 80), s1 = U01();
                x = s0 < 0.5 ? -m0 : m0;
                y = s1 < 0.5 ? -m1 : m1;
                break;
            }
            default: {
                x = (U01() - 0.5) * 20;
                double f = U01(), g = U01();
                y = x * (1 + (f - 0.5) * 0.2);
                if (g < 0.5) y = -y;
                break;
            }
        }
        double a = agar_atan2(y, x), b = atan2(y, x);
        bad += memcmp(&a, &b, 8) != 0;
        double ang = (i & 8) ? b : (U01() - 0.5) * 2 * M_PI;
        double s1 = agar_sin(ang), s2 = sin(ang), c1 = agar_cos(ang), c2 = cos(ang);
        bad += memcmp(&s1, &s2, 8) != 0;
        bad += memcmp(&c1, &c2, 8) != 0;
        double s3, c3;
        agar_sincos(ang, &s3, &c3); /* the paired form the kernels call */
        bad += memcmp(&s3, &s2, 8) != 0;
        bad += memcmp(&c3, &c2, 8) != 0;
    }
#undef U01
    for (int yy = -300; yy <= 300; ++yy) /* fresh cells and pellets sit on integer coordinates */
        for (int xx = -300; xx <= 300; ++xx) {
            double a = agar_atan2(yy, xx), b = atan2(yy, xx);
            bad += memcmp(&a, &b, 8) != 0;
        }
    return bad;
}
double oracle_pm_atan2(double y, double x) { return agar_atan2(y, x); }
double oracle_pm_sin(double x) { return agar_sin(x); }
double oracle_pm_cos(double x) { return agar_cos(x); }
void oracle_pm_dir(double dy, double dx, double* c, double* s) { agar_dir(dy, dx, c, s); }
double oracle_pm_round_dec(double x, double scale) { return agar_round_dec(x, scale); }
/* spatialHashTable.py:70-83 for one axis, exported for the closed-form pellet-rectangle proof in tests */
void oracle_axis_range(double p, double radius, int S, int* b0, int* b1) { axis_range(p, radius, S, b0, b1); }
/* Field.update() alone (field.py:85-92), for the hand-placed known-answer scenes of SURVEY.md App. B */
void oracle_field_update(OracleEnv* e) {
    e->h->n_events = 0;
    field_update(e);
    e->h->frame += 1;
}
