/*
 * agar_replay.cu — GPU replay buffer (uniform ring + prioritized with sum / min segment trees).
 * Replaces src/model/replay_buffer.py and src/model/common/segment_tree.py; C ABI in include/agar_replay.h.
 * HBM-bound: add / gather move 2 * state_len + action_len + 2 floats per transition with coalesced row copies;
 * the trees are float64 arrays of 2 * it_capacity nodes rebuilt level by level (a node is op(left, right), so the
 * tree is a function of the leaves only and batched updates commute).
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/agar_b200.h"
#include "../../include/agar_math.h"
#include "../../include/agar_replay.h"

struct ReplayDev {
    float *obs_t, *obs_tp1, *action, *reward;
    uint8_t* done;
    double *sum, *mn;    /* segment trees, 2 * itcap each (node 1 = root) */
    int32_t* counters;   /* [0] next_idx  [1] size  [2] n_valid of the last add */
    double* max_priority;
    int32_t* pos;        /* scratch: ring position per candidate of the last add (-1 = not stored) */
    int cap, itcap, L, AL, prioritized;
    double alpha, beta;
};
struct AgarReplay {
    ReplayDev d;
    int device, pos_cap;
    int64_t launches;
    char err[256];
};
static char g_rp_err[256] = "";

static int rp_fail(AgarReplay* rp, int code, const char* msg) {
    snprintf(rp ? rp->err : g_rp_err, 256, "%s", msg);
    return code;
}
#define RCU(call)                                                                       \
    do {                                                                                \
        cudaError_t _e = (call);                                                        \
        if (_e != cudaSuccess) return rp_fail(rp, AGAR_E_CUDA, cudaGetErrorString(_e)); \
    } while (0)

/* ---- add: ranks of the valid candidates (one CTA, chunked block scan), then one warp per stored transition */
__global__ void k_rp_scan(ReplayDev d, const uint8_t* __restrict__ valid, int n) {
    __shared__ int warp_tot[32];
    __shared__ int running;
    if (threadIdx.x == 0) running = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    /* pass 1: count */
    int total = 0;
    for (int base = 0; base < n; base += blockDim.x) {
        int i = base + threadIdx.x;
        int v = i < n && (valid == nullptr || valid[i]);
        unsigned b = __ballot_sync(0xffffffffu, v);
        if (lane == 0) warp_tot[wid] = __popc(b);
        __syncthreads();
        if (threadIdx.x == 0) {
            int s = 0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += warp_tot[w];
            running += s;
        }
        __syncthreads();
    }
    total = running;
    __syncthreads();
    const int next0 = d.counters[0], size0 = d.counters[1];
    const int skip = total > d.cap ? total - d.cap : 0; /* later adds of the same batch would overwrite these */
    if (threadIdx.x == 0) running = 0;
    __syncthreads();
    /* pass 2: positions */
    for (int base = 0; base < n; base += blockDim.x) {
        int i = base + threadIdx.x;
        int v = i < n && (valid == nullptr || valid[i]);
        unsigned b = __ballot_sync(0xffffffffu, v);
        if (lane == 0) warp_tot[wid] = __popc(b);
        __syncthreads();
        int before = running;
        for (int w = 0; w < wid; ++w) before += warp_tot[w];
        int rank = before + __popc(b & ((1u << lane) - 1));
        if (i < n) d.pos[i] = (v && rank >= skip) ? (int)(((long long)next0 + rank) % d.cap) : -1;
        __syncthreads();
        if (threadIdx.x == 0) {
            int s = 0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += warp_tot[w];
            running += s;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        d.counters[0] = (int)(((long long)next0 + total) % d.cap);
        long long sz = (long long)size0 + total;
        d.counters[1] = sz > d.cap ? d.cap : (int)sz;
        d.counters[2] = total;
    }
}
__global__ void k_rp_store(ReplayDev d, const float* __restrict__ obs_t, const float* __restrict__ action,
                           const float* __restrict__ reward, const float* __restrict__ obs_tp1,
                           const uint8_t* __restrict__ done, int n) {
    int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n) return;
    int p = d.pos[w];
    if (p < 0) return;
    const float* s0 = obs_t + (size_t)w * d.L;
    const float* s1 = obs_tp1 + (size_t)w * d.L;
    float* d0 = d.obs_t + (size_t)p * d.L;
    float* d1 = d.obs_tp1 + (size_t)p * d.L;
    for (int i = lane; i < d.L; i += 32) {
        d0[i] = s0[i];
        d1[i] = s1[i];
    }
    if (lane < d.AL) d.action[(size_t)p * d.AL + lane] = action[(size_t)w * d.AL + lane];
    if (lane == 0) {
        d.reward[p] = reward[w];
        d.done[p] = done[w];
        if (d.prioritized) { /* leaf = max_priority ** alpha (replay_buffer.py:111-112) */
            double v = agar_pow(*d.max_priority, d.alpha);
            d.sum[d.itcap + p] = v;
            d.mn[d.itcap + p] = v;
        }
    }
}
__global__ void k_rp_init(ReplayDev d) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 2 * d.itcap) {
        d.sum[i] = 0.0;
        d.mn[i] = INFINITY;
    }
    if (i == 0) {
        d.counters[0] = d.counters[1] = d.counters[2] = d.counters[3] = 0;
        *d.max_priority = 1.0;
    }
}
__global__ void k_rp_gather(ReplayDev d, const int32_t* __restrict__ idx, int batch, float* obs_t, float* action, float* reward,
                            float* obs_tp1, uint8_t* done) {
    int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= batch) return;
    int p = idx[w];
    if (p < 0 || p >= d.counters[1]) { /* the reference asserts 0 <= idx < len(storage); here: flag, read slot 0 */
        if (lane == 0) atomicOr(&d.counters[3], AGAR_RP_ERR_INDEX);
        p = 0;
    }
    const float* s0 = d.obs_t + (size_t)p * d.L;
    const float* s1 = d.obs_tp1 + (size_t)p * d.L;
    for (int i = lane; i < d.L; i += 32) {
        if (obs_t) obs_t[(size_t)w * d.L + i] = s0[i];
        if (obs_tp1) obs_tp1[(size_t)w * d.L + i] = s1[i];
    }
    if (action && lane < d.AL) action[(size_t)w * d.AL + lane] = d.action[(size_t)p * d.AL + lane];
    if (lane == 0) {
        if (reward) reward[w] = d.reward[p];
        if (done) done[w] = d.done[p];
    }
}
__global__ void k_rp_uniform_idx(ReplayDev d, const double* __restrict__ u, int batch, int32_t* idx) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch) return;
    int len = d.counters[1];
    if (len < 1) { /* random.randint(0, -1) raises in the reference; here: index 0 and a sticky error flag */
        atomicOr(&d.counters[3], AGAR_RP_ERR_EMPTY);
        idx[i] = 0;
        return;
    }
    int v = (int)(u[i] * (double)len); /* random.randint(0, len - 1) */
    idx[i] = v >= len ? len - 1 : v;
}
/* SegmentTree._reduce_helper for the prefix [0, end] (segment_tree.py:38-53): value[left child] + helper(right), i.e.
 * v1 + (v2 + (v3 + ...)) over the maximal nodes of the prefix */
__device__ double rp_prefix_sum(const ReplayDev& d, int end) {
    double parts[40];
    int np = 0;
    int node = 1, ns = 0, ne = d.itcap - 1;
    if (end < 0) return 0.0; /* callers guard len < 2; never descend with an end the loop cannot reach */
    if (end > ne) end = ne;
    while (true) {
        if (end == ne) {
            parts[np++] = d.sum[node];
            break;
        }
        int mid = (ns + ne) / 2;
        if (end <= mid) {
            node = 2 * node, ne = mid;
        } else {
            parts[np++] = d.sum[2 * node];
            node = 2 * node + 1, ns = mid + 1;
        }
    }
    double acc = parts[np - 1];
    for (int i = np - 2; i >= 0; --i) acc = parts[i] + acc;
    return acc;
}
__global__ void k_rp_per_sample(ReplayDev d, const double* __restrict__ u, int batch, int32_t* idx, double* weights) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch) return;
    const int len = d.counters[1];
    if (len < 2) { /* the reference recurses without end here (sum(0, len - 1) over an empty prefix); flag and return slot 0 */
        atomicOr(&d.counters[3], AGAR_RP_ERR_EMPTY);
        idx[i] = 0;
        if (weights) weights[i] = 0.0;
        return;
    }
    /* _sample_proportional (:113-120): sum(0, len - 1) reduces over [0, len - 2] (reduce() decrements `end`) */
    double total = rp_prefix_sum(d, len - 2);
    double mass = u[i] * total;
    int node = 1;
    while (node < d.itcap) { /* find_prefixsum_idx (segment_tree.py:107-126) */
        double l = d.sum[2 * node];
        if (l > mass)
            node = 2 * node;
        else {
            mass -= l;
            node = 2 * node + 1;
        }
    }
    int p = node - d.itcap;
    idx[i] = p;
    if (weights) { /* :157-166 */
        double root = d.sum[1];
        double p_min = d.mn[1] / root;
        double max_weight = agar_pow(p_min * len, -d.beta);
        double p_sample = d.sum[d.itcap + p] / root;
        weights[i] = agar_pow(p_sample * len, -d.beta) / max_weight;
    }
}
/* staged != 0: the launch carries batch * 4 bytes of dynamic shared memory and every CTA keeps a copy of idx there — the
 * "does a later entry name the same leaf" scan then runs on shared memory without an early exit (consecutive threads read
 * consecutive words): 2048 entries 123 us -> a few us */
__global__ void k_rp_set_priorities(ReplayDev d, const int32_t* __restrict__ idx, const double* __restrict__ prio, int batch, int staged) {
    extern __shared__ int32_t sh_idx[];
    if (staged) {
        for (int j = threadIdx.x; j < batch; j += blockDim.x) sh_idx[j] = idx[j];
        __syncthreads();
    }
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch) return;
    /* duplicates in idx: the reference applies them in order, the last one wins */
    if (staged) {
        const int32_t mine = sh_idx[i];
        bool later = false;
        for (int j = i + 1; j < batch; ++j) later |= sh_idx[j] == mine;
        if (later) return;
    } else {
        for (int j = i + 1; j < batch; ++j)
            if (idx[j] == idx[i]) return;
    }
    /* replay_buffer.py:203-204 asserts priority > 0 and 0 <= idx < len(storage): skip and flag instead of poisoning the trees */
    if (idx[i] < 0 || idx[i] >= d.counters[1]) {
        atomicOr(&d.counters[3], AGAR_RP_ERR_INDEX);
        return;
    }
    if (!(prio[i] > 0.0) || prio[i] > 1.7e308) {
        atomicOr(&d.counters[3], AGAR_RP_ERR_PRIORITY);
        return;
    }
    double v = agar_pow(prio[i], d.alpha);
    d.sum[d.itcap + idx[i]] = v;
    d.mn[d.itcap + idx[i]] = v;
}
/* _max_priority = max(_max_priority, priority) over the batch (replay_buffer.py:188): a maximum, so any order — one CTA, strided
 * loads and a shared-memory tree (a lone thread walking 2048 priorities took 190 us of a 750 us learner tick) */
__global__ void __launch_bounds__(256) k_rp_max_priority(ReplayDev d, const double* __restrict__ prio, int batch) {
    __shared__ double sm[256];
    double m = 0.0; /* priorities are > 0 (guarded by k_rp_set_priorities) */
    for (int i = threadIdx.x; i < batch; i += 256) {
        const double v = prio[i];
        if (v > m && v <= 1.7e308) m = v;
    }
    sm[threadIdx.x] = m;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off && sm[threadIdx.x + off] > sm[threadIdx.x]) sm[threadIdx.x] = sm[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0 && sm[0] > *d.max_priority) *d.max_priority = sm[0];
}

/* Repair the sum / min trees above the n leaves idx[0..n) (entries < 0 are skipped) in ONE launch: a node is
 * op(left child, right child) (segment_tree.py:76-87), so only the ancestors of touched leaves change, level by level.  One
 * CTA walks the levels with a barrier in between; threads that reach the same parent recompute the same value from the same
 * finished children (a benign duplicate).  Replaces 20 level-wide launches per add / priority update (the learner tick is
 * launch-bound: ~150 small kernels). */
__global__ void __launch_bounds__(1024) k_rp_fix_paths(ReplayDev d, const int32_t* __restrict__ idx, int n) {
    int levels = 0;
    while ((1 << levels) < d.itcap) ++levels;
    for (int lv = 1; lv <= levels; ++lv) {
        const int width = d.itcap >> lv; /* nodes on this level */
        if (width <= n) { /* near the root there are fewer nodes than touched leaves: recompute the level (same values, fewer loads) */
            for (int i = threadIdx.x; i < width; i += blockDim.x) {
                const int node = width + i;
                d.sum[node] = d.sum[2 * node] + d.sum[2 * node + 1];
                const double a = d.mn[2 * node], b = d.mn[2 * node + 1];
                d.mn[node] = b < a ? b : a;
            }
        } else {
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const int p = idx[i];
                if (p < 0 || p >= d.cap) continue;
                const int node = (d.itcap + p) >> lv;
                d.sum[node] = d.sum[2 * node] + d.sum[2 * node + 1];
                const double a = d.mn[2 * node], b = d.mn[2 * node + 1];
                d.mn[node] = b < a ? b : a; /* Python min(a, b): b only if b < a */
            }
        }
        __syncthreads();
    }
}
/* all nodes of one tree level: value[node] = op(value[2 node], value[2 node + 1]) (segment_tree.py:76-87) */
__global__ void k_rp_level(ReplayDev d, int first, int count) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    int node = first + i;
    d.sum[node] = d.sum[2 * node] + d.sum[2 * node + 1];
    double a = d.mn[2 * node], b = d.mn[2 * node + 1];
    d.mn[node] = b < a ? b : a;
}
/* few touched leaves (a learner's batch, a tick of a few thousand envs): one launch that walks their ancestors; many (a tick of
 * 65k envs touches a large part of every level anyway): one grid-wide launch per level */
static int rp_fix_paths(AgarReplay* rp, const int32_t* idx_dev, int n, cudaStream_t s) {
    if (n <= 8192) {
        k_rp_fix_paths<<<1, 1024, 0, s>>>(rp->d, idx_dev, n);
        rp->launches += 1;
    } else {
        for (int count = rp->d.itcap / 2; count >= 1; count /= 2) {
            k_rp_level<<<(count + 255) / 256, 256, 0, s>>>(rp->d, count, count);
            rp->launches += 1;
        }
    }
    RCU(cudaGetLastError());
    return AGAR_OK;
}

extern "C" int agar_replay_create(int capacity, int state_len, int action_len, int prioritized, double alpha, double beta,
                                  int device, AgarReplay** out) {
    AgarReplay* rp = nullptr;
    if (!out || capacity < 1 || state_len < 1 || action_len < 1 || action_len > 32) return rp_fail(nullptr, AGAR_E_INVALID, "bad arguments");
    if (prioritized && (alpha < 0 || beta <= 0)) return rp_fail(nullptr, AGAR_E_INVALID, "alpha >= 0 and beta > 0 required");
    RCU(cudaSetDevice(device));
    rp = (AgarReplay*)calloc(1, sizeof(AgarReplay));
    ReplayDev& d = rp->d;
    d.cap = capacity, d.L = state_len, d.AL = action_len, d.prioritized = prioritized, d.alpha = alpha, d.beta = beta;
    d.itcap = 1;
    while (d.itcap < capacity) d.itcap *= 2;
    rp->device = device;
    size_t cap = (size_t)capacity;
    if (cudaMalloc(&d.obs_t, cap * state_len * 4) != cudaSuccess || cudaMalloc(&d.obs_tp1, cap * state_len * 4) != cudaSuccess ||
        cudaMalloc(&d.action, cap * action_len * 4) != cudaSuccess || cudaMalloc(&d.reward, cap * 4) != cudaSuccess ||
        cudaMalloc(&d.done, cap) != cudaSuccess || cudaMalloc(&d.sum, (size_t)2 * d.itcap * 8) != cudaSuccess ||
        cudaMalloc(&d.mn, (size_t)2 * d.itcap * 8) != cudaSuccess || cudaMalloc(&d.counters, 16) != cudaSuccess ||
        cudaMalloc(&d.max_priority, 8) != cudaSuccess) {
        cudaGetLastError();
        free(rp);
        return rp_fail(nullptr, AGAR_E_NOMEM, "cudaMalloc of the replay storage failed");
    }
    k_rp_init<<<(2 * d.itcap + 255) / 256, 256>>>(d);
    RCU(cudaDeviceSynchronize());
    rp->launches = 1;
    *out = rp;
    return AGAR_OK;
}
extern "C" int agar_replay_destroy(AgarReplay* rp) {
    if (!rp) return AGAR_E_INVALID;
    cudaSetDevice(rp->device);
    ReplayDev& d = rp->d;
    cudaFree(d.obs_t), cudaFree(d.obs_tp1), cudaFree(d.action), cudaFree(d.reward), cudaFree(d.done);
    cudaFree(d.sum), cudaFree(d.mn), cudaFree(d.counters), cudaFree(d.max_priority);
    if (d.pos) cudaFree(d.pos);
    free(rp);
    return AGAR_OK;
}
extern "C" const char* agar_replay_last_error(const AgarReplay* rp) { return rp ? rp->err : g_rp_err; }
extern "C" int64_t agar_replay_launch_count(const AgarReplay* rp) { return rp ? rp->launches : 0; }
static int rp_counter(AgarReplay* rp, int which, void* stream) {
    int v = 0;
    if (cudaSetDevice(rp->device) != cudaSuccess) return AGAR_E_CUDA;
    if (cudaMemcpyAsync(&v, rp->d.counters + which, 4, cudaMemcpyDeviceToHost, (cudaStream_t)stream) != cudaSuccess) return AGAR_E_CUDA;
    if (cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) return AGAR_E_CUDA;
    return v;
}
/* sticky AGAR_RP_ERR_* bits raised on the device since create (sampling from a buffer that is too small, an index outside
 * [0, size), a priority <= 0): where the reference raises / asserts, the kernels skip the offending element and flag it */
extern "C" int agar_replay_error_flags(AgarReplay* rp, void* stream) { return rp ? rp_counter(rp, 3, stream) : AGAR_E_INVALID; }
extern "C" int agar_replay_size(AgarReplay* rp, void* stream) { return rp ? rp_counter(rp, 1, stream) : AGAR_E_INVALID; }
extern "C" int agar_replay_next_idx(AgarReplay* rp, void* stream) { return rp ? rp_counter(rp, 0, stream) : AGAR_E_INVALID; }

extern "C" int agar_replay_add_batch(AgarReplay* rp, const float* obs_t, const float* action, const float* reward,
                                     const float* obs_tp1, const uint8_t* done, const uint8_t* valid, int n, void* stream) {
    if (!rp || !obs_t || !action || !reward || !obs_tp1 || !done || n < 0) return AGAR_E_INVALID;
    if (n == 0) return AGAR_OK;
    RCU(cudaSetDevice(rp->device));
    cudaStream_t s = (cudaStream_t)stream;
    if (n > rp->pos_cap) {
        if (rp->d.pos) RCU(cudaFree(rp->d.pos));
        RCU(cudaMalloc(&rp->d.pos, (size_t)n * 4));
        rp->pos_cap = n;
    }
    k_rp_scan<<<1, 1024, 0, s>>>(rp->d, valid, n);
    k_rp_store<<<(int)(((size_t)n * 32 + 255) / 256), 256, 0, s>>>(rp->d, obs_t, action, reward, obs_tp1, done, n);
    RCU(cudaGetLastError());
    rp->launches += 2;
    if (rp->d.prioritized) return rp_fix_paths(rp, rp->d.pos, n, s);
    return AGAR_OK;
}
extern "C" int agar_replay_gather(AgarReplay* rp, const int32_t* idx_dev, int batch, float* obs_t, float* action, float* reward,
                                  float* obs_tp1, uint8_t* done, void* stream) {
    if (!rp || !idx_dev || batch < 0) return AGAR_E_INVALID;
    if (batch == 0) return AGAR_OK;
    RCU(cudaSetDevice(rp->device));
    k_rp_gather<<<(int)(((size_t)batch * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rp->d, idx_dev, batch, obs_t, action, reward,
                                                                                            obs_tp1, done);
    RCU(cudaGetLastError());
    rp->launches += 1;
    return AGAR_OK;
}
extern "C" int agar_replay_sample_uniform(AgarReplay* rp, const double* u_dev, int batch, int32_t* idx_out, float* obs_t,
                                          float* action, float* reward, float* obs_tp1, uint8_t* done, void* stream) {
    if (!rp || !u_dev || !idx_out || batch < 0) return AGAR_E_INVALID;
    if (batch == 0) return AGAR_OK;
    RCU(cudaSetDevice(rp->device));
    k_rp_uniform_idx<<<(batch + 255) / 256, 256, 0, (cudaStream_t)stream>>>(rp->d, u_dev, batch, idx_out);
    RCU(cudaGetLastError());
    rp->launches += 1;
    return agar_replay_gather(rp, idx_out, batch, obs_t, action, reward, obs_tp1, done, stream);
}
extern "C" int agar_replay_sample_prioritized(AgarReplay* rp, const double* u_dev, int batch, int32_t* idx_out, double* weights_out,
                                              float* obs_t, float* action, float* reward, float* obs_tp1, uint8_t* done,
                                              void* stream) {
    if (!rp || !u_dev || !idx_out || batch < 0) return AGAR_E_INVALID;
    if (!rp->d.prioritized) return rp_fail(rp, AGAR_E_UNSUPPORTED, "buffer was created without priorities");
    if (batch == 0) return AGAR_OK;
    RCU(cudaSetDevice(rp->device));
    k_rp_per_sample<<<(batch + 127) / 128, 128, 0, (cudaStream_t)stream>>>(rp->d, u_dev, batch, idx_out, weights_out);
    RCU(cudaGetLastError());
    rp->launches += 1;
    return agar_replay_gather(rp, idx_out, batch, obs_t, action, reward, obs_tp1, done, stream);
}
extern "C" int agar_replay_update_priorities(AgarReplay* rp, const int32_t* idx_dev, const double* priorities_dev, int batch,
                                             void* stream) {
    if (!rp || !idx_dev || !priorities_dev || batch < 0) return AGAR_E_INVALID;
    if (!rp->d.prioritized || batch == 0) return AGAR_OK; /* ReplayBuffer.update_priorities is a no-op (:69-70) */
    RCU(cudaSetDevice(rp->device));
    cudaStream_t s = (cudaStream_t)stream;
    const int staged = batch <= 12288; /* 48 KB of shared memory */
    k_rp_set_priorities<<<(batch + 127) / 128, 128, staged ? (size_t)batch * 4 : 0, s>>>(rp->d, idx_dev, priorities_dev, batch, staged);
    k_rp_max_priority<<<1, 256, 0, s>>>(rp->d, priorities_dev, batch);
    RCU(cudaGetLastError());
    rp->launches += 2;
    return rp_fix_paths(rp, idx_dev, batch, s);
}
