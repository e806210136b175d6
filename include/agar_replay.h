/*
 * agar_replay.h — C ABI of the GPU replay buffer (the step right after the env path).
 *
 * Replaces src/model/replay_buffer.py (ReplayBuffer :7-73, PrioritizedReplayBuffer :76-209) and
 * src/model/common/segment_tree.py for batches of transitions that never leave the GPU: the env's observation /
 * reward / done buffers (agar_b200.h) are appended straight into the ring.  Same conventions as agar_b200.h.
 *
 * Randomness is the caller's: every sampling entry takes uniforms in [0, 1) (the reference draws them with the
 * stdlib `random` module, replay_buffer.py:66,116), so results are reproducible and comparable with the reference.
 */
#ifndef AGAR_REPLAY_H
#define AGAR_REPLAY_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct AgarReplay AgarReplay; /* opaque */

/* ReplayBuffer(size) / PrioritizedReplayBuffer(size, alpha, beta) — replay_buffer.py:8-19,77-106 */
int agar_replay_create(int capacity, int state_len, int action_len, int prioritized, double alpha, double beta, int device,
                       AgarReplay** out);
int agar_replay_destroy(AgarReplay* rp);
const char* agar_replay_last_error(const AgarReplay* rp);
/* len(buffer) and _next_idx (synchronise `stream`) */
int agar_replay_size(AgarReplay* rp, void* stream);
int agar_replay_next_idx(AgarReplay* rp, void* stream);

/* ReplayBuffer.add (:24-31) / PrioritizedReplayBuffer.add (:108-113) for the n candidate transitions of one tick:
 * those with valid[i] != 0 (an experience was emitted, bot.py:204-217) are appended in index order.  All pointers
 * are device pointers: obs_t, obs_tp1 float[n][state_len], action float[n][action_len], reward float[n],
 * done uint8[n], valid uint8[n] (NULL = all valid). */
int agar_replay_add_batch(AgarReplay* rp, const float* obs_t, const float* action, const float* reward,
                          const float* obs_tp1, const uint8_t* done, const uint8_t* valid, int n, void* stream);

/* ReplayBuffer._encode_sample (:33-44): gather the transitions idx[0..batch) */
int agar_replay_gather(AgarReplay* rp, const int32_t* idx_dev, int batch, float* obs_t, float* action, float* reward,
                       float* obs_tp1, uint8_t* done, void* stream);
/* ReplayBuffer.sample (:46-67): idx[i] = randint(0, len - 1) drawn as floor(u[i] * len), then gather */
int agar_replay_sample_uniform(AgarReplay* rp, const double* u_dev, int batch, int32_t* idx_out, float* obs_t, float* action,
                               float* reward, float* obs_tp1, uint8_t* done, void* stream);
/* PrioritizedReplayBuffer.sample (:113-171): mass = u[i] * sum(0, len - 1), idx = find_prefixsum_idx(mass);
 * weights[i] = ((p_i * len) ** -beta) / max_weight */
int agar_replay_sample_prioritized(AgarReplay* rp, const double* u_dev, int batch, int32_t* idx_out, double* weights_out,
                                   float* obs_t, float* action, float* reward, float* obs_tp1, uint8_t* done, void* stream);
/* PrioritizedReplayBuffer.update_priorities (:173-195) */
int agar_replay_update_priorities(AgarReplay* rp, const int32_t* idx_dev, const double* priorities_dev, int batch, void* stream);
/* Sticky error bits raised ON THE DEVICE since create (synchronises `stream`).  Where the reference raises — sampling an empty
 * buffer (random.randint(0, -1), replay_buffer.py:66; the prioritized path recurses without end for len < 2, :116) — or asserts
 * (0 <= idx < len(storage), priority > 0, :203-204), the kernels skip the offending element (index 0 / weight 0 is returned,
 * trees are left untouched) and set a bit here instead of hanging or corrupting memory. */
enum { AGAR_RP_ERR_EMPTY = 1, AGAR_RP_ERR_INDEX = 2, AGAR_RP_ERR_PRIORITY = 4 };
int agar_replay_error_flags(AgarReplay* rp, void* stream);
/* number of kernels launched so far */
int64_t agar_replay_launch_count(const AgarReplay* rp);

#ifdef __cplusplus
}
#endif
#endif /* AGAR_REPLAY_H */
