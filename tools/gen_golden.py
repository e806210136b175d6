"""Generate tests/golden/*.npz by EXECUTING the reference (NILOIDE/A.I.gar, /root/reference/src) through
oracle/ref_harness.py.  Run in the build container only (the reference is absent on the GPU box).

Each fixture holds, for one (config, seed, env id): the float32 actions fed per frame, the env record after
selected frames, the running event hash after every frame, and every observation the NN agents received
(float32, as the API returns them).  tests/test_golden.py replays the actions through the C oracle (bit-exact)
and tests/test_gpu_parity.py through the CUDA path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import aigar_b200.layout as lay
from oracle import ref_harness as rh

CASES = {
    "cfg1_pellet": (dict(), 640, 3, 17, 80),
    "cfg3_1v1": (dict(num_nn=1, num_greedy=1, virus=True, split=True, eject=True), 480, 5, 2, 60),
    "cfg3_random_bot": (dict(num_nn=1, num_greedy=1, num_random=1, virus=True, split=True, eject=True), 320, 7, 1, 80),
    "cfg4_arena": (dict(num_nn=8, num_greedy=8, virus=True, split=True, eject=True), 96, 9, 4, 24),
    # obs_mode = AGAR_OBS_CANONICAL: the FOV grid always has G columns (the harness patches spatialHashTable.__init__)
    "cfg1_pellet_canonical": (dict(obs_mode=1), 640, 3, 17, 80),
    "cfg3_1v1_canonical": (dict(num_nn=1, num_greedy=1, virus=True, split=True, eject=True, obs_mode=1), 480, 5, 2, 60),
    # the "handcraft CNN" observation = the same grid encoder at G = CNN_INPUT_DIM_1 = 42, no extras
    # (src/model/bot.py:103-111,276-282; networkParameters.py:195-208)
    "cnn42_pellet_canonical": (dict(grid=42, obs_mode=1, overrides={"use_fovsize": 0, "use_totalmass": 0}), 320, 11, 3, 80),
    "cnn42_1v1": (dict(grid=42, num_nn=1, num_greedy=1, virus=True, split=True, eject=True), 160, 13, 1, 40),
    "cnn84_pellet_canonical": (dict(grid=84, obs_mode=1, overrides={"use_fovsize": 0, "use_totalmass": 0}), 96, 15, 2, 48),
    # ALL_PLAYER_GRID (networkParameters.py:88-91) instead of the self / enemy channels
    "all_player_1v1_canonical": (dict(num_nn=1, num_greedy=1, virus=True, split=True, eject=True, obs_mode=1,
                                      overrides={"all_player_grid": 1, "self_grid": 0, "enemy_grid": 0, "self_grid_lf": 0,
                                                 "enemy_grid_lf": 0}), 320, 17, 3, 80),
    # GRID_VIEW_ENABLED = False: Bot.getSimpleStateRepresentation (bot.py:511-548), 12 values per observation
    "simple_state_1v1": (dict(num_nn=1, num_greedy=1, virus=True, split=True, eject=True, grid_view=False), 480, 19, 1, 80),
    "simple_state_arena": (dict(num_nn=3, num_greedy=5, virus=True, split=True, eject=True, grid_view=False, frame_skip=3), 160, 23, 2, 40),
}


# LONG rollouts (round 2): the rare paths of the step — split, eject, merge (~400 frames after a split), virus eating /
# explosion, blob eating, virus-eats-blob, blob -> pellet — only occur once cells have grown, hundreds of frames after a
# reset.  Compact format: actions quantised to k/256 (exact in float32) and stored as uint8, turn flags as a bit mask,
# every `obs_stride`-th observation of each agent, a record every `every` frames, the running event hash EVERY frame and
# the tally of the reference's own event log.  tests/golden_util.py reads both formats.
LONG_CASES = {
    # name: (kw, frames, seed, env_id, record every, obs stride)
    "long_cfg1_pellet": (dict(), 1200, 21, 5, 300, 5),
    "long_cfg3_1v1": (dict(num_nn=1, num_greedy=1, virus=True, split=True, eject=True), 4000, 77, 0, 500, 10),
    "long_cfg4_arena": (dict(num_nn=8, num_greedy=8, virus=True, split=True, eject=True), 1600, 59, 0, 400, 8),
}


def generate_long(name, kw, frames, seed, env_id, every, obs_stride):
    cfg = lay.derive_config(event_cap=0, **kw)
    ref = rh.RefEnv(cfg, seed=seed, env_id=env_id)
    L = ref.layout
    A = max(L.n_agents, 1)
    rng = np.random.default_rng(seed * 7919 + env_id)
    actions_u8 = rng.integers(0, 256, size=(frames, A, 4), dtype=np.uint8)
    actions = (actions_u8.astype(np.float32) / np.float32(256.0)).astype(np.float32)
    recs, rec_frames, hashes = [ref.to_record().buf.copy()], [-1], []
    obs_list, obs_idx = [], []
    n_obs = [0] * A
    flags = np.zeros((frames, A), np.uint8)
    tally = np.zeros(16, np.int64)
    for t in range(frames):
        tr = ref.step(actions[t])
        for ev in ref.events:
            tally[ev[0]] += 1
        for a in range(L.n_agents):
            flags[t, a] = (int(tr[a]["observed"]) | int(tr[a]["valid"]) << 1 | int(tr[a]["done"]) << 2 |
                           int(tr[a]["need_action"]) << 3)
            if tr[a]["obs"] is not None:
                if n_obs[a] % obs_stride == 0:
                    obs_list.append(tr[a]["obs"].astype(np.float32))
                    obs_idx.append((t, a))
                n_obs[a] += 1
        hashes.append(ref.event_hash)
        if (t + 1) % every == 0 or t == frames - 1:
            recs.append(ref.to_record(tr).buf.copy())
            rec_frames.append(t)
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", name + ".npz")
    np.savez_compressed(out, fmt=2, kw=np.array(repr(kw)), seed=seed, env_id=env_id, actions_u8=actions_u8,
                        records=np.stack(recs), record_frames=np.array(rec_frames), event_hash=np.array(hashes, dtype=np.uint64),
                        obs=np.stack(obs_list), obs_index=np.array(obs_idx, dtype=np.int32).reshape(-1, 2), flags_u8=flags,
                        tally=tally)
    print(name, os.path.getsize(out), "bytes;", len(recs), "records;", len(obs_list), "observations; events:",
          {lay.EV_NAMES[i]: int(n) for i, n in enumerate(tally) if n})


def generate(name, kw, frames, seed, env_id, every):
    cfg = lay.derive_config(event_cap=0, **kw)
    ref = rh.RefEnv(cfg, seed=seed, env_id=env_id)
    L = ref.layout
    A = max(L.n_agents, 1)
    rng = np.random.default_rng(seed * 7919 + env_id)
    actions = rng.random((frames, A, 4)).astype(np.float32)
    recs, rec_frames, hashes = [ref.to_record().buf.copy()], [-1], []
    obs_list, obs_idx, flags = [], [], []
    for t in range(frames):
        tr = ref.step(actions[t])
        for a in range(L.n_agents):
            flags.append((t, a, int(tr[a]["observed"]), int(tr[a]["valid"]), int(tr[a]["done"]), int(tr[a]["need_action"])))
            if tr[a]["obs"] is not None:
                obs_list.append(tr[a]["obs"].astype(np.float32))
                obs_idx.append((t, a))
        hashes.append(ref.event_hash)
        if (t + 1) % every == 0 or t == frames - 1:
            recs.append(ref.to_record(tr).buf.copy())
            rec_frames.append(t)
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", name + ".npz")
    np.savez_compressed(out, kw=np.array(repr(kw)), seed=seed, env_id=env_id, actions=actions,
                        records=np.stack(recs), record_frames=np.array(rec_frames), event_hash=np.array(hashes, dtype=np.uint64),
                        obs=np.stack(obs_list) if obs_list else np.zeros((0, L.state_len), np.float32),
                        obs_index=np.array(obs_idx, dtype=np.int32).reshape(-1, 2), flags=np.array(flags, dtype=np.int32))
    print(name, os.path.getsize(out), "bytes;", len(recs), "records;", len(obs_list), "observations")


if __name__ == "__main__":
    for name, (kw, frames, seed, env_id, every) in CASES.items():
        if len(sys.argv) > 1 and name not in sys.argv[1:]:
            continue
        generate(name, kw, frames, seed, env_id, every)
    for name, args in LONG_CASES.items():
        if len(sys.argv) > 1 and name not in sys.argv[1:]:
            continue
        generate_long(name, *args)
