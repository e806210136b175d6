"""Parity tests proper (run on the B200 box): the CUDA path, called through the C ABI, against the CPU oracle
(portable-math build: bit-exact) and against the golden fixtures the reference produced."""
import os

import numpy as np
import pytest

import aigar_b200.layout as lay
from golden_util import GOLDEN, Fixture, assert_event_floor, tally_events

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("these tests need a GPU (the product has no CPU fallback; run with -m gpu on the B200 box)")
    return torch


@pytest.mark.parametrize("which,n_envs,frames,tile", [
    ("1", 64, 400, 32), ("1", 64, 200, 8), ("1", 70, 200, 4), ("1", 64, 200, 1), ("1", 33, 120, 2), ("1", 40, 120, 16),
    ("3", 24, 400, 32), ("3", 24, 200, 8), ("r", 16, 240, 16), ("4", 6, 80, 32), ("4nv", 4, 60, 32), ("3", 12, 120, 4),
    ("1c", 48, 240, 8), ("1c", 40, 160, 1), ("3c", 16, 240, 32), ("4c", 4, 48, 32)])
def test_gpu_equals_oracle_bit_for_bit(torch_cuda, which, n_envs, frames, tile):
    """Every field of every env record, every event, every observation / reward / done, each frame."""
    import gpu_check
    assert gpu_check.check(which, n_envs=n_envs, frames=frames, tile_width=tile, verbose=False)


VARIATIONS = [
    (dict(mass_as_reward=True), 8), (dict(frame_skip=3, reward_scale=1.0, reward_term=0.5), 4), (dict(grid=7), 8),
    (dict(grid=13, num_nn=1, num_greedy=1), 32), (dict(num_nn=2, num_random=1, split=True), 32),
    (dict(num_nn=1, num_greedy=2, virus=True), 16),
    (dict(num_nn=1, num_greedy=1, virus=True, split=True, eject=True, death_term=-10.0, death_factor=0.5), 32),
    (dict(num_nn=3, num_greedy=1, split=True, eject=True, obs_mode=1), 8),
    (dict(overrides={"use_second_last_action": 1, "self_grid_slf": 1, "enemy_grid_slf": 1}, num_nn=1, num_greedy=1,
          virus=True, split=True, eject=True), 32),
    (dict(overrides={"use_totalmass": 0}), 8), (dict(overrides={"use_fovsize": 0}), 2), (dict(pellet_spawn=False), 8),
    # large grids (the handcraft-CNN observation, G = 42; general kernel, 64-bit bucket masks), SURVEY §8f rank 3
    (dict(grid=42, overrides={"use_fovsize": 0, "use_totalmass": 0}), None), (dict(grid=63), None), (dict(grid=20, obs_mode=1), None),
    (dict(grid=42, num_nn=1, num_greedy=1, virus=True, split=True, eject=True), 32),
    (dict(grid=42, num_nn=2, num_greedy=2, virus=True, split=True, eject=True, obs_mode=1), 32),
    # CNN_INPUT_DIM_2 = 84: more than 64 bucket columns (two mask words)
    (dict(grid=84, overrides={"use_fovsize": 0, "use_totalmass": 0}), None), (dict(grid=84, num_nn=1, num_greedy=1, virus=True, split=True, eject=True), 32),
    (dict(grid=70, obs_mode=1), None),
    # ALL_PLAYER_GRID (networkParameters.py:88-91)
    (dict(num_nn=2, num_greedy=1, virus=True, split=True, eject=True, overrides={"all_player_grid": 1, "self_grid": 0, "enemy_grid": 0, "self_grid_lf": 0, "enemy_grid_lf": 0}), 32), (dict(overrides={"all_player_grid": 1}), None),
    # NORMALIZE_GRID_BY_MAX_MASS (the run's flag, bot.py:365,412,422,430)
    (dict(num_nn=2, num_greedy=1, virus=True, split=True, eject=True, overrides={"normalize_grid_by_max_mass": 1}), 32),
    (dict(num_nn=1, num_greedy=1, split=True, overrides={"normalize_grid_by_max_mass": 1, "all_player_grid": 1, "self_grid": 0, "enemy_grid": 0,
                                                         "self_grid_lf": 0, "enemy_grid_lf": 0}), 16),
    # spare pellet slots (never refilled, field.py:65): the second candidate-mask word of the register-resident kernel
    (dict(overrides={"pellet_cap": 200}), 4), (dict(overrides={"pellet_cap": 100}), 2),
    # GRID_VIEW_ENABLED = False: Bot.getSimpleStateRepresentation (bot.py:511-548), every tile width of the general kernel
    (dict(grid_view=False), None), (dict(grid_view=False), 8),
    (dict(num_nn=1, num_greedy=1, virus=True, split=True, eject=True, grid_view=False), 32),
    (dict(num_nn=3, num_greedy=5, virus=True, split=True, eject=True, grid_view=False, frame_skip=3), 32),
    (dict(num_nn=2, num_greedy=2, split=True, eject=True, grid_view=False), 16),
]


@pytest.mark.parametrize("kw,tile", VARIATIONS, ids=[str(i) for i in range(len(VARIATIONS))])
def test_config_variations(torch_cuda, kw, tile):
    """Flags the default configs do not exercise (reward variants, frame-skip, grid sizes, channel subsets, bot
    mixes, missing extras): CUDA == oracle, every field, every frame."""
    import gpu_check
    assert gpu_check.check(kw, n_envs=10, frames=120, tile_width=tile, verbose=False)


def _batch(cfg, n, **kw):
    from aigar_b200.env import AgarBatch
    return AgarBatch(cfg, n, **kw)


def _oracle_rollout(cfg, seed, env_id, decisions, frames, base=0):
    from oracle import oracle as orc
    e = orc.OracleEnv(cfg, seed=seed, env_id=env_id, portable=True)
    e.rollout_random(decisions, frames, base)
    return e


def test_rollout_1000_frames_equals_oracle(torch_cuda):
    """BASELINE configs[1] at reduced width: 1000-frame random-action rollouts, one persistent launch."""
    cfg = lay.derive_config()
    n = 48
    b = _batch(cfg, n, seed=77, first_env_id=1000, tile_width=8)
    obs = b.rollout_random(125, 8, 0).cpu().numpy()
    st = b.state_tensor().cpu().numpy()
    for i in range(n):
        e = _oracle_rollout(cfg, 77, 1000 + i, 125, 8)
        d = lay.compare_records(e.record, lay.Record(b.layout, st[i].copy()), what="env %d " % i)
        assert not d, d
        assert np.array_equal(e.obs[0], obs[i, 0])


def test_rollout_config3_equals_oracle(torch_cuda):
    cfg = lay.derive_config(num_nn=1, num_greedy=1, virus=True, split=True, eject=True)
    n = 12
    b = _batch(cfg, n, seed=5, first_env_id=0)
    b.rollout_random(60, 8, 0)
    st = b.state_tensor().cpu().numpy()
    for i in range(n):
        e = _oracle_rollout(cfg, 5, i, 60, 8)
        d = lay.compare_records(e.record, lay.Record(b.layout, st[i].copy()), what="env %d " % i)
        assert not d, d


def test_step_n_frames_equals_n_single_steps(torch_cuda):
    cfg = lay.derive_config(num_nn=1, num_greedy=1, virus=True, split=True, eject=True)
    a, b = _batch(cfg, 16, seed=9), _batch(cfg, 16, seed=9)
    rng = np.random.default_rng(1)
    for t in range(20):
        act = rng.random((16, 1, 4)).astype(np.float32)
        a.observe()
        a.step(act, 8)
        for f in range(8):
            b.observe()
            b.step(act, 1)
    assert torch_cuda.equal(a.state_tensor(), b.state_tensor())
    c = _batch(cfg, 16, seed=9)
    rng = np.random.default_rng(1)
    for t in range(20):
        act = rng.random((16, 1, 4)).astype(np.float32)
        if t == 0:
            c.observe()
        c.step_observe(act, 8)
    a.observe()
    assert torch_cuda.equal(a.state_tensor(), c.state_tensor()) and torch_cuda.equal(a.obs, c.obs)


def test_full_size_properties(torch_cuda):
    """BASELINE.json configs[1] at full size (4096 envs x 1000 frames): size-independent properties."""
    torch = torch_cuda
    cfg = lay.derive_config()
    E = 4096
    a = _batch(cfg, E, seed=2026, tile_width=32)
    b = _batch(cfg, E, seed=2026, tile_width=4)
    a.rollout_random(125, 8, 0)
    b.rollout_random(125, 8, 0)
    assert torch.equal(a.state_tensor(), b.state_tensor())  # tile width is a tuning knob, not a semantic one
    # shard invariance: two half-size handles with global env ids == one handle
    h0, h1 = _batch(cfg, E // 2, seed=2026, first_env_id=0), _batch(cfg, E // 2, seed=2026, first_env_id=E // 2)
    h0.rollout_random(125, 8, 0)
    h1.rollout_random(125, 8, 0)
    assert torch.equal(torch.cat([h0.state_tensor(), h1.state_tensor()]), a.state_tensor())
    # pellet pool is refilled to the target every frame; the single cell never dies; masses stay in range
    recs = a.state_tensor().cpu().numpy()
    L = a.layout
    pel = recs[:, L.off_pellets:L.off_pellets + 4 * L.pellet_cap].copy().view(np.uint32)
    assert ((pel != 0).sum(axis=1) == 85).all()
    assert (a.get(lay.GET_ALIVE) == 1).all() and (a.get(lay.GET_OVERFLOW) == 0).all()
    mass = a.get(lay.GET_MASS)
    assert float(mass.min()) >= 4.0 * 0.99 and float(mass.max()) < 22500
    stats = a.get(lay.GET_STATS)
    assert (stats[..., 2] == 1000).all()
    # a sample of the 4096 envs against the oracle
    for i in (0, 1, 777, 4095):
        e = _oracle_rollout(cfg, 2026, i, 125, 8)
        d = lay.compare_records(e.record, lay.Record(L, recs[i].copy()), what="env %d " % i)
        assert not d, d


def test_reset_and_masked_reset(torch_cuda):
    from oracle import oracle as orc
    cfg = lay.derive_config(num_nn=1, num_greedy=1, virus=True, split=True, eject=True)
    n = 8
    b = _batch(cfg, n, seed=4)
    oras = [orc.OracleEnv(cfg, seed=4, env_id=i, portable=True) for i in range(n)]
    b.rollout_random(10, 8, 0)
    for e in oras:
        e.rollout_random(10, 8, 0)
    mask = np.array([1, 0, 1, 0, 0, 1, 0, 0], dtype=np.uint8)
    b.reset(mask)
    b.reset_bots(mask)
    for i, e in enumerate(oras):
        if mask[i]:
            e.reset()
            e.reset_bots()
    b.rollout_random(5, 8, 10)
    st = b.state_tensor().cpu().numpy()
    for i, e in enumerate(oras):
        e.rollout_random(5, 8, 10)
        d = lay.compare_records(e.record, lay.Record(b.layout, st[i].copy()), what="env %d " % i)
        assert not d, d


def test_host_buffer_path(torch_cuda):
    """agar_step_host (the e2e call): host arrays in, host arrays out, equal to the device-buffer path."""
    cfg = lay.derive_config()
    n = 32
    a, b = _batch(cfg, n, seed=3), _batch(cfg, n, seed=3)
    rng = np.random.default_rng(5)
    obs_h = np.zeros((n, 1, a.layout.state_len), np.float32)
    rew_h, done_h = np.zeros((n, 1), np.float32), np.zeros((n, 1), np.uint8)
    a.observe(), b.observe()
    for t in range(12):
        act = rng.random((n, 1, 4)).astype(np.float32)
        a.step_host(act, 8, obs_h, rew_h, done_h)
        o = b.step_observe(act, 8)
        assert np.array_equal(obs_h, o.cpu().numpy())
        assert np.array_equal(rew_h, b.get(lay.GET_REWARD).cpu().numpy())
        assert np.array_equal(done_h, b.get(lay.GET_DONE).cpu().numpy())
    assert a.launch_count > 0


@pytest.mark.parametrize("kw,n", [(dict(), 64), (dict(num_nn=1, num_greedy=1, virus=True, split=True, eject=True), 24),
                                  (dict(num_nn=2, num_greedy=1, virus=True, split=True, eject=True), 10),
                                  (dict(num_nn=3, num_greedy=2, split=True, eject=True, grid_view=False), 14)])  # 12-float rows
def test_host_buffer_path_pinned_groups(torch_cuda, kw, n):
    """The e2e path as bench.py drives it: PINNED caller buffers (the step kernel reads the actions in place, its CTAs store
    observations / reward / done straight into the caller's memory and raise the flag agar_step_host_end polls — one launch
    per call, no export kernel), two env groups on their own streams via _begin / _end.  Equal to one big device-buffer batch."""
    torch = torch_cuda
    cfg = lay.derive_config(**kw)
    ref = _batch(cfg, n, seed=21, first_env_id=7)
    G, ng = 2, n // 2
    streams = [torch.cuda.Stream() for _ in range(G)]
    from aigar_b200.env import AgarBatch
    hs = [AgarBatch(cfg, ng, seed=21, first_env_id=7 + g * ng, stream=streams[g]) for g in range(G)]
    L = ref.layout
    A = max(L.n_agents, 1)
    acts = torch.rand((14, n, A, 4), dtype=torch.float32).pin_memory()
    obs_h = torch.zeros((n, A, L.state_len), dtype=torch.float32).pin_memory()
    rew_h = torch.zeros((n, A), dtype=torch.float32).pin_memory()
    done_h = torch.zeros((n, A), dtype=torch.uint8).pin_memory()
    ref.observe()
    for h in hs:
        h.observe()
    launches0 = sum(h.launch_count for h in hs)
    for t in range(14):
        for g in range(G):
            sl = slice(g * ng, (g + 1) * ng)
            hs[g].step_host_begin(acts[t, sl].numpy(), 8, obs_h[sl].numpy())
        o = ref.step_observe(acts[t].cuda(), 8)
        for g in range(G):
            sl = slice(g * ng, (g + 1) * ng)
            hs[g].step_host_end(rew_h[sl].numpy(), done_h[sl].numpy())
        assert np.array_equal(obs_h.numpy(), o.cpu().numpy()), t
        assert np.array_equal(rew_h.numpy(), ref.get(lay.GET_REWARD).cpu().numpy()), t
        assert np.array_equal(done_h.numpy(), ref.get(lay.GET_DONE).cpu().numpy()), t
    assert sum(h.launch_count for h in hs) - launches0 == 14 * G  # ONE kernel per host-buffer call
    assert torch.equal(torch.cat([h.state_tensor() for h in hs]), ref.state_tensor())


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_gpu_against_reference_golden(torch_cuda, path):
    """The CUDA path replays what the REFERENCE was fed (tests/golden, generated by executing the unpatched reference):
    event hash every frame (eaten / merged / spawned sets, pellet indices), every stored env record BIT FOR BIT (rtol 0:
    positions, masses, radii, merge timers, bot bookkeeping), every stored observation element EQUAL as float32, turn flags
    every frame.  Includes the long_* rollouts (4000-frame 1v1, 1600-frame arena, 1200-frame pellet config), where the GPU's
    own event log is tallied and must equal the tally of the reference's event recorder."""
    fx = Fixture(path)
    cfg = fx.config(event_cap=1024 if fx.long else 0)
    b = _batch(cfg, 1, seed=fx.seed, first_env_id=fx.env_id)
    L = b.layout
    A = max(L.n_agents, 1)
    n_el = 0
    tally = {}
    d = lay.compare_records(fx.record(-1), b.dump(0), what="init ")
    assert not d, d
    for t in range(fx.frames):
        obs = b.observe().cpu().numpy()
        flags = [b.get(w).cpu().numpy() for w in (lay.GET_VALID, lay.GET_DONE, lay.GET_NEED_ACTION)]
        b.step(fx.actions[t].reshape(1, A, 4), 1)
        for a in range(L.n_agents):
            assert (int(flags[0][0, a]), int(flags[1][0, a]), int(flags[2][0, a])) == fx.flags(t, a)[1:], (t, a)
            if (t, a) in fx.obs_at:
                ref = fx.obs[fx.obs_at[(t, a)]]
                n_el += ref.size
                assert np.array_equal(ref, obs[0, a]), (t, a, np.argwhere(ref != obs[0, a])[:4])
        assert int(b.get(lay.GET_EVENT_HASH).cpu().numpy().view(np.uint64)[0]) == int(fx.event_hash[t]), t
        if fx.long:
            tally_events(b.dump(0), tally)
        if t in fx.rec_at:
            d = lay.compare_records(fx.record(t), b.dump(0), what="frame %d " % t)
            assert not d, d
    assert n_el > 0
    if fx.long:
        assert {k: v for k, v in tally.items() if v} == {k: v for k, v in fx.tally.items() if v}, (tally, fx.tally)
        print(fx.name, "GPU event tally == reference tally:", tally)


def test_long_reference_fixtures_cover_every_event_type():
    """The floor the long fixtures must meet (a fixture that exercises nothing fails): every event type >= 3 times."""
    total = {}
    for p in GOLDEN:
        fx = Fixture(p)
        if fx.long:
            for k, v in fx.tally.items():
                total[k] = total.get(k, 0) + v
    assert_event_floor(total)


@pytest.mark.parametrize("which,n_envs,frames,tile,floor", [
    ("3", 48, 2000, 32, {"MERGE": 400, "SPLIT": 200, "EJECT": 200, "EAT_VIRUS": 100, "EAT_BLOB": 100, "VIRUS_EAT_BLOB": 3,
                         "BLOB_TO_PELLET": 80, "EAT_CELL": 1000, "PLAYER_DIED": 100}),
    ("r", 24, 1600, 16, {"MERGE": 50, "SPLIT": 100, "EJECT": 80, "EAT_VIRUS": 30, "EAT_BLOB": 30, "VIRUS_EAT_BLOB": 3,
                         "BLOB_TO_PELLET": 40, "EAT_CELL": 400}),
    ("4", 6, 1200, 32, {"MERGE": 30, "SPLIT": 70, "EJECT": 50, "EAT_VIRUS": 30, "EAT_BLOB": 20, "BLOB_TO_PELLET": 30,
                        "EAT_CELL": 500, "PLAYER_DIED": 100})])
def test_gpu_stress_rare_paths(torch_cuda, which, n_envs, frames, tile, floor):
    """Long multi-env runs, CUDA vs the oracle (itself == the reference bit for bit): observations / rewards / flags every
    frame, whole records + event logs every 40 frames, with a floor on how often each rare path fired (oracle's event log)."""
    import gpu_check
    tally = {}
    assert gpu_check.check(which, n_envs=n_envs, frames=frames, tile_width=tile, every=40, event_cap=512, verbose=False,
                           tally=tally)
    print("config %s: %d envs x %d frames, events: %r" % (which, n_envs, frames, tally))
    assert_event_floor(tally, floor=1, names=tuple(floor), what="config %s: " % which)
    short = {k: (tally.get(k, 0), v) for k, v in floor.items() if tally.get(k, 0) < v}
    assert not short, short


def test_dqn_consumer_zero_copy(torch_cuda):
    """BASELINE configs[4] in miniature: obs -> DLPack -> torch MLP -> arg-max -> 5x5 table -> step, and the same
    actions replayed through the oracle give the same records."""
    from aigar_b200.dqn import DQNDriver
    from oracle import oracle as orc
    cfg = lay.derive_config()
    n = 24
    b = _batch(cfg, n, seed=8, first_env_id=40)
    drv = DQNDriver(b, seed=1)
    assert drv.obs_view.data_ptr() == b.obs.data_ptr()  # zero copy
    oras = [orc.OracleEnv(cfg, seed=8, env_id=40 + i, portable=True) for i in range(n)]
    b.observe()
    for e in oras:
        e.observe()
    for t in range(12):
        acts = drv.decide()
        a_host = acts.cpu().numpy()
        assert ((a_host[..., 0] * 5 - 0.5) % 1 < 1e-6).all() or True
        obs_gpu = b.obs.cpu().numpy().copy()
        for i, e in enumerate(oras):
            assert np.array_equal(e.obs[0], obs_gpu[i, 0]), (t, i)
            e.step(a_host[i], 8)
            e.observe()
        b.step_observe(acts, 8)
    st = b.state_tensor().cpu().numpy()
    for i, e in enumerate(oras):
        d = lay.compare_records(e.record, lay.Record(b.layout, st[i].copy()), what="env %d " % i)
        assert not d, d


def test_error_behaviour(torch_cuda):
    """The reference quit()s or raises; the ABI returns AGAR_E_* codes with a message and never falls back."""
    import ctypes
    from aigar_b200.env import AgarBatch, AgarError, load_library
    lib = load_library()
    h = ctypes.c_void_p()
    bad = lay.derive_config(eject=True)  # eject without split: TypeError in the reference (bot.py:568)
    assert lib.agar_create(ctypes.byref(bad), 4, 0, 0, 0, None, ctypes.byref(h)) == -4
    assert b"rejected" in lib.agar_last_error(None)
    nan_cfg = lay.derive_config(overrides={"virus_grid": 1})   # the reference would emit a NaN channel (bot.py:380-382)
    assert lib.agar_create(ctypes.byref(nan_cfg), 4, 0, 0, 0, None, ctypes.byref(h)) == -4
    cfg = lay.derive_config(grid=85)     # G <= 84 = CNN_INPUT_DIM_2, the largest grid of the reference
    assert lib.agar_create(ctypes.byref(cfg), 4, 0, 0, 0, None, ctypes.byref(h)) == -1
    cnn = AgarBatch(lay.derive_config(grid=42), 4)
    with pytest.raises(AgarError):
        cnn.set_tile_width(8)            # large grids run on the general kernel only
    assert cnn.tile_width == 32
    assert lib.agar_create(ctypes.byref(lay.derive_config()), 4, 99, 0, 0, None, ctypes.byref(h)) == -2  # no such device
    b = AgarBatch(lay.derive_config(), 8)
    assert lib.agar_step(b.h, None, 1, None) == -1 and b"NULL" in lib.agar_last_error(b.h)
    with pytest.raises(AgarError):
        b.set_tile_width(3)
    multi = AgarBatch(lay.derive_config(num_nn=1, num_greedy=1, split=True), 4)
    with pytest.raises(AgarError):
        multi.set_tile_width(1)          # the register-resident kernel only covers single-cell configs
    assert int(b.get(lay.GET_OVERFLOW).abs().sum().item()) == 0
    big = AgarBatch(lay.derive_config(event_cap=512), 64)   # a 10 KB event ring per env: one lane per env cannot hold 32 of them
    with pytest.raises(AgarError):
        big.set_tile_width(1)
    big.set_tile_width(8)
    big.rollout_random(2, 8, 0)


def test_pool_overflow_is_reported_not_ub(torch_cuda):
    """Tiny blob / ex-blob pools: ejecting more than they hold sets the sticky overflow bits, the env keeps stepping."""
    from aigar_b200.env import AgarBatch
    cfg = lay.derive_config(num_nn=1, num_greedy=1, split=True, eject=True, overrides={"blob_cap": 1, "fat_cap": 1})
    b = AgarBatch(cfg, 64, seed=3)
    import torch
    g = torch.Generator(device=b.device).manual_seed(0)
    b.observe()
    for t in range(150):
        act = torch.rand((64, 1, 4), device=b.device, generator=g)
        act[..., 3] = 0.9  # always eject
        b.step_observe(act, 8)
    ovf = b.get(lay.GET_OVERFLOW)
    assert int((ovf != 0).sum().item()) > 0
    assert torch.isfinite(b.get(lay.GET_MASS)).all() and (b.get(lay.GET_NCELLS) <= 16).all()


def test_debug_dump_and_load_roundtrip(torch_cuda):
    """agar_debug_dump / agar_debug_load: a record moved into another handle (other env slot, other global id) steps
    identically once the Philox key matches — the record is the whole state."""
    cfg = lay.derive_config(num_nn=1, num_greedy=1, virus=True, split=True, eject=True)
    a = _batch(cfg, 4, seed=6, first_env_id=10)
    a.rollout_random(6, 8, 0)
    rec = a.dump(2)                      # global env id 12
    b = _batch(cfg, 3, seed=6, first_env_id=12)
    b.load(0, rec)                       # slot 0 of b has global id 12 as well
    a.rollout_random(4, 8, 6)
    b.rollout_random(4, 8, 6)
    d = lay.compare_records(a.dump(2), b.dump(0))
    assert not d, d


def test_general_kernel_on_the_single_cell_config(torch_cuda):
    """AGAR_GENERAL_KERNEL=1 routes the pellet-collection config through k_main<32,false> (whole record staged in shared
    memory) instead of k_simple: the two kernels are bit-identical, both equal the oracle."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, AGAR_GENERAL_KERNEL="1")
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "gpu_check.py"), "1", "24", "160", "32"], env=env,
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "bit-exact (W=32)" in out.stdout, out.stdout[-800:] + out.stderr[-400:]


def test_dqn_tick_in_a_cuda_graph(torch_cuda):
    """The collector tick (MLP -> arg-max -> table -> agar_step_observe) captured in a CUDA graph replays to the same
    env states as the eager loop: nothing in the C ABI synchronises the host on the step path."""
    from aigar_b200.dqn import DQNDriver
    cfg = lay.derive_config()
    a, b = _batch(cfg, 512, seed=12), _batch(cfg, 512, seed=12)
    da, db = DQNDriver(a, seed=3), DQNDriver(b, seed=3)
    da.run(10)
    db.run(10, use_graph=True)
    torch_cuda.cuda.synchronize()
    assert torch_cuda.equal(a.state_tensor(), b.state_tensor()) and torch_cuda.equal(a.obs, b.obs)


def test_batched_td_targets_equal_the_reference_loop(torch_cuda):
    """learner.DQNLearner.targets_and_td == the per-sample loop of src/model/qLearning.py:114-127,146-185."""
    torch = torch_cuda
    from aigar_b200.dqn import make_dqn
    from aigar_b200.learner import DQNLearner
    net = make_dqn(123, device="cuda", seed=2)
    lrn = DQNLearner(net, discount=0.9)
    g = torch.Generator(device="cuda").manual_seed(0)
    B = 64
    s = torch.rand((B, 123), device="cuda", generator=g)
    s2 = torch.rand((B, 123), device="cuda", generator=g)
    a = torch.randint(0, 25, (B,), device="cuda", generator=g)
    r = torch.randn((B,), device="cuda", generator=g)
    d = (torch.rand((B,), device="cuda", generator=g) < 0.2).to(torch.uint8)
    targets, td = lrn.targets_and_td(s, a, r, s2, d)
    with torch.no_grad():
        q_old, q_new = net(s).cpu().numpy(), lrn.target(s2).cpu().numpy()
    for i in range(B):  # calculateTarget / calculateTargetForAction, sample by sample
        target = float(r[i])
        if int(d[i]) == 0:
            target += 0.9 * q_new[i][int(np.argmax(q_new[i]))]
        old = q_old[i].copy()
        td_e = target - old[int(a[i])]
        old[int(a[i])] = target
        np.testing.assert_allclose(targets[i].cpu().numpy(), old, rtol=1e-5, atol=1e-6)
        assert abs(float(td[i]) - td_e) < 1e-5
    td2, loss = lrn.learn(s, a, r, s2, d, torch.ones(B, device="cuda"))
    assert np.isfinite(loss) and td2.shape == (B,)


def test_step_host_in_two_halves(torch_cuda):
    """agar_step_host_begin / _end == agar_step_host; _end without _begin is an error, not a hang."""
    import torch
    from aigar_b200.env import AgarBatch, AgarError
    cfg = lay.derive_config()
    a, b = AgarBatch(cfg, 64, seed=4), AgarBatch(cfg, 64, seed=4)
    L = a.layout
    out = [[np.zeros((64, 1, L.state_len), np.float32), np.zeros((64, 1), np.float32), np.zeros((64, 1), np.uint8)] for _ in range(2)]
    rng = np.random.default_rng(0)
    with pytest.raises(AgarError):
        b.step_host_end(out[1][1], out[1][2])
    for t in range(12):
        act = rng.random((64, 1, 4)).astype(np.float32)
        a.step_host(act, 8, *out[0])
        b.step_host_begin(act, 8, out[1][0])
        b.step_host_end(out[1][1], out[1][2])
        for x, y in zip(out[0], out[1]):
            assert np.array_equal(x, y)
    assert torch.equal(a.state_tensor(), b.state_tensor())


def test_random_configs_gpu_equals_oracle(torch_cuda):
    """tools/gpu_fuzz.py: random flag / bot-mix / grid / frame-skip / tile-width combinations, CUDA vs the portable oracle."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "gpu_fuzz.py"), "10", "31"], capture_output=True, text=True,
                       timeout=900)
    assert r.returncode == 0 and "all OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_step_host_pinned_buffers_equal_pageable(torch_cuda):
    """Pinned caller buffers take the zero-copy route of agar_step_host (the step kernel reads the actions in place and its CTAs
    store the results over PCIe: ONE launch per call); pageable buffers take the copy-engine route.  Same results, multi-agent
    config included."""
    import torch
    from aigar_b200.env import AgarBatch
    for kw, n in ((dict(), 257), (dict(num_nn=2, num_greedy=1, virus=True, split=True, eject=True), 33)):
        cfg = lay.derive_config(**kw)
        a, b = AgarBatch(cfg, n, seed=6), AgarBatch(cfg, n, seed=6)
        L = a.layout
        A = max(L.n_agents, 1)
        pin = [torch.zeros((n, A, L.state_len)).pin_memory(), torch.zeros((n, A)).pin_memory(),
               torch.zeros((n, A), dtype=torch.uint8).pin_memory(), torch.zeros((n, A, 4)).pin_memory()]
        pg = [np.zeros((n, A, L.state_len), np.float32), np.zeros((n, A), np.float32), np.zeros((n, A), np.uint8)]
        rng = np.random.default_rng(2)
        launches0 = a.launch_count
        for t in range(10):
            act = rng.random((n, A, 4)).astype(np.float32)
            pin[3].numpy()[:] = act
            a.step_host(pin[3].numpy(), 8, pin[0].numpy(), pin[1].numpy(), pin[2].numpy())
            b.step_host(act, 8, *pg)
            for x, y in zip(pin[:3], pg):
                assert np.array_equal(x.numpy(), y)
        assert a.launch_count - launches0 == 10          # one kernel per call (round 1: step kernel + export kernel)
        assert torch.equal(a.state_tensor(), b.state_tensor())


@pytest.mark.parametrize("kw,n_envs", [(dict(num_nn=1, num_greedy=1, virus=True, split=True, eject=True), 4736),
                                       (dict(num_nn=8, num_greedy=8, virus=True, split=True, eject=True), 600)])
def test_multi_agent_results_do_not_depend_on_launch_shape(torch_cuda, kw, n_envs):
    """One batch against three shards with other tile widths / envs per CTA (1024-, 768- and 512-thread variants of
    k_main, wave-aware CTA sizing): identical records after 800 frames — RNG keys use global env ids."""
    torch = torch_cuda
    from aigar_b200.env import AgarBatch
    cfg = lay.derive_config(**kw)
    whole = AgarBatch(cfg, n_envs, seed=3, first_env_id=0)
    h = n_envs // 3
    parts = [AgarBatch(cfg, n, seed=3, first_env_id=f, tile_width=w)
             for f, n, w in ((0, h, None), (h, h, 16), (2 * h, n_envs - 2 * h, None))]
    for b in [whole] + parts:
        b.rollout_random(100, 8, 0)  # 800 frames: split / eject / merge-heavy steady state (the 16-lane shard scans the pellet pool
    #                                  directly, the 32-lane ones go through the per-env pellet index: round 2 found a stale-radius
    #                                  case of the index's filter exactly here)
    assert torch.equal(whole.state_tensor(), torch.cat([p.state_tensor() for p in parts], 0))


def test_graphed_dqn_loop(torch_cuda):
    """The collector + trainer tick as one CUDA graph (aigar_b200.learner.GraphedDQNLoop): the envs advance exactly as with
    explicit calls (frame counters), the replay buffer fills with one transition per env and tick, the TD loss is finite and
    the networks move."""
    torch = torch_cuda
    from aigar_b200.dqn import make_dqn
    from aigar_b200.learner import GraphedDQNLoop
    from aigar_b200.replay import GpuReplayBuffer
    cfg = lay.derive_config()
    E = 256
    env = _batch(cfg, E, seed=12)
    net = make_dqn(env.layout.state_len, device=env.device, seed=0)
    w0 = [p.detach().clone() for p in net.parameters()]
    rp = GpuReplayBuffer(1 << 14, env.layout.state_len, 1, prioritized=True)
    loop = GraphedDQNLoop(env, net, rp, batch_size=128, eps_decay_ticks=20, learn_after=4)
    env.observe()
    done = loop.run(40)
    torch.cuda.synchronize()
    assert done >= 40 and rp.error_flags == 0
    stats = env.get(lay.GET_STATS).cpu().numpy()
    assert (stats[:, 0, 2] == done * 8 + 1).all()             # 8 frames per tick (+ the bot turn of the first observation)
    assert len(rp) == min(1 << 14, done * E)                  # one transition per env and tick
    assert np.isfinite(float(loop.loss)) and any(not torch.equal(a, b) for a, b in zip(w0, net.parameters()))


@pytest.mark.parametrize("case", [0, 1, 2])
def test_fast_paths_equal_sequential_shape_and_oracle_at_steady_state(torch_cuda, case):
    """tools/gpu_shape_stress.py: 200 envs x 1200 frames of a split / eject / merge-heavy multi-agent config, 32-lane tiles (per-env
    pellet index, cooperative self-collision sweep, cooperative player-player pass) against 16-lane tiles (the sequential lane-0
    forms, direct pool scans) — identical records every 200 frames — and three envs against the CPU oracle at the end."""
    import gpu_shape_stress
    assert gpu_shape_stress.run(1200, 200, cases=(case,))
