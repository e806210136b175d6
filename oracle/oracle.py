"""ctypes wrapper of the C oracle (oracle/agar_oracle.c) — TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
import ctypes
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

import aigar_b200.layout as lay  # noqa: E402

_libs = {}


def build(force=False):
    """Compile both oracle builds with the committed Makefile (gcc only)."""
    targets = [os.path.join(_HERE, n) for n in ("libagar_oracle.so", "libagar_oracle_pm.so")]
    src = os.path.join(_HERE, "agar_oracle.c")
    deps = [src] + [os.path.join(_ROOT, "include", n) for n in ("agar_b200.h", "agar_layout.h", "agar_math.h", "agar_libm_tables.h")]
    stale = force or any(not os.path.exists(t) or any(os.path.getmtime(t) < os.path.getmtime(d) for d in deps)
                         for t in targets)
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "-B"], stdout=subprocess.DEVNULL)
    return targets


def load(portable=False):
    key = bool(portable)
    if key in _libs:
        return _libs[key]
    build()
    lib = ctypes.CDLL(os.path.join(_HERE, "libagar_oracle_pm.so" if portable else "libagar_oracle.so"))
    vp, u64, i32 = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int
    lib.oracle_layout.argtypes = [ctypes.POINTER(lay.AgarConfig), ctypes.POINTER(lay.AgarLayout)]
    lib.oracle_layout.restype = i32
    lib.oracle_create.argtypes = [ctypes.POINTER(lay.AgarConfig), u64, u64]
    lib.oracle_create.restype = vp
    lib.oracle_destroy.argtypes = [vp]
    lib.oracle_record.argtypes = [vp]
    lib.oracle_record.restype = vp
    lib.oracle_record_bytes.argtypes = [vp]
    lib.oracle_record_bytes.restype = u64
    lib.oracle_load_record.argtypes = [vp, vp]
    lib.oracle_set_key.argtypes = [vp, u64, u64]
    lib.oracle_reset.argtypes = [vp]
    lib.oracle_reset_bots.argtypes = [vp]
    lib.oracle_observe.argtypes = [vp, vp, vp]
    lib.oracle_step.argtypes = [vp, vp, i32]
    lib.oracle_get_turn.argtypes = [vp, vp, vp, vp, vp]
    lib.oracle_field_update.argtypes = [vp]
    lib.oracle_pm_pow.argtypes = [ctypes.c_double, ctypes.c_double]
    lib.oracle_pm_pow.restype = ctypes.c_double
    lib.oracle_pm_pow_mismatches.argtypes = [u64, u64, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
    lib.oracle_pm_pow_mismatches.restype = u64
    lib.oracle_pm_trig_mismatches.argtypes = [u64, u64]
    lib.oracle_pm_trig_mismatches.restype = u64
    for name in ("oracle_pm_sin", "oracle_pm_cos"):
        getattr(lib, name).argtypes = [ctypes.c_double]
        getattr(lib, name).restype = ctypes.c_double
    lib.oracle_pm_atan2.argtypes = [ctypes.c_double, ctypes.c_double]
    lib.oracle_pm_atan2.restype = ctypes.c_double
    lib.oracle_pm_log.argtypes = [ctypes.c_double]
    lib.oracle_pm_log.restype = ctypes.c_double
    lib.oracle_pm_exp.argtypes = [ctypes.c_double]
    lib.oracle_pm_exp.restype = ctypes.c_double
    lib.oracle_pm_dir.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
    lib.oracle_pm_round_dec.argtypes = [ctypes.c_double, ctypes.c_double]
    lib.oracle_pm_round_dec.restype = ctypes.c_double
    lib.oracle_axis_range.argtypes = [ctypes.c_double, ctypes.c_double, i32, ctypes.POINTER(i32), ctypes.POINTER(i32)]
    lib.oracle_rollout_random.argtypes = [vp, i32, i32, ctypes.c_uint32, vp]
    lib.oracle_rollout_batch.argtypes = [ctypes.POINTER(lay.AgarConfig), i32, u64, u64, i32, i32,
                                         ctypes.POINTER(ctypes.c_double)]
    lib.oracle_rollout_batch.restype = u64
    _libs[key] = lib
    return lib


class OracleEnv(object):
    """One env stepped by the C oracle.  Mirrors oracle.ref_harness.RefEnv's driving interface."""

    def __init__(self, cfg, seed=0, env_id=0, portable=False):
        self.lib = load(portable)
        self.cfg = cfg
        self.layout = lay.AgarLayout()
        rc = self.lib.oracle_layout(ctypes.byref(cfg), ctypes.byref(self.layout))
        if rc != 0:
            raise ValueError("oracle_layout failed: %d" % rc)
        self.h = self.lib.oracle_create(ctypes.byref(cfg), seed, env_id)
        if not self.h:
            raise ValueError("oracle_create failed")
        n = int(self.layout.record_bytes)
        self._buf = (ctypes.c_uint8 * n).from_address(self.lib.oracle_record(self.h))
        self.record = lay.Record(self.layout, np.frombuffer(self._buf, dtype=np.uint8))
        a, L = max(self.layout.n_agents, 1), self.layout.state_len
        self.obs = np.zeros((a, L), dtype=np.float32)
        self.obs64 = np.zeros((a, L), dtype=np.float64)

    def __del__(self):
        try:
            if self.h:
                self.lib.oracle_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def reset(self):
        self.lib.oracle_reset(self.h)

    def reset_bots(self):
        self.lib.oracle_reset_bots(self.h)

    def load_record(self, rec):
        buf = np.ascontiguousarray(rec.buf if isinstance(rec, lay.Record) else rec, dtype=np.uint8)
        assert buf.nbytes == int(self.layout.record_bytes)
        self.lib.oracle_load_record(self.h, buf.ctypes.data)

    def set_key(self, seed, env_id):
        self.lib.oracle_set_key(self.h, seed, env_id)

    def observe(self):
        """First half of the NN bots' turn.  Returns per-agent turn info like RefEnv.step()."""
        self.lib.oracle_observe(self.h, self.obs.ctypes.data, self.obs64.ctypes.data)
        return self.turn()

    def turn(self):
        a = self.layout.n_agents
        rew = np.zeros(max(a, 1), dtype=np.float32)
        done = np.zeros(max(a, 1), dtype=np.uint8)
        valid = np.zeros(max(a, 1), dtype=np.uint8)
        need = np.zeros(max(a, 1), dtype=np.uint8)
        self.lib.oracle_get_turn(self.h, rew.ctypes.data, done.ctypes.data, valid.ctypes.data, need.ctypes.data)
        bots = self.record.players["bot"]
        out = []
        for i in range(a):
            observed = not bool(bots["skipping"][i])  # getStateRepresentation() was called (None when dead)
            has = bool(need[i])
            out.append({"observed": observed, "obs": self.obs64[i].copy() if has else None,
                        "obs32": self.obs[i].copy() if has else None,
                        "reward": float(bots["last_reward"][i]), "valid": bool(valid[i]), "done": bool(done[i]),
                        "need_action": bool(need[i])})
        return out

    def step(self, actions=None, n_frames=1):
        a = max(self.layout.n_agents, 1)
        act = np.zeros((a, 4), dtype=np.float32)
        if actions is not None:
            act[:] = np.asarray(actions, dtype=np.float32).reshape(a, 4)
        self.lib.oracle_step(self.h, act.ctypes.data, n_frames)

    def field_update(self):
        self.lib.oracle_field_update(self.h)

    def rollout_random(self, n_decisions, n_frames, decision_base=0):
        self.lib.oracle_rollout_random(self.h, n_decisions, n_frames, decision_base, self.obs.ctypes.data)

    def frame(self, actions=None):
        """observe + one frame, the unit RefEnv.step() performs."""
        t = self.observe()
        self.step(actions, 1)
        return t


def rollout_batch(cfg, n_envs, seed, first_env, n_decisions, n_threads, portable=False):
    lib = load(portable)
    mass = ctypes.c_double(0.0)
    steps = lib.oracle_rollout_batch(ctypes.byref(cfg), n_envs, seed, first_env, n_decisions, n_threads,
                                     ctypes.byref(mass))
    return int(steps), float(mass.value)
