import sys
sys.path.insert(0, 'tools'); sys.path.insert(0, '.')
import gpu_check
for kw, n, fr in ((dict(grid=42, overrides={"use_fovsize": 0, "use_totalmass": 0}), 16, 200),
              (dict(grid=63), 8, 120),
              (dict(grid=42, obs_mode=1), 8, 120),
              (dict(grid=20), 8, 120),
              (dict(grid=42, num_nn=1, num_greedy=1, virus=True, split=True, eject=True), 8, 200),
              (dict(grid=42, num_nn=2, num_greedy=2, virus=True, split=True, eject=True, obs_mode=1), 4, 120)):
    ok = gpu_check.check(kw, n_envs=n, frames=fr, verbose=False)
    print(kw, "OK" if ok else "FAILED", flush=True)
