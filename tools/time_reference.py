"""Time the UNPATCHED Python reference (NILOIDE/A.I.gar, /root/reference/src) on this machine's host cores — BASELINE.md §3,
SURVEY §8(d) "CPU baseline timing".  Runs in the build container only (the reference cannot travel to the GPU box);
writes profiles/r02_python_reference.json, which bench.py cites as cpu_baseline.python_reference.

    python -O tools/time_reference.py [seconds per measurement, default 6]

Measured per config (1 = pellet collection, 3 = 1-vs-greedy with viruses / split / eject, 4 = 8 NN + 8 greedy arena), with a
uniform random-action learner stub (the benchmark's driver):
  * Model.update()            frames/s, 1 process         (bots act, then Field.update — model.py:100-112)
  * Field.update() only       frames/s, 1 process         (field.py:85-92; commands frozen)
  * getStateRepresentation()  observations/s, 1 process   (bot.py:272-299)
  * Model.update()            frames/s, multiprocessing.Pool(n_cores) of independent envs, as the reference's own testers
                              run (aigar.py:549-554)
No harness patches are installed: stock numpy RNG, stock set-ordered candidates.  Only matplotlib / pygame are stubbed (they are
imported by model/model.py but never used on the step path) and `python -O` disables tracemalloc as the project's own job
scripts do (submission.sh:13).
"""
import json
import multiprocessing as mp
import os
import platform
import re
import sys
import time
import types

REF_SRC = os.environ.get("AGAR_REF_SRC", "/root/reference/src")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CONFIGS = {
    "1_pellet": dict(NUM_NN_BOTS=1, NUM_GREEDY_BOTS=0, VIRUS_SPAWN=False, ENABLE_SPLIT=False, ENABLE_EJECT=False),
    "3_1v1_greedy": dict(NUM_NN_BOTS=1, NUM_GREEDY_BOTS=1, VIRUS_SPAWN=True, ENABLE_SPLIT=True, ENABLE_EJECT=True),
    "4_arena": dict(NUM_NN_BOTS=8, NUM_GREEDY_BOTS=8, VIRUS_SPAWN=True, ENABLE_SPLIT=True, ENABLE_EJECT=True),
}


def _load():
    for name in ("matplotlib", "matplotlib.pyplot", "pygame", "pygame.gfxdraw"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].use = lambda *a, **k: None
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["pygame"].gfxdraw = sys.modules["pygame.gfxdraw"]
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    import model.model as mm
    return mm


def _params(subst):
    """networkParameters for a config, the reference's way: a text-patched copy executed as a module (aigar.py:270-298)."""
    src = open(os.path.join(REF_SRC, "model", "networkParameters.py")).read()
    for name, val in subst.items():
        src, n = re.subn(r"(?m)^(\s*)%s\s*=.*$" % name, lambda m: "%s%s = %r" % (m.group(1), name, val), src, count=1)
        assert n == 1, name
    mod = types.ModuleType("networkParameters_timed")
    exec(compile(src, "networkParameters_timed.py", "exec"), mod.__dict__)
    mod.GATHER_EXP = True
    return mod


class _RandomAlg(object):
    """Learner stub: uniform random actions (bot.py:127,189,206,224 are the only members the bot touches)."""
    discrete = False

    def __init__(self, n):
        import random
        self.n, self.rnd = n, random.Random(1)

    def __repr__(self):
        return "Stub"

    def reset(self):
        pass

    def decideMove(self, state, updateNoise=True):
        return None, [self.rnd.random() for _ in range(self.n)]


def make_model(cfg_name):
    mm = _load()
    subst = CONFIGS[cfg_name]
    params = _params(subst)
    model = mm.Model(False, False, params)
    n_act = 2 + int(subst["ENABLE_SPLIT"]) + int(subst["ENABLE_EJECT"])
    for _ in range(subst["NUM_NN_BOTS"]):
        model.createBot("NN", _RandomAlg(n_act), params)
    for _ in range(subst["NUM_GREEDY_BOTS"]):
        model.createBot("Greedy", None, params)
    model.initialize()
    return model


def time_update(cfg_name, seconds, warm=200):
    model = make_model(cfg_name)
    for _ in range(warm):
        model.update()
    n, t0 = 0, time.perf_counter()
    while True:
        for _ in range(50):
            model.update()
        n += 50
        dt = time.perf_counter() - t0
        if dt >= seconds:
            return n / dt


def time_field_only(cfg_name, seconds, warm=200):
    model = make_model(cfg_name)
    for _ in range(warm):
        model.update()
    f = model.field
    n, t0 = 0, time.perf_counter()
    while True:
        for _ in range(50):
            f.update()
        n += 50
        dt = time.perf_counter() - t0
        if dt >= seconds:
            return n / dt


def time_obs_only(cfg_name, seconds, warm=200):
    model = make_model(cfg_name)
    for _ in range(warm):
        model.update()
    bots = [b for b in model.getNNBots() if b.player.getIsAlive()]
    n, t0 = 0, time.perf_counter()
    while True:
        for b in bots:
            b.getStateRepresentation()
        n += len(bots)
        dt = time.perf_counter() - t0
        if dt >= seconds:
            return n / dt


def _pool_job(args):
    cfg_name, seconds = args
    return time_update(cfg_name, seconds)


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 6.0
    cores = len(os.sched_getaffinity(0))
    cpu = "unknown"
    for line in open("/proc/cpuinfo"):
        if line.startswith("model name"):
            cpu = line.split(":", 1)[1].strip()
            break
    out = {"what": "unpatched Python reference timed on this container's host cores (tools/time_reference.py)",
           "cpu_model": cpu, "cores": cores, "python": platform.python_version(), "optimized": not __debug__,
           "seconds_per_measurement": seconds, "configs": {}}
    for name in CONFIGS:
        r = {"model_update_frames_per_s_1core": time_update(name, seconds),
             "field_update_only_frames_per_s_1core": time_field_only(name, seconds),
             "get_state_representation_obs_per_s_1core": time_obs_only(name, seconds)}
        with mp.Pool(cores) as pool:
            rates = pool.map(_pool_job, [(name, seconds)] * cores)
        r["model_update_frames_per_s_pool"] = sum(rates)
        r["pool_processes"] = cores
        out["configs"][name] = r
        print(name, json.dumps(r), flush=True)
    path = os.path.join(ROOT, "profiles", "r02_python_reference.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
