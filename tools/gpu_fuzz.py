"""Random configurations, GPU vs portable oracle, every field every frame (run on a GPU box): python tools/gpu_fuzz.py [N] [SEED]"""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import aigar_b200.layout as lay
import gpu_check

N = int(sys.argv[1]) if len(sys.argv) > 1 else 24
rnd = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
FLAGS = ["pellet_grid", "self_grid", "wall_grid", "enemy_grid", "virus_grid", "self_grid_lf", "self_grid_slf", "enemy_grid_lf",
         "enemy_grid_slf", "use_fovsize", "use_last_fovsize", "use_totalmass", "use_last_action", "use_second_last_action"]
ok_all, ran = True, 0
for case in range(N):
    n_nn = rnd.choice([1, 1, 1, 2, 3])
    n_gr = rnd.choice([0, 0, 1, 1, 2, 4])
    n_rd = rnd.choice([0, 0, 0, 1])
    split = rnd.random() < 0.5
    eject = split and rnd.random() < 0.6          # eject without split is rejected (bot.py:568)
    kw = dict(num_nn=n_nn, num_greedy=n_gr, num_random=n_rd, virus=rnd.random() < 0.5, split=split, eject=eject,
              grid=rnd.choice([3, 5, 8, 11, 11, 14, 16, 19, 25]), frame_skip=rnd.choice([0, 1, 3, 7, 7, 9]),
              obs_mode=rnd.choice([0, 0, 1]), mass_as_reward=rnd.random() < 0.2,
              overrides={f: int(rnd.random() < 0.5) for f in rnd.sample(FLAGS, rnd.randint(0, 6))})
    if rnd.random() < 0.15:  # ALL_PLAYER_GRID replaces the self / enemy channels (networkParameters.py:88-91)
        kw["overrides"].update({"all_player_grid": 1, "self_grid": 0, "enemy_grid": 0, "self_grid_lf": 0, "self_grid_slf": 0,
                                "enemy_grid_lf": 0, "enemy_grid_slf": 0})
    cfg = lay.derive_config(**kw)
    try:
        lay.layout_for_config(cfg)
    except Exception as ex:
        print("case %d rejected by layout (%s): %r" % (case, ex, kw))
        continue
    tile = rnd.choice([None, None, 32, 8, 4, 16]) if n_nn + n_gr + n_rd > 1 or split or kw["virus"] else rnd.choice([None, 1, 2, 4, 8, 16, 32])
    try:
        ok = gpu_check.check(kw, n_envs=rnd.choice([3, 7, 12]), frames=rnd.choice([40, 80, 120]), seed=rnd.randint(0, 999),
                             tile_width=tile, verbose=False)
    except Exception as ex:
        msg = str(ex)
        if "rejected" in msg or "unsupported" in msg.lower() or "tile width" in msg:
            print("case %d not runnable (%s): %r tile %r" % (case, msg[:80], kw, tile))
            continue
        raise
    ran += 1
    print("case %d %s: %r tile %r" % (case, "OK" if ok else "FAILED", kw, tile), flush=True)
    ok_all = ok_all and ok
print("ran %d cases: %s" % (ran, "all OK" if ok_all else "FAILURES"))
sys.exit(0 if ok_all else 1)
