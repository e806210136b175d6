"""include/agar_math.h (the arithmetic of the CUDA kernels) against libm / Python on the game's domain."""
import ctypes
import math

import numpy as np

from oracle import oracle as orc


def _ulps(a, b):
    return abs(a - b) / np.spacing(abs(b))


def test_pow_within_16_ulps():
    lib = orc.load(portable=True)
    rng = np.random.default_rng(0)
    worst = 0.0
    xs = np.concatenate([rng.uniform(0.5, 90, 4000), rng.uniform(1, 22500, 4000), [1.0, 10.0, 22500.0, 4.0]])
    for x in xs:
        for y in (-0.35, 0.475, 0.32):
            worst = max(worst, _ulps(lib.oracle_pm_pow(float(x), y), math.pow(float(x), y)))
    assert worst <= 16.0, worst  # y*log(x) carries ~1 ulp of |y log x| <= 5


def test_direction_equals_cos_sin_of_atan2_within_ulps():
    lib = orc.load(portable=True)
    rng = np.random.default_rng(1)
    c, s = ctypes.c_double(), ctypes.c_double()
    for _ in range(5000):
        dx, dy = rng.uniform(-300, 300, 2)
        lib.oracle_pm_dir(dy, dx, ctypes.byref(c), ctypes.byref(s))
        a = math.atan2(dy, dx)
        assert abs(c.value - math.cos(a)) < 1e-15 and abs(s.value - math.sin(a)) < 1e-15
    lib.oracle_pm_dir(0.0, 0.0, ctypes.byref(c), ctypes.byref(s))
    assert (c.value, s.value) == (1.0, 0.0)  # atan2(0, 0) == 0


def test_round_dec_is_pythons_round():
    lib = orc.load(portable=True)
    rng = np.random.default_rng(2)
    vals = list(rng.uniform(-2, 3, 20000)) + [0.015625, 0.5, 0.000005, 1.0000050000000001, 0.3456750000000001,
                                              2.675, 1.005, 0.1 + 0.2, 1 / 64 + 1e-18, 0.999995, -0.015625]
    for v in vals:
        assert lib.oracle_pm_round_dec(float(v), 1e5) == round(float(v), 5), v
        assert lib.oracle_pm_round_dec(float(v), 1e3) == round(float(v), 3), v
