"""include/agar_math.h (the arithmetic of the CUDA kernels) against libm / Python on the game's domain."""
import ctypes
import math

import numpy as np

from oracle import oracle as orc


def _ulps(a, b):
    return abs(a - b) / np.spacing(abs(b))


def test_pow_is_bit_identical_to_libm():
    """agar_pow restates glibc's pow operation by operation on glibc's tables (include/agar_math.h): the bits CPython's
    math.pow returns, on 1e8 pseudo-random inputs of the path's domain + 1.44e6 masses / radii on a 1/64 grid."""
    lib = orc.load(portable=True)
    x, y = ctypes.c_double(), ctypes.c_double()
    bad = lib.oracle_pm_pow_mismatches(100_000_000, 1, ctypes.byref(x), ctypes.byref(y))
    assert bad == 0, (bad, x.value.hex(), y.value)
    rng = np.random.default_rng(0)
    for v in np.concatenate([rng.uniform(0.5, 90, 2000), [1.0, 10.0, 22500.0, 4.0]]):  # through Python's own math.pow
        for e in (-0.35, 0.475, 0.32, 2.0):
            assert lib.oracle_pm_pow(float(v), e) == math.pow(float(v), e)
    assert lib.oracle_pm_pow(7.0, 0.0) == 1.0


def test_atan2_sin_cos_are_bit_identical_to_libm():
    """agar_atan2 / agar_sin / agar_cos restate glibc's e_atan2.c / s_sin.c operation by operation on glibc's tables: the
    bits math.atan2 / math.sin / math.cos return, on 1e8 direction vectors + angles and the whole 601 x 601 integer lattice."""
    lib = orc.load(portable=True)
    assert lib.oracle_pm_trig_mismatches(100_000_000, 7) == 0
    rng = np.random.default_rng(1)
    c, s = ctypes.c_double(), ctypes.c_double()
    for _ in range(5000):  # through Python's own math module, as cell.py:49-57 calls it
        dx, dy = rng.uniform(-300, 300, 2)
        a = math.atan2(dy, dx)
        assert lib.oracle_pm_atan2(dy, dx) == a
        lib.oracle_pm_dir(dy, dx, ctypes.byref(c), ctypes.byref(s))
        assert c.value == math.cos(a) and s.value == math.sin(a)
    for y, x in ((0.0, 0.0), (0.0, -1.0), (-0.0, -1.0), (-0.0, 2.0), (1.0, 0.0), (-1.0, 0.0), (3.0, 3.0), (1e-300, 1e300),
                 (5.0, -1e-17), (-7.0, 7.0)):
        assert lib.oracle_pm_atan2(y, x) == math.atan2(y, x) and math.copysign(1, lib.oracle_pm_atan2(y, x)) == math.copysign(1, math.atan2(y, x))
    for a in (0.0, 0.126, -0.126, 0.855469, 2.426265, math.pi, -math.pi, math.pi / 2, 1e-9, 2.0 ** -27, 3.0, -2.5):
        for d in (-2, -1, 0, 1, 2):
            v = a
            for _ in range(abs(d)):
                v = math.nextafter(v, math.inf if d > 0 else -math.inf)
            assert lib.oracle_pm_sin(v) == math.sin(v) and lib.oracle_pm_cos(v) == math.cos(v), v
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
