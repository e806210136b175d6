"""The C-ABI library loads (no GPU needed) and exports every function include/agar_b200.h declares."""
import ctypes
import os
import re

import pytest

import aigar_b200.layout as lay

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    names = set()
    for header in ("agar_b200.h", "agar_replay.h"):
        src = open(os.path.join(ROOT, "include", header)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(agar_[a-z_]+)\s*\(", src))
    return sorted(names)


def test_header_declares_the_surface():
    names = declared_functions()
    for n in ("agar_create", "agar_reset", "agar_step", "agar_observe", "agar_get", "agar_step_host", "agar_debug_dump"):
        assert n in names


def test_library_exports_every_declared_symbol():
    from aigar_b200 import env
    lib = env.load_library()
    for name in declared_functions():
        assert hasattr(lib, name), name


def test_layout_entry_point_needs_no_gpu():
    from aigar_b200 import env
    lib = env.load_library()
    cfg = lay.derive_config(num_nn=1, num_greedy=1, virus=True, split=True, eject=True)
    out = lay.AgarLayout()
    assert lib.agar_layout_for_config(ctypes.byref(cfg), ctypes.byref(out)) == 0
    assert out.as_dict() == lay.layout_for_config(cfg).as_dict()
    bad = lay.derive_config(eject=True)
    assert lib.agar_layout_for_config(ctypes.byref(bad), ctypes.byref(out)) == -4


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly, never route through oracle/."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from aigar_b200.env import AgarBatch, AgarError
    with pytest.raises(AgarError):
        AgarBatch(lay.derive_config(), 4)
    src = open(os.path.join(ROOT, "a.i.gar_b200", "env.py")).read() + open(os.path.join(ROOT, "a.i.gar_b200", "__init__.py")).read()
    assert "oracle" not in src.replace("oracle/", "").lower() or "import oracle" not in src
