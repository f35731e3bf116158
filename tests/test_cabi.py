"""CPU tier: the C-ABI library loads and exports every symbol include/gfr_b200.h declares, the
ctypes structures have the C layout, and the product path fails loudly without a GPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gfr_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gfr_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from grid_fed_rl_b200 import build
    from grid_fed_rl_b200 import _native as nat
    build.build()
    return nat.load_library()


def test_header_symbols_are_exported(lib):
    names = _declared()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} is declared in gfr_b200.h but not exported"


def test_binding_covers_header(lib):
    from grid_fed_rl_b200 import _native as nat
    assert sorted(nat.SIGNATURES) == _declared()
    assert lib.gfr_abi_version() == 1


def test_struct_layouts():
    from grid_fed_rl_b200 import _native as nat
    # sizes of the C structs on LP64 (int32 x5 + pad, double, 32 pointers)
    assert C.sizeof(nat.FeederDesc) == 32 + 8 + 40 * 8
    assert C.sizeof(nat.SolverCfg) == 32
    assert C.sizeof(nat.EnvCfg) == 8 + 16 + 6 * 8 + 32 + 8
    assert C.sizeof(nat.StepOut) == 15 * 8
    assert C.sizeof(nat.SolOut) == 8 * 8


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import grid_fed_rl_b200 as m
    from grid_fed_rl_b200 import _native as nat
    with pytest.raises(nat.NativeLibraryMissing):
        m.BatchedGridEnvironment(m.IEEE13Bus(), 4)
    # straight through the C ABI: creating a feeder needs a device
    soa = m.compile_feeder(m.repair_topology(m.IEEE13Bus()), renewable_sources=["solar", "wind"])
    desc, keep = nat.make_feeder_desc(soa)
    h = C.c_void_p()
    rc = lib.gfr_feeder_create(C.byref(desc), 0, C.byref(h))
    assert rc == nat.GFR_E_CUDA and not h.value
    assert b"CPU fallback" in lib.gfr_last_error() or b"cuda" in lib.gfr_last_error().lower()


def test_bad_descriptions_are_rejected(lib):
    import numpy as np
    import grid_fed_rl_b200 as m
    from grid_fed_rl_b200 import _native as nat
    soa = m.compile_feeder(m.SimpleRadialFeeder(5))
    soa.parent = soa.parent.copy()
    soa.parent[3] = 4                      # a parent that does not precede its child
    desc, keep = nat.make_feeder_desc(soa)
    h = C.c_void_p()
    assert lib.gfr_feeder_create(C.byref(desc), 0, C.byref(h)) == nat.GFR_E_ARG
    assert b"parent" in lib.gfr_last_error()
