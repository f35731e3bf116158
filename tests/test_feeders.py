"""CPU tier: this package's feeder generators against the reference's, field by field.

``tests/golden/feeders.npz`` was written by ``oracle/ref_harness.py feeders`` from the UNMODIFIED
reference generators (``/root/reference/grid_fed_rl/feeders/ieee_feeders.py:22-378``,
``synthetic.py:23-252``, ``base.py:256-303``): every bus / line / load / generator field the hot path
reads, raw (as generated, after ``np.random.seed(0)`` for IEEE-34 / IEEE-123: deviation D4-i) and the
line list after ``repair_topology`` (D4-ii, iii).  "Topology ordering bit-exact" (BASELINE north_star)
starts here: same ids in the same positions, same impedances to the last bit."""
import json
import os

import numpy as np
import pytest

from oracle.ref_harness import FEEDER_SPECS, feeder_table, raw_feeder

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "feeders.npz")


@pytest.fixture(scope="module")
def frozen():
    z = np.load(GOLDEN, allow_pickle=False)
    return {k: z[k] for k in z.files}


def test_every_spec_is_frozen(frozen):
    specs = sorted({k.split("/")[0] for k in frozen})
    assert specs == sorted(FEEDER_SPECS)
    # the BASELINE configs' feeders are among them
    for need in ("ieee13", "ieee34", "ieee123", "synthetic1000:1000"):
        assert need in specs


@pytest.mark.parametrize("spec", FEEDER_SPECS)
def test_generator_equals_reference(frozen, spec):
    import grid_fed_rl_b200 as m
    raw = raw_feeder(None, spec, use_reference_classes=False)
    mine = feeder_table(raw)
    for k, v in mine.items():
        ref = frozen[f"{spec}/raw/{k}"]
        if k == "generators":
            assert json.loads(str(v)) == json.loads(str(ref)), f"{spec}: generators differ"
            continue
        assert v.shape == ref.shape, f"{spec}: {k} has {v.shape}, the reference {ref.shape}"
        assert np.array_equal(v, ref), f"{spec}: {k} differs from the reference generator"
    fixed = m.repair_topology(raw, keep_cycles=spec.startswith("mesh"))
    rep = feeder_table(fixed)
    for k in ("line_id", "line_from", "line_to", "line_r", "line_x", "line_rating"):
        assert np.array_equal(rep[k], frozen[f"{spec}/repaired/{k}"]), f"{spec}: repaired {k}"


def test_bench_feeders_compile_to_the_published_sizes():
    """IEEE-13 D=71 A=3, IEEE-34 D=164 A=2, IEEE-123 D=692 A=8, synthetic-1000 D=6320 A=403."""
    import grid_fed_rl_b200 as m
    from grid_fed_rl_b200.topology import compile_for_solver
    want = {"ieee13": (13, 71, 3, ["solar", "wind"], 2), "ieee34": (34, 164, 2, ["solar"], 2),
            "ieee123": (123, 692, 8, ["solar", "wind"], 8), "synthetic1000:1000": (1000, 6320, 403, ["solar", "wind"], 64)}
    for spec, (n, D, A, srcs, lanes) in want.items():
        f = m.repair_topology(raw_feeder(None, spec, use_reference_classes=False))
        soa, used = compile_for_solver(f, "newton", 0, renewable_sources=srcs)
        assert (soa.n_bus, soa.obs_dim, soa.act_dim) == (n, D, A), spec
        # ONE lane rule (gfr_auto_lanes in the native library; topology.auto_lanes calls it)
        assert used == lanes and soa.lanes_hint == lanes, (spec, used)
        assert int(soa.level_ptr[1:].max() - 0) == n
        assert max(int(b - a) for a, b in zip(soa.level_ptr[:-1], soa.level_ptr[1:])) <= max(lanes, 1)


def test_auto_lanes_rule_is_the_native_one():
    from grid_fed_rl_b200 import _native as nat
    from grid_fed_rl_b200.topology import auto_lanes
    lib = nat.load_library()
    for n in (2, 13, 20, 21, 34, 45, 46, 90, 123, 160, 161, 250, 400, 401, 1000, 1500, 1501, 5000):
        for solver, code in (("newton", nat.SOLVER_NEWTON), ("sweep", nat.SOLVER_SWEEP)):
            for depth in (0, 3, 9, 12):
                assert auto_lanes(n, solver, depth) == lib.gfr_auto_lanes(n, code, depth)
    assert auto_lanes(13, "newton", 4) == 2 and auto_lanes(34, "newton", 12) == 2
    assert auto_lanes(30, "newton", 7) == 4 and auto_lanes(13, "sweep", 4) == 1
    assert auto_lanes(123, "newton") == 8 and auto_lanes(123, "sweep") == 16 and auto_lanes(1000) == 64
