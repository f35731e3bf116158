"""The numpy oracle (oracle/port.py) against the frozen outputs of the reference itself
(tests/golden/*.npz, written by oracle/ref_harness.py).  This is what pins the oracle."""
import numpy as np
import pytest

from oracle import port
from tests.golden_util import (TOL_PU, feeder_for, golden_names, load_golden, replay_trace)


class _OneEnv:
    def __init__(self, feeder, kw):
        self.env = port.PortEnv(feeder, 1, **kw)

    def reset(self, noise4, start_time):
        return self.env.reset(np.asarray(noise4)[None, :], start_time=start_time)[0]

    def step(self, action, noise):
        out = self.env.step(action[None, :], noise[None, :])
        return {k: v[0] for k, v in out.items()}


@pytest.mark.parametrize("name", golden_names("trace_") + golden_names("meshtrace_"))
def test_port_env_matches_reference_trace(name):
    g = load_golden(name)
    exact = replay_trace(_OneEnv, g, ctx=name)
    assert exact >= 0.9 * g["obs"].shape[0]


@pytest.mark.parametrize("name", golden_names("solve_") + golden_names("meshsolve_"))
def test_port_solver_matches_reference(name):
    g = load_golden(name)
    f = feeder_for(g)
    net = port.DenseNetwork(f.buses, f.lines)
    tol, max_it = float(g["meta"][0]), int(g["meta"][1])
    sol = port.newton_raphson(net, g["p_spec"], tol, max_it)
    assert np.array_equal(sol["converged"], g["converged"])
    conv = g["converged"]
    assert np.all(np.abs(sol["iterations"].astype(int) - g["iterations"]) <= 1)
    if conv.any():
        for k in ("bus_voltages", "bus_angles", "line_flows", "losses"):
            assert np.max(np.abs(sol[k][conv] - g[k][conv])) <= TOL_PU, k
        assert np.allclose(sol["line_loadings"][conv], g["line_loadings"][conv], rtol=1e-7, atol=1e-14)
    # non-converged (divergent) cases: only the flags / iteration count are meaningful


def test_known_answer_appendix_a():
    """the reference's 3-bus fixture, loads {2: 0.1, 3: 0.05} pu, tol 1e-10."""
    from oracle.ref_harness import make_feeder
    f = make_feeder(None, "fixture3", use_reference_classes=False)
    net = port.DenseNetwork(f.buses, f.lines)
    sol = port.newton_raphson(net, np.array([[0.0, -0.1, -0.05]]), 1e-10, 50)
    assert sol["converged"][0] and sol["iterations"][0] == 4
    assert np.allclose(sol["bus_voltages"][0], [1, 0.998491585126075, 0.997739099627788], atol=1e-12)
    assert np.allclose(sol["bus_angles"][0], [0, -0.003004662358732, -0.004259387863609], atol=1e-12)
    assert np.allclose(sol["line_flows"][0], [0.15026346387586, 0.05003767014433], atol=1e-12)
    assert abs(sol["losses"][0] - 2.6346387586110437e-04) < 1e-13
    sol6 = port.newton_raphson(net, np.array([[0.0, -0.1, -0.05]]), 1e-6, 50)
    assert sol6["converged"][0] and sol6["iterations"][0] == 3
    # as shipped (no D2) the reference does not converge
    bad = port.newton_raphson(net, np.array([[0.0, -0.1, -0.05]]), 1e-6, 50, j11_fix=False)
    assert not bad["converged"][0] and bad["iterations"][0] == 50
