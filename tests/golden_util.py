"""Shared helpers: load a golden trace, rebuild its feeder with this package's own classes
(the reference is not present on the GPU box) and replay it through an env implementation."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# comparisons: V/theta/flows within 1e-8 pu (BASELINE north_star); flags and counts bit-exact
# except on steps whose margin to a threshold is below FLAG_MARGIN (SURVEY 8c caveat)
TOL_PU = 1e-8
FLAG_MARGIN = 1e-9


def golden_names(prefix):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, prefix + "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def feeder_for(g):
    """Feeder of a golden file, built from grid_fed_rl_b200's own generators + repair (D4)."""
    from oracle.ref_harness import make_feeder
    f = make_feeder(None, str(g["spec"]), use_reference_classes=False)
    if "meta" in g and g["meta"].size >= 10 and g["meta"][9] != 1.0:
        s = float(g["meta"][9])
        for ld in f.loads:
            ld.base_power *= s
            ld.active_power *= s
            ld.reactive_power *= s
    return f


def trace_kwargs(g):
    m = g["meta"]
    return dict(timestep=float(m[2]), episode_length=int(m[3]), tolerance=float(m[5]),
                max_iterations=int(m[6]), stochastic_loads=bool(m[7]), weather_variation=bool(m[8]),
                renewable_sources=[str(s) for s in g["renewable_sources"]]), float(m[4])


def obs_layout(n, m, L, G, Bt):
    """Slices of the observation vector (reference grid_env.py:753-783)."""
    o = 0
    out = {}
    out["vm"] = slice(o, o + 2 * n, 2); out["va"] = slice(o + 1, o + 2 * n, 2); o += 2 * n
    out["p"] = slice(o, o + 2 * m, 2); out["loading"] = slice(o + 1, o + 2 * m, 2); o += 2 * m
    out["freq"] = o; o += 1
    out["loads"] = slice(o, o + 2 * L); o += 2 * L
    out["gen"] = slice(o, o + G); o += G
    out["soc"] = slice(o, o + 2 * Bt, 2); out["bpow"] = slice(o + 1, o + 2 * Bt, 2)
    return out


def compare_step(got, g, t, lay, s_base, ctx="", check_iterations=True):
    """got: dict of per-env arrays (env 0 compared) vs golden step t."""
    def chk(a, b, tol, what):
        a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
        err = np.max(np.abs(a - b)) if a.size else 0.0
        assert err <= tol, f"{ctx} step {t}: {what} differs by {err:.3e} (tol {tol:.1e})"
    o, ref = got["obs"], g["obs"][t]
    chk(o[lay["vm"]], ref[lay["vm"]], TOL_PU, "Vm")
    chk(o[lay["va"]], ref[lay["va"]], TOL_PU, "Va")
    chk(o[lay["p"]] / s_base, ref[lay["p"]] / s_base, TOL_PU, "line P (pu)")
    chk(o[lay["loading"]], ref[lay["loading"]], 1e-7, "line loading")
    chk(o[lay["freq"]], ref[lay["freq"]], 1e-9, "frequency")
    chk(o[lay["loads"]], ref[lay["loads"]], 0.0, "static loads")
    chk(o[lay["gen"]], ref[lay["gen"]], 1e-6, "generation (W)")
    chk(o[lay["soc"]], ref[lay["soc"]], 1e-12, "soc")
    chk(o[lay["bpow"]], ref[lay["bpow"]], 1e-6, "battery power (W)")
    chk(got["losses"] / s_base, g["losses"][t] / s_base, TOL_PU, "losses (pu)")
    rtol = 1e-9 * max(1.0, abs(float(g["reward"][t])))
    chk(got["reward"], g["reward"][t], rtol + 1e-6, "reward")
    chk(got["episode_reward"], g["episode_reward"][t], 1e-9 * max(1.0, abs(float(g["episode_reward"][t]))) + 1e-5,
        "episode reward")
    assert bool(got["terminated"]) == bool(g["terminated"][t]), f"{ctx} step {t}: terminated"
    assert bool(got["error"]) == bool(g["error"][t]), f"{ctx} step {t}: error flag"
    assert bool(got["converged"]) == bool(g["converged"][t]), f"{ctx} step {t}: converged"
    if check_iterations:
        assert abs(int(got["iterations"]) - int(g["iterations"][t])) <= 1, f"{ctx} step {t}: iterations"
    assert int(got["current_step"]) == int(g["current_step"][t]), f"{ctx} step {t}: current_step"
    # threshold margins
    vm = ref[lay["vm"]]
    margin = min(np.min(np.abs(vm - 0.95)), np.min(np.abs(vm - 1.05)),
                 abs(ref[lay["freq"]] - 59.5), abs(ref[lay["freq"]] - 60.5))
    if margin >= FLAG_MARGIN:
        assert list(map(bool, got["violations"])) == list(map(bool, g["violations"][t])), \
            f"{ctx} step {t}: violation flags"
        assert int(got["viol_count"]) == int(g["viol_count"][t]), f"{ctx} step {t}: violation count"
        assert bool(got["truncated"]) == bool(g["truncated"][t]), f"{ctx} step {t}: truncated"
        return True
    return False


def replay_trace(env_factory, g, ctx="", check_iterations=True):
    """env_factory(feeder, kwargs) -> object with reset(noise4, start_time) -> obs[D] and
    step(action[A], noise[4+L]) -> dict of scalars/arrays for one env.  Returns #steps compared
    with exact flags."""
    f = feeder_for(g)
    kw, start_time = trace_kwargs(g)
    env = env_factory(f, kw)
    s_base = f.parameters.base_power * 1e6
    n, m, L = len(f.buses), len(f.lines), len(f.loads)
    D = g["obs"].shape[1]
    G_Bt = D - (2 * n + 2 * m + 1 + 2 * L)
    A = g["actions"].shape[1]
    Bt = G_Bt - A          # D-part = G + 2Bt, A = G + Bt
    G = A - Bt
    lay = obs_layout(n, m, L, G, Bt)
    ep = 0
    obs0 = env.reset(g["reset_noise"][ep], start_time)
    assert np.max(np.abs(np.asarray(obs0) - g["reset_obs"][ep])) <= 1e-9, f"{ctx}: reset obs"
    exact = 0
    for t in range(g["obs"].shape[0]):
        if g["reset_before"][t]:
            ep += 1
            obs0 = env.reset(g["reset_noise"][ep], start_time)
            assert np.max(np.abs(np.asarray(obs0) - g["reset_obs"][ep])) <= 1e-9, f"{ctx}: reset obs {ep}"
        got = env.step(g["actions"][t], g["noise"][t])
        exact += bool(compare_step(got, g, t, lay, s_base, ctx, check_iterations))
    return exact


def port_trace(g, tolerance=None, max_iterations=None):
    """Re-run a golden trace's inputs (actions, noise, reset draws) through the numpy oracle,
    optionally at another solver tolerance, and return a dict shaped like the golden file.
    Used where the compared solver is a different algorithm (sweep), so both sides run tight."""
    from oracle import port
    f = feeder_for(g)
    kw, start_time = trace_kwargs(g)
    if tolerance is not None:
        kw["tolerance"] = tolerance
    if max_iterations is not None:
        kw["max_iterations"] = max_iterations
    env = port.PortEnv(f, 1, **kw)
    T = g["obs"].shape[0]
    keys = ("obs", "reward", "terminated", "truncated", "error", "converged", "iterations",
            "max_voltage", "min_voltage", "losses", "violations", "viol_count", "current_step",
            "episode_reward")
    rec = {k: [] for k in keys}
    rec["reset_before"], rec["reset_obs"] = [], []
    ep = 0
    rec["reset_obs"].append(env.reset(g["reset_noise"][ep][None, :], start_time=start_time)[0])
    need = False
    for t in range(T):
        rec["reset_before"].append(need)
        if need:
            ep += 1
            rec["reset_obs"].append(env.reset(g["reset_noise"][ep][None, :], start_time=start_time)[0])
        out = env.step(g["actions"][t][None, :], g["noise"][t][None, :])
        for k in keys:
            rec[k].append(out[k][0])
        need = bool(out["terminated"][0] or out["truncated"][0])
    res = {k: np.array(v) for k, v in rec.items()}
    for k in ("actions", "noise", "reset_noise", "meta", "renewable_sources", "spec"):
        res[k] = g[k]
    if tolerance is not None:
        res["meta"] = g["meta"].copy()
        res["meta"][5] = tolerance
    return res
