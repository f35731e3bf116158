"""CPU tier: ``compat.RolloutBuffer`` against the reference's ``GridDataset``
(``/root/reference/grid_fed_rl/algorithms/base.py:180-265``).

``tests/golden/dataset_gridDataset.npz`` was written by ``oracle/ref_harness.py dataset``: the UNMODIFIED
reference class fed a seeded transition set (257 x 19 observations incl. one constant column, 3 actions),
with and without normalisation - its statistics, the arrays it holds, what ``get_all_data`` hands a learner
(float32) and its two denormalisers on probe vectors."""
import os

import numpy as np
import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dataset_gridDataset.npz")
FIELDS = ("observations", "actions", "rewards", "next_observations", "terminals")


@pytest.fixture(scope="module")
def g():
    z = np.load(GOLDEN, allow_pickle=False)
    return {k: z[k] for k in z.files}


def _filled(g, dtype, chunks=(100, 57, 100)):
    from grid_fed_rl_b200.compat import RolloutBuffer
    raw = {k: torch.as_tensor(g["raw_" + k].astype(np.float64)) for k in FIELDS}
    N, D, A = raw["observations"].shape[0], raw["observations"].shape[1], raw["actions"].shape[1]
    buf = RolloutBuffer(N, D, A, "cpu", dtype=dtype)
    o = 0
    for c in chunks:                                     # appended batch by batch, as the collectors do
        s = slice(o, o + c)
        assert buf.add(*(raw[k][s] for k in FIELDS)) == c
        o += c
    assert buf.size == N == int(g["plain_size"]) and buf.add(*(raw[k][:5] for k in FIELDS)) == 0   # full
    return buf


def test_plain_buffer_equals_griddataset(g):
    buf = _filled(g, torch.float64)
    data = buf.get_all_data()
    for k in FIELDS:
        assert np.array_equal(data[k].numpy(), g["plain_" + k].astype(np.float64)), k
        assert np.array_equal(data[k].float().numpy(), g["plain_all_" + k]), k
    pa, po = torch.from_numpy(g["probe_action"]), torch.from_numpy(g["probe_observation"])
    assert np.array_equal(buf.denormalize_action(pa).numpy(), g["plain_denorm_action"])
    assert np.array_equal(buf.denormalize_observation(po).numpy(), g["plain_denorm_observation"])
    out = buf.to_numpy()
    assert out["terminals"].dtype == bool and np.array_equal(out["terminals"], g["raw_terminals"])


def test_normalised_buffer_equals_griddataset(g):
    buf = _filled(g, torch.float64)
    buf.normalize()
    for name, mine in (("obs_mean", buf.obs_mean), ("obs_std", buf.obs_std), ("action_mean", buf.action_mean),
                       ("action_std", buf.action_std), ("reward_mean", buf.reward_mean), ("reward_std", buf.reward_std)):
        assert np.allclose(np.asarray(mine), g["stat_" + name], rtol=1e-13, atol=1e-15), name
    # the constant column: std = 0 + 1e-6 exactly, entries (x - mean) / 1e-6
    assert abs(float(buf.obs_std[4]) - 1e-6) < 1e-18
    data = buf.get_all_data()
    for k in FIELDS:
        ref = g["norm_" + k].astype(np.float64)
        scale = 1.0 if k not in ("observations", "next_observations") else None
        err = np.abs(data[k].numpy() - ref)
        if scale is None:                                # the constant column amplifies rounding by 1e6
            err[:, 4] *= 1e-6
        assert err.max() < 1e-11, k
        assert np.allclose(data[k].float().numpy(), g["norm_all_" + k], rtol=1e-6, atol=1e-5), k
    pa, po = torch.from_numpy(g["probe_action"]), torch.from_numpy(g["probe_observation"])
    assert np.allclose(buf.denormalize_action(pa.double()).numpy(), g["norm_denorm_action"], rtol=1e-6, atol=1e-6)
    assert np.allclose(buf.denormalize_observation(po.double()).numpy(), g["norm_denorm_observation"], rtol=1e-6, atol=1e-5)


def test_float32_buffer_is_what_learners_get(g):
    """The device buffer's default dtype is the learners' (torch.FloatTensor upstream, base.py:236-243)."""
    buf = _filled(g, torch.float32)
    for k in FIELDS:
        assert np.array_equal(buf.get_all_data()[k].numpy(), g["plain_all_" + k]), k
    buf.normalize()
    keep = [c for c in range(g["raw_observations"].shape[1]) if c != 4]
    assert np.allclose(buf.observations[:, keep].numpy(), g["norm_all_observations"][:, keep], rtol=2e-4, atol=2e-5)
    assert np.allclose(buf.rewards.numpy(), g["norm_all_rewards"], rtol=2e-4, atol=2e-5)
    batch = buf.sample_batch(64, generator=torch.Generator().manual_seed(0))
    assert set(batch) == set(FIELDS) and batch["observations"].shape == (64, 19) and batch["actions"].shape == (64, 3)
