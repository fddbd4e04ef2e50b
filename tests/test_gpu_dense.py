"""GPU parity of the dense-cost entry points (reference signatures that take C) — fp64 end to end, so the
bar is 1e-9 relative against the golden vectors generated from the reference."""
import os

import numpy as np
import pytest
import torch

from oracle import ot_dense

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


def load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    cfg = {k: float(v) for k, v in zip(z["cfg_keys"], z["cfg_vals"])}
    cfg["growth_iters"] = int(cfg["growth_iters"])
    return z, cfg


@pytest.mark.parametrize("name", ["ot_small_48x61_d6", "ot_growth_90x70_d20"])
def test_dense_signatures_match_reference_golden(golden_dir, name):
    from spadot_b200 import ot_solvers
    z, cfg = load(golden_dir, name)
    Cn, _ = ot_dense.median_normalised_cost(z["a"], z["b"])
    plan = ot_solvers.optimal_transport_duality_gap(C=Cn, G=z["G"], **cfg)
    assert plan.dtype == np.float64 and plan.shape == Cn.shape
    np.testing.assert_allclose(plan, z["plan_gap"], rtol=1e-9, atol=1e-300)
    plan2 = ot_solvers.transport_stablev2(C=Cn, G=z["G"], **cfg)
    np.testing.assert_allclose(plan2, z["plan_v2"], rtol=1e-9, atol=1e-300)
    gamma = ot_solvers.compute_transport_map(z["a"], z["b"], dict(cfg), C=Cn, G=z["G"].copy())
    np.testing.assert_allclose(gamma, z["plan_gap"], rtol=1e-9, atol=1e-300)


def test_dense_medium_and_torch_inputs(golden_dir):
    from spadot_b200 import ot_solvers
    z, cfg = load(golden_dir, "ot_medium_300x411_d20")
    Cn, _ = ot_dense.median_normalised_cost(z["a"], z["b"])
    plan = ot_solvers.optimal_transport_duality_gap(C=torch.from_numpy(Cn), G=torch.from_numpy(z["G"]), **cfg)
    np.testing.assert_allclose(plan[z["sample_i"], z["sample_j"]], z["plan_gap_samples"], rtol=1e-9, atol=1e-300)
    np.testing.assert_allclose(plan.sum(1), z["gap_row_sums"], rtol=1e-10)


def test_dense_lse_ragged_and_large():
    from spadot_b200.cuda_ops import DenseOps
    rng = np.random.default_rng(0)
    for n, m in [(1, 1), (7, 300), (1000, 33), (2049, 4097)]:
        C = rng.uniform(0, 5, (n, m))
        f, g = rng.normal(0, 0.3, n), rng.normal(0, 0.3, m)
        ops = DenseOps(C)
        Lr = ops.row_lse(ops.tensor(g), 0.05).cpu().numpy()
        Lc = ops.col_lse(ops.tensor(f), 0.05).cpu().numpy()
        t = (g[None, :] - C) / 0.05
        want_r = t.max(1) + np.log(np.exp(t - t.max(1, keepdims=True)).sum(1))
        t = (f[:, None] - C) / 0.05
        want_c = t.max(0) + np.log(np.exp(t - t.max(0, keepdims=True)).sum(0))
        np.testing.assert_allclose(Lr, want_r, rtol=0, atol=1e-11)
        np.testing.assert_allclose(Lc, want_c, rtol=0, atol=1e-11)
