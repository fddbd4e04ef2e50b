"""GPU parity tests of the streamed OT path against the oracle and the committed golden vectors.

Tolerances are north_star's: plan marginals 1e-5 relative, plan entries rtol 1e-4 (entries above
1e-8 of the plan maximum — the reference's own check uses atol=1e-8), transition table 1e-4 with
identical argmax.  Everything goes through the C ABI (spadot_b200._lib)."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import ot_dense, ot_logdomain

pytestmark = pytest.mark.gpu

CFG = dict(ot_dense.DEFAULT_OT_CONFIG)
MARG_RTOL = 1e-5
ENTRY_RTOL = 1e-4


def rel_max(got, want):
    return float(np.abs(got - want).max() / np.abs(want).max())


def entry_rel(got, want):
    big = want > 1e-8 * want.max()
    return float((np.abs(got - want)[big] / want[big]).max())


@pytest.fixture(scope="module")
def ot():
    from spadot_b200 import ot_solvers, sinkhorn
    from spadot_b200.cuda_ops import CudaOps
    torch.cuda.set_device(0)
    return ot_solvers, sinkhorn, CudaOps


# ------------------------------------------------------------------ K3: one pass, teacher forced
@pytest.mark.parametrize("n,m,d", [(747, 1966, 20), (1966, 1916, 20), (130, 97, 32), (1, 1, 3), (65, 63, 5), (300, 5000, 32)])
@pytest.mark.parametrize("eps", [1.0, 0.05])
def test_lse_pass_matches_oracle(ot, n, m, d, eps):
    _, sinkhorn, CudaOps = ot
    a, b, _, _ = ot_dense.synthetic_embeddings(n, m, d, seed=n + m)
    C = ot_dense.sqeuclidean(a, b)
    med = float(np.median(C))
    rng = np.random.default_rng(3)
    g = rng.normal(0, 0.3, m)
    f = rng.normal(0, 0.3, n)
    cost = ot_logdomain.CostOperator(a, b, median=med)
    ops = CudaOps(a, b)
    ops.set_median(med)
    Lr = ops.row_lse(ops.tensor(g), eps).cpu().numpy()
    Lc = ops.col_lse(ops.tensor(f), eps).cpu().numpy()
    # fp32 tile math: absolute error of a natural-log LSE (a relative error of the row sum)
    assert np.abs(Lr - cost.row_lse(g / eps, eps)).max() < 2e-5
    assert np.abs(Lc - cost.col_lse(f / eps, eps)).max() < 2e-5


def test_lse_pass_masked_and_empty_columns(ot):
    """-inf potentials (zero-mass spots) must drop out; a row whose every column is masked gives -inf."""
    _, _, CudaOps = ot
    a, b, _, _ = ot_dense.synthetic_embeddings(70, 90, 4, seed=5)
    ops = CudaOps(a, b)
    ops.set_median(3.0)
    g = np.zeros(90)
    g[10:] = -np.inf
    cost = ot_logdomain.CostOperator(a, b, median=3.0)
    Lr = ops.row_lse(ops.tensor(g), 0.5).cpu().numpy()
    assert np.abs(Lr - cost.row_lse(g / 0.5, 0.5)).max() < 2e-5
    g[:] = -np.inf
    Lr = ops.row_lse(ops.tensor(g), 0.5).cpu().numpy()
    assert np.all(np.isneginf(Lr))


# ------------------------------------------------------------------ K5: exact median
@pytest.mark.parametrize("n,m,d", [(747, 1966, 20), (33, 34, 7), (3000, 2500, 32), (1, 5, 2)])
def test_median_is_exact(ot, n, m, d):
    _, sinkhorn, CudaOps = ot
    a, b, _, _ = ot_dense.synthetic_embeddings(n, m, d, seed=n)
    want = float(np.median(ot_dense.sqeuclidean(a, b)))
    ops = CudaOps(a, b)
    info = {}
    got = sinkhorn.median_cost(ops, small_limit=1 << 20, info=info)
    assert got == pytest.approx(want, rel=1e-15)
    if n * m > (1 << 20):
        assert info["sweeps"] == 2 and info["candidates"] < n * m // 100


def test_median_scipy_bit_exact(ot):
    """Same summation order as scipy's cdist and no FMA contraction: the median is bit-identical."""
    import scipy.spatial.distance as ssd
    _, sinkhorn, CudaOps = ot
    a, b, _, _ = ot_dense.synthetic_embeddings(211, 190, 20, seed=9)
    want = float(np.median(ssd.cdist(a, b, metric="sqeuclidean")))
    got = sinkhorn.median_cost(CudaOps(a, b))
    assert got == want


@pytest.mark.parametrize("n", [1, 2, 7, 4096, 100001])
def test_device_radix_select_matches_sort(ot, n):
    """sdb_select_ranks_f64 (eight digit passes on the device, no host round trip) = sorted()[k], ties and zeros included,
    and equal to the host-driven digit loop it replaces (sinkhorn._select_rank)."""
    _, sinkhorn, CudaOps = ot
    a, b, _, _ = ot_dense.synthetic_embeddings(8, 9, 3, seed=1)
    ops = CudaOps(a, b)
    rng = np.random.default_rng(n)
    v = rng.gamma(2.0, 30.0, size=n)
    v[rng.integers(0, n, size=max(1, n // 5))] = v[0]            # ties
    if n > 2:
        v[1] = 0.0
    srt = np.sort(v)
    cand = ops.tensor(v)
    for ranks in ([0], [n - 1], [(n - 1) // 2, n // 2], [n // 3]):
        got = ops.select_ranks(cand, n, ranks)
        assert got == [float(srt[k]) for k in ranks]
        assert got[0] == sinkhorn._select_rank(ops, cand, n, ranks[0])


# ------------------------------------------------------------------ full solves
@pytest.mark.parametrize("n,m,d,seed", [(747, 1966, 20, 1), (1966, 1916, 20, 2), (400, 300, 32, 3)])
def test_duality_gap_solve_matches_oracle(ot, n, m, d, seed):
    ot_solvers, sinkhorn, CudaOps = ot
    a, b, la, lb = ot_dense.synthetic_embeddings(n, m, d, seed=seed)
    Cn, med = ot_dense.median_normalised_cost(a, b)
    info_ref = {}
    want = ot_dense.duality_gap_solve(Cn, np.ones(n), info=info_ref, **CFG)
    cp = ot_solvers.solve_coupling(a, b, CFG)
    got = cp.plan().cpu().numpy()
    assert cp.median == pytest.approx(med, rel=1e-14)
    assert cp.info["iters_per_stage"] == info_ref["iters_per_stage"]
    assert rel_max(got.sum(1), want.sum(1)) < MARG_RTOL
    assert rel_max(got.sum(0), want.sum(0)) < MARG_RTOL
    assert entry_rel(got, want) < ENTRY_RTOL
    # marginals straight from the potentials (the path used when the plan is never materialised)
    assert rel_max(cp.row_mass().cpu().numpy(), want.sum(1)) < MARG_RTOL
    assert rel_max(cp.col_mass().cpu().numpy(), want.sum(0)) < MARG_RTOL
    # K6 transition table
    tab = cp.transition_table(la, lb, 10, 10).cpu().numpy()
    tab_ref = ot_dense.transition_table(want, la, lb, 10, 10)
    assert rel_max(tab, tab_ref) < 1e-4
    assert np.array_equal(tab.argmax(1), tab_ref.argmax(1))
    assert np.array_equal(ot_dense.plot_ot_normalisation(tab).argmax(1), ot_dense.plot_ot_normalisation(tab_ref).argmax(1))


def test_stablev2_matches_oracle(ot):
    ot_solvers, _, _ = ot
    n, m, d = 300, 411, 20
    a, b, _, _ = ot_dense.synthetic_embeddings(n, m, d, seed=1993)
    Cn, _ = ot_dense.median_normalised_cost(a, b)
    cfg = dict(CFG, scaling_iter=600, extra_iter=200)
    want = ot_dense.transport_stablev2(C=Cn, G=np.ones(n), **cfg)
    cp = ot_solvers.solve_coupling(a, b, cfg, solver="stablev2")
    got = cp.plan().cpu().numpy()
    assert rel_max(got.sum(1), want.sum(1)) < MARG_RTOL
    assert rel_max(got.sum(0), want.sum(0)) < MARG_RTOL
    assert entry_rel(got, want) < ENTRY_RTOL


# ------------------------------------------------------------------ golden vectors (generated from the reference itself)
@pytest.mark.parametrize("name", ["ot_small_48x61_d6", "ot_growth_90x70_d20", "ot_medium_300x411_d20", "ot_wotcfg_130x97_d32"])
def test_golden_vectors(ot, golden_dir, name):
    ot_solvers, _, _ = ot
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    cfg = {k: float(v) for k, v in zip(z["cfg_keys"], z["cfg_vals"])}
    cfg["growth_iters"] = int(cfg["growth_iters"])
    a, b, G = z["a"], z["b"], z["G"]
    cp = ot_solvers.solve_coupling(a, b, cfg, G=G)
    assert cp.median == pytest.approx(float(z["median"]), rel=1e-14)
    got = cp.plan().cpu().numpy()
    assert rel_max(got.sum(1), z["gap_row_sums"]) < MARG_RTOL
    assert rel_max(got.sum(0), z["gap_col_sums"]) < MARG_RTOL
    if "plan_gap" in z:
        assert entry_rel(got, z["plan_gap"]) < ENTRY_RTOL
    else:
        s = z["plan_gap_samples"]
        g = got[z["sample_i"], z["sample_j"]]
        big = s > 1e-8 * s.max()
        assert (np.abs(g - s)[big] / s[big]).max() < ENTRY_RTOL
    tab = cp.transition_table(z["labels_a"], z["labels_b"], 10, 10).cpu().numpy()
    assert rel_max(tab, z["table_gap"]) < 1e-4
    assert np.array_equal(tab.argmax(1), z["table_gap"].argmax(1))
    # fixed-schedule solver
    cp2 = ot_solvers.solve_coupling(a, b, cfg, G=G, solver="stablev2", median=cp.median)
    got2 = cp2.plan().cpu().numpy()
    assert rel_max(got2.sum(1), z["v2_row_sums"]) < MARG_RTOL
    assert rel_max(got2.sum(0), z["v2_col_sums"]) < MARG_RTOL


def test_compute_transport_map_is_drop_in(ot, golden_dir):
    """Same call as utils/_train_utils.py:318; growth loop mutates config['G'] like ot_solvers.py:113-117."""
    ot_solvers, _, _ = ot
    z = np.load(os.path.join(golden_dir, "ot_growth_90x70_d20.npz"))
    cfg = {k: float(v) for k, v in zip(z["cfg_keys"], z["cfg_vals"])}
    cfg["growth_iters"] = int(cfg["growth_iters"])
    cfg["some_unknown_key"] = "ignored"          # **ignored, ot_solvers.py:178
    gamma = ot_solvers.compute_transport_map(torch.from_numpy(z["a"]), torch.from_numpy(z["b"]), cfg, G=z["G"].copy())
    assert isinstance(gamma, np.ndarray) and gamma.dtype == np.float64 and gamma.shape == (90, 70)
    assert entry_rel(gamma, z["plan_gap"]) < ENTRY_RTOL
    # after the loop config["G"] holds the growth vector fed to the last solve = row sums of iteration growth_iters-2
    assert rel_max(np.asarray(cfg["G"]), z["growth_row_sums"][-2]) < MARG_RTOL


def test_growth_dynamic_range(ot):
    """Growth rates spanning hundreds of orders of magnitude (OT_g.txt reaches 4e-93) stay finite."""
    ot_solvers, _, _ = ot
    n, m, d = 200, 180, 10
    a, b, _, _ = ot_dense.synthetic_embeddings(n, m, d, seed=4)
    G = np.exp(np.random.default_rng(1).uniform(-200, 2, n))
    Cn, _ = ot_dense.median_normalised_cost(a, b)
    want = ot_dense.duality_gap_solve(Cn, G, **CFG)
    cp = ot_solvers.solve_coupling(a, b, CFG, G=G)
    got_rows = cp.row_mass().cpu().numpy()
    w = want.sum(1)
    assert np.all(np.isfinite(got_rows))
    assert np.abs(np.log(got_rows) - np.log(w)).max() < 1e-5     # relative, row by row, down to 1e-90


def test_large_pass_linearity_properties(ot):
    """Size-independent properties at a size the CPU oracle cannot hold: shifting g by a constant c
    shifts every row LSE by c/eps; restricting columns to a partition and combining partial LSEs
    reproduces the full LSE (checksum of checksums)."""
    _, _, CudaOps = ot
    n, m, d = 20000, 30000, 32
    a, b, _, _ = ot_dense.synthetic_embeddings(n, m, d, seed=77)
    ops = CudaOps(a, b)
    ops.set_median(160.0)
    eps = 0.05
    g = ops.tensor(np.random.default_rng(0).normal(0, 0.2, m))
    L0 = ops.row_lse(g, eps).clone()
    L1 = ops.row_lse(g + 0.125, eps)
    assert float((L1 - L0 - 0.125 / eps).abs().max()) < 2e-5
    # spot-check 64 random rows against the fp64 oracle
    idx = np.random.default_rng(1).integers(0, n, 64)
    cost = ot_logdomain.CostOperator(a[idx], b, median=160.0)
    want = cost.row_lse(g.cpu().numpy() / eps, eps)
    assert np.abs(L0.cpu().numpy()[idx] - want).max() < 2e-5


def test_reductions_of_different_width_share_scratch(ot):
    """Regression: with > 256 reduction blocks the 4-wide / 10-wide partials used to overwrite the counter of
    the narrower reductions.  Interleave all three at a size that needs ~1000 blocks."""
    _, _, CudaOps = ot
    n, m = 250_000, 240_000
    a = np.zeros((n, 2))
    b = np.zeros((m, 2))
    ops = CudaOps(a, b, tc="off")
    rng = np.random.default_rng(0)
    f, g = rng.normal(0, 0.1, n), rng.normal(0, 0.1, m)
    L = rng.normal(0, 0.1, n)
    ft, gt, Lt = ops.tensor(f), ops.tensor(g), ops.tensor(L)
    z_n, z_m = ops.zeros(n), ops.zeros(m)
    for _ in range(3):
        t10 = ops.gap_terms(ft, Lt, z_n, gt, z_m, z_m, 0.5, 0.1, 5.0, 1.0 / n, 1.0 / m).cpu().numpy()
        c4 = ops.stage_criterion(ft, z_n, z_n, gt, z_m, z_m, 0.5).cpu().numpy()
        s1 = float(ops.sum_exp(Lt).item())
        assert s1 == pytest.approx(np.exp(L).sum(), rel=1e-12)
        assert c4[1] == pytest.approx(np.sum(np.exp(f / 0.5) ** 2), rel=1e-12)
        assert c4[3] == pytest.approx(np.sum(np.exp(g / 0.5) ** 2), rel=1e-12)
        assert t10[0] == pytest.approx(np.sum(np.exp(f / 0.5 + L)), rel=1e-12)
        assert t10[4] == pytest.approx(np.sum(np.exp(g / 0.5)), rel=1e-12)


def test_analyze_pipeline_wot_conventions_chickenheart_shapes(ot):
    """`SpaDOT analyze` on arrays: 3 adjacent pairs at ChickenHeart sizes, wot conventions (tau=10000, last growth
    iteration).  Domain transitions must reproduce the dense oracle: tables within 1e-4, identical argmax, identical
    dot-plot (min of row/col-normalised) argmax."""
    from spadot_b200 import analyze
    sizes, d = [747, 1966, 1916, 1967], 20
    rng = np.random.default_rng(1993)
    centres = rng.normal(0, 1.5, size=(10, d))
    embs, labs = [], []
    for t, n in enumerate(sizes):
        lab = rng.integers(0, 10, n)
        embs.append(centres[lab] + rng.normal(0, 0.5, size=(n, d)) + 0.1 * t)
        labs.append(lab)
    tables = analyze.ot_analysis(embs, labs, n_domains=[10] * 4)
    cfg = dict(ot_dense.DEFAULT_OT_CONFIG, **analyze.WOT_CONFIG)
    for t in range(3):
        gammas = ot_dense.compute_transport_map(embs[t], embs[t + 1], cfg, return_all=True)
        want = ot_dense.transition_table(gammas[-1], labs[t], labs[t + 1], 10, 10)
        assert rel_max(tables[t], want) < 1e-4
        assert np.array_equal(tables[t].argmax(1), want.argmax(1))
        assert np.array_equal(analyze.transition_probabilities(tables[t]).argmax(1),
                              ot_dense.plot_ot_normalisation(want).argmax(1))


def test_train_side_centroid_coupling_10x10(ot):
    """utils/_train_utils.py:309-321: 10 k-means centroids per timepoint, z=20, torch tensors in, gammas[0] out."""
    ot_solvers, _, _ = ot
    rng = np.random.default_rng(3)
    a, b = rng.normal(0, 1.5, (10, 20)), rng.normal(0, 1.5, (10, 20)) + 0.2
    cfg = dict(CFG)
    want = ot_dense.compute_transport_map(a, b, cfg)
    got_cpu_tensor = ot_solvers.compute_transport_map(torch.from_numpy(a), torch.from_numpy(b), dict(cfg))
    got_cuda_tensor = ot_solvers.compute_transport_map(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), dict(cfg))
    for got in (got_cpu_tensor, got_cuda_tensor):
        assert got.shape == (10, 10)
        assert rel_max(got.sum(1), want.sum(1)) < MARG_RTOL
        assert entry_rel(got, want) < ENTRY_RTOL
    # the consumer normalises rows (utils/_train_utils.py:296-300)
    row_norm = got_cpu_tensor / got_cpu_tensor.sum(axis=1, keepdims=True)
    want_norm = want / want.sum(axis=1, keepdims=True)
    assert np.abs(row_norm - want_norm).max() < 1e-5


def test_analyze_writes_growth_and_tables(ot, tmp_path):
    from spadot_b200 import analyze
    rng = np.random.default_rng(0)
    embs = [rng.normal(0, 1, (150, 8)), rng.normal(0, 1, (170, 8)) + 0.1, rng.normal(0, 1, (160, 8)) + 0.2]
    labs = [np.array([f"D{t}_{k}" for k in rng.integers(0, 4, e.shape[0])]) for t, e in enumerate(embs)]
    tables = analyze.ot_analysis(embs, labs, out_dir=str(tmp_path), prefix="run_")
    assert len(tables) == 2 and tables[0].shape == (4, 4)
    g = np.loadtxt(tmp_path / "OT_g.txt", skiprows=1, usecols=(1, 2, 3, 4))
    assert g.shape == (150 + 170, 4) and np.all(g[:, 0] == 1.0) and np.all(g > 0)
    ext = ".h5ad" if analyze.have_anndata() else ".npz"           # _analyze_utils.py:138 format when anndata is installed
    tab, rows, cols = analyze.read_transition_table(str(tmp_path / ("run_transition_table_0_1" + ext)))
    assert rows == ["D0_0", "D0_1", "D0_2", "D0_3"] and cols == ["D1_0", "D1_1", "D1_2", "D1_3"] and np.allclose(tab, tables[0])
    # growth column k+1 = row sums of the plan of growth iteration k (ot_solvers.py:116); the table sums to the last one
    assert tables[0].sum() == pytest.approx(g[:150, 3].sum(), rel=1e-6)


class _LoopbackDist:
    """The multi-rank code path on one GPU: a 2-rank group whose other rank owns no rows (its partial
    column LSE is -inf, its sums are 0), so every collective is the identity."""

    def __new__(cls, sinkhorn):
        class Loop(sinkhorn.Dist):
            def __init__(self):
                super().__init__(enabled=False)
                self.world = 2

            def sum_(self, t):
                self.collectives += 1
                return t

            def max_(self, t):
                self.collectives += 1
                return t

            def min_(self, t):
                self.collectives += 1
                return t

            def gather_cat(self, t):
                return t

            def native_comm(self, device):
                return getattr(self, "comm", None)
        return Loop()


def _world1_nccl_comm():
    """An NCCL communicator of ONE rank owned by libspadot_b200.so: its all-reduce is the identity, like the loop-back
    group's collectives, so the native multi-rank loop (sdb_sinkhorn_sweeps_dist) can run on a single GPU."""
    import ctypes
    from spadot_b200 import _lib
    lib = _lib.load()
    if not lib.sdb_nccl_available():
        pytest.skip("libnccl.so.2 cannot be bound in this process")
    buf = (ctypes.c_char * 128)()
    _lib.check(lib.sdb_nccl_unique_id(buf))
    comm = ctypes.c_void_p()
    _lib.check(lib.sdb_nccl_comm_create(buf, 1, 0, ctypes.byref(comm)))
    return comm


@pytest.mark.parametrize("n,m,d,tc,native", [(1500, 1300, 20, "off", False), (2600, 2100, 32, "on", False),
                                              (2600, 2100, 32, "on", True), (5000, 3300, 20, "on", True)])
def test_multi_rank_code_path_matches_native_loop(ot, n, m, d, tc, native):
    """Regression: the per-iteration Python loop used for world > 1 (row pass fused, column pass + all-reduce +
    potential update) must walk the same iterates as the native single-rank loop — same iterations per epsilon
    stage, same potentials.  (A stale cached bias vector once froze g inside a stage on this path only.)"""
    ot_solvers, sinkhorn, CudaOps = ot
    a, b, _, _ = ot_dense.synthetic_embeddings(n, m, d, seed=n)
    G = np.exp(np.random.default_rng(1).normal(0, 0.3, n))
    cfg = dict(CFG)
    cp1 = ot_solvers.solve_coupling(a, b, cfg, G=G, ops=CudaOps(a, b, tc=tc), dist=sinkhorn.Dist(enabled=False))
    loop = _LoopbackDist(sinkhorn)
    if native:
        loop.comm = _world1_nccl_comm()                 # sdb_sinkhorn_sweeps_dist: kernels + ncclAllReduce issued from C
    ops2 = CudaOps(a, b, tc=tc)
    ops2.PREDICT_MIN_PAIRS = 0                          # the predicted stabiliser rides along on the tensor-core case
    l0 = ops2.launches
    cp2 = ot_solvers.solve_coupling(a, b, cfg, G=G, ops=ops2, dist=loop)
    assert (ops2._pred is not None and ops2._pred.ok) == (tc == "on")
    assert loop.collectives > 0
    if tc == "on":
        # one collective per iteration except the first of a stage (3) + one per stopping-rule check
        total = cp2.info["total_iters"]
        assert loop.collectives <= 1.6 * total, (loop.collectives, total)
    assert cp1.median == cp2.median
    assert cp1.info["iters_per_stage"] == cp2.info["iters_per_stage"], (cp1.info, cp2.info)
    assert float((cp1.f - cp2.f).abs().max()) < 1e-6
    assert float((cp1.g - cp2.g).abs().max()) < 1e-6


def test_two_gpu_row_partition_matches_single_gpu():
    """tests/dist_gpu_check.py under torchrun (NCCL) when the box has >= 2 GPUs."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29533", os.path.join(root, "tests", "dist_gpu_check.py")],
                       capture_output=True, text=True, timeout=280)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert r.stdout.count("dist check") == 2, r.stdout


@pytest.mark.parametrize("n,m,d", [(1, 1, 3), (1, 700, 2), (65, 63, 5), (130, 4000, 7), (1966, 1916, 20), (300, 411, 70)])
def test_persistent_sweeps_match_launch_loop(ot, n, m, d):
    """sdb_sinkhorn_sweeps_persistent (one cooperative launch per batch of iterations) walks the same iterates as the
    five-launches-per-iteration loop: same iterations per stage, potentials equal to fp64 rounding of the LSE combine."""
    ot_solvers, sinkhorn, CudaOps = ot
    a, b, _, _ = ot_dense.synthetic_embeddings(n, m, d, seed=n + m)
    G = np.exp(np.random.default_rng(2).normal(0, 0.3, n))
    out = []
    for persistent in (True, False):
        ops = CudaOps(a, b, tc="off")
        ops.persistent = persistent
        ops.fused_solve = None       # the batch-of-iterations kernel under the host's stage loop (the one-launch solve has its own test)
        ops.SIMT_DOT_MAX = 0.0       # same tile arithmetic on both sides
        l0 = ops.launches
        cp = ot_solvers.solve_coupling(a, b, dict(CFG), G=G, ops=ops, dist=sinkhorn.Dist(enabled=False))
        out.append((cp, ops.launches - l0))
    (cp1, l1), (cp2, l2) = out
    assert l1 < l2 / 2, (l1, l2)                       # the persistent path really ran (far fewer launches)
    assert cp1.info["iters_per_stage"] == cp2.info["iters_per_stage"]
    assert float((cp1.f - cp2.f).abs().max()) < 1e-7
    assert float((cp1.g - cp2.g).abs().max()) < 1e-7
    assert abs(cp1.info["gap"]) <= 1e-8


@pytest.mark.parametrize("form", ["strips", "resident", "streamed"])
@pytest.mark.parametrize("n,m,d,tau", [(747, 1966, 20, 1000.0), (300, 411, 20, 1000.0), (130, 97, 6, 3.0), (1, 1, 3, 1000.0),
                                       (65, 64, 33, 1.5), (1966, 1916, 20, 1000.0), (2500, 3000, 32, 1000.0)])
def test_whole_solve_in_one_launch_matches_host_stage_loop(ot, n, m, d, tau, form):
    """sdb_sinkhorn_solve_persistent (six epsilon stages, stopping rules and tau bookkeeping on the device, two grid barriers
    per iteration) against the host-driven stage loop over the same tile code: same iterations per stage, same potentials,
    same frames; small tau forces absorptions (ot_func.cpp:778-819) so the row-by-row deferred absorb is exercised."""
    ot_solvers, sinkhorn, CudaOps = ot
    a, b, _, _ = ot_dense.synthetic_embeddings(n, m, d, seed=n + m)
    G = np.exp(np.random.default_rng(n).normal(0, 0.4, n))
    cfg = dict(CFG, tau=tau)
    out = []
    for fused in (True, False):
        ops = CudaOps(a, b, tc="off")
        resident = form != "streamed"
        ops.STRIP_FORM = form == "strips"                        # whole rows / columns of the cost matrix per CTA (owner computes),
        ops.RESIDENT_TILES = form == "resident"                  # 64x64 cost tiles resident in shared memory, or streamed tiles
        ops.RESIDENT_MAX_TILES_PER_CTA = 6                       # (the default gate keeps larger problems on streamed tiles)
        ops.SIMT_DOT_MAX = 0.0 if resident else ops.SIMT_DOT_MAX  # the resident form is a direct-difference form: compare like with like
        if not fused:
            ops.fused_solve = None
        l0 = ops.launches
        cp = ot_solvers.solve_coupling(a, b, cfg, G=G, ops=ops, dist=sinkhorn.Dist(enabled=False), median=1.7 * d)
        out.append((cp, ops.launches - l0))
    (c1, l1), (c2, l2) = out
    assert l1 <= 2 and l2 > 20                                  # one launch vs the batch-by-batch loop
    fits = c1.ops.strip_bytes(torch.cuda.get_device_properties(0).multi_processor_count) <= c1.ops.STRIP_SMEM_MAX
    if form == "strips" and fits:
        assert c1.ops.solve_form == 2                           # the strip form really ran
    elif form == "resident" and n * m <= 64 * 64 * 2 * 148 * 6:
        assert c1.ops.solve_form == 1
    if n * m > 1024:
        assert c1.info["iters_per_stage"] == c2.info["iters_per_stage"], (c1.info, c2.info)
    else:
        # a 1 x 1 problem has ONE term per sum: nothing averages the fp32 rounding of t (|t| ~ 120, ulp 8e-6), the two forms round
        # differently, and the stage criterion (5.6e-7 vs 1.2e-6 against the reference's 1e-6, profiles/r2_dbg_1x1.txt) can take
        # one more batch; both potentials are within that rounding of the fp64 oracle
        for it1, it2 in zip(c1.info["iters_per_stage"], c2.info["iters_per_stage"]):
            assert abs(it1 - it2) <= 5
    assert c1.info["gap"] == pytest.approx(c2.info["gap"], rel=0.05, abs=1e-12)      # a 1e-10 difference of O(1) sums of fp32-tile results
    # fp32-noise level: the host loop's gap-check row pass uses the launch path's column splits, the one-launch solve the
    # cooperative grid's, so their fp32 partial sums differ in the last bit (6e-8 on an LSE, 2e-9 on a potential)
    # (a single-term LSE is the fp32 rounding of t itself: half an ulp of |t| ~ 120 is 4e-6)
    lse_tol = 1e-6 if n * m > 1024 else 1e-5
    for name, tol in (("f", 4e-7), ("g", 4e-7), ("u", 4e-7), ("v", 4e-7), ("Lr", lse_tol), ("Lc", lse_tol)):
        x, y = getattr(c1.state, name), getattr(c2.state, name)
        assert float((x - y).abs().max()) < tol * max(1.0, float(y.abs().max())), name      # relative to the vector's magnitude


def test_whole_solve_in_one_launch_nan_and_max_iter(ot):
    ot_solvers, sinkhorn, CudaOps = ot
    a, b, _, _ = ot_dense.synthetic_embeddings(40, 50, 4, seed=3)
    G = np.ones(40)
    G[7] = 0.0                                                  # r log(r/p) with p = 0 -> NaN gap (ot_func.cpp:309-322)
    with pytest.raises(RuntimeError, match="Overflow encountered in duality gap computation"):
        ot_solvers.solve_coupling(a, b, dict(CFG), G=G, ops=CudaOps(a, b, tc="off"), dist=sinkhorn.Dist(enabled=False))
    with pytest.warns(RuntimeWarning, match="Reached max_iter"):
        cp = ot_solvers.solve_coupling(a, b, dict(CFG, max_iter=5), ops=CudaOps(a, b, tc="off"), dist=sinkhorn.Dist(enabled=False))
    assert cp.info["iters_per_stage"] == [5] * 6 and cp.info["max_iter_reached"]
