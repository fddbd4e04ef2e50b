"""GPU parity of the tcgen05/TMEM/TMA form of the streamed LSE pass (sdb_lse_pass_tc) against the fp64
oracle and against the SIMT form, then full solves with the tensor-core pass forced on."""
import numpy as np
import pytest
import torch

from oracle import ot_dense, ot_logdomain

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]

CFG = dict(ot_dense.DEFAULT_OT_CONFIG)


@pytest.fixture(scope="module")
def ot():
    from spadot_b200 import ot_solvers, sinkhorn
    from spadot_b200.cuda_ops import CudaOps
    torch.cuda.set_device(0)
    return ot_solvers, sinkhorn, CudaOps


@pytest.mark.parametrize("n,m,d", [(128, 256, 32), (747, 1966, 20), (1966, 1916, 20), (300, 5000, 32), (130, 97, 6),
                                   (1000, 1300, 48), (1, 1, 3), (5000, 40000, 32), (40000, 3000, 16)])
@pytest.mark.parametrize("eps", [1.0, 0.05])
def test_tc_pass_matches_oracle(ot, n, m, d, eps):
    _, _, CudaOps = ot
    a, b, _, _ = ot_dense.synthetic_embeddings(n, m, d, seed=n + m)
    rng = np.random.default_rng(3)
    if n * m <= 4_000_000:
        C = ot_dense.sqeuclidean(a, b)
        med = float(np.median(C))
        rows = np.arange(n)
        cols = np.arange(m)
    else:
        med = 2.0 * d * 2.5
        rows = rng.integers(0, n, 96)
        cols = rng.integers(0, m, 96)
    g = rng.normal(0, 0.3, m)
    f = rng.normal(0, 0.3, n)
    ops = CudaOps(a, b, tc="on")
    ops.set_median(med)
    Lr = ops.row_lse(ops.tensor(g), eps).cpu().numpy()
    Lc = ops.col_lse(ops.tensor(f), eps).cpu().numpy()
    want_r = ot_logdomain.CostOperator(a[rows], b, median=med).row_lse(g / eps, eps)
    want_c = ot_logdomain.CostOperator(a, b[cols], median=med).col_lse(f / eps, eps)
    assert np.abs(Lr[rows] - want_r).max() < 2e-5
    assert np.abs(Lc[cols] - want_c).max() < 2e-5


def test_tc_and_simt_passes_agree(ot):
    _, _, CudaOps = ot
    a, b, _, _ = ot_dense.synthetic_embeddings(3000, 2900, 32, seed=12)
    g = np.random.default_rng(0).normal(0, 0.3, 2900)
    o1 = CudaOps(a, b, tc="on")
    o2 = CudaOps(a, b, tc="off")
    for o in (o1, o2):
        o.set_median(160.0)
    L1 = o1.row_lse(o1.tensor(g), 0.05)
    L2 = o2.row_lse(o2.tensor(g), 0.05)
    assert float((L1 - L2).abs().max()) < 2e-5


def test_tc_masked_columns_and_empty_rows(ot):
    _, _, CudaOps = ot
    a, b, _, _ = ot_dense.synthetic_embeddings(200, 700, 8, seed=5)
    ops = CudaOps(a, b, tc="on")
    ops.set_median(20.0)
    g = np.zeros(700)
    g[300:] = -np.inf
    cost = ot_logdomain.CostOperator(a, b, median=20.0)
    Lr = ops.row_lse(ops.tensor(g), 0.5).cpu().numpy()
    assert np.abs(Lr - cost.row_lse(g / 0.5, 0.5)).max() < 2e-5
    g[:] = -np.inf
    assert np.all(np.isneginf(ops.row_lse(ops.tensor(g), 0.5).cpu().numpy()))


@pytest.mark.parametrize("n,m,d,seed", [(1966, 1916, 20, 2), (2500, 2048, 32, 3)])
def test_tc_full_solve_matches_oracle(ot, n, m, d, seed):
    ot_solvers, sinkhorn, CudaOps = ot
    a, b, la, lb = ot_dense.synthetic_embeddings(n, m, d, seed=seed)
    Cn, med = ot_dense.median_normalised_cost(a, b)
    info_ref = {}
    want = ot_dense.duality_gap_solve(Cn, np.ones(n), info=info_ref, **CFG)
    ops = CudaOps(a, b, tc="on")
    cp = ot_solvers.solve_coupling(a, b, CFG, ops=ops)
    got = cp.plan().cpu().numpy()
    assert cp.info["iters_per_stage"] == info_ref["iters_per_stage"]
    assert np.abs(got.sum(1) - want.sum(1)).max() / want.sum(1).max() < 1e-5
    assert np.abs(got.sum(0) - want.sum(0)).max() / want.sum(0).max() < 1e-5
    big = want > 1e-8 * want.max()
    assert (np.abs(got - want)[big] / want[big]).max() < 1e-4
    tab = cp.transition_table(la, lb, 10, 10).cpu().numpy()
    tab_ref = ot_dense.transition_table(want, la, lb, 10, 10)
    assert np.abs(tab - tab_ref).max() / tab_ref.max() < 1e-4
    assert np.array_equal(tab.argmax(1), tab_ref.argmax(1))


def test_tc_large_properties(ot):
    """At a size no CPU oracle can hold densely: shift equivariance + spot rows against fp64."""
    _, _, CudaOps = ot
    n, m, d = 60000, 50000, 32
    a, b, _, _ = ot_dense.synthetic_embeddings(n, m, d, seed=77)
    ops = CudaOps(a, b)
    assert ops.use_tc
    ops.set_median(160.0)
    eps = 0.05
    g = ops.tensor(np.random.default_rng(0).normal(0, 0.2, m))
    L0 = ops.row_lse(g, eps).clone()
    L1 = ops.row_lse(g + 0.125, eps)
    assert float((L1 - L0 - 0.125 / eps).abs().max()) < 2e-5
    idx = np.random.default_rng(1).integers(0, n, 64)
    want = ot_logdomain.CostOperator(a[idx], b, median=160.0).row_lse(g.cpu().numpy() / eps, eps)
    assert np.abs(L0.cpu().numpy()[idx] - want).max() < 2e-5


@pytest.mark.parametrize("n,m,d", [(3000, 2500, 32), (2047, 4099, 20), (5000, 6000, 10), (9000, 7000, 48)])
def test_tc_median_sweeps_are_exact(ot, n, m, d):
    """K5 on the tensor cores: same exact median as numpy, 2 sweeps."""
    _, sinkhorn, CudaOps = ot
    a, b, _, _ = ot_dense.synthetic_embeddings(n, m, d, seed=n)
    want = float(np.median(ot_dense.sqeuclidean(a, b)))
    ops = CudaOps(a, b, tc="on")
    info = {}
    got = sinkhorn.median_cost(ops, small_limit=1 << 20, info=info)
    assert got == pytest.approx(want, rel=1e-15)
    assert info["sweeps"] == 2 and info["candidates"] < n * m // 100
    # counts of the tensor-core and SIMT sweeps agree away from the bracket edges
    lo, hi = want * 0.9, want * 1.1
    h1, c1 = ops.cost_histogram(lo, hi, 64)
    ops.use_tc = False
    h2, c2 = ops.cost_histogram(lo, hi, 64)
    ops.use_tc = True
    total = n * m
    assert abs(int(c1[0]) - int(c2[0])) <= max(4, total // 100000)
    assert int((h1 - h2).abs().max()) <= max(8, total // 50000)


def test_full_size_1Mx1M_properties(ot):
    """BASELINE.json configs[3] at full size (1M x 1M, d=32), where no CPU oracle can hold the problem:
    (i) shift equivariance of the row pass, (ii) checksum of checksums — the LSE over all columns equals the
    log-sum-exp of the LSEs over a 3-way column partition (run as three separate couplings), (iii) 48 random
    rows against the fp64 oracle, (iv) one full Sinkhorn iteration keeps the potentials finite and its column
    marginal consistent with the row marginal (total mass identity sum_i R_i = sum_j C_j)."""
    _, sinkhorn, CudaOps = ot
    import bench
    n = m = 1_000_000
    x, y = bench.synth(n, m, 32)
    ops = CudaOps(x, y)
    assert ops.use_tc
    med, eps = 160.0, 0.05
    ops.set_median(med)
    rng = np.random.default_rng(0)
    g_np = rng.normal(0, 0.2, m)
    g = ops.tensor(g_np)
    L0 = ops.row_lse(g, eps).clone()
    L1 = ops.row_lse(g + 0.25, eps)
    assert float((L1 - L0 - 0.25 / eps).abs().max()) < 2e-5
    # (ii) column partition: mask out all but one third of the columns with -inf potentials
    parts = []
    for k in range(3):
        gk = torch.full_like(g, float("-inf"))
        gk[k::3] = g[k::3]
        parts.append(ops.row_lse(gk, eps).clone())
    combined = torch.logsumexp(torch.stack(parts), dim=0)
    assert float((combined - L0).abs().max()) < 2e-5
    # (iii) spot rows
    idx = rng.integers(0, n, 48)
    want = ot_logdomain.CostOperator(x[idx], y, median=med, block=8).row_lse(g_np / eps, eps)
    assert np.abs(L0.cpu().numpy()[idx] - want).max() < 2e-5
    # (iv) one iteration: mass seen from the rows equals mass seen from the columns
    st = sinkhorn._State(ops, np.ones(n), sinkhorn.Dist(enabled=False))
    st.u.copy_(st.f), st.v.copy_(st.g)
    sinkhorn._sweep(ops, st, sinkhorn.Dist(enabled=False), eps, 0.1 / 0.15, 5.0 / 5.05, np.log(1000.0), False)
    assert bool(torch.isfinite(st.f).all()) and bool(torch.isfinite(st.g).all())
    Lr = ops.row_lse(st.g, eps)
    Lc = ops.col_lse(st.f, eps)
    mass_rows = float(torch.exp(st.f / eps + Lr).sum())
    mass_cols = float(torch.exp(st.g / eps + Lc).sum())
    assert mass_rows == pytest.approx(mass_cols, rel=1e-5)


@pytest.mark.timeout(600)
def test_parity_8192x8192_three_way(ot):
    """SURVEY 8d's largest parity shape: the streamed tensor-core solve, libot_b200.so (the reference's dense ABI on
    the device) and the reference's own ot_func.cpp compiled unmodified (oracle/_ref), same inputs, d = 32.
    Tolerances are the north star's: marginals 1e-5 relative, plan entries rtol 1e-4, transition table 1e-4 with the
    same argmax; the two dense fp64 paths must agree to 1e-9."""
    from oracle import ref_lib
    from spadot_b200 import ot_func
    if not ref_lib.available():
        pytest.skip("oracle/_ref/libot_ref.so (the compiled reference) is not present")
    ot_solvers, sinkhorn, CudaOps = ot
    n = m = 8192
    a, b, la, lb = ot_dense.synthetic_embeddings(n, m, 32, seed=8192)
    Cn, med = ot_dense.median_normalised_cost(a, b)
    kw = dict(lambda1=CFG["lambda1"], lambda2=CFG["lambda2"], epsilon=CFG["epsilon"], tau=CFG["tau"],
              batch_size=CFG["batch_size"], tolerance=CFG["tolerance"], epsilon0=CFG["epsilon0"])
    want = ref_lib.duality_gap_solve(Cn, np.ones(n), **kw)
    dense_dev = ref_lib.duality_gap_solve(Cn, np.ones(n), L=ref_lib.bind(ot_func.LIB_PATH), **kw)
    big = want > 1e-8 * want.max()
    assert (np.abs(dense_dev - want)[big] / want[big]).max() < 1e-9
    del dense_dev
    ops = CudaOps(a, b)
    assert ops.use_tc
    cp = ot_solvers.solve_coupling(a, b, CFG, ops=ops)
    assert cp.median == pytest.approx(med, rel=1e-13)      # exact order statistic; the oracle's cdist may differ in the last bit
    got = cp.plan().cpu().numpy()
    assert np.abs(got.sum(1) - want.sum(1)).max() / want.sum(1).max() < 1e-5
    assert np.abs(got.sum(0) - want.sum(0)).max() / want.sum(0).max() < 1e-5
    assert (np.abs(got - want)[big] / want[big]).max() < 1e-4
    tab = cp.transition_table(la, lb, 10, 10).cpu().numpy()
    tab_ref = ot_dense.transition_table(want, la, lb, 10, 10)
    assert np.abs(tab - tab_ref).max() / tab_ref.max() < 1e-4
    assert np.array_equal(tab.argmax(1), tab_ref.argmax(1))


def _six_sweeps(sinkhorn, ops, n, dist, corrupt=None):
    """3 + 3 sweeps at the final-stage parameters through the public driver; optional tampering in between."""
    st = sinkhorn._State(ops, np.ones(n), dist)
    st.u.copy_(st.f)
    st.v.copy_(st.g)
    args = (0.05, 0.1 / 0.15, 5.0 / 5.05, float(np.log(1000.0)))
    sinkhorn._sweeps(ops, st, dist, *args, False, 3)
    if corrupt is not None:
        corrupt(ops)
    sinkhorn._sweeps(ops, st, dist, *args, False, 3)
    torch.cuda.synchronize()
    return st


@pytest.mark.parametrize("loop", ["native", "python"])
def test_predicted_stabiliser_matches_tracking_and_falls_back(ot, loop):
    """Predicted stabiliser of the tensor-core pass (sdb_lse_pass_tc_pred): same potentials as the tracking kernel; a
    prediction tampered with between two batches is caught by the finalize check, the batch is redone from the snapshot
    with tracking, and the result is still the tracking result."""
    _, sinkhorn, CudaOps = ot
    n, m = 3000, 2600
    a, b, _, _ = ot_dense.synthetic_embeddings(n, m, 32, seed=5)
    dist = sinkhorn.Dist(enabled=False)

    def make(predict):
        ops = CudaOps(a, b, tc="on")
        ops.set_median(160.0)
        ops.PREDICT, ops.PREDICT_MIN_PAIRS = predict, 0
        if loop == "python":
            ops.fused_sweeps = None                      # the per-iteration path used for world > 1 and by bench.py
        return ops

    ref = _six_sweeps(sinkhorn, make(False), n, dist)
    ops = make(True)
    got = _six_sweeps(sinkhorn, ops, n, dist)
    assert ops._pred is not None and ops._pred.ok and ops._pred.fresh["x"] is not None      # predictions really ran
    assert float((got.f - ref.f).abs().max()) < 2e-6 and float((got.g - ref.g).abs().max()) < 2e-6

    def tamper(o):
        o._pred.m["x"] += 400.0                           # every term of the next row pass underflows
    ops2 = make(True)
    got2 = _six_sweeps(sinkhorn, ops2, n, dist, corrupt=tamper)
    assert ops2._pred.ok is False                         # the misprediction was detected ...
    assert float((got2.f - ref.f).abs().max()) < 2e-6 and float((got2.g - ref.g).abs().max()) < 2e-6   # ... and repaired

    def tamper_low(o):
        o._pred.m["y"] -= 400.0                           # terms of the next column pass overflow
    ops3 = make(True)
    got3 = _six_sweeps(sinkhorn, ops3, n, dist, corrupt=tamper_low)
    assert ops3._pred.ok is False
    assert float((got3.f - ref.f).abs().max()) < 2e-6 and float((got3.g - ref.g).abs().max()) < 2e-6


def test_predicted_stabiliser_full_solve_same_iterations(ot):
    ot_solvers, sinkhorn, CudaOps = ot
    a, b, _, _ = ot_dense.synthetic_embeddings(2600, 2100, 20, seed=9)
    out = []
    for predict in (False, True):
        ops = CudaOps(a, b, tc="on")
        ops.PREDICT, ops.PREDICT_MIN_PAIRS = predict, 0
        out.append(ot_solvers.solve_coupling(a, b, dict(CFG), ops=ops, dist=sinkhorn.Dist(enabled=False)))
    cp0, cp1 = out
    assert cp1.ops._pred is not None and cp1.ops._pred.ok
    assert cp0.info["iters_per_stage"] == cp1.info["iters_per_stage"]
    assert float((cp0.f - cp1.f).abs().max()) < 2e-6 and float((cp0.g - cp1.g).abs().max()) < 2e-6


def test_transition_table_on_the_tensor_core_pass(ot):
    """K6 on the tcgen05 pass (label-permuted, tile-padded target spots, one variable-length split per label,
    sdb_lse_pass_tc_groups) against the dense oracle P0^T.T.P1 (ref: utils/_analyze_utils.py:135-137) and against the SIMT
    form; labels with an empty class, a class larger than one split, and spots outside [0, k1)."""
    ot_solvers, sinkhorn, CudaOps = ot
    n, m, d = 2300, 3100, 32
    a, b, la, lb = ot_dense.synthetic_embeddings(n, m, d, seed=41)
    lb = lb.copy()
    lb[lb == 3] = 4                       # class 3 is empty
    lb[:5] = 12                           # outside [0, 10): takes no part
    Cn, med = ot_dense.median_normalised_cost(a, b)
    want = ot_dense.duality_gap_solve(Cn, np.ones(n), **CFG)
    keep = lb < 10
    tab_ref = ot_dense.transition_table(want[:, keep], la, lb[keep], 10, 10)
    tabs = {}
    for tc in ("on", "off"):
        ops = CudaOps(a, b, tc=tc)
        if tc == "on":
            ops.MAX_SPLIT_COLS = 512      # forces classes to be cut into several splits
        cp = ot_solvers.solve_coupling(a, b, CFG, ops=ops, dist=sinkhorn.Dist(enabled=False))
        tabs[tc] = cp.transition_table(la, lb, 10, 10).cpu().numpy()
    for tc in ("on", "off"):
        assert np.abs(tabs[tc] - tab_ref).max() / tab_ref.max() < 1e-4, tc
        assert np.array_equal(tabs[tc].argmax(1), tab_ref.argmax(1))
        assert np.all(tabs[tc][:, 3] == 0.0)
    assert np.abs(tabs["on"] - tabs["off"]).max() / tab_ref.max() < 2e-5


@pytest.mark.parametrize("n,m,d,tau", [(2600, 2100, 20, 1000.0), (3000, 2600, 32, 2.0), (130, 300, 6, 1.5), (5000, 4100, 48, 1000.0)])
@pytest.mark.parametrize("predict", [False, True])
def test_fused_pass_and_update_matches_separate_kernels(ot, n, m, d, tau, predict):
    """sdb_lse_pass_tc_fused (the CTA completing a 128-row tile combines and updates it; tau bookkeeping deferred, two launches
    per iteration) against pass + sdb_finalize_update_pred + sdb_absorb (five launches): same partials, same combine code, so
    the iterates must agree to rounding, the frames after the flush exactly in value; small tau forces absorptions."""
    ot_solvers, sinkhorn, CudaOps = ot
    a, b, _, _ = ot_dense.synthetic_embeddings(n, m, d, seed=n)
    G = np.exp(np.random.default_rng(m).normal(0, 0.3, n))
    out = []
    for fused in (True, False):
        ops = CudaOps(a, b, tc="on")
        ops.FUSED_UPDATE = fused
        ops.PREDICT, ops.PREDICT_MIN_PAIRS = predict, 0
        l0 = ops.launches
        cp = ot_solvers.solve_coupling(a, b, dict(CFG, tau=tau), G=G, ops=ops, dist=sinkhorn.Dist(enabled=False), median=1.9 * d)
        out.append((cp, ops.launches - l0))
    (c1, l1), (c2, l2) = out
    assert l1 < 0.6 * l2, (l1, l2)
    assert c1.info["iters_per_stage"] == c2.info["iters_per_stage"]
    for name in ("f", "g", "u", "v", "Lr", "Lc"):
        x, y = getattr(c1.state, name), getattr(c2.state, name)
        assert float((x - y).abs().max()) < 1e-11, name
