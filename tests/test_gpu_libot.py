"""GPU parity of libot_b200.so (include/libot_b200.h) — the reference's own libot.so ABI on the device — against
the reference's native file compiled unmodified (oracle/_ref/libot_ref.so) through ONE ctypes driver
(oracle/ref_lib.py) pointed at either library, and against the numpy restatement where that is simpler.

fp64 everywhere; the only difference is the order of the long sums, so tolerances are 1e-10 relative or tighter
and iteration counts / return codes must be identical."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import ot_dense, ref_lib

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RTOL = 1e-10


@pytest.fixture(scope="module")
def libs():
    from spadot_b200 import ot_func
    torch.cuda.set_device(0)
    assert ot_func.lib.libot_b200_device_check() == 0
    if not ref_lib.available():
        pytest.skip("oracle/_ref/libot_ref.so (the compiled reference) is not present")
    return ot_func, ref_lib.bind(ot_func.LIB_PATH), ref_lib.lib()


def problem(n, m, d, seed, eps=0.3):
    a, b, _, _ = ot_dense.synthetic_embeddings(n, m, d, seed=seed)
    C, _ = ot_dense.median_normalised_cost(a, b)
    rng = np.random.default_rng(seed)
    u, v = rng.normal(0, 0.05, n), rng.normal(0, 0.05, m)
    return np.ascontiguousarray(C), u, v, rng


def close(x, y, rtol=RTOL):
    np.testing.assert_allclose(x, y, rtol=rtol, atol=1e-300)


@pytest.mark.parametrize("n,m", [(1, 1), (3, 700), (257, 129), (747, 1966)])
def test_update_k_and_update_R(libs, n, m):
    ot_func, dev, ref = libs
    C, u, v, rng = problem(n, m, 6, n + m)
    a, b = np.exp(rng.normal(0, 1, n)), np.exp(rng.normal(0, 1, m))
    out = {}
    for name, L in (("dev", dev), ("ref", ref)):
        K, K_, R = np.zeros_like(C), np.zeros_like(C), np.zeros_like(C)
        L.update_k_double(ref_lib._ptr(K), ref_lib._ptr(K_), ref_lib._ptr(C), ref_lib._ptr(u), ref_lib._ptr(v), 0.3, n, m)
        L.update_R_double(ref_lib._ptr(R), ref_lib._ptr(K), ref_lib._ptr(a), ref_lib._ptr(b), n, m)
        out[name] = (K, K_, R)
    for x, y in zip(out["dev"], out["ref"]):
        close(x, y, 1e-13)


def test_float_entry_points_match_numpy(libs):
    ot_func, _, _ = libs
    C, u, v, rng = problem(130, 97, 5, 3)
    C32, u32, v32 = C.astype(np.float32), u.astype(np.float32), v.astype(np.float32)
    K, K_ = np.zeros_like(C32), np.zeros_like(C32)
    ot_func.update_K_c(K, K_, C32, u32, v32, 0.3, use_float=True)
    np.testing.assert_allclose(K_, np.exp(-C32 / np.float32(0.3)), rtol=2e-6)
    np.testing.assert_allclose(K, np.exp((u32[:, None] + v32[None, :] - C32) / np.float32(0.3)), rtol=2e-6)
    a, b = np.exp(rng.normal(0, 1, 130)).astype(np.float32), np.exp(rng.normal(0, 1, 97)).astype(np.float32)
    R = np.zeros_like(C32)
    ot_func.update_R_c(R, K, a, b, use_float=True)
    np.testing.assert_allclose(R, K * a[:, None] * b[None, :], rtol=1e-6)
    dx, dy = np.ones(130) / 130, np.ones(97) / 97
    p, q = np.ones(130), np.ones(97)
    C64 = C32.astype(np.float64)
    want = ot_dense.primal(C64, K_.astype(np.float64), R.astype(np.float64), dx, dy, p, q, 0.3, 0.1, 5.0)
    got = ot_func.primal_c(C32, K_, R, dx, dy, p, q, a, b, 0.3, 0.1, 5.0, use_float=True)
    assert got == pytest.approx(want, rel=1e-4)
    wantd = ot_dense.dual(K_.astype(np.float64), R.astype(np.float64), dx, dy, p, q, a.astype(np.float64),
                          b.astype(np.float64), 0.3, 0.1, 5.0)
    gotd = ot_func.dual_c(C32, K_, R, dx, dy, p, q, a, b, 0.3, 0.1, 5.0, use_float=True)
    assert gotd == pytest.approx(wantd, rel=1e-4, abs=1e-5)
    gap = ot_func.compute_duality_gap_c(C32, K_, R, dx, dy, p, q, a, b, 0.3, 0.1, 5.0, use_float=True)
    assert gap == pytest.approx((want - wantd) / abs(want), rel=1e-3)
    assert ot_func.dummy_c(C, K_, R, dx, dy, p, q, a, b, 0.3, 0.1, 5.0) == 0.0


@pytest.mark.parametrize("n,m,tau", [(64, 80, 1000.0), (300, 411, 1.05), (747, 1966, 2.0), (5, 1200, 1000.0)])
def test_step1_process_matches_compiled_reference(libs, n, m, tau):
    """`iters` iterations incl. stabilisation: small tau forces absorptions (K rebuilt, a,b reset, u,v moved)."""
    _, dev, ref = libs
    eps, l1, l2 = 0.2, 0.1, 5.0
    C, u0, v0, rng = problem(n, m, 8, n * 7 + m)
    K0 = np.exp((u0[:, None] + v0[None, :] - C) / eps)
    dx, dy = np.ones(n) / n, np.ones(m) / m
    p, q = np.exp(rng.normal(0, 0.2, n)), np.ones(m)
    res = {}
    for name, L in (("dev", dev), ("ref", ref)):
        a, b = np.ones(n), np.ones(m)
        oa, ob = a.copy(), b.copy()
        K, u, v = K0.copy(), u0.copy(), v0.copy()
        its = []
        cur = 0
        for _ in range(3):                           # three calls of 4 iterations, state carried by the caller
            cur = ref_lib.step1(a, b, oa, ob, K, C, dx, dy, p, q, u, v, cur, 10 ** 7, 4, tau, l1, l2, l1 / (l1 + eps),
                                l2 / (l2 + eps), eps, L=L)
            its.append(cur)
        res[name] = (its, a, b, oa, ob, K, u, v)
    assert res["dev"][0] == res["ref"][0] == [4, 8, 12]
    for x, y in zip(res["dev"][1:], res["ref"][1:]):
        close(x, y)
    if tau < 10:
        assert np.abs(res["ref"][6] - u0).max() > 0, "the case was meant to absorb at least once"


def test_step1_process_max_iter_returns_minus_one(libs, capfd):
    _, dev, ref = libs
    C, u, v, _ = problem(40, 30, 4, 9)
    eps = 0.3
    out = []
    for L in (dev, ref):
        K = np.exp(-C / eps)
        a, b = np.ones(40), np.ones(30)
        oa, ob = a.copy(), b.copy()
        uu, vv = np.zeros(40), np.zeros(30)
        r = ref_lib.step1(a, b, oa, ob, K, C, np.ones(40) / 40, np.ones(30) / 30, np.ones(40), np.ones(30), uu, vv, 0, 3, 5,
                          1000.0, 0.1, 5.0, 0.25, 0.94, eps, L=L)
        out.append((r, a, b, oa, ob))
    assert out[0][0] == out[1][0] == -1
    for x, y in zip(out[0][1:], out[1][1:]):        # three iterations were applied before giving up
        close(x, y)


@pytest.mark.parametrize("n,m,d,seed,lam", [(90, 70, 20, 1, (0.1, 5.0)), (300, 411, 20, 2, (1.0, 50.0)), (747, 1966, 20, 3, (0.1, 5.0))])
def test_full_solve_through_update_process(libs, n, m, d, seed, lam):
    """optimal_transport_duality_gap's stock path (ot_solvers.py:264-290): update_k + update_process per epsilon stage,
    driven by the same Python for both libraries."""
    _, dev, ref = libs
    a, b, _, _ = ot_dense.synthetic_embeddings(n, m, d, seed=seed)
    C, _ = ot_dense.median_normalised_cost(a, b)
    G = np.exp(np.random.default_rng(seed).normal(0, 0.3, n))
    i_dev, i_ref = {}, {}
    T_dev = ref_lib.duality_gap_solve(C, G, lam[0], lam[1], 0.05, info=i_dev, L=dev)
    T_ref = ref_lib.duality_gap_solve(C, G, lam[0], lam[1], 0.05, info=i_ref, L=ref)
    assert i_dev["gap"] <= 1e-8 and i_ref["gap"] <= 1e-8
    close(T_dev.sum(1), T_ref.sum(1), 1e-9)
    close(T_dev.sum(0), T_ref.sum(0), 1e-9)
    big = T_ref > 1e-12 * T_ref.max()
    assert (np.abs(T_dev - T_ref)[big] / T_ref[big]).max() < 1e-8
    np.testing.assert_allclose(i_dev["f"], i_ref["f"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(i_dev["g"], i_ref["g"], rtol=0, atol=1e-9)
    close(i_dev["K"], i_ref["K"], 1e-9)


def test_standalone_primal_dual_gap(libs):
    _, dev, ref = libs
    C, u, v, rng = problem(211, 333, 7, 11)
    eps = 0.05
    K_ = np.exp(-C / eps)
    a, b = np.exp(rng.normal(0, 0.5, 211)), np.exp(rng.normal(0, 0.5, 333))
    R = np.ascontiguousarray(K_ * a[:, None] * b[None, :])
    R[3, 5] = 0.0                                     # log(0) -> nan_to_num path (ot_func.cpp:30-40, 413)
    dx, dy = np.ones(211) / 211, np.ones(333) / 333
    p, q = np.exp(rng.normal(0, 0.2, 211)), np.ones(333)
    got = ref_lib.gap_parts(C, K_, R, dx, dy, p, q, a, b, eps, 0.1, 5.0, L=dev)
    want = ref_lib.gap_parts(C, K_, R, dx, dy, p, q, a, b, eps, 0.1, 5.0, L=ref)
    for g, w in zip(got, want):
        assert g == pytest.approx(w, rel=1e-11)


def test_reference_driver_module_runs_on_the_drop_in(libs):
    """spadot_b200.ot_func (the mirror of the reference's ot_func.py) with ndpointer-checked numpy arrays:
    one final-stage update_process_c call equals the compiled reference."""
    ot_func, _, ref = libs
    n, m, eps, l1, l2 = 120, 150, 0.05, 0.1, 5.0
    C, _, _, _ = problem(n, m, 10, 21)
    dx, dy, p, q = np.ones(n) / n, np.ones(m) / m, np.ones(n), np.ones(m)
    got = {}
    for name in ("dev", "ref"):
        u, v, a, b = np.zeros(n), np.zeros(m), np.ones(n), np.ones(m)
        oa, ob = a.copy(), b.copy()
        K, K_, R = np.zeros_like(C), np.zeros_like(C), np.zeros_like(C)
        if name == "dev":
            ot_func.update_K_c(K, K_, C, u, v, eps)
            gap = ot_func.update_process_c(R, a, b, oa, ob, K, K_, C, dx, dy, p, q, u, v, 5, 5, 5, eps, 1e-8, 1000.0, l1, l2,
                                           l1 / (l1 + eps), l2 / (l2 + eps), 0, 1e7)
        else:
            P = ref_lib._ptr
            ref.update_k_double(P(K), P(K_), P(C), P(u), P(v), eps, n, m)
            gap = ref.update_process_double(P(R), P(a), P(b), P(oa), P(ob), P(K), P(K_), P(C), P(dx), P(dy), P(p), P(q), P(u),
                                            P(v), 5, 5, 5, eps, 1e-8, 1000.0, l1, l2, l1 / (l1 + eps), l2 / (l2 + eps), 0,
                                            10 ** 7, n, m)
        got[name] = (gap, R, a, b, oa, ob)
    assert got["dev"][0] <= 1e-8 and got["ref"][0] <= 1e-8
    for x, y in zip(got["dev"][1:], got["ref"][1:]):
        close(x, y, 1e-9)
    launches, h2d, d2h = ot_func.counters()
    assert launches > 0 and h2d > 0 and d2h > 0


# ------------------------------------------------------------------ the UNMODIFIED reference driver on the drop-in
def _reference_driver_on(lib_path, tag):
    from oracle import reference_loader
    if not (reference_loader.available() or reference_loader.compiled_available()):
        pytest.skip("neither /root/reference nor oracle/_ref/pyref (its byte-compiled driver) is present")
    return reference_loader.load_ot_solvers(lib_path=lib_path, tag=tag)


@pytest.mark.parametrize("name", ["ot_small_48x61_d6", "ot_growth_90x70_d20", "ot_medium_300x411_d20", "ot_wotcfg_130x97_d32"])
def test_unmodified_reference_solver_text_runs_on_libot_b200(libs, golden_dir, name):
    """INTEGRATION.md section 5, executed: the reference's own ot_func.py + ot_solvers.py (source where /root/reference is
    mounted, else the byte-compiled copies of oracle/_ref/pyref), with NOTHING changed but the library that
    `ctypes.cdll.LoadLibrary` (ot_func.py:10) resolves — libot_b200.so instead of libot.so.  Its
    `optimal_transport_duality_gap(use_C=True)` (ot_solvers.py:164-449, native loop :283-290) and `compute_transport_map`
    (:95-121) must reproduce the golden vectors that the same text produced on the reference's shipped libot.so."""
    import contextlib
    import io
    ot_func, _, _ = libs
    ref = _reference_driver_on(ot_func.LIB_PATH, "_spadot_ref_on_b200")
    bound = os.path.realpath(__import__("sys").modules["_spadot_ref_on_b200.ot_func"].lib._name)
    assert bound == os.path.realpath(ot_func.LIB_PATH)
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    cfg = {k: float(v) for k, v in zip(z["cfg_keys"], z["cfg_vals"])}
    for k in ("growth_iters", "batch_size", "scaling_iter", "inner_iter_max", "extra_iter"):
        cfg[k] = int(cfg[k])
    cfg.update(use_C=True, use_Py=False, profiling=False)
    a, b, G = z["a"], z["b"], z["G"]
    C = ot_dense.sqeuclidean(a, b)
    Cn = C / z["median"]
    launches0 = ot_func.counters()[0]
    with contextlib.redirect_stdout(io.StringIO()):
        plan = ref.optimal_transport_duality_gap(C=Cn.copy(), G=G.copy(), **cfg)
        gamma0 = ref.compute_transport_map(a, b, dict(cfg), G=G.copy())
    assert ot_func.counters()[0] > launches0            # the device library did the work
    np.testing.assert_allclose(plan.sum(1), z["gap_row_sums"], rtol=1e-9, atol=1e-300)
    np.testing.assert_allclose(plan.sum(0), z["gap_col_sums"], rtol=1e-9, atol=1e-300)
    if "plan_gap" in z:
        np.testing.assert_allclose(plan, z["plan_gap"], rtol=1e-8, atol=1e-300)
    else:
        np.testing.assert_allclose(plan[z["sample_i"], z["sample_j"]], z["plan_gap_samples"], rtol=1e-8, atol=1e-300)
    np.testing.assert_allclose(gamma0, plan, rtol=1e-9, atol=1e-300)       # gammas[0] of the growth loop, ot_solvers.py:121
