#!/usr/bin/env python
"""bench.py — Sinkhorn iterations/s of the streamed unbalanced OT coupling (BASELINE.json metric).

Workload (config.workload): SYN-OT, one coupling of N x M synthetic spots (default 1M x 1M, latent
dim 32; BASELINE.json configs[3]) at the final-stage regularisation eps=0.05, lambda=(0.1, 5)
(SpaDOT/config.yaml:39-57).  A *step* is one full Sinkhorn iteration = row pass + column pass +
potential updates + tau bookkeeping, i.e. 2*N*M kernel evaluations.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --steps K --warmup W    # the reference's own CPU path

N > 1: launched by torchrun, one rank per GPU; source spots (rows) are partitioned across ranks,
target spots are replicated, the column pass is followed by ONE all-reduce (sum) of the M-vector of
partial sums against a shift every rank already holds.  Total work is fixed => "scaling": "strong".

Reference arm: the reference's UNMODIFIED solver text (`compute_transport_map` ->
`optimal_transport_duality_gap`, SpaDOT/utils/OT_loss/ot_solvers.py:95-449) imported from /root/reference
where mounted, else from its byte-compiled copy in oracle/_ref/pyref, on a dense sample that fits host
memory, rescaled by N*M; BLAS threads are set explicitly (torchrun exports OMP_NUM_THREADS=1).
"""
import argparse
import contextlib
import io
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

EPS, LAM1, LAM2, TAU = 0.05, 0.1, 5.0, 1000.0
METRIC, UNIT = "sinkhorn_iters_per_sec", "iter/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="spadot_b200", choices=["spadot_b200", "reference"])
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--m", type=int, default=1_000_000)
    ap.add_argument("--d", type=int, default=32)
    ap.add_argument("--cpu-sample", type=int, default=None,
                    help="rows=cols of the dense CPU sample (default 8192 for --impl reference, 4096 for the cpu_baseline leg)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-aux", action="store_true", help="skip the train-epoch aux measurement (N=1 only)")
    ap.add_argument("--no-full-solve", action="store_true", help="skip the end-to-end full solve (e2e falls back to fixed steps)")
    ap.add_argument("--tc", default="auto", choices=["auto", "on", "off"], help="tensor-core (tcgen05) pass")
    return ap.parse_args()


def workload_name(a):
    return f"SYN-OT {a.n}x{a.m} d={a.d} eps={EPS} lambda=({LAM1},{LAM2}) fp32-tile/fp64-vector"


def synth(n, m, d, seed=1993):
    """Mixture of 10 Gaussians in R^d; fp64 like the reference's inputs."""
    rng = np.random.default_rng(seed)
    centres = rng.normal(0.0, 1.5, size=(10, d))
    x = centres[rng.integers(0, 10, n)] + rng.normal(0.0, 0.5, size=(n, d))
    y = centres[rng.integers(0, 10, m)] + rng.normal(0.0, 0.5, size=(m, d)) + 0.1
    return x, y


def ot_config():
    return dict(growth_iters=1, epsilon=EPS, epsilon0=1.0, lambda1=LAM1, lambda2=LAM2, tau=TAU, scaling_iter=3000,
                inner_iter_max=50, tolerance=1e-8, max_iter=1e7, batch_size=5, extra_iter=1000, use_Py=False, use_C=True,
                profiling=False)


# ----------------------------------------------------------------------------------------- CPU arms
def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


@contextlib.contextmanager
def blas_threads(n):
    """Set the BLAS / OpenMP pools to n threads whatever the launcher exported (torchrun sets OMP_NUM_THREADS=1 for
    nproc > 1, which silently cut round 1's multi-rank reference runs to one thread) and report what is really used."""
    used = {"threads": 1, "how": "threadpoolctl unavailable: library defaults"}
    try:
        from threadpoolctl import threadpool_info, threadpool_limits
        with threadpool_limits(limits=n):
            pools = [p for p in threadpool_info() if p.get("user_api") in ("blas", "openmp")]
            blas = [p["num_threads"] for p in pools if p.get("user_api") == "blas"]
            used["threads"] = int(max(blas)) if blas else int(max([p["num_threads"] for p in pools] or [1]))
            used["how"] = "threadpoolctl.threadpool_limits(%d); pools %s" % (n, [(p.get("internal_api"), p["num_threads"]) for p in pools])
            yield used
    except ImportError:
        yield used


def cpu_reference_arm(a, sample):
    """One full stock solve of the reference on a dense s x s sample of the workload (cost matrix + median + six
    epsilon stages to 1e-8, everything timed), per leg:
      (i)  use_C=True  — SpaDOT's default config: libot's native loop, single-threaded by construction;
      (ii) use_C=False, use_Py=True — the numpy/BLAS text, which is what wot runs in `SpaDOT analyze`.
    Iterations of that very solve are counted by the oracle port on the same inputs (identical counts are a tested
    property, tests/test_oracle_golden.py); value = iterations / wall time of the fastest leg, rescaled by N*M."""
    from oracle import ot_dense, reference_loader
    s = int(sample)
    x, y = synth(s, s, a.d)
    cfg = ot_config()
    scale = (float(s) * float(s)) / (float(a.n) * float(a.m))
    legs = []
    with blas_threads(host_threads()) as used:
        t0 = time.perf_counter()
        Cn, _ = ot_dense.median_normalised_cost(x, y)
        info = {}
        ot_dense.duality_gap_solve(Cn, np.ones(s), info=info, **{k: cfg[k] for k in
                                   ("lambda1", "lambda2", "epsilon", "batch_size", "tolerance", "tau", "epsilon0", "max_iter")})
        port_s = time.perf_counter() - t0
        del Cn
        iters = int(info["total_iters"])
        legs.append(("port", used["threads"], "oracle/ot_dense.py (numpy restatement, BLAS threads)", port_s))
        have_ref = reference_loader.available() or reference_loader.compiled_available()
        if have_ref:
            ref = reference_loader.load_ot_solvers()
            src = "source under /root/reference" if reference_loader.available() else "byte-compiled oracle/_ref/pyref + oracle/_ref/libot_ref.so"
            with contextlib.redirect_stdout(io.StringIO()):
                ref.compute_transport_map(*synth(256, 256, a.d), dict(cfg))                  # untimed warm-up (imports, page-in)
            for label, use_C, use_Py, thr in (("use_C=True: native libot loop, 1 thread by construction", True, False, 1),
                                              ("use_C=False,use_Py=True: numpy/BLAS text (= wot's solver)", False, True, used["threads"])):
                t0 = time.perf_counter()
                with contextlib.redirect_stdout(io.StringIO()):
                    ref.compute_transport_map(x, y, dict(cfg, use_C=use_C, use_Py=use_Py))
                legs.append(("reference", thr, f"reference compute_transport_map ({src}; {label})", time.perf_counter() - t0))
            legs = legs[1:]          # the port only counted the iterations
        kind, cores, what, dt = min(legs, key=lambda r: r[3])
        others = "; ".join(f"{w}: {iters / t:.3f} iter/s in {t:.1f}s on {c} thread(s)" for _, c, w, t in legs)
        return dict(value=iters / dt * scale, unit=UNIT, cores=cores, kind=kind,
                    sample=f"ONE stock solve of a dense fp64 {s}x{s} d={a.d} sample (sqeuclidean cost + median + 6 eps stages to 1e-8 = "
                           f"{iters} Sinkhorn iterations, all timed); legs at sample size: [{others}]; fastest rescaled by the N*M ratio "
                           f"{scale:.3e} to the full workload; host threads available {host_threads()}, BLAS: {used['how']}",
                    iterations=iters, sample_seconds=dt), dt / iters


def reference_main(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    s = a.cpu_sample or 8192
    base, sec_per_iter = cpu_reference_arm(a, s)
    scale = (float(s) ** 2) / (float(a.n) * float(a.m))
    line = dict(metric=METRIC, value=base["value"], unit=UNIT, n_gpus=a.gpus, steps=a.steps, warmup=a.warmup,
                ms_per_step=sec_per_iter * 1e3 / scale, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64",
                data="synthetic", impl="reference", config=dict(workload=workload_name(a)), cpu_baseline=base,
                e2e=dict(value=base["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.rows:
            if t < t0 or t > t1 + 0.2:
                continue
            parts = [p.strip() for p in line.split(",")]
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except Exception:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))


# ----------------------------------------------------------------------------------------- GPU arm
def gpu_main(a):
    import torch
    import torch.distributed as td
    from spadot_b200 import ot_solvers, sinkhorn
    from spadot_b200.cuda_ops import CudaOps

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        td.init_process_group("nccl", device_id=dev)
    dist = sinkhorn.Dist(enabled=world > 1)
    dist.native_comm(dev)          # the library's own NCCL communicator (multi-rank native loop): create it outside any timed region

    # rows partitioned across ranks; identical RNG stream on every rank
    r0, r1 = (a.n * rank) // world, (a.n * (rank + 1)) // world
    x_all, y = synth(a.n, a.m, a.d)
    x = np.ascontiguousarray(x_all[r0:r1])
    del x_all
    # fixed normalisation: E|x-y|^2 for this mixture (the exact median is K5's job; it is inside the full-solve e2e)
    median = float(2 * a.d * (1.5 ** 2 + 0.5 ** 2))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    ops = CudaOps(x, y, device=dev, tc=a.tc)
    ops.set_median(median)
    st = sinkhorn._State(ops, np.ones(ops.n), dist)
    st.u.copy_(st.f)
    st.v.copy_(st.g)
    a1, a2 = LAM1 / (LAM1 + EPS), LAM2 / (LAM2 + EPS)
    log_tau = math.log(TAU)
    pass_events = []

    def step():
        sinkhorn._sweep(ops, st, dist, EPS, a1, a2, log_tau, False)

    # instrument the dominant kernel (the LSE pass) with CUDA events on the launching stream
    class _T:
        on = False
    timed_lse = _T
    orig_call = ops._call

    def call_hook(name, *args):
        if timed_lse.on and name.startswith("sdb_lse_pass"):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            orig_call(name, *args)
            e1.record()
            pass_events.append((e0, e1))
        else:
            orig_call(name, *args)
    ops._call = call_hook

    for _ in range(a.warmup):
        step()
    barrier()
    clocks = ClockSampler(local) if rank == 0 else None
    launches0 = ops.launches
    coll0 = dist.collectives
    timed_lse.on = True
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    ev0.record()
    for _ in range(a.steps):
        step()
    ev1.record()
    barrier()
    t_wall1 = time.perf_counter()
    timed_lse.on = False
    collectives_per_step = (dist.collectives - coll0) / a.steps if world > 1 else 0      # read before anything else runs
    elapsed = torch.tensor([ev0.elapsed_time(ev1) / 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        td.all_reduce(elapsed, op=td.ReduceOp.MAX)
    elapsed = float(elapsed.item())
    launches = ops.launches - launches0
    pass_ms = [e0.elapsed_time(e1) for e0, e1 in pass_events]
    clk = clocks.stop(t_wall0, t_wall1) if clocks else None
    ops._call = orig_call
    finite = bool(torch.isfinite(st.f).all().item() and torch.isfinite(st.g).all().item())

    # ---- parity of the timed iterates (every N runs the same W+K iterations from the same inputs: the SCALE lines must
    # carry equal checksums, and the row LSE at the final g is checked against the fp64 oracle on 32 rows of rank 0)
    f_sum = float(dist.sum_(st.f.sum().reshape(1).clone()).item())
    f_abs = float(dist.sum_(st.f.abs().sum().reshape(1).clone()).item())
    g_sum, g_abs = float(st.g.sum().item()), float(st.g.abs().sum().item())
    Lr_dev = ops.row_lse(st.g, EPS)
    parity = dict(f_checksum=f_sum, f_abs_checksum=f_abs, g_checksum=g_sum, g_abs_checksum=g_abs, iterations=a.warmup + a.steps)
    if rank == 0:
        from oracle import ot_logdomain                       # the checker, never the thing measured
        idx = np.random.default_rng(7).integers(0, ops.n, 32)
        g_host = st.g.cpu().numpy()
        want = ot_logdomain.CostOperator(x[idx], y, median=median, block=8).row_lse(g_host / EPS, EPS)
        got = Lr_dev.cpu().numpy()[idx]
        f_host = st.f.cpu().numpy()[idx]
        parity.update(rows_checked=32, lse_max_abs_err=float(np.abs(got - want).max()), lse_mean_err=float((got - want).mean()),
                      checker="oracle/ot_logdomain.py fp64 row LSE at the final g", f_rows_sample=[float(v) for v in f_host[:4]])

    # ---- e2e (i): the public call a user makes — host buffers in, potentials on the host out; exact median + six
    # epsilon stages to 1e-8; host<->device copies inside the timed region.  BASELINE's "OT wall-time at N x M".
    xh = torch.from_numpy(x).pin_memory()
    yh = torch.from_numpy(y).pin_memory()
    del ops, st
    torch.cuda.empty_cache()
    cfg = ot_config()
    full = None
    if not a.no_full_solve:
        ws = max(2048, min(20000, a.n // 8))
        cpw = ot_solvers.solve_coupling(x[:max(1, ws // world)], y[:ws], cfg, dist=dist, device=dev)    # untimed warm-up (first-use costs)
        del cpw
        barrier()
        t0 = time.perf_counter()
        cp = ot_solvers.solve_coupling(xh, yh, cfg, dist=dist, device=dev)
        f_host, g_host2 = cp.f.cpu(), (cp.g.cpu() if rank == 0 else None)
        barrier()
        tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            td.all_reduce(tt, op=td.ReduceOp.MAX)
        full = dict(seconds=float(tt.item()), iterations=int(cp.info["total_iters"]), iters_per_stage=cp.info["iters_per_stage"],
                    gap=cp.info["gap"], median=cp.median, median_sweeps=cp.info.get("sweeps"),
                    h2d=(xh.numel() + yh.numel()) * 8, d2h=f_host.numel() * 8 + (g_host2.numel() * 8 if g_host2 is not None else 0))
        del cp
        torch.cuda.empty_cache()

    # ---- e2e (ii): fixed number of iterations from host buffers (same work per step as the timed region above)
    def e2e_fixed(n_steps):
        ops2 = CudaOps(xh, yh, device=dev, tc=a.tc)  # H2D of both spot sets + point preparation
        ops2.set_median(median)
        st2 = sinkhorn._State(ops2, np.ones(ops2.n), dist)
        st2.u.copy_(st2.f)
        st2.v.copy_(st2.g)
        sinkhorn._sweeps(ops2, st2, dist, EPS, a1, a2, log_tau, False, n_steps)   # native loop when single-rank
        out = st2.f.cpu(), (st2.g.cpu() if rank == 0 else None)      # D2H of the result
        return out + (ops2,)

    e2e_fixed(1)                                     # warm-up call (allocator growth, first-use costs), untimed
    barrier()
    t0 = time.perf_counter()
    f2, g2, ops2 = e2e_fixed(a.steps)
    barrier()
    e2e_t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        td.all_reduce(e2e_t, op=td.ReduceOp.MAX)
    e2e_t = float(e2e_t.item())
    h2d = (xh.numel() + yh.numel()) * 8
    d2h = f2.numel() * 8 + (g2.numel() * 8 if g2 is not None else 0)
    fixed = dict(value=a.steps / e2e_t, unit=UNIT, steps=a.steps, seconds=e2e_t, h2d_bytes_per_step=h2d / a.steps,
                 d2h_bytes_per_step=d2h / a.steps, api="CudaOps(host x, host y) + sinkhorn sweeps + potentials to host")

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        sm_max = float(peaks.get("sm_max_mhz", 1965.0))
        # algorithmic work of one pass launch on this rank: n_p*n_q pair evaluations,
        # each (2*dpad + 8) fp32 flop in the SIMT form and one ex2.
        n_loc, use_tc = ops2.n, bool(ops2.use_tc)
        pairs = float(n_loc) * float(a.m)
        flop_per_pair = 2 * ops2.X.dpad + 8
        avg_pass_s = (sum(pass_ms) / len(pass_ms)) / 1e3 if pass_ms else float("nan")
        n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
            key = f"lse_pass_{'tc' if ops2.use_tc else 'simt'} n={a.n} m={a.m} d={a.d} world={world}"
            traffic = tj.get(key, {}).get("dram_bytes_per_launch")
        except Exception:
            pass
        measured = None
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import pipe_peaks
            del ops2
            torch.cuda.empty_cache()
            measured = pipe_peaks.measure(dev)
        except Exception as exc:
            measured = dict(error=repr(exc)[:200])
        common = dict(traffic=traffic, traffic_source="ncu --set full, profiles/r2_traffic.json" if traffic else None,
                      algorithmic_hbm_bytes_per_launch=(n_loc + a.m) * (32 * 4 + 4) if a.d <= 32 else None,
                      pass_ms_avg=avg_pass_s * 1e3, pass_launches=len(pass_ms),
                      share_of_step=(sum(pass_ms) / 1e3) / elapsed, pairs_per_launch=pairs, measured_pipe_peaks=measured)
        if use_tc:
            # tensor-core form: the dot product runs on tcgen05; what is left per pair is one ex2 (SFU, 16/clk/SM)
            # and ~4 fp32 instructions, so the binding pipe is the SFU.
            sfu_nominal = n_sm * 16 * sm_max * 1e6 / 1e12
            achieved = pairs / avg_pass_s / 1e12
            sfu_measured = measured.get("mufu_ex2", {}).get("value") if isinstance(measured, dict) else None
            roofline = dict(bound="sfu", kernel="lse_pass_tc_kernel (sdb_lse_pass_tc_pred)", achieved=achieved,
                            peak=sfu_nominal, unit="Tex2/s", frac=achieved / sfu_nominal,
                            peak_source=f"nominal: {n_sm} SM x 16 MUFU lanes x {sm_max:.0f} MHz (sm_max_mhz of MEASURED_PEAKS.json, which has "
                                        "no SFU entry); measured beside it by the MUFU.EX2 micro-benchmark of this run (sdb_pipe_peak)",
                            peak_measured=sfu_measured, frac_of_measured=(achieved / sfu_measured) if sfu_measured else None,
                            frac_at_sampled_clock=(achieved / (n_sm * 16 * clk["sm_mhz"] * 1e6 / 1e12)) if clk and clk.get("sm_mhz") else None,
                            **common)
        else:
            fp32_peak = n_sm * 128 * 2 * sm_max * 1e6 / 1e12
            achieved = pairs * flop_per_pair / avg_pass_s / 1e12
            roofline = dict(bound="fp32", kernel="pair_tile_kernel<LseEpi> (sdb_lse_pass_simt)", achieved=achieved,
                            peak=fp32_peak, unit="TFLOP/s", frac=achieved / fp32_peak,
                            peak_source=f"{n_sm} SM x 128 FMA x 2 x {sm_max:.0f} MHz (sm_max_mhz of MEASURED_PEAKS.json; "
                                        "that file has no fp32 entry)",
                            ex2_per_s=pairs / avg_pass_s, **common)
        if full is not None:
            e2e = dict(value=full["iterations"] / full["seconds"], unit=UNIT, h2d_bytes_per_step=full["h2d"] / full["iterations"],
                       d2h_bytes_per_step=full["d2h"] / full["iterations"], steps=full["iterations"], seconds=full["seconds"],
                       api="spadot_b200.ot_solvers.solve_coupling(host x, host y, config): exact median (K5) + six eps stages to "
                           "tolerance 1e-8, potentials copied back to the host", iters_per_stage=full["iters_per_stage"],
                       gap=full["gap"], median=full["median"], median_sweeps=full["median_sweeps"], fixed_steps=fixed)
        else:
            e2e = dict(fixed)
        line = dict(metric=METRIC, value=a.steps / elapsed, unit=UNIT, n_gpus=world, steps=a.steps, warmup=a.warmup,
                    ms_per_step=elapsed * 1e3 / a.steps, higher_is_better=True, scaling="strong", vs_baseline=None,
                    dtype="f16x2-split tensor tiles + f32 epilogue / f64 vectors" if use_tc else "f32 tiles / f64 vectors", data="synthetic", impl="spadot_b200",
                    config=dict(workload=workload_name(a), rows_per_rank=n_loc, parallelism=f"row-partition x{world}",
                                l2_policy="inputs per pass (>=136 MB at 1M) exceed reuse; potentials rewritten every step",
                                median="analytic in the timed steps (K5 exact median is inside e2e)", finite=finite,
                                collectives_per_step=collectives_per_step),
                    e2e=e2e, e2e_full_solve_s=full["seconds"] if full else None, parity=parity,
                    gpu_launches=launches, clocks=clk, roofline=roofline)
        if world == 1 and not a.no_cpu_baseline:
            base, _ = cpu_reference_arm(a, a.cpu_sample or 4096)
            line["cpu_baseline"] = base
        if world == 1 and not a.no_aux:
            # BASELINE.json's metric also names "train s/epoch": the train inner loop (SVGP + GAT) at ChickenHeart
            # shapes, reference formulas in plain torch vs this repo's modules, both on this GPU (tools/train_epoch_bench.py)
            try:
                sys.path.insert(0, os.path.join(ROOT, "tools"))
                import train_epoch_bench
                torch.cuda.empty_cache()
                line["aux_train_epoch"] = train_epoch_bench.run(dev)
            except Exception as exc:       # the aux number must never take the headline down
                line["aux_train_epoch"] = dict(error=repr(exc)[:200])
            # BASELINE configs[2] ("SVGP+GAT training throughput" at 100k spots x 2000 SVGs): one timepoint of that shape, six
            # optimiser steps on 512-seed 2-hop batches (tools/syn_t_bench.py); the reference cannot run this shape at all
            try:
                import statistics
                import syn_t_bench
                torch.cuda.empty_cache()
                r = syn_t_bench.run(["--batches", "6", "--epoch2"])
                line["aux_syn_t_step"] = dict(workload=r["workload"], subgraph_nodes=r["subgraph_nodes"], subgraph_edges=r["subgraph_edges"],
                                              train_step_ms_median=1e3 * statistics.median(r["step_times_s"]),
                                              step_times_s=r["step_times_s"], seeds_per_s_median=512.0 / statistics.median(r["step_times_s"]),
                                              second_epoch=r.get("second_epoch"),
                                              graph_build_s=r["graph_build_s"], sample_batch_s=r["sample_batch_s"],
                                              all_latent_samples_s=r["all_latent_samples_s"], peak_mem_gb=r["peak_mem_gb"],
                                              device_time_ms={k: round(v["ms"], 3) for k, v in (r["step_breakdown_device_time"] or {}).items()
                                                              if isinstance(v, dict)})
            except Exception as exc:
                line["aux_syn_t_step"] = dict(error=repr(exc)[:200])
        print(json.dumps(line))
    if world > 1:
        td.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        reference_main(args)
    else:
        gpu_main(args)
