#!/usr/bin/env python
"""bench.py — Sinkhorn iterations/s of the streamed unbalanced OT coupling (BASELINE.json metric).

Workload (config.workload): SYN-OT, one coupling of N x M synthetic spots (default 1M x 1M, latent
dim 32; BASELINE.json configs[3]) at the final-stage regularisation eps=0.05, lambda=(0.1, 5)
(SpaDOT/config.yaml:39-57).  A *step* is one full Sinkhorn iteration = row pass + column pass +
potential updates + tau bookkeeping, i.e. 2*N*M kernel evaluations.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --steps K --warmup W    # the reference's own CPU path (oracle/_ref)

N > 1: launched by torchrun, one rank per GPU; source spots (rows) are partitioned across ranks,
target spots are replicated, the column pass is followed by the all-reduce pair of
spadot_b200.sinkhorn.combine_col_lse.  Total work is fixed => "scaling": "strong".
"""
import argparse
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
    ap.add_argument("--cpu-sample", type=int, default=8192, help="rows=cols of the dense CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-aux", action="store_true", help="skip the train-epoch aux measurement (N=1 only)")
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


# ----------------------------------------------------------------------------------------- CPU arms
def cpu_reference_arm(a, steps, warmup):
    """Times the reference's native inner loop (step1_process_double, ot_func.cpp:690-828, compiled
    unmodified into oracle/_ref) — or the oracle port when that library is absent — on a dense sample
    and rescales to the full workload's pair count (t_iter is proportional to N*M)."""
    from oracle import ot_dense, ref_lib
    s = a.cpu_sample
    x, y = synth(s, s, a.d)
    t0 = time.perf_counter()
    C, _ = ot_dense.median_normalised_cost(x, y)
    K = np.exp(-C / EPS)
    build_s = time.perf_counter() - t0
    I = J = s
    dx, dy = np.ones(I) / I, np.ones(J) / J
    p, q = np.ones(I), np.ones(J)
    u, v = np.zeros(I), np.zeros(J)
    av, bv = np.ones(I), np.ones(J)
    oa, ob = av.copy(), bv.copy()
    a1, a2 = LAM1 / (LAM1 + EPS), LAM2 / (LAM2 + EPS)
    scale = (float(s) * float(s)) / (float(a.n) * float(a.m))
    results = []

    def timed(run):
        run(warmup)
        t0 = time.perf_counter()
        run(steps)
        return time.perf_counter() - t0

    if ref_lib.available():
        # (i) the reference's native inner loop, compiled unmodified: single-threaded by construction
        def run_native(k):
            ref_lib.step1(av, bv, oa, ob, K, C, dx, dy, p, q, u, v, 0, 10 ** 7, k, TAU, LAM1, LAM2, a1, a2, EPS)
        results.append(("reference", 1, "libot_ref.so step1_process_double (ot_func.cpp:690-828, 1 thread)", timed(run_native)))

    # (ii) the numpy twin of the same update (ot_solvers.py:311-316) = what wot runs in `SpaDOT analyze`; BLAS threads
    state = dict(a=np.ones(I), b=np.ones(J))

    def run_numpy(k):
        for _ in range(k):
            state["a"] = (p / (K @ (state["b"] * dy))) ** a1 * np.exp(-u / (LAM1 + EPS))
            state["b"] = (q / (K.T @ (state["a"] * dx))) ** a2 * np.exp(-v / (LAM2 + EPS))
    results.append(("reference" if not ref_lib.available() else "port", os.cpu_count() or 1,
                    "numpy K.dot / K.T.dot path (ot_solvers.py:311-316, BLAS threads)", timed(run_numpy)))
    kind, cores, what, dt = min(results, key=lambda r: r[3])
    it_per_s_sample = steps / dt
    others = "; ".join(f"{w}: {steps / t:.2f} iter/s" for _, _, w, t in results)
    return dict(value=it_per_s_sample * scale, unit=UNIT, cores=cores, kind=kind,
                sample=f"dense fp64 {s}x{s} d={a.d} (K,C resident; {build_s:.1f}s to build, untimed), fastest of [{others}] "
                       f"at sample size ({steps} iterations), rescaled by N*M ratio {scale:.3e} to the full workload"), dt / steps


def reference_main(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base, ms = cpu_reference_arm(a, a.steps, a.warmup)
    line = dict(metric=METRIC, value=base["value"], unit=UNIT, n_gpus=a.gpus, steps=a.steps, warmup=a.warmup,
                ms_per_step=ms * 1e3 / ((float(a.cpu_sample) ** 2) / (float(a.n) * float(a.m))),
                higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64", data="synthetic",
                impl="reference", config=dict(workload=workload_name(a)), cpu_baseline=base,
                e2e=dict(value=base["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
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
    from spadot_b200 import sinkhorn
    from spadot_b200.cuda_ops import CudaOps

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        td.init_process_group("nccl", device_id=dev)
    dist = sinkhorn.Dist(enabled=world > 1)

    # rows partitioned across ranks; identical RNG stream on every rank
    r0, r1 = (a.n * rank) // world, (a.n * (rank + 1)) // world
    x_all, y = synth(a.n, a.m, a.d)
    x = np.ascontiguousarray(x_all[r0:r1])
    del x_all
    # fixed normalisation: E|x-y|^2 for this mixture (the exact median is K5's job, untimed here)
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
    elapsed = torch.tensor([ev0.elapsed_time(ev1) / 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        td.all_reduce(elapsed, op=td.ReduceOp.MAX)
    elapsed = float(elapsed.item())
    launches = ops.launches - launches0
    pass_ms = [e0.elapsed_time(e1) for e0, e1 in pass_events]
    clk = clocks.stop(t_wall0, t_wall1) if clocks else None
    finite = bool(torch.isfinite(st.f).all().item() and torch.isfinite(st.g).all().item())

    # ---- e2e: host buffers in, potentials out, through the public API, copies inside the timed region
    xh = torch.from_numpy(x).pin_memory()
    yh = torch.from_numpy(y).pin_memory()
    e2e_steps = a.steps

    ops2_holder = []

    def e2e_call(n_steps):
        ops2 = CudaOps(xh, yh, device=dev, tc=a.tc)  # H2D of both spot sets + point preparation
        ops2_holder[:] = [ops2]
        ops2.set_median(median)
        st2 = sinkhorn._State(ops2, np.ones(ops2.n), dist)
        st2.u.copy_(st2.f)
        st2.v.copy_(st2.g)
        sinkhorn._sweeps(ops2, st2, dist, EPS, a1, a2, log_tau, False, n_steps)   # native loop when single-rank
        return st2.f.cpu(), (st2.g.cpu() if rank == 0 else None)      # D2H of the result

    e2e_call(1)                                      # warm-up call (allocator growth, first-use costs), untimed
    barrier()
    t0 = time.perf_counter()
    f_host, g_host = e2e_call(e2e_steps)
    barrier()
    e2e_t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        td.all_reduce(e2e_t, op=td.ReduceOp.MAX)
    e2e_t = float(e2e_t.item())
    h2d = (xh.numel() + yh.numel()) * 8
    d2h = f_host.numel() * 8 + (g_host.numel() * 8 if g_host is not None else 0)

    if rank == 0:
        import json as _json
        peaks = {}
        try:
            peaks = _json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        sm_max = float(peaks.get("sm_max_mhz", 1965.0))
        # algorithmic work of one pass launch on this rank: n_p*n_q pair evaluations,
        # each (2*dpad + 8) fp32 flop in the SIMT form and one ex2.
        n_loc = ops.n
        pairs = float(n_loc) * float(a.m)
        flop_per_pair = 2 * ops.X.dpad + 8
        avg_pass_s = (sum(pass_ms) / len(pass_ms)) / 1e3 if pass_ms else float("nan")
        n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
        traffic = None
        try:
            tj = _json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
            key = f"lse_pass_{'tc' if ops.use_tc else 'simt'} n={a.n} m={a.m} d={a.d} world={world}"
            traffic = tj.get(key, {}).get("dram_bytes_per_launch")
        except Exception:
            pass
        common = dict(traffic=traffic, traffic_source="ncu --set full, profiles/r1_traffic.json" if traffic else None,
                      algorithmic_hbm_bytes_per_launch=(n_loc + a.m) * (ops.X.dpad + 2) * 4,
                      pass_ms_avg=avg_pass_s * 1e3, pass_launches=len(pass_ms),
                      share_of_step=(sum(pass_ms) / 1e3) / elapsed,
                      hbm_gbs_algorithmic=((n_loc + a.m) * (ops.X.dpad + 2) * 4) / avg_pass_s / 1e9,
                      pairs_per_launch=pairs)
        if ops.use_tc:
            # tensor-core form: the dot product runs on tcgen05; what is left per pair is one ex2 (SFU, 16/clk/SM)
            # and ~4 fp32 instructions, so the binding pipe is the SFU.
            sfu_peak = n_sm * 16 * sm_max * 1e6 / 1e12
            achieved = pairs / avg_pass_s / 1e12
            entry = "sdb_lse_pass_tc_pred, predicted stabiliser" if ops.predicting() else "sdb_lse_pass_tc"
            roofline = dict(bound="sfu", kernel="lse_pass_tc_kernel<%d> (%s)" % (ops.X.dp, entry), achieved=achieved,
                            peak=sfu_peak, unit="Tex2/s", frac=achieved / sfu_peak,
                            peak_source=f"{n_sm} SM x 16 MUFU lanes x {sm_max:.0f} MHz (sm_max_mhz of MEASURED_PEAKS.json; "
                                        "that file has no SFU entry)",
                            frac_at_sampled_clock=(achieved / (n_sm * 16 * clk["sm_mhz"] * 1e6 / 1e12)) if clk and clk.get("sm_mhz") else None,
                            tensor_tflops=pairs * 6 * ops.X.dp / avg_pass_s / 1e12,
                            tensor_frac_of_measured_bf16=(pairs * 6 * ops.X.dp / avg_pass_s / 1e12) / float(peaks.get("bf16_tflops", 1640.6)),
                            **common)
        else:
            fp32_peak = n_sm * 128 * 2 * sm_max * 1e6 / 1e12
            achieved = pairs * flop_per_pair / avg_pass_s / 1e12
            roofline = dict(bound="fp32", kernel="pair_tile_kernel<LseEpi> (sdb_lse_pass_simt)", achieved=achieved,
                            peak=fp32_peak, unit="TFLOP/s", frac=achieved / fp32_peak,
                            peak_source=f"{n_sm} SM x 128 FMA x 2 x {sm_max:.0f} MHz (sm_max_mhz of MEASURED_PEAKS.json; "
                                        "that file has no fp32 entry)",
                            ex2_per_s=pairs / avg_pass_s, **common)
        line = dict(metric=METRIC, value=a.steps / elapsed, unit=UNIT, n_gpus=world, steps=a.steps, warmup=a.warmup,
                    ms_per_step=elapsed * 1e3 / a.steps, higher_is_better=True, scaling="strong", vs_baseline=None,
                    dtype="f16x2-split tensor tiles + f32 epilogue / f64 vectors" if ops.use_tc else "f32 tiles / f64 vectors", data="synthetic", impl="spadot_b200",
                    config=dict(workload=workload_name(a), rows_per_rank=n_loc, parallelism=f"row-partition x{world}",
                                l2_policy="inputs per pass (>=136 MB at 1M) exceed reuse; potentials rewritten every step",
                                median="analytic (K5 exact median untimed)", finite=finite,
                                collectives_per_step=(dist.collectives - coll0) / a.steps if world > 1 else 0),
                    e2e=dict(value=e2e_steps / e2e_t, unit=UNIT, h2d_bytes_per_step=h2d / e2e_steps,
                             d2h_bytes_per_step=d2h / e2e_steps, steps=e2e_steps, seconds=e2e_t,
                             api="CudaOps(host x, host y) + sinkhorn sweeps + potentials to host"),
                    gpu_launches=launches, clocks=clk, roofline=roofline)
        if world == 1 and not a.no_cpu_baseline:
            base, _ = cpu_reference_arm(a, 10, 2)
            line["cpu_baseline"] = base
        if world == 1 and not a.no_aux:
            # BASELINE.json's metric also names "train s/epoch": the train inner loop (SVGP + GAT) at ChickenHeart
            # shapes, reference formulas in plain torch vs this repo's modules, both on this GPU (tools/train_epoch_bench.py)
            try:
                sys.path.insert(0, os.path.join(ROOT, "tools"))
                import train_epoch_bench
                del ops2_holder[:]
                torch.cuda.empty_cache()
                line["aux_train_epoch"] = train_epoch_bench.run(dev)
            except Exception as exc:       # the aux number must never take the headline down
                line["aux_train_epoch"] = dict(error=repr(exc)[:200])
        print(json.dumps(line))
    if world > 1:
        td.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        reference_main(args)
    else:
        gpu_main(args)
