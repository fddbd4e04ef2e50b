"""Host drivers of the streamed unbalanced Sinkhorn (the reference's OT hot path).

What the reference does with dense N x M fp64 matrices K, _K, R on one CPU thread
(SpaDOT/utils/OT_loss/ot_solvers.py:164-531 driving ot_func.cpp:587-930) is done here in
total potentials  f = u + eps*log a,  g = v + eps*log b  (SURVEY.md §3.3):

    f_i <- eps*alpha1*( log p_i - LSE_j[(g_j - C_ij)/eps] + log J )
    g_j <- eps*alpha2*( log q_j - LSE_i[(f_i - C_ij)/eps] + log I )

Each LSE is one streamed device pass that recomputes cost tiles from the embeddings, so
nothing of size N x M exists unless the caller asks for the dense plan.  Rows (source
spots) may be partitioned across ranks: the row pass is local, the column pass produces
per-rank partial sums over M columns against a shift every rank already holds (the previous
combined column LSE), combined with ONE all-reduce(SUM) per iteration that also carries the
tau and verification flags; only the first iteration of an epsilon stage, which has no such
shift yet, combines with max-then-sum — the only data-path collectives.

The drivers are written against the `ops` interface of cuda_ops.CudaOps.
"""
from __future__ import annotations

import math
import os

import numpy as np
import torch

EPSILON_SCALINGS = 5   # ot_solvers.py:217
NEG_INF = float("-inf")


def _safe_ratio(num2, den2):
    """sqrt(num2) / (1 + sqrt(den2)) with inf/inf -> nan like the reference's fp64 arithmetic."""
    a, b = math.sqrt(num2) if num2 == num2 else float("nan"), math.sqrt(den2) if den2 == den2 else float("nan")
    if math.isinf(a) and math.isinf(b):
        return float("nan")
    return a / (1.0 + b)


class Dist:
    """Thin wrapper over torch.distributed for the row-partitioned solve (no-op when single)."""

    def __init__(self, group=None, enabled=None):
        import torch.distributed as td
        self.td = td
        self.enabled = (td.is_available() and td.is_initialized()) if enabled is None else enabled
        self.group = group
        self.world = td.get_world_size(group) if self.enabled else 1
        self.rank = td.get_rank(group) if self.enabled else 0
        self.collectives = 0

    def sum_(self, t):
        if self.world > 1:
            self.td.all_reduce(t, op=self.td.ReduceOp.SUM, group=self.group)
            self.collectives += 1
        return t

    def max_(self, t):
        if self.world > 1:
            self.td.all_reduce(t, op=self.td.ReduceOp.MAX, group=self.group)
            self.collectives += 1
        return t

    def min_(self, t):
        if self.world > 1:
            self.td.all_reduce(t, op=self.td.ReduceOp.MIN, group=self.group)
            self.collectives += 1
        return t

    def _shared_comm(self, device):
        """The ncclComm_t torch.distributed already holds for this group (ProcessGroupNCCL._comm_ptr), or None.  The library
        binds the libnccl.so.2 this process has loaded (RTLD_NOLOAD), i.e. torch's, so the handle is valid there; sharing it
        saves the second communicator's creation (0.3-1.5 s on 2-8 ranks, once per process) and its buffers.  NCCL orders
        operations issued on one communicator from different streams in host issue order, and every rank issues torch's and
        the library's collectives in the same program order."""
        import ctypes
        if os.environ.get("SDB_NATIVE_COMM", "shared") != "shared":
            return None
        try:
            pg = self.group if self.group is not None else self.td.distributed_c10d._get_default_group()
            self.td.all_reduce(torch.zeros(1, device=device), group=self.group)        # makes sure the communicator exists
            ptr = int(pg._get_backend(torch.device(device))._comm_ptr())
            return ctypes.c_void_p(ptr) if ptr else None
        except Exception:
            return None

    def native_comm(self, device):
        """NCCL communicator for the native multi-rank loop (sdb_sinkhorn_sweeps_dist issues its all-reduce from C): torch's
        own communicator of this group when it can be reached (_shared_comm), else one created by the library
        (sdb_nccl_comm_create: rank 0 draws the unique id, torch.distributed ships the 128 bytes).  None when NCCL cannot be
        bound or the group is not a CUDA/NCCL one (e.g. the gloo test groups)."""
        if self.world == 1 or getattr(device, "type", "cpu") != "cuda":
            return None
        if not hasattr(self, "_native"):
            self._native = None
            try:
                import ctypes
                from . import _lib
                lib = _lib.load()
                usable = os.environ.get("SDB_NATIVE_DIST", "1") != "0" and lib.sdb_nccl_available() and self.td.get_backend(self.group) == "nccl"
                if usable:
                    self._native = self._shared_comm(device)
                    self.native_comm_kind = "torch's (shared)" if self._native is not None else "library's own"
                if usable and self._native is None:
                    buf = (ctypes.c_char * 128)()
                    if self.rank == 0:
                        _lib.check(lib.sdb_nccl_unique_id(buf), "sdb_nccl_unique_id")
                    t = torch.frombuffer(bytearray(bytes(buf)), dtype=torch.uint8).to(device)
                    self.td.broadcast(t, src=self.td.get_global_rank(self.group, 0) if self.group is not None else 0, group=self.group)
                    idb = (ctypes.c_char * 128).from_buffer_copy(bytes(t.cpu().numpy().tobytes()))
                    comm = ctypes.c_void_p()
                    _lib.check(lib.sdb_nccl_comm_create(idb, self.world, self.rank, ctypes.byref(comm)), "sdb_nccl_comm_create")
                    self._native = comm
            except Exception as exc:           # the Python-driven loop is always available
                import warnings
                warnings.warn(f"spadot_b200: native NCCL loop unavailable ({exc!r}); using the Python-driven multi-rank loop")
                self._native = None
        return self._native

    def gather_cat(self, t):
        """Concatenate variable-length 1-D tensors from all ranks (used by the median only)."""
        if self.world == 1:
            return t
        n = torch.tensor([t.numel()], dtype=torch.int64, device=t.device)
        sizes = [torch.zeros_like(n) for _ in range(self.world)]
        self.td.all_gather(sizes, n, group=self.group)
        mx = int(max(int(s.item()) for s in sizes))
        pad = torch.zeros(mx, dtype=t.dtype, device=t.device)
        pad[:t.numel()] = t
        bufs = [torch.empty_like(pad) for _ in range(self.world)]
        self.td.all_gather(bufs, pad, group=self.group)
        self.collectives += 2
        return torch.cat([b[:int(s.item())] for b, s in zip(bufs, sizes)])


def combine_col_lse(Lc_partial, dist: Dist):
    """LSE over all ranks' rows from per-rank partial LSEs: max all-reduce, then sum all-reduce."""
    if dist.world == 1:
        return Lc_partial
    mx = dist.max_(Lc_partial.clone())
    safe = torch.where(torch.isfinite(mx), mx, torch.zeros_like(mx))
    s = dist.sum_(torch.exp(Lc_partial - safe))
    return safe + torch.log(s)


class _State:
    """Per-solve device vectors."""

    def __init__(self, ops, G_local, dist: Dist):
        begin = getattr(ops, "begin_solve", None)          # new potentials: cached bias vectors / predictions are stale
        if begin is not None:
            begin(dist)
        self.n, self.m = ops.n, ops.m
        nt = torch.tensor([float(self.n)], dtype=torch.float64, device=ops.device)
        self.N = int(round(float(dist.sum_(nt).item())))
        p = ops.tensor(G_local)
        if p.numel() != self.n:
            raise ValueError("G must have one entry per (local) source spot")
        gsum = dist.sum_(p.sum().reshape(1).clone())
        self.q_value = float(gsum.item()) / self.N               # ot_solvers.py:224  np.average(G)
        self.p = p
        self.logp = torch.log(p)
        self.logq = torch.full((self.m,), math.log(self.q_value) if self.q_value > 0 else NEG_INF,
                               dtype=torch.float64, device=ops.device)
        self.f, self.g = ops.zeros(self.n), ops.zeros(self.m)
        self.u, self.v = ops.zeros(self.n), ops.zeros(self.m)
        self.la_old, self.lb_old = ops.zeros(self.n), ops.zeros(self.m)
        self.Lr = ops.zeros(self.n)
        self.Lc = ops.zeros(self.m)


def _predicting(ops):
    """The device ops may run passes with a predicted stabiliser (cuda_ops.CudaOps._Pred): work issued since a snapshot
    must then be verified with ops.settle() and redone from the snapshot if a prediction was out of range."""
    fn = getattr(ops, "predicting", None)
    return fn is not None and fn()


def _checked(ops, dist: Dist, fn):
    """Run fn() (passes that only write their own output); redo it once with tracking if a prediction failed."""
    out = fn()
    settle = getattr(ops, "settle", None)
    if settle is not None and not settle(dist):
        out = fn()
    return out


def _sweeps(ops, st: _State, dist: Dist, eps, alpha1, alpha2, log_tau, lr_known, n_sweeps, log_floor=NEG_INF):
    """n_sweeps iterations; one native call when the ops provide it (single rank), else a Python loop."""
    native = getattr(ops, "fused_sweeps", None)
    if native is not None and dist.world == 1 and n_sweeps > 0:
        snap = ops.snapshot(st) if _predicting(ops) else None
        native(st, eps, alpha1, alpha2, log_tau, log_floor, n_sweeps, lr_known)
        if snap is not None and not ops.settle(dist):
            ops.restore(st, snap)                                 # predictions are off now: this batch tracks
            native(st, eps, alpha1, alpha2, log_tau, log_floor, n_sweeps, lr_known)
        return
    # per-iteration path (row-partitioned solve, or ops without a native loop): one snapshot / verification per batch
    snap = ops.snapshot(st) if _predicting(ops) and n_sweeps > 0 else None
    _sweeps_multi(ops, st, dist, eps, alpha1, alpha2, log_tau, lr_known, n_sweeps, log_floor)
    if snap is not None and not ops.settle(dist):
        ops.restore(st, snap)
        _sweeps_multi(ops, st, dist, eps, alpha1, alpha2, log_tau, lr_known, n_sweeps, log_floor)


def _sweeps_multi(ops, st, dist, eps, alpha1, alpha2, log_tau, lr_known, n_sweeps, log_floor):
    """Iterations of the row-partitioned solve: Python-driven while something special is due (a known row LSE to consume,
    no common column shift yet: the first iteration of an epsilon stage), then the rest of the batch in ONE native call
    (sdb_sinkhorn_sweeps_dist: kernels + the NCCL all-reduce of every iteration issued from C)."""
    native = getattr(ops, "fused_sweeps_dist", None)
    ready = getattr(ops, "col_shift_ready", None)
    comm = dist.native_comm(ops.device) if (native is not None and ready is not None and dist.world > 1) else None
    for i in range(n_sweeps):
        if comm is not None and not (lr_known and i == 0) and ready(eps):
            native(st, eps, alpha1, alpha2, log_tau, log_floor, n_sweeps - i, comm)
            dist.collectives += n_sweeps - i
            return
        _sweep_body(ops, st, dist, eps, alpha1, alpha2, log_tau, lr_known and i == 0, log_floor)


def _sweep(ops, st: _State, dist: Dist, eps, alpha1, alpha2, log_tau, lr_known, log_floor=NEG_INF):
    """One Sinkhorn iteration, verified: see _sweep_body."""
    snap = ops.snapshot(st) if _predicting(ops) else None
    _sweep_body(ops, st, dist, eps, alpha1, alpha2, log_tau, lr_known, log_floor)
    if snap is not None and not ops.settle(dist):
        ops.restore(st, snap)
        _sweep_body(ops, st, dist, eps, alpha1, alpha2, log_tau, lr_known, log_floor)


def _sweep_body(ops, st: _State, dist: Dist, eps, alpha1, alpha2, log_tau, lr_known, log_floor=NEG_INF):
    """One Sinkhorn iteration = row update + column update (ot_func.cpp:587-687) + tau bookkeeping."""
    it = ops.tick()
    fused = getattr(ops, "fused_half_step", None)
    if fused is not None:
        fused("row", st, eps, alpha1, it, log_tau, log_floor, lse_known=lr_known)
    else:
        if not lr_known:
            ops.row_lse(st.g, eps, out=st.Lr)
        ops.potential_update("row", st.Lr, st.logp, eps, alpha1, math.log(st.m), st.f, st.u, st.la_old, it, log_tau, log_floor)
    merged = False
    ready = getattr(ops, "col_shift_ready", None) if dist.world > 1 else None
    if fused is not None and dist.world == 1:
        fused("col", st, eps, alpha2, it, log_tau, log_floor)
    elif ready is not None and ready(eps):
        # every rank holds the same per-column shift (the previous combined LSE): per-rank sums are addable, ONE collective;
        # the tau flag of the row update and the verification flag of the predicted passes ride in two extra slots
        vec = ops.col_partial_sums(st, eps, it)
        dist.sum_(vec)
        ops.col_update_from_sums(vec, st, eps, alpha2, it, log_tau, log_floor)
        merged = True
    else:
        seed = getattr(ops, "seed_col_shift", None) if dist.world > 1 else None
        if _predicting(ops) and seed is None:
            ops.col_lse(st.f, eps, out=st.Lc, predict=True)       # verified by the settle() of _sweep
        else:
            ops.col_lse(st.f, eps, out=st.Lc)
        if dist.world > 1:
            st.Lc.copy_(combine_col_lse(st.Lc, dist))             # first iteration of a stage: max, then sum
        ops.potential_update("col", st.Lc, st.logq, eps, alpha2, math.log(st.N), st.g, st.v, st.lb_old, it, log_tau, log_floor)
        if seed is not None:
            seed(st, eps)
    if dist.world > 1 and not merged:
        dist.max_(ops.absorb_flag_tensor())
    ops.absorb(it, st.f, st.g, st.u, st.v)                        # ot_func.cpp:778-819


class _StageTrace:
    """The reference's `profiling` switch (config.yaml:56; ot_solvers.py:244-245,437-444): one printed line per epsilon
    stage with wall time (after a device synchronisation) and iterations, plus an NVTX range per stage for nsys/ncu."""

    def __init__(self, ops, enabled):
        self.on = bool(enabled)
        self.cuda = self.on and getattr(getattr(ops, "device", None), "type", "cpu") == "cuda"
        self.t0 = 0.0

    def _sync(self):
        if self.cuda:
            torch.cuda.synchronize()

    def begin(self, e, eps):
        if not self.on:
            return
        import time
        print("Compute for epsilon scaling {} (epsilon = {:.6g})".format(e, eps))       # ot_solvers.py:245
        if self.cuda:
            torch.cuda.nvtx.range_push(f"sinkhorn stage {e} eps={eps:.4g}")
        self._sync()
        self.t0 = time.perf_counter()

    def end(self, e, n_it, gap):
        if not self.on:
            return
        import time
        self._sync()
        if self.cuda:
            torch.cuda.nvtx.range_pop()
        print("  stage {}: {} iterations, {:.3f} ms, criterion {:.3e}".format(e, n_it, 1e3 * (time.perf_counter() - self.t0), gap))


def solve_duality_gap(ops, G_local, lambda1, lambda2, epsilon, batch_size=5, tolerance=1e-8, tau=1000.0,
                      epsilon0=1.0, max_iter=1e7, dist: Dist | None = None, info: dict | None = None, profiling=False,
                      **ignored):
    """optimal_transport_duality_gap (ot_solvers.py:164-449) on the device.

    Returns the state (potentials f,g on the device, row LSE at the final g) and the final epsilon.
    Raises RuntimeError on a NaN gap like the reference (ot_solvers.py:446-447)."""
    dist = dist or Dist(enabled=False)
    st = _State(ops, G_local, dist)
    native = getattr(ops, "fused_solve", None)
    if native is not None and dist.world == 1 and not profiling:
        # small problems: the stage loop below runs on the device in one cooperative launch (sdb_sinkhorn_solve_persistent)
        res = native(st, lambda1, lambda2, epsilon, batch_size, tolerance, tau, epsilon0, max_iter)
        if res is not None:
            if res.max_iter_reached:
                import warnings
                warnings.warn("Reached max_iter with duality gap still above threshold. Returning", RuntimeWarning, stacklevel=2)
            if res.status == 1 or math.isnan(res.gap):
                raise RuntimeError("Overflow encountered in duality gap computation, please report this incident")
            if info is not None:
                info.update(iters_per_stage=[int(v) for v in res.iters], total_iters=int(res.total_iters), gap=float(res.gap),
                            epsilon_final=float(res.eps_final), max_iter_reached=bool(res.max_iter_reached),
                            stage_criteria=[float(v) for v in res.stage_gap])
            return st, float(res.eps_final)
    scale_factor = math.exp(-math.log(epsilon) / EPSILON_SCALINGS)
    eps_i = epsilon0 * scale_factor
    log_tau = math.log(tau)
    iters, total, gap, crits = [], 0, math.inf, []
    hit_max_iter = False
    trace = _StageTrace(ops, profiling and dist.rank == 0)
    for e in range(EPSILON_SCALINGS + 1):
        st.u.copy_(st.f)                                          # absorb, ot_solvers.py:249-252
        st.v.copy_(st.g)
        eps_i = eps_i / scale_factor
        trace.begin(e, eps_i)
        alpha1 = lambda1 / (lambda1 + eps_i)
        alpha2 = lambda2 / (lambda2 + eps_i)
        final = e == EPSILON_SCALINGS
        threshold = tolerance if final else 1e-6                  # ot_solvers.py:262
        n_inner = int(batch_size) if final else 5                 # ot_func.cpp:867
        st.la_old.zero_()
        st.lb_old.zero_()
        sumK, lr_known, gap, n_it = None, False, math.inf, 0
        while gap > threshold:
            _sweeps(ops, st, dist, eps_i, alpha1, alpha2, log_tau, lr_known, n_inner)
            n_it += n_inner
            lr_known = False
            if final:
                if sumK is None:
                    Lk = ops.row_lse(None, eps_i)
                    sumK = float(dist.sum_(ops.sum_exp(Lk)).item())
                if _predicting(ops):                              # also the next iteration's row pass
                    _checked(ops, dist, lambda: ops.row_lse(st.g, eps_i, out=st.Lr, predict=True))
                else:
                    ops.row_lse(st.g, eps_i, out=st.Lr)
                lr_known = True
                t = ops.gap_terms(st.f, st.Lr, st.logp, st.g, st.Lc, st.logq, eps_i, lambda1, lambda2,
                                  1.0 / st.N, 1.0 / st.m)
                if dist.world > 1:
                    rows = dist.sum_(t[:4].clone())
                    t = torch.cat([rows, t[4:]])
                t = t.cpu().numpy()
                IJ = float(st.N) * float(st.m)
                pri = lambda1 * t[2] + lambda2 * t[6] + (t[1] + t[5] - eps_i * t[0] + eps_i * sumK) / IJ
                dua = -lambda1 * t[3] - lambda2 * t[7] - eps_i * (t[0] - sumK) / IJ
                gap = (pri - dua) / abs(pri)                      # ot_func.cpp:543
            else:
                c = ops.stage_criterion(st.f, st.u, st.la_old, st.g, st.v, st.lb_old, eps_i)
                if dist.world > 1:
                    rows = dist.sum_(c[:2].clone())
                    c = torch.cat([rows, c[2:]])
                c = c.cpu().numpy()
                va = _safe_ratio(c[0], c[1])
                vb = _safe_ratio(c[2], c[3])
                gap = vb if vb > va else va                       # std::max(v1, v2), NaN semantics included
                if math.isnan(gap):
                    break        # `while (nan > threshold)` is false in the reference as well
            if n_it >= max_iter:
                # the reference announces this and gives up on the stage (ot_solvers.py:339-341, ot_func.cpp:821-824);
                # like its native path the iteration budget is per stage (current_iter is passed as 0, ot_solvers.py:288)
                import warnings
                warnings.warn("Reached max_iter with duality gap still above threshold. Returning", RuntimeWarning, stacklevel=2)
                hit_max_iter = True
                break
        iters.append(n_it)
        total += n_it
        crits.append(float(gap))
        trace.end(e, n_it, gap)
    if math.isnan(gap):
        raise RuntimeError("Overflow encountered in duality gap computation, please report this incident")
    if not lr_known:
        ops.row_lse(st.g, eps_i, out=st.Lr)
    if info is not None:
        info.update(iters_per_stage=iters, total_iters=total, gap=float(gap), epsilon_final=eps_i, max_iter_reached=hit_max_iter,
                    stage_criteria=crits)
    return st, eps_i


def solve_stablev2(ops, G_local, lambda1, lambda2, epsilon, scaling_iter=3000, tau=1000.0, epsilon0=1.0,
                   extra_iter=1000, inner_iter_max=50, dist: Dist | None = None, info: dict | None = None, **ignored):
    """transport_stablev2 (ot_solvers.py:452-531): fixed epsilon schedule, 1e-10 floor, no host syncs."""
    dist = dist or Dist(enabled=False)
    st = _State(ops, G_local, dist)
    warm = tau is not None
    eps_i = float(epsilon0 if warm else epsilon)
    log_tau = math.log(tau) if warm else math.inf
    log_floor = math.log(1e-10)                                   # ot_solvers.py:498
    alpha1 = lambda1 / (lambda1 + eps_i)
    alpha2 = lambda2 / (lambda2 + eps_i)
    idx, done = 0, 0
    total = int(scaling_iter)
    while done < total:
        chunk = min(int(inner_iter_max), total - done) if warm else total - done
        _sweeps(ops, st, dist, eps_i, alpha1, alpha2, log_tau, False, chunk, log_floor)
        done += chunk
        if warm and chunk == inner_iter_max:                      # ot_solvers.py:513-523
            idx += 1
            st.u.copy_(st.f)
            st.v.copy_(st.g)
            eps_i = (epsilon0 - epsilon) * math.exp(-idx) + epsilon
            alpha1 = lambda1 / (lambda1 + eps_i)
            alpha2 = lambda2 / (lambda2 + eps_i)
    # the extra iterations never absorb (ot_solvers.py:525-527)
    _sweeps(ops, st, dist, eps_i, alpha1, alpha2, math.inf, False, int(extra_iter), log_floor)
    ops.row_lse(st.g, eps_i, out=st.Lr)
    if info is not None:
        info.update(total_iters=int(scaling_iter) + int(extra_iter), epsilon_final=eps_i)
    return st, eps_i


# ------------------------------------------------------------------------------------------ median (K5)
def _select_rank(ops, cand, n_cand, k):
    """k-th smallest (0-based) of cand[:n_cand] (non-negative doubles) by 8 radix-256 digit passes."""
    prefix, remaining = 0, int(k)
    for shift in range(56, -1, -8):
        hist = ops.radix_digit_hist(cand, n_cand, shift, prefix).cpu().numpy()
        csum = np.cumsum(hist)
        digit = int(np.searchsorted(csum, remaining, side="right"))
        remaining -= int(csum[digit - 1]) if digit > 0 else 0
        prefix = (prefix << 8) | digit
    return float(np.array([prefix], dtype=np.uint64).view(np.float64)[0])


def median_cost(ops, dist: Dist | None = None, n_samples=1 << 24, n_bins=4096, seed=0, small_limit=1 << 22,
                info: dict | None = None):
    """Exact np.median of all N*M squared distances (ot_solvers.py:103) without materialising them.

    1. bracket the median from a random sample of pairs (fp64 distances);
    2. one streamed sweep: count costs below the bracket and histogram the bracket uniformly;
    3. one streamed sweep: count below the selected bin and re-evaluate its members by direct fp64
       differences; 4. radix-select the two middle order statistics among those candidates.
    Sweeps classify with fp32 distances of the centred points; brackets are widened by REL_MARGIN so the
    classification can never disagree with the exact fp64 value about membership of the true median."""
    import time
    dist = dist or Dist(enabled=False)
    REL_MARGIN = 1e-5     # >= 4x the worst-case relative error of the fp32 / fp16-split distances
    n, m = ops.n, ops.m
    phases, t_last = {}, time.perf_counter()

    def mark(name):      # host wall-clock between points that synchronise anyway (reported through `info`)
        nonlocal t_last
        now = time.perf_counter()
        phases[name] = phases.get(name, 0.0) + now - t_last
        t_last = now
    N = int(round(float(dist.sum_(torch.tensor([float(n)], dtype=torch.float64, device=ops.device)).item())))
    total = N * m
    k_lo, k_hi = (total - 1) // 2, total // 2
    sweeps = 0

    def collect(lo, hi, cap):
        nonlocal sweeps
        sweeps += 1
        cand, counts = ops.cost_collect(float(lo), float(hi), int(cap))
        below = int(dist.sum_(counts[:1].clone()).item())
        n_local = int(counts[1].item())
        # every rank must take the same branch (the gather below is collective): agree on "some rank overflowed its buffer"
        over = torch.tensor([int(n_local > cap)], dtype=torch.int32, device=counts.device)
        if int(dist.max_(over).item()):
            return None, below, n_local
        mark("collect_sweep")
        allc = dist.gather_cat(cand[:n_local].contiguous())
        mark("gather")
        return allc, below, int(allc.numel())

    lo, hi = 0.0, math.inf
    expect = None
    if total > small_limit:
        ns = int(max(1024, min(n_samples, total // 16) // dist.world))
        # the sample only brackets the median (the result is exact whatever it is): draw it where the points live
        sdev = ops.device if getattr(ops.device, "type", "cpu") == "cuda" else "cpu"
        gen = torch.Generator(device=sdev).manual_seed(seed + 7919 * dist.rank)
        ii = torch.randint(0, max(n, 1), (ns,), generator=gen, device=sdev)
        jj = torch.randint(0, m, (ns,), generator=gen, device=sdev)
        samp = torch.sort(dist.gather_cat(ops.pair_distances(ii, jj)))[0] if n > 0 else torch.zeros(0, dtype=torch.float64)
        S = int(samp.numel())
        mark("sample")
        width = 6.0 * 0.5 / math.sqrt(S)
        lo = float(samp[max(0, int((0.5 - width) * S))].item()) * (1 - REL_MARGIN)
        hi = float(samp[min(S - 1, int((0.5 + width) * S))].item()) * (1 + REL_MARGIN)
        if hi > lo > 0:
            sweeps += 1
            hist, counts = ops.cost_histogram(float(np.float32(lo)), float(np.float32(hi)), n_bins)
            hist = dist.sum_(hist).cpu().numpy()
            below = int(dist.sum_(counts[:1].clone()).item())
            mark("hist_sweep")
            csum = below + np.cumsum(hist)
            if below <= k_lo and csum[-1] > k_hi:
                b_lo = int(np.searchsorted(csum, k_lo, side="right"))
                b_hi = int(np.searchsorted(csum, k_hi, side="right"))
                lo0 = float(np.float32(lo))
                w = (float(np.float32(hi)) - lo0) / n_bins
                lo, hi = (lo0 + (b_lo - 1) * w) * (1 - REL_MARGIN), (lo0 + (b_hi + 2) * w) * (1 + REL_MARGIN)
                # the histogram already says how many candidates the refined bracket holds (per rank at most that)
                first = max(int(math.floor((lo - lo0) / w)) - 1, 0)
                last = min(int(math.ceil((hi - lo0) / w)) + 1, n_bins)
                expect = int(hist[first:last].sum())
            else:
                lo, hi = 0.0, math.inf     # sample bracket missed (vanishing probability): full collect
        else:
            lo, hi = 0.0, math.inf
    if not math.isfinite(hi):
        cap = max(1 << 16, int(min(total, 1 << 26)))
    else:
        cap = max(1 << 16, int(1.25 * expect) + (1 << 16)) if expect is not None else 1 << 24
    while True:
        cand, below, n_cand = collect(max(lo, 0.0), hi if math.isfinite(hi) else 3.0e38, cap)
        if cand is not None and below <= k_lo and k_hi < below + n_cand:
            break
        if cand is None and math.isfinite(hi) and cap < (1 << 28):
            cap *= 4
            continue
        if not math.isfinite(hi) and lo <= 0.0:
            raise RuntimeError("median_cost: candidate buffer too small for an exhaustive collect")
        lo, hi = 0.0, math.inf
        cap = int(min(total, 1 << 28))
    native = getattr(ops, "select_ranks", None)
    if native is not None:                       # eight digit passes on the device, one read-back (sdb_select_ranks_f64)
        vals = native(cand, n_cand, [k_lo - below] if k_hi == k_lo else [k_lo - below, k_hi - below])
        v_lo, v_hi = vals[0], vals[-1]
    else:
        v_lo = _select_rank(ops, cand, n_cand, k_lo - below)
        v_hi = v_lo if k_hi == k_lo else _select_rank(ops, cand, n_cand, k_hi - below)
    mark("select")
    if info is not None:
        info.update(sweeps=sweeps, candidates=n_cand, below=below, phases_s=phases)
    return 0.5 * (v_lo + v_hi)
