"""Device-side operations of the streamed Sinkhorn, one thin method per C-ABI entry point.

`CudaOps` owns the device buffers of one coupling problem (the local row slice of the
source spots `x`, all target spots `y`) and launches libspadot_b200.so kernels on torch's
current CUDA stream.  torch is used for allocation, streams and collectives only.

There is no CPU implementation of this interface in the product: the drivers in
`sinkhorn.py` are written against it so that the multi-rank logic can be exercised on
CPU by the test-suite, which injects an oracle-backed stand-in (tests/_numpy_ops.py).
"""
from __future__ import annotations

import ctypes
import math
import os

import numpy as np
import torch

from . import _lib, plans

NEG_INF = float("-inf")


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _round_up(a, b):
    return (a + b - 1) // b * b


class PointSet:
    """One spot set on the device: original fp64 coordinates (row-major), centred fp32 copy
    stored feature-major for the tile kernels, and fp64 squared norms of that copy."""

    def __init__(self, x64: torch.Tensor, center: torch.Tensor):
        assert x64.dtype == torch.float64 and x64.is_cuda and x64.is_contiguous()
        self.x64 = x64
        self.n, self.d = x64.shape
        self.dpad = _round_up(self.d, 4)
        self.ld = _round_up(self.n, 64) + 64
        self.xt = torch.empty((self.dpad, self.ld), dtype=torch.float32, device=x64.device)
        self.norms_sq = torch.empty(self.n, dtype=torch.float64, device=x64.device)     # |x_i|^2 of the fp32 copy
        # the SIMT pass evaluates |x_i - y_j|^2 by direct differences: its bias / finalize kernels take zero norms
        self.norms = torch.zeros(self.n, dtype=torch.float64, device=x64.device)
        st = torch.cuda.current_stream(x64.device).cuda_stream
        _lib.call("sdb_prep_points_f64", _ptr(x64), self.n, self.d, _ptr(center), _ptr(self.xt), self.ld, self.dpad,
                  _ptr(self.norms_sq), st)
        self.x16 = None          # fp16 hi/lo split for the tensor-core pass (built on demand)

    def build_split(self, center, prescale):
        """fp16 hi/lo split [n_pad][2*dp] of (x - center) * prescale (+ fp64 norms of the represented points) for
        sdb_lse_pass_tc.  Rebuilt in place when the prescale changes (CudaOps._prep)."""
        if self.x16 is None:
            self.dp = 16 if self.d <= 16 else 32 if self.d <= 32 else 64
            self.n_pad = _round_up(max(self.n, 1), 256)
            self.x16 = torch.empty((self.n_pad, 2 * self.dp), dtype=torch.float16, device=self.x64.device)
            self.norms16 = torch.empty(self.n, dtype=torch.float64, device=self.x64.device)
        st = torch.cuda.current_stream(self.x64.device).cuda_stream
        _lib.call("sdb_prep_points_split_f16_scaled", _ptr(self.x64), self.n, self.d, _ptr(center), float(prescale),
                  _ptr(self.x16), self.n_pad, self.dp, _ptr(self.norms16), st)


class VectorOps:
    """O(N+M) fp64 device vector kernels shared by the streamed and the dense-cost solvers."""

    def _init_vectors(self, device):
        _lib.require_device()
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.launches = 0
        self._tick = 0
        self.flag = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.scratch = torch.zeros(10 * 1024 + 2, dtype=torch.float64, device=self.device)
        self.out10 = torch.zeros(10, dtype=torch.float64, device=self.device)

    def _to_dev(self, a):
        if isinstance(a, torch.Tensor):
            t = a.detach()
        else:
            t = torch.from_numpy(np.ascontiguousarray(a))
        return t.to(device=self.device, dtype=torch.float64).contiguous()

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _call(self, name, *args):
        _lib.call(name, *args, self._stream())
        self.launches += 1

    def tick(self):
        """Monotonic sweep counter; the absorb flag stores the last sweep that exceeded tau."""
        self._tick += 1
        return self._tick

    def zeros(self, n, dtype=torch.float64):
        return torch.zeros(n, dtype=dtype, device=self.device)

    def tensor(self, a, dtype=torch.float64):
        return torch.as_tensor(np.asarray(a), dtype=dtype).to(self.device)

    def _side_norms(self, side, pot):
        return pot          # only read when a bias vector is requested (streamed solver overrides)

    def _c1(self, eps):
        return 0.0

    def potential_update(self, side, L, logmarg, eps, alpha, log_n_other, pot, frame, la_old, it, log_tau,
                         log_floor=NEG_INF):
        """side='row' updates f (local rows), side='col' updates g; see sdb_potential_update."""
        self._call("sdb_potential_update", pot.numel(), _ptr(L), _ptr(logmarg), _ptr(self._side_norms(side, pot)), eps,
                   alpha, log_n_other, self._c1(eps), _ptr(pot), _ptr(frame), _ptr(la_old), self._bias_out(side, pot, eps),
                   _ptr(self.flag), it, log_tau, log_floor)

    def _bias_out(self, side, pot, eps):
        """Device pointer of the fp32 bias vector the update should refresh for the next pass (0 = none)."""
        return 0

    def absorb_flag_tensor(self):
        return self.flag

    def absorb(self, it, f, g, u, v):
        self._call("sdb_absorb", f.numel(), g.numel(), _ptr(self.flag), it, _ptr(f), _ptr(g), _ptr(u), _ptr(v))

    def stage_criterion(self, f, u, la_old, g, v, lb_old, eps):
        self._call("sdb_stage_criterion", f.numel(), g.numel(), _ptr(f), _ptr(u), _ptr(la_old), _ptr(g), _ptr(v),
                   _ptr(lb_old), eps, _ptr(self.out10), _ptr(self.scratch))
        return self.out10[:4]        # read back (synchronously) by the caller before the next reduction overwrites it

    def gap_terms(self, f, Lr, logp, g, Lc, logq, eps, lam1, lam2, dx, dy):
        self._call("sdb_gap_terms", f.numel(), g.numel(), _ptr(f), _ptr(Lr), _ptr(logp), _ptr(g), _ptr(Lc), _ptr(logq),
                   eps, lam1, lam2, dx, dy, _ptr(self.out10), _ptr(self.scratch))
        return self.out10

    def sum_exp(self, L):
        self._call("sdb_sum_exp", L.numel(), _ptr(L), 0, 0.0, _ptr(self.out10), _ptr(self.scratch))
        return self.out10[:1]

    def row_mass(self, f, Lr, eps):
        out = torch.empty_like(f)
        self._call("sdb_row_mass", f.numel(), _ptr(f), _ptr(Lr), eps, 1.0 / self.m, _ptr(out))
        return out


class DenseOps(VectorOps):
    """`ops` for a caller-supplied dense cost matrix C (N x M fp64): every sweep reads C once from HBM."""

    def __init__(self, C, device=None):
        self._init_vectors(device)
        self.C = self._to_dev(C)
        if self.C.dim() != 2:
            raise ValueError("C must be a 2-D cost matrix")
        self.n, self.m = self.C.shape
        self.n_chunks = int(max(1, min(256, (self.n + 255) // 256, max(1, (148 * 8 * 256) // max(self.m, 1)))))
        self.partial = torch.empty(2 * self.n_chunks * max(self.m, 1), dtype=torch.float64, device=self.device)

    def set_median(self, median):
        pass

    def row_lse(self, g, eps, out=None):
        out = torch.empty(self.n, dtype=torch.float64, device=self.device) if out is None else out
        self._call("sdb_dense_row_lse_f64", _ptr(self.C), self.m, self.n, self.m, _ptr(g), eps, _ptr(out))
        return out

    def col_lse(self, f, eps, out=None):
        out = torch.empty(self.m, dtype=torch.float64, device=self.device) if out is None else out
        self._call("sdb_dense_col_lse_f64", _ptr(self.C), self.m, self.n, self.m, _ptr(f), eps, _ptr(out),
                   _ptr(self.partial), self.n_chunks)
        self.launches += 1
        return out

    def plan_dense(self, f, g, eps):
        plan = torch.empty((self.n, self.m), dtype=torch.float64, device=self.device)
        for r0 in range(0, self.n, 65535):
            r1 = min(self.n, r0 + 65535)
            self._call("sdb_dense_plan_f64", _ptr(self.C[r0:r1]), self.m, r1 - r0, self.m, _ptr(f[r0:r1]), _ptr(g), eps,
                       1.0 / self.m, _ptr(plan[r0:r1]))
        return plan


class CudaOps(VectorOps):
    SIMT_MAX_D = 128
    TC_MAX_D = 64
    TC_MIN_PAIRS = 1 << 22   # below this the tensor-core pipeline cannot fill; the SIMT pass is used
    TC_ITEM_OVERHEAD_TILES = float(os.environ.get("SDB_TC_ITEM_OVERHEAD", "2.0"))   # item switch cost in tile-times (split planner)
    MAX_SPLIT_COLS = 65536   # keeps fp32 running sums < 1e-6 relative (include/spadot_b200.h)
    TARGET_CTAS = 148 * 6
    PREDICT = os.environ.get("SDB_PREDICTED_MAX", "1") != "0"   # predicted stabiliser in the tensor-core pass (see _Pred)
    PREDICT_MIN_PAIRS = 1 << 28      # below this a batch of sweeps is too short for the verification read-back to pay
    PERSISTENT_MAX_PAIRS = 1 << 24   # SIMT problems up to this many pairs run their sweeps in one cooperative launch
    PERSISTENT_CTAS_PER_SM = int(os.environ.get("SDB_PERSISTENT_CTAS_PER_SM", "2"))
    FUSED_UPDATE = os.environ.get("SDB_TC_FUSED_UPDATE", "1") != "0"   # tensor-core native loop: pass + update in one kernel
    RESIDENT_TILES = os.environ.get("SDB_RESIDENT_TILES", "1") != "0"  # one-launch solve with the cost matrix resident on chip
    # measured r2 (profiles/r2_ch_time.txt): with 1-2 tiles per CTA the resident form wins (747 x 1966: 2.48 vs 2.83 ms per solve),
    # with 4 the per-tile reductions cost more than the streamed form's dot products (1966 x 1916: 4.0 vs 3.45 ms)
    RESIDENT_MAX_TILES_PER_CTA = 2
    STRIP_FORM = os.environ.get("SDB_STRIP_FORM", "1") != "0"   # owner-computes strips in the one-launch solve when a CTA can hold its rows and columns

    def __init__(self, x_local, y, device=None, tc="auto"):
        self._init_vectors(device)
        x64 = self._to_dev(x_local)
        y64 = self._to_dev(y)
        if x64.shape[1] != y64.shape[1]:
            raise ValueError("x and y must have the same feature dimension")
        if x64.shape[1] > self.SIMT_MAX_D:
            raise ValueError(f"latent dimension {x64.shape[1]} > {self.SIMT_MAX_D} is not supported by the SIMT pass")
        self.n, self.d = x64.shape
        self.m = y64.shape[0]
        # centre on the mean of y (replicated on every rank, so no collective is needed)
        center = torch.zeros(self.d, dtype=torch.float64, device=self.device)
        self._call("sdb_column_sums_f64", _ptr(y64), self.m, self.d, _ptr(center))
        center /= max(self.m, 1)
        self.center = center
        self.X = PointSet(x64, center)
        self.Y = PointSet(y64, center)
        self.launches += 2
        self.bias_x = torch.full((_round_up(max(self.n, 1), 256),), -1.0e30, dtype=torch.float32, device=self.device)
        self.bias_y = torch.full((_round_up(max(self.m, 1), 256),), -1.0e30, dtype=torch.float32, device=self.device)
        self._bias_key = {"x": None, "y": None}     # (potential data_ptr, eps) the bias vector was last built for
        if tc not in ("auto", "on", "off"):
            raise ValueError("tc must be 'auto', 'on' or 'off'")
        self.use_tc = tc == "on" or (tc == "auto" and self.d <= self.TC_MAX_D and self.n * self.m >= self.TC_MIN_PAIRS)
        if self.use_tc:
            if self.d > self.TC_MAX_D or self.n == 0:
                raise ValueError("the tensor-core pass needs 1 <= d <= 64 and at least one row")
            amax = torch.zeros(1, dtype=torch.float64, device=self.device)
            self._call("sdb_absmax_centered_f64", _ptr(x64), self.n, self.d, _ptr(center), _ptr(amax))
            self._call("sdb_absmax_centered_f64", _ptr(y64), self.m, self.d, _ptr(center), _ptr(amax))
            self.amax = float(amax.item())
            self.n_sm = torch.cuda.get_device_properties(self.device).multi_processor_count
            self._split_key = None
            self._build_splits(None)
        self._splits = {}
        self._partials = {}
        self._direct = True
        self.inv_med = 1.0
        self.persistent = os.environ.get("SDB_PERSISTENT", "1") != "0"      # development knob (A/B against the launch loop)
        self._barrier = None
        self._pred = None
        self._pred_allowed = False          # set per solve by begin_solve()

    # ------------------------------------------------------------------ plumbing
    def set_median(self, median: float):
        self.inv_med = 1.0 / float(median)

    # Exact exponent scale.  A pass evaluates 2^(bias_j + S * (x^.y^)) with x^ = (x - centre) * q split into fp16 hi + lo.
    # q^2 * S = 2*c1*log2(e) * (1 + TC_TRUNC_COMP), c1 = 1/(median*eps), and S is a POWER OF TWO: the only fp32 constant that
    # multiplies the accumulator is then exact.  (With q a fixed power of two, S was an arbitrary real rounded to fp32: a
    # relative error of up to 2^-24 = 6e-8 common to every term.  Dominant terms have S*(x^.y^) ~ 2*c1*|x|^2*log2(e), ~30
    # at eps = 0.05 but ~150 at eps = 0.01 for BASELINE's mixture, so that one rounding alone biased every LSE by up to
    # 6e-6 — and through the fixed point every plan entry by the same relative amount.)
    # TC_TRUNC_COMP compensates the tensor core's accumulate: tcgen05.mma adds each K=16 step into its fp32 accumulator with
    # truncation toward zero, i.e. a mean error of -1/2 ulp per step at the magnitude of the running sum.  The hi.hi chain
    # (issued last, the cross chains are 2^-11 of it) has ks = dp/16 steps with running sums ~ (j/ks) of the dot product D:
    # mean error -1/2 * sum_j (j/ks) * E[ulp/x] * D = -(ks+1)/4 * 0.7213 * 2^-23 * D for log-uniform mantissas
    # (tools/tc_precision.py measures the residual bias), which the prescale gives back.  What is left is the random part,
    # +-1/2 ulp per step and pair, which averages out over the columns of a sum.
    TC_TRUNC_COMP_PER_KSTEP = float(os.environ.get("SDB_TC_TRUNC_COMP", "1.0"))

    def _trunc_comp(self):
        ks = self.X.dp // 16
        return self.TC_TRUNC_COMP_PER_KSTEP * (ks + 1) / 4.0 * 0.7213 * 2.0 ** -23

    def _build_splits(self, eps):
        """(Re)build the fp16 hi/lo splits for the exponent scale of `eps` (None: plain power-of-two prescale, used by the
        median sweeps before any eps is known).  Largest |coordinate| * q stays in [2^12.5, 2^13.5): hi + lo carry 22 bits."""
        e = math.frexp(self.amax)[1] if self.amax > 0 else 0
        q0 = 2.0 ** (13 - e)                                  # amax * q0 in [2^12, 2^13)
        if eps is None:
            key, q, S = None, 2.0 * q0, None
        else:
            T = 2.0 * (self.inv_med / eps) * math.log2(math.e) * (1.0 + self._trunc_comp())
            k = round(math.log2(T / (2.0 * q0 * q0)) - 0.5)  # S = 2^k, q = sqrt(T/S) in [q0*sqrt2, 2*q0*sqrt2)
            S = 2.0 ** k
            q = math.sqrt(T / S)
            key = (eps, self.inv_med)
        self.X.build_split(self.center, q)
        self.Y.build_split(self.center, q)
        self.launches += 2
        self.prescale, self.tc_scale, self._split_key = q, S, key
        self._nx32 = None
        self._bias_key = {"x": None, "y": None}

    # SIMT tile arithmetic (see sdb_lse_pass_simt): dot-product tiles while the largest exponent 2*c1*log2(e)*|x||y| stays at or
    # below this, direct-difference tiles (twice the inner-loop work, rounding relative to the cost of a pair) beyond it.
    # BASELINE's mixture has 2*c1*log2(e)*|x||y| = 43 at eps = 0.05, 108 at 0.02, 215 at 0.01; measured with dot-product tiles
    # (profiles/r2_tc_precision.jsonl, r2_sweep_parity_a.jsonl): max LSE error 3e-6 / 1e-5 / 4e-5, marginals 1e-6 / 2e-6 / 1.3e-5.
    # 100 keeps eps >= ~0.025 on the fast form and moves smaller regularisations to the exact one.
    SIMT_DOT_MAX = float(os.environ.get("SDB_SIMT_DOT_MAX", "100"))

    def _xy_max(self):
        if getattr(self, "_xy_max_v", None) is None:
            nx = float(self.X.norms_sq.max().item()) if self.n else 0.0
            ny = float(self.Y.norms_sq.max().item()) if self.m else 0.0
            self._xy_max_v = math.sqrt(nx * ny)
        return self._xy_max_v

    def _prep(self, eps):
        """Called by every operation that takes eps, before anything derived from the point representation (norms, bias) is
        used: rebuilds the fp16 splits for the exponent scale of `eps` (tensor-core form), or picks the SIMT tile form."""
        if self.use_tc:
            if self._split_key != (eps, self.inv_med):
                self._build_splits(eps)
        else:
            direct = 2.0 * (self.inv_med / eps) * math.log2(math.e) * self._xy_max() > self.SIMT_DOT_MAX
            if direct != self._direct:
                self._direct = direct
                self._bias_key = {"x": None, "y": None}

    def _split_plan(self, n_p, n_q):
        key = (n_p, n_q)
        if key not in self._splits:
            ns = plans.simt_splits(n_p, n_q, self.TARGET_CTAS, self.MAX_SPLIT_COLS)
            bounds = torch.tensor(plans.aligned_bounds(n_q, ns), dtype=torch.int64, device=self.device)
            self._splits[key] = (bounds, ns)
        return self._splits[key]

    def _persist_plan(self, n_p, n_q, grid):
        key = ("persist", n_p, n_q, grid)
        if key not in self._splits:
            ns = plans.persistent_splits(n_p, n_q, grid, self.MAX_SPLIT_COLS)
            bounds = torch.tensor(plans.aligned_bounds(n_q, ns), dtype=torch.int64, device=self.device)
            self._splits[key] = (bounds, ns)
        return self._splits[key]

    def _partial(self, ns, n_p, tag=None):
        key = (ns, n_p) if tag is None else (tag, ns, n_p)
        if key not in self._partials:
            self._partials[key] = torch.empty((ns, n_p, 2), dtype=torch.float32, device=self.device)
        return self._partials[key]

    # ------------------------------------------------------------------ predicted stabiliser
    # The tensor-core epilogue is issue-bound and ~16 % of its instructions track the running row maximum.  A pass over
    # rows that were swept one iteration ago can instead be told an upper bound of each row's largest exponent (the
    # previous pass's LSE in log2 units, + 1): the finalize kernel verifies that the sums stayed inside [2^-60, 2^100] and
    # raises `bad` otherwise, in which case the caller restores its snapshot of the potentials and redoes the work with the
    # tracking kernel (and predictions stay off for the rest of that solve).  Predictions are only used when the last
    # pass over the same side ran with the same (eps, median): the first iteration of every epsilon stage tracks.
    class _Pred:
        def __init__(self, ops):
            self.m = {"x": torch.zeros(max(ops.n, 1), dtype=torch.float32, device=ops.device),
                      "y": torch.zeros(max(ops.m, 1), dtype=torch.float32, device=ops.device)}
            self.bad = torch.zeros(1, dtype=torch.int32, device=ops.device)
            self.fresh = {"x": None, "y": None}
            self.ok = True          # False after a misprediction, until the next solve
            self.used = False       # a predicted pass is in flight and not yet verified
            self.common_y = None    # (eps, inv_med) for which m["y"] is the rank-independent column shift (multi-rank solve)
            self.unmerged = False   # a predicted pass ran whose `bad` flag other ranks have not seen yet

    def _pred_state(self):
        if not self._pred_allowed:
            return None
        if self._pred is None:
            self._pred = CudaOps._Pred(self)
        return self._pred if self._pred.ok else None

    def predicting(self):
        """True when sweeps may use predictions, i.e. the caller should hold a snapshot for `settle` to fall back on."""
        return self._pred_state() is not None

    def snapshot(self, st):
        vec = (st.f, st.g, st.u, st.v, st.la_old, st.lb_old, st.Lr, st.Lc)
        return [t.clone() for t in vec], self.flag.clone(), self._tick

    def restore(self, st, snap):
        vec = (st.f, st.g, st.u, st.v, st.la_old, st.lb_old, st.Lr, st.Lc)
        for t, c in zip(vec, snap[0]):
            t.copy_(c)
        self.flag.copy_(snap[1])
        self._tick = snap[2]
        self._bias_key = {"x": None, "y": None}
        if getattr(self, "_flag2", None) is not None:
            self._flag2.fill_(-1)            # snapshots are taken at batch boundaries, after the pending absorption was flushed
            self._tile_counters.zero_()

    def settle(self, dist=None):
        """Verify the predicted passes issued since the last call (one small synchronising read-back).  False means a
        prediction was out of range: the results since the caller's snapshot are invalid and predictions are now off."""
        ps = self._pred
        if ps is None or not ps.used:
            return True
        ps.used = False
        flag = ps.bad
        if dist is not None and dist.world > 1 and ps.unmerged:
            # only when a predicted pass ran since the last merged column step (whose all-reduce carries the flag): the
            # host-side `unmerged` is the same on every rank, so every rank takes this branch or none does
            flag = dist.max_(flag)
        ps.unmerged = False
        if int(flag.item()) == 0:
            return True
        ps.bad.zero_()
        ps.ok = False
        ps.fresh = {"x": None, "y": None}
        ps.common_y = None
        return False

    # ------------------------------------------------------------------ K3 passes
    def _norms(self, P: PointSet):
        if self.use_tc:
            return P.norms16
        return P.norms if self._direct else P.norms_sq          # direct-difference tiles take zero norms

    def _tc_split_plan(self, n_p, n_q):
        """(tiles per split, number of splits) of one tensor-core pass, see plans.tc_split_plan."""
        key = ("tc", n_p, n_q)
        if key not in self._splits:
            self._splits[key] = plans.tc_split_plan(n_p, n_q, self.n_sm, self.MAX_SPLIT_COLS, self.TC_ITEM_OVERHEAD_TILES)
        return self._splits[key]

    def _lse(self, P: PointSet, Q: PointSet, bias, eps, out=None, bounds=None, ns=None, finalize=True, simt=False, row_m=None,
             m_next=None, bad=None):
        c1 = self.inv_med / eps
        scale = 2.0 * c1 * math.log2(math.e)
        if self.use_tc and not simt:
            assert self._split_key == (eps, self.inv_med), "CudaOps._prep(eps) must run before a tensor-core pass"
            tps, ns = self._tc_split_plan(P.n, Q.n)
            partial = self._partial(ns, P.n)
            self._call("sdb_lse_pass_tc_pred", _ptr(P.x16), P.n, P.n_pad, _ptr(Q.x16), Q.n, Q.n_pad, P.dp, _ptr(bias),
                       self.tc_scale, tps, self.n_sm, _ptr(row_m), _ptr(partial))
            norms = P.norms16
        else:
            if bounds is None:
                bounds, ns = self._split_plan(P.n, Q.n)
            partial = self._partial(ns, P.n)
            direct = simt or self._direct                       # (simt=True: the transition table's pass, always direct)
            self._call("sdb_lse_pass_simt", _ptr(P.xt), P.ld, P.n, _ptr(Q.xt), Q.ld, Q.n, P.dpad, _ptr(bias),
                       -0.5 * scale if direct else scale, _ptr(bounds), ns, _ptr(partial))
            norms = P.norms if direct else P.norms_sq
        if not finalize:
            return partial
        if out is None:
            out = torch.empty(P.n, dtype=torch.float64, device=self.device)
        self._call("sdb_lse_finalize_pred", _ptr(partial), ns, P.n, _ptr(norms), c1, _ptr(out), _ptr(m_next), _ptr(bad))
        return out

    def _pred_args(self, side_key, eps, allow):
        """(row_m, m_next, bad) tensors for a pass over the rows of side `side_key`, and bookkeeping of freshness."""
        ps = self._pred_state() if allow else None
        if ps is None:
            return None, None, None
        key = (eps, self.inv_med)
        use = ps.fresh[side_key] == key
        ps.fresh[side_key] = key
        ps.used = ps.used or use
        ps.unmerged = ps.unmerged or use
        return (ps.m[side_key] if use else None), ps.m[side_key], ps.bad

    def row_lse(self, g, eps, out=None, predict=False):
        """Lr_i = LSE_j[(g_j - C_ij)/eps] over all columns (natural log, fp64).  g=None means g=0.
        predict=True (the solver's own potential only): may use / refresh the predicted stabiliser; the caller must
        `settle()` before trusting the result."""
        self._prep(eps)
        self._call("sdb_make_bias", self.m, self.bias_y.numel(), _ptr(g), _ptr(self._norms(self.Y)), eps,
                   self.inv_med / eps, _ptr(self.bias_y))
        self._bias_key["y"] = (_ptr(g), eps, self.inv_med) if g is not None else None
        row_m, m_next, bad = self._pred_args("x", eps, predict and g is not None)
        return self._lse(self.X, self.Y, self.bias_y, eps, out, row_m=row_m, m_next=m_next, bad=bad)

    def begin_solve(self, dist=None):
        self._bias_key = {"x": None, "y": None}
        # whether this solve may predict is decided once and identically on every rank (settle() is collective)
        ok = bool(self.use_tc and self.PREDICT and self.n * self.m >= self.PREDICT_MIN_PAIRS)
        if dist is not None and dist.world > 1:
            ok = bool(int(dist.min_(torch.tensor([int(ok)], dtype=torch.int32, device=self.device)).item()))
        self._pred_allowed = ok
        if getattr(self, "_flag2", None) is not None:
            self._flag2.fill_(-1)
        if self._pred is not None:          # new potentials: old predictions mean nothing; a new solve may predict again
            self._pred.fresh = {"x": None, "y": None}
            self._pred.ok = True
            self._pred.used = False
            self._pred.common_y = None
            self._pred.unmerged = False
            self._pred.bad.zero_()

    def _persistent_eligible(self):
        return (not self.use_tc) and self.persistent and 0 < self.n * self.m <= self.PERSISTENT_MAX_PAIRS

    STRIP_SMEM_MAX = 223 * 1024        # SDB_STRIP_SMEM_MAX

    def strip_bytes(self, n_sm):
        """Shared memory per CTA of the strip form (include/spadot_b200.h, SDB_STRIP_SMEM_MAX)."""
        g = max(1, min(n_sm, self.n, self.m))
        ldm, ldn = (self.m + 3) & ~3, (self.n + 3) & ~3
        rp, cp = -(-self.n // g), -(-self.m // g)
        return 4 * (rp * ldm + cp * ldn + max(ldm, ldn, max(rp, cp) * self.X.dpad))

    def fused_solve(self, st, lambda1, lambda2, epsilon, batch_size, tolerance, tau, epsilon0, max_iter):
        """The whole duality-gap solve in ONE cooperative launch (sdb_sinkhorn_solve_persistent) for problems the persistent
        SIMT kernel serves; None when this problem is not one of them (the caller then runs the host-driven stage loop)."""
        if not self._persistent_eligible() or os.environ.get("SDB_FUSED_SOLVE", "1") == "0":
            return None
        d = self._sweep_desc(st, float(epsilon), 0.0, 0.0, 0.0, NEG_INF)
        d.norms_x, d.norms_y = _ptr(self.X.norms_sq), _ptr(self.Y.norms_sq)
        self._persistent_plan(d)
        R, C = (self.n + 63) // 64, (self.m + 63) // 64
        # forms of the kernel, best first: 2 = strips (every CTA owns whole rows and whole columns of the cost matrix in shared
        # memory - the ChickenHeart sizes), 1 = resident 64x64 tiles, 0 = streamed tiles
        forms = []
        if self.STRIP_FORM and self.strip_bytes(self._n_sm) <= self.STRIP_SMEM_MAX:
            forms.append(2)
        if self.RESIDENT_TILES and R * C <= 2 * self._n_sm * self.RESIDENT_MAX_TILES_PER_CTA:
            forms.append(1)
        forms.append(0)

        def plan(form):
            if form == 1:       # the whole scaled cost matrix stays in shared memory: one split per tile
                d.ns_row, d.ns_col = C, R
                d.partial_row = _ptr(self._partial(C, self.n, "rrow"))
                d.partial_col = _ptr(self._partial(R, self.m, "rcol"))
            else:
                self._persistent_plan(d)
        plan(forms[0])
        # the six regularisations in the host's own arithmetic (ot_solvers.py:218,240,254), so that the device walks the very
        # same epsilons as the host-driven loop and the reference
        scale_factor = math.exp(-math.log(epsilon) / 5)
        eps_i, stages = epsilon0 * scale_factor, []
        for _ in range(6):
            eps_i = eps_i / scale_factor
            stages.append(eps_i)
        p = _lib.SolveParams(float(lambda1), float(lambda2), float(epsilon), float(epsilon0), float(tolerance), float(tau),
                             float(max_iter), (ctypes.c_double * 6)(*stages), self._xy_max(), self.SIMT_DOT_MAX, int(batch_size),
                             int(forms[0]))
        if getattr(self, "_solve_ws", None) is None:
            tiles = (self.n + 63) // 64 + (self.m + 63) // 64
            self._solve_ws = dict(flag2=torch.zeros(2, dtype=torch.int32, device=self.device),
                                  counters=torch.zeros(tiles, dtype=torch.int32, device=self.device),
                                  # per-CTA partial sums + the strip form's tagged bias vectors (n + m 64-bit words)
                                  scratch=torch.zeros(_lib.SOLVE_MAX_CTAS * 10 + self.n + self.m, dtype=torch.float64, device=self.device),
                                  result=torch.zeros(ctypes.sizeof(_lib.SolveResult), dtype=torch.uint8, device=self.device))
        ws = self._solve_ws
        fn = _lib.load().sdb_sinkhorn_solve_persistent
        args = (ctypes.byref(d), ctypes.byref(p), self._tick + 1, _ptr(ws["flag2"]), _ptr(self._barrier), _ptr(ws["counters"]),
                _ptr(ws["scratch"]), _ptr(ws["result"]), self._stream())
        status = fn(*args)
        for form in forms[1:]:
            if status != -2:
                break
            # the strips / tiles do not fit next to whatever else occupies the SMs: the next form
            plan(form)
            p.reserved = form
            status = fn(*args)
        self.solve_form = int(p.reserved)
        _lib.check(status, "sdb_sinkhorn_solve_persistent")
        self.launches += 1
        raw = ws["result"].cpu().numpy().tobytes()                # the one synchronising read-back of the solve
        res = _lib.SolveResult.from_buffer_copy(raw)
        self._tick = max(self._tick, int(res.last_tick))
        self._bias_key = {"x": None, "y": None}
        return res

    def _persistent_plan(self, d):
        """Grid and column splits of the cooperative small-problem kernels: about one work item per CTA and pass."""
        if self._barrier is None:
            self._barrier = torch.zeros(1024, dtype=torch.int32, device=self.device)      # SDB_BARRIER_WORDS
            self._n_sm = torch.cuda.get_device_properties(self.device).multi_processor_count
        d.n_ctas = self._n_sm * self.PERSISTENT_CTAS_PER_SM
        b_row, d.ns_row = self._persist_plan(self.n, self.m, d.n_ctas)
        b_col, d.ns_col = self._persist_plan(self.m, self.n, d.n_ctas)
        d.bounds_row, d.bounds_col = _ptr(b_row), _ptr(b_col)
        d.partial_row = _ptr(self._partial(d.ns_row, self.n, "prow"))
        d.partial_col = _ptr(self._partial(d.ns_col, self.m, "pcol"))

    def _sweep_desc(self, st, eps, alpha1, alpha2, log_tau, log_floor):
        """sdb_sweep_desc of this problem's SIMT form (the tensor-core fields are filled by fused_sweeps)."""
        d = _lib.SweepDesc()
        d.n, d.m, d.n_total = self.n, self.m, st.N
        d.use_tc = 0
        d.dpad, d.n_ctas = self.X.dpad, 0
        d.xt, d.ldx, d.yt, d.ldy = _ptr(self.X.xt), self.X.ld, _ptr(self.Y.xt), self.Y.ld
        d.norms_x, d.norms_y = _ptr(self.X.norms), _ptr(self.Y.norms)
        d.bias_x, d.bias_y, d.m_bias = _ptr(self.bias_x), _ptr(self.bias_y), self.bias_y.numel()
        d.f, d.g, d.u, d.v = _ptr(st.f), _ptr(st.g), _ptr(st.u), _ptr(st.v)
        d.la_old, d.lb_old, d.Lr, d.Lc = _ptr(st.la_old), _ptr(st.lb_old), _ptr(st.Lr), _ptr(st.Lc)
        d.logp, d.logq, d.flag = _ptr(st.logp), _ptr(st.logq), _ptr(self.flag)
        d.eps, d.inv_med, d.alpha1, d.alpha2 = eps, self.inv_med, alpha1, alpha2
        d.log_tau, d.log_floor, d.pow2_scale = log_tau, log_floor, 1.0
        return d

    def _full_desc(self, st, eps, alpha1, alpha2, log_tau, log_floor, fused_update=True):
        """sdb_sweep_desc of this problem in the form in use (tensor-core or SIMT), without the prediction fields."""
        d = _lib.SweepDesc()
        d.n, d.m, d.n_total = self.n, self.m, st.N
        d.use_tc = int(self.use_tc)
        d.dpad, d.n_ctas = self.X.dpad, getattr(self, "n_sm", 0)
        d.xt, d.ldx, d.yt, d.ldy = _ptr(self.X.xt), self.X.ld, _ptr(self.Y.xt), self.Y.ld
        if self.use_tc:
            d.dp = self.X.dp
            d.x16, d.n_pad, d.y16, d.m_pad = _ptr(self.X.x16), self.X.n_pad, _ptr(self.Y.x16), self.Y.n_pad
            d.tps_row, d.ns_row = self._tc_split_plan(self.n, self.m)
            d.tps_col, d.ns_col = self._tc_split_plan(self.m, self.n)
            if self.FUSED_UPDATE and fused_update:
                # two launches per iteration: the CTA that completes a 128-row tile also combines and updates it
                if getattr(self, "_flag2", None) is None:
                    self._flag2 = torch.full((2,), -1, dtype=torch.int32, device=self.device)
                    self._tile_counters = torch.zeros((self.n + 127) // 128 + (self.m + 127) // 128, dtype=torch.int32, device=self.device)
                d.flag2, d.tile_counters = _ptr(self._flag2), _ptr(self._tile_counters)
            # the native loop multiplies 2*c1*log2(e) by this factor and rounds to fp32: exactly the power of two S
            d.pow2_scale = self.tc_scale / (2.0 * (self.inv_med / eps) * math.log2(math.e))
        else:
            b_row, d.ns_row = self._split_plan(self.n, self.m)
            b_col, d.ns_col = self._split_plan(self.m, self.n)
            d.bounds_row, d.bounds_col = _ptr(b_row), _ptr(b_col)
            d.pow2_scale = 1.0
            d.simt_direct = int(self._direct)
        d.norms_x, d.norms_y = _ptr(self._norms(self.X)), _ptr(self._norms(self.Y))
        d.partial_row, d.partial_col = _ptr(self._partial(d.ns_row, self.n)), _ptr(self._partial(d.ns_col, self.m))
        if self.n == self.m and d.ns_row == d.ns_col:
            # both passes would share one (ns, n, 2) workspace from the cache: give the column pass its own
            key = ("col", d.ns_col, self.m)
            if key not in self._partials:
                self._partials[key] = torch.empty((d.ns_col, self.m, 2), dtype=torch.float32, device=self.device)
            d.partial_col = _ptr(self._partials[key])
        d.bias_x, d.bias_y, d.m_bias = _ptr(self.bias_x), _ptr(self.bias_y), self.bias_y.numel()
        d.f, d.g, d.u, d.v = _ptr(st.f), _ptr(st.g), _ptr(st.u), _ptr(st.v)
        d.la_old, d.lb_old, d.Lr, d.Lc = _ptr(st.la_old), _ptr(st.lb_old), _ptr(st.Lr), _ptr(st.Lc)
        d.logp, d.logq, d.flag = _ptr(st.logp), _ptr(st.logq), _ptr(self.flag)
        d.eps, d.inv_med, d.alpha1, d.alpha2 = eps, self.inv_med, alpha1, alpha2
        d.log_tau, d.log_floor = log_tau, log_floor
        return d

    def fused_sweeps_dist(self, st, eps, alpha1, alpha2, log_tau, log_floor, n_sweeps, comm):
        """n_sweeps iterations of the ROW-PARTITIONED solve issued by the native loop sdb_sinkhorn_sweeps_dist: kernels and the
        one NCCL all-reduce per iteration back to back on the stream, no interpreter in between.  Requires the common column
        shift (col_shift_ready) and a bias_y that matches the current g — the state every Python-driven iteration leaves."""
        self._prep(eps)
        ps = self._pred_state()
        key = (eps, self.inv_med)
        assert ps is not None and ps.common_y == key and self.use_tc
        bkey = (_ptr(st.g), eps, self.inv_med)
        if self._bias_key["y"] != bkey:
            self._call("sdb_make_bias", self.m, self.bias_y.numel(), _ptr(st.g), _ptr(self._norms(self.Y)), eps, self.inv_med / eps,
                       _ptr(self.bias_y))
        d = self._full_desc(st, eps, alpha1, alpha2, log_tau, log_floor, fused_update=False)
        d.m_x, d.m_y, d.bad_flag = _ptr(ps.m["x"]), _ptr(ps.m["y"]), _ptr(ps.bad)
        d.pred_from_row = 0 if ps.fresh["x"] == key else 1
        d.pred_from_col = 0
        ps.fresh["x"] = key
        ps.used = True
        if getattr(self, "_sum_vec", None) is None:
            self._sum_vec = torch.zeros(self.m + 2, dtype=torch.float64, device=self.device)
        first = self._tick + 1
        self._tick += n_sweeps
        _lib.call("sdb_sinkhorn_sweeps_dist", ctypes.byref(d), comm, _ptr(self._sum_vec), int(n_sweeps), first, self._stream())
        self.launches += n_sweeps * 6
        self._bias_key = {"x": (_ptr(st.f), eps, self.inv_med), "y": (_ptr(st.g), eps, self.inv_med)}
        ps.unmerged = False          # every iteration's all-reduce carried the verification flag

    def fused_sweeps(self, st, eps, alpha1, alpha2, log_tau, log_floor, n_sweeps, lr_known_first):
        """n_sweeps full iterations issued by the native loop sdb_sinkhorn_sweeps (single rank)."""
        self._prep(eps)
        d = self._full_desc(st, eps, alpha1, alpha2, log_tau, log_floor)
        ps = self._pred_state()
        if ps is not None:
            key = (eps, self.inv_med)
            d.m_x, d.m_y, d.bad_flag = _ptr(ps.m["x"]), _ptr(ps.m["y"]), _ptr(ps.bad)
            d.pred_from_row = 0 if ps.fresh["x"] == key else (2 if lr_known_first else 1)
            d.pred_from_col = 0 if ps.fresh["y"] == key else 1
            if n_sweeps > (1 if lr_known_first else 0):
                ps.fresh["x"] = key
            ps.fresh["y"] = key
            ps.used = ps.used or d.pred_from_row < n_sweeps or d.pred_from_col < n_sweeps
        first = self._tick + 1
        self._tick += n_sweeps
        if self._persistent_eligible():
            # small problem: all n_sweeps iterations in one cooperative launch (grid barriers instead of launches);
            # column splits sized so that one pass is about one work item per CTA
            self._persistent_plan(d)
            _lib.call("sdb_sinkhorn_sweeps_persistent", ctypes.byref(d), int(n_sweeps), first, int(bool(lr_known_first)),
                      _ptr(self._barrier), self._stream())
            self.launches += 1 + (0 if lr_known_first else 1)
        else:
            _lib.call("sdb_sinkhorn_sweeps", ctypes.byref(d), int(n_sweeps), first, int(bool(lr_known_first)), self._stream())
            per_iter = 2 if (self.use_tc and self.FUSED_UPDATE) else 5
            self.launches += n_sweeps * per_iter + 1 + (0 if lr_known_first else 1)
        self._bias_key = {"x": None, "y": None}

    def fused_half_step(self, side, st, eps, alpha, it, log_tau, log_floor=NEG_INF, lse_known=False):
        """One half-iteration with 2 launches: streamed pass (bias vector left behind by the previous half-step)
        + sdb_finalize_update (LSE combine, potential update, tau flag, bias for the next pass).
        side='row': g -> Lr -> f;  side='col' (single rank only): f -> Lc -> g."""
        if side == "row":
            P, Q, pot_in, pot, L, logmarg, frame, la, bias_in, bias_out, kin, kout, n_other = \
                self.X, self.Y, st.g, st.f, st.Lr, st.logp, st.u, st.la_old, self.bias_y, self.bias_x, "y", "x", self.m
        else:
            P, Q, pot_in, pot, L, logmarg, frame, la, bias_in, bias_out, kin, kout, n_other = \
                self.Y, self.X, st.f, st.g, st.Lc, st.logq, st.v, st.lb_old, self.bias_x, self.bias_y, "x", "y", st.N
        self._prep(eps)
        c1 = self.inv_med / eps
        key = (_ptr(pot_in), eps, self.inv_med)
        if lse_known:
            # L already holds the LSE at the current potentials (final-stage gap check): plain update + bias
            self.potential_update(side, L, logmarg, eps, alpha, math.log(n_other), pot, frame, la, it, log_tau, log_floor)
            return
        if self._bias_key[kin] != key:
            self._call("sdb_make_bias", Q.n, bias_in.numel(), _ptr(pot_in), _ptr(self._norms(Q)), eps, c1, _ptr(bias_in))
            self._bias_key[kin] = key
        row_m, m_next, bad = self._pred_args("x" if side == "row" else "y", eps, True)
        partial = self._lse(P, Q, bias_in, eps, finalize=False, row_m=row_m)
        ns = partial.shape[0]
        self._call("sdb_finalize_update_pred", _ptr(partial), ns, P.n, _ptr(self._norms(P)), c1, _ptr(L), _ptr(logmarg), eps, alpha,
                   math.log(n_other), _ptr(pot), _ptr(frame), _ptr(la), _ptr(bias_out), _ptr(self.flag), it, log_tau, log_floor,
                   _ptr(m_next), _ptr(bad))
        self._bias_key[kout] = (_ptr(pot), eps, self.inv_med)

    def col_lse(self, f, eps, out=None, predict=False):
        """Lc_j = LSE_{i local}[(f_i - C_ij)/eps] over this rank's rows (predict: see row_lse)."""
        if self.n == 0:
            out = torch.empty(self.m, dtype=torch.float64, device=self.device) if out is None else out
            return out.fill_(NEG_INF)
        self._prep(eps)
        self._call("sdb_make_bias", self.n, self.bias_x.numel(), _ptr(f), _ptr(self._norms(self.X)), eps,
                   self.inv_med / eps, _ptr(self.bias_x))
        self._bias_key["x"] = (_ptr(f), eps, self.inv_med) if f is not None else None
        row_m, m_next, bad = self._pred_args("y", eps, predict and f is not None)
        return self._lse(self.Y, self.X, self.bias_x, eps, out, row_m=row_m, m_next=m_next, bad=bad)

    # ------------------------------------------------------------------ row-partitioned column step with one collective
    # (driver: sinkhorn._sweep_body; kernels: sdb_partial_sums_f64 / sdb_update_from_sums_f64, see include/spadot_b200.h)
    def col_shift_ready(self, eps):
        """True when every rank holds the same per-column shift for (eps, median): the column pass can then run against it
        and the per-rank sums are combined by ONE all-reduce(SUM).  Same answer on every rank by construction."""
        ps = self._pred_state()
        return ps is not None and ps.common_y == (eps, self.inv_med)

    def seed_col_shift(self, st, eps):
        """After a column step combined the long way (max + sum all-reduces): shift_j = combined LSE_j in log2 units + 1."""
        ps = self._pred_state()
        if ps is None:
            return
        self._prep(eps)
        t = (st.Lc + self._norms(self.Y) * (self.inv_med / eps)) * math.log2(math.e) + 1.0
        ps.m["y"].copy_(torch.where(torch.isfinite(t), t, torch.zeros_like(t)))
        ps.fresh["y"] = ps.common_y = (eps, self.inv_med)

    def col_partial_sums(self, st, eps, it):
        """This rank's column sums against the common shift, + the tau and verification riders: (M + 2) doubles."""
        self._prep(eps)
        ps = self._pred
        if getattr(self, "_sum_vec", None) is None:
            self._sum_vec = torch.zeros(self.m + 2, dtype=torch.float64, device=self.device)
        vec = self._sum_vec
        if self.n == 0:
            return vec.zero_()
        key = (_ptr(st.f), eps, self.inv_med)
        if self._bias_key["x"] != key:
            self._call("sdb_make_bias", self.n, self.bias_x.numel(), _ptr(st.f), _ptr(self._norms(self.X)), eps, self.inv_med / eps,
                       _ptr(self.bias_x))
            self._bias_key["x"] = key
        partial = self._lse(self.Y, self.X, self.bias_x, eps, finalize=False, row_m=ps.m["y"])
        self._call("sdb_partial_sums_f64", _ptr(partial), partial.shape[0], self.m, _ptr(ps.m["y"]), _ptr(vec), _ptr(self.flag), it,
                   _ptr(ps.bad))
        ps.used = True
        return vec

    def col_update_from_sums(self, vec, st, eps, alpha2, it, log_tau, log_floor=NEG_INF):
        ps = self._pred
        self._call("sdb_update_from_sums_f64", _ptr(vec), _ptr(ps.m["y"]), self.m, _ptr(self._norms(self.Y)), self.inv_med / eps,
                   _ptr(st.Lc), _ptr(st.logq), eps, alpha2, math.log(st.N), _ptr(st.g), _ptr(st.v), _ptr(st.lb_old), _ptr(self.bias_y),
                   _ptr(self.flag), it, log_tau, log_floor, _ptr(ps.bad))
        self._bias_key["y"] = (_ptr(st.g), eps, self.inv_med)
        ps.used = True
        ps.unmerged = False          # the all-reduce carried every rank's flag: `bad` is now the same everywhere

    # ------------------------------------------------------------------ vector updates
    def _side_norms(self, side, pot):
        return self._norms(self.X if side == "row" else self.Y)

    def _c1(self, eps):
        return self.inv_med / eps

    def potential_update(self, side, L, logmarg, eps, *args, **kw):
        self._prep(eps)
        return super().potential_update(side, L, logmarg, eps, *args, **kw)

    def _bias_out(self, side, pot, eps):
        # the potential changes in place: the cached bias vector of that side must follow it (the cache is keyed by
        # pointer, so a stale entry would otherwise be reused by the next pass)
        k, bias = ("x", self.bias_x) if side == "row" else ("y", self.bias_y)
        self._bias_key[k] = (_ptr(pot), eps, self.inv_med)
        return _ptr(bias)

    def plan_dense(self, f, g, eps):
        plan = torch.empty((self.n, self.m), dtype=torch.float64, device=self.device)
        self._call("sdb_plan_dense_f64", _ptr(self.X.x64), _ptr(self.Y.x64), self.n, self.m, self.d, _ptr(f), _ptr(g),
                   self.inv_med, eps, 1.0 / self.m, _ptr(plan))
        return plan

    # ------------------------------------------------------------------ K5 median primitives
    def pair_distances(self, ii, jj):
        ii = torch.as_tensor(ii, dtype=torch.int64).to(self.device)
        jj = torch.as_tensor(jj, dtype=torch.int64).to(self.device)
        out = torch.empty(ii.numel(), dtype=torch.float64, device=self.device)
        self._call("sdb_pair_distances_f64", _ptr(self.X.x64), _ptr(self.Y.x64), self.d, _ptr(ii), _ptr(jj), ii.numel(),
                   _ptr(out))
        return out

    def _sweep_norms(self):
        """fp32 |x_i|^2 and padded |y_j|^2 (+3e38 padding) of the fp16-split points, for the tensor-core sweeps."""
        if getattr(self, "_nx32", None) is None:
            self._nx32 = self.X.norms16.to(torch.float32).contiguous()
            self._ny32 = torch.full((self.Y.n_pad,), 3.0e38, dtype=torch.float32, device=self.device)
            self._ny32[:self.m] = self.Y.norms16.to(torch.float32)
        return self._nx32, self._ny32

    def cost_histogram(self, lo, hi, n_bins):
        hist = torch.zeros(n_bins, dtype=torch.int64, device=self.device)
        counts = torch.zeros(2, dtype=torch.int64, device=self.device)
        if self.use_tc and n_bins <= 4096:
            nx, ny = self._sweep_norms()
            tps, _ = self._tc_split_plan(self.n, self.m)
            self._call("sdb_cost_histogram_tc", _ptr(self.X.x16), self.n, self.X.n_pad, _ptr(self.Y.x16), self.m, self.Y.n_pad,
                       self.X.dp, _ptr(nx), _ptr(ny), -2.0 / (self.prescale * self.prescale), tps, self.n_sm, lo, hi, n_bins,
                       _ptr(hist), _ptr(counts))
            return hist, counts
        bounds, ns = self._split_plan(self.n, self.m)
        self._call("sdb_cost_histogram", _ptr(self.X.xt), self.X.ld, self.n, _ptr(self.Y.xt), self.Y.ld, self.m,
                   self.X.dpad, _ptr(bounds), ns, lo, hi, n_bins, _ptr(hist), _ptr(counts))
        return hist, counts

    def cost_collect(self, lo, hi, cap):
        cand = torch.empty(cap, dtype=torch.float64, device=self.device)
        counts = torch.zeros(2, dtype=torch.int64, device=self.device)
        if self.use_tc:
            nx, ny = self._sweep_norms()
            tps, _ = self._tc_split_plan(self.n, self.m)
            self._call("sdb_cost_collect_tc", _ptr(self.X.x16), self.n, self.X.n_pad, _ptr(self.Y.x16), self.m, self.Y.n_pad,
                       self.X.dp, _ptr(nx), _ptr(ny), -2.0 / (self.prescale * self.prescale), tps, self.n_sm, lo, hi,
                       _ptr(self.X.x64), _ptr(self.Y.x64), self.d, _ptr(cand), cap, _ptr(counts))
            return cand, counts
        bounds, ns = self._split_plan(self.n, self.m)
        self._call("sdb_cost_collect", _ptr(self.X.xt), self.X.ld, self.n, _ptr(self.Y.xt), self.Y.ld, self.m,
                   self.X.dpad, _ptr(bounds), ns, lo, hi, _ptr(self.X.x64), _ptr(self.Y.x64), self.d, _ptr(cand), cap,
                   _ptr(counts))
        return cand, counts

    def select_ranks(self, cand, n, ranks):
        """The ranks[s]-th smallest (0-based, one or two ranks) of cand[:n]: all eight digit passes on the device, one read-back."""
        import ctypes
        ranks = [int(r) for r in ranks]
        arr = (ctypes.c_uint64 * len(ranks))(*ranks)
        out = torch.empty(2, dtype=torch.float64, device=self.device)
        work = torch.empty(4128 // 8, dtype=torch.int64, device=self.device)       # SDB_SELECT_WORKSPACE_BYTES
        self._call("sdb_select_ranks_f64", _ptr(cand), int(n), ctypes.addressof(arr), len(ranks), _ptr(out), _ptr(work))
        return [float(v) for v in out[:len(ranks)].tolist()]

    def radix_digit_hist(self, cand, n, shift, prefix):
        hist = torch.zeros(256, dtype=torch.int64, device=self.device)
        self._call("sdb_radix_digit_hist", _ptr(cand), n, shift, prefix, _ptr(hist))
        return hist

    # ------------------------------------------------------------------ K6 transition table
    def _transition_table_tc(self, f, g, eps, lx, ly, k0, k1):
        """K6 on the tensor-core pass: target spots permuted by domain label, every label group padded to whole 256-column
        tiles, ONE pass with one variable-length split per label (sdb_lse_pass_tc_groups) -> (max, sum) per (label, row),
        then the rows are folded by their own labels.  Same point representation (fp16 hi/lo split, norms) as the solve."""
        self._prep(eps)
        valid = torch.nonzero((ly >= 0) & (ly < k1)).flatten()         # labels outside [0, k1) take no part (as in the SIMT form)
        lyv = ly[valid]
        counts = torch.bincount(lyv, minlength=k1)[:k1]
        counts_h = counts.cpu().numpy()
        # one split per label, cut further so that no split exceeds MAX_SPLIT_COLS columns (fp32 running sums)
        max_tiles = self.MAX_SPLIT_COLS // 256
        split_label, split_tiles, group_start, t = [], [0], np.zeros(k1, dtype=np.int64), 0
        for lab in range(k1):
            nt = int(-(-int(counts_h[lab]) // 256))
            group_start[lab] = t * 256
            while nt > 0:
                step = min(nt, max_tiles)
                t += step
                nt -= step
                split_label.append(lab)
                split_tiles.append(t)
        table = torch.zeros((k0, k1), dtype=torch.float64, device=self.device)
        if not split_label:
            return table
        m_pad = t * 256
        order = torch.argsort(lyv, stable=True)
        perm = valid[order]                                             # original column of each permuted column
        lp = lyv[order]
        first = torch.zeros(k1 + 1, dtype=torch.int64, device=self.device)
        first[1:] = torch.cumsum(counts, 0)
        gs = torch.from_numpy(group_start).to(self.device)
        pos = gs[lp] + (torch.arange(perm.numel(), device=self.device) - first[lp])        # padded slot of permuted column
        y_pad = torch.zeros((m_pad, self.d), dtype=torch.float64, device=self.device)
        y_pad[pos] = self.Y.x64[perm]
        g_pad = torch.full((m_pad,), NEG_INF, dtype=torch.float64, device=self.device)     # padding: bias = sentinel
        g_pad[pos] = g[perm]
        y16 = torch.empty((m_pad, 2 * self.Y.dp), dtype=torch.float16, device=self.device)
        norms = torch.empty(m_pad, dtype=torch.float64, device=self.device)
        self._call("sdb_prep_points_split_f16_scaled", _ptr(y_pad), m_pad, self.d, _ptr(self.center), float(self.prescale), _ptr(y16),
                   m_pad, self.Y.dp, _ptr(norms))
        bias = torch.empty(m_pad, dtype=torch.float32, device=self.device)
        c1 = self.inv_med / eps
        self._call("sdb_make_bias", m_pad, m_pad, _ptr(g_pad), _ptr(norms), eps, c1, _ptr(bias))
        kk = len(split_label)
        st_dev = torch.tensor(split_tiles, dtype=torch.int32, device=self.device)
        sl_dev = torch.tensor(split_label, dtype=torch.int64, device=self.device)
        chunk = max(1, 4096 // k0)                                      # sdb_transition_accumulate holds k0 x splits in shared memory
        partial = torch.empty((kk, self.n, 2), dtype=torch.float32, device=self.device)
        self._call("sdb_lse_pass_tc_groups", _ptr(self.X.x16), self.n, self.X.n_pad, _ptr(y16), m_pad, self.X.dp, _ptr(bias),
                   self.tc_scale, _ptr(st_dev), kk, self.n_sm, _ptr(partial))
        for s0 in range(0, kk, chunk):
            s1 = min(kk, s0 + chunk)
            sub = torch.zeros((k0, s1 - s0), dtype=torch.float64, device=self.device)
            self._call("sdb_transition_accumulate", _ptr(partial[s0:s1]), s1 - s0, self.n, _ptr(self.X.norms16), c1, _ptr(f), eps,
                       1.0 / self.m, _ptr(lx), k0, _ptr(sub))
            table.index_add_(1, sl_dev[s0:s1], sub)
        return table

    def transition_table(self, f, g, eps, labels_x, labels_y, k0, k1):
        """table[a,b] = sum_{i in a, j in b} exp((f_i+g_j-C_ij)/eps)/M for this rank's rows."""
        ly = torch.as_tensor(np.asarray(labels_y), dtype=torch.int64).to(self.device)
        lx = torch.as_tensor(np.asarray(labels_x), dtype=torch.int32).to(self.device)
        if self.use_tc and self.n > 0 and k0 <= 4096:
            return self._transition_table_tc(f, g, eps, lx, ly, k0, k1)
        perm = torch.argsort(ly, stable=True)
        counts = torch.bincount(ly, minlength=k1)[:k1]
        bounds = torch.zeros(k1 + 1, dtype=torch.int64, device=self.device)
        bounds[1:] = torch.cumsum(counts, 0)
        Yp = PointSet.__new__(PointSet)
        Yp.n, Yp.d, Yp.dpad, Yp.ld = self.Y.n, self.Y.d, self.Y.dpad, self.Y.ld
        Yp.xt = torch.zeros_like(self.Y.xt)
        Yp.xt[:, :self.m] = self.Y.xt[:, :self.m].index_select(1, perm)
        Yp.norms = self.Y.norms.index_select(0, perm)
        gp = g.index_select(0, perm).contiguous()
        bias = torch.empty(self.m, dtype=torch.float32, device=self.device)
        c1 = self.inv_med / eps
        self._call("sdb_make_bias", self.m, self.m, _ptr(gp), _ptr(Yp.norms), eps, c1, _ptr(bias))
        partial = self._lse(self.X, Yp, bias, eps, bounds=bounds, ns=k1, finalize=False, simt=True)
        table = torch.zeros((k0, k1), dtype=torch.float64, device=self.device)
        self._call("sdb_transition_accumulate", _ptr(partial), k1, self.n, _ptr(self.X.norms), c1, _ptr(f), eps,
                   1.0 / self.m, _ptr(lx), k0, _ptr(table))
        return table
