// O(N+M) fp64 vector kernels around the streamed Sinkhorn passes: point preparation, partial-LSE
// combination, potential updates, absorption bookkeeping, stopping rules (K4), plan / mass
// evaluation, radix-select helper (K5) and transition-table accumulation (K6).
// ref: SpaDOT/utils/OT_loss/ot_func.cpp:358-584,587-930; SpaDOT/utils/OT_loss/ot_solvers.py:102-158,449.
#include "sdb_common.cuh"

namespace {

// ------------------------------------------------------------------------------------ preparation
__global__ void column_sums_kernel(const double* __restrict__ x, int64_t n, int d, double* __restrict__ out) {
    // one block handles a contiguous slab of rows; thread k-strided accumulation keeps loads coalesced
    extern __shared__ double sh[];  // [blockDim.x]
    const int64_t rows_per_block = (n + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    const int64_t r1 = min(n, r0 + rows_per_block);
    const int64_t total = (r1 > r0) ? (r1 - r0) * d : 0;
    // each thread owns the feature (threadIdx.x % d) when blockDim.x % d == 0; general case: loop per feature
    for (int k = 0; k < d; ++k) {
        double acc = 0.0;
        for (int64_t r = r0 + threadIdx.x; r < r1; r += blockDim.x) acc += x[r * d + k];
        sh[threadIdx.x] = acc;
        __syncthreads();
        for (int s = blockDim.x / 2; s > 0; s >>= 1) {
            if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
            __syncthreads();
        }
        if (threadIdx.x == 0 && total > 0) atomicAdd(out + k, sh[0]);
        __syncthreads();
    }
}

__global__ void prep_points_kernel(const double* __restrict__ x, int64_t n, int d, const double* __restrict__ center,
                                   float* __restrict__ xt, int64_t ld, int dpad, double* __restrict__ norms) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ld) return;
    double nn = 0.0;
    for (int k = 0; k < dpad; ++k) {
        float v = 0.f;
        if (i < n && k < d) v = (float)(x[i * d + k] - center[k]);
        xt[(int64_t)k * ld + i] = v;   // coalesced across i
        nn += (double)v * (double)v;
    }
    if (i < n) norms[i] = nn;
}

// ------------------------------------------------------------------------------------ LSE combine + updates
__global__ void lse_finalize_kernel(const float2* __restrict__ partial, int n_splits, int64_t n,
                                    const double* __restrict__ norms, double c1, double* __restrict__ L, float* m_next, int* bad_flag) {
    sdb_launch_dependents();
    sdb_grid_dependency_wait();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    L[i] = sdb_combine_partials(partial, n_splits, n, i, norms[i] * c1, m_next, bad_flag);
}

__global__ void potential_update_kernel(int64_t n, const double* __restrict__ L, const double* __restrict__ logmarg,
                                        const double* __restrict__ norms, double eps, double alpha, double log_n_other,
                                        double c1, double* __restrict__ pot, const double* __restrict__ frame,
                                        double* __restrict__ la_old, float* __restrict__ bias, int* __restrict__ absorb_flag,
                                        int iter, double log_tau, double log_floor) {
    sdb_launch_dependents();
    sdb_grid_dependency_wait();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    sdb_update_row(i, L[i], logmarg[i], norms[i], eps, alpha, log_n_other, c1, pot, frame, la_old, bias, absorb_flag, iter, log_tau,
                   log_floor);
}

// finalize + potential update + next-pass bias in one launch (per-iteration launch count 9 -> 5)
__global__ void finalize_update_kernel(const float2* __restrict__ partial, int n_splits, int64_t n, const double* __restrict__ norms,
                                       double c1, double* __restrict__ L, const double* __restrict__ logmarg, double eps,
                                       double alpha, double log_n_other, double* __restrict__ pot, const double* __restrict__ frame,
                                       double* __restrict__ la_old, float* __restrict__ bias, int* __restrict__ absorb_flag, int iter,
                                       double log_tau, double log_floor, float* m_next, int* bad_flag) {
    sdb_launch_dependents();
    sdb_grid_dependency_wait();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double Li = sdb_combine_partials(partial, n_splits, n, i, norms[i] * c1, m_next, bad_flag);
    L[i] = Li;
    sdb_update_row(i, Li, logmarg[i], norms[i], eps, alpha, log_n_other, c1, pot, frame, la_old, bias, absorb_flag, iter, log_tau,
                   log_floor);
}

// ---- row-partitioned solve: the column step with ONE collective ("a shift s_j known identically on all ranks")
// Every rank sweeps its rows against the same per-column shift (the previous combined column LSE in log2 units, + 1: the
// predicted stabiliser), so the per-rank sums are directly addable: sums[j] = sum_s partial_s(j).sum * 2^(partial_s(j).max - shift[j]).
// Two more slots ride in the same all-reduce(SUM): sums[n] = 1 if this rank's row update exceeded tau in iteration `iter`
// (ref: ot_func.cpp:778-790), sums[n+1] = 1 if one of this rank's predicted passes was out of range.
__global__ void partial_sums_kernel(const float2* __restrict__ partial, int n_splits, int64_t n, const float* __restrict__ shift,
                                    double* __restrict__ sums, const int* __restrict__ absorb_flag, int iter,
                                    const int* __restrict__ bad_flag) {
    sdb_launch_dependents();
    sdb_grid_dependency_wait();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        sums[n] = (absorb_flag && *absorb_flag == iter) ? 1.0 : 0.0;
        sums[n + 1] = (bad_flag && *bad_flag != 0) ? 1.0 : 0.0;
    }
    if (i >= n) return;
    const double sh = (double)shift[i];
    double S = 0.0;
    for (int s = 0; s < n_splits; ++s) {
        const float2 ps = __ldcg(partial + (int64_t)s * n + i);
        if (ps.x > -1e29f && ps.y > 0.f) S += (double)ps.y * exp2((double)ps.x - sh);
    }
    sums[i] = S;
}

// After the all-reduce: LSE from the combined sums, potential update, next-pass bias, next shift, verification — the same on
// every rank, because every rank holds the same sums.
__global__ void update_from_sums_kernel(const double* __restrict__ sums, float* shift, int64_t n, const double* __restrict__ norms,
                                        double c1, double* __restrict__ L, const double* __restrict__ logmarg, double eps, double alpha,
                                        double log_n_other, double* __restrict__ pot, const double* __restrict__ frame,
                                        double* __restrict__ la_old, float* __restrict__ bias, int* __restrict__ absorb_flag, int iter,
                                        double log_tau, double log_floor, int* bad_flag) {
    sdb_launch_dependents();
    sdb_grid_dependency_wait();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        if (sums[n] > 0.0) atomicMax(absorb_flag, iter);
        if (bad_flag && sums[n + 1] > 0.0) atomicOr(bad_flag, 1);
    }
    if (i >= n) return;
    const double S = sums[i], M = (double)shift[i];
    double Li = -INFINITY;
    if (S > 0.0 && S < INFINITY) {
        Li = SDB_LN2 * (M + log2(S)) - norms[i] * c1;
        sdb_prediction_out(M, S, i, shift, bad_flag);             // shift[i] <- M + log2 S + 1; range check of the combined sum
    } else if (bad_flag) {
        atomicOr(bad_flag, 1);                                    // every term flushed, or an overflow: redo with tracking
    }
    L[i] = Li;
    sdb_update_row(i, Li, logmarg[i], norms[i], eps, alpha, log_n_other, c1, pot, frame, la_old, bias, absorb_flag, iter, log_tau,
                   log_floor);
}

__global__ void potential_update_deferred_kernel(int64_t n, const double* __restrict__ L, const double* __restrict__ logmarg,
                                                 const double* __restrict__ norms, double eps, double alpha, double log_n_other,
                                                 double c1, double* __restrict__ pot, double* __restrict__ frame,
                                                 double* __restrict__ la_old, float* __restrict__ bias, int* flag2, int iter,
                                                 double log_tau, double log_floor) {
    sdb_launch_dependents();
    sdb_grid_dependency_wait();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    sdb_update_row_deferred(i, L[i], logmarg[i], norms[i], eps, alpha, log_n_other, c1, pot, frame, la_old, bias, flag2, iter, log_tau,
                            log_floor);
}

__global__ void absorb_pending_kernel(int64_t n, int64_t m, const int* __restrict__ flag2, int last_tick, const double* __restrict__ f,
                                      const double* __restrict__ g, double* __restrict__ u, double* __restrict__ v, int* flag) {
    sdb_launch_dependents();
    sdb_grid_dependency_wait();
    if (flag2[last_tick & 1] != last_tick) return;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 && flag) *flag = last_tick;
    if (i < n) u[i] = f[i];
    if (i < m) v[i] = g[i];
}

__global__ void make_bias_kernel(int64_t n, int64_t n_pad, const double* __restrict__ pot, const double* __restrict__ norms,
                                 double eps, double c1, float* __restrict__ bias) {
    sdb_launch_dependents();
    sdb_grid_dependency_wait();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    if (i >= n) { bias[i] = SDB_NEG_SENTINEL; return; }
    const double p = pot ? pot[i] : 0.0;
    double b = SDB_LOG2E * (p / eps - norms[i] * c1);
    bias[i] = (b > (double)SDB_NEG_SENTINEL) ? (float)b : SDB_NEG_SENTINEL;
}

__global__ void absorb_kernel(int64_t n, int64_t m, const int* __restrict__ flag, int iter, const double* __restrict__ f,
                              const double* __restrict__ g, double* __restrict__ u, double* __restrict__ v) {
    sdb_launch_dependents();
    sdb_grid_dependency_wait();
    if (*flag != iter) return;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) u[i] = f[i];
    if (i < m) v[i] = g[i];
}

// ------------------------------------------------------------------------------------ stopping rules
__global__ void __launch_bounds__(256) stage_criterion_kernel(int64_t n, int64_t m, const double* __restrict__ f,
                                                              const double* __restrict__ u, const double* __restrict__ la_old,
                                                              const double* __restrict__ g, const double* __restrict__ v,
                                                              const double* __restrict__ lb_old, double eps, double* out4,
                                                              void* scratch) {
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double eu = exp(u[i] / eps);
        const double at = exp((f[i] - u[i]) / eps) * eu;     // _a = a * exp(u/eps)      ot_func.cpp:878-880
        const double df = at - exp(la_old[i]) * eu;          // _a - old_a*exp(u/eps)    ot_func.cpp:900
        acc[0] += df * df;
        acc[1] += at * at;
    }
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += stride) {
        const double ev = exp(v[j] / eps);
        const double bt = exp((g[j] - v[j]) / eps) * ev;
        const double df = bt - exp(lb_old[j]) * ev;
        acc[2] += df * df;
        acc[3] += bt * bt;
    }
    sdb_grid_reduce<4>(acc, scratch, out4);
}

__global__ void __launch_bounds__(256) gap_terms_kernel(int64_t n, int64_t m, const double* __restrict__ f,
                                                        const double* __restrict__ Lr, const double* __restrict__ logp,
                                                        const double* __restrict__ g, const double* __restrict__ Lc,
                                                        const double* __restrict__ logq, double eps, double lam1, double lam2,
                                                        double dx, double dy, double* out10, void* scratch) {
    double acc[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) acc[k] = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double R = exp(f[i] / eps + Lr[i]);
        const double p = exp(logp[i]);
        const double r = R * dy;
        acc[0] += R;
        acc[1] += f[i] * R;
        acc[2] += dx * (r * (log(r) - logp[i]) - r + p);      // fdiv, ot_func.cpp:309-322
        acc[3] += p * dx * (exp(-f[i] / lam1) - 1.0);         // fdivstarexp, ot_func.cpp:341-355
    }
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += stride) {
        const double R = exp(g[j] / eps + Lc[j]);
        const double q = exp(logq[j]);
        const double c = R * dx;
        acc[4] += R;
        acc[5] += g[j] * R;
        acc[6] += dy * (c * (log(c) - logq[j]) - c + q);
        acc[7] += q * dy * (exp(-g[j] / lam2) - 1.0);
    }
    sdb_grid_reduce<10>(acc, scratch, out10);
}

__global__ void __launch_bounds__(256) sum_exp_kernel(int64_t n, const double* __restrict__ L, const double* __restrict__ add,
                                                      double add_scale, double* out1, void* scratch) {
    double acc[1] = {0.0};
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        acc[0] += exp(L[i] + (add ? add[i] * add_scale : 0.0));
    sdb_grid_reduce<1>(acc, scratch, out1);
}

// ------------------------------------------------------------------------------------ plan / mass
__global__ void row_mass_kernel(int64_t n, const double* __restrict__ f, const double* __restrict__ Lr, double eps,
                                double inv_m, double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = exp(f[i] / eps + Lr[i]) * inv_m;
}

// 32 x 32 output tile per block of 256 threads (each thread 4 rows); y slab staged in shared memory.
__global__ void __launch_bounds__(256) plan_dense_kernel(const double* __restrict__ x, const double* __restrict__ y,
                                                         int64_t n, int64_t m, int d, const double* __restrict__ f,
                                                         const double* __restrict__ g, double inv_med, double eps,
                                                         double inv_m, double* __restrict__ plan) {
    extern __shared__ double shd[];      // xs[32][d+1], ys[32][d+1]
    double* xs = shd;
    double* ys = shd + 32 * (d + 1);
    const int64_t i0 = (int64_t)blockIdx.y * 32, j0 = (int64_t)blockIdx.x * 32;
    for (int idx = threadIdx.x; idx < 32 * d; idx += 256) {
        const int r = idx / d, k = idx % d;
        xs[r * (d + 1) + k] = (i0 + r < n) ? x[(i0 + r) * d + k] : 0.0;
        ys[r * (d + 1) + k] = (j0 + r < m) ? y[(j0 + r) * d + k] : 0.0;
    }
    __syncthreads();
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // ty in 0..7, rows ty, ty+8, ty+16, ty+24
    const int64_t j = j0 + tx;
    const double gj = (j < m) ? g[j] : 0.0;
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
        const int r = ty + rr * 8;
        const int64_t i = i0 + r;
        if (i >= n || j >= m) continue;
        double s = 0.0;
        for (int k = 0; k < d; ++k) {   // no FMA contraction: bit-identical to scipy's cdist loop
            const double df = xs[r * (d + 1) + k] - ys[tx * (d + 1) + k];
            s = __dadd_rn(s, __dmul_rn(df, df));
        }
        plan[i * m + j] = exp((f[i] + gj - s * inv_med) / eps) * inv_m;
    }
}

// ------------------------------------------------------------------------------------ median helpers
__global__ void pair_distances_kernel(const double* __restrict__ x, const double* __restrict__ y, int d,
                                      const int64_t* __restrict__ ii, const int64_t* __restrict__ jj, int64_t n_pairs,
                                      double* __restrict__ dist) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_pairs) return;
    const double* xi = x + ii[s] * d;
    const double* yj = y + jj[s] * d;
    double acc = 0.0;
    for (int k = 0; k < d; ++k) { const double df = xi[k] - yj[k]; acc = __dadd_rn(acc, __dmul_rn(df, df)); }
    dist[s] = acc;
}

__global__ void radix_digit_hist_kernel(const double* __restrict__ cand, unsigned long long n, int shift,
                                        unsigned long long prefix, unsigned long long* __restrict__ hist256) {
    __shared__ unsigned int sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned long long key = (unsigned long long)__double_as_longlong(cand[i]);
        const bool match = (shift >= 56) ? true : ((key >> (shift + 8)) == prefix);
        if (match) atomicAdd(&sh[(key >> shift) & 255ull], 1u);
    }
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(hist256 + threadIdx.x, (unsigned long long)sh[threadIdx.x]);
}

// Radix select of up to two order statistics without host round trips: eight (histogram, digit) kernel pairs on one stream.
// state: prefix[2], remaining[2], hist[2][256] (device).  Same digit rule as the host loop it replaces
// (sinkhorn._select_rank: digit = #{b : cumsum(hist)[b] <= remaining}).
struct SelectState {
    unsigned long long prefix[2];
    unsigned long long remaining[2];
    unsigned long long hist[2][256];
};

__global__ void select_init_kernel(SelectState* st, unsigned long long k0, unsigned long long k1) {
    const int t = threadIdx.x;
    st->hist[0][t] = 0ull;
    st->hist[1][t] = 0ull;
    if (t == 0) { st->prefix[0] = st->prefix[1] = 0ull; st->remaining[0] = k0; st->remaining[1] = k1; }
}

__global__ void select_hist_kernel(const double* __restrict__ cand, unsigned long long n, int shift, int n_ranks,
                                   SelectState* __restrict__ st) {
    __shared__ unsigned int sh[2][256];
    sh[0][threadIdx.x] = 0;
    sh[1][threadIdx.x] = 0;
    __syncthreads();
    const unsigned long long p0 = st->prefix[0], p1 = st->prefix[1];
    const bool same = n_ranks < 2 || p0 == p1;          // the two ranks usually share every digit but the last ones
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned long long key = (unsigned long long)__double_as_longlong(cand[i]);
        const unsigned long long hi = (shift >= 56) ? 0ull : (key >> (shift + 8));
        const unsigned int b = (unsigned int)((key >> shift) & 255ull);
        if (hi == p0) atomicAdd(&sh[0][b], 1u);
        if (!same && hi == p1) atomicAdd(&sh[1][b], 1u);
    }
    __syncthreads();
    if (sh[0][threadIdx.x]) {
        atomicAdd(&st->hist[0][threadIdx.x], (unsigned long long)sh[0][threadIdx.x]);
        if (same && n_ranks > 1) atomicAdd(&st->hist[1][threadIdx.x], (unsigned long long)sh[0][threadIdx.x]);
    }
    if (!same && sh[1][threadIdx.x]) atomicAdd(&st->hist[1][threadIdx.x], (unsigned long long)sh[1][threadIdx.x]);
}

__global__ void select_digit_kernel(SelectState* st, int shift, int n_ranks, double* __restrict__ out) {
    __shared__ unsigned long long cs[256];
    __shared__ unsigned int digit_s;
    const int t = threadIdx.x;
    for (int s = 0; s < n_ranks; ++s) {
        cs[t] = st->hist[s][t];
        if (t == 0) digit_s = 0u;
        __syncthreads();
        for (int o = 1; o < 256; o <<= 1) {                       // inclusive scan
            const unsigned long long v = t >= o ? cs[t - o] : 0ull;
            __syncthreads();
            cs[t] += v;
            __syncthreads();
        }
        const unsigned long long rem = st->remaining[s];
        if (cs[t] <= rem) atomicAdd(&digit_s, 1u);
        __syncthreads();
        const unsigned int digit = digit_s < 255u ? digit_s : 255u;
        if (t == 0) {
            st->remaining[s] = rem - (digit > 0 ? cs[digit - 1] : 0ull);
            const unsigned long long p = (st->prefix[s] << 8) | (unsigned long long)digit;
            st->prefix[s] = p;
            if (shift == 0) out[s] = __longlong_as_double((long long)p);
        }
        st->hist[s][t] = 0ull;
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------ transition table
__global__ void transition_accumulate_kernel(const float2* __restrict__ partial, int k1, int64_t n,
                                             const double* __restrict__ norms, double c1, const double* __restrict__ f,
                                             double eps, double inv_m, const int* __restrict__ label_row, int k0,
                                             double* __restrict__ table) {
    extern __shared__ double tab[];  // [k0*k1]
    for (int t = threadIdx.x; t < k0 * k1; t += blockDim.x) tab[t] = 0.0;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int a = label_row[i];
        if (a < 0 || a >= k0) continue;
        const double base = f[i] / eps - norms[i] * c1;
        for (int b = 0; b < k1; ++b) {
            const float2 ps = partial[(int64_t)b * n + i];
            if (ps.x > -1e29f && ps.y > 0.f)
                atomicAdd(&tab[a * k1 + b], exp(base + SDB_LN2 * ((double)ps.x + log2((double)ps.y))) * inv_m);
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < k0 * k1; t += blockDim.x)
        if (tab[t] != 0.0) atomicAdd(table + t, tab[t]);
}

inline unsigned blocks_for(int64_t n, int bs = 256) { return (unsigned)((n + bs - 1) / bs); }

}  // namespace

extern "C" {

int sdb_version(void) { return SDB_VERSION; }

const char* sdb_error_string(int status) {
    switch (status) {
        case 0: return "ok";
        case SDB_E_INVALID: return "spadot_b200: invalid argument";
        case SDB_E_UNSUPPORTED: return "spadot_b200: unsupported shape";
        case SDB_E_NOTFINITE: return "spadot_b200: non-finite value";
        case SDB_E_DRIVER: return "spadot_b200: CUDA driver entry point or libnccl unavailable";
        case SDB_E_NCCL: return "spadot_b200: NCCL call failed";
        default: return status > 0 ? cudaGetErrorString((cudaError_t)status) : "spadot_b200: unknown status";
    }
}

int sdb_device_check(void) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) return (int)e;
    return prop.major == 10 ? 0 : SDB_E_UNSUPPORTED;
}

int sdb_column_sums_f64(const double* x, int64_t n, int d, double* out, void* stream) {
    SDB_CHECK_ARG(x && out && n >= 0 && d > 0);
    if (n == 0) return 0;
    int grid = (int)((n + 4095) / 4096);
    if (grid > 592) grid = 592;
    column_sums_kernel<<<grid, 256, 256 * sizeof(double), sdb_stream(stream)>>>(x, n, d, out);
    SDB_LAUNCH_STATUS();
}

int sdb_prep_points_f64(const double* x, int64_t n, int d, const double* center, float* xt, int64_t ld, int dpad,
                        double* norms, void* stream) {
    SDB_CHECK_ARG(x && center && xt && norms && n >= 0 && d > 0 && dpad >= d && ld >= n);
    prep_points_kernel<<<blocks_for(ld), 256, 0, sdb_stream(stream)>>>(x, n, d, center, xt, ld, dpad, norms);
    SDB_LAUNCH_STATUS();
}

int sdb_lse_finalize(const float* partial, int n_splits, int64_t n, const double* norms, double c1, double* L, void* stream) {
    SDB_CHECK_ARG(partial && norms && L && n_splits > 0 && n >= 0);
    if (n == 0) return 0;
    return sdb_lse_finalize_pred(partial, n_splits, n, norms, c1, L, nullptr, nullptr, stream);
}

int sdb_lse_finalize_pred(const float* partial, int n_splits, int64_t n, const double* norms, double c1, double* L, float* m_next,
                          int* bad_flag, void* stream) {
    SDB_CHECK_ARG(partial && norms && L && n_splits > 0 && n >= 0);
    if (n == 0) return 0;
    return (int)sdb_launch(lse_finalize_kernel, dim3(blocks_for(n)), dim3(256), 0, sdb_stream(stream),
                           reinterpret_cast<const float2*>(partial), n_splits, n, norms, c1, L, m_next, bad_flag);
}

int sdb_potential_update(int64_t n, const double* L, const double* logmarg, const double* norms, double eps, double alpha,
                         double log_n_other, double c1, double* pot, const double* frame, double* la_old, float* bias,
                         int* absorb_flag, int iter, double log_tau, double log_floor, void* stream) {
    SDB_CHECK_ARG(L && logmarg && norms && pot && n >= 0 && eps > 0.0);
    if (n == 0) return 0;
    return (int)sdb_launch(potential_update_kernel, dim3(blocks_for(n)), dim3(256), 0, sdb_stream(stream), n, L, logmarg, norms, eps,
                           alpha, log_n_other, c1, pot, frame, la_old, bias, absorb_flag, iter, log_tau, log_floor);
}

int sdb_finalize_update(const float* partial, int n_splits, int64_t n, const double* norms, double c1, double* L,
                        const double* logmarg, double eps, double alpha, double log_n_other, double* pot, const double* frame,
                        double* la_old, float* bias, int* absorb_flag, int iter, double log_tau, double log_floor, void* stream) {
    SDB_CHECK_ARG(partial && norms && L && logmarg && pot && frame && la_old && bias && absorb_flag && n_splits > 0 && n >= 0 && eps > 0.0);
    if (n == 0) return 0;
    return sdb_finalize_update_pred(partial, n_splits, n, norms, c1, L, logmarg, eps, alpha, log_n_other, pot, frame, la_old, bias,
                                    absorb_flag, iter, log_tau, log_floor, nullptr, nullptr, stream);
}

int sdb_potential_update_deferred(int64_t n, const double* L, const double* logmarg, const double* norms, double eps, double alpha,
                                  double log_n_other, double c1, double* pot, double* frame, double* la_old, float* bias,
                                  int* flag2, int iter, double log_tau, double log_floor, void* stream) {
    SDB_CHECK_ARG(L && logmarg && norms && pot && frame && la_old && flag2 && n >= 0 && eps > 0.0);
    if (n == 0) return 0;
    return (int)sdb_launch(potential_update_deferred_kernel, dim3(blocks_for(n)), dim3(256), 0, sdb_stream(stream), n, L, logmarg, norms,
                           eps, alpha, log_n_other, c1, pot, frame, la_old, bias, flag2, iter, log_tau, log_floor);
}

int sdb_absorb_pending(int64_t n, int64_t m, const int* flag2, int last_tick, const double* f, const double* g, double* u,
                       double* v, int* flag, void* stream) {
    SDB_CHECK_ARG(flag2 && f && g && u && v && n >= 0 && m >= 0);
    const int64_t mx = n > m ? n : m;
    if (mx == 0) return 0;
    return (int)sdb_launch(absorb_pending_kernel, dim3(blocks_for(mx)), dim3(256), 0, sdb_stream(stream), n, m, flag2, last_tick, f, g, u,
                           v, flag);
}

int sdb_partial_sums_f64(const float* partial, int n_splits, int64_t n, const float* shift, double* sums, const int* absorb_flag,
                         int iter, const int* bad_flag, void* stream) {
    SDB_CHECK_ARG(partial && shift && sums && n_splits > 0 && n >= 0);
    return (int)sdb_launch(partial_sums_kernel, dim3(blocks_for(n > 0 ? n : 1)), dim3(256), 0, sdb_stream(stream),
                           reinterpret_cast<const float2*>(partial), n_splits, n, shift, sums, absorb_flag, iter, bad_flag);
}

int sdb_update_from_sums_f64(const double* sums, float* shift, int64_t n, const double* norms, double c1, double* L,
                             const double* logmarg, double eps, double alpha, double log_n_other, double* pot, const double* frame,
                             double* la_old, float* bias, int* absorb_flag, int iter, double log_tau, double log_floor, int* bad_flag,
                             void* stream) {
    SDB_CHECK_ARG(sums && shift && norms && L && logmarg && pot && frame && la_old && bias && absorb_flag && n > 0 && eps > 0.0);
    return (int)sdb_launch(update_from_sums_kernel, dim3(blocks_for(n)), dim3(256), 0, sdb_stream(stream), sums, shift, n, norms, c1, L,
                           logmarg, eps, alpha, log_n_other, pot, frame, la_old, bias, absorb_flag, iter, log_tau, log_floor, bad_flag);
}

int sdb_finalize_update_pred(const float* partial, int n_splits, int64_t n, const double* norms, double c1, double* L,
                             const double* logmarg, double eps, double alpha, double log_n_other, double* pot, const double* frame,
                             double* la_old, float* bias, int* absorb_flag, int iter, double log_tau, double log_floor, float* m_next,
                             int* bad_flag, void* stream) {
    SDB_CHECK_ARG(partial && norms && L && logmarg && pot && frame && la_old && bias && absorb_flag && n_splits > 0 && n >= 0 && eps > 0.0);
    if (n == 0) return 0;
    return (int)sdb_launch(finalize_update_kernel, dim3(blocks_for(n)), dim3(256), 0, sdb_stream(stream),
                           reinterpret_cast<const float2*>(partial), n_splits, n, norms, c1, L, logmarg, eps, alpha, log_n_other, pot,
                           frame, la_old, bias, absorb_flag, iter, log_tau, log_floor, m_next, bad_flag);
}

int sdb_make_bias(int64_t n, int64_t n_pad, const double* pot, const double* norms, double eps, double c1, float* bias,
                  void* stream) {
    SDB_CHECK_ARG(norms && bias && n >= 0 && n_pad >= n && eps > 0.0);
    if (n_pad == 0) return 0;
    return (int)sdb_launch(make_bias_kernel, dim3(blocks_for(n_pad)), dim3(256), 0, sdb_stream(stream), n, n_pad, pot, norms, eps, c1,
                           bias);
}

int sdb_absorb(int64_t n, int64_t m, const int* absorb_flag, int iter, const double* f, const double* g, double* u, double* v,
               void* stream) {
    SDB_CHECK_ARG(absorb_flag && f && g && u && v);
    const int64_t mx = n > m ? n : m;
    if (mx == 0) return 0;
    return (int)sdb_launch(absorb_kernel, dim3(blocks_for(mx)), dim3(256), 0, sdb_stream(stream), n, m, absorb_flag, iter, f, g, u, v);
}

int sdb_stage_criterion(int64_t n, int64_t m, const double* f, const double* u, const double* la_old, const double* g,
                        const double* v, const double* lb_old, double eps, double* out4, void* scratch, void* stream) {
    SDB_CHECK_ARG(f && u && la_old && g && v && lb_old && out4 && scratch);
    stage_criterion_kernel<<<sdb_reduce_grid(n > m ? n : m), 256, 0, sdb_stream(stream)>>>(n, m, f, u, la_old, g, v, lb_old, eps, out4, scratch);
    SDB_LAUNCH_STATUS();
}

int sdb_gap_terms(int64_t n, int64_t m, const double* f, const double* Lr, const double* logp, const double* g,
                  const double* Lc, const double* logq, double eps, double lam1, double lam2, double dx, double dy,
                  double* out10, void* scratch, void* stream) {
    SDB_CHECK_ARG(f && Lr && logp && g && Lc && logq && out10 && scratch);
    gap_terms_kernel<<<sdb_reduce_grid(n > m ? n : m), 256, 0, sdb_stream(stream)>>>(n, m, f, Lr, logp, g, Lc, logq, eps, lam1, lam2, dx, dy, out10, scratch);
    SDB_LAUNCH_STATUS();
}

int sdb_sum_exp(int64_t n, const double* L, const double* add, double add_scale, double* out1, void* scratch, void* stream) {
    SDB_CHECK_ARG(L && out1 && scratch);
    sum_exp_kernel<<<sdb_reduce_grid(n), 256, 0, sdb_stream(stream)>>>(n, L, add, add_scale, out1, scratch);
    SDB_LAUNCH_STATUS();
}

int sdb_row_mass(int64_t n, const double* f, const double* Lr, double eps, double inv_m, double* out, void* stream) {
    SDB_CHECK_ARG(f && Lr && out);
    if (n == 0) return 0;
    row_mass_kernel<<<blocks_for(n), 256, 0, sdb_stream(stream)>>>(n, f, Lr, eps, inv_m, out);
    SDB_LAUNCH_STATUS();
}

int sdb_plan_dense_f64(const double* x, const double* y, int64_t n, int64_t m, int d, const double* f, const double* g,
                       double inv_med, double eps, double inv_m, double* plan, void* stream) {
    SDB_CHECK_ARG(x && y && f && g && plan && d > 0 && d <= 400);   // 2*32*(d+1) doubles of shared memory
    if (n == 0 || m == 0) return 0;
    dim3 grid((unsigned)((m + 31) / 32), (unsigned)((n + 31) / 32));
    if (grid.y > 65535) return SDB_E_UNSUPPORTED;
    const size_t smem = sizeof(double) * 2 * 32 * (d + 1);
    cudaError_t e = cudaFuncSetAttribute(plan_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    plan_dense_kernel<<<grid, 256, smem, sdb_stream(stream)>>>(x, y, n, m, d, f, g, inv_med, eps, inv_m, plan);
    SDB_LAUNCH_STATUS();
}

int sdb_pair_distances_f64(const double* x, const double* y, int d, const int64_t* ii, const int64_t* jj, int64_t n_pairs,
                           double* dist, void* stream) {
    SDB_CHECK_ARG(x && y && ii && jj && dist && d > 0);
    if (n_pairs == 0) return 0;
    pair_distances_kernel<<<blocks_for(n_pairs), 256, 0, sdb_stream(stream)>>>(x, y, d, ii, jj, n_pairs, dist);
    SDB_LAUNCH_STATUS();
}

int sdb_radix_digit_hist(const double* cand, unsigned long long n, int shift, unsigned long long prefix,
                         unsigned long long* hist256, void* stream) {
    SDB_CHECK_ARG(cand && hist256 && shift >= 0 && shift <= 56 && (shift % 8) == 0);
    if (n == 0) return 0;
    unsigned grid = (unsigned)((n + 256 * 16 - 1) / (256 * 16));
    if (grid > 1184) grid = 1184;
    radix_digit_hist_kernel<<<grid, 256, 0, sdb_stream(stream)>>>(cand, n, shift, prefix, hist256);
    SDB_LAUNCH_STATUS();
}

int sdb_select_ranks_f64(const double* cand, unsigned long long n, const unsigned long long* ranks, int n_ranks, double* out,
                         void* workspace, void* stream) {
    SDB_CHECK_ARG(cand && ranks && out && workspace && n > 0 && (n_ranks == 1 || n_ranks == 2));
    for (int s = 0; s < n_ranks; ++s) SDB_CHECK_ARG(ranks[s] < n);
    cudaStream_t st = sdb_stream(stream);
    SelectState* state = reinterpret_cast<SelectState*>(workspace);
    select_init_kernel<<<1, 256, 0, st>>>(state, ranks[0], n_ranks > 1 ? ranks[1] : ranks[0]);
    unsigned grid = (unsigned)((n + 256 * 16 - 1) / (256 * 16));
    if (grid > 1184) grid = 1184;
    for (int shift = 56; shift >= 0; shift -= 8) {
        select_hist_kernel<<<grid, 256, 0, st>>>(cand, n, shift, n_ranks, state);
        select_digit_kernel<<<1, 256, 0, st>>>(state, shift, n_ranks, out);
    }
    SDB_LAUNCH_STATUS();
}

int sdb_transition_accumulate(const float* partial, int k1, int64_t n, const double* norms, double c1, const double* f,
                              double eps, double inv_m, const int* label_row, int k0, double* table, void* stream) {
    SDB_CHECK_ARG(partial && norms && f && label_row && table && k0 > 0 && k1 > 0 && k0 * k1 <= 4096);
    if (n == 0) return 0;
    unsigned grid = blocks_for(n);
    if (grid > 592) grid = 592;
    transition_accumulate_kernel<<<grid, 256, sizeof(double) * k0 * k1, sdb_stream(stream)>>>(
        reinterpret_cast<const float2*>(partial), k1, n, norms, c1, f, eps, inv_m, label_row, k0, table);
    SDB_LAUNCH_STATUS();
}

}  // extern "C"
