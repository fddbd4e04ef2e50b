// Dense-cost path: the same solver for callers that hand over an explicit N x M fp64 cost matrix
// (the reference's optimal_transport_duality_gap(C, ...) / transport_stablev2(C, ...) signatures,
// ref: SpaDOT/utils/OT_loss/ot_solvers.py:164,452 — this is the shape wot's `solver` slot calls).
// Nothing else of size N x M is created: K, _K and R of the reference (ot_solvers.py:266-274) are
// replaced by fp64 log-sum-exp sweeps over C, so each half-iteration reads C exactly once (HBM-bound,
// 8*N*M bytes) instead of the reference's gemv over K plus exp passes (ot_func.cpp:43-249,547-568).
#include "sdb_common.cuh"

namespace {

struct MS { double m, s; };

__device__ __forceinline__ void ms_add(MS& a, double t) {
    if (t > a.m) { a.s = a.s * exp(a.m - t) + 1.0; a.m = t; }
    else if (t > -INFINITY) a.s += exp(t - a.m);
}
__device__ __forceinline__ void ms_merge(MS& a, double m2, double s2) {
    if (m2 == -INFINITY) return;
    if (a.m == -INFINITY) { a.m = m2; a.s = s2; return; }
    const double mn = fmax(a.m, m2);
    a.s = a.s * exp(a.m - mn) + s2 * exp(m2 - mn);
    a.m = mn;
}

// one warp per row; lanes stride the row (coalesced 256 B per warp-load)
__global__ void __launch_bounds__(256) dense_row_lse_kernel(const double* __restrict__ C, int64_t ldc, int64_t n, int64_t m,
                                                            const double* __restrict__ g, double inv_eps,
                                                            double* __restrict__ L) {
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= n) return;
    const int lane = threadIdx.x & 31;
    const double* c = C + row * ldc;
    MS a{-INFINITY, 0.0};
    for (int64_t j = lane; j < m; j += 32) ms_add(a, ((g ? g[j] : 0.0) - c[j]) * inv_eps);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double m2 = __shfl_xor_sync(0xffffffffu, a.m, o);
        const double s2 = __shfl_xor_sync(0xffffffffu, a.s, o);
        ms_merge(a, m2, s2);
    }
    if (lane == 0) L[row] = (a.m == -INFINITY) ? -INFINITY : a.m + log(a.s);
}

// thread per column, blockIdx.y selects a contiguous chunk of rows; partial (max,sum) per (chunk, column)
__global__ void __launch_bounds__(256) dense_col_lse_partial_kernel(const double* __restrict__ C, int64_t ldc, int64_t n,
                                                                    int64_t m, const double* __restrict__ f, double inv_eps,
                                                                    int rows_per_chunk, double2* __restrict__ partial) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const int64_t i0 = (int64_t)blockIdx.y * rows_per_chunk;
    const int64_t i1 = min(n, i0 + rows_per_chunk);
    MS a{-INFINITY, 0.0};
    for (int64_t i = i0; i < i1; ++i) ms_add(a, ((f ? f[i] : 0.0) - C[i * ldc + j]) * inv_eps);
    partial[(int64_t)blockIdx.y * m + j] = make_double2(a.m, a.s);
}

__global__ void dense_col_lse_finish_kernel(const double2* __restrict__ partial, int n_chunks, int64_t m, double* __restrict__ L) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    MS a{-INFINITY, 0.0};
    for (int c = 0; c < n_chunks; ++c) {
        const double2 p = partial[(int64_t)c * m + j];
        ms_merge(a, p.x, p.y);
    }
    L[j] = (a.m == -INFINITY) ? -INFINITY : a.m + log(a.s);
}

__global__ void dense_plan_kernel(const double* __restrict__ C, int64_t ldc, int64_t n, int64_t m, const double* __restrict__ f,
                                  const double* __restrict__ g, double inv_eps, double inv_m, double* __restrict__ plan) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i = blockIdx.y;
    if (j >= m || i >= n) return;
    plan[i * m + j] = exp((f[i] + g[j] - C[i * ldc + j]) * inv_eps) * inv_m;
}

}  // namespace

extern "C" {

int sdb_dense_row_lse_f64(const double* C, int64_t ldc, int64_t n, int64_t m, const double* g, double eps, double* L, void* stream) {
    SDB_CHECK_ARG(C && L && n >= 0 && m >= 0 && ldc >= m && eps > 0.0);
    if (n == 0) return 0;
    dense_row_lse_kernel<<<(unsigned)((n + 7) / 8), 256, 0, sdb_stream(stream)>>>(C, ldc, n, m, g, 1.0 / eps, L);
    SDB_LAUNCH_STATUS();
}

int sdb_dense_col_lse_f64(const double* C, int64_t ldc, int64_t n, int64_t m, const double* f, double eps, double* L,
                          double* partial, int n_chunks, void* stream) {
    SDB_CHECK_ARG(C && L && partial && n >= 0 && m >= 0 && ldc >= m && eps > 0.0 && n_chunks > 0 && n_chunks <= 65535);
    if (m == 0) return 0;
    const int rows_per_chunk = (int)((n + n_chunks - 1) / n_chunks);
    dim3 grid((unsigned)((m + 255) / 256), (unsigned)n_chunks);
    dense_col_lse_partial_kernel<<<grid, 256, 0, sdb_stream(stream)>>>(C, ldc, n, m, f, 1.0 / eps, rows_per_chunk > 0 ? rows_per_chunk : 1,
                                                                      reinterpret_cast<double2*>(partial));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    dense_col_lse_finish_kernel<<<(unsigned)((m + 255) / 256), 256, 0, sdb_stream(stream)>>>(reinterpret_cast<const double2*>(partial),
                                                                                            n_chunks, m, L);
    SDB_LAUNCH_STATUS();
}

int sdb_dense_plan_f64(const double* C, int64_t ldc, int64_t n, int64_t m, const double* f, const double* g, double eps,
                       double inv_m, double* plan, void* stream) {
    SDB_CHECK_ARG(C && f && g && plan && ldc >= m && eps > 0.0 && n <= 65535);
    if (n == 0 || m == 0) return 0;
    dim3 grid((unsigned)((m + 255) / 256), (unsigned)n);
    dense_plan_kernel<<<grid, 256, 0, sdb_stream(stream)>>>(C, ldc, n, m, f, g, 1.0 / eps, inv_m, plan);
    SDB_LAUNCH_STATUS();
}

}  // extern "C"
