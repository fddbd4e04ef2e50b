// libot_b200.so — the reference's libot.so ABI (include/libot_b200.h) on B200 kernels (sm_100a only).
//
// What the reference does with one CPU thread over caller-owned dense host matrices
// (SpaDOT/utils/OT_loss/ot_func.cpp) is done here on the device: each entry point uploads its operands, runs the
// kernels below and downloads exactly what the reference mutates.  The expressions are the reference's (same
// divisions, pow / exp / log placement); long sums are tree reductions, so results agree to fp64 rounding.
//
// HBM picture of one Sinkhorn iteration (step1_process): K (m x n) is read twice — once by the row kernel
// (K (b*dy), one warp per row, coalesced along the row) and once by the column kernel (K^T (a*dx), one thread per
// column inside a row chunk, coalesced across threads) — which is the reference's own traffic (gemv + gemtv,
// ot_func.cpp:43-249); algorithmic bytes per iteration = 2 * m * n * sizeof(T).  The stabilisation (tau) decision is
// taken on the device: the update kernels raise a flag, and the absorption / K-rebuild kernels are launched every
// iteration and return immediately unless it is set, so the iteration loop never synchronises with the host.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "../../../include/libot_b200.h"

namespace {

std::mutex g_mu;
long long g_launches = 0, g_h2d = 0, g_d2h = 0;
int g_device_ok = -1;

// LIBOT_B200_TRACE=1: one stderr line per entry point with its wall time split (development aid; adds syncs).
bool tracing() {
    static int on = -1;
    if (on < 0) { const char* e = getenv("LIBOT_B200_TRACE"); on = (e && *e && *e != '0') ? 1 : 0; }
    return on == 1;
}
struct Trace {
    const char* name;
    std::chrono::steady_clock::time_point t0, last;
    double h2d = 0, d2h = 0, compute = 0, alloc = 0;
    explicit Trace(const char* n) : name(n), t0(std::chrono::steady_clock::now()), last(t0) {}
    double lap() {
        auto now = std::chrono::steady_clock::now();
        double ms = std::chrono::duration<double, std::milli>(now - last).count();
        last = now;
        return ms;
    }
    ~Trace() {
        if (!tracing()) return;
        double total = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        fprintf(stderr, "libot_b200 %-22s total %9.3f ms  alloc %8.3f  h2d %8.3f  compute %8.3f  d2h %8.3f\n", name, total, alloc, h2d,
                compute, d2h);
    }
};
Trace* g_trace = nullptr;
struct TraceScope {
    Trace t;
    explicit TraceScope(const char* n) : t(n) { g_trace = &t; }
    ~TraceScope() { g_trace = nullptr; }
};
void trace_compute_done() {
    if (tracing() && g_trace) { cudaDeviceSynchronize(); g_trace->compute += g_trace->lap(); }
}

[[noreturn]] void die(const char* what, cudaError_t e) {
    fprintf(stderr, "libot_b200: %s failed: %s. This library runs on a B200 (sm_100a) only and has no CPU fallback.\n", what,
            cudaGetErrorString(e));
    fflush(stderr);
    abort();
}
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) die(#x, e_); } while (0)
#define LAUNCHED() do { CK(cudaGetLastError()); ++g_launches; } while (0)

int device_check() {
    if (g_device_ok >= 0) return g_device_ok ? 0 : 1;
    int n = 0, dev = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { cudaGetLastError(); return 1; }
    cudaDeviceProp prop;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) { cudaGetLastError(); return 1; }
    g_device_ok = prop.major == 10 ? 1 : 0;
    return g_device_ok ? 0 : 1;
}

void require_device() {
    if (device_check() != 0) {
        fprintf(stderr, "libot_b200: no usable sm_100 (B200) CUDA device. This library has no CPU fallback.\n");
        fflush(stderr);
        abort();
    }
}

// Device memory pool.  cudaMalloc / cudaFree of matrix-sized blocks cost 1-150 ms each on the GPU box (measured with
// LIBOT_B200_TRACE: more than the PCIe copies they serve), so freed blocks are kept and handed out again: a request
// takes the smallest pooled block of at least its size and at most twice its size.  The pool only ever holds what
// one call had live at once; libot_b200_release() (or LIBOT_B200_POOL=0) gives the memory back.
struct PoolBlock { void* p; size_t bytes; };
std::vector<PoolBlock> g_pool;
bool pooling() {
    static int on = -1;
    if (on < 0) { const char* e = getenv("LIBOT_B200_POOL"); on = (e && *e == '0') ? 0 : 1; }
    return on == 1;
}
void* pool_acquire(size_t bytes, size_t* got) {
    int best = -1;
    for (int i = 0; i < (int)g_pool.size(); ++i)
        if (g_pool[i].bytes >= bytes && g_pool[i].bytes <= 2 * bytes + 4096 && (best < 0 || g_pool[i].bytes < g_pool[best].bytes)) best = i;
    if (best >= 0) {
        PoolBlock b = g_pool[best];
        g_pool.erase(g_pool.begin() + best);
        *got = b.bytes;
        return b.p;
    }
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaErrorMemoryAllocation && !g_pool.empty()) {          // make room and retry once
        cudaGetLastError();
        for (auto& b : g_pool) cudaFree(b.p);
        g_pool.clear();
        e = cudaMalloc(&p, bytes);
    }
    if (e != cudaSuccess) die("cudaMalloc", e);
    *got = bytes;
    return p;
}
void pool_release(void* p, size_t bytes) {
    if (!p) return;
    if (pooling()) g_pool.push_back({p, bytes});
    else cudaFree(p);
}
void pool_clear() {
    for (auto& b : g_pool) cudaFree(b.p);
    g_pool.clear();
}

// Device buffer living for one call (the reference mallocs / frees its temporaries per call as well).
template <typename T>
struct Buf {
    T* d = nullptr;
    size_t n = 0, cap_bytes = 0;
    Buf() {}
    explicit Buf(size_t count) { alloc(count); }
    Buf(const Buf&) = delete;
    Buf& operator=(const Buf&) = delete;
    ~Buf() { pool_release(d, cap_bytes); }
    void alloc(size_t count) {
        n = count;
        if (tracing() && g_trace) g_trace->lap();
        d = static_cast<T*>(pool_acquire(std::max<size_t>(count, 1) * sizeof(T), &cap_bytes));
        if (tracing() && g_trace) g_trace->alloc += g_trace->lap();
    }
    void up(const T* h) {
        if (n) CK(cudaMemcpy(d, h, n * sizeof(T), cudaMemcpyHostToDevice));
        g_h2d += (long long)(n * sizeof(T));
        if (tracing() && g_trace) g_trace->h2d += g_trace->lap();
    }
    void down(T* h) const {
        if (tracing() && g_trace) { cudaDeviceSynchronize(); g_trace->compute += g_trace->lap(); }
        if (n) CK(cudaMemcpy(h, d, n * sizeof(T), cudaMemcpyDeviceToHost));
        g_d2h += (long long)(n * sizeof(T));
        if (tracing() && g_trace) g_trace->d2h += g_trace->lap();
    }
    void zero() { CK(cudaMemset(d, 0, std::max<size_t>(n, 1) * sizeof(T))); }
};

template <typename T>
Buf<T>* uploaded(Buf<T>& b, const T* h, size_t count) {
    b.alloc(count);
    b.up(h);
    return &b;
}

// ------------------------------------------------------------------------------------------------ device helpers
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// K-wide block sum (blockDim.x == 256); result valid in thread 0..K-1 of warp 0 as out[k]
template <typename T, int K>
__device__ void block_sum_store(T (&vals)[K], T* dst) {
    __shared__ T sm[8][K];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        T v = warp_sum(vals[k]);
        if (lane == 0) sm[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < K) {
        T v = 0;
        for (int w = 0; w < 8; ++w) v += sm[w][threadIdx.x];
        dst[threadIdx.x] = v;
    }
}

// ot_func.cpp:30-40: +-inf -> +-MAXFLOAT (the float constant, also for doubles), everything else (NaN too) unchanged
template <typename T>
__device__ __forceinline__ T nan_to_num(T x) {
    if (isinf(x)) return x < 0 ? (T)(-FLT_MAX) : (T)FLT_MAX;
    return x;
}

// ------------------------------------------------------------------------------------------------ kernels
// update_k (ot_func.cpp:547-568) and the K rebuild of an absorption (:805-809; then K_ == nullptr and the kernel only
// runs when *flag == it).
template <typename T>
__global__ void k_update_k(T* __restrict__ K, T* __restrict__ K_, const T* __restrict__ C, const T* __restrict__ u,
                           const T* __restrict__ v, T eps, int m, int n, const int* flag, int it) {
    if (flag != nullptr && *flag != it) return;
    for (int i = blockIdx.y; i < m; i += gridDim.y) {
        const size_t row = (size_t)i * n;
        const T ui = u[i];
        for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
            const T c = C[row + j];
            if (K_ != nullptr) K_[row + j] = exp(-c / eps);
            K[row + j] = exp((ui + v[j] - c) / eps);
        }
    }
}

// update_R (ot_func.cpp:571-584)
template <typename T>
__global__ void k_update_R(T* __restrict__ R, const T* __restrict__ K, const T* __restrict__ a, const T* __restrict__ b, int m,
                           int n) {
    for (int i = blockIdx.y; i < m; i += gridDim.y) {
        const size_t row = (size_t)i * n;
        const T ai = a[i];
        for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x)
            R[row + j] = K[row + j] * ai * b[j];
    }
}

template <typename T>
__global__ void k_mul(const T* __restrict__ x, const T* __restrict__ y, T* __restrict__ out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = x[i] * y[i];
}

// First half of update_a_b (ot_func.cpp:610-636): s_i = sum_j K_ij (b_j dy_j) with one warp per row, then
// a_i = (p_i / s_i)^alpha1 * exp(-u_i / (lambda1 + eps)).  Also old_a <- a (:729-730), wa <- a*dx for the column
// kernel, and the tau flag (:778-784).
template <typename T>
__global__ void k_row_update(const T* __restrict__ K, const T* __restrict__ wb, const T* __restrict__ p, const T* __restrict__ u,
                             const T* __restrict__ dx, T* __restrict__ a, T* __restrict__ old_a, T* __restrict__ wa, T alpha1,
                             T lam_eps, T tau, int* flag, int it, int m, int n) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    for (int i = warp; i < m; i += n_warps) {
        const T* row = K + (size_t)i * n;
        T s0 = 0, s1 = 0, s2 = 0, s3 = 0;
        int j = lane;
        if (sizeof(T) == 8 && (n & 1) == 0 && ((uintptr_t)K & 15) == 0 && ((uintptr_t)wb & 15) == 0) {
            // 16-byte loads, four per lane in flight (2 KB per warp): the row is streamed once, this is the whole traffic
            const double2* r2 = reinterpret_cast<const double2*>(row);
            const double2* w2 = reinterpret_cast<const double2*>(wb);
            const int n2 = n >> 1;
            int q = lane;
            for (; q + 96 < n2; q += 128) {
                const double2 a0 = r2[q], a1 = r2[q + 32], a2 = r2[q + 64], a3 = r2[q + 96];
                const double2 b0 = w2[q], b1 = w2[q + 32], b2 = w2[q + 64], b3 = w2[q + 96];
                s0 += (T)(a0.x * b0.x + a0.y * b0.y);
                s1 += (T)(a1.x * b1.x + a1.y * b1.y);
                s2 += (T)(a2.x * b2.x + a2.y * b2.y);
                s3 += (T)(a3.x * b3.x + a3.y * b3.y);
            }
            for (; q < n2; q += 32) { const double2 a0 = r2[q], b0 = w2[q]; s0 += (T)(a0.x * b0.x + a0.y * b0.y); }
            j = n;
        }
        for (; j + 96 < n; j += 128) {
            s0 += row[j] * wb[j];
            s1 += row[j + 32] * wb[j + 32];
            s2 += row[j + 64] * wb[j + 64];
            s3 += row[j + 96] * wb[j + 96];
        }
        for (; j < n; j += 32) s0 += row[j] * wb[j];
        const T s = warp_sum((s0 + s1) + (s2 + s3));
        if (lane == 0) {
            const T an = pow(p[i] / s, alpha1) * exp(-u[i] / lam_eps);
            old_a[i] = a[i];
            a[i] = an;
            wa[i] = an * dx[i];
            if (an > tau) *flag = it;
        }
    }
}

// Column sums of M weighted by w over one chunk of rows: part[chunk][j] = sum_{i in chunk} M_ij w_i
// (gemtv, ot_func.cpp:173-249; also R^T dx of primal, :405-409).  One thread per column, coalesced across threads.
template <typename T>
__global__ void k_col_partial(const T* __restrict__ M, const T* __restrict__ w, T* __restrict__ part, int m, int n,
                              int rows_per_chunk) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int i0 = blockIdx.y * rows_per_chunk;
    const int i1 = min(m, i0 + rows_per_chunk);
    T s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    int i = i0;
    for (; i + 7 < i1; i += 8) {            // eight rows in flight per thread (2 KB per warp), streamed once
        const T m0 = M[(size_t)i * n + j], m1 = M[(size_t)(i + 1) * n + j];
        const T m2 = M[(size_t)(i + 2) * n + j], m3 = M[(size_t)(i + 3) * n + j];
        const T m4 = M[(size_t)(i + 4) * n + j], m5 = M[(size_t)(i + 5) * n + j];
        const T m6 = M[(size_t)(i + 6) * n + j], m7 = M[(size_t)(i + 7) * n + j];
        s0 += m0 * w[i];
        s1 += m1 * w[i + 1];
        s2 += m2 * w[i + 2];
        s3 += m3 * w[i + 3];
        s0 += m4 * w[i + 4];
        s1 += m5 * w[i + 5];
        s2 += m6 * w[i + 6];
        s3 += m7 * w[i + 7];
    }
    for (; i + 3 < i1; i += 4) {
        s0 += M[(size_t)i * n + j] * w[i];
        s1 += M[(size_t)(i + 1) * n + j] * w[i + 1];
        s2 += M[(size_t)(i + 2) * n + j] * w[i + 2];
        s3 += M[(size_t)(i + 3) * n + j] * w[i + 3];
    }
    for (; i < i1; ++i) s0 += M[(size_t)i * n + j] * w[i];
    part[(size_t)blockIdx.y * n + j] = (s0 + s1) + (s2 + s3);
}

// Second half of update_a_b (ot_func.cpp:642-668): b_j = (q_j / s_j)^alpha2 * exp(-v_j / (lambda2 + eps)).
template <typename T>
__global__ void k_col_finish(const T* __restrict__ part, int chunks, const T* __restrict__ q, const T* __restrict__ v,
                             const T* __restrict__ dy, T* __restrict__ b, T* __restrict__ old_b, T* __restrict__ wb, T alpha2,
                             T lam_eps, T tau, int* flag, int it, int n) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    T s = 0;
    for (int c = 0; c < chunks; ++c) s += part[(size_t)c * n + j];
    const T bn = pow(q[j] / s, alpha2) * exp(-v[j] / lam_eps);
    old_b[j] = b[j];
    b[j] = bn;
    wb[j] = bn * dy[j];
    if (bn > tau) *flag = it;
}

// Absorption, vector part (ot_func.cpp:792-803, 811-817): only when the flag was raised in iteration `it`.
template <typename T>
__global__ void k_absorb_vec(T* __restrict__ a, T* __restrict__ b, T* __restrict__ u, T* __restrict__ v, T* __restrict__ wa,
                             T* __restrict__ wb, const T* __restrict__ dx, const T* __restrict__ dy, T eps, int m, int n,
                             const int* flag, int it, int* dirty) {
    if (*flag != it) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) {
        u[i] = u[i] + eps * log(a[i]);
        a[i] = 1;
        wa[i] = dx[i];
    }
    if (i < n) {
        v[i] = v[i] + eps * log(b[i]);
        b[i] = 1;
        wb[i] = dy[i];
    }
    if (i == 0) *dirty = 1;
}

// Dual-evolution criterion of the first stages (ot_func.cpp:878-922): per-block partial sums of
// |_a - old_a e^{u/eps}|^2, |_a|^2 and the same for b.
template <typename T>
__global__ void k_criterion(const T* __restrict__ a, const T* __restrict__ old_a, const T* __restrict__ u, int m,
                            const T* __restrict__ b, const T* __restrict__ old_b, const T* __restrict__ v, int n, T eps,
                            T* __restrict__ part) {
    T s[4] = {0, 0, 0, 0};
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) {
        const T e = exp(u[i] / eps);
        const T ra = a[i] * e;
        const T t = ra - old_a[i] * e;
        s[0] += t * t;
        s[1] += ra * ra;
    }
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        const T e = exp(v[j] / eps);
        const T rb = b[j] * e;
        const T t = rb - old_b[j] * e;
        s[2] += t * t;
        s[3] += rb * rb;
    }
    block_sum_store<T, 4>(s, part + (size_t)blockIdx.x * 4);
}

// Matrix part of primal / dual (ot_func.cpp:397-431, 482-483), one warp per row:
//   t1_i = sum_j R_ij dy_j;  per-warp partials of  sum R*nan_to_num(log R) - R + K_,  sum R*C,  sum R - K_.
// COMPUTE_R: R_ij = K_ij a_i b_j is formed here (update_R, :571-584) and written out.
template <typename T, bool COMPUTE_R>
__global__ void k_gap_rows(const T* __restrict__ K, const T* __restrict__ K_, const T* __restrict__ C, T* __restrict__ R,
                           const T* __restrict__ a, const T* __restrict__ b, const T* __restrict__ dy, T* __restrict__ t1,
                           T* __restrict__ part, int m, int n) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    T s3 = 0, s4 = 0, am = 0;
    for (int i = warp; i < m; i += n_warps) {
        const size_t row = (size_t)i * n;
        const T ai = COMPUTE_R ? a[i] : (T)0;
        T rs = 0;
        for (int j = lane; j < n; j += 32) {
            T r;
            if (COMPUTE_R) {
                r = K[row + j] * ai * b[j];
                R[row + j] = r;
            } else {
                r = R[row + j];
            }
            const T k_ = K_[row + j];
            rs += r * dy[j];
            s3 += r * nan_to_num(log(r)) - r + k_;
            s4 += r * C[row + j];
            am += r - k_;
        }
        rs = warp_sum(rs);
        if (lane == 0) t1[i] = rs;
    }
    s3 = warp_sum(s3);
    s4 = warp_sum(s4);
    am = warp_sum(am);
    if (lane == 0) {
        part[(size_t)warp * 3 + 0] = s3;
        part[(size_t)warp * 3 + 1] = s4;
        part[(size_t)warp * 3 + 2] = am;
    }
}

// Vector part of primal / dual: fdiv (ot_func.cpp:309-322) of (t1, p, dx) and (t2, q, dy), fdivstarexp (:341-355) of
// the real duals.  t2_j is finished here from the column partials.  With u != nullptr the real duals are
// a_i e^{u_i/eps}, b_j e^{v_j/eps} (update_process, :878-884); otherwise a, b are taken as given.
template <typename T>
__global__ void k_gap_vec(const T* __restrict__ t1, const T* __restrict__ p, const T* __restrict__ dx, int m,
                          const T* __restrict__ part2, int chunks, const T* __restrict__ q, const T* __restrict__ dy, int n,
                          const T* __restrict__ a, const T* __restrict__ b, const T* __restrict__ u, const T* __restrict__ v,
                          T eps, T lambda1, T lambda2, T* __restrict__ out_part) {
    T s[4] = {0, 0, 0, 0};
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) {
        const T x = t1[i];
        s[0] += dx[i] * (x * log(x / p[i]) - x + p[i]);
        const T ra = u != nullptr ? a[i] * exp(u[i] / eps) : a[i];
        s[2] += (p[i] * dx[i]) * (exp((-eps * log(ra)) / lambda1) - 1);
    }
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        T x = 0;
        for (int c = 0; c < chunks; ++c) x += part2[(size_t)c * n + j];
        s[1] += dy[j] * (x * log(x / q[j]) - x + q[j]);
        const T rb = v != nullptr ? b[j] * exp(v[j] / eps) : b[j];
        s[3] += (q[j] * dy[j]) * (exp((-eps * log(rb)) / lambda2) - 1);
    }
    block_sum_store<T, 4>(s, out_part + (size_t)blockIdx.x * 4);
}

// ------------------------------------------------------------------------------------------------ launch plans
constexpr int kThreads = 256;
constexpr int kRedBlocks = 128;       // blocks of the vector reductions (partials summed on the host, fixed order)

dim3 grid2d(int m, int n) {
    int gx = std::min((n + kThreads - 1) / kThreads, 64);
    int gy = std::min(m, 32768);
    return dim3((unsigned)std::max(gx, 1), (unsigned)std::max(gy, 1));
}

int row_warp_blocks(int m) {               // one warp per row, 8 warps per block, at most 16 blocks per SM in flight
    return std::max(1, std::min((m + 7) / 8, 148 * 16));
}

struct ColPlan { int tiles, chunks, rows_per_chunk; };
ColPlan col_plan(int m, int n) {
    ColPlan c;
    c.tiles = std::max(1, (n + kThreads - 1) / kThreads);
    int want = std::max(1, (148 * 8 + c.tiles - 1) / c.tiles);          // ~8 blocks per SM in total
    int max_chunks = std::max(1, (m + 31) / 32);                        // at least 32 rows per chunk
    c.chunks = std::min(std::min(want, max_chunks), 65535);
    c.rows_per_chunk = std::max(1, (m + c.chunks - 1) / c.chunks);
    c.chunks = std::max(1, (m + c.rows_per_chunk - 1) / c.rows_per_chunk);
    return c;
}

template <typename T>
void host_block_sums(const Buf<T>& part, int blocks, int width, double* out) {
    std::vector<T> h((size_t)blocks * width);
    CK(cudaMemcpy(h.data(), part.d, h.size() * sizeof(T), cudaMemcpyDeviceToHost));
    g_d2h += (long long)(h.size() * sizeof(T));
    for (int k = 0; k < width; ++k) {
        T acc = 0;                                  // accumulate in T like the reference does
        for (int b = 0; b < blocks; ++b) acc += h[(size_t)b * width + k];
        out[k] = (double)acc;
    }
}

// ------------------------------------------------------------------------------------------------ primal / dual
template <typename T>
struct GapParts { T pri, dua; };

// R on the device is either given (COMPUTE_R = false) or formed from the stabilised K, a, b and written to dR.
template <typename T, bool COMPUTE_R>
GapParts<T> gap_device(const T* dK, const T* dK_, const T* dC, T* dR, const T* da, const T* db, const T* du, const T* dv,
                       const T* ddx, const T* ddy, const T* dp, const T* dq, T eps, T lambda1, T lambda2, int m, int n) {
    const int rb = row_warp_blocks(m);
    const int n_warps = rb * (kThreads / 32);
    const ColPlan cp = col_plan(m, n);
    Buf<T> t1((size_t)m), part((size_t)n_warps * 3), part2((size_t)cp.chunks * n), vpart((size_t)kRedBlocks * 4);
    k_gap_rows<T, COMPUTE_R><<<rb, kThreads>>>(dK, dK_, dC, dR, da, db, ddy, t1.d, part.d, m, n);
    LAUNCHED();
    k_col_partial<T><<<dim3(cp.tiles, cp.chunks), kThreads>>>(dR, ddx, part2.d, m, n, cp.rows_per_chunk);
    LAUNCHED();
    k_gap_vec<T><<<kRedBlocks, kThreads>>>(t1.d, dp, ddx, m, part2.d, cp.chunks, dq, ddy, n, da, db, du, dv, eps, lambda1,
                                           lambda2, vpart.d);
    LAUNCHED();
    double ms[3], vs[4];
    host_block_sums(part, n_warps, 3, ms);
    host_block_sums(vpart, kRedBlocks, 4, vs);
    const T mn = (T)((double)m * (double)n);
    GapParts<T> g;
    g.pri = lambda1 * (T)vs[0] + lambda2 * (T)vs[1] + (eps * (T)ms[0] + (T)ms[1]) / mn;        // ot_func.cpp:433-435
    g.dua = -(lambda1 * (T)vs[2]) + -(lambda2 * (T)vs[3]) + -eps * (T)ms[2] / mn;              // ot_func.cpp:485-489
    return g;
}

template <typename T>
GapParts<T> gap_from_host(const T* C, const T* K, const T* R, const T* dx, const T* dy, const T* p, const T* q, const T* a,
                          const T* b, T eps, T lambda1, T lambda2, int m, int n) {
    std::lock_guard<std::mutex> lock(g_mu);
    TraceScope ts("primal/dual/gap");
    require_device();
    const size_t mn = (size_t)m * (size_t)n;
    Buf<T> dC, dK, dR, ddx, ddy, dp, dq, da, db;
    uploaded(dC, C, mn); uploaded(dK, K, mn); uploaded(dR, R, mn);
    uploaded(ddx, dx, (size_t)m); uploaded(ddy, dy, (size_t)n); uploaded(dp, p, (size_t)m); uploaded(dq, q, (size_t)n);
    uploaded(da, a, (size_t)m); uploaded(db, b, (size_t)n);
    if (m <= 0 || n <= 0) { GapParts<T> z; z.pri = 0; z.dua = 0; return z; }
    return gap_device<T, false>(nullptr, dK.d, dC.d, dR.d, da.d, db.d, (const T*)nullptr, (const T*)nullptr, ddx.d, ddy.d, dp.d,
                                dq.d, eps, lambda1, lambda2, m, n);
}

template <typename T>
void update_k_host(T* K, T* K_, const T* C, const T* u, const T* v, T eps, int m, int n) {
    std::lock_guard<std::mutex> lock(g_mu);
    TraceScope ts("update_k");
    require_device();
    if (m <= 0 || n <= 0) return;
    const size_t mn = (size_t)m * (size_t)n;
    Buf<T> dC, du, dv, dK(mn), dK_(mn);
    uploaded(dC, C, mn); uploaded(du, u, (size_t)m); uploaded(dv, v, (size_t)n);
    k_update_k<T><<<grid2d(m, n), kThreads>>>(dK.d, dK_.d, dC.d, du.d, dv.d, eps, m, n, nullptr, 0);
    LAUNCHED();
    dK_.down(K_);
    dK.down(K);
}

template <typename T>
void update_R_host(T* R, const T* K, const T* a, const T* b, int m, int n) {
    std::lock_guard<std::mutex> lock(g_mu);
    TraceScope ts("update_R");
    require_device();
    if (m <= 0 || n <= 0) return;
    const size_t mn = (size_t)m * (size_t)n;
    Buf<T> dK, da, db, dR(mn);
    uploaded(dK, K, mn); uploaded(da, a, (size_t)m); uploaded(db, b, (size_t)n);
    k_update_R<T><<<grid2d(m, n), kThreads>>>(dR.d, dK.d, da.d, db.d, m, n);
    LAUNCHED();
    dR.down(R);
}

// ------------------------------------------------------------------------------------------------ the iteration
struct Solve {                     // device state of step1_process / update_process (double only, like the reference)
    int m, n;
    Buf<double> K, C, a, b, old_a, old_b, dx, dy, p, q, u, v, wa, wb, part;
    Buf<int> flags;                // [0] iteration that last exceeded tau, [1] K was rebuilt at least once
    ColPlan cp;
    int it = 0;                    // iteration counter of this call (the tau flag stores it)

    Solve(const double* hK, const double* hC, const double* ha, const double* hb, const double* hoa, const double* hob,
          const double* hdx, const double* hdy, const double* hp, const double* hq, const double* hu, const double* hv, int m_,
          int n_)
        : m(m_), n(n_) {
        const size_t mn = (size_t)m * (size_t)n;
        uploaded(K, hK, mn); uploaded(C, hC, mn);
        uploaded(a, ha, (size_t)m); uploaded(b, hb, (size_t)n); uploaded(old_a, hoa, (size_t)m); uploaded(old_b, hob, (size_t)n);
        uploaded(dx, hdx, (size_t)m); uploaded(dy, hdy, (size_t)n); uploaded(p, hp, (size_t)m); uploaded(q, hq, (size_t)n);
        uploaded(u, hu, (size_t)m); uploaded(v, hv, (size_t)n);
        wa.alloc((size_t)m); wb.alloc((size_t)n);
        cp = col_plan(m, n);
        part.alloc((size_t)cp.chunks * n);
        flags.alloc(2);
        flags.zero();
        k_mul<double><<<(n + kThreads - 1) / kThreads, kThreads>>>(b.d, dy.d, wb.d, n);      // ot_func.cpp:610-612
        LAUNCHED();
    }

    // step1_process (ot_func.cpp:690-828)
    int step1(int cur_iter, int max_iter, int iters, double tau, double lambda1, double lambda2, double alpha1, double alpha2,
              double eps) {
        const int vec_blocks = (std::max(m, n) + kThreads - 1) / kThreads;
        for (int k = 0; k < iters; ++k) {
            cur_iter += 1;
            ++it;
            k_row_update<double><<<row_warp_blocks(m), kThreads>>>(K.d, wb.d, p.d, u.d, dx.d, a.d, old_a.d, wa.d, alpha1,
                                                                   lambda1 + eps, tau, flags.d, it, m, n);
            LAUNCHED();
            k_col_partial<double><<<dim3(cp.tiles, cp.chunks), kThreads>>>(K.d, wa.d, part.d, m, n, cp.rows_per_chunk);
            LAUNCHED();
            k_col_finish<double><<<cp.tiles, kThreads>>>(part.d, cp.chunks, q.d, v.d, dy.d, b.d, old_b.d, wb.d, alpha2,
                                                         lambda2 + eps, tau, flags.d, it, n);
            LAUNCHED();
            k_absorb_vec<double><<<vec_blocks, kThreads>>>(a.d, b.d, u.d, v.d, wa.d, wb.d, dx.d, dy.d, eps, m, n, flags.d, it,
                                                           flags.d + 1);
            LAUNCHED();
            k_update_k<double><<<grid2d(m, n), kThreads>>>(K.d, nullptr, C.d, u.d, v.d, eps, m, n, flags.d, it);
            LAUNCHED();
            if (cur_iter >= max_iter) {
                printf("Reached max_iter with duality gap still above threshold. Returning");     // ot_func.cpp:822
                return -1;
            }
        }
        return cur_iter;
    }

    void download(double* ha, double* hb, double* hoa, double* hob, double* hK, double* hu, double* hv) {
        int h[2];
        CK(cudaMemcpy(h, flags.d, sizeof(h), cudaMemcpyDeviceToHost));
        a.down(ha); b.down(hb); old_a.down(hoa); old_b.down(hob); u.down(hu); v.down(hv);
        if (h[1]) K.down(hK);          // K only changes when an absorption rebuilt it
    }
};

}  // namespace

// ================================================================================================ exports
extern "C" {

int libot_b200_version(void) { return 100; }

int libot_b200_device_check(void) {
    std::lock_guard<std::mutex> lock(g_mu);
    return device_check();
}

void libot_b200_release(void) {
    std::lock_guard<std::mutex> lock(g_mu);
    pool_clear();
}

void libot_b200_counters(long long* launches, long long* h2d_bytes, long long* d2h_bytes) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (launches) *launches = g_launches;
    if (h2d_bytes) *h2d_bytes = g_h2d;
    if (d2h_bytes) *d2h_bytes = g_d2h;
}

float dummy_float(float*, float*, float*, float*, float*, float*, float*, float*, float*, float, float, float, int, int) { return 0; }
double dummy_double(double*, double*, double*, double*, double*, double*, double*, double*, double*, double, double, double, int,
                    int) { return 0; }

float primal_float(float* C, float* K, float* R, float* dx, float* dy, float* p, float* q, float* a, float* b, float epsilon,
                   float lambda1, float lambda2, int m, int n) {
    return gap_from_host<float>(C, K, R, dx, dy, p, q, a, b, epsilon, lambda1, lambda2, m, n).pri;
}
double primal_double(double* C, double* K, double* R, double* dx, double* dy, double* p, double* q, double* a, double* b,
                     double epsilon, double lambda1, double lambda2, int m, int n) {
    return gap_from_host<double>(C, K, R, dx, dy, p, q, a, b, epsilon, lambda1, lambda2, m, n).pri;
}
float dual_float(float* C, float* K, float* R, float* dx, float* dy, float* p, float* q, float* a, float* b, float epsilon,
                 float lambda1, float lambda2, int m, int n) {
    return gap_from_host<float>(C, K, R, dx, dy, p, q, a, b, epsilon, lambda1, lambda2, m, n).dua;
}
double dual_double(double* C, double* K, double* R, double* dx, double* dy, double* p, double* q, double* a, double* b,
                   double epsilon, double lambda1, double lambda2, int m, int n) {
    return gap_from_host<double>(C, K, R, dx, dy, p, q, a, b, epsilon, lambda1, lambda2, m, n).dua;
}
float compute_duality_gap_float(float* C, float* K, float* R, float* dx, float* dy, float* p, float* q, float* a, float* b,
                                float epsilon, float lambda1, float lambda2, int m, int n) {
    GapParts<float> g = gap_from_host<float>(C, K, R, dx, dy, p, q, a, b, epsilon, lambda1, lambda2, m, n);
    return (g.pri - g.dua) / std::abs(g.pri);                                                     // ot_func.cpp:543
}
double compute_duality_gap_double(double* C, double* K, double* R, double* dx, double* dy, double* p, double* q, double* a,
                                  double* b, double epsilon, double lambda1, double lambda2, int m, int n) {
    GapParts<double> g = gap_from_host<double>(C, K, R, dx, dy, p, q, a, b, epsilon, lambda1, lambda2, m, n);
    return (g.pri - g.dua) / std::abs(g.pri);
}

void update_k_float(float* K, float* K_, float* C, float* u, float* v, float epsilon, int m, int n) {
    update_k_host<float>(K, K_, C, u, v, epsilon, m, n);
}
void update_k_double(double* K, double* K_, double* C, double* u, double* v, double epsilon, int m, int n) {
    update_k_host<double>(K, K_, C, u, v, epsilon, m, n);
}
void update_R_float(float* R, float* K, float* a, float* b, int m, int n) { update_R_host<float>(R, K, a, b, m, n); }
void update_R_double(double* R, double* K, double* a, double* b, int m, int n) { update_R_host<double>(R, K, a, b, m, n); }

int step1_process_double(double* a, double* b, double* old_a, double* old_b, double* K, double* C, double* dx, double* dy,
                         double* p, double* q, double* u, double* v, int cur_iter, int max_iter, int iters, double tau,
                         double lambda1, double lambda2, double alpha1, double alpha2, double epsilon, int m, int n) {
    std::lock_guard<std::mutex> lock(g_mu);
    TraceScope ts("step1_process");
    require_device();
    if (m <= 0 || n <= 0 || iters <= 0) return cur_iter;
    Solve s(K, C, a, b, old_a, old_b, dx, dy, p, q, u, v, m, n);
    const int r = s.step1(cur_iter, max_iter, iters, tau, lambda1, lambda2, alpha1, alpha2, epsilon);
    s.download(a, b, old_a, old_b, K, u, v);
    return r;
}

double update_process_double(double* R, double* a, double* b, double* old_a, double* old_b, double* K, double* _K, double* C,
                             double* dx, double* dy, double* p, double* q, double* u, double* v, int epsilon_scalings,
                             int cur_epsilon_scaling, int batch_size, double epsilon, double threshold, double tau,
                             double lambda1, double lambda2, double alpha1, double alpha2, int cur_iter, int max_iter, int m,
                             int n) {
    std::lock_guard<std::mutex> lock(g_mu);
    TraceScope ts("update_process");
    require_device();
    double gap = 1e100;                                                                           // ot_func.cpp:861
    if (m <= 0 || n <= 0) return gap;
    const bool final_stage = cur_epsilon_scaling == epsilon_scalings;
    const size_t mn = (size_t)m * (size_t)n;
    Solve s(K, C, a, b, old_a, old_b, dx, dy, p, q, u, v, m, n);
    Buf<double> dK_, dR, crit;
    if (final_stage) {                  // _K and R are only touched by the duality gap of the last stage (:887-895)
        uploaded(dK_, _K, mn);
        dR.alloc(mn);
    } else {
        crit.alloc((size_t)kRedBlocks * 4);
    }
    bool r_written = false;
    while (gap > threshold) {
        const int iters = final_stage ? batch_size : 5;                                           // ot_func.cpp:867
        cur_iter = s.step1(cur_iter, max_iter, iters, tau, lambda1, lambda2, alpha1, alpha2, epsilon);
        if (final_stage) {
            GapParts<double> g = gap_device<double, true>(s.K.d, dK_.d, s.C.d, dR.d, s.a.d, s.b.d, s.u.d, s.v.d, s.dx.d, s.dy.d,
                                                          s.p.d, s.q.d, epsilon, lambda1, lambda2, m, n);
            r_written = true;
            gap = (g.pri - g.dua) / std::abs(g.pri);
        } else {
            k_criterion<double><<<kRedBlocks, kThreads>>>(s.a.d, s.old_a.d, s.u.d, m, s.b.d, s.old_b.d, s.v.d, n, epsilon, crit.d);
            LAUNCHED();
            double c[4];
            host_block_sums(crit, kRedBlocks, 4, c);
            const double v1 = std::sqrt(c[0]) / (1 + std::sqrt(c[1]));
            const double v2 = std::sqrt(c[2]) / (1 + std::sqrt(c[3]));
            gap = std::max(v1, v2);                                                               // ot_func.cpp:922
        }
    }
    s.download(a, b, old_a, old_b, K, u, v);
    if (r_written) dR.down(R);
    return gap;
}

}  // extern "C"
