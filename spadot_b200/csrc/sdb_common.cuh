// Shared helpers for libspadot_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/spadot_b200.h"

#define SDB_LOG2E 1.4426950408889634074
#define SDB_LN2   0.6931471805599453094

#define SDB_CHECK_ARG(cond) do { if (!(cond)) return SDB_E_INVALID; } while (0)
#define SDB_LAUNCH_STATUS() do { cudaError_t e__ = cudaGetLastError(); return (int)e__; } while (0)

static inline cudaStream_t sdb_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Programmatic dependent launch (PDL).  Inside sdb_sinkhorn_sweeps the kernels of one iteration are launched with
// the programmatic-stream-serialization attribute: each kernel signals `launch_dependents` as soon as it is set up
// and executes `griddepcontrol.wait` before it touches anything its predecessor writes, so launch latency, barrier /
// TMEM set-up and the first operand tiles of a pass overlap the tail of the previous kernel.  Every kernel of the
// chain waits in every thread, which makes completion transitive along the stream.  Outside that loop sdb_pdl is 0,
// launches are ordinary and both instructions are no-ops.
extern thread_local int sdb_pdl;
__device__ __forceinline__ void sdb_grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void sdb_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t sdb_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = sdb_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// Opt-in to more than 48 KB of dynamic shared memory.  The attribute is per function AND per device, so the "already set"
// memo is indexed by the current device (a process that drives several GPUs sets it on each); a racing second thread at
// worst sets the same value again.
#define SDB_MAX_DEVICES 64
template <typename F>
static inline cudaError_t sdb_ensure_smem(F kern, size_t smem, size_t (&memo)[SDB_MAX_DEVICES]) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= SDB_MAX_DEVICES) return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (smem > memo[dev]) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        memo[dev] = smem;
    }
    return cudaSuccess;
}

__device__ __forceinline__ float sdb_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ double sdb_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic K-wide fp64 grid reduction finished by the last block to arrive.
// scratch: SDB_REDUCE_WIDTH*SDB_REDUCE_BLOCKS doubles of per-block partials followed by one unsigned counter
// at a FIXED offset shared by every K (zero before first use; the kernel leaves it zero again), so reductions of
// different widths can share one scratch buffer on a stream.  gridDim.x <= SDB_REDUCE_BLOCKS, blockDim.x == 256.
template <int K>
__device__ void sdb_grid_reduce(double (&vals)[K], void* scratch_v, double* out) {
    double* scratch = reinterpret_cast<double*>(scratch_v);
    static_assert(K <= SDB_REDUCE_WIDTH, "reduction wider than the scratch layout");
    unsigned* counter = reinterpret_cast<unsigned*>(scratch + SDB_REDUCE_WIDTH * SDB_REDUCE_BLOCKS);
    __shared__ double sm[8][K];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double v = sdb_warp_sum(vals[k]);
        if (lane == 0) sm[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < K) {
        double v = 0.0;
        for (int w = 0; w < 8; ++w) v += sm[w][threadIdx.x];
        scratch[(size_t)blockIdx.x * K + threadIdx.x] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned prev = atomicAdd(counter, 1u);
        is_last = (prev == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = 0.0;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] += scratch[(size_t)b * K + k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double v = sdb_warp_sum(acc[k]);
        if (lane == 0) sm[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < K) {
        double v = 0.0;
        for (int w = 0; w < 8; ++w) v += sm[w][threadIdx.x];
        out[threadIdx.x] = v;
    }
    if (threadIdx.x == 0) *counter = 0u;
}

static inline int sdb_reduce_grid(int64_t n) {
    int64_t b = (n + 255) / 256;
    if (b < 1) b = 1;
    if (b > SDB_REDUCE_BLOCKS) b = SDB_REDUCE_BLOCKS;
    return (int)b;
}

// ---------------------------------------------------------------------------------------------------------------
// One row of the Sinkhorn potential update, shared by the vector kernels (sdb_vectors.cu) and the persistent
// small-problem kernel (sdb_pairs.cu).  No __restrict__ / read-only loads here: inside the persistent kernel these
// buffers are rewritten by other CTAs between grid barriers.
// Predicted stabiliser: a pass may be given, per row, an upper bound m_i of its largest exponent
// instead of tracking the running maximum; the partials then carry (m_i, sum 2^(t - m_i)).  After combining, the
// prediction for the NEXT pass over the same rows is  M + log2 S + 1  (>= this pass's largest exponent), and the pass was
// sound iff  2^-60 < S < 2^100: no term overflowed and the dominant terms were not flushed to zero.
#define SDB_PRED_S_MIN 8.673617379884035e-19     /* 2^-60  */
#define SDB_PRED_S_MAX 1.2676506002282294e30     /* 2^100 */
__device__ __forceinline__ void sdb_prediction_out(double M, double S, int64_t i, float* m_next, int* bad_flag) {
    if (m_next) m_next[i] = (M > -INFINITY && S > 0.0) ? (float)(M + log2(S)) + 1.0f : 0.0f;
    if (bad_flag && M > -INFINITY && !(S > SDB_PRED_S_MIN && S < SDB_PRED_S_MAX)) atomicOr(bad_flag, 1);
}

__device__ __forceinline__ double sdb_combine_partials(const float2* partial, int n_splits, int64_t n, int64_t i, double norm_c1,
                                                       float* m_next = nullptr, int* bad_flag = nullptr) {
    double M = -INFINITY;
    if (n_splits <= 8) {
        // all partials of the row in flight at once (one L2 round trip instead of 2*n_splits dependent ones: this combine sits on
        // the serial tail of every fused pass + update launch)
        float2 pv[8];
#pragma unroll
        for (int s = 0; s < 8; ++s) pv[s] = (s < n_splits) ? __ldcg(partial + (int64_t)s * n + i) : make_float2(SDB_NEG_SENTINEL, 0.f);
        bool any = false;
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            any |= (s < n_splits) && (pv[s].x > -1e29f);
            if (pv[s].x > -1e29f && pv[s].y > 0.f) M = fmax(M, (double)pv[s].x);
        }
        if (!(M > -INFINITY)) {
            if (m_next) m_next[i] = 0.0f;
            if (bad_flag && any) atomicOr(bad_flag, 1);
            return -INFINITY;
        }
        double S = 0.0;
#pragma unroll
        for (int s = 0; s < 8; ++s)
            if (pv[s].x > -1e29f && pv[s].y > 0.f) S += (double)pv[s].y * exp2((double)pv[s].x - M);
        sdb_prediction_out(M, S, i, m_next, bad_flag);
        return SDB_LN2 * (M + log2(S)) - norm_c1;
    }
    for (int s = 0; s < n_splits; ++s) {
        const float2 ps = __ldcg(partial + (int64_t)s * n + i);      // L2: written by other CTAs (possibly of this very kernel)
        if (ps.x > -1e29f && ps.y > 0.f) M = fmax(M, (double)ps.x);
    }
    if (!(M > -INFINITY)) {
        // nothing valid: either every column is masked, or (predicted pass) every term was flushed to zero
        bool any = false;
        for (int s = 0; s < n_splits; ++s) any |= (__ldcg(partial + (int64_t)s * n + i).x > -1e29f);
        if (m_next) m_next[i] = 0.0f;
        if (bad_flag && any) atomicOr(bad_flag, 1);
        return -INFINITY;
    }
    double S = 0.0;
    for (int s = 0; s < n_splits; ++s) {
        const float2 ps = __ldcg(partial + (int64_t)s * n + i);
        if (ps.x > -1e29f && ps.y > 0.f) S += (double)ps.y * exp2((double)ps.x - M);
    }
    sdb_prediction_out(M, S, i, m_next, bad_flag);
    return SDB_LN2 * (M + log2(S)) - norm_c1;
}

// The same combination by one warp per row: lanes take splits lane, lane+32, ... so all partials of a row are in
// flight at once (one L2 round trip instead of 2*n_splits dependent ones); every lane returns the result.
__device__ __forceinline__ double sdb_combine_partials_warp(const float2* partial, int n_splits, int64_t n, int64_t i, double norm_c1) {
    const int lane = threadIdx.x & 31;
    // pass 1: the row maximum (exact in fp32: the partial maxima are floats), loads of all splits in flight at once
    float2 first = make_float2(SDB_NEG_SENTINEL, 0.f);
    float mx = -INFINITY;
    for (int s = lane; s < n_splits; s += 32) {
        const float2 ps = __ldcg(partial + (int64_t)s * n + i);
        if (s == lane) first = ps;
        if (ps.x > -1e29f && ps.y > 0.f) mx = fmaxf(mx, ps.x);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (!(mx > -INFINITY)) return -INFINITY;
    // pass 2: one fp64 exp2 per (lane, split) against the common maximum, then a plain fp64 warp sum
    const double M = (double)mx;
    double S = 0.0;
    for (int s = lane; s < n_splits; s += 32) {
        const float2 ps = (s == lane) ? first : __ldcg(partial + (int64_t)s * n + i);
        if (ps.x > -1e29f && ps.y > 0.f) S += (double)ps.y * exp2((double)ps.x - M);
    }
    S = sdb_warp_sum(S);
    return SDB_LN2 * (M + log2(S)) - norm_c1;
}

// pot_i <- eps*alpha*(logmarg_i - LA_i) with la_old, tau flag and next-pass bias (see sdb_potential_update in the header)
__device__ __forceinline__ void sdb_update_row(int64_t i, double Li, double logmarg_i, double norm_i, double eps, double alpha,
                                               double log_n_other, double c1, double* pot, const double* frame, double* la_old,
                                               float* bias, int* absorb_flag, int iter, double log_tau, double log_floor) {
    const double fr = frame ? frame[i] : 0.0;
    if (la_old) la_old[i] = (pot[i] - fr) / eps;
    double LA = Li - log_n_other;
    if (log_floor > -INFINITY) {   // K(b dy) + 1e-10 of ot_solvers.py:501-502 in total potentials
        const double t = log_floor - fr / eps;
        const double hi = fmax(LA, t), lo = fmin(LA, t);
        LA = (lo == -INFINITY) ? hi : hi + log1p(exp(lo - hi));
    }
    const double nv = eps * alpha * (logmarg_i - LA);
    pot[i] = nv;
    if (bias) {
        const double b = SDB_LOG2E * (nv / eps - norm_i * c1);
        bias[i] = (b > (double)SDB_NEG_SENTINEL) ? (float)b : SDB_NEG_SENTINEL;  // NaN -> sentinel too
    }
    if (absorb_flag && (nv - fr) / eps > log_tau) atomicMax(absorb_flag, iter);
}

// The same update with the tau bookkeeping deferred instead of a separate absorb launch: updaters of tick t stamp
// flag2[t & 1]; an updater of tick t+1 that finds flag2[t & 1] == t first absorbs its own row (frame <- potential), which is
// all `absorb` is in total potentials (ref: ot_func.cpp:792-819).  Whoever reads the frames next (stopping rules) must flush
// the last tick's pending absorption first (sdb_absorb_pending).
__device__ __forceinline__ void sdb_update_row_deferred(int64_t i, double Li, double logmarg_i, double norm_i, double eps, double alpha,
                                                        double log_n_other, double c1, double* pot, double* frame, double* la_old,
                                                        float* bias, int* flag2, int iter, double log_tau, double log_floor) {
    const bool pending = (*reinterpret_cast<volatile int*>(flag2 + ((iter - 1) & 1)) == iter - 1);
    const double old = pot[i];
    double fr = frame[i];
    if (pending) { fr = old; frame[i] = old; }
    la_old[i] = (old - fr) / eps;
    double LA = Li - log_n_other;
    if (log_floor > -INFINITY) {
        const double t = log_floor - fr / eps;
        const double hi = fmax(LA, t), lo = fmin(LA, t);
        LA = (lo == -INFINITY) ? hi : hi + log1p(exp(lo - hi));
    }
    const double nv = eps * alpha * (logmarg_i - LA);
    pot[i] = nv;
    if (bias) {
        const double b = SDB_LOG2E * (nv / eps - norm_i * c1);
        bias[i] = (b > (double)SDB_NEG_SENTINEL) ? (float)b : SDB_NEG_SENTINEL;
    }
    if ((nv - fr) / eps > log_tau) atomicMax(flag2 + (iter & 1), iter);
}
