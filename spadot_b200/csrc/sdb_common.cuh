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
