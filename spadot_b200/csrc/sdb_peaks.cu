// Pipe micro-benchmarks: the measured denominators of the roofline of the streamed Sinkhorn pass.
//
// SURVEY.md §8(d) bounds K3 by the SFU (ex2) and FP32 pipes and asks for the nominal figures
// (148 SM x 16 MUFU lanes x clock; 148 x 128 FMA lanes x 2 x clock) "to be confirmed by a
// micro-benchmark on the box": these kernels are that confirmation.  Each thread runs `iters` rounds of
// 8 independent dependency chains of one instruction kind, so with >= 8 resident warps per
// sub-partition the pipe under test is the only limiter; the caller times the launch with CUDA events
// and divides `ops` (returned through *ops_out on the host) by the duration.
#include "sdb_common.cuh"

namespace {

enum { PIPE_MUFU_EX2 = 0, PIPE_FFMA = 1, PIPE_FFMA2 = 2, PIPE_MUFU_FFMA2 = 3 };

template <int KIND>
__global__ void __launch_bounds__(512, 2) pipe_peak_kernel(int iters, float seed, float* __restrict__ out) {
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = seed + 0.001f * (float)(threadIdx.x + k);
    if constexpr (KIND == PIPE_MUFU_EX2) {
        // x <- 2^(-x): contracts to the fixed point 0.641, never leaves [0.5, 1]; the negation is a source modifier
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] = sdb_ex2(-v[k]);
        }
    } else if constexpr (KIND == PIPE_FFMA) {
        const float a = 0.999f, b = 1e-3f;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] = fmaf(v[k], a, b);
        }
    } else if constexpr (KIND == PIPE_FFMA2) {
        uint64_t w[4];
        uint64_t a2, b2;
        asm("mov.b64 %0, {%1, %1};" : "=l"(a2) : "f"(0.999f));
        asm("mov.b64 %0, {%1, %1};" : "=l"(b2) : "f"(1e-3f));
#pragma unroll
        for (int k = 0; k < 4; ++k) asm("mov.b64 %0, {%1, %2};" : "=l"(w[k]) : "f"(v[2 * k]), "f"(v[2 * k + 1]));
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int k = 0; k < 4; ++k) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(w[k]) : "l"(a2), "l"(b2));
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) asm("mov.b64 {%0, %1}, %2;" : "=f"(v[2 * k]), "=f"(v[2 * k + 1]) : "l"(w[k]));
    } else {
        // the instruction mix of the predicted-stabiliser epilogue per two pairs: FFMA2, FADD2, 2 x MUFU.EX2, FADD2
        uint64_t acc[4];
        uint64_t a2, nm2;
        asm("mov.b64 %0, {%1, %1};" : "=l"(a2) : "f"(0.5f));
        asm("mov.b64 %0, {%1, %1};" : "=l"(nm2) : "f"(-0.25f));
#pragma unroll
        for (int k = 0; k < 4; ++k) asm("mov.b64 %0, {%1, %1};" : "=l"(acc[k]) : "f"(0.f));
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    uint64_t t2, x2;
                    asm("mov.b64 %0, {%1, %2};" : "=l"(x2) : "f"(v[2 * k]), "f"(v[2 * k + 1]));
                    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(t2) : "l"(x2), "l"(a2), "l"(nm2));
                    asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(t2) : "l"(nm2));
                    float t0, t1;
                    asm("mov.b64 {%0, %1}, %2;" : "=f"(t0), "=f"(t1) : "l"(t2));
                    v[2 * k] = sdb_ex2(t0);
                    v[2 * k + 1] = sdb_ex2(t1);
                    asm("mov.b64 %0, {%1, %2};" : "=l"(x2) : "f"(v[2 * k]), "f"(v[2 * k + 1]));
                    asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(acc[k]) : "l"(x2));
                }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float s0, s1;
            asm("mov.b64 {%0, %1}, %2;" : "=f"(s0), "=f"(s1) : "l"(acc[k]));
            v[2 * k] += s0;
            v[2 * k + 1] += s1;
        }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += v[k];
    if (s == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;      // keeps the chains alive, never true
}

}  // namespace

extern "C" int sdb_pipe_peak(int kind, int n_ctas, int iters, float* out, double* ops_out_host, void* stream) {
    SDB_CHECK_ARG(out && n_ctas > 0 && iters > 0 && kind >= 0 && kind <= 3);
    const dim3 grid(n_ctas), block(512);
    cudaStream_t st = sdb_stream(stream);
    // operations of the kind under test per launch (ex2 evaluations, or fp32 FMAs = 2 flop each)
    const double threads = (double)n_ctas * 512.0;
    switch (kind) {
        case PIPE_MUFU_EX2: pipe_peak_kernel<PIPE_MUFU_EX2><<<grid, block, 0, st>>>(iters, 0.6f, out); if (ops_out_host) *ops_out_host = threads * iters * 32.0; break;
        case PIPE_FFMA: pipe_peak_kernel<PIPE_FFMA><<<grid, block, 0, st>>>(iters, 0.6f, out); if (ops_out_host) *ops_out_host = threads * iters * 32.0; break;
        case PIPE_FFMA2: pipe_peak_kernel<PIPE_FFMA2><<<grid, block, 0, st>>>(iters, 0.6f, out); if (ops_out_host) *ops_out_host = threads * iters * 64.0; break;
        default: pipe_peak_kernel<PIPE_MUFU_FFMA2><<<grid, block, 0, st>>>(iters, 0.6f, out); if (ops_out_host) *ops_out_host = threads * iters * 32.0; break;
    }
    SDB_LAUNCH_STATUS();
}
