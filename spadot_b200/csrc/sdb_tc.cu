// K3, tensor-core form: streamed "flash-Sinkhorn" half-iteration on tcgen05 / TMEM / TMA (sm_100a).
//
//   L2_i = log2 sum_j 2^( bias_j + scale * x_i . y_j )
//
// A persistent, warp-specialised CTA per SM owns 128 rows at a time and streams 256-column tiles:
//   warp 0      TMA producer  : cp.async.bulk.tensor tiles of the fp16 hi/lo split embeddings (SWIZZLE_32/64/128B)
//                               and the 256-float bias slice, into mbarrier rings
//   warp 1      MMA issuer    : x.y^T as three tcgen05.mma chains (hi*hi + hi*lo + lo*hi, fp32 accumulate) into one
//                               of two 128x256 fp32 TMEM accumulators
//   warps 2..9  epilogue      : all 8 warps drain each accumulator (32 TMEM lanes x 128 columns per warp) while the
//                               MMA of the next tile fills the other one: tcgen05.ld -> t = scale*d + bias_j ->
//                               lazily stabilised sum of ex2 per row in registers; one (max,sum) pair per row leaves the SM.
// The N x M cost / kernel matrix never exists in any memory.  The 2-term fp16 split keeps 22 mantissa
// bits of every coordinate, so the tile math is fp32-accurate (plain TF32/BF16 is not).
// Replaces gemv/gemtv + update_k of ref: SpaDOT/utils/OT_loss/ot_func.cpp:43-249,547-568.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "sdb_common.cuh"

#include <type_traits>

namespace {

constexpr int TILE_M = 128;          // rows per CTA item (UMMA M)
constexpr int TILE_N = 256;          // columns per streamed tile (UMMA N)
constexpr int BIAS_STAGES = 4;
constexpr int MAX_EPI_WARPS = 16;
constexpr int SWEEP_BINS = 4096;     // shared-memory histogram of the median sweeps
enum { MODE_LSE = 0, MODE_HIST = 1, MODE_COLLECT = 2 };

// ------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
template <int W>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t (&r)[W]) {
    if constexpr (W == 32) tmem_ld32(taddr, r); else tmem_ld16(taddr, r);
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
// packed fp32x2 arithmetic (sm_100: FFMA2 / FADD2 issue one instruction for two lanes of work)
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

// 2^x for a packed pair on the FMA/ALU pipes (no MUFU): Cody-Waite split x = n + f, |f| <= 0.5, degree-5 minimax
// polynomial for 2^f (max rel. error 2.4e-7 in fp32, the same class as ex2.approx's 2^-22), then the integer n is
// added into the exponent field.  Used for a fixed fraction of the exponentials so that the SFU (16 ex2/clk/SM)
// and the FMA pipe share the load.  x must be <= 127; values below -126 are clamped (they are 0 in the sum).
__device__ __forceinline__ uint64_t exp2_poly2(uint64_t x2) {
    float x0, x1;
    unpack2(x2, x0, x1);
    x0 = fminf(fmaxf(x0, -126.f), 127.f);                          // (an exponent above 127 must stay huge, not wrap: the
    x1 = fminf(fmaxf(x1, -126.f), 127.f);                          //  predicted-stabiliser range check relies on it)
    x2 = pack2(x0, x1);
    const uint64_t magic = pack2(12582912.f, 12582912.f);        // 1.5 * 2^23: rounds to nearest integer
    const uint64_t r2 = fadd2(x2, magic);
    const uint64_t n2 = fadd2(r2, pack2(-12582912.f, -12582912.f));
    const uint64_t f2 = ffma2(n2, pack2(-1.f, -1.f), x2);
    uint64_t p2 = pack2(0.0013390863314270973f, 0.0013390863314270973f);
    p2 = ffma2(p2, f2, pack2(0.009676031768321991f, 0.009676031768321991f));
    p2 = ffma2(p2, f2, pack2(0.055503569543361664f, 0.055503569543361664f));
    p2 = ffma2(p2, f2, pack2(0.2402210682630539f, 0.2402210682630539f));
    p2 = ffma2(p2, f2, pack2(0.6931471824645996f, 0.6931471824645996f));
    p2 = ffma2(p2, f2, pack2(1.0000001192092896f, 1.0000001192092896f));
    float r0, r1, p0, p1;
    unpack2(r2, r0, r1);
    unpack2(p2, p0, p1);
    const float e0 = __int_as_float(__float_as_int(p0) + (__float_as_int(r0) << 23));
    const float e1 = __int_as_float(__float_as_int(p1) + (__float_as_int(r1) << 23));
    return pack2(e0, e1);
}

// K-major, swizzled UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor layout, version 1):
//   [0,14) start>>4, [16,30) LBO>>4 (=1: unused for swizzled K-major), [32,46) SBO>>4 (8 rows * row bytes),
//   [46,48) version=1, [61,64) layout: 2=SW128, 4=SW64, 6=SW32.
template <int DP>
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    constexpr uint64_t row_bytes = DP * 2;
    constexpr uint64_t layout = (DP == 64) ? 2 : (DP == 32) ? 4 : 6;
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((8 * row_bytes) >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= layout << 61;
    return d;
}

template <int DP>
struct TcSmem {
    static constexpr int STAGES = (DP == 64) ? 2 : 4;
    static constexpr int A_BYTES = TILE_M * DP * 2;     // one of (hi, lo)
    static constexpr int B_BYTES = TILE_N * DP * 2;
    static constexpr int A_STAGES = (DP == 64) ? 1 : 2;                 // row-tile operand double buffer: the next item's rows
                                                                        // load while the current item's last tiles still multiply
    static constexpr int OFF_A = 0;                                     // [A_STAGES][2][A_BYTES]
    static constexpr int OFF_B = OFF_A + A_STAGES * 2 * A_BYTES;        // [STAGES][2][B_BYTES]
    static constexpr int OFF_BIAS = OFF_B + STAGES * 2 * B_BYTES;       // [BIAS_STAGES][TILE_N] float
    static constexpr int OFF_MERGE = OFF_BIAS + BIAS_STAGES * TILE_N * 4;  // [3][TILE_M] float2
    static constexpr int OFF_BAR = OFF_MERGE + 3 * TILE_M * 8;          // barriers
    static constexpr int N_BARS = 2 * STAGES + 2 * A_STAGES + 4 + 2 * BIAS_STAGES;
    static constexpr int OFF_TMEM = OFF_BAR + N_BARS * 8;
    static constexpr int OFF_HIST = OFF_TMEM + 16;                      // [SWEEP_BINS] uint32 (median sweeps only)
    static constexpr int TOTAL = OFF_HIST + 1024;                       // + alignment slack (LSE pass)
    static constexpr int OFF_SDIST = OFF_HIST + SWEEP_BINS * 4;        // [32][256] float (histogram sweep only)
    static constexpr int TOTAL_SWEEP = OFF_SDIST + 32 * 256 * 4 + 1024;
    static_assert(TOTAL_SWEEP <= 227 * 1024, "shared-memory budget of one CTA");
};

// Rare paths of the median sweeps, kept out of line so the hot loop stays small (an inlined fp64 loop per
// accumulator element blew the instruction cache: the sweep ran 20x slower than the LSE pass).
__device__ __noinline__ void sweep_hist_rare(uint32_t mask, const float* sd, float lo, float inv_width, int n_bins,
                                             uint32_t* shist, unsigned long long* inside) {
    while (mask) {
        const int k = __ffs(mask) - 1;
        mask &= mask - 1;
        int bin = (int)((sd[k * 256] - lo) * inv_width);
        bin = min(max(bin, 0), n_bins - 1);
        atomicAdd(shist + bin, 1u);
        ++(*inside);
    }
}
__device__ __noinline__ void sweep_collect_rare(uint32_t mask, const double* xi, const double* y64, int64_t col0, int d,
                                                double* cand, unsigned long long cap, unsigned long long* counter) {
    while (mask) {
        const int k = __ffs(mask) - 1;
        mask &= mask - 1;
        const double* yj = y64 + (col0 + k) * d;
        double s2 = 0.0;
        for (int q = 0; q < d; ++q) { const double df = xi[q] - yj[q]; s2 = __dadd_rn(s2, __dmul_rn(df, df)); }
        const unsigned long long slot = atomicAdd(counter, 1ull);
        if (slot < cap) cand[slot] = s2;
    }
}

struct TcArgs;
__device__ __forceinline__ void split_range(const TcArgs& a, int sp, int& t0, int& t1);

struct TcArgs {
    sdb_tc_update upd;       // upd.counters != NULL: the CTA completing a row tile also combines and updates it (fused pass + update)
    const int* split_tiles;  // optional: n_splits + 1 tile boundaries of variable-length splits (label groups, K6); else uniform
    const float* bias;       // padded to a multiple of TILE_N with SDB_NEG_SENTINEL
    const float* row_m;      // predicted stabiliser per row (FIXED_M kernels), else unused
    float scale;
    int64_t n_p;
    int n_row_tiles, n_col_tiles, tiles_per_split, n_splits;
    float2* partial;         // [n_splits][n_p]
    // median sweeps (MODE_HIST / MODE_COLLECT): bias carries |y_j|^2 (+huge padding), scale = -2 * 2^(-2*pow2_exp)
    const float* row_norms;  // |x_i|^2 as float, n_p entries
    float lo, hi, inv_width;
    int n_bins;
    unsigned long long* hist;     // [n_bins]
    unsigned long long* counts;   // [2]: below, inside/appended
    const double* x64;            // original coordinates, row-major (n_p, d) / (n_q, d)
    const double* y64;
    int d;
    int64_t n_q;
    double* cand;
    unsigned long long cap;
};

__device__ __forceinline__ void split_range(const TcArgs& a, int sp, int& t0, int& t1) {
    if (a.split_tiles) {
        t0 = __ldg(a.split_tiles + sp);
        t1 = __ldg(a.split_tiles + sp + 1);
    } else {
        t0 = sp * a.tiles_per_split;
        t1 = min(a.n_col_tiles, t0 + a.tiles_per_split);
    }
}

template <int DP, int EPI_WARPS, bool PACKED, int POLY_EVERY = 0, int MODE = MODE_LSE, bool PIPE = false, bool XTILE = false,
          bool FIXED_M = false, bool RUNSUM = false>
__global__ void __launch_bounds__(32 * (2 + EPI_WARPS), 1)
lse_pass_tc_kernel(const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmQ, TcArgs a) {
    using S = TcSmem<DP>;
    constexpr int N_EPI_WARPS = EPI_WARPS;
    constexpr int STAGES = S::STAGES;
    constexpr int KSTEPS = DP / 16;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem + S::OFF_A;
    uint8_t* sB = smem + S::OFF_B;
    float* sBias = reinterpret_cast<float*>(smem + S::OFF_BIAS);
    float2* sMerge = reinterpret_cast<float2*>(smem + S::OFF_MERGE);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
    uint64_t* full = bars;                      // [STAGES]
    uint64_t* empty = full + STAGES;            // [STAGES]
    uint64_t* a_full = empty + STAGES;          // [A_STAGES]
    uint64_t* a_empty = a_full + S::A_STAGES;   // [A_STAGES]
    uint64_t* t_full = a_empty + S::A_STAGES;   // [2]
    uint64_t* t_empty = t_full + 2;             // [2]
    uint64_t* b_full = t_empty + 2;             // [BIAS_STAGES]
    uint64_t* b_empty = b_full + BIAS_STAGES;   // [BIAS_STAGES]
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(smem + S::OFF_TMEM);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
        for (int i = 0; i < S::A_STAGES; ++i) { mbar_init(a_full + i, 1); mbar_init(a_empty + i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(t_full + i, 1); mbar_init(t_empty + i, N_EPI_WARPS); }
        for (int i = 0; i < BIAS_STAGES; ++i) { mbar_init(b_full + i, 1); mbar_init(b_empty + i, N_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmP) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_holder)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    // PDL (sdb_common.cuh): the next kernel of the stream may be scheduled from here on; this kernel's operand tiles
    // and MMAs do not depend on its predecessor, only the bias vector and the output buffer do (waits below).
    sdb_launch_dependents();

    const int n_items = a.n_row_tiles * a.n_splits;

    if (warp == 0) {
        // =============================================================== TMA producer
        if (lane == 0) {
            uint32_t tile_ctr = 0, item_ctr = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_ctr) {
                const int rt = item % a.n_row_tiles, sp = item / a.n_row_tiles;
                int t0, t1;
                split_range(a, sp, t0, t1);
                const int as = item_ctr % S::A_STAGES;
                uint8_t* sAs = sA + (size_t)as * 2 * S::A_BYTES;
                mbar_wait(a_empty + as, ((item_ctr / S::A_STAGES) & 1) ^ 1);
                mbar_expect_tx(a_full + as, 2 * S::A_BYTES);
                tma_load_2d(sAs, &tmP, a_full + as, 0, rt * TILE_M);
                tma_load_2d(sAs + S::A_BYTES, &tmP, a_full + as, DP, rt * TILE_M);
                for (int t = t0; t < t1; ++t, ++tile_ctr) {
                    const int s = tile_ctr % STAGES;
                    mbar_wait(empty + s, ((tile_ctr / STAGES) & 1) ^ 1);
                    mbar_expect_tx(full + s, 2 * S::B_BYTES);
                    uint8_t* dst = sB + (size_t)s * 2 * S::B_BYTES;
                    tma_load_2d(dst, &tmQ, full + s, 0, t * TILE_N);
                    tma_load_2d(dst + S::B_BYTES, &tmQ, full + s, DP, t * TILE_N);
                    const int bs = tile_ctr % BIAS_STAGES;
                    mbar_wait(b_empty + bs, ((tile_ctr / BIAS_STAGES) & 1) ^ 1);
                    if (tile_ctr == 0) sdb_grid_dependency_wait();     // the bias was written by the previous kernel
                    mbar_expect_tx(b_full + bs, TILE_N * 4);
                    bulk_load_1d(sBias + bs * TILE_N, a.bias + (size_t)t * TILE_N, TILE_N * 4, b_full + bs);
                }
            }
        }
    } else if (warp == 1) {
        // =============================================================== MMA issuer
        // instruction descriptor: D=f32 (bit 4), A=B=f16 K-major, N>>3 at [17,23), M>>4 at [24,29)
        constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(TILE_N >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
        uint32_t tile_ctr = 0, item_ctr = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_ctr) {
            const int sp = item / a.n_row_tiles;
            int t0, t1;
            split_range(a, sp, t0, t1);
            const int as = item_ctr % S::A_STAGES;
            const uint64_t dAh = make_desc<DP>(smem_u32(sA + (size_t)as * 2 * S::A_BYTES));
            const uint64_t dAl = make_desc<DP>(smem_u32(sA + (size_t)as * 2 * S::A_BYTES + S::A_BYTES));
            mbar_wait(a_full + as, (item_ctr / S::A_STAGES) & 1);
            for (int t = t0; t < t1; ++t, ++tile_ctr) {
                const int s = tile_ctr % STAGES;
                const int acc = tile_ctr & 1;
                mbar_wait(full + s, (tile_ctr / STAGES) & 1);
                mbar_wait(t_empty + acc, ((tile_ctr >> 1) & 1) ^ 1);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t bsm = smem_u32(sB + (size_t)s * 2 * S::B_BYTES);
                    const uint64_t dBh = make_desc<DP>(bsm);
                    const uint64_t dBl = make_desc<DP>(bsm + S::B_BYTES);
                    const uint32_t d_tmem = tmem_base + acc * TILE_N;
                    // The tensor core adds into its fp32 accumulator with truncation, i.e. every add at full magnitude
                    // biases the dot product toward zero by ~half an ulp.  The two cross chains (2^-11 of the magnitude)
                    // therefore go first, while the accumulator is small, and the hi.hi chain last: two full-magnitude
                    // adds instead of six (measured at 8192^2: plan entries biased +5e-6 before).
#pragma unroll
                    for (int k = 0; k < KSTEPS; ++k) umma_f16(d_tmem, dAh + 2 * k, dBl + 2 * k, idesc, k > 0);
#pragma unroll
                    for (int k = 0; k < KSTEPS; ++k) umma_f16(d_tmem, dAl + 2 * k, dBh + 2 * k, idesc, 1);
#pragma unroll
                    for (int k = 0; k < KSTEPS; ++k) umma_f16(d_tmem, dAh + 2 * k, dBh + 2 * k, idesc, 1);
                    umma_commit(empty + s);
                    umma_commit(t_full + acc);
                    if (t == t1 - 1) umma_commit(a_empty + as);
                }
                __syncwarp();
            }
        }
    } else {
        // =============================================================== epilogue warps
        // Every tile is drained by all epilogue warps: warp (quad, part) owns TMEM lanes [32*quad, +32) and a
        // 256/PARTS-column slice of the current accumulator, so the MMA of tile t+1 (other accumulator) overlaps
        // the epilogue of tile t.  tcgen05.ld of the next chunk is in flight while this one is processed.  The
        // stabiliser m_used only moves when a chunk maximum exceeds it by more than 2^64 (floating point is
        // scale-free: no precision is lost, terms 2^-126 below the stabiliser are irrelevant to the sum).
        constexpr int PARTS = EPI_WARPS / 4;
        constexpr int COLS = TILE_N / PARTS;         // columns per warp per tile
        constexpr int CH = (EPI_WARPS == 8) ? 32 : 16;
        constexpr int NCH = COLS / CH;
        sdb_grid_dependency_wait();        // before anything is written to global memory (partials, histograms)
        const int ew = warp - 2;
        const int part = ew >> 2;
        const int quad = warp & 3;                   // TMEM lane quadrant accessible to this warp
        const int row_in_tile = quad * 32 + lane;
        const float scale = a.scale;
        const uint64_t scale2 = pack2(scale, scale);
        const uint32_t bias_base = smem_u32(sBias) + part * COLS * 4;
        if constexpr (MODE != MODE_LSE) {
            // ----------------------------------------------------------- median sweeps (K5 on the tensor cores)
            // dist = |x_i|^2 + |y_j|^2 - 2 x_i.y_j from the same accumulator tiles; count dist < lo, and for the few
            // pairs inside [lo, hi) either bin them (shared-memory histogram) or re-evaluate them exactly in fp64
            // from the original coordinates and append them to the candidate list.
            uint32_t* shist = reinterpret_cast<uint32_t*>(smem + S::OFF_HIST);
            if constexpr (MODE == MODE_HIST) {
                for (int bI = threadIdx.x - 64; bI < SWEEP_BINS; bI += 32 * EPI_WARPS) shist[bI] = 0u;
                asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
            }
            unsigned long long below = 0ull, inside = 0ull;
            uint32_t tile_ctr = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const int rt = item % a.n_row_tiles, sp = item / a.n_row_tiles;
                int t0, t1;
                split_range(a, sp, t0, t1);
                const int64_t row = (int64_t)rt * TILE_M + row_in_tile;
                // u = |x_i|^2 + |y_j|^2 - 2 x_i.y_j - lo; rows beyond n_p (and padded columns, |y_j|^2 >= 3e38) come out huge
                // and positive: neither below nor inside
                const float nxl = row < a.n_p ? a.row_norms[row] - a.lo : 3.0e38f;
                const float width = a.hi - a.lo;
                for (int t = t0; t < t1; ++t, ++tile_ctr) {
                    const int acc = tile_ctr & 1;
                    const int bs = tile_ctr % BIAS_STAGES;
                    mbar_wait(b_full + bs, (tile_ctr / BIAS_STAGES) & 1);
                    mbar_wait(t_full + acc, (tile_ctr >> 1) & 1);
                    tc_fence_after();
                    const uint32_t bias_s = bias_base + bs * TILE_N * 4;
                    const uint32_t tbase = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * TILE_N + part * COLS;
                    uint32_t d[2][CH];
                    tmem_ld<CH>(tbase, d[0]);
                    uint32_t cnt = 0;
                    // Classification on sign bits: u = dist - lo, v = u - (hi - lo).  below <=> sign(u); inside <=> !sign(u) & sign(v).
                    // Each element shifts its bit into a 32-bit mask with one funnel shift, so the integer pipe sees 3
                    // instructions per pair (the compare / select form it replaces needed ~9 and made the sweeps ALU-bound:
                    // 0.66-0.73 s per sweep at 1M x 1M against 0.29 s for an LSE pass).  Element k of a chunk lands in bit 31-k.
#pragma unroll
                    for (int c = 0; c < NCH; ++c) {
                        tmem_ld_wait();
                        if (c + 1 < NCH) tmem_ld<CH>(tbase + (c + 1) * CH, d[(c + 1) & 1]);
                        float u[CH];
                        uint32_t below_mask = 0, mask = 0;
#pragma unroll
                        for (int k4 = 0; k4 < CH / 4; ++k4) {
                            const float4 b = lds128(bias_s + (c * CH + k4 * 4) * 4);
                            const float ny[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const float uv = fmaf(scale, __uint_as_float(d[c & 1][k4 * 4 + q]), ny[q]) + nxl;
                                u[k4 * 4 + q] = uv;
                                const uint32_t ub = __float_as_uint(uv), vb = __float_as_uint(uv - width);
                                below_mask = __funnelshift_l(ub, below_mask, 1);
                                mask = __funnelshift_l(~ub & vb, mask, 1);
                            }
                        }
                        cnt += __popc(below_mask);
                        if (mask) {
                            mask = __brev(mask);                       // bit k = element k again
                            const int64_t col0 = (int64_t)t * TILE_N + part * COLS + c * CH;
                            if constexpr (MODE == MODE_HIST) {
                                float* sd = reinterpret_cast<float*>(smem + S::OFF_SDIST) + (threadIdx.x - 64);
#pragma unroll
                                for (int k = 0; k < CH; ++k) sd[k * 256] = u[k];
                                sweep_hist_rare(mask, sd, 0.f, a.inv_width, a.n_bins, shist, &inside);
                            } else {
                                sweep_collect_rare(mask, a.x64 + row * a.d, a.y64, col0, a.d, a.cand, a.cap, a.counts + 1);
                            }
                        }
                    }
                    below += cnt;
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) { mbar_arrive(t_empty + acc); mbar_arrive(b_empty + bs); }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                below += __shfl_xor_sync(0xffffffffu, below, o);
                inside += __shfl_xor_sync(0xffffffffu, inside, o);
            }
            if (lane == 0 && below) atomicAdd(a.counts, below);
            if constexpr (MODE == MODE_HIST) {
                if (lane == 0 && inside) atomicAdd(a.counts + 1, inside);
                asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
                for (int bI = threadIdx.x - 64; bI < a.n_bins; bI += 32 * EPI_WARPS)
                    if (shist[bI]) atomicAdd(a.hist + bI, (unsigned long long)shist[bI]);
            }
        } else {
        uint32_t tile_ctr = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int rt = item % a.n_row_tiles, sp = item / a.n_row_tiles;
            int t0, t1;
            split_range(a, sp, t0, t1);
            float m_used = SDB_NEG_SENTINEL, ssum = 0.f;
            if constexpr (FIXED_M) {
                // predicted stabiliser: an upper bound of this row's largest exponent, supplied by the caller
                const int64_t prow = (int64_t)rt * TILE_M + row_in_tile;
                m_used = prow < a.n_p ? __ldg(a.row_m + prow) : 0.f;
            }
            if constexpr (XTILE) {
                // Cross-tile software pipeline.  Same chunk pipeline as PIPE, but it does not drain at tile boundaries: the
                // first two tcgen05.ld of tile t+1 and the FP32 stage of its first chunk are issued inside the last two
                // chunks of tile t, and tile t's accumulator is released as soon as its last TMEM read has landed.  In
                // the PIPE form both warps of a sub-partition hit the boundary together (same barrier) and the SFU
                // idles for one TMEM round trip + one FP32 stage per tile (~13 % of the tile).
                static_assert(NCH == 4, "the cross-tile pipeline is written for four chunks per tile");
                uint32_t d[2][CH];
                uint64_t tv2[2][CH / 2];
                float cmv[2];
                auto stage_a = [&](int c, uint32_t bias_s) {            // FP32 stage: packed fma + chunk maximum
#pragma unroll
                    for (int k4 = 0; k4 < CH / 4; ++k4) {
                        const float4 b = lds128(bias_s + (c * CH + k4 * 4) * 4);
                        tv2[c & 1][k4 * 2 + 0] = ffma2(scale2, pack2(__uint_as_float(d[c & 1][k4 * 4 + 0]), __uint_as_float(d[c & 1][k4 * 4 + 1])), pack2(b.x, b.y));
                        tv2[c & 1][k4 * 2 + 1] = ffma2(scale2, pack2(__uint_as_float(d[c & 1][k4 * 4 + 2]), __uint_as_float(d[c & 1][k4 * 4 + 3])), pack2(b.z, b.w));
                    }
                    float m0 = SDB_NEG_SENTINEL, m1 = SDB_NEG_SENTINEL, m2 = SDB_NEG_SENTINEL, m3 = SDB_NEG_SENTINEL;
#pragma unroll
                    for (int k = 0; k < CH / 2; k += 2) {
                        float a0, a1, a2, a3;
                        unpack2(tv2[c & 1][k], a0, a1);
                        unpack2(tv2[c & 1][k + 1], a2, a3);
                        m0 = fmaxf(m0, a0); m1 = fmaxf(m1, a1); m2 = fmaxf(m2, a2); m3 = fmaxf(m3, a3);
                    }
                    cmv[c & 1] = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
                };
                auto stabilise = [&](int c) {                           // branch-free lazy stabiliser (see PIPE)
                    const float cm = cmv[c & 1];
                    const float m_new = (cm > m_used + 64.f) ? cm : m_used;
                    ssum *= sdb_ex2(m_used - m_new);
                    m_used = m_new;
                };
                auto sfu_stage = [&](int c) {                           // SFU stage of chunk c
                    const uint64_t nm2 = pack2(-m_used, -m_used);
                    uint64_t acc2[CH / 2];
#pragma unroll
                    for (int k = 0; k < CH / 2; ++k) {
                        float a0, a1;
                        unpack2(fadd2(tv2[c & 1][k], nm2), a0, a1);
                        acc2[k] = pack2(sdb_ex2(a0), sdb_ex2(a1));
                    }
#pragma unroll
                    for (int w = CH / 4; w > 0; w >>= 1)
#pragma unroll
                        for (int k = 0; k < w; ++k) acc2[k] = fadd2(acc2[k], acc2[k + w]);
                    float e0, e1;
                    unpack2(acc2[0], e0, e1);
                    ssum += e0 + e1;
                };
                auto tile_tbase = [&](uint32_t ctr) { return tmem_base + ((uint32_t)(quad * 32) << 16) + (ctr & 1) * TILE_N + part * COLS; };
                auto tile_bias = [&](uint32_t ctr) { return bias_base + (ctr % BIAS_STAGES) * TILE_N * 4; };
                // item prologue: first tile's first two loads and its first FP32 stage
                mbar_wait(b_full + tile_ctr % BIAS_STAGES, (tile_ctr / BIAS_STAGES) & 1);
                mbar_wait(t_full + (tile_ctr & 1), (tile_ctr >> 1) & 1);
                tc_fence_after();
                tmem_ld<CH>(tile_tbase(tile_ctr), d[0]);
                tmem_ld_wait();
                tmem_ld<CH>(tile_tbase(tile_ctr) + CH, d[1]);
                stage_a(0, tile_bias(tile_ctr));
                // one tile; HAS_NEXT is a compile-time flag so that the steady-state body has no data-dependent branches
                // (a branch around the next tile's FP32 stage would put it in another basic block than the SFU stage it
                // is meant to overlap)
                auto tile_body = [&](auto has_next_c) {
                    constexpr bool HAS_NEXT = decltype(has_next_c)::value;
                    const int acc = tile_ctr & 1;
                    const int bs = tile_ctr % BIAS_STAGES;
                    const uint32_t bias_s = tile_bias(tile_ctr);
                    const uint32_t tbase = tile_tbase(tile_ctr);
                    const uint32_t nctr = tile_ctr + 1;
                    // chunk 0
                    stabilise(0);
                    tmem_ld_wait();                                   // chunk 1 landed in d[1]
                    tmem_ld<CH>(tbase + 2 * CH, d[0]);
                    stage_a(1, bias_s);
                    sfu_stage(0);
                    // chunk 1
                    stabilise(1);
                    tmem_ld_wait();                                   // chunk 2 landed in d[0]
                    tmem_ld<CH>(tbase + 3 * CH, d[1]);
                    stage_a(2, bias_s);
                    sfu_stage(1);
                    // the next tile's accumulator and bias (long finished: the MMA runs a whole tile ahead)
                    if constexpr (HAS_NEXT) {
                        mbar_wait(b_full + nctr % BIAS_STAGES, (nctr / BIAS_STAGES) & 1);
                        mbar_wait(t_full + (nctr & 1), (nctr >> 1) & 1);
                        tc_fence_after();
                    }
                    // chunk 2
                    stabilise(2);
                    tmem_ld_wait();                                   // chunk 3 landed in d[1]: this accumulator is drained
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(t_empty + acc);
                    if constexpr (HAS_NEXT) tmem_ld<CH>(tile_tbase(nctr), d[0]);
                    stage_a(3, bias_s);
                    sfu_stage(2);
                    // chunk 3
                    stabilise(3);
                    if constexpr (HAS_NEXT) {
                        tmem_ld_wait();                               // next tile's chunk 0 landed in d[0]
                        tmem_ld<CH>(tile_tbase(nctr) + CH, d[1]);
                        stage_a(0, tile_bias(nctr));
                    }
                    sfu_stage(3);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(b_empty + bs);
                    ++tile_ctr;
                };
                for (int t = t0; t < t1 - 1; ++t) tile_body(std::true_type{});
                tile_body(std::false_type{});
            } else if constexpr (PIPE) {
                // Software-pipelined form: the FP32 stage of chunk c+1 (t = scale*d + bias, chunk maximum) and the SFU
                // stage of chunk c (ex2, sum) sit in one basic block with no dependence between them, so the scheduler
                // interleaves FMA-pipe and MUFU work of the same warp instead of running them as alternating phases.
                // RUNSUM (predicted stabiliser only: no rescale ever happens): the exponentials go into four packed running
                // sums that live across chunks and tiles and are folded once per item - no reduction tree, no scalar adds and no
                // end-of-chunk dependency on the last MUFU results (each FADD2 depends only on a pair issued four pairs earlier).
                uint64_t racc[4] = {pack2(0.f, 0.f), pack2(0.f, 0.f), pack2(0.f, 0.f), pack2(0.f, 0.f)};
                for (int t = t0; t < t1; ++t, ++tile_ctr) {
                    const int acc = tile_ctr & 1;
                    const int bs = tile_ctr % BIAS_STAGES;
                    mbar_wait(b_full + bs, (tile_ctr / BIAS_STAGES) & 1);
                    mbar_wait(t_full + acc, (tile_ctr >> 1) & 1);
                    tc_fence_after();
                    const uint32_t bias_s = bias_base + bs * TILE_N * 4;
                    const uint32_t tbase = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * TILE_N + part * COLS;
                    uint32_t d[2][CH];
                    uint64_t tv2[2][CH / 2];
                    float cmv[2];
                    auto stage_a = [&](int c) {            // FP32 stage: packed fma + chunk maximum
#pragma unroll
                        for (int k4 = 0; k4 < CH / 4; ++k4) {
                            const float4 b = lds128(bias_s + (c * CH + k4 * 4) * 4);
                            tv2[c & 1][k4 * 2 + 0] = ffma2(scale2, pack2(__uint_as_float(d[c & 1][k4 * 4 + 0]), __uint_as_float(d[c & 1][k4 * 4 + 1])), pack2(b.x, b.y));
                            tv2[c & 1][k4 * 2 + 1] = ffma2(scale2, pack2(__uint_as_float(d[c & 1][k4 * 4 + 2]), __uint_as_float(d[c & 1][k4 * 4 + 3])), pack2(b.z, b.w));
                        }
                        if constexpr (!FIXED_M) {
                            float m0 = SDB_NEG_SENTINEL, m1 = SDB_NEG_SENTINEL, m2 = SDB_NEG_SENTINEL, m3 = SDB_NEG_SENTINEL;
#pragma unroll
                            for (int k = 0; k < CH / 2; k += 2) {
                                float a0, a1, a2, a3;
                                unpack2(tv2[c & 1][k], a0, a1);
                                unpack2(tv2[c & 1][k + 1], a2, a3);
                                m0 = fmaxf(m0, a0); m1 = fmaxf(m1, a1); m2 = fmaxf(m2, a2); m3 = fmaxf(m3, a3);
                            }
                            cmv[c & 1] = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
                        }
                    };
                    tmem_ld<CH>(tbase, d[0]);
                    tmem_ld_wait();
                    tmem_ld<CH>(tbase + CH, d[1]);
                    stage_a(0);
#pragma unroll
                    for (int c = 0; c < NCH; ++c) {
                        if constexpr (!FIXED_M) {
                            const float cm = cmv[c & 1];
                            const float m_new = (cm > m_used + 64.f) ? cm : m_used;
                            ssum *= sdb_ex2(m_used - m_new);
                            m_used = m_new;
                        }
                        const uint64_t nm2 = pack2(-m_used, -m_used);
                        if (c + 1 < NCH) {
                            tmem_ld_wait();                                   // chunk c+1 has landed in d[(c+1)&1]
                            if (c + 2 < NCH) tmem_ld<CH>(tbase + (c + 2) * CH, d[c & 1]);   // d[c&1] was consumed by stage_a(c)
                            stage_a(c + 1);
                        }
                        if constexpr (FIXED_M && RUNSUM) {
#pragma unroll
                            for (int k = 0; k < CH / 2; ++k) {                 // SFU stage of chunk c, running sums
                                float a0, a1;
                                unpack2(fadd2(tv2[c & 1][k], nm2), a0, a1);
                                racc[k & 3] = fadd2(racc[k & 3], pack2(sdb_ex2(a0), sdb_ex2(a1)));
                            }
                        } else {
                        uint64_t acc2[CH / 2];
#pragma unroll
                        for (int k = 0; k < CH / 2; ++k) {                     // SFU stage of chunk c
                            const uint64_t x2 = fadd2(tv2[c & 1][k], nm2);
                            if (POLY_EVERY > 0 && (k % POLY_EVERY) == POLY_EVERY - 1) {
                                acc2[k] = exp2_poly2(x2);                      // development knob: this pair on the FMA pipe
                            } else {
                                float a0, a1;
                                unpack2(x2, a0, a1);
                                acc2[k] = pack2(sdb_ex2(a0), sdb_ex2(a1));
                            }
                        }
#pragma unroll
                        for (int w = CH / 4; w > 0; w >>= 1)
#pragma unroll
                            for (int k = 0; k < w; ++k) acc2[k] = fadd2(acc2[k], acc2[k + w]);
                        float e0, e1;
                        unpack2(acc2[0], e0, e1);
                        ssum += e0 + e1;
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) { mbar_arrive(t_empty + acc); mbar_arrive(b_empty + bs); }
                }
                if constexpr (FIXED_M && RUNSUM) {
                    float e0, e1;
                    unpack2(fadd2(fadd2(racc[0], racc[1]), fadd2(racc[2], racc[3])), e0, e1);
                    ssum = e0 + e1;
                }
            } else {
            for (int t = t0; t < t1; ++t, ++tile_ctr) {
                const int acc = tile_ctr & 1;
                const int bs = tile_ctr % BIAS_STAGES;
                mbar_wait(b_full + bs, (tile_ctr / BIAS_STAGES) & 1);
                mbar_wait(t_full + acc, (tile_ctr >> 1) & 1);
                tc_fence_after();
                const uint32_t bias_s = bias_base + bs * TILE_N * 4;
                const uint32_t tbase = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * TILE_N + part * COLS;
                uint32_t d[2][CH];
                tmem_ld<CH>(tbase, d[0]);
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    tmem_ld_wait();
                    if (c + 1 < NCH) tmem_ld<CH>(tbase + (c + 1) * CH, d[(c + 1) & 1]);
                    float tv[CH];
#pragma unroll
                    for (int k4 = 0; k4 < CH / 4; ++k4) {
                        const float4 b = lds128(bias_s + (c * CH + k4 * 4) * 4);
                        if constexpr (PACKED) {
                            const uint64_t r0 = ffma2(scale2, pack2(__uint_as_float(d[c & 1][k4 * 4 + 0]), __uint_as_float(d[c & 1][k4 * 4 + 1])), pack2(b.x, b.y));
                            const uint64_t r1 = ffma2(scale2, pack2(__uint_as_float(d[c & 1][k4 * 4 + 2]), __uint_as_float(d[c & 1][k4 * 4 + 3])), pack2(b.z, b.w));
                            unpack2(r0, tv[k4 * 4 + 0], tv[k4 * 4 + 1]);
                            unpack2(r1, tv[k4 * 4 + 2], tv[k4 * 4 + 3]);
                        } else {
                            tv[k4 * 4 + 0] = fmaf(scale, __uint_as_float(d[c & 1][k4 * 4 + 0]), b.x);
                            tv[k4 * 4 + 1] = fmaf(scale, __uint_as_float(d[c & 1][k4 * 4 + 1]), b.y);
                            tv[k4 * 4 + 2] = fmaf(scale, __uint_as_float(d[c & 1][k4 * 4 + 2]), b.z);
                            tv[k4 * 4 + 3] = fmaf(scale, __uint_as_float(d[c & 1][k4 * 4 + 3]), b.w);
                        }
                    }
                    float cm0 = tv[0], cm1 = tv[1], cm2 = tv[2], cm3 = tv[3];
#pragma unroll
                    for (int k = 4; k < CH; k += 4) {
                        cm0 = fmaxf(cm0, tv[k]); cm1 = fmaxf(cm1, tv[k + 1]);
                        cm2 = fmaxf(cm2, tv[k + 2]); cm3 = fmaxf(cm3, tv[k + 3]);
                    }
                    const float cm = fmaxf(fmaxf(cm0, cm1), fmaxf(cm2, cm3));
                    // branch-free lazy stabiliser: moves only when the chunk maximum exceeds it by more than 2^64
                    // (first chunk of an item, or a much closer column shows up); otherwise the rescale is *1.
                    const float m_new = (cm > m_used + 64.f) ? cm : m_used;
                    ssum *= sdb_ex2(m_used - m_new);
                    m_used = m_new;
                    if constexpr (PACKED) {
                        const uint64_t nm2 = pack2(-m_used, -m_used);
                        uint64_t acc2[CH / 2];
#pragma unroll
                        for (int k = 0; k < CH; k += 2) {
                            const uint64_t x2 = fadd2(pack2(tv[k], tv[k + 1]), nm2);
                            if (POLY_EVERY > 0 && ((k / 2) % POLY_EVERY) == POLY_EVERY - 1) {
                                acc2[k / 2] = exp2_poly2(x2);       // FMA-pipe exponential
                            } else {
                                float a0, a1;
                                unpack2(x2, a0, a1);
                                acc2[k / 2] = pack2(sdb_ex2(a0), sdb_ex2(a1));
                            }
                        }
#pragma unroll
                        for (int w = CH / 4; w > 0; w >>= 1)
#pragma unroll
                            for (int k = 0; k < w; ++k) acc2[k] = fadd2(acc2[k], acc2[k + w]);
                        float e0, e1;
                        unpack2(acc2[0], e0, e1);
                        ssum += e0 + e1;
                    } else {
#pragma unroll
                        for (int k = 0; k < CH; ++k) tv[k] = sdb_ex2(tv[k] - m_used);
#pragma unroll
                        for (int w = CH / 2; w > 0; w >>= 1)
#pragma unroll
                            for (int k = 0; k < w; ++k) tv[k] += tv[k + w];
                        ssum += tv[0];
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { mbar_arrive(t_empty + acc); mbar_arrive(b_empty + bs); }
            }
            }
            // merge the column slices' row statistics and emit one (max, sum) per row
            if (part > 0) sMerge[(part - 1) * TILE_M + row_in_tile] = make_float2(m_used, ssum);
            asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
            if (part == 0) {
#pragma unroll
                for (int q = 0; q < PARTS - 1; ++q) {
                    const float2 o = sMerge[q * TILE_M + row_in_tile];
                    const float mn = fmaxf(m_used, o.x);
                    ssum = ssum * sdb_ex2(m_used - mn) + o.y * sdb_ex2(o.x - mn);
                    m_used = mn;
                }
                const int64_t row = (int64_t)rt * TILE_M + row_in_tile;
                if (row < a.n_p) a.partial[(int64_t)sp * a.n_p + row] = make_float2(m_used, ssum);
            }
            asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
            if (a.upd.counters != nullptr) {
                // fused update: the partials of this item were stored before the bar.sync above; one thread's acq_rel atomic on
                // the row tile's arrival counter publishes them and tells whether every split of the tile is in
                uint32_t* s_last = tmem_holder + 2;
                if (threadIdx.x == 64) {
                    unsigned prev;
                    asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(prev) : "l"(a.upd.counters + rt) : "memory");
                    const bool last = (prev == (unsigned)a.n_splits - 1u);
                    if (last) a.upd.counters[rt] = 0u;
                    *s_last = last ? 1u : 0u;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
                if (*s_last != 0u && part == 0) {
                    const int64_t row = (int64_t)rt * TILE_M + row_in_tile;
                    if (row < a.n_p) {
                        const sdb_tc_update& u = a.upd;
                        const double norm = u.norms[row];
                        const double Li = sdb_combine_partials(a.partial, a.n_splits, a.n_p, row, norm * u.c1, u.m_next, u.bad_flag);
                        u.L[row] = Li;
                        sdb_update_row_deferred(row, Li, u.logmarg[row], norm, u.eps, u.alpha, u.log_n_other, u.c1, u.pot, u.frame, u.la_old,
                                                u.bias_out, u.flag2, u.tick, u.log_tau, u.log_floor);
                    }
                }
                asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
            }
        }
        }  // MODE_LSE
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

// ------------------------------------------------------------------------------------------- fp16 split preparation
__global__ void prep_split_kernel(const double* __restrict__ x, int64_t n, int d, const double* __restrict__ center,
                                  double pow2, __half* __restrict__ out, int64_t n_pad, int dp, double* __restrict__ norms) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    __half* row = out + i * (2 * dp);
    double nn = 0.0;
    const double inv = 1.0 / pow2;
    for (int k = 0; k < dp; ++k) {
        __half hi = __float2half_rn(0.f), lo = hi;
        if (i < n && k < d) {
            const double v = (x[i * d + k] - center[k]) * pow2;
            hi = __double2half(v);
            const double rem = v - (double)__half2float(hi);
            lo = __double2half(rem);
            const double rep = ((double)__half2float(hi) + (double)__half2float(lo)) * inv;
            nn += rep * rep;
        }
        row[k] = hi;
        row[dp + k] = lo;
    }
    if (i < n) norms[i] = nn;
}

__global__ void absmax_centered_kernel(const double* __restrict__ x, int64_t n, int d, const double* __restrict__ center,
                                       unsigned long long* __restrict__ out_bits) {
    double mx = 0.0;
    const int64_t total = n * d;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride)
        mx = fmax(mx, fabs(x[e] - center[e % d]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0 && mx > 0.0) atomicMax(out_bits, (unsigned long long)__double_as_longlong(mx));
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
    return fn;
}

int make_tmap(CUtensorMap* tm, const void* base, int64_t rows_pad, int dp, int box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return SDB_E_DRIVER;
    cuuint64_t gdim[2] = {(cuuint64_t)(2 * dp), (cuuint64_t)rows_pad};
    cuuint64_t gstride[1] = {(cuuint64_t)(2 * dp) * 2};
    cuuint32_t box[2] = {(cuuint32_t)dp, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMapSwizzle sw = dp == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : dp == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : SDB_E_DRIVER;
}

template <int DP, int EPI_WARPS, bool PACKED, int POLY_EVERY = 0, bool PIPE = false, bool XTILE = false, bool FIXED_M = false,
          bool RUNSUM = false>
int launch_tc_v(const CUtensorMap& tmP, const CUtensorMap& tmQ, const TcArgs& a, int n_ctas, cudaStream_t st) {
    auto kern = lse_pass_tc_kernel<DP, EPI_WARPS, PACKED, POLY_EVERY, MODE_LSE, PIPE, XTILE, FIXED_M, RUNSUM>;
    constexpr int smem = TcSmem<DP>::TOTAL;
    static size_t attr_set[SDB_MAX_DEVICES] = {0};
    { cudaError_t e = sdb_ensure_smem(kern, (size_t)smem, attr_set); if (e != cudaSuccess) return (int)e; }
    return (int)sdb_launch(kern, dim3(n_ctas), dim3(32 * (2 + EPI_WARPS)), smem, st, tmP, tmQ, a);
}

template <int DP, int MODE>
int launch_tc_sweep(const CUtensorMap& tmP, const CUtensorMap& tmQ, const TcArgs& a, int n_ctas, cudaStream_t st) {
    auto kern = lse_pass_tc_kernel<DP, 8, false, 0, MODE>;
    constexpr int smem = TcSmem<DP>::TOTAL_SWEEP;
    static size_t attr_set[SDB_MAX_DEVICES] = {0};
    { cudaError_t e = sdb_ensure_smem(kern, (size_t)smem, attr_set); if (e != cudaSuccess) return (int)e; }
    kern<<<n_ctas, 32 * (2 + 8), smem, st>>>(tmP, tmQ, a);
    SDB_LAUNCH_STATUS();
}

// Tuning variant (development knob SDB_TC_VARIANT): 0 scalar epilogue, 1 packed f32x2 (all exponentials on the SFU),
// 2/3/4 packed with every 2nd/3rd/4th pair of exponentials evaluated by the FMA-pipe polynomial,
// 5 packed + software-pipelined (FP32 stage of chunk c+1 interleaved with the SFU stage of chunk c),
// 6 the same pipeline carried across tile boundaries.
int tc_variant() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("SDB_TC_VARIANT");
        v = e ? atoi(e) : 5;
        if (v < 0 || v > 6) v = 5;
    }
    return v;
}

template <int DP>
int launch_tc(const CUtensorMap& tmP, const CUtensorMap& tmQ, const TcArgs& a, int n_ctas, cudaStream_t st) {
    if (a.row_m != nullptr) {                                                                                // predicted stabiliser
        static int runsum = -1;          // development knob SDB_TC_RUNSUM=1: running sums instead of the per-chunk reduction tree
                                         // (measured r2: 0.793 vs 0.809 of the SFU peak at 131072^2 - not adopted)
        if (runsum < 0) { const char* e = getenv("SDB_TC_RUNSUM"); runsum = (e && *e == '1') ? 1 : 0; }
        static int poly = -1;            // development knob SDB_TC_POLY=4|8|16: every 4th/8th/16th pair of exponentials by the
                                         // FMA-pipe polynomial inside the pipelined predicted pass
        if (poly < 0) { const char* e = getenv("SDB_TC_POLY"); poly = e ? atoi(e) : 0; }
        if (poly == 4) return launch_tc_v<DP, 8, true, 4, true, false, true, false>(tmP, tmQ, a, n_ctas, st);
        if (poly == 8) return launch_tc_v<DP, 8, true, 8, true, false, true, false>(tmP, tmQ, a, n_ctas, st);
        if (poly == 16) return launch_tc_v<DP, 8, true, 16, true, false, true, false>(tmP, tmQ, a, n_ctas, st);
        return runsum ? launch_tc_v<DP, 8, true, 0, true, false, true, true>(tmP, tmQ, a, n_ctas, st)
                      : launch_tc_v<DP, 8, true, 0, true, false, true, false>(tmP, tmQ, a, n_ctas, st);
    }
    switch (tc_variant()) {
        case 0: return launch_tc_v<DP, 8, false>(tmP, tmQ, a, n_ctas, st);
        case 2: return launch_tc_v<DP, 8, true, 2>(tmP, tmQ, a, n_ctas, st);
        case 3: return launch_tc_v<DP, 8, true, 3>(tmP, tmQ, a, n_ctas, st);
        case 4: return launch_tc_v<DP, 8, true, 4>(tmP, tmQ, a, n_ctas, st);
        case 5: return launch_tc_v<DP, 8, true, 0, true>(tmP, tmQ, a, n_ctas, st);
        case 6: return launch_tc_v<DP, 8, true, 0, true, true>(tmP, tmQ, a, n_ctas, st);
        default: return launch_tc_v<DP, 8, true>(tmP, tmQ, a, n_ctas, st);
    }
}

}  // namespace

extern "C" {

int sdb_absmax_centered_f64(const double* x, int64_t n, int d, const double* center, double* out_max, void* stream) {
    SDB_CHECK_ARG(x && center && out_max && d > 0 && n >= 0);
    if (n == 0) return 0;
    int grid = (int)((n * d + 256 * 8 - 1) / (256 * 8));
    if (grid > 1184) grid = 1184;
    absmax_centered_kernel<<<grid, 256, 0, sdb_stream(stream)>>>(x, n, d, center, reinterpret_cast<unsigned long long*>(out_max));
    SDB_LAUNCH_STATUS();
}

int sdb_prep_points_split_f16(const double* x, int64_t n, int d, const double* center, int pow2_exp, void* out16, int64_t n_pad,
                              int dp, double* norms, void* stream) {
    return sdb_prep_points_split_f16_scaled(x, n, d, center, ldexp(1.0, pow2_exp), out16, n_pad, dp, norms, stream);
}

int sdb_prep_points_split_f16_scaled(const double* x, int64_t n, int d, const double* center, double prescale, void* out16,
                                     int64_t n_pad, int dp, double* norms, void* stream) {
    SDB_CHECK_ARG(x && center && out16 && norms && d > 0 && dp >= d && (dp == 16 || dp == 32 || dp == 64) && n_pad >= n &&
                  (n_pad % TILE_N) == 0 && prescale > 0.0);
    prep_split_kernel<<<(unsigned)((n_pad + 255) / 256), 256, 0, sdb_stream(stream)>>>(x, n, d, center, prescale,
                                                                                       reinterpret_cast<__half*>(out16), n_pad, dp, norms);
    SDB_LAUNCH_STATUS();
}

int sdb_lse_pass_tc(const void* p16, int64_t n_p, int64_t n_p_pad, const void* q16, int64_t n_q, int64_t n_q_pad, int dp,
                    const float* bias_padded, float scale, int tiles_per_split, int n_ctas, float* partial, void* stream) {
    return sdb_lse_pass_tc_pred(p16, n_p, n_p_pad, q16, n_q, n_q_pad, dp, bias_padded, scale, tiles_per_split, n_ctas, nullptr, partial,
                                stream);
}

int sdb_lse_pass_tc_pred(const void* p16, int64_t n_p, int64_t n_p_pad, const void* q16, int64_t n_q, int64_t n_q_pad, int dp,
                         const float* bias_padded, float scale, int tiles_per_split, int n_ctas, const float* row_m, float* partial,
                         void* stream) {
    return sdb_lse_pass_tc_fused(p16, n_p, n_p_pad, q16, n_q, n_q_pad, dp, bias_padded, scale, tiles_per_split, n_ctas, row_m, partial,
                                 nullptr, stream);
}

int sdb_lse_pass_tc_fused(const void* p16, int64_t n_p, int64_t n_p_pad, const void* q16, int64_t n_q, int64_t n_q_pad, int dp,
                          const float* bias_padded, float scale, int tiles_per_split, int n_ctas, const float* row_m, float* partial,
                          const sdb_tc_update* upd, void* stream) {
    SDB_CHECK_ARG(p16 && q16 && bias_padded && partial && n_p > 0 && n_q > 0 && tiles_per_split > 0 && n_ctas > 0);
    if (upd) SDB_CHECK_ARG(upd->counters && upd->norms && upd->L && upd->logmarg && upd->pot && upd->frame && upd->la_old && upd->flag2 &&
                           upd->eps > 0.0);
    SDB_CHECK_ARG((n_p_pad % TILE_N) == 0 && (n_q_pad % TILE_N) == 0 && n_p_pad >= n_p && n_q_pad >= n_q);
    SDB_CHECK_ARG(((uintptr_t)p16 % 128) == 0 && ((uintptr_t)q16 % 128) == 0 && ((uintptr_t)bias_padded % 16) == 0);
    if (!(dp == 16 || dp == 32 || dp == 64)) return SDB_E_UNSUPPORTED;
    CUtensorMap tmP, tmQ;
    int rc = make_tmap(&tmP, p16, n_p_pad, dp, TILE_M);
    if (rc) return rc;
    rc = make_tmap(&tmQ, q16, n_q_pad, dp, TILE_N);
    if (rc) return rc;
    TcArgs a{};
    if (upd) a.upd = *upd;
    a.bias = bias_padded;
    a.row_m = row_m;
    a.scale = scale;
    a.n_p = n_p;
    a.n_row_tiles = (int)((n_p + TILE_M - 1) / TILE_M);
    a.n_col_tiles = (int)((n_q + TILE_N - 1) / TILE_N);
    a.tiles_per_split = tiles_per_split;
    a.n_splits = (a.n_col_tiles + tiles_per_split - 1) / tiles_per_split;   // never empty, by construction
    a.partial = reinterpret_cast<float2*>(partial);
    cudaStream_t st = sdb_stream(stream);
    switch (dp) {
        case 16: return launch_tc<16>(tmP, tmQ, a, n_ctas, st);
        case 32: return launch_tc<32>(tmP, tmQ, a, n_ctas, st);
        default: return launch_tc<64>(tmP, tmQ, a, n_ctas, st);
    }
}

int sdb_lse_pass_tc_groups(const void* p16, int64_t n_p, int64_t n_p_pad, const void* q16, int64_t n_q_pad, int dp,
                           const float* bias_padded, float scale, const int* split_tiles, int n_splits, int n_ctas, float* partial,
                           void* stream) {
    SDB_CHECK_ARG(p16 && q16 && bias_padded && partial && split_tiles && n_p > 0 && n_q_pad > 0 && n_splits > 0 && n_ctas > 0);
    SDB_CHECK_ARG((n_p_pad % TILE_N) == 0 && (n_q_pad % TILE_N) == 0 && n_p_pad >= n_p);
    SDB_CHECK_ARG(((uintptr_t)p16 % 128) == 0 && ((uintptr_t)q16 % 128) == 0 && ((uintptr_t)bias_padded % 16) == 0);
    if (!(dp == 16 || dp == 32 || dp == 64)) return SDB_E_UNSUPPORTED;
    CUtensorMap tmP, tmQ;
    int rc = make_tmap(&tmP, p16, n_p_pad, dp, TILE_M);
    if (rc) return rc;
    rc = make_tmap(&tmQ, q16, n_q_pad, dp, TILE_N);
    if (rc) return rc;
    TcArgs a{};
    a.split_tiles = split_tiles;
    a.bias = bias_padded;
    a.scale = scale;
    a.n_p = n_p;
    a.n_row_tiles = (int)((n_p + TILE_M - 1) / TILE_M);
    a.n_col_tiles = (int)(n_q_pad / TILE_N);
    a.tiles_per_split = 0;
    a.n_splits = n_splits;
    a.partial = reinterpret_cast<float2*>(partial);
    cudaStream_t st = sdb_stream(stream);
    switch (dp) {
        case 16: return launch_tc<16>(tmP, tmQ, a, n_ctas, st);
        case 32: return launch_tc<32>(tmP, tmQ, a, n_ctas, st);
        default: return launch_tc<64>(tmP, tmQ, a, n_ctas, st);
    }
}

// Tensor-core forms of sdb_cost_histogram / sdb_cost_collect (same contract; distances from the fp16 hi/lo split
// points: dist = row_norms[i] + col_norms_padded[j] - 2 x_i.y_j, col_norms_padded = |y_j|^2 for j < n_q and
// >= 3e38 for the padding).  scale must be -2 * 2^(-2*pow2_exp).  hist / counts accumulate (zero them first).
static int tc_sweep_common(int mode, const void* p16, int64_t n_p, int64_t n_p_pad, const void* q16, int64_t n_q, int64_t n_q_pad,
                           int dp, const float* row_norms, const float* col_norms_padded, float scale, int tiles_per_split,
                           int n_ctas, TcArgs& a, void* stream) {
    if (!p16 || !q16 || !row_norms || !col_norms_padded || n_p <= 0 || n_q <= 0 || tiles_per_split <= 0 || n_ctas <= 0)
        return SDB_E_INVALID;
    if ((n_p_pad % TILE_N) || (n_q_pad % TILE_N) || n_p_pad < n_p || n_q_pad < n_q) return SDB_E_INVALID;
    if (!(dp == 16 || dp == 32 || dp == 64)) return SDB_E_UNSUPPORTED;
    CUtensorMap tmP, tmQ;
    int rc = make_tmap(&tmP, p16, n_p_pad, dp, TILE_M);
    if (rc) return rc;
    rc = make_tmap(&tmQ, q16, n_q_pad, dp, TILE_N);
    if (rc) return rc;
    a.bias = col_norms_padded;
    a.scale = scale;
    a.n_p = n_p;
    a.n_q = n_q;
    a.row_norms = row_norms;
    a.n_row_tiles = (int)((n_p + TILE_M - 1) / TILE_M);
    a.n_col_tiles = (int)((n_q + TILE_N - 1) / TILE_N);
    a.tiles_per_split = tiles_per_split;
    a.n_splits = (a.n_col_tiles + tiles_per_split - 1) / tiles_per_split;
    a.partial = nullptr;
    cudaStream_t st = sdb_stream(stream);
    if (mode == MODE_HIST) {
        switch (dp) {
            case 16: return launch_tc_sweep<16, MODE_HIST>(tmP, tmQ, a, n_ctas, st);
            case 32: return launch_tc_sweep<32, MODE_HIST>(tmP, tmQ, a, n_ctas, st);
            default: return launch_tc_sweep<64, MODE_HIST>(tmP, tmQ, a, n_ctas, st);
        }
    }
    switch (dp) {
        case 16: return launch_tc_sweep<16, MODE_COLLECT>(tmP, tmQ, a, n_ctas, st);
        case 32: return launch_tc_sweep<32, MODE_COLLECT>(tmP, tmQ, a, n_ctas, st);
        default: return launch_tc_sweep<64, MODE_COLLECT>(tmP, tmQ, a, n_ctas, st);
    }
}

int sdb_cost_histogram_tc(const void* p16, int64_t n_p, int64_t n_p_pad, const void* q16, int64_t n_q, int64_t n_q_pad, int dp,
                          const float* row_norms, const float* col_norms_padded, float scale, int tiles_per_split, int n_ctas,
                          float lo, float hi, int n_bins, unsigned long long* hist, unsigned long long* counts2, void* stream) {
    SDB_CHECK_ARG(hist && counts2 && n_bins > 0 && n_bins <= SWEEP_BINS && hi > lo);
    TcArgs a{};
    a.lo = lo; a.hi = hi; a.inv_width = (float)n_bins / (hi - lo); a.n_bins = n_bins; a.hist = hist; a.counts = counts2;
    return tc_sweep_common(MODE_HIST, p16, n_p, n_p_pad, q16, n_q, n_q_pad, dp, row_norms, col_norms_padded, scale, tiles_per_split,
                           n_ctas, a, stream);
}

int sdb_cost_collect_tc(const void* p16, int64_t n_p, int64_t n_p_pad, const void* q16, int64_t n_q, int64_t n_q_pad, int dp,
                        const float* row_norms, const float* col_norms_padded, float scale, int tiles_per_split, int n_ctas,
                        float lo, float hi, const double* x, const double* y, int d, double* cand, unsigned long long cap,
                        unsigned long long* counts2, void* stream) {
    SDB_CHECK_ARG(x && y && cand && counts2 && d > 0);
    TcArgs a{};
    a.lo = lo; a.hi = hi; a.counts = counts2; a.x64 = x; a.y64 = y; a.d = d; a.cand = cand; a.cap = cap;
    return tc_sweep_common(MODE_COLLECT, p16, n_p, n_p_pad, q16, n_q, n_q_pad, dp, row_norms, col_norms_padded, scale,
                           tiles_per_split, n_ctas, a, stream);
}

}  // extern "C"
