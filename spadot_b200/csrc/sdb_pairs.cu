// K3 (SIMT fp32 form) and K5 sweeps: one tile engine that streams 64x64 tiles of point pairs.
//
// The cost / kernel matrix is never materialised: a CTA keeps a 64-row slab of the "row" point
// set in shared memory (feature-major, so the inner product loop reads conflict-free float4s),
// streams 64-column slabs of the other set through a cp.async double buffer, forms the 4x4
// register micro-tile of x_i.y_j (or |x_i-y_j|^2) per thread and hands it to an epilogue functor:
//   LseEpi      online (max, sum 2^(t-max)) per row           -> Sinkhorn half-iteration
//   HistEpi     count-below + uniform histogram of the costs   -> median bracketing
//   CollectEpi  count-below + fp64 re-evaluation of bracket    -> exact median candidates
// Replaces gemv/gemtv/update_k of ref: SpaDOT/utils/OT_loss/ot_func.cpp:43-249,547-568 and the
// cost + np.median of ref: SpaDOT/utils/OT_loss/ot_solvers.py:102-103.
#include "sdb_common.cuh"

namespace {

constexpr int BM = 64;   // rows per CTA
constexpr int BN = 64;   // columns per streamed tile
constexpr int NT = 256;  // 16 x 16 threads, 4x4 micro-tile each
constexpr int MAX_DPAD = 128;

struct PairArgs {
    const float* pt; int64_t ldp; int64_t n_p;
    const float* qt; int64_t ldq; int64_t n_q;
    int dpad;
    const float* bias;            // may be null (treated as 0 / valid)
    const int64_t* split_bounds;  // device, n_splits+1 entries
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

// --------------------------------------------------------------------------------------------
// One work item = (64-row slab, column split).  Used by the one-item-per-CTA kernel below and by the persistent
// small-problem kernel, whose CTAs walk many items per pass.
template <bool DIRECT, class Epi>
__device__ __forceinline__ void pair_tile_item(const PairArgs& a, const typename Epi::Params& ep, int row_tile, int split, float* smem) {
    const int dpad = a.dpad;
    float* Ps = smem;                       // [dpad][BM]
    float* Qs = Ps + dpad * BM;             // [2][dpad][BN]
    float* Bs = Qs + 2 * dpad * BN;         // [2][BN]

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t row0 = (int64_t)row_tile * BM;
    const int64_t c0 = a.split_bounds[split], c1 = a.split_bounds[split + 1];
    const int64_t j_begin = c0 & ~int64_t(3);
    const int n_tiles = (c1 > c0) ? (int)((c1 - j_begin + BN - 1) / BN) : 0;

    Epi epi(ep, row0 + ty * 4, split, a.n_p);

    // rows slab, once
    for (int idx = tid; idx < dpad * (BM / 4); idx += NT) {
        int k = idx / (BM / 4), c4 = idx % (BM / 4);
        cp_async16(Ps + k * BM + c4 * 4, a.pt + (int64_t)k * a.ldp + row0 + c4 * 4);
    }
    auto load_q = [&](int t, int buf) {
        const int64_t j0 = j_begin + (int64_t)t * BN;
        float* q = Qs + buf * dpad * BN;
        for (int idx = tid; idx < dpad * (BN / 4); idx += NT) {
            int k = idx / (BN / 4), c4 = idx % (BN / 4);
            cp_async16(q + k * BN + c4 * 4, a.qt + (int64_t)k * a.ldq + j0 + c4 * 4);
        }
        if (tid < BN) {
            int64_t j = j0 + tid;
            float b = SDB_NEG_SENTINEL;
            if (j >= c0 && j < c1) b = a.bias ? __ldcg(a.bias + j) : 0.f;    // L2: rewritten between passes
            Bs[buf * BN + tid] = b;
        }
    };
    if (n_tiles > 0) load_q(0, 0);
    cp_async_commit();

    for (int t = 0; t < n_tiles; ++t) {
        const int buf = t & 1;
        if (t + 1 < n_tiles) load_q(t + 1, buf ^ 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();

        const float* q = Qs + buf * dpad * BN;
        float acc[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;

#pragma unroll 4
        for (int k = 0; k < dpad; ++k) {
            const float4 pa = *reinterpret_cast<const float4*>(Ps + k * BM + ty * 4);
            const float4 qb = *reinterpret_cast<const float4*>(q + k * BN + tx * 4);
            const float pr[4] = {pa.x, pa.y, pa.z, pa.w};
            const float qc[4] = {qb.x, qb.y, qb.z, qb.w};
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (DIRECT) {
                        const float df = pr[r] - qc[c];
                        acc[r][c] = fmaf(df, df, acc[r][c]);
                    } else {
                        acc[r][c] = fmaf(pr[r], qc[c], acc[r][c]);
                    }
                }
        }
        const float4 bv = *reinterpret_cast<const float4*>(Bs + buf * BN + tx * 4);
        const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
        epi.tile(acc, bb, j_begin + (int64_t)t * BN + tx * 4);
        __syncthreads();
    }
    cp_async_wait<0>();
    epi.finish(tx);
}

template <bool DIRECT, class Epi>
__global__ void __launch_bounds__(NT) pair_tile_kernel(PairArgs a, typename Epi::Params ep) {
    extern __shared__ __align__(16) float smem[];
    pair_tile_item<DIRECT, Epi>(a, ep, (int)blockIdx.x, (int)blockIdx.y, smem);
}

// -------------------------------------------------------------------------------------------- LSE
struct LseEpi {
    // The SIMT pass works on direct differences, acc = |x_i - y_j|^2 (like scipy's cdist behind ref: ot_solvers.py:102): no
    // |x|^2 + |y|^2 - 2 x.y cancellation, so the fp32 rounding of the tile math is relative to the COST of a pair, not to
    // |x||y| (at eps = 0.01 the dot-product form put 1.3e-5 on the marginals, the direct form 1e-6; the inner loop pays one
    // extra FADD per feature, which the latency-bound problems this path serves do not notice).
    // scale = scale_hi + scale_lo: the exponent scale -c1*log2(e) as an fp32 pair.  A single fp32 rounds it by up to 2^-24
    // relative, and that error is common to every term of every row: dominant terms have scale*x.y ~ 150 at eps = 0.01, i.e.
    // an LSE bias of up to 9e-6 that goes one-to-one into the plan (measured: marginals off by 1.3e-5 at eps = 0.01).
    struct Params { float scale, scale_lo; float2* partial; };
    Params p; int64_t row; int split; int64_t n_p;
    float m[4], s[4];
    __device__ LseEpi(const Params& p_, int64_t row_, int split_, int64_t n_p_) : p(p_), row(row_), split(split_), n_p(n_p_) {
#pragma unroll
        for (int r = 0; r < 4; ++r) { m[r] = SDB_NEG_SENTINEL; s[r] = 0.f; }
    }
    __device__ __forceinline__ void tile(const float (&acc)[4][4], const float (&b)[4], int64_t) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const float t0 = fmaf(p.scale, acc[r][0], fmaf(p.scale_lo, acc[r][0], b[0]));
            const float t1 = fmaf(p.scale, acc[r][1], fmaf(p.scale_lo, acc[r][1], b[1]));
            const float t2 = fmaf(p.scale, acc[r][2], fmaf(p.scale_lo, acc[r][2], b[2]));
            const float t3 = fmaf(p.scale, acc[r][3], fmaf(p.scale_lo, acc[r][3], b[3]));
            const float mn = fmaxf(m[r], fmaxf(fmaxf(t0, t1), fmaxf(t2, t3)));
            s[r] = s[r] * sdb_ex2(m[r] - mn) + ((sdb_ex2(t0 - mn) + sdb_ex2(t1 - mn)) + (sdb_ex2(t2 - mn) + sdb_ex2(t3 - mn)));
            m[r] = mn;
        }
    }
    __device__ __forceinline__ void finish(int tx) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int o = 1; o < 16; o <<= 1) {
                const float m2 = __shfl_xor_sync(0xffffffffu, m[r], o);
                const float s2 = __shfl_xor_sync(0xffffffffu, s[r], o);
                const float mn = fmaxf(m[r], m2);
                s[r] = s[r] * sdb_ex2(m[r] - mn) + s2 * sdb_ex2(m2 - mn);
                m[r] = mn;
            }
            if (tx == 0 && row + r < n_p) p.partial[(int64_t)split * n_p + row + r] = make_float2(m[r], s[r]);
        }
    }
};

// -------------------------------------------------------------------------------------------- median sweeps
struct HistEpi {
    struct Params { float lo, hi, inv_width; int n_bins; unsigned long long* hist; unsigned long long* counts; };
    Params p; int64_t row; int64_t n_p;
    unsigned long long below, inside;
    __device__ HistEpi(const Params& p_, int64_t row_, int, int64_t n_p_) : p(p_), row(row_), n_p(n_p_), below(0), inside(0) {}
    __device__ __forceinline__ void tile(const float (&acc)[4][4], const float (&b)[4], int64_t) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            if (row + r >= n_p) continue;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (b[c] < -1e29f) continue;
                const float dist = acc[r][c];
                if (dist < p.lo) { ++below; }
                else if (dist < p.hi) {
                    int bin = (int)((dist - p.lo) * p.inv_width);
                    bin = min(max(bin, 0), p.n_bins - 1);
                    atomicAdd(p.hist + bin, 1ull);
                    ++inside;
                }
            }
        }
    }
    __device__ __forceinline__ void finish(int) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            below += __shfl_xor_sync(0xffffffffu, below, o);
            inside += __shfl_xor_sync(0xffffffffu, inside, o);
        }
        if ((threadIdx.x & 31) == 0) {
            if (below) atomicAdd(p.counts, below);
            if (inside) atomicAdd(p.counts + 1, inside);
        }
    }
};

struct CollectEpi {
    struct Params { float lo, hi; const double* x; const double* y; int d; double* cand; unsigned long long cap; unsigned long long* counts; };
    Params p; int64_t row; int64_t n_p;
    unsigned long long below;
    __device__ CollectEpi(const Params& p_, int64_t row_, int, int64_t n_p_) : p(p_), row(row_), n_p(n_p_), below(0) {}
    __device__ __forceinline__ void tile(const float (&acc)[4][4], const float (&b)[4], int64_t col) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            if (row + r >= n_p) continue;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (b[c] < -1e29f) continue;
                const float dist = acc[r][c];
                if (dist < p.lo) { ++below; }
                else if (dist < p.hi) {
                    const double* xi = p.x + (row + r) * p.d;
                    const double* yj = p.y + (col + c) * p.d;
                    double s = 0.0;
                    for (int k = 0; k < p.d; ++k) { const double df = xi[k] - yj[k]; s = __dadd_rn(s, __dmul_rn(df, df)); }
                    const unsigned long long slot = atomicAdd(p.counts + 1, 1ull);
                    if (slot < p.cap) p.cand[slot] = s;
                }
            }
        }
    }
    __device__ __forceinline__ void finish(int) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) below += __shfl_xor_sync(0xffffffffu, below, o);
        if ((threadIdx.x & 31) == 0 && below) atomicAdd(p.counts, below);
    }
};

template <bool DIRECT, class Epi>
int launch_pairs(const PairArgs& a, const typename Epi::Params& ep, int n_splits, cudaStream_t st) {
    if (!a.pt || !a.qt || !a.split_bounds) return SDB_E_INVALID;
    if (a.dpad <= 0 || a.dpad > MAX_DPAD || (a.dpad & 3)) return SDB_E_UNSUPPORTED;
    if ((a.ldp & 3) || (a.ldq & 3) || n_splits <= 0 || n_splits > 65535) return SDB_E_INVALID;
    if (a.ldp < ((a.n_p + BM - 1) / BM) * BM || a.ldq < ((a.n_q + BN - 1) / BN) * BN + BN) return SDB_E_INVALID;
    if (a.n_p <= 0) return 0;
    const size_t smem = sizeof(float) * ((size_t)a.dpad * BM + 2 * (size_t)a.dpad * BN + 2 * BN);
    auto kern = pair_tile_kernel<DIRECT, Epi>;
    static size_t smem_set[SDB_MAX_DEVICES] = {0};      // per instantiation and device
    { cudaError_t e = sdb_ensure_smem(kern, smem, smem_set); if (e != cudaSuccess) return (int)e; }
    dim3 grid((unsigned)((a.n_p + BM - 1) / BM), (unsigned)n_splits);
    kern<<<grid, NT, smem, st>>>(a, ep);
    SDB_LAUNCH_STATUS();
}

// -------------------------------------------------------------------------------------------- persistent sweeps
// Small problems (ChickenHeart timepoints: ~2000 x 2000 spots) are latency-bound: five launches of a few
// microseconds of work each per iteration.  This kernel runs `n_sweeps` whole iterations in ONE cooperative launch:
//   row pass over (slab, split) items -> grid barrier -> combine + update f -> grid barrier ->
//   column pass -> grid barrier -> combine + update g -> grid barrier -> absorb (u <- f, v <- g if tau was exceeded)
// Buffers that change inside the kernel (potentials, bias vectors, partials, the flag) are read with plain loads or
// cp.async.cg (L2), never through the read-only path; the grid barrier's fences order them.
struct PersistArgs {
    PairArgs row, col;                 // row pass: P = x, Q = y, bias = bias_y;  column pass: P = y, Q = x, bias = bias_x
    int ns_row, ns_col;
    float scale, scale_lo;
    int direct;
    float2* partial_row; float2* partial_col;
    const double* norms_x; const double* norms_y;
    float* bias_x; float* bias_y;
    double *f, *g, *u, *v, *la_old, *lb_old, *Lr, *Lc;
    const double* logp; const double* logq;
    int* flag;
    double eps, c1, alpha1, alpha2, log_m, log_N, log_tau, log_floor;
    int n_sweeps, first_tick, lr_known_first;
    unsigned int* barrier;             // [2]: arrival counter, generation (zero-initialised once by the host)
};

// Grid barrier (all CTAs are co-resident: cooperative launch).  One acq_rel atomic per CTA on the arrival counter;
// the last arriver resets it and releases the next generation, the others poll the generation with acquire loads.
// Release/acquire at gpu scope (instead of two membar.sc) orders every thread's earlier writes before, and every
// later read after, the barrier: the bar.sync pair extends that from thread 0 to the whole CTA.
// TWO-LEVEL arrival: the CTAs are spread over SDB_BARRIER_GROUPS group counters (128 bytes apart, so they live in different
// L2 sectors) and only the last arriver of each group touches the top counter - ~20 + 16 same-address atomics deep instead of
// ~300 (same-address atomics serialise in L2; with one counter the barrier cost 6 us of a 17 us ChickenHeart phase).
// bar[0] = top counter, bar[1] = generation, bar[32 + 32*g] = counter of group g  (SDB_BARRIER_WORDS unsigned ints, zeroed).
// Measured (ChickenHeart 1966 x 1916, 95 iterations): 3.70 ms with 16 groups against 3.45 ms with ONE counter - the second
// dependent atomic of the group's last arriver costs more than the same-address serialisation it saves - so the group count is 1
// (one counter, one atomic per CTA); the two-level code stays for grids far beyond 300 CTAs.
constexpr unsigned SDB_BARRIER_GROUPS = 1;
__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int& gen) {
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int arrived;
        bool released = false;
        if constexpr (SDB_BARRIER_GROUPS == 1) {
            asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(arrived) : "l"(bar) : "memory");
            if (arrived == gridDim.x - 1) {
                asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(bar), "r"(0u) : "memory");
                asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar + 1) : "memory");
                released = true;
            }
        } else {
            const unsigned n_groups = min(SDB_BARRIER_GROUPS, gridDim.x);
            const unsigned g = blockIdx.x % n_groups;
            const unsigned in_group = (gridDim.x - g + n_groups - 1) / n_groups;
            unsigned int* sub = bar + 32 + 32 * g;
            asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(arrived) : "l"(sub) : "memory");
            if (arrived == in_group - 1) {
                asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(sub), "r"(0u) : "memory");
                asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(arrived) : "l"(bar) : "memory");
                if (arrived == n_groups - 1) {
                    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(bar), "r"(0u) : "memory");
                    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar + 1) : "memory");
                    released = true;
                }
            }
        }
        if (!released) {
            // poll with relaxed loads (an acquire load per poll invalidates L1 every time: CCTL.IVALL was 4 % of the one-launch
            // solve's samples), then one acquire fence once the new generation is seen
            unsigned int gg;
            do {
                asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(gg) : "l"(bar + 1) : "memory");
            } while (gg == gen);
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
        }
    }
    gen += 1u;
    __syncthreads();
}

__global__ void __launch_bounds__(NT) sinkhorn_persistent_kernel(PersistArgs a) {
    extern __shared__ __align__(16) float smem[];
    unsigned int gen = *reinterpret_cast<volatile unsigned int*>(a.barrier + 1);
    const int64_t n = a.row.n_p, m = a.col.n_p;
    const int row_tiles = (int)((n + BM - 1) / BM), col_tiles = (int)((m + BM - 1) / BM);
    const int64_t gtid = (int64_t)blockIdx.x * NT + threadIdx.x, gsize = (int64_t)gridDim.x * NT;
    const int64_t gwarp = gtid >> 5, n_warps = gsize >> 5;
    const int lane = threadIdx.x & 31;
    for (int sweep = 0; sweep < a.n_sweeps; ++sweep) {
        const int tick = a.first_tick + sweep;
        const bool skip_row_pass = (sweep == 0 && a.lr_known_first);
        if (!skip_row_pass) {
            LseEpi::Params ep{a.scale, a.scale_lo, a.partial_row};
            for (int item = blockIdx.x; item < row_tiles * a.ns_row; item += gridDim.x) {
                if (a.direct) pair_tile_item<true, LseEpi>(a.row, ep, item % row_tiles, item / row_tiles, smem);
                else pair_tile_item<false, LseEpi>(a.row, ep, item % row_tiles, item / row_tiles, smem);
                __syncthreads();
            }
            grid_barrier(a.barrier, gen);
        }
        for (int64_t i = gwarp; i < n; i += n_warps) {               // one warp per row: lanes over the splits
            double Li;
            if (skip_row_pass) Li = a.Lr[i];
            else Li = sdb_combine_partials_warp(a.partial_row, a.ns_row, n, i, a.norms_x[i] * a.c1);
            if (lane == 0) {
                a.Lr[i] = Li;
                sdb_update_row(i, Li, a.logp[i], a.norms_x[i], a.eps, a.alpha1, a.log_m, a.c1, a.f, a.u, a.la_old, a.bias_x, a.flag,
                               tick, a.log_tau, a.log_floor);
            }
        }
        grid_barrier(a.barrier, gen);
        {
            LseEpi::Params ep{a.scale, a.scale_lo, a.partial_col};
            for (int item = blockIdx.x; item < col_tiles * a.ns_col; item += gridDim.x) {
                if (a.direct) pair_tile_item<true, LseEpi>(a.col, ep, item % col_tiles, item / col_tiles, smem);
                else pair_tile_item<false, LseEpi>(a.col, ep, item % col_tiles, item / col_tiles, smem);
                __syncthreads();
            }
        }
        grid_barrier(a.barrier, gen);
        for (int64_t j = gwarp; j < m; j += n_warps) {
            const double Lj = sdb_combine_partials_warp(a.partial_col, a.ns_col, m, j, a.norms_y[j] * a.c1);
            if (lane == 0) {
                a.Lc[j] = Lj;
                sdb_update_row(j, Lj, a.logq[j], a.norms_y[j], a.eps, a.alpha2, a.log_N, a.c1, a.g, a.v, a.lb_old, a.bias_y, a.flag,
                               tick, a.log_tau, a.log_floor);
            }
        }
        grid_barrier(a.barrier, gen);
        if (*reinterpret_cast<volatile int*>(a.flag) == tick) {          // ot_func.cpp:792-819 in total potentials
            for (int64_t i = gtid; i < n; i += gsize) a.u[i] = a.f[i];
            for (int64_t j = gtid; j < m; j += gsize) a.v[j] = a.g[j];
        }
        // u, v are next written/read by the same warps' lane 0 / these threads only after later barriers
    }
}


// -------------------------------------------------------------------------------------------- whole solve, one launch
// optimal_transport_duality_gap (ref: SpaDOT/utils/OT_loss/ot_solvers.py:240-449 driving update_process<double>,
// ot_func.cpp:831-930) for a small problem in ONE cooperative launch: the six epsilon stages, their stopping rules
// (stage 0-4 criterion :897-922, final-stage duality gap :493-544) and every iteration run on the device; the host
// reads one small result record at the end.  An iteration costs TWO grid barriers: the (max, sum) partials of a
// 64-row slab are combined, and the potential updated, by whichever CTA finishes the slab's last column split
// (a per-slab arrival counter), so there is no separate update phase between a pass and the next one.
// tau bookkeeping without a phase of its own: updaters of tick t stamp flag[t & 1]; an updater of tick t+1 that finds
// flag[t & 1] == t first absorbs its own row (frame <- potential), which is all `absorb` is in total potentials.
struct SolveArgs {
    PairArgs row, col;
    int ns_row, ns_col;
    float2* partial_row; float2* partial_col;
    float* bias_x; float* bias_y;
    double *f, *g, *u, *v, *la_old, *lb_old, *Lr, *Lc;
    const double* logp; const double* logq;
    const double* norms_x; const double* norms_y;   // |x_i|^2, |y_j|^2 of the fp32 points (dot-product stages)
    double eps_stage[6], xy_max, dot_limit;
    int* flag;                          // caller's absorb flag (left consistent: last tick that exceeded tau)
    int* flag2;                         // [2] ping-pong flags of this kernel
    int res_tiles;                      // resident form: tiles of shared memory per CTA (>= the tiles any CTA owns)
    int strip_rows, strip_cols;         // strip form: rows / columns owned by every CTA (the last owners get fewer)
    double inv_med, lambda1, lambda2, epsilon, epsilon0, tolerance, log_tau, log_m, log_N, dx, dy;
    int batch_size; long long max_iter; int first_tick;
    unsigned int* barrier;              // [2]
    unsigned int* counters;             // [row_tiles + col_tiles], zero on entry, left zero
    double* scratch;                    // [gridDim.x * 10]
    sdb_solve_result* result;
};

template <int K>
__device__ void solve_grid_sums(double (&acc)[K], double* scratch, unsigned int* bar, unsigned int& gen, double (&out)[K]) {
    __shared__ double sm[32][K];
    __shared__ double tot[K];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = (int)(blockDim.x >> 5);
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const double v = sdb_warp_sum(acc[k]);
        if (lane == 0) sm[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < K) {
        double v = 0.0;
        for (int w = 0; w < n_warps; ++w) v += sm[w][threadIdx.x];
        scratch[(size_t)blockIdx.x * K + threadIdx.x] = v;
    }
    grid_barrier(bar, gen);
    // every CTA adds the per-CTA partials in the same order: identical sums, identical decisions everywhere
    for (int k = warp; k < K; k += n_warps) {
        double v = 0.0;
        for (unsigned b = lane; b < gridDim.x; b += 32) v += __ldcg(scratch + (size_t)b * K + k);
        v = sdb_warp_sum(v);
        if (lane == 0) tot[k] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) out[k] = tot[k];
    __syncthreads();
}

__device__ __forceinline__ double solve_safe_ratio(double num2, double den2) {       // sinkhorn._safe_ratio
    const double a = sqrt(num2), b = sqrt(den2);
    if (isinf(a) && isinf(b)) return NAN;
    return a / (1.0 + b);
}

// One pass over (slab, split) items; the CTA that completes a slab combines its partials and, with UPDATE, updates the
// potential of those 64 rows (pot, frame, la_old, bias_out, flags) - else it only stores the LSE.
// Combine the (max, sum) partials of the 64 rows of slab `tile` (ns splits each) and, with UPDATE, update their potentials:
// the work of whichever CTA completed the slab.  4 threads per row, splits strided over them.
template <bool UPDATE, int MAXP = 4>
__device__ __forceinline__ void solve_finish_slab(int tile, int64_t n, const float2* partial, int ns, double* L, const double* norms,
                                                  double c1n, const double* logmarg, double eps, double alpha, double log_n_other,
                                                  double* pot, double* frame, double* la_old, float* bias_out, int* flag2, int tick,
                                                  double log_tau, bool pending) {
        // 4 threads per row, splits strided over them; every partial is loaded once (all loads in flight together)
        const int64_t i = (int64_t)tile * BM + (threadIdx.x >> 2);
        const int sub = threadIdx.x & 3;
        // MAXP partials per thread stay in registers (4 * MAXP splits); beyond that they are reloaded for the sum
        float2 pv[MAXP];
        float mx = -INFINITY;
#pragma unroll
        for (int q = 0; q < MAXP; ++q) {
            const int sp = sub + 4 * q;
            pv[q] = (i < n && sp < ns) ? __ldcg(partial + (int64_t)sp * n + i) : make_float2(SDB_NEG_SENTINEL, 0.f);
        }
#pragma unroll
        for (int q = 0; q < MAXP; ++q)
            if (pv[q].x > -1e29f && pv[q].y > 0.f) mx = fmaxf(mx, pv[q].x);
        if (i < n)
            for (int sp = sub + 4 * MAXP; sp < ns; sp += 4) {
                const float2 ps = __ldcg(partial + (int64_t)sp * n + i);
                if (ps.x > -1e29f && ps.y > 0.f) mx = fmaxf(mx, ps.x);
            }
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        double S = 0.0;
        if (mx > -INFINITY) {
#pragma unroll
            for (int q = 0; q < MAXP; ++q)
                if (pv[q].x > -1e29f && pv[q].y > 0.f) S += (double)pv[q].y * exp2((double)pv[q].x - (double)mx);
            if (i < n)
                for (int sp = sub + 4 * MAXP; sp < ns; sp += 4) {
                    const float2 ps = __ldcg(partial + (int64_t)sp * n + i);
                    if (ps.x > -1e29f && ps.y > 0.f) S += (double)ps.y * exp2((double)ps.x - (double)mx);
                }
        }
        S += __shfl_xor_sync(0xffffffffu, S, 1);
        S += __shfl_xor_sync(0xffffffffu, S, 2);
        if (sub == 0 && i < n) {
            const double nc = norms[i] * c1n;
            const double Li = (mx > -INFINITY) ? SDB_LN2 * ((double)mx + log2(S)) - nc : -INFINITY;
            L[i] = Li;
            if (UPDATE) {
                const double old = pot[i];
                double fr = frame[i];
                if (pending) { fr = old; frame[i] = old; }           // absorb of the previous tick, row by row
                la_old[i] = (old - fr) / eps;
                const double nv = eps * alpha * (logmarg[i] - (Li - log_n_other));
                pot[i] = nv;
                const double b = SDB_LOG2E * (nv / eps - nc);
                bias_out[i] = (b > (double)SDB_NEG_SENTINEL) ? (float)b : SDB_NEG_SENTINEL;
                if ((nv - fr) / eps > log_tau) atomicMax(flag2 + (tick & 1), tick);
            }
        }
    }

// `direct`: tile arithmetic of this stage (see sdb_lse_pass_simt); c1n = c1 for dot-product tiles (row norm subtracted from the
// LSE, norm inside the bias), 0 for direct ones.
template <bool UPDATE>
__device__ void solve_pass(const PairArgs& pa, bool direct, float scale_hi, float scale_lo, float2* partial, int ns,
                           unsigned int* counters, double* L, const double* norms, double c1n, const double* logmarg, double eps,
                           double alpha, double log_n_other, double* pot, double* frame, double* la_old, float* bias_out,
                           int* flag2, int tick, double log_tau, float* smem) {
    __shared__ int s_last;
    const int64_t n = pa.n_p;
    const int tiles = (int)((n + BM - 1) / BM);
    LseEpi::Params ep{scale_hi, scale_lo, partial};
    const bool pending = UPDATE && (*reinterpret_cast<volatile int*>(flag2 + ((tick - 1) & 1)) == tick - 1);
    for (int item = blockIdx.x; item < tiles * ns; item += gridDim.x) {
        const int tile = item % tiles;
        if (direct) pair_tile_item<true, LseEpi>(pa, ep, tile, item / tiles, smem);
        else pair_tile_item<false, LseEpi>(pa, ep, tile, item / tiles, smem);
        // the partials of this item were stored by the CTA's threads before the bar.sync; thread 0's acq_rel atomic at gpu scope
        // publishes them (release is cumulative over what happened-before it) and, for the last arriver, acquires the others'
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned prev;
            asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(prev) : "l"(counters + tile) : "memory");
            s_last = (prev == (unsigned)ns - 1u);
            if (s_last) counters[tile] = 0u;
        }
        __syncthreads();
        if (s_last)
            solve_finish_slab<UPDATE>(tile, n, partial, ns, L, norms, c1n, logmarg, eps, alpha, log_n_other, pot, frame, la_old, bias_out,
                                      flag2, tick, log_tau, pending);
        __syncthreads();
    }
}

// ---- resident cost tiles -----------------------------------------------------------------------------------------------------
// A ChickenHeart-sized coupling has a few million pairs: its whole squared-distance matrix D_ij = |x_i - y_j|^2 (fp32, direct
// differences) fits in the shared memory of the 148 SMs (4 x 16 KB tiles per CTA, two CTAs per SM hold 1 184 tiles = 4.8 M
// pairs).  Built ONCE per solve and kept on chip, it turns a half-iteration into  t = fma(s, D, bias),  ex2,  add  per pair
// (s = -c1*log2(e) as an fp32 pair, one rounding at the magnitude of t) - no dot products, no coordinate traffic - and BOTH
// passes read the same tile (rows reduce along x, columns along y).
// Tile k of a CTA is global tile blockIdx.x + k*gridDim.x (row-tile major); layout [rr][tid][4]: thread (ty, tx) owns rows
// 4*ty + rr, columns 4*tx .. 4*tx + 3, so every access is a conflict-free LDS.128.
constexpr int RES_MAX_TILES = 6;

__device__ void res_build_tiles(const SolveArgs& a, float* tiles, int n_owned, int C) {
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t n = a.row.n_p, m = a.col.n_p;
    const int dpad = a.row.dpad;
    for (int k = 0; k < n_owned; ++k) {
        const int tile = blockIdx.x + k * gridDim.x;
        const int64_t r0 = (int64_t)(tile / C) * BM + ty * 4, c0 = (int64_t)(tile % C) * BN + tx * 4;
        float acc[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
        for (int q = 0; q < dpad; ++q) {                       // padded rows / columns of xt, yt are zero-filled and in bounds
            const float4 xa = *reinterpret_cast<const float4*>(a.row.pt + (int64_t)q * a.row.ldp + r0);
            const float4 yb = *reinterpret_cast<const float4*>(a.col.pt + (int64_t)q * a.col.ldp + c0);
            const float xr[4] = {xa.x, xa.y, xa.z, xa.w}, yc[4] = {yb.x, yb.y, yb.z, yb.w};
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) { const float df = xr[r] - yc[c]; acc[r][c] = fmaf(df, df, acc[r][c]); }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            float4 v;
            float* e = reinterpret_cast<float*>(&v);
#pragma unroll
            for (int c = 0; c < 4; ++c)
                e[c] = (r0 + r < n && c0 + c < m) ? acc[r][c] : 1.0e30f;      // masked pair: scale * 1e30 stays finite, 2^that = 0
            *reinterpret_cast<float4*>(tiles + ((size_t)k * 4 + r) * (NT * 4) + tid * 4) = v;
        }
    }
    __syncthreads();
}

// One half-iteration over the resident tiles.  ROWS: reduce along the columns of each tile with bias = bias of the columns
// (NULL = 0: the sum of exp(-C/eps)), one (max, sum) per (column tile, row); else along the rows with the bias of the rows, one
// per (row tile, column).  Then the same slab-completion protocol as solve_pass.
template <bool UPDATE, bool ROWS>
__device__ void res_pass(const SolveArgs& a, const float* tiles, int n_owned, int R, int C, float s_hi, float s_lo, const float* bias,
                         float2* partial,
                         unsigned int* counters, double* L, const double* norms, const double* logmarg, double eps, double alpha,
                         double log_n_other, double* pot, double* frame, double* la_old, float* bias_out, int tick, float2* red) {
    __shared__ int s_last[RES_MAX_TILES];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4, lane = tid & 31, warp = tid >> 5;
    const int64_t n = a.row.n_p, m = a.col.n_p;
    const int64_t n_out = ROWS ? n : m;
    const int ns = ROWS ? C : R;
    const bool pending = UPDATE && (*reinterpret_cast<volatile int*>(a.flag2 + ((tick - 1) & 1)) == tick - 1);
    // all the bias loads of this CTA's tiles first: one L2 round trip for the whole pass instead of one per tile
    float4 bq[RES_MAX_TILES];
#pragma unroll
    for (int k = 0; k < RES_MAX_TILES; ++k) {
        bq[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < n_owned && bias) {
            const int tile = blockIdx.x + k * gridDim.x;
            if (ROWS) bq[k] = __ldcg(reinterpret_cast<const float4*>(bias + (int64_t)(tile % C) * BN + tx * 4));       // padded to 256
            else bq[k] = __ldcg(reinterpret_cast<const float4*>(bias + (int64_t)(tile / C) * BM + ty * 4));
        }
    }
#pragma unroll
    for (int k = 0; k < RES_MAX_TILES; ++k) {
        if (k >= n_owned) break;
        const int tile = blockIdx.x + k * gridDim.x;
        const int rt = tile / C, ct = tile % C;
        float4 cs[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) cs[r] = *reinterpret_cast<const float4*>(tiles + ((size_t)k * 4 + r) * (NT * 4) + tid * 4);
        if (ROWS) {
            const float4 b = bq[k];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float t0 = fmaf(s_hi, cs[r].x, fmaf(s_lo, cs[r].x, b.x)), t1 = fmaf(s_hi, cs[r].y, fmaf(s_lo, cs[r].y, b.y));
                const float t2 = fmaf(s_hi, cs[r].z, fmaf(s_lo, cs[r].z, b.z)), t3 = fmaf(s_hi, cs[r].w, fmaf(s_lo, cs[r].w, b.w));
                float mx = fmaxf(fmaxf(t0, t1), fmaxf(t2, t3));
                float sm = (sdb_ex2(t0 - mx) + sdb_ex2(t1 - mx)) + (sdb_ex2(t2 - mx) + sdb_ex2(t3 - mx));
#pragma unroll
                for (int o = 1; o < 16; o <<= 1) {                 // the 16 threads of a row sit in one half-warp
                    const float m2 = __shfl_xor_sync(0xffffffffu, mx, o), s2 = __shfl_xor_sync(0xffffffffu, sm, o);
                    const float mn = fmaxf(mx, m2);
                    sm = sm * sdb_ex2(mx - mn) + s2 * sdb_ex2(m2 - mn);
                    mx = mn;
                }
                const int64_t row = (int64_t)rt * BM + ty * 4 + r;
                if (tx == 0 && row < n) partial[(int64_t)ct * n + row] = make_float2(mx, sm);
            }
        } else {
            const float br[4] = {bq[k].x, bq[k].y, bq[k].z, bq[k].w};
            float mxc[4], smc[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float d0 = reinterpret_cast<const float*>(&cs[0])[c], d1 = reinterpret_cast<const float*>(&cs[1])[c];
                const float d2 = reinterpret_cast<const float*>(&cs[2])[c], d3 = reinterpret_cast<const float*>(&cs[3])[c];
                const float t0 = fmaf(s_hi, d0, fmaf(s_lo, d0, br[0])), t1 = fmaf(s_hi, d1, fmaf(s_lo, d1, br[1]));
                const float t2 = fmaf(s_hi, d2, fmaf(s_lo, d2, br[2])), t3 = fmaf(s_hi, d3, fmaf(s_lo, d3, br[3]));
                float mx = fmaxf(fmaxf(t0, t1), fmaxf(t2, t3));
                float sm = (sdb_ex2(t0 - mx) + sdb_ex2(t1 - mx)) + (sdb_ex2(t2 - mx) + sdb_ex2(t3 - mx));
                // the two ty of this warp (lanes tx and tx + 16)
                const float m2 = __shfl_xor_sync(0xffffffffu, mx, 16), s2 = __shfl_xor_sync(0xffffffffu, sm, 16);
                const float mn = fmaxf(mx, m2);
                smc[c] = sm * sdb_ex2(mx - mn) + s2 * sdb_ex2(m2 - mn);
                mxc[c] = mn;
            }
            if (lane < 16) {
#pragma unroll
                for (int c = 0; c < 4; ++c) red[((size_t)k * (NT / 32) + warp) * BN + tx * 4 + c] = make_float2(mxc[c], smc[c]);
            }
        }
    }
    if (!ROWS) {
        __syncthreads();
        // the 8 warps' contributions to every column of every owned tile
        for (int idx = tid; idx < n_owned * BN; idx += NT) {
            const int k = idx / BN, cl = idx % BN;
            const int tile = blockIdx.x + k * gridDim.x;
            const float2* rk = red + (size_t)k * (NT / 32) * BN;
            float mx = rk[cl].x, sm = rk[cl].y;
#pragma unroll
            for (int w = 1; w < NT / 32; ++w) {
                const float2 o = rk[w * BN + cl];
                const float mn = fmaxf(mx, o.x);
                sm = sm * sdb_ex2(mx - mn) + o.y * sdb_ex2(o.x - mn);
                mx = mn;
            }
            const int64_t col = (int64_t)(tile % C) * BN + cl;
            if (col < m) partial[(int64_t)(tile / C) * m + col] = make_float2(mx, sm);
        }
    }
    // slab completion for all owned tiles at once: one bar.sync, the arrivals issued by different threads in parallel
    __syncthreads();
    if (tid < n_owned) {
        const int tile = blockIdx.x + tid * gridDim.x;
        const int slab = ROWS ? tile / C : tile % C;
        unsigned prev;
        asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(prev) : "l"(counters + slab) : "memory");
        const bool last = (prev == (unsigned)ns - 1u);
        if (last) counters[slab] = 0u;
        s_last[tid] = last ? 1 : 0;
    }
    __syncthreads();
    for (int k = 0; k < n_owned; ++k) {
        if (!s_last[k]) continue;
        const int tile = blockIdx.x + k * gridDim.x;
        solve_finish_slab<UPDATE, 8>(ROWS ? tile / C : tile % C, n_out, partial, ns, L, norms, 0.0, logmarg, eps, alpha, log_n_other, pot, frame,
                                  la_old, bias_out, a.flag2, tick, a.log_tau, pending);
    }
    __syncthreads();
}

// ---- strips: the owner-computes form ------------------------------------------------------------------------------------------
// For the reference's own example sizes (ChickenHeart: <= 1967 spots per side) a CTA can hold, in shared memory, BOTH the rows of
// D_ij = |x_i - y_j|^2 it owns (n/grid full rows) and the columns it owns (m/grid full columns).  A half-iteration then needs no
// partial results at all: the owner reduces its whole rows (a warp per row), updates their potential on the spot and the only
// traffic between CTAs is the updated vector itself - one read of the other side's bias after the grid barrier.  Against the
// tile forms this removes, per half-iteration, the partial stores, the per-slab arrival atomics and the last arriver's combine
// from the dependent chain between two barriers (profiles/r2_ncu_solve_kernel_hot_lines.txt: 48 % of the one-launch solve was
// waiting in barriers for exactly that tail).  Same tile arithmetic as the resident form: t = fma(s_hi, D, fma(s_lo, D, bias)).
constexpr int STRIP_REG_LEN = 2048;      // rows up to this length are reduced from registers
constexpr int STRIP_NT = 512;            // threads per CTA of the strip form: 16 warps = one owned row per warp at ChickenHeart sizes

constexpr int STRIP_DREG = 32;           // points of up to this (padded) dimension are held in registers while a strip is built

__device__ void strip_build(const float* __restrict__ pt, int64_t ldp, const float* __restrict__ qt, int64_t ldq, int64_t n_q,
                            int dpad, int first, int n_owned, int ld, float* strip, float* stage) {
    // stage: the owned points' coordinates, [n_owned][dpad] (borrowed space, see the caller)
    for (int idx = threadIdx.x; idx < n_owned * dpad; idx += blockDim.x) {
        const int r = idx / dpad, q = idx - r * dpad;
        stage[idx] = pt[(int64_t)q * ldp + first + r];
    }
    __syncthreads();
    for (int j = threadIdx.x; j < ld; j += blockDim.x) {
        if (j >= n_q) {
            for (int r = 0; r < n_owned; ++r) strip[(size_t)r * ld + j] = 1.0e30f;    // padding: scale * 1e30 stays finite, 2^that = 0
        } else if (dpad <= STRIP_DREG) {
            float yq[STRIP_DREG];                                // this column's coordinates: read once (coalesced over j)
#pragma unroll
            for (int q = 0; q < STRIP_DREG; ++q) yq[q] = q < dpad ? qt[(int64_t)q * ldq + j] : 0.f;
            for (int r = 0; r < n_owned; ++r) {
                float acc = 0.f;
#pragma unroll
                for (int q = 0; q < STRIP_DREG; ++q)             // the accumulation order of res_build_tiles / the direct tiles
                    if (q < dpad) { const float df = stage[r * dpad + q] - yq[q]; acc = fmaf(df, df, acc); }
                strip[(size_t)r * ld + j] = acc;
            }
        } else {
            for (int r = 0; r < n_owned; ++r) {
                float acc = 0.f;
                for (int q = 0; q < dpad; ++q) { const float df = stage[r * dpad + q] - qt[(int64_t)q * ldq + j]; acc = fmaf(df, df, acc); }
                strip[(size_t)r * ld + j] = acc;
            }
        }
    }
    __syncthreads();
}

// One half-iteration over this CTA's strip: (max, sum) of 2^t over each owned row, then its LSE and, with UPDATE, the potential
// update of solve_finish_slab - by the warp that reduced the row.
//
// DATAFLOW instead of a grid barrier between the passes of a batch.  The other side's bias vector travels as 64-bit words
// (fp32 bias | 31-bit production number | 1 tau bit): `tin` is polled (relaxed gpu-scope loads, eight in flight per thread) until
// every entry carries production number `expect`, and the owner publishes its own updated entries to `tout` with number
// `out_seq`.  A pass can only start when every CTA has finished the previous one (every CTA owns at least one row and one
// column, so its entries are part of what is waited for), which is also what makes overwriting the arrays safe.  The value and
// its number arrive in one single-copy-atomic load, so no fence is needed for them; everything else an update writes (pot,
// frame, la_old, L) is read by other CTAs only after the full barrier that ends the batch.  The tau bookkeeping rides in the
// tags: bit 63 says "this update exceeded tau", the OR over a staged vector is returned in `bits`.
// tin == NULL: zero bias, nothing to wait for.
template <bool UPDATE>
__device__ void strip_pass(const float* strip, int first, int n_owned, int len, int ld, float s_hi, float s_lo,
                           const unsigned long long* tin, unsigned expect, float* sbias, double* L, const double* logmarg, double eps,
                           double alpha, double log_n_other, double* pot, double* frame, double* la_old, unsigned long long* tout,
                           unsigned out_seq, int* flag2, int tick, double log_tau, bool pend_base, bool add_bits, bool& pending_out,
                           bool& bits_out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = (int)(blockDim.x >> 5);
    int my_bits = 0;
    for (int base = 0; base < ld; base += 8 * (int)blockDim.x) {
        unsigned long long v[8];
        if (tin) {
            unsigned spins = 0;
            bool ok;
            do {
                ok = true;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int j = base + k * (int)blockDim.x + (int)threadIdx.x;
                    v[k] = 0ull;
                    if (j < len) asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v[k]) : "l"(tin + j) : "memory");
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int j = base + k * (int)blockDim.x + (int)threadIdx.x;
                    if (j < len && (((unsigned)(v[k] >> 32)) & 0x7fffffffu) != expect) ok = false;
                }
                if (!ok && ++spins > (1u << 22)) __trap();       // a protocol error must end the launch, not hang the device
            } while (!ok);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int j = base + k * (int)blockDim.x + (int)threadIdx.x;
            if (j < ld) {
                float b = 0.f;
                if (tin && j < len) { b = __uint_as_float((unsigned)v[k]); my_bits |= (int)(v[k] >> 63); }
                sbias[j] = b;
            }
        }
    }
    const bool bits = __syncthreads_or(my_bits) != 0;            // (also the barrier between staging and use)
    bits_out = bits;
    const bool pending = UPDATE && (pend_base || (add_bits && bits));
    pending_out = pending;
    // Lane k of a warp finishes (LSE, update, publication) the k-th row the warp reduced: the rows' fp64 tails run side by side
    // instead of one after the other.  What the update reads from global memory is requested before the reductions, so that
    // its L2 round trip overlaps them.
    const int my_r = warp + lane * n_warps;                     // the row this lane finishes
    double u_old = 0.0, u_fr = 0.0, u_lm = 0.0;
    if (UPDATE && my_r < n_owned) { u_old = pot[first + my_r]; u_fr = frame[first + my_r]; u_lm = logmarg[first + my_r]; }
    float my_mall = SDB_NEG_SENTINEL;
    double my_S = 0.0;
    for (int r0 = warp; r0 < n_owned; r0 += 32 * n_warps) {     // (more than 32 rows per warp: finish them group by group)
    for (int r = r0, k = 0; r < n_owned && k < 32; r += n_warps, ++k) {
        const float* row = strip + (size_t)r * ld;
        float mall;
        double S;
        if (ld <= STRIP_REG_LEN) {
            // the whole row in registers (64 values per lane): every exponent first, one maximum for the warp, then 64 independent
            // ex2 per lane - no running maximum, no dependency from one chunk to the next (a CTA runs alone on its SM: 255
            // registers per thread are there to be used)
            float t[STRIP_REG_LEN / 128][4];
            float mx = SDB_NEG_SENTINEL;
#pragma unroll
            for (int q = 0; q < STRIP_REG_LEN / 128; ++q) {
                const int j0 = lane * 4 + q * 128;               // ld is a multiple of 4: conflict-free LDS.128
                if (j0 < ld) {
                    const float4 dv = *reinterpret_cast<const float4*>(row + j0);
                    const float4 bv = *reinterpret_cast<const float4*>(sbias + j0);
                    t[q][0] = fmaf(s_hi, dv.x, fmaf(s_lo, dv.x, bv.x)); t[q][1] = fmaf(s_hi, dv.y, fmaf(s_lo, dv.y, bv.y));
                    t[q][2] = fmaf(s_hi, dv.z, fmaf(s_lo, dv.z, bv.z)); t[q][3] = fmaf(s_hi, dv.w, fmaf(s_lo, dv.w, bv.w));
                } else {
                    t[q][0] = t[q][1] = t[q][2] = t[q][3] = SDB_NEG_SENTINEL;
                }
                mx = fmaxf(mx, fmaxf(fmaxf(t[q][0], t[q][1]), fmaxf(t[q][2], t[q][3])));
            }
            mall = mx;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) mall = fmaxf(mall, __shfl_xor_sync(0xffffffffu, mall, o));
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
            for (int q = 0; q < STRIP_REG_LEN / 128; ++q) {
                s0 += sdb_ex2(t[q][0] - mall); s1 += sdb_ex2(t[q][1] - mall);
                s2 += sdb_ex2(t[q][2] - mall); s3 += sdb_ex2(t[q][3] - mall);
            }
            S = (double)((s0 + s1) + (s2 + s3));
        } else {
            float mx = SDB_NEG_SENTINEL, sm = 0.f;
            for (int j0 = lane * 4; j0 < ld; j0 += 128) {
                const float4 dv = *reinterpret_cast<const float4*>(row + j0);
                const float4 bv = *reinterpret_cast<const float4*>(sbias + j0);
                const float t0 = fmaf(s_hi, dv.x, fmaf(s_lo, dv.x, bv.x)), t1 = fmaf(s_hi, dv.y, fmaf(s_lo, dv.y, bv.y));
                const float t2 = fmaf(s_hi, dv.z, fmaf(s_lo, dv.z, bv.z)), t3 = fmaf(s_hi, dv.w, fmaf(s_lo, dv.w, bv.w));
                const float cm = fmaxf(fmaxf(t0, t1), fmaxf(t2, t3));
                if (cm > mx) { sm *= sdb_ex2(mx - cm); mx = cm; }
                sm += (sdb_ex2(t0 - mx) + sdb_ex2(t1 - mx)) + (sdb_ex2(t2 - mx) + sdb_ex2(t3 - mx));
            }
            mall = mx;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) mall = fmaxf(mall, __shfl_xor_sync(0xffffffffu, mall, o));
            S = (double)sm * (double)sdb_ex2(mx - mall);
        }
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) S += __shfl_xor_sync(0xffffffffu, S, o);      // the 32 lane sums are combined in fp64
        if (lane == k) { my_mall = mall; my_S = S; }
    }
        const int fin_r = r0 + lane * n_warps;
        if (fin_r < n_owned) {
            const int64_t i = (int64_t)first + fin_r;
            const float mall = my_mall;
            const double S = my_S;
            if (UPDATE && r0 != warp) { u_old = pot[i]; u_fr = frame[i]; u_lm = logmarg[i]; }      // later groups: no prefetch
            const bool valid = mall > -1e29f && S > 0.0;
            const double Li = valid ? SDB_LN2 * ((double)mall + log2(S)) : -INFINITY;
            L[i] = Li;
            if (UPDATE) {
                const double old = u_old;
                double fr = u_fr;
                if (pending) { fr = old; frame[i] = old; }       // absorb of the previous tick, row by row
                la_old[i] = (old - fr) / eps;
                const double nv = eps * alpha * (u_lm - (Li - log_n_other));
                pot[i] = nv;
                const double b = SDB_LOG2E * (nv / eps);
                const float bf = (b > (double)SDB_NEG_SENTINEL) ? (float)b : SDB_NEG_SENTINEL;
                const bool over = (nv - fr) / eps > log_tau;
                if (over) atomicMax(flag2 + (tick & 1), tick);   // the batch-end checks read this one (after a full barrier)
                const unsigned long long word = ((unsigned long long)(out_seq | (over ? 0x80000000u : 0u)) << 32) |
                                                (unsigned long long)__float_as_uint(bf);
                asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(tout + i), "l"(word) : "memory");
            }
        }
    }
    __syncthreads();                                             // sbias is rewritten by the next pass
}

__device__ __forceinline__ unsigned long long strip_word(double b, unsigned seq, bool over) {
    const float bf = (b > (double)SDB_NEG_SENTINEL) ? (float)b : SDB_NEG_SENTINEL;
    return ((unsigned long long)(seq | (over ? 0x80000000u : 0u)) << 32) | (unsigned long long)__float_as_uint(bf);
}

// FORM: 0 streamed tiles (coordinates re-read every pass), 1 resident 64x64 cost tiles, 2 strips (owner computes whole rows)
template <int FORM>
__global__ void __launch_bounds__(FORM == 2 ? STRIP_NT : NT) sinkhorn_solve_kernel(SolveArgs a) {
    constexpr bool RESIDENT = (FORM == 1);
    constexpr bool STRIPS = (FORM == 2);
    constexpr int NTK = STRIPS ? STRIP_NT : NT;                  // threads per CTA of this form
    extern __shared__ __align__(16) float smem[];
    unsigned int gen = *reinterpret_cast<volatile unsigned int*>(a.barrier + 1);
    const int64_t n = a.row.n_p, m = a.col.n_p;
    const int row_tiles = (int)((n + BM - 1) / BM);
    const int64_t gtid = (int64_t)blockIdx.x * NTK + threadIdx.x, gsize = (int64_t)gridDim.x * NTK;
    unsigned int* cnt_row = a.counters;
    unsigned int* cnt_col = a.counters + row_tiles;
    double eps = a.eps_stage[0];
    int tick = a.first_tick - 1;
    double gap = INFINITY, sumK = 0.0;
    int status = 0, total = 0;
    bool lr_known = false;
    PairArgs row0 = a.row;                                               // row pass with g = 0 (sum of exp(-C/eps))
    row0.bias = nullptr;
    // resident form: this CTA's tiles of the scaled cost matrix live in shared memory for a whole epsilon stage
    const int col_tiles = (int)((m + BN - 1) / BN);
    int n_owned = 0;
    float2* red = nullptr;
    if constexpr (RESIDENT) {
        for (int k = 0; k < RES_MAX_TILES; ++k) n_owned += (blockIdx.x + k * (int)gridDim.x < row_tiles * col_tiles) ? 1 : 0;
        red = reinterpret_cast<float2*>(smem + (size_t)a.res_tiles * 4 * NT * 4);
        res_build_tiles(a, smem, n_owned, col_tiles);                     // once: the squared distances do not depend on epsilon
    }
    // strip form: the rows and the columns of D this CTA owns, and a staging buffer for the other side's bias
    const int ldm = (int)((m + 3) & ~(int64_t)3), ldn = (int)((n + 3) & ~(int64_t)3);
    float *srow = nullptr, *scol = nullptr, *sbias = nullptr;
    int first_row = 0, n_rows = 0, first_col = 0, n_cols = 0;
    // tagged bias vectors of the strip form (see strip_pass), behind the per-CTA partial sums in `scratch`
    unsigned long long* tagx = reinterpret_cast<unsigned long long*>(a.scratch + (size_t)SDB_SOLVE_MAX_CTAS * 10);
    unsigned long long* tagy = tagx + n;
    unsigned seq_x = 0, seq_y = 0;                   // productions of bias_x / bias_y so far (the arrays start zeroed)
    bool fl_row_prev = false, pend = false, bits = false;
    if constexpr (STRIPS) {
        srow = smem;
        scol = srow + (size_t)a.strip_rows * ldm;
        sbias = scol + (size_t)a.strip_cols * ldn;
        // balanced partition: with gridDim.x <= min(n, m) every CTA owns at least one row and one column (the dataflow
        // protocol of strip_pass relies on it), and never more than strip_rows / strip_cols
        first_row = (int)(((int64_t)blockIdx.x * n) / gridDim.x);
        first_col = (int)(((int64_t)blockIdx.x * m) / gridDim.x);
        n_rows = (int)(((int64_t)(blockIdx.x + 1) * n) / gridDim.x) - first_row;
        n_cols = (int)(((int64_t)(blockIdx.x + 1) * m) / gridDim.x) - first_col;
        // (the coordinate staging borrows the bias buffer: max(ldm, ldn) floats; the host checked strip_rows/cols * dpad fits)
        strip_build(a.row.pt, a.row.ldp, a.col.pt, a.col.ldp, m, a.row.dpad, first_row, n_rows, ldm, srow, sbias);
        strip_build(a.col.pt, a.col.ldp, a.row.pt, a.row.ldp, n, a.row.dpad, first_col, n_cols, ldn, scol, sbias);
        __syncthreads();
    }
    for (int e = 0; e <= 5 && status == 0; ++e) {
        eps = a.eps_stage[e];                                            // ot_solvers.py:218,240,254 (host arithmetic)
        const double c1 = a.inv_med / eps;
        const double alpha1 = a.lambda1 / (a.lambda1 + eps), alpha2 = a.lambda2 / (a.lambda2 + eps);
        const bool direct = RESIDENT || STRIPS || 2.0 * c1 * SDB_LOG2E * a.xy_max > a.dot_limit;
        const double c1n = direct ? 0.0 : c1;
        const double sc = direct ? -c1 * SDB_LOG2E : 2.0 * c1 * SDB_LOG2E;
        const float sc_hi = (float)sc, sc_lo = (float)(sc - (double)sc_hi);
        const double scd = -c1 * SDB_LOG2E;                               // the sum of exp(-C/eps) always runs on direct tiles
        const float scd_hi = (float)scd, scd_lo = (float)(scd - (double)scd_hi);
        const bool final_stage = (e == 5);
        const double threshold = final_stage ? a.tolerance : 1e-6;       // ot_solvers.py:262
        const int n_inner = final_stage ? a.batch_size : 5;              // ot_func.cpp:867
        // stage start: absorb (u <- f, v <- g), old_a = old_b = 1, bias of the first row pass at the new epsilon
        for (int64_t i = gtid; i < n; i += gsize) { a.u[i] = a.f[i]; a.la_old[i] = 0.0; }
        for (int64_t j = gtid; j < m; j += gsize) {
            const double gj = a.g[j];
            a.v[j] = gj;
            a.lb_old[j] = 0.0;
            const double b = SDB_LOG2E * (gj / eps - a.norms_y[j] * c1n);
            a.bias_y[j] = (b > (double)SDB_NEG_SENTINEL) ? (float)b : SDB_NEG_SENTINEL;
            if constexpr (STRIPS) tagy[j] = strip_word(b, seq_y + 1, false);
        }
        if constexpr (STRIPS) { ++seq_y; fl_row_prev = false; pend = false; }
        if (gtid == 0) { a.flag2[0] = -1; a.flag2[1] = -1; }              // frames are fresh: nothing pending (ticks are >= 0)
        grid_barrier(a.barrier, gen);
        long long n_it = 0;
        gap = INFINITY;
        lr_known = false;
        bool have_sumK = false;
        while (true) {
            for (int sw = 0; sw < n_inner; ++sw) {
                ++tick;
                if (sw == 0 && lr_known) {
                    // Lr holds the row LSE at the current g (the gap check's pass): plain update of f
                    const bool pending = (*reinterpret_cast<volatile int*>(a.flag2 + ((tick - 1) & 1)) == tick - 1);
                    for (int64_t i = gtid; i < n; i += gsize) {
                        const double old = a.f[i];
                        double fr = a.u[i];
                        if (pending) { fr = old; a.u[i] = old; }
                        a.la_old[i] = (old - fr) / eps;
                        const double nv = eps * alpha1 * (a.logp[i] - (a.Lr[i] - a.log_m));
                        a.f[i] = nv;
                        const double b = SDB_LOG2E * (nv / eps - a.norms_x[i] * c1n);
                        a.bias_x[i] = (b > (double)SDB_NEG_SENTINEL) ? (float)b : SDB_NEG_SENTINEL;
                        const bool over = (nv - fr) / eps > a.log_tau;
                        if (over) atomicMax(a.flag2 + (tick & 1), tick);
                        if constexpr (STRIPS) tagx[i] = strip_word(b, seq_x + 1, over);
                    }
                    if constexpr (STRIPS) { ++seq_x; pend = pending; }
                } else {
                    if constexpr (STRIPS)
                        {
                            ++seq_x;
                            strip_pass<true>(srow, first_row, n_rows, (int)m, ldm, sc_hi, sc_lo, tagy, seq_y, sbias, a.Lr, a.logp, eps, alpha1,
                                             a.log_m, a.f, a.u, a.la_old, tagx, seq_x, a.flag2, tick, a.log_tau, fl_row_prev, true, pend, bits);
                        }
                    else if constexpr (RESIDENT)
                        res_pass<true, true>(a, smem, n_owned, row_tiles, col_tiles, sc_hi, sc_lo, a.bias_y, a.partial_row, cnt_row, a.Lr, a.norms_x, a.logp,
                                             eps, alpha1, a.log_m, a.f, a.u, a.la_old, a.bias_x, tick, red);
                    else
                    solve_pass<true>(a.row, direct, sc_hi, sc_lo, a.partial_row, a.ns_row, cnt_row, a.Lr, a.norms_x, c1n, a.logp, eps, alpha1,
                                     a.log_m, a.f, a.u, a.la_old, a.bias_x, a.flag2, tick, a.log_tau, smem);
                }
                // strips: the column pass waits for the rows' tagged entries themselves; only the vector-loop update above (it
                // wrote rows it does not own) needs the full barrier
                if (!STRIPS || (sw == 0 && lr_known)) grid_barrier(a.barrier, gen);
                if constexpr (STRIPS) {
                    ++seq_y;
                    bool pend_unused;
                    strip_pass<true>(scol, first_col, n_cols, (int)n, ldn, sc_hi, sc_lo, tagx, seq_x, sbias, a.Lc, a.logq, eps, alpha2, a.log_N,
                                     a.g, a.v, a.lb_old, tagy, seq_y, a.flag2, tick, a.log_tau, pend, false, pend_unused, bits);
                    fl_row_prev = bits;                                   // the tau bits of this tick's row updates
                } else if constexpr (RESIDENT)
                    res_pass<true, false>(a, smem, n_owned, row_tiles, col_tiles, sc_hi, sc_lo, a.bias_x, a.partial_col, cnt_col, a.Lc, a.norms_y, a.logq, eps,
                                          alpha2, a.log_N, a.g, a.v, a.lb_old, a.bias_y, tick, red);
                else
                solve_pass<true>(a.col, direct, sc_hi, sc_lo, a.partial_col, a.ns_col, cnt_col, a.Lc, a.norms_y, c1n, a.logq, eps, alpha2,
                                 a.log_N, a.g, a.v, a.lb_old, a.bias_y, a.flag2, tick, a.log_tau, smem);
                if (!STRIPS || sw == n_inner - 1) grid_barrier(a.barrier, gen);     // strips: the batch's checks read every row
            }
            n_it += n_inner;
            lr_known = false;
            // the last tick's absorption is still pending: apply it here, row by row, before the frames are read
            const bool pending = (*reinterpret_cast<volatile int*>(a.flag2 + (tick & 1)) == tick);
            if (final_stage) {
                if (!have_sumK) {
                    if constexpr (STRIPS)
                        {
                            bool p_unused, b_unused;
                            strip_pass<false>(srow, first_row, n_rows, (int)m, ldm, scd_hi, scd_lo, nullptr, 0u, sbias, a.Lr, nullptr, eps, 0.0,
                                              0.0, nullptr, nullptr, nullptr, nullptr, 0u, a.flag2, tick, a.log_tau, false, false, p_unused,
                                              b_unused);
                        }
                    else if constexpr (RESIDENT)
                        res_pass<false, true>(a, smem, n_owned, row_tiles, col_tiles, scd_hi, scd_lo, nullptr, a.partial_row, cnt_row, a.Lr, a.norms_x, nullptr, eps,
                                              0.0, 0.0, nullptr, nullptr, nullptr, nullptr, tick, red);
                    else
                    solve_pass<false>(row0, true, scd_hi, scd_lo, a.partial_row, a.ns_row, cnt_row, a.Lr, a.norms_x, 0.0, nullptr, eps, 0.0, 0.0,
                                      nullptr, nullptr, nullptr, nullptr, a.flag2, tick, a.log_tau, smem);
                    grid_barrier(a.barrier, gen);
                    double acc1[1] = {0.0}, out1[1];
                    for (int64_t i = gtid; i < n; i += gsize) acc1[0] += exp(a.Lr[i]);
                    solve_grid_sums<1>(acc1, a.scratch, a.barrier, gen, out1);
                    sumK = out1[0];
                    have_sumK = true;
                }
                // row LSE at the new g: the gap's row marginal, and the next iteration's row pass
                if constexpr (STRIPS)
                {
                    bool p_unused, b_unused;
                    strip_pass<false>(srow, first_row, n_rows, (int)m, ldm, sc_hi, sc_lo, tagy, seq_y, sbias, a.Lr, nullptr, eps, 0.0, 0.0,
                                      nullptr, nullptr, nullptr, nullptr, 0u, a.flag2, tick, a.log_tau, false, false, p_unused, b_unused);
                }
                else if constexpr (RESIDENT)
                    res_pass<false, true>(a, smem, n_owned, row_tiles, col_tiles, sc_hi, sc_lo, a.bias_y, a.partial_row, cnt_row, a.Lr, a.norms_x, nullptr, eps,
                                          0.0, 0.0, nullptr, nullptr, nullptr, nullptr, tick, red);
                else
                solve_pass<false>(a.row, direct, sc_hi, sc_lo, a.partial_row, a.ns_row, cnt_row, a.Lr, a.norms_x, c1n, nullptr, eps, 0.0, 0.0,
                                  nullptr, nullptr, nullptr, nullptr, a.flag2, tick, a.log_tau, smem);
                grid_barrier(a.barrier, gen);
                lr_known = true;
                double acc[8], t[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[k] = 0.0;
                for (int64_t i = gtid; i < n; i += gsize) {               // gap_terms_kernel (sdb_vectors.cu), ot_func.cpp:358-544
                    if (pending) a.u[i] = a.f[i];
                    const double fi = a.f[i], lp = a.logp[i];
                    const double R = exp(fi / eps + a.Lr[i]), p = exp(lp), r = R * a.dy;
                    acc[0] += R;
                    acc[1] += fi * R;
                    acc[2] += a.dx * (r * (log(r) - lp) - r + p);
                    acc[3] += p * a.dx * (exp(-fi / a.lambda1) - 1.0);
                }
                for (int64_t j = gtid; j < m; j += gsize) {
                    if (pending) a.v[j] = a.g[j];
                    const double gj = a.g[j], lq = a.logq[j];
                    const double R = exp(gj / eps + a.Lc[j]), q = exp(lq), c = R * a.dx;
                    acc[4] += R;
                    acc[5] += gj * R;
                    acc[6] += a.dy * (c * (log(c) - lq) - c + q);
                    acc[7] += q * a.dy * (exp(-gj / a.lambda2) - 1.0);
                }
                solve_grid_sums<8>(acc, a.scratch, a.barrier, gen, t);
                const double IJ = (double)(1.0 / a.dx) * (double)m;
                const double pri = a.lambda1 * t[2] + a.lambda2 * t[6] + (t[1] + t[5] - eps * t[0] + eps * sumK) / IJ;
                const double dua = -a.lambda1 * t[3] - a.lambda2 * t[7] - eps * (t[0] - sumK) / IJ;
                gap = (pri - dua) / fabs(pri);                            // ot_func.cpp:543
            } else {
                double acc[4] = {0.0, 0.0, 0.0, 0.0}, c[4];
                for (int64_t i = gtid; i < n; i += gsize) {               // stage_criterion_kernel, ot_func.cpp:878-922
                    if (pending) a.u[i] = a.f[i];
                    const double ui = a.u[i];
                    const double eu = exp(ui / eps);
                    const double at = exp((a.f[i] - ui) / eps) * eu;
                    const double df = at - exp(a.la_old[i]) * eu;
                    acc[0] += df * df;
                    acc[1] += at * at;
                }
                for (int64_t j = gtid; j < m; j += gsize) {
                    if (pending) a.v[j] = a.g[j];
                    const double vj = a.v[j];
                    const double ev = exp(vj / eps);
                    const double bt = exp((a.g[j] - vj) / eps) * ev;
                    const double df = bt - exp(a.lb_old[j]) * ev;
                    acc[2] += df * df;
                    acc[3] += bt * bt;
                }
                solve_grid_sums<4>(acc, a.scratch, a.barrier, gen, c);
                const double va = solve_safe_ratio(c[0], c[1]), vb = solve_safe_ratio(c[2], c[3]);
                gap = (vb > va) ? vb : va;                                // std::max(v1, v2) with its NaN behaviour
            }
            if (gap != gap) break;                                        // `while (nan > threshold)` is false
            if (!(gap > threshold)) break;
            if (n_it >= a.max_iter) { status = 2; break; }                // ot_func.cpp:821-824
        }
        total += (int)n_it;
        if (gtid == 0) { a.result->iters[e] = (int)n_it; a.result->stage_gap[e] = gap; }
        if (status == 2) { status = 0; if (gtid == 0) a.result->max_iter_reached = 1; }        // gives up on the stage, keeps going
        if (gap != gap && final_stage) status = 1;       // a NaN ends its stage (ot_func.cpp:866); only the last one is fatal (:446)
    }
    if (!lr_known && status != 1) {
        // row LSE at the final g (plan row sums, the growth loop's next G)
        // (bias_y still holds the last stage's form: reuse that stage's choice)
        const double c1 = a.inv_med / eps;
        const bool direct = RESIDENT || STRIPS || 2.0 * c1 * SDB_LOG2E * a.xy_max > a.dot_limit;
        const double sc = direct ? -c1 * SDB_LOG2E : 2.0 * c1 * SDB_LOG2E;
        const float sc_hi = (float)sc;
        if constexpr (STRIPS)
        {
            bool p_unused, b_unused;
            strip_pass<false>(srow, first_row, n_rows, (int)m, ldm, sc_hi, (float)(sc - (double)sc_hi), tagy, seq_y, sbias, a.Lr, nullptr, eps,
                              0.0, 0.0, nullptr, nullptr, nullptr, nullptr, 0u, a.flag2, tick, a.log_tau, false, false, p_unused, b_unused);
        }
        else if constexpr (RESIDENT)
            res_pass<false, true>(a, smem, n_owned, row_tiles, col_tiles, sc_hi, (float)(sc - (double)sc_hi), a.bias_y, a.partial_row, cnt_row,
                                  a.Lr, a.norms_x, nullptr, eps, 0.0, 0.0, nullptr, nullptr, nullptr, nullptr, tick, red);
        else
        solve_pass<false>(a.row, direct, sc_hi, (float)(sc - (double)sc_hi), a.partial_row, a.ns_row, cnt_row, a.Lr, a.norms_x,
                          direct ? 0.0 : c1, nullptr, eps, 0.0, 0.0, nullptr, nullptr, nullptr, nullptr, a.flag2, tick, a.log_tau, smem);
    }
    if (gtid == 0) {
        a.result->total_iters = total;
        a.result->status = status;
        a.result->gap = gap;
        a.result->eps_final = eps;
        a.result->last_tick = tick;
        if (a.flag2[tick & 1] == tick) *a.flag = tick;
    }
}

}  // namespace

extern "C" int sdb_sinkhorn_sweeps_persistent(const sdb_sweep_desc* d, int n_sweeps, int first_tick, int lr_known_first,
                                              unsigned int* barrier2, void* stream) {
    SDB_CHECK_ARG(d && barrier2 && n_sweeps >= 0 && d->n > 0 && d->m > 0 && d->eps > 0.0);
    if (d->use_tc) return SDB_E_UNSUPPORTED;
    if (d->dpad <= 0 || d->dpad > MAX_DPAD || (d->dpad & 3)) return SDB_E_UNSUPPORTED;
    SDB_CHECK_ARG(d->xt && d->yt && d->bounds_row && d->bounds_col && d->partial_row && d->partial_col && d->bias_x && d->bias_y);
    SDB_CHECK_ARG(d->f && d->g && d->u && d->v && d->la_old && d->lb_old && d->Lr && d->Lc && d->logp && d->logq && d->flag);
    SDB_CHECK_ARG(d->ns_row > 0 && d->ns_col > 0 && !(d->ldx & 3) && !(d->ldy & 3));
    if (n_sweeps == 0) return 0;
    const double c1 = d->inv_med / d->eps;
    PersistArgs a;
    a.row = PairArgs{d->xt, d->ldx, d->n, d->yt, d->ldy, d->m, d->dpad, d->bias_y, d->bounds_row};
    a.col = PairArgs{d->yt, d->ldy, d->m, d->xt, d->ldx, d->n, d->dpad, d->bias_x, d->bounds_col};
    a.ns_row = d->ns_row; a.ns_col = d->ns_col;
    a.direct = d->simt_direct ? 1 : 0;
    const double sc = a.direct ? -c1 * SDB_LOG2E : 2.0 * c1 * SDB_LOG2E;   // direct: t = bias_j - c1*log2(e)*|x_i - y_j|^2
    a.scale = (float)sc;
    a.scale_lo = (float)(sc - (double)a.scale);
    a.partial_row = reinterpret_cast<float2*>(d->partial_row); a.partial_col = reinterpret_cast<float2*>(d->partial_col);
    a.norms_x = d->norms_x; a.norms_y = d->norms_y; a.bias_x = d->bias_x; a.bias_y = d->bias_y;
    a.f = d->f; a.g = d->g; a.u = d->u; a.v = d->v; a.la_old = d->la_old; a.lb_old = d->lb_old; a.Lr = d->Lr; a.Lc = d->Lc;
    a.logp = d->logp; a.logq = d->logq; a.flag = d->flag;
    a.eps = d->eps; a.c1 = c1; a.alpha1 = d->alpha1; a.alpha2 = d->alpha2;
    a.log_m = log((double)d->m); a.log_N = log((double)d->n_total); a.log_tau = d->log_tau; a.log_floor = d->log_floor;
    a.n_sweeps = n_sweeps; a.first_tick = first_tick; a.lr_known_first = lr_known_first;
    a.barrier = barrier2;
    cudaStream_t st = sdb_stream(stream);
    if (!lr_known_first) {       // bias of the first row pass from the current g (a previous call may have used another eps)
        int rc = sdb_make_bias(d->m, d->m_bias, d->g, d->norms_y, d->eps, c1, d->bias_y, stream);
        if (rc) return rc;
    }
    const size_t smem = sizeof(float) * ((size_t)d->dpad * BM + 2 * (size_t)d->dpad * BN + 2 * BN);
    static size_t smem_set[SDB_MAX_DEVICES] = {0};
    { cudaError_t e0 = sdb_ensure_smem(sinkhorn_persistent_kernel, smem, smem_set); if (e0 != cudaSuccess) return (int)e0; }
    int n_sm = 0, per_sm = 0;
    {
        int dev = 0;
        cudaError_t e1 = cudaGetDevice(&dev);
        if (e1 == cudaSuccess) e1 = cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        if (e1 != cudaSuccess) return (int)e1;
    }
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sinkhorn_persistent_kernel, NT, smem);
    if (e != cudaSuccess) return (int)e;
    if (per_sm < 1) return SDB_E_UNSUPPORTED;
    // grid: the caller's choice (d->n_ctas, it sized the splits for it), capped by what can be co-resident
    const int64_t row_items = ((d->n + BM - 1) / BM) * d->ns_row, col_items = ((d->m + BM - 1) / BM) * d->ns_col;
    int64_t want = d->n_ctas > 0 ? d->n_ctas : (row_items > col_items ? row_items : col_items);
    const int64_t cap = (int64_t)n_sm * per_sm;
    int grid = (int)(want < cap ? want : cap);
    if (grid < 1) grid = 1;
    // a launch that died inside a barrier must not poison the next one: start every launch from a clean counter pair
    e = cudaMemsetAsync(barrier2, 0, SDB_BARRIER_WORDS * sizeof(unsigned int), st);
    if (e != cudaSuccess) return (int)e;
    void* params[] = {&a};
    e = cudaLaunchCooperativeKernel((const void*)sinkhorn_persistent_kernel, dim3((unsigned)grid), dim3(NT), params, smem, st);
    return (int)e;
}

extern "C" int sdb_sinkhorn_solve_persistent(const sdb_sweep_desc* d, const sdb_solve_params* p, int first_tick, int* flag2,
                                             unsigned int* barrier2, unsigned int* counters, double* scratch,
                                             sdb_solve_result* result, void* stream) {
    SDB_CHECK_ARG(d && p && flag2 && barrier2 && counters && scratch && result && d->n > 0 && d->m > 0);
    if (d->use_tc) return SDB_E_UNSUPPORTED;
    if (d->dpad <= 0 || d->dpad > MAX_DPAD || (d->dpad & 3)) return SDB_E_UNSUPPORTED;
    SDB_CHECK_ARG(d->xt && d->yt && d->bounds_row && d->bounds_col && d->partial_row && d->partial_col && d->bias_x && d->bias_y);
    SDB_CHECK_ARG(d->f && d->g && d->u && d->v && d->la_old && d->lb_old && d->Lr && d->Lc && d->logp && d->logq && d->flag);
    SDB_CHECK_ARG(d->ns_row > 0 && d->ns_col > 0 && !(d->ldx & 3) && !(d->ldy & 3));
    SDB_CHECK_ARG(p->epsilon > 0.0 && p->epsilon0 > 0.0 && p->batch_size > 0 && p->tau > 0.0 && p->max_iter > 0);
    SolveArgs a;
    a.row = PairArgs{d->xt, d->ldx, d->n, d->yt, d->ldy, d->m, d->dpad, d->bias_y, d->bounds_row};
    a.col = PairArgs{d->yt, d->ldy, d->m, d->xt, d->ldx, d->n, d->dpad, d->bias_x, d->bounds_col};
    a.ns_row = d->ns_row; a.ns_col = d->ns_col;
    a.partial_row = reinterpret_cast<float2*>(d->partial_row); a.partial_col = reinterpret_cast<float2*>(d->partial_col);
    a.bias_x = d->bias_x; a.bias_y = d->bias_y;
    a.f = d->f; a.g = d->g; a.u = d->u; a.v = d->v; a.la_old = d->la_old; a.lb_old = d->lb_old; a.Lr = d->Lr; a.Lc = d->Lc;
    a.logp = d->logp; a.logq = d->logq; a.flag = d->flag; a.flag2 = flag2;
    SDB_CHECK_ARG(d->norms_x && d->norms_y);
    a.norms_x = d->norms_x; a.norms_y = d->norms_y;
    for (int k = 0; k < 6; ++k) { SDB_CHECK_ARG(p->eps_stage[k] > 0.0); a.eps_stage[k] = p->eps_stage[k]; }
    a.xy_max = p->xy_max; a.dot_limit = p->dot_limit;
    a.inv_med = d->inv_med; a.lambda1 = p->lambda1; a.lambda2 = p->lambda2; a.epsilon = p->epsilon; a.epsilon0 = p->epsilon0;
    a.tolerance = p->tolerance; a.log_tau = log(p->tau);
    a.log_m = log((double)d->m); a.log_N = log((double)d->n_total); a.dx = 1.0 / (double)d->n_total; a.dy = 1.0 / (double)d->m;
    a.batch_size = p->batch_size; a.max_iter = (long long)p->max_iter; a.first_tick = first_tick;
    a.barrier = barrier2; a.counters = counters; a.scratch = scratch; a.result = result;
    cudaStream_t st = sdb_stream(stream);
    int dev = 0, n_sm = 0, per_sm = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return (int)e;
    const int64_t R = (d->n + BM - 1) / BM, C = (d->m + BN - 1) / BN;
    const bool resident = p->reserved == 1, strips = p->reserved == 2;
    size_t smem = 0;
    int grid = 0;
    a.res_tiles = 0;
    a.strip_rows = a.strip_cols = 0;
    if (strips) {
        // one CTA per SM; every CTA keeps ceil(n/grid) full rows and ceil(m/grid) full columns of D plus one bias vector
        static size_t smem_set_s[SDB_MAX_DEVICES] = {0};
        int64_t g = n_sm;                               // every CTA must own at least one row and one column (dataflow protocol)
        if (g > d->n) g = d->n;
        if (g > d->m) g = d->m;
        const int64_t rp = (d->n + g - 1) / g, cp = (d->m + g - 1) / g;
        const int64_t ldm = (d->m + 3) & ~(int64_t)3, ldn = (d->n + 3) & ~(int64_t)3;
        // the last buffer stages the other side's bias in every pass and, once, the owned points' coordinates
        int64_t stage = ldm > ldn ? ldm : ldn;
        if ((rp > cp ? rp : cp) * d->dpad > stage) stage = (rp > cp ? rp : cp) * d->dpad;
        const size_t need = sizeof(float) * (size_t)(rp * ldm + cp * ldn + stage);
        if (need > SDB_STRIP_SMEM_MAX) return SDB_E_UNSUPPORTED;
        cudaError_t e0 = sdb_ensure_smem(sinkhorn_solve_kernel<2>, need, smem_set_s);
        if (e0 != cudaSuccess) return (int)e0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sinkhorn_solve_kernel<2>, STRIP_NT, need);
        if (e != cudaSuccess) return (int)e;
        if (per_sm < 1) return SDB_E_UNSUPPORTED;
        grid = (int)g; smem = need; a.strip_rows = (int)rp; a.strip_cols = (int)cp;
    } else if (resident) {
        // resident cost tiles: the caller sized the partial buffers for one split per tile (ns_row = C, ns_col = R); find a
        // co-resident grid whose CTAs can hold their share of the R*C tiles (16 KB each) in shared memory
        SDB_CHECK_ARG(d->ns_row == C && d->ns_col == R);
        static size_t smem_set_r[SDB_MAX_DEVICES] = {0};
        for (int want_per_sm = 2; want_per_sm >= 1 && grid == 0; --want_per_sm) {
            const int64_t g = (int64_t)n_sm * want_per_sm;
            const int64_t T = (R * C + g - 1) / g;
            if (T > RES_MAX_TILES) continue;
            const size_t need = (size_t)T * 4 * NT * 4 * sizeof(float) + (size_t)T * (NT / 32) * BN * sizeof(float2);
            if (need > 220 * 1024) continue;
            cudaError_t e0 = sdb_ensure_smem(sinkhorn_solve_kernel<1>, need, smem_set_r);
            if (e0 != cudaSuccess) return (int)e0;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sinkhorn_solve_kernel<1>, NT, need);
            if (e != cudaSuccess) return (int)e;
            if (per_sm >= want_per_sm) { grid = (int)g; smem = need; a.res_tiles = (int)T; }
        }
        if (grid == 0) return SDB_E_UNSUPPORTED;      // too many tiles for the on-chip memory: the caller takes the streamed form
    } else {
        smem = sizeof(float) * ((size_t)d->dpad * BM + 2 * (size_t)d->dpad * BN + 2 * BN);
        static size_t smem_set[SDB_MAX_DEVICES] = {0};
        { cudaError_t e0 = sdb_ensure_smem(sinkhorn_solve_kernel<0>, smem, smem_set); if (e0 != cudaSuccess) return (int)e0; }
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sinkhorn_solve_kernel<0>, NT, smem);
        if (e != cudaSuccess) return (int)e;
        if (per_sm < 1) return SDB_E_UNSUPPORTED;
        const int64_t row_items = R * d->ns_row, col_items = C * d->ns_col;
        int64_t want = d->n_ctas > 0 ? d->n_ctas : (row_items > col_items ? row_items : col_items);
        const int64_t cap = (int64_t)n_sm * per_sm;
        grid = (int)(want < cap ? want : cap);
        if (grid < 1) grid = 1;
    }
    if (grid > SDB_SOLVE_MAX_CTAS) grid = SDB_SOLVE_MAX_CTAS;
    e = cudaMemsetAsync(barrier2, 0, SDB_BARRIER_WORDS * sizeof(unsigned int), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(result, 0, sizeof(sdb_solve_result), st);
    // strip form: the tagged bias vectors live behind the per-CTA sums in `scratch`; production number 0 = "nothing yet"
    if (e == cudaSuccess && strips)
        e = cudaMemsetAsync(scratch + (size_t)SDB_SOLVE_MAX_CTAS * 10, 0, sizeof(double) * (size_t)(d->n + d->m), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(counters, 0, sizeof(unsigned int) * (size_t)((d->n + BM - 1) / BM + (d->m + BM - 1) / BM), st);
    if (e != cudaSuccess) return (int)e;
    void* params[] = {&a};
    e = cudaLaunchCooperativeKernel(strips ? (const void*)sinkhorn_solve_kernel<2>
                                           : resident ? (const void*)sinkhorn_solve_kernel<1> : (const void*)sinkhorn_solve_kernel<0>,
                                    dim3((unsigned)grid), dim3(strips ? STRIP_NT : NT), params, smem, st);
    return (int)e;
}

extern "C" int sdb_lse_pass_simt(const float* pt, int64_t ldp, int64_t n_p, const float* qt, int64_t ldq, int64_t n_q,
                                 int dpad, const float* bias, double scale, const int64_t* split_bounds, int n_splits,
                                 float* partial, void* stream) {
    SDB_CHECK_ARG(bias && partial);
    PairArgs a{pt, ldp, n_p, qt, ldq, n_q, dpad, bias, split_bounds};
    const float s_hi = (float)scale;
    LseEpi::Params ep{s_hi, (float)(scale - (double)s_hi), reinterpret_cast<float2*>(partial)};
    return scale < 0.0 ? launch_pairs<true, LseEpi>(a, ep, n_splits, sdb_stream(stream))
                       : launch_pairs<false, LseEpi>(a, ep, n_splits, sdb_stream(stream));
}

extern "C" int sdb_cost_histogram(const float* pt, int64_t ldp, int64_t n_p, const float* qt, int64_t ldq, int64_t n_q,
                                  int dpad, const int64_t* split_bounds, int n_splits, float lo, float hi, int n_bins,
                                  unsigned long long* hist, unsigned long long* counts2, void* stream) {
    SDB_CHECK_ARG(hist && counts2 && n_bins > 0 && hi > lo);
    PairArgs a{pt, ldp, n_p, qt, ldq, n_q, dpad, nullptr, split_bounds};
    HistEpi::Params ep{lo, hi, (float)n_bins / (hi - lo), n_bins, hist, counts2};
    return launch_pairs<true, HistEpi>(a, ep, n_splits, sdb_stream(stream));
}

extern "C" int sdb_cost_collect(const float* pt, int64_t ldp, int64_t n_p, const float* qt, int64_t ldq, int64_t n_q,
                                int dpad, const int64_t* split_bounds, int n_splits, float lo, float hi,
                                const double* x, const double* y, int d, double* cand, unsigned long long cap,
                                unsigned long long* counts2, void* stream) {
    SDB_CHECK_ARG(x && y && cand && counts2 && d > 0);
    PairArgs a{pt, ldp, n_p, qt, ldq, n_q, dpad, nullptr, split_bounds};
    CollectEpi::Params ep{lo, hi, x, y, d, cand, cap, counts2};
    return launch_pairs<true, CollectEpi>(a, ep, n_splits, sdb_stream(stream));
}
