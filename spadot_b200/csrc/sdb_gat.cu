// K2 / K2b: fused GAT edge-softmax + aggregation, forward and backward, over a CSR-by-destination graph.
//
// Restates what torch_geometric.nn.GATConv does between its linear layer and its bias add
// (used at ref: SpaDOT/model/encoder.py:41-45,56-58): per edge j->i and head h
//     e_ij = leaky_relu(a_src[j,h] + a_dst[i,h], slope);  alpha_ij = softmax_j(e_ij) (+1e-16 in the denominator)
//     out[i,h,:] = sum_j alpha_ij * feat[j,h,:]
// PyG materialises (E,H) logits, (E,H,C) messages and scatters them; here one CTA per destination node
// does the segment softmax in registers/warp shuffles and gathers each source row exactly once
// (coalesced, 16 B per thread where C allows), writing only out (n,H,C) and alpha (E,H).  The
// backward uses the by-source CSR so there are no atomics and results are deterministic.
// HBM/L2 gather-bound:  bytes ~ E*H*C*s (row gathers) + n*H*C*s (write) + E*(4 + 2*H*s).
#include "sdb_common.cuh"
#include <stdlib.h>

namespace {

template <typename T> __device__ __forceinline__ T t_exp(T x);
template <> __device__ __forceinline__ float t_exp<float>(float x) { return __expf(x); }
template <> __device__ __forceinline__ double t_exp<double>(double x) { return exp(x); }

template <typename T>
__device__ __forceinline__ T warp_max(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
template <typename T>
__device__ __forceinline__ T warp_sum_t(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// reductions over aligned groups of eight lanes (all 32 lanes of the warp take part in the shuffles)
template <typename T>
__device__ __forceinline__ T oct_max(T v) {
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
template <typename T>
__device__ __forceinline__ T oct_sum(T v) {
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// 16-byte vector access: 4 floats or 2 doubles per thread per load (used when C is a multiple of the width)
template <typename T> struct Vec;
template <> struct Vec<float> { static constexpr int W = 4; using type = float4; };
template <> struct Vec<double> { static constexpr int W = 2; using type = double2; };
template <typename T>
__device__ __forceinline__ void vec_load(const T* p, T (&v)[Vec<T>::W]) {
    const typename Vec<T>::type t = *reinterpret_cast<const typename Vec<T>::type*>(p);
    const T* e = reinterpret_cast<const T*>(&t);
#pragma unroll
    for (int q = 0; q < Vec<T>::W; ++q) v[q] = e[q];
}
template <typename T>
__device__ __forceinline__ void vec_store(T* p, const T (&v)[Vec<T>::W]) {
    typename Vec<T>::type t;
    T* e = reinterpret_cast<T*>(&t);
#pragma unroll
    for (int q = 0; q < Vec<T>::W; ++q) e[q] = v[q];
    *reinterpret_cast<typename Vec<T>::type*>(p) = t;
}

// ======================================================================================== per-node form
// One node at a time, every thread of the CTA cooperating.  This is the general form (any H, any C, any degree); the tile
// kernels below call it for the tiles that exceed their shared-memory budget.  Must be called by ALL threads of the CTA.

// SH ("shared features", the aggregate-first layers of spadot_b200/gat.py): the gathered rows are (n, C) and every head reads
// the SAME row - out[i,h,:] = sum_j alpha^h_ij x_j - instead of its own slice of an (n, H, C) row.

// forward: segment softmax of node i (one warp per head), then the weighted sum of its gathered source rows
template <typename T, bool SH>
__device__ void gat_fwd_node(const T* __restrict__ feat, const T* __restrict__ a_src, const T* __restrict__ a_dst,
                             const int32_t* __restrict__ col, int64_t i, int64_t e0, int64_t e1, int H, int C, T slope,
                             T* __restrict__ out, T* __restrict__ alpha) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5, nt = blockDim.x;
    for (int h = warp; h < H; h += nw) {
        const T ad = a_dst[i * H + h];
        T mx = -INFINITY;
        for (int64_t e = e0 + lane; e < e1; e += 32) {
            T r = a_src[(int64_t)col[e] * H + h] + ad;
            r = r > T(0) ? r : r * slope;
            mx = max(mx, r);
        }
        mx = warp_max(mx);
        T sum = T(0);
        for (int64_t e = e0 + lane; e < e1; e += 32) {
            T r = a_src[(int64_t)col[e] * H + h] + ad;
            r = r > T(0) ? r : r * slope;
            sum += t_exp(r - mx);
        }
        sum = warp_sum_t(sum) + T(1e-16);
        for (int64_t e = e0 + lane; e < e1; e += 32) {
            T r = a_src[(int64_t)col[e] * H + h] + ad;
            r = r > T(0) ? r : r * slope;
            alpha[e * H + h] = t_exp(r - mx) / sum;
        }
    }
    __syncthreads();
    // aggregation: thread owns feature columns c, c+nt, ...; each source row is read once, coalesced
    const int HC = H * C;
    constexpr int W = Vec<T>::W;
    if (C % W == 0) {
        for (int c = threadIdx.x * W; c < HC; c += nt * W) {
            const int h = c / C;
            T acc[W];
#pragma unroll
            for (int q = 0; q < W; ++q) acc[q] = T(0);
            const int fc = SH ? c - h * C : c;
            for (int64_t e = e0; e < e1; ++e) {
                const T a = alpha[e * H + h];
                T v[W];
                vec_load(feat + (int64_t)col[e] * (SH ? C : HC) + fc, v);
#pragma unroll
                for (int q = 0; q < W; ++q) acc[q] += a * v[q];
            }
            vec_store(out + i * HC + c, acc);
        }
    } else {
        for (int c = threadIdx.x; c < HC; c += nt) {
            const int h = c / C;
            T acc = T(0);
            for (int64_t e = e0; e < e1; ++e) acc += alpha[e * H + h] * feat[(int64_t)col[e] * (SH ? C : HC) + (SH ? c - h * C : c)];
            out[i * HC + c] = acc;
        }
    }
    __syncthreads();
}

// backward, by destination:  dlogit[e,h] = d loss / d (a_src[j,h] + a_dst[i,h]);  grad_a_dst[i,h] = sum_e dlogit[e,h]
template <typename T, bool SH>
__device__ void gat_bdst_node(const T* __restrict__ feat, const T* __restrict__ a_src, const T* __restrict__ a_dst,
                              const int32_t* __restrict__ col, int64_t i, int64_t e0, int64_t e1, int H, int C, T slope,
                              const T* __restrict__ alpha, const T* __restrict__ grad_out, T* __restrict__ dlogit,
                              T* __restrict__ grad_a_dst) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int HC = H * C;
    const int64_t n_eh = (e1 - e0) * H;
    // d alpha[e,h] = <grad_out[i,h,:], feat[j,h,:]>  (one warp per (edge, head), coalesced over C)
    for (int64_t p = warp; p < n_eh; p += nw) {
        const int64_t e = e0 + p / H;
        const int h = (int)(p % H);
        const T* go = grad_out + i * HC + h * C;
        const T* fj = SH ? feat + (int64_t)col[e] * C : feat + (int64_t)col[e] * HC + h * C;
        T acc = T(0);
        constexpr int W = Vec<T>::W;
        if (C % W == 0) {
            for (int c = lane * W; c < C; c += 32 * W) {
                T g[W], f[W];
                vec_load(go + c, g);
                vec_load(fj + c, f);
#pragma unroll
                for (int q = 0; q < W; ++q) acc += g[q] * f[q];
            }
        } else {
            for (int c = lane; c < C; c += 32) acc += go[c] * fj[c];
        }
        acc = warp_sum_t(acc);
        if (lane == 0) dlogit[e * H + h] = acc;
    }
    __syncthreads();
    for (int h = warp; h < H; h += nw) {
        T s = T(0);
        for (int64_t e = e0 + lane; e < e1; e += 32) s += alpha[e * H + h] * dlogit[e * H + h];
        s = warp_sum_t(s);
        const T ad = a_dst[i * H + h];
        T gd = T(0);
        for (int64_t e = e0 + lane; e < e1; e += 32) {
            const T r = a_src[(int64_t)col[e] * H + h] + ad;
            T g = alpha[e * H + h] * (dlogit[e * H + h] - s);
            g = r > T(0) ? g : g * slope;
            dlogit[e * H + h] = g;
            gd += g;
        }
        gd = warp_sum_t(gd);
        if (lane == 0) grad_a_dst[i * H + h] = gd;
    }
    __syncthreads();
}

// backward, by source:  grad_feat[j,h,:] = sum_{e: j->i} alpha[e,h] * grad_out[i,h,:];  grad_a_src[j,h] = sum_e dlogit[e,h]
// prefix form: only the edges into the first n_dst destinations exist for this layer; in the by-destination order they
// are the edge ids below rowptr[n_dst] (e_limit), so the full by-source lists are filtered instead of rebuilt
template <typename T>
__device__ void gat_bsrc_node(const int32_t* __restrict__ src_dst, const int32_t* __restrict__ src_eid, int64_t j, int64_t p0,
                              int64_t p1, int64_t e_limit, int H, int C, const T* __restrict__ alpha,
                              const T* __restrict__ dlogit, const T* __restrict__ grad_out, T* __restrict__ grad_feat,
                              T* __restrict__ grad_a_src) {
    const int HC = H * C, nt = blockDim.x;
    constexpr int W = Vec<T>::W;
    if (C % W == 0) {
        for (int c = threadIdx.x * W; c < HC; c += nt * W) {
            const int h = c / C;
            T acc[W];
#pragma unroll
            for (int q = 0; q < W; ++q) acc[q] = T(0);
            for (int64_t p = p0; p < p1; ++p) {
                if (src_eid[p] >= e_limit) continue;
                const T a = alpha[(int64_t)src_eid[p] * H + h];
                T v[W];
                vec_load(grad_out + (int64_t)src_dst[p] * HC + c, v);
#pragma unroll
                for (int q = 0; q < W; ++q) acc[q] += a * v[q];
            }
            vec_store(grad_feat + j * HC + c, acc);
        }
    } else {
        for (int c = threadIdx.x; c < HC; c += nt) {
            const int h = c / C;
            T acc = T(0);
            for (int64_t p = p0; p < p1; ++p)
                if (src_eid[p] < e_limit) acc += alpha[(int64_t)src_eid[p] * H + h] * grad_out[(int64_t)src_dst[p] * HC + c];
            grad_feat[j * HC + c] = acc;
        }
    }
    for (int h = threadIdx.x; h < H; h += nt) {
        T acc = T(0);
        for (int64_t p = p0; p < p1; ++p)
            if (src_eid[p] < e_limit) acc += dlogit[(int64_t)src_eid[p] * H + h];
        grad_a_src[j * H + h] = acc;
    }
}

// by-source backward with shared features: grad_x[j,:] = sum over heads and out-edges of alpha[e,h] * grad_out[i,h,:]  (n_src, C)
template <typename T>
__device__ void gat_bsrc_node_shared(const int32_t* __restrict__ src_dst, const int32_t* __restrict__ src_eid, int64_t j, int64_t p0,
                                     int64_t p1, int64_t e_limit, int H, int C, const T* __restrict__ alpha,
                                     const T* __restrict__ dlogit, const T* __restrict__ grad_out, T* __restrict__ grad_feat,
                                     T* __restrict__ grad_a_src) {
    const int HC = H * C, nt = blockDim.x;
    constexpr int W = Vec<T>::W;
    if (C % W == 0) {
        for (int c = threadIdx.x * W; c < C; c += nt * W) {
            T acc[W];
#pragma unroll
            for (int q = 0; q < W; ++q) acc[q] = T(0);
            for (int64_t p = p0; p < p1; ++p) {
                if (src_eid[p] >= e_limit) continue;
                for (int h = 0; h < H; ++h) {
                    const T a = alpha[(int64_t)src_eid[p] * H + h];
                    T v[W];
                    vec_load(grad_out + (int64_t)src_dst[p] * HC + h * C + c, v);
#pragma unroll
                    for (int q = 0; q < W; ++q) acc[q] += a * v[q];
                }
            }
            vec_store(grad_feat + j * C + c, acc);
        }
    } else {
        for (int c = threadIdx.x; c < C; c += nt) {
            T acc = T(0);
            for (int64_t p = p0; p < p1; ++p)
                if (src_eid[p] < e_limit)
                    for (int h = 0; h < H; ++h) acc += alpha[(int64_t)src_eid[p] * H + h] * grad_out[(int64_t)src_dst[p] * HC + h * C + c];
            grad_feat[j * C + c] = acc;
        }
    }
    for (int h = threadIdx.x; h < H; h += nt) {
        T acc = T(0);
        for (int64_t p = p0; p < p1; ++p)
            if (src_eid[p] < e_limit) acc += dlogit[(int64_t)src_eid[p] * H + h];
        grad_a_src[j * H + h] = acc;
    }
}

template <typename T, bool SH>
__global__ void __launch_bounds__(128) gat_fwd_kernel(const T* __restrict__ feat, const T* __restrict__ a_src,
                                                      const T* __restrict__ a_dst, const int64_t* __restrict__ rowptr,
                                                      const int32_t* __restrict__ col, const int32_t* __restrict__ order,
                                                      int64_t n, int H, int C, T slope, T* __restrict__ out,
                                                      T* __restrict__ alpha) {
    const int64_t i = order ? order[blockIdx.x] : blockIdx.x;   // locality order of the CTAs (L2 reuse of gathered rows)
    gat_fwd_node<T, SH>(feat, a_src, a_dst, col, i, rowptr[i], rowptr[i + 1], H, C, slope, out, alpha);
}

template <typename T, bool SH>
__global__ void __launch_bounds__(128) gat_bwd_dst_kernel(const T* __restrict__ feat, const T* __restrict__ a_src,
                                                          const T* __restrict__ a_dst, const int64_t* __restrict__ rowptr,
                                                          const int32_t* __restrict__ col, const int32_t* __restrict__ order,
                                                          int64_t n, int H, int C, T slope, const T* __restrict__ alpha,
                                                          const T* __restrict__ grad_out, T* __restrict__ dlogit,
                                                          T* __restrict__ grad_a_dst) {
    const int64_t i = order ? order[blockIdx.x] : blockIdx.x;
    gat_bdst_node<T, SH>(feat, a_src, a_dst, col, i, rowptr[i], rowptr[i + 1], H, C, slope, alpha, grad_out, dlogit, grad_a_dst);
}

template <typename T, bool SH>
__global__ void __launch_bounds__(128) gat_bwd_src_kernel(const int64_t* __restrict__ src_rowptr, const int32_t* __restrict__ src_dst,
                                                          const int32_t* __restrict__ src_eid, const int32_t* __restrict__ order,
                                                          const int64_t* __restrict__ e_limit_ptr, int H, int C,
                                                          const T* __restrict__ alpha,
                                                          const T* __restrict__ dlogit, const T* __restrict__ grad_out,
                                                          T* __restrict__ grad_feat, T* __restrict__ grad_a_src) {
    const int64_t j = order ? order[blockIdx.x] : blockIdx.x;
    if constexpr (SH)
        gat_bsrc_node_shared<T>(src_dst, src_eid, j, src_rowptr[j], src_rowptr[j + 1], *e_limit_ptr, H, C, alpha, dlogit, grad_out,
                                grad_feat, grad_a_src);
    else
        gat_bsrc_node<T>(src_dst, src_eid, j, src_rowptr[j], src_rowptr[j + 1], *e_limit_ptr, H, C, alpha, dlogit, grad_out, grad_feat,
                         grad_a_src);
}

// ======================================================================================== tile form
// GT_TD consecutive nodes of the locality order per CTA.  Neighbouring spots share most of their neighbours (k = 30 nearest
// + self: eight Z-curve neighbours reference 50-90 DISTINCT sources through their ~250 edges), so the tile first
// de-duplicates the rows it gathers (shared-memory hash set, then a rank sort so that the order - and with it every
// floating-point sum - is reproducible), computes every edge logit ONCE, keeps the attention weights in shared memory as
// a small dense (distinct row) x (tile member) matrix per head, and then streams every distinct row exactly once:
//     forward / by-source backward:  acc[member][cols] += w[h][row][member] * row[cols]      (thread = column group)
//     by-destination backward:       d alpha[row][member] = <grad_out[member], row> per head  (thread = column group,
//                                    eight partial sums per lane reduced across the warp with a 9-shuffle butterfly)
// That cuts the L2 gather traffic of the per-node form (E*H*C*s bytes) by the sharing factor (3-5x) and turns the
// per-edge dependent loads into independent, unrollable ones.  The weights that are zero (a row some member does not
// read) are multiplied like the others: a dense 8-wide update is cheaper than a branch per (row, member).
// A tile whose edges or distinct rows do not fit (hubs) runs the per-node form above, member by member.
constexpr int GT_TD = 8;          // tile members (destinations; sources in the by-source pass)
constexpr int GT_EMAX = 512;      // edges of a tile held in shared memory
constexpr int GT_UMAX = 256;      // distinct gathered rows of a tile (8 x 32 edges fit even when nothing is shared)
constexpr int GT_UCHUNK = 128;    // by-destination backward: rows per pass through the per-warp partial sums
constexpr int GT_HMAX = 4;        // heads (the reference's num_heads; more heads run the per-node kernels)
constexpr int GT_HASH = 1024;     // hash-set slots (>= 2 * GT_EMAX: the probe loop always ends)
constexpr int GT_THREADS = 256;

template <typename T, int LEAD, int UROWS>
struct GatTile {
    T w[LEAD][UROWS][GT_TD];      // (heads, GT_UMAX): dense attention weights;  (warps, GT_UCHUNK): per-warp partial dot products
    T ev[GT_EMAX][GT_HMAX];       // per edge and head: logit -> alpha (forward), alpha (by source), d alpha (by destination)
    int32_t tab[GT_HASH];
    int32_t e_nbr[GT_EMAX];       // the row an edge gathers (-1: filtered out by the prefix form)
    int32_t e_gid[GT_EMAX];       // position of the edge in the by-destination order (index into alpha / dlogit)
    int32_t u_tmp[GT_UMAX];
    int32_t u_id[GT_UMAX];        // distinct rows, ascending
    uint16_t e_loc[GT_EMAX];      // index of e_nbr in u_id
    uint8_t e_own[GT_EMAX];       // tile member the edge belongs to
    int64_t node[GT_TD];
    int64_t p0[GT_TD];
    int len[GT_TD];
    int off[GT_TD + 1];
    int n_u;
};

// Phase A of every tile kernel: members, their edges, the distinct rows.  Returns false (uniformly) when the tile does not fit.
template <typename S>
__device__ bool gat_tile_edges(S& sm, const int64_t* __restrict__ ptr, const int32_t* __restrict__ nbr,
                               const int32_t* __restrict__ eid, int64_t e_limit, const int32_t* __restrict__ order, int64_t n) {
    const int tid = threadIdx.x;
    const int64_t base = (int64_t)blockIdx.x * GT_TD;
    if (tid < GT_TD) {
        int64_t node = -1, a = 0, b = 0;
        if (base + tid < n) {
            node = order ? (int64_t)order[base + tid] : base + tid;
            a = ptr[node];
            b = ptr[node + 1];
        }
        sm.node[tid] = node;
        sm.p0[tid] = a;
        sm.len[tid] = (b - a > (int64_t)GT_EMAX) ? GT_EMAX + 1 : (int)(b - a);
    }
    for (int i = tid; i < GT_HASH; i += GT_THREADS) sm.tab[i] = -1;
    if (tid == 0) sm.n_u = 0;
    __syncthreads();
    if (tid == 0) {
        int acc = 0;
#pragma unroll
        for (int m = 0; m < GT_TD; ++m) { sm.off[m] = acc; acc += sm.len[m]; }
        sm.off[GT_TD] = acc;
    }
    __syncthreads();
    const int Et = sm.off[GT_TD];
    if (Et > GT_EMAX) return false;
    for (int s = tid; s < Et; s += GT_THREADS) {
        int m = 0;
#pragma unroll
        for (int k = 1; k < GT_TD; ++k) m += (s >= sm.off[k]) ? 1 : 0;
        const int64_t p = sm.p0[m] + (s - sm.off[m]);
        const int32_t g = eid ? eid[p] : (int32_t)p;
        int32_t v = nbr[p];
        if ((int64_t)g >= e_limit) v = -1;
        sm.e_nbr[s] = v;
        sm.e_gid[s] = g;
        sm.e_own[s] = (uint8_t)m;
        if (v >= 0) {
            unsigned slot = ((unsigned)v * 2654435761u) >> 22;          // 10 bits
            while (true) {
                const int32_t old = atomicCAS(&sm.tab[slot], -1, v);
                if (old == -1) {
                    const int idx = atomicAdd(&sm.n_u, 1);
                    if (idx < GT_UMAX) sm.u_tmp[idx] = v;
                    break;
                }
                if (old == v) break;
                slot = (slot + 1) & (GT_HASH - 1);
            }
        }
    }
    __syncthreads();
    const int U = sm.n_u;
    if (U > GT_UMAX) return false;
    for (int u = tid; u < U; u += GT_THREADS) {                         // rank sort: distinct keys, so ranks are a permutation
        const int32_t my = sm.u_tmp[u];
        int r = 0;
        for (int k = 0; k < U; ++k) r += (sm.u_tmp[k] < my) ? 1 : 0;
        sm.u_id[r] = my;
    }
    __syncthreads();
    for (int s = tid; s < Et; s += GT_THREADS) {
        const int32_t v = sm.e_nbr[s];
        if (v < 0) continue;
        int lo = 0, hi = U;                                             // lower bound of v in u_id[0, U): v is present
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (sm.u_id[mid] <= v) lo = mid; else hi = mid;
        }
        sm.e_loc[s] = (uint16_t)lo;
    }
    return true;                                                       // the caller synchronises before reading e_loc
}

// Phase B of the forward and of the by-source backward: out[member] = sum over distinct rows of w[h][row][member] * rows[row].
template <typename T, int W, typename S>
__device__ __forceinline__ void gat_tile_aggregate(const S& sm, const T* __restrict__ rows, T* __restrict__ out, int U, int H, int C) {
    const int HC = H * C;
    for (int c = (blockIdx.y * GT_THREADS + threadIdx.x) * W; c < HC; c += gridDim.y * GT_THREADS * W) {
        const int h = c / C;
        T acc[GT_TD][W];
#pragma unroll
        for (int t = 0; t < GT_TD; ++t)
#pragma unroll
            for (int q = 0; q < W; ++q) acc[t][q] = T(0);
        // eight independent row loads in flight per thread (the batch is hoisted above the FMAs); a software prefetch into L1
        // (prefetch.global.L1, distance 8 / 16) and three CTAs per SM at unroll 4 were measured and change nothing (+-3 %)
#pragma unroll 8
        for (int u = 0; u < U; ++u) {
            T v[W];
            if constexpr (W == 1) v[0] = rows[(int64_t)sm.u_id[u] * HC + c];
            else vec_load(rows + (int64_t)sm.u_id[u] * HC + c, *reinterpret_cast<T(*)[Vec<T>::W]>(&v[0]));
            T a[GT_TD];
            constexpr int VW = Vec<T>::W;
#pragma unroll
            for (int t = 0; t < GT_TD; t += VW) vec_load(&sm.w[h][u][t], *reinterpret_cast<T(*)[VW]>(&a[t]));
#pragma unroll
            for (int t = 0; t < GT_TD; ++t)
#pragma unroll
                for (int q = 0; q < W; ++q) acc[t][q] += a[t] * v[q];
        }
#pragma unroll
        for (int t = 0; t < GT_TD; ++t) {
            const int64_t node = sm.node[t];
            if (node < 0) continue;
            if constexpr (W == 1) out[node * HC + c] = acc[t][0];
            else vec_store(out + node * HC + c, *reinterpret_cast<T(*)[Vec<T>::W]>(&acc[t][0]));
        }
    }
}

// Shared features, forward: rows are (n, C) and every head weighs the same row.  The CTA is cut into H groups of
// GT_THREADS / H threads (H in {1, 2, 4}); the groups work on the SAME columns at the same time, so a row piece is fetched
// from L2 once and served to the other heads by L1, and each thread keeps its eight accumulators of ONE head.
template <typename T, int W, typename S>
__device__ __forceinline__ void gat_tile_aggregate_shared_fwd(const S& sm, const T* __restrict__ rows, T* __restrict__ out, int U, int H,
                                                              int C) {
    constexpr int VW = Vec<T>::W;
    const int tpg = GT_THREADS / H;
    const int h = threadIdx.x / tpg, j = threadIdx.x - h * tpg;
    for (int c = (blockIdx.y * tpg + j) * W; c < C; c += gridDim.y * tpg * W) {
        T acc[GT_TD][W];
#pragma unroll
        for (int t = 0; t < GT_TD; ++t)
#pragma unroll
            for (int q = 0; q < W; ++q) acc[t][q] = T(0);
#pragma unroll 8
        for (int u = 0; u < U; ++u) {
            T v[W];
            if constexpr (W == 1) v[0] = rows[(int64_t)sm.u_id[u] * C + c];
            else vec_load(rows + (int64_t)sm.u_id[u] * C + c, v);
            T a[GT_TD];
#pragma unroll
            for (int t = 0; t < GT_TD; t += VW) vec_load(&sm.w[h][u][t], *reinterpret_cast<T(*)[VW]>(&a[t]));
#pragma unroll
            for (int t = 0; t < GT_TD; ++t)
#pragma unroll
                for (int q = 0; q < W; ++q) acc[t][q] += a[t] * v[q];
        }
#pragma unroll
        for (int t = 0; t < GT_TD; ++t) {
            const int64_t node = sm.node[t];
            if (node < 0) continue;
            if constexpr (W == 1) out[(node * H + h) * C + c] = acc[t][0];
            else vec_store(out + (node * H + h) * C + c, acc[t]);
        }
    }
}

// Shared features, by-source backward: grad_x[member, cols] = sum over heads and distinct rows of w[h][row][member] *
// grad_out[row, h, cols] - the heads are summed inside the kernel, the output is (n_src, C).
template <typename T, int W, typename S>
__device__ __forceinline__ void gat_tile_aggregate_shared_bsrc(const S& sm, const T* __restrict__ rows, T* __restrict__ out, int U, int H,
                                                               int C) {
    constexpr int VW = Vec<T>::W;
    const int HC = H * C;
    for (int c = (blockIdx.y * GT_THREADS + threadIdx.x) * W; c < C; c += gridDim.y * GT_THREADS * W) {
        T acc[GT_TD][W];
#pragma unroll
        for (int t = 0; t < GT_TD; ++t)
#pragma unroll
            for (int q = 0; q < W; ++q) acc[t][q] = T(0);
#pragma unroll 2
        for (int u = 0; u < U; ++u) {
            const T* row = rows + (int64_t)sm.u_id[u] * HC + c;
            for (int h = 0; h < H; ++h) {
                T v[W];
                if constexpr (W == 1) v[0] = row[h * C];
                else vec_load(row + h * C, v);
                T a[GT_TD];
#pragma unroll
                for (int t = 0; t < GT_TD; t += VW) vec_load(&sm.w[h][u][t], *reinterpret_cast<T(*)[VW]>(&a[t]));
#pragma unroll
                for (int t = 0; t < GT_TD; ++t)
#pragma unroll
                    for (int q = 0; q < W; ++q) acc[t][q] += a[t] * v[q];
            }
        }
#pragma unroll
        for (int t = 0; t < GT_TD; ++t) {
            const int64_t node = sm.node[t];
            if (node < 0) continue;
            if constexpr (W == 1) out[node * C + c] = acc[t][0];
            else vec_store(out + node * C + c, acc[t]);
        }
    }
}

extern __shared__ __align__(16) unsigned char gat_smem_raw[];

// MODE 0: forward.  members = destinations, (ptr, nbr) = by-destination CSR, rows = feat, out = out; alpha is WRITTEN.
// MODE 1: by-source backward.  members = sources, (ptr, nbr, eid) = by-source lists, rows = grad_out, out = grad_feat;
//         alpha and dlogit are READ, grad_a (= grad_a_src) is written.
template <typename T, int MODE, bool SH>
__global__ void __launch_bounds__(GT_THREADS, 2) gat_tile_agg_kernel(const T* __restrict__ rows, const T* __restrict__ a_src,
                                                                     const T* __restrict__ a_dst, const int64_t* __restrict__ ptr,
                                                                     const int32_t* __restrict__ nbr, const int32_t* __restrict__ eid,
                                                                     const int64_t* __restrict__ e_limit_ptr,
                                                                     const int32_t* __restrict__ order, int64_t n, int H, int C, T slope,
                                                                     T* __restrict__ alpha, const T* __restrict__ dlogit,
                                                                     T* __restrict__ out, T* __restrict__ grad_a) {
    using S = GatTile<T, GT_HMAX, GT_UMAX>;
    S& sm = *reinterpret_cast<S*>(gat_smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t e_limit = e_limit_ptr ? *e_limit_ptr : INT64_MAX;
    if (!gat_tile_edges(sm, ptr, nbr, eid, e_limit, order, n)) {
        for (int m = 0; m < GT_TD; ++m) {                               // a tile that does not fit: node by node (uniform branch)
            const int64_t node = sm.node[m];
            if (node < 0) continue;
            const int64_t p0 = sm.p0[m], p1 = ptr[node + 1];
            if constexpr (MODE == 0) gat_fwd_node<T, SH>(rows, a_src, a_dst, nbr, node, p0, p1, H, C, slope, out, alpha);
            else if constexpr (SH) gat_bsrc_node_shared<T>(nbr, eid, node, p0, p1, e_limit, H, C, alpha, dlogit, rows, out, grad_a);
            else gat_bsrc_node<T>(nbr, eid, node, p0, p1, e_limit, H, C, alpha, dlogit, rows, out, grad_a);
        }
        return;
    }
    const int Et = sm.off[GT_TD], U = sm.n_u;
    {   // clear the dense weights of the rows in use
        T* wf = &sm.w[0][0][0];
        const int per_h = U * GT_TD;
        for (int i = tid; i < H * per_h; i += GT_THREADS) {
            const int h = i / per_h;
            wf[h * (GT_UMAX * GT_TD) + (i - h * per_h)] = T(0);
        }
        // (U == 0: nothing to clear and the division above is never reached)
    }
    if constexpr (MODE == 0) {
        for (int i = tid; i < Et * H; i += GT_THREADS) {                // every logit once
            const int s = i / H, h = i - s * H;
            T r = a_src[(int64_t)sm.e_nbr[s] * H + h] + a_dst[sm.node[sm.e_own[s]] * H + h];
            sm.ev[s][h] = r > T(0) ? r : r * slope;
        }
        __syncthreads();
        {   // segment softmax: eight lanes per (member, head) - the 8 x H <= 32 segments of a tile run side by side
            const int seg = tid >> 3, gl = tid & 7;
            const bool on = seg < GT_TD * H;
            const int m = on ? seg / H : 0, h = on ? seg - m * H : 0;
            const int s0 = on ? sm.off[m] : 0, s1 = on ? sm.off[m + 1] : 0;
            T mx = -INFINITY;
            for (int s = s0 + gl; s < s1; s += 8) mx = max(mx, sm.ev[s][h]);
            mx = oct_max(mx);
            T sum = T(0);
            for (int s = s0 + gl; s < s1; s += 8) {
                const T e = t_exp(sm.ev[s][h] - mx);
                sm.ev[s][h] = e;
                sum += e;
            }
            sum = oct_sum(sum) + T(1e-16);
            for (int s = s0 + gl; s < s1; s += 8) {
                const T a = sm.ev[s][h] / sum;
                sm.ev[s][h] = a;
                alpha[(int64_t)sm.e_gid[s] * H + h] = a;
            }
        }
    } else {
        for (int i = tid; i < Et * H; i += GT_THREADS) {
            const int s = i / H, h = i - s * H;
            sm.ev[s][h] = sm.e_nbr[s] >= 0 ? alpha[(int64_t)sm.e_gid[s] * H + h] : T(0);
        }
        {   // grad_a_src[member, head] = sum of its edges' dlogit (eight lanes per segment)
            const int seg = tid >> 3, gl = tid & 7;
            const bool on = seg < GT_TD * H;
            const int m = on ? seg / H : 0, h = on ? seg - m * H : 0;
            const int64_t node = on ? sm.node[m] : -1;
            T acc = T(0);
            if (node >= 0)
                for (int s = sm.off[m] + gl; s < sm.off[m + 1]; s += 8)
                    if (sm.e_nbr[s] >= 0) acc += dlogit[(int64_t)sm.e_gid[s] * H + h];
            acc = oct_sum(acc);
            if (node >= 0 && gl == 0) grad_a[node * H + h] = acc;
        }
    }
    __syncthreads();
    for (int i = tid; i < Et * H; i += GT_THREADS) {                    // scatter into the dense form (+=: parallel edges add up)
        const int s = i / H, h = i - s * H;
        if (sm.e_nbr[s] >= 0) atomicAdd(&sm.w[h][sm.e_loc[s]][sm.e_own[s]], sm.ev[s][h]);
    }
    __syncthreads();
    if constexpr (SH && MODE == 0) {
        if (C % Vec<T>::W == 0) gat_tile_aggregate_shared_fwd<T, Vec<T>::W>(sm, rows, out, U, H, C);
        else gat_tile_aggregate_shared_fwd<T, 1>(sm, rows, out, U, H, C);
    } else if constexpr (SH) {
        if (C % Vec<T>::W == 0) gat_tile_aggregate_shared_bsrc<T, Vec<T>::W>(sm, rows, out, U, H, C);
        else gat_tile_aggregate_shared_bsrc<T, 1>(sm, rows, out, U, H, C);
    } else {
        if (C % Vec<T>::W == 0) gat_tile_aggregate<T, Vec<T>::W>(sm, rows, out, U, H, C);
        else gat_tile_aggregate<T, 1>(sm, rows, out, U, H, C);
    }
}

// By-destination backward in tile form.  Needs C = 2 * W * tph with tph (threads per head) in {32, 64, 128, 256}: every thread
// owns two W-wide column groups of ONE head, half a head apart, so a warp never straddles heads and the butterfly is uniform.
// Wider heads (C = S * 2 * W * 256, the aggregate-first layers of spadot_b200/gat.py: C = in_channels) are cut into S column
// slices of 2 * W * 256 that take turns - the whole CTA on one slice, partial dot products added up in shared memory.
template <typename T, bool SH>
__global__ void __launch_bounds__(GT_THREADS, 2) gat_tile_bdst_kernel(const T* __restrict__ feat, const T* __restrict__ a_src,
                                                                      const T* __restrict__ a_dst, const int64_t* __restrict__ rowptr,
                                                                      const int32_t* __restrict__ col, const int32_t* __restrict__ order,
                                                                      int64_t n, int H, int C, T slope, const T* __restrict__ alpha,
                                                                      const T* __restrict__ grad_out, T* __restrict__ dlogit,
                                                                      T* __restrict__ grad_a_dst) {
    using S = GatTile<T, GT_THREADS / 32, GT_UCHUNK>;
    S& sm = *reinterpret_cast<S*>(gat_smem_raw);
    constexpr int W = Vec<T>::W;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (!gat_tile_edges(sm, rowptr, col, nullptr, INT64_MAX, order, n)) {
        for (int m = 0; m < GT_TD; ++m) {
            const int64_t node = sm.node[m];
            if (node < 0) continue;
            gat_bdst_node<T, SH>(feat, a_src, a_dst, col, node, sm.p0[m], rowptr[node + 1], H, C, slope, alpha, grad_out, dlogit, grad_a_dst);
        }
        return;
    }
    const int Et = sm.off[GT_TD], U = sm.n_u;
    const int HC = H * C;
    const int NS = max(1, C / (2 * W * GT_THREADS));   // column slices per head (1 unless a head is wider than the CTA covers)
    const int Cs = C / NS;                // slice width
    const int tph = Cs / (2 * W);         // threads per slice
    const int hps = GT_THREADS / tph;     // slices per sweep (1 when NS > 1)
    const int wph = tph >> 5;             // warps per slice
    const int hl = tid / tph;             // slice slot of this thread within a sweep
    const int j = tid - hl * tph;
    const int n_sl = H * NS;
    __syncthreads();                      // e_loc
    // gridDim.y > 1 (few tiles, hps == 1): CTA y takes the heads h with h % gridDim.y == y - their slices here, their
    // softmax backward below
    for (int h0 = 0; h0 < n_sl; h0 += hps) {
        if (gridDim.y > 1 && (h0 / NS) % (int)gridDim.y != (int)blockIdx.y) continue;
        const int sl = h0 + hl;
        const bool live = sl < n_sl;      // uniform per warp (tph is a multiple of 32)
        const int c0 = live ? (sl / NS) * C + (sl % NS) * Cs + j * W : j * W, c1 = c0 + (Cs >> 1);
        // column of the gathered row: the head's slice of an (n, H, C) row, or the same (n, C) row for every head
        const int r0 = SH ? (live ? (sl % NS) * Cs + j * W : j * W) : c0, r1c = r0 + (Cs >> 1);
        const int RS = SH ? C : HC;
        T go[GT_TD][2 * W];
#pragma unroll
        for (int t = 0; t < GT_TD; ++t) {
            const int64_t node = sm.node[t];
            if (live && node >= 0) {
                vec_load(grad_out + node * HC + c0, *reinterpret_cast<T(*)[W]>(&go[t][0]));
                vec_load(grad_out + node * HC + c1, *reinterpret_cast<T(*)[W]>(&go[t][W]));
            } else {
#pragma unroll
                for (int q = 0; q < 2 * W; ++q) go[t][q] = T(0);
            }
        }
        for (int ub = 0; ub < U; ub += GT_UCHUNK) {
        const int ue = min(U, ub + GT_UCHUNK);
#pragma unroll 2
        for (int u = ub; u < ue; ++u) {
            T v[2 * W];
            const T* row = feat + (int64_t)sm.u_id[u] * RS;
            vec_load(row + r0, *reinterpret_cast<T(*)[W]>(&v[0]));
            vec_load(row + r1c, *reinterpret_cast<T(*)[W]>(&v[W]));
            T p[GT_TD];
#pragma unroll
            for (int t = 0; t < GT_TD; ++t) {
                T acc = go[t][0] * v[0];
#pragma unroll
                for (int q = 1; q < 2 * W; ++q) acc += go[t][q] * v[q];
                p[t] = acc;
            }
            // eight sums across 32 lanes in 9 shuffles: halve the number of values while doubling the lanes that own each
            T q4[4], q2[2];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const bool up = lane & 16;
                const T send = up ? p[k] : p[k + 4], keep = up ? p[k + 4] : p[k];
                q4[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
            }
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const bool up = lane & 8;
                const T send = up ? q4[k] : q4[k + 2], keep = up ? q4[k + 2] : q4[k];
                q2[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
            }
            T r1;
            {
                const bool up = lane & 4;
                const T send = up ? q2[0] : q2[1], keep = up ? q2[1] : q2[0];
                r1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
            }
            r1 += __shfl_xor_sync(0xffffffffu, r1, 2);
            r1 += __shfl_xor_sync(0xffffffffu, r1, 1);
            // lane bits 4,3,2 chose members +4, +2, +1
            if ((lane & 3) == 0) sm.w[warp][u - ub][((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1)] = r1;
        }
        __syncthreads();
        for (int i = tid; i < Et * hps; i += GT_THREADS) {              // d alpha of the edges that read these rows, heads of this sweep
            const int s = i / hps, k = i - s * hps;
            const int loc = sm.e_loc[s];
            if (h0 + k >= n_sl || loc < ub || loc >= ue) continue;
            T acc = T(0);
            for (int wi = 0; wi < wph; ++wi) acc += sm.w[k * wph + wi][loc - ub][sm.e_own[s]];
            const int hh = (h0 + k) / NS;
            if ((h0 + k) % NS == 0) sm.ev[s][hh] = acc;       // first slice of the head (slices of one head never share a sweep)
            else sm.ev[s][hh] += acc;
        }
        __syncthreads();
        }
    }
    // softmax backward per (member, head): dlogit = alpha (d alpha - sum alpha d alpha), through the LeakyReLU
    {   // eight lanes per (member, head): all segments of the tile side by side
        const int seg = tid >> 3, gl = tid & 7;
        const bool on = seg < GT_TD * H && (gridDim.y == 1 || (seg % H) % (int)gridDim.y == (int)blockIdx.y);
        const int m = on ? seg / H : 0, h = on ? seg - m * H : 0;
        const int64_t node = on ? sm.node[m] : -1;
        const int s0 = node >= 0 ? sm.off[m] : 0, s1 = node >= 0 ? sm.off[m + 1] : 0;
        T sa = T(0);
        for (int s = s0 + gl; s < s1; s += 8) sa += alpha[(int64_t)sm.e_gid[s] * H + h] * sm.ev[s][h];
        sa = oct_sum(sa);
        const T ad = node >= 0 ? a_dst[node * H + h] : T(0);
        T gd = T(0);
        for (int s = s0 + gl; s < s1; s += 8) {
            const T r = a_src[(int64_t)sm.e_nbr[s] * H + h] + ad;
            T g = alpha[(int64_t)sm.e_gid[s] * H + h] * (sm.ev[s][h] - sa);
            g = r > T(0) ? g : g * slope;
            dlogit[(int64_t)sm.e_gid[s] * H + h] = g;
            gd += g;
        }
        gd = oct_sum(gd);
        if (node >= 0 && gl == 0) grad_a_dst[node * H + h] = gd;
    }
}

// Which form runs.  The tile form pays off when consecutive CTA members share neighbours, i.e. when the caller supplies a
// locality order (spadot_b200/gat.py does for graphs of >= 20000 nodes: Z-curve of the spot coordinates, or reverse
// Cuthill-McKee); without one - small graphs, where these kernels are launch-bound anyway - the per-node form runs.
// SDB_GAT_TILES=0 / 1 forces the per-node / the tile form (A/B measurements, tests).
static bool gat_use_tiles(int H, const int32_t* order) {
    if (H > GT_HMAX) return false;
    const char* e = getenv("SDB_GAT_TILES");
    if (e && e[0] == '0') return false;
    if (e && e[0] == '1') return true;
    return order != nullptr;
}
template <typename T>
static bool gat_bdst_tile_shape(int C) {
    constexpr int W = Vec<T>::W;
    if (C % (2 * W)) return false;
    const int tph = C / (2 * W);
    if (tph > GT_THREADS) return tph % GT_THREADS == 0;      // S slices of a full CTA each
    return tph == 32 || tph == 64 || tph == 128 || tph == 256;
}

// Few tiles (a prefix layer with 512 destinations is 64 tiles on 148 SMs): the column sweeps of a tile are dealt over
// gridDim.y CTAs, each of which repeats the tile's (cheap) phase A; `sweeps` = column sweeps one CTA would otherwise do.
static int gat_column_splits(int64_t n, int sweeps) {
    const int64_t tiles = (n + GT_TD - 1) / GT_TD;
    const int64_t want = (4 * 148 + tiles - 1) / tiles;
    return (int)max((int64_t)1, min((int64_t)sweeps, min(want, (int64_t)16)));
}

template <typename T, int MODE, bool SH, typename... Args>
static int gat_launch_agg(int64_t n, int ny, cudaStream_t st, Args... args) {
    static size_t memo[SDB_MAX_DEVICES];
    using S = GatTile<T, GT_HMAX, GT_UMAX>;
    auto kern = gat_tile_agg_kernel<T, MODE, SH>;
    cudaError_t e = sdb_ensure_smem(kern, sizeof(S), memo);
    if (e != cudaSuccess) return (int)e;
    kern<<<dim3((unsigned)((n + GT_TD - 1) / GT_TD), (unsigned)ny), GT_THREADS, sizeof(S), st>>>(args...);
    SDB_LAUNCH_STATUS();
}
template <typename T, bool SH, typename... Args>
static int gat_launch_bdst(int64_t n, int ny, cudaStream_t st, Args... args) {
    static size_t memo[SDB_MAX_DEVICES];
    using S = GatTile<T, GT_THREADS / 32, GT_UCHUNK>;
    auto kern = gat_tile_bdst_kernel<T, SH>;
    cudaError_t e = sdb_ensure_smem(kern, sizeof(S), memo);
    if (e != cudaSuccess) return (int)e;
    kern<<<dim3((unsigned)((n + GT_TD - 1) / GT_TD), (unsigned)ny), GT_THREADS, sizeof(S), st>>>(args...);
    SDB_LAUNCH_STATUS();
}

// SH: feat is (n_src, C), shared by the heads; out (n, H, C).  The shared forward's tile form splits the CTA into H groups.
template <typename T, bool SH>
int gat_forward_t(const void* feat, const void* a_src, const void* a_dst, const int64_t* rowptr, const int32_t* col,
                  const int32_t* order, int64_t n, int H, int C, double slope, void* out, void* alpha, cudaStream_t st) {
    if (gat_use_tiles(H, order) && (!SH || GT_THREADS % H == 0))
        return gat_launch_agg<T, 0, SH>(n, gat_column_splits(n, SH ? (C * H + GT_THREADS * (int)Vec<T>::W - 1) / (GT_THREADS * (int)Vec<T>::W)
                                                                      : (H * C + GT_THREADS * (int)Vec<T>::W - 1) / (GT_THREADS * (int)Vec<T>::W)),
                                        st, (const T*)feat, (const T*)a_src, (const T*)a_dst, rowptr, col, (const int32_t*)nullptr,
                                        (const int64_t*)nullptr, order, n, H, C, (T)slope, (T*)alpha, (const T*)nullptr, (T*)out,
                                        (T*)nullptr);
    gat_fwd_kernel<T, SH><<<(unsigned)n, 128, 0, st>>>((const T*)feat, (const T*)a_src, (const T*)a_dst, rowptr, col, order, n, H, C,
                                                      (T)slope, (T*)out, (T*)alpha);
    SDB_LAUNCH_STATUS();
}

// SH: grad_feat is (n_src, C) - the sum over the heads happens inside the by-source kernel.
template <typename T, bool SH>
int gat_backward_t(const void* feat, const void* a_src, const void* a_dst, const int64_t* rowptr, const int32_t* col,
                   const int64_t* src_rowptr, const int32_t* src_dst, const int32_t* src_eid, const int32_t* order_dst,
                   const int32_t* order_src, int64_t n_dst, int64_t n_src, int H, int C, double slope, const void* alpha,
                   const void* grad_out, void* dlogit, void* grad_feat, void* grad_a_src, void* grad_a_dst, cudaStream_t st) {
    cudaError_t e;
    if (gat_use_tiles(H, order_dst) && gat_bdst_tile_shape<T>(C)) {
        // heads over gridDim.y only when one slice fills the CTA (threads per head >= GT_THREADS) and H splits evenly
        const int tph = C / (2 * (int)Vec<T>::W);
        int ny = 1;
        if (tph >= GT_THREADS) {
            ny = gat_column_splits(n_dst, H);
            while (H % ny) --ny;
        }
        const int rc = gat_launch_bdst<T, SH>(n_dst, ny, st, (const T*)feat, (const T*)a_src, (const T*)a_dst, rowptr, col, order_dst, n_dst,
                                              H, C, (T)slope, (const T*)alpha, (const T*)grad_out, (T*)dlogit, (T*)grad_a_dst);
        if (rc) return rc;
    } else {
        gat_bwd_dst_kernel<T, SH><<<(unsigned)n_dst, 128, 0, st>>>((const T*)feat, (const T*)a_src, (const T*)a_dst, rowptr, col,
                                                                  order_dst, n_dst, H, C, (T)slope, (const T*)alpha, (const T*)grad_out,
                                                                  (T*)dlogit, (T*)grad_a_dst);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    if (gat_use_tiles(H, order_src))
        return gat_launch_agg<T, 1, SH>(n_src, gat_column_splits(n_src, ((SH ? C : H * C) + GT_THREADS * (int)Vec<T>::W - 1) / (GT_THREADS * (int)Vec<T>::W)),
                                        st, (const T*)grad_out, (const T*)nullptr, (const T*)nullptr, src_rowptr, src_dst, src_eid,
                                        rowptr + n_dst, order_src, n_src, H, C, (T)slope, const_cast<T*>((const T*)alpha),
                                        (const T*)dlogit, (T*)grad_feat, (T*)grad_a_src);
    gat_bwd_src_kernel<T, SH><<<(unsigned)n_src, 128, 0, st>>>(src_rowptr, src_dst, src_eid, order_src, rowptr + n_dst, H, C,
                                                              (const T*)alpha, (const T*)dlogit, (const T*)grad_out, (T*)grad_feat,
                                                              (T*)grad_a_src);
    SDB_LAUNCH_STATUS();
}

}  // namespace

extern "C" {

int sdb_gat_forward(const void* feat, const void* a_src, const void* a_dst, const int64_t* rowptr, const int32_t* col,
                    const int32_t* node_order, int64_t n, int H, int C, double negative_slope, int is_double, void* out,
                    void* alpha, void* stream) {
    SDB_CHECK_ARG(feat && a_src && a_dst && rowptr && col && out && alpha && n >= 0 && H > 0 && C > 0);
    if (n == 0) return 0;
    if (n > 2147483647LL) return SDB_E_UNSUPPORTED;
    return is_double ? gat_forward_t<double, false>(feat, a_src, a_dst, rowptr, col, node_order, n, H, C, negative_slope, out, alpha, sdb_stream(stream))
                     : gat_forward_t<float, false>(feat, a_src, a_dst, rowptr, col, node_order, n, H, C, negative_slope, out, alpha, sdb_stream(stream));
}

int sdb_gat_backward_prefix(const void* feat, const void* a_src, const void* a_dst, const int64_t* rowptr, const int32_t* col,
                            const int64_t* src_rowptr, const int32_t* src_dst, const int32_t* src_eid, const int32_t* order_dst,
                            const int32_t* order_src, int64_t n_dst, int64_t n_src, int H, int C, double negative_slope,
                            int is_double, const void* alpha, const void* grad_out, void* dlogit, void* grad_feat,
                            void* grad_a_src, void* grad_a_dst, void* stream) {
    SDB_CHECK_ARG(feat && a_src && a_dst && rowptr && col && src_rowptr && src_dst && src_eid && alpha && grad_out && dlogit &&
                  grad_feat && grad_a_src && grad_a_dst && n_dst >= 0 && n_src >= n_dst && H > 0 && C > 0);
    if (n_src == 0) return 0;
    if (n_src > 2147483647LL) return SDB_E_UNSUPPORTED;
    if (n_dst == 0) return SDB_E_INVALID;                 // a layer without destinations has no backward
    return is_double ? gat_backward_t<double, false>(feat, a_src, a_dst, rowptr, col, src_rowptr, src_dst, src_eid, order_dst, order_src,
                                              n_dst, n_src, H, C, negative_slope, alpha, grad_out, dlogit, grad_feat, grad_a_src,
                                              grad_a_dst, sdb_stream(stream))
                     : gat_backward_t<float, false>(feat, a_src, a_dst, rowptr, col, src_rowptr, src_dst, src_eid, order_dst, order_src,
                                             n_dst, n_src, H, C, negative_slope, alpha, grad_out, dlogit, grad_feat, grad_a_src,
                                             grad_a_dst, sdb_stream(stream));
}

int sdb_gat_forward_shared(const void* feat, const void* a_src, const void* a_dst, const int64_t* rowptr, const int32_t* col,
                           const int32_t* node_order, int64_t n, int H, int C, double negative_slope, int is_double, void* out,
                           void* alpha, void* stream) {
    SDB_CHECK_ARG(feat && a_src && a_dst && rowptr && col && out && alpha && n >= 0 && H > 0 && C > 0);
    if (n == 0) return 0;
    if (n > 2147483647LL) return SDB_E_UNSUPPORTED;
    return is_double ? gat_forward_t<double, true>(feat, a_src, a_dst, rowptr, col, node_order, n, H, C, negative_slope, out, alpha, sdb_stream(stream))
                     : gat_forward_t<float, true>(feat, a_src, a_dst, rowptr, col, node_order, n, H, C, negative_slope, out, alpha, sdb_stream(stream));
}

int sdb_gat_backward_shared(const void* feat, const void* a_src, const void* a_dst, const int64_t* rowptr, const int32_t* col,
                            const int64_t* src_rowptr, const int32_t* src_dst, const int32_t* src_eid, const int32_t* order_dst,
                            const int32_t* order_src, int64_t n_dst, int64_t n_src, int H, int C, double negative_slope,
                            int is_double, const void* alpha, const void* grad_out, void* dlogit, void* grad_feat,
                            void* grad_a_src, void* grad_a_dst, void* stream) {
    SDB_CHECK_ARG(feat && a_src && a_dst && rowptr && col && src_rowptr && src_dst && src_eid && alpha && grad_out && dlogit &&
                  grad_feat && grad_a_src && grad_a_dst && n_dst >= 0 && n_src >= n_dst && H > 0 && C > 0);
    if (n_src == 0) return 0;
    if (n_src > 2147483647LL) return SDB_E_UNSUPPORTED;
    if (n_dst == 0) return SDB_E_INVALID;
    return is_double ? gat_backward_t<double, true>(feat, a_src, a_dst, rowptr, col, src_rowptr, src_dst, src_eid, order_dst, order_src,
                                                    n_dst, n_src, H, C, negative_slope, alpha, grad_out, dlogit, grad_feat, grad_a_src,
                                                    grad_a_dst, sdb_stream(stream))
                     : gat_backward_t<float, true>(feat, a_src, a_dst, rowptr, col, src_rowptr, src_dst, src_eid, order_dst, order_src,
                                                   n_dst, n_src, H, C, negative_slope, alpha, grad_out, dlogit, grad_feat, grad_a_src,
                                                   grad_a_dst, sdb_stream(stream));
}

int sdb_gat_backward(const void* feat, const void* a_src, const void* a_dst, const int64_t* rowptr, const int32_t* col,
                     const int64_t* src_rowptr, const int32_t* src_dst, const int32_t* src_eid, const int32_t* node_order, int64_t n,
                     int H, int C, double negative_slope, int is_double, const void* alpha, const void* grad_out, void* dlogit,
                     void* grad_feat, void* grad_a_src, void* grad_a_dst, void* stream) {
    SDB_CHECK_ARG(feat && a_src && a_dst && rowptr && col && src_rowptr && src_dst && src_eid && alpha && grad_out && dlogit &&
                  grad_feat && grad_a_src && grad_a_dst && n >= 0 && H > 0 && C > 0);
    if (n == 0) return 0;
    return sdb_gat_backward_prefix(feat, a_src, a_dst, rowptr, col, src_rowptr, src_dst, src_eid, node_order, node_order, n, n, H, C,
                                   negative_slope, is_double, alpha, grad_out, dlogit, grad_feat, grad_a_src, grad_a_dst, stream);
}

}  // extern "C"

// ---------------------------------------------------------------------------------------- spatial kNN graph
// Brute-force k nearest neighbours in 2-D/3-D (k <= 32) for the GAT's spatial graph
// (ref: SpaDOT/utils/_utils.py:52-100 _Cal_Spatial_Net: sklearn NearestNeighbors(k+1) minus self).
// One thread per query keeps its k best (distance, index) pairs sorted by (distance, index); candidate
// coordinates are streamed through shared memory.  fp64 distances so pixel-scale coordinates do not collide.
namespace {

constexpr int KNN_MAX_K = 32;
constexpr int KNN_TILE = 1024;

__global__ void __launch_bounds__(128) knn_kernel(const double* __restrict__ pts, int64_t n, int dim, int k, int32_t* __restrict__ out_idx,
                                                  double* __restrict__ out_dist) {
    __shared__ double tile[KNN_TILE * 3];
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double qx[3] = {0.0, 0.0, 0.0};
    if (q < n) for (int d = 0; d < dim; ++d) qx[d] = pts[q * dim + d];
    double bd[KNN_MAX_K];
    int32_t bi[KNN_MAX_K];
    for (int t = 0; t < KNN_MAX_K; ++t) { bd[t] = INFINITY; bi[t] = 0x7fffffff; }
    for (int64_t base = 0; base < n; base += KNN_TILE) {
        const int cnt = (int)min((int64_t)KNN_TILE, n - base);
        __syncthreads();
        for (int e = threadIdx.x; e < cnt * dim; e += blockDim.x) tile[e] = pts[base * dim + e];
        __syncthreads();
        if (q >= n) continue;
        for (int c = 0; c < cnt; ++c) {
            const int64_t j = base + c;
            if (j == q) continue;
            double d2 = 0.0;
            for (int d = 0; d < dim; ++d) { const double df = qx[d] - tile[c * dim + d]; d2 += df * df; }
            if (d2 < bd[k - 1] || (d2 == bd[k - 1] && (int32_t)j < bi[k - 1])) {
                int pos = k - 1;
                while (pos > 0 && (bd[pos - 1] > d2 || (bd[pos - 1] == d2 && bi[pos - 1] > (int32_t)j))) {
                    bd[pos] = bd[pos - 1];
                    bi[pos] = bi[pos - 1];
                    --pos;
                }
                bd[pos] = d2;
                bi[pos] = (int32_t)j;
            }
        }
    }
    if (q < n)
        for (int t = 0; t < k; ++t) {
            out_idx[q * k + t] = bi[t];
            if (out_dist) out_dist[q * k + t] = sqrt(bd[t]);
        }
}

// Grid-binned form for 2-D coordinates: O(n k) instead of O(n^2).  Points are sorted by the cell of a uniform grid (about
// four points per cell); a query walks the rings of cells around its own cell and stops as soon as its k-th best distance is
// strictly smaller than the distance to the border of the block of cells it has already visited - every unvisited point lies
// outside that block - so the result is exactly the brute-force one, including the (distance, index) order.
__global__ void __launch_bounds__(128) knn_grid_kernel(const double* __restrict__ pts, const int32_t* __restrict__ order,
                                                       const int32_t* __restrict__ cell_of, const int32_t* __restrict__ cell_start,
                                                       int64_t n, int gx, int gy, double x0, double y0, double h, int k,
                                                       int32_t* __restrict__ out_idx, double* __restrict__ out_dist) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // sorted slot: neighbouring threads, neighbouring cells
    if (s >= n) return;
    const double qx = pts[s * 2], qy = pts[s * 2 + 1];
    const int32_t qi = order[s];
    const int cell = cell_of[s], cx = cell % gx, cy = cell / gx;
    double bd[KNN_MAX_K];
    int32_t bi[KNN_MAX_K];
    for (int t = 0; t < KNN_MAX_K; ++t) { bd[t] = INFINITY; bi[t] = 0x7fffffff; }
    const int rmax = max(max(cx, gx - 1 - cx), max(cy, gy - 1 - cy));
    for (int r = 0; r <= rmax; ++r) {
        const int y_lo = max(cy - r, 0), y_hi = min(cy + r, gy - 1);
        for (int yy = y_lo; yy <= y_hi; ++yy) {
            const bool edge_row = (yy == cy - r) || (yy == cy + r);
            // an edge row of the ring is scanned in full, an interior row only at its two end cells
            const int step = edge_row ? 1 : max(2 * r, 1);
            for (int xx = cx - r; xx <= cx + r; xx += step) {
                if (xx < 0 || xx >= gx) continue;
                const int c = yy * gx + xx;
                for (int p = cell_start[c]; p < cell_start[c + 1]; ++p) {
                    const int32_t j = order[p];
                    if (j == qi) continue;
                    double d2 = 0.0;
                    { const double df = qx - pts[(int64_t)p * 2]; d2 += df * df; }
                    { const double df = qy - pts[(int64_t)p * 2 + 1]; d2 += df * df; }
                    if (d2 < bd[k - 1] || (d2 == bd[k - 1] && j < bi[k - 1])) {
                        int pos = k - 1;
                        while (pos > 0 && (bd[pos - 1] > d2 || (bd[pos - 1] == d2 && bi[pos - 1] > j))) {
                            bd[pos] = bd[pos - 1];
                            bi[pos] = bi[pos - 1];
                            --pos;
                        }
                        bd[pos] = d2;
                        bi[pos] = j;
                    }
                }
            }
        }
        // distance from the query to the border of the visited block [cx-r, cx+r] x [cy-r, cy+r] (sides at the grid's edge do not
        // count: nothing lies beyond them); a relative margin covers the rounding of the cell assignment
        double dmin = INFINITY;
        if (cx - r > 0) dmin = fmin(dmin, qx - (x0 + (double)(cx - r) * h));
        if (cx + r < gx - 1) dmin = fmin(dmin, (x0 + (double)(cx + r + 1) * h) - qx);
        if (cy - r > 0) dmin = fmin(dmin, qy - (y0 + (double)(cy - r) * h));
        if (cy + r < gy - 1) dmin = fmin(dmin, (y0 + (double)(cy + r + 1) * h) - qy);
        dmin -= 1e-9 * h;
        if (dmin > 0.0 && bd[k - 1] < dmin * dmin) break;
    }
    for (int t = 0; t < k; ++t) {
        out_idx[(int64_t)qi * k + t] = bi[t];
        if (out_dist) out_dist[(int64_t)qi * k + t] = sqrt(bd[t]);
    }
}

}  // namespace

extern "C" int sdb_knn_grid_f64(const double* pts_sorted, const int32_t* order, const int32_t* cell_of_sorted, const int32_t* cell_start,
                                int64_t n, int gx, int gy, double x0, double y0, double h, int k, int32_t* out_idx, double* out_dist,
                                void* stream) {
    SDB_CHECK_ARG(pts_sorted && order && cell_of_sorted && cell_start && out_idx && n >= 0 && gx > 0 && gy > 0 && h > 0.0 && k >= 1 &&
                  k <= KNN_MAX_K);
    if (n == 0) return 0;
    if (k > n - 1 || n > 2147483647LL) return SDB_E_INVALID;
    knn_grid_kernel<<<(unsigned)((n + 127) / 128), 128, 0, sdb_stream(stream)>>>(pts_sorted, order, cell_of_sorted, cell_start, n, gx, gy,
                                                                                   x0, y0, h, k, out_idx, out_dist);
    SDB_LAUNCH_STATUS();
}

extern "C" int sdb_knn_f64(const double* pts, int64_t n, int dim, int k, int32_t* out_idx, double* out_dist, void* stream) {
    SDB_CHECK_ARG(pts && out_idx && n >= 0 && dim >= 1 && dim <= 3 && k >= 1 && k <= KNN_MAX_K);
    if (n == 0) return 0;
    if (k > n - 1) return SDB_E_INVALID;
    knn_kernel<<<(unsigned)((n + 127) / 128), 128, 0, sdb_stream(stream)>>>(pts, n, dim, k, out_idx, out_dist);
    SDB_LAUNCH_STATUS();
}
