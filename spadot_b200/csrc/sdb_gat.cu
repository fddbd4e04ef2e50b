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

// ---------------------------------------------------------------------------------------- forward
template <typename T>
__global__ void __launch_bounds__(128) gat_fwd_kernel(const T* __restrict__ feat, const T* __restrict__ a_src,
                                                      const T* __restrict__ a_dst, const int64_t* __restrict__ rowptr,
                                                      const int32_t* __restrict__ col, const int32_t* __restrict__ order,
                                                      int64_t n, int H, int C, T slope, T* __restrict__ out,
                                                      T* __restrict__ alpha) {
    const int64_t i = order ? order[blockIdx.x] : blockIdx.x;   // locality order of the CTAs (L2 reuse of gathered rows)
    const int64_t e0 = rowptr[i], e1 = rowptr[i + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // segment softmax: one warp per head (round-robin when H > 4)
    for (int h = warp; h < H; h += 4) {
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
    // aggregation: thread owns feature columns c, c+128, ...; each source row is read once, coalesced
    const int HC = H * C;
    constexpr int W = Vec<T>::W;
    if (C % W == 0) {
        for (int c = threadIdx.x * W; c < HC; c += 128 * W) {
            const int h = c / C;
            T acc[W];
#pragma unroll
            for (int q = 0; q < W; ++q) acc[q] = T(0);
            for (int64_t e = e0; e < e1; ++e) {
                const T a = alpha[e * H + h];
                T v[W];
                vec_load(feat + (int64_t)col[e] * HC + c, v);
#pragma unroll
                for (int q = 0; q < W; ++q) acc[q] += a * v[q];
            }
            vec_store(out + i * HC + c, acc);
        }
    } else {
        for (int c = threadIdx.x; c < HC; c += 128) {
            const int h = c / C;
            T acc = T(0);
            for (int64_t e = e0; e < e1; ++e) acc += alpha[e * H + h] * feat[(int64_t)col[e] * HC + c];
            out[i * HC + c] = acc;
        }
    }
}

// ---------------------------------------------------------------------------------------- backward, by destination
// dlogit[e,h] = d loss / d (a_src[j,h] + a_dst[i,h]);  grad_a_dst[i,h] = sum_e dlogit[e,h]
template <typename T>
__global__ void __launch_bounds__(128) gat_bwd_dst_kernel(const T* __restrict__ feat, const T* __restrict__ a_src,
                                                          const T* __restrict__ a_dst, const int64_t* __restrict__ rowptr,
                                                          const int32_t* __restrict__ col, const int32_t* __restrict__ order,
                                                          int64_t n, int H, int C, T slope, const T* __restrict__ alpha,
                                                          const T* __restrict__ grad_out, T* __restrict__ dlogit,
                                                          T* __restrict__ grad_a_dst) {
    const int64_t i = order ? order[blockIdx.x] : blockIdx.x;
    const int64_t e0 = rowptr[i], e1 = rowptr[i + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int HC = H * C;
    const int64_t n_eh = (e1 - e0) * H;
    // d alpha[e,h] = <grad_out[i,h,:], feat[j,h,:]>  (one warp per (edge, head), coalesced over C)
    for (int64_t p = warp; p < n_eh; p += 4) {
        const int64_t e = e0 + p / H;
        const int h = (int)(p % H);
        const T* go = grad_out + i * HC + h * C;
        const T* fj = feat + (int64_t)col[e] * HC + h * C;
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
    for (int h = warp; h < H; h += 4) {
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
}

// ---------------------------------------------------------------------------------------- backward, by source
// grad_feat[j,h,:] = sum_{e: j->i} alpha[e,h] * grad_out[i,h,:];  grad_a_src[j,h] = sum_e dlogit[e,h]
template <typename T>
__global__ void __launch_bounds__(128) gat_bwd_src_kernel(const int64_t* __restrict__ src_rowptr, const int32_t* __restrict__ src_dst,
                                                          const int32_t* __restrict__ src_eid, const int32_t* __restrict__ order,
                                                          const int64_t* __restrict__ e_limit_ptr, int H, int C,
                                                          const T* __restrict__ alpha,
                                                          const T* __restrict__ dlogit, const T* __restrict__ grad_out,
                                                          T* __restrict__ grad_feat, T* __restrict__ grad_a_src) {
    const int64_t j = order ? order[blockIdx.x] : blockIdx.x;
    const int64_t p0 = src_rowptr[j], p1 = src_rowptr[j + 1];
    // prefix form: only the edges into the first n_dst destinations exist for this layer; in the by-destination order they
    // are the edge ids below rowptr[n_dst] (*e_limit_ptr), so the full by-source lists are filtered instead of rebuilt
    const int64_t e_limit = *e_limit_ptr;
    const int HC = H * C;
    constexpr int W = Vec<T>::W;
    if (C % W == 0) {
        for (int c = threadIdx.x * W; c < HC; c += 128 * W) {
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
        for (int c = threadIdx.x; c < HC; c += 128) {
            const int h = c / C;
            T acc = T(0);
            for (int64_t p = p0; p < p1; ++p)
                if (src_eid[p] < e_limit) acc += alpha[(int64_t)src_eid[p] * H + h] * grad_out[(int64_t)src_dst[p] * HC + c];
            grad_feat[j * HC + c] = acc;
        }
    }
    for (int h = threadIdx.x; h < H; h += 128) {
        T acc = T(0);
        for (int64_t p = p0; p < p1; ++p)
            if (src_eid[p] < e_limit) acc += dlogit[(int64_t)src_eid[p] * H + h];
        grad_a_src[j * H + h] = acc;
    }
}

template <typename T>
int gat_forward_t(const void* feat, const void* a_src, const void* a_dst, const int64_t* rowptr, const int32_t* col,
                  const int32_t* order, int64_t n, int H, int C, double slope, void* out, void* alpha, cudaStream_t st) {
    gat_fwd_kernel<T><<<(unsigned)n, 128, 0, st>>>((const T*)feat, (const T*)a_src, (const T*)a_dst, rowptr, col, order, n, H, C,
                                                  (T)slope, (T*)out, (T*)alpha);
    SDB_LAUNCH_STATUS();
}

template <typename T>
int gat_backward_t(const void* feat, const void* a_src, const void* a_dst, const int64_t* rowptr, const int32_t* col,
                   const int64_t* src_rowptr, const int32_t* src_dst, const int32_t* src_eid, const int32_t* order_dst,
                   const int32_t* order_src, int64_t n_dst, int64_t n_src, int H, int C, double slope, const void* alpha,
                   const void* grad_out, void* dlogit, void* grad_feat, void* grad_a_src, void* grad_a_dst, cudaStream_t st) {
    gat_bwd_dst_kernel<T><<<(unsigned)n_dst, 128, 0, st>>>((const T*)feat, (const T*)a_src, (const T*)a_dst, rowptr, col, order_dst,
                                                          n_dst, H, C, (T)slope, (const T*)alpha, (const T*)grad_out, (T*)dlogit,
                                                          (T*)grad_a_dst);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    gat_bwd_src_kernel<T><<<(unsigned)n_src, 128, 0, st>>>(src_rowptr, src_dst, src_eid, order_src, rowptr + n_dst, H, C,
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
    return is_double ? gat_forward_t<double>(feat, a_src, a_dst, rowptr, col, node_order, n, H, C, negative_slope, out, alpha, sdb_stream(stream))
                     : gat_forward_t<float>(feat, a_src, a_dst, rowptr, col, node_order, n, H, C, negative_slope, out, alpha, sdb_stream(stream));
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
    return is_double ? gat_backward_t<double>(feat, a_src, a_dst, rowptr, col, src_rowptr, src_dst, src_eid, order_dst, order_src,
                                              n_dst, n_src, H, C, negative_slope, alpha, grad_out, dlogit, grad_feat, grad_a_src,
                                              grad_a_dst, sdb_stream(stream))
                     : gat_backward_t<float>(feat, a_src, a_dst, rowptr, col, src_rowptr, src_dst, src_eid, order_dst, order_src,
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
