// K7: Lloyd iterations of k-means on the device (per-epoch `_update_Kmeans`, ref: SpaDOT/utils/_train_utils.py:255-269,
// and `KMeans_Clustering` / `Adaptive_clustering`, ref: SpaDOT/utils/_analyze_utils.py:10-105 — both call
// sklearn.cluster.KMeans(n_init=10) on the host).  One launch assigns every sample to its nearest centre
// (first minimum on ties, like sklearn's lloyd_iter), flags label changes and accumulates the per-cluster
// sums in shared memory; a second tiny launch forms the new centres and the squared centre shift.
#include "sdb_common.cuh"

namespace {

constexpr int KM_MAX_K = 64;
constexpr int KM_MAX_KD = 4096;   // doubles of shared memory for the privatised sums

__global__ void __launch_bounds__(256) kmeans_assign_kernel(const double* __restrict__ X, const double* __restrict__ centers,
                                                            int64_t n, int d, int k, int32_t* __restrict__ labels,
                                                            int* __restrict__ changed, double* __restrict__ sums,
                                                            double* __restrict__ counts, int accumulate) {
    extern __shared__ double sh[];           // centres [k*d], csq [k], then (accumulate) sums [k*d], counts [k]
    double* c_s = sh;
    double* csq = c_s + k * d;
    double* s_s = csq + k;
    double* n_s = s_s + k * d;
    for (int t = threadIdx.x; t < k * d; t += blockDim.x) c_s[t] = centers[t];
    if (accumulate) for (int t = threadIdx.x; t < k * d + k; t += blockDim.x) s_s[t] = 0.0;
    __syncthreads();
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
        double s = 0.0;
        for (int q = 0; q < d; ++q) s += c_s[j * d + q] * c_s[j * d + q];
        csq[j] = s;
    }
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    bool any_changed = false;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double* x = X + i * d;
        // sklearn's lloyd_iter minimises  |c_j|^2 - 2 x.c_j  (the |x|^2 term is common to all j)
        double best = INFINITY;
        int bj = 0;
        for (int j = 0; j < k; ++j) {
            double dot = 0.0;
            for (int q = 0; q < d; ++q) dot += x[q] * c_s[j * d + q];
            const double v = csq[j] - 2.0 * dot;
            if (v < best) { best = v; bj = j; }
        }
        if (labels[i] != bj) { labels[i] = bj; any_changed = true; }
        if (accumulate) {
            for (int q = 0; q < d; ++q) atomicAdd(&s_s[bj * d + q], x[q]);
            atomicAdd(&n_s[bj], 1.0);
        }
    }
    if (any_changed) *changed = 1;
    if (accumulate) {
        __syncthreads();
        for (int t = threadIdx.x; t < k * d; t += blockDim.x) if (s_s[t] != 0.0) atomicAdd(sums + t, s_s[t]);
        for (int t = threadIdx.x; t < k; t += blockDim.x) if (n_s[t] != 0.0) atomicAdd(counts + t, n_s[t]);
    }
}

__global__ void kmeans_update_kernel(const double* __restrict__ sums, const double* __restrict__ counts,
                                     const double* __restrict__ centers_old, double* __restrict__ centers_new, int d, int k,
                                     double* __restrict__ shift_sq) {
    __shared__ double acc[256];
    double s = 0.0;
    for (int t = threadIdx.x; t < k * d; t += blockDim.x) {
        const int j = t / d;
        const double c = counts[j] > 0.0 ? sums[t] / counts[j] : centers_old[t];
        centers_new[t] = c;
        const double df = c - centers_old[t];
        s += df * df;
    }
    acc[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) acc[threadIdx.x] += acc[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *shift_sq = acc[0];
}

__global__ void __launch_bounds__(256) kmeans_inertia_kernel(const double* __restrict__ X, const double* __restrict__ centers,
                                                             const int32_t* __restrict__ labels, int64_t n, int d, double* out1,
                                                             void* scratch) {
    double acc[1] = {0.0};
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double* x = X + i * d;
        const double* c = centers + (int64_t)labels[i] * d;
        double s = 0.0;
        for (int q = 0; q < d; ++q) { const double df = x[q] - c[q]; s += df * df; }
        acc[0] += s;
    }
    sdb_grid_reduce<1>(acc, scratch, out1);
}

// Whole Lloyd runs on the device, one CTA per run (the n_init restarts of one fit go out in ONE launch).  For the
// sizes SpaDOT calls k-means on every epoch (a few thousand spots x 20 latent dims, k <= 20) an iteration is a few
// microseconds of arithmetic; driven from the host it cost ~130 us (two launches, three synchronising read-backs).
// Same arithmetic as kmeans_assign_kernel / kmeans_update_kernel, same stopping rules as sklearn's _kmeans_single_lloyd
// (strict label convergence, else squared centre shift <= tol, then one more E-step).  A run that meets an empty
// cluster stops with status 1 and is redone by the host-driven path, which reproduces sklearn's relocation.
__global__ void __launch_bounds__(1024) kmeans_lloyd_runs_kernel(const double* __restrict__ X, int64_t n, int d, int k,
                                                                 const double* __restrict__ centers_init, int max_iter, double tol,
                                                                 int32_t* __restrict__ labels_out, double* __restrict__ centers_out,
                                                                 double* __restrict__ inertia_out, int32_t* __restrict__ n_iter_out,
                                                                 int32_t* __restrict__ status_out) {
    extern __shared__ double sh[];           // centres [k*d], csq [k], sums [k*d], counts [k], reduction [32]
    double* c_s = sh;
    double* csq = c_s + k * d;
    double* s_s = csq + k;
    double* n_s = s_s + k * d;
    double* red = n_s + k;
    const int run = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    int32_t* labels = labels_out + (int64_t)run * n;
    for (int t = tid; t < k * d; t += nt) c_s[t] = centers_init[(int64_t)run * k * d + t];
    for (int64_t i = tid; i < n; i += nt) labels[i] = -1;
    __syncthreads();
    auto block_sum = [&](double v) -> double {           // result valid in every thread
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        __syncthreads();
        if ((tid & 31) == 0) red[tid >> 5] = v;
        __syncthreads();
        double r = 0.0;
        for (int w = 0; w < (nt >> 5); ++w) r += red[w];
        return r;
    };
    auto e_step = [&](bool accumulate) -> bool {         // labels from the current centres; true if any label moved
        for (int j = tid; j < k; j += nt) {
            double q2 = 0.0;
            for (int q = 0; q < d; ++q) q2 += c_s[j * d + q] * c_s[j * d + q];
            csq[j] = q2;
        }
        if (accumulate) for (int t = tid; t < k * d + k; t += nt) s_s[t] = 0.0;      // sums and counts are contiguous
        __syncthreads();
        bool moved = false;
        for (int64_t i = tid; i < n; i += nt) {
            const double* x = X + i * d;
            double best = INFINITY;
            int bj = 0;
            for (int j = 0; j < k; ++j) {
                double dot = 0.0;
                for (int q = 0; q < d; ++q) dot += x[q] * c_s[j * d + q];
                const double v = csq[j] - 2.0 * dot;
                if (v < best) { best = v; bj = j; }
            }
            if (labels[i] != bj) { labels[i] = bj; moved = true; }
            if (accumulate) {
                for (int q = 0; q < d; ++q) atomicAdd(&s_s[bj * d + q], x[q]);
                atomicAdd(&n_s[bj], 1.0);
            }
        }
        return __syncthreads_or(moved ? 1 : 0) != 0;
    };
    bool strict = false;
    int n_iter = 0, status = 0;
    for (int it = 0; it < max_iter; ++it) {
        n_iter = it + 1;
        const bool moved = e_step(true);
        bool empty = false;
        for (int j = 0; j < k; ++j) empty |= (n_s[j] == 0.0);
        if (empty) { status = 1; break; }                 // uniform: every thread reads the same shared counts
        double sh_part = 0.0;
        for (int t = tid; t < k * d; t += nt) {
            const double c = s_s[t] / n_s[t / d];
            const double df = c - c_s[t];
            sh_part += df * df;
            s_s[t] = c;                                     // new centre, parked until every thread has read the old one
        }
        const double shift = block_sum(sh_part);
        for (int t = tid; t < k * d; t += nt) c_s[t] = s_s[t];
        __syncthreads();
        if (!moved) { strict = true; break; }
        if (shift <= tol) break;
    }
    if (status == 0 && !strict) e_step(false);            // labels consistent with the final centres
    double part = 0.0;
    if (status == 0) {
        for (int64_t i = tid; i < n; i += nt) {
            const double* x = X + i * d;
            const double* c = c_s + (int64_t)labels[i] * d;
            double q2 = 0.0;
            for (int q = 0; q < d; ++q) { const double df = x[q] - c[q]; q2 += df * df; }
            part += q2;
        }
    }
    const double inertia = block_sum(part);
    for (int t = tid; t < k * d; t += nt) centers_out[(int64_t)run * k * d + t] = c_s[t];
    if (tid == 0) { inertia_out[run] = inertia; n_iter_out[run] = n_iter; status_out[run] = status; }
}

}  // namespace

extern "C" {

int sdb_kmeans_lloyd_runs(const double* X, int64_t n, int d, int k, const double* centers_init, int n_runs, int max_iter, double tol,
                          int32_t* labels_out, double* centers_out, double* inertia_out, int32_t* n_iter_out, int32_t* status_out,
                          void* stream) {
    SDB_CHECK_ARG(X && centers_init && labels_out && centers_out && inertia_out && n_iter_out && status_out);
    SDB_CHECK_ARG(n > 0 && d > 0 && k > 0 && n_runs > 0 && max_iter > 0);
    if (k > KM_MAX_K || k * d > KM_MAX_KD) return SDB_E_UNSUPPORTED;
    const size_t smem = sizeof(double) * (2 * (size_t)k * d + 2 * k + 32);
    static size_t smem_set = 0;
    if (smem > smem_set) {
        cudaError_t e = cudaFuncSetAttribute(kmeans_lloyd_runs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        smem_set = smem;
    }
    kmeans_lloyd_runs_kernel<<<n_runs, 1024, smem, sdb_stream(stream)>>>(X, n, d, k, centers_init, max_iter, tol, labels_out, centers_out,
                                                                        inertia_out, n_iter_out, status_out);
    SDB_LAUNCH_STATUS();
}

int sdb_kmeans_assign(const double* X, const double* centers, int64_t n, int d, int k, int32_t* labels, int* changed, double* sums,
                      double* counts, void* stream) {
    SDB_CHECK_ARG(X && centers && labels && changed && n >= 0 && d > 0 && k > 0);
    if (k > KM_MAX_K || k * d > KM_MAX_KD) return SDB_E_UNSUPPORTED;
    if (n == 0) return 0;
    const int accumulate = (sums && counts) ? 1 : 0;
    const size_t smem = sizeof(double) * ((size_t)k * d + k + (accumulate ? (size_t)k * d + k : 0));
    int grid = (int)((n + 255) / 256);
    if (grid > 592) grid = 592;
    static size_t smem_set = 0;
    if (smem > smem_set) {
        cudaError_t e = cudaFuncSetAttribute(kmeans_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        smem_set = smem;
    }
    kmeans_assign_kernel<<<grid, 256, smem, sdb_stream(stream)>>>(X, centers, n, d, k, labels, changed, sums, counts, accumulate);
    SDB_LAUNCH_STATUS();
}

int sdb_kmeans_update(const double* sums, const double* counts, const double* centers_old, double* centers_new, int d, int k,
                      double* shift_sq, void* stream) {
    SDB_CHECK_ARG(sums && counts && centers_old && centers_new && shift_sq && d > 0 && k > 0);
    kmeans_update_kernel<<<1, 256, 0, sdb_stream(stream)>>>(sums, counts, centers_old, centers_new, d, k, shift_sq);
    SDB_LAUNCH_STATUS();
}

int sdb_kmeans_inertia(const double* X, const double* centers, const int32_t* labels, int64_t n, int d, double* out1, void* scratch,
                       void* stream) {
    SDB_CHECK_ARG(X && centers && labels && out1 && scratch && d > 0);
    kmeans_inertia_kernel<<<sdb_reduce_grid(n), 256, 0, sdb_stream(stream)>>>(X, centers, labels, n, d, out1, scratch);
    SDB_LAUNCH_STATUS();
}

}  // extern "C"
