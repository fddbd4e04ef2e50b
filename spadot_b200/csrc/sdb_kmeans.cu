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

}  // namespace

extern "C" {

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
