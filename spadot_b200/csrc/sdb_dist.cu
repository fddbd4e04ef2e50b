// Row-partitioned solve, native loop ("NCCL communicator handle for multi-GPU variants").
//
// One process per GPU; every rank holds its rows of x and all of y.  An iteration is
//   row pass + update of the local f      (local, tensor-core or SIMT form, predicted stabiliser where valid)
//   column pass against the COMMON shift  (row_m = previous combined column LSE in log2 units + 1, identical on every rank)
//   sdb_partial_sums_f64                  (per-rank sums + tau / verification riders, M + 2 doubles)
//   ncclAllReduce(sum)                    (the one collective of the iteration, on the caller's stream)
//   sdb_update_from_sums_f64              (identical g, bias, shift, flags on every rank)
//   sdb_absorb
// issued back to back from C, so there is no per-iteration interpreter or framework cost between the kernels and the
// collective.  NCCL is bound at run time from the libnccl.so.2 the process already has (the one PyTorch loaded): the library
// has no link-time dependency on it, and single-GPU use never touches it.
#include "sdb_common.cuh"

#include <dlfcn.h>
#include <string.h>

namespace {

typedef struct { char internal[128]; } nccl_unique_id;           // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128)
typedef void* nccl_comm_t;
typedef int (*fn_get_unique_id)(nccl_unique_id*);
typedef int (*fn_comm_init_rank)(nccl_comm_t*, int, nccl_unique_id, int);
typedef int (*fn_comm_destroy)(nccl_comm_t);
typedef int (*fn_all_reduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t);
typedef const char* (*fn_error_string)(int);
constexpr int NCCL_FLOAT64 = 8, NCCL_SUM = 0;                     // ncclFloat64, ncclSum (nccl.h)

struct NcclApi {
    void* handle = nullptr;
    fn_get_unique_id get_unique_id = nullptr;
    fn_comm_init_rank comm_init_rank = nullptr;
    fn_comm_destroy comm_destroy = nullptr;
    fn_all_reduce all_reduce = nullptr;
    fn_error_string error_string = nullptr;
    bool ok = false;
};

NcclApi& nccl() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);      // the copy already mapped into the process (PyTorch's)
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW);
    if (!h) return api;
    api.handle = h;
    api.get_unique_id = (fn_get_unique_id)dlsym(h, "ncclGetUniqueId");
    api.comm_init_rank = (fn_comm_init_rank)dlsym(h, "ncclCommInitRank");
    api.comm_destroy = (fn_comm_destroy)dlsym(h, "ncclCommDestroy");
    api.all_reduce = (fn_all_reduce)dlsym(h, "ncclAllReduce");
    api.error_string = (fn_error_string)dlsym(h, "ncclGetErrorString");
    api.ok = api.get_unique_id && api.comm_init_rank && api.comm_destroy && api.all_reduce;
    return api;
}

}  // namespace

extern "C" {

int sdb_nccl_available(void) { return nccl().ok ? 1 : 0; }

int sdb_nccl_unique_id(void* id128_host) {
    SDB_CHECK_ARG(id128_host);
    if (!nccl().ok) return SDB_E_DRIVER;
    nccl_unique_id id;
    const int r = nccl().get_unique_id(&id);
    if (r != 0) return SDB_E_NCCL;
    memcpy(id128_host, id.internal, 128);
    return 0;
}

int sdb_nccl_comm_create(const void* id128_host, int world, int rank, void** comm_out) {
    SDB_CHECK_ARG(id128_host && comm_out && world >= 1 && rank >= 0 && rank < world);
    if (!nccl().ok) return SDB_E_DRIVER;
    nccl_unique_id id;
    memcpy(id.internal, id128_host, 128);
    nccl_comm_t comm = nullptr;
    const int r = nccl().comm_init_rank(&comm, world, id, rank);
    if (r != 0) return SDB_E_NCCL;
    *comm_out = comm;
    return 0;
}

int sdb_nccl_comm_destroy(void* comm) {
    if (!comm) return 0;
    if (!nccl().ok) return SDB_E_DRIVER;
    return nccl().comm_destroy(comm) == 0 ? 0 : SDB_E_NCCL;
}

int sdb_nccl_allreduce_sum_f64(void* comm, double* buf, int64_t count, void* stream) {
    SDB_CHECK_ARG(comm && buf && count >= 0);
    if (!nccl().ok) return SDB_E_DRIVER;
    return nccl().all_reduce(buf, buf, (size_t)count, NCCL_FLOAT64, NCCL_SUM, comm, sdb_stream(stream)) == 0 ? 0 : SDB_E_NCCL;
}

// n_sweeps iterations of the row-partitioned solve.  Preconditions (the Python driver establishes them with the first
// iteration of an epsilon stage, which combines with max-then-sum): `shift` (= d->m_y, M floats) holds the common per-column
// shift for this (eps, median), d->bias_y the bias of the current g, predictions for the local rows in d->m_x are valid
// from sweep d->pred_from_row on (or d->m_x == NULL: the row pass tracks its maximum).  sums: M + 2 doubles of workspace.
int sdb_sinkhorn_sweeps_dist(const sdb_sweep_desc* d, void* comm, double* sums, int n_sweeps, int first_tick, void* stream) {
    SDB_CHECK_ARG(d && comm && sums && n_sweeps >= 0 && d->m > 0 && d->n >= 0 && d->eps > 0.0);
    SDB_CHECK_ARG(d->m_y && d->bad_flag && d->flag && d->partial_col && d->g && d->v && d->lb_old && d->Lc && d->logq && d->bias_y);
    if (!nccl().ok) return SDB_E_DRIVER;
    const double c1 = d->inv_med / d->eps;
    const double scale = 2.0 * c1 * SDB_LOG2E;
    const double log_m = log((double)d->m), log_N = log((double)d->n_total);
    const bool tc = d->use_tc != 0;
    const double simt_scale = d->simt_direct ? -c1 * SDB_LOG2E : scale;
    int rc = 0;
    for (int i = 0; i < n_sweeps; ++i) {
        const int tick = first_tick + i;
        if (d->n > 0) {
            // ---- row pass + update of the local f
            const bool pred_row = tc && d->m_x && i >= d->pred_from_row;
            if (tc)
                rc = sdb_lse_pass_tc_pred(d->x16, d->n, d->n_pad, d->y16, d->m, d->m_pad, d->dp, d->bias_y, (float)(scale * d->pow2_scale),
                                          d->tps_row, d->n_ctas, pred_row ? d->m_x : nullptr, d->partial_row, stream);
            else
                rc = sdb_lse_pass_simt(d->xt, d->ldx, d->n, d->yt, d->ldy, d->m, d->dpad, d->bias_y, simt_scale, d->bounds_row, d->ns_row,
                                       d->partial_row, stream);
            if (rc) return rc;
            rc = sdb_finalize_update_pred(d->partial_row, d->ns_row, d->n, d->norms_x, c1, d->Lr, d->logp, d->eps, d->alpha1, log_m, d->f, d->u,
                                          d->la_old, d->bias_x, d->flag, tick, d->log_tau, d->log_floor, tc ? d->m_x : nullptr,
                                          tc ? d->bad_flag : nullptr, stream);
            if (rc) return rc;
            // ---- column pass over the local rows against the common shift
            if (tc)
                rc = sdb_lse_pass_tc_pred(d->y16, d->m, d->m_pad, d->x16, d->n, d->n_pad, d->dp, d->bias_x, (float)(scale * d->pow2_scale),
                                          d->tps_col, d->n_ctas, d->m_y, d->partial_col, stream);
            else
                rc = sdb_lse_pass_simt(d->yt, d->ldy, d->m, d->xt, d->ldx, d->n, d->dpad, d->bias_x, simt_scale, d->bounds_col, d->ns_col,
                                       d->partial_col, stream);
            if (rc) return rc;
            rc = sdb_partial_sums_f64(d->partial_col, d->ns_col, d->m, d->m_y, sums, d->flag, tick, d->bad_flag, stream);
        } else {
            rc = (int)cudaMemsetAsync(sums, 0, sizeof(double) * (size_t)(d->m + 2), sdb_stream(stream));   // a rank without rows adds nothing
        }
        if (rc) return rc;
        if (nccl().all_reduce(sums, sums, (size_t)(d->m + 2), NCCL_FLOAT64, NCCL_SUM, comm, sdb_stream(stream)) != 0) return SDB_E_NCCL;
        rc = sdb_update_from_sums_f64(sums, d->m_y, d->m, d->norms_y, c1, d->Lc, d->logq, d->eps, d->alpha2, log_N, d->g, d->v, d->lb_old,
                                      d->bias_y, d->flag, tick, d->log_tau, d->log_floor, d->bad_flag, stream);
        if (rc) return rc;
        rc = sdb_absorb(d->n, d->m, d->flag, tick, d->f, d->g, d->u, d->v, stream);
        if (rc) return rc;
    }
    return 0;
}

}  // extern "C"
