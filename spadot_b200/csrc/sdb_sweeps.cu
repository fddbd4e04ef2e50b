// Native sweep loop: `n_sweeps` full Sinkhorn iterations issued from C in one call.
//
// The reference keeps its inner loop native for the same reason (step1_process<T> runs `iters` updates per
// call, ref: SpaDOT/utils/OT_loss/ot_func.cpp:690-828): at ChickenHeart sizes a streamed pass takes a few
// microseconds of GPU time, so a Python-driven launch sequence (5 launches x ~13 us of interpreter + ctypes
// per iteration) is host-bound.  This entry point issues pass -> finalize/update -> pass -> finalize/update ->
// absorb for every iteration back to back on the caller's stream (single rank; the row-partitioned solve
// needs a collective inside the iteration and is driven from Python).
#include "sdb_common.cuh"

#include <cstdlib>

thread_local int sdb_pdl = 0;

namespace {
bool pdl_enabled() {                 // development knob: SDB_PDL=0 launches the loop's kernels without overlap
    static int on = -1;
    if (on < 0) { const char* e = getenv("SDB_PDL"); on = (e && *e == '0') ? 0 : 1; }
    return on == 1;
}
struct PdlScope {
    explicit PdlScope(bool on) { sdb_pdl = on ? 1 : 0; }
    ~PdlScope() { sdb_pdl = 0; }
};
}  // namespace

extern "C" int sdb_sinkhorn_sweeps(const sdb_sweep_desc* d, int n_sweeps, int first_tick, int lr_known_first, void* stream) {
    SDB_CHECK_ARG(d && n_sweeps >= 0 && d->n > 0 && d->m > 0 && d->eps > 0.0);
    const double c1 = d->inv_med / d->eps;
    const double scale = 2.0 * c1 * SDB_LOG2E;      // tensor-core form: exponent = bias_j + scale * x_i.y_j (norms inside the bias)
    // SIMT form: dot-product tiles with the same scale, or direct-difference tiles (negative scale, zero norm vectors)
    const double scale_direct = d->simt_direct ? -c1 * SDB_LOG2E : scale;
    const double log_m = log((double)d->m), log_N = log((double)d->n_total);
    int rc = 0;
    PdlScope pdl(pdl_enabled() && d->use_tc);      // the SIMT pass kernel is not part of the PDL chain
    auto pass = [&](bool row, bool predicted) -> int {
        if (d->use_tc) {
            return row ? sdb_lse_pass_tc_pred(d->x16, d->n, d->n_pad, d->y16, d->m, d->m_pad, d->dp, d->bias_y,
                                              (float)(scale * d->pow2_scale), d->tps_row, d->n_ctas, predicted ? d->m_x : nullptr,
                                              d->partial_row, stream)
                       : sdb_lse_pass_tc_pred(d->y16, d->m, d->m_pad, d->x16, d->n, d->n_pad, d->dp, d->bias_x,
                                              (float)(scale * d->pow2_scale), d->tps_col, d->n_ctas, predicted ? d->m_y : nullptr,
                                              d->partial_col, stream);
        }
        return row ? sdb_lse_pass_simt(d->xt, d->ldx, d->n, d->yt, d->ldy, d->m, d->dpad, d->bias_y, scale_direct, d->bounds_row,
                                       d->ns_row, d->partial_row, stream)
                   : sdb_lse_pass_simt(d->yt, d->ldy, d->m, d->xt, d->ldx, d->n, d->dpad, d->bias_x, scale_direct, d->bounds_col,
                                       d->ns_col, d->partial_col, stream);
    };
    const bool pred = d->use_tc && d->m_x && d->m_y && d->bad_flag;
    if (d->use_tc && d->flag2 && d->tile_counters) {
        // fused pass + update: two launches per iteration (the CTA that completes a row tile updates it), tau bookkeeping
        // deferred to the next update of each row, flushed once at the end of the batch
        const int64_t row_tiles = (d->n + 127) / 128;
        auto fused = [&](bool row, bool predicted, int tick) -> int {
            sdb_tc_update u{};
            u.counters = row ? d->tile_counters : d->tile_counters + row_tiles;
            u.norms = row ? d->norms_x : d->norms_y;
            u.c1 = c1;
            u.L = row ? d->Lr : d->Lc;
            u.logmarg = row ? d->logp : d->logq;
            u.eps = d->eps;
            u.alpha = row ? d->alpha1 : d->alpha2;
            u.log_n_other = row ? log_m : log_N;
            u.pot = row ? d->f : d->g;
            u.frame = row ? d->u : d->v;
            u.la_old = row ? d->la_old : d->lb_old;
            u.bias_out = row ? d->bias_x : d->bias_y;
            u.flag2 = d->flag2;
            u.tick = tick;
            u.log_tau = d->log_tau;
            u.log_floor = d->log_floor;
            u.m_next = pred ? (row ? d->m_x : d->m_y) : nullptr;
            u.bad_flag = pred ? d->bad_flag : nullptr;
            return row ? sdb_lse_pass_tc_fused(d->x16, d->n, d->n_pad, d->y16, d->m, d->m_pad, d->dp, d->bias_y,
                                               (float)(scale * d->pow2_scale), d->tps_row, d->n_ctas, predicted ? d->m_x : nullptr,
                                               d->partial_row, &u, stream)
                       : sdb_lse_pass_tc_fused(d->y16, d->m, d->m_pad, d->x16, d->n, d->n_pad, d->dp, d->bias_x,
                                               (float)(scale * d->pow2_scale), d->tps_col, d->n_ctas, predicted ? d->m_y : nullptr,
                                               d->partial_col, &u, stream);
        };
        if (!(lr_known_first && n_sweeps > 0)) {
            rc = sdb_make_bias(d->m, d->m_bias, d->g, d->norms_y, d->eps, c1, d->bias_y, stream);
            if (rc) return rc;
        }
        for (int i = 0; i < n_sweeps; ++i) {
            const int tick = first_tick + i;
            if (i == 0 && lr_known_first)
                rc = sdb_potential_update_deferred(d->n, d->Lr, d->logp, d->norms_x, d->eps, d->alpha1, log_m, c1, d->f, d->u, d->la_old,
                                                   d->bias_x, d->flag2, tick, d->log_tau, d->log_floor, stream);
            else
                rc = fused(true, pred && i >= d->pred_from_row, tick);
            if (rc) return rc;
            rc = fused(false, pred && i >= d->pred_from_col, tick);
            if (rc) return rc;
        }
        if (n_sweeps > 0) rc = sdb_absorb_pending(d->n, d->m, d->flag2, first_tick + n_sweeps - 1, d->f, d->g, d->u, d->v, d->flag, stream);
        return rc;
    }
    if (!(lr_known_first && n_sweeps > 0)) {
        // bias of the first row pass from the current g (a previous call may have used another eps)
        rc = sdb_make_bias(d->m, d->m_bias, d->g, d->norms_y, d->eps, c1, d->bias_y, stream);
        if (rc) return rc;
    }
    for (int i = 0; i < n_sweeps; ++i) {
        const int tick = first_tick + i;
        if (i == 0 && lr_known_first) {
            // Lr already holds the row LSE at the current g (final-stage gap check)
            rc = sdb_potential_update(d->n, d->Lr, d->logp, d->norms_x, d->eps, d->alpha1, log_m, c1, d->f, d->u, d->la_old, d->bias_x,
                                      d->flag, tick, d->log_tau, d->log_floor, stream);
        } else {
            rc = pass(true, pred && i >= d->pred_from_row);
            if (rc) return rc;
            rc = sdb_finalize_update_pred(d->partial_row, d->ns_row, d->n, d->norms_x, c1, d->Lr, d->logp, d->eps, d->alpha1, log_m, d->f,
                                          d->u, d->la_old, d->bias_x, d->flag, tick, d->log_tau, d->log_floor, pred ? d->m_x : nullptr,
                                          pred ? d->bad_flag : nullptr, stream);
        }
        if (rc) return rc;
        rc = pass(false, pred && i >= d->pred_from_col);
        if (rc) return rc;
        rc = sdb_finalize_update_pred(d->partial_col, d->ns_col, d->m, d->norms_y, c1, d->Lc, d->logq, d->eps, d->alpha2, log_N, d->g,
                                      d->v, d->lb_old, d->bias_y, d->flag, tick, d->log_tau, d->log_floor, pred ? d->m_y : nullptr,
                                      pred ? d->bad_flag : nullptr, stream);
        if (rc) return rc;
        rc = sdb_absorb(d->n, d->m, d->flag, tick, d->f, d->g, d->u, d->v, stream);
        if (rc) return rc;
    }
    return 0;
}
