/*
 * spadot_b200.h — C ABI of libspadot_b200.so (sm_100a).
 *
 * Drop-in boundary for SpaDOT's optimal-transport hot path.  Every
 * entry point is `extern "C"`, takes plain pointers / sizes / scalars, returns an int
 * status (0 = ok, >0 = cudaError_t, <0 = SDB_E_*; sdb_error_string() decodes) and is
 * asynchronous on the `stream` argument (a cudaStream_t passed as void*).  The caller
 * owns every buffer including workspaces; nothing is allocated behind its back and
 * there is no global state.  Pointers are DEVICE pointers unless the name ends in
 * `_host`.  There is no CPU fallback anywhere in this library.
 *
 * Citations `ref:` are relative to /root/reference/SpaDOT/.
 *
 * Conventions of the streamed ("flash") Sinkhorn:
 *   - points are centred and stored feature-major fp32:  xt[k*ld + i]  (k < dpad, i < n),
 *     ld a multiple of 64 and >= round_up(n,64)+64, padding zero-filled;
 *   - cost  C_ij = (|x_i|^2 + |y_j|^2 - 2 x_i.y_j) * inv_med      (ref: utils/OT_loss/ot_solvers.py:102-103)
 *   - a pass over rows i and columns j computes, in base 2,
 *         L2_i = log2 sum_j 2^( bias_j + scale * x_i.y_j ),   scale = 2*c1*log2(e),  c1 = inv_med/eps
 *         bias_j = log2(e) * ( g_j/eps - |y_j|^2 * c1 )
 *     which is the reference's  K(b*dy)  matvec (ref: utils/OT_loss/ot_func.cpp:43-170, 610-636)
 *     without ever forming K; the column pass is the same kernel with the roles swapped
 *     (ref: ot_func.cpp:173-249, 642-668).
 *   - potentials f,g, norms, marginals and every N- or M-length reduction are fp64.
 */
#ifndef SPADOT_B200_H
#define SPADOT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDB_VERSION 100

#define SDB_E_INVALID   (-1)  /* bad argument (null pointer, size, alignment) */
#define SDB_E_UNSUPPORTED (-2)/* shape outside what the kernel was built for */
#define SDB_E_NOTFINITE (-3)  /* NaN / overflow detected (ref: ot_solvers.py:446-447) */
#define SDB_E_DRIVER    (-4)  /* CUDA driver entry point / libnccl unavailable */
#define SDB_E_NCCL      (-5)  /* an NCCL call failed */

#define SDB_NEG_SENTINEL (-1.0e30f) /* finite stand-in for -inf in fp32 bias vectors */

int sdb_version(void);
const char* sdb_error_string(int status);
/* 0 when a CUDA device of compute capability 10.x is usable, else a status. */
int sdb_device_check(void);

/* ------------------------------------------------------------------ point preparation */
/* out[k] += sum_i x[i*d+k]   (x row-major n x d, fp64).  out must be zeroed by the caller. */
int sdb_column_sums_f64(const double* x, int64_t n, int d, double* out, void* stream);
/* xt[k*ld+i] = (float)(x[i*d+k] - center[k]);  norms[i] = sum_k (double)xt[k*ld+i]^2
 * (norms of the ROUNDED coordinates, so the streamed cost is the exact squared distance of
 * the points the kernels actually see).  Rows k in [d,dpad) and columns i in [n,ld) are zeroed. */
int sdb_prep_points_f64(const double* x, int64_t n, int d, const double* center,
                        float* xt, int64_t ld, int dpad, double* norms, void* stream);

/* ------------------------------------------------------------------ K3: streamed LSE pass (SIMT fp32) */
/* partial[(s*n_p + i)*2 + {0,1}] = (max, sum) of 2^(bias_j + scale*A_ij - max) over the columns of split s, where the sign
 * of `scale` selects the tile arithmetic:
 *   scale > 0: A_ij = x_i.y_j  (dot-product form, scale = 2*c1*log2(e), bias_j = log2(e)*(g_j/eps - |y_j|^2*c1); the row
 *              norm is subtracted by the finalize kernels).  Half the inner-loop work, but its fp32 rounding is relative to
 *              |x||y|: fine while scale*|x||y| stays below ~40 (eps = 0.05 on median-normalised costs).
 *   scale < 0: A_ij = |x_i - y_j|^2 by direct differences like scipy's cdist (ref: ot_solvers.py:102), scale = -c1*log2(e),
 *              bias_j = log2(e)*g_j/eps: pass an all-zero `norms` vector to sdb_make_bias, sdb_lse_finalize* and
 *              sdb_potential_update.  Rounding relative to the COST of a pair: what small eps needs (at eps = 0.01 the dot
 *              form put 1.3e-5 on the plan's marginals, this form 1e-6).
 * Split s covers the
 * columns j in [split_bounds[s], split_bounds[s+1]) (device array of n_splits+1 entries; keep
 * each split <= 65536 columns so the fp32 running sums stay below 1e-6 relative error).
 * dpad <= 128, dpad % 4 == 0.  Grid = ceil(n_p/64) x n_splits CTAs of 256 threads.  `scale` is a double: the kernel applies
 * it as an fp32 (hi, lo) pair, t = fma(hi, d, fma(lo, d, bias)), so its rounding is not a bias common to every term.
 * replaces: gemv/gemtv + update_k of ref: utils/OT_loss/ot_func.cpp:43-249,547-568. */
int sdb_lse_pass_simt(const float* pt, int64_t ldp, int64_t n_p,
                      const float* qt, int64_t ldq, int64_t n_q, int dpad,
                      const float* bias, double scale,
                      const int64_t* split_bounds, int n_splits,
                      float* partial, void* stream);

/* ------------------------------------------------------------------ K3: streamed LSE pass (tcgen05 / TMEM / TMA) */
/* max over all entries of |x[i*d+k] - center[k]|, written as a double (out must be zeroed). */
int sdb_absmax_centered_f64(const double* x, int64_t n, int d, const double* center, double* out_max, void* stream);
/* fp16 hi/lo split of the centred points scaled by 2^pow2_exp:  out16[i*(2*dp) + k] = hi, out16[i*(2*dp) + dp + k] = lo
 * with hi = fp16(v), lo = fp16(v - hi), v = (x[i*d+k]-center[k]) * 2^pow2_exp (|v| must stay < 32768);
 * norms[i] = sum_k ((hi+lo) * 2^-pow2_exp)^2 in fp64.  dp in {16,32,64} >= d; rows [n, n_pad) are zero; n_pad % 256 == 0. */
int sdb_prep_points_split_f16(const double* x, int64_t n, int d, const double* center, int pow2_exp,
                              void* out16, int64_t n_pad, int dp, double* norms, void* stream);
/* The same with an arbitrary positive prescale q instead of 2^pow2_exp: v = (x-center)*q, norms of (hi+lo)/q.  The streamed
 * solver folds the whole exponent scale into the points, q^2 * S = 2*c1*log2(e) with S a power of two, so that the `scale`
 * argument of sdb_lse_pass_tc is exactly representable in fp32 (an fp32-rounded scale is a 2^-24 relative error common to
 * every term of every row: a systematic bias of the plan that grows like 1/eps). */
int sdb_prep_points_split_f16_scaled(const double* x, int64_t n, int d, const double* center, double prescale,
                                     void* out16, int64_t n_pad, int dp, double* norms, void* stream);
/* Same contract as sdb_lse_pass_simt with scale already multiplied by 2^(-2*pow2_exp); splits are
 * runs of `tiles_per_split` 256-column tiles: n_splits = ceil(ceil(n_q/256)/tiles_per_split) and
 * partial holds n_splits*n_p (max,sum) pairs.  bias_padded has n_q_pad entries (see sdb_make_bias).
 * Persistent grid of n_ctas CTAs (use the SM count), 320 threads, one CTA per SM, 512 TMEM columns.
 * x.y^T = hi*hi + hi*lo + lo*hi on tcgen05.mma kind::f16 with fp32 accumulation in TMEM. */
int sdb_lse_pass_tc(const void* p16, int64_t n_p, int64_t n_p_pad, const void* q16, int64_t n_q, int64_t n_q_pad,
                    int dp, const float* bias_padded, float scale, int tiles_per_split, int n_ctas,
                    float* partial, void* stream);

/* The same pass with VARIABLE-length splits: split s covers the 256-column tiles [split_tiles[s], split_tiles[s+1]) (device
 * array of n_splits + 1 ints, every split non-empty, split_tiles[n_splits] = n_q_pad / 256).  Used by the transition table
 * (K6): columns permuted by domain label, every label group padded to a tile multiple (padding bias = SDB_NEG_SENTINEL), one
 * (max, sum) per (label, row).  replaces the dense P0^T.T.P1 of ref: utils/_analyze_utils.py:135-137. */
int sdb_lse_pass_tc_groups(const void* p16, int64_t n_p, int64_t n_p_pad, const void* q16, int64_t n_q_pad, int dp,
                           const float* bias_padded, float scale, const int* split_tiles, int n_splits, int n_ctas, float* partial,
                           void* stream);

/* L[i] = ln2 * log2( sum_s sum_s,i * 2^(max_s,i - M_i) ) + ln2*M_i - norms[i]*c1   (fp64)
 *      = LSE_j[(g_j - C_ij)/eps]; -inf when every partial is empty. */
int sdb_lse_finalize(const float* partial, int n_splits, int64_t n, const double* norms,
                     double c1, double* L, void* stream);
/* Predicted stabiliser.  A pass may take, per row, an upper bound m_i of its largest exponent (row_m) instead of tracking the
 * running maximum (one max tree + stabiliser update per 32-column chunk: ~16 % of the epilogue's instructions, and the
 * epilogue is issue-bound): the partials then hold (m_i, sum_j 2^(t_ij - m_i)).  The *_pred finalizers below write the
 * prediction for the NEXT pass over the same rows, m_next[i] = M + log2 S + 1 (an upper bound of this pass's largest
 * exponent), and OR 1 into *bad_flag unless 2^-60 < S < 2^100 (no term overflowed, the dominant terms were not flushed to
 * zero); a caller that passed row_m must check the flag and, if it is set, redo the pass with tracking (row_m = NULL).
 * m_next / bad_flag may be NULL; with both NULL the *_pred entry points are the plain ones. */
int sdb_lse_finalize_pred(const float* partial, int n_splits, int64_t n, const double* norms, double c1, double* L,
                          float* m_next, int* bad_flag, void* stream);
int sdb_finalize_update_pred(const float* partial, int n_splits, int64_t n, const double* norms, double c1, double* L,
                             const double* logmarg, double eps, double alpha, double log_n_other, double* pot,
                             const double* frame, double* la_old, float* bias, int* absorb_flag, int iter,
                             double log_tau, double log_floor, float* m_next, int* bad_flag, void* stream);
/* Row-partitioned solve, column step with ONE collective.  Every rank runs its column pass against the SAME
 * per-column shift (row_m of sdb_lse_pass_tc_pred = the previous combined column LSE in log2 units + 1, kept in `shift`), so
 * the per-rank sums are addable:
 *   sdb_partial_sums_f64:     sums[j] = sum over splits of partial.sum * 2^(partial.max - shift[j])  (j < n), and two riders
 *                             sums[n] = (*absorb_flag == iter), sums[n+1] = (*bad_flag != 0)   [n+2 doubles]
 *   caller:                   all-reduce(SUM) of the n+2 doubles (NCCL)
 *   sdb_update_from_sums_f64: L[j] = ln2*(shift[j] + log2 sums[j]) - norms[j]*c1, then exactly sdb_potential_update on it,
 *                             shift[j] <- shift[j] + log2 sums[j] + 1 for the next iteration, *bad_flag |= 1 unless
 *                             2^-60 < sums[j] < 2^100 or if sums[n+1] > 0, atomicMax(absorb_flag, iter) if sums[n] > 0.
 * Identical inputs on every rank => identical g, bias, shift and flags on every rank without further exchange. */
int sdb_partial_sums_f64(const float* partial, int n_splits, int64_t n, const float* shift, double* sums, const int* absorb_flag,
                         int iter, const int* bad_flag, void* stream);
int sdb_update_from_sums_f64(const double* sums, float* shift, int64_t n, const double* norms, double c1, double* L,
                             const double* logmarg, double eps, double alpha, double log_n_other, double* pot, const double* frame,
                             double* la_old, float* bias, int* absorb_flag, int iter, double log_tau, double log_floor, int* bad_flag,
                             void* stream);
/* FUSED pass + update (tensor-core form): the CTA that finishes the LAST column split of a 128-row tile (per-tile arrival
 * counter) combines that tile's partials and performs exactly sdb_finalize_update_pred on its rows, so an iteration of the
 * native loop is two launches (row pass, column pass) instead of five; tau bookkeeping is deferred (see
 * sdb_update_row_deferred in csrc/sdb_common.cuh): flag2 = two ints (ping-pong by tick parity, initialise to -1), and
 * sdb_absorb_pending flushes the last tick before anything reads the frames.  counters: ceil(n_p/128) zero-initialised uints
 * (left zero). */
typedef struct sdb_tc_update {
    unsigned int* counters;
    const double* norms; double c1;
    double* L; const double* logmarg; double eps, alpha, log_n_other;
    double* pot; double* frame; double* la_old; float* bias_out;
    int* flag2; int32_t tick, reserved; double log_tau, log_floor;
    float* m_next; int* bad_flag;
} sdb_tc_update;
int sdb_lse_pass_tc_fused(const void* p16, int64_t n_p, int64_t n_p_pad, const void* q16, int64_t n_q, int64_t n_q_pad,
                          int dp, const float* bias_padded, float scale, int tiles_per_split, int n_ctas,
                          const float* row_m, float* partial, const sdb_tc_update* upd, void* stream);
/* pot <- update from a known LSE (as sdb_potential_update) with the deferred tau bookkeeping. */
int sdb_potential_update_deferred(int64_t n, const double* L, const double* logmarg, const double* norms, double eps, double alpha,
                                  double log_n_other, double c1, double* pot, double* frame, double* la_old, float* bias,
                                  int* flag2, int iter, double log_tau, double log_floor, void* stream);
/* if flag2[last_tick & 1] == last_tick: u <- f, v <- g, *flag = last_tick (the legacy single flag, may be NULL). */
int sdb_absorb_pending(int64_t n, int64_t m, const int* flag2, int last_tick, const double* f, const double* g, double* u,
                       double* v, int* flag, void* stream);
/* sdb_lse_pass_tc with the predicted stabiliser: row_m[i] for i < n_p (NULL = track the maximum, i.e. sdb_lse_pass_tc). */
int sdb_lse_pass_tc_pred(const void* p16, int64_t n_p, int64_t n_p_pad, const void* q16, int64_t n_q, int64_t n_q_pad,
                         int dp, const float* bias_padded, float scale, int tiles_per_split, int n_ctas,
                         const float* row_m, float* partial, void* stream);

/* Potential update of one side (ref: ot_func.cpp:610-668 in total potentials):
 *   la_old[i] = (pot[i] - frame[i]) / eps                       (log of the reference's old_a)
 *   pot[i]    = eps*alpha*( logmarg[i] - LA_i ),  LA_i = L[i] - log_n_other, or with the fixed-schedule
 *               solver's 1e-10 floor (ref: ot_solvers.py:498-502): logaddexp(LA_i, log_floor - frame[i]/eps)
 *               (pass log_floor = -inf to disable)
 *   bias[i]   = log2(e)*( pot[i]/eps - norms[i]*c1 )            (fp32, clamped to SDB_NEG_SENTINEL)
 *   if (pot[i]-frame[i])/eps > log_tau: atomicMax(absorb_flag, iter)   (ref: ot_func.cpp:778-790)
 * la_old / bias / absorb_flag may be NULL. */
int sdb_potential_update(int64_t n, const double* L, const double* logmarg, const double* norms,
                         double eps, double alpha, double log_n_other, double c1,
                         double* pot, const double* frame, double* la_old, float* bias,
                         int* absorb_flag, int iter, double log_tau, double log_floor, void* stream);
/* sdb_lse_finalize + sdb_potential_update fused (all outputs mandatory): combines the partials into L, updates pot,
 * la_old, the absorb flag, and writes bias[i] (i < n only) = the bias vector of the NEXT pass over the other side. */
int sdb_finalize_update(const float* partial, int n_splits, int64_t n, const double* norms, double c1, double* L,
                        const double* logmarg, double eps, double alpha, double log_n_other, double* pot,
                        const double* frame, double* la_old, float* bias, int* absorb_flag, int iter,
                        double log_tau, double log_floor, void* stream);
/* bias[i] = log2(e)*( pot[i]/eps - norms[i]*c1 ) for i < n (pot may be NULL = 0);
 * bias[i] = SDB_NEG_SENTINEL for n <= i < n_pad (column padding of the tensor-core pass). */
int sdb_make_bias(int64_t n, int64_t n_pad, const double* pot, const double* norms, double eps, double c1,
                  float* bias, void* stream);
/* if (*absorb_flag == iter) { u = f; v = g; }      (ref: ot_func.cpp:792-819) */
int sdb_absorb(int64_t n, int64_t m, const int* absorb_flag, int iter,
               const double* f, const double* g, double* u, double* v, void* stream);

/* ------------------------------------------------------------------ native sweep loop (single rank) */
/* Everything one Sinkhorn iteration touches, so that `n_sweeps` iterations can be issued from C in one call
 * (the reference's step1_process_double runs `iters` updates per call, ref: ot_func.cpp:690-828, :1261).
 * All pointers are device pointers owned by the caller. */
typedef struct sdb_sweep_desc {
    int64_t n, m, n_total;            /* local source spots, target spots, source spots over all ranks (= n here) */
    int32_t use_tc, dpad, dp, n_ctas; /* pass form; SIMT feature padding; tensor-core K per term; persistent grid */
    const float* xt; int64_t ldx;     /* SIMT operands (sdb_prep_points_f64) */
    const float* yt; int64_t ldy;
    const void* x16; int64_t n_pad;   /* tensor-core operands (sdb_prep_points_split_f16) */
    const void* y16; int64_t m_pad;
    const double* norms_x; const double* norms_y;     /* norms of the representation in use */
    const int64_t* bounds_row; const int64_t* bounds_col;   /* SIMT column splits of the row / column pass */
    int32_t ns_row, ns_col, tps_row, tps_col;          /* number of splits; tiles per split (tensor-core form) */
    float* partial_row; float* partial_col;            /* (ns_row, n, 2) and (ns_col, m, 2) workspaces */
    float* bias_x; float* bias_y; int64_t m_bias;      /* bias vectors (padded to 256); padded length of bias_y */
    double *f, *g, *u, *v, *la_old, *lb_old, *Lr, *Lc;
    const double* logp; const double* logq;
    int* flag;
    double eps, inv_med, alpha1, alpha2, log_tau, log_floor, pow2_scale;   /* pow2_scale = 2^(-2*pow2_exp) */
    /* predicted stabiliser (tensor-core form; see sdb_lse_pass_tc_pred): per-row predictions of the row / column pass,
     * written by every finalize when non-NULL; sweep i uses them in its row pass if i >= pred_from_row and in its column
     * pass if i >= pred_from_col.  The caller checks *bad_flag after the call and redoes the batch with pred_from_* past
     * n_sweeps if it is set. */
    float* m_x; float* m_y; int* bad_flag;
    int32_t pred_from_row, pred_from_col;
    /* SIMT form only: 0 = dot-product tiles (norms_x / norms_y hold |x|^2, |y|^2), 1 = direct-difference tiles (they hold
     * zeros); see sdb_lse_pass_simt. */
    int32_t simt_direct, reserved0;
    /* tensor-core form: non-NULL flag2 (2 ints) + tile_counters (ceil(n/128) + ceil(m/128) uints) select the fused
     * pass + update loop (two launches per iteration, see sdb_lse_pass_tc_fused). */
    int* flag2; unsigned int* tile_counters;
} sdb_sweep_desc;
/* Issues n_sweeps x [row pass, finalize+update f, column pass, finalize+update g, absorb] on `stream`.
 * lr_known_first != 0: Lr already holds the row LSE at the current g, the first row pass is skipped
 * (final-stage gap check, see spadot_b200/sinkhorn.py).  Sweep i stamps the absorb flag with first_tick + i. */
int sdb_sinkhorn_sweeps(const sdb_sweep_desc* d, int n_sweeps, int first_tick, int lr_known_first, void* stream);
/* The same n_sweeps iterations in ONE cooperative launch (SIMT form only, use_tc == 0): CTAs walk the (slab, split)
 * items of each pass and meet at a grid barrier between pass and update, so a ChickenHeart-sized iteration costs four
 * barriers instead of five launches.  barrier2: SDB_BARRIER_WORDS unsigned ints owned by the caller (zeroed by every call, reusable across
 * calls).  Grid = d->n_ctas CTAs (capped by co-residency; 0 = one per work item); the caller sizes the column splits
 * for it.  Same tile and update code as sdb_sinkhorn_sweeps; the partials of a row are combined by a warp instead of a
 * thread, so iterates agree to fp64 rounding of that sum. */
int sdb_sinkhorn_sweeps_persistent(const sdb_sweep_desc* d, int n_sweeps, int first_tick, int lr_known_first,
                                   unsigned int* barrier2, void* stream);

/* The whole duality-gap solve of a small problem (SIMT form) in ONE cooperative launch: six epsilon stages, stage 0-4
 * criterion, final-stage duality gap, tau bookkeeping - everything optimal_transport_duality_gap does around its native
 * loop (ref: ot_solvers.py:240-449, ot_func.cpp:831-930).  Two grid barriers per iteration: the CTA that finishes the last
 * column split of a 64-row slab combines the slab's partials and updates its potentials.  On return (after the stream
 * synchronises) f, g, u, v, Lr (row LSE at the final g), Lc hold the final state and *result the iteration counts.
 * Caller-owned workspaces: flag2 (2 ints), barrier2 (SDB_BARRIER_WORDS uints), counters (ceil(n/64) + ceil(m/64) uints), scratch
 * (SDB_SOLVE_MAX_CTAS * 10 + n + m doubles: per-CTA sums, then the strip form's tagged bias vectors), result (device).
 * Iterations are stamped first_tick, first_tick + 1, ... */
#define SDB_SOLVE_MAX_CTAS 1024
#define SDB_STRIP_SMEM_MAX (223 * 1024)  /* strip form: 4*(ceil(n/g)*pad4(m) + ceil(m/g)*pad4(n) + max(pad4(n), pad4(m), ceil(max(n,m)/g)*dpad))
                                          * bytes with g = min(#SM, n, m) must not exceed this */
#define SDB_BARRIER_WORDS 1024   /* grid barrier of the cooperative kernels: top counter, generation, 16 group counters 128 B apart */
typedef struct sdb_solve_params {
    double lambda1, lambda2, epsilon, epsilon0, tolerance, tau, max_iter;
    double eps_stage[6];         /* the six regularisations, computed by the caller exactly as ot_solvers.py:218,240,254 does */
    double xy_max;               /* max_i |x_i| * max_j |y_j| of the prepared points */
    double dot_limit;            /* a stage uses dot-product tiles while 2*c1*log2(e)*xy_max <= dot_limit, else direct ones */
    int32_t batch_size;
    int32_t reserved;            /* 1 = RESIDENT form: every CTA keeps its 64x64 tiles of the scaled cost matrix
                                  * c1*log2(e)*|x_i - y_j|^2 in shared memory for a whole epsilon stage (a ChickenHeart-sized
                                  * problem fits on chip), so a half-iteration is bias - cost, ex2, add per pair.  The caller
                                  * then sizes partial_row / partial_col for ns_row = ceil(m/64), ns_col = ceil(n/64) splits;
                                  * SDB_E_UNSUPPORTED when the tiles do not fit (use 0 = streamed tiles).
                                  * 2 = STRIP form (owner computes): one CTA per SM holds ceil(n/grid) whole rows and
                                  * ceil(m/grid) whole columns of |x_i - y_j|^2 plus one bias vector in shared memory
                                  * (<= SDB_STRIP_SMEM_MAX: the ChickenHeart sizes, up to about 2000 x 2000); a half-iteration has no
                                  * partial results and no arrival counters - the owner reduces, updates, and the only
                                  * exchange is the updated vector, which travels as 64-bit (bias, production number, tau bit)
                                  * words that the consumers poll: inside a batch of iterations there is NO grid barrier, a
                                  * pass starts when the entries it reads carry the number it expects.  partial_* / ns_* are
                                  * not used.  SDB_E_UNSUPPORTED when the strips do not fit. */
} sdb_solve_params;
typedef struct sdb_solve_result {
    int32_t iters[6];            /* iterations per epsilon stage */
    int32_t total_iters;
    int32_t status;              /* 0 ok, 1 NaN gap (ref: ot_solvers.py:446-447) */
    int32_t max_iter_reached;    /* a stage gave up at max_iter (ref: ot_func.cpp:821-824) */
    int32_t last_tick;
    double gap, eps_final;
    double stage_gap[6];         /* criterion value that ended each stage (stage 0-4: relative change, stage 5: duality gap) */
} sdb_solve_result;
int sdb_sinkhorn_solve_persistent(const sdb_sweep_desc* d, const sdb_solve_params* p, int first_tick, int* flag2,
                                  unsigned int* barrier2, unsigned int* counters, double* scratch,
                                  sdb_solve_result* result, void* stream);

/* ------------------------------------------------------------------ row-partitioned solve: NCCL inside the ABI */
/* The library binds NCCL at run time from the libnccl.so.2 already loaded in the process (PyTorch's); nothing here is needed
 * for single-GPU use.  One communicator per process/GPU: rank 0 calls sdb_nccl_unique_id, ships the 128 bytes to the other
 * ranks by any means (the Python driver broadcasts them with torch.distributed), every rank calls sdb_nccl_comm_create. */
int sdb_nccl_available(void);
int sdb_nccl_unique_id(void* id128_host);
int sdb_nccl_comm_create(const void* id128_host, int world, int rank, void** comm_out);
int sdb_nccl_comm_destroy(void* comm);
int sdb_nccl_allreduce_sum_f64(void* comm, double* buf, int64_t count, void* stream);
/* n_sweeps iterations of the row-partitioned solve, all launches and the ONE all-reduce per iteration issued from C on
 * `stream`: row pass + update of the local f; column pass against the common shift d->m_y (see sdb_partial_sums_f64);
 * ncclAllReduce(sum) of the M + 2 doubles in `sums`; sdb_update_from_sums_f64; sdb_absorb.  Preconditions: d->m_y holds the
 * common shift for (eps, median), d->bias_y matches the current g, d->bad_flag / d->m_x as in sdb_sinkhorn_sweeps
 * (the row pass uses predictions from sweep d->pred_from_row on).  A rank with n == 0 rows contributes zeros. */
int sdb_sinkhorn_sweeps_dist(const sdb_sweep_desc* d, void* comm, double* sums, int n_sweeps, int first_tick, void* stream);

/* ------------------------------------------------------------------ K4: stopping rules */
/* Stage 0-4 rule (ref: ot_func.cpp:897-922).  out[0..3] =
 *   sum (a~ - old_a e^{u/eps})^2, sum a~^2, sum (b~ - old_b e^{v/eps})^2, sum b~^2
 * with a~ = exp((f-u)/eps)*exp(u/eps).
 * scratch (all reductions): SDB_REDUCE_WIDTH*SDB_REDUCE_BLOCKS doubles + one zero-initialised uint after them. */
#define SDB_REDUCE_BLOCKS 1024
#define SDB_REDUCE_WIDTH 10
int sdb_stage_criterion(int64_t n, int64_t m, const double* f, const double* u, const double* la_old,
                        const double* g, const double* v, const double* lb_old, double eps,
                        double* out4, void* scratch, void* stream);
/* Final-stage duality-gap ingredients (ref: ot_func.cpp:358-544, ot_solvers.py:132-158).  out[0..9]:
 *  [0] sum_i Rrow_i  [1] sum_i f_i Rrow_i  [2] sum_i dx (r log(r/p) - r + p), r = Rrow_i/m
 *  [3] sum_i p_i dx (exp(-f_i/lam1) - 1)
 *  [4] sum_j Rcol_j  [5] sum_j g_j Rcol_j  [6] sum_j dy (c log(c/q) - c + q), c = Rcol_j/n
 *  [7] sum_j q_j dy (exp(-g_j/lam2) - 1)   [8],[9] spare
 * with Rrow_i = exp(f_i/eps + Lr_i), Rcol_j = exp(g_j/eps + Lc_j); dx = 1/N_total, dy = 1/M are explicit
 * because n may be one rank's row slice.  scratch as above (10 wide). */
int sdb_gap_terms(int64_t n, int64_t m, const double* f, const double* Lr, const double* logp,
                  const double* g, const double* Lc, const double* logq,
                  double eps, double lam1, double lam2, double dx, double dy,
                  double* out10, void* scratch, void* stream);
/* out[0] = sum_i exp(L[i] + add[i]/eps_or_1)  generic fp64 sum of exp (used for sum_K and masses);
 * add may be NULL. */
int sdb_sum_exp(int64_t n, const double* L, const double* add, double add_scale, double* out1,
                void* scratch, void* stream);

/* ------------------------------------------------------------------ plan / marginals */
/* plan[i*m+j] = exp((f_i + g_j - C_ij)/eps) * inv_m with C_ij = sum_k (x_ik-y_jk)^2 * inv_med computed
 * by direct fp64 differences (ref: ot_solvers.py:102-103,449; ot_func.cpp:571-584). x,y row-major fp64. */
int sdb_plan_dense_f64(const double* x, const double* y, int64_t n, int64_t m, int d,
                       const double* f, const double* g, double inv_med, double eps, double inv_m,
                       double* plan, void* stream);
/* out[i] = exp(f_i/eps + Lr_i) * inv_m   — plan row sums = next growth vector (ref: ot_solvers.py:116). */
int sdb_row_mass(int64_t n, const double* f, const double* Lr, double eps, double inv_m, double* out,
                 void* stream);

/* ------------------------------------------------------------------ K5: exact median of all N*M costs */
/* dist[s] = sum_k (x[ii[s]*d+k] - y[jj[s]*d+k])^2   fp64, for sampled index pairs. */
int sdb_pair_distances_f64(const double* x, const double* y, int d, const int64_t* ii, const int64_t* jj,
                           int64_t n_pairs, double* dist, void* stream);
/* One streamed sweep over all pairs (fp32 direct differences on the prepared points):
 *   counts[0] += #pairs with D32 <  lo;  pairs with lo <= D32 < hi are binned uniformly into
 *   hist[0..n_bins) (counts[1] = their total).  n_bins <= 4096. */
int sdb_cost_histogram(const float* pt, int64_t ldp, int64_t n_p, const float* qt, int64_t ldq, int64_t n_q,
                       int dpad, const int64_t* split_bounds, int n_splits, float lo, float hi, int n_bins,
                       unsigned long long* hist, unsigned long long* counts2, void* stream);
/* Final sweep: counts[0] += #pairs with D32 < lo; every pair with lo <= D32 < hi has its cost
 * recomputed by direct fp64 differences from x,y (row-major fp64, ORIGINAL coordinates) and is
 * appended to cand (capacity cap; counts[1] = number appended, may exceed cap => overflow). */
int sdb_cost_collect(const float* pt, int64_t ldp, int64_t n_p, const float* qt, int64_t ldq, int64_t n_q,
                     int dpad, const int64_t* split_bounds, int n_splits, float lo, float hi,
                     const double* x, const double* y, int d,
                     double* cand, unsigned long long cap, unsigned long long* counts2, void* stream);
/* Tensor-core forms of the two sweeps above for the fp16 hi/lo split points of sdb_prep_points_split_f16:
 * dist = row_norms[i] + col_norms_padded[j] - 2 x_i.y_j with fp32 norms (col_norms_padded[j] >= 3e38 for j >= n_q),
 * scale = -2 * 2^(-2*pow2_exp); same persistent tcgen05/TMEM/TMA pipeline as sdb_lse_pass_tc, n_bins <= 4096. */
int sdb_cost_histogram_tc(const void* p16, int64_t n_p, int64_t n_p_pad, const void* q16, int64_t n_q, int64_t n_q_pad,
                          int dp, const float* row_norms, const float* col_norms_padded, float scale,
                          int tiles_per_split, int n_ctas, float lo, float hi, int n_bins,
                          unsigned long long* hist, unsigned long long* counts2, void* stream);
int sdb_cost_collect_tc(const void* p16, int64_t n_p, int64_t n_p_pad, const void* q16, int64_t n_q, int64_t n_q_pad,
                        int dp, const float* row_norms, const float* col_norms_padded, float scale,
                        int tiles_per_split, int n_ctas, float lo, float hi, const double* x, const double* y, int d,
                        double* cand, unsigned long long cap, unsigned long long* counts2, void* stream);
/* hist256[b] = #{ i : (key(cand[i]) >> shift) & 255 == b and key(cand[i]) >> (shift+8) == prefix }
 * (key = IEEE bits of a non-negative double); one radix-select digit. hist256 must be zeroed. */
int sdb_radix_digit_hist(const double* cand, unsigned long long n, int shift, unsigned long long prefix,
                         unsigned long long* hist256, void* stream);
/* out[s] (device) = the ranks[s]-th smallest (0-based; ranks is a HOST array of n_ranks = 1 or 2 values) of the n non-negative
 * doubles cand[] - the two middle order statistics np.median averages (ref: ot_solvers.py:103).  Eight radix-256 digit passes
 * issued back to back on `stream`, no host round trip in between.  workspace: SDB_SELECT_WORKSPACE_BYTES of device memory. */
#define SDB_SELECT_WORKSPACE_BYTES 4128
int sdb_select_ranks_f64(const double* cand, unsigned long long n, const unsigned long long* ranks, int n_ranks, double* out,
                         void* workspace, void* stream);

/* ------------------------------------------------------------------ K6: domain transition table */
/* table[a*k1+b] += sum_{i: label_row[i]=a} exp(f_i/eps + ln2*(M + log2 S) - norms_i*c1) * inv_m
 * where (M,S) = partial of column segment b (columns sorted by label; one split per label).
 * replaces wot transition_table  P0^T . T . P1  (ref: utils/_analyze_utils.py:135-137). */
int sdb_transition_accumulate(const float* partial, int k1, int64_t n, const double* norms, double c1,
                              const double* f, double eps, double inv_m, const int* label_row, int k0,
                              double* table, void* stream);

/* ------------------------------------------------------------------ dense-cost path (caller supplies C) */
/* For the reference signatures that take an explicit cost matrix (ref: utils/OT_loss/ot_solvers.py:164,452).
 * C is row-major fp64 with leading dimension ldc; one sweep reads C once.  fp64 throughout.
 * L[i] = LSE_j[(g_j - C_ij)/eps]  (g may be NULL = 0). */
int sdb_dense_row_lse_f64(const double* C, int64_t ldc, int64_t n, int64_t m, const double* g, double eps,
                          double* L, void* stream);
/* L[j] = LSE_i[(f_i - C_ij)/eps]; partial: workspace of 2*n_chunks*m doubles (rows are swept in n_chunks chunks). */
int sdb_dense_col_lse_f64(const double* C, int64_t ldc, int64_t n, int64_t m, const double* f, double eps,
                          double* L, double* partial, int n_chunks, void* stream);
/* plan[i*m+j] = exp((f_i + g_j - C_ij)/eps) * inv_m   (R / J, ref: ot_solvers.py:449); n <= 65535 per call. */
int sdb_dense_plan_f64(const double* C, int64_t ldc, int64_t n, int64_t m, const double* f, const double* g,
                       double eps, double inv_m, double* plan, void* stream);

/* ------------------------------------------------------------------ K1: SVGP kernel blocks (train inner loop) */
/* out[i*ldo+j] = k(|x_i - y_j|^2) (+ jitter when i == j), x (n,dim), y (m,dim) row-major fp64, dim <= 3.
 * kernel_type 0 Gaussian exp(-d2/scale), 1 Cauchy 1/(1+d2/scale), 2 Quadratic 1 - d2/(d2+scale)
 * replaces Kernel.forward + _add_diagonal_jitter (ref: model/svgp.py:116-125, 33-41). */
int sdb_kernel_block_f64(const double* x, int64_t n, const double* y, int64_t m, int dim, int kernel_type,
                         double scale, double jitter, double* out, int64_t ldo, void* stream);
/* out[i] = k(|x_i - y_i|^2): kernel_matrix(..., diag_only=True) without the n x n block (ref: svgp.py:43-45). */
int sdb_kernel_diag_f64(const double* x, const double* y, int64_t n, int dim, int kernel_type, double scale,
                        double* out, void* stream);
/* out[i] = w_i^T A w_i for the rows w_i of W (b,m), A (m,m) row-major: the O(b m^2) form of the (b,m,m)
 * batched product in _compute_l3_term (ref: svgp.py:96-104). */
int sdb_quad_form_rows_f64(const double* W, const double* A, int64_t b, int m, double* out, void* stream);

/* ------------------------------------------------------------------ K2 / K2b: GAT edge-softmax + aggregation */
/* Graph in CSR by destination: rowptr (n+1, int64), col (E, int32) = source node of each incoming edge
 * (self loops included, PyG GATConv add_self_loops semantics are the caller's job).  feat (n,H,C),
 * a_src / a_dst (n,H) in float or double (is_double).  Writes out (n,H,C) = sum_j alpha_ij feat_j and
 * alpha (E,H) = softmax_j(leaky_relu(a_src[j]+a_dst[i])) with PyG's +1e-16 denominator.
 * node_order (n, int32, may be NULL): the order in which nodes are dealt to CTAs.  With a locality-preserving order (Z-curve of
 * the spot coordinates, reverse Cuthill-McKee) the kernels run in TILE form: eight consecutive nodes of the order per CTA, the
 * rows they gather de-duplicated in shared memory (neighbouring spots share most neighbours), every distinct row streamed once
 * per tile; results are reproducible and independent of the order up to rounding.  Without an order (small graphs) one CTA
 * per node runs.  Tiles that exceed the shared-memory budget (512 edges / 256 distinct rows) and layers with more than four
 * heads use the per-node form.  SDB_GAT_TILES=0/1 in the environment forces either form.
 * replaces torch_geometric GATConv message passing used at ref: model/encoder.py:41-45,56-58. */
int sdb_gat_forward(const void* feat, const void* a_src, const void* a_dst, const int64_t* rowptr, const int32_t* col,
                    const int32_t* node_order, int64_t n, int H, int C, double negative_slope, int is_double,
                    void* out, void* alpha, void* stream);
/* Backward of the above.  src_rowptr / src_dst / src_eid: the same edges grouped by SOURCE node
 * (destination node and position in the by-destination order).  dlogit (E,H) is workspace.
 * Outputs grad_feat (n,H,C), grad_a_src (n,H), grad_a_dst (n,H).  Deterministic (no global atomics; in the tile form
 * parallel edges - the same (source, target) pair listed twice - are merged by a shared-memory add, the only order-dependent sum). */
int sdb_gat_backward(const void* feat, const void* a_src, const void* a_dst, const int64_t* rowptr, const int32_t* col,
                     const int64_t* src_rowptr, const int32_t* src_dst, const int32_t* src_eid, const int32_t* node_order,
                     int64_t n, int H, int C, double negative_slope, int is_double, const void* alpha,
                     const void* grad_out, void* dlogit, void* grad_feat, void* grad_a_src, void* grad_a_dst,
                     void* stream);
/* The same backward for a layer whose destinations are the first n_dst nodes of a graph with n_src >= n_dst source nodes
 * (hop-ordered mini-batches: seeds first, then the nodes each hop adds; ref: utils/_train_utils.py:80-85 NeighborLoader order,
 * model/SpaDOT.py:83-84 only rows [:batch_size] of the encoder output are used).  rowptr / col are the by-destination CSR of the
 * WHOLE batch graph: its first rowptr[n_dst] edges are exactly the layer's edges, so nothing is rebuilt; the by-source lists
 * (also of the whole graph) are filtered on the fly with src_eid < rowptr[n_dst].  order_dst permutes [0,n_dst), order_src
 * [0,n_src) (either may be NULL).  feat/a_src/grad_feat/grad_a_src have n_src rows, a_dst/grad_out/grad_a_dst n_dst rows.
 * sdb_gat_forward serves this form unchanged with n = n_dst. */
int sdb_gat_backward_prefix(const void* feat, const void* a_src, const void* a_dst, const int64_t* rowptr, const int32_t* col,
                            const int64_t* src_rowptr, const int32_t* src_dst, const int32_t* src_eid, const int32_t* order_dst,
                            const int32_t* order_src, int64_t n_dst, int64_t n_src, int H, int C, double negative_slope,
                            int is_double, const void* alpha, const void* grad_out, void* dlogit, void* grad_feat,
                            void* grad_a_src, void* grad_a_dst, void* stream);
/* The same message passing with SHARED features: feat is (n_src, C) and every head weighs the same row,
 *     out[i,h,:] = sum_j alpha^h_ij feat[j,:]            (n, H, C),
 * which is what a GAT layer needs when the linear map is applied AFTER the aggregation (out_h = W_h sum_j alpha^h_ij x_j equals
 * sum_j alpha^h_ij W_h x_j; spadot_b200/gat.py does this for prefix layers, where it leaves the GEMM n_dst instead of n_src rows;
 * ref: model/encoder.py:41-45,56-58 states the layers, torch_geometric's GATConv always transforms first).  a_src / a_dst are
 * the caller's attention scalars (n_src,H) / (n,H).  Backward: grad_out (n_dst,H,C) -> grad_feat (n_src, C), the sum over the
 * heads done inside the kernel; everything else as sdb_gat_forward / sdb_gat_backward_prefix. */
int sdb_gat_forward_shared(const void* feat, const void* a_src, const void* a_dst, const int64_t* rowptr, const int32_t* col,
                           const int32_t* node_order, int64_t n, int H, int C, double negative_slope, int is_double,
                           void* out, void* alpha, void* stream);
int sdb_gat_backward_shared(const void* feat, const void* a_src, const void* a_dst, const int64_t* rowptr, const int32_t* col,
                            const int64_t* src_rowptr, const int32_t* src_dst, const int32_t* src_eid, const int32_t* order_dst,
                            const int32_t* order_src, int64_t n_dst, int64_t n_src, int H, int C, double negative_slope,
                            int is_double, const void* alpha, const void* grad_out, void* dlogit, void* grad_feat,
                            void* grad_a_src, void* grad_a_dst, void* stream);

/* k nearest OTHER points (self excluded) of every point, sorted by (distance, index); pts (n,dim) fp64, dim <= 3,
 * k <= 32 and k <= n-1.  out_idx (n,k) int32, out_dist (n,k) fp64 Euclidean distances (may be NULL).
 * replaces sklearn NearestNeighbors in _Cal_Spatial_Net (ref: utils/_utils.py:65-69). */
int sdb_knn_f64(const double* pts, int64_t n, int dim, int k, int32_t* out_idx, double* out_dist, void* stream);
/* The same result for 2-D points in O(n k): pts_sorted (n,2) = the points sorted by the cell of a uniform gx x gy grid of
 * pitch h with origin (x0, y0), cell = min(gy-1, floor((y-y0)/h))*gx + min(gx-1, floor((x-x0)/h)); order[s] = original index of
 * sorted slot s, cell_of_sorted[s] its cell, cell_start (gx*gy+1) the CSR offsets of the cells.  out_idx / out_dist are
 * indexed by ORIGINAL point index.  A query scans rings of cells until its k-th distance is below the distance to the
 * visited block's border, so neighbours and their (distance, index) order are exactly those of sdb_knn_f64. */
int sdb_knn_grid_f64(const double* pts_sorted, const int32_t* order, const int32_t* cell_of_sorted, const int32_t* cell_start,
                     int64_t n, int gx, int gy, double x0, double y0, double h, int k, int32_t* out_idx, double* out_dist,
                     void* stream);

/* ------------------------------------------------------------------ K7: k-means Lloyd iterations */
/* One Lloyd E-step (+ accumulation for the M-step): labels[i] <- argmin_j |c_j|^2 - 2 x_i.c_j (first minimum on ties),
 * *changed set to 1 when any label moved; with sums (k,d) / counts (k) non-NULL (zeroed by the caller) the per-cluster
 * coordinate sums and sizes are accumulated.  X (n,d), centers (k,d) row-major fp64; k <= 64, k*d <= 4096.
 * replaces sklearn.cluster.KMeans' lloyd_iter behind ref: utils/_train_utils.py:255-269, utils/_analyze_utils.py:10-39. */
int sdb_kmeans_assign(const double* X, const double* centers, int64_t n, int d, int k, int32_t* labels, int* changed,
                      double* sums, double* counts, void* stream);
/* centers_new = sums / counts (empty clusters keep their centre); *shift_sq = sum |centers_new - centers_old|^2. */
int sdb_kmeans_update(const double* sums, const double* counts, const double* centers_old, double* centers_new, int d,
                      int k, double* shift_sq, void* stream);
/* out1[0] = sum_i |x_i - c_{labels[i]}|^2 (direct differences); scratch as for the other reductions. */
int sdb_kmeans_inertia(const double* X, const double* centers, const int32_t* labels, int64_t n, int d, double* out1,
                       void* scratch, void* stream);
/* n_runs complete Lloyd runs in ONE launch, one CTA per run (the n_init restarts of a fit; ref: sklearn's
 * _kmeans_single_lloyd behind utils/_train_utils.py:262-266).  centers_init: (n_runs, k, d).  Outputs per run: labels (n),
 * centres (k, d), inertia, iterations, status (0 = done, 1 = met an empty cluster: redo that run with the host-driven
 * kernels above, which implement sklearn's relocation).  Meant for n*d up to about half a million (one SM streams X from L2). */
int sdb_kmeans_lloyd_runs(const double* X, int64_t n, int d, int k, const double* centers_init, int n_runs, int max_iter,
                          double tol, int32_t* labels_out, double* centers_out, double* inertia_out, int32_t* n_iter_out,
                          int32_t* status_out, void* stream);

/* ------------------------------------------------------------------ pipe micro-benchmarks (roofline denominators) */
/* One launch of n_ctas x 512 threads running `iters` rounds of independent chains of one instruction kind:
 * kind 0 = MUFU.EX2 (ex2.approx.ftz.f32), 1 = FFMA, 2 = FFMA2 (fma.rn.f32x2), 3 = the epilogue's mix per two pairs
 * (FFMA2, FADD2, 2 x MUFU.EX2, FADD2; counted in ex2).  *ops_out_host (HOST pointer, may be NULL) receives the number of
 * operations of that kind the launch executes (ex2 evaluations, or FMAs = 2 flop each); the caller times the launch.
 * `out` needs n_ctas*512 floats and is never written.  SURVEY.md 8(d): "to be confirmed by a micro-benchmark on the box". */
int sdb_pipe_peak(int kind, int n_ctas, int iters, float* out, double* ops_out_host, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SPADOT_B200_H */
