/* TEST INFRASTRUCTURE ONLY — multi-threaded (pthreads) C port of the log-domain half-iteration.
 *
 * Restates, for sizes whose dense fp64 K does not fit in host memory, what the reference's
 * gemv/gemtv + update_k compute (SpaDOT/utils/OT_loss/ot_func.cpp:43-249,547-568):
 *     L_i = log sum_j exp( col_term_j - C_ij / eps ),  C_ij = sum_k (x_ik - y_jk)^2 * inv_med
 * (direct differences, like scipy cdist behind ot_solvers.py:102).  fp64 throughout.
 * Used by tests (checked against oracle/ot_logdomain.py) and as bench.py's "port" CPU baseline.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>

typedef struct {
    const double *x, *y, *col_term;
    int64_t i0, i1, m;
    int d;
    double inv_med, eps;
    double* out;
} lse_job;

static void* lse_worker(void* arg) {
    lse_job* J = (lse_job*)arg;
    const double scale = J->inv_med / J->eps;
    for (int64_t i = J->i0; i < J->i1; ++i) {
        const double* xi = J->x + i * J->d;
        double mx = -INFINITY, acc = 0.0;
        for (int64_t j = 0; j < J->m; ++j) {
            const double* yj = J->y + j * J->d;
            double s = 0.0;
            for (int k = 0; k < J->d; ++k) { const double df = xi[k] - yj[k]; s += df * df; }
            const double t = J->col_term[j] - s * scale;
            if (t > mx) { acc = acc * exp(mx - t) + 1.0; mx = t; }
            else if (t > -INFINITY) acc += exp(t - mx);
        }
        J->out[i] = (mx == -INFINITY) ? -INFINITY : mx + log(acc);
    }
    return NULL;
}

void oracle_row_lse(const double* x, const double* y, int64_t n, int64_t m, int d, const double* col_term,
                    double inv_med, double eps, double* out, int n_threads) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    pthread_t th[256];
    lse_job jobs[256];
    const int64_t chunk = (n + n_threads - 1) / n_threads;
    int used = 0;
    for (int t = 0; t < n_threads; ++t) {
        const int64_t i0 = t * chunk, i1 = (i0 + chunk < n) ? i0 + chunk : n;
        if (i0 >= i1) break;
        jobs[t] = (lse_job){x, y, col_term, i0, i1, m, d, inv_med, eps, out};
        pthread_create(&th[t], NULL, lse_worker, &jobs[t]);
        ++used;
    }
    for (int t = 0; t < used; ++t) pthread_join(th[t], NULL);
}

/* one full Sinkhorn iteration in total potentials; f,g updated in place */
void oracle_sinkhorn_iteration(const double* x, const double* y, int64_t n, int64_t m, int d, const double* logp,
                               const double* logq, double inv_med, double eps, double alpha1, double alpha2, double* f,
                               double* g, int n_threads) {
    const int64_t mx = n > m ? n : m;
    double* term = (double*)malloc(sizeof(double) * (size_t)mx);
    double* work = (double*)malloc(sizeof(double) * (size_t)mx);
    for (int64_t j = 0; j < m; ++j) term[j] = g[j] / eps;
    oracle_row_lse(x, y, n, m, d, term, inv_med, eps, work, n_threads);
    for (int64_t i = 0; i < n; ++i) f[i] = eps * alpha1 * (logp[i] - work[i] + log((double)m));
    for (int64_t i = 0; i < n; ++i) term[i] = f[i] / eps;
    oracle_row_lse(y, x, m, n, d, term, inv_med, eps, work, n_threads);
    for (int64_t j = 0; j < m; ++j) g[j] = eps * alpha2 * (logq[j] - work[j] + log((double)n));
    free(term);
    free(work);
}
