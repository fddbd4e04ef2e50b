/*
 * libot_b200.h — the reference's own native ABI (libot.so), served by B200 kernels.
 *
 * SpaDOT binds its OT inner loop through ctypes to `libot.so`
 * (ref: SpaDOT/utils/OT_loss/ot_func.py:8-10, built from ot_func.cpp).  `libot_b200.so` exports the
 * SAME fourteen symbols with the SAME signatures and the same calling convention:
 *
 *   - every array is a caller-owned, C-contiguous, row-major HOST buffer (numpy), passed as a raw
 *     pointer; matrices are (m, n) with `int m, int n` trailing; scalars by value;
 *   - buffers documented "in/out" are mutated in place exactly where the reference mutates them;
 *   - no handles, no contexts, no error codes besides the reference's own (-1 from step1_process,
 *     NaN gap); calls are serialised by one process-wide mutex; between calls the library keeps
 *     three counters and a pool of the device blocks the last calls used (cudaMalloc/cudaFree of
 *     matrix-sized blocks cost more than the PCIe copies; libot_b200_release() or the environment
 *     variable LIBOT_B200_POOL=0 returns to allocate-per-call) — never any problem data;
 *   - there is NO CPU fallback: without a usable sm_100 device every entry point prints a message
 *     to stderr and aborts the process (the ABI has no error channel to report it through).
 *
 * Each call uploads its operands, runs dense fp64 (or fp32 for *_float) kernels on the device and
 * downloads what the reference would have mutated.  Arithmetic follows the reference expression by
 * expression (same divisions, pow/exp/log placement, nan_to_num clamp to +-FLT_MAX); only the
 * order of the long sums differs (tree reductions), i.e. results agree to fp64 rounding, not bit
 * for bit.  Because the contract keeps K, _K, C, R on the host, every call is PCIe-bound — the
 * streamed API of spadot_b200.h is the fast path; this library exists so that the reference's
 * ot_func.py / ot_solvers.py run unmodified on the GPU by changing one path.
 *
 * Differences from the reference, all outside its defined behaviour: element counts are computed
 * in 64 bits (the reference indexes `i*n+j` in `int`, ot_func.cpp:413,565, and overflows beyond
 * 2^31 elements); m == 0 or n == 0 is a no-op.
 */
#ifndef LIBOT_B200_H
#define LIBOT_B200_H

#ifdef __cplusplus
extern "C" {
#endif

/* ref: ot_func.cpp:938-975 — placeholders used by the reference to time the ctypes call overhead; return 0. */
float  dummy_float (float*  C, float*  K, float*  R, float*  dx, float*  dy, float*  p, float*  q, float*  a, float*  b,
                    float  epsilon, float  lambda1, float  lambda2, int m, int n);
double dummy_double(double* C, double* K, double* R, double* dx, double* dy, double* p, double* q, double* a, double* b,
                    double epsilon, double lambda1, double lambda2, int m, int n);

/* ref: ot_func.cpp:977-1043 (primal<T> :358-462).  All inputs read-only.
 * F1(R dy, p) + F2(R^T dx, q) + (eps * sum(R*nan_to_num(log R) - R + K) + sum(R*C)) / (m*n),
 * F(x, p) = lambda * sum_k d_k (x_k log(x_k/p_k) - x_k + p_k).  `a`, `b` are unused (as in the reference). */
float  primal_float (float*  C, float*  K, float*  R, float*  dx, float*  dy, float*  p, float*  q, float*  a, float*  b,
                     float  epsilon, float  lambda1, float  lambda2, int m, int n);
double primal_double(double* C, double* K, double* R, double* dx, double* dy, double* p, double* q, double* a, double* b,
                     double epsilon, double lambda1, double lambda2, int m, int n);

/* ref: ot_func.cpp:1045-1111 (dual<T> :465-490).
 * -F1*(a) - F2*(b) - eps * sum(R - K) / (m*n),  F*(x) = lambda * sum_k p_k d_k (exp(-eps log(x_k)/lambda) - 1). */
float  dual_float (float*  C, float*  K, float*  R, float*  dx, float*  dy, float*  p, float*  q, float*  a, float*  b,
                   float  epsilon, float  lambda1, float  lambda2, int m, int n);
double dual_double(double* C, double* K, double* R, double* dx, double* dy, double* p, double* q, double* a, double* b,
                   double epsilon, double lambda1, double lambda2, int m, int n);

/* ref: ot_func.cpp:1113-1179 (compute_duality_gap<T> :493-544):  (primal - dual) / |primal|. */
float  compute_duality_gap_float (float*  C, float*  K, float*  R, float*  dx, float*  dy, float*  p, float*  q, float*  a,
                                  float*  b, float  epsilon, float  lambda1, float  lambda2, int m, int n);
double compute_duality_gap_double(double* C, double* K, double* R, double* dx, double* dy, double* p, double* q, double* a,
                                  double* b, double epsilon, double lambda1, double lambda2, int m, int n);

/* ref: ot_func.cpp:1181-1223 (update_k<T> :547-568).  out: K_ = exp(-C/eps), K = exp((u_i + v_j - C_ij)/eps). */
void update_k_float (float*  K, float*  K_, float*  C, float*  u, float*  v, float  epsilon, int m, int n);
void update_k_double(double* K, double* K_, double* C, double* u, double* v, double epsilon, int m, int n);

/* ref: ot_func.cpp:1225-1259 (update_R<T> :571-584).  out: R_ij = K_ij * a_i * b_j. */
void update_R_float (float*  R, float*  K, float*  a, float*  b, int m, int n);
void update_R_double(double* R, double* K, double* a, double* b, int m, int n);

/* ref: ot_func.cpp:1261-1311 (step1_process<T> :690-828; update_a_b<T> :587-687, gemv :43-170, gemtv :173-249).
 * `iters` Sinkhorn iterations.  Each: old_a <- a, old_b <- b;
 *   a <- (p / (K (b*dy)))^alpha1 * exp(-u/(lambda1+eps));  b <- (q / (K^T (a*dx)))^alpha2 * exp(-v/(lambda2+eps));
 *   if any a_i > tau or b_j > tau:  u += eps log a, v += eps log b, K <- exp((u_i+v_j-C_ij)/eps), a <- 1, b <- 1;
 *   ++cur_iter; if cur_iter >= max_iter: print the reference's message to stdout and return -1.
 * in/out: a, b, old_a, old_b, K, u, v.  Returns the new iteration count. */
int step1_process_double(double* a, double* b, double* old_a, double* old_b, double* K, double* C, double* dx, double* dy,
                         double* p, double* q, double* u, double* v, int cur_iter, int max_iter, int iters, double tau,
                         double lambda1, double lambda2, double alpha1, double alpha2, double epsilon, int m, int n);

/* ref: ot_func.cpp:1313-1373 (update_process<T> :831-930).  One epsilon stage of optimal_transport_duality_gap
 * (ref: ot_solvers.py:283-290): repeat { step1_process(iters = batch_size in the last stage, else 5); then either the
 * duality gap of R = a K b against _K (last stage; R is written) or the dual-evolution criterion
 * max(|_a - old_a e^{u/eps}| / (1+|_a|), same for b) } until the value is <= threshold.  The 28-parameter form the
 * reference passes through ctypes (ot_func.py:561-567).  in/out: R (last stage only), a, b, old_a, old_b, K, u, v.
 * Returns the final gap / criterion. */
double update_process_double(double* R, double* a, double* b, double* old_a, double* old_b, double* K, double* _K, double* C,
                             double* dx, double* dy, double* p, double* q, double* u, double* v, int epsilon_scalings,
                             int cur_epsilon_scaling, int batch_size, double epsilon, double threshold, double tau,
                             double lambda1, double lambda2, double alpha1, double alpha2, int cur_iter, int max_iter,
                             int m, int n);

/* Not part of the reference ABI: library/device identification for tests (0 = a usable sm_100 device is present). */
int libot_b200_device_check(void);
int libot_b200_version(void);
/* Frees every pooled device block. */
void libot_b200_release(void);
/* Cumulative counters since load: kernels launched, bytes uploaded, bytes downloaded (for the bench's accounting). */
void libot_b200_counters(long long* launches, long long* h2d_bytes, long long* d2h_bytes);

#ifdef __cplusplus
}
#endif
#endif
