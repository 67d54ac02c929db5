/*
 * finmath_b200.h — C ABI of the B200-native Monte-Carlo path-simulation backend for finmath-lib.
 *
 * This is the drop-in boundary: a JNI shim (finmath-lib_b200/java/), C++ and Python ctypes all
 * bind exactly these symbols.  Plain pointers, sizes and opaque 64-bit handles only — no torch / CUDA types.
 * Every function returns 0 on success and a non-zero FMB_E* code on failure; fmb_last_error() gives the thread-local
 * message.  There is NO CPU fallback: every entry point that computes fails with FMB_ENODEVICE when no sm_100 device
 * can be initialised.
 *
 * Citations: J/ = /root/reference/src/main/java/net/finmath/ (finmath-lib 6.1.3-SNAPSHOT).  The reference has no FFI;
 * each group names the Java interface whose methods the JNI shim routes to it.
 *
 * Concurrency: all entry points are thread-safe (the reference calls RandomVariable ops from one pool thread per
 * process component, J/montecarlo/process/EulerSchemeFromProcessModel.java:199,:232-269, and frees from GC threads).
 * Work is stream-ordered on one compute stream per process; functions returning host values synchronise.
 */
#ifndef FINMATH_B200_H
#define FINMATH_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef uint64_t fmb_handle;          /* device-resident vector of doubles (one RandomVariable's realizations); 0 = none */

enum {
	FMB_OK = 0,
	FMB_EINVAL = 1,        /* bad argument (IllegalArgumentException on the Java side) */
	FMB_ENODEVICE = 2,     /* no usable sm_100 GPU / CUDA failure at init — never falls back to the CPU */
	FMB_ENOMEM = 3,        /* device allocation failed (OutOfMemoryError on the Java side) */
	FMB_ECUDA = 4,         /* CUDA runtime error, message in fmb_last_error() */
	FMB_EHANDLE = 5,       /* unknown / freed handle */
	FMB_EUNSUPPORTED = 6   /* UnsupportedOperationException */
};

/* ---- context ------------------------------------------------------------------------------------------------- */
int fmb_init(int device);                       /* idempotent; device = CUDA ordinal (LOCAL_RANK under torchrun) */
int fmb_shutdown(void);
int fmb_is_initialized(void);
const char* fmb_last_error(void);               /* thread-local, never NULL */
int fmb_device_count(int* count);
int fmb_device_name(char* buf, int len);
int fmb_synchronize(void);
/* 0 = STRICT (default): reference operation order, no FMA contraction -> as close to the JVM's arithmetic as the
 * device's exp/log allow.  1 = FAST (fused LMM / Black-Scholes kernels): FMA contraction, functional schemes carry the
 * log-state instead of log(exp(y)) per step (a capped rate carries log(cap)); still within 1e-12 of the reference. */
int fmb_set_fp_mode(int mode);
int fmb_get_fp_mode(int* mode);
/* device event timing on the library's compute stream (used by bench.py; torch.cuda.Event cannot see this stream) */
int fmb_timer_start(void);
int fmb_timer_stop_ms(float* ms);
int fmb_kernel_launch_count(uint64_t* count);   /* number of kernels this library has launched so far */

/* ---- memory: RandomVariableFactory.createRandomVariable(time, double[]) J/montecarlo/RandomVariableFactory.java:30-95,
 *      RandomVariable.getRealizations()/get(i)/size() J/stochastic/RandomVariable.java:60-110 ------------------------- */
int fmb_rv_create(uint64_t n, fmb_handle* out);                          /* uninitialised */
int fmb_rv_upload(const double* host, uint64_t n, fmb_handle* out);       /* copies (the Java ctor does not, its array is) */
int fmb_rv_fill(double value, uint64_t n, fmb_handle* out);
int fmb_rv_download(fmb_handle h, double* host, uint64_t n);              /* getRealizations(): a copy */
int fmb_rv_get(fmb_handle h, uint64_t i, double* out);                    /* get(i) */
int fmb_rv_size(fmb_handle h, uint64_t* n);
int fmb_rv_retain(fmb_handle h);                                          /* +1 reference (aliased process values) */
int fmb_rv_free(fmb_handle h);                                            /* -1 reference; memory returns to the pool */
int fmb_rv_free_many(const fmb_handle* handles, uint64_t count);            /* fmb_rv_free for each (0 entries skipped): one call when a process with thousands of values is collected */
int fmb_rv_device_ptr(fmb_handle h, void** dptr);                         /* raw device pointer (interop / tests) */
int fmb_pool_stats(uint64_t* bytes_in_use, uint64_t* bytes_cached, uint64_t* live_handles);
int fmb_pool_trim(void);                                                  /* release cached blocks to the driver */

/* ---- element-wise RandomVariable arithmetic, semantics of J/montecarlo/RandomVariableFromDoubleArray.java:742-1504.
 *      An operand is a handle, or — when the handle is 0 — the scalar next to it broadcast to every path (the
 *      deterministic branches of the reference).  Result is always a new handle. ----------------------------------- */
enum {  /* fmb_rv_unary: out = f(x, a) */
	FMB_U_SQUARED = 0, FMB_U_SQRT = 1, FMB_U_EXP = 2, FMB_U_LOG = 3, FMB_U_SIN = 4, FMB_U_COS = 5, FMB_U_INVERT = 6,
	FMB_U_ABS = 7, FMB_U_ISNAN = 8, FMB_U_EXPM1 = 9,
	FMB_U_ADD = 10,   /* x + a   :775 */
	FMB_U_SUB = 11,   /* x - a   :790 */
	FMB_U_BUS = 12,   /* a - x   :805 */
	FMB_U_MULT = 13,  /* x * a   :820 */
	FMB_U_DIV = 14,   /* x / a   :835 */
	FMB_U_VID = 15,   /* a / x   :850 */
	FMB_U_CAP = 16,   /* Math.min(x, a) — NaN-propagating, -0.0 < +0.0  :745 */
	FMB_U_FLOOR = 17, /* Math.max(x, a)  :760 */
	FMB_U_POW = 18,   /* Math.pow(x, a); a == 0.5 is sqrt, a == 2.0 is x*x bit-for-bit (T/montecarlo/RandomVariableTest.java:101-127) */
	FMB_U_ICDF_NORMAL = 19   /* NormalDistribution.inverseCumulativeDistribution(x), AS241 (J/functions/NormalDistribution.java:47-162): the
	                            transform BrownianMotionFromRandomNumberGenerator / IndependentIncrementsFromICDF apply to uniforms */
};
enum {  /* fmb_rv_binary: out = f(x, y) */
	FMB_B_ADD = 0, FMB_B_SUB = 1, FMB_B_MULT = 2, FMB_B_DIV = 3, FMB_B_CAP = 4, FMB_B_FLOOR = 5
};
enum {  /* fmb_rv_ternary: out = f(x, y, z, a) */
	FMB_T_ADD_PRODUCT = 0,    /* x + y * z            :1385-1427 */
	FMB_T_ADD_PRODUCT_D = 1,  /* x + y * a            :1365-1391 */
	FMB_T_ADD_RATIO = 2,      /* x + y / z            :1440-1458 */
	FMB_T_SUB_RATIO = 3,      /* x - y / z            :1461-1479 */
	FMB_T_ACCRUE = 4,         /* x * (1 + y * a)      :1278-1303 */
	FMB_T_DISCOUNT = 5,       /* x / (1.0 + y * a)    :1306-1331 */
	FMB_T_CHOOSE = 6          /* x >= 0.0 ? y : z     :1341-1363 */
};
int fmb_rv_unary(int op, fmb_handle x, double a, fmb_handle* out);
int fmb_rv_binary(int op, fmb_handle x, double sx, fmb_handle y, double sy, fmb_handle* out);
int fmb_rv_ternary(int op, fmb_handle x, double sx, fmb_handle y, double sy, fmb_handle z, double sz, double a, fmb_handle* out);

/* ((1 + r_0 d_0)(1 + r_1 d_1) ... (1 + r_{n-1} d_{n-1}) - 1) / divisor in ONE pass: the multi-period forward rate of
 * LIBORMarketModelFromCovarianceModel.getForwardRate (J/montecarlo/interestrate/models/LIBORMarketModelFromCovarianceModel.java:1288-1302),
 * which the reference builds with one accrue() pass per period.  Same operations and order (r_0.mult(d_0).add(1.0), accrue per further
 * period, sub(1.0).div(divisor)): bit-identical to the op-by-op evaluation. */
int fmb_rv_accrue_chain(int n, const fmb_handle* rates, const double* period_lengths, double divisor, fmb_handle* out);
/* Every prefix of an accrual chain in ONE pass: out[k] = start * (1 + r_0 d_0) * ... * (1 + r_k d_k), k < n - the spot-measure numeraire at
 * every tenor date (LIBORMarketModelFromCovarianceModel.java:1050-1069 builds it with one accrue() pass per date).  Same operations and
 * order as the chain of accrue() calls: bit-identical. */
int fmb_rv_accrue_prefix(int n, double start, const fmb_handle* rates, const double* period_lengths, fmb_handle* out);

/* A chain of element-wise operations evaluated in ONE pass over the vectors (deferred evaluation in the host binding: a result that
 * is only consumed by the next operation never becomes a vector in HBM).  acc = leaves[start_leaf][i]; instruction k replaces acc by
 * its operation with acc at operand position `pos`; the other operands are leaf vectors or broadcast scalars.  Same device functions,
 * order and roundings as fmb_rv_unary/binary/ternary, so the result is bit-identical to issuing the operations one by one
 * (RandomVariableFromDoubleArray.java:742-1504 applied repeatedly).
 * code: 8 bytes per instruction {kind (0 unary, 1 binary, 2 ternary), opcode, pos, refA, refB, refC, 0, 0}; a ref with bit 7 set is
 * scalars[ref & 127], otherwise leaves[ref].  unary: refA = the op's double argument (scalar).  binary: refA = the other operand.
 * ternary: refA, refB = the other two operands in x, y, z order, refC = the op's double argument (scalar). */
#define FMB_CHAIN_MAX_INSTR 16
#define FMB_CHAIN_MAX_LEAVES 8
#define FMB_CHAIN_MAX_SCALARS 24
int fmb_rv_eval_chain(int n_instr, const unsigned char* code, int start_leaf, const fmb_handle* leaves, int n_leaves,
                      const double* scalars, int n_scalars, fmb_handle* out);

/* ---- reductions (getAverage/getVariance/getMin/getMax ... :262-428).  Sums are accumulated in double-double
 *      (block + warp tree, the last CTA merges the per-CTA partials and writes the result straight into mapped host memory: one
 *      launch, 16 bytes back), returned as out2[0] = hi, out2[1] = lo; the caller divides by n.  MIN/MAX return the value in
 *      out2[0].  With a communicator (fmb_comm_init) the result covers all shards (empty shards are ignored by MIN/MAX). --------- */
enum {
	FMB_R_SUM = 0,            /* sum x_i */
	FMB_R_SUM_PRODUCT = 1,    /* sum x_i * w_i */
	FMB_R_CENTERED_M2 = 2,    /* sum (x_i - a)^2 */
	FMB_R_CENTERED_M2_W = 3,  /* sum (x_i - a)^2 * w_i */
	FMB_R_MIN = 4,
	FMB_R_MAX = 5
};
int fmb_rv_reduce(int op, fmb_handle x, fmb_handle w, double a, double* out2);
/* The same sums for up to 64 vectors of one length in ONE launch, one synchronisation and (sharded) one exchange: out2[2i], out2[2i+1] =
 * (hi, lo) of sum_p f(x_i[p]).  Replaces the sequence of getAverage calls behind the LIBOR market model's numeraire adjustment
 * (LIBORMarketModelFromCovarianceModel.java:859-876: E[N(0) / N(T_i)] for every tenor date). */
enum {
	FMB_RM_SUM = 0,               /* f(x) = x */
	FMB_RM_SUM_INVERT_MULT = 1    /* f(x) = (1 / x) * a: RandomVariable.invert().mult(a) */
};
int fmb_rv_reduce_many(int op, int count, const fmb_handle* x, double a, double* out2);
/* Order statistics (getQuantile / getQuantileExpectation / getHistogram, :445-575) without a sort and without moving path data between
 * GPUs; with a communicator all three cover the logical vector over all shards (only a 256-bin histogram, a few counts or partial sums
 * are exchanged).  Order = Arrays.sort(double[]): -0.0 < +0.0, NaN above everything. */
/* element of rank `rank` (0-based) of the sorted order: MSB-first radix select, 8 passes of 8 bits over order-preserving keys */
int fmb_rv_select(fmb_handle x, uint64_t rank, double* out);
/* counts[j] = number of elements <= pts[j] (any order of pts, at most 511 per call; NaN elements are never counted) */
int fmb_rv_count_le(fmb_handle x, const double* pts, int npts, uint64_t* counts);
/* out4[0] + out4[1] = double-double sum of the elements strictly between lo and hi, out4[2] = #{x <= lo}, out4[3] = #{x < hi} */
int fmb_rv_range_sum(fmb_handle x, double lo, double hi, double* out4);

/* ---- MT19937 + AS241 Brownian driver: BrownianMotionFromMersenneRandomNumbers
 *      J/montecarlo/BrownianMotionFromMersenneRandomNumbers.java:141-191, MersenneTwister J/randomnumbers/MersenneTwister.java:26-51,
 *      AS241 J/functions/NormalDistribution.java:67-162. -------------------------------------------------------------- */
/* tempered 32-bit outputs word_offset .. word_offset+n of `new MersenneTwister(seed)`, produced on the device via jump-ahead */
int fmb_mt_words(int64_t seed, uint64_t word_offset, uint64_t n, uint32_t* host_out);
/* nextDouble() outputs uniform_offset .. +n (two words each) */
int fmb_mt_uniforms(int64_t seed, uint64_t uniform_offset, uint64_t n, double* host_out);
/* inverseCumulativeDistribution on the device (test hook for the AS241 kernel) */
int fmb_icdf(const double* host_p, uint64_t n, double* host_out);
/* Brownian increments for paths [path_offset, path_offset+paths) of the single sequential stream:
 * out[t*F+f] = handle of length `paths`, value ICDF(u_{((path_offset+p)*T+t)*F+f}) * sqrt_dt[t].
 * All T*F vectors live in one slab in [t][f][path] order. */
int fmb_bm_generate(int32_t seed, int T, int F, uint64_t paths, uint64_t path_offset, const double* sqrt_dt, fmb_handle* out);

/* The uniforms themselves in the same [t][f][path] layout and draw order (out[t*F+f][p] = u_{((path_offset+p)*T+t)*F+f}, bit-exact):
 * IndependentIncrementsFromICDF (J/montecarlo/IndependentIncrementsFromICDF.java:173-206) applies its own inverse distribution functions
 * to them; seed is the long the reference hands to MersenneTwister(seed). */
int fmb_uniforms_generate(int64_t seed, int T, int F, uint64_t paths, uint64_t path_offset, fmb_handle* out);

/* ---- fused Euler schemes: EulerSchemeFromProcessModel J/montecarlo/process/EulerSchemeFromProcessModel.java:170-326.
 *      scheme: 0 EULER, 1 PREDICTOR_CORRECTOR, 2 EULER_FUNCTIONAL, 3 PREDICTOR_CORRECTOR_FUNCTIONAL (:68-73).
 *      dW = the T*F handles of fmb_bm_generate (or any handles of length `paths`).  dt[t] = t_{i+1} - t_i.
 *      out = (T+1)*N handles, [timeIndex][component]; a component frozen at a step aliases the previous handle
 *      (same handle value, one more reference), like discreteProcess[t][c] = discreteProcess[t-1][c] (:285). --------- */
/* BlackScholesModel J/montecarlo/assetderivativevaluation/models/BlackScholesModel.java:60-139; N=1, F>=1 (only dW[.,0] used) */
int fmb_euler_black_scholes(int scheme, int T, int F, uint64_t paths, const double* dt, const fmb_handle* dW,
                            double initial_value, double risk_free_rate, double volatility, fmb_handle* out);
/* HestonModel .../models/HestonModel.java:325-420; N=2 (asset, variance), F=2; heston_scheme 0 REFLECTION, 1 FULL_TRUNCATION.
 * risk_free_rate[t] per step (constant model: all equal). */
int fmb_euler_heston(int scheme, int heston_scheme, int T, uint64_t paths, const double* dt, const fmb_handle* dW,
                     double initial_value, const double* risk_free_rate, double volatility, double theta, double kappa, double xi, double rho,
                     fmb_handle* out);
/* LIBORMarketModelFromCovarianceModel J/montecarlo/interestrate/models/LIBORMarketModelFromCovarianceModel.java:1080-1223.
 * N components, F factors.  factor_loading[t][j][k] (deterministic table, = sigma_j(t_i) * F[j][k]); variance[t][j]
 * (= getCovariance(t,j,j), the "-1/2 sigma^2" term); first_live[t] = firstForwardRateIndex at time index t (:1127-1130);
 * period_length[j]; initial_state[j] = Y_j(0) (log L_j(0) for LOGNORMAL); measure 0 SPOT / 1 TERMINAL; state_space 0 NORMAL / 1 LOGNORMAL;
 * libor_cap (Inf = none). */
int fmb_euler_lmm(int scheme, int measure, int state_space, double libor_cap, int T, int N, int F, uint64_t paths,
                  const double* dt, const fmb_handle* dW, const double* initial_state, const double* period_length,
                  const double* factor_loading, const double* variance, const int32_t* first_live, fmb_handle* out);
/* HullWhiteModel J/montecarlo/interestrate/models/HullWhiteModel.java:277-424; N=2, F=2, scheme EULER.  Per-step deterministic
 * coefficients from the host: drift0[t], drift1[t] multiply the short-rate state x0; fl[t][4] = (l00,l01,l10,l11). */
int fmb_euler_hull_white(int T, uint64_t paths, const double* dt, const fmb_handle* dW, const double* drift0, const double* drift1,
                         const double* fl, fmb_handle* out);

/* ---- regression: MonteCarloConditionalExpectationRegression
 *      J/montecarlo/conditionalexpectation/MonteCarloConditionalExpectationRegression.java:97-150.  basis[i] == 0 means
 *      the deterministic basis function basis_scalar[i].  Moments are SUMS over the local paths in double-double:
 *      XtX_hi/lo[K*K] (symmetric, full) and Xty_hi/lo[K]; the caller all-reduces across shards and divides by n. ------ */
int fmb_regression_moments(int K, const fmb_handle* basis, const double* basis_scalar, fmb_handle y,
                           double* XtX_hi, double* XtX_lo, double* Xty_hi, double* Xty_lo);
/* x = pinv(A) b with the commons-math3 SingularValueDecomposition solver's cut-off (one-sided Jacobi SVD on the host, K x K) */
int fmb_regression_solve_svd(int K, const double* A, const double* b, double* x, double* cond);
/* b_0*x_0 then addProduct(b_i, x_i) in order (:103-107) */
int fmb_regression_predict(int K, const fmb_handle* basis, const double* basis_scalar, const double* x, fmb_handle* out);

/* Device-resident form of the same regression: nothing returns to the host, so a Bermudan backward induction queues one exercise
 * date after the other without a round trip (MonteCarloConditionalExpectationRegression.java:97-150 is one getConditionalExpectation
 * call per exercise date, BermudanSwaption.java:150-156).
 *   fit: ONE pass accumulates the local moments (last CTA merges the per-CTA partials); with a communicator (fmb_comm_init) the
 *   shards' moments are all-gathered on the compute stream and merged in rank order; mean = sum / n_global; the K x K system is solved
 *   on the device with the same Jacobi SVD code as fmb_regression_solve_svd.  The result is a small device vector behind `fit`:
 *   XtX, Xty, coefficients, condition number (fmb_regression_fit_get downloads it; only tests / getLinearRegressionParameters do).
 *   cached_fit != 0: XtX is taken from that earlier fit (the reference caches its solver per estimator instance, :125-138).
 *   n_global = logical number of paths over all shards (0: the local length).
 *   conditional_expectation = fit + predict_fit; basis_pred == NULL: predict on the estimator's basis functions. */
int fmb_regression_fit(int K, const fmb_handle* basis, const double* basis_scalar, fmb_handle y, uint64_t n_global, fmb_handle cached_fit,
                       fmb_handle* fit);
int fmb_regression_fit_get(fmb_handle fit, int K, double* XtX, double* Xty, double* x, double* cond);
int fmb_regression_predict_fit(int K, const fmb_handle* basis, const double* basis_scalar, fmb_handle fit, fmb_handle* out);
int fmb_regression_conditional_expectation(int K, const fmb_handle* basis, const double* basis_scalar, fmb_handle y, uint64_t n_global,
                                           fmb_handle cached_fit, int Kp, const fmb_handle* basis_pred, const double* basis_pred_scalar,
                                           fmb_handle* fit, fmb_handle* out);

/* ---- multi-GPU: one process per GPU, paths sharded by MT19937 jump-ahead (fmb_bm_generate's path_offset).  The only exchange is the
 *      reduction partials of fmb_rv_reduce and the regression moments of fmb_regression_fit: an NCCL all-gather of a few doubles on
 *      the library's compute stream followed by a rank-ordered merge kernel, so every rank holds identical bits.  With a communicator,
 *      fmb_rv_reduce returns the result over ALL shards.  NCCL is bound at run time (dlopen libnccl.so.2, or FMB_NCCL_LIB).
 *      Rank 0 creates the 128-byte id, the host distributes it (any channel), every rank calls fmb_comm_init. ---------------------- */
int fmb_comm_unique_id(unsigned char* id, int len);
int fmb_comm_init(const unsigned char* id, int len, int rank, int world);
int fmb_comm_shutdown(void);
int fmb_comm_info(int* rank, int* world, uint64_t* exchanges);
/* Peer-memory exchange (one node, 2..8 ranks, one process per GPU): instead of an NCCL all-gather per reduction, every rank stores its
 * partials straight into the gather buffers of all ranks over NVLink (CUDA IPC mappings) from one small kernel and waits for the others'
 * flags - the payloads are tens of bytes, so the latency of the exchange is what counts.  Every rank calls fmb_comm_peer_handle (64
 * bytes out), the host gathers the handles in rank order (any channel) and every rank calls fmb_comm_peer_open at the same point of the
 * call sequence.  FMB_EUNSUPPORTED (no peer access between the devices): the NCCL path simply stays in use; handles == NULL switches
 * back to it (for the ranks that could map their peers when another rank could not). */
int fmb_comm_peer_handle(unsigned char* handle, int len);
int fmb_comm_peer_open(const unsigned char* handles, int len);

/* ---- micro-benchmarks used by bench.py to measure the roofline denominators on the box itself ---------------------- */
int fmb_bench_dfma_tflops(double* tflops);          /* dependent-chain-free DFMA loop on all SMs: FP64 pipe peak */
int fmb_bench_copy_gbs(uint64_t bytes, double* gbs); /* device copy read+write bandwidth */

#ifdef __cplusplus
}
#endif
#endif
