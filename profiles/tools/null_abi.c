/* null_abi.c — PROFILING STUB, not part of the product and never loaded by it.
 *
 * Exports the symbols of include/finmath_b200.h with bodies that only hand out handles and remember vector lengths: no device,
 * no arithmetic.  profiles/tools/host_profile.py points the ctypes binding at it to measure what the HOST side of a valuation costs
 * (Python mirror + binding) without a GPU, e.g. in the build container.  Reductions return 0.5, downloads return zeros.
 *   gcc -O2 -shared -fPIC -I include profiles/tools/null_abi.c -o /tmp/libnull_abi.so
 */
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#include "finmath_b200.h"

#define MAXH (1u << 22)
static uint64_t next_h = 0x1000, launches = 0;
static uint64_t *sizes;
static fmb_handle mk(uint64_t n) { if (!sizes) sizes = calloc(MAXH, 8); fmb_handle h = next_h++; sizes[h % MAXH] = n; return h; }
static uint64_t sz(fmb_handle h) { return sizes ? sizes[h % MAXH] : 0; }

int fmb_init(int d) { return 0; }
int fmb_shutdown(void) { return 0; }
int fmb_is_initialized(void) { return 1; }
const char* fmb_last_error(void) { return "null abi"; }
int fmb_device_count(int* c) { *c = 1; return 0; }
int fmb_device_name(char* b, int l) { strncpy(b, "null", l); return 0; }
int fmb_synchronize(void) { return 0; }
int fmb_set_fp_mode(int m) { return 0; }
int fmb_get_fp_mode(int* m) { *m = 0; return 0; }
int fmb_timer_start(void) { return 0; }
int fmb_timer_stop_ms(float* ms) { *ms = 1; return 0; }
int fmb_kernel_launch_count(uint64_t* c) { *c = launches; return 0; }
int fmb_rv_create(uint64_t n, fmb_handle* o) { *o = mk(n); return 0; }
int fmb_rv_upload(const double* h, uint64_t n, fmb_handle* o) { *o = mk(n); return 0; }
int fmb_rv_fill(double v, uint64_t n, fmb_handle* o) { *o = mk(n); launches++; return 0; }
int fmb_rv_download(fmb_handle h, double* host, uint64_t n) { memset(host, 0, n * 8); return 0; }
int fmb_rv_get(fmb_handle h, uint64_t i, double* o) { *o = 0; return 0; }
int fmb_rv_size(fmb_handle h, uint64_t* n) { *n = sz(h); return 0; }
int fmb_rv_retain(fmb_handle h) { return 0; }
int fmb_rv_free(fmb_handle h) { return 0; }
int fmb_rv_free_many(const fmb_handle* h, uint64_t n) { return 0; }
int fmb_rv_device_ptr(fmb_handle h, void** p) { *p = 0; return 0; }
int fmb_pool_stats(uint64_t* a, uint64_t* b, uint64_t* c) { if (a) *a = 0; if (b) *b = 0; if (c) *c = 0; return 0; }
int fmb_pool_trim(void) { return 0; }
int fmb_rv_unary(int op, fmb_handle x, double a, fmb_handle* o) { *o = mk(sz(x)); launches++; return 0; }
int fmb_rv_binary(int op, fmb_handle x, double sx, fmb_handle y, double sy, fmb_handle* o) { *o = mk(sz(x ? x : y)); launches++; return 0; }
int fmb_rv_ternary(int op, fmb_handle x, double sx, fmb_handle y, double sy, fmb_handle z, double sz_, double a, fmb_handle* o) {
	*o = mk(sz(x ? x : (y ? y : z))); launches++; return 0; }
int fmb_rv_accrue_chain(int n, const fmb_handle* r, const double* d, double div, fmb_handle* o) { *o = mk(sz(r[0])); launches++; return 0; }
int fmb_rv_eval_chain(int n, const unsigned char* code, int s, const fmb_handle* l, int nl, const double* sc, int ns, fmb_handle* o) { *o = mk(sz(l[0])); launches++; return 0; }
int fmb_rv_accrue_prefix(int n, double s0, const fmb_handle* r, const double* d, fmb_handle* o) { for (int k = 0; k < n; k++) o[k] = mk(sz(r[0])); launches++; return 0; }
int fmb_rv_reduce_many(int op, int n, const fmb_handle* x, double a, double* o) { for (int k = 0; k < n; k++) { o[2 * k] = 0.5 * (double)sz(x[k]); o[2 * k + 1] = 0; } launches++; return 0; }
int fmb_rv_reduce(int op, fmb_handle x, fmb_handle w, double a, double* o) { o[0] = 0.5 * (double)sz(x); o[1] = 0; launches++; return 0; }
int fmb_rv_select(fmb_handle x, uint64_t r, double* o) { *o = 0; return 0; }
int fmb_rv_range_sum(fmb_handle x, double lo, double hi, double* o) { o[0] = o[1] = o[2] = o[3] = 0; return 0; }
int fmb_rv_count_le(fmb_handle s, const double* p, int n, uint64_t* c) { for (int i = 0; i < n; i++) c[i] = 0; return 0; }
int fmb_mt_words(int64_t s, uint64_t o, uint64_t n, uint32_t* out) { memset(out, 0, 4 * n); return 0; }
int fmb_mt_uniforms(int64_t s, uint64_t o, uint64_t n, double* out) { memset(out, 0, 8 * n); return 0; }
int fmb_icdf(const double* p, uint64_t n, double* o) { memset(o, 0, 8 * n); return 0; }
int fmb_bm_generate(int32_t seed, int T, int F, uint64_t paths, uint64_t off, const double* sq, fmb_handle* out) {
	for (int i = 0; i < T * F; i++) out[i] = mk(paths); launches += 10; return 0; }
int fmb_uniforms_generate(int64_t seed, int T, int F, uint64_t paths, uint64_t off, fmb_handle* out) { for (int i = 0; i < T * F; i++) out[i] = mk(paths); launches += 10; return 0; }
static void proc(int T, int N, uint64_t paths, fmb_handle* out) { for (int j = 0; j < N; j++) out[j] = 0; for (int i = N; i < (T + 1) * N; i++) out[i] = mk(paths); launches++; }
int fmb_euler_black_scholes(int s, int T, int F, uint64_t paths, const double* dt, const fmb_handle* dW, double a, double b, double c, fmb_handle* out) { proc(T, 1, paths, out); return 0; }
int fmb_euler_heston(int s, int hs, int T, uint64_t paths, const double* dt, const fmb_handle* dW, double iv, const double* r, double v, double th, double k, double xi, double rho, fmb_handle* out) { proc(T, 2, paths, out); return 0; }
int fmb_euler_hull_white(int T, uint64_t paths, const double* dt, const fmb_handle* dW, const double* d0, const double* d1, const double* fl, fmb_handle* out) { proc(T, 2, paths, out); return 0; }
int fmb_euler_lmm(int scheme, int measure, int ss, double cap, int T, int N, int F, uint64_t paths, const double* dt, const fmb_handle* dW, const double* is,
                  const double* pl, const double* fl, const double* var, const int32_t* first, fmb_handle* out) {
	for (int j = 0; j < N; j++) out[j] = 0;
	for (int t = 1; t <= T; t++) for (int j = 0; j < N; j++) out[t * N + j] = j >= first[t - 1] ? mk(paths) : out[(t - 1) * N + j];
	launches++; return 0; }
int fmb_regression_moments(int K, const fmb_handle* b, const double* bs, fmb_handle y, double* a, double* c, double* d, double* e) {
	for (int i = 0; i < K * K; i++) { a[i] = (i % (K + 1) == 0) ? 1.0 : 0.0; c[i] = 0; } for (int i = 0; i < K; i++) { d[i] = 1; e[i] = 0; } launches++; return 0; }
int fmb_regression_solve_svd(int K, const double* A, const double* b, double* x, double* cond) { for (int i = 0; i < K; i++) x[i] = 1; if (cond) *cond = 1; return 0; }
int fmb_regression_predict(int K, const fmb_handle* b, const double* bs, const double* x, fmb_handle* o) { uint64_t n = 0; for (int i = 0; i < K; i++) if (b[i]) n = sz(b[i]); *o = mk(n); launches++; return 0; }
int fmb_regression_fit(int K, const fmb_handle* b, const double* bs, fmb_handle y, uint64_t ng, fmb_handle c, fmb_handle* fit) { *fit = mk(96); launches++; return 0; }
int fmb_regression_fit_get(fmb_handle f, int K, double* a, double* b, double* x, double* cond) { if (x) for (int i = 0; i < K; i++) x[i] = 1; if (cond) *cond = 1; return 0; }
int fmb_regression_predict_fit(int K, const fmb_handle* b, const double* bs, fmb_handle fit, fmb_handle* o) { uint64_t n = 0; for (int i = 0; i < K; i++) if (b[i]) n = sz(b[i]); *o = mk(n); launches++; return 0; }
int fmb_regression_conditional_expectation(int K, const fmb_handle* b, const double* bs, fmb_handle y, uint64_t ng, fmb_handle c, int Kp, const fmb_handle* bp,
                                           const double* bps, fmb_handle* fit, fmb_handle* o) { *fit = mk(96); *o = mk(sz(y)); launches += 2; return 0; }
int fmb_comm_unique_id(unsigned char* id, int len) { memset(id, 0, 128); return 0; }
int fmb_comm_init(const unsigned char* id, int len, int r, int w) { return 0; }
int fmb_comm_shutdown(void) { return 0; }
int fmb_comm_peer_handle(unsigned char* h, int len) { return 6; }
int fmb_comm_peer_open(const unsigned char* h, int len) { return 6; }
int fmb_comm_info(int* r, int* w, uint64_t* e) { if (r) *r = 0; if (w) *w = 1; if (e) *e = 0; return 0; }
int fmb_bench_dfma_tflops(double* t) { *t = 1; return 0; }
int fmb_bench_copy_gbs(uint64_t b, double* g) { *g = 1; return 0; }
