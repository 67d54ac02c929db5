// Test shim: the device exp/log of finmath-lib_b200/csrc/fmb_math.cuh compiled for the host (their host branch differs
// only in the reciprocal seed), so that their accuracy can be measured against mpmath without a GPU.
#include "../finmath-lib_b200/csrc/fmb_math.cuh"
extern "C" {
void shim_exp(const double* x, double* y, long n) { for (long i = 0; i < n; i++) y[i] = fmb::fexp(x[i]); }
void shim_log(const double* x, double* y, long n) { for (long i = 0; i < n; i++) y[i] = fmb::flog(x[i]); }
void shim_exp2(const double* x, double* y, long n) { for (long i = 0; i + 1 < n; i += 2) fmb::fexp2(x[i], x[i + 1], y[i], y[i + 1]); }
void shim_log2(const double* x, double* y, long n) { for (long i = 0; i + 1 < n; i += 2) fmb::flog2(x[i], x[i + 1], y[i], y[i + 1]); }
}
