// Reductions over paths and the regression's moment accumulation.
//
// getAverage/getVariance (J/montecarlo/RandomVariableFromDoubleArray.java:286-380) are sequential Kahan sums in the
// reference; a parallel reduction cannot reproduce their rounding order, so every sum here is accumulated in
// double-double (error-free TwoSum per element, warp-shuffle + shared-memory tree in double-double, per-block partials
// combined on the host in fixed order): the result is within one rounding of the exact sum of the same addends and is
// deterministic.  The addends themselves (x*w, (x-a)^2, b_i*b_j) are the rounded products the reference forms.
//
// Streaming kernels: HBM-bound, 8 B read per element and operand; grid = 4 CTAs per SM.
#include "fmb_common.cuh"
#include <cub/cub.cuh>
#include <cmath>
#include <algorithm>

namespace fmb {

struct dd { double hi, lo; };

__host__ __device__ __forceinline__ void twoSum(double a, double b, double& s, double& e) {
	s = a + b;
	const double bb = s - a;
	e = (a - (s - bb)) + (b - bb);
}
__host__ __device__ __forceinline__ void ddAdd(dd& acc, double x) {
	double s, e;
	twoSum(acc.hi, x, s, e);
	acc.hi = s;
	acc.lo += e;
}
__host__ __device__ __forceinline__ void ddMerge(dd& a, const dd& b) {
	double s, e;
	twoSum(a.hi, b.hi, s, e);
	e += a.lo + b.lo;
	twoSum(s, e, a.hi, a.lo);
}
__device__ __forceinline__ dd ddShflDown(const dd& v, int delta) {
	dd r;
	r.hi = __shfl_down_sync(0xffffffffu, v.hi, delta);
	r.lo = __shfl_down_sync(0xffffffffu, v.lo, delta);
	return r;
}
__device__ __forceinline__ void warpReduceDd(dd& v) {
#pragma unroll
	for (int d = 16; d > 0; d >>= 1) { dd o = ddShflDown(v, d); ddMerge(v, o); }
}

__device__ __forceinline__ double jminD(double a, double b) {
	if (a != a) return a;
	if (b != b) return b;
	if (a == 0.0 && b == 0.0) return signbit(a) ? a : b;
	return (a <= b) ? a : b;
}
__device__ __forceinline__ double jmaxD(double a, double b) {
	if (a != a) return a;
	if (b != b) return b;
	if (a == 0.0 && b == 0.0) return signbit(a) ? b : a;
	return (a >= b) ? a : b;
}

static const int RED_THREADS = 256;

template <int OP> __device__ __forceinline__ double addend(double x, double w, double a) {
	switch (OP) {
	case FMB_R_SUM: return x;
	case FMB_R_SUM_PRODUCT: return x * w;
	case FMB_R_CENTERED_M2: return (x - a) * (x - a);
	case FMB_R_CENTERED_M2_W: return (x - a) * (x - a) * w;
	}
	return x;
}

template <int OP> __global__ void __launch_bounds__(RED_THREADS) sumKernel(const double* __restrict__ x, const double* __restrict__ w,
		double a, uint64_t n, double* __restrict__ partials /* [grid][2] */) {
	dd acc0 = {0.0, 0.0}, acc1 = {0.0, 0.0};
	const uint64_t stride = (uint64_t)gridDim.x * RED_THREADS;
	uint64_t i = blockIdx.x * (uint64_t)RED_THREADS + threadIdx.x;
	for (; i + stride < n; i += 2 * stride) {
		const double x0 = x[i], x1 = x[i + stride];
		const double w0 = (OP == FMB_R_SUM_PRODUCT || OP == FMB_R_CENTERED_M2_W) ? w[i] : 0.0;
		const double w1 = (OP == FMB_R_SUM_PRODUCT || OP == FMB_R_CENTERED_M2_W) ? w[i + stride] : 0.0;
		ddAdd(acc0, addend<OP>(x0, w0, a));
		ddAdd(acc1, addend<OP>(x1, w1, a));
	}
	if (i < n) {
		const double w0 = (OP == FMB_R_SUM_PRODUCT || OP == FMB_R_CENTERED_M2_W) ? w[i] : 0.0;
		ddAdd(acc0, addend<OP>(x[i], w0, a));
	}
	ddMerge(acc0, acc1);
	warpReduceDd(acc0);
	__shared__ dd sh[RED_THREADS / 32];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (lane == 0) sh[warp] = acc0;
	__syncthreads();
	if (threadIdx.x == 0) {
		dd t = sh[0];
		for (int k = 1; k < RED_THREADS / 32; k++) ddMerge(t, sh[k]);
		partials[2 * blockIdx.x] = t.hi;
		partials[2 * blockIdx.x + 1] = t.lo;
	}
}

template <bool IS_MAX> __global__ void __launch_bounds__(RED_THREADS) minMaxKernel(const double* __restrict__ x, uint64_t n, double* __restrict__ partials) {
	const uint64_t stride = (uint64_t)gridDim.x * RED_THREADS;
	uint64_t i = blockIdx.x * (uint64_t)RED_THREADS + threadIdx.x;
	double m = x[i < n ? i : 0];
	for (; i < n; i += stride) m = IS_MAX ? jmaxD(m, x[i]) : jminD(m, x[i]);
#pragma unroll
	for (int d = 16; d > 0; d >>= 1) {
		const double o = __shfl_down_sync(0xffffffffu, m, d);
		m = IS_MAX ? jmaxD(m, o) : jminD(m, o);
	}
	__shared__ double sh[RED_THREADS / 32];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (lane == 0) sh[warp] = m;
	__syncthreads();
	if (threadIdx.x == 0) {
		double t = sh[0];
		for (int k = 1; k < RED_THREADS / 32; k++) t = IS_MAX ? jmaxD(t, sh[k]) : jminD(t, sh[k]);
		partials[blockIdx.x] = t;
	}
}

// ---- regression moments: all K(K+1)/2 + K sums in ONE pass over the paths (the reference makes 27 passes for K = 6,
//      MonteCarloConditionalExpectationRegression.java:128-144).  Reads 8 B per stochastic basis function + 8 B for y per path.
struct BasisArgs {
	const double* ptr[8];
	double scalar[8];
};

template <int K> __global__ void __launch_bounds__(RED_THREADS) momentsKernel(BasisArgs b, const double* __restrict__ y, uint64_t n,
		double* __restrict__ partials /* [grid][M][2] */) {
	constexpr int M = K * (K + 1) / 2 + K;
	dd acc[M];
#pragma unroll
	for (int m = 0; m < M; m++) { acc[m].hi = 0.0; acc[m].lo = 0.0; }
	const uint64_t stride = (uint64_t)gridDim.x * RED_THREADS;
	for (uint64_t i = blockIdx.x * (uint64_t)RED_THREADS + threadIdx.x; i < n; i += stride) {
		double v[K];
#pragma unroll
		for (int k = 0; k < K; k++) v[k] = b.ptr[k] ? b.ptr[k][i] : b.scalar[k];
		const double yy = y[i];
		int m = 0;
#pragma unroll
		for (int p = 0; p < K; p++) {
#pragma unroll
			for (int q = p; q < K; q++) { ddAdd(acc[m], v[p] * v[q]); m++; }
		}
#pragma unroll
		for (int p = 0; p < K; p++) { ddAdd(acc[m], yy * v[p]); m++; }
	}
	__shared__ dd sh[RED_THREADS / 32][M];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
	for (int m = 0; m < M; m++) {
		warpReduceDd(acc[m]);
		if (lane == 0) sh[warp][m] = acc[m];
	}
	__syncthreads();
	if (threadIdx.x < M) {
		dd t = sh[0][threadIdx.x];
		for (int k = 1; k < RED_THREADS / 32; k++) ddMerge(t, sh[k][threadIdx.x]);
		partials[((size_t)blockIdx.x * M + threadIdx.x) * 2] = t.hi;
		partials[((size_t)blockIdx.x * M + threadIdx.x) * 2 + 1] = t.lo;
	}
}

// conditional expectation: b_0*x_0, then + b_i*x_i in order (…Regression.java:103-107); 8 B per stochastic basis read, 8 B written
struct PredictCoef { double x[8]; };
template <int K> __global__ void __launch_bounds__(256) predictKernelV(BasisArgs b, PredictCoef c, double* __restrict__ out, uint64_t n) {
	const uint64_t stride = (uint64_t)gridDim.x * 256;
	for (uint64_t i = blockIdx.x * (uint64_t)256 + threadIdx.x; i < n; i += stride) {
		double ce = (b.ptr[0] ? b.ptr[0][i] : b.scalar[0]) * c.x[0];
#pragma unroll
		for (int k = 1; k < K; k++) ce = ce + (b.ptr[k] ? b.ptr[k][i] : b.scalar[k]) * c.x[k];
		out[i] = ce;
	}
}

__global__ void countLeKernel(const double* __restrict__ sorted, uint64_t n, const double* __restrict__ pts, int npts, unsigned long long* __restrict__ counts) {
	const int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= npts) return;
	const double p = pts[k];
	uint64_t lo = 0, hi = n;                       // first index with sorted[idx] > p
	while (lo < hi) {
		const uint64_t mid = (lo + hi) >> 1;
		if (sorted[mid] <= p) lo = mid + 1; else hi = mid;
	}
	counts[k] = lo;
}

static int reduceGrid() { return ctx().smCount * 4; }

template <int K> static void launchMoments(const BasisArgs& b, const double* y, uint64_t n, double* partials, int grid, cudaStream_t s) {
	momentsKernel<K><<<grid, RED_THREADS, 0, s>>>(b, y, n, partials);
}
template <int K> static void launchPredict(const BasisArgs& b, const PredictCoef& c, double* out, uint64_t n, int grid, cudaStream_t s) {
	predictKernelV<K><<<grid, 256, 0, s>>>(b, c, out, n);
}

} // namespace fmb

using namespace fmb;

extern "C" {

int fmb_rv_reduce(int op, fmb_handle x, fmb_handle w, double a, double* out2) {
	FMB_TRY(requireInit());
	if (!out2) return FMB_EINVAL;
	if (op < 0 || op > FMB_R_MAX) { setError("unknown reduction %d", op); return FMB_EINVAL; }
	Context& c = ctx();
	Vec* vx;
	FMB_TRY(lookup(x, &vx));
	const uint64_t n = vx->n;
	const double* wp = nullptr;
	if (op == FMB_R_SUM_PRODUCT || op == FMB_R_CENTERED_M2_W) {
		if (w == 0) { setError("reduction %d needs a weight vector", op); return FMB_EINVAL; }
		FMB_TRY(lookupPtr(w, n, &wp));
	}
	out2[0] = 0.0; out2[1] = 0.0;
	if (n == 0) { out2[0] = (op >= FMB_R_MIN) ? NAN : 0.0; return FMB_OK; }
	int grid = (int)std::min<uint64_t>((uint64_t)reduceGrid(), (n + RED_THREADS - 1) / RED_THREADS);
	std::lock_guard<std::mutex> lk(c.scratchMu);
	FMB_TRY(ensureScratch((size_t)grid * 2 * sizeof(double), (size_t)grid * 2 * sizeof(double)));
	double* dpart = (double*)c.scratch;
	double* hpart = (double*)c.pinned;
	switch (op) {
	case FMB_R_SUM: sumKernel<FMB_R_SUM><<<grid, RED_THREADS, 0, c.stream>>>(vx->ptr, wp, a, n, dpart); break;
	case FMB_R_SUM_PRODUCT: sumKernel<FMB_R_SUM_PRODUCT><<<grid, RED_THREADS, 0, c.stream>>>(vx->ptr, wp, a, n, dpart); break;
	case FMB_R_CENTERED_M2: sumKernel<FMB_R_CENTERED_M2><<<grid, RED_THREADS, 0, c.stream>>>(vx->ptr, wp, a, n, dpart); break;
	case FMB_R_CENTERED_M2_W: sumKernel<FMB_R_CENTERED_M2_W><<<grid, RED_THREADS, 0, c.stream>>>(vx->ptr, wp, a, n, dpart); break;
	case FMB_R_MIN: minMaxKernel<false><<<grid, RED_THREADS, 0, c.stream>>>(vx->ptr, n, dpart); break;
	case FMB_R_MAX: minMaxKernel<true><<<grid, RED_THREADS, 0, c.stream>>>(vx->ptr, n, dpart); break;
	}
	countLaunch();
	FMB_CUDA(cudaGetLastError());
	const size_t cnt = (op >= FMB_R_MIN) ? grid : 2 * (size_t)grid;
	FMB_CUDA(cudaMemcpyAsync(hpart, dpart, cnt * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
	FMB_CUDA(cudaStreamSynchronize(c.stream));
	if (op >= FMB_R_MIN) {
		double m = hpart[0];
		for (int b = 1; b < grid; b++) {
			const double v = hpart[b];
			if (m != m) break;
			if (v != v) { m = v; break; }
			if (op == FMB_R_MIN) { if (v < m || (v == 0.0 && m == 0.0 && std::signbit(v))) m = v; }
			else { if (v > m || (v == 0.0 && m == 0.0 && !std::signbit(v))) m = v; }
		}
		out2[0] = m;
	} else {
		dd t = { hpart[0], hpart[1] };
		for (int b = 1; b < grid; b++) { dd o = { hpart[2 * b], hpart[2 * b + 1] }; ddMerge(t, o); }
		out2[0] = t.hi; out2[1] = t.lo;
	}
	return FMB_OK;
}

static int fillBasis(int K, const fmb_handle* basis, const double* basis_scalar, uint64_t* n, BasisArgs* b) {
	if (K < 1 || K > 8) { setError("regression kernels support 1..8 basis functions (got %d)", K); return FMB_EUNSUPPORTED; }
	bool any = false;
	for (int k = 0; k < K; k++) {
		b->ptr[k] = nullptr;
		b->scalar[k] = basis_scalar ? basis_scalar[k] : 0.0;
		if (basis[k] == 0) continue;
		Vec* v;
		FMB_TRY(lookup(basis[k], &v));
		if (any && v->n != *n) { setError("basis function sizes differ"); return FMB_EINVAL; }
		*n = v->n; any = true;
		b->ptr[k] = v->ptr;
	}
	for (int k = K; k < 8; k++) { b->ptr[k] = nullptr; b->scalar[k] = 0.0; }
	return FMB_OK;
}

int fmb_regression_moments(int K, const fmb_handle* basis, const double* basis_scalar, fmb_handle y,
                           double* XtX_hi, double* XtX_lo, double* Xty_hi, double* Xty_lo) {
	FMB_TRY(requireInit());
	if (!basis || !XtX_hi || !XtX_lo || !Xty_hi || !Xty_lo) return FMB_EINVAL;
	Context& c = ctx();
	BasisArgs b;
	Vec* vy;
	FMB_TRY(lookup(y, &vy));
	uint64_t n = vy->n;
	FMB_TRY(fillBasis(K, basis, basis_scalar, &n, &b));
	if (n != vy->n) { setError("basis functions and dependents differ in size"); return FMB_EINVAL; }
	const int M = K * (K + 1) / 2 + K;
	int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)c.smCount * 2, (n + RED_THREADS - 1) / RED_THREADS));
	std::lock_guard<std::mutex> lk(c.scratchMu);
	const size_t bytes = (size_t)grid * M * 2 * sizeof(double);
	FMB_TRY(ensureScratch(bytes, bytes));
	double* dpart = (double*)c.scratch;
	double* hpart = (double*)c.pinned;
	switch (K) {
	case 1: launchMoments<1>(b, vy->ptr, n, dpart, grid, c.stream); break;
	case 2: launchMoments<2>(b, vy->ptr, n, dpart, grid, c.stream); break;
	case 3: launchMoments<3>(b, vy->ptr, n, dpart, grid, c.stream); break;
	case 4: launchMoments<4>(b, vy->ptr, n, dpart, grid, c.stream); break;
	case 5: launchMoments<5>(b, vy->ptr, n, dpart, grid, c.stream); break;
	case 6: launchMoments<6>(b, vy->ptr, n, dpart, grid, c.stream); break;
	case 7: launchMoments<7>(b, vy->ptr, n, dpart, grid, c.stream); break;
	case 8: launchMoments<8>(b, vy->ptr, n, dpart, grid, c.stream); break;
	}
	countLaunch();
	FMB_CUDA(cudaGetLastError());
	FMB_CUDA(cudaMemcpyAsync(hpart, dpart, bytes, cudaMemcpyDeviceToHost, c.stream));
	FMB_CUDA(cudaStreamSynchronize(c.stream));
	std::vector<dd> tot(M);
	for (int m = 0; m < M; m++) tot[m] = dd{ hpart[2 * m], hpart[2 * m + 1] };
	for (int blk = 1; blk < grid; blk++)
		for (int m = 0; m < M; m++) { dd o = { hpart[((size_t)blk * M + m) * 2], hpart[((size_t)blk * M + m) * 2 + 1] }; ddMerge(tot[m], o); }
	int m = 0;
	for (int p = 0; p < K; p++) for (int q = p; q < K; q++) {
		XtX_hi[p * K + q] = XtX_hi[q * K + p] = tot[m].hi;
		XtX_lo[p * K + q] = XtX_lo[q * K + p] = tot[m].lo;
		m++;
	}
	for (int p = 0; p < K; p++) { Xty_hi[p] = tot[m].hi; Xty_lo[p] = tot[m].lo; m++; }
	return FMB_OK;
}

// One-sided Jacobi (Hestenes) SVD of the K x K matrix A = U S V^T; x = V S^+ U^T b with the commons-math3 3.6.1
// SingularValueDecomposition solver's cut-off  tol = max(K * s_max * 2^-52, sqrt(2^-1022))  (third-party jar).
int fmb_regression_solve_svd(int K, const double* A, const double* b, double* x, double* cond) {
	if (K < 1 || !A || !b || !x) { setError("solve_svd: bad argument"); return FMB_EINVAL; }
	std::vector<double> U(A, A + (size_t)K * K), V((size_t)K * K, 0.0);
	for (int i = 0; i < K; i++) V[(size_t)i * K + i] = 1.0;
	for (int sweep = 0; sweep < 60; sweep++) {
		bool rotated = false;
		for (int p = 0; p < K - 1; p++) for (int q = p + 1; q < K; q++) {
			double alpha = 0, beta = 0, gamma = 0;
			for (int i = 0; i < K; i++) {
				const double up = U[(size_t)i * K + p], uq = U[(size_t)i * K + q];
				alpha += up * up; beta += uq * uq; gamma += up * uq;
			}
			if (gamma == 0.0 || std::fabs(gamma) <= 1e-300) continue;
			if (std::fabs(gamma) <= 0x1.0p-53 * std::sqrt(alpha * beta)) continue;
			rotated = true;
			const double zeta = (beta - alpha) / (2.0 * gamma);
			const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
			const double cs = 1.0 / std::sqrt(1.0 + t * t), sn = cs * t;
			for (int i = 0; i < K; i++) {
				const double up = U[(size_t)i * K + p], uq = U[(size_t)i * K + q];
				U[(size_t)i * K + p] = cs * up - sn * uq;
				U[(size_t)i * K + q] = sn * up + cs * uq;
				const double vp = V[(size_t)i * K + p], vq = V[(size_t)i * K + q];
				V[(size_t)i * K + p] = cs * vp - sn * vq;
				V[(size_t)i * K + q] = sn * vp + cs * vq;
			}
		}
		if (!rotated) break;
	}
	std::vector<double> s(K);
	double smax = 0.0, smin = INFINITY;
	for (int j = 0; j < K; j++) {
		double nn = 0;
		for (int i = 0; i < K; i++) nn += U[(size_t)i * K + j] * U[(size_t)i * K + j];
		s[j] = std::sqrt(nn);
		smax = std::max(smax, s[j]); smin = std::min(smin, s[j]);
	}
	if (cond) *cond = smax / smin;
	const double tol = std::max((double)K * smax * 0x1.0p-52, std::sqrt(0x1.0p-1022));
	for (int k = 0; k < K; k++) x[k] = 0.0;
	for (int j = 0; j < K; j++) {
		if (s[j] <= tol) continue;
		double ub = 0;                              // (u_j . b) / s_j, with u_j = U[:,j] / s_j
		for (int i = 0; i < K; i++) ub += U[(size_t)i * K + j] * b[i];
		const double wgt = ub / (s[j] * s[j]);
		for (int k = 0; k < K; k++) x[k] += V[(size_t)k * K + j] * wgt;
	}
	return FMB_OK;
}

int fmb_regression_predict(int K, const fmb_handle* basis, const double* basis_scalar, const double* x, fmb_handle* out) {
	FMB_TRY(requireInit());
	if (!basis || !x || !out) return FMB_EINVAL;
	Context& c = ctx();
	BasisArgs b;
	uint64_t n = 0;
	FMB_TRY(fillBasis(K, basis, basis_scalar, &n, &b));
	bool any = false;
	for (int k = 0; k < K; k++) any = any || b.ptr[k];
	if (!any) { setError("predict: all basis functions deterministic; stays on the host"); return FMB_EINVAL; }
	PredictCoef pc;
	for (int k = 0; k < 8; k++) pc.x[k] = k < K ? x[k] : 0.0;
	double* dst;
	FMB_TRY(newVec(n, out, &dst));
	if (n == 0) return FMB_OK;
	const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)c.smCount * 8, (n + 255) / 256));
	switch (K) {
	case 1: launchPredict<1>(b, pc, dst, n, grid, c.stream); break;
	case 2: launchPredict<2>(b, pc, dst, n, grid, c.stream); break;
	case 3: launchPredict<3>(b, pc, dst, n, grid, c.stream); break;
	case 4: launchPredict<4>(b, pc, dst, n, grid, c.stream); break;
	case 5: launchPredict<5>(b, pc, dst, n, grid, c.stream); break;
	case 6: launchPredict<6>(b, pc, dst, n, grid, c.stream); break;
	case 7: launchPredict<7>(b, pc, dst, n, grid, c.stream); break;
	case 8: launchPredict<8>(b, pc, dst, n, grid, c.stream); break;
	}
	countLaunch();
	FMB_CUDA(cudaGetLastError());
	return FMB_OK;
}

int fmb_rv_sorted(fmb_handle x, fmb_handle* out) {
	FMB_TRY(requireInit());
	if (!out) return FMB_EINVAL;
	Context& c = ctx();
	Vec* vx;
	FMB_TRY(lookup(x, &vx));
	const uint64_t n = vx->n;
	if (n > 0x7fffffffull) { setError("sort: more than 2^31-1 elements"); return FMB_EUNSUPPORTED; }
	double* dst;
	FMB_TRY(newVec(n, out, &dst));
	if (n == 0) return FMB_OK;
	size_t tmpBytes = 0;
	FMB_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tmpBytes, vx->ptr, dst, (int)n, 0, 64, c.stream));
	void* tmp;
	FMB_TRY(poolAlloc(tmpBytes, &tmp));
	cudaError_t e = cub::DeviceRadixSort::SortKeys(tmp, tmpBytes, vx->ptr, dst, (int)n, 0, 64, c.stream);
	countLaunch(4);
	poolFree(tmp, tmpBytes);
	if (e != cudaSuccess) { setError("sort: %s", cudaGetErrorString(e)); return FMB_ECUDA; }
	return FMB_OK;
}

int fmb_rv_count_le(fmb_handle sorted, const double* pts, int npts, uint64_t* counts) {
	FMB_TRY(requireInit());
	if (npts < 0 || (npts && (!pts || !counts))) return FMB_EINVAL;
	if (npts == 0) return FMB_OK;
	Context& c = ctx();
	Vec* vs;
	FMB_TRY(lookup(sorted, &vs));
	void* dp; void* dc;
	FMB_TRY(poolAlloc(npts * sizeof(double), &dp));
	FMB_TRY(poolAlloc(npts * sizeof(uint64_t), &dc));
	cudaError_t e = cudaMemcpyAsync(dp, pts, npts * sizeof(double), cudaMemcpyHostToDevice, c.stream);
	countLeKernel<<<(npts + 127) / 128, 128, 0, c.stream>>>(vs->ptr, vs->n, (const double*)dp, npts, (unsigned long long*)dc);
	countLaunch();
	if (e == cudaSuccess) e = cudaMemcpyAsync(counts, dc, npts * sizeof(uint64_t), cudaMemcpyDeviceToHost, c.stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(c.stream);
	poolFree(dp, npts * sizeof(double));
	poolFree(dc, npts * sizeof(uint64_t));
	if (e != cudaSuccess) { setError("count_le: %s", cudaGetErrorString(e)); return FMB_ECUDA; }
	return FMB_OK;
}

} // extern "C"
