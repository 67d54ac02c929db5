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

// Warp reduction of CNT double-doubles per lane with 1/5 of the shuffles of CNT separate trees: in every round a lane hands one half of
// its values to the partner (lane ^ D) and merges the partner's other half into its own, so the number of values per lane halves while
// the lanes covered by each double: after five rounds a lane holds the complete sums of at most two values.  ddMerge is commutative and
// the partners are those of the shuffle tree (16, 8, 4, 2, 1), so every sum is the one warpReduceDd produces, bit for bit.
template <int CNT, int D> struct DdButterfly {
	static constexpr int HALF = (CNT + 1) / 2;
	static __device__ __forceinline__ void run(dd* acc, int lane) {
		const bool upper = (lane & D) != 0;
#pragma unroll
		for (int j = 0; j < HALF; j++) {
			const dd a = acc[j];
			dd b = {0.0, 0.0};
			if (j + HALF < CNT) b = acc[j + HALF];
			dd keep, send, recv;
			keep.hi = upper ? b.hi : a.hi; keep.lo = upper ? b.lo : a.lo;
			send.hi = upper ? a.hi : b.hi; send.lo = upper ? a.lo : b.lo;
			recv.hi = __shfl_xor_sync(0xffffffffu, send.hi, D);
			recv.lo = __shfl_xor_sync(0xffffffffu, send.lo, D);
			ddMerge(keep, recv);
			acc[j] = keep;
		}
		DdButterfly<HALF, D / 2>::run(acc, lane);
	}
	// index (among the CNT values the round started with) of the value this lane ends up holding in slot j; ok: it is a real one
	static __device__ __forceinline__ int origin(int j, int lane, bool& ok) {
		int pos = DdButterfly<HALF, D / 2>::origin(j, lane, ok);
		if (lane & D) pos += HALF;
		ok = ok && pos < CNT;
		return pos;
	}
};
template <int CNT> struct DdButterfly<CNT, 0> {
	static constexpr int LEFT = CNT;
	static __device__ __forceinline__ void run(dd*, int) {}
	static __device__ __forceinline__ int origin(int j, int, bool& ok) { ok = j < CNT; return j; }
};
template <int M> struct DdButterflyLeft { static constexpr int value = (((((M + 1) / 2 + 1) / 2 + 1) / 2 + 1) / 2 + 1) / 2; };

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

// ---- last-block finalisation -------------------------------------------------------------------------------------------
// Every CTA writes its partial, makes it visible (__threadfence) and takes a ticket; the CTA that draws the last ticket merges all
// partials in a FIXED order (so the result is deterministic for a given grid) and writes the final value - to mapped pinned host
// memory when the caller is waiting for a double, to the communicator's send buffer when shards are exchanged, or it goes straight on
// to the regression solve.  atomicInc wraps the ticket back to zero, ready for the next launch (launches of these kernels are
// serialised on the compute stream).
__device__ __forceinline__ bool drawLastTicketOf(unsigned int* ticket, unsigned int total) {
	__shared__ bool isLast;
	__threadfence();
	__syncthreads();
	if (threadIdx.x == 0) isLast = (atomicInc(ticket, total - 1) == total - 1);
	__syncthreads();
	if (isLast) __threadfence();
	return isLast;
}
__device__ __forceinline__ bool drawLastTicket(unsigned int* ticket) { return drawLastTicketOf(ticket, gridDim.x); }

template <int OP> __device__ __forceinline__ double addend(double x, double w, double a) {
	switch (OP) {
	case FMB_R_SUM: return x;
	case FMB_R_SUM_PRODUCT: return x * w;
	case FMB_R_CENTERED_M2: return (x - a) * (x - a);
	case FMB_R_CENTERED_M2_W: return (x - a) * (x - a) * w;
	}
	return x;
}

// block-level merge of one double-double per thread; result valid in thread 0
__device__ __forceinline__ dd blockReduceDd(dd v) {
	__shared__ dd sh[RED_THREADS / 32];
	warpReduceDd(v);
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	__syncthreads();                                   // (the buffer may still be read from a previous call)
	if (lane == 0) sh[warp] = v;
	__syncthreads();
	if (threadIdx.x == 0) { for (int k = 1; k < RED_THREADS / 32; k++) ddMerge(v, sh[k]); }
	return v;
}

// out2[0..1] = (hi, lo) of the sum over all n elements
// shard merge in the finalising CTA (peer exchange): payload (hi, lo) of this rank -> all ranks -> rank-ordered double-double merge
__device__ inline void finishSumShards(const PeerArgs& px, dd t, double* __restrict__ out2) {
	__shared__ double payload[2];
	if (threadIdx.x == 0) { payload[0] = t.hi; payload[1] = t.lo; }
	__syncthreads();
	peerExchangeBlock(px, payload, 2);
	if (threadIdx.x == 0) {
		const double* g = peerGathered(px);
		dd m = { __ldcg(g), __ldcg(g + 1) };
		for (int r = 1; r < px.world; r++) { const dd o = { __ldcg(g + 2 * r), __ldcg(g + 2 * r + 1) }; ddMerge(m, o); }
		out2[0] = m.hi; out2[1] = m.lo;
	}
}

// px.world > 1: the finalising CTA also exchanges the result with the other ranks (peer memory) and writes the merged value to out2
template <int OP> __global__ void __launch_bounds__(RED_THREADS) sumKernel(const double* __restrict__ x, const double* __restrict__ w,
		double a, uint64_t n, double* __restrict__ partials /* [grid][2] */, unsigned int* ticket, double* __restrict__ out2, const PeerArgs px) {
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
	dd t = blockReduceDd(acc0);
	if (gridDim.x == 1) {
		if (px.world > 1) { finishSumShards(px, t, out2); return; }
		if (threadIdx.x == 0) { out2[0] = t.hi; out2[1] = t.lo; }
		return;
	}
	if (threadIdx.x == 0) { partials[2 * blockIdx.x] = t.hi; partials[2 * blockIdx.x + 1] = t.lo; }
	if (!drawLastTicket(ticket)) return;
	dd m = {0.0, 0.0};
	for (unsigned int b = threadIdx.x; b < gridDim.x; b += RED_THREADS) { const dd o = { __ldcg(partials + 2 * b), __ldcg(partials + 2 * b + 1) }; ddMerge(m, o); }
	m = blockReduceDd(m);
	if (px.world > 1) { finishSumShards(px, m, out2); return; }
	if (threadIdx.x == 0) { out2[0] = m.hi; out2[1] = m.lo; }
}

// out2[0] = min / max with Math.min / Math.max semantics (NaN-propagating, -0.0 < +0.0), out2[1] = 1.0 (this shard holds data)
template <bool IS_MAX> __device__ __forceinline__ double blockReduceMinMax(double m) {
	__shared__ double sh[RED_THREADS / 32];
#pragma unroll
	for (int d = 16; d > 0; d >>= 1) {
		const double o = __shfl_down_sync(0xffffffffu, m, d);
		m = IS_MAX ? jmaxD(m, o) : jminD(m, o);
	}
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	__syncthreads();
	if (lane == 0) sh[warp] = m;
	__syncthreads();
	if (threadIdx.x == 0) { for (int k = 1; k < RED_THREADS / 32; k++) m = IS_MAX ? jmaxD(m, sh[k]) : jminD(m, sh[k]); }
	return m;
}

// (value, "this shard holds data") of this rank -> all ranks -> min / max over the shards that hold data, in rank order
template <bool IS_MAX> __device__ inline void finishMinMaxShards(const PeerArgs& px, double v, double* __restrict__ out2) {
	__shared__ double payload[2];
	if (threadIdx.x == 0) { payload[0] = v; payload[1] = 1.0; }
	__syncthreads();
	peerExchangeBlock(px, payload, 2);
	if (threadIdx.x == 0) {
		const double* g = peerGathered(px);
		bool any = false;
		double t = 0.0;
		for (int r = 0; r < px.world; r++) {
			if (__ldcg(g + 2 * r + 1) == 0.0) continue;
			const double o = __ldcg(g + 2 * r);
			t = any ? (IS_MAX ? jmaxD(t, o) : jminD(t, o)) : o;
			any = true;
		}
		out2[0] = any ? t : NAN;
		out2[1] = any ? 1.0 : 0.0;
	}
}

template <bool IS_MAX> __global__ void __launch_bounds__(RED_THREADS) minMaxKernel(const double* __restrict__ x, uint64_t n, double* __restrict__ partials,
		unsigned int* ticket, double* __restrict__ out2, const PeerArgs px) {
	const uint64_t stride = (uint64_t)gridDim.x * RED_THREADS;
	uint64_t i = blockIdx.x * (uint64_t)RED_THREADS + threadIdx.x;
	double m = x[i < n ? i : 0];
	for (; i < n; i += stride) m = IS_MAX ? jmaxD(m, x[i]) : jminD(m, x[i]);
	m = blockReduceMinMax<IS_MAX>(m);
	if (gridDim.x == 1) {
		if (px.world > 1) { finishMinMaxShards<IS_MAX>(px, m, out2); return; }
		if (threadIdx.x == 0) { out2[0] = m; out2[1] = 1.0; }
		return;
	}
	if (threadIdx.x == 0) partials[blockIdx.x] = m;
	if (!drawLastTicket(ticket)) return;
	double t = __ldcg(partials + (threadIdx.x < gridDim.x ? threadIdx.x : 0));
	for (unsigned int b = threadIdx.x; b < gridDim.x; b += RED_THREADS) { const double o = __ldcg(partials + b); t = IS_MAX ? jmaxD(t, o) : jminD(t, o); }
	t = blockReduceMinMax<IS_MAX>(t);
	if (px.world > 1) { finishMinMaxShards<IS_MAX>(px, t, out2); return; }
	if (threadIdx.x == 0) { out2[0] = t; out2[1] = 1.0; }
}

// merge of the gathered shard results in rank order (one launch, one warp): sums as double-double pairs, min / max with the
// reference's semantics; shards without data (flag 0) are skipped, so an empty shard never turns a result into NaN.
// gathered: [world][count] doubles; kind 0: count/2 (hi, lo) pairs -> out[2m], out[2m+1]; kind 1 / 2: (value, flag) -> out[0] = min / max
__global__ void mergeShardsKernel(const double* __restrict__ gathered, int world, int count, int kind, double* __restrict__ out) {
	const int m = threadIdx.x;
	if (kind == 0) {
		if (2 * m >= count) return;
		dd t = { gathered[2 * m], gathered[2 * m + 1] };
		for (int r = 1; r < world; r++) { const dd o = { gathered[(size_t)r * count + 2 * m], gathered[(size_t)r * count + 2 * m + 1] }; ddMerge(t, o); }
		out[2 * m] = t.hi; out[2 * m + 1] = t.lo;
	} else if (m == 0) {
		bool any = false;
		double t = 0.0;
		for (int r = 0; r < world; r++) {
			if (gathered[(size_t)r * count + 1] == 0.0) continue;
			const double v = gathered[(size_t)r * count];
			t = any ? (kind == 2 ? jmaxD(t, v) : jminD(t, v)) : v;
			any = true;
		}
		out[0] = any ? t : NAN;
		out[1] = any ? 1.0 : 0.0;
	}
}

// ---- the same sum over MANY vectors in one launch ------------------------------------------------------------------------------
// The LMM's numeraire adjustment needs E[N(0) / N(T_i)] for every tenor date (LIBORMarketModelFromCovarianceModel.java:859-876 evaluates
// them one getAverage at a time: one kernel, one host synchronisation and - sharded - one rendezvous of all ranks per date).  Here the
// sums of up to 64 vectors are ONE launch: blockIdx.y = vector, per-vector last-block merge exactly as sumKernel, and the CTA that
// completes the LAST vector exchanges all results with the other ranks at once (peer memory), merges the shards and writes them out.
static const int REDUCE_MANY_MAX = 64;
struct ManyArgs { const double* x[REDUCE_MANY_MAX]; };
template <int OP> __device__ __forceinline__ double addendMany(double x, double a) { return OP == FMB_RM_SUM ? x : (1.0 / x) * a; }

template <int OP> __global__ void __launch_bounds__(RED_THREADS) sumManyKernel(const __grid_constant__ ManyArgs v, int count, double a, uint64_t n,
		double* __restrict__ partials /* [count][gridDim.x][2] */, unsigned int* tickets /* [count] per vector, [count] = vectors done */,
		double* __restrict__ results /* [count][2], device */, double* __restrict__ out /* [count][2] */, const PeerArgs px) {
	const int vi = blockIdx.y;
	const double* __restrict__ x = v.x[vi];
	dd acc0 = {0.0, 0.0}, acc1 = {0.0, 0.0};
	const uint64_t stride = (uint64_t)gridDim.x * RED_THREADS;
	uint64_t i = blockIdx.x * (uint64_t)RED_THREADS + threadIdx.x;
	for (; i + stride < n; i += 2 * stride) {
		const double x0 = x[i], x1 = x[i + stride];
		ddAdd(acc0, addendMany<OP>(x0, a));
		ddAdd(acc1, addendMany<OP>(x1, a));
	}
	if (i < n) ddAdd(acc0, addendMany<OP>(x[i], a));
	ddMerge(acc0, acc1);
	dd t = blockReduceDd(acc0);
	if (gridDim.x > 1) {
		if (threadIdx.x == 0) { partials[((size_t)vi * gridDim.x + blockIdx.x) * 2] = t.hi; partials[((size_t)vi * gridDim.x + blockIdx.x) * 2 + 1] = t.lo; }
		if (!drawLastTicketOf(tickets + vi, gridDim.x)) return;
		dd m = {0.0, 0.0};
		const double* row = partials + (size_t)vi * gridDim.x * 2;
		for (unsigned int b = threadIdx.x; b < gridDim.x; b += RED_THREADS) { const dd o = { __ldcg(row + 2 * b), __ldcg(row + 2 * b + 1) }; ddMerge(m, o); }
		t = blockReduceDd(m);
	}
	if (threadIdx.x == 0) { results[2 * vi] = t.hi; results[2 * vi + 1] = t.lo; }
	if (!drawLastTicketOf(tickets + count, gridDim.y)) return;
	// every vector's local sum is in results: exchange (sharded), merge in rank order, write
	if (px.world > 1) {
		peerExchangeBlock(px, results, 2 * count);
		const double* g = peerGathered(px);
		for (int m = threadIdx.x; m < count; m += RED_THREADS) {
			dd s = { __ldcg(g + 2 * m), __ldcg(g + 2 * m + 1) };
			for (int r = 1; r < px.world; r++) { const dd o = { __ldcg(g + (size_t)r * 2 * count + 2 * m), __ldcg(g + (size_t)r * 2 * count + 2 * m + 1) }; ddMerge(s, o); }
			out[2 * m] = s.hi; out[2 * m + 1] = s.lo;
		}
	} else {
		for (int m = threadIdx.x; m < 2 * count; m += RED_THREADS) out[m] = __ldcg(results + m);
	}
}

// ---- regression --------------------------------------------------------------------------------------------------------
// x = pinv(A) b for the symmetric K x K moment matrix: one-sided Jacobi (Hestenes) SVD A = U S V^T, x = V S^+ U^T b with the
// commons-math3 3.6.1 SingularValueDecomposition solver's cut-off  tol = max(K * s_max * 2^-52, sqrt(2^-1022))  (third-party jar,
// SURVEY.md §8c).  ONE implementation for the host entry point (fmb_regression_solve_svd) and for the device-resident regression, so
// both give the same coefficients.  U holds A on entry (row-major, K x K); V, s are work space.
__host__ __device__ inline void jacobiPinvSolve(int K, double* U, double* V, double* s, const double* b, double* x, double* cond) {
	for (int i = 0; i < K * K; i++) V[i] = 0.0;
	for (int i = 0; i < K; i++) V[i * K + i] = 1.0;
	for (int sweep = 0; sweep < 60; sweep++) {
		bool rotated = false;
		for (int p = 0; p < K - 1; p++) for (int q = p + 1; q < K; q++) {
			double alpha = 0, beta = 0, gamma = 0;
			for (int i = 0; i < K; i++) {
				const double up = U[i * K + p], uq = U[i * K + q];
				alpha += up * up; beta += uq * uq; gamma += up * uq;
			}
			if (gamma == 0.0 || fabs(gamma) <= 1e-300) continue;
			if (fabs(gamma) <= 0x1.0p-53 * sqrt(alpha * beta)) continue;
			rotated = true;
			const double zeta = (beta - alpha) / (2.0 * gamma);
			const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
			const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
			for (int i = 0; i < K; i++) {
				const double up = U[i * K + p], uq = U[i * K + q];
				U[i * K + p] = cs * up - sn * uq;
				U[i * K + q] = sn * up + cs * uq;
				const double vp = V[i * K + p], vq = V[i * K + q];
				V[i * K + p] = cs * vp - sn * vq;
				V[i * K + q] = sn * vp + cs * vq;
			}
		}
		if (!rotated) break;
	}
	double smax = 0.0, smin = INFINITY;
	for (int j = 0; j < K; j++) {
		double nn = 0;
		for (int i = 0; i < K; i++) nn += U[i * K + j] * U[i * K + j];
		s[j] = sqrt(nn);
		smax = s[j] > smax ? s[j] : smax; smin = s[j] < smin ? s[j] : smin;
	}
	if (cond) *cond = smax / smin;
	const double tolA = (double)K * smax * 0x1.0p-52, tolB = 1.4916681462400413e-154 /* sqrt(2^-1022) */;
	const double tol = tolA > tolB ? tolA : tolB;
	for (int k = 0; k < K; k++) x[k] = 0.0;
	for (int j = 0; j < K; j++) {
		if (s[j] <= tol) continue;
		double ub = 0;                              // (u_j . b) / s_j, with u_j = U[:,j] / s_j
		for (int i = 0; i < K; i++) ub += U[i * K + j] * b[i];
		const double wgt = ub / (s[j] * s[j]);
		for (int k = 0; k < K; k++) x[k] += V[k * K + j] * wgt;
	}
}

// K <= 8: one-sided Jacobi with the ROUND-ROBIN (tournament) ordering - a sweep is KP - 1 rounds of KP / 2 rotations on disjoint column
// pairs (KP = K rounded up to even), and the rotations of a round are independent.  The device version runs them on ONE WARP: lane = (pair,
// row), 8 lanes per pair; the three column dot products of a pair are 8-lane butterfly reductions, every lane of the pair evaluates the
// rotation, each lane updates its own row of U and V in shared memory.  Compact code (one small loop body - a fully unrolled single-thread
// version spent more time on cold instruction-cache lines than on arithmetic) and ~40 rounds instead of ~120 sequential rotations.
// The host version below (fmb_regression_solve_svd, K <= 8) performs the same operations in the same order - including the butterfly's
// summation tree ((t0+t4)+(t2+t6))+((t1+t5)+(t3+t7)) - so both sides produce identical coefficients.
__host__ __device__ inline void roundRobinPair(int KP, int r, int m, int& p, int& q) {
	const int R = KP - 1;
	const int pa = m == 0 ? KP - 1 : (r + m) % R, pb = m == 0 ? r : (r - m + R) % R;
	p = pa < pb ? pa : pb; q = pa < pb ? pb : pa;
}
// rotation of the column pair from its three dot products; false: the pair is already orthogonal to working precision
__host__ __device__ inline bool jacobiRotation(double alpha, double beta, double gamma, double& cs, double& sn) {
	cs = 1.0; sn = 0.0;
	if (gamma == 0.0 || fabs(gamma) <= 1e-300 || gamma * gamma <= 0x1.0p-106 * (alpha * beta)) return false;
	const double zeta = (beta - alpha) / (2.0 * gamma);
	const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
	cs = 1.0 / sqrt(1.0 + t * t);
	sn = cs * t;
	return true;
}
// singular values, condition number and x = V S^+ U^T b from the rotated U (columns u_j s_j) and V; sequential, K x K
__host__ __device__ inline void jacobiFinish(int K, const double* U, const double* V, double* s, const double* b, double* x, double* cond) {
	double smax = 0.0, smin = INFINITY;
	for (int j = 0; j < K; j++) {
		double nn = 0;
		for (int i = 0; i < K; i++) nn += U[i * K + j] * U[i * K + j];
		s[j] = sqrt(nn);
		smax = s[j] > smax ? s[j] : smax; smin = s[j] < smin ? s[j] : smin;
	}
	if (cond) *cond = smax / smin;
	const double tolA = (double)K * smax * 0x1.0p-52, tolB = 1.4916681462400413e-154 /* sqrt(2^-1022) */;
	const double tol = tolA > tolB ? tolA : tolB;
	for (int k = 0; k < K; k++) x[k] = 0.0;
	for (int j = 0; j < K; j++) {
		if (s[j] <= tol) continue;
		double ub = 0;                              // (u_j . b) / s_j, with u_j = U[:,j] / s_j
		for (int i = 0; i < K; i++) ub += U[i * K + j] * b[i];
		const double wgt = ub / (s[j] * s[j]);
		for (int k = 0; k < K; k++) x[k] += V[k * K + j] * wgt;
	}
}
inline double tree8(const double* t) { return ((t[0] + t[4]) + (t[2] + t[6])) + ((t[1] + t[5]) + (t[3] + t[7])); }
// host twin of the warp solver (K <= 8)
inline void jacobiPinvSolveRoundRobinHost(int K, double* U, double* V, double* s, const double* b, double* x, double* cond) {
	const int KP = (K + 1) & ~1, H = KP / 2;
	for (int i = 0; i < K * K; i++) V[i] = 0.0;
	for (int i = 0; i < K; i++) V[i * K + i] = 1.0;
	for (int sweep = 0; sweep < 60 && K > 1; sweep++) {
		bool rotated = false;
		for (int r = 0; r < KP - 1; r++) {
			double cs[4], sn[4];
			bool on[4];
			for (int m = 0; m < H; m++) {                      // all rotations of the round from the SAME U
				int p, q;
				roundRobinPair(KP, r, m, p, q);
				on[m] = false; cs[m] = 1.0; sn[m] = 0.0;
				if (q >= K) continue;
				double ta[8] = {0}, tb[8] = {0}, tg[8] = {0};
				for (int i = 0; i < K; i++) { const double up = U[i * K + p], uq = U[i * K + q]; ta[i] = up * up; tb[i] = uq * uq; tg[i] = up * uq; }
				on[m] = jacobiRotation(tree8(ta), tree8(tb), tree8(tg), cs[m], sn[m]);
			}
			for (int m = 0; m < H; m++) {
				if (!on[m]) continue;
				int p, q;
				roundRobinPair(KP, r, m, p, q);
				rotated = true;
				for (int i = 0; i < K; i++) {
					const double up = U[i * K + p], uq = U[i * K + q];
					U[i * K + p] = cs[m] * up - sn[m] * uq;
					U[i * K + q] = sn[m] * up + cs[m] * uq;
					const double vp = V[i * K + p], vq = V[i * K + q];
					V[i * K + p] = cs[m] * vp - sn[m] * vq;
					V[i * K + q] = sn[m] * vp + cs[m] * vq;
				}
			}
		}
		if (!rotated) break;
	}
	jacobiFinish(K, U, V, s, b, x, cond);
}
#ifdef __CUDACC__
// one warp; U (= A on entry), V: K x K in shared memory; s, b, x: K doubles in shared memory.  All 32 lanes must call.
__device__ inline void jacobiPinvSolveWarp(int K, double* U, double* V, double* s, const double* b, double* x, double* cond) {
	const int lane = threadIdx.x & 31, g = lane >> 3, i = lane & 7;
	const int KP = (K + 1) & ~1, H = KP / 2;
	for (int idx = lane; idx < K * K; idx += 32) V[idx] = (idx / K == idx % K) ? 1.0 : 0.0;
	__syncwarp();
	for (int sweep = 0; sweep < 60 && K > 1; sweep++) {
		bool rotated = false;
		for (int r = 0; r < KP - 1; r++) {
			int p = 0, q = 0;
			roundRobinPair(KP, r, g < H ? g : 0, p, q);
			const bool pairOk = g < H && q < K, rowOk = pairOk && i < K;
			const double up = rowOk ? U[i * K + p] : 0.0, uq = rowOk ? U[i * K + q] : 0.0;
			double alpha = up * up, beta = uq * uq, gamma = up * uq;
#pragma unroll
			for (int o = 4; o > 0; o >>= 1) {
				alpha = alpha + __shfl_xor_sync(0xffffffffu, alpha, o);
				beta = beta + __shfl_xor_sync(0xffffffffu, beta, o);
				gamma = gamma + __shfl_xor_sync(0xffffffffu, gamma, o);
			}
			double cs, sn;
			const bool on = jacobiRotation(alpha, beta, gamma, cs, sn) && pairOk;
			__syncwarp();                                      // every lane has read the round's U before anybody writes
			if (on && rowOk) {
				U[i * K + p] = cs * up - sn * uq;
				U[i * K + q] = sn * up + cs * uq;
				const double vp = V[i * K + p], vq = V[i * K + q];
				V[i * K + p] = cs * vp - sn * vq;
				V[i * K + q] = sn * vp + cs * vq;
			}
			rotated = __any_sync(0xffffffffu, on) || rotated;
			__syncwarp();
		}
		if (!rotated) break;
	}
	if (lane == 0) jacobiFinish(K, U, V, s, b, x, cond);
	__syncwarp();
}
#endif

// asynchronous 8-byte global -> shared copies (LDGSTS) of a thread into its own slots
// (shared-window addresses are computed once by the caller: __cvta_generic_to_shared inside a loop costs an S2UR per copy)
__device__ __forceinline__ void cpAsync8(unsigned smemDst, const double* gsrc) {
	asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(smemDst), "l"(gsrc) : "memory");
}
// the same, ordered after the value read from that slot has arrived
__device__ __forceinline__ void cpAsync8After(unsigned smemDst, const double* gsrc, double readBefore) {
	asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(smemDst), "l"(gsrc), "d"(readBefore) : "memory");
}
__device__ __forceinline__ double ldShared(unsigned smemSrc) {
	double v;
	asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(smemSrc) : "memory");
	return v;
}
__device__ __forceinline__ void cpAsyncCommit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cpAsyncWait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
// stages of the moments kernel's copy ring: as many (2..6) as fit 42 KB of the 48 KB of static shared memory next to the reduction scratch
#define MOMENTS_STAGES(K) (21 / ((K) + 1) < 2 ? 2 : (21 / ((K) + 1) > 6 ? 6 : 21 / ((K) + 1)))

struct BasisArgs {
	const double* ptr[8];
	double scalar[8];
};

// device-resident result of a fit (one small vector behind a handle): XtX[K*K] (stride K) | Xty[8] | x[8] | cond
static const int FIT_XTX = 0, FIT_XTY = 64, FIT_X = 72, FIT_COND = 80, FIT_DOUBLES = 96;

struct FitArgs {
	int world;                 // shards gathered in src ([world][M][2]); 1: the local moments
	double n;                  // logical number of paths (all shards)
	const double* cachedFit;   // non-null: XtX of an earlier fit of the same estimator (the reference caches its solver per instance, :125-138)
	double* fit;               // FIT_DOUBLES doubles
};

// Called by one whole CTA.  src: moments as (hi, lo) pairs in the order (p <= q pairs row by row, then y*b_p).
template <int K> __device__ void regressionFinish(const double* __restrict__ src, const BasisArgs& b, const FitArgs& f) {
	constexpr int M = K * (K + 1) / 2 + K;
	__shared__ double mean[M];
	if (threadIdx.x < M) {
		const int m = threadIdx.x;
		dd t = { __ldcg(src + 2 * m), __ldcg(src + 2 * m + 1) };
		for (int r = 1; r < f.world; r++) { const dd o = { __ldcg(src + ((size_t)r * M + m) * 2), __ldcg(src + ((size_t)r * M + m) * 2 + 1) }; ddMerge(t, o); }
		mean[m] = (t.hi + t.lo) / f.n;
	}
	__shared__ double U[K * K], V[K * K], sv[K], rhs[K], xs[K], condS;
	__syncthreads();
	if (threadIdx.x == 0) {
		int m = 0;
		for (int p = 0; p < K; p++) for (int q = p; q < K; q++) {
			double v = mean[m++];
			// deterministic x deterministic: the reference's mult() stays a scalar and its average is the product itself
			if (b.ptr[p] == nullptr && b.ptr[q] == nullptr) v = b.scalar[p] * b.scalar[q];
			if (f.cachedFit) v = f.cachedFit[FIT_XTX + p * K + q];
			U[p * K + q] = U[q * K + p] = v;
			f.fit[FIT_XTX + p * K + q] = f.fit[FIT_XTX + q * K + p] = v;
		}
		for (int p = 0; p < K; p++) { rhs[p] = mean[m++]; f.fit[FIT_XTY + p] = rhs[p]; }
	}
	__syncthreads();
	if (threadIdx.x < 32) {                            // one warp solves (round-robin Jacobi, lanes = (pair, row))
		jacobiPinvSolveWarp(K, U, V, sv, rhs, xs, &condS);
		if (threadIdx.x < K) f.fit[FIT_X + threadIdx.x] = xs[threadIdx.x];
		if (threadIdx.x == 0) f.fit[FIT_COND] = condS;
	}
}

// All K(K+1)/2 + K sums in ONE pass over the paths (the reference makes 27 passes for K = 6, MonteCarloConditionalExpectationRegression
// .java:128-144).  Reads 8 B per stochastic basis function + 8 B for y per path.  The last CTA merges the per-CTA partials and
//   FINISH 0: writes the local moments ((hi, lo) pairs) to momOut (mapped host memory, or the communicator's send buffer),
//   FINISH 1: (single GPU) goes straight on to the solve: no second launch, nothing returns to the host.
template <int K, int FINISH> __global__ void __launch_bounds__(RED_THREADS) momentsKernel(BasisArgs b, const double* __restrict__ y, uint64_t n,
		double* __restrict__ partials /* [grid][M][2] */, unsigned int* ticket, double* __restrict__ momOut, FitArgs f, const PeerArgs px) {
	constexpr int M = K * (K + 1) / 2 + K;
	dd acc[M];
#pragma unroll
	for (int m = 0; m < M; m++) { acc[m].hi = 0.0; acc[m].lo = 0.0; }
	const uint64_t stride = (uint64_t)gridDim.x * RED_THREADS;
	uint64_t i = blockIdx.x * (uint64_t)RED_THREADS + threadIdx.x;
	// The kernel runs at one CTA per SM (27 double-double accumulators per thread for K = 6), so other warps do not hide memory latency, and
	// a prefetch held in registers cannot go deeper than one element: every thread keeps its next D elements on the way as asynchronous
	// 8-byte copies into its OWN shared-memory slots (no other thread reads them: the only synchronisation is cp.async.wait_group),
	// D x (K+1) x 2 KB per CTA.  Against the one-element register prefetch (K = 6): 238 -> 174 registers, 66 -> 64.5 us per regression at
	// 1 M paths (L2-resident inputs), 304 -> 228 us at 8 M (inputs from HBM).
	constexpr int D = MOMENTS_STAGES(K);
	__shared__ double stage[D][K + 1][RED_THREADS];
	const int tid = threadIdx.x;
	const unsigned slot0 = (unsigned)__cvta_generic_to_shared(&stage[0][0][tid]);
	constexpr unsigned ROW = RED_THREADS * sizeof(double), STAGE = (K + 1) * ROW;
#pragma unroll
	for (int d = 0; d < D; d++) {
		const uint64_t idx = i + d * stride;
#pragma unroll
		for (int k = 0; k < K; k++) if (!b.ptr[k]) stage[d][k][tid] = b.scalar[k];      // a deterministic basis function: its slots hold the constant
		if (idx < n) {
#pragma unroll
			for (int k = 0; k < K; k++) if (b.ptr[k]) cpAsync8(slot0 + d * STAGE + k * ROW, b.ptr[k] + idx);
			cpAsync8(slot0 + d * STAGE + K * ROW, y + idx);
		}
		cpAsyncCommit();
	}
	int s = 0;
	for (; i < n; i += stride) {
		cpAsyncWait<D - 1>();                              // the oldest of the D groups in flight: this element
		const unsigned cur = slot0 + s * STAGE;
		double v[K];
#pragma unroll
		for (int k = 0; k < K; k++) v[k] = ldShared(cur + k * ROW);
		const double yy = ldShared(cur + K * ROW);
		{
			// refill the slot just read (the values are operands of the asm, so the reads have completed before the copies are issued)
			const uint64_t idx = i + D * stride;
			if (idx < n) {
#pragma unroll
				for (int k = 0; k < K; k++) if (b.ptr[k]) cpAsync8After(cur + k * ROW, b.ptr[k] + idx, v[k]);
				cpAsync8After(cur + K * ROW, y + idx, yy);
			}
			cpAsyncCommit();
		}
		s = (s + 1 == D) ? 0 : s + 1;
		int m = 0;
#pragma unroll
		for (int p = 0; p < K; p++) {
#pragma unroll
			for (int q = p; q < K; q++) { ddAdd(acc[m], v[p] * v[q]); m++; }
		}
#pragma unroll
		for (int p = 0; p < K; p++) { ddAdd(acc[m], yy * v[p]); m++; }
	}
	cpAsyncWait<0>();
	__shared__ dd sh[RED_THREADS / 32][M];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	DdButterfly<M, 16>::run(acc, lane);
#pragma unroll
	for (int j = 0; j < DdButterflyLeft<M>::value; j++) {
		bool ok;
		const int m = DdButterfly<M, 16>::origin(j, lane, ok);
		if (ok) sh[warp][m] = acc[j];
	}
	__syncthreads();
	if (threadIdx.x < M) {
		dd t = sh[0][threadIdx.x];
		for (int k = 1; k < RED_THREADS / 32; k++) ddMerge(t, sh[k][threadIdx.x]);
		partials[((size_t)blockIdx.x * M + threadIdx.x) * 2] = t.hi;
		partials[((size_t)blockIdx.x * M + threadIdx.x) * 2 + 1] = t.lo;
	}
	if (!drawLastTicket(ticket)) return;
	// merge over the CTAs: S lanes per moment take the CTAs s, s+S, ... in order, then a shuffle tree over the S lanes (fixed order)
	constexpr int S = (M * 8 <= RED_THREADS) ? 8 : ((M * 4 <= RED_THREADS) ? 4 : 2);
	const int m = threadIdx.x / S, sl = threadIdx.x % S;
	dd t = {0.0, 0.0};
	if (m < M) {
		// four partials are fetched (L2 latency each) before they are merged in order
		for (unsigned int blk = sl; blk < gridDim.x; blk += 4 * S) {
			dd o[4];
#pragma unroll
			for (int u = 0; u < 4; u++) {
				const unsigned int bb = blk + u * S;
				const bool ok = bb < gridDim.x;
				o[u].hi = ok ? __ldcg(partials + ((size_t)bb * M + m) * 2) : 0.0;
				o[u].lo = ok ? __ldcg(partials + ((size_t)bb * M + m) * 2 + 1) : 0.0;
			}
#pragma unroll
			for (int u = 0; u < 4; u++) if (blk + u * S < gridDim.x) ddMerge(t, o[u]);
		}
	}
#pragma unroll
	for (int d = S / 2; d > 0; d >>= 1) {
		dd o;
		o.hi = __shfl_down_sync(0xffffffffu, t.hi, d, S);
		o.lo = __shfl_down_sync(0xffffffffu, t.lo, d, S);
		ddMerge(t, o);
	}
	if (m < M && sl == 0) { momOut[2 * m] = t.hi; momOut[2 * m + 1] = t.lo; }
	if (FINISH == 1) {
		__threadfence_block();
		__syncthreads();
		regressionFinish<K>(momOut, b, f);
	}
	if (FINISH == 2) {
		// sharded: the local moments go to all ranks through peer memory, then every rank merges the shards in rank order and solves
		__threadfence_block();
		__syncthreads();
		peerExchangeBlock(px, momOut, 2 * M);
		regressionFinish<K>(peerGathered(px), b, f);
	}
}

// after the all-gather of the shards' moments: merge in rank order + solve (one CTA)
template <int K> __global__ void __launch_bounds__(64) regressionSolveKernel(const double* __restrict__ gathered, BasisArgs b, FitArgs f) {
	regressionFinish<K>(gathered, b, f);
}

// conditional expectation: b_0*x_0, then + b_i*x_i in order (…Regression.java:103-107); 8 B per stochastic basis read, 8 B written.
// The coefficients come from the device-resident fit (uniform loads): the host never sees them on this path.
template <int K> __global__ void __launch_bounds__(256) predictKernelV(BasisArgs b, const double* __restrict__ coef, double* __restrict__ out, uint64_t n) {
	double c[K];
#pragma unroll
	for (int k = 0; k < K; k++) c[k] = __ldcg(coef + k);
	const uint64_t stride = (uint64_t)gridDim.x * 256;
	for (uint64_t i = blockIdx.x * (uint64_t)256 + threadIdx.x; i < n; i += stride) {
		double ce = (b.ptr[0] ? b.ptr[0][i] : b.scalar[0]) * c[0];
#pragma unroll
		for (int k = 1; k < K; k++) ce = ce + (b.ptr[k] ? b.ptr[k][i] : b.scalar[k]) * c[k];
		out[i] = ce;
	}
}

static int reduceGrid() { return ctx().smCount * 4; }
// CTAs of the moments kernel: one per SM - its accumulators leave room for one resident CTA per SM only, so a second wave would just
// double the partials the last CTA has to merge (FMB_MOMENTS_WAVES=2: the earlier geometry, for A/B runs: 80 vs 70 us per regression at 1 M)
static int momentsGrid(uint64_t n) {
	static const int waves = (getenv("FMB_MOMENTS_WAVES") && atoi(getenv("FMB_MOMENTS_WAVES")) == 2) ? 2 : 1;
	return (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)ctx().smCount * waves, (n + RED_THREADS - 1) / RED_THREADS));
}

int commAllGather(int count);                      // fmb_comm.cu
void peerArgsNext(PeerArgs& px);                   // fmb_comm.cu

template <int K> static void launchMoments(int finish, const BasisArgs& b, const double* y, uint64_t n, double* partials, unsigned int* ticket, double* momOut,
		const FitArgs& f, int grid, cudaStream_t s, const PeerArgs& px = PeerArgs()) {
	if (finish == 2) momentsKernel<K, 2><<<grid, RED_THREADS, 0, s>>>(b, y, n, partials, ticket, momOut, f, px);
	else if (finish) momentsKernel<K, 1><<<grid, RED_THREADS, 0, s>>>(b, y, n, partials, ticket, momOut, f, px);
	else momentsKernel<K, 0><<<grid, RED_THREADS, 0, s>>>(b, y, n, partials, ticket, momOut, f, px);
}
template <int K> static void launchSolve(const double* gathered, const BasisArgs& b, const FitArgs& f, cudaStream_t s) {
	regressionSolveKernel<K><<<1, 64, 0, s>>>(gathered, b, f);
}
template <int K> static void launchPredict(const BasisArgs& b, const double* coef, double* out, uint64_t n, int grid, cudaStream_t s) {
	predictKernelV<K><<<grid, 256, 0, s>>>(b, coef, out, n);
}
#define FMB_K_SWITCH(K, CALL) \
	switch (K) { case 1: CALL(1); break; case 2: CALL(2); break; case 3: CALL(3); break; case 4: CALL(4); break; \
	             case 5: CALL(5); break; case 6: CALL(6); break; case 7: CALL(7); break; default: CALL(8); break; }

static int fillBasis(int K, const fmb_handle* basis, const double* basis_scalar, uint64_t* n, bool* any, BasisArgs* b) {
	if (K < 1 || K > 8) { setError("regression kernels support 1..8 basis functions (got %d)", K); return FMB_EUNSUPPORTED; }
	for (int k = 0; k < K; k++) {
		b->ptr[k] = nullptr;
		b->scalar[k] = basis_scalar ? basis_scalar[k] : 0.0;
		if (basis[k] == 0) continue;
		Vec* v;
		FMB_TRY(lookup(basis[k], &v));
		if (*any && v->n != *n) { setError("basis function sizes differ"); return FMB_EINVAL; }
		*n = v->n; *any = true;
		b->ptr[k] = v->ptr;
	}
	for (int k = K; k < 8; k++) { b->ptr[k] = nullptr; b->scalar[k] = 0.0; }
	return FMB_OK;
}

// enqueue: local moments (+ all-gather + merge) + solve -> fit (device).  Caller holds scratchMu.
static int enqueueFit(int K, const BasisArgs& b, const double* y, uint64_t nLocal, double nGlobal, const double* cachedFit, double* fit) {
	Context& c = ctx();
	const int M = K * (K + 1) / 2 + K;
	const int grid = momentsGrid(nLocal);
	FMB_TRY(ensureScratch(0, (size_t)(grid * M * 2 + COMM_MAX_DOUBLES) * sizeof(double)));
	double* dpart = (double*)c.scratch;
	double* momLocal = c.comm.active ? c.comm.sendBuf : dpart + (size_t)grid * M * 2;
	FitArgs f;
	f.world = 1; f.n = nGlobal; f.cachedFit = cachedFit; f.fit = fit;
	// sharded with peer exchange: ONE kernel accumulates, exchanges over NVLink, merges the shards and solves (an empty shard has no
	// accumulation kernel: it takes part through the stand-alone exchange kernel below - same protocol)
	const bool fused = c.comm.active && c.comm.peer && nLocal > 0;
	const int finish = fused ? 2 : (c.comm.active ? 0 : 1);
	PeerArgs px;
	if (fused) {
		if (*c.comm.peerErrHost) { setError("peer exchange: a rank did not arrive (timed out)"); return FMB_ECUDA; }
		peerArgsNext(px);
		f.world = c.comm.world;
		momLocal = dpart + (size_t)grid * M * 2;
	}
	if (nLocal == 0) {
		// an empty shard contributes zero moments (it still takes part in the exchange)
		FMB_CUDA(cudaMemsetAsync(momLocal, 0, (size_t)M * 2 * sizeof(double), c.stream));
		if (finish) { setError("regression on an empty vector"); return FMB_EINVAL; }
	} else {
#define CALL(KV) launchMoments<KV>(finish, b, y, nLocal, dpart, c.ticket, momLocal, f, grid, c.stream, px)
		FMB_K_SWITCH(K, CALL)
#undef CALL
		countLaunch();
		FMB_CUDA(cudaGetLastError());
	}
	if (c.comm.active && !fused) {
		FMB_TRY(commAllGather(2 * M));
		f.world = c.comm.world;
#define CALL(KV) launchSolve<KV>(c.comm.gatherBuf, b, f, c.stream)
		FMB_K_SWITCH(K, CALL)
#undef CALL
		countLaunch();
		FMB_CUDA(cudaGetLastError());
	}
	return FMB_OK;
}

static int enqueuePredict(int K, const BasisArgs& b, const double* coef, uint64_t n, fmb_handle* out) {
	Context& c = ctx();
	double* dst;
	FMB_TRY(newVec(n, out, &dst));
	if (n == 0) return FMB_OK;
	const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)c.smCount * 8, (n + 255) / 256));
#define CALL(KV) launchPredict<KV>(b, coef, dst, n, grid, c.stream)
	FMB_K_SWITCH(K, CALL)
#undef CALL
	countLaunch();
	FMB_CUDA(cudaGetLastError());
	return FMB_OK;
}

} // namespace fmb

using namespace fmb;

extern "C" {

// With a communicator (fmb_comm_init) the result covers ALL shards: local kernel -> all-gather on the compute stream -> rank-ordered
// merge kernel -> mapped host memory; without one it is the local result, written by the finalising CTA straight to mapped host memory.
int fmb_rv_reduce(int op, fmb_handle x, fmb_handle w, double a, double* out2) {
	FMB_TRY(requireInit());
	PinScope pins;
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
	const bool isMinMax = op >= FMB_R_MIN;
	out2[0] = 0.0; out2[1] = 0.0;
	if (n == 0 && !c.comm.active) { out2[0] = isMinMax ? NAN : 0.0; return FMB_OK; }
	const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)reduceGrid(), (n + RED_THREADS - 1) / RED_THREADS));
	std::lock_guard<std::mutex> lk(c.scratchMu);
	FMB_TRY(ensureScratch(0, (size_t)grid * 2 * sizeof(double)));
	double* dpart = (double*)c.scratch;
	// sharded with peer exchange: the finalising CTA exchanges over NVLink, merges the shards and writes the result (one launch); an
	// empty shard has no reduction kernel and takes part through the stand-alone exchange kernel (same protocol)
	const bool fused = c.comm.active && c.comm.peer && n > 0;
	PeerArgs px;
	if (fused) {
		if (*c.comm.peerErrHost) { setError("peer exchange: a rank did not arrive (timed out)"); return FMB_ECUDA; }
		peerArgsNext(px);
	}
	double* dst = (c.comm.active && !fused) ? c.comm.sendBuf : c.hostResultDev;
	if (n == 0) {
		FMB_CUDA(cudaMemsetAsync(dst, 0, 2 * sizeof(double), c.stream));          // (0, 0): a zero sum / "no data" for min and max
	} else {
		switch (op) {
		case FMB_R_SUM: sumKernel<FMB_R_SUM><<<grid, RED_THREADS, 0, c.stream>>>(vx->ptr, wp, a, n, dpart, c.ticket, dst, px); break;
		case FMB_R_SUM_PRODUCT: sumKernel<FMB_R_SUM_PRODUCT><<<grid, RED_THREADS, 0, c.stream>>>(vx->ptr, wp, a, n, dpart, c.ticket, dst, px); break;
		case FMB_R_CENTERED_M2: sumKernel<FMB_R_CENTERED_M2><<<grid, RED_THREADS, 0, c.stream>>>(vx->ptr, wp, a, n, dpart, c.ticket, dst, px); break;
		case FMB_R_CENTERED_M2_W: sumKernel<FMB_R_CENTERED_M2_W><<<grid, RED_THREADS, 0, c.stream>>>(vx->ptr, wp, a, n, dpart, c.ticket, dst, px); break;
		case FMB_R_MIN: minMaxKernel<false><<<grid, RED_THREADS, 0, c.stream>>>(vx->ptr, n, dpart, c.ticket, dst, px); break;
		case FMB_R_MAX: minMaxKernel<true><<<grid, RED_THREADS, 0, c.stream>>>(vx->ptr, n, dpart, c.ticket, dst, px); break;
		}
		countLaunch();
		FMB_CUDA(cudaGetLastError());
	}
	if (c.comm.active && !fused) {
		FMB_TRY(commAllGather(2));
		mergeShardsKernel<<<1, 32, 0, c.stream>>>(c.comm.gatherBuf, c.comm.world, 2, isMinMax ? (op == FMB_R_MIN ? 1 : 2) : 0, c.hostResultDev);
		countLaunch();
		FMB_CUDA(cudaGetLastError());
	}
	FMB_CUDA(cudaStreamSynchronize(c.stream));
	out2[0] = c.hostResult[0];
	out2[1] = isMinMax ? 0.0 : c.hostResult[1];
	return FMB_OK;
}

// sums of `count` (<= 64) vectors of one length in one launch: out2[2i], out2[2i+1] = (hi, lo) of sum_p f(x_i[p]); with a communicator over
// all shards.  FMB_RM_SUM: f(x) = x; FMB_RM_SUM_INVERT_MULT: f(x) = (1 / x) * a  (RandomVariable.invert().mult(a), summed).
int fmb_rv_reduce_many(int op, int count, const fmb_handle* x, double a, double* out2) {
	FMB_TRY(requireInit());
	PinScope pins;
	if (!x || !out2 || count < 1 || count > REDUCE_MANY_MAX || (op != FMB_RM_SUM && op != FMB_RM_SUM_INVERT_MULT)) { setError("reduce_many: bad argument"); return FMB_EINVAL; }
	Context& c = ctx();
	ManyArgs v;
	uint64_t n = 0;
	for (int i = 0; i < REDUCE_MANY_MAX; i++) v.x[i] = nullptr;
	for (int i = 0; i < count; i++) {
		Vec* vx;
		if (x[i] == 0) { setError("reduce_many: vector %d is not a device vector", i); return FMB_EINVAL; }
		FMB_TRY(lookup(x[i], &vx));
		if (i > 0 && vx->n != n) { setError("operand sizes differ (%llu vs %llu)", (unsigned long long)vx->n, (unsigned long long)n); return FMB_EINVAL; }
		n = vx->n;
		v.x[i] = vx->ptr;
	}
	for (int i = 0; i < 2 * count; i++) out2[i] = 0.0;
	if (n == 0 && !c.comm.active) return FMB_OK;
	// CTAs per vector: about eight elements per thread at least, and about eight CTAs per SM over all vectors (a few long vectors get
	// many CTAs each, many vectors few)
	const uint64_t gxCap = std::max<uint64_t>(32, ((uint64_t)c.smCount * 8 + count - 1) / count);
	const int gx = (int)std::max<uint64_t>(1, std::min<uint64_t>(gxCap, (n + (uint64_t)RED_THREADS * 8 - 1) / ((uint64_t)RED_THREADS * 8)));
	std::lock_guard<std::mutex> lk(c.scratchMu);
	FMB_TRY(ensureScratch(0, ((size_t)count * gx * 2 + 2 * (size_t)count) * sizeof(double)));
	double* dpart = (double*)c.scratch;
	double* results = dpart + (size_t)count * gx * 2;
	const bool fused = c.comm.active && c.comm.peer && n > 0;
	PeerArgs px;
	if (fused) {
		if (*c.comm.peerErrHost) { setError("peer exchange: a rank did not arrive (timed out)"); return FMB_ECUDA; }
		peerArgsNext(px);
	}
	double* dst = (c.comm.active && !fused) ? c.comm.sendBuf : c.hostResultDev;
	if (n == 0) {
		FMB_CUDA(cudaMemsetAsync(dst, 0, 2 * (size_t)count * sizeof(double), c.stream));
	} else {
		const dim3 grid(gx, count);
		if (op == FMB_RM_SUM) sumManyKernel<FMB_RM_SUM><<<grid, RED_THREADS, 0, c.stream>>>(v, count, a, n, dpart, c.ticketMany, results, dst, px);
		else sumManyKernel<FMB_RM_SUM_INVERT_MULT><<<grid, RED_THREADS, 0, c.stream>>>(v, count, a, n, dpart, c.ticketMany, results, dst, px);
		countLaunch();
		FMB_CUDA(cudaGetLastError());
	}
	if (c.comm.active && !fused) {
		FMB_TRY(commAllGather(2 * count));
		mergeShardsKernel<<<1, REDUCE_MANY_MAX, 0, c.stream>>>(c.comm.gatherBuf, c.comm.world, 2 * count, 0, c.hostResultDev);
		countLaunch();
		FMB_CUDA(cudaGetLastError());
	}
	FMB_CUDA(cudaStreamSynchronize(c.stream));
	for (int i = 0; i < 2 * count; i++) out2[i] = c.hostResult[i];
	return FMB_OK;
}

// LOCAL moment sums as double-double pairs on the host (the caller exchanges them between shards itself and divides by n): the
// host-resident variant of the regression, kept for hosts that exchange partials outside the library.
int fmb_regression_moments(int K, const fmb_handle* basis, const double* basis_scalar, fmb_handle y,
                           double* XtX_hi, double* XtX_lo, double* Xty_hi, double* Xty_lo) {
	FMB_TRY(requireInit());
	PinScope pins;
	if (!basis || !XtX_hi || !XtX_lo || !Xty_hi || !Xty_lo) return FMB_EINVAL;
	Context& c = ctx();
	BasisArgs b;
	Vec* vy;
	FMB_TRY(lookup(y, &vy));
	uint64_t n = vy->n;
	bool any = false;
	FMB_TRY(fillBasis(K, basis, basis_scalar, &n, &any, &b));
	if (n != vy->n) { setError("basis functions and dependents differ in size"); return FMB_EINVAL; }
	const int M = K * (K + 1) / 2 + K;
	std::lock_guard<std::mutex> lk(c.scratchMu);
	if (n == 0) {
		for (int i = 0; i < 2 * M; i++) c.hostResult[i] = 0.0;
	} else {
		const int grid = momentsGrid(n);                        // (the same geometry as the device-resident fit: identical moments, bit for bit)
		FMB_TRY(ensureScratch(0, (size_t)grid * M * 2 * sizeof(double)));
		FitArgs f = { 1, 0.0, nullptr, nullptr };
#define CALL(KV) launchMoments<KV>(0, b, vy->ptr, n, (double*)c.scratch, c.ticket, c.hostResultDev, f, grid, c.stream)
		FMB_K_SWITCH(K, CALL)
#undef CALL
		countLaunch();
		FMB_CUDA(cudaGetLastError());
		FMB_CUDA(cudaStreamSynchronize(c.stream));
	}
	const double* tot = c.hostResult;
	int m = 0;
	for (int p = 0; p < K; p++) for (int q = p; q < K; q++) {
		XtX_hi[p * K + q] = XtX_hi[q * K + p] = tot[2 * m];
		XtX_lo[p * K + q] = XtX_lo[q * K + p] = tot[2 * m + 1];
		m++;
	}
	for (int p = 0; p < K; p++) { Xty_hi[p] = tot[2 * m]; Xty_lo[p] = tot[2 * m + 1]; m++; }
	return FMB_OK;
}

int fmb_regression_solve_svd(int K, const double* A, const double* b, double* x, double* cond) {
	if (K < 1 || !A || !b || !x) { setError("solve_svd: bad argument"); return FMB_EINVAL; }
	std::vector<double> U(A, A + (size_t)K * K), V((size_t)K * K, 0.0), s(K);
	double condLocal = 0.0;
	double* cp = cond ? cond : &condLocal;
	if (K <= 8) jacobiPinvSolveRoundRobinHost(K, U.data(), V.data(), s.data(), b, x, cp);    // the device-resident regression's algorithm: identical coefficients
	else jacobiPinvSolve(K, U.data(), V.data(), s.data(), b, x, cp);
	return FMB_OK;
}

int fmb_regression_predict(int K, const fmb_handle* basis, const double* basis_scalar, const double* x, fmb_handle* out) {
	FMB_TRY(requireInit());
	PinScope pins;
	if (!basis || !x || !out) return FMB_EINVAL;
	Context& c = ctx();
	BasisArgs b;
	uint64_t n = 0;
	bool any = false;
	FMB_TRY(fillBasis(K, basis, basis_scalar, &n, &any, &b));
	if (!any) { setError("predict: all basis functions deterministic; stays on the host"); return FMB_EINVAL; }
	// the coefficients travel through a small pool block (stream-ordered: the block is reused only after the kernel that reads it)
	void* coef;
	FMB_TRY(poolAlloc(8 * sizeof(double), &coef));
	std::lock_guard<std::mutex> lk(c.scratchMu);
	int rc = ensureScratch(8 * sizeof(double), 0);
	if (rc == FMB_OK) {
		double* h = (double*)c.pinned;
		for (int k = 0; k < 8; k++) h[k] = k < K ? x[k] : 0.0;
		cudaError_t e = cudaMemcpyAsync(coef, h, 8 * sizeof(double), cudaMemcpyHostToDevice, c.stream);
		if (e == cudaSuccess) e = cudaStreamSynchronize(c.stream);          // the pinned staging buffer is shared
		if (e != cudaSuccess) { setError("predict: %s", cudaGetErrorString(e)); rc = FMB_ECUDA; }
	}
	if (rc == FMB_OK) rc = enqueuePredict(K, b, (const double*)coef, n, out);
	poolFree(coef, 8 * sizeof(double));
	return rc;
}

// ---- device-resident regression (no host round trip): MonteCarloConditionalExpectationRegression.java:97-150 -----------------
int fmb_regression_fit(int K, const fmb_handle* basis, const double* basis_scalar, fmb_handle y, uint64_t n_global, fmb_handle cached_fit,
                       fmb_handle* fit) {
	FMB_TRY(requireInit());
	PinScope pins;
	if (!basis || !fit) return FMB_EINVAL;
	Context& c = ctx();
	BasisArgs b;
	Vec* vy;
	FMB_TRY(lookup(y, &vy));
	uint64_t n = vy->n;
	bool any = false;
	FMB_TRY(fillBasis(K, basis, basis_scalar, &n, &any, &b));
	if (n != vy->n) { setError("basis functions and dependents differ in size"); return FMB_EINVAL; }
	if (n_global == 0) n_global = n;
	const double* cachedPtr = nullptr;
	if (cached_fit) FMB_TRY(lookupPtr(cached_fit, FIT_DOUBLES, &cachedPtr));
	double* fitPtr;
	FMB_TRY(newVec(FIT_DOUBLES, fit, &fitPtr));
	std::lock_guard<std::mutex> lk(c.scratchMu);
	const int rc = enqueueFit(K, b, vy->ptr, n, (double)n_global, cachedPtr, fitPtr);
	if (rc != FMB_OK) { releaseRef(*fit); *fit = 0; }
	return rc;
}

int fmb_regression_fit_get(fmb_handle fit, int K, double* XtX, double* Xty, double* x, double* cond) {
	FMB_TRY(requireInit());
	PinScope pins;
	if (K < 1 || K > 8) return FMB_EINVAL;
	const double* p;
	FMB_TRY(lookupPtr(fit, FIT_DOUBLES, &p));
	double h[FIT_DOUBLES];
	FMB_CUDA(cudaMemcpyAsync(h, p, sizeof(h), cudaMemcpyDeviceToHost, ctx().stream));
	FMB_CUDA(cudaStreamSynchronize(ctx().stream));
	if (XtX) for (int i = 0; i < K * K; i++) XtX[i] = h[FIT_XTX + i];
	if (Xty) for (int i = 0; i < K; i++) Xty[i] = h[FIT_XTY + i];
	if (x) for (int i = 0; i < K; i++) x[i] = h[FIT_X + i];
	if (cond) *cond = h[FIT_COND];
	return FMB_OK;
}

int fmb_regression_predict_fit(int K, const fmb_handle* basis, const double* basis_scalar, fmb_handle fit, fmb_handle* out) {
	FMB_TRY(requireInit());
	PinScope pins;
	if (!basis || !out) return FMB_EINVAL;
	BasisArgs b;
	uint64_t n = 0;
	bool any = false;
	FMB_TRY(fillBasis(K, basis, basis_scalar, &n, &any, &b));
	if (!any) { setError("predict: all basis functions deterministic; stays on the host"); return FMB_EINVAL; }
	const double* p;
	FMB_TRY(lookupPtr(fit, FIT_DOUBLES, &p));
	return enqueuePredict(K, b, p + FIT_X, n, out);
}

int fmb_regression_conditional_expectation(int K, const fmb_handle* basis, const double* basis_scalar, fmb_handle y, uint64_t n_global,
                                           fmb_handle cached_fit, int Kp, const fmb_handle* basis_pred, const double* basis_pred_scalar,
                                           fmb_handle* fit, fmb_handle* out) {
	if (!fit || !out) return FMB_EINVAL;
	if (Kp != K) { setError("conditional_expectation: %d estimator and %d predictor basis functions", K, Kp); return FMB_EINVAL; }
	FMB_TRY(fmb_regression_fit(K, basis, basis_scalar, y, n_global, cached_fit, fit));
	const int rc = fmb_regression_predict_fit(Kp, basis_pred ? basis_pred : basis, basis_pred ? basis_pred_scalar : basis_scalar, *fit, out);
	if (rc != FMB_OK) { releaseRef(*fit); *fit = 0; }
	return rc;
}

} // extern "C"
