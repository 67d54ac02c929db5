// Fused Euler / log-Euler path evolution: one kernel per simulation instead of dozens of array passes per time step.
//
// Replaces EulerSchemeFromProcessModel.doPrecalculateProcess (J/montecarlo/process/EulerSchemeFromProcessModel.java:170-326)
// for the four models of the path.  One thread owns one path for the whole time loop (lanes = consecutive paths), so
//   * every load of dW[t][f][path] and every store of X[t+1][component][path] is coalesced (getProcessValue(t, c) stays
//     a contiguous device vector), and
//   * the per-path arithmetic runs in exactly the reference's order (j ascending, k ascending, no FMA contraction in
//     STRICT mode: this file is compiled with -fmad=false).
// The kernels are FP64-pipe bound (double exp/log per component-step), not HBM bound; see DESIGN.md for the numbers.
//
// Algorithmic HBM bytes per path-step: 8*F read (increments) + 8*live(t) written (process values).
#include "fmb_common.cuh"
#include "fmb_math.cuh"
#include <cmath>
#include <algorithm>

namespace fmb {

__device__ __forceinline__ double jminE(double a, double b) {
	if (a != a) return a;
	if (a == 0.0 && b == 0.0 && signbit(b)) return b;
	return (a <= b) ? a : b;
}
__device__ __forceinline__ double jmaxE(double a, double b) {
	if (a != a) return a;
	if (a == 0.0 && b == 0.0 && signbit(a)) return b;
	return (a >= b) ? a : b;
}

#ifndef FMB_LMM_PREFETCH
#define FMB_LMM_PREFETCH 0   // 1: fetch the Brownian increments of step t+1 while step t computes (measured 1 % slower: six more live registers)
#endif
#ifndef FMB_LMM_U
#define FMB_LMM_U 2          // live rates processed together per thread (ILP); 2 measured best on B200 (profiles/r01_notes.md)
#endif

#ifndef FMB_LMM_MINB
#define FMB_LMM_MINB 5
#endif

enum { SCHEME_EULER = 0, SCHEME_PC = 1, SCHEME_EULER_FUNCTIONAL = 2, SCHEME_PC_FUNCTIONAL = 3 };

// a*b + c: STRICT = rounded product then rounded sum (the JVM never contracts); FAST (fmb_set_fp_mode(1)) = one fused multiply-add.
template <bool FAST> __device__ __forceinline__ double mad(double a, double b, double c) { return FAST ? fma(a, b, c) : a * b + c; }

// ---------------------------------------------------------------------------------------------------------------
// Black-Scholes: BlackScholesModel.java:60-139.  Y += (r - sigma^2/2) dt + sigma dW ; X = exp(Y); functional schemes
// re-apply log every step (:235-237 of the Euler scheme).  The drift does not depend on the state, so the corrector adds
// ((mu - mu)/2)*dt = +0.0 and PREDICTOR_CORRECTOR reproduces EULER bit for bit.
// ---------------------------------------------------------------------------------------------------------------
template <bool FAST> __global__ void __launch_bounds__(256) eulerBlackScholesKernel(int functional, int T, int F, uint64_t P, const double* __restrict__ dt,
		const double* const* __restrict__ dW, double* const* __restrict__ X, double x0, double y0, double ylog0, double drift, double sigma) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t p = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; p < P; p += stride) {
		double x = x0, y = functional ? ylog0 : y0;
		for (int t = 0; t < T; t++) {
			if (functional && !FAST && t > 0) y = flog(x);          // FAST: log(exp(y)) == y up to one rounding, the state is carried
			const double w = dW[(size_t)t * F][p];
			y = mad<FAST>(drift, dt[t], y);
			y = mad<FAST>(w, sigma, y);
			x = fexp(y);
			X[t + 1][p] = x;
		}
	}
}

// ---------------------------------------------------------------------------------------------------------------
// Heston: HestonModel.java:336-420 (full truncation / reflection).  Components (asset, variance), two factors.
// ---------------------------------------------------------------------------------------------------------------
struct HestonParams { double x0, v0, ylog0, theta, kappa, xi, rho, rhoBar; int hestonScheme, scheme; };

__device__ __forceinline__ void hestonDrift(const HestonParams& h, double v, double r, double& var, double& mu0, double& mu1) {
	var = h.hestonScheme == 1 ? jmaxE(v, 0.0) : fabs(v);
	mu0 = r - var / 2.0;
	mu1 = (h.theta - var) * h.kappa;
}

__global__ void __launch_bounds__(256) eulerHestonKernel(HestonParams h, int T, uint64_t P, const double* __restrict__ dt,
		const double* __restrict__ rate, const double* const* __restrict__ dW, double* const* __restrict__ X) {
	const bool functional = (h.scheme == SCHEME_EULER_FUNCTIONAL || h.scheme == SCHEME_PC_FUNCTIONAL);
	const bool pc = (h.scheme == SCHEME_PC || h.scheme == SCHEME_PC_FUNCTIONAL);
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t p = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; p < P; p += stride) {
		double s = h.x0, v = h.v0, y0 = h.ylog0, y1 = h.v0;
		for (int t = 0; t < T; t++) {
			const double w0 = dW[2 * (size_t)t][p], w1 = dW[2 * (size_t)t + 1][p];
			const double d = dt[t], r = rate[t];
			double var, mu0, mu1;
			hestonDrift(h, v, r, var, mu0, mu1);
			const double vol = sqrt(var);
			if (functional) { y0 = (t == 0) ? h.ylog0 : flog(s); y1 = v; }
			y0 = y0 + mu0 * d;
			y0 = y0 + vol * w0;
			y0 = y0 + w1 * 0.0;
			const double volv = vol * h.xi;
			y1 = y1 + mu1 * d;
			y1 = y1 + (volv * h.rho) * w0;
			y1 = y1 + (volv * h.rhoBar) * w1;
			s = fexp(y0);
			v = y1;
			if (pc) {
				double varP, mu0P, mu1P;
				hestonDrift(h, v, r, varP, mu0P, mu1P);
				y0 = y0 + ((mu0P - mu0) / 2.0) * d;
				y1 = y1 + ((mu1P - mu1) / 2.0) * d;
				s = fexp(y0);
				v = y1;
			}
			X[2 * (size_t)(t + 1)][p] = s;
			X[2 * (size_t)(t + 1) + 1][p] = v;
		}
	}
}

// ---------------------------------------------------------------------------------------------------------------
// Hull-White: HullWhiteModel.java:367-424.  Identity state-space transform; per-step deterministic coefficients.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) eulerHullWhiteKernel(int T, uint64_t P, const double* __restrict__ dt, const double* __restrict__ c0,
		const double* __restrict__ c1, const double* __restrict__ fl, const double* const* __restrict__ dW, double* const* __restrict__ X) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t p = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; p < P; p += stride) {
		double x0 = 0.0, x1 = 0.0;
		for (int t = 0; t < T; t++) {
			const double w0 = dW[2 * (size_t)t][p], w1 = dW[2 * (size_t)t + 1][p];
			const double d = dt[t];
			const double mu0 = x0 * c0[t], mu1 = x0 * c1[t];
			double y0 = x0 + mu0 * d;
			y0 = y0 + w0 * fl[4 * t + 0];
			y0 = y0 + w1 * fl[4 * t + 1];
			double y1 = x1 + mu1 * d;
			y1 = y1 + w0 * fl[4 * t + 2];
			y1 = y1 + w1 * fl[4 * t + 3];
			x0 = y0; x1 = y1;
			X[2 * (size_t)(t + 1)][p] = x0;
			X[2 * (size_t)(t + 1) + 1][p] = x1;
		}
	}
}

// ---------------------------------------------------------------------------------------------------------------
// LIBOR market model: LIBORMarketModelFromCovarianceModel.java:1124-1223 inside the Euler scheme.
//   drift (spot):   a_j = 1/(L_j*(d/d) + 1/d) [*L_j if lognormal];  S_k += a_j*fl_jk;  mu_j = sum_k S_k*fl_jk  (+ -0.5*var_j)
//   drift (terminal): the mirrored suffix sum with a_j = 1/(L_j*(d/(-d)) + 1/(-d)), mu_j formed BEFORE S is updated.
// State: the current forward rates of the thread's path live in shared memory as Lsh[j][thread] (conflict-free, j is a
// run-time loop so the drift's prefix sum over the rate index runs sequentially per lane in the reference's order).
// Y (non-functional schemes) and mu (predictor-corrector) live in a block-private, L2-resident global scratch.
// ---------------------------------------------------------------------------------------------------------------
struct LmmParams {
	int scheme, measure, hasCap, capFix;
	double cap, logCap;
	int T, N, F, recStride;
	const double* dt;        // [T]
	const int* firstLive;    // [T]
	const double* rec;       // [T][N][recStride]: invv, hv, (bits of) X[t+1][j] row pointer, fl[0..F), pad
	const double* x0;        // [N]  X_j(0) (host libm)
	const double* y0;        // [N]  Y_j(0)
	const double* ylog0;     // [N]  inverse transform of X_j(0) (host libm), used at the first step of functional schemes
	unsigned long long* tileCounter;   // zero at launch: next 32-path tile to hand to a warp
};

// One (t, j) record is read by every thread of every block in the same order: 16-byte uniform loads, L1-resident.
template <int FT> struct LmmRec {
	double invv, hv;
	double* xrow;
	double fl[FT > 0 ? FT : 16];
	// layout: invv, hv, row pointer, fl[0..F), padded to an even number of doubles (F = 3: 48 bytes, three 16-byte loads)
	__device__ __forceinline__ void load(const double* __restrict__ r, int F) {
		const double2 a = __ldg(reinterpret_cast<const double2*>(r));
		invv = a.x; hv = a.y;
		if (FT > 0) {
			double v[(FT + 2) & ~1];
#pragma unroll
			for (int k = 0; k < ((FT + 2) & ~1); k += 2) {
				const double2 c = __ldg(reinterpret_cast<const double2*>(r) + 1 + k / 2);
				v[k] = c.x; v[k + 1] = c.y;
			}
			xrow = reinterpret_cast<double*>(__double_as_longlong(v[0]));
#pragma unroll
			for (int k = 0; k < FT; k++) fl[k] = v[1 + k];
		} else {
			xrow = reinterpret_cast<double*>(__double_as_longlong(__ldg(r + 2)));
			for (int k = 0; k < F; k++) fl[k] = __ldg(r + 3 + k);
		}
	}
};

// U consecutive live rates of one path at once (i = position in processing order; j = first+i for the spot measure,
// N-1-i for the terminal measure).  Per rate the operations and their order are exactly those of the scalar recipe; the
// only cross-rate dependency is the running factor sum S, so the U log / exp / reciprocal chains overlap (ILP U).
template <int FT, bool LOGN, int MODE, bool SPOT, int U, bool CORRECTOR, bool PARTIAL, bool FAST, bool FIRST>
__device__ __forceinline__ void lmmChunk(const LmmParams& q, const double* __restrict__ rec0, int recStep, int i0, int jBeg, int colStep, int F,
		bool functional, double d, const double* w, double* S, double* L0, double* Y0, size_t mOff, uint64_t pOff, int cnt,
		const double* __restrict__ logTab) {
	// rec0 / L0 / Y0 point at the chunk's first rate (record, shared-memory state, scratch column; the predictor drift column is Y0 + mOff);
	// recStep / colStep move them to the next rate in processing order (the caller advances them chunk by chunk, so no index
	// multiplications are left in the loop).  i0 = position of the first rate in processing order (j = jBeg +- i).
	// PARTIAL: only the first cnt (< U) rates are real; the others recompute rate cnt-1 and are masked out of S and of every store,
	// so a short remainder costs one chunk latency instead of cnt sequential ones.
	constexpr int FMAX = FT > 0 ? FT : 16;
	LmmRec<FT> r[U];
	double L[U], a[U], mu[U], y[U], Ln[U];
	int co[U];                                                        // column offset of rate u relative to rate j0
#pragma unroll
	for (int u = 0; u < U; u++) {
		const int uu = (PARTIAL && u >= cnt) ? cnt - 1 : u;
		co[u] = uu * colStep;
		r[u].load(rec0 + uu * recStep, F);
		L[u] = L0[co[u]];
	}
	// The logarithms depend on the state only: start them before the drift needs the records.  Logarithm and reciprocal run their fast
	// paths unconditionally and share ONE cold fix-up branch (special arguments), so that all chains of the chunk stay in one basic block.
	// At the first step of a functional scheme the state is the host's log of X(0) (FIRST: its own instantiation).
	const bool fromLog = !CORRECTOR && !(MODE == 1 || (MODE == 2 && !functional));
	bool regular = true;
	if (!CORRECTOR) {
		if (!fromLog) {
#pragma unroll
			for (int u = 0; u < U; u++) y[u] = Y0[co[u]];
		} else if (FIRST) {
#pragma unroll
			for (int u = 0; u < U; u++) { const int i = i0 + ((PARTIAL && u >= cnt) ? cnt - 1 : u); y[u] = q.ylog0[SPOT ? jBeg + i : jBeg - i]; }
		} else if (LOGN) {
			regular = flogNFast<U>(logTab, L, y);
		} else {
#pragma unroll
			for (int u = 0; u < U; u++) y[u] = L[u];
		}
	}
	double den[U];
#pragma unroll
	for (int u = 0; u < U; u++) den[u] = (SPOT ? L[u] : -L[u]) + r[u].invv;   // L * (d / +-d) == +-L exactly (ratio is +1 under the spot measure, -1 under the terminal measure)
	regular = frcpNFast<U>(den, a) & regular;                                 // == 1.0 / den, bit for bit
	if (!regular) {
		if (fromLog && !FIRST && LOGN) flogNSlow<U>(L, y);
		frcpNSlow<U>(den, a);
	}
#pragma unroll
	for (int u = 0; u < U; u++) if (LOGN) a[u] = a[u] * L[u];
#pragma unroll
	for (int u = 0; u < U; u++) {
		const bool valid = !PARTIAL || u < cnt;
		if (SPOT && valid) {
#pragma unroll
			for (int k = 0; k < FMAX; k++) if (k < F) S[k] = mad<FAST>(a[u], r[u].fl[k], S[k]);
		}
		double m = FAST ? S[0] * r[u].fl[0] : S[0] * r[u].fl[0] + 0.0;
#pragma unroll
		for (int k = 1; k < FMAX; k++) if (k < F) m = mad<FAST>(S[k], r[u].fl[k], m);
		if (!SPOT && valid) {
#pragma unroll
			for (int k = 0; k < FMAX; k++) if (k < F) S[k] = mad<FAST>(a[u], r[u].fl[k], S[k]);
		}
		if (LOGN) m = m + r[u].hv;
		mu[u] = m;
	}
	if (!CORRECTOR) {
#pragma unroll
		for (int u = 0; u < U; u++) {
			y[u] = mad<FAST>(mu[u], d, y[u]);
#pragma unroll
			for (int k = 0; k < FMAX; k++) if (k < F) y[u] = mad<FAST>(w[k], r[u].fl[k], y[u]);
		}
	} else {
#pragma unroll
		for (int u = 0; u < U; u++) {
			y[u] = Y0[co[u]];
			y[u] = mad<FAST>((mu[u] - Y0[mOff + co[u]]) / 2.0, d, y[u]);
		}
	}
	// X = exp(Y) and Math.min(X, cap), with one cold branch for both: unless the cap is a zero or NaN (hasCap == 1), min is
	// (X > cap ? cap : X) bit for bit (NaN stays NaN, no signed-zero case); q.cap is +infinity when there is no cap
	double pe[U];
	int ke[U];
	bool plain = (q.hasCap != 1);
	if (LOGN) plain = fexpNParts<U>(y, pe, ke) & plain;
	if (plain) {
#pragma unroll
		for (int u = 0; u < U; u++) { Ln[u] = LOGN ? fexpScaleFast(pe[u], ke[u]) : y[u]; Ln[u] = (Ln[u] > q.cap) ? q.cap : Ln[u]; }
	} else {
#pragma unroll
		for (int u = 0; u < U; u++) {
			Ln[u] = LOGN ? expFinish(pe[u], ke[u], y[u]) : y[u];
			Ln[u] = (q.hasCap == 1) ? jminE(Ln[u], q.cap) : ((Ln[u] > q.cap) ? q.cap : Ln[u]);
		}
	}
#pragma unroll
	for (int u = 0; u < U; u++) {
		if (PARTIAL && u >= cnt) continue;
		// carried state of a capped rate at the END of a step: log(cap), what the functional scheme would re-derive from X
		if (FAST && (MODE != 2 || CORRECTOR) && q.capFix && Ln[u] == q.cap) y[u] = q.logCap;
		L0[co[u]] = Ln[u];
		if (MODE != 0) Y0[co[u]] = y[u];
		if (MODE == 2 && !CORRECTOR) Y0[mOff + co[u]] = mu[u];
		else *reinterpret_cast<double*>(reinterpret_cast<char*>(r[u].xrow) + pOff) = Ln[u];
	}
}

// One time step of one path: all live rates in chunks of U, then (predictor-corrector) the corrector pass.
template <int FT, bool LOGN, int MODE, bool SPOT, bool FAST, bool FIRST>
__device__ __forceinline__ void lmmTimeStep(const LmmParams& q, int t, int N, int F, int BD, bool functional, const double* const* __restrict__ dW,
		uint64_t p, uint64_t pOff, double* wNext, double* Lcol, double* Ybuf, size_t mOff, const double* __restrict__ logTab) {
	constexpr int FMAX = FT > 0 ? FT : 16;
	constexpr int U = FMB_LMM_U;
	const int first = q.firstLive[t];
	double w[FMAX], S[FMAX];
#pragma unroll
#if FMB_LMM_PREFETCH
	for (int k = 0; k < FMAX; k++) { w[k] = wNext[k]; S[k] = 0.0; }
#else
	for (int k = 0; k < FMAX; k++) { w[k] = (k < F) ? dW[(size_t)t * F + k][p] : 0.0; S[k] = 0.0; }
#endif
#if FMB_LMM_PREFETCH
	if (t + 1 < q.T) {
#pragma unroll
		for (int k = 0; k < FMAX; k++) if (k < F) wNext[k] = dW[(size_t)(t + 1) * F + k][p];
	}
#endif
	if (first >= N) return;
	const double d = q.dt[t];
	const int live = N - first, jBeg = SPOT ? first : N - 1;
	const int RS = FT > 0 ? ((FT + 4) & ~1) : q.recStride;                     // doubles per (t,j) record
	const int recStep = SPOT ? RS : -RS, colStep = SPOT ? BD : -BD;
	const double* recBeg = q.rec + ((size_t)t * N + jBeg) * RS;
	const double* rp = recBeg;
	double* Lp = Lcol + jBeg * BD;
	double* Yp = Ybuf + jBeg * BD;
	int i = 0;
	for (; i + U <= live; i += U, rp += U * recStep, Lp += U * colStep, Yp += U * colStep)
		lmmChunk<FT, LOGN, MODE, SPOT, U, false, false, FAST, FIRST>(q, rp, recStep, i, jBeg, colStep, F, functional, d, w, S, Lp, Yp, mOff, pOff, U, logTab);
	if (i < live)
		lmmChunk<FT, LOGN, MODE, SPOT, U, false, true, FAST, FIRST>(q, rp, recStep, i, jBeg, colStep, F, functional, d, w, S, Lp, Yp, mOff, pOff, live - i, logTab);
	if (MODE == 2) {
		// corrector: drift re-evaluated on the predicted rates (EulerSchemeFromProcessModel.java:292-314)
#pragma unroll
		for (int k = 0; k < FMAX; k++) S[k] = 0.0;
		rp = recBeg; Lp = Lcol + jBeg * BD; Yp = Ybuf + jBeg * BD;
		for (i = 0; i + U <= live; i += U, rp += U * recStep, Lp += U * colStep, Yp += U * colStep)
			lmmChunk<FT, LOGN, MODE, SPOT, U, true, false, FAST, false>(q, rp, recStep, i, jBeg, colStep, F, functional, d, w, S, Lp, Yp, mOff, pOff, U, logTab);
		if (i < live)
			lmmChunk<FT, LOGN, MODE, SPOT, U, true, true, FAST, false>(q, rp, recStep, i, jBeg, colStep, F, functional, d, w, S, Lp, Yp, mOff, pOff, live - i, logTab);
	}
}

// MODE 0: EULER_FUNCTIONAL (state = L in shared memory only).  MODE 1: EULER (Y carried in scratch).
// MODE 2: PREDICTOR_CORRECTOR[_FUNCTIONAL] (Y and the predictor drift in scratch).
template <int FT, bool LOGN, int MODE, bool SPOT, bool FAST> __global__ void __launch_bounds__(128, FMB_LMM_MINB) eulerLmmKernel(LmmParams q, uint64_t P,
		const double* const* __restrict__ dW, double* __restrict__ scratch) {
	extern __shared__ double Lsh[];                       // [N][blockDim]
	const int BD = blockDim.x, tid = threadIdx.x;
	const int N = q.N, F = FT > 0 ? FT : q.F;
	constexpr int FMAX = FT > 0 ? FT : 16;
	constexpr int U = FMB_LMM_U;
	const bool functional = (MODE == 0) || (MODE == 2 && q.scheme == SCHEME_PC_FUNCTIONAL);
	double* Ybuf = scratch + (size_t)blockIdx.x * 2 * N * BD + tid;      // [N][BD], this thread's column
	const size_t mOff = (size_t)N * BD;                                   // the predictor drift columns follow the Y columns
	double* Lcol = Lsh + tid;
	// the log table (3 KB) behind the state store: dynamic per-lane indices are cheap in shared memory
	double* logTab = Lsh + (size_t)N * BD;
	for (int i = tid; i < 384; i += BD) logTab[i] = kLogTab[i];
	__syncthreads();

	// Warps take 32-path tiles from a global counter (no block-wide barriers anywhere: every thread only touches its own column), so the
	// resident warps stay busy until the paths run out instead of each block owning a fixed share.
	const uint64_t tiles = (P + 31) / 32;
	const int lane = tid & 31;
	for (;;) {
		unsigned long long tile = 0;
		if (lane == 0) tile = atomicAdd(q.tileCounter, 1ull);
		tile = __shfl_sync(0xffffffffu, tile, 0);
		if (tile >= tiles) break;
		const uint64_t p = tile * 32 + lane;
		if (p >= P) continue;                             // the tail lanes of the last tile: the next counter value ends the loop for the whole warp
		for (int j = 0; j < N; j++) {
			Lcol[j * BD] = q.x0[j];
			if (MODE != 0) Ybuf[(size_t)j * BD] = q.y0[j];
		}
		uint64_t pOff = p * sizeof(double);
		asm volatile("" : "+l"(pOff));                    // keep the byte offset in registers (otherwise it is re-derived from tile and tid in every chunk)
		// Brownian increments come from HBM (~1 us away): fetch step t+1 while step t computes
		double wNext[FMAX];
#pragma unroll
		for (int k = 0; k < FMAX; k++) wNext[k] = (k < F) ? dW[k][p] : 0.0;
		// the first step of a functional scheme starts from the host's log X(0): its own instantiation, no per-chunk test
		lmmTimeStep<FT, LOGN, MODE, SPOT, FAST, true>(q, 0, N, F, BD, functional, dW, p, pOff, wNext, Lcol, Ybuf, mOff, logTab);
		for (int t = 1; t < q.T; t++)
			lmmTimeStep<FT, LOGN, MODE, SPOT, FAST, false>(q, t, N, F, BD, functional, dW, p, pOff, wNext, Lcol, Ybuf, mOff, logTab);
	}
}

// ---- host helpers -------------------------------------------------------------------------------------------------
struct DeviceBlob {                // one pool allocation holding all parameter tables of a launch
	void* base = nullptr;
	size_t bytes = 0;
	std::vector<unsigned char> host;
	size_t add(const void* src, size_t n) {
		size_t off = (host.size() + 15) & ~(size_t)15;
		host.resize(off + n);
		if (src) memcpy(host.data() + off, src, n);
		return off;
	}
	int upload() {
		bytes = std::max<size_t>(host.size(), 16);
		FMB_TRY(poolAlloc(bytes, &base));
		FMB_CUDA(cudaMemcpyAsync(base, host.data(), host.size(), cudaMemcpyHostToDevice, ctx().stream));
		FMB_CUDA(cudaStreamSynchronize(ctx().stream));      // host vector may go away
		return FMB_OK;
	}
	template <class Tp> const Tp* at(size_t off) const { return reinterpret_cast<const Tp*>((const unsigned char*)base + off); }
	void release() { if (base) poolFree(base, bytes); base = nullptr; }
};

static int gatherIncrements(const fmb_handle* dW, int count, uint64_t paths, std::vector<const double*>& ptrs) {
	ptrs.resize(count);
	for (int i = 0; i < count; i++) {
		if (dW[i] == 0) { setError("Brownian increment %d is not a device vector", i); return FMB_EINVAL; }
		FMB_TRY(lookupPtr(dW[i], paths, &ptrs[i]));
	}
	return FMB_OK;
}

// process storage: one slab [(T+1)*N][paths]; rows that are never written (time 0 = deterministic, frozen components) get no view
struct ProcessStore {
	Slab* slab = nullptr;
	std::vector<double*> rowPtr;    // device row pointers, nullptr when the entry is not materialised
};

static int launchCheck(const char* what) {
	countLaunch();
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) { setError("%s launch failed: %s", what, cudaGetErrorString(e)); return FMB_ECUDA; }
	return FMB_OK;
}

} // namespace fmb

using namespace fmb;

extern "C" {

int fmb_euler_black_scholes(int scheme, int T, int F, uint64_t paths, const double* dt, const fmb_handle* dW,
                            double initial_value, double risk_free_rate, double volatility, fmb_handle* out) {
	FMB_TRY(requireInit());
	if (scheme < 0 || scheme > 3 || T <= 0 || F <= 0 || paths == 0 || !dt || !dW || !out) { setError("euler_black_scholes: bad argument"); return FMB_EINVAL; }
	Context& c = ctx();
	std::vector<const double*> inc;
	FMB_TRY(gatherIncrements(dW, T * F, paths, inc));
	Slab* slab;
	FMB_TRY(newSlab((size_t)T * paths * sizeof(double), &slab));
	std::vector<double*> rows(T + 1, nullptr);
	for (int t = 1; t <= T; t++) rows[t] = (double*)slab->base + (size_t)(t - 1) * paths;
	DeviceBlob blob;
	const size_t oDt = blob.add(dt, T * sizeof(double));
	const size_t oInc = blob.add(inc.data(), inc.size() * sizeof(double*));
	const size_t oRows = blob.add(rows.data(), rows.size() * sizeof(double*));
	int rc = blob.upload();
	if (rc == FMB_OK) {
		// BlackScholesModel.java:76-78: drift = r - sigma^2/2 (host scalars), initial state log(S0), X0 = exp(log S0)
		const double drift = risk_free_rate - (volatility * volatility) / 2;
		const double y0 = std::log(initial_value), x0 = std::exp(y0), ylog0 = std::log(x0);
		const int functional = (scheme == SCHEME_EULER_FUNCTIONAL || scheme == SCHEME_PC_FUNCTIONAL) ? 1 : 0;
		const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)c.smCount * 8, (paths + 255) / 256));
		if (c.fpMode.load() == 1)
			eulerBlackScholesKernel<true><<<grid, 256, 0, c.stream>>>(functional, T, F, paths, blob.at<double>(oDt), blob.at<const double*>(oInc),
			                                                         (double* const*)blob.at<double*>(oRows), x0, y0, ylog0, drift, volatility);
		else
			eulerBlackScholesKernel<false><<<grid, 256, 0, c.stream>>>(functional, T, F, paths, blob.at<double>(oDt), blob.at<const double*>(oInc),
			                                                          (double* const*)blob.at<double*>(oRows), x0, y0, ylog0, drift, volatility);
		rc = launchCheck("euler_black_scholes");
	}
	if (rc == FMB_OK) {
		out[0] = 0;
		for (int t = 1; t <= T; t++) out[t] = newView(slab, rows[t], paths);
	} else { poolFree(slab->base, slab->bytes); delete slab; }
	blob.release();
	return rc;
}

int fmb_euler_heston(int scheme, int heston_scheme, int T, uint64_t paths, const double* dt, const fmb_handle* dW,
                     double initial_value, const double* risk_free_rate, double volatility, double theta, double kappa, double xi, double rho,
                     fmb_handle* out) {
	FMB_TRY(requireInit());
	if (scheme < 0 || scheme > 3 || heston_scheme < 0 || heston_scheme > 1 || T <= 0 || paths == 0 || !dt || !dW || !risk_free_rate || !out) {
		setError("euler_heston: bad argument"); return FMB_EINVAL;
	}
	Context& c = ctx();
	std::vector<const double*> inc;
	FMB_TRY(gatherIncrements(dW, T * 2, paths, inc));
	Slab* slab;
	FMB_TRY(newSlab((size_t)T * 2 * paths * sizeof(double), &slab));
	std::vector<double*> rows((size_t)(T + 1) * 2, nullptr);
	for (int t = 1; t <= T; t++) for (int k = 0; k < 2; k++) rows[(size_t)t * 2 + k] = (double*)slab->base + ((size_t)(t - 1) * 2 + k) * paths;
	DeviceBlob blob;
	const size_t oDt = blob.add(dt, T * sizeof(double));
	const size_t oRate = blob.add(risk_free_rate, T * sizeof(double));
	const size_t oInc = blob.add(inc.data(), inc.size() * sizeof(double*));
	const size_t oRows = blob.add(rows.data(), rows.size() * sizeof(double*));
	int rc = blob.upload();
	if (rc == FMB_OK) {
		HestonParams h;
		const double y0 = std::log(initial_value);
		h.x0 = std::exp(y0); h.ylog0 = std::log(h.x0); h.v0 = volatility * volatility;
		h.theta = theta; h.kappa = kappa; h.xi = xi; h.rho = rho;
		h.rhoBar = std::sqrt(((rho * rho) - 1) * -1);      // HestonModel.java:182
		h.hestonScheme = heston_scheme; h.scheme = scheme;
		if (scheme == SCHEME_EULER || scheme == SCHEME_PC) h.ylog0 = y0;
		const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)c.smCount * 8, (paths + 255) / 256));
		eulerHestonKernel<<<grid, 256, 0, c.stream>>>(h, T, paths, blob.at<double>(oDt), blob.at<double>(oRate), blob.at<const double*>(oInc),
		                                             (double* const*)blob.at<double*>(oRows));
		rc = launchCheck("euler_heston");
	}
	if (rc == FMB_OK) {
		out[0] = out[1] = 0;
		for (size_t i = 2; i < rows.size(); i++) out[i] = newView(slab, rows[i], paths);
	} else { poolFree(slab->base, slab->bytes); delete slab; }
	blob.release();
	return rc;
}

int fmb_euler_hull_white(int T, uint64_t paths, const double* dt, const fmb_handle* dW, const double* drift0, const double* drift1,
                         const double* fl, fmb_handle* out) {
	FMB_TRY(requireInit());
	if (T <= 0 || paths == 0 || !dt || !dW || !drift0 || !drift1 || !fl || !out) { setError("euler_hull_white: bad argument"); return FMB_EINVAL; }
	Context& c = ctx();
	std::vector<const double*> inc;
	FMB_TRY(gatherIncrements(dW, T * 2, paths, inc));
	Slab* slab;
	FMB_TRY(newSlab((size_t)T * 2 * paths * sizeof(double), &slab));
	std::vector<double*> rows((size_t)(T + 1) * 2, nullptr);
	for (int t = 1; t <= T; t++) for (int k = 0; k < 2; k++) rows[(size_t)t * 2 + k] = (double*)slab->base + ((size_t)(t - 1) * 2 + k) * paths;
	DeviceBlob blob;
	const size_t oDt = blob.add(dt, T * sizeof(double));
	const size_t oC0 = blob.add(drift0, T * sizeof(double));
	const size_t oC1 = blob.add(drift1, T * sizeof(double));
	const size_t oFl = blob.add(fl, (size_t)T * 4 * sizeof(double));
	const size_t oInc = blob.add(inc.data(), inc.size() * sizeof(double*));
	const size_t oRows = blob.add(rows.data(), rows.size() * sizeof(double*));
	int rc = blob.upload();
	if (rc == FMB_OK) {
		const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)c.smCount * 8, (paths + 255) / 256));
		eulerHullWhiteKernel<<<grid, 256, 0, c.stream>>>(T, paths, blob.at<double>(oDt), blob.at<double>(oC0), blob.at<double>(oC1), blob.at<double>(oFl),
		                                                blob.at<const double*>(oInc), (double* const*)blob.at<double*>(oRows));
		rc = launchCheck("euler_hull_white");
	}
	if (rc == FMB_OK) {
		out[0] = out[1] = 0;
		for (size_t i = 2; i < rows.size(); i++) out[i] = newView(slab, rows[i], paths);
	} else { poolFree(slab->base, slab->bytes); delete slab; }
	blob.release();
	return rc;
}

int fmb_euler_lmm(int scheme, int measure, int state_space, double libor_cap, int T, int N, int F, uint64_t paths,
                  const double* dt, const fmb_handle* dW, const double* initial_state, const double* period_length,
                  const double* factor_loading, const double* variance, const int32_t* first_live, fmb_handle* out) {
	FMB_TRY(requireInit());
	if (scheme < 0 || scheme > 3 || measure < 0 || measure > 1 || state_space < 0 || state_space > 1 || T <= 0 || N <= 0 || F <= 0 || F > 16 ||
	    paths == 0 || !dt || !dW || !initial_state || !period_length || !factor_loading || !variance || !first_live || !out) {
		setError("euler_lmm: bad argument"); return FMB_EINVAL;
	}
	Context& c = ctx();
	std::vector<const double*> inc;
	FMB_TRY(gatherIncrements(dW, T * F, paths, inc));

	// which (t, j) are materialised: component j is written at time index t+1 iff j >= first_live[t]
	size_t liveRows = 0;
	for (int t = 0; t < T; t++) { if (first_live[t] < 0) { setError("euler_lmm: negative first_live"); return FMB_EINVAL; } liveRows += (size_t)std::max(0, N - first_live[t]); }
	Slab* slab = nullptr;
	if (liveRows) FMB_TRY(newSlab(liveRows * paths * sizeof(double), &slab));
	std::vector<double*> rows((size_t)(T + 1) * N, nullptr);
	{
		size_t r = 0;
		for (int t = 0; t < T; t++) for (int j = first_live[t]; j < N; j++) rows[(size_t)(t + 1) * N + j] = (double*)slab->base + (r++) * paths;
	}
	std::vector<double> x0(N), ylog0(N);
	const int RS = (F + 4) & ~1;                    // invv, hv, row pointer, F factor loadings, padded to 16 bytes
	std::vector<double> rec((size_t)T * N * RS, 0.0);
	for (int j = 0; j < N; j++) {
		double x = state_space == 1 ? std::exp(initial_state[j]) : initial_state[j];       // applyStateSpaceTransform :1199-1212 (host scalars at t=0)
		if (!std::isinf(libor_cap)) x = (x != x) ? x : std::min(x, libor_cap);
		x0[j] = x;
		ylog0[j] = state_space == 1 ? std::log(x) : x;
	}
	for (int t = 0; t < T; t++) for (int j = 0; j < N; j++) {
		double* r = &rec[((size_t)t * N + j) * RS];
		const double value = measure == 0 ? period_length[j] : -period_length[j];          // Scalar.of(+-periodLength).discount(...) :1149,:1167
		// (the discount's other factor, periodLength / value = +-1, is the sign applied to L in the kernel)
		r[0] = 1.0 / value;
		r[1] = variance[(size_t)t * N + j] * -0.5;                                         // :1187 addProduct(variance, -0.5)
		double* rowp = rows[(size_t)(t + 1) * N + j];
		memcpy(&r[2], &rowp, sizeof(double*));
		for (int k = 0; k < F; k++) r[3 + k] = factor_loading[((size_t)t * N + j) * F + k];
	}
	DeviceBlob blob;
	const size_t oDt = blob.add(dt, T * sizeof(double));
	const size_t oFirst = blob.add(first_live, T * sizeof(int32_t));
	const size_t oRec = blob.add(rec.data(), rec.size() * sizeof(double));
	const size_t oX0 = blob.add(x0.data(), N * sizeof(double));
	const size_t oY0 = blob.add(initial_state, N * sizeof(double));
	const size_t oYl = blob.add(ylog0.data(), N * sizeof(double));
	const size_t oInc = blob.add(inc.data(), inc.size() * sizeof(double*));
	const unsigned long long zeroCounter = 0;
	const size_t oCounter = blob.add(&zeroCounter, sizeof(zeroCounter));
	int rc = blob.upload();
	void* scratch = nullptr;
	size_t scratchBytes = 0;
	if (rc == FMB_OK && liveRows) {
		LmmParams q;
		// FAST (fmb_set_fp_mode(1)): FMA contraction, and the functional schemes carry Y instead of re-deriving it as log(exp(Y)) every
		// step (equal up to one rounding of Y per step; a capped rate carries log(cap)).  Only for the log-normal model.
		const bool fast = (c.fpMode.load() == 1) && state_space == 1;
		const bool functionalScheme = (scheme == SCHEME_EULER_FUNCTIONAL || scheme == SCHEME_PC_FUNCTIONAL);
		int kernelScheme = scheme;
		if (fast && functionalScheme) kernelScheme = (scheme == SCHEME_EULER_FUNCTIONAL) ? SCHEME_EULER : SCHEME_PC;
		q.scheme = kernelScheme; q.measure = measure; q.hasCap = (std::isinf(libor_cap) && libor_cap > 0) ? 0 : ((libor_cap == 0.0 || std::isnan(libor_cap)) ? 1 : 2); q.cap = libor_cap;
		q.capFix = (fast && functionalScheme && q.hasCap) ? 1 : 0; q.logCap = q.hasCap ? std::log(libor_cap) : 0.0;
		q.T = T; q.N = N; q.F = F; q.recStride = RS;
		q.dt = blob.at<double>(oDt); q.firstLive = blob.at<int>(oFirst); q.rec = blob.at<double>(oRec);
		q.x0 = blob.at<double>(oX0); q.y0 = blob.at<double>(oY0); q.ylog0 = blob.at<double>(oYl);
		q.tileCounter = const_cast<unsigned long long*>(blob.at<unsigned long long>(oCounter));
		// block size: the shared-memory column store is 8*N bytes per thread
		int BD = 128;
		if (const char* e = getenv("FMB_LMM_BD")) { const int v = atoi(e); if (v == 32 || v == 64 || v == 128) BD = v; }   // tuning hook (profiles/r01_notes.md)
		while (BD > 32 && (size_t)BD * N * sizeof(double) > 196 * 1024) BD >>= 1;
		const size_t smem = (size_t)BD * N * sizeof(double) + 384 * sizeof(double);      // state store + log table
		if (smem > 220 * 1024) { setError("euler_lmm: %d components exceed the shared-memory state store", N); rc = FMB_EUNSUPPORTED; }
		if (rc == FMB_OK) {
			const int perSm = (int)std::max<size_t>(1, std::min<size_t>(16, (220 * 1024) / std::max<size_t>(smem, 1)));
			const uint64_t tiles = (paths + BD - 1) / BD;
			const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)c.smCount * perSm, tiles));
			const int mode = kernelScheme == SCHEME_EULER_FUNCTIONAL ? 0 : (kernelScheme == SCHEME_EULER ? 1 : 2);
			scratchBytes = mode != 0 ? (size_t)grid * 2 * N * BD * sizeof(double) : 16;
			rc = poolAlloc(scratchBytes, &scratch);
			if (rc == FMB_OK) {
				auto launch = [&](auto kernel) {
					cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
					kernel<<<grid, BD, smem, c.stream>>>(q, paths, blob.at<const double*>(oInc), (double*)scratch);
				};
#define LMM_SPOT(FTV, LOGNV, MODEV, FASTV) \
				if (measure == 0) launch(eulerLmmKernel<FTV, LOGNV, MODEV, true, FASTV>); else launch(eulerLmmKernel<FTV, LOGNV, MODEV, false, FASTV>);
#define LMM_MODE(FTV, LOGNV) \
				switch (mode) { case 0: LMM_SPOT(FTV, LOGNV, 0, false) break; case 1: LMM_SPOT(FTV, LOGNV, 1, false) break; default: LMM_SPOT(FTV, LOGNV, 2, false) break; }
#define LMM_FAST(FTV) if (mode == 1) { LMM_SPOT(FTV, true, 1, true) } else { LMM_SPOT(FTV, true, 2, true) }
#define LMM_LOGN(FTV) if (fast) { LMM_FAST(FTV) } else if (state_space == 1) { LMM_MODE(FTV, true) } else { LMM_MODE(FTV, false) }
				switch (F) {
				case 1: LMM_LOGN(1) break;
				case 2: LMM_LOGN(2) break;
				case 3: LMM_LOGN(3) break;
				default: LMM_LOGN(0) break;
				}
#undef LMM_LOGN
#undef LMM_FAST
#undef LMM_SPOT
#undef LMM_MODE
				rc = launchCheck("euler_lmm");
			}
		}
	}
	if (rc == FMB_OK) {
		// handles: time 0 deterministic (0); a frozen component aliases the previous time index (one more reference)
		for (int j = 0; j < N; j++) out[j] = 0;
		for (int t = 1; t <= T; t++) for (int j = 0; j < N; j++) {
			const size_t i = (size_t)t * N + j;
			if (rows[i]) out[i] = newView(slab, rows[i], paths);
			else { out[i] = out[i - N]; if (out[i]) fmb_rv_retain(out[i]); }
		}
	} else if (slab) { poolFree(slab->base, slab->bytes); delete slab; }
	if (scratch) poolFree(scratch, scratchBytes);
	blob.release();
	return rc;
}

} // extern "C"
