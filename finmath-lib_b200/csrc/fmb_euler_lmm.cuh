// LIBOR market model Euler kernel (device code + launcher template), shared by fmb_euler_lmm_f*.cu: one translation unit per
// compile-time factor count so that the 16 instantiations of each F build in parallel and every F = 1..8 is a first-class path
// (the reference's own LMM test uses 6 factors, T/montecarlo/interestrate/LIBORMarketModelValuationTest.java:75).
//
// LIBORMarketModelFromCovarianceModel.java:1124-1223 inside the Euler scheme (EulerSchemeFromProcessModel.java:170-326).
//   drift (spot):   a_j = 1/(L_j*(d/d) + 1/d) [*L_j if lognormal];  S_k += a_j*fl_jk;  mu_j = sum_k S_k*fl_jk  (+ -0.5*var_j)
//   drift (terminal): the mirrored suffix sum with a_j = 1/(L_j*(d/(-d)) + 1/(-d)), mu_j formed BEFORE S is updated.
// State: the current forward rates of the thread's path live in shared memory as Lsh[j][thread] (conflict-free, j is a
// run-time loop so the drift's prefix sum over the rate index runs sequentially per lane in the reference's order).
// Y (non-functional schemes) and mu (predictor-corrector) live in a block-private, L2-resident global scratch.
//
// Factor vectors (the running sums S_k, the increments w_k, the loadings fl_jk of a record): registers for FT = 1..8 (every loop
// over k fully unrolled); for FT = 0 (any F, used above 8) S and w are per-thread shared-memory columns and the loadings are
// re-read from the L1-resident record, so that no instantiation has a register array indexed at run time (no local-memory spills).
#pragma once
#include "fmb_common.cuh"
#include "fmb_math.cuh"
#include <cmath>
#include <algorithm>

namespace fmb {

#ifndef FMB_LMM_U
#define FMB_LMM_U 2          // live rates processed together per thread (ILP); 2 measured best on B200 (profiles/r01_notes.md)
#endif

enum { SCHEME_EULER = 0, SCHEME_PC = 1, SCHEME_EULER_FUNCTIONAL = 2, SCHEME_PC_FUNCTIONAL = 3 };

// resident CTAs (128 threads) per SM the register budget is set for: 5 up to three factors (<= 102 registers, measured best,
// profiles/r01_notes.md), fewer as the per-thread factor vectors grow (2*U*F + 4*F more registers than F = 0 would need)
template <int FT> struct LmmOcc { static constexpr int minBlocks = FT == 0 ? 4 : (FT <= 3 ? 5 : (FT == 4 ? 4 : 3)); };

// a*b + c: STRICT = rounded product then rounded sum (the JVM never contracts); FAST (fmb_set_fp_mode(1)) = one fused multiply-add.
template <bool FAST> __device__ __forceinline__ double mad(double a, double b, double c) { return FAST ? fma(a, b, c) : a * b + c; }

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

// per-thread vector over the factors
template <int FT> struct FacVec {
	double v[FT];
	__device__ __forceinline__ double get(int k) const { return v[k]; }
	__device__ __forceinline__ void set(int k, double x) { v[k] = x; }
};
template <> struct FacVec<0> {
	double* p;               // shared-memory column of this thread, element k at p[k * stride]
	int stride;
	__device__ __forceinline__ double get(int k) const { return p[k * stride]; }
	__device__ __forceinline__ void set(int k, double x) { p[k * stride] = x; }
};

// One (t, j) record is read by every thread of every block in the same order: 16-byte uniform loads, L1-resident.
template <int FT> struct LmmRec {
	double invv, hv;
	double* xrow;
	double flv[FT > 0 ? FT : 1];
	const double* flp;
	// layout: invv, hv, row pointer, fl[0..F), padded to an even number of doubles (F = 3: 48 bytes, three 16-byte loads)
	__device__ __forceinline__ void load(const double* __restrict__ r) {
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
			for (int k = 0; k < FT; k++) flv[k] = v[1 + k];
		} else {
			xrow = reinterpret_cast<double*>(__double_as_longlong(__ldg(r + 2)));
			flp = r + 3;
		}
	}
	__device__ __forceinline__ double fl(int k) const { return FT > 0 ? flv[k] : __ldg(flp + k); }
};

// U consecutive live rates of one path at once (i = position in processing order; j = first+i for the spot measure,
// N-1-i for the terminal measure).  Per rate the operations and their order are exactly those of the scalar recipe; the
// only cross-rate dependency is the running factor sum S, so the U log / exp / reciprocal chains overlap (ILP U).
template <int FT, bool LOGN, int MODE, bool SPOT, int U, bool CORRECTOR, bool PARTIAL, bool FAST, bool FIRST>
__device__ __forceinline__ void lmmChunk(const LmmParams& q, const double* __restrict__ rec0, int recStep, int i0, int jBeg, int colStep, int Frt,
		bool functional, double d, const FacVec<FT>& w, FacVec<FT>& S, double* L0, double* Y0, size_t mOff, uint64_t pOff, int cnt,
		const double* __restrict__ logTab) {
	// rec0 / L0 / Y0 point at the chunk's first rate (record, shared-memory state, scratch column; the predictor drift column is Y0 + mOff);
	// recStep / colStep move them to the next rate in processing order (the caller advances them chunk by chunk, so no index
	// multiplications are left in the loop).  i0 = position of the first rate in processing order (j = jBeg +- i).
	// PARTIAL: only the first cnt (< U) rates are real; the others recompute rate cnt-1 and are masked out of S and of every store,
	// so a short remainder costs one chunk latency instead of cnt sequential ones.
	const int F = FT > 0 ? FT : Frt;
	LmmRec<FT> r[U];
	double L[U], a[U], mu[U], y[U], Ln[U];
	int co[U];                                                        // column offset of rate u relative to rate j0
#pragma unroll
	for (int u = 0; u < U; u++) {
		const int uu = (PARTIAL && u >= cnt) ? cnt - 1 : u;
		co[u] = uu * colStep;
		r[u].load(rec0 + uu * recStep);
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
			for (int k = 0; k < F; k++) S.set(k, mad<FAST>(a[u], r[u].fl(k), S.get(k)));
		}
		double m = FAST ? S.get(0) * r[u].fl(0) : S.get(0) * r[u].fl(0) + 0.0;
#pragma unroll
		for (int k = 1; k < F; k++) m = mad<FAST>(S.get(k), r[u].fl(k), m);
		if (!SPOT && valid) {
#pragma unroll
			for (int k = 0; k < F; k++) S.set(k, mad<FAST>(a[u], r[u].fl(k), S.get(k)));
		}
		if (LOGN) m = m + r[u].hv;
		mu[u] = m;
	}
	if (!CORRECTOR) {
#pragma unroll
		for (int u = 0; u < U; u++) {
			y[u] = mad<FAST>(mu[u], d, y[u]);
#pragma unroll
			for (int k = 0; k < F; k++) y[u] = mad<FAST>(w.get(k), r[u].fl(k), y[u]);
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
__device__ __forceinline__ void lmmTimeStep(const LmmParams& q, int t, int N, int Frt, int BD, bool functional, const double* const* __restrict__ dW,
		uint64_t p, uint64_t pOff, double* facCol, double* Lcol, double* Ybuf, size_t mOff, const double* __restrict__ logTab) {
	constexpr int U = FMB_LMM_U;
	const int F = FT > 0 ? FT : Frt;
	const int first = q.firstLive[t];
	FacVec<FT> w, S;
	if constexpr (FT == 0) { S.p = facCol; S.stride = BD; w.p = facCol + (size_t)F * BD; w.stride = BD; }
#pragma unroll
	for (int k = 0; k < F; k++) { w.set(k, dW[(size_t)t * F + k][p]); S.set(k, 0.0); }
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
		for (int k = 0; k < F; k++) S.set(k, 0.0);
		rp = recBeg; Lp = Lcol + jBeg * BD; Yp = Ybuf + jBeg * BD;
		for (i = 0; i + U <= live; i += U, rp += U * recStep, Lp += U * colStep, Yp += U * colStep)
			lmmChunk<FT, LOGN, MODE, SPOT, U, true, false, FAST, false>(q, rp, recStep, i, jBeg, colStep, F, functional, d, w, S, Lp, Yp, mOff, pOff, U, logTab);
		if (i < live)
			lmmChunk<FT, LOGN, MODE, SPOT, U, true, true, FAST, false>(q, rp, recStep, i, jBeg, colStep, F, functional, d, w, S, Lp, Yp, mOff, pOff, live - i, logTab);
	}
}

// MODE 0: EULER_FUNCTIONAL (state = L in shared memory only).  MODE 1: EULER (Y carried in scratch).
// MODE 2: PREDICTOR_CORRECTOR[_FUNCTIONAL] (Y and the predictor drift in scratch).
// Shared memory: Lsh[N][BD] | log table (384 doubles) | FT == 0 only: S[F][BD], w[F][BD].
template <int FT, bool LOGN, int MODE, bool SPOT, bool FAST> __global__ void __launch_bounds__(128, LmmOcc<FT>::minBlocks) eulerLmmKernel(LmmParams q, uint64_t P,
		const double* const* __restrict__ dW, double* __restrict__ scratch) {
	extern __shared__ double Lsh[];                       // [N][blockDim]
	const int BD = blockDim.x, tid = threadIdx.x;
	const int N = q.N, F = FT > 0 ? FT : q.F;
	const bool functional = (MODE == 0) || (MODE == 2 && q.scheme == SCHEME_PC_FUNCTIONAL);
	double* Ybuf = scratch + (size_t)blockIdx.x * 2 * N * BD + tid;      // [N][BD], this thread's column
	const size_t mOff = (size_t)N * BD;                                   // the predictor drift columns follow the Y columns
	double* Lcol = Lsh + tid;
	// the log table (3 KB) behind the state store: dynamic per-lane indices are cheap in shared memory
	double* logTab = Lsh + (size_t)N * BD;
	double* facCol = logTab + 384 + tid;                                  // FT == 0: this thread's S / w columns
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
		// the first step of a functional scheme starts from the host's log X(0): its own instantiation, no per-chunk test
		lmmTimeStep<FT, LOGN, MODE, SPOT, FAST, true>(q, 0, N, F, BD, functional, dW, p, pOff, facCol, Lcol, Ybuf, mOff, logTab);
		for (int t = 1; t < q.T; t++)
			lmmTimeStep<FT, LOGN, MODE, SPOT, FAST, false>(q, t, N, F, BD, functional, dW, p, pOff, facCol, Lcol, Ybuf, mOff, logTab);
	}
}

// ---- launch -------------------------------------------------------------------------------------------------------
struct LmmLaunch {
	LmmParams q;
	uint64_t paths;
	const double* const* dW;   // device array of T*F increment pointers
	int mode;                  // 0 EULER_FUNCTIONAL, 1 EULER, 2 predictor-corrector
	bool fast, logn, spot;
	int BD;                    // threads per CTA (32, 64 or 128)
	size_t smem;
	cudaStream_t stream;
	// out: geometry chosen by the launcher (the caller sizes the scratch from it before the launch: two-phase call)
	int grid;
	double* scratch;
};

// phase 0 (scratch == nullptr): choose the grid (resident CTAs per SM from the occupancy calculator x SMs, capped by the tiles) and
// return; phase 1: launch.  One function per compile-time F (fmb_euler_lmm_f*.cu).
typedef int (*LmmLaunchFn)(LmmLaunch&, int smCount, bool launchNow);

template <int FT, bool LOGN, int MODE, bool SPOT, bool FAST> static int lmmLaunchOne(LmmLaunch& a, int smCount, bool launchNow) {
	auto kernel = eulerLmmKernel<FT, LOGN, MODE, SPOT, FAST>;
	if (!launchNow) {
		FMB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)a.smem));
		int perSm = 0;
		FMB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kernel, a.BD, a.smem));
		if (perSm < 1) { setError("euler_lmm: kernel does not fit one SM (%zu bytes shared memory)", a.smem); return FMB_EUNSUPPORTED; }
		const uint64_t tiles = (a.paths + a.BD - 1) / a.BD;
		a.grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)smCount * perSm, tiles));
		return FMB_OK;
	}
	kernel<<<a.grid, a.BD, a.smem, a.stream>>>(a.q, a.paths, a.dW, a.scratch);
	return FMB_OK;
}

template <int FT> static int lmmLaunchF(LmmLaunch& a, int smCount, bool launchNow) {
#define LMM_SPOT(LOGNV, MODEV, FASTV) \
	return a.spot ? lmmLaunchOne<FT, LOGNV, MODEV, true, FASTV>(a, smCount, launchNow) : lmmLaunchOne<FT, LOGNV, MODEV, false, FASTV>(a, smCount, launchNow);
#define LMM_MODE(LOGNV) \
	switch (a.mode) { case 0: LMM_SPOT(LOGNV, 0, false) case 1: LMM_SPOT(LOGNV, 1, false) default: LMM_SPOT(LOGNV, 2, false) }
	if (a.fast) { if (a.mode == 1) { LMM_SPOT(true, 1, true) } else { LMM_SPOT(true, 2, true) } }
	if (a.logn) { LMM_MODE(true) }
	LMM_MODE(false)
#undef LMM_SPOT
#undef LMM_MODE
}

#define FMB_LMM_DECLARE(FT) int lmmLaunchF##FT(LmmLaunch& a, int smCount, bool launchNow);
FMB_LMM_DECLARE(0) FMB_LMM_DECLARE(1) FMB_LMM_DECLARE(2) FMB_LMM_DECLARE(3) FMB_LMM_DECLARE(4)
FMB_LMM_DECLARE(5) FMB_LMM_DECLARE(6) FMB_LMM_DECLARE(7) FMB_LMM_DECLARE(8)
int lmmLaunchLanePerRate(LmmLaunch& a, int smCount);   // experiment, fmb_euler_lmm_shuffle.cu
#define FMB_LMM_DEFINE(FT) int lmmLaunchF##FT(LmmLaunch& a, int smCount, bool launchNow) { return lmmLaunchF<FT>(a, smCount, launchNow); }

} // namespace fmb
