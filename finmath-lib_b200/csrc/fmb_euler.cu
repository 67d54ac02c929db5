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
#include "fmb_euler_lmm.cuh"

namespace fmb {

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
// Two-component, two-factor models (Heston, Hull-White) with the increments streamed through shared memory by the bulk asynchronous
// copy engine (TMA, cp.async.bulk + mbarrier).  The plain kernels above issue two dependent 8-byte loads per lane and step; with a
// thousand sequential steps per path ncu shows long_scoreboard as the dominant stall (52 %) and 256-byte row segments that scatter over
// thousands of DRAM pages (0.55 of the HBM peak, profiles/r01_notes.md).  Here a CTA owns a tile of 2*NT consecutive paths (two paths
// per thread: ILP 2 on the exp / log / sqrt chains); one elected thread keeps S stages of KS time steps in flight - per stage 2*KS bulk
// copies of the tile's row segments (8*2*NT contiguous bytes each, 4 KB at NT = 256) that complete on the stage's "full" mbarrier; the
// warps release a stage through its "empty" mbarrier.  No block-wide barrier in the time loop.  Results are written with 16-byte
// stores (the thread's two adjacent paths), 16*NT contiguous bytes per row and CTA.  Same arithmetic, same order: bit-identical to the
// plain kernels.  Needs an even number of paths (16-byte aligned rows); the plain kernels remain for odd path counts.
// Algorithmic HBM bytes per path-step: 16 read + 16 written.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarInit(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smemAddr(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbarExpectTx(uint64_t* bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarArrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smemAddr(bar)) : "memory"); }
__device__ __forceinline__ void mbarWait(uint64_t* bar, uint32_t parity) {
	uint32_t done;
	do {
		asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smemAddr(bar)), "r"(parity) : "memory");
	} while (!done);
}
__device__ __forceinline__ void bulkLoad(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             :: "r"(smemAddr(sdst)), "l"(gsrc), "r"(bytes), "r"(smemAddr(bar)) : "memory");
}

struct HestonStep {
	HestonParams h;
	const double* dt; const double* rate;
	bool functional, pc;
	__device__ __forceinline__ void init(double* a, double* b, double* ya, double* yb) const {
#pragma unroll
		for (int u = 0; u < 2; u++) { a[u] = h.x0; b[u] = h.v0; ya[u] = h.ylog0; yb[u] = h.v0; }
	}
	// two paths at once; per path exactly the operations of eulerHestonKernel
	__device__ __forceinline__ void step(int t, const double* w0, const double* w1, double* s, double* v, double* y0, double* y1) const {
		const double d = __ldg(dt + t), r = __ldg(rate + t);
		double var[2], mu0[2], mu1[2], vol[2];
#pragma unroll
		for (int u = 0; u < 2; u++) { hestonDrift(h, v[u], r, var[u], mu0[u], mu1[u]); vol[u] = sqrt(var[u]); }
		if (functional) {
			if (t == 0) { y0[0] = h.ylog0; y0[1] = h.ylog0; } else flogN<2>(s, y0);
			y1[0] = v[0]; y1[1] = v[1];
		}
#pragma unroll
		for (int u = 0; u < 2; u++) {
			y0[u] = y0[u] + mu0[u] * d;
			y0[u] = y0[u] + vol[u] * w0[u];
			y0[u] = y0[u] + w1[u] * 0.0;
			const double volv = vol[u] * h.xi;
			y1[u] = y1[u] + mu1[u] * d;
			y1[u] = y1[u] + (volv * h.rho) * w0[u];
			y1[u] = y1[u] + (volv * h.rhoBar) * w1[u];
		}
		fexpN<2>(y0, s);
		v[0] = y1[0]; v[1] = y1[1];
		if (pc) {
#pragma unroll
			for (int u = 0; u < 2; u++) {
				double varP, mu0P, mu1P;
				hestonDrift(h, v[u], r, varP, mu0P, mu1P);
				y0[u] = y0[u] + ((mu0P - mu0[u]) / 2.0) * d;
				y1[u] = y1[u] + ((mu1P - mu1[u]) / 2.0) * d;
			}
			fexpN<2>(y0, s);
			v[0] = y1[0]; v[1] = y1[1];
		}
	}
};

struct HullWhiteStep {
	const double* dt; const double* c0; const double* c1; const double* fl;
	__device__ __forceinline__ void init(double* a, double* b, double* ya, double* yb) const {
#pragma unroll
		for (int u = 0; u < 2; u++) { a[u] = 0.0; b[u] = 0.0; ya[u] = 0.0; yb[u] = 0.0; }
	}
	__device__ __forceinline__ void step(int t, const double* w0, const double* w1, double* x0, double* x1, double*, double*) const {
		const double d = __ldg(dt + t), k0 = __ldg(c0 + t), k1 = __ldg(c1 + t);
		const double2 fa = __ldg(reinterpret_cast<const double2*>(fl) + 2 * t), fb = __ldg(reinterpret_cast<const double2*>(fl) + 2 * t + 1);
#pragma unroll
		for (int u = 0; u < 2; u++) {
			const double mu0 = x0[u] * k0, mu1 = x0[u] * k1;
			double y0 = x0[u] + mu0 * d;
			y0 = y0 + w0[u] * fa.x;
			y0 = y0 + w1[u] * fa.y;
			double y1 = x1[u] + mu1 * d;
			y1 = y1 + w0[u] * fb.x;
			y1 = y1 + w1[u] * fb.y;
			x0[u] = y0; x1[u] = y1;
		}
	}
};

template <class Model, int NT, int KS, int S> __global__ void __launch_bounds__(NT) eulerTwoFactorTmaKernel(Model m, int T, uint64_t P,
		const double* const* __restrict__ dW, double* const* __restrict__ X) {
	constexpr int TP = 2 * NT;                                   // paths per tile
	extern __shared__ __align__(128) unsigned char smemRaw[];
	double* stage = reinterpret_cast<double*>(smemRaw);          // [S][KS * 2 rows][TP]
	uint64_t* full = reinterpret_cast<uint64_t*>(smemRaw + (size_t)S * KS * 2 * TP * sizeof(double));
	uint64_t* empty = full + S;
	const int tid = threadIdx.x, lane = tid & 31;
	if (tid == 0) {
		for (int s = 0; s < S; s++) { mbarInit(full + s, 1); mbarInit(empty + s, NT / 32); }
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	const uint64_t tiles = (P + TP - 1) / TP;
	const int chunksPerTile = (T + KS - 1) / KS;
	// chunk sequence of this CTA: (tile, chunk in tile), tiles blockIdx.x, + gridDim.x, ...; g counts them, stage = g % S, phase = (g / S) & 1
	uint64_t myTiles = blockIdx.x < tiles ? (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
	const uint64_t totalChunks = myTiles * chunksPerTile;
	uint64_t gIssue = 0;                                          // (thread 0) next chunk to request
	auto issue = [&](uint64_t g) {
		const uint64_t tileIdx = blockIdx.x + (g / chunksPerTile) * gridDim.x;
		const int ck = (int)(g % chunksPerTile);
		const uint64_t p0 = tileIdx * TP;
		const uint32_t n = (uint32_t)min((uint64_t)TP, P - p0);
		const int t0 = ck * KS, steps = min(KS, T - t0);
		const int s = (int)(g % S);
		if (g >= (uint64_t)S) mbarWait(empty + s, (uint32_t)(((g / S) - 1) & 1));     // every warp has finished reading the stage's previous chunk
		mbarExpectTx(full + s, (uint32_t)(steps * 2) * n * 8u);
		double* dst = stage + (size_t)s * KS * 2 * TP;
		for (int r = 0; r < steps * 2; r++) bulkLoad(dst + (size_t)r * TP, dW[(size_t)t0 * 2 + r] + p0, n * 8u, full + s);
	};
	if (tid == 0) { for (; gIssue < (uint64_t)(S - 1) && gIssue < totalChunks; gIssue++) issue(gIssue); }
	uint64_t g = 0;
	for (uint64_t tileIdx = blockIdx.x; tileIdx < tiles; tileIdx += gridDim.x) {
		const uint64_t p0 = tileIdx * TP;
		const uint64_t p = p0 + 2 * (uint64_t)tid;
		const bool active = p < P;                                // P is even: both paths of an active thread exist
		double a[2], b[2], ya[2], yb[2];
		m.init(a, b, ya, yb);
		for (int ck = 0; ck < chunksPerTile; ck++, g++) {
			if (tid == 0 && gIssue < totalChunks) { issue(gIssue); gIssue++; }
			const int s = (int)(g % S);
			mbarWait(full + s, (uint32_t)((g / S) & 1));
			const double* src = stage + (size_t)s * KS * 2 * TP + 2 * tid;
			const int t0 = ck * KS, steps = min(KS, T - t0);
			if (active) {
#pragma unroll
				for (int k = 0; k < KS; k++) {
					if (k < steps) {
						const double2 w0 = *reinterpret_cast<const double2*>(src + (size_t)(2 * k) * TP);
						const double2 w1 = *reinterpret_cast<const double2*>(src + (size_t)(2 * k + 1) * TP);
						const double w0v[2] = { w0.x, w0.y }, w1v[2] = { w1.x, w1.y };
						m.step(t0 + k, w0v, w1v, a, b, ya, yb);
						*reinterpret_cast<double2*>(X[2 * (size_t)(t0 + k + 1)] + p) = make_double2(a[0], a[1]);
						*reinterpret_cast<double2*>(X[2 * (size_t)(t0 + k + 1) + 1] + p) = make_double2(b[0], b[1]);
					}
				}
			}
			__syncwarp();
			if (lane == 0) mbarArrive(empty + s);
		}
	}
}

template <class Model, int NT, int KS, int S> static int launchTwoFactorTmaCfg(const Model& m, int T, uint64_t paths, const double* const* dW, double* const* X) {
	Context& c = ctx();
	auto kernel = eulerTwoFactorTmaKernel<Model, NT, KS, S>;
	const size_t smem = (size_t)S * KS * 2 * (2 * NT) * sizeof(double) + 2 * S * sizeof(uint64_t);
	FMB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	int perSm = 0;
	FMB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kernel, NT, smem));
	const uint64_t tiles = (paths + 2 * NT - 1) / (2 * NT);
	const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)c.smCount * std::max(perSm, 1), tiles));
	kernel<<<grid, NT, smem, c.stream>>>(m, T, paths, dW, X);
	return FMB_OK;
}
// (threads per CTA, time steps per stage, stages): FMB_TMA_CFG selects one for A/B runs; the defaults are the measured best per model
// (profiles/r02_notes.md)
template <class Model> static int launchTwoFactorTma(const Model& m, int T, uint64_t paths, const double* const* dW, double* const* X, int defaultCfg) {
	int cfg = defaultCfg;
	if (const char* e = getenv("FMB_TMA_CFG")) cfg = atoi(e);
	switch (cfg) {
	case 1: return launchTwoFactorTmaCfg<Model, 256, 4, 3>(m, T, paths, dW, X);
	case 2: return launchTwoFactorTmaCfg<Model, 256, 8, 2>(m, T, paths, dW, X);
	case 3: return launchTwoFactorTmaCfg<Model, 128, 4, 4>(m, T, paths, dW, X);
	case 4: return launchTwoFactorTmaCfg<Model, 128, 8, 3>(m, T, paths, dW, X);
	case 5: return launchTwoFactorTmaCfg<Model, 128, 2, 6>(m, T, paths, dW, X);
	default: return launchTwoFactorTmaCfg<Model, 256, 2, 4>(m, T, paths, dW, X);
	}
}
// the bulk-copy kernels need 16-byte aligned row segments: an even number of paths (FMB_EULER_TMA=0 forces the plain kernels: A/B runs)
static bool useTwoFactorTma(uint64_t paths, bool byDefault) {
	if (paths % 2 != 0 || paths < 2) return false;
	const char* e = getenv("FMB_EULER_TMA");
	return e ? atoi(e) != 0 : byDefault;
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
	PinScope pins;
	if (scheme < 0 || scheme > 3 || T <= 0 || F <= 0 || !dt || !dW || !out) { setError("euler_black_scholes: bad argument"); return FMB_EINVAL; }
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
	if (rc == FMB_OK && paths > 0) {                           // (a rank may own no paths: zero-length vectors, nothing to launch)
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
	PinScope pins;
	if (scheme < 0 || scheme > 3 || heston_scheme < 0 || heston_scheme > 1 || T <= 0 || !dt || !dW || !risk_free_rate || !out) {
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
	if (rc == FMB_OK && paths > 0) {
		HestonParams h;
		const double y0 = std::log(initial_value);
		h.x0 = std::exp(y0); h.ylog0 = std::log(h.x0); h.v0 = volatility * volatility;
		h.theta = theta; h.kappa = kappa; h.xi = xi; h.rho = rho;
		h.rhoBar = std::sqrt(((rho * rho) - 1) * -1);      // HestonModel.java:182
		h.hestonScheme = heston_scheme; h.scheme = scheme;
		if (scheme == SCHEME_EULER || scheme == SCHEME_PC) h.ylog0 = y0;
		// (Heston is bound by its exp / log / sqrt chains, not by the increment loads: the plain kernel measured faster, profiles/r02_notes.md)
		if (useTwoFactorTma(paths, false)) {
			HestonStep m;
			m.h = h; m.dt = blob.at<double>(oDt); m.rate = blob.at<double>(oRate);
			m.functional = (scheme == SCHEME_EULER_FUNCTIONAL || scheme == SCHEME_PC_FUNCTIONAL);
			m.pc = (scheme == SCHEME_PC || scheme == SCHEME_PC_FUNCTIONAL);
			rc = launchTwoFactorTma(m, T, paths, blob.at<const double*>(oInc), (double* const*)blob.at<double*>(oRows), 0);
		} else {
			const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)c.smCount * 8, (paths + 255) / 256));
			eulerHestonKernel<<<grid, 256, 0, c.stream>>>(h, T, paths, blob.at<double>(oDt), blob.at<double>(oRate), blob.at<const double*>(oInc),
			                                             (double* const*)blob.at<double*>(oRows));
		}
		if (rc == FMB_OK) rc = launchCheck("euler_heston");
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
	PinScope pins;
	if (T <= 0 || !dt || !dW || !drift0 || !drift1 || !fl || !out) { setError("euler_hull_white: bad argument"); return FMB_EINVAL; }
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
	if (rc == FMB_OK && paths > 0) {
		// (Hull-White has no transcendental in the step: HBM-bound, 0.92 of the measured copy bandwidth with the bulk-copy pipeline vs 0.68 without)
		if (useTwoFactorTma(paths, true)) {
			HullWhiteStep m;
			m.dt = blob.at<double>(oDt); m.c0 = blob.at<double>(oC0); m.c1 = blob.at<double>(oC1); m.fl = blob.at<double>(oFl);
			rc = launchTwoFactorTma(m, T, paths, blob.at<const double*>(oInc), (double* const*)blob.at<double*>(oRows), 0);
		} else {
			const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)c.smCount * 8, (paths + 255) / 256));
			eulerHullWhiteKernel<<<grid, 256, 0, c.stream>>>(T, paths, blob.at<double>(oDt), blob.at<double>(oC0), blob.at<double>(oC1), blob.at<double>(oFl),
			                                                blob.at<const double*>(oInc), (double* const*)blob.at<double*>(oRows));
		}
		if (rc == FMB_OK) rc = launchCheck("euler_hull_white");
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
	PinScope pins;
	if (scheme < 0 || scheme > 3 || measure < 0 || measure > 1 || state_space < 0 || state_space > 1 || T <= 0 || N <= 0 || F <= 0 || F > 64 ||
	    !dt || !dW || !initial_state || !period_length || !factor_loading || !variance || !first_live || !out) {
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
	if (rc == FMB_OK && liveRows && paths > 0) {
		LmmLaunch L;
		LmmParams& q = L.q;
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
		// compile-time factor counts 1..8 (factor vectors in registers); above that the run-time-F kernel (factor vectors in shared memory)
		static const LmmLaunchFn launchers[9] = { lmmLaunchF0, lmmLaunchF1, lmmLaunchF2, lmmLaunchF3, lmmLaunchF4, lmmLaunchF5, lmmLaunchF6, lmmLaunchF7, lmmLaunchF8 };
		const LmmLaunchFn launcher = launchers[F <= 8 ? F : 0];
		const size_t facDoubles = F <= 8 ? 0 : 2 * (size_t)F;            // per-thread S and w columns of the run-time-F kernel
		// block size: the shared-memory column store is 8*(N + facDoubles) bytes per thread
		int BD = 128;
		if (const char* e = getenv("FMB_LMM_BD")) { const int v = atoi(e); if (v == 32 || v == 64 || v == 128) BD = v; }   // tuning hook (profiles/r01_notes.md)
		while (BD > 32 && (size_t)BD * (N + facDoubles) * sizeof(double) > 196 * 1024) BD >>= 1;
		L.smem = (size_t)BD * (N + facDoubles) * sizeof(double) + 384 * sizeof(double);      // state store + log table (+ factor columns)
		if (L.smem > 220 * 1024) { setError("euler_lmm: %d components x %d factors exceed the shared-memory state store", N, F); rc = FMB_EUNSUPPORTED; }
		if (rc == FMB_OK) {
			L.paths = paths; L.dW = blob.at<const double*>(oInc); L.BD = BD; L.stream = c.stream; L.scratch = nullptr; L.grid = 0;
			L.mode = kernelScheme == SCHEME_EULER_FUNCTIONAL ? 0 : (kernelScheme == SCHEME_EULER ? 1 : 2);
			L.fast = fast; L.logn = state_space == 1; L.spot = measure == 0;
			rc = launcher(L, c.smCount, false);                           // geometry (occupancy of this instantiation)
		}
		if (rc == FMB_OK) {
			scratchBytes = L.mode != 0 ? (size_t)L.grid * 2 * N * BD * sizeof(double) : 16;
			rc = poolAlloc(scratchBytes, &scratch);
		}
		if (rc == FMB_OK) {
			L.scratch = (double*)scratch;
			// FMB_LMM_VARIANT=shuffle: the experimental lane-per-rate kernel (warp-shuffle prefix sums, fmb_euler_lmm_shuffle.cu) for A/B runs
			const char* variant = getenv("FMB_LMM_VARIANT");
			int vr = FMB_EUNSUPPORTED;
			if (variant && strcmp(variant, "shuffle") == 0) vr = lmmLaunchLanePerRate(L, c.smCount);
			rc = vr == FMB_EUNSUPPORTED ? launcher(L, c.smCount, true) : vr;
			if (rc == FMB_OK) rc = launchCheck("euler_lmm");
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
