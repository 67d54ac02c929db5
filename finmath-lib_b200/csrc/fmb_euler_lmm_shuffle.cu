// EXPERIMENT (A/B only, selected with FMB_LMM_VARIANT=shuffle): the lane-per-RATE layout the north star sketches for the LMM drift -
// "a per-path prefix sum over the forward-rate index done with warp shuffles".  A warp owns a tile of 32 paths and walks through them one
// path at a time; within a path the lanes are the live forward rates j = first .. N-1 (two slots of 32 when more than 32 are live): every
// lane evaluates log / reciprocal / exp of ITS rate, the running factor sums S_k(j) = sum_{i <= j} a_i fl_ik are inclusive warp scans
// (__shfl_up, 5 steps per factor), and the new rates go through a [rate][path] shared-memory tile so that the stores X[t+1][j][path] stay
// coalesced (lanes = paths in the flush).  The scan changes the summation tree of S_k (the reference adds j ascending): results agree with
// the production kernel to ~1e-15, not bit for bit.
// Restricted to what the A/B needs: spot measure, log-normal, EULER_FUNCTIONAL, F = 3, N <= 64, finite positive rates.
// Measured against eulerLmmKernel<3,1,0,1,0> on C4 in profiles/r02_notes.md; this file is not on the default path.
#include "fmb_euler_lmm.cuh"

namespace fmb {

static const int LPR_WARPS = 4;           // warps per CTA; each has its own state / tile

__global__ void __launch_bounds__(32 * LPR_WARPS) eulerLmmLanePerRateKernel(LmmParams q, uint64_t P, const double* const* __restrict__ dW, int NP) {
	extern __shared__ double sm[];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int N = q.N;
	// per warp: state L[32 paths][NP] | tile [N][33] | increments w[3][32]
	double* Lst = sm + (size_t)warp * (32 * NP + N * 33 + 96);
	double* tile = Lst + 32 * NP;
	double* wsm = tile + N * 33;
	const uint64_t tiles = (P + 31) / 32;
	const int RS = 6;                      // doubles per (t, j) record for F = 3: invv, hv, row pointer, fl[3]
	for (;;) {
		unsigned long long tileIdx = 0;
		if (lane == 0) tileIdx = atomicAdd(q.tileCounter, 1ull);
		tileIdx = __shfl_sync(0xffffffffu, tileIdx, 0);
		if (tileIdx >= tiles) break;
		const uint64_t p0 = tileIdx * 32;
		const int nPaths = (int)min((uint64_t)32, P - p0);
		for (int i = lane; i < 32 * N; i += 32) Lst[(i / N) * NP + (i % N)] = q.x0[i % N];
		__syncwarp();
		for (int t = 0; t < q.T; t++) {
			const int first = q.firstLive[t];
			const int live = N - first;
			if (live <= 0) continue;
			const double d = q.dt[t];
			if (lane < nPaths) {
#pragma unroll
				for (int k = 0; k < 3; k++) wsm[k * 32 + lane] = dW[(size_t)t * 3 + k][p0 + lane];
			}
			__syncwarp();
			const int slots = (live + 31) / 32;
			for (int pl = 0; pl < nPaths; pl++) {
				double carry0 = 0.0, carry1 = 0.0, carry2 = 0.0;
				const double w0 = wsm[pl], w1 = wsm[32 + pl], w2 = wsm[64 + pl];
				for (int s = 0; s < slots; s++) {
					const int j = first + s * 32 + lane;
					const bool active = j < N;
					const int jj = active ? j : N - 1;
					const double* r = q.rec + ((size_t)t * N + jj) * RS;
					const double invv = __ldg(r), hv = __ldg(r + 1), fl0 = __ldg(r + 3), fl1 = __ldg(r + 4), fl2 = __ldg(r + 5);
					const double L = Lst[pl * NP + jj];
					double y = (t == 0) ? q.ylog0[jj] : flog(L);
					double a = (1.0 / (L + invv)) * L;
					if (!active) a = 0.0;
					// inclusive scans of a * fl_k over the lanes (rates ascending)
					double s0 = a * fl0, s1 = a * fl1, s2 = a * fl2;
#pragma unroll
					for (int o = 1; o < 32; o <<= 1) {
						const double u0 = __shfl_up_sync(0xffffffffu, s0, o), u1 = __shfl_up_sync(0xffffffffu, s1, o), u2 = __shfl_up_sync(0xffffffffu, s2, o);
						if (lane >= o) { s0 = s0 + u0; s1 = s1 + u1; s2 = s2 + u2; }
					}
					s0 = s0 + carry0; s1 = s1 + carry1; s2 = s2 + carry2;
					carry0 = __shfl_sync(0xffffffffu, s0, 31); carry1 = __shfl_sync(0xffffffffu, s1, 31); carry2 = __shfl_sync(0xffffffffu, s2, 31);
					double mu = s0 * fl0 + 0.0;
					mu = s1 * fl1 + mu;
					mu = s2 * fl2 + mu;
					mu = mu + hv;
					y = mu * d + y;
					y = w0 * fl0 + y;
					y = w1 * fl1 + y;
					y = w2 * fl2 + y;
					double Ln = fexp(y);
					Ln = (Ln > q.cap) ? q.cap : Ln;
					if (active) { Lst[pl * NP + j] = Ln; tile[j * 33 + pl] = Ln; }
				}
			}
			__syncwarp();
			// flush: lanes = paths, one coalesced row segment per live rate
			for (int j = first; j < N; j++) {
				const double* r = q.rec + ((size_t)t * N + j) * RS;
				double* row = reinterpret_cast<double*>(__double_as_longlong(__ldg(r + 2)));
				if (lane < nPaths) row[p0 + lane] = tile[j * 33 + lane];
			}
			__syncwarp();
		}
	}
}

// returns FMB_EUNSUPPORTED when the configuration is outside the experiment's scope (the caller then uses the production kernel)
int lmmLaunchLanePerRate(LmmLaunch& a, int smCount) {
	const LmmParams& q = a.q;
	if (!(q.F == 3 && a.logn && a.spot && a.mode == 0 && !a.fast && q.N <= 64 && q.hasCap != 1)) return FMB_EUNSUPPORTED;
	const int NP = q.N | 1;
	const size_t smem = (size_t)LPR_WARPS * (32 * NP + q.N * 33 + 96) * sizeof(double);
	FMB_CUDA(cudaFuncSetAttribute(eulerLmmLanePerRateKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	int perSm = 0;
	FMB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, eulerLmmLanePerRateKernel, 32 * LPR_WARPS, smem));
	const uint64_t tiles = (a.paths + 31) / 32;
	const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)smCount * std::max(perSm, 1), (tiles + LPR_WARPS - 1) / LPR_WARPS));
	eulerLmmLanePerRateKernel<<<grid, 32 * LPR_WARPS, smem, a.stream>>>(q, a.paths, a.dW, NP);
	return FMB_OK;
}

} // namespace fmb
