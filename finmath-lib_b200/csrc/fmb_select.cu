// Order statistics without a sort and without moving path data between GPUs.
//
// getQuantile / getQuantileExpectation / getHistogram of the reference clone and SORT the vector
// (J/montecarlo/RandomVariableFromDoubleArray.java:445-575).  All three only need
//   * the element of a given rank in the sorted order            -> MSB-first radix SELECT on order-preserving 64-bit keys (8 passes of 8 bits:
//                                                                   a 256-bin histogram of the elements that still match the prefix, then the bin
//                                                                   holding the rank),
//   * counts of elements <= given thresholds                      -> one pass, binary search in the (sorted) thresholds per element,
//   * the sum of the elements strictly between two values + counts -> one pass in double-double.
// Every pass streams the vector once (HBM-bound: 8 B per element), and with path shards the only exchange is the 256-bin histogram
// (or the handful of counts / partial sums) through the communicator of fmb_comm.cu - the shards themselves never move.
// Order: the one of Arrays.sort(double[]): -0.0 < +0.0, NaN above everything (all NaN equal).
#include "fmb_common.cuh"
#include <cmath>
#include <algorithm>

namespace fmb {

int commAllGather(int count);                      // fmb_comm.cu

// order-preserving key of a double (Double.compare order): negative values reversed, NaN canonical at the top
__device__ __forceinline__ unsigned long long orderKey(double x) {
	unsigned long long b = (unsigned long long)__double_as_longlong(x);
	if (x != x) return 0xffffffffffffffffull;
	return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double keyToDouble(unsigned long long k) {
	if (k == 0xffffffffffffffffull) return NAN;
	const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
	return __longlong_as_double((long long)b);
}

struct SelectState { unsigned long long prefix, rank; };

static const int SEL_THREADS = 256;

// histogram over the next 8 key bits of the elements whose leading 8*pass bits equal state->prefix
__global__ void __launch_bounds__(SEL_THREADS) selectHistogramKernel(const double* __restrict__ x, uint64_t n, int pass, const SelectState* __restrict__ state,
		unsigned long long* __restrict__ hist) {
	__shared__ unsigned int sh[256];
	sh[threadIdx.x] = 0;
	__syncthreads();
	const unsigned long long prefix = state->prefix;
	const int shift = 56 - 8 * pass;
	const uint64_t stride = (uint64_t)gridDim.x * SEL_THREADS;
	for (uint64_t i = blockIdx.x * (uint64_t)SEL_THREADS + threadIdx.x; i < n; i += stride) {
		const unsigned long long k = orderKey(x[i]);
		if (pass == 0 || (k >> (shift + 8)) == prefix) atomicAdd(&sh[(unsigned int)(k >> shift) & 255u], 1u);
	}
	__syncthreads();
	if (sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
}

// one CTA of 256 threads: sum the shards' histograms, find the bin that holds the rank, extend the prefix, clear the local histogram.
// After the last pass the prefix is the key of the selected element: out[0] = its value.
__global__ void __launch_bounds__(256) selectPickKernel(const unsigned long long* __restrict__ hists /* [world][256] */, int world, int pass, SelectState* state,
		unsigned long long* __restrict__ histLocal, double* __restrict__ out) {
	__shared__ unsigned long long cnt[256];
	unsigned long long c = 0;
	for (int r = 0; r < world; r++) c += hists[(size_t)r * 256 + threadIdx.x];
	cnt[threadIdx.x] = c;
	__syncthreads();
	if (threadIdx.x == 0) {
		unsigned long long rank = state->rank, before = 0;
		int bin = 255;
		for (int b = 0; b < 256; b++) {
			if (rank < before + cnt[b]) { bin = b; break; }
			before += cnt[b];
		}
		state->rank = rank - before;
		state->prefix = (state->prefix << 8) | (unsigned long long)bin;
		if (pass == 7) { out[0] = keyToDouble(state->prefix); out[1] = 0.0; }
	}
	histLocal[threadIdx.x] = 0;
}

__global__ void selectInitKernel(SelectState* state, unsigned long long rank, unsigned long long* hist) {
	if (threadIdx.x == 0) { state->prefix = 0; state->rank = rank; }
	hist[threadIdx.x] = 0;
}

// counts[j] += number of elements x with sortedPts[j-1] < x <= sortedPts[j] (bucket j; bucket npts = everything above, NaN included)
__global__ void __launch_bounds__(SEL_THREADS) bucketCountKernel(const double* __restrict__ x, uint64_t n, const double* __restrict__ pts, int npts,
		unsigned long long* __restrict__ counts) {
	extern __shared__ unsigned char shraw[];
	double* sp = reinterpret_cast<double*>(shraw);
	unsigned int* sc = reinterpret_cast<unsigned int*>(sp + npts);
	for (int j = threadIdx.x; j < npts; j += SEL_THREADS) sp[j] = pts[j];
	for (int j = threadIdx.x; j <= npts; j += SEL_THREADS) sc[j] = 0;
	__syncthreads();
	const uint64_t stride = (uint64_t)gridDim.x * SEL_THREADS;
	for (uint64_t i = blockIdx.x * (uint64_t)SEL_THREADS + threadIdx.x; i < n; i += stride) {
		const double v = x[i];
		int lo = 0, hi = npts;                     // first j with v <= sp[j]  (NaN: never, -> npts)
		while (lo < hi) {
			const int mid = (lo + hi) >> 1;
			if (v <= sp[mid]) hi = mid; else lo = mid + 1;
		}
		atomicAdd(&sc[lo], 1u);
	}
	__syncthreads();
	for (int j = threadIdx.x; j <= npts; j += SEL_THREADS) if (sc[j]) atomicAdd(&counts[j], (unsigned long long)sc[j]);
}

// out (as (hi, lo) pairs): sum of the elements with lo < x < hi in double-double; count of x <= lo; count of x < hi
struct dd2 { double hi, lo; };
__device__ __forceinline__ void twoSum2(double a, double b, double& s, double& e) { s = a + b; const double bb = s - a; e = (a - (s - bb)) + (b - bb); }
__device__ __forceinline__ void ddMerge2(dd2& a, const dd2& b) { double s, e; twoSum2(a.hi, b.hi, s, e); e += a.lo + b.lo; twoSum2(s, e, a.hi, a.lo); }

__global__ void __launch_bounds__(SEL_THREADS) rangeSumKernel(const double* __restrict__ x, uint64_t n, double lo, double hi, double* __restrict__ partials /* [grid][4] */,
		unsigned int* ticket, double* __restrict__ out6) {
	dd2 acc = {0.0, 0.0};
	unsigned long long cLe = 0, cLt = 0;
	const unsigned long long kLo = orderKey(lo), kHi = orderKey(hi);
	const uint64_t stride = (uint64_t)gridDim.x * SEL_THREADS;
	for (uint64_t i = blockIdx.x * (uint64_t)SEL_THREADS + threadIdx.x; i < n; i += stride) {
		const double v = x[i];
		const unsigned long long k = orderKey(v);            // (key comparisons: -0.0 < +0.0 and NaN on top, like the sort the reference uses)
		if (k <= kLo) cLe++;
		if (k < kHi) cLt++;
		if (k > kLo && k < kHi) { double s, e; twoSum2(acc.hi, v, s, e); acc.hi = s; acc.lo += e; }
	}
	__shared__ dd2 shs[SEL_THREADS];
	__shared__ unsigned long long sha[SEL_THREADS], shb[SEL_THREADS];
	shs[threadIdx.x] = acc; sha[threadIdx.x] = cLe; shb[threadIdx.x] = cLt;
	__syncthreads();
	for (int d = SEL_THREADS / 2; d > 0; d >>= 1) {
		if ((int)threadIdx.x < d) { ddMerge2(shs[threadIdx.x], shs[threadIdx.x + d]); sha[threadIdx.x] += sha[threadIdx.x + d]; shb[threadIdx.x] += shb[threadIdx.x + d]; }
		__syncthreads();
	}
	if (threadIdx.x == 0) {
		double* p = partials + 4 * (size_t)blockIdx.x;
		p[0] = shs[0].hi; p[1] = shs[0].lo; p[2] = (double)sha[0]; p[3] = (double)shb[0];      // counts < 2^53: exact
	}
	__shared__ bool isLast;
	__threadfence();
	__syncthreads();
	if (threadIdx.x == 0) isLast = (atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1);
	__syncthreads();
	if (!isLast) return;
	__threadfence();
	if (threadIdx.x == 0) {                        // the grid is small (<= 4 CTAs per SM): a sequential, fixed-order merge
		dd2 t = {0.0, 0.0};
		double a = 0.0, b = 0.0;
		for (unsigned int g = 0; g < gridDim.x; g++) {
			const dd2 o = { __ldcg(partials + 4 * g), __ldcg(partials + 4 * g + 1) };
			ddMerge2(t, o);
			a += __ldcg(partials + 4 * g + 2); b += __ldcg(partials + 4 * g + 3);
		}
		out6[0] = t.hi; out6[1] = t.lo; out6[2] = a; out6[3] = 0.0; out6[4] = b; out6[5] = 0.0;
	}
}

// sum over the shards of `count` gathered doubles (plain counts: exact)
__global__ void sumShardsKernel(const double* __restrict__ gathered, int world, int count, double* __restrict__ out) {
	const int j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= count) return;
	double s = 0.0;
	for (int r = 0; r < world; r++) s += gathered[(size_t)r * count + j];
	out[j] = s;
}
__global__ void mergeDdShardsKernel(const double* __restrict__ gathered, int world, int pairs, double* __restrict__ out) {
	const int m = threadIdx.x;
	if (m >= pairs) return;
	dd2 t = { gathered[2 * m], gathered[2 * m + 1] };
	for (int r = 1; r < world; r++) { const dd2 o = { gathered[(size_t)r * 2 * pairs + 2 * m], gathered[(size_t)r * 2 * pairs + 2 * m + 1] }; ddMerge2(t, o); }
	out[2 * m] = t.hi; out[2 * m + 1] = t.lo;
}
__global__ void u64ToDoubleKernel(const unsigned long long* __restrict__ in, double* __restrict__ out, int count) {
	const int j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j < count) out[j] = (double)in[j];
}

static int selGrid(uint64_t n) { return (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)ctx().smCount * 4, (n + SEL_THREADS - 1) / SEL_THREADS)); }

} // namespace fmb

using namespace fmb;

extern "C" {

// element of rank `rank` (0-based) in the sorted order of the logical vector (all shards when a communicator is active)
int fmb_rv_select(fmb_handle x, uint64_t rank, double* out) {
	FMB_TRY(requireInit());
	PinScope pins;
	if (!out) return FMB_EINVAL;
	Context& c = ctx();
	Vec* vx;
	FMB_TRY(lookup(x, &vx));
	const uint64_t n = vx->n;
	if (!c.comm.active && rank >= n) { setError("select: rank %llu out of range (size %llu)", (unsigned long long)rank, (unsigned long long)n); return FMB_EINVAL; }
	std::lock_guard<std::mutex> lk(c.scratchMu);
	// scratch: [0, 256) local histogram (u64) | state
	FMB_TRY(ensureScratch(0, 512 * sizeof(unsigned long long)));
	unsigned long long* hist = (unsigned long long*)c.scratch;
	SelectState* state = (SelectState*)(hist + 256);
	const bool sharded = c.comm.active;
	// with shards the histogram is accumulated straight into the communicator's send buffer
	unsigned long long* histBuf = sharded ? (unsigned long long*)c.comm.sendBuf : hist;
	selectInitKernel<<<1, 256, 0, c.stream>>>(state, rank, histBuf);
	const int grid = selGrid(n);
	for (int pass = 0; pass < 8; pass++) {
		if (n) selectHistogramKernel<<<grid, SEL_THREADS, 0, c.stream>>>(vx->ptr, n, pass, state, histBuf);
		if (sharded) FMB_TRY(commAllGather(256));
		selectPickKernel<<<1, 256, 0, c.stream>>>(sharded ? (const unsigned long long*)c.comm.gatherBuf : hist, sharded ? c.comm.world : 1, pass, state, histBuf,
		                                         c.hostResultDev);
	}
	countLaunch(17);
	FMB_CUDA(cudaGetLastError());
	FMB_CUDA(cudaStreamSynchronize(c.stream));
	*out = c.hostResult[0];
	return FMB_OK;
}

// counts[j] = number of elements <= pts[j] over the logical vector (any order of pts; NaN elements are never counted)
int fmb_rv_count_le(fmb_handle x, const double* pts, int npts, uint64_t* counts) {
	FMB_TRY(requireInit());
	PinScope pins;
	if (npts < 0 || (npts && (!pts || !counts))) return FMB_EINVAL;
	if (npts == 0) return FMB_OK;
	if (npts > COMM_MAX_DOUBLES - 1) { setError("count_le: at most %d thresholds per call", COMM_MAX_DOUBLES - 1); return FMB_EUNSUPPORTED; }
	Context& c = ctx();
	Vec* vx;
	FMB_TRY(lookup(x, &vx));
	// sorted copy of the thresholds (NaN thresholds count nothing: kept out of the search)
	std::vector<std::pair<double, int>> order;
	for (int j = 0; j < npts; j++) if (pts[j] == pts[j]) order.push_back({pts[j], j});
	std::sort(order.begin(), order.end());
	const int m = (int)order.size();
	std::lock_guard<std::mutex> lk(c.scratchMu);
	FMB_TRY(ensureScratch((size_t)(npts + 1) * sizeof(double), (size_t)(2 * npts + 4) * sizeof(double)));
	double* hp = (double*)c.pinned;
	for (int j = 0; j < m; j++) hp[j] = order[j].first;
	double* dpts = (double*)c.scratch;
	unsigned long long* dcnt = (unsigned long long*)(dpts + npts + 1);
	for (int j = 0; j < npts; j++) counts[j] = 0;
	if (m > 0) {
		FMB_CUDA(cudaMemcpyAsync(dpts, hp, (size_t)m * sizeof(double), cudaMemcpyHostToDevice, c.stream));
		FMB_CUDA(cudaMemsetAsync(dcnt, 0, (size_t)(m + 1) * sizeof(unsigned long long), c.stream));
		if (vx->n) bucketCountKernel<<<selGrid(vx->n), SEL_THREADS, (size_t)m * sizeof(double) + (size_t)(m + 1) * sizeof(unsigned int), c.stream>>>(vx->ptr, vx->n, dpts, m, dcnt);
		double* res = c.comm.active ? c.comm.sendBuf : c.hostResultDev;
		u64ToDoubleKernel<<<(m + 127) / 128, 128, 0, c.stream>>>(dcnt, res, m);
		countLaunch(2);
		if (c.comm.active) {
			FMB_TRY(commAllGather(m));
			sumShardsKernel<<<(m + 127) / 128, 128, 0, c.stream>>>(c.comm.gatherBuf, c.comm.world, m, c.hostResultDev);
			countLaunch();
		}
		FMB_CUDA(cudaGetLastError());
		FMB_CUDA(cudaStreamSynchronize(c.stream));
		uint64_t cum = 0;
		for (int j = 0; j < m; j++) { cum += (uint64_t)c.hostResult[j]; counts[order[j].second] = cum; }
	}
	return FMB_OK;
}

// out[0] + out[1] = sum of the elements strictly between lo and hi (sort order), out[2] = #{x <= lo}, out[3] = #{x < hi}, all shards
int fmb_rv_range_sum(fmb_handle x, double lo, double hi, double* out4) {
	FMB_TRY(requireInit());
	PinScope pins;
	if (!out4) return FMB_EINVAL;
	Context& c = ctx();
	Vec* vx;
	FMB_TRY(lookup(x, &vx));
	const uint64_t n = vx->n;
	const int grid = selGrid(n);
	std::lock_guard<std::mutex> lk(c.scratchMu);
	FMB_TRY(ensureScratch(0, (size_t)grid * 4 * sizeof(double)));
	double* dst = c.comm.active ? c.comm.sendBuf : c.hostResultDev;
	if (n == 0) FMB_CUDA(cudaMemsetAsync(dst, 0, 6 * sizeof(double), c.stream));
	else rangeSumKernel<<<grid, SEL_THREADS, 0, c.stream>>>(vx->ptr, n, lo, hi, (double*)c.scratch, c.ticket, dst);
	countLaunch();
	if (c.comm.active) {
		FMB_TRY(commAllGather(6));
		mergeDdShardsKernel<<<1, 32, 0, c.stream>>>(c.comm.gatherBuf, c.comm.world, 3, c.hostResultDev);
		countLaunch();
	}
	FMB_CUDA(cudaGetLastError());
	FMB_CUDA(cudaStreamSynchronize(c.stream));
	out4[0] = c.hostResult[0]; out4[1] = c.hostResult[1];
	out4[2] = c.hostResult[2] + c.hostResult[3];
	out4[3] = c.hostResult[4] + c.hostResult[5];
	return FMB_OK;
}

} // extern "C"
