// Roofline denominators measured on the box itself: FP64 pipe peak (DFMA) and device copy bandwidth.
// MEASURED_PEAKS.json has no FP64 entry, and the Euler kernels are FP64-pipe bound.
#include "fmb_common.cuh"

namespace fmb {

__global__ void __launch_bounds__(256) dfmaKernel(double* out, int iters, double a, double b) {
	double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
	for (int i = 0; i < iters; i++) {
		x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
		x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
	}
	const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
	if (s == 123.456) out[0] = s;                  // keep the chains alive
}

__global__ void __launch_bounds__(256) copyKernel(const double2* __restrict__ in, double2* __restrict__ out, uint64_t n2) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n2; i += stride) out[i] = in[i];
}

} // namespace fmb

using namespace fmb;

extern "C" {

int fmb_bench_dfma_tflops(double* tflops) {
	FMB_TRY(requireInit());
	if (!tflops) return FMB_EINVAL;
	Context& c = ctx();
	void* d;
	FMB_TRY(poolAlloc(64, &d));
	const int iters = 1 << 14, grid = c.smCount * 8;
	cudaEvent_t e0, e1;
	FMB_CUDA(cudaEventCreate(&e0));
	FMB_CUDA(cudaEventCreate(&e1));
	double best = 0.0;
	for (int rep = 0; rep < 5; rep++) {
		FMB_CUDA(cudaEventRecord(e0, c.stream));
		dfmaKernel<<<grid, 256, 0, c.stream>>>((double*)d, iters, 0.999999, 1e-9);
		countLaunch();
		FMB_CUDA(cudaEventRecord(e1, c.stream));
		FMB_CUDA(cudaEventSynchronize(e1));
		float ms;
		FMB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
		const double flops = 2.0 * 8.0 * iters * (double)grid * 256.0;
		if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
	}
	cudaEventDestroy(e0); cudaEventDestroy(e1);
	poolFree(d, 64);
	*tflops = best;
	return FMB_OK;
}

int fmb_bench_copy_gbs(uint64_t bytes, double* gbs) {
	FMB_TRY(requireInit());
	if (!gbs || bytes < 32) return FMB_EINVAL;
	Context& c = ctx();
	bytes &= ~(uint64_t)15;
	void *a, *b;
	FMB_TRY(poolAlloc(bytes, &a));
	FMB_TRY(poolAlloc(bytes, &b));
	FMB_CUDA(cudaMemsetAsync(a, 0, bytes, c.stream));
	cudaEvent_t e0, e1;
	FMB_CUDA(cudaEventCreate(&e0));
	FMB_CUDA(cudaEventCreate(&e1));
	double best = 0.0;
	for (int rep = 0; rep < 6; rep++) {
		FMB_CUDA(cudaEventRecord(e0, c.stream));
		copyKernel<<<c.smCount * 16, 256, 0, c.stream>>>((const double2*)a, (double2*)b, bytes / 16);
		countLaunch();
		FMB_CUDA(cudaEventRecord(e1, c.stream));
		FMB_CUDA(cudaEventSynchronize(e1));
		float ms;
		FMB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
		if (rep > 0) best = std::max(best, 2.0 * bytes / (ms * 1e-3) / 1e9);
	}
	cudaEventDestroy(e0); cudaEventDestroy(e1);
	poolFree(a, bytes); poolFree(b, bytes);
	*gbs = best;
	return FMB_OK;
}

} // extern "C"
