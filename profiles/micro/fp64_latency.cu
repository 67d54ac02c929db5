// Micro-benchmark (profiling aid, not part of the library): FP64 pipe latency / throughput on B200 as a function of
// warps per SM sub-partition and independent chains per thread.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP> __global__ void chain(double* out, int iters, double a, double b, long long* cyc) {
	double x[ILP];
#pragma unroll
	for (int i = 0; i < ILP; i++) x[i] = threadIdx.x + i;
	long long t0 = clock64();
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int i = 0; i < ILP; i++) x[i] = fma(x[i], a, b);
	}
	long long t1 = clock64();
	double s = 0;
#pragma unroll
	for (int i = 0; i < ILP; i++) s += x[i];
	if (s == 1.2345) out[0] = s;
	if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP> void run(int warpsPerSM) {
	double* d; long long* c; cudaMalloc(&d, 8); cudaMalloc(&c, 8);
	int iters = 4096;
	int threads = warpsPerSM * 32;
	chain<ILP><<<148, threads>>>(d, iters, 0.999, 1e-3, c);
	chain<ILP><<<148, threads>>>(d, iters, 0.999, 1e-3, c);
	long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
	double cycPerInstrPerWarp = (double)h / (iters * ILP);
	double warpInstrPerCycPerSM = (double)iters * ILP * warpsPerSM / h;
	printf("warps/SM %2d ILP %d: %.2f cycles per DFMA per warp; %.3f warp-DFMA/cycle/SM (peak 2.0)\n", warpsPerSM, ILP, cycPerInstrPerWarp, warpInstrPerCycPerSM);
	cudaFree(d); cudaFree(c);
}
int main() {
	for (int w : {1, 4, 8, 16, 20, 32}) { run<1>(w); run<2>(w); run<4>(w); run<8>(w); }
	return 0;
}
