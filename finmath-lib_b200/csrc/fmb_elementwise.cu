// Element-wise RandomVariable arithmetic on device-resident vectors.
// Semantics: J/montecarlo/RandomVariableFromDoubleArray.java:742-1504 (one streaming pass per op there; one kernel here).
// HBM-bound: 8 B read per vector operand + 8 B written per element.  Grid-stride, two elements per thread per trip,
// 16-byte vector accesses when every pointer is 16-byte aligned.  Compiled with -fmad=false: x + y*z is a rounded
// multiply followed by a rounded add, as on the JVM.
#include "fmb_common.cuh"
#include "fmb_math.cuh"
#include "fmb_icdf.cuh"

namespace fmb {

// Java Math.min / Math.max: NaN-propagating, -0.0 < +0.0 (CUDA fmin/fmax drop NaNs, so written out)
__device__ __forceinline__ double jmin(double a, double b) {
	if (a != a) return a;
	if (a == 0.0 && b == 0.0 && signbit(b)) return b;
	return (a <= b) ? a : b;
}
__device__ __forceinline__ double jmax(double a, double b) {
	if (a != a) return a;
	if (a == 0.0 && b == 0.0 && signbit(a)) return b;
	return (a >= b) ? a : b;
}

template <int OP> __device__ __forceinline__ double unaryOp(double x, double a) {
	switch (OP) {
	case FMB_U_SQUARED: return x * x;
	case FMB_U_SQRT: return sqrt(x);
	case FMB_U_EXP: return fexp(x);
	case FMB_U_LOG: return flog(x);
	case FMB_U_SIN: return sin(x);
	case FMB_U_COS: return cos(x);
	case FMB_U_INVERT: return 1.0 / x;
	case FMB_U_ABS: return fabs(x);
	case FMB_U_ISNAN: return (x != x) ? 1.0 : 0.0;
	case FMB_U_EXPM1: return expm1(x);
	case FMB_U_ADD: return x + a;
	case FMB_U_SUB: return x - a;
	case FMB_U_BUS: return a - x;
	case FMB_U_MULT: return x * a;
	case FMB_U_DIV: return x / a;
	case FMB_U_VID: return a / x;
	case FMB_U_CAP: return jmin(x, a);
	case FMB_U_FLOOR: return jmax(x, a);
	case FMB_U_POW: return pow(x, a);
	case FMB_U_ICDF_NORMAL: return inverseCumulativeNormal(x);
	}
	return x;
}
template <int OP> __device__ __forceinline__ double binaryOp(double x, double y) {
	switch (OP) {
	case FMB_B_ADD: return x + y;
	case FMB_B_SUB: return x - y;
	case FMB_B_MULT: return x * y;
	case FMB_B_DIV: return x / y;
	case FMB_B_CAP: return jmin(x, y);
	case FMB_B_FLOOR: return jmax(x, y);
	}
	return x;
}
template <int OP> __device__ __forceinline__ double ternaryOp(double x, double y, double z, double a) {
	switch (OP) {
	case FMB_T_ADD_PRODUCT: return x + y * z;
	case FMB_T_ADD_PRODUCT_D: return x + y * a;
	case FMB_T_ADD_RATIO: return x + y / z;
	case FMB_T_SUB_RATIO: return x - y / z;
	case FMB_T_ACCRUE: return x * (1 + y * a);
	case FMB_T_DISCOUNT: return x / (1.0 + y * a);
	case FMB_T_CHOOSE: return x >= 0.0 ? y : z;
	}
	return x;
}

__device__ __forceinline__ double2 ld2(const double* p, uint64_t i, double s) {
	if (p) return *reinterpret_cast<const double2*>(p + i);
	return make_double2(s, s);
}
__device__ __forceinline__ double ld1(const double* p, uint64_t i, double s) { return p ? p[i] : s; }

template <int OP> __global__ void __launch_bounds__(256) unaryKernel(const double* __restrict__ x, double a, double* __restrict__ out, uint64_t n, int vec) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	if (vec) {
		const uint64_t n2 = n >> 1;
		for (; i < n2; i += stride) {
			const double2 v = *reinterpret_cast<const double2*>(x + 2 * i);
			*reinterpret_cast<double2*>(out + 2 * i) = make_double2(unaryOp<OP>(v.x, a), unaryOp<OP>(v.y, a));
		}
		if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) out[n - 1] = unaryOp<OP>(x[n - 1], a);
	} else {
		for (; i < n; i += stride) out[i] = unaryOp<OP>(x[i], a);
	}
}
template <int OP> __global__ void __launch_bounds__(256) binaryKernel(const double* __restrict__ x, double sx, const double* __restrict__ y, double sy,
		double* __restrict__ out, uint64_t n, int vec) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	if (vec) {
		const uint64_t n2 = n >> 1;
		for (; i < n2; i += stride) {
			const double2 a = ld2(x, 2 * i, sx), b = ld2(y, 2 * i, sy);
			*reinterpret_cast<double2*>(out + 2 * i) = make_double2(binaryOp<OP>(a.x, b.x), binaryOp<OP>(a.y, b.y));
		}
		if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) out[n - 1] = binaryOp<OP>(ld1(x, n - 1, sx), ld1(y, n - 1, sy));
	} else {
		for (; i < n; i += stride) out[i] = binaryOp<OP>(ld1(x, i, sx), ld1(y, i, sy));
	}
}
template <int OP> __global__ void __launch_bounds__(256) ternaryKernel(const double* __restrict__ x, double sx, const double* __restrict__ y, double sy,
		const double* __restrict__ z, double sz, double a, double* __restrict__ out, uint64_t n, int vec) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	if (vec) {
		const uint64_t n2 = n >> 1;
		for (; i < n2; i += stride) {
			const double2 p = ld2(x, 2 * i, sx), q = ld2(y, 2 * i, sy), r = ld2(z, 2 * i, sz);
			*reinterpret_cast<double2*>(out + 2 * i) = make_double2(ternaryOp<OP>(p.x, q.x, r.x, a), ternaryOp<OP>(p.y, q.y, r.y, a));
		}
		if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0)
			out[n - 1] = ternaryOp<OP>(ld1(x, n - 1, sx), ld1(y, n - 1, sy), ld1(z, n - 1, sz), a);
	} else {
		for (; i < n; i += stride) out[i] = ternaryOp<OP>(ld1(x, i, sx), ld1(y, i, sy), ld1(z, i, sz), a);
	}
}

__global__ void __launch_bounds__(256) fillKernel(double* __restrict__ out, double v, uint64_t n) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += stride) out[i] = v;
}

// ---------------------------------------------------------------------------------------------------------------
// Chains of element-wise operations in ONE pass (fmb_rv_eval_chain).  The host binding defers RandomVariable arithmetic: a result
// that is only consumed by the next operation never becomes a vector in HBM.  acc = leaf[start][i]; every instruction replaces acc by
// op(.., acc, ..) with acc at operand position `pos` and the other operands taken from leaf vectors or broadcast scalars.  The
// operations are the same device functions as the one-op kernels, applied in the same order with the same roundings (no
// contraction), so a chain is bit-identical to the sequence of single operations.
// ---------------------------------------------------------------------------------------------------------------
struct ChainInstr { unsigned char kind, op, pos, refA, refB, refC, pad0, pad1; };    // ref: bit 7 set = scalars[ref & 127], else leaf[ref]
struct ChainProg {
	int n, start;
	ChainInstr code[FMB_CHAIN_MAX_INSTR];
	const double* leaf[FMB_CHAIN_MAX_LEAVES];
	double scalar[FMB_CHAIN_MAX_SCALARS];
};

__device__ __forceinline__ double unaryDyn(int op, double x, double a) {
	switch (op) {
	case FMB_U_SQUARED: return unaryOp<FMB_U_SQUARED>(x, a);
	case FMB_U_SQRT: return unaryOp<FMB_U_SQRT>(x, a);
	case FMB_U_EXP: return unaryOp<FMB_U_EXP>(x, a);
	case FMB_U_LOG: return unaryOp<FMB_U_LOG>(x, a);
	case FMB_U_SIN: return unaryOp<FMB_U_SIN>(x, a);
	case FMB_U_COS: return unaryOp<FMB_U_COS>(x, a);
	case FMB_U_INVERT: return unaryOp<FMB_U_INVERT>(x, a);
	case FMB_U_ABS: return unaryOp<FMB_U_ABS>(x, a);
	case FMB_U_ISNAN: return unaryOp<FMB_U_ISNAN>(x, a);
	case FMB_U_EXPM1: return unaryOp<FMB_U_EXPM1>(x, a);
	case FMB_U_ADD: return unaryOp<FMB_U_ADD>(x, a);
	case FMB_U_SUB: return unaryOp<FMB_U_SUB>(x, a);
	case FMB_U_BUS: return unaryOp<FMB_U_BUS>(x, a);
	case FMB_U_MULT: return unaryOp<FMB_U_MULT>(x, a);
	case FMB_U_DIV: return unaryOp<FMB_U_DIV>(x, a);
	case FMB_U_VID: return unaryOp<FMB_U_VID>(x, a);
	case FMB_U_CAP: return unaryOp<FMB_U_CAP>(x, a);
	case FMB_U_FLOOR: return unaryOp<FMB_U_FLOOR>(x, a);
	case FMB_U_ICDF_NORMAL: return unaryOp<FMB_U_ICDF_NORMAL>(x, a);
	default: return unaryOp<FMB_U_POW>(x, a);
	}
}
__device__ __forceinline__ double binaryDyn(int op, double x, double y) {
	switch (op) {
	case FMB_B_ADD: return binaryOp<FMB_B_ADD>(x, y);
	case FMB_B_SUB: return binaryOp<FMB_B_SUB>(x, y);
	case FMB_B_MULT: return binaryOp<FMB_B_MULT>(x, y);
	case FMB_B_DIV: return binaryOp<FMB_B_DIV>(x, y);
	case FMB_B_CAP: return binaryOp<FMB_B_CAP>(x, y);
	default: return binaryOp<FMB_B_FLOOR>(x, y);
	}
}
__device__ __forceinline__ double ternaryDyn(int op, double x, double y, double z, double a) {
	switch (op) {
	case FMB_T_ADD_PRODUCT: return ternaryOp<FMB_T_ADD_PRODUCT>(x, y, z, a);
	case FMB_T_ADD_PRODUCT_D: return ternaryOp<FMB_T_ADD_PRODUCT_D>(x, y, z, a);
	case FMB_T_ADD_RATIO: return ternaryOp<FMB_T_ADD_RATIO>(x, y, z, a);
	case FMB_T_SUB_RATIO: return ternaryOp<FMB_T_SUB_RATIO>(x, y, z, a);
	case FMB_T_ACCRUE: return ternaryOp<FMB_T_ACCRUE>(x, y, z, a);
	case FMB_T_DISCOUNT: return ternaryOp<FMB_T_DISCOUNT>(x, y, z, a);
	default: return ternaryOp<FMB_T_CHOOSE>(x, y, z, a);
	}
}

// Two elements per thread.  Phase 1 issues the loads of ALL leaf vectors back to back (a fetch inside the interpreter loop would wait
// ~1 us for HBM once per instruction) and parks the values in the thread's own shared-memory column, where the interpreter can index
// them dynamically; phase 2 walks the instructions.  No barrier: a column is private to its thread.  (Four elements per thread measured
// slower: 124 registers, and the dispatch is not the dominant cost - profiles/r01_notes.md.)
__global__ void __launch_bounds__(256) chainKernel(const __grid_constant__ ChainProg p, double* __restrict__ out, uint64_t n, int nLeaves, int vec) {
	__shared__ double2 col[FMB_CHAIN_MAX_LEAVES][256];
	const int tid = threadIdx.x;
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	const uint64_t pairs = (n + 1) >> 1;
	for (uint64_t i2 = blockIdx.x * (uint64_t)blockDim.x + tid; i2 < pairs; i2 += stride) {
		const uint64_t i = 2 * i2;
		const bool two = i + 1 < n;
		double2 v[FMB_CHAIN_MAX_LEAVES];
#pragma unroll
		for (int l = 0; l < FMB_CHAIN_MAX_LEAVES; l++) {
			if (l < nLeaves) {
				if (two && vec) v[l] = *reinterpret_cast<const double2*>(p.leaf[l] + i);
				else v[l] = make_double2(p.leaf[l][i], two ? p.leaf[l][i + 1] : 0.0);
			}
		}
#pragma unroll
		for (int l = 0; l < FMB_CHAIN_MAX_LEAVES; l++) if (l < nLeaves) col[l][tid] = v[l];
		double2 acc = col[p.start][tid];
		for (int k = 0; k < p.n; k++) {
			const ChainInstr c = p.code[k];
			if (c.kind == 0) {
				const double a = p.scalar[c.refA & 127];
				acc.x = unaryDyn(c.op, acc.x, a);
				acc.y = unaryDyn(c.op, acc.y, a);
			} else {
				double2 u;
				if (c.refA & 128) { const double sv = p.scalar[c.refA & 127]; u = make_double2(sv, sv); } else u = col[c.refA][tid];
				if (c.kind == 1) {
					if (c.pos == 0) { acc.x = binaryDyn(c.op, acc.x, u.x); acc.y = binaryDyn(c.op, acc.y, u.y); }
					else { acc.x = binaryDyn(c.op, u.x, acc.x); acc.y = binaryDyn(c.op, u.y, acc.y); }
				} else {
					double2 w;
					if (c.refB & 128) { const double sv = p.scalar[c.refB & 127]; w = make_double2(sv, sv); } else w = col[c.refB][tid];
					const double a = p.scalar[c.refC & 127];
					if (c.pos == 0) { acc.x = ternaryDyn(c.op, acc.x, u.x, w.x, a); acc.y = ternaryDyn(c.op, acc.y, u.y, w.y, a); }
					else if (c.pos == 1) { acc.x = ternaryDyn(c.op, u.x, acc.x, w.x, a); acc.y = ternaryDyn(c.op, u.y, acc.y, w.y, a); }
					else { acc.x = ternaryDyn(c.op, u.x, w.x, acc.x, a); acc.y = ternaryDyn(c.op, u.y, w.y, acc.y, a); }
				}
			}
		}
		if (two && vec) *reinterpret_cast<double2*>(out + i) = acc;
		else { out[i] = acc.x; if (two) out[i + 1] = acc.y; }
	}
}

static inline int aligned16(const void* p) { return p == nullptr || (((uintptr_t)p) & 15) == 0; }

// grid sized as a multiple of the SM count (8 resident CTAs of 256 threads per SM), capped by the work
static inline int ewGrid(uint64_t n) {
	uint64_t want = (n / 2 + 255) / 256;
	uint64_t cap = (uint64_t)ctx().smCount * 8;
	if (want > cap) want = cap;
	if (want < 1) want = 1;
	return (int)want;
}

#define LAUNCH_UNARY(OPC) case OPC: unaryKernel<OPC><<<grid, 256, 0, c.stream>>>(xp, a, dst, n, vec); break;
#define LAUNCH_BINARY(OPC) case OPC: binaryKernel<OPC><<<grid, 256, 0, c.stream>>>(xp, sx, yp, sy, dst, n, vec); break;
#define LAUNCH_TERNARY(OPC) case OPC: ternaryKernel<OPC><<<grid, 256, 0, c.stream>>>(xp, sx, yp, sy, zp, sz, a, dst, n, vec); break;

static int finishLaunch(fmb_handle* out) {
	countLaunch();
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) {
		setError("element-wise kernel launch failed: %s", cudaGetErrorString(e));
		fmb_rv_free(*out);
		*out = 0;
		return FMB_ECUDA;
	}
	return FMB_OK;
}


// ---- multi-period forward rate: ((1 + L_a d_a)(1 + L_{a+1} d_{a+1}) ... - 1) / (T_b - T_a) in ONE pass.
// LIBORMarketModelFromCovarianceModel.getForwardRate (:1288-1302) builds it with one accrue() pass per period - up to 20 passes for the
// long rate of a Bermudan's basis functions, 210 over a valuation.  Same operations in the same order (l.mult(d).add(1.0), then
// accrue = x * (1 + y * a) per period, then sub(1.0).div(length)), so the result is bit-identical to the op-by-op evaluation.
struct AccrueChainArgs {
	const double* rate[32];
	double delta[32];
};
__global__ void __launch_bounds__(256) accrueChainKernel(AccrueChainArgs a, int n, const double* __restrict__ accIn, int finalize, double divisor,
		double* __restrict__ out, uint64_t len) {
	const uint64_t stride = (uint64_t)gridDim.x * 256;
	for (uint64_t i = blockIdx.x * (uint64_t)256 + threadIdx.x; i < len; i += stride) {
		double acc;
		int k = 0;
		if (accIn) acc = accIn[i];
		else { acc = a.rate[0][i] * a.delta[0] + 1.0; k = 1; }
		for (; k < n; k++) acc = acc * (1 + a.rate[k][i] * a.delta[k]);
		out[i] = finalize ? (acc - 1.0) / divisor : acc;
	}
}

// every prefix of the accrual chain in one pass: out_k = start * (1 + r_0 d_0) * ... * (1 + r_k d_k), k < n (the spot-measure numeraire at every
// tenor date, LIBORMarketModelFromCovarianceModel.java:1050-1069: each accrue() there is one more array pass)
struct AccruePrefixArgs {
	const double* rate[64];
	double delta[64];
};
__global__ void __launch_bounds__(256) accruePrefixKernel(const __grid_constant__ AccruePrefixArgs a, int n, double start, const double* __restrict__ accIn,
		double* __restrict__ out /* [n][len] */, uint64_t len) {
	const uint64_t stride = (uint64_t)gridDim.x * 256;
	for (uint64_t i = blockIdx.x * (uint64_t)256 + threadIdx.x; i < len; i += stride) {
		double acc = accIn ? accIn[i] : start;
		for (int k = 0; k < n; k++) {
			acc = acc * (1 + a.rate[k][i] * a.delta[k]);
			out[(size_t)k * len + i] = acc;
		}
	}
}

} // namespace fmb

using namespace fmb;

extern "C" {

int fmb_rv_fill(double value, uint64_t n, fmb_handle* out) {
	FMB_TRY(requireInit());
	PinScope pins;
	if (!out) return FMB_EINVAL;
	double* dst;
	FMB_TRY(newVec(n, out, &dst));
	if (n == 0) return FMB_OK;
	fillKernel<<<ewGrid(n), 256, 0, ctx().stream>>>(dst, value, n);
	return finishLaunch(out);
}

int fmb_rv_unary(int opcode, fmb_handle x, double a, fmb_handle* out) {
	FMB_TRY(requireInit());
	PinScope pins;
	if (!out) return FMB_EINVAL;
	if (opcode < 0 || opcode > FMB_U_ICDF_NORMAL) { setError("unknown unary op %d", opcode); return FMB_EINVAL; }
	Context& c = ctx();
	Vec* vx;
	FMB_TRY(lookup(x, &vx));
	const uint64_t n = vx->n;
	const double* xp = vx->ptr;
	// Math.pow(x, 0.5) / Math.pow(x, 2.0) must equal sqrt / x*x bit-for-bit (T/montecarlo/RandomVariableTest.java:101-127)
	if (opcode == FMB_U_POW && a == 0.5) opcode = FMB_U_SQRT;
	else if (opcode == FMB_U_POW && a == 2.0) opcode = FMB_U_SQUARED;
	double* dst;
	FMB_TRY(newVec(n, out, &dst));
	if (n == 0) return FMB_OK;
	const int vec = aligned16(xp) && aligned16(dst);
	const int grid = ewGrid(n);
	switch (opcode) {
		LAUNCH_UNARY(FMB_U_SQUARED) LAUNCH_UNARY(FMB_U_SQRT) LAUNCH_UNARY(FMB_U_EXP) LAUNCH_UNARY(FMB_U_LOG) LAUNCH_UNARY(FMB_U_SIN)
		LAUNCH_UNARY(FMB_U_COS) LAUNCH_UNARY(FMB_U_INVERT) LAUNCH_UNARY(FMB_U_ABS) LAUNCH_UNARY(FMB_U_ISNAN) LAUNCH_UNARY(FMB_U_EXPM1)
		LAUNCH_UNARY(FMB_U_ADD) LAUNCH_UNARY(FMB_U_SUB) LAUNCH_UNARY(FMB_U_BUS) LAUNCH_UNARY(FMB_U_MULT) LAUNCH_UNARY(FMB_U_DIV)
		LAUNCH_UNARY(FMB_U_VID) LAUNCH_UNARY(FMB_U_CAP) LAUNCH_UNARY(FMB_U_FLOOR) LAUNCH_UNARY(FMB_U_POW) LAUNCH_UNARY(FMB_U_ICDF_NORMAL)
	}
	return finishLaunch(out);
}

// length of the result = length of the vector operands (all must agree); at least one operand must be a vector
static int commonLength(const fmb_handle* hs, int cnt, uint64_t* n) {
	*n = 0;
	bool any = false;
	for (int i = 0; i < cnt; i++) {
		if (hs[i] == 0) continue;
		Vec* v;
		FMB_TRY(lookup(hs[i], &v));
		if (any && v->n != *n) { setError("operand sizes differ (%llu vs %llu)", (unsigned long long)v->n, (unsigned long long)*n); return FMB_EINVAL; }
		*n = v->n; any = true;
	}
	if (!any) { setError("all operands are scalars; deterministic arithmetic stays on the host"); return FMB_EINVAL; }
	return FMB_OK;
}

int fmb_rv_binary(int opcode, fmb_handle x, double sx, fmb_handle y, double sy, fmb_handle* out) {
	FMB_TRY(requireInit());
	PinScope pins;
	if (!out) return FMB_EINVAL;
	if (opcode < 0 || opcode > FMB_B_FLOOR) { setError("unknown binary op %d", opcode); return FMB_EINVAL; }
	Context& c = ctx();
	uint64_t n;
	const fmb_handle hs[2] = { x, y };
	FMB_TRY(commonLength(hs, 2, &n));
	const double *xp, *yp;
	FMB_TRY(lookupPtr(x, 0, &xp));
	FMB_TRY(lookupPtr(y, 0, &yp));
	double* dst;
	FMB_TRY(newVec(n, out, &dst));
	if (n == 0) return FMB_OK;
	const int vec = aligned16(xp) && aligned16(yp) && aligned16(dst);
	const int grid = ewGrid(n);
	switch (opcode) {
		LAUNCH_BINARY(FMB_B_ADD) LAUNCH_BINARY(FMB_B_SUB) LAUNCH_BINARY(FMB_B_MULT) LAUNCH_BINARY(FMB_B_DIV) LAUNCH_BINARY(FMB_B_CAP)
		LAUNCH_BINARY(FMB_B_FLOOR)
	}
	return finishLaunch(out);
}

int fmb_rv_ternary(int opcode, fmb_handle x, double sx, fmb_handle y, double sy, fmb_handle z, double sz, double a, fmb_handle* out) {
	FMB_TRY(requireInit());
	PinScope pins;
	if (!out) return FMB_EINVAL;
	if (opcode < 0 || opcode > FMB_T_CHOOSE) { setError("unknown ternary op %d", opcode); return FMB_EINVAL; }
	Context& c = ctx();
	uint64_t n;
	const fmb_handle hs[3] = { x, y, z };
	FMB_TRY(commonLength(hs, 3, &n));
	const double *xp, *yp, *zp;
	FMB_TRY(lookupPtr(x, 0, &xp));
	FMB_TRY(lookupPtr(y, 0, &yp));
	FMB_TRY(lookupPtr(z, 0, &zp));
	double* dst;
	FMB_TRY(newVec(n, out, &dst));
	if (n == 0) return FMB_OK;
	const int vec = aligned16(xp) && aligned16(yp) && aligned16(zp) && aligned16(dst);
	const int grid = ewGrid(n);
	switch (opcode) {
		LAUNCH_TERNARY(FMB_T_ADD_PRODUCT) LAUNCH_TERNARY(FMB_T_ADD_PRODUCT_D) LAUNCH_TERNARY(FMB_T_ADD_RATIO) LAUNCH_TERNARY(FMB_T_SUB_RATIO)
		LAUNCH_TERNARY(FMB_T_ACCRUE) LAUNCH_TERNARY(FMB_T_DISCOUNT) LAUNCH_TERNARY(FMB_T_CHOOSE)
	}
	return finishLaunch(out);
}

int fmb_rv_eval_chain(int n_instr, const unsigned char* code, int start_leaf, const fmb_handle* leaves, int n_leaves,
                      const double* scalars, int n_scalars, fmb_handle* out) {
	FMB_TRY(requireInit());
	PinScope pins;
	if (!out || !code || !leaves || n_instr < 1 || n_instr > FMB_CHAIN_MAX_INSTR || n_leaves < 1 || n_leaves > FMB_CHAIN_MAX_LEAVES ||
	    n_scalars < 0 || n_scalars > FMB_CHAIN_MAX_SCALARS || (n_scalars > 0 && !scalars) || start_leaf < 0 || start_leaf >= n_leaves) {
		setError("eval_chain: bad argument");
		return FMB_EINVAL;
	}
	ChainProg p;
	memset(&p, 0, sizeof(p));
	p.n = n_instr; p.start = start_leaf;
	uint64_t n;
	FMB_TRY(commonLength(leaves, n_leaves, &n));
	for (int i = 0; i < n_leaves; i++) {
		if (leaves[i] == 0) { setError("eval_chain: leaf %d is not a device vector", i); return FMB_EINVAL; }
		FMB_TRY(lookupPtr(leaves[i], n, &p.leaf[i]));
	}
	for (int i = 0; i < n_scalars; i++) p.scalar[i] = scalars[i];
	auto refOk = [&](unsigned char r) { return (r & 128) ? (int)(r & 127) < n_scalars : (int)r < n_leaves; };
	for (int k = 0; k < n_instr; k++) {
		ChainInstr c;
		memcpy(&c, code + 8 * k, sizeof(c));
		bool ok = c.kind <= 2;
		if (ok && c.kind == 0) {
			ok = c.op <= FMB_U_ICDF_NORMAL && (c.refA & 128) && refOk(c.refA);
			// Math.pow(x, 0.5) / Math.pow(x, 2.0) must equal sqrt / x*x bit-for-bit, as in fmb_rv_unary
			if (ok && c.op == FMB_U_POW && p.scalar[c.refA & 127] == 0.5) c.op = FMB_U_SQRT;
			else if (ok && c.op == FMB_U_POW && p.scalar[c.refA & 127] == 2.0) c.op = FMB_U_SQUARED;
		} else if (ok && c.kind == 1) {
			ok = c.op <= FMB_B_FLOOR && c.pos <= 1 && refOk(c.refA);
		} else if (ok) {
			ok = c.op <= FMB_T_CHOOSE && c.pos <= 2 && refOk(c.refA) && refOk(c.refB) && (c.refC & 128) && refOk(c.refC);
		}
		if (!ok) { setError("eval_chain: malformed instruction %d", k); return FMB_EINVAL; }
		p.code[k] = c;
	}
	double* dst;
	FMB_TRY(newVec(n, out, &dst));
	if (n == 0) return FMB_OK;
	int vec = aligned16(dst);
	for (int i = 0; i < n_leaves; i++) vec = vec && aligned16(p.leaf[i]);
	chainKernel<<<ewGrid(n), 256, 0, ctx().stream>>>(p, dst, n, n_leaves, vec);
	return finishLaunch(out);
}

int fmb_rv_accrue_chain(int n, const fmb_handle* rates, const double* period_lengths, double divisor, fmb_handle* out) {
	FMB_TRY(requireInit());
	PinScope pins;
	if (n < 1 || !rates || !period_lengths || !out) { setError("accrue_chain: bad argument"); return FMB_EINVAL; }
	uint64_t len = 0;
	std::vector<const double*> ptr(n);
	for (int k = 0; k < n; k++) {
		Vec* v;
		if (rates[k] == 0) { setError("accrue_chain: rate %d is not a device vector", k); return FMB_EINVAL; }
		FMB_TRY(lookup(rates[k], &v));
		if (k > 0 && v->n != len) { setError("operand sizes differ (%llu vs %llu)", (unsigned long long)v->n, (unsigned long long)len); return FMB_EINVAL; }
		len = v->n;
		ptr[k] = v->ptr;
	}
	double* dst;
	FMB_TRY(newVec(len, out, &dst));
	if (len == 0) return FMB_OK;
	const int grid = ewGrid(len);
	const double* accIn = nullptr;
	for (int k0 = 0; k0 < n; k0 += 32) {                       // (more than 32 periods: the partial product is carried in the result vector)
		AccrueChainArgs a;
		const int m = std::min(32, n - k0);
		for (int k = 0; k < 32; k++) { a.rate[k] = k < m ? ptr[k0 + k] : nullptr; a.delta[k] = k < m ? period_lengths[k0 + k] : 0.0; }
		accrueChainKernel<<<grid, 256, 0, ctx().stream>>>(a, m, accIn, k0 + m >= n ? 1 : 0, divisor, dst, len);
		countLaunch();
		accIn = dst;
	}
	FMB_CUDA(cudaGetLastError());
	return FMB_OK;
}

int fmb_rv_accrue_prefix(int n, double start, const fmb_handle* rates, const double* period_lengths, fmb_handle* out) {
	FMB_TRY(requireInit());
	PinScope pins;
	if (n < 1 || !rates || !period_lengths || !out) { setError("accrue_prefix: bad argument"); return FMB_EINVAL; }
	uint64_t len = 0;
	std::vector<const double*> ptr(n);
	for (int k = 0; k < n; k++) {
		Vec* v;
		if (rates[k] == 0) { setError("accrue_prefix: rate %d is not a device vector", k); return FMB_EINVAL; }
		FMB_TRY(lookup(rates[k], &v));
		if (k > 0 && v->n != len) { setError("operand sizes differ (%llu vs %llu)", (unsigned long long)v->n, (unsigned long long)len); return FMB_EINVAL; }
		len = v->n;
		ptr[k] = v->ptr;
	}
	Slab* slab = nullptr;
	FMB_TRY(newSlab(std::max<size_t>(8, (size_t)n * len * sizeof(double)), &slab));
	double* base = (double*)slab->base;
	if (len > 0) {
		const int grid = ewGrid(len);
		const double* accIn = nullptr;
		for (int k0 = 0; k0 < n; k0 += 64) {                       // (more than 64 periods: the running product is carried by the last output)
			AccruePrefixArgs a;
			const int m = std::min(64, n - k0);
			for (int k = 0; k < 64; k++) { a.rate[k] = k < m ? ptr[k0 + k] : nullptr; a.delta[k] = k < m ? period_lengths[k0 + k] : 0.0; }
			accruePrefixKernel<<<grid, 256, 0, ctx().stream>>>(a, m, start, accIn, base + (size_t)k0 * len, len);
			countLaunch();
			accIn = base + (size_t)(k0 + m - 1) * len;
		}
		const cudaError_t e = cudaGetLastError();
		if (e != cudaSuccess) { poolFree(slab->base, slab->bytes); delete slab; setError("accrue_prefix: %s", cudaGetErrorString(e)); return FMB_ECUDA; }
	}
	for (int k = 0; k < n; k++) out[k] = newView(slab, base + (size_t)k * len, len);
	return FMB_OK;
}

} // extern "C"
