// Internal header of libfinmath_b200: context, stream, device-memory pool, handle table.
// Everything here is host-side plumbing behind include/finmath_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <cstring>
#include <atomic>
#include <mutex>
#include <vector>
#include <map>
#include <unordered_map>
#include <string>
#include "../../include/finmath_b200.h"

namespace fmb {

void setError(const char* fmt, ...);

#define FMB_CUDA(call)                                                                                   \
	do {                                                                                                   \
		cudaError_t e__ = (call);                                                                          \
		if (e__ != cudaSuccess) {                                                                          \
			fmb::setError("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e__));   \
			return (e__ == cudaErrorMemoryAllocation) ? FMB_ENOMEM : FMB_ECUDA;                            \
		}                                                                                                  \
	} while (0)

#define FMB_TRY(expr)            \
	do {                         \
		int rc__ = (expr);       \
		if (rc__ != FMB_OK) return rc__; \
	} while (0)

struct Slab {            // one cudaMalloc'ed region shared by several vectors (Brownian increments, process values)
	void* base;
	size_t bytes;
	int refs;
};

struct Vec {             // one RandomVariable's realizations on the device
	double* ptr;
	uint64_t n;
	int refs;
	Slab* slab;          // non-null: ptr is a view into slab->base
	size_t bytes;        // own allocation size (slab == nullptr)
};

struct Context {
	bool initialized = false;
	int device = -1;
	int smCount = 0;
	cudaStream_t stream = nullptr;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	std::atomic<int> fpMode{0};
	std::atomic<uint64_t> launches{0};
	std::mutex mu;
	std::unordered_map<uint64_t, Vec*> table;
	uint64_t nextHandle = 0x1000;
	std::map<size_t, std::vector<void*>> freeLists;
	uint64_t bytesInUse = 0, bytesCached = 0;
	// small pinned + device scratch for reductions and parameter tables
	void* pinned = nullptr; size_t pinnedBytes = 0;
	void* scratch = nullptr; size_t scratchBytes = 0;
	std::mutex scratchMu;       // serialises users of pinned/scratch (reductions, table uploads)
};

Context& ctx();
int requireInit();

// pool
int poolAlloc(size_t bytes, void** out);            // stream-ordered reuse on ctx().stream
void poolFree(void* p, size_t bytes);
size_t roundBytes(size_t bytes);

// handles
int newVec(uint64_t n, fmb_handle* h, double** ptr);                    // own allocation
int newSlab(size_t bytes, Slab** slab);
fmb_handle newView(Slab* slab, double* ptr, uint64_t n);                 // refs = 1, slab->refs++
int lookup(fmb_handle h, Vec** v);                                       // no ref change
int lookupPtr(fmb_handle h, uint64_t expectN, const double** ptr);       // h == 0 -> nullptr; checks length if expectN != 0
int ensureScratch(size_t pinnedBytes, size_t deviceBytes);

inline void countLaunch(int n = 1) { ctx().launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

inline int gridFor(uint64_t n, int block, int perThread = 1) {
	uint64_t g = (n + (uint64_t)block * perThread - 1) / ((uint64_t)block * perThread);
	uint64_t cap = (uint64_t)ctx().smCount * 16;
	if (g > cap) g = cap;
	if (g < 1) g = 1;
	return (int)g;
}

} // namespace fmb
