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

// Multi-GPU exchange of reduction partials (one process per GPU): an NCCL communicator on the library's own compute stream, so that the
// all-gather of a few doubles is stream-ordered between the local reduction kernel and the merge / solve kernel (no host round trip).
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy the host process already loaded, e.g. torch's, else the system's).
struct Comm {
	bool active = false;
	int rank = 0, world = 1;
	void* lib = nullptr;         // dlopen handle
	void* comm = nullptr;        // ncclComm_t
	double* sendBuf = nullptr;   // device, COMM_MAX_DOUBLES
	double* gatherBuf = nullptr; // device, world * COMM_MAX_DOUBLES (peer mode: the half of peerBase the last exchange filled)
	uint64_t exchanges = 0;
	// peer-memory exchange (fmb_comm_peer_open): every rank stores its partials straight into the gather buffers of all ranks over
	// NVLink; one allocation per rank: gather[2][world][COMM_MAX_DOUBLES] doubles, then flags[2][world] (two halves used alternately)
	bool peer = false;
	double* ncclGatherBuf = nullptr;             // the NCCL path's gather buffer while the peer path is active
	double* peerBase = nullptr;                  // this rank's allocation (exported with cudaIpcGetMemHandle)
	double* peerGather[8] = {nullptr};           // base of every rank's allocation as mapped into this process (own included)
	unsigned int* peerErrDev = nullptr;          // set by a kernel that gave up waiting for a peer
	volatile unsigned int* peerErrHost = nullptr;
};
static const int COMM_MAX_DOUBLES = 512;      // 2 * (8*9/2 + 8) = 88 doubles for the largest regression (K = 8); 256 for a radix-select histogram

// ---- peer-memory exchange (fmb_comm.cu sets it up; the reduction kernels call peerExchangeBlock from their finalising CTA) -------------
static const size_t PEER_HALF_DOUBLES = 8 * (size_t)COMM_MAX_DOUBLES;          // gather[world <= 8][COMM_MAX_DOUBLES]
static const size_t PEER_FLAGS_OFFSET = 2 * PEER_HALF_DOUBLES;                 // in 8-byte words: flags[2][8] behind the two halves
static const size_t PEER_ALLOC_BYTES = (PEER_FLAGS_OFFSET + 16) * sizeof(double);
struct PeerArgs {
	int rank = 0, world = 1, half = 0;       // world == 1: no exchange
	unsigned long long seq = 0;              // value of the flags of this exchange (unique, non-zero, the same on all ranks)
	double* base[8] = {nullptr};             // every rank's buffer as mapped into this process
	unsigned int* err = nullptr;
};
#ifdef __CUDACC__
// Called by ALL threads of one CTA (blockDim >= world).  Stores `count` doubles of payload into slot `rank` of every rank's gather
// buffer (NVLink stores), releases this rank's flag at every destination and waits until every rank's flag has arrived here.  On return
// the payloads of all ranks are at peerGathered(px, count) as [world][count].
__device__ __forceinline__ const double* peerGathered(const PeerArgs& px) { return px.base[px.rank] + px.half * PEER_HALF_DOUBLES; }
__device__ inline void peerExchangeBlock(const PeerArgs& px, const double* payload, int count) {
	const int tid = threadIdx.x;
	for (int d = 0; d < px.world; d++) {
		double* dst = px.base[d] + px.half * PEER_HALF_DOUBLES + (size_t)px.rank * count;
		for (int i = tid; i < count; i += blockDim.x) dst[i] = payload[i];
	}
	__threadfence_system();
	__syncthreads();
	if (tid < px.world) {
		unsigned long long* f = reinterpret_cast<unsigned long long*>(px.base[tid] + PEER_FLAGS_OFFSET) + px.half * 8 + px.rank;
		asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(f), "l"(px.seq) : "memory");
		const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(px.base[px.rank] + PEER_FLAGS_OFFSET) + px.half * 8 + tid;
		unsigned long long v;
		const long long t0 = clock64();
		for (;;) {
			asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
			if (v == px.seq) break;
			if (clock64() - t0 > 60000000000ll) { atomicExch(px.err, 1u); break; }     // ~30 s: a peer never arrived (reported by the next reduction)
		}
	}
	__syncthreads();
}
#endif

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
	std::mutex scratchMu;       // serialises users of pinned/scratch (reductions, table uploads) and keeps their launch sequences contiguous on the stream
	// results of reductions: written by the finalising kernel straight into mapped pinned host memory (no separate device-to-host copy)
	double* hostResult = nullptr;      // host address, COMM_MAX_DOUBLES doubles
	double* hostResultDev = nullptr;   // the same memory as the device sees it
	unsigned int* ticket = nullptr;    // device counter of the last-block finalisation (zero between launches)
	unsigned int* ticketMany = nullptr; // 65 counters of fmb_rv_reduce_many (one per vector + vectors done)
	Comm comm;
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
int lookup(fmb_handle h, Vec** v);                                       // pins the vector for the lifetime of the caller's PinScope
int releaseRef(fmb_handle h);                                            // -1 reference; frees at zero (fmb_rv_free)
// Every entry point that dereferences handles opens a PinScope first: lookup() then takes one reference per handle and the scope
// gives them back when the entry point returns (after its kernels are enqueued), so a concurrent fmb_rv_free from a GC / cleaner
// thread (the documented threading contract) can neither delete a Vec that is being read nor hand its block to another allocation
// before the kernel that reads it is in the stream.
struct PinScope {
	std::vector<fmb_handle> held;
	PinScope* prev;
	PinScope();
	~PinScope();
};
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
