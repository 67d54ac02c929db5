// Multi-GPU exchange of reduction partials: one process per GPU, paths sharded by MT19937 jump-ahead.
//
// The only data that ever crosses GPUs is the handful of double-double partial sums behind getAverage / getVariance / getMin /
// getMax (J/montecarlo/RandomVariableFromDoubleArray.java:262-428) and the K(K+1)/2 + K regression moments
// (J/montecarlo/conditionalexpectation/MonteCarloConditionalExpectationRegression.java:118-150).  They are all-gathered with NCCL
// ON THE LIBRARY'S COMPUTE STREAM, between the kernel that produced the local partials and the kernel that merges them in rank
// order - so a regression needs no host round trip at all, and every rank computes identical bits (a sum all-reduce would add the
// hi and lo words separately in an order NCCL chooses; gathering and merging in rank order keeps the double-double exact and
// deterministic).
//
// NCCL is bound at run time with dlopen("libnccl.so.2"): inside a host process that already loaded NCCL (Python with torch) this
// resolves to that very copy, in a JVM to the system library; the library itself has no link-time dependency on it.  The 128-byte
// unique id is created on rank 0 (fmb_comm_unique_id) and distributed by the host through whatever channel it has (torch.distributed
// in the Python binding, any socket in a JVM).
#include "fmb_common.cuh"
#include <dlfcn.h>

namespace fmb {

// the few NCCL entry points used, declared here so that no NCCL header is needed to build (ABI: nccl.h of NCCL 2.x)
struct NcclUniqueId { char internal[128]; };
typedef int (*ncclGetUniqueId_t)(NcclUniqueId*);
typedef int (*ncclCommInitRank_t)(void**, int, NcclUniqueId, int);
typedef int (*ncclCommDestroy_t)(void*);
typedef int (*ncclAllGather_t)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef const char* (*ncclGetErrorString_t)(int);
static const int kNcclFloat64 = 8;

static struct {
	ncclGetUniqueId_t getUniqueId = nullptr;
	ncclCommInitRank_t commInitRank = nullptr;
	ncclCommDestroy_t commDestroy = nullptr;
	ncclAllGather_t allGather = nullptr;
	ncclGetErrorString_t errorString = nullptr;
} g_nccl;

static int bindNccl() {
	Comm& m = ctx().comm;
	if (m.lib) return FMB_OK;
	const char* names[] = { getenv("FMB_NCCL_LIB"), "libnccl.so.2", "libnccl.so" };
	for (const char* name : names) {
		if (!name || !*name) continue;
		m.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
		if (m.lib) break;
	}
	if (!m.lib) { setError("NCCL not found (dlopen libnccl.so.2: %s); set FMB_NCCL_LIB", dlerror()); return FMB_EUNSUPPORTED; }
	g_nccl.getUniqueId = (ncclGetUniqueId_t)dlsym(m.lib, "ncclGetUniqueId");
	g_nccl.commInitRank = (ncclCommInitRank_t)dlsym(m.lib, "ncclCommInitRank");
	g_nccl.commDestroy = (ncclCommDestroy_t)dlsym(m.lib, "ncclCommDestroy");
	g_nccl.allGather = (ncclAllGather_t)dlsym(m.lib, "ncclAllGather");
	g_nccl.errorString = (ncclGetErrorString_t)dlsym(m.lib, "ncclGetErrorString");
	if (!g_nccl.getUniqueId || !g_nccl.commInitRank || !g_nccl.commDestroy || !g_nccl.allGather || !g_nccl.errorString) {
		setError("libnccl lacks an expected symbol");
		dlclose(m.lib); m.lib = nullptr;
		return FMB_EUNSUPPORTED;
	}
	return FMB_OK;
}

// all-gather `count` doubles per rank from comm.sendBuf into comm.gatherBuf ([world][count]) on the compute stream
int commAllGather(int count) {
	Context& c = ctx();
	Comm& m = c.comm;
	if (!m.active) return FMB_OK;
	if (count > COMM_MAX_DOUBLES) { setError("exchange of %d doubles exceeds the buffer (%d)", count, COMM_MAX_DOUBLES); return FMB_EINVAL; }
	const int rc = g_nccl.allGather(m.sendBuf, m.gatherBuf, (size_t)count, kNcclFloat64, m.comm, c.stream);
	if (rc != 0) { setError("ncclAllGather failed: %s", g_nccl.errorString(rc)); return FMB_ECUDA; }
	m.exchanges++;
	return FMB_OK;
}

} // namespace fmb

using namespace fmb;

extern "C" {

int fmb_comm_unique_id(unsigned char* id, int len) {
	FMB_TRY(requireInit());
	if (!id || len < 128) { setError("comm_unique_id: the buffer must hold 128 bytes"); return FMB_EINVAL; }
	FMB_TRY(bindNccl());
	NcclUniqueId u;
	const int rc = g_nccl.getUniqueId(&u);
	if (rc != 0) { setError("ncclGetUniqueId failed: %s", g_nccl.errorString(rc)); return FMB_ECUDA; }
	memcpy(id, u.internal, 128);
	return FMB_OK;
}

int fmb_comm_init(const unsigned char* id, int len, int rank, int world) {
	FMB_TRY(requireInit());
	if (!id || len < 128 || world < 1 || rank < 0 || rank >= world) { setError("comm_init: bad argument"); return FMB_EINVAL; }
	Context& c = ctx();
	Comm& m = c.comm;
	std::lock_guard<std::mutex> lk(c.scratchMu);
	if (m.active) { setError("communicator already initialised (rank %d of %d)", m.rank, m.world); return FMB_EINVAL; }
	if (world == 1) { m.rank = 0; m.world = 1; return FMB_OK; }
	FMB_TRY(bindNccl());
	NcclUniqueId u;
	memcpy(u.internal, id, 128);
	void* comm = nullptr;
	const int rc = g_nccl.commInitRank(&comm, world, u, rank);
	if (rc != 0) { setError("ncclCommInitRank failed: %s", g_nccl.errorString(rc)); return FMB_ECUDA; }
	FMB_CUDA(cudaMalloc((void**)&m.sendBuf, COMM_MAX_DOUBLES * sizeof(double)));
	FMB_CUDA(cudaMalloc((void**)&m.gatherBuf, (size_t)world * COMM_MAX_DOUBLES * sizeof(double)));
	m.comm = comm; m.rank = rank; m.world = world; m.active = true; m.exchanges = 0;
	return FMB_OK;
}

int fmb_comm_shutdown(void) {
	Context& c = ctx();
	Comm& m = c.comm;
	if (!m.active) return FMB_OK;
	cudaStreamSynchronize(c.stream);
	g_nccl.commDestroy(m.comm);
	cudaFree(m.sendBuf); cudaFree(m.gatherBuf);
	m.sendBuf = m.gatherBuf = nullptr; m.comm = nullptr; m.active = false; m.rank = 0; m.world = 1;
	return FMB_OK;
}

int fmb_comm_info(int* rank, int* world, uint64_t* exchanges) {
	const Comm& m = ctx().comm;
	if (rank) *rank = m.rank;
	if (world) *world = m.active ? m.world : 1;
	if (exchanges) *exchanges = m.exchanges;
	return FMB_OK;
}

} // extern "C"
