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

// ---- peer-memory exchange ----------------------------------------------------------------------------------------------------
// The payloads are a few hundred bytes: what an exchange costs is latency, and an NCCL all-gather of 16 bytes per rank takes ~13 us on
// 8 GPUs (launch + protocol).  With the gather buffers of all ranks mapped into every process (CUDA IPC: one process per GPU), ONE
// small kernel per rank stores its partials into everybody's buffer over NVLink, releases a flag per destination and waits for the flags of
// the others.  The reductions and the regression go one step further and do this from the finalising CTA of the very kernel that
// produced the partials, which then also merges the shards (and solves): no second and third launch per exchange (fmb_reduce.cu); the
// stand-alone kernel below serves the remaining callers (order statistics, empty shards).  Two buffer halves are used alternately: a rank can only be one exchange ahead of the slowest one (it needs
// everybody's flag to finish an exchange), so the half it overwrites has been consumed everywhere.
__global__ void __launch_bounds__(128) peerGatherKernel(const double* __restrict__ send, int count, PeerArgs px) { peerExchangeBlock(px, send, count); }

// arguments of the next peer exchange (advances the exchange counter; comm.gatherBuf = where the gathered partials will be)
void peerArgsNext(PeerArgs& px) {
	Comm& m = ctx().comm;
	px.rank = m.rank; px.world = m.world; px.half = (int)(m.exchanges & 1); px.seq = m.exchanges + 1; px.err = m.peerErrDev;
	for (int r = 0; r < 8; r++) px.base[r] = m.peerGather[r];
	m.gatherBuf = m.peerBase + px.half * PEER_HALF_DOUBLES;
	m.exchanges++;
}

// all-gather `count` doubles per rank from comm.sendBuf into comm.gatherBuf ([world][count]) on the compute stream
int commAllGather(int count) {
	Context& c = ctx();
	Comm& m = c.comm;
	if (!m.active) return FMB_OK;
	if (count > COMM_MAX_DOUBLES) { setError("exchange of %d doubles exceeds the buffer (%d)", count, COMM_MAX_DOUBLES); return FMB_EINVAL; }
	if (m.peer) {
		if (*m.peerErrHost) { setError("peer exchange: a rank did not arrive (timed out)"); return FMB_ECUDA; }
		PeerArgs px;
		peerArgsNext(px);
		peerGatherKernel<<<1, 128, 0, c.stream>>>(m.sendBuf, count, px);
		FMB_CUDA(cudaGetLastError());
		countLaunch();
		return FMB_OK;
	}
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
	if (m.peer) {
		for (int r = 0; r < m.world; r++) if (r != m.rank && m.peerGather[r]) cudaIpcCloseMemHandle(m.peerGather[r]);
		for (int r = 0; r < 8; r++) m.peerGather[r] = nullptr;
		m.gatherBuf = m.ncclGatherBuf;
		m.peer = false;
	}
	// (the exported buffer itself is NOT freed here: another rank may still hold its mapping - freeing before every importer has closed is
	// undefined; 64 KB, reused by a later fmb_comm_peer_handle, released with the process)
	if (m.peerErrHost) { cudaFreeHost((void*)m.peerErrHost); m.peerErrHost = nullptr; m.peerErrDev = nullptr; }
	cudaFree(m.sendBuf); cudaFree(m.gatherBuf);
	m.sendBuf = m.gatherBuf = m.ncclGatherBuf = nullptr; m.comm = nullptr; m.active = false; m.rank = 0; m.world = 1;
	return FMB_OK;
}

// Peer-memory exchange, step 1: this rank's buffer (allocated on the first call) as a 64-byte CUDA IPC handle.  The host gathers the
// handles of all ranks (any channel) and hands them to fmb_comm_peer_open.
int fmb_comm_peer_handle(unsigned char* handle, int len) {
	FMB_TRY(requireInit());
	if (!handle || len < (int)sizeof(cudaIpcMemHandle_t)) { setError("comm_peer_handle: the buffer must hold %d bytes", (int)sizeof(cudaIpcMemHandle_t)); return FMB_EINVAL; }
	Context& c = ctx();
	Comm& m = c.comm;
	std::lock_guard<std::mutex> lk(c.scratchMu);
	if (!m.active || m.world < 2 || m.world > 8) { setError("comm_peer_handle: needs a communicator of 2..8 ranks (fmb_comm_init first)"); return FMB_EUNSUPPORTED; }
	if (m.peer) { setError("comm_peer_handle: the peer exchange is already open"); return FMB_EINVAL; }
	if (m.peerBase) {
		FMB_CUDA(cudaMemset(m.peerBase, 0, PEER_ALLOC_BYTES));      // a buffer kept from an earlier communicator: its flags are stale
	} else {
		FMB_CUDA(cudaMalloc((void**)&m.peerBase, PEER_ALLOC_BYTES));
		FMB_CUDA(cudaMemset(m.peerBase, 0, PEER_ALLOC_BYTES));
	}
	if (!m.peerErrHost) {
		unsigned int* eh = nullptr;
		FMB_CUDA(cudaHostAlloc((void**)&eh, sizeof(unsigned int), cudaHostAllocMapped));
		*eh = 0;
		m.peerErrHost = eh;
		FMB_CUDA(cudaHostGetDevicePointer((void**)&m.peerErrDev, eh, 0));
	}
	cudaIpcMemHandle_t h;
	FMB_CUDA(cudaIpcGetMemHandle(&h, m.peerBase));
	memcpy(handle, &h, sizeof(h));
	return FMB_OK;
}

// step 2: handles = world x 64 bytes in rank order (this rank's own entry is ignored).  From then on the exchanges of the reductions and
// regressions go through peer memory instead of NCCL.  Every rank must have called fmb_comm_peer_handle before any rank calls this, and
// all ranks must switch at the same point of their (identical) call sequences.
int fmb_comm_peer_open(const unsigned char* handles, int len) {
	FMB_TRY(requireInit());
	Context& c = ctx();
	Comm& m = c.comm;
	std::lock_guard<std::mutex> lk(c.scratchMu);
	if (!m.active || !m.peerBase) { setError("comm_peer_open: call fmb_comm_peer_handle first"); return FMB_EINVAL; }
	if (!handles) {                                        // NULL: back to the NCCL path (a rank could not map its peers: nobody may use them)
		if (m.peer) {
			FMB_CUDA(cudaStreamSynchronize(c.stream));
			for (int r = 0; r < m.world; r++) if (r != m.rank && m.peerGather[r]) cudaIpcCloseMemHandle(m.peerGather[r]);
			for (int r = 0; r < 8; r++) m.peerGather[r] = nullptr;
			m.gatherBuf = m.ncclGatherBuf;
			m.peer = false;
		}
		return FMB_OK;
	}
	if (len < m.world * (int)sizeof(cudaIpcMemHandle_t)) { setError("comm_peer_open: %d handles of %d bytes expected", m.world, (int)sizeof(cudaIpcMemHandle_t)); return FMB_EINVAL; }
	if (m.peer) return FMB_OK;
	FMB_CUDA(cudaStreamSynchronize(c.stream));
	for (int r = 0; r < m.world; r++) {
		if (r == m.rank) { m.peerGather[r] = m.peerBase; continue; }
		cudaIpcMemHandle_t h;
		memcpy(&h, handles + (size_t)r * sizeof(h), sizeof(h));
		void* p = nullptr;
		const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
		if (e != cudaSuccess) {
			for (int q = 0; q < r; q++) if (q != m.rank && m.peerGather[q]) { cudaIpcCloseMemHandle(m.peerGather[q]); m.peerGather[q] = nullptr; }
			setError("comm_peer_open: cannot map the buffer of rank %d (%s)", r, cudaGetErrorString(e));
			cudaGetLastError();
			return FMB_EUNSUPPORTED;
		}
		m.peerGather[r] = (double*)p;
	}
	m.ncclGatherBuf = m.gatherBuf;
	m.exchanges &= ~1ull;                                  // (all ranks share the call sequence: the same count everywhere; start on half 0)
	m.peer = true;
	return FMB_OK;
}

int fmb_comm_info(int* rank, int* world, uint64_t* exchanges) {
	const Comm& m = ctx().comm;
	if (m.peer && m.peerErrHost && *m.peerErrHost) { setError("peer exchange: a rank did not arrive (timed out)"); return FMB_ECUDA; }
	if (rank) *rank = m.rank;
	if (world) *world = m.active ? m.world : 1;
	if (exchanges) *exchanges = m.exchanges;
	return FMB_OK;
}

} // extern "C"
