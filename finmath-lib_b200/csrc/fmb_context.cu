// Context, device-memory pool, handle table, upload/download.  Host-side plumbing of libfinmath_b200.
#include "fmb_common.cuh"

namespace fmb {

static thread_local char g_err[1024] = "";

void setError(const char* fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
}

Context& ctx() {
	static Context c;
	return c;
}

int requireInit() {
	if (!ctx().initialized) {
		// lazy init so that a JNI caller does not need an explicit init: FMB_DEVICE, else LOCAL_RANK (one process per GPU under
		// torchrun), else device 0.  A multi-process job without either must call fmb_init itself: silently putting every rank on
		// device 0 would be wrong.
		const char* d = getenv("FMB_DEVICE");
		if (!d) d = getenv("LOCAL_RANK");
		if (!d) {
			const char* w = getenv("WORLD_SIZE");
			if (w && atoi(w) > 1) {
				setError("WORLD_SIZE=%s but neither FMB_DEVICE nor LOCAL_RANK is set: call fmb_init(device) explicitly", w);
				return FMB_EINVAL;
			}
		}
		return fmb_init(d ? atoi(d) : 0);
	}
	// the current device is per host thread; callers may be pool / GC threads
	if (cudaSetDevice(ctx().device) != cudaSuccess) { cudaGetLastError(); setError("cudaSetDevice(%d) failed", ctx().device); return FMB_ECUDA; }
	return FMB_OK;
}

size_t roundBytes(size_t bytes) {
	if (bytes == 0) bytes = 8;
	return (bytes + 511) & ~(size_t)511;
}

int poolAlloc(size_t bytes, void** out) {
	Context& c = ctx();
	bytes = roundBytes(bytes);
	{
		std::lock_guard<std::mutex> lk(c.mu);
		auto it = c.freeLists.find(bytes);
		if (it != c.freeLists.end() && !it->second.empty()) {
			*out = it->second.back();
			it->second.pop_back();
			c.bytesCached -= bytes;
			c.bytesInUse += bytes;
			return FMB_OK;
		}
	}
	cudaError_t e = cudaMalloc(out, bytes);
	if (e != cudaSuccess) {
		cudaGetLastError();
		fmb_pool_trim();                       // give cached blocks back and retry once
		e = cudaMalloc(out, bytes);
		if (e != cudaSuccess) {
			cudaGetLastError();
			setError("device allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
			return FMB_ENOMEM;
		}
	}
	std::lock_guard<std::mutex> lk(c.mu);
	c.bytesInUse += bytes;
	return FMB_OK;
}

void poolFree(void* p, size_t bytes) {
	Context& c = ctx();
	bytes = roundBytes(bytes);
	std::lock_guard<std::mutex> lk(c.mu);
	c.freeLists[bytes].push_back(p);
	c.bytesInUse -= bytes;
	c.bytesCached += bytes;
}

int newVec(uint64_t n, fmb_handle* h, double** ptr) {
	void* p = nullptr;
	FMB_TRY(poolAlloc(n * sizeof(double), &p));
	Vec* v = new Vec{(double*)p, n, 1, nullptr, n * sizeof(double)};
	Context& c = ctx();
	std::lock_guard<std::mutex> lk(c.mu);
	*h = c.nextHandle++;
	c.table[*h] = v;
	if (ptr) *ptr = v->ptr;
	return FMB_OK;
}

int newSlab(size_t bytes, Slab** slab) {
	void* p = nullptr;
	FMB_TRY(poolAlloc(bytes, &p));
	*slab = new Slab{p, bytes, 0};
	return FMB_OK;
}

fmb_handle newView(Slab* slab, double* ptr, uint64_t n) {
	Vec* v = new Vec{ptr, n, 1, slab, 0};
	Context& c = ctx();
	std::lock_guard<std::mutex> lk(c.mu);
	slab->refs++;
	fmb_handle h = c.nextHandle++;
	c.table[h] = v;
	return h;
}

static thread_local PinScope* g_scope = nullptr;

PinScope::PinScope() : prev(g_scope) { g_scope = this; }
PinScope::~PinScope() {
	g_scope = prev;
	for (fmb_handle h : held) releaseRef(h);
}

int lookup(fmb_handle h, Vec** v) {
	Context& c = ctx();
	std::lock_guard<std::mutex> lk(c.mu);
	auto it = c.table.find(h);
	if (it == c.table.end()) {
		setError("unknown or freed handle 0x%llx", (unsigned long long)h);
		return FMB_EHANDLE;
	}
	*v = it->second;
	if (g_scope) { it->second->refs++; g_scope->held.push_back(h); }
	return FMB_OK;
}

int releaseRef(fmb_handle h) {
	Context& c = ctx();
	Vec* v = nullptr;
	Slab* dead = nullptr;
	{
		std::lock_guard<std::mutex> lk(c.mu);
		auto it = c.table.find(h);
		if (it == c.table.end()) { setError("unknown or freed handle 0x%llx", (unsigned long long)h); return FMB_EHANDLE; }
		v = it->second;
		if (--v->refs > 0) return FMB_OK;
		c.table.erase(it);
		if (v->slab && --v->slab->refs == 0) dead = v->slab;
	}
	if (!v->slab) poolFree(v->ptr, v->bytes);
	if (dead) { poolFree(dead->base, dead->bytes); delete dead; }
	delete v;
	return FMB_OK;
}

int lookupPtr(fmb_handle h, uint64_t expectN, const double** ptr) {
	if (h == 0) { *ptr = nullptr; return FMB_OK; }
	Vec* v;
	FMB_TRY(lookup(h, &v));
	if (expectN != 0 && v->n != expectN) {
		setError("size mismatch: handle 0x%llx has %llu elements, expected %llu", (unsigned long long)h,
		         (unsigned long long)v->n, (unsigned long long)expectN);
		return FMB_EINVAL;
	}
	*ptr = v->ptr;
	return FMB_OK;
}

int ensureScratch(size_t pinnedBytes, size_t deviceBytes) {
	Context& c = ctx();
	if (c.pinnedBytes < pinnedBytes) {
		if (c.pinned) cudaFreeHost(c.pinned);
		c.pinned = nullptr; c.pinnedBytes = 0;
		FMB_CUDA(cudaMallocHost(&c.pinned, pinnedBytes));
		c.pinnedBytes = pinnedBytes;
	}
	if (c.scratchBytes < deviceBytes) {
		if (c.scratch) { cudaStreamSynchronize(c.stream); cudaFree(c.scratch); }
		c.scratch = nullptr; c.scratchBytes = 0;
		FMB_CUDA(cudaMalloc(&c.scratch, deviceBytes));
		c.scratchBytes = deviceBytes;
	}
	return FMB_OK;
}

} // namespace fmb

using namespace fmb;

extern "C" {

const char* fmb_last_error(void) { return g_err; }

int fmb_device_count(int* count) {
	if (!count) return FMB_EINVAL;
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess) { cudaGetLastError(); n = 0; }
	*count = n;
	return FMB_OK;
}

int fmb_init(int device) {
	Context& c = ctx();
	static std::mutex initMu;
	std::lock_guard<std::mutex> lk(initMu);
	if (c.initialized) {
		if (c.device != device) { setError("already initialised on device %d", c.device); return FMB_EINVAL; }
		return FMB_OK;
	}
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n == 0) {
		cudaGetLastError();
		setError("no CUDA device available (%s); this library has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
		return FMB_ENODEVICE;
	}
	if (device < 0 || device >= n) { setError("device %d out of range (0..%d)", device, n - 1); return FMB_EINVAL; }
	cudaDeviceProp prop;
	FMB_CUDA(cudaGetDeviceProperties(&prop, device));
	if (prop.major != 10) {
		setError("device %d is sm_%d%d; this library contains sm_100a code only", device, prop.major, prop.minor);
		return FMB_ENODEVICE;
	}
	FMB_CUDA(cudaSetDevice(device));
	FMB_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
	FMB_CUDA(cudaEventCreate(&c.ev0));
	FMB_CUDA(cudaEventCreate(&c.ev1));
	c.device = device;
	c.smCount = prop.multiProcessorCount;
	FMB_CUDA(cudaHostAlloc((void**)&c.hostResult, COMM_MAX_DOUBLES * sizeof(double), cudaHostAllocMapped));
	FMB_CUDA(cudaHostGetDevicePointer((void**)&c.hostResultDev, c.hostResult, 0));
	FMB_CUDA(cudaMalloc((void**)&c.ticket, 64 + 66 * sizeof(unsigned int)));
	FMB_CUDA(cudaMemsetAsync(c.ticket, 0, 64 + 66 * sizeof(unsigned int), c.stream));
	c.ticketMany = c.ticket + 16;
	c.initialized = true;
	return ensureScratch(1 << 16, 1 << 20);
}

int fmb_is_initialized(void) { return ctx().initialized ? 1 : 0; }

int fmb_shutdown(void) {
	Context& c = ctx();
	if (!c.initialized) return FMB_OK;
	cudaSetDevice(c.device);
	cudaStreamSynchronize(c.stream);
	{
		std::lock_guard<std::mutex> lk(c.mu);
		for (auto& kv : c.table) {
			Vec* v = kv.second;
			if (!v->slab) cudaFree(v->ptr);
			delete v;
		}
		c.table.clear();
		for (auto& kv : c.freeLists) for (void* p : kv.second) cudaFree(p);
		c.freeLists.clear();
		c.bytesCached = c.bytesInUse = 0;
	}
	fmb_comm_shutdown();
	if (c.pinned) cudaFreeHost(c.pinned);
	if (c.scratch) cudaFree(c.scratch);
	if (c.hostResult) cudaFreeHost(c.hostResult);
	if (c.ticket) cudaFree(c.ticket);
	c.hostResult = c.hostResultDev = nullptr; c.ticket = nullptr; c.ticketMany = nullptr;
	c.pinned = c.scratch = nullptr; c.pinnedBytes = c.scratchBytes = 0;
	cudaEventDestroy(c.ev0); cudaEventDestroy(c.ev1);
	cudaStreamDestroy(c.stream);
	c.initialized = false;
	return FMB_OK;
}

int fmb_device_name(char* buf, int len) {
	FMB_TRY(requireInit());
	cudaDeviceProp prop;
	FMB_CUDA(cudaGetDeviceProperties(&prop, ctx().device));
	snprintf(buf, len, "%s", prop.name);
	return FMB_OK;
}

int fmb_synchronize(void) {
	FMB_TRY(requireInit());
	FMB_CUDA(cudaStreamSynchronize(ctx().stream));
	return FMB_OK;
}

int fmb_set_fp_mode(int mode) {
	if (mode != 0 && mode != 1) { setError("fp mode must be 0 (strict) or 1 (fast)"); return FMB_EINVAL; }
	ctx().fpMode.store(mode);
	return FMB_OK;
}
int fmb_get_fp_mode(int* mode) { if (!mode) return FMB_EINVAL; *mode = ctx().fpMode.load(); return FMB_OK; }

int fmb_timer_start(void) {
	FMB_TRY(requireInit());
	FMB_CUDA(cudaEventRecord(ctx().ev0, ctx().stream));
	return FMB_OK;
}
int fmb_timer_stop_ms(float* ms) {
	FMB_TRY(requireInit());
	FMB_CUDA(cudaEventRecord(ctx().ev1, ctx().stream));
	FMB_CUDA(cudaEventSynchronize(ctx().ev1));
	FMB_CUDA(cudaEventElapsedTime(ms, ctx().ev0, ctx().ev1));
	return FMB_OK;
}
int fmb_kernel_launch_count(uint64_t* count) { if (!count) return FMB_EINVAL; *count = ctx().launches.load(); return FMB_OK; }

int fmb_rv_create(uint64_t n, fmb_handle* out) {
	FMB_TRY(requireInit());
	if (!out) return FMB_EINVAL;
	return newVec(n, out, nullptr);
}

int fmb_rv_upload(const double* host, uint64_t n, fmb_handle* out) {
	FMB_TRY(requireInit());
	if (!out || (!host && n)) { setError("null argument"); return FMB_EINVAL; }
	double* p;
	FMB_TRY(newVec(n, out, &p));
	if (n) {
		FMB_CUDA(cudaMemcpyAsync(p, host, n * sizeof(double), cudaMemcpyHostToDevice, ctx().stream));
		FMB_CUDA(cudaStreamSynchronize(ctx().stream));   // the caller may reuse its buffer immediately
	}
	return FMB_OK;
}

int fmb_rv_download(fmb_handle h, double* host, uint64_t n) {
	FMB_TRY(requireInit());
	PinScope pins;
	Vec* v;
	FMB_TRY(lookup(h, &v));
	if (n != v->n) { setError("download of %llu elements from a vector of %llu", (unsigned long long)n, (unsigned long long)v->n); return FMB_EINVAL; }
	if (n) {
		FMB_CUDA(cudaMemcpyAsync(host, v->ptr, n * sizeof(double), cudaMemcpyDeviceToHost, ctx().stream));
		FMB_CUDA(cudaStreamSynchronize(ctx().stream));
	}
	return FMB_OK;
}

int fmb_rv_get(fmb_handle h, uint64_t i, double* out) {
	FMB_TRY(requireInit());
	PinScope pins;
	Vec* v;
	FMB_TRY(lookup(h, &v));
	if (i >= v->n) { setError("index %llu out of bounds (size %llu)", (unsigned long long)i, (unsigned long long)v->n); return FMB_EINVAL; }
	FMB_CUDA(cudaMemcpyAsync(out, v->ptr + i, sizeof(double), cudaMemcpyDeviceToHost, ctx().stream));
	FMB_CUDA(cudaStreamSynchronize(ctx().stream));
	return FMB_OK;
}

int fmb_rv_size(fmb_handle h, uint64_t* n) {
	PinScope pins;
	Vec* v;
	FMB_TRY(lookup(h, &v));
	*n = v->n;
	return FMB_OK;
}

int fmb_rv_device_ptr(fmb_handle h, void** dptr) {
	PinScope pins;
	Vec* v;
	FMB_TRY(lookup(h, &v));
	*dptr = v->ptr;
	return FMB_OK;
}

int fmb_rv_retain(fmb_handle h) {
	Context& c = ctx();
	std::lock_guard<std::mutex> lk(c.mu);
	auto it = c.table.find(h);
	if (it == c.table.end()) { setError("unknown or freed handle 0x%llx", (unsigned long long)h); return FMB_EHANDLE; }
	it->second->refs++;
	return FMB_OK;
}

int fmb_rv_free(fmb_handle h) {
	if (h == 0) return FMB_OK;
	return releaseRef(h);
}

int fmb_rv_free_many(const fmb_handle* handles, uint64_t count) {
	if (count && !handles) return FMB_EINVAL;
	for (uint64_t i = 0; i < count; i++) if (handles[i]) releaseRef(handles[i]);      // unknown handles are skipped, like free() from a cleaner
	return FMB_OK;
}

int fmb_pool_stats(uint64_t* bytes_in_use, uint64_t* bytes_cached, uint64_t* live_handles) {
	Context& c = ctx();
	std::lock_guard<std::mutex> lk(c.mu);
	if (bytes_in_use) *bytes_in_use = c.bytesInUse;
	if (bytes_cached) *bytes_cached = c.bytesCached;
	if (live_handles) *live_handles = c.table.size();
	return FMB_OK;
}

int fmb_pool_trim(void) {
	Context& c = ctx();
	if (!c.initialized) return FMB_OK;
	std::vector<void*> toFree;
	{
		std::lock_guard<std::mutex> lk(c.mu);
		for (auto& kv : c.freeLists) { for (void* p : kv.second) toFree.push_back(p); kv.second.clear(); }
		c.bytesCached = 0;
	}
	if (!toFree.empty()) {
		cudaStreamSynchronize(c.stream);
		for (void* p : toFree) cudaFree(p);
	}
	return FMB_OK;
}

} // extern "C"
