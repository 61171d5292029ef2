/*
 * ctbd_runtime.cu -- device selection, stream, stream-ordered memory pool, copies, events.
 * Implements the "runtime" and "memory" groups of include/ctb_device.h.
 */
#include <stdlib.h>
#include <unordered_map>
#include <mutex>
#include "ctbd_common.cuh"

namespace ctbd {

Runtime& rt() { static Runtime r; return r; }

static std::mutex g_mutex;
static std::unordered_map<void*, size_t> g_allocs;

int fail(const char* what, cudaError_t e, const char* file, int line)
{
	snprintf(rt().err, sizeof(rt().err), "%s: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
	return -1;
}

int fail_msg(const char* msg)
{
	snprintf(rt().err, sizeof(rt().err), "%s", msg);
	return -1;
}

int upload(const void* host, size_t bytes, void** dev)
{
	*dev = nullptr;
	int rc = ctbd_malloc(dev, bytes);
	if (rc < 0) { return rc; }
	if (bytes > 0) { CTBD_CUDA(cudaMemcpyAsync(*dev, host, bytes, cudaMemcpyHostToDevice, rt().stream)); }
	return 0;
}

} // namespace ctbd

using namespace ctbd;

extern "C" {

int ctbd_init(int device)
{
	std::lock_guard<std::mutex> lock(g_mutex);
	Runtime& r = rt();
	if (r.ready && (device < 0 || device == r.device)) { return 0; }
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if (e != cudaSuccess || ndev == 0) {
		snprintf(r.err, sizeof(r.err), "no CUDA device available (%s); the engine has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
		return -1;
	}
	if (device < 0)
	{
		const char* env = getenv("CTB_DEVICE");
		if (env != nullptr) { device = atoi(env); }
		else {
			/* one process per GPU under torchrun */
			const char* lr = getenv("LOCAL_RANK");
			device = (lr != nullptr) ? atoi(lr) % ndev : 0;
		}
	}
	if (device >= ndev) { snprintf(r.err, sizeof(r.err), "device %d out of range (%d devices)", device, ndev); return -1; }
	if (r.ready && r.stream != nullptr) { cudaStreamSynchronize(r.stream); cudaStreamDestroy(r.stream); r.stream = nullptr; r.ready = false; }
	CTBD_CUDA(cudaSetDevice(device));
	cudaDeviceProp prop;
	CTBD_CUDA(cudaGetDeviceProperties(&prop, device));
	if (prop.major < 10) {
		snprintf(r.err, sizeof(r.err), "device %d is sm_%d%d; this engine is built for sm_100a only", device, prop.major, prop.minor);
		return -1;
	}
	r.device = device;
	r.sm_count = prop.multiProcessorCount;
	r.smem_optin = (int)prop.sharedMemPerBlockOptin;
	CTBD_CUDA(cudaStreamCreateWithFlags(&r.stream, cudaStreamNonBlocking));
	/* keep freed blocks in the pool: the sweep re-allocates the same sizes over and over */
	cudaMemPool_t pool;
	CTBD_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
	uint64_t threshold = UINT64_MAX;
	CTBD_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold));
	r.ready = true;
	return 0;
}

int ctbd_shutdown(void)
{
	std::lock_guard<std::mutex> lock(g_mutex);
	Runtime& r = rt();
	if (!r.ready) { return 0; }
	cudaStreamSynchronize(r.stream);
	cudaStreamDestroy(r.stream);
	r.stream = nullptr;
	r.ready = false;
	return 0;
}

int ctbd_backend(void) { return 1; }
const char* ctbd_last_error(void) { return rt().err; }
long long ctbd_launch_count(void) { return rt().launches; }
int ctbd_sm_count(void) { return rt().sm_count; }
void* ctbd_stream(void) { return (void*)rt().stream; }

int ctbd_event_create(void** ev)
{
	CTBD_REQUIRE_INIT();
	cudaEvent_t e;
	CTBD_CUDA(cudaEventCreate(&e));
	*ev = (void*)e;
	return 0;
}
int ctbd_event_record(void* ev) { CTBD_CUDA(cudaEventRecord((cudaEvent_t)ev, rt().stream)); return 0; }
int ctbd_event_elapsed_ms(void* ev_start, void* ev_stop, float* ms)
{
	CTBD_CUDA(cudaEventSynchronize((cudaEvent_t)ev_stop));
	CTBD_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)ev_start, (cudaEvent_t)ev_stop));
	return 0;
}
int ctbd_event_destroy(void* ev) { CTBD_CUDA(cudaEventDestroy((cudaEvent_t)ev)); return 0; }

int ctbd_malloc(void** dptr, size_t bytes)
{
	CTBD_REQUIRE_INIT();
	const size_t nb = bytes > 0 ? bytes : 16;
	CTBD_CUDA(cudaMallocAsync(dptr, nb, rt().stream));
	CTBD_CUDA(cudaMemsetAsync(*dptr, 0, nb, rt().stream));
	{
		std::lock_guard<std::mutex> lock(g_mutex);
		g_allocs[*dptr] = nb;
		rt().bytes_in_use += (long long)nb;
	}
	return 0;
}

int ctbd_free(void* dptr)
{
	if (dptr == nullptr) { return 0; }
	{
		std::lock_guard<std::mutex> lock(g_mutex);
		auto it = g_allocs.find(dptr);
		if (it != g_allocs.end()) { rt().bytes_in_use -= (long long)it->second; g_allocs.erase(it); }
	}
	CTBD_CUDA(cudaFreeAsync(dptr, rt().stream));
	return 0;
}

int ctbd_memset_zero(void* dptr, size_t bytes) { CTBD_CUDA(cudaMemsetAsync(dptr, 0, bytes, rt().stream)); return 0; }

int ctbd_h2d(void* dptr, const void* hptr, size_t bytes)
{
	CTBD_REQUIRE_INIT();
	if (bytes == 0) { return 0; }
	CTBD_CUDA(cudaMemcpyAsync(dptr, hptr, bytes, cudaMemcpyHostToDevice, rt().stream));
	/* the caller may reuse or release 'hptr' right away (pinned staging buffers included) */
	CTBD_CUDA(cudaStreamSynchronize(rt().stream));
	return 0;
}

int ctbd_d2h(void* hptr, const void* dptr, size_t bytes)
{
	CTBD_REQUIRE_INIT();
	if (bytes == 0) { return 0; }
	CTBD_CUDA(cudaMemcpyAsync(hptr, dptr, bytes, cudaMemcpyDeviceToHost, rt().stream));
	CTBD_CUDA(cudaStreamSynchronize(rt().stream));
	return 0;
}

int ctbd_d2d(void* dst, const void* src, size_t bytes)
{
	if (bytes == 0) { return 0; }
	CTBD_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, rt().stream));
	return 0;
}

int ctbd_sync(void)
{
	if (!rt().ready) { return 0; }
	CTBD_CUDA(cudaStreamSynchronize(rt().stream));
	return 0;
}

int ctbd_host_alloc(void** hptr, size_t bytes)
{
	CTBD_REQUIRE_INIT();
	CTBD_CUDA(cudaMallocHost(hptr, bytes > 0 ? bytes : 16));
	return 0;
}
int ctbd_host_free(void* hptr) { CTBD_CUDA(cudaFreeHost(hptr)); return 0; }

long long ctbd_bytes_in_use(void) { return rt().bytes_in_use; }

} // extern "C"
