/*
 * ctbd_runtime.cu -- device selection, stream, stream-ordered memory pool, copies, events.
 * Implements the "runtime" and "memory" groups of include/ctb_device.h.
 */
#include <stdlib.h>
#include <dlfcn.h>
#include <sys/mman.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <unistd.h>
#include <stddef.h>
#include <omp.h>
#include <unordered_map>
#include <mutex>
#include <vector>
#include <algorithm>
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
	int rc = ctbd_malloc_noinit(dev, bytes);      /* fully overwritten by the copy */
	if (rc < 0) { return rc; }
	if (bytes > 0) { CTBD_CUDA(cudaMemcpyAsync(*dev, host, bytes, cudaMemcpyHostToDevice, rt().stream)); }
	return 0;
}

} // namespace ctbd

using namespace ctbd;

/* ---- block-wise host <-> device transfers through a persistent pinned staging ring ------------------------------
 * Host tensors are thousands of separately allocated (pageable) blocks; the device layout is one packed buffer.  The blocks
 * are streamed through two pinned chunks: while chunk c is in flight on the copy engine the host packs (or unpacks) chunk
 * c+1, so a transfer costs about max(host memcpy, PCIe) instead of pinned allocation + pack + copy in sequence. */
struct ncclUniqueIdBytes { char internal[CTBD_UNIQUE_ID_BYTES]; };   /* layout of ncclUniqueId (nccl.h), passed by value */

namespace {
constexpr size_t STAGE_CHUNK = (size_t)8 << 20;      /* small chunks: the un-overlapped head (first pack) and tail (last DMA) stay short */
constexpr int STAGE_SLOTS = 4;
struct StageRing
{
	void* buf[STAGE_SLOTS] = { nullptr };
	cudaEvent_t ev[STAGE_SLOTS] = { nullptr };
	bool busy[STAGE_SLOTS] = { false };
};
StageRing g_ring;

int ring_init()
{
	if (g_ring.buf[0] != nullptr) { return 0; }
	for (int i = 0; i < STAGE_SLOTS; i++) {
		CTBD_CUDA(cudaMallocHost(&g_ring.buf[i], STAGE_CHUNK));
		CTBD_CUDA(cudaEventCreateWithFlags(&g_ring.ev[i], cudaEventDisableTiming));
	}
	return 0;
}
int ring_wait(int i)
{
	if (g_ring.busy[i]) { CTBD_CUDA(cudaEventSynchronize(g_ring.ev[i])); g_ring.busy[i] = false; }
	return 0;
}
} // namespace

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
	for (int i = 0; i < STAGE_SLOTS; i++) {
		if (g_ring.buf[i] != nullptr) { cudaFreeHost(g_ring.buf[i]); g_ring.buf[i] = nullptr; }
		if (g_ring.ev[i] != nullptr) { cudaEventDestroy(g_ring.ev[i]); g_ring.ev[i] = nullptr; }
		g_ring.busy[i] = false;
	}
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

int ctbd_malloc_noinit(void** dptr, size_t bytes)
{
	CTBD_REQUIRE_INIT();
	const size_t nb = bytes > 0 ? bytes : 16;
	CTBD_CUDA(cudaMallocAsync(dptr, nb, rt().stream));
	{
		std::lock_guard<std::mutex> lock(g_mutex);
		g_allocs[*dptr] = nb;
		rt().bytes_in_use += (long long)nb;
	}
	return 0;
}

int ctbd_malloc(void** dptr, size_t bytes)
{
	int rc = ctbd_malloc_noinit(dptr, bytes);
	if (rc < 0) { return rc; }
	CTBD_CUDA(cudaMemsetAsync(*dptr, 0, bytes > 0 ? bytes : 16, rt().stream));
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


/* pieces of host blocks laid out in consecutive staging chunks (each chunk = one contiguous device range) */
namespace {
struct Piece { int b; int64_t boff, n; size_t pos; };          /* block, byte offset inside the block, bytes, position in the chunk */
struct Chunk { int64_t dev, len; size_t p0, p1; };
constexpr int64_t PIECE_MAX = (int64_t)256 << 10;               /* split large blocks so that the host copy parallelises */

void plan_chunks(int nblk, const int64_t* dev_off, const int64_t* nbytes, std::vector<Piece>& pieces, std::vector<Chunk>& chunks)
{
	Chunk c; c.dev = -1; c.len = 0; c.p0 = 0; c.p1 = 0;
	for (int b = 0; b < nblk; b++) {
		int64_t done = 0;
		while (done < nbytes[b]) {
			if (c.len > 0 && (c.dev + c.len != dev_off[b] + done || (size_t)c.len == STAGE_CHUNK)) { c.p1 = pieces.size(); chunks.push_back(c); c.len = 0; }
			if (c.len == 0) { c.dev = dev_off[b] + done; c.p0 = pieces.size(); }
			const int64_t n = std::min<int64_t>(std::min<int64_t>(nbytes[b] - done, (int64_t)STAGE_CHUNK - c.len), PIECE_MAX);
			Piece pc; pc.b = b; pc.boff = done; pc.n = n; pc.pos = (size_t)c.len; pieces.push_back(pc);
			c.len += n; done += n;
		}
	}
	if (c.len > 0) { c.p1 = pieces.size(); chunks.push_back(c); }
}

int g_copy_world = 1;      /* ranks sharing the host cores (set by ctbd_dist_init) */

/* host threads that pack / unpack the staging chunks: at most 8, and never more than this rank's share of the host cores -- with one
 * process per GPU, 8 ranks x 8 spinning OpenMP threads on a 32-core box made every host-struct call 15x slower (round-1 record) */
int host_copy_threads()
{
	const char* env = getenv("CTB_COPY_THREADS");
	int nt = (env != nullptr) ? atoi(env) : std::min(8, std::max(1, omp_get_num_procs() / std::max(1, g_copy_world)));
	if (nt < 1) { nt = 1; }
	return nt;
}
} // namespace

int ctbd_h2d_blocks(void* dptr, int nblk, const void* const* hptrs, const int64_t* dst_off, const int64_t* nbytes)
{
	CTBD_REQUIRE_INIT();
	CTBD_CUDA(cudaSetDevice(rt().device));      /* may run on a helper thread of the host side (the current device is per thread) */
	if (ring_init() < 0) { return -1; }
	std::vector<Piece> pieces; std::vector<Chunk> chunks;
	plan_chunks(nblk, dst_off, nbytes, pieces, chunks);
	const int nt = host_copy_threads();
	for (size_t ci = 0; ci < chunks.size(); ci++)
	{
		const int slot = (int)(ci % STAGE_SLOTS);
		if (ring_wait(slot) < 0) { return -1; }
		char* stage = (char*)g_ring.buf[slot];
		const long p0 = (long)chunks[ci].p0, p1 = (long)chunks[ci].p1;
		/* pack: pageable host blocks -> pinned chunk, in parallel (the previous chunk is on the copy engine meanwhile) */
		#pragma omp parallel for schedule(dynamic, 2) num_threads(nt) if (p1 - p0 > 2)
		for (long q = p0; q < p1; q++) { memcpy(stage + pieces[q].pos, (const char*)hptrs[pieces[q].b] + pieces[q].boff, (size_t)pieces[q].n); }
		CTBD_CUDA(cudaMemcpyAsync((char*)dptr + chunks[ci].dev, stage, (size_t)chunks[ci].len, cudaMemcpyHostToDevice, rt().stream));
		CTBD_CUDA(cudaEventRecord(g_ring.ev[slot], rt().stream));
		g_ring.busy[slot] = true;
	}
	/* the ring stays owned by this layer, so the copies may complete asynchronously */
	return 0;
}

int ctbd_d2h_blocks(const void* dptr, int nblk, void* const* hptrs, const int64_t* src_off, const int64_t* nbytes)
{
	CTBD_REQUIRE_INIT();
	if (ring_init() < 0) { return -1; }
	std::vector<Piece> pieces; std::vector<Chunk> chunks;
	plan_chunks(nblk, src_off, nbytes, pieces, chunks);
	const int nt = host_copy_threads();
	auto issue = [&](size_t ci) -> int {
		const int slot = (int)(ci % STAGE_SLOTS);
		if (ring_wait(slot) < 0) { return -1; }
		CTBD_CUDA(cudaMemcpyAsync(g_ring.buf[slot], (const char*)dptr + chunks[ci].dev, (size_t)chunks[ci].len, cudaMemcpyDeviceToHost, rt().stream));
		CTBD_CUDA(cudaEventRecord(g_ring.ev[slot], rt().stream));
		g_ring.busy[slot] = true;
		return 0;
	};
	/* the copy engine runs STAGE_SLOTS - 1 chunks ahead of the unpacking threads */
	for (size_t ci = 0; ci + 1 < (size_t)STAGE_SLOTS && ci < chunks.size(); ci++) { if (issue(ci) < 0) { return -1; } }
	for (size_t ci = 0; ci < chunks.size(); ci++)
	{
		const int slot = (int)(ci % STAGE_SLOTS);
		if (ci + STAGE_SLOTS - 1 < chunks.size() && issue(ci + STAGE_SLOTS - 1) < 0) { return -1; }
		if (ring_wait(slot) < 0) { return -1; }
		const char* stage = (const char*)g_ring.buf[slot];
		const long p0 = (long)chunks[ci].p0, p1 = (long)chunks[ci].p1;
		/* unpack in parallel: the destination blocks are freshly allocated, so this is where their pages get faulted in */
		#pragma omp parallel for schedule(dynamic, 2) num_threads(nt) if (p1 - p0 > 2)
		for (long q = p0; q < p1; q++) { memcpy((char*)hptrs[pieces[q].b] + pieces[q].boff, stage + pieces[q].pos, (size_t)pieces[q].n); }
	}
	return 0;
}

/* Fault in the pages of freshly allocated host blocks with all copy threads (MADV_POPULATE_WRITE where the kernel has it, else one
 * write per page).  Called while the device is still computing the result that will land in these blocks, so the page faults of
 * the result download are off the critical path. */
int ctbd_host_prefault(int nblk, void* const* hptrs, const int64_t* nbytes)
{
	std::vector<std::pair<char*, int64_t>> pieces;
	const int64_t PIECE = (int64_t)2 << 20;
	for (int b = 0; b < nblk; b++) {
		for (int64_t off = 0; off < nbytes[b]; off += PIECE) { pieces.push_back(std::make_pair((char*)hptrs[b] + off, std::min<int64_t>(PIECE, nbytes[b] - off))); }
	}
	const long np = (long)pieces.size();
	const int nt = host_copy_threads();
	static int have_populate = 1;
	#pragma omp parallel for schedule(dynamic, 4) num_threads(nt) if (np > 8)
	for (long q = 0; q < np; q++) {
		char* p0 = pieces[q].first; char* p1 = p0 + pieces[q].second;
		bool done = false;
#ifdef MADV_POPULATE_WRITE
		if (have_populate) {
			char* a0 = (char*)(((uintptr_t)p0 + 4095) & ~(uintptr_t)4095);
			char* a1 = (char*)((uintptr_t)p1 & ~(uintptr_t)4095);
			if (a1 > a0) {
				if (madvise(a0, (size_t)(a1 - a0), MADV_POPULATE_WRITE) == 0) { done = true; p0[0] = 0; p1[-1] = 0; }
				else { have_populate = 0; }
			}
		}
#endif
		if (!done) { for (char* p = p0; p < p1; p += 4096) { *(volatile char*)p = 0; } p1[-1] = 0; }
	}
	return 0;
}

/* ---- distributed exchange: NCCL bound at run time, or a host callback ---- */
namespace {
struct NcclApi
{
	void* lib = nullptr;
	int (*GetUniqueId)(void*) = nullptr;
	int (*CommInitRank)(void**, int, ncclUniqueIdBytes, int) = nullptr;
	int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
	int (*CommDestroy)(void*) = nullptr;
	const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
void* g_comm = nullptr;
int g_rank = 0, g_world = 1;
ctbd_allgather_fn g_ag_fn = nullptr;
void* g_ag_ctx = nullptr;

int nccl_load()
{
	if (g_nccl.lib != nullptr) { return 0; }
	const char* names[] = { "libnccl.so.2", "libnccl.so" };
	for (const char* n : names) { g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (g_nccl.lib != nullptr) { break; } }
	if (g_nccl.lib == nullptr) { return fail_msg("NCCL: libnccl.so.2 not found (needed for more than one GPU)"); }
	g_nccl.GetUniqueId    = (int (*)(void*))dlsym(g_nccl.lib, "ncclGetUniqueId");
	g_nccl.CommInitRank   = (int (*)(void**, int, ncclUniqueIdBytes, int))dlsym(g_nccl.lib, "ncclCommInitRank");
	g_nccl.AllGather      = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(g_nccl.lib, "ncclAllGather");
	g_nccl.CommDestroy    = (int (*)(void*))dlsym(g_nccl.lib, "ncclCommDestroy");
	g_nccl.GetErrorString = (const char* (*)(int))dlsym(g_nccl.lib, "ncclGetErrorString");
	if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllGather || !g_nccl.CommDestroy) { return fail_msg("NCCL: missing symbols in libnccl"); }
	return 0;
}
int nccl_fail(const char* what, int rc)
{
	snprintf(rt().err, sizeof(rt().err), "%s: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "NCCL error");
	return -1;
}
} // namespace

namespace {
struct PeerBuffer { void* local = nullptr; std::vector<void*> ptrs; };

/* Barrier over peer-mapped arrival flags: rank r writes the epoch into slot r of every rank's flag array (posted NVLink stores, after a
 * system-scope fence) and waits until all slots of its own array have reached the epoch.  One tiny kernel on the stream instead of a
 * collective launch; the kernel boundary before it orders the peer stores of the preceding launch ahead of the arrival flag. */
PeerBuffer* g_flags = nullptr;
bool g_flags_failed = false;
unsigned long long g_epoch = 0;
struct FlagPtrs { unsigned long long* p[8]; };

__global__ void flag_barrier_kernel(FlagPtrs peers, int rank, int world, unsigned long long epoch)
{
	const int p = threadIdx.x;
	if (p < world) {
		__threadfence_system();
		volatile unsigned long long* arrive = peers.p[p] + rank;
		*arrive = epoch;
		volatile unsigned long long* seen = peers.p[rank] + p;
		const long long t0 = clock64();
		/* a dead peer must not hang the device for ever, and the exchange must never continue with a partially filled buffer:
		 * after about 9 s the kernel traps, the stream reports the error and every later call of the layer returns < 0 */
		while (*seen < epoch) { if (clock64() - t0 > (1ll << 34)) { __trap(); } }
		__threadfence_system();
	}
}
}

int ctbd_dist_unique_id(void* id_out)
{
	if (nccl_load() < 0) { return -1; }
	const int rc = g_nccl.GetUniqueId(id_out);
	return rc == 0 ? 0 : nccl_fail("ncclGetUniqueId", rc);
}

int ctbd_dist_init(int rank, int world, const void* unique_id)
{
	CTBD_REQUIRE_INIT();
	if (world < 1 || rank < 0 || rank >= world) { return fail_msg("dist: bad rank / world"); }
	g_rank = rank; g_world = world; g_copy_world = world;
	if (world == 1 || unique_id == nullptr) { return 0; }      /* single rank, or the host supplies the collective by callback */
	if (nccl_load() < 0) { return -1; }
	ncclUniqueIdBytes id;
	memcpy(&id, unique_id, sizeof(id));
	CTBD_CUDA(cudaSetDevice(rt().device));
	const int rc = g_nccl.CommInitRank(&g_comm, world, id, rank);
	return rc == 0 ? 0 : nccl_fail("ncclCommInitRank", rc);
}

int ctbd_dist_set_allgather(ctbd_allgather_fn fn, void* ctx) { g_ag_fn = fn; g_ag_ctx = ctx; return 0; }

int ctbd_peer_buffer_destroy(void* handle);
int ctbd_dist_finalize(void)
{
	if (g_flags != nullptr) { PeerBuffer* f = g_flags; g_flags = nullptr; g_flags_failed = true; ctbd_peer_buffer_destroy(f); }
	g_flags_failed = false; g_epoch = 0;
	if (g_comm != nullptr) { cudaStreamSynchronize(rt().stream); g_nccl.CommDestroy(g_comm); g_comm = nullptr; }
	g_rank = 0; g_world = 1; g_copy_world = 1; g_ag_fn = nullptr; g_ag_ctx = nullptr;
	return 0;
}

int ctbd_allgather(const void* sendbuf, void* recvbuf, size_t bytes_per_rank)
{
	if (g_world == 1) {
		if (sendbuf != recvbuf) { CTBD_CUDA(cudaMemcpyAsync(recvbuf, sendbuf, bytes_per_rank, cudaMemcpyDeviceToDevice, rt().stream)); }
		return 0;
	}
	if (g_ag_fn != nullptr) { return g_ag_fn(g_ag_ctx, sendbuf, recvbuf, bytes_per_rank, (void*)rt().stream); }
	if (g_comm == nullptr) { return fail_msg("dist: no communicator (call ctbd_dist_init with the NCCL unique id, or register a callback)"); }
	const int rc = g_nccl.AllGather(sendbuf, recvbuf, bytes_per_rank, /* ncclUint8 */ 1, g_comm, rt().stream);
	rt().launches++;
	return rc == 0 ? 0 : nccl_fail("ncclAllGather", rc);
}

int ctbd_peer_buffer_create(size_t bytes, void** handle);

int ctbd_barrier(void)
{
	if (g_world == 1) { return 0; }
	if (g_comm != nullptr && g_ag_fn == nullptr && !g_flags_failed && g_world <= 8 && getenv("CTB_NCCL_BARRIER") == nullptr)
	{
		if (g_flags == nullptr) {
			void* hnd = nullptr;
			g_flags_failed = true;      /* the creation below runs barriers of its own (the NCCL form) */
			if (ctbd_peer_buffer_create(sizeof(unsigned long long) * 8, &hnd) == 0) { g_flags = (PeerBuffer*)hnd; g_flags_failed = false; g_epoch = 0; }
		}
		if (g_flags != nullptr) {
			FlagPtrs fp;
			for (int p = 0; p < 8; p++) { fp.p[p] = (unsigned long long*)g_flags->ptrs[(size_t)(p < g_world ? p : 0)]; }
			g_epoch++;
			flag_barrier_kernel<<<1, 32, 0, rt().stream>>>(fp, g_rank, g_world, g_epoch);
			CTBD_LAUNCH_CHECK();
			return 0;
		}
	}
	/* an 8-byte all-gather on the stream: completes on a rank only after every rank has reached it, and kernel boundaries make the
	 * peer stores issued before it visible system-wide */
	static void* scratch = nullptr;      /* sized for the largest supported world (8 ranks + the send slot), so a re-init with more ranks stays in bounds */
	if (scratch == nullptr) { if (ctbd_malloc(&scratch, (size_t)8 * 9) < 0) { return -1; } }
	if (g_world > 8) { return fail_msg("dist: barrier supports at most 8 ranks"); }
	return ctbd_allgather(scratch, (char*)scratch + 8, 8);
}



int ctbd_peer_buffer_create(size_t bytes, void** handle)
{
	CTBD_REQUIRE_INIT();
	*handle = nullptr;
	if (g_world == 1 || g_comm == nullptr) { return fail_msg("peer buffer: needs the NCCL communicator of ctbd_dist_init"); }
	if (getenv("CTB_NO_PEER") != nullptr) { return fail_msg("peer buffer: disabled by CTB_NO_PEER"); }
	PeerBuffer* pb = new PeerBuffer();
	pb->ptrs.assign((size_t)g_world, nullptr);
	/* plain cudaMalloc: pool memory is not exportable through legacy CUDA IPC */
	cudaError_t e = cudaMalloc(&pb->local, bytes > 0 ? bytes : 256);
	if (e != cudaSuccess) { delete pb; return fail("cudaMalloc (peer buffer)", e, __FILE__, __LINE__); }
	cudaMemset(pb->local, 0, bytes > 0 ? bytes : 256);      /* synchronous: done before any peer can learn the handle */
	cudaIpcMemHandle_t mine;
	e = cudaIpcGetMemHandle(&mine, pb->local);
	int ok = (e == cudaSuccess) ? 1 : 0;
	/* exchange the handles (and the success flags) with the collective itself */
	const size_t slot = 128;
	static_assert(sizeof(cudaIpcMemHandle_t) + sizeof(int) <= 128, "handle slot too small");
	std::vector<unsigned char> hbuf(slot * (size_t)(g_world + 1), 0);
	memcpy(hbuf.data(), &mine, sizeof(mine));
	memcpy(hbuf.data() + sizeof(mine), &ok, sizeof(int));
	void* dbuf = nullptr;
	if (ctbd_malloc(&dbuf, slot * (size_t)(g_world + 1)) < 0) { cudaFree(pb->local); delete pb; return -1; }
	int rc = ctbd_h2d(dbuf, hbuf.data(), slot);
	if (rc == 0) { rc = ctbd_allgather(dbuf, (char*)dbuf + slot, slot); }
	if (rc == 0) { rc = ctbd_d2h(hbuf.data() + slot, (char*)dbuf + slot, slot * (size_t)g_world); }
	ctbd_free(dbuf);
	if (rc < 0) { cudaFree(pb->local); delete pb; return -1; }
	bool all_ok = true;
	for (int p = 0; p < g_world; p++) { int f = 0; memcpy(&f, hbuf.data() + slot * (size_t)(p + 1) + sizeof(mine), sizeof(int)); all_ok = all_ok && (f == 1); }
	for (int p = 0; p < g_world && all_ok; p++) {
		if (p == g_rank) { pb->ptrs[p] = pb->local; continue; }
		cudaIpcMemHandle_t hp;
		memcpy(&hp, hbuf.data() + slot * (size_t)(p + 1), sizeof(hp));
		e = cudaIpcOpenMemHandle(&pb->ptrs[p], hp, cudaIpcMemLazyEnablePeerAccess);
		if (e != cudaSuccess) { all_ok = false; (void)cudaGetLastError(); }
	}
	/* every rank must take the same decision: agree on success with one more tiny all-gather */
	{
		int flag = all_ok ? 1 : 0;
		std::vector<int> flags((size_t)g_world * 2 + 2, 0);
		void* dflag = nullptr;
		if (ctbd_malloc(&dflag, 8 * (size_t)(g_world + 1)) < 0) { all_ok = false; }
		else {
			long long f64 = flag;
			std::vector<long long> all((size_t)g_world + 1, 0);
			all[0] = f64;
			if (ctbd_h2d(dflag, all.data(), 8) < 0 || ctbd_allgather(dflag, (char*)dflag + 8, 8) < 0 || ctbd_d2h(all.data() + 1, (char*)dflag + 8, 8 * (size_t)g_world) < 0) { all_ok = false; }
			else { for (int p = 0; p < g_world; p++) { all_ok = all_ok && (all[(size_t)p + 1] == 1); } }
			ctbd_free(dflag);
		}
	}
	if (!all_ok) {
		for (int p = 0; p < g_world; p++) { if (p != g_rank && pb->ptrs[p] != nullptr) { cudaIpcCloseMemHandle(pb->ptrs[p]); } }
		cudaFree(pb->local); delete pb;
		return fail_msg("peer buffer: CUDA IPC mapping of the peers failed");
	}
	*handle = pb;
	return 0;
}

int ctbd_peer_buffer_ptrs(void* handle, void** ptrs)
{
	PeerBuffer* pb = (PeerBuffer*)handle;
	for (int p = 0; p < g_world; p++) { ptrs[p] = pb->ptrs[(size_t)p]; }
	return 0;
}

int ctbd_peer_buffer_destroy(void* handle)
{
	PeerBuffer* pb = (PeerBuffer*)handle;
	if (pb == nullptr) { return 0; }
	cudaStreamSynchronize(rt().stream);
	for (int p = 0; p < (int)pb->ptrs.size(); p++) { if (p != g_rank && pb->ptrs[(size_t)p] != nullptr) { cudaIpcCloseMemHandle(pb->ptrs[(size_t)p]); } }
	/* peers may still be unmapping: make sure nobody writes any more before the memory goes away */
	ctbd_barrier();
	cudaStreamSynchronize(rt().stream);
	cudaFree(pb->local);
	delete pb;
	return 0;
}

/* ---- NVSwitch multicast buffers ------------------------------------------------------------------------------------------
 * One allocation per rank (cuMemCreate), all bound to ONE multicast object (cuMulticastCreate / BindMem): a store to the multicast
 * address is replicated by the switch into the buffers of all ranks, so the all-gather of the sharded effective Hamiltonian costs
 * every GPU one copy of its slice on the wire instead of world - 1 (17 MB instead of 120 MB per matvec at D = 4096 on 8 GPUs).
 * The driver API is bound at run time (dlopen libcuda.so.1: the library must load on hosts without a driver); the multicast handle
 * travels from rank 0 to the other processes as a POSIX file descriptor over a unix-domain socket (SCM_RIGHTS). */
namespace {
typedef int CUres;
typedef unsigned long long CUhandle;      /* CUmemGenericAllocationHandle */
typedef unsigned long long CUdptr;        /* CUdeviceptr */
struct CuMcProp { unsigned int numDevices; size_t size; unsigned long long handleTypes; unsigned long long flags; };      /* CUmulticastObjectProp */
struct CuLocation { int type; int id; };                                                                                   /* CUmemLocation */
struct CuAllocProp { int type; int requestedHandleTypes; CuLocation location; void* win32HandleMetaData;                  /* CUmemAllocationProp */
	struct { unsigned char compressionType, gpuDirectRDMACapable; unsigned short usage; unsigned char reserved[4]; } allocFlags; };
struct CuAccessDesc { CuLocation location; int flags; };                                                                   /* CUmemAccessDesc */
struct CudaDrv
{
	void* lib = nullptr;
	CUres (*DeviceGet)(int*, int) = nullptr;
	CUres (*DeviceGetAttribute)(int*, int, int) = nullptr;
	CUres (*MulticastCreate)(CUhandle*, const CuMcProp*) = nullptr;
	CUres (*MulticastAddDevice)(CUhandle, int) = nullptr;
	CUres (*MulticastBindMem)(CUhandle, size_t, CUhandle, size_t, size_t, unsigned long long) = nullptr;
	CUres (*MulticastUnbind)(CUhandle, int, size_t, size_t) = nullptr;
	CUres (*MulticastGetGranularity)(size_t*, const CuMcProp*, int) = nullptr;
	CUres (*MemCreate)(CUhandle*, size_t, const CuAllocProp*, unsigned long long) = nullptr;
	CUres (*MemRelease)(CUhandle) = nullptr;
	CUres (*MemGetAllocationGranularity)(size_t*, const CuAllocProp*, int) = nullptr;
	CUres (*MemAddressReserve)(CUdptr*, size_t, size_t, CUdptr, unsigned long long) = nullptr;
	CUres (*MemAddressFree)(CUdptr, size_t) = nullptr;
	CUres (*MemMap)(CUdptr, size_t, size_t, CUhandle, unsigned long long) = nullptr;
	CUres (*MemUnmap)(CUdptr, size_t) = nullptr;
	CUres (*MemSetAccess)(CUdptr, size_t, const CuAccessDesc*, size_t) = nullptr;
	CUres (*MemExportToShareableHandle)(void*, CUhandle, int, unsigned long long) = nullptr;
	CUres (*MemImportFromShareableHandle)(CUhandle*, void*, int) = nullptr;
};
CudaDrv g_drv;
int drv_load()
{
	if (g_drv.lib != nullptr) { return 0; }
	g_drv.lib = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
	if (g_drv.lib == nullptr) { return fail_msg("multicast: libcuda.so.1 not found"); }
	bool ok = true;
	auto sym = [&](const char* n) { void* f = dlsym(g_drv.lib, n); if (f == nullptr) { ok = false; } return f; };
	g_drv.DeviceGet = (decltype(g_drv.DeviceGet))sym("cuDeviceGet");
	g_drv.DeviceGetAttribute = (decltype(g_drv.DeviceGetAttribute))sym("cuDeviceGetAttribute");
	g_drv.MulticastCreate = (decltype(g_drv.MulticastCreate))sym("cuMulticastCreate");
	g_drv.MulticastAddDevice = (decltype(g_drv.MulticastAddDevice))sym("cuMulticastAddDevice");
	g_drv.MulticastBindMem = (decltype(g_drv.MulticastBindMem))sym("cuMulticastBindMem");
	g_drv.MulticastUnbind = (decltype(g_drv.MulticastUnbind))sym("cuMulticastUnbind");
	g_drv.MulticastGetGranularity = (decltype(g_drv.MulticastGetGranularity))sym("cuMulticastGetGranularity");
	g_drv.MemCreate = (decltype(g_drv.MemCreate))sym("cuMemCreate");
	g_drv.MemRelease = (decltype(g_drv.MemRelease))sym("cuMemRelease");
	g_drv.MemGetAllocationGranularity = (decltype(g_drv.MemGetAllocationGranularity))sym("cuMemGetAllocationGranularity");
	g_drv.MemAddressReserve = (decltype(g_drv.MemAddressReserve))sym("cuMemAddressReserve");
	g_drv.MemAddressFree = (decltype(g_drv.MemAddressFree))sym("cuMemAddressFree");
	g_drv.MemMap = (decltype(g_drv.MemMap))sym("cuMemMap");
	g_drv.MemUnmap = (decltype(g_drv.MemUnmap))sym("cuMemUnmap");
	g_drv.MemSetAccess = (decltype(g_drv.MemSetAccess))sym("cuMemSetAccess");
	g_drv.MemExportToShareableHandle = (decltype(g_drv.MemExportToShareableHandle))sym("cuMemExportToShareableHandle");
	g_drv.MemImportFromShareableHandle = (decltype(g_drv.MemImportFromShareableHandle))sym("cuMemImportFromShareableHandle");
	if (!ok) { dlclose(g_drv.lib); g_drv.lib = nullptr; return fail_msg("multicast: the driver lacks the multicast / virtual memory entry points"); }
	return 0;
}

/* every rank contributes 'slot' bytes; all[p * slot ...] = contribution of rank p */
int small_allgather(const void* mine, size_t slot, std::vector<unsigned char>& all)
{
	all.assign(slot * (size_t)g_world, 0);
	void* dbuf = nullptr;
	if (ctbd_malloc(&dbuf, slot * (size_t)(g_world + 1)) < 0) { return -1; }
	int rc = ctbd_h2d(dbuf, mine, slot);
	if (rc == 0) { rc = ctbd_allgather(dbuf, (char*)dbuf + slot, slot); }
	if (rc == 0) { rc = ctbd_d2h(all.data(), (char*)dbuf + slot, slot * (size_t)g_world); }
	ctbd_free(dbuf);
	return rc;
}
/* all ranks agree on success */
bool all_agree(bool mine)
{
	long long f = mine ? 1 : 0;
	std::vector<unsigned char> all;
	if (small_allgather(&f, 8, all) < 0) { return false; }
	bool ok = true;
	for (int p = 0; p < g_world; p++) { long long v; memcpy(&v, all.data() + 8 * (size_t)p, 8); ok = ok && (v == 1); }
	return ok;
}

int send_fd(int sock, int fd)
{
	struct msghdr msg; memset(&msg, 0, sizeof(msg));
	char cbuf[CMSG_SPACE(sizeof(int))]; memset(cbuf, 0, sizeof(cbuf));
	char data = 'f';
	struct iovec io; io.iov_base = &data; io.iov_len = 1;
	msg.msg_iov = &io; msg.msg_iovlen = 1; msg.msg_control = cbuf; msg.msg_controllen = sizeof(cbuf);
	struct cmsghdr* cm = CMSG_FIRSTHDR(&msg);
	cm->cmsg_level = SOL_SOCKET; cm->cmsg_type = SCM_RIGHTS; cm->cmsg_len = CMSG_LEN(sizeof(int));
	memcpy(CMSG_DATA(cm), &fd, sizeof(int));
	return sendmsg(sock, &msg, 0) == 1 ? 0 : -1;
}
int recv_fd(int sock)
{
	struct msghdr msg; memset(&msg, 0, sizeof(msg));
	char cbuf[CMSG_SPACE(sizeof(int))]; memset(cbuf, 0, sizeof(cbuf));
	char data = 0;
	struct iovec io; io.iov_base = &data; io.iov_len = 1;
	msg.msg_iov = &io; msg.msg_iovlen = 1; msg.msg_control = cbuf; msg.msg_controllen = sizeof(cbuf);
	if (recvmsg(sock, &msg, 0) != 1) { return -1; }
	struct cmsghdr* cm = CMSG_FIRSTHDR(&msg);
	if (cm == nullptr || cm->cmsg_level != SOL_SOCKET || cm->cmsg_type != SCM_RIGHTS) { return -1; }
	int fd = -1;
	memcpy(&fd, CMSG_DATA(cm), sizeof(int));
	return fd;
}

struct McBuffer
{
	CUhandle mc = 0, mem = 0;
	CUdptr local = 0, mcva = 0;
	size_t size = 0;
	int dev = 0;
	bool bound = false;
};
void mc_release(McBuffer* b)
{
	if (b == nullptr) { return; }
	if (b->mcva != 0) { g_drv.MemUnmap(b->mcva, b->size); g_drv.MemAddressFree(b->mcva, b->size); }
	if (b->bound) { g_drv.MulticastUnbind(b->mc, b->dev, 0, b->size); }
	if (b->local != 0) { g_drv.MemUnmap(b->local, b->size); g_drv.MemAddressFree(b->local, b->size); }
	if (b->mem != 0) { g_drv.MemRelease(b->mem); }
	if (b->mc != 0) { g_drv.MemRelease(b->mc); }
	delete b;
}
int g_mc_seq = 0;
} // namespace

/* collective; on success *local_ptr is this rank's buffer and *mc_ptr the multicast address (only multimem stores may touch it).
 * Returns < 0 on every rank when any rank cannot take part (no NVSwitch multicast, no fabric manager, ...). */
int ctbd_mc_buffer_create(size_t bytes, void** handle, void** local_ptr, void** mc_ptr)
{
	CTBD_REQUIRE_INIT();
	*handle = nullptr; *local_ptr = nullptr; *mc_ptr = nullptr;
	if (g_world == 1 || g_comm == nullptr) { return fail_msg("multicast buffer: needs the NCCL communicator of ctbd_dist_init"); }
	if (getenv("CTB_NO_MULTICAST") != nullptr || getenv("CTB_NO_PEER") != nullptr) { return fail_msg("multicast buffer: disabled by the environment"); }
	CTBD_CUDA(cudaSetDevice(rt().device));
	CTBD_CUDA(cudaFree(nullptr));
	bool ok = (drv_load() == 0);
	McBuffer* b = new McBuffer();
	int dev = 0, supported = 0;
	if (ok) { ok = g_drv.DeviceGet(&dev, rt().device) == 0 && g_drv.DeviceGetAttribute(&supported, /* CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED */ 132, dev) == 0 && supported != 0; }
	b->dev = dev;
	ok = all_agree(ok);
	if (!ok) { mc_release(b); return fail_msg("multicast buffer: multicast is not supported on every device of the team"); }

	const int FD = 1;      /* CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR */
	CuMcProp mp; memset(&mp, 0, sizeof(mp));
	mp.numDevices = (unsigned)g_world; mp.handleTypes = FD; mp.size = bytes > 0 ? bytes : 1;
	size_t gran = 0, mgran = 0;
	CuAllocProp ap; memset(&ap, 0, sizeof(ap));
	ap.type = 1 /* PINNED */; ap.requestedHandleTypes = FD; ap.location.type = 1 /* DEVICE */; ap.location.id = dev;
	ok = g_drv.MulticastGetGranularity(&gran, &mp, /* RECOMMENDED */ 1) == 0 && g_drv.MemGetAllocationGranularity(&mgran, &ap, /* RECOMMENDED */ 1) == 0 && gran > 0 && mgran > 0;
	if (ok) { if (mgran > gran) { gran = mgran; } b->size = ((mp.size + gran - 1) / gran) * gran; mp.size = b->size; }

	/* rank 0 creates the multicast object and hands its descriptor to the other processes */
	char sockname[64]; memset(sockname, 0, sizeof(sockname));
	int lsock = -1, fd0 = -1;
	if (g_rank == 0 && ok)
	{
		ok = g_drv.MulticastCreate(&b->mc, &mp) == 0 && g_drv.MemExportToShareableHandle(&fd0, b->mc, FD, 0) == 0;
		if (ok) {
			snprintf(sockname + 1, sizeof(sockname) - 2, "ctb_mc_%d_%d", (int)getpid(), g_mc_seq++);      /* abstract socket: leading NUL */
			lsock = socket(AF_UNIX, SOCK_STREAM, 0);
			struct sockaddr_un sa; memset(&sa, 0, sizeof(sa)); sa.sun_family = AF_UNIX;
			memcpy(sa.sun_path, sockname, sizeof(sockname));
			ok = lsock >= 0 && bind(lsock, (struct sockaddr*)&sa, (socklen_t)(offsetof(struct sockaddr_un, sun_path) + 1 + strlen(sockname + 1))) == 0 && listen(lsock, 16) == 0;
		}
	}
	struct { char name[64]; long long ok; } slot; memset(&slot, 0, sizeof(slot));
	memcpy(slot.name, sockname, sizeof(sockname)); slot.ok = ok ? 1 : 0;
	std::vector<unsigned char> all;
	if (small_allgather(&slot, sizeof(slot), all) < 0) { ok = false; }
	for (int p = 0; p < g_world && ok; p++) { long long f; memcpy(&f, all.data() + sizeof(slot) * (size_t)p + 64, 8); ok = (f == 1); }
	if (ok && g_rank == 0) {
		for (int p = 1; p < g_world && ok; p++) {
			const int c = accept(lsock, nullptr, nullptr);
			ok = (c >= 0) && send_fd(c, fd0) == 0;
			if (c >= 0) { close(c); }
		}
	}
	else if (ok) {
		struct sockaddr_un sa; memset(&sa, 0, sizeof(sa)); sa.sun_family = AF_UNIX;
		memcpy(sa.sun_path, all.data(), 64);
		const socklen_t len = (socklen_t)(offsetof(struct sockaddr_un, sun_path) + 1 + strlen((const char*)all.data() + 1));
		const int c = socket(AF_UNIX, SOCK_STREAM, 0);
		int fd = -1;
		ok = (c >= 0);
		for (int attempt = 0; ok && attempt < 2000; attempt++) {
			if (connect(c, (struct sockaddr*)&sa, len) == 0) { fd = recv_fd(c); break; }
			usleep(1000);
		}
		if (c >= 0) { close(c); }
		ok = ok && fd >= 0 && g_drv.MemImportFromShareableHandle(&b->mc, (void*)(uintptr_t)fd, FD) == 0;
		if (fd >= 0) { close(fd); }
	}
	if (lsock >= 0) { close(lsock); }
	if (fd0 >= 0) { close(fd0); }
	ok = all_agree(ok);
	if (!ok) { mc_release(b); return fail_msg("multicast buffer: could not create / share the multicast object"); }

	ok = g_drv.MulticastAddDevice(b->mc, dev) == 0;
	ok = all_agree(ok);      /* every device has to be added before memory is bound */
	if (ok) { ok = g_drv.MemCreate(&b->mem, b->size, &ap, 0) == 0; }
	if (ok) { ok = g_drv.MulticastBindMem(b->mc, 0, b->mem, 0, b->size, 0) == 0; b->bound = ok; }
	CuAccessDesc ad; ad.location.type = 1; ad.location.id = dev; ad.flags = 3 /* READWRITE */;
	if (ok) { ok = g_drv.MemAddressReserve(&b->local, b->size, gran, 0, 0) == 0 && g_drv.MemMap(b->local, b->size, 0, b->mem, 0) == 0 && g_drv.MemSetAccess(b->local, b->size, &ad, 1) == 0; }
	ok = all_agree(ok);      /* all ranks bound: the multicast address may be mapped */
	if (ok) { ok = g_drv.MemAddressReserve(&b->mcva, b->size, gran, 0, 0) == 0 && g_drv.MemMap(b->mcva, b->size, 0, b->mc, 0) == 0 && g_drv.MemSetAccess(b->mcva, b->size, &ad, 1) == 0; }
	if (ok) { ok = cudaMemset((void*)b->local, 0, b->size) == cudaSuccess && cudaDeviceSynchronize() == cudaSuccess; }
	ok = all_agree(ok);
	if (!ok) { (void)cudaGetLastError(); mc_release(b); return fail_msg("multicast buffer: binding / mapping failed"); }
	*handle = b; *local_ptr = (void*)b->local; *mc_ptr = (void*)b->mcva;
	return 0;
}

int ctbd_mc_buffer_destroy(void* handle)
{
	McBuffer* b = (McBuffer*)handle;
	if (b == nullptr) { return 0; }
	cudaStreamSynchronize(rt().stream);
	ctbd_barrier();      /* nobody stores to the team any more */
	cudaStreamSynchronize(rt().stream);
	mc_release(b);
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

/* ---- CUDA graphs of launch sequences (include/ctb_device.h) ---- */
extern "C" int ctbd_graph_capture_begin(void)
{
	CTBD_REQUIRE_INIT();
	if (getenv("CTB_NO_GRAPH") != nullptr) { return 1; }
	const cudaError_t e = cudaStreamBeginCapture(rt().stream, cudaStreamCaptureModeThreadLocal);
	if (e != cudaSuccess) { (void)cudaGetLastError(); return 1; }
	return 0;
}

extern "C" int ctbd_graph_capture_end(void** graph)
{
	*graph = nullptr;
	cudaGraph_t g = nullptr;
	cudaError_t e = cudaStreamEndCapture(rt().stream, &g);
	if (e != cudaSuccess || g == nullptr) { (void)cudaGetLastError(); if (g != nullptr) { cudaGraphDestroy(g); } return 1; }
	cudaGraphExec_t ex = nullptr;
	e = cudaGraphInstantiate(&ex, g, 0);
	cudaGraphDestroy(g);
	if (e != cudaSuccess || ex == nullptr) { (void)cudaGetLastError(); return 1; }
	*graph = (void*)ex;
	return 0;
}

extern "C" int ctbd_graph_launch(void* graph)
{
	CTBD_CUDA(cudaGraphLaunch((cudaGraphExec_t)graph, rt().stream));
	rt().launches++;
	return 0;
}

extern "C" int ctbd_graph_destroy(void* graph)
{
	if (graph != nullptr) { cudaGraphExecDestroy((cudaGraphExec_t)graph); }
	return 0;
}
