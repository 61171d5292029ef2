/*
 * ctbd_remap.cu -- logical-index remaps between packed block-sparse layouts resident on the device.
 *
 * One gather kernel covers the structural primitives of the reference that move entries between
 * block structures (src/tensor/block_sparse_tensor.c): _transpose :785 (incl. the conjugating
 * variant :880), _flatten_axes :950, _split_axis :1123, _slice :1446 and
 * _multiply_pointwise_vector :1654.  Every destination entry computes its logical multi-index,
 * maps it to the source logical multi-index of the operation and reads the source entry through
 * the source sector tables; writes are fully coalesced (destination order), reads are coalesced
 * along the innermost source run.  HBM-bound: 2 x sizeof(T) bytes per stored entry.
 */
#include <vector>
#include <algorithm>
#include "ctbd_common.cuh"

namespace ctbd {

struct LayoutDev
{
	int ndim, dtype;
	int64_t dim[CTBD_MAXDIM];
	int nsec[CTBD_MAXDIM];
	const int32_t* sec_of[CTBD_MAXDIM];
	const int32_t* pos_of[CTBD_MAXDIM];
	const int32_t* secstart[CTBD_MAXDIM];
	const int32_t* log_of[CTBD_MAXDIM];
	int64_t ngrid;
	const int64_t* grid_off;
	int nblk;
	const int64_t* blk_grid;
	const int64_t* blk_off;
	int64_t nstore;
};

struct Layout
{
	LayoutDev d;
	void* arena;    /* one device allocation holding all tables */
};

struct RemapParams
{
	int op, i_ax, conj, scale_ax;
	int perm[CTBD_MAXDIM];
	const int64_t* ind;
	const double* scale;
};

template <typename T> __device__ __forceinline__ T conj_if(T v, int) { return v; }
template <> __device__ __forceinline__ double2 conj_if<double2>(double2 v, int c) { if (c) { v.y = -v.y; } return v; }
__device__ __forceinline__ double  scale_by(double v, double s)  { return v * s; }
__device__ __forceinline__ double2 scale_by(double2 v, double s) { return make_double2(v.x * s, v.y * s); }
template <typename T> __device__ __forceinline__ T zero_of();
template <> __device__ __forceinline__ double zero_of<double>() { return 0.0; }
template <> __device__ __forceinline__ double2 zero_of<double2>() { return make_double2(0.0, 0.0); }

template <typename T>
__global__ void __launch_bounds__(256) remap_kernel(const LayoutDev D, const LayoutDev S, const RemapParams p, T* __restrict__ dst, const T* __restrict__ src)
{
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < D.nstore; e += stride)
	{
		/* stored block containing entry e */
		int lo = 0, hi = D.nblk - 1;
		while (lo < hi) {
			const int mid = (lo + hi + 1) >> 1;
			if (D.blk_off[mid] <= e) { lo = mid; } else { hi = mid - 1; }
		}
		int64_t cell = D.blk_grid[lo];
		int64_t r = e - D.blk_off[lo];
		int sec[CTBD_MAXDIM];
		#pragma unroll
		for (int i = CTBD_MAXDIM - 1; i >= 0; i--) {
			if (i < D.ndim) { sec[i] = (int)(cell % D.nsec[i]); cell /= D.nsec[i]; }
		}
		int64_t ld[CTBD_MAXDIM], ls[CTBD_MAXDIM];
		#pragma unroll
		for (int i = CTBD_MAXDIM - 1; i >= 0; i--) {
			if (i < D.ndim) {
				const int s0 = D.secstart[i][sec[i]];
				const int bd = D.secstart[i][sec[i] + 1] - s0;
				const int pos = (int)(r % bd); r /= bd;
				ld[i] = D.log_of[i][s0 + pos];
			}
		}
		switch (p.op)
		{
			case CTBD_REMAP_TRANSPOSE:
				#pragma unroll
				for (int i = 0; i < CTBD_MAXDIM; i++) { if (i < D.ndim) { ls[p.perm[i]] = ld[i]; } }
				break;
			case CTBD_REMAP_FLATTEN:
				#pragma unroll
				for (int i = 0; i < CTBD_MAXDIM; i++) {
					if (i < D.ndim) {
						if (i < p.i_ax) { ls[i] = ld[i]; }
						else if (i == p.i_ax) { ls[i] = ld[i] / S.dim[i + 1]; ls[i + 1] = ld[i] % S.dim[i + 1]; }
						else { ls[i + 1] = ld[i]; }
					}
				}
				break;
			case CTBD_REMAP_SPLIT:
				#pragma unroll
				for (int i = 0; i < CTBD_MAXDIM; i++) {
					if (i < D.ndim) {
						if (i < p.i_ax) { ls[i] = ld[i]; }
						else if (i == p.i_ax) { ls[i] = ld[i] * D.dim[i + 1]; }
						else if (i == p.i_ax + 1) { ls[i - 1] += ld[i]; }
						else { ls[i - 1] = ld[i]; }
					}
				}
				break;
			case CTBD_REMAP_SLICE:
				#pragma unroll
				for (int i = 0; i < CTBD_MAXDIM; i++) { if (i < D.ndim) { ls[i] = (i == p.i_ax) ? p.ind[ld[i]] : ld[i]; } }
				break;
			default:
				#pragma unroll
				for (int i = 0; i < CTBD_MAXDIM; i++) { if (i < D.ndim) { ls[i] = ld[i]; } }
				break;
		}
		int64_t scell = 0, soff = 0;
		#pragma unroll
		for (int i = 0; i < CTBD_MAXDIM; i++) {
			if (i < S.ndim) {
				const int s = S.sec_of[i][ls[i]];
				scell = scell * S.nsec[i] + s;
				soff = soff * (S.secstart[i][s + 1] - S.secstart[i][s]) + S.pos_of[i][ls[i]];
			}
		}
		const int64_t sbase = S.grid_off[scell];
		T v = (sbase >= 0) ? src[sbase + soff] : zero_of<T>();
		v = conj_if<T>(v, p.conj);
		if (p.scale_ax >= 0) { v = scale_by(v, p.scale[ld[p.scale_ax]]); }
		dst[e] = v;
	}
}

/* scatter form (CTBD_REMAP_UNSLICE): every SOURCE entry is written to the destination entry whose index on axis i_ax is
 * ind[source index]; used to merge the column slices computed by different GPUs back into the full tensor */
template <typename T>
__global__ void __launch_bounds__(256) unslice_kernel(const LayoutDev D, const LayoutDev S, const RemapParams p, T* __restrict__ dst, const T* __restrict__ src)
{
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < S.nstore; e += stride)
	{
		int lo = 0, hi = S.nblk - 1;
		while (lo < hi) {
			const int mid = (lo + hi + 1) >> 1;
			if (S.blk_off[mid] <= e) { lo = mid; } else { hi = mid - 1; }
		}
		int64_t cell = S.blk_grid[lo];
		int64_t r = e - S.blk_off[lo];
		int sec[CTBD_MAXDIM];
		#pragma unroll
		for (int i = CTBD_MAXDIM - 1; i >= 0; i--) {
			if (i < S.ndim) { sec[i] = (int)(cell % S.nsec[i]); cell /= S.nsec[i]; }
		}
		int64_t ld[CTBD_MAXDIM];
		#pragma unroll
		for (int i = CTBD_MAXDIM - 1; i >= 0; i--) {
			if (i < S.ndim) {
				const int s0 = S.secstart[i][sec[i]];
				const int bd = S.secstart[i][sec[i] + 1] - s0;
				const int pos = (int)(r % bd); r /= bd;
				const int64_t ls = S.log_of[i][s0 + pos];
				ld[i] = (i == p.i_ax) ? p.ind[ls] : ls;
			}
		}
		int64_t dcell = 0, doff = 0;
		#pragma unroll
		for (int i = 0; i < CTBD_MAXDIM; i++) {
			if (i < D.ndim) {
				const int s = D.sec_of[i][ld[i]];
				dcell = dcell * D.nsec[i] + s;
				doff = doff * (D.secstart[i][s + 1] - D.secstart[i][s]) + D.pos_of[i][ld[i]];
			}
		}
		const int64_t dbase = D.grid_off[dcell];
		if (dbase >= 0) { dst[dbase + doff] = src[e]; }
	}
}

/* ---- batched strided 2-d copy ---- */
struct CopyTile { int32_t desc, row0; };
struct CopyPlan { int dtype; int ndesc, ntiles; ctbd_copy2d* descs; CopyTile* tiles; };
static constexpr int COPY_ROWS = 16;      /* rows per tile: 8 warps x 2 rows */

template <typename T>
__global__ void __launch_bounds__(256) copy2d_kernel(int ntiles, const CopyTile* __restrict__ tiles, const ctbd_copy2d* __restrict__ descs,
	const T* __restrict__ src, T* __restrict__ dst)
{
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	for (int t = blockIdx.x; t < ntiles; t += gridDim.x)
	{
		const CopyTile tl = tiles[t];
		const ctbd_copy2d d = descs[tl.desc];
		const int r1 = min(tl.row0 + COPY_ROWS, d.rows);
		for (int i = tl.row0 + warp; i < r1; i += 8) {
			const T* __restrict__ s = src + d.src_off + (int64_t)i * d.src_ld;
			T* __restrict__ o = dst + d.dst_off + (int64_t)i * d.dst_ld;
			for (int j = lane; j < d.cols; j += 32) { o[j] = s[j]; }
		}
	}
}

struct CopySrcs { const void* p[8]; int64_t stride; int vec_ok; };

/* source rows may live in peer memory: 16-byte loads where the run allows it, every lane keeps several loads in flight */
template <typename T>
__global__ void __launch_bounds__(256) copy2d_multi_kernel(int ntiles, const CopyTile* __restrict__ tiles, const ctbd_copy2d* __restrict__ descs,
	const CopySrcs srcs, T* __restrict__ dst)
{
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	for (int t = blockIdx.x; t < ntiles; t += gridDim.x)
	{
		const CopyTile tl = tiles[t];
		const ctbd_copy2d d = descs[tl.desc];
		const int64_t q = d.src_off / srcs.stride, off = d.src_off - q * srcs.stride;
		const T* __restrict__ sbase = reinterpret_cast<const T*>(srcs.p[q]) + off;
		const int r1 = min(tl.row0 + COPY_ROWS, d.rows);
		const bool vec = (sizeof(T) == 8) && srcs.vec_ok && ((d.cols & 1) == 0) && ((d.src_ld & 1) == 0) && ((d.dst_ld & 1) == 0) && ((off & 1) == 0) && ((d.dst_off & 1) == 0);
		for (int i = tl.row0 + warp; i < r1; i += 8) {
			const T* __restrict__ s = sbase + (int64_t)i * d.src_ld;
			T* __restrict__ o = dst + d.dst_off + (int64_t)i * d.dst_ld;
			if (vec) {
				const double2* __restrict__ s2 = reinterpret_cast<const double2*>(s);
				double2* __restrict__ o2 = reinterpret_cast<double2*>(o);
				for (int j = lane; j < d.cols / 2; j += 32) { o2[j] = s2[j]; }
			}
			else { for (int j = lane; j < d.cols; j += 32) { o[j] = s[j]; } }
		}
	}
}

struct CopyDsts { void* p[8]; int n; int vec_ok; };

/* one local read, the same strided block written to several (peer-mapped) destinations: 16 bytes per lane, whole row runs per warp */
template <typename T>
__global__ void __launch_bounds__(256) copy2d_push_kernel(int ntiles, const CopyTile* __restrict__ tiles, const ctbd_copy2d* __restrict__ descs,
	const T* __restrict__ src, const CopyDsts dsts)
{
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	for (int t = blockIdx.x; t < ntiles; t += gridDim.x)
	{
		const CopyTile tl = tiles[t];
		const ctbd_copy2d d = descs[tl.desc];
		const int r1 = min(tl.row0 + COPY_ROWS, d.rows);
		const bool vec = (sizeof(T) == 8) && dsts.vec_ok && ((d.cols & 1) == 0) && ((d.src_ld & 1) == 0) && ((d.dst_ld & 1) == 0) && ((d.src_off & 1) == 0) && ((d.dst_off & 1) == 0);
		for (int i = tl.row0 + warp; i < r1; i += 8) {
			const T* __restrict__ s = src + d.src_off + (int64_t)i * d.src_ld;
			const int64_t doff = d.dst_off + (int64_t)i * d.dst_ld;
			if (vec) {
				const double2* __restrict__ s2 = reinterpret_cast<const double2*>(s);
				for (int j = lane; j < d.cols / 2; j += 32) {
					const double2 v = s2[j];
					#pragma unroll
					for (int q = 0; q < 8; q++) { if (q < dsts.n) { reinterpret_cast<double2*>(reinterpret_cast<T*>(dsts.p[q]) + doff)[j] = v; } }
				}
			}
			else {
				for (int j = lane; j < d.cols; j += 32) {
					const T v = s[j];
					#pragma unroll
					for (int q = 0; q < 8; q++) { if (q < dsts.n) { (reinterpret_cast<T*>(dsts.p[q]) + doff)[j] = v; } }
				}
			}
		}
	}
}

} // namespace ctbd

using namespace ctbd;

extern "C" {

int ctbd_layout_create(const struct ctbd_layout_host* h, void** layout)
{
	CTBD_REQUIRE_INIT();
	Layout* L = new Layout();
	memset(&L->d, 0, sizeof(L->d));
	L->arena = nullptr;
	/* pack all tables into one host buffer (8-byte aligned pieces), upload once */
	size_t bytes = 0;
	auto reserve = [&](size_t n) { size_t off = bytes; bytes += (n + 7) & ~(size_t)7; return off; };
	size_t o_sec[CTBD_MAXDIM], o_pos[CTBD_MAXDIM], o_start[CTBD_MAXDIM], o_log[CTBD_MAXDIM];
	for (int i = 0; i < h->ndim; i++) {
		o_sec[i]   = reserve((size_t)h->dim[i] * 4);
		o_pos[i]   = reserve((size_t)h->dim[i] * 4);
		o_start[i] = reserve((size_t)(h->nsec[i] + 1) * 4);
		o_log[i]   = reserve((size_t)h->dim[i] * 4);
	}
	const size_t o_grid = reserve((size_t)h->ngrid * 8);
	const size_t o_bg   = reserve((size_t)(h->nblk > 0 ? h->nblk : 1) * 8);
	const size_t o_bo   = reserve((size_t)(h->nblk + 1) * 8);
	std::vector<unsigned char> buf(bytes > 0 ? bytes : 8, 0);
	for (int i = 0; i < h->ndim; i++) {
		memcpy(&buf[o_sec[i]],   h->sec_of[i],   (size_t)h->dim[i] * 4);
		memcpy(&buf[o_pos[i]],   h->pos_of[i],   (size_t)h->dim[i] * 4);
		memcpy(&buf[o_start[i]], h->secstart[i], (size_t)(h->nsec[i] + 1) * 4);
		memcpy(&buf[o_log[i]],   h->log_of[i],   (size_t)h->dim[i] * 4);
	}
	memcpy(&buf[o_grid], h->grid_off, (size_t)h->ngrid * 8);
	if (h->nblk > 0) { memcpy(&buf[o_bg], h->blk_grid, (size_t)h->nblk * 8); }
	memcpy(&buf[o_bo], h->blk_off, (size_t)(h->nblk + 1) * 8);
	if (upload(buf.data(), buf.size(), &L->arena) < 0) { delete L; return -1; }
	/* the staging vector dies at return: make sure the copy has been consumed */
	CTBD_CUDA(cudaStreamSynchronize(rt().stream));
	unsigned char* base = (unsigned char*)L->arena;
	LayoutDev& d = L->d;
	d.ndim = h->ndim; d.dtype = h->dtype; d.ngrid = h->ngrid; d.nblk = h->nblk; d.nstore = h->nstore;
	for (int i = 0; i < h->ndim; i++) {
		d.dim[i] = h->dim[i]; d.nsec[i] = h->nsec[i];
		d.sec_of[i]   = (const int32_t*)(base + o_sec[i]);
		d.pos_of[i]   = (const int32_t*)(base + o_pos[i]);
		d.secstart[i] = (const int32_t*)(base + o_start[i]);
		d.log_of[i]   = (const int32_t*)(base + o_log[i]);
	}
	d.grid_off = (const int64_t*)(base + o_grid);
	d.blk_grid = (const int64_t*)(base + o_bg);
	d.blk_off  = (const int64_t*)(base + o_bo);
	*layout = L;
	return 0;
}

int ctbd_layout_destroy(void* layout)
{
	Layout* L = (Layout*)layout;
	if (L == nullptr) { return 0; }
	ctbd_free(L->arena);
	delete L;
	return 0;
}

int ctbd_copy_plan_create(int dtype, int n, const struct ctbd_copy2d* descs_host, void** plan)
{
	CTBD_REQUIRE_INIT();
	if (dtype != CTBD_F64 && dtype != CTBD_C128) { return fail_msg("copy plan: unsupported dtype"); }
	CopyPlan* p = new CopyPlan();
	p->dtype = dtype; p->ndesc = n; p->descs = nullptr; p->tiles = nullptr;
	std::vector<CopyTile> tiles;
	for (int i = 0; i < n; i++) {
		if (descs_host[i].cols <= 0) { continue; }
		for (int r0 = 0; r0 < descs_host[i].rows; r0 += COPY_ROWS) { CopyTile t; t.desc = i; t.row0 = r0; tiles.push_back(t); }
	}
	p->ntiles = (int)tiles.size();
	int rc = upload(descs_host, (size_t)n * sizeof(ctbd_copy2d), (void**)&p->descs);
	rc |= upload(tiles.data(), tiles.size() * sizeof(CopyTile), (void**)&p->tiles);
	if (rc < 0) { ctbd_copy_plan_destroy(p); return -1; }
	*plan = p;
	return 0;
}

int ctbd_copy_plan_run(void* plan, const void* src, void* dst)
{
	CopyPlan* p = (CopyPlan*)plan;
	if (p == nullptr || p->ntiles == 0) { return 0; }
	const int grid = std::min(p->ntiles, rt().sm_count * 16);
	if (p->dtype == CTBD_F64) { copy2d_kernel<double><<<grid, 256, 0, rt().stream>>>(p->ntiles, p->tiles, p->descs, (const double*)src, (double*)dst); }
	else                      { copy2d_kernel<double2><<<grid, 256, 0, rt().stream>>>(p->ntiles, p->tiles, p->descs, (const double2*)src, (double2*)dst); }
	CTBD_LAUNCH_CHECK();
	return 0;
}

int ctbd_copy_plan_run_multi(void* plan, int nsrc, const void* const* srcs, int64_t src_stride, void* dst)
{
	CopyPlan* p = (CopyPlan*)plan;
	if (p == nullptr || p->ntiles == 0) { return 0; }
	if (nsrc < 1 || nsrc > 8 || src_stride <= 0) { return fail_msg("copy plan: between 1 and 8 sources"); }
	CopySrcs cs;
	for (int i = 0; i < 8; i++) { cs.p[i] = srcs[i < nsrc ? i : 0]; }
	cs.stride = src_stride;
	bool aligned = (((uintptr_t)dst) & 15) == 0;
	for (int i = 0; i < nsrc; i++) { aligned = aligned && ((((uintptr_t)srcs[i]) & 15) == 0); }
	cs.vec_ok = aligned ? 1 : 0;
	const int grid = std::min(p->ntiles, rt().sm_count * 8);
	if (p->dtype == CTBD_F64) { copy2d_multi_kernel<double><<<grid, 256, 0, rt().stream>>>(p->ntiles, p->tiles, p->descs, cs, (double*)dst); }
	else                      { copy2d_multi_kernel<double2><<<grid, 256, 0, rt().stream>>>(p->ntiles, p->tiles, p->descs, cs, (double2*)dst); }
	CTBD_LAUNCH_CHECK();
	return 0;
}

int ctbd_copy_plan_run_push(void* plan, const void* src, int ndst, void* const* dsts)
{
	CopyPlan* p = (CopyPlan*)plan;
	if (p == nullptr || p->ntiles == 0) { return 0; }
	if (ndst < 1 || ndst > 8) { return fail_msg("copy plan: between 1 and 8 destinations"); }
	CopyDsts cd;
	bool aligned = (((uintptr_t)src) & 15) == 0;
	for (int i = 0; i < 8; i++) { cd.p[i] = dsts[i < ndst ? i : 0]; }
	for (int i = 0; i < ndst; i++) { aligned = aligned && ((((uintptr_t)dsts[i]) & 15) == 0); }
	cd.n = ndst; cd.vec_ok = aligned ? 1 : 0;
	const int grid = std::min(p->ntiles, rt().sm_count * 8);
	if (p->dtype == CTBD_F64) { copy2d_push_kernel<double><<<grid, 256, 0, rt().stream>>>(p->ntiles, p->tiles, p->descs, (const double*)src, cd); }
	else                      { copy2d_push_kernel<double2><<<grid, 256, 0, rt().stream>>>(p->ntiles, p->tiles, p->descs, (const double2*)src, cd); }
	CTBD_LAUNCH_CHECK();
	return 0;
}

int ctbd_copy_plan_destroy(void* plan)
{
	CopyPlan* p = (CopyPlan*)plan;
	if (p == nullptr) { return 0; }
	ctbd_free(p->descs); ctbd_free(p->tiles);
	delete p;
	return 0;
}

int ctbd_remap(const struct ctbd_remap_args* a)
{
	CTBD_REQUIRE_INIT();
	const Layout* D = (const Layout*)a->dst_layout;
	const Layout* S = (const Layout*)a->src_layout;
	if (D->d.dtype != S->d.dtype) { return fail_msg("remap: dtype mismatch"); }
	if (D->d.nstore == 0) { return 0; }
	if (a->op == CTBD_REMAP_UNSLICE)
	{
		if (S->d.nstore == 0) { return 0; }
		RemapParams p;
		memset(&p, 0, sizeof(p));
		p.op = a->op; p.i_ax = a->i_ax; p.scale_ax = -1;
		void* ind_dev = nullptr;
		if (upload(a->ind, (size_t)S->d.dim[a->i_ax] * sizeof(int64_t), &ind_dev) < 0) { return -1; }
		p.ind = (const int64_t*)ind_dev;
		int64_t blocks = std::min<int64_t>(ceil_div(S->d.nstore, 256), (int64_t)rt().sm_count * 16);
		if (D->d.dtype == CTBD_F64) { unslice_kernel<double><<<(int)blocks, 256, 0, rt().stream>>>(D->d, S->d, p, (double*)a->dst, (const double*)a->src); }
		else if (D->d.dtype == CTBD_C128) { unslice_kernel<double2><<<(int)blocks, 256, 0, rt().stream>>>(D->d, S->d, p, (double2*)a->dst, (const double2*)a->src); }
		else { return fail_msg("remap: unsupported dtype"); }
		CTBD_LAUNCH_CHECK();
		ctbd_free(ind_dev);
		return 0;
	}
	RemapParams p;
	memset(&p, 0, sizeof(p));
	p.op = a->op; p.i_ax = a->i_ax; p.conj = a->conj; p.scale_ax = a->scale_ax; p.scale = a->scale;
	for (int i = 0; i < CTBD_MAXDIM; i++) { p.perm[i] = a->perm[i]; }
	void* ind_dev = nullptr;
	if (a->op == CTBD_REMAP_SLICE) {
		if (upload(a->ind, (size_t)D->d.dim[a->i_ax] * sizeof(int64_t), &ind_dev) < 0) { return -1; }
		p.ind = (const int64_t*)ind_dev;
	}
	const int threads = 256;
	int64_t blocks = ceil_div(D->d.nstore, threads);
	const int64_t maxb = (int64_t)rt().sm_count * 16;
	if (blocks > maxb) { blocks = maxb; }
	if (D->d.dtype == CTBD_F64) {
		remap_kernel<double><<<(int)blocks, threads, 0, rt().stream>>>(D->d, S->d, p, (double*)a->dst, (const double*)a->src);
	}
	else if (D->d.dtype == CTBD_C128) {
		remap_kernel<double2><<<(int)blocks, threads, 0, rt().stream>>>(D->d, S->d, p, (double2*)a->dst, (const double2*)a->src);
	}
	else { return fail_msg("remap: unsupported dtype"); }
	CTBD_LAUNCH_CHECK();
	if (ind_dev != nullptr) { ctbd_free(ind_dev); }
	return 0;
}

} // extern "C"
