/*
 * mps_io.c -- save_mps / load_mps in the reference's on-disk format, without libhdf5.
 *
 * Reference: src/state/mps.c:1219-1303 (save_mps) and :1309-1460 (load_mps), on top of src/util/hdf5_util.c.  The file is an HDF5
 * container with, on the root group, the attributes "nsites" (scalar int32), "qsite" (int32[d]) and "qbond_<i>" (int32[D_i],
 * i = 0 .. nsites) and the datasets "tensor_<i>" -- the DENSE site tensor [D_i, d, D_{i+1}], IEEE double or the compound {r, i} of
 * two doubles for complex numbers (mps.c:1232-1268, hdf5_util.c:327-351).
 *
 * libhdf5 is not available where this engine is built, so the container is written and parsed here, in the one dialect that
 * H5Fcreate(..., H5P_DEFAULT, H5P_DEFAULT) of libhdf5 1.8 - 1.14 produces and that all 86 fixture files of the reference's test-suite
 * use (the byte layouts below were taken from those fixtures; oracle/hdf5_v0.py is an independent reader of the same dialect):
 *   superblock version 0 with 8-byte offsets and lengths; the root group as a symbol table (version-1 B-tree node "TREE", local
 *   heap "HEAP", symbol-table node "SNOD"); version-1 object headers (with continuation blocks when reading); version-1 attribute,
 *   dataspace and datatype messages; version-2 fill-value and version-3 contiguous (or compact) layout messages.
 * Anything else (chunked or filtered datasets, new-style groups, superblock >= 2) is rejected with -1 and a message on stderr, like
 * every other failure of the two functions.
 *
 * Pure host code: no device call; the tensors are the reference's host structs.
 */
#ifndef _GNU_SOURCE
#define _GNU_SOURCE      /* qsort_r, pread */
#endif
#include <stdlib.h>
#include <errno.h>
#include <fcntl.h>
#include <unistd.h>
#include "ctb_internal.h"
#include "chemtensor_b200.h"

#define H5_UNDEF 0xFFFFFFFFFFFFFFFFull

/* ------------------------------------------------------------------------------------------------ */
/* little-endian byte buffer                                                                         */
/* ------------------------------------------------------------------------------------------------ */

struct bytebuf { unsigned char* p; size_t n, cap; };

static void bb_reserve(struct bytebuf* b, size_t extra)
{
	if (b->n + extra > b->cap) {
		while (b->n + extra > b->cap) { b->cap = b->cap ? 2 * b->cap : 4096; }
		b->p = realloc(b->p, b->cap);
	}
}
static void bb_put(struct bytebuf* b, const void* src, size_t n) { bb_reserve(b, n); memcpy(b->p + b->n, src, n); b->n += n; }
static void bb_zero(struct bytebuf* b, size_t n) { bb_reserve(b, n); memset(b->p + b->n, 0, n); b->n += n; }
static void bb_u8(struct bytebuf* b, unsigned v) { unsigned char c = (unsigned char)v; bb_put(b, &c, 1); }
static void bb_u16(struct bytebuf* b, unsigned v) { unsigned char c[2] = { (unsigned char)(v & 0xFF), (unsigned char)((v >> 8) & 0xFF) }; bb_put(b, c, 2); }
static void bb_u32(struct bytebuf* b, uint32_t v) { unsigned char c[4]; for (int i = 0; i < 4; i++) { c[i] = (unsigned char)((v >> (8 * i)) & 0xFF); } bb_put(b, c, 4); }
static void bb_u64(struct bytebuf* b, uint64_t v) { unsigned char c[8]; for (int i = 0; i < 8; i++) { c[i] = (unsigned char)((v >> (8 * i)) & 0xFF); } bb_put(b, c, 8); }
static void bb_pad8(struct bytebuf* b) { while (b->n & 7) { bb_u8(b, 0); } }
static void bb_set_u64(struct bytebuf* b, size_t at, uint64_t v) { for (int i = 0; i < 8; i++) { b->p[at + i] = (unsigned char)((v >> (8 * i)) & 0xFF); } }
static void bb_set_u32(struct bytebuf* b, size_t at, uint32_t v) { for (int i = 0; i < 4; i++) { b->p[at + i] = (unsigned char)((v >> (8 * i)) & 0xFF); } }
static void bb_set_u16(struct bytebuf* b, size_t at, unsigned v) { b->p[at] = (unsigned char)(v & 0xFF); b->p[at + 1] = (unsigned char)((v >> 8) & 0xFF); }

static size_t pad8(size_t n) { return (n + 7) & ~(size_t)7; }

/* ------------------------------------------------------------------------------------------------ */
/* message encoders                                                                                  */
/* ------------------------------------------------------------------------------------------------ */

/* datatype message bodies (version 1) */
static void enc_dtype_i32(struct bytebuf* b)
{
	bb_u8(b, 0x10); bb_u8(b, 0x08); bb_u8(b, 0); bb_u8(b, 0);      /* class 0 (fixed point) version 1; little endian, signed */
	bb_u32(b, 4);
	bb_u16(b, 0); bb_u16(b, 32);                                 /* bit offset, precision */
}
static void enc_dtype_f64(struct bytebuf* b)
{
	bb_u8(b, 0x11); bb_u8(b, 0x20); bb_u8(b, 0x3F); bb_u8(b, 0);   /* class 1 (floating point) version 1; little endian, implied mantissa bit, sign at bit 63 */
	bb_u32(b, 8);
	bb_u16(b, 0); bb_u16(b, 64);                                 /* bit offset, precision */
	bb_u8(b, 52); bb_u8(b, 11); bb_u8(b, 0); bb_u8(b, 52);          /* exponent location / size, mantissa location / size */
	bb_u32(b, 1023);                                             /* exponent bias */
}
/* compound { double r; double i; } (hdf5_util.c:327-351), version-1 member encoding */
static void enc_dtype_c128(struct bytebuf* b)
{
	bb_u8(b, 0x16); bb_u8(b, 2); bb_u8(b, 0); bb_u8(b, 0);         /* class 6 (compound) version 1; two members */
	bb_u32(b, 16);
	const char* names[2] = { "r", "i" };
	for (int m = 0; m < 2; m++) {
		bb_put(b, names[m], 2); bb_zero(b, 6);                     /* name, padded to a multiple of 8 */
		bb_u32(b, (uint32_t)(8 * m));                              /* byte offset of the member */
		bb_u8(b, 0); bb_zero(b, 3);                                /* dimensionality 0, reserved */
		bb_u32(b, 0); bb_u32(b, 0);                                /* dimension permutation, reserved */
		bb_zero(b, 16);                                           /* four dimension sizes */
		enc_dtype_f64(b);
	}
}
/* dataspace message body (version 1): scalar for rank 0, else simple with the maximum dimensions equal to the dimensions */
static void enc_dataspace(struct bytebuf* b, int rank, const uint64_t* dims)
{
	bb_u8(b, 1); bb_u8(b, (unsigned)rank); bb_u8(b, rank > 0 ? 1 : 0); bb_u8(b, 0); bb_u32(b, 0);
	for (int i = 0; i < rank; i++) { bb_u64(b, dims[i]); }
	for (int i = 0; i < rank; i++) { bb_u64(b, dims[i]); }
}
/* header of an object-header message; returns the position of the size field for patching */
static size_t msg_begin(struct bytebuf* b, unsigned type)
{
	bb_u16(b, type);
	const size_t at = b->n;
	bb_u16(b, 0); bb_u8(b, 0); bb_zero(b, 3);
	return at;
}
static void msg_end(struct bytebuf* b, size_t size_at)
{
	bb_pad8(b);
	bb_set_u16(b, size_at, (unsigned)(b->n - (size_at + 6)));
}
/* attribute message (version 1) with int32 data: rank 0 (scalar) or rank 1 */
static int enc_attribute_i32(struct bytebuf* b, const char* name, int rank, uint64_t count, const int32_t* data)
{
	const size_t name_size = strlen(name) + 1;
	const size_t body = 8 + pad8(name_size) + 16 + pad8(8 + (size_t)rank * 16) + pad8((size_t)count * 4);
	if (body > 65528) {
		fprintf(stderr, "chemtensor_b200: attribute '%s' with %llu entries does not fit an object-header message (the reference's format has the same limit)\n", name, (unsigned long long)count);
		return -1;
	}
	const size_t at = msg_begin(b, 0x000C);
	bb_u8(b, 1); bb_u8(b, 0);
	bb_u16(b, (unsigned)name_size); bb_u16(b, 12); bb_u16(b, (unsigned)(8 + rank * 16));
	bb_put(b, name, name_size); bb_pad8(b);
	enc_dtype_i32(b); bb_pad8(b);
	enc_dataspace(b, rank, &count); bb_pad8(b);
	for (uint64_t i = 0; i < count; i++) { bb_u32(b, (uint32_t)data[i]); }
	msg_end(b, at);
	return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* dense <-> block-sparse on host structs (reference block_sparse_to_dense_tensor :563, dense_to_block_sparse_tensor_entries :641)  */
/* ------------------------------------------------------------------------------------------------ */

/* visit every entry of every stored block: position of the entry in the dense row-major tensor */
static void scatter_blocks(const struct block_sparse_tensor* t, void* dense, int to_dense)
{
	const size_t esize = ctb_sizeof_dtype(t->dtype);
	const int nd = t->ndim;
	ct_long ngrid = 1;
	for (int i = 0; i < nd; i++) { ngrid *= t->dim_blocks[i]; }
	ct_long stride[CTB_MAXDIM];
	{ ct_long st = 1; for (int i = nd - 1; i >= 0; i--) { stride[i] = st; st *= t->dim_logical[i]; } }
	/* logical indices of every sector, per axis */
	ct_long** idx_of = ctb_malloc((size_t)(nd > 0 ? nd : 1) * sizeof(ct_long*));
	ct_long** start = ctb_malloc((size_t)(nd > 0 ? nd : 1) * sizeof(ct_long*));
	for (int i = 0; i < nd; i++) {
		idx_of[i] = ctb_malloc((size_t)t->dim_logical[i] * sizeof(ct_long));
		start[i] = ctb_calloc((size_t)t->dim_blocks[i] + 1, sizeof(ct_long));
		ct_long n = 0;
		for (ct_long s = 0; s < t->dim_blocks[i]; s++) {
			start[i][s] = n;
			for (ct_long j = 0; j < t->dim_logical[i]; j++) { if (t->qnums_logical[i][j] == t->qnums_blocks[i][s]) { idx_of[i][n++] = j; } }
		}
		start[i][t->dim_blocks[i]] = n;
	}
	for (ct_long cell = 0; cell < ngrid; cell++)
	{
		struct dense_tensor* blk = t->blocks[cell];
		if (blk == NULL) { continue; }
		int sec[CTB_MAXDIM];
		{ ct_long c = cell; for (int i = nd - 1; i >= 0; i--) { sec[i] = (int)(c % t->dim_blocks[i]); c /= t->dim_blocks[i]; } }
		ct_long numel = 1;
		for (int i = 0; i < nd; i++) { numel *= blk->dim[i]; }
		ct_long pos[CTB_MAXDIM] = { 0 };
		for (ct_long e = 0; e < numel; e++)
		{
			ct_long off = 0;
			for (int i = 0; i < nd; i++) { off += idx_of[i][start[i][sec[i]] + pos[i]] * stride[i]; }
			if (to_dense) { memcpy((char*)dense + (size_t)off * esize, (const char*)blk->data + (size_t)e * esize, esize); }
			else          { memcpy((char*)blk->data + (size_t)e * esize, (const char*)dense + (size_t)off * esize, esize); }
			for (int i = nd - 1; i >= 0; i--) { if (++pos[i] < blk->dim[i]) { break; } pos[i] = 0; }
		}
	}
	for (int i = 0; i < nd; i++) { ctb_free(idx_of[i]); ctb_free(start[i]); }
	ctb_free(idx_of); ctb_free(start);
}

/* ------------------------------------------------------------------------------------------------ */
/* save_mps                                                                                          */
/* ------------------------------------------------------------------------------------------------ */

static int cmp_name_idx(const void* a, const void* b, void* names)
{
	const char (*nm)[32] = names;
	return strcmp(nm[*(const int*)a], nm[*(const int*)b]);
}

int save_mps(const char* filename, const struct mps* mps)
{
	const int L = mps->nsites;
	if (L <= 0) { fprintf(stderr, "chemtensor_b200: save_mps: no sites\n"); return -1; }
	const int dtype = mps->a[0].dtype;
	if (dtype != CT_DOUBLE_REAL && dtype != CT_DOUBLE_COMPLEX) { fprintf(stderr, "chemtensor_b200: save_mps: only double precision tensors are stored\n"); return -1; }
	const size_t esize = ctb_sizeof_dtype(dtype);
	int rc = 0;

	/* ---- root object header: symbol-table message + attributes ---- */
	struct bytebuf root = { NULL, 0, 0 };
	bb_u8(&root, 1); bb_u8(&root, 0); bb_u16(&root, (unsigned)(1 + 2 + (L + 1))); bb_u32(&root, 1); bb_u32(&root, 0); bb_u32(&root, 0);   /* version, nmsg, ref count, size (patched), pad */
	const size_t symtab_at = msg_begin(&root, 0x0011);
	const size_t symtab_body = root.n;
	bb_u64(&root, 0); bb_u64(&root, 0);      /* B-tree and heap addresses, patched below */
	msg_end(&root, symtab_at);
	const int32_t nsites32 = L;
	rc |= enc_attribute_i32(&root, "nsites", 0, 1, &nsites32);
	rc |= enc_attribute_i32(&root, "qsite", 1, (uint64_t)mps->d, mps->qsite);
	for (int i = 0; i <= L && rc == 0; i++) {
		char nm[32];
		snprintf(nm, sizeof(nm), "qbond_%i", i);
		const struct block_sparse_tensor* t = (i < L) ? &mps->a[i] : &mps->a[L - 1];
		const int ax = (i < L) ? 0 : 2;
		rc |= enc_attribute_i32(&root, nm, 1, (uint64_t)t->dim_logical[ax], t->qnums_logical[ax]);
	}
	if (rc < 0) { free(root.p); return -1; }
	bb_set_u32(&root, 8, (uint32_t)(root.n - 16));

	/* ---- names, sorted as the symbol table wants them ---- */
	char (*names)[32] = ctb_malloc((size_t)L * 32);
	int* order = ctb_malloc((size_t)L * sizeof(int));
	for (int i = 0; i < L; i++) { snprintf(names[i], 32, "tensor_%i", i); order[i] = i; }
	qsort_r(order, (size_t)L, sizeof(int), cmp_name_idx, names);

	/* ---- local heap data segment: "" at offset 0, then the names (8-byte aligned) ---- */
	struct bytebuf heapd = { NULL, 0, 0 };
	bb_zero(&heapd, 8);
	uint64_t* name_off = ctb_malloc((size_t)L * sizeof(uint64_t));
	for (int k = 0; k < L; k++) {
		const int i = order[k];
		name_off[i] = heapd.n;
		bb_put(&heapd, names[i], strlen(names[i]) + 1); bb_pad8(&heapd);
	}

	/* ---- dataset object headers (all the same size) ---- */
	const int leaf_k = (L + 1) / 2 > 4 ? (L + 1) / 2 : 4;      /* one symbol-table node holds 2 K entries */
	const size_t addr_root = 96;
	const size_t addr_btree = addr_root + root.n;
	const size_t btree_size = 8 + 16 + (2 * 16 + 1) * 8 + 2 * 16 * 8;      /* internal node K = 16 */
	const size_t addr_heap = addr_btree + btree_size;
	const size_t addr_heapd = addr_heap + 32;
	const size_t addr_snod = addr_heapd + heapd.n;
	const size_t snod_size = 8 + (size_t)(2 * leaf_k) * 40;
	size_t addr = addr_snod + snod_size;
	struct bytebuf* dhdr = ctb_calloc((size_t)L, sizeof(struct bytebuf));
	uint64_t* addr_dhdr = ctb_malloc((size_t)L * sizeof(uint64_t));
	size_t* layout_addr_at = ctb_malloc((size_t)L * sizeof(size_t));
	uint64_t* nbytes = ctb_malloc((size_t)L * sizeof(uint64_t));
	for (int i = 0; i < L; i++)
	{
		const struct block_sparse_tensor* t = &mps->a[i];
		struct bytebuf* h = &dhdr[i];
		bb_u8(h, 1); bb_u8(h, 0); bb_u16(h, 4); bb_u32(h, 1); bb_u32(h, 0); bb_u32(h, 0);
		uint64_t dims[CTB_MAXDIM];
		uint64_t numel = 1;
		for (int k = 0; k < t->ndim; k++) { dims[k] = (uint64_t)t->dim_logical[k]; numel *= dims[k]; }
		nbytes[i] = numel * esize;
		size_t at = msg_begin(h, 0x0001); enc_dataspace(h, t->ndim, dims); msg_end(h, at);
		at = msg_begin(h, 0x0003); h->p[at + 2] = 1;      /* flags: constant message */
		if (dtype == CT_DOUBLE_REAL) { enc_dtype_f64(h); } else { enc_dtype_c128(h); }
		msg_end(h, at);
		at = msg_begin(h, 0x0005); h->p[at + 2] = 1; bb_u8(h, 2); bb_u8(h, 2); bb_u8(h, 2); bb_u8(h, 1); bb_u32(h, 0); msg_end(h, at);      /* fill value v2: late allocation, write if set, defined, size 0 */
		at = msg_begin(h, 0x0008); bb_u8(h, 3); bb_u8(h, 1); layout_addr_at[i] = h->n; bb_u64(h, 0); bb_u64(h, nbytes[i]); msg_end(h, at);      /* layout v3, contiguous */
		bb_set_u32(h, 8, (uint32_t)(h->n - 16));
		addr_dhdr[i] = addr;
		addr += h->n;
	}
	/* raw data behind all metadata */
	uint64_t* addr_data = ctb_malloc((size_t)L * sizeof(uint64_t));
	for (int i = 0; i < L; i++) { addr = pad8(addr); addr_data[i] = addr; addr += nbytes[i]; bb_set_u64(&dhdr[i], layout_addr_at[i], addr_data[i]); }
	const uint64_t addr_eof = addr;

	/* ---- assemble the metadata block ---- */
	struct bytebuf f = { NULL, 0, 0 };
	static const unsigned char sig[8] = { 0x89, 'H', 'D', 'F', '\r', '\n', 0x1A, '\n' };
	bb_put(&f, sig, 8);
	bb_u8(&f, 0); bb_u8(&f, 0); bb_u8(&f, 0); bb_u8(&f, 0); bb_u8(&f, 0);      /* superblock, free-space, root group versions, reserved, shared header version */
	bb_u8(&f, 8); bb_u8(&f, 8); bb_u8(&f, 0);                                  /* size of offsets, size of lengths, reserved */
	bb_u16(&f, (unsigned)leaf_k); bb_u16(&f, 16);                              /* group leaf node K, group internal node K */
	bb_u32(&f, 0);                                                           /* file consistency flags */
	bb_u64(&f, 0); bb_u64(&f, H5_UNDEF); bb_u64(&f, addr_eof); bb_u64(&f, H5_UNDEF);      /* base, free-space info, end of file, driver info */
	bb_u64(&f, 0); bb_u64(&f, addr_root); bb_u32(&f, 1); bb_u32(&f, 0); bb_u64(&f, addr_btree); bb_u64(&f, addr_heap);      /* root symbol-table entry, cache type 1 */
	CTB_REQUIRE(f.n == addr_root);
	bb_set_u64(&root, symtab_body, addr_btree); bb_set_u64(&root, symtab_body + 8, addr_heap);
	bb_put(&f, root.p, root.n);
	/* B-tree node: one child (the symbol-table node) */
	bb_put(&f, "TREE", 4); bb_u8(&f, 0); bb_u8(&f, 0); bb_u16(&f, 1); bb_u64(&f, H5_UNDEF); bb_u64(&f, H5_UNDEF);
	bb_u64(&f, 0); bb_u64(&f, addr_snod); bb_u64(&f, name_off[order[L - 1]]);      /* key 0 (""), child, key 1 (largest name of the child) */
	bb_zero(&f, addr_heap - f.n);
	/* local heap */
	bb_put(&f, "HEAP", 4); bb_u8(&f, 0); bb_zero(&f, 3); bb_u64(&f, heapd.n); bb_u64(&f, 1); bb_u64(&f, addr_heapd);      /* free list: none (H5HL_FREE_NULL) */
	bb_put(&f, heapd.p, heapd.n);
	/* symbol-table node */
	bb_put(&f, "SNOD", 4); bb_u8(&f, 1); bb_u8(&f, 0); bb_u16(&f, (unsigned)L);
	for (int k = 0; k < L; k++) { const int i = order[k]; bb_u64(&f, name_off[i]); bb_u64(&f, addr_dhdr[i]); bb_u32(&f, 0); bb_u32(&f, 0); bb_zero(&f, 16); }
	bb_zero(&f, addr_snod + snod_size - f.n);
	for (int i = 0; i < L; i++) { CTB_REQUIRE(f.n == addr_dhdr[i]); bb_put(&f, dhdr[i].p, dhdr[i].n); }

	FILE* fp = fopen(filename, "wb");
	if (fp == NULL) { fprintf(stderr, "chemtensor_b200: save_mps: cannot create '%s': %s\n", filename, strerror(errno)); rc = -1; }
	if (rc == 0 && fwrite(f.p, 1, f.n, fp) != f.n) { rc = -1; }
	uint64_t at = f.n;
	for (int i = 0; i < L && rc == 0; i++)
	{
		static const unsigned char zeros[8] = { 0 };
		if (addr_data[i] > at) { if (fwrite(zeros, 1, (size_t)(addr_data[i] - at), fp) != (size_t)(addr_data[i] - at)) { rc = -1; break; } at = addr_data[i]; }
		void* dense = ctb_calloc(nbytes[i] > 0 ? (size_t)nbytes[i] : 1, 1);
		scatter_blocks(&mps->a[i], dense, 1);
		if (fwrite(dense, 1, (size_t)nbytes[i], fp) != (size_t)nbytes[i]) { rc = -1; }
		at += nbytes[i];
		ctb_free(dense);
	}
	if (fp != NULL && fclose(fp) != 0) { rc = -1; }
	if (rc < 0 && fp != NULL) { fprintf(stderr, "chemtensor_b200: save_mps: writing '%s' failed\n", filename); }

	for (int i = 0; i < L; i++) { free(dhdr[i].p); }
	ctb_free(dhdr); ctb_free(addr_dhdr); ctb_free(layout_addr_at); ctb_free(nbytes); ctb_free(addr_data);
	ctb_free(name_off); ctb_free(order); ctb_free(names);
	free(heapd.p); free(root.p); free(f.p);
	return rc;
}

/* ------------------------------------------------------------------------------------------------ */
/* load_mps                                                                                          */
/* ------------------------------------------------------------------------------------------------ */

struct h5file { int fd; uint64_t size; };

static int h5_read(const struct h5file* f, uint64_t off, void* buf, size_t n)
{
	if (off > f->size || n > f->size - off) { return -1; }
	size_t done = 0;
	while (done < n) {
		const ssize_t r = pread(f->fd, (char*)buf + done, n - done, (off_t)(off + done));
		if (r <= 0) { return -1; }
		done += (size_t)r;
	}
	return 0;
}
static uint64_t le64(const unsigned char* p) { uint64_t v = 0; for (int i = 7; i >= 0; i--) { v = (v << 8) | p[i]; } return v; }
static uint32_t le32(const unsigned char* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
static unsigned le16(const unsigned char* p) { return (unsigned)p[0] | ((unsigned)p[1] << 8); }

struct h5msg { unsigned type; uint64_t off; unsigned size; };

/* all messages of a version-1 object header, continuation blocks included */
static int h5_object_header(const struct h5file* f, uint64_t addr, struct h5msg** out, int* nout)
{
	unsigned char pre[16];
	if (h5_read(f, addr, pre, 16) < 0 || pre[0] != 1) { return -1; }
	const int nmsg = (int)le16(pre + 2);
	struct h5msg* msgs = ctb_calloc((size_t)(nmsg > 0 ? nmsg : 1), sizeof(*msgs));
	uint64_t blk_off[64], blk_len[64];
	int nblk = 1, cur = 0, n = 0;
	blk_off[0] = addr + 16; blk_len[0] = le32(pre + 8);
	while (cur < nblk && n < nmsg)
	{
		uint64_t pos = blk_off[cur], end = blk_off[cur] + blk_len[cur];
		cur++;
		while (pos + 8 <= end && n < nmsg) {
			unsigned char mh[8];
			if (h5_read(f, pos, mh, 8) < 0) { ctb_free(msgs); return -1; }
			msgs[n].type = le16(mh); msgs[n].size = le16(mh + 2); msgs[n].off = pos + 8;
			if (msgs[n].type == 0x0010 && nblk < 64) {
				unsigned char c[16];
				if (h5_read(f, pos + 8, c, 16) < 0) { ctb_free(msgs); return -1; }
				blk_off[nblk] = le64(c); blk_len[nblk] = le64(c + 8); nblk++;
			}
			pos += 8 + msgs[n].size;
			n++;
		}
	}
	*out = msgs; *nout = n;
	return 0;
}

enum h5kind { H5K_I32, H5K_I64, H5K_F64, H5K_C128, H5K_OTHER };

/* kind of a datatype message body; *len = encoded length */
static enum h5kind h5_datatype(const unsigned char* p, size_t avail, size_t* len)
{
	if (avail < 8) { *len = 0; return H5K_OTHER; }
	const int cls = p[0] & 0x0F, ver = p[0] >> 4;
	const uint32_t size = le32(p + 4);
	if (cls == 0) { *len = 12; return (size == 4) ? H5K_I32 : (size == 8 ? H5K_I64 : H5K_OTHER); }
	if (cls == 1) { *len = 20; return (size == 8) ? H5K_F64 : H5K_OTHER; }
	if (cls == 6)
	{
		const int nmemb = p[1] | (p[2] << 8);
		size_t pos = 8;
		int ok = (nmemb == 2 && size == 16);
		for (int m = 0; m < nmemb; m++) {
			size_t e = pos;
			while (e < avail && p[e] != 0) { e++; }
			if (e >= avail) { *len = 0; return H5K_OTHER; }
			const size_t nlen = e - pos + 1;
			if (ok && !(nlen == 2 && p[pos] == (m == 0 ? 'r' : 'i'))) { ok = 0; }
			uint32_t moff = 0;
			if (ver == 1) { pos += pad8(nlen); if (pos + 32 > avail) { *len = 0; return H5K_OTHER; } moff = le32(p + pos); pos += 4 + 1 + 3 + 4 + 4 + 16; }
			else if (ver == 2) { pos += pad8(nlen); if (pos + 4 > avail) { *len = 0; return H5K_OTHER; } moff = le32(p + pos); pos += 4; }
			else { pos += nlen; const int nb = size < 256 ? 1 : (size < 65536 ? 2 : 4); if (pos + (size_t)nb > avail) { *len = 0; return H5K_OTHER; } for (int q = nb - 1; q >= 0; q--) { moff = (moff << 8) | p[pos + q]; } pos += (size_t)nb; }
			size_t ml = 0;
			const enum h5kind mk = h5_datatype(p + pos, avail - pos, &ml);
			if (ml == 0) { *len = 0; return H5K_OTHER; }
			if (mk != H5K_F64 || moff != (uint32_t)(8 * m)) { ok = 0; }
			pos += ml;
		}
		*len = pos;
		return ok ? H5K_C128 : H5K_OTHER;
	}
	*len = 0;
	return H5K_OTHER;
}

/* dataspace message body: rank and dimensions */
static int h5_dataspace(const unsigned char* p, size_t avail, int* rank, uint64_t* dims)
{
	if (avail < 4) { return -1; }
	const int ver = p[0];
	*rank = p[1];
	if (*rank > CTB_MAXDIM) { return -1; }
	const size_t pos = (ver == 1) ? 8 : 4;
	if (pos + (size_t)(*rank) * 8 > avail) { return -1; }
	for (int i = 0; i < *rank; i++) { dims[i] = le64(p + pos + 8 * (size_t)i); }
	return 0;
}

struct h5attr { char name[64]; int rank; uint64_t count; int32_t* data; };

static int h5_attribute(const struct h5file* f, const struct h5msg* m, struct h5attr* a)
{
	unsigned char* p = ctb_malloc(m->size > 0 ? m->size : 1);
	int rc = -1;
	a->data = NULL;
	if (h5_read(f, m->off, p, m->size) == 0 && m->size >= 8 && p[0] == 1)
	{
		const size_t name_size = le16(p + 2), dt_size = le16(p + 4), ds_size = le16(p + 6);
		size_t pos = 8;
		if (pos + pad8(name_size) + pad8(dt_size) + pad8(ds_size) <= m->size && name_size >= 1 && name_size <= sizeof(a->name))
		{
			memcpy(a->name, p + pos, name_size); a->name[name_size - 1] = 0;
			pos += pad8(name_size);
			size_t dl = 0;
			const enum h5kind k = h5_datatype(p + pos, dt_size, &dl);
			pos += pad8(dt_size);
			uint64_t dims[CTB_MAXDIM];
			if ((k == H5K_I32 || k == H5K_I64) && h5_dataspace(p + pos, ds_size, &a->rank, dims) == 0 && a->rank <= 1)
			{
				pos += pad8(ds_size);
				a->count = (a->rank == 0) ? 1 : dims[0];
				const size_t w = (k == H5K_I32) ? 4 : 8;
				if (pos + a->count * w <= m->size) {
					a->data = ctb_malloc((size_t)(a->count > 0 ? a->count : 1) * sizeof(int32_t));
					for (uint64_t i = 0; i < a->count; i++) { a->data[i] = (k == H5K_I32) ? (int32_t)le32(p + pos + 4 * i) : (int32_t)(int64_t)le64(p + pos + 8 * i); }
					rc = 0;
				}
			}
			else { rc = 1; }      /* an attribute of another kind: skipped */
		}
	}
	ctb_free(p);
	return rc;
}

/* object header address of the dataset 'name' in the symbol-table group (B-tree 'addr'), 0 if absent */
static uint64_t h5_find(const struct h5file* f, uint64_t addr, uint64_t heap_data, const char* name, int depth)
{
	unsigned char h[8];
	if (depth > 16 || h5_read(f, addr, h, 8) < 0) { return 0; }
	if (memcmp(h, "TREE", 4) == 0)
	{
		const int nent = (int)le16(h + 6);
		for (int i = 0; i < nent; i++) {
			unsigned char c[8];
			if (h5_read(f, addr + 8 + 16 + 8 + (uint64_t)i * 16, c, 8) < 0) { return 0; }
			const uint64_t r = h5_find(f, le64(c), heap_data, name, depth + 1);
			if (r != 0) { return r; }
		}
		return 0;
	}
	if (memcmp(h, "SNOD", 4) == 0)
	{
		const int nsym = (int)le16(h + 6);
		for (int i = 0; i < nsym; i++) {
			unsigned char e[16];
			char nm[64] = { 0 };
			if (h5_read(f, addr + 8 + (uint64_t)i * 40, e, 16) < 0) { return 0; }
			const uint64_t noff = heap_data + le64(e);
			const size_t nread = (f->size - noff < sizeof(nm) - 1) ? (size_t)(f->size - noff) : sizeof(nm) - 1;
			if (noff >= f->size || h5_read(f, noff, nm, nread) < 0) { return 0; }
			if (strcmp(nm, name) == 0) { return le64(e + 8); }
		}
	}
	return 0;
}

int load_mps(const char* filename, struct mps* mps)
{
	struct h5file f;
	f.fd = open(filename, O_RDONLY);
	if (f.fd < 0) { fprintf(stderr, "chemtensor_b200: load_mps: cannot open '%s': %s\n", filename, strerror(errno)); return -1; }
	f.size = (uint64_t)lseek(f.fd, 0, SEEK_END);
	int rc = -1;
	struct h5msg* msgs = NULL; int nmsg = 0;
	struct h5attr* attrs = NULL; int nattr = 0;
	qnumber** qbonds = NULL; ct_long* dim_bonds = NULL;
	int nsites = 0;
	const char* why = "not an HDF5 file of the supported dialect (superblock version 0, 8-byte offsets)";
	unsigned char sb[96];
	static const unsigned char sig[8] = { 0x89, 'H', 'D', 'F', '\r', '\n', 0x1A, '\n' };
	if (h5_read(&f, 0, sb, 96) < 0 || memcmp(sb, sig, 8) != 0 || sb[8] != 0 || sb[13] != 8 || sb[14] != 8) { goto done; }
	const uint64_t root_ohdr = le64(sb + 56 + 8);
	uint64_t btree = le64(sb + 56 + 24), heap = le64(sb + 56 + 32);
	const int cached = (le32(sb + 56 + 16) == 1);
	why = "unreadable root object header";
	if (h5_object_header(&f, root_ohdr, &msgs, &nmsg) < 0) { goto done; }
	attrs = ctb_calloc((size_t)(nmsg > 0 ? nmsg : 1), sizeof(*attrs));
	int have_symtab = cached;
	for (int i = 0; i < nmsg; i++) {
		if (msgs[i].type == 0x0011) {
			unsigned char c[16];
			if (h5_read(&f, msgs[i].off, c, 16) == 0) { btree = le64(c); heap = le64(c + 8); have_symtab = 1; }
		}
		else if (msgs[i].type == 0x000C) {
			if (h5_attribute(&f, &msgs[i], &attrs[nattr]) == 0) { nattr++; }
		}
	}
	why = "the root group is not a symbol-table group";
	if (!have_symtab) { goto done; }
	unsigned char hh[32];
	why = "unreadable local heap";
	if (h5_read(&f, heap, hh, 32) < 0 || memcmp(hh, "HEAP", 4) != 0) { goto done; }
	const uint64_t heap_data = le64(hh + 24);

	/* attributes: nsites, qsite, qbond_<i> (reference mps.c:1318-1385) */
	const struct h5attr* a_nsites = NULL; const struct h5attr* a_qsite = NULL;
	for (int i = 0; i < nattr; i++) {
		if (strcmp(attrs[i].name, "nsites") == 0) { a_nsites = &attrs[i]; }
		if (strcmp(attrs[i].name, "qsite") == 0)  { a_qsite = &attrs[i]; }
	}
	why = "attribute 'nsites' or 'qsite' missing or invalid";
	if (a_nsites == NULL || a_nsites->count != 1 || a_nsites->data[0] <= 0 || a_qsite == NULL || a_qsite->count == 0) { goto done; }
	nsites = a_nsites->data[0];
	const ct_long d = (ct_long)a_qsite->count;
	qbonds = ctb_calloc((size_t)nsites + 1, sizeof(qnumber*));
	dim_bonds = ctb_calloc((size_t)nsites + 1, sizeof(ct_long));
	why = "a 'qbond_<i>' attribute is missing";
	for (int i = 0; i <= nsites; i++) {
		char nm[32];
		snprintf(nm, sizeof(nm), "qbond_%i", i);
		const struct h5attr* q = NULL;
		for (int k = 0; k < nattr; k++) { if (strcmp(attrs[k].name, nm) == 0) { q = &attrs[k]; break; } }
		if (q == NULL || q->count == 0) { goto done; }
		dim_bonds[i] = (ct_long)q->count;
		qbonds[i] = q->data;      /* borrowed */
	}

	/* datasets: type of tensor_0 decides the numeric type (mps.c:1387-1417) */
	int dtype = -1;
	for (int i = 0; i < nsites; i++)
	{
		char nm[32];
		snprintf(nm, sizeof(nm), "tensor_%i", i);
		why = "a 'tensor_<i>' dataset is missing";
		const uint64_t oh = h5_find(&f, btree, heap_data, nm, 0);
		if (oh == 0) { goto fail_tensors; }
		struct h5msg* dm = NULL; int ndm = 0;
		why = "unreadable dataset header";
		if (h5_object_header(&f, oh, &dm, &ndm) < 0) { goto fail_tensors; }
		enum h5kind kind = H5K_OTHER;
		int rank = -1; uint64_t dims[CTB_MAXDIM];
		uint64_t data_addr = H5_UNDEF, data_size = 0;
		int compact = 0;
		int bad = 0;
		for (int k = 0; k < ndm; k++)
		{
			unsigned char buf[512];
			const size_t n = dm[k].size < sizeof(buf) ? dm[k].size : sizeof(buf);
			if (dm[k].type != 0x0001 && dm[k].type != 0x0003 && dm[k].type != 0x0008) { continue; }
			if (h5_read(&f, dm[k].off, buf, n) < 0) { bad = 1; break; }
			if (dm[k].type == 0x0001) { if (h5_dataspace(buf, n, &rank, dims) < 0) { bad = 1; } }
			else if (dm[k].type == 0x0003) { size_t dl; kind = h5_datatype(buf, n, &dl); }
			else {
				if (buf[0] != 3) { bad = 1; }
				else if (buf[1] == 1) { data_addr = le64(buf + 2); data_size = le64(buf + 10); }
				else if (buf[1] == 0) { compact = 1; data_size = le16(buf + 2); data_addr = dm[k].off + 4; }
				else { bad = 1; }      /* chunked */
			}
		}
		ctb_free(dm);
		why = "a dataset uses a layout, type or shape outside the reference's MPS format";
		if (bad || rank != 3 || (kind != H5K_F64 && kind != H5K_C128)) { goto fail_tensors; }
		const int dt_i = (kind == H5K_F64) ? CT_DOUBLE_REAL : CT_DOUBLE_COMPLEX;
		if (i == 0) {
			dtype = dt_i;
			allocate_mps((enum numeric_type)dtype, nsites, d, a_qsite->data, dim_bonds, (const qnumber**)qbonds, mps);
		}
		why = "tensor dimensions or types do not match the bond attributes";
		if (dt_i != dtype || (ct_long)dims[0] != dim_bonds[i] || (ct_long)dims[1] != d || (ct_long)dims[2] != dim_bonds[i + 1]) { goto fail_tensors; }
		const uint64_t nb = dims[0] * dims[1] * dims[2] * ctb_sizeof_dtype(dtype);
		void* dense = ctb_calloc(nb > 0 ? (size_t)nb : 1, 1);
		why = "tensor data could not be read";
		if (data_addr != H5_UNDEF) {      /* an undefined address = never written = zeros */
			if (data_size < nb || h5_read(&f, data_addr, dense, (size_t)nb) < 0) { ctb_free(dense); goto fail_tensors; }
		}
		(void)compact;
		scatter_blocks(&mps->a[i], dense, 0);      /* entries outside the conserving blocks are dropped, as dense_to_block_sparse_tensor_entries does */
		ctb_free(dense);
	}
	rc = 0;
	goto done;
fail_tensors:
	if (dtype >= 0) { delete_mps(mps); }
done:
	if (rc < 0) { fprintf(stderr, "chemtensor_b200: load_mps: '%s': %s\n", filename, why); }
	for (int i = 0; i < nattr; i++) { ctb_free(attrs[i].data); }
	ctb_free(attrs); ctb_free(msgs); ctb_free(qbonds); ctb_free(dim_bonds);
	close(f.fd);
	return rc;
}
