// hb_api.cu -- the C ABI (include/harry_b200.h): context, device-resident mesh, and the
// host-buffer entry points that replace the reference's calls.
#include "hb_internal.cuh"
#include <mutex>
#include <unordered_map>
#include <utility>

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static std::string g_create_err;

int hb_fail(hb_ctx *ctx, int code, const char *fmt, ...)
{
	char buf[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof buf, fmt, ap);
	va_end(ap);
	if (ctx) ctx->err = buf;
	else g_create_err = buf;
	return code;
}

static const char *device_error_text(int code)
{
	switch (code) {
	case 1: return "face with more than 65535 edges or negative size";
	case 2: return "edge origin or twin out of range";
	case 3: return "vertex order entry out of range";
	case 4: return "face order entry out of range";
	case 5: return "fan walk does not terminate (inconsistent twin table)";
	case 6: return "mixed-type interpretation group";
	case 7: return "attribute binding out of range";
	case 8: return "emit_type stream shorter than the number of emissions";
	default: return "unknown device error";
	}
}

int hb_check_device_error(hb_ctx *ctx, const char *what)
{
	HB_CUDA(ctx, cudaMemcpyAsync(ctx->h_err, ctx->d_err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
	HB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	if (*ctx->h_err) {
		const int code = *ctx->h_err;
		*ctx->h_err = 0;
		cudaMemsetAsync(ctx->d_err, 0, sizeof(int), ctx->stream);
		return hb_fail(ctx, code == 6 ? HB_ERR_UNSUPPORTED : HB_ERR_INVALID, "%s: %s", what, device_error_text(code));
	}
	return 0;
}

extern "C" int hb_ctx_create(int device, hb_ctx **out)
{
	*out = nullptr;
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if (e != cudaSuccess || ndev == 0)
		return hb_fail(nullptr, HB_ERR_CUDA, "no CUDA device (%s); the attribute path has no CPU fallback", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
	if (device < 0 || device >= ndev) return hb_fail(nullptr, HB_ERR_INVALID, "device %d out of range (%d devices)", device, ndev);
	hb_ctx *ctx = new hb_ctx();
	ctx->device = device;
#define CREATE_TRY(call)                                                                                      \
	do {                                                                                                      \
		cudaError_t e2 = (call);                                                                              \
		if (e2 != cudaSuccess) {                                                                              \
			hb_fail(nullptr, HB_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e2));                    \
			delete ctx;                                                                                       \
			return HB_ERR_CUDA;                                                                               \
		}                                                                                                     \
	} while (0)
	CREATE_TRY(cudaSetDevice(device));
	CREATE_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
	CREATE_TRY(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
	CREATE_TRY(cudaEventCreateWithFlags(&ctx->ev_alloc, cudaEventDisableTiming));
	CREATE_TRY(cudaEventCreateWithFlags(&ctx->ev_up[0], cudaEventDisableTiming));
	CREATE_TRY(cudaEventCreateWithFlags(&ctx->ev_up[1], cudaEventDisableTiming));
	for (int i = 0; i < 6; ++i) CREATE_TRY(cudaEventCreate(&ctx->ev[i]));
	CREATE_TRY(cudaMalloc((void **)&ctx->d_err, sizeof(int)));
	CREATE_TRY(cudaMemset(ctx->d_err, 0, sizeof(int)));
	CREATE_TRY(cudaMallocHost((void **)&ctx->h_err, sizeof(int)));
	*ctx->h_err = 0;
	cudaDeviceProp prop;
	CREATE_TRY(cudaGetDeviceProperties(&prop, device));
	ctx->sm_count = prop.multiProcessorCount;
	// keep freed blocks in the stream-ordered pool instead of returning them to the driver
	cudaMemPool_t pool;
	if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
		uint64_t thresh = UINT64_MAX;
		cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh);
	}
#undef CREATE_TRY
	*out = ctx;
	return 0;
}

extern "C" void hb_ctx_destroy(hb_ctx *ctx)
{
	if (!ctx) return;
	cudaSetDevice(ctx->device);
	if (ctx->stream) cudaStreamSynchronize(ctx->stream);
	if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
	if (ctx->ev_alloc) cudaEventDestroy(ctx->ev_alloc);
	for (int i = 0; i < 2; ++i)
		if (ctx->ev_up[i]) cudaEventDestroy(ctx->ev_up[i]);
	for (int i = 0; i < 6; ++i)
		if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
	for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
	if (ctx->d_err) cudaFree(ctx->d_err);
	if (ctx->h_err) cudaFreeHost(ctx->h_err);
	if (ctx->stream) cudaStreamDestroy(ctx->stream);
	delete ctx;
}

extern "C" const char *hb_last_error(hb_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

extern "C" void hb_last_timing(hb_ctx *ctx, float *kernel_ms, float *copy_ms)
{
	if (kernel_ms) *kernel_ms = ctx->kernel_ms;
	if (copy_ms) *copy_ms = ctx->copy_ms;
}

extern "C" uint64_t hb_kernel_launches(hb_ctx *ctx) { return ctx->launches; }

cudaEvent_t hb_prof_event(hb_ctx *ctx)
{
	cudaEvent_t e = nullptr;
	cudaEventCreate(&e);
	ctx->ev_pool.push_back(e);
	return e;
}

// Per-kernel timing: while enabled, every launch is bracketed by CUDA events on the context
// stream.  hb_ctx_profile_report() synchronizes, writes one line per kernel name
// ("name launches total_ms\n") into buf and clears the records.
extern "C" int hb_ctx_profile(hb_ctx *ctx, int enable)
{
	ctx->profiling = enable != 0;
	return 0;
}
extern "C" int hb_ctx_profile_report(hb_ctx *ctx, char *buf, size_t len)
{
	HB_CUDA(ctx, cudaSetDevice(ctx->device));
	HB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	struct Acc { const char *name; int n; double ms; };
	std::vector<Acc> acc;
	for (auto &r : ctx->prof) {
		float ms = 0.f;
		cudaEventElapsedTime(&ms, r.a, r.b);
		bool found = false;
		for (auto &a : acc)
			if (strcmp(a.name, r.name) == 0) { a.n++; a.ms += ms; found = true; break; }
		if (!found) acc.push_back(Acc{ r.name, 1, ms });
	}
	size_t pos = 0;
	if (len) buf[0] = 0;
	for (auto &a : acc) {
		int w = snprintf(buf + pos, pos < len ? len - pos : 0, "%s %d %.6f\n", a.name, a.n, a.ms);
		if (w < 0 || pos + (size_t)w >= len) break;
		pos += (size_t)w;
	}
	for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
	ctx->ev_pool.clear();
	ctx->prof.clear();
	return 0;
}
// events on the context stream for timing a region of device-resident stages
extern "C" int hb_ctx_mark(hb_ctx *ctx, int idx)
{
	if (idx < 0 || idx >= 6) return hb_fail(ctx, HB_ERR_INVALID, "mark index out of range");
	HB_CUDA(ctx, cudaSetDevice(ctx->device));
	HB_CUDA(ctx, cudaEventRecord(ctx->ev[idx], ctx->stream));
	return 0;
}
extern "C" int hb_ctx_elapsed(hb_ctx *ctx, int a, int b, float *ms)
{
	if (a < 0 || a >= 6 || b < 0 || b >= 6) return hb_fail(ctx, HB_ERR_INVALID, "mark index out of range");
	HB_CUDA(ctx, cudaEventSynchronize(ctx->ev[b]));
	HB_CUDA(ctx, cudaEventElapsedTime(ms, ctx->ev[a], ctx->ev[b]));
	return 0;
}

extern "C" int hb_ctx_sync(hb_ctx *ctx)
{
	HB_CUDA(ctx, cudaSetDevice(ctx->device));
	HB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return hb_check_device_error(ctx, "sync");
}

// ------------------------------------------------------------------------------------------------
int hb_dalloc(hb_dmesh *m, void **p, size_t bytes)
{
	if (*p) return 0; // already allocated by an earlier run over the same mesh (sizes are per-mesh constants)
	if (bytes == 0) bytes = 16;
	cudaError_t e = cudaMallocAsync(p, bytes, m->ctx->stream);
	if (e != cudaSuccess) {
		*p = nullptr;
		return hb_fail(m->ctx, e == cudaErrorMemoryAllocation ? HB_ERR_NOMEM : HB_ERR_CUDA, "cudaMallocAsync(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
	}
	m->allocs.push_back(*p);
	return 0;
}

static int storage_type(int type, int q)
{
	if (q == 0) return type;
	if (q <= 8) return HB_UCHAR;
	if (q <= 16) return HB_USHORT;
	if (q <= 32) return HB_UINT;
	return HB_ULONG;
}

void hb_fill_list_params(ListParams &p, const hb_list_desc &L)
{
	uint8_t *rows = p.rows;
	memset(&p, 0, sizeof p);
	p.rows = rows;
	p.nrows = L.nrows;
	p.stride = L.stride;
	p.ncomp = L.ncomp;
	p.target = L.target;
	uint32_t pos = 0;
	int uni = -2;
	for (int j = 0; j < L.ncomp; ++j) {
		p.type[j] = L.type[j];
		p.quant[j] = L.quant[j];
		p.stype[j] = (uint8_t)storage_type(L.type[j], L.quant[j]);
		p.size[j] = (uint8_t)hb_type_size(p.stype[j]);
		p.offset[j] = L.offset[j];
		p.sym_off[j] = (uint16_t)pos;
		pos += p.size[j];
		uni = (uni == -2 || uni == p.stype[j]) ? p.stype[j] : -1;
	}
	p.sym_stride = pos;
	p.uniform_stype = uni < 0 ? -1 : uni;
}

static int validate_list(hb_ctx *ctx, const hb_list_desc &L, bool for_coder)
{
	if (L.ncomp > HB_MAX_COMP) return hb_fail(ctx, HB_ERR_INVALID, "list: %d components (max %d)", L.ncomp, HB_MAX_COMP);
	if (L.nrows && L.ncomp && !L.rows) return hb_fail(ctx, HB_ERR_INVALID, "list: rows == NULL");
	for (int j = 0; j < L.ncomp; ++j) {
		if (L.type[j] >= HB_TYPE_NONE) return hb_fail(ctx, HB_ERR_INVALID, "list: bad type %d", L.type[j]);
		const int sz = hb_type_size(L.type[j]);
		if (L.quant[j] > 8 * sz) return hb_fail(ctx, HB_ERR_INVALID, "list: %d-bit quantization in a %d-byte slot", L.quant[j], sz);
		if ((uint32_t)L.offset[j] + sz > L.stride) return hb_fail(ctx, HB_ERR_INVALID, "list: component %d outside the row", j);
		if (for_coder) {
			const int st = storage_type(L.type[j], L.quant[j]);
			if (st == HB_DOUBLE) return hb_fail(ctx, HB_ERR_UNSUPPORTED, "double lists: the reference reads masks[8] out of bounds (prediction.h:33-44)");
			if (L.quant[j] > 31 && L.quant[j] != 8 * hb_type_size(st)) return hb_fail(ctx, HB_ERR_UNSUPPORTED, "quantization with more than 31 bits (undefined in the reference)");
		}
	}
	return 0;
}

static int upload(hb_dmesh *m, void **dst, const void *src, size_t bytes)
{
	HB_TRY(hb_dalloc(m, dst, bytes));
	if (!(bytes && src)) return 0;
	hb_ctx *ctx = m->ctx;
	if (m->async_copy) {
		// the buffer comes from the stream-ordered pool of ctx->stream: the copy stream may touch it
		// once it has waited for an event recorded behind the allocation
		HB_CUDA(ctx, cudaEventRecord(ctx->ev_alloc, ctx->stream));
		HB_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_alloc, 0));
		HB_CUDA(ctx, cudaMemcpyAsync(*dst, src, bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
	} else {
		HB_CUDA(ctx, cudaMemcpyAsync(*dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
	}
	return 0;
}

static int add_list(hb_dmesh *m, const hb_list_desc &L, bool for_coder)
{
	HB_TRY(validate_list(m->ctx, L, for_coder));
	DevList dl;
	memset(&dl.p, 0, sizeof dl.p);
	hb_fill_list_params(dl.p, L);
	void *rows = nullptr;
	HB_TRY(upload(m, &rows, L.ncomp ? L.rows : nullptr, (size_t)L.nrows * L.stride));
	dl.p.rows = (uint8_t *)rows;
	void *b = nullptr;
	HB_TRY(hb_dalloc(m, &b, 3 * (size_t)L.stride + 16));
	dl.d_bounds = (uint8_t *)b;
	HB_CUDA(m->ctx, cudaMemsetAsync(b, 0, 3 * (size_t)L.stride + 16, m->ctx->stream));
	m->lists.push_back(dl);
	return 0;
}

extern "C" void hb_dmesh_free(hb_dmesh *m)
{
	if (!m) return;
	cudaSetDevice(m->ctx->device);
	if (m->async_copy) cudaStreamSynchronize(m->ctx->copy_stream); // nothing may still be landing in these buffers
	for (void *p : m->allocs) cudaFreeAsync(p, m->ctx->stream);
	cudaStreamSynchronize(m->ctx->stream);
	delete m;
}

static int build_slot_table(hb_ctx *ctx, const int32_t *off, const uint16_t *lists, int nregs, int nlists, std::vector<int16_t> &slot, std::vector<int> *counts)
{
	slot.assign((size_t)(nregs > 0 ? nregs : 1) * (nlists > 0 ? nlists : 1), -1);
	if (counts) counts->assign(nregs > 0 ? nregs : 1, 0);
	for (int r = 0; r < nregs; ++r) {
		const int nb = off[r + 1] - off[r];
		if (nb < 0) return hb_fail(ctx, HB_ERR_INVALID, "region table: negative binding count");
		if (counts) (*counts)[r] = nb;
		for (int a = 0; a < nb; ++a) {
			const int l = lists[off[r] + a];
			if (l >= nlists) return hb_fail(ctx, HB_ERR_INVALID, "region table: list %d out of range", l);
			if (slot[(size_t)r * nlists + l] >= 0) return hb_fail(ctx, HB_ERR_UNSUPPORTED, "list %d bound twice in one region", l);
			slot[(size_t)r * nlists + l] = (int16_t)a;
		}
	}
	return 0;
}

// vertex_only: the caller will only reconstruct vertex lists (hb_attr_decode of a mesh whose face and
// corner lists carry no components): the face-side arrays (order_f, face regions, face / corner bindings --
// 280 MB on the 10M-vertex mesh) are then neither uploaded nor ranked
static int dmesh_upload_impl(hb_ctx *ctx, const hb_mesh_desc *d, hb_dmesh *m, bool vertex_only = false)
{
	m->ctx = ctx;
	m->nv = d->nv; m->nf = d->nf; m->ne = d->ne;
	m->norder = d->norder;
	m->has_order_f = d->order_f != nullptr && !vertex_only;
	m->norder_f = vertex_only ? 0 : (d->order_f ? d->norder_f : d->nf);
	m->nb_face = d->nb_face; m->nb_vtx = d->nb_vtx; m->nb_corner = d->nb_corner;
	m->nregs_face = d->nregs_face; m->nregs_vtx = d->nregs_vtx; m->nlists = d->nlists;
	if (d->ne && (!d->edges || !d->face_off)) return hb_fail(ctx, HB_ERR_INVALID, "mesh: edges / face_off == NULL");
	if (d->nf && d->face_off[d->nf] != d->ne) return hb_fail(ctx, HB_ERR_INVALID, "mesh: face_off[nf] != ne");
	if (d->norder && !d->order) return hb_fail(ctx, HB_ERR_INVALID, "mesh: order == NULL");
	if ((d->nv && !d->vtx_regs) || (d->nf && !d->face_regs)) return hb_fail(ctx, HB_ERR_INVALID, "mesh: region arrays == NULL");
	if (d->nlists && !d->lists) return hb_fail(ctx, HB_ERR_INVALID, "mesh: lists == NULL");
	HB_TRY(build_slot_table(ctx, d->off_reg_vtx, d->reg_vtxlist, d->nregs_vtx, d->nlists, m->h_slot_vtx, nullptr));
	HB_TRY(build_slot_table(ctx, d->off_reg_face, d->reg_facelist, d->nregs_face, d->nlists, m->h_slot_face, nullptr));
	HB_TRY(build_slot_table(ctx, d->off_reg_corner, d->reg_cornerlist, d->nregs_face, d->nlists, m->h_slot_corner, &m->reg_ncorner));
	m->any_corner = false;
	for (int c : m->reg_ncorner) m->any_corner = m->any_corner || c > 0;
	// a list may only be bound through the class it declares (Attr::target)
	for (int l = 0; l < d->nlists; ++l) {
		const int cls = d->lists[l].target;
		for (int r = 0; r < d->nregs_vtx; ++r)
			if (m->h_slot_vtx[(size_t)r * d->nlists + l] >= 0 && cls != HB_VTX) return hb_fail(ctx, HB_ERR_UNSUPPORTED, "list %d bound to vertices but declared with target %d", l, cls);
		for (int r = 0; r < d->nregs_face; ++r) {
			if (m->h_slot_face[(size_t)r * d->nlists + l] >= 0 && cls != HB_FACE) return hb_fail(ctx, HB_ERR_UNSUPPORTED, "list %d bound to faces but declared with target %d", l, cls);
			if (m->h_slot_corner[(size_t)r * d->nlists + l] >= 0 && cls != HB_CORNER) return hb_fail(ctx, HB_ERR_UNSUPPORTED, "list %d bound to corners but declared with target %d", l, cls);
		}
	}
	HB_TRY(upload(m, (void **)&m->d_edges_raw, d->edges, 12 * (size_t)d->ne));
	HB_TRY(upload(m, (void **)&m->d_face_off, d->face_off, sizeof(uint32_t) * ((size_t)d->nf + 1)));
	HB_TRY(upload(m, (void **)&m->d_order, d->order, 8 * (size_t)d->norder));
	if (d->order_f && !vertex_only) HB_TRY(upload(m, (void **)&m->d_order_f, d->order_f, 8 * (size_t)d->norder_f));
	HB_TRY(upload(m, (void **)&m->d_vtx_regs, d->vtx_regs, sizeof(uint16_t) * (size_t)d->nv));
	// everything K0 / K3 / K4 read is on its way: first milestone of the copy stream
	if (m->async_copy) HB_CUDA(ctx, cudaEventRecord(ctx->ev_up[0], ctx->copy_stream));
	if (!vertex_only) {
		HB_TRY(upload(m, (void **)&m->d_face_regs, d->face_regs, sizeof(uint16_t) * (size_t)d->nf));
		HB_TRY(upload(m, (void **)&m->d_bind_face, d->bind_face_attr, sizeof(uint32_t) * (size_t)d->nf * d->nb_face));
	}
	HB_TRY(upload(m, (void **)&m->d_bind_vtx, d->bind_vtx_attr, sizeof(uint32_t) * (size_t)d->nv * d->nb_vtx));
	if (!vertex_only) HB_TRY(upload(m, (void **)&m->d_bind_corner, d->bind_corner_attr, sizeof(uint32_t) * (size_t)d->ne * d->nb_corner));
	if (vertex_only) m->any_corner = false;
	HB_TRY(upload(m, (void **)&m->d_slot_vtx, m->h_slot_vtx.data(), sizeof(int16_t) * m->h_slot_vtx.size()));
	HB_TRY(upload(m, (void **)&m->d_slot_face, m->h_slot_face.data(), sizeof(int16_t) * m->h_slot_face.size()));
	HB_TRY(upload(m, (void **)&m->d_slot_corner, m->h_slot_corner.data(), sizeof(int16_t) * m->h_slot_corner.size()));
	for (int l = 0; l < d->nlists; ++l) {
		HB_TRY(add_list(m, d->lists[l], true));
		if (d->emit_type && d->emit_type[l]) {
			if (!d->emit_count) return hb_fail(ctx, HB_ERR_INVALID, "mesh: emit_type without emit_count");
			DevList &dl = m->lists.back();
			dl.emit_count = d->emit_count[l];
			HB_TRY(upload(m, (void **)&dl.d_emit_type, d->emit_type[l], dl.emit_count));
		}
	}
	if (m->async_copy) HB_CUDA(ctx, cudaEventRecord(ctx->ev_up[1], ctx->copy_stream));
	return 0;
}

// host-buffer entry points: connectivity stages under the tail of the upload, everything else behind it
static int upload_overlapped(hb_ctx *ctx, const hb_mesh_desc *mesh, hb_dmesh *m, bool vertex_only)
{
	m->async_copy = true;
	HB_TRY(dmesh_upload_impl(ctx, mesh, m, vertex_only));
	HB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_up[0], 0));
	HB_TRY(hb_build_conn(m));
	bool need_v = false;
	for (int l = 0; l < m->nlists; ++l)
		if (m->lists[l].p.target == HB_VTX && m->lists[l].p.ncomp) need_v = true;
	if (need_v) HB_TRY(hb_build_vertex_candidates(m));
	HB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_up[1], 0));
	return 0;
}

extern "C" int hb_dmesh_upload(hb_ctx *ctx, const hb_mesh_desc *d, hb_dmesh **out)
{
	*out = nullptr;
	HB_CUDA(ctx, cudaSetDevice(ctx->device));
	hb_dmesh *m = new hb_dmesh();
	int rc = dmesh_upload_impl(ctx, d, m);
	if (rc == 0) {
		cudaError_t e = cudaStreamSynchronize(ctx->stream); // the host buffers may be reused after return
		if (e != cudaSuccess) rc = hb_fail(ctx, HB_ERR_CUDA, "upload failed: %s", cudaGetErrorString(e));
	}
	if (rc) { hb_dmesh_free(m); return rc; }
	*out = m;
	return 0;
}

extern "C" int hb_dmesh_quantize(hb_dmesh *m, uint32_t l, const uint8_t *new_quant, const uint8_t *groups)
{
	if (l >= m->lists.size()) return hb_fail(m->ctx, HB_ERR_INVALID, "list %u out of range", l);
	HB_CUDA(m->ctx, cudaSetDevice(m->ctx->device));
	HB_TRY(hb_list_bounds(m, l));
	HB_TRY(hb_list_scale(m, l, groups));
	return hb_list_requant(m, l, new_quant);
}

extern "C" int hb_dmesh_dequantize(hb_dmesh *m, uint32_t l)
{
	if (l >= m->lists.size()) return hb_fail(m->ctx, HB_ERR_INVALID, "list %u out of range", l);
	HB_CUDA(m->ctx, cudaSetDevice(m->ctx->device));
	uint8_t zero[HB_MAX_COMP] = { 0 };
	return hb_list_requant(m, l, zero);
}

// bounds rows for list l supplied by the host (decode side: min / max come from the .hry header,
// the scale from set_scale)
extern "C" int hb_dmesh_set_bounds(hb_dmesh *m, uint32_t l, const void *min_row, const void *max_row, const void *scale_row)
{
	if (l >= m->lists.size()) return hb_fail(m->ctx, HB_ERR_INVALID, "list %u out of range", l);
	DevList &dl = m->lists[l];
	const size_t s = dl.p.stride;
	HB_CUDA(m->ctx, cudaSetDevice(m->ctx->device));
	if (min_row) HB_CUDA(m->ctx, cudaMemcpyAsync(dl.d_bounds, min_row, s, cudaMemcpyHostToDevice, m->ctx->stream));
	if (max_row) HB_CUDA(m->ctx, cudaMemcpyAsync(dl.d_bounds + s, max_row, s, cudaMemcpyHostToDevice, m->ctx->stream));
	if (scale_row) HB_CUDA(m->ctx, cudaMemcpyAsync(dl.d_bounds + 2 * s, scale_row, s, cudaMemcpyHostToDevice, m->ctx->stream));
	HB_CUDA(m->ctx, cudaStreamSynchronize(m->ctx->stream));
	return 0;
}

extern "C" int hb_dmesh_fetch_bounds(hb_dmesh *m, uint32_t l, void *min_row, void *max_row, void *scale_row)
{
	if (l >= m->lists.size()) return hb_fail(m->ctx, HB_ERR_INVALID, "list %u out of range", l);
	DevList &dl = m->lists[l];
	const size_t s = dl.p.stride;
	HB_CUDA(m->ctx, cudaSetDevice(m->ctx->device));
	if (min_row) HB_CUDA(m->ctx, cudaMemcpyAsync(min_row, dl.d_bounds, s, cudaMemcpyDeviceToHost, m->ctx->stream));
	if (max_row) HB_CUDA(m->ctx, cudaMemcpyAsync(max_row, dl.d_bounds + s, s, cudaMemcpyDeviceToHost, m->ctx->stream));
	if (scale_row) HB_CUDA(m->ctx, cudaMemcpyAsync(scale_row, dl.d_bounds + 2 * s, s, cudaMemcpyDeviceToHost, m->ctx->stream));
	HB_CUDA(m->ctx, cudaStreamSynchronize(m->ctx->stream));
	return hb_check_device_error(m->ctx, "bounds");
}

extern "C" int hb_dmesh_fetch_rows(hb_dmesh *m, uint32_t l, void *rows_out)
{
	if (l >= m->lists.size()) return hb_fail(m->ctx, HB_ERR_INVALID, "list %u out of range", l);
	DevList &dl = m->lists[l];
	HB_CUDA(m->ctx, cudaSetDevice(m->ctx->device));
	const size_t bytes = (size_t)dl.p.nrows * dl.p.stride;
	if (bytes) HB_CUDA(m->ctx, cudaMemcpyAsync(rows_out, dl.p.rows, bytes, cudaMemcpyDeviceToHost, m->ctx->stream));
	HB_CUDA(m->ctx, cudaStreamSynchronize(m->ctx->stream));
	return hb_check_device_error(m->ctx, "fetch rows");
}

// keep / restore a device copy of all rows and quantization states (bench loops re-run stages
// that work in place)
extern "C" int hb_dmesh_snapshot(hb_dmesh *m)
{
	HB_CUDA(m->ctx, cudaSetDevice(m->ctx->device));
	for (DevList &dl : m->lists) {
		const size_t bytes = (size_t)dl.p.nrows * dl.p.stride;
		HB_TRY(hb_dalloc(m, (void **)&dl.d_rows_backup, bytes));
		if (bytes) HB_CUDA(m->ctx, cudaMemcpyAsync(dl.d_rows_backup, dl.p.rows, bytes, cudaMemcpyDeviceToDevice, m->ctx->stream));
		memcpy(dl.backup_quant, dl.p.quant, HB_MAX_COMP);
	}
	return 0;
}
extern "C" int hb_dmesh_restore(hb_dmesh *m)
{
	HB_CUDA(m->ctx, cudaSetDevice(m->ctx->device));
	for (DevList &dl : m->lists) {
		if (!dl.d_rows_backup) return hb_fail(m->ctx, HB_ERR_INVALID, "restore without snapshot");
		const size_t bytes = (size_t)dl.p.nrows * dl.p.stride;
		if (bytes) HB_CUDA(m->ctx, cudaMemcpyAsync(dl.p.rows, dl.d_rows_backup, bytes, cudaMemcpyDeviceToDevice, m->ctx->stream));
		hb_list_desc tmp;
		tmp.ncomp = (uint16_t)dl.p.ncomp; tmp.rows = dl.p.rows; tmp.nrows = dl.p.nrows; tmp.stride = dl.p.stride; tmp.target = (uint8_t)dl.p.target;
		for (int j = 0; j < dl.p.ncomp; ++j) { tmp.type[j] = dl.p.type[j]; tmp.quant[j] = dl.backup_quant[j]; tmp.offset[j] = dl.p.offset[j]; }
		hb_fill_list_params(dl.p, tmp);
	}
	// derived data is recomputed by the next encode / decode
	m->conn_ready = m->vcand_ready = m->ccand_ready = false;
	m->encoded = false;
	return 0;
}

// diagnostics of the speculative vertex decoder for list l: sweeps, hypothesis sweeps, plain sweeps,
// ranks finalized by hypothesis sweeps, sweeps stopped by an unknown offset
extern "C" int hb_dmesh_decode_stats(hb_dmesh *m, uint32_t l, uint64_t *out8)
{
	if (l >= m->lists.size()) return hb_fail(m->ctx, HB_ERR_INVALID, "list %u out of range", l);
	for (int k = 0; k < 8; ++k) out8[k] = 0;
	if (!m->lists[l].d_spec_stats) return 0;
	HB_CUDA(m->ctx, cudaSetDevice(m->ctx->device));
	HB_CUDA(m->ctx, cudaMemcpyAsync(out8, m->lists[l].d_spec_stats, 64, cudaMemcpyDeviceToHost, m->ctx->stream));
	HB_CUDA(m->ctx, cudaStreamSynchronize(m->ctx->stream));
	return 0;
}

extern "C" int hb_dmesh_encode(hb_dmesh *m)
{
	HB_CUDA(m->ctx, cudaSetDevice(m->ctx->device));
	return hb_encode_lists(m);
}

extern "C" int hb_dmesh_decode(hb_dmesh *m)
{
	HB_CUDA(m->ctx, cudaSetDevice(m->ctx->device));
	return hb_decode_lists(m);
}

// ------------------------------------------------------------------------------------------------
// stream download
// ------------------------------------------------------------------------------------------------
// Output streams live in page-locked host memory (device -> host copies into pageable memory run at
// a fraction of the link rate).  Page-locking is expensive, so blocks are recycled through a small
// process-wide cache: hb_streams_free() parks them, the next fetch of a similar size reuses them.
namespace {
struct PinnedCache {
	std::mutex mu;
	std::vector<std::pair<void *, size_t>> free_blocks;
	std::unordered_map<void *, std::pair<size_t, bool>> live; // ptr -> (capacity, pinned)
	size_t cached_bytes = 0;
	static constexpr size_t kMaxCached = (size_t)8 << 30;
	void *alloc(size_t bytes)
	{
		if (bytes == 0) bytes = 1;
		bytes = (bytes + 4095) & ~(size_t)4095;
		{
			std::lock_guard<std::mutex> g(mu);
			size_t best = free_blocks.size();
			for (size_t k = 0; k < free_blocks.size(); ++k)
				if (free_blocks[k].second >= bytes && free_blocks[k].second <= 2 * bytes + (1 << 20) && (best == free_blocks.size() || free_blocks[k].second < free_blocks[best].second)) best = k;
			if (best != free_blocks.size()) {
				void *p = free_blocks[best].first;
				const size_t cap = free_blocks[best].second;
				free_blocks.erase(free_blocks.begin() + (long)best);
				cached_bytes -= cap;
				live[p] = { cap, true };
				return p;
			}
		}
		void *p = nullptr;
		bool pinned = cudaHostAlloc(&p, bytes, cudaHostAllocDefault) == cudaSuccess;
		if (!pinned) { cudaGetLastError(); p = malloc(bytes); }
		if (!p) return nullptr;
		std::lock_guard<std::mutex> g(mu);
		live[p] = { bytes, pinned };
		return p;
	}
	void release(void *p)
	{
		if (!p) return;
		std::unique_lock<std::mutex> g(mu);
		auto it = live.find(p);
		if (it == live.end()) { g.unlock(); free(p); return; }
		const size_t cap = it->second.first;
		const bool pinned = it->second.second;
		live.erase(it);
		if (pinned && cached_bytes + cap <= kMaxCached) {
			free_blocks.emplace_back(p, cap);
			cached_bytes += cap;
			return;
		}
		g.unlock();
		if (pinned) cudaFreeHost(p); else free(p);
	}
};
PinnedCache g_pinned;
} // namespace

extern "C" void hb_streams_free(hb_streams *s)
{
	if (!s) return;
	if (s->lists) {
		for (int l = 0; l < s->nlists; ++l) { g_pinned.release(s->lists[l].type); g_pinned.release(s->lists[l].aux); g_pinned.release(s->lists[l].symbols); g_pinned.release(s->lists[l].hist); }
		free(s->lists);
	}
	g_pinned.release(s->reg_vtx);
	g_pinned.release(s->reg_face);
	free(s);
}

__global__ void k_region_stream(const uint32_t *__restrict__ ent, const uint4 *__restrict__ he, const uint16_t *__restrict__ regs, uint32_t n, int is_face, uint16_t *__restrict__ out)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint32_t e = is_face ? he[ent[i]].w : ent[i];
	out[i] = regs[e];
}

extern "C" int hb_dmesh_fetch_streams(hb_dmesh *m, hb_streams **out)
{
	hb_ctx *ctx = m->ctx;
	*out = nullptr;
	if (!m->encoded) return hb_fail(ctx, HB_ERR_INVALID, "fetch_streams before encode");
	HB_CUDA(ctx, cudaSetDevice(ctx->device));
	hb_streams *s = (hb_streams *)calloc(1, sizeof(hb_streams));
	if (!s) return hb_fail(ctx, HB_ERR_NOMEM, "out of host memory");
	s->n_vtx = m->norder;
	s->n_face = m->norder_f;
	s->nlists = m->nlists;
	// a mesh with a single vertex (face) region has an all-zero region stream: NULL stands for it
	const bool reg_v = m->nregs_vtx > 1, reg_f = m->nregs_face > 1;
	s->reg_vtx = reg_v ? (uint16_t *)g_pinned.alloc(sizeof(uint16_t) * ((size_t)m->norder + 1)) : nullptr;
	s->reg_face = reg_f ? (uint16_t *)g_pinned.alloc(sizeof(uint16_t) * ((size_t)m->norder_f + 1)) : nullptr;
	s->lists = (hb_list_streams *)calloc((size_t)m->nlists + 1, sizeof(hb_list_streams));
	int rc = 0;
	uint16_t *d_reg = nullptr;
	const uint32_t nmax = m->norder > m->norder_f ? m->norder : m->norder_f;
#define FETCH_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { rc = hb_fail(ctx, HB_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e__)); goto fail; } } while (0)
	if ((reg_v && !s->reg_vtx) || (reg_f && !s->reg_face) || !s->lists) { rc = hb_fail(ctx, HB_ERR_NOMEM, "out of host memory"); goto fail; }
	// region symbol streams (io.h:109-116): region of every traversed vertex / face
	FETCH_CUDA(cudaMallocAsync((void **)&d_reg, sizeof(uint16_t) * ((size_t)nmax + 1), ctx->stream));
	if (m->norder && reg_v) {
		k_region_stream<<<hb_div_up(m->norder, 256), 256, 0, ctx->stream>>>(m->d_ord_v, m->d_he, m->d_vtx_regs, m->norder, 0, d_reg);
		ctx->launches++;
		FETCH_CUDA(cudaMemcpyAsync(s->reg_vtx, d_reg, sizeof(uint16_t) * m->norder, cudaMemcpyDeviceToHost, ctx->stream));
	}
	if (m->norder_f && reg_f) {
		k_region_stream<<<hb_div_up(m->norder_f, 256), 256, 0, ctx->stream>>>(m->d_ford_h, m->d_he, m->d_face_regs, m->norder_f, 1, d_reg);
		ctx->launches++;
		FETCH_CUDA(cudaMemcpyAsync(s->reg_face, d_reg, sizeof(uint16_t) * m->norder_f, cudaMemcpyDeviceToHost, ctx->stream));
	}
	FETCH_CUDA(cudaFreeAsync(d_reg, ctx->stream));
	// counts first (sizes of the compacted streams)
	for (int l = 0; l < m->nlists; ++l) {
		DevList &dl = m->lists[l];
		hb_list_streams &ls = s->lists[l];
		ls.sym_stride = dl.p.sym_stride;
		ls.n_emit = 0;
		ls.n_data = 0;
		if (!dl.n_elems) continue;
		if (dl.d_ek) FETCH_CUDA(cudaMemcpyAsync(&ls.n_emit, dl.d_ek + dl.n_elems, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
		else ls.n_emit = dl.n_elems;
		FETCH_CUDA(cudaMemcpyAsync(&ls.n_data, dl.d_dord + dl.n_elems, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
		FETCH_CUDA(cudaMemcpyAsync(ls.type_hist, dl.d_type_hist, sizeof(uint64_t) * 4, cudaMemcpyDeviceToHost, ctx->stream));
	}
	FETCH_CUDA(cudaStreamSynchronize(ctx->stream));
	for (int l = 0; l < m->nlists; ++l) {
		DevList &dl = m->lists[l];
		hb_list_streams &ls = s->lists[l];
		// every emission is a DATA row (no shared attribute rows): the type and history-offset streams are
		// all zero -- NULL stands for them, nothing is copied
		const bool all_data = ls.type_hist[HB_HIST] + ls.type_hist[HB_LHIST] == 0;
		ls.type = all_data ? nullptr : (uint8_t *)g_pinned.alloc((size_t)ls.n_emit + 1);
		ls.aux = all_data ? nullptr : (uint32_t *)g_pinned.alloc(sizeof(uint32_t) * ((size_t)ls.n_emit + 1));
		ls.symbols = (uint8_t *)g_pinned.alloc((size_t)ls.n_data * ls.sym_stride + 1);
		ls.hist = (uint64_t *)g_pinned.alloc(sizeof(uint64_t) * ((size_t)ls.sym_stride * 256 + 4));
		if ((!all_data && (!ls.type || !ls.aux)) || !ls.symbols || !ls.hist) { rc = hb_fail(ctx, HB_ERR_NOMEM, "out of host memory"); goto fail; }
		memset(ls.hist, 0, sizeof(uint64_t) * ((size_t)ls.sym_stride * 256 + 4));
		if (!dl.n_elems) continue;
		if (ls.n_emit && !all_data) {
			FETCH_CUDA(cudaMemcpyAsync(ls.type, dl.d_type, ls.n_emit, cudaMemcpyDeviceToHost, ctx->stream));
			FETCH_CUDA(cudaMemcpyAsync(ls.aux, dl.d_aux, sizeof(uint32_t) * ls.n_emit, cudaMemcpyDeviceToHost, ctx->stream));
		}
		if ((size_t)ls.n_data * ls.sym_stride) FETCH_CUDA(cudaMemcpyAsync(ls.symbols, dl.d_sym, (size_t)ls.n_data * ls.sym_stride, cudaMemcpyDeviceToHost, ctx->stream));
		FETCH_CUDA(cudaMemcpyAsync(ls.hist, dl.d_hist, sizeof(uint64_t) * ((size_t)ls.sym_stride * 256 + 4), cudaMemcpyDeviceToHost, ctx->stream));
	}
	FETCH_CUDA(cudaStreamSynchronize(ctx->stream));
	for (int l = 0; l < m->nlists; ++l) {
		hb_list_streams &ls = s->lists[l];
		for (int k = 0; k < 4; ++k) ls.type_hist[k] = ls.hist[(size_t)ls.sym_stride * 256 + k];
	}
	rc = hb_check_device_error(ctx, "attribute encode");
	if (rc) goto fail;
	*out = s;
	return 0;
fail:
	hb_streams_free(s);
	return rc;
#undef FETCH_CUDA
}

// ------------------------------------------------------------------------------------------------
// host-buffer entry points (the drop-in boundary)
// ------------------------------------------------------------------------------------------------
struct PhaseTimer {
	hb_ctx *ctx;
	explicit PhaseTimer(hb_ctx *c) : ctx(c) { ctx->kernel_ms = ctx->copy_ms = 0.f; }
	void mark(int i) { cudaEventRecord(ctx->ev[i], ctx->stream); }
	void finish(int n_marks)
	{
		// marks: 0 start, 1 after H2D, 2 after kernels, 3 after D2H
		cudaEventSynchronize(ctx->ev[n_marks - 1]);
		float a = 0, b = 0, c = 0;
		cudaEventElapsedTime(&a, ctx->ev[0], ctx->ev[1]);
		cudaEventElapsedTime(&b, ctx->ev[1], ctx->ev[2]);
		if (n_marks > 3) cudaEventElapsedTime(&c, ctx->ev[2], ctx->ev[3]);
		ctx->kernel_ms = b;
		ctx->copy_ms = a + c;
	}
};

static int single_list_mesh(hb_ctx *ctx, const hb_list_desc *list, hb_dmesh *m)
{
	m->ctx = ctx;
	m->nlists = 1;
	return add_list(m, *list, false);
}

extern "C" int hb_bounds(hb_ctx *ctx, const hb_list_desc *list, void *min_row, void *max_row)
{
	HB_CUDA(ctx, cudaSetDevice(ctx->device));
	hb_dmesh *m = new hb_dmesh();
	PhaseTimer t(ctx);
	t.mark(0);
	int rc = single_list_mesh(ctx, list, m);
	t.mark(1);
	if (rc == 0) rc = hb_list_bounds(m, 0);
	t.mark(2);
	if (rc == 0 && list->ncomp) rc = hb_dmesh_fetch_bounds(m, 0, min_row, max_row, nullptr);
	t.mark(3);
	t.finish(4);
	hb_dmesh_free(m);
	return rc;
}

extern "C" int hb_requant(hb_ctx *ctx, hb_list_desc *list, const uint8_t *new_quant, const void *min_row, const void *scale_row)
{
	HB_CUDA(ctx, cudaSetDevice(ctx->device));
	hb_dmesh *m = new hb_dmesh();
	PhaseTimer t(ctx);
	t.mark(0);
	int rc = single_list_mesh(ctx, list, m);
	if (rc == 0 && list->ncomp) rc = hb_dmesh_set_bounds(m, 0, min_row, nullptr, scale_row);
	t.mark(1);
	if (rc == 0) rc = hb_list_requant(m, 0, new_quant);
	t.mark(2);
	if (rc == 0 && list->ncomp) rc = hb_dmesh_fetch_rows(m, 0, list->rows);
	t.mark(3);
	t.finish(4);
	if (rc == 0)
		for (int j = 0; j < list->ncomp; ++j) list->quant[j] = new_quant[j];
	hb_dmesh_free(m);
	return rc;
}

extern "C" int hb_twin_match(hb_ctx *ctx, uint32_t nv, uint32_t nf, const uint32_t *face_off, const void *org, uint32_t org_stride,
                             void *edges_out)
{
	if (nf && (!face_off || !org || !edges_out)) return hb_fail(ctx, HB_ERR_INVALID, "twin_match: NULL argument");
	if (org_stride != 4 && org_stride != 12) return hb_fail(ctx, HB_ERR_INVALID, "twin_match: org_stride must be 4 (packed) or 12 (edge records)");
	const uint32_t ne = nf ? face_off[nf] : 0;
	if (nf && face_off[0] != 0) return hb_fail(ctx, HB_ERR_INVALID, "twin_match: face_off[0] != 0");
	HB_CUDA(ctx, cudaSetDevice(ctx->device));
	hb_dmesh *m = new hb_dmesh();
	m->ctx = ctx;
	PhaseTimer t(ctx);
	t.mark(0);
	uint32_t *d_face_off = nullptr, *d_org = nullptr, *d_out = nullptr;
	int rc = hb_dalloc_t(m, &d_face_off, (size_t)nf + 1);
	if (rc == 0) rc = hb_dalloc_t(m, &d_out, (size_t)ne * 3);
	if (rc == 0 && org_stride == 4) rc = hb_dalloc_t(m, &d_org, (size_t)ne);
	auto up = [&](void *dst, const void *src, size_t bytes) {
		if (rc != 0 || !bytes) return;
		if (cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)
			rc = hb_fail(ctx, HB_ERR_CUDA, "twin_match: upload failed: %s", cudaGetErrorString(cudaGetLastError()));
	};
	up(d_face_off, face_off, sizeof(uint32_t) * ((size_t)nf + 1));
	if (org_stride == 4) up(d_org, org, sizeof(uint32_t) * (size_t)ne);
	else up(d_out, org, 12 * (size_t)ne); // the records themselves: org read in place, twin words overwritten
	t.mark(1);
	if (rc == 0) rc = hb_twin_build(m, nv, nf, ne, d_face_off, org_stride == 4 ? d_org : d_out, org_stride / 4, d_out);
	t.mark(2);
	if (rc == 0 && ne && cudaMemcpyAsync(edges_out, d_out, 12 * (size_t)ne, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
		rc = hb_fail(ctx, HB_ERR_CUDA, "twin_match: download failed: %s", cudaGetErrorString(cudaGetLastError()));
	t.mark(3);
	t.finish(4);
	if (rc == 0) rc = hb_check_device_error(ctx, "twin matching");
	hb_dmesh_free(m);
	return rc;
}

extern "C" int hb_attr_encode(hb_ctx *ctx, const hb_mesh_desc *mesh, hb_streams **out)
{
	*out = nullptr;
	HB_CUDA(ctx, cudaSetDevice(ctx->device));
	hb_dmesh *m = new hb_dmesh();
	PhaseTimer t(ctx);
	t.mark(0);
	int rc = upload_overlapped(ctx, mesh, m, false);
	t.mark(1);
	if (rc == 0) rc = hb_encode_lists(m);
	t.mark(2);
	if (rc == 0) rc = hb_dmesh_fetch_streams(m, out);
	t.mark(3);
	t.finish(4);
	hb_dmesh_free(m);
	return rc;
}

extern "C" int hb_attr_decode(hb_ctx *ctx, const hb_mesh_desc *mesh)
{
	HB_CUDA(ctx, cudaSetDevice(ctx->device));
	hb_dmesh *m = new hb_dmesh();
	PhaseTimer t(ctx);
	t.mark(0);
	bool vertex_only = true;
	for (int l = 0; l < mesh->nlists; ++l)
		if (mesh->lists[l].target != HB_VTX && mesh->lists[l].ncomp && mesh->lists[l].nrows) vertex_only = false;
	int rc = upload_overlapped(ctx, mesh, m, vertex_only);
	t.mark(1);
	if (rc == 0) rc = hb_decode_lists(m);
	t.mark(2);
	for (int l = 0; rc == 0 && l < mesh->nlists; ++l)
		if (mesh->lists[l].ncomp && mesh->lists[l].nrows) rc = hb_dmesh_fetch_rows(m, (uint32_t)l, mesh->lists[l].rows);
	t.mark(3);
	t.finish(4);
	if (rc == 0) rc = hb_check_device_error(ctx, "attribute decode");
	hb_dmesh_free(m);
	return rc;
}
