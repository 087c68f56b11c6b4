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

extern "C" int hb_ctx_create(int device, hb_ctx **out) { return hb_ctx_create_prio(device, 0, out); }

extern "C" int hb_ctx_create_prio(int device, int high_priority, hb_ctx **out)
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
	// a high-priority context: the blocks of its kernels are placed before pending blocks of other contexts' kernels
	// (the chain-bound decode of a single mesh next to an encode job that only needs the SMs the chain leaves idle)
	int prio_lo = 0, prio_hi = 0;
	CREATE_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
	const int prio = high_priority ? prio_hi : prio_lo;
	CREATE_TRY(cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prio));
	CREATE_TRY(cudaStreamCreateWithPriority(&ctx->copy_stream, cudaStreamNonBlocking, prio));
	CREATE_TRY(cudaStreamCreateWithPriority(&ctx->out_stream, cudaStreamNonBlocking, prio));
	for (int i = 0; i < 6; ++i) CREATE_TRY(cudaEventCreate(&ctx->ev[i]));
	CREATE_TRY(cudaMalloc((void **)&ctx->d_err, sizeof(int)));
	CREATE_TRY(cudaMemset(ctx->d_err, 0, sizeof(int)));
	CREATE_TRY(cudaMallocHost((void **)&ctx->h_err, sizeof(int)));
	*ctx->h_err = 0;
	CREATE_TRY(cudaMallocHost((void **)&ctx->h_flag, sizeof(uint32_t)));
	CREATE_TRY(cudaEventCreateWithFlags(&ctx->ev_flag, cudaEventDisableTiming));
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
	for (hb_ctx::RowEntry &e : ctx->rows_cached) cudaFreeAsync(e.dev, ctx->stream);
	ctx->rows_cached.clear();
	if (ctx->stream) cudaStreamSynchronize(ctx->stream);
	if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
	if (ctx->out_stream) { cudaStreamSynchronize(ctx->out_stream); cudaStreamDestroy(ctx->out_stream); }
	for (int i = 0; i < 6; ++i)
		if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
	for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
	if (ctx->d_err) cudaFree(ctx->d_err);
	if (ctx->h_err) cudaFreeHost(ctx->h_err);
	if (ctx->h_flag) cudaFreeHost(ctx->h_flag);
	if (ctx->ev_flag) cudaEventDestroy(ctx->ev_flag);
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
extern "C" uint64_t hb_h2d_bytes(hb_ctx *ctx) { return ctx->h2d_bytes; }

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

// everything queued on `other` so far happens before whatever is queued on `ctx` from now on (two contexts of one
// device working on independent meshes: e.g. the encode of one mesh next to the decode of another)
extern "C" int hb_ctx_wait(hb_ctx *ctx, hb_ctx *other)
{
	if (ctx->device != other->device) return hb_fail(ctx, HB_ERR_INVALID, "hb_ctx_wait: contexts of different devices");
	HB_CUDA(ctx, cudaSetDevice(ctx->device));
	cudaEvent_t e;
	HB_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
	HB_CUDA(ctx, cudaEventRecord(e, other->stream));
	HB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, e, 0));
	HB_CUDA(ctx, cudaEventDestroy(e)); // released once it has completed
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
	// upload buffers of the host-buffer entry points come from the copy stream (the upload does not have to wait for
	// the kernels of the previous mesh / group that are still queued on the compute stream)
	cudaStream_t st = m->alloc_on_copy_stream ? m->ctx->copy_stream : m->ctx->stream;
	cudaError_t e = cudaMallocAsync(p, bytes, st);
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

// ---- gathered uploads (hb_internal.cuh, hb_dmesh::up_pending) -------------------------------------------------------
#define UP_CHUNK (64u << 10)   // bytes one CTA moves per turn
#define UP_CTAS 48             // 48 x 256 threads x 4 x 16 bytes in flight: far more than link rate x link latency
// every CTA takes chunks of the concatenated descriptor list in turn; the HOST side is always read with aligned 16-byte
// loads (a misaligned head / tail goes byte by byte), the device side is written as its alignment allows
__global__ void __launch_bounds__(256) k_upload_gather(const hb_dmesh::UpDesc *__restrict__ descs, const uint32_t *__restrict__ first_chunk, uint32_t n, uint32_t total)
{
	for (uint32_t c = blockIdx.x; c < total; c += gridDim.x) {
		uint32_t lo = 0, hi = n;
		while (hi - lo > 1) {
			const uint32_t mid = (lo + hi) >> 1;
			if (first_chunk[mid] <= c) lo = mid;
			else hi = mid;
		}
		const hb_dmesh::UpDesc d = descs[lo];
		const unsigned long long off = (unsigned long long)(c - first_chunk[lo]) * UP_CHUNK;
		const uint32_t len = (uint32_t)(d.bytes - off < UP_CHUNK ? d.bytes - off : UP_CHUNK);
		const uint8_t *s = (const uint8_t *)d.src + off;
		uint8_t *t = (uint8_t *)d.dst + off;
		const uint32_t head = min(len, (uint32_t)((16u - (uint32_t)((uintptr_t)s & 15u)) & 15u));
		if (threadIdx.x < head) t[threadIdx.x] = s[threadIdx.x];
		const uint32_t nvec = (len - head) >> 4;
		const uint4 *sv = (const uint4 *)(s + head);
		uint8_t *tv = t + head;
		const uint32_t dal = (uint32_t)((uintptr_t)tv & 15u);
		if (dal == 0) {
#pragma unroll 4
			for (uint32_t v = threadIdx.x; v < nvec; v += 256) ((uint4 *)tv)[v] = sv[v];
		} else if ((dal & 3u) == 0) {
#pragma unroll 4
			for (uint32_t v = threadIdx.x; v < nvec; v += 256) {
				const uint4 q = sv[v];
				uint32_t *w = (uint32_t *)(tv + 16 * (size_t)v);
				w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w;
			}
		} else {
			for (uint32_t v = threadIdx.x; v < nvec; v += 256) {
				const uint4 q = sv[v];
				const uint32_t w[4] = { q.x, q.y, q.z, q.w };
				for (int k = 0; k < 16; ++k) tv[16 * (size_t)v + k] = (uint8_t)(w[k >> 2] >> (8 * (k & 3)));
			}
		}
		for (uint32_t k = head + 16 * nvec + threadIdx.x; k < len; k += 256) t[k] = s[k];
	}
}
// hands the queued descriptors to one kernel on the upload stream; called in front of everything that orders itself
// behind the uploads (events, the face-order check)
static int flush_uploads(hb_dmesh *m)
{
	if (m->up_pending.empty()) return 0;
	hb_ctx *ctx = m->ctx;
	cudaStream_t st = m->async_copy ? ctx->copy_stream : ctx->stream;
	const size_t n = m->up_pending.size();
	if (m->upload_mode == 2) {
		// one call for all copies of the stage: the copy engines run them back to back
		std::vector<void *> dsts(n), srcs(n);
		std::vector<size_t> sizes(n);
		for (size_t k = 0; k < n; ++k) { dsts[k] = m->up_pending[k].dst; srcs[k] = const_cast<void *>(m->up_pending[k].src); sizes[k] = (size_t)m->up_pending[k].bytes; }
		cudaMemcpyAttributes at;
		memset(&at, 0, sizeof(at));
		at.srcAccessOrder = cudaMemcpySrcAccessOrderStream; // the caller's buffers stay valid until the call returns, which is behind the copies
		size_t idx0 = 0, fail = 0;
		if (cudaMemcpyBatchAsync(dsts.data(), srcs.data(), sizes.data(), n, &at, &idx0, 1, &fail, st) == cudaSuccess) {
			m->up_pending.clear();
			return 0;
		}
		cudaGetLastError(); // a driver without batched copies: the kernel below
	}
	std::vector<uint32_t> first((size_t)n + 1, 0);
	for (size_t k = 0; k < n; ++k) {
		const unsigned long long nc = (m->up_pending[k].bytes + UP_CHUNK - 1) / UP_CHUNK;
		if (first[k] + nc > 0xffffffffull) return hb_fail(ctx, HB_ERR_UNSUPPORTED, "batch: upload stage larger than 2^48 bytes");
		first[k + 1] = first[k] + (uint32_t)nc;
	}
	hb_dmesh::UpDesc *d_desc = nullptr;
	uint32_t *d_first = nullptr;
	const bool keep = m->alloc_on_copy_stream;
	m->alloc_on_copy_stream = m->async_copy;
	int rc = hb_dalloc(m, (void **)&d_desc, sizeof(hb_dmesh::UpDesc) * n);
	if (rc == 0) rc = hb_dalloc(m, (void **)&d_first, sizeof(uint32_t) * (n + 1));
	m->alloc_on_copy_stream = keep;
	if (rc) return rc;
	// (pageable sources: staged by the runtime before the calls return)
	HB_CUDA(ctx, cudaMemcpyAsync(d_desc, m->up_pending.data(), sizeof(hb_dmesh::UpDesc) * n, cudaMemcpyHostToDevice, st));
	HB_CUDA(ctx, cudaMemcpyAsync(d_first, first.data(), sizeof(uint32_t) * (n + 1), cudaMemcpyHostToDevice, st));
	const uint32_t total = first[n];
	k_upload_gather<<<std::min<uint32_t>(total, UP_CTAS), 256, 0, st>>>(d_desc, d_first, (uint32_t)n, total);
	ctx->launches++;
	HB_CUDA(ctx, cudaGetLastError());
	m->up_pending.clear();
	return 0;
}

// host -> device copy of one uploaded array (on the copy stream for the host-buffer entry points)
static int copy_in(hb_dmesh *m, void *dst, const void *src, size_t bytes)
{
	if (!(bytes && src)) return 0;
	hb_ctx *ctx = m->ctx;
	if (m->gather_uploads) {
		// page-locked (allocated or registered) host memory is visible to the device: queue it for the gather kernel
		cudaPointerAttributes at;
		if (cudaPointerGetAttributes(&at, src) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer) {
			m->up_pending.push_back(hb_dmesh::UpDesc{ at.devicePointer, dst, (unsigned long long)bytes });
			ctx->h2d_bytes += bytes;
			return 0;
		}
		cudaGetLastError(); // (older runtimes report unregistered memory as an error)
	}
	HB_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, m->async_copy ? ctx->copy_stream : ctx->stream));
	ctx->h2d_bytes += bytes;
	return 0;
}
static int upload(hb_dmesh *m, void **dst, const void *src, size_t bytes)
{
	HB_TRY(hb_dalloc(m, dst, bytes));
	return copy_in(m, *dst, src, bytes);
}

// ---- row cache (hb_internal.cuh) ------------------------------------------------------------------
static void row_cache_drop(hb_ctx *ctx, const void *host)
{
	for (size_t k = 0; k < ctx->rows_cached.size();)
		if (ctx->rows_cached[k].host == host) {
			cudaFreeAsync(ctx->rows_cached[k].dev, ctx->stream);
			ctx->rows_cached.erase(ctx->rows_cached.begin() + (long)k);
		} else ++k;
}
// device rows of `L` if the cache holds them in exactly this state (removed from the cache: the caller owns them now)
static uint8_t *row_cache_take(hb_ctx *ctx, const hb_list_desc &L)
{
	if (!ctx->row_cache) return nullptr;
	for (size_t k = 0; k < ctx->rows_cached.size(); ++k) {
		hb_ctx::RowEntry &e = ctx->rows_cached[k];
		if (e.host != L.rows || e.nrows != L.nrows || e.stride != L.stride || e.ncomp != L.ncomp) continue;
		bool same = true;
		for (int j = 0; j < L.ncomp; ++j) same = same && e.quant[j] == L.quant[j];
		if (!same) continue;
		uint8_t *dev = e.dev;
		ctx->rows_cached.erase(ctx->rows_cached.begin() + (long)k);
		return dev;
	}
	return nullptr;
}
// hands the rows buffer of list 0 of a single-segment device mesh over to the cache (it is taken off the mesh's free list)
static void row_cache_put(hb_dmesh *m, int l, const void *host)
{
	hb_ctx *ctx = m->ctx;
	if (!ctx->row_cache || m->nseg != 1 || !host) return;
	DevList &dl = m->lists[l];
	if (!dl.p.rows || !dl.p.ncomp || !dl.p.nrows) return;
	for (size_t k = 0; k < m->allocs.size(); ++k)
		if (m->allocs[k] == dl.p.rows) { m->allocs.erase(m->allocs.begin() + (long)k); break; }
	row_cache_drop(ctx, host);
	while (ctx->rows_cached.size() >= 8) row_cache_drop(ctx, ctx->rows_cached.front().host);
	hb_ctx::RowEntry e;
	e.host = host; e.nrows = dl.p.nrows; e.stride = dl.p.stride; e.dev = dl.p.rows; e.ncomp = dl.p.ncomp;
	for (int j = 0; j < HB_MAX_COMP; ++j) e.quant[j] = j < dl.p.ncomp ? dl.p.quant[j] : 0;
	ctx->rows_cached.push_back(e);
	dl.p.rows = nullptr;
}
extern "C" int hb_ctx_set_row_cache(hb_ctx *ctx, int enable)
{
	HB_CUDA(ctx, cudaSetDevice(ctx->device));
	ctx->row_cache = enable != 0;
	if (!enable)
		while (!ctx->rows_cached.empty()) row_cache_drop(ctx, ctx->rows_cached.front().host);
	return 0;
}

static bool same_format(const hb_list_desc &a, const hb_list_desc &b)
{
	if (a.ncomp != b.ncomp || a.stride != b.stride || a.target != b.target) return false;
	for (int j = 0; j < a.ncomp; ++j)
		if (a.type[j] != b.type[j] || a.quant[j] != b.quant[j] || a.offset[j] != b.offset[j]) return false;
	return true;
}

// list l of all segments: rows concatenated (segment bases padded to multiples of 16 rows), one bounds triple per segment
static int add_list(hb_dmesh *m, const hb_mesh_desc *descs, const hb_list_desc *single, uint32_t nseg, int l, bool for_coder)
{
	hb_ctx *ctx = m->ctx;
	const hb_list_desc &L0 = single ? *single : descs[0].lists[l];
	DevList dl;
	memset(&dl.p, 0, sizeof dl.p);
	dl.h_rowbase.assign((size_t)nseg + 1, 0);
	dl.h_rownum.assign((size_t)nseg, 0);
	uint64_t acc = 0;
	for (uint32_t s = 0; s < nseg; ++s) {
		const hb_list_desc &L = single ? *single : descs[s].lists[l];
		HB_TRY(validate_list(ctx, L, for_coder));
		if (s && !same_format(L0, L)) return hb_fail(ctx, HB_ERR_INVALID, "batch: list %d of mesh %u has another format than in mesh 0", l, s);
		dl.h_rowbase[s] = (uint32_t)acc;
		dl.h_rownum[s] = L.nrows;
		acc += L.nrows;
		if (s + 1 < nseg) acc = (acc + 15) & ~(uint64_t)15;
		if (acc >= 0xfffffff0ull) return hb_fail(ctx, HB_ERR_UNSUPPORTED, "batch: more than 2^32 rows in list %d", l);
	}
	dl.h_rowbase[nseg] = (uint32_t)acc;
	hb_list_desc Lc = L0;
	Lc.nrows = (uint32_t)acc;
	hb_fill_list_params(dl.p, Lc);
	// the rows may still be on the device: only for calls that CONTINUE a pipeline (requant, encode) -- set_bounds and
	// decode start one, their host rows are new
	void *rows = m->take_cached_rows && nseg == 1 && L0.ncomp && L0.nrows ? row_cache_take(ctx, L0) : nullptr;
	const bool cached = rows != nullptr;
	if (cached) {
		m->allocs.push_back(rows);
		// the buffer was last touched on the compute stream: an upload stream that fills sibling buffers need not wait,
		// the kernels reading it run on the compute stream anyway
	} else {
		HB_TRY(hb_dalloc(m, &rows, (size_t)acc * L0.stride));
	}
	dl.p.rows = (uint8_t *)rows;
	dl.rows_from_cache = cached;
	if (L0.ncomp && !cached)
		for (uint32_t s = 0; s < nseg; ++s) {
			const hb_list_desc &L = single ? *single : descs[s].lists[l];
			HB_TRY(copy_in(m, dl.p.rows + (size_t)dl.h_rowbase[s] * L0.stride, L.rows, (size_t)L.nrows * L0.stride));
		}
	// segment tables of the list: bases (nseg + 1) and row counts (nseg), back to back
	std::vector<uint32_t> tab(dl.h_rowbase);
	tab.insert(tab.end(), dl.h_rownum.begin(), dl.h_rownum.end());
	void *dtab = nullptr;
	HB_TRY(upload(m, &dtab, tab.data(), sizeof(uint32_t) * tab.size())); // (pageable source: staged before the call returns)
	dl.d_rowbase = (uint32_t *)dtab;
	dl.d_rownum = dl.d_rowbase + nseg + 1;
	dl.bounds_pitch = (3 * (size_t)L0.stride + 16 + 15) & ~(size_t)15;
	void *b = nullptr;
	HB_TRY(hb_dalloc(m, &b, dl.bounds_pitch * nseg));
	dl.d_bounds = (uint8_t *)b;
	HB_CUDA(ctx, cudaMemsetAsync(b, 0, dl.bounds_pitch * nseg, m->alloc_on_copy_stream ? ctx->copy_stream : ctx->stream));
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
	if (m->ev_up[0]) cudaEventDestroy(m->ev_up[0]);
	if (m->ev_up[1]) cudaEventDestroy(m->ev_up[1]);
	if (m->ev_done) cudaEventDestroy(m->ev_done);
	delete m;
}

// device memory of a group goes back to the pool behind the work queued on `st` (pipelined batches); the host object
// stays (its segment tables are still needed for the stream views)
static void dmesh_release_device(hb_dmesh *m, cudaStream_t st)
{
	for (void *p : m->allocs) cudaFreeAsync(p, st);
	m->allocs.clear();
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

static bool same_region_table(const int32_t *off_a, const uint16_t *la, const int32_t *off_b, const uint16_t *lb, int nregs)
{
	for (int r = 0; r <= nregs; ++r)
		if (off_a[r] != off_b[r]) return false;
	for (int k = 0; k < off_a[nregs]; ++k)
		if (la[k] != lb[k]) return false;
	return true;
}

// Upload of `nseg` meshes of one schema as one device mesh (hb_internal.cuh, "segments").
// ---- encoder uploads: is the face order needed at all? ---------------------------------------------------------
// The traversal order of the faces (order_f, 8 bytes per face: 160 MB on the 10M-vertex mesh) only decides WHICH
// element of a FACE / CORNER list emits a row as DATA and which ones refer back to it (attrcode.h:296-319); the
// residuals of such a list depend on it through the same first-reference rule.  When no FACE / CORNER list carries
// components, the mesh has one face region (the region stream is constant) and no row of those lists is bound twice,
// every element emits DATA with an empty row whatever the order: the streams are those of the index order.  The first
// two conditions are properties of the schema; the third is checked on the device from the binding arrays, which are
// then uploaded FIRST: the answer (4 bytes) is back while the rest of the upload is still on the link.
static bool order_f_maybe_unneeded(const hb_mesh_desc *descs, uint32_t nseg)
{
	const hb_mesh_desc *d = &descs[0];
	bool any = false;
	for (uint32_t s = 0; s < nseg; ++s) any = any || (descs[s].nf && descs[s].order_f);
	if (!any || d->nregs_face != 1 || !d->lists) return false;
	const char *env = getenv("HARRY_B200_KEEP_ORDER_F"); // A/B runs and tests: always upload the face order
	if (env && env[0] == '1') return false;
	for (uint32_t s = 0; s < nseg; ++s) {
		const hb_mesh_desc &ds = descs[s];
		if (ds.nf == 0 && ds.nlists == d->nlists && ds.lists) continue; // no faces: nothing to order
		if (!ds.order_f || ds.norder_f != ds.nf || ds.nlists != d->nlists || !ds.lists) return false;
		for (int l = 0; l < ds.nlists; ++l)
			if (ds.lists[l].target != HB_VTX && ds.lists[l].ncomp) return false;
	}
	return true;
}
// element i of segment s binds row bind[i * nb + slot] of the segment's rows [0, nrows[s]): one bit per row
__global__ void __launch_bounds__(256) k_bind_twice(const uint32_t *__restrict__ bind, const uint32_t *__restrict__ elem_base, uint32_t nseg, uint32_t nb, uint32_t slot,
                                                    const uint32_t *__restrict__ bitbase, uint32_t *__restrict__ bits, uint32_t *__restrict__ twice)
{
	const uint32_t n = elem_base[nseg];
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const uint32_t s = hb_seg_find(elem_base, nseg, i);
		const uint32_t row = bind[(size_t)i * nb + slot];
		const uint32_t w0 = bitbase[s], nrow = (bitbase[s + 1] - w0) * 32u; // (rounded up: a row behind the end is caught by the encoder proper)
		if (row >= nrow) { *twice = 1u; continue; } // unbound / out of range: not the plain case
		const uint32_t bit = 1u << (row & 31u);
		if (atomicOr(&bits[w0 + (row >> 5)], bit) & bit) *twice = 1u;
	}
}

// vertex_only: the caller will only reconstruct vertex lists (hb_attr_decode of a mesh whose face and
// corner lists carry no components): the face-side arrays (order_f, face regions, face / corner bindings --
// 280 MB on the 10M-vertex mesh) are then neither uploaded nor ranked
static int dmesh_upload_impl(hb_ctx *ctx, const hb_mesh_desc *descs, uint32_t nseg, hb_dmesh *m, bool vertex_only = false, bool for_encoder = false)
{
	if (nseg == 0) return hb_fail(ctx, HB_ERR_INVALID, "batch: no mesh");
	const hb_mesh_desc *d = &descs[0];
	m->ctx = ctx;
	m->nseg = nseg;
	m->alloc_on_copy_stream = m->async_copy;
	m->gather_uploads = m->async_copy && nseg > 1; // (a single mesh is a dozen large copies: nothing to gain)
	m->upload_mode = 2;
	if (const char *env = getenv("HARRY_B200_GATHER_UPLOADS")) { // A/B runs and tests: 0 = one copy per array, 1 = gather kernel, 2 = batched copies
		m->upload_mode = atoi(env);
		m->gather_uploads = m->async_copy && m->upload_mode != 0;
	}
	// (a mesh without faces says nothing about the face order: its pointer may be anything)
	bool any_order_f = false;
	for (uint32_t s = 0; s < nseg; ++s) any_order_f = any_order_f || (descs[s].nf && descs[s].order_f != nullptr);
	m->has_order_f = any_order_f && !vertex_only;
	m->nb_face = d->nb_face; m->nb_vtx = d->nb_vtx; m->nb_corner = d->nb_corner;
	m->nregs_face = d->nregs_face; m->nregs_vtx = d->nregs_vtx; m->nlists = d->nlists;
	if (d->nlists && !d->lists) return hb_fail(ctx, HB_ERR_INVALID, "mesh: lists == NULL");
	// ---- segment bases -----------------------------------------------------------------------------
	m->h_vbase.assign((size_t)nseg + 1, 0); m->h_fbase.assign((size_t)nseg + 1, 0); m->h_ebase.assign((size_t)nseg + 1, 0);
	m->h_obase.assign((size_t)nseg + 1, 0); m->h_ofbase.assign((size_t)nseg + 1, 0);
	uint64_t av = 0, af = 0, ae = 0, ao = 0, aof = 0;
	for (uint32_t s = 0; s < nseg; ++s) {
		const hb_mesh_desc &ds = descs[s];
		if (ds.ne && (!ds.edges || !ds.face_off)) return hb_fail(ctx, HB_ERR_INVALID, "mesh: edges / face_off == NULL");
		if (ds.nf && ds.face_off[ds.nf] != ds.ne) return hb_fail(ctx, HB_ERR_INVALID, "mesh: face_off[nf] != ne");
		if (ds.nf && ds.face_off[0] != 0) return hb_fail(ctx, HB_ERR_INVALID, "mesh: face_off[0] != 0");
		if (ds.norder && !ds.order) return hb_fail(ctx, HB_ERR_INVALID, "mesh: order == NULL");
		if ((ds.nv && !ds.vtx_regs) || (ds.nf && !ds.face_regs)) return hb_fail(ctx, HB_ERR_INVALID, "mesh: region arrays == NULL");
		if (ds.nf && (ds.order_f != nullptr) != any_order_f) return hb_fail(ctx, HB_ERR_INVALID, "batch: a face order for some meshes only (mesh %u)", s);
		if (s) {
			if (ds.nlists != d->nlists || ds.nb_face != d->nb_face || ds.nb_vtx != d->nb_vtx || ds.nb_corner != d->nb_corner || ds.nregs_face != d->nregs_face ||
			    ds.nregs_vtx != d->nregs_vtx || (ds.emit_type != nullptr) != (d->emit_type != nullptr) ||
			    !same_region_table(d->off_reg_vtx, d->reg_vtxlist, ds.off_reg_vtx, ds.reg_vtxlist, d->nregs_vtx) ||
			    !same_region_table(d->off_reg_face, d->reg_facelist, ds.off_reg_face, ds.reg_facelist, d->nregs_face) ||
			    !same_region_table(d->off_reg_corner, d->reg_cornerlist, ds.off_reg_corner, ds.reg_cornerlist, d->nregs_face))
				return hb_fail(ctx, HB_ERR_INVALID, "batch: mesh %u has another schema (lists, regions, bindings) than mesh 0", s);
		}
		m->h_vbase[s] = (uint32_t)av; m->h_fbase[s] = (uint32_t)af; m->h_ebase[s] = (uint32_t)ae; m->h_obase[s] = (uint32_t)ao; m->h_ofbase[s] = (uint32_t)aof;
		av += ds.nv; af += ds.nf; ae += ds.ne; ao += ds.norder;
		aof += vertex_only ? 0 : (ds.order_f ? ds.norder_f : ds.nf);
		if (av >= 0x7fffffffull || af + nseg >= 0x7fffffffull || ae >= 0x7fffffffull || ao >= 0x7fffffffull)
			return hb_fail(ctx, HB_ERR_UNSUPPORTED, "batch: more than 2^31 vertices / faces / half-edges (split the batch)");
	}
	m->h_vbase[nseg] = (uint32_t)av; m->h_fbase[nseg] = (uint32_t)af; m->h_ebase[nseg] = (uint32_t)ae; m->h_obase[nseg] = (uint32_t)ao; m->h_ofbase[nseg] = (uint32_t)aof;
	m->nv = (uint32_t)av; m->nf = (uint32_t)af; m->ne = (uint32_t)ae; m->norder = (uint32_t)ao; m->norder_f = (uint32_t)aof;
	{
		std::vector<uint32_t> tab;
		tab.reserve(5 * ((size_t)nseg + 1));
		for (const std::vector<uint32_t> *v : { &m->h_vbase, &m->h_fbase, &m->h_ebase, &m->h_obase, &m->h_ofbase }) tab.insert(tab.end(), v->begin(), v->end());
		HB_TRY(upload(m, (void **)&m->d_segtab, tab.data(), sizeof(uint32_t) * tab.size()));
		const size_t w = (size_t)nseg + 1;
		m->d_vbase = m->d_segtab; m->d_fbase = m->d_segtab + w; m->d_ebase = m->d_segtab + 2 * w; m->d_obase = m->d_segtab + 3 * w; m->d_ofbase = m->d_segtab + 4 * w;
	}
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
	cudaStream_t up_stream = m->async_copy ? ctx->copy_stream : ctx->stream;
	// ---- encoder upload whose face order may be unneeded: bindings first, the check right behind them -------------
	bool late = false;
	if (for_encoder && !vertex_only && m->async_copy && order_f_maybe_unneeded(descs, nseg)) {
		late = true;
		HB_TRY(hb_dalloc(m, (void **)&m->d_face_regs, sizeof(uint16_t) * (size_t)m->nf));
		HB_TRY(hb_dalloc(m, (void **)&m->d_bind_face, sizeof(uint32_t) * (size_t)m->nf * d->nb_face));
		HB_TRY(hb_dalloc(m, (void **)&m->d_bind_corner, sizeof(uint32_t) * (size_t)m->ne * d->nb_corner));
		for (uint32_t s = 0; s < nseg; ++s) {
			const hb_mesh_desc &ds = descs[s];
			HB_TRY(copy_in(m, m->d_bind_face + (size_t)m->h_fbase[s] * d->nb_face, ds.bind_face_attr, sizeof(uint32_t) * (size_t)ds.nf * d->nb_face));
			HB_TRY(copy_in(m, m->d_bind_corner + (size_t)m->h_ebase[s] * d->nb_corner, ds.bind_corner_attr, sizeof(uint32_t) * (size_t)ds.ne * d->nb_corner));
		}
		HB_TRY(flush_uploads(m));
		uint32_t *d_twice = nullptr;
		HB_TRY(hb_dalloc(m, (void **)&d_twice, sizeof(uint32_t)));
		HB_CUDA(ctx, cudaMemsetAsync(d_twice, 0, sizeof(uint32_t), up_stream));
		for (int l = 0; l < d->nlists; ++l) {
			const int cls = d->lists[l].target;
			if (cls != HB_FACE && cls != HB_CORNER) continue;
			const int slot = cls == HB_FACE ? m->h_slot_face[(size_t)l] : m->h_slot_corner[(size_t)l]; // region 0 is the only one
			const uint32_t nelem = cls == HB_FACE ? m->nf : m->ne;
			if (slot < 0 || nelem == 0) continue;
			std::vector<uint32_t> bitbase((size_t)nseg + 1, 0);
			for (uint32_t s = 0; s < nseg; ++s) bitbase[s + 1] = bitbase[s] + (descs[s].lists[l].nrows + 31u) / 32u;
			uint32_t *d_bitbase = nullptr, *d_bits = nullptr;
			HB_TRY(upload(m, (void **)&d_bitbase, bitbase.data(), sizeof(uint32_t) * bitbase.size()));
			if (bitbase[nseg]) {
				HB_TRY(hb_dalloc(m, (void **)&d_bits, sizeof(uint32_t) * (size_t)bitbase[nseg]));
				HB_CUDA(ctx, cudaMemsetAsync(d_bits, 0, sizeof(uint32_t) * (size_t)bitbase[nseg], up_stream));
			}
			const uint32_t grid = std::min<uint32_t>(hb_div_up(nelem, 256), (uint32_t)ctx->sm_count * 8u);
			k_bind_twice<<<grid, 256, 0, up_stream>>>(cls == HB_FACE ? m->d_bind_face : m->d_bind_corner, cls == HB_FACE ? m->d_fbase : m->d_ebase, nseg,
			                                          cls == HB_FACE ? d->nb_face : d->nb_corner, (uint32_t)slot, d_bitbase, d_bits, d_twice);
			ctx->launches++;
			HB_CUDA(ctx, cudaGetLastError());
		}
		HB_CUDA(ctx, cudaMemcpyAsync(ctx->h_flag, d_twice, sizeof(uint32_t), cudaMemcpyDeviceToHost, up_stream));
		HB_CUDA(ctx, cudaEventRecord(ctx->ev_flag, up_stream));
	}
	// ---- allocate, then copy segment by segment ----------------------------------------------------------
	const bool up_faces = !vertex_only;
	HB_TRY(hb_dalloc(m, (void **)&m->d_edges_raw, 12 * (size_t)m->ne));
	if (nseg == 1) HB_TRY(hb_dalloc(m, (void **)&m->d_face_off, sizeof(uint32_t) * ((size_t)m->nf + 1)));
	else HB_TRY(hb_dalloc(m, (void **)&m->d_face_off_raw, sizeof(uint32_t) * ((size_t)m->nf + nseg)));
	HB_TRY(hb_dalloc(m, (void **)&m->d_order, 8 * (size_t)m->norder));
	if (m->has_order_f && !late) HB_TRY(hb_dalloc(m, (void **)&m->d_order_f, 8 * (size_t)m->norder_f));
	HB_TRY(hb_dalloc(m, (void **)&m->d_vtx_regs, sizeof(uint16_t) * (size_t)m->nv));
	uint32_t *face_off_dst = nseg == 1 ? m->d_face_off : m->d_face_off_raw;
	for (uint32_t s = 0; s < nseg; ++s) {
		const hb_mesh_desc &ds = descs[s];
		HB_TRY(copy_in(m, m->d_edges_raw + 12 * (size_t)m->h_ebase[s], ds.edges, 12 * (size_t)ds.ne));
		HB_TRY(copy_in(m, face_off_dst + m->h_fbase[s] + s, ds.face_off, sizeof(uint32_t) * ((size_t)ds.nf + 1)));
		HB_TRY(copy_in(m, m->d_order + 8 * (size_t)m->h_obase[s], ds.order, 8 * (size_t)ds.norder));
		if (m->has_order_f && !late) HB_TRY(copy_in(m, m->d_order_f + 8 * (size_t)m->h_ofbase[s], ds.order_f, 8 * (size_t)ds.norder_f));
		if (d->nregs_vtx > 1) HB_TRY(copy_in(m, m->d_vtx_regs + m->h_vbase[s], ds.vtx_regs, sizeof(uint16_t) * (size_t)ds.nv));
	}
	// a single region: every entry is 0 by definition -- cleared on the device instead of uploaded
	if (d->nregs_vtx <= 1 && m->nv) HB_CUDA(ctx, cudaMemsetAsync(m->d_vtx_regs, 0, sizeof(uint16_t) * (size_t)m->nv, up_stream));
	// everything K0 / K3 / K4 read is on its way: first milestone of the copy stream
	HB_TRY(flush_uploads(m));
	if (m->async_copy) {
		HB_CUDA(ctx, cudaEventCreateWithFlags(&m->ev_up[0], cudaEventDisableTiming));
		HB_CUDA(ctx, cudaEventCreateWithFlags(&m->ev_up[1], cudaEventDisableTiming));
		HB_CUDA(ctx, cudaEventRecord(m->ev_up[0], ctx->copy_stream));
	}
	if (up_faces && !late) {
		HB_TRY(hb_dalloc(m, (void **)&m->d_face_regs, sizeof(uint16_t) * (size_t)m->nf));
		HB_TRY(hb_dalloc(m, (void **)&m->d_bind_face, sizeof(uint32_t) * (size_t)m->nf * d->nb_face));
		HB_TRY(hb_dalloc(m, (void **)&m->d_bind_corner, sizeof(uint32_t) * (size_t)m->ne * d->nb_corner));
	}
	HB_TRY(hb_dalloc(m, (void **)&m->d_bind_vtx, sizeof(uint32_t) * (size_t)m->nv * d->nb_vtx));
	for (uint32_t s = 0; s < nseg; ++s) {
		const hb_mesh_desc &ds = descs[s];
		if (up_faces && !late) {
			if (d->nregs_face > 1) HB_TRY(copy_in(m, m->d_face_regs + m->h_fbase[s], ds.face_regs, sizeof(uint16_t) * (size_t)ds.nf));
			HB_TRY(copy_in(m, m->d_bind_face + (size_t)m->h_fbase[s] * d->nb_face, ds.bind_face_attr, sizeof(uint32_t) * (size_t)ds.nf * d->nb_face));
			HB_TRY(copy_in(m, m->d_bind_corner + (size_t)m->h_ebase[s] * d->nb_corner, ds.bind_corner_attr, sizeof(uint32_t) * (size_t)ds.ne * d->nb_corner));
		}
		HB_TRY(copy_in(m, m->d_bind_vtx + (size_t)m->h_vbase[s] * d->nb_vtx, ds.bind_vtx_attr, sizeof(uint32_t) * (size_t)ds.nv * d->nb_vtx));
	}
	if (up_faces && d->nregs_face <= 1 && m->nf) HB_CUDA(ctx, cudaMemsetAsync(m->d_face_regs, 0, sizeof(uint16_t) * (size_t)m->nf, up_stream));
	if (vertex_only) m->any_corner = false;
	HB_TRY(upload(m, (void **)&m->d_slot_vtx, m->h_slot_vtx.data(), sizeof(int16_t) * m->h_slot_vtx.size()));
	HB_TRY(upload(m, (void **)&m->d_slot_face, m->h_slot_face.data(), sizeof(int16_t) * m->h_slot_face.size()));
	HB_TRY(upload(m, (void **)&m->d_slot_corner, m->h_slot_corner.data(), sizeof(int16_t) * m->h_slot_corner.size()));
	for (int l = 0; l < d->nlists; ++l) {
		HB_TRY(add_list(m, descs, nullptr, nseg, l, true));
		if (d->emit_type) {
			// drained type symbols: all segments or none, concatenated in segment order
			DevList &dl = m->lists.back();
			bool any = false, all = true;
			uint64_t total = 0;
			dl.h_emitbase.assign((size_t)nseg + 1, 0);
			for (uint32_t s = 0; s < nseg; ++s) {
				dl.h_emitbase[s] = (uint32_t)total;
				if (descs[s].nv == 0 && descs[s].nf == 0) continue; // an empty mesh emits nothing: neither given nor missing
				const bool has = descs[s].emit_type && descs[s].emit_type[l];
				any = any || has; all = all && has;
				if (has && !descs[s].emit_count) return hb_fail(ctx, HB_ERR_INVALID, "mesh: emit_type without emit_count");
				if (has) total += descs[s].emit_count[l];
			}
			dl.h_emitbase[nseg] = (uint32_t)total;
			if (any && !all) return hb_fail(ctx, HB_ERR_INVALID, "batch: emit_type[%d] given for some meshes only", l);
			if (all) {
				dl.emit_count = (uint32_t)total;
				HB_TRY(hb_dalloc(m, (void **)&dl.d_emit_type, total));
				for (uint32_t s = 0; s < nseg; ++s) {
					if (descs[s].nv == 0 && descs[s].nf == 0) continue;
					HB_TRY(copy_in(m, dl.d_emit_type + dl.h_emitbase[s], descs[s].emit_type[l], descs[s].emit_count[l]));
				}
			}
		}
	}
	HB_TRY(flush_uploads(m));
	if (late) {
		// the answer has been back for a while: the link is still busy with what was queued behind the check
		HB_CUDA(ctx, cudaEventSynchronize(ctx->ev_flag));
		if (*ctx->h_flag) {
			HB_TRY(hb_dalloc(m, (void **)&m->d_order_f, 8 * (size_t)m->norder_f));
			for (uint32_t s = 0; s < nseg; ++s) HB_TRY(copy_in(m, m->d_order_f + 8 * (size_t)m->h_ofbase[s], descs[s].order_f, 8 * (size_t)descs[s].norder_f));
			m->order_f_late = true;
		} else {
			m->has_order_f = false; // index order, gate corner 0 (norder_f == nf was required: the segment tables are the same)
		}
	}
	HB_TRY(flush_uploads(m));
	m->gather_uploads = false;
	if (m->async_copy) HB_CUDA(ctx, cudaEventRecord(m->ev_up[1], ctx->copy_stream));
	m->alloc_on_copy_stream = false;
	return 0;
}

// host-buffer entry points: connectivity stages under the tail of the upload, everything else behind it
static int upload_overlapped(hb_ctx *ctx, const hb_mesh_desc *meshes, uint32_t nseg, hb_dmesh *m, bool vertex_only, bool for_encoder)
{
	m->async_copy = true;
	HB_TRY(dmesh_upload_impl(ctx, meshes, nseg, m, vertex_only, for_encoder));
	HB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, m->ev_up[m->order_f_late ? 1 : 0], 0));
	HB_TRY(hb_build_conn(m));
	bool need_v = false;
	for (int l = 0; l < m->nlists; ++l)
		if (m->lists[l].p.target == HB_VTX && m->lists[l].p.ncomp) need_v = true;
	if (need_v) HB_TRY(hb_build_vertex_candidates(m));
	HB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, m->ev_up[1], 0));
	return 0;
}

extern "C" int hb_dmesh_upload_batch(hb_ctx *ctx, const hb_mesh_desc *meshes, uint32_t n, hb_dmesh **out)
{
	*out = nullptr;
	HB_CUDA(ctx, cudaSetDevice(ctx->device));
	hb_dmesh *m = new hb_dmesh();
	int rc = dmesh_upload_impl(ctx, meshes, n, m);
	if (rc == 0) {
		cudaError_t e = cudaStreamSynchronize(ctx->stream); // the host buffers may be reused after return
		if (e != cudaSuccess) rc = hb_fail(ctx, HB_ERR_CUDA, "upload failed: %s", cudaGetErrorString(e));
	}
	if (rc) { hb_dmesh_free(m); return rc; }
	*out = m;
	return 0;
}
extern "C" int hb_dmesh_upload(hb_ctx *ctx, const hb_mesh_desc *d, hb_dmesh **out) { return hb_dmesh_upload_batch(ctx, d, 1, out); }
extern "C" uint32_t hb_dmesh_segments(hb_dmesh *m) { return m->nseg; }

extern "C" int hb_dmesh_quantize(hb_dmesh *m, uint32_t l, const uint8_t *new_quant, const uint8_t *groups)
{
	if (l >= m->lists.size()) return hb_fail(m->ctx, HB_ERR_INVALID, "list %u out of range", l);
	HB_CUDA(m->ctx, cudaSetDevice(m->ctx->device));
	HB_TRY(hb_list_bounds(m, l, groups)); // min, max and (tail of the same kernel) scale rows of every segment
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
// the scale from set_scale); a batch passes the rows of its meshes back to back (nseg * stride bytes each)
extern "C" int hb_dmesh_set_bounds(hb_dmesh *m, uint32_t l, const void *min_row, const void *max_row, const void *scale_row)
{
	if (l >= m->lists.size()) return hb_fail(m->ctx, HB_ERR_INVALID, "list %u out of range", l);
	DevList &dl = m->lists[l];
	const size_t s = dl.p.stride;
	hb_ctx *ctx = m->ctx;
	HB_CUDA(ctx, cudaSetDevice(ctx->device));
	const void *src[3] = { min_row, max_row, scale_row };
	for (int w = 0; w < 3; ++w)
		if (src[w] && s) HB_CUDA(ctx, cudaMemcpy2DAsync(dl.d_bounds + w * s, dl.bounds_pitch, src[w], s, s, m->nseg, cudaMemcpyHostToDevice, ctx->stream));
	HB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return 0;
}

extern "C" int hb_dmesh_fetch_bounds(hb_dmesh *m, uint32_t l, void *min_row, void *max_row, void *scale_row)
{
	if (l >= m->lists.size()) return hb_fail(m->ctx, HB_ERR_INVALID, "list %u out of range", l);
	DevList &dl = m->lists[l];
	const size_t s = dl.p.stride;
	hb_ctx *ctx = m->ctx;
	HB_CUDA(ctx, cudaSetDevice(ctx->device));
	void *dst[3] = { min_row, max_row, scale_row };
	for (int w = 0; w < 3; ++w)
		if (dst[w] && s) HB_CUDA(ctx, cudaMemcpy2DAsync(dst[w], s, dl.d_bounds + w * s, dl.bounds_pitch, s, m->nseg, cudaMemcpyDeviceToHost, ctx->stream));
	HB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return hb_check_device_error(ctx, "bounds");
}

// rows of list l of segment `seg` (nrows_seg * stride bytes)
extern "C" int hb_dmesh_fetch_rows_seg(hb_dmesh *m, uint32_t seg, uint32_t l, void *rows_out)
{
	if (l >= m->lists.size() || seg >= m->nseg) return hb_fail(m->ctx, HB_ERR_INVALID, "list %u / mesh %u out of range", l, seg);
	DevList &dl = m->lists[l];
	HB_CUDA(m->ctx, cudaSetDevice(m->ctx->device));
	const size_t bytes = (size_t)dl.h_rownum[seg] * dl.p.stride;
	if (bytes) HB_CUDA(m->ctx, cudaMemcpyAsync(rows_out, dl.p.rows + (size_t)dl.h_rowbase[seg] * dl.p.stride, bytes, cudaMemcpyDeviceToHost, m->ctx->stream));
	HB_CUDA(m->ctx, cudaStreamSynchronize(m->ctx->stream));
	return hb_check_device_error(m->ctx, "fetch rows");
}
extern "C" int hb_dmesh_fetch_rows(hb_dmesh *m, uint32_t l, void *rows_out) { return hb_dmesh_fetch_rows_seg(m, 0, l, rows_out); }

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

// per segment: emission and DATA-row offsets of its first element (entries nseg: the totals)
__global__ void k_segment_offsets(const uint32_t *__restrict__ elem_base, uint32_t nseg, const uint32_t *__restrict__ ek, const uint32_t *__restrict__ dord, const uint32_t *__restrict__ dup,
                                  uint32_t *__restrict__ out)
{
	const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s > nseg) return;
	const uint32_t e = elem_base[s];
	out[s] = ek ? ek[e] : e;
	out[nseg + 1 + s] = dup && *dup == 0u ? out[s] : dord[e]; // no row is shared: every emission is a DATA row
}

// ------------------------------------------------------------------------------------------------
// Streams of a device mesh -> page-locked host memory.  One block per (list, stream) for ALL segments; the
// hb_streams of segment s are views into those blocks (hb_batch_streams owns them).  Two phases so that a
// pipelined batch can queue the copies of group g and collect them after it has queued the work of group g + 1:
//   begin   small read-back of the per-segment offsets (one synchronisation of `st`), then the sized copies
//   finish  (after `st` was synchronised) fills the per-mesh views
// ------------------------------------------------------------------------------------------------
struct hb_batch_streams_priv {
	std::vector<void *> blocks;   // page-locked blocks (g_pinned)
	std::vector<hb_list_streams *> list_arrays;
};
struct FetchState {
	hb_dmesh *m = nullptr;
	uint32_t seg0 = 0;             // index of the group's first mesh in the caller's batch
	std::vector<uint32_t *> offs;  // per list: page-locked [2 * (nseg + 1)] emission / DATA offsets
	std::vector<uint8_t *> type;
	std::vector<uint32_t *> aux;
	std::vector<uint8_t *> sym;
	std::vector<uint64_t *> hist;
	uint16_t *reg_vtx = nullptr, *reg_face = nullptr;
	uint16_t *d_reg = nullptr;
};

static void *priv_alloc(hb_batch_streams_priv *pv, size_t bytes)
{
	void *p = g_pinned.alloc(bytes);
	if (p) pv->blocks.push_back(p);
	return p;
}

extern "C" void hb_batch_streams_free(hb_batch_streams *b)
{
	if (!b) return;
	hb_batch_streams_priv *pv = (hb_batch_streams_priv *)b->priv;
	if (pv) {
		for (void *p : pv->blocks) g_pinned.release(p);
		for (hb_list_streams *l : pv->list_arrays) free(l);
		delete pv;
	}
	free(b->mesh);
	free(b);
}

#define FETCH_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return hb_fail(ctx, HB_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e__)); } while (0)
static int fetch_begin(hb_dmesh *m, cudaStream_t st, hb_batch_streams_priv *pv, FetchState &fs)
{
	hb_ctx *ctx = m->ctx;
	if (!m->encoded) return hb_fail(ctx, HB_ERR_INVALID, "fetch_streams before encode");
	const uint32_t nseg = m->nseg;
	fs.m = m;
	const bool reg_v = m->nregs_vtx > 1, reg_f = m->nregs_face > 1;
	// region symbol streams (io.h:109-116): region of every traversed vertex / face; a mesh with a single vertex (face)
	// region has an all-zero stream: NULL stands for it
	if (reg_v || reg_f) {
		const uint32_t nmax = m->norder > m->norder_f ? m->norder : m->norder_f;
		FETCH_CUDA(cudaMallocAsync((void **)&fs.d_reg, sizeof(uint16_t) * (2 * (size_t)nmax + 2), st));
		if (m->norder && reg_v) {
			fs.reg_vtx = (uint16_t *)priv_alloc(pv, sizeof(uint16_t) * ((size_t)m->norder + 1));
			if (!fs.reg_vtx) return hb_fail(ctx, HB_ERR_NOMEM, "out of host memory");
			k_region_stream<<<hb_div_up(m->norder, 256), 256, 0, st>>>(m->d_ord_v, m->d_he, m->d_vtx_regs, m->norder, 0, fs.d_reg);
			ctx->launches++;
			FETCH_CUDA(cudaMemcpyAsync(fs.reg_vtx, fs.d_reg, sizeof(uint16_t) * m->norder, cudaMemcpyDeviceToHost, st));
		}
		if (m->norder_f && reg_f) {
			fs.reg_face = (uint16_t *)priv_alloc(pv, sizeof(uint16_t) * ((size_t)m->norder_f + 1));
			if (!fs.reg_face) return hb_fail(ctx, HB_ERR_NOMEM, "out of host memory");
			k_region_stream<<<hb_div_up(m->norder_f, 256), 256, 0, st>>>(m->d_ford_h, m->d_he, m->d_face_regs, m->norder_f, 1, fs.d_reg + nmax + 1);
			ctx->launches++;
			FETCH_CUDA(cudaMemcpyAsync(fs.reg_face, fs.d_reg + nmax + 1, sizeof(uint16_t) * m->norder_f, cudaMemcpyDeviceToHost, st));
		}
		FETCH_CUDA(cudaFreeAsync(fs.d_reg, st));
	}
	const int nl = m->nlists;
	fs.offs.assign(nl, nullptr); fs.type.assign(nl, nullptr); fs.aux.assign(nl, nullptr); fs.sym.assign(nl, nullptr); fs.hist.assign(nl, nullptr);
	// per-segment offsets and histograms (with the type counters) first: they size the other copies
	std::vector<uint32_t *> d_offs(nl, nullptr);
	for (int l = 0; l < nl; ++l) {
		DevList &dl = m->lists[l];
		if (!dl.n_elems && !dl.d_hist) continue; // list not coded (no class / no corner bindings)
		const size_t hist_pitch = (size_t)dl.p.sym_stride * 256 + 4;
		fs.offs[l] = (uint32_t *)priv_alloc(pv, sizeof(uint32_t) * 2 * ((size_t)nseg + 1));
		fs.hist[l] = (uint64_t *)priv_alloc(pv, sizeof(uint64_t) * hist_pitch * nseg);
		if (!fs.offs[l] || !fs.hist[l]) return hb_fail(ctx, HB_ERR_NOMEM, "out of host memory");
		const int cls = dl.p.target;
		const uint32_t *elem_base = cls == HB_VTX ? m->d_obase : cls == HB_FACE ? m->d_ofbase : m->d_cebase;
		if (dl.nocomp_fast) { // offsets follow from the segment tables and the type counters (filled in below)
			FETCH_CUDA(cudaMemcpyAsync(fs.hist[l], dl.d_hist, sizeof(uint64_t) * hist_pitch * nseg, cudaMemcpyDeviceToHost, st));
			continue;
		}
		FETCH_CUDA(cudaMallocAsync((void **)&d_offs[l], sizeof(uint32_t) * 2 * ((size_t)nseg + 1), st));
		if (dl.n_elems) {
			k_segment_offsets<<<hb_div_up(nseg + 1, 128), 128, 0, st>>>(elem_base, nseg, dl.d_ek, dl.d_dord, dl.d_dup_active ? dl.d_dup : nullptr, d_offs[l]);
			ctx->launches++;
		} else {
			FETCH_CUDA(cudaMemsetAsync(d_offs[l], 0, sizeof(uint32_t) * 2 * ((size_t)nseg + 1), st));
		}
		FETCH_CUDA(cudaMemcpyAsync(fs.offs[l], d_offs[l], sizeof(uint32_t) * 2 * ((size_t)nseg + 1), cudaMemcpyDeviceToHost, st));
		FETCH_CUDA(cudaMemcpyAsync(fs.hist[l], dl.d_hist, sizeof(uint64_t) * hist_pitch * nseg, cudaMemcpyDeviceToHost, st));
		FETCH_CUDA(cudaFreeAsync(d_offs[l], st));
	}
	FETCH_CUDA(cudaStreamSynchronize(st));
	for (int l = 0; l < nl; ++l) {
		DevList &dl = m->lists[l];
		if (!fs.offs[l]) continue;
		const size_t hist_pitch = (size_t)dl.p.sym_stride * 256 + 4;
		if (dl.nocomp_fast) {
			const std::vector<uint32_t> &eb = dl.p.target == HB_VTX ? m->h_obase : m->h_ofbase;
			uint32_t nd = 0;
			bool any_hist = false;
			for (uint32_t sg = 0; sg <= nseg; ++sg) {
				fs.offs[l][sg] = eb[sg];
				fs.offs[l][nseg + 1 + sg] = nd;
				if (sg < nseg) {
					nd += (uint32_t)fs.hist[l][hist_pitch * sg + HB_DATA];
					any_hist = any_hist || fs.hist[l][hist_pitch * sg + HB_HIST] != 0;
				}
			}
			if (any_hist) HB_TRY(hb_nocomp_finish(m, l, st)); // shared rows in a list without components: history offsets after all
		}
		const uint32_t n_emit = fs.offs[l][nseg], n_data = fs.offs[l][2 * nseg + 1];
		// every emission of every segment is a DATA row (no shared attribute rows): the type and history-offset
		// streams are all zero -- NULL stands for them, nothing is copied
		bool all_data = true;
		for (uint32_t sg = 0; sg < nseg; ++sg) {
			const uint64_t *th = fs.hist[l] + hist_pitch * sg + (size_t)dl.p.sym_stride * 256;
			all_data = all_data && th[HB_HIST] + th[HB_LHIST] == 0;
		}
		if (!all_data && n_emit) {
			fs.type[l] = (uint8_t *)priv_alloc(pv, (size_t)n_emit + 1);
			fs.aux[l] = (uint32_t *)priv_alloc(pv, sizeof(uint32_t) * ((size_t)n_emit + 1));
			if (!fs.type[l] || !fs.aux[l]) return hb_fail(ctx, HB_ERR_NOMEM, "out of host memory");
			FETCH_CUDA(cudaMemcpyAsync(fs.type[l], dl.d_type, n_emit, cudaMemcpyDeviceToHost, st));
			FETCH_CUDA(cudaMemcpyAsync(fs.aux[l], dl.d_aux, sizeof(uint32_t) * n_emit, cudaMemcpyDeviceToHost, st));
		}
		const size_t sym_bytes = (size_t)n_data * dl.p.sym_stride;
		fs.sym[l] = (uint8_t *)priv_alloc(pv, sym_bytes + 1);
		if (!fs.sym[l]) return hb_fail(ctx, HB_ERR_NOMEM, "out of host memory");
		if (sym_bytes) FETCH_CUDA(cudaMemcpyAsync(fs.sym[l], dl.d_sym, sym_bytes, cudaMemcpyDeviceToHost, st));
	}
	return 0;
}

// views of segment sg (after `st` has been synchronised)
static void fetch_view(const FetchState &fs, hb_batch_streams_priv *pv, uint32_t sg, hb_streams *s)
{
	hb_dmesh *m = fs.m;
	const uint32_t nseg = m->nseg;
	memset(s, 0, sizeof *s);
	s->n_vtx = m->h_obase[sg + 1] - m->h_obase[sg];
	s->n_face = m->h_ofbase[sg + 1] - m->h_ofbase[sg];
	s->nlists = m->nlists;
	s->reg_vtx = fs.reg_vtx ? fs.reg_vtx + m->h_obase[sg] : nullptr;
	s->reg_face = fs.reg_face ? fs.reg_face + m->h_ofbase[sg] : nullptr;
	s->lists = (hb_list_streams *)calloc((size_t)m->nlists + 1, sizeof(hb_list_streams));
	pv->list_arrays.push_back(s->lists);
	for (int l = 0; l < m->nlists; ++l) {
		DevList &dl = m->lists[l];
		hb_list_streams &ls = s->lists[l];
		ls.sym_stride = dl.p.sym_stride;
		if (!fs.offs[l]) continue;
		const size_t hist_pitch = (size_t)dl.p.sym_stride * 256 + 4;
		const uint32_t e0 = fs.offs[l][sg], e1 = fs.offs[l][sg + 1], d0 = fs.offs[l][nseg + 1 + sg], d1 = fs.offs[l][nseg + 1 + sg + 1];
		ls.n_emit = e1 - e0;
		ls.n_data = d1 - d0;
		ls.hist = fs.hist[l] + hist_pitch * sg;
		for (int k = 0; k < 4; ++k) ls.type_hist[k] = ls.hist[(size_t)ls.sym_stride * 256 + k];
		const bool all_data = ls.type_hist[HB_HIST] + ls.type_hist[HB_LHIST] == 0;
		ls.type = (all_data || !fs.type[l]) ? nullptr : fs.type[l] + e0;
		ls.aux = (all_data || !fs.aux[l]) ? nullptr : fs.aux[l] + e0;
		ls.symbols = fs.sym[l] + (size_t)d0 * ls.sym_stride;
	}
}
#undef FETCH_CUDA

extern "C" int hb_dmesh_fetch_streams_batch(hb_dmesh *m, hb_batch_streams **out)
{
	hb_ctx *ctx = m->ctx;
	*out = nullptr;
	HB_CUDA(ctx, cudaSetDevice(ctx->device));
	hb_batch_streams *b = (hb_batch_streams *)calloc(1, sizeof(hb_batch_streams));
	hb_batch_streams_priv *pv = new hb_batch_streams_priv();
	if (!b) { delete pv; return hb_fail(ctx, HB_ERR_NOMEM, "out of host memory"); }
	b->priv = pv;
	b->n = m->nseg;
	b->mesh = (hb_streams *)calloc(m->nseg, sizeof(hb_streams));
	FetchState fs;
	int rc = b->mesh ? fetch_begin(m, ctx->stream, pv, fs) : hb_fail(ctx, HB_ERR_NOMEM, "out of host memory");
	if (rc == 0 && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = hb_fail(ctx, HB_ERR_CUDA, "stream download failed: %s", cudaGetErrorString(cudaGetLastError()));
	if (rc == 0) rc = hb_check_device_error(ctx, "attribute encode");
	if (rc) { hb_batch_streams_free(b); return rc; }
	for (uint32_t sg = 0; sg < m->nseg; ++sg) fetch_view(fs, pv, sg, &b->mesh[sg]);
	*out = b;
	return 0;
}

// single mesh: the same download, handed out as a standalone hb_streams (every array its own page-locked block,
// released one by one by hb_streams_free)
extern "C" int hb_dmesh_fetch_streams(hb_dmesh *m, hb_streams **out)
{
	*out = nullptr;
	if (m->nseg != 1) return hb_fail(m->ctx, HB_ERR_INVALID, "fetch_streams on a batch: use hb_dmesh_fetch_streams_batch");
	hb_batch_streams *b = nullptr;
	HB_TRY(hb_dmesh_fetch_streams_batch(m, &b));
	hb_batch_streams_priv *pv = (hb_batch_streams_priv *)b->priv;
	hb_streams *s = (hb_streams *)calloc(1, sizeof(hb_streams));
	*s = b->mesh[0];
	// blocks that are not handed out (the offset tables) go back to the cache now; the others are released by
	// hb_streams_free through the pointers of the hb_streams (all views start at offset 0 for a single segment)
	for (void *p : pv->blocks) {
		bool used = p == s->reg_vtx || p == s->reg_face;
		for (int l = 0; l < s->nlists; ++l) used = used || p == s->lists[l].type || p == s->lists[l].aux || p == s->lists[l].symbols || p == s->lists[l].hist;
		if (!used) g_pinned.release(p);
	}
	delete pv;
	free(b->mesh);
	free(b);
	*out = s;
	return 0;
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
	m->nseg = 1;
	return add_list(m, nullptr, list, 1, 0, false);
}

extern "C" int hb_bounds(hb_ctx *ctx, const hb_list_desc *list, void *min_row, void *max_row)
{
	HB_CUDA(ctx, cudaSetDevice(ctx->device));
	hb_dmesh *m = new hb_dmesh();
	PhaseTimer t(ctx);
	t.mark(0);
	int rc = single_list_mesh(ctx, list, m);
	t.mark(1);
	if (rc == 0) rc = hb_list_bounds(m, 0, nullptr);
	t.mark(2);
	if (rc == 0 && list->ncomp) rc = hb_dmesh_fetch_bounds(m, 0, min_row, max_row, nullptr);
	t.mark(3);
	t.finish(4);
	if (rc == 0) row_cache_put(m, 0, list->rows); // requant of the same rows usually follows (quant.h:222-242)
	hb_dmesh_free(m);
	return rc;
}

extern "C" int hb_requant(hb_ctx *ctx, hb_list_desc *list, const uint8_t *new_quant, const void *min_row, const void *scale_row)
{
	HB_CUDA(ctx, cudaSetDevice(ctx->device));
	hb_dmesh *m = new hb_dmesh();
	PhaseTimer t(ctx);
	t.mark(0);
	m->take_cached_rows = true;
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
	if (rc == 0) row_cache_put(m, 0, list->rows); // the coder (or another requant) usually follows
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
	// the fill pass hands out one slot per corner: the offsets must be a monotone CSR with degrees the 16-bit local
	// edge index can hold, or the sum of the degrees would exceed ne (one pass over nf + 1 words)
	for (uint32_t f = 0; f < nf; ++f) {
		if (face_off[f + 1] < face_off[f]) return hb_fail(ctx, HB_ERR_INVALID, "twin_match: face_off is not monotone at face %u", f);
		if (face_off[f + 1] - face_off[f] > 0xffffu) return hb_fail(ctx, HB_ERR_INVALID, "twin_match: face %u has more than 65535 corners", f);
	}
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
	// the caller's records (possibly the origins themselves, in place) are only overwritten by a table that is valid
	if (rc == 0) rc = hb_check_device_error(ctx, "twin matching");
	if (rc == 0 && ne && cudaMemcpyAsync(edges_out, d_out, 12 * (size_t)ne, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
		rc = hb_fail(ctx, HB_ERR_CUDA, "twin_match: download failed: %s", cudaGetErrorString(cudaGetLastError()));
	t.mark(3);
	t.finish(4);
	hb_dmesh_free(m);
	return rc;
}

// ------------------------------------------------------------------------------------------------
// attribute coder on host buffers: single mesh = batch of one
// ------------------------------------------------------------------------------------------------
extern "C" int hb_attr_encode(hb_ctx *ctx, const hb_mesh_desc *mesh, hb_streams **out)
{
	*out = nullptr;
	HB_CUDA(ctx, cudaSetDevice(ctx->device));
	hb_dmesh *m = new hb_dmesh();
	m->take_cached_rows = true;
	PhaseTimer t(ctx);
	t.mark(0);
	int rc = upload_overlapped(ctx, mesh, 1, m, false, true);
	t.mark(1);
	if (rc == 0) rc = hb_encode_lists(m);
	t.mark(2);
	if (rc == 0) rc = hb_dmesh_fetch_streams(m, out);
	t.mark(3);
	t.finish(4);
	hb_dmesh_free(m);
	return rc;
}

static bool decode_is_vertex_only(const hb_mesh_desc *mesh)
{
	for (int l = 0; l < mesh->nlists; ++l)
		if (mesh->lists[l].target != HB_VTX && mesh->lists[l].ncomp && mesh->lists[l].nrows) return false;
	return true;
}

extern "C" int hb_attr_decode(hb_ctx *ctx, const hb_mesh_desc *mesh)
{
	HB_CUDA(ctx, cudaSetDevice(ctx->device));
	hb_dmesh *m = new hb_dmesh();
	PhaseTimer t(ctx);
	t.mark(0);
	int rc = upload_overlapped(ctx, mesh, 1, m, decode_is_vertex_only(mesh), false);
	t.mark(1);
	if (rc == 0) rc = hb_decode_lists(m);
	t.mark(2);
	for (int l = 0; rc == 0 && l < mesh->nlists; ++l)
		if (mesh->lists[l].ncomp && mesh->lists[l].nrows) rc = hb_dmesh_fetch_rows(m, (uint32_t)l, mesh->lists[l].rows);
	t.mark(3);
	t.finish(4);
	if (rc == 0) rc = hb_check_device_error(ctx, "attribute decode");
	for (int l = 0; rc == 0 && l < mesh->nlists; ++l) row_cache_put(m, l, mesh->lists[l].rows); // requant(clear) may follow (main.cc:104-108)
	hb_dmesh_free(m);
	return rc;
}

// ------------------------------------------------------------------------------------------------
// Batches of independent meshes on host buffers (BASELINE configs[4]; the reference runs one process per mesh,
// main.cc:93-122).  The batch is cut into groups of meshes (about 117M half-edges each, HARRY_B200_GROUP_HALF_EDGES); every group
// is one device mesh (one launch per stage over all its meshes).  Software pipeline over the groups:
//   copy stream     upload of group g + 1
//   compute stream  kernels of group g (wait for its upload only)
//   out stream      download of group g - 1
// The host waits exactly once per group (for the per-mesh stream sizes of the group before) and once at the end.
// ------------------------------------------------------------------------------------------------
static uint64_t group_half_edges()
{
	const char *env = getenv("HARRY_B200_GROUP_HALF_EDGES"); // read per call: tests force many small groups
	const uint64_t v = env ? strtoull(env, nullptr, 10) : 0;
	return v ? v : (uint64_t)112 << 20; // ~195 meshes of 100K vertices: 585 (mesh, component) chains = four per SM in the scan decoder
}
static void make_groups(const hb_mesh_desc *meshes, uint32_t n, std::vector<uint32_t> &starts)
{
	const uint64_t budget = group_half_edges();
	starts.clear();
	uint64_t acc = 0;
	for (uint32_t i = 0; i < n; ++i) {
		if (i == 0 || acc + meshes[i].ne > budget) { starts.push_back(i); acc = 0; }
		acc += meshes[i].ne;
	}
	starts.push_back(n);
}

// HARRY_B200_TRACE_BATCH=1: per group, when its upload / kernels / download began and ended (ms from the start of the
// call, CUDA events on the three streams), printed to stderr at the end of hb_encode_batch / hb_decode_batch
struct BatchTrace {
	bool on = false;
	cudaEvent_t base = nullptr;
	struct Row { cudaEvent_t e[6] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr }; };
	std::vector<Row> rows;
	void begin(hb_ctx *ctx, size_t G)
	{
		const char *env = getenv("HARRY_B200_TRACE_BATCH");
		on = env && env[0] == '1';
		if (!on) return;
		rows.resize(G);
		cudaEventCreate(&base);
		cudaEventRecord(base, ctx->copy_stream);
	}
	void mark(size_t g, int k, cudaStream_t st)
	{
		if (!on) return;
		cudaEventCreate(&rows[g].e[k]);
		cudaEventRecord(rows[g].e[k], st);
	}
	void report(const char *what)
	{
		if (!on) return;
		cudaDeviceSynchronize();
		fprintf(stderr, "[trace] %s: group: upload begin-end | kernels begin-end | download begin-end (ms)\n", what);
		for (size_t g = 0; g < rows.size(); ++g) {
			float t[6] = { -1, -1, -1, -1, -1, -1 };
			for (int k = 0; k < 6; ++k)
				if (rows[g].e[k]) { cudaEventElapsedTime(&t[k], base, rows[g].e[k]); cudaEventDestroy(rows[g].e[k]); }
			fprintf(stderr, "[trace]   %2zu: %8.2f - %8.2f | %8.2f - %8.2f | %8.2f - %8.2f\n", g, t[0], t[1], t[2], t[3], t[4], t[5]);
		}
		cudaEventDestroy(base);
	}
};

extern "C" int hb_encode_batch(hb_ctx *ctx, const hb_mesh_desc *meshes, uint32_t n, const hb_quant_req *q, uint32_t nq, void *const *bounds_out, hb_batch_streams **out)
{
	*out = nullptr;
	if (n == 0) return hb_fail(ctx, HB_ERR_INVALID, "batch: no mesh");
	HB_CUDA(ctx, cudaSetDevice(ctx->device));
	hb_batch_streams *b = (hb_batch_streams *)calloc(1, sizeof(hb_batch_streams));
	hb_batch_streams_priv *pv = new hb_batch_streams_priv();
	if (!b) { delete pv; return hb_fail(ctx, HB_ERR_NOMEM, "out of host memory"); }
	b->priv = pv;
	b->n = n;
	b->mesh = (hb_streams *)calloc(n, sizeof(hb_streams));
	std::vector<uint32_t> starts;
	make_groups(meshes, n, starts);
	const size_t G = starts.size() - 1;
	std::vector<hb_dmesh *> dm(G, nullptr);
	std::vector<FetchState> fs(G);
	int rc = b->mesh ? 0 : hb_fail(ctx, HB_ERR_NOMEM, "out of host memory");
	BatchTrace tr;
	tr.begin(ctx, G);
	auto upload_group = [&](size_t g) -> int {
		dm[g] = new hb_dmesh();
		dm[g]->async_copy = true;
		tr.mark(g, 0, ctx->copy_stream);
		const int r = dmesh_upload_impl(ctx, meshes + starts[g], starts[g + 1] - starts[g], dm[g], false, true);
		tr.mark(g, 1, ctx->copy_stream);
		return r;
	};
	auto run_group = [&](size_t g) -> int {
		hb_dmesh *m = dm[g];
		HB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, m->ev_up[1], 0));
		tr.mark(g, 2, ctx->stream);
		for (uint32_t k = 0; k < nq; ++k) {
			if (q[k].list >= (uint32_t)m->nlists) return hb_fail(ctx, HB_ERR_INVALID, "batch: quantization request for list %u", q[k].list);
			HB_TRY(hb_list_bounds(m, q[k].list, q[k].groups));
			HB_TRY(hb_list_requant(m, q[k].list, q[k].new_quant));
		}
		HB_TRY(hb_encode_lists(m));
		HB_CUDA(ctx, cudaEventCreateWithFlags(&m->ev_done, cudaEventDisableTiming));
		HB_CUDA(ctx, cudaEventRecord(m->ev_done, ctx->stream));
		tr.mark(g, 3, ctx->stream);
		return 0;
	};
	auto download_group = [&](size_t g) -> int {
		hb_dmesh *m = dm[g];
		HB_CUDA(ctx, cudaStreamWaitEvent(ctx->out_stream, m->ev_done, 0));
		tr.mark(g, 4, ctx->out_stream);
		fs[g].seg0 = starts[g];
		for (uint32_t k = 0; k < nq; ++k) { // bounds rows of the group's meshes: min, max, scale per mesh
			if (!bounds_out || !bounds_out[k]) continue;
			DevList &dl = m->lists[q[k].list];
			const size_t s3 = 3 * (size_t)dl.p.stride;
			if (s3) HB_CUDA(ctx, cudaMemcpy2DAsync((uint8_t *)bounds_out[k] + s3 * starts[g], s3, dl.d_bounds, dl.bounds_pitch, s3, m->nseg, cudaMemcpyDeviceToHost, ctx->out_stream));
		}
		HB_TRY(fetch_begin(m, ctx->out_stream, pv, fs[g]));
		tr.mark(g, 5, ctx->out_stream);
		dmesh_release_device(m, ctx->out_stream);
		return 0;
	};
	if (rc == 0) rc = upload_group(0);
	for (size_t g = 0; rc == 0 && g < G; ++g) {
		if (g + 1 < G) rc = upload_group(g + 1);
		if (rc == 0) rc = run_group(g);
		if (rc == 0 && g >= 1) rc = download_group(g - 1);
	}
	if (rc == 0) rc = download_group(G - 1);
	if (rc == 0 && cudaStreamSynchronize(ctx->out_stream) != cudaSuccess) rc = hb_fail(ctx, HB_ERR_CUDA, "batch download failed: %s", cudaGetErrorString(cudaGetLastError()));
	if (rc == 0) rc = hb_check_device_error(ctx, "batch encode");
	if (rc == 0)
		for (size_t g = 0; g < G; ++g)
			for (uint32_t sg = 0; sg < dm[g]->nseg; ++sg) fetch_view(fs[g], pv, sg, &b->mesh[starts[g] + sg]);
	cudaStreamSynchronize(ctx->out_stream);
	tr.report("hb_encode_batch");
	for (hb_dmesh *m : dm) hb_dmesh_free(m);
	if (rc) { hb_batch_streams_free(b); return rc; }
	*out = b;
	return 0;
}

extern "C" int hb_decode_batch(hb_ctx *ctx, const hb_mesh_desc *meshes, uint32_t n, const hb_dequant_req *q, uint32_t nq)
{
	if (n == 0) return hb_fail(ctx, HB_ERR_INVALID, "batch: no mesh");
	HB_CUDA(ctx, cudaSetDevice(ctx->device));
	std::vector<uint32_t> starts;
	make_groups(meshes, n, starts);
	const size_t G = starts.size() - 1;
	std::vector<hb_dmesh *> dm(G, nullptr);
	bool vertex_only = true;
	for (uint32_t i = 0; i < n; ++i) vertex_only = vertex_only && decode_is_vertex_only(&meshes[i]);
	int rc = 0;
	BatchTrace tr;
	tr.begin(ctx, G);
	auto upload_group = [&](size_t g) -> int {
		dm[g] = new hb_dmesh();
		dm[g]->async_copy = true;
		hb_dmesh *m = dm[g];
		tr.mark(g, 0, ctx->copy_stream);
		HB_TRY(dmesh_upload_impl(ctx, meshes + starts[g], starts[g + 1] - starts[g], m, vertex_only));
		for (uint32_t k = 0; k < nq; ++k) { // bounds rows of the lists to dequantize (min, max, scale per mesh)
			if (q[k].list >= (uint32_t)m->nlists || !q[k].bounds) return hb_fail(ctx, HB_ERR_INVALID, "batch: dequantization request for list %u", q[k].list);
			DevList &dl = m->lists[q[k].list];
			const size_t s3 = 3 * (size_t)dl.p.stride;
			if (s3) HB_CUDA(ctx, cudaMemcpy2DAsync(dl.d_bounds, dl.bounds_pitch, (const uint8_t *)q[k].bounds + s3 * starts[g], s3, s3, m->nseg, cudaMemcpyHostToDevice, ctx->copy_stream));
		}
		HB_CUDA(ctx, cudaEventRecord(m->ev_up[1], ctx->copy_stream));
		tr.mark(g, 1, ctx->copy_stream);
		return 0;
	};
	auto run_group = [&](size_t g) -> int {
		hb_dmesh *m = dm[g];
		HB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, m->ev_up[1], 0));
		tr.mark(g, 2, ctx->stream);
		HB_TRY(hb_decode_lists(m));
		uint8_t zero[HB_MAX_COMP] = { 0 };
		for (uint32_t k = 0; k < nq; ++k) HB_TRY(hb_list_requant(m, q[k].list, zero));
		HB_CUDA(ctx, cudaEventCreateWithFlags(&m->ev_done, cudaEventDisableTiming));
		HB_CUDA(ctx, cudaEventRecord(m->ev_done, ctx->stream));
		tr.mark(g, 3, ctx->stream);
		// rows back to the caller's lists
		HB_CUDA(ctx, cudaStreamWaitEvent(ctx->out_stream, m->ev_done, 0));
		tr.mark(g, 4, ctx->out_stream);
		for (int l = 0; l < m->nlists; ++l) {
			DevList &dl = m->lists[l];
			if (!dl.p.ncomp) continue;
			for (uint32_t sg = 0; sg < m->nseg; ++sg) {
				const hb_list_desc &L = meshes[starts[g] + sg].lists[l];
				const size_t bytes = (size_t)L.nrows * L.stride;
				if (bytes) HB_CUDA(ctx, cudaMemcpyAsync(L.rows, dl.p.rows + (size_t)dl.h_rowbase[sg] * L.stride, bytes, cudaMemcpyDeviceToHost, ctx->out_stream));
			}
		}
		tr.mark(g, 5, ctx->out_stream);
		dmesh_release_device(m, ctx->out_stream);
		return 0;
	};
	rc = upload_group(0);
	for (size_t g = 0; rc == 0 && g < G; ++g) {
		if (g + 1 < G) rc = upload_group(g + 1);
		if (rc == 0) rc = run_group(g);
		// at most three groups resident: wait for the downloads of the group before the last one
		if (rc == 0 && g >= 2) { cudaEventSynchronize(dm[g - 2]->ev_done); }
	}
	if (cudaStreamSynchronize(ctx->out_stream) != cudaSuccess && rc == 0) rc = hb_fail(ctx, HB_ERR_CUDA, "batch download failed: %s", cudaGetErrorString(cudaGetLastError()));
	if (rc == 0) rc = hb_check_device_error(ctx, "batch decode");
	tr.report("hb_decode_batch");
	for (hb_dmesh *m : dm) hb_dmesh_free(m);
	return rc;
}
