// hb_internal.cuh -- shared declarations of the sm_100a attribute-path library.
//
// Layout of the device-resident mesh (all in HBM, see DESIGN.md "Data layout"):
//   he[ne]            uint4 per half-edge: {origin vertex, twin half-edge, local|degree<<16, face}
//   vrank[nv]         position of the vertex in the traversal order (0xffffffff = never visited)
//   ord_h/ord_v[n]    half-edge / vertex of traversal position i
//   vc_off/vc_tri     CSR of accepted parallelograms per traversal position, as triples of RANKS
//   rp[l]             "rank-space" value rows of list l: the row bound to element i, components
//                     transposed into a packed power-of-two record (fast path) or u64 containers
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/harry_b200.h"

#define HB_NONE 0xffffffffu

// ------------------------------------------------------------------------------------------------
// host-side objects
// ------------------------------------------------------------------------------------------------
struct hb_ctx {
	int device = 0;
	cudaStream_t stream = nullptr;
	// host-buffer entry points upload on a second stream: the connectivity stages (K0, K3, K4) start as soon as
	// their arrays have landed and run under the rest of the upload
	cudaStream_t copy_stream = nullptr;
	cudaStream_t out_stream = nullptr;  // device -> host copies of pipelined batches
	cudaEvent_t ev[6] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
	std::string err;
	uint64_t launches = 0;
	uint64_t h2d_bytes = 0; // mesh arrays and rows copied host -> device (copy_in)
	float kernel_ms = 0.f, copy_ms = 0.f;
	int *d_err = nullptr;   // device error flag (fan walk overflow, binding out of range)
	int *h_err = nullptr;   // pinned mirror
	uint32_t *h_flag = nullptr; // pinned: answer of the 'is any face / corner row bound twice' check of an encoder upload
	cudaEvent_t ev_flag = nullptr;
	int sm_count = 148;
	// Optional row cache (hb_ctx_set_row_cache): the device copy of a list's rows survives from one host-buffer call to
	// the next of the same pipeline (set_bounds -> requant -> encode; decode -> requant(clear)), keyed by the host
	// pointer of the rows, so the rows travel once per direction instead of once per call.
	struct RowEntry { const void *host; uint32_t nrows, stride; uint8_t *dev; uint8_t quant[HB_MAX_COMP]; int ncomp; };
	bool row_cache = false;
	std::vector<RowEntry> rows_cached;
	// optional per-kernel timing (CUDA events around every launch on the context stream)
	bool profiling = false;
	struct ProfRec { const char *name; cudaEvent_t a, b; };
	std::vector<ProfRec> prof;
	std::vector<cudaEvent_t> ev_pool;
};
cudaEvent_t hb_prof_event(hb_ctx *ctx);

// argument block of the vertex-reconstruction kernels (hb_decode_spec.cuh, hb_decode_scan.cuh): one per segment.
// All indices are global (positions in the concatenated traversal order); a kernel works on [base, n).
struct SpecArgs {
	const uint8_t *kind;      // per rank: 0 skip, 1 DATA, 2 copy from src[i] (HIST)
	const uint32_t *src;      // kind 2: owning rank
	const uint32_t *cand_off; // n + 1
	const uint32_t *cand;     // rank triples
	const void *resid;        // compact records: residuals (read only)
	void *x;                  // compact records: values (in/out)
	uint32_t base;            // first rank of the segment
	uint32_t n;               // end of the segment (exclusive)
	int bits[4];              // quantization bits per component (prediction.h:22-25)
	unsigned long long *stats; // [0] sweeps, [1] hypothesis sweeps, [2] plain sweeps (segment 0 only)
	const void *srec;         // hb_decode_scan.cuh: one ScanRec per rank
};
struct ListParams {
	uint8_t *rows;          // AoS rows in HBM (nrows * stride)
	uint32_t nrows, stride;
	int ncomp;
	int target;
	uint32_t sym_stride;    // bytes of one residual row
	int uniform_stype;      // storage type shared by all components, or -1
	uint8_t type[HB_MAX_COMP], quant[HB_MAX_COMP], stype[HB_MAX_COMP], size[HB_MAX_COMP];
	uint16_t offset[HB_MAX_COMP], sym_off[HB_MAX_COMP];
};

// decode: up to four components of one storage type, reconstructed together on compact rank-space records
struct SpecGroup {
	int st = 0, ncomp = 0;
	int comp[4] = { 0, 0, 0, 0 };
	uint8_t *d_cres = nullptr, *d_cx = nullptr; // residual / value records
	SpecArgs *d_args = nullptr;                 // one argument block per segment
};

struct DevList {
	ListParams p;
	// bounds rows (dequantized layout, `stride` bytes each): min, max, scale -- one triple per segment (mesh of a
	// batch), `bounds_pitch` bytes apart
	uint8_t *d_bounds = nullptr;
	size_t bounds_pitch = 0;
	// segments: rows of mesh s are [rowbase[s], rowbase[s] + rownum[s]) of the concatenated list (bases padded to
	// multiples of 16 rows, so that every segment starts on a 16-byte boundary whatever the stride)
	std::vector<uint32_t> h_rowbase, h_rownum;
	uint32_t *d_rowbase = nullptr, *d_rownum = nullptr;
	std::vector<uint32_t> h_emitbase; // decode: offsets of the segments in the concatenated type-symbol stream
	// encode outputs (device)
	uint32_t n_elems = 0;       // elements of the list's target class (emission slots)
	uint32_t n_emit = 0, n_data = 0;
	uint8_t *d_type = nullptr;
	uint32_t *d_aux = nullptr;
	uint8_t *d_sym = nullptr;
	unsigned long long *d_hist = nullptr;      // sym_stride * 256
	unsigned long long *d_type_hist = nullptr; // 4
	// per-element work arrays (element = emission slot of the list's target class)
	uint32_t *d_erow = nullptr;   // bound attribute row or HB_NONE
	uint32_t *d_ek = nullptr;     // emission index (exclusive scan of bound flags); nullptr: == element index
	uint32_t *d_first = nullptr;  // per attribute row: first referencing element (its DATA emission)
	uint32_t *d_dord = nullptr;   // exclusive scan of DATA flags (n_elems + 1)
	unsigned long long *d_rp = nullptr; // rank-space value records, ncomp u64 containers per element
	// speculative vertex decode (hb_decode_spec.cuh)
	uint8_t *d_kind = nullptr;
	uint32_t *d_src = nullptr;
	uint8_t *d_cx = nullptr;            // encode: packed rank-space value records
	std::vector<SpecGroup> groups;      // decode: component groups (hb_decode.cu)
	unsigned long long *d_spec_stats = nullptr;
	uint32_t *d_chain = nullptr;        // float lists: per segment {ranks reading their predecessor, DATA ranks}
	void *d_srec = nullptr;             // hb_decode_scan.cuh records
	uint32_t *d_wide = nullptr;         // encode: elements deferred to the warp-per-element kernel
	uint8_t *d_emit_type = nullptr;     // decode: drained type symbols (optional)
	uint32_t emit_count = 0;
	bool rows_from_cache = false;       // the rows were already on the device (row cache): nothing was uploaded
	uint32_t *d_dup = nullptr;          // device flag: some row is referenced twice (else the "no shared row" shortcuts apply)
	bool d_dup_active = false;          // the flag is maintained for this encode (hb_prepare_list_elems)
	bool nocomp_fast = false;           // encode: zero-component list coded by the two-pass path (hb_encode.cu)
	uint8_t *d_done = nullptr;          // decode: corner wavefront flags
	uint32_t *d_remaining = nullptr;
	uint8_t *d_rows_backup = nullptr;   // hb_dmesh_snapshot
	uint8_t backup_quant[HB_MAX_COMP];
};

struct hb_dmesh {
	hb_ctx *ctx = nullptr;
	// A device mesh is the concatenation of `nseg` independent meshes (a batch; nseg == 1 for a single mesh): vertices,
	// faces, half-edges, traversal orders and attribute rows of segment s occupy [base[s], base[s + 1]) of the global
	// index spaces.  Prediction never crosses a segment (a candidate needs vertices of the same connected component
	// that were coded earlier), so every kernel behind the ingest kernels runs on the concatenation unchanged.
	uint32_t nseg = 1;
	std::vector<uint32_t> h_vbase, h_fbase, h_ebase, h_obase, h_ofbase; // nseg + 1 each
	uint32_t *d_segtab = nullptr;                                        // the five tables, back to back
	const uint32_t *d_vbase = nullptr, *d_fbase = nullptr, *d_ebase = nullptr, *d_obase = nullptr, *d_ofbase = nullptr;
	uint32_t *d_cebase = nullptr;          // corner-element base per segment (computed on the device)
	uint32_t *d_face_off_raw = nullptr;    // per-segment CSR offsets as uploaded (nseg > 1); d_face_off is the global CSR
	std::vector<SpecArgs> h_spec_args; // kernel argument blocks staged for upload (kept until the mesh is freed)
	uint32_t nv = 0, nf = 0, ne = 0, norder = 0, norder_f = 0;
	uint16_t nb_face = 0, nb_vtx = 0, nb_corner = 0, nregs_face = 0, nregs_vtx = 0, nlists = 0;
	bool has_order_f = false;
	bool order_f_late = false; // order_f was uploaded behind everything else (the face stages wait for the whole upload)
	bool async_copy = false;     // uploads go to ctx->copy_stream (hb_attr_encode / hb_attr_decode)
	bool alloc_on_copy_stream = false; // while the upload buffers are being allocated
	// Batches: a group of ~200 meshes is ~1500 host arrays of 0.4-7 MB each.  One cudaMemcpyAsync per array moves them at
	// 46 GB/s on a 55.6 GB/s link; arrays in page-locked host memory are therefore queued as descriptors and issued per
	// upload stage as ONE cudaMemcpyBatchAsync (54 GB/s).  Runtimes without it: one kernel that reads the host memory
	// directly (k_upload_gather, 49-50 GB/s).
	struct UpDesc { const void *src; void *dst; unsigned long long bytes; };
	bool gather_uploads = false;
	int upload_mode = 2;         // 1: k_upload_gather reads the host arrays; 2: cudaMemcpyBatchAsync (falls back to 1)
	std::vector<UpDesc> up_pending;
	bool take_cached_rows = false;     // rows may come from the context's row cache (requant / encode continue a pipeline)
	cudaEvent_t ev_up[2] = { nullptr, nullptr }; // copy-stream milestones: connectivity landed / everything landed
	cudaEvent_t ev_done = nullptr;               // kernels of this mesh / group finished (pipelined batches)
	std::vector<void *> allocs;
	// uploaded
	uint8_t *d_edges_raw = nullptr;
	uint32_t *d_face_off = nullptr;
	uint8_t *d_order = nullptr, *d_order_f = nullptr;
	uint16_t *d_vtx_regs = nullptr, *d_face_regs = nullptr;
	uint32_t *d_bind_face = nullptr, *d_bind_vtx = nullptr, *d_bind_corner = nullptr;
	// region tables: slot of list l in region r, or -1 (nregs * nlists, int16)
	int16_t *d_slot_vtx = nullptr, *d_slot_face = nullptr, *d_slot_corner = nullptr;
	std::vector<int16_t> h_slot_vtx, h_slot_face, h_slot_corner;
	std::vector<int> reg_ncorner;   // corner bindings per face region
	bool any_corner = false;
	// derived connectivity
	uint4 *d_he = nullptr;
	uint32_t *d_vrank = nullptr, *d_ord_h = nullptr, *d_ord_v = nullptr;
	uint32_t *d_frank = nullptr, *d_ford_h = nullptr; // face rank by face, gate half-edge by face rank
	uint32_t *d_he_celem = nullptr;   // half-edge -> corner element
	uint32_t *d_lh = nullptr;         // local-history offsets [nb_corner][n_corner_elems]
	int *d_reg_ncorner = nullptr;     // corner bindings per face region
	uint32_t *d_cbase = nullptr;    // corner-element base per face rank (norder_f + 1)
	uint32_t n_corner_elems = 0;
	uint32_t *d_vc_off = nullptr, *d_vc_tri = nullptr, *d_vc_stage = nullptr; uint32_t vc_total = 0;
	void *d_vc_wide = nullptr;          // wide-fan control block and scratch (hb_conn.cu)
	uint32_t *d_vc_wslot = nullptr;     // vertex -> wide slot
	uint32_t *d_vc_wbits = nullptr;     // one bit per vertex: registered as wide
	uint32_t *d_vc_wnodes = nullptr, *d_vc_wpos = nullptr, *d_vc_wwork = nullptr, *d_vc_worder = nullptr, *d_vc_warena = nullptr;
	uint32_t vc_wide_cap = 0;
	uint32_t *d_cc_off = nullptr, *d_cc_idx = nullptr; uint32_t cc_total = 0;
	uint32_t *d_celem_h = nullptr;  // half-edge of corner element
	bool conn_ready = false, vcand_ready = false, ccand_ready = false;
	std::vector<DevList> lists;
	void *d_walks = nullptr;            // decode: argument blocks of the generic chain walker
	bool encoded = false;
};

// segment lookup: the s < nseg with base[s] <= x < base[s + 1] (empty segments are skipped)
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t hb_seg_find(const uint32_t *__restrict__ base, uint32_t nseg, uint32_t x)
{
	uint32_t lo = 0, hi = nseg;
	while (hi - lo > 1) {
		const uint32_t mid = (lo + hi) >> 1;
		if (__ldg(base + mid) <= x) lo = mid;
		else hi = mid;
	}
	return lo;
}
#endif

// error plumbing ---------------------------------------------------------------------------------
int hb_fail(hb_ctx *ctx, int code, const char *fmt, ...);
#define HB_CUDA(ctx, call)                                                                          \
	do {                                                                                            \
		cudaError_t e__ = (call);                                                                   \
		if (e__ != cudaSuccess)                                                                     \
			return hb_fail((ctx), HB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
	} while (0)
#define HB_TRY(expr)                                                                                \
	do {                                                                                            \
		int rc__ = (expr);                                                                          \
		if (rc__ != 0) return rc__;                                                                 \
	} while (0)
// every kernel launch goes through this (counts launches, checks the launch)
#define HB_LAUNCH(ctx, kernel, grid, block, smem, ...)                                              \
	do {                                                                                            \
		cudaEvent_t pa__ = nullptr, pb__ = nullptr;                                                 \
		if ((ctx)->profiling) {                                                                     \
			pa__ = hb_prof_event(ctx);                                                              \
			pb__ = hb_prof_event(ctx);                                                              \
			cudaEventRecord(pa__, (ctx)->stream);                                                   \
		}                                                                                           \
		kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);                            \
		(ctx)->launches++;                                                                          \
		if (pa__) {                                                                                 \
			cudaEventRecord(pb__, (ctx)->stream);                                                   \
			(ctx)->prof.push_back(hb_ctx::ProfRec{ #kernel, pa__, pb__ });                          \
		}                                                                                           \
		HB_CUDA((ctx), cudaGetLastError());                                                         \
	} while (0)

static inline uint32_t hb_div_up(uint64_t a, uint32_t b) { return (uint32_t)((a + b - 1) / b); }

// device memory helpers (stream-ordered pool)
int hb_dalloc(hb_dmesh *m, void **p, size_t bytes);
template <typename T> static inline int hb_dalloc_t(hb_dmesh *m, T **p, size_t n) { return hb_dalloc(m, (void **)p, n * sizeof(T)); }
int hb_check_device_error(hb_ctx *ctx, const char *what);

// stages (implemented across the .cu files) ----------------------------------------------------
int hb_scan_exclusive_u32(hb_ctx *ctx, const uint32_t *d_in, uint32_t *d_out, uint32_t n, uint32_t *d_total /* device, may be null */,
                          const uint32_t *d_skip_if_zero = nullptr /* device flag: 0 = return at once */);
int hb_build_conn(hb_dmesh *m);             // he[], ranks, orders
int hb_build_vertex_candidates(hb_dmesh *m); // vc_off / vc_tri
int hb_build_corner_candidates(hb_dmesh *m); // cc_off / cc_idx
int hb_prepare_list_elems(hb_dmesh *m, int l, bool need_rp, bool decode);
int hb_encode_lists(hb_dmesh *m);
int hb_nocomp_finish(hb_dmesh *m, int l, cudaStream_t st);
int hb_decode_lists(hb_dmesh *m);
int hb_list_bounds(hb_dmesh *m, uint32_t l, const uint8_t *groups /* or nullptr: min and max rows only */);
int hb_list_scale(hb_dmesh *m, uint32_t l, const uint8_t *groups);
int hb_list_requant(hb_dmesh *m, uint32_t l, const uint8_t *new_quant);
void hb_fill_list_params(ListParams &p, const hb_list_desc &L);
int hb_twin_build(hb_dmesh *m, uint32_t nv, uint32_t nf, uint32_t ne, const uint32_t *d_face_off, const uint32_t *d_org,
                  uint32_t os, uint32_t *d_out); // hb_twin.cu

// ------------------------------------------------------------------------------------------------
// device arithmetic
// ------------------------------------------------------------------------------------------------
#ifdef __CUDACC__

__host__ __device__ static inline int hb_type_size(int t)
{
	switch (t) {
	case HB_FLOAT: case HB_UINT: case HB_INT: return 4;
	case HB_DOUBLE: case HB_ULONG: case HB_LONG: return 8;
	case HB_USHORT: case HB_SHORT: return 2;
	case HB_UCHAR: case HB_CHAR: return 1;
	default: return 0;
	}
}

// half-edge record helpers (conn.h:63-70,137-144 in flattened form)
__device__ __forceinline__ uint32_t he_next(uint32_t h, uint32_t ld)
{
	const uint32_t le = ld & 0xffffu, deg = ld >> 16;
	return (le + 1 == deg) ? h - le : h + 1;
}
__device__ __forceinline__ uint32_t he_prev(uint32_t h, uint32_t ld)
{
	const uint32_t le = ld & 0xffffu, deg = ld >> 16;
	return (le == 0) ? h + deg - 1 : h - 1;
}

template <typename T> struct hb_unsigned;
template <> struct hb_unsigned<uint8_t> { typedef uint8_t type; };
template <> struct hb_unsigned<int8_t> { typedef uint8_t type; };
template <> struct hb_unsigned<uint16_t> { typedef uint16_t type; };
template <> struct hb_unsigned<int16_t> { typedef uint16_t type; };
template <> struct hb_unsigned<uint32_t> { typedef uint32_t type; };
template <> struct hb_unsigned<int32_t> { typedef uint32_t type; };
template <> struct hb_unsigned<uint64_t> { typedef uint64_t type; };
template <> struct hb_unsigned<int64_t> { typedef uint64_t type; };

// Integer residual arithmetic over a storage type T; `bits` = quantization bits, or the width
// of T for unquantized components (formats/hry/prediction.h:22-31).  Sums and differences wrap
// modulo 2^width (computed in the unsigned twin), comparisons keep T's signedness.
template <typename T> struct IntOps {
	typedef typename hb_unsigned<T>::type U;
	static __device__ __forceinline__ T mask(int bits) { return bits == (int)(8 * sizeof(T)) ? (T)-1 : (T)((1u << bits) - 1u); }
	// prediction.h:121-137
	static __device__ __forceinline__ T predict(T v0, T v1, T v2, int bits)
	{
		const T hi = mask(bits);
		if (v1 < v2) {
			const T d = (T)((U)v2 - (U)v1);
			return d > v0 ? (T)0 : (T)((U)v0 - (U)d);
		}
		const T d = (T)((U)v1 - (U)v2);
		const T v = (T)((U)v0 + (U)d);
		return (v > hi || v < v0) ? hi : v;
	}
	// same with the mask precomputed (hot loops)
	static __device__ __forceinline__ T predict_hi(T v0, T v1, T v2, T hi)
	{
		if (v1 < v2) {
			const T d = (T)((U)v2 - (U)v1);
			return d > v0 ? (T)0 : (T)((U)v0 - (U)d);
		}
		const T d = (T)((U)v1 - (U)v2);
		const T v = (T)((U)v0 + (U)d);
		return (v > hi || v < v0) ? hi : v;
	}
	static __device__ __forceinline__ T dec_hi(T delta, T pred, T hi)
	{
		const T room = (T)((U)hi - (U)pred);
		if (pred == (T)0) return delta;
		const T pm1 = (T)((U)pred - (U)1);
		const T bal = room < pm1 ? room : pm1;
		const T half = (T)(delta >> 1);
		if (half > bal) {
			if (room >= pred) return (T)((U)pred + (U)delta - (U)bal - (U)1);
			return (T)((U)pred - (U)delta + (U)bal);
		}
		return (T)((U)pred + ((U)half ^ ((delta & 1) ? (U) ~(U)0 : (U)0)));
	}
	// prediction.h:81-99
	static __device__ __forceinline__ T enc(T raw, T pred, int bits)
	{
		const T room = (T)((U)mask(bits) - (U)pred);
		if (pred == (T)0) return raw;
		const T bal = room < pred ? room : pred;
		const bool below = raw < pred;
		const T d = below ? (T)((U)pred - (U)raw) : (T)((U)raw - (U)pred);
		if (d > bal) return (T)((U)d + (U)bal);
		return below ? (T)(((U)d << 1) - (U)1) : (T)((U)d << 1);
	}
	// prediction.h:46-63
	static __device__ __forceinline__ T dec(T delta, T pred, int bits)
	{
		const T room = (T)((U)mask(bits) - (U)pred);
		if (pred == (T)0) return delta;
		const T pm1 = (T)((U)pred - (U)1);
		const T bal = room < pm1 ? room : pm1;
		const T half = (T)(delta >> 1);
		if (half > bal) {
			if (room >= pred) return (T)((U)pred + (U)delta - (U)bal - (U)1);
			return (T)((U)pred - (U)delta + (U)bal);
		}
		return (T)((U)pred + ((U)half ^ ((delta & 1) ? (U) ~(U)0 : (U)0)));
	}
};

// transform.h:19-23 -- order-preserving bit map of IEEE-754 floats (an involution)
__device__ __forceinline__ uint32_t hb_flip_f32(uint32_t i) { return i ^ ((0u - (i >> 31)) >> 1); }

// sign/zero extension of a container holding a value of integer storage type st
__device__ __forceinline__ long long hb_bits_to_i64(unsigned long long b, int st)
{
	switch (st) {
	case HB_CHAR: return (int8_t)b;
	case HB_UCHAR: return (uint8_t)b;
	case HB_SHORT: return (int16_t)b;
	case HB_USHORT: return (uint16_t)b;
	case HB_INT: return (int32_t)b;
	case HB_UINT: return (uint32_t)b;
	default: return (long long)b;
	}
}

__device__ __forceinline__ int hb_stype_bits(int st, int q) { return q == 0 ? 8 * hb_type_size(st) : q; }

// runtime-typed scalar ops on u64 bit containers (generic path)
__device__ __forceinline__ unsigned long long hb_predict(int st, unsigned long long v0, unsigned long long v1, unsigned long long v2, int q)
{
	const int bits = hb_stype_bits(st, q);
	switch (st) {
	case HB_FLOAT: return __float_as_uint(__fadd_rn(__uint_as_float((uint32_t)v0), __fsub_rn(__uint_as_float((uint32_t)v1), __uint_as_float((uint32_t)v2))));
	case HB_UCHAR: return IntOps<uint8_t>::predict((uint8_t)v0, (uint8_t)v1, (uint8_t)v2, bits);
	case HB_CHAR: return (uint8_t)IntOps<int8_t>::predict((int8_t)v0, (int8_t)v1, (int8_t)v2, bits);
	case HB_USHORT: return IntOps<uint16_t>::predict((uint16_t)v0, (uint16_t)v1, (uint16_t)v2, bits);
	case HB_SHORT: return (uint16_t)IntOps<int16_t>::predict((int16_t)v0, (int16_t)v1, (int16_t)v2, bits);
	case HB_UINT: return IntOps<uint32_t>::predict((uint32_t)v0, (uint32_t)v1, (uint32_t)v2, bits);
	case HB_INT: return (uint32_t)IntOps<int32_t>::predict((int32_t)v0, (int32_t)v1, (int32_t)v2, bits);
	case HB_ULONG: return IntOps<uint64_t>::predict(v0, v1, v2, bits);
	case HB_LONG: return (unsigned long long)IntOps<int64_t>::predict((int64_t)v0, (int64_t)v1, (int64_t)v2, bits);
	default: return 0;
	}
}
__device__ __forceinline__ unsigned long long hb_enc(int st, unsigned long long raw, unsigned long long pred, int q)
{
	const int bits = hb_stype_bits(st, q);
	switch (st) {
	case HB_FLOAT: return IntOps<uint32_t>::enc(hb_flip_f32((uint32_t)raw), hb_flip_f32((uint32_t)pred), bits);
	case HB_UCHAR: return IntOps<uint8_t>::enc((uint8_t)raw, (uint8_t)pred, bits);
	case HB_CHAR: return (uint8_t)IntOps<int8_t>::enc((int8_t)raw, (int8_t)pred, bits);
	case HB_USHORT: return IntOps<uint16_t>::enc((uint16_t)raw, (uint16_t)pred, bits);
	case HB_SHORT: return (uint16_t)IntOps<int16_t>::enc((int16_t)raw, (int16_t)pred, bits);
	case HB_UINT: return IntOps<uint32_t>::enc((uint32_t)raw, (uint32_t)pred, bits);
	case HB_INT: return (uint32_t)IntOps<int32_t>::enc((int32_t)raw, (int32_t)pred, bits);
	case HB_ULONG: return IntOps<uint64_t>::enc(raw, pred, bits);
	case HB_LONG: return (unsigned long long)IntOps<int64_t>::enc((int64_t)raw, (int64_t)pred, bits);
	default: return 0;
	}
}
__device__ __forceinline__ unsigned long long hb_dec(int st, unsigned long long delta, unsigned long long pred, int q)
{
	const int bits = hb_stype_bits(st, q);
	switch (st) {
	case HB_FLOAT: return hb_flip_f32(IntOps<uint32_t>::dec((uint32_t)delta, hb_flip_f32((uint32_t)pred), bits));
	case HB_UCHAR: return IntOps<uint8_t>::dec((uint8_t)delta, (uint8_t)pred, bits);
	case HB_CHAR: return (uint8_t)IntOps<int8_t>::dec((int8_t)delta, (int8_t)pred, bits);
	case HB_USHORT: return IntOps<uint16_t>::dec((uint16_t)delta, (uint16_t)pred, bits);
	case HB_SHORT: return (uint16_t)IntOps<int16_t>::dec((int16_t)delta, (int16_t)pred, bits);
	case HB_UINT: return IntOps<uint32_t>::dec((uint32_t)delta, (uint32_t)pred, bits);
	case HB_INT: return (uint32_t)IntOps<int32_t>::dec((int32_t)delta, (int32_t)pred, bits);
	case HB_ULONG: return IntOps<uint64_t>::dec(delta, pred, bits);
	case HB_LONG: return (unsigned long long)IntOps<int64_t>::dec((int64_t)delta, (int64_t)pred, bits);
	default: return 0;
	}
}

// transform::divround for the integer accumulator types (transform.h:90-91): (n + (d >> 1)) / d
// with C++ truncating division; fast paths avoid the 64-bit divide for the common small cases.
__device__ __forceinline__ long long hb_divround_i64(long long n, int k)
{
	const long long s = n + (long long)(k >> 1);
	if (k == 1) return s;
	if (s >= 0 && s < 0x7fffffffLL) return (long long)((uint32_t)s / (uint32_t)k);
	return s / (long long)k;
}

// attrcode.h:199-205: keep the candidate closest to the mean, later candidates win ties
__device__ __forceinline__ float hb_closest_step(float res, float p, float avg)
{
	const float rd = avg > res ? __fsub_rn(avg, res) : __fsub_rn(res, avg);
	const float pd = avg > p ? __fsub_rn(avg, p) : __fsub_rn(p, avg);
	return rd < pd ? res : p;
}

// AoS rows follow mixing::Fmt offsets (no alignment padding, mixing.h:60), so a component may be
// misaligned for its size: fall back to byte accesses then.
__device__ __forceinline__ unsigned long long hb_ld_bits(const uint8_t *p, int size)
{
	if ((((size_t)p) & (size_t)(size - 1)) == 0) {
		switch (size) {
		case 1: return *p;
		case 2: return *(const uint16_t *)p;
		case 4: return *(const uint32_t *)p;
		case 8: return *(const unsigned long long *)p;
		default: return 0;
		}
	}
	unsigned long long v = 0;
	for (int b = 0; b < size; ++b) v |= (unsigned long long)p[b] << (8 * b);
	return v;
}
__device__ __forceinline__ void hb_st_bits(uint8_t *p, int size, unsigned long long v)
{
	if ((((size_t)p) & (size_t)(size - 1)) == 0) {
		switch (size) {
		case 1: *p = (uint8_t)v; return;
		case 2: *(uint16_t *)p = (uint16_t)v; return;
		case 4: *(uint32_t *)p = (uint32_t)v; return;
		case 8: *(unsigned long long *)p = v; return;
		default: return;
		}
	}
	for (int b = 0; b < size; ++b) p[b] = (uint8_t)(v >> (8 * b));
}

#endif // __CUDACC__
