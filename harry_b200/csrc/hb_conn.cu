// hb_conn.cu -- connectivity-side kernels of the attribute path (no attribute values touched):
//   K0 flatten_halfedges   raw Conn::edgeorg records -> 16-byte half-edge records
//   K3 rank kernels        traversal order -> vrank / ord_h / ord_v, face ranks, corner elements
//   K4 fan_gather          per traversed vertex (corner): ordered fan walk, emit the accepted
//                          parallelogram rank triples (corner candidates) as a CSR
// plus the device-wide exclusive scan used to turn counts into CSR offsets.
// Reference semantics: formats/hry/attrcode.h:83-106 (TFAN_IT), :117-134 (use_paral),
// :135-154 (use_corner), :155-171 (paral); structs/conn.h:123-160.
#include "hb_internal.cuh"

// ------------------------------------------------------------------------------------------------
// exclusive scan (single pass, decoupled look-back), out[n] receives the total
// ------------------------------------------------------------------------------------------------
#define SCAN_THREADS 256
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *total)
{
	__shared__ uint32_t warp_sums[SCAN_THREADS / 32];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t x = v;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
		if (lane >= d) x += y;
	}
	if (lane == 31) warp_sums[warp] = x;
	__syncthreads();
	if (warp == 0) {
		uint32_t s = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0;
#pragma unroll
		for (int d = 1; d < SCAN_THREADS / 32; d <<= 1) {
			const uint32_t y = __shfl_up_sync(0xffffffffu, s, d);
			if (lane >= d) s += y;
		}
		if (lane < SCAN_THREADS / 32) warp_sums[lane] = s;
	}
	__syncthreads();
	const uint32_t base = warp ? warp_sums[warp - 1] : 0;
	*total = warp_sums[SCAN_THREADS / 32 - 1];
	__syncthreads();
	return base + x - v;
}

// Single pass with decoupled look-back: a tile publishes its aggregate, looks back over the status
// words of its predecessors until it meets an inclusive prefix, and publishes its own.  Tiles are
// numbered by an atomic ticket (a tile never waits for one that has not started).  Status word:
// bits 62..63 = 0 empty / 1 aggregate / 2 inclusive prefix, low 32 bits = value.
#define SCAN_VEC (SCAN_ITEMS / 4)
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_lookback(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, unsigned long long *__restrict__ status,
                                                               uint32_t *__restrict__ ticket, uint32_t n, uint32_t ntiles, uint32_t *__restrict__ out_total2,
                                                               const uint32_t *__restrict__ skip_if_zero)
{
	__shared__ uint32_t s_tile, s_prefix;
	if (skip_if_zero && *skip_if_zero == 0u) return; // the caller's consumers do not need the result (decided on the device)
	if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
	__syncthreads();
	const uint32_t tile = s_tile;
	const uint32_t base = tile * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
	uint32_t v[SCAN_ITEMS], s = 0;
	if (base + SCAN_ITEMS <= n && ((((size_t)in) & 15u) == 0)) {
#pragma unroll
		for (int q = 0; q < SCAN_VEC; ++q) {
			const uint4 w = ((const uint4 *)(in + base))[q];
			v[4 * q] = w.x; v[4 * q + 1] = w.y; v[4 * q + 2] = w.z; v[4 * q + 3] = w.w;
		}
	} else {
#pragma unroll
		for (int k = 0; k < SCAN_ITEMS; ++k) v[k] = base + k < n ? in[base + k] : 0;
	}
#pragma unroll
	for (int k = 0; k < SCAN_ITEMS; ++k) s += v[k];
	uint32_t total;
	uint32_t ex = block_exclusive_scan(s, &total);
	if (threadIdx.x == 0) {
		volatile unsigned long long *st = status;
		if (tile == 0) {
			st[0] = (2ull << 62) | total;
			s_prefix = 0;
		} else {
			st[tile] = (1ull << 62) | total;
		}
	}
	if (tile > 0 && threadIdx.x < 32) {
		// warp 0 looks back 32 tiles at a time
		volatile unsigned long long *st = status;
		uint32_t acc = 0;
		int j = (int)tile - 1;
		for (;;) {
			const int k = j - (int)threadIdx.x;
			unsigned long long w = k >= 0 ? st[k] : (2ull << 62);
			// wait until every word of this window is published
			while (__any_sync(0xffffffffu, (w >> 62) == 0)) w = k >= 0 ? st[k] : (2ull << 62);
			const unsigned incl = __ballot_sync(0xffffffffu, (w >> 62) == 2);
			const int first = incl ? __ffs((int)incl) - 1 : 32; // nearest inclusive prefix in the window
			uint32_t part = (int)threadIdx.x <= first ? (uint32_t)w : 0u;
#pragma unroll
			for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
			acc += part;
			if (incl) break;
			j -= 32;
		}
		if (threadIdx.x == 0) {
			s_prefix = acc;
			__threadfence();
			st[tile] = (2ull << 62) | (unsigned long long)(acc + total);
		}
	}
	__syncthreads();
	ex += s_prefix;
	if (base + SCAN_ITEMS <= n && ((((size_t)out) & 15u) == 0)) {
#pragma unroll
		for (int q = 0; q < SCAN_VEC; ++q) {
			uint4 w;
			w.x = ex; ex += v[4 * q];
			w.y = ex; ex += v[4 * q + 1];
			w.z = ex; ex += v[4 * q + 2];
			w.w = ex; ex += v[4 * q + 3];
			((uint4 *)(out + base))[q] = w;
		}
	} else {
#pragma unroll
		for (int k = 0; k < SCAN_ITEMS; ++k) {
			if (base + k < n) out[base + k] = ex;
			ex += v[k];
		}
	}
	if (tile == ntiles - 1 && threadIdx.x == 0) {
		out[n] = s_prefix + total;
		if (out_total2) *out_total2 = s_prefix + total;
	}
}
__global__ void k_scan_empty(uint32_t *out, uint32_t *out_total2)
{
	out[0] = 0;
	if (out_total2) *out_total2 = 0;
}

// d_out must hold n + 1 entries; d_out[n] = total.  in == out is allowed.
int hb_scan_exclusive_u32(hb_ctx *ctx, const uint32_t *d_in, uint32_t *d_out, uint32_t n, uint32_t *d_total, const uint32_t *d_skip_if_zero)
{
	const uint32_t ntiles = n ? hb_div_up(n, SCAN_TILE) : 0;
	if (!ntiles) {
		HB_LAUNCH(ctx, k_scan_empty, 1, 1, 0, d_out, d_total);
		return 0;
	}
	unsigned long long *d_status = nullptr;
	HB_CUDA(ctx, cudaMallocAsync((void **)&d_status, sizeof(unsigned long long) * ((size_t)ntiles + 1), ctx->stream));
	HB_CUDA(ctx, cudaMemsetAsync(d_status, 0, sizeof(unsigned long long) * ((size_t)ntiles + 1), ctx->stream));
	HB_LAUNCH(ctx, k_scan_lookback, ntiles, SCAN_THREADS, 0, d_in, d_out, d_status, (uint32_t *)(d_status + ntiles), n, ntiles, d_total, d_skip_if_zero);
	HB_CUDA(ctx, cudaFreeAsync(d_status, ctx->stream));
	return 0;
}

// ------------------------------------------------------------------------------------------------
// K0: raw 12-byte {org, twin_face, twin_edge:16} records -> uint4 {org, twin, le | deg << 16, face}
// The raw records of a segment (mesh of a batch) carry indices local to it; the segment of a face is found by
// binary search over the face bases (nseg == 1: no search).  A record that is flagged as invalid is stored in a
// SAFE form (origin 0, its own twin; faces of impossible degree as isolated one-edge faces), so the kernels
// behind K0 -- which run before the host looks at the error flag -- never leave their arrays.
// ------------------------------------------------------------------------------------------------
struct SegView {
	uint32_t nseg;
	const uint32_t *vbase, *fbase, *ebase, *obase, *ofbase;
};

// global CSR offsets of the concatenated faces from the per-segment ones (segment s uploaded nf_s + 1 offsets
// starting at 0, stored at fbase[s] + s)
#define FOFF_PER_THREAD 8 // (one face per thread made this copy-sized kernel wait for CTA slots: 1.6 TB/s)
__global__ void __launch_bounds__(256) k_global_face_off(const uint32_t *__restrict__ raw, SegView sv, uint32_t nf, uint32_t *__restrict__ face_off)
{
	// one binary search per CTA, then a step or two forward: a segment is thousands of faces long
	__shared__ uint32_t s_seg;
	const uint32_t f0 = blockIdx.x * (blockDim.x * FOFF_PER_THREAD);
	if (threadIdx.x == 0) s_seg = hb_seg_find(sv.fbase, sv.nseg, min(f0, nf ? nf - 1 : 0u));
	__syncthreads();
	uint32_t s = s_seg;
#pragma unroll
	for (int k = 0; k < FOFF_PER_THREAD; ++k) {
		const uint32_t f = f0 + k * blockDim.x + threadIdx.x;
		if (f > nf) return;
		if (f == nf) { face_off[f] = sv.ebase[sv.nseg]; return; }
		while (s + 1 < sv.nseg && f >= sv.fbase[s + 1]) ++s;
		face_off[f] = raw[f + s] + sv.ebase[s];
	}
}

__global__ void __launch_bounds__(256) k_flatten_halfedges(const uint32_t *__restrict__ raw, const uint32_t *__restrict__ face_off, uint4 *__restrict__ he,
                                                           uint32_t nf, uint32_t ne, SegView sv, int *err)
{
	const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
	if (f >= nf) return;
	const uint32_t s = hb_seg_find(sv.fbase, sv.nseg, f);
	const uint32_t vb = sv.vbase[s], fb = sv.fbase[s];
	const uint32_t nv_s = sv.vbase[s + 1] - vb, nf_s = sv.fbase[s + 1] - fb;
	const uint32_t b = face_off[f], e = face_off[f + 1];
	const uint32_t deg = e - b;
	if (e < b || e > ne) { atomicExch(err, 1); return; }
	if (deg > 0xffffu) {
		atomicExch(err, 1);
		for (uint32_t h = b; h < e; ++h) he[h] = make_uint4(vb, h, 0u | (1u << 16), f);
		return;
	}
	for (uint32_t h = b; h < e; ++h) {
		uint32_t org = raw[3 * (size_t)h];
		const uint32_t tf = raw[3 * (size_t)h + 1], te = raw[3 * (size_t)h + 2] & 0xffffu;
		uint32_t tw = h;
		if (org >= nv_s || tf >= nf_s) {
			atomicExch(err, 2);
			if (org >= nv_s) org = 0;
		} else {
			const uint32_t tb = face_off[tf + fb];
			if (tb + te >= face_off[tf + fb + 1]) atomicExch(err, 2);
			else tw = tb + te;
		}
		he[h] = make_uint4(org + vb, tw, (h - b) | (deg << 16), f);
	}
}

// order[i] = fepair {u32 face; u16 edge} -> half-edge, vertex, and the vertex rank (first visit)
__global__ void __launch_bounds__(256) k_vertex_order(const uint32_t *__restrict__ order, const uint32_t *__restrict__ face_off, const uint4 *__restrict__ he, uint32_t n, SegView sv,
                                                      uint32_t *__restrict__ ord_h, uint32_t *__restrict__ ord_v, uint32_t *__restrict__ vrank, int *err)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint32_t s = hb_seg_find(sv.obase, sv.nseg, i);
	const uint32_t fb = sv.fbase[s], nf_s = sv.fbase[s + 1] - fb;
	const uint32_t fl = order[2 * (size_t)i], e = order[2 * (size_t)i + 1] & 0xffffu;
	const uint32_t f = fl + fb;
	// an invalid entry takes the first half-edge of its segment (any valid one keeps the later kernels in bounds)
	if (fl >= nf_s || face_off[f] + e >= face_off[f + 1]) { atomicExch(err, 3); ord_h[i] = sv.ebase[s]; ord_v[i] = sv.vbase[s]; return; }
	const uint32_t h = face_off[f] + e;
	const uint32_t v = he[h].x;
	ord_h[i] = h;
	ord_v[i] = v;
	atomicMin(&vrank[v], i);
}

// face order -> frank[f], gate half-edge per rank, degree per rank (for the corner-element scan)
__global__ void __launch_bounds__(256) k_face_order(const uint32_t *__restrict__ order_f, const uint32_t *__restrict__ face_off, uint32_t n, SegView sv,
                                                    uint32_t *__restrict__ frank, uint32_t *__restrict__ ford_h, uint32_t *__restrict__ fdeg, int *err)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint32_t s = hb_seg_find(sv.ofbase, sv.nseg, i);
	const uint32_t fb = sv.fbase[s], nf_s = sv.fbase[s + 1] - fb;
	uint32_t fl = i - sv.ofbase[s], e = 0; // no face order: faces in index order, gate corner 0 (the decoder, attrcode.h:543-548)
	if (order_f) {
		fl = order_f[2 * (size_t)i];
		e = order_f[2 * (size_t)i + 1] & 0xffffu;
	}
	const uint32_t f = fl + fb;
	if (fl >= nf_s || face_off[f] + e >= face_off[f + 1]) { atomicExch(err, 4); ford_h[i] = sv.ebase[s]; fdeg[i] = 0; return; }
	atomicMin(&frank[f], i);
	ford_h[i] = face_off[f] + e;
	fdeg[i] = face_off[f + 1] - face_off[f];
}

// corner elements in emission order: face rank fr, corners starting at the gate corner
// (attrcode.h:405-414); celem_h[ce] = half-edge, he_celem[h] = corner element.  Every face is traversed exactly
// once, so there are exactly ne corner elements; anything else is an invalid face order (flagged, never out of bounds:
// the arrays were cleared to element / half-edge 0).
__global__ void __launch_bounds__(256) k_corner_elems(const uint32_t *__restrict__ ford_h, const uint32_t *__restrict__ cbase, const uint4 *__restrict__ he, uint32_t n, uint32_t ne,
                                                      uint32_t *__restrict__ celem_h, uint32_t *__restrict__ he_celem, int *err)
{
	const uint32_t fr = blockIdx.x * blockDim.x + threadIdx.x;
	if (fr >= n) return;
	if (fr == 0 && cbase[n] != ne) atomicExch(err, 4);
	const uint32_t hg = ford_h[fr];
	const uint32_t ld = he[hg].z;
	const uint32_t deg = ld >> 16, base = cbase[fr];
	const uint32_t gate = ld & 0xffffu, f0 = hg - gate;
	if (cbase[fr + 1] - base != deg || base + deg > ne) { atomicExch(err, 4); return; } // a face listed twice / a gate outside its face
	for (uint32_t j = 0; j < deg; ++j) {
		uint32_t le = gate + j;
		if (le >= deg) le -= deg;
		celem_h[base + j] = f0 + le;
		he_celem[f0 + le] = base + j;
	}
}
// corner-element base of every segment
__global__ void k_seg_corner_base(const uint32_t *__restrict__ cbase, SegView sv, uint32_t *__restrict__ cebase)
{
	const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s <= sv.nseg) cebase[s] = cbase[sv.ofbase[s]];
}

// ------------------------------------------------------------------------------------------------
// fan walk (TFAN_IT, attrcode.h:83-106).  `visit(e, rec)` is called for every fan half-edge in the
// reference's order: forward over twin/next until back at the start or at a border (twin == self),
// then backward from prev(start) over twin/prev.
// ------------------------------------------------------------------------------------------------
template <typename Visit>
__device__ __forceinline__ bool fan_walk(const uint4 *__restrict__ he, uint32_t ein, uint32_t max_steps, Visit &&visit)
{
	const uint4 rin = he[ein];
	uint32_t e = ein, steps = 0;
	uint4 rec = rin;
	for (;;) {
		visit(e, rec);
		const uint32_t t = rec.y;
		if (t == e) break; // border -> backward
		const uint4 rt = he[t];
		e = he_next(t, rt.z);
		if (e == ein) return true;
		rec = he[e];
		if (++steps > max_steps) return false;
	}
	e = he_prev(ein, rin.z);
	rec = he[e];
	if (rec.y == e) return true;
	e = rec.y;
	do {
		rec = he[e];
		visit(e, rec);
		e = he_prev(e, rec.z);
		rec = he[e];
		if (rec.y == e) break;
		e = rec.y;
		if (++steps > max_steps) return false;
	} while (e != ein);
	return true;
}

// A candidate is kept when its three vertices were coded earlier and lie in the vertex's region (attrcode.h:117-121).
// The ranks are tested one by one AS SOON AS a vertex is known: most fan faces of a vertex fail on the first or second
// rank (their other corners come later in the traversal), and the kernel is bound by the number of scattered loads --
// a face that fails early never loads the rest of the neighbouring face's records nor the remaining ranks.
struct ParalSink {
	const uint32_t *vrank;
	const uint16_t *vtx_regs; // nullptr: one vertex region
	uint32_t self;   // traversal position of the vertex being coded
	uint16_t reg;
	uint32_t count;
	uint32_t *out;   // nullptr in the counting pass
	__device__ __forceinline__ void accept(uint32_t v0, uint32_t v1, uint32_t vo, uint32_t r0, uint32_t r1, uint32_t ro)
	{
		if (vtx_regs && (vtx_regs[v0] != reg || vtx_regs[v1] != reg || vtx_regs[vo] != reg)) return;
		if (out) {
			out[3 * (size_t)count] = r0;
			out[3 * (size_t)count + 1] = r1;
			out[3 * (size_t)count + 2] = ro;
		}
		++count;
	}
};

// Pass 1 (MODE_STAGE): one fan walk per traversed vertex; the first VC_STAGE accepted
// parallelograms are parked in a fixed-size staging slot, the count goes to cnt[].  After the scan
// of the counts k_vertex_candidates_compact streams the staged triples into the CSR; only vertices
// with more than VC_STAGE parallelograms (poles, closing vertices) walk their fan a second time
// (MODE_FILL, restricted to those vertices).
#define VC_STAGE 4
struct ParalStageSink {
	const uint32_t *vrank;
	const uint16_t *vtx_regs;
	uint32_t self;
	uint16_t reg;
	uint32_t count;
	uint32_t *stage; // VC_STAGE triples
	__device__ __forceinline__ void accept(uint32_t v0, uint32_t v1, uint32_t vo, uint32_t r0, uint32_t r1, uint32_t ro)
	{
		if (vtx_regs && (vtx_regs[v0] != reg || vtx_regs[v1] != reg || vtx_regs[vo] != reg)) return;
		if (count < VC_STAGE) {
			stage[3 * count] = r0;
			stage[3 * count + 1] = r1;
			stage[3 * count + 2] = ro;
		}
		++count;
	}
};
// attrcode.h:155-171 (paral) applied to fan half-edge e
template <typename Sink>
__device__ __forceinline__ void paral_visit_t(const uint4 *__restrict__ he, uint32_t e, const uint4 &rec, Sink &sink)
{
	const uint32_t deg = rec.z >> 16;
	if (deg == 3) {
		// the face across the edge opposite to the vertex: (org(t), org(next t), org(next next t))
		const uint32_t e1 = he_next(e, rec.z);
		const uint32_t t = he[e1].y;
		if (t == e1) return;
		const uint4 rt = he[t];
		const uint32_t r0 = sink.vrank[rt.x];
		if (r0 >= sink.self) return;
		const uint32_t tn = he_next(t, rt.z);
		const uint4 rtn = he[tn];
		const uint32_t r1 = sink.vrank[rtn.x];
		if (r1 >= sink.self) return;
		const uint32_t vo = he[he_next(tn, rtn.z)].x;
		const uint32_t ro = sink.vrank[vo];
		if (ro >= sink.self) return;
		sink.accept(rt.x, rtn.x, vo, r0, r1, ro);
		return;
	}
	const uint32_t e0 = he_next(e, rec.z), e1 = he_prev(e, rec.z);
	const uint4 q0 = he[e0];
	const uint32_t v0 = q0.x;
	const uint32_t r0 = sink.vrank[v0];
	if (r0 >= sink.self) return;
	const uint32_t v1 = he[e1].x;
	const uint32_t r1 = sink.vrank[v1];
	if (r1 >= sink.self) return;
	if (deg > 4) {
		// the second "parallelogram" of an n-gon degenerates to (v0, v1, v1) (Appendix C.3); it follows the regular one
		const uint32_t vo = he[he_next(e0, q0.z)].x;
		const uint32_t ro = sink.vrank[vo];
		if (ro < sink.self) sink.accept(v0, v1, vo, r0, r1, ro);
		sink.accept(v0, v1, v1, r0, r1, r1);
		return;
	}
	const uint32_t vo = he[he_next(e0, q0.z)].x;
	const uint32_t ro = sink.vrank[vo];
	if (ro >= sink.self) return;
	sink.accept(v0, v1, vo, r0, r1, ro);
}

// Wide fans.  A fan is a linked list (e -> next(twin(e))): a thread that has not closed it after
// VC_WALK_CAP steps hands the vertex over to the wide path, which collects the vertex's half-edges
// with one streaming pass, ranks the list by pointer doubling and evaluates the parallelograms of
// all fan positions in parallel -- a 4472-face sphere pole costs one thread ~5 ms, the wide path
// well under a millisecond.  A batch of meshes has many such vertices (two poles per sphere): each registers a
// slot, a per-vertex slot table routes the half-edges to their fan in the collecting pass.  More than VC_MAXWIDE
// such vertices: the rest walk sequentially.
#define VC_WALK_CAP 64
#define VC_MAXWIDE 4096
#define VC_WIDE_MARK 0xffffffffu
#define VC_WIDE_ARENA (1u << 21)  // fan nodes of all wide vertices of one launch (split evenly between them)
struct WideCtl {
	uint32_t n;                    // wide vertices registered (may exceed VC_MAXWIDE)
	uint32_t per;                  // arena nodes per wide vertex
	uint32_t vtx[VC_MAXWIDE], rank[VC_MAXWIDE], deg[VC_MAXWIDE], base[VC_MAXWIDE + 1], fill[VC_MAXWIDE], ncand[VC_MAXWIDE];
};

__global__ void __launch_bounds__(256) k_vertex_candidates_stage(const uint4 *__restrict__ he, const uint32_t *__restrict__ ord_h, const uint32_t *__restrict__ ord_v,
                                                                  const uint32_t *__restrict__ vrank, const uint16_t *__restrict__ vtx_regs, uint32_t n, uint32_t ne,
                                                                  uint32_t *__restrict__ cnt, uint32_t *__restrict__ stage, WideCtl *__restrict__ wide, uint32_t *__restrict__ vslot,
                                                                  uint32_t *__restrict__ wbits, uint32_t walk_cap, int *err)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	// a parallelogram needs three vertices coded earlier (two for the degenerate n-gon case): the
	// first two traversal positions cannot have any, whatever their fan looks like
	if (i < 2) { cnt[i] = 0; return; }
	ParalStageSink sink;
	sink.vrank = vrank;
	sink.vtx_regs = vtx_regs;
	sink.self = i;
	sink.reg = vtx_regs ? vtx_regs[ord_v[i]] : (uint16_t)0;
	sink.count = 0;
	uint32_t st[3 * VC_STAGE];
	sink.stage = st;
	bool ok = fan_walk(he, ord_h[i], walk_cap, [&](uint32_t e, const uint4 &rec) { paral_visit_t(he, e, rec, sink); });
	if (!ok) {
		const uint32_t slot = atomicAdd(&wide->n, 1u);
		if (slot < VC_MAXWIDE) {
			wide->vtx[slot] = ord_v[i];
			wide->rank[slot] = i;
			vslot[ord_v[i]] = slot;
			atomicOr(&wbits[ord_v[i] >> 5], 1u << (ord_v[i] & 31u));
			cnt[i] = 0; // set by k_wide_rank
			stage[(size_t)i * 3 * VC_STAGE] = VC_WIDE_MARK;
			return;
		}
		sink.count = 0;
		ok = fan_walk(he, ord_h[i], ne + 2, [&](uint32_t e, const uint4 &rec) { paral_visit_t(he, e, rec, sink); });
		if (!ok) atomicExch(err, 5);
	}
	cnt[i] = sink.count;
	if (sink.count) {
		uint4 *o = (uint4 *)(stage + (size_t)i * 3 * VC_STAGE);
		o[0] = make_uint4(st[0], st[1], st[2], st[3]);
		if (sink.count > 1) o[1] = make_uint4(st[4], st[5], st[6], st[7]);
		if (sink.count > 2) o[2] = make_uint4(st[8], st[9], st[10], st[11]);
	}
}

// Pass 2: staged triples -> CSR (streaming); vertices that overflowed the staging slot walk again
__global__ void __launch_bounds__(256) k_vertex_candidates_compact(const uint4 *__restrict__ he, const uint32_t *__restrict__ ord_h, const uint32_t *__restrict__ ord_v,
                                                                    const uint32_t *__restrict__ vrank, const uint16_t *__restrict__ vtx_regs, uint32_t n, uint32_t ne,
                                                                    const uint32_t *__restrict__ off, const uint32_t *__restrict__ stage, uint32_t *__restrict__ tri, uint32_t cap, int *err)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint32_t o0 = off[i], K = off[i + 1] - o0;
	if (K == 0) return;
	// every fan half-edge yields at most two candidates and belongs to one vertex: 2 ne triples hold any valid input
	if ((unsigned long long)o0 + K > cap) { atomicExch(err, 3); return; }
	if (stage[(size_t)i * 3 * VC_STAGE] == VC_WIDE_MARK) return; // written by k_wide_copy
	if (K <= VC_STAGE) {
		const uint32_t *st = stage + (size_t)i * 3 * VC_STAGE;
		for (uint32_t w = 0; w < 3 * K; ++w) tri[3 * (size_t)o0 + w] = st[w];
		return;
	}
	ParalSink sink;
	sink.vrank = vrank;
	sink.vtx_regs = vtx_regs;
	sink.self = i;
	sink.reg = vtx_regs ? vtx_regs[ord_v[i]] : (uint16_t)0;
	sink.count = 0;
	sink.out = tri + 3 * (size_t)o0;
	const bool ok = fan_walk(he, ord_h[i], ne + 2, [&](uint32_t e, const uint4 &rec) { paral_visit_t(he, e, rec, sink); });
	if (!ok) atomicExch(err, 5);
}

// ---- wide fans ---------------------------------------------------------------------------------
// half-edges whose origin is a wide vertex: count per vertex (SCATTER = false), then scatter into
// contiguous node lists and note every node's index (SCATTER = true)
template <bool SCATTER>
__global__ void __launch_bounds__(256) k_wide_collect(const uint4 *__restrict__ he, uint32_t ne, WideCtl *__restrict__ wide, const uint32_t *__restrict__ vslot,
                                                      const uint32_t *__restrict__ wbits, uint32_t *__restrict__ nodes, uint32_t *__restrict__ pos)
{
	const uint32_t nw = min(wide->n, (uint32_t)VC_MAXWIDE);
	if (nw == 0) return; // the usual mesh: no wide fan, nothing to stream
	const uint32_t cap_per_slot = wide->per;
	for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += gridDim.x * blockDim.x) {
		// one bit per vertex in front of the slot table: the bitmap of a 10M-vertex mesh is 1.2 MB and stays in cache, the
		// slot table (4 bytes per vertex) would be gathered from L2 for every half-edge
		// (a packed copy of the origins, written by K0 for this pass, cost K0 four times what it saved here)
		const uint32_t v = __ldg((const uint32_t *)(he + e));
		if (!((__ldg(wbits + (v >> 5)) >> (v & 31u)) & 1u)) continue;
		const uint32_t w = vslot[v];
		if (w == HB_NONE) continue;
		if (!SCATTER) atomicAdd(&wide->deg[w], 1u);
		else {
			const uint32_t k = atomicAdd(&wide->fill[w], 1u);
			if (k >= cap_per_slot) continue; // fan larger than its share of the arena: k_wide_rank walks it sequentially
			const uint32_t idx = wide->base[w] + k;
			nodes[idx] = e;
			pos[e] = idx;
		}
	}
}
// evenly split arena: slot w owns [w * per, (w + 1) * per)
__global__ void k_wide_even_bases(WideCtl *wide)
{
	const uint32_t nw = min(wide->n, (uint32_t)VC_MAXWIDE);
	const uint32_t per = VC_WIDE_ARENA / (nw ? nw : 1u);
	if (threadIdx.x == 0) wide->per = per;
	for (uint32_t w = threadIdx.x; w <= nw; w += blockDim.x) wide->base[w] = w * per;
	for (uint32_t w = threadIdx.x; w < nw; w += blockDim.x) wide->fill[w] = 0;
}
__global__ void k_wide_fill_to_deg(WideCtl *wide)
{
	const uint32_t nw = min(wide->n, (uint32_t)VC_MAXWIDE);
	for (uint32_t w = threadIdx.x; w < nw; w += blockDim.x) wide->deg[w] = wide->fill[w];
}

// One CTA per wide vertex: rank its fan by pointer doubling (distance to the tail of the
// e -> next(twin(e)) path, the path cut in front of the gate half-edge when the fan is closed),
// turn the ranks into the visiting order of fan_walk (forward from the gate, then backward from
// the gate's predecessor), and evaluate paral() for every position in parallel.  Triples go to
// `arena` (two per node at most) in visiting order; the count goes to cnt[rank].
#define WIDE_T 1024
__global__ void __launch_bounds__(WIDE_T) k_wide_rank(const uint4 *__restrict__ he, const uint32_t *__restrict__ ord_h, const uint32_t *__restrict__ vrank,
                                                      const uint16_t *__restrict__ vtx_regs, WideCtl *__restrict__ wide, const uint32_t *__restrict__ nodes,
                                                      const uint32_t *__restrict__ pos, uint32_t *nxtA, uint32_t *nxtB, uint32_t *dA, uint32_t *dB, uint32_t *lastA, uint32_t *lastB,
                                                      uint32_t *__restrict__ order, uint32_t *__restrict__ arena, uint32_t *__restrict__ cnt, uint32_t *__restrict__ stage, uint32_t ne, int *err)
{
	const uint32_t NIL = 0xffffffffu;
	const uint32_t slot = blockIdx.x, t = threadIdx.x;
	if (slot >= min(wide->n, (uint32_t)VC_MAXWIDE)) return;
	const uint32_t b = wide->base[slot], d = wide->deg[slot], self = wide->rank[slot];
	const uint32_t ein = ord_h[self];
	const uint16_t reg = vtx_regs ? vtx_regs[wide->vtx[slot]] : (uint16_t)0;
	__shared__ uint32_t s_warp[32], s_carry, s_total;
	if (d > wide->per) {
		// the fan does not fit its share of the arena (more than VC_WIDE_ARENA / #wide vertices half-edges around one
		// vertex): one thread walks it in order, like the stage kernel does for the vertices beyond VC_MAXWIDE
		if (t == 0) {
			ParalStageSink sink;
			sink.vrank = vrank; sink.vtx_regs = vtx_regs; sink.self = self; sink.reg = reg; sink.count = 0;
			uint32_t st[3 * VC_STAGE];
			sink.stage = st;
			if (!fan_walk(he, ein, ne + 2, [&](uint32_t e, const uint4 &rec) { paral_visit_t(he, e, rec, sink); })) atomicExch(err, 5);
			cnt[self] = sink.count;
			for (uint32_t w = 0; w < 3 * VC_STAGE; ++w) stage[(size_t)self * 3 * VC_STAGE + w] = w < 3 * min(sink.count, (uint32_t)VC_STAGE) ? st[w] : 0u; // clears the wide mark
			wide->ncand[slot] = 0;
		}
		return;
	}
	for (uint32_t k = t; k < d; k += WIDE_T) {
		const uint32_t e = nodes[b + k];
		const uint4 rec = he[e];
		uint32_t sidx = NIL;
		if (rec.y != e) {
			const uint32_t e2 = he_next(rec.y, he[rec.y].z);
			if (e2 != ein) sidx = pos[e2] - b;
		}
		nxtA[b + k] = sidx;
		dA[b + k] = sidx == NIL ? 0u : 1u;
		lastA[b + k] = k;
		order[b + k] = NIL;
	}
	__syncthreads();
	for (uint32_t span = 1; span < d; span <<= 1) {
		for (uint32_t k = t; k < d; k += WIDE_T) {
			const uint32_t sidx = nxtA[b + k];
			if (sidx != NIL && sidx < d) {
				dB[b + k] = dA[b + k] + dA[b + sidx];
				nxtB[b + k] = nxtA[b + sidx];
				lastB[b + k] = lastA[b + sidx];
			} else {
				dB[b + k] = dA[b + k];
				nxtB[b + k] = NIL;
				lastB[b + k] = lastA[b + k];
			}
		}
		__syncthreads();
		uint32_t *tp;
		tp = nxtA; nxtA = nxtB; nxtB = tp;
		tp = dA; dA = dB; dB = tp;
		tp = lastA; lastA = lastB; lastB = tp;
	}
	const uint32_t kin = pos[ein] - b;
	if (kin >= d || nxtA[b + kin] != NIL) { if (t == 0) atomicExch(err, 5); return; } // the gate's path does not end: inconsistent twins
	const uint32_t D = dA[b + kin], tail = lastA[b + kin];
	// visiting position: nodes behind the gate D - dtail, nodes in front of it (open fan) dtail
	for (uint32_t k = t; k < d; k += WIDE_T) {
		if (lastA[b + k] != tail) continue; // another fan of a non-manifold vertex
		const uint32_t dt = dA[b + k];
		const uint32_t p = dt <= D ? D - dt : dt;
		if (p < d) order[b + p] = nodes[b + k];
	}
	if (t == 0) s_carry = 0;
	__syncthreads();
	for (uint32_t p0 = 0; p0 < d; p0 += WIDE_T) {
		const uint32_t p = p0 + t;
		ParalStageSink sink;
		sink.vrank = vrank;
		sink.vtx_regs = vtx_regs;
		sink.self = self;
		sink.reg = reg;
		sink.count = 0;
		uint32_t st[3 * VC_STAGE];
		sink.stage = st;
		if (p < d) {
			const uint32_t e = order[b + p];
			if (e != NIL) paral_visit_t(he, e, he[e], sink);
		}
		// exclusive scan of the per-position counts (0..2)
		const uint32_t c = sink.count;
		uint32_t x = c;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const uint32_t up = __shfl_up_sync(0xffffffffu, x, o);
			if ((int)(t & 31) >= o) x += up;
		}
		if ((t & 31) == 31) s_warp[t >> 5] = x;
		__syncthreads();
		if (t < 32) {
			uint32_t w = s_warp[t];
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) {
				const uint32_t up = __shfl_up_sync(0xffffffffu, w, o);
				if ((int)t >= o) w += up;
			}
			s_warp[t] = w;
			if (t == 31) s_total = w;
		}
		__syncthreads();
		const uint32_t excl = s_carry + x - c + ((t >> 5) ? s_warp[(t >> 5) - 1] : 0u);
		for (uint32_t j = 0; j < c; ++j) {
			uint32_t *o = arena + 3 * (size_t)(2 * b + excl + j);
			o[0] = st[3 * j]; o[1] = st[3 * j + 1]; o[2] = st[3 * j + 2];
		}
		__syncthreads();
		if (t == 0) s_carry += s_total;
		__syncthreads();
	}
	if (t == 0) { wide->ncand[slot] = s_carry; cnt[self] = s_carry; }
}
__global__ void __launch_bounds__(256) k_wide_copy(const WideCtl *__restrict__ wide, const uint32_t *__restrict__ off, const uint32_t *__restrict__ arena, uint32_t *__restrict__ tri, uint32_t cap)
{
	const uint32_t slot = blockIdx.x;
	if (slot >= min(wide->n, (uint32_t)VC_MAXWIDE)) return;
	const uint32_t nc = wide->ncand[slot], b = wide->base[slot], o0 = off[wide->rank[slot]];
	if ((unsigned long long)o0 + nc > cap) return; // flagged by k_vertex_candidates_compact
	for (uint32_t w = threadIdx.x; w < 3 * nc; w += blockDim.x) tri[3 * (size_t)o0 + w] = arena[3 * (size_t)(2 * b) + w];
}

// corner candidates: fan faces coded earlier (face rank smaller) in the same face region
// (attrcode.h:135-154, 272-288); candidate = corner element of that fan half-edge
template <bool FILL>
__global__ void __launch_bounds__(256) k_corner_candidates(const uint4 *__restrict__ he, const uint32_t *__restrict__ celem_h, const uint32_t *__restrict__ he_celem,
                                                            const uint32_t *__restrict__ frank, const uint16_t *__restrict__ face_regs, const int *__restrict__ reg_ncorner,
                                                            uint32_t n, uint32_t ne, uint32_t *__restrict__ cnt_or_off, uint32_t *__restrict__ idx, int *err)
{
	const uint32_t ce = blockIdx.x * blockDim.x + threadIdx.x;
	if (ce >= n) return;
	const uint32_t h = celem_h[ce];
	const uint32_t f = he[h].w;
	const uint16_t reg = face_regs[f];
	uint32_t count = 0;
	if (reg_ncorner[reg] > 0) { // Appendix C.9: the reference walks the fan anyway; no output effect
		const uint32_t fr = frank[f];
		uint32_t *out = FILL ? idx + cnt_or_off[ce] : nullptr;
		const bool ok = fan_walk(he, h, ne + 2, [&](uint32_t e, const uint4 &rec) {
			const uint32_t f2 = rec.w;
			if (frank[f2] >= fr || face_regs[f2] != reg) return;
			if (FILL) out[count] = he_celem[e];
			++count;
		});
		if (!ok) atomicExch(err, 5);
	}
	if (!FILL) cnt_or_off[ce] = count;
}

// ------------------------------------------------------------------------------------------------
// host drivers
// ------------------------------------------------------------------------------------------------
static SegView seg_view(const hb_dmesh *m)
{
	SegView sv;
	sv.nseg = m->nseg;
	sv.vbase = m->d_vbase; sv.fbase = m->d_fbase; sv.ebase = m->d_ebase; sv.obase = m->d_obase; sv.ofbase = m->d_ofbase;
	return sv;
}

int hb_build_conn(hb_dmesh *m)
{
	hb_ctx *ctx = m->ctx;
	if (m->conn_ready) return 0;
	const SegView sv = seg_view(m);
	HB_TRY(hb_dalloc_t(m, &m->d_he, (size_t)m->ne + 1));
	HB_TRY(hb_dalloc_t(m, &m->d_vrank, (size_t)m->nv + 1));
	HB_TRY(hb_dalloc_t(m, &m->d_ord_h, (size_t)m->norder + 1));
	HB_TRY(hb_dalloc_t(m, &m->d_ord_v, (size_t)m->norder + 1));
	HB_TRY(hb_dalloc_t(m, &m->d_frank, (size_t)m->nf + 1));
	HB_TRY(hb_dalloc_t(m, &m->d_ford_h, (size_t)m->norder_f + 1));
	HB_TRY(hb_dalloc_t(m, &m->d_cbase, (size_t)m->norder_f + 2));
	HB_CUDA(ctx, cudaMemsetAsync(m->d_vrank, 0xff, sizeof(uint32_t) * ((size_t)m->nv + 1), ctx->stream));
	HB_CUDA(ctx, cudaMemsetAsync(m->d_frank, 0xff, sizeof(uint32_t) * ((size_t)m->nf + 1), ctx->stream));
	if (m->nseg > 1) {
		HB_TRY(hb_dalloc_t(m, &m->d_face_off, (size_t)m->nf + 1));
		HB_LAUNCH(ctx, k_global_face_off, hb_div_up((uint64_t)m->nf + 1, 256 * FOFF_PER_THREAD), 256, 0, m->d_face_off_raw, sv, m->nf, m->d_face_off);
	}
	if (m->nf) HB_LAUNCH(ctx, k_flatten_halfedges, hb_div_up(m->nf, 256), 256, 0, (const uint32_t *)m->d_edges_raw, m->d_face_off, m->d_he, m->nf, m->ne, sv, ctx->d_err);
	if (m->norder)
		HB_LAUNCH(ctx, k_vertex_order, hb_div_up(m->norder, 256), 256, 0, (const uint32_t *)m->d_order, m->d_face_off, m->d_he, m->norder, sv, m->d_ord_h, m->d_ord_v, m->d_vrank, ctx->d_err);
	// face ranks and gate half-edges are only needed by face lists with components (or partly bound ones), corner lists
	// and the face-region stream: the zero-component face list of a plain PLY is coded straight from order_f
	bool need_face_order = m->any_corner || m->nregs_face > 1;
	for (int l = 0; l < m->nlists; ++l) {
		const ListParams &p = m->lists[l].p;
		if (p.target != HB_FACE) continue;
		bool everywhere = true;
		for (int r = 0; r < m->nregs_face; ++r) everywhere = everywhere && m->h_slot_face[(size_t)r * m->nlists + l] >= 0;
		if (p.ncomp > 0 || !everywhere) need_face_order = true;
	}
	if (m->norder_f && need_face_order)
		HB_LAUNCH(ctx, k_face_order, hb_div_up(m->norder_f, 256), 256, 0, m->has_order_f ? (const uint32_t *)m->d_order_f : (const uint32_t *)nullptr, m->d_face_off, m->norder_f, sv, m->d_frank, m->d_ford_h, m->d_cbase, ctx->d_err);
	// corner elements are only materialized when some region binds corner lists
	m->n_corner_elems = 0;
	if (m->any_corner && m->norder_f) {
		HB_TRY(hb_scan_exclusive_u32(ctx, m->d_cbase, m->d_cbase, m->norder_f, nullptr));
		m->n_corner_elems = m->ne; // every face is traversed once (k_corner_elems flags anything else): no read-back
		HB_TRY(hb_dalloc_t(m, &m->d_celem_h, (size_t)m->ne + 1));
		HB_TRY(hb_dalloc_t(m, &m->d_he_celem, (size_t)m->ne + 1));
		HB_TRY(hb_dalloc_t(m, &m->d_cebase, (size_t)m->nseg + 1));
		HB_CUDA(ctx, cudaMemsetAsync(m->d_celem_h, 0, sizeof(uint32_t) * ((size_t)m->ne + 1), ctx->stream));
		HB_CUDA(ctx, cudaMemsetAsync(m->d_he_celem, 0, sizeof(uint32_t) * ((size_t)m->ne + 1), ctx->stream));
		HB_LAUNCH(ctx, k_corner_elems, hb_div_up(m->norder_f, 256), 256, 0, m->d_ford_h, m->d_cbase, m->d_he, m->norder_f, m->ne, m->d_celem_h, m->d_he_celem, ctx->d_err);
		HB_LAUNCH(ctx, k_seg_corner_base, hb_div_up(m->nseg + 1, 128), 128, 0, m->d_cbase, sv, m->d_cebase);
		HB_TRY(hb_dalloc_t(m, &m->d_reg_ncorner, m->reg_ncorner.size() + 1));
		HB_CUDA(ctx, cudaMemcpyAsync(m->d_reg_ncorner, m->reg_ncorner.data(), sizeof(int) * m->reg_ncorner.size(), cudaMemcpyHostToDevice, ctx->stream));
	}
	m->conn_ready = true;
	return 0;
}

// No host synchronisation: the wide-fan kernels are launched unconditionally and return at once when the stage
// pass registered no wide vertex; the CSR of triples is allocated at its upper bound (two candidates per fan
// half-edge, every half-edge in the fan of one vertex).
int hb_build_vertex_candidates(hb_dmesh *m)
{
	hb_ctx *ctx = m->ctx;
	if (m->vcand_ready) return 0;
	HB_TRY(hb_build_conn(m));
	const uint32_t n = m->norder;
	HB_TRY(hb_dalloc_t(m, &m->d_vc_off, (size_t)n + 2));
	m->vc_total = 0;
	if (n) {
		HB_TRY(hb_dalloc_t(m, &m->d_vc_stage, (size_t)n * 3 * VC_STAGE + 4));
		uint32_t *stage = m->d_vc_stage;
		HB_TRY(hb_dalloc(m, &m->d_vc_wide, sizeof(WideCtl)));
		WideCtl *wide = (WideCtl *)m->d_vc_wide;
		HB_CUDA(ctx, cudaMemsetAsync(wide, 0, sizeof(WideCtl), ctx->stream));
		HB_TRY(hb_dalloc_t(m, &m->d_vc_wslot, (size_t)m->nv + 1));
		HB_TRY(hb_dalloc_t(m, &m->d_vc_wbits, ((size_t)m->nv >> 5) + 2));
		// (the slot table is only read where the bitmap says a slot was written: no need to clear its 4 bytes per vertex)
		HB_CUDA(ctx, cudaMemsetAsync(m->d_vc_wbits, 0, sizeof(uint32_t) * (((size_t)m->nv >> 5) + 2), ctx->stream));
		const uint16_t *vregs = m->nregs_vtx > 1 ? m->d_vtx_regs : nullptr; // one vertex region: no region test, no region gathers
		HB_LAUNCH(ctx, k_vertex_candidates_stage, hb_div_up(n, 256), 256, 0, m->d_he, m->d_ord_h, m->d_ord_v, m->d_vrank, vregs, n, m->ne, m->d_vc_off, stage, wide,
		          m->d_vc_wslot, m->d_vc_wbits, (uint32_t)VC_WALK_CAP, ctx->d_err);
		const size_t cap = VC_WIDE_ARENA;
		m->vc_wide_cap = (uint32_t)cap;
		HB_TRY(hb_dalloc_t(m, &m->d_vc_wpos, (size_t)m->ne + 1));
		HB_TRY(hb_dalloc_t(m, &m->d_vc_wnodes, cap + 1));
		HB_TRY(hb_dalloc_t(m, &m->d_vc_wwork, 6 * cap + 6));
		HB_TRY(hb_dalloc_t(m, &m->d_vc_worder, cap + 1));
		HB_TRY(hb_dalloc_t(m, &m->d_vc_warena, 6 * cap + 6));
		uint32_t *nodes = m->d_vc_wnodes, *pos = m->d_vc_wpos, *work = m->d_vc_wwork, *order = m->d_vc_worder, *arena = m->d_vc_warena;
		HB_LAUNCH(ctx, k_wide_even_bases, 1, 256, 0, wide);
		HB_LAUNCH(ctx, k_wide_collect<true>, (uint32_t)ctx->sm_count * 8, 256, 0, m->d_he, m->ne, wide, m->d_vc_wslot, m->d_vc_wbits, nodes, pos);
		HB_LAUNCH(ctx, k_wide_fill_to_deg, 1, 256, 0, wide);
		HB_LAUNCH(ctx, k_wide_rank, VC_MAXWIDE, WIDE_T, 0, m->d_he, m->d_ord_h, m->d_vrank, vregs, wide, nodes, pos, work, work + cap, work + 2 * cap, work + 3 * cap,
		          work + 4 * cap, work + 5 * cap, order, arena, m->d_vc_off, stage, m->ne, ctx->d_err);
		HB_TRY(hb_scan_exclusive_u32(ctx, m->d_vc_off, m->d_vc_off, n, nullptr));
		const uint64_t tri_cap = 2 * (uint64_t)m->ne + 8;
		if (tri_cap > 0xffffffffull) return hb_fail(ctx, HB_ERR_UNSUPPORTED, "mesh (or batch) with more than 2^31 half-edges");
		m->vc_total = (uint32_t)tri_cap;
		HB_TRY(hb_dalloc_t(m, &m->d_vc_tri, 3 * (size_t)tri_cap + 3));
		HB_LAUNCH(ctx, k_vertex_candidates_compact, hb_div_up(n, 256), 256, 0, m->d_he, m->d_ord_h, m->d_ord_v, m->d_vrank, vregs, n, m->ne, m->d_vc_off, stage, m->d_vc_tri, (uint32_t)tri_cap, ctx->d_err);
		HB_LAUNCH(ctx, k_wide_copy, VC_MAXWIDE, 256, 0, wide, m->d_vc_off, arena, m->d_vc_tri, (uint32_t)tri_cap);
	} else {
		HB_CUDA(ctx, cudaMemsetAsync(m->d_vc_off, 0, sizeof(uint32_t) * 2, ctx->stream));
	}
	m->vcand_ready = true;
	return 0;
}

int hb_build_corner_candidates(hb_dmesh *m)
{
	hb_ctx *ctx = m->ctx;
	if (m->ccand_ready) return 0;
	HB_TRY(hb_build_conn(m));
	const uint32_t n = m->n_corner_elems;
	HB_TRY(hb_dalloc_t(m, &m->d_cc_off, (size_t)n + 2));
	m->cc_total = 0;
	if (n) {
		int *d_ncorner = m->d_reg_ncorner;
		HB_LAUNCH(ctx, k_corner_candidates<false>, hb_div_up(n, 256), 256, 0, m->d_he, m->d_celem_h, m->d_he_celem, m->d_frank, m->d_face_regs, d_ncorner, n, m->ne, m->d_cc_off, (uint32_t *)nullptr, ctx->d_err);
		HB_TRY(hb_scan_exclusive_u32(ctx, m->d_cc_off, m->d_cc_off, n, nullptr));
		HB_CUDA(ctx, cudaMemcpyAsync(&m->cc_total, m->d_cc_off + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
		HB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
		HB_TRY(hb_check_device_error(ctx, "corner fan walk"));
		HB_TRY(hb_dalloc_t(m, &m->d_cc_idx, (size_t)m->cc_total + 1));
		HB_LAUNCH(ctx, k_corner_candidates<true>, hb_div_up(n, 256), 256, 0, m->d_he, m->d_celem_h, m->d_he_celem, m->d_frank, m->d_face_regs, d_ncorner, n, m->ne, m->d_cc_off, m->d_cc_idx, ctx->d_err);
	} else {
		HB_CUDA(ctx, cudaMemsetAsync(m->d_cc_off, 0, sizeof(uint32_t) * 2, ctx->stream));
	}
	m->ccand_ready = true;
	return 0;
}
