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
// exclusive scan (3-phase), out[n] receives the total
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

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, uint32_t *__restrict__ tile_sums, uint32_t n)
{
	const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
	uint32_t v[SCAN_ITEMS], s = 0;
#pragma unroll
	for (int k = 0; k < SCAN_ITEMS; ++k) {
		v[k] = base + k < n ? in[base + k] : 0;
		s += v[k];
	}
	uint32_t total;
	uint32_t ex = block_exclusive_scan(s, &total);
#pragma unroll
	for (int k = 0; k < SCAN_ITEMS; ++k) {
		if (base + k < n) out[base + k] = ex;
		ex += v[k];
	}
	if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_sums(uint32_t *__restrict__ tile_sums, uint32_t ntiles, uint32_t *__restrict__ out_total, uint32_t *__restrict__ out_total2)
{
	uint32_t carry = 0;
	for (uint32_t base = 0; base < ntiles; base += SCAN_THREADS) {
		const uint32_t i = base + threadIdx.x;
		const uint32_t v = i < ntiles ? tile_sums[i] : 0;
		uint32_t total;
		const uint32_t ex = block_exclusive_scan(v, &total);
		if (i < ntiles) tile_sums[i] = carry + ex;
		carry += total;
	}
	if (threadIdx.x == 0) {
		*out_total = carry;
		if (out_total2) *out_total2 = carry;
	}
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_add(uint32_t *__restrict__ out, const uint32_t *__restrict__ tile_sums, uint32_t n)
{
	const uint32_t add = tile_sums[blockIdx.x];
	const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
#pragma unroll
	for (int k = 0; k < SCAN_ITEMS; ++k)
		if (base + k < n) out[base + k] += add;
}

// d_out must hold n + 1 entries; d_out[n] = total.  in == out is allowed.
int hb_scan_exclusive_u32(hb_ctx *ctx, const uint32_t *d_in, uint32_t *d_out, uint32_t n, uint32_t *d_total)
{
	const uint32_t ntiles = n ? hb_div_up(n, SCAN_TILE) : 0;
	uint32_t *d_sums = nullptr;
	HB_CUDA(ctx, cudaMallocAsync((void **)&d_sums, sizeof(uint32_t) * (ntiles + 1), ctx->stream));
	if (ntiles) HB_LAUNCH(ctx, k_scan_tiles, ntiles, SCAN_THREADS, 0, d_in, d_out, d_sums, n);
	HB_LAUNCH(ctx, k_scan_sums, 1, SCAN_THREADS, 0, d_sums, ntiles, d_out + n, d_total);
	if (ntiles > 1) HB_LAUNCH(ctx, k_scan_add, ntiles, SCAN_THREADS, 0, d_out, d_sums, n);
	HB_CUDA(ctx, cudaFreeAsync(d_sums, ctx->stream));
	return 0;
}

// ------------------------------------------------------------------------------------------------
// K0: raw 12-byte {org, twin_face, twin_edge:16} records -> uint4 {org, twin, le | deg << 16, face}
// ------------------------------------------------------------------------------------------------
__global__ void k_flatten_halfedges(const uint32_t *__restrict__ raw, const uint32_t *__restrict__ face_off, uint4 *__restrict__ he, uint32_t nf, uint32_t nv, int *err)
{
	const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
	if (f >= nf) return;
	const uint32_t b = face_off[f], e = face_off[f + 1];
	const uint32_t deg = e - b;
	if (deg > 0xffffu || e < b) { atomicExch(err, 1); return; }
	for (uint32_t h = b; h < e; ++h) {
		const uint32_t org = raw[3 * (size_t)h], tf = raw[3 * (size_t)h + 1], te = raw[3 * (size_t)h + 2] & 0xffffu;
		uint32_t tw = h;
		if (org >= nv || tf >= nf) {
			atomicExch(err, 2);
		} else {
			const uint32_t tb = face_off[tf];
			if (tb + te >= face_off[tf + 1]) atomicExch(err, 2);
			else tw = tb + te;
		}
		he[h] = make_uint4(org, tw, (h - b) | (deg << 16), f);
	}
}

// order[i] = fepair {u32 face; u16 edge} -> half-edge, vertex, and the vertex rank (first visit)
__global__ void k_vertex_order(const uint32_t *__restrict__ order, const uint32_t *__restrict__ face_off, const uint4 *__restrict__ he, uint32_t n, uint32_t nf,
                               uint32_t *__restrict__ ord_h, uint32_t *__restrict__ ord_v, uint32_t *__restrict__ vrank, int *err)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint32_t f = order[2 * (size_t)i], e = order[2 * (size_t)i + 1] & 0xffffu;
	if (f >= nf || face_off[f] + e >= face_off[f + 1]) { atomicExch(err, 3); ord_h[i] = 0; ord_v[i] = 0; return; }
	const uint32_t h = face_off[f] + e;
	const uint32_t v = he[h].x;
	ord_h[i] = h;
	ord_v[i] = v;
	atomicMin(&vrank[v], i);
}

// face order -> frank[f], gate half-edge per rank, degree per rank (for the corner-element scan)
__global__ void k_face_order(const uint32_t *__restrict__ order_f, const uint32_t *__restrict__ face_off, uint32_t n, uint32_t nf,
                             uint32_t *__restrict__ frank, uint32_t *__restrict__ ford_h, uint32_t *__restrict__ fdeg, int *err)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint32_t f = i, e = 0;
	if (order_f) {
		f = order_f[2 * (size_t)i];
		e = order_f[2 * (size_t)i + 1] & 0xffffu;
	}
	if (f >= nf || face_off[f] + e >= face_off[f + 1]) { atomicExch(err, 4); ford_h[i] = 0; fdeg[i] = 0; return; }
	atomicMin(&frank[f], i);
	ford_h[i] = face_off[f] + e;
	fdeg[i] = face_off[f + 1] - face_off[f];
}

// corner elements in emission order: face rank fr, corners starting at the gate corner
// (attrcode.h:405-414); celem_h[ce] = half-edge, he_celem[h] = corner element
__global__ void k_corner_elems(const uint32_t *__restrict__ ford_h, const uint32_t *__restrict__ cbase, const uint4 *__restrict__ he, uint32_t n,
                               uint32_t *__restrict__ celem_h, uint32_t *__restrict__ he_celem)
{
	const uint32_t fr = blockIdx.x * blockDim.x + threadIdx.x;
	if (fr >= n) return;
	const uint32_t hg = ford_h[fr];
	const uint32_t ld = he[hg].z;
	const uint32_t deg = ld >> 16, base = cbase[fr];
	const uint32_t gate = ld & 0xffffu, f0 = hg - gate;
	for (uint32_t j = 0; j < deg; ++j) {
		uint32_t le = gate + j;
		if (le >= deg) le -= deg;
		celem_h[base + j] = f0 + le;
		he_celem[f0 + le] = base + j;
	}
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

struct ParalSink {
	const uint32_t *vrank;
	const uint16_t *vtx_regs;
	uint32_t self;   // traversal position of the vertex being coded
	uint16_t reg;
	uint32_t count;
	uint32_t *out;   // nullptr in the counting pass
	// attrcode.h:117-121: all three vertices coded earlier and in the same region
	__device__ __forceinline__ void offer(uint32_t v0, uint32_t v1, uint32_t vo)
	{
		const uint32_t r0 = vrank[v0], r1 = vrank[v1], ro = vrank[vo];
		if (r0 >= self || r1 >= self || ro >= self) return;
		if (vtx_regs[v0] != reg || vtx_regs[v1] != reg || vtx_regs[vo] != reg) return;
		if (out) {
			out[3 * (size_t)count] = r0;
			out[3 * (size_t)count + 1] = r1;
			out[3 * (size_t)count + 2] = ro;
		}
		++count;
	}
};

// attrcode.h:155-171 (paral) applied to fan half-edge e
__device__ __forceinline__ void paral_visit(const uint4 *__restrict__ he, uint32_t e, const uint4 &rec, ParalSink &sink)
{
	const uint32_t deg = rec.z >> 16;
	if (deg == 3) {
		const uint32_t e1 = he_next(e, rec.z);
		const uint32_t t = he[e1].y;
		if (t == e1) return;
		const uint4 rt = he[t];
		const uint32_t tn = he_next(t, rt.z);
		const uint4 rtn = he[tn];
		const uint32_t tnn = he_next(tn, rtn.z);
		sink.offer(rt.x, rtn.x, he[tnn].x);
		return;
	}
	const uint32_t e0 = he_next(e, rec.z), e1 = he_prev(e, rec.z);
	const uint4 r0 = he[e0];
	const uint32_t v0 = r0.x, v1 = he[e1].x;
	sink.offer(v0, v1, he[he_next(e0, r0.z)].x);
	if (deg > 4) sink.offer(v0, v1, v1); // second "parallelogram" of an n-gon degenerates (Appendix C.3)
}

template <bool FILL>
__global__ void __launch_bounds__(256) k_vertex_candidates(const uint4 *__restrict__ he, const uint32_t *__restrict__ ord_h, const uint32_t *__restrict__ ord_v,
                                                            const uint32_t *__restrict__ vrank, const uint16_t *__restrict__ vtx_regs, uint32_t n, uint32_t ne,
                                                            uint32_t *__restrict__ cnt_or_off, uint32_t *__restrict__ tri, int *err)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	ParalSink sink;
	sink.vrank = vrank;
	sink.vtx_regs = vtx_regs;
	sink.self = i;
	sink.reg = vtx_regs[ord_v[i]];
	sink.count = 0;
	sink.out = FILL ? tri + 3 * (size_t)cnt_or_off[i] : nullptr;
	// a vertex visited twice would see itself as coded the second time; the reference marks a
	// vertex coded after its first visit (attrcode.h:218), which vrank (first visit) reproduces.
	const bool ok = fan_walk(he, ord_h[i], ne + 2, [&](uint32_t e, const uint4 &rec) { paral_visit(he, e, rec, sink); });
	if (!ok) atomicExch(err, 5);
	if (!FILL) cnt_or_off[i] = sink.count;
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
int hb_build_conn(hb_dmesh *m)
{
	hb_ctx *ctx = m->ctx;
	if (m->conn_ready) return 0;
	HB_TRY(hb_dalloc_t(m, &m->d_he, (size_t)m->ne + 1));
	HB_TRY(hb_dalloc_t(m, &m->d_vrank, (size_t)m->nv + 1));
	HB_TRY(hb_dalloc_t(m, &m->d_ord_h, (size_t)m->norder + 1));
	HB_TRY(hb_dalloc_t(m, &m->d_ord_v, (size_t)m->norder + 1));
	HB_TRY(hb_dalloc_t(m, &m->d_frank, (size_t)m->nf + 1));
	HB_TRY(hb_dalloc_t(m, &m->d_ford_h, (size_t)m->norder_f + 1));
	HB_TRY(hb_dalloc_t(m, &m->d_cbase, (size_t)m->norder_f + 2));
	HB_CUDA(ctx, cudaMemsetAsync(m->d_vrank, 0xff, sizeof(uint32_t) * ((size_t)m->nv + 1), ctx->stream));
	HB_CUDA(ctx, cudaMemsetAsync(m->d_frank, 0xff, sizeof(uint32_t) * ((size_t)m->nf + 1), ctx->stream));
	if (m->nf) HB_LAUNCH(ctx, k_flatten_halfedges, hb_div_up(m->nf, 256), 256, 0, (const uint32_t *)m->d_edges_raw, m->d_face_off, m->d_he, m->nf, m->nv, ctx->d_err);
	if (m->norder)
		HB_LAUNCH(ctx, k_vertex_order, hb_div_up(m->norder, 256), 256, 0, (const uint32_t *)m->d_order, m->d_face_off, m->d_he, m->norder, m->nf, m->d_ord_h, m->d_ord_v, m->d_vrank, ctx->d_err);
	if (m->norder_f)
		HB_LAUNCH(ctx, k_face_order, hb_div_up(m->norder_f, 256), 256, 0, m->has_order_f ? (const uint32_t *)m->d_order_f : (const uint32_t *)nullptr, m->d_face_off, m->norder_f, m->nf, m->d_frank, m->d_ford_h, m->d_cbase, ctx->d_err);
	// corner elements are only materialized when some region binds corner lists
	m->n_corner_elems = 0;
	if (m->any_corner && m->norder_f) {
		HB_TRY(hb_scan_exclusive_u32(ctx, m->d_cbase, m->d_cbase, m->norder_f, nullptr));
		uint32_t total = 0;
		HB_CUDA(ctx, cudaMemcpyAsync(&total, m->d_cbase + m->norder_f, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
		HB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
		HB_TRY(hb_check_device_error(ctx, "connectivity"));
		m->n_corner_elems = total;
		HB_TRY(hb_dalloc_t(m, &m->d_celem_h, (size_t)total + 1));
		HB_TRY(hb_dalloc_t(m, &m->d_he_celem, (size_t)m->ne + 1));
		HB_CUDA(ctx, cudaMemsetAsync(m->d_he_celem, 0xff, sizeof(uint32_t) * ((size_t)m->ne + 1), ctx->stream));
		HB_LAUNCH(ctx, k_corner_elems, hb_div_up(m->norder_f, 256), 256, 0, m->d_ford_h, m->d_cbase, m->d_he, m->norder_f, m->d_celem_h, m->d_he_celem);
		HB_TRY(hb_dalloc_t(m, &m->d_reg_ncorner, m->reg_ncorner.size() + 1));
		HB_CUDA(ctx, cudaMemcpyAsync(m->d_reg_ncorner, m->reg_ncorner.data(), sizeof(int) * m->reg_ncorner.size(), cudaMemcpyHostToDevice, ctx->stream));
	}
	m->conn_ready = true;
	return 0;
}

int hb_build_vertex_candidates(hb_dmesh *m)
{
	hb_ctx *ctx = m->ctx;
	if (m->vcand_ready) return 0;
	HB_TRY(hb_build_conn(m));
	const uint32_t n = m->norder;
	HB_TRY(hb_dalloc_t(m, &m->d_vc_off, (size_t)n + 2));
	m->vc_total = 0;
	if (n) {
		HB_LAUNCH(ctx, k_vertex_candidates<false>, hb_div_up(n, 256), 256, 0, m->d_he, m->d_ord_h, m->d_ord_v, m->d_vrank, m->d_vtx_regs, n, m->ne, m->d_vc_off, (uint32_t *)nullptr, ctx->d_err);
		HB_TRY(hb_scan_exclusive_u32(ctx, m->d_vc_off, m->d_vc_off, n, nullptr));
		HB_CUDA(ctx, cudaMemcpyAsync(&m->vc_total, m->d_vc_off + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
		HB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
		HB_TRY(hb_check_device_error(ctx, "vertex fan walk"));
		HB_TRY(hb_dalloc_t(m, &m->d_vc_tri, 3 * (size_t)m->vc_total + 3));
		HB_LAUNCH(ctx, k_vertex_candidates<true>, hb_div_up(n, 256), 256, 0, m->d_he, m->d_ord_h, m->d_ord_v, m->d_vrank, m->d_vtx_regs, n, m->ne, m->d_vc_off, m->d_vc_tri, ctx->d_err);
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
