// hb_encode.cu -- encode side of the attribute coder (AttrCoder<WR>::encode,
// formats/hry/attrcode.h:321-416) as data-parallel kernels:
//   elem_rows      element -> bound attribute row (or none), per list
//   first_ref      first reference of every attribute row (atomicMin) = its DATA emission
//                  (GlobalHistory, attrcode.h:23-53)
//   lhist          per-(corner slot, vertex) local history (LocalHistory, attrcode.h:54-80)
//   gather_rp      rows -> rank-space value records (SoA-by-traversal-order transposition)
//   K5 encode_main prediction from the fan candidates, residual (pred::encodeDelta), byte-plane
//                  symbol rows, per-context 256-bin histograms, type / history-offset streams
#include "hb_lists.cuh"
#include <type_traits>
#include "hb_decode_spec.cuh" // SpecRec: packed rank-space records

// ------------------------------------------------------------------------------------------------
// The binding tables hold row indices local to the segment (mesh of a batch) of the element; the segment is found
// from the entity (vertex / face / half-edge) by binary search over `ent_base`.
struct RowSeg {
	uint32_t nseg;
	const uint32_t *ent_base;         // vbase / fbase / ebase of the class
	const uint32_t *rowbase, *rownum; // of the list
};
template <int CLS>
__global__ void __launch_bounds__(256) k_elem_rows(ElemCtx c, int l, RowSeg rs, uint32_t *__restrict__ erow, uint32_t *__restrict__ bound_flag, int *err)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= c.n) return;
	uint32_t row, entity;
	int a;
	bool ok = elem_lookup<CLS>(c, i, l, row, a, entity);
	if (ok) {
		const uint32_t s = hb_seg_find(rs.ent_base, rs.nseg, entity);
		if (row >= rs.rownum[s]) { atomicExch(err, 7); ok = false; }
		row += rs.rowbase[s];
	}
	erow[i] = ok ? row : HB_NONE;
	if (bound_flag) bound_flag[i] = ok ? 1u : 0u;
}

// *dup (optional) is raised when some row is referenced twice: without shared rows every emission is a DATA row, its
// ordinal is its emission index, and the consumers skip the first-reference gather and the ordinal scan
__global__ void __launch_bounds__(256) k_first_ref(const uint32_t *__restrict__ erow, uint32_t n, uint32_t *__restrict__ first, uint32_t *dup)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint32_t row = erow[i];
	if (row == HB_NONE) return;
	const uint32_t old = atomicMin(&first[row], i);
	if (dup && old != HB_NONE) *dup = 1u;
}

// corner lists: an element answered by the local history never reaches the global history
// (attrcode.h:377-381), so it must not count as a reference
__global__ void __launch_bounds__(256) k_first_ref_corner(const uint32_t *__restrict__ erow, uint32_t n, const uint4 *__restrict__ he, const uint32_t *__restrict__ celem_h,
                                                          const uint16_t *__restrict__ face_regs, const int16_t *__restrict__ slot_corner, uint32_t nlists, int l,
                                                          const uint32_t *__restrict__ lh, uint32_t ncel, uint32_t *__restrict__ first)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint32_t row = erow[i];
	if (row == HB_NONE) return;
	const int slot = slot_corner[(uint32_t)face_regs[he[celem_h[i]].w] * nlists + (uint32_t)l];
	if (lh[(size_t)slot * ncel + i] != HB_NONE) return;
	atomicMin(&first[row], i);
}

// decode with drained type symbols: the DATA emission owns its row
__global__ void __launch_bounds__(256) k_owner_from_types(const uint32_t *__restrict__ erow, const uint32_t *__restrict__ ek, const uint8_t *__restrict__ types, uint32_t ntypes,
                                                          uint32_t n, uint32_t *__restrict__ first, int *err)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint32_t row = erow[i];
	if (row == HB_NONE) return;
	const uint32_t k = ek ? ek[i] : i;
	if (k >= ntypes) { atomicExch(err, 8); return; }
	if (types[k] == HB_DATA) first[row] = i;
}

__global__ void __launch_bounds__(256) k_data_flags(const uint32_t *__restrict__ erow, const uint32_t *__restrict__ first, uint32_t n, uint32_t *__restrict__ dflag,
                                                    const uint32_t *__restrict__ skip_if_zero)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	if (skip_if_zero && *skip_if_zero == 0u) return;
	const uint32_t row = erow[i];
	dflag[i] = (row != HB_NONE && first[row] == i) ? 1u : 0u;
}

// rows (AoS, by attribute index) -> rank-space records: u64 container per component, element-major
__global__ void __launch_bounds__(256) k_gather_rp(ListParams p, const uint32_t *__restrict__ erow, uint32_t n, unsigned long long *__restrict__ rp)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint32_t row = erow[i];
	if (row == HB_NONE) return;
	const uint8_t *src = p.rows + (size_t)row * p.stride;
	for (int j = 0; j < p.ncomp; ++j) rp[(size_t)i * p.ncomp + j] = hb_ld_bits(src + p.offset[j], p.size[j]);
}

// ------------------------------------------------------------------------------------------------
// local history of corner bindings
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_vinc_count(const uint4 *__restrict__ he, const uint32_t *__restrict__ celem_h, const uint16_t *__restrict__ face_regs,
                                                     const int *__restrict__ reg_ncorner, uint32_t n, uint32_t *__restrict__ cnt)
{
	const uint32_t ce = blockIdx.x * blockDim.x + threadIdx.x;
	if (ce >= n) return;
	const uint4 r = he[celem_h[ce]];
	if (reg_ncorner[face_regs[r.w]] > 0) atomicAdd(&cnt[r.x], 1u);
}
__global__ void __launch_bounds__(256) k_vinc_fill(const uint4 *__restrict__ he, const uint32_t *__restrict__ celem_h, const uint16_t *__restrict__ face_regs,
                                                    const int *__restrict__ reg_ncorner, uint32_t n, const uint32_t *__restrict__ off, uint32_t *__restrict__ cursor, uint32_t *__restrict__ vinc)
{
	const uint32_t ce = blockIdx.x * blockDim.x + threadIdx.x;
	if (ce >= n) return;
	const uint4 r = he[celem_h[ce]];
	if (reg_ncorner[face_regs[r.w]] > 0) vinc[off[r.x] + atomicAdd(&cursor[r.x], 1u)] = ce;
}

__device__ void sort_u32(uint32_t *a, uint32_t n)
{
	if (n < 32) { // insertion sort
		for (uint32_t i = 1; i < n; ++i) {
			const uint32_t x = a[i];
			uint32_t j = i;
			while (j > 0 && a[j - 1] > x) { a[j] = a[j - 1]; --j; }
			a[j] = x;
		}
		return;
	}
	// heap sort for the rare huge fan (sphere poles)
	for (uint32_t start = n / 2; start-- > 0;) {
		uint32_t root = start;
		for (;;) {
			uint32_t child = 2 * root + 1;
			if (child >= n) break;
			if (child + 1 < n && a[child] < a[child + 1]) ++child;
			if (a[root] >= a[child]) break;
			const uint32_t t = a[root]; a[root] = a[child]; a[child] = t;
			root = child;
		}
	}
	for (uint32_t end = n; end-- > 1;) {
		const uint32_t t = a[0]; a[0] = a[end]; a[end] = t;
		uint32_t root = 0;
		for (;;) {
			uint32_t child = 2 * root + 1;
			if (child >= end) break;
			if (child + 1 < end && a[child] < a[child + 1]) ++child;
			if (a[root] >= a[child]) break;
			const uint32_t u = a[root]; a[root] = a[child]; a[child] = u;
			root = child;
		}
	}
}

// One thread per vertex: its corner elements in emission order, and per binding slot the list of
// distinct attribute indices seen so far.  A hit yields the offset size-1-pos (attrcode.h:67-74).
__global__ void __launch_bounds__(128) k_lhist(const uint4 *__restrict__ he, const uint32_t *__restrict__ celem_h, const uint16_t *__restrict__ face_regs,
                                                const int *__restrict__ reg_ncorner, const uint32_t *__restrict__ bind_corner, uint32_t nb_corner, uint32_t nv,
                                                const uint32_t *__restrict__ off, uint32_t *__restrict__ vinc, uint32_t *__restrict__ dist, uint32_t ncel, uint32_t *__restrict__ lh)
{
	const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
	if (v >= nv) return;
	const uint32_t b = off[v], cnt = off[v + 1] - b;
	if (cnt == 0) return;
	sort_u32(vinc + b, cnt);
	for (uint32_t a = 0; a < nb_corner; ++a) {
		uint32_t nd = 0;
		for (uint32_t k = 0; k < cnt; ++k) {
			const uint32_t ce = vinc[b + k];
			const uint32_t h = celem_h[ce];
			if ((uint32_t)reg_ncorner[face_regs[he[h].w]] <= a) continue;
			const uint32_t idx = bind_corner[(size_t)h * nb_corner + a];
			uint32_t hit = HB_NONE;
			for (uint32_t q = 0; q < nd; ++q)
				if (dist[b + q] == idx) { hit = nd - 1 - q; break; }
			lh[(size_t)a * ncel + ce] = hit;
			if (hit == HB_NONE) dist[b + nd++] = idx;
		}
	}
}

// ------------------------------------------------------------------------------------------------
// Lists without components (the face list of a PLY without face properties: 20M elements at configs[1]).  The
// reference still emits one type symbol per element (io.h:90-94): DATA for the first reference of a row, HIST after.
// Two passes instead of the six of the general path (element rows, first reference, DATA flags, scan, K5): pass 1
// finds the first reference of every row, pass 2 writes the type symbols and counts them per segment.  History
// offsets need the scan of the DATA flags: only when pass 2 counted a HIST emission (checked when the streams are
// fetched, hb_nocomp_finish) -- for identity bindings it never runs.
// ------------------------------------------------------------------------------------------------
struct NoCompArgs {
	const uint32_t *order_f;   // FACE: traversal order (fepair records) or nullptr = index order
	const uint32_t *ord_v;     // VTX: vertex of every traversal position
	const uint16_t *regs;
	const int16_t *slot;       // [region * nlists + list]
	const uint32_t *bind;
	uint32_t nb, nlists, n;
	int l;
	RowSeg rs;                 // ent_base = fbase / vbase
	const uint32_t *elem_base; // ofbase / obase
	uint32_t *erow, *first;
	uint8_t *type;
	unsigned long long *type_hist; // per segment, hist_pitch words apart
	size_t hist_pitch;
};
template <int CLS>
__device__ __forceinline__ uint32_t nocomp_row(const NoCompArgs &a, uint32_t i, int *err)
{
	uint32_t ent, s;
	if (CLS == CLS_FACE) {
		s = hb_seg_find(a.elem_base, a.rs.nseg, i);
		const uint32_t fb = a.rs.ent_base[s];
		const uint32_t fl = a.order_f ? a.order_f[2 * (size_t)i] : i - a.elem_base[s];
		if (fl >= a.rs.ent_base[s + 1] - fb) { atomicExch(err, 4); return HB_NONE; }
		ent = fl + fb;
	} else {
		ent = a.ord_v[i];
		s = hb_seg_find(a.rs.ent_base, a.rs.nseg, ent);
	}
	const int sl = a.slot[(uint32_t)a.regs[ent] * a.nlists + (uint32_t)a.l];
	if (sl < 0) return HB_NONE;
	const uint32_t row = a.bind[(size_t)ent * a.nb + (uint32_t)sl];
	if (row >= a.rs.rownum[s]) { atomicExch(err, 7); return HB_NONE; }
	return row + a.rs.rowbase[s];
}
// pass 1: first reference of every row; *dup is raised when some row is referenced twice (never for identity bindings)
template <int CLS>
__global__ void __launch_bounds__(256) k_nocomp_first(NoCompArgs a, uint32_t *dup, int *err)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= a.n) return;
	const uint32_t row = nocomp_row<CLS>(a, i, err);
	if (row != HB_NONE && atomicMin(&a.first[row], i) != HB_NONE) *dup = 1u;
}
// pass 2.  No row is shared (the usual case): every emission is a DATA row, the type stream stays the zeros it was
// cleared to and only the per-segment counts are written -- nothing is gathered.  Otherwise: type symbols per element.
template <int CLS>
__global__ void __launch_bounds__(256) k_nocomp_types(NoCompArgs a, const uint32_t *dup, int *err)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (*dup == 0u) {
		for (uint32_t sg = i; sg < a.rs.nseg; sg += gridDim.x * blockDim.x)
			a.type_hist[(size_t)sg * a.hist_pitch + HB_DATA] = a.elem_base[sg + 1] - a.elem_base[sg]; // bound in every region
		return;
	}
	int t = -1;
	if (i < a.n) {
		const uint32_t row = nocomp_row<CLS>(a, i, err);
		a.erow[i] = row;
		if (row != HB_NONE) {
			t = a.first[row] == i ? HB_DATA : HB_HIST;
			a.type[i] = (uint8_t)t;
		}
	}
	// per-segment counts: one atomic per warp, type and segment (a warp rarely straddles a boundary)
	const uint32_t seg = a.rs.nseg > 1 && i < a.n ? hb_seg_find(a.elem_base, a.rs.nseg, i) : 0u;
	const uint32_t seg0 = __shfl_sync(0xffffffffu, seg, 0);
	const bool uniform = __all_sync(0xffffffffu, seg == seg0 || t < 0);
	for (int ty = 0; ty < 2; ++ty) {
		if (uniform) {
			const unsigned mk = __ballot_sync(0xffffffffu, t == ty);
			if (mk && (threadIdx.x & 31u) == 0) atomicAdd(&a.type_hist[(size_t)seg0 * a.hist_pitch + ty], (unsigned long long)__popc(mk));
		} else if (t == ty) {
			atomicAdd(&a.type_hist[(size_t)seg * a.hist_pitch + ty], 1ull);
		}
	}
}
// history offsets of the HIST emissions (attrcode.h:43-52): tidx - 1 - g
__global__ void __launch_bounds__(256) k_nocomp_aux(const uint32_t *__restrict__ erow, const uint32_t *__restrict__ first, const uint32_t *__restrict__ dord, uint32_t n, uint32_t *__restrict__ aux)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint32_t row = erow[i];
	uint32_t v = 0;
	if (row != HB_NONE && first[row] != i) v = dord[i] - 1u - dord[first[row]];
	aux[i] = v;
}

// ------------------------------------------------------------------------------------------------
// K5: prediction + residual + byte-plane symbols + histograms
// ------------------------------------------------------------------------------------------------
#define ENC_THREADS 256
#define HIST_SMEM_CTX 24 // contexts histogrammed in shared memory; the rest goes to global atomics
#define ENC_WIDE_K 48     // vertices with more candidates (poles, huge fans) are coded by a whole warp

struct EncodeArgs {
	const uint32_t *erow, *ek, *first, *dord; // ek == nullptr: emission index == element index
	const unsigned long long *rp;             // rank-space values (VTX / CORNER)
	const uint32_t *cand_off, *cand;          // VTX: rank triples, CORNER: corner elements
	const uint32_t *lh;                       // CORNER: local-history offsets of this list's slot table base
	uint32_t ncel;
	const uint16_t *face_regs;
	const int16_t *slot_corner;
	uint32_t nlists;
	const uint32_t *celem_h;
	const uint4 *he;
	uint8_t *type;
	uint32_t *aux;
	uint8_t *sym;
	unsigned long long *hist, *type_hist;
	uint32_t n;
	int l;
	uint32_t *wide;       // [0] = count, [1..] = element indices whose candidate count exceeds ENC_WIDE_K
	uint32_t wide_cap;
	const uint32_t *dup;       // device flag or nullptr: 0 = no row is shared (every emission a DATA row, ordinal = emission index)
	// histograms are kept per segment (mesh of a batch): hist + s * hist_pitch, type counters behind the contexts
	uint32_t nseg;
	const uint32_t *elem_base; // first element of every segment (nseg + 1)
	size_t hist_pitch;         // in u64 words
};
// segment of the elements of this block: the common one, or HB_NONE when the block straddles a boundary
// (then every thread adds to the histograms of its own segment with global atomics)
__device__ __forceinline__ uint32_t block_segment(const EncodeArgs &a, uint32_t first, uint32_t &mine, uint32_t i)
{
	mine = 0;
	if (a.nseg <= 1) return 0;
	const uint32_t last = min(first + ENC_THREADS, a.n) - 1;
	const uint32_t s0 = hb_seg_find(a.elem_base, a.nseg, first), s1 = hb_seg_find(a.elem_base, a.nseg, last);
	if (s0 == s1) { mine = s0; return s0; }
	mine = hb_seg_find(a.elem_base, a.nseg, i < a.n ? i : last);
	return HB_NONE;
}

template <int CLS>
__global__ void __launch_bounds__(ENC_THREADS) k_encode_main(ListParams p, EncodeArgs a)
{
	extern __shared__ uint32_t s_hist[]; // [min(sym_stride, HIST_SMEM_CTX)][256] + 4 type counters
	const int nctx_s = (int)(p.sym_stride < HIST_SMEM_CTX ? p.sym_stride : HIST_SMEM_CTX);
	for (int k = threadIdx.x; k < nctx_s * 256 + 4; k += ENC_THREADS) s_hist[k] = 0;
	__syncthreads();
	uint32_t *s_type = s_hist + nctx_s * 256;

	const uint32_t i = blockIdx.x * ENC_THREADS + threadIdx.x;
	uint32_t myseg;
	const uint32_t bseg = block_segment(a, blockIdx.x * ENC_THREADS, myseg, i);
	const bool agg = bseg != HB_NONE; // one segment in this block: shared-memory histograms
	unsigned long long *ghist = a.hist + (size_t)myseg * a.hist_pitch, *gtype = a.type_hist + (size_t)myseg * a.hist_pitch;
	const bool fast = a.dup && *a.dup == 0u; // no row is shared: every emission is a DATA row, its ordinal the emission index
	const uint32_t row = i < a.n ? a.erow[i] : HB_NONE;
	int t = -1;                 // emission type of this element, -1 = no emission
	unsigned long long res[HB_MAX_COMP > 8 ? 8 : HB_MAX_COMP]; // residuals of the first 8 components (register resident)
	if (row != HB_NONE) {
		const uint32_t k = a.ek ? a.ek[i] : i;
		const uint32_t fi = fast ? i : a.first[row];
		t = HB_DATA;
		uint32_t aux = 0;
		bool lhit = false;
		if (CLS == CLS_CORNER) {
			// the local history is consulted first (attrcode.h:377-381); it is keyed by binding
			// slot, so a hit may even name a row this list has not emitted (first[row] unset)
			const uint32_t h = a.celem_h[i];
			const int slot = a.slot_corner[(uint32_t)a.face_regs[a.he[h].w] * a.nlists + (uint32_t)a.l];
			const uint32_t lo = a.lh[(size_t)slot * a.ncel + i];
			if (lo != HB_NONE) { t = HB_LHIST; aux = lo; lhit = true; }
		}
		if (!lhit && fi != i) {
			t = HB_HIST;
			aux = a.dord[i] - 1u - a.dord[fi]; // tidx - 1 - g, attrcode.h:43-52
		}
		if (!fast) { // (the streams were cleared: all DATA, no offsets)
			a.type[k] = (uint8_t)t;
			a.aux[k] = aux;
		}
		if (t == HB_DATA && CLS == CLS_VTX && a.wide && a.cand_off[i + 1] - a.cand_off[i] > ENC_WIDE_K) {
			const uint32_t slot = atomicAdd(&a.wide[0], 1u);
			if (slot < a.wide_cap) { a.wide[1 + slot] = i; t = -2; } // residual + histogram by k_encode_wide
		}
		if (t == HB_DATA) {
			const uint32_t d = fast ? k : a.dord[i];
			uint8_t *out = a.sym + (size_t)d * p.sym_stride;
			uint32_t c0 = 0, K = 0;
			if (CLS != CLS_FACE) { c0 = a.cand_off[i]; K = a.cand_off[i + 1] - c0; }
			for (int j = 0; j < p.ncomp; ++j) {
				const int st = p.stype[j], q = p.quant[j];
				unsigned long long raw, pred = 0;
				if (CLS == CLS_FACE) {
					// face prediction is always 0 (Appendix C.2): no rank-space copy needed
					raw = hb_ld_bits(p.rows + (size_t)row * p.stride + p.offset[j], p.size[j]);
				} else {
					raw = a.rp[(size_t)i * p.ncomp + j];
					if (CLS == CLS_VTX) {
						pred = combine_candidates(st, K, [&](uint32_t kk) {
							const uint32_t *tr = a.cand + 3 * (size_t)(c0 + kk);
							return hb_predict(st, a.rp[(size_t)tr[0] * p.ncomp + j], a.rp[(size_t)tr[1] * p.ncomp + j], a.rp[(size_t)tr[2] * p.ncomp + j], q);
						});
					} else {
						pred = combine_candidates(st, K, [&](uint32_t kk) { return a.rp[(size_t)a.cand[c0 + kk] * p.ncomp + j]; });
					}
				}
				const unsigned long long r = hb_enc(st, raw, pred, q);
				hb_st_bits(out + p.sym_off[j], p.size[j], r);
				if (j < 8) {
					res[j] = r;
				} else { // rare wide rows: no aggregation
					for (int b = 0; b < p.size[j]; ++b) {
						const uint32_t ctx = p.sym_off[j] + b, s = (uint32_t)(r >> (8 * b)) & 0xffu;
						if (agg && (int)ctx < nctx_s) atomicAdd(&s_hist[ctx * 256 + s], 1u);
						else atomicAdd(&ghist[(size_t)ctx * 256 + s], 1ull);
					}
				}
			}
		}
	}
	// Histogram update, warp-aggregated: residuals of a smooth mesh are tiny, so most lanes hit
	// the same bin (the high bytes are almost always 0).  Lanes with equal symbols are grouped with
	// __match_any_sync and one lane adds the group size.
	{
		const bool is_data = t == HB_DATA;
		const int ncj = p.ncomp < 8 ? p.ncomp : 8;
		for (int j = 0; j < ncj; ++j) {
			for (int b = 0; b < p.size[j]; ++b) {
				const uint32_t ctx = p.sym_off[j] + b;
				if (!agg) { // block across a segment boundary (rare): plain global atomics on the thread's own segment
					if (is_data) atomicAdd(&ghist[(size_t)ctx * 256 + ((uint32_t)(res[j] >> (8 * b)) & 0xffu)], 1ull);
					continue;
				}
				const uint32_t s = is_data ? (uint32_t)(res[j] >> (8 * b)) & 0xffu : 0x100u + (threadIdx.x & 31u);
				const unsigned grp = __match_any_sync(0xffffffffu, s);
				if (is_data && (int)(__ffs(grp) - 1) == (int)(threadIdx.x & 31u)) {
					const uint32_t cnt = __popc(grp);
					if ((int)ctx < nctx_s) atomicAdd(&s_hist[ctx * 256 + s], cnt);
					else atomicAdd(&ghist[(size_t)ctx * 256 + s], (unsigned long long)cnt);
				}
			}
		}
		// emission-type counts: one shared-memory atomic per warp and type
		if (agg) {
			for (int ty = 0; ty < 3; ++ty) {
				const unsigned m = __ballot_sync(0xffffffffu, t == ty || (ty == HB_DATA && t == -2));
				if (m && (threadIdx.x & 31u) == 0) atomicAdd(&s_type[ty], (uint32_t)__popc(m));
			}
		} else if (t >= 0 || t == -2) {
			atomicAdd(&gtype[t == -2 ? HB_DATA : t], 1ull);
		}
	}
	__syncthreads();
	if (agg) {
		unsigned long long *bh = a.hist + (size_t)bseg * a.hist_pitch, *bt = a.type_hist + (size_t)bseg * a.hist_pitch;
		for (int k = threadIdx.x; k < nctx_s * 256; k += ENC_THREADS)
			if (s_hist[k]) atomicAdd(&bh[k], (unsigned long long)s_hist[k]);
		if (threadIdx.x < 3 && s_type[threadIdx.x]) atomicAdd(&bt[threadIdx.x], (unsigned long long)s_type[threadIdx.x]);
	}
}

// One warp per wide vertex element: the candidate predictions are summed with a warp-strided loop
// and a shuffle reduction.  Integer storage types only add in int64 (order independent, so this is
// exact); a float component keeps the reference's in-order double accumulation on lane 0.
__global__ void __launch_bounds__(128) k_encode_wide(ListParams p, EncodeArgs a)
{
	const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	const uint32_t nw = a.wide[0] < a.wide_cap ? a.wide[0] : a.wide_cap;
	if (w >= nw) return;
	const uint32_t i = a.wide[1 + w];
	const uint32_t c0 = a.cand_off[i], K = a.cand_off[i + 1] - c0;
	const bool fast = a.dup && *a.dup == 0u;
	uint8_t *out = a.sym + (size_t)(fast ? (a.ek ? a.ek[i] : i) : a.dord[i]) * p.sym_stride;
	unsigned long long *ghist = a.hist + (size_t)hb_seg_find(a.elem_base, a.nseg, i) * a.hist_pitch;
	for (int j = 0; j < p.ncomp; ++j) {
		const int st = p.stype[j], q = p.quant[j];
		unsigned long long pred;
		if (st == HB_FLOAT) {
			pred = 0;
			if (lane == 0)
				pred = combine_candidates(st, K, [&](uint32_t kk) {
					const uint32_t *tr = a.cand + 3 * (size_t)(c0 + kk);
					return hb_predict(st, a.rp[(size_t)tr[0] * p.ncomp + j], a.rp[(size_t)tr[1] * p.ncomp + j], a.rp[(size_t)tr[2] * p.ncomp + j], q);
				});
		} else {
			unsigned long long sum = 0;
			for (uint32_t kk = lane; kk < K; kk += 32) {
				const uint32_t *tr = a.cand + 3 * (size_t)(c0 + kk);
				const unsigned long long v = hb_predict(st, a.rp[(size_t)tr[0] * p.ncomp + j], a.rp[(size_t)tr[1] * p.ncomp + j], a.rp[(size_t)tr[2] * p.ncomp + j], q);
				sum += st == HB_ULONG ? v : (unsigned long long)hb_bits_to_i64(v, st);
			}
#pragma unroll
			for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
			if (st == HB_ULONG) {
				pred = (sum + ((unsigned long long)K >> 1)) / (unsigned long long)K;
			} else {
				pred = (unsigned long long)hb_divround_i64((long long)sum, (int)K);
				const int size = hb_type_size(st);
				if (size < 8) pred &= (1ull << (8 * size)) - 1ull;
			}
		}
		if (lane == 0) {
			const unsigned long long r = hb_enc(st, a.rp[(size_t)i * p.ncomp + j], pred, q);
			hb_st_bits(out + p.sym_off[j], p.size[j], r);
			for (int b = 0; b < p.size[j]; ++b)
				atomicAdd(&ghist[(size_t)(p.sym_off[j] + b) * 256 + ((uint32_t)(r >> (8 * b)) & 0xffu)], 1ull);
		}
	}
}

// ------------------------------------------------------------------------------------------------
// K5, packed fast path: vertex lists whose components share one integer storage type (every quantized
// list).  Values live in rank space as one aligned record per vertex (8 bytes for three 16-bit
// components), so a parallelogram operand is ONE load instead of one 8-byte container per component,
// and the arithmetic runs in T instead of runtime-typed u64 containers.  Same results as
// k_encode_main<CLS_VTX>.
// ------------------------------------------------------------------------------------------------
template <typename T, int NC>
__global__ void __launch_bounds__(256) k_gather_packed(ListParams p, const uint32_t *__restrict__ erow, uint32_t n, SpecRec<T, NC> *__restrict__ out)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint32_t row = erow[i];
	SpecRec<T, NC> r;
#pragma unroll
	for (int j = 0; j < (int)(sizeof(r.c) / sizeof(T)); ++j) r.c[j] = 0;
	if (row != HB_NONE) {
		const uint8_t *src = p.rows + (size_t)row * p.stride;
#pragma unroll
		for (int j = 0; j < NC; ++j) r.c[j] = (T)hb_ld_bits(src + p.offset[j], (int)sizeof(T));
	}
	out[i] = r;
}

template <typename T, int NC>
__global__ void __launch_bounds__(ENC_THREADS) k_encode_vtx_packed(ListParams p, EncodeArgs a, const SpecRec<T, NC> *__restrict__ rec, uint32_t chunks_per_cta)
{
	typedef SpecRec<T, NC> Rec;
	typedef typename std::conditional<(sizeof(T) <= 2), uint32_t, unsigned long long>::type Acc;
	extern __shared__ uint32_t s_hist[]; // [NC * sizeof(T)][256] + 4 type counters
	constexpr int NCTX = NC * (int)sizeof(T);
	for (int k = threadIdx.x; k < NCTX * 256 + 4; k += ENC_THREADS) s_hist[k] = 0;
	__syncthreads();
	uint32_t *s_type = s_hist + NCTX * 256;
	int bits[NC];
#pragma unroll
	for (int j = 0; j < NC; ++j) bits[j] = hb_stype_bits(p.stype[j], p.quant[j]);
	// A CTA codes a contiguous run of chunks of ENC_THREADS elements and keeps ONE set of shared-memory histograms for
	// the segment it is in: they are cleared and flushed when the segment changes and at the end, not once per chunk
	// (zeroing + flushing 1536 bins cost a sixth of the instructions of the one-chunk-per-CTA version).
	const bool fast = a.dup && *a.dup == 0u; // no row is shared: no element-row / first-reference / ordinal loads, no type / offset stores
	uint32_t cur_seg = HB_NONE;
	auto flush = [&]() {
		if (cur_seg == HB_NONE) return;
		__syncthreads();
		unsigned long long *bh = a.hist + (size_t)cur_seg * a.hist_pitch, *bt = a.type_hist + (size_t)cur_seg * a.hist_pitch;
		for (int k = threadIdx.x; k < NCTX * 256; k += ENC_THREADS) {
			const uint32_t v = s_hist[k];
			if (v) { atomicAdd(&bh[k], (unsigned long long)v); s_hist[k] = 0; }
		}
		if (threadIdx.x < 3 && s_type[threadIdx.x]) { atomicAdd(&bt[threadIdx.x], (unsigned long long)s_type[threadIdx.x]); s_type[threadIdx.x] = 0; }
		__syncthreads();
	};
	const uint32_t nchunks = (a.n + ENC_THREADS - 1) / ENC_THREADS;
	const uint32_t c_begin = blockIdx.x * chunks_per_cta, c_end = min(nchunks, c_begin + chunks_per_cta);
	for (uint32_t chunk = c_begin; chunk < c_end; ++chunk) {
		const uint32_t i = chunk * ENC_THREADS + threadIdx.x;
		uint32_t myseg;
		const uint32_t bseg = block_segment(a, chunk * ENC_THREADS, myseg, i);
		const bool agg = bseg != HB_NONE; // one segment in this chunk: shared-memory histograms
		if (agg && bseg != cur_seg) { flush(); cur_seg = bseg; }
		const uint32_t row = i < a.n ? (fast ? 0u : a.erow[i]) : HB_NONE; // (the shortcut is only taken by lists every region binds)
		int t = -1;
		T res[NC];
#pragma unroll
		for (int j = 0; j < NC; ++j) res[j] = 0;
		if (row != HB_NONE) {
			const uint32_t k = a.ek ? a.ek[i] : i;
			t = HB_DATA;
			if (!fast) {
				const uint32_t fi = a.first[row];
				uint32_t aux = 0;
				if (fi != i) {
					t = HB_HIST;
					aux = a.dord[i] - 1u - a.dord[fi]; // tidx - 1 - g, attrcode.h:43-52
				}
				a.type[k] = (uint8_t)t;
				a.aux[k] = aux;
			}
			const uint32_t c0 = a.cand_off[i], K = a.cand_off[i + 1] - c0;
			if (t == HB_DATA && a.wide && K > ENC_WIDE_K) {
				const uint32_t slot = atomicAdd(&a.wide[0], 1u);
				if (slot < a.wide_cap) { a.wide[1 + slot] = i; t = -2; } // residual + histogram by k_encode_wide_packed
			}
			if (t == HB_DATA) {
				const Rec raw = rec[i];
				Acc sum[NC];
#pragma unroll
				for (int j = 0; j < NC; ++j) sum[j] = 0;
				for (uint32_t kk = 0; kk < K; ++kk) {
					const uint32_t *tr = a.cand + 3 * (size_t)(c0 + kk);
					const Rec v0 = rec[tr[0]], v1 = rec[tr[1]], v2 = rec[tr[2]];
#pragma unroll
					for (int j = 0; j < NC; ++j) sum[j] += (Acc)IntOps<T>::predict(v0.c[j], v1.c[j], v2.c[j], bits[j]);
				}
				uint8_t *out = a.sym + (size_t)(fast ? k : a.dord[i]) * p.sym_stride;
#pragma unroll
				for (int j = 0; j < NC; ++j) {
					const T pred = K == 0 ? (T)0 : (K == 1 ? (T)sum[j] : (K == 2 ? (T)((sum[j] + 1) >> 1) : (T)hb_divround_i64((long long)sum[j], (int)K)));
					res[j] = IntOps<T>::enc(raw.c[j], pred, bits[j]);
					hb_st_bits(out + p.sym_off[j], (int)sizeof(T), res[j]);
				}
			}
		}
		// warp-aggregated histogram update (see k_encode_main)
		const bool is_data = t == HB_DATA;
#pragma unroll
		for (int j = 0; j < NC; ++j) {
#pragma unroll
			for (int b = 0; b < (int)sizeof(T); ++b) {
				const uint32_t ctx = p.sym_off[j] + b;
				if (!agg) { // chunk across a segment boundary (rare): plain global atomics on the thread's own segment
					if (is_data) atomicAdd(&a.hist[(size_t)myseg * a.hist_pitch + (size_t)ctx * 256 + (((uint32_t)res[j] >> (8 * b)) & 0xffu)], 1ull);
					continue;
				}
				const uint32_t s = is_data ? ((uint32_t)res[j] >> (8 * b)) & 0xffu : 0x100u + (threadIdx.x & 31u);
				const unsigned grp = __match_any_sync(0xffffffffu, s);
				if (is_data && (int)(__ffs(grp) - 1) == (int)(threadIdx.x & 31u)) atomicAdd(&s_hist[ctx * 256 + s], (uint32_t)__popc(grp));
			}
		}
		if (agg) {
			for (int ty = 0; ty < 3; ++ty) {
				const unsigned m = __ballot_sync(0xffffffffu, t == ty || (ty == HB_DATA && t == -2));
				if (m && (threadIdx.x & 31u) == 0) atomicAdd(&s_type[ty], (uint32_t)__popc(m));
			}
		} else if (t >= 0 || t == -2) {
			atomicAdd(&a.type_hist[(size_t)myseg * a.hist_pitch + (t == -2 ? HB_DATA : t)], 1ull);
		}
	}
	flush();
}

// one CTA per wide vertex (integer sums are order independent)
#define ENC_WIDE_T 256
template <typename T, int NC>
__global__ void __launch_bounds__(ENC_WIDE_T) k_encode_wide_packed(ListParams p, EncodeArgs a, const SpecRec<T, NC> *__restrict__ rec)
{
	typedef SpecRec<T, NC> Rec;
	const uint32_t nw = a.wide[0] < a.wide_cap ? a.wide[0] : a.wide_cap;
	__shared__ unsigned long long s_sum[NC];
	for (uint32_t w = blockIdx.x; w < nw; w += gridDim.x) {
		if (threadIdx.x < NC) s_sum[threadIdx.x] = 0ull;
		__syncthreads();
		const uint32_t i = a.wide[1 + w];
		const uint32_t c0 = a.cand_off[i], K = a.cand_off[i + 1] - c0;
		unsigned long long sum[NC];
#pragma unroll
		for (int j = 0; j < NC; ++j) sum[j] = 0;
		for (uint32_t kk = threadIdx.x; kk < K; kk += ENC_WIDE_T) {
			const uint32_t *tr = a.cand + 3 * (size_t)(c0 + kk);
			const Rec v0 = rec[tr[0]], v1 = rec[tr[1]], v2 = rec[tr[2]];
#pragma unroll
			for (int j = 0; j < NC; ++j) sum[j] += (unsigned long long)IntOps<T>::predict(v0.c[j], v1.c[j], v2.c[j], hb_stype_bits(p.stype[j], p.quant[j]));
		}
#pragma unroll
		for (int j = 0; j < NC; ++j) {
#pragma unroll
			for (int d = 16; d > 0; d >>= 1) sum[j] += __shfl_xor_sync(0xffffffffu, sum[j], d);
			if ((threadIdx.x & 31) == 0) atomicAdd(&s_sum[j], sum[j]);
		}
		__syncthreads();
		if (threadIdx.x == 0) {
			const Rec raw = rec[i];
			const bool fast = a.dup && *a.dup == 0u;
			uint8_t *out = a.sym + (size_t)(fast ? (a.ek ? a.ek[i] : i) : a.dord[i]) * p.sym_stride;
			unsigned long long *ghist = a.hist + (size_t)hb_seg_find(a.elem_base, a.nseg, i) * a.hist_pitch;
#pragma unroll
			for (int j = 0; j < NC; ++j) {
				const T pred = (T)hb_divround_i64((long long)s_sum[j], (int)K);
				const T r = IntOps<T>::enc(raw.c[j], pred, hb_stype_bits(p.stype[j], p.quant[j]));
				hb_st_bits(out + p.sym_off[j], (int)sizeof(T), r);
				for (int b = 0; b < (int)sizeof(T); ++b)
					atomicAdd(&ghist[(size_t)(p.sym_off[j] + b) * 256 + (((uint32_t)r >> (8 * b)) & 0xffu)], 1ull);
			}
		}
		__syncthreads();
	}
}

template <typename T, int NC>
static int encode_vtx_packed_nc(hb_dmesh *m, DevList &dl, const EncodeArgs &a)
{
	hb_ctx *ctx = m->ctx;
	const uint32_t n = dl.n_elems;
	HB_TRY(hb_dalloc(m, (void **)&dl.d_cx, sizeof(SpecRec<T, NC>) * ((size_t)n + 1)));
	SpecRec<T, NC> *rec = (SpecRec<T, NC> *)dl.d_cx;
	HB_LAUNCH(ctx, (k_gather_packed<T, NC>), hb_div_up(n, 256), 256, 0, dl.p, dl.d_erow, n, rec);
	const size_t smem = sizeof(uint32_t) * ((size_t)NC * sizeof(T) * 256 + 4);
	// contiguous runs of chunks per CTA, about 16 CTAs per SM in all (8 resident at 32 registers: two waves)
	const uint32_t nchunks = hb_div_up(n, ENC_THREADS);
	const uint32_t per = hb_div_up(nchunks, (uint32_t)ctx->sm_count * 16u);
	HB_LAUNCH(ctx, (k_encode_vtx_packed<T, NC>), hb_div_up(nchunks, per ? per : 1u), ENC_THREADS, smem, dl.p, a, rec, per ? per : 1u);
	if (a.wide) HB_LAUNCH(ctx, (k_encode_wide_packed<T, NC>), 2 * ctx->sm_count, ENC_WIDE_T, 0, dl.p, a, rec);
	return 0;
}
template <typename T>
static int encode_vtx_packed(hb_dmesh *m, DevList &dl, const EncodeArgs &a)
{
	switch (dl.p.ncomp) {
	case 1: return encode_vtx_packed_nc<T, 1>(m, dl, a);
	case 2: return encode_vtx_packed_nc<T, 2>(m, dl, a);
	case 3: return encode_vtx_packed_nc<T, 3>(m, dl, a);
	default: return encode_vtx_packed_nc<T, 4>(m, dl, a);
	}
}
static bool packed_eligible(const ListParams &p)
{
	if (p.ncomp < 1 || p.ncomp > 4 || p.target != CLS_VTX) return false;
	const int st = p.uniform_stype;
	if (st != HB_UCHAR && st != HB_USHORT && st != HB_UINT) return false;
	for (int j = 0; j < p.ncomp; ++j)
		if (p.sym_off[j] != j * hb_type_size(st)) return false;
	return true;
}

// ------------------------------------------------------------------------------------------------
// host driver
// ------------------------------------------------------------------------------------------------
static ElemCtx make_elem_ctx(hb_dmesh *m, int cls)
{
	ElemCtx c;
	c.he = m->d_he;
	c.ord_v = m->d_ord_v;
	c.ford_h = m->d_ford_h;
	c.celem_h = m->d_celem_h;
	c.vtx_regs = m->d_vtx_regs;
	c.face_regs = m->d_face_regs;
	c.nlists = m->nlists;
	if (cls == CLS_VTX) { c.bind = m->d_bind_vtx; c.slot = m->d_slot_vtx; c.nb = m->nb_vtx; c.n = m->norder; }
	else if (cls == CLS_FACE) { c.bind = m->d_bind_face; c.slot = m->d_slot_face; c.nb = m->nb_face; c.n = m->norder_f; }
	else { c.bind = m->d_bind_corner; c.slot = m->d_slot_corner; c.nb = m->nb_corner; c.n = m->n_corner_elems; }
	return c;
}

// true when every region of the class binds list l (then emission index == element index)
static bool bound_everywhere(hb_dmesh *m, int cls, int l)
{
	const std::vector<int16_t> &slot = cls == CLS_VTX ? m->h_slot_vtx : cls == CLS_FACE ? m->h_slot_face : m->h_slot_corner;
	const int nregs = cls == CLS_VTX ? m->nregs_vtx : m->nregs_face;
	for (int r = 0; r < nregs; ++r)
		if (slot[(size_t)r * m->nlists + l] < 0) return false;
	return true;
}

int hb_prepare_list_elems(hb_dmesh *m, int l, bool need_rp, bool decode)
{
	hb_ctx *ctx = m->ctx;
	DevList &dl = m->lists[l];
	const int cls = dl.p.target;
	ElemCtx c = make_elem_ctx(m, cls);
	const uint32_t n = c.n;
	dl.n_elems = n;
	HB_TRY(hb_dalloc_t(m, &dl.d_erow, (size_t)n + 1));
	HB_TRY(hb_dalloc_t(m, &dl.d_first, (size_t)dl.p.nrows + 1));
	HB_TRY(hb_dalloc_t(m, &dl.d_dord, (size_t)n + 2));
	dl.d_ek = nullptr;
	const bool everywhere = bound_everywhere(m, cls, l);
	if (!everywhere) HB_TRY(hb_dalloc_t(m, &dl.d_ek, (size_t)n + 2));
	HB_CUDA(ctx, cudaMemsetAsync(dl.d_first, 0xff, sizeof(uint32_t) * ((size_t)dl.p.nrows + 1), ctx->stream));
	HB_CUDA(ctx, cudaMemsetAsync(dl.d_dord, 0, sizeof(uint32_t) * ((size_t)n + 2), ctx->stream));
	// "no row is shared" shortcut (decided on the device): vertex / face lists that every region binds, first-reference
	// ownership (no drained type symbols), encode side
	dl.d_dup_active = false;
	if (!decode && everywhere && cls != CLS_CORNER && n) {
		HB_TRY(hb_dalloc_t(m, &dl.d_dup, 1));
		HB_CUDA(ctx, cudaMemsetAsync(dl.d_dup, 0, sizeof(uint32_t), ctx->stream));
		dl.d_dup_active = true;
	}
	uint32_t *const dup_arg = dl.d_dup_active ? dl.d_dup : nullptr;
	if (n) {
		const uint32_t g = hb_div_up(n, 256);
		RowSeg rs;
		rs.nseg = m->nseg;
		rs.ent_base = cls == CLS_VTX ? m->d_vbase : cls == CLS_FACE ? m->d_fbase : m->d_ebase;
		rs.rowbase = dl.d_rowbase; rs.rownum = dl.d_rownum;
		if (cls == CLS_VTX) HB_LAUNCH(ctx, k_elem_rows<CLS_VTX>, g, 256, 0, c, l, rs, dl.d_erow, dl.d_ek, ctx->d_err);
		else if (cls == CLS_FACE) HB_LAUNCH(ctx, k_elem_rows<CLS_FACE>, g, 256, 0, c, l, rs, dl.d_erow, dl.d_ek, ctx->d_err);
		else HB_LAUNCH(ctx, k_elem_rows<CLS_CORNER>, g, 256, 0, c, l, rs, dl.d_erow, dl.d_ek, ctx->d_err);
		if (dl.d_ek) HB_TRY(hb_scan_exclusive_u32(ctx, dl.d_ek, dl.d_ek, n, nullptr));
		if (decode && dl.d_emit_type)
			HB_LAUNCH(ctx, k_owner_from_types, g, 256, 0, dl.d_erow, dl.d_ek, dl.d_emit_type, dl.emit_count, n, dl.d_first, ctx->d_err);
		else if (!decode && cls == CLS_CORNER && m->d_lh)
			HB_LAUNCH(ctx, k_first_ref_corner, g, 256, 0, dl.d_erow, n, m->d_he, m->d_celem_h, m->d_face_regs, m->d_slot_corner, (uint32_t)m->nlists, l, m->d_lh, m->n_corner_elems, dl.d_first);
		else
			HB_LAUNCH(ctx, k_first_ref, g, 256, 0, dl.d_erow, n, dl.d_first, dup_arg);
		if (!decode) { // ordinal of every DATA emission: positions in the symbol stream and history offsets (encode only)
			HB_LAUNCH(ctx, k_data_flags, g, 256, 0, dl.d_erow, dl.d_first, n, dl.d_dord, (const uint32_t *)dup_arg);
			HB_TRY(hb_scan_exclusive_u32(ctx, dl.d_dord, dl.d_dord, n, nullptr, dup_arg));
		}
		if (need_rp && dl.p.ncomp) {
			HB_TRY(hb_dalloc_t(m, &dl.d_rp, (size_t)n * dl.p.ncomp + 1));
			HB_LAUNCH(ctx, k_gather_rp, g, 256, 0, dl.p, dl.d_erow, n, dl.d_rp);
		}
	} else if (dl.d_ek) {
		HB_CUDA(ctx, cudaMemsetAsync(dl.d_ek, 0, sizeof(uint32_t) * 2, ctx->stream));
	}
	return 0;
}

static int build_lhist(hb_dmesh *m)
{
	hb_ctx *ctx = m->ctx;
	if (m->d_lh || !m->n_corner_elems || !m->nb_corner) return 0;
	const uint32_t ncel = m->n_corner_elems;
	uint32_t *off = nullptr, *cursor = nullptr, *vinc = nullptr, *dist = nullptr;
	HB_TRY(hb_dalloc_t(m, &off, (size_t)m->nv + 2));
	HB_TRY(hb_dalloc_t(m, &cursor, (size_t)m->nv + 1));
	HB_TRY(hb_dalloc_t(m, &vinc, (size_t)ncel + 1));
	HB_TRY(hb_dalloc_t(m, &dist, (size_t)ncel + 1));
	HB_TRY(hb_dalloc_t(m, &m->d_lh, (size_t)m->nb_corner * ncel + 1));
	HB_CUDA(ctx, cudaMemsetAsync(off, 0, sizeof(uint32_t) * ((size_t)m->nv + 2), ctx->stream));
	HB_CUDA(ctx, cudaMemsetAsync(cursor, 0, sizeof(uint32_t) * ((size_t)m->nv + 1), ctx->stream));
	HB_CUDA(ctx, cudaMemsetAsync(m->d_lh, 0xff, sizeof(uint32_t) * ((size_t)m->nb_corner * ncel + 1), ctx->stream));
	const uint32_t g = hb_div_up(ncel, 256);
	HB_LAUNCH(ctx, k_vinc_count, g, 256, 0, m->d_he, m->d_celem_h, m->d_face_regs, m->d_reg_ncorner, ncel, off);
	HB_TRY(hb_scan_exclusive_u32(ctx, off, off, m->nv, nullptr));
	HB_LAUNCH(ctx, k_vinc_fill, g, 256, 0, m->d_he, m->d_celem_h, m->d_face_regs, m->d_reg_ncorner, ncel, off, cursor, vinc);
	HB_LAUNCH(ctx, k_lhist, hb_div_up(m->nv, 128), 128, 0, m->d_he, m->d_celem_h, m->d_face_regs, m->d_reg_ncorner, m->d_bind_corner, (uint32_t)m->nb_corner, m->nv, off, vinc, dist, ncel, m->d_lh);
	return 0;
}

// a list without components that every region binds: type symbols in two passes (see k_nocomp_first)
static int encode_nocomp(hb_dmesh *m, int l)
{
	hb_ctx *ctx = m->ctx;
	DevList &dl = m->lists[l];
	const int cls = dl.p.target;
	const uint32_t n = cls == CLS_VTX ? m->norder : m->norder_f;
	dl.n_elems = n;
	dl.nocomp_fast = true;
	dl.d_ek = nullptr;
	const size_t hist_pitch = 4;
	HB_TRY(hb_dalloc_t(m, &dl.d_erow, (size_t)n + 1));
	HB_TRY(hb_dalloc_t(m, &dl.d_first, (size_t)dl.p.nrows + 1));
	HB_TRY(hb_dalloc_t(m, &dl.d_type, (size_t)n + 1));
	HB_TRY(hb_dalloc_t(m, &dl.d_hist, hist_pitch * m->nseg));
	dl.d_type_hist = dl.d_hist;
	HB_TRY(hb_dalloc_t(m, &dl.d_dup, 1));
	HB_CUDA(ctx, cudaMemsetAsync(dl.d_first, 0xff, sizeof(uint32_t) * ((size_t)dl.p.nrows + 1), ctx->stream));
	HB_CUDA(ctx, cudaMemsetAsync(dl.d_hist, 0, sizeof(unsigned long long) * hist_pitch * m->nseg, ctx->stream));
	HB_CUDA(ctx, cudaMemsetAsync(dl.d_type, 0, (size_t)n + 1, ctx->stream));
	HB_CUDA(ctx, cudaMemsetAsync(dl.d_dup, 0, sizeof(uint32_t), ctx->stream));
	if (!n) return 0;
	NoCompArgs a;
	a.order_f = m->has_order_f ? (const uint32_t *)m->d_order_f : nullptr;
	a.ord_v = m->d_ord_v;
	a.regs = cls == CLS_VTX ? m->d_vtx_regs : m->d_face_regs;
	a.slot = cls == CLS_VTX ? m->d_slot_vtx : m->d_slot_face;
	a.bind = cls == CLS_VTX ? m->d_bind_vtx : m->d_bind_face;
	a.nb = cls == CLS_VTX ? m->nb_vtx : m->nb_face;
	a.nlists = m->nlists; a.n = n; a.l = l;
	a.rs.nseg = m->nseg;
	a.rs.ent_base = cls == CLS_VTX ? m->d_vbase : m->d_fbase;
	a.rs.rowbase = dl.d_rowbase; a.rs.rownum = dl.d_rownum;
	a.elem_base = cls == CLS_VTX ? m->d_obase : m->d_ofbase;
	a.erow = dl.d_erow; a.first = dl.d_first; a.type = dl.d_type; a.type_hist = dl.d_type_hist; a.hist_pitch = hist_pitch;
	const uint32_t g = hb_div_up(n, 256);
	if (cls == CLS_VTX) {
		HB_LAUNCH(ctx, k_nocomp_first<CLS_VTX>, g, 256, 0, a, dl.d_dup, ctx->d_err);
		HB_LAUNCH(ctx, k_nocomp_types<CLS_VTX>, g, 256, 0, a, dl.d_dup, ctx->d_err);
	} else {
		HB_LAUNCH(ctx, k_nocomp_first<CLS_FACE>, g, 256, 0, a, dl.d_dup, ctx->d_err);
		HB_LAUNCH(ctx, k_nocomp_types<CLS_FACE>, g, 256, 0, a, dl.d_dup, ctx->d_err);
	}
	return 0;
}

// the streams of a zero-component list are being fetched and it has HIST emissions: their history offsets
int hb_nocomp_finish(hb_dmesh *m, int l, cudaStream_t st)
{
	hb_ctx *ctx = m->ctx;
	DevList &dl = m->lists[l];
	const uint32_t n = dl.n_elems;
	HB_TRY(hb_dalloc_t(m, &dl.d_dord, (size_t)n + 2));
	HB_TRY(hb_dalloc_t(m, &dl.d_aux, (size_t)n + 1));
	cudaStream_t keep = ctx->stream;
	ctx->stream = st; // the helpers launch on the context stream
	int rc = 0;
	do {
		const uint32_t g = hb_div_up(n, 256);
		k_data_flags<<<g, 256, 0, st>>>(dl.d_erow, dl.d_first, n, dl.d_dord, (const uint32_t *)nullptr);
		ctx->launches++;
		if ((rc = hb_scan_exclusive_u32(ctx, dl.d_dord, dl.d_dord, n, nullptr)) != 0) break;
		k_nocomp_aux<<<g, 256, 0, st>>>(dl.d_erow, dl.d_first, dl.d_dord, n, dl.d_aux);
		ctx->launches++;
	} while (0);
	ctx->stream = keep;
	return rc;
}

int hb_encode_lists(hb_dmesh *m)
{
	hb_ctx *ctx = m->ctx;
	HB_TRY(hb_build_conn(m));
	bool need_v = false, need_c = false;
	for (int l = 0; l < m->nlists; ++l) {
		const int cls = m->lists[l].p.target;
		if (cls == CLS_VTX && m->lists[l].p.ncomp) need_v = true;
		if (cls == CLS_CORNER) need_c = true;
	}
	if (need_v) HB_TRY(hb_build_vertex_candidates(m));
	if (need_c && m->any_corner) {
		HB_TRY(hb_build_corner_candidates(m));
		HB_TRY(build_lhist(m));
	}
	for (int l = 0; l < m->nlists; ++l) {
		DevList &dl = m->lists[l];
		const ListParams &p = dl.p;
		const int cls = p.target;
		if (cls != CLS_VTX && cls != CLS_FACE && cls != CLS_CORNER) { dl.n_elems = 0; continue; }
		if (cls == CLS_CORNER && !m->any_corner) { dl.n_elems = 0; continue; }
		dl.nocomp_fast = false;
		if (p.ncomp == 0 && (cls == CLS_FACE || cls == CLS_VTX) && bound_everywhere(m, cls, l)) {
			HB_TRY(encode_nocomp(m, l));
			continue;
		}
		const bool packed = cls == CLS_VTX && packed_eligible(p);
		HB_TRY(hb_prepare_list_elems(m, l, cls != CLS_FACE && !packed, false));
		const uint32_t n = dl.n_elems;
		HB_TRY(hb_dalloc_t(m, &dl.d_type, (size_t)n + 1));
		HB_TRY(hb_dalloc_t(m, &dl.d_aux, (size_t)n + 1));
		HB_TRY(hb_dalloc_t(m, &dl.d_sym, (size_t)n * p.sym_stride + 8));
		const size_t hist_pitch = (size_t)p.sym_stride * 256 + 4; // per segment: the contexts, then the four type counters
		HB_TRY(hb_dalloc_t(m, &dl.d_hist, hist_pitch * m->nseg));
		dl.d_type_hist = dl.d_hist + (size_t)p.sym_stride * 256;
		HB_CUDA(ctx, cudaMemsetAsync(dl.d_hist, 0, sizeof(unsigned long long) * hist_pitch * m->nseg, ctx->stream));
		if (!n) continue;
		EncodeArgs a;
		a.erow = dl.d_erow; a.ek = dl.d_ek; a.first = dl.d_first; a.dord = dl.d_dord;
		a.rp = dl.d_rp;
		a.cand_off = cls == CLS_VTX ? m->d_vc_off : m->d_cc_off;
		a.cand = cls == CLS_VTX ? m->d_vc_tri : m->d_cc_idx;
		a.lh = m->d_lh; a.ncel = m->n_corner_elems;
		a.face_regs = m->d_face_regs; a.slot_corner = m->d_slot_corner; a.nlists = m->nlists;
		a.celem_h = m->d_celem_h; a.he = m->d_he;
		a.type = dl.d_type; a.aux = dl.d_aux; a.sym = dl.d_sym; a.hist = dl.d_hist; a.type_hist = dl.d_type_hist;
		a.n = n; a.l = l;
		a.dup = dl.d_dup_active ? dl.d_dup : nullptr;
		if (a.dup) { // the fast path leaves the all-zero type / offset streams alone
			HB_CUDA(ctx, cudaMemsetAsync(dl.d_type, 0, (size_t)n + 1, ctx->stream));
			HB_CUDA(ctx, cudaMemsetAsync(dl.d_aux, 0, sizeof(uint32_t) * ((size_t)n + 1), ctx->stream));
		}
		a.nseg = m->nseg;
		a.elem_base = cls == CLS_VTX ? m->d_obase : cls == CLS_FACE ? m->d_ofbase : m->d_cebase;
		a.hist_pitch = hist_pitch;
		a.wide = nullptr; a.wide_cap = 0;
		if (cls == CLS_VTX && p.ncomp) {
			a.wide_cap = 4096 + 4 * m->nseg;
			HB_TRY(hb_dalloc_t(m, &dl.d_wide, (size_t)a.wide_cap + 1));
			HB_CUDA(ctx, cudaMemsetAsync(dl.d_wide, 0, sizeof(uint32_t), ctx->stream));
			a.wide = dl.d_wide;
		}
		const int nctx_s = (int)(p.sym_stride < HIST_SMEM_CTX ? p.sym_stride : HIST_SMEM_CTX);
		const size_t smem = sizeof(uint32_t) * ((size_t)nctx_s * 256 + 4);
		const uint32_t g = hb_div_up(n, ENC_THREADS);
		if (packed) {
			if (p.uniform_stype == HB_UCHAR) HB_TRY(encode_vtx_packed<uint8_t>(m, dl, a));
			else if (p.uniform_stype == HB_USHORT) HB_TRY(encode_vtx_packed<uint16_t>(m, dl, a));
			else HB_TRY(encode_vtx_packed<uint32_t>(m, dl, a));
		} else if (cls == CLS_VTX) {
			HB_LAUNCH(ctx, k_encode_main<CLS_VTX>, g, ENC_THREADS, smem, p, a);
			if (a.wide) HB_LAUNCH(ctx, k_encode_wide, hb_div_up((uint64_t)a.wide_cap * 32, 128), 128, 0, p, a);
		}
		else if (cls == CLS_FACE) HB_LAUNCH(ctx, k_encode_main<CLS_FACE>, g, ENC_THREADS, smem, p, a);
		else HB_LAUNCH(ctx, k_encode_main<CLS_CORNER>, g, ENC_THREADS, smem, p, a);
	}
	m->encoded = true;
	return 0;
}
