// hb_decode.cu -- decode side: reconstruction of attribute values from residual rows
// (AttrDecoder<RD>::decode, formats/hry/attrcode.h:443-550, after the host drained the symbol
// stream -- SURVEY.md 3.2).
//
// Scheduling (DESIGN.md "Decode"): a row can be reconstructed once every candidate row of its
// prediction is reconstructed.
//   * FACE lists: prediction is always 0 -> embarrassingly parallel map.
//   * CORNER lists: shallow DAG (depth <= fan size).  Level-synchronous wavefronts: each sweep
//     reconstructs every element whose candidates are all done; levels are discovered on the fly.
//   * VTX lists: the DAG over the traversal order is a CHAIN (every new vertex is predicted across
//     a gate that ends in the vertex decoded just before it; depth ~ N).  One thread per
//     (list, component) walks the chain in traversal order over rank-space records; parallelism
//     comes from components, lists and -- for batches -- meshes.
#include "hb_lists.cuh"
#include "hb_decode_spec.cuh"
#include "hb_decode_scan.cuh"
#include <math.h>
#include <stdlib.h>

// ------------------------------------------------------------------------------------------------
// VTX: sequential walker for the lists the scan kernel does not take -- lossless float lists (the prediction picks
// ONE candidate, `closest to the mean` in float arithmetic: no map algebra to compose) and lists with mixed or wide
// storage types.  The chain x[i] = f(x[i - 1], older values) is walked in order by ONE thread per (segment, component)
// -- components are independent chains -- but that thread only does the dependent arithmetic.  Its warp works as a
// two-stage pipeline, 32 ranks per batch:
//   PREPARE (all lanes, one rank each, one batch ahead): element kind, candidate triples, residual, and every operand
//           value that is already final (from a shared-memory ring of the last WALK_RING values, else from L2), parked
//           as a 128-byte work item in shared memory.  Its loads are in flight while lane 0 executes.
//   EXECUTE (lane 0): the items of the current batch in order; only operands produced inside the last two batches are
//           read at that point (from the ring).  get_prediction / decodeDelta exactly as in hb_lists.cuh.
// ------------------------------------------------------------------------------------------------
struct WalkArgs {
	int ncomp, comp;             // this chain reconstructs component `comp`
	int stype, quant;
	const uint32_t *erow, *first, *cand_off, *cand;
	unsigned long long *rp;      // rank-space containers, ncomp per element: residual in, value out
	uint32_t base, n; // ranks [base, n) of one segment
};
#define WALK_RING 2048
#define WALK_KIN 4

// mean of the candidates and the prediction from it (attrcode.h:182-208), candidates given by `get(k)`; the division by
// 1, 2 or 4 candidates is an exact scaling (same bits as the IEEE division, without its latency)
template <typename Get>
__device__ __forceinline__ unsigned long long walk_combine(int st, uint32_t K, Get &&get)
{
	if (K == 0) return 0;
	if (st == HB_FLOAT && (K == 1 || K == 2 || K == 4)) {
		double sum = 0.0;
		float p[4];
#pragma unroll
		for (uint32_t k = 0; k < 4; ++k)
			if (k < K) { p[k] = __uint_as_float((uint32_t)get(k)); sum = __dadd_rn(sum, (double)p[k]); }
		const float avg = __double2float_rn(__dmul_rn(sum, K == 1 ? 1.0 : (K == 2 ? 0.5 : 0.25)));
		float res = FLT_MAX;
#pragma unroll
		for (uint32_t k = 0; k < 4; ++k)
			if (k < K) res = hb_closest_step(res, p[k], avg);
		return __float_as_uint(res);
	}
	return combine_candidates(st, K, get);
}

__global__ void __launch_bounds__(32) k_decode_vertex_walk(const WalkArgs *__restrict__ args)
{
	const WalkArgs a = args[blockIdx.x];
	const uint32_t lane = threadIdx.x;
	const int st = a.stype, q = a.quant, nc = a.ncomp, j = a.comp;
	const uint32_t base = a.base, n = a.n;
	const uint32_t *__restrict__ erow = a.erow, *__restrict__ first = a.first, *__restrict__ coff = a.cand_off, *__restrict__ cand = a.cand;
	unsigned long long *rp = a.rp;
	__shared__ unsigned long long s_ring[WALK_RING];
	__shared__ __align__(16) unsigned long long s_item[2][32][16];
	if (base >= n) return;
	auto ring = [&](uint32_t r) -> unsigned long long & { return s_ring[(r - base) & (WALK_RING - 1)]; };
	// a final value: from the ring if it is still there (nothing at or behind `front` has been produced yet), else from L2
	auto final_value = [&](uint32_t r, uint32_t front) -> unsigned long long {
		if (r + WALK_RING >= front + 64) return ring(r);
		return __ldcg(rp + (size_t)r * nc + j);
	};
	// registers of PREPARE between issue() and finish()
	uint32_t p_kind = 0, p_K = 0, p_c0 = 0, p_late = 0, p_src = 0;
	unsigned long long p_res = 0, p_op[3 * WALK_KIN];
	// issue(b0): ranks [b0, b0 + 32); everything below `front` (the start of the batch that executes meanwhile) is final
	auto issue = [&](uint32_t b0, uint32_t front) {
		const uint32_t i = b0 + lane;
		p_kind = 0; p_K = 0; p_late = 0; p_res = 0; p_src = 0; p_c0 = 0;
		if (i >= n) return;
		const uint32_t row = __ldg(erow + i);
		if (row == HB_NONE) return;
		const uint32_t fi = __ldg(first + row);
		p_res = __ldcg(rp + (size_t)i * nc + j);
		if (fi != i) { // HIST reference: the value of the owning element (attrcode.h:463-466)
			p_kind = 2;
			p_src = fi;
			if (fi < front) { p_res = final_value(fi, front); p_kind = 0; }
			else if (fi > i) p_kind = 0; // not emitted yet: rows start zeroed -- keeps what is there
			return;
		}
		p_kind = 1;
		p_c0 = __ldg(coff + i);
		p_K = __ldg(coff + i + 1) - p_c0;
		if (p_K > WALK_KIN) { p_kind = 3; return; }
#pragma unroll
		for (int w = 0; w < 3 * WALK_KIN; ++w) {
			p_op[w] = 0;
			if ((uint32_t)(w / 3) < p_K) {
				const uint32_t r = __ldg(cand + 3 * (size_t)p_c0 + w);
				if (r < front) p_op[w] = final_value(r, front);
				else { p_late |= 1u << w; p_op[w] = r; }
			}
		}
	};
	auto finish = [&](uint32_t buf) {
		unsigned long long *it = &s_item[buf][lane][0];
		it[0] = (unsigned long long)(p_kind | (p_K << 2) | (p_late << 8)) | ((unsigned long long)(p_kind == 2 ? p_src : p_c0) << 32);
		it[1] = p_res;
		if (p_kind == 1) {
#pragma unroll
			for (int w = 0; w < 3 * WALK_KIN; ++w) it[2 + w] = p_op[w];
		}
	};
	issue(base, base);
	finish(0);
	__syncwarp();
	for (uint32_t b0 = base; b0 < n; b0 += 32) {
		const uint32_t buf = ((b0 - base) >> 5) & 1u;
		const bool more = b0 + 32 < n;
		if (more) issue(b0 + 32, b0);
		if (lane == 0) {
			const uint32_t nl = min(32u, n - b0);
#pragma unroll 1
			for (uint32_t t = 0; t < nl; ++t) {
				const unsigned long long *it = &s_item[buf][t][0];
				const unsigned long long h = it[0];
				const uint32_t kind = (uint32_t)h & 3u, K = ((uint32_t)h >> 2) & 0x3fu, late = ((uint32_t)h >> 8) & 0xfffu, aux = (uint32_t)(h >> 32);
				const uint32_t i = b0 + t;
				unsigned long long val = it[1];
				if (kind == 1) {
					const unsigned long long pred = walk_combine(st, K, [&](uint32_t k) -> unsigned long long {
						unsigned long long v[3];
#pragma unroll
						for (int o = 0; o < 3; ++o) {
							const unsigned long long x = it[2 + 3 * k + o];
							v[o] = ((late >> (3 * k + o)) & 1u) ? ring((uint32_t)x) : x;
						}
						return hb_predict(st, v[0], v[1], v[2], q);
					});
					val = hb_dec(st, val, pred, q);
				} else if (kind == 2) {
					val = ring(aux); // owner inside the last two batches
				} else if (kind == 3) {
					// more candidates than an item holds (poles, closing vertices): straight from the CSR
					const uint32_t Kf = __ldg(coff + i + 1) - aux;
					const unsigned long long pred = combine_candidates(st, Kf, [&](uint32_t k) -> unsigned long long {
						unsigned long long v[3];
#pragma unroll
						for (int o = 0; o < 3; ++o) {
							const uint32_t r = __ldg(cand + 3 * (size_t)(aux + k) + o);
							v[o] = r + WALK_RING > i ? ring(r) : __ldcg(rp + (size_t)r * nc + j);
						}
						return hb_predict(st, v[0], v[1], v[2], q);
					});
					val = hb_dec(st, val, pred, q);
				}
				ring(i) = val;
				rp[(size_t)i * nc + j] = val;
			}
		}
		__syncwarp();
		if (more) finish(buf ^ 1u);
		__syncwarp();
	}
}

// ------------------------------------------------------------------------------------------------
// Lossless float vertex lists: which reconstruction per segment?
//   chain-like traversal (triangle meshes: almost every vertex is predicted across a gate that ends in the vertex
//   coded just before it, the DAG is a chain of depth ~N)          -> k_decode_vertex_walkf, a pipelined sequential walk
//   shallow DAG (polygon meshes: candidates come from the own face) -> k_decode_vertex_spec, chunk-parallel Jacobi sweeps
// k_chain_stat counts, per segment, the DATA ranks that read rank i - 1; both kernels are launched and the one that is
// not chosen for a segment returns at once (no host synchronisation).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_chain_stat(const uint8_t *__restrict__ kind, const uint32_t *__restrict__ cand_off, const uint32_t *__restrict__ cand, uint32_t n,
                                                    const uint32_t *__restrict__ elem_base, uint32_t nseg, uint32_t *__restrict__ chain /* [nseg][2]: predecessor readers, DATA ranks */)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n || kind[i] != 1) return;
	const uint32_t c0 = cand_off[i], K = min(cand_off[i + 1] - c0, 8u);
	bool pred = false;
	for (uint32_t w = 0; w < 3 * K; ++w) pred = pred || cand[3 * (size_t)c0 + w] + 1 == i;
	const uint32_t s = hb_seg_find(elem_base, nseg, i);
	atomicAdd(&chain[2 * s + 1], 1u);
	if (pred) atomicAdd(&chain[2 * s], 1u);
}
// measured: sphere 99.99 % predecessor readers (walk 14 ms, Jacobi 80 ms for 35K vertices), polygon grid 58.6 % (walk 150 ms,
// Jacobi 39 ms for 361K vertices) -- three quarters is the line
__device__ __forceinline__ bool chain_like(const uint32_t *chain, uint32_t seg) { return 4ull * chain[2 * seg] >= 3ull * chain[2 * seg + 1]; }

#define WALKF_RING 1024
// L2-coherent load of a whole compact record (4, 8 or 16 bytes)
template <int NC> __device__ __forceinline__ SpecRec<uint32_t, NC> walkf_ldcg(const SpecRec<uint32_t, NC> *p)
{
	SpecRec<uint32_t, NC> r;
	constexpr int RS = (int)(sizeof(r) / 4);
	if (RS == 4) { const uint4 v = __ldcg((const uint4 *)p); r.c[0] = v.x; r.c[1] = v.y; r.c[RS > 2 ? 2 : 0] = v.z; r.c[RS > 3 ? 3 : 0] = v.w; }
	else if (RS == 2) { const uint2 v = __ldcg((const uint2 *)p); r.c[0] = v.x; r.c[RS > 1 ? 1 : 0] = v.y; }
	else r.c[0] = __ldcg((const unsigned int *)p);
	return r;
}
// One warp per segment.  PREPARE: 32 lanes = the 32 ranks of the next batch (records of all components at once);
// EXECUTE: lanes 0..NC-1 = the components of the list, which share kinds, candidates and late operands and so walk the
// chain in lockstep.  Arithmetic of spec_step<.., FP = true> (attrcode.h:182-208, prediction.h:64-72).
__device__ __forceinline__ uint32_t walkf_lds(uint32_t addr) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }
__device__ __forceinline__ uint4 walkf_lds4(uint32_t addr)
{
	uint4 v;
	asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
	return v;
}
__device__ __forceinline__ void walkf_sts(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }

template <int NC>
__global__ void __launch_bounds__(32) k_decode_vertex_walkf(const SpecArgs *__restrict__ args, const uint32_t *__restrict__ chain)
{
	typedef SpecRec<uint32_t, NC> Rec;
	constexpr int RS = (int)(sizeof(Rec) / 4);
	// work item of one rank (words): 0 kind | K << 2, 1 aux (CSR offset / owning rank), 2 K in full, 3 pad, 4.. residual record,
	// then per candidate four words: the shared-memory ADDRESSES of its three operand records (a value parked in the item
	// itself, or the ring slot a rank of the last two batches will have filled by then), then the parked operand records.
	// EXECUTE reads every operand with the same two loads, no selects.
	constexpr int OFF_ADDR = (4 + RS + 3) & ~3, OFF_VAL = OFF_ADDR + 16, ITEM = (OFF_VAL + 12 * RS + 3) & ~3; // 16-byte aligned vector loads
	if (!chain_like(chain, blockIdx.x)) return;
	const SpecArgs a = args[blockIdx.x];
	const uint32_t lane = threadIdx.x;
	const uint32_t base = a.base, n = a.n;
	const Rec *__restrict__ resid = (const Rec *)a.resid;
	Rec *x = (Rec *)a.x;
	__shared__ Rec s_ring[WALKF_RING];
	__shared__ __align__(16) uint32_t s_item[2][32][ITEM];
	if (base >= n) return;
	uint32_t ringb, itemb; // shared-window addresses, pinned in registers
	asm volatile("mov.u32 %0, %1;" : "=r"(ringb) : "r"((uint32_t)__cvta_generic_to_shared(s_ring)));
	asm volatile("mov.u32 %0, %1;" : "=r"(itemb) : "r"((uint32_t)__cvta_generic_to_shared(s_item)));
	auto ring = [&](uint32_t r) -> Rec & { return s_ring[(r - base) & (WALKF_RING - 1)]; };
	auto ring_addr = [&](uint32_t r) -> uint32_t { return ringb + 4u * RS * ((r - base) & (WALKF_RING - 1)); };
	auto final_value = [&](uint32_t r, uint32_t front) -> Rec {
		if (r + WALKF_RING >= front) return ring(r);
		return walkf_ldcg<NC>(x + r);
	};
	uint32_t p_kind = 0, p_K = 0, p_late = 0, p_aux = 0;
	Rec p_res, p_op[12];
	auto issue = [&](uint32_t b0, uint32_t front) {
		const uint32_t i = b0 + lane;
		p_kind = 0; p_K = 0; p_late = 0; p_aux = 0;
		if (i >= n) return;
		const uint32_t kd = a.kind[i];
		if (kd == 0) { p_res = walkf_ldcg<NC>(x + i); return; } // not bound: keeps what is there
		if (kd == 2) {
			const uint32_t fi = a.src[i];
			if (fi < front) p_res = final_value(fi, front);
			else { p_kind = 2; p_aux = fi; }
			return;
		}
		p_res = resid[i];
		p_aux = a.cand_off[i];
		p_K = a.cand_off[i + 1] - p_aux;
		if (p_K > 4) { p_kind = 3; return; }
		p_kind = 1;
#pragma unroll
		for (int w = 0; w < 12; ++w) {
			if ((uint32_t)(w / 3) < p_K) {
				const uint32_t r = a.cand[3 * (size_t)p_aux + w];
				if (r < front) p_op[w] = final_value(r, front);
				else { p_late |= 1u << w; p_op[w].c[0] = r; }
			}
		}
	};
	auto finish = [&](uint32_t buf) {
		uint32_t *it = &s_item[buf][lane][0];
		const uint32_t ib = itemb + 4u * ITEM * (buf * 32u + lane);
		it[0] = p_kind | (min(p_K, 63u) << 2);
		it[1] = p_kind == 2 ? ring_addr(p_aux) : p_aux;
		it[2] = p_K; // in full (a pole has hundreds of candidates)
#pragma unroll
		for (int e = 0; e < RS; ++e) it[4 + e] = p_res.c[e];
		if (p_kind == 1) {
#pragma unroll
			for (int w = 0; w < 12; ++w)
				if ((uint32_t)(w / 3) < p_K) {
					const bool late = (p_late >> w) & 1u;
					it[OFF_ADDR + 4 * (w / 3) + (w % 3)] = late ? ring_addr(p_op[w].c[0]) : ib + 4u * (OFF_VAL + w * RS);
					if (!late) {
#pragma unroll
						for (int e = 0; e < RS; ++e) it[OFF_VAL + w * RS + e] = p_op[w].c[e];
					}
				}
		}
	};
	const uint32_t j = lane < (uint32_t)NC ? lane : 0u; // my component in EXECUTE
	issue(base, base);
	finish(0);
	__syncwarp();
	for (uint32_t b0 = base; b0 < n; b0 += 32) {
		const uint32_t buf = ((b0 - base) >> 5) & 1u;
		const bool more = b0 + 32 < n;
		if (more) issue(b0 + 32, b0);
		if (lane < (uint32_t)NC) {
			const uint32_t nl = min(32u, n - b0);
			uint32_t ib = itemb + 4u * ITEM * (buf * 32u);
			uint32_t ra = ring_addr(b0) + 4u * j;
			const uint32_t ring_end = ringb + 4u * RS * WALKF_RING;
			// one candidate prediction: v0 + (v1 - v2), operands through the parked addresses
			auto cand_pred = [&](uint32_t k) -> float {
				const uint4 ad = walkf_lds4(ib + 4u * (OFF_ADDR + 4 * k));
				const float v0 = __uint_as_float(walkf_lds(ad.x + 4u * j)), v1 = __uint_as_float(walkf_lds(ad.y + 4u * j)), v2 = __uint_as_float(walkf_lds(ad.z + 4u * j));
				return __fadd_rn(v0, __fsub_rn(v1, v2));
			};
#pragma unroll 1
			for (uint32_t t = 0; t < nl; ++t, ib += 4u * ITEM) {
				const uint4 h4 = walkf_lds4(ib);
				const uint32_t kind = h4.x & 3u, K = (h4.x >> 2) & 0x3fu, aux = h4.y;
				const uint32_t i = b0 + t;
				uint32_t val = walkf_lds(ib + 4u * (4 + j));
				if (kind == 1) {
					uint32_t pred = 0;
					if (K == 2) { // the usual vertex: two parallelograms
						const float p0 = cand_pred(0), p1 = cand_pred(1);
						const float avg = __double2float_rn(__dmul_rn(__dadd_rn((double)p0, (double)p1), 0.5)); // exact scaling == IEEE division by 2
						pred = __float_as_uint(hb_closest_step(hb_closest_step(FLT_MAX, p0, avg), p1, avg));
					} else if (K == 1) {
						pred = __float_as_uint(cand_pred(0)); // mean of one candidate: itself
					} else if (K) {
						double sum = 0.0;
						for (uint32_t k = 0; k < K; ++k) sum = __dadd_rn(sum, (double)cand_pred(k));
						const float avg = __double2float_rn(__ddiv_rn(sum, (double)(int)K));
						float best = FLT_MAX;
						for (uint32_t k = 0; k < K; ++k) best = hb_closest_step(best, cand_pred(k), avg);
						pred = __float_as_uint(best);
					}
					val = hb_flip_f32(IntOps<uint32_t>::dec(val, hb_flip_f32(pred), 32));
				} else if (kind == 2) {
					val = walkf_lds(aux + 4u * j);
				} else if (kind == 3) {
					// more candidates than an item holds (poles, closing vertices): straight from the CSR, two passes
					const uint32_t *tri = a.cand + 3 * (size_t)aux;
					const uint32_t Kf = h4.z;
					auto getv = [&](uint32_t r) -> float { return __uint_as_float(r + WALKF_RING > i ? ring(r).c[j] : __ldcg(&x[r].c[j])); };
					double sum = 0.0;
					for (uint32_t k = 0; k < Kf; ++k) sum = __dadd_rn(sum, (double)__fadd_rn(getv(tri[3 * k]), __fsub_rn(getv(tri[3 * k + 1]), getv(tri[3 * k + 2]))));
					const float avg = __double2float_rn(__ddiv_rn(sum, (double)(int)Kf));
					float best = FLT_MAX;
					for (uint32_t k = 0; k < Kf; ++k) best = hb_closest_step(best, __fadd_rn(getv(tri[3 * k]), __fsub_rn(getv(tri[3 * k + 1]), getv(tri[3 * k + 2]))), avg);
					val = hb_flip_f32(IntOps<uint32_t>::dec(val, hb_flip_f32(__float_as_uint(best)), 32));
				}
				walkf_sts(ra, val);
				ra += 4u * RS;
				if (ra >= ring_end) ra -= 4u * RS * WALKF_RING;
				x[i].c[j] = val;
			}
		}
		__syncwarp();
		if (more) finish(buf ^ 1u);
		__syncwarp();
	}
}
static int launch_walkf(hb_ctx *ctx, int ncomp, const SpecArgs *d_args, uint32_t nseg, const uint32_t *chain)
{
	switch (ncomp) {
	case 1: HB_LAUNCH(ctx, k_decode_vertex_walkf<1>, nseg, 32, 0, d_args, chain); break;
	case 2: HB_LAUNCH(ctx, k_decode_vertex_walkf<2>, nseg, 32, 0, d_args, chain); break;
	case 3: HB_LAUNCH(ctx, k_decode_vertex_walkf<3>, nseg, 32, 0, d_args, chain); break;
	default: HB_LAUNCH(ctx, k_decode_vertex_walkf<4>, nseg, 32, 0, d_args, chain); break;
	}
	return 0;
}

// ---- speculative path helpers ---------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_spec_prep(const uint32_t *__restrict__ erow, const uint32_t *__restrict__ first, uint32_t n, uint8_t *__restrict__ kind, uint32_t *__restrict__ src)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint32_t row = erow[i];
	uint8_t k = 0;
	uint32_t s = 0;
	if (row != HB_NONE) {
		const uint32_t fi = first[row];
		if (fi == i) k = 1;
		else if (fi < i) { k = 2; s = fi; }
	}
	kind[i] = k;
	src[i] = s;
}

// A vertex list is reconstructed in GROUPS of up to four components that share one storage type (u8 / u16 / u32 / float):
// every group gets compact rank-space records of its own and runs through the scan decoder (integers) or the float
// kernels; components of other storage types (signed, 64-bit) are left to the generic walker.
struct CompMap {
	int n;          // components of the group
	int8_t comp[4]; // their indices in the list
};
// AoS rows -> compact rank-space records (ncp elements of esize bytes per element)
__global__ void __launch_bounds__(256) k_gather_compact(ListParams p, CompMap cm, const uint32_t *__restrict__ erow, uint32_t n, uint8_t *__restrict__ out, int esize, int ncp)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint32_t row = erow[i];
	uint8_t *dst = out + (size_t)i * ncp * esize;
	if (row == HB_NONE) {
		for (int j = 0; j < ncp; ++j) hb_st_bits(dst + j * esize, esize, 0);
		return;
	}
	const uint8_t *srcp = p.rows + (size_t)row * p.stride;
	for (int j = 0; j < ncp; ++j) hb_st_bits(dst + j * esize, esize, j < cm.n ? hb_ld_bits(srcp + p.offset[cm.comp[j]], esize) : 0);
}

__global__ void __launch_bounds__(256) k_scatter_compact(ListParams p, CompMap cm, const uint32_t *__restrict__ erow, const uint8_t *__restrict__ kind, uint32_t n, const uint8_t *__restrict__ x, int esize, int ncp)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n || kind[i] != 1) return;
	uint8_t *dst = p.rows + (size_t)erow[i] * p.stride;
	const uint8_t *srcp = x + (size_t)i * ncp * esize;
	for (int j = 0; j < cm.n; ++j) hb_st_bits(dst + p.offset[cm.comp[j]], esize, hb_ld_bits(srcp + j * esize, esize));
}

template <typename T, int NC, bool FP>
static int launch_spec_nc(hb_ctx *ctx, const SpecArgs *d_args, uint32_t nseg, const uint32_t *chain)
{
	const int threads = spec_threads<T, NC>();
	const size_t smem = (size_t)threads * 4 * SPEC_HB * sizeof(SpecRec<T, NC>);
	HB_CUDA(ctx, cudaFuncSetAttribute(k_decode_vertex_spec<T, NC, FP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	HB_LAUNCH(ctx, (k_decode_vertex_spec<T, NC, FP>), nseg, threads, smem, d_args, chain);
	return 0;
}
template <typename T, bool FP>
static int launch_spec(hb_ctx *ctx, int ncomp, const SpecArgs *d_args, uint32_t nseg, const uint32_t *chain = nullptr)
{
	switch (ncomp) {
	case 1: return launch_spec_nc<T, 1, FP>(ctx, d_args, nseg, chain);
	case 2: return launch_spec_nc<T, 2, FP>(ctx, d_args, nseg, chain);
	case 3: return launch_spec_nc<T, 3, FP>(ctx, d_args, nseg, chain);
	default: return launch_spec_nc<T, 4, FP>(ctx, d_args, nseg, chain);
	}
}

// cluster launch of the verified-scan kernel (integer lists): one cluster per component
template <typename T, int NTB, int MINB>
static int launch_scan_ntb(hb_ctx *ctx, int ncomp, const SpecArgs *d_args, int cluster, uint32_t nseg)
{
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3((uint32_t)cluster * (uint32_t)ncomp * nseg); // one cluster per (segment, component)
	cfg.blockDim = dim3(NTB);
	cfg.dynamicSmemBytes = 0;
	cfg.stream = ctx->stream;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeClusterDimension;
	attr[0].val.clusterDim.x = cluster;
	attr[0].val.clusterDim.y = 1;
	attr[0].val.clusterDim.z = 1;
	cfg.attrs = attr;
	cfg.numAttrs = 1;
	if (cluster > 8) HB_CUDA(ctx, cudaFuncSetAttribute((k_decode_vertex_scan<T, NTB, MINB>), cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
	cudaEvent_t pa = nullptr, pb = nullptr;
	if (ctx->profiling) { pa = hb_prof_event(ctx); pb = hb_prof_event(ctx); cudaEventRecord(pa, ctx->stream); }
	HB_CUDA(ctx, cudaLaunchKernelEx(&cfg, (k_decode_vertex_scan<T, NTB, MINB>), d_args, (uint32_t)ncomp));
	ctx->launches++;
	if (pa) { cudaEventRecord(pb, ctx->stream); ctx->prof.push_back(hb_ctx::ProfRec{ "k_decode_vertex_scan", pa, pb }); }
	return 0;
}
static int scan_ntb(uint32_t chains, int sm_count)
{
	const char *env = getenv("HARRY_B200_SCAN_NTB"); // (read per call: tests switch it)
	if (env && (atoi(env) == 128 || atoi(env) == 256 || atoi(env) == 512)) return atoi(env);
	// the fixed cost of a sweep (two dependent memory round trips, four barriers, ~1000 dependent instructions) barely
	// depends on the window length: the more chains an SM holds, the better it hides it (measured, 197 spheres of 100K
	// vertices: 512 threads 8.3 ms per 196 meshes, 256: 6.3 ms, 128: 5.1 ms)
	if (2u * chains >= 7u * (uint32_t)sm_count) return 128;
	return chains >= 2u * (uint32_t)sm_count - 2u ? 256 : 512;
}
template <typename T>
static int launch_scan(hb_ctx *ctx, int ncomp, const SpecArgs *d_args, int cluster, uint32_t nseg, int ntb)
{
	if (ntb == 128) {
		// more chains than four per SM: the 64-register build keeps eight of them resident per SM
		const char *env = getenv("HARRY_B200_SCAN_DENSE"); // A/B runs and tests (read per call): 0 = never, 1 = always
		const bool dense = env ? atoi(env) != 0 : (uint64_t)nseg * (uint32_t)ncomp > 4ull * (uint32_t)ctx->sm_count;
		if (dense && cluster == 1) return launch_scan_ntb<T, 128, 8>(ctx, ncomp, d_args, cluster, nseg);
		return launch_scan_ntb<T, 128, 4>(ctx, ncomp, d_args, cluster, nseg);
	}
	if (ntb == 256) return launch_scan_ntb<T, 256, 2>(ctx, ncomp, d_args, cluster, nseg);
	return launch_scan_ntb<T, 512, 1>(ctx, ncomp, d_args, cluster, nseg);
}
// CTAs per cluster: the window is about one cut-border length (~ sqrt(2 n) on a regular mesh);
// two ranks per thread
static int scan_cluster_size(uint32_t n, uint32_t nseg, int ntb)
{
	static const char *env = getenv("HARRY_B200_SCAN_CLUSTER");
	if (env && atoi(env) > 0) return atoi(env) > SCAN_MAXC ? SCAN_MAXC : atoi(env);
	// a batch fills the machine with chains: one CTA per chain (no cluster hand-offs)
	if (nseg > 1 && ntb < 512) return 1;
	const double need = (nseg > 1 ? 0.9 : 1.2) * sqrt(2.0 * (double)n) / (double)ntb;
	int c = 1;
	while (c < SCAN_MAXC && (double)c < need) c <<= 1;
	return c;
}

// components the compact kernels take: unsigned integer storage types up to 32 bits, and float
static bool group_type(int st) { return st == HB_UCHAR || st == HB_USHORT || st == HB_UINT || st == HB_FLOAT; }
// groups of up to four components of one storage type, in list order; `rest` = the components left to the generic walker
static void plan_groups(const ListParams &p, std::vector<SpecGroup> &out, std::vector<int> &rest)
{
	std::vector<SpecGroup> groups;
	rest.clear();
	for (int j = 0; j < p.ncomp; ++j) {
		const int st = p.stype[j];
		if (!group_type(st)) { rest.push_back(j); continue; }
		SpecGroup *g = nullptr;
		for (SpecGroup &c : groups)
			if (c.st == st && c.ncomp < 4) g = &c;
		if (!g) { groups.push_back(SpecGroup()); g = &groups.back(); g->st = st; g->ncomp = 0; }
		g->comp[g->ncomp++] = j;
	}
	// the same plan as in the last decode of this mesh: keep it, its buffers are reused
	bool same = groups.size() == out.size();
	for (size_t k = 0; same && k < groups.size(); ++k) {
		same = groups[k].st == out[k].st && groups[k].ncomp == out[k].ncomp;
		for (int j = 0; same && j < groups[k].ncomp; ++j) same = groups[k].comp[j] == out[k].comp[j];
	}
	if (!same) out = groups;
}

// reconstruct the component groups of one vertex list (compact records per group; connectivity records shared)
static int decode_vertex_groups(hb_dmesh *m, int l)
{
	hb_ctx *ctx = m->ctx;
	DevList &dl = m->lists[l];
	const ListParams &p = dl.p;
	const uint32_t n = dl.n_elems;
	const uint32_t nseg = m->nseg;
	const uint32_t g = hb_div_up(n, 256);
	HB_TRY(hb_dalloc_t(m, &dl.d_kind, (size_t)n + 1));
	HB_TRY(hb_dalloc_t(m, &dl.d_src, (size_t)n + 1));
	HB_TRY(hb_dalloc_t(m, &dl.d_spec_stats, 8));
	HB_CUDA(ctx, cudaMemsetAsync(dl.d_spec_stats, 0, 64, ctx->stream));
	HB_LAUNCH(ctx, k_spec_prep, g, 256, 0, dl.d_erow, dl.d_first, n, dl.d_kind, dl.d_src);
	bool any_int = false, any_float = false;
	for (const SpecGroup &sg : dl.groups) { any_int = any_int || sg.st != HB_FLOAT; any_float = any_float || sg.st == HB_FLOAT; }
	ScanRec *srec = nullptr;
	if (any_int) {
		HB_TRY(hb_dalloc(m, &dl.d_srec, sizeof(ScanRec) * ((size_t)n + 1)));
		srec = (ScanRec *)dl.d_srec;
		HB_LAUNCH(ctx, k_scan_prep, g, 256, 0, dl.d_kind, dl.d_src, m->d_vc_off, m->d_vc_tri, n, srec);
	}
	if (any_float) {
		// lossless float components: sequential walk for chain-like segments, Jacobi sweeps for shallow ones (decided on the device)
		HB_TRY(hb_dalloc_t(m, &dl.d_chain, 2 * (size_t)nseg));
		HB_CUDA(ctx, cudaMemsetAsync(dl.d_chain, 0, sizeof(uint32_t) * 2 * nseg, ctx->stream));
		HB_LAUNCH(ctx, k_chain_stat, g, 256, 0, dl.d_kind, m->d_vc_off, m->d_vc_tri, n, m->d_obase, nseg, dl.d_chain);
		HB_CUDA(ctx, cudaMemcpyAsync(dl.d_spec_stats + 7, dl.d_chain, 8, cudaMemcpyDeviceToDevice, ctx->stream)); // diagnostics: segment 0
	}
	uint32_t max_n = 0;
	for (uint32_t sg = 0; sg < nseg; ++sg) max_n = std::max(max_n, m->h_obase[sg + 1] - m->h_obase[sg]);
	for (size_t gi = 0; gi < dl.groups.size(); ++gi) {
		SpecGroup &sgp = dl.groups[gi];
		const int st = sgp.st, esize = hb_type_size(st), nc = sgp.ncomp, ncp = nc == 3 ? 4 : nc;
		CompMap cm;
		cm.n = nc;
		for (int j = 0; j < 4; ++j) cm.comp[j] = (int8_t)(j < nc ? sgp.comp[j] : 0);
		HB_TRY(hb_dalloc_t(m, &sgp.d_cres, (size_t)(n + 1) * ncp * esize));
		HB_TRY(hb_dalloc_t(m, &sgp.d_cx, (size_t)(n + 1) * ncp * esize));
		HB_LAUNCH(ctx, k_gather_compact, g, 256, 0, p, cm, dl.d_erow, n, sgp.d_cres, esize, ncp);
		HB_CUDA(ctx, cudaMemsetAsync(sgp.d_cx, 0, (size_t)(n + 1) * ncp * esize, ctx->stream));
		// one argument block per segment (mesh of a batch): the kernels run one CTA / cluster set per block
		m->h_spec_args.resize(nseg);
		for (uint32_t sg = 0; sg < nseg; ++sg) {
			SpecArgs &a = m->h_spec_args[sg];
			a.srec = srec;
			a.kind = dl.d_kind; a.src = dl.d_src; a.cand_off = m->d_vc_off; a.cand = m->d_vc_tri;
			a.resid = sgp.d_cres; a.x = sgp.d_cx;
			a.base = m->h_obase[sg]; a.n = m->h_obase[sg + 1];
			a.stats = sg == 0 && gi == 0 ? dl.d_spec_stats : nullptr;
			for (int j = 0; j < 4; ++j) a.bits[j] = j < nc ? (p.quant[sgp.comp[j]] ? p.quant[sgp.comp[j]] : 8 * esize) : 8 * esize;
		}
		HB_TRY(hb_dalloc_t(m, &sgp.d_args, nseg));
		// pageable source: the runtime stages it before cudaMemcpyAsync returns, no synchronisation needed
		HB_CUDA(ctx, cudaMemcpyAsync(sgp.d_args, m->h_spec_args.data(), sizeof(SpecArgs) * nseg, cudaMemcpyHostToDevice, ctx->stream));
		static const bool single_cta = getenv("HARRY_B200_SPEC1") != nullptr; // A/B switch: single-CTA kernel
		if (st != HB_FLOAT && !single_cta) {
			const int ntb = scan_ntb(nseg * (uint32_t)nc, ctx->sm_count);
			const int cl = scan_cluster_size(max_n, nseg, ntb);
			if (st == HB_UCHAR) HB_TRY(launch_scan<uint8_t>(ctx, nc, sgp.d_args, cl, nseg, ntb));
			else if (st == HB_USHORT) HB_TRY(launch_scan<uint16_t>(ctx, nc, sgp.d_args, cl, nseg, ntb));
			else HB_TRY(launch_scan<uint32_t>(ctx, nc, sgp.d_args, cl, nseg, ntb));
		} else if (st == HB_UCHAR) HB_TRY((launch_spec<uint8_t, false>(ctx, nc, sgp.d_args, nseg)));
		else if (st == HB_USHORT) HB_TRY((launch_spec<uint16_t, false>(ctx, nc, sgp.d_args, nseg)));
		else if (st == HB_UINT) HB_TRY((launch_spec<uint32_t, false>(ctx, nc, sgp.d_args, nseg)));
		else {
			HB_TRY(launch_walkf(ctx, nc, sgp.d_args, nseg, dl.d_chain));
			HB_TRY((launch_spec<uint32_t, true>(ctx, nc, sgp.d_args, nseg, dl.d_chain)));
		}
		HB_LAUNCH(ctx, k_scatter_compact, g, 256, 0, p, cm, dl.d_erow, dl.d_kind, n, sgp.d_cx, esize, ncp);
	}
	return 0;
}

// rank-space records -> AoS rows (DATA elements own their row)
__global__ void __launch_bounds__(256) k_scatter_rp(ListParams p, const uint32_t *__restrict__ erow, const uint32_t *__restrict__ first, uint32_t n, const unsigned long long *__restrict__ rp,
                                                     uint32_t comp_mask /* components reconstructed in rp (the others went through compact records) */)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint32_t row = erow[i];
	if (row == HB_NONE || first[row] != i) return;
	uint8_t *dst = p.rows + (size_t)row * p.stride;
	for (int j = 0; j < p.ncomp; ++j)
		if ((comp_mask >> j) & 1u) hb_st_bits(dst + p.offset[j], p.size[j], rp[(size_t)i * p.ncomp + j]);
}

// ------------------------------------------------------------------------------------------------
// FACE: prediction 0 (Appendix C.2) -> decodeDelta(delta, 0): identity for integer storage types,
// un-flip for floats (prediction.h:64-72)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_decode_face(ListParams p, const uint32_t *__restrict__ erow, const uint32_t *__restrict__ first, uint32_t n)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint32_t row = erow[i];
	if (row == HB_NONE || first[row] != i) return;
	uint8_t *dst = p.rows + (size_t)row * p.stride;
	for (int j = 0; j < p.ncomp; ++j) {
		const unsigned long long d = hb_ld_bits(dst + p.offset[j], p.size[j]);
		hb_st_bits(dst + p.offset[j], p.size[j], hb_dec(p.stype[j], d, 0, p.quant[j]));
	}
}

// ------------------------------------------------------------------------------------------------
// CORNER: level-synchronous wavefront sweeps
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long ld_cg_u64(const unsigned long long *p) { return __ldcg(p); }

// One sweep: every not-yet-done DATA element whose candidate rows are all available is
// reconstructed.  A candidate row that is emitted LATER than this element (or never) reads as
// zero, exactly like the reference's zero-initialised rows (see harry_b200.h, emit_type).
__global__ void __launch_bounds__(256) k_decode_corner_sweep(ListParams p, const uint32_t *__restrict__ erow, const uint32_t *__restrict__ first,
                                                              const uint32_t *__restrict__ cand_off, const uint32_t *__restrict__ cand, uint32_t n,
                                                              unsigned long long *rp, volatile uint8_t *done, uint32_t *__restrict__ remaining,
                                                              const uint32_t *__restrict__ remaining_before /* nullptr: first sweep */)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	// sweeps are launched several at a time without asking the host in between: once nothing is left they return at once
	if (remaining_before && *remaining_before == 0u) return;
	const uint32_t row = erow[i];
	if (row == HB_NONE || first[row] != i || done[i]) return;
	const uint32_t c0 = cand_off[i], K = cand_off[i + 1] - c0;
	bool ready = true;
	for (uint32_t k = 0; k < K && ready; ++k) {
		const uint32_t o = first[erow[cand[c0 + k]]];
		if (o < i) ready = done[o] != 0; // o == HB_NONE or o > i: not emitted yet -> zeros
	}
	if (!ready) { atomicAdd(remaining, 1u); return; }
	__threadfence();
	for (int j = 0; j < p.ncomp; ++j) {
		const int st = p.stype[j];
		const unsigned long long pred = combine_candidates(st, K, [&](uint32_t kk) -> unsigned long long {
			const uint32_t o = first[erow[cand[c0 + kk]]];
			return o < i ? ld_cg_u64(rp + (size_t)o * p.ncomp + j) : 0ull;
		});
		rp[(size_t)i * p.ncomp + j] = hb_dec(st, ld_cg_u64(rp + (size_t)i * p.ncomp + j), pred, p.quant[j]);
	}
	__threadfence();
	done[i] = 1;
}

// ------------------------------------------------------------------------------------------------
int hb_decode_lists(hb_dmesh *m)
{
	hb_ctx *ctx = m->ctx;
	HB_TRY(hb_build_conn(m));
	bool need_v = false, need_c = false;
	for (int l = 0; l < m->nlists; ++l) {
		const int cls = m->lists[l].p.target;
		if (cls == CLS_VTX && m->lists[l].p.ncomp) need_v = true;
		if (cls == CLS_CORNER && m->lists[l].p.ncomp) need_c = true;
	}
	if (need_v) HB_TRY(hb_build_vertex_candidates(m));
	if (need_c && m->any_corner) HB_TRY(hb_build_corner_candidates(m));

	std::vector<WalkArgs> walks;
	std::vector<int> walk_lists;
	std::vector<uint32_t> walk_masks;
	for (int l = 0; l < m->nlists; ++l) {
		DevList &dl = m->lists[l];
		const ListParams &p = dl.p;
		const int cls = p.target;
		if (p.ncomp == 0 || (cls != CLS_VTX && cls != CLS_FACE && cls != CLS_CORNER)) continue;
		if (cls == CLS_CORNER && !m->any_corner) continue;
		std::vector<int> rest;
		if (cls == CLS_VTX) plan_groups(p, dl.groups, rest);
		const bool all_grouped = cls == CLS_VTX && rest.empty();
		HB_TRY(hb_prepare_list_elems(m, l, cls != CLS_FACE && !all_grouped, true));
		const uint32_t n = dl.n_elems;
		if (!n) continue;
		if (cls == CLS_VTX && !dl.groups.empty()) HB_TRY(decode_vertex_groups(m, l));
		if (all_grouped) {
		} else if (cls == CLS_FACE) {
			HB_LAUNCH(ctx, k_decode_face, hb_div_up(n, 256), 256, 0, p, dl.d_erow, dl.d_first, n);
		} else if (cls == CLS_VTX) {
			WalkArgs w;
			w.ncomp = p.ncomp;
			w.erow = dl.d_erow; w.first = dl.d_first; w.cand_off = m->d_vc_off; w.cand = m->d_vc_tri; w.rp = dl.d_rp;
			for (uint32_t sg = 0; sg < m->nseg; ++sg) { // one chain per segment (mesh of a batch) and component
				w.base = m->h_obase[sg]; w.n = m->h_obase[sg + 1];
				for (int j : rest) { // the components no compact group takes (signed and 64-bit storage types)
					w.comp = j; w.stype = p.stype[j]; w.quant = p.quant[j];
					walks.push_back(w);
				}
			}
			uint32_t mask = 0;
			for (int j : rest) mask |= 1u << j;
			walk_lists.push_back(l);
			walk_masks.push_back(mask);
		} else {
			HB_TRY(hb_dalloc_t(m, &dl.d_done, (size_t)n + 1));   // kept with the list: reused by the next decode of this mesh
			// the DAG is shallow (a handful of levels): CORNER_SWEEPS sweeps go out back to back, each with its own counter
			// of the elements it had to leave; the host looks at the last counter of the round only
			constexpr uint32_t CORNER_SWEEPS = 4;
			HB_TRY(hb_dalloc_t(m, &dl.d_remaining, CORNER_SWEEPS));
			uint8_t *done = dl.d_done;
			uint32_t *remaining = dl.d_remaining;
			HB_CUDA(ctx, cudaMemsetAsync(done, 0, (size_t)n + 1, ctx->stream));
			uint32_t prev = 0xffffffffu;
			for (uint32_t round = 0;; ++round) {
				HB_CUDA(ctx, cudaMemsetAsync(remaining, 0, sizeof(uint32_t) * CORNER_SWEEPS, ctx->stream));
				for (uint32_t k = 0; k < CORNER_SWEEPS; ++k)
					HB_LAUNCH(ctx, k_decode_corner_sweep, hb_div_up(n, 256), 256, 0, p, dl.d_erow, dl.d_first, m->d_cc_off, m->d_cc_idx, n, dl.d_rp, done, remaining + k,
					          k ? remaining + k - 1 : (const uint32_t *)nullptr);
				uint32_t rem = 0;
				HB_CUDA(ctx, cudaMemcpyAsync(&rem, remaining + CORNER_SWEEPS - 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
				HB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
				if (rem == 0) break; // (a sweep that found nothing left wrote nothing: the counters behind it stay 0)
				if (rem >= prev) return hb_fail(ctx, HB_ERR_INVALID, "corner decode: dependency cycle (%u elements stuck)", rem);
				prev = rem;
			}
			HB_LAUNCH(ctx, k_scatter_rp, hb_div_up(n, 256), 256, 0, p, dl.d_erow, dl.d_first, n, dl.d_rp, 0xffffffffu);
		}
	}
	if (!walks.empty()) {
		HB_TRY(hb_dalloc(m, &m->d_walks, sizeof(WalkArgs) * walks.size()));
		WalkArgs *d_walks = (WalkArgs *)m->d_walks;
		HB_CUDA(ctx, cudaMemcpyAsync(d_walks, walks.data(), sizeof(WalkArgs) * walks.size(), cudaMemcpyHostToDevice, ctx->stream));
		HB_LAUNCH(ctx, k_decode_vertex_walk, (uint32_t)walks.size(), 32, 0, d_walks); // (the pageable source was staged by the copy call)
		for (size_t k = 0; k < walk_lists.size(); ++k) {
			DevList &dl = m->lists[walk_lists[k]];
			HB_LAUNCH(ctx, k_scatter_rp, hb_div_up(dl.n_elems, 256), 256, 0, dl.p, dl.d_erow, dl.d_first, dl.n_elems, dl.d_rp, walk_masks[k]);
		}
	}
	return 0;
}
