// hb_quant.cu -- quantization kernels:
//   K1 bounds_reduce   per-component min / max over all rows       (structs/quant.h:30-38)
//      scale           per interpretation-group scale (tiny)       (structs/quant.h:46-96)
//   K2/K8 requant      float/int <-> fixed point, in place, AoS    (structs/quant.h:98-221)
//
// Mapping: one thread per (row, component) scalar, flattened as e = row * ncomp + comp, so that
// for the usual all-float rows consecutive lanes touch consecutive 4-byte words (fully coalesced
// 128-byte requests) even though the layout is AoS.  The thread count is a multiple of ncomp, so
// each thread keeps one component for the whole grid-stride loop.
//
// Every kernel is segmented: blockIdx.y = segment (mesh of a batch, hb_internal.cuh), the rows of a
// segment are [rowbase, rowbase + rownum) of the concatenated list, its bounds rows live at
// bounds + segment * pitch.  A single mesh is one segment.
#include "hb_internal.cuh"

#include <float.h>
#include <string.h>
#include <stdlib.h>

// monotone u64 sort keys: unsigned order of the key == numeric order of the value
__device__ __forceinline__ unsigned long long key_of(unsigned long long bits, int type)
{
	switch (type) {
	case HB_FLOAT: { const uint32_t b = (uint32_t)bits; return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u); }
	case HB_DOUBLE: return bits ^ ((bits >> 63) ? ~0ull : 0x8000000000000000ull);
	case HB_LONG: return bits ^ 0x8000000000000000ull;
	case HB_INT: return (unsigned long long)((long long)(int32_t)bits) ^ 0x8000000000000000ull;
	case HB_SHORT: return (unsigned long long)((long long)(int16_t)bits) ^ 0x8000000000000000ull;
	case HB_CHAR: return (unsigned long long)((long long)(int8_t)bits) ^ 0x8000000000000000ull;
	default: return bits; // unsigned types
	}
}
__device__ __forceinline__ unsigned long long value_of(unsigned long long key, int type)
{
	switch (type) {
	case HB_FLOAT: { const uint32_t k = (uint32_t)key; return k ^ ((k >> 31) ? 0x80000000u : 0xffffffffu); }
	case HB_DOUBLE: return key ^ ((key >> 63) ? 0x8000000000000000ull : ~0ull);
	case HB_LONG: case HB_INT: case HB_SHORT: case HB_CHAR: return key ^ 0x8000000000000000ull;
	default: return key;
	}
}
// numeric_limits<T>::max() / ::min() as bit patterns (quant.h:32-33; min() of a floating type is
// the smallest positive normal -- Appendix C.1)
__device__ __forceinline__ unsigned long long seed_bits(int type, bool for_min)
{
	switch (type) {
	case HB_FLOAT: return for_min ? 0x7f7fffffu : 0x00800000u;
	case HB_DOUBLE: return for_min ? 0x7fefffffffffffffull : 0x0010000000000000ull;
	case HB_ULONG: return for_min ? ~0ull : 0ull;
	case HB_LONG: return for_min ? 0x7fffffffffffffffull : 0x8000000000000000ull;
	case HB_UINT: return for_min ? 0xffffffffull : 0ull;
	case HB_INT: return for_min ? 0x7fffffffull : 0x80000000ull;
	case HB_USHORT: return for_min ? 0xffffull : 0ull;
	case HB_SHORT: return for_min ? 0x7fffull : 0x8000ull;
	case HB_UCHAR: return for_min ? 0xffull : 0ull;
	default: return for_min ? 0x7full : 0x80ull;
	}
}


// Reduction scratch per (segment, component): four u64 words, ALL combined with atomicMax so that a zero-filled
// buffer is the identity: [0] ~(smallest key), [1] largest key, [2] ~(first row holding -0.0), [3] ~(first row
// holding +0.0).  Behind the words of a segment: one u32 ticket counter -- the last block of a segment to finish
// writes the bounds rows (and the scale row when the interpretation groups are known): no init / finish kernels.
struct BoundsSeg {
	const uint32_t *rowbase, *rownum;
	unsigned long long *scratch;   // [nseg][4 * ncomp]
	uint32_t *ticket;              // [nseg]
	uint32_t *partial;             // flat float kernel: [nseg][blocks][4 * ncomp] per-block results
	uint8_t *bounds;               // [nseg] triples of rows, `pitch` bytes apart
	size_t pitch;
	uint8_t groups[HB_MAX_COMP];   // interpretation-group leader per component (quant.h:54-91)
	int have_groups;
	int *err;
};

// quant.h:46-96 for one segment (groups whose members share one type)
__device__ void scale_segment(const ListParams &p, uint8_t *__restrict__ bounds, const uint8_t *groups, int *err)
{
	const uint8_t *mnr = bounds, *mxr = bounds + p.stride;
	uint8_t *scr = bounds + 2 * (size_t)p.stride;
	for (int k = 0; k < p.ncomp; ++k) {
		if (groups[k] != k) continue;
		const int t = p.type[k];
		const int size = hb_type_size(t);
		unsigned long long s_bits;
		if (t == HB_FLOAT) {
			float s = FLT_MIN;
			for (int j = 0; j < p.ncomp; ++j) {
				if (groups[j] != k) continue;
				if (p.type[j] != t) { atomicExch(err, 6); continue; }
				const float r = __fsub_rn(__uint_as_float((uint32_t)hb_ld_bits(mxr + p.offset[j], 4)), __uint_as_float((uint32_t)hb_ld_bits(mnr + p.offset[j], 4)));
				s = s < r ? r : s;
			}
			s_bits = __float_as_uint(s);
		} else if (t == HB_DOUBLE) {
			double s = DBL_MIN;
			for (int j = 0; j < p.ncomp; ++j) {
				if (groups[j] != k) continue;
				if (p.type[j] != t) { atomicExch(err, 6); continue; }
				const double r = __dsub_rn(__longlong_as_double((long long)hb_ld_bits(mxr + p.offset[j], 8)), __longlong_as_double((long long)hb_ld_bits(mnr + p.offset[j], 8)));
				s = s < r ? r : s;
			}
			s_bits = (unsigned long long)__double_as_longlong(s);
		} else {
			const bool sg = t == HB_LONG || t == HB_INT || t == HB_SHORT || t == HB_CHAR;
			long long s_i = 0;
			unsigned long long s_u = 0;
			if (sg) s_i = hb_bits_to_i64(seed_bits(t, false), t);
			for (int j = 0; j < p.ncomp; ++j) {
				if (groups[j] != k) continue;
				if (p.type[j] != t) { atomicExch(err, 6); continue; }
				unsigned long long r = hb_ld_bits(mxr + p.offset[j], size) - hb_ld_bits(mnr + p.offset[j], size);
				if (size < 8) r &= (1ull << (8 * size)) - 1ull;
				if (sg) {
					const long long ri = hb_bits_to_i64(r, t);
					s_i = s_i < ri ? ri : s_i;
				} else {
					s_u = s_u < r ? r : s_u;
				}
			}
			s_bits = sg ? (unsigned long long)s_i : s_u;
		}
		for (int j = 0; j < p.ncomp; ++j)
			if (groups[j] == k) hb_st_bits(scr + p.offset[j], size, s_bits);
	}
}

// the block's partial results are in the scratch words: the last block of the segment turns them into rows
// [0] min, [1] max (, [2] scale) -- `stride` bytes each, components at their row offsets
// true in the block that finishes last in its segment (every block's results are visible to it)
__device__ bool bounds_is_last(const BoundsSeg &b, uint32_t seg, uint32_t blocks_per_seg)
{
	__shared__ uint32_t s_last;
	__syncthreads();
	if (threadIdx.x == 0) {
		__threadfence();
		s_last = atomicAdd(&b.ticket[seg], 1u) == blocks_per_seg - 1 ? 1u : 0u;
	}
	__syncthreads();
	if (s_last) __threadfence();
	return s_last != 0;
}
// sc: the combined words of the segment (global scratch, or shared memory of the last block)
__device__ void bounds_finish(const ListParams &p, const BoundsSeg &b, uint32_t seg, const volatile unsigned long long *sc)
{
	uint8_t *bounds = b.bounds + (size_t)seg * b.pitch;
	if ((int)threadIdx.x < p.ncomp) {
		const int j = threadIdx.x;
		const int type = p.type[j];
		const int size = hb_type_size(type);
		unsigned long long mn = value_of(~sc[4 * j + 0], type);
		const unsigned long long mx = value_of(sc[4 * j + 1], type);
		if (type == HB_FLOAT || type == HB_DOUBLE) {
			// minimum is a zero: its sign is that of the first zero in row order (`e < cur` is false
			// between -0.0 and +0.0, so a later zero never replaces an earlier one)
			const unsigned long long mag = type == HB_FLOAT ? (mn & 0x7fffffffu) : (mn & 0x7fffffffffffffffull);
			if (mag == 0) {
				const bool neg_first = ~sc[4 * j + 2] < ~sc[4 * j + 3];
				mn = neg_first ? (type == HB_FLOAT ? 0x80000000ull : 0x8000000000000000ull) : 0ull;
			}
		}
		hb_st_bits(bounds + p.offset[j], size, mn);
		hb_st_bits(bounds + p.stride + p.offset[j], size, mx);
	}
	__syncthreads();
	if (threadIdx.x == 0 && b.have_groups) scale_segment(p, bounds, b.groups, b.err);
}

// generic runtime-typed reduction: one (row, component) scalar per thread and step
__global__ void __launch_bounds__(256) k_bounds_reduce(ListParams p, BoundsSeg b, uint32_t rows_per_step)
{
	const uint32_t seg = blockIdx.y;
	const uint32_t r0 = b.rowbase[seg], nrows = b.rownum[seg];
	const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t total_threads = rows_per_step * (uint32_t)p.ncomp;
	const int j = (int)(tid % (uint32_t)p.ncomp);
	const int type = p.type[j];
	const int size = hb_type_size(type);
	const bool is_fp = type == HB_FLOAT || type == HB_DOUBLE;
	unsigned long long kmin = key_of(seed_bits(type, true), type), kmax = key_of(seed_bits(type, false), type);
	unsigned long long zneg = ~0ull, zpos = ~0ull;
	const uint8_t *base = p.rows + (size_t)r0 * p.stride + p.offset[j];
	for (uint32_t row = tid < total_threads ? tid / (uint32_t)p.ncomp : nrows; row < nrows; row += rows_per_step) {
		const unsigned long long bits = hb_ld_bits(base + (size_t)row * p.stride, size);
		if (is_fp) {
			// NaN never replaces a bound (both comparisons are false)
			const bool nan = type == HB_FLOAT ? ((bits & 0x7fffffffu) > 0x7f800000u) : ((bits & 0x7fffffffffffffffull) > 0x7ff0000000000000ull);
			if (nan) continue;
			// -0.0 == +0.0 for the reference's `<`: remember which zero came first in row order
			const unsigned long long mag = type == HB_FLOAT ? (bits & 0x7fffffffu) : (bits & 0x7fffffffffffffffull);
			if (mag == 0) {
				const bool neg = type == HB_FLOAT ? (bits >> 31) != 0 : (bits >> 63) != 0;
				if (neg) zneg = zneg < row ? zneg : row;
				else zpos = zpos < row ? zpos : row;
			}
		}
		const unsigned long long k = key_of(bits, type);
		kmin = k < kmin ? k : kmin;
		kmax = k > kmax ? k : kmax;
	}
	// block-level combine in shared memory (one slot per component), then one global atomic per
	// block and component instead of one per thread
	__shared__ unsigned long long s_red[4 * HB_MAX_COMP];
	for (int k = threadIdx.x; k < 4 * p.ncomp; k += blockDim.x) s_red[k] = 0ull;
	__syncthreads();
	atomicMax(&s_red[4 * j + 0], ~kmin);
	atomicMax(&s_red[4 * j + 1], kmax);
	if (zneg != ~0ull) atomicMax(&s_red[4 * j + 2], ~zneg);
	if (zpos != ~0ull) atomicMax(&s_red[4 * j + 3], ~zpos);
	__syncthreads();
	unsigned long long *sc = b.scratch + (size_t)seg * 4 * p.ncomp;
	for (int k = threadIdx.x; k < 4 * p.ncomp; k += blockDim.x)
		if (s_red[k]) atomicMax(&sc[k], s_red[k]);
	if (bounds_is_last(b, seg, gridDim.x)) bounds_finish(p, b, seg, b.scratch + (size_t)seg * 4 * p.ncomp);
}

// quant.h:98-112, integer flavour, evaluated in T (narrow types compute in int and truncate)
__device__ __forceinline__ unsigned long long rescale_int(int t, unsigned long long val, unsigned long long from, unsigned long long to)
{
	switch (t) {
	case HB_ULONG: return from ? val / from * to + val % from * to / from : 0;
	case HB_LONG: { const long long v = (long long)val, f = (long long)from, o = (long long)to; return f ? (unsigned long long)(v / f * o + v % f * o / f) : 0; }
	case HB_UINT: { const uint32_t v = (uint32_t)val, f = (uint32_t)from, o = (uint32_t)to; return f ? (uint32_t)(v / f * o + v % f * o / f) : 0; }
	case HB_INT: { const int32_t v = (int32_t)val, f = (int32_t)from, o = (int32_t)to; return f ? (uint32_t)((uint32_t)(v / f) * (uint32_t)o + (uint32_t)((int32_t)((uint32_t)(v % f) * (uint32_t)o) / f)) : 0; }
	case HB_USHORT: { const int v = (uint16_t)val, f = (uint16_t)from, o = (uint16_t)to; return f ? (uint16_t)(v / f * o + v % f * o / f) : 0; }
	case HB_SHORT: { const int v = (int16_t)val, f = (int16_t)from, o = (int16_t)to; return f ? (uint16_t)(int16_t)(v / f * o + v % f * o / f) : 0; }
	case HB_UCHAR: { const int v = (uint8_t)val, f = (uint8_t)from, o = (uint8_t)to; return f ? (uint8_t)(v / f * o + v % f * o / f) : 0; }
	default: { const int v = (int8_t)val, f = (int8_t)from, o = (int8_t)to; return f ? (uint8_t)(int8_t)(v / f * o + v % f * o / f) : 0; }
	}
}

struct RequantParams {
	uint8_t dq[HB_MAX_COMP]; // destination quantization bits per component
};
struct RequantSeg {
	const uint32_t *rowbase, *rownum;
	const uint8_t *bounds;
	size_t pitch;
};

// quant.h:114-214 for every (row, component); four separately rounded float operations
// (no FMA contraction), truncating float -> integer conversion.
__global__ void __launch_bounds__(256) k_requant(ListParams p, RequantParams rq, RequantSeg sg_, uint32_t rows_per_step)
{
	const uint32_t seg = blockIdx.y;
	const uint32_t r0 = sg_.rowbase[seg], nrows = sg_.rownum[seg];
	const uint8_t *bounds = sg_.bounds + (size_t)seg * sg_.pitch;
	const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t total_threads = rows_per_step * (uint32_t)p.ncomp;
	if (tid >= total_threads) return;
	const int j = (int)(tid % (uint32_t)p.ncomp);
	const int t = p.type[j];
	const int sq = p.quant[j], dq = rq.dq[j];
	if (sq == 0 && dq == 0) return; // quant.h:118-121
	const int tsize = hb_type_size(t);
	const int st_src = sq ? (sq <= 8 ? HB_UCHAR : sq <= 16 ? HB_USHORT : sq <= 32 ? HB_UINT : HB_ULONG) : t;
	const int st_dst = dq ? (dq <= 8 ? HB_UCHAR : dq <= 16 ? HB_USHORT : dq <= 32 ? HB_UINT : HB_ULONG) : t;
	const int ssize = hb_type_size(st_src), dsize = hb_type_size(st_dst);
	const unsigned long long m_src = sq ? (unsigned long long)(long long)(int32_t)((1u << sq) - 1u) : 0;
	const unsigned long long m_dst = dq ? (unsigned long long)(long long)(int32_t)((1u << dq) - 1u) : 0;
	const unsigned long long mn_b = hb_ld_bits(bounds + p.offset[j], tsize);
	const unsigned long long sc_b = hb_ld_bits(bounds + 2 * (size_t)p.stride + p.offset[j], tsize);
	const float mn_f = __uint_as_float((uint32_t)mn_b), sc_f = __uint_as_float((uint32_t)sc_b);
	const float m_dst_f = (float)(int32_t)m_dst, m_src_f = (float)(int32_t)m_src;
	const bool sg = t == HB_LONG || t == HB_INT || t == HB_SHORT || t == HB_CHAR;
	uint8_t *base = p.rows + (size_t)r0 * p.stride + p.offset[j];
	for (uint32_t row = tid / (uint32_t)p.ncomp; row < nrows; row += rows_per_step) {
		uint8_t *ptr = base + (size_t)row * p.stride;
		unsigned long long q;
		if (sq) {
			q = hb_ld_bits(ptr, ssize);
		} else if (t == HB_FLOAT) {
			const float x = __uint_as_float((uint32_t)hb_ld_bits(ptr, 4));
			const float d = __fadd_rn(__fmul_rn(__fdiv_rn(__fsub_rn(x, mn_f), sc_f), m_dst_f), 0.5f);
			q = (unsigned long long)d; // quant.h:135
		} else {
			const unsigned long long v = hb_ld_bits(ptr, tsize) - mn_b;
			q = rescale_int(t, v, sc_b, m_dst);
			if (sg) q = (unsigned long long)hb_bits_to_i64(q, t);
		}
		if (sq && dq) q = q / m_src * m_dst + q % m_src * m_dst / m_src; // quant.h:167-169
		if (dq) {
			hb_st_bits(ptr, dsize, q); // the other bytes of the slot keep their old contents
		} else if (t == HB_FLOAT) {
			const float x = __fadd_rn(__fmul_rn(__fdiv_rn((float)q, m_src_f), sc_f), mn_f); // quant.h:182
			hb_st_bits(ptr, 4, __float_as_uint(x));
		} else {
			hb_st_bits(ptr, tsize, rescale_int(t, q, m_src, sc_b) + mn_b);
		}
	}
}

// ------------------------------------------------------------------------------------------------
// Fast paths for the usual list: every component float32, rows tightly packed (stride == 4 * ncomp),
// unquantized or at most 31 bits.  The rows of a segment are then one flat float array that starts on a
// 16-byte boundary: 128-bit loads and stores, FOUR independent vectors in flight per thread and step;
// the threads of a segment are a multiple of ncomp, so the component of each of a thread's four lanes
// never changes.  Same arithmetic as the generic kernels.
// ------------------------------------------------------------------------------------------------
#ifndef FLAT_UNROLL
#define FLAT_UNROLL 4
#endif
static bool flat_f32(const ListParams &p)
{
	if (p.ncomp < 1 || p.ncomp > 4 || p.stride != 4u * (uint32_t)p.ncomp || (((size_t)p.rows) & 15u)) return false;
	for (int j = 0; j < p.ncomp; ++j)
		if (p.type[j] != HB_FLOAT || p.offset[j] != 4 * j) return false;
	return true;
}

template <int NC>
__global__ void __launch_bounds__(256) k_bounds_reduce_f32(ListParams p, BoundsSeg b)
{
	constexpr uint32_t ncomp = NC;
	const uint32_t seg = blockIdx.y;
	const uint32_t nscal = b.rownum[seg] * ncomp;
	const uint4 *__restrict__ rows4 = (const uint4 *)(p.rows + (size_t)b.rowbase[seg] * p.stride);
	const uint32_t *__restrict__ rows1 = (const uint32_t *)rows4;
	const uint32_t G = gridDim.x * blockDim.x; // a multiple of ncomp
	const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t nvec = nscal >> 2;
	// Running bounds as floats: fminf / fmaxf return the other operand when one is a NaN (a NaN never replaces a bound,
	// quant.h:34-35) and cost one instruction each.  They order -0 below +0 where the reference's `<` sees a tie, so a
	// ZERO minimum gets its sign from the first zero in row order (zneg / zpos; a zero maximum cannot occur: the seed
	// numeric_limits<float>::min() is positive and only a larger value replaces it).
	float fmin_[4], fmax_[4];
	uint32_t zneg[4], zpos[4];
#pragma unroll
	for (int m = 0; m < 4; ++m) {
		fmin_[m] = FLT_MAX;   // numeric_limits<float>::max()
		fmax_[m] = FLT_MIN;   // numeric_limits<float>::min() (quant.h:33)
		zneg[m] = zpos[m] = 0xffffffffu;
	}
	auto take = [&](int m, uint32_t bits) {
		const float x = __uint_as_float(bits);
		fmin_[m] = fminf(fmin_[m], x);
		fmax_[m] = fmaxf(fmax_[m], x);
	};
	auto zeros = [&](const uint4 &q, uint32_t e) {
		const uint32_t w[4] = { q.x, q.y, q.z, q.w };
#pragma unroll
		for (int m = 0; m < 4; ++m)
			if ((w[m] << 1) == 0u) {
				const uint32_t row = (e + (uint32_t)m) / ncomp;
				if (w[m] >> 31) zneg[m] = min(zneg[m], row);
				else zpos[m] = min(zpos[m], row);
			}
	};
	for (uint32_t v = tid; v < nvec; v += FLAT_UNROLL * G) {
		uint4 q[FLAT_UNROLL];
#pragma unroll
		for (int u = 0; u < FLAT_UNROLL; ++u) {
			const uint32_t vv = v + (uint32_t)u * G;
			q[u] = vv < nvec ? __ldcs(rows4 + vv) : make_uint4(0x7fc00000u, 0x7fc00000u, 0x7fc00000u, 0x7fc00000u); // NaN: ignored
		}
#pragma unroll
		for (int u = 0; u < FLAT_UNROLL; ++u) {
			take(0, q[u].x); take(1, q[u].y); take(2, q[u].z); take(3, q[u].w);
			if (min(min(q[u].x << 1, q[u].y << 1), min(q[u].z << 1, q[u].w << 1)) == 0u) zeros(q[u], 4u * (v + (uint32_t)u * G));
		}
	}
	// order-preserving integer keys from here on (the reductions below compare unsigned)
	uint32_t kmin[4], kmax[4];
#pragma unroll
	for (int m = 0; m < 4; ++m) {
		const uint32_t bn = __float_as_uint(fmin_[m]) & (fmin_[m] == 0.f ? 0x7fffffffu : 0xffffffffu); // a zero minimum as +0: its sign is decided by zneg / zpos
		const uint32_t bx = __float_as_uint(fmax_[m]);
		kmin[m] = bn ^ ((uint32_t)((int32_t)bn >> 31) | 0x80000000u);
		kmax[m] = bx ^ ((uint32_t)((int32_t)bx >> 31) | 0x80000000u);
	}
	// per-component accumulators of this thread, then a shuffle reduction over the warp, one shared-memory
	// slot per warp, and one set of global atomics per block
	uint32_t amin[NC], amax[NC], azn[NC], azp[NC];
#pragma unroll
	for (int c = 0; c < NC; ++c) { amin[c] = 0xffffffffu; amax[c] = 0u; azn[c] = azp[c] = 0xffffffffu; }
#pragma unroll
	for (int m = 0; m < 4; ++m) {
		const uint32_t j = (4u * tid + (uint32_t)m) % ncomp;
#pragma unroll
		for (int c = 0; c < NC; ++c)
			if (j == (uint32_t)c) { amin[c] = min(amin[c], kmin[m]); amax[c] = max(amax[c], kmax[m]); azn[c] = min(azn[c], zneg[m]); azp[c] = min(azp[c], zpos[m]); }
	}
	// the up to three scalars behind the last full vector
	if (tid == 0) {
		for (uint32_t e = nvec << 2; e < nscal; ++e) {
			const uint32_t bits = rows1[e], j = e % ncomp, row = e / ncomp;
			if ((bits & 0x7fffffffu) > 0x7f800000u) continue;
			const uint32_t k = bits ^ ((bits >> 31) ? 0xffffffffu : 0x80000000u);
#pragma unroll
			for (int c = 0; c < NC; ++c)
				if (j == (uint32_t)c) {
					if ((bits & 0x7fffffffu) == 0u) { if (bits >> 31) azn[c] = min(azn[c], row); else azp[c] = min(azp[c], row); }
					amin[c] = min(amin[c], k); amax[c] = max(amax[c], k);
				}
		}
	}
#pragma unroll
	for (int c = 0; c < NC; ++c) {
		amin[c] = __reduce_min_sync(0xffffffffu, amin[c]);
		amax[c] = __reduce_max_sync(0xffffffffu, amax[c]);
		azn[c] = __reduce_min_sync(0xffffffffu, azn[c]);
		azp[c] = __reduce_min_sync(0xffffffffu, azp[c]);
	}
	__shared__ uint32_t s_w[8][4 * NC];
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (lane == 0) {
#pragma unroll
		for (int c = 0; c < NC; ++c) { s_w[warp][4 * c] = amin[c]; s_w[warp][4 * c + 1] = amax[c]; s_w[warp][4 * c + 2] = azn[c]; s_w[warp][4 * c + 3] = azp[c]; }
	}
	__syncthreads();
	// one slot of partial results per block -- a thousand blocks hammering a dozen addresses with atomics cost more than
	// the whole streaming loop of a 120 MB list; the block that finishes last folds the slots
	uint32_t *part = b.partial + ((size_t)seg * gridDim.x + blockIdx.x) * 4 * NC;
	if (threadIdx.x < 4 * NC) {
		const uint32_t k = threadIdx.x;
		uint32_t v = s_w[0][k];
		for (int w = 1; w < 8; ++w) v = (k & 3) == 1 ? max(v, s_w[w][k]) : min(v, s_w[w][k]);
		part[k] = v;
	}
	if (!bounds_is_last(b, seg, gridDim.x)) return;
	{
		constexpr uint32_t W = 4 * NC, GROUPS = 256 / W;
		__shared__ uint32_t s_f[GROUPS][W];
		__shared__ unsigned long long s_fin[W];
		const uint32_t k = threadIdx.x % W, g = threadIdx.x / W;
		if (g < GROUPS) {
			uint32_t v = (k & 3) == 1 ? 0u : 0xffffffffu;
			const uint32_t *all = b.partial + (size_t)seg * gridDim.x * W;
#pragma unroll 4
			for (uint32_t blk = g; blk < gridDim.x; blk += GROUPS) {
				const uint32_t x = __ldcg(all + (size_t)blk * W + k); // written by other SMs: L2
				v = (k & 3) == 1 ? max(v, x) : min(v, x);
			}
			s_f[g][k] = v;
		}
		__syncthreads();
		if (threadIdx.x < W) {
			uint32_t v = s_f[0][k];
			for (uint32_t gg = 1; gg < GROUPS; ++gg) v = (k & 3) == 1 ? max(v, s_f[gg][k]) : min(v, s_f[gg][k]);
			// the finish reads u64 words combined with max: [0] ~min key, [1] max key, [2] / [3] ~first row of -0.0 / +0.0 (0 = none)
			s_fin[k] = (k & 3) == 1 ? (unsigned long long)v : ((k & 3) == 0 || v != 0xffffffffu ? ~(unsigned long long)v : 0ull);
		}
		__syncthreads();
		bounds_finish(p, b, seg, s_fin);
	}
}

struct FlatRequant {
	float m_src[4], m_dst[4];
	uint32_t smask[4], dmask[4]; // low-byte masks of the source / destination storage type (0: the float itself)
	uint8_t sq[4], dq[4];
};
struct FlatLane { float mn, sc, m_src, m_dst; uint32_t smask, dmask, sq, dq; };
// one scalar: quant.h:135 (float -> fixed), :167-169 (fixed -> fixed), :182 (fixed -> float); the bytes of the
// slot above the destination storage type keep their old contents
__device__ __forceinline__ uint32_t requant_f32_lane(uint32_t bits, const FlatLane &r)
{
	if (r.sq == 0 && r.dq == 0) return bits;
	unsigned long long q;
	if (r.sq) q = bits & r.smask;
	else q = (unsigned long long)__fadd_rn(__fmul_rn(__fdiv_rn(__fsub_rn(__uint_as_float(bits), r.mn), r.sc), r.m_dst), 0.5f);
	if (r.sq && r.dq) {
		const unsigned long long ms = (unsigned long long)(long long)(int32_t)((1u << r.sq) - 1u), md = (unsigned long long)(long long)(int32_t)((1u << r.dq) - 1u);
		q = q / ms * md + q % ms * md / ms;
	}
	if (r.dq) return (bits & ~r.dmask) | ((uint32_t)q & r.dmask);
	return __float_as_uint(__fadd_rn(__fmul_rn(__fdiv_rn((float)q, r.m_src), r.sc), r.mn));
}
// MODE 0: every component has its own source / destination quantization; 1: all components float -> the same number
// of bits (the quantizer of `-q`); 2: all components the same number of bits -> float (requant(clear))
template <int MODE>
__global__ void __launch_bounds__(256) k_requant_f32(ListParams p, FlatRequant r, RequantSeg sg)
{
	const uint32_t seg = blockIdx.y;
	const uint32_t ncomp = (uint32_t)p.ncomp;
	const uint32_t nscal = sg.rownum[seg] * ncomp;
	uint4 *__restrict__ rows4 = (uint4 *)(p.rows + (size_t)sg.rowbase[seg] * p.stride);
	uint32_t *__restrict__ rows1 = (uint32_t *)rows4;
	// bounds rows: min at row 0, scale at row 2, `ncomp` floats each
	const float *__restrict__ bounds = (const float *)(sg.bounds + (size_t)seg * sg.pitch);
	const uint32_t G = gridDim.x * blockDim.x; // a multiple of ncomp
	const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t nvec = nscal >> 2;
	FlatLane L[4];
#pragma unroll
	for (int m = 0; m < 4; ++m) {
		const int j = (int)((4u * tid + (uint32_t)m) % ncomp);
		L[m].mn = bounds[j]; L[m].sc = bounds[2 * ncomp + j];
		const int ju = MODE == 0 ? j : 0;
		L[m].m_src = r.m_src[ju]; L[m].m_dst = r.m_dst[ju];
		L[m].smask = r.smask[ju]; L[m].dmask = r.dmask[ju]; L[m].sq = r.sq[ju]; L[m].dq = r.dq[ju];
	}
	auto one = [&](uint32_t bits, const FlatLane &l) -> uint32_t {
		if (MODE == 1) {
			const unsigned long long q = (unsigned long long)__fadd_rn(__fmul_rn(__fdiv_rn(__fsub_rn(__uint_as_float(bits), l.mn), l.sc), L[0].m_dst), 0.5f);
			return (bits & ~L[0].dmask) | ((uint32_t)q & L[0].dmask);
		}
		if (MODE == 2) return __float_as_uint(__fadd_rn(__fmul_rn(__fdiv_rn((float)(unsigned long long)(bits & L[0].smask), L[0].m_src), l.sc), l.mn));
		return requant_f32_lane(bits, l);
	};
	for (uint32_t v = tid; v < nvec; v += FLAT_UNROLL * G) {
		uint4 q[FLAT_UNROLL];
#pragma unroll
		for (int u = 0; u < FLAT_UNROLL; ++u) {
			const uint32_t vv = v + (uint32_t)u * G;
			if (vv < nvec) q[u] = rows4[vv];
		}
#pragma unroll
		for (int u = 0; u < FLAT_UNROLL; ++u) {
			const uint32_t vv = v + (uint32_t)u * G;
			if (vv < nvec) {
				q[u].x = one(q[u].x, L[0]); q[u].y = one(q[u].y, L[1]); q[u].z = one(q[u].z, L[2]); q[u].w = one(q[u].w, L[3]);
				rows4[vv] = q[u];
			}
		}
	}
	if (tid == 0)
		for (uint32_t e = nvec << 2; e < nscal; ++e) {
			const int j = (int)(e % ncomp);
			FlatLane l;
			l.mn = bounds[j]; l.sc = bounds[2 * ncomp + j]; l.m_src = r.m_src[j]; l.m_dst = r.m_dst[j];
			l.smask = r.smask[j]; l.dmask = r.dmask[j]; l.sq = r.sq[j]; l.dq = r.dq[j];
			rows1[e] = requant_f32_lane(rows1[e], l);
		}
}

// largest segment of the list (rows)
static uint32_t max_seg_rows(const DevList &dl)
{
	uint32_t mx = 0;
	for (uint32_t r : dl.h_rownum) mx = r > mx ? r : mx;
	return mx;
}

// blocks per segment of a flat kernel: the whole grid (blocks * nseg) should be ~8 blocks of 256 per SM, a
// multiple of ncomp per segment, not more than the work of the largest segment
template <typename K>
static uint32_t flat_blocks(hb_ctx *ctx, K kernel, uint32_t nvec, uint32_t ncomp, uint32_t nseg)
{
	// exactly one wave: as many blocks as are resident at once (a second, partly filled wave costs as much as a full one
	// in a kernel that lasts a few tens of microseconds)
	static const char *env = getenv("HARRY_B200_FLAT_BLOCKS_PER_SM"); // tuning knob
	int resident = 0;
	if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kernel, 256, 0) != cudaSuccess || resident < 1) resident = 4;
	const uint32_t per_sm = env && atoi(env) > 0 ? (uint32_t)atoi(env) : (uint32_t)resident;
	uint32_t blocks = (uint32_t)((uint64_t)ctx->sm_count * per_sm / nseg);
	const uint32_t need = hb_div_up(nvec ? nvec : 1, 256 * FLAT_UNROLL);
	if (blocks > need) blocks = need;
	if (blocks == 0) blocks = 1;
	blocks = blocks / ncomp * ncomp; // a multiple of ncomp (rounded down: never more than one wave)
	if (blocks == 0) blocks = ncomp;
	return blocks;
}

static uint32_t pick_rows_per_step(hb_ctx *ctx, const ListParams &p, uint32_t nrows, uint32_t nseg)
{
	// ~8 resident blocks of 256 threads per SM over all segments, threads = rows_per_step * ncomp
	const uint64_t want_threads = (uint64_t)ctx->sm_count * 8 * 256 / nseg + 256;
	uint64_t rps = want_threads / (uint64_t)p.ncomp;
	if (rps > nrows) rps = nrows;
	if (rps == 0) rps = 1;
	return (uint32_t)rps;
}

// set_bounds of every segment of list l; with `groups` the scale rows too (set_scale, quant.h:46-96)
int hb_list_bounds(hb_dmesh *m, uint32_t l, const uint8_t *groups)
{
	hb_ctx *ctx = m->ctx;
	DevList &dl = m->lists[l];
	const ListParams &p = dl.p;
	if (p.ncomp == 0) return 0;
	for (int j = 0; j < p.ncomp; ++j)
		if (p.quant[j]) return hb_fail(ctx, HB_ERR_INVALID, "set_bounds on a quantized list (the reference only calls it on unquantized data)");
	const uint32_t nseg = m->nseg;
	const size_t words = (size_t)nseg * 4 * p.ncomp;
	unsigned long long *scratch = nullptr;
	HB_CUDA(ctx, cudaMallocAsync((void **)&scratch, sizeof(unsigned long long) * words + sizeof(uint32_t) * nseg, ctx->stream));
	HB_CUDA(ctx, cudaMemsetAsync(scratch, 0, sizeof(unsigned long long) * words + sizeof(uint32_t) * nseg, ctx->stream));
	BoundsSeg b;
	memset(&b, 0, sizeof b);
	b.rowbase = dl.d_rowbase; b.rownum = dl.d_rownum;
	b.scratch = scratch; b.ticket = (uint32_t *)(scratch + words);
	b.bounds = dl.d_bounds; b.pitch = dl.bounds_pitch;
	b.have_groups = groups != nullptr;
	if (groups) memcpy(b.groups, groups, p.ncomp);
	b.err = ctx->d_err;
	const uint32_t mxrows = max_seg_rows(dl);
	uint32_t *partial = nullptr;
	if (flat_f32(p) && (uint64_t)mxrows * p.ncomp < 0xffffffffull) {
		const uint32_t nvec = (mxrows * (uint32_t)p.ncomp) >> 2, nc = (uint32_t)p.ncomp;
		const dim3 grid(p.ncomp == 1 ? flat_blocks(ctx, k_bounds_reduce_f32<1>, nvec, nc, nseg) : p.ncomp == 2 ? flat_blocks(ctx, k_bounds_reduce_f32<2>, nvec, nc, nseg)
		                : p.ncomp == 3 ? flat_blocks(ctx, k_bounds_reduce_f32<3>, nvec, nc, nseg) : flat_blocks(ctx, k_bounds_reduce_f32<4>, nvec, nc, nseg), nseg);
		HB_CUDA(ctx, cudaMallocAsync((void **)&partial, sizeof(uint32_t) * 4 * p.ncomp * (size_t)grid.x * nseg, ctx->stream));
		b.partial = partial;
		switch (p.ncomp) {
		case 1: HB_LAUNCH(ctx, k_bounds_reduce_f32<1>, grid, 256, 0, p, b); break;
		case 2: HB_LAUNCH(ctx, k_bounds_reduce_f32<2>, grid, 256, 0, p, b); break;
		case 3: HB_LAUNCH(ctx, k_bounds_reduce_f32<3>, grid, 256, 0, p, b); break;
		default: HB_LAUNCH(ctx, k_bounds_reduce_f32<4>, grid, 256, 0, p, b); break;
		}
	} else {
		const uint32_t rps = pick_rows_per_step(ctx, p, mxrows, nseg);
		const dim3 grid(hb_div_up((uint64_t)rps * p.ncomp, 256), nseg);
		HB_LAUNCH(ctx, k_bounds_reduce, grid, 256, 0, p, b, rps);
	}
	HB_CUDA(ctx, cudaFreeAsync(scratch, ctx->stream));
	if (partial) HB_CUDA(ctx, cudaFreeAsync(partial, ctx->stream));
	return 0;
}

__global__ void k_scale(ListParams p, BoundsSeg b)
{
	if (threadIdx.x != 0) return;
	scale_segment(p, b.bounds + (size_t)blockIdx.x * b.pitch, b.groups, b.err);
}

// set_scale alone (the bounds rows were supplied by the caller)
int hb_list_scale(hb_dmesh *m, uint32_t l, const uint8_t *groups)
{
	hb_ctx *ctx = m->ctx;
	DevList &dl = m->lists[l];
	if (dl.p.ncomp == 0) return 0;
	BoundsSeg b;
	memset(&b, 0, sizeof b);
	b.bounds = dl.d_bounds; b.pitch = dl.bounds_pitch;
	b.have_groups = 1;
	memcpy(b.groups, groups, dl.p.ncomp);
	b.err = ctx->d_err;
	HB_LAUNCH(ctx, k_scale, m->nseg, 32, 0, dl.p, b);
	return 0;
}

int hb_list_requant(hb_dmesh *m, uint32_t l, const uint8_t *new_quant)
{
	hb_ctx *ctx = m->ctx;
	DevList &dl = m->lists[l];
	ListParams &p = dl.p;
	if (p.ncomp == 0) return 0;
	RequantParams rq;
	bool any = false;
	for (int j = 0; j < p.ncomp; ++j) {
		rq.dq[j] = new_quant[j];
		if (new_quant[j] > 8 * hb_type_size(p.type[j])) return hb_fail(ctx, HB_ERR_INVALID, "requant: %d bits do not fit component %d", new_quant[j], j);
		if (p.quant[j] == 0 && new_quant[j] == 0) continue;
		any = true;
		if (p.type[j] == HB_DOUBLE) return hb_fail(ctx, HB_ERR_UNSUPPORTED, "requant: double lists (the reference shifts an int by >= 32 bits, undefined)");
		if (p.quant[j] > 31 || new_quant[j] > 31) return hb_fail(ctx, HB_ERR_UNSUPPORTED, "requant: more than 31 bits (the reference computes 1 << q in int, undefined)");
	}
	const uint32_t nseg = m->nseg;
	const uint32_t mxrows = max_seg_rows(dl);
	RequantSeg sg;
	sg.rowbase = dl.d_rowbase; sg.rownum = dl.d_rownum; sg.bounds = dl.d_bounds; sg.pitch = dl.bounds_pitch;
	bool flat = any && mxrows && flat_f32(p) && (uint64_t)mxrows * p.ncomp < 0xffffffffull;
	for (int j = 0; flat && j < p.ncomp; ++j) flat = p.quant[j] <= 31 && new_quant[j] <= 31;
	if (flat) {
		FlatRequant fr;
		memset(&fr, 0, sizeof fr);
		bool uni_q = true, uni_d = true; // float -> one width / one width -> float
		for (int j = 0; j < p.ncomp; ++j) {
			const int sq = p.quant[j], dq = new_quant[j];
			fr.sq[j] = (uint8_t)sq; fr.dq[j] = (uint8_t)dq;
			fr.m_src[j] = sq ? (float)(int32_t)((1u << sq) - 1u) : 0.f;
			fr.m_dst[j] = dq ? (float)(int32_t)((1u << dq) - 1u) : 0.f;
			fr.smask[j] = sq ? (sq <= 8 ? 0xffu : sq <= 16 ? 0xffffu : 0xffffffffu) : 0u;
			fr.dmask[j] = dq ? (dq <= 8 ? 0xffu : dq <= 16 ? 0xffffu : 0xffffffffu) : 0u;
			uni_q = uni_q && sq == 0 && dq != 0 && dq == new_quant[0];
			uni_d = uni_d && dq == 0 && sq != 0 && sq == p.quant[0];
		}
		const uint32_t nvec = (mxrows * (uint32_t)p.ncomp) >> 2, nc = (uint32_t)p.ncomp;
		if (uni_q) { const dim3 grid(flat_blocks(ctx, k_requant_f32<1>, nvec, nc, nseg), nseg); HB_LAUNCH(ctx, k_requant_f32<1>, grid, 256, 0, p, fr, sg); }
		else if (uni_d) { const dim3 grid(flat_blocks(ctx, k_requant_f32<2>, nvec, nc, nseg), nseg); HB_LAUNCH(ctx, k_requant_f32<2>, grid, 256, 0, p, fr, sg); }
		else { const dim3 grid(flat_blocks(ctx, k_requant_f32<0>, nvec, nc, nseg), nseg); HB_LAUNCH(ctx, k_requant_f32<0>, grid, 256, 0, p, fr, sg); }
	} else if (any && mxrows) {
		const uint32_t rps = pick_rows_per_step(ctx, p, mxrows, nseg);
		const dim3 grid(hb_div_up((uint64_t)rps * p.ncomp, 256), nseg);
		HB_LAUNCH(ctx, k_requant, grid, 256, 0, p, rq, sg, rps);
	}
	for (int j = 0; j < p.ncomp; ++j) p.quant[j] = new_quant[j];
	hb_list_desc tmp;
	tmp.ncomp = (uint16_t)p.ncomp;
	for (int j = 0; j < p.ncomp; ++j) { tmp.type[j] = p.type[j]; tmp.quant[j] = p.quant[j]; tmp.offset[j] = p.offset[j]; }
	tmp.rows = p.rows; tmp.nrows = p.nrows; tmp.stride = p.stride; tmp.target = (uint8_t)p.target;
	hb_fill_list_params(p, tmp);
	return 0;
}
