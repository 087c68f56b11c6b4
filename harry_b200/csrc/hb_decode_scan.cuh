// hb_decode_scan.cuh -- vertex reconstruction as a verified prefix scan of floor-affine maps.
//
// Problem (AttrDecoder::vtx_post, formats/hry/attrcode.h:443-470): in traversal order
//   x[i] = decodeDelta(residual[i], divround(sum_k predict(x[a_k], x[b_k], x[c_k]), K))
// and one operand of (almost) every rank i is rank i-1, so the dependency DAG is a chain of depth N.
// The other operands lie about one cut-border length behind (the previous "ring" of the traversal).
//
// Observation: with all other operands known, rank i is a function of x[i-1] alone, and in its
// regular regime (no saturation in predict(), no escape in decodeDelta()) that function is
//   f_i(x) = floor((x + A_i) / M_i) + B_i          M_i = K (number of parallelograms)
// The family  x -> floor((x + A) / M) + B  is closed under composition:
//   f2(f1(x)) = floor((x + A1 + M1 * r) / (M1 * M2)) + B2 + q,   B1 + A2 = q * M2 + r, 0 <= r < M2
// and once M exceeds the value range the map degenerates to a step  B + [x >= T]  (kept as M == CAP).
// Saturated / escaped regimes are CONSTANT maps (decodeDelta returns delta or hi - delta there,
// prediction.h:46-63), which are in the family as well.  So a window of ranks whose other operands are
// final can be reconstructed by a parallel prefix composition instead of a sequential walk.
//
// Exactness does not rest on that algebra.  Every sweep (one thread per rank of the window):
//   A. window = ranks [done, done + slots).  Each thread reads the record of its rank (k_scan_prep:
//      which operand is rank i - 1, the highest other operand) and the operand values, notes whether
//      its rank reads a window rank other than its predecessor (the first such rank E cuts the window:
//      nothing at or after E can be validated this sweep), and builds the map of its rank: the regular
//      floor-affine form, or -- when a previous sweep left a guess g for x[i-1] and the exact step
//      disagrees with the regular form at g -- the constant f_i(g).
//   B. inclusive composition scan inside each warp; the warp totals travel to the other CTAs of the
//      cluster (st.async completing a transaction barrier); every CTA classifies the totals -- a total
//      is a step, so fed with the two possible outputs of its predecessor it either gives one value
//      (anchor) or copies the predecessor's choice -- and a warp finds the value in front of it from the
//      nearest anchor with a few bit operations -> the presumed input of every thread.
//   C. every thread runs the EXACT reference step (ScanOps::predict / dec == IntOps, prediction.h)
//      from its presumed input; one repair turn inside the warp redoes the step of a thread whose left
//      neighbour produced something else.
//   D. verification: the boundary in front of a thread is good iff its left neighbour's exact result
//      is the input the thread used.  Thread 0 starts from x[done - 1], which is final; by induction
//      every rank up to the first bad boundary is the sequential result.  done = min(that, E).
// Machine mapping: one thread-block cluster per (list, component) -- the components of an integer list
// are independent chains, so they run on different SMs.  A sweep is bounded by dependent-instruction
// latency and by the two hand-offs (warp totals; values + limits through a cluster barrier), not by
// bandwidth: the window is dealt evenly to the CTAs so that an SM holds only a few active warps, the
// record of the next sweep is copied into shared memory behind the barrier (cp.async).
// Progress is at least one rank per sweep for ANY input (thread 0 is always exact); the algebra only
// decides how far `done` moves.  Ranks with more than SCAN_WIDE candidates (sphere poles) end the
// window and are evaluated by a whole CTA when `done` reaches them (integer sums commute); windows
// of a few ranks (higher-order dependence on the chain: the first ring of a sphere, irregular
// triangulations) are handed to a sequential walker (one warp: prepare 32 ranks ahead / execute).
#pragma once
#include <cooperative_groups.h>
#include "hb_decode_spec.cuh"

namespace cgs = cooperative_groups;

#define SCAN_NTB 512
#define SCAN_NWARP (SCAN_NTB / 32)
#define SCAN_KIN 4            // candidates held inline in a ScanRec
#define SCAN_WIDE 32          // more candidates: evaluated cooperatively when first in the window
#define SCAN_MAXC 16
#define SCAN_SEQ_TRIGGER 48   // a sweep that advances by fewer ranks hands over to the sequential walker
#define SCAN_SEQ_MIN 256
#define SCAN_SEQ_MAX 2048

// Per rank, written once by k_scan_prep (connectivity only, shared by all sweeps): everything a sweep
// needs besides values, in one 64-byte line.  Which operand is the predecessor rank i - 1 is structural,
// so the classification happens here and not in the sweeps.
struct alignas(16) ScanRec {
	uint32_t hdr;               // bits 0..1 kind (SpecArgs::kind); bits 2..9: per inline candidate 0 = no operand is rank i-1,
	                            // 1 = v0 is, 2 = v1 is, 3 = other pattern; bit 10: not a regular rank; bits 11..: K
	uint32_t aux;               // kind 1: CSR offset of the first candidate; kind 2: source rank
	uint32_t farp1;             // 1 + highest operand rank other than i - 1 (0: none)
	uint32_t tri[3 * SCAN_KIN]; // the first four candidates (rank triples)
	uint32_t pad;
};
#define SCAN_HDR_IRREGULAR 0x400u

__global__ void __launch_bounds__(256) k_scan_prep(const uint8_t *__restrict__ kind, const uint32_t *__restrict__ src, const uint32_t *__restrict__ cand_off,
                                                   const uint32_t *__restrict__ cand, uint32_t n, ScanRec *__restrict__ out)
{
	const uint32_t iraw = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t i = iraw < n ? iraw : n - 1; // threads behind the end redo the last rank (their words are not stored)
	const uint32_t kd = kind[i];
	const uint32_t c0 = cand_off[i], K = cand_off[i + 1] - c0;
	const uint32_t pr = i - 1; // 0xffffffff for i == 0: never equal to an operand
	uint32_t hdr = kd | (K << 11), farp1 = 0, npred = 0;
	uint32_t tri[3 * SCAN_KIN];
#pragma unroll
	for (int w = 0; w < 3 * SCAN_KIN; ++w) tri[w] = 0;
	if (kd == 2) {
		if (src[i] != pr) farp1 = src[i] + 1;
	} else if (kd == 1 && K <= SCAN_WIDE) { // (a wider rank ends every window and is evaluated from the CSR by a whole CTA: its
		                                     // record is the header alone -- scanning the 4472 candidates of a pole here cost 0.5 ms)
		for (uint32_t j = 0; j < K; ++j) {
			const uint32_t r0 = cand[3 * (size_t)(c0 + j)], r1 = cand[3 * (size_t)(c0 + j) + 1], r2 = cand[3 * (size_t)(c0 + j) + 2];
			const uint32_t nv = (uint32_t)(r0 == pr) + (uint32_t)(r1 == pr) + (uint32_t)(r2 == pr);
			if (r0 != pr) farp1 = max(farp1, r0 + 1);
			if (r1 != pr) farp1 = max(farp1, r1 + 1);
			if (r2 != pr) farp1 = max(farp1, r2 + 1);
			uint32_t code = 0;
			if (nv == 1 && r0 == pr) code = 1;
			else if (nv == 1 && r1 == pr) code = 2;
			else if (nv) code = 3;
			if (code) ++npred;
			if (code == 3) hdr |= SCAN_HDR_IRREGULAR;
			if (j < SCAN_KIN) {
				hdr |= code << (2 + 2 * j);
				tri[3 * j] = r0; tri[3 * j + 1] = r1; tri[3 * j + 2] = r2;
			}
		}
		if (npred > 1 || (K > SCAN_KIN && K <= SCAN_WIDE)) hdr |= SCAN_HDR_IRREGULAR;
	}
	if (iraw >= n) return;
	uint4 *o = (uint4 *)(out + i);
	o[0] = make_uint4(hdr, kd == 2 ? src[i] : c0, farp1, tri[0]);
	o[1] = make_uint4(tri[1], tri[2], tri[3], tri[4]);
	o[2] = make_uint4(tri[5], tri[6], tri[7], tri[8]);
	o[3] = make_uint4(tri[9], tri[10], tri[11], 0u);
}

// ---- floor-affine maps --------------------------------------------------------------------------
// floor(c / m), m > 0
template <typename SW, typename W> __device__ __noinline__ SW scan_floordiv_slow(SW c, W m)
{
	return c >= 0 ? (SW)((W)c / m) : -(SW)(((W)(-c) + m - 1) / m);
}

// x -> floor((x + A) / M) + B with 0 <= A < M <= CAP = 2^width, for 0 <= x < CAP.  M == CAP is the
// degenerate step B + [x + A >= CAP]; a constant is (CAP, 0, c).
// 8/16-bit storage types: M = 2^sm when sm < 32 (the common case: 1, 2 or 4 parallelograms, constants,
// steps -- shifts only); other divisors keep M in sm >> 8 with 0xff in the low byte.
template <typename T> struct alignas(16) FMap {
	typedef uint32_t W;
	typedef int32_t SW;
	static constexpr int CB = 8 * (int)sizeof(T);
	uint32_t sm, A;
	int32_t B;
	uint32_t pad_;
	static __device__ __forceinline__ FMap identity() { FMap f; f.sm = 0; f.A = 0; f.B = 0; return f; }
	static __device__ __forceinline__ FMap constant(SW c, int cb) { FMap f; f.sm = (uint32_t)cb; f.A = 0; f.B = c; return f; }
	__device__ __forceinline__ uint32_t divisor() const { return sm < 32u ? 1u << sm : sm >> 8; }
	__device__ __forceinline__ bool is_step(int cb) const { return sm == (uint32_t)cb; } // takes the values B and B + 1 only
	static __device__ __forceinline__ FMap make(uint32_t M, uint32_t A, SW B) // M <= CAP
	{
		FMap f;
		f.sm = (M & (M - 1u)) == 0u ? (uint32_t)(31 - __clz((int)M)) : ((M << 8) | 0xffu);
		f.A = A;
		f.B = B;
		return f;
	}
	// regular form of one rank: floor((x + araw) / K) + b
	static __device__ __forceinline__ FMap affine(uint32_t K, SW araw, SW b)
	{
		FMap f;
		if ((K & (K - 1u)) == 0u) {
			const int s = 31 - __clz((int)K);
			f.sm = (uint32_t)s;
			f.A = (uint32_t)araw & (K - 1u);
			f.B = b + (araw >> s);
		} else {
			const SW q = scan_floordiv_slow<SW, W>(araw, K);
			f.sm = (K << 8) | 0xffu;
			f.A = (uint32_t)(araw - q * (SW)K);
			f.B = b + q;
		}
		return f;
	}
	__device__ __forceinline__ SW eval(SW x) const
	{
		if (sm < 32u) return ((x + (SW)A) >> sm) + B; // arithmetic shift == floor
		return scan_floordiv_slow<SW, W>(x + (SW)A, sm >> 8) + B;
	}
};
// 32-bit storage types: general form only (64-bit parameters, 128-bit products)
template <> struct alignas(16) FMap<uint32_t> {
	typedef unsigned long long W;
	typedef long long SW;
	W M, A;
	SW B;
	unsigned long long pad_;
	static __device__ __forceinline__ W cap() { return (W)1 << 32; }
	__device__ __forceinline__ bool is_step(int) const { return M == cap(); }
	static __device__ __forceinline__ FMap identity() { FMap f; f.M = 1; f.A = 0; f.B = 0; return f; }
	static __device__ __forceinline__ FMap constant(SW c, int) { FMap f; f.M = cap(); f.A = 0; f.B = c; return f; }
	static __device__ __forceinline__ SW floordiv(SW c, W m)
	{
		if ((m & (m - 1)) == 0) return c >> (63 - __clzll((long long)m));
		return scan_floordiv_slow<SW, W>(c, m);
	}
	static __device__ __forceinline__ FMap affine(uint32_t K, SW araw, SW b)
	{
		FMap f;
		const SW q = floordiv(araw, (W)K);
		f.M = K;
		f.A = (W)(araw - q * (SW)K);
		f.B = b + q;
		return f;
	}
	__device__ __forceinline__ SW eval(SW x) const { return floordiv(x + (SW)A, M) + B; }
};

// apply f1 first, then f2:  floor((floor((x + A1) / M1) + B1 + A2) / M2) + B2
//   = floor((x + A1 + M1 * r) / (M1 * M2)) + B2 + q   with B1 + A2 = q * M2 + r, 0 <= r < M2;
// a product M >= CAP is folded back to the step form (x + A < 2 M there, so the quotient is 0 or 1).
// cb = quantization bits of the component: every value is < 2^cb, so divisors fold at 2^cb
template <typename T> __device__ __forceinline__ FMap<T> scan_compose(const FMap<T> &f1, const FMap<T> &f2, int cb)
{
	const uint32_t CAPV = 1u << cb;
	FMap<T> f;
	if (f1.sm == (uint32_t)cb && f2.sm == (uint32_t)cb) {
		// step after step (all that is left once ~cb halvings have been composed): f1 takes the
		// values B1 and B1 + 1
		const uint32_t y0 = f1.B + (int)f2.A >= (int)CAPV ? 1u : 0u;
		const uint32_t y1 = f1.B + 1 + (int)f2.A >= (int)CAPV ? 1u : 0u;
		f.sm = (uint32_t)cb;
		f.A = y1 != y0 ? f1.A : 0u;
		f.B = f2.B + (int)y0;
		return f;
	}
	const int Cc = f1.B + (int)f2.A;
	if ((f1.sm | f2.sm) < 32u) {
		// both divisors are powers of two (1, 2, 4 parallelograms; constants; steps): shifts only
		const uint32_t r = (uint32_t)Cc & ((1u << f2.sm) - 1u);
		const uint32_t s = f1.sm + f2.sm;
		f.B = f2.B + (Cc >> f2.sm);
		f.sm = s;
		f.A = f1.A + (r << f1.sm);
		if (s > (uint32_t)cb) {
			const unsigned long long thr = (1ull << s) - ((unsigned long long)f1.A + ((unsigned long long)r << f1.sm));
			f.sm = (uint32_t)cb;
			f.A = thr >= CAPV ? 0u : CAPV - (uint32_t)thr;
		}
		return f;
	}
	// some divisor is not a power of two (3, 5, 6, 7 .. parallelograms).  |Cc| < 2^24 and M2 <= 2^16
	// for in-range data: a float quotient with one correction step is exact there, and for anything
	// else the map is garbage that the verification discards anyway.
	const uint32_t M1 = f1.divisor(), M2 = f2.divisor();
	int q = __float2int_rd(__fdividef((float)Cc, (float)M2));
	int rr = Cc - q * (int)M2;
	if (rr < 0) { --q; rr += (int)M2; }
	if (rr >= (int)M2) { ++q; rr -= (int)M2; }
	const uint32_t r = (uint32_t)rr;
	const unsigned long long M = (unsigned long long)M1 * M2;
	const unsigned long long A = (unsigned long long)f1.A + (unsigned long long)M1 * r;
	f.B = f2.B + q;
	if (M >= CAPV) {
		const unsigned long long thr = M - A; // >= 1
		f.sm = (uint32_t)cb;
		f.A = thr >= CAPV ? 0u : CAPV - (uint32_t)thr;
	} else {
		const uint32_t m = (uint32_t)M;
		f.sm = (m & (m - 1u)) == 0u ? (uint32_t)(31 - __clz((int)m)) : ((m << 8) | 0xffu);
		f.A = (uint32_t)A;
	}
	return f;
}
template <> __device__ __noinline__ FMap<uint32_t> scan_compose<uint32_t>(const FMap<uint32_t> &f1, const FMap<uint32_t> &f2, int)
{
	typedef unsigned long long W;
	typedef long long SW;
	typedef unsigned __int128 WW;
	const W CAP = FMap<uint32_t>::cap();
	const SW Cc = f1.B + (SW)f2.A;
	const SW q = FMap<uint32_t>::floordiv(Cc, f2.M);
	const W r = (W)(Cc - q * (SW)f2.M);
	const WW M = (WW)f1.M * (WW)f2.M;
	const WW A = (WW)f1.A + (WW)f1.M * (WW)r;
	FMap<uint32_t> f;
	f.B = f2.B + q;
	if (M >= (WW)CAP) {
		const WW thr = M - A;
		f.M = CAP;
		f.A = thr >= (WW)CAP ? (W)0 : (W)((WW)CAP - thr);
	} else {
		f.M = (W)M;
		f.A = (W)A;
	}
	return f;
}

template <typename T> __device__ __forceinline__ FMap<T> scan_shfl_up(const FMap<T> &m, int d)
{
	FMap<T> r;
	r.sm = __shfl_up_sync(0xffffffffu, m.sm, d);
	r.A = __shfl_up_sync(0xffffffffu, m.A, d);
	r.B = __shfl_up_sync(0xffffffffu, m.B, d);
	return r;
}
template <> __device__ __forceinline__ FMap<uint32_t> scan_shfl_up<uint32_t>(const FMap<uint32_t> &m, int d)
{
	FMap<uint32_t> r;
	r.M = __shfl_up_sync(0xffffffffu, m.M, d);
	r.A = __shfl_up_sync(0xffffffffu, m.A, d);
	r.B = __shfl_up_sync(0xffffffffu, m.B, d);
	return r;
}
template <typename T> __device__ __forceinline__ FMap<T> scan_shfl(const FMap<T> &m, int lane)
{
	FMap<T> r;
	r.sm = __shfl_sync(0xffffffffu, m.sm, lane);
	r.A = __shfl_sync(0xffffffffu, m.A, lane);
	r.B = __shfl_sync(0xffffffffu, m.B, lane);
	return r;
}
template <> __device__ __forceinline__ FMap<uint32_t> scan_shfl<uint32_t>(const FMap<uint32_t> &m, int lane)
{
	FMap<uint32_t> r;
	r.M = __shfl_sync(0xffffffffu, m.M, lane);
	r.A = __shfl_sync(0xffffffffu, m.A, lane);
	r.B = __shfl_sync(0xffffffffu, m.B, lane);
	return r;
}

// ---- cluster plumbing: asynchronous remote shared-memory stores that complete a transaction barrier
// in the destination CTA (st.async + mbarrier complete_tx) -- a consumer warp waits for exactly the
// bytes it was promised instead of for a barrier over the whole cluster
__device__ __forceinline__ uint32_t scan_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t scan_mapa(uint32_t addr, uint32_t cta)
{
	uint32_t r;
	asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
	return r;
}
__device__ __forceinline__ void scan_st_async_v4(uint32_t raddr, uint32_t rbar, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
	asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(raddr), "r"(a), "r"(b), "r"(c), "r"(d), "r"(rbar) : "memory");
}
__device__ __forceinline__ void scan_mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void scan_mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
	asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.release.cta.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool scan_mbar_try_wait(uint32_t bar, uint32_t parity)
{
	uint32_t ok;
	asm volatile("{\n\t.reg .pred P_OUT;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P_OUT, [%1], %2;\n\tselp.b32 %0, 1, 0, P_OUT;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
	return ok != 0;
}

// Shared-memory accesses by 32-bit shared-window address.  (In a cluster kernel nvcc derives the
// address of a dynamically indexed shared array that is captured by a lambda through the cluster
// window -- an S2R SR_CgaCtaId in front of every access; the sequential walker cannot afford that.)
template <int BYTES> __device__ __forceinline__ uint32_t scan_lds(uint32_t addr);
template <> __device__ __forceinline__ uint32_t scan_lds<1>(uint32_t addr) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }
template <> __device__ __forceinline__ uint32_t scan_lds<2>(uint32_t addr) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }
template <> __device__ __forceinline__ uint32_t scan_lds<4>(uint32_t addr) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }
template <int BYTES> __device__ __forceinline__ void scan_sts(uint32_t addr, uint32_t v);
template <> __device__ __forceinline__ void scan_sts<1>(uint32_t addr, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
template <> __device__ __forceinline__ void scan_sts<2>(uint32_t addr, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
template <> __device__ __forceinline__ void scan_sts<4>(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ uint4 scan_lds_v4(uint32_t addr)
{
	uint4 v;
	asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
	return v;
}
__device__ __forceinline__ void scan_sts_v4(uint32_t addr, uint4 v) { asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }

// L2-coherent scalar load (values are rewritten between sweeps by other SMs of the cluster)
__device__ __forceinline__ uint32_t scan_ld(const uint8_t *p) { return __ldcg((const unsigned char *)p); }
__device__ __forceinline__ uint32_t scan_ld(const uint16_t *p) { return __ldcg((const unsigned short *)p); }
__device__ __forceinline__ uint32_t scan_ld(const uint32_t *p) { return __ldcg((const unsigned int *)p); }

// Values travel as 32-bit words inside the kernel.  For 8/16-bit storage types the reference's
// modulo-2^width arithmetic (prediction.h:46-63,121-137) is restated on 32-bit words -- the native
// narrow types cost a byte-permute after every operation:
//   predict: v1 < v2 ? max(v0 - (v2 - v1), 0) : min(v0 + (v1 - v2), hi)   (the wrapped sum of the
//            reference is < v0 exactly when the true sum exceeds the type, which also selects hi)
//   dec:     the reference expression with every intermediate reduced modulo 2^width
template <typename T> struct ScanOps {
	static constexpr uint32_t TM = (1u << (8 * sizeof(T))) - 1u;
	static __device__ __forceinline__ uint32_t predict(uint32_t v0, uint32_t v1, uint32_t v2, uint32_t hi)
	{
		const int s = (int)v0 + (int)v1 - (int)v2;
		return v1 < v2 ? (uint32_t)max(s, 0) : (uint32_t)min(s, (int)hi);
	}
	static __device__ __forceinline__ uint32_t dec(uint32_t delta, uint32_t pred, uint32_t hi)
	{
		const uint32_t room = (hi - pred) & TM;
		const uint32_t pm1 = (pred - 1u) & TM;
		const uint32_t bal = min(room, pm1);
		const uint32_t half = delta >> 1;
		uint32_t r = (delta & 1u) ? pred + ~half : pred + half;
		if (half > bal) r = room >= pred ? pred + delta - bal - 1u : pred - delta + bal;
		if (pred == 0u) r = delta;
		return r & TM;
	}
};
template <> struct ScanOps<uint32_t> {
	static __device__ __forceinline__ uint32_t predict(uint32_t v0, uint32_t v1, uint32_t v2, uint32_t hi) { return IntOps<uint32_t>::predict_hi(v0, v1, v2, hi); }
	static __device__ __forceinline__ uint32_t dec(uint32_t delta, uint32_t pred, uint32_t hi) { return IntOps<uint32_t>::dec_hi(delta, pred, hi); }
};

// (sum + K / 2) / K of transform.h:90-91 for K > 2
template <typename W> __device__ __noinline__ W scan_divk(W sum, uint32_t K) { return (sum + (W)(K >> 1)) / (W)K; }
template <typename T> __device__ __forceinline__ uint32_t scan_mean(typename FMap<T>::W sum, uint32_t K)
{
	typedef typename FMap<T>::W W;
	if (K == 2) return (uint32_t)(T)((sum + 1) >> 1);
	if (K <= 1) return (uint32_t)(T)sum; // K == 0: sum == 0
	if (K == 4) return (uint32_t)(T)((sum + 2) >> 2);
	return (uint32_t)(T)scan_divk<W>(sum, K);
}

enum { SCAN_NONE = 0, SCAN_CONST = 1, SCAN_COPY = 2, SCAN_STD = 3, SCAN_OPAQUE = 4 };

// the exact step of a regular rank for one component: one candidate reads the predecessor value `cur`
// at operand position varpos (0: v0, 1: v1); S = sum of the other candidates' predictions
template <typename T>
__device__ __forceinline__ uint32_t scan_std_step(uint32_t cur, uint32_t va, uint32_t vb, uint32_t delta, typename FMap<T>::W S, uint32_t K, uint32_t varpos, uint32_t hi)
{
	typedef typename FMap<T>::W W;
	const uint32_t p = varpos ? ScanOps<T>::predict(va, cur, vb, hi) : ScanOps<T>::predict(cur, va, vb, hi);
	return ScanOps<T>::dec(delta, scan_mean<T>(S + (W)p, K), hi);
}

// One component of the generic step (any number of candidates, any operand pattern).  Operand r:
// rank `predrank` reads pv, everything else comes from the value array.
template <typename T>
__device__ __noinline__ uint32_t scan_generic_step(const uint32_t *__restrict__ cand, const T *xc, uint32_t cs, uint32_t predrank, uint32_t pv,
                                                   uint32_t c0, uint32_t K, uint32_t delta, uint32_t hi)
{
	typedef typename FMap<T>::W W;
	W sum = 0;
	for (uint32_t k = 0; k < K; ++k) {
		const uint32_t *tr = cand + 3 * (size_t)(c0 + k);
		uint32_t v[3];
#pragma unroll
		for (int o = 0; o < 3; ++o) {
			const uint32_t r = __ldg(tr + o);
			v[o] = r == predrank ? pv : scan_ld(xc + (size_t)r * cs);
		}
		sum += (W)ScanOps<T>::predict(v[0], v[1], v[2], hi);
	}
	return ScanOps<T>::dec(delta, scan_mean<T>(sum, K), hi);
}

#ifdef SCAN_DEBUG
#define SCAN_CLK(v) const long long v = clock64()
#define SCAN_PROBE(i, ta, tb) pr[i] += (tb) - (ta)
#else
#define SCAN_CLK(v)
#define SCAN_PROBE(i, ta, tb)
#endif

// NTB threads per CTA: 512 (one CTA per SM; a single mesh wants the widest window), or 256 / 128 with two / four CTAs
// per SM -- a batch has more chains than SMs, and resident chains hide each other's latencies and barriers
// MINB = CTAs per SM the register allocation is sized for: 512 / NTB, or 8 with 128 threads (64 registers, ~250 bytes of
// spills) when there are more than four chains per SM to keep resident -- measured on spheres of 100K vertices: 195 meshes
// (585 chains) 5.1 ms with 4 CTAs per SM, 390 meshes (1170 chains) 8.2 ms with 8 per SM = 20.9 instead of 26.2 us per mesh
template <typename T, int NTB, int MINB>
__global__ void __launch_bounds__(NTB, MINB) k_decode_vertex_scan(const SpecArgs *__restrict__ args, uint32_t ncomp)
{
	constexpr uint32_t NWARP = NTB / 32;
	typedef FMap<T> Map;
	typedef typename Map::W W;
	typedef typename Map::SW SW;
	constexpr int NC = 1;                              // components per thread group (one cluster per component)
	constexpr uint32_t G = 32;                         // ranks per warp
	constexpr uint32_t NSEG = NWARP * G;          // ranks per CTA and sweep
	cgs::cluster_group cluster = cgs::this_cluster();
	const uint32_t C = cluster.num_blocks();
	const uint32_t crank = cluster.block_rank();
	const uint32_t job = blockIdx.x / C;               // (list, component)
	const uint32_t list = job / ncomp;
	const uint32_t c = job % ncomp;                    // my component
	const SpecArgs a = args[list];
	const uint32_t RS = ncomp == 3 ? 4 : ncomp;        // elements per value record
	const ScanRec *__restrict__ srec = (const ScanRec *)a.srec;
	const uint32_t n = a.n, base = a.base;              // this chain: ranks [base, n) of the concatenated order
	const uint32_t t = threadIdx.x, lane = t & 31, warp = t >> 5;
	const uint32_t grp = lane;
	constexpr bool live = true;
	const uint32_t q = t;                               // my rank slot within the CTA
	const T *__restrict__ rc = (const T *)a.resid + c;  // component views: element r at [r * RS]
	T *xc = (T *)a.x + c;
	const uint32_t NSEGT = C * NSEG;                    // ranks per window

	__shared__ Map s_all[SCAN_MAXC * NWARP]; // warp totals of the whole cluster (pushed by their owners), active warps only
	__shared__ __align__(8) unsigned long long s_bar; // transaction barrier of the incoming totals
	__shared__ uint32_t s_eva[SCAN_MAXC * NWARP];  // per element: its output if the predecessor says B
	__shared__ uint32_t s_eam[SCAN_MAXC * NWARP / 32], s_ebm[SCAN_MAXC * NWARP / 32]; // anchor / broken-shortcut bit masks
	__shared__ uint32_t s_ndE[SCAN_MAXC];       // per CTA: first rank that reads a non-final window value
	__shared__ uint32_t s_ndF[SCAN_MAXC];       // per CTA: first rank behind a failed boundary check
	__shared__ uint32_t s_start[NSEG + 1][NC];  // presumed start value of every slot (+ of the next CTA's first slot)
	__shared__ uint32_t s_minE, s_minF;
	__shared__ uint32_t s_stop;                 // length of the sequential stretch just walked (pushed by CTA 0)
	__shared__ unsigned long long s_sum[4];
	__shared__ T s_seq[SCAN_SEQ_MAX];           // values of a sequential stretch
	__shared__ __align__(16) uint32_t s_item[2][32][16]; // work items of the sequential stretch (two batches)
	__shared__ __align__(16) uint4 s_pref[NTB][3];  // record of the rank this thread will most likely own in the next sweep

	const uint32_t hi = (uint32_t)IntOps<T>::mask(a.bits[c]);
	const int cb = a.bits[c];
	const int cbw = cb;
	const int lgC = 31 - __clz((int)C);

	uint32_t done = base, gend = base, est = NSEGT;
	bool widehead = false, seqmode = false, lastseq = false;
	uint32_t seqlen = SCAN_SEQ_MIN, smallrun = 0;
	uint32_t lastadv = 0, pref_i = 0xffffffffu; // advance of the last sweep; rank whose record sits in s_pref[t]
	uint32_t prefb;
	asm volatile("mov.u32 %0, %1;" : "=r"(prefb) : "r"(scan_smem_u32(&s_pref[threadIdx.x][0])));
	unsigned long long sweeps = 0, fails = 0, wides = 0, capped = 0, nseq = 0, nfallback = 0;
	long long cyA = 0, cyB = 0, cyC = 0, cyD = 0;
#ifdef SCAN_DEBUG
	long long pr[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
#endif
	auto csync = [&]() { if (C > 1) cluster.sync(); else __syncthreads(); };
	const uint32_t bar = scan_smem_u32(&s_bar);
	// shared-window addresses of the sequential walker's buffers, pinned in registers (an opaque move keeps
	// the compiler from re-deriving them in front of every access)
	uint32_t seqb, itemb;
	asm volatile("mov.u32 %0, %1;" : "=r"(seqb) : "r"(scan_smem_u32(s_seq)));
	asm volatile("mov.u32 %0, %1;" : "=r"(itemb) : "r"(scan_smem_u32(s_item)));
	uint32_t bar_parity = 0;
	if (t == 0) {
		scan_mbar_init(bar, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	cluster.sync();

	while (done < n) {
		const long long tA = clock64();
		// ---------------------------------------------------------------- wide rank at the head of the window
		// (a sweep that cannot move `done` has found one: see phase A)
		if (widehead) {
			const uint32_t K0 = __ldg(&srec[done].hdr) >> 11;
			if (crank == 0) {
				if (t < 4) s_sum[t] = 0ull;
				__syncthreads();
				const uint32_t c0 = __ldg(&srec[done].aux);
				unsigned long long sum = 0ull;
				if (live) {
					for (uint32_t k = q; k < K0; k += NSEG) {
						const uint32_t *tr = a.cand + 3 * (size_t)(c0 + k);
						const uint32_t v0 = scan_ld(xc + (size_t)tr[0] * RS), v1 = scan_ld(xc + (size_t)tr[1] * RS), v2 = scan_ld(xc + (size_t)tr[2] * RS);
						sum += (unsigned long long)ScanOps<T>::predict(v0, v1, v2, hi);
					}
					atomicAdd(&s_sum[0], sum);
				}
				__syncthreads();
				if (t == 0) xc[(size_t)done * RS] = (T)ScanOps<T>::dec(rc[(size_t)done * RS], (uint32_t)(T)hb_divround_i64((long long)s_sum[0], (int)K0), hi);
			}
			csync();
			++done;
			if (gend < done) gend = done;
			widehead = false;
			++wides;
			++sweeps;
			cyD += clock64() - tA;
			continue;
		}
		// ---------------------------------------------------------------- sequential stretch (one warp)
		// The window was cut after a few ranks: some operand other than rank i - 1 is a few ranks old
		// (the first ring of a sphere -- every vertex predicted from the two before it -- and most
		// vertices of an irregular triangulation).  Warp 0 of CTA 0 then walks the next `seqlen`
		// ranks in order, 32 at a time: (P) in parallel, every lane fetches the record of its rank
		// and every operand that is already known -- final values from L2, values of earlier batches
		// of this stretch from shared memory -- and sums the candidates that are complete;
		// (S) the lanes take turns: a turn only reads the operands produced inside the batch.
		if (seqmode) {
			const uint32_t len = n - done < seqlen ? n - done : seqlen;
			if (crank == 0 && warp == 0) {
				// Two roles in one warp.  PREPARE (all lanes, one rank each, one batch ahead): fetch
				// the record, every operand that is already known (final values from L2, values of
				// earlier batches from the stretch buffer), sum the complete candidates, and park a
				// 64-byte work item in shared memory.  EXECUTE (lane 0): walk the 32 items of the
				// current batch in order; only operands produced inside the last two batches are
				// read at that point.  The loads of PREPARE are in flight while lane 0 executes.
				uint32_t stop = len;
				constexpr int TB = (int)sizeof(T);
				// item words: 0 kind | K << 2 | late mask << 8, 1 residual, 2 known sum (or the value for kinds 0 / 2),
				// 3 aux, 4..15 operands: the value if known, else the index into the stretch buffer
				// PREPARE is split in two so that its second wave of loads (operand values) is in
				// flight while lane 0 executes: issue() reads the record (prefetched into L2 two
				// batches ahead) and starts the operand loads, finish() turns them into the item.
				uint32_t p_tri[3 * SCAN_KIN], p_ov[3 * SCAN_KIN], p_kd = 0, p_K = 0, p_aux = 0, p_res = 0, p_late = 0, p_cv = 0;
				bool p_valid = false, p_inl = false;
				auto issue = [&](uint32_t b0) -> bool {
					const uint32_t k0 = b0 + lane;
					p_valid = k0 < len;
					const uint32_t i = done + (p_valid ? k0 : 0);
					const uint4 *sr = (const uint4 *)(srec + i);
					const uint4 q0 = __ldg(sr), q1 = __ldg(sr + 1), q2 = __ldg(sr + 2), q3 = __ldg(sr + 3);
					if ((unsigned long long)i + 64 < n) asm volatile("prefetch.global.L2 [%0];" ::"l"(srec + i + 64));
					p_tri[0] = q0.w; p_tri[1] = q1.x; p_tri[2] = q1.y; p_tri[3] = q1.z; p_tri[4] = q1.w; p_tri[5] = q2.x;
					p_tri[6] = q2.y; p_tri[7] = q2.z; p_tri[8] = q2.w; p_tri[9] = q3.x; p_tri[10] = q3.y; p_tri[11] = q3.z;
					p_kd = q0.x & 3u; p_K = q0.x >> 11; p_aux = q0.y;
					p_res = rc[(size_t)i * RS];
					p_inl = p_valid && p_kd == 1 && p_K <= SCAN_KIN;
					p_late = 0;
#pragma unroll
					for (int w = 0; w < 3 * SCAN_KIN; ++w) {
						p_ov[w] = 0;
						if (p_inl && (uint32_t)(w / 3) < p_K) {
							const uint32_t r = p_tri[w];
							if (r < done) p_ov[w] = scan_ld(xc + (size_t)r * RS);
							else if (r + 32 < done + b0) p_ov[w] = scan_lds<TB>(seqb + TB * (r - done)); // produced at least a batch before the one executing now
							else { p_late |= 1u << w; p_ov[w] = r - done; }
						}
					}
					p_cv = 0;
					if (p_valid && p_kd == 0) p_cv = scan_ld(xc + (size_t)i * RS);
					if (p_valid && p_kd == 2 && p_aux < done) p_cv = scan_ld(xc + (size_t)p_aux * RS);
					return p_valid && p_kd == 1 && p_K > SCAN_WIDE;
				};
				// item (48 bytes): word 0 = kind | K << 2 | slot-0 late mask << 8 | slot-1 late mask << 11 | slot-1 used << 14,
				// 1 residual, 2 sum of the complete candidates (or the value / stretch index for kinds 0 / 2),
				// 3 reciprocal of K, 4..6 operands of slot 0, 7..9 operands of slot 1 (the value if known, else the
				// stretch index).  Kinds: 0 value known, 1 regular (at most two candidates read this or the
				// previous batch), 2 copy of a stretch value, 3 generic step.
				auto finish = [&](uint32_t buf, bool wide) {
					W sknown = 0;
					uint32_t nl2 = 0, m0 = 0, m1 = 0, s0[3] = { 0, 0, 0 }, s1[3] = { 0, 0, 0 };
#pragma unroll
					for (int cj = 0; cj < SCAN_KIN; ++cj) {
						if (!(p_inl && (uint32_t)cj < p_K)) continue;
						const uint32_t lm = (p_late >> (3 * cj)) & 7u;
						if (lm == 0) sknown += (W)ScanOps<T>::predict(p_ov[3 * cj], p_ov[3 * cj + 1], p_ov[3 * cj + 2], hi);
						else {
							if (nl2 == 0) { m0 = lm; s0[0] = p_ov[3 * cj]; s0[1] = p_ov[3 * cj + 1]; s0[2] = p_ov[3 * cj + 2]; }
							else if (nl2 == 1) { m1 = lm; s1[0] = p_ov[3 * cj]; s1[1] = p_ov[3 * cj + 1]; s1[2] = p_ov[3 * cj + 2]; }
							++nl2;
						}
					}
					// (sum + K / 2) / K as a multiplication: exact for sum < 2^27, K <= 32 (K == 2: a shift by one)
					uint32_t w2 = (uint32_t)sknown, w3 = p_K > 1 ? (uint32_t)((0x100000000ull + p_K - 1) / p_K) : 0u, kind = p_kd;
					if (p_valid && p_kd == 0) w2 = p_cv;
					if (p_valid && p_kd == 2) {
						if (p_aux < done) { w2 = p_cv; kind = 0; }      // a known value
						else w2 = p_aux - done;                           // stretch index
					}
					if (p_valid && p_kd == 1 && (!p_inl || nl2 > 2 || sknown > (W)0xffffffffu)) { kind = 3; w3 = p_aux; } // generic step
					if (!p_valid) kind = 0;
					const uint32_t o = itemb + 64 * (buf * 32 + lane);
					scan_sts_v4(o, make_uint4(kind | (p_K << 2) | (m0 << 8) | (m1 << 11) | ((nl2 > 1 ? 1u : 0u) << 14) | (wide ? 0x80000000u : 0u), p_res, w2, w3));
					scan_sts_v4(o + 16, make_uint4(s0[0], s0[1], s0[2], s1[0]));
					scan_sts_v4(o + 32, make_uint4(s1[1], s1[2], 0u, 0u));
				};
				bool wide_next = issue(0);
				finish(0, wide_next);
				__syncwarp();
				for (uint32_t b0 = 0; b0 < len; b0 += 32) {
					const uint32_t buf = (b0 >> 5) & 1u;
					const unsigned wm = __ballot_sync(0xffffffffu, wide_next);
					const uint32_t nl = wm ? (uint32_t)__ffs((int)wm) - 1u : min(32u, len - b0);
					const bool more = b0 + 32 < len && !wm;
					SCAN_CLK(q0c);
					if (more) wide_next = issue(b0 + 32);
					SCAN_CLK(q1c);
					if (lane == 0) {
#pragma unroll 1
						for (uint32_t j = 0; j < nl; ++j) {
							const uint32_t it = itemb + 64 * (buf * 32 + j);
							const uint4 h = scan_lds_v4(it), oa = scan_lds_v4(it + 16), ob = scan_lds_v4(it + 32);
							const uint32_t kind = h.x & 3u, K = (h.x >> 2) & 0x3fu;
							uint32_t val = h.z;
							if (kind == 1) {
								// branch-free: slot 0 always holds a candidate here (a rank without late
								// operands was folded into kind 0 ... or has K == 0), slot 1 is masked
								const uint32_t m = h.x >> 8;
								const uint32_t v0 = (m & 1u) ? scan_lds<TB>(seqb + TB * oa.x) : oa.x;
								const uint32_t v1 = (m & 2u) ? scan_lds<TB>(seqb + TB * oa.y) : oa.y;
								const uint32_t v2 = (m & 4u) ? scan_lds<TB>(seqb + TB * oa.z) : oa.z;
								const uint32_t u0 = (m & 8u) ? scan_lds<TB>(seqb + TB * oa.w) : oa.w;
								const uint32_t u1 = (m & 16u) ? scan_lds<TB>(seqb + TB * ob.x) : ob.x;
								const uint32_t u2 = (m & 32u) ? scan_lds<TB>(seqb + TB * ob.y) : ob.y;
								W sum = (W)h.z;
								if (m & 7u) sum += (W)ScanOps<T>::predict(v0, v1, v2, hi);
								if (m & 64u) sum += (W)ScanOps<T>::predict(u0, u1, u2, hi);
								uint32_t pred;
								if constexpr (sizeof(T) < 4) pred = K <= 1 ? (uint32_t)sum : __umulhi((uint32_t)sum + (K >> 1), h.w);
								else pred = scan_mean<T>(sum, K);
								val = ScanOps<T>::dec(h.y, pred, hi);
							} else if (kind == 2) {
								val = scan_lds<TB>(seqb + TB * h.z);
							} else if (kind == 3) {
								val = scan_generic_step<T>(a.cand, xc, RS, 0xffffffffu, 0u, h.w, K, h.y, hi); // earlier ranks of the stretch are in the value array already
							}
							scan_sts<TB>(seqb + TB * (b0 + j), val);
							xc[(size_t)(done + b0 + j) * RS] = (T)val;
						}
					}
					__syncwarp();
					SCAN_CLK(q2c);
					if (more) finish(buf ^ 1u, wide_next);
					__syncwarp();
					SCAN_CLK(q3c);
					SCAN_PROBE(0, q0c, q1c); SCAN_PROBE(1, q1c, q2c); SCAN_PROBE(2, q2c, q3c);
					if (wm) { stop = b0 + nl; break; }
				}
				if (lane < C) {
					uint32_t *dst = C > 1 ? cluster.map_shared_rank(&s_stop, lane) : &s_stop;
					*dst = stop;
				}
			}
			csync();
			const uint32_t stop = s_stop;
			csync(); // s_stop may be rewritten by the next stretch
			if (stop == 0) widehead = true;
			done += stop;
			if (gend < done) gend = done;
			seqmode = false;
			++nseq;
			++sweeps;
			cyD += clock64() - tA;
			continue;
		}
		// ---------------------------------------------------------------- phase A: operands and the map of my rank
		// the window is dealt evenly to the CTAs of the cluster, whole warps each (instruction issue
		// per SM is what bounds a sweep); warps beyond `wpc` sit the sweep out
		uint32_t wpc = (((est + est / 8 + 8 + G * C - 1) >> lgC)) / G; // active warps per CTA
		if (wpc > NWARP) wpc = NWARP;
		const uint32_t per = wpc * G;                       // slots per CTA
		// warp totals travel to the CTA that owns them and to every CTA behind it; thread 0 announces
		// how many bytes this CTA is going to receive in this sweep
		if (t == 0 && C > 1) scan_mbar_arrive_expect_tx(bar, (crank + 1u) * wpc * (uint32_t)sizeof(Map));
		const uint32_t nact = per << lgC;                   // slots of the window
		const uint32_t gs = crank * per + q;                // my slot in the window
		const unsigned long long i64 = (unsigned long long)done + gs;
		const bool active = live && warp < wpc && i64 < n;
		const uint32_t i = active ? (uint32_t)i64 : n;
		const unsigned long long wend64 = (unsigned long long)done + nact;
		const uint32_t wend = wend64 < n ? (uint32_t)wend64 : n;
		const uint32_t x0 = done > base ? scan_ld(xc + (size_t)(done - 1) * RS) : 0u;
		if (active && (unsigned long long)i + est < n) {
			// the record and residual this slot will most likely need in the next sweep: pull them into L2 now
			asm volatile("prefetch.global.L2 [%0];" ::"l"(srec + i + est));
			if ((lane & 7) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(rc + (size_t)(i + est) * RS));
		}
		const uint32_t x0w = x0;

		uint32_t mode = SCAN_NONE, Kk = 0, vpos = 0, ra = 0, rb = 0, rd = 0;
		bool wr = false;
		W S = 0;
		uint32_t Ecand = 0xffffffffu;
		Map tm = Map::identity();
		if (active) {
			// the record: copied into shared memory during the previous sweep's barrier if this thread
			// owns the rank it expected to own (cp.async below), else from L2 now
			uint4 q0, q1, q2, q3;
			if (i == pref_i) {
				asm volatile("cp.async.wait_all;" ::: "memory");
				q0 = scan_lds_v4(prefb); q1 = scan_lds_v4(prefb + 16); q2 = scan_lds_v4(prefb + 32);
				q3 = make_uint4(0u, 0u, 0u, 0u);
				if ((q0.x >> 11) > 3u) q3 = __ldg((const uint4 *)(srec + i) + 3); // only a fourth candidate lives there
			} else {
				const uint4 *sr = (const uint4 *)(srec + i);
				q0 = __ldg(sr); q1 = __ldg(sr + 1); q2 = __ldg(sr + 2); q3 = __ldg(sr + 3);
			}
			const uint32_t tri[3 * SCAN_KIN] = { q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, q3.z };
			const uint32_t hdr = q0.x, kd = hdr & 3u, K = hdr >> 11, aux = q0.y, farp1 = q0.z;
			const uint32_t res = rc[(size_t)i * RS];
			// operand values of the inline candidates: all loads are issued before any is used
			// (the slot of the predecessor rank is loaded too and ignored)
			uint32_t ov[3 * SCAN_KIN];
#pragma unroll
			for (int w = 0; w < 3 * SCAN_KIN; ++w) {
				ov[w] = 0;
				if (kd == 1 && (uint32_t)(w / 3) < K) ov[w] = scan_ld(xc + (size_t)tri[w] * RS);
			}
			// a guess for x[i - 1] exists if a previous sweep computed it with all operands final
			const bool have_guess = i == done || i - 1 < gend;
			uint32_t g = 0;
			if (i == done) g = x0;
			else if (have_guess) g = scan_ld(xc + (size_t)(i - 1) * RS);
			rd = res;
			Kk = K;
			wr = kd != 0;
			if (farp1 > done || (kd == 1 && K > SCAN_WIDE)) {
				// reads a window value that is not final (or is a wide rank, evaluated when `done`
				// reaches it): nothing from here on can be validated in this sweep
				Ecand = i;
				mode = SCAN_CONST;
				wr = false;
				tm = Map::constant(0, cb);
			} else if (kd == 0) {
				mode = SCAN_CONST;
				ra = scan_ld(xc + (size_t)i * RS);
				tm = Map::constant((SW)ra, cb);
			} else if (kd == 2) {
				if (i > 0 && aux == i - 1) {
					mode = SCAN_COPY;
				} else {
					mode = SCAN_CONST;
					ra = scan_ld(xc + (size_t)aux * RS);
					tm = Map::constant((SW)ra, cb);
				}
			} else if (hdr & SCAN_HDR_IRREGULAR) {
				mode = SCAN_OPAQUE;
				S = (W)aux;
				tm = Map::constant((SW)scan_generic_step<T>(a.cand, xc, RS, i - 1, g, aux, K, res, hi), cb);
			} else {
				W Sacc = 0;
				uint32_t va = 0, vb = 0, vp = 0;
				bool var = false;
#pragma unroll
				for (int j = 0; j < SCAN_KIN; ++j) {
					if ((uint32_t)j >= K) continue;
					const uint32_t code = (hdr >> (2 + 2 * j)) & 3u;
					if (code == 0) {
						Sacc += (W)ScanOps<T>::predict(ov[3 * j], ov[3 * j + 1], ov[3 * j + 2], hi);
					} else {
						var = true;
						vp = code - 1u;
						va = code == 1 ? ov[3 * j + 1] : ov[3 * j];
						vb = ov[3 * j + 2];
					}
				}
				if (!var) {
					mode = SCAN_CONST;
					ra = ScanOps<T>::dec(res, scan_mean<T>(Sacc, K), hi);
					tm = Map::constant((SW)ra, cb);
				} else {
					mode = SCAN_STD;
					vpos = vp;
					ra = va;
					rb = vb;
					S = Sacc;
					const SW half = (SW)(res >> 1);
					const SW sb = (res & 1) ? -(half + 1) : half;
					tm = Map::affine(K, (SW)va - (SW)vb + (SW)Sacc + (SW)(K >> 1), sb);
					if (have_guess) {
						// the head of the window knows its predecessor exactly: a constant, exactly
						const uint32_t yg = scan_std_step<T>(g, va, vb, res, Sacc, K, vp, hi);
						if (i == done || (uint32_t)(T)tm.eval((SW)g) != yg) tm = Map::constant((SW)yg, cb);
					}
				}
			}
		}
		// A map only serves the presumed inputs of LATER ranks (a rank's own value comes from the exact
		// step).  A general-divisor map (3, 5, 6.. parallelograms: the closing vertex of a ring) drags
		// its warp onto the slow composition path; if the window is cut one or two ranks behind it
		// anyway, it is not composed -- the repair turn of phase C fixes the one rank in between.
		if constexpr (sizeof(T) < 4) {
			const uint32_t cutmask = __ballot_sync(0xffffffffu, Ecand != 0xffffffffu);
			if (active && mode == SCAN_STD && tm.sm >= 32u && ((cutmask >> lane) & 6u)) tm = Map::constant(0, cb);
		}
		const long long tB = clock64();
		// ---------------------------------------------------------------- phase B: presumed start values
		// B1. inclusive composition scan inside every active warp; the warp totals go to every CTA
		//     of the cluster (distributed shared memory), element index = cta * wpc + warp.
		Map inc = tm;
		if (warp < wpc) {
#pragma unroll 1
			for (uint32_t d = 1; d < 32; d <<= 1) {
				const Map up = scan_shfl_up<T>(inc, (int)d);
				if (lane >= d) inc = scan_compose<T>(up, inc, cb);
			}
			const Map tot = scan_shfl<T>(inc, 31);
			if (C == 1) {
				if (lane == 0) s_all[warp] = tot; // a single CTA: plain shared memory, block barrier below
			} else if (crank + lane < C) {
				static_assert(sizeof(Map) % 16 == 0, "totals travel as 16-byte vectors");
				const uint32_t *w = (const uint32_t *)&tot;
				const uint32_t dst = scan_mapa(scan_smem_u32(&s_all[crank * wpc + warp]), crank + lane), dbar = scan_mapa(bar, crank + lane);
#pragma unroll
				for (int v = 0; v < (int)(sizeof(Map) / 16); ++v) scan_st_async_v4(dst + 16 * v, dbar, w[4 * v], w[4 * v + 1], w[4 * v + 2], w[4 * v + 3]);
			}
		}
		SCAN_CLK(p1);
		if (t == 0) { s_minE = 0xffffffffu; s_minF = 0xffffffffu; }
		SCAN_CLK(p2);
		SCAN_CLK(p3);
		// [1] wait for the totals promised to this CTA (no barrier: the producers complete the
		// transaction count that thread 0 armed at the top of the sweep)
		if (C == 1) {
			__syncthreads();
		} else {
			while (!scan_mbar_try_wait(bar, bar_parity)) { }
			bar_parity ^= 1u;
		}
		SCAN_CLK(p4);
		// B2. the value in front of my warp: out[j], out[k] = value after warp total k, j = my element
		//     - 1.  Once ~cb halvings are composed a total is a step: it outputs B or B + 1.  Fed
		//     with the two possible outputs of its predecessor, element k yields a (pred said B) or b
		//     (pred said B + 1).  a == b anchors the chain (so does a total that is constant over the
		//     whole value range, and element 0, whose input x[done - 1] is known); b == a + 1 copies
		//     the predecessor's choice.  The nearest anchor at or before element j decides.  Every
		//     CTA classifies the elements in front of its last warp once, one element per thread, and
		//     publishes two bit masks; a warp then finds its anchor with a few bit operations instead
		//     of a scan over the cluster.  An element whose predecessor is not a step yet (long runs
		//     of single-parallelogram vertices) breaks the shortcut: the totals from the anchor on
		//     are then evaluated one after the other.
		const uint32_t nW = (crank + 1u) * wpc;             // elements that reached this CTA
		{
			uint32_t va = 0;
			bool anch = false, bad = false;
			if (t < nW) {
				const Map f = s_all[t];
				if (t == 0) {
					va = (uint32_t)(T)f.eval((SW)x0);
					anch = true;
				} else {
					const Map pf = s_all[t - 1];
					const bool ps = pf.is_step(cb);
					const uint32_t c0 = (uint32_t)(T)f.eval(0), c1 = (uint32_t)(T)f.eval((SW)hi);
					va = (uint32_t)(T)f.eval(pf.B);
					const uint32_t vb = (uint32_t)(T)f.eval(pf.B + 1);
					if (c0 == c1) { va = c0; anch = true; }
					else if (ps && va == vb) anch = true;
					else bad = !ps || vb != va + 1u;
				}
				s_eva[t] = va;
			}
			if (t < SCAN_MAXC * NWARP) {
				const uint32_t am = __ballot_sync(0xffffffffu, anch), bm = __ballot_sync(0xffffffffu, bad);
				if (lane == 0) { s_eam[warp] = am; s_ebm[warp] = bm; }
			}
		}
		__syncthreads();
		uint32_t vstart = x0;
		if (warp < wpc) {
			const int j = (int)(crank * wpc + warp) - 1;
			if (j >= 0) {
				// nearest anchor at or before element j; any bad element in (anchor, j]
				int wd = j >> 5;
				uint32_t am = s_eam[wd] & (0xffffffffu >> (31 - (j & 31)));
				uint32_t badbits = 0, bmask = 0xffffffffu >> (31 - (j & 31));
				while (am == 0) { // element 0 is always an anchor
					badbits |= s_ebm[wd] & bmask;
					--wd;
					am = s_eam[wd];
					bmask = 0xffffffffu;
				}
				const int hb = 31 - __clz((int)am);
				const int ka = (wd << 5) + hb;
				badbits |= s_ebm[wd] & bmask & ~(0xffffffffu >> (31 - hb));
				const uint32_t aka = s_eva[ka];
				if (!badbits) {
					vstart = ka == j ? aka : s_eva[j] + (aka - (uint32_t)s_all[ka].B);
				} else {
					SW v = (SW)aka;
					for (int kk = ka + 1; kk <= j; ++kk) v = (SW)(uint32_t)(T)s_all[kk].eval(v);
					vstart = (uint32_t)(T)v;
					if (lane == 0) ++nfallback;
				}
			}
		}
		const uint32_t vend = (uint32_t)(T)scan_shfl<T>(inc, 31).eval((SW)vstart);
		Map ex = scan_shfl_up<T>(inc, 1);
		if (lane == 0) ex = Map::identity();
		uint32_t start = (uint32_t)(T)ex.eval((SW)vstart);
		if (gs == 0) start = x0; // final by construction
		if (warp < wpc) s_start[q][0] = start; // (slot `per` belongs to the line below)
		if (warp + 1 == wpc && lane == 0) s_start[per][0] = vend;
		__syncthreads();
		const long long tC = clock64();
		// ---------------------------------------------------------------- phase C: the exact step
		auto exact = [&](uint32_t from) -> uint32_t {
			if (mode == SCAN_CONST) return ra;
			if (mode == SCAN_STD) return scan_std_step<T>(from, ra, rb, rd, S, Kk, vpos, hi);
			if (mode == SCAN_OPAQUE) return scan_generic_step<T>(a.cand, xc, RS, i - 1, from, (uint32_t)S, Kk, rd, hi);
			return from;
		};
		uint32_t cur = exact(start);
		// one repair turn inside the warp: a thread whose presumed input differs from what its left
		// neighbour actually produced redoes its step from that value (an isolated bad map -- a rank
		// whose map was not composed, see phase A -- then costs nothing)
		{
			const uint32_t pc = __shfl_up_sync(0xffffffffu, cur, 1);
			if (active && lane > 0 && pc != start) {
				start = pc;
				cur = exact(pc);
			}
		}
		const long long tD = clock64();
		// ---------------------------------------------------------------- phase D: verify, publish, write
		// the boundary in front of a thread is good iff its left neighbour's final result is the
		// input the thread used; across warps: the presumed start of the next warp's first thread
		uint32_t ndE = Ecand, ndF = 0xffffffffu;
		{
			const uint32_t pc = __shfl_up_sync(0xffffffffu, cur, 1);
			if (active) {
				if (lane > 0 && pc != start) ndF = i;
				const bool last = gs + 1 >= nact; // nothing after me in this window
				if (lane == 31 && !last && cur != s_start[q + 1][0]) ndF = min(ndF, i + 1);
				if (wr) xc[(size_t)i * RS] = (T)cur;
			}
		}
		ndE = __reduce_min_sync(0xffffffffu, ndE);
		ndF = __reduce_min_sync(0xffffffffu, ndF);
		if (lane == 0) {
			if (ndE != 0xffffffffu) atomicMin(&s_minE, ndE);
			if (ndF != 0xffffffffu) atomicMin(&s_minF, ndF);
		}
		// next sweep: if it advances like the last one did, this thread owns rank i + lastadv -- start
		// copying that record into shared memory now, behind the barrier
		asm volatile("cp.async.wait_all;" ::: "memory");
		pref_i = 0xffffffffu;
		if (active && lastadv && (unsigned long long)i + lastadv < n) {
			pref_i = i + lastadv;
			const ScanRec *src = srec + pref_i;
#pragma unroll
			for (int v = 0; v < 3; ++v) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(prefb + 16 * v), "l"((const uint4 *)src + v) : "memory");
			asm volatile("cp.async.commit_group;" ::: "memory");
		}
		SCAN_CLK(p5);
		__syncthreads();
		SCAN_CLK(p6);
		if (t < C) {
			uint32_t *dE = C > 1 ? cluster.map_shared_rank(&s_ndE[crank], t) : &s_ndE[crank];
			uint32_t *dF = C > 1 ? cluster.map_shared_rank(&s_ndF[crank], t) : &s_ndF[crank];
			*dE = s_minE;
			*dF = s_minF;
		}
		csync(); // [2] values and the two limits visible
		SCAN_CLK(p7);
		uint32_t E = wend, F = 0xffffffffu;
		for (uint32_t r = 0; r < C; ++r) { E = min(E, s_ndE[r]); F = min(F, s_ndF[r]); }
		const uint32_t newdone = min(E, F);
		if (F < E) ++fails;
		else if (newdone == wend && wend < n) ++capped;
		if (gend < E) gend = E;            // [newdone, E) now holds this sweep's values: guesses for the next one
		if (newdone == done) widehead = true; // only a wide head rank stops slot 0
		else if (E - done < SCAN_SEQ_TRIGGER && newdone < n) {
			// the window itself is short (not a verification failure): the second time in a row a
			// stretch is walked sequentially, twice as long as the previous one if that was one too
			if (++smallrun >= 2) {
				seqmode = true;
				seqlen = lastseq ? (seqlen * 2 < SCAN_SEQ_MAX ? seqlen * 2 : SCAN_SEQ_MAX) : SCAN_SEQ_MIN;
				lastseq = true;
			}
		} else { smallrun = 0; lastseq = false; }
		const uint32_t adv = newdone - done;
		lastadv = adv;
		if (newdone == wend && wend < n) est = est * 2 < NSEGT ? est * 2 : NSEGT;
		else est = adv > est - est / 8 ? adv : est - est / 8;
		done = newdone;
		++sweeps;
		SCAN_PROBE(0, tB, p1); SCAN_PROBE(1, p1, p2); SCAN_PROBE(2, p2, p3); SCAN_PROBE(3, p3, p4); SCAN_PROBE(4, p4, tC); SCAN_PROBE(5, tD, p5); SCAN_PROBE(6, p5, p6); SCAN_PROBE(7, p6, p7);
		{ const long long tE = clock64(); cyA += tB - tA; cyB += tC - tB; cyC += tD - tC; cyD += tE - tD; }
	}
#ifdef SCAN_DEBUG
	if (t == 0 && crank == 0 && nseq > 20) printf("seq probes: issue %lld exec %lld finish %lld total cycles over %llu stretches\n", pr[0], pr[1], pr[2], nseq);
	if (t == 0 && nseq <= 20) printf("cta %u A %lld probes: warpscan %lld sync %lld lvl2 %lld csync1 %lld start %lld walk %lld | verify+write %lld sync %lld push+csync2 %lld (per sweep, %llu sweeps)\n", crank, cyA / (long long)sweeps, pr[0] / (long long)sweeps, pr[1] / (long long)sweeps, pr[2] / (long long)sweeps, pr[3] / (long long)sweeps, pr[4] / (long long)sweeps, cyC / (long long)sweeps, pr[5] / (long long)sweeps, pr[6] / (long long)sweeps, pr[7] / (long long)sweeps, sweeps);
#endif
#ifdef SCAN_DEBUG
	if (crank == 0 && t == 0) printf("comp %u: sweeps %llu fails %llu capped %llu fallback %llu nseq %llu cycles %lld\n", c, sweeps, fails, capped, nfallback, nseq, cyA + cyB + cyC + cyD);
#endif
	if (crank == 0 && t == 0 && a.stats && c == 0) {
		a.stats[0] = sweeps; a.stats[1] = fails | (capped << 32); a.stats[2] = nfallback | (wides << 32) | (nseq << 40); a.stats[3] = n - base;
		a.stats[4] = (unsigned long long)cyA; a.stats[5] = (unsigned long long)cyB; a.stats[6] = (unsigned long long)cyC; a.stats[7] = (unsigned long long)cyD;
	}
}
