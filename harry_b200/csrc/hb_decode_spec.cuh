// hb_decode_spec.cuh -- speculative chunk-parallel reconstruction of vertex lists.
//
// Problem: x[i] = decodeDelta(residual[i], prediction(x[deps(i)])), deps(i) < i, and in a CBM
// traversal one dependency is (almost) always rank i-1, so the level DAG is a chain of depth ~N.
//
// Scheme (exact for ANY input; the speculation only decides how fast `done` advances):
//   ranks [0, done) are final.  The window [done, done + T*B) is cut into T chunks of B consecutive
//   ranks, one thread each.
//
//   HYPOTHESIS MODE (B = 8, integer lists).  Every sweep is a Jacobi step in three phases:
//     1. compute: thread t recomputes its chunk sequentially from the STORED state of the window,
//        three times: assuming the value of its predecessor rank (start - 1) is the stored value
//        + e, e in {-1, 0, +1}, independently per component.  Why: averaged parallelogram prediction
//        halves an upstream error per step, so after a few sweeps every stored value is within +-1
//        of the truth, but a +-1 offset can persist indefinitely through the rounding (e.g. a
//        coordinate that is constant along a ring has two adjacent fixed points).
//        The three runs give, per component, a map  offset-in -> offset-out  (end value relative to
//        the stored end value), or "unknown" if outside +-1.
//     2. resolve: an inclusive prefix composition (block scan) of these maps, starting from offset 0
//        for chunk 0 (its predecessor is final), yields the TRUE offset-in of every chunk, as long as
//        every map along the way is defined.  A chunk is valid iff all chunks before it are valid,
//        its offset-in is known, and every other window value it read is unchanged by the
//        resolution (range query on prefix sums of per-chunk "changed" flags).
//     3. write: valid chunks store the selected trajectories -- exact by induction: each was
//        computed sequentially from final values, from the exact value of its predecessor rank and
//        from window values equal to their final ones; done = start of the first invalid chunk.
//        Invalid chunks store their offset-0 trajectory (plain Jacobi refresh).
//     Chunk 0 is always valid, so progress >= B ranks per sweep.
//   PLAIN MODE (adaptive B, used for lossless float lists -- they pick ONE candidate, no averaging,
//   no contraction -- and whenever hypothesis sweeps stop paying): in-place recomputation; the
//   first rank whose value changed bounds the exact prefix; chunk 0 is exact after the sweep.
#pragma once
#include "hb_lists.cuh"

#define SPEC_THREADS 1024
#define SPEC_HB 8            // chunk length in hypothesis mode
#define SPEC_B_MAX 1024

template <typename T, int NC> struct SpecRec;
template <typename T> struct alignas(sizeof(T) * 1) SpecRec<T, 1> { T c[1]; };
template <typename T> struct alignas(sizeof(T) * 2) SpecRec<T, 2> { T c[2]; };
template <typename T> struct alignas(sizeof(T) * 4) SpecRec<T, 3> { T c[4]; };
template <typename T> struct alignas(sizeof(T) * 4) SpecRec<T, 4> { T c[4]; };

// one reconstruction step for all components of a rank.  `get(r)` returns the record of rank r.
// FP = lossless float list (T == uint32_t holding the IEEE bits).
template <typename T, int NC, bool FP, typename Get>
__device__ __forceinline__ SpecRec<T, NC> spec_step(const SpecArgs &a, Get &&get, uint32_t c0, uint32_t K, const SpecRec<T, NC> &res)
{
	SpecRec<T, NC> out = res;
	const uint32_t *__restrict__ tri = a.cand + 3 * (size_t)c0;
	if (!FP) {
		long long sum[NC];
#pragma unroll
		for (int j = 0; j < NC; ++j) sum[j] = 0;
		for (uint32_t k = 0; k < K; ++k) {
			const SpecRec<T, NC> v0 = get(tri[3 * k]), v1 = get(tri[3 * k + 1]), v2 = get(tri[3 * k + 2]);
#pragma unroll
			for (int j = 0; j < NC; ++j) sum[j] += (long long)IntOps<T>::predict(v0.c[j], v1.c[j], v2.c[j], a.bits[j]);
		}
#pragma unroll
		for (int j = 0; j < NC; ++j) {
			const T pred = K ? (T)hb_divround_i64(sum[j], (int)K) : (T)0;
			out.c[j] = IntOps<T>::dec(res.c[j], pred, a.bits[j]);
		}
	} else {
		double sum[NC];
#pragma unroll
		for (int j = 0; j < NC; ++j) sum[j] = 0.0;
		for (uint32_t k = 0; k < K; ++k) {
			const SpecRec<T, NC> v0 = get(tri[3 * k]), v1 = get(tri[3 * k + 1]), v2 = get(tri[3 * k + 2]);
#pragma unroll
			for (int j = 0; j < NC; ++j)
				sum[j] = __dadd_rn(sum[j], (double)__fadd_rn(__uint_as_float(v0.c[j]), __fsub_rn(__uint_as_float(v1.c[j]), __uint_as_float(v2.c[j]))));
		}
		float avg[NC], best[NC];
#pragma unroll
		for (int j = 0; j < NC; ++j) { avg[j] = K ? __double2float_rn(__ddiv_rn(sum[j], (double)(int)K)) : 0.f; best[j] = FLT_MAX; }
		for (uint32_t k = 0; k < K; ++k) {
			const SpecRec<T, NC> v0 = get(tri[3 * k]), v1 = get(tri[3 * k + 1]), v2 = get(tri[3 * k + 2]);
#pragma unroll
			for (int j = 0; j < NC; ++j)
				best[j] = hb_closest_step(best[j], __fadd_rn(__uint_as_float(v0.c[j]), __fsub_rn(__uint_as_float(v1.c[j]), __uint_as_float(v2.c[j]))), avg[j]);
		}
#pragma unroll
		for (int j = 0; j < NC; ++j) {
			const uint32_t pred = K ? __float_as_uint(best[j]) : 0u;
			out.c[j] = (T)hb_flip_f32(IntOps<uint32_t>::dec((uint32_t)res.c[j], hb_flip_f32(pred), 32));
		}
	}
	return out;
}

template <typename T, int NC>
__device__ __forceinline__ bool spec_equal(const SpecRec<T, NC> &a, const SpecRec<T, NC> &b)
{
	bool eq = true;
#pragma unroll
	for (int j = 0; j < NC; ++j) eq = eq && a.c[j] == b.c[j];
	return eq;
}

// ---- offset maps ----------------------------------------------------------------------------------
// per component 3 entries (offset-in -1, 0, +1 -> index 0, 1, 2) of 2 bits: offset-out index, 3 = unknown.
// component j occupies bits [6j, 6j + 6).
__device__ __forceinline__ uint32_t map_get(uint32_t m, int j, uint32_t e) { return e == 3 ? 3u : (m >> (6 * j + 2 * e)) & 3u; }
// apply a first, then b.  Per component: c[e] = a[e] == unknown ? unknown : b[a[e]]; with b extended
// by a fourth entry "unknown -> unknown" this is one table lookup per entry.
__device__ __forceinline__ uint32_t map_compose(uint32_t a, uint32_t b, int nc)
{
	uint32_t c = 0;
#pragma unroll
	for (int j = 0; j < 4; ++j) {
		if (j >= nc) break;
		const uint32_t tab = ((b >> (6 * j)) & 0x3fu) | 0xc0u; // 4 entries of 2 bits
		const uint32_t aj = (a >> (6 * j)) & 0x3fu;
		const uint32_t c0 = (tab >> (2 * (aj & 3u))) & 3u;
		const uint32_t c1 = (tab >> (2 * ((aj >> 2) & 3u))) & 3u;
		const uint32_t c2 = (tab >> (2 * ((aj >> 4) & 3u))) & 3u;
		c |= (c0 | (c1 << 2) | (c2 << 4)) << (6 * j);
	}
	return c;
}
#define SPEC_MAP_IDENTITY 0x00924924u // every component: 0 -> 0, 1 -> 1, 2 -> 2  (binary 100100 repeated)

// inclusive block scan of maps (composition in thread order)
__device__ __forceinline__ uint32_t block_scan_maps(uint32_t m, int nc, uint32_t *s_warp /* 32 */)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t up = __shfl_up_sync(0xffffffffu, m, d);
		if (lane >= d) m = map_compose(up, m, nc);
	}
	if (lane == 31) s_warp[warp] = m;
	__syncthreads();
	if (warp == 0) {
		uint32_t w = lane < (int)(blockDim.x >> 5) ? s_warp[lane] : SPEC_MAP_IDENTITY;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const uint32_t up = __shfl_up_sync(0xffffffffu, w, d);
			if (lane >= d) w = map_compose(up, w, nc);
		}
		s_warp[lane] = w;
	}
	__syncthreads();
	if (warp > 0) m = map_compose(s_warp[warp - 1], m, nc);
	__syncthreads();
	return m;
}

__device__ __forceinline__ uint32_t block_scan_u32(uint32_t v, uint32_t *s_warp /* 32 */)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t x = v;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t up = __shfl_up_sync(0xffffffffu, x, d);
		if (lane >= d) x += up;
	}
	if (lane == 31) s_warp[warp] = x;
	__syncthreads();
	if (warp == 0) {
		uint32_t w = lane < (int)(blockDim.x >> 5) ? s_warp[lane] : 0u;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const uint32_t up = __shfl_up_sync(0xffffffffu, w, d);
			if (lane >= d) w += up;
		}
		s_warp[lane] = w;
	}
	__syncthreads();
	const uint32_t incl = x + (warp ? s_warp[warp - 1] : 0);
	__syncthreads();
	return incl - v; // exclusive
}

// one CTA per vertex list
// shared-memory trajectories: element (slot, k) of thread t lives at [(slot * SPEC_HB + k) * nthreads + t]
// (slot 0..2 = offset hypotheses, slot 3 = the stored values) -> conflict-free, no local memory
template <typename Rec> struct SpecTraj {
	Rec *base;
	uint32_t nthreads, t;
	__device__ __forceinline__ Rec &at(int slot, uint32_t k) const { return base[((uint32_t)slot * SPEC_HB + k) * nthreads + t]; }
};

template <typename T, int NC> __host__ __device__ constexpr int spec_threads()
{
	// 4 slots * SPEC_HB records per thread within ~200 KB of shared memory
	return (int)sizeof(SpecRec<T, NC>) <= 4 ? 1024 : ((int)sizeof(SpecRec<T, NC>) <= 8 ? 768 : 384);
}

template <typename T, int NC, bool FP>
__global__ void __launch_bounds__(1024, 1) k_decode_vertex_spec(const SpecArgs *__restrict__ args, const uint32_t *__restrict__ chain)
{
	extern __shared__ __align__(16) unsigned char s_dyn[];
	// float lists: chain-like segments are reconstructed by k_decode_vertex_walkf (hb_decode.cu, k_chain_stat)
	if (chain && 4ull * chain[2 * blockIdx.x] >= 3ull * chain[2 * blockIdx.x + 1]) return;
	const SpecArgs a = args[blockIdx.x];
	typedef SpecRec<T, NC> Rec;
	const Rec *__restrict__ resid = (const Rec *)a.resid;
	Rec *x = (Rec *)a.x;
	const uint32_t n = a.n;
	const uint32_t t = threadIdx.x;
	__shared__ uint32_t s_min, s_warp[32], s_excl[SPEC_THREADS], s_incl[SPEC_THREADS];
	__shared__ uint8_t s_inner[SPEC_THREADS];
	const uint32_t NT = blockDim.x;
	SpecTraj<Rec> traj;
	traj.base = (Rec *)s_dyn;
	traj.nthreads = NT;
	traj.t = t;
	uint32_t done = a.base, B = FP ? SPEC_HB : 4 * SPEC_HB; // this chain: ranks [base, n) of the concatenated order
	bool hyp = !FP;            // hypothesis mode
	int poor = 0;              // consecutive hypothesis sweeps that advanced by <= 2 chunks
	unsigned long long sweeps = 0, hsweeps = 0, hadv = 0, unk = 0;
	long long cyc_h = 0, cyc_p = 0;
	uint32_t est = 0;          // decayed maximum of the recent advance (ranks per sweep) ~ one "ring"
	while (done < n) {
		uint32_t newdone;
		const long long tc0 = clock64();
		const bool was_hyp = hyp;
		if (hyp) {
			// ---------------- phase 1: three trajectories from the stored state --------------------
			// A rank becomes final two sweeps after the ranks it depends on: recomputing more than
			// ~2.5x the recent advance is wasted work (and this single SM is instruction bound).
			uint32_t nact = (5 * est / 2) / SPEC_HB + 48;
			if (nact > NT) nact = NT;
			const unsigned long long start64 = (unsigned long long)done + (unsigned long long)t * SPEC_HB;
			const bool active = start64 < n && t < nact;
			const uint32_t start = active ? (uint32_t)start64 : n;
			const uint32_t len = active ? ((n - start < SPEC_HB) ? n - start : SPEC_HB) : 0;
			uint32_t minread = 0xffffffffu; // lowest window rank read besides the predecessor rank
			uint32_t map = SPEC_MAP_IDENTITY;
			if (active) {
				Rec pst = resid[a.base];
				const bool pred_in_window = start > done; // chunk 0: the predecessor is final
				if (pred_in_window) pst = x[start - 1];
				for (uint32_t k = 0; k < len; ++k) traj.at(3, k) = x[start + k];
				uint32_t c0 = a.cand_off[start];
				for (uint32_t k = 0; k < len; ++k) {
					const uint32_t i = start + k;
					const uint32_t c1 = a.cand_off[i + 1];
					const int kind = a.kind[i];
#pragma unroll
					for (int e = 0; e < 3; ++e) {
						auto get = [&](uint32_t r) -> Rec {
							if (r >= start) return traj.at(e, r - start);
							if (pred_in_window && r == start - 1) {
								Rec v = pst;
#pragma unroll
								for (int j = 0; j < NC; ++j) v.c[j] = (T)(v.c[j] + (T)(e - 1));
								return v;
							}
							if (r >= done && r < minread) minread = r;
							return x[r];
						};
						if (kind == 1) traj.at(e, k) = spec_step<T, NC, FP>(a, get, c0, c1 - c0, resid[i]);
						else if (kind == 2) traj.at(e, k) = get(a.src[i]);
						else traj.at(e, k) = traj.at(3, k);
					}
					c0 = c1;
				}
				// offset map: end value under offset-in e, relative to the stored end value
				map = 0;
#pragma unroll
				for (int j = 0; j < NC; ++j)
#pragma unroll
					for (int e = 0; e < 3; ++e) {
						const long long d = (long long)traj.at(e, len - 1).c[j] - (long long)traj.at(3, len - 1).c[j] + 1;
						map |= (uint32_t)((d >= 0 && d <= 2) ? d : 3) << (6 * j + 2 * e);
					}
			}
			__syncthreads(); // every read of the stored state is done
			// ---------------- phase 2: resolve the offsets across chunks ---------------------------
			s_incl[t] = block_scan_maps(map, NC, s_warp);
			__syncthreads();
			const uint32_t before = t == 0 ? SPEC_MAP_IDENTITY : s_incl[t - 1];
			uint32_t ein[NC]; // true offset-in index per component (3 = unknown)
			bool known = true;
#pragma unroll
			for (int j = 0; j < NC; ++j) { ein[j] = map_get(before, j, 1u); known = known && ein[j] != 3u; }
			// which of my ranks change under the selected trajectory?
			bool chg_inner = false, chg_last = false;
			if (active && known) {
				for (uint32_t k = 0; k < len; ++k) {
					bool ch = false;
#pragma unroll
					for (int j = 0; j < NC; ++j) ch = ch || traj.at(ein[j], k).c[j] != traj.at(3, k).c[j];
					if (k + 1 == len) chg_last = ch;
					else chg_inner = chg_inner || ch;
				}
			}
			s_inner[t] = chg_inner ? 1 : 0;
			s_excl[t] = block_scan_u32((chg_inner ? 1u : 0u) + (chg_last ? 1u : 0u), s_warp);
			if (t == 0) s_min = 0xffffffffu;
			__syncthreads();
			bool valid = !active || known;
			if (active && known && minread != 0xffffffffu && t > 0) {
				// other window reads: chunks [dep, t-2] must be entirely unchanged, chunk t-1 unchanged
				// except (possibly) its last rank, which the offset hypothesis accounts for
				const uint32_t dep = (minread - done) / SPEC_HB;
				const uint32_t changed_before = s_excl[t - 1] - s_excl[dep]; // chunks [dep, t-2]
				if (changed_before != 0 || s_inner[t - 1]) valid = false;
			}
			if (!valid) atomicMin(&s_min, t);
			__syncthreads();
			const uint32_t first_bad = s_min; // first invalid chunk (chunk 0 is always valid)
			// ---------------- phase 3: write ----------------------------------------------------------
			if (active) {
				const bool sel = t < first_bad;
				for (uint32_t k = 0; k < len; ++k) {
					const Rec o = traj.at(3, k);
					Rec v = o;
#pragma unroll
					for (int j = 0; j < NC; ++j) v.c[j] = traj.at(sel ? ein[j] : 1, k).c[j];
					if (!spec_equal<T, NC>(v, o)) x[start + k] = v;
				}
			}
			const unsigned long long wend = (unsigned long long)done + (unsigned long long)(first_bad == 0xffffffffu ? nact : first_bad) * SPEC_HB;
			newdone = wend < n ? (uint32_t)wend : n;
			const uint32_t adv = newdone - done;
			// policy: hypothesis sweeps are the default.  After 3 poor ones in a row (no contraction
			// here, e.g. single-candidate predictions at the start of a component) ONE plain sweep
			// with a long exact chunk 0 carries the progress, then hypotheses get another chance.
			est = adv > est - est / 8 ? adv : est - est / 8;
			if (adv <= 2 * SPEC_HB) ++poor;
			else { poor = 0; B = 4 * SPEC_HB; }
			if (poor >= 3) hyp = false;
			++hsweeps;
			hadv += adv;
			if (first_bad != 0xffffffffu && t == first_bad && !known) ++unk;
			__syncthreads();
		} else {
			// ---------------- plain mode: in-place recomputation ---------------------------------------
			if (t == 0) s_min = 0xffffffffu;
			__syncthreads();
			const unsigned long long start64 = (unsigned long long)done + (unsigned long long)t * B;
			uint32_t fc = 0xffffffffu;
			if (start64 < n) {
				const uint32_t start = (uint32_t)start64;
				const uint32_t end = (n - start < B) ? n : start + B;
				uint32_t c0 = a.cand_off[start];
				auto get = [&](uint32_t r) -> Rec { return x[r]; };
				for (uint32_t i = start; i < end; ++i) {
					const uint32_t c1 = a.cand_off[i + 1];
					const int kind = a.kind[i];
					if (kind) {
						const Rec old = x[i];
						const Rec nw = kind == 2 ? x[a.src[i]] : spec_step<T, NC, FP>(a, get, c0, c1 - c0, resid[i]);
						if (!spec_equal<T, NC>(old, nw)) {
							x[i] = nw;
							if (fc == 0xffffffffu) fc = i;
						}
					}
					c0 = c1;
				}
				// chunk 0 starts at `done` and reads only final values: it is exact after this sweep
				// whether or not it changed; if it changed, nothing behind it is validated
				if (t == 0 && fc != 0xffffffffu) fc = end;
			}
			if (fc != 0xffffffffu) atomicMin(&s_min, fc);
			__syncthreads();
			const uint32_t p = s_min;
			const unsigned long long wend = (unsigned long long)done + (unsigned long long)NT * B;
			newdone = p != 0xffffffffu ? p : (wend < n ? (uint32_t)wend : n);
			const uint32_t adv = newdone - done;
			if (FP) {
				if (adv <= 2 * B && B < SPEC_B_MAX) B <<= 1;
				else if (adv >= 16 * B && B > SPEC_HB) B >>= 1;
			} else {
				if (B < SPEC_B_MAX) B <<= 1; // the next fallback uses a longer exact chunk
				hyp = true;
				poor = 2;                    // one more poor hypothesis sweep falls back again
			}
			__syncthreads();
		}
		done = newdone;
		++sweeps;
		if (was_hyp) cyc_h += clock64() - tc0; else cyc_p += clock64() - tc0;
	}
	if (t == 0 && a.stats) { a.stats[0] = sweeps; a.stats[1] = hsweeps; a.stats[2] = sweeps - hsweeps; a.stats[3] = hadv; a.stats[5] = (unsigned long long)cyc_h; a.stats[6] = (unsigned long long)cyc_p; }
	if (unk && a.stats) atomicAdd(&a.stats[4], unk);
}
