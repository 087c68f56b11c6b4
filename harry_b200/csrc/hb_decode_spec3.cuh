// hb_decode_spec3.cuh -- speculative vertex reconstruction, cluster version with the offset
// hypotheses mapped onto lanes.
//
// Algorithm and exactness argument: hb_decode_spec.cuh (hypothesis mode).  Machine mapping:
//   * one thread-block cluster (8 CTAs) per list; the window is a run of chunks of SPEC_HB ranks,
//     chunk g = cta_rank * CPC + q.
//   * a chunk is served by a group of 4 adjacent lanes: lanes 0..2 walk the chunk under the offset
//     hypotheses -1 / 0 / +1, all four lanes cooperate on the loads, the flags and the write-back.
//     1024 threads per CTA = 32 warps hide the shared-memory / ALU latencies of the walk, and the
//     sequential part of a sweep is SPEC_HB steps instead of 3 * SPEC_HB.
//   * the stored state is read-only until the write phase: candidate offsets, rank triples,
//     residuals and all operand values that lie outside the chunk are fetched up front (two
//     dependent waves of independent loads) into shared memory; candidates that do not depend on the
//     hypothesis are predicted once.
//   * maps / change counts are scanned per CTA in shared memory and stitched across CTAs through a
//     few words of global scratch ordered by barrier.cluster (release / acquire at cluster scope).
#pragma once
#include <cooperative_groups.h>
#include <type_traits>
#include "hb_decode_spec.cuh"

namespace cg3 = cooperative_groups;

#define SPEC3_KCAP 2        // candidates per rank cached in shared memory
#define SPEC3_WIDE 48       // ranks with more candidates block speculation and are evaluated by a whole warp
#define SPEC3_CLUSTER 8
#define SPEC3_OPS (SPEC_HB * SPEC3_KCAP * 3)
#define SPEC3_MAXCPC 256    // chunks per CTA (upper bound)
#ifndef SPEC3_CPC_SMALL
#define SPEC3_CPC_SMALL 128
#endif

#define SPEC3_MAXCHUNKS (SPEC3_CLUSTER * SPEC3_MAXCPC)
struct Spec3Scratch {
	uint32_t first_bad[2];
	uint32_t plain_first[2];
	uint32_t pad[28];
	uint32_t map[SPEC3_MAXCHUNKS];   // offset map of every window chunk
	uint32_t flg[SPEC3_MAXCHUNKS];   // per component j, hypothesis e: bit (6j + 2e) inner changed, bit (6j + 2e + 1) last changed
	uint32_t dep[SPEC3_MAXCHUNKS];   // chunk of the lowest other-window rank read, or 0xffffffff
};

// In-place inclusive scan of the first `nvalid` of 2 * blockDim.x maps in shared memory (composition
// in index order).  Entries beyond nvalid are identities; warps that only hold such entries skip the
// arithmetic (they still join the barriers).
__device__ __forceinline__ void block_scan_maps_x2(uint32_t *s_all, int nc, uint32_t *s_warp, uint32_t *s_prev /* blockDim.x words */, uint32_t nvalid)
{
	const uint32_t t = threadIdx.x;
	const int lane = t & 31, warp = t >> 5;
	const bool warp_live = (uint32_t)(warp * 64) < nvalid;
	uint32_t a0 = SPEC_MAP_IDENTITY, m = SPEC_MAP_IDENTITY;
	if (warp_live) {
		a0 = s_all[2 * t];
		m = map_compose(a0, s_all[2 * t + 1], nc);
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const uint32_t up = __shfl_up_sync(0xffffffffu, m, d);
			if (lane >= d) m = map_compose(up, m, nc);
		}
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
	if (warp_live) {
		if (warp > 0) m = map_compose(s_warp[warp - 1], m, nc);
		s_prev[t] = m;
	}
	__syncthreads();
	if (warp_live) {
		const uint32_t pre = t ? s_prev[t - 1] : SPEC_MAP_IDENTITY;
		s_all[2 * t] = map_compose(pre, a0, nc);
		s_all[2 * t + 1] = m;
	}
	__syncthreads();
}
// in-place exclusive scan of the first `nvalid` of 2 * blockDim.x counts (the rest are zero and are
// not needed: range queries only touch indices below nvalid)
__device__ __forceinline__ void block_scan_u32_x2(uint32_t *s_all, uint32_t *s_warp, uint32_t nvalid)
{
	const uint32_t t = threadIdx.x;
	const int lane = t & 31, warp = t >> 5;
	const bool warp_live = (uint32_t)(warp * 64) < nvalid;
	uint32_t a0 = 0, v = 0, x = 0;
	if (warp_live) {
		a0 = s_all[2 * t];
		v = a0 + s_all[2 * t + 1];
		x = v;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const uint32_t up = __shfl_up_sync(0xffffffffu, x, d);
			if (lane >= d) x += up;
		}
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
	if (warp_live) {
		const uint32_t ex = x - v + (warp ? s_warp[warp - 1] : 0u);
		s_all[2 * t] = ex;
		s_all[2 * t + 1] = ex + a0;
	}
	__syncthreads();
}

template <typename T, int NC> __host__ __device__ constexpr int spec3_cpc()
{
	return (int)sizeof(SpecRec<T, NC>) <= 8 ? SPEC3_CPC_SMALL : 128;
}
// per chunk: 3 trajectories + stored + residual (5 * HB records), OPS operand records, OPS codes,
// HB candidate counts, HB kinds, HB candidate offsets
template <typename T, int NC> __host__ __device__ constexpr size_t spec3_smem()
{
	return (size_t)spec3_cpc<T, NC>() * ((5 * SPEC_HB + SPEC3_OPS) * sizeof(SpecRec<T, NC>) + SPEC3_OPS + 2 * SPEC_HB + 4 * SPEC_HB);
	// = 3 trajectories + stored + residual records, operand cache, codes, counts, kinds, offsets
}

template <typename T, int NC>
__global__ void __launch_bounds__((4 * spec3_cpc<T, NC>()), 1) k_decode_vertex_spec3(const SpecArgs *__restrict__ args, Spec3Scratch *__restrict__ scratch_all,
                                                                                    uint32_t *__restrict__ g_excl_all, uint8_t *__restrict__ g_inner_all)
{
	typedef SpecRec<T, NC> Rec;
	typedef typename std::conditional<(sizeof(T) <= 2), uint32_t, unsigned long long>::type Acc;
	constexpr uint32_t CPC = (uint32_t)spec3_cpc<T, NC>();
	cg3::cluster_group cluster = cg3::this_cluster();
	const uint32_t C = cluster.num_blocks();
	const uint32_t rank = cluster.block_rank();
	const uint32_t list = blockIdx.x / C;
	const SpecArgs a = args[list];
	Spec3Scratch *sc = scratch_all + list;
	(void)g_excl_all; (void)g_inner_all;
	const Rec *__restrict__ resid = (const Rec *)a.resid;
	Rec *x = (Rec *)a.x;
	const Rec *xs = (const Rec *)a.x; // stored state (read-only until the write phase)
	const uint32_t n = a.n;
	const uint32_t t = threadIdx.x;
	const uint32_t q = t >> 2;          // chunk within the CTA
	const uint32_t L = t & 3;           // lane within the chunk group: 0..2 = hypothesis, 3 = helper
	const uint32_t g = q * C + rank;    // global chunk index in the window: interleaved over the cluster (load balance)

	extern __shared__ __align__(16) unsigned char s_dyn[];
	Rec *s_traj = (Rec *)s_dyn;                                   // [((k * CPC + q) * 3 + e)]   hypotheses e = 0..2
	Rec *s_old = s_traj + (size_t)SPEC_HB * CPC * 3;               // [k * CPC + q]  stored values
	Rec *s_res = s_old + (size_t)SPEC_HB * CPC;                    // [k * CPC + q]  residuals
	Rec *s_opv = s_res + (size_t)SPEC_HB * CPC;                    // [op * CPC + q] operand values / precomputed candidates
	uint32_t *s_coff = (uint32_t *)(s_opv + (size_t)SPEC3_OPS * CPC); // [k * CPC + q]
	uint8_t *s_code = (uint8_t *)(s_coff + (size_t)SPEC_HB * CPC); // [op * CPC + q]
	uint8_t *s_K = s_code + (size_t)SPEC3_OPS * CPC;               // [k * CPC + q]  candidates (255 = slow path)
	uint8_t *s_kind = s_K + (size_t)SPEC_HB * CPC;                 // [k * CPC + q]
	__shared__ uint32_t s_warp[32], s_all[SPEC3_MAXCHUNKS], s_excl[SPEC3_MAXCHUNKS];
	__shared__ uint8_t s_inner[SPEC3_MAXCHUNKS];
	__shared__ uint32_t s_first;
	const uint32_t NCH = C * CPC; // window chunks = 2 * blockDim.x

	auto TR = [&](uint32_t e, uint32_t k) -> Rec & { return s_traj[(k * CPC + q) * 3 + e]; };

	T hi[NC];
#pragma unroll
	for (int c = 0; c < NC; ++c) hi[c] = IntOps<T>::mask(a.bits[c]);
	uint32_t done = 0, Bp = 4 * SPEC_HB, est = 0;
	bool hyp = true;
	int poor = 0;
	uint32_t parity = 0, pparity = 0;
	unsigned long long sweeps = 0, hsweeps = 0, hadv = 0;
	long long cy0 = 0, cy1 = 0, cy2 = 0, cy3 = 0;
	while (done < n) {
		const long long tA = clock64();
		uint32_t newdone;
		if (hyp) {
			uint32_t nact = (5 * est / 2) / SPEC_HB + 64;
			if (nact > NCH) nact = NCH;
			const unsigned long long start64 = (unsigned long long)done + (unsigned long long)g * SPEC_HB;
			const bool active = start64 < n && g < nact;
			const uint32_t start = active ? (uint32_t)start64 : n;
			const uint32_t len = active ? ((n - start < SPEC_HB) ? n - start : SPEC_HB) : 0;
			const bool pred_in_window = start > done;
			uint32_t minread = 0xffffffffu;
			Rec pst = resid[0];
			// ------------------------------------------------------------------ phase 0: cooperative hoisted loads
			if (active) {
				if (pred_in_window) pst = xs[start - 1];
				// wave 0: lane L fetches ranks k = L and L + 4
#pragma unroll
				for (int h = 0; h < SPEC_HB / 4; ++h) {
					const uint32_t k = L + 4 * h;
					if (k < len) {
						const uint32_t b = a.cand_off[start + k], e2 = a.cand_off[start + k + 1];
						s_old[k * CPC + q] = xs[start + k];
						s_res[k * CPC + q] = resid[start + k];
						s_kind[k * CPC + q] = a.kind[start + k];
						s_coff[k * CPC + q] = b;
						s_K[k * CPC + q] = (uint8_t)(e2 - b > SPEC3_WIDE ? 254 : (e2 - b > SPEC3_KCAP ? 255 : e2 - b));
					}
				}
			}
			__syncwarp();
			// a chunk holding a wide rank is not speculated on: it blocks the exact prefix until `done`
			// reaches the wide rank, which is then evaluated by a whole warp (plain branch below)
			uint32_t blocked = 0;
			if (active) {
#pragma unroll
				for (int h = 0; h < SPEC_HB / 4; ++h) {
					const uint32_t k = L + 4 * h;
					if (k < len && s_kind[k * CPC + q] == 1 && s_K[k * CPC + q] == 254) blocked = 1;
				}
			}
			blocked |= __shfl_xor_sync(0xffffffffu, blocked, 1);
			blocked |= __shfl_xor_sync(0xffffffffu, blocked, 2);
			const bool compute = active && !blocked;
			if (compute) {
				// waves 1 + 2: lane L owns operands op = L * 12 .. L * 12 + 11 (= ranks 2L, 2L + 1)
#pragma unroll
				for (int half = 0; half < 2; ++half) {
					uint32_t r6[6];
					bool have[6];
#pragma unroll
					for (int m = 0; m < 6; ++m) {
						const uint32_t op = L * 12 + half * 6 + m;
						const uint32_t k = op / (3 * SPEC3_KCAP), j = (op / 3) % SPEC3_KCAP, o = op % 3;
						const uint32_t K = k < len && s_kind[k * CPC + q] == 1 ? s_K[k * CPC + q] : 0;
						have[m] = K != 255 && j < K;
						r6[m] = have[m] ? a.cand[3 * (size_t)(s_coff[k * CPC + q] + j) + o] : 0;
					}
					Rec v6[6];
					bool far6[6];
#pragma unroll
					for (int m = 0; m < 6; ++m) {
						const uint32_t r = r6[m];
						far6[m] = have[m] && !(r >= start) && !(pred_in_window && r == start - 1);
						v6[m] = xs[far6[m] ? r : start];
					}
#pragma unroll
					for (int m = 0; m < 6; ++m) {
						if (!have[m]) continue;
						const uint32_t op = L * 12 + half * 6 + m;
						const uint32_t r = r6[m];
						uint8_t code = 0xff;
						if (r >= start) code = (uint8_t)(r - start);
						else if (pred_in_window && r == start - 1) code = 0xfe;
						else {
							s_opv[op * CPC + q] = v6[m];
							if (r >= done && r < minread) minread = r;
						}
						s_code[op * CPC + q] = code;
					}
				}
				// window reads of slow-path ranks and of HIST copies (lane L: ranks L, L + 4)
#pragma unroll
				for (int h = 0; h < SPEC_HB / 4; ++h) {
					const uint32_t k = L + 4 * h;
					if (k >= len) continue;
					const uint8_t kd = s_kind[k * CPC + q];
					if (kd == 1 && s_K[k * CPC + q] == 255) {
						const uint32_t b = s_coff[k * CPC + q], e2 = a.cand_off[start + k + 1];
						for (uint32_t w = 3 * b; w < 3 * e2; ++w) {
							const uint32_t r = a.cand[w];
							if (r < start && !(pred_in_window && r == start - 1) && r >= done && r < minread) minread = r;
						}
					} else if (kd == 2) {
						const uint32_t r = a.src[start + k];
						if (r < start && !(pred_in_window && r == start - 1) && r >= done && r < minread) minread = r;
					}
				}
			}
			__syncwarp();
			if (compute) {
				// candidates whose three operands are all outside the chunk do not depend on the
				// hypothesis: predict them once (lane L: candidates 4L .. 4L + 3)
#pragma unroll
				for (int m = 0; m < 4; ++m) {
					const uint32_t cj = L * 4 + m;
					const uint32_t k = cj / SPEC3_KCAP, j = cj % SPEC3_KCAP;
					if (k >= len || s_kind[k * CPC + q] != 1) continue;
					const uint32_t K = s_K[k * CPC + q];
					if (K == 255 || j >= K) continue;
					const uint32_t op0 = cj * 3;
					if (s_code[op0 * CPC + q] == 0xff && s_code[(op0 + 1) * CPC + q] == 0xff && s_code[(op0 + 2) * CPC + q] == 0xff) {
						const Rec v0 = s_opv[op0 * CPC + q], v1 = s_opv[(op0 + 1) * CPC + q], v2 = s_opv[(op0 + 2) * CPC + q];
						Rec pv = v0;
#pragma unroll
						for (int c = 0; c < NC; ++c) pv.c[c] = IntOps<T>::predict_hi(v0.c[c], v1.c[c], v2.c[c], hi[c]);
						s_opv[op0 * CPC + q] = pv;
						s_code[op0 * CPC + q] = 0xfd;
					}
				}
			}
			// the group's lowest other-window read
			minread = min(minread, __shfl_xor_sync(0xffffffffu, minread, 1));
			minread = min(minread, __shfl_xor_sync(0xffffffffu, minread, 2));
			__syncwarp();
			const long long tB = clock64();
			// ------------------------------------------------------------------ phase 1: lane e walks the chunk under hypothesis e
			uint32_t mybits = 0, myflags = 0;
			if (compute && L < 3) {
				const uint32_t e = L;
#pragma unroll 1
				for (uint32_t k = 0; k < len; ++k) {
					const uint32_t K = s_K[k * CPC + q];
					const uint32_t kd = s_kind[k * CPC + q];
					const Rec res = s_res[k * CPC + q];
					Rec out;
					if (kd == 1 && K != 255) {
						Acc sum[NC];
#pragma unroll
						for (int c = 0; c < NC; ++c) sum[c] = 0;
#pragma unroll
						for (int j = 0; j < SPEC3_KCAP; ++j) {
							if ((uint32_t)j >= K) continue;
							const uint32_t op0 = (k * SPEC3_KCAP + (uint32_t)j) * 3;
							const uint8_t code0 = s_code[op0 * CPC + q];
							if (code0 == 0xfd) {
								const Rec pv = s_opv[op0 * CPC + q];
#pragma unroll
								for (int c = 0; c < NC; ++c) sum[c] += (Acc)pv.c[c];
								continue;
							}
							Rec v[3];
#pragma unroll
							for (int o = 0; o < 3; ++o) {
								const uint8_t code = o == 0 ? code0 : s_code[(op0 + o) * CPC + q];
								if (code == 0xff) v[o] = s_opv[(op0 + o) * CPC + q];
								else if (code == 0xfe) {
									v[o] = pst;
#pragma unroll
									for (int c = 0; c < NC; ++c) v[o].c[c] = (T)(v[o].c[c] + (T)((int)e - 1));
								} else v[o] = TR(e, code);
							}
#pragma unroll
							for (int c = 0; c < NC; ++c) sum[c] += (Acc)IntOps<T>::predict_hi(v[0].c[c], v[1].c[c], v[2].c[c], hi[c]);
						}
						out = res;
#pragma unroll
						for (int c = 0; c < NC; ++c) {
							// (sum + (K >> 1)) / K for K in {0, 1, 2} (transform.h:90-91)
							const T pred = K == 2 ? (T)((sum[c] + 1) >> 1) : (K == 1 ? (T)sum[c] : (T)0);
							out.c[c] = IntOps<T>::dec_hi(res.c[c], pred, hi[c]);
						}
					} else if (kd != 0) {
						auto get_slow = [&](uint32_t r) -> Rec {
							if (r >= start) return TR(e, r - start);
							if (pred_in_window && r == start - 1) {
								Rec v = pst;
#pragma unroll
								for (int c = 0; c < NC; ++c) v.c[c] = (T)(v.c[c] + (T)((int)e - 1));
								return v;
							}
							return xs[r];
						};
						if (kd == 1) {
							const uint32_t b = s_coff[k * CPC + q];
							out = spec_step<T, NC, false>(a, get_slow, b, a.cand_off[start + k + 1] - b, res);
						} else {
							out = get_slow(a.src[start + k]);
						}
					} else {
						out = s_old[k * CPC + q];
					}
					TR(e, k) = out;
				}
				// my entries of the offset map (end value under hypothesis e relative to the stored end
				// value) and my change flags (which of my ranks differ from the stored ones), per component
				const Rec last = TR(e, len - 1), olast = s_old[(len - 1) * CPC + q];
#pragma unroll
				for (int j = 0; j < NC; ++j) {
					const long long d = (long long)last.c[j] - (long long)olast.c[j] + 1;
					mybits |= (uint32_t)((d >= 0 && d <= 2) ? d : 3) << (6 * j + 2 * e);
				}
				for (uint32_t k = 0; k < len; ++k) {
					const Rec v = TR(e, k), o = s_old[k * CPC + q];
#pragma unroll
					for (int j = 0; j < NC; ++j)
						if (v.c[j] != o.c[j]) myflags |= ((k + 1 == len) ? 2u : 1u) << (6 * j + 2 * e);
				}
			}
			uint32_t map = mybits, flg = myflags;
			map |= __shfl_xor_sync(0xffffffffu, map, 1);
			map |= __shfl_xor_sync(0xffffffffu, map, 2);
			flg |= __shfl_xor_sync(0xffffffffu, flg, 1);
			flg |= __shfl_xor_sync(0xffffffffu, flg, 2);
			if (!active) { map = SPEC_MAP_IDENTITY; flg = 0; }
			if (blocked) { map = 0x00ffffffu; flg = 0x80000000u; } // every offset unknown; bit 31: never valid
			if (L == 0) {
				sc->map[g] = map;
				sc->flg[g] = flg;
				sc->dep[g] = (active && minread != 0xffffffffu) ? (minread - done) / SPEC_HB : 0xffffffffu;
			}
			if (t == 0) s_first = 0xffffffffu;
			const long long tC = clock64();
			// ------------------------------------------------------------------ phase 2: resolve
			// One cluster barrier (release / acquire at cluster scope) publishes every chunk's record;
			// then every CTA redundantly derives the offsets, the change counts, the validity of every
			// chunk and the first invalid one -- identical results, no further exchange.
			cluster.sync(); // [B]
			uint32_t f0 = 0, f1 = 0, d0 = 0xffffffffu, d1 = 0xffffffffu;
			{
				const bool ld = 2 * t < nact;
				s_all[2 * t] = ld ? __ldcg(&sc->map[2 * t]) : SPEC_MAP_IDENTITY;
				s_all[2 * t + 1] = ld ? __ldcg(&sc->map[2 * t + 1]) : SPEC_MAP_IDENTITY;
				if (ld) { f0 = __ldcg(&sc->flg[2 * t]); f1 = __ldcg(&sc->flg[2 * t + 1]); d0 = __ldcg(&sc->dep[2 * t]); d1 = __ldcg(&sc->dep[2 * t + 1]); }
			}
			__syncthreads();
			block_scan_maps_x2(s_all, NC, s_warp, s_excl, nact);
			// per window chunk i (two per thread): true offsets -> selected change flags
			bool kn[2] = { true, true };
			uint32_t cntv[2] = { 0, 0 };
			uint8_t inn[2] = { 0, 0 };
#pragma unroll
			for (int h = 0; h < 2; ++h) {
				const uint32_t i = 2 * t + h;
				if (i >= nact) continue;
				const uint32_t bf = i == 0 ? SPEC_MAP_IDENTITY : s_all[i - 1];
				const uint32_t fl = h ? f1 : f0;
				uint32_t inner = 0, lastc = 0;
#pragma unroll
				for (int j = 0; j < NC; ++j) {
					const uint32_t ej = map_get(bf, j, 1u);
					if (ej == 3u) { kn[h] = false; continue; }
					const uint32_t two = (fl >> (6 * j + 2 * ej)) & 3u;
					inner |= two & 1u;
					lastc |= two >> 1;
				}
				if (kn[h]) { cntv[h] = inner + lastc; inn[h] = (uint8_t)inner; }
			}
			__syncthreads(); // all reads of s_all[i - 1] done before s_excl (scratch of the scan above) is reused
			s_excl[2 * t] = cntv[0];
			s_excl[2 * t + 1] = cntv[1];
			s_inner[2 * t] = inn[0];
			s_inner[2 * t + 1] = inn[1];
			__syncthreads();
			block_scan_u32_x2(s_excl, s_warp, nact);
#pragma unroll
			for (int h = 0; h < 2; ++h) {
				const uint32_t i = 2 * t + h;
				if (i >= nact) continue;
				bool valid = kn[h] && !((h ? f1 : f0) >> 31);
				const uint32_t dp = h ? d1 : d0;
				if (valid && dp != 0xffffffffu && i > 0) {
					// other window reads: chunks [dep, i-2] entirely unchanged, chunk i-1 unchanged except
					// (possibly) its last rank, which the offset hypothesis accounts for
					if (s_excl[i - 1] - s_excl[dp] != 0 || s_inner[i - 1]) valid = false;
				}
				if (!valid) atomicMin(&s_first, i);
			}
			__syncthreads();
			const uint32_t first_bad = s_first;
			// my chunk's true offsets
			uint32_t ein[NC];
			{
				const uint32_t before = g == 0 ? SPEC_MAP_IDENTITY : (g < nact ? s_all[g - 1] : SPEC_MAP_IDENTITY);
#pragma unroll
				for (int j = 0; j < NC; ++j) ein[j] = map_get(before, j, 1u);
			}
			const long long tD = clock64();
			// ------------------------------------------------------------------ phase 3: write (lane L: ranks L, L + 4)
			if (compute) {
				const bool sel = g < first_bad;
#pragma unroll
				for (int h = 0; h < SPEC_HB / 4; ++h) {
					const uint32_t k = L + 4 * h;
					if (k >= len) continue;
					const Rec o = s_old[k * CPC + q];
					Rec v = o;
#pragma unroll
					for (int j = 0; j < NC; ++j) v.c[j] = TR(sel ? ein[j] : 1u, k).c[j];
					if (!spec_equal<T, NC>(v, o)) x[start + k] = v;
				}
			}
			const unsigned long long wend = (unsigned long long)done + (unsigned long long)(first_bad == 0xffffffffu ? nact : first_bad) * SPEC_HB;
			newdone = wend < n ? (uint32_t)wend : n;
			const uint32_t adv = newdone - done;
			est = adv > est - est / 8 ? adv : est - est / 8;
			if (adv <= 2 * SPEC_HB) ++poor;
			else { poor = 0; Bp = 4 * SPEC_HB; }
			if (poor >= 3 || adv == 0) hyp = false;
			++hsweeps;
			hadv += adv;
			parity ^= 1;
			cluster.sync(); // [F] writes visible before the next sweep reads the stored state
			{ const long long tE = clock64(); cy0 += tB - tA; cy1 += tC - tB; cy2 += tD - tC; cy3 += tE - tD; }
		} else {
			// ------------------------------------------------------------------ plain sweep (CTA 0, one warp):
			// no contraction here; a long exact chunk 0 carries the progress, a few speculative chunks ride along
			if (rank == 0 && t == 0) sc->plain_first[pparity ^ 1] = 0xffffffffu; // for the next plain sweep
			const uint32_t B = Bp;
			const uint32_t K0 = a.cand_off[done + 1] - a.cand_off[done];
			const bool wide0 = a.kind[done] == 1 && K0 > SPEC3_WIDE;
			if (rank == 0 && t < 32 && wide0) {
				// rank `done` is a wide vertex (pole / huge fan): all its operands are final; the warp
				// sums the candidate predictions with a strided loop + shuffle reduction (integer sums
				// are order independent, so this equals the sequential result)
				const uint32_t c0 = a.cand_off[done];
				long long sum[NC];
#pragma unroll
				for (int c = 0; c < NC; ++c) sum[c] = 0;
				for (uint32_t k = t; k < K0; k += 32) {
					const uint32_t *tr = a.cand + 3 * (size_t)(c0 + k);
					const Rec v0 = x[tr[0]], v1 = x[tr[1]], v2 = x[tr[2]];
#pragma unroll
					for (int c = 0; c < NC; ++c) sum[c] += (long long)IntOps<T>::predict_hi(v0.c[c], v1.c[c], v2.c[c], hi[c]);
				}
#pragma unroll
				for (int c = 0; c < NC; ++c)
#pragma unroll
					for (int d = 16; d > 0; d >>= 1) sum[c] += __shfl_xor_sync(0xffffffffu, sum[c], d);
				if (t == 0) {
					const Rec res = resid[done];
					Rec out = res;
#pragma unroll
					for (int c = 0; c < NC; ++c) out.c[c] = IntOps<T>::dec_hi(res.c[c], (T)hb_divround_i64(sum[c], (int)K0), hi[c]);
					x[done] = out;
					atomicMin(&sc->plain_first[pparity], done + 1);
				}
			} else if (rank == 0 && t < 32) {
				const unsigned long long start64 = (unsigned long long)done + (unsigned long long)t * B;
				uint32_t fc = 0xffffffffu;
				if (start64 < n) {
					const uint32_t start = (uint32_t)start64;
					const uint32_t end = (n - start < B) ? n : start + B;
					uint32_t c0 = a.cand_off[start];
					uint32_t stop = end;
					auto get = [&](uint32_t r) -> Rec { return x[r]; };
					for (uint32_t i = start; i < end; ++i) {
						const uint32_t c1 = a.cand_off[i + 1];
						const int kd = a.kind[i];
						if (kd == 1 && c1 - c0 > SPEC3_WIDE) { // not recomputed here: nothing from here on is validated
							stop = i;
							if (fc == 0xffffffffu) fc = i;
							break;
						}
						if (kd) {
							const Rec old = x[i];
							const Rec nw = kd == 2 ? x[a.src[i]] : spec_step<T, NC, false>(a, get, c0, c1 - c0, resid[i]);
							if (!spec_equal<T, NC>(old, nw)) {
								x[i] = nw;
								if (fc == 0xffffffffu) fc = i;
							}
						}
						c0 = c1;
					}
					if (t == 0 && fc != 0xffffffffu) fc = stop; // chunk 0 is exact up to where it stopped
				}
				if (fc != 0xffffffffu) atomicMin(&sc->plain_first[pparity], fc);
			}
			cluster.sync();
			const uint32_t p = __ldcg(&sc->plain_first[pparity]);
			const unsigned long long wend = (unsigned long long)done + (unsigned long long)32 * B;
			newdone = p != 0xffffffffu ? p : (wend < n ? (uint32_t)wend : n);
			if (Bp < SPEC_B_MAX) Bp <<= 1;
			hyp = true;
			poor = 2;
			pparity ^= 1;
			cy3 += clock64() - tA;
		}
		done = newdone;
		++sweeps;
	}
	if (rank == 0 && t == 0 && a.stats) {
		a.stats[0] = sweeps; a.stats[1] = hsweeps; a.stats[2] = sweeps - hsweeps; a.stats[3] = hadv;
		a.stats[4] = (unsigned long long)cy0; a.stats[5] = (unsigned long long)cy1; a.stats[6] = (unsigned long long)cy2; a.stats[7] = (unsigned long long)cy3;
	}
}
