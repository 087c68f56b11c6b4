// hb_decode_spec2.cuh -- cluster version of the speculative vertex reconstruction (integer lists).
//
// Same algorithm and exactness argument as hb_decode_spec.cuh (hypothesis mode): Jacobi sweeps over
// a window of chunks of SPEC_HB consecutive ranks; three trajectories per chunk (predecessor value
// offset -1 / 0 / +1 per component); prefix composition of the per-chunk offset maps; a chunk is
// valid if its offset-in is known and every other window value it read is unchanged; valid chunks
// are final.  What changes is the mapping onto the machine:
//
//   * one thread-block CLUSTER (up to 8 CTAs = 8 SMs) per list.  Chunk g = cta_rank * NT + t.
//     The window state lives in HBM/L2 (x[]), the scans exchange one map / one count per CTA through
//     global scratch, ordered by barrier.cluster (cooperative_groups cluster.sync()).
//   * the stored state is read-only while trajectories are computed, so every global load of a
//     chunk -- candidate offsets, rank triples, residuals, and the operand values that lie outside
//     the chunk -- is issued up front into shared memory (two dependent load waves), and the
//     sequential part of the sweep is pure ALU + shared memory.
//   * ranks with more than SPEC2_KCAP candidates (poles, high-valence vertices) take a slow path
//     with direct loads.
#pragma once
#include <cooperative_groups.h>
#include <type_traits>
#include "hb_decode_spec.cuh"

namespace cg = cooperative_groups;

#define SPEC2_KCAP 2        // candidates per rank cached in shared memory
#define SPEC2_CLUSTER 8
#define SPEC2_OPS (SPEC_HB * SPEC2_KCAP * 3)

struct Spec2Scratch {           // global scratch of one list (double-buffered where needed)
	uint32_t cta_map[SPEC2_CLUSTER];
	uint32_t cta_cnt[SPEC2_CLUSTER];
	uint32_t first_bad[2];
	uint32_t plain_first[2];
	uint32_t pad[12];
};

template <typename T, int NC> __host__ __device__ constexpr int spec2_threads()
{
	return (int)sizeof(SpecRec<T, NC>) <= 4 ? 512 : ((int)sizeof(SpecRec<T, NC>) <= 8 ? 256 : 128);
}
// per thread: (4 trajectories + residuals) * SPEC_HB records, SPEC2_OPS operand records, SPEC2_OPS
// operand codes, SPEC_HB candidate counts, SPEC_HB kinds
template <typename T, int NC> __host__ __device__ constexpr size_t spec2_smem()
{
	return (size_t)spec2_threads<T, NC>() * ((5 * SPEC_HB + SPEC2_OPS) * sizeof(SpecRec<T, NC>) + SPEC2_OPS + 2 * SPEC_HB);
}

template <typename T, int NC>
__global__ void __launch_bounds__((spec2_threads<T, NC>()), 1) k_decode_vertex_spec2(const SpecArgs *__restrict__ args, Spec2Scratch *__restrict__ scratch_all,
                                                                                      uint32_t *__restrict__ g_excl_all, uint8_t *__restrict__ g_inner_all)
{
	typedef SpecRec<T, NC> Rec;
	typedef typename std::conditional<(sizeof(T) <= 2), uint32_t, unsigned long long>::type Acc;
	cg::cluster_group cluster = cg::this_cluster();
	const uint32_t C = cluster.num_blocks();
	const uint32_t rank = cluster.block_rank();
	const uint32_t list = blockIdx.x / C;
	const SpecArgs a = args[list];
	Spec2Scratch *sc = scratch_all + list;
	const uint32_t NT = blockDim.x;
	uint32_t *g_excl = g_excl_all + (size_t)list * SPEC2_CLUSTER * 512;
	uint8_t *g_inner = g_inner_all + (size_t)list * SPEC2_CLUSTER * 512;
	const Rec *__restrict__ resid = (const Rec *)a.resid;
	Rec *x = (Rec *)a.x;
	const Rec *xs = (const Rec *)a.x; // stored state (read-only during phases 0-2)
	const uint32_t n = a.n;
	const uint32_t t = threadIdx.x;
	const uint32_t g = rank * NT + t;  // global chunk index in the window

	extern __shared__ __align__(16) unsigned char s_dyn[];
	Rec *s_traj = (Rec *)s_dyn;                                    // [(slot * HB + k) * NT + t]: slots 0-2 hypotheses, 3 stored, 4 residual
	Rec *s_opv = s_traj + (size_t)5 * SPEC_HB * NT;                 // [op * NT + t]
	uint8_t *s_code = (uint8_t *)(s_opv + (size_t)SPEC2_OPS * NT);   // [op * NT + t]
	uint8_t *s_K = s_code + (size_t)SPEC2_OPS * NT;                  // [k * NT + t] candidates (255 = slow path)
	uint8_t *s_kind = s_K + (size_t)SPEC_HB * NT;                    // [k * NT + t]
	__shared__ uint32_t s_warp[32], s_incl[512], s_bcast[4];

	auto TR = [&](uint32_t slot, uint32_t k) -> Rec & { return s_traj[(slot * SPEC_HB + k) * NT + t]; };

	T hi[NC];
#pragma unroll
	for (int c = 0; c < NC; ++c) hi[c] = IntOps<T>::mask(a.bits[c]);
	uint32_t done = 0, Bp = 4 * SPEC_HB, est = 0;
	bool hyp = true;
	int poor = 0;
	uint32_t parity = 0, pparity = 0; // double-buffer selectors of first_bad / plain_first
	unsigned long long sweeps = 0, hsweeps = 0, hadv = 0;
	long long cy0 = 0, cy1 = 0, cy2 = 0, cy3 = 0;
	while (done < n) {
		const long long tA = clock64();
		uint32_t newdone;
		if (hyp) {
			// ------------------------------------------------------------------ phase 0: hoisted loads
			uint32_t nact = (5 * est / 2) / SPEC_HB + 64;
			if (nact > C * NT) nact = C * NT;
			const unsigned long long start64 = (unsigned long long)done + (unsigned long long)g * SPEC_HB;
			const bool active = start64 < n && g < nact;
			const uint32_t start = active ? (uint32_t)start64 : n;
			const uint32_t len = active ? ((n - start < SPEC_HB) ? n - start : SPEC_HB) : 0;
			const bool pred_in_window = start > done;
			uint32_t minread = 0xffffffffu;
			Rec pst = resid[0];
			if (active) {
				uint32_t coff[SPEC_HB + 1];
#pragma unroll
				for (int k = 0; k <= SPEC_HB; ++k) coff[k] = (uint32_t)k <= len ? a.cand_off[start + k] : 0;
				if (pred_in_window) pst = xs[start - 1];
#pragma unroll
				for (int k = 0; k < SPEC_HB; ++k) {
					if ((uint32_t)k >= len) continue;
					TR(3, k) = xs[start + k];
					TR(4, k) = resid[start + k];
					const uint8_t kd = a.kind[start + k];
					const uint32_t K = coff[k + 1] - coff[k];
					s_kind[k * NT + t] = kd;
					s_K[k * NT + t] = (uint8_t)(K > SPEC2_KCAP ? 255 : K);
				}
				// wave 1: all rank triples of the cached candidates in one batch of independent loads
				uint32_t tri[SPEC2_OPS];
#pragma unroll
				for (int k = 0; k < SPEC_HB; ++k) {
					const uint32_t K = (uint32_t)k < len ? coff[k + 1] - coff[k] : 0;
#pragma unroll
					for (int j = 0; j < SPEC2_KCAP; ++j)
#pragma unroll
						for (int o = 0; o < 3; ++o)
							tri[(k * SPEC2_KCAP + j) * 3 + o] = (K <= SPEC2_KCAP && (uint32_t)j < K) ? a.cand[3 * (size_t)(coff[k] + j) + o] : 0xffffffffu;
				}
				// wave 2: operand values.  A candidate whose three operands all lie outside the chunk
				// does not depend on the offset hypothesis: its prediction is computed here, once.
#pragma unroll 1
				for (uint32_t cj = 0; cj < SPEC_HB * SPEC2_KCAP; ++cj) {
					const uint32_t op0 = cj * 3;
					// tri[] is indexed dynamically here only through this select chain-free trick:
					uint32_t r3[3];
#pragma unroll
					for (int q = 0; q < SPEC2_OPS; ++q)
						if ((uint32_t)q / 3 == cj) r3[q % 3] = tri[q];
					if (r3[0] == 0xffffffffu) continue; // candidate absent
					bool far[3];
					Rec v[3];
#pragma unroll
					for (int o = 0; o < 3; ++o) {
						const uint32_t r = r3[o];
						far[o] = !(r >= start) && !(pred_in_window && r == start - 1);
						v[o] = xs[far[o] ? r : start];
						if (far[o] && r >= done && r < minread) minread = r;
					}
					if (far[0] && far[1] && far[2]) {
						Rec pv = v[0];
#pragma unroll
						for (int c = 0; c < NC; ++c) pv.c[c] = IntOps<T>::predict_hi(v[0].c[c], v[1].c[c], v[2].c[c], hi[c]);
						s_opv[op0 * NT + t] = pv;
						s_code[op0 * NT + t] = 0xfd; // precomputed candidate value
					} else {
#pragma unroll
						for (int o = 0; o < 3; ++o) {
							const uint32_t r = r3[o];
							uint8_t code = 0xff;
							if (r >= start) code = (uint8_t)(r - start);
							else if (pred_in_window && r == start - 1) code = 0xfe;
							else s_opv[(op0 + o) * NT + t] = v[o];
							s_code[(op0 + o) * NT + t] = code;
						}
					}
				}
				// window reads of the slow-path ranks and of HIST copies
				for (uint32_t k = 0; k < len; ++k) {
					const uint8_t kd = s_kind[k * NT + t];
					if (kd == 1 && s_K[k * NT + t] == 255) {
						const uint32_t b = a.cand_off[start + k], e2 = a.cand_off[start + k + 1];
						for (uint32_t q = 3 * b; q < 3 * e2; ++q) {
							const uint32_t r = a.cand[q];
							if (r < start && !(pred_in_window && r == start - 1) && r >= done && r < minread) minread = r;
						}
					} else if (kd == 2) {
						const uint32_t r = a.src[start + k];
						if (r < start && !(pred_in_window && r == start - 1) && r >= done && r < minread) minread = r;
					}
				}
			}
			const long long tB = clock64();
			// ------------------------------------------------------------------ phase 1: three trajectories (ALU + smem)
			uint32_t map = SPEC_MAP_IDENTITY;
			if (active) {
#pragma unroll 1
				for (uint32_t k = 0; k < len; ++k) {
					const uint32_t K = s_K[k * NT + t];
					const uint32_t kd = s_kind[k * NT + t];
					const Rec res = TR(4, k);
#pragma unroll 1
					for (uint32_t e = 0; e < 3; ++e) {
						Rec out;
						if (kd == 1 && K != 255) {
							Acc sum[NC];
#pragma unroll
							for (int c = 0; c < NC; ++c) sum[c] = 0;
#pragma unroll
							for (int j = 0; j < SPEC2_KCAP; ++j) {
								if ((uint32_t)j >= K) continue;
								const uint32_t op0 = (k * SPEC2_KCAP + (uint32_t)j) * 3;
								const uint8_t code0 = s_code[op0 * NT + t];
								if (code0 == 0xfd) {
									const Rec pv = s_opv[op0 * NT + t];
#pragma unroll
									for (int c = 0; c < NC; ++c) sum[c] += (Acc)pv.c[c];
									continue;
								}
								Rec v[3];
#pragma unroll
								for (int o = 0; o < 3; ++o) {
									const uint8_t code = o == 0 ? code0 : s_code[(op0 + o) * NT + t];
									if (code == 0xff) v[o] = s_opv[(op0 + o) * NT + t];
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
								const uint32_t b = a.cand_off[start + k];
								out = spec_step<T, NC, false>(a, get_slow, b, a.cand_off[start + k + 1] - b, res);
							} else {
								out = get_slow(a.src[start + k]);
							}
						} else {
							out = TR(3, k);
						}
						TR(e, k) = out;
					}
				}
				map = 0;
#pragma unroll
				for (int j = 0; j < NC; ++j)
#pragma unroll
					for (int e = 0; e < 3; ++e) {
						const long long d = (long long)TR(e, len - 1).c[j] - (long long)TR(3, len - 1).c[j] + 1;
						map |= (uint32_t)((d >= 0 && d <= 2) ? d : 3) << (6 * j + 2 * e);
					}
			}
			const long long tC = clock64();
			// ------------------------------------------------------------------ phase 2: resolve across the cluster
			// (barrier.cluster has release / acquire semantics at cluster scope: the global-memory
			//  exchanges below need no extra fence)
			const uint32_t incl = block_scan_maps(map, NC, s_warp);
			s_incl[t] = incl;
			if (t == NT - 1) sc->cta_map[rank] = incl;
			if (g == 0) sc->first_bad[parity ^ 1] = 0xffffffffu;
			cluster.sync(); // [B] CTA maps visible
			if (t == 0) {
				uint32_t pre = SPEC_MAP_IDENTITY;
				for (uint32_t r = 0; r < rank; ++r) pre = map_compose(pre, __ldcg(&sc->cta_map[r]), NC);
				s_bcast[0] = pre;
			}
			__syncthreads();
			const uint32_t local_before = t == 0 ? SPEC_MAP_IDENTITY : s_incl[t - 1];
			const uint32_t before = map_compose(s_bcast[0], local_before, NC);
			uint32_t ein[NC];
			bool known = true;
#pragma unroll
			for (int j = 0; j < NC; ++j) { ein[j] = map_get(before, j, 1u); known = known && ein[j] != 3u; }
			bool chg_inner = false, chg_last = false;
			if (active && known) {
				for (uint32_t k = 0; k < len; ++k) {
					bool ch = false;
#pragma unroll
					for (int j = 0; j < NC; ++j) ch = ch || TR(ein[j], k).c[j] != TR(3, k).c[j];
					if (k + 1 == len) chg_last = ch;
					else chg_inner = chg_inner || ch;
				}
			}
			const uint32_t cnt = (chg_inner ? 1u : 0u) + (chg_last ? 1u : 0u);
			const uint32_t lexcl = block_scan_u32(cnt, s_warp);
			// CTA-local exclusive counts + per-CTA totals; readers add the CTA prefix themselves
			g_excl[g] = lexcl;
			g_inner[g] = chg_inner ? 1 : 0;
			if (t == NT - 1) sc->cta_cnt[rank] = lexcl + cnt;
			cluster.sync(); // [C] counts visible
			bool valid = !active || known;
			if (active && known && minread != 0xffffffffu && g > 0) {
				const uint32_t dep = (minread - done) / SPEC_HB;
				auto gexcl = [&](uint32_t q) -> uint32_t { // global exclusive count of chunk q
					uint32_t pre = 0;
					const uint32_t qc = q / NT;
					for (uint32_t r = 0; r < qc; ++r) pre += __ldcg(&sc->cta_cnt[r]);
					return pre + __ldcg(&g_excl[q]);
				};
				const uint32_t changed_before = gexcl(g - 1) - gexcl(dep); // chunks [dep, g-2]
				if (changed_before != 0 || __ldcg(&g_inner[g - 1])) valid = false;
			}
			if (!valid) atomicMin(&sc->first_bad[parity], g);
			cluster.sync(); // [E] first invalid chunk known
			const uint32_t first_bad = __ldcg(&sc->first_bad[parity]);
			const long long tD = clock64();
			// ------------------------------------------------------------------ phase 3: write
			if (active) {
				const bool sel = g < first_bad;
				for (uint32_t k = 0; k < len; ++k) {
					const Rec o = TR(3, k);
					Rec v = o;
#pragma unroll
					for (int j = 0; j < NC; ++j) v.c[j] = TR(sel ? ein[j] : 1, k).c[j];
					if (!spec_equal<T, NC>(v, o)) x[start + k] = v;
				}
			}
			const unsigned long long wend = (unsigned long long)done + (unsigned long long)(first_bad == 0xffffffffu ? nact : first_bad) * SPEC_HB;
			newdone = wend < n ? (uint32_t)wend : n;
			const uint32_t adv = newdone - done;
			est = adv > est - est / 8 ? adv : est - est / 8;
			if (adv <= 2 * SPEC_HB) ++poor;
			else { poor = 0; Bp = 4 * SPEC_HB; }
			if (poor >= 3) hyp = false;
			++hsweeps;
			hadv += adv;
			parity ^= 1;
			cluster.sync(); // [F] writes visible before the next sweep reads the stored state
			{ const long long tE = clock64(); cy0 += tB - tA; cy1 += tC - tB; cy2 += tD - tC; cy3 += tE - tD; }
		} else {
			// ------------------------------------------------------------------ plain sweep (CTA 0, one warp):
			// no contraction here; a long exact chunk 0 carries the progress, a few speculative chunks ride along
			if (g == 0) sc->plain_first[pparity ^ 1] = 0xffffffffu; // for the next plain sweep
			const uint32_t B = Bp;
			if (rank == 0 && t < 32) {
				const unsigned long long start64 = (unsigned long long)done + (unsigned long long)t * B;
				uint32_t fc = 0xffffffffu;
				if (start64 < n) {
					const uint32_t start = (uint32_t)start64;
					const uint32_t end = (n - start < B) ? n : start + B;
					uint32_t c0 = a.cand_off[start];
					auto get = [&](uint32_t r) -> Rec { return x[r]; };
					for (uint32_t i = start; i < end; ++i) {
						const uint32_t c1 = a.cand_off[i + 1];
						const int kd = a.kind[i];
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
					if (t == 0 && fc != 0xffffffffu) fc = end; // chunk 0 is exact after the sweep
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
	if (g == 0 && a.stats) {
		a.stats[0] = sweeps; a.stats[1] = hsweeps; a.stats[2] = sweeps - hsweeps; a.stats[3] = hadv;
		a.stats[4] = (unsigned long long)cy0; a.stats[5] = (unsigned long long)cy1; a.stats[6] = (unsigned long long)cy2; a.stats[7] = (unsigned long long)cy3;
	}
}
