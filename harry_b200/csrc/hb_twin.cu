// hb_twin.cu -- twin matching of a polygon soup on the device.
//
// Replaces the hash join of mesh::Builder (structs/conn.h:164-214: add_edge / face_end / set_org) that the
// reference's readers run while they parse faces (formats/ply/reader.cc:349-353, formats/obj/reader.rl): an
// unordered_map from the DIRECTED edge (a, b) to the half-edge that first carried it, probed with (b, a) by every
// later half-edge, matched entries erased.  The result depends on the file order of the half-edges only through
// the half-edges that share an UNDIRECTED edge {lo, hi}; seen in index order they drive a three-state machine
//     empty --h--> stored(h)
//     stored(s) --h, opposite direction (or lo == hi)--> empty, twin(s) = h, twin(h) = s      (fmerge + erase)
//     stored(s) --h, same direction--> stored(s), h stays a border                            (insert keeps the old entry)
// and whatever is still stored at the end is a border (twin = itself, Conn::add_face :87-91).
//
// Device formulation (no sort, three streaming passes over the faces + one scan):
//   k_twin_scatter<false>  every half-edge hashes its undirected edge into one of M buckets (M ~ ne / 4, power of two,
//                   the counters fit the L2) and counts it
//   k_scan_lookback exclusive scan of the counters (hb_conn.cu)
//   k_twin_scatter<true>   every half-edge takes a slot of its bucket and leaves a 16-byte entry {lo, hi | dir << 31, h, f}
//   k_twin_resolve  every half-edge reads its bucket (a few consecutive entries), collects the members of its
//                   group (same lo, hi).  Groups of one or two -- every edge of a manifold mesh -- are decided in
//                   place and the record of h is written by its own thread (coalesced).  Larger groups
//                   (non-manifold edges) are replayed in index order by their smallest member, which writes the
//                   records of all members.
// The slots inside a bucket are handed out by atomics, so their order differs from run to run; the result does
// not depend on it.
#include "hb_internal.cuh"

static constexpr int TW_THREADS = 256;

__device__ __forceinline__ unsigned long long tw_mix(unsigned long long x)
{
	x ^= x >> 33;
	x *= 0xff51afd7ed558ccdULL;
	x ^= x >> 33;
	x *= 0xc4ceb9fe1a85ec53ULL;
	x ^= x >> 33;
	return x;
}

// Two-level bucket index.  The high bits hash the CELL (lo >> 4, hi >> 4), the low TW_FINE bits hash the edge
// itself: the edges of one cell share a run of 2^TW_FINE consecutive buckets.  Where the file order of the faces
// follows the vertex numbering (every mesh a scanner or a modelling tool writes), consecutive half-edges fall
// into a handful of cells, so the counters, the entry writes of the fill pass and the bucket reads of the
// resolve pass of one warp touch a few sectors instead of one per half-edge (10M-vertex sphere: 6.1 -> 2.7 ms);
// with randomly numbered vertices every cell holds about one edge and this is a plain hash (6.1 ms either way).
// A cell has at most 256 distinct edges over 8 buckets, so a high-valence vertex cannot overload a bucket.
static constexpr int TW_CELL = 4, TW_FINE = 3;
__device__ __forceinline__ uint32_t tw_bucket(uint32_t lo, uint32_t hi, uint32_t mask)
{
	const unsigned long long fine = tw_mix(((unsigned long long)lo << 32) | hi);
	const unsigned long long coarse = tw_mix(((unsigned long long)(lo >> TW_CELL) << 32) | (hi >> TW_CELL));
	return (uint32_t)((coarse << TW_FINE) | (fine & ((1u << TW_FINE) - 1u))) & mask;
}

// One thread per face, its corners in batches: the loads of a batch (origins, counters / bucket bounds, entries)
// are independent of each other, so a thread keeps several of them in flight instead of one dependent chain per
// half-edge.  Measured on the 10M-vertex sphere (59 996 352 half-edges): the scatter passes gain from batches of 4
// (fill 0.76 -> 0.63 ms, randomly numbered vertices 2.59 -> 1.84 ms); the resolve pass loses -- 48 / 64 / 76
// registers at batch 1 / 2 / 4 cost more occupancy than the extra loads in flight give back (1.53 / 1.97 / 2.79 ms;
// randomly numbered vertices 2.96 ms at every batch size) -- so it walks its corners one by one.
static constexpr int TW_BATCH = 4, TW_RESOLVE_BATCH = 1;

// FILL = false: count the bucket sizes; FILL = true: take slots and write the entries
template <bool FILL>
__global__ void __launch_bounds__(TW_THREADS)
k_twin_scatter(const uint32_t *__restrict__ face_off, const uint32_t *__restrict__ org, uint32_t os, uint32_t nf, uint32_t nv,
               uint32_t ne, uint32_t *__restrict__ cursor, uint32_t mask, uint4 *__restrict__ entries, int *err)
{
	const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
	if (f >= nf) return;
	const uint32_t b = face_off[f], e = face_off[f + 1];
	if (e < b || e - b > 0xffffu || e > ne) {
		atomicExch(err, 1);
		return;
	}
	if (b == e) return;
	const uint32_t first = org[(size_t)b * os];
	uint32_t a = first;
	for (uint32_t h0 = b; h0 < e; h0 += TW_BATCH) {
		uint32_t v[TW_BATCH + 1];
		v[0] = a;
#pragma unroll
		for (int k = 0; k < TW_BATCH; ++k) {
			const uint32_t h = h0 + k;
			v[k + 1] = h >= e ? 0u : (h + 1 == e ? first : org[(size_t)(h + 1) * os]);
		}
		uint32_t slot[TW_BATCH];
		bool bad = false;
#pragma unroll
		for (int k = 0; k < TW_BATCH; ++k) {
			if (h0 + k >= e) continue;
			if (v[k] >= nv || v[k + 1] >= nv) bad = true;
		}
		if (bad) {
			atomicExch(err, 2);
			return;
		}
#pragma unroll
		for (int k = 0; k < TW_BATCH; ++k) {
			if (h0 + k >= e) continue;
			const uint32_t lo = min(v[k], v[k + 1]), hi = max(v[k], v[k + 1]);
			slot[k] = atomicAdd(&cursor[tw_bucket(lo, hi, mask)], 1u);
		}
		if (FILL) {
#pragma unroll
			for (int k = 0; k < TW_BATCH; ++k) {
				if (h0 + k >= e) continue;
				const uint32_t lo = min(v[k], v[k + 1]), hi = max(v[k], v[k + 1]);
				entries[slot[k]] = make_uint4(lo, hi | (v[k] > v[k + 1] ? 0x80000000u : 0u), h0 + k, f);
			}
		}
		a = v[TW_BATCH];
	}
}

__device__ __forceinline__ void tw_store(uint32_t *out, uint32_t h, uint32_t org, uint32_t tf, uint32_t te)
{
	out[3 * (size_t)h] = org;
	out[3 * (size_t)h + 1] = tf;
	out[3 * (size_t)h + 2] = te; // u16 local edge, pad bytes zero
}

// non-manifold edge (more than two half-edges on {lo, hi}): replay of the group in index order by its smallest
// member, which writes the records of all members
__device__ __noinline__ void tw_replay_group(const uint32_t *__restrict__ face_off, const uint4 *__restrict__ entries, uint32_t s0,
                                             uint32_t s1, uint32_t lo, uint32_t hi, uint32_t g, uint32_t *out)
{
	bool has = false, started = false;
	uint32_t cur = 0, st_h = 0, st_f = 0, st_dir = 0;
	for (uint32_t it = 0; it < g; ++it) {
		uint4 best = make_uint4(0, 0, 0xffffffffu, 0);
		bool found = false;
		for (uint32_t s = s0; s < s1; ++s) {
			const uint4 en = entries[s];
			if (en.x == lo && (en.y & 0x7fffffffu) == hi && (!started || en.z > cur) && (!found || en.z < best.z)) {
				best = en;
				found = true;
			}
		}
		started = true;
		cur = best.z;
		const uint32_t bd = best.y >> 31, borg = bd ? hi : lo;
		if (has && (st_dir != bd || lo == hi)) {
			tw_store(out, best.z, borg, st_f, st_h - face_off[st_f]);
			tw_store(out, st_h, st_dir ? hi : lo, best.w, best.z - face_off[best.w]);
			has = false;
		} else if (has) {
			tw_store(out, best.z, borg, best.w, best.z - face_off[best.w]);
		} else {
			has = true;
			st_h = best.z;
			st_f = best.w;
			st_dir = bd;
		}
	}
	if (has) tw_store(out, st_h, st_dir ? hi : lo, st_f, st_h - face_off[st_f]);
}

template <int B>
__global__ void __launch_bounds__(TW_THREADS)
k_twin_resolve(const uint32_t *__restrict__ face_off, const uint32_t *org, uint32_t os, uint32_t nf, uint32_t ne,
               const uint32_t *__restrict__ bucket_end, uint32_t mask, const uint4 *__restrict__ entries, uint32_t *out)
{
	const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
	if (f >= nf) return;
	const uint32_t b = face_off[f], e = face_off[f + 1];
	if (b >= e || e > ne || e - b > 0xffffu) return; // malformed faces were flagged by the count pass
	const uint32_t first = org[(size_t)b * os];
	uint32_t a = first;
	for (uint32_t h0 = b; h0 < e; h0 += B) {
		// per corner of the batch: size of its group, smallest member, the (last seen) other member
		uint32_t v[B + 1], s0[B], s1[B], g[B], minh[B], oh[B], of[B], odir = 0;
		v[0] = a;
#pragma unroll
		for (int k = 0; k < B; ++k) {
			const uint32_t h = h0 + k;
			v[k + 1] = h >= e ? 0u : (h + 1 == e ? first : org[(size_t)(h + 1) * os]);
		}
#pragma unroll
		for (int k = 0; k < B; ++k) {
			s0[k] = s1[k] = 0;
			if (h0 + k >= e) continue;
			const uint32_t bk = tw_bucket(min(v[k], v[k + 1]), max(v[k], v[k + 1]), mask);
			s0[k] = bk ? bucket_end[bk - 1] : 0u;
			s1[k] = bucket_end[bk];
		}
#pragma unroll
		for (int k = 0; k < B; ++k) {
			const uint32_t lo = min(v[k], v[k + 1]), hi = max(v[k], v[k + 1]), h = h0 + k;
			g[k] = 0;
			minh[k] = 0xffffffffu;
			oh[k] = h;
			of[k] = f;
			for (uint32_t s = s0[k]; s < s1[k]; ++s) {
				const uint4 en = entries[s];
				if (en.x == lo && (en.y & 0x7fffffffu) == hi) {
					++g[k];
					minh[k] = min(minh[k], en.z);
					if (en.z != h) {
						oh[k] = en.z;
						of[k] = en.w;
						odir = (odir & ~(1u << k)) | ((en.y >> 31) << k);
					}
				}
			}
		}
#pragma unroll
		for (int k = 0; k < B; ++k) {
			const uint32_t h = h0 + k;
			if (h >= e) continue;
			const uint32_t lo = min(v[k], v[k + 1]), hi = max(v[k], v[k + 1]), dir = v[k] > v[k + 1] ? 1u : 0u;
			if (g[k] <= 2) {
				uint32_t tf = f, te = h - b;
				if (g[k] == 2 && (((odir >> k) & 1u) != dir || lo == hi)) {
					tf = of[k];
					te = oh[k] - face_off[of[k]];
				}
				tw_store(out, h, v[k], tf, te);
			} else if (minh[k] == h) {
				const uint32_t bk = tw_bucket(lo, hi, mask);
				tw_replay_group(face_off, entries, bk ? bucket_end[bk - 1] : 0u, bucket_end[bk], lo, hi, g[k], out);
			}
		}
		a = v[B];
	}
}

// d_org: u32 every `os` words (1: packed array, 3: the org field of 12-byte edge records).  d_out: ne records of
// 12 bytes.  d_org may alias d_out (in-place operation on the reference's own records): the org word of a record
// is only ever rewritten with the value it already holds, and nobody reads the twin words.
int hb_twin_build(hb_dmesh *m, uint32_t nv, uint32_t nf, uint32_t ne, const uint32_t *d_face_off, const uint32_t *d_org,
                  uint32_t os, uint32_t *d_out)
{
	hb_ctx *ctx = m->ctx;
	if (!nf || !ne) return 0;
	if (nv > 0x80000000u) return hb_fail(ctx, HB_ERR_UNSUPPORTED, "twin_match: more than 2^31 vertices");
	uint32_t nb = 1024;
	while ((uint64_t)nb * 4 < ne) nb <<= 1;
	uint32_t *d_cursor = nullptr;
	uint4 *d_entries = nullptr;
	HB_TRY(hb_dalloc_t(m, &d_cursor, (size_t)nb + 1));
	HB_TRY(hb_dalloc_t(m, &d_entries, (size_t)ne));
	HB_CUDA(ctx, cudaMemsetAsync(d_cursor, 0, sizeof(uint32_t) * ((size_t)nb + 1), ctx->stream));
	const uint32_t grid = hb_div_up(nf, TW_THREADS);
	HB_LAUNCH(ctx, k_twin_scatter<false>, grid, TW_THREADS, 0, d_face_off, d_org, os, nf, nv, ne, d_cursor, nb - 1, d_entries, ctx->d_err);
	HB_TRY(hb_scan_exclusive_u32(ctx, d_cursor, d_cursor, nb, nullptr));
	// after the fill pass cursor[b] is the END of bucket b (its start is cursor[b - 1])
	HB_LAUNCH(ctx, k_twin_scatter<true>, grid, TW_THREADS, 0, d_face_off, d_org, os, nf, nv, ne, d_cursor, nb - 1, d_entries, ctx->d_err);
	static_assert(TW_RESOLVE_BATCH == 1, "the launch below names the instantiation (the profile report prints it)");
	HB_LAUNCH(ctx, k_twin_resolve<1>, grid, TW_THREADS, 0, d_face_off, d_org, os, nf, ne, d_cursor, nb - 1, d_entries, d_out);
	return 0;
}
