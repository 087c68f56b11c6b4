// hb_decode_spec.cuh -- speculative chunk-parallel reconstruction of vertex lists.
//
// Problem: x[i] = decodeDelta(residual[i], prediction(x[deps(i)])), deps(i) < i, and in a CBM
// traversal one dependency is (almost) always rank i-1, so the DAG is a chain of depth ~N.
//
// Scheme (exact for any input, see DESIGN.md "Speculative wavefront"):
//   * ranks [0, done) are final.  The window [done, done + T*B) is cut into T chunks of B
//     consecutive ranks; thread t recomputes chunk t SEQUENTIALLY, reading whatever values the
//     other chunks currently hold (possibly stale) and writing its results in place.
//   * each thread records the first rank of its chunk whose value CHANGED in this sweep; p = the
//     minimum over all chunks.  Every rank in [done, p) was recomputed from dependencies of lower
//     rank that did not change during the sweep, so by induction on the rank all of [done, p) equal
//     the sequential solution: done = p.
//   * chunk 0 starts at `done` and only reads final values, so it is exact after the sweep whether
//     or not it changed (if it changed, p = its end): progress >= B ranks per sweep, i.e. never
//     slower than a sequential walk, even if nothing else converges.
//   * why it is fast: with averaged parallelogram prediction x[i] ~ (x[i-1] + known terms) / 2 + delta,
//     an error in x[i-1] is halved at every step, so a chunk started from a wrong value is right
//     after ~log2(range) steps; whole rings of the traversal converge in 2-3 sweeps.
//   * B adapts: if a sweep advances by no more than one chunk (no contraction, e.g. lossless float
//     lists pick ONE candidate), B grows so that the exact chunk 0 carries the progress.
#pragma once
#include "hb_lists.cuh"

#define SPEC_THREADS 1024
#define SPEC_B_MIN 8
#define SPEC_B_MAX 1024

template <typename T, int NC> struct SpecRec;
template <typename T> struct alignas(sizeof(T) * 1) SpecRec<T, 1> { T c[1]; };
template <typename T> struct alignas(sizeof(T) * 2) SpecRec<T, 2> { T c[2]; };
template <typename T> struct alignas(sizeof(T) * 4) SpecRec<T, 3> { T c[4]; };
template <typename T> struct alignas(sizeof(T) * 4) SpecRec<T, 4> { T c[4]; };

struct SpecArgs {
	const uint8_t *kind;      // per rank: 0 skip, 1 DATA, 2 copy from src[i] (HIST)
	const uint32_t *src;      // kind 2: owning rank
	const uint32_t *cand_off; // n + 1
	const uint32_t *cand;     // rank triples
	const void *resid;        // compact records: residuals (read only)
	void *x;                  // compact records: values (in/out)
	uint32_t n;
	int bits[4];              // quantization bits per component (prediction.h:22-25)
	unsigned long long *stats; // [0] sweeps, [1] sequential steps (diagnostics)
};

// one reconstruction step for all components of a rank; FP = lossless float list (T == uint32_t bits)
template <typename T, int NC, bool FP>
__device__ __forceinline__ SpecRec<T, NC> spec_step(const SpecArgs &a, const SpecRec<T, NC> *x, uint32_t i, uint32_t c0, uint32_t K, const SpecRec<T, NC> &res)
{
	SpecRec<T, NC> out = res;
	const uint32_t *__restrict__ tri = a.cand + 3 * (size_t)c0;
	if (!FP) {
		long long sum[NC];
#pragma unroll
		for (int j = 0; j < NC; ++j) sum[j] = 0;
		for (uint32_t k = 0; k < K; ++k) {
			const SpecRec<T, NC> v0 = x[tri[3 * k]], v1 = x[tri[3 * k + 1]], v2 = x[tri[3 * k + 2]];
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
			const SpecRec<T, NC> v0 = x[tri[3 * k]], v1 = x[tri[3 * k + 1]], v2 = x[tri[3 * k + 2]];
#pragma unroll
			for (int j = 0; j < NC; ++j)
				sum[j] = __dadd_rn(sum[j], (double)__fadd_rn(__uint_as_float(v0.c[j]), __fsub_rn(__uint_as_float(v1.c[j]), __uint_as_float(v2.c[j]))));
		}
		float avg[NC], best[NC];
#pragma unroll
		for (int j = 0; j < NC; ++j) { avg[j] = K ? __double2float_rn(__ddiv_rn(sum[j], (double)(int)K)) : 0.f; best[j] = FLT_MAX; }
		for (uint32_t k = 0; k < K; ++k) {
			const SpecRec<T, NC> v0 = x[tri[3 * k]], v1 = x[tri[3 * k + 1]], v2 = x[tri[3 * k + 2]];
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

// one CTA per vertex list
template <typename T, int NC, bool FP>
__global__ void __launch_bounds__(SPEC_THREADS, 1) k_decode_vertex_spec(const SpecArgs *__restrict__ args)
{
	const SpecArgs a = args[blockIdx.x];
	typedef SpecRec<T, NC> Rec;
	const Rec *__restrict__ resid = (const Rec *)a.resid;
	Rec *x = (Rec *)a.x;
	const uint32_t n = a.n;
	__shared__ uint32_t s_first_changed;
	uint32_t done = 0, B = SPEC_B_MIN;
	unsigned long long sweeps = 0;
	while (done < n) {
		if (threadIdx.x == 0) s_first_changed = 0xffffffffu;
		__syncthreads();
		const unsigned long long start64 = (unsigned long long)done + (unsigned long long)threadIdx.x * B;
		uint32_t fc = 0xffffffffu;
		if (start64 < n) {
			const uint32_t start = (uint32_t)start64;
			const uint32_t end = (n - start < B) ? n : start + B;
			uint32_t c0 = a.cand_off[start];
			for (uint32_t i = start; i < end; ++i) {
				const uint32_t c1 = a.cand_off[i + 1];
				const int kind = a.kind[i];
				if (kind) {
					const Rec old = x[i];
					const Rec nw = kind == 2 ? x[a.src[i]] : spec_step<T, NC, FP>(a, (const Rec *)x, i, c0, c1 - c0, resid[i]);
					if (!spec_equal<T, NC>(old, nw)) {
						x[i] = nw;
						if (fc == 0xffffffffu) fc = i;
					}
				}
				c0 = c1;
			}
			// chunk 0 starts at `done` and reads only final values: it is exact after this sweep
			// whether or not it changed; if it changed, nothing behind it is validated
			if (threadIdx.x == 0 && fc != 0xffffffffu) fc = end;
		}
		if (fc != 0xffffffffu) atomicMin(&s_first_changed, fc);
		__syncthreads();
		const uint32_t p = s_first_changed;
		const unsigned long long wend = (unsigned long long)done + (unsigned long long)SPEC_THREADS * B;
		const uint32_t newdone = p != 0xffffffffu ? p : (wend < n ? (uint32_t)wend : n);
		const uint32_t adv = newdone - done;
		// adapt the chunk length: no speculation benefit -> longer exact chunk 0; plenty -> shorter
		if (adv <= 2 * B && B < SPEC_B_MAX) B <<= 1;
		else if (adv >= 64 * B && B > SPEC_B_MIN) B >>= 1;
		done = newdone;
		++sweeps;
		__syncthreads();
	}
	if (threadIdx.x == 0 && a.stats) { a.stats[0] = sweeps; }
}
