// hb_lists.cuh -- device helpers shared by the encode and decode kernels: element -> (entity,
// region, binding slot, row) lookup for the three target classes, and the candidate combiner
// (AbsAttrCoder::get_prediction, formats/hry/attrcode.h:182-208).
#pragma once
#include "hb_internal.cuh"

#include <float.h>

enum { CLS_FACE = HB_FACE, CLS_VTX = HB_VTX, CLS_CORNER = HB_CORNER };

// Elements of a target class in emission order:
//   VTX     i = traversal position, entity = ord_v[i]
//   FACE    i = face rank,          entity = face of ford_h[i]
//   CORNER  i = corner element,     entity = half-edge celem_h[i] (bindings are per half-edge)
struct ElemCtx {
	const uint4 *he;
	const uint32_t *ord_v, *ford_h, *celem_h;
	const uint16_t *vtx_regs, *face_regs;
	const uint32_t *bind;  // binding table of the class
	const int16_t *slot;   // [region * nlists + list] -> binding slot or -1
	uint32_t nb, nlists, n;
};

template <int CLS>
__device__ __forceinline__ bool elem_lookup(const ElemCtx &c, uint32_t i, int l, uint32_t &row, int &a, uint32_t &entity)
{
	uint16_t reg;
	if (CLS == CLS_VTX) {
		entity = c.ord_v[i];
		reg = c.vtx_regs[entity];
	} else if (CLS == CLS_FACE) {
		entity = c.he[c.ford_h[i]].w;
		reg = c.face_regs[entity];
	} else {
		entity = c.celem_h[i];
		reg = c.face_regs[c.he[entity].w];
	}
	a = c.slot[(uint32_t)reg * c.nlists + (uint32_t)l];
	if (a < 0) return false;
	row = c.bind[(size_t)entity * c.nb + (uint32_t)a];
	return true;
}

// get_prediction for one component.  `get(k)` returns the k-th candidate value (bit container) in
// fan order.  Integer storage types: mean in int64 (uint64 for ULONG) with round-half-up
// truncating division, then truncated to T.  Float: mean accumulated in double IN ORDER, rounded
// to float, and the candidate closest to it wins (strict <, so later candidates win ties).
template <typename Get>
__device__ __forceinline__ unsigned long long combine_candidates(int st, uint32_t K, Get &&get)
{
	if (K == 0) return 0;
	if (st == HB_FLOAT) {
		double sum = 0.0;
		for (uint32_t k = 0; k < K; ++k) sum = __dadd_rn(sum, (double)__uint_as_float((uint32_t)get(k)));
		const float avg = __double2float_rn(__ddiv_rn(sum, (double)(int)K));
		float res = FLT_MAX;
		for (uint32_t k = 0; k < K; ++k) res = hb_closest_step(res, __uint_as_float((uint32_t)get(k)), avg);
		return __float_as_uint(res);
	}
	if (st == HB_ULONG) {
		unsigned long long sum = 0;
		for (uint32_t k = 0; k < K; ++k) sum += get(k);
		const unsigned long long d = (unsigned long long)K;
		return (sum + (d >> 1)) / d;
	}
	unsigned long long sum = 0;
	for (uint32_t k = 0; k < K; ++k) sum += (unsigned long long)hb_bits_to_i64(get(k), st);
	unsigned long long out = (unsigned long long)hb_divround_i64((long long)sum, (int)K);
	const int size = hb_type_size(st);
	if (size < 8) out &= (1ull << (8 * size)) - 1ull;
	return out;
}
