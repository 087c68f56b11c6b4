/*
 * harry_oracle.c -- CPU restatement of the Harry attribute path.  TEST INFRASTRUCTURE ONLY
 * (see harry_oracle.h for who may call it and for the parity-pinning status).
 *
 * Plain single-threaded C99.  Every function cites the reference code it restates
 * (paths relative to the reference tree).  Nothing here is copied from the reference: the
 * semantics (integer promotion, rounding, evaluation order) are re-derived and then pinned
 * against the real reference by tests/test_oracle_pinned.py.
 */
#include "harry_oracle.h"

#include <float.h>
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static char g_err[512];
const char *ho_last_error(void) { return g_err; }
#define FAIL(code, ...) do { snprintf(g_err, sizeof g_err, __VA_ARGS__); return (code); } while (0)

/* ------------------------------------------------------------------------------------------ */
/* Type model: structs/mixing.h:18-19 (SIZES / Type), :58,:101-108 (storage type of a quantized  */
/* component).                                                                                 */
/* ------------------------------------------------------------------------------------------ */
static const int TYPE_SIZE[11] = { 4, 8, 8, 8, 4, 4, 2, 2, 1, 1, 0 };

static int quant_storage_type(int q)
{
	if (q <= 8) return HB_UCHAR;
	if (q <= 16) return HB_USHORT;
	if (q <= 32) return HB_UINT;
	if (q <= 64) return HB_ULONG;
	return HB_TYPE_NONE;
}
static int comp_stype(const hb_list_desc *L, int j)
{
	return L->quant[j] ? quant_storage_type(L->quant[j]) : L->type[j];
}
static int type_is_signed(int t) { return t == HB_LONG || t == HB_INT || t == HB_SHORT || t == HB_CHAR; }

/* raw little-endian bits of one component, zero-extended into a u64 container */
static uint64_t ld_bits(const uint8_t *p, int t)
{
	uint64_t v = 0;
	memcpy(&v, p, (size_t)TYPE_SIZE[t]);
	return v;
}
static void st_bits(uint8_t *p, int t, uint64_t v) { memcpy(p, &v, (size_t)TYPE_SIZE[t]); }

/* value of an integer-typed container as int64 (sign- or zero-extended): View::get<int64_t>,
 * structs/mixing.h:245-270 */
static int64_t bits_to_i64(uint64_t b, int t)
{
	switch (t) {
	case HB_CHAR: return (int8_t)b;
	case HB_UCHAR: return (uint8_t)b;
	case HB_SHORT: return (int16_t)b;
	case HB_USHORT: return (uint16_t)b;
	case HB_INT: return (int32_t)b;
	case HB_UINT: return (uint32_t)b;
	default: return (int64_t)b;
	}
}
static float bits_to_f32(uint64_t b) { uint32_t u = (uint32_t)b; float f; memcpy(&f, &u, 4); return f; }
static uint64_t f32_to_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

/* ------------------------------------------------------------------------------------------ */
/* Scalar prediction / residual arithmetic: formats/hry/prediction.h:22-147.                   */
/* One instantiation per integer width/signedness, mirroring the reference's templates over T:  */
/* additions/subtractions wrap modulo 2^width (done in the unsigned twin type U to stay defined */
/* for 32/64-bit signed T), comparisons and right shifts keep T's signedness.                    */
/* ------------------------------------------------------------------------------------------ */
#define DEFINE_INT_OPS(T, U, N, WIDTH)                                                            \
	/* prediction.h:27-31 */                                                                      \
	static T N##_mask(int bits) { return bits == WIDTH ? (T)-1 : (T)((1u << bits) - 1u); }        \
	/* prediction.h:121-137: parallelogram v0 + v1 - v2 saturated to [0, mask] */                  \
	static T N##_predict(T v0, T v1, T v2, int bits)                                              \
	{                                                                                             \
		const T hi = N##_mask(bits);                                                              \
		if (v1 < v2) {                                                                            \
			const T d = (T)((U)v2 - (U)v1);                                                       \
			if (d > v0) return (T)0;                                                              \
			return (T)((U)v0 - (U)d);                                                             \
		} else {                                                                                  \
			const T d = (T)((U)v1 - (U)v2);                                                       \
			const T v = (T)((U)v0 + (U)d);                                                        \
			if (v > hi || v < v0) return hi;                                                      \
			return v;                                                                             \
		}                                                                                         \
	}                                                                                             \
	/* prediction.h:81-99: balanced zig-zag around pred, escape past min(pred, max - pred) */     \
	static T N##_enc(T raw, T pred, int bits)                                                     \
	{                                                                                             \
		const T room = (T)((U)N##_mask(bits) - (U)pred);                                          \
		if (pred == (T)0) return raw;                                                             \
		const T bal = room < pred ? room : pred;                                                  \
		if (raw < pred) {                                                                         \
			const T d = (T)((U)pred - (U)raw);                                                    \
			if (d > bal) return (T)((U)d + (U)bal);                                               \
			return (T)(((U)d << 1) - (U)1);                                                       \
		} else {                                                                                  \
			const T d = (T)((U)raw - (U)pred);                                                    \
			if (d > bal) return (T)((U)d + (U)bal);                                               \
			return (T)((U)d << 1);                                                                \
		}                                                                                         \
	}                                                                                             \
	/* prediction.h:46-63 */                                                                      \
	static T N##_dec(T delta, T pred, int bits)                                                   \
	{                                                                                             \
		const T room = (T)((U)N##_mask(bits) - (U)pred);                                          \
		if (pred == (T)0) return delta;                                                           \
		const T pm1 = (T)((U)pred - (U)1);                                                        \
		const T bal = room < pm1 ? room : pm1;                                                    \
		const T half = (T)(delta >> 1);                                                           \
		if (half > bal) {                                                                         \
			if (room >= pred) return (T)((U)pred + (U)delta - (U)bal - (U)1);                     \
			return (T)((U)pred - (U)delta + (U)bal);                                              \
		}                                                                                         \
		return (T)((U)pred + ((U)half ^ ((delta & 1) ? (U) ~(U)0 : (U)0)));                       \
	}

DEFINE_INT_OPS(uint8_t, uint8_t, u8, 8)
DEFINE_INT_OPS(int8_t, uint8_t, i8, 8)
DEFINE_INT_OPS(uint16_t, uint16_t, u16, 16)
DEFINE_INT_OPS(int16_t, uint16_t, i16, 16)
DEFINE_INT_OPS(uint32_t, uint32_t, u32, 32)
DEFINE_INT_OPS(int32_t, uint32_t, i32, 32)
DEFINE_INT_OPS(uint64_t, uint64_t, u64, 64)
DEFINE_INT_OPS(int64_t, uint64_t, i64, 64)

/* transform.h:19-23: order-preserving map of IEEE-754 bits (negative floats get their low 31
 * bits inverted); an involution.  The int2uint/uint2int XOR with masks[sizeof(T)] is the
 * identity for float because masks[4] == 0 (prediction.h:33-44). */
static uint32_t flip_f32(uint32_t i) { return i ^ ((0u - (i >> 31)) >> 1); }

/* prediction.h:22-25 */
static int stype_bits(int stype, int q) { return q == 0 ? TYPE_SIZE[stype] * 8 : q; }

uint64_t ho_predict(int stype, uint64_t v0, uint64_t v1, uint64_t v2, int q)
{
	const int bits = stype_bits(stype, q);
	switch (stype) {
	case HB_FLOAT: /* prediction.h:138-142 */
		return f32_to_bits(bits_to_f32(v0) + (bits_to_f32(v1) - bits_to_f32(v2)));
	case HB_UCHAR: return u8_predict((uint8_t)v0, (uint8_t)v1, (uint8_t)v2, bits);
	case HB_CHAR: return (uint8_t)i8_predict((int8_t)v0, (int8_t)v1, (int8_t)v2, bits);
	case HB_USHORT: return u16_predict((uint16_t)v0, (uint16_t)v1, (uint16_t)v2, bits);
	case HB_SHORT: return (uint16_t)i16_predict((int16_t)v0, (int16_t)v1, (int16_t)v2, bits);
	case HB_UINT: return u32_predict((uint32_t)v0, (uint32_t)v1, (uint32_t)v2, bits);
	case HB_INT: return (uint32_t)i32_predict((int32_t)v0, (int32_t)v1, (int32_t)v2, bits);
	case HB_ULONG: return u64_predict(v0, v1, v2, bits);
	case HB_LONG: return (uint64_t)i64_predict((int64_t)v0, (int64_t)v1, (int64_t)v2, bits);
	default: return 0;
	}
}

uint64_t ho_encode_delta(int stype, uint64_t raw, uint64_t pred, int q)
{
	const int bits = stype_bits(stype, q);
	switch (stype) {
	case HB_FLOAT: /* prediction.h:101-114 */
		return u32_enc(flip_f32((uint32_t)raw), flip_f32((uint32_t)pred), bits);
	case HB_UCHAR: return u8_enc((uint8_t)raw, (uint8_t)pred, bits);
	case HB_CHAR: return (uint8_t)i8_enc((int8_t)raw, (int8_t)pred, bits);
	case HB_USHORT: return u16_enc((uint16_t)raw, (uint16_t)pred, bits);
	case HB_SHORT: return (uint16_t)i16_enc((int16_t)raw, (int16_t)pred, bits);
	case HB_UINT: return u32_enc((uint32_t)raw, (uint32_t)pred, bits);
	case HB_INT: return (uint32_t)i32_enc((int32_t)raw, (int32_t)pred, bits);
	case HB_ULONG: return u64_enc(raw, pred, bits);
	case HB_LONG: return (uint64_t)i64_enc((int64_t)raw, (int64_t)pred, bits);
	default: return 0;
	}
}

uint64_t ho_decode_delta(int stype, uint64_t delta, uint64_t pred, int q)
{
	const int bits = stype_bits(stype, q);
	switch (stype) {
	case HB_FLOAT: /* prediction.h:64-72 */
		return flip_f32(u32_dec((uint32_t)delta, flip_f32((uint32_t)pred), bits));
	case HB_UCHAR: return u8_dec((uint8_t)delta, (uint8_t)pred, bits);
	case HB_CHAR: return (uint8_t)i8_dec((int8_t)delta, (int8_t)pred, bits);
	case HB_USHORT: return u16_dec((uint16_t)delta, (uint16_t)pred, bits);
	case HB_SHORT: return (uint16_t)i16_dec((int16_t)delta, (int16_t)pred, bits);
	case HB_UINT: return u32_dec((uint32_t)delta, (uint32_t)pred, bits);
	case HB_INT: return (uint32_t)i32_dec((int32_t)delta, (int32_t)pred, bits);
	case HB_ULONG: return u64_dec(delta, pred, bits);
	case HB_LONG: return (uint64_t)i64_dec((int64_t)delta, (int64_t)pred, bits);
	default: return 0;
	}
}

/* arith/msb.h:7-13: isolate the most significant set bit */
uint32_t ho_msb(uint32_t x)
{
	for (int s = 1; s < 32; s <<= 1) x |= x >> s;
	return x & ~(x >> 1);
}

/* ------------------------------------------------------------------------------------------ */
/* Quantization: structs/quant.h                                                               */
/* ------------------------------------------------------------------------------------------ */
static int check_list(const hb_list_desc *L)
{
	if (!L || L->ncomp > HB_MAX_COMP) FAIL(HB_ERR_INVALID, "list: bad component count");
	if (L->nrows && !L->rows && L->ncomp) FAIL(HB_ERR_INVALID, "list: rows == NULL");
	for (int j = 0; j < L->ncomp; ++j) {
		if (L->type[j] >= HB_TYPE_NONE) FAIL(HB_ERR_INVALID, "list: bad type");
		if (L->quant[j] > 8 * TYPE_SIZE[L->type[j]]) FAIL(HB_ERR_INVALID, "list: quant wider than slot");
		if ((uint32_t)L->offset[j] + (uint32_t)TYPE_SIZE[L->type[j]] > L->stride) FAIL(HB_ERR_INVALID, "list: offset out of row");
	}
	return 0;
}

/* quant.h:30-38.  min seeds with numeric_limits<T>::max(), max seeds with
 * numeric_limits<T>::min() -- for floating point that is the smallest positive normal, not
 * lowest().  Updates are `e < cur ? e : cur` / `e > cur ? e : cur`. */
int ho_bounds(const hb_list_desc *L, void *min_row, void *max_row)
{
	int rc = check_list(L);
	if (rc) return rc;
	const uint8_t *rows = (const uint8_t *)L->rows;
	for (int j = 0; j < L->ncomp; ++j) {
		const int t = L->type[j];
		const size_t off = L->offset[j];
		uint8_t *pmin = (uint8_t *)min_row + off, *pmax = (uint8_t *)max_row + off;
		if (t == HB_FLOAT) {
			float mn = FLT_MAX, mx = FLT_MIN;
			for (uint32_t i = 0; i < L->nrows; ++i) {
				float e;
				memcpy(&e, rows + (size_t)i * L->stride + off, 4);
				mn = e < mn ? e : mn;
				mx = e > mx ? e : mx;
			}
			memcpy(pmin, &mn, 4);
			memcpy(pmax, &mx, 4);
		} else if (t == HB_DOUBLE) {
			double mn = DBL_MAX, mx = DBL_MIN;
			for (uint32_t i = 0; i < L->nrows; ++i) {
				double e;
				memcpy(&e, rows + (size_t)i * L->stride + off, 8);
				mn = e < mn ? e : mn;
				mx = e > mx ? e : mx;
			}
			memcpy(pmin, &mn, 8);
			memcpy(pmax, &mx, 8);
		} else if (t == HB_ULONG) {
			uint64_t mn = UINT64_MAX, mx = 0;
			for (uint32_t i = 0; i < L->nrows; ++i) {
				uint64_t e = ld_bits(rows + (size_t)i * L->stride + off, t);
				mn = e < mn ? e : mn;
				mx = e > mx ? e : mx;
			}
			st_bits(pmin, t, mn);
			st_bits(pmax, t, mx);
		} else {
			int64_t mn, mx;
			switch (t) {
			case HB_LONG: mn = INT64_MAX; mx = INT64_MIN; break;
			case HB_UINT: mn = UINT32_MAX; mx = 0; break;
			case HB_INT: mn = INT32_MAX; mx = INT32_MIN; break;
			case HB_USHORT: mn = UINT16_MAX; mx = 0; break;
			case HB_SHORT: mn = INT16_MAX; mx = INT16_MIN; break;
			case HB_UCHAR: mn = UINT8_MAX; mx = 0; break;
			default: mn = INT8_MAX; mx = INT8_MIN; break;
			}
			for (uint32_t i = 0; i < L->nrows; ++i) {
				int64_t e = bits_to_i64(ld_bits(rows + (size_t)i * L->stride + off, t), t);
				mn = e < mn ? e : mn;
				mx = e > mx ? e : mx;
			}
			st_bits(pmin, t, (uint64_t)mn);
			st_bits(pmax, t, (uint64_t)mx);
		}
	}
	return 0;
}

/* quant.h:46-96.  range = max - min in the component's own type; the group leader's scale is the
 * maximum of numeric_limits<T>::min() and every member's range; then broadcast to the members.
 * Restated for groups whose members share one type (all reference readers produce such groups). */
int ho_scale(const hb_list_desc *L, const uint8_t *groups, const void *min_row, const void *max_row,
             void *scale_row)
{
	int rc = check_list(L);
	if (rc) return rc;
	for (int j = 0; j < L->ncomp; ++j)
		if (groups[j] >= L->ncomp || L->type[groups[j]] != L->type[j])
			FAIL(HB_ERR_UNSUPPORTED, "scale: mixed-type interpretation group");
	for (int k = 0; k < L->ncomp; ++k) {
		if (groups[k] != k) continue; /* k is a leader */
		const int t = L->type[k];
		if (t == HB_FLOAT) {
			float s = FLT_MIN;
			for (int j = 0; j < L->ncomp; ++j) {
				if (groups[j] != k) continue;
				float mn, mx;
				memcpy(&mn, (const uint8_t *)min_row + L->offset[j], 4);
				memcpy(&mx, (const uint8_t *)max_row + L->offset[j], 4);
				const float range = mx - mn;
				s = s < range ? range : s;
			}
			for (int j = 0; j < L->ncomp; ++j)
				if (groups[j] == k) memcpy((uint8_t *)scale_row + L->offset[j], &s, 4);
		} else if (t == HB_DOUBLE) {
			double s = DBL_MIN;
			for (int j = 0; j < L->ncomp; ++j) {
				if (groups[j] != k) continue;
				double mn, mx;
				memcpy(&mn, (const uint8_t *)min_row + L->offset[j], 8);
				memcpy(&mx, (const uint8_t *)max_row + L->offset[j], 8);
				const double range = mx - mn;
				s = s < range ? range : s;
			}
			for (int j = 0; j < L->ncomp; ++j)
				if (groups[j] == k) memcpy((uint8_t *)scale_row + L->offset[j], &s, 8);
		} else {
			/* integer types: seed = numeric_limits<T>::min() (0 for unsigned, lowest for signed) */
			const int sg = type_is_signed(t);
			int64_t s_i = 0;
			uint64_t s_u = 0;
			switch (t) {
			case HB_LONG: s_i = INT64_MIN; break;
			case HB_INT: s_i = INT32_MIN; break;
			case HB_SHORT: s_i = INT16_MIN; break;
			case HB_CHAR: s_i = INT8_MIN; break;
			default: break;
			}
			for (int j = 0; j < L->ncomp; ++j) {
				if (groups[j] != k) continue;
				const uint64_t mn = ld_bits((const uint8_t *)min_row + L->offset[j], t);
				const uint64_t mx = ld_bits((const uint8_t *)max_row + L->offset[j], t);
				uint64_t range = mx - mn; /* wraps in T after truncation below */
				if (TYPE_SIZE[t] < 8) range &= (1ull << (8 * TYPE_SIZE[t])) - 1ull;
				if (sg) {
					const int64_t r = bits_to_i64(range, t);
					s_i = s_i < r ? r : s_i;
				} else {
					s_u = s_u < range ? range : s_u;
				}
			}
			for (int j = 0; j < L->ncomp; ++j)
				if (groups[j] == k) st_bits((uint8_t *)scale_row + L->offset[j], t, sg ? (uint64_t)s_i : s_u);
		}
	}
	return 0;
}

/* quant.h:98-112, integer flavour: val / from * to + val % from * to / from, evaluated in T.
 * For T narrower than int the arithmetic happens in int and the result is truncated to T. */
static uint64_t rescale_int(int t, uint64_t val, uint64_t from, uint64_t to)
{
	switch (t) {
	case HB_ULONG: return val / from * to + val % from * to / from;
	case HB_LONG: { int64_t v = (int64_t)val, f = (int64_t)from, o = (int64_t)to; return (uint64_t)(v / f * o + v % f * o / f); }
	case HB_UINT: { uint32_t v = (uint32_t)val, f = (uint32_t)from, o = (uint32_t)to; return (uint32_t)(v / f * o + v % f * o / f); }
	case HB_INT: { int32_t v = (int32_t)val, f = (int32_t)from, o = (int32_t)to; return (uint32_t)(int32_t)((uint32_t)(v / f) * (uint32_t)o + (uint32_t)((int32_t)((uint32_t)(v % f) * (uint32_t)o) / f)); }
	case HB_USHORT: { int v = (uint16_t)val, f = (uint16_t)from, o = (uint16_t)to; return (uint16_t)(v / f * o + v % f * o / f); }
	case HB_SHORT: { int v = (int16_t)val, f = (int16_t)from, o = (int16_t)to; return (uint16_t)(int16_t)(v / f * o + v % f * o / f); }
	case HB_UCHAR: { int v = (uint8_t)val, f = (uint8_t)from, o = (uint8_t)to; return (uint8_t)(v / f * o + v % f * o / f); }
	default: { int v = (int8_t)val, f = (int8_t)from, o = (int8_t)to; return (uint8_t)(int8_t)(v / f * o + v % f * o / f); }
	}
}

/* quant.h:114-221: one list, all rows, in place. */
int ho_requant(hb_list_desc *L, const uint8_t *new_quant, const void *min_row, const void *scale_row)
{
	int rc = check_list(L);
	if (rc) return rc;
	uint8_t *rows = (uint8_t *)L->rows;
	for (int j = 0; j < L->ncomp; ++j) {
		const int t = L->type[j];
		const int sq = L->quant[j], dq = new_quant[j];
		if (dq > 8 * TYPE_SIZE[t]) FAIL(HB_ERR_INVALID, "requant: quant wider than slot");
		if (sq == 0 && dq == 0) continue; /* quant.h:118-121: bytewise copy onto itself */
		if (t == HB_DOUBLE) FAIL(HB_ERR_UNSUPPORTED, "requant: double lists (reference shifts an int by >= 32, undefined)");
		if (sq > 31 || dq > 31) FAIL(HB_ERR_UNSUPPORTED, "requant: > 31 bits (reference computes 1 << q in int, undefined)");
		const int st_src = sq ? quant_storage_type(sq) : t;
		const int st_dst = dq ? quant_storage_type(dq) : t;
		const uint64_t m_src = sq ? (uint64_t)(int64_t)(int32_t)((1u << sq) - 1u) : 0;
		const uint64_t m_dst = dq ? (uint64_t)(int64_t)(int32_t)((1u << dq) - 1u) : 0;
		const size_t off = L->offset[j];
		const uint64_t mn_b = ld_bits((const uint8_t *)min_row + off, t);
		const uint64_t sc_b = ld_bits((const uint8_t *)scale_row + off, t);
		for (uint32_t i = 0; i < L->nrows; ++i) {
			uint8_t *p = rows + (size_t)i * L->stride + off;
			uint64_t q;
			if (sq) {
				q = ld_bits(p, st_src); /* quant.h:124-131 */
			} else if (t == HB_FLOAT) {
				/* quant.h:135: four separately rounded float ops, then a truncating conversion */
				const float x = bits_to_f32(ld_bits(p, t));
				volatile float a = x - bits_to_f32(mn_b);
				volatile float b = a / bits_to_f32(sc_b);
				volatile float c = b * (float)(int32_t)m_dst;
				volatile float d = c + 0.5f;
				q = (uint64_t)d;
			} else {
				/* quant.h:140-163: (x - min) in T, then the integer rescale in T */
				uint64_t v = ld_bits(p, t) - mn_b;
				q = rescale_int(t, v, sc_b, m_dst);
				if (type_is_signed(t)) q = (uint64_t)bits_to_i64(q, t);
			}
			if (sq && dq) q = q / m_src * m_dst + q % m_src * m_dst / m_src; /* quant.h:167-169 */
			if (dq) {
				st_bits(p, st_dst, q); /* quant.h:171-178: other bytes of the slot stay */
			} else if (t == HB_FLOAT) {
				/* quant.h:182 */
				volatile float a = (float)q / (float)(int32_t)m_src;
				volatile float b = a * bits_to_f32(sc_b);
				volatile float c = b + bits_to_f32(mn_b);
				st_bits(p, t, f32_to_bits(c));
			} else {
				/* quant.h:187-209: rescale<T>(q, M, scale) + min, in T */
				uint64_t r = rescale_int(t, q, m_src, sc_b) + mn_b;
				st_bits(p, t, r);
			}
		}
		L->quant[j] = (uint8_t)dq;
	}
	return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Attribute coder                                                                             */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
	uint32_t *v;
	size_t n, cap;
} u32vec;
static int u32vec_push(u32vec *a, uint32_t x)
{
	if (a->n == a->cap) {
		size_t nc = a->cap ? a->cap * 2 : 16;
		uint32_t *p = (uint32_t *)realloc(a->v, nc * sizeof(uint32_t));
		if (!p) return -1;
		a->v = p;
		a->cap = nc;
	}
	a->v[a->n++] = x;
	return 0;
}
typedef struct {
	uint8_t *v;
	size_t n, cap;
} u8vec;
static int u8vec_append(u8vec *a, const uint8_t *src, size_t len)
{
	if (a->n + len > a->cap) {
		size_t nc = a->cap ? a->cap * 2 : 256;
		while (nc < a->n + len) nc *= 2;
		uint8_t *p = (uint8_t *)realloc(a->v, nc);
		if (!p) return -1;
		a->v = p;
		a->cap = nc;
	}
	if (len) memcpy(a->v + a->n, src, len);
	a->n += len;
	return 0;
}

typedef struct {
	int stype[HB_MAX_COMP], size[HB_MAX_COMP], q[HB_MAX_COMP];
	uint32_t sym_off[HB_MAX_COMP]; /* byte position of the component inside a residual row */
	uint32_t sym_stride;
} list_info;

typedef struct {
	const hb_mesh_desc *m;
	uint32_t *org, *tw, *hface; /* per half-edge: origin vertex, twin half-edge, face */
	uint8_t *vtx_done, *face_done; /* AbsAttrCoder::vtx_is_encoded / face_is_encoded, attrcode.h:109-110 */
	list_info *li;
	u32vec cand; /* vertex: triples (v0, v1, vo); corner: half-edge indices */
} coder;

static void coder_free(coder *c)
{
	free(c->org); free(c->tw); free(c->hface); free(c->vtx_done); free(c->face_done); free(c->li); free(c->cand.v);
}

static uint32_t he_next(const coder *c, uint32_t h)
{
	const uint32_t f = c->hface[h];
	return h + 1 == c->m->face_off[f + 1] ? c->m->face_off[f] : h + 1; /* conn.h:63-66,137-140 */
}
static uint32_t he_prev(const coder *c, uint32_t h)
{
	const uint32_t f = c->hface[h];
	return h == c->m->face_off[f] ? c->m->face_off[f + 1] - 1 : h - 1; /* conn.h:67-70,141-144 */
}

static int coder_init(coder *c, const hb_mesh_desc *m)
{
	memset(c, 0, sizeof *c);
	c->m = m;
	if (m->ne && (!m->edges || !m->face_off)) FAIL(HB_ERR_INVALID, "mesh: edges/face_off == NULL");
	if (m->face_off && m->face_off[m->nf] != m->ne) FAIL(HB_ERR_INVALID, "mesh: face_off[nf] != ne");
	c->org = (uint32_t *)malloc(sizeof(uint32_t) * (m->ne + 1));
	c->tw = (uint32_t *)malloc(sizeof(uint32_t) * (m->ne + 1));
	c->hface = (uint32_t *)malloc(sizeof(uint32_t) * (m->ne + 1));
	c->vtx_done = (uint8_t *)calloc(m->nv + 1, 1);
	c->face_done = (uint8_t *)calloc(m->nf + 1, 1);
	c->li = (list_info *)calloc(m->nlists + 1, sizeof(list_info));
	if (!c->org || !c->tw || !c->hface || !c->vtx_done || !c->face_done || !c->li) FAIL(HB_ERR_NOMEM, "out of memory");
	const uint8_t *e = (const uint8_t *)m->edges;
	for (uint32_t f = 0; f < m->nf; ++f)
		for (uint32_t h = m->face_off[f]; h < m->face_off[f + 1]; ++h) c->hface[h] = f;
	for (uint32_t h = 0; h < m->ne; ++h) {
		uint32_t o, tf;
		uint16_t te;
		memcpy(&o, e + 12 * (size_t)h, 4);
		memcpy(&tf, e + 12 * (size_t)h + 4, 4);
		memcpy(&te, e + 12 * (size_t)h + 8, 2);
		if (o >= m->nv) FAIL(HB_ERR_INVALID, "mesh: origin out of range");
		if (tf >= m->nf || m->face_off[tf] + te >= m->face_off[tf + 1]) FAIL(HB_ERR_INVALID, "mesh: twin out of range");
		c->org[h] = o;
		c->tw[h] = m->face_off[tf] + te; /* conn.h:123-126 */
	}
	for (int l = 0; l < m->nlists; ++l) {
		const hb_list_desc *L = &m->lists[l];
		int rc = check_list(L);
		if (rc) return rc;
		uint32_t pos = 0;
		for (int j = 0; j < L->ncomp; ++j) {
			const int st = comp_stype(L, j);
			if (st == HB_DOUBLE) FAIL(HB_ERR_UNSUPPORTED, "double lists: the reference reads masks[8] out of bounds (prediction.h:33-44)");
			if (L->quant[j] > 31 && L->quant[j] != 8 * TYPE_SIZE[st]) FAIL(HB_ERR_UNSUPPORTED, "quantization > 31 bits");
			c->li[l].stype[j] = st;
			c->li[l].size[j] = TYPE_SIZE[st];
			c->li[l].q[j] = L->quant[j];
			c->li[l].sym_off[j] = pos;
			pos += (uint32_t)TYPE_SIZE[st];
		}
		c->li[l].sym_stride = pos;
	}
	return 0;
}

/* Fan walk, attrcode.h:83-106 (TFAN_IT): forward over twin/next until the start edge or a border
 * (twin == self); after a border, backward from prev(start) over twin/prev.  Every visited
 * half-edge has the fan's centre vertex as origin.  Visited half-edges are appended to `out`. */
static int fan_walk(const coder *c, uint32_t ein, u32vec *out)
{
	const uint32_t cap = c->m->ne + 2;
	uint32_t steps = 0;
	uint32_t e = ein, t;
	for (;;) {
		if (u32vec_push(out, e)) FAIL(HB_ERR_NOMEM, "out of memory");
		t = c->tw[e];
		if (t == e) break; /* border: go backward */
		e = he_next(c, t);
		if (e == ein) return 0;
		if (++steps > cap) FAIL(HB_ERR_INVALID, "fan walk does not terminate (inconsistent twin table)");
	}
	e = he_prev(c, ein);
	t = c->tw[e];
	if (t == e) return 0;
	e = t;
	do {
		if (u32vec_push(out, e)) FAIL(HB_ERR_NOMEM, "out of memory");
		e = he_prev(c, e);
		t = c->tw[e];
		if (t == e) break;
		e = t;
		if (++steps > cap) FAIL(HB_ERR_INVALID, "fan walk does not terminate (inconsistent twin table)");
	} while (e != ein);
	return 0;
}

/* attrcode.h:117-134 (use_paral): accept the parallelogram iff all three vertices are already
 * coded and lie in region r. */
static int push_paral(coder *c, uint32_t v0, uint32_t v1, uint32_t vo, uint16_t r)
{
	const hb_mesh_desc *m = c->m;
	if (!c->vtx_done[v0] || !c->vtx_done[v1] || !c->vtx_done[vo]) return 0;
	if (m->vtx_regs[v0] != r || m->vtx_regs[v1] != r || m->vtx_regs[vo] != r) return 0;
	if (u32vec_push(&c->cand, v0) || u32vec_push(&c->cand, v1) || u32vec_push(&c->cand, vo)) FAIL(HB_ERR_NOMEM, "out of memory");
	return 0;
}

/* attrcode.h:155-171 (paral) for every fan edge, attrcode.h:173-176 (tfan).
 * Triangle: the parallelogram across the edge opposite the centre vertex.  Otherwise the polygon's
 * own neighbours (next, prev, next-next); polygons with more than 4 edges add a second one whose
 * third operand equals the second (reference quirk). */
static int gather_vertex_candidates(coder *c, uint32_t ein, uint16_t r)
{
	const hb_mesh_desc *m = c->m;
	u32vec fan = { 0, 0, 0 };
	int rc = fan_walk(c, ein, &fan);
	c->cand.n = 0;
	for (size_t k = 0; rc == 0 && k < fan.n; ++k) {
		const uint32_t e = fan.v[k];
		const uint32_t f = c->hface[e];
		const uint32_t deg = m->face_off[f + 1] - m->face_off[f];
		if (deg == 3) {
			const uint32_t e1 = he_next(c, e);
			const uint32_t t = c->tw[e1];
			if (t == e1) continue;
			const uint32_t tn = he_next(c, t);
			rc = push_paral(c, c->org[t], c->org[tn], c->org[he_next(c, tn)], r);
		} else {
			const uint32_t e0 = he_next(c, e), e1 = he_prev(c, e);
			rc = push_paral(c, c->org[e0], c->org[e1], c->org[he_next(c, e0)], r);
			if (rc == 0 && deg > 4) rc = push_paral(c, c->org[e0], c->org[e1], c->org[e1], r);
		}
	}
	free(fan.v);
	return rc;
}

/* attrcode.h:135-154 (use_corner) over attrcode.h:177-180 (tfan_corner): candidates are the
 * corners at the same vertex in fan faces that are already coded and in face region r.
 * Candidate = half-edge index (the corner binding is indexed by half-edge, attr.h:127-130). */
static int gather_corner_candidates(coder *c, uint32_t ein, uint16_t r)
{
	u32vec fan = { 0, 0, 0 };
	int rc = fan_walk(c, ein, &fan);
	c->cand.n = 0;
	for (size_t k = 0; rc == 0 && k < fan.n; ++k) {
		const uint32_t f = c->hface[fan.v[k]];
		if (!c->face_done[f] || c->m->face_regs[f] != r) continue;
		if (u32vec_push(&c->cand, fan.v[k])) { snprintf(g_err, sizeof g_err, "out of memory"); rc = HB_ERR_NOMEM; }
	}
	free(fan.v);
	return rc;
}

/* attrcode.h:182-208 (get_prediction) for one component.
 * cands: k candidate values (bit containers).  Integer storage types: mean in int64 (uint64 for
 * ULONG) with transform::divround (transform.h:90-91), truncated to T.  Float: mean in double in
 * candidate order, rounded to float, then the candidate closest to it (strict <, later wins).  */
static uint64_t combine_candidates(int stype, const uint64_t *cands, size_t k)
{
	if (k == 0) return 0;
	if (stype == HB_FLOAT) {
		double sum = 0.0;
		for (size_t i = 0; i < k; ++i) sum = sum + (double)bits_to_f32(cands[i]);
		const float avg = (float)(sum / (double)(int)k);
		float res = FLT_MAX;
		for (size_t i = 0; i < k; ++i) {
			const float p = bits_to_f32(cands[i]);
			const float rd = avg > res ? avg - res : res - avg;
			const float pd = avg > p ? avg - p : p - avg;
			res = rd < pd ? res : p;
		}
		return f32_to_bits(res);
	}
	if (stype == HB_ULONG) {
		uint64_t sum = 0;
		for (size_t i = 0; i < k; ++i) sum += cands[i];
		const uint64_t d = (uint64_t)(int)k;
		return (sum + (d >> 1)) / d;
	}
	uint64_t sum = 0;
	for (size_t i = 0; i < k; ++i) sum += (uint64_t)bits_to_i64(cands[i], stype);
	const int64_t d = (int64_t)(int)k;
	const int64_t avg = (int64_t)(sum + (uint64_t)(d >> 1)) / d;
	uint64_t out = (uint64_t)avg;
	if (TYPE_SIZE[stype] < 8) out &= (1ull << (8 * TYPE_SIZE[stype])) - 1ull;
	return out;
}

/* Prediction row of list l for a vertex: per component, predict() of every accepted
 * parallelogram (attrcode.h:123-132) combined by combine_candidates. */
static int predict_vertex_row(coder *c, int l, int slot, uint64_t *pred)
{
	const hb_mesh_desc *m = c->m;
	const hb_list_desc *L = &m->lists[l];
	const list_info *li = &c->li[l];
	const size_t k = c->cand.n / 3;
	uint64_t stackbuf[64], *tmp = stackbuf;
	if (k > 64) {
		tmp = (uint64_t *)malloc(k * sizeof(uint64_t));
		if (!tmp) FAIL(HB_ERR_NOMEM, "out of memory");
	}
	const uint8_t *rows = (const uint8_t *)L->rows;
	for (int j = 0; j < L->ncomp; ++j) {
		for (size_t i = 0; i < k; ++i) {
			const uint32_t r0 = m->bind_vtx_attr[(size_t)c->cand.v[3 * i] * m->nb_vtx + slot];
			const uint32_t r1 = m->bind_vtx_attr[(size_t)c->cand.v[3 * i + 1] * m->nb_vtx + slot];
			const uint32_t ro = m->bind_vtx_attr[(size_t)c->cand.v[3 * i + 2] * m->nb_vtx + slot];
			if (r0 >= L->nrows || r1 >= L->nrows || ro >= L->nrows) { if (tmp != stackbuf) free(tmp); FAIL(HB_ERR_INVALID, "vertex binding out of range"); }
			const uint64_t a = ld_bits(rows + (size_t)r0 * L->stride + L->offset[j], li->stype[j]);
			const uint64_t b = ld_bits(rows + (size_t)r1 * L->stride + L->offset[j], li->stype[j]);
			const uint64_t o = ld_bits(rows + (size_t)ro * L->stride + L->offset[j], li->stype[j]);
			tmp[i] = ho_predict(li->stype[j], a, b, o, li->q[j]);
		}
		pred[j] = combine_candidates(li->stype[j], tmp, k);
	}
	if (tmp != stackbuf) free(tmp);
	return 0;
}

/* Same for a corner: candidates are single rows (pred::predict_face is the identity,
 * prediction.h:149-164). */
static int predict_corner_row(coder *c, int l, int slot, uint64_t *pred)
{
	const hb_mesh_desc *m = c->m;
	const hb_list_desc *L = &m->lists[l];
	const list_info *li = &c->li[l];
	const size_t k = c->cand.n;
	uint64_t stackbuf[64], *tmp = stackbuf;
	if (k > 64) {
		tmp = (uint64_t *)malloc(k * sizeof(uint64_t));
		if (!tmp) FAIL(HB_ERR_NOMEM, "out of memory");
	}
	const uint8_t *rows = (const uint8_t *)L->rows;
	for (int j = 0; j < L->ncomp; ++j) {
		for (size_t i = 0; i < k; ++i) {
			const uint32_t row = m->bind_corner_attr[(size_t)c->cand.v[i] * m->nb_corner + slot];
			if (row >= L->nrows) { if (tmp != stackbuf) free(tmp); FAIL(HB_ERR_INVALID, "corner binding out of range"); }
			tmp[i] = ld_bits(rows + (size_t)row * L->stride + L->offset[j], li->stype[j]);
		}
		pred[j] = combine_candidates(li->stype[j], tmp, k);
	}
	if (tmp != stackbuf) free(tmp);
	return 0;
}

/* growable per-list output */
typedef struct {
	u8vec type, sym;
	u32vec aux;
	uint32_t *first_tidx; /* GlobalHistory::tidxlist, attrcode.h:23-53 */
	uint32_t tidx;
} list_out;

#define UNSET 0xffffffffu

static int emit_type(list_out *o, uint8_t t, uint32_t aux)
{
	if (u8vec_append(&o->type, &t, 1) || u32vec_push(&o->aux, aux)) FAIL(HB_ERR_NOMEM, "out of memory");
	return 0;
}

/* attrcode.h:334-342 / 356-364 / 383-391: global history test, then the residual row. */
static int emit_ghist_or_data(coder *c, list_out *o, int l, uint32_t idx, const uint64_t *pred)
{
	const hb_list_desc *L = &c->m->lists[l];
	const list_info *li = &c->li[l];
	if (idx >= L->nrows) FAIL(HB_ERR_INVALID, "binding out of range");
	if (o->first_tidx[idx] != UNSET) return emit_type(o, HB_HIST, o->tidx - 1 - o->first_tidx[idx]);
	o->first_tidx[idx] = o->tidx++;
	int rc = emit_type(o, HB_DATA, 0);
	if (rc) return rc;
	uint8_t row[HB_MAX_COMP * 8];
	const uint8_t *src = (const uint8_t *)L->rows + (size_t)idx * L->stride;
	for (int j = 0; j < L->ncomp; ++j) {
		const uint64_t raw = ld_bits(src + L->offset[j], li->stype[j]);
		const uint64_t res = ho_encode_delta(li->stype[j], raw, pred[j], li->q[j]);
		memcpy(row + li->sym_off[j], &res, (size_t)li->size[j]);
	}
	if (u8vec_append(&o->sym, row, li->sym_stride)) FAIL(HB_ERR_NOMEM, "out of memory");
	return 0;
}

void ho_streams_free(hb_streams *s)
{
	if (!s) return;
	if (s->lists) {
		for (int l = 0; l < s->nlists; ++l) {
			free(s->lists[l].type); free(s->lists[l].aux); free(s->lists[l].symbols); free(s->lists[l].hist);
		}
		free(s->lists);
	}
	free(s->reg_vtx); free(s->reg_face);
	free(s);
}

static uint32_t order_halfedge(const hb_mesh_desc *m, const void *order, uint32_t i, int *ok)
{
	uint32_t f;
	uint16_t e;
	if (!order) { *ok = i < m->nf; return *ok ? m->face_off[i] : 0; } /* decoder's face order: (i, 0) */
	memcpy(&f, (const uint8_t *)order + 8 * (size_t)i, 4);
	memcpy(&e, (const uint8_t *)order + 8 * (size_t)i + 4, 2);
	*ok = f < m->nf && m->face_off[f] + e < m->face_off[f + 1];
	return *ok ? m->face_off[f] + e : 0;
}

/* per-(corner slot, vertex) local history, attrcode.h:54-80 */
typedef struct { uint32_t *v; uint32_t n, cap; } lhist_cell;

int ho_attr_encode(const hb_mesh_desc *m, hb_streams **out)
{
	coder c;
	int rc = coder_init(&c, m);
	list_out *lo = NULL;
	lhist_cell *lh = NULL;
	hb_streams *s = NULL;
	uint64_t pred[HB_MAX_COMP];
	const uint32_t nface_order = m->order_f ? m->norder_f : m->nf;
	if (rc) goto done;
	lo = (list_out *)calloc(m->nlists + 1, sizeof(list_out));
	s = (hb_streams *)calloc(1, sizeof(hb_streams));
	if (m->nb_corner) lh = (lhist_cell *)calloc((size_t)m->nb_corner * m->nv + 1, sizeof(lhist_cell));
	if (!lo || !s || (m->nb_corner && !lh)) { rc = HB_ERR_NOMEM; snprintf(g_err, sizeof g_err, "out of memory"); goto done; }
	for (int l = 0; l < m->nlists; ++l) {
		lo[l].first_tidx = (uint32_t *)malloc(sizeof(uint32_t) * (m->lists[l].nrows + 1));
		if (!lo[l].first_tidx) { rc = HB_ERR_NOMEM; goto done; }
		memset(lo[l].first_tidx, 0xff, sizeof(uint32_t) * (m->lists[l].nrows + 1));
	}
	s->n_vtx = m->norder;
	s->n_face = nface_order;
	s->nlists = m->nlists;
	s->reg_vtx = (uint16_t *)malloc(sizeof(uint16_t) * (m->norder + 1));
	s->reg_face = (uint16_t *)malloc(sizeof(uint16_t) * (nface_order + 1));
	s->lists = (hb_list_streams *)calloc(m->nlists + 1, sizeof(hb_list_streams));
	if (!s->reg_vtx || !s->reg_face || !s->lists) { rc = HB_ERR_NOMEM; goto done; }

	/* vertices in traversal order: attrcode.h:398-404 -> vtx_post :321-344 -> vtx :209-225 */
	for (uint32_t i = 0; i < m->norder && rc == 0; ++i) {
		int ok;
		const uint32_t h = order_halfedge(m, m->order, i, &ok);
		if (!ok) { rc = HB_ERR_INVALID; snprintf(g_err, sizeof g_err, "order[%u] out of range", i); break; }
		const uint32_t v = c.org[h];
		const uint16_t r = m->vtx_regs[v];
		if (r >= m->nregs_vtx) { rc = HB_ERR_INVALID; snprintf(g_err, sizeof g_err, "vertex region out of range"); break; }
		rc = gather_vertex_candidates(&c, h, r);
		c.vtx_done[v] = 1;
		s->reg_vtx[i] = r;
		for (int a = 0; rc == 0 && a < m->off_reg_vtx[r + 1] - m->off_reg_vtx[r]; ++a) {
			const int l = m->reg_vtxlist[m->off_reg_vtx[r] + a];
			const uint32_t idx = m->bind_vtx_attr[(size_t)v * m->nb_vtx + a];
			rc = predict_vertex_row(&c, l, a, pred);
			if (rc == 0) rc = emit_ghist_or_data(&c, &lo[l], l, idx, pred);
		}
	}
	/* faces then their corners: attrcode.h:405-414 */
	for (uint32_t i = 0; i < nface_order && rc == 0; ++i) {
		int ok;
		const uint32_t hg = order_halfedge(m, m->order_f, i, &ok);
		if (!ok) { rc = HB_ERR_INVALID; snprintf(g_err, sizeof g_err, "order_f[%u] out of range", i); break; }
		const uint32_t f = c.hface[hg];
		const uint16_t r = m->face_regs[f];
		if (r >= m->nregs_face) { rc = HB_ERR_INVALID; snprintf(g_err, sizeof g_err, "face region out of range"); break; }
		/* face_post :345-365.  The neighbour loop (:245-254) offers the face itself, which is not
		 * yet marked coded, so there is never a candidate: prediction 0. */
		s->reg_face[i] = r;
		c.face_done[f] = 1;
		memset(pred, 0, sizeof pred);
		for (int a = 0; rc == 0 && a < m->off_reg_face[r + 1] - m->off_reg_face[r]; ++a) {
			const int l = m->reg_facelist[m->off_reg_face[r] + a];
			rc = emit_ghist_or_data(&c, &lo[l], l, m->bind_face_attr[(size_t)f * m->nb_face + a], pred);
		}
		/* corners, starting at the gate corner: corner_post :367-393 -> corner :272-288 */
		const uint32_t f0 = m->face_off[f], deg = m->face_off[f + 1] - f0;
		const int ncs = m->off_reg_corner[r + 1] - m->off_reg_corner[r];
		for (uint32_t j = 0; j < deg && rc == 0 && ncs > 0; ++j) {
			uint32_t h = hg + j;
			if (h >= f0 + deg) h -= deg;
			c.face_done[f] = 0;
			rc = gather_corner_candidates(&c, h, r);
			c.face_done[f] = 1;
			const uint32_t v = c.org[h];
			for (int a = 0; rc == 0 && a < ncs; ++a) {
				const int l = m->reg_cornerlist[m->off_reg_corner[r] + a];
				const uint32_t idx = m->bind_corner_attr[(size_t)h * m->nb_corner + a];
				/* LocalHistory::insert :67-74 */
				lhist_cell *cell = &lh[(size_t)a * m->nv + v];
				uint32_t hit = UNSET;
				for (uint32_t p = 0; p < cell->n; ++p)
					if (cell->v[p] == idx) { hit = cell->n - 1 - p; break; }
				if (hit != UNSET) { rc = emit_type(&lo[l], HB_LHIST, hit); continue; }
				if (cell->n == cell->cap) {
					uint32_t nc = cell->cap ? cell->cap * 2 : 4;
					uint32_t *p = (uint32_t *)realloc(cell->v, nc * sizeof(uint32_t));
					if (!p) { rc = HB_ERR_NOMEM; break; }
					cell->v = p;
					cell->cap = nc;
				}
				cell->v[cell->n++] = idx;
				rc = predict_corner_row(&c, l, a, pred);
				if (rc == 0) rc = emit_ghist_or_data(&c, &lo[l], l, idx, pred);
			}
		}
	}
	/* finalize: move vectors into the stream structs and count the byte-plane histograms
	 * (= AdaptiveStatisticsModule::C[] - 1 after coding, stat_adaptive.h:117-124, model.h:57-66) */
	for (int l = 0; l < m->nlists && rc == 0; ++l) {
		hb_list_streams *ls = &s->lists[l];
		ls->n_emit = (uint32_t)lo[l].type.n;
		ls->n_data = lo[l].tidx;
		ls->sym_stride = c.li[l].sym_stride;
		ls->type = lo[l].type.v; lo[l].type.v = NULL;
		ls->aux = lo[l].aux.v; lo[l].aux.v = NULL;
		ls->symbols = lo[l].sym.v; lo[l].sym.v = NULL;
		ls->hist = (uint64_t *)calloc((size_t)ls->sym_stride * 256 + 1, sizeof(uint64_t));
		if (!ls->hist) { rc = HB_ERR_NOMEM; break; }
		for (uint32_t k = 0; k < ls->n_data; ++k)
			for (uint32_t p = 0; p < ls->sym_stride; ++p) ls->hist[p * 256 + ls->symbols[(size_t)k * ls->sym_stride + p]]++;
		for (uint32_t k = 0; k < ls->n_emit; ++k) ls->type_hist[ls->type[k]]++;
	}
done:
	if (lo) {
		for (int l = 0; l < m->nlists; ++l) { free(lo[l].type.v); free(lo[l].aux.v); free(lo[l].sym.v); free(lo[l].first_tidx); }
		free(lo);
	}
	if (lh) {
		for (size_t k = 0; k < (size_t)m->nb_corner * m->nv; ++k) free(lh[k].v);
		free(lh);
	}
	coder_free(&c);
	if (rc) { ho_streams_free(s); s = NULL; }
	*out = s;
	return rc;
}

/* In-place reconstruction of one DATA row: attrcode.h:461 / :492 / :517. */
static void reconstruct_row(coder *c, int l, uint32_t idx, const uint64_t *pred)
{
	const hb_list_desc *L = &c->m->lists[l];
	const list_info *li = &c->li[l];
	uint8_t *row = (uint8_t *)L->rows + (size_t)idx * L->stride;
	for (int j = 0; j < L->ncomp; ++j) {
		const uint64_t delta = ld_bits(row + L->offset[j], li->stype[j]);
		st_bits(row + L->offset[j], li->stype[j], ho_decode_delta(li->stype[j], delta, pred[j], li->q[j]));
	}
}

/* Decode-side candidate gathering differs from the encoder in one point: the reference decoder's
 * rows start zeroed and are filled at their DATA emission (attrcode.h:458-461), so a candidate
 * row that has not been emitted yet contributes zeros.  `done[l][row]` tracks emission. */
static int predict_vertex_row_dec(coder *c, int l, int slot, uint8_t *const *done, uint64_t *pred)
{
	const hb_mesh_desc *m = c->m;
	const hb_list_desc *L = &m->lists[l];
	const list_info *li = &c->li[l];
	const size_t k = c->cand.n / 3;
	uint64_t stackbuf[64], *tmp = stackbuf;
	if (k > 64) {
		tmp = (uint64_t *)malloc(k * sizeof(uint64_t));
		if (!tmp) FAIL(HB_ERR_NOMEM, "out of memory");
	}
	const uint8_t *rows = (const uint8_t *)L->rows;
	for (int j = 0; j < L->ncomp; ++j) {
		for (size_t i = 0; i < k; ++i) {
			uint64_t v[3];
			for (int t = 0; t < 3; ++t) {
				const uint32_t r = m->bind_vtx_attr[(size_t)c->cand.v[3 * i + t] * m->nb_vtx + slot];
				if (r >= L->nrows) { if (tmp != stackbuf) free(tmp); FAIL(HB_ERR_INVALID, "vertex binding out of range"); }
				v[t] = done[l][r] ? ld_bits(rows + (size_t)r * L->stride + L->offset[j], li->stype[j]) : 0;
			}
			tmp[i] = ho_predict(li->stype[j], v[0], v[1], v[2], li->q[j]);
		}
		pred[j] = combine_candidates(li->stype[j], tmp, k);
	}
	if (tmp != stackbuf) free(tmp);
	return 0;
}

static int predict_corner_row_dec(coder *c, int l, int slot, uint8_t *const *done, uint64_t *pred)
{
	const hb_mesh_desc *m = c->m;
	const hb_list_desc *L = &m->lists[l];
	const list_info *li = &c->li[l];
	const size_t k = c->cand.n;
	uint64_t stackbuf[64], *tmp = stackbuf;
	if (k > 64) {
		tmp = (uint64_t *)malloc(k * sizeof(uint64_t));
		if (!tmp) FAIL(HB_ERR_NOMEM, "out of memory");
	}
	const uint8_t *rows = (const uint8_t *)L->rows;
	for (int j = 0; j < L->ncomp; ++j) {
		for (size_t i = 0; i < k; ++i) {
			const uint32_t r = m->bind_corner_attr[(size_t)c->cand.v[i] * m->nb_corner + slot];
			if (r >= L->nrows) { if (tmp != stackbuf) free(tmp); FAIL(HB_ERR_INVALID, "corner binding out of range"); }
			tmp[i] = done[l][r] ? ld_bits(rows + (size_t)r * L->stride + L->offset[j], li->stype[j]) : 0;
		}
		pred[j] = combine_candidates(li->stype[j], tmp, k);
	}
	if (tmp != stackbuf) free(tmp);
	return 0;
}

/* Is this emission of list l the one that carried the row?  With the drained type stream: the
 * emission's type symbol is DATA.  Without: the first reference of the row (see harry_b200.h). */
static int is_data_emission(const hb_mesh_desc *m, int l, uint32_t *cursor, uint8_t *const *done, uint32_t idx)
{
	const uint32_t k = cursor[l]++;
	if (m->emit_type && m->emit_type[l]) return m->emit_type[l][k] == HB_DATA;
	return !done[l][idx];
}

/* attrcode.h:534-550 with the symbol drain factored out: the rows hold residuals,
 * the binding tables are complete. */
int ho_attr_decode(const hb_mesh_desc *m)
{
	coder c;
	int rc = coder_init(&c, m);
	uint8_t **done = NULL;
	uint32_t *cursor = NULL;
	uint64_t pred[HB_MAX_COMP];
	const uint32_t nface_order = m->order_f ? m->norder_f : m->nf;
	if (rc) goto done_;
	done = (uint8_t **)calloc(m->nlists + 1, sizeof(uint8_t *));
	cursor = (uint32_t *)calloc(m->nlists + 1, sizeof(uint32_t));
	if (!done || !cursor) { rc = HB_ERR_NOMEM; goto done_; }
	for (int l = 0; l < m->nlists; ++l) {
		done[l] = (uint8_t *)calloc(m->lists[l].nrows + 1, 1);
		if (!done[l]) { rc = HB_ERR_NOMEM; goto done_; }
	}
	for (uint32_t i = 0; i < m->norder && rc == 0; ++i) { /* vtx_post :443-470 */
		int ok;
		const uint32_t h = order_halfedge(m, m->order, i, &ok);
		if (!ok) { rc = HB_ERR_INVALID; snprintf(g_err, sizeof g_err, "order[%u] out of range", i); break; }
		const uint32_t v = c.org[h];
		const uint16_t r = m->vtx_regs[v];
		if (r >= m->nregs_vtx) { rc = HB_ERR_INVALID; snprintf(g_err, sizeof g_err, "vertex region out of range"); break; }
		rc = gather_vertex_candidates(&c, h, r);
		c.vtx_done[v] = 1;
		for (int a = 0; rc == 0 && a < m->off_reg_vtx[r + 1] - m->off_reg_vtx[r]; ++a) {
			const int l = m->reg_vtxlist[m->off_reg_vtx[r] + a];
			const uint32_t idx = m->bind_vtx_attr[(size_t)v * m->nb_vtx + a];
			if (idx >= m->lists[l].nrows) { rc = HB_ERR_INVALID; snprintf(g_err, sizeof g_err, "binding out of range"); break; }
			if (!is_data_emission(m, l, cursor, done, idx)) continue;
			rc = predict_vertex_row_dec(&c, l, a, done, pred);
			if (rc == 0) reconstruct_row(&c, l, idx, pred);
			done[l][idx] = 1;
		}
	}
	for (uint32_t i = 0; i < nface_order && rc == 0; ++i) { /* face_post :476-501, corner_post :502-531 */
		int ok;
		const uint32_t hg = order_halfedge(m, m->order_f, i, &ok);
		if (!ok) { rc = HB_ERR_INVALID; snprintf(g_err, sizeof g_err, "order_f[%u] out of range", i); break; }
		const uint32_t f = c.hface[hg];
		const uint16_t r = m->face_regs[f];
		if (r >= m->nregs_face) { rc = HB_ERR_INVALID; snprintf(g_err, sizeof g_err, "face region out of range"); break; }
		c.face_done[f] = 1;
		memset(pred, 0, sizeof pred);
		for (int a = 0; rc == 0 && a < m->off_reg_face[r + 1] - m->off_reg_face[r]; ++a) {
			const int l = m->reg_facelist[m->off_reg_face[r] + a];
			const uint32_t idx = m->bind_face_attr[(size_t)f * m->nb_face + a];
			if (idx >= m->lists[l].nrows) { rc = HB_ERR_INVALID; snprintf(g_err, sizeof g_err, "binding out of range"); break; }
			if (!is_data_emission(m, l, cursor, done, idx)) continue;
			reconstruct_row(&c, l, idx, pred);
			done[l][idx] = 1;
		}
		const uint32_t f0 = m->face_off[f], deg = m->face_off[f + 1] - f0;
		const int ncs = m->off_reg_corner[r + 1] - m->off_reg_corner[r];
		for (uint32_t j = 0; j < deg && rc == 0 && ncs > 0; ++j) {
			uint32_t h = hg + j;
			if (h >= f0 + deg) h -= deg;
			c.face_done[f] = 0;
			rc = gather_corner_candidates(&c, h, r);
			c.face_done[f] = 1;
			for (int a = 0; rc == 0 && a < ncs; ++a) {
				const int l = m->reg_cornerlist[m->off_reg_corner[r] + a];
				const uint32_t idx = m->bind_corner_attr[(size_t)h * m->nb_corner + a];
				if (idx >= m->lists[l].nrows) { rc = HB_ERR_INVALID; snprintf(g_err, sizeof g_err, "binding out of range"); break; }
				if (!is_data_emission(m, l, cursor, done, idx)) continue;
				rc = predict_corner_row_dec(&c, l, a, done, pred);
				if (rc == 0) reconstruct_row(&c, l, idx, pred);
				done[l][idx] = 1;
			}
		}
	}
done_:
	if (done) {
		for (int l = 0; l < m->nlists; ++l) free(done[l]);
		free(done);
	}
	free(cursor);
	coder_free(&c);
	return rc;
}

/* ------------------------------------------------------------------------------------------ */
/* Debug / design aid (not part of the parity surface): CPU simulation of the speculative        */
/* chunk-parallel reconstruction used by harry_b200/csrc/hb_decode_spec.cuh, for one quantized   */
/* vertex list with identity bindings.  Prints the advance of the exact prefix per sweep.        */
/* ------------------------------------------------------------------------------------------ */
int ho_debug_spec_sim(const hb_mesh_desc *m, int l, uint32_t B, uint32_t T, uint32_t max_print)
{
	coder c;
	int rc = coder_init(&c, m);
	if (rc) return rc;
	const hb_list_desc *L = &m->lists[l];
	const list_info *li = &c.li[l];
	const uint32_t n = m->norder;
	uint32_t *off = (uint32_t *)calloc(n + 1, 4);
	u32vec tri = { 0, 0, 0 };
	uint32_t *rank = (uint32_t *)malloc(4 * (m->nv + 1));
	memset(rank, 0xff, 4 * (m->nv + 1));
	uint32_t *ordv = (uint32_t *)malloc(4 * (n + 1));
	for (uint32_t i = 0; i < n; ++i) {
		int ok;
		const uint32_t h = order_halfedge(m, m->order, i, &ok);
		const uint32_t v = c.org[h];
		ordv[i] = v;
		gather_vertex_candidates(&c, h, m->vtx_regs[v]);
		c.vtx_done[v] = 1;
		rank[v] = i;
		off[i] = (uint32_t)(tri.n / 3);
		for (size_t k = 0; k < c.cand.n; ++k) u32vec_push(&tri, rank[c.cand.v[k]]);
	}
	off[n] = (uint32_t)(tri.n / 3);
	const int nc = L->ncomp;
	/* residuals = rows (decode input); x = working values */
	uint64_t *res = (uint64_t *)calloc((size_t)n * nc + 1, 8), *x = (uint64_t *)calloc((size_t)n * nc + 1, 8);
	for (uint32_t i = 0; i < n; ++i)
		for (int j = 0; j < nc; ++j) res[(size_t)i * nc + j] = ld_bits((const uint8_t *)L->rows + (size_t)m->bind_vtx_attr[(size_t)ordv[i] * m->nb_vtx] * L->stride + L->offset[j], li->stype[j]);
	uint32_t done = 0, sweeps = 0;
	uint64_t steps = 0;
	uint64_t tmp[4096];
	while (done < n) {
		uint32_t p = 0xffffffffu;
		/* emulate concurrent chunks: Jacobi semantics (all chunks read the previous sweep's values
		 * of other chunks) via a snapshot would be expensive; run chunks in DESCENDING order so a
		 * chunk never sees this sweep's results of an earlier chunk. */
		uint32_t nchunks = T;
		for (int64_t t = (int64_t)nchunks - 1; t >= 0; --t) {
			const uint64_t start = (uint64_t)done + (uint64_t)t * B;
			if (start >= n) continue;
			const uint32_t end = (uint32_t)((start + B < n) ? start + B : n);
			uint32_t fc = 0xffffffffu;
			for (uint32_t i = (uint32_t)start; i < end; ++i) {
				const uint32_t K = off[i + 1] - off[i];
				int changed = 0;
				for (int j = 0; j < nc; ++j) {
					for (uint32_t k = 0; k < K && k < 4096; ++k) {
						const uint32_t *tr = tri.v + 3 * (size_t)(off[i] + k);
						tmp[k] = ho_predict(li->stype[j], x[(size_t)tr[0] * nc + j], x[(size_t)tr[1] * nc + j], x[(size_t)tr[2] * nc + j], li->q[j]);
					}
					const uint64_t pred = combine_candidates(li->stype[j], tmp, K);
					const uint64_t nv = ho_decode_delta(li->stype[j], res[(size_t)i * nc + j], pred, li->q[j]);
					if (nv != x[(size_t)i * nc + j]) { x[(size_t)i * nc + j] = nv; changed = 1; }
				}
				if (changed && fc == 0xffffffffu) fc = i;
				++steps;
			}
			if (t == 0 && fc != 0xffffffffu) fc = end;
			if (fc < p) p = fc;
		}
		const uint64_t wend = (uint64_t)done + (uint64_t)T * B;
		const uint32_t nd = p != 0xffffffffu ? p : (uint32_t)(wend < n ? wend : n);
		if (sweeps < max_print) printf("sweep %u: done %u -> %u (+%u)\n", sweeps, done, nd, nd - done);
		done = nd;
		++sweeps;
	}
	printf("n=%u B=%u T=%u sweeps=%u steps=%llu (%.2f per rank), sequential-equivalent depth %llu\n", n, B, T, sweeps, (unsigned long long)steps, (double)steps / n, (unsigned long long)sweeps * B);
	free(off); free(tri.v); free(rank); free(ordv); free(res); free(x);
	coder_free(&c);
	return 0;
}

/* Debug: traversal-order candidate triples (ranks) of the vertices, CSR.  Caller frees with free(). */
int ho_debug_vertex_candidates(const hb_mesh_desc *m, uint32_t **off_out, uint32_t **tri_out)
{
	coder c;
	int rc = coder_init(&c, m);
	if (rc) return rc;
	const uint32_t n = m->norder;
	uint32_t *off = (uint32_t *)calloc(n + 1, 4);
	u32vec tri = { 0, 0, 0 };
	uint32_t *rank = (uint32_t *)malloc(4 * (m->nv + 1));
	memset(rank, 0xff, 4 * (m->nv + 1));
	for (uint32_t i = 0; i < n && rc == 0; ++i) {
		int ok;
		const uint32_t h = order_halfedge(m, m->order, i, &ok);
		const uint32_t v = c.org[h];
		rc = gather_vertex_candidates(&c, h, m->vtx_regs[v]);
		c.vtx_done[v] = 1;
		if (rank[v] == 0xffffffffu) rank[v] = i;
		off[i] = (uint32_t)(tri.n / 3);
		for (size_t k = 0; k < c.cand.n; ++k) u32vec_push(&tri, rank[c.cand.v[k]]);
	}
	off[n] = (uint32_t)(tri.n / 3);
	free(rank);
	coder_free(&c);
	*off_out = off;
	*tri_out = tri.v;
	return rc;
}

/* ------------------------------------------------------------------------------------------ */
/* Twin matching: mesh::Builder, structs/conn.h:164-214.        */
/*                                                                                             */
/* The reference keeps an unordered_map from the DIRECTED edge (a, b) to the half-edge that    */
/* first carried it.  Half-edges arrive in file order (face by face, corner by corner:         */
/* set_org :203-210 calls add_edge(last, vtx) for corners 1.., face_end :199-202 closes the    */
/* loop with add_edge(last, start)); half-edge (f, e) runs from org(f, e) to org(f, e+1 mod n). */
/* add_edge(a, b) :178-190: if (b, a) is in the map, merge with the half-edge stored there and  */
/* ERASE the entry; otherwise insert (a, b) -- std::unordered_map::insert keeps an existing     */
/* entry, so a second half-edge with the same direction stays a border (its twin is itself,     */
/* Conn::add_face :87-91).  Restated here with an open-addressing table (linear probing,        */
/* tombstones) instead of the node-based map; only the lookup / insert-if-absent / erase        */
/* semantics matter.                                                                            */
/* ------------------------------------------------------------------------------------------ */
typedef struct tw_slot { uint32_t a, b, he; uint32_t state; } tw_slot; /* state: 0 empty, 1 full, 2 erased */

static uint64_t tw_hash(uint32_t a, uint32_t b)
{
	uint64_t x = ((uint64_t)a << 32) | b;
	x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
	return x;
}
static tw_slot *tw_find(tw_slot *tab, uint64_t mask, uint32_t a, uint32_t b)
{
	uint64_t i = tw_hash(a, b) & mask;
	for (;; i = (i + 1) & mask) {
		if (tab[i].state == 0) return NULL;
		if (tab[i].state == 1 && tab[i].a == a && tab[i].b == b) return &tab[i];
	}
}
static void tw_insert_absent(tw_slot *tab, uint64_t mask, uint32_t a, uint32_t b, uint32_t he)
{
	uint64_t i = tw_hash(a, b) & mask;
	tw_slot *grave = NULL;
	for (;; i = (i + 1) & mask) {
		if (tab[i].state == 0) break;
		if (tab[i].state == 2) { if (!grave) grave = &tab[i]; continue; }
		if (tab[i].a == a && tab[i].b == b) return; /* insert() of an existing key is a no-op */
	}
	tw_slot *s = grave ? grave : &tab[i];
	s->a = a; s->b = b; s->he = he; s->state = 1;
}

int ho_twin_match(uint32_t nv, uint32_t nf, const uint32_t *face_off, const void *org_in, uint32_t org_stride, void *edges_out)
{
	if (nf && (!face_off || !edges_out)) FAIL(HB_ERR_INVALID, "twin_match: NULL argument");
	if (org_stride != 4 && org_stride != 12) FAIL(HB_ERR_INVALID, "twin_match: org_stride must be 4 or 12");
	uint32_t ne = nf ? face_off[nf] : 0;
	if (ne && !org_in) FAIL(HB_ERR_INVALID, "twin_match: NULL org");
	uint8_t *out = (uint8_t *)edges_out;
	uint32_t *org = (uint32_t *)malloc(sizeof(uint32_t) * (ne ? ne : 1)); /* packed copy (org_in may alias edges_out) */
	if (!org) FAIL(HB_ERR_NOMEM, "twin_match: out of memory");
	for (uint32_t h = 0; h < ne; ++h) memcpy(&org[h], (const uint8_t *)org_in + (size_t)h * org_stride, 4);
	uint32_t *he_face = (uint32_t *)malloc(sizeof(uint32_t) * (ne ? ne : 1));
	uint32_t *twin = (uint32_t *)malloc(sizeof(uint32_t) * (ne ? ne : 1));
	uint64_t cap = 16;
	while (cap < 2 * (uint64_t)ne) cap <<= 1;
	tw_slot *tab = (tw_slot *)calloc(cap, sizeof(tw_slot));
	if (!he_face || !twin || !tab) { free(org); free(he_face); free(twin); free(tab); FAIL(HB_ERR_NOMEM, "twin_match: out of memory"); }
	int rc = 0;
	for (uint32_t f = 0; f < nf && !rc; ++f) {
		uint32_t o = face_off[f], n = face_off[f + 1] - o;
		if (face_off[f + 1] < o || face_off[f + 1] > ne || n > 0xffffu) { rc = HB_ERR_INVALID; break; } /* fepair::e is 16 bits wide (structs/conn.h:22-31) */
		for (uint32_t e = 0; e < n; ++e) {
			uint32_t h = o + e;
			he_face[h] = f;
			twin[h] = h; /* Conn::add_face: every half-edge starts as its own twin */
			if (org[h] >= nv) { rc = HB_ERR_INVALID; break; }
		}
	}
	for (uint32_t f = 0; f < nf && !rc; ++f) {
		uint32_t o = face_off[f], n = face_off[f + 1] - o;
		for (uint32_t e = 0; e < n; ++e) {
			uint32_t h = o + e, a = org[h], b = org[o + (e + 1 == n ? 0 : e + 1)];
			tw_slot *t = tw_find(tab, cap - 1, b, a);
			if (t) {
				twin[t->he] = h; /* Conn::fmerge :154-158 */
				twin[h] = t->he;
				t->state = 2;
			} else {
				tw_insert_absent(tab, cap - 1, a, b, h);
			}
		}
	}
	if (!rc) {
		for (uint32_t h = 0; h < ne; ++h) {
			uint32_t t = twin[h], tf = he_face[t];
			uint16_t te = (uint16_t)(t - face_off[tf]), pad = 0;
			memcpy(out + 12 * (size_t)h, &org[h], 4);
			memcpy(out + 12 * (size_t)h + 4, &tf, 4);
			memcpy(out + 12 * (size_t)h + 8, &te, 2);
			memcpy(out + 12 * (size_t)h + 10, &pad, 2);
		}
	}
	free(org); free(he_face); free(twin); free(tab);
	if (rc) FAIL(rc, "twin_match: malformed face_off (decreasing, beyond ne, face with more than 65535 corners) or vertex index >= nv");
	return 0;
}
