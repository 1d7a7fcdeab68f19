// Device-side numerical primitives for the B200 ASTC block compressor.
// Semantics follow the reference's scalar vecmathlib back-end op for op (file:line cites are relative to
// /root/reference/Source): IEEE binary32, round-to-nearest, no FMA contraction (compile with -fmad=false),
// a<b?a:b min/max, x86 cvttps2dq float->int. See DESIGN.md "numerical contract".
#pragma once

struct f4 {
	float x, y, z, w;
};
struct i4 {
	int x, y, z, w;
};

ASTC_FN f4 mk4(float a, float b, float c, float d) { f4 r = {a, b, c, d}; return r; }
ASTC_FN f4 splat4(float a) { f4 r = {a, a, a, a}; return r; }
ASTC_FN i4 mki4(int a, int b, int c, int d) { i4 r = {a, b, c, d}; return r; }
ASTC_FN float lane(const f4& v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }
ASTC_FN void set_lane(f4& v, int i, float a) { if (i == 0) v.x = a; else if (i == 1) v.y = a; else if (i == 2) v.z = a; else v.w = a; }
ASTC_FN int lanei(const i4& v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }

ASTC_FN f4 operator+(f4 a, f4 b) { return mk4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
ASTC_FN f4 operator-(f4 a, f4 b) { return mk4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
ASTC_FN f4 operator*(f4 a, f4 b) { return mk4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
ASTC_FN f4 operator/(f4 a, f4 b) { return mk4(a.x / b.x, a.y / b.y, a.z / b.z, a.w / b.w); }
ASTC_FN f4 operator*(f4 a, float b) { return mk4(a.x * b, a.y * b, a.z * b, a.w * b); }
ASTC_FN f4 operator/(f4 a, float b) { return mk4(a.x / b, a.y / b, a.z / b, a.w / b); }

// astc::min / astc::max and vector min/max: "a < b ? a : b" (astcenc_mathlib.h:168-223,
// astcenc_vecmathlib_none_4.h:840-859) - returns b when either is NaN
ASTC_FN float minf(float a, float b) { return a < b ? a : b; }
ASTC_FN float maxf(float a, float b) { return a > b ? a : b; }
ASTC_FN int mini(int a, int b) { return a < b ? a : b; }
ASTC_FN int maxi(int a, int b) { return a > b ? a : b; }
ASTC_FN f4 min4(f4 a, f4 b) { return mk4(minf(a.x, b.x), minf(a.y, b.y), minf(a.z, b.z), minf(a.w, b.w)); }
ASTC_FN f4 max4(f4 a, f4 b) { return mk4(maxf(a.x, b.x), maxf(a.y, b.y), maxf(a.z, b.z), maxf(a.w, b.w)); }
// astc::clamp(v, mn, mx) (astcenc_mathlib.h:272-279): NaN -> mn
ASTC_FN float clampf(float v, float mn, float mx) { if (v > mx) return mx; if (v > mn) return v; return mn; }
ASTC_FN int clampi(int v, int mn, int mx) { if (v > mx) return mx; if (v > mn) return v; return mn; }
ASTC_FN float clamp1f(float v) { return clampf(v, 0.0f, 1.0f); }
// vector clamp(lo, hi, a) = min(max(a, lo), hi) (astcenc_vecmathlib_common_4.h:225-229): NaN -> lo
ASTC_FN float vclampf(float lo, float hi, float a) { return minf(maxf(a, lo), hi); }
ASTC_FN f4 vclamp4(float lo, float hi, f4 a) { return mk4(vclampf(lo, hi, a.x), vclampf(lo, hi, a.y), vclampf(lo, hi, a.z), vclampf(lo, hi, a.w)); }
ASTC_FN float clampzo(float a) { return vclampf(0.0f, 1.0f, a); }
// abs(a) = max(0 - a, a) (SSE, astcenc_vecmathlib_sse_4.h:808) - same as fabsf for non-NaN
ASTC_FN float absf(float a) { return fabsf(a); }

// hadd_s = (l0+l2)+(l1+l3) (astcenc_vecmathlib_none_4.h:907); hadd_rgb_s = (l0+l1)+l2 (common_4.h:287)
ASTC_FN float hadd_s(f4 a) { return (a.x + a.z) + (a.y + a.w); }
ASTC_FN float hadd_rgb_s(f4 a) { return (a.x + a.y) + a.z; }
ASTC_FN float dot_s(f4 a, f4 b) { return hadd_s(a * b); }
ASTC_FN float dot3_s(f4 a, f4 b) { f4 m = a * b; return (m.x + m.y) + m.z; }
ASTC_FN float hmin_s(f4 a) { return minf(minf(a.x, a.y), minf(a.z, a.w)); }
ASTC_FN float hmax_s(f4 a) { return maxf(maxf(a.x, a.y), maxf(a.z, a.w)); }

// The 4-lane accumulator behind vfloatacc + haccumulate (astcenc_vecmathlib.h:93-97,
// common_4.h:270-282, avx2_8.h:868-896): element i of the stream goes to lane i mod 4.
struct acc4 {
	float l[4];
	int n;
};
ASTC_FN void acc_init(acc4& a) { a.l[0] = a.l[1] = a.l[2] = a.l[3] = 0.0f; a.n = 0; }
ASTC_FN void acc_add(acc4& a, float v) { a.l[a.n & 3] = a.l[a.n & 3] + v; a.n++; }
// restart lane assignment at lane 0 (a new vector loop starts), keeping the sums
ASTC_FN void acc_restart(acc4& a) { a.n = 0; }
ASTC_FN float acc_sum(const acc4& a) { return (a.l[0] + a.l[2]) + (a.l[1] + a.l[3]); }

// float_to_int = C truncation with x86 cvttps2dq semantics for NaN / out of range
// (astcenc_vecmathlib_none_4.h:983, sse_4.h:942)
ASTC_FN int f2i(float a) {
	if (!(a > -2147483904.0f && a < 2147483648.0f)) {
		return (int)0x80000000u;
	}
	return (int)a;
}
// float_to_int_rtn / astc::flt2int_rtn = trunc(a + 0.5f) (none_4.h:994, mathlib.h:328)
ASTC_FN int f2i_rtn(float a) { return f2i(a + 0.5f); }
// round() = nearest even (none_4.h:875 / _MM_FROUND_TO_NEAREST_INT)
ASTC_FN float round_ne(float a) { return ASTC_RINT(a); }

ASTC_FN uint32_t f_as_u(float f) { return ASTC_F2U(f); }
ASTC_FN float u_as_f(uint32_t u) { return ASTC_U2F(u); }

// atan / atan2 approximations (astcenc_vecmathlib.h:275-306)
ASTC_FN float change_sign(float a, float b) { return u_as_f(f_as_u(a) ^ (f_as_u(b) & 0x80000000u)); }
ASTC_FN float approx_atan(float x) {
	const float PI_OVER_TWO = 1.57079632679489661923f;
	bool c = absf(x) > 1.0f;
	float z = change_sign(PI_OVER_TWO, x);
	float y = c ? 1.0f / x : x;
	y = y / (y * y * 0.28f + 1.0f);
	return c ? z - y : y;
}
ASTC_FN float approx_atan2(float y, float x) {
	const float PI = 3.14159265358979323846f;
	float z = approx_atan(absf(y / x));
	bool xmask = x < 0.0f;
	return change_sign(xmask ? PI - z : z, y);
}

// normalize / normalize_safe (astcenc_vecmathlib.h:353-371): 4-lane dot
ASTC_FN f4 normalize4(f4 a) {
	float len = dot_s(a, a);
	float s = sqrtf(len);
	return mk4(a.x / s, a.y / s, a.z / s, a.w / s);
}
ASTC_FN f4 normalize_safe4(f4 a, f4 safe) {
	float len = dot_s(a, a);
	if (len != 0.0f) {
		float s = sqrtf(len);
		return mk4(a.x / s, a.y / s, a.z / s, a.w / s);
	}
	return safe;
}
ASTC_FN f4 unit4() { return splat4(0.5f); }
ASTC_FN f4 unit3() { return mk4(0.577350258827209473f, 0.577350258827209473f, 0.577350258827209473f, 0.0f); }
ASTC_FN f4 unit2() { return mk4(0.707106769084930420f, 0.707106769084930420f, 0.0f, 0.0f); }

// fp16 <-> fp32 (F16C semantics: round to nearest even) (astcenc_vecmathlib_sse_4.h:967-1000)
ASTC_NOINLINE uint16_t float_to_sf16(float f) {
	uint32_t u = f_as_u(f);
	uint32_t sign = (u >> 16) & 0x8000u;
	uint32_t exp = (u >> 23) & 0xFF;
	uint32_t mant = u & 0x7FFFFFu;
	if (exp == 0xFF) {
		if (mant) {
			return (uint16_t)(sign | 0x7C00u | 0x200u | (mant >> 13));
		}
		return (uint16_t)(sign | 0x7C00u);
	}
	int e = (int)exp - 127 + 15;
	if (e >= 31) {
		return (uint16_t)(sign | 0x7C00u);
	}
	if (e <= 0) {
		if (e < -10) {
			return (uint16_t)sign;
		}
		mant |= 0x800000u;
		int shift = 14 - e;
		uint32_t half = mant >> shift;
		uint32_t rem = mant & ((1u << shift) - 1);
		uint32_t halfway = 1u << (shift - 1);
		if (rem > halfway || (rem == halfway && (half & 1))) {
			half++;
		}
		return (uint16_t)(sign | half);
	}
	uint32_t half = ((uint32_t)e << 10) | (mant >> 13);
	uint32_t rem = mant & 0x1FFFu;
	if (rem > 0x1000u || (rem == 0x1000u && (half & 1))) {
		half++;
	}
	return (uint16_t)(sign | half);
}

ASTC_NOINLINE float sf16_to_float(uint16_t h) {
	uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
	uint32_t exp = (h >> 10) & 0x1F;
	uint32_t mant = h & 0x3FFu;
	if (exp == 0) {
		if (mant == 0) {
			return u_as_f(sign);
		}
		// subnormal
		int e = -1;
		do {
			e++;
			mant <<= 1;
		} while (!(mant & 0x400u));
		mant &= 0x3FFu;
		return u_as_f(sign | ((uint32_t)(127 - 15 - e) << 23) | (mant << 13));
	}
	if (exp == 31) {
		return u_as_f(sign | 0x7F800000u | (mant << 13) | (mant ? 0x400000u : 0));
	}
	return u_as_f(sign | ((exp + 127 - 15) << 23) | (mant << 13));
}

// lns_to_sf16 (astcenc_vecmathlib.h:536-556)
ASTC_FN int lns_to_sf16(int p) {
	int mc = p & 0x7FF;
	int ec = (int)((unsigned int)p >> 11);
	int mt;
	if (mc < 512) {
		mt = mc * 3;
	} else if (mc < 1536) {
		mt = mc * 4 - 512;
	} else {
		mt = mc * 5 - 2048;
	}
	int res = (ec << 10) | (int)((unsigned int)mt >> 3);
	return res < 0x7BFF ? res : 0x7BFF;
}

// unorm16_to_sf16 (astcenc_vecmathlib.h:503-531)
ASTC_FN int unorm16_to_sf16(int p) {
	if (p == 0xFFFF) {
		return 0x3C00;
	}
	if (p < 4) {
		return p << 8;
	}
	// lz = clz(p) - 16 on a 32-bit value in [4, 0xFFFE]
	int lz = ASTC_CLZ((unsigned int)p) - 16;
	p = p * (1 << (lz + 1));
	p &= 0xFFFF;
	p = (int)((unsigned int)p >> 6);
	p |= (14 - lz) << 10;
	return p;
}

// float_to_lns (astcenc_vecmathlib.h:566-620)
ASTC_NOINLINE float float_to_lns(float a) {
	uint32_t ai = f_as_u(a);
	int exp = (int)((ai >> 23) & 0xFF) - 126;
	float mant = u_as_f((ai & 0x807FFFFFu) | 0x3F000000u);
	bool mask_underflow_nan = !(a > (1.0f / 67108864.0f));
	bool mask_infinity = a >= 65536.0f;
	bool exp_lt_m13 = exp < -13;
	float a1a = a * 33554432.0f;
	float a1b = (mant - 0.5f) * 4096;
	float v = exp_lt_m13 ? a1a : a1b;
	int e = exp_lt_m13 ? 0 : exp + 14;
	bool lt_384 = v < 384.0f;
	bool le_1408 = v <= 1408.0f;
	float a2a = v * (4.0f / 3.0f);
	float a2b = v + 128.0f;
	float a2c = (v + 512.0f) * (4.0f / 5.0f);
	v = a2c;
	if (le_1408) v = a2b;
	if (lt_384) v = a2a;
	v = v + (static_cast<float>(e) * 2048.0f) + 1.0f;
	if (mask_infinity) v = 65535.0f;
	if (mask_underflow_nan) v = 0.0f;
	return v;
}

