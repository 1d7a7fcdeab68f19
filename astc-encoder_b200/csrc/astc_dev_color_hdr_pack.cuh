// HDR endpoint quantisers, lane-parallel (what astcenc_color_quantize.cpp:856-1906 computes). Included by astc_dev_color.cuh.
//
// The HDR endpoint formats are "try sub-mode after sub-mode until one can represent the value" ladders: 5 sub-modes + a
// flat fallback for RGB+offset (format 7), 8 + fallback for RGB (format 11), 3 + fallback for the alpha pair, 2 + the large
// range form for luminance. The sub-modes are independent of each other, only their PRIORITY matters. So the warp tries
// them all at once - attempt k on lane k, every lane running the same code on per-mode parameters it unpacks from small
// packed tables (shift amounts, field widths, and a routing table that says which bit of which intermediate goes into
// which spare bit of the output bytes) - and the attempt with the highest priority that succeeded
// (__ffs(__ballot_sync(ok))) stores its bytes. Cost = one attempt instead of up to nine in a row on a single lane.
//
// Every attempt is the arithmetic the format defines (scale by 2^-k, round to nearest, keep the top bits through the
// quantiser) in the order the reference evaluates it - the results are the reference's bytes; only who computes what differs.

// result of one attempt: up to 8 output bytes (little endian in two words) and whether the sub-mode could hold the value
struct HdrTry {
	uint32_t lo, hi;
	bool ok;
};
ASTC_FN uint32_t hdr_b4(int a, int b, int c, int d) { return (uint32_t)(a & 0xFF) | ((uint32_t)(b & 0xFF) << 8) | ((uint32_t)(c & 0xFF) << 16) | ((uint32_t)(d & 0xFF) << 24); }

// lowest lane whose attempt succeeded (-1: none); one-lane builds pass their single attempt through
ASTC_FN int hdr_first_ok(bool ok, int lane) {
#if defined(ASTC_ONE_LANE)
	return ok ? lane : -1;
#else
	(void)lane;
	unsigned int m = __ballot_sync(0xffffffffu, ok);
	return m ? __ffs((int)m) - 1 : -1;
#endif
}

// Run attempts 0..n-1 (priority order) over the lanes, store the bytes of the first that succeeds; the last one never fails.
template <typename TryFn>
ASTC_FN void hdr_try_all(int lane, int n, int nbytes, uint8_t* output, TryFn attempt) {
	int base = 0;
	ASTC_NOUNROLL
	while (base < n) {
		int k = base + lane;
		HdrTry t;
		t.lo = t.hi = 0;
		t.ok = false;
		if (k < n) {
			t = attempt(k);
		}
		int win = hdr_first_ok(t.ok, lane);
		if (win >= 0) {
			if (lane == win) {
				for (int i = 0; i < nbytes; i++) {
					output[i] = (uint8_t)((i < 4 ? t.lo >> (8 * i) : t.hi >> (8 * (i - 4))) & 0xFF);
				}
			}
			return;
		}
		base += ASTC_WARP;
	}
}

// quantise `value` so that the bits in `topmask` survive the round trip: step the input down until they do (:856-919)
ASTC_NOINLINE uint8_t quant_retain_top_bits(QuantCtx q, uint8_t value, int topmask) {
	while (true) {
		uint8_t quantval = (uint8_t)quant_color(q, value);
		if (((value ^ quantval) & topmask) == 0) {
			return quantval;
		}
		value--;
	}
}

// one routed bit: descriptor = (source index << 4) | bit position, sources in src[]
ASTC_FN int hdr_route(uint32_t desc, const int src[4]) {
	return (src[(desc >> 4) & 3] >> (desc & 15)) & 1;
}
#define HDR_NIB(word, k) ((int)(((word) >> (4 * (k))) & 0xF))

// ---- format 7: RGB + offset ("RGBO") --------------------------------------------------------------------------------
// Sub-mode m = 0..4: the major component is kept with 9..11 bits, the two others as differences to it, the offset with
// 5..8 bits, all scaled by 2^-sh. Per-mode parameters (nibble m of each word):
//   sh       scale shift                     5 5 6 7 8
//   gb_bits  width of the difference fields  5 6 5 6 7
//   s_bits   width of the offset field       7 5 8 7 6
//   cut0/1   value ranges: differences <= 1 << cut0, offset <= 1 << cut1
// Routing (source 0 = major, 1 / 2 = differences, 3 = offset; entry k = spare bit k): which bit the sub-mode parks where.
ASTC_FN HdrTry hdr_rgbo_attempt(int m, f4 color, f4 color_bak, int majcomp, QuantCtx q) {
	HdrTry out;
	out.lo = out.hi = 0;
	out.ok = false;
	if (m == 5) {
		// flat form: 7 bits per component at 1/512 (:1221-1250)
		float v0 = clampf(color_bak.x, 0.0f, 65020.0f), v1 = clampf(color_bak.y, 0.0f, 65020.0f), v2 = clampf(color_bak.z, 0.0f, 65020.0f);
		int i0 = f2i_rtn(v0 * (1.0f / 512.0f)), i1 = f2i_rtn(v1 * (1.0f / 512.0f)), i2 = f2i_rtn(v2 * (1.0f / 512.0f));
		float c0 = static_cast<float>(i0) * 512.0f, c1 = static_cast<float>(i1) * 512.0f, c2 = static_cast<float>(i2) * 512.0f;
		float rgb_errorsum = (c0 - v0) + (c1 - v1) + (c2 - v2);
		float v3 = color_bak.w + rgb_errorsum * (1.0f / 3.0f);
		v3 = clampf(v3, 0.0f, 65020.0f);
		int i3 = f2i_rtn(v3 * (1.0f / 512.0f));
		out.lo = hdr_b4(quant_retain_top_bits(q, (uint8_t)((i0 & 0x3f) | 0xC0), 0xF0), quant_retain_top_bits(q, (uint8_t)((i1 & 0x7f) | 0x80), 0xF0),
		                quant_retain_top_bits(q, (uint8_t)((i2 & 0x7f) | 0x80), 0xF0), quant_retain_top_bits(q, (uint8_t)((i3 & 0x7f) | ((i0 & 0x40) << 1)), 0xF0));
		out.ok = true;
		return out;
	}
	const uint32_t SH = 0x87655u, GB_BITS = 0x76565u, S_BITS = 0x67857u, CUT0 = 0xFDBBAu, CUT1 = 0xEEEACu;
	// routing of spare bits 0..6 per sub-mode
	const uint8_t ROUTE[5][7] = {
		{0x09, 0x08, 0x07, 0x0A, 0x06, 0x36, 0x35},
		{0x08, 0x15, 0x07, 0x25, 0x06, 0x0A, 0x09},
		{0x09, 0x08, 0x07, 0x06, 0x37, 0x36, 0x35},
		{0x08, 0x15, 0x07, 0x25, 0x06, 0x36, 0x35},
		{0x16, 0x15, 0x26, 0x25, 0x06, 0x07, 0x35}};
	float r_base = color.x;
	float g_base = color.x - color.y;
	float b_base = color.x - color.z;
	float s_base = color.w;
	float cut0 = static_cast<float>(1 << HDR_NIB(CUT0, m)), cut1 = static_cast<float>(1 << HDR_NIB(CUT1, m));
	if (g_base > cut0 || b_base > cut0 || s_base > cut1) {
		return out;
	}
	int sh = HDR_NIB(SH, m);
	float mode_rscale = static_cast<float>(1 << sh);
	float mode_scale = 1.0f / mode_rscale;
	int mode_enc = m < 4 ? (m | (majcomp << 2)) : (majcomp | 0xC);
	int src[4];
	// major component: low 6 bits + two mode bits on top
	int r_intval = f2i_rtn(r_base * mode_scale);
	uint8_t r_quantval = quant_retain_top_bits(q, (uint8_t)((r_intval & 0x3f) | ((mode_enc & 3) << 6)), 0xC0);
	r_intval = (r_intval & ~0x3f) | (r_quantval & 0x3f);
	float r_fval = static_cast<float>(r_intval) * mode_rscale;
	// the two differences against the quantised major component
	float g_fval = clampf(r_fval - color.y, 0.0f, 65535.0f);
	float b_fval = clampf(r_fval - color.z, 0.0f, 65535.0f);
	int g_intval = f2i_rtn(g_fval * mode_scale);
	int b_intval = f2i_rtn(b_fval * mode_scale);
	int gb_limit = 1 << HDR_NIB(GB_BITS, m);
	if (g_intval >= gb_limit || b_intval >= gb_limit) {
		return out;
	}
	src[0] = r_intval; src[1] = g_intval; src[2] = b_intval; src[3] = 0;
	int g_field = (g_intval & 0x1f) | ((mode_enc & 0x4) << 5) | (hdr_route(ROUTE[m][0], src) << 6) | (hdr_route(ROUTE[m][1], src) << 5);
	int b_field = (b_intval & 0x1f) | ((mode_enc & 0x8) << 4) | (hdr_route(ROUTE[m][2], src) << 6) | (hdr_route(ROUTE[m][3], src) << 5);
	uint8_t g_quantval = quant_retain_top_bits(q, (uint8_t)g_field, 0xF0);
	uint8_t b_quantval = quant_retain_top_bits(q, (uint8_t)b_field, 0xF0);
	g_intval = (g_intval & ~0x1f) | (g_quantval & 0x1f);
	b_intval = (b_intval & ~0x1f) | (b_quantval & 0x1f);
	g_fval = static_cast<float>(g_intval) * mode_rscale;
	b_fval = static_cast<float>(b_intval) * mode_rscale;
	// the offset absorbs the mean error made on the colour
	float rgb_errorsum = (r_fval - color.x) + (r_fval - g_fval - color.y) + (r_fval - b_fval - color.z);
	float s_fval = clampf(s_base + rgb_errorsum * (1.0f / 3.0f), 0.0f, 1e9f);
	int s_intval = f2i_rtn(s_fval * mode_scale);
	if (s_intval >= (1 << HDR_NIB(S_BITS, m))) {
		return out;
	}
	src[3] = s_intval;
	int s_field = (s_intval & 0x1f) | (hdr_route(ROUTE[m][6], src) << 5) | (hdr_route(ROUTE[m][5], src) << 6) | (hdr_route(ROUTE[m][4], src) << 7);
	uint8_t s_quantval = quant_retain_top_bits(q, (uint8_t)s_field, 0xF0);
	out.lo = hdr_b4(r_quantval, g_quantval, b_quantval, s_quantval);
	out.ok = true;
	return out;
}

ASTC_NOINLINE void quantize_hdr_rgbo(int lane, f4 color, uint8_t output[4], QuantCtx q) {   // :925-1250
	color.x = color.x + color.w;
	color.y = color.y + color.w;
	color.z = color.z + color.w;
	color = vclamp4(0.0f, 65535.0f, color);
	f4 color_bak = color;
	// the largest component leads; the others are coded against it
	int majcomp = (color.x > color.y && color.x > color.z) ? 0 : (color.y > color.z ? 1 : 2);
	if (majcomp == 1) {
		color = mk4(color.y, color.x, color.z, color.w);
	} else if (majcomp == 2) {
		color = mk4(color.z, color.y, color.x, color.w);
	}
	hdr_try_all(lane, 6, 4, output, [&](int k) { return hdr_rgbo_attempt(k, color, color_bak, majcomp, q); });
}

// ---- format 11: RGB, two endpoints ------------------------------------------------------------------------------------
// Value a = the major component of endpoint 1, b0 / b1 = its distance to the two other components of endpoint 1, c = its
// distance to the major component of endpoint 0, d0 / d1 = what is left for endpoint 0's other components (signed).
// Sub-modes 7 (finest, smallest range) .. 0; attempt k tries sub-mode 7 - k, attempt 8 is the flat form. Per sub-mode:
//   sh scale shift 7 7 6 6 5 5 4 4; widths of the b / c / d fields (the a field is 9..12 bits); log2 of the three value ranges.
// Routing sources: 0 = a, 1 = b0 or b1 / d0 or d1 (by slot), 2 = c, 3 = the other of the pair - see the table.
ASTC_FN HdrTry hdr_rgb_attempt(int k, f4 color0, f4 color1, f4 color0_bak, f4 color1_bak, int majcomp, QuantCtx q) {
	HdrTry out;
	out.lo = out.hi = 0;
	out.ok = false;
	if (k == 8) {
		// flat form: 8 bits per component at 1/256, the last pair at 1/512 with the format's marker bit (:1777-1788)
		float v[6] = {color0_bak.x, color1_bak.x, color0_bak.y, color1_bak.y, color0_bak.z, color1_bak.z};
		int o[6];
		for (int i = 0; i < 6; i++) {
			float c = clampf(v[i], 0.0f, 65020.0f);
			o[i] = i < 4 ? quant_color(q, f2i_rtn(c * 1.0f / 256.0f)) : quant_retain_top_bits(q, (uint8_t)(f2i_rtn(c * 1.0f / 512.0f) + 128), 0xC0);
		}
		out.lo = hdr_b4(o[0], o[1], o[2], o[3]);
		out.hi = hdr_b4(o[4], o[5], 0, 0);
		out.ok = true;
		return out;
	}
	int mode = 7 - k;
	// nibble `mode` of each word
	const uint32_t SH = 0x44556677u, B_BITS = 0x67687687u, C_BITS = 0x77867766u, D_BITS = 0x65656767u;
	const uint32_t CUTB = 0xABBDDCFEu, CUTC = 0xBBDBDDDDu, CUTD = 0x98A9BCCDu;
	// spare bits 0..5: sources 0 a, 1 b0, 2 b1, 3 c, 4 d0, 5 d1 (3 bits) | position (4 bits)
	const uint8_t ROUTE[8][6] = {
		{0x16, 0x26, 0x46, 0x56, 0x45, 0x55},
		{0x16, 0x26, 0x17, 0x27, 0x45, 0x55},
		{0x09, 0x36, 0x46, 0x56, 0x45, 0x55},
		{0x16, 0x26, 0x09, 0x36, 0x45, 0x55},
		{0x16, 0x26, 0x17, 0x27, 0x09, 0x0A},
		{0x09, 0x0A, 0x37, 0x36, 0x45, 0x55},
		{0x16, 0x26, 0x0B, 0x36, 0x09, 0x0A},
		{0x09, 0x0A, 0x0B, 0x36, 0x45, 0x55}};
	float a_base = clampf(color1.x, 0.0f, 65535.0f);
	float b0_base = a_base - color1.y;
	float b1_base = a_base - color1.z;
	float c_base = a_base - color0.x;
	float d0_base = a_base - b0_base - c_base - color0.y;
	float d1_base = a_base - b1_base - c_base - color0.z;
	float b_cutoff = static_cast<float>(1 << HDR_NIB(CUTB, mode));
	float c_cutoff = static_cast<float>(1 << HDR_NIB(CUTC, mode));
	float d_cutoff = static_cast<float>(1 << HDR_NIB(CUTD, mode));
	if (b0_base > b_cutoff || b1_base > b_cutoff || c_base > c_cutoff || fabsf(d0_base) > d_cutoff || fabsf(d1_base) > d_cutoff) {
		return out;
	}
	int sh = HDR_NIB(SH, mode);
	float mode_rscale = static_cast<float>(1 << sh);
	float mode_scale = 1.0f / mode_rscale;
	int v[6];      // a, b0, b1, c, d0, d1 as scaled integers
	// a: its low 8 bits go through the quantiser as they are
	int a_intval = f2i_rtn(a_base * mode_scale);
	int a_quantval = quant_color(q, a_intval & 0xFF);
	a_intval = (a_intval & ~0xFF) | a_quantval;
	float a_fval = static_cast<float>(a_intval) * mode_rscale;
	// c against the quantised a
	float c_fval = clampf(a_fval - color0.x, 0.0f, 65535.0f);
	int c_intval = f2i_rtn(c_fval * mode_scale);
	if (c_intval >= (1 << HDR_NIB(C_BITS, mode))) {
		return out;
	}
	uint8_t c_quantval = quant_retain_top_bits(q, (uint8_t)((c_intval & 0x3f) | ((mode & 1) << 7) | ((a_intval & 0x100) >> 2)), 0xC0);
	c_intval = (c_intval & ~0x3F) | (c_quantval & 0x3F);
	c_fval = static_cast<float>(c_intval) * mode_rscale;
	// b0, b1 against the quantised a
	float b0_fval = clampf(a_fval - color1.y, 0.0f, 65535.0f);
	float b1_fval = clampf(a_fval - color1.z, 0.0f, 65535.0f);
	int b0_intval = f2i_rtn(b0_fval * mode_scale);
	int b1_intval = f2i_rtn(b1_fval * mode_scale);
	int b_limit = 1 << HDR_NIB(B_BITS, mode);
	if (b0_intval >= b_limit || b1_intval >= b_limit) {
		return out;
	}
	v[0] = a_intval; v[1] = b0_intval; v[2] = b1_intval; v[3] = c_intval; v[4] = 0; v[5] = 0;
	auto route = [&](int slot) { uint32_t d = ROUTE[mode][slot]; return (v[(d >> 4) & 7] >> (d & 15)) & 1; };
	uint8_t b0_quantval = quant_retain_top_bits(q, (uint8_t)((b0_intval & 0x3f) | (route(0) << 6) | (((mode >> 1) & 1) << 7)), 0xC0);
	uint8_t b1_quantval = quant_retain_top_bits(q, (uint8_t)((b1_intval & 0x3f) | (route(1) << 6) | (((mode >> 2) & 1) << 7)), 0xC0);
	b0_intval = (b0_intval & ~0x3f) | (b0_quantval & 0x3f);
	b1_intval = (b1_intval & ~0x3f) | (b1_quantval & 0x3f);
	b0_fval = static_cast<float>(b0_intval) * mode_rscale;
	b1_fval = static_cast<float>(b1_intval) * mode_rscale;
	// d0, d1: what the quantised a, b, c leave over (signed)
	float d0_fval = clampf(a_fval - b0_fval - c_fval - color0.y, -65535.0f, 65535.0f);
	float d1_fval = clampf(a_fval - b1_fval - c_fval - color0.z, -65535.0f, 65535.0f);
	int d0_intval = f2i_rtn(d0_fval * mode_scale);
	int d1_intval = f2i_rtn(d1_fval * mode_scale);
	int d_limit = 1 << (HDR_NIB(D_BITS, mode) - 1);
	if (abs(d0_intval) >= d_limit || abs(d1_intval) >= d_limit) {
		return out;
	}
	v[1] = b0_intval; v[2] = b1_intval; v[4] = d0_intval; v[5] = d1_intval;
	int d0_field = (d0_intval & 0x1f) | (route(2) << 6) | (route(4) << 5) | ((majcomp & 1) << 7);
	int d1_field = (d1_intval & 0x1f) | (route(3) << 6) | (route(5) << 5) | (((majcomp >> 1) & 1) << 7);
	uint8_t d0_quantval = quant_retain_top_bits(q, (uint8_t)d0_field, 0xF0);
	uint8_t d1_quantval = quant_retain_top_bits(q, (uint8_t)d1_field, 0xF0);
	out.lo = hdr_b4(a_quantval, c_quantval, b0_quantval, b1_quantval);
	out.hi = hdr_b4(d0_quantval, d1_quantval, 0, 0);
	out.ok = true;
	return out;
}

ASTC_NOINLINE void quantize_hdr_rgb(int lane, f4 color0, f4 color1, uint8_t output[6], QuantCtx q) {   // :1253-1788
	color0 = vclamp4(0.0f, 65535.0f, color0);
	color1 = vclamp4(0.0f, 65535.0f, color1);
	f4 color0_bak = color0;
	f4 color1_bak = color1;
	int majcomp = (color1.x > color1.y && color1.x > color1.z) ? 0 : (color1.y > color1.z ? 1 : 2);
	if (majcomp == 1) {
		color0 = mk4(color0.y, color0.x, color0.z, color0.w);
		color1 = mk4(color1.y, color1.x, color1.z, color1.w);
	} else if (majcomp == 2) {
		color0 = mk4(color0.z, color0.y, color0.x, color0.w);
		color1 = mk4(color1.z, color1.y, color1.x, color1.w);
	}
	hdr_try_all(lane, 9, 6, output, [&](int k) { return hdr_rgb_attempt(k, color0, color1, color0_bak, color1_bak, majcomp, q); });
}

// ---- HDR alpha pair (:1820-1890): precision levels 2, 1, 0 (attempts 0..2), then the flat 7-bit form ----------------------
ASTC_FN HdrTry hdr_alpha_attempt(int k, int ialpha0, int ialpha1, QuantCtx q) {
	HdrTry out;
	out.lo = out.hi = 0;
	out.ok = false;
	if (k == 3) {
		out.lo = hdr_b4(quant_color(q, ((ialpha0 + 256) >> 9) | 0x80), quant_color(q, ((ialpha1 + 256) >> 9) | 0x80), 0, 0);
		out.ok = true;
		return out;
	}
	int i = 2 - k;
	int val0 = (ialpha0 + (128 >> i)) >> (8 - i);
	int val1 = (ialpha1 + (128 >> i)) >> (8 - i);
	int v6 = (val0 & 0x7F) | ((i & 1) << 7);
	int v6e = quant_color(q, v6);
	if ((v6 ^ v6e) & 0x80) {
		return out;
	}
	val0 = (val0 & ~0x7f) | (v6e & 0x7f);
	int diffval = val1 - val0;
	int cutoff = 32 >> i;
	if (diffval < -cutoff || diffval >= cutoff) {
		return out;
	}
	int v7 = ((i & 2) << 6) | ((val0 >> 7) << (6 - i)) | (diffval & (2 * cutoff - 1));
	int v7e = quant_color(q, v7);
	int keep = i == 0 ? 0xE0 : (i == 1 ? 0xF0 : 0xF8);      // the bits of the second byte that carry the base
	if ((v7 ^ v7e) & keep) {
		return out;
	}
	out.lo = hdr_b4(v6e, v7e, 0, 0);
	out.ok = true;
	return out;
}

ASTC_NOINLINE void quantize_hdr_alpha(int lane, float alpha0, float alpha1, uint8_t output[2], QuantCtx q) {
	int ialpha0 = f2i_rtn(clampf(alpha0, 0.0f, 65280.0f));
	int ialpha1 = f2i_rtn(clampf(alpha1, 0.0f, 65280.0f));
	hdr_try_all(lane, 4, 2, output, [&](int k) { return hdr_alpha_attempt(k, ialpha0, ialpha1, q); });
}

// ---- HDR luminance (:1659-1817): small range at 1/32 (attempt 0) or 1/64 (attempt 1), else the large range form (2) -------
ASTC_FN HdrTry hdr_luminance_attempt(int k, int ilum0, int ilum1, QuantCtx q) {
	HdrTry out;
	out.lo = out.hi = 0;
	out.ok = false;
	if (k == 2) {
		// large range: both ends at 1/256, rounding the pair up or down together - whichever lands closer
		int upper_v0 = clampi((ilum0 + 128) >> 8, 0, 255), upper_v1 = clampi((ilum1 + 128) >> 8, 0, 255);
		int lower_v0 = clampi((ilum1 + 256) >> 8, 0, 255), lower_v1 = clampi(ilum0 >> 8, 0, 255);
		int u0 = (upper_v0 << 8) - ilum0, u1 = (upper_v1 << 8) - ilum1;
		int l0 = ((lower_v1 << 8) + 128) - ilum0, l1 = ((lower_v0 << 8) - 128) - ilum1;
		bool upper = (u0 * u0 + u1 * u1) < (l0 * l0 + l1 * l1);
		out.lo = hdr_b4(quant_color(q, upper ? upper_v0 : lower_v0), quant_color(q, upper ? upper_v1 : lower_v1), 0, 0);
		out.ok = true;
		return out;
	}
	if (ilum1 - ilum0 > 2048) {
		return out;
	}
	// k = 0: 11-bit base, 4-bit difference; k = 1: 10-bit base, 5-bit difference, marker bit set
	int sh = 5 + k, top = k == 0 ? 2047 : 1023, dmax = k == 0 ? 15 : 31;
	int lowval = clampi((ilum0 + (16 << k)) >> sh, 0, top);
	int highval = clampi((ilum1 + (16 << k)) >> sh, 0, top);
	int v0 = (lowval & 0x7F) | (k << 7);
	int v0e = quant_color(q, v0);
	if ((v0e & 0x80) != (k << 7)) {
		return out;
	}
	lowval = (lowval & ~0x7F) | (v0e & 0x7F);
	int diffval = highval - lowval;
	if (diffval < 0 || diffval > dmax) {
		return out;
	}
	int keep = k == 0 ? 0xF0 : 0xE0;
	int v1 = ((lowval >> (3 - k)) & keep) | diffval;
	int v1e = quant_color(q, v1);
	if ((v1e & keep) != (v1 & keep)) {
		return out;
	}
	out.lo = hdr_b4(v0e, v1e, 0, 0);
	out.ok = true;
	return out;
}

// returns the format that was used (small range when one of its two forms fits)
ASTC_NOINLINE int quantize_hdr_luminance(int lane, f4 color0, f4 color1, uint8_t output[2], QuantCtx q) {
	float lum0 = hadd_rgb_s(color0) * (1.0f / 3.0f);
	float lum1 = hadd_rgb_s(color1) * (1.0f / 3.0f);
	if (lum1 < lum0) {
		float avg = (lum0 + lum1) * 0.5f;
		lum0 = avg;
		lum1 = avg;
	}
	int ilum1 = f2i_rtn(lum1);
	int ilum0 = f2i_rtn(lum0);
	// (every lane can tell which attempt wins from the flags alone: recompute the winner index for the return value)
	int base = 0, win = -1;
	ASTC_NOUNROLL
	while (win < 0) {
		int k = base + lane;
		HdrTry t;
		t.lo = t.hi = 0;
		t.ok = false;
		if (k < 3) {
			t = hdr_luminance_attempt(k, ilum0, ilum1, q);
		}
		int wl = hdr_first_ok(t.ok, lane);
		if (wl >= 0) {
			win = base + wl;
			if (lane == wl) {
				output[0] = (uint8_t)(t.lo & 0xFF);
				output[1] = (uint8_t)((t.lo >> 8) & 0xFF);
			}
		}
		base += ASTC_WARP;
	}
	return win < 2 ? FMT_HDR_LUMINANCE_SMALL_RANGE : FMT_HDR_LUMINANCE_LARGE_RANGE;
}
