// HDR endpoint quantisers (astcenc_color_quantize.cpp:856-1906). Included by astc_dev_color.cuh.

ASTC_NOINLINE uint8_t quant_retain_top_bits(QuantCtx q, uint8_t value, int topmask) {   // :856-919
	int perform_loop;
	uint8_t quantval;
	do {
		quantval = (uint8_t)quant_color(q, value);
		perform_loop = (value & topmask) != (quantval & topmask);
		if ((quantval & topmask) > (value & topmask)) {
			value--;
		} else if ((quantval & topmask) < (value & topmask)) {
			value--;
		}
	} while (perform_loop);
	return quantval;
}

ASTC_NOINLINE void quantize_hdr_rgbo(f4 color, uint8_t output[4], QuantCtx q) {   // :925-1250
	color.x = color.x + color.w;
	color.y = color.y + color.w;
	color.z = color.z + color.w;
	color = vclamp4(0.0f, 65535.0f, color);
	f4 color_bak = color;
	int majcomp;
	if (color.x > color.y && color.x > color.z) {
		majcomp = 0;
	} else if (color.y > color.z) {
		majcomp = 1;
	} else {
		majcomp = 2;
	}
	switch (majcomp) {
	case 1: color = mk4(color.y, color.x, color.z, color.w); break;
	case 2: color = mk4(color.z, color.y, color.x, color.w); break;
	default: break;
	}
	const int mode_bits[5][3] = {{11, 5, 7}, {11, 6, 5}, {10, 5, 8}, {9, 6, 7}, {8, 7, 6}};
	const float mode_cutoffs[5][2] = {{1024, 4096}, {2048, 1024}, {2048, 16384}, {8192, 16384}, {32768, 16384}};
	const float mode_rscales[5] = {32.0f, 32.0f, 64.0f, 128.0f, 256.0f};
	const float mode_scales[5] = {1.0f / 32.0f, 1.0f / 32.0f, 1.0f / 64.0f, 1.0f / 128.0f, 1.0f / 256.0f};
	float r_base = color.x;
	float g_base = color.x - color.y;
	float b_base = color.x - color.z;
	float s_base = color.w;
	for (int mode = 0; mode < 5; mode++) {
		if (g_base > mode_cutoffs[mode][0] || b_base > mode_cutoffs[mode][0] || s_base > mode_cutoffs[mode][1]) {
			continue;
		}
		int mode_enc = mode < 4 ? (mode | (majcomp << 2)) : (majcomp | 0xC);
		float mode_scale = mode_scales[mode];
		float mode_rscale = mode_rscales[mode];
		int gb_intcutoff = 1 << mode_bits[mode][1];
		int s_intcutoff = 1 << mode_bits[mode][2];
		int r_intval = f2i_rtn(r_base * mode_scale);
		int r_lowbits = r_intval & 0x3f;
		r_lowbits |= (mode_enc & 3) << 6;
		uint8_t r_quantval = quant_retain_top_bits(q, (uint8_t)r_lowbits, 0xC0);
		r_intval = (r_intval & ~0x3f) | (r_quantval & 0x3f);
		float r_fval = static_cast<float>(r_intval) * mode_rscale;
		float g_fval = r_fval - color.y;
		float b_fval = r_fval - color.z;
		g_fval = clampf(g_fval, 0.0f, 65535.0f);
		b_fval = clampf(b_fval, 0.0f, 65535.0f);
		int g_intval = f2i_rtn(g_fval * mode_scale);
		int b_intval = f2i_rtn(b_fval * mode_scale);
		if (g_intval >= gb_intcutoff || b_intval >= gb_intcutoff) {
			continue;
		}
		int g_lowbits = g_intval & 0x1f;
		int b_lowbits = b_intval & 0x1f;
		int bit0 = 0, bit1 = 0, bit2 = 0, bit3 = 0;
		switch (mode) {
		case 0: case 2: bit0 = (r_intval >> 9) & 1; break;
		case 1: case 3: bit0 = (r_intval >> 8) & 1; break;
		case 4: case 5: bit0 = (g_intval >> 6) & 1; break;
		}
		switch (mode) {
		case 0: case 1: case 2: case 3: bit2 = (r_intval >> 7) & 1; break;
		case 4: case 5: bit2 = (b_intval >> 6) & 1; break;
		}
		switch (mode) {
		case 0: case 2: bit1 = (r_intval >> 8) & 1; break;
		case 1: case 3: case 4: case 5: bit1 = (g_intval >> 5) & 1; break;
		}
		switch (mode) {
		case 0: bit3 = (r_intval >> 10) & 1; break;
		case 2: bit3 = (r_intval >> 6) & 1; break;
		case 1: case 3: case 4: case 5: bit3 = (b_intval >> 5) & 1; break;
		}
		g_lowbits |= (mode_enc & 0x4) << 5;
		b_lowbits |= (mode_enc & 0x8) << 4;
		g_lowbits |= bit0 << 6;
		g_lowbits |= bit1 << 5;
		b_lowbits |= bit2 << 6;
		b_lowbits |= bit3 << 5;
		uint8_t g_quantval = quant_retain_top_bits(q, (uint8_t)g_lowbits, 0xF0);
		uint8_t b_quantval = quant_retain_top_bits(q, (uint8_t)b_lowbits, 0xF0);
		g_intval = (g_intval & ~0x1f) | (g_quantval & 0x1f);
		b_intval = (b_intval & ~0x1f) | (b_quantval & 0x1f);
		g_fval = static_cast<float>(g_intval) * mode_rscale;
		b_fval = static_cast<float>(b_intval) * mode_rscale;
		float rgb_errorsum = (r_fval - color.x) + (r_fval - g_fval - color.y) + (r_fval - b_fval - color.z);
		float s_fval = s_base + rgb_errorsum * (1.0f / 3.0f);
		s_fval = clampf(s_fval, 0.0f, 1e9f);
		int s_intval = f2i_rtn(s_fval * mode_scale);
		if (s_intval >= s_intcutoff) {
			continue;
		}
		int s_lowbits = s_intval & 0x1f;
		int bit4, bit5, bit6;
		switch (mode) {
		case 1: bit6 = (r_intval >> 9) & 1; break;
		default: bit6 = (s_intval >> 5) & 1; break;
		}
		switch (mode) {
		case 4: bit5 = (r_intval >> 7) & 1; break;
		case 1: bit5 = (r_intval >> 10) & 1; break;
		default: bit5 = (s_intval >> 6) & 1; break;
		}
		switch (mode) {
		case 2: bit4 = (s_intval >> 7) & 1; break;
		default: bit4 = (r_intval >> 6) & 1; break;
		}
		s_lowbits |= bit6 << 5;
		s_lowbits |= bit5 << 6;
		s_lowbits |= bit4 << 7;
		uint8_t s_quantval = quant_retain_top_bits(q, (uint8_t)s_lowbits, 0xF0);
		output[0] = r_quantval;
		output[1] = g_quantval;
		output[2] = b_quantval;
		output[3] = s_quantval;
		return;
	}
	// failed to encode with any of the modes above: encode as flat RGB (mode 5)
	float vals[4] = {color_bak.x, color_bak.y, color_bak.z, color_bak.w};
	int ivals[4];
	float cvals[3];
	for (int i = 0; i < 3; i++) {
		vals[i] = clampf(vals[i], 0.0f, 65020.0f);
		ivals[i] = f2i_rtn(vals[i] * (1.0f / 512.0f));
		cvals[i] = static_cast<float>(ivals[i]) * 512.0f;
	}
	float rgb_errorsum = (cvals[0] - vals[0]) + (cvals[1] - vals[1]) + (cvals[2] - vals[2]);
	vals[3] += rgb_errorsum * (1.0f / 3.0f);
	vals[3] = clampf(vals[3], 0.0f, 65020.0f);
	ivals[3] = f2i_rtn(vals[3] * (1.0f / 512.0f));
	int encvals[4];
	encvals[0] = (ivals[0] & 0x3f) | 0xC0;
	encvals[1] = (ivals[1] & 0x7f) | 0x80;
	encvals[2] = (ivals[2] & 0x7f) | 0x80;
	encvals[3] = (ivals[3] & 0x7f) | ((ivals[0] & 0x40) << 1);
	for (int i = 0; i < 4; i++) {
		output[i] = quant_retain_top_bits(q, (uint8_t)encvals[i], 0xF0);
	}
}

ASTC_NOINLINE void quantize_hdr_rgb(f4 color0, f4 color1, uint8_t output[6], QuantCtx q) {   // :1253-1788
	color0 = vclamp4(0.0f, 65535.0f, color0);
	color1 = vclamp4(0.0f, 65535.0f, color1);
	f4 color0_bak = color0;
	f4 color1_bak = color1;
	int majcomp;
	if (color1.x > color1.y && color1.x > color1.z) {
		majcomp = 0;
	} else if (color1.y > color1.z) {
		majcomp = 1;
	} else {
		majcomp = 2;
	}
	switch (majcomp) {
	case 1:
		color0 = mk4(color0.y, color0.x, color0.z, color0.w);
		color1 = mk4(color1.y, color1.x, color1.z, color1.w);
		break;
	case 2:
		color0 = mk4(color0.z, color0.y, color0.x, color0.w);
		color1 = mk4(color1.z, color1.y, color1.x, color1.w);
		break;
	default: break;
	}
	float a_base = color1.x;
	a_base = clampf(a_base, 0.0f, 65535.0f);
	float b0_base = a_base - color1.y;
	float b1_base = a_base - color1.z;
	float c_base = a_base - color0.x;
	float d0_base = a_base - b0_base - c_base - color0.y;
	float d1_base = a_base - b1_base - c_base - color0.z;
	const int mode_bits[8][4] = {{9, 7, 6, 7}, {9, 8, 6, 6}, {10, 6, 7, 7}, {10, 7, 7, 6},
	                                    {11, 8, 6, 5}, {11, 6, 8, 6}, {12, 7, 7, 5}, {12, 6, 7, 6}};
	const float mode_cutoffs[8][4] = {{16384, 8192, 8192, 8}, {32768, 8192, 4096, 8}, {4096, 8192, 4096, 4}, {8192, 8192, 2048, 4},
	                                         {8192, 2048, 512, 2}, {2048, 8192, 1024, 2}, {2048, 2048, 256, 1}, {1024, 2048, 512, 1}};
	const float mode_scales[8] = {1.0f / 128.0f, 1.0f / 128.0f, 1.0f / 64.0f, 1.0f / 64.0f, 1.0f / 32.0f, 1.0f / 32.0f, 1.0f / 16.0f, 1.0f / 16.0f};
	const float mode_rscales[8] = {128.0f, 128.0f, 64.0f, 64.0f, 32.0f, 32.0f, 16.0f, 16.0f};
	for (int mode = 7; mode >= 0; mode--) {
		float b_cutoff = mode_cutoffs[mode][0];
		float c_cutoff = mode_cutoffs[mode][1];
		float d_cutoff = mode_cutoffs[mode][2];
		if (b0_base > b_cutoff || b1_base > b_cutoff || c_base > c_cutoff || fabsf(d0_base) > d_cutoff || fabsf(d1_base) > d_cutoff) {
			continue;
		}
		float mode_scale = mode_scales[mode];
		float mode_rscale = mode_rscales[mode];
		int b_intcutoff = 1 << mode_bits[mode][1];
		int c_intcutoff = 1 << mode_bits[mode][2];
		int d_intcutoff = 1 << (mode_bits[mode][3] - 1);
		int a_intval = f2i_rtn(a_base * mode_scale);
		int a_lowbits = a_intval & 0xFF;
		int a_quantval = quant_color(q, a_lowbits);
		int a_uquantval = a_quantval;
		a_intval = (a_intval & ~0xFF) | a_uquantval;
		float a_fval = static_cast<float>(a_intval) * mode_rscale;
		float c_fval = a_fval - color0.x;
		c_fval = clampf(c_fval, 0.0f, 65535.0f);
		int c_intval = f2i_rtn(c_fval * mode_scale);
		if (c_intval >= c_intcutoff) {
			continue;
		}
		int c_lowbits = c_intval & 0x3f;
		c_lowbits |= (mode & 1) << 7;
		c_lowbits |= (a_intval & 0x100) >> 2;
		uint8_t c_quantval = quant_retain_top_bits(q, (uint8_t)c_lowbits, 0xC0);
		c_intval = (c_intval & ~0x3F) | (c_quantval & 0x3F);
		c_fval = static_cast<float>(c_intval) * mode_rscale;
		float b0_fval = a_fval - color1.y;
		float b1_fval = a_fval - color1.z;
		b0_fval = clampf(b0_fval, 0.0f, 65535.0f);
		b1_fval = clampf(b1_fval, 0.0f, 65535.0f);
		int b0_intval = f2i_rtn(b0_fval * mode_scale);
		int b1_intval = f2i_rtn(b1_fval * mode_scale);
		if (b0_intval >= b_intcutoff || b1_intval >= b_intcutoff) {
			continue;
		}
		int b0_lowbits = b0_intval & 0x3f;
		int b1_lowbits = b1_intval & 0x3f;
		int bit0 = 0, bit1 = 0;
		switch (mode) {
		case 0: case 1: case 3: case 4: case 6: bit0 = (b0_intval >> 6) & 1; break;
		case 2: case 5: case 7: bit0 = (a_intval >> 9) & 1; break;
		}
		switch (mode) {
		case 0: case 1: case 3: case 4: case 6: bit1 = (b1_intval >> 6) & 1; break;
		case 2: bit1 = (c_intval >> 6) & 1; break;
		case 5: case 7: bit1 = (a_intval >> 10) & 1; break;
		}
		b0_lowbits |= bit0 << 6;
		b1_lowbits |= bit1 << 6;
		b0_lowbits |= ((mode >> 1) & 1) << 7;
		b1_lowbits |= ((mode >> 2) & 1) << 7;
		uint8_t b0_quantval = quant_retain_top_bits(q, (uint8_t)b0_lowbits, 0xC0);
		uint8_t b1_quantval = quant_retain_top_bits(q, (uint8_t)b1_lowbits, 0xC0);
		b0_intval = (b0_intval & ~0x3f) | (b0_quantval & 0x3f);
		b1_intval = (b1_intval & ~0x3f) | (b1_quantval & 0x3f);
		b0_fval = static_cast<float>(b0_intval) * mode_rscale;
		b1_fval = static_cast<float>(b1_intval) * mode_rscale;
		float d0_fval = a_fval - b0_fval - c_fval - color0.y;
		float d1_fval = a_fval - b1_fval - c_fval - color0.z;
		d0_fval = clampf(d0_fval, -65535.0f, 65535.0f);
		d1_fval = clampf(d1_fval, -65535.0f, 65535.0f);
		int d0_intval = f2i_rtn(d0_fval * mode_scale);
		int d1_intval = f2i_rtn(d1_fval * mode_scale);
		if (abs(d0_intval) >= d_intcutoff || abs(d1_intval) >= d_intcutoff) {
			continue;
		}
		int d0_lowbits = d0_intval & 0x1f;
		int d1_lowbits = d1_intval & 0x1f;
		int bit2 = 0, bit3 = 0, bit4, bit5;
		switch (mode) {
		case 0: case 2: bit2 = (d0_intval >> 6) & 1; break;
		case 1: case 4: bit2 = (b0_intval >> 7) & 1; break;
		case 3: bit2 = (a_intval >> 9) & 1; break;
		case 5: bit2 = (c_intval >> 7) & 1; break;
		case 6: case 7: bit2 = (a_intval >> 11) & 1; break;
		}
		switch (mode) {
		case 0: case 2: bit3 = (d1_intval >> 6) & 1; break;
		case 1: case 4: bit3 = (b1_intval >> 7) & 1; break;
		case 3: case 5: case 6: case 7: bit3 = (c_intval >> 6) & 1; break;
		}
		switch (mode) {
		case 4: case 6:
			bit4 = (a_intval >> 9) & 1;
			bit5 = (a_intval >> 10) & 1;
			break;
		default:
			bit4 = (d0_intval >> 5) & 1;
			bit5 = (d1_intval >> 5) & 1;
			break;
		}
		d0_lowbits |= bit2 << 6;
		d1_lowbits |= bit3 << 6;
		d0_lowbits |= bit4 << 5;
		d1_lowbits |= bit5 << 5;
		d0_lowbits |= (majcomp & 1) << 7;
		d1_lowbits |= ((majcomp >> 1) & 1) << 7;
		uint8_t d0_quantval = quant_retain_top_bits(q, (uint8_t)d0_lowbits, 0xF0);
		uint8_t d1_quantval = quant_retain_top_bits(q, (uint8_t)d1_lowbits, 0xF0);
		output[0] = (uint8_t)a_quantval;
		output[1] = c_quantval;
		output[2] = b0_quantval;
		output[3] = b1_quantval;
		output[4] = d0_quantval;
		output[5] = d1_quantval;
		return;
	}
	// flat (no-submode) fallback
	float vals[6] = {color0_bak.x, color1_bak.x, color0_bak.y, color1_bak.y, color0_bak.z, color1_bak.z};
	for (int i = 0; i < 6; i++) {
		vals[i] = clampf(vals[i], 0.0f, 65020.0f);
	}
	for (int i = 0; i < 4; i++) {
		int idx = f2i_rtn(vals[i] * 1.0f / 256.0f);
		output[i] = (uint8_t)quant_color(q, idx);
	}
	for (int i = 4; i < 6; i++) {
		int idx = f2i_rtn(vals[i] * 1.0f / 512.0f) + 128;
		output[i] = quant_retain_top_bits(q, (uint8_t)idx, 0xC0);
	}
}

ASTC_FN void quantize_hdr_rgb_ldr_alpha(f4 color0, f4 color1, uint8_t output[8], QuantCtx q) {   // :1791-1817
	float scale = 1.0f / 257.0f;
	float a0 = clampf(color0.w * scale, 0.0f, 255.0f);
	float a1 = clampf(color1.w * scale, 0.0f, 255.0f);
	output[6] = (uint8_t)quant_color_f(q, f2i_rtn(a0), a0);
	output[7] = (uint8_t)quant_color_f(q, f2i_rtn(a1), a1);
	quantize_hdr_rgb(color0, color1, output, q);
}

ASTC_NOINLINE void quantize_hdr_luminance_large_range(f4 color0, f4 color1, uint8_t output[2], QuantCtx q) {   // :1659-1720
	float lum0 = hadd_rgb_s(color0) * (1.0f / 3.0f);
	float lum1 = hadd_rgb_s(color1) * (1.0f / 3.0f);
	if (lum1 < lum0) {
		float avg = (lum0 + lum1) * 0.5f;
		lum0 = avg;
		lum1 = avg;
	}
	int ilum1 = f2i_rtn(lum1);
	int ilum0 = f2i_rtn(lum0);
	int upper_v0 = (ilum0 + 128) >> 8;
	int upper_v1 = (ilum1 + 128) >> 8;
	upper_v0 = clampi(upper_v0, 0, 255);
	upper_v1 = clampi(upper_v1, 0, 255);
	int lower_v0 = (ilum1 + 256) >> 8;
	int lower_v1 = ilum0 >> 8;
	lower_v0 = clampi(lower_v0, 0, 255);
	lower_v1 = clampi(lower_v1, 0, 255);
	int upper0_dec = upper_v0 << 8;
	int upper1_dec = upper_v1 << 8;
	int lower0_dec = (lower_v1 << 8) + 128;
	int lower1_dec = (lower_v0 << 8) - 128;
	int upper0_diff = upper0_dec - ilum0;
	int upper1_diff = upper1_dec - ilum1;
	int lower0_diff = lower0_dec - ilum0;
	int lower1_diff = lower1_dec - ilum1;
	int upper_error = (upper0_diff * upper0_diff) + (upper1_diff * upper1_diff);
	int lower_error = (lower0_diff * lower0_diff) + (lower1_diff * lower1_diff);
	int v0, v1;
	if (upper_error < lower_error) {
		v0 = upper_v0;
		v1 = upper_v1;
	} else {
		v0 = lower_v0;
		v1 = lower_v1;
	}
	output[0] = (uint8_t)quant_color(q, v0);
	output[1] = (uint8_t)quant_color(q, v1);
}

ASTC_NOINLINE bool try_quantize_hdr_luminance_small_range(f4 color0, f4 color1, uint8_t output[2], QuantCtx q) {   // :1723-1817
	float lum0 = hadd_rgb_s(color0) * (1.0f / 3.0f);
	float lum1 = hadd_rgb_s(color1) * (1.0f / 3.0f);
	if (lum1 < lum0) {
		float avg = (lum0 + lum1) * 0.5f;
		lum0 = avg;
		lum1 = avg;
	}
	int ilum1 = f2i_rtn(lum1);
	int ilum0 = f2i_rtn(lum0);
	if (ilum1 - ilum0 > 2048) {
		return false;
	}
	int lowval, highval, diffval;
	int v0, v1, v0e, v1e, v0d, v1d;
	lowval = (ilum0 + 16) >> 5;
	highval = (ilum1 + 16) >> 5;
	lowval = clampi(lowval, 0, 2047);
	highval = clampi(highval, 0, 2047);
	v0 = lowval & 0x7F;
	v0e = quant_color(q, v0);
	v0d = v0e;
	if (v0d < 0x80) {
		lowval = (lowval & ~0x7F) | v0d;
		diffval = highval - lowval;
		if (diffval >= 0 && diffval <= 15) {
			v1 = ((lowval >> 3) & 0xF0) | diffval;
			v1e = quant_color(q, v1);
			v1d = v1e;
			if ((v1d & 0xF0) == (v1 & 0xF0)) {
				output[0] = (uint8_t)v0e;
				output[1] = (uint8_t)v1e;
				return true;
			}
		}
	}
	lowval = (ilum0 + 32) >> 6;
	highval = (ilum1 + 32) >> 6;
	lowval = clampi(lowval, 0, 1023);
	highval = clampi(highval, 0, 1023);
	v0 = (lowval & 0x7F) | 0x80;
	v0e = quant_color(q, v0);
	v0d = v0e;
	if ((v0d & 0x80) == 0) {
		return false;
	}
	lowval = (lowval & ~0x7F) | (v0d & 0x7F);
	diffval = highval - lowval;
	if (diffval < 0 || diffval > 31) {
		return false;
	}
	v1 = ((lowval >> 2) & 0xE0) | diffval;
	v1e = quant_color(q, v1);
	v1d = v1e;
	if ((v1d & 0xE0) != (v1 & 0xE0)) {
		return false;
	}
	output[0] = (uint8_t)v0e;
	output[1] = (uint8_t)v1e;
	return true;
}

ASTC_NOINLINE void quantize_hdr_alpha(float alpha0, float alpha1, uint8_t output[2], QuantCtx q) {   // :1820-1890
	alpha0 = clampf(alpha0, 0.0f, 65280.0f);
	alpha1 = clampf(alpha1, 0.0f, 65280.0f);
	int ialpha0 = f2i_rtn(alpha0);
	int ialpha1 = f2i_rtn(alpha1);
	int val0, val1, diffval;
	int v6, v7, v6e, v7e, v6d, v7d;
	for (int i = 2; i >= 0; i--) {
		val0 = (ialpha0 + (128 >> i)) >> (8 - i);
		val1 = (ialpha1 + (128 >> i)) >> (8 - i);
		v6 = (val0 & 0x7F) | ((i & 1) << 7);
		v6e = quant_color(q, v6);
		v6d = v6e;
		if ((v6 ^ v6d) & 0x80) {
			continue;
		}
		val0 = (val0 & ~0x7f) | (v6d & 0x7f);
		diffval = val1 - val0;
		int cutoff = 32 >> i;
		int mask = 2 * cutoff - 1;
		if (diffval < -cutoff || diffval >= cutoff) {
			continue;
		}
		v7 = ((i & 2) << 6) | ((val0 >> 7) << (6 - i)) | (diffval & mask);
		v7e = quant_color(q, v7);
		v7d = v7e;
		const int testbits[3] = {0xE0, 0xF0, 0xF8};
		if ((v7 ^ v7d) & testbits[i]) {
			continue;
		}
		output[0] = (uint8_t)v6e;
		output[1] = (uint8_t)v7e;
		return;
	}
	val0 = (ialpha0 + 256) >> 9;
	val1 = (ialpha1 + 256) >> 9;
	v6 = val0 | 0x80;
	v7 = val1 | 0x80;
	output[0] = (uint8_t)quant_color(q, v6);
	output[1] = (uint8_t)quant_color(q, v7);
}

ASTC_FN void quantize_hdr_rgb_alpha(f4 color0, f4 color1, uint8_t output[8], QuantCtx q) {   // :1893-1906
	quantize_hdr_rgb(color0, color1, output, q);
	quantize_hdr_alpha(color0.w, color1.w, output + 6, q);
}
