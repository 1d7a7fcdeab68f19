// HDR endpoint unpackers (astcenc_color_unquantize.cpp:310-841). Included by astc_dev_color.cuh.

ASTC_FN int safe_signed_lsh(int val, int shift) {
	return (int)((unsigned int)val << shift);
}

ASTC_NOINLINE void hdr_rgbo_unpack(const uint8_t in[4], i4& o0, i4& o1) {   // :310-500
	int v0 = in[0], v1 = in[1], v2 = in[2], v3 = in[3];
	int modeval = ((v0 & 0xC0) >> 6) | (((v1 & 0x80) >> 7) << 2) | (((v2 & 0x80) >> 7) << 3);
	int majcomp, mode;
	if ((modeval & 0xC) != 0xC) {
		majcomp = modeval >> 2;
		mode = modeval & 3;
	} else if (modeval != 0xF) {
		majcomp = modeval & 3;
		mode = 4;
	} else {
		majcomp = 0;
		mode = 5;
	}
	int red = v0 & 0x3F, green = v1 & 0x1F, blue = v2 & 0x1F, scale = v3 & 0x1F;
	int bit0 = (v1 >> 6) & 1, bit1 = (v1 >> 5) & 1, bit2 = (v2 >> 6) & 1, bit3 = (v2 >> 5) & 1;
	int bit4 = (v3 >> 7) & 1, bit5 = (v3 >> 6) & 1, bit6 = (v3 >> 5) & 1;
	int oh = 1 << mode;
	if (oh & 0x30) green |= bit0 << 6;
	if (oh & 0x3A) green |= bit1 << 5;
	if (oh & 0x30) blue |= bit2 << 6;
	if (oh & 0x3A) blue |= bit3 << 5;
	if (oh & 0x3D) scale |= bit6 << 5;
	if (oh & 0x2D) scale |= bit5 << 6;
	if (oh & 0x04) scale |= bit4 << 7;
	if (oh & 0x3B) red |= bit4 << 6;
	if (oh & 0x04) red |= bit3 << 6;
	if (oh & 0x10) red |= bit5 << 7;
	if (oh & 0x0F) red |= bit2 << 7;
	if (oh & 0x05) red |= bit1 << 8;
	if (oh & 0x0A) red |= bit0 << 8;
	if (oh & 0x05) red |= bit0 << 9;
	if (oh & 0x02) red |= bit6 << 9;
	if (oh & 0x01) red |= bit3 << 10;
	if (oh & 0x02) red |= bit5 << 10;
	const int shamts[6] = {1, 1, 2, 3, 4, 5};
	int shamt = shamts[mode];
	red <<= shamt;
	green <<= shamt;
	blue <<= shamt;
	scale <<= shamt;
	if (mode != 5) {
		green = red - green;
		blue = red - blue;
	}
	int temp;
	switch (majcomp) {
	case 1: temp = red; red = green; green = temp; break;
	case 2: temp = red; red = blue; blue = temp; break;
	default: break;
	}
	int red0 = red - scale, green0 = green - scale, blue0 = blue - scale;
	red = maxi(red, 0);
	green = maxi(green, 0);
	blue = maxi(blue, 0);
	red0 = maxi(red0, 0);
	green0 = maxi(green0, 0);
	blue0 = maxi(blue0, 0);
	o0 = mki4(red0 << 4, green0 << 4, blue0 << 4, 0x7800);
	o1 = mki4(red << 4, green << 4, blue << 4, 0x7800);
}

ASTC_NOINLINE void hdr_rgb_unpack(const uint8_t in[6], i4& o0, i4& o1) {   // :503-700
	int v0 = in[0], v1 = in[1], v2 = in[2], v3 = in[3], v4 = in[4], v5 = in[5];
	int modeval = ((v1 & 0x80) >> 7) | (((v2 & 0x80) >> 7) << 1) | (((v3 & 0x80) >> 7) << 2);
	int majcomp = ((v4 & 0x80) >> 7) | (((v5 & 0x80) >> 7) << 1);
	if (majcomp == 3) {
		o0 = mki4(v0 << 8, v2 << 8, (v4 & 0x7F) << 9, 0x7800);
		o1 = mki4(v1 << 8, v3 << 8, (v5 & 0x7F) << 9, 0x7800);
		return;
	}
	int a = v0 | ((v1 & 0x40) << 2);
	int b0 = v2 & 0x3f, b1 = v3 & 0x3f, c = v1 & 0x3f, d0 = v4 & 0x7f, d1 = v5 & 0x7f;
	const int dbits_tab[8] = {7, 6, 7, 6, 5, 6, 5, 6};
	int dbits = dbits_tab[modeval];
	int bit0 = (v2 >> 6) & 1, bit1 = (v3 >> 6) & 1, bit2 = (v4 >> 6) & 1, bit3 = (v5 >> 6) & 1;
	int bit4 = (v4 >> 5) & 1, bit5 = (v5 >> 5) & 1;
	int oh = 1 << modeval;
	if (oh & 0xA4) a |= bit0 << 9;
	if (oh & 0x8) a |= bit2 << 9;
	if (oh & 0x50) a |= bit4 << 9;
	if (oh & 0x50) a |= bit5 << 10;
	if (oh & 0xA0) a |= bit1 << 10;
	if (oh & 0xC0) a |= bit2 << 11;
	if (oh & 0x4) c |= bit1 << 6;
	if (oh & 0xE8) c |= bit3 << 6;
	if (oh & 0x20) c |= bit2 << 7;
	if (oh & 0x5B) { b0 |= bit0 << 6; b1 |= bit1 << 6; }
	if (oh & 0x12) { b0 |= bit2 << 7; b1 |= bit3 << 7; }
	if (oh & 0xAF) { d0 |= bit4 << 5; d1 |= bit5 << 5; }
	if (oh & 0x5) { d0 |= bit2 << 6; d1 |= bit3 << 6; }
	int sx = 32 - dbits;
	int d0x = safe_signed_lsh(d0, sx) >> sx;
	int d1x = safe_signed_lsh(d1, sx) >> sx;
	d0 = d0x;
	d1 = d1x;
	int vs = (modeval >> 1) ^ 3;
	a = safe_signed_lsh(a, vs);
	b0 = safe_signed_lsh(b0, vs);
	b1 = safe_signed_lsh(b1, vs);
	c = safe_signed_lsh(c, vs);
	d0 = safe_signed_lsh(d0, vs);
	d1 = safe_signed_lsh(d1, vs);
	int red1 = a, green1 = a - b0, blue1 = a - b1;
	int red0 = a - c, green0 = a - b0 - c - d0, blue0 = a - b1 - c - d1;
	red0 = clampi(red0, 0, 4095);
	green0 = clampi(green0, 0, 4095);
	blue0 = clampi(blue0, 0, 4095);
	red1 = clampi(red1, 0, 4095);
	green1 = clampi(green1, 0, 4095);
	blue1 = clampi(blue1, 0, 4095);
	int t0, t1;
	switch (majcomp) {
	case 1: t0 = red0; t1 = red1; red0 = green0; red1 = green1; green0 = t0; green1 = t1; break;
	case 2: t0 = red0; t1 = red1; red0 = blue0; red1 = blue1; blue0 = t0; blue1 = t1; break;
	default: break;
	}
	o0 = mki4(red0 << 4, green0 << 4, blue0 << 4, 0x7800);
	o1 = mki4(red1 << 4, green1 << 4, blue1 << 4, 0x7800);
}

ASTC_FN void hdr_luminance_small_range_unpack(const uint8_t in[2], i4& o0, i4& o1) {   // :735-770
	int v0 = in[0], v1 = in[1];
	int y0, y1;
	if (v0 & 0x80) {
		y0 = ((v1 & 0xE0) << 4) | ((v0 & 0x7F) << 2);
		y1 = (v1 & 0x1F) << 2;
	} else {
		y0 = ((v1 & 0xF0) << 4) | ((v0 & 0x7F) << 1);
		y1 = (v1 & 0xF) << 1;
	}
	y1 += y0;
	if (y1 > 0xFFF) {
		y1 = 0xFFF;
	}
	o0 = mki4(y0 << 4, y0 << 4, y0 << 4, 0x7800);
	o1 = mki4(y1 << 4, y1 << 4, y1 << 4, 0x7800);
}

ASTC_FN void hdr_luminance_large_range_unpack(const uint8_t in[2], i4& o0, i4& o1) {   // :773-798
	int v0 = in[0], v1 = in[1];
	int y0, y1;
	if (v1 >= v0) {
		y0 = v0 << 4;
		y1 = v1 << 4;
	} else {
		y0 = (v1 << 4) + 8;
		y1 = (v0 << 4) - 8;
	}
	o0 = mki4(y0 << 4, y0 << 4, y0 << 4, 0x7800);
	o1 = mki4(y1 << 4, y1 << 4, y1 << 4, 0x7800);
}

ASTC_NOINLINE void hdr_alpha_unpack(const uint8_t in[2], int& o0, int& o1) {   // :801-838
	int v6 = in[0], v7 = in[1];
	int selector = ((v6 >> 7) & 1) | ((v7 >> 6) & 2);
	v6 &= 0x7F;
	v7 &= 0x7F;
	if (selector == 3) {
		o0 = v6 << 5;
		o1 = v7 << 5;
	} else {
		v6 |= (v7 << (selector + 1)) & 0x780;
		v7 &= (0x3f >> selector);
		v7 ^= 32 >> selector;
		v7 -= 32 >> selector;
		v6 = v6 << (4 - selector);
		v7 = safe_signed_lsh(v7, 4 - selector);
		v7 = clampi(v6 + v7, 0, 0xFFF);
		o0 = v6;
		o1 = v7;
	}
	o0 <<= 4;
	o1 <<= 4;
}
