// Host-side table construction for the B200 ASTC compressor. See astc_host_tables.h for the reference citations.
#include "astc_host_tables.h"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>

namespace astc_host {

#include "astc_percentile_data.inc"

// ---------------------------------------------------------------------------------------------
// BISE: quant level -> (bits, trits, quints), sequence bit count
// (astcenc_integer_sequence.cpp:301-357 btq_counts / ise_sizes, :419 get_ise_sequence_bitcount)
// ---------------------------------------------------------------------------------------------
static const uint8_t LEVEL_COUNT[21] = {2, 3, 4, 5, 6, 8, 10, 12, 16, 20, 24, 32, 40, 48, 64, 80, 96, 128, 160, 192, 255};

unsigned int get_quant_level(int q) {
	return q == QUANT_256 ? 256u : LEVEL_COUNT[q];
}

void ise_btq(int q, unsigned int& bits, unsigned int& trits, unsigned int& quints) {
	// levels = 2^bits * (3 if trits) * (5 if quints)
	unsigned int levels = get_quant_level(q);
	trits = 0;
	quints = 0;
	if ((levels % 3) == 0) {
		trits = 1;
		levels /= 3;
	} else if ((levels % 5) == 0) {
		quints = 1;
		levels /= 5;
	}
	bits = 0;
	while ((1u << bits) < levels) {
		bits++;
	}
}

unsigned int ise_sequence_bitcount(unsigned int count, int q) {
	if (q < 0 || q > QUANT_256) {
		return 1024;
	}
	unsigned int bits, trits, quints;
	ise_btq(q, bits, trits, quints);
	// trits: 8 bits per 5 values, quints: 7 bits per 3 values, rounded up over the whole sequence
	if (trits) {
		return ((8 + 5 * bits) * count + 4) / 5;
	}
	if (quints) {
		return ((7 + 3 * bits) * count + 2) / 3;
	}
	return bits * count;
}

// ---------------------------------------------------------------------------------------------
// Trit / quint block decode per the ASTC format, and the encoder-side inverse
// (astcenc_integer_sequence.cpp:28-298)
// ---------------------------------------------------------------------------------------------
static void decode_trit_block(unsigned int T, uint8_t t[5]) {
	auto bit = [&](unsigned int v, int b) { return (v >> b) & 1u; };
	unsigned int C;
	unsigned int t4, t3, t2, t1, t0;
	if (((T >> 2) & 7) == 7) {
		C = (((T >> 5) & 7) << 2) | (T & 3);
		t4 = 2;
		t3 = 2;
	} else {
		C = T & 0x1F;
		if (((T >> 5) & 3) == 3) {
			t4 = 2;
			t3 = bit(T, 7);
		} else {
			t4 = bit(T, 7);
			t3 = (T >> 5) & 3;
		}
	}
	if ((C & 3) == 3) {
		t2 = 2;
		t1 = bit(C, 4);
		t0 = (bit(C, 3) << 1) | (bit(C, 2) & ~bit(C, 3) & 1);
	} else if (((C >> 2) & 3) == 3) {
		t2 = 2;
		t1 = 2;
		t0 = C & 3;
	} else {
		t2 = bit(C, 4);
		t1 = (C >> 2) & 3;
		t0 = (bit(C, 1) << 1) | (bit(C, 0) & ~bit(C, 1) & 1);
	}
	t[0] = (uint8_t)t0; t[1] = (uint8_t)t1; t[2] = (uint8_t)t2; t[3] = (uint8_t)t3; t[4] = (uint8_t)t4;
}

static void decode_quint_block(unsigned int Q, uint8_t q[3]) {
	auto bit = [&](unsigned int v, int b) { return (v >> b) & 1u; };
	unsigned int q2, q1, q0;
	if (((Q >> 1) & 3) == 3 && ((Q >> 5) & 3) == 0) {
		q2 = (bit(Q, 0) << 2) | ((bit(Q, 4) & ~bit(Q, 0) & 1) << 1) | (bit(Q, 3) & ~bit(Q, 0) & 1);
		q1 = 4;
		q0 = 4;
	} else {
		unsigned int C;
		if (((Q >> 1) & 3) == 3) {
			q2 = 4;
			C = (((Q >> 3) & 3) << 3) | ((~(Q >> 5) & 3) << 1) | bit(Q, 0);
		} else {
			q2 = (Q >> 5) & 3;
			C = Q & 0x1F;
		}
		if ((C & 7) == 5) {
			q1 = 4;
			q0 = (C >> 3) & 3;
		} else {
			q1 = (C >> 3) & 3;
			q0 = C & 7;
		}
	}
	q[0] = (uint8_t)q0; q[1] = (uint8_t)q1; q[2] = (uint8_t)q2;
}

static void build_ise_tables(ConstTables& ct) {
	// Several packed values decode to the same trit/quint tuple. The encoder table keeps the last
	// (highest) packed value, as the reference tables do (the reference keeps the highest packed value of each tuple).
	bool seen_t[3][3][3][3][3];
	memset(seen_t, 0, sizeof(seen_t));
	for (unsigned int T = 0; T < 256; T++) {
		uint8_t t[5];
		decode_trit_block(T, t);
		for (int k = 0; k < 5; k++) {
			ct.trits_of_integer[T][k] = t[k];
		}
	}
	for (int T = 255; T >= 0; T--) {
		const uint8_t* t = ct.trits_of_integer[T];
		bool& s = seen_t[t[4]][t[3]][t[2]][t[1]][t[0]];
		if (!s) {
			s = true;
			ct.integer_of_trits[t[4]][t[3]][t[2]][t[1]][t[0]] = (uint8_t)T;
		}
	}
	bool seen_q[5][5][5];
	memset(seen_q, 0, sizeof(seen_q));
	for (unsigned int Q = 0; Q < 128; Q++) {
		uint8_t q[3];
		decode_quint_block(Q, q);
		for (int k = 0; k < 3; k++) {
			ct.quints_of_integer[Q][k] = q[k];
		}
	}
	for (int Q = 127; Q >= 0; Q--) {
		const uint8_t* q = ct.quints_of_integer[Q];
		bool& s = seen_q[q[2]][q[1]][q[0]];
		if (!s) {
			s = true;
			ct.integer_of_quints[q[2]][q[1]][q[0]] = (uint8_t)Q;
		}
	}
}

// ---------------------------------------------------------------------------------------------
// Colour endpoint quantisation tables (astcenc_quantization.cpp:27, :338, :780)
// unquant rule = ASTC format "colour endpoint unquantisation"
// ---------------------------------------------------------------------------------------------
static unsigned int color_unquant_value(int q, unsigned int p) {
	unsigned int bits, trits, quints;
	ise_btq(q, bits, trits, quints);
	if (!trits && !quints) {
		// bit replication to 8 bits
		unsigned int v = p << (8 - bits);
		int remaining = 8 - (int)bits;
		while (remaining > 0) {
			int shift = remaining - (int)bits;
			remaining -= (int)bits;
			v |= shift > 0 ? (p << shift) : (p >> -shift);
		}
		return v & 0xFF;
	}
	unsigned int D = p >> bits;
	unsigned int m = p & ((1u << bits) - 1);
	unsigned int A = (m & 1) ? 0x1FF : 0;
	unsigned int b = (m >> 1) & 1, c = (m >> 2) & 1, d = (m >> 3) & 1, e = (m >> 4) & 1, f = (m >> 5) & 1;
	unsigned int B = 0, C = 0;
	if (trits) {
		switch (bits) {
		case 1: C = 204; B = 0; break;
		case 2: C = 93; B = b * 0x116; break;                                  // b000b0bb0
		case 3: C = 44; B = c * 0x10A + b * 0x085; break;                      // cb000cbcb
		case 4: C = 22; B = d * 0x104 + c * 0x082 + b * 0x041; break;          // dcb000dcb
		case 5: C = 11; B = e * 0x102 + d * 0x081 + c * 0x040 + b * 0x020; break;   // edcb000ed
		case 6: C = 5; B = f * 0x101 + e * 0x080 + d * 0x040 + c * 0x020 + b * 0x010; break;  // fedcb000f
		}
	} else {
		switch (bits) {
		case 1: C = 113; B = 0; break;
		case 2: C = 54; B = b * 0x10C; break;                                  // b0000bb00
		case 3: C = 26; B = c * 0x105 + b * 0x082; break;                      // cb0000cbc
		case 4: C = 13; B = d * 0x102 + c * 0x081 + b * 0x040; break;          // dcb0000dc
		case 5: C = 6; B = e * 0x101 + d * 0x080 + c * 0x040 + b * 0x020; break;    // edcb0000e
		}
	}
	unsigned int T = D * C + B;
	T ^= A;
	T = (A & 0x80) | (T >> 2);
	return T & 0xFF;
}

static void build_color_tables(ConstTables& ct) {
	memset(ct.color_uquant_to_scrambled_pquant, 0, sizeof(ct.color_uquant_to_scrambled_pquant));
	memset(ct.color_scrambled_pquant_to_uquant, 0, sizeof(ct.color_scrambled_pquant_to_uquant));
	for (int qi = 0; qi < 17; qi++) {
		int q = QUANT_6 + qi;
		unsigned int levels = get_quant_level(q);
		bool valid[256];
		memset(valid, 0, sizeof(valid));
		for (unsigned int p = 0; p < levels; p++) {
			unsigned int u = color_unquant_value(q, p);
			ct.color_scrambled_pquant_to_uquant[qi][p] = (uint8_t)u;
			valid[u] = true;
		}
		// value -> (nearest level rounding ties down, nearest level rounding ties up)
		for (unsigned int i = 0; i < 256; i++) {
			unsigned int min_dist = 256, lo = 256, hi = 0;
			for (unsigned int v = 0; v < 256; v++) {
				if (!valid[v]) {
					continue;
				}
				unsigned int dist = i > v ? i - v : v - i;
				if (dist < min_dist) {
					min_dist = dist;
					lo = v;
					hi = v;
				} else if (dist == min_dist) {
					lo = v < lo ? v : lo;
					hi = v > hi ? v : hi;
				}
			}
			ct.color_unquant_to_uquant[qi][2 * i] = (uint8_t)lo;
			ct.color_unquant_to_uquant[qi][2 * i + 1] = (uint8_t)hi;
		}
		// unquantised value -> packed (ISE) value; non-level inputs map via their nearest (ties up) level
		for (unsigned int i = 0; i < 256; i++) {
			unsigned int u = ct.color_unquant_to_uquant[qi][2 * i + 1];
			for (unsigned int p = 0; p < levels; p++) {
				if (ct.color_scrambled_pquant_to_uquant[qi][p] == u) {
					ct.color_uquant_to_scrambled_pquant[qi][i] = (uint8_t)p;
					break;
				}
			}
		}
	}
	// quant_mode_table[pairs][bits] = highest level whose 2*pairs-integer sequence fits (astcenc_quantization.cpp:802)
	for (int i = 0; i < 10; i++) {
		for (int j = 0; j < 128; j++) {
			int best = -1;
			if (i > 0) {
				for (int q = QUANT_2; q <= QUANT_256; q++) {
					if (ise_sequence_bitcount(2 * i, q) <= (unsigned int)j) {
						best = q;
					}
				}
			}
			ct.quant_mode_table[i][j] = (int8_t)best;
		}
	}
}

// ---------------------------------------------------------------------------------------------
// Weight quantisation transfer tables (astcenc_weight_quant_xfer_tables.cpp:26)
// unquant rule = ASTC format "weight unquantisation"
// ---------------------------------------------------------------------------------------------
static unsigned int weight_unquant_value(int q, unsigned int p) {
	unsigned int bits, trits, quints;
	ise_btq(q, bits, trits, quints);
	unsigned int T;
	if (!trits && !quints) {
		unsigned int v = p << (6 - bits);
		int remaining = 6 - (int)bits;
		while (remaining > 0) {
			int shift = remaining - (int)bits;
			remaining -= (int)bits;
			v |= shift > 0 ? (p << shift) : (p >> -shift);
		}
		T = v & 0x3F;
	} else if (bits == 0) {
		static const uint8_t t3[3] = {0, 32, 63};
		static const uint8_t q5[5] = {0, 16, 32, 47, 63};
		T = trits ? t3[p] : q5[p];
	} else {
		unsigned int D = p >> bits;
		unsigned int m = p & ((1u << bits) - 1);
		unsigned int A = (m & 1) ? 0x7F : 0;
		unsigned int b = (m >> 1) & 1, c = (m >> 2) & 1;
		unsigned int B = 0, C = 0;
		if (trits) {
			switch (bits) {
			case 1: C = 50; B = 0; break;
			case 2: C = 23; B = b * 0x45; break;              // b000b0b
			case 3: C = 11; B = c * 0x42 + b * 0x21; break;   // cb000cb
			}
		} else {
			switch (bits) {
			case 1: C = 28; B = 0; break;
			case 2: C = 13; B = b * 0x42; break;              // b0000b0
			}
		}
		T = D * C + B;
		T ^= A;
		T = (A & 0x20) | (T >> 2);
	}
	if (T > 32) {
		T += 1;
	}
	return T;
}

static void build_weight_tables(ConstTables& ct) {
	memset(ct.weight_quant, 0, sizeof(ct.weight_quant));
	for (int q = 0; q < 12; q++) {
		WeightQuantTable& wt = ct.weight_quant[q];
		unsigned int levels = get_quant_level(q);
		unsigned int unq[32];
		for (unsigned int p = 0; p < levels; p++) {
			unq[p] = weight_unquant_value(q, p);
			wt.unscramble_and_unquant_map[p] = (uint8_t)unq[p];
		}
		// sorted order
		unsigned int n = 0;
		for (unsigned int v = 0; v <= 64; v++) {
			for (unsigned int p = 0; p < levels; p++) {
				if (unq[p] == v) {
					wt.quant_to_unquant[n] = (uint8_t)v;
					wt.scramble_map[n] = (uint8_t)p;
					n++;
				}
			}
		}
		for (unsigned int i = 0; i < levels; i++) {
			unsigned int v = wt.quant_to_unquant[i];
			unsigned int prev = wt.quant_to_unquant[i == 0 ? 0 : i - 1];
			unsigned int next = wt.quant_to_unquant[i == levels - 1 ? i : i + 1];
			wt.prev_next_values[v] = (uint16_t)((next << 8) | prev);
		}
	}
}

static void build_angular_tables(ConstTables& ct) {
	// astcenc_weight_align.cpp:72-84 - evaluated by host libm in fp32
	const float PI = 3.14159265358979323846f;
	for (unsigned int i = 0; i < 32; i++) {
		float angle_step = static_cast<float>(i + 1);
		for (unsigned int j = 0; j < 64; j++) {
			ct.sin_table[j][i] = static_cast<float>(sinf((2.0f * PI / (64 - 1.0f)) * angle_step * static_cast<float>(j)));
			ct.cos_table[j][i] = static_cast<float>(cosf((2.0f * PI / (64 - 1.0f)) * angle_step * static_cast<float>(j)));
		}
	}
}

const ConstTables& const_tables() {
	static ConstTables* ct = nullptr;
	if (!ct) {
		ConstTables* t = new ConstTables;
		build_ise_tables(*t);
		build_color_tables(*t);
		build_weight_tables(*t);
		build_angular_tables(*t);
		ct = t;
	}
	return *ct;
}

// ---------------------------------------------------------------------------------------------
// Block modes (astcenc_block_sizes.cpp:36-98)
// ---------------------------------------------------------------------------------------------
static bool decode_block_mode_2d(unsigned int mode, unsigned int& wx, unsigned int& wy, bool& dual,
                                 unsigned int& quant_mode, unsigned int& weight_bits) {
	unsigned int base_quant = (mode >> 4) & 1;
	unsigned int H = (mode >> 9) & 1;
	unsigned int D = (mode >> 10) & 1;
	unsigned int A = (mode >> 5) & 3;
	wx = 0;
	wy = 0;
	if ((mode & 3) != 0) {
		base_quant |= (mode & 3) << 1;
		unsigned int B = (mode >> 7) & 3;
		switch ((mode >> 2) & 3) {
		case 0: wx = B + 4; wy = A + 2; break;
		case 1: wx = B + 8; wy = A + 2; break;
		case 2: wx = A + 2; wy = B + 8; break;
		case 3:
			B &= 1;
			if (mode & 0x100) {
				wx = B + 2;
				wy = A + 2;
			} else {
				wx = A + 2;
				wy = B + 6;
			}
			break;
		}
	} else {
		base_quant |= ((mode >> 2) & 3) << 1;
		if (((mode >> 2) & 3) == 0) {
			return false;
		}
		unsigned int B = (mode >> 9) & 3;
		switch ((mode >> 7) & 3) {
		case 0: wx = 12; wy = A + 2; break;
		case 1: wx = A + 2; wy = 12; break;
		case 2: wx = A + 6; wy = B + 6; D = 0; H = 0; break;
		case 3:
			switch ((mode >> 5) & 3) {
			case 0: wx = 6; wy = 10; break;
			case 1: wx = 10; wy = 6; break;
			default: return false;
			}
			break;
		}
	}
	unsigned int weight_count = wx * wy * (D + 1);
	quant_mode = (base_quant - 2) + 6 * H;
	dual = D != 0;
	weight_bits = ise_sequence_bitcount(weight_count, (int)quant_mode);
	return weight_count <= 64 && weight_bits >= 24 && weight_bits <= 96;
}

// ---------------------------------------------------------------------------------------------
// Decimation tables (astcenc_block_sizes.cpp:252-486)
// ---------------------------------------------------------------------------------------------
static void init_decimation_info_2d(unsigned int tx, unsigned int ty, unsigned int wx, unsigned int wy, DecimationInfo& di) {
	unsigned int texels = tx * ty;
	unsigned int weights = wx * wy;
	memset(&di, 0, sizeof(di));

	static uint8_t weight_count_of_texel[MAX_TEXELS];
	static uint8_t grid_weights_of_texel[MAX_TEXELS][4];
	static uint8_t weights_of_texel[MAX_TEXELS][4];
	static uint8_t texel_count_of_weight[MAX_WEIGHTS];
	static uint8_t texels_of_weight[MAX_WEIGHTS][MAX_TEXELS];
	static uint8_t texel_weights_of_weight[MAX_WEIGHTS][MAX_TEXELS];
	memset(weight_count_of_texel, 0, sizeof(weight_count_of_texel));
	memset(texel_count_of_weight, 0, sizeof(texel_count_of_weight));

	for (unsigned int y = 0; y < ty; y++) {
		for (unsigned int x = 0; x < tx; x++) {
			unsigned int texel = y * tx + x;
			unsigned int x_weight = (((1024 + tx / 2) / (tx - 1)) * x * (wx - 1) + 32) >> 6;
			unsigned int y_weight = (((1024 + ty / 2) / (ty - 1)) * y * (wy - 1) + 32) >> 6;
			unsigned int xf = x_weight & 0xF, yf = y_weight & 0xF;
			unsigned int xi = x_weight >> 4, yi = y_weight >> 4;
			unsigned int qweight[4];
			qweight[0] = xi + yi * wx;
			qweight[1] = qweight[0] + 1;
			qweight[2] = qweight[0] + wx;
			qweight[3] = qweight[2] + 1;
			unsigned int prod = xf * yf;
			unsigned int weight[4];
			weight[3] = (prod + 8) >> 4;
			weight[1] = xf - weight[3];
			weight[2] = yf - weight[3];
			weight[0] = 16 - xf - yf + weight[3];
			for (unsigned int i = 0; i < 4; i++) {
				if (weight[i] != 0) {
					unsigned int c = weight_count_of_texel[texel];
					grid_weights_of_texel[texel][c] = (uint8_t)qweight[i];
					weights_of_texel[texel][c] = (uint8_t)weight[i];
					weight_count_of_texel[texel]++;
					unsigned int tc = texel_count_of_weight[qweight[i]];
					texels_of_weight[qweight[i]][tc] = (uint8_t)texel;
					texel_weights_of_weight[qweight[i]][tc] = (uint8_t)weight[i];
					texel_count_of_weight[qweight[i]]++;
				}
			}
		}
	}

	uint8_t max_texel_weight_count = 0;
	for (unsigned int i = 0; i < texels; i++) {
		di.texel_weight_count[i] = weight_count_of_texel[i];
		if (di.texel_weight_count[i] > max_texel_weight_count) {
			max_texel_weight_count = di.texel_weight_count[i];
		}
		for (unsigned int j = 0; j < weight_count_of_texel[i]; j++) {
			di.texel_weight_contribs_int[j][i] = weights_of_texel[i][j];
			di.texel_weight_contribs_float[j][i] = static_cast<float>(weights_of_texel[i][j]) * (1.0f / 16.0f);
			di.texel_weights[j][i] = grid_weights_of_texel[i][j];
		}
		// unused taps stay {weight 0, contribution 0}
	}
	di.max_texel_weight_count = max_texel_weight_count;

	unsigned int off = 0;
	for (unsigned int i = 0; i < weights; i++) {
		unsigned int cnt = texel_count_of_weight[i];
		di.weight_texel_count[i] = (uint8_t)cnt;
		di.weight_texel_offset[i] = (uint16_t)off;
		for (unsigned int j = 0; j < cnt; j++) {
			uint8_t texel = texels_of_weight[i][j];
			di.weight_texels[off + j] = texel;
			di.weight_texel_contribs[off + j] = static_cast<float>(texel_weights_of_weight[i][j]);
			di.texel_contrib_for_weight[off + j] = 0.0f;
			for (unsigned int k = 0; k < 4; k++) {
				uint8_t dttw = di.texel_weights[k][texel];
				float dttwf = di.texel_weight_contribs_float[k][texel];
				if (dttw == i && dttwf != 0.0f) {
					di.texel_contrib_for_weight[off + j] = dttwf;
					break;
				}
			}
		}
		off += cnt;
	}
	di.weight_texel_offset[weights] = (uint16_t)off;

	di.texel_count = (uint8_t)texels;
	di.weight_count = (uint8_t)weights;
	di.weight_x = (uint8_t)wx;
	di.weight_y = (uint8_t)wy;
	di.weight_z = 1;
}

// ---------------------------------------------------------------------------------------------
// 3D block modes and decimation tables (astcenc_block_sizes.cpp:152-250, :450-700)
// ---------------------------------------------------------------------------------------------
static bool decode_block_mode_3d(unsigned int mode, unsigned int& wx, unsigned int& wy, unsigned int& wz, bool& dual,
                                 unsigned int& quant_mode, unsigned int& weight_bits) {
	unsigned int base_quant = (mode >> 4) & 1;
	unsigned int H = (mode >> 9) & 1;
	unsigned int D = (mode >> 10) & 1;
	unsigned int A = (mode >> 5) & 3;
	wx = wy = wz = 0;
	if ((mode & 3) != 0) {
		base_quant |= (mode & 3) << 1;
		wx = A + 2;
		wy = ((mode >> 7) & 3) + 2;
		wz = ((mode >> 2) & 3) + 2;
	} else {
		unsigned int sel = (mode >> 2) & 3;
		base_quant |= sel << 1;
		if (sel == 0) {
			return false;
		}
		unsigned int B = (mode >> 9) & 3;
		unsigned int layout = (mode >> 7) & 3;
		if (layout != 3) {
			D = 0;
			H = 0;
		}
		if (layout == 0) {
			wx = 6; wy = B + 2; wz = A + 2;
		} else if (layout == 1) {
			wx = A + 2; wy = 6; wz = B + 2;
		} else if (layout == 2) {
			wx = A + 2; wy = B + 2; wz = 6;
		} else {
			wx = wy = wz = 2;
			if (A == 0) wx = 6;
			else if (A == 1) wy = 6;
			else if (A == 2) wz = 6;
			else return false;
		}
	}
	unsigned int weight_count = wx * wy * wz * (D + 1);
	quant_mode = (base_quant - 2) + 6 * H;
	dual = D != 0;
	weight_bits = ise_sequence_bitcount(weight_count, (int)quant_mode);
	return weight_count <= 64 && weight_bits >= 24 && weight_bits <= 96;
}

// Fills the two directions of a DecimationInfo from per-texel tap lists (shared tail of the 2D and 3D builders).
struct TapLists {
	uint8_t weight_count_of_texel[MAX_TEXELS];
	uint8_t grid_weights_of_texel[MAX_TEXELS][4];
	uint8_t weights_of_texel[MAX_TEXELS][4];
	uint8_t texel_count_of_weight[MAX_WEIGHTS];
	uint8_t texels_of_weight[MAX_WEIGHTS][MAX_TEXELS];
	uint8_t texel_weights_of_weight[MAX_WEIGHTS][MAX_TEXELS];
};

static void add_tap(TapLists& tl, unsigned int texel, unsigned int grid_weight, unsigned int contribution) {
	if (contribution == 0) {
		return;
	}
	unsigned int c = tl.weight_count_of_texel[texel]++;
	tl.grid_weights_of_texel[texel][c] = (uint8_t)grid_weight;
	tl.weights_of_texel[texel][c] = (uint8_t)contribution;
	unsigned int tc = tl.texel_count_of_weight[grid_weight]++;
	tl.texels_of_weight[grid_weight][tc] = (uint8_t)texel;
	tl.texel_weights_of_weight[grid_weight][tc] = (uint8_t)contribution;
}

static void finish_decimation_info(const TapLists& tl, unsigned int texels, unsigned int weights, DecimationInfo& di) {
	uint8_t max_texel_weight_count = 0;
	for (unsigned int i = 0; i < texels; i++) {
		di.texel_weight_count[i] = tl.weight_count_of_texel[i];
		if (di.texel_weight_count[i] > max_texel_weight_count) {
			max_texel_weight_count = di.texel_weight_count[i];
		}
		for (unsigned int j = 0; j < tl.weight_count_of_texel[i]; j++) {
			di.texel_weight_contribs_int[j][i] = tl.weights_of_texel[i][j];
			di.texel_weight_contribs_float[j][i] = static_cast<float>(tl.weights_of_texel[i][j]) * (1.0f / 16.0f);
			di.texel_weights[j][i] = tl.grid_weights_of_texel[i][j];
		}
	}
	di.max_texel_weight_count = max_texel_weight_count;
	unsigned int off = 0;
	for (unsigned int i = 0; i < weights; i++) {
		unsigned int cnt = tl.texel_count_of_weight[i];
		di.weight_texel_count[i] = (uint8_t)cnt;
		di.weight_texel_offset[i] = (uint16_t)off;
		for (unsigned int j = 0; j < cnt; j++) {
			uint8_t texel = tl.texels_of_weight[i][j];
			di.weight_texels[off + j] = texel;
			di.weight_texel_contribs[off + j] = static_cast<float>(tl.texel_weights_of_weight[i][j]);
			di.texel_contrib_for_weight[off + j] = 0.0f;
			for (unsigned int k = 0; k < 4; k++) {
				uint8_t dttw = di.texel_weights[k][texel];
				float dttwf = di.texel_weight_contribs_float[k][texel];
				if (dttw == i && dttwf != 0.0f) {
					di.texel_contrib_for_weight[off + j] = dttwf;
					break;
				}
			}
		}
		off += cnt;
	}
	di.weight_texel_offset[weights] = (uint16_t)off;
	di.texel_count = (uint8_t)texels;
	di.weight_count = (uint8_t)weights;
}

// Simplex interpolation over a 3D weight grid: every texel takes the corner of its cell, the opposite corner and the two
// corners on the path between them that follows the fractions in descending order (:497-583).
static void init_decimation_info_3d(unsigned int tx, unsigned int ty, unsigned int tz, unsigned int wx, unsigned int wy, unsigned int wz,
                                    DecimationInfo& di) {
	memset(&di, 0, sizeof(di));
	static TapLists tl;
	memset(tl.weight_count_of_texel, 0, sizeof(tl.weight_count_of_texel));
	memset(tl.texel_count_of_weight, 0, sizeof(tl.texel_count_of_weight));
	const int stride[3] = {1, (int)wx, (int)(wx * wy)};
	for (unsigned int z = 0; z < tz; z++) {
		for (unsigned int y = 0; y < ty; y++) {
			for (unsigned int x = 0; x < tx; x++) {
				unsigned int texel = (z * ty + y) * tx + x;
				int g[3];
				g[0] = (int)((((1024 + tx / 2) / (tx - 1)) * x * (wx - 1) + 32) >> 6);
				g[1] = (int)((((1024 + ty / 2) / (ty - 1)) * y * (wy - 1) + 32) >> 6);
				g[2] = (int)((((1024 + tz / 2) / (tz - 1)) * z * (wz - 1) + 32) >> 6);
				int f[3] = {g[0] & 0xF, g[1] & 0xF, g[2] & 0xF};
				int base = ((g[2] >> 4) * (int)wy + (g[1] >> 4)) * (int)wx + (g[0] >> 4);
				// order the axes by descending fraction; ties resolve as the reference's comparison triple does
				// (fs > ft, ft > fp, fs > fp): cases 7 3 5 4 2 0, the impossible 1 and 6 fall to case 0
				bool s_gt_t = f[0] > f[1], t_gt_p = f[1] > f[2], s_gt_p = f[0] > f[2];
				int first, second, third;
				if (s_gt_t && t_gt_p && s_gt_p) { first = 0; second = 1; third = 2; }
				else if (!s_gt_t && t_gt_p && s_gt_p) { first = 1; second = 0; third = 2; }
				else if (s_gt_t && !t_gt_p && s_gt_p) { first = 0; second = 2; third = 1; }
				else if (s_gt_t && !t_gt_p && !s_gt_p) { first = 2; second = 0; third = 1; }
				else if (!s_gt_t && t_gt_p && !s_gt_p) { first = 1; second = 2; third = 0; }
				else { first = 2; second = 1; third = 0; }
				int q0 = base;
				int q1 = q0 + stride[first];
				int q2 = q1 + stride[second];
				int q3 = base + stride[0] + stride[1] + stride[2];
				add_tap(tl, texel, (unsigned int)q0, (unsigned int)(16 - f[first]));
				add_tap(tl, texel, (unsigned int)q1, (unsigned int)(f[first] - f[second]));
				add_tap(tl, texel, (unsigned int)q2, (unsigned int)(f[second] - f[third]));
				add_tap(tl, texel, (unsigned int)q3, (unsigned int)f[third]);
			}
		}
	}
	finish_decimation_info(tl, tx * ty * tz, wx * wy * wz, di);
	di.weight_x = (uint8_t)wx;
	di.weight_y = (uint8_t)wy;
	di.weight_z = (uint8_t)wz;
}

// ---------------------------------------------------------------------------------------------
// k-means texel subset (astcenc_block_sizes.cpp:717-754, astcenc_mathlib.cpp:32-48)
// ---------------------------------------------------------------------------------------------
static inline uint64_t rotl64(uint64_t v, int c) {
	return (v << c) | (v >> (64 - c));
}

static void assign_kmeans_texels(BlockSizeTables& bsd) {
	if (bsd.texel_count <= MAX_KMEANS_TEXELS) {
		for (uint8_t i = 0; i < bsd.texel_count; i++) {
			bsd.kmeans_texels[i] = i;
		}
		return;
	}
	uint64_t s[2] = {0xfaf9e171cea1ec6bULL, 0xf1b318cc06af5d71ULL};
	bool seen[MAX_TEXELS];
	for (unsigned int i = 0; i < bsd.texel_count; i++) {
		seen[i] = false;
	}
	unsigned int set = 0;
	while (set < (unsigned int)MAX_KMEANS_TEXELS) {
		uint64_t s0 = s[0], s1 = s[1];
		uint64_t res = s0 + s1;
		s1 ^= s0;
		s[0] = rotl64(s0, 24) ^ s1 ^ (s1 << 16);
		s[1] = rotl64(s1, 37);
		uint8_t texel = static_cast<uint8_t>(res);
		texel = texel % bsd.texel_count;
		if (!seen[texel]) {
			bsd.kmeans_texels[set++] = texel;
			seen[texel] = true;
		}
	}
}

// ---------------------------------------------------------------------------------------------
// Partition tables (astcenc_partition_tables.cpp)
// ---------------------------------------------------------------------------------------------
static uint32_t hash52(uint32_t inp) {
	inp ^= inp >> 15;
	inp *= 0xEEDE0891;
	inp ^= inp >> 5;
	inp += inp << 16;
	inp ^= inp >> 7;
	inp ^= inp >> 3;
	inp ^= inp << 6;
	inp ^= inp >> 17;
	return inp;
}

static uint8_t select_partition(int seed, int x, int y, int z, int partition_count, bool small_block) {
	if (small_block) {
		x <<= 1;
		y <<= 1;
		z <<= 1;
	}
	seed += (partition_count - 1) * 1024;
	uint32_t rnum = hash52(seed);
	uint8_t sd[12];
	sd[0] = rnum & 0xF;
	sd[1] = (rnum >> 4) & 0xF;
	sd[2] = (rnum >> 8) & 0xF;
	sd[3] = (rnum >> 12) & 0xF;
	sd[4] = (rnum >> 16) & 0xF;
	sd[5] = (rnum >> 20) & 0xF;
	sd[6] = (rnum >> 24) & 0xF;
	sd[7] = (rnum >> 28) & 0xF;
	sd[8] = (rnum >> 18) & 0xF;
	sd[9] = (rnum >> 22) & 0xF;
	sd[10] = (rnum >> 26) & 0xF;
	sd[11] = ((rnum >> 30) | (rnum << 2)) & 0xF;
	for (int i = 0; i < 12; i++) {
		sd[i] = (uint8_t)(sd[i] * sd[i]);
	}
	int sh1, sh2;
	if (seed & 1) {
		sh1 = (seed & 2 ? 4 : 5);
		sh2 = (partition_count == 3 ? 6 : 5);
	} else {
		sh1 = (partition_count == 3 ? 6 : 5);
		sh2 = (seed & 2 ? 4 : 5);
	}
	int sh3 = (seed & 0x10) ? sh1 : sh2;
	sd[0] >>= sh1; sd[1] >>= sh2; sd[2] >>= sh1; sd[3] >>= sh2;
	sd[4] >>= sh1; sd[5] >>= sh2; sd[6] >>= sh1; sd[7] >>= sh2;
	sd[8] >>= sh3; sd[9] >>= sh3; sd[10] >>= sh3; sd[11] >>= sh3;
	int a = sd[0] * x + sd[1] * y + sd[10] * z + (rnum >> 14);
	int b = sd[2] * x + sd[3] * y + sd[11] * z + (rnum >> 10);
	int c = sd[4] * x + sd[5] * y + sd[8] * z + (rnum >> 6);
	int d = sd[6] * x + sd[7] * y + sd[9] * z + (rnum >> 2);
	a &= 0x3F; b &= 0x3F; c &= 0x3F; d &= 0x3F;
	if (partition_count <= 3) d = 0;
	if (partition_count <= 2) c = 0;
	if (partition_count <= 1) b = 0;
	if (a >= b && a >= c && a >= d) return 0;
	if (b >= c && b >= d) return 1;
	if (c >= d) return 2;
	return 3;
}

static bool generate_one_partition_info_entry(BlockSizeTables& bsd, unsigned int partition_count, unsigned int partition_index,
                                              unsigned int remap_index, PartitionInfo& pi) {
	int texels_per_block = bsd.texel_count;
	bool small_block = texels_per_block < 32;
	int counts[4] = {0, 0, 0, 0};
	int texel_idx = 0;
	memset(&pi, 0, sizeof(pi));
	for (unsigned int z = 0; z < bsd.dim_z; z++) {
		for (unsigned int y = 0; y < bsd.dim_y; y++) {
			for (unsigned int x = 0; x < bsd.dim_x; x++) {
				uint8_t part = select_partition((int)partition_index, (int)x, (int)y, (int)z, (int)partition_count, small_block);
				pi.texels_of_partition[part][counts[part]++] = (uint8_t)texel_idx;
				pi.partition_of_texel[texel_idx] = part;
				texel_idx++;
			}
		}
	}
	if (counts[0] == 0) pi.partition_count = 0;
	else if (counts[1] == 0) pi.partition_count = 1;
	else if (counts[2] == 0) pi.partition_count = 2;
	else if (counts[3] == 0) pi.partition_count = 3;
	else pi.partition_count = 4;
	pi.partition_index = (uint16_t)partition_index;
	for (int i = 0; i < 4; i++) {
		pi.partition_texel_count[i] = (uint8_t)counts[i];
	}
	bool valid = pi.partition_count == partition_count;
	if (partition_count >= 2) {
		uint64_t* bitmaps = bsd.coverage_bitmaps[partition_count] + (size_t)remap_index * partition_count;
		for (unsigned int i = 0; i < partition_count; i++) {
			bitmaps[i] = 0;
		}
		unsigned int n = bsd.texel_count < MAX_KMEANS_TEXELS ? bsd.texel_count : (unsigned int)MAX_KMEANS_TEXELS;
		for (unsigned int i = 0; i < n; i++) {
			unsigned int idx = bsd.kmeans_texels[i];
			bitmaps[pi.partition_of_texel[idx]] |= 1ULL << i;
		}
	}
	return valid;
}

static const int BIT_PATTERN_WORDS = (MAX_TEXELS * 2 + 63) / 64;

static void canonical_pattern(unsigned int texel_count, const uint8_t* partition_of_texel, uint64_t* pat) {
	for (int i = 0; i < BIT_PATTERN_WORDS; i++) {
		pat[i] = 0;
	}
	int mapped[4] = {-1, -1, -1, -1};
	int n = 0;
	for (unsigned int i = 0; i < texel_count; i++) {
		int index = partition_of_texel[i];
		if (mapped[index] < 0) {
			mapped[index] = n++;
		}
		uint64_t x = (uint64_t)mapped[index];
		pat[i >> 5] |= x << (2 * (i & 0x1F));
	}
}

static void build_partition_table(BlockSizeTables& bsd, bool can_omit, unsigned int cutoff, unsigned int pc, uint64_t* patterns) {
	PartitionInfo* ptab = bsd.partitionings[pc];
	unsigned int next_index = 0;
	bsd.partitioning_count_selected[pc - 1] = 0;
	bsd.partitioning_count_all[pc - 1] = 0;
	for (int i = 0; i < MAX_PARTITIONINGS; i++) {
		bsd.partitioning_packed_index[pc - 2][i] = 0xFFFF;
	}
	if (can_omit && pc > cutoff) {
		return;
	}
	unsigned int max_iter = can_omit ? 1 : 2;
	uint8_t build[MAX_PARTITIONINGS];
	memset(build, 0, sizeof(build));
	for (unsigned int x = 0; x < max_iter; x++) {
		for (unsigned int i = 0; i < (unsigned int)MAX_PARTITIONINGS; i++) {
			if (x == 1 && build[i]) {
				continue;
			}
			bool keep_useful = generate_one_partition_info_entry(bsd, pc, i, next_index, ptab[next_index]);
			if (x == 0 && !keep_useful) {
				continue;
			}
			uint64_t* mine = patterns + (size_t)next_index * BIT_PATTERN_WORDS;
			canonical_pattern(bsd.texel_count, ptab[next_index].partition_of_texel, mine);
			bool keep_canonical = true;
			for (unsigned int j = 0; j < next_index; j++) {
				if (memcmp(mine, patterns + (size_t)j * BIT_PATTERN_WORDS, sizeof(uint64_t) * BIT_PATTERN_WORDS) == 0) {
					keep_canonical = false;
					break;
				}
			}
			if (keep_useful && keep_canonical) {
				if (x == 0) {
					bsd.partitioning_packed_index[pc - 2][i] = (uint16_t)next_index;
					bsd.partitioning_count_selected[pc - 1]++;
					bsd.partitioning_count_all[pc - 1]++;
					build[i] = 1;
					next_index++;
				}
			} else if (x == 1) {
				bsd.partitioning_packed_index[pc - 2][i] = (uint16_t)next_index;
				bsd.partitioning_count_all[pc - 1]++;
				next_index++;
			}
		}
	}
}

bool is_legal_3d_block_size(unsigned int x, unsigned int y, unsigned int z) {
	// the ten footprints of the format: each axis 3..6, sizes descending by at most one step from x to z
	static const uint8_t legal[10][3] = {{3, 3, 3}, {4, 3, 3}, {4, 4, 3}, {4, 4, 4}, {5, 4, 4}, {5, 5, 4}, {5, 5, 5}, {6, 5, 5}, {6, 6, 5}, {6, 6, 6}};
	for (int i = 0; i < 10; i++) {
		if (legal[i][0] == x && legal[i][1] == y && legal[i][2] == z) {
			return true;
		}
	}
	return false;
}

bool is_legal_2d_block_size(unsigned int x, unsigned int y) {
	static const uint8_t legal[14][2] = {{4, 4}, {5, 4}, {5, 5}, {6, 5}, {6, 6}, {8, 5}, {8, 6}, {8, 8},
	                                     {10, 5}, {10, 6}, {10, 8}, {10, 10}, {12, 10}, {12, 12}};
	for (int i = 0; i < 14; i++) {
		if (legal[i][0] == x && legal[i][1] == y) {
			return true;
		}
	}
	return false;
}

static void unpack_percentiles(unsigned int x, unsigned int y, float* table) {
	for (int i = 0; i < MAX_BLOCK_MODES; i++) {
		table[i] = 1.0f;
	}
	for (int r = 0; r < 14; r++) {
		if (PCT_ROWS[r][0] == x && PCT_ROWS[r][1] == y) {
			for (unsigned int k = 0; k < PCT_ROWS[r][2]; k++) {
				const uint32_t* e = PCT_ENTRIES[PCT_ROWS[r][3] + k];
				float f;
				memcpy(&f, &e[1], 4);
				table[e[0]] = f;
			}
		}
	}
}

// astcenc_block_sizes.cpp:822-1002 (four passes define the packed block-mode / decimation order)
// construct_block_size_descriptor_3d (astcenc_block_sizes.cpp:1025-1190): every grid that fits and every legal block mode is
// kept (no percentile selection for 3D); one-plane modes first, then two-plane modes.
static void build_modes_3d(BlockSizeTables& bsd, unsigned int tx, unsigned int ty, unsigned int tz) {
	int decimation_mode_index[7 * 64];
	for (auto& v : decimation_mode_index) {
		v = -1;
	}
	unsigned int dm_count = 0;
	for (unsigned int wx = 2; wx <= tx; wx++) {
		for (unsigned int wy = 2; wy <= ty; wy++) {
			for (unsigned int wz = 2; wz <= tz; wz++) {
				unsigned int weight_count = wx * wy * wz;
				if (weight_count > (unsigned int)MAX_WEIGHTS) {
					continue;
				}
				decimation_mode_index[wz * 64 + wy * 8 + wx] = (int)dm_count;
				init_decimation_info_3d(tx, ty, tz, wx, wy, wz, bsd.decimation_tables[dm_count]);
				int maxprec_1 = -1, maxprec_2 = -1;
				for (int q = 0; q < 12; q++) {
					unsigned int b1 = ise_sequence_bitcount(weight_count, q);
					if (b1 >= 24 && b1 <= 96) maxprec_1 = q;
					unsigned int b2 = ise_sequence_bitcount(2 * weight_count, q);
					if (b2 >= 24 && b2 <= 96) maxprec_2 = q;
				}
				if (2 * weight_count > (unsigned int)MAX_WEIGHTS) {
					maxprec_2 = -1;
				}
				DecimationMode& dm = bsd.decimation_modes[dm_count];
				dm.maxprec_1plane = (int8_t)maxprec_1;
				dm.maxprec_2planes = (int8_t)maxprec_2;
				dm.refprec_1plane = maxprec_1 == -1 ? 0 : 0xFFFF;
				dm.refprec_2planes = maxprec_2 == -1 ? 0 : 0xFFFF;
				dm_count++;
			}
		}
	}
	for (unsigned int i = dm_count; i < (unsigned int)MAX_DECIMATION_MODES; i++) {
		bsd.decimation_modes[i].maxprec_1plane = -1;
		bsd.decimation_modes[i].maxprec_2planes = -1;
		bsd.decimation_modes[i].refprec_1plane = 0;
		bsd.decimation_modes[i].refprec_2planes = 0;
	}
	bsd.decimation_mode_count_always = 0;
	bsd.decimation_mode_count_selected = dm_count;
	bsd.decimation_mode_count_all = dm_count;

	for (int i = 0; i < MAX_BLOCK_MODES; i++) {
		bsd.block_mode_packed_index[i] = 0xFFFF;
	}
	unsigned int packed = 0, counts[2] = {0, 0};
	for (unsigned int pass = 0; pass < 2; pass++) {
		for (unsigned int i = 0; i < (unsigned int)MAX_BLOCK_MODES; i++) {
			if (bsd.block_mode_packed_index[i] != 0xFFFF) {
				continue;
			}
			unsigned int wx, wy, wz, quant_mode, weight_bits;
			bool dual;
			if (!decode_block_mode_3d(i, wx, wy, wz, dual, quant_mode, weight_bits) || wx > tx || wy > ty || wz > tz) {
				continue;
			}
			if ((pass == 0) == dual) {
				continue;
			}
			if ((dual ? 109 : 111) - (int)weight_bits <= 0) {
				continue;
			}
			BlockMode& bm = bsd.block_modes[packed];
			bm.decimation_mode = (uint8_t)decimation_mode_index[wz * 64 + wy * 8 + wx];
			bm.quant_mode = (uint8_t)quant_mode;
			bm.weight_bits = (uint8_t)weight_bits;
			bm.is_dual_plane = dual ? 1 : 0;
			bm.mode_index = (uint16_t)i;
			bsd.block_mode_packed_index[i] = (uint16_t)packed;
			counts[pass]++;
			packed++;
		}
	}
	bsd.block_mode_count_1plane_always = 0;
	bsd.block_mode_count_1plane_selected = counts[0];
	bsd.block_mode_count_1plane_2plane_selected = counts[0] + counts[1];
	bsd.block_mode_count_all = counts[0] + counts[1];
}

static void build_partitions(BlockSizeTables& bsd, bool can_omit_modes, unsigned int partition_count_cutoff) {
	bsd.partitionings[1] = new PartitionInfo[1];
	for (int pc = 2; pc <= 4; pc++) {
		bsd.partitionings[pc] = new PartitionInfo[MAX_PARTITIONINGS];
		bsd.coverage_bitmaps[pc] = new uint64_t[(size_t)MAX_PARTITIONINGS * pc];
	}
	generate_one_partition_info_entry(bsd, 1, 0, 0, bsd.partitionings[1][0]);
	bsd.partitioning_count_selected[0] = 1;
	bsd.partitioning_count_all[0] = 1;
	uint64_t* patterns = new uint64_t[(size_t)MAX_PARTITIONINGS * BIT_PATTERN_WORDS];
	for (unsigned int pc = 2; pc <= 4; pc++) {
		build_partition_table(bsd, can_omit_modes, partition_count_cutoff, pc, patterns);
	}
	delete[] patterns;
}

BlockSizeTables* build_block_size_tables(unsigned int tx, unsigned int ty, unsigned int tz, bool can_omit_modes,
                                         unsigned int partition_count_cutoff, float mode_cutoff) {
	BlockSizeTables* bp = new BlockSizeTables;
	BlockSizeTables& bsd = *bp;
	memset(bp, 0, sizeof(*bp));
	bsd.dim_x = (uint8_t)tx;
	bsd.dim_y = (uint8_t)ty;
	bsd.dim_z = (uint8_t)tz;
	bsd.texel_count = (uint8_t)(tx * ty * tz);
	bsd.decimation_tables = new DecimationInfo[MAX_DECIMATION_MODES];
	if (tz > 1) {
		build_modes_3d(bsd, tx, ty, tz);
		assign_kmeans_texels(bsd);
		build_partitions(bsd, can_omit_modes, partition_count_cutoff);
		return bp;
	}

	int decimation_mode_index[12 * 16 + 12 + 16];
	for (auto& v : decimation_mode_index) {
		v = -1;
	}
	float percentiles[MAX_BLOCK_MODES];
	unpack_percentiles(tx, ty, percentiles);
	const float always_cutoff = 0.0f;

	unsigned int packed_bm_idx = 0, packed_dm_idx = 0;
	unsigned int bm_counts[4] = {0, 0, 0, 0}, dm_counts[4] = {0, 0, 0, 0};
	for (int i = 0; i < MAX_BLOCK_MODES; i++) {
		bsd.block_mode_packed_index[i] = 0xFFFF;
	}
	unsigned int limit = can_omit_modes ? 3 : 4;
	for (unsigned int j = 0; j < limit; j++) {
		for (unsigned int i = 0; i < (unsigned int)MAX_BLOCK_MODES; i++) {
			if (bsd.block_mode_packed_index[i] != 0xFFFF) {
				continue;
			}
			unsigned int wx, wy, quant_mode, weight_bits;
			bool dual;
			bool valid = decode_block_mode_2d(i, wx, wy, dual, quant_mode, weight_bits);
			if (!valid || wx > tx || wy > ty) {
				continue;
			}
			if ((j <= 1 && dual) || (j == 2 && !dual)) {
				continue;
			}
			if (dual) {
				if ((109 - (int)weight_bits) <= 0) continue;
			} else {
				if ((111 - (int)weight_bits) <= 0) continue;
			}
			bool hit = j == 0 ? percentiles[i] <= always_cutoff : percentiles[i] <= mode_cutoff;
			if (j != 3 && !hit) {
				continue;
			}
			int dmode = decimation_mode_index[wy * 16 + wx];
			if (dmode < 0) {
				// construct_dt_entry_2d (astcenc_block_sizes.cpp:756-815)
				unsigned int weight_count = wx * wy;
				bool try_2planes = (2 * weight_count) <= 64;
				init_decimation_info_2d(tx, ty, wx, wy, bsd.decimation_tables[packed_dm_idx]);
				int maxprec_1 = -1, maxprec_2 = -1;
				for (int q = 0; q < 12; q++) {
					unsigned int b1 = ise_sequence_bitcount(weight_count, q);
					if (b1 >= 24 && b1 <= 96) maxprec_1 = q;
					if (try_2planes) {
						unsigned int b2 = ise_sequence_bitcount(2 * weight_count, q);
						if (b2 >= 24 && b2 <= 96) maxprec_2 = q;
					}
				}
				DecimationMode& dm = bsd.decimation_modes[packed_dm_idx];
				dm.maxprec_1plane = (int8_t)maxprec_1;
				dm.maxprec_2planes = (int8_t)maxprec_2;
				dm.refprec_1plane = 0;
				dm.refprec_2planes = 0;
				decimation_mode_index[wy * 16 + wx] = (int)packed_dm_idx;
				dmode = (int)packed_dm_idx;
				dm_counts[j]++;
				packed_dm_idx++;
			}
			BlockMode& bm = bsd.block_modes[packed_bm_idx];
			bm.decimation_mode = (uint8_t)dmode;
			bm.quant_mode = (uint8_t)quant_mode;
			bm.is_dual_plane = dual ? 1 : 0;
			bm.weight_bits = (uint8_t)weight_bits;
			bm.mode_index = (uint16_t)i;
			DecimationMode& dm = bsd.decimation_modes[dmode];
			if (dual) {
				dm.refprec_2planes |= (uint16_t)(1u << quant_mode);
			} else {
				dm.refprec_1plane |= (uint16_t)(1u << quant_mode);
			}
			bsd.block_mode_packed_index[i] = (uint16_t)packed_bm_idx;
			packed_bm_idx++;
			bm_counts[j]++;
		}
	}
	bsd.block_mode_count_1plane_always = bm_counts[0];
	bsd.block_mode_count_1plane_selected = bm_counts[0] + bm_counts[1];
	bsd.block_mode_count_1plane_2plane_selected = bm_counts[0] + bm_counts[1] + bm_counts[2];
	bsd.block_mode_count_all = bm_counts[0] + bm_counts[1] + bm_counts[2] + bm_counts[3];
	bsd.decimation_mode_count_always = dm_counts[0];
	bsd.decimation_mode_count_selected = dm_counts[0] + dm_counts[1] + dm_counts[2];
	bsd.decimation_mode_count_all = dm_counts[0] + dm_counts[1] + dm_counts[2] + dm_counts[3];
	for (unsigned int i = bsd.decimation_mode_count_all; i < (unsigned int)MAX_DECIMATION_MODES; i++) {
		bsd.decimation_modes[i].maxprec_1plane = -1;
		bsd.decimation_modes[i].maxprec_2planes = -1;
	}
	assign_kmeans_texels(bsd);
	build_partitions(bsd, can_omit_modes, partition_count_cutoff);
	return bp;
}

void free_block_size_tables(BlockSizeTables* t) {
	if (!t) {
		return;
	}
	delete[] t->decimation_tables;
	delete[] t->partitionings[1];
	for (int pc = 2; pc <= 4; pc++) {
		delete[] t->partitionings[pc];
		delete[] t->coverage_bitmaps[pc];
	}
	delete t;
}

}  // namespace astc_host
