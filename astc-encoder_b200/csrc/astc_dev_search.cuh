// B200-native ASTC block compressor: warp-cooperative device code, part 2
// (per-mode weight quantisation + error, endpoint format choice, candidate refinement, error scoring,
//  weight realignment, partition search, physical packing, the compress_block driver).
#pragma once
#include "astc_dev_core.cuh"

// su (union scratch) sub-layouts. Each phase owns the whole union while it runs.
//   quantise+error pass : per-lane quantised weight rows                  UQ_ROW_STRIDE * 32
//   endpoint formats    : best_error/format tables, combined tables       see EF_* below
//   refinement          : undecimated weights, int weights, realign staging
//   partition search    : mismatch counts, ordering, histogram, k-means state, candidate errors
#define UQ_ROW_STRIDE 68    /* 17 words: conflict-free lane-private rows */

// =============================================================================================
// Per block mode: quantise the decimated ideal weights and measure the weight-set error
// (astcenc_compress_symbolic.cpp:434-485 / :803-868 + ideal_endpoints.cpp:688-842, :974-1080).
// Lanes over block modes; the texel error sum keeps the reference's 4-lane accumulator order per mode.
// =============================================================================================
ASTC_COOP void quantize_and_score_modes(WCtx& w, unsigned int start_mode, unsigned int end_mode, int nplanes, unsigned int partition_count,
                                        int max_weight_quant, float min_wt_cutoff1, float min_wt_cutoff2) {
	const DevBsd& bsd = *w.bsd;
	const int8_t free_bits_for_partition_count[4] = {115 - 4, 111 - 4 - 10, 108 - 4 - 10, 105 - 4 - 10};
	uint8_t* uqrow = w.su + w.lane * UQ_ROW_STRIDE;
	int T = w.T;
	for (unsigned int i = start_mode + (unsigned int)w.lane; i < end_mode; i += ASTC_WARP) {
		const DevBlockMode bm = bsd.block_modes[i];
		if (bm.quant_mode > max_weight_quant) {
			w.mode_err[i] = 1e38f;
			continue;
		}
		if (nplanes == 1) {
			int bitcount = free_bits_for_partition_count[partition_count - 1] - bm.weight_bits;
			if (bitcount <= 0) {
				w.mode_err[i] = 1e38f;
				continue;
			}
		}
		DecView di = dec_view(bsd, bm.decimation_mode);
		int W = di.W;
		float low1, high1, low2 = 0.0f, high2 = 1.0f;
		mode_low_high(w, bm, 0, min_wt_cutoff1, low1, high1);
		WeightQuantizer z1 = make_weight_quantizer(low1, high1, bm.quant_mode);
		WeightQuantizer z2 = z1;
		const float* ideal1 = w.dwi + di.dm->dwi_offset;
		for (int k = 0; k < W; k++) {
			uqrow[k] = (uint8_t)quantize_weight(z1, ideal1[k]);
		}
		if (nplanes == 2) {
			mode_low_high(w, bm, 1, min_wt_cutoff2, low2, high2);
			z2 = make_weight_quantizer(low2, high2, bm.quant_mode);
			const float* ideal2 = ideal1 + W;
			for (int k = 0; k < W; k++) {
				uqrow[32 + k] = (uint8_t)quantize_weight(z2, ideal2[k]);
			}
		}
		// compute_error_of_weight_set_1plane / _2planes
		float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
		const float* eiw1 = w.eiw[0];
		const float* eis1 = w.eis[0];
		const float* eiw2 = w.eiw[1];
		const float* eis2 = w.eis[1];
		for (int t = 0; t < T; t++) {
			float cur1, cur2 = 0.0f;
			if (di.max_twc > 2) {
				cur1 = (quantized_weight_value(z1, uqrow[di.tw[t]]) * contrib_f(di.tc[t]) +
				        quantized_weight_value(z1, uqrow[di.tw[T + t]]) * contrib_f(di.tc[T + t])) +
				       (quantized_weight_value(z1, uqrow[di.tw[2 * T + t]]) * contrib_f(di.tc[2 * T + t]) +
				        quantized_weight_value(z1, uqrow[di.tw[3 * T + t]]) * contrib_f(di.tc[3 * T + t]));
				if (nplanes == 2) {
					cur2 = (quantized_weight_value(z2, uqrow[32 + di.tw[t]]) * contrib_f(di.tc[t]) +
					        quantized_weight_value(z2, uqrow[32 + di.tw[T + t]]) * contrib_f(di.tc[T + t])) +
					       (quantized_weight_value(z2, uqrow[32 + di.tw[2 * T + t]]) * contrib_f(di.tc[2 * T + t]) +
					        quantized_weight_value(z2, uqrow[32 + di.tw[3 * T + t]]) * contrib_f(di.tc[3 * T + t]));
				}
			} else if (di.max_twc > 1) {
				cur1 = (quantized_weight_value(z1, uqrow[di.tw[t]]) * contrib_f(di.tc[t]) +
				        quantized_weight_value(z1, uqrow[di.tw[T + t]]) * contrib_f(di.tc[T + t]));
				if (nplanes == 2) {
					cur2 = (quantized_weight_value(z2, uqrow[32 + di.tw[t]]) * contrib_f(di.tc[t]) +
					        quantized_weight_value(z2, uqrow[32 + di.tw[T + t]]) * contrib_f(di.tc[T + t]));
				}
			} else {
				cur1 = quantized_weight_value(z1, uqrow[t]);
				if (nplanes == 2) {
					cur2 = quantized_weight_value(z2, uqrow[32 + t]);
				}
			}
			float diff = cur1 - eiw1[t];
			float error = diff * diff * eis1[t];
			if (nplanes == 2) {
				float diff2 = cur2 - eiw2[t];
				float error2 = diff2 * diff2 * eis2[t];
				error = error + error2;
			}
			acc[t & 3] = acc[t & 3] + error;
		}
		w.mode_err[i] = (acc[0] + acc[2]) + (acc[1] + acc[3]);
	}
	wsync();
}

// =============================================================================================
// Endpoint format choice (astcenc_pick_best_endpoint_format.cpp)
// =============================================================================================
struct EncodingChoiceErrors {
	float rgb_scale_error, rgb_luma_error, luminance_error, alpha_drop_error;
	bool can_offset_encode, can_blue_contract;
};

struct ProcessedLine {
	f4 amod;
	f4 bs;
};

ASTC_FN f4 dot3_splat(f4 a, f4 b) {
	float d = dot3_s(a, b);
	return mk4(d, d, d, 0.0f);
}

// layout of su during endpoint format selection
struct EfTables {
	float best_error[4][21][4];
	float combined_error[21][13];
	uint8_t format_of_choice[4][21][4];
	uint8_t combined_format[21][13][4];
};

// compute_encoding_choice_errors :222-312 with compute_error_squared_rgb_single_partition :72-219.
// One chain per (partition, accumulator a, texel index mod 4).
ASTC_COOP void compute_encoding_choice_errors(WCtx& w, const PartView& pi, int ep0slot, int ep1slot, EncodingChoiceErrors eci[4]) {
	int pc = (int)pi.partition_count;
	PartitionMetrics pms[4];
	compute_avgs_and_dirs_3_comp_rgb(w, pi, pms);
	ProcessedLine uncor[4], samec[4], rgbl[4], lum[4];
	for (int i = 0; i < pc; i++) {
		f4 uncor_a = pms[i].avg;
		f4 uncor_b = normalize_safe4(pms[i].dir, unit3());
		f4 samec_b = normalize_safe4(pms[i].avg, unit3());
		f4 luma_a = pms[i].avg;
		f4 luma_b = unit3();
		uncor[i].amod = uncor_a - uncor_b * dot3_splat(uncor_a, uncor_b);
		uncor[i].bs = uncor_b;
		samec[i].amod = splat4(0.0f);
		samec[i].bs = samec_b;
		rgbl[i].amod = luma_a - luma_b * dot3_splat(luma_a, luma_b);
		rgbl[i].bs = luma_b;
		lum[i].amod = splat4(0.0f);
		lum[i].bs = unit3();
	}
	f4 ews = w.bi.channel_weight;
	float default_a = default_alpha(w);
	int nchains = pc * 20;
	for (int id = w.lane; id < nchains; id += ASTC_WARP) {
		int p = id / 20;
		int r = id - p * 20;
		int a = r >> 2;
		int l = r & 3;
		const uint8_t* tix = pi.texels + pi.start[p];
		int n = pi.count[p];
		f4 amod, bs;
		if (a == 1) { amod = uncor[p].amod; bs = uncor[p].bs; }
		else if (a == 2) { amod = samec[p].amod; bs = samec[p].bs; }
		else if (a == 3) { amod = rgbl[p].amod; bs = rgbl[p].bs; }
		else { amod = lum[p].amod; bs = lum[p].bs; }
		float s = 0.0f;
		for (int j = l; j < n; j += 4) {
			int t = tix[j];
			float term;
			if (a == 0) {
				float alpha_diff = w.blk[3][t] - default_a;
				term = alpha_diff * alpha_diff;
			} else {
				float dr = w.blk[0][t], dg = w.blk[1][t], db = w.blk[2][t];
				float param = dr * bs.x + dg * bs.y + db * bs.z;
				float dist0, dist1, dist2;
				if (a == 1 || a == 3) {
					dist0 = (amod.x + param * bs.x) - dr;
					dist1 = (amod.y + param * bs.y) - dg;
					dist2 = (amod.z + param * bs.z) - db;
				} else {
					dist0 = (param * bs.x) - dr;
					dist1 = (param * bs.y) - dg;
					dist2 = (param * bs.z) - db;
				}
				term = dist0 * dist0 * ews.x + dist1 * dist1 * ews.y + dist2 * dist2 * ews.z;
			}
			s = s + term;
		}
		w.tmpf[id] = s;
	}
	wsync();
	for (int i = 0; i < pc; i++) {
		const float* a = w.tmpf + i * 20;
		float alpha_drop_error = ((a[0] + a[2]) + (a[1] + a[3])) * ews.w;
		float uncorr_rgb_error = (a[4] + a[6]) + (a[5] + a[7]);
		float samechroma_rgb_error = (a[8] + a[10]) + (a[9] + a[11]);
		float rgb_luma_error = (a[12] + a[14]) + (a[13] + a[15]);
		float luminance_rgb_error = (a[16] + a[18]) + (a[17] + a[19]);
		f4 d = w.ep[ep1slot + i] - w.ep[ep0slot + i];
		const float lim = 0.12f * 65535.0f;
		eci[i].can_offset_encode = (absf(d.x) < lim) && (absf(d.y) < lim) && (absf(d.z) < lim);
		eci[i].rgb_scale_error = (samechroma_rgb_error - uncorr_rgb_error) * 0.7f;
		eci[i].rgb_luma_error = (rgb_luma_error - uncorr_rgb_error) * 1.5f;
		eci[i].luminance_error = (luminance_rgb_error - uncorr_rgb_error) * 3.0f;
		eci[i].alpha_drop_error = alpha_drop_error * 3.0f;
		eci[i].can_blue_contract = !is_luminance(w);
	}
	wsync();
}

// compute_color_error_for_every_integer_count_and_quant_level :315-675. Lanes over quant levels.
ASTC_COOP void compute_color_error_tables(WCtx& w, const PartView& pi, int partition_index, const EncodingChoiceErrors& eci, int ep0slot, int ep1slot, EfTables* ef) {
	bool encode_hdr_rgb = w.bi.rgb_lns0 != 0;
	bool encode_hdr_alpha = w.bi.alpha_lns0 != 0;
	f4 error_weight = w.bi.channel_weight;
	int partition_size = pi.count[partition_index];
	const float den[17] = {5 * 5, 7 * 7, 9 * 9, 11 * 11, 15 * 15, 19 * 19, 23 * 23, 31 * 31, 39 * 39, 47 * 47,
	                       63 * 63, 79 * 79, 95 * 95, 127 * 127, 159 * 159, 191 * 191, 255 * 255};
	f4 ep0 = w.ep[ep0slot + partition_index];
	f4 ep1 = w.ep[ep1slot + partition_index];
	float ep1_min = hmin_s(mk4(ep1.x, ep1.y, ep1.z, ep1.x));
	ep1_min = maxf(ep1_min, 0.0f);
	float error_weight_rgbsum = hadd_rgb_s(error_weight);
	float range_upper_limit_rgb = encode_hdr_rgb ? 61440.0f : 65535.0f;
	float range_upper_limit_alpha = encode_hdr_alpha ? 61440.0f : 65535.0f;
	f4 offset = mk4(range_upper_limit_rgb, range_upper_limit_rgb, range_upper_limit_rgb, range_upper_limit_alpha);
	f4 zero = splat4(0.0f);
	f4 ep0_high = max4(ep0 - offset, zero);
	f4 ep1_high = max4(ep1 - offset, zero);
	f4 ep0_low = min4(ep0, zero);
	f4 ep1_low = min4(ep1, zero);
	f4 sum_range_error = (ep0_low * ep0_low) + (ep1_low * ep1_low) + (ep0_high * ep0_high) + (ep1_high * ep1_high);
	float rgb_range_error = dot3_s(sum_range_error, error_weight) * 0.5f * static_cast<float>(partition_size);
	float alpha_range_error = sum_range_error.w * error_weight.w * 0.5f * static_cast<float>(partition_size);
	float (*best_error)[4] = ef->best_error[partition_index];
	uint8_t (*format_of_choice)[4] = ef->format_of_choice[partition_index];

	if (encode_hdr_rgb) {
		float af, cf;
		if (ep1.x > ep1.y && ep1.x > ep1.z) {
			af = ep1.x;
			cf = ep1.x - ep0.x;
		} else if (ep1.y > ep1.z) {
			af = ep1.y;
			cf = ep1.y - ep0.y;
		} else {
			af = ep1.z;
			cf = ep1.z - ep0.z;
		}
		float bf = af - ep1_min;
		f4 prd = mk4(ep1.x - cf, ep1.y - cf, ep1.z - cf, 0.0f);
		f4 pdif = prd - mk4(ep0.x, ep0.y, ep0.z, 0.0f);
		float df = hmax_s(mk4(absf(pdif.x), absf(pdif.y), absf(pdif.z), absf(pdif.w)));
		int b = static_cast<int>(clampf(bf, 0.0f, 65536.0f));
		int c = static_cast<int>(clampf(cf, 0.0f, 65536.0f));
		int d = static_cast<int>(clampf(df, 0.0f, 65536.0f));
		int rgbo_mode = 5;
		if (b < 32768 && c < 16384) rgbo_mode = 4;
		if (b < 8192 && c < 16384) rgbo_mode = 3;
		if (b < 2048 && c < 16384) rgbo_mode = 2;
		if (b < 2048 && c < 1024) rgbo_mode = 1;
		if (b < 1024 && c < 4096) rgbo_mode = 0;
		int rgb_mode = 8;
		if (b < 16384 && c < 8192 && d < 8192) rgb_mode = 0;
		if (b < 32768 && c < 8192 && d < 4096) rgb_mode = 1;
		if (b < 4096 && c < 8192 && d < 4096) rgb_mode = 2;
		if (b < 8192 && c < 8192 && d < 2048) rgb_mode = 3;
		if (b < 8192 && c < 2048 && d < 512) rgb_mode = 4;
		if (b < 2048 && c < 8192 && d < 1024) rgb_mode = 5;
		if (b < 2048 && c < 2048 && d < 256) rgb_mode = 6;
		if (b < 1024 && c < 2048 && d < 512) rgb_mode = 7;
		const float rgbo_error_scales[6] = {4.0f, 4.0f, 16.0f, 64.0f, 256.0f, 1024.0f};
		const float rgb_error_scales[9] = {64.0f, 64.0f, 16.0f, 16.0f, 4.0f, 4.0f, 1.0f, 1.0f, 384.0f};
		float mode7mult = rgbo_error_scales[rgbo_mode] * 0.0015f;
		float mode11mult = rgb_error_scales[rgb_mode] * 0.010f;
		float lum_high = hadd_rgb_s(ep1) * (1.0f / 3.0f);
		float lum_low = hadd_rgb_s(ep0) * (1.0f / 3.0f);
		float lumdif = lum_high - lum_low;
		float mode23mult = lumdif < 960 ? 4.0f : lumdif < 3968 ? 16.0f : 128.0f;
		mode23mult *= 0.0005f;
		for (int i = w.lane; i <= QUANT_256; i += ASTC_WARP) {
			format_of_choice[i][3] = static_cast<uint8_t>(encode_hdr_alpha ? FMT_HDR_RGBA : FMT_HDR_RGB_LDR_ALPHA);
			format_of_choice[i][2] = FMT_HDR_RGB;
			format_of_choice[i][1] = FMT_HDR_RGB_SCALE;
			format_of_choice[i][0] = FMT_HDR_LUMINANCE_LARGE_RANGE;
			if (i < QUANT_16) {
				best_error[i][3] = ERROR_CALC_DEFAULT;
				best_error[i][2] = ERROR_CALC_DEFAULT;
				best_error[i][1] = ERROR_CALC_DEFAULT;
				best_error[i][0] = ERROR_CALC_DEFAULT;
				continue;
			}
			float base_quant_error = ((65536.0f * 65536.0f / 18.0f) / den[i - QUANT_6]) * static_cast<float>(partition_size);
			float rgb_quantization_error = error_weight_rgbsum * base_quant_error * 2.0f;
			float alpha_quantization_error = error_weight.w * base_quant_error * 2.0f;
			float rgba_quantization_error = rgb_quantization_error + alpha_quantization_error;
			best_error[i][3] = rgba_quantization_error + rgb_range_error + alpha_range_error;
			best_error[i][2] = (rgb_quantization_error * mode11mult) + rgb_range_error + eci.alpha_drop_error;
			best_error[i][1] = (rgb_quantization_error * mode7mult) + rgb_range_error + eci.alpha_drop_error + eci.rgb_luma_error;
			best_error[i][0] = (rgb_quantization_error * mode23mult) + rgb_range_error + eci.alpha_drop_error + eci.luminance_error;
		}
	} else {
		float base_quant_error_rgb = error_weight_rgbsum * static_cast<float>(partition_size);
		float base_quant_error_a = error_weight.w * static_cast<float>(partition_size);
		float base_quant_error_rgba = base_quant_error_rgb + base_quant_error_a;
		float error_scale_bc_rgba = eci.can_blue_contract ? 0.625f : 1.0f;
		float error_scale_bc_rgb = eci.can_blue_contract ? 0.5f : 1.0f;
		for (int i = w.lane; i <= QUANT_256; i += ASTC_WARP) {
			if (i < QUANT_6) {
				best_error[i][3] = ERROR_CALC_DEFAULT;
				best_error[i][2] = ERROR_CALC_DEFAULT;
				best_error[i][1] = ERROR_CALC_DEFAULT;
				best_error[i][0] = ERROR_CALC_DEFAULT;
				format_of_choice[i][3] = FMT_RGBA;
				format_of_choice[i][2] = FMT_RGB;
				format_of_choice[i][1] = FMT_RGB_SCALE;
				format_of_choice[i][0] = FMT_LUMINANCE;
				continue;
			}
			float error_scale_oe_rgba = eci.can_offset_encode ? 0.5f : 1.0f;
			float error_scale_oe_rgb = eci.can_offset_encode ? 0.25f : 1.0f;
			if (i >= QUANT_192) {
				error_scale_oe_rgba = 1.0f;
				error_scale_oe_rgb = 1.0f;
			}
			float base_quant_error = (65536.0f * 65536.0f / 18.0f) / den[i - QUANT_6];
			float quant_error_rgb = base_quant_error_rgb * base_quant_error;
			float quant_error_rgba = base_quant_error_rgba * base_quant_error;
			best_error[i][3] = quant_error_rgba * error_scale_bc_rgba * error_scale_oe_rgba + rgb_range_error + alpha_range_error;
			format_of_choice[i][3] = FMT_RGBA;
			float full_ldr_rgb_error = quant_error_rgb * error_scale_bc_rgb * error_scale_oe_rgb + rgb_range_error + eci.alpha_drop_error;
			float rgbs_alpha_error = quant_error_rgba + eci.rgb_scale_error + rgb_range_error + alpha_range_error;
			if (rgbs_alpha_error < full_ldr_rgb_error) {
				best_error[i][2] = rgbs_alpha_error;
				format_of_choice[i][2] = FMT_RGB_SCALE_ALPHA;
			} else {
				best_error[i][2] = full_ldr_rgb_error;
				format_of_choice[i][2] = FMT_RGB;
			}
			float ldr_rgbs_error = quant_error_rgb + rgb_range_error + eci.alpha_drop_error + eci.rgb_scale_error;
			float lum_alpha_error = quant_error_rgba + rgb_range_error + alpha_range_error + eci.luminance_error;
			if (ldr_rgbs_error < lum_alpha_error) {
				best_error[i][1] = ldr_rgbs_error;
				format_of_choice[i][1] = FMT_RGB_SCALE;
			} else {
				best_error[i][1] = lum_alpha_error;
				format_of_choice[i][1] = FMT_LUMINANCE_ALPHA;
			}
			best_error[i][0] = quant_error_rgb + rgb_range_error + eci.alpha_drop_error + eci.luminance_error;
			format_of_choice[i][0] = FMT_LUMINANCE;
		}
	}
	wsync();
}

// N-partition combination tables (:728-1093): lanes over quant levels, the combination scan inside a
// level is sequential (the reference's "<=" update order matters).
ASTC_COOP void multi_partition_find_best_combination(WCtx& w, int pc, EfTables* ef) {
	int width = pc == 2 ? 7 : pc == 3 ? 10 : 13;
	for (int quant = w.lane; quant <= QUANT_256; quant += ASTC_WARP) {
		for (int j = 0; j < width; j++) {
			ef->combined_error[quant][j] = ERROR_CALC_DEFAULT;
		}
		if (quant < QUANT_6) {
			continue;
		}
		for (int i = 0; i < 4; i++) {
			for (int j = 0; j < 4; j++) {
				int low2 = mini(i, j);
				int high2 = maxi(i, j);
				if ((high2 - low2) > 1) {
					continue;
				}
				if (pc == 2) {
					int intcnt = i + j;
					float errorterm = minf(ef->best_error[0][quant][i] + ef->best_error[1][quant][j], 1e10f);
					if (errorterm <= ef->combined_error[quant][intcnt]) {
						ef->combined_error[quant][intcnt] = errorterm;
						ef->combined_format[quant][intcnt][0] = ef->format_of_choice[0][quant][i];
						ef->combined_format[quant][intcnt][1] = ef->format_of_choice[1][quant][j];
					}
					continue;
				}
				for (int k = 0; k < 4; k++) {
					int low3 = mini(k, low2);
					int high3 = maxi(k, high2);
					if ((high3 - low3) > 1) {
						continue;
					}
					if (pc == 3) {
						int intcnt = i + j + k;
						float errorterm = minf(ef->best_error[0][quant][i] + ef->best_error[1][quant][j] + ef->best_error[2][quant][k], 1e10f);
						if (errorterm <= ef->combined_error[quant][intcnt]) {
							ef->combined_error[quant][intcnt] = errorterm;
							ef->combined_format[quant][intcnt][0] = ef->format_of_choice[0][quant][i];
							ef->combined_format[quant][intcnt][1] = ef->format_of_choice[1][quant][j];
							ef->combined_format[quant][intcnt][2] = ef->format_of_choice[2][quant][k];
						}
						continue;
					}
					for (int l = 0; l < 4; l++) {
						int low4 = mini(l, low3);
						int high4 = maxi(l, high3);
						if ((high4 - low4) > 1) {
							continue;
						}
						int intcnt = i + j + k + l;
						float errorterm = minf(ef->best_error[0][quant][i] + ef->best_error[1][quant][j] + ef->best_error[2][quant][k] + ef->best_error[3][quant][l], 1e10f);
						if (errorterm <= ef->combined_error[quant][intcnt]) {
							ef->combined_error[quant][intcnt] = errorterm;
							ef->combined_format[quant][intcnt][0] = ef->format_of_choice[0][quant][i];
							ef->combined_format[quant][intcnt][1] = ef->format_of_choice[1][quant][j];
							ef->combined_format[quant][intcnt][2] = ef->format_of_choice[2][quant][k];
							ef->combined_format[quant][intcnt][3] = ef->format_of_choice[3][quant][l];
						}
					}
				}
			}
		}
	}
	wsync();
}

// one_partition_/N-partition _find_best_combination_for_bitcount (:678-725, :768-1093)
ASTC_NOINLINE float find_best_combination_for_bitcount(int pc, const EfTables* ef, int bits_available, uint8_t& best_quant_level, uint8_t& best_quant_level_mod, uint8_t* best_formats) {
	const DevConstTables* ct = ASTC_CT;
	if (pc == 1) {
		int best_integer_count = 0;
		float best_integer_count_error = ERROR_CALC_DEFAULT;
		for (int integer_count = 1; integer_count <= 4; integer_count++) {
			int quant_level = ct->quant_mode_table[integer_count][bits_available];
			if (quant_level < QUANT_6) {
				continue;
			}
			float integer_count_error = ef->best_error[0][quant_level][integer_count - 1];
			if (integer_count_error < best_integer_count_error) {
				best_integer_count_error = integer_count_error;
				best_integer_count = integer_count - 1;
			}
		}
		int ql = ct->quant_mode_table[best_integer_count + 1][bits_available];
		best_quant_level = static_cast<uint8_t>(ql);
		best_quant_level_mod = best_quant_level;
		best_formats[0] = FMT_LUMINANCE;
		if (ql >= QUANT_6) {
			best_formats[0] = ef->format_of_choice[0][ql][best_integer_count];
		}
		return best_integer_count_error;
	}
	int best_integer_count = 0;
	float best_integer_count_error = ERROR_CALC_DEFAULT;
	int first = pc;
	int last = pc == 2 ? 8 : 9;
	int mod_bits = pc == 2 ? 2 : pc == 3 ? 5 : 8;
	for (int integer_count = first; integer_count <= last; integer_count++) {
		int quant_level = ct->quant_mode_table[integer_count][bits_available];
		if (quant_level < QUANT_6) {
			break;
		}
		float integer_count_error = ef->combined_error[quant_level][integer_count - first];
		if (integer_count_error < best_integer_count_error) {
			best_integer_count_error = integer_count_error;
			best_integer_count = integer_count;
		}
	}
	int ql = ct->quant_mode_table[best_integer_count][bits_available];
	int ql_mod = ct->quant_mode_table[best_integer_count][bits_available + mod_bits];
	best_quant_level = static_cast<uint8_t>(ql);
	best_quant_level_mod = static_cast<uint8_t>(ql_mod);
	if (ql >= QUANT_6) {
		for (int i = 0; i < pc; i++) {
			best_formats[i] = ef->combined_format[ql][best_integer_count - first][i];
		}
	} else {
		for (int i = 0; i < pc; i++) {
			best_formats[i] = FMT_LUMINANCE;
		}
	}
	return best_integer_count_error;
}

// candidate record kept in w.cand (8 bytes each)
struct Candidate {
	uint16_t block_mode;       // packed index
	uint8_t quant_level, quant_level_mod;
	uint8_t formats[4];
};

ASTC_FN int mode_bitcount(const DevBlockMode& bm, int nplanes, int pc) {
	const int8_t free_bits_for_partition_count[4] = {115 - 4, 111 - 4 - 10, 108 - 4 - 10, 105 - 4 - 10};
	return nplanes == 2 ? 109 - bm.weight_bits : free_bits_for_partition_count[pc - 1] - bm.weight_bits;
}

// compute_ideal_endpoint_formats :1096-1357. Returns the candidate count; candidates go to w.cand.
ASTC_COOP unsigned int compute_ideal_endpoint_formats(WCtx& w, const PartView& pi, int ep0slot, int ep1slot, int nplanes,
                                                      unsigned int start_block_mode, unsigned int end_block_mode) {
	const DevBsd& bsd = *w.bsd;
	int pc = (int)pi.partition_count;
	EncodingChoiceErrors eci[4];
	compute_encoding_choice_errors(w, pi, ep0slot, ep1slot, eci);
	EfTables* ef = reinterpret_cast<EfTables*>(w.su);
	for (int i = 0; i < pc; i++) {
		compute_color_error_tables(w, pi, i, eci[i], ep0slot, ep1slot, ef);
	}
	if (pc >= 2) {
		multi_partition_find_best_combination(w, pc, ef);
	}
	// total error per mode (overwrites the weight error in place)
	for (unsigned int i = start_block_mode + (unsigned int)w.lane; i < end_block_mode; i += ASTC_WARP) {
		float qwt = w.mode_err[i];
		if (qwt >= ERROR_CALC_DEFAULT) {
			w.mode_err[i] = ERROR_CALC_DEFAULT;
			continue;
		}
		uint8_t ql, qlm, fmts[4];
		float error_of_best = find_best_combination_for_bitcount(pc, ef, mode_bitcount(bsd.block_modes[i], nplanes, pc), ql, qlm, fmts);
		w.mode_err[i] = error_of_best + qwt;
	}
	wsync();
	// the tune_candidate_limit lowest totals, lowest index first among equals (:1286-1333)
	unsigned int limit = w.cfg->tune_candidate_limit;
	unsigned int count = 0;
	Candidate* cands = reinterpret_cast<Candidate*>(w.cand);
	for (unsigned int k = 0; k < limit; k++) {
		float best = ERROR_CALC_DEFAULT;
		int best_idx = 0x7FFFFFFF;
		for (unsigned int i = start_block_mode + (unsigned int)w.lane; i < end_block_mode; i += ASTC_WARP) {
			float e = w.mode_err[i];
			if (e < best) {
				best = e;
				best_idx = (int)i;
			}
		}
		wargmin(best, best_idx);
		if (!(best < ERROR_CALC_DEFAULT)) {
			break;
		}
		if (w.lane == 0) {
			w.mode_err[best_idx] = ERROR_CALC_DEFAULT;
			Candidate c;
			c.block_mode = (uint16_t)best_idx;
			c.formats[0] = c.formats[1] = c.formats[2] = c.formats[3] = 0;
			find_best_combination_for_bitcount(pc, ef, mode_bitcount(bsd.block_modes[best_idx], nplanes, pc), c.quant_level, c.quant_level_mod, c.formats);
			cands[k] = c;
		}
		count++;
		wsync();
	}
	wsync();
	return count;
}

// =============================================================================================
// Least-squares endpoint refit (astcenc_ideal_endpoints_and_weights.cpp:1099-1650)
// =============================================================================================
ASTC_NOINLINE f4 compute_rgbo_vector(f4 rgba_weight_sum, f4 weight_weight_sum, f4 rgbq_sum, float psum) {
	float X = rgba_weight_sum.x, Y = rgba_weight_sum.y, Z = rgba_weight_sum.z;
	float P = weight_weight_sum.x, Q = weight_weight_sum.y, R = weight_weight_sum.z;
	float S = psum;
	float PP = P * P, QQ = Q * Q, RR = R * R;
	float SZmRR = S * Z - RR;
	float DT = SZmRR * Y - Z * QQ;
	float YP = Y * P, QX = Q * X, YX = Y * X;
	float mZYP = -Z * YP, mZQX = -Z * QX, mRYX = -R * YX;
	float ZQP = Z * Q * P, RYP = R * YP, RQX = R * QX;
	float rdet = 1.0f / (DT * X + mZYP * P);
	f4 mat0 = mk4(DT, ZQP, RYP, mZYP);
	f4 mat1 = mk4(ZQP, SZmRR * X - Z * PP, RQX, mZQX);
	f4 mat2 = mk4(RYP, RQX, (S * Y - QQ) * X - Y * PP, mRYX);
	f4 mat3 = mk4(mZYP, mZQX, mRYX, Z * YX);
	f4 vect = rgbq_sum * rdet;
	return mk4(dot_s(mat0, vect), dot_s(mat1, vect), dot_s(mat2, vect), dot_s(mat3, vect));
}

ASTC_FN f4 sel4(f4 a, f4 b, bool m0, bool m1, bool m2, bool m3) {
	return mk4(m0 ? b.x : a.x, m1 ? b.y : a.y, m2 ? b.z : a.z, m3 ? b.w : a.w);
}

ASTC_FN void rgbo_fallback(f4& rgbo, const f4& v0, const f4& v1) {
	float dd = dot_s(rgbo, rgbo);
	if (dd != dd) {
		float avgdif = hadd_rgb_s(v1 - v0) * (1.0f / 3.0f);
		avgdif = maxf(avgdif, 0.0f);
		f4 avg = (v0 + v1) * 0.5f;
		f4 ep0 = avg - splat4(avgdif) * 0.5f;
		rgbo = mk4(ep0.x, ep0.y, ep0.z, avgdif);
	}
}

// refinement scratch inside su, laid out from the block size at run time
struct RefineScratch {
	float* undec[2];      // [T] undecimated float weights per plane
	float* texel_err;     // [T] per-texel error terms for the ordered sums
	float* uqf;           // [64] realign: float copy of the quantised weights
	float* stage;         // [12][stage_stride] realign: per-texel error vectors of one weight
	int stage_stride;
	uint8_t* iw[2];       // [T] integer undecimated weights (0..64) per plane
};

ASTC_FN RefineScratch make_refine_scratch(const WCtx& w) {
	RefineScratch r;
	int Tp = (w.T + 3) & ~3;
	float* f = reinterpret_cast<float*>(w.su);
	r.undec[0] = f;
	r.undec[1] = f + Tp;
	r.texel_err = f + 2 * Tp;
	r.uqf = f + 3 * Tp;
	r.stage = f + 3 * Tp + 64;
	r.stage_stride = (int)w.bsd->max_weight_texel_count;
	uint8_t* b = reinterpret_cast<uint8_t*>(r.stage + 12 * r.stage_stride);
	r.iw[0] = b;
	r.iw[1] = b + Tp;
	return r;
}

// undecimate the quantised weights of `planes` planes: lanes over texels
ASTC_COOP void undecimate_weights(WCtx& w, const DecView& di, const uint8_t* uquant, int planes, RefineScratch* rs) {
	int T = w.T;
	for (int id = w.lane; id < T * planes; id += ASTC_WARP) {
		int pl = id >= T ? 1 : 0;
		int t = id - pl * T;
		const uint8_t* uq = uquant + pl * 32;
		float v;
		if (di.max_twc == 1) {
			v = static_cast<float>(uq[t]) * (1.0f / 64.0f);
		} else if (di.max_twc <= 2) {
			v = (static_cast<float>(uq[di.tw[t]]) * (1.0f / 64.0f)) * contrib_f(di.tc[t]) +
			    (static_cast<float>(uq[di.tw[T + t]]) * (1.0f / 64.0f)) * contrib_f(di.tc[T + t]);
		} else {
			v = ((static_cast<float>(uq[di.tw[t]]) * (1.0f / 64.0f)) * contrib_f(di.tc[t]) +
			     (static_cast<float>(uq[di.tw[T + t]]) * (1.0f / 64.0f)) * contrib_f(di.tc[T + t])) +
			    ((static_cast<float>(uq[di.tw[2 * T + t]]) * (1.0f / 64.0f)) * contrib_f(di.tc[2 * T + t]) +
			     (static_cast<float>(uq[di.tw[3 * T + t]]) * (1.0f / 64.0f)) * contrib_f(di.tc[3 * T + t]));
		}
		rs->undec[pl][t] = v;
	}
	wsync();
}

// recompute_ideal_colors_1plane :1146-1366. Chains per partition:
//   0 left, 1 middle, 2 right, 3 weight_weight, 4-7 color_vec_x, 8-11 color_vec_y, 12-13 scale_vec, 14-17 rgba_sum
ASTC_COOP void recompute_ideal_colors_1plane(WCtx& w, const PartView& pi, const DecView& di, RefineScratch* rs) {
	unsigned int pc = pi.partition_count;
	undecimate_weights(w, di, w.work_weights, 1, rs);
	const float* undec = rs->undec[0];
	f4 color_weight = w.bi.channel_weight;
	float ls_weight = hadd_rgb_s(color_weight);
	// phase A: per-partition colour sums (needed for scale_dir) - chains 14-17
	if (pc > 1) {
		for (int id = w.lane; id < (int)pc * 4; id += ASTC_WARP) {
			int p = id >> 2;
			int c = id & 3;
			const uint8_t* tix = pi.texels + pi.start[p];
			int n = pi.count[p];
			const float* d = w.blk[c];
			float s = 0.0f;
			for (int j = 0; j < n; j++) {
				s = s + d[tix[j]];
			}
			w.tmpf[96 + id] = s;
		}
		wsync();
	}
	f4 scale_dir[4], rgba_weight_sum[4];
	for (unsigned int p = 0; p < pc; p++) {
		f4 rgba_sum = pc > 1 ? mk4(w.tmpf[96 + p * 4], w.tmpf[96 + p * 4 + 1], w.tmpf[96 + p * 4 + 2], w.tmpf[96 + p * 4 + 3])
		                     : w.bi.data_mean * static_cast<float>(w.T);
		rgba_sum = rgba_sum * color_weight;
		rgba_weight_sum[p] = max4(color_weight * static_cast<float>(pi.count[p]), splat4(1e-17f));
		f4 q = rgba_sum / rgba_weight_sum[p];
		scale_dir[p] = normalize4(mk4(q.x, q.y, q.z, 0.0f));
	}
	// phase B: the weighted sums - 14 chains per partition; min/max terms by lanes over texels
	for (int id = w.lane; id < (int)pc * 14; id += ASTC_WARP) {
		int p = id / 14;
		int ch = id - p * 14;
		const uint8_t* tix = pi.texels + pi.start[p];
		int n = pi.count[p];
		f4 sd = scale_dir[p];
		float s = ch == 3 ? 1e-17f : 0.0f;
		for (int j = 0; j < n; j++) {
			int t = tix[j];
			float idx0 = undec[t];
			float om_idx0 = 1.0f - idx0;
			float term;
			if (ch == 0) term = om_idx0 * om_idx0;
			else if (ch == 1) term = om_idx0 * idx0;
			else if (ch == 2) term = idx0 * idx0;
			else if (ch == 3) term = idx0;
			else if (ch < 8) {
				float cw = w.blk[ch - 4][t];
				float cwi = cw * idx0;
				term = cw - cwi;
			} else if (ch < 12) {
				term = w.blk[ch - 8][t] * idx0;
			} else {
				f4 rgba = texel4(w, t);
				float scale = dot3_s(sd, rgba);
				term = (ch == 12 ? om_idx0 : idx0) * (scale * ls_weight);
			}
			s = s + term;
		}
		w.tmpf[id] = s;
	}
	float wmin1[4], wmax1[4], scale_min[4], scale_max[4];
	for (unsigned int p = 0; p < pc; p++) {
		const uint8_t* tix = pi.texels + pi.start[p];
		int n = pi.count[p];
		float a = 1.0f, b = 0.0f, c = 1e10f, d = 0.0f;
		for (int j = w.lane; j < n; j += ASTC_WARP) {
			int t = tix[j];
			float idx0 = undec[t];
			a = minf(idx0, a);
			b = maxf(idx0, b);
			float scale = dot3_s(scale_dir[p], texel4(w, t));
			c = minf(scale, c);
			d = maxf(scale, d);
		}
		wmin1[p] = wmin_f(a);
		wmax1[p] = wmax_f(b);
		scale_min[p] = wmin_f(c);
		scale_max[p] = wmax_f(d);
	}
	wsync();
	// phase C: the solves, one lane per partition
	for (unsigned int i = (unsigned int)w.lane; i < pc; i += ASTC_WARP) {
		const float* a = w.tmpf + i * 14;
		float left_sum_s = a[0], middle_sum_s = a[1], right_sum_s = a[2], weight_weight_sum_s = a[3];
		f4 color_vec_x = mk4(a[4], a[5], a[6], a[7]);
		f4 color_vec_y = mk4(a[8], a[9], a[10], a[11]);
		f4 scale_vec = mk4(a[12], a[13], 0.0f, 0.0f);
		f4 left_sum = splat4(left_sum_s) * color_weight;
		f4 middle_sum = splat4(middle_sum_s) * color_weight;
		f4 right_sum = splat4(right_sum_s) * color_weight;
		f4 lmrs_sum = mk4(left_sum_s, middle_sum_s, right_sum_s, 0.0f) * ls_weight;
		color_vec_x = color_vec_x * color_weight;
		color_vec_y = color_vec_y * color_weight;
		float scalediv = scale_min[i] / maxf(scale_max[i], 1e-10f);
		scalediv = clamp1f(scalediv);
		f4 sds = scale_dir[i] * scale_max[i];
		f4 rgbs = mk4(sds.x, sds.y, sds.z, scalediv);
		f4 e0 = w.ep[EP_WORK_0 + i], e1 = w.ep[EP_WORK_1 + i];
		if (wmin1[i] >= wmax1[i] * 0.999f) {
			f4 avg = (color_vec_x + color_vec_y) / rgba_weight_sum[i];
			e0 = sel4(e0, avg, avg.x == avg.x, avg.y == avg.y, avg.z == avg.z, avg.w == avg.w);
			e1 = sel4(e1, avg, avg.x == avg.x, avg.y == avg.y, avg.z == avg.z, avg.w == avg.w);
			rgbs = mk4(sds.x, sds.y, sds.z, 1.0f);
		} else {
			f4 color_det1 = (left_sum * right_sum) - (middle_sum * middle_sum);
			f4 color_rdet1 = splat4(1.0f) / color_det1;
			float ls_det1 = (lmrs_sum.x * lmrs_sum.z) - (lmrs_sum.y * lmrs_sum.y);
			float ls_rdet1 = 1.0f / ls_det1;
			f4 color_mss1 = (left_sum * left_sum) + (splat4(2.0f) * middle_sum * middle_sum) + (right_sum * right_sum);
			float ls_mss1 = (lmrs_sum.x * lmrs_sum.x) + (2.0f * lmrs_sum.y * lmrs_sum.y) + (lmrs_sum.z * lmrs_sum.z);
			f4 ep0 = (right_sum * color_vec_x - middle_sum * color_vec_y) * color_rdet1;
			f4 ep1 = (left_sum * color_vec_y - middle_sum * color_vec_x) * color_rdet1;
			f4 thr = color_mss1 * 1e-4f;
			bool m[4];
			for (int c = 0; c < 4; c++) {
				bool det = absf(lane(color_det1, c)) > lane(thr, c);
				bool notnan = (lane(ep0, c) == lane(ep0, c)) && (lane(ep1, c) == lane(ep1, c));
				m[c] = det && notnan;
			}
			e0 = sel4(e0, ep0, m[0], m[1], m[2], m[3]);
			e1 = sel4(e1, ep1, m[0], m[1], m[2], m[3]);
			float scale_ep0 = (lmrs_sum.z * scale_vec.x - lmrs_sum.y * scale_vec.y) * ls_rdet1;
			float scale_ep1 = (lmrs_sum.x * scale_vec.y - lmrs_sum.y * scale_vec.x) * ls_rdet1;
			if (fabsf(ls_det1) > (ls_mss1 * 1e-4f) && scale_ep0 == scale_ep0 && scale_ep1 == scale_ep1 && scale_ep0 < scale_ep1) {
				float scalediv2 = scale_ep0 / scale_ep1;
				f4 sdsm = scale_dir[i] * scale_ep1;
				rgbs = mk4(sdsm.x, sdsm.y, sdsm.z, scalediv2);
			}
		}
		w.ep[EP_WORK_0 + i] = e0;
		w.ep[EP_WORK_1 + i] = e1;
		w.ep[EP_RGBS + i] = rgbs;
		if (w.bi.rgb_lns0 || w.bi.alpha_lns0) {
			f4 weight_weight_sum = splat4(weight_weight_sum_s) * color_weight;
			float psum = right_sum_s * hadd_rgb_s(color_weight);
			f4 rgbq_sum = color_vec_x + color_vec_y;
			rgbq_sum.w = hadd_rgb_s(color_vec_y);
			f4 rgbovec = compute_rgbo_vector(rgba_weight_sum[i], weight_weight_sum, rgbq_sum, psum);
			rgbo_fallback(rgbovec, e0, e1);
			w.ep[EP_RGBO + i] = rgbovec;
		}
	}
	wsync();
}

// recompute_ideal_colors_2planes :1369-1650. Chains:
//   0-2 left/middle/right plane 1, 3-5 plane 2, 6-9 color_vec_x, 10-13 color_vec_y, 14-15 scale_vec, 16-19 weight_weight_sum
ASTC_COOP void recompute_ideal_colors_2planes(WCtx& w, const DecView& di, int plane2_component, RefineScratch* rs) {
	int T = w.T;
	undecimate_weights(w, di, w.work_weights, 2, rs);
	const float* undec1 = rs->undec[0];
	const float* undec2 = rs->undec[1];
	f4 color_weight = w.bi.channel_weight;
	float ls_weight = hadd_rgb_s(color_weight);
	f4 rgba_weight_sum = max4(color_weight * static_cast<float>(T), splat4(1e-17f));
	f4 scale_dir = normalize4(mk4(w.bi.data_mean.x, w.bi.data_mean.y, w.bi.data_mean.z, 0.0f));
	for (int ch = w.lane; ch < 20; ch += ASTC_WARP) {
		float s = ch >= 16 ? 1e-17f : 0.0f;
		for (int j = 0; j < T; j++) {
			float idx0 = undec1[j];
			float om_idx0 = 1.0f - idx0;
			float idx1 = undec2[j];
			float om_idx1 = 1.0f - idx1;
			float term;
			if (ch == 0) term = om_idx0 * om_idx0;
			else if (ch == 1) term = om_idx0 * idx0;
			else if (ch == 2) term = idx0 * idx0;
			else if (ch == 3) term = om_idx1 * om_idx1;
			else if (ch == 4) term = om_idx1 * idx1;
			else if (ch == 5) term = idx1 * idx1;
			else if (ch < 10) {
				int c = ch - 6;
				float color_idx = c == plane2_component ? idx1 : idx0;
				float cw = w.blk[c][j];
				float cwi = cw * color_idx;
				term = cw - cwi;
			} else if (ch < 14) {
				int c = ch - 10;
				float color_idx = c == plane2_component ? idx1 : idx0;
				term = w.blk[c][j] * color_idx;
			} else if (ch < 16) {
				float scale = dot3_s(scale_dir, texel4(w, j));
				term = (ch == 14 ? om_idx0 : idx0) * (ls_weight * scale);
			} else {
				int c = ch - 16;
				term = c == plane2_component ? idx1 : idx0;
			}
			s = s + term;
		}
		w.tmpf[ch] = s;
	}
	float a = 1.0f, b = 0.0f, a2 = 1.0f, b2 = 0.0f, c = 1e10f, d = 0.0f;
	for (int j = w.lane; j < T; j += ASTC_WARP) {
		float idx0 = undec1[j];
		float idx1 = undec2[j];
		a = minf(idx0, a);
		b = maxf(idx0, b);
		a2 = minf(idx1, a2);
		b2 = maxf(idx1, b2);
		float scale = dot3_s(scale_dir, texel4(w, j));
		c = minf(scale, c);
		d = maxf(scale, d);
	}
	float wmin1 = wmin_f(a), wmax1 = wmax_f(b), wmin2 = wmin_f(a2), wmax2 = wmax_f(b2);
	float scale_min = wmin_f(c), scale_max = wmax_f(d);
	wsync();
	if (w.lane == 0) {
		const float* t = w.tmpf;
		bool p2[4] = {plane2_component == 0, plane2_component == 1, plane2_component == 2, plane2_component == 3};
		float left1_sum_s = t[0], middle1_sum_s = t[1], right1_sum_s = t[2];
		float left2_sum_s = t[3], middle2_sum_s = t[4], right2_sum_s = t[5];
		f4 color_vec_x = mk4(t[6], t[7], t[8], t[9]);
		f4 color_vec_y = mk4(t[10], t[11], t[12], t[13]);
		f4 scale_vec = mk4(t[14], t[15], 0.0f, 0.0f);
		f4 weight_weight_sum = mk4(t[16], t[17], t[18], t[19]);
		f4 left1_sum = splat4(left1_sum_s) * color_weight;
		f4 middle1_sum = splat4(middle1_sum_s) * color_weight;
		f4 right1_sum = splat4(right1_sum_s) * color_weight;
		f4 lmrs_sum = mk4(left1_sum_s, middle1_sum_s, right1_sum_s, 0.0f) * ls_weight;
		f4 left2_sum = splat4(left2_sum_s) * color_weight;
		f4 middle2_sum = splat4(middle2_sum_s) * color_weight;
		f4 right2_sum = splat4(right2_sum_s) * color_weight;
		color_vec_x = color_vec_x * color_weight;
		color_vec_y = color_vec_y * color_weight;
		float scalediv = scale_min / maxf(scale_max, 1e-10f);
		scalediv = clamp1f(scalediv);
		f4 sds = scale_dir * scale_max;
		f4 rgbs_vector = mk4(sds.x, sds.y, sds.z, scalediv);
		f4 e0 = w.ep[EP_WORK_0], e1 = w.ep[EP_WORK_1];
		if (wmin1 >= wmax1 * 0.999f) {
			f4 avg = (color_vec_x + color_vec_y) / rgba_weight_sum;
			bool m[4];
			for (int k = 0; k < 4; k++) {
				m[k] = !p2[k] && (lane(avg, k) == lane(avg, k));
			}
			e0 = sel4(e0, avg, m[0], m[1], m[2], m[3]);
			e1 = sel4(e1, avg, m[0], m[1], m[2], m[3]);
			rgbs_vector = mk4(sds.x, sds.y, sds.z, 1.0f);
		} else {
			f4 color_det1 = (left1_sum * right1_sum) - (middle1_sum * middle1_sum);
			f4 color_rdet1 = splat4(1.0f) / color_det1;
			float ls_det1 = (lmrs_sum.x * lmrs_sum.z) - (lmrs_sum.y * lmrs_sum.y);
			float ls_rdet1 = 1.0f / ls_det1;
			f4 color_mss1 = (left1_sum * left1_sum) + (splat4(2.0f) * middle1_sum * middle1_sum) + (right1_sum * right1_sum);
			float ls_mss1 = (lmrs_sum.x * lmrs_sum.x) + (2.0f * lmrs_sum.y * lmrs_sum.y) + (lmrs_sum.z * lmrs_sum.z);
			f4 ep0 = (right1_sum * color_vec_x - middle1_sum * color_vec_y) * color_rdet1;
			f4 ep1 = (left1_sum * color_vec_y - middle1_sum * color_vec_x) * color_rdet1;
			float scale_ep0 = (lmrs_sum.z * scale_vec.x - lmrs_sum.y * scale_vec.y) * ls_rdet1;
			float scale_ep1 = (lmrs_sum.x * scale_vec.y - lmrs_sum.y * scale_vec.x) * ls_rdet1;
			f4 thr = color_mss1 * 1e-4f;
			bool m[4];
			for (int k = 0; k < 4; k++) {
				bool det = absf(lane(color_det1, k)) > lane(thr, k);
				bool notnan = (lane(ep0, k) == lane(ep0, k)) && (lane(ep1, k) == lane(ep1, k));
				m[k] = !p2[k] && det && notnan;
			}
			e0 = sel4(e0, ep0, m[0], m[1], m[2], m[3]);
			e1 = sel4(e1, ep1, m[0], m[1], m[2], m[3]);
			if (fabsf(ls_det1) > (ls_mss1 * 1e-4f) && scale_ep0 == scale_ep0 && scale_ep1 == scale_ep1 && scale_ep0 < scale_ep1) {
				float scalediv2 = scale_ep0 / scale_ep1;
				f4 sdsm = scale_dir * scale_ep1;
				rgbs_vector = mk4(sdsm.x, sdsm.y, sdsm.z, scalediv2);
			}
		}
		if (wmin2 >= wmax2 * 0.999f) {
			f4 avg = (color_vec_x + color_vec_y) / rgba_weight_sum;
			bool m[4];
			for (int k = 0; k < 4; k++) {
				m[k] = p2[k] && (lane(avg, k) == lane(avg, k));
			}
			e0 = sel4(e0, avg, m[0], m[1], m[2], m[3]);
			e1 = sel4(e1, avg, m[0], m[1], m[2], m[3]);
		} else {
			f4 color_det2 = (left2_sum * right2_sum) - (middle2_sum * middle2_sum);
			f4 color_rdet2 = splat4(1.0f) / color_det2;
			f4 color_mss2 = (left2_sum * left2_sum) + (splat4(2.0f) * middle2_sum * middle2_sum) + (right2_sum * right2_sum);
			f4 ep0 = (right2_sum * color_vec_x - middle2_sum * color_vec_y) * color_rdet2;
			f4 ep1 = (left2_sum * color_vec_y - middle2_sum * color_vec_x) * color_rdet2;
			f4 thr = color_mss2 * 1e-4f;
			bool m[4];
			for (int k = 0; k < 4; k++) {
				bool det = absf(lane(color_det2, k)) > lane(thr, k);
				bool notnan = (lane(ep0, k) == lane(ep0, k)) && (lane(ep1, k) == lane(ep1, k));
				m[k] = p2[k] && det && notnan;
			}
			e0 = sel4(e0, ep0, m[0], m[1], m[2], m[3]);
			e1 = sel4(e1, ep1, m[0], m[1], m[2], m[3]);
		}
		w.ep[EP_WORK_0] = e0;
		w.ep[EP_WORK_1] = e1;
		w.ep[EP_RGBS] = rgbs_vector;
		if (w.bi.rgb_lns0 || w.bi.alpha_lns0) {
			weight_weight_sum = weight_weight_sum * color_weight;
			f4 rsel = mk4(p2[0] ? right2_sum.x : right1_sum.x, p2[1] ? right2_sum.y : right1_sum.y, p2[2] ? right2_sum.z : right1_sum.z,
			              p2[3] ? right2_sum.w : right1_sum.w);
			float psum = dot3_s(rsel, color_weight);
			f4 rgbq_sum = color_vec_x + color_vec_y;
			rgbq_sum.w = hadd_rgb_s(color_vec_y);
			f4 rgbo_vector = compute_rgbo_vector(rgba_weight_sum, weight_weight_sum, rgbq_sum, psum);
			rgbo_fallback(rgbo_vector, e0, e1);
			w.ep[EP_RGBO] = rgbo_vector;
		}
	}
	wsync();
}

// =============================================================================================
// Decompress-and-diff scoring (astcenc_decompress_symbolic.cpp:89-618)
// =============================================================================================
ASTC_FN bool u8_mask(const WCtx& w) { return w.bi.decode_unorm8 || w.cfg->profile == PRF_LDR_SRGB; }

ASTC_FN int lerp1(bool u8, int c0, int c1, int w1) {   // lerp_color_int :37-61
	int w0 = 64 - w1;
	int color = (c0 * w0) + (c1 * w1) + 32;
	color = color >> 6;
	if (u8) {
		color = (color >> 8) * 257;
	}
	return color;
}

// unpack_weights :89-167: integer infill into rs->iw
ASTC_COOP void unpack_weights(WCtx& w, const DecView& di, const uint8_t* weights, int planes, RefineScratch* rs) {
	int T = w.T;
	for (int id = w.lane; id < T * planes; id += ASTC_WARP) {
		int pl = id >= T ? 1 : 0;
		int t = id - pl * T;
		const uint8_t* uq = weights + pl * 32;
		int s = 8;
		s += uq[di.tw[t]] * di.tc[t];
		s += uq[di.tw[T + t]] * di.tc[T + t];
		s += uq[di.tw[2 * T + t]] * di.tc[2 * T + t];
		s += uq[di.tw[3 * T + t]] * di.tc[3 * T + t];
		rs->iw[pl][t] = (uint8_t)(s >> 4);
	}
	wsync();
}

// compute_symbolic_block_difference_{2plane,1plane,1plane_1partition} (:313-618) on the candidate in
// w.work_weights / w.work_colors with header hdr. Per-texel terms by lanes over texels, then the
// reference's summation order: 4-lane accumulator (1 partition, 1 plane, no RGBM) or one scalar chain.
ASTC_COOP float compute_symbolic_block_difference(WCtx& w, const ScbHdr& hdr, const PartView& pi, const DecView& di, bool dual, RefineScratch* rs) {
	if (hdr.block_type == SYM_BTYPE_ERROR) {
		return ERROR_CALC_DEFAULT;
	}
	const DevConfig& cfg = *w.cfg;
	int T = w.T;
	unsigned int pc = hdr.partition_count;
	bool rgbm = (cfg.flags & FLG_MAP_RGBM) != 0;
	bool fast = !dual && pc == 1 && !rgbm;
	unpack_weights(w, di, w.work_weights, dual ? 2 : 1, rs);
	bool u8 = u8_mask(w);
	i4 ep0[4], ep1[4];
	for (unsigned int p = 0; p < pc; p++) {
		bool rgb_lns, a_lns;
		unpack_color_endpoints(cfg.profile, hdr.color_formats[p], w.work_colors + p * 8, rgb_lns, a_lns, ep0[p], ep1[p]);
	}
	f4 cw = w.bi.channel_weight;
	bool reject = false;
	for (int t = w.lane; t < T; t += ASTC_WARP) {
		int p = pc > 1 ? pi.partition_of_texel[t] : 0;
		int w1 = rs->iw[0][t];
		int w2 = dual ? rs->iw[1][t] : w1;
		int pc2 = hdr.plane2_component;
		i4 e0 = ep0[0], e1 = ep1[0];
		if (p == 1) { e0 = ep0[1]; e1 = ep1[1]; }
		else if (p == 2) { e0 = ep0[2]; e1 = ep1[2]; }
		else if (p == 3) { e0 = ep0[3]; e1 = ep1[3]; }
		float cr = (float)lerp1(u8, e0.x, e1.x, (dual && pc2 == 0) ? w2 : w1);
		float cg = (float)lerp1(u8, e0.y, e1.y, (dual && pc2 == 1) ? w2 : w1);
		float cb = (float)lerp1(u8, e0.z, e1.z, (dual && pc2 == 2) ? w2 : w1);
		float ca = (float)lerp1(u8, e0.w, e1.w, (dual && pc2 == 3) ? w2 : w1);
		float orr = w.blk[0][t], og = w.blk[1][t], ob = w.blk[2][t], oa = w.blk[3][t];
		float metric;
		if (fast) {
			float er = minf(absf(orr - cr), 1e15f);
			float eg = minf(absf(og - cg), 1e15f);
			float eb = minf(absf(ob - cb), 1e15f);
			float ea = minf(absf(oa - ca), 1e15f);
			er = er * er;
			eg = eg * eg;
			eb = eb * eb;
			ea = ea * ea;
			metric = er * cw.x + eg * cw.y + eb * cw.z + ea * cw.w;
		} else {
			f4 color = mk4(cr, cg, cb, ca);
			f4 old = mk4(orr, og, ob, oa);
			if (rgbm) {
				if (color.w == 0.0f) {
					reject = true;
				}
				color = mk4(color.x * color.w * cfg.rgbm_m_scale, color.y * color.w * cfg.rgbm_m_scale, color.z * color.w * cfg.rgbm_m_scale, 1.0f);
				old = mk4(old.x * old.w * cfg.rgbm_m_scale, old.y * old.w * cfg.rgbm_m_scale, old.z * old.w * cfg.rgbm_m_scale, 1.0f);
			}
			f4 error = old - color;
			error = min4(mk4(absf(error.x), absf(error.y), absf(error.z), absf(error.w)), splat4(1e15f));
			error = error * error;
			metric = minf(dot_s(error, cw), ERROR_CALC_DEFAULT);
		}
		rs->texel_err[t] = metric;
	}
	// The reference returns -1e30 at the first texel (in its iteration order) whose decoded alpha is 0; any
	// such texel makes the result -1e30, so the order does not matter for the rejection itself.
	reject = wany(reject);
	wsync();
	if (reject) {
		return -ERROR_CALC_DEFAULT;
	}
	if (fast) {
		for (int l = w.lane; l < 4; l += ASTC_WARP) {
			float s = 0.0f;
			for (int t = l; t < T; t += 4) {
				s = s + rs->texel_err[t];
			}
			w.tmpf[l] = s;
		}
		wsync();
		float r = (w.tmpf[0] + w.tmpf[2]) + (w.tmpf[1] + w.tmpf[3]);
		wsync();
		return r;
	}
	if (w.lane == 0) {
		float summa = 0.0f;
		if (dual || pc == 1) {
			for (int t = 0; t < T; t++) {
				summa += rs->texel_err[t];
			}
		} else {
			for (int t = 0; t < T; t++) {
				summa += rs->texel_err[pi.texels[t]];
			}
		}
		w.tmpf[0] = summa;
	}
	wsync();
	float r = w.tmpf[0];
	wsync();
	return r;
}

// =============================================================================================
// Weight realignment (astcenc_compress_symbolic.cpp:69-350)
// =============================================================================================
ASTC_COOP bool realign_weights(WCtx& w, const ScbHdr& hdr, const PartView& pi, const DevBlockMode& bm, const DecView& di, RefineScratch* rs) {
	const DevConstTables* ct = ASTC_CT;
	const DevConfig& cfg = *w.cfg;
	unsigned int pc = hdr.partition_count;
	const uint16_t* prev_next = ct->wq_prev_next[bm.quant_mode];
	int weight_count = di.W;
	int T = w.T;
	bool decimated = weight_count != T;
	unsigned int max_plane = bm.is_dual_plane;
	int plane2_component = hdr.plane2_component;
	i4 endpnt0[4], endpnt1[4];
	for (unsigned int p = 0; p < pc; p++) {
		bool rgb_hdr, alpha_hdr;
		unpack_color_endpoints(cfg.profile, hdr.color_formats[p], w.work_colors + p * 8, rgb_hdr, alpha_hdr, endpnt0[p], endpnt1[p]);
	}
	f4 ew = w.bi.channel_weight;
	bool adjustments = false;
	uint8_t* dec_weights_uquant = w.work_weights;
	for (unsigned int pl = 0; pl <= max_plane; pl++) {
		// plane_mask lanes are zeroed: for plane 1 that is the plane-2 component, for plane 2 all others
		f4 endpnt0f[4], offset[4];
		for (unsigned int p = 0; p < pc; p++) {
			i4 epd = mki4(endpnt1[p].x - endpnt0[p].x, endpnt1[p].y - endpnt0[p].y, endpnt1[p].z - endpnt0[p].z, endpnt1[p].w - endpnt0[p].w);
			bool m0 = (plane2_component == 0) != (pl == 1);
			bool m1 = (plane2_component == 1) != (pl == 1);
			bool m2 = (plane2_component == 2) != (pl == 1);
			bool m3 = (plane2_component == 3) != (pl == 1);
			if (m0) epd.x = 0;
			if (m1) epd.y = 0;
			if (m2) epd.z = 0;
			if (m3) epd.w = 0;
			endpnt0f[p] = mk4((float)endpnt0[p].x, (float)endpnt0[p].y, (float)endpnt0[p].z, (float)endpnt0[p].w);
			offset[p] = mk4((float)epd.x, (float)epd.y, (float)epd.z, (float)epd.w) * (1.0f / 64.0f);
		}
		if (!decimated) {
			// realign_weights_undecimated :69-185 - texels are independent
			for (int texel = w.lane; texel < T; texel += ASTC_WARP) {
				int uqw = dec_weights_uquant[texel];
				uint32_t pn = prev_next[uqw];
				int uqw_down = pn & 0xFF;
				int uqw_up = (pn >> 8) & 0xFF;
				float weight_base = static_cast<float>(uqw);
				float weight_down = static_cast<float>(uqw_down - uqw);
				float weight_up = static_cast<float>(uqw_up - uqw);
				int partition = pc > 1 ? pi.partition_of_texel[texel] : 0;
				f4 color_offset = offset[0], color_base = endpnt0f[0];
				if (partition == 1) { color_offset = offset[1]; color_base = endpnt0f[1]; }
				else if (partition == 2) { color_offset = offset[2]; color_base = endpnt0f[2]; }
				else if (partition == 3) { color_offset = offset[3]; color_base = endpnt0f[3]; }
				f4 color = color_base + color_offset * weight_base;
				f4 orig_color = texel4(w, texel);
				f4 color_diff = color - orig_color;
				f4 color_diff_down = color_diff + color_offset * weight_down;
				f4 color_diff_up = color_diff + color_offset * weight_up;
				float error_base = dot_s(color_diff * color_diff, ew);
				float error_down = dot_s(color_diff_down * color_diff_down, ew);
				float error_up = dot_s(color_diff_up * color_diff_up, ew);
				if ((error_up < error_base) && (error_up < error_down) && (uqw < 64)) {
					dec_weights_uquant[texel] = static_cast<uint8_t>(uqw_up);
					adjustments = true;
				} else if ((error_down < error_base) && (uqw > 0)) {
					dec_weights_uquant[texel] = static_cast<uint8_t>(uqw_down);
					adjustments = true;
				}
			}
			wsync();
		} else {
			// realign_weights_decimated :188-350 - weights are visited in order; a changed weight feeds the
			// following ones. Per weight: lanes over its texels, then 12 ordered chains (3 vectors x 4 lanes).
			for (int we = w.lane; we < weight_count; we += ASTC_WARP) {
				rs->uqf[we] = static_cast<float>(dec_weights_uquant[we]);
			}
			wsync();
			for (int we = 0; we < weight_count; we++) {
				int uqw = dec_weights_uquant[we];
				uint32_t pn = prev_next[uqw];
				float uqw_base = rs->uqf[we];
				float uqw_down = static_cast<float>(pn & 0xFF);
				float uqw_up = static_cast<float>((pn >> 8) & 0xFF);
				float uqw_diff_down = uqw_down - uqw_base;
				float uqw_diff_up = uqw_up - uqw_base;
				int off = di.wto[we];
				int cnt = di.wto[we + 1] - off;
				for (int te = w.lane; te < cnt; te += ASTC_WARP) {
					int texel = di.wt[off + te];
					float tw_base = contrib_f(di.wc[off + te]);
					float weight_base = (rs->uqf[di.tw[texel]] * contrib_f(di.tc[texel]) + rs->uqf[di.tw[T + texel]] * contrib_f(di.tc[T + texel])) +
					                    (rs->uqf[di.tw[2 * T + texel]] * contrib_f(di.tc[2 * T + texel]) + rs->uqf[di.tw[3 * T + texel]] * contrib_f(di.tc[3 * T + texel]));
					float weight_down = weight_base + uqw_diff_down * tw_base - weight_base;
					float weight_up = weight_base + uqw_diff_up * tw_base - weight_base;
					int partition = pc > 1 ? pi.partition_of_texel[texel] : 0;
					f4 color_offset = offset[0], color_base = endpnt0f[0];
					if (partition == 1) { color_offset = offset[1]; color_base = endpnt0f[1]; }
					else if (partition == 2) { color_offset = offset[2]; color_base = endpnt0f[2]; }
					else if (partition == 3) { color_offset = offset[3]; color_base = endpnt0f[3]; }
					f4 color = color_base + color_offset * weight_base;
					f4 orig_color = texel4(w, texel);
					f4 color_diff = color - orig_color;
					f4 color_down_diff = color_diff + color_offset * weight_down;
					f4 color_up_diff = color_diff + color_offset * weight_up;
					f4 b = color_diff * color_diff;
					f4 dn = color_down_diff * color_down_diff;
					f4 up = color_up_diff * color_up_diff;
					rs->stage[0 * rs->stage_stride + te] = b.x; rs->stage[1 * rs->stage_stride + te] = b.y; rs->stage[2 * rs->stage_stride + te] = b.z; rs->stage[3 * rs->stage_stride + te] = b.w;
					rs->stage[4 * rs->stage_stride + te] = dn.x; rs->stage[5 * rs->stage_stride + te] = dn.y; rs->stage[6 * rs->stage_stride + te] = dn.z; rs->stage[7 * rs->stage_stride + te] = dn.w;
					rs->stage[8 * rs->stage_stride + te] = up.x; rs->stage[9 * rs->stage_stride + te] = up.y; rs->stage[10 * rs->stage_stride + te] = up.z; rs->stage[11 * rs->stage_stride + te] = up.w;
				}
				wsync();
				for (int ch = w.lane; ch < 12; ch += ASTC_WARP) {
					float s = 0.0f;
					for (int te = 0; te < cnt; te++) {
						s = s + rs->stage[ch * rs->stage_stride + te];
					}
					w.tmpf[ch] = s * lane(ew, ch & 3);
				}
				wsync();
				const float* t = w.tmpf;
				float error_base = (t[0] + t[2]) + (t[1] + t[3]);
				float error_down = (t[4] + t[6]) + (t[5] + t[7]);
				float error_up = (t[8] + t[10]) + (t[9] + t[11]);
				if ((error_up < error_base) && (error_up < error_down) && (uqw < 64)) {
					if (w.lane == 0) {
						rs->uqf[we] = uqw_up;
						dec_weights_uquant[we] = static_cast<uint8_t>(uqw_up);
					}
					adjustments = true;
				} else if ((error_down < error_base) && (uqw > 0)) {
					if (w.lane == 0) {
						rs->uqf[we] = uqw_down;
						dec_weights_uquant[we] = static_cast<uint8_t>(uqw_down);
					}
					adjustments = true;
				}
				wsync();
			}
		}
		dec_weights_uquant += 32;
	}
	return wany(adjustments);
}

// =============================================================================================
// Physical block packing (astcenc_symbolic_physical.cpp:102-286, astcenc_integer_sequence.cpp:493-648).
// Pure bit twiddling on ~100 values: executed by lane 0.
// =============================================================================================
// The 128-bit block is assembled in two 64-bit registers; fields never overlap, so OR-ing a field in is
// equivalent to the reference's masked byte writes (symbolic_physical.cpp:63-85).
struct Bits128 {
	uint64_t lo, hi;
};

ASTC_FN void put_bits(Bits128& b, unsigned int value, unsigned int bitcount, unsigned int bitoffset) {
	uint64_t v = (uint64_t)(value & ((1u << bitcount) - 1u));
	if (bitoffset < 64) {
		b.lo |= v << bitoffset;
		if (bitoffset + bitcount > 64) {
			b.hi |= v >> (64 - bitoffset);
		}
	} else if (bitoffset < 128) {
		b.hi |= v << (bitoffset - 64);
	}
}

ASTC_FN uint64_t brev64(uint64_t v) {
#if defined(ASTC_HOSTSIM)
	v = ((v >> 1) & 0x5555555555555555ULL) | ((v & 0x5555555555555555ULL) << 1);
	v = ((v >> 2) & 0x3333333333333333ULL) | ((v & 0x3333333333333333ULL) << 2);
	v = ((v >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((v & 0x0F0F0F0F0F0F0F0FULL) << 4);
	v = ((v >> 8) & 0x00FF00FF00FF00FFULL) | ((v & 0x00FF00FF00FF00FFULL) << 8);
	v = ((v >> 16) & 0x0000FFFF0000FFFFULL) | ((v & 0x0000FFFF0000FFFFULL) << 16);
	return (v >> 32) | (v << 32);
#else
	return __brevll(v);
#endif
}

// encode_ise (astcenc_integer_sequence.cpp:493-648); get(i) yields the i-th value to encode
template <typename GetFn>
ASTC_FN void encode_ise(int quant_level, unsigned int character_count, GetFn get, Bits128& out, unsigned int bit_offset) {
	const DevConstTables* ct = ASTC_CT;
	unsigned int bits, trits, quints;
	ise_btq(quant_level, bits, trits, quints);
	unsigned int mask = (1u << bits) - 1;
	if (trits) {
		unsigned int i = 0;
		while (i < character_count) {
			unsigned int v0 = get(i);
			unsigned int v1 = (i + 1 < character_count) ? get(i + 1) : 0u;
			unsigned int v2 = (i + 2 < character_count) ? get(i + 2) : 0u;
			unsigned int v3 = (i + 3 < character_count) ? get(i + 3) : 0u;
			unsigned int v4 = (i + 4 < character_count) ? get(i + 4) : 0u;
			unsigned int T = ct->integer_of_trits[(((((v4 >> bits) * 3 + (v3 >> bits)) * 3 + (v2 >> bits)) * 3 + (v1 >> bits)) * 3) + (v0 >> bits)];
			put_bits(out, (v0 & mask) | (((T >> 0) & 3) << bits), bits + 2, bit_offset);
			bit_offset += bits + 2;
			if (++i >= character_count) break;
			put_bits(out, (v1 & mask) | (((T >> 2) & 3) << bits), bits + 2, bit_offset);
			bit_offset += bits + 2;
			if (++i >= character_count) break;
			put_bits(out, (v2 & mask) | (((T >> 4) & 1) << bits), bits + 1, bit_offset);
			bit_offset += bits + 1;
			if (++i >= character_count) break;
			put_bits(out, (v3 & mask) | (((T >> 5) & 3) << bits), bits + 2, bit_offset);
			bit_offset += bits + 2;
			if (++i >= character_count) break;
			put_bits(out, (v4 & mask) | (((T >> 7) & 1) << bits), bits + 1, bit_offset);
			bit_offset += bits + 1;
			++i;
		}
	} else if (quints) {
		unsigned int i = 0;
		while (i < character_count) {
			unsigned int v0 = get(i);
			unsigned int v1 = (i + 1 < character_count) ? get(i + 1) : 0u;
			unsigned int v2 = (i + 2 < character_count) ? get(i + 2) : 0u;
			unsigned int Q = ct->integer_of_quints[((v2 >> bits) * 5 + (v1 >> bits)) * 5 + (v0 >> bits)];
			put_bits(out, (v0 & mask) | (((Q >> 0) & 7) << bits), bits + 3, bit_offset);
			bit_offset += bits + 3;
			if (++i >= character_count) break;
			put_bits(out, (v1 & mask) | (((Q >> 3) & 3) << bits), bits + 2, bit_offset);
			bit_offset += bits + 2;
			if (++i >= character_count) break;
			put_bits(out, (v2 & mask) | (((Q >> 5) & 3) << bits), bits + 2, bit_offset);
			bit_offset += bits + 2;
			++i;
		}
	} else {
		for (unsigned int i = 0; i < character_count; i++) {
			put_bits(out, get(i), bits, bit_offset);
			bit_offset += bits;
		}
	}
}

// Writes the 16 physical bytes of the best block (header hdr, arrays in w.best_*) to out. Lane 0 only.
ASTC_NOINLINE void symbolic_to_physical(const WCtx& w, const ScbHdr& scb, uint8_t* out) {
	Bits128 pcb;
	pcb.lo = 0;
	pcb.hi = 0;
	if (scb.block_type == SYM_BTYPE_CONST_U16 || scb.block_type == SYM_BTYPE_CONST_F16) {
		// FC FD FF .. FF (UNORM16) or FC FF FF .. FF (FP16), then the four 16-bit components
		pcb.lo = scb.block_type == SYM_BTYPE_CONST_U16 ? 0xFFFFFFFFFFFFFDFCULL : 0xFFFFFFFFFFFFFFFCULL;
		pcb.hi = ((uint64_t)(scb.constant_color[0] & 0xFFFF)) | ((uint64_t)(scb.constant_color[1] & 0xFFFF) << 16) |
		         ((uint64_t)(scb.constant_color[2] & 0xFFFF) << 32) | ((uint64_t)(scb.constant_color[3] & 0xFFFF) << 48);
	} else {
		const DevBsd& bsd = *w.bsd;
		const DevConstTables* ct = ASTC_CT;
		unsigned int partition_count = scb.partition_count;
		const DevBlockMode bm = bsd.block_modes[bsd.block_mode_packed_index[scb.block_mode]];
		int weight_count = bsd.dec_modes[bm.decimation_mode].weight_count;
		int weight_quant_method = bm.quant_mode;
		float weight_quant_levels = static_cast<float>(quant_level_count(weight_quant_method));
		int is_dual_plane = bm.is_dual_plane;
		const uint8_t* scramble = ct->wq_scramble_map[weight_quant_method];
		int real_weight_count = is_dual_plane ? 2 * weight_count : weight_count;
		int bits_for_weights = (int)ise_sequence_bitcount((unsigned int)real_weight_count, weight_quant_method);
		const uint8_t* bw = w.best_weights;
		// the i-th weight in transmission order: planes interleave for dual-plane modes (:127-150)
		auto get_weight = [&](unsigned int i) -> unsigned int {
			unsigned int src = is_dual_plane ? ((i >> 1) + ((i & 1) ? 32u : 0u)) : i;
			float uqw = static_cast<float>(bw[src]);
			float qw = (uqw / 64.0f) * (weight_quant_levels - 1.0f);
			int qwi = static_cast<int>(qw + 0.5f);
			return scramble[qwi];
		};
		Bits128 wb;
		wb.lo = 0;
		wb.hi = 0;
		encode_ise(weight_quant_method, (unsigned int)real_weight_count, get_weight, wb, 0);
		// the weight stream is stored bit-reversed from the top of the block (:153-156)
		pcb.lo = brev64(wb.hi);
		pcb.hi = brev64(wb.lo);
		put_bits(pcb, scb.block_mode, 11, 0);
		put_bits(pcb, partition_count - 1, 2, 11);
		int below_weights_pos = 128 - bits_for_weights;
		if (partition_count > 1) {
			put_bits(pcb, scb.partition_index, 6, 13);
			put_bits(pcb, scb.partition_index >> 6, 10 - 6, 19);
			if (scb.color_formats_matched) {
				put_bits(pcb, (unsigned int)scb.color_formats[0] << 2, 6, 13 + 10);
			} else {
				int low_class = 4;
				for (unsigned int i = 0; i < partition_count; i++) {
					int class_of_format = scb.color_formats[i] >> 2;
					low_class = mini(class_of_format, low_class);
				}
				if (low_class == 3) {
					low_class = 2;
				}
				int encoded_type = low_class + 1;
				int bitpos = 2;
				for (unsigned int i = 0; i < partition_count; i++) {
					int classbit_of_format = (scb.color_formats[i] >> 2) - low_class;
					encoded_type |= classbit_of_format << bitpos;
					bitpos++;
				}
				for (unsigned int i = 0; i < partition_count; i++) {
					int lowbits_of_format = scb.color_formats[i] & 3;
					encoded_type |= lowbits_of_format << bitpos;
					bitpos += 2;
				}
				int encoded_type_lowpart = encoded_type & 0x3F;
				int encoded_type_highpart = encoded_type >> 6;
				int encoded_type_highpart_size = (3 * (int)partition_count) - 4;
				int encoded_type_highpart_pos = 128 - bits_for_weights - encoded_type_highpart_size;
				put_bits(pcb, (unsigned int)encoded_type_lowpart, 6, 13 + 10);
				put_bits(pcb, (unsigned int)encoded_type_highpart, (unsigned int)encoded_type_highpart_size, (unsigned int)encoded_type_highpart_pos);
				below_weights_pos -= encoded_type_highpart_size;
			}
		} else {
			put_bits(pcb, scb.color_formats[0], 4, 13);
		}
		if (is_dual_plane) {
			put_bits(pcb, (unsigned int)scb.plane2_component, 2, (unsigned int)(below_weights_pos - 2));
		}
		// colour values: partition after partition, 2 * (class + 1) values each (:268-285)
		const uint8_t* pack_table = ct->color_uquant_to_scrambled_pquant[scb.quant_mode - QUANT_6];
		unsigned int n0 = 2u * (scb.color_formats[0] >> 2) + 2u;
		unsigned int n1 = partition_count > 1 ? 2u * (scb.color_formats[1] >> 2) + 2u : 0u;
		unsigned int n2 = partition_count > 2 ? 2u * (scb.color_formats[2] >> 2) + 2u : 0u;
		unsigned int n3 = partition_count > 3 ? 2u * (scb.color_formats[3] >> 2) + 2u : 0u;
		const uint8_t* bc = w.best_colors;
		auto get_color = [&](unsigned int i) -> unsigned int {
			unsigned int p = 0;
			if (i >= n0) { i -= n0; p = 1; if (i >= n1) { i -= n1; p = 2; if (i >= n2) { i -= n2; p = 3; } } }
			return pack_table[bc[p * 8 + i]];
		};
		encode_ise(scb.quant_mode, n0 + n1 + n2 + n3, get_color, pcb, scb.partition_count == 1 ? 17 : 19 + 10);
	}
	for (int i = 0; i < 8; i++) {
		out[i] = (uint8_t)(pcb.lo >> (8 * i));
		out[8 + i] = (uint8_t)(pcb.hi >> (8 * i));
	}
}

#include "astc_dev_partition.cuh"
#include "astc_dev_driver.cuh"
