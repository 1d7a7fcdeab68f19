// ORACLE: second half of astc_codec.cpp (endpoint format choice, refinement, scoring, partition search,
// physical packing, block driver). Included inside namespace ao.

// =============================================================================================
// Endpoint format choice (astcenc_pick_best_endpoint_format.cpp)
// =============================================================================================
struct EncodingChoiceErrors {
	float rgb_scale_error, rgb_luma_error, luminance_error, alpha_drop_error;
	bool can_offset_encode, can_blue_contract;
};

// compute_error_squared_rgb_single_partition :72-219
static void compute_error_squared_rgb_single_partition(const PartitionInfo& pi, int partition_index, const ImageBlock& blk,
                                                       const ProcessedLine& uncor, float& uncor_err, const ProcessedLine& samec, float& samec_err,
                                                       const ProcessedLine& rgbl, float& rgbl_err, const ProcessedLine& lum, float& l_err, float& a_drop_err) {
	f4 ews = blk.channel_weight;
	unsigned int texel_count = pi.partition_texel_count[partition_index];
	const uint8_t* texel_indexes = pi.texels_of_partition[partition_index];
	float default_a = default_alpha(blk);
	acc4 a_drop, u, s, r, l;
	acc_init(a_drop);
	acc_init(u);
	acc_init(s);
	acc_init(r);
	acc_init(l);
	for (unsigned int i = 0; i < texel_count; i++) {
		unsigned int tix = texel_indexes[i];
		float da = blk.data_a[tix];
		float alpha_diff = da - default_a;
		alpha_diff = alpha_diff * alpha_diff;
		acc_add(a_drop, alpha_diff);
		float dr = blk.data_r[tix], dg = blk.data_g[tix], db = blk.data_b[tix];

		float param = dr * uncor.bs.x + dg * uncor.bs.y + db * uncor.bs.z;
		float dist0 = (uncor.amod.x + param * uncor.bs.x) - dr;
		float dist1 = (uncor.amod.y + param * uncor.bs.y) - dg;
		float dist2 = (uncor.amod.z + param * uncor.bs.z) - db;
		float error = dist0 * dist0 * ews.x + dist1 * dist1 * ews.y + dist2 * dist2 * ews.z;
		acc_add(u, error);

		param = dr * samec.bs.x + dg * samec.bs.y + db * samec.bs.z;
		dist0 = (param * samec.bs.x) - dr;
		dist1 = (param * samec.bs.y) - dg;
		dist2 = (param * samec.bs.z) - db;
		error = dist0 * dist0 * ews.x + dist1 * dist1 * ews.y + dist2 * dist2 * ews.z;
		acc_add(s, error);

		param = dr * rgbl.bs.x + dg * rgbl.bs.y + db * rgbl.bs.z;
		dist0 = (rgbl.amod.x + param * rgbl.bs.x) - dr;
		dist1 = (rgbl.amod.y + param * rgbl.bs.y) - dg;
		dist2 = (rgbl.amod.z + param * rgbl.bs.z) - db;
		error = dist0 * dist0 * ews.x + dist1 * dist1 * ews.y + dist2 * dist2 * ews.z;
		acc_add(r, error);

		param = dr * lum.bs.x + dg * lum.bs.y + db * lum.bs.z;
		dist0 = (param * lum.bs.x) - dr;
		dist1 = (param * lum.bs.y) - dg;
		dist2 = (param * lum.bs.z) - db;
		error = dist0 * dist0 * ews.x + dist1 * dist1 * ews.y + dist2 * dist2 * ews.z;
		acc_add(l, error);
	}
	a_drop_err = acc_sum(a_drop) * ews.w;
	uncor_err = acc_sum(u);
	samec_err = acc_sum(s);
	rgbl_err = acc_sum(r);
	l_err = acc_sum(l);
}

static inline f4 dot3_splat(f4 a, f4 b) {   // dot3(): (d,d,d,0)
	float d = dot3_s(a, b);
	return mk4(d, d, d, 0.0f);
}

// compute_encoding_choice_errors :222-312
static void compute_encoding_choice_errors(const ImageBlock& blk, const PartitionInfo& pi, const Endpoints& ep, EncodingChoiceErrors eci[4]) {
	int pc = pi.partition_count;
	PartitionMetrics pms[4];
	compute_avgs_and_dirs_3_comp_rgb(pi, blk, pms);
	for (int i = 0; i < pc; i++) {
		PartitionMetrics& pm = pms[i];
		f4 uncor_a = pm.avg;
		f4 uncor_b = normalize_safe4(pm.dir, unit3());
		f4 samec_b = normalize_safe4(pm.avg, unit3());
		f4 luma_a = pm.avg;
		f4 luma_b = unit3();
		ProcessedLine uncor, samec, rgbl, lum;
		uncor.amod = uncor_a - uncor_b * dot3_splat(uncor_a, uncor_b);
		uncor.bs = uncor_b;
		samec.amod = splat4(0.0f);
		samec.bs = samec_b;
		rgbl.amod = luma_a - luma_b * dot3_splat(luma_a, luma_b);
		rgbl.bs = luma_b;
		lum.amod = splat4(0.0f);
		lum.bs = unit3();
		float uncorr_rgb_error, samechroma_rgb_error, rgb_luma_error, luminance_rgb_error, alpha_drop_error;
		compute_error_squared_rgb_single_partition(pi, i, blk, uncor, uncorr_rgb_error, samec, samechroma_rgb_error,
		                                           rgbl, rgb_luma_error, lum, luminance_rgb_error, alpha_drop_error);
		f4 endpt0 = ep.endpt0[i];
		f4 endpt1 = ep.endpt1[i];
		f4 d = endpt1 - endpt0;
		const float lim = 0.12f * 65535.0f;
		bool can_offset_encode = (absf(d.x) < lim) && (absf(d.y) < lim) && (absf(d.z) < lim);
		eci[i].rgb_scale_error = (samechroma_rgb_error - uncorr_rgb_error) * 0.7f;
		eci[i].rgb_luma_error = (rgb_luma_error - uncorr_rgb_error) * 1.5f;
		eci[i].luminance_error = (luminance_rgb_error - uncorr_rgb_error) * 3.0f;
		eci[i].alpha_drop_error = alpha_drop_error * 3.0f;
		eci[i].can_offset_encode = can_offset_encode;
		eci[i].can_blue_contract = !is_luminance(blk);
	}
}

// compute_color_error_for_every_integer_count_and_quant_level :315-675
static void compute_color_error_for_every_integer_count_and_quant_level(bool encode_hdr_rgb, bool encode_hdr_alpha, int partition_index,
                                                                        const PartitionInfo& pi, const EncodingChoiceErrors& eci, const Endpoints& ep,
                                                                        f4 error_weight, float best_error[21][4], uint8_t format_of_choice[21][4]) {
	int partition_size = pi.partition_texel_count[partition_index];
	static const float den[17] = {5 * 5, 7 * 7, 9 * 9, 11 * 11, 15 * 15, 19 * 19, 23 * 23, 31 * 31, 39 * 39, 47 * 47,
	                              63 * 63, 79 * 79, 95 * 95, 127 * 127, 159 * 159, 191 * 191, 255 * 255};
	float baseline_quant_error[17];
	for (int i = 0; i < 17; i++) {
		baseline_quant_error[i] = (65536.0f * 65536.0f / 18.0f) / den[i];
	}
	f4 ep0 = ep.endpt0[partition_index];
	f4 ep1 = ep.endpt1[partition_index];
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
		static const float rgbo_error_scales[6] = {4.0f, 4.0f, 16.0f, 64.0f, 256.0f, 1024.0f};
		static const float rgb_error_scales[9] = {64.0f, 64.0f, 16.0f, 16.0f, 4.0f, 4.0f, 1.0f, 1.0f, 384.0f};
		float mode7mult = rgbo_error_scales[rgbo_mode] * 0.0015f;
		float mode11mult = rgb_error_scales[rgb_mode] * 0.010f;
		float lum_high = hadd_rgb_s(ep1) * (1.0f / 3.0f);
		float lum_low = hadd_rgb_s(ep0) * (1.0f / 3.0f);
		float lumdif = lum_high - lum_low;
		float mode23mult = lumdif < 960 ? 4.0f : lumdif < 3968 ? 16.0f : 128.0f;
		mode23mult *= 0.0005f;
		for (int i = QUANT_2; i < QUANT_16; i++) {
			best_error[i][3] = ERROR_CALC_DEFAULT;
			best_error[i][2] = ERROR_CALC_DEFAULT;
			best_error[i][1] = ERROR_CALC_DEFAULT;
			best_error[i][0] = ERROR_CALC_DEFAULT;
			format_of_choice[i][3] = static_cast<uint8_t>(encode_hdr_alpha ? FMT_HDR_RGBA : FMT_HDR_RGB_LDR_ALPHA);
			format_of_choice[i][2] = FMT_HDR_RGB;
			format_of_choice[i][1] = FMT_HDR_RGB_SCALE;
			format_of_choice[i][0] = FMT_HDR_LUMINANCE_LARGE_RANGE;
		}
		for (int i = QUANT_16; i <= QUANT_256; i++) {
			float base_quant_error = baseline_quant_error[i - QUANT_6] * static_cast<float>(partition_size);
			float rgb_quantization_error = error_weight_rgbsum * base_quant_error * 2.0f;
			float alpha_quantization_error = error_weight.w * base_quant_error * 2.0f;
			float rgba_quantization_error = rgb_quantization_error + alpha_quantization_error;
			float full_hdr_rgba_error = rgba_quantization_error + rgb_range_error + alpha_range_error;
			best_error[i][3] = full_hdr_rgba_error;
			format_of_choice[i][3] = static_cast<uint8_t>(encode_hdr_alpha ? FMT_HDR_RGBA : FMT_HDR_RGB_LDR_ALPHA);
			float full_hdr_rgb_error = (rgb_quantization_error * mode11mult) + rgb_range_error + eci.alpha_drop_error;
			best_error[i][2] = full_hdr_rgb_error;
			format_of_choice[i][2] = FMT_HDR_RGB;
			float hdr_rgb_scale_error = (rgb_quantization_error * mode7mult) + rgb_range_error + eci.alpha_drop_error + eci.rgb_luma_error;
			best_error[i][1] = hdr_rgb_scale_error;
			format_of_choice[i][1] = FMT_HDR_RGB_SCALE;
			float hdr_luminance_error = (rgb_quantization_error * mode23mult) + rgb_range_error + eci.alpha_drop_error + eci.luminance_error;
			best_error[i][0] = hdr_luminance_error;
			format_of_choice[i][0] = FMT_HDR_LUMINANCE_LARGE_RANGE;
		}
	} else {
		for (int i = QUANT_2; i < QUANT_6; i++) {
			best_error[i][3] = ERROR_CALC_DEFAULT;
			best_error[i][2] = ERROR_CALC_DEFAULT;
			best_error[i][1] = ERROR_CALC_DEFAULT;
			best_error[i][0] = ERROR_CALC_DEFAULT;
			format_of_choice[i][3] = FMT_RGBA;
			format_of_choice[i][2] = FMT_RGB;
			format_of_choice[i][1] = FMT_RGB_SCALE;
			format_of_choice[i][0] = FMT_LUMINANCE;
		}
		float base_quant_error_rgb = error_weight_rgbsum * static_cast<float>(partition_size);
		float base_quant_error_a = error_weight.w * static_cast<float>(partition_size);
		float base_quant_error_rgba = base_quant_error_rgb + base_quant_error_a;
		float error_scale_bc_rgba = eci.can_blue_contract ? 0.625f : 1.0f;
		float error_scale_oe_rgba = eci.can_offset_encode ? 0.5f : 1.0f;
		float error_scale_bc_rgb = eci.can_blue_contract ? 0.5f : 1.0f;
		float error_scale_oe_rgb = eci.can_offset_encode ? 0.25f : 1.0f;
		for (int i = QUANT_6; i <= QUANT_256; i++) {
			if (i >= QUANT_192) {
				error_scale_oe_rgba = 1.0f;
				error_scale_oe_rgb = 1.0f;
			}
			float base_quant_error = baseline_quant_error[i - QUANT_6];
			float quant_error_rgb = base_quant_error_rgb * base_quant_error;
			float quant_error_rgba = base_quant_error_rgba * base_quant_error;
			float full_ldr_rgba_error = quant_error_rgba * error_scale_bc_rgba * error_scale_oe_rgba + rgb_range_error + alpha_range_error;
			best_error[i][3] = full_ldr_rgba_error;
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
			float luminance_error = quant_error_rgb + rgb_range_error + eci.alpha_drop_error + eci.luminance_error;
			best_error[i][0] = luminance_error;
			format_of_choice[i][0] = FMT_LUMINANCE;
		}
	}
}

// one_partition_find_best_combination_for_bitcount :678-725
static float one_partition_find_best_combination_for_bitcount(const float best_combined_error[21][4], const uint8_t best_combined_format[21][4],
                                                              int bits_available, uint8_t& best_quant_level, uint8_t& best_format) {
	const ConstTables& ct = const_tables();
	int best_integer_count = 0;
	float best_integer_count_error = ERROR_CALC_DEFAULT;
	for (int integer_count = 1; integer_count <= 4; integer_count++) {
		int quant_level = ct.quant_mode_table[integer_count][bits_available];
		if (quant_level < QUANT_6) {
			continue;
		}
		float integer_count_error = best_combined_error[quant_level][integer_count - 1];
		if (integer_count_error < best_integer_count_error) {
			best_integer_count_error = integer_count_error;
			best_integer_count = integer_count - 1;
		}
	}
	int ql = ct.quant_mode_table[best_integer_count + 1][bits_available];
	best_quant_level = static_cast<uint8_t>(ql);
	best_format = FMT_LUMINANCE;
	if (ql >= QUANT_6) {
		best_format = best_combined_format[ql][best_integer_count];
	}
	return best_integer_count_error;
}

// N-partition combination tables (:728-1093). combined[quant][intcnt] with intcnt = sum of per-partition
// integer-count indices; partitions' counts may differ by at most one.
static void multi_partition_find_best_combination(int pc, const float best_error[4][21][4], const uint8_t best_format[4][21][4],
                                                  float combined_error[21][13], uint8_t combined_format[21][13][4]) {
	int width = pc == 2 ? 7 : pc == 3 ? 10 : 13;
	for (int i = QUANT_2; i <= QUANT_256; i++) {
		for (int j = 0; j < width; j++) {
			combined_error[i][j] = ERROR_CALC_DEFAULT;
		}
	}
	for (int quant = QUANT_6; quant <= QUANT_256; quant++) {
		for (int i = 0; i < 4; i++) {
			for (int j = 0; j < 4; j++) {
				int low2 = mini(i, j);
				int high2 = maxi(i, j);
				if ((high2 - low2) > 1) {
					continue;
				}
				if (pc == 2) {
					int intcnt = i + j;
					float errorterm = minf(best_error[0][quant][i] + best_error[1][quant][j], 1e10f);
					if (errorterm <= combined_error[quant][intcnt]) {
						combined_error[quant][intcnt] = errorterm;
						combined_format[quant][intcnt][0] = best_format[0][quant][i];
						combined_format[quant][intcnt][1] = best_format[1][quant][j];
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
						float errorterm = minf(best_error[0][quant][i] + best_error[1][quant][j] + best_error[2][quant][k], 1e10f);
						if (errorterm <= combined_error[quant][intcnt]) {
							combined_error[quant][intcnt] = errorterm;
							combined_format[quant][intcnt][0] = best_format[0][quant][i];
							combined_format[quant][intcnt][1] = best_format[1][quant][j];
							combined_format[quant][intcnt][2] = best_format[2][quant][k];
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
						float errorterm = minf(best_error[0][quant][i] + best_error[1][quant][j] + best_error[2][quant][k] + best_error[3][quant][l], 1e10f);
						if (errorterm <= combined_error[quant][intcnt]) {
							combined_error[quant][intcnt] = errorterm;
							combined_format[quant][intcnt][0] = best_format[0][quant][i];
							combined_format[quant][intcnt][1] = best_format[1][quant][j];
							combined_format[quant][intcnt][2] = best_format[2][quant][k];
							combined_format[quant][intcnt][3] = best_format[3][quant][l];
						}
					}
				}
			}
		}
	}
}

static float multi_partition_find_best_combination_for_bitcount(int pc, const float combined_error[21][13], const uint8_t combined_format[21][13][4],
                                                                int bits_available, uint8_t& best_quant_level, uint8_t& best_quant_level_mod, uint8_t* best_formats) {
	const ConstTables& ct = const_tables();
	int best_integer_count = 0;
	float best_integer_count_error = ERROR_CALC_DEFAULT;
	int first = pc;                       // 2, 3, 4
	int last = pc == 2 ? 8 : 9;           // :783, :908, :1044
	int mod_bits = pc == 2 ? 2 : pc == 3 ? 5 : 8;
	for (int integer_count = first; integer_count <= last; integer_count++) {
		int quant_level = ct.quant_mode_table[integer_count][bits_available];
		if (quant_level < QUANT_6) {
			break;
		}
		float integer_count_error = combined_error[quant_level][integer_count - first];
		if (integer_count_error < best_integer_count_error) {
			best_integer_count_error = integer_count_error;
			best_integer_count = integer_count;
		}
	}
	int ql = ct.quant_mode_table[best_integer_count][bits_available];
	int ql_mod = ct.quant_mode_table[best_integer_count][bits_available + mod_bits];
	best_quant_level = static_cast<uint8_t>(ql);
	best_quant_level_mod = static_cast<uint8_t>(ql_mod);
	if (ql >= QUANT_6) {
		for (int i = 0; i < pc; i++) {
			best_formats[i] = combined_format[ql][best_integer_count - first][i];
		}
	} else {
		for (int i = 0; i < pc; i++) {
			best_formats[i] = FMT_LUMINANCE;
		}
	}
	return best_integer_count_error;
}

// compute_ideal_endpoint_formats :1096-1357
static unsigned int compute_ideal_endpoint_formats(const PartitionInfo& pi, const ImageBlock& blk, const Endpoints& ep, const int8_t* qwt_bitcounts,
                                                   const float* qwt_errors, unsigned int tune_candidate_limit, unsigned int start_block_mode,
                                                   unsigned int end_block_mode, uint8_t partition_format_specifiers[8][4], int block_mode[8],
                                                   uint8_t quant_level[8], uint8_t quant_level_mod[8], WorkBuf& tmp) {
	int pc = pi.partition_count;
	bool encode_hdr_rgb = blk.rgb_lns0 != 0;
	bool encode_hdr_alpha = blk.alpha_lns0 != 0;
	EncodingChoiceErrors eci[4];
	compute_encoding_choice_errors(blk, pi, ep, eci);
	float best_error[4][21][4];
	uint8_t format_of_choice[4][21][4];
	for (int i = 0; i < pc; i++) {
		compute_color_error_for_every_integer_count_and_quant_level(encode_hdr_rgb, encode_hdr_alpha, i, pi, eci[i], ep, blk.channel_weight,
		                                                            best_error[i], format_of_choice[i]);
	}
	float* errors_of_best_combination = tmp.errors_of_best_combination;
	uint8_t* best_quant_levels = tmp.best_quant_levels;
	uint8_t* best_quant_levels_mod = tmp.best_quant_levels_mod;

	float error_of_best_combination = ERROR_CALC_DEFAULT;
	int index_of_best_combination = -1;
	float combined_error[21][13];
	uint8_t combined_format[21][13][4];
	if (pc >= 2) {
		multi_partition_find_best_combination(pc, best_error, format_of_choice, combined_error, combined_format);
	}
	for (unsigned int i = start_block_mode; i < end_block_mode; i++) {
		if (qwt_errors[i] >= ERROR_CALC_DEFAULT) {
			errors_of_best_combination[i] = ERROR_CALC_DEFAULT;
			continue;
		}
		float error_of_best;
		if (pc == 1) {
			error_of_best = one_partition_find_best_combination_for_bitcount(best_error[0], format_of_choice[0], qwt_bitcounts[i],
			                                                                 best_quant_levels[i], tmp.best_ep_formats[i][0]);
			best_quant_levels_mod[i] = best_quant_levels[i];
		} else {
			error_of_best = multi_partition_find_best_combination_for_bitcount(pc, combined_error, combined_format, qwt_bitcounts[i],
			                                                                   best_quant_levels[i], best_quant_levels_mod[i], tmp.best_ep_formats[i]);
		}
		float total_error = error_of_best + qwt_errors[i];
		errors_of_best_combination[i] = total_error;
		if (total_error < error_of_best_combination) {
			error_of_best_combination = total_error;
			index_of_best_combination = (int)i;
		}
	}
	int best_error_weights[8];
	best_error_weights[0] = index_of_best_combination;
	if (index_of_best_combination >= 0) {
		errors_of_best_combination[index_of_best_combination] = ERROR_CALC_DEFAULT;
	}
	for (unsigned int i = 1; i < tune_candidate_limit; i++) {
		// lowest index among the entries holding the minimum error below 1e30 (:1286-1333)
		int best_error_index = -1;
		float best_ep_error = ERROR_CALC_DEFAULT;
		for (unsigned int j = start_block_mode; j < end_block_mode; j++) {
			float err = errors_of_best_combination[j];
			if (err < best_ep_error) {
				best_ep_error = err;
				best_error_index = (int)j;
			}
		}
		best_error_weights[i] = best_error_index;
		if (best_error_index >= 0) {
			errors_of_best_combination[best_error_index] = ERROR_CALC_DEFAULT;
		} else {
			break;
		}
	}
	for (unsigned int i = 0; i < tune_candidate_limit; i++) {
		if (best_error_weights[i] < 0) {
			return i;
		}
		block_mode[i] = best_error_weights[i];
		quant_level[i] = best_quant_levels[best_error_weights[i]];
		quant_level_mod[i] = best_quant_levels_mod[best_error_weights[i]];
		for (int j = 0; j < pc; j++) {
			partition_format_specifiers[i][j] = tmp.best_ep_formats[best_error_weights[i]][j];
		}
	}
	return tune_candidate_limit;
}

// =============================================================================================
// Least-squares endpoint refit (astcenc_ideal_endpoints_and_weights.cpp:1099-1650)
// =============================================================================================
static inline f4 compute_rgbo_vector(f4 rgba_weight_sum, f4 weight_weight_sum, f4 rgbq_sum, float psum) {   // :1099-1143
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

static inline f4 sel4(f4 a, f4 b, bool m0, bool m1, bool m2, bool m3) {
	return mk4(m0 ? b.x : a.x, m1 ? b.y : a.y, m2 ? b.z : a.z, m3 ? b.w : a.w);
}

static void undecimate_weights(const DecimationInfo& di, const uint8_t* uquant, unsigned int texel_count, float* out) {
	float dec_weight[MAX_WEIGHTS];
	for (unsigned int i = 0; i < di.weight_count; i++) {
		dec_weight[i] = static_cast<float>(uquant[i]) * (1.0f / 64.0f);
	}
	for (unsigned int i = 0; i < texel_count; i++) {
		if (di.max_texel_weight_count == 1) out[i] = dec_weight[i];
		else if (di.max_texel_weight_count <= 2) out[i] = bilinear_infill_2(di, dec_weight, i);
		else out[i] = bilinear_infill(di, dec_weight, i);
	}
}

static void rgbo_fallback(const ImageBlock& blk, f4& rgbo, const f4& v0, const f4& v1) {
	(void)blk;
	if (dot_s(rgbo, rgbo) != dot_s(rgbo, rgbo)) {
		float avgdif = hadd_rgb_s(v1 - v0) * (1.0f / 3.0f);
		avgdif = maxf(avgdif, 0.0f);
		f4 avg = (v0 + v1) * 0.5f;
		f4 ep0 = avg - splat4(avgdif) * 0.5f;
		rgbo = mk4(ep0.x, ep0.y, ep0.z, avgdif);
	}
}

// recompute_ideal_colors_1plane :1146-1366
static void recompute_ideal_colors_1plane(const ImageBlock& blk, const PartitionInfo& pi, const DecimationInfo& di, const uint8_t* dec_weights_uquant,
                                          Endpoints& ep, f4 rgbs_vectors[4], f4 rgbo_vectors[4]) {
	unsigned int total_texel_count = blk.texel_count;
	unsigned int pc = pi.partition_count;
	float undec_weight[MAX_TEXELS];
	undecimate_weights(di, dec_weights_uquant, total_texel_count, undec_weight);

	f4 rgba_sum = blk.data_mean * static_cast<float>(blk.texel_count);
	for (unsigned int i = 0; i < pc; i++) {
		unsigned int texel_count = pi.partition_texel_count[i];
		const uint8_t* texel_indexes = pi.texels_of_partition[i];
		if (pc > 1) {
			rgba_sum = splat4(0.0f);
			for (unsigned int j = 0; j < texel_count; j++) {
				rgba_sum = rgba_sum + texel4(blk, texel_indexes[j]);
			}
		}
		rgba_sum = rgba_sum * blk.channel_weight;
		f4 rgba_weight_sum = max4(blk.channel_weight * static_cast<float>(texel_count), splat4(1e-17f));
		f4 q = rgba_sum / rgba_weight_sum;
		f4 scale_dir = normalize4(mk4(q.x, q.y, q.z, 0.0f));
		float scale_max = 0.0f, scale_min = 1e10f;
		float wmin1 = 1.0f, wmax1 = 0.0f;
		float left_sum_s = 0.0f, middle_sum_s = 0.0f, right_sum_s = 0.0f;
		f4 color_vec_x = splat4(0.0f), color_vec_y = splat4(0.0f), scale_vec = splat4(0.0f);
		float weight_weight_sum_s = 1e-17f;
		f4 color_weight = blk.channel_weight;
		float ls_weight = hadd_rgb_s(color_weight);
		for (unsigned int j = 0; j < texel_count; j++) {
			unsigned int tix = texel_indexes[j];
			f4 rgba = texel4(blk, tix);
			float idx0 = undec_weight[tix];
			float om_idx0 = 1.0f - idx0;
			wmin1 = minf(idx0, wmin1);
			wmax1 = maxf(idx0, wmax1);
			float scale = dot3_s(scale_dir, rgba);
			scale_min = minf(scale, scale_min);
			scale_max = maxf(scale, scale_max);
			left_sum_s += om_idx0 * om_idx0;
			middle_sum_s += om_idx0 * idx0;
			right_sum_s += idx0 * idx0;
			weight_weight_sum_s += idx0;
			f4 cwprod = rgba;
			f4 cwiprod = cwprod * idx0;
			color_vec_y = color_vec_y + cwiprod;
			color_vec_x = color_vec_x + (cwprod - cwiprod);
			scale_vec = scale_vec + mk4(om_idx0, idx0, 0.0f, 0.0f) * (scale * ls_weight);
		}
		f4 left_sum = splat4(left_sum_s) * color_weight;
		f4 middle_sum = splat4(middle_sum_s) * color_weight;
		f4 right_sum = splat4(right_sum_s) * color_weight;
		f4 lmrs_sum = mk4(left_sum_s, middle_sum_s, right_sum_s, 0.0f) * ls_weight;
		color_vec_x = color_vec_x * color_weight;
		color_vec_y = color_vec_y * color_weight;
		float scalediv = scale_min / maxf(scale_max, 1e-10f);
		scalediv = clamp1f(scalediv);
		f4 sds = scale_dir * scale_max;
		rgbs_vectors[i] = mk4(sds.x, sds.y, sds.z, scalediv);
		if (wmin1 >= wmax1 * 0.999f) {
			f4 avg = (color_vec_x + color_vec_y) / rgba_weight_sum;
			ep.endpt0[i] = sel4(ep.endpt0[i], avg, avg.x == avg.x, avg.y == avg.y, avg.z == avg.z, avg.w == avg.w);
			ep.endpt1[i] = sel4(ep.endpt1[i], avg, avg.x == avg.x, avg.y == avg.y, avg.z == avg.z, avg.w == avg.w);
			rgbs_vectors[i] = mk4(sds.x, sds.y, sds.z, 1.0f);
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
			ep.endpt0[i] = sel4(ep.endpt0[i], ep0, m[0], m[1], m[2], m[3]);
			ep.endpt1[i] = sel4(ep.endpt1[i], ep1, m[0], m[1], m[2], m[3]);
			float scale_ep0 = (lmrs_sum.z * scale_vec.x - lmrs_sum.y * scale_vec.y) * ls_rdet1;
			float scale_ep1 = (lmrs_sum.x * scale_vec.y - lmrs_sum.y * scale_vec.x) * ls_rdet1;
			if (fabsf(ls_det1) > (ls_mss1 * 1e-4f) && scale_ep0 == scale_ep0 && scale_ep1 == scale_ep1 && scale_ep0 < scale_ep1) {
				float scalediv2 = scale_ep0 / scale_ep1;
				f4 sdsm = scale_dir * scale_ep1;
				rgbs_vectors[i] = mk4(sdsm.x, sdsm.y, sdsm.z, scalediv2);
			}
		}
		if (blk.rgb_lns0 || blk.alpha_lns0) {
			f4 weight_weight_sum = splat4(weight_weight_sum_s) * color_weight;
			float psum = right_sum_s * hadd_rgb_s(color_weight);
			f4 rgbq_sum = color_vec_x + color_vec_y;
			rgbq_sum.w = hadd_rgb_s(color_vec_y);
			f4 rgbovec = compute_rgbo_vector(rgba_weight_sum, weight_weight_sum, rgbq_sum, psum);
			rgbo_vectors[i] = rgbovec;
			rgbo_fallback(blk, rgbo_vectors[i], ep.endpt0[i], ep.endpt1[i]);
		}
	}
}

// recompute_ideal_colors_2planes :1369-1650
static void recompute_ideal_colors_2planes(const ImageBlock& blk, const BlockSizeTables& bsd, const DecimationInfo& di, const uint8_t* uq1, const uint8_t* uq2,
                                           Endpoints& ep, f4& rgbs_vector, f4& rgbo_vector, int plane2_component) {
	unsigned int total_texel_count = blk.texel_count;
	float undec1[MAX_TEXELS], undec2[MAX_TEXELS];
	undecimate_weights(di, uq1, total_texel_count, undec1);
	undecimate_weights(di, uq2, total_texel_count, undec2);
	unsigned int texel_count = bsd.texel_count;
	f4 rgba_weight_sum = max4(blk.channel_weight * static_cast<float>(texel_count), splat4(1e-17f));
	f4 scale_dir = normalize4(mk4(blk.data_mean.x, blk.data_mean.y, blk.data_mean.z, 0.0f));
	float scale_max = 0.0f, scale_min = 1e10f;
	float wmin1 = 1.0f, wmax1 = 0.0f, wmin2 = 1.0f, wmax2 = 0.0f;
	float left1_sum_s = 0.0f, middle1_sum_s = 0.0f, right1_sum_s = 0.0f;
	float left2_sum_s = 0.0f, middle2_sum_s = 0.0f, right2_sum_s = 0.0f;
	f4 color_vec_x = splat4(0.0f), color_vec_y = splat4(0.0f), scale_vec = splat4(0.0f);
	f4 weight_weight_sum = splat4(1e-17f);
	bool p2[4] = {plane2_component == 0, plane2_component == 1, plane2_component == 2, plane2_component == 3};
	f4 color_weight = blk.channel_weight;
	float ls_weight = hadd_rgb_s(color_weight);
	for (unsigned int j = 0; j < texel_count; j++) {
		f4 rgba = texel4(blk, j);
		float idx0 = undec1[j];
		float om_idx0 = 1.0f - idx0;
		wmin1 = minf(idx0, wmin1);
		wmax1 = maxf(idx0, wmax1);
		float scale = dot3_s(scale_dir, rgba);
		scale_min = minf(scale, scale_min);
		scale_max = maxf(scale, scale_max);
		left1_sum_s += om_idx0 * om_idx0;
		middle1_sum_s += om_idx0 * idx0;
		right1_sum_s += idx0 * idx0;
		float idx1 = undec2[j];
		float om_idx1 = 1.0f - idx1;
		wmin2 = minf(idx1, wmin2);
		wmax2 = maxf(idx1, wmax2);
		left2_sum_s += om_idx1 * om_idx1;
		middle2_sum_s += om_idx1 * idx1;
		right2_sum_s += idx1 * idx1;
		f4 color_idx = mk4(p2[0] ? idx1 : idx0, p2[1] ? idx1 : idx0, p2[2] ? idx1 : idx0, p2[3] ? idx1 : idx0);
		f4 cwprod = rgba;
		f4 cwiprod = cwprod * color_idx;
		color_vec_y = color_vec_y + cwiprod;
		color_vec_x = color_vec_x + (cwprod - cwiprod);
		scale_vec = scale_vec + mk4(om_idx0, idx0, 0.0f, 0.0f) * (ls_weight * scale);
		weight_weight_sum = weight_weight_sum + color_idx;
	}
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
	rgbs_vector = mk4(sds.x, sds.y, sds.z, scalediv);
	if (wmin1 >= wmax1 * 0.999f) {
		f4 avg = (color_vec_x + color_vec_y) / rgba_weight_sum;
		bool m[4];
		for (int c = 0; c < 4; c++) {
			m[c] = !p2[c] && (lane(avg, c) == lane(avg, c));
		}
		ep.endpt0[0] = sel4(ep.endpt0[0], avg, m[0], m[1], m[2], m[3]);
		ep.endpt1[0] = sel4(ep.endpt1[0], avg, m[0], m[1], m[2], m[3]);
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
		for (int c = 0; c < 4; c++) {
			bool det = absf(lane(color_det1, c)) > lane(thr, c);
			bool notnan = (lane(ep0, c) == lane(ep0, c)) && (lane(ep1, c) == lane(ep1, c));
			m[c] = !p2[c] && det && notnan;
		}
		ep.endpt0[0] = sel4(ep.endpt0[0], ep0, m[0], m[1], m[2], m[3]);
		ep.endpt1[0] = sel4(ep.endpt1[0], ep1, m[0], m[1], m[2], m[3]);
		if (fabsf(ls_det1) > (ls_mss1 * 1e-4f) && scale_ep0 == scale_ep0 && scale_ep1 == scale_ep1 && scale_ep0 < scale_ep1) {
			float scalediv2 = scale_ep0 / scale_ep1;
			f4 sdsm = scale_dir * scale_ep1;
			rgbs_vector = mk4(sdsm.x, sdsm.y, sdsm.z, scalediv2);
		}
	}
	if (wmin2 >= wmax2 * 0.999f) {
		f4 avg = (color_vec_x + color_vec_y) / rgba_weight_sum;
		bool m[4];
		for (int c = 0; c < 4; c++) {
			m[c] = p2[c] && (lane(avg, c) == lane(avg, c));
		}
		ep.endpt0[0] = sel4(ep.endpt0[0], avg, m[0], m[1], m[2], m[3]);
		ep.endpt1[0] = sel4(ep.endpt1[0], avg, m[0], m[1], m[2], m[3]);
	} else {
		f4 color_det2 = (left2_sum * right2_sum) - (middle2_sum * middle2_sum);
		f4 color_rdet2 = splat4(1.0f) / color_det2;
		f4 color_mss2 = (left2_sum * left2_sum) + (splat4(2.0f) * middle2_sum * middle2_sum) + (right2_sum * right2_sum);
		f4 ep0 = (right2_sum * color_vec_x - middle2_sum * color_vec_y) * color_rdet2;
		f4 ep1 = (left2_sum * color_vec_y - middle2_sum * color_vec_x) * color_rdet2;
		f4 thr = color_mss2 * 1e-4f;
		bool m[4];
		for (int c = 0; c < 4; c++) {
			bool det = absf(lane(color_det2, c)) > lane(thr, c);
			bool notnan = (lane(ep0, c) == lane(ep0, c)) && (lane(ep1, c) == lane(ep1, c));
			m[c] = p2[c] && det && notnan;
		}
		ep.endpt0[0] = sel4(ep.endpt0[0], ep0, m[0], m[1], m[2], m[3]);
		ep.endpt1[0] = sel4(ep.endpt1[0], ep1, m[0], m[1], m[2], m[3]);
	}
	if (blk.rgb_lns0 || blk.alpha_lns0) {
		weight_weight_sum = weight_weight_sum * color_weight;
		f4 rsel = mk4(p2[0] ? right2_sum.x : right1_sum.x, p2[1] ? right2_sum.y : right1_sum.y, p2[2] ? right2_sum.z : right1_sum.z,
		              p2[3] ? right2_sum.w : right1_sum.w);
		float psum = dot3_s(rsel, color_weight);
		f4 rgbq_sum = color_vec_x + color_vec_y;
		rgbq_sum.w = hadd_rgb_s(color_vec_y);
		rgbo_vector = compute_rgbo_vector(rgba_weight_sum, weight_weight_sum, rgbq_sum, psum);
		rgbo_fallback(blk, rgbo_vector, ep.endpt0[0], ep.endpt1[0]);
	}
}

// =============================================================================================
// Decompress-and-diff scoring (astcenc_decompress_symbolic.cpp:89-618)
// =============================================================================================
static void unpack_weights(const BlockSizeTables& bsd, const SymbolicBlock& scb, const DecimationInfo& di, bool is_dual_plane, int* w1, int* w2) {   // :89-167
	for (unsigned int i = 0; i < bsd.texel_count; i++) {
		int s1 = 8, s2 = 8;
		for (int j = 0; j < 4; j++) {
			int tw = di.texel_weights[j][i];
			int c = di.texel_weight_contribs_int[j][i];
			s1 += scb.weights[tw] * c;
			if (is_dual_plane) {
				s2 += scb.weights[tw + PLANE2_OFFSET] * c;
			}
		}
		w1[i] = s1 >> 4;
		if (is_dual_plane) {
			w2[i] = s2 >> 4;
		}
	}
}

static inline bool u8_mask(int profile, const ImageBlock& blk) {   // get_u8_component_mask astcenc_internal.h:1790
	return blk.decode_unorm8 || profile == PRF_LDR_SRGB;
}

static inline i4 lerp_color_int(bool u8, i4 c0, i4 c1, i4 w) {   // :37-61
	i4 r;
	int* rp = &r.x;
	const int* a = &c0.x;
	const int* b = &c1.x;
	const int* wp = &w.x;
	for (int k = 0; k < 4; k++) {
		int w1 = wp[k];
		int w0 = 64 - w1;
		int color = (a[k] * w0) + (b[k] * w1) + 32;
		color = color >> 6;
		if (u8) {
			color = (color >> 8) * 257;
		}
		rp[k] = color;
	}
	return r;
}

// compute_symbolic_block_difference_2plane :313-404 and _1plane :407-502 share the scalar summa form
static float compute_symbolic_block_difference(const Config& config, const BlockSizeTables& bsd, const SymbolicBlock& scb, const ImageBlock& blk, bool dual) {
	if (scb.block_type == SYM_BTYPE_ERROR) {
		return ERROR_CALC_DEFAULT;
	}
	unsigned int pc = scb.partition_count;
	unsigned int packed_part = pc >= 2 ? bsd.partitioning_packed_index[pc - 2][scb.partition_index] : 0;
	const PartitionInfo& pi = bsd.partitionings[pc][packed_part];
	const BlockMode& bm = bsd.block_modes[bsd.block_mode_packed_index[scb.block_mode]];
	const DecimationInfo& di = bsd.decimation_tables[bm.decimation_mode];
	int w1[MAX_TEXELS], w2[MAX_TEXELS];
	unpack_weights(bsd, scb, di, dual, w1, w2);
	bool u8 = u8_mask(config.profile, blk);
	float summa = 0.0f;
	for (unsigned int p = 0; p < pc; p++) {
		i4 ep0, ep1;
		bool rgb_lns, a_lns;
		unpack_color_endpoints(config.profile, scb.color_formats[p], scb.color_values[p], rgb_lns, a_lns, ep0, ep1);
		unsigned int n = dual ? bsd.texel_count : pi.partition_texel_count[p];
		for (unsigned int j = 0; j < n; j++) {
			unsigned int tix = dual ? j : pi.texels_of_partition[p][j];
			i4 w = mki4(w1[tix], w1[tix], w1[tix], w1[tix]);
			if (dual) {
				if (scb.plane2_component == 0) w.x = w2[tix];
				else if (scb.plane2_component == 1) w.y = w2[tix];
				else if (scb.plane2_component == 2) w.z = w2[tix];
				else if (scb.plane2_component == 3) w.w = w2[tix];
			}
			i4 ci = lerp_color_int(u8, ep0, ep1, w);
			f4 color = mk4((float)ci.x, (float)ci.y, (float)ci.z, (float)ci.w);
			f4 old = texel4(blk, tix);
			if (config.flags & FLG_MAP_RGBM) {
				if (color.w == 0.0f) {
					return -ERROR_CALC_DEFAULT;
				}
				color = mk4(color.x * color.w * config.rgbm_m_scale, color.y * color.w * config.rgbm_m_scale, color.z * color.w * config.rgbm_m_scale, 1.0f);
				old = mk4(old.x * old.w * config.rgbm_m_scale, old.y * old.w * config.rgbm_m_scale, old.z * old.w * config.rgbm_m_scale, 1.0f);
			}
			f4 error = old - color;
			error = min4(mk4(absf(error.x), absf(error.y), absf(error.z), absf(error.w)), splat4(1e15f));
			error = error * error;
			summa += minf(dot_s(error, blk.channel_weight), ERROR_CALC_DEFAULT);
		}
	}
	return summa;
}

// compute_symbolic_block_difference_1plane_1partition :505-618
static float compute_symbolic_block_difference_1plane_1partition(const Config& config, const BlockSizeTables& bsd, const SymbolicBlock& scb, const ImageBlock& blk) {
	if (scb.block_type == SYM_BTYPE_ERROR) {
		return ERROR_CALC_DEFAULT;
	}
	const BlockMode& bm = bsd.block_modes[bsd.block_mode_packed_index[scb.block_mode]];
	const DecimationInfo& di = bsd.decimation_tables[bm.decimation_mode];
	int w1[MAX_TEXELS];
	unpack_weights(bsd, scb, di, false, w1, nullptr);
	i4 ep0, ep1;
	bool rgb_lns, a_lns;
	unpack_color_endpoints(config.profile, scb.color_formats[0], scb.color_values[0], rgb_lns, a_lns, ep0, ep1);
	bool u8 = u8_mask(config.profile, blk);
	acc4 acc;
	acc_init(acc);
	f4 cw = blk.channel_weight;
	for (unsigned int i = 0; i < bsd.texel_count; i++) {
		i4 ci = lerp_color_int(u8, ep0, ep1, mki4(w1[i], w1[i], w1[i], w1[i]));
		float er = minf(absf(blk.data_r[i] - (float)ci.x), 1e15f);
		float eg = minf(absf(blk.data_g[i] - (float)ci.y), 1e15f);
		float eb = minf(absf(blk.data_b[i] - (float)ci.z), 1e15f);
		float ea = minf(absf(blk.data_a[i] - (float)ci.w), 1e15f);
		er = er * er;
		eg = eg * eg;
		eb = eb * eb;
		ea = ea * ea;
		float metric = er * cw.x + eg * cw.y + eb * cw.z + ea * cw.w;
		acc_add(acc, metric);
	}
	return acc_sum(acc);
}

// =============================================================================================
// Weight realignment (astcenc_compress_symbolic.cpp:69-350)
// =============================================================================================
static bool realign_weights(int decode_mode, const BlockSizeTables& bsd, const ImageBlock& blk, SymbolicBlock& scb, bool decimated) {
	unsigned int pc = scb.partition_count;
	unsigned int packed_part = pc >= 2 ? bsd.partitioning_packed_index[pc - 2][scb.partition_index] : 0;
	const PartitionInfo& pi = bsd.partitionings[pc][packed_part];
	const BlockMode& bm = bsd.block_modes[bsd.block_mode_packed_index[scb.block_mode]];
	const WeightQuantTable& qat = const_tables().weight_quant[bm.quant_mode];
	const DecimationInfo& di = bsd.decimation_tables[bm.decimation_mode];
	unsigned int weight_count = di.weight_count;
	unsigned int max_plane = bm.is_dual_plane;
	int plane2_component = scb.plane2_component;
	bool plane_mask[4] = {plane2_component == 0, plane2_component == 1, plane2_component == 2, plane2_component == 3};
	i4 endpnt0[4], endpnt1[4];
	f4 endpnt0f[4], offset[4];
	for (unsigned int p = 0; p < pc; p++) {
		bool rgb_hdr, alpha_hdr;
		unpack_color_endpoints(decode_mode, scb.color_formats[p], scb.color_values[p], rgb_hdr, alpha_hdr, endpnt0[p], endpnt1[p]);
	}
	uint8_t* dec_weights_uquant = scb.weights;
	bool adjustments = false;
	f4 ew = blk.channel_weight;
	for (unsigned int pl = 0; pl <= max_plane; pl++) {
		for (unsigned int p = 0; p < pc; p++) {
			i4 epd = mki4(endpnt1[p].x - endpnt0[p].x, endpnt1[p].y - endpnt0[p].y, endpnt1[p].z - endpnt0[p].z, endpnt1[p].w - endpnt0[p].w);
			if (plane_mask[0]) epd.x = 0;
			if (plane_mask[1]) epd.y = 0;
			if (plane_mask[2]) epd.z = 0;
			if (plane_mask[3]) epd.w = 0;
			endpnt0f[p] = mk4((float)endpnt0[p].x, (float)endpnt0[p].y, (float)endpnt0[p].z, (float)endpnt0[p].w);
			offset[p] = mk4((float)epd.x, (float)epd.y, (float)epd.z, (float)epd.w) * (1.0f / 64.0f);
		}
		if (!decimated) {
			// realign_weights_undecimated :69-185
			for (unsigned int texel = 0; texel < bsd.texel_count; texel++) {
				int uqw = dec_weights_uquant[texel];
				uint32_t prev_and_next = qat.prev_next_values[uqw];
				int uqw_down = prev_and_next & 0xFF;
				int uqw_up = (prev_and_next >> 8) & 0xFF;
				float weight_base = static_cast<float>(uqw);
				float weight_down = static_cast<float>(uqw_down - uqw);
				float weight_up = static_cast<float>(uqw_up - uqw);
				unsigned int partition = pi.partition_of_texel[texel];
				f4 color_offset = offset[partition];
				f4 color_base = endpnt0f[partition];
				f4 color = color_base + color_offset * weight_base;
				f4 orig_color = texel4(blk, texel);
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
		} else {
			// realign_weights_decimated :188-350
			float uq_weightsf[MAX_WEIGHTS];
			for (unsigned int we = 0; we < weight_count; we++) {
				uq_weightsf[we] = static_cast<float>(dec_weights_uquant[we]);
			}
			for (unsigned int we = 0; we < weight_count; we++) {
				int uqw = dec_weights_uquant[we];
				uint32_t prev_and_next = qat.prev_next_values[uqw];
				float uqw_base = uq_weightsf[we];
				float uqw_down = static_cast<float>(prev_and_next & 0xFF);
				float uqw_up = static_cast<float>((prev_and_next >> 8) & 0xFF);
				float uqw_diff_down = uqw_down - uqw_base;
				float uqw_diff_up = uqw_up - uqw_base;
				f4 error_basev = splat4(0.0f), error_downv = splat4(0.0f), error_upv = splat4(0.0f);
				unsigned int off = di.weight_texel_offset[we];
				unsigned int texels_to_evaluate = di.weight_texel_count[we];
				for (unsigned int te = 0; te < texels_to_evaluate; te++) {
					unsigned int texel = di.weight_texels[off + te];
					float tw_base = di.texel_contrib_for_weight[off + te];
					float weight_base = (uq_weightsf[di.texel_weights[0][texel]] * di.texel_weight_contribs_float[0][texel] +
					                     uq_weightsf[di.texel_weights[1][texel]] * di.texel_weight_contribs_float[1][texel]) +
					                    (uq_weightsf[di.texel_weights[2][texel]] * di.texel_weight_contribs_float[2][texel] +
					                     uq_weightsf[di.texel_weights[3][texel]] * di.texel_weight_contribs_float[3][texel]);
					float weight_down = weight_base + uqw_diff_down * tw_base - weight_base;
					float weight_up = weight_base + uqw_diff_up * tw_base - weight_base;
					unsigned int partition = pi.partition_of_texel[texel];
					f4 color_offset = offset[partition];
					f4 color_base = endpnt0f[partition];
					f4 color = color_base + color_offset * weight_base;
					f4 orig_color = texel4(blk, texel);
					f4 color_diff = color - orig_color;
					f4 color_down_diff = color_diff + color_offset * weight_down;
					f4 color_up_diff = color_diff + color_offset * weight_up;
					error_basev = error_basev + color_diff * color_diff;
					error_downv = error_downv + color_down_diff * color_down_diff;
					error_upv = error_upv + color_up_diff * color_up_diff;
				}
				float error_base = hadd_s(error_basev * ew);
				float error_down = hadd_s(error_downv * ew);
				float error_up = hadd_s(error_upv * ew);
				if ((error_up < error_base) && (error_up < error_down) && (uqw < 64)) {
					uq_weightsf[we] = uqw_up;
					dec_weights_uquant[we] = static_cast<uint8_t>(uqw_up);
					adjustments = true;
				} else if ((error_down < error_base) && (uqw > 0)) {
					uq_weightsf[we] = uqw_down;
					dec_weights_uquant[we] = static_cast<uint8_t>(uqw_down);
					adjustments = true;
				}
			}
		}
		dec_weights_uquant += PLANE2_OFFSET;
		for (int c = 0; c < 4; c++) {
			plane_mask[c] = !plane_mask[c];
		}
	}
	return adjustments;
}

// =============================================================================================
// Physical block packing (astcenc_symbolic_physical.cpp:102-286, astcenc_integer_sequence.cpp:493-648)
// =============================================================================================
static inline void write_bits(unsigned int value, unsigned int bitcount, unsigned int bitoffset, uint8_t* ptr) {
	unsigned int mask = (1u << bitcount) - 1;
	value &= mask;
	ptr += bitoffset >> 3;
	bitoffset &= 7;
	value <<= bitoffset;
	mask <<= bitoffset;
	mask = ~mask;
	ptr[0] &= mask;
	ptr[0] |= value;
	ptr[1] &= mask >> 8;
	ptr[1] |= value >> 8;
}

static void encode_ise(int quant_level, unsigned int character_count, const uint8_t* input_data, uint8_t* output_data, unsigned int bit_offset) {
	const ConstTables& ct = const_tables();
	unsigned int bits, trits, quints;
	ise_btq(quant_level, bits, trits, quints);
	unsigned int mask = (1u << bits) - 1;
	if (trits) {
		unsigned int i = 0;
		unsigned int full_trit_blocks = character_count / 5;
		for (unsigned int j = 0; j < full_trit_blocks; j++) {
			unsigned int i4v = input_data[i + 4] >> bits;
			unsigned int i3 = input_data[i + 3] >> bits;
			unsigned int i2 = input_data[i + 2] >> bits;
			unsigned int i1 = input_data[i + 1] >> bits;
			unsigned int i0 = input_data[i + 0] >> bits;
			uint8_t T = ct.integer_of_trits[i4v][i3][i2][i1][i0];
			static const uint8_t tbits[5] = {2, 2, 1, 2, 1};
			static const uint8_t tshift[5] = {0, 2, 4, 5, 7};
			for (int k = 0; k < 5; k++) {
				uint8_t pack = (uint8_t)((input_data[i++] & mask) | (((T >> tshift[k]) & ((1 << tbits[k]) - 1)) << bits));
				write_bits(pack, bits + tbits[k], bit_offset, output_data);
				bit_offset += bits + tbits[k];
			}
		}
		if (i != character_count) {
			unsigned int i4v = 0;
			unsigned int i3 = i + 3 >= character_count ? 0 : input_data[i + 3] >> bits;
			unsigned int i2 = i + 2 >= character_count ? 0 : input_data[i + 2] >> bits;
			unsigned int i1 = i + 1 >= character_count ? 0 : input_data[i + 1] >> bits;
			unsigned int i0 = input_data[i + 0] >> bits;
			uint8_t T = ct.integer_of_trits[i4v][i3][i2][i1][i0];
			static const uint8_t tbits[4] = {2, 2, 1, 2};
			static const uint8_t tshift[4] = {0, 2, 4, 5};
			for (unsigned int j = 0; i < character_count; i++, j++) {
				uint8_t pack = (uint8_t)((input_data[i] & mask) | (((T >> tshift[j]) & ((1 << tbits[j]) - 1)) << bits));
				write_bits(pack, bits + tbits[j], bit_offset, output_data);
				bit_offset += bits + tbits[j];
			}
		}
	} else if (quints) {
		unsigned int i = 0;
		unsigned int full_quint_blocks = character_count / 3;
		static const uint8_t qbits[3] = {3, 2, 2};
		static const uint8_t qshift[3] = {0, 3, 5};
		for (unsigned int j = 0; j < full_quint_blocks; j++) {
			unsigned int i2 = input_data[i + 2] >> bits;
			unsigned int i1 = input_data[i + 1] >> bits;
			unsigned int i0 = input_data[i + 0] >> bits;
			uint8_t T = ct.integer_of_quints[i2][i1][i0];
			for (int k = 0; k < 3; k++) {
				uint8_t pack = (uint8_t)((input_data[i++] & mask) | (((T >> qshift[k]) & ((1 << qbits[k]) - 1)) << bits));
				write_bits(pack, bits + qbits[k], bit_offset, output_data);
				bit_offset += bits + qbits[k];
			}
		}
		if (i != character_count) {
			unsigned int i2 = 0;
			unsigned int i1 = i + 1 >= character_count ? 0 : input_data[i + 1] >> bits;
			unsigned int i0 = input_data[i + 0] >> bits;
			uint8_t T = ct.integer_of_quints[i2][i1][i0];
			for (unsigned int j = 0; i < character_count; i++, j++) {
				uint8_t pack = (uint8_t)((input_data[i] & mask) | (((T >> qshift[j]) & ((1 << qbits[j]) - 1)) << bits));
				write_bits(pack, bits + qbits[j], bit_offset, output_data);
				bit_offset += bits + qbits[j];
			}
		}
	} else {
		for (unsigned int i = 0; i < character_count; i++) {
			write_bits(input_data[i], bits, bit_offset, output_data);
			bit_offset += bits;
		}
	}
}

static inline int bitrev8(int p) {
	p = ((p & 0x0F) << 4) | ((p >> 4) & 0x0F);
	p = ((p & 0x33) << 2) | ((p >> 2) & 0x33);
	p = ((p & 0x55) << 1) | ((p >> 1) & 0x55);
	return p;
}

void symbolic_to_physical(const BlockSizeTables& bsd, const SymbolicBlock& scb, uint8_t pcb_out[16]) {
	// the reference writes up to one byte past some fields via its 2-byte write_bits; give it slack
	uint8_t pcb[18];
	memset(pcb, 0, sizeof(pcb));
	if (scb.block_type == SYM_BTYPE_CONST_U16 || scb.block_type == SYM_BTYPE_CONST_F16) {
		static const uint8_t cbytes_u16[8] = {0xFC, 0xFD, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF};
		static const uint8_t cbytes_f16[8] = {0xFC, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF};
		const uint8_t* cb = scb.block_type == SYM_BTYPE_CONST_U16 ? cbytes_u16 : cbytes_f16;
		for (int i = 0; i < 8; i++) {
			pcb_out[i] = cb[i];
		}
		for (int i = 0; i < 4; i++) {
			pcb_out[2 * i + 8] = (uint8_t)(scb.constant_color[i] & 0xFF);
			pcb_out[2 * i + 9] = (uint8_t)((scb.constant_color[i] >> 8) & 0xFF);
		}
		return;
	}
	const ConstTables& ct = const_tables();
	unsigned int partition_count = scb.partition_count;
	uint8_t weightbuf[18];
	memset(weightbuf, 0, sizeof(weightbuf));
	const BlockMode& bm = bsd.block_modes[bsd.block_mode_packed_index[scb.block_mode]];
	const DecimationInfo& di = bsd.decimation_tables[bm.decimation_mode];
	int weight_count = di.weight_count;
	int weight_quant_method = bm.quant_mode;
	float weight_quant_levels = static_cast<float>(get_quant_level(weight_quant_method));
	int is_dual_plane = bm.is_dual_plane;
	const WeightQuantTable& qat = ct.weight_quant[weight_quant_method];
	int real_weight_count = is_dual_plane ? 2 * weight_count : weight_count;
	int bits_for_weights = (int)ise_sequence_bitcount((unsigned int)real_weight_count, weight_quant_method);
	uint8_t weights[64];
	if (is_dual_plane) {
		for (int i = 0; i < weight_count; i++) {
			float uqw = static_cast<float>(scb.weights[i]);
			float qw = (uqw / 64.0f) * (weight_quant_levels - 1.0f);
			int qwi = static_cast<int>(qw + 0.5f);
			weights[2 * i] = qat.scramble_map[qwi];
			uqw = static_cast<float>(scb.weights[i + PLANE2_OFFSET]);
			qw = (uqw / 64.0f) * (weight_quant_levels - 1.0f);
			qwi = static_cast<int>(qw + 0.5f);
			weights[2 * i + 1] = qat.scramble_map[qwi];
		}
	} else {
		for (int i = 0; i < weight_count; i++) {
			float uqw = static_cast<float>(scb.weights[i]);
			float qw = (uqw / 64.0f) * (weight_quant_levels - 1.0f);
			int qwi = static_cast<int>(qw + 0.5f);
			weights[i] = qat.scramble_map[qwi];
		}
	}
	encode_ise(weight_quant_method, (unsigned int)real_weight_count, weights, weightbuf, 0);
	for (int i = 0; i < 16; i++) {
		pcb[i] = static_cast<uint8_t>(bitrev8(weightbuf[15 - i]));
	}
	write_bits(scb.block_mode, 11, 0, pcb);
	write_bits(partition_count - 1, 2, 11, pcb);
	int below_weights_pos = 128 - bits_for_weights;
	if (partition_count > 1) {
		write_bits(scb.partition_index, 6, 13, pcb);
		write_bits(scb.partition_index >> 6, 10 - 6, 19, pcb);
		if (scb.color_formats_matched) {
			write_bits((unsigned int)scb.color_formats[0] << 2, 6, 13 + 10, pcb);
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
			write_bits((unsigned int)encoded_type_lowpart, 6, 13 + 10, pcb);
			write_bits((unsigned int)encoded_type_highpart, (unsigned int)encoded_type_highpart_size, (unsigned int)encoded_type_highpart_pos, pcb);
			below_weights_pos -= encoded_type_highpart_size;
		}
	} else {
		write_bits(scb.color_formats[0], 4, 13, pcb);
	}
	if (is_dual_plane) {
		write_bits((unsigned int)scb.plane2_component, 2, (unsigned int)(below_weights_pos - 2), pcb);
	}
	uint8_t values_to_encode[32];
	int valuecount_to_encode = 0;
	const uint8_t* pack_table = ct.color_uquant_to_scrambled_pquant[scb.quant_mode - QUANT_6];
	for (unsigned int i = 0; i < scb.partition_count; i++) {
		int vals = 2 * (scb.color_formats[i] >> 2) + 2;
		for (int j = 0; j < vals; j++) {
			values_to_encode[j + valuecount_to_encode] = pack_table[scb.color_values[i][j]];
		}
		valuecount_to_encode += vals;
	}
	encode_ise(scb.quant_mode, (unsigned int)valuecount_to_encode, values_to_encode, pcb, scb.partition_count == 1 ? 17 : 19 + 10);
	memcpy(pcb_out, pcb, 16);
}

// =============================================================================================
// Partition search (astcenc_find_best_partitioning.cpp)
// =============================================================================================
static void kmeans_init(const ImageBlock& blk, unsigned int texel_count, unsigned int partition_count, f4 cluster_centers[4]) {   // :60-143
	unsigned int clusters_selected = 0;
	float distances[MAX_TEXELS];
	unsigned int sample = 145897 % texel_count;
	f4 center_color = texel4(blk, sample);
	cluster_centers[clusters_selected] = center_color;
	clusters_selected++;
	float distance_sum = 0.0f;
	for (unsigned int i = 0; i < texel_count; i++) {
		f4 diff = texel4(blk, i) - center_color;
		float distance = dot_s(diff * diff, blk.channel_weight);
		distance_sum += distance;
		distances[i] = distance;
	}
	static const float cluster_cutoffs[9] = {0.626220f, 0.932770f, 0.275454f, 0.318558f, 0.240113f, 0.009190f, 0.347661f, 0.731960f, 0.156391f};
	unsigned int cutoff = (clusters_selected - 1) + 3 * (partition_count - 2);
	while (true) {
		float summa = 0.0f;
		float distance_cutoff = distance_sum * cluster_cutoffs[cutoff++];
		for (sample = 0; sample < texel_count; sample++) {
			summa += distances[sample];
			if (summa >= distance_cutoff) {
				break;
			}
		}
		sample = sample < texel_count - 1 ? sample : texel_count - 1;
		center_color = texel4(blk, sample);
		cluster_centers[clusters_selected++] = center_color;
		if (clusters_selected >= partition_count) {
			break;
		}
		distance_sum = 0.0f;
		for (unsigned int i = 0; i < texel_count; i++) {
			f4 diff = texel4(blk, i) - center_color;
			float distance = dot_s(diff * diff, blk.channel_weight);
			distance = minf(distance, distances[i]);
			distance_sum += distance;
			distances[i] = distance;
		}
	}
}

static void kmeans_assign(const ImageBlock& blk, unsigned int texel_count, unsigned int partition_count, const f4 cluster_centers[4], uint8_t* partition_of_texel) {   // :146-207
	uint8_t partition_texel_count[4] = {0, 0, 0, 0};
	for (unsigned int i = 0; i < texel_count; i++) {
		float best_distance = 3.402823466e+38f;
		unsigned int best_partition = 0;
		f4 color = texel4(blk, i);
		for (unsigned int j = 0; j < partition_count; j++) {
			f4 diff = color - cluster_centers[j];
			float distance = dot_s(diff * diff, blk.channel_weight);
			if (distance < best_distance) {
				best_distance = distance;
				best_partition = j;
			}
		}
		partition_of_texel[i] = static_cast<uint8_t>(best_partition);
		partition_texel_count[best_partition]++;
	}
	bool problem_case;
	do {
		problem_case = false;
		for (unsigned int i = 0; i < partition_count; i++) {
			if (partition_texel_count[i] == 0) {
				partition_texel_count[partition_of_texel[i]]--;
				partition_texel_count[i]++;
				partition_of_texel[i] = static_cast<uint8_t>(i);
				problem_case = true;
			}
		}
	} while (problem_case);
}

static void kmeans_update(const ImageBlock& blk, unsigned int texel_count, unsigned int partition_count, f4 cluster_centers[4], const uint8_t* partition_of_texel) {   // :210-243
	f4 color_sum[4] = {splat4(0.0f), splat4(0.0f), splat4(0.0f), splat4(0.0f)};
	uint8_t partition_texel_count[4] = {0, 0, 0, 0};
	for (unsigned int i = 0; i < texel_count; i++) {
		uint8_t partition = partition_of_texel[i];
		color_sum[partition] = color_sum[partition] + texel4(blk, i);
		partition_texel_count[partition]++;
	}
	for (unsigned int i = 0; i < partition_count; i++) {
		float scale = 1.0f / static_cast<float>(partition_texel_count[i]);
		cluster_centers[i] = color_sum[i] * scale;
	}
}

static inline int popc64(uint64_t v) { return __builtin_popcountll(v); }
static inline int min3i(int a, int b, int c) { return mini(mini(a, b), c); }
static inline int min4i(int a, int b, int c, int d) { return mini(mini(a, b), mini(c, d)); }

static inline uint8_t partition_mismatch2(const uint64_t a[2], const uint64_t b[2]) {   // :253-263
	int v1 = popc64(a[0] ^ b[0]) + popc64(a[1] ^ b[1]);
	int v2 = popc64(a[0] ^ b[1]) + popc64(a[1] ^ b[0]);
	return static_cast<uint8_t>(mini(v1, v2) / 2);
}

static inline uint8_t partition_mismatch3(const uint64_t a[3], const uint64_t b[3]) {   // :273-304
	int p00 = popc64(a[0] ^ b[0]), p01 = popc64(a[0] ^ b[1]), p02 = popc64(a[0] ^ b[2]);
	int p10 = popc64(a[1] ^ b[0]), p11 = popc64(a[1] ^ b[1]), p12 = popc64(a[1] ^ b[2]);
	int p20 = popc64(a[2] ^ b[0]), p21 = popc64(a[2] ^ b[1]), p22 = popc64(a[2] ^ b[2]);
	int v0 = mini(p11 + p22, p12 + p21) + p00;
	int v1 = mini(p10 + p22, p12 + p20) + p01;
	int v2 = mini(p10 + p21, p11 + p20) + p02;
	return static_cast<uint8_t>(min3i(v0, v1, v2) / 2);
}

static inline uint8_t partition_mismatch4(const uint64_t a[4], const uint64_t b[4]) {   // :314-353
	int p00 = popc64(a[0] ^ b[0]), p01 = popc64(a[0] ^ b[1]), p02 = popc64(a[0] ^ b[2]), p03 = popc64(a[0] ^ b[3]);
	int p10 = popc64(a[1] ^ b[0]), p11 = popc64(a[1] ^ b[1]), p12 = popc64(a[1] ^ b[2]), p13 = popc64(a[1] ^ b[3]);
	int p20 = popc64(a[2] ^ b[0]), p21 = popc64(a[2] ^ b[1]), p22 = popc64(a[2] ^ b[2]), p23 = popc64(a[2] ^ b[3]);
	int p30 = popc64(a[3] ^ b[0]), p31 = popc64(a[3] ^ b[1]), p32 = popc64(a[3] ^ b[2]), p33 = popc64(a[3] ^ b[3]);
	int mx23 = mini(p22 + p33, p23 + p32);
	int mx13 = mini(p21 + p33, p23 + p31);
	int mx12 = mini(p21 + p32, p22 + p31);
	int mx03 = mini(p20 + p33, p23 + p30);
	int mx02 = mini(p20 + p32, p22 + p30);
	int mx01 = mini(p21 + p30, p20 + p31);
	int v0 = p00 + min3i(p11 + mx23, p12 + mx13, p13 + mx12);
	int v1 = p01 + min3i(p10 + mx23, p12 + mx03, p13 + mx02);
	int v2 = p02 + min3i(p11 + mx03, p10 + mx13, p13 + mx01);
	int v3 = p03 + min3i(p11 + mx02, p12 + mx01, p10 + mx12);
	return static_cast<uint8_t>(min4i(v0, v1, v2, v3) / 2);
}

// compute_kmeans_partition_ordering :458-509 (+ count_partition_mismatch_bits :365, counting sort :412)
static unsigned int compute_kmeans_partition_ordering(const BlockSizeTables& bsd, const ImageBlock& blk, unsigned int partition_count, uint16_t* partition_ordering) {
	f4 cluster_centers[4];
	uint8_t texel_partitions[MAX_TEXELS];
	for (unsigned int i = 0; i < 3; i++) {
		if (i == 0) {
			kmeans_init(blk, bsd.texel_count, partition_count, cluster_centers);
		} else {
			kmeans_update(blk, bsd.texel_count, partition_count, cluster_centers, texel_partitions);
		}
		kmeans_assign(blk, bsd.texel_count, partition_count, cluster_centers, texel_partitions);
	}
	uint64_t bitmaps[4] = {0, 0, 0, 0};
	unsigned int texels_to_process = bsd.texel_count < MAX_KMEANS_TEXELS ? bsd.texel_count : (unsigned int)MAX_KMEANS_TEXELS;
	for (unsigned int i = 0; i < texels_to_process; i++) {
		unsigned int idx = bsd.kmeans_texels[i];
		bitmaps[texel_partitions[idx]] |= 1ULL << i;
	}
	uint8_t mismatch_counts[MAX_PARTITIONINGS];
	unsigned int active_count = bsd.partitioning_count_selected[partition_count - 1];
	const uint64_t* cov = bsd.coverage_bitmaps[partition_count];
	for (unsigned int i = 0; i < active_count; i++) {
		if (partition_count == 2) mismatch_counts[i] = partition_mismatch2(bitmaps, cov + (size_t)i * 2);
		else if (partition_count == 3) mismatch_counts[i] = partition_mismatch3(bitmaps, cov + (size_t)i * 3);
		else mismatch_counts[i] = partition_mismatch4(bitmaps, cov + (size_t)i * 4);
	}
	uint16_t mscount[MAX_KMEANS_TEXELS];
	memset(mscount, 0, sizeof(mscount));
	for (unsigned int i = 0; i < active_count; i++) {
		mscount[mismatch_counts[i]]++;
	}
	uint16_t sum = 0;
	for (unsigned int i = 0; i < texels_to_process; i++) {
		uint16_t cnt = mscount[i];
		mscount[i] = sum;
		sum = (uint16_t)(sum + cnt);
	}
	for (unsigned int i = 0; i < active_count; i++) {
		unsigned int idx = mscount[mismatch_counts[i]]++;
		partition_ordering[idx] = static_cast<uint16_t>(i);
	}
	return active_count;
}

static void insert_result(unsigned int max_values, float this_error, unsigned int this_partition, float* best_errors, unsigned int* best_partitions) {   // :512-548
	if (this_error >= best_errors[max_values - 1]) {
		return;
	}
	for (unsigned int i = 0; i < max_values; i++) {
		if (this_error > best_errors[i]) {
			continue;
		}
		for (unsigned int j = max_values - 1; j > i; j--) {
			best_errors[j] = best_errors[j - 1];
			best_partitions[j] = best_partitions[j - 1];
		}
		best_errors[i] = this_error;
		best_partitions[i] = this_partition;
		break;
	}
}

// find_best_partition_candidates :551-780
static unsigned int find_best_partition_candidates(const BlockSizeTables& bsd, const ImageBlock& blk, unsigned int partition_count,
                                                   unsigned int partition_search_limit, unsigned int best_partitions[8], unsigned int requested_candidates) {
	unsigned int texels_per_block = bsd.texel_count;
	float weight_imprecision_estim = 0.055f;
	if (texels_per_block <= 20) weight_imprecision_estim = 0.03f;
	else if (texels_per_block <= 31) weight_imprecision_estim = 0.04f;
	else if (texels_per_block <= 41) weight_imprecision_estim = 0.05f;
	weight_imprecision_estim = weight_imprecision_estim * weight_imprecision_estim;

	uint16_t partition_sequence[MAX_PARTITIONINGS];
	unsigned int sequence_len = compute_kmeans_partition_ordering(bsd, blk, partition_count, partition_sequence);
	partition_search_limit = partition_search_limit < sequence_len ? partition_search_limit : sequence_len;
	requested_candidates = partition_search_limit < requested_candidates ? partition_search_limit : requested_candidates;
	bool uses_alpha = !is_constant_channel(blk, 3);

	float uncor_best_errors[8], samec_best_errors[8];
	unsigned int uncor_best_partitions[8] = {0, 0, 0, 0, 0, 0, 0, 0};
	unsigned int samec_best_partitions[8] = {0, 0, 0, 0, 0, 0, 0, 0};
	for (unsigned int i = 0; i < requested_candidates; i++) {
		uncor_best_errors[i] = ERROR_CALC_DEFAULT;
		samec_best_errors[i] = ERROR_CALC_DEFAULT;
	}
	for (unsigned int i = 0; i < partition_search_limit; i++) {
		unsigned int partition = partition_sequence[i];
		const PartitionInfo& pi = bsd.partitionings[partition_count][partition];
		PartitionMetrics pms[4];
		f4 uncor_b[4], samec_b[4];
		ProcessedLine uncor_pl[4], samec_pl[4];
		float line_lengths[4];
		if (uses_alpha) {
			compute_avgs_and_dirs_4_comp(pi, blk, pms);
		} else {
			compute_avgs_and_dirs_3_comp_rgb(pi, blk, pms);
		}
		for (unsigned int j = 0; j < partition_count; j++) {
			f4 a = pms[j].avg;
			uncor_b[j] = normalize_safe4(pms[j].dir, uses_alpha ? unit4() : unit3());
			samec_b[j] = normalize_safe4(pms[j].avg, uses_alpha ? unit4() : unit3());
			if (uses_alpha) {
				uncor_pl[j].amod = a - uncor_b[j] * splat4(dot_s(a, uncor_b[j]));
			} else {
				uncor_pl[j].amod = a - uncor_b[j] * dot3_splat(a, uncor_b[j]);
			}
			uncor_pl[j].bs = uncor_b[j];
			samec_pl[j].amod = splat4(0.0f);
			samec_pl[j].bs = samec_b[j];
		}
		float uncor_error = 0.0f, samec_error = 0.0f;
		compute_error_squared(pi, blk, uses_alpha ? 4 : 3, uncor_pl, samec_pl, line_lengths, uncor_error, samec_error);
		for (unsigned int j = 0; j < partition_count; j++) {
			float tpp = static_cast<float>(pi.partition_texel_count[j]);
			f4 error_weights = splat4(tpp * weight_imprecision_estim);
			f4 uncor_vector = uncor_b[j] * line_lengths[j];
			f4 samec_vector = samec_b[j] * line_lengths[j];
			if (uses_alpha) {
				uncor_error += dot_s(uncor_vector * uncor_vector, error_weights);
				samec_error += dot_s(samec_vector * samec_vector, error_weights);
			} else {
				uncor_error += dot3_s(uncor_vector * uncor_vector, error_weights);
				samec_error += dot3_s(samec_vector * samec_vector, error_weights);
			}
		}
		insert_result(requested_candidates, uncor_error, partition, uncor_best_errors, uncor_best_partitions);
		insert_result(requested_candidates, samec_error, partition, samec_best_errors, samec_best_partitions);
	}
	unsigned int interleave[16];
	for (unsigned int i = 0; i < requested_candidates; i++) {
		interleave[2 * i] = bsd.partitionings[partition_count][uncor_best_partitions[i]].partition_index;
		interleave[2 * i + 1] = bsd.partitionings[partition_count][samec_best_partitions[i]].partition_index;
	}
	uint64_t bitmasks[1024 / 64];
	memset(bitmasks, 0, sizeof(bitmasks));
	unsigned int emitted = 0;
	for (unsigned int i = 0; i < requested_candidates * 2; i++) {
		unsigned int partition = interleave[i];
		unsigned int word = partition / 64;
		unsigned int bit = partition % 64;
		bool written = (bitmasks[word] & (1ull << bit)) != 0;
		if (!written) {
			best_partitions[emitted] = partition;
			bitmasks[word] |= 1ull << bit;
			emitted++;
			if (emitted == requested_candidates) {
				break;
			}
		}
	}
	return emitted;
}

// =============================================================================================
// Trials (astcenc_compress_symbolic.cpp:353-1044)
// =============================================================================================
static inline float min_ep_cutoff(float e0, float e1, float cur) {
	float ep = (1.0f - e0) / (e1 - e0);
	bool use = (ep > 0.5f) && (ep < cur);
	return use ? ep : cur;
}

// The candidate refinement loop shared by both trial kinds (:504-699 and :886-1044)
static float refine_candidates(const Config& config, const BlockSizeTables& bsd, const ImageBlock& blk, const PartitionInfo& pi, bool dual,
                               unsigned int partition_count, unsigned int partition_index, int plane2_component, const Endpoints& base_ep,
                               unsigned int candidate_count, const uint8_t partition_format_specifiers[8][4], const int block_mode_index[8],
                               const uint8_t color_quant_level[8], const uint8_t color_quant_level_mod[8], const uint8_t* dec_weights_uquant,
                               float tune_errorval_threshold, SymbolicBlock& scb) {
	float best_errorval_in_mode = ERROR_CALC_DEFAULT;
	float best_errorval_in_scb = scb.errorval;
	bool fast_diff = !dual && (partition_count == 1) && !(config.flags & FLG_MAP_RGBM);
	for (unsigned int i = 0; i < candidate_count; i++) {
		const int bm_packed_index = block_mode_index[i];
		const BlockMode& qw_bm = bsd.block_modes[bm_packed_index];
		const DecimationInfo& di = bsd.decimation_tables[qw_bm.decimation_mode];
		f4 rgbs_colors[4], rgbo_colors[4];
		for (int k = 0; k < 4; k++) {
			rgbs_colors[k] = splat4(0.0f);
			rgbo_colors[k] = splat4(0.0f);
		}
		SymbolicBlock workscb;
		memset(&workscb, 0, sizeof(workscb));
		Endpoints workep = base_ep;
		const uint8_t* u8_weight_src = dec_weights_uquant + MAX_WEIGHTS * bm_packed_index;
		for (unsigned int j = 0; j < di.weight_count; j++) {
			workscb.weights[j] = u8_weight_src[j];
			if (dual) {
				workscb.weights[j + PLANE2_OFFSET] = u8_weight_src[j + PLANE2_OFFSET];
			}
		}
		for (unsigned int l = 0; l < config.tune_refinement_limit; l++) {
			if (dual) {
				recompute_ideal_colors_2planes(blk, bsd, di, workscb.weights, workscb.weights + PLANE2_OFFSET, workep, rgbs_colors[0], rgbo_colors[0], plane2_component);
				workscb.color_formats[0] = pack_color_endpoints(workep.endpt0[0], workep.endpt1[0], rgbs_colors[0], rgbo_colors[0],
				                                                partition_format_specifiers[i][0], workscb.color_values[0], color_quant_level[i]);
				workscb.partition_count = 1;
				workscb.partition_index = 0;
				workscb.quant_mode = color_quant_level[i];
				workscb.color_formats_matched = 0;
				workscb.block_mode = qw_bm.mode_index;
				workscb.plane2_component = static_cast<int8_t>(plane2_component);
				workscb.block_type = SYM_BTYPE_NONCONST;
			} else {
				recompute_ideal_colors_1plane(blk, pi, di, workscb.weights, workep, rgbs_colors, rgbo_colors);
				bool all_same = color_quant_level[i] != color_quant_level_mod[i];
				for (unsigned int j = 0; j < partition_count; j++) {
					workscb.color_formats[j] = pack_color_endpoints(workep.endpt0[j], workep.endpt1[j], rgbs_colors[j], rgbo_colors[j],
					                                                partition_format_specifiers[i][j], workscb.color_values[j], color_quant_level[i]);
					all_same = all_same && workscb.color_formats[j] == workscb.color_formats[0];
				}
				workscb.color_formats_matched = 0;
				if (partition_count >= 2 && all_same) {
					uint8_t colorvals[4][8];
					memset(colorvals, 0, sizeof(colorvals));
					uint8_t color_formats_mod[4] = {0, 0, 0, 0};
					bool all_same_mod = true;
					for (unsigned int j = 0; j < partition_count; j++) {
						color_formats_mod[j] = pack_color_endpoints(workep.endpt0[j], workep.endpt1[j], rgbs_colors[j], rgbo_colors[j],
						                                            partition_format_specifiers[i][j], colorvals[j], color_quant_level_mod[i]);
						if (color_formats_mod[j] != color_formats_mod[0]) {
							all_same_mod = false;
							break;
						}
					}
					if (all_same_mod) {
						workscb.color_formats_matched = 1;
						for (unsigned int j = 0; j < 4; j++) {
							for (unsigned int k = 0; k < 8; k++) {
								workscb.color_values[j][k] = colorvals[j][k];
							}
							workscb.color_formats[j] = color_formats_mod[j];
						}
					}
				}
				workscb.partition_count = static_cast<uint8_t>(partition_count);
				workscb.partition_index = static_cast<uint16_t>(partition_index);
				workscb.plane2_component = -1;
				workscb.quant_mode = workscb.color_formats_matched ? color_quant_level_mod[i] : color_quant_level[i];
				workscb.block_mode = qw_bm.mode_index;
				workscb.block_type = SYM_BTYPE_NONCONST;
			}

			if (l == 0) {
				float errorval = fast_diff ? compute_symbolic_block_difference_1plane_1partition(config, bsd, workscb, blk)
				                           : compute_symbolic_block_difference(config, bsd, workscb, blk, dual);
				if (errorval == -ERROR_CALC_DEFAULT) {
					errorval = -errorval;
					workscb.block_type = SYM_BTYPE_ERROR;
				}
				best_errorval_in_mode = minf(errorval, best_errorval_in_mode);
				unsigned int iters_remaining = config.tune_refinement_limit - l;
				float threshold = (0.045f * static_cast<float>(iters_remaining)) + 1.08f;
				if (errorval > (threshold * best_errorval_in_scb)) {
					break;
				}
				if (errorval < best_errorval_in_scb) {
					best_errorval_in_scb = errorval;
					workscb.errorval = errorval;
					scb = workscb;
					if (errorval < tune_errorval_threshold) {
						i = candidate_count;
						break;
					}
				}
			}
			bool adjustments = realign_weights(config.profile, bsd, blk, workscb, di.weight_count != bsd.texel_count);
			float errorval = fast_diff ? compute_symbolic_block_difference_1plane_1partition(config, bsd, workscb, blk)
			                           : compute_symbolic_block_difference(config, bsd, workscb, blk, dual);
			if (errorval == -ERROR_CALC_DEFAULT) {
				errorval = -errorval;
				workscb.block_type = SYM_BTYPE_ERROR;
			}
			best_errorval_in_mode = minf(errorval, best_errorval_in_mode);
			unsigned int iters_remaining = config.tune_refinement_limit - 1 - l;
			float threshold = (0.045f * static_cast<float>(iters_remaining)) + 1.0f;
			if (errorval > (threshold * best_errorval_in_scb)) {
				break;
			}
			if (errorval < best_errorval_in_scb) {
				best_errorval_in_scb = errorval;
				workscb.errorval = errorval;
				scb = workscb;
				if (errorval < tune_errorval_threshold) {
					i = candidate_count;
					break;
				}
			}
			if (!adjustments) {
				break;
			}
		}
	}
	return best_errorval_in_mode;
}

// compress_symbolic_block_for_partition_1plane :353-712
static float compress_symbolic_block_for_partition_1plane(const Config& config, const BlockSizeTables& bsd, const ImageBlock& blk, bool only_always,
                                                          float tune_errorval_threshold, unsigned int partition_count, unsigned int partition_index,
                                                          SymbolicBlock& scb, WorkBuf& tmp, int quant_limit) {
	int max_weight_quant = mini((int)QUANT_32, quant_limit);
	unsigned int packed_part = partition_count >= 2 ? bsd.partitioning_packed_index[partition_count - 2][partition_index] : 0;
	const PartitionInfo& pi = bsd.partitionings[partition_count][packed_part];
	EndpointsAndWeights& ei = tmp.ei1;
	compute_ideal_colors_and_weights_1plane(blk, pi, ei);

	float* dec_weights_ideal = tmp.dec_weights_ideal;
	uint8_t* dec_weights_uquant = tmp.dec_weights_uquant;
	unsigned int max_decimation_modes = only_always ? bsd.decimation_mode_count_always : bsd.decimation_mode_count_selected;
	uint16_t refmask = (uint16_t)((1u << (max_weight_quant + 1)) - 1);
	for (unsigned int i = 0; i < max_decimation_modes; i++) {
		if ((bsd.decimation_modes[i].refprec_1plane & refmask) == 0) {
			continue;
		}
		compute_ideal_weights_for_decimation(ei, bsd.decimation_tables[i], dec_weights_ideal + i * MAX_WEIGHTS);
	}
	f4 min_ep = splat4(10.0f);
	for (unsigned int i = 0; i < partition_count; i++) {
		const f4& e0 = ei.ep.endpt0[i];
		const f4& e1 = ei.ep.endpt1[i];
		min_ep.x = min_ep_cutoff(e0.x, e1.x, min_ep.x);
		min_ep.y = min_ep_cutoff(e0.y, e1.y, min_ep.y);
		min_ep.z = min_ep_cutoff(e0.z, e1.z, min_ep.z);
		min_ep.w = min_ep_cutoff(e0.w, e1.w, min_ep.w);
	}
	float min_wt_cutoff = hmin_s(min_ep);

	compute_angular_endpoints_1plane(only_always, bsd, dec_weights_ideal, (unsigned int)max_weight_quant, tmp);

	float* weight_low_value = tmp.weight_low_value1;
	float* weight_high_value = tmp.weight_high_value1;
	int8_t* qwt_bitcounts = tmp.qwt_bitcounts;
	float* qwt_errors = tmp.qwt_errors;
	static const int8_t free_bits_for_partition_count[4] = {115 - 4, 111 - 4 - 10, 108 - 4 - 10, 105 - 4 - 10};
	unsigned int max_block_modes = only_always ? bsd.block_mode_count_1plane_always : bsd.block_mode_count_1plane_selected;
	for (unsigned int i = 0; i < max_block_modes; i++) {
		const BlockMode& bm = bsd.block_modes[i];
		if (bm.quant_mode > max_weight_quant) {
			qwt_errors[i] = 1e38f;
			continue;
		}
		int bitcount = free_bits_for_partition_count[partition_count - 1] - bm.weight_bits;
		if (bitcount <= 0) {
			qwt_errors[i] = 1e38f;
			continue;
		}
		if (weight_high_value[i] > 1.02f * min_wt_cutoff) {
			weight_high_value[i] = 1.0f;
		}
		const DecimationInfo& di = bsd.decimation_tables[bm.decimation_mode];
		qwt_bitcounts[i] = static_cast<int8_t>(bitcount);
		float dec_weights_uquantf[MAX_WEIGHTS];
		compute_quantized_weights_for_decimation(di, weight_low_value[i], weight_high_value[i], dec_weights_ideal + MAX_WEIGHTS * bm.decimation_mode,
		                                         dec_weights_uquantf, dec_weights_uquant + MAX_WEIGHTS * i, bm.quant_mode);
		qwt_errors[i] = compute_error_of_weight_set_1plane(ei, di, dec_weights_uquantf);
	}
	uint8_t partition_format_specifiers[8][4];
	int block_mode_index[8];
	uint8_t color_quant_level[8], color_quant_level_mod[8];
	unsigned int candidate_count = compute_ideal_endpoint_formats(pi, blk, ei.ep, qwt_bitcounts, qwt_errors, config.tune_candidate_limit, 0, max_block_modes,
	                                                              partition_format_specifiers, block_mode_index, color_quant_level, color_quant_level_mod, tmp);
	return refine_candidates(config, bsd, blk, pi, false, partition_count, partition_index, -1, ei.ep, candidate_count, partition_format_specifiers,
	                         block_mode_index, color_quant_level, color_quant_level_mod, dec_weights_uquant, tune_errorval_threshold, scb);
}

// compress_symbolic_block_for_partition_2planes :715-1044
static float compress_symbolic_block_for_partition_2planes(const Config& config, const BlockSizeTables& bsd, const ImageBlock& blk, float tune_errorval_threshold,
                                                           unsigned int plane2_component, SymbolicBlock& scb, WorkBuf& tmp, int quant_limit) {
	int max_weight_quant = mini((int)QUANT_32, quant_limit);
	EndpointsAndWeights& ei1 = tmp.ei1;
	EndpointsAndWeights& ei2 = tmp.ei2;
	compute_ideal_colors_and_weights_2planes(bsd, blk, plane2_component, ei1, ei2);
	float* dec_weights_ideal = tmp.dec_weights_ideal;
	uint8_t* dec_weights_uquant = tmp.dec_weights_uquant;
	uint16_t refmask = (uint16_t)((1u << (max_weight_quant + 1)) - 1);
	for (unsigned int i = 0; i < bsd.decimation_mode_count_selected; i++) {
		if ((bsd.decimation_modes[i].refprec_2planes & refmask) == 0) {
			continue;
		}
		compute_ideal_weights_for_decimation(ei1, bsd.decimation_tables[i], dec_weights_ideal + i * MAX_WEIGHTS);
		compute_ideal_weights_for_decimation(ei2, bsd.decimation_tables[i], dec_weights_ideal + i * MAX_WEIGHTS + PLANE2_OFFSET);
	}
	f4 min_ep1 = splat4(10.0f), min_ep2 = splat4(10.0f);
	{
		const f4& a0 = ei1.ep.endpt0[0];
		const f4& a1 = ei1.ep.endpt1[0];
		min_ep1 = mk4(min_ep_cutoff(a0.x, a1.x, 10.0f), min_ep_cutoff(a0.y, a1.y, 10.0f), min_ep_cutoff(a0.z, a1.z, 10.0f), min_ep_cutoff(a0.w, a1.w, 10.0f));
		const f4& b0 = ei2.ep.endpt0[0];
		const f4& b1 = ei2.ep.endpt1[0];
		min_ep2 = mk4(min_ep_cutoff(b0.x, b1.x, 10.0f), min_ep_cutoff(b0.y, b1.y, 10.0f), min_ep_cutoff(b0.z, b1.z, 10.0f), min_ep_cutoff(b0.w, b1.w, 10.0f));
	}
	f4 m1 = min_ep1;
	set_lane(m1, (int)plane2_component, ERROR_CALC_DEFAULT);
	float min_wt_cutoff1 = hmin_s(m1);
	f4 m2 = splat4(ERROR_CALC_DEFAULT);
	set_lane(m2, (int)plane2_component, lane(min_ep2, (int)plane2_component));
	float min_wt_cutoff2 = hmin_s(m2);

	compute_angular_endpoints_2planes(bsd, dec_weights_ideal, (unsigned int)max_weight_quant, tmp);

	int8_t* qwt_bitcounts = tmp.qwt_bitcounts;
	float* qwt_errors = tmp.qwt_errors;
	unsigned int start_2plane = bsd.block_mode_count_1plane_selected;
	unsigned int end_2plane = bsd.block_mode_count_1plane_2plane_selected;
	for (unsigned int i = start_2plane; i < end_2plane; i++) {
		const BlockMode& bm = bsd.block_modes[i];
		if (bm.quant_mode > max_weight_quant) {
			qwt_errors[i] = 1e38f;
			continue;
		}
		qwt_bitcounts[i] = static_cast<int8_t>(109 - bm.weight_bits);
		if (tmp.weight_high_value1[i] > 1.02f * min_wt_cutoff1) {
			tmp.weight_high_value1[i] = 1.0f;
		}
		if (tmp.weight_high_value2[i] > 1.02f * min_wt_cutoff2) {
			tmp.weight_high_value2[i] = 1.0f;
		}
		unsigned int decimation_mode = bm.decimation_mode;
		const DecimationInfo& di = bsd.decimation_tables[decimation_mode];
		float dec_weights_uquantf[MAX_WEIGHTS];
		compute_quantized_weights_for_decimation(di, tmp.weight_low_value1[i], tmp.weight_high_value1[i], dec_weights_ideal + MAX_WEIGHTS * decimation_mode,
		                                         dec_weights_uquantf, dec_weights_uquant + MAX_WEIGHTS * i, bm.quant_mode);
		compute_quantized_weights_for_decimation(di, tmp.weight_low_value2[i], tmp.weight_high_value2[i],
		                                         dec_weights_ideal + MAX_WEIGHTS * decimation_mode + PLANE2_OFFSET, dec_weights_uquantf + PLANE2_OFFSET,
		                                         dec_weights_uquant + MAX_WEIGHTS * i + PLANE2_OFFSET, bm.quant_mode);
		qwt_errors[i] = compute_error_of_weight_set_2planes(ei1, ei2, di, dec_weights_uquantf, dec_weights_uquantf + PLANE2_OFFSET);
	}
	uint8_t partition_format_specifiers[8][4];
	int block_mode_index[8];
	uint8_t color_quant_level[8], color_quant_level_mod[8];
	// merge_endpoints :37-66
	Endpoints epm;
	epm.partition_count = 1;
	epm.endpt0[0] = ei1.ep.endpt0[0];
	epm.endpt1[0] = ei1.ep.endpt1[0];
	set_lane(epm.endpt0[0], (int)plane2_component, lane(ei2.ep.endpt0[0], (int)plane2_component));
	set_lane(epm.endpt1[0], (int)plane2_component, lane(ei2.ep.endpt1[0], (int)plane2_component));
	const PartitionInfo& pi = bsd.partitionings[1][0];
	unsigned int candidate_count = compute_ideal_endpoint_formats(pi, blk, epm, qwt_bitcounts, qwt_errors, config.tune_candidate_limit, start_2plane, end_2plane,
	                                                              partition_format_specifiers, block_mode_index, color_quant_level, color_quant_level_mod, tmp);
	return refine_candidates(config, bsd, blk, pi, true, 1, 0, (int)plane2_component, epm, candidate_count, partition_format_specifiers, block_mode_index,
	                         color_quant_level, color_quant_level_mod, dec_weights_uquant, tune_errorval_threshold, scb);
}

// prepare_block_statistics :1047-1159
static float prepare_block_statistics(int texels_per_block, const ImageBlock& blk) {
	float rs = 0.0f, gs = 0.0f, bs = 0.0f, as = 0.0f;
	float rr_var = 0.0f, gg_var = 0.0f, bb_var = 0.0f, aa_var = 0.0f;
	float rg_cov = 0.0f, rb_cov = 0.0f, ra_cov = 0.0f, gb_cov = 0.0f, ga_cov = 0.0f, ba_cov = 0.0f;
	float weight_sum = 0.0f;
	for (int i = 0; i < texels_per_block; i++) {
		float weight = hadd_s(blk.channel_weight) / 4.0f;
		weight_sum += weight;
		float r = blk.data_r[i], g = blk.data_g[i], b = blk.data_b[i], a = blk.data_a[i];
		float rw = r * weight;
		rs += rw;
		rr_var += r * rw;
		rg_cov += g * rw;
		rb_cov += b * rw;
		ra_cov += a * rw;
		float gw = g * weight;
		gs += gw;
		gg_var += g * gw;
		gb_cov += b * gw;
		ga_cov += a * gw;
		float bw = b * weight;
		bs += bw;
		bb_var += b * bw;
		ba_cov += a * bw;
		float aw = a * weight;
		as += aw;
		aa_var += a * aw;
	}
	float rpt = 1.0f / maxf(weight_sum, 1e-7f);
	rr_var -= rs * (rs * rpt);
	rg_cov -= gs * (rs * rpt);
	rb_cov -= bs * (rs * rpt);
	ra_cov -= as * (rs * rpt);
	gg_var -= gs * (gs * rpt);
	gb_cov -= bs * (gs * rpt);
	ga_cov -= as * (gs * rpt);
	bb_var -= bs * (bs * rpt);
	ba_cov -= as * (bs * rpt);
	aa_var -= as * (as * rpt);
	rg_cov *= 1.0f / sqrtf(rr_var * gg_var);
	rb_cov *= 1.0f / sqrtf(rr_var * bb_var);
	ra_cov *= 1.0f / sqrtf(rr_var * aa_var);
	gb_cov *= 1.0f / sqrtf(gg_var * bb_var);
	ga_cov *= 1.0f / sqrtf(gg_var * aa_var);
	ba_cov *= 1.0f / sqrtf(bb_var * aa_var);
	if (rg_cov != rg_cov) rg_cov = 1.0f;
	if (rb_cov != rb_cov) rb_cov = 1.0f;
	if (ra_cov != ra_cov) ra_cov = 1.0f;
	if (gb_cov != gb_cov) gb_cov = 1.0f;
	if (ga_cov != ga_cov) ga_cov = 1.0f;
	if (ba_cov != ba_cov) ba_cov = 1.0f;
	float lowest_correlation = minf(fabsf(rg_cov), fabsf(rb_cov));
	lowest_correlation = minf(lowest_correlation, fabsf(ra_cov));
	lowest_correlation = minf(lowest_correlation, fabsf(gb_cov));
	lowest_correlation = minf(lowest_correlation, fabsf(ga_cov));
	lowest_correlation = minf(lowest_correlation, fabsf(ba_cov));
	return lowest_correlation;
}

static void constant_color_u16(const ImageBlock& blk, SymbolicBlock& scb) {
	scb.block_type = SYM_BTYPE_CONST_U16;
	f4 c = vclamp4(0.0f, 1.0f, blk.origin_texel) * 65535.0f;
	scb.constant_color[0] = f2i_rtn(c.x);
	scb.constant_color[1] = f2i_rtn(c.y);
	scb.constant_color[2] = f2i_rtn(c.z);
	scb.constant_color[3] = f2i_rtn(c.w);
}

// compress_block :1162-1455
void compress_block(const Context& ctx, const ImageBlock& blk, uint8_t pcb[16]) {
	const Config& config = ctx.config;
	const BlockSizeTables& bsd = *ctx.bsd;
	WorkBuf& tmp = *static_cast<WorkBuf*>(ctx.work);
	int decode_mode = config.profile;
	SymbolicBlock scb;
	memset(&scb, 0, sizeof(scb));

	bool block_is_l = is_luminance(blk);
	float block_is_l_scale = block_is_l ? 1.0f / 1.5f : 1.0f;
	bool block_is_la = is_luminancealpha(blk);
	float block_is_la_scale = block_is_la ? 1.0f / 1.05f : 1.0f;
	int max_partitions = (int)config.tune_partition_count_limit;
	unsigned int requested_partition_indices[3] = {config.tune_2partition_index_limit, config.tune_3partition_index_limit, config.tune_4partition_index_limit};
	unsigned int requested_partition_trials[3] = {config.tune_2partitioning_candidate_limit, config.tune_3partitioning_candidate_limit,
	                                              config.tune_4partitioning_candidate_limit};

	if (blk.data_min.x == blk.data_max.x && blk.data_min.y == blk.data_max.y && blk.data_min.z == blk.data_max.z && blk.data_min.w == blk.data_max.w) {
		scb.partition_count = 0;
		if (decode_mode == PRF_HDR || decode_mode == PRF_HDR_RGB_LDR_A) {
			scb.block_type = SYM_BTYPE_CONST_F16;
			scb.constant_color[0] = float_to_sf16(blk.origin_texel.x);
			scb.constant_color[1] = float_to_sf16(blk.origin_texel.y);
			scb.constant_color[2] = float_to_sf16(blk.origin_texel.z);
			scb.constant_color[3] = float_to_sf16(blk.origin_texel.w);
		} else {
			constant_color_u16(blk, scb);
		}
		symbolic_to_physical(bsd, scb, pcb);
		return;
	}

	float error_weight_sum = hadd_s(blk.channel_weight) * bsd.texel_count;
	float error_threshold = config.tune_db_limit * error_weight_sum * block_is_l_scale * block_is_la_scale;

	scb.errorval = ERROR_CALC_DEFAULT;
	scb.block_type = SYM_BTYPE_ERROR;
	float best_errorvals_for_pcount[4] = {ERROR_CALC_DEFAULT, ERROR_CALC_DEFAULT, ERROR_CALC_DEFAULT, ERROR_CALC_DEFAULT};
	float exit_thresholds_for_pcount[4] = {0.0f, config.tune_2partition_early_out_limit_factor, config.tune_3partition_early_out_limit_factor, 0.0f};
	float errorval_mult[2] = {1.0f / config.tune_mse_overshoot, 1.0f};
	const float errorval_overshoot = 1.0f / config.tune_mse_overshoot;
	int start_trial = 1;
	if (config.tune_search_mode0_enable >= 0.85f && bsd.dim_z == 1) {      // (:1287: no mode-0 trial for 3D block sizes)
		start_trial = 0;
	}
	int quant_limit = QUANT_32;
	bool done = false;
	for (int i = start_trial; i < 2 && !done; i++) {
		float errorval = compress_symbolic_block_for_partition_1plane(config, bsd, blk, i == 0, error_threshold * errorval_mult[i] * errorval_overshoot,
		                                                              1, 0, scb, tmp, QUANT_32);
		if (scb.block_type != SYM_BTYPE_ERROR) {
			quant_limit = bsd.block_modes[bsd.block_mode_packed_index[scb.block_mode]].quant_mode;
		}
		best_errorvals_for_pcount[0] = minf(best_errorvals_for_pcount[0], errorval);
		if (errorval < (error_threshold * errorval_mult[i])) {
			done = true;
		}
	}
	if (!done) {
		float lowest_correl = prepare_block_statistics(bsd.texel_count, blk);
		bool block_skip_two_plane = lowest_correl > config.tune_2plane_early_out_limit_correlation;
		for (int i = 3; i >= 0 && !done; i--) {
			if (block_skip_two_plane) {
				continue;
			}
			if (blk.grayscale && i != 3) {
				continue;
			}
			if (is_constant_channel(blk, i)) {
				continue;
			}
			float errorval = compress_symbolic_block_for_partition_2planes(config, bsd, blk, error_threshold * errorval_overshoot, (unsigned int)i, scb, tmp, quant_limit);
			if (errorval > (best_errorvals_for_pcount[0] * 1.85f)) {
				break;
			}
			if (errorval < error_threshold) {
				done = true;
			}
		}
	}
	for (int partition_count = 2; partition_count <= max_partitions && !done; partition_count++) {
		unsigned int partition_indices[8];
		unsigned int requested_indices = requested_partition_indices[partition_count - 2];
		unsigned int requested_trials = requested_partition_trials[partition_count - 2];
		requested_trials = requested_trials < requested_indices ? requested_trials : requested_indices;
		unsigned int actual_trials = find_best_partition_candidates(bsd, blk, (unsigned int)partition_count, requested_indices, partition_indices, requested_trials);
		float best_error_in_prev = best_errorvals_for_pcount[partition_count - 2];
		for (unsigned int i = 0; i < actual_trials && !done; i++) {
			float errorval = compress_symbolic_block_for_partition_1plane(config, bsd, blk, false, error_threshold * errorval_overshoot,
			                                                              (unsigned int)partition_count, partition_indices[i], scb, tmp, quant_limit);
			best_errorvals_for_pcount[partition_count - 1] = minf(best_errorvals_for_pcount[partition_count - 1], errorval);
			float best_error = best_errorvals_for_pcount[partition_count - 1];
			float best_error_scale = exit_thresholds_for_pcount[partition_count - 1] * 1.85f;
			if (best_error > (best_error_in_prev * best_error_scale)) {
				done = true;
				break;
			}
			if (errorval < error_threshold) {
				done = true;
				break;
			}
		}
		if (done) {
			break;
		}
		float best_error = best_errorvals_for_pcount[partition_count - 1];
		float best_error_scale = exit_thresholds_for_pcount[partition_count - 1];
		if (best_error > (best_error_in_prev * best_error_scale)) {
			done = true;
		}
	}
	if (scb.block_type == SYM_BTYPE_ERROR) {
		constant_color_u16(blk, scb);
	}
	symbolic_to_physical(bsd, scb, pcb);
}

// brent_kung_prefix_sum (astcenc_compute_variance.cpp:52-100) on one float lane: in-place inclusive prefix sum whose
// association order is the reference's reduction tree followed by its expansion tree.
static void brent_kung_prefix_sum(float* d, size_t items, size_t stride) {
	if (items < 2) {
		return;
	}
	size_t lc_stride = 2;
	size_t log2_stride = 1;
	do {
		size_t step = lc_stride >> 1;
		size_t start = lc_stride - 1;
		size_t iters = items >> log2_stride;
		float* da = d + (start * stride);
		ptrdiff_t ofs = -static_cast<ptrdiff_t>(step * stride);
		size_t ofs_stride = stride << log2_stride;
		while (iters) {
			*da = *da + da[ofs];
			da += ofs_stride;
			iters--;
		}
		log2_stride += 1;
		lc_stride <<= 1;
	} while (lc_stride <= items);
	do {
		log2_stride -= 1;
		lc_stride >>= 1;
		size_t step = lc_stride >> 1;
		size_t start = step + lc_stride - 1;
		size_t iters = (items - step) >> log2_stride;
		float* da = d + (start * stride);
		ptrdiff_t ofs = -static_cast<ptrdiff_t>(step * stride);
		size_t ofs_stride = stride << log2_stride;
		while (iters) {
			*da = *da + da[ofs];
			da += ofs_stride;
			iters--;
		}
	} while (lc_stride > 2);
}

// The alpha channel of the swizzled input texel as the averaging pass sees it (compute_variance.cpp:158-360)
static inline float alpha_for_average(const void* data, int data_type, unsigned int dim_x, unsigned int x, unsigned int y, int swz_a) {
	size_t o = (4 * (size_t)dim_x * y) + 4 * (size_t)x;
	if (data_type == 0) {
		const uint8_t* p = static_cast<const uint8_t*>(data) + o;
		int v = swz_a < 4 ? p[swz_a] : (swz_a == 4 ? 0 : 255);
		return static_cast<float>(v) * (1.0f / 255.0f);
	}
	if (data_type == 1) {
		const uint16_t* p = static_cast<const uint16_t*>(data) + o;
		int v = swz_a < 4 ? p[swz_a] : (swz_a == 4 ? 0 : 0x3C00);
		// float16_to_float(vint4) of the F16C builds saturates the packed value (astcenc_vecmathlib_sse_4.h:1001)
		return sf16_to_float((uint16_t)(v > 0x7FFF ? 0x7FFF : v));
	}
	const float* p = static_cast<const float*>(data) + o;
	return swz_a < 4 ? p[swz_a] : (swz_a == 4 ? 0.0f : 1.0f);
}

// compute_averages + compute_pixel_region_variance for a 2D image (astcenc_entry.cpp:1056-1108,
// astcenc_compute_variance.cpp:103-500): per 32x32 tile a summed-area table of the alpha channel, box average of radius r.
static void compute_alpha_averages(const void* data, int data_type, unsigned int dim_x, unsigned int dim_y, const int swz[4], unsigned int radius, float* averages) {
	const size_t step = 32;
	size_t kerneldim = 2 * (size_t)radius + 1;
	std::vector<float> buf((step + kerneldim) * (step + kerneldim));
	float alpha_kdim = static_cast<float>(2 * radius + 1);
	float alpha_rsamples = 1.0f / (alpha_kdim * alpha_kdim);
	for (size_t oy = 0; oy < dim_y; oy += step) {
		size_t size_y = std::min(step, (size_t)dim_y - oy);
		for (size_t ox = 0; ox < dim_x; ox += step) {
			size_t size_x = std::min(step, (size_t)dim_x - ox);
			size_t padsize_x = size_x + kerneldim, padsize_y = size_y + kerneldim;
			size_t yst = padsize_x;
			for (size_t y = 1; y < padsize_y; y++) {
				size_t y_src = (y - 1) + oy;
				y_src = y_src <= radius ? 0 : y_src - radius;
				y_src = std::min(y_src, (size_t)dim_y - 1);
				for (size_t x = 1; x < padsize_x; x++) {
					size_t x_src = (x - 1) + ox;
					x_src = x_src <= radius ? 0 : x_src - radius;
					x_src = std::min(x_src, (size_t)dim_x - 1);
					buf[y * yst + x] = alpha_for_average(data, data_type, dim_x, (unsigned int)x_src, (unsigned int)y_src, swz[3]);
				}
			}
			for (size_t y = 0; y < padsize_y; y++) buf[y * yst] = 0.0f;
			for (size_t x = 0; x < padsize_x; x++) buf[x] = 0.0f;
			for (size_t y = 1; y < padsize_y; y++) {
				brent_kung_prefix_sum(&buf[y * yst + 1], padsize_x - 1, 1);
			}
			for (size_t x = 1; x < padsize_x; x++) {
				brent_kung_prefix_sum(&buf[1 * yst + x], padsize_y - 1, yst);
			}
			for (size_t y = 0; y < size_y; y++) {
				size_t y_src = y + radius;
				size_t y_low = y_src - radius, y_high = y_src + radius + 1;
				for (size_t x = 0; x < size_x; x++) {
					size_t x_src = x + radius;
					size_t x_low = x_src - radius, x_high = x_src + radius + 1;
					float vasum = buf[y_low * yst + x_low] - buf[y_low * yst + x_high] - buf[y_high * yst + x_low] + buf[y_high * yst + x_high];
					averages[(y + oy) * dim_x + (x + ox)] = vasum * alpha_rsamples;
				}
			}
		}
	}
}

void compress_image(const Context& ctx, const void* data, int data_type, unsigned int dim_x, unsigned int dim_y, const int swz[4], uint8_t* out, unsigned int dim_z) {
	unsigned int bx = ctx.bsd->dim_x, by = ctx.bsd->dim_y, bz = ctx.bsd->dim_z;
	unsigned int blocks_x = (dim_x + bx - 1) / bx;
	unsigned int blocks_y = (dim_y + by - 1) / by;
	unsigned int blocks_z = (dim_z + bz - 1) / bz;
	ImageBlock blk;
	std::vector<float> alpha_averages;
	// (the alpha averages only steer 2D block sizes, astcenc_entry.cpp:975; volumes with a radius are not restated)
	unsigned int radius = bz == 1 && dim_z == 1 ? ctx.config.a_scale_radius : 0;
	if (radius != 0) {
		alpha_averages.resize((size_t)dim_x * dim_y);
		compute_alpha_averages(data, data_type, dim_x, dim_y, swz, radius, alpha_averages.data());
	}
	for (unsigned int z = 0; z < blocks_z; z++) {
	for (unsigned int y = 0; y < blocks_y; y++) {
		for (unsigned int x = 0; x < blocks_x; x++) {
			// alpha-scale RDO (astcenc_entry.cpp:973-1003): blocks whose footprint has (almost) no alpha become constant zero
			bool use_full_block = true;
			if (radius != 0) {
				size_t start_x = (size_t)x * bx, end_x = std::min((size_t)dim_x, start_x + bx);
				size_t start_y = (size_t)y * by, end_y = std::min((size_t)dim_y, start_y + by);
				size_t x_footprint = bx + 2 * ((size_t)radius - 1);
				size_t y_footprint = by + 2 * ((size_t)radius - 1);
				float footprint = static_cast<float>(x_footprint * y_footprint);
				float threshold = 0.9f / (255.0f * footprint);
				use_full_block = false;
				for (size_t ay = start_y; ay < end_y && !use_full_block; ay++) {
					for (size_t ax = start_x; ax < end_x; ax++) {
						if (alpha_averages[ay * dim_x + ax] > threshold) {
							use_full_block = true;
							break;
						}
					}
				}
			}
			if (use_full_block) {
				load_block(ctx, data, data_type, dim_x, dim_y, x * bx, y * by, swz, blk, dim_z, z * bz);
				if (ctx.config.flags & FLG_USE_ALPHA_WEIGHT) {
					float alpha_scale = blk.data_max.w * (1.0f / 65535.0f);
					blk.channel_weight = mk4(ctx.config.cw_r_weight * alpha_scale, ctx.config.cw_g_weight * alpha_scale,
					                         ctx.config.cw_b_weight * alpha_scale, ctx.config.cw_a_weight);
				}
			} else {
				blk.origin_texel = splat4(0.0f);
				blk.data_min = splat4(0.0f);
				blk.data_mean = splat4(0.0f);
				blk.data_max = splat4(0.0f);
				blk.grayscale = true;
			}
			compress_block(ctx, blk, out + (((size_t)z * blocks_y + y) * blocks_x + x) * 16);
		}
	}
	}
}
