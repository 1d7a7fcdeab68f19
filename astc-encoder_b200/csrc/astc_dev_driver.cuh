// B200-native ASTC block compressor: trial functions and the compress_block search driver
// (astcenc_compress_symbolic.cpp:353-1455). Control flow is uniform across the warp.
#pragma once

ASTC_FN float min_ep_cutoff(float e0, float e1, float cur) {
	float ep = (1.0f - e0) / (e1 - e0);
	bool use = (ep > 0.5f) && (ep < cur);
	return use ? ep : cur;
}

// Quantise the decimated ideal weights of one winning block mode straight into w.work_weights
// (the reference keeps every mode's quantised set in dec_weights_uquant; we recompute the winners).
ASTC_COOP void quantize_candidate_weights(WCtx& w, const DevBlockMode& bm, const DecView& di, int nplanes, float cutoff1, float cutoff2) {
	int W = di.W;
	for (int id = w.lane; id < W * nplanes; id += ASTC_WARP) {
		int pl = id >= W ? 1 : 0;
		int k = id - pl * W;
		float low, high;
		mode_low_high(w, bm, pl, pl ? cutoff2 : cutoff1, low, high);
		WeightQuantizer z = make_weight_quantizer(low, high, bm.quant_mode);
		w.work_weights[pl * 32 + k] = (uint8_t)quantize_weight(z, w.dwi[di.dm->dwi_offset + pl * W + k]);
	}
	wsync();
}

ASTC_COOP void copy_work_to_best(WCtx& w) {
	for (int i = w.lane; i < 64; i += ASTC_WARP) {
		w.best_weights[i] = w.work_weights[i];
	}
	for (int i = w.lane; i < 32; i += ASTC_WARP) {
		w.best_colors[i] = w.work_colors[i];
	}
	wsync();
}

// The candidate refinement loop shared by both trial kinds (:504-699 and :886-1044).
// base endpoints are in slots EP_BASE_*; returns best_errorval_in_mode.
ASTC_COOP float refine_candidates(WCtx& w, const PartView& pi, bool dual, unsigned int partition_count, unsigned int partition_index,
                                  int plane2_component, unsigned int candidate_count, float cutoff1, float cutoff2,
                                  float tune_errorval_threshold, ScbHdr& scb) {
	const DevBsd& bsd = *w.bsd;
	const DevConfig& cfg = *w.cfg;
	RefineScratch rsv = make_refine_scratch(w);
	RefineScratch* rs = &rsv;
	float best_errorval_in_mode = ERROR_CALC_DEFAULT;
	float best_errorval_in_scb = scb.errorval;
	const Candidate* cands = reinterpret_cast<const Candidate*>(w.cand);
	for (unsigned int i = 0; i < candidate_count; i++) {
		Candidate cd = cands[i];
		const DevBlockMode qw_bm = bsd.block_modes[cd.block_mode];
		DecView di = dec_view(bsd, qw_bm.decimation_mode);
		quantize_candidate_weights(w, qw_bm, di, dual ? 2 : 1, cutoff1, cutoff2);
		for (int k = w.lane; k < 4; k += ASTC_WARP) {
			w.ep[EP_WORK_0 + k] = w.ep[EP_BASE_0 + k];
			w.ep[EP_WORK_1 + k] = w.ep[EP_BASE_1 + k];
			w.ep[EP_RGBS + k] = splat4(0.0f);
			w.ep[EP_RGBO + k] = splat4(0.0f);
		}
		for (int k = w.lane; k < 32; k += ASTC_WARP) {
			w.work_colors[k] = 0;
		}
		wsync();
		ScbHdr work;
		work.errorval = 0.0f;
		work.color_formats[0] = work.color_formats[1] = work.color_formats[2] = work.color_formats[3] = 0;
		work.constant_color[0] = work.constant_color[1] = work.constant_color[2] = work.constant_color[3] = 0;
		for (unsigned int l = 0; l < cfg.tune_refinement_limit; l++) {
			if (dual) {
				recompute_ideal_colors_2planes(w, di, plane2_component, rs);
				if (w.lane == 0) {
					w.tmpf[0] = (float)pack_color_endpoints(w.ep[EP_WORK_0], w.ep[EP_WORK_1], w.ep[EP_RGBS], w.ep[EP_RGBO], cd.formats[0], w.work_colors, cd.quant_level);
				}
				wsync();
				work.color_formats[0] = (uint8_t)w.tmpf[0];
				wsync();
				work.partition_count = 1;
				work.partition_index = 0;
				work.quant_mode = cd.quant_level;
				work.color_formats_matched = 0;
				work.block_mode = qw_bm.mode_index;
				work.plane2_component = static_cast<int8_t>(plane2_component);
				work.block_type = SYM_BTYPE_NONCONST;
			} else {
				recompute_ideal_colors_1plane(w, pi, di, rs);
				for (unsigned int j = (unsigned int)w.lane; j < partition_count; j += ASTC_WARP) {
					w.tmpf[j] = (float)pack_color_endpoints(w.ep[EP_WORK_0 + j], w.ep[EP_WORK_1 + j], w.ep[EP_RGBS + j], w.ep[EP_RGBO + j],
					                                        cd.formats[j], w.work_colors + j * 8, cd.quant_level);
				}
				wsync();
				bool all_same = cd.quant_level != cd.quant_level_mod;
				for (unsigned int j = 0; j < partition_count; j++) {
					work.color_formats[j] = (uint8_t)w.tmpf[j];
					all_same = all_same && work.color_formats[j] == work.color_formats[0];
				}
				wsync();
				work.color_formats_matched = 0;
				if (partition_count >= 2 && all_same) {
					for (int k = w.lane; k < 32; k += ASTC_WARP) {
						w.mod_colors[k] = 0;
					}
					wsync();
					for (unsigned int j = (unsigned int)w.lane; j < partition_count; j += ASTC_WARP) {
						w.tmpf[j] = (float)pack_color_endpoints(w.ep[EP_WORK_0 + j], w.ep[EP_WORK_1 + j], w.ep[EP_RGBS + j], w.ep[EP_RGBO + j],
						                                        cd.formats[j], w.mod_colors + j * 8, cd.quant_level_mod);
					}
					wsync();
					uint8_t color_formats_mod[4] = {0, 0, 0, 0};
					bool all_same_mod = true;
					for (unsigned int j = 0; j < partition_count; j++) {
						color_formats_mod[j] = (uint8_t)w.tmpf[j];
						if (color_formats_mod[j] != color_formats_mod[0]) {
							all_same_mod = false;
							// the reference stops packing at the first mismatch; later formats stay 0 but are unused
							break;
						}
					}
					wsync();
					if (all_same_mod) {
						work.color_formats_matched = 1;
						for (int k = w.lane; k < 32; k += ASTC_WARP) {
							w.work_colors[k] = w.mod_colors[k];
						}
						for (unsigned int j = 0; j < 4; j++) {
							work.color_formats[j] = color_formats_mod[j];
						}
						wsync();
					}
				}
				work.partition_count = static_cast<uint8_t>(partition_count);
				work.partition_index = static_cast<uint16_t>(partition_index);
				work.plane2_component = -1;
				work.quant_mode = work.color_formats_matched ? cd.quant_level_mod : cd.quant_level;
				work.block_mode = qw_bm.mode_index;
				work.block_type = SYM_BTYPE_NONCONST;
			}

			bool stop_all = false;
			for (unsigned int j = 0; j < partition_count; j++) {
				TRACE("refine cand=%u l=%u fmt[%u]=%u colors %u %u %u %u %u %u %u %u\n", i, l, j, work.color_formats[j], w.work_colors[j*8], w.work_colors[j*8+1], w.work_colors[j*8+2], w.work_colors[j*8+3],
				      w.work_colors[j*8+4], w.work_colors[j*8+5], w.work_colors[j*8+6], w.work_colors[j*8+7]);
			}
			if (l == 0) {
				float errorval = compute_symbolic_block_difference(w, work, pi, di, dual, rs);
				TRACE_F("err_pre", errorval);
				if (errorval == -ERROR_CALC_DEFAULT) {
					errorval = -errorval;
					work.block_type = SYM_BTYPE_ERROR;
				}
				best_errorval_in_mode = minf(errorval, best_errorval_in_mode);
				unsigned int iters_remaining = cfg.tune_refinement_limit - l;
				float threshold = (0.045f * static_cast<float>(iters_remaining)) + 1.08f;
				if (errorval > (threshold * best_errorval_in_scb)) {
					break;
				}
				if (errorval < best_errorval_in_scb) {
					best_errorval_in_scb = errorval;
					work.errorval = errorval;
					scb = work;
					copy_work_to_best(w);
					if (errorval < tune_errorval_threshold) {
						stop_all = true;
					}
				}
			}
			if (stop_all) {
				i = candidate_count;
				break;
			}
			bool adjustments = realign_weights(w, work, pi, qw_bm, di, rs);
			float errorval = compute_symbolic_block_difference(w, work, pi, di, dual, rs);
			TRACE_F("err_post", errorval);
			if (errorval == -ERROR_CALC_DEFAULT) {
				errorval = -errorval;
				work.block_type = SYM_BTYPE_ERROR;
			}
			best_errorval_in_mode = minf(errorval, best_errorval_in_mode);
			unsigned int iters_remaining = cfg.tune_refinement_limit - 1 - l;
			float threshold = (0.045f * static_cast<float>(iters_remaining)) + 1.0f;
			if (errorval > (threshold * best_errorval_in_scb)) {
				break;
			}
			if (errorval < best_errorval_in_scb) {
				best_errorval_in_scb = errorval;
				work.errorval = errorval;
				scb = work;
				copy_work_to_best(w);
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
ASTC_COOP float compress_symbolic_block_for_partition_1plane(WCtx& w, bool only_always, float tune_errorval_threshold, unsigned int partition_count,
                                                             unsigned int partition_index, ScbHdr& scb, int quant_limit) {
	const DevBsd& bsd = *w.bsd;
	int max_weight_quant = mini((int)QUANT_32, quant_limit);
	PartView pi = part_view(bsd, partition_count, partition_index);
	compute_ideal_colors_and_weights_1plane(w, pi);

	unsigned int max_decimation_modes = only_always ? bsd.decimation_mode_count_always : bsd.decimation_mode_count_selected;
	uint16_t refmask = (uint16_t)((1u << (max_weight_quant + 1)) - 1);
	for (unsigned int i = 0; i < max_decimation_modes; i++) {
		if ((bsd.dec_modes[i].refprec_1plane & refmask) == 0) {
			continue;
		}
		DecView di = dec_view(bsd, i);
		compute_ideal_weights_for_decimation(w, di, 1, w.dwi + di.dm->dwi_offset, 0);
	}
	f4 min_ep = splat4(10.0f);
	for (unsigned int i = 0; i < partition_count; i++) {
		f4 e0 = w.ep[EP_EI1_0 + i];
		f4 e1 = w.ep[EP_EI1_1 + i];
		min_ep.x = min_ep_cutoff(e0.x, e1.x, min_ep.x);
		min_ep.y = min_ep_cutoff(e0.y, e1.y, min_ep.y);
		min_ep.z = min_ep_cutoff(e0.z, e1.z, min_ep.z);
		min_ep.w = min_ep_cutoff(e0.w, e1.w, min_ep.w);
	}
	float min_wt_cutoff = hmin_s(min_ep);
	TRACE("trial1p pc=%u pidx=%u only_always=%d maxq=%d\n", partition_count, partition_index, (int)only_always, max_weight_quant);
	for (unsigned int i = 0; i < partition_count; i++) {
		TRACE("ep0[%u] %08x %08x %08x %08x ep1 %08x %08x %08x %08x\n", i, ASTC_F2U(w.ep[EP_EI1_0 + i].x), ASTC_F2U(w.ep[EP_EI1_0 + i].y), ASTC_F2U(w.ep[EP_EI1_0 + i].z), ASTC_F2U(w.ep[EP_EI1_0 + i].w),
		      ASTC_F2U(w.ep[EP_EI1_1 + i].x), ASTC_F2U(w.ep[EP_EI1_1 + i].y), ASTC_F2U(w.ep[EP_EI1_1 + i].z), ASTC_F2U(w.ep[EP_EI1_1 + i].w));
	}
	for (int t = 0; t < w.T; t++) {
		TRACE("eiw[%d] %08x wes %08x\n", t, ASTC_F2U(w.eiw[0][t]), ASTC_F2U(w.eis[0][t]));
	}
	for (unsigned int i = 0; i < max_decimation_modes; i++) {
		if ((bsd.dec_modes[i].refprec_1plane & refmask) == 0) continue;
		for (int k = 0; k < bsd.dec_modes[i].weight_count; k++) {
			TRACE("dwi[%u][%d] %08x\n", i, k, ASTC_F2U(w.dwi[bsd.dec_modes[i].dwi_offset + k]));
		}
	}
	TRACE_F("min_wt_cutoff", min_wt_cutoff);

	compute_angular_endpoints(w, only_always, 1, (unsigned int)max_weight_quant);
	for (unsigned int i = 0; i < max_decimation_modes; i++) {
		if ((bsd.dec_modes[i].refprec_1plane & refmask) == 0) continue;
		for (int k = 0; k < 16; k++) {
			TRACE("lowhigh[%u][%d] %08x\n", i, k, ASTC_F2U(w.lowhigh[(i * 2) * 16 + k]));
		}
	}

	unsigned int max_block_modes = only_always ? bsd.block_mode_count_1plane_always : bsd.block_mode_count_1plane_selected;
	quantize_and_score_modes(w, 0, max_block_modes, 1, partition_count, max_weight_quant, min_wt_cutoff, min_wt_cutoff);

	for (unsigned int i = 0; i < max_block_modes; i++) {
		TRACE("qwt_err[%u] %08x\n", i, ASTC_F2U(w.mode_err[i]));
	}
	unsigned int candidate_count = compute_ideal_endpoint_formats(w, pi, EP_EI1_0, EP_EI1_1, 1, 0, max_block_modes);
	for (unsigned int i = 0; i < candidate_count; i++) {
		const Candidate* cc = reinterpret_cast<const Candidate*>(w.cand) + i;
		TRACE("cand[%u] mode=%u ql=%u qlm=%u fmt=%u %u %u %u\n", i, cc->block_mode, cc->quant_level, cc->quant_level_mod, cc->formats[0], cc->formats[1], cc->formats[2], cc->formats[3]);
	}
	for (int k = w.lane; k < 4; k += ASTC_WARP) {
		w.ep[EP_BASE_0 + k] = w.ep[EP_EI1_0 + k];
		w.ep[EP_BASE_1 + k] = w.ep[EP_EI1_1 + k];
	}
	wsync();
	return refine_candidates(w, pi, false, partition_count, partition_index, -1, candidate_count, min_wt_cutoff, min_wt_cutoff, tune_errorval_threshold, scb);
}

// compress_symbolic_block_for_partition_2planes :715-1044
ASTC_COOP float compress_symbolic_block_for_partition_2planes(WCtx& w, float tune_errorval_threshold, unsigned int plane2_component, ScbHdr& scb, int quant_limit) {
	const DevBsd& bsd = *w.bsd;
	int max_weight_quant = mini((int)QUANT_32, quant_limit);
	compute_ideal_colors_and_weights_2planes(w, plane2_component);
	uint16_t refmask = (uint16_t)((1u << (max_weight_quant + 1)) - 1);
	for (unsigned int i = 0; i < bsd.decimation_mode_count_selected; i++) {
		if ((bsd.dec_modes[i].refprec_2planes & refmask) == 0) {
			continue;
		}
		DecView di = dec_view(bsd, i);
		compute_ideal_weights_for_decimation(w, di, 2, w.dwi + di.dm->dwi_offset, di.W);
	}
	f4 a0 = w.ep[EP_EI1_0], a1 = w.ep[EP_EI1_1], b0 = w.ep[EP_EI2_0], b1 = w.ep[EP_EI2_1];
	f4 min_ep1 = mk4(min_ep_cutoff(a0.x, a1.x, 10.0f), min_ep_cutoff(a0.y, a1.y, 10.0f), min_ep_cutoff(a0.z, a1.z, 10.0f), min_ep_cutoff(a0.w, a1.w, 10.0f));
	f4 min_ep2 = mk4(min_ep_cutoff(b0.x, b1.x, 10.0f), min_ep_cutoff(b0.y, b1.y, 10.0f), min_ep_cutoff(b0.z, b1.z, 10.0f), min_ep_cutoff(b0.w, b1.w, 10.0f));
	f4 m1 = min_ep1;
	set_lane(m1, (int)plane2_component, ERROR_CALC_DEFAULT);
	float min_wt_cutoff1 = hmin_s(m1);
	f4 m2 = splat4(ERROR_CALC_DEFAULT);
	set_lane(m2, (int)plane2_component, lane(min_ep2, (int)plane2_component));
	float min_wt_cutoff2 = hmin_s(m2);

	compute_angular_endpoints(w, false, 2, (unsigned int)max_weight_quant);

	unsigned int start_2plane = bsd.block_mode_count_1plane_selected;
	unsigned int end_2plane = bsd.block_mode_count_1plane_2plane_selected;
	quantize_and_score_modes(w, start_2plane, end_2plane, 2, 1, max_weight_quant, min_wt_cutoff1, min_wt_cutoff2);

	// merge_endpoints :37-66
	f4 epm0 = a0, epm1 = a1;
	set_lane(epm0, (int)plane2_component, lane(b0, (int)plane2_component));
	set_lane(epm1, (int)plane2_component, lane(b1, (int)plane2_component));
	if (w.lane == 0) {
		w.ep[EP_BASE_0] = epm0;
		w.ep[EP_BASE_1] = epm1;
	}
	wsync();
	PartView pi = part_view_packed(bsd, 1, 0);
	unsigned int candidate_count = compute_ideal_endpoint_formats(w, pi, EP_BASE_0, EP_BASE_1, 2, start_2plane, end_2plane);
	return refine_candidates(w, pi, true, 1, 0, (int)plane2_component, candidate_count, min_wt_cutoff1, min_wt_cutoff2, tune_errorval_threshold, scb);
}

// prepare_block_statistics :1047-1159 - 15 ordered chains over the texels
ASTC_COOP float prepare_block_statistics(WCtx& w) {
	int T = w.T;
	float weight = hadd_s(w.bi.channel_weight) / 4.0f;
	// chains: 0 rs, 1 gs, 2 bs, 3 as, 4 rr, 5 gg, 6 bb, 7 aa, 8 rg, 9 rb, 10 ra, 11 gb, 12 ga, 13 ba, 14 weight_sum
	for (int ch = w.lane; ch < 15; ch += ASTC_WARP) {
		float s = 0.0f;
		for (int i = 0; i < T; i++) {
			float r = w.blk[0][i], g = w.blk[1][i], b = w.blk[2][i], a = w.blk[3][i];
			float rw = r * weight, gw = g * weight, bw = b * weight, aw = a * weight;
			float term;
			switch (ch) {
			case 0: term = rw; break;
			case 1: term = gw; break;
			case 2: term = bw; break;
			case 3: term = aw; break;
			case 4: term = r * rw; break;
			case 5: term = g * gw; break;
			case 6: term = b * bw; break;
			case 7: term = a * aw; break;
			case 8: term = g * rw; break;
			case 9: term = b * rw; break;
			case 10: term = a * rw; break;
			case 11: term = b * gw; break;
			case 12: term = a * gw; break;
			case 13: term = a * bw; break;
			default: term = weight; break;
			}
			s += term;
		}
		w.tmpf[ch] = s;
	}
	wsync();
	const float* t = w.tmpf;
	float rs = t[0], gs = t[1], bs = t[2], as = t[3];
	float rr_var = t[4], gg_var = t[5], bb_var = t[6], aa_var = t[7];
	float rg_cov = t[8], rb_cov = t[9], ra_cov = t[10], gb_cov = t[11], ga_cov = t[12], ba_cov = t[13];
	float weight_sum = t[14];
	wsync();
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

ASTC_FN void constant_color_u16(const WCtx& w, ScbHdr& scb) {
	scb.block_type = SYM_BTYPE_CONST_U16;
	f4 c = vclamp4(0.0f, 1.0f, w.bi.origin_texel) * 65535.0f;
	scb.constant_color[0] = f2i_rtn(c.x);
	scb.constant_color[1] = f2i_rtn(c.y);
	scb.constant_color[2] = f2i_rtn(c.z);
	scb.constant_color[3] = f2i_rtn(c.w);
}

// compress_block :1162-1455. The block is already loaded (w.blk, w.bi). Writes 16 bytes to out.
ASTC_COOP void compress_block(WCtx& w, uint8_t* out) {
	const DevBsd& bsd = *w.bsd;
	const DevConfig& config = *w.cfg;
	int decode_mode = config.profile;
	ScbHdr scb;
	scb.block_type = SYM_BTYPE_ERROR;
	scb.partition_count = 0;
	scb.color_formats_matched = 0;
	scb.plane2_component = -1;
	scb.block_mode = 0;
	scb.partition_index = 0;
	scb.color_formats[0] = scb.color_formats[1] = scb.color_formats[2] = scb.color_formats[3] = 0;
	scb.quant_mode = 0;
	scb.errorval = 0.0f;
	scb.constant_color[0] = scb.constant_color[1] = scb.constant_color[2] = scb.constant_color[3] = 0;

	const BlkInfo& bi = w.bi;
	if (bi.data_min.x == bi.data_max.x && bi.data_min.y == bi.data_max.y && bi.data_min.z == bi.data_max.z && bi.data_min.w == bi.data_max.w) {
		scb.partition_count = 0;
		if (decode_mode == PRF_HDR || decode_mode == PRF_HDR_RGB_LDR_A) {
			scb.block_type = SYM_BTYPE_CONST_F16;
			scb.constant_color[0] = float_to_sf16(bi.origin_texel.x);
			scb.constant_color[1] = float_to_sf16(bi.origin_texel.y);
			scb.constant_color[2] = float_to_sf16(bi.origin_texel.z);
			scb.constant_color[3] = float_to_sf16(bi.origin_texel.w);
		} else {
			constant_color_u16(w, scb);
		}
		if (w.lane == 0) {
			symbolic_to_physical(w, scb, out);
		}
		wsync();
		return;
	}

	bool block_is_l = is_luminance(w);
	float block_is_l_scale = block_is_l ? 1.0f / 1.5f : 1.0f;
	bool block_is_la = is_luminancealpha(w);
	float block_is_la_scale = block_is_la ? 1.0f / 1.05f : 1.0f;
	int max_partitions = (int)config.tune_partition_count_limit;

	TRACE("blk min %08x %08x %08x %08x max %08x %08x %08x %08x mean %08x %08x %08x %08x gray=%d\n", ASTC_F2U(bi.data_min.x), ASTC_F2U(bi.data_min.y), ASTC_F2U(bi.data_min.z), ASTC_F2U(bi.data_min.w),
	      ASTC_F2U(bi.data_max.x), ASTC_F2U(bi.data_max.y), ASTC_F2U(bi.data_max.z), ASTC_F2U(bi.data_max.w), ASTC_F2U(bi.data_mean.x), ASTC_F2U(bi.data_mean.y), ASTC_F2U(bi.data_mean.z), ASTC_F2U(bi.data_mean.w), (int)bi.grayscale);
	float error_weight_sum = hadd_s(bi.channel_weight) * bsd.texel_count;
	float error_threshold = config.tune_db_limit * error_weight_sum * block_is_l_scale * block_is_la_scale;
	TRACE_F("error_threshold", error_threshold);

	scb.errorval = ERROR_CALC_DEFAULT;
	scb.block_type = SYM_BTYPE_ERROR;
	float best_errorvals_for_pcount[4] = {ERROR_CALC_DEFAULT, ERROR_CALC_DEFAULT, ERROR_CALC_DEFAULT, ERROR_CALC_DEFAULT};
	float exit_thresholds_for_pcount[4] = {0.0f, config.tune_2partition_early_out_limit_factor, config.tune_3partition_early_out_limit_factor, 0.0f};
	float errorval_mult[2] = {1.0f / config.tune_mse_overshoot, 1.0f};
	const float errorval_overshoot = 1.0f / config.tune_mse_overshoot;
	int start_trial = 1;
	if (config.tune_search_mode0_enable >= 0.85f) {
		start_trial = 0;
	}
	int quant_limit = QUANT_32;
	bool done = false;
	for (int i = start_trial; i < 2 && !done; i++) {
		float errorval = compress_symbolic_block_for_partition_1plane(w, i == 0, error_threshold * errorval_mult[i] * errorval_overshoot, 1, 0, scb, QUANT_32);
		if (scb.block_type != SYM_BTYPE_ERROR) {
			quant_limit = bsd.block_modes[bsd.block_mode_packed_index[scb.block_mode]].quant_mode;
		}
		best_errorvals_for_pcount[0] = minf(best_errorvals_for_pcount[0], errorval);
		if (errorval < (error_threshold * errorval_mult[i])) {
			done = true;
		}
	}
	if (!done) {
		float lowest_correl = prepare_block_statistics(w);
		bool block_skip_two_plane = lowest_correl > config.tune_2plane_early_out_limit_correlation;
		for (int i = 3; i >= 0 && !done; i--) {
			if (block_skip_two_plane) {
				continue;
			}
			if (bi.grayscale && i != 3) {
				continue;
			}
			if (is_constant_channel(w, i)) {
				continue;
			}
			float errorval = compress_symbolic_block_for_partition_2planes(w, error_threshold * errorval_overshoot, (unsigned int)i, scb, quant_limit);
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
		unsigned int requested_indices = config.tune_partition_index_limit[partition_count - 2];
		unsigned int requested_trials = config.tune_partitioning_candidate_limit[partition_count - 2];
		requested_trials = requested_trials < requested_indices ? requested_trials : requested_indices;
		unsigned int actual_trials = find_best_partition_candidates(w, (unsigned int)partition_count, requested_indices, partition_indices, requested_trials);
		float best_error_in_prev = best_errorvals_for_pcount[partition_count - 2];
		for (unsigned int i = 0; i < actual_trials && !done; i++) {
			float errorval = compress_symbolic_block_for_partition_1plane(w, false, error_threshold * errorval_overshoot, (unsigned int)partition_count,
			                                                              partition_indices[i], scb, quant_limit);
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
		constant_color_u16(w, scb);
	}
	if (w.lane == 0) {
		symbolic_to_physical(w, scb, out);
	}
	wsync();
}

// Carve the per-warp arena (layout computed by the host, see astc_host_tables.cpp: plan_arena)
ASTC_FN void init_wctx(WCtx& w, int lane, const DevBsd* bsd, const DevConfig* cfg, uint8_t* arena) {
	w.lane = lane;
	w.bsd = bsd;
	w.cfg = cfg;
	int T = bsd->texel_count;
	w.T = T;
	int Tp = (T + 3) & ~3;
	float* blk = reinterpret_cast<float*>(arena + bsd->off_blk);
	w.blk[0] = blk;
	w.blk[1] = blk + Tp;
	w.blk[2] = blk + 2 * Tp;
	w.blk[3] = blk + 3 * Tp;
	float* ei = reinterpret_cast<float*>(arena + bsd->off_ei);
	w.eiw[0] = ei;
	w.eis[0] = ei + Tp;
	w.eiw[1] = ei + 2 * Tp;
	w.eis[1] = ei + 3 * Tp;
	w.ei_const_wes[0] = w.ei_const_wes[1] = false;
	w.ep = reinterpret_cast<f4*>(arena + bsd->off_ep);
	w.dwi = reinterpret_cast<float*>(arena + bsd->off_dwi);
	w.lowhigh = reinterpret_cast<float*>(arena + bsd->off_lowhigh);
	w.mode_err = reinterpret_cast<float*>(arena + bsd->off_mode_err);
	uint8_t* scbp = arena + bsd->off_scb;
	w.best_weights = scbp;
	w.best_colors = scbp + 64;
	w.work_weights = scbp + 96;
	w.work_colors = scbp + 160;
	w.mod_colors = scbp + 192;
	uint8_t* sc = arena + bsd->off_scratch;
	w.tmpf = reinterpret_cast<float*>(sc);
	w.cand = sc + 512;
	w.su = sc + 512 + 64;
}
