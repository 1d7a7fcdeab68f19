// ORACLE (test infrastructure, not product code): CPU restatement of the reference's per-block
// compressor. See astc_codec.h. Scalar C++, IEEE fp32, no FMA contraction (-ffp-contract=off).
#include "astc_codec.h"
#include <vector>
#include <algorithm>
#include <cstddef>

#include <cstdlib>
#include <cstring>
#include <cmath>

namespace ao {

static const float ERROR_CALC_DEFAULT = 1e30f;
static const int TUNE_MAX_ANGULAR_QUANT = 7;
static const int TUNE_MAX_TRIAL_CANDIDATES = 8;
static const int TUNE_MAX_PARTITIONING_CANDIDATES = 8;

// =============================================================================================
// Config (astcenc_entry.cpp:65-135 presets, :504-723 astcenc_config_init, :434-501 validate_config)
// =============================================================================================
struct Preset {
	float quality;
	unsigned int partition_count_limit, p2_index_limit, p3_index_limit, p4_index_limit, block_mode_limit,
	             refinement_limit, candidate_limit, p2_cand_limit, p3_cand_limit, p4_cand_limit;
	float db_limit_a_base, db_limit_b_base, mse_overshoot, p2_early_out, p3_early_out, plane2_correlation, search_mode0;
};

static const Preset PRESETS_HIGH[6] = {
	{0.0f, 2, 10, 6, 4, 43, 2, 2, 2, 2, 2, 85.2f, 63.2f, 3.5f, 1.00f, 1.00f, 0.85f, 0.0f},
	{10.0f, 3, 18, 10, 8, 55, 3, 3, 2, 2, 2, 85.2f, 63.2f, 3.5f, 1.00f, 1.00f, 0.90f, 0.0f},
	{60.0f, 4, 34, 28, 16, 77, 3, 3, 2, 2, 2, 95.0f, 70.0f, 2.5f, 1.10f, 1.05f, 0.95f, 0.0f},
	{98.0f, 4, 82, 60, 30, 94, 4, 4, 3, 2, 2, 105.0f, 77.0f, 10.0f, 1.35f, 1.15f, 0.97f, 0.0f},
	{99.0f, 4, 256, 128, 64, 98, 4, 6, 8, 6, 4, 200.0f, 200.0f, 10.0f, 1.60f, 1.40f, 0.98f, 0.0f},
	{100.0f, 4, 512, 512, 512, 100, 4, 8, 8, 8, 8, 200.0f, 200.0f, 10.0f, 2.00f, 2.00f, 0.99f, 0.0f}};
static const Preset PRESETS_MID[6] = {
	{0.0f, 2, 10, 6, 4, 43, 2, 2, 2, 2, 2, 85.2f, 63.2f, 3.5f, 1.00f, 1.00f, 0.80f, 1.0f},
	{10.0f, 3, 18, 12, 10, 55, 3, 3, 2, 2, 2, 85.2f, 63.2f, 3.5f, 1.00f, 1.00f, 0.85f, 1.0f},
	{60.0f, 3, 34, 28, 16, 77, 3, 3, 2, 2, 2, 95.0f, 70.0f, 3.0f, 1.10f, 1.05f, 0.90f, 1.0f},
	{98.0f, 4, 82, 60, 30, 94, 4, 4, 3, 2, 2, 105.0f, 77.0f, 10.0f, 1.40f, 1.20f, 0.95f, 0.0f},
	{99.0f, 4, 256, 128, 64, 98, 4, 6, 8, 6, 3, 200.0f, 200.0f, 10.0f, 1.60f, 1.40f, 0.98f, 0.0f},
	{100.0f, 4, 256, 256, 256, 100, 4, 8, 8, 8, 8, 200.0f, 200.0f, 10.0f, 2.00f, 2.00f, 0.99f, 0.0f}};
static const Preset PRESETS_LOW[6] = {
	{0.0f, 2, 10, 6, 4, 40, 2, 2, 2, 2, 2, 85.0f, 63.0f, 3.5f, 1.00f, 1.00f, 0.80f, 1.0f},
	{10.0f, 2, 18, 12, 10, 55, 3, 3, 2, 2, 2, 85.0f, 63.0f, 3.5f, 1.00f, 1.00f, 0.85f, 1.0f},
	{60.0f, 3, 34, 28, 16, 77, 3, 3, 2, 2, 2, 95.0f, 70.0f, 3.5f, 1.10f, 1.05f, 0.90f, 1.0f},
	{98.0f, 4, 82, 60, 30, 93, 4, 4, 3, 2, 2, 105.0f, 77.0f, 10.0f, 1.30f, 1.20f, 0.97f, 1.0f},
	{99.0f, 4, 256, 128, 64, 98, 4, 6, 8, 5, 2, 200.0f, 200.0f, 10.0f, 1.60f, 1.40f, 0.98f, 1.0f},
	{100.0f, 4, 256, 256, 256, 100, 4, 8, 8, 8, 8, 200.0f, 200.0f, 10.0f, 2.00f, 2.00f, 0.99f, 1.0f}};

// approximate exp2 / log2 / pow used for the dB limit (astcenc_vecmathlib.h:402-483)
static float approx_exp2(float x) {
	x = vclampf(-126.99999f, 129.0f, x);
	int ipart = f2i(x - 0.5f);
	float fpart = x - static_cast<float>(ipart);
	float iexp = u_as_f((uint32_t)((ipart + 127) << 23));
	float p = 1.8775767e-3f;
	p = (p * fpart) + 8.9893397e-3f;
	p = (p * fpart) + 5.5826318e-2f;
	p = (p * fpart) + 2.4015361e-1f;
	p = (p * fpart) + 6.9315308e-1f;
	p = (p * fpart) + 9.9999994e-1f;
	return iexp * p;
}

static float approx_log2(float x) {
	uint32_t i = f_as_u(x);
	float e = static_cast<float>((int)((i & 0x7F800000u) >> 23) - 127);
	float m = u_as_f((i & 0x007FFFFFu) | 0x3F800000u);
	float p = 0.0596515482674574969533f;
	p = (p * m) + -0.465725644288844778798f;
	p = (p * m) + 1.48116647521213171641f;
	p = (p * m) + -2.52074962577807006663f;
	p = (p * m) + 2.8882704548164776201f;
	p = p * (m - 1.0f);
	return p + e;
}

static float approx_pow(float x, float y) {
	if (y == 0.0f) {
		return 1.0f;
	}
	return approx_exp2(approx_log2(x) * y);
}

int config_init(int profile, unsigned int block_x, unsigned int block_y, float quality, unsigned int flags, Config& cfg, unsigned int block_z) {
	memset(&cfg, 0, sizeof(cfg));
	block_z = block_z < 1 ? 1 : block_z;      // (Z == 0 is accepted for 2D sizes, :527)
	bool legal = block_z <= 1 ? is_legal_2d_block_size(block_x, block_y) : is_legal_3d_block_size(block_x, block_y, block_z);
	if (!legal) {
		return 4;   // ASTCENC_ERR_BAD_BLOCK_SIZE
	}
	cfg.block_x = block_x;
	cfg.block_y = block_y;
	cfg.block_z = block_z;
	float texels = static_cast<float>(block_x * block_y * block_z);
	float ltexels = logf(texels) / logf(10.0f);
	if (quality < 0.0f || quality > 100.0f) {
		return 6;   // ASTCENC_ERR_BAD_QUALITY
	}
	unsigned int texels_int = block_x * block_y * block_z;
	const Preset* presets = texels_int < 25 ? PRESETS_HIGH : texels_int < 64 ? PRESETS_MID : PRESETS_LOW;
	size_t end;
	for (end = 0; end < 6; end++) {
		if (presets[end].quality >= quality) {
			break;
		}
	}
	size_t start = end == 0 ? 0 : end - 1;
	if (start == end) {
		const Preset& p = presets[start];
		cfg.tune_partition_count_limit = p.partition_count_limit;
		cfg.tune_2partition_index_limit = p.p2_index_limit;
		cfg.tune_3partition_index_limit = p.p3_index_limit;
		cfg.tune_4partition_index_limit = p.p4_index_limit;
		cfg.tune_block_mode_limit = p.block_mode_limit;
		cfg.tune_refinement_limit = p.refinement_limit;
		cfg.tune_candidate_limit = p.candidate_limit;
		cfg.tune_2partitioning_candidate_limit = p.p2_cand_limit;
		cfg.tune_3partitioning_candidate_limit = p.p3_cand_limit;
		cfg.tune_4partitioning_candidate_limit = p.p4_cand_limit;
		cfg.tune_db_limit = maxf(p.db_limit_a_base - 35 * ltexels, p.db_limit_b_base - 19 * ltexels);
		cfg.tune_mse_overshoot = p.mse_overshoot;
		cfg.tune_2partition_early_out_limit_factor = p.p2_early_out;
		cfg.tune_3partition_early_out_limit_factor = p.p3_early_out;
		cfg.tune_2plane_early_out_limit_correlation = p.plane2_correlation;
		cfg.tune_search_mode0_enable = p.search_mode0;
	} else {
		const Preset& a = presets[start];
		const Preset& b = presets[end];
		float wt_range = b.quality - a.quality;
		float wa = (b.quality - quality) / wt_range;
		float wb = (quality - a.quality) / wt_range;
#define LERP(f) ((a.f * wa) + (b.f * wb))
#define LERPI(f) f2i_rtn((static_cast<float>(a.f) * wa) + (static_cast<float>(b.f) * wb))
		cfg.tune_partition_count_limit = (unsigned int)LERPI(partition_count_limit);
		cfg.tune_2partition_index_limit = (unsigned int)LERPI(p2_index_limit);
		cfg.tune_3partition_index_limit = (unsigned int)LERPI(p3_index_limit);
		cfg.tune_4partition_index_limit = (unsigned int)LERPI(p4_index_limit);
		cfg.tune_block_mode_limit = (unsigned int)LERPI(block_mode_limit);
		cfg.tune_refinement_limit = (unsigned int)LERPI(refinement_limit);
		cfg.tune_candidate_limit = (unsigned int)LERPI(candidate_limit);
		cfg.tune_2partitioning_candidate_limit = (unsigned int)LERPI(p2_cand_limit);
		cfg.tune_3partitioning_candidate_limit = (unsigned int)LERPI(p3_cand_limit);
		cfg.tune_4partitioning_candidate_limit = (unsigned int)LERPI(p4_cand_limit);
		cfg.tune_db_limit = maxf(LERP(db_limit_a_base) - 35 * ltexels, LERP(db_limit_b_base) - 19 * ltexels);
		cfg.tune_mse_overshoot = LERP(mse_overshoot);
		cfg.tune_2partition_early_out_limit_factor = LERP(p2_early_out);
		cfg.tune_3partition_early_out_limit_factor = LERP(p3_early_out);
		cfg.tune_2plane_early_out_limit_correlation = LERP(plane2_correlation);
		cfg.tune_search_mode0_enable = LERP(search_mode0);
#undef LERP
#undef LERPI
	}
	cfg.cw_r_weight = 1.0f;
	cfg.cw_g_weight = 1.0f;
	cfg.cw_b_weight = 1.0f;
	cfg.cw_a_weight = 1.0f;
	cfg.a_scale_radius = 0;
	cfg.rgbm_m_scale = 0.0f;
	cfg.profile = profile;
	switch (profile) {
	case PRF_LDR:
	case PRF_LDR_SRGB:
		break;
	case PRF_HDR_RGB_LDR_A:
	case PRF_HDR:
		cfg.tune_db_limit = 999.0f;
		cfg.tune_search_mode0_enable = 0.0f;
		break;
	default:
		return 5;   // ASTCENC_ERR_BAD_PROFILE
	}
	const unsigned int all_flags = FLG_MAP_NORMAL | FLG_MAP_RGBM | FLG_USE_ALPHA_WEIGHT | FLG_USE_PERCEPTUAL |
	                               FLG_USE_DECODE_UNORM8 | FLG_DECOMPRESS_ONLY | FLG_SELF_DECOMPRESS_ONLY;
	if (flags & ~all_flags) {
		return 8;   // ASTCENC_ERR_BAD_FLAGS
	}
	if ((flags & FLG_MAP_NORMAL) && (flags & FLG_MAP_RGBM)) {
		return 8;
	}
	if ((flags & FLG_USE_DECODE_UNORM8) && (profile == PRF_HDR || profile == PRF_HDR_RGB_LDR_A)) {
		return 11;   // ASTCENC_ERR_BAD_DECODE_MODE
	}
	if (flags & FLG_MAP_NORMAL) {
		cfg.tune_partition_count_limit = cfg.tune_partition_count_limit + 1u < 4u ? cfg.tune_partition_count_limit + 1u : 4u;
		cfg.cw_g_weight = 0.0f;
		cfg.cw_b_weight = 0.0f;
		cfg.tune_2partition_early_out_limit_factor *= 1.5f;
		cfg.tune_3partition_early_out_limit_factor *= 1.5f;
		cfg.tune_2plane_early_out_limit_correlation = 0.99f;
		cfg.tune_db_limit *= 1.03f;
	} else if (flags & FLG_MAP_RGBM) {
		cfg.rgbm_m_scale = 5.0f;
		cfg.cw_a_weight = 2.0f * cfg.rgbm_m_scale;
	} else if (flags & FLG_USE_PERCEPTUAL) {
		cfg.cw_r_weight = 0.30f * 2.25f;
		cfg.cw_g_weight = 0.59f * 2.25f;
		cfg.cw_b_weight = 0.11f * 2.25f;
	}
	cfg.flags = flags;
	return 0;
}

static unsigned int clampu(unsigned int v, unsigned int mn, unsigned int mx) {
	if (v > mx) return mx;
	if (v > mn) return v;
	return mn;
}

int config_finalize(Config& c) {
	c.rgbm_m_scale = maxf(c.rgbm_m_scale, 1.0f);
	c.tune_partition_count_limit = clampu(c.tune_partition_count_limit, 1u, 4u);
	c.tune_2partition_index_limit = clampu(c.tune_2partition_index_limit, 1u, 1024u);
	c.tune_3partition_index_limit = clampu(c.tune_3partition_index_limit, 1u, 1024u);
	c.tune_4partition_index_limit = clampu(c.tune_4partition_index_limit, 1u, 1024u);
	c.tune_block_mode_limit = clampu(c.tune_block_mode_limit, 1u, 100u);
	c.tune_refinement_limit = c.tune_refinement_limit > 1u ? c.tune_refinement_limit : 1u;
	c.tune_candidate_limit = clampu(c.tune_candidate_limit, 1u, (unsigned int)TUNE_MAX_TRIAL_CANDIDATES);
	c.tune_2partitioning_candidate_limit = clampu(c.tune_2partitioning_candidate_limit, 1u, (unsigned int)TUNE_MAX_PARTITIONING_CANDIDATES);
	c.tune_3partitioning_candidate_limit = clampu(c.tune_3partitioning_candidate_limit, 1u, (unsigned int)TUNE_MAX_PARTITIONING_CANDIDATES);
	c.tune_4partitioning_candidate_limit = clampu(c.tune_4partitioning_candidate_limit, 1u, (unsigned int)TUNE_MAX_PARTITIONING_CANDIDATES);
	c.tune_db_limit = maxf(c.tune_db_limit, 0.0f);
	c.tune_mse_overshoot = maxf(c.tune_mse_overshoot, 1.0f);
	c.tune_2partition_early_out_limit_factor = maxf(c.tune_2partition_early_out_limit_factor, 0.0f);
	c.tune_3partition_early_out_limit_factor = maxf(c.tune_3partition_early_out_limit_factor, 0.0f);
	c.tune_2plane_early_out_limit_correlation = maxf(c.tune_2plane_early_out_limit_correlation, 0.0f);
	float max_weight = maxf(maxf(c.cw_r_weight, c.cw_g_weight), maxf(c.cw_b_weight, c.cw_a_weight));
	if (max_weight > 0.0f) {
		max_weight /= 1000.0f;
		c.cw_r_weight = maxf(c.cw_r_weight, max_weight);
		c.cw_g_weight = maxf(c.cw_g_weight, max_weight);
		c.cw_b_weight = maxf(c.cw_b_weight, max_weight);
		c.cw_a_weight = maxf(c.cw_a_weight, max_weight);
	} else {
		return 3;   // ASTCENC_ERR_BAD_PARAM
	}
	// astcenc_context_alloc :814-821: dB limit -> per-texel squared error
	if (c.profile == PRF_LDR || c.profile == PRF_LDR_SRGB) {
		c.tune_db_limit = approx_pow(0.1f, c.tune_db_limit * 0.1f) * 65535.0f * 65535.0f;
	} else {
		c.tune_db_limit = 0.0f;
	}
	return 0;
}

// =============================================================================================
// Working buffers (compression_working_buffers, astcenc_internal.h:953-1040)
// =============================================================================================
struct Endpoints {
	unsigned int partition_count;
	f4 endpt0[4];
	f4 endpt1[4];
};

struct EndpointsAndWeights {
	bool is_constant_weight_error_scale;
	Endpoints ep;
	float weights[MAX_TEXELS];
	float weight_error_scale[MAX_TEXELS];
};

struct WorkBuf {
	EndpointsAndWeights ei1, ei2;
	float dec_weights_ideal[MAX_DECIMATION_MODES * MAX_WEIGHTS];
	uint8_t dec_weights_uquant[MAX_BLOCK_MODES * MAX_WEIGHTS];
	float errors_of_best_combination[MAX_BLOCK_MODES];
	uint8_t best_quant_levels[MAX_BLOCK_MODES];
	uint8_t best_quant_levels_mod[MAX_BLOCK_MODES];
	uint8_t best_ep_formats[MAX_BLOCK_MODES][4];
	int8_t qwt_bitcounts[MAX_BLOCK_MODES];
	float qwt_errors[MAX_BLOCK_MODES];
	float weight_low_value1[MAX_BLOCK_MODES];
	float weight_high_value1[MAX_BLOCK_MODES];
	float weight_low_values1[MAX_DECIMATION_MODES][TUNE_MAX_ANGULAR_QUANT + 1];
	float weight_high_values1[MAX_DECIMATION_MODES][TUNE_MAX_ANGULAR_QUANT + 1];
	float weight_low_value2[MAX_BLOCK_MODES];
	float weight_high_value2[MAX_BLOCK_MODES];
	float weight_low_values2[MAX_DECIMATION_MODES][TUNE_MAX_ANGULAR_QUANT + 1];
	float weight_high_values2[MAX_DECIMATION_MODES][TUNE_MAX_ANGULAR_QUANT + 1];
};

Context* context_create(const Config& cfg) {
	Context* ctx = new Context;
	ctx->config = cfg;
	bool can_omit = (cfg.flags & FLG_SELF_DECOMPRESS_ONLY) != 0;
	ctx->bsd = build_block_size_tables(cfg.block_x, cfg.block_y, cfg.block_z < 1 ? 1 : cfg.block_z, can_omit, cfg.tune_partition_count_limit,
	                                   static_cast<float>(cfg.tune_block_mode_limit) / 100.0f);
	ctx->work = new WorkBuf;
	return ctx;
}

void context_destroy(Context* ctx) {
	if (!ctx) {
		return;
	}
	free_block_size_tables(ctx->bsd);
	delete static_cast<WorkBuf*>(ctx->work);
	delete ctx;
}

static inline f4 texel4(const ImageBlock& b, unsigned int i) { return mk4(b.data_r[i], b.data_g[i], b.data_b[i], b.data_a[i]); }
static inline f4 texel3(const ImageBlock& b, unsigned int i) { return mk4(b.data_r[i], b.data_g[i], b.data_b[i], 0.0f); }
static inline float default_alpha(const ImageBlock& b) { return b.alpha_lns0 ? static_cast<float>(0x7800) : static_cast<float>(0xFFFF); }
static inline bool is_constant_channel(const ImageBlock& b, int ch) { return lane(b.data_min, ch) == lane(b.data_max, ch); }
static inline bool is_luminance(const ImageBlock& b) {
	float da = default_alpha(b);
	bool alpha1 = (b.data_min.w == da) && (b.data_max.w == da);
	return b.grayscale && alpha1;
}
static inline bool is_luminancealpha(const ImageBlock& b) {
	float da = default_alpha(b);
	bool alpha1 = (b.data_min.w == da) && (b.data_max.w == da);
	return b.grayscale && !alpha1;
}

// =============================================================================================
// Block load (astcenc_image.cpp:162-342)
// =============================================================================================
void load_block(const Context& ctx, const void* data, int data_type, unsigned int dim_x, unsigned int dim_y,
                unsigned int pos_x, unsigned int pos_y, const int swz[4], ImageBlock& blk, unsigned int dim_z, unsigned int pos_z) {
	const BlockSizeTables& bsd = *ctx.bsd;
	int profile = ctx.config.profile;
	bool needs_swz = swz[0] != 0 || swz[1] != 1 || swz[2] != 2 || swz[3] != 3;
	bool needs_hdr = profile == PRF_HDR || profile == PRF_HDR_RGB_LDR_A;
	bool fast = !needs_swz && !needs_hdr && data_type == 0 && bsd.dim_z == 1;      // (astcenc_entry.cpp:946-947)

	blk.texel_count = bsd.texel_count;
	blk.decode_unorm8 = (ctx.config.flags & FLG_USE_DECODE_UNORM8) != 0;
	blk.channel_weight = mk4(ctx.config.cw_r_weight, ctx.config.cw_g_weight, ctx.config.cw_b_weight, ctx.config.cw_a_weight);

	f4 dmin = splat4(1e38f), dmax = splat4(-1e38f), dmean = splat4(0.0f);
	bool gray = true;
	unsigned int idx = 0;
	if (fast) {
		// load_image_block_fast_ldr :278-342
		const uint8_t* plane = static_cast<const uint8_t*>(data);
		for (unsigned int y = pos_y; y < pos_y + bsd.dim_y; y++) {
			unsigned int yi = y < dim_y - 1 ? y : dim_y - 1;
			for (unsigned int x = pos_x; x < pos_x + bsd.dim_x; x++) {
				unsigned int xi = x < dim_x - 1 ? x : dim_x - 1;
				const uint8_t* p = plane + (4 * (size_t)dim_x * yi) + (4 * xi);
				f4 v = mk4(static_cast<float>(p[0]), static_cast<float>(p[1]), static_cast<float>(p[2]), static_cast<float>(p[3])) * (65535.0f / 255.0f);
				dmin = min4(dmin, v);
				dmean = dmean + v;
				dmax = max4(dmax, v);
				gray = gray && (v.x == v.y) && (v.x == v.z);
				blk.data_r[idx] = v.x;
				blk.data_g[idx] = v.y;
				blk.data_b[idx] = v.z;
				blk.data_a[idx] = v.w;
				idx++;
			}
		}
		blk.origin_texel = texel4(blk, 0) / 65535.0f;
		blk.rgb_lns0 = 0;
		blk.alpha_lns0 = 0;
		blk.data_min = dmin;
		blk.data_mean = dmean / static_cast<float>(bsd.texel_count);
		blk.data_max = dmax;
		blk.grayscale = gray;
		return;
	}

	// load_image_block :162-275
	float mean_scale = 1.0f / static_cast<float>(bsd.texel_count);
	uint8_t rgb_lns = needs_hdr ? 1 : 0;
	uint8_t a_lns = profile == PRF_HDR ? 1 : 0;
	for (unsigned int z = 0; z < bsd.dim_z; z++) {
	unsigned int zi = pos_z + z < dim_z - 1 ? pos_z + z : dim_z - 1;
	for (unsigned int y = 0; y < bsd.dim_y; y++) {
		unsigned int yi = pos_y + y < dim_y - 1 ? pos_y + y : dim_y - 1;
		for (unsigned int x = 0; x < bsd.dim_x; x++) {
			unsigned int xi = pos_x + x < dim_x - 1 ? pos_x + x : dim_x - 1;
			size_t off = (4 * (size_t)dim_x * dim_y * zi) + (4 * (size_t)dim_x * yi) + (4 * xi);
			f4 v;
			if (data_type == 0) {
				const uint8_t* p = static_cast<const uint8_t*>(data) + off;
				v = mk4(static_cast<float>(p[0]), static_cast<float>(p[1]), static_cast<float>(p[2]), static_cast<float>(p[3])) / 255.0f;
			} else if (data_type == 1) {
				const uint16_t* p = static_cast<const uint16_t*>(data) + off;
				v = mk4(sf16_to_float(p[0]), sf16_to_float(p[1]), sf16_to_float(p[2]), sf16_to_float(p[3]));
			} else {
				const float* p = static_cast<const float*>(data) + off;
				v = mk4(p[0], p[1], p[2], p[3]);
			}
			if (needs_swz) {
				float s[6] = {v.x, v.y, v.z, v.w, 0.0f, 1.0f};
				v = mk4(s[swz[0]], s[swz[1]], s[swz[2]], s[swz[3]]);
			}
			f4 un = vclamp4(0.0f, 65535.0f, v * 65535.0f);
			if (rgb_lns || a_lns) {
				f4 l = mk4(float_to_lns(v.x), float_to_lns(v.y), float_to_lns(v.z), float_to_lns(v.w));
				v = mk4(rgb_lns ? l.x : un.x, rgb_lns ? l.y : un.y, rgb_lns ? l.z : un.z, a_lns ? l.w : un.w);
			} else {
				v = un;
			}
			dmin = min4(dmin, v);
			dmean = dmean + v * mean_scale;
			dmax = max4(dmax, v);
			gray = gray && (v.x == v.y) && (v.x == v.z);
			blk.data_r[idx] = v.x;
			blk.data_g[idx] = v.y;
			blk.data_b[idx] = v.z;
			blk.data_a[idx] = v.w;
			idx++;
		}
	}
	}
	blk.rgb_lns0 = rgb_lns;
	blk.alpha_lns0 = a_lns;
	f4 enc = texel4(blk, 0);
	f4 enc_unorm = enc / 65535.0f;
	f4 enc_lns = splat4(0.0f);
	if (rgb_lns || a_lns) {
		enc_lns = mk4(sf16_to_float((uint16_t)lns_to_sf16(f2i(enc.x))), sf16_to_float((uint16_t)lns_to_sf16(f2i(enc.y))),
		              sf16_to_float((uint16_t)lns_to_sf16(f2i(enc.z))), sf16_to_float((uint16_t)lns_to_sf16(f2i(enc.w))));
	}
	blk.origin_texel = mk4(rgb_lns ? enc_lns.x : enc_unorm.x, rgb_lns ? enc_lns.y : enc_unorm.y,
	                       rgb_lns ? enc_lns.z : enc_unorm.z, a_lns ? enc_lns.w : enc_unorm.w);
	blk.data_min = dmin;
	blk.data_mean = dmean;
	blk.data_max = dmax;
	blk.grayscale = gray;
}

// =============================================================================================
// Averages and directions (astcenc_averages_and_directions.cpp)
// =============================================================================================
struct PartitionMetrics {
	f4 avg;
	f4 dir;
};

// compute_partition_averages_rgba :218-385 / _rgb :47-215 (ncomp = 4 or 3)
static void compute_partition_averages(const PartitionInfo& pi, const ImageBlock& blk, int ncomp, f4 averages[4]) {
	unsigned int pc = pi.partition_count;
	unsigned int texel_count = blk.texel_count;
	f4 mean = ncomp == 4 ? blk.data_mean : mk4(blk.data_mean.x, blk.data_mean.y, blk.data_mean.z, 0.0f);
	if (pc == 1) {
		averages[0] = mean;
		return;
	}
	// masked haccumulate: texel i adds into lane (i mod 4) of its partition's accumulator
	float acc[3][4][4];
	memset(acc, 0, sizeof(acc));
	const float* chan[4] = {blk.data_r, blk.data_g, blk.data_b, blk.data_a};
	for (unsigned int i = 0; i < texel_count; i++) {
		unsigned int p = pi.partition_of_texel[i];
		if (p < pc - 1) {
			for (int c = 0; c < ncomp; c++) {
				acc[p][c][i & 3] = acc[p][c][i & 3] + chan[c][i];
			}
		}
	}
	f4 block_total = mean * static_cast<float>(blk.texel_count);
	f4 rest = block_total;
	for (unsigned int p = 0; p < pc - 1; p++) {
		f4 total = splat4(0.0f);
		for (int c = 0; c < ncomp; c++) {
			set_lane(total, c, (acc[p][c][0] + acc[p][c][2]) + (acc[p][c][1] + acc[p][c][3]));
		}
		rest = rest - total;
		averages[p] = total / static_cast<float>(pi.partition_texel_count[p]);
	}
	averages[pc - 1] = rest / static_cast<float>(pi.partition_texel_count[pc - 1]);
}

// Shared sign-split direction estimate. ncomp lanes are live, the rest are zero.
static void compute_dirs(const PartitionInfo& pi, const float* c0, const float* c1, const float* c2, const float* c3, int ncomp,
                         const f4 averages[4], PartitionMetrics pm[4]) {
	unsigned int pc = pi.partition_count;
	for (unsigned int p = 0; p < pc; p++) {
		const uint8_t* texel_indexes = pi.texels_of_partition[p];
		unsigned int texel_count = pi.partition_texel_count[p];
		f4 average = averages[p];
		pm[p].avg = average;
		f4 sum_xp = splat4(0.0f), sum_yp = splat4(0.0f), sum_zp = splat4(0.0f), sum_wp = splat4(0.0f);
		for (unsigned int i = 0; i < texel_count; i++) {
			unsigned int iwt = texel_indexes[i];
			f4 d = mk4(c0[iwt], c1[iwt], ncomp > 2 ? c2[iwt] : 0.0f, ncomp > 3 ? c3[iwt] : 0.0f);
			d = d - average;
			f4 zero = splat4(0.0f);
			sum_xp = sum_xp + (d.x > 0.0f ? d : zero);
			sum_yp = sum_yp + (d.y > 0.0f ? d : zero);
			if (ncomp > 2) {
				sum_zp = sum_zp + (d.z > 0.0f ? d : zero);
			}
			if (ncomp > 3) {
				sum_wp = sum_wp + (d.w > 0.0f ? d : zero);
			}
		}
		float prod_xp = dot_s(sum_xp, sum_xp);
		float prod_yp = dot_s(sum_yp, sum_yp);
		f4 best_vector = sum_xp;
		float best_sum = prod_xp;
		if (prod_yp > best_sum) {
			best_vector = sum_yp;
			best_sum = prod_yp;
		}
		if (ncomp > 2) {
			float prod_zp = dot_s(sum_zp, sum_zp);
			if (prod_zp > best_sum) {
				best_vector = sum_zp;
				best_sum = prod_zp;
			}
		}
		if (ncomp > 3) {
			float prod_wp = dot_s(sum_wp, sum_wp);
			if (prod_wp > best_sum) {
				best_vector = sum_wp;
			}
		}
		pm[p].dir = best_vector;
	}
}

static void compute_avgs_and_dirs_4_comp(const PartitionInfo& pi, const ImageBlock& blk, PartitionMetrics pm[4]) {   // :388-456
	f4 averages[4];
	compute_partition_averages(pi, blk, 4, averages);
	compute_dirs(pi, blk.data_r, blk.data_g, blk.data_b, blk.data_a, 4, averages, pm);
}

static void compute_avgs_and_dirs_3_comp(const PartitionInfo& pi, const ImageBlock& blk, unsigned int omitted, PartitionMetrics pm[4]) {   // :459-565
	f4 averages[4] = {splat4(0.0f), splat4(0.0f), splat4(0.0f), splat4(0.0f)};
	compute_partition_averages(pi, blk, 4, averages);
	const float* vr = blk.data_r;
	const float* vg = blk.data_g;
	const float* vb = blk.data_b;
	for (int i = 0; i < 4; i++) {
		f4 a = averages[i];
		if (omitted == 0) averages[i] = mk4(a.y, a.z, a.w, 0.0f);
		else if (omitted == 1) averages[i] = mk4(a.x, a.z, a.w, 0.0f);
		else if (omitted == 2) averages[i] = mk4(a.x, a.y, a.w, 0.0f);
		else averages[i] = mk4(a.x, a.y, a.z, 0.0f);
	}
	if (omitted == 0) {
		vr = blk.data_g;
		vg = blk.data_b;
		vb = blk.data_a;
	} else if (omitted == 1) {
		vg = blk.data_b;
		vb = blk.data_a;
	} else if (omitted == 2) {
		vb = blk.data_a;
	}
	compute_dirs(pi, vr, vg, vb, nullptr, 3, averages, pm);
}

static void compute_avgs_and_dirs_3_comp_rgb(const PartitionInfo& pi, const ImageBlock& blk, PartitionMetrics pm[4]) {   // :568-628
	f4 averages[4];
	compute_partition_averages(pi, blk, 3, averages);
	compute_dirs(pi, blk.data_r, blk.data_g, blk.data_b, nullptr, 3, averages, pm);
}

static void compute_avgs_and_dirs_2_comp(const PartitionInfo& pt, const ImageBlock& blk, unsigned int comp1, unsigned int comp2, PartitionMetrics pm[4]) {   // :631-720
	const float* chan[4] = {blk.data_r, blk.data_g, blk.data_b, blk.data_a};
	const float* vr = chan[comp1];
	const float* vg = chan[comp2];
	f4 averages[4];
	unsigned int pc = pt.partition_count;
	for (unsigned int p = 0; p < pc; p++) {
		f4 average = mk4(lane(blk.data_mean, (int)comp1), lane(blk.data_mean, (int)comp2), 0.0f, 0.0f);
		if (pc > 1) {
			average = splat4(0.0f);
			unsigned int n = pt.partition_texel_count[p];
			for (unsigned int i = 0; i < n; i++) {
				unsigned int iwt = pt.texels_of_partition[p][i];
				average = average + mk4(vr[iwt], vg[iwt], 0.0f, 0.0f);
			}
			average = average / static_cast<float>(n);
		}
		averages[p] = average;
	}
	compute_dirs(pt, vr, vg, nullptr, nullptr, 2, averages, pm);
}

// compute_error_squared_rgba :723-840 / compute_error_squared_rgb :843-945
struct ProcessedLine {
	f4 amod;
	f4 bs;
};

static void compute_error_squared(const PartitionInfo& pi, const ImageBlock& blk, int ncomp, const ProcessedLine uncor[4], const ProcessedLine samec[4],
                                  float line_lengths[4], float& uncor_error, float& samec_error) {
	unsigned int pc = pi.partition_count;
	acc4 uacc, sacc;
	acc_init(uacc);
	acc_init(sacc);
	f4 ew = blk.channel_weight;
	for (unsigned int p = 0; p < pc; p++) {
		const uint8_t* texel_indexes = pi.texels_of_partition[p];
		unsigned int texel_count = pi.partition_texel_count[p];
		f4 ub = uncor[p].bs, ua = uncor[p].amod, sb = samec[p].bs;
		float lo = 1e10f, hi = -1e10f;
		acc_restart(uacc);
		acc_restart(sacc);
		for (unsigned int i = 0; i < texel_count; i++) {
			unsigned int t = texel_indexes[i];
			float r = blk.data_r[t], g = blk.data_g[t], b = blk.data_b[t], a = blk.data_a[t];
			float uparam, uerr, serr;
			if (ncomp == 4) {
				uparam = (r * ub.x) + (g * ub.y) + (b * ub.z) + (a * ub.w);
				float d0 = (ua.x - r) + (uparam * ub.x);
				float d1 = (ua.y - g) + (uparam * ub.y);
				float d2 = (ua.z - b) + (uparam * ub.z);
				float d3 = (ua.w - a) + (uparam * ub.w);
				uerr = (ew.x * d0 * d0) + (ew.y * d1 * d1) + (ew.z * d2 * d2) + (ew.w * d3 * d3);
				float sparam = (r * sb.x) + (g * sb.y) + (b * sb.z) + (a * sb.w);
				float s0 = sparam * sb.x - r;
				float s1 = sparam * sb.y - g;
				float s2 = sparam * sb.z - b;
				float s3 = sparam * sb.w - a;
				serr = (ew.x * s0 * s0) + (ew.y * s1 * s1) + (ew.z * s2 * s2) + (ew.w * s3 * s3);
			} else {
				uparam = (r * ub.x) + (g * ub.y) + (b * ub.z);
				float d0 = (ua.x - r) + (uparam * ub.x);
				float d1 = (ua.y - g) + (uparam * ub.y);
				float d2 = (ua.z - b) + (uparam * ub.z);
				uerr = (ew.x * d0 * d0) + (ew.y * d1 * d1) + (ew.z * d2 * d2);
				float sparam = (r * sb.x) + (g * sb.y) + (b * sb.z);
				float s0 = sparam * sb.x - r;
				float s1 = sparam * sb.y - g;
				float s2 = sparam * sb.z - b;
				serr = (ew.x * s0 * s0) + (ew.y * s1 * s1) + (ew.z * s2 * s2);
			}
			lo = minf(uparam, lo);
			hi = maxf(uparam, hi);
			acc_add(uacc, uerr);
			acc_add(sacc, serr);
		}
		float linelen = hi - lo;
		line_lengths[p] = maxf(linelen, 1e-7f);
	}
	uncor_error = acc_sum(uacc);
	samec_error = acc_sum(sacc);
}

// =============================================================================================
// Ideal endpoints and weights (astcenc_ideal_endpoints_and_weights.cpp:107-683)
// =============================================================================================
static void compute_ideal_colors_and_weights_1_comp(const ImageBlock& blk, const PartitionInfo& pi, EndpointsAndWeights& ei, unsigned int component) {   // :107-206
	unsigned int pc = pi.partition_count;
	ei.ep.partition_count = pc;
	const float* chan[4] = {blk.data_r, blk.data_g, blk.data_b, blk.data_a};
	const float* data_vr = chan[component];
	float error_weight = lane(blk.channel_weight, (int)component);
	bool is_constant_wes = true;
	float partition0_len_sq = 0.0f;
	for (unsigned int i = 0; i < pc; i++) {
		float lowvalue = 1e10f, highvalue = -1e10f;
		unsigned int n = pi.partition_texel_count[i];
		for (unsigned int j = 0; j < n; j++) {
			float value = data_vr[pi.texels_of_partition[i][j]];
			lowvalue = minf(value, lowvalue);
			highvalue = maxf(value, highvalue);
		}
		if (highvalue <= lowvalue) {
			lowvalue = 0.0f;
			highvalue = 1e-7f;
		}
		float length = highvalue - lowvalue;
		float length_squared = length * length;
		float scale = 1.0f / length;
		if (i == 0) {
			partition0_len_sq = length_squared;
		} else {
			is_constant_wes = is_constant_wes && length_squared == partition0_len_sq;
		}
		for (unsigned int j = 0; j < n; j++) {
			unsigned int tix = pi.texels_of_partition[i][j];
			float value = (data_vr[tix] - lowvalue) * scale;
			value = clamp1f(value);
			ei.weights[tix] = value;
			ei.weight_error_scale[tix] = length_squared * error_weight;
		}
		ei.ep.endpt0[i] = blk.data_min;
		ei.ep.endpt1[i] = blk.data_max;
		set_lane(ei.ep.endpt0[i], (int)component, lowvalue);
		set_lane(ei.ep.endpt1[i], (int)component, highvalue);
	}
	ei.is_constant_weight_error_scale = is_constant_wes;
}

// Shared tail for the 2/3/4 component variants: project on the line, normalise, weight error scale.
static void ideal_project(const ImageBlock& blk, const PartitionInfo& pi, EndpointsAndWeights& ei, const PartitionMetrics pms[4], int ncomp,
                          const float* c0, const float* c1, const float* c2, const float* c3, float error_weight, f4 lowv[4], f4 highv[4]) {
	unsigned int pc = pi.partition_count;
	bool is_constant_wes = true;
	float partition0_len_sq = 0.0f;
	for (unsigned int i = 0; i < pc; i++) {
		f4 dir = pms[i].dir;
		float dsum = ncomp == 2 ? hadd_s(dir) : hadd_rgb_s(dir);
		if (dsum < 0.0f) {
			dir = splat4(0.0f) - dir;
		}
		f4 la = pms[i].avg;
		f4 lb = normalize_safe4(dir, ncomp == 2 ? unit2() : ncomp == 3 ? unit3() : unit4());
		float lowparam = 1e10f, highparam = -1e10f;
		unsigned int n = pi.partition_texel_count[i];
		for (unsigned int j = 0; j < n; j++) {
			unsigned int tix = pi.texels_of_partition[i][j];
			f4 point = mk4(c0[tix], c1[tix], ncomp > 2 ? c2[tix] : 0.0f, ncomp > 3 ? c3[tix] : 0.0f);
			float param = ncomp == 3 ? dot3_s(point - la, lb) : dot_s(point - la, lb);
			ei.weights[tix] = param;
			lowparam = minf(param, lowparam);
			highparam = maxf(param, highparam);
		}
		if (highparam <= lowparam) {
			lowparam = 0.0f;
			highparam = 1e-7f;
		}
		float length = highparam - lowparam;
		float length_squared = length * length;
		float scale = 1.0f / length;
		if (i == 0) {
			partition0_len_sq = length_squared;
		} else {
			is_constant_wes = is_constant_wes && length_squared == partition0_len_sq;
		}
		for (unsigned int j = 0; j < n; j++) {
			unsigned int tix = pi.texels_of_partition[i][j];
			float idx = (ei.weights[tix] - lowparam) * scale;
			idx = clamp1f(idx);
			ei.weights[tix] = idx;
			ei.weight_error_scale[tix] = length_squared * error_weight;
		}
		lowv[i] = la + lb * lowparam;
		highv[i] = la + lb * highparam;
	}
	ei.is_constant_weight_error_scale = is_constant_wes;
}

static void compute_ideal_colors_and_weights_2_comp(const ImageBlock& blk, const PartitionInfo& pi, EndpointsAndWeights& ei, int comp1, int comp2) {   // :217-351
	ei.ep.partition_count = pi.partition_count;
	const float* chan[4] = {blk.data_r, blk.data_g, blk.data_b, blk.data_a};
	f4 cw = blk.channel_weight;
	// hadd_s over a 2-lane swizzle (other lanes zero): (l0 + 0) + (l1 + 0)
	float error_weight = ((lane(cw, comp1) + 0.0f) + (lane(cw, comp2) + 0.0f)) / 2.0f;
	PartitionMetrics pms[4];
	compute_avgs_and_dirs_2_comp(pi, blk, (unsigned int)comp1, (unsigned int)comp2, pms);
	f4 lowv[4], highv[4];
	ideal_project(blk, pi, ei, pms, 2, chan[comp1], chan[comp2], nullptr, nullptr, error_weight, lowv, highv);
	for (unsigned int i = 0; i < pi.partition_count; i++) {
		f4 ep0 = blk.data_min, ep1 = blk.data_max;
		set_lane(ep0, comp1, lowv[i].x);
		set_lane(ep1, comp1, highv[i].x);
		set_lane(ep0, comp2, lowv[i].y);
		set_lane(ep1, comp2, highv[i].y);
		ei.ep.endpt0[i] = ep0;
		ei.ep.endpt1[i] = ep1;
	}
}

static void compute_ideal_colors_and_weights_3_comp(const ImageBlock& blk, const PartitionInfo& pi, EndpointsAndWeights& ei, unsigned int omitted) {   // :354-517
	ei.ep.partition_count = pi.partition_count;
	f4 cw = blk.channel_weight;
	const float *vr, *vg, *vb;
	float error_weight;
	// hadd_s over a 3-lane swizzle (lane 3 zero): (l0 + l2) + (l1 + 0)
	if (omitted == 0) {
		error_weight = (cw.x + cw.z) + (cw.y + 0.0f);   // reference swizzles <0,1,2> here too (:377)
		vr = blk.data_g; vg = blk.data_b; vb = blk.data_a;
	} else if (omitted == 1) {
		error_weight = (cw.x + cw.w) + (cw.z + 0.0f);
		vr = blk.data_r; vg = blk.data_b; vb = blk.data_a;
	} else if (omitted == 2) {
		error_weight = (cw.x + cw.w) + (cw.y + 0.0f);
		vr = blk.data_r; vg = blk.data_g; vb = blk.data_a;
	} else {
		error_weight = (cw.x + cw.z) + (cw.y + 0.0f);
		vr = blk.data_r; vg = blk.data_g; vb = blk.data_b;
	}
	error_weight = error_weight * (1.0f / 3.0f);
	PartitionMetrics pms[4];
	if (omitted == 3) {
		compute_avgs_and_dirs_3_comp_rgb(pi, blk, pms);
	} else {
		compute_avgs_and_dirs_3_comp(pi, blk, omitted, pms);
	}
	f4 lowv[4], highv[4];
	ideal_project(blk, pi, ei, pms, 3, vr, vg, vb, nullptr, error_weight, lowv, highv);
	for (unsigned int i = 0; i < pi.partition_count; i++) {
		f4 e0 = lowv[i], e1 = highv[i];
		f4 bmin = blk.data_min, bmax = blk.data_max;
		switch (omitted) {
		case 0:
			ei.ep.endpt0[i] = mk4(bmin.x, e0.x, e0.y, e0.z);
			ei.ep.endpt1[i] = mk4(bmax.x, e1.x, e1.y, e1.z);
			break;
		case 1:
			ei.ep.endpt0[i] = mk4(e0.x, bmin.y, e0.y, e0.z);
			ei.ep.endpt1[i] = mk4(e1.x, bmax.y, e1.y, e1.z);
			break;
		case 2:
			ei.ep.endpt0[i] = mk4(e0.x, e0.y, bmin.z, e0.z);
			ei.ep.endpt1[i] = mk4(e1.x, e1.y, bmax.z, e1.z);
			break;
		default:
			ei.ep.endpt0[i] = mk4(e0.x, e0.y, e0.z, bmin.w);
			ei.ep.endpt1[i] = mk4(e1.x, e1.y, e1.z, bmax.w);
			break;
		}
	}
}

static void compute_ideal_colors_and_weights_4_comp(const ImageBlock& blk, const PartitionInfo& pi, EndpointsAndWeights& ei) {   // :520-609
	ei.ep.partition_count = pi.partition_count;
	float error_weight = hadd_s(blk.channel_weight) / 4.0f;
	PartitionMetrics pms[4];
	compute_avgs_and_dirs_4_comp(pi, blk, pms);
	f4 lowv[4], highv[4];
	ideal_project(blk, pi, ei, pms, 4, blk.data_r, blk.data_g, blk.data_b, blk.data_a, error_weight, lowv, highv);
	for (unsigned int i = 0; i < pi.partition_count; i++) {
		ei.ep.endpt0[i] = lowv[i];
		ei.ep.endpt1[i] = highv[i];
	}
}

static void compute_ideal_colors_and_weights_1plane(const ImageBlock& blk, const PartitionInfo& pi, EndpointsAndWeights& ei) {   // :612-627
	bool uses_alpha = !is_constant_channel(blk, 3);
	if (uses_alpha) {
		compute_ideal_colors_and_weights_4_comp(blk, pi, ei);
	} else {
		compute_ideal_colors_and_weights_3_comp(blk, pi, ei, 3);
	}
}

static void compute_ideal_colors_and_weights_2planes(const BlockSizeTables& bsd, const ImageBlock& blk, unsigned int plane2_component,
                                                     EndpointsAndWeights& ei1, EndpointsAndWeights& ei2) {   // :630-683
	const PartitionInfo& pi = bsd.partitionings[1][0];
	bool uses_alpha = !is_constant_channel(blk, 3);
	switch (plane2_component) {
	case 0:
		if (uses_alpha) compute_ideal_colors_and_weights_3_comp(blk, pi, ei1, 0);
		else compute_ideal_colors_and_weights_2_comp(blk, pi, ei1, 1, 2);
		compute_ideal_colors_and_weights_1_comp(blk, pi, ei2, 0);
		break;
	case 1:
		if (uses_alpha) compute_ideal_colors_and_weights_3_comp(blk, pi, ei1, 1);
		else compute_ideal_colors_and_weights_2_comp(blk, pi, ei1, 0, 2);
		compute_ideal_colors_and_weights_1_comp(blk, pi, ei2, 1);
		break;
	case 2:
		if (uses_alpha) compute_ideal_colors_and_weights_3_comp(blk, pi, ei1, 2);
		else compute_ideal_colors_and_weights_2_comp(blk, pi, ei1, 0, 1);
		compute_ideal_colors_and_weights_1_comp(blk, pi, ei2, 2);
		break;
	default:
		compute_ideal_colors_and_weights_3_comp(blk, pi, ei1, 3);
		compute_ideal_colors_and_weights_1_comp(blk, pi, ei2, 3);
		break;
	}
}

// bilinear_infill_vla / _2 (:38-104): (w0*c0 + w1*c1) + (w2*c2 + w3*c3)
static inline float bilinear_infill(const DecimationInfo& di, const float* weights, unsigned int t) {
	return (weights[di.texel_weights[0][t]] * di.texel_weight_contribs_float[0][t] +
	        weights[di.texel_weights[1][t]] * di.texel_weight_contribs_float[1][t]) +
	       (weights[di.texel_weights[2][t]] * di.texel_weight_contribs_float[2][t] +
	        weights[di.texel_weights[3][t]] * di.texel_weight_contribs_float[3][t]);
}
static inline float bilinear_infill_2(const DecimationInfo& di, const float* weights, unsigned int t) {
	return (weights[di.texel_weights[0][t]] * di.texel_weight_contribs_float[0][t] +
	        weights[di.texel_weights[1][t]] * di.texel_weight_contribs_float[1][t]);
}
static inline float infill_any(const DecimationInfo& di, const float* weights, unsigned int t) {
	if (di.max_texel_weight_count > 2) return bilinear_infill(di, weights, t);
	if (di.max_texel_weight_count > 1) return bilinear_infill_2(di, weights, t);
	return weights[t];
}

// compute_error_of_weight_set_1plane :688-749
static float compute_error_of_weight_set_1plane(const EndpointsAndWeights& eai, const DecimationInfo& di, const float* dec_weight_quant_uvalue) {
	acc4 acc;
	acc_init(acc);
	unsigned int texel_count = di.texel_count;
	for (unsigned int i = 0; i < texel_count; i++) {
		float current = infill_any(di, dec_weight_quant_uvalue, i);
		float diff = current - eai.weights[i];
		float error = diff * diff * eai.weight_error_scale[i];
		acc_add(acc, error);
	}
	return acc_sum(acc);
}

// compute_error_of_weight_set_2planes :752-842
static float compute_error_of_weight_set_2planes(const EndpointsAndWeights& eai1, const EndpointsAndWeights& eai2, const DecimationInfo& di,
                                                 const float* uvalue1, const float* uvalue2) {
	acc4 acc;
	acc_init(acc);
	unsigned int texel_count = di.texel_count;
	for (unsigned int i = 0; i < texel_count; i++) {
		float diff = infill_any(di, uvalue1, i) - eai1.weights[i];
		float error1 = diff * diff * eai1.weight_error_scale[i];
		diff = infill_any(di, uvalue2, i) - eai2.weights[i];
		float error2 = diff * diff * eai2.weight_error_scale[i];
		acc_add(acc, error1 + error2);
	}
	return acc_sum(acc);
}

// compute_ideal_weights_for_decimation :845-971
static void compute_ideal_weights_for_decimation(const EndpointsAndWeights& ei, const DecimationInfo& di, float* dec_weight_ideal_value) {
	unsigned int texel_count = di.texel_count;
	unsigned int weight_count = di.weight_count;
	if (texel_count == weight_count) {
		for (unsigned int i = 0; i < texel_count; i++) {
			dec_weight_ideal_value[i] = ei.weights[i];
		}
		return;
	}
	bool constant_wes = ei.is_constant_weight_error_scale;
	float wes0 = ei.weight_error_scale[0];
	for (unsigned int i = 0; i < weight_count; i++) {
		float weight_weight = 1e-10f;
		float initial_weight = 0.0f;
		unsigned int off = di.weight_texel_offset[i];
		unsigned int cnt = di.weight_texel_count[i];
		for (unsigned int j = 0; j < cnt; j++) {
			unsigned int texel = di.weight_texels[off + j];
			float weight = di.weight_texel_contribs[off + j];
			float wes = constant_wes ? wes0 : ei.weight_error_scale[texel];
			float contrib_weight = weight * wes;
			weight_weight += contrib_weight;
			initial_weight += ei.weights[texel] * contrib_weight;
		}
		dec_weight_ideal_value[i] = initial_weight / weight_weight;
	}
	float infilled[MAX_TEXELS];
	for (unsigned int i = 0; i < texel_count; i++) {
		infilled[i] = di.max_texel_weight_count <= 2 ? bilinear_infill_2(di, dec_weight_ideal_value, i) : bilinear_infill(di, dec_weight_ideal_value, i);
	}
	const float stepsize = 0.25f;
	const float chd_scale = -16.0f;
	for (unsigned int i = 0; i < weight_count; i++) {
		float weight_val = dec_weight_ideal_value[i];
		float error_change0 = 1e-10f;
		float error_change1 = 0.0f;
		unsigned int off = di.weight_texel_offset[i];
		unsigned int cnt = di.weight_texel_count[i];
		for (unsigned int j = 0; j < cnt; j++) {
			unsigned int texel = di.weight_texels[off + j];
			float contrib_weight = di.weight_texel_contribs[off + j];
			float wes = constant_wes ? wes0 : ei.weight_error_scale[texel];
			float scale = wes * contrib_weight;
			float old_weight = infilled[texel];
			float ideal_weight = ei.weights[texel];
			error_change0 += contrib_weight * scale;
			error_change1 += (old_weight - ideal_weight) * scale;
		}
		float step = (error_change1 * chd_scale) / error_change0;
		step = vclampf(-stepsize, stepsize, step);
		dec_weight_ideal_value[i] = weight_val + step;
	}
}

// compute_quantized_weights_for_decimation :974-1080
static void compute_quantized_weights_for_decimation(const DecimationInfo& di, float low_bound, float high_bound, const float* dec_weight_ideal_value,
                                                     float* weight_set_out, uint8_t* quantized_weight_set, int quant_level) {
	int weight_count = di.weight_count;
	const WeightQuantTable& qat = const_tables().weight_quant[quant_level];
	static const float quant_levels_m1[12] = {1.0f, 2.0f, 3.0f, 4.0f, 5.0f, 7.0f, 9.0f, 11.0f, 15.0f, 19.0f, 23.0f, 31.0f};
	int steps_m1 = (int)get_quant_level(quant_level) - 1;
	float quant_level_m1 = quant_levels_m1[quant_level];
	if (high_bound <= low_bound) {
		low_bound = 0.0f;
		high_bound = 1.0f;
	}
	float rscale = high_bound - low_bound;
	float scale = 1.0f / rscale;
	float scaled_low_bound = low_bound * scale;
	rscale *= 1.0f / 64.0f;
	for (int i = 0; i < weight_count; i++) {
		float ix = dec_weight_ideal_value[i] * scale - scaled_low_bound;
		ix = clampzo(ix);
		float ix1 = ix * quant_level_m1;
		int weightl = f2i(ix1);
		int weighth = mini(weightl + 1, steps_m1);
		int ixli = qat.quant_to_unquant[weightl];
		int ixhi = qat.quant_to_unquant[weighth];
		float ixl = static_cast<float>(ixli);
		float ixh = static_cast<float>(ixhi);
		bool mask = (ixl + ixh) < (128.0f * ix);
		int weight = mask ? ixhi : ixli;
		ixl = mask ? ixh : ixl;
		weight_set_out[i] = ixl * rscale + low_bound;
		quantized_weight_set[i] = (uint8_t)weight;
	}
}

// =============================================================================================
// Angular weight-range search (astcenc_weight_align.cpp)
// =============================================================================================
static const uint8_t STEPS_FOR_QUANT_LEVEL[12] = {2, 3, 4, 5, 6, 8, 10, 12, 16, 20, 24, 32};

static void compute_angular_endpoints_for_quant_levels(unsigned int weight_count, const float* dec_weight_ideal_value, unsigned int max_quant_level,
                                                       float low_value[8], float high_value[8]) {   // :256-355
	const ConstTables& ct = const_tables();
	unsigned int max_quant_steps = STEPS_FOR_QUANT_LEVEL[max_quant_level];
	unsigned int max_angular_steps = STEPS_FOR_QUANT_LEVEL[max_quant_level];

	// compute_angular_offsets :94-157
	float angular_offsets[32];
	int isamplev[MAX_WEIGHTS];
	for (unsigned int i = 0; i < weight_count; i++) {
		float sample = clampzo(dec_weight_ideal_value[i]) * (64 - 1.0f);
		isamplev[i] = f2i_rtn(sample);
	}
	const float mult = 1.0f / (2.0f * 3.14159265358979323846f);
	for (unsigned int s = 0; s < max_angular_steps; s++) {
		float anglesum_x = 0.0f, anglesum_y = 0.0f;
		for (unsigned int j = 0; j < weight_count; j++) {
			anglesum_x += ct.cos_table[isamplev[j]][s];
			anglesum_y += ct.sin_table[isamplev[j]][s];
		}
		float angle = approx_atan2(anglesum_y, anglesum_x);
		angle = (angle == angle) ? angle : 0.0f;
		angular_offsets[s] = angle * mult;
	}

	// compute_lowest_and_highest_weight :160-253
	float lowest_weight[32];
	int weight_span[32];
	float error[32], cut_low_weight_error[32], cut_high_weight_error[32];
	float min_weight = 3.402823466e+38f, max_weight = -3.402823466e+38f;
	for (unsigned int i = 0; i < weight_count; i++) {
		min_weight = minf(dec_weight_ideal_value[i], min_weight);
		max_weight = maxf(dec_weight_ideal_value[i], max_weight);
	}
	for (unsigned int sp = 0; sp < max_angular_steps; sp++) {
		float rcp_stepsize = static_cast<float>(sp) + 1.0f;
		float offset = angular_offsets[sp];
		float minidx = round_ne(min_weight * rcp_stepsize - offset);
		float maxidx = round_ne(max_weight * rcp_stepsize - offset);
		float errval = 0.0f, cut_low = 0.0f, cut_high = 0.0f;
		for (unsigned int j = 0; j < weight_count; j++) {
			float sval = dec_weight_ideal_value[j] * rcp_stepsize - offset;
			float svalrte = round_ne(sval);
			float diff = sval - svalrte;
			errval += diff * diff;
			if (svalrte == minidx) {
				cut_low = cut_low + 1.0f - 2.0f * diff;
			}
			if (svalrte == maxidx) {
				cut_high = cut_high + 1.0f + 2.0f * diff;
			}
		}
		int span = f2i(maxidx - minidx + 1.0f);
		span = mini(span, (int)max_quant_steps + 3);
		span = maxi(span, 2);
		lowest_weight[sp] = minidx;
		weight_span[sp] = span;
		float ssize = 1.0f / rcp_stepsize;
		float errscale = ssize * ssize;
		error[sp] = errval * errscale;
		cut_low_weight_error[sp] = cut_low * errscale;
		cut_high_weight_error[sp] = cut_high * errscale;
	}

	// :281-331
	float best_err[36], best_idx[36], best_cut[36];
	for (unsigned int i = 0; i < (max_quant_steps + 4); i++) {
		best_err[i] = ERROR_CALC_DEFAULT;
		best_idx[i] = -1.0f;
		best_cut[i] = 0.0f;
	}
	for (unsigned int i = 0; i < max_angular_steps; i++) {
		float i_flt = static_cast<float>(i);
		int idx_span = weight_span[i];
		float error_cut_low = error[i] + cut_low_weight_error[i];
		float error_cut_high = error[i] + cut_high_weight_error[i];
		float error_cut_low_high = error[i] + cut_low_weight_error[i] + cut_high_weight_error[i];
		if (best_err[idx_span] > error[i]) {
			best_err[idx_span] = error[i];
			best_idx[idx_span] = i_flt;
			best_cut[idx_span] = 0.0f;
		}
		if (best_err[idx_span - 1] > error_cut_low) {
			best_err[idx_span - 1] = error_cut_low;
			best_idx[idx_span - 1] = i_flt;
			best_cut[idx_span - 1] = 1.0f;
		}
		if (best_err[idx_span - 1] > error_cut_high) {
			best_err[idx_span - 1] = error_cut_high;
			best_idx[idx_span - 1] = i_flt;
			best_cut[idx_span - 1] = 0.0f;
		}
		if (best_err[idx_span - 2] > error_cut_low_high) {
			best_err[idx_span - 2] = error_cut_low_high;
			best_idx[idx_span - 2] = i_flt;
			best_cut[idx_span - 2] = 1.0f;
		}
	}
	for (unsigned int i = 0; i <= max_quant_level; i++) {
		unsigned int q = STEPS_FOR_QUANT_LEVEL[i];
		int bsi = (int)best_idx[q];
		bsi = maxi(0, bsi);
		float lwi = lowest_weight[bsi] + best_cut[q];
		float hwi = lwi + static_cast<float>(q) - 1.0f;
		float stepsize = 1.0f / (1.0f + static_cast<float>(bsi));
		low_value[i] = (angular_offsets[bsi] + lwi) * stepsize;
		high_value[i] = (angular_offsets[bsi] + hwi) * stepsize;
	}
}

static void compute_angular_endpoints_1plane(bool only_always, const BlockSizeTables& bsd, const float* dec_weight_ideal_value,
                                             unsigned int max_weight_quant, WorkBuf& tmp) {   // :358-423
	unsigned int max_dm = only_always ? bsd.decimation_mode_count_always : bsd.decimation_mode_count_selected;
	for (unsigned int i = 0; i < max_dm; i++) {
		const DecimationMode& dm = bsd.decimation_modes[i];
		uint16_t mask = (uint16_t)((1u << (max_weight_quant + 1)) - 1);
		if ((dm.refprec_1plane & mask) == 0) {
			continue;
		}
		unsigned int weight_count = bsd.decimation_tables[i].weight_count;
		unsigned int max_precision = (unsigned int)dm.maxprec_1plane;
		if (max_precision > (unsigned int)TUNE_MAX_ANGULAR_QUANT) max_precision = TUNE_MAX_ANGULAR_QUANT;
		if (max_precision > max_weight_quant) max_precision = max_weight_quant;
		compute_angular_endpoints_for_quant_levels(weight_count, dec_weight_ideal_value + i * MAX_WEIGHTS, max_precision,
		                                           tmp.weight_low_values1[i], tmp.weight_high_values1[i]);
	}
	unsigned int max_bm = only_always ? bsd.block_mode_count_1plane_always : bsd.block_mode_count_1plane_selected;
	for (unsigned int i = 0; i < max_bm; i++) {
		const BlockMode& bm = bsd.block_modes[i];
		if (bm.quant_mode <= TUNE_MAX_ANGULAR_QUANT) {
			tmp.weight_low_value1[i] = tmp.weight_low_values1[bm.decimation_mode][bm.quant_mode];
			tmp.weight_high_value1[i] = tmp.weight_high_values1[bm.decimation_mode][bm.quant_mode];
		} else {
			tmp.weight_low_value1[i] = 0.0f;
			tmp.weight_high_value1[i] = 1.0f;
		}
	}
}

static void compute_angular_endpoints_2planes(const BlockSizeTables& bsd, const float* dec_weight_ideal_value, unsigned int max_weight_quant, WorkBuf& tmp) {   // :426-500
	for (unsigned int i = 0; i < bsd.decimation_mode_count_selected; i++) {
		const DecimationMode& dm = bsd.decimation_modes[i];
		uint16_t mask = (uint16_t)((1u << (max_weight_quant + 1)) - 1);
		if ((dm.refprec_2planes & mask) == 0) {
			continue;
		}
		unsigned int weight_count = bsd.decimation_tables[i].weight_count;
		unsigned int max_precision = (unsigned int)dm.maxprec_2planes;
		if (max_precision > (unsigned int)TUNE_MAX_ANGULAR_QUANT) max_precision = TUNE_MAX_ANGULAR_QUANT;
		if (max_precision > max_weight_quant) max_precision = max_weight_quant;
		compute_angular_endpoints_for_quant_levels(weight_count, dec_weight_ideal_value + i * MAX_WEIGHTS, max_precision,
		                                           tmp.weight_low_values1[i], tmp.weight_high_values1[i]);
		compute_angular_endpoints_for_quant_levels(weight_count, dec_weight_ideal_value + i * MAX_WEIGHTS + PLANE2_OFFSET, max_precision,
		                                           tmp.weight_low_values2[i], tmp.weight_high_values2[i]);
	}
	for (unsigned int i = bsd.block_mode_count_1plane_selected; i < bsd.block_mode_count_1plane_2plane_selected; i++) {
		const BlockMode& bm = bsd.block_modes[i];
		if (bm.quant_mode <= TUNE_MAX_ANGULAR_QUANT) {
			tmp.weight_low_value1[i] = tmp.weight_low_values1[bm.decimation_mode][bm.quant_mode];
			tmp.weight_high_value1[i] = tmp.weight_high_values1[bm.decimation_mode][bm.quant_mode];
			tmp.weight_low_value2[i] = tmp.weight_low_values2[bm.decimation_mode][bm.quant_mode];
			tmp.weight_high_value2[i] = tmp.weight_high_values2[bm.decimation_mode][bm.quant_mode];
		} else {
			tmp.weight_low_value1[i] = 0.0f;
			tmp.weight_high_value1[i] = 1.0f;
			tmp.weight_low_value2[i] = 0.0f;
			tmp.weight_high_value2[i] = 1.0f;
		}
	}
}

#include "astc_codec_part2.inl"
#include "astc_decode.inl"

}  // namespace ao
