// Host-side configuration logic of the B200 ASTC compressor: presets, astcenc_config_init,
// validate_config and the dB -> squared-error conversion of astcenc_context_alloc.
// Behavioural spec: /root/reference/Source/astcenc_entry.cpp:65-135 (preset tables), :215-227
// (validate_cpu_float), :434-501 (validate_config), :504-723 (astcenc_config_init), :814-821.
// Built with -ffp-contract=off so the float arithmetic matches the reference's invariance builds.
#include "astc_host_config.h"
#include <cmath>
#include <cstring>

namespace astc_host {

static inline float maxf(float a, float b) { return a > b ? a : b; }
static inline float vclampf(float lo, float hi, float a) { float m = a > lo ? a : lo; return m < hi ? m : hi; }
static inline uint32_t f_as_u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u_as_f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline int f2i(float a) { if (!(a > -2147483904.0f && a < 2147483648.0f)) return (int)0x80000000u; return (int)a; }
static inline int f2i_rtn(float a) { return f2i(a + 0.5f); }

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

astcenc_error validate_cpu_float() {
	volatile float xprec_testval = 2.51f;
	float store = xprec_testval + 12582912.0f;
	float q = store - 12582912.0f;
	if (q != 3.0f) {
		return ASTCENC_ERR_BAD_CPU_FLOAT;
	}
	return ASTCENC_SUCCESS;
}

static astcenc_error validate_profile(astcenc_profile profile) {
	switch (static_cast<int>(profile)) {
	case ASTCENC_PRF_LDR_SRGB:
	case ASTCENC_PRF_LDR:
	case ASTCENC_PRF_HDR_RGB_LDR_A:
	case ASTCENC_PRF_HDR:
		return ASTCENC_SUCCESS;
	default:
		return ASTCENC_ERR_BAD_PROFILE;
	}
}

static astcenc_error validate_block_size(unsigned int block_x, unsigned int block_y, unsigned int block_z) {
	bool is_legal = ((block_z <= 1) && is_legal_2d_block_size(block_x, block_y)) || ((block_z >= 2) && is_legal_3d_block_size(block_x, block_y, block_z));
	if (!is_legal) {
		return ASTCENC_ERR_BAD_BLOCK_SIZE;
	}
	return ASTCENC_SUCCESS;
}

static int popcount32(unsigned int v) {
	int c = 0;
	while (v) {
		c += v & 1;
		v >>= 1;
	}
	return c;
}

static astcenc_error validate_flags(astcenc_profile profile, unsigned int flags) {
	unsigned int exMask = ~ASTCENC_ALL_FLAGS;
	if (popcount32(flags & exMask) != 0) {
		return ASTCENC_ERR_BAD_FLAGS;
	}
	exMask = ASTCENC_FLG_MAP_NORMAL | ASTCENC_FLG_MAP_RGBM;
	if (popcount32(flags & exMask) > 1) {
		return ASTCENC_ERR_BAD_FLAGS;
	}
	bool is_unorm8 = (flags & ASTCENC_FLG_USE_DECODE_UNORM8) != 0;
	bool is_hdr = (profile == ASTCENC_PRF_HDR) || (profile == ASTCENC_PRF_HDR_RGB_LDR_A);
	if (is_unorm8 && is_hdr) {
		return ASTCENC_ERR_BAD_DECODE_MODE;
	}
	return ASTCENC_SUCCESS;
}

static unsigned int clampu(unsigned int v, unsigned int mn, unsigned int mx) {
	if (v > mx) return mx;
	if (v > mn) return v;
	return mn;
}

astcenc_error validate_config(astcenc_config& config) {
	astcenc_error status = validate_profile(config.profile);
	if (status != ASTCENC_SUCCESS) return status;
	status = validate_flags(config.profile, config.flags);
	if (status != ASTCENC_SUCCESS) return status;
	status = validate_block_size(config.block_x, config.block_y, config.block_z);
	if (status != ASTCENC_SUCCESS) return status;

	config.rgbm_m_scale = maxf(config.rgbm_m_scale, 1.0f);
	config.tune_partition_count_limit = clampu(config.tune_partition_count_limit, 1u, 4u);
	config.tune_2partition_index_limit = clampu(config.tune_2partition_index_limit, 1u, 1024u);
	config.tune_3partition_index_limit = clampu(config.tune_3partition_index_limit, 1u, 1024u);
	config.tune_4partition_index_limit = clampu(config.tune_4partition_index_limit, 1u, 1024u);
	config.tune_block_mode_limit = clampu(config.tune_block_mode_limit, 1u, 100u);
	config.tune_refinement_limit = config.tune_refinement_limit > 1u ? config.tune_refinement_limit : 1u;
	config.tune_candidate_limit = clampu(config.tune_candidate_limit, 1u, 8u);
	config.tune_2partitioning_candidate_limit = clampu(config.tune_2partitioning_candidate_limit, 1u, 8u);
	config.tune_3partitioning_candidate_limit = clampu(config.tune_3partitioning_candidate_limit, 1u, 8u);
	config.tune_4partitioning_candidate_limit = clampu(config.tune_4partitioning_candidate_limit, 1u, 8u);
	config.tune_db_limit = maxf(config.tune_db_limit, 0.0f);
	config.tune_mse_overshoot = maxf(config.tune_mse_overshoot, 1.0f);
	config.tune_2partition_early_out_limit_factor = maxf(config.tune_2partition_early_out_limit_factor, 0.0f);
	config.tune_3partition_early_out_limit_factor = maxf(config.tune_3partition_early_out_limit_factor, 0.0f);
	config.tune_2plane_early_out_limit_correlation = maxf(config.tune_2plane_early_out_limit_correlation, 0.0f);
	float max_weight = maxf(maxf(config.cw_r_weight, config.cw_g_weight), maxf(config.cw_b_weight, config.cw_a_weight));
	if (max_weight > 0.0f) {
		max_weight /= 1000.0f;
		config.cw_r_weight = maxf(config.cw_r_weight, max_weight);
		config.cw_g_weight = maxf(config.cw_g_weight, max_weight);
		config.cw_b_weight = maxf(config.cw_b_weight, max_weight);
		config.cw_a_weight = maxf(config.cw_a_weight, max_weight);
	} else {
		return ASTCENC_ERR_BAD_PARAM;
	}
	return ASTCENC_SUCCESS;
}

astcenc_error config_init(astcenc_profile profile, unsigned int block_x, unsigned int block_y, unsigned int block_z, float quality, unsigned int flags,
                          astcenc_config* configp) {
	astcenc_error status = validate_cpu_float();
	if (status != ASTCENC_SUCCESS) return status;
	astcenc_config& config = *configp;
	memset(&config, 0, sizeof(config));
	block_z = block_z > 1u ? block_z : 1u;
	status = validate_block_size(block_x, block_y, block_z);
	if (status != ASTCENC_SUCCESS) return status;
	config.block_x = block_x;
	config.block_y = block_y;
	config.block_z = block_z;
	float texels = static_cast<float>(block_x * block_y * block_z);
	float ltexels = logf(texels) / logf(10.0f);
	if (quality < ASTCENC_PRE_FASTEST || quality > ASTCENC_PRE_EXHAUSTIVE) {
		return ASTCENC_ERR_BAD_QUALITY;
	}
	size_t texels_int = block_x * block_y * block_z;
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
		config.tune_partition_count_limit = p.partition_count_limit;
		config.tune_2partition_index_limit = p.p2_index_limit;
		config.tune_3partition_index_limit = p.p3_index_limit;
		config.tune_4partition_index_limit = p.p4_index_limit;
		config.tune_block_mode_limit = p.block_mode_limit;
		config.tune_refinement_limit = p.refinement_limit;
		config.tune_candidate_limit = p.candidate_limit;
		config.tune_2partitioning_candidate_limit = p.p2_cand_limit;
		config.tune_3partitioning_candidate_limit = p.p3_cand_limit;
		config.tune_4partitioning_candidate_limit = p.p4_cand_limit;
		config.tune_db_limit = maxf(p.db_limit_a_base - 35 * ltexels, p.db_limit_b_base - 19 * ltexels);
		config.tune_mse_overshoot = p.mse_overshoot;
		config.tune_2partition_early_out_limit_factor = p.p2_early_out;
		config.tune_3partition_early_out_limit_factor = p.p3_early_out;
		config.tune_2plane_early_out_limit_correlation = p.plane2_correlation;
		config.tune_search_mode0_enable = p.search_mode0;
	} else {
		const Preset& a = presets[start];
		const Preset& b = presets[end];
		float wt_range = b.quality - a.quality;
		float wa = (b.quality - quality) / wt_range;
		float wb = (quality - a.quality) / wt_range;
#define LERP(f) ((a.f * wa) + (b.f * wb))
#define LERPI(f) f2i_rtn((static_cast<float>(a.f) * wa) + (static_cast<float>(b.f) * wb))
		config.tune_partition_count_limit = (unsigned int)LERPI(partition_count_limit);
		config.tune_2partition_index_limit = (unsigned int)LERPI(p2_index_limit);
		config.tune_3partition_index_limit = (unsigned int)LERPI(p3_index_limit);
		config.tune_4partition_index_limit = (unsigned int)LERPI(p4_index_limit);
		config.tune_block_mode_limit = (unsigned int)LERPI(block_mode_limit);
		config.tune_refinement_limit = (unsigned int)LERPI(refinement_limit);
		config.tune_candidate_limit = (unsigned int)LERPI(candidate_limit);
		config.tune_2partitioning_candidate_limit = (unsigned int)LERPI(p2_cand_limit);
		config.tune_3partitioning_candidate_limit = (unsigned int)LERPI(p3_cand_limit);
		config.tune_4partitioning_candidate_limit = (unsigned int)LERPI(p4_cand_limit);
		config.tune_db_limit = maxf(LERP(db_limit_a_base) - 35 * ltexels, LERP(db_limit_b_base) - 19 * ltexels);
		config.tune_mse_overshoot = LERP(mse_overshoot);
		config.tune_2partition_early_out_limit_factor = LERP(p2_early_out);
		config.tune_3partition_early_out_limit_factor = LERP(p3_early_out);
		config.tune_2plane_early_out_limit_correlation = LERP(plane2_correlation);
		config.tune_search_mode0_enable = LERP(search_mode0);
#undef LERP
#undef LERPI
	}
	config.cw_r_weight = 1.0f;
	config.cw_g_weight = 1.0f;
	config.cw_b_weight = 1.0f;
	config.cw_a_weight = 1.0f;
	config.a_scale_radius = 0;
	config.rgbm_m_scale = 0.0f;
	config.profile = profile;
	switch (static_cast<int>(profile)) {
	case ASTCENC_PRF_LDR:
	case ASTCENC_PRF_LDR_SRGB:
		break;
	case ASTCENC_PRF_HDR_RGB_LDR_A:
	case ASTCENC_PRF_HDR:
		config.tune_db_limit = 999.0f;
		config.tune_search_mode0_enable = 0.0f;
		break;
	default:
		return ASTCENC_ERR_BAD_PROFILE;
	}
	status = validate_flags(profile, flags);
	if (status != ASTCENC_SUCCESS) return status;
	if (flags & ASTCENC_FLG_MAP_NORMAL) {
		config.tune_partition_count_limit = config.tune_partition_count_limit + 1u < 4u ? config.tune_partition_count_limit + 1u : 4u;
		config.cw_g_weight = 0.0f;
		config.cw_b_weight = 0.0f;
		config.tune_2partition_early_out_limit_factor *= 1.5f;
		config.tune_3partition_early_out_limit_factor *= 1.5f;
		config.tune_2plane_early_out_limit_correlation = 0.99f;
		config.tune_db_limit *= 1.03f;
	} else if (flags & ASTCENC_FLG_MAP_RGBM) {
		config.rgbm_m_scale = 5.0f;
		config.cw_a_weight = 2.0f * config.rgbm_m_scale;
	} else if (flags & ASTCENC_FLG_USE_PERCEPTUAL) {
		config.cw_r_weight = 0.30f * 2.25f;
		config.cw_g_weight = 0.59f * 2.25f;
		config.cw_b_weight = 0.11f * 2.25f;
	}
	config.flags = flags;
	return ASTCENC_SUCCESS;
}

void make_device_config(const astcenc_config& c, DevConfig& d) {
	memset(&d, 0, sizeof(d));
	d.profile = (int)c.profile;
	d.flags = c.flags;
	d.cw[0] = c.cw_r_weight;
	d.cw[1] = c.cw_g_weight;
	d.cw[2] = c.cw_b_weight;
	d.cw[3] = c.cw_a_weight;
	d.rgbm_m_scale = c.rgbm_m_scale;
	d.tune_partition_count_limit = c.tune_partition_count_limit;
	d.tune_partition_index_limit[0] = c.tune_2partition_index_limit;
	d.tune_partition_index_limit[1] = c.tune_3partition_index_limit;
	d.tune_partition_index_limit[2] = c.tune_4partition_index_limit;
	d.tune_refinement_limit = c.tune_refinement_limit;
	d.tune_candidate_limit = c.tune_candidate_limit;
	d.tune_partitioning_candidate_limit[0] = c.tune_2partitioning_candidate_limit;
	d.tune_partitioning_candidate_limit[1] = c.tune_3partitioning_candidate_limit;
	d.tune_partitioning_candidate_limit[2] = c.tune_4partitioning_candidate_limit;
	// astcenc_context_alloc :814-821: dB limit -> per-texel squared error
	if (c.profile == ASTCENC_PRF_LDR || c.profile == ASTCENC_PRF_LDR_SRGB) {
		d.tune_db_limit = approx_pow(0.1f, c.tune_db_limit * 0.1f) * 65535.0f * 65535.0f;
	} else {
		d.tune_db_limit = 0.0f;
	}
	d.tune_mse_overshoot = c.tune_mse_overshoot;
	d.tune_2partition_early_out_limit_factor = c.tune_2partition_early_out_limit_factor;
	d.tune_3partition_early_out_limit_factor = c.tune_3partition_early_out_limit_factor;
	d.tune_2plane_early_out_limit_correlation = c.tune_2plane_early_out_limit_correlation;
	d.tune_search_mode0_enable = c.tune_search_mode0_enable;
}

}  // namespace astc_host
