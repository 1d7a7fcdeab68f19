// ORACLE (test infrastructure, not product code): CPU restatement of the reference CLI's image error metrics,
// astcenccli_error_metrics.cpp:109-413 (compute_error_metrics). Same per-texel float arithmetic (the reference's
// polynomial log2, astcenc_vecmathlib.h:416-441; dot products in DPPS order), same raster-order double accumulators
// (:31-58), returned instead of printed. Pinned against the reference build's printed figures
// (oracle/ref_metrics_shim.cpp, tests/golden/make_golden_metrics.py -> tests/golden/golden_metrics.npz).
#include <math.h>

struct OracleErrorMetrics {
	double psnr, alpha_psnr, rgb_psnr, rgb_peak, peak_psnr, mpsnr, log_rmse, mean_angular_error, worst_angular_error;
	double sum_squared_error[4];
};

static inline float em_clampf(float lo, float hi, float v) {
	// SSE min(max(v, lo), hi): a NaN input yields lo
	float t = v > lo ? v : lo;
	return t < hi ? t : hi;
}

static void em_load(const void* data, int type, unsigned int pitch_x, unsigned int x, unsigned int y, float c[4]) {
	size_t i = 4 * ((size_t)pitch_x * y + x);
	for (int k = 0; k < 4; k++) {
		if (type == 0) {
			c[k] = (float)static_cast<const uint8_t*>(data)[i + k] / 255.0f;                   // :165-175
		} else if (type == 1) {
			c[k] = em_clampf(0.0f, 65504.0f, sf16_to_float(static_cast<const uint16_t*>(data)[i + k]));   // :176-188
		} else {
			c[k] = em_clampf(0.0f, 65504.0f, static_cast<const float*>(data)[i + k]);          // :189-201
		}
	}
}

static inline float em_log2(float x) {      // astcenc_vecmathlib.h:416-441
	int32_t i;
	memcpy(&i, &x, 4);
	float e = (float)(((i & 0x7F800000) >> 23) - 127);
	int32_t mi = (i & 0x007FFFFF) | 0x3F800000;
	float m;
	memcpy(&m, &mi, 4);
	float p = 0.0596515482674574969533f;
	p = p * m + -0.465725644288844778798f;
	p = p * m + 1.48116647521213171641f;
	p = p * m + -2.52074962577807006663f;
	p = p * m + 2.8882704548164776201f;
	p = p * (m - 1.0f);
	return p + e;
}

static inline float em_mpsnr_operator(float val, int fstop) {      // :69-80
	uint32_t us = 0x3f800000u + ((uint32_t)fstop << 23);
	float scale;
	memcpy(&scale, &us, 4);
	val = powf(val * scale, 1.0f / 2.2f);
	float v = val * 255.0f;
	return v < 0.0f ? 0.0f : (v > 255.0f ? 255.0f : v);      // astc::clamp: scalar, (v < lo) ? lo : (v > hi ? hi : v)
}

static inline float em_mpsnr_sumdiff(float v1, float v2, int lo, int hi) {      // :93-107
	float summa = 0.0f;
	for (int i = lo; i <= hi; i++) {
		float d = em_mpsnr_operator(v1, i) - em_mpsnr_operator(v2, i);
		summa += d * d;
	}
	return summa;
}

static inline void em_normal(const float c[4], float n[3]) {      // :269-275, normalize_safe astcenc_vecmathlib.h:362
	float x = (c[0] - 0.5f) * 2.0f, y = (c[1] - 0.5f) * 2.0f, z = (c[2] - 0.5f) * 2.0f;
	float len = (x * x + y * y) + (z * z + 0.0f);
	if (len != 0.0f) {
		float s = sqrtf(len);
		n[0] = x / s; n[1] = y / s; n[2] = z / s;
	} else {
		n[0] = n[1] = n[2] = 0.57735026918962576451f;
	}
}

extern "C" __attribute__((visibility("default")))
int oracle_error_metrics(int hdr, int normal, int input_components,
                         const void* data1, int type1, unsigned int w1, unsigned int h1,
                         const void* data2, int type2, unsigned int w2, unsigned int h2,
                         int fstop_lo, int fstop_hi, OracleErrorMetrics* out) {
	static const int componentmasks[5] = {0x00, 0x07, 0x0C, 0x07, 0x0F};
	if (input_components < 1 || input_components > 4) {
		return 1;
	}
	int mask = componentmasks[input_components];
	unsigned int dim_x = w1 < w2 ? w1 : w2;
	unsigned int dim_y = h1 < h2 ? h1 : h2;
	double es[4] = {0, 0, 0, 0}, as[4] = {0, 0, 0, 0}, ls[4] = {0, 0, 0, 0}, ms[4] = {0, 0, 0, 0};
	double mean_ang = 0.0, worst_ang = 0.0, rgb_peak = 0.0;
	for (unsigned int y = 0; y < dim_y; y++) {
		for (unsigned int x = 0; x < dim_x; x++) {
			float c1[4], c2[4];
			em_load(data1, type1, w1, x, y, c1);
			em_load(data2, type2, w2, x, y, c2);
			for (int k = 0; k < 3; k++) {
				if ((double)c1[k] > rgb_peak) rgb_peak = (double)c1[k];
			}
			float d[4];
			for (int k = 0; k < 4; k++) {
				d[k] = c1[k] - c2[k];
				es[k] += (double)(d[k] * d[k]);
				float ad = k < 3 ? d[k] * c1[3] : d[k];
				as[k] += (double)(ad * ad);
			}
			if (hdr) {
				for (int k = 0; k < 3; k++) {
					float l = em_log2(c1[k]) - em_log2(c2[k]);
					ls[k] += (double)(l * l);
					ms[k] += (double)em_mpsnr_sumdiff(c1[k], c2[k], fstop_lo, fstop_hi);
				}
			}
			if (normal) {
				float n1[3], n2[3];
				em_normal(c1, n1);
				em_normal(c2, n2);
				float dt = (n1[0] * n2[0] + n1[1] * n2[1]) + (n1[2] * n2[2] + 0.0f);
				dt = em_clampf(-1.0f, 1.0f, dt);
				float rad_to_degrees = 180.0f / 3.14159265358979323846f;
				double deg = acos((double)dt) * (double)rad_to_degrees;
				mean_ang += deg / (dim_x * dim_y * 1u);
				worst_ang = worst_ang > deg ? worst_ang : deg;
			}
		}
	}
	double pixels = (double)(dim_x * dim_y * 1u);
	double samples = 0.0, num = 0.0, alpha_num = 0.0, log_num = 0.0, mpsnr_num = 0.0;
	for (int c = 0; c < 4; c++) {
		if (mask & (1 << c)) {
			num += es[c];
			alpha_num += as[c];
			if (c < 3) {
				log_num += ls[c];
				mpsnr_num += ms[c];
			}
			samples += pixels;
		}
	}
	double stopcount = (double)(fstop_hi - fstop_lo + 1);
	double mpsnr_denom = pixels * 3.0 * stopcount * 255.0 * 255.0;
	memset(out, 0, sizeof(*out));
	double psnr = num == 0.0 ? 999.0 : 10.0 * log10(samples / num);
	double rgb_psnr = psnr;
	out->psnr = psnr;
	out->alpha_psnr = psnr;
	if (mask & 8) {
		out->alpha_psnr = alpha_num == 0.0 ? 999.0 : 10.0 * log10(samples / alpha_num);
		double rgb_num = es[0] + es[1] + es[2];
		rgb_psnr = rgb_num == 0.0 ? 999.0 : 10.0 * log10(pixels * 3.0 / rgb_num);
	}
	out->rgb_psnr = rgb_psnr;
	out->rgb_peak = rgb_peak;
	if (hdr) {
		out->peak_psnr = rgb_psnr + 20.0 * log10(rgb_peak);
		out->mpsnr = mpsnr_num == 0.0 ? 999.0 : 10.0 * log10(mpsnr_denom / mpsnr_num);
		out->log_rmse = sqrt(log_num / pixels);
	}
	if (normal) {
		out->mean_angular_error = mean_ang;
		out->worst_angular_error = worst_ang;
	}
	for (int c = 0; c < 4; c++) {
		out->sum_squared_error[c] = es[c];
	}
	return 0;
}
