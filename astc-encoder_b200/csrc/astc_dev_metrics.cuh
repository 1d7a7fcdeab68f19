// B200-native ASTC codec: image error metrics on the device.
//
// What the reference's CLI computes on the host after a -t* round trip (astcenccli_error_metrics.cpp:109-413,
// compute_error_metrics): per-channel sums of squared differences, the same with RGB scaled by the original alpha,
// and for HDR images log2 / multi-exposure (mPSNR) differences, for normal maps the angular error. Per-texel values
// use the reference's float arithmetic (its polynomial log2, astcenc_vecmathlib.h:416-441; DPPS-ordered dot products);
// the sums are double accumulators there (:31-58) - here the image is reduced in parallel (thread -> warp -> CTA ->
// a second single-CTA pass in fixed order), so results are deterministic but differ from the reference's raster-order
// double sums in the last bits (tests/test_metrics.py states the tolerance).
#pragma once

#define ASTC_METRIC_SUMS 17
// [0..3] squared error r,g,b,a; [4..7] alpha-scaled; [8..10] log2 error r,g,b; [11..13] mPSNR error r,g,b;
// [14] sum of angular error / pixel count; [15] worst angular error (max); [16] rgb peak of image 1 (max)
enum { MS_ERR = 0, MS_AERR = 4, MS_LOG = 8, MS_MPSNR = 11, MS_ANG_MEAN = 14, MS_ANG_WORST = 15, MS_PEAK = 16 };

struct MetricImage {
	const void* data;
	int type;                // ASTCENC_TYPE_*
	unsigned int dim_x;      // row pitch in texels
};

struct MetricArgs {
	MetricImage img1, img2;
	unsigned int dim_x, dim_y;     // the compared intersection
	int hdr, normal;
	int fstop_lo, fstop_hi;
	double inv_pixels;
	double* partials;              // [gridDim.x][ASTC_METRIC_SUMS]
};

static __device__ __forceinline__ float4 metric_load(const MetricImage& im, unsigned int x, unsigned int y) {
	size_t i = 4 * ((size_t)im.dim_x * y + x);
	float4 c;
	if (im.type == 0) {
		uchar4 v = *reinterpret_cast<const uchar4*>(static_cast<const uint8_t*>(im.data) + i);
		c = make_float4((float)v.x / 255.0f, (float)v.y / 255.0f, (float)v.z / 255.0f, (float)v.w / 255.0f);
	} else if (im.type == 1) {
		ushort4 v = *reinterpret_cast<const ushort4*>(static_cast<const uint16_t*>(im.data) + i);
		c = make_float4(sf16_to_float(v.x), sf16_to_float(v.y), sf16_to_float(v.z), sf16_to_float(v.w));
	} else {
		c = *reinterpret_cast<const float4*>(static_cast<const float*>(im.data) + i);
	}
	if (im.type != 0) {
		// clamp(0, 65504, c) = min(max(c, 0), 65504) (:187, :200)
		c.x = fminf(fmaxf(c.x, 0.0f), 65504.0f);
		c.y = fminf(fmaxf(c.y, 0.0f), 65504.0f);
		c.z = fminf(fmaxf(c.z, 0.0f), 65504.0f);
		c.w = fminf(fmaxf(c.w, 0.0f), 65504.0f);
	}
	return c;
}

// the reference's approximate log2 (astcenc_vecmathlib.h:416-441), unfused
static __device__ __forceinline__ float metric_log2(float x) {
	int i = __float_as_int(x);
	float e = (float)(((i & 0x7F800000) >> 23) - 127);
	float m = __int_as_float((i & 0x007FFFFF) | 0x3F800000);
	float p = 0.0596515482674574969533f;
	p = p * m + -0.465725644288844778798f;
	p = p * m + 1.48116647521213171641f;
	p = p * m + -2.52074962577807006663f;
	p = p * m + 2.8882704548164776201f;
	p = p * (m - 1.0f);
	return p + e;
}

// mpsnr_operator / mpsnr_sumdiff (astcenccli_error_metrics.cpp:69-107)
static __device__ __forceinline__ float metric_mpsnr_operator(float val, int fstop) {
	float scale = __uint_as_float(0x3f800000u + ((unsigned int)fstop << 23));
	val = powf(val * scale, 1.0f / 2.2f);
	return fminf(fmaxf(val * 255.0f, 0.0f), 255.0f);
}
static __device__ __forceinline__ float metric_mpsnr_sumdiff(float v1, float v2, int lo, int hi) {
	float summa = 0.0f;
	for (int i = lo; i <= hi; i++) {
		float d = metric_mpsnr_operator(v1, i) - metric_mpsnr_operator(v2, i);
		summa += d * d;
	}
	return summa;
}

// (c - 0.5) * 2, normalize_safe(xyz, unit3) (:269-275; astcenc_vecmathlib.h:362-371), dot in DPPS order
static __device__ __forceinline__ float3 metric_normal(float4 c) {
	float x = (c.x - 0.5f) * 2.0f, y = (c.y - 0.5f) * 2.0f, z = (c.z - 0.5f) * 2.0f;
	float len = (x * x + y * y) + (z * z + 0.0f);
	if (len != 0.0f) {
		float s = sqrtf(len);
		return make_float3(x / s, y / s, z / s);
	}
	float u = 0.57735026918962576451f;      // unit3()
	return make_float3(u, u, u);
}

__global__ void __launch_bounds__(256)
astc_error_metrics_kernel(const __grid_constant__ MetricArgs a) {
	double acc[ASTC_METRIC_SUMS];
	for (int k = 0; k < ASTC_METRIC_SUMS; k++) {
		acc[k] = 0.0;
	}
	size_t n = (size_t)a.dim_x * a.dim_y;
	for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x) {
		unsigned int y = (unsigned int)(p / a.dim_x);
		unsigned int x = (unsigned int)(p - (size_t)y * a.dim_x);
		float4 c1 = metric_load(a.img1, x, y);
		float4 c2 = metric_load(a.img2, x, y);
		acc[MS_PEAK] = fmax(fmax(fmax((double)c1.x, (double)c1.y), (double)c1.z), acc[MS_PEAK]);
		float dx = c1.x - c2.x, dy = c1.y - c2.y, dz = c1.z - c2.z, dw = c1.w - c2.w;
		acc[MS_ERR + 0] += (double)(dx * dx);
		acc[MS_ERR + 1] += (double)(dy * dy);
		acc[MS_ERR + 2] += (double)(dz * dz);
		acc[MS_ERR + 3] += (double)(dw * dw);
		float ax = dx * c1.w, ay = dy * c1.w, az = dz * c1.w;
		acc[MS_AERR + 0] += (double)(ax * ax);
		acc[MS_AERR + 1] += (double)(ay * ay);
		acc[MS_AERR + 2] += (double)(az * az);
		acc[MS_AERR + 3] += (double)(dw * dw);
		if (a.hdr) {
			float lx = metric_log2(c1.x) - metric_log2(c2.x);
			float ly = metric_log2(c1.y) - metric_log2(c2.y);
			float lz = metric_log2(c1.z) - metric_log2(c2.z);
			acc[MS_LOG + 0] += (double)(lx * lx);
			acc[MS_LOG + 1] += (double)(ly * ly);
			acc[MS_LOG + 2] += (double)(lz * lz);
			acc[MS_MPSNR + 0] += (double)metric_mpsnr_sumdiff(c1.x, c2.x, a.fstop_lo, a.fstop_hi);
			acc[MS_MPSNR + 1] += (double)metric_mpsnr_sumdiff(c1.y, c2.y, a.fstop_lo, a.fstop_hi);
			acc[MS_MPSNR + 2] += (double)metric_mpsnr_sumdiff(c1.z, c2.z, a.fstop_lo, a.fstop_hi);
		}
		if (a.normal) {
			float3 n1 = metric_normal(c1);
			float3 n2 = metric_normal(c2);
			float d = (n1.x * n2.x + n1.y * n2.y) + (n1.z * n2.z + 0.0f);
			d = fminf(fmaxf(d, -1.0f), 1.0f);
			float rad_to_degrees = 180.0f / 3.14159265358979323846f;
			double deg = acos((double)d) * (double)rad_to_degrees;
			acc[MS_ANG_MEAN] += deg * a.inv_pixels;
			acc[MS_ANG_WORST] = fmax(acc[MS_ANG_WORST], deg);
		}
	}
	// thread -> warp -> CTA, fixed order
	__shared__ double red[8][ASTC_METRIC_SUMS];
	for (int k = 0; k < ASTC_METRIC_SUMS; k++) {
		double v = acc[k];
		bool is_max = k >= MS_ANG_WORST;
		for (int o = 16; o > 0; o >>= 1) {
			double t = __shfl_xor_sync(0xffffffffu, v, o);
			v = is_max ? fmax(v, t) : v + t;
		}
		if ((threadIdx.x & 31) == 0) {
			red[threadIdx.x >> 5][k] = v;
		}
	}
	__syncthreads();
	if (threadIdx.x < ASTC_METRIC_SUMS) {
		int k = threadIdx.x;
		bool is_max = k >= MS_ANG_WORST;
		double v = red[0][k];
		for (unsigned int wi = 1; wi < blockDim.x / 32; wi++) {
			v = is_max ? fmax(v, red[wi][k]) : v + red[wi][k];
		}
		a.partials[(size_t)blockIdx.x * ASTC_METRIC_SUMS + k] = v;
	}
}

// second pass: one thread per sum walks the CTA partials in order
__global__ void astc_error_metrics_finish_kernel(const double* partials, int ctas, double* out) {
	int k = threadIdx.x;
	if (k < ASTC_METRIC_SUMS) {
		bool is_max = k >= MS_ANG_WORST;
		double v = partials[k];
		for (int c = 1; c < ctas; c++) {
			double t = partials[(size_t)c * ASTC_METRIC_SUMS + k];
			v = is_max ? fmax(v, t) : v + t;
		}
		out[k] = v;
	}
}
