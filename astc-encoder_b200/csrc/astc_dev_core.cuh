// B200-native ASTC block compressor: warp-cooperative device code, part 1
// (platform layer, block load, ideal endpoints/weights, decimated weights, angular search, weight quantisation).
//
// Execution model: ONE WARP OWNS ONE IMAGE BLOCK for its whole search. All lanes follow the same
// control flow (the decision tree only depends on the block); lanes are spread over independent
// work items (texels, grid weights, decimation grids, block modes, accumulation chains, candidate
// partitionings). Every fp32 sum whose order is fixed by the reference is computed as a *chain*: one
// lane adds the terms in the reference's order; chains run side by side in different lanes.
// The per-warp working set lives in an arena (shared memory when it fits, else global memory).
//
// The same source compiles for the host with ASTC_HOSTSIM (1 "lane", used by tests/hostsim to check
// lane-independent logic against the oracle without a GPU). It is not a product path.
#pragma once

#include <stdint.h>
#include <math.h>
#include <string.h>
#include "astc_dev_tables.h"

#if defined(ASTC_HOSTSIM)
	#define ASTC_FN static inline
	#define ASTC_NOINLINE static
	#define ASTC_COOP static
	#define ASTC_WARP 1
	#define ASTC_RINT(a) nearbyintf(a)
	static inline uint32_t astc_f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
	static inline float astc_u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
	#define ASTC_F2U(f) astc_f2u(f)
	#define ASTC_U2F(u) astc_u2f(u)
	#define ASTC_CLZ(v) __builtin_clz(v)
	#define ASTC_POPCLL(v) __builtin_popcountll(v)
	#define ASTC_LDG(p) (*(p))
	static const DevConstTables* g_astc_ct;
	#define ASTC_CT g_astc_ct
#else
	#define ASTC_FN static __device__ __forceinline__
	#define ASTC_NOINLINE static __device__ __noinline__
	#define ASTC_COOP static __device__ __noinline__
	#if defined(ASTC_DEBUG_SINGLE_LANE)
		#define ASTC_WARP 1       /* debug build: lane 0 of each warp does all the work serially */
	#else
		#define ASTC_WARP 32
	#endif
	#define ASTC_RINT(a) rintf(a)
	#define ASTC_F2U(f) __float_as_uint(f)
	#define ASTC_U2F(u) __uint_as_float(u)
	#define ASTC_CLZ(v) __clz((int)(v))
	#define ASTC_POPCLL(v) __popcll(v)
	#define ASTC_LDG(p) __ldg(p)
	__constant__ const DevConstTables* g_astc_ct;
	#define ASTC_CT g_astc_ct
#endif

#include "astc_dev_math.cuh"
#include "astc_dev_color.cuh"

// Optional tracing of intermediate values (debug builds only: -DASTC_TRACE), printed by lane 0.
#if defined(ASTC_TRACE)
	#include <stdio.h>
	#define TRACE(...) do { if (w.lane == 0) printf(__VA_ARGS__); } while (0)
	#define TRACE_F(name, v) TRACE("%s %08x\n", name, ASTC_F2U(v))
#else
	#define TRACE(...) do { } while (0)
	#define TRACE_F(name, v) do { } while (0)
#endif

static const float ERROR_CALC_DEFAULT = 1e30f;
#define TUNE_MAX_ANGULAR_QUANT 7
#define TUNE_MAX_TRIAL_CANDIDATES 8

enum { FLG_MAP_NORMAL = 1, FLG_USE_DECODE_UNORM8 = 2, FLG_USE_ALPHA_WEIGHT = 4, FLG_USE_PERCEPTUAL = 8,
       FLG_DECOMPRESS_ONLY = 16, FLG_SELF_DECOMPRESS_ONLY = 32, FLG_MAP_RGBM = 64 };
enum { SYM_BTYPE_ERROR = 0, SYM_BTYPE_CONST_F16 = 1, SYM_BTYPE_CONST_U16 = 2, SYM_BTYPE_NONCONST = 3 };

// ---------------------------------------------------------------------------------------------
// Warp primitives
// ---------------------------------------------------------------------------------------------
#if defined(ASTC_HOSTSIM) || defined(ASTC_DEBUG_SINGLE_LANE)
ASTC_FN void wsync() {}
ASTC_FN float wmin_f(float v) { return v; }
ASTC_FN float wmax_f(float v) { return v; }
ASTC_FN bool wall(bool p) { return p; }
ASTC_FN bool wany(bool p) { return p; }
ASTC_FN void wargmin(float& err, int& idx) {}
ASTC_FN int wsame_key_rank(int key, int lane) { (void)key; (void)lane; return 0; }
ASTC_FN int wsame_key_count(int key) { (void)key; return 1; }
#else
ASTC_FN void wsync() { __syncwarp(); }
ASTC_FN float wmin_f(float v) {
	for (int o = 16; o > 0; o >>= 1) {
		float t = __shfl_xor_sync(0xffffffffu, v, o);
		v = t < v ? t : v;
	}
	return v;
}
ASTC_FN float wmax_f(float v) {
	for (int o = 16; o > 0; o >>= 1) {
		float t = __shfl_xor_sync(0xffffffffu, v, o);
		v = t > v ? t : v;
	}
	return v;
}
ASTC_FN bool wall(bool p) { return __all_sync(0xffffffffu, p) != 0; }
ASTC_FN bool wany(bool p) { return __any_sync(0xffffffffu, p) != 0; }
// lowest error, lowest index among equal errors; every lane receives the winner
ASTC_FN void wargmin(float& err, int& idx) {
	for (int o = 16; o > 0; o >>= 1) {
		float e2 = __shfl_xor_sync(0xffffffffu, err, o);
		int i2 = __shfl_xor_sync(0xffffffffu, idx, o);
		bool take = (e2 < err) || (e2 == err && (unsigned int)i2 < (unsigned int)idx);
		err = take ? e2 : err;
		idx = take ? i2 : idx;
	}
}
ASTC_FN int wsame_key_rank(int key, int lane) {
	unsigned int m = __match_any_sync(0xffffffffu, key);
	return __popc(m & ((1u << lane) - 1));
}
ASTC_FN int wsame_key_count(int key) {
	return __popc(__match_any_sync(0xffffffffu, key));
}
#endif

// ---------------------------------------------------------------------------------------------
// Per-warp context
// ---------------------------------------------------------------------------------------------
struct BlkInfo {
	f4 origin_texel, data_min, data_mean, data_max, channel_weight;
	bool grayscale, decode_unorm8;
	uint8_t rgb_lns0, alpha_lns0;
};

struct ScbHdr {            // scalar part of symbolic_compressed_block (arrays live in the arena)
	uint8_t block_type, partition_count, color_formats_matched;
	int8_t plane2_component;
	uint16_t block_mode, partition_index;
	uint8_t color_formats[4];
	uint8_t quant_mode;
	float errorval;
	int constant_color[4];
};

struct WCtx {
	int lane;
	const DevBsd* bsd;
	const DevConfig* cfg;
	int T;                 // texels per block
	float* blk[4];         // r, g, b, a            [T] each
	float* eiw[2];         // ideal weights, plane 1 / plane 2 fit
	float* eis[2];         // weight_error_scale
	bool ei_const_wes[2];
	f4* ep;                // endpoint slots, see EP_* below
	float* dwi;            // decimated ideal weights, packed per grid
	float* lowhigh;        // [decimation mode][plane][quant 0..7][low, high]
	float* mode_err;       // per packed block mode
	uint8_t* best_weights; // [64]   best-so-far symbolic block arrays
	uint8_t* best_colors;  // [4][8]
	uint8_t* work_weights; // [64]   candidate under refinement
	uint8_t* work_colors;  // [4][8]
	uint8_t* mod_colors;   // [4][8]
	float* tmpf;           // 128 floats for partial-sum exchange
	uint8_t* cand;         // candidate list
	uint8_t* su;           // big union scratch
	BlkInfo bi;
};

// endpoint slots in w.ep (f4 units)
enum { EP_EI1_0 = 0, EP_EI1_1 = 4, EP_EI2_0 = 8, EP_EI2_1 = 12, EP_WORK_0 = 16, EP_WORK_1 = 20, EP_RGBS = 24, EP_RGBO = 28, EP_BASE_0 = 32, EP_BASE_1 = 36, EP_COUNT = 40 };

ASTC_FN float cw_lane(const WCtx& w, int c) { return lane(w.bi.channel_weight, c); }
ASTC_FN f4 texel4(const WCtx& w, int i) { return mk4(w.blk[0][i], w.blk[1][i], w.blk[2][i], w.blk[3][i]); }
ASTC_FN float default_alpha(const WCtx& w) { return w.bi.alpha_lns0 ? static_cast<float>(0x7800) : static_cast<float>(0xFFFF); }
ASTC_FN bool is_constant_channel(const WCtx& w, int ch) { return lane(w.bi.data_min, ch) == lane(w.bi.data_max, ch); }
ASTC_FN bool is_luminance(const WCtx& w) {
	float da = default_alpha(w);
	bool alpha1 = (w.bi.data_min.w == da) && (w.bi.data_max.w == da);
	return w.bi.grayscale && alpha1;
}
ASTC_FN bool is_luminancealpha(const WCtx& w) {
	float da = default_alpha(w);
	bool alpha1 = (w.bi.data_min.w == da) && (w.bi.data_max.w == da);
	return w.bi.grayscale && !alpha1;
}

// ---------------------------------------------------------------------------------------------
// Table accessors
// ---------------------------------------------------------------------------------------------
struct DecView {
	const DevDecMode* dm;
	const uint8_t* tw;        // [4][T]
	const uint8_t* tc;        // [4][T]
	const uint16_t* wto;      // [W + 1]
	const uint8_t* wt;
	const uint8_t* wc;
	int T, W, max_twc;
};

ASTC_FN DecView dec_view(const DevBsd& bsd, unsigned int d) {
	DecView v;
	const DevDecMode* dm = bsd.dec_modes + d;
	const uint8_t* blob = bsd.dec_blob + dm->blob_offset;
	v.dm = dm;
	v.T = bsd.texel_count;
	v.W = dm->weight_count;
	v.max_twc = dm->max_texel_weight_count;
	v.tw = blob;
	v.tc = blob + 4 * v.T;
	v.wto = reinterpret_cast<const uint16_t*>(blob + dm->wto_offset);
	v.wt = blob + dm->wt_offset;
	v.wc = blob + dm->wc_offset;
	return v;
}

struct PartView {
	const uint8_t* base;
	const uint8_t* partition_of_texel;
	const uint8_t* texels;    // concatenated texels_of_partition
	unsigned int partition_count;
	unsigned int partition_index;
	uint8_t count[4];
	uint8_t start[4];
};

ASTC_FN PartView part_view_packed(const DevBsd& bsd, unsigned int pc, unsigned int packed) {
	PartView v;
	const uint8_t* e = bsd.partitions[pc] + (size_t)packed * bsd.part_stride;
	v.base = e;
	v.partition_count = pc;
	v.partition_index = (unsigned int)e[0] | ((unsigned int)e[1] << 8);
	unsigned int s = 0;
	for (int i = 0; i < 4; i++) {
		v.count[i] = e[2 + i];
		v.start[i] = (uint8_t)s;
		s += e[2 + i];
	}
	v.partition_of_texel = e + ASTC_PART_HDR;
	v.texels = e + ASTC_PART_HDR + bsd.texel_count;
	return v;
}

ASTC_FN PartView part_view(const DevBsd& bsd, unsigned int pc, unsigned int partition_index) {
	unsigned int packed = pc >= 2 ? bsd.partitioning_packed_index[pc - 2][partition_index] : 0;
	return part_view_packed(bsd, pc, packed);
}

ASTC_FN unsigned int quant_level_count(int q) {
	const uint16_t levels[21] = {2, 3, 4, 5, 6, 8, 10, 12, 16, 20, 24, 32, 40, 48, 64, 80, 96, 128, 160, 192, 256};
	return levels[q];
}

// bits / trits / quints of a BISE level (astcenc_integer_sequence.cpp:301-327)
ASTC_FN void ise_btq(int q, unsigned int& bits, unsigned int& trits, unsigned int& quints) {
	const uint8_t b[21] = {1, 0, 2, 0, 1, 3, 1, 2, 4, 2, 3, 5, 3, 4, 6, 4, 5, 7, 5, 6, 8};
	const uint8_t t[21] = {0, 1, 0, 0, 1, 0, 0, 1, 0, 0, 1, 0, 0, 1, 0, 0, 1, 0, 0, 1, 0};
	const uint8_t u[21] = {0, 0, 0, 1, 0, 0, 1, 0, 0, 1, 0, 0, 1, 0, 0, 1, 0, 0, 1, 0, 0};
	bits = b[q];
	trits = t[q];
	quints = u[q];
}

ASTC_FN unsigned int ise_sequence_bitcount(unsigned int count, int q) {   // :419-435
	unsigned int bits, trits, quints;
	ise_btq(q, bits, trits, quints);
	if (trits) return ((8 + 5 * bits) * count + 4) / 5;
	if (quints) return ((7 + 3 * bits) * count + 2) / 3;
	return bits * count;
}

// =============================================================================================
// Block load (astcenc_image.cpp:162-342). Lanes over texels; the per-channel mean is a chain.
// =============================================================================================
ASTC_COOP void load_block(WCtx& w, const DevImage& img, unsigned int pos_x, unsigned int pos_y) {
	const DevBsd& bsd = *w.bsd;
	const DevConfig& cfg = *w.cfg;
	int profile = cfg.profile;
	bool needs_swz = img.swz[0] != 0 || img.swz[1] != 1 || img.swz[2] != 2 || img.swz[3] != 3;
	bool needs_hdr = profile == PRF_HDR || profile == PRF_HDR_RGB_LDR_A;
	bool fast = !needs_swz && !needs_hdr && img.data_type == 0;
	uint8_t rgb_lns = needs_hdr ? 1 : 0;
	uint8_t a_lns = profile == PRF_HDR ? 1 : 0;
	int T = w.T;
	unsigned int bx = bsd.dim_x;

	f4 dmin = splat4(1e38f), dmax = splat4(-1e38f);
	bool gray = true;
	for (int t = w.lane; t < T; t += ASTC_WARP) {
		unsigned int x = pos_x + (unsigned int)t % bx;
		unsigned int y = pos_y + (unsigned int)t / bx;
		unsigned int xi = x < img.dim_x - 1 ? x : img.dim_x - 1;
		unsigned int yi = y < img.dim_y - 1 ? y : img.dim_y - 1;
		size_t off = (4 * (size_t)img.dim_x * yi) + (4 * xi);
		f4 v;
		if (fast) {
			uint32_t px = ASTC_LDG(reinterpret_cast<const uint32_t*>(static_cast<const uint8_t*>(img.data) + off));
			v = mk4(static_cast<float>(px & 0xFF), static_cast<float>((px >> 8) & 0xFF), static_cast<float>((px >> 16) & 0xFF),
			        static_cast<float>(px >> 24)) * (65535.0f / 255.0f);
		} else {
			if (img.data_type == 0) {
				uint32_t px = ASTC_LDG(reinterpret_cast<const uint32_t*>(static_cast<const uint8_t*>(img.data) + off));
				v = mk4(static_cast<float>(px & 0xFF), static_cast<float>((px >> 8) & 0xFF), static_cast<float>((px >> 16) & 0xFF),
				        static_cast<float>(px >> 24)) / 255.0f;
			} else if (img.data_type == 1) {
				const uint16_t* p = static_cast<const uint16_t*>(img.data) + off;
				v = mk4(sf16_to_float(p[0]), sf16_to_float(p[1]), sf16_to_float(p[2]), sf16_to_float(p[3]));
			} else {
				const float* p = static_cast<const float*>(img.data) + off;
				v = mk4(p[0], p[1], p[2], p[3]);
			}
			if (needs_swz) {
				float s0 = img.swz[0] < 4 ? lane(v, img.swz[0]) : (img.swz[0] == 4 ? 0.0f : 1.0f);
				float s1 = img.swz[1] < 4 ? lane(v, img.swz[1]) : (img.swz[1] == 4 ? 0.0f : 1.0f);
				float s2 = img.swz[2] < 4 ? lane(v, img.swz[2]) : (img.swz[2] == 4 ? 0.0f : 1.0f);
				float s3 = img.swz[3] < 4 ? lane(v, img.swz[3]) : (img.swz[3] == 4 ? 0.0f : 1.0f);
				v = mk4(s0, s1, s2, s3);
			}
			f4 un = vclamp4(0.0f, 65535.0f, v * 65535.0f);
			if (rgb_lns || a_lns) {
				f4 l = mk4(float_to_lns(v.x), float_to_lns(v.y), float_to_lns(v.z), float_to_lns(v.w));
				v = mk4(rgb_lns ? l.x : un.x, rgb_lns ? l.y : un.y, rgb_lns ? l.z : un.z, a_lns ? l.w : un.w);
			} else {
				v = un;
			}
		}
		dmin = min4(dmin, v);
		dmax = max4(dmax, v);
		gray = gray && (v.x == v.y) && (v.x == v.z);
		w.blk[0][t] = v.x;
		w.blk[1][t] = v.y;
		w.blk[2][t] = v.z;
		w.blk[3][t] = v.w;
	}
	// block data never holds NaN (the clamps and float_to_lns flush it), so min/max are order independent
	dmin = mk4(wmin_f(dmin.x), wmin_f(dmin.y), wmin_f(dmin.z), wmin_f(dmin.w));
	dmax = mk4(wmax_f(dmax.x), wmax_f(dmax.y), wmax_f(dmax.z), wmax_f(dmax.w));
	gray = wall(gray);
	wsync();
	// per-channel mean: chain in texel order
	float mean_scale = 1.0f / static_cast<float>(T);
	for (int c = w.lane; c < 4; c += ASTC_WARP) {
		float s = 0.0f;
		const float* d = w.blk[c];
		if (fast) {
			for (int t = 0; t < T; t++) s = s + d[t];
			s = s / static_cast<float>(T);
		} else {
			for (int t = 0; t < T; t++) s = s + d[t] * mean_scale;
		}
		w.tmpf[c] = s;
	}
	wsync();
	BlkInfo& bi = w.bi;
	bi.data_mean = mk4(w.tmpf[0], w.tmpf[1], w.tmpf[2], w.tmpf[3]);
	bi.data_min = dmin;
	bi.data_max = dmax;
	bi.grayscale = gray;
	bi.decode_unorm8 = (cfg.flags & FLG_USE_DECODE_UNORM8) != 0;
	bi.channel_weight = mk4(cfg.cw[0], cfg.cw[1], cfg.cw[2], cfg.cw[3]);
	f4 enc = texel4(w, 0);
	if (fast) {
		bi.origin_texel = enc / 65535.0f;
		bi.rgb_lns0 = 0;
		bi.alpha_lns0 = 0;
	} else {
		f4 enc_unorm = enc / 65535.0f;
		f4 enc_lns = splat4(0.0f);
		if (rgb_lns || a_lns) {
			enc_lns = mk4(sf16_to_float((uint16_t)lns_to_sf16(f2i(enc.x))), sf16_to_float((uint16_t)lns_to_sf16(f2i(enc.y))),
			              sf16_to_float((uint16_t)lns_to_sf16(f2i(enc.z))), sf16_to_float((uint16_t)lns_to_sf16(f2i(enc.w))));
		}
		bi.origin_texel = mk4(rgb_lns ? enc_lns.x : enc_unorm.x, rgb_lns ? enc_lns.y : enc_unorm.y,
		                      rgb_lns ? enc_lns.z : enc_unorm.z, a_lns ? enc_lns.w : enc_unorm.w);
		bi.rgb_lns0 = rgb_lns;
		bi.alpha_lns0 = a_lns;
	}
	if (cfg.flags & FLG_USE_ALPHA_WEIGHT) {   // astcenc_entry.cpp:1017-1024
		float alpha_scale = bi.data_max.w * (1.0f / 65535.0f);
		bi.channel_weight = mk4(cfg.cw[0] * alpha_scale, cfg.cw[1] * alpha_scale, cfg.cw[2] * alpha_scale, cfg.cw[3]);
	}
	wsync();
}

// =============================================================================================
// Averages and directions (astcenc_averages_and_directions.cpp:47-720)
// =============================================================================================
struct PartitionMetrics {
	f4 avg;
	f4 dir;
};

// Partition means. For >= 2 partitions the reference accumulates with masked haccumulate: texel i adds
// to lane (i mod 4) of its partition's accumulator -> one chain per (partition, channel, i mod 4).
// ncomp = 4 (rgba) or 3 (rgb, lane 3 zero).
ASTC_COOP void compute_partition_averages(WCtx& w, const PartView& pi, int ncomp, f4 averages[4]) {
	unsigned int pc = pi.partition_count;
	f4 mean = ncomp == 4 ? w.bi.data_mean : mk4(w.bi.data_mean.x, w.bi.data_mean.y, w.bi.data_mean.z, 0.0f);
	if (pc == 1) {
		averages[0] = mean;
		return;
	}
	int T = w.T;
	int nchains = (int)(pc - 1) * ncomp * 4;
	for (int id = w.lane; id < nchains; id += ASTC_WARP) {
		int l = id & 3;
		int c = (id >> 2) % ncomp;
		unsigned int p = (unsigned int)((id >> 2) / ncomp);
		const float* d = w.blk[c];
		float s = 0.0f;
		for (int i = l; i < T; i += 4) {
			if (pi.partition_of_texel[i] == p) {
				s = s + d[i];
			}
		}
		w.tmpf[id] = s;
	}
	wsync();
	f4 block_total = mean * static_cast<float>(T);
	f4 rest = block_total;
	for (unsigned int p = 0; p < pc - 1; p++) {
		f4 total = splat4(0.0f);
		for (int c = 0; c < ncomp; c++) {
			const float* a = w.tmpf + ((int)p * ncomp + c) * 4;
			set_lane(total, c, (a[0] + a[2]) + (a[1] + a[3]));
		}
		rest = rest - total;
		averages[p] = total / static_cast<float>(pi.count[p]);
	}
	averages[pc - 1] = rest / static_cast<float>(pi.count[pc - 1]);
	wsync();
}

// Sign-split direction sums: one chain per (partition, split axis K, component c), in partition-texel order.
ASTC_COOP void compute_dirs(WCtx& w, const PartView& pi, const float* c0, const float* c1, const float* c2, const float* c3, int ncomp,
                            const f4 averages[4], PartitionMetrics pm[4]) {
	unsigned int pc = pi.partition_count;
	const float* chan[4] = {c0, c1, c2, c3};
	int per_part = ncomp * ncomp;
	int nchains = (int)pc * per_part;
	for (int id = w.lane; id < nchains; id += ASTC_WARP) {
		int p = id / per_part;
		int r = id - p * per_part;
		int K = r / ncomp;
		int c = r - K * ncomp;
		float avgK = lane(averages[p], K);
		float avgc = lane(averages[p], c);
		const float* dK = chan[K];
		const float* dc = chan[c];
		const uint8_t* tix = pi.texels + pi.start[p];
		int n = pi.count[p];
		float s = 0.0f;
		for (int i = 0; i < n; i++) {
			int t = tix[i];
			float vK = dK[t] - avgK;
			float vc = dc[t] - avgc;
			s = s + (vK > 0.0f ? vc : 0.0f);
		}
		w.tmpf[id] = s;
	}
	wsync();
	for (unsigned int p = 0; p < pc; p++) {
		const float* a = w.tmpf + (int)p * per_part;
		f4 sums[4];
		for (int K = 0; K < 4; K++) {
			sums[K] = splat4(0.0f);
		}
		for (int K = 0; K < ncomp; K++) {
			for (int c = 0; c < ncomp; c++) {
				set_lane(sums[K], c, a[K * ncomp + c]);
			}
		}
		f4 best_vector = sums[0];
		float best_sum = dot_s(sums[0], sums[0]);
		for (int K = 1; K < ncomp; K++) {
			float prod = dot_s(sums[K], sums[K]);
			if (prod > best_sum) {
				best_vector = sums[K];
				best_sum = prod;
			}
		}
		pm[p].avg = averages[p];
		pm[p].dir = best_vector;
	}
	wsync();
}

ASTC_COOP void compute_avgs_and_dirs_4_comp(WCtx& w, const PartView& pi, PartitionMetrics pm[4]) {
	f4 averages[4];
	compute_partition_averages(w, pi, 4, averages);
	compute_dirs(w, pi, w.blk[0], w.blk[1], w.blk[2], w.blk[3], 4, averages, pm);
}

ASTC_COOP void compute_avgs_and_dirs_3_comp(WCtx& w, const PartView& pi, unsigned int omitted, PartitionMetrics pm[4]) {
	f4 averages[4] = {splat4(0.0f), splat4(0.0f), splat4(0.0f), splat4(0.0f)};
	compute_partition_averages(w, pi, 4, averages);
	const float* vr = w.blk[0];
	const float* vg = w.blk[1];
	const float* vb = w.blk[2];
	for (int i = 0; i < 4; i++) {
		f4 a = averages[i];
		if (omitted == 0) averages[i] = mk4(a.y, a.z, a.w, 0.0f);
		else if (omitted == 1) averages[i] = mk4(a.x, a.z, a.w, 0.0f);
		else if (omitted == 2) averages[i] = mk4(a.x, a.y, a.w, 0.0f);
		else averages[i] = mk4(a.x, a.y, a.z, 0.0f);
	}
	if (omitted == 0) {
		vr = w.blk[1];
		vg = w.blk[2];
		vb = w.blk[3];
	} else if (omitted == 1) {
		vg = w.blk[2];
		vb = w.blk[3];
	} else if (omitted == 2) {
		vb = w.blk[3];
	}
	compute_dirs(w, pi, vr, vg, vb, vb, 3, averages, pm);
}

ASTC_COOP void compute_avgs_and_dirs_3_comp_rgb(WCtx& w, const PartView& pi, PartitionMetrics pm[4]) {
	f4 averages[4];
	compute_partition_averages(w, pi, 3, averages);
	compute_dirs(w, pi, w.blk[0], w.blk[1], w.blk[2], w.blk[2], 3, averages, pm);
}

ASTC_COOP void compute_avgs_and_dirs_2_comp(WCtx& w, const PartView& pt, unsigned int comp1, unsigned int comp2, PartitionMetrics pm[4]) {
	const float* vr = w.blk[comp1];
	const float* vg = w.blk[comp2];
	f4 averages[4];
	unsigned int pc = pt.partition_count;
	// only ever called with the single-partition table (2-plane trials), keep the general form anyway
	for (unsigned int p = 0; p < pc; p++) {
		f4 average = mk4(lane(w.bi.data_mean, (int)comp1), lane(w.bi.data_mean, (int)comp2), 0.0f, 0.0f);
		if (pc > 1) {
			average = splat4(0.0f);
			unsigned int n = pt.count[p];
			for (unsigned int i = 0; i < n; i++) {
				unsigned int iwt = pt.texels[pt.start[p] + i];
				average = average + mk4(vr[iwt], vg[iwt], 0.0f, 0.0f);
			}
			average = average / static_cast<float>(n);
		}
		averages[p] = average;
	}
	compute_dirs(w, pt, vr, vg, vg, vg, 2, averages, pm);
}

// =============================================================================================
// Ideal endpoints and weights (astcenc_ideal_endpoints_and_weights.cpp:107-683)
// =============================================================================================
// One-component fit (:107-206). which = 0/1 selects the ei slot, epslot the endpoint slot pair.
ASTC_COOP void compute_ideal_colors_and_weights_1_comp(WCtx& w, const PartView& pi, int which, int ep0slot, int ep1slot, unsigned int component) {
	unsigned int pc = pi.partition_count;
	const float* data_vr = w.blk[component];
	float error_weight = cw_lane(w, (int)component);
	float* weights = w.eiw[which];
	float* wes = w.eis[which];
	bool is_constant_wes = true;
	float partition0_len_sq = 0.0f;
	for (unsigned int i = 0; i < pc; i++) {
		const uint8_t* tix = pi.texels + pi.start[i];
		int n = pi.count[i];
		float lowvalue = 1e10f, highvalue = -1e10f;
		for (int j = w.lane; j < n; j += ASTC_WARP) {
			float value = data_vr[tix[j]];
			lowvalue = minf(value, lowvalue);
			highvalue = maxf(value, highvalue);
		}
		lowvalue = wmin_f(lowvalue);
		highvalue = wmax_f(highvalue);
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
		for (int j = w.lane; j < n; j += ASTC_WARP) {
			int t = tix[j];
			float value = (data_vr[t] - lowvalue) * scale;
			value = clamp1f(value);
			weights[t] = value;
			wes[t] = length_squared * error_weight;
		}
		f4 e0 = w.bi.data_min, e1 = w.bi.data_max;
		set_lane(e0, (int)component, lowvalue);
		set_lane(e1, (int)component, highvalue);
		if (w.lane == 0) {
			w.ep[ep0slot + i] = e0;
			w.ep[ep1slot + i] = e1;
		}
	}
	w.ei_const_wes[which] = is_constant_wes;
	wsync();
}

// Projection on the partition lines (shared tail of the 2/3/4 component fits).
ASTC_COOP void ideal_project(WCtx& w, const PartView& pi, int which, const PartitionMetrics pms[4], int ncomp,
                             const float* c0, const float* c1, const float* c2, const float* c3, float error_weight, f4 lowv[4], f4 highv[4]) {
	unsigned int pc = pi.partition_count;
	float* weights = w.eiw[which];
	float* wes = w.eis[which];
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
		const uint8_t* tix = pi.texels + pi.start[i];
		int n = pi.count[i];
		float lowparam = 1e10f, highparam = -1e10f;
		for (int j = w.lane; j < n; j += ASTC_WARP) {
			int t = tix[j];
			f4 point = mk4(c0[t], c1[t], ncomp > 2 ? c2[t] : 0.0f, ncomp > 3 ? c3[t] : 0.0f);
			float param = ncomp == 3 ? dot3_s(point - la, lb) : dot_s(point - la, lb);
			weights[t] = param;
			lowparam = minf(param, lowparam);
			highparam = maxf(param, highparam);
		}
		// params may be NaN only if the block data is, which load_block excludes -> order independent
		lowparam = wmin_f(lowparam);
		highparam = wmax_f(highparam);
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
		wsync();
		for (int j = w.lane; j < n; j += ASTC_WARP) {
			int t = tix[j];
			float idx = (weights[t] - lowparam) * scale;
			idx = clamp1f(idx);
			weights[t] = idx;
			wes[t] = length_squared * error_weight;
		}
		lowv[i] = la + lb * lowparam;
		highv[i] = la + lb * highparam;
	}
	w.ei_const_wes[which] = is_constant_wes;
	wsync();
}

ASTC_COOP void compute_ideal_colors_and_weights_2_comp(WCtx& w, const PartView& pi, int which, int ep0slot, int ep1slot, int comp1, int comp2) {   // :217-351
	f4 cw = w.bi.channel_weight;
	float error_weight = ((lane(cw, comp1) + 0.0f) + (lane(cw, comp2) + 0.0f)) / 2.0f;
	PartitionMetrics pms[4];
	compute_avgs_and_dirs_2_comp(w, pi, (unsigned int)comp1, (unsigned int)comp2, pms);
	f4 lowv[4], highv[4];
	ideal_project(w, pi, which, pms, 2, w.blk[comp1], w.blk[comp2], w.blk[comp2], w.blk[comp2], error_weight, lowv, highv);
	for (unsigned int i = 0; i < pi.partition_count; i++) {
		f4 e0 = w.bi.data_min, e1 = w.bi.data_max;
		set_lane(e0, comp1, lowv[i].x);
		set_lane(e1, comp1, highv[i].x);
		set_lane(e0, comp2, lowv[i].y);
		set_lane(e1, comp2, highv[i].y);
		if (w.lane == 0) {
			w.ep[ep0slot + i] = e0;
			w.ep[ep1slot + i] = e1;
		}
	}
	wsync();
}

ASTC_COOP void compute_ideal_colors_and_weights_3_comp(WCtx& w, const PartView& pi, int which, int ep0slot, int ep1slot, unsigned int omitted) {   // :354-517
	f4 cw = w.bi.channel_weight;
	const float *vr, *vg, *vb;
	float error_weight;
	if (omitted == 0) {
		error_weight = (cw.x + cw.z) + (cw.y + 0.0f);
		vr = w.blk[1]; vg = w.blk[2]; vb = w.blk[3];
	} else if (omitted == 1) {
		error_weight = (cw.x + cw.w) + (cw.z + 0.0f);
		vr = w.blk[0]; vg = w.blk[2]; vb = w.blk[3];
	} else if (omitted == 2) {
		error_weight = (cw.x + cw.w) + (cw.y + 0.0f);
		vr = w.blk[0]; vg = w.blk[1]; vb = w.blk[3];
	} else {
		error_weight = (cw.x + cw.z) + (cw.y + 0.0f);
		vr = w.blk[0]; vg = w.blk[1]; vb = w.blk[2];
	}
	error_weight = error_weight * (1.0f / 3.0f);
	PartitionMetrics pms[4];
	if (omitted == 3) {
		compute_avgs_and_dirs_3_comp_rgb(w, pi, pms);
	} else {
		compute_avgs_and_dirs_3_comp(w, pi, omitted, pms);
	}
	f4 lowv[4], highv[4];
	ideal_project(w, pi, which, pms, 3, vr, vg, vb, vb, error_weight, lowv, highv);
	for (unsigned int i = 0; i < pi.partition_count; i++) {
		f4 e0 = lowv[i], e1 = highv[i];
		f4 bmin = w.bi.data_min, bmax = w.bi.data_max;
		f4 r0, r1;
		switch (omitted) {
		case 0: r0 = mk4(bmin.x, e0.x, e0.y, e0.z); r1 = mk4(bmax.x, e1.x, e1.y, e1.z); break;
		case 1: r0 = mk4(e0.x, bmin.y, e0.y, e0.z); r1 = mk4(e1.x, bmax.y, e1.y, e1.z); break;
		case 2: r0 = mk4(e0.x, e0.y, bmin.z, e0.z); r1 = mk4(e1.x, e1.y, bmax.z, e1.z); break;
		default: r0 = mk4(e0.x, e0.y, e0.z, bmin.w); r1 = mk4(e1.x, e1.y, e1.z, bmax.w); break;
		}
		if (w.lane == 0) {
			w.ep[ep0slot + i] = r0;
			w.ep[ep1slot + i] = r1;
		}
	}
	wsync();
}

ASTC_COOP void compute_ideal_colors_and_weights_4_comp(WCtx& w, const PartView& pi, int which, int ep0slot, int ep1slot) {   // :520-609
	float error_weight = hadd_s(w.bi.channel_weight) / 4.0f;
	PartitionMetrics pms[4];
	compute_avgs_and_dirs_4_comp(w, pi, pms);
	f4 lowv[4], highv[4];
	ideal_project(w, pi, which, pms, 4, w.blk[0], w.blk[1], w.blk[2], w.blk[3], error_weight, lowv, highv);
	for (unsigned int i = 0; i < pi.partition_count; i++) {
		if (w.lane == 0) {
			w.ep[ep0slot + i] = lowv[i];
			w.ep[ep1slot + i] = highv[i];
		}
	}
	wsync();
}

ASTC_COOP void compute_ideal_colors_and_weights_1plane(WCtx& w, const PartView& pi) {   // :612-627
	bool uses_alpha = !is_constant_channel(w, 3);
	if (uses_alpha) {
		compute_ideal_colors_and_weights_4_comp(w, pi, 0, EP_EI1_0, EP_EI1_1);
	} else {
		compute_ideal_colors_and_weights_3_comp(w, pi, 0, EP_EI1_0, EP_EI1_1, 3);
	}
}

ASTC_COOP void compute_ideal_colors_and_weights_2planes(WCtx& w, unsigned int plane2_component) {   // :630-683
	PartView pi = part_view_packed(*w.bsd, 1, 0);
	bool uses_alpha = !is_constant_channel(w, 3);
	switch (plane2_component) {
	case 0:
		if (uses_alpha) compute_ideal_colors_and_weights_3_comp(w, pi, 0, EP_EI1_0, EP_EI1_1, 0);
		else compute_ideal_colors_and_weights_2_comp(w, pi, 0, EP_EI1_0, EP_EI1_1, 1, 2);
		break;
	case 1:
		if (uses_alpha) compute_ideal_colors_and_weights_3_comp(w, pi, 0, EP_EI1_0, EP_EI1_1, 1);
		else compute_ideal_colors_and_weights_2_comp(w, pi, 0, EP_EI1_0, EP_EI1_1, 0, 2);
		break;
	case 2:
		if (uses_alpha) compute_ideal_colors_and_weights_3_comp(w, pi, 0, EP_EI1_0, EP_EI1_1, 2);
		else compute_ideal_colors_and_weights_2_comp(w, pi, 0, EP_EI1_0, EP_EI1_1, 0, 1);
		break;
	default:
		compute_ideal_colors_and_weights_3_comp(w, pi, 0, EP_EI1_0, EP_EI1_1, 3);
		break;
	}
	compute_ideal_colors_and_weights_1_comp(w, pi, 1, EP_EI2_0, EP_EI2_1, plane2_component);
}

// bilinear infill (:38-104): (w0*c0 + w1*c1) + (w2*c2 + w3*c3); contributions are exact multiples of 1/16
ASTC_FN float contrib_f(uint8_t c) { return static_cast<float>(c) * (1.0f / 16.0f); }

ASTC_FN float bilinear_infill(const DecView& di, const float* weights, int t) {
	int T = di.T;
	return (weights[di.tw[t]] * contrib_f(di.tc[t]) + weights[di.tw[T + t]] * contrib_f(di.tc[T + t])) +
	       (weights[di.tw[2 * T + t]] * contrib_f(di.tc[2 * T + t]) + weights[di.tw[3 * T + t]] * contrib_f(di.tc[3 * T + t]));
}
ASTC_FN float bilinear_infill_2(const DecView& di, const float* weights, int t) {
	int T = di.T;
	return (weights[di.tw[t]] * contrib_f(di.tc[t]) + weights[di.tw[T + t]] * contrib_f(di.tc[T + t]));
}

// compute_ideal_weights_for_decimation (:845-971) for one grid; nplanes = 1 or 2 (second plane: ei slot 1,
// output at out + plane2_off). Lanes over grid weights / texels; the per-weight sums are chains.
ASTC_COOP void compute_ideal_weights_for_decimation(WCtx& w, const DecView& di, int nplanes, float* out, int plane2_off) {
	int T = di.T;
	int W = di.W;
	if (T == W) {
		for (int i = w.lane; i < T * nplanes; i += ASTC_WARP) {
			int pl = i >= T ? 1 : 0;
			int t = i - pl * T;
			out[pl * plane2_off + t] = w.eiw[pl][t];
		}
		wsync();
		return;
	}
	float* infilled = reinterpret_cast<float*>(w.su);   // [nplanes][T]
	for (int id = w.lane; id < W * nplanes; id += ASTC_WARP) {
		int pl = id >= W ? 1 : 0;
		int i = id - pl * W;
		const float* eiw = w.eiw[pl];
		const float* eis = w.eis[pl];
		bool constant_wes = w.ei_const_wes[pl];
		float wes0 = eis[0];
		float weight_weight = 1e-10f;
		float initial_weight = 0.0f;
		int off = di.wto[i];
		int end = di.wto[i + 1];
		for (int j = off; j < end; j++) {
			int texel = di.wt[j];
			float weight = static_cast<float>(di.wc[j]);
			float wes = constant_wes ? wes0 : eis[texel];
			float contrib_weight = weight * wes;
			weight_weight += contrib_weight;
			initial_weight += eiw[texel] * contrib_weight;
		}
		out[pl * plane2_off + i] = initial_weight / weight_weight;
	}
	wsync();
	for (int id = w.lane; id < T * nplanes; id += ASTC_WARP) {
		int pl = id >= T ? 1 : 0;
		int t = id - pl * T;
		const float* src = out + pl * plane2_off;
		infilled[id] = di.max_twc <= 2 ? bilinear_infill_2(di, src, t) : bilinear_infill(di, src, t);
	}
	wsync();
	const float stepsize = 0.25f;
	const float chd_scale = -16.0f;
	for (int id = w.lane; id < W * nplanes; id += ASTC_WARP) {
		int pl = id >= W ? 1 : 0;
		int i = id - pl * W;
		const float* eiw = w.eiw[pl];
		const float* eis = w.eis[pl];
		const float* inf = infilled + pl * T;
		bool constant_wes = w.ei_const_wes[pl];
		float wes0 = eis[0];
		float weight_val = out[pl * plane2_off + i];
		float error_change0 = 1e-10f;
		float error_change1 = 0.0f;
		int off = di.wto[i];
		int end = di.wto[i + 1];
		for (int j = off; j < end; j++) {
			int texel = di.wt[j];
			float contrib_weight = static_cast<float>(di.wc[j]);
			float wes = constant_wes ? wes0 : eis[texel];
			float scale = wes * contrib_weight;
			float old_weight = inf[texel];
			float ideal_weight = eiw[texel];
			error_change0 += contrib_weight * scale;
			error_change1 += (old_weight - ideal_weight) * scale;
		}
		float step = (error_change1 * chd_scale) / error_change0;
		step = vclampf(-stepsize, stepsize, step);
		out[pl * plane2_off + i] = weight_val + step;
	}
	wsync();
}

// =============================================================================================
// Angular weight-range search (astcenc_weight_align.cpp:94-355). One lane owns one (grid, plane) and walks
// all its angular steps; every sum over the grid's weights is a chain.
// =============================================================================================
ASTC_NOINLINE void compute_angular_endpoints_for_quant_levels(int weight_count, const float* dwi, unsigned int max_quant_level, float* lowhigh /* [8][2] */) {
	const uint8_t steps_for_quant_level[12] = {2, 3, 4, 5, 6, 8, 10, 12, 16, 20, 24, 32};
	const DevConstTables* ct = ASTC_CT;
	int max_quant_steps = steps_for_quant_level[max_quant_level];
	int max_angular_steps = max_quant_steps;

	float min_weight = 3.402823466e+38f, max_weight = -3.402823466e+38f;
	for (int i = 0; i < weight_count; i++) {
		float v = dwi[i];
		min_weight = minf(v, min_weight);
		max_weight = maxf(v, max_weight);
	}
	float angular_offsets[ASTC_ANGULAR_STEPS];
	float lowest_weight[ASTC_ANGULAR_STEPS];
	float best_err[ASTC_ANGULAR_STEPS + 4], best_idx[ASTC_ANGULAR_STEPS + 4], best_cut[ASTC_ANGULAR_STEPS + 4];
	for (int i = 0; i < max_quant_steps + 4; i++) {
		best_err[i] = ERROR_CALC_DEFAULT;
		best_idx[i] = -1.0f;
		best_cut[i] = 0.0f;
	}
	const float mult = 1.0f / (2.0f * 3.14159265358979323846f);
	for (int sp = 0; sp < max_angular_steps; sp++) {
		// compute_angular_offsets :94-157
		float anglesum_x = 0.0f, anglesum_y = 0.0f;
		for (int j = 0; j < weight_count; j++) {
			float sample = clampzo(dwi[j]) * (64 - 1.0f);
			int isample = f2i_rtn(sample);
			anglesum_x += ASTC_LDG(&ct->cos_table[isample][sp]);
			anglesum_y += ASTC_LDG(&ct->sin_table[isample][sp]);
		}
		float angle = approx_atan2(anglesum_y, anglesum_x);
		angle = (angle == angle) ? angle : 0.0f;
		float offset = angle * mult;
		angular_offsets[sp] = offset;
		// compute_lowest_and_highest_weight :160-253
		float rcp_stepsize = static_cast<float>(sp) + 1.0f;
		float minidx = ASTC_RINT(min_weight * rcp_stepsize - offset);
		float maxidx = ASTC_RINT(max_weight * rcp_stepsize - offset);
		float errval = 0.0f, cut_low = 0.0f, cut_high = 0.0f;
		for (int j = 0; j < weight_count; j++) {
			float sval = dwi[j] * rcp_stepsize - offset;
			float svalrte = ASTC_RINT(sval);
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
		span = mini(span, max_quant_steps + 3);
		span = maxi(span, 2);
		lowest_weight[sp] = minidx;
		float ssize = 1.0f / rcp_stepsize;
		float errscale = ssize * ssize;
		float error = errval * errscale;
		float cut_low_weight_error = cut_low * errscale;
		float cut_high_weight_error = cut_high * errscale;
		// :298-330
		float i_flt = static_cast<float>(sp);
		float error_cut_low = error + cut_low_weight_error;
		float error_cut_high = error + cut_high_weight_error;
		float error_cut_low_high = error + cut_low_weight_error + cut_high_weight_error;
		if (best_err[span] > error) {
			best_err[span] = error;
			best_idx[span] = i_flt;
			best_cut[span] = 0.0f;
		}
		if (best_err[span - 1] > error_cut_low) {
			best_err[span - 1] = error_cut_low;
			best_idx[span - 1] = i_flt;
			best_cut[span - 1] = 1.0f;
		}
		if (best_err[span - 1] > error_cut_high) {
			best_err[span - 1] = error_cut_high;
			best_idx[span - 1] = i_flt;
			best_cut[span - 1] = 0.0f;
		}
		if (best_err[span - 2] > error_cut_low_high) {
			best_err[span - 2] = error_cut_low_high;
			best_idx[span - 2] = i_flt;
			best_cut[span - 2] = 1.0f;
		}
	}
	for (unsigned int i = 0; i <= max_quant_level; i++) {
		int q = steps_for_quant_level[i];
		int bsi = (int)best_idx[q];
		bsi = maxi(0, bsi);
		float lwi = lowest_weight[bsi] + best_cut[q];
		float hwi = lwi + static_cast<float>(q) - 1.0f;
		float stepsize = 1.0f / (1.0f + static_cast<float>(bsi));
		lowhigh[2 * i] = (angular_offsets[bsi] + lwi) * stepsize;
		lowhigh[2 * i + 1] = (angular_offsets[bsi] + hwi) * stepsize;
	}
}

// compute_angular_endpoints_1plane / _2planes (:358-500): lanes over (grid, plane)
ASTC_COOP void compute_angular_endpoints(WCtx& w, bool only_always, int nplanes, unsigned int max_weight_quant) {
	const DevBsd& bsd = *w.bsd;
	unsigned int max_dm = (nplanes == 1 && only_always) ? bsd.decimation_mode_count_always : bsd.decimation_mode_count_selected;
	uint16_t mask = (uint16_t)((1u << (max_weight_quant + 1)) - 1);
	int items = (int)max_dm * nplanes;
	for (int id = w.lane; id < items; id += ASTC_WARP) {
		int d = id / nplanes;
		int pl = id - d * nplanes;
		const DevDecMode& dm = bsd.dec_modes[d];
		uint16_t ref = nplanes == 1 ? dm.refprec_1plane : dm.refprec_2planes;
		if ((ref & mask) == 0) {
			continue;
		}
		unsigned int max_precision = (unsigned int)(nplanes == 1 ? dm.maxprec_1plane : dm.maxprec_2planes);
		if (max_precision > TUNE_MAX_ANGULAR_QUANT) max_precision = TUNE_MAX_ANGULAR_QUANT;
		if (max_precision > max_weight_quant) max_precision = max_weight_quant;
		compute_angular_endpoints_for_quant_levels(dm.weight_count, w.dwi + dm.dwi_offset + pl * dm.weight_count, max_precision,
		                                           w.lowhigh + (d * 2 + pl) * 16);
	}
	wsync();
}

// The (low, high) weight range of a packed block mode, including the "snap high to 1.0" rule of
// astcenc_compress_symbolic.cpp:459 / :819-827.
ASTC_FN void mode_low_high(const WCtx& w, const DevBlockMode& bm, int plane, float min_wt_cutoff, float& low, float& high) {
	if (bm.quant_mode <= TUNE_MAX_ANGULAR_QUANT) {
		const float* lh = w.lowhigh + (bm.decimation_mode * 2 + plane) * 16 + bm.quant_mode * 2;
		low = lh[0];
		high = lh[1];
	} else {
		low = 0.0f;
		high = 1.0f;
	}
	if (high > 1.02f * min_wt_cutoff) {
		high = 1.0f;
	}
}

// compute_quantized_weights_for_decimation (:974-1080), one weight
struct WeightQuantizer {
	float scale, scaled_low_bound, quant_level_m1, rscale, low_bound;
	int steps_m1;
	const uint8_t* q2u;
};

ASTC_FN WeightQuantizer make_weight_quantizer(float low_bound, float high_bound, int quant_level) {
	const float quant_levels_m1[12] = {1.0f, 2.0f, 3.0f, 4.0f, 5.0f, 7.0f, 9.0f, 11.0f, 15.0f, 19.0f, 23.0f, 31.0f};
	WeightQuantizer z;
	z.steps_m1 = (int)quant_level_count(quant_level) - 1;
	z.quant_level_m1 = quant_levels_m1[quant_level];
	if (high_bound <= low_bound) {
		low_bound = 0.0f;
		high_bound = 1.0f;
	}
	float rscale = high_bound - low_bound;
	z.scale = 1.0f / rscale;
	z.scaled_low_bound = low_bound * z.scale;
	z.rscale = rscale * (1.0f / 64.0f);
	z.low_bound = low_bound;
	z.q2u = ASTC_CT->wq_quant_to_unquant[quant_level];
	return z;
}

ASTC_FN int quantize_weight(const WeightQuantizer& z, float ideal) {
	float ix = ideal * z.scale - z.scaled_low_bound;
	ix = clampzo(ix);
	float ix1 = ix * z.quant_level_m1;
	int weightl = f2i(ix1);
	int weighth = mini(weightl + 1, z.steps_m1);
	int ixli = z.q2u[weightl];
	int ixhi = z.q2u[weighth];
	float ixl = static_cast<float>(ixli);
	float ixh = static_cast<float>(ixhi);
	bool mask = (ixl + ixh) < (128.0f * ix);
	return mask ? ixhi : ixli;
}
ASTC_FN float quantized_weight_value(const WeightQuantizer& z, int uq) {
	return static_cast<float>(uq) * z.rscale + z.low_bound;
}
