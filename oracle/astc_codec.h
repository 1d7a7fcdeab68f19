// ORACLE (test infrastructure, not product code): CPU restatement of the reference's per-block
// compressor compress_block() (astcenc_compress_symbolic.cpp:1162) and everything beneath it.
// Plain scalar C++; each function cites the reference file:line it follows.
#ifndef ASTC_ORACLE_CODEC_H
#define ASTC_ORACLE_CODEC_H

#include "astc_math.h"
#include "astc_tables.h"
#include "astc_color.h"

namespace ao {

static const unsigned int FLG_MAP_NORMAL = 1 << 0;
static const unsigned int FLG_USE_DECODE_UNORM8 = 1 << 1;
static const unsigned int FLG_USE_ALPHA_WEIGHT = 1 << 2;
static const unsigned int FLG_USE_PERCEPTUAL = 1 << 3;
static const unsigned int FLG_DECOMPRESS_ONLY = 1 << 4;
static const unsigned int FLG_SELF_DECOMPRESS_ONLY = 1 << 5;
static const unsigned int FLG_MAP_RGBM = 1 << 6;

// Mirror of astcenc_config (astcenc.h:427-605) - only what the encoder consumes.
struct Config {
	int profile;
	unsigned int flags;
	unsigned int block_x, block_y, block_z;
	float cw_r_weight, cw_g_weight, cw_b_weight, cw_a_weight;
	unsigned int a_scale_radius;
	float rgbm_m_scale;
	unsigned int tune_partition_count_limit;
	unsigned int tune_2partition_index_limit, tune_3partition_index_limit, tune_4partition_index_limit;
	unsigned int tune_block_mode_limit;
	unsigned int tune_refinement_limit;
	unsigned int tune_candidate_limit;
	unsigned int tune_2partitioning_candidate_limit, tune_3partitioning_candidate_limit, tune_4partitioning_candidate_limit;
	float tune_db_limit;
	float tune_mse_overshoot;
	float tune_2partition_early_out_limit_factor, tune_3partition_early_out_limit_factor;
	float tune_2plane_early_out_limit_correlation;
	float tune_search_mode0_enable;
};

// astcenc_config_init (astcenc_entry.cpp:504-723). Returns 0 on success, else the astcenc_error value.
int config_init(int profile, unsigned int block_x, unsigned int block_y, float quality, unsigned int flags, Config& cfg, unsigned int block_z = 1);
// validate_config clamps + dB->error conversion done by astcenc_context_alloc (astcenc_entry.cpp:434-501, 814-821)
int config_finalize(Config& cfg);

struct ImageBlock {   // image_block (astcenc_internal.h:749-884)
	float data_r[MAX_TEXELS], data_g[MAX_TEXELS], data_b[MAX_TEXELS], data_a[MAX_TEXELS];
	uint8_t texel_count;
	f4 origin_texel, data_min, data_mean, data_max, channel_weight;
	bool grayscale;
	bool decode_unorm8;
	uint8_t rgb_lns0, alpha_lns0;
};

struct SymbolicBlock {   // symbolic_compressed_block (astcenc_internal.h:1077-1134)
	uint8_t block_type;
	uint8_t partition_count;
	uint8_t color_formats_matched;
	int8_t plane2_component;
	uint16_t block_mode;
	uint16_t partition_index;
	uint8_t color_formats[4];
	uint8_t quant_mode;
	float errorval;
	int constant_color[4];
	uint8_t color_values[4][8];
	uint8_t weights[MAX_WEIGHTS];
};

enum { SYM_BTYPE_ERROR = 0, SYM_BTYPE_CONST_F16 = 1, SYM_BTYPE_CONST_U16 = 2, SYM_BTYPE_NONCONST = 3 };

struct Context {
	Config config;
	BlockSizeTables* bsd;
	void* work;   // per-context scratch (single threaded)
};

Context* context_create(const Config& cfg);   // cfg already finalized
void context_destroy(Context* ctx);

// data_type: 0 = U8, 1 = F16, 2 = F32 (astcenc_type). swz: 4 entries of astcenc_swz.
// load_image_block / load_image_block_fast_ldr (astcenc_image.cpp:162, :278)
void load_block(const Context& ctx, const void* data, int data_type, unsigned int dim_x, unsigned int dim_y,
                unsigned int pos_x, unsigned int pos_y, const int swz[4], ImageBlock& blk, unsigned int dim_z = 1, unsigned int pos_z = 0);

// compress_block (astcenc_compress_symbolic.cpp:1162)
void compress_block(const Context& ctx, const ImageBlock& blk, uint8_t pcb[16]);

// whole-image loop of compress_image (astcenc_entry.cpp:891-1043); a volume is dim_z slices of dim_x * dim_y texels, contiguous
// astcenc_decompress_image (astcenc_entry.cpp:1274-1385); swz uses astcenc_swz numbering (6 = Z)
void decompress_image(const Context& ctx, const uint8_t* data, void* out, int data_type, unsigned int dim_x, unsigned int dim_y, const int swz[4],
                      unsigned int dim_z = 1);
void compress_image(const Context& ctx, const void* data, int data_type, unsigned int dim_x, unsigned int dim_y,
                    const int swz[4], uint8_t* out, unsigned int dim_z = 1);

void symbolic_to_physical(const BlockSizeTables& bsd, const SymbolicBlock& scb, uint8_t pcb[16]);

}  // namespace ao
#endif
