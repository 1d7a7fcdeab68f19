// Host-side construction of the static / derived ASTC tables (once per context), later packed into the
// compact device layout of astc_dev_tables.h by pack_device_tables().
//
// Follows (reference file:line, relative to /root/reference/Source):
//   block modes / decimation tables   astcenc_block_sizes.cpp:36-98, 252-486, 717-754, 822-1002
//   partition tables                  astcenc_partition_tables.cpp:114-165, 276-470
//   BISE tables + bit counts          astcenc_integer_sequence.cpp:28-330, 419-435
//   colour / weight quant tables      astcenc_quantization.cpp:27-850, astcenc_weight_quant_xfer_tables.cpp:26
//   percentile tables (data only)     astcenc_percentile_tables.cpp:1165
//   sin/cos tables                    astcenc_weight_align.cpp:72-84
#ifndef ASTC_HOST_TABLES_H
#define ASTC_HOST_TABLES_H

#include <cstdint>
#include <cstddef>

namespace astc_host {

static const int MAX_TEXELS = 216;        // largest block (6x6x6)
static const int MAX_WEIGHTS = 64;
static const int PLANE2_OFFSET = 32;
static const int MAX_PARTITIONINGS = 1024;
static const int MAX_BLOCK_MODES = 2048;
static const int MAX_DECIMATION_MODES = 87;
static const int MAX_KMEANS_TEXELS = 64;
static const int MAX_WT_ENTRIES = MAX_TEXELS * 4;

enum { QUANT_2 = 0, QUANT_3, QUANT_4, QUANT_5, QUANT_6, QUANT_8, QUANT_10, QUANT_12, QUANT_16, QUANT_20, QUANT_24,
       QUANT_32, QUANT_40, QUANT_48, QUANT_64, QUANT_80, QUANT_96, QUANT_128, QUANT_160, QUANT_192, QUANT_256 };

struct BlockMode {
	uint16_t mode_index;      // the physical 11-bit mode
	uint8_t decimation_mode;  // packed decimation index
	uint8_t quant_mode;       // weight quant level
	uint8_t weight_bits;
	uint8_t is_dual_plane;
};

struct DecimationMode {
	int8_t maxprec_1plane;
	int8_t maxprec_2planes;
	uint16_t refprec_1plane;
	uint16_t refprec_2planes;
};

// One weight grid: bilinear tables in both directions. The weight->texel direction is a CSR list.
struct DecimationInfo {
	uint8_t texel_count;
	uint8_t max_texel_weight_count;
	uint8_t weight_count;
	uint8_t weight_x;
	uint8_t weight_y;
	uint8_t weight_z;
	uint8_t texel_weight_count[MAX_TEXELS];
	uint8_t texel_weights[4][MAX_TEXELS];
	uint8_t texel_weight_contribs_int[4][MAX_TEXELS];
	float texel_weight_contribs_float[4][MAX_TEXELS];
	uint8_t weight_texel_count[MAX_WEIGHTS];
	uint16_t weight_texel_offset[MAX_WEIGHTS + 1];
	uint8_t weight_texels[MAX_WT_ENTRIES];
	float weight_texel_contribs[MAX_WT_ENTRIES];     // integer contribution 1..16 as float
	float texel_contrib_for_weight[MAX_WT_ENTRIES];  // same contribution / 16
};

struct PartitionInfo {
	uint16_t partition_count;
	uint16_t partition_index;                      // the physical 10-bit seed
	uint8_t partition_texel_count[4];
	uint8_t partition_of_texel[MAX_TEXELS];
	uint8_t texels_of_partition[4][MAX_TEXELS];
};

struct WeightQuantTable {
	uint8_t quant_to_unquant[32];
	uint8_t scramble_map[32];
	uint8_t unscramble_and_unquant_map[32];
	uint16_t prev_next_values[65];
};

// Process-wide constant tables (generated at start-up from the format rules).
struct ConstTables {
	uint8_t integer_of_trits[3][3][3][3][3];
	uint8_t trits_of_integer[256][5];
	uint8_t integer_of_quints[5][5][5];
	uint8_t quints_of_integer[128][3];
	uint8_t color_unquant_to_uquant[17][512];
	uint8_t color_uquant_to_scrambled_pquant[17][256];
	uint8_t color_scrambled_pquant_to_uquant[17][256];
	int8_t quant_mode_table[10][128];
	WeightQuantTable weight_quant[12];
	float sin_table[64][32];
	float cos_table[64][32];
};

const ConstTables& const_tables();

unsigned int get_quant_level(int quant);
unsigned int ise_sequence_bitcount(unsigned int character_count, int quant_level);
void ise_btq(int quant_level, unsigned int& bits, unsigned int& trits, unsigned int& quints);

// Everything derived from (block size, mode cutoff, partition cutoff).
struct BlockSizeTables {
	uint8_t dim_x, dim_y, dim_z, texel_count;
	unsigned int decimation_mode_count_always, decimation_mode_count_selected, decimation_mode_count_all;
	unsigned int block_mode_count_1plane_always, block_mode_count_1plane_selected,
	             block_mode_count_1plane_2plane_selected, block_mode_count_all;
	unsigned int partitioning_count_selected[4], partitioning_count_all[4];
	DecimationMode decimation_modes[MAX_DECIMATION_MODES];
	DecimationInfo* decimation_tables;               // [decimation_mode_count_all]
	uint16_t block_mode_packed_index[MAX_BLOCK_MODES];
	BlockMode block_modes[MAX_BLOCK_MODES];
	PartitionInfo* partitionings[5];                 // [1]=1 partition (1 entry), [2..4]
	uint16_t partitioning_packed_index[3][MAX_PARTITIONINGS];
	uint8_t kmeans_texels[MAX_KMEANS_TEXELS];
	uint64_t* coverage_bitmaps[5];                   // [pc][packed * pc + p], pc = 2..4
};

bool is_legal_2d_block_size(unsigned int x, unsigned int y);
bool is_legal_3d_block_size(unsigned int x, unsigned int y, unsigned int z);

// z == 1: a 2D block size (percentile-selected modes); z > 1: a 3D block size (every legal mode, astcenc_block_sizes.cpp:1025-1190)
BlockSizeTables* build_block_size_tables(unsigned int x, unsigned int y, unsigned int z, bool can_omit_modes,
                                         unsigned int partition_count_cutoff, float mode_cutoff);
void free_block_size_tables(BlockSizeTables* t);

}  // namespace astc_host
#endif
