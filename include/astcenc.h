/*
 * astcenc.h - C ABI of the B200-native ASTC block compressor (libastcenc_b200.so).
 *
 * This is the drop-in boundary: the library exports exactly the entry points that the reference
 * codec's public header declares, with identical names, argument meaning, struct layouts and error
 * behaviour, so the reference command line tool, its unit tests and any application written against
 * the reference API can be relinked against this library unchanged.
 *
 * Reference interface replaced by each declaration (file:line in /root/reference/Source):
 *   astcenc_error / profile / swz / type enums ....... astcenc.h:207-236, 241-256, 284-313, 318-329
 *   astcenc_config ................................... astcenc.h:427-605
 *   astcenc_image .................................... astcenc.h:613-629
 *   astcenc_block_info ............................... astcenc.h:637-717
 *   astcenc_config_init .............................. astcenc.h:725-749  (impl astcenc_entry.cpp:504)
 *   astcenc_context_alloc ............................ astcenc.h:751-772  (impl astcenc_entry.cpp:726)
 *   astcenc_compress_image ........................... astcenc.h:774-799  (impl astcenc_entry.cpp:1113)
 *   astcenc_compress_reset / _cancel ................. astcenc.h:801-823  (impl astcenc_entry.cpp:1231, :1251)
 *   astcenc_decompress_image / _reset ................ astcenc.h:825-861  (impl astcenc_entry.cpp:1274, :1389)
 *   astcenc_context_free ............................. astcenc.h:863-868  (impl astcenc_entry.cpp:862)
 *   astcenc_get_block_info ........................... astcenc.h:870-884  (impl astcenc_entry.cpp:1401)
 *   astcenc_get_error_string ......................... astcenc.h:886-894  (impl astcenc_entry.cpp:1520)
 *
 * Threading contract (same as the reference): a context compresses one image at a time; up to
 * `thread_count` callers may enter astcenc_compress_image() with distinct thread_index values, any
 * subset may actually call, and all of them return once the image is complete. Here the first
 * arrival drives the GPU (upload, kernel, download) and the others simply wait.
 *
 * Extensions (prefix astcenc_b200_) expose device-resident buffers and slab sharding for multi-GPU use.
 */
#ifndef ASTCENC_B200_PUBLIC_H
#define ASTCENC_B200_PUBLIC_H

#include <stddef.h>
#include <stdint.h>

#if defined(__cplusplus)
	#define ASTCENC_B200_EXTERN extern "C"
#else
	#include <stdbool.h>
	#define ASTCENC_B200_EXTERN
#endif
#define ASTCENC_PUBLIC ASTCENC_B200_EXTERN __attribute__((visibility("default")))

struct astcenc_context;

enum astcenc_error {
	ASTCENC_SUCCESS = 0,
	ASTCENC_ERR_OUT_OF_MEM,
	ASTCENC_ERR_BAD_CPU_FLOAT,
	ASTCENC_ERR_BAD_PARAM,
	ASTCENC_ERR_BAD_BLOCK_SIZE,
	ASTCENC_ERR_BAD_PROFILE,
	ASTCENC_ERR_BAD_QUALITY,
	ASTCENC_ERR_BAD_SWIZZLE,
	ASTCENC_ERR_BAD_FLAGS,
	ASTCENC_ERR_BAD_CONTEXT,
	ASTCENC_ERR_NOT_IMPLEMENTED,
	ASTCENC_ERR_BAD_DECODE_MODE
};

enum astcenc_profile { ASTCENC_PRF_LDR_SRGB = 0, ASTCENC_PRF_LDR, ASTCENC_PRF_HDR_RGB_LDR_A, ASTCENC_PRF_HDR };

/* search quality presets (any value in [0, 100] is accepted and interpolated) */
static const float ASTCENC_PRE_FASTEST = 0.0f;
static const float ASTCENC_PRE_FAST = 10.0f;
static const float ASTCENC_PRE_MEDIUM = 60.0f;
static const float ASTCENC_PRE_THOROUGH = 98.0f;
static const float ASTCENC_PRE_VERYTHOROUGH = 99.0f;
static const float ASTCENC_PRE_EXHAUSTIVE = 100.0f;

enum astcenc_swz { ASTCENC_SWZ_R = 0, ASTCENC_SWZ_G = 1, ASTCENC_SWZ_B = 2, ASTCENC_SWZ_A = 3, ASTCENC_SWZ_0 = 4, ASTCENC_SWZ_1 = 5, ASTCENC_SWZ_Z = 6 };

struct astcenc_swizzle {
	enum astcenc_swz r;
	enum astcenc_swz g;
	enum astcenc_swz b;
	enum astcenc_swz a;
};

enum astcenc_type { ASTCENC_TYPE_U8 = 0, ASTCENC_TYPE_F16 = 1, ASTCENC_TYPE_F32 = 2 };

ASTCENC_B200_EXTERN typedef void (*astcenc_progress_callback)(float);

static const unsigned int ASTCENC_FLG_MAP_NORMAL = 1 << 0;
static const unsigned int ASTCENC_FLG_USE_DECODE_UNORM8 = 1 << 1;
static const unsigned int ASTCENC_FLG_USE_ALPHA_WEIGHT = 1 << 2;
static const unsigned int ASTCENC_FLG_USE_PERCEPTUAL = 1 << 3;
static const unsigned int ASTCENC_FLG_DECOMPRESS_ONLY = 1 << 4;
static const unsigned int ASTCENC_FLG_SELF_DECOMPRESS_ONLY = 1 << 5;
static const unsigned int ASTCENC_FLG_MAP_RGBM = 1 << 6;
static const unsigned int ASTCENC_ALL_FLAGS = (1 << 0) | (1 << 1) | (1 << 2) | (1 << 3) | (1 << 4) | (1 << 5) | (1 << 6);

struct astcenc_config {
	enum astcenc_profile profile;
	unsigned int flags;
	unsigned int block_x;
	unsigned int block_y;
	unsigned int block_z;
	float cw_r_weight;
	float cw_g_weight;
	float cw_b_weight;
	float cw_a_weight;
	unsigned int a_scale_radius;
	float rgbm_m_scale;
	unsigned int tune_partition_count_limit;
	unsigned int tune_2partition_index_limit;
	unsigned int tune_3partition_index_limit;
	unsigned int tune_4partition_index_limit;
	unsigned int tune_block_mode_limit;
	unsigned int tune_refinement_limit;
	unsigned int tune_candidate_limit;
	unsigned int tune_2partitioning_candidate_limit;
	unsigned int tune_3partitioning_candidate_limit;
	unsigned int tune_4partitioning_candidate_limit;
	float tune_db_limit;
	float tune_mse_overshoot;
	float tune_2partition_early_out_limit_factor;
	float tune_3partition_early_out_limit_factor;
	float tune_2plane_early_out_limit_correlation;
	float tune_search_mode0_enable;
	astcenc_progress_callback progress_callback;
};

/* data[z] points at slice z: dim_x * dim_y RGBA texels of data_type, row-major, tightly packed */
struct astcenc_image {
	unsigned int dim_x;
	unsigned int dim_y;
	unsigned int dim_z;
	enum astcenc_type data_type;
	void** data;
};

struct astcenc_block_info {
	enum astcenc_profile profile;
	unsigned int block_x;
	unsigned int block_y;
	unsigned int block_z;
	unsigned int texel_count;
	bool is_error_block;
	bool is_constant_block;
	bool is_hdr_block;
	bool is_dual_plane_block;
	unsigned int partition_count;
	unsigned int partition_index;
	unsigned int dual_plane_component;
	unsigned int color_endpoint_modes[4];
	unsigned int color_level_count;
	unsigned int weight_level_count;
	unsigned int weight_x;
	unsigned int weight_y;
	unsigned int weight_z;
	float color_endpoints[4][2][4];
	float weight_values_plane1[216];
	float weight_values_plane2[216];
	uint8_t partition_assignment[216];
};

ASTCENC_PUBLIC enum astcenc_error astcenc_config_init(enum astcenc_profile profile, unsigned int block_x, unsigned int block_y, unsigned int block_z,
                                                      float quality, unsigned int flags, struct astcenc_config* config);

ASTCENC_PUBLIC enum astcenc_error astcenc_context_alloc(const struct astcenc_config* config, unsigned int thread_count,
                                                        struct astcenc_context** context, const struct astcenc_context* parent_context);

ASTCENC_PUBLIC enum astcenc_error astcenc_compress_image(struct astcenc_context* context, struct astcenc_image* image, const struct astcenc_swizzle* swizzle,
                                                         uint8_t* data_out, size_t data_len, unsigned int thread_index);

ASTCENC_PUBLIC enum astcenc_error astcenc_compress_reset(struct astcenc_context* context);

ASTCENC_PUBLIC enum astcenc_error astcenc_compress_cancel(struct astcenc_context* context);

ASTCENC_PUBLIC enum astcenc_error astcenc_decompress_image(struct astcenc_context* context, const uint8_t* data, size_t data_len,
                                                           struct astcenc_image* image_out, const struct astcenc_swizzle* swizzle, unsigned int thread_index);

ASTCENC_PUBLIC enum astcenc_error astcenc_decompress_reset(struct astcenc_context* context);

ASTCENC_PUBLIC void astcenc_context_free(struct astcenc_context* context);

ASTCENC_PUBLIC enum astcenc_error astcenc_get_block_info(struct astcenc_context* context, const uint8_t data[16], struct astcenc_block_info* info);

ASTCENC_PUBLIC const char* astcenc_get_error_string(enum astcenc_error status);

/* ---- B200 extensions ------------------------------------------------------------------------ */

/* Behaviour notes for the ten reference entry points above, where the GPU implementation differs in timing (never in
 * results) from the reference's thread pool:
 *  - progress_callback is called once per 2D slice (at 100 % for a 2D image); the reference calls it per 16-block ticket
 *    (astcenc_entry.cpp:937). astcenc_compress_cancel() takes effect between slices: a pipeline pass that is already
 *    enqueued runs to its end (the reference stops handing out tickets, astcenc_internal_entry.h:219).
 *  - Every entry point runs on the CUDA device that was current in astcenc_context_alloc() and restores the caller's
 *    current device before it returns.
 *  - Block sizes: the fourteen 2D and the ten 3D footprints (block_z > 1: 3x3x3 .. 6x6x6, astcenc_block_sizes.cpp:1025). With a
 *    3D block size the slices of astcenc_image::data are uploaded into one volume and compressed in one pass, blocks in
 *    (z, y, x) order like the reference's (astcenc_entry.cpp:1036); a_scale_radius is ignored like the reference does
 *    (astcenc_entry.cpp:975). The astcenc_b200_* device-resident and multi-GPU entry points take 2D block sizes only
 *    (ASTCENC_ERR_NOT_IMPLEMENTED otherwise).
 *  - A context owns ONE set of search scratch buffers: passes of the same context are serialised - on the host by a mutex,
 *    on the device by an event chain - whatever streams they are enqueued on. Use one context per concurrent stream.
 */

/* Compress block rows [block_row0, block_row0 + block_rows) of an image that is ALREADY RESIDENT in device
 * memory (d_pixels: the whole dim_x * dim_y image), writing block_rows * blocks_x * 16 bytes to the device
 * buffer d_out. Enqueued on `cuda_stream` (a cudaStream_t passed as void*, 0 = the context's own stream);
 * returns without synchronising. Used for slab sharding across GPUs and by bench.py's device-resident timing.
 * Calls on different streams of the same context are ordered one after the other (see above). */
ASTCENC_PUBLIC enum astcenc_error astcenc_b200_compress_device(struct astcenc_context* context, const void* d_pixels, enum astcenc_type data_type,
                                                               unsigned int dim_x, unsigned int dim_y, const struct astcenc_swizzle* swizzle,
                                                               unsigned int block_row0, unsigned int block_rows, uint8_t* d_out, void* cuda_stream);

/**
 * Multi-GPU sharding: one process per GPU, each with its own context; the only exchange is the gather of the compressed
 * payload over NCCL (loaded at run time: "libnccl.so.2"; ASTCENC_ERR_NOT_IMPLEMENTED when it cannot be loaded).
 * The reference has no counterpart - its unit of distribution is the block ticket of one process
 * (astcenc_internal_entry.h:97-324); the payload layout the ranks agree on is astcenc_entry.cpp:1036.
 *
 * astcenc_b200_comm_unique_id   rank 0 makes the 128-byte rendezvous id; the caller ships it to the other ranks
 * astcenc_b200_comm_init        collective: joins `world` ranks (world == 1: no communicator, the calls below run locally)
 * astcenc_b200_slab_rows        the block rows [first, first + rows) rank `rank` of `world` owns for an image of height dim_y
 * astcenc_b200_compress_image_sharded
 *     collective, SLAB MODE: every rank passes the same 2D image (host pointer, only the rows of its own slab are read and
 *     uploaded), compresses its slab, and the slabs are gathered into the root's device buffer (grouped ncclSend/ncclRecv)
 *     and copied to data_out on the root (data_out is ignored elsewhere). Bytes are identical to astcenc_compress_image().
 * astcenc_b200_compress_batch
 *     collective, BATCH MODE: image i of `images` belongs to rank i % world (other entries are not touched on this rank and
 *     may be NULL); all images share size and type. Each rank searches its images one after the other with the upload of
 *     the next image overlapped (two device image buffers); payload i arrives in data_out[i] on the root.
 * astcenc_b200_comm_last_timing device milliseconds of the last sharded call on this rank: search, and the gather alone
 */
ASTCENC_PUBLIC enum astcenc_error astcenc_b200_comm_unique_id(void* id_out, size_t id_bytes);
ASTCENC_PUBLIC enum astcenc_error astcenc_b200_comm_init(struct astcenc_context* context, int rank, int world, const void* id, size_t id_bytes);
ASTCENC_PUBLIC enum astcenc_error astcenc_b200_comm_free(struct astcenc_context* context);
ASTCENC_PUBLIC enum astcenc_error astcenc_b200_slab_rows(struct astcenc_context* context, unsigned int dim_y, int rank, int world,
                                                         unsigned int* first_block_row, unsigned int* block_rows);
ASTCENC_PUBLIC enum astcenc_error astcenc_b200_compress_image_sharded(struct astcenc_context* context, struct astcenc_image* image,
                                                                      const struct astcenc_swizzle* swizzle, uint8_t* data_out, size_t data_len, int root);
ASTCENC_PUBLIC enum astcenc_error astcenc_b200_compress_batch(struct astcenc_context* context, struct astcenc_image* const* images, unsigned int image_count,
                                                              const struct astcenc_swizzle* swizzle, uint8_t* const* data_out, size_t data_len_each, int root);
ASTCENC_PUBLIC enum astcenc_error astcenc_b200_comm_last_timing(struct astcenc_context* context, float* compress_ms, float* gather_ms);

/* Number of kernel launches issued by this context so far (bench.py reports it as gpu_launches). */
ASTCENC_PUBLIC unsigned long long astcenc_b200_launch_count(struct astcenc_context* context);

/**
 * Per-stage timing of the wave pipeline (measurement aid). enable != 0 makes the following
 * astcenc_b200_compress_device() calls record a CUDA event after every kernel launch; a later call with non-NULL
 * arrays returns, for the last such call, the summed launch durations and launch counts per kernel
 * (0 setup, 1 refine, 2 prepare, 3 emit). Leave it off for production use.
 */
ASTCENC_PUBLIC enum astcenc_error astcenc_b200_stage_timing(struct astcenc_context* context, int enable, float stage_ms[4], unsigned int stage_launches[4]);

/* Milliseconds of device time (CUDA events on the launching stream) spent in the compress kernel by the most
 * recent astcenc_compress_image() call on this context, and its H2D / D2H byte counts. */
ASTCENC_PUBLIC enum astcenc_error astcenc_b200_last_timing(struct astcenc_context* context, float* kernel_ms, size_t* h2d_bytes, size_t* d2h_bytes);

/**
 * Image error metrics on the device: what the reference's command line tool computes on the host after a -t* round trip
 * (Source/astcenccli_error_metrics.cpp:109-413, compute_error_metrics(); it prints, this returns). img1 is the
 * original, img2 the decoded image (host pointers, any mix of U8 / F16 / F32; only the intersection is compared).
 * input_components (1-4) selects the channels that count, as in the reference (1: RGB, 2: BA, 3: RGB, 4: RGBA).
 * Fields not requested (HDR / normal-map metrics) are zero. PSNR values are 999.0 for identical images.
 */
struct astcenc_b200_error_metrics {
	double psnr;                  /* "PSNR (LDR-RGBA)" when alpha counts, else "PSNR (LDR-RGB)" */
	double alpha_psnr;            /* "Alpha-weighted PSNR" (= psnr when alpha does not count) */
	double rgb_psnr;              /* "PSNR (LDR-RGB)" */
	double rgb_peak;              /* largest R/G/B value of img1 */
	double peak_psnr;             /* HDR: "PSNR (RGB norm to peak)" */
	double mpsnr;                 /* HDR: "mPSNR (RGB)" over fstop_lo..fstop_hi */
	double log_rmse;              /* HDR: "LogRMSE (RGB)" */
	double mean_angular_error;    /* normal maps: degrees */
	double worst_angular_error;   /* normal maps: degrees */
	double sum_squared_error[4];  /* per channel, values scaled to 0..1 (LDR) */
};
ASTCENC_PUBLIC enum astcenc_error astcenc_b200_compute_error_metrics(struct astcenc_context* context, int compute_hdr_metrics, int compute_normal_metrics,
                                                                     int input_components, const struct astcenc_image* img1, const struct astcenc_image* img2,
                                                                     int fstop_lo, int fstop_hi, struct astcenc_b200_error_metrics* metrics);

/**
 * The .astc container (Docs/FileFormat.md; Source/astcenccli_image_load_store.cpp:2573-2760, load_cimage / store_cimage):
 * a 16-byte header - magic 13 AB A1 5C, block dimensions (one byte each), image dimensions (24 bits each, little endian)
 * - followed by the blocks as astcenc_compress_image() wrote them.
 * store: data_len must be the size of the block grid the header describes.
 * load: call with data == NULL to get the header and the payload size in *data_len, then again with a buffer;
 * corrupt files (bad magic, zero dimensions, size overflow, fewer payload bytes than the header promises) give
 * ASTCENC_ERR_BAD_PARAM, a buffer smaller than the payload ASTCENC_ERR_OUT_OF_MEM.
 */
struct astcenc_b200_cimage_header {
	unsigned int block_x, block_y, block_z;
	unsigned int dim_x, dim_y, dim_z;
};
ASTCENC_PUBLIC enum astcenc_error astcenc_b200_store_cimage(const char* filename, const struct astcenc_b200_cimage_header* header,
                                                            const uint8_t* data, size_t data_len);
ASTCENC_PUBLIC enum astcenc_error astcenc_b200_load_cimage(const char* filename, struct astcenc_b200_cimage_header* header,
                                                           uint8_t* data, size_t data_capacity, size_t* data_len);

/**
 * The same payload in a KTX 1 container (Source/astcenccli_image_load_store.cpp:870-905 header, :1294-1440
 * load_ktx_compressed_image / store_ktx_compressed_image): 64-byte header with glInternalFormat =
 * GL_COMPRESSED_RGBA_ASTC_<footprint> (or the SRGB8_ALPHA8 variant when is_srgb), one mip level, no key/value data, then
 * the 32-bit payload size and the blocks. load accepts files of either byte order, skips key/value data and reads the first
 * mip level; the two-call protocol (data == NULL first) and the error codes are those of astcenc_b200_load_cimage, a footprint
 * without a GL enum gives ASTCENC_ERR_BAD_BLOCK_SIZE on store.
 */
ASTCENC_PUBLIC enum astcenc_error astcenc_b200_store_ktx_cimage(const char* filename, const struct astcenc_b200_cimage_header* header, int is_srgb,
                                                                const uint8_t* data, size_t data_len);
ASTCENC_PUBLIC enum astcenc_error astcenc_b200_load_ktx_cimage(const char* filename, struct astcenc_b200_cimage_header* header, int* is_srgb,
                                                               uint8_t* data, size_t data_capacity, size_t* data_len);

#endif
