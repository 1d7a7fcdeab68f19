// B200-native ASTC codec: the .astc container.
//
// The reference's command line tool stores compressed images as a 16-byte header followed by the raw blocks
// (Docs/FileFormat.md; astcenccli_image_load_store.cpp:2573-2760, load_cimage / store_cimage). These entry points
// read and write that container for callers of the C ABI; the payload is exactly what astcenc_compress_image()
// produced / astcenc_decompress_image() consumes. Header fields are single bytes, so there is no endianness issue.
#include "../../include/astcenc.h"
#include <cstdio>
#include <cstring>

namespace {

const unsigned char k_magic[4] = {0x13, 0xAB, 0xA1, 0x5C};      // 0x5CA1AB13, least significant byte first

struct RawHeader {
	unsigned char magic[4];
	unsigned char block_x, block_y, block_z;
	unsigned char dim_x[3], dim_y[3], dim_z[3];
};
static_assert(sizeof(RawHeader) == 16, "the .astc header is 16 bytes");

inline unsigned int get24(const unsigned char b[3]) {
	return (unsigned int)b[0] + ((unsigned int)b[1] << 8) + ((unsigned int)b[2] << 16);
}
inline void put24(unsigned char b[3], unsigned int v) {
	b[0] = (unsigned char)(v & 0xFF);
	b[1] = (unsigned char)((v >> 8) & 0xFF);
	b[2] = (unsigned char)((v >> 16) & 0xFF);
}

// multiply with overflow detection (the header can describe 2^72 blocks)
inline bool mul_ok(size_t a, size_t b, size_t& r) {
	return !__builtin_mul_overflow(a, b, &r);
}

// parse + validate the header the way load_cimage does (:2610-2662): magic, non-zero dimensions, payload size
astcenc_error parse_header(const RawHeader& raw, astcenc_b200_cimage_header* hdr, size_t* payload) {
	if (memcmp(raw.magic, k_magic, 4) != 0) {
		return ASTCENC_ERR_BAD_PARAM;
	}
	// (zero block dimensions are read as 1, :2626-2629)
	size_t bx = raw.block_x ? raw.block_x : 1, by = raw.block_y ? raw.block_y : 1, bz = raw.block_z ? raw.block_z : 1;
	size_t dx = get24(raw.dim_x), dy = get24(raw.dim_y), dz = get24(raw.dim_z);
	if (dx == 0 || dy == 0 || dz == 0) {
		return ASTCENC_ERR_BAD_PARAM;
	}
	size_t n = 0;
	if (!mul_ok((dx + bx - 1) / bx, (dy + by - 1) / by, n) || !mul_ok(n, (dz + bz - 1) / bz, n) || !mul_ok(n, 16, n)) {
		return ASTCENC_ERR_BAD_PARAM;
	}
	hdr->block_x = (unsigned int)bx;
	hdr->block_y = (unsigned int)by;
	hdr->block_z = (unsigned int)bz;
	hdr->dim_x = (unsigned int)dx;
	hdr->dim_y = (unsigned int)dy;
	hdr->dim_z = (unsigned int)dz;
	*payload = n;
	return ASTCENC_SUCCESS;
}

// ---- KTX 1 (astcenccli_image_load_store.cpp:870-905 header, :1294-1440 load / store of compressed images) ----
const unsigned char k_ktx_magic[12] = {0xAB, 0x4B, 0x54, 0x58, 0x20, 0x31, 0x31, 0xBB, 0x0D, 0x0A, 0x1A, 0x0A};
const uint32_t k_gl_rgba = 0x1908;

struct KtxHeader {
	unsigned char magic[12];
	uint32_t endianness, gl_type, gl_type_size, gl_format, gl_internal_format, gl_base_internal_format;
	uint32_t pixel_width, pixel_height, pixel_depth, number_of_array_elements, number_of_faces, number_of_mipmap_levels, bytes_of_key_value_data;
};
static_assert(sizeof(KtxHeader) == 64, "the KTX 1 header is 64 bytes");

// GL_COMPRESSED_RGBA_ASTC_<x>x<y>_KHR = 0x93B0 + i, ..._<x>x<y>x<z>_OES = 0x93C0 + j, the sRGB8_ALPHA8 variants + 0x20
// (the table at :760-835)
const unsigned char k_fp2d[14][2] = {{4, 4}, {5, 4}, {5, 5}, {6, 5}, {6, 6}, {8, 5}, {8, 6}, {8, 8}, {10, 5}, {10, 6}, {10, 8}, {10, 10}, {12, 10}, {12, 12}};
const unsigned char k_fp3d[10][3] = {{3, 3, 3}, {4, 3, 3}, {4, 4, 3}, {4, 4, 4}, {5, 4, 4}, {5, 5, 4}, {5, 5, 5}, {6, 5, 5}, {6, 6, 5}, {6, 6, 6}};

uint32_t gl_format_of(unsigned int x, unsigned int y, unsigned int z, bool srgb) {
	for (unsigned int i = 0; i < 14; i++) {
		if (z == 1 && k_fp2d[i][0] == x && k_fp2d[i][1] == y) {
			return 0x93B0u + i + (srgb ? 0x20u : 0u);
		}
	}
	for (unsigned int i = 0; i < 10; i++) {
		if (k_fp3d[i][0] == x && k_fp3d[i][1] == y && k_fp3d[i][2] == z) {
			return 0x93C0u + i + (srgb ? 0x20u : 0u);
		}
	}
	return 0;
}

bool footprint_of(uint32_t fmt, astcenc_b200_cimage_header* hdr, int* srgb) {
	*srgb = 0;
	if (fmt >= 0x93D0u && fmt <= 0x93E9u) {
		*srgb = 1;
		fmt -= 0x20u;
	}
	if (fmt >= 0x93B0u && fmt <= 0x93BDu) {
		hdr->block_x = k_fp2d[fmt - 0x93B0u][0];
		hdr->block_y = k_fp2d[fmt - 0x93B0u][1];
		hdr->block_z = 1;
		return true;
	}
	if (fmt >= 0x93C0u && fmt <= 0x93C9u) {
		hdr->block_x = k_fp3d[fmt - 0x93C0u][0];
		hdr->block_y = k_fp3d[fmt - 0x93C0u][1];
		hdr->block_z = k_fp3d[fmt - 0x93C0u][2];
		return true;
	}
	return false;
}

inline uint32_t bswap32(uint32_t v) {
	return (v >> 24) | ((v >> 8) & 0xFF00u) | ((v << 8) & 0xFF0000u) | (v << 24);
}

}      // namespace

extern "C" {

astcenc_error astcenc_b200_store_cimage(const char* filename, const astcenc_b200_cimage_header* hdr, const uint8_t* data, size_t data_len) {
	if (!filename || !hdr || (!data && data_len)) {
		return ASTCENC_ERR_BAD_PARAM;
	}
	if (hdr->block_x == 0 || hdr->block_x > 255 || hdr->block_y == 0 || hdr->block_y > 255 || hdr->block_z == 0 || hdr->block_z > 255 ||
	    hdr->dim_x == 0 || hdr->dim_x >= (1u << 24) || hdr->dim_y == 0 || hdr->dim_y >= (1u << 24) || hdr->dim_z == 0 || hdr->dim_z >= (1u << 24)) {
		return ASTCENC_ERR_BAD_PARAM;
	}
	RawHeader raw;
	memcpy(raw.magic, k_magic, 4);
	raw.block_x = (unsigned char)hdr->block_x;
	raw.block_y = (unsigned char)hdr->block_y;
	raw.block_z = (unsigned char)hdr->block_z;
	put24(raw.dim_x, hdr->dim_x);
	put24(raw.dim_y, hdr->dim_y);
	put24(raw.dim_z, hdr->dim_z);
	// the payload must be the block grid the header describes
	astcenc_b200_cimage_header check;
	size_t payload = 0;
	if (parse_header(raw, &check, &payload) != ASTCENC_SUCCESS || payload != data_len) {
		return ASTCENC_ERR_BAD_PARAM;
	}
	FILE* f = fopen(filename, "wb");
	if (!f) {
		return ASTCENC_ERR_BAD_PARAM;
	}
	bool ok = fwrite(&raw, sizeof(raw), 1, f) == 1 && (data_len == 0 || fwrite(data, 1, data_len, f) == data_len);
	ok = (fclose(f) == 0) && ok;
	return ok ? ASTCENC_SUCCESS : ASTCENC_ERR_BAD_PARAM;
}

astcenc_error astcenc_b200_load_cimage(const char* filename, astcenc_b200_cimage_header* hdr, uint8_t* data, size_t data_capacity, size_t* data_len) {
	if (!filename || !hdr || !data_len) {
		return ASTCENC_ERR_BAD_PARAM;
	}
	*data_len = 0;
	FILE* f = fopen(filename, "rb");
	if (!f) {
		return ASTCENC_ERR_BAD_PARAM;
	}
	RawHeader raw;
	astcenc_error status = ASTCENC_SUCCESS;
	size_t payload = 0;
	if (fread(&raw, sizeof(raw), 1, f) != 1) {
		status = ASTCENC_ERR_BAD_PARAM;
	} else {
		status = parse_header(raw, hdr, &payload);
	}
	if (status == ASTCENC_SUCCESS) {
		// a file shorter than its header promises is corrupt (:2674-2680); checked before anything is allocated or
		// read, so absurd sizes fail here too
		long at = ftell(f);
		if (at < 0 || fseek(f, 0, SEEK_END) != 0) {
			status = ASTCENC_ERR_BAD_PARAM;
		} else {
			long end = ftell(f);
			if (end < at || (size_t)(end - at) < payload) {
				status = ASTCENC_ERR_BAD_PARAM;
			}
			fseek(f, at, SEEK_SET);
		}
	}
	if (status == ASTCENC_SUCCESS) {
		*data_len = payload;
		if (data) {
			// with a buffer: read the payload (capacity too small: OUT_OF_MEM, *data_len says what is needed)
			if (data_capacity < payload) {
				status = ASTCENC_ERR_OUT_OF_MEM;
			} else if (payload && fread(data, 1, payload, f) != payload) {
				status = ASTCENC_ERR_BAD_PARAM;
			}
		}
	}
	fclose(f);
	return status;
}

// bytes of the block grid a header describes (false on zero dimensions / overflow)
static bool grid_bytes(const astcenc_b200_cimage_header& h, size_t* out) {
	if (h.block_x == 0 || h.block_y == 0 || h.block_z == 0 || h.dim_x == 0 || h.dim_y == 0 || h.dim_z == 0) {
		return false;
	}
	unsigned long long bx = ((unsigned long long)h.dim_x + h.block_x - 1) / h.block_x;
	unsigned long long by = ((unsigned long long)h.dim_y + h.block_y - 1) / h.block_y;
	unsigned long long bz = ((unsigned long long)h.dim_z + h.block_z - 1) / h.block_z;
	unsigned long long n = bx * by;
	if (by != 0 && n / by != bx) return false;
	unsigned long long m = n * bz;
	if (bz != 0 && m / bz != n) return false;
	if (m > (~0ull) / 16) return false;
	*out = (size_t)(m * 16);
	return true;
}

astcenc_error astcenc_b200_store_ktx_cimage(const char* filename, const astcenc_b200_cimage_header* hdr, int is_srgb, const uint8_t* data, size_t data_len) {
	if (!filename || !hdr || (!data && data_len) || data_len > 0xFFFFFFFFull || hdr->dim_x == 0 || hdr->dim_y == 0 || hdr->dim_z == 0) {
		return ASTCENC_ERR_BAD_PARAM;
	}
	uint32_t fmt = gl_format_of(hdr->block_x, hdr->block_y, hdr->block_z, is_srgb != 0);
	if (fmt == 0) {
		return ASTCENC_ERR_BAD_BLOCK_SIZE;
	}
	// the payload must be exactly the block grid of the header (as store_cimage checks); 2D footprints describe 2D images
	size_t want = 0;
	if (!grid_bytes(*hdr, &want) || want != data_len || (hdr->block_z == 1 && hdr->dim_z != 1)) {
		return ASTCENC_ERR_BAD_PARAM;
	}
	KtxHeader k;
	memcpy(k.magic, k_ktx_magic, 12);
	k.endianness = 0x04030201u;
	k.gl_type = 0;
	k.gl_type_size = 1;
	k.gl_format = 0;
	k.gl_internal_format = fmt;
	k.gl_base_internal_format = k_gl_rgba;
	k.pixel_width = hdr->dim_x;
	k.pixel_height = hdr->dim_y;
	k.pixel_depth = hdr->dim_z == 1 ? 0 : hdr->dim_z;
	k.number_of_array_elements = 0;
	k.number_of_faces = 1;
	k.number_of_mipmap_levels = 1;
	k.bytes_of_key_value_data = 0;
	FILE* f = fopen(filename, "wb");
	if (!f) {
		return ASTCENC_ERR_BAD_PARAM;
	}
	uint32_t len32 = (uint32_t)data_len;
	bool ok = fwrite(&k, sizeof(k), 1, f) == 1 && fwrite(&len32, 4, 1, f) == 1 && (data_len == 0 || fwrite(data, 1, data_len, f) == data_len);
	ok = (fclose(f) == 0) && ok;
	return ok ? ASTCENC_SUCCESS : ASTCENC_ERR_BAD_PARAM;
}

astcenc_error astcenc_b200_load_ktx_cimage(const char* filename, astcenc_b200_cimage_header* hdr, int* is_srgb, uint8_t* data, size_t data_capacity,
                                           size_t* data_len) {
	if (!filename || !hdr || !is_srgb || !data_len) {
		return ASTCENC_ERR_BAD_PARAM;
	}
	*data_len = 0;
	FILE* f = fopen(filename, "rb");
	if (!f) {
		return ASTCENC_ERR_BAD_PARAM;
	}
	astcenc_error status = ASTCENC_SUCCESS;
	KtxHeader k;
	bool swapped = false;
	if (fread(&k, sizeof(k), 1, f) != 1 || memcmp(k.magic, k_ktx_magic, 12) != 0 || (k.endianness != 0x04030201u && k.endianness != 0x01020304u)) {
		status = ASTCENC_ERR_BAD_PARAM;
	} else {
		if (k.endianness == 0x01020304u) {
			// written on a machine of the other byte order: every 32-bit field is reversed (:1321-1326)
			swapped = true;
			uint32_t* fields = &k.endianness;
			for (int i = 0; i < 13; i++) {
				fields[i] = bswap32(fields[i]);
			}
		}
		if (k.gl_type != 0 || k.gl_format != 0 || k.gl_type_size != 1 || k.gl_base_internal_format != k_gl_rgba ||
		    !footprint_of(k.gl_internal_format, hdr, is_srgb) || k.pixel_width == 0 || k.pixel_height == 0) {
			status = ASTCENC_ERR_BAD_PARAM;      // not an ASTC payload this library understands
		}
	}
	uint32_t len32 = 0;
	if (status == ASTCENC_SUCCESS) {
		hdr->dim_x = k.pixel_width;
		hdr->dim_y = k.pixel_height;
		hdr->dim_z = k.pixel_depth == 0 ? 1 : k.pixel_depth;
		if (fseek(f, (long)k.bytes_of_key_value_data, SEEK_CUR) != 0 || fread(&len32, 4, 1, f) != 1) {
			status = ASTCENC_ERR_BAD_PARAM;
		} else if (swapped) {
			len32 = bswap32(len32);
		}
	}
	if (status == ASTCENC_SUCCESS) {
		long at = ftell(f);
		if (at < 0 || fseek(f, 0, SEEK_END) != 0) {
			status = ASTCENC_ERR_BAD_PARAM;
		} else {
			long end = ftell(f);
			if (end < at || (size_t)(end - at) < (size_t)len32) {
				status = ASTCENC_ERR_BAD_PARAM;      // truncated payload
			}
			// the first mip level must hold the block grid of the header (a larger value would be trailing garbage,
			// a smaller one a short payload that only fails later in decompression)
			size_t want = 0;
			if (status == ASTCENC_SUCCESS && (!grid_bytes(*hdr, &want) || want != (size_t)len32)) {
				status = ASTCENC_ERR_BAD_PARAM;
			}
			fseek(f, at, SEEK_SET);
		}
	}
	if (status == ASTCENC_SUCCESS) {
		*data_len = len32;
		if (data) {
			if (data_capacity < (size_t)len32) {
				status = ASTCENC_ERR_OUT_OF_MEM;
			} else if (len32 && fread(data, 1, len32, f) != len32) {
				status = ASTCENC_ERR_BAD_PARAM;
			}
		}
	}
	fclose(f);
	return status;
}

}
