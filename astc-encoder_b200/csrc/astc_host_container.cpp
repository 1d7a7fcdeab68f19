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

}
