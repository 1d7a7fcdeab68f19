// TEST TOOL: stand-alone driver of the 32-lane host simulation (tests/hostsim/hostsim.cpp built with
// -DASTC_HOSTSIM_LANES32) for sanitizer runs:
//   g++ -std=c++14 -O1 -g -fsanitize=thread -pthread -ffp-contract=off -DASTC_HOSTSIM_LANES32=1 -x c++ \
//       tests/hostsim/lanes32_main.cpp tests/hostsim/hostsim.cpp astc-encoder_b200/csrc/astc_host_tables.cpp \
//       astc-encoder_b200/csrc/astc_host_config.cpp -o /tmp/lanes32_tsan
//   /tmp/lanes32_tsan <profile 0-3> <block_x> <block_y> <quality> <width> <height> <seed> [kind: 0 noise, 1 two-colour stripes, 2 hdr,
//                     3 = stripes + decode of the result, 4 = alpha-scale pre-pass (radius 2) on stripes with transparent rows]
//                    [block_z depth: a 3D block size on a volume of <height> x <depth> rows of noise / stripes, compressed and decoded]
// ThreadSanitizer sees every shared-memory access of every lane; the warp collectives are its only synchronisation, so a
// report is a missing __syncwarp() between a producer lane and a consumer lane. Prints an FNV hash of the output blocks.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

extern "C" int hostsim_compress_image(int profile, unsigned int bx, unsigned int by, float quality, unsigned int flags,
                                      const void* data, int data_type, unsigned int dim_x, unsigned int dim_y, const int* swz, uint8_t* out);
extern "C" int hostsim_decompress_image(int profile, unsigned int bx, unsigned int by, unsigned int flags, const uint8_t* blocks, void* out, int data_type,
                                        unsigned int dim_x, unsigned int dim_y, const int* swz);
extern "C" void hostsim_set_a_scale_radius(unsigned int r);
extern "C" int hostsim_compress_volume(int profile, unsigned int bx, unsigned int by, unsigned int bz, float quality, unsigned int flags,
                                       const void* data, int data_type, unsigned int dim_x, unsigned int dim_y, unsigned int dim_z, const int* swz, uint8_t* out);
extern "C" int hostsim_decompress_volume(int profile, unsigned int bx, unsigned int by, unsigned int bz, unsigned int flags, const uint8_t* blocks, void* out, int data_type,
                                         unsigned int dim_x, unsigned int dim_y, unsigned int dim_z, const int* swz);

int main(int argc, char** argv) {
	if (argc < 8) {
		fprintf(stderr, "usage: %s profile bx by quality width height seed [kind]\n", argv[0]);
		return 2;
	}
	int profile = atoi(argv[1]);
	unsigned int bx = (unsigned int)atoi(argv[2]), by = (unsigned int)atoi(argv[3]);
	float quality = (float)atof(argv[4]);
	unsigned int w = (unsigned int)atoi(argv[5]), h = (unsigned int)atoi(argv[6]);
	uint32_t s = (uint32_t)atoi(argv[7]) * 2654435761u + 12345u;
	int kind = argc > 8 ? atoi(argv[8]) : 0;
	auto rnd = [&]() { s = s * 1664525u + 1013904223u; return s >> 8; };
	std::vector<uint8_t> img8((size_t)w * h * 4);
	std::vector<uint16_t> img16((size_t)w * h * 4);
	uint8_t cell[2][4];
	for (int c = 0; c < 2; c++) for (int k = 0; k < 4; k++) cell[c][k] = (uint8_t)rnd();
	for (unsigned int y = 0; y < h; y++) {
		for (unsigned int x = 0; x < w; x++) {
			for (int k = 0; k < 4; k++) {
				size_t i = ((size_t)y * w + x) * 4 + k;
				if (kind == 1 || kind == 3 || kind == 4) {
					img8[i] = (uint8_t)(cell[((x * 3 + y * 5) / 7) & 1][k] + (rnd() & 7));      // two colours in diagonal stripes + a little noise
				} else {
					img8[i] = (uint8_t)rnd();
				}
				// half floats: sign 0, exponent 5..20, random mantissa (finite, positive)
				img16[i] = (uint16_t)(((5 + rnd() % 16) << 10) | (rnd() & 0x3FF));
			}
		}
	}
	unsigned int flags = 32;      // ASTCENC_FLG_SELF_DECOMPRESS_ONLY
	if (kind == 4) {
		for (unsigned int y = 0; y < h; y += 3) {
			for (unsigned int x = 0; x < w; x++) img8[((size_t)y * w + x) * 4 + 3] = 0;      // transparent rows
		}
		flags |= 4;               // ASTCENC_FLG_USE_ALPHA_WEIGHT
		hostsim_set_a_scale_radius(2);
	}
	if (kind == 3) {
		flags = 0;                // full tables: the result is decoded below
	}
	if (argc > 10) {
		// 3D block size: the image rows are cut into `depth` slices of h / depth rows
		unsigned int bz = (unsigned int)atoi(argv[9]), depth = (unsigned int)atoi(argv[10]);
		unsigned int hs = h / depth;
		unsigned int nb3 = ((w + bx - 1) / bx) * ((hs + by - 1) / by) * ((depth + bz - 1) / bz);
		std::vector<uint8_t> out3((size_t)nb3 * 16);
		int rc3 = hostsim_compress_volume(profile, bx, by, bz, quality, 0, img8.data(), 0, w, hs, depth, nullptr, out3.data());
		uint64_t hash3 = 1469598103934665603ull;
		for (uint8_t b : out3) { hash3 = (hash3 ^ b) * 1099511628211ull; }
		if (rc3 == 0) {
			std::vector<uint8_t> dec((size_t)w * hs * depth * 4);
			rc3 = hostsim_decompress_volume(profile, bx, by, bz, 0, out3.data(), dec.data(), 0, w, hs, depth, nullptr);
			for (uint8_t b : dec) { hash3 = (hash3 ^ b) * 1099511628211ull; }
		}
		printf("rc %d blocks %u hash %016llx\n", rc3, nb3, (unsigned long long)hash3);
		return rc3;
	}
	unsigned int nb = ((w + bx - 1) / bx) * ((h + by - 1) / by);
	std::vector<uint8_t> out((size_t)nb * 16);
	int rc = hostsim_compress_image(profile, bx, by, quality, flags, kind == 2 ? (const void*)img16.data() : (const void*)img8.data(),
	                                kind == 2 ? 1 : 0, w, h, nullptr, out.data());
	uint64_t hash = 1469598103934665603ull;
	for (uint8_t b : out) { hash = (hash ^ b) * 1099511628211ull; }
	if (kind == 3 && rc == 0) {
		std::vector<uint8_t> dec((size_t)w * h * 4);
		rc = hostsim_decompress_image(profile, bx, by, 0, out.data(), dec.data(), 0, w, h, nullptr);
		for (uint8_t b : dec) { hash = (hash ^ b) * 1099511628211ull; }
	}
	printf("rc %d blocks %u hash %016llx\n", rc, nb, (unsigned long long)hash);
	return rc;
}
