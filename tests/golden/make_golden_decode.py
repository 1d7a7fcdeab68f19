"""Generates tests/golden/golden_decode.npz from the UNMODIFIED reference build (oracle/_ref): decompression fixtures.

Each case = a block stream (taken from golden_blocks.npz, i.e. real encodings, or seeded random / void-extent bytes, i.e.
mostly invalid encodings) x (profile, block size, output type, flags, swizzle); stored value = the bytes of the image the
reference's astcenc_decompress_image writes. Consumed by tests/test_decode.py (oracle and host simulation on CPU,
CUDA path on the GPU)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from astc_ref import *  # noqa: E402,F401,F403

# name, source ("golden:<case>" | "random" | "void"), (h, w), profile, bx, by, out_type, flags, swizzle
DCASES = [
    ("photo_4x4_u8", "golden:photo_4x4_fast", (64, 64), PRF_LDR, 4, 4, TYPE_U8, 0, (0, 1, 2, 3)),
    ("photo_6x6_u8", "golden:photo_6x6_medium", (96, 96), PRF_LDR, 6, 6, TYPE_U8, 0, (0, 1, 2, 3)),
    ("photo_6x6_f16", "golden:photo_6x6_medium", (96, 96), PRF_LDR, 6, 6, TYPE_F16, 0, (0, 1, 2, 3)),
    ("photo_6x6_f32_srgb", "golden:photo_6x6_medium", (96, 96), PRF_LDR_SRGB, 6, 6, TYPE_F32, 0, (0, 1, 2, 3)),
    ("photo_8x8_u8_selfdec", "golden:photo_8x8_thorough", (64, 64), PRF_LDR, 8, 8, TYPE_U8, FLG_SELF_DECOMPRESS_ONLY, (0, 1, 2, 3)),
    ("photo_12x12_u8", "golden:photo_12x12_thorough", (60, 60), PRF_LDR, 12, 12, TYPE_U8, 0, (0, 1, 2, 3)),
    ("photo_10x5_f16_deconly", "golden:photo_10x5_perceptual", (50, 50), PRF_LDR, 10, 5, TYPE_F16, FLG_DECOMPRESS_ONLY, (0, 1, 2, 3)),
    ("voronoi_8x8_u8", "golden:voronoi_8x8_thorough", (64, 64), PRF_LDR, 8, 8, TYPE_U8, 0, (0, 1, 2, 3)),
    ("normalmap_6x6_z", "golden:normalmap_6x6", (60, 60), PRF_LDR, 6, 6, TYPE_U8, 0, (0, 3, 6, 5)),
    ("normalmap_6x6_z_f32", "golden:normalmap_6x6", (60, 60), PRF_LDR, 6, 6, TYPE_F32, 0, (0, 3, 6, 4)),
    ("swizzle_bgr1", "golden:photo_6x6_medium", (96, 96), PRF_LDR, 6, 6, TYPE_U8, 0, (2, 1, 0, 5)),
    ("odd_size_6x6", "golden:alphamask_odd_6x6", (53, 47), PRF_LDR, 6, 6, TYPE_U8, 0, (0, 1, 2, 3)),
    ("one_texel", "golden:one_texel", (1, 1), PRF_LDR, 6, 6, TYPE_F16, 0, (0, 1, 2, 3)),
    ("hdr_6x6_f16", "golden:hdr_f16_6x6_medium", (60, 60), PRF_HDR, 6, 6, TYPE_F16, 0, (0, 1, 2, 3)),
    ("hdr_6x6_f32", "golden:hdr_f16_6x6_medium", (60, 60), PRF_HDR, 6, 6, TYPE_F32, 0, (0, 1, 2, 3)),
    ("hdr_6x6_u8", "golden:hdr_f16_6x6_medium", (60, 60), PRF_HDR, 6, 6, TYPE_U8, 0, (0, 1, 2, 3)),
    ("hdr_ldra_4x4_f16", "golden:hdr_f32_rgb_ldra_4x4", (32, 32), PRF_HDR_RGB_LDR_A, 4, 4, TYPE_F16, 0, (0, 1, 2, 3)),
    ("hdr_blocks_as_ldr", "golden:hdr_f16_8x8_thorough", (48, 48), PRF_LDR, 8, 8, TYPE_F16, 0, (0, 1, 2, 3)),
    ("random_4x4_ldr_u8", "random", (40, 44), PRF_LDR, 4, 4, TYPE_U8, 0, (0, 1, 2, 3)),
    ("random_6x6_hdr_f16", "random", (60, 66), PRF_HDR, 6, 6, TYPE_F16, 0, (0, 1, 2, 3)),
    ("random_8x5_ldra_f32", "random", (40, 48), PRF_HDR_RGB_LDR_A, 8, 5, TYPE_F32, 0, (0, 1, 2, 3)),
    ("random_12x12_srgb_u8", "random", (96, 96), PRF_LDR_SRGB, 12, 12, TYPE_U8, 0, (0, 1, 2, 3)),
    ("random_6x6_selfdec", "random", (60, 60), PRF_LDR, 6, 6, TYPE_U8, FLG_SELF_DECOMPRESS_ONLY, (0, 1, 2, 3)),
    ("void_6x6_ldr_u8", "void", (48, 48), PRF_LDR, 6, 6, TYPE_U8, 0, (0, 1, 2, 3)),
    ("void_6x6_hdr_f16", "void", (48, 48), PRF_HDR, 6, 6, TYPE_F16, 0, (0, 1, 2, 3)),
    ("void_5x5_hdr_u8", "void", (50, 50), PRF_HDR, 5, 5, TYPE_U8, 0, (0, 1, 2, 3)),
]


def make_blocks(source, size, bx, by, golden, seed):
    h, w = size
    n = ((w + bx - 1) // bx) * ((h + by - 1) // by)
    if source.startswith("golden:"):
        b = golden[source[7:]]
        assert b.size == n * 16, (source, b.size, n)
        return b
    rng = np.random.default_rng(seed)
    rb = rng.integers(0, 256, size=n * 16, dtype=np.uint8)
    if source == "random":
        return rb
    vb = rb.reshape(-1, 16).copy()     # void-extent (constant colour) headers: U16 / F16, valid / reserved-bit / extent errors
    vb[:, 0] = 0xFC
    vb[:, 1] = (vb[:, 1] & 0xFC) | 0x01 | (rng.integers(0, 2, size=len(vb), dtype=np.uint8) << 1)
    vb[::2, 1] |= 0x0C
    vb[::4, 2:8] = 0xFF
    vb[::4, 1] |= 0xF0
    return vb.reshape(-1)


def main():
    ref = ref_lib()
    golden = np.load(os.path.join(HERE, "golden_blocks.npz"))
    out = {}
    for i, (name, source, size, prof, bx, by, ot, fl, swz) in enumerate(DCASES):
        blocks = make_blocks(source, size, bx, by, golden, 100 + i)
        img = ref.decompress(blocks, size[1], size[0], prof, bx, by, ot, fl, swz)
        out[name] = np.ascontiguousarray(img).view(np.uint8).reshape(-1)
        print(name, blocks.size // 16, "blocks ->", out[name].size, "bytes")
    np.savez_compressed(os.path.join(HERE, "golden_decode.npz"), **out)


if __name__ == "__main__":
    main()
