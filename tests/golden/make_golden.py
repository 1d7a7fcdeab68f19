"""Generates tests/golden/golden_blocks.npz from the UNMODIFIED reference build (oracle/_ref, built from
/root/reference by oracle/Makefile). Run here, where the reference exists; the fixture travels to the GPU box.

Each case = (image generator, size, seed) x (profile, block size, quality, flags, swizzle); stored value = the
reference's physical blocks (16 bytes per block). The same case list is consumed by tests/test_oracle_golden.py
(oracle vs fixture, CPU) and tests/test_gpu_parity.py (CUDA path vs fixture, GPU)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import astc_images as I  # noqa: E402
from astc_ref import *  # noqa: E402,F401,F403

S = FLG_SELF_DECOMPRESS_ONLY

# name, generator, (h, w), dtype, profile, bx, by, quality, flags, swizzle
CASES = [
    ("photo_4x4_fast", "photo_like", (64, 64), "u8", PRF_LDR, 4, 4, PRE_FAST, S, (0, 1, 2, 3)),
    ("photo_6x6_medium", "photo_like", (96, 96), "u8", PRF_LDR, 6, 6, PRE_MEDIUM, S, (0, 1, 2, 3)),
    ("photo_8x8_thorough", "photo_like", (64, 64), "u8", PRF_LDR, 8, 8, PRE_THOROUGH, S, (0, 1, 2, 3)),
    ("photo_5x4_fastest", "photo_like", (40, 40), "u8", PRF_LDR, 5, 4, PRE_FASTEST, S, (0, 1, 2, 3)),
    ("photo_5x5_q35", "photo_like", (50, 50), "u8", PRF_LDR, 5, 5, 35.0, S, (0, 1, 2, 3)),
    ("photo_6x5_verythorough", "photo_like", (30, 36), "u8", PRF_LDR, 6, 5, PRE_VERYTHOROUGH, S, (0, 1, 2, 3)),
    ("photo_8x5_medium_unorm8", "photo_like", (40, 48), "u8", PRF_LDR, 8, 5, PRE_MEDIUM, S | FLG_USE_DECODE_UNORM8, (0, 1, 2, 3)),
    ("photo_8x6_srgb", "photo_like", (48, 48), "u8", PRF_LDR_SRGB, 8, 6, PRE_MEDIUM, S, (0, 1, 2, 3)),
    ("photo_10x5_perceptual", "photo_like", (50, 50), "u8", PRF_LDR, 10, 5, PRE_MEDIUM, S | FLG_USE_PERCEPTUAL, (0, 1, 2, 3)),
    ("photo_10x6_alphaweight", "alpha_mask", (48, 50), "u8", PRF_LDR, 10, 6, PRE_MEDIUM, S | FLG_USE_ALPHA_WEIGHT, (0, 1, 2, 3)),
    ("photo_10x8_medium", "photo_like", (64, 60), "u8", PRF_LDR, 10, 8, PRE_MEDIUM, S, (0, 1, 2, 3)),
    ("photo_10x10_medium", "photo_like", (60, 60), "u8", PRF_LDR, 10, 10, PRE_MEDIUM, S, (0, 1, 2, 3)),
    ("photo_12x10_medium", "photo_like", (60, 60), "u8", PRF_LDR, 12, 10, PRE_MEDIUM, S, (0, 1, 2, 3)),
    ("photo_12x12_thorough", "photo_like", (60, 60), "u8", PRF_LDR, 12, 12, PRE_THOROUGH, S, (0, 1, 2, 3)),
    ("photo_6x6_exhaustive", "photo_like", (24, 24), "u8", PRF_LDR, 6, 6, PRE_EXHAUSTIVE, S, (0, 1, 2, 3)),
    ("noise_6x6_medium", "uniform_noise", (48, 48), "u8", PRF_LDR, 6, 6, PRE_MEDIUM, S, (0, 1, 2, 3)),
    ("noise_4x4_thorough", "uniform_noise", (32, 32), "u8", PRF_LDR, 4, 4, PRE_THOROUGH, S, (0, 1, 2, 3)),
    ("gradient_6x6_medium", "smooth_gradient", (60, 60), "u8", PRF_LDR, 6, 6, PRE_MEDIUM, S, (0, 1, 2, 3)),
    ("voronoi_6x6_medium", "voronoi_flat", (60, 60), "u8", PRF_LDR, 6, 6, PRE_MEDIUM, S, (0, 1, 2, 3)),
    ("voronoi_8x8_thorough", "voronoi_flat", (64, 64), "u8", PRF_LDR, 8, 8, PRE_THOROUGH, S, (0, 1, 2, 3)),
    ("constant_6x6", "constant", (13, 17), "u8", PRF_LDR, 6, 6, PRE_MEDIUM, S, (0, 1, 2, 3)),
    ("alphamask_odd_6x6", "alpha_mask", (53, 47), "u8", PRF_LDR, 6, 6, PRE_MEDIUM, S, (0, 1, 2, 3)),
    ("one_texel", "photo_like", (1, 1), "u8", PRF_LDR, 6, 6, PRE_MEDIUM, S, (0, 1, 2, 3)),
    ("normalmap_6x6", "photo_like", (60, 60), "u8", PRF_LDR, 6, 6, PRE_MEDIUM, S | FLG_MAP_NORMAL, (0, 0, 0, 1)),
    ("rgbm_6x6", "photo_like", (60, 60), "u8", PRF_LDR, 6, 6, PRE_MEDIUM, S | FLG_MAP_RGBM, (0, 1, 2, 3)),
    ("swizzle_bgr1", "photo_like", (48, 48), "u8", PRF_LDR, 6, 6, PRE_MEDIUM, S, (2, 1, 0, 5)),
    ("no_selfdecompress_flag", "photo_like", (48, 48), "u8", PRF_LDR, 6, 6, PRE_MEDIUM, 0, (0, 1, 2, 3)),
    ("ldr_f32_input", "photo_like", (48, 48), "f32", PRF_LDR, 6, 6, PRE_MEDIUM, S, (0, 1, 2, 3)),
    ("ldr_f16_input", "photo_like", (48, 48), "f16", PRF_LDR, 6, 6, PRE_MEDIUM, S, (0, 1, 2, 3)),
    ("hdr_f16_6x6_medium", "hdr_noise", (60, 60), "f16", PRF_HDR, 6, 6, PRE_MEDIUM, S, (0, 1, 2, 3)),
    ("hdr_f32_rgb_ldra_4x4", "hdr_noise", (32, 32), "f32", PRF_HDR_RGB_LDR_A, 4, 4, PRE_MEDIUM, S, (0, 1, 2, 3)),
    ("hdr_f16_8x8_thorough", "hdr_noise", (48, 48), "f16", PRF_HDR, 8, 8, PRE_THOROUGH, S, (0, 1, 2, 3)),
    ("naninf_ldr", "naninf", (24, 24), "f32", PRF_LDR, 6, 6, PRE_MEDIUM, S, (0, 1, 2, 3)),
    ("naninf_hdr", "naninf", (24, 24), "f32", PRF_HDR, 6, 6, PRE_MEDIUM, S, (0, 1, 2, 3)),
]


def make_image(gen, size, dtype, seed=1234):
    h, w = size
    if gen == "naninf":
        rng = np.random.default_rng(seed)
        img = rng.uniform(-1, 2, size=(h, w, 4)).astype(np.float32)
        img[5 % h, 5 % w, 0] = np.nan
        img[10 % h, 11 % w, 2] = np.inf
        img[20 % h, 20 % w, 3] = -np.inf
        return img
    if gen == "hdr_noise":
        return I.hdr_noise(h, w, seed, dtype={"f16": np.float16, "f32": np.float32}[dtype])
    img = getattr(I, gen)(h, w) if gen == "constant" else getattr(I, gen)(h, w, seed=seed)
    if dtype == "f32":
        return (img / 255.0).astype(np.float32)
    if dtype == "f16":
        return (img / 255.0).astype(np.float16)
    return img


def main():
    ref = ref_lib()
    out = {}
    for (name, gen, size, dtype, prof, bx, by, q, fl, swz) in CASES:
        img = make_image(gen, size, dtype)
        out[name] = ref.compress(img, prof, bx, by, q, fl, swz=swz)
        print(name, img.shape, img.dtype, len(out[name]) // 16, "blocks")
    np.savez_compressed(os.path.join(HERE, "golden_blocks.npz"), **out)


if __name__ == "__main__":
    main()
