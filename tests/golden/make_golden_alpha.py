"""Generates tests/golden/golden_alpha.npz from the UNMODIFIED reference build (oracle/_ref): the alpha-scale pre-pass
(astcenc_config::a_scale_radius, CLI `-a <radius>`). Stored value = the reference's physical blocks.
Consumed by tests/test_alpha_scale.py."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import astc_images as I  # noqa: E402
from astc_ref import *  # noqa: E402,F401,F403

FL = FLG_USE_ALPHA_WEIGHT | FLG_SELF_DECOMPRESS_ONLY
# name, (h, w), dtype, profile, bx, by, quality, flags, swizzle, radius
ACASES = [
    ("a1_6x6_medium", (70, 90), "u8", PRF_LDR, 6, 6, PRE_MEDIUM, FL, (0, 1, 2, 3), 1),
    ("a2_6x6_medium", (70, 90), "u8", PRF_LDR, 6, 6, PRE_MEDIUM, FL, (0, 1, 2, 3), 2),
    ("a1_4x4_fast", (64, 64), "u8", PRF_LDR, 4, 4, PRE_FAST, FL, (0, 1, 2, 3), 1),
    ("a3_8x8_medium_tiles", (100, 133), "u8", PRF_LDR, 8, 8, PRE_MEDIUM, FL, (0, 1, 2, 3), 3),
    ("a1_5x4_odd", (65, 97), "u8", PRF_LDR, 5, 4, PRE_MEDIUM, FL, (0, 1, 2, 3), 1),
    ("a1_6x6_f16", (70, 90), "f16", PRF_LDR, 6, 6, PRE_MEDIUM, FL, (0, 1, 2, 3), 1),
    ("a2_6x6_f32_hdr", (48, 60), "f32", PRF_HDR, 6, 6, PRE_MEDIUM, FL, (0, 1, 2, 3), 2),
    ("a8_6x6_wide", (33, 65), "u8", PRF_LDR, 6, 6, PRE_MEDIUM, FL, (0, 1, 2, 3), 8),
    ("a1_swizzle_alpha_from_red", (60, 60), "u8", PRF_LDR, 6, 6, PRE_MEDIUM, FL, (0, 1, 2, 0), 1),
    ("a1_no_alpha_weight_flag", (60, 60), "u8", PRF_LDR, 6, 6, PRE_MEDIUM, FLG_SELF_DECOMPRESS_ONLY, (0, 1, 2, 3), 1),
]


def make_alpha_image(size, dtype, seed):
    """alpha-mask image with fully transparent regions and a few isolated opaque texels (they keep their neighbourhood alive)"""
    h, w = size
    rng = np.random.default_rng(seed)
    img = I.alpha_mask(h, w).copy()
    img[:h // 2, :w // 3, 3] = 0
    img[h // 2:, w // 2:, 3] = 0
    img[rng.integers(0, h, size=6), rng.integers(0, w, size=6), 3] = 255
    if dtype == "u8":
        return img
    f = img.astype(np.float32) / 255.0
    return f.astype(np.float16 if dtype == "f16" else np.float32)


def main():
    ref = ref_lib()
    out = {}
    for i, (name, size, dt, prof, bx, by, q, fl, swz, r) in enumerate(ACASES):
        img = make_alpha_image(size, dt, 40 + i)
        out[name] = ref.compress(img, prof, bx, by, q, fl, swz=swz, a_scale_radius=r)
        plain = ref.compress(img, prof, bx, by, q, fl, swz=swz)
        print(name, out[name].size // 16, "blocks,", len(block_diff(out[name], plain)), "changed by the pre-pass")
    np.savez_compressed(os.path.join(HERE, "golden_alpha.npz"), **out)


if __name__ == "__main__":
    main()
