"""Dev tool (GPU box): wide parity sweep - every 2D block size x preset x content kind x profile on small images, the
product against the reference build (or the oracle when the build is absent). Prints mismatching configurations.
    python tools/parity_sweep.py [size]"""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from astc_ref import *
import astc_images as I

size = int(sys.argv[1]) if len(sys.argv) > 1 else 72
prod = AstcencLib(os.path.join(ROOT, "astc-encoder_b200", "libastcenc_b200.so"))
chk = ref_lib() if have_ref() else None
orc = Oracle()
BLOCKS = [(4, 4), (5, 4), (5, 5), (6, 5), (6, 6), (8, 5), (8, 6), (8, 8), (10, 5), (10, 6), (10, 8), (10, 10), (12, 10), (12, 12)]
PRESETS = [("fastest", PRE_FASTEST), ("fast", PRE_FAST), ("medium", PRE_MEDIUM), ("thorough", PRE_THOROUGH), ("verythorough", PRE_VERYTHOROUGH),
           ("exhaustive", PRE_EXHAUSTIVE)]
if os.environ.get("SWEEP_PRESETS"):
    PRESETS = [p for p in PRESETS if p[0] in os.environ["SWEEP_PRESETS"].split(",")]
SEED = int(os.environ.get("SWEEP_SEED", "0"))
BUDGET_S = float(os.environ.get("SWEEP_SECONDS", "300"))
rng = np.random.default_rng(99)


def contents():
    w, h = size + 5, size          # not a multiple of most block sizes: edge blocks are padded
    S0 = SEED * 100
    N = (0, 1, 2, 3)
    yield "photo", I.photo_like(h, w, seed=S0 + 31), PRF_LDR, 0, N
    yield "photo-srgb", I.photo_like(h, w, seed=S0 + 32), PRF_LDR_SRGB, 0, N
    yield "noise", I.uniform_noise(h, w, seed=S0 + 33), PRF_LDR, 0, N
    yield "voronoi", I.voronoi_flat(h, w, seed=S0 + 34), PRF_LDR, 0, N
    yield "alpha-mask", I.alpha_mask(h, w, seed=S0 + 35), PRF_LDR, FLG_USE_ALPHA_WEIGHT, N
    yield "gradient", I.smooth_gradient(h, w, seed=S0 + 36), PRF_LDR, FLG_USE_PERCEPTUAL, N
    yield "normal-map", I.photo_like(h, w, seed=S0 + 37), PRF_LDR, FLG_MAP_NORMAL, (0, 0, 0, 1)      # the CLI's -normal swizzle rrrg
    yield "hdr-f16", I.hdr_noise(h, w, seed=S0 + 38), PRF_HDR, 0, N
    yield "hdr-ldra-f32", I.hdr_noise(h, w, seed=S0 + 39, dtype=np.float32), PRF_HDR_RGB_LDR_A, 0, N
    yield "rgbm", I.photo_like(h, w, seed=S0 + 40), PRF_LDR, FLG_MAP_RGBM, N
    yield "unorm8", I.photo_like(h, w, seed=S0 + 41), PRF_LDR, FLG_USE_DECODE_UNORM8, N
    yield "luminance rrr1", I.photo_like(h, w, seed=S0 + 42), PRF_LDR, 0, (0, 0, 0, 5)
    yield "swizzle bgra", I.alpha_mask(h, w, seed=S0 + 43), PRF_LDR, 0, (2, 1, 0, 3)
    yield "hdr-f32 in LDR profile", np.clip(I.hdr_noise(h, w, seed=S0 + 44, dtype=np.float32), 0, 1), PRF_LDR, 0, N


bad = 0
total = 0
skipped = 0
t00 = time.time()
for cname, img, prof, fl, swz in contents():
    for bx, by in BLOCKS:
        for pname, q in PRESETS:
            if time.time() - t00 > BUDGET_S:
                continue
            try:
                g = prod.compress(img, prof, bx, by, q, fl | FLG_SELF_DECOMPRESS_ONLY, swz=swz)
            except RuntimeError as e:
                skipped += 1
                print("SKIP %-14s %2dx%-2d %-12s %s" % (cname, bx, by, pname, e)); sys.stdout.flush()
                continue
            r = chk.compress(img, prof, bx, by, q, fl | FLG_SELF_DECOMPRESS_ONLY, swz=swz, threads=16) if chk else orc.compress(img, prof, bx, by, q, fl | FLG_SELF_DECOMPRESS_ONLY, swz=list(swz))
            d = block_diff(g, r)
            total += 1
            if len(d):
                bad += 1
                print("DIFF %-14s %2dx%-2d %-12s blocks %d of %d: %s" % (cname, bx, by, pname, len(d), len(g) // 16, d[:5])); sys.stdout.flush()
    print("done %-14s %.0f s" % (cname, time.time() - t00)); sys.stdout.flush()
print("configurations %d, mismatching %d, skipped %d" % (total, bad, skipped))
