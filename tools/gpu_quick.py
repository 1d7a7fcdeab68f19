"""Quick on-GPU sanity run (dev tool): product library vs the reference build / oracle on a few images."""
import sys, time, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from astc_ref import *
import astc_images as I

PROD = os.environ.get("ASTC_LIB", os.path.join(ROOT, "astc-encoder_b200", "libastcenc_b200.so"))
QUICK = int(os.environ.get("ASTC_QUICK", "0"))
prod = AstcencLib(PROD)
chk = ref_lib() if have_ref() else None
orc = Oracle()

def check(name, img, prof, bx, by, q, fl=0, swz=(0, 1, 2, 3)):
    t0 = time.time(); g = prod.compress(img, prof, bx, by, q, fl | FLG_SELF_DECOMPRESS_ONLY, swz=swz); t1 = time.time()
    r = chk.compress(img, prof, bx, by, q, fl | FLG_SELF_DECOMPRESS_ONLY, swz=swz) if chk else orc.compress(img, prof, bx, by, q, fl | FLG_SELF_DECOMPRESS_ONLY, swz=list(swz))
    t2 = time.time()
    d = block_diff(g, r)
    print("%-22s %s %dx%d q=%g fl=%d blocks=%d diff=%d  gpu %.3fs ref %.3fs %s" % (name, img.shape, bx, by, q, fl, len(g) // 16, len(d), t1 - t0, t2 - t1, d[:6]))
    sys.stdout.flush()
    return len(d)

bad = 0
img = I.photo_like(256, 256)
if QUICK:
    small = I.photo_like(24 * QUICK, 24, seed=11)
    bad += check("tiny 4x4", small, PRF_LDR, 4, 4, 10)
    bad += check("tiny 6x6", small, PRF_LDR, 6, 6, 60)
    print("TOTAL DIFF BLOCKS", bad)
    sys.exit(1 if bad else 0)
bad += check("photo", img, PRF_LDR, 6, 6, 60)
bad += check("photo", img, PRF_LDR, 4, 4, 10)
bad += check("photo", img, PRF_LDR, 8, 8, 98)
bad += check("photo unorm8", img, PRF_LDR, 5, 5, 60, FLG_USE_DECODE_UNORM8)
bad += check("photo srgb 12x12", img, PRF_LDR_SRGB, 12, 12, 60)
bad += check("noise", I.uniform_noise(120, 120), PRF_LDR, 6, 6, 60)
bad += check("gradient", I.smooth_gradient(128, 128), PRF_LDR, 6, 6, 60)
bad += check("voronoi", I.voronoi_flat(126, 126), PRF_LDR, 6, 6, 60)
bad += check("const", I.constant(64, 64), PRF_LDR, 6, 6, 60)
bad += check("alpha mask odd", I.alpha_mask(101, 77), PRF_LDR, 6, 6, 60)
bad += check("hdr f16", I.hdr_noise(96, 96), PRF_HDR, 6, 6, 60)
bad += check("hdr f32 ldra", I.hdr_noise(96, 96, dtype=np.float32), PRF_HDR_RGB_LDR_A, 4, 4, 60)
bad += check("normal", img, PRF_LDR, 6, 6, 60, FLG_MAP_NORMAL, swz=(0, 0, 0, 1))
bad += check("rgbm", img, PRF_LDR, 6, 6, 60, FLG_MAP_RGBM)
big = I.photo_like(1024, 1024, seed=7)
bad += check("photo 1k", big, PRF_LDR, 6, 6, 60)
print("TOTAL DIFF BLOCKS", bad)
sys.exit(1 if bad else 0)
