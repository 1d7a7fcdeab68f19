"""Dev tool (GPU box): parity sweep of the 3D block sizes - every footprint x preset x content kind on small volumes with ragged
edges, the product against the reference build. Prints mismatching configurations.
    python tools/parity_sweep_3d.py"""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from astc_ref import *
import astc_images as I
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
ref = ref_lib()
FOOT = [(3, 3, 3), (4, 3, 3), (4, 4, 3), (4, 4, 4), (5, 4, 4), (5, 5, 4), (5, 5, 5), (6, 5, 5), (6, 6, 5), (6, 6, 6)]
PRESETS = [("fastest", PRE_FASTEST), ("fast", PRE_FAST), ("medium", PRE_MEDIUM), ("thorough", PRE_THOROUGH)]
W, H, D = 31, 26, 13


def vol(kind, seed):
    if kind == "hdr":
        return np.ascontiguousarray(I.hdr_noise(H * D, W, seed=seed).reshape(D, H, W, 4))
    return np.ascontiguousarray(getattr(I, kind)(H * D, W, seed=seed).reshape(D, H, W, 4))


KINDS = [("photo_like", PRF_LDR, 0), ("uniform_noise", PRF_LDR_SRGB, 0), ("voronoi_flat", PRF_LDR, 0), ("alpha_mask", PRF_LDR, FLG_USE_ALPHA_WEIGHT),
         ("smooth_gradient", PRF_LDR, FLG_USE_PERCEPTUAL), ("hdr", PRF_HDR, 0), ("hdr", PRF_HDR_RGB_LDR_A, 0)]
n = bad = 0
t0 = time.time()
for fp in FOOT:
    for pname, q in PRESETS:
        for k, (kind, prof, fl) in enumerate(KINDS):
            v = vol(kind, 100 + k)
            ctx = pkg.Context(pkg.config_init(prof, fp[0], fp[1], q, fl, block_z=fp[2]))
            try:
                got = ctx.compress_image(v)
            finally:
                ctx.close()
            want = ref.compress_volume(v, prof, fp[0], fp[1], fp[2], q, flags=fl, threads=8)
            d = len(block_diff(got, want))
            n += 1
            if d:
                bad += 1
                print("DIFF %dx%dx%d %s %s prof %d: %d of %d blocks" % (fp + (pname, kind, prof, d, len(got) // 16)), flush=True)
print("3D sweep: %d configurations (%d footprints x %d presets x %d content kinds / profiles, %dx%dx%d volumes), %d mismatching, %.0f s" %
      (n, len(FOOT), len(PRESETS), len(KINDS), W, H, D, bad, time.time() - t0))
