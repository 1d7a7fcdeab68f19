"""Dev tool: time the kernel on a 1024^2 photo (29k blocks), normal vs coherence probe (set by env)."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from astc_ref import *
import astc_images as I
prod = AstcencLib(os.path.join(ROOT, "astc-encoder_b200", "libastcenc_b200.so"))
img = I.photo_like(1024, 1024, seed=7)
for i in range(3):
    t0 = time.time()
    prod.compress(img, PRF_LDR, 6, 6, PRE_MEDIUM, FLG_SELF_DECOMPRESS_ONLY)
    print(os.environ.get("ASTCENC_B200_COHERENCE_PROBE", "0"), os.environ.get("ASTCENC_B200_WARPS", "-"), "%.4f s" % (time.time() - t0)); sys.stdout.flush()
