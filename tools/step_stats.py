"""Dev tool: run the -DASTC_STEP_STATS build (make libastcenc_b200_stats.so) on one image; the emit kernel prints the
refinement-step timing by step kind."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from astc_ref import *
import astc_images as I
prod = AstcencLib(os.path.join(ROOT, "astc-encoder_b200", "libastcenc_b200_stats.so"))
img = I.photo_like(2048, 2048, seed=3)
prod.compress(img, PRF_LDR, 6, 6, PRE_MEDIUM, FLG_SELF_DECOMPRESS_ONLY)
