"""Dev tool: run the product library on one image with search-limiting config overrides (to vary the hot code
footprint) - meant to be run under `ncu --metrics ...` to compare issue rates and instruction-cache hit rates."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from astc_ref import *
import astc_images as I

prod = AstcencLib(os.path.join(ROOT, "astc-encoder_b200", "libastcenc_b200.so"))
img = I.photo_like(2048, 2048, seed=3)
variants = {
    "full": {},
    "1part": dict(tune_partition_count_limit=1),
    "1part_no2plane": dict(tune_partition_count_limit=1, tune_2plane_early_out_limit_correlation=0.0),
    "1part_no2plane_1cand_1ref": dict(tune_partition_count_limit=1, tune_2plane_early_out_limit_correlation=0.0, tune_candidate_limit=1, tune_refinement_limit=1),
}
which = sys.argv[1:] or list(variants)
for name in which:
    t0 = time.time()
    prod.compress(img, PRF_LDR, 6, 6, PRE_MEDIUM, FLG_SELF_DECOMPRESS_ONLY, **variants[name])
    print(name, "%.3f s" % (time.time() - t0)); sys.stdout.flush()
