"""Dev tool: compress ONE block with the tracing build and print the intermediate values."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from astc_ref import *
import astc_images as I
img = I.photo_like(24, 24, seed=11)[:4, :4].copy() if len(sys.argv) < 3 else I.photo_like(24, 24, seed=11)[:6, :6].copy()
bx = img.shape[0]
q = 10.0 if bx == 4 else 60.0
if sys.argv[1] == "gpu":
    lib = AstcencLib(os.path.join(ROOT, "astc-encoder_b200", "libastcenc_b200_trace.so"))
    out = lib.compress(img, PRF_LDR, bx, bx, q, FLG_SELF_DECOMPRESS_ONLY)
else:
    hs = C.CDLL(os.path.join(ROOT, "tests", "_build", "libhostsim_trace.so"))
    hs.hostsim_compress_image.argtypes = [C.c_int, C.c_uint, C.c_uint, C.c_float, C.c_uint, C.c_void_p, C.c_int, C.c_uint, C.c_uint, C.POINTER(C.c_int), C.c_void_p]
    out = np.zeros(16, np.uint8)
    hs.hostsim_compress_image(PRF_LDR, bx, bx, q, FLG_SELF_DECOMPRESS_ONLY, img.ctypes.data, 0, bx, bx, None, out.ctypes.data)
sys.stdout.flush()
print("OUT", bytes(out).hex())
