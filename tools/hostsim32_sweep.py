"""Dev tool (CPU): randomised parity sweep of the 32-lane host simulation (tests/_build/libhostsim32.so, built by the tests) against
the oracle: random footprint, preset, content kind, flags and (small, non-multiple) image size.
    python tools/hostsim32_sweep.py [seconds]"""
import sys, time, ctypes as C, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
from astc_ref import *
import astc_images as I
lib = C.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', '_build', 'libhostsim32.so'))
lib.hostsim_compress_image.argtypes = [C.c_int, C.c_uint, C.c_uint, C.c_float, C.c_uint, C.c_void_p, C.c_int, C.c_uint, C.c_uint, C.POINTER(C.c_int), C.c_void_p]
orc = Oracle()
S = FLG_SELF_DECOMPRESS_ONLY
rng = np.random.default_rng(2026)
BLOCKS = [(4, 4), (5, 4), (5, 5), (6, 5), (6, 6), (8, 5), (8, 6), (8, 8), (10, 5), (10, 6), (10, 8), (10, 10), (12, 10), (12, 12)]
PRE = [PRE_FASTEST, PRE_FAST, PRE_MEDIUM, PRE_THOROUGH]
t_end = time.time() + float(sys.argv[1]) if len(sys.argv) > 1 else time.time() + 600
n = bad = 0
while time.time() < t_end:
    bx, by = BLOCKS[rng.integers(len(BLOCKS))]
    q = PRE[rng.integers(len(PRE))]
    kind = int(rng.integers(8))
    h, w = by * int(rng.integers(1, 4)) + int(rng.integers(0, 3)), bx * int(rng.integers(1, 4)) + int(rng.integers(0, 3))
    seed = int(rng.integers(1 << 30))
    fl, prof, swz = S, PRF_LDR, None
    if kind == 0: img = I.photo_like(h, w, seed=seed)
    elif kind == 1: img = I.uniform_noise(h, w, seed=seed)
    elif kind == 2: img = I.voronoi_flat(h, w, cell=max(3, bx - 1), seed=seed)
    elif kind == 3: img = I.alpha_mask(h, w, seed=seed); fl |= FLG_USE_ALPHA_WEIGHT
    elif kind == 4: img = I.hdr_noise(h, w, seed=seed); prof = PRF_HDR
    elif kind == 5: img = I.hdr_noise(h, w, seed=seed, dtype=np.float32); prof = PRF_HDR_RGB_LDR_A
    elif kind == 6: img = I.photo_like(h, w, seed=seed); fl |= FLG_MAP_NORMAL; swz = (0, 0, 0, 1)
    else: img = I.smooth_gradient(h, w, seed=seed); prof = PRF_LDR_SRGB; fl |= FLG_USE_PERCEPTUAL
    img = np.ascontiguousarray(img)
    nb = ((w + bx - 1) // bx) * ((h + by - 1) // by)
    out = np.zeros(nb * 16, np.uint8)
    dt = {np.dtype(np.uint8): 0, np.dtype(np.float16): 1, np.dtype(np.float32): 2}[img.dtype]
    sw = (C.c_int * 4)(*swz) if swz else None
    rc = lib.hostsim_compress_image(prof, bx, by, q, fl, img.ctypes.data, dt, w, h, sw, out.ctypes.data)
    want = orc.compress(img, prof, bx, by, q, fl, swz=list(swz) if swz else None)
    d = len(block_diff(out, want))
    n += 1
    if rc != 0 or d:
        bad += 1
        print("DIFF", bx, by, q, kind, h, w, seed, rc, d, flush=True)
print("runs", n, "bad", bad)
