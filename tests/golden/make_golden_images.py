"""Generates tests/golden/golden_images.npz: REAL-CONTENT fixtures from the reference's own test images.

For a set of images of /root/reference/Test/Images/Small (every class the reference's image test harness runs:
LDR-RGB, LDR-RGBA, LDR-L, LDR-XY, LDRS-RGBA, HDR-RGB, HDR-RGBA - Test/astc_test_image.py:44-47, Test/testlib/encoder.py:269-333)
the fixture holds the decoded pixels and the blocks the UNMODIFIED reference produces for them at 4x4 -fast, 6x6 -medium and
8x8 -thorough with the harness's own switches (-cl / -cs / -ch / -cH, -normal for the XY maps).

The blocks come from the reference COMMAND LINE TOOL (tools/_build/astcenc-ref = Source/astcenccli_*.cpp + the reference
library, built by tools/relink_reference.sh), i.e. through the reference's own image loaders. The pixels are decoded here
(PIL for PNG, small parsers for DDS L8 / KTX, OpenCV for Radiance .hdr); they are accepted only if the reference LIBRARY fed
with them reproduces the tool's output byte for byte - so the stored pixels are, to the bit, what the tool compressed.

/root/reference does not exist on the GPU box: run this here, commit the .npz.
    python tests/golden/make_golden_images.py
"""
import os
import struct
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from astc_ref import *  # noqa: E402,F401,F403

ROOT = os.path.dirname(os.path.dirname(HERE))
SMALL = "/root/reference/Test/Images/Small"
CLI = os.path.join(ROOT, "tools", "_build", "astcenc-ref")

# (key, file, harness profile switch, flags beyond SELF_DECOMPRESS_ONLY, swizzle, extra CLI switches)
IMAGES = [
    ("ldr-rgb-00", "LDR-RGB/ldr-rgb-00.png", "-cl", 0, (0, 1, 2, 3), []),
    ("ldr-rgb-03", "LDR-RGB/ldr-rgb-03.png", "-cl", 0, (0, 1, 2, 3), []),
    ("ldr-rgb-05", "LDR-RGB/ldr-rgb-05.png", "-cl", 0, (0, 1, 2, 3), []),
    ("ldr-rgb-10", "LDR-RGB/ldr-rgb-10.png", "-cl", 0, (0, 1, 2, 3), []),          # 127 x 128: ragged edges
    ("ldr-rgba-00", "LDR-RGBA/ldr-rgba-00.png", "-cl", 0, (0, 1, 2, 3), []),
    ("ldr-rgba-02", "LDR-RGBA/ldr-rgba-02.png", "-cl", 0, (0, 1, 2, 3), []),
    ("ldr-l-00", "LDR-L/ldr-l-00-3.dds", "-cl", 0, (0, 1, 2, 3), []),
    ("ldr-l-01", "LDR-L/ldr-l-01-3.dds", "-cl", 0, (0, 1, 2, 3), []),
    ("ldr-xy-00", "LDR-XY/ldr-xy-00.png", "-cl", FLG_MAP_NORMAL, (0, 0, 0, 1), ["-normal"]),
    ("ldr-xy-02", "LDR-XY/ldr-xy-02.png", "-cl", FLG_MAP_NORMAL, (0, 0, 0, 1), ["-normal"]),
    ("ldrs-rgba-00", "LDRS-RGBA/ldrs-rgba-00.png", "-cs", 0, (0, 1, 2, 3), []),
    ("hdr-rgb-00", "HDR-RGB/hdr-rgb-00.hdr", "-ch", 0, (0, 1, 2, 3), []),
    ("hdr-rgb-rgb32", "HDR-RGB/hdr-rgb-rgb32.ktx", "-ch", 0, (0, 1, 2, 3), []),
    ("hdr-rgba-rgba16", "HDR-RGBA/hdr-rgba-rgba16.ktx", "-cH", 0, (0, 1, 2, 3), []),
    ("hdr-rgba-rgba32", "HDR-RGBA/hdr-rgba-rgba32.ktx", "-cH", 0, (0, 1, 2, 3), []),
]
CONFIGS = [("4x4", 4, "-fast", PRE_FAST), ("6x6", 6, "-medium", PRE_MEDIUM), ("8x8", 8, "-thorough", PRE_THOROUGH)]
PROFILE = {"-cl": PRF_LDR, "-cs": PRF_LDR_SRGB, "-ch": PRF_HDR_RGB_LDR_A, "-cH": PRF_HDR}


def load_pixels(path):
    if path.endswith(".png"):
        from PIL import Image
        return np.ascontiguousarray(np.asarray(Image.open(path).convert("RGBA"), dtype=np.uint8))
    if path.endswith(".dds"):
        b = open(path, "rb").read()
        assert b[:4] == b"DDS "
        h, w = struct.unpack_from("<II", b, 12)
        bits = struct.unpack_from("<I", b, 4 + 72 + 12)[0]
        assert bits == 8, "only the L8 files of the small set are handled"
        lum = np.frombuffer(b, dtype=np.uint8, count=h * w, offset=128).reshape(h, w)
        # (the volume files of the set hold `depth` slices: the harness compresses them as 2D with 2D block sizes -> slice 0..; the tool
        #  treats a 2D block size on a volume as an array of slices; the fixture keeps the first slice only when depth > 1)
        return np.ascontiguousarray(np.stack([lum, lum, lum, np.full_like(lum, 255)], axis=-1))
    if path.endswith(".hdr"):
        import cv2
        bgr = cv2.imread(path, cv2.IMREAD_UNCHANGED)
        rgb = bgr[..., ::-1].astype(np.float32)
        rgba = np.concatenate([rgb, np.ones(rgb.shape[:2] + (1,), np.float32)], axis=-1)
        return np.ascontiguousarray(rgba.astype(np.float16))
    if path.endswith(".ktx"):
        b = open(path, "rb").read()
        hdr = struct.unpack_from("<13I", b, 12)
        gl_type, _, gl_format = hdr[1], hdr[2], hdr[3]
        w, h, kv = hdr[6], hdr[7], hdr[12]
        comps = {6403: 1, 33319: 2, 6407: 3, 6408: 4}[gl_format]
        dt = {5131: np.float16, 5126: np.float32}[gl_type]
        off = 64 + kv + 4
        row = w * comps * np.dtype(dt).itemsize
        stride = (row + 3) & ~3
        out = np.zeros((h, w, 4), np.float32)
        out[..., 3] = 1.0
        for y in range(h):
            px = np.frombuffer(b, dtype=dt, count=w * comps, offset=off + y * stride).reshape(w, comps).astype(np.float32)
            out[y, :, :comps] = px
            if comps == 1:      # (the tool replicates a single channel into RGB; not used by the files picked above)
                out[y, :, 1] = px[:, 0]
                out[y, :, 2] = px[:, 0]
        return np.ascontiguousarray(out.astype(np.float16))
    raise ValueError(path)


def cli_blocks(path, switch, block, preset, extra):
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, "o.astc")
        r = subprocess.run([CLI, switch, path, out, block, preset, "-silent", "-j", "4"] + extra, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        b = open(out, "rb").read()
        return np.frombuffer(b, dtype=np.uint8)[16:].copy(), b[:16]


def main():
    assert os.path.exists(CLI), "run tools/relink_reference.sh first"
    ref = ref_lib()
    store = {}
    index = []
    for key, rel, switch, fl, swz, extra in IMAGES:
        path = os.path.join(SMALL, rel)
        px = load_pixels(path)
        if px.ndim == 3 and rel.endswith(".dds"):
            pass
        store["px_" + key] = px
        for bname, b, preset, q in CONFIGS:
            want, hdr = cli_blocks(path, switch, bname, preset, extra)
            dim_x = hdr[7] | (hdr[8] << 8) | (hdr[9] << 16)
            dim_y = hdr[10] | (hdr[11] << 8) | (hdr[12] << 16)
            dim_z = hdr[13] | (hdr[14] << 8) | (hdr[15] << 16)
            if dim_z > 1:
                # a volume: the tool compressed every slice; the fixture keeps slice 0 (2D path of this repo)
                per = len(want) // dim_z
                want = want[:per]
            assert (dim_x, dim_y) == (px.shape[1], px.shape[0]), (key, dim_x, dim_y, px.shape)
            got = np.frombuffer(bytes(ref.compress(px, PROFILE[switch], b, b, q, FLG_SELF_DECOMPRESS_ONLY | fl, swz=swz, threads=4)), dtype=np.uint8)
            d = block_diff(got, want)
            assert len(d) == 0, "%s %s: the library fed with the decoded pixels differs from the tool in %d blocks - pixel decoding is off" % (key, bname, len(d))
            store["blk_%s_%s" % (key, bname)] = want
        index.append("%s|%s|%d|%d|%s" % (key, switch, fl, PROFILE[switch], ",".join(str(s) for s in swz)))
        print(key, px.shape, px.dtype, "ok")
    store["index"] = np.array(index)
    np.savez_compressed(os.path.join(HERE, "golden_images.npz"), **store)
    print("wrote golden_images.npz", os.path.getsize(os.path.join(HERE, "golden_images.npz")), "bytes")


if __name__ == "__main__":
    main()
