"""Generates tests/golden/golden_metrics.npz: image pairs and the figures the UNMODIFIED reference
compute_error_metrics() (astcenccli_error_metrics.cpp:109-413) prints for them, via oracle/_ref/libastcenc_ref_metrics.so
(oracle/Makefile builds it from /root/reference plus oracle/ref_metrics_shim.cpp). Run in the build container:
    make -C oracle && python tests/golden/make_golden_metrics.py
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from astc_ref import ref_error_metrics, METRIC_FIELDS


def cases():
    rng = np.random.default_rng(20260923)
    a = rng.integers(0, 256, (45, 67, 4), dtype=np.uint8)
    b = np.clip(a.astype(int) + rng.integers(-9, 10, a.shape), 0, 255).astype(np.uint8)
    yield "ldr_rgba", a, b, dict(components=4)
    yield "ldr_rgb", a, b, dict(components=3)
    yield "ldr_la", a, b, dict(components=2)
    yield "ldr_l", a, b, dict(components=1)
    yield "ldr_identical", a, a.copy(), dict(components=4)
    yield "ldr_normal", a, b, dict(components=3, normal=True)
    yield "ldr_cropped", a, b[:40, :60].copy(), dict(components=4)
    h1 = np.exp2(rng.uniform(-8, 8, (33, 41, 4))).astype(np.float16)
    h2 = (h1.astype(np.float32) * rng.uniform(0.95, 1.05, h1.shape)).astype(np.float16)
    yield "hdr_f16", h1, h2, dict(components=3, hdr=True, fstop_lo=-10, fstop_hi=10)
    yield "hdr_f16_rgba", h1, h2, dict(components=4, hdr=True, fstop_lo=-5, fstop_hi=5)
    f1 = np.exp2(rng.uniform(-6, 6, (29, 35, 4))).astype(np.float32)
    f2 = (f1 * rng.uniform(0.9, 1.1, f1.shape)).astype(np.float32)
    yield "hdr_f32", f1, f2, dict(components=4, hdr=True, fstop_lo=-3, fstop_hi=8)
    yield "mixed_u8_f32", a, (b.astype(np.float32) / 255.0), dict(components=4)
    z = np.zeros((16, 16, 4), dtype=np.uint8)
    z[..., :3] = 128
    yield "normal_degenerate", z, a[:16, :16].copy(), dict(components=3, normal=True)
    zf = np.full((12, 20, 4), 0.5, dtype=np.float32)      # zero-length normals: normalize_safe() falls back to unit3()
    yield "normal_zero_f32", zf, (a[:12, :20].astype(np.float32) / 255.0), dict(components=3, normal=True)


if __name__ == "__main__":
    out = {}
    names = []
    for name, i1, i2, kw in cases():
        r = ref_error_metrics(i1, i2, **kw)
        names.append(name)
        out[name + "/img1"] = i1
        out[name + "/img2"] = i2
        out[name + "/args"] = np.array([int(kw.get("hdr", False)), int(kw.get("normal", False)), kw.get("components", 4),
                                        kw.get("fstop_lo", -10), kw.get("fstop_hi", 10)], dtype=np.int32)
        out[name + "/ref"] = np.array([r.get(k, np.nan) for k in METRIC_FIELDS], dtype=np.float64)
        print(name, r)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_metrics.npz"), **out)
