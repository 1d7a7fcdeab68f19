"""CPU: host-side logic of the product (config presets, table construction, C-ABI surface, sharding maths)
and the host-simulated device source against the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import astc_images as I
from astc_ref import *  # noqa: F401,F403
from golden.make_golden import CASES, make_image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PRESET_POINTS = [0.0, 5.0, 10.0, 35.0, 60.0, 80.0, 98.0, 98.5, 99.0, 99.5, 100.0]
BLOCKS = [(4, 4), (5, 4), (5, 5), (6, 5), (6, 6), (8, 5), (8, 6), (8, 8), (10, 5), (10, 6), (10, 8), (10, 10), (12, 10), (12, 12)]
FIELDS = [f for f, _ in Config._fields_ if f not in ("progress_callback",)]


@pytest.mark.parametrize("profile", [PRF_LDR_SRGB, PRF_LDR, PRF_HDR_RGB_LDR_A, PRF_HDR])
def test_config_init_matches_reference(pkg, reference, profile):
    """astcenc_config_init: every field equal to the reference for every block size, preset and interpolated quality."""
    for (bx, by) in BLOCKS:
        for q in PRESET_POINTS:
            for fl in (0, FLG_MAP_NORMAL, FLG_MAP_RGBM, FLG_USE_PERCEPTUAL | FLG_SELF_DECOMPRESS_ONLY):
                want = reference.config(profile, bx, by, q, fl)
                got = pkg.config_init(profile, bx, by, q, fl)
                for f in FIELDS:
                    assert getattr(got, f) == getattr(want, f), (bx, by, q, fl, f)


def test_config_init_errors(pkg):
    lib = pkg.lib()
    cfg = pkg.Config()
    assert lib.astcenc_config_init(PRF_LDR, 7, 7, 1, 60.0, 0, C.byref(cfg)) == 4      # BAD_BLOCK_SIZE
    assert lib.astcenc_config_init(PRF_LDR, 6, 6, 1, 101.0, 0, C.byref(cfg)) == 6     # BAD_QUALITY
    assert lib.astcenc_config_init(PRF_LDR, 6, 6, 1, -1.0, 0, C.byref(cfg)) == 6
    assert lib.astcenc_config_init(7, 6, 6, 1, 60.0, 0, C.byref(cfg)) == 5            # BAD_PROFILE
    assert lib.astcenc_config_init(PRF_LDR, 6, 6, 1, 60.0, 1 << 9, C.byref(cfg)) == 8  # BAD_FLAGS
    assert lib.astcenc_config_init(PRF_LDR, 6, 6, 1, 60.0, FLG_MAP_NORMAL | FLG_MAP_RGBM, C.byref(cfg)) == 8
    assert lib.astcenc_config_init(PRF_HDR, 6, 6, 1, 60.0, FLG_USE_DECODE_UNORM8, C.byref(cfg)) == 11  # BAD_DECODE_MODE
    assert lib.astcenc_config_init(PRF_LDR, 4, 4, 4, 60.0, 0, C.byref(cfg)) == 0      # 3D block sizes: the ten footprints
    assert lib.astcenc_config_init(PRF_LDR, 4, 4, 5, 60.0, 0, C.byref(cfg)) == 4      # ... and nothing else (BAD_BLOCK_SIZE)
    assert lib.astcenc_get_error_string(3) == b"ASTCENC_ERR_BAD_PARAM"
    assert lib.astcenc_get_error_string(99) is None


def test_library_exports_the_declared_abi(pkg):
    """Every function include/astcenc.h declares is exported by the shared library."""
    hdr = open(os.path.join(ROOT, "include", "astcenc.h")).read()
    names = set(re.findall(r"\b(astcenc_(?:b200_)?[a-z_0-9]+)\s*\(", hdr))
    names -= {"astcenc_progress_callback"}
    assert {"astcenc_config_init", "astcenc_context_alloc", "astcenc_compress_image", "astcenc_compress_reset", "astcenc_compress_cancel",
            "astcenc_decompress_image", "astcenc_decompress_reset", "astcenc_context_free", "astcenc_get_block_info",
            "astcenc_get_error_string"} <= names
    lib = pkg.lib()
    for n in names:
        assert hasattr(lib, n), n


def test_no_cpu_fallback_without_gpu(pkg):
    """Without a CUDA device a compression context cannot be created: the product never compresses on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    cfg = pkg.config_init(PRF_LDR, 6, 6, PRE_MEDIUM)
    with pytest.raises(pkg.AstcencError) as e:
        pkg.Context(cfg)
    assert e.value.code == 9   # ASTCENC_ERR_BAD_CONTEXT
    lib = pkg.lib()
    ctx = C.c_void_p()
    assert lib.astcenc_context_alloc(C.byref(cfg), 0, C.byref(ctx), None) == 3   # thread_count 0 -> BAD_PARAM (checked first)
    assert lib.astcenc_context_alloc(None, 1, C.byref(ctx), None) == 3            # neither config nor parent


def test_product_sources_do_not_reference_the_oracle():
    """The shipped path must not import, include or link anything under oracle/."""
    pk = os.path.join(ROOT, "astc-encoder_b200")
    for base, _, files in os.walk(pk):
        for f in files:
            if f.endswith((".cu", ".cuh", ".h", ".cpp", ".py", ".inc")) or f == "Makefile":
                txt = open(os.path.join(base, f), errors="ignore").read()
                assert "oracle/" not in txt and "astc_oracle" not in txt, os.path.join(base, f)


def test_table_counts_match_reference_probe(hostsim):
    """Arena plan is sane and the per-config table counts equal the numbers measured on the reference
    (block modes / decimation grids / partitionings per config)."""
    S = FLG_SELF_DECOMPRESS_ONLY
    assert 0 < hostsim.hostsim_arena_bytes(PRF_LDR, 6, 6, PRE_MEDIUM, S) <= 14 * 1024   # 16 warps/SM fit in shared memory
    assert 0 < hostsim.hostsim_arena_bytes(PRF_LDR, 4, 4, PRE_FAST, S) <= 8 * 1024
    assert hostsim.hostsim_arena_bytes(PRF_LDR, 12, 12, PRE_EXHAUSTIVE, 0) > 0


SIM_CASES = [c for c in CASES if c[2][0] * c[2][1] <= 64 * 64]


@pytest.mark.parametrize("case", SIM_CASES, ids=[c[0] for c in SIM_CASES])
def test_device_source_host_simulation_matches_golden(case, hostsim, golden):
    """The kernel source compiled for the host with one simulated lane reproduces the reference's blocks."""
    name, gen, size, dtype, prof, bx, by, q, fl, swz = case
    img = np.ascontiguousarray(make_image(gen, size, dtype))
    h, w = img.shape[:2]
    dt = {np.dtype(np.uint8): 0, np.dtype(np.float16): 1, np.dtype(np.float32): 2}[img.dtype]
    out = np.zeros(((w + bx - 1) // bx) * ((h + by - 1) // by) * 16, np.uint8)
    sw = (C.c_int * 4)(*swz)
    assert hostsim.hostsim_compress_image(prof, bx, by, q, fl, img.ctypes.data, dt, w, h, sw, out.ctypes.data) == 0
    assert len(block_diff(out, golden[name])) == 0


def test_slab_rows_partition(pkg):
    for blocks_y in (1, 2, 7, 683, 1000):
        for world in (1, 2, 3, 4, 8):
            rows = [pkg.slab_rows(blocks_y, r, world) for r in range(world)]
            assert rows[0][0] == 0 and rows[-1][1] == blocks_y
            for a, b in zip(rows, rows[1:]):
                assert a[1] == b[0]
