"""Alpha-scale pre-pass (second "next" row): astcenc_config::a_scale_radius.

Fixtures: tests/golden/golden_alpha.npz, written by the unmodified reference build (tests/golden/make_golden_alpha.py).
CPU: oracle restatement and host simulation of the device source; GPU: the CUDA path through the C ABI."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))
from astc_ref import *  # noqa: E402,F401,F403
from make_golden_alpha import ACASES, make_alpha_image  # noqa: E402

DT = {"u8": TYPE_U8, "f16": TYPE_F16, "f32": TYPE_F32}


@pytest.fixture(scope="module")
def golden_alpha():
    return np.load(os.path.join(HERE, "golden", "golden_alpha.npz"))


@pytest.mark.parametrize("i", range(len(ACASES)), ids=[c[0] for c in ACASES])
def test_oracle_alpha_scale_matches_reference_fixture(i, oracle, golden_alpha):
    name, size, dt, prof, bx, by, q, fl, swz, r = ACASES[i]
    got = oracle.compress(make_alpha_image(size, dt, 40 + i), prof, bx, by, q, fl, swz=list(swz), a_scale_radius=r)
    assert len(block_diff(got, golden_alpha[name])) == 0


@pytest.mark.parametrize("i", range(len(ACASES)), ids=[c[0] for c in ACASES])
def test_device_source_host_simulation_alpha_scale(i, hostsim, golden_alpha):
    name, size, dt, prof, bx, by, q, fl, swz, r = ACASES[i]
    img = np.ascontiguousarray(make_alpha_image(size, dt, 40 + i))
    out = np.zeros(golden_alpha[name].size, dtype=np.uint8)
    sw = (C.c_int * 4)(*swz)
    hostsim.hostsim_set_a_scale_radius(r)
    try:
        assert hostsim.hostsim_compress_image(prof, bx, by, q, fl, img.ctypes.data, DT[dt], size[1], size[0], sw, out.ctypes.data) == 0
    finally:
        hostsim.hostsim_set_a_scale_radius(0)
    assert len(block_diff(out, golden_alpha[name])) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(len(ACASES)), ids=[c[0] for c in ACASES])
def test_cuda_alpha_scale_matches_reference_fixture(i, pkg, golden_alpha):
    name, size, dt, prof, bx, by, q, fl, swz, r = ACASES[i]
    cfg = pkg.config_init(prof, bx, by, q, fl, a_scale_radius=r)
    ctx = pkg.Context(cfg)
    try:
        got = ctx.compress_image(make_alpha_image(size, dt, 40 + i), swizzle=swz)
    finally:
        ctx.close()
    assert len(block_diff(got, golden_alpha[name])) == 0


@pytest.mark.gpu
def test_cuda_alpha_scale_large_image_against_oracle(pkg, oracle):
    """Many 32x32 tiles, slabs: 300x260 image in two block-row slabs equals the oracle's whole-image result."""
    import torch
    img = make_alpha_image((260, 300), "u8", 77)
    want = oracle.compress(img, PRF_LDR, 6, 6, PRE_FAST, FLG_USE_ALPHA_WEIGHT | FLG_SELF_DECOMPRESS_ONLY, a_scale_radius=2)
    cfg = pkg.config_init(PRF_LDR, 6, 6, PRE_FAST, FLG_USE_ALPHA_WEIGHT | FLG_SELF_DECOMPRESS_ONLY, a_scale_radius=2)
    ctx = pkg.Context(cfg)
    try:
        assert len(block_diff(ctx.compress_image(img), want)) == 0
        nbx, nby = ctx.blocks(300, 260)
        d_img = torch.from_numpy(img).cuda()
        d_out = torch.zeros(nbx * nby * 16, dtype=torch.uint8, device="cuda")
        half = nby // 2
        ctx.compress_device(d_img.data_ptr(), pkg.TYPE_U8, 300, 260, d_out.data_ptr(), block_row0=0, block_rows=half)
        ctx.compress_device(d_img.data_ptr(), pkg.TYPE_U8, 300, 260, d_out.data_ptr() + half * nbx * 16, block_row0=half, block_rows=nby - half)
        torch.cuda.synchronize()
        assert len(block_diff(d_out.cpu().numpy(), want)) == 0
    finally:
        ctx.close()
