"""The device source on 32 emulated lanes (tests/hostsim/simt_emul.h): one host thread per lane of a warp, the CUDA warp
primitives as collectives over the threads. Where the one-lane simulation checks the kernels' logic, this build runs the
lane-parallel protocol itself without a GPU: shuffle partners, ordered sums spread over lanes, warp-uniform decisions,
__syncwarp between producer and consumer lanes. Results must equal the oracle's byte for byte, and ThreadSanitizer - for
which the collectives are the only synchronisation between lanes - must stay silent. Small images only (~0.2 s per block)."""
import ctypes as C
import os
import subprocess
import numpy as np
import pytest
from astc_ref import *
import astc_images as I

S = FLG_SELF_DECOMPRESS_ONLY
_DT = {np.dtype(np.uint8): TYPE_U8, np.dtype(np.float16): TYPE_F16, np.dtype(np.float32): TYPE_F32}


def sim_compress(lib, img, prof, bx, by, q, flags=S, swz=None):
    img = np.ascontiguousarray(img)
    h, w = img.shape[:2]
    out = np.zeros(((w + bx - 1) // bx) * ((h + by - 1) // by) * 16, np.uint8)
    sw = (C.c_int * 4)(*swz) if swz is not None else None
    rc = lib.hostsim_compress_image(prof, bx, by, q, flags, img.ctypes.data, _DT[img.dtype], w, h, sw, out.ctypes.data)
    assert rc == 0
    return out


CASES = [
    ("photo 6x6 medium", lambda: I.photo_like(18, 24, seed=4), PRF_LDR, 6, 6, PRE_MEDIUM, S),
    ("noise 6x6 medium", lambda: I.uniform_noise(12, 12, seed=7), PRF_LDR, 6, 6, PRE_MEDIUM, S),
    ("voronoi 4x4 thorough", lambda: I.voronoi_flat(12, 16, seed=5), PRF_LDR, 4, 4, PRE_THOROUGH, S),
    ("voronoi 8x8 medium", lambda: I.voronoi_flat(16, 16, cell=5, seed=9), PRF_LDR, 8, 8, PRE_MEDIUM, S),
    ("hdr 6x5 medium", lambda: I.hdr_noise(10, 12, seed=6), PRF_HDR, 6, 5, PRE_MEDIUM, S),
    ("alpha 5x5 fast", lambda: I.alpha_mask(15, 15, seed=8), PRF_LDR, 5, 5, PRE_FAST, S | FLG_USE_ALPHA_WEIGHT),
    ("normal 6x6 medium", lambda: I.photo_like(12, 12, seed=10), PRF_LDR, 6, 6, PRE_MEDIUM, S | FLG_MAP_NORMAL),
    ("photo 12x12 medium", lambda: I.photo_like(24, 24, seed=11), PRF_LDR, 12, 12, PRE_MEDIUM, S),
]


@pytest.mark.parametrize("name,gen,prof,bx,by,q,fl", CASES, ids=[c[0] for c in CASES])
def test_wave_pipeline_on_32_lanes_matches_oracle(hostsim32, name, gen, prof, bx, by, q, fl):
    img = gen()
    want = Oracle().compress(img, prof, bx, by, q, fl)
    assert np.array_equal(sim_compress(hostsim32, img, prof, bx, by, q, fl), want)


@pytest.mark.parametrize("bx,by,q", [(6, 6, PRE_MEDIUM), (8, 8, PRE_MEDIUM), (10, 5, PRE_MEDIUM), (12, 10, PRE_MEDIUM)], ids=["6x6", "8x8", "10x5", "12x10"])
def test_wavefront_replay_of_realign_weights(hostsim32, bx, by, q):
    """Noisy blocks move most of their weights: realign_weights must take its wavefront replay (fronts k = x + 2y) - asserted through
    the simulation's counter - and still decide every weight on the state the reference's index-order loop shows it."""
    img = I.uniform_noise(2 * by, 3 * bx, seed=21 + bx)
    hostsim32.hostsim_wavefront_replays.restype = C.c_uint
    hostsim32.hostsim_wavefront_replays()
    got = sim_compress(hostsim32, img, PRF_LDR, bx, by, q)
    assert hostsim32.hostsim_wavefront_replays() > 0
    assert np.array_equal(got, Oracle().compress(img, PRF_LDR, bx, by, q, S))


@pytest.mark.parametrize("driver", ["lockstep", "warp"])
def test_single_kernel_drivers_on_32_lanes(hostsim32, monkeypatch, driver):
    monkeypatch.setenv("HOSTSIM_DRIVER", driver)
    img = I.photo_like(12, 18, seed=12)
    want = Oracle().compress(img, PRF_LDR, 6, 6, PRE_MEDIUM, S)
    assert np.array_equal(sim_compress(hostsim32, img, PRF_LDR, 6, 6, PRE_MEDIUM), want)


def test_alpha_scale_rdo_on_32_lanes(hostsim32):
    img = I.alpha_mask(24, 24, seed=13)
    want = Oracle().compress(img, PRF_LDR, 6, 6, PRE_MEDIUM, S | FLG_USE_ALPHA_WEIGHT, a_scale_radius=2)
    hostsim32.hostsim_set_a_scale_radius(2)
    try:
        got = sim_compress(hostsim32, img, PRF_LDR, 6, 6, PRE_MEDIUM, S | FLG_USE_ALPHA_WEIGHT)
    finally:
        hostsim32.hostsim_set_a_scale_radius(0)
    assert np.array_equal(got, want)


def test_decode_on_32_lanes(hostsim32):
    orc = Oracle()
    for img, prof, bx, by, ot in ((I.photo_like(20, 26, seed=14), PRF_LDR, 6, 6, TYPE_U8), (I.hdr_noise(16, 16, seed=15), PRF_HDR, 8, 8, TYPE_F16)):
        blocks = orc.compress(img, prof, bx, by, PRE_FAST, 0)
        h, w = img.shape[:2]
        want = orc.decompress(blocks, w, h, prof, bx, by, out_type=ot)
        got = np.zeros_like(want)
        rc = hostsim32.hostsim_decompress_image(prof, bx, by, 0, blocks.ctypes.data, got.ctypes.data, ot, w, h, None)
        assert rc == 0
        assert np.array_equal(got.view(np.uint8), want.view(np.uint8))


@pytest.mark.parametrize("args", ["1 6 6 60 18 12 1 0", "1 6 6 60 24 18 2 1", "3 6 6 60 12 12 3 2", "1 4 4 98 12 12 4 1", "1 8 8 60 16 16 5 1",
                                  "1 6 6 10 18 18 6 3", "1 6 6 60 18 18 7 4",      # these two: + decode of the result, alpha-scale pre-pass
                                  "1 4 4 60 8 32 8 0 4 4", "1 5 5 60 10 30 9 1 4 6"])      # 3D block sizes 4x4x4 / 5x5x4 on 8x8x4 / 10x5x6 volumes, + decode
def test_no_data_race_between_lanes(lanes32_tsan, args):
    """ThreadSanitizer over the whole wave pipeline: any report is a missing __syncwarp() between lanes (removing one is
    detected at once - tried by hand on pack_work_endpoints)."""
    if lanes32_tsan is None:
        pytest.skip("g++ -fsanitize=thread not available")
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=1 exitcode=66")
    r = subprocess.run([lanes32_tsan] + args.split(), capture_output=True, text=True, env=env, timeout=600)
    assert "ThreadSanitizer" not in r.stderr, r.stderr[-3000:]
    assert r.returncode == 0 and r.stdout.startswith("rc 0"), (r.returncode, r.stdout, r.stderr[-1000:])
