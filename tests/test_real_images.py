"""Real content: the reference's own test images (Test/Images/Small, every class its image harness runs) and the blocks
the reference COMMAND LINE TOOL produced for them (tests/golden/golden_images.npz, made by tests/golden/make_golden_images.py),
plus the reference's tools RELINKED against this library (tools/relink_reference.sh -> tools/_build):

CPU : the oracle port on the fixtures; the relinked binaries exist and import nothing but what libastcenc_b200.so exports
GPU : the CUDA path on every fixture (15 images x 3 configurations, all profiles); `astcenc-b200 -cl/-cs/-ch/-cH` vs
      `astcenc-ref` byte for byte on files; the reference's UnitTest/test_encode.cpp + test_decode.cpp on this library
"""
import os
import subprocess

import numpy as np
import pytest

from astc_ref import *  # noqa: F401,F403

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
BUILD = os.path.join(ROOT, "tools", "_build")
S = FLG_SELF_DECOMPRESS_ONLY
CONFIGS = [("4x4", 4, PRE_FAST, "-fast"), ("6x6", 6, PRE_MEDIUM, "-medium"), ("8x8", 8, PRE_THOROUGH, "-thorough")]


@pytest.fixture(scope="module")
def images():
    z = np.load(os.path.join(HERE, "golden", "golden_images.npz"))
    out = []
    for row in z["index"]:
        key, switch, fl, prof, swz = str(row).split("|")
        out.append(dict(key=key, switch=switch, flags=int(fl), profile=int(prof), swz=tuple(int(s) for s in swz.split(",")), px=z["px_" + key],
                        blocks={b: z["blk_%s_%s" % (key, b)] for b, _, _, _ in CONFIGS}))
    return out


def _ensure_relinked():
    if not os.path.exists(os.path.join(BUILD, "astcenc-b200")):
        if not os.path.isdir("/root/reference/Source"):
            pytest.skip("relinked reference tools not built and no reference sources here")
        r = subprocess.run(["bash", os.path.join(ROOT, "tools", "relink_reference.sh")], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_fixture_covers_the_harness_classes(images):
    keys = [i["key"] for i in images]
    for cls in ("ldr-rgb-", "ldr-rgba-", "ldr-l-", "ldr-xy-", "ldrs-rgba-", "hdr-rgb-", "hdr-rgba-"):
        assert any(k.startswith(cls) for k in keys), cls
    assert {i["profile"] for i in images} == {PRF_LDR, PRF_LDR_SRGB, PRF_HDR_RGB_LDR_A, PRF_HDR}


def test_oracle_on_real_images(images, oracle):
    """The CPU restatement on real content: every image, every configuration (~20 s)."""
    for im in images:
        for bname, b, q, _ in CONFIGS:
            got = oracle.compress(im["px"], im["profile"], b, b, q, S | im["flags"], swz=im["swz"])
            d = block_diff(got, im["blocks"][bname])
            assert len(d) == 0, (im["key"], bname, len(d))


def test_relinked_tools_link_against_this_library(pkg):
    """The reference CLI and the reference unit tests, compiled against include/astcenc.h and linked against
    libastcenc_b200.so: every astcenc_* symbol they import is one this library exports."""
    _ensure_relinked()
    lib_syms = subprocess.run(["nm", "-D", "--defined-only", pkg.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in lib_syms.splitlines() if l.strip()}
    for exe in ("astcenc-b200", "unittests-b200"):
        path = os.path.join(BUILD, exe)
        assert os.path.exists(path)
        und = subprocess.run(["nm", "-D", "--undefined-only", path], capture_output=True, text=True).stdout
        need = {l.split()[-1] for l in und.splitlines() if " astcenc_" in l}
        assert need and need <= exported, need - exported
        ldd = subprocess.run(["ldd", path], capture_output=True, text=True).stdout
        assert "libastcenc_b200.so" in ldd and "not found" not in ldd.split("libastcenc_b200.so")[1].split("\n")[0]


@pytest.mark.gpu
def test_cuda_on_real_images(images, pkg):
    """Block-for-block equality with the reference tool on the reference's Test/Images content: every image class,
    4x4 -fast / 6x6 -medium / 8x8 -thorough, LDR / sRGB / HDR / HDR+LDR-alpha profiles, normal-map switch."""
    for bname, b, q, _ in CONFIGS:
        ctxs = {}
        for im in images:
            k = (im["profile"], im["flags"])
            if k not in ctxs:
                ctxs[k] = pkg.Context(pkg.config_init(im["profile"], b, b, q, S | im["flags"]))
            got = ctxs[k].compress_image(im["px"], swizzle=im["swz"])
            d = block_diff(got, im["blocks"][bname])
            assert len(d) == 0, "%s %s: %d of %d blocks differ from the reference tool, first %s" % (im["key"], bname, len(d), len(got) // 16, d[:5])
        for c in ctxs.values():
            c.close()


@pytest.mark.gpu
def test_relinked_cli_matches_reference_cli(images, tmp_path):
    """The reference's command line tool linked against this library (astcenc-b200) writes the same .astc files as the
    same tool linked against the reference library (astcenc-ref)."""
    _ensure_relinked()
    from PIL import Image
    ref_cli = os.path.join(BUILD, "astcenc-ref")
    for im in images:
        if im["px"].dtype != np.uint8 or im["key"] not in ("ldr-rgb-00", "ldr-rgba-02", "ldr-xy-00", "ldrs-rgba-00", "ldr-rgb-10"):
            continue
        src = os.path.join(str(tmp_path), im["key"] + ".png")
        Image.fromarray(im["px"], "RGBA").save(src)
        extra = ["-normal"] if im["flags"] & FLG_MAP_NORMAL else []
        for bname, b, q, preset in CONFIGS[:2]:
            out_b = os.path.join(str(tmp_path), "b200.astc")
            r = subprocess.run([os.path.join(BUILD, "astcenc-b200"), im["switch"], src, out_b, bname, preset, "-silent"] + extra, capture_output=True, text=True)
            assert r.returncode == 0, r.stdout + r.stderr
            data_b = open(out_b, "rb").read()
            assert np.array_equal(np.frombuffer(data_b[16:], np.uint8), im["blocks"][bname]), (im["key"], bname)
            if os.path.exists(ref_cli):
                out_r = os.path.join(str(tmp_path), "ref.astc")
                r = subprocess.run([ref_cli, im["switch"], src, out_r, bname, preset, "-silent", "-j", "4"] + extra, capture_output=True, text=True)
                assert r.returncode == 0, r.stdout + r.stderr
                assert data_b == open(out_r, "rb").read(), (im["key"], bname)
    # and the round trip the tool itself offers: -tl prints the PSNR after decompressing on the device
    src = os.path.join(str(tmp_path), "ldr-rgb-00.png")
    r = subprocess.run([os.path.join(BUILD, "astcenc-b200"), "-tl", src, os.path.join(str(tmp_path), "rt.png"), "6x6", "-medium"], capture_output=True, text=True)
    assert r.returncode == 0 and "PSNR" in r.stdout, r.stdout + r.stderr
    if os.path.exists(ref_cli):
        r2 = subprocess.run([ref_cli, "-tl", src, os.path.join(str(tmp_path), "rt2.png"), "6x6", "-medium", "-j", "4"], capture_output=True, text=True)
        psnr = [l for l in r.stdout.splitlines() if "PSNR" in l]
        psnr2 = [l for l in r2.stdout.splitlines() if "PSNR" in l]
        assert psnr == psnr2, (psnr, psnr2)


@pytest.mark.gpu
def test_reference_unit_tests_pass_on_this_library():
    """Source/UnitTest/test_encode.cpp and test_decode.cpp (GoogleTest), built against include/astcenc.h + libastcenc_b200.so."""
    _ensure_relinked()
    r = subprocess.run([os.path.join(BUILD, "unittests-b200")], capture_output=True, text=True, timeout=600)
    tail = r.stdout[-3000:]
    assert r.returncode == 0, tail + r.stderr[-1000:]
    assert "[  PASSED  ] 18 tests" in r.stdout, tail
