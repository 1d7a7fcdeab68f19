"""CPU: the oracle port against the golden vectors produced by the unmodified reference build
(tests/golden/make_golden.py), and - where the reference build is present - directly against it."""
import numpy as np
import pytest

import astc_images as I
from astc_ref import *  # noqa: F401,F403
from golden.make_golden import CASES, make_image


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_oracle_matches_golden(case, oracle, golden):
    name, gen, size, dtype, prof, bx, by, q, fl, swz = case
    img = make_image(gen, size, dtype)
    got = oracle.compress(img, prof, bx, by, q, fl, swz=list(swz))
    want = golden[name]
    assert got.shape == want.shape
    assert len(block_diff(got, want)) == 0


def test_golden_fixture_is_reproducible(reference, golden):
    """The committed fixture is what the reference build emits (guards against a stale .npz)."""
    for case in CASES[:6]:
        name, gen, size, dtype, prof, bx, by, q, fl, swz = case
        img = make_image(gen, size, dtype)
        assert np.array_equal(reference.compress(img, prof, bx, by, q, fl, swz=swz), golden[name])


def test_oracle_matches_reference_on_larger_image(oracle, reference):
    img = I.photo_like(192, 160, seed=99)
    for (bx, by, q) in [(6, 6, PRE_MEDIUM), (4, 4, PRE_FAST)]:
        r = reference.compress(img, PRF_LDR, bx, by, q, FLG_SELF_DECOMPRESS_ONLY)
        o = oracle.compress(img, PRF_LDR, bx, by, q, FLG_SELF_DECOMPRESS_ONLY)
        assert len(block_diff(r, o)) == 0


def test_reference_thread_invariance(reference):
    """The reference's output does not depend on the number of caller threads."""
    img = I.photo_like(96, 96, seed=5)
    a = reference.compress(img, PRF_LDR, 6, 6, PRE_MEDIUM, FLG_SELF_DECOMPRESS_ONLY, threads=1)
    b = reference.compress(img, PRF_LDR, 6, 6, PRE_MEDIUM, FLG_SELF_DECOMPRESS_ONLY, threads=4)
    assert np.array_equal(a, b)


def test_constant_block_golden_vector(oracle):
    """Known-answer: a constant LDR block is FC FD FF FF FF FF FF FF + 4 x UNORM16 (the layout the
    reference's own functional test byte-compares, Test/astc_test_functional.py:457)."""
    img = I.constant(4, 4, rgba=(255, 128, 0, 255))
    out = oracle.compress(img, PRF_LDR, 4, 4, PRE_MEDIUM, FLG_SELF_DECOMPRESS_ONLY)
    assert bytes(out[:8]) == bytes([0xFC, 0xFD, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF])
    vals = out[8:16].view(np.uint16)
    assert list(vals) == [65535, 128 * 257, 0, 65535]
