"""The .astc container: astcenc_b200_store_cimage / astcenc_b200_load_cimage against the
reference's own fixture files (tests/golden/golden_container.npz = Test/Data/*.astc, see make_golden_container.py) and
the format description (Docs/FileFormat.md). Host-only code: runs without a GPU."""
import os
import numpy as np
import pytest
from astc_ref import ROOT

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_container.npz"))


def _write(tmp_path, name):
    p = tmp_path / name
    p.write_bytes(GOLD[name].tobytes())
    return str(p)


@pytest.mark.parametrize("name,block,dim,nblocks", [
    ("LDR-A-1x1.astc", (6, 6, 1), (1, 1, 1), 1), ("LDRS-A-1x1.astc", (6, 6, 1), (1, 1, 1), 1), ("HDR-A-1x1.astc", (6, 6, 1), (1, 1, 1), 1),
    ("ldr.astc", (4, 4, 1), (8, 8, 1), 4), ("hdr.astc", (4, 4, 1), (8, 8, 1), 4)])
def test_load_reference_fixtures(pkg, tmp_path, name, block, dim, nblocks):
    data, hdr = pkg.load_cimage(_write(tmp_path, name))
    assert (hdr["block_x"], hdr["block_y"], hdr["block_z"]) == block
    assert (hdr["dim_x"], hdr["dim_y"], hdr["dim_z"]) == dim
    assert data.nbytes == nblocks * 16
    assert data.tobytes() == GOLD[name].tobytes()[16:]
    # writing it back reproduces the reference's file byte for byte
    out = str(tmp_path / ("re_" + name))
    pkg.store_cimage(out, data, dim[0], dim[1], block[0], block[1])
    assert open(out, "rb").read() == GOLD[name].tobytes()


@pytest.mark.parametrize("name", ["negative_magic.astc", "negative_huge.astc", "negative_overflow.astc", "negative_short.astc"])
def test_corrupt_files_are_refused(pkg, tmp_path, name):
    """astc_test_functional.py:2195-2250: wrong magic, absurd size, size overflow, truncated payload."""
    with pytest.raises(pkg.AstcencError):
        pkg.load_cimage(_write(tmp_path, name))


def test_bad_block_size_is_refused_by_config_init(pkg, tmp_path):
    """negative_block_size.astc (5x6) loads - the container does not know which footprints exist - and the codec
    refuses it, as in the reference (astc_test_functional.py:2252-2262; astcenc_entry.cpp:497 -> BAD_BLOCK_SIZE)."""
    data, hdr = pkg.load_cimage(_write(tmp_path, "negative_block_size.astc"))
    assert (hdr["block_x"], hdr["block_y"]) == (5, 6)
    with pytest.raises(pkg.AstcencError) as e:
        pkg.config_init(pkg.PRF_LDR, hdr["block_x"], hdr["block_y"], 60.0)
    assert e.value.code == pkg.ERR_BAD_BLOCK_SIZE


def test_round_trip_and_header_layout(pkg, tmp_path):
    rng = np.random.default_rng(1)
    w, h, bx, by = 70001, 13, 12, 10      # a dimension that needs the third size byte
    n = ((w + bx - 1) // bx) * ((h + by - 1) // by)
    blocks = rng.integers(0, 256, n * 16, dtype=np.uint8)
    p = str(tmp_path / "big.astc")
    pkg.store_cimage(p, blocks, w, h, bx, by)
    raw = open(p, "rb").read()
    assert raw[:4] == bytes([0x13, 0xAB, 0xA1, 0x5C]) and raw[4:7] == bytes([bx, by, 1])
    assert raw[7:10] == bytes([w & 0xFF, (w >> 8) & 0xFF, w >> 16]) and raw[10:13] == bytes([h, 0, 0]) and raw[13:16] == bytes([1, 0, 0])
    data, hdr = pkg.load_cimage(p)
    assert np.array_equal(data, blocks) and hdr["dim_x"] == w and hdr["dim_y"] == h
    # payload size must match the header's block grid
    with pytest.raises(pkg.AstcencError):
        pkg.store_cimage(str(tmp_path / "bad.astc"), blocks[:-16], w, h, bx, by)
    with pytest.raises(pkg.AstcencError):
        pkg.store_cimage(str(tmp_path / "bad.astc"), blocks, 1 << 24, h, bx, by)
    with pytest.raises(pkg.AstcencError):
        pkg.load_cimage(str(tmp_path / "missing.astc"))


@pytest.mark.gpu
def test_fixture_blocks_decode_like_the_oracle(pkg, tmp_path):
    """The reference's ldr.astc tile through load_cimage -> astcenc_decompress_image equals the oracle's decode."""
    from astc_ref import Oracle, PRF_LDR
    data, hdr = pkg.load_cimage(_write(tmp_path, "ldr.astc"))
    cfg = pkg.config_init(pkg.PRF_LDR, hdr["block_x"], hdr["block_y"], 60.0, flags=pkg.FLG_DECOMPRESS_ONLY)
    ctx = pkg.Context(cfg)
    try:
        img = ctx.decompress_image(data, hdr["dim_x"], hdr["dim_y"])
    finally:
        ctx.close()
    want = Oracle().decompress(data, hdr["dim_x"], hdr["dim_y"], PRF_LDR, hdr["block_x"], hdr["block_y"])
    assert np.array_equal(np.asarray(img).reshape(-1), np.asarray(want).reshape(-1))
