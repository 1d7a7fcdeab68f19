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


# ---- KTX 1 (astcenccli_image_load_store.cpp:870-905 header, :1294-1440 load / store of compressed images) ----
KTX_MAGIC = bytes([0xAB, 0x4B, 0x54, 0x58, 0x20, 0x31, 0x31, 0xBB, 0x0D, 0x0A, 0x1A, 0x0A])


def _u32(raw, i):
    return int.from_bytes(raw[i:i + 4], "little")


@pytest.mark.parametrize("bx,by,srgb,glfmt", [(4, 4, False, 0x93B0), (6, 6, False, 0x93B4), (12, 12, False, 0x93BD), (6, 5, True, 0x93D3), (10, 8, True, 0x93DA)])
def test_ktx_round_trip_and_header_fields(pkg, tmp_path, bx, by, srgb, glfmt):
    """Header fields as store_ktx_compressed_image writes them (:1404-1418) and the GL enums of the footprint table (:725-753)."""
    rng = np.random.default_rng(bx * 16 + by)
    w, h = 37, 29
    n = ((w + bx - 1) // bx) * ((h + by - 1) // by)
    blocks = rng.integers(0, 256, n * 16, dtype=np.uint8)
    p = str(tmp_path / "t.ktx")
    pkg.store_ktx_cimage(p, blocks, w, h, bx, by, is_srgb=srgb)
    raw = open(p, "rb").read()
    assert raw[:12] == KTX_MAGIC and len(raw) == 64 + 4 + n * 16
    fields = [_u32(raw, 12 + 4 * i) for i in range(13)]
    # endianness, glType, glTypeSize, glFormat, glInternalFormat, glBaseInternalFormat (GL_RGBA), width, height, depth, array elements, faces, mip levels, kv bytes
    assert fields == [0x04030201, 0, 1, 0, glfmt, 0x1908, w, h, 0, 0, 1, 1, 0]
    assert _u32(raw, 64) == n * 16 and raw[68:] == blocks.tobytes()
    data, hdr, is_srgb = pkg.load_ktx_cimage(p)
    assert np.array_equal(data, blocks) and is_srgb == srgb
    assert (hdr["block_x"], hdr["block_y"], hdr["block_z"], hdr["dim_x"], hdr["dim_y"], hdr["dim_z"]) == (bx, by, 1, w, h, 1)


def test_ktx_other_byte_order_and_key_value_data(pkg, tmp_path):
    """A file written on a big-endian machine (every 32-bit field reversed, :1321-1326) with key/value data to skip (:1343)."""
    blocks = np.arange(32, dtype=np.uint8)
    fields = [0x04030201, 0, 1, 0, 0x93B4, 0x1908, 7, 6, 0, 0, 1, 1, 8]
    raw = KTX_MAGIC + b"".join(v.to_bytes(4, "big") for v in fields) + b"KEYVALUE" + (32).to_bytes(4, "big") + blocks.tobytes()
    p = tmp_path / "be.ktx"
    p.write_bytes(raw)
    data, hdr, is_srgb = pkg.load_ktx_cimage(str(p))
    assert np.array_equal(data, blocks) and not is_srgb
    assert (hdr["block_x"], hdr["block_y"], hdr["dim_x"], hdr["dim_y"]) == (6, 6, 7, 6)


def test_ktx_refusals(pkg, tmp_path):
    good = [0x04030201, 0, 1, 0, 0x93B4, 0x1908, 6, 6, 0, 0, 1, 1, 0]

    def write(name, fields=good, magic=KTX_MAGIC, payload=bytes(16), length=16):
        p = tmp_path / name
        p.write_bytes(magic + b"".join(v.to_bytes(4, "little") for v in fields) + length.to_bytes(4, "little") + payload)
        return str(p)
    pkg.load_ktx_cimage(write("ok.ktx"))
    for name, kw in (("magic.ktx", dict(magic=b"\x00" + KTX_MAGIC[1:])), ("endian.ktx", dict(fields=[0x11223344] + good[1:])),
                     ("uncompressed.ktx", dict(fields=good[:1] + [0x1401] + good[2:])), ("format.ktx", dict(fields=good[:4] + [0x8058] + good[5:])),
                     ("base.ktx", dict(fields=good[:5] + [0x1907] + good[6:])), ("short.ktx", dict(payload=bytes(8)))):
        with pytest.raises(pkg.AstcencError):
            pkg.load_ktx_cimage(write(name, **kw))
    # a level size that is not the block grid of the header: trailing garbage / a short first level
    for name, kw in (("long.ktx", dict(payload=bytes(32), length=32)), ("dims.ktx", dict(fields=good[:6] + [13, 6] + good[8:]))):
        with pytest.raises(pkg.AstcencError):
            pkg.load_ktx_cimage(write(name, **kw))
    with pytest.raises(pkg.AstcencError) as e:
        pkg.store_ktx_cimage(str(tmp_path / "x.ktx"), bytes(16), 7, 7, 7, 7)      # no GL enum for a 7x7 footprint
    assert e.value.code == pkg.ERR_BAD_BLOCK_SIZE
    # store checks the payload against the header's block grid like store_cimage does, and refuses volumes with a 2D footprint
    with pytest.raises(pkg.AstcencError) as e:
        pkg.store_ktx_cimage(str(tmp_path / "y.ktx"), bytes(32), 6, 6, 6, 6)
    assert e.value.code == pkg.ERR_BAD_PARAM
    with pytest.raises(pkg.AstcencError):
        pkg.store_ktx_cimage(str(tmp_path / "z.ktx"), bytes(32), 6, 6, 6, 6, dim_z=2)
