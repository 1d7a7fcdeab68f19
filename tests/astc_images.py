"""Seeded synthetic test images shared by the tests and bench.py."""
import numpy as np


def _value_noise(rng, h, w, cell, channels):
    gh, gw = h // cell + 2, w // cell + 2
    g = rng.random((gh, gw, channels), dtype=np.float32)
    ys = np.arange(h, dtype=np.float32) / cell
    xs = np.arange(w, dtype=np.float32) / cell
    y0 = ys.astype(np.int32); x0 = xs.astype(np.int32)
    fy = (ys - y0)[:, None, None]; fx = (xs - x0)[None, :, None]
    fy = fy * fy * (3 - 2 * fy); fx = fx * fx * (3 - 2 * fx)
    a = g[y0][:, x0]; b = g[y0][:, x0 + 1]; c = g[y0 + 1][:, x0]; d = g[y0 + 1][:, x0 + 1]
    return (a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy


def _jitter_voronoi(rng, h, w, cell, channels):
    gh, gw = h // cell + 3, w // cell + 3
    py = (np.arange(gh)[:, None] - 1 + rng.random((gh, gw))) * cell
    px = (np.arange(gw)[None, :] - 1 + rng.random((gh, gw))) * cell
    cols = rng.random((gh, gw, channels)).astype(np.float32)
    yy = np.arange(h, dtype=np.float32)[:, None]
    xx = np.arange(w, dtype=np.float32)[None, :]
    cy = (np.arange(h) // cell + 1)[:, None]
    cx = (np.arange(w) // cell + 1)[None, :]
    best = np.full((h, w), 1e30, np.float32)
    out = np.zeros((h, w, channels), np.float32)
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            iy = np.broadcast_to(cy + dy, (h, w))
            ix = np.broadcast_to(cx + dx, (h, w))
            d = (yy - py[iy, ix]) ** 2 + (xx - px[iy, ix]) ** 2
            m = d < best
            best = np.where(m, d, best)
            out[m] = cols[iy, ix][m]
    return out


def photo_like(h, w, seed=1234):
    """Texture-like content. Images larger than 1024 in either axis are a mirrored tiling of a 1024-sized
    base image (keeps generation fast; block statistics are those of the base image)."""
    if h > 1024 or w > 1024:
        bh, bw = min(h, 1024), min(w, 1024)
        base = _photo_like_base(bh, bw, seed)
        ny, nx = (h + bh - 1) // bh, (w + bw - 1) // bw
        rows = []
        for ty in range(ny):
            row = []
            for tx in range(nx):
                t = base
                if ty & 1:
                    t = t[::-1]
                if tx & 1:
                    t = t[:, ::-1]
                row.append(t)
            rows.append(np.concatenate(row, axis=1))
        return np.ascontiguousarray(np.concatenate(rows, axis=0)[:h, :w])
    return _photo_like_base(h, w, seed)


def _photo_like_base(h, w, seed=1234):
    """Multi-octave colour noise + Voronoi edges + flat patches + partly varying alpha: a mix of block types
    (constant, smooth, edge, textured) comparable to real texture content."""
    rng = np.random.default_rng(seed)
    img = np.zeros((h, w, 4), np.float32)
    amp = 1.0; tot = 0.0
    for cell in (256, 64, 16, 4):
        if cell <= max(4, min(h, w) // 2):
            img += amp * _value_noise(rng, h, w, cell, 4)
            tot += amp
        amp *= 0.5
    img /= max(tot, 1e-6)
    # jittered-grid Voronoi cells with per-cell tint -> hard edges (9-neighbour search, vectorised)
    lab_cols = _jitter_voronoi(rng, h, w, 48, 4)
    mix = _value_noise(rng, h, w, 128, 1)
    img = img * (0.55 + 0.45 * mix) + lab_cols * (0.45 * (1 - mix))
    # flat patches and opaque alpha regions
    flat = _value_noise(rng, h, w, 96, 1)[..., 0] > 0.72
    img[flat] = np.round(img[flat] * 6) / 6
    opaque = _value_noise(rng, h, w, 160, 1)[..., 0] > 0.45
    img[..., 3][opaque] = 1.0
    return (np.clip(img, 0, 1) * 255.0 + 0.5).astype(np.uint8)


def uniform_noise(h, w, seed=1234):                   # s1: nothing exits early
    return np.random.default_rng(seed).integers(0, 256, size=(h, w, 4), dtype=np.uint8)


def smooth_gradient(h, w, seed=1234):                 # s2: mode0 exits
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.stack([xx / w, yy / h, (xx + yy) / (w + h), np.ones_like(xx)], axis=-1)
    img += rng.normal(0, 0.01, img.shape).astype(np.float32)
    return (np.clip(img, 0, 1) * 255.0 + 0.5).astype(np.uint8)


def voronoi_flat(h, w, cell=9, seed=1234):            # s3: partition search stress
    rng = np.random.default_rng(seed)
    n = max(2, (h * w) // (cell * cell))
    pts = rng.random((n, 2)) * [h, w]
    cols = rng.integers(0, 256, size=(n, 4), dtype=np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]
    best = np.full((h, w), 1e30); lab = np.zeros((h, w), np.int32)
    for i in range(n):
        d = (yy - pts[i, 0]) ** 2 + (xx - pts[i, 1]) ** 2
        m = d < best
        best[m] = d[m]; lab[m] = i
    return cols[lab]


def constant(h, w, rgba=(12, 200, 99, 255)):          # s4
    img = np.zeros((h, w, 4), np.uint8)
    img[:] = rgba
    return img


def alpha_mask(h, w, seed=1234):                      # s5
    img = photo_like(h, w, seed)
    rng = np.random.default_rng(seed + 1)
    img[..., 3] = np.where(_value_noise(rng, h, w, 24, 1)[..., 0] > 0.5, 255, 0).astype(np.uint8)
    return img


def hdr_noise(h, w, seed=1234, dtype=np.float16):     # s6
    rng = np.random.default_rng(seed)
    base = _value_noise(rng, h, w, 32, 3) * 16 - 8 + rng.normal(0, 0.3, (h, w, 3)).astype(np.float32)
    rgb = np.exp2(base).astype(np.float32)
    a = _value_noise(rng, h, w, 40, 1)
    return np.concatenate([rgb, a], axis=-1).astype(dtype)
