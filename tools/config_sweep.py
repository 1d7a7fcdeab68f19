"""Dev tool (GPU box): the concrete inputs of SURVEY.md §8(d) beyond the bench line - for each configuration the product
through the C ABI with host buffers (best of N), the reference build on all host threads (best of N), and the share of
byte-identical blocks; plus decode and error-metric throughput. Writes a markdown table to stdout.
    python tools/config_sweep.py [--quick]
"""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from astc_ref import *
import astc_images as I

quick = "--quick" in sys.argv
tail_only = "--tail" in sys.argv          # only the rows after (s2)


def tiled(base, h, w):
    """mirror-tile a small base image (the voronoi generator is O(cells x texels))"""
    ny, nx = (h + base.shape[0] - 1) // base.shape[0], (w + base.shape[1] - 1) // base.shape[1]
    rows = []
    for ty in range(ny):
        row = [(base[::-1] if ty & 1 else base)[:, ::-1] if tx & 1 else (base[::-1] if ty & 1 else base) for tx in range(nx)]
        rows.append(np.concatenate(row, axis=1))
    return np.ascontiguousarray(np.concatenate(rows, axis=0)[:h, :w])
prod = AstcencLib(os.path.join(ROOT, "astc-encoder_b200", "libastcenc_b200.so"))
ref = ref_lib() if have_ref() else None
threads = os.cpu_count() or 1


def best(fn, n):
    ts = []
    out = None
    for _ in range(n):
        t0 = time.perf_counter(); out = fn(); ts.append(time.perf_counter() - t0)
    return min(ts), out


def compress_row(name, img, prof, bx, by, q, flags=0, reps=3, ref_reps=2):
    h, w = img.shape[:2]
    cfg = prod.config(prof, bx, by, q, flags | FLG_SELF_DECOMPRESS_ONLY)
    ctx = C.c_void_p()
    assert prod.lib.astcenc_context_alloc(C.byref(cfg), 1, C.byref(ctx), None) == 0
    prod.compress_ctx(ctx, img, bx, by)                     # warm-up (allocations)
    tg, g = best(lambda: prod.compress_ctx(ctx, img, bx, by), reps)
    prod.lib.astcenc_context_free(ctx)
    line = "| %s | %dx%d %s | %dx%d q=%g | %.1f |" % (name, w, h, img.dtype, bx, by, q, w * h / tg / 1e6)
    if ref is not None:
        rcfg = ref.config(prof, bx, by, q, flags | FLG_SELF_DECOMPRESS_ONLY)
        rctx = C.c_void_p()
        assert ref.lib.astcenc_context_alloc(C.byref(rcfg), threads, C.byref(rctx), None) == 0
        tr, r = best(lambda: ref.compress_ctx(rctx, img, bx, by, threads=threads), ref_reps)
        ref.lib.astcenc_context_free(rctx)
        same = 100.0 * (1.0 - len(block_diff(g, r)) / (len(g) // 16))
        line += " %.1f (%d thr) | %.1fx | %.3f %% |" % (w * h / tr / 1e6, threads, tr / tg, same)
    else:
        line += " - | - | - |"
    print(line); sys.stdout.flush()
    return g


print("| input | image | block / quality | B200 e2e MT/s | reference MT/s | ratio | identical blocks |")
print("|---|---|---|---|---|---|---|")
S = 2 if quick else 1
big = I.photo_like(4096 // S, 4096 // S)
if not tail_only:
    compress_row("(1) crop", I.photo_like(512, 512, seed=5), PRF_LDR, 4, 4, PRE_FAST, reps=5, ref_reps=5)
blocks66 = compress_row("(2) photo-like", big, PRF_LDR, 6, 6, PRE_MEDIUM, ref_reps=1 if tail_only else 2)
if not tail_only:
    compress_row("(3) photo-like", big, PRF_LDR, 8, 8, PRE_THOROUGH, reps=2, ref_reps=1)
    hdr = I.hdr_noise(2048 // S, 2048 // S)
    compress_row("(4) HDR noise -cH", hdr, PRF_HDR, 6, 6, PRE_MEDIUM)
    compress_row("(4) HDR noise -ch", hdr, PRF_HDR_RGB_LDR_A, 6, 6, PRE_MEDIUM)
    compress_row("(s1) uniform noise", I.uniform_noise(2048 // S, 2048 // S), PRF_LDR, 6, 6, PRE_MEDIUM)
    compress_row("(s2) gradients", I.smooth_gradient(2048 // S, 2048 // S), PRF_LDR, 6, 6, PRE_MEDIUM)
compress_row("(s3) voronoi (256^2 tile)", tiled(I.voronoi_flat(256, 256), 2048 // S, 2048 // S), PRF_LDR, 6, 6, PRE_MEDIUM)
compress_row("(s5) alpha masks", I.alpha_mask(2048 // S, 2048 // S), PRF_LDR, 6, 6, PRE_MEDIUM)

# decode + metrics on configuration (2)
h, w = big.shape[:2]
td, dec = best(lambda: prod.decompress(blocks66, w, h, PRF_LDR, 6, 6), 3)
line = "\ndecompress %dx%d 6x6 -> U8: B200 e2e %.1f MT/s" % (w, h, w * h / td / 1e6)
if ref is not None:
    tr, rdec = best(lambda: ref.decompress(blocks66, w, h, PRF_LDR, 6, 6), 2)
    line += ", reference (1 thread, as the CLI decodes) %.1f MT/s, images identical: %s" % (w * h / tr / 1e6, bool(np.array_equal(dec, rdec)))
print(line)
tm, m = best(lambda: prod.error_metrics(big, dec), 3)
orc = Oracle()
to, mo = best(lambda: orc.error_metrics(big, dec), 1)
print("error metrics %dx%d RGBA8: B200 e2e %.1f MT/s (PSNR %.4f dB), CPU restatement (1 thread) %.1f MT/s (PSNR %.4f dB)" %
      (w, h, w * h / tm / 1e6, m["psnr"], w * h / to / 1e6, mo["psnr"]))
