"""dev tool (GPU box): device-resident path with float input - experiments"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import astc_images as I
from astc_ref import block_diff
import __graft_entry__ as g
pkg = g.load_package()
dev = torch.device("cuda", 0)
dims = [int(a) for a in sys.argv[1:]] or [96]
def run_dev(ctx, ptr, dtype_id, dim, stream=None, swz=(0, 1, 2, 3)):
    nbx, nby = ctx.blocks(dim, dim)
    d_out = torch.zeros(nbx * nby * 16, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    ctx.compress_device(ptr, dtype_id, dim, dim, d_out.data_ptr(), stream=stream.cuda_stream if stream is not None else 0, swizzle=swz)
    torch.cuda.synchronize()
    return d_out.cpu().numpy()
st = torch.cuda.Stream(device=dev)
for dim in dims:
  img16 = I.hdr_noise(dim, dim, seed=7)
  img8 = I.photo_like(dim, dim, seed=7)
  for name, prof, img, swz in [("u8 swizzled (generic load path)", 1, img8, (2, 1, 0, 3)), ("f16 hdr", 3, img16, (0, 1, 2, 3)), ("f16 ldr", 1, img16, (0, 1, 2, 3)), ("f16 hdr dev first", 3, img16, (0, 1, 2, 3))]:
    ctx = pkg.Context(pkg.config_init(prof, 6, 6, 60.0, 32))
    tid = {np.dtype(np.uint8): 0, np.dtype(np.float16): 1, np.dtype(np.float32): 2}[img.dtype]
    t = torch.from_numpy(img.view(np.uint8)).to(dev)
    if "first" in name:
        a = run_dev(ctx, t.data_ptr(), tid, dim, st, swz)
        h = ctx.compress_image(img, swizzle=swz)
    else:
        h = ctx.compress_image(img, swizzle=swz)
        a = run_dev(ctx, t.data_ptr(), tid, dim, st, swz)
    b = run_dev(ctx, t.data_ptr(), tid, dim, None, swz)
    big = torch.zeros(t.numel() + 4096, dtype=torch.uint8, device=dev)
    off = (-big.data_ptr()) % 4096
    big[off:off + t.numel()] = t.reshape(-1)
    c = run_dev(ctx, big.data_ptr() + off, tid, dim, st, swz)
    print(dim, name, "| torch stream", len(block_diff(a, h)), "| ctx stream", len(block_diff(b, h)), "| 4096-aligned copy", len(block_diff(c, h)), "| a==b", np.array_equal(a, b))
    ctx.close()
