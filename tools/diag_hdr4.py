"""dev tool (GPU box): device-resident float input - which property of the source pointer matters?"""
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
dim = 128
st = torch.cuda.Stream(device=dev)
def run_dev(ctx, ptr, dtype_id, dim):
    nbx, nby = ctx.blocks(dim, dim)
    d_out = torch.zeros(nbx * nby * 16, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    ctx.compress_device(ptr, dtype_id, dim, dim, d_out.data_ptr(), stream=st.cuda_stream)
    torch.cuda.synchronize()
    return d_out.cpu().numpy()
img = I.hdr_noise(dim, dim, seed=7)
ctx = pkg.Context(pkg.config_init(3, 6, 6, 60.0, 32))
h = ctx.compress_image(img)
t = torch.from_numpy(img.view(np.uint8)).to(dev)
n = t.numel()
print("t ptr %x" % t.data_ptr(), "numel", n)
a = run_dev(ctx, t.data_ptr(), 1, dim)
print("plain tensor:", len(block_diff(a, h)))
c = t.clone()
print("clone ptr %x:" % c.data_ptr(), len(block_diff(run_dev(ctx, c.data_ptr(), 1, dim), h)))
for pad_fill in (0, 0xAB):
    big = torch.full((n + 3 * 8192,), pad_fill, dtype=torch.uint8, device=dev)
    base = (-big.data_ptr()) % 4096 + 4096
    for off in (0, 8, 16, 64, 256, 512, 1024, 2048):
        o = base + off
        big[o:o + n] = t.reshape(-1)
        r = run_dev(ctx, big.data_ptr() + o, 1, dim)
        d = block_diff(r, h)
        print("fill %02x offset %4d: %d diffs" % (pad_fill, off, len(d)), d[:8])
        big[o:o + n] = pad_fill
# does the pass modify its input?
t2 = torch.from_numpy(img.view(np.uint8)).to(dev)
run_dev(ctx, t2.data_ptr(), 1, dim)
print("input unchanged after the pass:", t2.cpu().numpy().tobytes() == img.tobytes())
ctx.close()
