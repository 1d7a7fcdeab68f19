"""dev tool: N device-resident passes of config 1 (4096^2 6x6 -medium) - the command ncu wraps. python tools/one_pass.py [passes] [dim]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import astc_images as I
import __graft_entry__ as g
pkg = g.load_package()
passes = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
dev = torch.device("cuda", 0)
img = np.ascontiguousarray(I.photo_like(dim, dim, seed=2024))
d_img = torch.from_numpy(img).to(dev)
ctx = pkg.Context(pkg.config_init(1, 6, 6, 60.0, 32))
nbx, nby = ctx.blocks(dim, dim)
d_out = torch.zeros(nbx * nby * 16, dtype=torch.uint8, device=dev)
st = torch.cuda.Stream(device=dev)
torch.cuda.synchronize()
for _ in range(passes):
    ctx.compress_device(d_img.data_ptr(), 0, dim, dim, d_out.data_ptr(), stream=st.cuda_stream)
    torch.cuda.synchronize()
print("launches", ctx.launch_count())
ctx.close()
