"""dev tool (GPU box): why does the device-resident path differ for F16 input?"""
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
dim = 256
def run_dev(ctx, t, dtype_id, dim):
    nbx, nby = ctx.blocks(dim, dim)
    d_out = torch.zeros(nbx * nby * 16, dtype=torch.uint8, device=dev)
    st = torch.cuda.Stream(device=dev)
    torch.cuda.synchronize()
    ctx.compress_device(t.data_ptr(), dtype_id, dim, dim, d_out.data_ptr(), stream=st.cuda_stream)
    torch.cuda.synchronize()
    return d_out.cpu().numpy()
for prof, gen, dt in [(1, "photo_like", np.uint8), (3, "hdr_noise", np.float16), (3, "hdr_noise", np.float32), (1, "hdr_noise", np.float16)]:
    img = getattr(I, gen)(dim, dim, seed=7)
    if img.dtype != dt:
        img = img.astype(dt)
    cfg = pkg.config_init(prof, 6, 6, 60.0, 32)
    ctx = pkg.Context(cfg)
    tid = {np.dtype(np.uint8): 0, np.dtype(np.float16): 1, np.dtype(np.float32): 2}[img.dtype]
    t_u8 = torch.from_numpy(img.view(np.uint8)).to(dev)
    back = t_u8.cpu().numpy().tobytes() == img.tobytes()
    a = run_dev(ctx, t_u8, tid, dim)            # device path first, fresh context
    h = ctx.compress_image(img)
    b = run_dev(ctx, t_u8, tid, dim)
    t_nat = torch.from_numpy(img).to(dev)
    c = run_dev(ctx, t_nat, tid, dim)
    print(prof, gen, img.dtype, "bytes ok", back, "dev-first vs host", len(block_diff(a, h)), "dev-after", len(block_diff(b, h)), "native tensor", len(block_diff(c, h)),
          "zeros?", int(a.sum()) == 0, "first blocks", a[:16].tolist(), h[:16].tolist())
    ctx.close()
