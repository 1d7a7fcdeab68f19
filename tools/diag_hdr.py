"""dev tool (GPU box): run-to-run determinism and host-vs-device path equality for a config (default: HDR F16 6x6 medium)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import astc_images as I
from astc_ref import block_diff, ref_lib, have_ref
import __graft_entry__ as g
pkg = g.load_package()
dim = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
prof = int(sys.argv[2]) if len(sys.argv) > 2 else 3
img = I.hdr_noise(dim, dim, seed=2024) if prof >= 2 else I.photo_like(dim, dim, seed=2024)
cfg = pkg.config_init(prof, 6, 6, 60.0, 32)
outs = {}
for bands in ("1", "4"):
    os.environ["ASTCENC_B200_UPLOAD_BANDS"] = bands
    ctx = pkg.Context(cfg)
    for rep in range(3):
        outs["host_b%s_%d" % (bands, rep)] = ctx.compress_image(img).copy()
    dev = torch.device("cuda", 0)
    d_img = torch.from_numpy(img.view(np.uint8) if img.dtype != np.uint8 else img).to(dev)
    nbx, nby = ctx.blocks(dim, dim)
    d_out = torch.zeros(nbx * nby * 16, dtype=torch.uint8, device=dev)
    st = torch.cuda.Stream(device=dev)
    for rep in range(3):
        ctx.compress_device(d_img.data_ptr(), pkg.TYPE_F16 if img.dtype != np.uint8 else pkg.TYPE_U8, dim, dim, d_out.data_ptr(), stream=st.cuda_stream)
        torch.cuda.synchronize()
        outs["dev_b%s_%d" % (bands, rep)] = d_out.cpu().numpy().copy()
    ctx.close()
if have_ref():
    outs["ref"] = np.frombuffer(bytes(ref_lib().compress(img, prof, 6, 6, 60.0, 32, threads=16)), dtype=np.uint8)
keys = list(outs)
base = outs["ref"] if "ref" in outs else outs[keys[0]]
for k in keys:
    d = block_diff(outs[k], base)
    print(k, "diff vs ref:", len(d), d[:6])
