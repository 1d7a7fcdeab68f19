"""dev tool (GPU box): sweep environment knobs (read at context creation) on one resident image; prints ms per pass.
   python tools/sweep_knobs.py [--dim 4096] [--lib path.so] "K1=v1,K2=v2" "K1=v3" ...     ("-" = no knobs)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import astc_images as I
import __graft_entry__ as g
pkg = g.load_package()
args = sys.argv[1:]
dim = 4096
if "--dim" in args:
    i = args.index("--dim"); dim = int(args[i + 1]); del args[i:i + 2]
if "--lib" in args:
    i = args.index("--lib"); os.environ["ASTCENC_B200_LIB"] = os.path.abspath(args[i + 1]); del args[i:i + 2]
    pkg.LIB_PATH = os.environ["ASTCENC_B200_LIB"]
vals = args
touched = set()
dev = torch.device("cuda", 0)
img = np.ascontiguousarray(I.photo_like(dim, dim, seed=2024))
d_img = torch.from_numpy(img).to(dev)
st = torch.cuda.Stream(device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ref = None
for v in vals:
    for k in touched:
        os.environ.pop(k, None)
    if v != "-":
        for kv in v.split(","):
            k, x = kv.split("=")
            os.environ[k] = x
            touched.add(k)
    ctx = pkg.Context(pkg.config_init(1, 6, 6, 60.0, 32))
    nbx, nby = ctx.blocks(dim, dim)
    d_out = torch.zeros(nbx * nby * 16, dtype=torch.uint8, device=dev)
    ts = []
    with torch.cuda.stream(st):
        for rep in range(6):
            flush.zero_()
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record(st)
            ctx.compress_device(d_img.data_ptr(), 0, dim, dim, d_out.data_ptr(), stream=st.cuda_stream)
            b.record(st)
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
    out = d_out.cpu().numpy()
    if ref is None:
        ref = out
    print("%-60s %.2f ms (min %.2f)  same=%s" % (v, float(np.median(ts[2:])), min(ts[2:]), np.array_equal(out, ref)), flush=True)
    ctx.close()
