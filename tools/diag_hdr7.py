"""dev tool (GPU box): compare the per-block records (texels, block statistics) of a host-path and a device-path pass"""
import os, sys, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import astc_images as I
from astc_ref import block_diff
import __graft_entry__ as g
pkg = g.load_package()
lib = pkg.lib()
lib.astcenc_b200_debug_records.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
cudart = C.CDLL([l.split()[-1] for l in open("/proc/self/maps") if "libcudart" in l][0]) if any("libcudart" in l for l in open("/proc/self/maps")) else None
dev = torch.device("cuda", 0)
dim = int(sys.argv[1]) if len(sys.argv) > 1 else 128
img = I.hdr_noise(dim, dim, seed=7)
ctx = pkg.Context(pkg.config_init(3, 6, 6, 60.0, 32))
nbx, nby = ctx.blocks(dim, dim)
nblk = nbx * nby
def records():
    p = C.c_void_p(); rb = C.c_size_t(); cap = C.c_size_t()
    lib.astcenc_b200_debug_records(ctx.handle, C.byref(p), C.byref(rb), C.byref(cap))
    n = nblk * rb.value
    # wrap the raw pointer with torch through the cuda array interface
    class Raw:
        pass
    r = Raw()
    r.__cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (p.value, False), "version": 2}
    return torch.as_tensor(r, device=dev).cpu().numpy().reshape(nblk, rb.value).copy(), rb.value
h = ctx.compress_image(img)
rh, rb = records()
t = torch.from_numpy(img.view(np.uint8)).to(dev)
d_out = torch.zeros(nblk * 16, dtype=torch.uint8, device=dev)
st = torch.cuda.Stream(device=dev)
torch.cuda.synchronize()
ctx.compress_device(t.data_ptr(), 1, dim, dim, d_out.data_ptr(), stream=st.cuda_stream)
torch.cuda.synchronize()
rd, _ = records()
print("blocks", nblk, "record bytes", rb, "block diffs", len(block_diff(d_out.cpu().numpy(), h)))
# record = arena head [0, A_PERSIST) then texels 4 x Tp floats
T = 36
head = rb - 4 * T * 4
texh = rh[:, head:].view(np.float32); texd = rd[:, head:].view(np.float32)
print("records with different texels:", int((texh != texd).any(axis=1).sum()), "of", nblk)
bih = rh[:, :112]; bid = rd[:, :112]
print("records with different BlkInfo:", int((bih != bid).any(axis=1).sum()))
k = np.where((texh != texd).any(axis=1))[0]
if len(k):
    b = k[0]
    print("first differing record", b, "texel diffs at", np.where(texh[b] != texd[b])[0][:20], texh[b][:8], texd[b][:8])
kb = np.where((bih != bid).any(axis=1))[0]
if len(kb):
    b = kb[0]
    print("BlkInfo diff record", b, "bytes", np.where(bih[b] != bid[b])[0][:40])
    print(" host", bih[b].view(np.float32)[:20]); print(" dev ", bid[b].view(np.float32)[:20])
ctx.close()
