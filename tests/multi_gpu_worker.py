"""Worker of tests/test_gpu_fullsize.py::test_two_ranks_nccl_slab_and_batch (and a manual check under torchrun):
slab mode and batch mode of the library over NCCL against the single-GPU payloads.
    python -m torch.distributed.run --nproc-per-node N tests/multi_gpu_worker.py <outdir>"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


def main():
    import torch
    import torch.distributed as dist
    import astc_images as I
    from __graft_entry__ import load_package
    pkg = load_package()
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")          # only to ship the NCCL id: the payload gather is the library's own NCCL
    box = [pkg.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    cfg = pkg.config_init(pkg.PRF_LDR, 6, 6, pkg.PRE_MEDIUM, pkg.FLG_SELF_DECOMPRESS_ONLY)
    ctx = pkg.Context(cfg)
    ctx.comm_init(rank, world, box[0])
    ok = True
    # slab mode: ragged heights, more ranks than sensible slabs included
    for (h, w) in ((1000, 777), (64, 64), (7, 300)):
        img = I.photo_like(h, w, seed=h)
        got = ctx.compress_image_sharded(img)
        if rank == 0:
            solo = pkg.Context(cfg)
            want = solo.compress_image(img)
            solo.close()
            ok = ok and np.array_equal(got, want)
    # batch mode: image i on rank i % world, including a ragged last round
    n = 2 * world + 1
    base = I.photo_like(400, 400, seed=9)
    imgs_all = [np.ascontiguousarray(np.roll(base, 13 * i, axis=1)) for i in range(n)]
    images = [imgs_all[i] if i % world == rank else None for i in range(n)]
    nb = ((400 + 5) // 6) ** 2 * 16
    outs = [np.zeros(nb, np.uint8) for _ in range(n)] if rank == 0 else None
    ctx.compress_batch(images, outs)
    if rank == 0:
        solo = pkg.Context(cfg)
        for i in range(n):
            ok = ok and np.array_equal(outs[i], solo.compress_image(imgs_all[i]))
        solo.close()
    ctx.close()
    flag = torch.tensor([1 if ok else 0])
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    if rank == 0 and ok and len(sys.argv) > 1:
        open(os.path.join(sys.argv[1], "ok"), "w").write("ok\n")
    if not int(flag[0]):
        sys.exit(1)


if __name__ == "__main__":
    main()
