"""N > 1 host logic on CPU: world_size 2 over gloo.

The multi-GPU path shards an image by block-row slabs (astc-encoder_b200.slab_rows), every rank compresses its slab
with no data-path collective, and the payload is gathered on rank 0. Here the per-rank compressor is the ORACLE
(the checker - there is no GPU in this test); what is under test is the slab arithmetic, the gather plumbing and the
claim that slabs compose byte for byte into the whole-image result.
"""
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmpdir, dim_x, dim_y, bx, by):
    import torch
    import torch.distributed as dist
    from astc_ref import Oracle, PRF_LDR, PRE_FAST, FLG_SELF_DECOMPRESS_ONLY
    import astc_images as I
    from __graft_entry__ import load_package
    pkg = load_package()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        img = I.photo_like(dim_y, dim_x, seed=5)
        blocks_x = (dim_x + bx - 1) // bx
        blocks_y = (dim_y + by - 1) // by
        r0, r1 = pkg.slab_rows(blocks_y, rank, world)
        # every rank sees the whole image (as on the GPUs) and compresses only its block rows
        slab = np.ascontiguousarray(img[r0 * by:min(r1 * by, dim_y)])
        orc = Oracle()
        mine = np.frombuffer(orc.compress(slab, PRF_LDR, bx, by, PRE_FAST, FLG_SELF_DECOMPRESS_ONLY), dtype=np.uint8)
        assert mine.size == (r1 - r0) * blocks_x * 16
        # ragged gather: slabs may differ by one block row -> pad to the largest, trim on the root
        sizes = [(pkg.slab_rows(blocks_y, r, world)[1] - pkg.slab_rows(blocks_y, r, world)[0]) * blocks_x * 16 for r in range(world)]
        buf = torch.zeros(max(sizes), dtype=torch.uint8)
        buf[:mine.size] = torch.from_numpy(mine.copy())
        gathered = [torch.zeros(max(sizes), dtype=torch.uint8) for _ in range(world)] if rank == 0 else None
        dist.gather(buf, gathered, dst=0)
        if rank == 0:
            whole = np.concatenate([g.numpy()[:n] for g, n in zip(gathered, sizes)])
            full = np.frombuffer(orc.compress(img, PRF_LDR, bx, by, PRE_FAST, FLG_SELF_DECOMPRESS_ONLY), dtype=np.uint8)
            np.save(os.path.join(tmpdir, "ok.npy"), np.array([int(np.array_equal(whole, full)), whole.size, full.size]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("dims", [(96, 100, 6, 6), (64, 36, 4, 4), (50, 30, 8, 5)])
def test_two_ranks_compose_whole_image(tmp_path, dims, oracle):
    import torch.multiprocessing as mp
    dim_x, dim_y, bx, by = dims
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path), dim_x, dim_y, bx, by), nprocs=2, join=True)
    ok = np.load(os.path.join(str(tmp_path), "ok.npy"))
    assert ok[1] == ok[2]
    assert ok[0] == 1, "slabs of two ranks do not compose into the whole-image payload"


def test_slab_rows_cover_all_block_rows(pkg):
    for blocks_y in (1, 2, 3, 7, 683):
        for world in (1, 2, 3, 4, 8):
            spans = [pkg.slab_rows(blocks_y, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == blocks_y
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
