"""world_size-2 gloo tests (CPU) of the N>1 host logic: frames and row bands are sharded with no
data-path collective; the only communication is the barrier/max-time plumbing of bench.py."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_py as O
    import lanczos_hls_b200 as lz
    from lanczos_hls_b200.sharding import band_range, frame_range

    # ---- row bands of one image: each rank computes its band from its own halo rows only
    iw, ih, n, d, a, c = 60, 50, 17, 10, 3, 3
    ow, oh = O.out_dims(iw, ih, n, d)
    img = O.xorshift_bytes(ih * iw * c).reshape(ih, iw, c)
    desc = lz.make_desc(iw, ih, ow, oh, c, a, n, d)
    r0, r1 = band_range(oh, rank, world)
    in0, inn = lz.band_input_rows(desc, r0, r1 - r0)
    full = O.upscale(img, ow, oh, a, n, d)
    # the band only needs rows [in0, in0+inn): zero everything else and the band must not change
    masked = np.zeros_like(img)
    masked[in0:in0 + inn] = img[in0:in0 + inn]
    band = O.upscale(masked, ow, oh, a, n, d, rows=(r0, r1 - r0))
    ok_band = bool(np.array_equal(band, full[r0:r1]))

    # ---- frame batch: contiguous frame ranges, checksum of checksums via the timing all-reduce path
    f0, f1 = frame_range(7, rank, world)
    local = torch.tensor([float(f1 - f0), float(sum(range(f0, f1)))], dtype=torch.float64)
    dist.all_reduce(local, op=dist.ReduceOp.SUM)
    t = torch.tensor([0.5 + rank], dtype=torch.float64)  # bench.py takes MAX over ranks of the step time
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.barrier()
    q.put((rank, ok_band, (r0, r1), local.tolist(), t.item()))
    dist.destroy_process_group()


def test_two_rank_sharding():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res)                       # every band equals its slice of the full result
    assert res[0][2][1] == res[1][2][0] and res[0][2][0] == 0 and res[1][2][1] == 85   # bands tile [0, out_h)
    assert res[0][3] == [7.0, 21.0]                     # 7 frames, each exactly once
    assert res[0][4] == 1.5                             # max over ranks
