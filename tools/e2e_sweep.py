"""Host-buffer path (lanczos_b200_upscale_host) vs number of streams and frames per call: Gpix/s and PCIe GB/s."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lanczos_hls_b200 as lz
IW, IH, OW, OH, C = 1920, 1080, 3840, 2160, 3
DEV = int(os.environ.get("LOCAL_RANK", "0"))      # under torchrun: one process per GPU, all running at once
torch.cuda.set_device(DEV)
if os.environ.get("LZB_NCCL"):                    # same process state as bench.py under torchrun
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", DEV))
    dist.barrier()
if os.environ.get("LZB_BIND"):
    print("bound to", len(lz.bind_host_to_device(DEV) or []), "cpus", flush=True)
for frames in (16, 32):
    hin = lz.PinnedBuffer(frames * IH * IW * C); hout = lz.PinnedBuffer(frames * OH * OW * C)
    hin.array[:] = np.random.default_rng(1).integers(100, 156, hin.nbytes, dtype=np.uint8)
    h_in = hin.array.reshape(frames, IH, IW, C); h_out = hout.array.reshape(frames, OH, OW, C)
    for ns in (1, 2, 3, 4, 6, 8):
        lz.upscale(h_in, OW, OH, n_streams=ns, out=h_out, device=DEV)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(4):
            lz.upscale(h_in, OW, OH, n_streams=ns, out=h_out, device=DEV)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 4
        print(f"gpu {DEV} frames {frames} streams {ns}: {frames*OW*OH/dt/1e9:6.2f} Gpix/s  D2H {hout.nbytes/dt/1e9:5.1f} GB/s  H2D {hin.nbytes/dt/1e9:5.1f} GB/s", flush=True)
    hin.free(); hout.free()
# raw copy ceilings
n = 16 * OH * OW * C
h = lz.PinnedBuffer(n); d = torch.empty(n, dtype=torch.uint8, device="cuda")
ht = torch.from_numpy(h.array)
for name, fn in (("D2H", lambda: ht.copy_(d, non_blocking=True)), ("H2D", lambda: d.copy_(ht, non_blocking=True))):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(4): fn()
    torch.cuda.synchronize(); print(name, "raw cudaMemcpyAsync pinned:", round(4 * n / (time.perf_counter() - t0) / 1e9, 1), "GB/s")
