"""Aggregate pinned-memory copy bandwidth with one process per GPU copying at the same time (torchrun):
the platform ceiling of the host-buffer path when several ranks share one host."""
import os, time, torch, torch.distributed as dist
rank, lrank, world = (int(os.environ.get(k, "0")) for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"))
torch.cuda.set_device(lrank)
dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
n = 400 << 20
h = torch.empty(n, dtype=torch.uint8, pin_memory=True); d = torch.empty(n, dtype=torch.uint8, device="cuda")
h2 = torch.empty(n // 4, dtype=torch.uint8, pin_memory=True); d2 = torch.empty(n // 4, dtype=torch.uint8, device="cuda")
s2 = torch.cuda.Stream()
for name in ("D2H", "H2D", "D2H + H2D/4 concurrently"):
    def go():
        if name == "D2H": h.copy_(d, non_blocking=True)
        elif name == "H2D": d.copy_(h, non_blocking=True)
        else:
            h.copy_(d, non_blocking=True)
            with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
    go(); torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): go()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    t = torch.tensor([dt], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        b = n if name != "D2H + H2D/4 concurrently" else n + n // 4
        print(f"{world} ranks, {name}: {5 * b / t.item() / 1e9:.1f} GB/s per rank, {world * 5 * b / t.item() / 1e9:.1f} GB/s aggregate", flush=True)
dist.destroy_process_group()
