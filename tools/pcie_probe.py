"""Developer probe: pinned H2D / D2H bandwidth per rank, alone and concurrently (run under torchrun)."""
import os, time, torch, torch.distributed as dist
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(rank)
if world > 1:
    os.environ.pop("NCCL_DEBUG", None)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{rank}"))
n = 512 * 1024 * 1024  # 2 GiB of float32
h = torch.empty(n, dtype=torch.float32, pin_memory=True); h.fill_(1.0)
d = torch.empty(n, dtype=torch.float32, device="cuda")
def bw(fn, reps=3):
    fn(); torch.cuda.synchronize()
    if world > 1: dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return reps * n * 4 / (time.perf_counter() - t0) / 1e9
r1 = bw(lambda: d.copy_(h, non_blocking=True)); r2 = bw(lambda: h.copy_(d, non_blocking=True))
print(f"rank {rank}/{world}: concurrent H2D {r1:.1f} GB/s  D2H {r2:.1f} GB/s", flush=True)
if world > 1:
    for r in range(world):
        dist.barrier(); torch.cuda.synchronize()
        if r == rank:
            a = n * 4 * 3; t0 = time.perf_counter()
            for _ in range(3): d.copy_(h, non_blocking=True)
            torch.cuda.synchronize(); print(f"rank {rank} alone H2D {a / (time.perf_counter() - t0) / 1e9:.1f} GB/s", flush=True)
        dist.barrier()
    dist.destroy_process_group()
