"""torch.distributed (NCCL) all-reduce time by message size on this box; launch with torchrun.  Also times the same
collective issued through the ctypes callback path the sharded VEGAS loop uses (host overhead of one Python call)."""
import os
import sys
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
for n in (4, 4096, 65540, 65536 * 8 + 4, 16 * 2**20, 160 * 2**20):
    t = torch.ones(n, dtype=torch.float64, device=dev)
    for _ in range(5):
        dist.all_reduce(t)
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    reps = 50 if n < 2**22 else 5
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0 = time.perf_counter()
    a.record()
    for _ in range(reps):
        dist.all_reduce(t)
    b.record()
    h1 = time.perf_counter()
    torch.cuda.synchronize()
    dt_dev = a.elapsed_time(b) * 1e-3 / reps
    if rank == 0:
        print(f"world={world} all_reduce fp64 x {n:>10d} ({n*8/1e6:9.3f} MB): device {dt_dev*1e6:9.1f} us/op, host enqueue "
              f"{(h1-h0)/reps*1e6:7.1f} us/op, algbw {n*8/dt_dev/1e9:7.1f} GB/s", flush=True)
dist.destroy_process_group()
