"""Platform check for the e2e number: plain pinned device->host copies, one process per GPU, all at once.
torchrun --nproc-per-node N tests/tools/d2h_bw.py   (or plain python for one GPU)"""
import os, time
import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 90316800                                    # int16 samples of one bench step
dev = torch.zeros(n, dtype=torch.int16, device="cuda")
host = torch.empty(n, dtype=torch.int16).pin_memory()
for _ in range(3):
    host.copy_(dev, non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
reps = 20
for _ in range(reps):
    host.copy_(dev, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
gbs = torch.tensor([n * 2 * reps / dt / 1e9], dtype=torch.float64, device="cuda")
if world > 1:
    lo = gbs.clone(); dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    tot = gbs.clone(); dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    if rank == 0:
        print(f"{world} GPUs copying at once: min {float(lo):.1f} GB/s per GPU, {float(tot):.1f} GB/s total")
    dist.destroy_process_group()
else:
    print(f"1 GPU: {float(gbs):.1f} GB/s")
