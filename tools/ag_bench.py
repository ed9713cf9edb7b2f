"""All-gather bandwidth of the partitioned step's big exchanges (torchrun, N GPUs): in-place
all_gather_into_tensor of a [N_pad, 128] fp32 array, like PartitionedLinkStep does for Z, H and dH."""
import os
import sys
import time

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=dev)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
    per = (n + world - 1) // world
    full = torch.empty(per * world, 128, dtype=torch.float32, device=dev)
    mine = full[rank * per:(rank + 1) * per]
    mine.fill_(rank)
    for _ in range(3):
        dist.all_gather_into_tensor(full, mine)
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    reps = 5
    for _ in range(reps):
        dist.all_gather_into_tensor(full, mine)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    recv = (world - 1) * per * 512
    if rank == 0:
        print(f"world={world} bytes={per * world * 512 / 1e9:.1f} GB  {ms:.2f} ms  recv/rank {recv / ms / 1e6:.0f} GB/s  "
              f"env={ {k: v for k, v in os.environ.items() if k.startswith('NCCL_') and k != 'NCCL_DEBUG_FILE'} }", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
