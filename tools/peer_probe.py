"""Probe (torchrun, >= 2 GPUs): which way of mapping a peer's buffer lets dl_push_slice write into it."""
import ctypes
import os
import sys
import traceback

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from disenlink_b200._lib import lib, stream_of  # noqa: E402


def push(src, ptrs, nbytes, dev):
    arr = (ctypes.c_void_p * len(ptrs))(*ptrs)
    rc = lib().dl_push_slice(src.data_ptr(), arr, len(ptrs), nbytes, stream_of(dev))
    torch.cuda.synchronize(dev)
    return rc


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=dev)
    n = 1 << 20
    say = lambda *a: print(f"[rank {rank}]", *a, flush=True)
    # --- A: CUDA IPC handle exported by torch, opened by dl_ipc_open with THIS device current
    try:
        buf = torch.full((world, n), -1.0, device=dev)
        h = buf.untyped_storage()._share_cuda_()
        objs = [None] * world
        hb = bytes(h[1])
        say("handle len", len(hb), "first", hb[:1])
        if len(hb) == 66:                 # torch: version byte, type byte (b'c' = cudaMalloc block), 64-byte handle
            say("type", hb[1:2])
            hb = hb[2:]
        dist.all_gather_object(objs, (hb, int(h[3])))
        ptrs = []
        for r, (hb, off) in enumerate(objs):
            if r == rank:
                continue
            base = ctypes.c_void_p()
            rc = lib().dl_ipc_open(hb, ctypes.byref(base))
            say("ipc open peer", r, "rc", rc, "base", hex(base.value or 0), "off", off)
            ptrs.append(base.value + off + rank * n * 4)
        dist.barrier()
        src = torch.full((n,), float(rank + 10), device=dev)
        rc = push(src, ptrs, n * 4, dev)
        dist.barrier()
        torch.cuda.synchronize(dev)
        say("A ipc push rc", rc, "rows now", [float(buf[r, 0]) for r in range(world)])
    except Exception:
        say("A failed:", traceback.format_exc()[-800:].replace("\n", " | "))
    # --- B: torch symmetric memory
    try:
        import torch.distributed._symmetric_memory as symm_mem
        t = symm_mem.empty((world, n), dtype=torch.float32, device=dev)
        t.fill_(-1.0)
        hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
        say("symm ptrs", [hex(p) for p in hdl.buffer_ptrs])
        ptrs = [p + rank * n * 4 for r, p in enumerate(hdl.buffer_ptrs) if r != rank]
        dist.barrier()
        src = torch.full((n,), float(rank + 20), device=dev)
        rc = push(src, ptrs, n * 4, dev)
        dist.barrier()
        torch.cuda.synchronize(dev)
        say("B symm push rc", rc, "rows now", [float(t[r, 0]) for r in range(world)])
    except Exception:
        say("B failed:", traceback.format_exc()[-1500:].replace("\n", " | "))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
