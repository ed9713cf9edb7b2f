"""All-to-all of row blocks between the ranks of one node, three ways (torchrun, N GPUs): what bounds the
halo exchange of the partitioned step?  Every rank sends `rows` 512-byte rows to every peer.

  push        dl_push_rows over CUDA-IPC peer mappings, index list per peer (what PartitionedLinkStep does)
  push_contig the same kernel without index lists (contiguous source block per peer)
  push_slice  dl_push_slice: ONE contiguous block written to every peer (round-1 all-gather pattern)
  nccl_a2a    torch.distributed.all_to_all_single (NCCL send / recv pairs)

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 \\
        tools/a2a_bench.py [rows_per_peer]
"""
import ctypes
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from disenlink_b200._lib import DlPushDesc, check, lib, stream_of  # noqa: E402
from disenlink_b200.partition import PeerExchange  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
    D = 128
    src = torch.randn(rows * world, D, device=dev)                 # block q = what goes to peer q
    dst = torch.zeros(rows * world, D, device=dev)                 # block p = what arrives from peer p
    px = PeerExchange(world, rank, dev)
    assert px.register(dst), "peer mapping unavailable"
    bases = px.peers[id(dst)]
    peers = [q for q in range(world) if q != rank]
    row_bytes = D * 4
    idx = {q: (torch.arange(rows, device=dev, dtype=torch.int32) + q * rows) for q in peers}

    def push(with_idx):
        descs = (DlPushDesc * len(peers))()
        for i, q in enumerate(peers):
            descs[i].dst = bases[q] + rank * rows * row_bytes
            descs[i].src_idx = idx[q].data_ptr() if with_idx else None
            descs[i].dst_idx = None
            descs[i].mask = None
            descs[i].n = rows
        if with_idx:
            check(lib().dl_push_rows(src.data_ptr(), row_bytes, 0, descs, len(peers), stream_of(dev)), "push")
        else:                                                     # contiguous: one launch per peer, source offset
            for i, q in enumerate(peers):
                one = (DlPushDesc * 1)()
                one[0].dst, one[0].src_idx, one[0].dst_idx, one[0].mask, one[0].n = descs[i].dst, None, None, None, rows
                check(lib().dl_push_rows(src.data_ptr() + q * rows * row_bytes, row_bytes, 0, one, 1, stream_of(dev)), "push")
        px.barrier()

    def push_slice():
        arr = (ctypes.c_void_p * len(peers))(*[bases[q] + rank * rows * row_bytes for q in peers])
        check(lib().dl_push_slice(src.data_ptr() + rank * rows * row_bytes, arr, len(peers), rows * row_bytes,
                                  stream_of(dev)), "push_slice")
        px.barrier()

    def nccl_a2a():
        dist.all_to_all_single(dst, src)

    out = {}
    for name, fn in (("push", lambda: push(True)), ("push_contig", lambda: push(False)), ("push_slice", push_slice),
                     ("nccl_a2a", nccl_a2a)):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        reps = 4
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / reps], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        out[name] = {"ms": round(ms, 3), "GB/s_out_per_rank": round((world - 1) * rows * row_bytes / ms / 1e6, 1)}
    if rank == 0:
        print(json.dumps({"world": world, "rows_per_peer": rows, "row_bytes": row_bytes, **out}), flush=True)
    px.barrier()
    torch.cuda.synchronize()
    px.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
