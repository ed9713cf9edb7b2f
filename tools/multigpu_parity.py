"""Run under torchrun on N GPUs: the node-partitioned step (NCCL all-gathers between the kernels)
must reproduce the single-GPU step on the rows each rank owns: routing (kstar, w) bit for bit --
the canonical arithmetic does not depend on the decomposition -- and every float within 2e-6 of
the tensor's max-abs (a row that crosses a 2048-entry range boundary is summed piecewise, and the
cut positions depend on the partition; rows inside one range are bitwise equal).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/multigpu_parity.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from disenlink_b200.partition import PartitionedLinkStep  # noqa: E402


def main():
    world = int(os.environ["WORLD_SIZE"])
    rank = int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    N, E, K, d, P = 300_000, 3_000_000, 8, 16, 600_000
    src, dst = bench.gen_edges(N, E, 0, dev)
    g = torch.Generator(device=dev).manual_seed(5)
    u = torch.sort(torch.randint(0, N, (P,), generator=g, device=dev)).values
    v = torch.randint(0, N, (P,), generator=g, device=dev)
    lab = (torch.rand(P, generator=g, device=dev) < 0.2).float()
    wts = torch.full((P,), 1.0 / P, device=dev)
    part_step = PartitionedLinkStep(src, dst, N, u, v, lab, wts, K, d, 0.5, 1.0, world=world, rank=rank,
                                    device=dev)
    part = part_step.part
    Zfull = bench.gen_Z(part.n_pad, K, d, 0, dev)
    Zin = torch.zeros_like(Zfull)
    Zin[part.lo:part.hi] = Zfull[part.lo:part.hi]        # a rank only has its own rows before the gather
    pushed = part_step.register_input(Zin)
    if rank == 0:
        print('exchange:', 'NVLink push' if pushed else 'NCCL all-gather', flush=True)
    part_step.run(Zin)
    single = PartitionedLinkStep(src, dst, N, u, v, lab, wts, K, d, 0.5, 1.0, world=1, rank=0, device=dev)
    Zs = Zfull[:N].clone()
    single.run(Zs)
    torch.cuda.synchronize()
    lo, hi = part.lo, part.hi
    ok = True
    e0, e1 = int(single.graph.rowptr[lo]), int(single.graph.rowptr[hi])
    nl = part_step.graph.nnz
    ok &= (e1 - e0 == nl)
    ok &= bool(torch.equal(part_step.graph.col[:nl], single.graph.col[e0:e1]))
    ok &= bool(torch.equal(part_step.kstar[:nl], single.kstar[e0:e1]))
    ok &= bool(torch.equal(part_step.w[:nl], single.w[e0:e1]))
    if not ok:
        print(f"rank {rank}: integer / routing outputs differ")
    worst = 0.0
    for name in ("H", "s", "r", "dZ", "dH"):
        a, b = getattr(part_step, name)[lo:hi], getattr(single, name)[lo:hi]
        rel = float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
        worst = max(worst, rel)
        if rel > 2e-6:
            ok = False
            print(f"rank {rank}: {name} rel err {rel:.3e}")
    rel = float((part_step.prob[:P] - single.prob[:P]).abs().max())
    worst = max(worst, rel)
    ok &= rel < 2e-6
    ok &= abs(float(part_step.loss) - float(single.loss)) < 1e-6 * abs(float(single.loss))
    print(f"rank {rank}: routing bitwise equal, worst float rel err {worst:.2e}")
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"multigpu parity world={world}: {'OK' if int(flag.item()) else 'MISMATCH'} "
              f"(nnz_local={part_step.graph.nnz}, loss={float(single.loss):.6f})")
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
