"""Run under torchrun on N GPUs: the node-partitioned step (rank-local storage, halo pushes over NVLink
peer memory) must reproduce the single-GPU step on the rows each rank owns: integers and routing
(col, kstar, w) bit for bit -- the canonical arithmetic does not depend on the decomposition -- and every
float within 1e-5 of the tensor's max-abs (a row that crosses a 2048-entry range boundary is summed
piecewise, and the cut positions depend on the partition; rows inside one range are bitwise equal).
Two steps with Z changed in between (exchange ordering).  bench.py --gpus N emits the same block as
"parity" in its JSON line.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/multigpu_parity.py [workload]
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    world = int(os.environ["WORLD_SIZE"])
    rank = int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    out = bench.multi_gpu_parity(world, rank, dev, sys.argv[1] if len(sys.argv) > 1 else "tiny")
    if rank == 0:
        print(json.dumps(out), flush=True)
        print(f"multigpu parity world={world}: {'OK' if out['ok'] else 'MISMATCH'}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if out["ok"] else 1)


if __name__ == "__main__":
    main()
