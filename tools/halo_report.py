#!/usr/bin/env python
"""Halo volume of the node-partitioned run on the real graphs the reference ships, for the nnz-balanced split of
the given numbering against partition.locality_partition (host-side integer work; no GPU needed):

    python tools/halo_report.py > profiles/r02_locality_partition.json

Per graph and world size: halo rows summed over the ranks (= rows of Z pushed per step), the share of CSR entries
whose column lives on another rank, and max / mean entries per rank."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from disenlink_b200 import data as D  # noqa: E402
from disenlink_b200.partition import NodePartition, locality_partition  # noqa: E402

REF = os.path.join(ROOT, "baseline", "_ref")


def halo_stats(src, dst, n, bounds):
    s_, d_ = src.numpy(), dst.numpy()
    key = np.unique(np.concatenate([s_ * n + d_, d_ * n + s_]))
    rows, cols = key // n, key % n
    owner = np.searchsorted(np.asarray(bounds[1:]), np.arange(n), side="right")
    remote = owner[rows] != owner[cols]
    halo = int(np.unique(owner[rows][remote].astype(np.int64) * n + cols[remote]).size)
    load = np.bincount(owner[rows], minlength=len(bounds) - 1)
    return {"halo_rows": halo, "remote_entry_share": round(float(remote.mean()), 4),
            "max_over_mean_entries": round(float(load.max() / load.mean()), 3)}


def graphs():
    p = os.path.join(ROOT, "tests", "golden", "pubmed_graph.npz")
    if os.path.exists(p):
        g = np.load(p)
        yield "pubmed", torch.from_numpy(g["src"].astype(np.int64)), torch.from_numpy(g["dst"].astype(np.int64)), int(g["N"])
    for name in ("cora", "citeseer"):
        raw = os.path.join(REF, "data", name, "raw")
        if os.path.exists(raw):
            x, ei, _ = D.read_planetoid(raw, name)
            yield name, ei[0], ei[1], x.shape[0]
    p = os.path.join(REF, "data_pre_false", "chameleon", "raw", "chameleon.npz")
    if os.path.exists(p):
        x, ei, _ = D.read_wikipedia_npz(p)
        yield "chameleon", ei[0], ei[1], x.shape[0]
    p = os.path.join(REF, "data", "squirrel", "geom_gcn", "raw", "out1_graph_edges.txt")
    if os.path.exists(p):
        e = torch.from_numpy(np.loadtxt(p, skiprows=1, dtype=np.int64).T.copy())
        yield "squirrel", e[0], e[1], int(e.max()) + 1
    for name in ("Amherst41", "JohnsHopkins55"):
        p = os.path.join(REF, "data", "facebook100", name + ".mat")
        if os.path.exists(p):
            x, ei, _ = D.read_fb100(p)
            yield "fb100-" + name, ei[0], ei[1], x.shape[0]
    p = os.path.join(REF, "data", "twitch", "DE")
    if os.path.exists(p):
        x, ei, _ = D.read_twitch(p, "DE")
        yield "twitch-DE", ei[0], ei[1], x.shape[0]


def main():
    out = {}
    for name, src, dst, n in graphs():
        rec = {"N": n, "edge_columns": int(src.numel())}
        for world in (2, 4, 8):
            base = NodePartition.nnz_balanced(src, dst, n, world, 0).bounds
            t0 = time.perf_counter()
            order, bounds = locality_partition(src, dst, n, world)
            dt = time.perf_counter() - t0
            rec[f"world{world}"] = {"given_numbering": halo_stats(src, dst, n, base),
                                    "locality_partition": halo_stats(order.relabel(src), order.relabel(dst), n, bounds),
                                    "partition_s": round(dt, 3)}
        out[name] = rec
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
