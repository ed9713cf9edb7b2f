"""PyG-free readers for the data formats the training script feeds the hot path, plus the script's
edge split (host side).

    read_planetoid(raw_dir, name)   main_disentangled.py:117-123  (Planetoid raw pickles: cora, citeseer, pubmed)
    read_wikipedia_npz(path)        main_disentangled.py:73-79,98-101 (chameleon / squirrel / crocodile .npz)
    row_standardize(x)              main_disentangled.py:100  (x - mean_row) / std_row
    split_edges(E, seed, device)    main_disentangled.py:134-135  (85 / 10 / 5 train / test / val over edge columns)

The readers return torch tensors on the host (features, `edge_index [2,E] int64`); nothing here touches the
GPU library.  Planetoid semantics follow torch_geometric.io.read_planetoid_data (external, version
unpinned in the reference): x = vstack(allx, tx) with the test rows put back at their ids, edges from
the adjacency dict, self-loops removed, symmetrised, coalesced (row-major order).
"""
from __future__ import annotations

import os
import pickle

import numpy as np
import torch


def _load_pickle(path):
    with open(path, "rb") as f:
        return pickle.load(f, encoding="latin1")


def read_planetoid(raw_dir: str, name: str, sparse_x: bool = False):
    """-> (x [N,F] float32 dense or sparse-COO, edge_index [2,E] int64, y [N] int64).
    Cora gives N=2708, E=10556, F=1433 like PyG's Planetoid dataset."""
    part = {k: _load_pickle(os.path.join(raw_dir, f"ind.{name}.{k}")) for k in ("x", "tx", "allx", "y", "ty", "ally", "graph")}
    test_index = np.loadtxt(os.path.join(raw_dir, f"ind.{name}.test.index"), dtype=np.int64)
    sorted_test = np.sort(test_index)
    import scipy.sparse as sp
    tx, ty = part["tx"].tolil(), np.asarray(part["ty"])
    if name.lower() == "citeseer":
        # isolated test nodes are missing from tx / ty: insert zero rows (PyG does the same)
        n_full = int(sorted_test[-1] - sorted_test[0] + 1)
        tx_ext = sp.lil_matrix((n_full, tx.shape[1]), dtype=tx.dtype)
        tx_ext[sorted_test - sorted_test[0], :] = tx
        ty_ext = np.zeros((n_full, ty.shape[1]), dtype=ty.dtype)
        ty_ext[sorted_test - sorted_test[0], :] = ty
        tx, ty = tx_ext, ty_ext
    x = sp.vstack([part["allx"], tx]).tolil()
    y = np.vstack([np.asarray(part["ally"]), ty]).argmax(axis=1)
    x[test_index, :] = x[sorted_test, :]
    y[test_index] = y[sorted_test]
    n = x.shape[0]
    rows, cols = [], []
    for src, nbrs in part["graph"].items():
        rows.extend([src] * len(nbrs))
        cols.extend(nbrs)
    rows, cols = np.asarray(rows, np.int64), np.asarray(cols, np.int64)
    keep = rows != cols
    rows, cols = rows[keep], cols[keep]
    key = np.unique(np.concatenate([rows * n + cols, cols * n + rows]))
    edge_index = torch.from_numpy(np.stack([key // n, key % n]))
    xc = x.tocoo()
    if sparse_x:
        xt = torch.sparse_coo_tensor(np.stack([xc.row, xc.col]), xc.data.astype(np.float32), (n, x.shape[1])).coalesce()
    else:
        xt = torch.from_numpy(np.asarray(x.todense(), dtype=np.float32))
    return xt, edge_index, torch.from_numpy(y.astype(np.int64))


def read_wikipedia_npz(path: str, coalesce: bool = False):
    """-> (x [N,F] float32, edge_index [2,E] int64, y [N] int64).
    Default = the reference's own reader (dataset.py:119-124): `edges.t()` exactly as stored -- directed
    columns, duplicates KEPT, no to_undirected (commented out at :123).  chameleon: N=2277, F=128,
    E=72202 columns of which 9410 are duplicates and 100 self loops.  The duplicates matter downstream:
    they change the 85/10/5 split sizes (main_disentangled.py:134) and which positives the `== 1`
    loss mask drops (:175-178,195).  coalesce=True gives the sorted de-duplicated columns (62792) that
    PyG's coalescing loaders would."""
    d = np.load(path, allow_pickle=True)
    x = torch.from_numpy(np.asarray(d["features"], dtype=np.float32))
    e = np.asarray(d["edges"], dtype=np.int64)
    n = x.shape[0]
    if coalesce:
        key = np.unique(e[:, 0] * n + e[:, 1])
        edge_index = torch.from_numpy(np.stack([key // n, key % n]))
    else:
        edge_index = torch.from_numpy(np.ascontiguousarray(e.T))
    y = torch.from_numpy(np.asarray(d["label"] if "label" in d.files else d["target"]).astype(np.int64))
    return x, edge_index, y


def row_standardize(x: torch.Tensor) -> torch.Tensor:
    """(x - mean over features) / unbiased std over features, per node (main_disentangled.py:100)."""
    return (x - x.mean(dim=1, keepdim=True)) / x.std(dim=1, keepdim=True)


def split_edges(n_edges: int, seed: int = 0, device="cpu"):
    """The script's split of the edge columns: train_test_split(range(E), train_size=0.85), then
    train_test_split(rest, train_size=2/3) -> (train, test, val) index tensors with sklearn's sizes
    (floor(train_size * n) for the first part).  The permutation is torch's (seeded); sklearn's own
    unseeded shuffle cannot be reproduced."""
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    perm = torch.randperm(int(n_edges), generator=g, device=device)
    n_train = int(np.floor(0.85 * n_edges))
    rest = n_edges - n_train
    n_test = int(np.floor(rest * (2.0 / 3.0)))
    return perm[:n_train], perm[n_train:n_train + n_test], perm[n_train + n_test:]
