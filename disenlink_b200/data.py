"""PyG-free readers for the data formats the training script feeds the hot path, plus the script's
edge split (host side).

    read_planetoid(raw_dir, name)   main_disentangled.py:117-123  (Planetoid raw pickles: cora, citeseer, pubmed)
    read_wikipedia_npz(path)        main_disentangled.py:73-79,98-101 (chameleon / squirrel / crocodile .npz)
    read_webkb(raw_dir)             main_disentangled.py:69-71,89-94  (texas / wisconsin / cornell, PyG WebKB raw files)
    read_fb100(mat_path)            main_disentangled.py:61-63,104-108 via other_hetero_datasets.py:131-154 (Facebook100 .mat)
    read_twitch(dir, lang)          main_disentangled.py:61-63,104-114 via load_data.py:21-69 (twitch-e csv / json)
    read_pyg_data(path)             main_disentangled.py:123-128      (mini/year*.pt: a pickled PyG Data object)
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


def read_webkb(raw_dir: str, to_undirected: bool = True):
    """texas / wisconsin / cornell as torch_geometric.datasets.WebKB reads them (external, version unpinned in
    the reference): `out1_node_feature_label.txt` (id TAB comma-separated features TAB label, one header line)
    and `out1_graph_edges.txt` (src TAB dst).  -> (x [N,F] float32, edge_index [2,E] int64, y [N] int64).
    PyG 2.0.x (the reference's vintage: cpython-39 byte code, 2022) symmetrises the edge list and coalesces it;
    later releases keep it directed -- `to_undirected=False` gives those columns (still coalesced, row-major).
    The hot path symmetrises the training edges itself (main_disentangled.py:141), so the choice only changes
    the number of edge columns the 85/10/5 split sees."""
    rows = open(os.path.join(raw_dir, "out1_node_feature_label.txt")).read().split("\n")[1:]
    rows = [r.split("\t") for r in rows if r.strip()]
    x = torch.tensor([[float(v) for v in r[1].split(",")] for r in rows], dtype=torch.float32)
    y = torch.tensor([int(r[2]) for r in rows], dtype=torch.int64)
    n = x.shape[0]
    e = np.loadtxt(os.path.join(raw_dir, "out1_graph_edges.txt"), skiprows=1, dtype=np.int64).reshape(-1, 2)
    key = e[:, 0] * n + e[:, 1]
    if to_undirected:
        key = np.concatenate([key, e[:, 1] * n + e[:, 0]])
    key = np.unique(key)
    return x, torch.from_numpy(np.stack([key // n, key % n])), y


def _one_hot_columns(col: np.ndarray) -> np.ndarray:
    """sklearn.preprocessing.label_binarize(col, classes=np.unique(col)) restated: one indicator column per
    distinct value -- except that TWO distinct values give a single column (1 = the larger value) and ONE
    distinct value gives a single all-zero column (sklearn's binary / degenerate cases)."""
    vals = np.unique(col)
    if vals.size <= 2:
        return (col == vals[-1]).astype(np.float64)[:, None] if vals.size == 2 else np.zeros((col.size, 1))
    return (col[:, None] == vals[None, :]).astype(np.float64)


def read_fb100(mat_path: str):
    """One Facebook100 school (`data/facebook100/<name>.mat`: adjacency `A`, `local_info` [N,7]) as the
    reference builds it (load_data.py:11-19, other_hetero_datasets.py:131-154): edge_index = A.nonzero() (scipy
    returns the entries of a CSC matrix in row-major order), label = gender - 1 (-1 = unlabeled), features = the other six attribute
    columns (student/faculty flag, major, second major, dorm, year, high school) one-hot encoded with
    label_binarize and concatenated.  -> (x [N,F] float32, edge_index [2,E] int64, y [N] int64)."""
    import scipy.io
    import scipy.sparse as sp
    mat = scipy.io.loadmat(mat_path)
    A = sp.coo_matrix(mat["A"])
    keep = A.data != 0
    n = A.shape[0]
    key = np.unique(A.row[keep].astype(np.int64) * n + A.col[keep].astype(np.int64))
    edge_index = torch.from_numpy(np.stack([key // n, key % n]))
    meta = np.asarray(mat["local_info"]).astype(np.int64)
    y = torch.from_numpy(meta[:, 1] - 1)
    cols = [meta[:, 0]] + [meta[:, c] for c in range(2, meta.shape[1])]
    x = np.hstack([_one_hot_columns(c) for c in cols])
    return torch.from_numpy(x.astype(np.float32)), edge_index, y


def read_twitch(lang_dir: str, lang: str, n_raw_features: int = 3170):
    """One twitch-e language graph as load_data.py:21-69 builds it: nodes and `mature` labels from
    `musae_<lang>_target.csv` (column 5 = node id, first occurrence wins), edges from `musae_<lang>_edges.csv`
    (duplicates collapse, row-major order = csr .nonzero()), bag-of-games features from
    `musae_<lang>_features.json` as 0/1 over 3170 ids with the all-zero columns removed; labels reordered so
    that label[i] belongs to node id i.  The training script appends the reversed columns itself
    (main_disentangled.py:112-114).  -> (x [N,F] float32, edge_index [2,E] int64, y [N] int64)."""
    import csv
    import json
    ids, lab, seen = [], [], set()
    with open(os.path.join(lang_dir, f"musae_{lang}_target.csv")) as f:
        rd = csv.reader(f)
        next(rd)
        for row in rd:
            nid = int(row[5])
            if nid not in seen:
                seen.add(nid)
                ids.append(nid)
                lab.append(int(row[2] == "True"))
    ids, lab = np.asarray(ids, np.int64), np.asarray(lab, np.int64)
    n = lab.size
    pos_of_id = np.empty(n, np.int64)
    pos_of_id[ids] = np.arange(n)                    # ids are a permutation of 0..n-1
    y = lab[pos_of_id]
    e = np.loadtxt(os.path.join(lang_dir, f"musae_{lang}_edges.csv"), delimiter=",", skiprows=1, dtype=np.int64).reshape(-1, 2)
    key = np.unique(e[:, 0] * n + e[:, 1])
    edge_index = torch.from_numpy(np.stack([key // n, key % n]))
    with open(os.path.join(lang_dir, f"musae_{lang}_features.json")) as f:
        feats = json.load(f)
    x = np.zeros((n, n_raw_features), np.float32)
    for node, fs in feats.items():
        if int(node) < n:
            x[int(node), np.asarray(fs, dtype=np.int64)] = 1.0
    x = x[:, x.sum(axis=0) != 0]
    return torch.from_numpy(x), edge_index, torch.from_numpy(y)


class _PygBag:
    """Stand-in for the torch_geometric classes inside a pickled Data object: keeps the pickled state."""

    def __setstate__(self, state):
        self.__dict__.update(state if isinstance(state, dict) else {"state": state})


class _PygPickle:
    """pickle_module for torch.load: torch_geometric.* classes resolve to _PygBag, torch's own tensor
    rebuilders and plain containers are allowed, everything else is refused (the file is foreign input)."""
    __name__ = "disenlink_b200.data._PygPickle"
    ALLOWED = {("collections", "OrderedDict"), ("torch._utils", "_rebuild_tensor_v2"), ("torch._utils", "_rebuild_parameter"),
               ("torch", "LongStorage"), ("torch", "FloatStorage"), ("torch", "IntStorage"), ("torch", "DoubleStorage"),
               ("torch", "BoolStorage"), ("torch", "ByteStorage"), ("torch", "HalfStorage"), ("torch", "ShortStorage"),
               ("torch", "CharStorage"), ("builtins", "dict"), ("builtins", "list"), ("builtins", "tuple"), ("builtins", "set")}

    class Unpickler(pickle.Unpickler):
        def find_class(self, module, name):
            if module.split(".")[0] == "torch_geometric":
                return _PygBag
            if (module, name) in _PygPickle.ALLOWED:
                return super().find_class(module, name)
            raise pickle.UnpicklingError(f"refusing to unpickle {module}.{name}")

    @staticmethod
    def load(f, **kw):
        return _PygPickle.Unpickler(f, **kw).load()


_torch_load = torch.load       # (the script harness swaps torch.load for a wrapper that lands here)


def read_pyg_data(path: str) -> dict:
    """A `torch.save`d torch_geometric Data object (mini/year<i>.pt, main_disentangled.py:123-128) without PyG:
    -> the dict of its attributes (x, edge_index, y, num_nodes, ...) as stored."""
    obj = _torch_load(path, map_location="cpu", pickle_module=_PygPickle, weights_only=False)
    store = obj.__dict__.get("_store", obj)
    mapping = store.__dict__.get("_mapping", store.__dict__)
    return {k: v for k, v in mapping.items() if not k.startswith("_")}


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
