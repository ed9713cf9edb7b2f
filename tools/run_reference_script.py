#!/usr/bin/env python
"""Run the reference's UNMODIFIED main_disentangled.py with `from model import Disentangle` resolving
to this repository's implementation (or, with --model reference, to the reference's own model.py).

    python tools/run_reference_script.py [--model ours|reference] -- --dataset cora --epochs 3 --run 1 ...

The script (staged byte for byte into baseline/_ref/ by tools/stage_reference.sh) imports
torch_geometric, torch_sparse-backed loaders and friends at module level (main_disentangled.py:2-17);
none of them is installed here and none is on the hot path.  This harness registers minimal stand-ins
in sys.modules for exactly what the datasets with shipped raw files touch (cora, citeseer, chameleon,
texas / wisconsin / cornell, twitch-e, fb100, year):

    torch_geometric.datasets.Planetoid       -> disenlink_b200.data.read_planetoid on data/cora/raw
    torch_geometric.transforms (T.Compose, T.NormalizeFeatures: built at :56, never applied)
    torch_geometric.utils.structured_negative_sampling -> ops.structured_negative_sampling on the GPU
                                                (numpy restatement of PyG's loop on the CPU)
    dataset.WikipediaNetwork                 -> disenlink_b200.data.read_wikipedia_npz (dataset.py:119-124:
                                                duplicates kept, no to_undirected)
    torch_geometric.datasets.WebKB           -> disenlink_b200.data.read_webkb (texas, wisconsin, cornell)
    other_hetero_datasets.load_nc_dataset    -> read_twitch / read_fb100 behind the NCDataset surface the script
                                                reads (.graph['edge_index' | 'node_feat' | 'num_nodes'], .label;
                                                other_hetero_datasets.py:78-154); other names raise
    torch.load('mini/year<i>.pt')            -> read_pyg_data (a pickled PyG Data object; main_disentangled.py:124)

and then executes the script with runpy, cwd = baseline/_ref, so `./data/` and `data_pre_false/` resolve
to the staged files.  Nothing in the script is edited.
"""
import argparse
import os
import runpy
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


class _Data:
    def __init__(self, x, edge_index, y):
        self.x, self.edge_index, self.y = x, edge_index, y

    def to(self, device):
        return _Data(self.x.to(device), self.edge_index.to(device), self.y.to(device))


class _Dataset:
    def __init__(self, data):
        self._data = data

    def __getitem__(self, i):
        assert i == 0
        return self._data


def install_stubs(use_ours: bool):
    import numpy as np
    import torch
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from disenlink_b200 import data as dl_data

    def planetoid(root, name, transform=None):
        x, ei, y = dl_data.read_planetoid(os.path.join(root, name.lower(), "raw"), name.lower())
        return _Dataset(_Data(x, ei, y))

    def wikipedia(root, name, geom_gcn_preprocess=True, **kw):
        path = os.path.join(root, name, "raw", name + ".npz")
        if not os.path.exists(path):
            # main_disentangled.py:77 also builds the geom-gcn copy, whose only use is the unused y (:80)
            path = os.path.join("data_pre_false", name, "raw", name + ".npz")
        x, ei, y = dl_data.read_wikipedia_npz(path, coalesce=False)
        return _Dataset(_Data(x, ei, y))

    def webkb(root, name, **kw):
        x, ei, y = dl_data.read_webkb(os.path.join(root, name, "raw"))
        return _Dataset(_Data(x, ei, y))

    class _NC:                                                  # NCDataset surface (other_hetero_datasets.py:22-76)
        def __init__(self, x, ei, y):
            self.graph = {"edge_index": ei, "edge_feat": None, "node_feat": x, "num_nodes": int(x.shape[0])}
            self.label = y

    def load_nc_dataset(dataname, sub_dataname=""):
        if dataname == "twitch-e":
            if sub_dataname not in ("DE", "ENGB", "ES", "FR", "PTBR", "RU", "TW"):
                sub_dataname = "DE"                             # other_hetero_datasets.py:84-86
            return _NC(*dl_data.read_twitch(os.path.join("data", "twitch", sub_dataname), sub_dataname))
        if dataname == "fb100":
            if sub_dataname not in ("Penn94", "Amherst41", "Cornell5", "JohnsHopkins55", "Reed98"):
                sub_dataname = "Penn94"                         # other_hetero_datasets.py:89-91
            return _NC(*dl_data.read_fb100(os.path.join("data", "facebook100", sub_dataname + ".mat")))
        raise RuntimeError(f"{dataname}: raw files are not shipped with the reference (not available in this harness)")

    torch_load = torch.load

    def load_pt(f, *a, **kw):
        if isinstance(f, str) and f.startswith("mini/year"):
            d = dl_data.read_pyg_data(f)
            return _Data(d["x"], d["edge_index"], d.get("y", torch.zeros(d["x"].shape[0], dtype=torch.int64)))
        return torch_load(f, *a, **kw)
    torch.load = load_pt

    def structured_negative_sampling(edge_index, num_nodes=None, contains_neg_self_loops=True):
        if edge_index.is_cuda:
            from disenlink_b200 import ops
            structured_negative_sampling.calls += 1
            return ops.structured_negative_sampling(edge_index, num_nodes, seed=structured_negative_sampling.calls)
        # PyG's loop restated for the CPU reference arm
        n = int(edge_index.max()) + 1 if num_nodes is None else int(num_nodes)
        row, col = edge_index[0].numpy(), edge_index[1].numpy()
        pos = np.unique(row.astype(np.int64) * n + col)
        rng = np.random.default_rng(structured_negative_sampling.calls)
        structured_negative_sampling.calls += 1
        k = rng.integers(0, n, row.size)
        bad = np.isin(row.astype(np.int64) * n + k, pos)
        while bad.any():
            k[bad] = rng.integers(0, n, int(bad.sum()))
            bad = np.isin(row.astype(np.int64) * n + k, pos)
        return edge_index[0], edge_index[1], torch.from_numpy(k)
    structured_negative_sampling.calls = 0

    def _unused(*a, **k):
        raise RuntimeError("not available in this harness (its raw files are not shipped with the reference)")

    tg = types.ModuleType("torch_geometric")
    tg.datasets = types.ModuleType("torch_geometric.datasets")
    tg.datasets.Planetoid, tg.datasets.WebKB, tg.datasets.Amazon = planetoid, webkb, _unused
    tg.transforms = types.ModuleType("torch_geometric.transforms")
    tg.transforms.Compose = lambda ts: ts
    tg.transforms.NormalizeFeatures = lambda: None
    tg.utils = types.ModuleType("torch_geometric.utils")
    tg.utils.structured_negative_sampling = structured_negative_sampling
    tg.utils.to_dense_adj = tg.utils.homophily = tg.utils.degree = _unused
    ds = types.ModuleType("dataset")
    ds.WikipediaNetwork = wikipedia
    oh = types.ModuleType("other_hetero_datasets")
    oh.load_nc_dataset = load_nc_dataset
    for m in (tg, tg.datasets, tg.transforms, tg.utils, ds, oh):
        sys.modules[m.__name__] = m
    if use_ours:
        import integration.model as ours                       # the 3-line shim of INTEGRATION.md
        sys.modules["model"] = ours
    else:
        sys.path.insert(0, REF_DIR)                            # the reference's own model.py


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="ours", choices=["ours", "reference"])
    args, rest = ap.parse_known_args()
    rest = [a for a in rest if a != "--"]
    script = os.path.join(REF_DIR, "main_disentangled.py")
    if not os.path.exists(script):
        raise SystemExit("baseline/_ref/main_disentangled.py is missing: run tools/stage_reference.sh where /root/reference exists")
    install_stubs(args.model == "ours")
    os.chdir(REF_DIR)
    sys.argv = [script] + rest
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
