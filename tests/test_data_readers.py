"""Host-side readers and the edge split against the numbers PyG's datasets give for
the files shipped with the reference.  The data live under /root/reference, which does not exist on
the GPU box: skipped there."""
import os

import numpy as np
import pytest
import torch

from disenlink_b200 import data

CORA = "/root/reference/data/cora/raw"
CHAM = "/root/reference/data_pre_false/chameleon/raw/chameleon.npz"


@pytest.mark.skipif(not os.path.exists(CORA), reason="reference data not present")
def test_planetoid_cora_matches_pyg_canonical_numbers():
    x, ei, y = data.read_planetoid(CORA, "cora")
    assert tuple(x.shape) == (2708, 1433) and tuple(ei.shape) == (2, 10556) and tuple(y.shape) == (2708,)
    assert int(x.sum()) == 49216 and int(y.max()) == 6                       # binary bag of words, 7 classes
    key = ei[0] * 2708 + ei[1]
    assert bool((key[1:] > key[:-1]).all())                                  # coalesced, row-major
    assert not bool((ei[0] == ei[1]).any())                                  # self-loops removed
    rev = torch.sort(ei[1] * 2708 + ei[0]).values
    assert torch.equal(rev, key)                                             # symmetric
    xs, _, _ = data.read_planetoid(CORA, "cora", sparse_x=True)
    assert xs._nnz() == 49216 and torch.equal(xs.to_dense(), x)


@pytest.mark.skipif(not os.path.exists(CHAM), reason="reference data not present")
def test_wikipedia_npz_chameleon():
    x, ei, y = data.read_wikipedia_npz(CHAM)
    assert tuple(x.shape) == (2277, 128) and tuple(y.shape) == (2277,)
    # the reference's reader keeps the stored columns as they are (dataset.py:119-124): 72202 columns,
    # 9410 of them duplicates, 100 self loops, not symmetrised
    assert ei.shape[1] == 72202
    raw = np.load(CHAM, allow_pickle=True)["edges"]
    assert np.array_equal(ei.numpy(), raw.T)
    key = ei[0] * 2277 + ei[1]
    assert key.numel() - torch.unique(key).numel() == 9410 and int((ei[0] == ei[1]).sum()) == 100
    _, eic, _ = data.read_wikipedia_npz(CHAM, coalesce=True)
    assert eic.shape[1] == 62792 and torch.equal(eic[0] * 2277 + eic[1], torch.unique(key))
    xs = data.row_standardize(x)
    assert float(xs.mean(dim=1).abs().max()) < 1e-5
    assert float((xs.std(dim=1) - 1).abs().max()) < 1e-4


def test_split_sizes_match_sklearn():
    from sklearn.model_selection import train_test_split
    for n in (10556, 62792, 17, 100):
        a, b = train_test_split(range(n), train_size=0.85)
        c, d = train_test_split(b, train_size=2 / 3)
        tr, te, va = data.split_edges(n, seed=3)
        assert (len(tr), len(te), len(va)) == (len(a), len(c), len(d))
        allidx = torch.cat([tr, te, va]).numpy()
        assert np.array_equal(np.sort(allidx), np.arange(n))
    assert torch.equal(data.split_edges(1000, 5)[0], data.split_edges(1000, 5)[0])


def test_batched_projection_equals_per_factor_mlps():
    """Disentangle.project (one GEMM + one batched GEMM) == the K per-factor MLPs of model.py:106,
    dense and sparse x, both Factor (nhid == 1) and Factor2."""
    from disenlink_b200.model import Disentangle
    torch.manual_seed(0)
    x = torch.randn(40, 30)
    x[x.abs() < 1.2] = 0
    for nhid in (1, 16):
        m = Disentangle(30, nhid, 8, nfactor=4, beta=0.5, t=1)
        ref = torch.stack([f(x) for f in m.factors], dim=1)
        assert float((m.project(x) - ref).abs().max()) < 1e-6
        assert float((m.project(x.to_sparse()) - ref).abs().max()) < 1e-6
        m.project(x).square().sum().backward()
        g1 = [p.grad.clone() for p in m.parameters()]
        m.zero_grad()
        ref = torch.stack([f(x) for f in m.factors], dim=1)
        ref.square().sum().backward()
        for a, p in zip(g1, m.parameters()):
            assert float((a - p.grad).abs().max()) <= 1e-5 * max(float(p.grad.abs().max()), 1e-12)


# ----------------------------------------------------------------------------------------------
# the other datasets of hyperparameters_setting whose raw files ship with the reference, read from the staged
# byte-for-byte copies in baseline/_ref (tools/stage_reference.sh); skipped where nothing is staged
# ----------------------------------------------------------------------------------------------
from conftest import ROOT  # noqa: E402

REF = os.path.join(ROOT, "baseline", "_ref")
staged = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "load_data.py")), reason="baseline/_ref not staged")


@pytest.fixture(scope="module")
def ref_load_data():
    """The reference's own load_data.py, imported as is.  It spells the integer dtype `np.int` (removed from
    numpy 1.24): the alias is restored for the duration of the import and the calls, nothing is edited."""
    import importlib.util
    had = hasattr(np, "int")
    if not had:
        np.int = int
    cwd = os.getcwd()
    os.chdir(REF)                                   # its paths are relative ('data/...')
    spec = importlib.util.spec_from_file_location("_ref_load_data", os.path.join(REF, "load_data.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    yield m
    os.chdir(cwd)
    if not had:
        del np.int


@staged
@pytest.mark.parametrize("lang", ["PTBR", "DE"])
def test_twitch_reader_equals_the_reference_loader(ref_load_data, lang):
    A, label, features = ref_load_data.load_twitch(lang)                     # load_data.py:21-69
    ei_ref = np.stack(A.nonzero())                                           # other_hetero_datasets.py:120
    x, ei, y = data.read_twitch(os.path.join(REF, "data", "twitch", lang), lang)
    assert np.array_equal(ei.numpy(), ei_ref)
    assert np.array_equal(y.numpy(), label)
    assert x.dtype == torch.float32 and np.array_equal(x.numpy(), features.astype(np.float32))


@staged
@pytest.mark.parametrize("name,n,f,e", [("Amherst41", 2235, 1193, 181908), ("JohnsHopkins55", 5180, 2406, 373172),
                                        ("Reed98", 962, 745, 37624)])
def test_fb100_reader_equals_the_reference_construction(ref_load_data, name, n, f, e):
    """other_hetero_datasets.py:131-154 on top of load_data.load_fb100, with the same sklearn call."""
    from sklearn.preprocessing import label_binarize
    A, meta = ref_load_data.load_fb100(name)
    ei_ref = np.stack(A.nonzero())
    meta = meta.astype(np.int64)
    vals = np.hstack((np.expand_dims(meta[:, 0], 1), meta[:, 2:]))
    feats = np.hstack([label_binarize(vals[:, c], classes=np.unique(vals[:, c])) for c in range(vals.shape[1])])
    x, ei, y = data.read_fb100(os.path.join(REF, "data", "facebook100", name + ".mat"))
    assert tuple(x.shape) == (n, f) and ei.shape[1] == e                      # the LINKX paper's sizes
    assert np.array_equal(ei.numpy(), ei_ref)
    assert np.array_equal(x.numpy(), feats.astype(np.float32))
    assert np.array_equal(y.numpy(), meta[:, 1] - 1)


def test_one_hot_columns_follow_label_binarize_in_the_degenerate_cases():
    from sklearn.preprocessing import label_binarize
    for col in (np.array([3, 3, 3]), np.array([0, 7, 7, 0]), np.array([5, 1, 9, 1]), np.array([2, 0, 1, 2, 0])):
        assert np.array_equal(data._one_hot_columns(col), label_binarize(col, classes=np.unique(col)))


@staged
@pytest.mark.parametrize("name,n,e_dir", [("texas", 183, 325), ("wisconsin", 251, 515), ("cornell", 183, 298)])
def test_webkb_reader(name, n, e_dir):
    raw = os.path.join(REF, "data", name, "raw")
    x, ei, y = data.read_webkb(raw)
    assert tuple(x.shape) == (n, 1703) and int(y.max()) == 4 and tuple(y.shape) == (n,)
    key = ei[0] * n + ei[1]
    assert bool((key[1:] > key[:-1]).all())                                  # coalesced, row-major
    assert torch.equal(torch.sort(ei[1] * n + ei[0]).values, key)            # symmetrised (PyG 2.0.x)
    _, eid, _ = data.read_webkb(raw, to_undirected=False)
    assert eid.shape[1] == e_dir                                             # PyG's documented directed counts
    lines = [ln.split("\t") for ln in open(os.path.join(raw, "out1_graph_edges.txt")).read().split("\n")[1:] if ln.strip()]
    want = np.unique(np.array([[int(a), int(b)] for a, b in lines]), axis=0)
    assert np.array_equal(eid.numpy().T, want)


@staged
def test_pyg_pickle_reader_and_its_allow_list(tmp_path):
    import pickle
    import zipfile
    path = os.path.join(REF, "mini", "year9.pt")
    d = data.read_pyg_data(path)
    assert set(d) >= {"x", "edge_index", "num_nodes"}
    n = int(d["num_nodes"])
    assert tuple(d["x"].shape) == (n, 128) and d["edge_index"].shape[0] == 2 and int(d["edge_index"].max()) < n
    z = zipfile.ZipFile(path)                                                # the raw storages, without any unpickling
    ei_raw = np.frombuffer(z.read("archive/data/0"), dtype=np.int64)
    x_raw = np.frombuffer(z.read("archive/data/1"), dtype=np.float32)
    assert np.array_equal(d["edge_index"].numpy().reshape(-1), ei_raw)
    assert np.array_equal(d["x"].numpy().reshape(-1), x_raw)
    # a pickle that reaches for anything outside torch's tensor rebuilders is refused, not executed
    assert data._PygPickle.load(_dump(tmp_path / "p.pkl", {"a": [1, 2]})) == {"a": [1, 2]}     # plain containers: allowed
    with pytest.raises(pickle.UnpicklingError):
        data._PygPickle.load(_dump(tmp_path / "q.pkl", os.getcwd))            # a global outside the allow-list


def _dump(path, obj):
    import pickle
    with open(path, "wb") as f:
        pickle.dump(obj, f)
    return open(path, "rb")


@staged
@pytest.mark.parametrize("argv", [
    ("--dataset", "texas", "--beta", "0.6", "--nfactor", "5", "--nhidden", "512", "--nembed", "32"),      # hyperparameters_setting:5
    ("--dataset", "fb100", "--sub_dataset", "Reed98", "--beta", "0.5", "--nfactor", "5", "--nhidden", "256", "--nembed", "32"),
    ("--dataset", "year", "--miniid", "9", "--beta", "0.7", "--nfactor", "2", "--nhidden", "64", "--nembed", "16", "--epochs", "1", "--m", "1"),
])
def test_unmodified_script_runs_over_the_readers_with_the_reference_model(argv):
    """The staged main_disentangled.py, unedited, with ITS OWN model.py on the host, fed by these readers through
    the stand-ins of tools/run_reference_script.py: the data path is the one the CUDA module gets on the GPU box
    (tests/test_gpu_script.py); this pins it where the reference itself can run."""
    import re
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "run_reference_script.py"), "--model", "reference",
                          "--", "--epochs", "2", "--run", "1", *argv], capture_output=True, text=True, timeout=900,
                         env={**os.environ, "CUDA_VISIBLE_DEVICES": ""})
    assert out.returncode == 0, out.stderr[-2000:]
    epochs = re.findall(r"epoch: (\d+) loss: ([0-9.eE+-]+) val_auc: ([0-9.eE+-]+)", out.stdout)
    assert len(epochs) == (int(argv[argv.index("--epochs") + 1]) if "--epochs" in argv else 2), out.stdout[-2000:]
    assert all(float(e[1]) == float(e[1]) for e in epochs)
    assert 0.3 < float(re.search(r"test auc: ([0-9.eE+-]+)", out.stdout).group(1)) <= 1.0


@staged
def test_planetoid_citeseer_matches_pyg_canonical_numbers():
    """CiteSeer has 15 isolated test nodes missing from tx / ty: PyG inserts zero rows for them."""
    x, ei, y = data.read_planetoid(os.path.join(REF, "data", "citeseer", "raw"), "citeseer")
    assert tuple(x.shape) == (3327, 3703) and tuple(ei.shape) == (2, 9104) and int(y.max()) == 5
    assert int((x.sum(1) == 0).sum()) == 15
    key = ei[0] * 3327 + ei[1]
    assert bool((key[1:] > key[:-1]).all()) and not bool((ei[0] == ei[1]).any())
    assert torch.equal(torch.sort(ei[1] * 3327 + ei[0]).values, key)
