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
