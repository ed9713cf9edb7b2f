"""End-to-end training on the scalable path (examples/train_link.py = the loop of
main_disentangled.py:131-221): batched projection -> CSR graph -> device negative sampling -> fused
attention / aggregation / decoder / BCE forward+backward -> Adam -> device AUC."""
import argparse
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples"))


def test_training_learns_planted_partition():
    import train_link
    torch.manual_seed(0)
    args = argparse.Namespace(nfactor=4, nhidden=64, nembed=16, beta=0.7, temperature=1, m=2, lr=0.01,
                              epochs=40, seed=0, log_every=1000, standardize=True)
    x, edge_index, _ = train_link.synthetic(n=2000, seed=0)
    out = train_link.run(args, x, edge_index, torch.device("cuda:0"), log=lambda *_: None)
    losses = [l for l, _ in out["history"]]
    assert losses[-1] < 0.8 * losses[0]
    assert out["best_val_auc"] > 0.75 and out["test_auc"] > 0.75
    # same seed, same run: the whole loop is deterministic
    torch.manual_seed(0)
    out2 = train_link.run(args, x, edge_index, torch.device("cuda:0"), log=lambda *_: None)
    assert out2["history"][:5] == out["history"][:5]
