import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_names(prefix=""):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, prefix + "*.npz")))


def load_golden(name):
    """Fixture arrays; the big fixtures store ids as int32 and alias the loss positives to the
    training edges (see make_golden.py COMPACT) -- both undone here."""
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    for k, v in list(g.items()):
        if v.dtype == np.int32:
            g[k] = v.astype(np.int64)
    if "loss_neg_u" in g and "loss_pos_u" not in g:
        g["loss_pos_u"], g["loss_pos_v"] = g["src"], g["dst"]
    return g


# module_*: whole-module fixtures; *_graph: an edge list only (no reference outputs)
GRAPH_FIXTURES = [n for n in golden_names() if not n.startswith("module_") and not n.endswith("_graph")]


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o
    o.lib()
    return o
