"""Pubmed fixtures for BASELINE.json configs[2] ("Pubmed with larger factor count K=8, d=64").

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_pubmed.py

The dense reference cannot run Pubmed at K=8 ((9K+6) N^2 4 B = 121 GB at N = 19 717, SURVEY.md
section 6), and the reference ships Pubmed without its `allx` feature file, so two kinds of fixture
are committed:

  pubmed_graph.npz         the REAL Pubmed graph (ind.pubmed.graph parsed like PyG's Planetoid
                           reader: self loops dropped, symmetrised, coalesced; N = 19 717,
                           88 648 directed columns) -- the GPU path is compared with the oracle
                           on it at K=8, d=8 and d=64 (tests/test_gpu_scale_parity.py)
  pubmed_sub_K8_d64.npz    the dense reference ITSELF (model.py, unmodified) on the subgraph
  pubmed_sub_K8_d8.npz     induced by a breadth-first ball around the largest hub (1 000 / 3 000
                           nodes), K = 8, d = 64 / 8, in the standard fixture format of
                           make_golden.py (so every fixture test picks them up)
"""
import os
import pickle
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402

REF = mg.REF


def load_pubmed_graph():
    with open(os.path.join(REF, "data/pubmed/raw/ind.pubmed.graph"), "rb") as f:
        graph = pickle.load(f, encoding="latin1")
    rows, cols = [], []
    for a, nb in graph.items():
        for b in nb:
            if a != b:
                rows += [a, b]
                cols += [b, a]
    n = max(graph.keys()) + 1
    key = np.unique(np.asarray(rows, np.int64) * n + np.asarray(cols, np.int64))
    return n, key // n, key % n


def bfs_ball(n, src, dst, size):
    order = np.argsort(src, kind="stable")
    s, d = src[order], dst[order]
    ptr = np.searchsorted(s, np.arange(n + 1))
    deg = np.diff(ptr)
    seen = np.zeros(n, bool)
    start = int(np.argmax(deg))
    ball, frontier = [start], [start]
    seen[start] = True
    while frontier and len(ball) < size:
        nxt = []
        for a in frontier:
            for b in d[ptr[a]:ptr[a + 1]]:
                if not seen[b] and len(ball) < size:
                    seen[b] = True
                    ball.append(int(b))
                    nxt.append(int(b))
        frontier = nxt
    return np.sort(np.asarray(ball, np.int64))


def sub_case(name, n, src, dst, size, K, d, beta, seed):
    rng = np.random.default_rng(seed)
    nodes = bfs_ball(n, src, dst, size)
    remap = -np.ones(n, np.int64)
    remap[nodes] = np.arange(nodes.size)
    keep = (remap[src] >= 0) & (remap[dst] >= 0)
    s_all, d_all = remap[src[keep]], remap[dst[keep]]
    ns = nodes.size
    E = s_all.size
    perm = rng.permutation(E)
    n_tr = int(round(0.85 * E))
    tr, va = perm[:n_tr], perm[n_tr:]
    Z = (rng.standard_normal((ns, K, d)) * (0.9 / d ** 0.25)).astype(np.float32)
    m = 2
    negs = [mg.structured_negatives(s_all, d_all, ns, rng) for _ in range(m)]
    neg_tr = (np.concatenate([s_all[tr]] * m), np.concatenate([k[tr] for k in negs]))
    neg_va = (np.concatenate([s_all[va]] * m), np.concatenate([k[va] for k in negs]))
    val_u = np.concatenate([s_all[va], neg_va[0]])
    val_v = np.concatenate([d_all[va], neg_va[1]])
    pu = np.concatenate([val_u, rng.integers(0, ns, size=2000)])
    pv = np.concatenate([val_v, rng.integers(0, ns, size=2000)])
    mg.COMPACT = True
    mg.run_reference(name, s_all[tr], d_all[tr], ns, Z, beta=beta, T=1.0, pu=pu, pv=pv,
                     loss_pos=(s_all[tr], d_all[tr]), loss_neg=neg_tr, m=m,
                     extra=dict(n_val_pos=va.size, val_u=val_u, val_v=val_v))
    mg.COMPACT = False


def main():
    n, src, dst = load_pubmed_graph()
    assert n == 19717 and src.size == 88648, (n, src.size)
    np.savez_compressed(os.path.join(HERE, "pubmed_graph.npz"), N=n, src=src.astype(np.int32), dst=dst.astype(np.int32))
    print("pubmed_graph: N=%d directed columns=%d" % (n, src.size))
    sub_case("pubmed_sub_K8_d64", n, src, dst, 1000, K=8, d=64, beta=0.6, seed=5)
    sub_case("pubmed_sub_K8_d8", n, src, dst, 3000, K=8, d=8, beta=0.6, seed=6)


if __name__ == "__main__":
    main()
