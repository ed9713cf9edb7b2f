"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every CUDA entry point is called through
the C ABI (ctypes binding in disenlink_b200/_lib.py) and compared with the CPU oracle (oracle/),
which is itself pinned to the dense reference by tests/test_oracle_golden.py, and directly with
the committed golden vectors of the reference.

Bars (BASELINE.json north_star): integer / index outputs bit-exact; routing (kstar, w) bit-exact
against the oracle because both use the same canonical fp32 order; everything continuous within
1e-5 relative (of the tensor's max-abs) of the reference, and within 2e-6 of the oracle.
"""
import numpy as np
import pytest
import torch

from conftest import GRAPH_FIXTURES, load_golden

pytestmark = pytest.mark.gpu

REF_TOL = 1e-5     # vs the dense reference (golden vectors)
ORA_TOL = 2e-6     # vs the CPU oracle (same algorithm, different summation order in places)
DEV = "cuda:0"


def relerr(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    if b.size == 0:
        return 0.0
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def t(x, dtype=None):
    return torch.as_tensor(np.ascontiguousarray(x), device=DEV, dtype=dtype)


@pytest.fixture(scope="module")
def dl():
    import disenlink_b200.ops as ops
    from disenlink_b200.graph import Graph
    from disenlink_b200 import _lib
    _lib.lib()   # must load: no fallback
    return ops, Graph


def random_graph(rng, n, e, hubs=()):
    src = rng.integers(0, n, size=e)
    dst = rng.integers(0, n, size=e)
    for node, deg in hubs:
        src = np.concatenate([src, np.full(deg, node)])
        dst = np.concatenate([dst, rng.choice(n, size=deg, replace=False)])
    src = np.concatenate([src, src[:e // 20], np.arange(0, n, 97)])   # duplicates, self loops
    dst = np.concatenate([dst, dst[:e // 20], np.arange(0, n, 97)])
    return src.astype(np.int64), dst.astype(np.int64)


# ------------------------------------------------------------------------------------------
# integer work: bit-exact
# ------------------------------------------------------------------------------------------
def check_graph(oracle, Graph, src, dst, n):
    g = Graph.from_edges(t(src), t(dst), n)
    rowptr, col = oracle.csr_from_edges(src, dst, n)
    assert g.nnz == col.size
    assert np.array_equal(g.rowptr.cpu().numpy(), rowptr)
    assert np.array_equal(g.col.cpu().numpy(), col)
    perm, off = oracle.degree_buckets(rowptr)
    assert np.array_equal(g.bucket_off_host.numpy(), off)
    assert np.array_equal(g.perm.cpu().numpy()[:n], perm)
    if col.size:
        assert np.array_equal(g.rev_index().cpu().numpy(), oracle.rev_index(rowptr, col))
    deg = np.diff(rowptr)
    hub_rows = perm[:g.n_hub]
    assert np.all(deg[hub_rows] >= 512) and (g.n_hub == int((deg >= 512).sum()))
    nseg = (deg[hub_rows] + 511) // 512
    assert np.array_equal(g.hub_seg_ptr.cpu().numpy(), np.concatenate([[0], np.cumsum(nseg)]))
    if g.n_hub_items:
        assert np.array_equal(g.item_hub.cpu().numpy()[:g.n_hub_items],
                              np.repeat(np.arange(g.n_hub), nseg))
    return g, rowptr, col


@pytest.mark.parametrize("name", GRAPH_FIXTURES)
def test_csr_build_fixture(dl, oracle, name):
    _, Graph = dl
    gd = load_golden(name)
    g, rowptr, col = check_graph(oracle, Graph, gd["src"], gd["dst"], int(gd["N"]))
    assert np.array_equal(g.rows().cpu().numpy(), gd["ref_rows"])
    assert np.array_equal(g.col.cpu().numpy().astype(np.int64), gd["ref_cols"])


@pytest.mark.parametrize("n,e,hubs", [(1, 0, ()), (5, 3, ()), (3000, 20000, ((7, 2000), (11, 600), (13, 512))),
                                      (200000, 1500000, ((5, 70000),))])
def test_csr_build_random(dl, oracle, n, e, hubs):
    _, Graph = dl
    rng = np.random.default_rng(n + e)
    if e == 0:
        src = dst = np.zeros(0, np.int64)
    else:
        src, dst = random_graph(rng, n, e, hubs)
    check_graph(oracle, Graph, src, dst, n)


def test_csr_build_rejects_out_of_range(dl):
    _, Graph = dl
    from disenlink_b200._lib import DlError
    with pytest.raises(DlError):
        Graph.from_edges(t(np.array([0, 5])), t(np.array([1, 2])), 4)


def test_csr_from_dense_and_asymmetry(dl, oracle):
    _, Graph = dl
    from disenlink_b200._lib import DlError
    rng = np.random.default_rng(5)
    n = 77
    a = (rng.random((n, n)) < 0.1).astype(np.float32)
    a = np.maximum(a, a.T)
    g = Graph.from_dense(t(a))
    r, c = np.nonzero(a)
    assert np.array_equal(g.rows().cpu().numpy(), r)
    assert np.array_equal(g.col.cpu().numpy().astype(np.int64), c)
    g.assert_symmetric()
    a[3, 9], a[9, 3] = 1.0, 0.0
    with pytest.raises(DlError):
        Graph.from_dense(t(a)).assert_symmetric()


# ------------------------------------------------------------------------------------------
# forward / backward against the oracle and the golden vectors
# ------------------------------------------------------------------------------------------
def run_all(ops, Graph, oracle, src, dst, n, Z, beta, T, pu, pv, dS=None, Gin=None):
    """GPU and oracle results for every kernel on one input; returns dict of (gpu, oracle)."""
    g = Graph.from_edges(t(src), t(dst), n)
    rowptr, col = oracle.csr_from_edges(src, dst, n)
    Zt = t(Z)
    kstar, w, s = ops.edge_attn_fwd(g, Zt, T)
    H = ops.factor_spmm_fwd(g, Zt, kstar, w, s, beta)
    o_k, o_w, o_s = oracle.edge_attn_fwd(rowptr, col, Z, T)
    o_H = oracle.factor_spmm_fwd(rowptr, col, Z, o_k, o_w, o_s, beta)
    batch = ops.PairBatch(t(pu), t(pv), n)
    logit, prob = ops.pair_score_fwd(Zt, H, batch, T)
    o_logit, o_prob = oracle.pair_score_fwd(pu, pv, Z, o_H, T)
    rng = np.random.default_rng(0)
    if dS is None:
        dS = (rng.standard_normal(pu.size) / max(pu.size, 1)).astype(np.float32)
    dZp, dHp = ops.pair_score_bwd(Zt, H, batch, t(dS), T)
    o_dZp, o_dHp = oracle.pair_score_bwd(pu, pv, Z, o_H, dS, T)
    if Gin is None:
        Gin = rng.standard_normal(Z.shape).astype(np.float32)
    dZ0 = rng.standard_normal(Z.shape).astype(np.float32)
    dZ, r = ops.factor_bwd(g, Zt, t(Gin), kstar, w, s, beta, T, dZ=t(dZ0).clone())
    o_dZ, o_r = oracle.factor_bwd(rowptr, col, Z, Gin, o_k, o_w, o_s, beta, T, dZ_init=dZ0,
                                  return_r=True)
    c = lambda x: x.cpu().numpy()
    return dict(kstar=(c(kstar), o_k), w=(c(w), o_w), s=(c(s), o_s), H=(c(H), o_H),
                logit=(c(logit), o_logit), prob=(c(prob), o_prob), dZp=(c(dZp), o_dZp),
                dHp=(c(dHp), o_dHp), r=(c(r), o_r), dZ=(c(dZ), o_dZ)), g


def assert_matches_oracle(res, tol=ORA_TOL):
    k, ok = res["kstar"]
    assert np.array_equal(k, ok), f"{(k != ok).sum()} routing mismatches vs oracle"
    w, ow = res["w"]
    assert np.array_equal(w.view(np.uint32), ow.view(np.uint32)), "w not bit-identical to oracle"
    for name in ("s", "H", "logit", "prob", "dZp", "dHp", "r", "dZ"):
        a, b = res[name]
        e = relerr(a, b)
        assert e < (tol if name in ("s", "H", "logit", "prob") else 5 * tol), f"{name}: rel err {e:.3e}"


@pytest.mark.parametrize("name", GRAPH_FIXTURES)
def test_fixture_vs_oracle_and_reference(dl, oracle, name):
    ops, Graph = dl
    gd = load_golden(name)
    n, K, d = int(gd["N"]), int(gd["K"]), int(gd["d"])
    beta, T = float(gd["beta"]), float(gd["T"])
    res, g = run_all(ops, Graph, oracle, gd["src"], gd["dst"], n, gd["Z"], beta, T, gd["pu"], gd["pv"])
    assert_matches_oracle(res)
    # against the dense reference itself
    k = res["kstar"][0]
    bad = np.nonzero(k != gd["ref_kstar"])[0]
    assert np.all(gd["ref_margin"][bad] < 1e-5)
    if bad.size == 0:
        assert relerr(res["w"][0], gd["ref_w"]) < REF_TOL
        assert relerr(res["s"][0], gd["ref_s"]) < REF_TOL
        assert relerr(res["H"][0].reshape(n, -1), gd["ref_H"]) < REF_TOL
        assert relerr(res["prob"][0], gd["ref_prob"]) < REF_TOL


@pytest.mark.parametrize("name", [n for n in GRAPH_FIXTURES if "ref_dZ" in load_golden(n)])
def test_fixture_training_loss_and_grad(dl, oracle, name):
    """main_disentangled.py:195 on pair lists through the autograd bindings vs the reference's
    loss and dL/dZ."""
    ops, Graph = dl
    import torch.nn.functional as F
    gd = load_golden(name)
    n, beta, T, m = int(gd["N"]), float(gd["beta"]), float(gd["T"]), float(gd["m"])
    g = Graph.from_edges(t(gd["src"]), t(gd["dst"]), n)
    pu, pv = oracle.pairs_exactly_once(gd["loss_pos_u"], gd["loss_pos_v"], n)
    nu, nv = oracle.pairs_exactly_once(gd["loss_neg_u"], gd["loss_neg_v"], n)
    pos, neg = ops.PairBatch(t(pu), t(pv), n), ops.PairBatch(t(nu), t(nv), n)
    Z = t(gd["Z"]).requires_grad_(True)
    H = ops.factor_aggregate(Z, g, beta, T)
    pp, pn = ops.pair_score(Z, H, pos, T), ops.pair_score(Z, H, neg, T)
    loss = F.binary_cross_entropy(pp, torch.ones_like(pp)) + F.binary_cross_entropy(pn, torch.zeros_like(pn)) / m
    loss.backward()
    assert abs(loss.item() - float(gd["ref_loss"])) <= REF_TOL * abs(float(gd["ref_loss"]))
    assert relerr(Z.grad.cpu().numpy(), gd["ref_dZ"]) < 5 * REF_TOL


SHAPES = [(8, 16), (8, 8), (8, 64), (5, 32), (5, 64), (3, 32), (10, 32), (10, 64), (20, 32), (4, 32),
          (3, 8), (2, 8), (3, 4), (2, 4), (1, 16),        # fast-path instantiations
          (4, 6), (2, 12), (7, 20), (3, 1), (2, 130), (6, 24)]  # runtime-generic path


@pytest.mark.parametrize("K,d", SHAPES)
def test_shapes_with_hub_rows(dl, oracle, K, d):
    ops, Graph = dl
    rng = np.random.default_rng(K * 1000 + d)
    n = 1500
    src, dst = random_graph(rng, n, 6000, hubs=((3, 1300), (40, 520)))
    Z = (rng.standard_normal((n, K, d)) * (0.7 / np.sqrt(d))).astype(np.float32)
    pu = np.concatenate([np.repeat(rng.integers(0, n, 300), 6), np.full(1200, 9)])  # 9: pair hub
    pv = rng.integers(0, n, pu.size)
    res, g = run_all(ops, Graph, oracle, src, dst, n, Z, 0.6, 1.0 if d != 8 else 2.0, pu, pv)
    assert g.n_hub >= 2
    assert_matches_oracle(res)


def test_empty_and_isolated(dl, oracle):
    ops, Graph = dl
    rng = np.random.default_rng(3)
    n, K, d = 10, 8, 16
    Z = rng.standard_normal((n, K, d)).astype(np.float32)
    g = Graph.from_edges(t(np.zeros(0, np.int64)), t(np.zeros(0, np.int64)), n)
    kstar, w, s = ops.edge_attn_fwd(g, t(Z), 1.0)
    H = ops.factor_spmm_fwd(g, t(Z), kstar, w, s, 0.7)
    assert torch.all(s == 1)
    assert relerr(H.cpu().numpy(), np.float32(0.7) * Z) < 1e-7
    empty = ops.PairBatch(t(np.zeros(0, np.int64)), t(np.zeros(0, np.int64)), n)
    lg, pb = ops.pair_score_fwd(t(Z), H, empty, 1.0)
    assert lg.numel() == 0 and pb.numel() == 0
    dZ, dH = ops.pair_score_bwd(t(Z), H, empty, t(np.zeros(0, np.float32)), 1.0)
    assert float(dZ.abs().max()) == 0.0 and float(dH.abs().max()) == 0.0


def test_run_to_run_bitwise_determinism(dl):
    ops, Graph = dl
    rng = np.random.default_rng(9)
    n, K, d = 20000, 8, 16
    src, dst = random_graph(rng, n, 150000, hubs=((1, 5000),))
    Z = t((rng.standard_normal((n, K, d)) * 0.2).astype(np.float32))
    G = t(rng.standard_normal((n, K, d)).astype(np.float32))
    g = Graph.from_edges(t(src), t(dst), n)
    batch = ops.PairBatch(t(rng.integers(0, n, 50000)), t(rng.integers(0, n, 50000)), n)
    dS = t(rng.standard_normal(50000).astype(np.float32))
    outs = []
    for _ in range(2):
        kstar, w, s = ops.edge_attn_fwd(g, Z, 1.0)
        H = ops.factor_spmm_fwd(g, Z, kstar, w, s, 0.5)
        dZ, r = ops.factor_bwd(g, Z, G, kstar, w, s, 0.5, 1.0)
        lg, pb = ops.pair_score_fwd(Z, H, batch, 1.0)
        dZp, dHp = ops.pair_score_bwd(Z, H, batch, dS, 1.0)
        outs.append([x.clone() for x in (kstar, w, s, H, dZ, r, lg, pb, dZp, dHp)])
    for a, b in zip(*outs):
        assert torch.equal(a, b)
    # routing is bitwise symmetric
    rev = g.rev_index()
    assert torch.equal(outs[0][0], outs[0][0][rev])
    assert torch.equal(outs[0][1], outs[0][1][rev])


@pytest.mark.parametrize("K,d", [(8, 16), (5, 32), (3, 7)])
def test_saved_normaliser_sj(dl, K, d):
    """dl_factor_spmm_fwd's optional sj_out is exactly s[col, kstar] per entry (streamed and
    fallback paths), and handing it to the backward changes nothing, bit for bit."""
    ops, Graph = dl
    rng = np.random.default_rng(K * 100 + d)
    n = 6000
    src, dst = random_graph(rng, n, 50000, hubs=((2, 3000),))
    Z = t((rng.standard_normal((n, K, d)) * 0.3).astype(np.float32))
    G = t(rng.standard_normal((n, K, d)).astype(np.float32))
    g = Graph.from_edges(t(src), t(dst), n)
    kstar, w, s = ops.edge_attn_fwd(g, Z, 1.0)
    sj = torch.full((g.nnz,), float("nan"), dtype=torch.float32, device=DEV)
    H1 = ops.factor_spmm_fwd(g, Z, kstar, w, s, 0.5, sj=sj)
    H0 = ops.factor_spmm_fwd(g, Z, kstar, w, s, 0.5)
    assert torch.equal(H0, H1)
    assert torch.equal(sj, s[g.col.long(), kstar.long()])
    dZ0, r0 = ops.factor_bwd(g, Z, G, kstar, w, s, 0.5, 1.0)
    dZ1, r1 = ops.factor_bwd(g, Z, G, kstar, w, s, 0.5, 1.0, sj=sj)
    assert torch.equal(dZ0, dZ1) and torch.equal(r0, r1)
    # pre-scaled aggregation (slices divided by s beforehand): same H up to the rounding of Z/s
    H2 = ops.factor_spmm_fwd(g, Z, kstar, w, s, 0.5, zs=torch.empty_like(Z))
    assert relerr(H2.cpu().numpy(), H0.cpu().numpy()) < ORA_TOL
    with pytest.raises(ValueError):
        ops.factor_spmm_fwd(g, Z, kstar, w, s, 0.5, sj=sj, zs=torch.empty_like(Z))


def test_large_graph_properties(dl):
    """Size-independent properties at a size the dense reference cannot touch (N = 1M,
    nnz ~ 2*10^7): attention columns sum to one, kstar/w symmetric, scores symmetric in (u,v),
    isolated rows give H = beta * Z, gradient of a linear functional matches a finite difference."""
    ops, Graph = dl
    gen = torch.Generator(device=DEV).manual_seed(0)
    n, K, d, E = 1_000_000, 8, 16, 10_000_000
    u = (torch.rand(E, device=DEV, generator=gen) ** 3 * n).long().clamp_(max=n - 1)
    v = (torch.rand(E, device=DEV, generator=gen) ** 3 * n).long().clamp_(max=n - 1)
    u = (u * 7919 + 13) % n
    v = (v * 7919 + 13) % n
    g = Graph.from_edges(u, v, n)
    assert g.n_hub > 0
    Z = torch.randn(n, K, d, device=DEV, generator=gen) * 0.25
    kstar, w, s = ops.edge_attn_fwd(g, Z, 1.0)
    rev = g.rev_index()
    assert torch.equal(kstar, kstar[rev]) and torch.equal(w, w[rev])
    # column sums of att_k are 1 wherever a node has an entry routed to k
    att = w / s[g.col.long(), kstar.long()]
    colsum = torch.zeros(n * K, device=DEV, dtype=torch.float64)
    colsum.index_add_(0, g.col.long() * K + kstar.long(), att.double())
    routed = torch.zeros(n * K, device=DEV, dtype=torch.bool)
    routed[g.col.long() * K + kstar.long()] = True
    assert float((colsum[routed] - 1).abs().max()) < 1e-4
    assert float(colsum[~routed].abs().max()) == 0.0
    H = ops.factor_spmm_fwd(g, Z, kstar, w, s, 0.5)
    iso = (g.degrees() == 0)
    if iso.any():
        assert torch.equal(H[iso], 0.5 * Z[iso])
    P = 2_000_000
    pu = torch.randint(0, n, (P,), device=DEV, generator=gen)
    pv = torch.randint(0, n, (P,), device=DEV, generator=gen)
    a = ops.pair_score_fwd(Z, H, ops.PairBatch(pu, pv, n), 1.0)[0]
    b = ops.pair_score_fwd(Z, H, ops.PairBatch(pv, pu, n), 1.0)[0]
    assert torch.equal(a, b)
    assert torch.isfinite(H).all() and torch.isfinite(a).all()


# ------------------------------------------------------------------------------------------
# symmetric attention (every undirected edge once) and the asymmetric-adjacency guard
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("K,d", [(8, 16), (5, 32), (8, 8), (3, 32), (10, 32)])
def test_symmetric_attention_equals_two_sided(dl, oracle, K, d):
    """dl_edge_attn_fwd_sym: kstar / w bit-identical to the two-sided kernel and to the oracle, s
    within the summation-order tolerance; shapes without a factor-per-lane kernel fall back."""
    ops, Graph = dl
    from disenlink_b200 import _lib
    rng = np.random.default_rng(K * 7 + d)
    n = 30000
    src, dst = random_graph(rng, n, 200000, hubs=((3, 9000), (77, 600)))
    Z = (rng.standard_normal((n, K, d)) * (0.8 / np.sqrt(np.sqrt(d)))).astype(np.float32)
    g = Graph.from_edges(t(src), t(dst), n)
    g.sym_min_nnz = 0
    upper, eidx = g.sym_view()
    rowptr, col = oracle.csr_from_edges(src, dst, n)
    rows = np.repeat(np.arange(n), np.diff(rowptr))
    deg = np.diff(rowptr)
    # primary entry of an edge: the one whose row has the larger degree (ties: smaller id), diagonal included
    up = (rows == col) | (deg[rows] > deg[col]) | ((deg[rows] == deg[col]) & (rows < col))
    assert upper.nnz == int(up.sum())
    assert np.array_equal(upper.col.cpu().numpy(), col[up])
    assert np.array_equal(upper.rowptr.cpu().numpy(), np.concatenate([[0], np.cumsum(np.bincount(rows[up], minlength=n))]))
    # eidx: own position for primary entries, ~(the mirror's position) for secondary ones
    pos = np.cumsum(up) - up                                   # exclusive count of primaries
    key = rows.astype(np.int64) * n + col
    mirror = np.searchsorted(key, col.astype(np.int64) * n + rows)
    assert np.array_equal(key[mirror], col.astype(np.int64) * n + rows)
    exp = np.where(up, pos, ~pos[mirror])
    assert np.array_equal(eidx.cpu().numpy(), exp)
    lower, lmirror = g.sym_lower_view()
    assert np.array_equal(lower.col.cpu().numpy(), col[~up])
    assert np.array_equal(lmirror.cpu().numpy(), pos[mirror][~up])
    assert np.array_equal(lower.rowptr.cpu().numpy(), rowptr - upper.rowptr.cpu().numpy())
    k1, w1, s1 = (x.clone() for x in ops.edge_attn_fwd(g, t(Z), 1.0))
    g.flags = _lib.DL_F_NO_SYM
    k0, w0, s0 = ops.edge_attn_fwd(g, t(Z), 1.0)
    assert torch.equal(k0, k1) and torch.equal(w0, w1)
    assert relerr(s1.cpu().numpy(), s0.cpu().numpy()) < ORA_TOL
    o_k, o_w, o_s = oracle.edge_attn_fwd(rowptr, col, Z, 1.0)
    assert np.array_equal(k1.cpu().numpy(), o_k)
    assert np.array_equal(w1.cpu().numpy().view(np.uint32), o_w.view(np.uint32))
    assert relerr(s1.cpu().numpy(), o_s) < ORA_TOL


def test_asymmetric_adjacency_forward_ok_backward_refuses(dl, oracle):
    """Disentangle.forward(x, adj) accepts any adj; the gather-only backward needs a symmetric
    pattern and must say so instead of returning wrong gradients."""
    ops, Graph = dl
    from disenlink_b200._lib import DlError
    rng = np.random.default_rng(2)
    n, K, d = 300, 3, 8
    a = (rng.random((n, n)) < 0.03).astype(np.float32)          # NOT symmetrised
    g = Graph.from_dense(t(a))
    g.sym_min_nnz = 0
    assert g.sym_view() is None
    Z = (rng.standard_normal((n, K, d)) * 0.5).astype(np.float32)
    r, c = np.nonzero(a)
    rowptr = np.concatenate([[0], np.cumsum(np.bincount(r, minlength=n))]).astype(np.int64)
    o_H, o_k, o_w, o_s = oracle.factor_fwd(rowptr, c.astype(np.int32), Z, 0.7, 1.0)
    Zt = t(Z).requires_grad_(True)
    H, kstar, w, s = ops.factor_aggregate(Zt, g, 0.7, 1.0, return_attention=True)
    assert np.array_equal(kstar.cpu().numpy(), o_k)
    assert relerr(H.detach().cpu().numpy(), o_H) < ORA_TOL
    with pytest.raises(DlError, match="symmetric"):
        H.sum().backward()


def test_backward_pass2_single_writer_invariant(oracle):
    """The fire-and-forget vector reduction of backward pass 2 (red.global.add.v4.f32) is deterministic
    only because every row has ONE direct writer per launch.  A -DDL_DEBUG_SINGLE_WRITER build of bwd_fl.cu
    (tools/build_variant.sh dbgsw bwd_fl.cu -DDL_DEBUG_SINGLE_WRITER) counts the direct writers of every
    row and fails the call with DL_EINTERNAL on a second one; here it runs over a hub-heavy graph."""
    import ctypes
    import os
    from disenlink_b200 import _lib
    from disenlink_b200.graph import Graph
    path = os.path.join(os.path.dirname(_lib.LIB_PATH), "_variants", "lib_dbgsw.so")
    if not os.path.exists(path):
        pytest.skip("debug variant not built")
    dbg = ctypes.CDLL(path)
    dbg.dl_factor_bwd.restype, dbg.dl_factor_bwd.argtypes = _lib.SIGNATURES["dl_factor_bwd"]
    import disenlink_b200.ops as ops
    rng = np.random.default_rng(12)
    n, K, d = 60000, 8, 16
    src, dst = random_graph(rng, n, 900000, hubs=((1, 30000), (2, 2049), (3, 2048), (4, 4097)))
    g = Graph.from_edges(t(src), t(dst), n)
    Z = t((rng.standard_normal((n, K, d)) * 0.3).astype(np.float32))
    G = t(rng.standard_normal((n, K, d)).astype(np.float32))
    kstar, w, s = ops.edge_attn_fwd(g, Z, 1.0)
    g.flags = _lib.DL_F_NO_SYM                       # the two-sided pass 2, the path dl_factor_bwd takes
    dZ_ref, r_ref = ops.factor_bwd(g, Z, G, kstar, w, s, 0.5, 1.0)
    dZ, r = torch.zeros_like(Z), torch.empty_like(s)
    dev = torch.device(DEV)
    rc = dbg.dl_factor_bwd(g.ref, Z.data_ptr(), G.data_ptr(), kstar.data_ptr(), w.data_ptr(), s.data_ptr(), None,
                           ops._sr_scratch(g, s).data_ptr(), int(s.shape[0]), ops._x_scratch(g).data_ptr(), K, d,
                           0.5, 0.5, 1.0, dZ.data_ptr(), r.data_ptr(), g.hub_scratch(K * d).data_ptr(),
                           _lib.stream_of(dev))
    assert rc == 0, f"debug build reported rc={rc} (-6 = a row had two direct writers)"
    assert torch.equal(dZ, dZ_ref) and torch.equal(r, r_ref)


@pytest.mark.parametrize("K,d", [(8, 16), (8, 8), (5, 16), (5, 32), (3, 32), (10, 32)])
def test_symmetric_backward_pass2_equals_two_sided(dl, oracle, K, d):
    """dl_factor_bwd_edges_sym (every undirected edge evaluated once: coefficients computed on the primary
    entries, read back by the secondary ones) against the two-sided pass 2 and the oracle; (10, 32) has no
    factor-per-lane kernel and must take the regular path."""
    ops, Graph = dl
    from disenlink_b200 import _lib
    rng = np.random.default_rng(K * 31 + d)
    n = 40000
    src, dst = random_graph(rng, n, 300000, hubs=((3, 9000), (77, 2049), (5, 600)))
    Z = (rng.standard_normal((n, K, d)) * (0.8 / np.sqrt(np.sqrt(d)))).astype(np.float32)
    G = rng.standard_normal((n, K, d)).astype(np.float32)
    dZ0 = rng.standard_normal((n, K, d)).astype(np.float32)
    g = Graph.from_edges(t(src), t(dst), n)
    g.sym_min_nnz = 0
    kstar, w, s = ops.edge_attn_fwd(g, t(Z), 1.0)
    plan = ops.bwd_plan(g, K, d)
    assert plan["mode"] == ("sym" if (K, d) != (10, 32) else "x")
    dZ1, r1 = ops.factor_bwd(g, t(Z), t(G), kstar, w, s, 0.5, 1.0, dZ=t(dZ0).clone())
    dZ1b, _ = ops.factor_bwd(g, t(Z), t(G), kstar, w, s, 0.5, 1.0, dZ=t(dZ0).clone())
    assert torch.equal(dZ1, dZ1b)                                   # run-to-run bitwise
    g.flags = _lib.DL_F_NO_SYM
    assert ops.bwd_plan(g, K, d)["mode"] == "x"
    dZ2, r2 = ops.factor_bwd(g, t(Z), t(G), kstar, w, s, 0.5, 1.0, dZ=t(dZ0).clone())
    assert torch.equal(r1, r2)
    assert relerr(dZ1.cpu().numpy(), dZ2.cpu().numpy()) < 5 * ORA_TOL
    rowptr, col = oracle.csr_from_edges(src, dst, n)
    o_k, o_w, o_s = oracle.edge_attn_fwd(rowptr, col, Z, 1.0)
    o_dZ = oracle.factor_bwd(rowptr, col, Z, G, o_k, o_w, o_s, 0.5, 1.0, dZ_init=dZ0)
    assert relerr(dZ1.cpu().numpy(), o_dZ) < 5 * ORA_TOL
