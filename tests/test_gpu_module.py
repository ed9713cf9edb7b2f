"""GPU tests of the drop-in module API (disenlink_b200.model) against the reference's own outputs
(golden vectors produced by /root/reference/model.py): same constructor, state_dict and forward
contract as model.py:91-114, dense [N,N] link_pred for a dense adj, LinkScorer for a Graph."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-5


def relerr(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def dense_adj_sym(src, dst, n):
    adj = torch.zeros(n, n)
    adj[torch.as_tensor(src), torch.as_tensor(dst)] = 1
    adj = adj + adj.t()
    adj[adj != 0] = 1
    return adj


def build(g):
    from disenlink_b200.model import Disentangle
    m = Disentangle(int(g["F"]), int(g["nhid"]), int(g["d"]), nfactor=int(g["K"]), beta=float(g["beta"]), t=1)
    m.load_state_dict({k[3:]: torch.from_numpy(g[k]) for k in g if k.startswith("sd.")}, strict=True)
    return m.to(DEV)


@pytest.mark.parametrize("fixture", ["module_small", "module_nhid1"])
def test_forward_matches_reference_dense_contract(fixture):
    g = load_golden(fixture)
    n = int(g["N"])
    m = build(g)
    x = torch.from_numpy(g["x"]).to(DEV)
    adj = dense_adj_sym(g["src"], g["dst"], n).to(DEV)
    H, link_pred = m(x, adj)
    assert tuple(H.shape) == (n, int(g["K"]) * int(g["d"])) and tuple(link_pred.shape) == (n, n)
    assert relerr(H.detach().cpu().numpy(), g["ref_H"]) < TOL
    assert relerr(link_pred.detach().cpu().numpy(), g["ref_link_pred"]) < TOL
    # second call hits the cached graph and is bitwise identical
    H2, lp2 = m(x, adj)
    assert torch.equal(H, H2) and torch.equal(link_pred, lp2)


def test_training_step_gradients_match_reference():
    g = load_golden("module_small")
    n = int(g["N"])
    m = build(g)
    x = torch.from_numpy(g["x"]).to(DEV)
    adj = dense_adj_sym(g["src"], g["dst"], n).to(DEV)
    pu, pv = torch.from_numpy(g["pu"]).to(DEV), torch.from_numpy(g["pv"]).to(DEV)
    lab = torch.from_numpy(g["lab"]).to(DEV)
    # (a) the unmodified script's way: index the dense link_pred
    H, link_pred = m(x, adj)
    loss = F.binary_cross_entropy(link_pred[pu, pv], lab)
    loss.backward()
    assert abs(loss.item() - float(g["ref_loss"])) < TOL * abs(float(g["ref_loss"]))
    grads_dense = {k: p.grad.detach().cpu().numpy().copy() for k, p in m.named_parameters()}
    for k, gr in grads_dense.items():
        assert relerr(gr, g["grad." + k]) < 5 * TOL, k
    # (b) the scalable way: Graph handle + LinkScorer on the pair list
    from disenlink_b200.graph import Graph
    m.zero_grad()
    graph = Graph.from_edges(torch.from_numpy(g["src"]).to(DEV), torch.from_numpy(g["dst"]).to(DEV), n)
    H2, scorer = m(x, graph)
    loss2 = F.binary_cross_entropy(scorer(torch.stack([pu, pv])), lab)
    loss2.backward()
    assert abs(loss2.item() - float(g["ref_loss"])) < TOL * abs(float(g["ref_loss"]))
    for k, p in m.named_parameters():
        assert relerr(p.grad.cpu().numpy(), g["grad." + k]) < 5 * TOL, k
    # boolean-mask indexing of the scorer follows the script's row-major order
    mask = torch.zeros(n, n, dtype=torch.bool, device=DEV)
    mask[pu, pv] = True
    assert relerr(scorer[mask].detach().cpu().numpy(), link_pred[mask].detach().cpu().numpy()) < TOL


def test_disentangle_layer_returns_reference_triplet():
    """Disentangle_layer.forward -> (h_list, alpha0 [K,N,N], att list), model.py:49-77."""
    from disenlink_b200.model import Disentangle_layer
    g = load_golden("tiny_generic")
    n, K, d = int(g["N"]), int(g["K"]), int(g["d"])
    Z = torch.from_numpy(g["Z"]).to(DEV)
    adj = dense_adj_sym(g["src"], g["dst"], n).to(DEV)
    layer = Disentangle_layer(K, float(g["beta"]), t=float(g["T"]))
    h_list, alpha0, att = layer([Z[:, k, :] for k in range(K)], adj)
    H = torch.cat(h_list, dim=1).cpu().numpy()
    assert relerr(H, g["ref_H"]) < TOL
    rows, cols = g["ref_rows"], g["ref_cols"]
    att_st = torch.stack(att, 0).cpu().numpy()
    assert relerr(att_st[g["ref_kstar"], rows, cols], g["ref_att"]) < TOL
    off = att_st.copy()
    off[g["ref_kstar"], rows, cols] = 0
    assert np.abs(off).max() == 0.0
    Zc = g["Z"].astype(np.float64)
    want = np.exp(np.einsum("ikd,jkd->kij", Zc, Zc) / float(g["T"]))
    assert relerr(alpha0.cpu().numpy(), want) < TOL


def test_link_bce_loss_matches_script_semantics():
    """ops.link_bce_loss == main_disentangled.py:195 on the chameleon fixture (pairs exactly once,
    mean BCE, negatives / m), loss and dL/dZ."""
    from disenlink_b200 import ops
    from disenlink_b200.graph import Graph
    g = load_golden("chameleon_K5_d32")
    n, beta, T, m = int(g["N"]), float(g["beta"]), float(g["T"]), float(g["m"])
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    graph = Graph.from_edges(t(g["src"]), t(g["dst"]), n)
    pu, pv = ops.pairs_exactly_once(t(g["loss_pos_u"]), t(g["loss_pos_v"]), n)
    nu, nv = ops.pairs_exactly_once(t(g["loss_neg_u"]), t(g["loss_neg_v"]), n)
    batch = ops.PairBatch(torch.cat([pu, nu]), torch.cat([pv, nv]), n)
    lab = torch.cat([torch.ones(pu.numel()), torch.zeros(nu.numel())]).to(DEV)
    wts = torch.cat([torch.full((pu.numel(),), 1.0 / pu.numel()),
                     torch.full((nu.numel(),), 1.0 / (m * nu.numel()))]).to(DEV)
    Z = t(g["Z"]).requires_grad_(True)
    loss, prob, H = ops.link_bce_loss(Z, graph, batch, lab, wts, beta, T)
    loss.backward()
    assert abs(loss.item() - float(g["ref_loss"])) < TOL * abs(float(g["ref_loss"]))
    assert relerr(Z.grad.cpu().numpy(), g["ref_dZ"]) < 5 * TOL
    assert relerr(H.cpu().numpy().reshape(n, -1), g["ref_H"]) < TOL


@pytest.mark.parametrize("P", [0, 1, 1000, 1 << 20])
def test_link_bce_kernel_vs_oracle_and_torch(P):
    """dl_link_bce == oracle.bce_weighted == torch's own F.binary_cross_entropy + autograd through a
    sigmoid, including saturated scores (p == 0 or 1 exactly: clamped log, zero gradient)."""
    from disenlink_b200 import ops
    from oracle import oracle
    rng = np.random.default_rng(P)
    logit = (rng.standard_normal(P) * 6).astype(np.float32)
    if P >= 1000:
        logit[:8] = [200, -200, 120, -120, 17, -17, 0, 40]          # exact 1 / 0 after sigmoid in fp32
    y = (rng.random(P) < 0.3).astype(np.float32)
    w = rng.random(P).astype(np.float32) / max(P, 1)
    S = torch.from_numpy(logit).to(DEV).requires_grad_(True)
    prob = torch.sigmoid(S)
    yt, wt = torch.from_numpy(y).to(DEV), torch.from_numpy(w).to(DEV)
    loss, dS = ops.link_bce(prob.detach(), yt, wt)
    want_loss, want_dS = oracle.bce_weighted(prob.detach().cpu().numpy(), y, w)
    assert abs(loss.item() - want_loss) <= 1e-6 * max(abs(want_loss), 1e-30) + 1e-12
    if P:
        assert np.abs(dS.cpu().numpy() - want_dS).max() <= 1e-6 * max(np.abs(want_dS).max(), 1e-30)
        ref = (F.binary_cross_entropy(prob, yt, reduction="none") * wt).sum()
        ref.backward()
        assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item())
        assert np.abs(dS.cpu().numpy() - S.grad.cpu().numpy()).max() <= 1e-6 * float(S.grad.abs().max())
        # unweighted variant (weights = NULL) and loss-only variant
        l1, none = ops.link_bce(prob.detach(), yt, None, want_grad=False)
        assert none is None
        assert abs(l1.item() - oracle.bce_weighted(prob.detach().cpu().numpy(), y)[0]) <= 1e-6 * abs(l1.item())
        # bitwise run-to-run determinism of the reduction
        l2, _ = ops.link_bce(prob.detach(), yt, wt)
        assert l2.item() == loss.item()


@pytest.mark.parametrize("case,P", [("distinct", 1000), ("ties", 100000), ("saturated", 1 << 20),
                                    ("signed_zero", 5000), ("one_tie_group", 70000), ("distinct", 1 << 22)])
def test_roc_auc_kernel_vs_oracle_and_sklearn(case, P):
    """dl_roc_auc: the integer 2U is bit-exact against the oracle, the AUC matches sklearn's
    roc_auc_score to 1e-12 (the script's metric, main_disentangled.py:204,219)."""
    from disenlink_b200 import ops
    from oracle import oracle
    from sklearn.metrics import roc_auc_score
    rng = np.random.default_rng(P)
    y = (rng.random(P) < 0.17).astype(np.float32)
    if case == "distinct":
        sc = rng.standard_normal(P).astype(np.float32)
    elif case == "ties":
        sc = (rng.integers(0, 211, P) / 211.0).astype(np.float32)
    elif case == "saturated":
        sc = (1.0 / (1.0 + np.exp(-(rng.standard_normal(P) * 30)))).astype(np.float32)
    elif case == "signed_zero":
        sc = np.where(rng.random(P) < 0.5, np.float32(0.0), np.float32(-0.0)).astype(np.float32)
        sc[:100] = rng.standard_normal(100).astype(np.float32)
    else:
        sc = np.full(P, 0.25, np.float32)
        sc[:7] = [0.1, 0.9, 0.25, 0.3, 0.2, 0.25, 1.0]
    out = ops.roc_auc_stats(torch.from_numpy(sc).to(DEV), torch.from_numpy(y).to(DEV)).cpu().numpy()
    auc, two_u, n_pos, n_neg = oracle.roc_auc(sc, y)
    assert (int(out[1]), int(out[2]), int(out[3])) == (n_pos, n_neg, 0)
    assert int(out[4]) == two_u
    assert out[0] == auc
    assert abs(out[0] - roc_auc_score(y, sc)) < 1e-12
    assert ops.roc_auc(torch.from_numpy(sc).to(DEV), torch.from_numpy(y).to(DEV)) == auc


def test_roc_auc_error_cases():
    from disenlink_b200 import ops
    sc = torch.rand(100, device=DEV)
    with pytest.raises(ValueError):
        ops.roc_auc(sc, torch.ones(100, device=DEV))
    with pytest.raises(ValueError):
        ops.roc_auc(sc, torch.zeros(100, device=DEV))
    bad = sc.clone()
    bad[3] = float("nan")
    lab = (torch.rand(100, device=DEV) < 0.5).float()
    lab[0], lab[1] = 1.0, 0.0
    with pytest.raises(ValueError):
        ops.roc_auc(bad, lab)
    with pytest.raises(ValueError):
        ops.roc_auc(torch.empty(0, device=DEV), torch.empty(0, device=DEV))


@pytest.mark.parametrize("n,e,seed", [(50, 400, 0), (300, 6000, 7), (20000, 200000, 123456789012345)])
def test_structured_negative_sampling_vs_oracle(n, e, seed):
    """dl_structured_negative_sampling == the numpy oracle bit for bit; PyG's contract holds."""
    from disenlink_b200 import ops
    from oracle import oracle
    rng = np.random.default_rng(n)
    src, dst = rng.integers(0, n, e), rng.integers(0, n, e)
    ei = torch.from_numpy(np.stack([src, dst])).to(DEV)
    i, j, k = ops.structured_negative_sampling(ei, n, seed=seed)
    assert np.array_equal(i.cpu().numpy(), src) and np.array_equal(j.cpu().numpy(), dst)
    assert np.array_equal(k.cpu().numpy(), oracle.structured_negative_sampling(src, dst, n, seed=seed))
    keys = set((src * n + dst).tolist())
    assert all((int(a) * n + int(b)) not in keys for a, b in zip(src, k.cpu().numpy()))


def test_structured_negative_sampling_dense_rows():
    """A node adjacent to all but one node always gets that node (fallback after rejected draws);
    a node adjacent to every node has no negative: error, like an endless loop in PyG."""
    from disenlink_b200 import ops
    from oracle import oracle
    n = 40
    src = np.concatenate([np.zeros(n - 1, np.int64), np.array([1, 2], np.int64)])
    dst = np.concatenate([np.delete(np.arange(n), 17), np.array([5, 9], np.int64)])
    ei = torch.from_numpy(np.stack([src, dst])).to(DEV)
    i, j, k = ops.structured_negative_sampling(ei, n, seed=3, max_tries=4)
    k = k.cpu().numpy()
    assert (k[:n - 1] == 17).all()
    assert np.array_equal(k, oracle.structured_negative_sampling(src, dst, n, seed=3, max_tries=4))
    full = torch.from_numpy(np.stack([np.zeros(n, np.int64), np.arange(n)])).to(DEV)
    with pytest.raises(RuntimeError):
        ops.structured_negative_sampling(full, n, seed=0, max_tries=2)


def test_structured_negative_sampling_large_properties():
    """2e7 edges: no sampled (i, k) is an edge (checked through the CSR on the device)."""
    from disenlink_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(5)
    n, e = 2_000_000, 20_000_000
    ei = torch.randint(0, n, (2, e), device=DEV, generator=g)
    i, j, k = ops.structured_negative_sampling(ei, n, seed=11)
    assert int(k.min()) >= 0 and int(k.max()) < n
    edge_keys = torch.unique(ei[0] * n + ei[1])
    pos = torch.searchsorted(edge_keys, i * n + k).clamp_(max=edge_keys.numel() - 1)
    assert not bool((edge_keys[pos] == i * n + k).any())
    cnt = torch.bincount(k, minlength=n).float()
    assert float(((cnt - e / n) ** 2 / (e / n)).sum()) < 1.2 * n


@pytest.mark.parametrize("n_bytes", [0, 16, 4096 + 16, 1_000_003])
def test_push_slice_kernel(n_bytes):
    """dl_push_slice (the all-gather push of the partitioned step) with local buffers standing in for
    the peers' copies: every destination receives exactly the slice, nothing around it is touched,
    a byte tail that is not a multiple of 16 included."""
    import ctypes
    from disenlink_b200._lib import check, lib, stream_of
    dev = torch.device(DEV)
    g = torch.Generator(device=DEV).manual_seed(n_bytes)
    src = torch.randint(0, 255, (n_bytes + 32,), dtype=torch.uint8, device=DEV, generator=g)
    dsts = [torch.full((n_bytes + 64,), 7, dtype=torch.uint8, device=DEV) for _ in range(3)]
    arr = (ctypes.c_void_p * 3)(*[d.data_ptr() + 16 for d in dsts])
    check(lib().dl_push_slice(src.data_ptr() + 16, arr, 3, n_bytes, stream_of(dev)), "dl_push_slice")
    torch.cuda.synchronize()
    for d in dsts:
        assert torch.equal(d[16:16 + n_bytes], src[16:16 + n_bytes])
        assert bool((d[:16] == 7).all()) and bool((d[16 + n_bytes:] == 7).all())
    # misaligned pointers are rejected, not silently mis-copied
    bad = (ctypes.c_void_p * 1)(dsts[0].data_ptr() + 4)
    assert lib().dl_push_slice(src.data_ptr() + 16, bad, 1, 64, stream_of(dev)) != 0


@pytest.mark.parametrize("K,d", [(8, 16), (5, 32), (3, 7)])
def test_fused_exchange_kernels_write_every_peer(K, d):
    """dl_factor_spmm_fwd_push / dl_pair_score_bwd_push with local buffers standing in for the peers'
    copies: every peer receives, bit for bit, the rows the kernel wrote locally (streamed shapes push
    from the row epilogues, chain and empty-row kernels; other shapes through dl_push_slice), and
    nothing else in the peer arrays is touched."""
    import ctypes
    from disenlink_b200 import ops
    from disenlink_b200._lib import check, lib, ptr, stream_of
    from disenlink_b200.graph import Graph
    rng = np.random.default_rng(K * 10 + d)
    n = 5000
    src = np.concatenate([rng.integers(0, n, 40000), np.full(3000, 7)])
    dst = np.concatenate([rng.integers(0, n, 40000), rng.choice(n, 3000, replace=False)])
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    g = Graph.from_edges(t(src), t(dst), n)
    Z = t((rng.standard_normal((n, K, d)) * 0.3).astype(np.float32))
    kstar, w, s = ops.edge_attn_fwd(g, Z, 1.0)
    H_ref = ops.factor_spmm_fwd(g, Z, kstar, w, s, 0.5)
    dev = torch.device(DEV)
    peers = [torch.full_like(Z, 123.0) for _ in range(2)]
    arr = (ctypes.c_void_p * 2)(*[p.data_ptr() for p in peers])
    H = torch.empty_like(Z)
    check(lib().dl_factor_spmm_fwd_push(g.ref, ptr(Z), ptr(kstar), ptr(w), ptr(s), K, d, 0.5, 0.5, ptr(H), None,
                                        ptr(g.hub_scratch(K * d)), arr, 2, stream_of(dev)), "spmm push")
    assert torch.equal(H, H_ref)
    for p in peers:
        assert torch.equal(p, H)
    # attention + row sums: s goes to the peers
    s_peers = [torch.full_like(s, 9.0) for _ in range(2)]
    arr = (ctypes.c_void_p * 2)(*[p.data_ptr() for p in s_peers])
    k2, w2, s2 = torch.empty_like(kstar), torch.empty_like(w), torch.empty_like(s)
    check(lib().dl_edge_attn_fwd_push(g.ref, ptr(Z), K, d, 1.0, ptr(k2), ptr(w2), ptr(s2), ptr(g.hub_scratch(K)),
                                      arr, 2, stream_of(dev)), "attn push")
    assert torch.equal(k2, kstar) and torch.equal(w2, w) and torch.equal(s2, s)
    for p in s_peers:
        assert torch.equal(p, s)
    # backward pass 1: r goes to the peers, dZ stays local
    G = t(rng.standard_normal((n, K, d)).astype(np.float32))
    dZ_ref, r_ref = torch.zeros_like(Z), torch.empty_like(s)
    ops.factor_bwd_gather(g, Z, G, kstar, w, s, 0.5, dZ_ref, r_ref)
    r_peers = [torch.full_like(s, 9.0) for _ in range(2)]
    arr = (ctypes.c_void_p * 2)(*[p.data_ptr() for p in r_peers])
    dZ1, r1 = torch.zeros_like(Z), torch.empty_like(s)
    check(lib().dl_factor_bwd_gather_push(g.ref, ptr(Z), ptr(G), ptr(kstar), ptr(w), ptr(s), K, d, 0.5, 0.5, ptr(dZ1),
                                          ptr(r1), None, None, ptr(g.hub_scratch(K * d)), arr, 2, stream_of(dev)),
          "bwd gather push")
    assert torch.equal(dZ1, dZ_ref) and torch.equal(r1, r_ref)
    for p in r_peers:
        assert torch.equal(p, r_ref)
    # decoder backward: dH goes to the peers, dZ stays local
    P = 30000
    batch = ops.PairBatch(t(rng.integers(0, n, P)), t(rng.integers(0, n, P)), n)
    dS = t(rng.standard_normal(P).astype(np.float32))
    dZ_ref, dH_ref = ops.pair_score_bwd(Z, H, batch, dS, 1.0)
    inc, inc_pair = batch.incidence()
    peers = [torch.full_like(Z, -5.0) for _ in range(3)]
    arr = (ctypes.c_void_p * 3)(*[p.data_ptr() for p in peers])
    dZ, dH = torch.empty_like(Z), torch.empty_like(Z)
    check(lib().dl_pair_score_bwd_push(inc.ref, ptr(inc_pair), ptr(Z), ptr(H), ptr(dS), K, d, 1.0, ptr(dZ), ptr(dH),
                                       ptr(inc.hub_scratch(2 * K * d)), arr, 3, stream_of(dev)), "pair bwd push")
    assert torch.equal(dZ, dZ_ref) and torch.equal(dH, dH_ref)
    for p in peers:
        assert torch.equal(p, dH)


def test_graph_cache_identity_and_sparse_adj():
    """One module, several adjacencies (train vs eval, a new split): the cached CSR must follow the tensor
    OBJECT, not its address (the allocator reuses a freed adjacency's address); sparse COO / CSR
    adjacencies are accepted."""
    g = load_golden("module_small")
    n = int(g["N"])
    m = build(g)
    x = torch.from_numpy(g["x"]).to(DEV)
    adj1 = dense_adj_sym(g["src"], g["dst"], n).to(DEV)
    H1, _ = m(x, adj1)
    ptr1 = adj1.data_ptr()
    del adj1
    adj2 = dense_adj_sym(g["src"][:60], g["dst"][:60], n).to(DEV)      # usually lands at the same address
    H2, _ = m(x, adj2)
    from disenlink_b200.model import Disentangle
    fresh = Disentangle(int(g["F"]), int(g["nhid"]), int(g["d"]), nfactor=int(g["K"]), beta=float(g["beta"]), t=1)
    fresh.load_state_dict(m.state_dict())
    H2_ref, _ = fresh.to(DEV)(x, adj2)
    assert torch.equal(H2, H2_ref), f"stale graph served (address reused: {adj2.data_ptr() == ptr1})"
    assert not torch.equal(H1, H2)
    # in-place edit of the same tensor bumps its version: the CSR is rebuilt
    adj2[0, :] = 0
    adj2[:, 0] = 0
    H3, _ = m(x, adj2)
    assert torch.equal(H3[0], m.beta * m.project(x)[0].reshape(-1).detach()) or torch.allclose(
        H3[0], m.beta * m.project(x)[0].reshape(-1).detach(), rtol=1e-6, atol=1e-7)
    # sparse layouts
    adj_d = dense_adj_sym(g["src"], g["dst"], n).to(DEV)
    Hd, sc = m(x, adj_d)
    for sp in (adj_d.to_sparse(), adj_d.to_sparse_csr()):
        Hs, scorer = m(x, sp)
        assert torch.equal(Hs, Hd)


@pytest.mark.parametrize("row_floats,masked", [(8, False), (128, False), (128, True), (160, True)])
def test_push_rows_kernel(row_floats, masked):
    """dl_push_rows / dl_need_masks: indexed rows (and routed slices) land where the descriptor says,
    everything else in the destination stays untouched (peers emulated by local buffers)."""
    import ctypes
    from disenlink_b200._lib import DlPushDesc, check, lib, stream_of
    rng = np.random.default_rng(row_floats + masked)
    n_src, n_dst = 5000, 7000
    src = torch.from_numpy(rng.standard_normal((n_src, row_floats)).astype(np.float32)).to(DEV)
    K = 5 if row_floats == 160 else 8                      # slices of 2^n vectors: (8, d=16) and (5, d=32)
    vpf = row_floats // K // 4 if masked else 0
    peers, descs_keep = [], []
    descs = (DlPushDesc * 3)()
    expect = []
    for q in range(3):
        n = [1200, 0, 3000][q]
        sidx = torch.from_numpy(rng.choice(n_src, n, replace=False).astype(np.int32)).to(DEV)
        contiguous = q == 0
        didx = None if contiguous else torch.from_numpy(rng.choice(n_dst, n, replace=False).astype(np.int32)).to(DEV)
        mask = torch.from_numpy(rng.integers(0, 2 ** K, n_src).astype(np.int32)).to(DEV) if masked else None
        dst = torch.full((n_dst, row_floats), -7.0, device=DEV)
        base = 100 if contiguous else 0
        exp = dst.clone()
        rows = (torch.arange(n, device=DEV) + base) if contiguous else didx.long()
        val = src[sidx.long()]
        if masked:
            keep = ((mask[sidx.long()].long()[:, None] >> torch.arange(K, device=DEV)[None, :]) & 1).bool()
            keep = keep.repeat_interleave(row_floats // K, dim=1)
            val = torch.where(keep, val, exp[rows])
        exp[rows] = val
        expect.append(exp)
        peers.append(dst)
        descs_keep.append((sidx, didx, mask))
        descs[q].dst = dst.data_ptr() + base * row_floats * 4
        descs[q].src_idx = sidx.data_ptr()
        descs[q].dst_idx = didx.data_ptr() if didx is not None else None
        descs[q].mask = mask.data_ptr() if mask is not None else None
        descs[q].n = n
    check(lib().dl_push_rows(src.data_ptr(), row_floats * 4, vpf, descs, 3, stream_of(torch.device(DEV))), "dl_push_rows")
    torch.cuda.synchronize()
    for dst, exp in zip(peers, expect):
        assert torch.equal(dst, exp)


def test_need_masks_kernel():
    """masks[p][row] = OR of (1 << kstar) over the row's entries whose column lies in owner p's halo block."""
    from disenlink_b200._lib import check, lib, ptr, stream_of
    from disenlink_b200.graph import Graph
    rng = np.random.default_rng(4)
    n_own, n_tot, K, world = 3000, 9000, 8, 4
    deg = rng.integers(0, 40, n_own)
    rowptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    col = np.concatenate([np.sort(rng.choice(n_tot, d_, replace=False)) for d_ in deg]).astype(np.int32)
    kstar = rng.integers(0, K, col.size).astype(np.uint8)
    halo_off = np.array([n_own, n_own + 1500, n_own + 1500, n_own + 4000, n_tot], np.int32)   # block 1 empty (self)
    g = Graph(torch.from_numpy(rowptr).to(DEV), torch.from_numpy(col).to(DEV), n_own, row_base=0, n_global=n_tot)
    masks = torch.full((world, n_own), -1, dtype=torch.int32, device=DEV)
    check(lib().dl_need_masks(g.ref, ptr(torch.from_numpy(kstar).to(DEV)), ptr(torch.from_numpy(halo_off).to(DEV)),
                              world, ptr(masks), stream_of(torch.device(DEV))), "dl_need_masks")
    rows = np.repeat(np.arange(n_own), deg)
    exp = np.zeros((world, n_own), np.int64)
    part = np.searchsorted(halo_off, col, side="right") - 1
    sel = col >= n_own
    np.bitwise_or.at(exp, (part[sel], rows[sel]), 1 << kstar[sel].astype(np.int64))
    assert np.array_equal(masks.cpu().numpy().astype(np.int64), exp)


def test_projection_3xtf32_keeps_fp32_accuracy():
    """Disentangle.project on tensor cores (three TF32 GEMMs per product): forward within 1e-5 of fp64 and
    parameter gradients within 2e-4, both at least 10x closer to fp64 than plain TF32."""
    from disenlink_b200.model import Disentangle
    torch.manual_seed(0)
    n, Fdim, nhid, d, K = 3000, 500, 256, 32, 5
    x = torch.randn(n, Fdim, device=DEV)
    m = Disentangle(Fdim, nhid, d, nfactor=K, beta=0.5, t=1).to(DEV)
    m64 = Disentangle(Fdim, nhid, d, nfactor=K, beta=0.5, t=1).double().to(DEV)
    m64.load_state_dict({k: v.double() for k, v in m.state_dict().items()})
    Z64 = m64.project(x.double())
    Z64.square().sum().backward()
    g64 = [p.grad.clone() for p in m64.parameters()]

    def run(mode, tf32):
        m.zero_grad()
        m.projection = mode
        old = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        try:
            Z = m.project(x)
            Z.square().sum().backward()
        finally:
            torch.backends.cuda.matmul.allow_tf32 = old
        ez = float((Z.detach().double() - Z64.detach()).abs().max() / Z64.detach().abs().max())
        eg = max(float((p.grad.double() - g).abs().max() / g.abs().max()) for p, g in zip(m.parameters(), g64))
        return ez, eg
    e32, g32 = run("fp32", False)
    e3, g3 = run("3xtf32", False)
    etf, gtf = run("fp32", True)                       # plain TF32: what 3xTF32 must beat by an order of magnitude
    print("projection max rel err vs fp64 (Z, grads): fp32 %.2e %.2e, 3xtf32 %.2e %.2e, plain tf32 %.2e %.2e"
          % (e32, g32, e3, g3, etf, gtf))
    assert e3 < 1e-5 and e3 < etf / 10
    assert g3 < 2e-4 and g3 < gtf / 10
