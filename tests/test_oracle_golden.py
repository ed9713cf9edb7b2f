"""Pins the CPU oracle (oracle/) to the dense reference through the committed golden vectors.

The fixtures were produced by tests/golden/make_golden.py from /root/reference/model.py
(unchanged).  Tolerances: integers bit-exact; floats 1e-5 relative to the tensor's max-abs
(BASELINE.json north_star: "within 1e-5 relative (fp32)").  Hard routing is discontinuous, so
entries whose reference top-1/top-2 softmax margin is below 1e-5 may legitimately route
differently under a different fp32 summation order; the test counts them, bounds their margin and
excludes what they touch -- it never ignores a flip on a well-separated entry.
"""
import numpy as np
import pytest

from conftest import GRAPH_FIXTURES, load_golden

RTOL = 1e-5
TIE_MARGIN = 1e-5


def relerr(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    den = max(np.abs(b).max(), 1e-30) if b.size else 1.0
    return float(np.abs(a - b).max() / den) if b.size else 0.0


def routed(oracle, g):
    Z = g["Z"]
    rowptr, col = oracle.csr_from_edges(g["src"], g["dst"], int(g["N"]))
    H, ks, w, s = oracle.factor_fwd(rowptr, col, Z, float(g["beta"]), float(g["T"]))
    return rowptr, col, H, ks, w, s


@pytest.mark.parametrize("name", GRAPH_FIXTURES)
def test_csr_equals_adj_sym_nonzero(oracle, name):
    g = load_golden(name)
    rowptr, col = oracle.csr_from_edges(g["src"], g["dst"], int(g["N"]))
    assert np.array_equal(oracle.rows_of(rowptr), g["ref_rows"])
    assert np.array_equal(col.astype(np.int64), g["ref_cols"])
    rev = oracle.rev_index(rowptr, col)
    rows = oracle.rows_of(rowptr)
    assert np.array_equal(rows[rev], col.astype(np.int64))
    assert np.array_equal(col[rev].astype(np.int64), rows)


def flips(g, ks):
    bad = np.nonzero(ks != g["ref_kstar"])[0]
    assert np.all(g["ref_margin"][bad] < TIE_MARGIN), (
        f"{bad.size} routing mismatches, worst margin {g['ref_margin'][bad].max():.3e}")
    return bad


@pytest.mark.parametrize("name", GRAPH_FIXTURES)
def test_routing_and_attention(oracle, name):
    g = load_golden(name)
    rowptr, col, H, ks, w, s = routed(oracle, g)
    bad = flips(g, ks)
    ok = np.ones(ks.size, bool)
    ok[bad] = False
    assert relerr(w[ok], g["ref_w"][ok]) < RTOL
    rows = oracle.rows_of(rowptr)
    # rows with a flipped entry have a different s by construction; everything else must agree
    clean_row = np.ones(int(g["N"]), bool)
    clean_row[rows[bad]] = False
    assert relerr(s[clean_row], g["ref_s"][clean_row]) < RTOL
    att = oracle.att_values(rowptr, col, ks, w, s)
    ok_att = ok & clean_row[col]
    assert relerr(att[ok_att], g["ref_att"][ok_att]) < RTOL
    # symmetry of the routing (what the atomic-free backward relies on)
    rev = oracle.rev_index(rowptr, col)
    assert np.array_equal(ks, ks[rev])
    assert np.array_equal(w.view(np.uint32), w[rev].view(np.uint32))


@pytest.mark.parametrize("name", GRAPH_FIXTURES)
def test_embeddings_and_scores(oracle, name):
    g = load_golden(name)
    N, K, d = int(g["N"]), int(g["K"]), int(g["d"])
    rowptr, col, H, ks, w, s = routed(oracle, g)
    bad = flips(g, ks)
    rows = oracle.rows_of(rowptr)
    # a flip at (i,j) changes s[i,*] hence H of i and of every neighbour of i
    dirty = np.zeros(N, bool)
    dirty[rows[bad]] = True
    dirty_n = dirty.copy()
    dirty_n[rows[dirty[col]]] = True
    Href = g["ref_H"].reshape(N, K, d)
    assert relerr(H[~dirty_n], Href[~dirty_n]) < RTOL
    _, prob = oracle.pair_score_fwd(g["pu"], g["pv"], g["Z"], H, float(g["T"]))
    okp = ~(dirty_n[g["pu"]] | dirty_n[g["pv"]])
    assert relerr(prob[okp], g["ref_prob"][okp]) < RTOL
    if "ref_link_pred" in g:
        uu, vv = np.meshgrid(np.arange(N), np.arange(N), indexing="ij")
        _, full = oracle.pair_score_fwd(uu.reshape(-1), vv.reshape(-1), g["Z"], H, float(g["T"]))
        assert not dirty_n.any()
        assert relerr(full.reshape(N, N), g["ref_link_pred"]) < RTOL


def script_loss_and_grad(oracle, g, Z, H):
    """main_disentangled.py:195 on pair lists: pairs occurring exactly once, mean BCE, neg / m."""
    N, T = int(g["N"]), float(g["T"])
    pu, pv = oracle.pairs_exactly_once(g["loss_pos_u"], g["loss_pos_v"], N)
    nu, nv = oracle.pairs_exactly_once(g["loss_neg_u"], g["loss_neg_v"], N)
    m = float(g["m"])
    _, pp = oracle.pair_score_fwd(pu, pv, Z, H, T)
    _, pn = oracle.pair_score_fwd(nu, nv, Z, H, T)
    loss = oracle.bce_mean(pp, 1.0) + oracle.bce_mean(pn, 0.0) / m
    dS = np.concatenate([oracle.bce_mean_grad_logit(pp, 1.0),
                         oracle.bce_mean_grad_logit(pn, 0.0) / np.float32(m)]).astype(np.float32)
    u = np.concatenate([pu, nu])
    v = np.concatenate([pv, nv])
    return loss, u, v, dS


@pytest.mark.parametrize("name", [n for n in GRAPH_FIXTURES if "ref_dZ" in load_golden(n)])
def test_loss_and_gradients(oracle, name):
    g = load_golden(name)
    rowptr, col, H, ks, w, s = routed(oracle, g)
    bad = flips(g, ks)
    loss, u, v, dS = script_loss_and_grad(oracle, g, g["Z"], H)
    dZ_dec, dH = oracle.pair_score_bwd(u, v, g["Z"], H, dS, float(g["T"]))
    dZ = oracle.factor_bwd(rowptr, col, g["Z"], dH, ks, w, s, float(g["beta"]), float(g["T"]),
                           dZ_init=dZ_dec)
    if bad.size == 0:
        assert abs(loss - float(g["ref_loss"])) <= RTOL * abs(float(g["ref_loss"]))
        assert relerr(dZ, g["ref_dZ"]) < 5 * RTOL
    else:
        # near-tie flips perturb a few rows; the bulk must still agree
        assert abs(loss - float(g["ref_loss"])) <= 1e-3 * abs(float(g["ref_loss"]))
        err = np.abs(dZ - g["ref_dZ"]).reshape(dZ.shape[0], -1).max(1) / np.abs(g["ref_dZ"]).max()
        assert np.mean(err < 5 * RTOL) > 0.9


def test_expf_accuracy(oracle):
    x = np.concatenate([np.linspace(-87, 88, 20001), np.random.default_rng(0).normal(0, 3, 20000)])
    x = x.astype(np.float32)
    got = oracle.expf(x).astype(np.float64)
    want = np.exp(x.astype(np.float64))
    ulp = np.abs(got - want) / np.spacing(want.astype(np.float32)).astype(np.float64)
    assert ulp.max() < 1.5
    assert oracle.expf(np.float32(0.0)) == 1.0
    assert np.isinf(oracle.expf(np.float32(100.0)))
    assert oracle.expf(np.float32(-200.0)) <= 1.5e-45
    assert np.isnan(oracle.expf(np.float32(np.nan)))


def test_dot_is_symmetric_and_accurate(oracle):
    rng = np.random.default_rng(1)
    for d in (1, 3, 4, 6, 8, 12, 16, 32, 64, 100, 128):
        x = rng.standard_normal(d).astype(np.float32)
        y = rng.standard_normal(d).astype(np.float32)
        a, b = oracle.dot(x, y), oracle.dot(y, x)
        assert a.view(np.uint32) == b.view(np.uint32)
        assert abs(float(a) - float(np.dot(x.astype(np.float64), y.astype(np.float64)))) < 1e-5 * d


def test_bce_weighted_matches_torch(oracle):
    """oracle.bce_weighted against torch's own BCE + autograd through a sigmoid (CPU), saturated
    scores included.  [ref: main_disentangled.py:195]"""
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(5)
    logit = (rng.standard_normal(4096) * 6).astype(np.float32)
    logit[:6] = [200, -200, 120, -120, 17, -17]
    y = (rng.random(4096) < 0.3).astype(np.float32)
    w = rng.random(4096).astype(np.float32)
    S = torch.from_numpy(logit).requires_grad_(True)
    p = torch.sigmoid(S)
    ref = (F.binary_cross_entropy(p, torch.from_numpy(y), reduction="none") * torch.from_numpy(w)).sum()
    ref.backward()
    loss, dS = oracle.bce_weighted(p.detach().numpy(), y, w)
    assert abs(loss - ref.item()) <= 2e-6 * abs(ref.item())
    assert np.abs(dS - S.grad.numpy()).max() <= 1e-6 * float(S.grad.abs().max())


@pytest.mark.parametrize("case", ["distinct", "ties", "saturated", "signed_zero"])
def test_roc_auc_oracle_matches_sklearn(oracle, case):
    """oracle.roc_auc against sklearn.metrics.roc_auc_score (the function the script calls,
    main_disentangled.py:204,219)."""
    from sklearn.metrics import roc_auc_score
    rng = np.random.default_rng(11)
    n = 5000
    y = (rng.random(n) < 0.2).astype(np.float32)
    if case == "distinct":
        sc = rng.random(n).astype(np.float32)
    elif case == "ties":
        sc = (rng.integers(0, 37, n) / 37.0).astype(np.float32)
    elif case == "saturated":
        sc = 1.0 / (1.0 + np.exp(-(rng.standard_normal(n) * 30))).astype(np.float32)     # many exact 0 / 1
        sc = sc.astype(np.float32)
    else:
        sc = np.where(rng.random(n) < 0.5, np.float32(0.0), np.float32(-0.0)).astype(np.float32)
        sc[:100] = rng.standard_normal(100).astype(np.float32)
    auc, two_u, n_pos, n_neg = oracle.roc_auc(sc, y)
    assert n_pos == int(y.sum()) and n_neg == n - n_pos
    assert abs(auc - roc_auc_score(y, sc)) < 1e-12


def test_philox_known_answers(oracle):
    """Random123's published known-answer vectors for Philox4x32-10 pin the draw stream of the
    negative sampler."""
    z = oracle.philox4x32_10([0], [0], [0], [0], 0, 0)
    assert [int(x[0]) for x in z] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = 0xFFFFFFFF
    o = oracle.philox4x32_10([f], [f], [f], [f], f, f)
    assert [int(x[0]) for x in o] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    p = oracle.philox4x32_10([0x243f6a88], [0x85a308d3], [0x13198a2e], [0x03707344], 0xa4093822, 0x299f31d0)
    assert [int(x[0]) for x in p] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_structured_negative_sampling_oracle_properties(oracle):
    """PyG semantics (main_disentangled.py:160): (i, k) is never an edge column, k is uniform over
    the non-neighbours, the stream depends on the seed only."""
    rng = np.random.default_rng(2)
    n, e = 300, 6000
    src, dst = rng.integers(0, n, e), rng.integers(0, n, e)
    k = oracle.structured_negative_sampling(src, dst, n, seed=7)
    keys = set((src * n + dst).tolist())
    assert all((int(i) * n + int(kk)) not in keys for i, kk in zip(src, k))
    assert k.min() >= 0 and k.max() < n
    assert np.array_equal(k, oracle.structured_negative_sampling(src, dst, n, seed=7))
    assert not np.array_equal(k, oracle.structured_negative_sampling(src, dst, n, seed=8))
    # roughly uniform: chi-square over nodes, 6000 draws into 300 bins (mean 20)
    cnt = np.bincount(k, minlength=n)
    assert ((cnt - e / n) ** 2 / (e / n)).sum() < 2.0 * n
