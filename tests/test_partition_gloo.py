"""World-size-2 test of the node-partitioned orchestration (disenlink_b200/partition.py) on CPU.

torch.distributed with the gloo backend; the CUDA kernels are replaced by a backend that calls the
CPU oracle on the rank's LOCAL graph, so what is under test is the host logic: nnz-balanced split
points, owned-entry selection, the halo (remote columns + pair endpoints), the local index space and the
remapped CSR / pair / incidence lists, the send lists both sides agree on, and the exchange sequence
(Z, s, H for the pair endpoints, prob, dH, r) over the torch.distributed point-to-point path.
The partitioned result must equal the single-process oracle result bit for bit.
"""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from disenlink_b200.partition import (HaloPlan, NodeOrder, NodePartition, PartitionedLinkStep, locality_partition,
                                      owned_entries, pair_shard)


class _LocalGraph:
    """rowptr padded to n_tot rows (halo rows are empty) so the oracle sees a square local problem."""

    def __init__(self, rowptr, col, n_own, n_tot):
        rp = np.full(n_tot + 1, int(rowptr[-1]), np.int64)
        rp[:n_own + 1] = rowptr.numpy()
        self.rowptr, self.col = rp, col.numpy().astype(np.int32)
        self.nnz, self.n_own, self.n_tot = int(self.col.size), n_own, n_tot


class OracleBackend:
    """Test double for CudaBackend: same interface, numpy + oracle, writes only owned rows."""

    def __init__(self):
        from oracle import oracle
        self.o = oracle

    def build_csr(self, src, dst, part):
        rows, cols = owned_entries(src, dst, part)
        key = np.unique(rows.numpy() * part.n_global + cols.numpy())
        r, c = key // part.n_global, key % part.n_global
        rowptr = np.zeros(part.n_local + 1, np.int64)
        np.cumsum(np.bincount(r, minlength=part.n_local), out=rowptr[1:])
        return torch.from_numpy(rowptr), torch.from_numpy(c.astype(np.int32))

    def make_graph(self, rowptr, col_local, n_own, n_tot):
        return _LocalGraph(rowptr, col_local, n_own, n_tot)

    def build_incidence(self, u, v, part):
        un, vn = u.numpy(), v.numpy()
        P = un.size
        node = np.stack([un, vn], 1).reshape(-1)                # incidence 2p = u side, 2p + 1 = v side
        other = np.stack([vn, un], 1).reshape(-1)
        pair = np.repeat(np.arange(P), 2)
        keep = (node >= part.lo) & (node < part.hi)
        node, other, pair = node[keep] - part.lo, other[keep], pair[keep]
        order = np.argsort(node, kind="stable")
        ptr = np.zeros(part.n_local + 1, np.int64)
        np.cumsum(np.bincount(node, minlength=part.n_local), out=ptr[1:])
        return (torch.from_numpy(ptr), torch.from_numpy(other[order].astype(np.int32)),
                torch.from_numpy(pair[order].astype(np.int32)))

    def make_pairs(self, u_loc, v_loc, n_tot):
        class B:
            pass
        b = B()
        b.u, b.v = u_loc.numpy(), v_loc.numpy()
        return b

    def edge_attn_fwd(self, g, Z, T, kstar, w, s):
        ks, ww, ss = self.o.edge_attn_fwd(g.rowptr, g.col, Z.numpy(), T)
        kstar[:g.nnz] = torch.from_numpy(ks)
        w[:g.nnz] = torch.from_numpy(ww)
        s[:g.n_own] = torch.from_numpy(ss[:g.n_own])

    def factor_spmm_fwd(self, g, Z, kstar, w, s, beta, H, sj=None, zs=None):
        out = self.o.factor_spmm_fwd(g.rowptr, g.col, Z.numpy(), kstar[:g.nnz].numpy(), w[:g.nnz].numpy(),
                                     s.numpy(), beta)
        H[:g.n_own] = torch.from_numpy(out[:g.n_own])

    def pair_score_fwd(self, Z, H, shard, T, prob_slice):
        _, prob = self.o.pair_score_fwd(shard.u, shard.v, Z.numpy(), H.numpy(), T)
        prob_slice.copy_(torch.from_numpy(prob))

    def pair_score_bwd(self, inc, inc_pair, Z, H, dS, T, dZ, dH):
        # one-sided: the `other` endpoints are shifted into a shadow copy of the arrays, so only the owned
        # endpoint of every incidence receives its term
        n_own, n_tot = inc.n_own, inc.n_tot
        rows = np.repeat(np.arange(n_own), np.diff(inc.rowptr[:n_own + 1]))
        Z2, H2 = np.concatenate([Z.numpy()] * 2), np.concatenate([H.numpy()] * 2)
        a, b = self.o.pair_score_bwd(rows, inc.col.astype(np.int64) + n_tot, Z2, H2,
                                     dS.numpy()[inc_pair.numpy()], T)
        dZ[:n_own] = torch.from_numpy(a[:n_own])
        dH[:n_own] = torch.from_numpy(b[:n_own])

    def link_bce(self, prob, labels, weights, dS):
        loss, ds = self.o.bce_weighted(prob.numpy(), labels.numpy(), weights.numpy())
        dS.copy_(torch.from_numpy(ds))
        return torch.tensor(loss, dtype=torch.float32)

    def factor_bwd_gather(self, g, Z, G, kstar, w, s, beta, dZ, r, plan=None):
        n = g.n_own
        dz = np.zeros((g.n_tot,) + tuple(dZ.shape[1:]), np.float32)
        dz[:n] = dZ.numpy()
        rr = np.zeros(tuple(r.shape), np.float32)
        self.o.factor_bwd_gather(g.rowptr, g.col, Z.numpy(), G.numpy(), kstar[:g.nnz].numpy(),
                                 w[:g.nnz].numpy(), s.numpy(), beta, dz, rr)
        dZ.copy_(torch.from_numpy(dz[:n]))
        r[:n] = torch.from_numpy(rr[:n])

    def factor_bwd_edges(self, g, Z, G, kstar, w, s, r, beta, T, dZ, sj=None, plan=None):
        assert plan is None
        n = g.n_own
        dz = np.zeros((g.n_tot,) + tuple(dZ.shape[1:]), np.float32)
        dz[:n] = dZ.numpy()
        self.o.factor_bwd_edges(g.rowptr, g.col, Z.numpy(), G.numpy(), s.numpy(), r.numpy(), beta, T, dz)
        dZ.copy_(torch.from_numpy(dz[:n]))


def make_inputs(n=203, e=1500, K=3, d=8, P=901, seed=0):
    rng = np.random.default_rng(seed)
    src = torch.from_numpy(rng.integers(0, n, e))
    dst = torch.from_numpy(rng.integers(0, n, e))
    u = torch.from_numpy(np.sort(rng.integers(0, n, P)))
    v = torch.from_numpy(rng.integers(0, n, P))
    lab = torch.from_numpy((rng.random(P) < 0.3).astype(np.float32))
    wts = torch.full((P,), 1.0 / P)
    Z = torch.from_numpy((rng.standard_normal((n, K, d)) * 0.4).astype(np.float32))
    return src, dst, u, v, lab, wts, Z


CASES = {"default": dict(), "empty_ranks": dict(n=3, e=4, K=2, d=4, P=2, seed=3),   # 3 nodes over 4 ranks
         "locality": dict()}            # nodes renumbered by locality_partition, explicit split points
LOCALITY_WORLD = 4


def run_step(world, rank, group=None, case="default"):
    src, dst, u, v, lab, wts, Z = make_inputs(**CASES[case])
    n, K, d = Z.shape
    bounds = None
    if case == "locality":
        order, b = locality_partition(src, dst, n, LOCALITY_WORLD)
        src, dst, u, v, Z = order.relabel(src), order.relabel(dst), order.relabel(u), order.relabel(v), order.rows_to_new(Z)
        bounds = b if world == LOCALITY_WORLD else None
    step = PartitionedLinkStep(src, dst, n, u, v, lab, wts, K, d, 0.6, 1.0, world=world, rank=rank,
                               group=group, backend=OracleBackend(), device=torch.device("cpu"), bounds=bounds)
    part = step.part
    step.Z_own.copy_(Z[part.lo:part.hi])              # each rank only has its own rows before the exchange
    step.run()
    step.run()                                        # a second step over the same buffers (ordering / WAR guard)
    return step, part


def _worker(rank, world, port, out_dir, case="default"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        step, part = run_step(world, rank, case=case)
        n = step.n_own
        torch.save({"dZ": step.dZ.clone(), "H": step.H[:n].clone(), "prob": step.prob.clone(),
                    "loss": step.loss.clone(), "lo": part.lo, "hi": part.hi, "s": step.s[:n].clone(),
                    "r": step.r[:n].clone(), "nnz": step.graph.nnz, "n_halo": step.plan.n_halo,
                    "halo": step.plan.halo.clone(), "Zhalo": step.Z[n:].clone(), "shalo": step.s[n:].clone(),
                    "bounds": part.bounds}, os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_partition_bounds():
    for n, world in [(10, 2), (11, 4), (3, 8), (50_000_000, 8)]:
        parts = [NodePartition(n, world, r) for r in range(world)]
        assert parts[0].lo == 0 and parts[-1].hi == n
        assert all(a.hi == b.lo for a, b in zip(parts[:-1], parts[1:]))
        assert sum(p.n_local for p in parts) == n
    per, lo, hi = pair_shard(10, 4, 3)
    assert (per, lo, hi) == (3, 9, 10)


def test_nnz_balanced_split_points():
    """Power-law degrees: equal node ranges are badly unbalanced, the degree-prefix split is not."""
    rng = np.random.default_rng(0)
    n, e, world = 20000, 300000, 8
    ids = (rng.random(e) ** 3 * n).astype(np.int64)            # hubs at the low ids (no shuffle: worst case)
    src, dst = torch.from_numpy(ids), torch.from_numpy(rng.integers(0, n, e))
    parts = [NodePartition.nnz_balanced(src, dst, n, world, r) for r in range(world)]
    assert parts[0].lo == 0 and parts[-1].hi == n and all(a.hi == b.lo for a, b in zip(parts[:-1], parts[1:]))
    assert all(p.bounds == parts[0].bounds for p in parts)

    def load(ps):
        return [int(owned_entries(src, dst, p)[0].numel()) for p in ps]
    bal, eq = load(parts), load([NodePartition(n, world, r) for r in range(world)])
    assert max(bal) < 1.1 * (sum(bal) / world)
    assert max(eq) > 2.0 * (sum(eq) / world)


def test_halo_plan_single_rank_is_identity():
    part = NodePartition(10, 1, 0)
    plan = HaloPlan(part, torch.empty(0, dtype=torch.int64), torch.empty(0, dtype=torch.int64))
    assert plan.n_tot == 10 and plan.n_halo == 0
    ids = torch.tensor([0, 3, 9])
    assert torch.equal(plan.to_local(ids), ids)


import pytest  # noqa: E402


@pytest.mark.parametrize("world", [2, 4, 8])
def test_gloo_ranks_equal_single_process(tmp_path, world):
    single, part1 = run_step(1, 0)
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    outs = [torch.load(os.path.join(tmp_path, f"rank{r}.pt")) for r in range(world)]
    assert all(o["bounds"] == outs[0]["bounds"] for o in outs)
    assert sum(o["nnz"] for o in outs) == single.graph.nnz
    assert all(0 < o["n_halo"] < part1.n_global for o in outs)          # rank-local storage: own + halo only
    dZ = torch.cat([o["dZ"] for o in outs])
    assert torch.equal(dZ, single.dZ)
    assert torch.equal(torch.cat([o["H"] for o in outs]), single.H)
    assert torch.equal(torch.cat([o["s"] for o in outs]), single.s)
    assert torch.equal(torch.cat([o["r"] for o in outs]), single.r)
    for o in outs:
        assert torch.equal(o["prob"][:single.P], single.prob[:single.P])
        assert torch.equal(o["loss"], single.loss)
        # the halo rows hold exactly the owners' values
        assert torch.equal(o["Zhalo"], single.Z[o["halo"]])
        assert torch.equal(o["shalo"], single.s[o["halo"]])


def test_gloo_ranks_that_own_nothing(tmp_path):
    """3 nodes over 4 ranks: at least one rank owns no node (zero-row arrays, empty send lists, no pairs to
    score); the ranks still agree with one process bit for bit."""
    world = 4
    single, _ = run_step(1, 0, case="empty_ranks")
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    mp.spawn(_worker, args=(world, port, str(tmp_path), "empty_ranks"), nprocs=world, join=True)
    outs = [torch.load(os.path.join(tmp_path, f"rank{r}.pt")) for r in range(world)]
    assert any(o["hi"] == o["lo"] for o in outs)
    assert torch.equal(torch.cat([o["dZ"] for o in outs]), single.dZ)
    assert torch.equal(torch.cat([o["H"] for o in outs]), single.H)
    assert torch.equal(torch.cat([o["r"] for o in outs]), single.r)
    for o in outs:
        assert torch.equal(o["prob"][:single.P], single.prob[:single.P])
        assert torch.equal(o["loss"], single.loss)


# ----------------------------------------------------------------------------------------------
# the partitioned training loop (examples/train_link_partitioned.py): replicated MLPs, summed gradients
# ----------------------------------------------------------------------------------------------
def _trainer_cls():
    import importlib.util
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "train_link_partitioned.py")
    spec = importlib.util.spec_from_file_location("_train_link_partitioned", path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.PartitionedLinkTrainer


def run_training(world, rank, steps=3):
    from disenlink_b200.model import Disentangle
    src, dst, u, v, lab, wts, _ = make_inputs(n=151, e=900, K=3, d=8, P=700, seed=5)
    n, K, d, F_ = 151, 3, 8, 12
    x = torch.randn(n, F_, generator=torch.Generator().manual_seed(7))
    torch.manual_seed(100 + rank)                        # different initial weights per rank: the trainer broadcasts rank 0's
    model = Disentangle(F_, 16, d, nfactor=K, beta=0.6, t=1)
    step = PartitionedLinkStep(src, dst, n, u, v, lab, wts, K, d, 0.6, 1.0, world=world, rank=rank,
                               backend=OracleBackend(), device=torch.device("cpu"))
    opt = torch.optim.Adam(model.parameters(), lr=0.01)
    tr = _trainer_cls()(model, x[step.part.lo:step.part.hi], step, opt)
    losses = [float(tr.train_step()) for _ in range(steps)]
    return losses, [p.detach().clone() for p in model.parameters()], tr.scores().clone()


def _train_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        losses, params, prob = run_training(world, rank)
        torch.save({"losses": losses, "params": params, "prob": prob}, os.path.join(out_dir, f"train{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_gloo_partitioned_training_matches_single_process(tmp_path):
    """Three Adam steps of the partitioned loop at world 2 against one process: same losses, same weights (up to
    the summation order of the gradient all-reduce), identical replicas on the ranks."""
    losses1, params1, prob1 = run_training(1, 0)
    assert losses1[-1] < losses1[0]
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    mp.spawn(_train_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    outs = [torch.load(os.path.join(tmp_path, f"train{r}.pt")) for r in range(2)]
    for a, b in zip(outs[0]["params"], outs[1]["params"]):
        assert torch.equal(a, b)                                              # replicas stay identical
    for o in outs:
        assert np.allclose(o["losses"], losses1, rtol=1e-5, atol=0)
        for a, b in zip(o["params"], params1):
            assert float((a - b).abs().max()) <= 1e-5 * max(float(b.abs().max()), 1e-6)
        assert float((o["prob"] - prob1).abs().max()) < 1e-5


# ----------------------------------------------------------------------------------------------
# locality reorder ("optional locality/degree reorder; keep the permutation and un-permute outputs")
# ----------------------------------------------------------------------------------------------
def _halo_stats(src, dst, n, bounds):
    """-> (halo rows summed over the ranks, share of remote CSR entries, max / mean entries per rank)."""
    s_, d_ = src.numpy(), dst.numpy()
    key = np.unique(np.concatenate([s_ * n + d_, d_ * n + s_]))
    rows, cols = key // n, key % n
    owner = np.searchsorted(np.asarray(bounds[1:]), np.arange(n), side="right")
    remote = owner[rows] != owner[cols]
    halo = np.unique(owner[rows][remote].astype(np.int64) * n + cols[remote]).size
    load = np.bincount(owner[rows], minlength=len(bounds) - 1)
    return halo, float(remote.mean()), float(load.max() / load.mean())


def test_locality_partition_shrinks_the_halo_of_a_real_graph():
    """The real Pubmed citation graph (committed fixture), 8 ranks: the nnz-balanced split of the given numbering
    against locality_partition's renumbering + split points."""
    gd = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pubmed_graph.npz"))
    src, dst, n = torch.from_numpy(gd["src"].astype(np.int64)), torch.from_numpy(gd["dst"].astype(np.int64)), int(gd["N"])
    world = 8
    base = NodePartition.nnz_balanced(src, dst, n, world, 0).bounds
    halo0, remote0, bal0 = _halo_stats(src, dst, n, base)
    order, bounds = locality_partition(src, dst, n, world)
    assert bounds[0] == 0 and bounds[-1] == n and all(a <= b for a, b in zip(bounds[:-1], bounds[1:]))
    assert torch.equal(torch.sort(order.new_of_old).values, torch.arange(n))         # a permutation
    halo1, remote1, bal1 = _halo_stats(order.relabel(src), order.relabel(dst), n, bounds)
    assert halo1 < 0.5 * halo0 and remote1 < 0.5 * remote0, (halo0, halo1, remote0, remote1)
    assert bal1 < 1.10
    order2, bounds2 = locality_partition(src, dst, n, world)                          # deterministic
    assert bounds2 == bounds and torch.equal(order2.new_of_old, order.new_of_old)


def test_renumbered_graph_gives_the_same_result_after_unpermuting():
    src, dst, u, v, lab, wts, Z = make_inputs()
    n, K, d = Z.shape
    plain, _ = run_step(1, 0)
    order, _ = locality_partition(src, dst, n, LOCALITY_WORLD)
    moved, _ = run_step(1, 0, case="locality")
    tol = lambda a, b: float((a - b).abs().max()) <= 2e-5 * max(float(b.abs().max()), 1e-6)  # noqa: E731
    assert tol(order.rows_to_old(moved.H), plain.H) and tol(order.rows_to_old(moved.dZ), plain.dZ)
    assert tol(order.rows_to_old(moved.s), plain.s)
    assert tol(moved.prob[:plain.P], plain.prob[:plain.P]) and tol(moved.loss, plain.loss)


def test_gloo_ranks_with_locality_split_points_equal_single_process(tmp_path):
    world = LOCALITY_WORLD
    single, _ = run_step(1, 0, case="locality")
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    mp.spawn(_worker, args=(world, port, str(tmp_path), "locality"), nprocs=world, join=True)
    outs = [torch.load(os.path.join(tmp_path, f"rank{r}.pt")) for r in range(world)]
    src, dst, *_ = make_inputs()
    _, bounds = locality_partition(src, dst, single.part.n_global, world)
    assert outs[0]["bounds"] == bounds
    assert torch.equal(torch.cat([o["dZ"] for o in outs]), single.dZ)
    assert torch.equal(torch.cat([o["H"] for o in outs]), single.H)
    assert torch.equal(torch.cat([o["r"] for o in outs]), single.r)
    for o in outs:
        assert torch.equal(o["prob"][:single.P], single.prob[:single.P]) and torch.equal(o["loss"], single.loss)
