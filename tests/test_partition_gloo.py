"""World-size-2 test of the node-partitioned orchestration (disenlink_b200/partition.py) on CPU.

torch.distributed with the gloo backend; the CUDA kernels are replaced by a backend that calls
the CPU oracle on the rank's owned rows, so what is under test is the host logic: partition
bounds and padding, owned-entry selection, pair sharding, incidence ranges and the exchange
sequence (six in-place all-gathers per step).  The partitioned result must equal the
single-process oracle result bit for bit.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from disenlink_b200.partition import NodePartition, PartitionedLinkStep, owned_entries, pair_shard


class OracleBackend:
    """Test double for CudaBackend: same interface, numpy + oracle, writes only owned rows."""

    def __init__(self):
        from oracle import oracle
        self.o = oracle

    def build_graph(self, src, dst, part):
        rows, cols = owned_entries(src, dst, part)
        key = np.unique((rows.numpy() + part.lo) * part.n_pad + cols.numpy())
        r, c = key // part.n_pad, (key % part.n_pad).astype(np.int32)
        rowptr = np.zeros(part.n_pad + 1, np.int64)          # global row ids; foreign rows stay empty
        np.cumsum(np.bincount(r, minlength=part.n_pad), out=rowptr[1:])

        class G:
            pass
        g = G()
        g.rowptr, g.col, g.nnz, g.part = rowptr, c, int(c.size), part
        return g

    def build_pairs(self, u, v, part):
        P = int(u.numel())
        per, p_lo, p_hi = pair_shard(P, part.world, part.rank)

        class B:
            pass
        shard = B()
        shard.u, shard.v = u[p_lo:p_hi].numpy(), v[p_lo:p_hi].numpy()
        inc = B()
        inc.u, inc.v, inc.part = u.numpy(), v.numpy(), part
        inc.nnz = int(((u >= part.lo) & (u < part.hi)).sum() + ((v >= part.lo) & (v < part.hi)).sum())
        inc.n_hub = 0
        return shard, inc, None

    def edge_attn_fwd(self, g, Z, T, kstar, w, s):
        ks, ww, ss = self.o.edge_attn_fwd(g.rowptr, g.col, Z.numpy(), T)
        kstar[:g.nnz] = torch.from_numpy(ks)
        w[:g.nnz] = torch.from_numpy(ww)
        s[g.part.lo:g.part.hi] = torch.from_numpy(ss[g.part.lo:g.part.hi])

    def factor_spmm_fwd(self, g, Z, kstar, w, s, beta, H, sj=None, zs=None):
        out = self.o.factor_spmm_fwd(g.rowptr, g.col, Z.numpy(), kstar[:g.nnz].numpy(), w[:g.nnz].numpy(),
                                     s.numpy(), beta)
        H[g.part.lo:g.part.hi] = torch.from_numpy(out[g.part.lo:g.part.hi])

    def pair_score_fwd(self, Z, H, shard, T, prob_slice):
        _, prob = self.o.pair_score_fwd(shard.u, shard.v, Z.numpy(), H.numpy(), T)
        prob_slice.copy_(torch.from_numpy(prob))

    def pair_score_bwd(self, inc, inc_pair, Z, H, dS, T, dZ, dH):
        P = inc.u.size
        a, b = self.o.pair_score_bwd(inc.u, inc.v, Z.numpy(), H.numpy(), dS[:P].numpy(), T)
        lo, hi = inc.part.lo, inc.part.hi
        dZ[lo:hi] = torch.from_numpy(a[lo:hi])
        dH[lo:hi] = torch.from_numpy(b[lo:hi])

    def link_bce(self, prob, labels, weights, dS):
        loss, ds = self.o.bce_weighted(prob.numpy(), labels.numpy(), weights.numpy())
        dS.copy_(torch.from_numpy(ds))
        return torch.tensor(loss, dtype=torch.float32)

    def factor_bwd_gather(self, g, Z, G, kstar, w, s, beta, dZ, r, peers=None, x=None):
        lo, hi = g.part.lo, g.part.hi
        dz, rr = dZ.numpy().copy(), np.zeros(tuple(r.shape), np.float32)
        self.o.factor_bwd_gather(g.rowptr, g.col, Z.numpy(), G.numpy(), kstar[:g.nnz].numpy(),
                                 w[:g.nnz].numpy(), s.numpy(), beta, dz, rr)
        dZ[lo:hi] = torch.from_numpy(dz[lo:hi])
        r[lo:hi] = torch.from_numpy(rr[lo:hi])
        return False                                   # x (per-entry dots for pass 2) not filled

    def factor_bwd_edges(self, g, Z, G, kstar, w, s, r, beta, T, dZ, sj=None, x=None):
        assert x is None
        lo, hi = g.part.lo, g.part.hi
        dz = dZ.numpy().copy()
        self.o.factor_bwd_edges(g.rowptr, g.col, Z.numpy(), G.numpy(), s.numpy(), r.numpy(), beta, T, dz)
        dZ[lo:hi] = torch.from_numpy(dz[lo:hi])


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


def run_step(world, rank, group=None):
    src, dst, u, v, lab, wts, Z = make_inputs()
    n, K, d = Z.shape
    step = PartitionedLinkStep(src, dst, n, u, v, lab, wts, K, d, 0.6, 1.0, world=world, rank=rank,
                               group=group, backend=OracleBackend(), device=torch.device("cpu"))
    part = step.part
    Zf = torch.zeros(part.n_pad, K, d)
    Zf[part.lo:part.hi] = Z[part.lo:part.hi]          # each rank only has its own rows before the gather
    step.run(Zf)
    return step, part


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        step, part = run_step(world, rank)
        torch.save({"dZ": step.dZ[part.lo:part.hi].clone(), "H": step.H.clone(), "prob": step.prob.clone(),
                    "loss": step.loss.clone(), "lo": part.lo, "hi": part.hi, "s": step.s.clone(),
                    "nnz": step.graph.nnz}, os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_partition_bounds():
    for n, world in [(10, 2), (11, 4), (3, 8), (50_000_000, 8)]:
        parts = [NodePartition(n, world, r) for r in range(world)]
        assert parts[0].lo == 0 and parts[-1].hi == n
        assert all(a.hi == b.lo for a, b in zip(parts[:-1], parts[1:]))
        assert all(p.n_pad == p.per * world >= n for p in parts)
        cover = sum(p.n_local for p in parts)
        assert cover == n
    per, lo, hi = pair_shard(10, 4, 3)
    assert (per, lo, hi) == (3, 9, 10)


def test_two_rank_gloo_equals_single_process(tmp_path):
    single, part1 = run_step(1, 0)
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    outs = [torch.load(os.path.join(tmp_path, f"rank{r}.pt")) for r in range(2)]
    n = part1.n_global
    assert sum(o["nnz"] for o in outs) == single.graph.nnz
    dZ = torch.cat([o["dZ"] for o in outs])
    assert torch.equal(dZ, single.dZ[:n])
    for o in outs:
        assert torch.equal(o["H"][:n], single.H[:n])
        assert torch.equal(o["s"][:n], single.s[:n])
        assert torch.equal(o["prob"][:single.P], single.prob[:single.P])
        assert torch.equal(o["loss"], single.loss)
