"""Node-partitioned execution of the hot path across the GPUs of one node (NCCL over NVLink).

The reference is single-process; for graphs that do not fit one device the
nodes are split into `world` equal contiguous ranges.  Rank r owns rows [lo, hi) of the symmetric
CSR (columns stay global), their Z / H / s / r / dZ rows and an equal share of every pair batch.
Per-node arrays are allocated full size ([n_pad, ...], n_pad = world * ceil(N / world)); a kernel
call reads any gathered row and writes only the owned slice (dl_graph.row_base), and the slice is
then exchanged IN PLACE with one all-gather -- no reductions, no atomics, owner computes.  Per
step (forward + backward) the exchange steps are

    all-gather Z -> attention -> all-gather s -> aggregation -> all-gather H -> pair scores
    -> all-gather prob -> decoder backward (owned nodes, all incident pairs) -> all-gather dH
    -> backward pass 1 -> all-gather r -> backward pass 2

Each row's result is produced by the same kernel code walking the same column-sorted row as on one
GPU, so the partitioned result equals the single-GPU result bit for bit.

The orchestration below is backend-agnostic: the product backend (`CudaBackend`) calls the CUDA
kernels; tests inject a CPU backend to exercise partition bounds, padding, pair sharding and the
exchange sequence over gloo with world_size 2.
"""
from __future__ import annotations

from dataclasses import dataclass

import os

import torch
import torch.distributed as dist


@dataclass
class NodePartition:
    n_global: int
    world: int
    rank: int

    @property
    def per(self) -> int:
        return (self.n_global + self.world - 1) // self.world

    @property
    def n_pad(self) -> int:
        return self.per * self.world

    @property
    def lo(self) -> int:
        return min(self.rank * self.per, self.n_global)

    @property
    def hi(self) -> int:
        return min(self.lo + self.per, self.n_global)

    @property
    def n_local(self) -> int:
        return self.hi - self.lo

    def owner_of(self, node: torch.Tensor) -> torch.Tensor:
        return torch.div(node, self.per, rounding_mode="floor")


def owned_entries(src: torch.Tensor, dst: torch.Tensor, part: NodePartition):
    """Directed (local_row, global_col) entries of the symmetrised adjacency whose row this rank
    owns: (s,t) if s is owned, (t,s) if t is owned (duplicates are collapsed by the CSR build).
    [ref: main_disentangled.py:141  adj_sym = adj + adj.t()]"""
    ms = (src >= part.lo) & (src < part.hi)
    mt = (dst >= part.lo) & (dst < part.hi)
    rows = torch.cat([src[ms], dst[mt]]) - part.lo
    cols = torch.cat([dst[ms], src[mt]])
    return rows, cols


def pair_shard(P: int, world: int, rank: int):
    """Contiguous equal share of a pair batch (padded length per rank, [p_lo, p_hi))."""
    per = (P + world - 1) // world
    lo = min(rank * per, P)
    return per, lo, min(lo + per, P)


def all_gather_rows(full: torch.Tensor, part: NodePartition, group=None) -> None:
    """In-place all-gather of the owned row block of a full-size [n_pad, ...] array."""
    if part.world == 1:
        return
    mine = full[part.rank * part.per:(part.rank + 1) * part.per]
    dist.all_gather_into_tensor(full, mine, group=group)


class PeerExchange:
    """All-gather as a push over NVLink peer memory (dl_push_slice, csrc/peer_copy.cu).

    Every full-size array is registered once: its CUDA IPC handle goes round the process group and
    each rank maps the peers' copies.  A gather is then ONE kernel -- the owner writes its slice
    into the same position of every peer's copy -- followed by a barrier (a 1-element all-reduce),
    instead of NCCL's ring all-gather (33.4 ms for 25.6 GB on 8 B200 whatever the settings).
    The ranks must be processes on one node with P2P access between all GPUs; otherwise, or with
    DL_NO_PUSH=1, `available` is False and the caller keeps using NCCL."""

    def __init__(self, world: int, rank: int, device, group=None):
        self.world, self.rank, self.group, self.device = world, rank, group, device
        self.peers = {}          # data_ptr of the local array -> (ctypes pointer array class, [peer base ptrs])
        self._opened = {}        # IPC handle bytes -> base address of the peer's block in this process
        self.available = world > 1 and not os.environ.get("DL_NO_PUSH") and torch.device(device).type == "cuda"
        if self.available:
            self._flag = torch.zeros(1, dtype=torch.float32, device=device)

    def register(self, t: torch.Tensor) -> bool:
        """Collective: every rank registers its copy of the same array.  -> False when the mapping
        is not possible (the exchange then stays on NCCL for every array)."""
        if not self.available:
            return False
        import ctypes
        from ._lib import lib
        mine = None
        try:
            h = t.untyped_storage()._share_cuda_()
            hb = bytes(h[1])
            # torch: [version byte][type byte: b'c' = a cudaMalloc block][64-byte cudaIpcMemHandle_t]
            if len(hb) == 66 and hb[1:2] == b"c":
                hb = hb[2:]
            if len(hb) == 64:
                mine = (hb, int(h[3]) + t.storage_offset() * t.element_size())
        except Exception:                                   # pragma: no cover - allocator without IPC
            mine = None
        objs = [None] * self.world
        dist.all_gather_object(objs, mine, group=self.group)
        bases, ok = [], all(o is not None for o in objs)
        if ok:
            with torch.cuda.device(self.device):
                for r, (hb, off) in enumerate(objs):
                    if r == self.rank:
                        continue
                    base = self._opened.get(hb)             # one cudaMalloc block may hold several arrays
                    if base is None:
                        out = ctypes.c_void_p()
                        # opened with THIS device current: a peer mapping kernels of this device can use
                        if lib().dl_ipc_open(hb, ctypes.byref(out)) != 0 or not out.value:
                            ok = False
                            break
                        base = self._opened[hb] = out.value
                    bases.append(base + off)
        flag = torch.tensor([1.0 if ok else 0.0], device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if flag.item() < 1.0:
            self.available = False
            self.peers.clear()
            return False
        self.peers[t.data_ptr()] = ((ctypes.c_void_p * len(bases)), bases)
        return True

    def peer_array(self, full: torch.Tensor):
        """-> ctypes array of the peers' base pointers of `full` (for the *_push kernels), or None."""
        reg = self.peers.get(full.data_ptr()) if self.available else None
        if reg is None:
            return None
        arr_t, bases = reg
        return arr_t(*bases)

    def barrier(self) -> None:
        dist.all_reduce(self._flag, group=self.group)

    def gather(self, full: torch.Tensor, per: int) -> bool:
        """Push rows [rank*per, (rank+1)*per) of `full` to every peer, then barrier.  -> False if
        `full` was not registered (the caller falls back to NCCL)."""
        reg = self.peers.get(full.data_ptr()) if self.available else None
        if reg is None:
            return False
        from ._lib import check, lib, stream_of
        arr_t, bases = reg
        row_bytes = (full[0].numel() if full.dim() > 1 else 1) * full.element_size()
        off, n_bytes = self.rank * per * row_bytes, per * row_bytes
        if n_bytes % 16 or full.data_ptr() % 16:           # same on every rank: 16-byte vectors only
            return False
        dst = arr_t(*[b + off for b in bases])
        with torch.cuda.device(self.device):
            check(lib().dl_push_slice(full.data_ptr() + off, dst, len(bases), n_bytes, stream_of(self.device)),
                  "dl_push_slice")
        dist.all_reduce(self._flag, group=self.group)      # barrier: every peer's push has completed
        return True


def all_gather_flat(full: torch.Tensor, per: int, rank: int, world: int, group=None) -> None:
    if world == 1:
        return
    dist.all_gather_into_tensor(full, full[rank * per:(rank + 1) * per], group=group)


class CudaBackend:
    """The product backend: every call goes to libdisenlink_b200.so."""

    def __init__(self):
        from . import ops
        from .graph import Graph
        self.ops, self.Graph = ops, Graph

    def build_graph(self, src, dst, part: NodePartition):
        from ._lib import check, lib, ptr, stream_of
        rows, cols = owned_entries(src, dst, part)
        dev = src.device
        E = int(rows.numel())
        L = lib()
        with torch.cuda.device(dev):
            rowptr = torch.empty(part.n_local + 1, dtype=torch.int64, device=dev)
            col = torch.empty(max(E, 1), dtype=torch.int32, device=dev)
            meta = torch.zeros(2, dtype=torch.int64, device=dev)
            ws_bytes = L.dl_csr_build_workspace_bytes(E, part.n_local)
            ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
            check(L.dl_csr_build_rect(ptr(rows.contiguous()), ptr(cols.contiguous()), E, part.n_local,
                                      part.n_global, 0, ptr(rowptr), ptr(col), ptr(meta),
                                      meta[1:].data_ptr(), ptr(ws), ws_bytes, stream_of(dev)),
                  "dl_csr_build_rect")
            nnz = int(meta[0].item())
            del ws
            col = col[:nnz].clone()
        return self.Graph(rowptr, col, part.n_local, row_base=part.lo, n_global=part.n_pad)

    def build_pairs(self, u, v, part: NodePartition):
        """-> (local shard PairBatch for scoring, incidence graph of owned nodes over ALL pairs,
        inc_pair)."""
        from ._lib import check, lib, ptr, stream_of
        ops = self.ops
        P = int(u.numel())
        per, p_lo, p_hi = pair_shard(P, part.world, part.rank)
        shard = ops.PairBatch(u[p_lo:p_hi], v[p_lo:p_hi], part.n_pad)
        dev = u.device
        u32, v32 = u.to(torch.int32).contiguous(), v.to(torch.int32).contiguous()
        L = lib()
        with torch.cuda.device(dev):
            inc_ptr = torch.empty(part.n_local + 1, dtype=torch.int64, device=dev)
            inc_other = torch.empty(max(2 * P, 1), dtype=torch.int32, device=dev)
            inc_pair = torch.empty(max(2 * P, 1), dtype=torch.int32, device=dev)
            ws_bytes = L.dl_pair_incidence_workspace_bytes(P, part.n_local + 1)
            ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
            check(L.dl_pair_incidence_range(ptr(u32), ptr(v32), P, part.lo, part.hi, ptr(inc_ptr),
                                            ptr(inc_other), ptr(inc_pair), ptr(ws), ws_bytes,
                                            stream_of(dev)), "dl_pair_incidence_range")
            m = int(inc_ptr[-1].item())
            del ws
            inc_other, inc_pair = inc_other[:m].clone(), inc_pair[:m].clone()
        inc = self.Graph(inc_ptr, inc_other, part.n_local, row_base=part.lo, n_global=part.n_pad)
        return shard, inc, inc_pair

    def entry_scratch(self, g):
        return self.ops._x_scratch(g)

    # -- kernels (all write only the owned rows of the full-size outputs) --
    def edge_attn_fwd(self, g, Z, T, kstar, w, s, peers=None):
        if peers is None:
            self.ops.edge_attn_fwd(g, Z, T, out=(kstar, w, s))
            return
        from ._lib import check, lib, ptr, stream_of
        K, d = int(Z.shape[1]), int(Z.shape[2])
        dev = Z.device
        with torch.cuda.device(dev):          # attention + row sums with the all-gather of s fused in
            check(lib().dl_edge_attn_fwd_push(g.ref, ptr(Z), K, d, float(T), ptr(kstar), ptr(w), ptr(s),
                                              ptr(g.hub_scratch(K)), peers, len(peers), stream_of(dev)),
                  "dl_edge_attn_fwd_push")

    def factor_spmm_fwd(self, g, Z, kstar, w, s, beta, H, sj=None, zs=None, peers=None):
        if peers is None:
            self.ops.factor_spmm_fwd(g, Z, kstar, w, s, beta, out=H, sj=sj, zs=zs)
            return
        from ._lib import check, lib, ptr, stream_of
        K, d = int(Z.shape[1]), int(Z.shape[2])
        dev = Z.device
        with torch.cuda.device(dev):          # aggregation with the all-gather of H fused in
            check(lib().dl_factor_spmm_fwd_push(g.ref, ptr(Z), ptr(kstar), ptr(w), ptr(s), K, d, float(beta),
                                                self.ops.one_minus(beta), ptr(H), ptr(sj) if sj is not None else None,
                                                ptr(g.hub_scratch(K * d)), peers, len(peers), stream_of(dev)),
                  "dl_factor_spmm_fwd_push")

    def pair_score_fwd(self, Z, H, shard, T, prob_slice):
        self.ops.pair_score_fwd(Z, H, shard, T, out=(None, prob_slice))

    def pair_score_bwd(self, inc, inc_pair, Z, H, dS, T, dZ, dH, peers=None):
        from ._lib import check, lib, ptr, stream_of
        K, d = int(Z.shape[1]), int(Z.shape[2])
        dev = Z.device
        with torch.cuda.device(dev):
            if peers is not None:             # decoder backward with the all-gather of dH fused in
                check(lib().dl_pair_score_bwd_push(inc.ref, ptr(inc_pair), ptr(Z), ptr(H), ptr(dS), K, d, float(T),
                                                   ptr(dZ), ptr(dH), ptr(inc.hub_scratch(2 * K * d)), peers,
                                                   len(peers), stream_of(dev)), "dl_pair_score_bwd_push")
                return
            check(lib().dl_pair_score_bwd(inc.ref, ptr(inc_pair), ptr(Z), ptr(H), ptr(dS), K, d, float(T),
                                          ptr(dZ), ptr(dH), ptr(inc.hub_scratch(2 * K * d)),
                                          stream_of(dev)), "dl_pair_score_bwd")

    def link_bce(self, prob, labels, weights, dS):
        """-> loss (0-dim tensor); dS [P] written in place."""
        loss, _ = self.ops.link_bce(prob, labels, weights, want_grad=True, dS=dS)
        return loss

    def factor_bwd_gather(self, g, Z, G, kstar, w, s, beta, dZ, r, peers=None, x=None):
        """-> True when x was filled (backward pass 1; with peers the all-gather of r is fused in)."""
        return self.ops.factor_bwd_gather(g, Z, G, kstar, w, s, beta, dZ, r, x=x, peers=peers)

    def factor_bwd_edges(self, g, Z, G, kstar, w, s, r, beta, T, dZ, sj=None, x=None):
        self.ops.factor_bwd_edges(g, Z, G, kstar, w, s, r, beta, T, dZ, sj=sj, x=x)


class PartitionedLinkStep:
    """One forward+backward pass of the hot path on a node-partitioned graph.

    Buffers are allocated once; `run(Z_full)` expects the owned rows of Z_full to hold this rank's
    factor embeddings and leaves dL/dZ of the owned rows in `self.dZ[lo:hi]`."""

    def __init__(self, src, dst, n_global, u, v, labels, weights, K, d, beta, T,
                 world=1, rank=0, group=None, backend=None, device=None, mark=None):
        self.mark = mark if mark is not None else (lambda name: None)  # phase boundary hook (bench)
        self.part = NodePartition(int(n_global), int(world), int(rank))
        self.group = group
        self.be = backend if backend is not None else CudaBackend()
        self.K, self.d, self.beta, self.T = int(K), int(d), float(beta), float(T)
        dev = device if device is not None else src.device
        self.device = dev
        part = self.part
        self.graph = self.be.build_graph(src, dst, part)
        self.P = int(u.numel())
        self.p_per, self.p_lo, self.p_hi = pair_shard(self.P, part.world, part.rank)
        self.shard, self.inc, self.inc_pair = self.be.build_pairs(u, v, part)
        P_pad = self.p_per * part.world
        f32 = dict(dtype=torch.float32, device=dev)
        self.labels = torch.zeros(P_pad, **f32)
        self.labels[:self.P] = labels
        self.weights = torch.zeros(P_pad, **f32)
        self.weights[:self.P] = weights
        nnz = self.graph.nnz
        self.kstar = torch.empty(max(nnz, 1), dtype=torch.uint8, device=dev)
        self.w = torch.empty(max(nnz, 1), **f32)
        # one rank: the aggregation gathers slices pre-divided by s (scratch = the dH buffer, idle
        # during the forward); several ranks: s[col, kstar] per local entry goes forward -> pass 2
        self.prescale = (part.world == 1) and not (getattr(self.graph, "flags", 0) & 8)    # _lib.DL_F_NO_PRESCALE
        self.sj = None if self.prescale else torch.empty(max(nnz, 1), **f32)
        # <G[j,k*], Z[i,k*]> per entry, pass 1 -> pass 2 (the graph's per-entry scratch, shared with the
        # symmetric attention's packed records, which are dead by then)
        self.x = self.be.entry_scratch(self.graph) if hasattr(self.be, "entry_scratch") else None
        self.s = torch.ones(part.n_pad, K, **f32)
        self.r = torch.zeros(part.n_pad, K, **f32)
        self.H = torch.zeros(part.n_pad, K, d, **f32)
        self.dZ = torch.zeros(part.n_pad, K, d, **f32)
        self.dH = torch.zeros(part.n_pad, K, d, **f32)
        self.prob = torch.zeros(P_pad, **f32)
        self.dS = torch.zeros(P_pad, **f32)
        self.loss = None
        # exchange: push over NVLink peer memory when the ranks can map each other's buffers
        self.px = PeerExchange(part.world, part.rank, dev, group) if isinstance(self.be, CudaBackend) else None
        if self.px is not None and self.px.available:
            for t in (self.s, self.H, self.prob, self.dH, self.r):
                if not self.px.register(t):
                    break

    def register_input(self, Z) -> bool:
        """Collective, optional: map the peers' copies of the caller's Z so that its all-gather is
        pushed too (otherwise Z goes through NCCL)."""
        return bool(self.px is not None and self.px.available and self.px.register(Z))

    def _fused_peers(self, full):
        """Peer pointers for a kernel that pushes its own output (None: exchange after the kernel)."""
        if self.px is None or os.environ.get("DL_NO_FUSED_PUSH"):
            return None
        return self.px.peer_array(full)

    def _gather_rows(self, full):
        if not (self.px is not None and self.px.gather(full, self.part.per)):
            all_gather_rows(full, self.part, self.group)

    def _gather_flat(self, full, per):
        if not (self.px is not None and self.px.gather(full, per)):
            all_gather_flat(full, per, self.part.rank, self.part.world, self.group)

    def forward(self, Z):
        part, be, g, mark = self.part, self.be, self.graph, self.mark
        mark("begin")
        self._gather_rows(Z)
        mark("ag_Z")
        sp = self._fused_peers(self.s)
        if sp is not None:
            be.edge_attn_fwd(g, Z, self.T, self.kstar, self.w, self.s, sp)
            mark("attn_fwd")
            self.px.barrier()
        else:
            be.edge_attn_fwd(g, Z, self.T, self.kstar, self.w, self.s)
            mark("attn_fwd")
            self._gather_rows(self.s)
        mark("ag_s")
        hp = self._fused_peers(self.H)
        if hp is not None:                    # the exchange of H rides on the kernel: only a barrier follows
            be.factor_spmm_fwd(g, Z, self.kstar, self.w, self.s, self.beta, self.H, self.sj, None, hp)
            mark("spmm_fwd")
            self.px.barrier()
        else:
            be.factor_spmm_fwd(g, Z, self.kstar, self.w, self.s, self.beta, self.H, self.sj,
                               self.dH if self.prescale else None)
            mark("spmm_fwd")
            self._gather_rows(self.H)
        mark("ag_H")
        lo = part.rank * self.p_per
        if self.p_hi > self.p_lo:
            be.pair_score_fwd(Z, self.H, self.shard, self.T, self.prob[lo:lo + (self.p_hi - self.p_lo)])
        mark("pair_fwd")
        self._gather_flat(self.prob, self.p_per)
        mark("ag_prob")
        return self.H, self.prob

    def loss_and_grad_logit(self):
        """weighted BCE over all pairs and dL/dlogit (every rank evaluates the full P-vector; it
        is P floats), with F.binary_cross_entropy's clamps (see dl_link_bce)."""
        self.loss = self.be.link_bce(self.prob, self.labels, self.weights, self.dS)
        self.mark("loss")
        return self.loss

    def backward(self, Z):
        part, be, g, mark = self.part, self.be, self.graph, self.mark
        dp = self._fused_peers(self.dH)
        if dp is not None:
            be.pair_score_bwd(self.inc, self.inc_pair, Z, self.H, self.dS, self.T, self.dZ, self.dH, dp)
            mark("pair_bwd")
            self.px.barrier()
        else:
            be.pair_score_bwd(self.inc, self.inc_pair, Z, self.H, self.dS, self.T, self.dZ, self.dH)
            mark("pair_bwd")
            self._gather_rows(self.dH)
        mark("ag_dH")
        rp = self._fused_peers(self.r)
        if rp is not None:
            xv = be.factor_bwd_gather(g, Z, self.dH, self.kstar, self.w, self.s, self.beta, self.dZ, self.r, rp,
                                      x=self.x)
            mark("bwd_gather")
            self.px.barrier()
        else:
            xv = be.factor_bwd_gather(g, Z, self.dH, self.kstar, self.w, self.s, self.beta, self.dZ, self.r,
                                      x=self.x)
            mark("bwd_gather")
            self._gather_rows(self.r)
        mark("ag_r")
        be.factor_bwd_edges(g, Z, self.dH, self.kstar, self.w, self.s, self.r, self.beta, self.T, self.dZ, self.sj,
                            x=self.x if xv else None)
        mark("bwd_edges")
        return self.dZ

    def run(self, Z):
        self.forward(Z)
        self.loss_and_grad_logit()
        return self.backward(Z)
