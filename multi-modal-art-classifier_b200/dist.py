"""Multi-GPU execution of the path (nothing like it exists in the reference, which is
single-process): one process per GPU, ``torch.distributed`` (NCCL over NVLink / NVSwitch) for the
plumbing.

GNN: the graph is partitioned **by destination node**.  Every node type is cut into ``world``
contiguous id ranges; rank p owns the destination rows of its ranges and every edge that points at
them, so the per-row neighbour lists -- and their order -- are those of the single-GPU CSR.  Source
rows owned by another rank are *boundary* ("halo") rows: before each conv layer every rank packs
the rows some other rank needs (``agx_pack_rows``) and one NCCL **all-gather** per node type appends
all ranks' boundary rows to the local feature table; in the backward pass the gradient of that
gathered region is **reduce-scattered** back to the owners (``agx_unpack_rows_add``).  Weights are
replicated; their gradients (one flat arena, ``FlatAdam``) take one **all-reduce**; BatchNorm
statistics over the rows of all ranks take two small float64 all-reduces per layer
(``agx_bn_forward_phase``), the loss one.  With these the N-GPU step is arithmetically the
single-GPU step on the whole graph.

For the block-diagonal N-times replicated graph of BASELINE config 5 partitioned block by block no
edge is cut, every boundary list is empty and no feature rows move.

Heads: batch-sharded data parallel, one all-reduce of the gradient arena per step, and the class
weighted CE normaliser (sum of w_y) all-reduced so the weighted mean equals the single-GPU one
(``functional.cross_entropy(..., group=)``).

``GraphPartition`` is host-side index arithmetic in plain torch (runs on CPU or CUDA tensors, no
kernels): it is covered by world_size-2 ``gloo`` tests on CPU.
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch

from . import ops
from . import _lib as L
from ._lib import check, lib, ptr, stream_ptr


def split_bounds(n: int, world: int) -> List[int]:
    """``world + 1`` boundaries of contiguous ranges whose sizes differ by at most one."""
    base, rem = divmod(int(n), int(world))
    out = [0]
    for q in range(world):
        out.append(out[-1] + base + (1 if q < rem else 0))
    return out


def balanced_bounds(edge_index_dict, num_nodes: Dict[str, int], world: int) -> Dict[str, List[int]]:
    """Contiguous destination ranges per node type that carry about the same number of INCOMING
    edges each (the aggregation work of a rank is its edges, not its rows: with Zipf-distributed
    tags or artists an equal-rows split leaves one rank with most of the edges).  Types without
    incoming edges are split by rows."""
    deg = {t: None for t in num_nodes}
    for (s, r, d), ei in edge_index_dict.items():
        c = torch.bincount(ei[1], minlength=int(num_nodes[d])).to(torch.int64)
        deg[d] = c if deg[d] is None else deg[d] + c
    out = {}
    for t, n in num_nodes.items():
        if deg[t] is None or int(deg[t].sum()) == 0 or n < world:
            out[t] = split_bounds(n, world)
            continue
        # weight = edges + 1 per row so that empty rows are spread too
        cum = torch.cumsum(deg[t] + 1, 0)
        total = int(cum[-1])
        b = [0]
        for q in range(1, world):
            target = total * q // world
            idx = int(torch.searchsorted(cum, torch.tensor(target, device=cum.device)))
            b.append(min(max(idx, b[-1]), n))
        b.append(n)
        out[t] = b
    return out


class GraphPartition:
    """Destination-node partition of one heterograph, from the point of view of ``rank``.

    Feature table of node type t on this rank ("extended" table):
        rows [0, n_owned[t])                                 the rows this rank owns
        row  n_owned[t] + q * max_boundary[t] + j            the j-th boundary row of rank q
    (the second region exists only if some rank has boundary rows of type t).

    ``edge_index[(s, rel, d)]``: the edges whose destination this rank owns, in their original
    order, row 0 = source index into the extended table of s, row 1 = destination index into the
    owned rows of d.
    """

    def __init__(self, edge_index_dict, num_nodes: Dict[str, int], world: int, rank: int,
                 bounds: Optional[Dict[str, List[int]]] = None, replicated=(), scattered=()):
        self.world, self.rank = int(world), int(rank)
        self.num_nodes = OrderedDict((t, int(n)) for t, n in num_nodes.items())
        # ``replicated`` node types (the tiny ones -- style, genre, field, media,
        # movement, gallery, ...) live on EVERY rank with all their rows:
        #   replicated -> partitioned relation : every source row is local, nothing is exchanged
        #   partitioned -> replicated relation : a rank owns the edges whose SOURCE it owns and
        #       produces PARTIAL neighbour sums for all destination rows (``partial`` relations: the
        #       conv layer all-reduces them, [N_t, F] floats instead of the source rows)
        #   replicated -> replicated relation  : computed by every rank, redundantly
        self.replicated = set(replicated)
        for t in self.replicated:
            if t not in self.num_nodes:
                raise ValueError(f'replicated type {t} is not a node type')
        # ``scattered`` node types (the mid-sized ones: tag, artist) are cut into ``world`` equal
        # chunks of ceil(N / world) rows like any partitioned type -- their dense work (products,
        # BatchNorm, dropout) is done once, by the owner -- but a relation INTO them from a plain
        # partitioned type (artwork -> tag) keeps its edges with their SOURCE: every rank produces
        # partial neighbour sums for all N rows and one **reduce-scatter** hands each owner the full
        # sums of its chunk (half the volume of the replicated types' all-reduce, no redundant
        # dense work).  Their rows reach the ranks that need them as sources (tag -> artwork) through
        # the boundary-row all-gather like those of every partitioned type.
        self.scattered = set(scattered)
        for t in self.scattered:
            if t not in self.num_nodes or t in self.replicated:
                raise ValueError(f'scattered type {t} must be a non-replicated node type')
        self.chunk = {t: -(-self.num_nodes[t] // world) for t in self.scattered}
        self.bounds = {t: list(bounds[t]) if bounds and t in bounds else split_bounds(n, world)
                       for t, n in self.num_nodes.items()}
        for t in self.scattered:
            self.bounds[t] = [min(q * self.chunk[t], self.num_nodes[t]) for q in range(world + 1)]
        for t, b in self.bounds.items():
            if len(b) != world + 1 or b[0] != 0 or b[-1] != self.num_nodes[t] or \
                    any(b[i] > b[i + 1] for i in range(world)):
                raise ValueError(f'bounds of {t} must be {world + 1} ascending ids from 0 to N')
        dev = next(iter(edge_index_dict.values())).device
        bt = {t: torch.tensor(b, dtype=torch.int64, device=dev) for t, b in self.bounds.items()}

        def owner(t, ids):
            return torch.bucketize(ids, bt[t][1:], right=True)

        # 1. boundary rows: sources of edges whose destination lives on another rank
        is_b = {t: torch.zeros(n, dtype=torch.bool, device=dev) for t, n in self.num_nodes.items()}
        own_dst = {}
        rep = self.replicated
        self.partial: "OrderedDict[tuple, torch.Tensor]" = OrderedDict()
        self.scatter: "OrderedDict[tuple, torch.Tensor]" = OrderedDict()
        sc = self.scattered
        for (s, r, d), ei in edge_index_dict.items():
            if ei.numel() and (int(ei[0].max()) >= self.num_nodes[s] or int(ei[0].min()) < 0 or
                               int(ei[1].max()) >= self.num_nodes[d] or int(ei[1].min()) < 0):
                raise IndexError('edge_index contains node ids outside [0, num_nodes)')
            if d in rep:
                if s in rep:                      # every rank computes the whole relation
                    own_dst[(s, r, d)] = torch.full_like(ei[1], rank)
                else:                             # edges follow their SOURCE; sums are partial
                    own_dst[(s, r, d)] = owner(s, ei[0])
                    # max(in-degree over ALL ranks' edges, 1): the divisor of scatter-mean
                    self.partial[(s, r, d)] = torch.bincount(
                        ei[1], minlength=self.num_nodes[d]).clamp(min=1).to(torch.float32)
                continue
            if d in sc and s not in rep and s not in sc:
                # edges follow their SOURCE; destination ids stay global, in a table of
                # world * chunk rows (the reduce-scatter's equal pieces; the tail rows are empty)
                own_dst[(s, r, d)] = owner(s, ei[0])
                self.scatter[(s, r, d)] = torch.bincount(
                    ei[1], minlength=world * self.chunk[d]).clamp(min=1).to(torch.float32)
                continue
            od = owner(d, ei[1])
            own_dst[(s, r, d)] = od
            if s not in rep:
                is_b[s][ei[0][owner(s, ei[0]) != od]] = True

        # 2. slots of the boundary rows inside their owner's packed send buffer
        self.n_owned: Dict[str, int] = {}
        self.n_ext: Dict[str, int] = {}
        self.max_boundary: Dict[str, int] = {}
        self.n_boundary: Dict[str, List[int]] = {}
        self.boundary_idx: Dict[str, torch.Tensor] = {}
        self.ext_global: Dict[str, torch.Tensor] = {}
        ext_of_global = {}
        for t, n in self.num_nodes.items():
            if t in rep:                          # all rows, in global order, no boundary region
                self.n_boundary[t] = [0] * world
                self.max_boundary[t] = 0
                self.n_owned[t] = self.n_ext[t] = n
                self.boundary_idx[t] = torch.zeros(0, dtype=torch.int32, device=dev)
                ext_of_global[t] = torch.arange(n, dtype=torch.int64, device=dev)
                self.ext_global[t] = ext_of_global[t]
                continue
            lo, hi = self.bounds[t][rank], self.bounds[t][rank + 1]
            csum0 = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev),
                               torch.cumsum(is_b[t].to(torch.int64), 0)])
            at_bounds = csum0[bt[t]]
            cnt = (at_bounds[1:] - at_bounds[:-1]).tolist()
            B = max(cnt) if cnt else 0
            self.n_boundary[t] = [int(c) for c in cnt]
            self.max_boundary[t] = int(B)
            self.n_owned[t] = hi - lo
            self.n_ext[t] = (hi - lo) + (world * B if B > 0 else 0)
            self.boundary_idx[t] = torch.nonzero(is_b[t][lo:hi]).flatten().to(torch.int32)
            ids = torch.arange(n, dtype=torch.int64, device=dev)
            own = owner(t, ids)
            slot = csum0[1:] - 1 - at_bounds[own]
            ext_of_global[t] = torch.where(own == rank, ids - lo, (hi - lo) + own * B + slot)
            eg = torch.full((self.n_ext[t],), -1, dtype=torch.int64, device=dev)
            eg[:hi - lo] = ids[lo:hi]
            if B > 0:
                bi = torch.nonzero(is_b[t]).flatten()
                eg[(hi - lo) + own[bi] * B + slot[bi]] = bi
            self.ext_global[t] = eg

        # 3. the edges this rank owns, re-indexed, original order kept
        self.edge_index = OrderedDict()
        for (s, r, d), ei in edge_index_dict.items():
            m = own_dst[(s, r, d)] == rank
            dst0 = 0 if (d in rep or (s, r, d) in self.scatter) else self.bounds[d][rank]
            self.edge_index[(s, r, d)] = torch.stack(
                [ext_of_global[s][ei[0][m]], ei[1][m] - dst0], dim=0).contiguous()

    @property
    def has_halo(self) -> bool:
        return any(b > 0 for b in self.max_boundary.values())

    def owned(self, t: str, x_global: torch.Tensor) -> torch.Tensor:
        """This rank's rows of a per-node tensor of type t (features, labels): all rows of a
        replicated type."""
        if t in self.replicated:
            return x_global
        return x_global[self.bounds[t][self.rank]:self.bounds[t][self.rank + 1]]

    def halo_rows(self) -> int:
        """Rows this rank receives per exchange (all types)."""
        return sum(self.world * b for b in self.max_boundary.values())


class _HaloFn(torch.autograd.Function):
    """x_owned [n, F] -> extended table [n + world * B, F] (pack, all-gather); backward:
    reduce-scatter of the gathered region's gradient + scatter-add into the owned rows."""

    @staticmethod
    def forward(ctx, x, idx, B, group):
        import torch.distributed as dist
        world = dist.get_world_size(group)
        n, F = x.shape
        x = x.contiguous()
        ext = torch.empty(n + world * B, F, dtype=torch.float32, device=x.device)
        ext[:n].copy_(x)
        send = ops.zeros((B, F), x.device) if idx.numel() < B else \
            torch.empty(B, F, dtype=torch.float32, device=x.device)
        check(lib().agx_pack_rows(ptr(x), x.stride(0), ptr(idx), idx.numel(), F, ptr(send),
                                  stream_ptr()), 'agx_pack_rows')
        dist.all_gather_into_tensor(ext[n:], send, group=group)
        ctx.save_for_backward(idx)
        ctx.meta = (n, F, B, group)
        return ext

    @staticmethod
    def backward(ctx, d_ext):
        import torch.distributed as dist
        (idx,) = ctx.saved_tensors
        n, F, B, group = ctx.meta
        d_ext = d_ext.contiguous()
        dx = d_ext[:n].clone()
        recv = torch.empty(B, F, dtype=torch.float32, device=d_ext.device)
        dist.reduce_scatter_tensor(recv, d_ext[n:], group=group)
        check(lib().agx_unpack_rows_add(ptr(dx), dx.stride(0), ptr(idx), idx.numel(), F, ptr(recv),
                                        stream_ptr()), 'agx_unpack_rows_add')
        return dx, None, None, None


class HaloExchange:
    """Runtime side of a ``GraphPartition``: extends owned feature tables by the boundary rows of
    all ranks.  Input features that do not change between steps are exchanged once (cached on
    tensor identity + version)."""

    def __init__(self, part: GraphPartition, group, device):
        self.group = group
        self.n_owned = dict(part.n_owned)
        self.n_ext = dict(part.n_ext)
        self.max_boundary = dict(part.max_boundary)
        self.idx = {t: part.boundary_idx[t].to(device=device, dtype=torch.int32).contiguous()
                    for t in part.boundary_idx}
        self._cache: Dict[str, tuple] = {}
        self.rows_exchanged = 0

    def extend(self, x_dict, cache: bool = False):
        out = OrderedDict()
        for t, x in x_dict.items():
            B = self.max_boundary.get(t, 0)
            if B == 0:
                out[t] = x
                continue
            if x.shape[0] != self.n_owned[t]:
                raise ValueError(f"x['{t}'] has {x.shape[0]} rows, this rank owns {self.n_owned[t]}")
            if cache and not x.requires_grad:
                key = (x.data_ptr(), tuple(x.shape), x._version)
                hit = self._cache.get(t)
                if hit is not None and hit[0] == key:
                    out[t] = hit[1]
                    continue
            ext = _HaloFn.apply(x, self.idx[t], B, self.group)
            self.rows_exchanged += ext.shape[0] - x.shape[0]
            if cache and not x.requires_grad:
                self._cache[t] = (key, ext)
            out[t] = ext
        return out


    @torch.no_grad()
    def refresh(self, x_dict):
        """New values were copied into cached (static) input tables: redo their exchange INTO the
        cached extended tables (addresses unchanged, so a captured CUDA graph reads the new rows)."""
        for t, x in x_dict.items():
            hit = self._cache.get(t)
            if hit is None or self.max_boundary.get(t, 0) == 0:
                continue
            key = (x.data_ptr(), tuple(x.shape), x._version)
            if hit[0] != key:
                hit[1].copy_(_HaloFn.apply(x, self.idx[t], self.max_boundary[t], self.group))
                self._cache[t] = (key, hit[1])


@dataclass
class DistContext:
    """What the hetero module and the trainers need to run one rank of a multi-GPU job."""
    group: object
    rank: int
    world: int
    num_nodes_global: Dict[str, int]            # rows of every node type over all ranks
    halo: Optional[HaloExchange] = None
    _counts: Optional[dict] = None
    replicated: frozenset = frozenset()         # node types every rank holds completely
    partial: Optional[dict] = None              # edge type -> global max(in-degree, 1) [N_dst]
    scatter: Optional[dict] = None              # edge type -> the same over world * chunk rows
    n_owned: Optional[dict] = None              # node type -> rows this rank owns

    def __deepcopy__(self, memo):           # process groups are not copyable; share the context
        return self

    def plan_rows(self, num_nodes: Dict[str, int]):
        """(rows of every type's table as a SOURCE, rows as a DESTINATION or None) for graph.get_plan
        given the rows this rank holds: source tables carry the boundary rows behind the owned
        ones; a scatter relation produces sums for the destination rows of ALL ranks."""
        num_dst = None
        if self.halo is not None:
            num_dst, num_nodes = num_nodes, {t: self.halo.n_ext.get(t, n)
                                             for t, n in num_nodes.items()}
        if self.scatter:
            num_dst = dict(num_dst if num_dst is not None else num_nodes)
            num_dst.update({et: int(c.shape[0]) for et, c in self.scatter.items()})
        return num_nodes, num_dst

    def counts(self, types, device) -> torch.Tensor:
        """float64 [len(types)] global row counts (BatchNorm over the rows of all ranks)."""
        if self._counts is None:
            self._counts = {}
        key = (tuple(types), str(device))
        c = self._counts.get(key)
        if c is None:
            c = torch.tensor([float(self.num_nodes_global[t]) for t in types], dtype=torch.float64,
                             device=device)
            self._counts[key] = c
        return c


def block_context(group, num_nodes_block: Dict[str, int]) -> DistContext:
    """Context of the block-diagonal replicated graph, one block per rank (no edge is cut)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    return DistContext(group, dist.get_rank(group), world,
                       {t: int(n) * world for t, n in num_nodes_block.items()}, None)


def partition_context(part: GraphPartition, group, device) -> DistContext:
    halo = HaloExchange(part, group, device) if part.has_halo else None
    return DistContext(group, part.rank, part.world, dict(part.num_nodes), halo,
                       replicated=frozenset(part.replicated),
                       partial={k: v.to(device) for k, v in part.partial.items()},
                       scatter={k: v.to(device) for k, v in part.scatter.items()},
                       n_owned=dict(part.n_owned))


class PeerAllReduce:
    """In-place sum all-reduce of SMALL float32 / float64 tensors by one agx kernel over NVLink peer
    memory (agx_peer_allreduce, csrc/agx_comm.cu) instead of an NCCL call: the BatchNorm-statistic,
    loss and label-normaliser reductions of a multi-GPU step (4 B .. 18 KB) and the heads' gradient
    arena (180 KB) are purely latency-bound.  The symmetric buffers are allocated and exchanged by
    ``torch.distributed._symmetric_memory`` (plumbing); larger tensors keep using NCCL."""

    MAX_BYTES = 256 * 1024

    def __init__(self, group, device):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        nbytes = int(lib().agx_peer_allreduce_buffer_bytes(self.MAX_BYTES, self.world))
        self.buf = symm.empty(nbytes, dtype=torch.uint8, device=device)
        self.buf.zero_()
        self.handle = symm.rendezvous(self.buf, group)
        self.ptrs = torch.tensor([int(p) for p in self.handle.buffer_ptrs], dtype=torch.int64,
                                 device=device)
        self.epoch = torch.zeros(4, dtype=torch.int64, device=device)
        torch.cuda.synchronize(device)
        dist.barrier(group)                 # every rank has zeroed its flags before the first call
        self.calls = 0

    def fits(self, t: torch.Tensor) -> bool:
        return t.is_cuda and t.is_contiguous() and t.dtype in (torch.float32, torch.float64) and \
            0 < t.numel() * t.element_size() <= self.MAX_BYTES

    def __call__(self, t: torch.Tensor) -> torch.Tensor:
        dt = L.F64 if t.dtype == torch.float64 else L.F32
        check(lib().agx_peer_allreduce(ptr(self.ptrs), self.rank, self.world, ptr(t), ptr(t),
                                       t.numel(), dt, ptr(self.epoch), self.MAX_BYTES, stream_ptr()),
              'agx_peer_allreduce')
        self.calls += 1
        return t


_PEER: Dict[int, object] = {}          # id(group) -> PeerAllReduce | False (set-up failed)
PEER_STATUS: Dict[str, str] = {}


def enable_peer_allreduce(group, device) -> bool:
    """Set up the peer-memory all-reduce for ``group`` (collective call: every rank of the group).
    Returns False -- and the small reductions stay on NCCL -- when symmetric memory cannot be
    established on this machine (the reason is kept in ``dist.PEER_STATUS``)."""
    import os
    import torch.distributed as dist
    key = id(group)
    if key in _PEER:
        return bool(_PEER[key])
    ok = torch.ones(1, dtype=torch.int32, device=device)
    par = None
    if os.environ.get('AGX_PEER_ALLREDUCE', '1') == '0':
        ok.zero_()
        PEER_STATUS['reason'] = 'disabled by AGX_PEER_ALLREDUCE=0'
    else:
        try:
            par = PeerAllReduce(group, device)
        except Exception as e:  # noqa: BLE001  (no P2P / fabric handles: NCCL keeps doing the job)
            ok.zero_()
            PEER_STATUS['reason'] = f'{type(e).__name__}: {e}'[:300]
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)     # all ranks or none
    if int(ok.item()) != 1:
        par = None
    _PEER[key] = par if par is not None else False
    PEER_STATUS['mode'] = 'peer-memory kernel (agx_peer_allreduce)' if par is not None else 'nccl'
    return par is not None


def small_all_reduce_(t: torch.Tensor, group) -> torch.Tensor:
    """Sum all-reduce in place: the peer-memory kernel when it is set up for ``group`` and the
    tensor is small, NCCL otherwise."""
    par = _PEER.get(id(group))
    if par and par.fits(t):
        return par(t)
    import torch.distributed as dist
    dist.all_reduce(t, group=group)
    return t


def reduce_scatter_rows(own: torch.Tensor, full: torch.Tensor, group):
    """own [chunk, F] <- this rank's piece of the sum over the ranks of full [world * chunk, F];
    returns the async work handle (NCCL ``ncclReduceScatter`` on the communicator's stream).  The
    ``gloo`` backend of the CPU test-suite has no reduce-scatter: all-reduce + slice there."""
    import torch.distributed as dist
    if dist.get_backend(group) == 'gloo':
        tmp = full.clone()
        dist.all_reduce(tmp, group=group)
        r = dist.get_rank(group)
        own.copy_(tmp[r * own.shape[0]:(r + 1) * own.shape[0]])
        return None
    return dist.reduce_scatter_tensor(own, full, group=group, async_op=True)


def all_reduce_(t: torch.Tensor, group) -> torch.Tensor:
    import torch.distributed as dist
    dist.all_reduce(t, group=group)
    return t


def broadcast_(t: torch.Tensor, group, src: int = 0) -> torch.Tensor:
    import torch.distributed as dist
    dist.broadcast(t, src=dist.get_global_rank(group, src) if group is not None else src,
                   group=group)
    return t
