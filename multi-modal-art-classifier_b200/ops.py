"""Tensor-level wrappers over the C ABI (include/agx.h): every function takes CUDA torch tensors,
passes raw device pointers + the current stream to ``libagx.so`` and returns torch tensors.
torch is used for allocation and stream plumbing only."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib as L
from ._lib import check, lib, ptr, stream_ptr


# ------------------------------------------------------------------------------------------------
# per-kernel-class timing hook (bench.py's roofline leg): CUDA events on the launching stream
# ------------------------------------------------------------------------------------------------
class KernelTimer:
    """When installed as ``ops.TIMER`` every aggregation / GEMM launch group is bracketed by CUDA
    events on the current stream and its algorithmic bytes / flops are recorded."""

    def __init__(self):
        self.records = []           # (kind, algorithmic_bytes, flops, start_event, end_event)

    def begin(self):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(torch.cuda.current_stream())
        return ev

    def end(self, kind, nbytes, flops, start):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(torch.cuda.current_stream())
        self.records.append((kind, nbytes, flops, start, ev))

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for kind, nbytes, flops, s, e in self.records:
            d = out.setdefault(kind, {'launches': 0, 'ms': 0.0, 'bytes': 0, 'flops': 0})
            d['launches'] += 1
            d['ms'] += s.elapsed_time(e)
            d['bytes'] += nbytes
            d['flops'] += flops
        return out


def _summary_min(timers):
    """Per launch position the fastest of the repeated steps (same launch sequence in every
    timer), summed per kernel class: {'kind': {'launches', 'ms', 'bytes', 'flops'}} of ONE step."""
    torch.cuda.synchronize()
    n = min(len(t.records) for t in timers)
    out = {}
    for i in range(n):
        kind, nbytes, flops, _, _ = timers[0].records[i]
        ms = min(t.records[i][3].elapsed_time(t.records[i][4]) for t in timers
                 if t.records[i][0] == kind)
        d = out.setdefault(kind, {'launches': 0, 'ms': 0.0, 'bytes': 0, 'flops': 0})
        d['launches'] += 1
        d['ms'] += ms
        d['bytes'] += nbytes
        d['flops'] += flops
    return out


KernelTimer.summary_min = staticmethod(_summary_min)
TIMER: Optional[KernelTimer] = None
_DEBUG = bool(int(__import__('os').environ.get('AGX_DEBUG', '0')))


# ------------------------------------------------------------------------------------------------
# side stream for INDEPENDENT launches of one fused layer (the edge-balanced aggregation of the
# long-row relations next to the row-parallel one of the others: both leave part of the machine
# idle in their tails).  Under stream capture the branch becomes a parallel path of the graph.
# ------------------------------------------------------------------------------------------------
_NO_SIDE = __import__('os').environ.get('AGX_NO_SIDE_STREAMS') is not None
_SIDE: dict = {}


class Fork:
    """``fk = fork(device)``; ``with fk: <launches for the side stream>``; launches for the main
    stream; ``fk.join()``.  A no-op (everything stays on the current stream) on CPU tensors (the
    test-suite's restated kernels), while per-kernel timing is on, and with AGX_NO_SIDE_STREAMS."""

    def __init__(self, device):
        self.main = self.side = None
        if _NO_SIDE or TIMER is not None or device.type != 'cuda':
            return
        self.main = torch.cuda.current_stream(device)
        side = _SIDE.get(device.index)
        if side is None:
            side = _SIDE[device.index] = torch.cuda.Stream(device)
        self.side = side
        side.wait_stream(self.main)
        self._ctx = None

    def __enter__(self):
        if self.side is not None:
            self._ctx = torch.cuda.stream(self.side)
            self._ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.side is not None:
            self._ctx.__exit__(*exc)
        return False

    def join(self):
        if self.side is not None:
            self.main.wait_stream(self.side)
            self.side = None


def fork(device) -> Fork:
    return Fork(device)


# Work nobody reads before the END of the backward pass (weight gradients that are added straight
# into the optimizer's buffers) goes to a third stream and is NOT joined by the layer that issues
# it: it runs under the BatchNorm backward and the next layer's backward.  ``Deferred.join`` is
# queued as an autograd-engine callback (runs when the backward pass finishes, also inside a stream
# capture), so that after ``loss.backward()`` the caller's stream has waited for everything.  The
# tensors those launches read are kept alive until then (autograd frees a layer's saved tensors and
# incoming gradients when its backward returns; the caching allocator would hand their memory to
# the main stream while the deferred launch still reads it).
_DEFER: dict = {}


class Deferred:
    def __init__(self, device):
        self.stream = torch.cuda.Stream(device)
        self.main = None
        self.keep: list = []
        self.pending = False

    def join(self):
        if self.pending:
            self.main.wait_stream(self.stream)
            self.pending = False
        self.keep.clear()


class _DeferCtx:
    def __init__(self, d, main):
        self.d, self.main = d, main
        self.keep = d.keep            # launches inside add what they must keep alive

    def __enter__(self):
        self.d.stream.wait_stream(self.main)
        self._ctx = torch.cuda.stream(self.d.stream)
        self._ctx.__enter__()
        return self

    def __exit__(self, *exc):
        self._ctx.__exit__(*exc)
        return False


class _NoDefer:
    keep = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def defer(device, keep=()):
    """Context manager: launches inside go to the deferred stream (after everything queued on the
    current stream so far); ``keep``: tensors those launches read / write.  Only valid inside a
    backward pass (the join is an engine callback).  No-op where ``fork`` is one."""
    if _NO_SIDE or TIMER is not None or device.type != 'cuda':
        return _NoDefer()
    d = _DEFER.get(device.index)
    if d is None:
        d = _DEFER[device.index] = Deferred(device)
    main = torch.cuda.current_stream(device)
    if d.pending and d.main is not None and d.main != main:
        d.join()                                  # (another stream took over: settle the old one)
    d.main = main
    d.pending = True
    # one callback per use (join() is idempotent): a flag "already queued for this pass" would
    # survive a backward pass that died with an exception and silence the join of the next one
    torch.autograd.Variable._execution_engine.queue_callback(d.join)
    d.keep.extend(t for t in keep if t is not None)
    return _DeferCtx(d, main)


def aggregation_bytes(n_edges: int, n_rows: int, F: int, elem: int = 4) -> int:
    """Algorithmic HBM bytes of one relation's aggregation pass:
    E*(F*s + 4) gathered rows + column ids, (N+1)*4 row pointers, N*F*s output rows."""
    return n_edges * (F * elem + 4) + (n_rows + 1) * 4 + n_rows * F * elem


# ------------------------------------------------------------------------------------------------
# K1: CSR / CSC
# ------------------------------------------------------------------------------------------------
@dataclass
class CSR:
    """Compressed rows of one relation: ``col[rowptr[i]:rowptr[i+1]]`` are row i's neighbours in
    edge-list order; ``eid`` is the stable permutation; ``cnt = max(degree, 1)`` as float."""
    rowptr: torch.Tensor
    col: torch.Tensor
    eid: torch.Tensor
    cnt: torch.Tensor
    n_rows: int
    n_cols: int
    n_edges: int
    max_degree: int = -1             # filled by HeteroPlan (one read-back at plan time); -1 unknown

    @property
    def avg_degree(self) -> float:
        return self.n_edges / max(self.n_rows, 1)

    @property
    def long_rows(self) -> bool:
        """Few / skewed rows: a warp per row would serialise thousands of gathers behind one warp
        (artwork -> style has 32 rows of ~3.6k edges; Zipf tags reach 40k), so the edge-balanced
        kernel is used instead."""
        return self.avg_degree > LONG_ROW_AVG_DEGREE or self.max_degree > LONG_ROW_MAX_DEGREE


LONG_ROW_AVG_DEGREE = 12.0     # agg_rows stages <= 16 neighbours per output row
LONG_ROW_MAX_DEGREE = 256


def csr_build(edge_lists: Sequence[Tuple[torch.Tensor, torch.Tensor, int, int]],
              check_range: bool = True, buffers: Optional[list] = None) -> List[CSR]:
    """``edge_lists``: (keys int64 [E], vals int64 [E], n_rows, n_cols) per relation.  One batched
    stable radix sort (K1) for all of them.  Raises IndexError for out-of-range indices, like the
    reference's ``index_select`` / ``scatter_add_``.

    ``buffers``: an (initially empty) list that receives the output / workspace tensors of each
    batch; passing the same list again re-sorts INTO those tensors (same addresses: a captured
    CUDA graph that reads the CSR stays valid) and defers the range check to ``buffers`` users
    (``err`` flags are ``buffers[i][4]``)."""
    out: List[CSR] = []
    for bi, base in enumerate(range(0, len(edge_lists), L.MAX_CSR_RELS)):
        reuse = None
        if buffers is not None and bi < len(buffers):
            reuse = buffers[bi]
        res, bufs = _csr_build_batch(edge_lists[base:base + L.MAX_CSR_RELS],
                                     check_range and reuse is None, reuse)
        if buffers is not None and reuse is None:
            buffers.append(bufs)
        out.extend(res)
    return out


def _csr_build_batch(edge_lists, check_range, reuse=None):
    n = len(edge_lists)
    dev = edge_lists[0][0].device
    arr = (L.EdgeList * n)()
    keep = []
    tot_e = tot_r = 0
    for i, (keys, vals, n_rows, n_cols) in enumerate(edge_lists):
        L.require_cuda(keys, 'edge keys')
        if keys.dtype != torch.int64 or vals.dtype != torch.int64:
            raise TypeError('edge_index must be int64')
        keys = keys.contiguous()
        vals = vals.contiguous()
        keep += [keys, vals]
        arr[i] = L.EdgeList(ptr(keys), ptr(vals), keys.numel(), int(n_rows), int(n_cols))
        tot_e += keys.numel()
        tot_r += int(n_rows)
    ws_bytes = lib().agx_csr_workspace_bytes(tot_e, tot_r)
    if reuse is not None:
        rowptr, col, eid, cnt, err, ws = reuse
        if rowptr.numel() != tot_r + n or col.numel() != max(tot_e, 1) or ws.numel() < ws_bytes:
            raise ValueError('csr_build: reused buffers do not match the edge lists')
    else:
        rowptr = torch.empty(tot_r + n, dtype=torch.int32, device=dev)
        col = torch.empty(max(tot_e, 1), dtype=torch.int32, device=dev)
        eid = torch.empty(max(tot_e, 1), dtype=torch.int32, device=dev)
        cnt = torch.empty(max(tot_r, 1), dtype=torch.float32, device=dev)
        err = torch.zeros(1, dtype=torch.int32, device=dev)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    check(lib().agx_csr_build(arr, n, ptr(rowptr), ptr(col), ptr(eid), ptr(cnt), ptr(err), ptr(ws),
                              ws_bytes, stream_ptr()), 'agx_csr_build')
    if check_range and int(err.item()) != 0:
        raise IndexError('edge_index contains node ids outside [0, num_nodes)')
    res = []
    e0 = r0 = 0
    for i, (keys, vals, n_rows, n_cols) in enumerate(edge_lists):
        e, r = keys.numel(), int(n_rows)
        res.append(CSR(rowptr[r0 + i:r0 + i + r + 1], col[e0:e0 + e], eid[e0:e0 + e],
                       cnt[r0:r0 + r], r, int(n_cols), e))
        e0 += e
        r0 += r
    return res, [rowptr, col, eid, cnt, err, ws]


def coalesce_undirected(row: torch.Tensor, col: torch.Tensor, n_nodes: int) -> torch.Tensor:
    """a-2, non-bipartite store: unique (row, col) pairs of the symmetrised list, ascending."""
    L.require_cuda(row, 'edge_index')
    e = row.numel()
    dev = row.device
    out_r = torch.empty(max(2 * e, 1), dtype=torch.int64, device=dev)
    out_c = torch.empty(max(2 * e, 1), dtype=torch.int64, device=dev)
    count = torch.zeros(1, dtype=torch.int64, device=dev)
    ws_bytes = lib().agx_coalesce_workspace_bytes(e)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    row, col = row.contiguous(), col.contiguous()
    check(lib().agx_coalesce_undirected(ptr(row), ptr(col), e, int(n_nodes), ptr(out_r), ptr(out_c),
                                        ptr(count), ptr(ws), ws_bytes, stream_ptr()),
          'agx_coalesce_undirected')
    k = int(count.item())
    return torch.stack([out_r[:k], out_c[:k]], dim=0)


# ------------------------------------------------------------------------------------------------
# K2/K3: aggregation
# ------------------------------------------------------------------------------------------------
@dataclass
class RelArg:
    csr: CSR
    x: torch.Tensor
    mean_rows: bool = False          # divide by the row's own count (forward scatter-mean)
    nbr_scale: Optional[torch.Tensor] = None   # per-neighbour counts (transpose of scatter-mean)
    edge_w: Optional[torch.Tensor] = None      # per-slot weights (GATConv attention coefficients)
    edge_w_idx: Optional[torch.Tensor] = None  # int32: slot e uses edge_w[edge_w_idx[e]]


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return L.F32
    if t.dtype == torch.bfloat16:
        return L.BF16
    raise TypeError(f'unsupported feature dtype {t.dtype}')


def _rel_struct(a: RelArg) -> L.Rel:
    if a.x.stride(-1) != 1:
        raise ValueError('feature rows must be contiguous')
    if a.edge_w is not None and (a.edge_w.dtype != torch.float32 or not a.edge_w.is_contiguous()):
        raise TypeError('edge_w must be contiguous float32')
    if a.edge_w_idx is not None and (a.edge_w_idx.dtype != torch.int32 or
                                     not a.edge_w_idx.is_contiguous()):
        raise TypeError('edge_w_idx must be contiguous int32')
    return L.Rel(ptr(a.csr.rowptr), ptr(a.csr.col), ptr(a.x), a.x.stride(0),
                 ptr(a.csr.cnt) if a.mean_rows else None, ptr(a.nbr_scale), ptr(a.edge_w),
                 ptr(a.edge_w_idx))


def aggregate_rows(groups: Sequence[tuple], F: int):
    """``groups``: (out [n_rows, F], relations, accumulate[, bias [F]]).  One launch per <=24 groups;
    each output row is the sum over the group's relations of the (scaled) neighbour sums."""
    if not groups:
        return
    dt = _dtype_code(groups[0][0])
    if _DEBUG:
        for out, rels, acc, *_ in groups:
            print(f'[agx] agg_rows F={F} rows={out.shape[0]} acc={acc} rels=' + ', '.join(
                f'(E={a.csr.n_edges} avg={a.csr.avg_degree:.1f} max={a.csr.max_degree} '
                f'mean={a.mean_rows} scale={a.nbr_scale is not None} x={tuple(a.x.shape)})'
                for a in rels), flush=True)
    for base in range(0, len(groups), L.MAX_GROUPS):
        part = groups[base:base + L.MAX_GROUPS]
        arr = (L.RowGroup * len(part))()
        for i, (out, rels, acc, *rest) in enumerate(part):
            if len(rels) > L.MAX_REL_PER_GROUP:
                raise ValueError('more than 8 relations in one aggregation group')
            g = arr[i]
            g.out, g.ldo, g.n_rows, g.n_rel, g.accumulate = ptr(out), out.stride(0), out.shape[0], \
                len(rels), int(acc)
            if rest and rest[0] is not None:
                if rest[0].dtype != torch.float32 or rest[0].numel() != F or \
                        not rest[0].is_contiguous():
                    raise ValueError('row-group bias must be contiguous float32 [F]')
                g.bias = ptr(rest[0])
            for j, a in enumerate(rels):
                g.rel[j] = _rel_struct(a)
        t0 = TIMER.begin() if TIMER is not None else None
        check(lib().agx_aggregate_rows(arr, len(part), F, dt, stream_ptr()), 'agx_aggregate_rows')
        if TIMER is not None:
            esz = 4 if dt == L.F32 else 2
            nb = sum(aggregation_bytes(a.csr.n_edges, 0, F, esz) + (g_[0].shape[0] + 1) * 4
                     for g_ in part for a in g_[1])
            nb += sum(g_[0].shape[0] * F * esz for g_ in part)
            TIMER.end('agg_rows', nb, 0, t0)


# Arrival counters of agg_chunks: the kernel needs them zero on entry and leaves them zero, so one
# zero-initialised buffer per device is handed out in slices, round robin (launches on one stream
# are ordered; concurrent launches on different streams get different slices until the ring wraps).
_COUNTER_RING = {}
_COUNTER_RING_INTS = 1 << 20


def _counter_slice(dev: torch.device, n: int) -> torch.Tensor:
    buf, off = _COUNTER_RING.get(dev, (None, 0))
    if buf is None or n > buf.numel():
        buf, off = torch.zeros(max(_COUNTER_RING_INTS, 2 * n), dtype=torch.int32, device=dev), 0
    if off + n > buf.numel():
        off = 0
    _COUNTER_RING[dev] = (buf, off + n)
    return buf[off:off + n]


def aggregate_chunks(segs: Sequence[Tuple[torch.Tensor, RelArg]], F: int):
    """Edge-balanced aggregation for long-row relations; ``segs``: (out [n_rows, F], relation)."""
    if not segs:
        return
    dt = _dtype_code(segs[0][0])
    dev = segs[0][0].device
    for base in range(0, len(segs), L.MAX_CHUNK_SEGS):
        part = segs[base:base + L.MAX_CHUNK_SEGS]
        arr = (L.ChunkSeg * len(part))()
        sizes = [lib().agx_chunk_frag_floats(a.csr.n_edges, F) for _, a in part]
        ncnt = [lib().agx_chunk_counters(a.csr.n_edges) for _, a in part]
        frag = torch.empty(sum(sizes), dtype=torch.float32, device=dev)
        counters = _counter_slice(dev, sum(ncnt))
        off = coff = 0
        for i, (out, a) in enumerate(part):
            s = arr[i]
            s.rel = _rel_struct(a)
            s.out, s.ldo, s.n_rows, s.n_edges = ptr(out), out.stride(0), out.shape[0], a.csr.n_edges
            s.frag = frag.data_ptr() + 4 * off
            s.counters = counters.data_ptr() + 4 * coff
            off += sizes[i]
            coff += ncnt[i]
        if _DEBUG:
            print(f'[agx] agg_chunks F={F} ' + ', '.join(
                f'(rows={out.shape[0]} E={a.csr.n_edges} max={a.csr.max_degree} '
                f'mean={a.mean_rows} scale={a.nbr_scale is not None})' for out, a in part),
                flush=True)
        t0 = TIMER.begin() if TIMER is not None else None
        check(lib().agx_aggregate_chunks(arr, len(part), F, dt, stream_ptr()),
              'agx_aggregate_chunks')
        if TIMER is not None:
            esz = 4 if dt == L.F32 else 2
            TIMER.end('agg_chunks', sum(aggregation_bytes(a.csr.n_edges, out.shape[0], F, esz)
                                        for out, a in part), 0, t0)


def aggregate(out: torch.Tensor, rel: RelArg, F: int):
    """Single relation, picks the row-parallel or the edge-balanced kernel from the mean degree."""
    if rel.csr.long_rows:
        aggregate_chunks([(out, rel)], F)
    else:
        aggregate_rows([(out, [rel], False)], F)


# ------------------------------------------------------------------------------------------------
# GATConv scalar passes (agx_gat.cu)
# ------------------------------------------------------------------------------------------------
@dataclass
class GatArg:
    """One relation of an attention layer; per-edge arrays in CSR order."""
    csr: CSR
    a_l: torch.Tensor                # [n_src]
    a_r: torch.Tensor                # [n_dst]
    alpha: torch.Tensor              # [n_edges]
    dalpha: Optional[torch.Tensor] = None
    de: Optional[torch.Tensor] = None
    da_r: Optional[torch.Tensor] = None
    long_rows: Optional[torch.Tensor] = None   # int32: ALL rows with > GAT_LONG_ROW edges


def gat_long_rows(csr: CSR) -> Optional[torch.Tensor]:
    """The rows of ``csr`` that agx_gat_edge_softmax must be told about (one read-back: call once
    per graph, not per step)."""
    if csr.n_rows == 0 or csr.n_edges <= L.GAT_LONG_ROW:
        return None
    deg = csr.rowptr[1:] - csr.rowptr[:-1]
    rows = torch.nonzero(deg > L.GAT_LONG_ROW).view(-1).to(torch.int32)
    return rows.contiguous() if rows.numel() else None


def gat_edge_softmax(rels: Sequence[GatArg], slope: float, backward: bool = False):
    """alpha = softmax over each destination row of leaky_relu(a_l[src] + a_r[dst]) (forward), or
    de / da_r from alpha and dalpha (backward); <= 24 relations per launch."""
    fn = lib().agx_gat_edge_softmax_bwd if backward else lib().agx_gat_edge_softmax
    what = 'agx_gat_edge_softmax_bwd' if backward else 'agx_gat_edge_softmax'
    for base in range(0, len(rels), L.MAX_GAT_RELS):
        part = rels[base:base + L.MAX_GAT_RELS]
        arr = (L.GatRel * len(part))()
        for i, a in enumerate(part):
            for t in (a.a_l, a.a_r, a.alpha, a.dalpha, a.de, a.da_r):
                if t is not None and (t.dtype != torch.float32 or not t.is_contiguous()):
                    raise TypeError('attention arrays must be contiguous float32')
            if a.a_r.numel() != a.csr.n_rows or a.a_l.numel() != a.csr.n_cols or \
                    a.alpha.numel() < a.csr.n_edges:
                raise ValueError('attention array sizes do not match the CSR')
            if a.long_rows is not None and a.long_rows.dtype != torch.int32:
                raise TypeError('long_rows must be int32')
            arr[i] = L.GatRel(ptr(a.csr.rowptr), ptr(a.csr.col), ptr(a.a_l), ptr(a.a_r),
                              ptr(a.alpha), ptr(a.dalpha), ptr(a.de), ptr(a.da_r),
                              ptr(a.long_rows), a.csr.n_rows,
                              0 if a.long_rows is None else a.long_rows.numel())
        check(fn(arr, len(part), float(slope), stream_ptr()), what)


def sddmm(segs: Sequence[Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]],
          F: int):
    """``segs``: (row int32 [E], col int32 [E], a [*, F], b [*, F], out [E]):
    out[e] = <a[row[e]], b[col[e]]>; <= 24 segments per launch."""
    for base in range(0, len(segs), L.MAX_SDDMM_SEGS):
        part = segs[base:base + L.MAX_SDDMM_SEGS]
        arr = (L.SddmmSeg * len(part))()
        for i, (row, col, a, b, out) in enumerate(part):
            if a.stride(-1) != 1 or b.stride(-1) != 1 or a.shape[1] != F or b.shape[1] != F:
                raise ValueError('sddmm operands must be [*, F] with contiguous rows')
            if row.dtype != torch.int32 or col.dtype != torch.int32 or row.numel() != col.numel() \
                    or out.numel() < row.numel():
                raise ValueError('sddmm index arrays must be int32 of equal length')
            arr[i] = L.SddmmSeg(ptr(row), ptr(col), ptr(a), a.stride(0), ptr(b), b.stride(0),
                                ptr(out), row.numel(), 0)
        check(lib().agx_sddmm(arr, len(part), F, stream_ptr()), 'agx_sddmm')


# ------------------------------------------------------------------------------------------------
# K4: grouped GEMM
# ------------------------------------------------------------------------------------------------
TC_MIN_ROWS = 512        # agx_gemm_tc.cu takes problems with at least this many output rows
TC_MAX_K = 256           # ... and at most this much reduction per launch (accumulator error)


class GemmBatch:
    """Collects problems ``C (+)= sum_s opA_s @ opB_s (+ bias)`` and launches them together.
    Operands are given as 2-D tensor *views* (any strides): opA [M, K], opB [K, N]."""

    def __init__(self):
        self.problems: List[L.GemmProblem] = []
        self.segs: List[L.GemmSeg] = []
        self._keep = []
        self._transposes: list = []
        self._later: list = []        # follow-up batches (K-split tall problems), run in order
        self._wave = None

    def add(self, C_out: torch.Tensor, segs: Sequence[tuple], bias: Optional[torch.Tensor] = None,
            accumulate: bool = False, row_scale: Optional[torch.Tensor] = None,
            split_k: int = 1, skip_flag: Optional[torch.Tensor] = None):
        """``segs``: (opA [M,K], opB [K,N]) or (opA, opB, A_mask, B_mask)."""
        M, N = C_out.shape
        # a tall problem whose segments add up to more than the tensor-core kernel's K budget
        # (256: accumulator error, agx_gemm_tc.cu) is cut into consecutive launches of <= 256 each,
        # the later ones accumulating onto the first
        if (M >= TC_MIN_ROWS and split_k <= 1 and row_scale is None and len(segs) > 1 and
                sum(sg[0].shape[1] for sg in segs) > TC_MAX_K and
                all(sg[0].shape[1] <= TC_MAX_K for sg in segs) and self._wave is None):
            groups, cur, k = [], [], 0
            for sg in segs:
                if cur and k + sg[0].shape[1] > TC_MAX_K:
                    groups.append(cur)
                    cur, k = [], 0
                cur.append(sg)
                k += sg[0].shape[1]
            groups.append(cur)
            self.add(C_out, groups[0], bias=bias, accumulate=accumulate, skip_flag=skip_flag)
            for gi, grp in enumerate(groups[1:]):
                while len(self._later) <= gi:
                    self._later.append(GemmBatch())
                self._later[gi]._wave = gi + 1
                self._later[gi].add(C_out, grp, accumulate=True, skip_flag=skip_flag)
                self._later[gi]._wave = None
            return
        if C_out.stride(1) != 1:
            raise ValueError('C must be row-major')
        p = L.GemmProblem()
        p.C, p.ldc, p.bias, p.row_scale = ptr(C_out), C_out.stride(0), ptr(bias), ptr(row_scale)
        p.M, p.N, p.accumulate = M, N, int(accumulate)
        p.seg_begin, p.seg_count = len(self.segs), len(segs)
        p.split_k = split_k
        p.skip_flag = ptr(skip_flag)
        if split_k > 1:
            part = torch.empty(split_k * M * N, dtype=torch.float32, device=C_out.device)
            self._keep.append(part)
            p.partial = ptr(part)
        for sg in segs:
            A, B = sg[0], sg[1]
            Am = sg[2] if len(sg) > 2 else None
            Bm = sg[3] if len(sg) > 3 else None
            if (M >= TC_MIN_ROWS and split_k <= 1 and Bm is None and B.stride(1) == 1 and
                    B.stride(0) != 1 and B.shape[0] * B.shape[1] <= (1 << 20)):
                # dY @ W with W given [K, N] row-major: the tensor-core kernel wants both operands
                # K-contiguous, so hand it a transposed copy of the (small) weight
                Bt = torch.empty(B.shape[1], B.shape[0], dtype=torch.float32, device=B.device)
                self._transposes.append((Bt, B))
                self._keep.append(Bt)
                B = Bt.t()
            if A.shape[0] != M or B.shape[1] != N or A.shape[1] != B.shape[0]:
                raise ValueError(f'gemm shape mismatch: A {tuple(A.shape)} B {tuple(B.shape)} '
                                 f'C {tuple(C_out.shape)}')
            if Am is not None and Am.stride() != A.stride():
                raise ValueError('A_mask must have the strides of A')
            if Bm is not None and Bm.stride() != B.stride():
                raise ValueError('B_mask must have the strides of B')
            s = L.GemmSeg(ptr(A), A.stride(0), A.stride(1), ptr(B), B.stride(0), B.stride(1),
                          ptr(Am), ptr(Bm), A.shape[1], 0)
            self.segs.append(s)
            self._keep += [A, B, Am, Bm]
        self.problems.append(p)
        self._keep += [C_out, bias, row_scale, skip_flag]

    def run(self, keep_into: Optional[list] = None):
        """``keep_into``: receives the tensors the launches use (operands, split-K partials,
        transposed weight copies) -- for launches on a stream that is joined later (ops.defer)."""
        if self._transposes:
            transpose_many(self._transposes)
            self._transposes = []
        i = 0
        n = len(self.problems)
        while i < n:
            # take problems while both tables fit
            j, nseg = i, 0
            while j < n and j - i < L.MAX_GEMM_PROBLEMS and \
                    nseg + self.problems[j].seg_count <= L.MAX_GEMM_SEGS:
                nseg += self.problems[j].seg_count
                j += 1
            if j == i:
                raise ValueError('a GEMM problem has more than 64 segments')
            parr = (L.GemmProblem * (j - i))()
            sarr = (L.GemmSeg * max(nseg, 1))()
            so = 0
            for k in range(i, j):
                src = self.problems[k]
                q = parr[k - i]
                C.memmove(C.byref(q), C.byref(src), C.sizeof(L.GemmProblem))
                for t in range(src.seg_count):
                    sarr[so + t] = self.segs[src.seg_begin + t]
                q.seg_begin = so
                so += src.seg_count
            if _DEBUG:
                for k in range(i, j):
                    src = self.problems[k]
                    ks = [self.segs[src.seg_begin + t].K for t in range(src.seg_count)]
                    print(f'[agx] gemm M={src.M} N={src.N} K={ks} split_k={src.split_k} '
                          f'acc={src.accumulate}', flush=True)
                print('[agx] gemm launch', flush=True)
            t0 = TIMER.begin() if TIMER is not None else None
            check(lib().agx_gemm_grouped(parr, j - i, sarr, max(nseg, 1), stream_ptr()),
                  'agx_gemm_grouped')
            if TIMER is not None:
                fl = nb = 0
                for k in range(i, j):
                    src = self.problems[k]
                    nb += src.M * src.N * 4
                    for t in range(src.seg_count):
                        K = self.segs[src.seg_begin + t].K
                        fl += 2 * src.M * src.N * K
                        nb += (src.M + src.N) * K * 4
                TIMER.end('gemm', nb, fl, t0)
            i = j
        if keep_into is not None:
            keep_into.extend(t for t in self._keep if t is not None)
        self.problems, self.segs, self._keep = [], [], []
        later, self._later = self._later, []
        for gb in later:
            gb.run(keep_into)


def split_k_for(k_rows: int, slab: int = 384, max_split: int = 512) -> int:
    return max(1, min(max_split, (k_rows + slab - 1) // slab))


# ------------------------------------------------------------------------------------------------
# K5/K6: fused head step on the tensor cores (agx_head_step)
# ------------------------------------------------------------------------------------------------
@dataclass
class HeadArg:
    """One Linear head with <= 64 outputs over the virtual concatenation of ``parts``."""
    parts: Sequence[torch.Tensor]            # [feat] or [feat, emb], float32 [B, width]
    weight: torch.Tensor                     # [C, K] float32 (a row slice of nn.Linear.weight)
    bias: Optional[torch.Tensor]
    d_weight: torch.Tensor                   # gradient destinations (same shapes)
    d_bias: Optional[torch.Tensor]
    loss: str = 'ce'                         # 'ce' | 'smooth_l1'
    labels: Optional[torch.Tensor] = None    # int64 [B]
    class_w: Optional[torch.Tensor] = None   # float32 [C]
    coef: float = 1.0
    target: Optional[torch.Tensor] = None    # float32 [B, C] view (row stride = target.stride(0))
    inv_count: float = 0.0
    mask: Optional[torch.Tensor] = None      # explicit multiplicative dropout mask [B, K]
    logits: Optional[torch.Tensor] = None    # optional output [B, C]


def head_step_supported(parts: Sequence[torch.Tensor], n_out: int) -> bool:
    """Shapes agx_head_step takes (see include/agx.h): first part a multiple of 64 wide, total
    width a multiple of 128, float32 rows 16-byte aligned."""
    if not parts or len(parts) > 2:
        return False
    k = sum(p.shape[1] for p in parts)
    ok = parts[0].shape[1] % 64 == 0 and k % 128 == 0 and n_out >= 1
    return ok and all(p.is_cuda and p.dtype == torch.float32 and p.stride(1) == 1 and
                      p.stride(0) % 4 == 0 and p.data_ptr() % 16 == 0 for p in parts)


class HeadStep:
    """Caller-owned buffers of one fused head step (workspace, label normaliser, loss) -- allocated
    once per batch shape so that the step is CUDA-graph capturable."""

    def __init__(self):
        self.ws = None
        self.norm = None
        self.loss = None
        self._key = None

    def run(self, heads: Sequence[HeadArg], p_drop: float = 0.0,
            seed_state: Optional[torch.Tensor] = None, group=None, accumulate: bool = False):
        n = len(heads)
        if n < 1 or n > L.MAX_HEADS:
            raise ValueError(f'1..{L.MAX_HEADS} heads per step')
        B = heads[0].parts[0].shape[0]
        dev = heads[0].weight.device
        arr = (L.Head * n)()
        keep = []
        for i, h in enumerate(heads):
            a = arr[i]
            for k, p_ in enumerate(h.parts):
                L.require_cuda(p_, 'head input')
                if p_.dtype != torch.float32 or p_.stride(1) != 1 or p_.shape[0] != B:
                    raise TypeError('head inputs must be float32 [B, width] with contiguous rows')
                a.part[k], a.ld[k], a.width[k] = ptr(p_), p_.stride(0), p_.shape[1]
            a.C = h.weight.shape[0]
            a.loss = L.HEAD_CE if h.loss == 'ce' else L.HEAD_SMOOTH_L1
            if h.weight.stride(1) != 1 or h.d_weight.stride(1) != 1:
                raise ValueError('head weight / gradient rows must be contiguous')
            a.weight, a.ldw, a.bias = ptr(h.weight), h.weight.stride(0), ptr(h.bias)
            if h.mask is not None:
                a.mask, a.ld_mask = ptr(h.mask), h.mask.stride(0)
            if h.loss == 'ce':
                lab = h.labels.to(device=dev, dtype=torch.int64).contiguous()
                keep.append(lab)
                a.labels, a.class_w, a.coef = ptr(lab), ptr(h.class_w), float(h.coef)
            else:
                if h.target.stride(1) != 1:
                    raise ValueError('projector target rows must be contiguous')
                a.target, a.ld_target, a.inv_count = ptr(h.target), h.target.stride(0), float(h.inv_count)
            if h.logits is not None:
                a.logits, a.ld_logits = ptr(h.logits), h.logits.stride(0)
            a.d_weight, a.ld_dw, a.d_bias = ptr(h.d_weight), h.d_weight.stride(0), ptr(h.d_bias)
        nbytes = lib().agx_head_step_workspace_bytes(arr, n, B)
        if nbytes == 0:
            check(-1, 'agx_head_step_workspace_bytes')
        key = (n, B, nbytes, str(dev))
        if self._key != key:
            self.ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            self.norm = torch.ones(L.MAX_HEADS, dtype=torch.float32, device=dev)
            self.loss = torch.zeros(1, dtype=torch.float32, device=dev)
            self._key = key
        check(lib().agx_head_step_prepare(arr, n, B, ptr(self.norm), ptr(self.ws), nbytes,
                                          stream_ptr()), 'agx_head_step_prepare')
        if group is not None and any(h.loss == 'ce' for h in heads):
            from .dist import small_all_reduce_    # weighted mean over the batch shards of all ranks
            small_all_reduce_(self.norm, group)
        check(lib().agx_head_step(arr, n, B, float(p_drop),
                                  ptr(seed_state) if p_drop > 0 else None, ptr(self.norm),
                                  ptr(self.loss), int(accumulate), ptr(self.ws), nbytes,
                                  stream_ptr()), 'agx_head_step')
        return self.loss


# ------------------------------------------------------------------------------------------------
# small batched ops
# ------------------------------------------------------------------------------------------------
def sum_arrays(items: Sequence[tuple]):
    """``items``: (out, [in_0..in_k][, bias]) -- out = sum of inputs (same numel, contiguous)
    (+ bias [F] broadcast over the rows of out [*, F])."""
    for base in range(0, len(items), L.MAX_TENSORS):
        part = items[base:base + L.MAX_TENSORS]
        arr = (L.SumDesc * len(part))()
        for i, (out, ins, *rest) in enumerate(part):
            d = arr[i]
            d.out, d.n_in, d.numel = ptr(out), len(ins), out.numel()
            if rest and rest[0] is not None:
                b = rest[0]
                if b.dtype != torch.float32 or not b.is_contiguous() or out.dim() != 2 or \
                        b.numel() != out.shape[1] or not out.is_contiguous():
                    raise ValueError('sum_arrays bias must be contiguous float32 [out.shape[1]]')
                d.bias, d.bias_F = ptr(b), b.numel()
            for k, t in enumerate(ins):
                if not t.is_contiguous() or t.numel() != out.numel():
                    raise ValueError('sum_arrays needs contiguous equally sized tensors')
                d.inp[k] = ptr(t)
        check(lib().agx_sum_arrays(arr, len(part), stream_ptr()), 'agx_sum_arrays')


def colsum(items: Sequence[Tuple[torch.Tensor, torch.Tensor, bool]]):
    """``items``: (x [rows, F], out [F], accumulate)."""
    if not items:
        return
    dev = items[0][0].device
    for base in range(0, len(items), L.MAX_TENSORS):
        part = items[base:base + L.MAX_TENSORS]
        arr = (L.ColsumDesc * len(part))()
        rows = sum(x.shape[0] for x, _, _ in part)
        maxf = max(x.shape[1] for x, _, _ in part)
        for i, (x, out, acc) in enumerate(part):
            arr[i] = L.ColsumDesc(ptr(x), x.stride(0), ptr(out), x.shape[0], x.shape[1], int(acc), 0)
        n_ws = lib().agx_colsum_workspace_floats(rows, len(part), maxf)
        ws = torch.empty(n_ws, dtype=torch.float32, device=dev)
        check(lib().agx_colsum(arr, len(part), ptr(ws), n_ws, stream_ptr()), 'agx_colsum')


def dropout_mask(shape, p: float, seed_state: torch.Tensor) -> torch.Tensor:
    """Multiplicative mask ``bernoulli(1-p) / (1-p)`` from Philox; ``seed_state`` = device
    uint64-as-int64 [2] (key, counter)."""
    mask = torch.empty(shape, dtype=torch.float32, device=seed_state.device)
    check(lib().agx_dropout_mask(ptr(mask), mask.numel(), float(p), ptr(seed_state), stream_ptr()),
          'agx_dropout_mask')
    return mask


def fill_(t: torch.Tensor, v: float):
    check(lib().agx_fill_f32(ptr(t), t.numel(), float(v), stream_ptr()), 'agx_fill_f32')
    return t


def zeros(shape, device) -> torch.Tensor:
    return fill_(torch.empty(shape, dtype=torch.float32, device=device), 0.0)


def scale_mask(x: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    y = torch.empty_like(x)
    check(lib().agx_scale_mask(ptr(x), ptr(mask), ptr(y), x.numel(), stream_ptr()),
          'agx_scale_mask')
    return y


def gather_rows(table: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    out = torch.empty(idx.numel(), table.shape[1], dtype=torch.float32, device=table.device)
    check(lib().agx_gather_rows(ptr(table), table.stride(0), ptr(idx), idx.numel(), table.shape[1],
                                ptr(out), out.stride(0), stream_ptr()), 'agx_gather_rows')
    return out


def is_identity(x: torch.Tensor) -> torch.Tensor:
    """Device int32 flag: 1 iff ``x`` is a square identity (one-hot node features)."""
    flag = torch.zeros(2, dtype=torch.int32, device=x.device)
    if x.dim() != 2 or x.shape[0] != x.shape[1]:
        return flag[:1]
    check(lib().agx_is_identity(ptr(x), x.stride(0), x.shape[0], ptr(flag), flag.data_ptr() + 4,
                                stream_ptr()), 'agx_is_identity')
    return flag[:1]


def transpose_many(pairs: Sequence[Tuple[torch.Tensor, torch.Tensor]]):
    """``pairs``: (out [cols, rows], inp [rows, cols]) or (out, inp, accumulate) -- all transposes
    in one launch per 48; ``accumulate``: out += inp^T."""
    for base in range(0, len(pairs), L.MAX_TENSORS):
        part = pairs[base:base + L.MAX_TENSORS]
        arr = (L.TransposeDesc * len(part))()
        for i, item in enumerate(part):
            out, inp = item[0], item[1]
            acc = bool(item[2]) if len(item) > 2 else False
            if inp.stride(1) != 1 or out.stride(1) != 1 or out.shape != (inp.shape[1], inp.shape[0]):
                raise ValueError('transpose_many: row-major [r, c] -> [c, r] expected')
            arr[i] = L.TransposeDesc(ptr(inp), inp.stride(0), ptr(out), out.stride(0), inp.shape[0],
                                     inp.shape[1], int(acc), 0)
        check(lib().agx_transpose_batched(arr, len(part), stream_ptr()), 'agx_transpose_batched')


def transpose_into(out: torch.Tensor, inp: torch.Tensor, only_if_flag: Optional[torch.Tensor] = None):
    check(lib().agx_transpose(ptr(inp), inp.stride(0), inp.shape[0], inp.shape[1], ptr(out),
                              out.stride(0), ptr(only_if_flag), stream_ptr()), 'agx_transpose')


def adam_step(param, grad, exp_avg, exp_avg_sq, step, lr, betas=(0.9, 0.999), eps=1e-8,
              weight_decay=0.0):
    check(lib().agx_adam_step(ptr(param), ptr(grad), ptr(exp_avg), ptr(exp_avg_sq), param.numel(),
                              lr, betas[0], betas[1], eps, weight_decay, ptr(step), stream_ptr()),
          'agx_adam_step')


def pack_rows(x: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    out = torch.empty(idx.numel(), x.shape[1], dtype=torch.float32, device=x.device)
    check(lib().agx_pack_rows(ptr(x), x.stride(0), ptr(idx), idx.numel(), x.shape[1], ptr(out),
                              stream_ptr()), 'agx_pack_rows')
    return out


def unpack_rows_add_(x: torch.Tensor, idx: torch.Tensor, src: torch.Tensor):
    check(lib().agx_unpack_rows_add(ptr(x), x.stride(0), ptr(idx), idx.numel(), x.shape[1],
                                    ptr(src), stream_ptr()), 'agx_unpack_rows_add')
