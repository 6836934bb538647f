"""TEST-ONLY restatement of the tensor-level wrappers in ``mmac_b200.ops`` with plain torch on CPU.

The product has no CPU execution path (``ops`` calls ``libagx.so`` and raises for CPU tensors).  To
check the HOST-SIDE logic of the autograd functions without a GPU -- which kernels are asked to
combine which buffers with which index arrays, the gradient formulas, the wave ordering of the
grouped GEMMs -- the CPU suite swaps the ``ops`` entry points for the functions below, each written
from the C-ABI contract in ``include/agx.h`` (not from the CUDA source), and compares the result
with the oracle.  Nothing under ``multi-modal-art-classifier_b200/`` imports this file.
"""
from __future__ import annotations

import contextlib

import torch

from mmac_b200 import _lib as L
from mmac_b200 import ops


def _csr_build(edge_lists, check_range=True, buffers=None):
    out = []
    for keys, vals, n_rows, n_cols in edge_lists:
        if check_range and keys.numel() and (int(keys.min()) < 0 or int(keys.max()) >= n_rows or
                                             int(vals.min()) < 0 or int(vals.max()) >= n_cols):
            raise IndexError('edge_index contains node ids outside [0, num_nodes)')
        order = torch.sort(keys, stable=True).indices
        deg = torch.bincount(keys, minlength=n_rows)
        rowptr = torch.zeros(n_rows + 1, dtype=torch.int32)
        rowptr[1:] = torch.cumsum(deg, 0)
        out.append(ops.CSR(rowptr, vals[order].to(torch.int32), order.to(torch.int32),
                           deg.clamp(min=1).float(), int(n_rows), int(n_cols), int(keys.numel())))
    if buffers is not None and not buffers:
        buffers.append([None, None, None, None, torch.zeros(1, dtype=torch.int32), None])
    return out


def _rel_sum(a: ops.RelArg, n_rows: int, F: int) -> torch.Tensor:
    """agx_rel_t contract: out[i] = sum over slots e of row i of x[col[e]] * scale(e)."""
    c = a.csr
    E = c.n_edges
    deg = (c.rowptr[1:] - c.rowptr[:-1]).long()
    rows = torch.repeat_interleave(torch.arange(c.n_rows), deg)
    col = c.col[:E].long()
    w = torch.ones(E, dtype=torch.float64)
    if a.nbr_scale is not None:
        w = w / a.nbr_scale.double()[col]
    if a.edge_w is not None:
        idx = a.edge_w_idx[:E].long() if a.edge_w_idx is not None else torch.arange(E)
        w = w * a.edge_w.double()[idx]
    out = torch.zeros(n_rows, F, dtype=torch.float64)
    out.index_add_(0, rows, a.x.double()[col][:, :F] * w[:, None])
    if a.mean_rows:
        out = out / c.cnt.double()[:, None]
    return out


def _aggregate_rows(groups, F):
    for out, rels, acc, *rest in groups:
        tot = sum(_rel_sum(a, out.shape[0], F) for a in rels)
        if acc:
            tot = tot + out.double()
        if rest and rest[0] is not None:
            tot = tot + rest[0].double()[None, :]
        out.copy_(tot.to(out.dtype))


def _aggregate_chunks(segs, F):
    for out, a in segs:
        out.copy_(_rel_sum(a, out.shape[0], F).to(out.dtype))


class _GemmBatch:
    """C (+)= sum_s opA_s @ opB_s (+ bias), masks multiply the operands element-wise."""

    def __init__(self):
        self.problems = []

    def add(self, C_out, segs, bias=None, accumulate=False, row_scale=None, split_k=1,
            skip_flag=None):
        assert row_scale is None and skip_flag is None
        M, N = C_out.shape
        for sg in segs:
            A, B = sg[0], sg[1]
            assert A.shape[0] == M and B.shape[1] == N and A.shape[1] == B.shape[0], \
                (A.shape, B.shape, C_out.shape)
        if split_k > 1:
            assert len(segs) == 1
        self.problems.append((C_out, list(segs), bias, accumulate))

    def run(self):
        for C_out, segs, bias, acc in self.problems:
            tot = torch.zeros(C_out.shape, dtype=torch.float64)
            for sg in segs:
                A, B = sg[0].double(), sg[1].double()
                if len(sg) > 2 and sg[2] is not None:
                    A = A * sg[2].double()
                if len(sg) > 3 and sg[3] is not None:
                    B = B * sg[3].double()
                tot += A @ B
            if bias is not None:
                tot += bias.double()[None, :]
            if acc:
                tot += C_out.double()
            C_out.copy_(tot.to(C_out.dtype))
        self.problems = []


def _rows_of(csr):
    deg = (csr.rowptr[1:] - csr.rowptr[:-1]).long()
    return torch.repeat_interleave(torch.arange(csr.n_rows), deg)


def _gat_edge_softmax(rels, slope, backward=False):
    for a in rels:
        c = a.csr
        E = c.n_edges
        if c.n_rows == 0:
            continue
        rows, col = _rows_of(c), c.col[:E].long()
        raw = a.a_l.double()[col] + a.a_r.double()[rows]
        if not backward:
            e = torch.where(raw > 0, raw, slope * raw)
            m = torch.full((c.n_rows,), float('-inf'), dtype=torch.float64).scatter_reduce(
                0, rows, e, reduce='amax', include_self=True)
            ex = torch.exp(e - m[rows])
            den = torch.zeros(c.n_rows, dtype=torch.float64).index_add_(0, rows, ex)
            a.alpha[:E] = (ex / (den[rows] + 1e-16)).float()
        else:
            al, dal = a.alpha[:E].double(), a.dalpha[:E].double()
            s = torch.zeros(c.n_rows, dtype=torch.float64).index_add_(0, rows, al * dal)
            de = al * (dal - s[rows]) * torch.where(raw > 0, 1.0, slope)
            a.de[:E] = de.float()
            a.da_r.copy_(torch.zeros(c.n_rows, dtype=torch.float64).index_add_(0, rows, de).float())


def _sddmm(segs, F):
    for row, col, a, b, out in segs:
        E = row.numel()
        out[:E] = (a.double()[row.long()] * b.double()[col.long()]).sum(1).float()


def _sum_arrays(items):
    for out, ins, *rest in items:
        tot = sum(t.double().reshape(out.shape) for t in ins)
        if rest and rest[0] is not None:
            tot = tot + rest[0].double()[None, :]
        out.copy_(tot.to(out.dtype))


def _colsum(items):
    for x, out, acc in items:
        tot = x.double().sum(0)
        out.copy_((tot + out.double() if acc else tot).float())


def _transpose_many(pairs):
    for out, inp in pairs:
        out.copy_(inp.t())


def _fill_(t, v):
    return t.fill_(v)


def _is_identity(x):
    n = x.shape[0]
    return torch.tensor(int(x.shape[0] == x.shape[1] and bool(torch.equal(x, torch.eye(n)))))


_PATCHES = {
    'csr_build': _csr_build, 'aggregate_rows': _aggregate_rows,
    'aggregate_chunks': _aggregate_chunks, 'GemmBatch': _GemmBatch,
    'gat_edge_softmax': _gat_edge_softmax, 'sddmm': _sddmm, 'sum_arrays': _sum_arrays,
    'colsum': _colsum, 'transpose_many': _transpose_many, 'is_identity': _is_identity,
    'fill_': _fill_,
}


@contextlib.contextmanager
def cpu_ops():
    """Inside the block ``mmac_b200.ops`` computes with torch on CPU (see the module docstring)."""
    saved = {k: getattr(ops, k) for k in _PATCHES}
    saved_req = L.require_cuda
    try:
        for k, v in _PATCHES.items():
            setattr(ops, k, v)
        L.require_cuda = lambda t, name: None
        yield
    finally:
        for k, v in saved.items():
            setattr(ops, k, v)
        L.require_cuda = saved_req
