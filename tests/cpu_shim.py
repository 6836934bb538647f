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

    def run(self, keep_into=None):
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
    for item in pairs:
        out, inp = item[0], item[1]
        if len(item) > 2 and item[2]:
            out.add_(inp.t())
        else:
            out.copy_(inp.t())


def _fill_(t, v):
    return t.fill_(v)


def _is_identity(x):
    n = x.shape[0]
    return torch.tensor(int(x.shape[0] == x.shape[1] and bool(torch.equal(x, torch.eye(n)))))


def _dropout_mask(shape, p, seed_state):
    """Stand-in for the Philox kernel: a mask that is a deterministic function of the (key,
    counter) state, so that equal states give equal masks (not the kernel's bit stream)."""
    gen = torch.Generator().manual_seed(int((int(seed_state[0]) * 31 + int(seed_state[1])) % (2 ** 62)))
    return (torch.rand(shape, generator=gen) >= p).float() / (1.0 - p)


def _scale_mask(x, mask):
    return x * mask


def _zeros(shape, device):
    return torch.zeros(shape, dtype=torch.float32, device=device)


# ------------------------------------------------------------------------------------------------
# pointer-level restatement of the entry points functional.py calls on ``lib()`` directly
# (BatchNorm, log_softmax / nll): host addresses of CPU tensors viewed through numpy
# ------------------------------------------------------------------------------------------------
def _view(addr, rows, cols, ld=None, dtype=torch.float32):
    """torch view of ``rows x cols`` elements at host address ``addr`` (row stride ``ld``)."""
    import ctypes
    import numpy as np
    if not addr:
        return None
    ld = cols if ld is None else int(ld)
    ct = {torch.float32: ctypes.c_float, torch.int64: ctypes.c_int64, torch.int32: ctypes.c_int32,
          torch.float64: ctypes.c_double}[dtype]
    n = (rows - 1) * ld + cols if rows > 0 else 0
    if n == 0:
        return torch.empty(rows, cols, dtype=dtype)
    flat = torch.from_numpy(np.ctypeslib.as_array((ct * n).from_address(int(addr))))
    return torch.as_strided(flat, (rows, cols), (ld, 1))


def _bn_outputs(d, y64, N, F):
    """y (optional) and relu(y) * dropout of one agx_bn_desc_t: mask tensor, or the (key, counter,
    offset) form -- here the stand-in stream of _dropout_mask at counter + offset / 4."""
    y32 = y64.float()
    if d.y:
        _view(d.y, N, F).copy_(y32)
    if d.y_act:
        a = torch.relu(y32.double())
        if d.dmask:
            a = a * _view(d.dmask, N, F).double()
        elif d.drop_seed and d.drop_p > 0:
            st = _view(d.drop_seed, 1, 2, dtype=torch.int64).view(-1)
            m = _dropout_mask((N, F), float(d.drop_p), [int(st[0]), int(st[1]) + int(d.drop_offset) // 4])
            a = a * m.double()
        _view(d.y_act, N, F).copy_(a.float())


class _FakeLib:
    """The agx_* functions with device pointers replaced by host addresses (include/agx.h
    semantics, float64 arithmetic, results rounded to float32 once)."""

    def agx_last_error(self):
        return b'cpu shim'

    def agx_bn_workspace_floats(self, rows, n, F):
        return 16

    def agx_bn_forward(self, arr, n, F, training, momentum, eps, ws, n_ws, stream):
        for i in range(n):
            d = arr[i]
            N = d.n_rows
            x = _view(d.x, N, F).double()
            w, b = _view(d.weight, 1, F).double(), _view(d.bias, 1, F).double()
            rm, rv = _view(d.running_mean, 1, F), _view(d.running_var, 1, F)
            if training:
                mean = x.mean(0, keepdim=True)
                var = x.var(0, unbiased=False, keepdim=True)
                if rm is not None:
                    rm.copy_(((1 - momentum) * rm.double() + momentum * mean).float())
                    rv.copy_(((1 - momentum) * rv.double() + momentum * var * N / (N - 1)).float())
            else:
                mean, var = rm.double(), rv.double()
            invstd = 1.0 / torch.sqrt(var + eps)
            y = (x - mean) * invstd * w + b
            _view(d.save_mean, 1, F).copy_(mean.float())
            _view(d.save_invstd, 1, F).copy_(invstd.float())
            _bn_outputs(d, y, N, F)
        return 0

    def agx_bn_backward(self, arr, n, F, training, ws, n_ws, stream):
        for i in range(n):
            d = arr[i]
            N = d.n_rows
            x = _view(d.x, N, F).double()
            g = torch.zeros(N, F, dtype=torch.float64)
            if d.dy:
                g = g + _view(d.dy, N, F).double()
            if d.dy_act:
                ga = _view(d.dy_act, N, F).double() * (_view(d.y, N, F) > 0).double()
                if d.dmask:
                    ga = ga * _view(d.dmask, N, F).double()
                else:
                    ga = ga * (float(d.act_scale) if d.act_scale else 1.0)
                g = g + ga
            w = _view(d.weight, 1, F).double()
            mean, invstd = _view(d.save_mean, 1, F).double(), _view(d.save_invstd, 1, F).double()
            xhat = (x - mean) * invstd
            if d.dweight:
                dw = _view(d.dweight, 1, F)
                dw.copy_((dw.double() + (g * xhat).sum(0, keepdim=True)).float())
            if d.dbias:
                db = _view(d.dbias, 1, F)
                db.copy_((db.double() + g.sum(0, keepdim=True)).float())
            if d.dx:
                if training:
                    dx = w * invstd * (g - g.mean(0, keepdim=True) -
                                       xhat * (g * xhat).mean(0, keepdim=True))
                else:
                    dx = g * w * invstd
                _view(d.dx, N, F).copy_(dx.float())
        return 0

    # -- multi-GPU BatchNorm: statistics over the rows of all ranks (phases, see agx.h) ----------
    def agx_bn_forward_phase(self, arr, n, F, training, momentum, eps, ws, n_ws, phases, sums,
                             counts, stream):
        S = _view(sums, n, 2 * F, dtype=torch.float64)
        cnt = _view(counts, 1, n, dtype=torch.float64).view(-1)
        for i in range(n):
            d = arr[i]
            N = d.n_rows
            x = _view(d.x, N, F).double()
            if phases & 1:
                S[i, :F] = x.sum(0)
                S[i, F:] = (x * x).sum(0)
            if phases & 4:
                c = float(cnt[i])
                mean = (S[i, :F] / c).view(1, F)
                var = (S[i, F:] / c).view(1, F) - mean * mean
                rm, rv = _view(d.running_mean, 1, F), _view(d.running_var, 1, F)
                if rm is not None:
                    rm.copy_(((1 - momentum) * rm.double() + momentum * mean).float())
                    rv.copy_(((1 - momentum) * rv.double() + momentum * var * c / (c - 1)).float())
                invstd = 1.0 / torch.sqrt(var + eps)
                w, b = _view(d.weight, 1, F).double(), _view(d.bias, 1, F).double()
                _view(d.save_mean, 1, F).copy_(mean.float())
                _view(d.save_invstd, 1, F).copy_(invstd.float())
                _bn_outputs(d, (x - mean) * invstd * w + b, N, F)
        return 0

    def agx_bn_backward_phase(self, arr, n, F, training, ws, n_ws, phases, totals, counts, stream):
        T = _view(totals, n, 2 * F, dtype=torch.float64)
        cnt = _view(counts, 1, n, dtype=torch.float64).view(-1)
        for i in range(n):
            d = arr[i]
            N = d.n_rows
            x = _view(d.x, N, F).double()
            g = torch.zeros(N, F, dtype=torch.float64)
            if d.dy:
                g = g + _view(d.dy, N, F).double()
            if d.dy_act:
                ga = _view(d.dy_act, N, F).double() * (_view(d.y, N, F) > 0).double()
                if d.dmask:
                    ga = ga * _view(d.dmask, N, F).double()
                else:
                    ga = ga * (float(d.act_scale) if d.act_scale else 1.0)
                g = g + ga
            mean, invstd = _view(d.save_mean, 1, F).double(), _view(d.save_invstd, 1, F).double()
            xhat = (x - mean) * invstd
            if phases & 1:
                T[i, :F] = g.sum(0)
                T[i, F:] = (g * xhat).sum(0)
                if d.dweight:
                    dw = _view(d.dweight, 1, F)
                    dw.copy_((dw.double() + (g * xhat).sum(0, keepdim=True)).float())
                if d.dbias:
                    db = _view(d.dbias, 1, F)
                    db.copy_((db.double() + g.sum(0, keepdim=True)).float())
            if phases & 2 and d.dx:
                c = float(cnt[i])
                w = _view(d.weight, 1, F).double()
                dx = w * invstd * (g - T[i, :F].view(1, F) / c - xhat * T[i, F:].view(1, F) / c)
                _view(d.dx, N, F).copy_(dx.float())
        return 0

    # -- boundary rows of a destination partition ----------------------------------------------
    def agx_pack_rows(self, x, ld, idx, n_idx, F, out, stream):
        if n_idx:
            rows = _view(idx, 1, n_idx, dtype=torch.int32).view(-1).long()
            xv = _view(x, int(rows.max()) + 1, F, ld)
            _view(out, n_idx, F).copy_(xv[rows])
        return 0

    def agx_unpack_rows_add(self, x, ld, idx, n_idx, F, src, stream):
        if n_idx:
            rows = _view(idx, 1, n_idx, dtype=torch.int32).view(-1).long()
            xv = _view(x, int(rows.max()) + 1, F, ld)
            xv.index_add_(0, rows, _view(src, n_idx, F))
        return 0

    @staticmethod
    def _weights(labels, class_w, n, C):
        y = _view(labels, 1, n, dtype=torch.int64).view(-1)
        ok = (y >= 0) & (y < C)
        wy = ok.double()
        if class_w:
            wy = wy * _view(class_w, 1, C).double().view(-1)[y.clamp(0, C - 1)]
        return y.clamp(0, C - 1), wy

    def agx_log_softmax_nll(self, logits, ld, n, C, labels, class_w, logp, ldp, loss_sum, row_ws,
                            stream):
        lp = torch.log_softmax(_view(logits, n, C, ld).double(), dim=1)
        if logp:
            _view(logp, n, C, ldp).copy_(lp.float())
        if labels:
            y, wy = self._weights(labels, class_w, n, C)
            nll = -lp[torch.arange(n), y]
            _view(loss_sum, 1, 2).copy_(torch.stack([(wy * nll).sum(), wy.sum()]).float().view(1, 2))
        return 0

    def agx_nll_forward(self, logp, ld, n, C, labels, class_w, loss_sum, row_ws, stream):
        lp = _view(logp, n, C, ld).double()
        y, wy = self._weights(labels, class_w, n, C)
        nll = -lp[torch.arange(n), y]
        _view(loss_sum, 1, 2).copy_(torch.stack([(wy * nll).sum(), wy.sum()]).float().view(1, 2))
        return 0

    def agx_nll_backward(self, n, C, labels, class_w, loss_sum, gscale, coef, dlogp, ld, stream):
        y, wy = self._weights(labels, class_w, n, C)
        ls = _view(loss_sum, 1, 2).double().view(-1)
        gs = _view(gscale, 1, 1).double().view(-1)[0]
        out = torch.zeros(n, C, dtype=torch.float64)
        out[torch.arange(n), y] = -coef * gs * wy / ls[1]
        _view(dlogp, n, C, ld).copy_(out.float())
        return 0

    def agx_loss_finish(self, loss_sum, coef, loss, accumulate, stream):
        ls = _view(loss_sum, 1, 2).double().view(-1)
        out = _view(loss, 1, 1)
        v = coef * ls[0] / ls[1]
        out.copy_((out.double() + v if accumulate else v).float().view(1, 1))
        return 0

    def agx_log_softmax_nll_bwd(self, logp, ldp, n, C, labels, class_w, loss_sum, gscale, coef,
                                dlogp, lddp, dlogits, ld, stream):
        lp = _view(logp, n, C, ldp).double()
        sm = torch.exp(lp)
        out = torch.zeros(n, C, dtype=torch.float64)
        if labels:
            y, wy = self._weights(labels, class_w, n, C)
            ls = _view(loss_sum, 1, 2).double().view(-1)
            gs = _view(gscale, 1, 1).double().view(-1)[0]
            onehot = torch.zeros(n, C, dtype=torch.float64)
            onehot[torch.arange(n), y] = 1.0
            out = coef * gs * wy[:, None] * (sm - onehot) / ls[1]
        if dlogp:
            g = _view(dlogp, n, C, lddp).double()
            out = out + g - sm * g.sum(1, keepdim=True)
        _view(dlogits, n, C, ld).copy_(out.float())
        return 0


_PATCHES = {
    'csr_build': _csr_build, 'aggregate_rows': _aggregate_rows,
    'aggregate_chunks': _aggregate_chunks, 'GemmBatch': _GemmBatch,
    'gat_edge_softmax': _gat_edge_softmax, 'sddmm': _sddmm, 'sum_arrays': _sum_arrays,
    'colsum': _colsum, 'transpose_many': _transpose_many, 'is_identity': _is_identity,
    'fill_': _fill_, 'scale_mask': _scale_mask, 'zeros': _zeros, 'dropout_mask': _dropout_mask,
}


def _coalesce_undirected(row, col, n_nodes):
    """agx_coalesce_undirected contract: unique (row, col) pairs of the symmetrised list, ascending
    by row * n_nodes + col."""
    key = torch.cat([row * n_nodes + col, col * n_nodes + row])
    key = torch.unique(key, sorted=True)
    return torch.stack([key // n_nodes, key % n_nodes], dim=0)


_PATCHES['coalesce_undirected'] = _coalesce_undirected


@contextlib.contextmanager
def cpu_ops():
    """Inside the block ``mmac_b200.ops`` computes with torch on CPU (see the module docstring)."""
    from mmac_b200 import dist as AD
    from mmac_b200 import functional as AF
    saved = {k: getattr(ops, k) for k in _PATCHES}
    saved_req = L.require_cuda
    saved_dev = L.compute_device
    saved_mod = [(m, m.lib, m.stream_ptr) for m in (AF, AD)]
    fake = _FakeLib()
    try:
        for k, v in _PATCHES.items():
            setattr(ops, k, v)
        L.require_cuda = lambda t, name: None
        L.compute_device = lambda: torch.device('cpu')
        for m, _, _ in saved_mod:
            m.lib = lambda: fake
            m.stream_ptr = lambda: 0
        yield
    finally:
        for k, v in saved.items():
            setattr(ops, k, v)
        L.require_cuda = saved_req
        L.compute_device = saved_dev
        for m, lib_, sp_ in saved_mod:
            m.lib, m.stream_ptr = lib_, sp_
