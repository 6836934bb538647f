"""``FlatAdam``: ``torch.optim.Adam`` arithmetic (the reference's optimizer,
/root/reference/src/train_gnn_embeddings.py:144, src/train_projector.py:34,
src/train_new_multimodal_multitask.py:56) as ONE agx launch over a flat parameter arena.

On first use every initialised parameter is re-pointed at a slice of one contiguous float32
buffer (and its ``.grad`` at a slice of one gradient buffer), so a step is a single elementwise
kernel instead of ~250 per-tensor updates, and ``zero_grad`` is one fill.  The step counter lives
on the device so a captured CUDA graph advances the bias corrections on replay.

Parameters that never receive a gradient (dead ``conv_out`` relations, unused lazy ``lins``)
behave as under torch's Adam, which skips ``grad is None``: with zero first/second moments the
update is exactly zero.
"""
from __future__ import annotations

from typing import Iterable, List

import torch
from torch import nn

from . import ops
from ._lib import check, lib, ptr, stream_ptr


class FlatAdam:
    def __init__(self, params: Iterable[nn.Parameter], lr: float = 1e-3, betas=(0.9, 0.999),
                 eps: float = 1e-8, weight_decay: float = 0.0):
        self._params: List[nn.Parameter] = list(params)
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), betas, eps, weight_decay
        self.flat = self.grad = self.exp_avg = self.exp_avg_sq = self.step_t = None
        self._members: List[nn.Parameter] = []

    # ------------------------------------------------------------------------------------------
    def _flatten(self):
        members = [p for p in self._params
                   if not isinstance(p, nn.parameter.UninitializedParameter) and p.requires_grad]
        if not members:
            raise ValueError('FlatAdam: no initialised parameters')
        dev = members[0].device
        if dev.type != 'cuda':
            raise RuntimeError('FlatAdam needs CUDA parameters: this package has no CPU path')
        for p in members:
            if p.dtype != torch.float32 or p.device != dev:
                raise TypeError('FlatAdam: parameters must be float32 on one device')
        sizes = [((p.numel() + 3) // 4) * 4 for p in members]       # keep 16-byte alignment
        total = sum(sizes)
        self.flat = ops.zeros(total, dev)
        self.grad = ops.zeros(total, dev)
        self.exp_avg = ops.zeros(total, dev)
        self.exp_avg_sq = ops.zeros(total, dev)
        self.step_t = torch.zeros(1, dtype=torch.int32, device=dev)
        off = 0
        with torch.no_grad():
            for p, sz in zip(members, sizes):
                n = p.numel()
                view = self.flat[off:off + n].view(p.shape)
                view.copy_(p.data)
                g_old = p.grad
                p.data = view
                p.grad = self.grad[off:off + n].view(p.shape)
                if g_old is not None:
                    p.grad.copy_(g_old)
                off += sz
        self._members = members
        self._member_ids = {id(p) for p in members}

    def flatten(self):
        """Move the parameters into the flat arena now (otherwise done by the first step)."""
        if self.flat is None:
            self._flatten()
        return self

    def _ensure(self):
        if self.flat is None:
            self._flatten()
        else:
            # a parameter materialised after flattening would silently miss updates
            for p in self._params:
                if p.requires_grad and not isinstance(p, nn.parameter.UninitializedParameter) \
                        and id(p) not in self._member_ids:
                    raise RuntimeError('FlatAdam: a parameter was initialised after the first '
                                       'step; run one forward before the first step')
            for p in self._members:     # zero_grad(set_to_none=True) elsewhere detaches the views
                if p.grad is None:
                    raise RuntimeError('FlatAdam: parameter .grad was reset to None; use '
                                       'FlatAdam.zero_grad()')

    # ------------------------------------------------------------------------------------------
    def zero_grad(self, set_to_none: bool = False):
        if self.flat is None:
            for p in self._params:
                if not isinstance(p, nn.parameter.UninitializedParameter):
                    p.grad = None
            return
        ops.fill_(self.grad, 0.0)

    @torch.no_grad()
    def step(self):
        self._ensure()
        self.step_t += 1
        ops.adam_step(self.flat, self.grad, self.exp_avg, self.exp_avg_sq, self.step_t, self.lr,
                      self.betas, self.eps, self.weight_decay)

    def state_dict(self):
        return {'step': self.step_t, 'exp_avg': self.exp_avg, 'exp_avg_sq': self.exp_avg_sq,
                'lr': self.lr, 'betas': self.betas, 'eps': self.eps}


class FlatSGD(FlatAdam):
    """``torch.optim.SGD(lr, momentum)`` (the ContextNet optimizer,
    /root/reference/src/train_baseline_context.py:49) over the same flat arena as ``FlatAdam``."""

    def __init__(self, params, lr: float = 1e-3, momentum: float = 0.0, weight_decay: float = 0.0):
        super().__init__(params, lr=lr, weight_decay=weight_decay)
        self.momentum = float(momentum)

    @torch.no_grad()
    def step(self):
        self._ensure()
        self.step_t += 1
        check(lib().agx_sgd_step(ptr(self.flat), ptr(self.grad), ptr(self.exp_avg), self.flat.numel(),
                                 self.lr, self.momentum, self.weight_decay, stream_ptr()),
              'agx_sgd_step')

    def state_dict(self):
        return {'step': self.step_t, 'momentum_buffer': self.exp_avg, 'lr': self.lr,
                'momentum': self.momentum}
