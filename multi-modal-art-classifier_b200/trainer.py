"""Device-side training loops with the semantics of the reference's scripts:

``GNNTrainer``   ``hetero_training()`` / ``hetero_test()`` / ``save_embeddings()`` of
                 /root/reference/src/train_gnn_embeddings.py:39-66,82-93,144-156 (full-graph
                 forward, ``nll_loss`` over all artwork nodes, ``backward``, Adam lr=0.01).
``HeadTrainer``  the mini-batch loops of src/train_new_multimodal_multitask.py:62-90 and
                 src/train_projector.py:39-59 (Adam lr=3e-4) on precomputed backbone features.

The whole step (zero_grad -> forward -> loss -> backward -> Adam) is captured once into a CUDA
graph and replayed: ~150 agx launches per GNN step otherwise cost more host time than device time.
"""
from __future__ import annotations

import copy
from collections import OrderedDict
from typing import Dict, Optional

import torch

from . import functional as AF
from . import ops
from .graph import get_plan
from .heads import multitask_loss, projector_loss
from .optim import FlatAdam


class _StateSnapshot:
    """Everything a training step mutates -- parameters and Adam state (flat arenas), module
    buffers (BatchNorm running statistics), dropout Philox counters.  CUDA-graph capture needs
    warm-up steps on the capture stream; they are real steps, so the state is put back afterwards
    and N calls of ``train_step`` / ``step`` stay N optimizer steps."""

    def __init__(self, model: torch.nn.Module, opt: FlatAdam):
        self.model, self.opt = model, opt
        self.opt_state = [t.clone() for t in (opt.flat, opt.exp_avg, opt.exp_avg_sq, opt.step_t)]
        self.buffers = {k: v.clone() for k, v in model.named_buffers()}
        # Philox streams: the per-rank one and the one shared by the replicated node types
        self.seeds = {}
        for name, m in model.named_modules():
            for attr in ('_seed', '_seed_common'):
                if hasattr(m, attr):
                    v = getattr(m, attr)
                    self.seeds[(name, attr)] = None if v is None else v.clone()

    @torch.no_grad()
    def restore(self):
        for t, saved in zip((self.opt.flat, self.opt.exp_avg, self.opt.exp_avg_sq, self.opt.step_t),
                            self.opt_state):
            t.copy_(saved)
        for k, v in self.model.named_buffers():
            if k in self.buffers:
                v.copy_(self.buffers[k])
        for name, m in self.model.named_modules():
            for attr in ('_seed', '_seed_common'):
                cur = getattr(m, attr, None)
                if (name, attr) in self.seeds and cur is not None:
                    if self.seeds[(name, attr)] is None:
                        cur[1] = 0                      # the stream had not been started
                    else:
                        cur.copy_(self.seeds[(name, attr)])


def _accuracy(logp: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """``predicted.argmax(dim=1).eq(labels).sum() / N`` (train_gnn_embeddings.py:20-21)."""
    return logp.argmax(dim=1).eq(labels).sum() / logp.shape[0]


class GNNTrainer:
    """``dist_ctx`` (dist.DistContext) makes this one rank of a multi-GPU job: ``x_dict`` / ``labels``
    are the rows this rank owns, ``edge_index_dict`` its edges (dist.GraphPartition.edge_index, or
    its own block of the replicated graph); the loss, BatchNorm statistics and weight gradients are
    those of the whole graph (NCCL all-reduces inside the captured step)."""

    def __init__(self, model: torch.nn.Module, x_dict, edge_index_dict, labels: torch.Tensor,
                 lr: float = 0.01, use_cuda_graph: bool = True, node_type: str = 'artwork',
                 dist_ctx=None):
        self.model = model
        self.ctx = dist_ctx
        self.group = dist_ctx.group if dist_ctx is not None else None
        dev0 = next(v for v in x_dict.values() if torch.is_tensor(v)).device
        if self.group is not None and dev0.type == 'cuda':
            from .dist import enable_peer_allreduce          # small reductions: one NVLink kernel
            enable_peer_allreduce(self.group, dev0)
        if dist_ctx is not None:
            from .hetero import HeteroModule
            for m in model.modules():
                if isinstance(m, HeteroModule):
                    m.set_distributed(dist_ctx)
        self._id_flags = {}
        self._identity_seen = {}
        self._plan = None
        from .data import Identity, identity_tensor
        ref = next(v for v in x_dict.values() if torch.is_tensor(v))
        # data.Identity(n) markers (one-hot features): device-side eye tensors, never uploaded
        self.x = OrderedDict((k, identity_tensor(v.n, ref.device) if isinstance(v, Identity) else v)
                             for k, v in x_dict.items())
        self.ei = OrderedDict(edge_index_dict)
        self.y = labels.to(next(iter(self.x.values())).device).to(torch.int64).contiguous()
        self.node_type = node_type
        self.lr = lr
        self.use_cuda_graph = use_cuda_graph
        self.opt: Optional[FlatAdam] = None
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        self._loss = None
        self._out = None
        self._emb = None
        self.launches_per_step = 0
        self._stage = None
        self._prefetched = False
        self._refresh_graph = None

    # -- one optimisation step, eager ------------------------------------------------------------
    def _step_eager(self):
        self.opt.zero_grad()
        emb, out = self.model(self.x, self.ei)
        pending = []
        loss = AF.nll_loss(out[0][self.node_type], self.y, self.group,
                           None if self.ctx is None else self.ctx.num_nodes_global[self.node_type],
                           pending)
        loss.backward()
        if self.group is not None:          # every rank holds d(global loss)/dW of its own rows
            from .dist import all_reduce_
            all_reduce_(self.opt.grad, self.group)
        self.opt.step()
        for w in pending:                   # the loss value's all-reduce ran beside the backward
            w.wait()
        return loss, emb, out

    def _lazy_init(self):
        if self.opt is None:
            self.model.train()
            with torch.no_grad():                       # materialise lazy weights (:146-147)
                self.model(self.x, self.ei)
            self.opt = FlatAdam(self.model.parameters(), lr=self.lr)
            self.opt.flatten()
            if self.group is not None:      # replicated weights: rank 0's initialisation
                from .dist import broadcast_
                broadcast_(self.opt.flat, self.group)

    def train_step(self) -> torch.Tensor:
        """``hetero_training()``: returns the (device) loss of this step."""
        self._lazy_init()
        self.model.train()
        if not self.use_cuda_graph or getattr(self, '_eager_only', False):
            self._loss, self._emb, self._out = self._step_eager()
            return self._loss
        if self._graph is None:
            snap = _StateSnapshot(self.model, self.opt)
            self._capture()
            snap.restore()
        self._graph.replay()
        return self._loss

    def _capture(self):
        from ._lib import launch_count
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            # the first step also flattens the parameters (FlatAdam) and builds the plan: keep
            # every allocation and host sync out of the capture
            for _ in range(2):
                self._step_eager()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._graph = torch.cuda.CUDAGraph()
        n0 = launch_count()
        # NCCL's watchdog thread polls events while this thread captures: thread-local error mode
        mode = 'thread_local' if self.group is not None else 'global'
        with torch.cuda.graph(self._graph, capture_error_mode=mode):
            self._loss, self._emb, self._out = self._step_eager()
        self.launches_per_step = launch_count() - n0

    # -- fresh inputs from the host (the reference passes x_dict / edge_index_dict every call) ------
    def update_inputs(self, host_x, host_ei):
        """Copy a new epoch's features and edge lists from (pinned) host memory into the static
        device tensors, re-sort the CSR/CSC in place and re-check the one-hot assumption on the
        device.  Nothing synchronises; ``verify_inputs()`` reads the flags back."""
        for k, v in host_x.items():
            if self._keeps_identity(k, v):
                continue
            self.x[k].copy_(v, non_blocking=True)
        for k, v in host_ei.items():
            self.ei[k].copy_(v, non_blocking=True)
        self._inputs_changed()

    def _keeps_identity(self, k, v) -> bool:
        """``v`` is a data.Identity marker for a type whose features already are the declared
        identity of that size: nothing to copy.  A marker for a type that holds dense features
        cannot be honoured in place (the captured step was planned for dense features)."""
        from .data import Identity, is_declared_identity
        if not isinstance(v, Identity):
            if is_declared_identity(self.x[k]):
                raise ValueError(f"x['{k}'] was declared Identity({self.x[k].shape[0]}); dense "
                                 f"features need a new trainer")
            return False
        if not (is_declared_identity(self.x[k]) and self.x[k].shape[0] == v.n):
            raise ValueError(f"x['{k}'] = {v!r} but the trainer holds dense features of shape "
                             f"{tuple(self.x[k].shape)}: re-create the trainer")
        return True

    # -- the same, one step ahead: host -> device copies overlap the previous step -----------------
    def prefetch_inputs(self, host_x, host_ei):
        """Start copying the NEXT step's inputs from pinned host memory into device staging
        buffers on a copy stream (returns at once; the running step is not disturbed).
        ``consume_prefetched()`` moves them into the state of the captured step.  Two staging sets
        alternate, so the copy for step i+2 may start as soon as the one for step i+1 is done
        (it only waits until step i's set has been read out)."""
        dev = next(iter(self.x.values())).device
        from .data import is_declared_identity
        if self._stage is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._stages = [(OrderedDict((k, torch.empty_like(v)) for k, v in self.x.items()
                                         if not is_declared_identity(v)),
                             OrderedDict((k, torch.empty_like(v)) for k, v in self.ei.items()))
                            for _ in range(2)]
            self._stage = self._stages[0]
            self._ev_staged = [torch.cuda.Event(), torch.cuda.Event()]
            self._ev_consumed = [None, None]
            self._pf_issued = 0
            self._pf_queue = []
            self._refresh_graphs = [None, None]
            self._slot_flags = [None, None]
        if len(self._pf_queue) >= 2:
            raise RuntimeError('prefetch_inputs(): two prefetched steps are already waiting')
        slot = self._pf_issued % 2
        self._pf_issued += 1
        stage = self._stages[slot]
        with torch.cuda.stream(self._copy_stream):
            if self._ev_consumed[slot] is not None:     # this set was read out
                self._copy_stream.wait_event(self._ev_consumed[slot])
            for k, v in host_x.items():
                if self._keeps_identity(k, v):
                    continue
                stage[0][k].copy_(v, non_blocking=True)
            for k, v in host_ei.items():
                stage[1][k].copy_(v, non_blocking=True)
            self._ev_staged[slot].record(self._copy_stream)
        self._pf_queue.append(slot)
        self._prefetched = True

    def consume_prefetched(self):
        """The oldest prefetched inputs become the state of the captured step (waits for the host
        copy on the device, not on the host)."""
        if not getattr(self, '_prefetched', False) or not self._pf_queue:
            raise RuntimeError('consume_prefetched() without prefetch_inputs()')
        slot = self._pf_queue.pop(0)
        self._stage = self._stages[slot]
        cur = torch.cuda.current_stream()
        cur.wait_event(self._ev_staged[slot])
        fast = getattr(self, '_plan', None) is not None and \
            not getattr(self, '_eager_only', False) and \
            (self.ctx is None or self.ctx.halo is None)
        if not fast:
            for k, v in self._stage[0].items():
                self.x[k].copy_(v, non_blocking=True)
            for k, v in self._stage[1].items():
                self.ei[k].copy_(v, non_blocking=True)
            self._inputs_changed()
        elif self._refresh_graphs[slot] is not None:
            self._refresh_graphs[slot].replay()
            self._id_flags = self._slot_flags[slot]
        else:
            self._refresh_from_stage()
            if self.use_cuda_graph:
                # the ~40 launches of the refresh (feature copies, batched radix sort, one-hot
                # checks) as ONE graph launch from now on (one graph per staging set)
                g = torch.cuda.CUDAGraph()
                mode = 'thread_local' if self.group is not None else 'global'
                with torch.cuda.graph(g, capture_error_mode=mode):
                    self._refresh_from_stage()
                g.replay()            # (the flag tensors of the captured run hold values now)
                self._refresh_graphs[slot] = g
                self._slot_flags[slot] = self._id_flags
                self._refresh_graph = g
        ev = torch.cuda.Event()
        ev.record(cur)
        self._ev_consumed[slot] = ev
        self._prefetched = bool(self._pf_queue)

    def _refresh_from_stage(self):
        """Prefetched inputs -> the state the captured step reads: features are copied into the
        static tensors; the CSR / CSC are re-sorted STRAIGHT from the staged edge lists (the step
        never reads ``self.ei`` itself -- 22 device-to-device copies less per step); the one-hot
        checks re-run.  No allocation that outlives the call, no host synchronisation:
        capturable."""
        for k, v in self._stage[0].items():
            self.x[k].copy_(v, non_blocking=True)
        self._plan.rebuild_(self._stage[1])
        self._refresh_identity_flags()

    def _inputs_changed(self):
        from .hetero import _is_identity_input
        num_nodes = {t: v.shape[0] for t, v in self.x.items()}
        num_dst = None
        if self.ctx is not None:
            num_nodes, num_dst = self.ctx.plan_rows(num_nodes)
        if any(type(m).__name__ == 'GATConv' for m in self.model.modules()):
            # GATConv sorts self-loop augmented edge lists whose LENGTH depends on the data
            # (functional.GATPlan): no in-place re-sort under a captured step; steps that follow an
            # input update re-plan and run eagerly
            self._graph = None
            self._plan = None
            self._eager_only = True
        else:
            self._plan = get_plan(self.ei, num_nodes, num_dst=num_dst)   # version bump -> re-sort
        if self.ctx is not None and self.ctx.halo is not None:
            self.ctx.halo.refresh(self.x)
        self._refresh_identity_flags()

    def _refresh_identity_flags(self):
        from .hetero import _is_identity_input
        self._id_flags = {}
        from .data import is_declared_identity
        for t, v in self.x.items():
            if is_declared_identity(v):
                continue
            if v.dim() == 2 and v.shape[0] == v.shape[1] and v.shape[0] >= 2:
                was_identity = self._identity_seen.get(t)
                if was_identity is None:
                    was_identity = self._identity_seen[t] = _is_identity_input(v)
                if was_identity:
                    self._id_flags[t] = ops.is_identity(v)

    def verify_inputs(self):
        """Synchronising checks of the last ``update_inputs``: node ids in range, and features
        that were one-hot when the step was planned / captured are still one-hot."""
        if getattr(self, '_plan', None) is not None:
            self._plan.check()
        for t, f in self._id_flags.items():
            if int(f.item()) != 1:
                raise RuntimeError(f"x['{t}'] is no longer an identity matrix: re-create the "
                                   f"trainer (the captured step assumed one-hot features)")

    def verify_inputs_async(self):
        """The checks of ``verify_inputs`` without stalling the stream: the flags are copied to
        pinned host memory behind the work already queued; the returned callable waits for that
        copy only and raises like ``verify_inputs``.  Lets a caller queue the next step before it
        looks at the previous one."""
        flags = []
        if getattr(self, '_plan', None) is not None:
            flags += [b[4].reshape(1).to(torch.int32) for b in self._plan._buffers]
        n_err = len(flags)
        names = list(self._id_flags.keys())
        flags += [self._id_flags[t].reshape(1).to(torch.int32) for t in names]
        if not flags:
            return lambda: None
        dev_flags = torch.cat(flags)
        host = torch.empty(dev_flags.shape, dtype=torch.int32, pin_memory=True)
        host.copy_(dev_flags, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()

        def check():
            ev.synchronize()
            if any(int(v) != 0 for v in host[:n_err]):
                raise IndexError('edge_index contains node ids outside [0, num_nodes)')
            for t, v in zip(names, host[n_err:]):
                if int(v) != 1:
                    raise RuntimeError(f"x['{t}'] is no longer an identity matrix: re-create the "
                                       f"trainer (the captured step assumed one-hot features)")
        return check

    # -- evaluation --------------------------------------------------------------------------------
    @torch.no_grad()
    def evaluate(self, x_dict=None, edge_index_dict=None, labels=None):
        """``hetero_test()`` on one graph: (loss, accuracy); BatchNorm in eval mode, dropout as
        traced.  On the trainer's own graph (no arguments) with ``use_cuda_graph``
        the evaluation forward is captured once and replayed, like the training step (its ~60
        launches are otherwise bound by the Python launch path: 3-4 ms instead of 0.5)."""
        own = x_dict is None and edge_index_dict is None and labels is None
        if own and self.use_cuda_graph and not getattr(self, '_eager_only', False):
            self._lazy_init()
            if getattr(self, '_eval_graph', None) is None:
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    self._evaluate_eager(None, None, None)           # plans, workspaces
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                mode = 'thread_local' if self.group is not None else 'global'
                with torch.cuda.graph(g, capture_error_mode=mode):
                    self._eval_out = self._evaluate_eager(None, None, None)
                self._eval_graph = g
            self._eval_graph.replay()
            return self._eval_out
        return self._evaluate_eager(x_dict, edge_index_dict, labels)

    def _evaluate_eager(self, x_dict, edge_index_dict, labels):
        self._check_foreign_graph(x_dict, edge_index_dict)
        self.model.eval()
        x = self.x if x_dict is None else x_dict
        ei = self.ei if edge_index_dict is None else edge_index_dict
        y = self.y if labels is None else labels.to(torch.int64)
        _, out = self.model(x, ei)
        logp = out[0][self.node_type]
        loss = AF.nll_loss(logp, y, self.group)        # mean over the rows of all ranks
        self.model.train()
        if self.group is None:
            return loss, _accuracy(logp, y)
        from .dist import all_reduce_
        # (torch.full, not torch.tensor: no host -> device copy, so the call can be captured)
        hits = torch.stack([logp.argmax(dim=1).eq(y).sum().to(torch.float64),
                            torch.full((), float(y.numel()), dtype=torch.float64, device=y.device)])
        hits = all_reduce_(hits, self.group)
        return loss, (hits[0] / hits[1]).to(torch.float32)

    def _check_foreign_graph(self, x_dict, edge_index_dict):
        """Multi-GPU: the boundary-row exchange and the partial-sum relations are those of the
        TRAINING partition; another graph (the reference's validation / test graphs,
        train_gnn_embeddings.py:57-58) needs its own partition.  On the block partition every rank
        evaluates its own block of the other graph, which is what this accepts."""
        if (x_dict is not None or edge_index_dict is not None) and self.ctx is not None and \
                (self.ctx.halo is not None or self.ctx.partial or self.ctx.scatter):
            raise NotImplementedError(
                'evaluate() / embeddings() on a graph other than the training partition: build a '
                'GraphPartition + partition_context for that graph and a GNNTrainer on it (the '
                "training partition's boundary lists and global in-degrees do not apply)")

    @torch.no_grad()
    def embeddings(self, x_dict=None, edge_index_dict=None) -> Dict[str, torch.Tensor]:
        """``save_embeddings()``: deep copy, eval mode, forward; returns the embedding dict (the
        reference saves ``emb['artwork']``; ``emb['style']`` / ``emb['genre']`` feed the heads)."""
        self._check_foreign_graph(x_dict, edge_index_dict)
        clone = copy.deepcopy(self.model)
        clone.eval()
        emb, _ = clone(self.x if x_dict is None else x_dict,
                       self.ei if edge_index_dict is None else edge_index_dict)
        return emb


class HeadTrainer:
    """One fused step per mini-batch on device-resident (or freshly copied) features."""

    def __init__(self, head: torch.nn.Module, kind: str = 'multitask', lr: float = 3e-4,
                 w_style=None, w_genre=None, group=None, use_cuda_graph: bool = False,
                 precision: str = 'fp32'):
        """``group``: batch-sharded data parallel -- ``step`` gets this rank's shard of the batch;
        weights are replicated (rank 0's), gradients all-reduced (one NCCL call on the arena).
        ``precision``: 'fp32' = exact float32 arithmetic (parity rel 1e-5); 'bf16' = the fused
        tensor-core step (``train_step_tc``: bf16 operands, float32 accumulation, rel 2e-2 -- the
        precision class of the reference's fp16 autocast, train_new_multimodal_multitask.py:76)."""
        assert kind in ('multitask', 'projector')
        assert precision in ('fp32', 'bf16')
        self.head = head
        self.kind = kind
        self.precision = precision
        self.opt = FlatAdam(head.parameters(), lr=lr).flatten()
        self.w_style, self.w_genre = w_style, w_genre
        self.group = group
        self.world = 1
        if group is not None:
            import torch.distributed as dist
            from .dist import broadcast_, enable_peer_allreduce
            self.world = dist.get_world_size(group)
            broadcast_(self.opt.flat, group)
            if self.opt.flat.is_cuda:
                enable_peer_allreduce(group, self.opt.flat.device)
        # CUDA graph: a head step is ~20 small launches, i.e. launch-bound from Python; the batch is
        # copied into static buffers and the captured step replayed (one batch shape per trainer)
        self.use_cuda_graph = use_cuda_graph
        self._graph = None
        self._static = None
        self._loss = None

    def step(self, feat, *rest) -> torch.Tensor:
        if not self.use_cuda_graph:
            return self._step_eager(feat, *rest)
        batch = (feat, *rest)
        if self._static is not None and any(a.shape != b.shape or a.dtype != b.dtype
                                            for a, b in zip(batch, self._static)):
            self._graph = self._static = None           # new batch shape: capture again
        if self._graph is None:
            self._static = [torch.empty_like(t, device=self.opt.flat.device) for t in batch]
            for st, t in zip(self._static, batch):
                st.copy_(t, non_blocking=True)
            snap = _StateSnapshot(self.head, self.opt)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):                      # warm-up steps are real steps
                    self._step_eager(*self._static)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self._graph = torch.cuda.CUDAGraph()
            mode = 'thread_local' if self.group is not None else 'global'
            with torch.cuda.graph(self._graph, capture_error_mode=mode):
                self._loss = self._step_eager(*self._static)
            snap.restore()
            self._graph.replay()
            return self._loss
        for st, t in zip(self._static, batch):
            st.copy_(t, non_blocking=True)
        self._graph.replay()
        return self._loss

    # -- the next batch one step ahead: its host -> device copy overlaps the running step -----------
    def prefetch(self, feat, *rest):
        """Start copying the NEXT mini-batch from (pinned) host memory into device staging buffers
        on a copy stream; returns at once.  ``step_prefetched()`` runs the step on it.  (What a
        ``DataLoader(pin_memory=True)`` + ``.to(device, non_blocking=True)`` loop does for the
        reference, src/train_new_multimodal_multitask.py:62-90.)"""
        batch = (feat, *rest)
        dev = self.opt.flat.device
        stage = getattr(self, '_stage', None)
        if stage is None or any(a.shape != b.shape or a.dtype != b.dtype
                                for a, b in zip(batch, stage)):
            self._stage = [torch.empty_like(t, device=dev) for t in batch]
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._ev_staged = torch.cuda.Event()
            self._ev_consumed = None
        with torch.cuda.stream(self._copy_stream):
            if self._ev_consumed is not None:           # the staging buffers were read out
                self._copy_stream.wait_event(self._ev_consumed)
            for st, t in zip(self._stage, batch):
                st.copy_(t, non_blocking=True)
            self._ev_staged.record(self._copy_stream)
        self._staged = True

    def step_prefetched(self) -> torch.Tensor:
        if not getattr(self, '_staged', False):
            raise RuntimeError('step_prefetched() without prefetch()')
        cur = torch.cuda.current_stream()
        cur.wait_event(self._ev_staged)
        self._staged = False
        ready = self.use_cuda_graph and self._graph is not None and all(
            a.shape == b.shape and a.dtype == b.dtype for a, b in zip(self._stage, self._static))
        if not ready:
            loss = self.step(*self._stage)              # first batch / new shape / eager trainer
            self._ev_consumed = torch.cuda.Event()
            self._ev_consumed.record(cur)
            return loss
        for st, t in zip(self._static, self._stage):
            st.copy_(t, non_blocking=True)
        self._ev_consumed = torch.cuda.Event()
        self._ev_consumed.record(cur)                   # the next prefetch may start under the step
        self._graph.replay()
        return self._loss

    def _step_tc(self, feat, *rest) -> torch.Tensor:
        """One fused kernel for forward + loss + backward, then (all-reduce,) Adam: every parameter
        of the head receives its gradient from the step, so the arena is written, not cleared."""
        if self.kind == 'multitask':
            emb_s, emb_g, y_s, y_g = rest
            loss = self.head.train_step_tc(feat, emb_s, emb_g, y_s, y_g, self.w_style, self.w_genre,
                                           group=self.group, accumulate=False)
        else:
            (target,) = rest
            loss = self.head.train_step_tc(feat, target, feat.shape[0] * self.world,
                                           accumulate=False)
        if self.group is not None:
            from .dist import small_all_reduce_      # 180 KB of head gradients: peer-memory kernel
            small_all_reduce_(self.opt.grad, self.group)
            loss = small_all_reduce_(loss.detach().clone(), self.group)
        self.opt.step()
        return loss

    def _step_eager(self, feat, *rest) -> torch.Tensor:
        self.head.train()
        if self.precision == 'bf16':
            ok = self.head.tc_supported(feat, rest[0]) if self.kind == 'multitask' else \
                self.head.tc_supported(feat)
            if not ok:
                raise ValueError("precision='bf16': these shapes are not supported by agx_head_step "
                                 "(feature width a multiple of 64, total width of 128, <= 64 classes)")
            return self._step_tc(feat, *rest)
        self.opt.zero_grad()
        if self.kind == 'multitask':
            emb_s, emb_g, y_s, y_g = rest
            out = self.head(feat, emb_s, emb_g)
            loss = multitask_loss(out, y_s, y_g, self.w_style, self.w_genre, self.group)
        else:
            (target,) = rest             # equal shards: global rows = world * local rows
            loss = projector_loss(self.head(feat), target, feat.shape[0] * self.world)
        loss.backward()
        if self.group is not None:
            from .dist import small_all_reduce_
            small_all_reduce_(self.opt.grad, self.group)
            if self.kind == 'projector':
                loss = small_all_reduce_(loss.detach().clone(), self.group)
        self.opt.step()
        return loss
