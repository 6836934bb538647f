"""ctypes binding of ``libagx.so`` (include/agx.h).

The library is the product: there is no CPU or eager-PyTorch fallback.  If ``libagx.so`` cannot be
loaded (and cannot be built with nvcc) importing this module's ``lib()`` raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

from . import _build

c_i32, c_i64, c_f32 = C.c_int32, C.c_int64, C.c_float
vp = C.c_void_p

MAX_CSR_RELS = 40
MAX_REL_PER_GROUP = 8
MAX_GROUPS = 24
MAX_CHUNK_SEGS = 24
CHUNK_EDGES = 128
MAX_GEMM_PROBLEMS = 24
MAX_GEMM_SEGS = 64
MAX_TENSORS = 48
MAX_GAT_RELS = 24
MAX_SDDMM_SEGS = 24
GAT_LONG_ROW = 1024
F32, BF16, F64 = 0, 1, 2


class EdgeList(C.Structure):
    _fields_ = [('keys', vp), ('vals', vp), ('n_edges', c_i64), ('n_rows', c_i64),
                ('n_cols', c_i64)]


class Rel(C.Structure):
    _fields_ = [('rowptr', vp), ('col', vp), ('x', vp), ('ldx', c_i64), ('row_cnt', vp),
                ('nbr_scale', vp), ('edge_w', vp), ('edge_w_idx', vp)]


class RowGroup(C.Structure):
    _fields_ = [('out', vp), ('ldo', c_i64), ('n_rows', c_i32), ('n_rel', c_i32),
                ('accumulate', c_i32), ('relu_dmask', c_i32), ('rel', Rel * MAX_REL_PER_GROUP),
                ('bias', vp)]


class ChunkSeg(C.Structure):
    _fields_ = [('rel', Rel), ('out', vp), ('ldo', c_i64), ('n_rows', c_i32), ('n_edges', c_i32),
                ('frag', vp), ('counters', vp)]


class GemmSeg(C.Structure):
    _fields_ = [('A', vp), ('a_rs', c_i64), ('a_cs', c_i64), ('B', vp), ('b_rs', c_i64),
                ('b_cs', c_i64), ('A_mask', vp), ('B_mask', vp), ('K', c_i32), ('pad_', c_i32)]


class GemmProblem(C.Structure):
    _fields_ = [('C', vp), ('ldc', c_i64), ('bias', vp), ('row_scale', vp), ('M', c_i32),
                ('N', c_i32), ('accumulate', c_i32), ('seg_begin', c_i32), ('seg_count', c_i32),
                ('split_k', c_i32), ('partial', vp), ('skip_flag', vp)]


class TransposeDesc(C.Structure):
    _fields_ = [('inp', vp), ('ld_in', c_i64), ('out', vp), ('ld_out', c_i64), ('rows', c_i32),
                ('cols', c_i32), ('accumulate', c_i32), ('pad_', c_i32)]


class SumDesc(C.Structure):
    _fields_ = [('out', vp), ('inp', vp * 8), ('n_in', c_i32), ('numel', c_i64), ('bias', vp),
                ('bias_F', c_i64)]


class GatRel(C.Structure):
    _fields_ = [('rowptr', vp), ('col', vp), ('a_l', vp), ('a_r', vp), ('alpha', vp),
                ('dalpha', vp), ('de', vp), ('da_r', vp), ('long_rows', vp), ('n_rows', c_i32),
                ('n_long', c_i32)]


class SddmmSeg(C.Structure):
    _fields_ = [('row', vp), ('col', vp), ('a', vp), ('lda', c_i64), ('b', vp), ('ldb', c_i64),
                ('out', vp), ('n_edges', c_i32), ('pad_', c_i32)]


class BnDesc(C.Structure):
    _fields_ = [('x', vp), ('y', vp), ('y_act', vp), ('dmask', vp), ('weight', vp), ('bias', vp),
                ('running_mean', vp), ('running_var', vp), ('save_mean', vp), ('save_invstd', vp),
                ('n_rows', c_i32), ('drop_p', c_f32), ('drop_seed', vp), ('drop_offset', c_i64)]


class BnBwdDesc(C.Structure):
    _fields_ = [('x', vp), ('y', vp), ('dy', vp), ('dy_act', vp), ('dmask', vp), ('weight', vp),
                ('save_mean', vp), ('save_invstd', vp), ('dx', vp), ('dweight', vp), ('dbias', vp),
                ('n_rows', c_i32), ('act_scale', c_f32)]


class ColsumDesc(C.Structure):
    _fields_ = [('x', vp), ('ldx', c_i64), ('out', vp), ('n_rows', c_i32), ('F', c_i32),
                ('accumulate', c_i32), ('pad_', c_i32)]


MAX_HEADS = 4
HEAD_CE, HEAD_SMOOTH_L1 = 0, 1


class Head(C.Structure):
    _fields_ = [('part', vp * 2), ('ld', c_i64 * 2), ('width', c_i32 * 2), ('C', c_i32),
                ('loss', c_i32), ('weight', vp), ('ldw', c_i64), ('bias', vp), ('mask', vp),
                ('ld_mask', c_i64), ('labels', vp), ('class_w', vp), ('coef', c_f32),
                ('inv_count', c_f32), ('target', vp), ('ld_target', c_i64), ('logits', vp),
                ('ld_logits', c_i64), ('d_weight', vp), ('ld_dw', c_i64), ('d_bias', vp)]


class PlanRel(C.Structure):
    _fields_ = [('rowptr', vp), ('col', vp), ('cnt', vp), ('t_rowptr', vp), ('t_col', vp),
                ('n_src', c_i32), ('n_dst', c_i32), ('n_edges', c_i32), ('long_rows', c_i32),
                ('t_long_rows', c_i32), ('pad_', c_i32)]


class SageLayer(C.Structure):
    _fields_ = [('rel', PlanRel), ('mean', c_i32), ('f_src', c_i32), ('f_dst', c_i32),
                ('out_channels', c_i32), ('x_src', vp), ('ld_src', c_i64), ('x_dst', vp),
                ('ld_dst', c_i64), ('w_l', vp), ('b_l', vp), ('w_r', vp)]


_SIGS = {
    'agx_version': (C.c_int, []),
    'agx_last_error': (C.c_char_p, []),
    'agx_launch_count': (C.c_uint64, []),
    'agx_kernel_inventory': (C.c_int, [C.c_char_p, C.c_size_t]),
    'agx_csr_workspace_bytes': (C.c_size_t, [c_i64, c_i64]),
    'agx_csr_build': (C.c_int, [C.POINTER(EdgeList), C.c_int, vp, vp, vp, vp, vp, vp, C.c_size_t,
                                vp]),
    'agx_coalesce_workspace_bytes': (C.c_size_t, [c_i64]),
    'agx_coalesce_undirected': (C.c_int, [vp, vp, c_i64, c_i64, vp, vp, vp, vp, C.c_size_t, vp]),
    'agx_aggregate_rows': (C.c_int, [C.POINTER(RowGroup), C.c_int, C.c_int, C.c_int, vp]),
    'agx_chunk_frag_floats': (C.c_size_t, [c_i64, C.c_int]),
    'agx_chunk_counters': (C.c_size_t, [c_i64]),
    'agx_aggregate_chunks': (C.c_int, [C.POINTER(ChunkSeg), C.c_int, C.c_int, C.c_int, vp]),
    'agx_gemm_grouped': (C.c_int, [C.POINTER(GemmProblem), C.c_int, C.POINTER(GemmSeg), C.c_int,
                                   vp]),
    'agx_sum_arrays': (C.c_int, [C.POINTER(SumDesc), C.c_int, vp]),
    'agx_bn_workspace_floats': (C.c_size_t, [c_i64, C.c_int, C.c_int]),
    'agx_bn_forward': (C.c_int, [C.POINTER(BnDesc), C.c_int, C.c_int, C.c_int, c_f32, c_f32, vp,
                                 C.c_size_t, vp]),
    'agx_bn_backward': (C.c_int, [C.POINTER(BnBwdDesc), C.c_int, C.c_int, C.c_int, vp, C.c_size_t,
                                  vp]),
    'agx_bn_forward_phase': (C.c_int, [C.POINTER(BnDesc), C.c_int, C.c_int, C.c_int, c_f32, c_f32,
                                       vp, C.c_size_t, C.c_int, vp, vp, vp]),
    'agx_bn_backward_phase': (C.c_int, [C.POINTER(BnBwdDesc), C.c_int, C.c_int, C.c_int, vp,
                                        C.c_size_t, C.c_int, vp, vp, vp]),
    'agx_colsum_workspace_floats': (C.c_size_t, [c_i64, C.c_int, C.c_int]),
    'agx_colsum': (C.c_int, [C.POINTER(ColsumDesc), C.c_int, vp, C.c_size_t, vp]),
    'agx_log_softmax_nll': (C.c_int, [vp, c_i64, c_i32, c_i32, vp, vp, vp, c_i64, vp, vp, vp]),
    'agx_nll_forward': (C.c_int, [vp, c_i64, c_i32, c_i32, vp, vp, vp, vp, vp]),
    'agx_nll_backward': (C.c_int, [c_i32, c_i32, vp, vp, vp, vp, c_f32, vp, c_i64, vp]),
    'agx_loss_finish': (C.c_int, [vp, c_f32, vp, C.c_int, vp]),
    'agx_log_softmax_nll_bwd': (C.c_int, [vp, c_i64, c_i32, c_i32, vp, vp, vp, vp, c_f32, vp,
                                          c_i64, vp, c_i64, vp]),
    'agx_adam_step': (C.c_int, [vp, vp, vp, vp, c_i64, c_f32, c_f32, c_f32, c_f32, c_f32, vp, vp]),
    'agx_dropout_mask': (C.c_int, [vp, c_i64, c_f32, vp, vp]),
    'agx_smooth_l1_workspace_floats': (C.c_size_t, []),
    'agx_smooth_l1': (C.c_int, [vp, vp, c_i64, vp, vp, vp, vp]),
    'agx_gat_edge_softmax': (C.c_int, [C.POINTER(GatRel), C.c_int, c_f32, vp]),
    'agx_gat_edge_softmax_bwd': (C.c_int, [C.POINTER(GatRel), C.c_int, c_f32, vp]),
    'agx_sddmm': (C.c_int, [C.POINTER(SddmmSeg), C.c_int, c_i32, vp]),
    'agx_graph_plan_create': (C.c_int, [C.POINTER(EdgeList), C.c_int, vp, C.POINTER(vp)]),
    'agx_graph_plan_relation': (C.c_int, [vp, C.c_int, C.POINTER(PlanRel)]),
    'agx_graph_plan_destroy': (C.c_int, [vp]),
    'agx_sage_layer_workspace_bytes': (C.c_size_t, [C.POINTER(SageLayer)]),
    'agx_sage_layer_fwd': (C.c_int, [C.POINTER(SageLayer), vp, c_i64, C.c_int, vp, vp, C.c_size_t,
                                     vp]),
    'agx_sage_layer_bwd': (C.c_int, [C.POINTER(SageLayer), vp, vp, c_i64, vp, vp, vp, vp, c_i64, vp,
                                     c_i64, C.c_int, vp, C.c_size_t, vp]),
    'agx_peer_allreduce_buffer_bytes': (C.c_size_t, [C.c_size_t, C.c_int]),
    'agx_peer_allreduce': (C.c_int, [vp, C.c_int, C.c_int, vp, vp, c_i64, C.c_int, vp, C.c_size_t,
                                     vp]),
    'agx_head_step_workspace_bytes': (C.c_size_t, [C.POINTER(Head), C.c_int, c_i32]),
    'agx_head_step_prepare': (C.c_int, [C.POINTER(Head), C.c_int, c_i32, vp, vp, C.c_size_t, vp]),
    'agx_head_step': (C.c_int, [C.POINTER(Head), C.c_int, c_i32, c_f32, vp, vp, vp, C.c_int, vp,
                                C.c_size_t, vp]),
    'agx_mse': (C.c_int, [vp, vp, c_i64, vp, vp, vp, vp]),
    'agx_tanh': (C.c_int, [vp, vp, c_i64, vp]),
    'agx_tanh_bwd': (C.c_int, [vp, vp, vp, c_i64, vp]),
    'agx_sgd_step': (C.c_int, [vp, vp, vp, c_i64, c_f32, c_f32, c_f32, vp]),
    'agx_fill_f32': (C.c_int, [vp, c_i64, c_f32, vp]),
    'agx_scale_mask': (C.c_int, [vp, vp, vp, c_i64, vp]),
    'agx_gather_rows': (C.c_int, [vp, c_i64, vp, c_i64, c_i32, vp, c_i64, vp]),
    'agx_is_identity': (C.c_int, [vp, c_i64, c_i32, vp, vp, vp]),
    'agx_transpose': (C.c_int, [vp, c_i64, c_i32, c_i32, vp, c_i64, vp, vp]),
    'agx_transpose_batched': (C.c_int, [C.POINTER(TransposeDesc), C.c_int, vp]),
    'agx_pack_rows': (C.c_int, [vp, c_i64, vp, c_i32, c_i32, vp, vp]),
    'agx_unpack_rows_add': (C.c_int, [vp, c_i64, vp, c_i32, c_i32, vp, vp]),
}

EXPORTED = tuple(_SIGS.keys())

_lib: Optional[C.CDLL] = None


class AgxError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load (building first if the in-tree .so is missing or stale and nvcc is present)."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    if not _build.is_current():
        try:
            _build.build()
        except Exception as e:  # noqa: BLE001
            if not os.path.exists(path):
                raise AgxError(
                    f'libagx.so is missing and could not be built ({e}); the sm_100a CUDA library '
                    f'is the only execution path of this package') from e
    h = C.CDLL(path)
    for name, (res, args) in _SIGS.items():
        fn = getattr(h, name)          # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = h
    return h


def check(rc: int, what: str = ''):
    if rc != 0:
        msg = lib().agx_last_error().decode(errors='replace')
        raise AgxError(f'{what or "agx call"} failed ({rc}): {msg}')


def ptr(t: Optional[torch.Tensor]):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def launch_count() -> int:
    """Kernels launched (or captured) by libagx.so so far in this process."""
    return int(lib().agx_launch_count())


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise AgxError(f'{name} must be a CUDA tensor: this package has no CPU execution path')


def compute_device() -> torch.device:
    """Where host inputs are staged for computation: the current CUDA device.  There is no CPU
    execution path, so this raises without one."""
    if not torch.cuda.is_available():
        raise AgxError('no CUDA device: the kernels of libagx.so are the only execution path')
    return torch.device('cuda', torch.cuda.current_device())
