"""The layer-level C ABI used from plain C (tests/c/test_sage_layer.c): a host program without
Python or torch links against libagx.so, builds a graph plan, runs one SAGEConv / GraphConv
relation forward + backward and checks the results against loops written from the operator's
definition.  Without a GPU the program is compiled and linked only (every symbol it uses is
declared in include/agx.h and exported by the library)."""
import os
import subprocess

import pytest
import torch

import util
import mmac_b200  # noqa: F401
from mmac_b200 import _build

SRC = os.path.join(util.ROOT, 'tests', 'c', 'test_sage_layer.c')
CUDA = '/usr/local/cuda'


def _compile(out_path):
    lib_dir = os.path.dirname(_build.build())
    cmd = ['gcc', '-std=c99', '-Wall', '-Werror', SRC, '-I', os.path.join(util.ROOT, 'include'),
           '-I', os.path.join(CUDA, 'include'), '-L', lib_dir, '-lagx',
           '-L', os.path.join(CUDA, 'lib64'), '-lcudart', '-lm', f'-Wl,-rpath,{lib_dir}',
           f'-Wl,-rpath,{os.path.join(CUDA, "lib64")}', '-o', out_path]
    subprocess.check_call(cmd)
    return out_path


def test_c_host_program_compiles_and_links(tmp_path):
    exe = _compile(str(tmp_path / 'test_sage_layer'))
    assert os.path.exists(exe)
    used = subprocess.check_output(['nm', '-u', exe]).decode()
    for sym in ('agx_graph_plan_create', 'agx_graph_plan_relation', 'agx_graph_plan_destroy',
                'agx_sage_layer_workspace_bytes', 'agx_sage_layer_fwd', 'agx_sage_layer_bwd'):
        assert sym in used, sym


@pytest.mark.gpu
def test_c_host_program_runs_a_sage_layer(tmp_path):
    exe = _compile(str(tmp_path / 'test_sage_layer'))
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=300)
    out = r.stdout.decode()
    print(out)
    assert r.returncode == 0 and out.strip().endswith('OK'), out


@pytest.mark.gpu
def test_graph_plan_handle_matches_python_plan():
    """agx_graph_plan_create (C-owned CSR / CSC buffers) against the Python-side HeteroPlan on the
    same edge lists: identical row pointers, neighbour order and long-row flags."""
    import ctypes as C
    from mmac_b200 import _lib as L
    from mmac_b200 import synth
    from mmac_b200.graph import get_plan
    dev = 'cuda:0'
    g = synth.make_artgraph('small', features='dense')
    data = mmac_b200.ToUndirected()(g)
    ei = {k: v.to(dev) for k, v in data.edge_index_dict.items()}
    n = data.num_nodes_dict
    keys = list(ei.keys())[:8]
    pl = get_plan({k: ei[k] for k in keys}, n)
    arr = (L.EdgeList * len(keys))()
    keep = []
    for i, (s, r, d) in enumerate(keys):
        dst, src = ei[(s, r, d)][1].contiguous(), ei[(s, r, d)][0].contiguous()
        keep += [dst, src]
        arr[i] = L.EdgeList(dst.data_ptr(), src.data_ptr(), dst.numel(), n[d], n[s])
    handle = C.c_void_p()
    L.check(L.lib().agx_graph_plan_create(arr, len(keys), L.stream_ptr(), C.byref(handle)),
            'agx_graph_plan_create')
    try:
        for i, k in enumerate(keys):
            pr = L.PlanRel()
            L.check(L.lib().agx_graph_plan_relation(handle, i, C.byref(pr)), 'agx_graph_plan_relation')
            rel = pl[k]
            assert (pr.n_src, pr.n_dst, pr.n_edges) == (rel.n_src, rel.n_dst, rel.n_edges)
            assert bool(pr.long_rows) == rel.csr.long_rows and bool(pr.t_long_rows) == rel.csc.long_rows

            def view(addr, count, dtype):
                buf = torch.empty(count, dtype=dtype, device=dev)
                C.cdll.LoadLibrary('libcudart.so').cudaMemcpy(
                    C.c_void_p(buf.data_ptr()), C.c_void_p(addr), C.c_size_t(count * buf.element_size()), 3)
                return buf
            assert torch.equal(view(pr.rowptr, rel.n_dst + 1, torch.int32), rel.csr.rowptr)
            assert torch.equal(view(pr.col, rel.n_edges, torch.int32), rel.csr.col)
            assert torch.equal(view(pr.cnt, rel.n_dst, torch.float32), rel.csr.cnt)
            assert torch.equal(view(pr.t_rowptr, rel.n_src + 1, torch.int32), rel.csc.rowptr)
            assert torch.equal(view(pr.t_col, rel.n_edges, torch.int32), rel.csc.col)
    finally:
        L.lib().agx_graph_plan_destroy(handle)
