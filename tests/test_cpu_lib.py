"""CPU suite, part 2: the C-ABI library loads and exports every symbol include/agx.h declares, the
ctypes mirrors of the descriptor structs have the C sizes, and the host-side module logic (tracing,
state-dict layout, loud failure without CUDA) works.  No kernel is launched here."""
import copy
import ctypes
import os
import re
import subprocess
import sys
import tempfile

import pytest
import torch

import util
import mmac_b200 as agx
from mmac_b200 import _lib as L
from oracle import graph_oracle as go

HEADER = os.path.join(util.ROOT, 'include', 'agx.h')


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(agx_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    lib = agx.lib()
    declared = _declared_functions()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f'{name} declared in agx.h but not exported by libagx.so'
    assert sorted(L.EXPORTED) == declared, set(L.EXPORTED) ^ set(declared)
    assert lib.agx_version() == 100
    buf = ctypes.create_string_buffer(4096)
    assert lib.agx_kernel_inventory(buf, 4096) == 0
    kernels = buf.value.decode().split(',')
    assert 'agg_rows' in kernels and 'radix_scatter' in kernels and 'gemm_f32' in kernels


def test_ctypes_structs_match_c_layout():
    names = {'agx_edge_list_t': L.EdgeList, 'agx_rel_t': L.Rel, 'agx_row_group_t': L.RowGroup,
             'agx_chunk_seg_t': L.ChunkSeg, 'agx_gemm_seg_t': L.GemmSeg,
             'agx_gemm_problem_t': L.GemmProblem, 'agx_sum_desc_t': L.SumDesc,
             'agx_bn_desc_t': L.BnDesc, 'agx_bn_bwd_desc_t': L.BnBwdDesc,
             'agx_colsum_desc_t': L.ColsumDesc, 'agx_gat_rel_t': L.GatRel,
             'agx_sddmm_seg_t': L.SddmmSeg, 'agx_head_t': L.Head, 'agx_plan_rel_t': L.PlanRel,
             'agx_sage_layer_t': L.SageLayer}
    body = '\n'.join(f'printf("{n} %zu\\n", sizeof({n}));' for n in names)
    consts = ['AGX_MAX_CSR_RELS', 'AGX_MAX_REL_PER_GROUP', 'AGX_MAX_GROUPS', 'AGX_MAX_CHUNK_SEGS',
              'AGX_CHUNK_EDGES', 'AGX_MAX_GEMM_PROBLEMS', 'AGX_MAX_GEMM_SEGS', 'AGX_MAX_TENSORS',
              'AGX_MAX_GAT_RELS', 'AGX_MAX_SDDMM_SEGS', 'AGX_GAT_LONG_ROW', 'AGX_MAX_HEADS']
    body += '\n' + '\n'.join(f'printf("{c} %d\\n", (int){c});' for c in consts)
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, 't.c')
        with open(src, 'w') as fh:
            fh.write(f'#include <stdio.h>\n#include "{HEADER}"\nint main(void){{{body} return 0;}}')
        exe = os.path.join(d, 't')
        subprocess.check_call(['gcc', '-std=c99', '-Wall', '-Werror', src, '-o', exe])
        out = dict(line.split() for line in subprocess.check_output([exe]).decode().splitlines())
    for n, cls in names.items():
        assert int(out[n]) == ctypes.sizeof(cls), (n, out[n], ctypes.sizeof(cls))
    assert int(out['AGX_MAX_CSR_RELS']) == L.MAX_CSR_RELS
    assert int(out['AGX_MAX_REL_PER_GROUP']) == L.MAX_REL_PER_GROUP
    assert int(out['AGX_MAX_GROUPS']) == L.MAX_GROUPS
    assert int(out['AGX_MAX_CHUNK_SEGS']) == L.MAX_CHUNK_SEGS
    assert int(out['AGX_CHUNK_EDGES']) == L.CHUNK_EDGES
    assert int(out['AGX_MAX_GEMM_PROBLEMS']) == L.MAX_GEMM_PROBLEMS
    assert int(out['AGX_MAX_GEMM_SEGS']) == L.MAX_GEMM_SEGS
    assert int(out['AGX_MAX_TENSORS']) == L.MAX_TENSORS
    assert int(out['AGX_MAX_GAT_RELS']) == L.MAX_GAT_RELS
    assert int(out['AGX_MAX_SDDMM_SEGS']) == L.MAX_SDDMM_SEGS
    assert int(out['AGX_GAT_LONG_ROW']) == L.GAT_LONG_ROW
    assert int(out['AGX_MAX_HEADS']) == L.MAX_HEADS


def test_invalid_arguments_return_error_codes_not_crashes():
    lib = agx.lib()
    assert lib.agx_aggregate_rows(None, 0, 128, 0, None) == -1
    assert b'n_groups' in lib.agx_last_error()
    assert lib.agx_gemm_grouped(None, 1, None, 1, None) == -1
    assert lib.agx_csr_build(None, 0, None, None, None, None, None, None, 0, None) == -1
    assert lib.agx_adam_step(None, None, None, None, 10, 0.1, 0.9, 0.999, 1e-8, 0.0, None, None) == -1
    assert lib.agx_csr_workspace_bytes(1000, 10) > 4 * 4 * 1000
    assert lib.agx_gat_edge_softmax(None, 0, 0.2, None) == -1
    assert lib.agx_gat_edge_softmax_bwd(None, 1, 0.2, None) == -1
    assert lib.agx_sddmm(None, 1, 128, None) == -1


def test_product_model_has_reference_state_dict_layout():
    g, ei, md = util.undirected_graph('tiny')
    prod = agx.HeteroSGNN(agx.SAGEConv, torch.nn.ReLU(), 'sum', 128, 32, md, 2, 0.4, True, False)
    orc = go.HeteroSGNNOracle(go.SAGEConv, torch.nn.ReLU(), 'sum', 128, 32, md, 2, 0.4, True, False)
    assert list(prod.state_dict().keys()) == list(orc.state_dict().keys())
    k = 'gnn.convs.0.artist__field_rel__field.lin_l.weight'
    assert isinstance(prod.state_dict()[k], torch.nn.parameter.UninitializedParameter)
    with torch.no_grad():
        orc(g.x_dict, ei)
    util.copy_state(orc, prod)
    assert prod.state_dict()[k].shape == (128, g.num_nodes_dict['artist'])
    clone = copy.deepcopy(prod)                       # save_embeddings deep-copies the model
    assert torch.equal(clone.state_dict()[k], prod.state_dict()[k])
    pg = agx.HeteroSGNN(agx.GraphConv, torch.nn.ReLU(), 'sum', 128, 18, md, 2, 0.4, True, False)
    og = go.HeteroSGNNOracle(go.GraphConv, torch.nn.ReLU(), 'sum', 128, 18, md, 2, 0.4, True, False)
    assert list(pg.state_dict().keys()) == list(og.state_dict().keys())
    assert any('.lin_rel.' in n for n in pg.state_dict())


def test_dead_activation_is_not_scheduled():
    _, _, md = util.undirected_graph('tiny')
    prod = agx.HeteroSGNN(agx.SAGEConv, torch.nn.ReLU(), 'sum', 128, 32, md, 2, 0.4, True, False)
    hm = prod.gnn
    drops = [n for n in hm._graph.nodes if n.op == 'call_function' and 'dropout' in str(n.target)]
    assert len(drops) == 2
    assert drops[0].name not in hm._live and drops[1].name in hm._live
    assert drops[0].kwargs.get('training', True) is True         # baked in by tracing
    fused = [v for v in hm._fusion.values() if v[1] is not None]
    assert len(fused) == 1 and abs(fused[0][1][1] - 0.4) < 1e-12


def test_no_cpu_execution_path():
    g, ei, md = util.undirected_graph('tiny')
    prod = agx.HeteroSGNN(agx.SAGEConv, torch.nn.ReLU(), 'sum', 128, 32, md, 2, 0.0, True, False)
    with pytest.raises(agx.AgxError):
        prod(g.x_dict, ei)                            # CPU tensors: must fail loudly
    head = agx.NewMultiModalMultiTaskHead(128, {'style': 32, 'genre': 18}, 0.0)
    with pytest.raises(agx.AgxError):
        head(torch.randn(4, 768), torch.randn(4, 128), torch.randn(4, 128))
    with pytest.raises(RuntimeError):
        agx.FlatAdam(head.parameters(), lr=1e-3).step()


def test_heads_state_dict_keys():
    h = agx.NewMultiModalMultiTaskHead(128, {'style': 32, 'genre': 18}, 0.3, feat_size=2048)
    assert sorted(h.state_dict()) == ['class_genre.1.bias', 'class_genre.1.weight',
                                      'class_style.1.bias', 'class_style.1.weight']
    assert h.class_style[1].weight.shape == (32, 2176)
    assert sorted(agx.LabelProjectorHead(128).state_dict()) == ['encoder.bias', 'encoder.weight']
    assert sorted(agx.NewMultiModalSingleTaskHead(128, 18, 0.1).state_dict()) == \
        ['classifier.1.bias', 'classifier.1.weight']


def test_oracle_is_not_imported_by_the_product():
    pkg = os.path.join(util.ROOT, 'multi-modal-art-classifier_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', txt, flags=re.M), f


def test_bench_reference_arm_json_contract():
    """`bench.py --impl reference` (the CPU oracle timed on the host cores) prints ONE JSON line
    with the contract's keys; bounded here to the tiny graph."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(util.ROOT, 'bench.py'), '--impl', 'reference',
                        '--cpu-size', 'tiny', '--size', 'tiny', '--steps', '1', '--warmup', '0'],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    lines = [ln for ln in r.stdout.decode().splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'hetero_gnn_aggregated_edges_per_s'
    assert d['unit'] == 'edges/s' and d['higher_is_better'] is True and d['value'] > 0
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1
    assert d['e2e'] == {'value': d['value'], 'unit': 'edges/s', 'h2d_bytes_per_step': 0,
                        'd2h_bytes_per_step': 0}
    for k in ('n_gpus', 'steps', 'warmup', 'ms_per_step', 'scaling', 'vs_baseline', 'dtype', 'data',
              'config'):
        assert k in d
