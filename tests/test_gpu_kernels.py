"""GPU parity tests of the individual agx kernels (through the C ABI) against the CPU oracle.
Integer / index work is compared bit-exactly; float32 work to rel 1e-5 (north_star)."""
import numpy as np
import pytest
import torch

import util
from util import RTOL_F32, rel_err
import mmac_b200 as agx
from mmac_b200 import ops, synth
from mmac_b200 import functional as AF
from oracle import graph_oracle as go

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _rand_rel(gen, n_src, n_dst, e, empty_tail=0):
    hi = max(n_dst - empty_tail, 1)
    return torch.stack([torch.randint(0, n_src, (e,), generator=gen),
                        torch.randint(0, hi, (e,), generator=gen)])


# ---------------------------------------------------------------- K1
def _check_csr(csr, keys, vals, n_rows):
    rowptr, col, eid = go.csr_build(keys.numpy(), vals.numpy(), n_rows)
    assert np.array_equal(csr.rowptr.cpu().numpy().astype(np.int64), rowptr)
    assert np.array_equal(csr.col.cpu().numpy().astype(np.int64), col)
    assert np.array_equal(csr.eid.cpu().numpy().astype(np.int64), eid)
    deg = np.maximum(np.diff(rowptr), 1).astype(np.float32)
    assert np.array_equal(csr.cnt.cpu().numpy(), deg)


@pytest.mark.parametrize('case', ['one', 'many', 'skew', 'empty_rel', 'single_row', 'big'])
def test_csr_build_bit_exact(case):
    gen = torch.Generator().manual_seed(11)
    if case == 'one':
        rels = [(_rand_rel(gen, 50, 30, 500, empty_tail=4), 50, 30)]
    elif case == 'many':
        rels = [(_rand_rel(gen, 10 + 7 * i, 5 + 13 * i, 40 * i + (i % 3)), 10 + 7 * i, 5 + 13 * i)
                for i in range(1, 18)]
    elif case == 'skew':       # few destination rows, long rows (artwork -> style), duplicates
        ei = torch.stack([torch.randint(0, 4000, (20000,), generator=gen),
                          torch.multinomial(torch.tensor([0.5, 0.3, 0.1, 0.05, 0.05]), 20000,
                                            replacement=True, generator=gen)])
        rels = [(ei, 4000, 5), (torch.stack([ei[1], ei[0]]), 5, 4000)]
    elif case == 'empty_rel':
        rels = [(_rand_rel(gen, 9, 9, 30), 9, 9), (torch.zeros(2, 0, dtype=torch.int64), 5, 7),
                (_rand_rel(gen, 3, 700, 100), 3, 700)]
    elif case == 'single_row':
        rels = [(torch.stack([torch.arange(300), torch.zeros(300, dtype=torch.int64)]), 300, 1)]
    else:
        rels = [(_rand_rel(gen, 100000, 70000, 600000), 100000, 70000),
                (_rand_rel(gen, 33, 100000, 250000), 33, 100000)]
    lists = [(ei[1].to(DEV), ei[0].to(DEV), nd, ns) for ei, ns, nd in rels]
    built = ops.csr_build(lists)
    for csr, (ei, ns, nd) in zip(built, rels):
        _check_csr(csr, ei[1], ei[0], nd)


def test_csr_build_rejects_out_of_range():
    ei = torch.tensor([[0, 1, 2], [0, 5, 1]])
    with pytest.raises(IndexError):
        ops.csr_build([(ei[1].to(DEV), ei[0].to(DEV), 3, 3)])


def test_plan_csr_and_csc_on_artgraph():
    g, ei, md = util.undirected_graph('small')
    plan = agx.HeteroPlan({k: v.to(DEV) for k, v in ei.items()}, g.num_nodes_dict)
    assert list(plan.rels.keys()) == list(ei.keys())
    for k, v in ei.items():
        r = plan[k]
        _check_csr(r.csr, v[1], v[0], r.n_dst)
        _check_csr(r.csc, v[0], v[1], r.n_src)


@pytest.mark.parametrize('size', ['tiny', 'small'])
def test_to_undirected_bit_exact(size):
    g = synth.make_artgraph(size)
    ref = go.to_undirected(g.edge_index_dict)
    got = agx.to_undirected_dict({k: v.to(DEV) for k, v in g.edge_index_dict.items()},
                                 g.num_nodes_dict)
    assert list(got.keys()) == list(ref.keys())
    for k in ref:
        assert torch.equal(got[k].cpu(), ref[k]), k
    # ToUndirected transform on the container, CPU tensors in -> CPU tensors out
    data = agx.ToUndirected()(synth.make_artgraph(size))
    assert data.edge_types == list(ref.keys())
    assert all(torch.equal(data[k].edge_index, ref[k]) for k in ref)


def test_coalesce_edge_cases():
    row = torch.tensor([3, 3, 0, 2, 2, 2], dtype=torch.int64)
    col = torch.tensor([3, 1, 0, 1, 1, 0], dtype=torch.int64)     # self loops + duplicates
    ref = go.to_undirected({('a', 'r', 'a'): torch.stack([row, col])})[('a', 'r', 'a')]
    got = ops.coalesce_undirected(row.to(DEV), col.to(DEV), 4)
    assert torch.equal(got.cpu(), ref)
    got0 = ops.coalesce_undirected(torch.zeros(0, dtype=torch.int64, device=DEV),
                                   torch.zeros(0, dtype=torch.int64, device=DEV), 4)
    assert got0.shape == (2, 0)


# ---------------------------------------------------------------- K2 / K3
@pytest.mark.parametrize('F', [128, 64, 32, 18, 200, 260])
@pytest.mark.parametrize('reduce', ['mean', 'add'])
def test_aggregate_rows_matches_scatter(F, reduce):
    gen = torch.Generator().manual_seed(F)
    n_src, n_dst, e = 300, 500, 2500
    ei = _rand_rel(gen, n_src, n_dst, e, empty_tail=20)
    x = torch.randn(n_src, F, generator=gen)
    ref = go.propagate(x, ei, n_dst, reduce)
    (csr,) = ops.csr_build([(ei[1].to(DEV), ei[0].to(DEV), n_dst, n_src)])
    out = torch.full((n_dst, F), float('nan'), device=DEV)
    ops.aggregate_rows([(out, [ops.RelArg(csr, x.to(DEV), mean_rows=reduce == 'mean')], False)], F)
    # same summation order as CPU scatter_add_ (edge order inside each row): bit-exact
    assert torch.equal(out.cpu(), ref)


@pytest.mark.parametrize('F', [128, 32, 18])
@pytest.mark.parametrize('reduce', ['mean', 'add'])
def test_aggregate_chunks_long_rows(F, reduce):
    gen = torch.Generator().manual_seed(100 + F)
    n_src, n_dst, e = 5000, 40, 30000
    p = 1.0 / torch.arange(1, n_dst - 4, dtype=torch.float64)          # Zipf rows, 5 empty rows
    dst = torch.multinomial(p / p.sum(), e, replacement=True, generator=gen)
    ei = torch.stack([torch.randint(0, n_src, (e,), generator=gen), dst])
    x = torch.randn(n_src, F, generator=gen)
    ref = go.propagate(x, ei, n_dst, reduce)
    (csr,) = ops.csr_build([(ei[1].to(DEV), ei[0].to(DEV), n_dst, n_src)])
    out = torch.full((n_dst, F), float('nan'), device=DEV)
    ops.aggregate_chunks([(out, ops.RelArg(csr, x.to(DEV), mean_rows=reduce == 'mean'))], F)
    assert rel_err(out, ref) <= RTOL_F32
    assert torch.all(out[-5:] == 0)
    out2 = torch.empty_like(out)
    ops.aggregate_chunks([(out2, ops.RelArg(csr, x.to(DEV), mean_rows=reduce == 'mean'))], F)
    assert torch.equal(out, out2)                                      # run-to-run reproducible


@pytest.mark.parametrize('F', [128, 64, 32, 200, 260, 18])
def test_aggregate_chunks_many_segments_empty_rows_and_counter_reset(F):
    """One launch over several relations: leading / interior / trailing rows without edges (the
    kernel, not a memset, clears them), rows spanning hundreds of chunks next to 1-edge rows, a
    relation without any edge; repeated launches reuse the self-resetting arrival counters and
    must be bit-identical."""
    gen = torch.Generator().manual_seed(300 + F)
    specs = [(3000, 60, 50000, 'zipf'), (64, 7, 9000, 'uniform'), (500, 300, 1200, 'sparse'),
             (40, 25, 0, 'uniform'), (2000, 3, 130, 'uniform'), (900, 1, 129, 'uniform')]
    segs, refs, keep = [], [], []
    for n_src, n_dst, e, kind in specs:
        if kind == 'zipf':                      # rows 0-2 and the last 4 rows empty, rest Zipf
            p = torch.zeros(n_dst, dtype=torch.float64)
            p[3:n_dst - 4] = 1.0 / torch.arange(1, n_dst - 6, dtype=torch.float64)
            dst = torch.multinomial(p / p.sum(), e, replacement=True, generator=gen)
        elif kind == 'sparse':                  # most rows empty, runs of empty rows > 32
            live = torch.randperm(n_dst, generator=gen)[:40]
            dst = live[torch.randint(0, 40, (e,), generator=gen)]
        else:
            dst = torch.randint(0, n_dst, (e,), generator=gen)
        ei = torch.stack([torch.randint(0, n_src, (e,), generator=gen), dst])
        x = torch.randn(n_src, F, generator=gen)
        mean = (len(segs) % 2 == 0)
        refs.append(go.propagate(x, ei, n_dst, 'mean' if mean else 'add'))
        (csr,) = ops.csr_build([(ei[1].to(DEV), ei[0].to(DEV), n_dst, n_src)])
        csr.max_degree = int(torch.bincount(dst, minlength=n_dst).max()) if e else 0
        out = torch.full((n_dst, F), float('nan'), device=DEV)
        segs.append((out, ops.RelArg(csr, x.to(DEV), mean_rows=mean)))
        keep.append((x, ei))
    ops.aggregate_chunks(segs, F)
    first = [o.clone() for o, _ in segs]
    for (out, _), ref in zip(segs, refs):
        assert not torch.isnan(out).any()
        assert rel_err(out, ref) <= RTOL_F32
        empty = ref.abs().sum(1) == 0
        assert torch.all(out.cpu()[empty] == 0)
    for _ in range(3):
        for o, _a in segs:
            o.fill_(float('nan'))
        ops.aggregate_chunks(segs, F)
        for (o, _a), f in zip(segs, first):
            assert torch.equal(o, f)


def test_aggregate_multi_relation_group_and_transpose_scale():
    gen = torch.Generator().manual_seed(4)
    n_dst, F = 700, 128
    rels, xs = [], []
    for n_src, e in ((32, 700), (18, 650), (90, 3000)):
        rels.append(_rand_rel(gen, n_src, n_dst, e))
        xs.append(torch.randn(n_src, F, generator=gen))
    ref = sum(go.propagate(x, ei, n_dst, 'mean') for x, ei in zip(xs, rels))
    csrs = ops.csr_build([(ei[1].to(DEV), ei[0].to(DEV), n_dst, x.shape[0]) for x, ei in zip(xs, rels)])
    out = torch.empty(n_dst, F, device=DEV)
    ops.aggregate_rows([(out, [ops.RelArg(c, x.to(DEV), mean_rows=True) for c, x in zip(csrs, xs)],
                         False)], F)
    assert rel_err(out, ref) <= RTOL_F32
    # transpose of scatter-mean: gradient wrt x_src of sum(out * g)
    ei, x = rels[2], xs[2].clone().requires_grad_(True)
    gout = torch.randn(n_dst, F, generator=gen)
    (go.propagate(x, ei, n_dst, 'mean') * gout).sum().backward()
    fwd, bwd = ops.csr_build([(ei[1].to(DEV), ei[0].to(DEV), n_dst, 90),
                              (ei[0].to(DEV), ei[1].to(DEV), 90, n_dst)])
    dx = torch.empty(90, F, device=DEV)
    arg = ops.RelArg(bwd, gout.to(DEV), nbr_scale=fwd.cnt)
    ops.aggregate_chunks([(dx, arg)], F)
    assert rel_err(dx, x.grad) <= RTOL_F32
    dx2 = torch.empty(90, F, device=DEV)
    ops.aggregate_rows([(dx2, [arg], False)], F)
    assert rel_err(dx2, x.grad) <= RTOL_F32


def test_aggregate_bf16():
    gen = torch.Generator().manual_seed(8)
    n_src, n_dst, e, F = 400, 300, 4000, 128
    ei = _rand_rel(gen, n_src, n_dst, e)
    x = torch.randn(n_src, F, generator=gen).bfloat16()
    ref = go.propagate(x.float(), ei, n_dst, 'mean')
    (csr,) = ops.csr_build([(ei[1].to(DEV), ei[0].to(DEV), n_dst, n_src)])
    out = torch.empty(n_dst, F, device=DEV, dtype=torch.bfloat16)
    ops.aggregate_rows([(out, [ops.RelArg(csr, x.to(DEV), mean_rows=True)], False)], F)
    assert rel_err(out.float(), ref) <= util.RTOL_BF16


@pytest.mark.parametrize('F', [128, 64, 200])
def test_aggregate_chunks_bf16(F):
    """bfloat16 feature storage through the edge-balanced kernel (float32 accumulation, rel 2e-2):
    long rows crossing CTAs, rows inside a chunk, empty rows."""
    gen = torch.Generator().manual_seed(500 + F)
    n_src, n_dst, e = 4000, 50, 60000
    p = torch.zeros(n_dst, dtype=torch.float64)
    p[2:n_dst - 3] = 1.0 / torch.arange(1, n_dst - 4, dtype=torch.float64)
    dst = torch.multinomial(p / p.sum(), e, replacement=True, generator=gen)
    ei = torch.stack([torch.randint(0, n_src, (e,), generator=gen), dst])
    x = torch.randn(n_src, F, generator=gen).bfloat16()
    for reduce in ('mean', 'add'):
        ref = go.propagate(x.float(), ei, n_dst, reduce)
        (csr,) = ops.csr_build([(ei[1].to(DEV), ei[0].to(DEV), n_dst, n_src)])
        out = torch.full((n_dst, F), float('nan'), device=DEV, dtype=torch.bfloat16)
        ops.aggregate_chunks([(out, ops.RelArg(csr, x.to(DEV), mean_rows=reduce == 'mean'))], F)
        assert not torch.isnan(out.float()).any()
        assert rel_err(out.float(), ref) <= util.RTOL_BF16
        assert torch.all(out[:2].float() == 0) and torch.all(out[-3:].float() == 0)


# ---------------------------------------------------------------- K4
@pytest.mark.parametrize('M,K,N', [(300, 128, 128), (1000, 200, 32), (77, 5, 18), (513, 129, 130),
                                   (4, 768, 50)])
def test_gemm_nt_nn_tn(M, K, N):
    gen = torch.Generator().manual_seed(M + K + N)
    A = torch.randn(M, K, generator=gen)
    W = torch.randn(N, K, generator=gen)
    b = torch.randn(N, generator=gen)
    Ad, Wd, bd = A.to(DEV), W.to(DEV), b.to(DEV)
    C = torch.empty(M, N, device=DEV)
    gb = ops.GemmBatch()
    gb.add(C, [(Ad, Wd.t())], bias=bd)                               # X W^T + b
    G = torch.randn(M, N, generator=gen)
    Gd = G.to(DEV)
    dX = torch.empty(M, K, device=DEV)
    gb.add(dX, [(Gd, Wd)])                                            # G W
    dW = torch.empty(N, K, device=DEV)
    gb.add(dW, [(Gd.t(), Ad)], split_k=3)                             # G^T X, split reduction
    dW1 = torch.zeros(N, K, device=DEV) + 1.0
    gb.add(dW1, [(Gd.t(), Ad)], accumulate=True)
    gb.run()
    A64, W64, G64 = A.double(), W.double(), G.double()
    assert rel_err(C, A64 @ W64.t() + b.double()) <= RTOL_F32
    assert rel_err(dX, G64 @ W64) <= RTOL_F32
    assert rel_err(dW, G64.t() @ A64) <= RTOL_F32
    assert rel_err(dW1, G64.t() @ A64 + 1.0) <= RTOL_F32


@pytest.mark.parametrize('M,N,K', [(128, 128, 20000), (32, 128, 116475), (128, 32, 5000),
                                   (100, 72, 3001), (128, 128, 2048), (8, 8, 4100),
                                   (64, 128, 1000003)])      # the 16x graph: > 140 * 768 rows
def test_gemm_tensor_core_long_k_weight_gradient(M, N, K):
    """dW = dOut^T X over a very long row dimension takes the split-K tcgen05 path (MN-major
    operands, accumulators drained every 256 rows): float32 accuracy (rel 1e-5 against float64),
    same answer as the exact FFMA kernel to that tolerance, reproducible, accumulate / bias /
    strided views handled by the split reduction."""
    gen = torch.Generator().manual_seed(M + N + K)
    dout = torch.randn(K, M, generator=gen)
    x = torch.randn(K, N + 8, generator=gen)[:, 4:4 + N]           # a column slice: ld != N
    c0 = torch.randn(M, N, generator=gen)
    ref = dout.double().t() @ x.double()
    dd, xd = dout.to(DEV), x.to(DEV)
    sk = ops.split_k_for(K)
    outs = []
    for _ in range(2):
        dw = torch.full((M, N), float('nan'), device=DEV)
        dwa = c0.to(DEV).clone()
        gb = ops.GemmBatch()
        gb.add(dw, [(dd.t(), xd)], split_k=sk)
        gb.add(dwa, [(dd.t(), xd)], split_k=sk, accumulate=True)
        gb.run()
        outs.append((dw, dwa))
    assert rel_err(outs[0][0], ref) <= RTOL_F32
    assert rel_err(outs[0][1], ref + c0.double()) <= RTOL_F32
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize('M,N,Ks', [(4096, 128, [128]), (5000, 128, [128, 128, 64]),
                                    (3001, 32, [128]), (2049, 18, [128, 96]), (1500, 72, [200])])
def test_gemm_tensor_core_3xtf32(M, N, Ks):
    """Tall X W^T problems take the tcgen05 path (agx_gemm_tc.cu): three TF32 products per
    element must reproduce float32 accuracy (rel 1e-5 against float64)."""
    gen = torch.Generator().manual_seed(M + N)
    As = [torch.randn(M, K, generator=gen) for K in Ks]
    Ws = [torch.randn(N, K, generator=gen) / 8 for K in Ks]
    b = torch.randn(N, generator=gen)
    C0 = torch.randn(M, N, generator=gen)
    ref = sum(a.double() @ w.double().t() for a, w in zip(As, Ws)) + b.double()
    Ad = [a.to(DEV) for a in As]
    Wd = [w.to(DEV) for w in Ws]
    C = torch.full((M, N), float('nan'), device=DEV)
    gb = ops.GemmBatch()
    gb.add(C, [(a, w.t()) for a, w in zip(Ad, Wd)], bias=b.to(DEV))
    Cacc = C0.to(DEV).clone()
    gb.add(Cacc, [(a, w.t()) for a, w in zip(Ad, Wd)], bias=b.to(DEV), accumulate=True)
    gb.run()
    assert rel_err(C, ref) <= RTOL_F32
    assert rel_err(Cacc, ref + C0.double()) <= RTOL_F32
    # an ill-conditioned case: large common offset, the lo terms matter
    A2 = (torch.randn(M, Ks[0], generator=gen) * 1e-3 + 7.0)
    W2 = torch.randn(N, Ks[0], generator=gen)
    C2 = torch.empty(M, N, device=DEV)
    gb = ops.GemmBatch()
    gb.add(C2, [(A2.to(DEV), W2.to(DEV).t())])
    gb.run()
    assert rel_err(C2, A2.double() @ W2.double().t()) <= RTOL_F32


def test_gemm_multi_segment_masks_and_slices():
    gen = torch.Generator().manual_seed(1)
    B, Fv, Fe, Cn = 200, 768, 128, 32
    feat, emb = torch.randn(B, Fv, generator=gen), torch.randn(B, Fe, generator=gen)
    W = torch.randn(Cn, Fv + Fe, generator=gen) / 30
    m1 = (torch.rand(B, Fv, generator=gen) > 0.3).float() / 0.7
    m2 = (torch.rand(B, Fe, generator=gen) > 0.3).float() / 0.7
    ref = torch.cat([feat * m1, emb * m2], 1).double() @ W.double().t()
    Wd = W.to(DEV)
    C = torch.empty(B, Cn, device=DEV)
    gb = ops.GemmBatch()
    gb.add(C, [(feat.to(DEV), Wd[:, :Fv].t(), m1.to(DEV), None),
               (emb.to(DEV), Wd[:, Fv:].t(), m2.to(DEV), None)])
    gb.run()
    assert rel_err(C, ref) <= RTOL_F32


# ---------------------------------------------------------------- small kernels
def test_dropout_mask_statistics_and_determinism():
    seed = torch.tensor([1234, 0], dtype=torch.int64, device=DEV)
    m = ops.dropout_mask((1000, 128), 0.4, seed)
    vals = torch.unique(m).cpu()
    assert vals.numel() == 2 and vals[0] == 0 and abs(float(vals[1]) - 1 / 0.6) < 1e-6
    keep = float((m > 0).float().mean())
    assert abs(keep - 0.6) < 0.01
    assert torch.equal(m, ops.dropout_mask((1000, 128), 0.4, seed))
    seed[1] += 32000
    assert not torch.equal(m, ops.dropout_mask((1000, 128), 0.4, seed))


def test_identity_detection_and_transpose():
    eye = torch.eye(300, device=DEV)
    assert int(ops.is_identity(eye).item()) == 1
    eye[17, 200] = 1e-3
    assert int(ops.is_identity(eye).item()) == 0
    assert int(ops.is_identity(torch.randn(5, 7, device=DEV)).item()) == 0
    a = torch.randn(70, 130, device=DEV)
    out = torch.empty(130, 70, device=DEV)
    ops.transpose_into(out, a)
    assert torch.equal(out, a.t())


def test_adam_matches_torch():
    gen = torch.Generator().manual_seed(2)
    p0 = torch.randn(1000, generator=gen)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=0.01)
    p = p0.to(DEV)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    step = torch.zeros(1, dtype=torch.int32, device=DEV)
    for it in range(5):
        g = torch.randn(1000, generator=gen)
        ref.grad = g.clone()
        opt.step()
        step += 1
        ops.adam_step(p, g.to(DEV), m, v, step, 0.01)
    assert rel_err(p, ref.data) <= RTOL_F32


def test_log_softmax_nll_and_ce_weighted():
    gen = torch.Generator().manual_seed(6)
    for n, c in ((500, 32), (333, 18), (64, 100), (501, 20), (77, 8)):   # c % 4 == 0, c <= 32: float4 kernels
        x = torch.randn(n, c, generator=gen) * 3
        y = torch.randint(0, c, (n,), generator=gen)
        w = torch.rand(c, generator=gen) + 0.5
        xr = x.clone().requires_grad_(True)
        ref = torch.nn.functional.cross_entropy(xr, y, weight=w) * 0.5
        ref.backward()
        xd = x.to(DEV).requires_grad_(True)
        loss = AF.cross_entropy(xd, y.to(DEV), w.to(DEV), coef=0.5)
        loss.backward()
        assert rel_err(loss, ref) <= RTOL_F32
        assert rel_err(xd.grad, xr.grad) <= RTOL_F32
        # log_softmax followed by the reference's F.nll_loss call shape
        xr2 = x.clone().requires_grad_(True)
        l2 = torch.nn.functional.nll_loss(torch.log_softmax(xr2, 1), y)
        l2.backward()
        xd2 = x.to(DEV).requires_grad_(True)
        lp = AF.log_softmax(xd2, 1)
        l2d = AF.nll_loss(lp, y.to(DEV))
        l2d.backward()
        assert rel_err(lp, torch.log_softmax(x, 1)) <= RTOL_F32
        assert rel_err(l2d, l2) <= RTOL_F32
        assert rel_err(xd2.grad, xr2.grad) <= RTOL_F32
        # an arbitrary upstream gradient on the log-probabilities (the un-fused backward)
        r = torch.randn(n, c, generator=gen)
        xr3 = x.clone().requires_grad_(True)
        (torch.log_softmax(xr3, 1) * r).sum().backward()
        xd3 = x.to(DEV).requires_grad_(True)
        (AF.log_softmax(xd3, 1) * r.to(DEV)).sum().backward()
        assert rel_err(xd3.grad, xr3.grad) <= RTOL_F32


def test_smooth_l1():
    gen = torch.Generator().manual_seed(7)
    o, t = torch.randn(64, 128, generator=gen) * 2, torch.randn(64, 128, generator=gen)
    orr = o.clone().requires_grad_(True)
    ref = torch.nn.functional.smooth_l1_loss(orr, t)
    ref.backward()
    od = o.to(DEV).requires_grad_(True)
    loss = AF.smooth_l1_loss(od, t.to(DEV))
    loss.backward()
    assert rel_err(loss, ref) <= RTOL_F32
    assert rel_err(od.grad, orr.grad) <= RTOL_F32


@pytest.mark.parametrize('training', [True, False])
def test_batch_norm_act_batched(training):
    gen = torch.Generator().manual_seed(12)
    sizes = [700, 32, 18, 129]
    F = 128
    xs = [torch.randn(n, F, generator=gen) * (1 + i) + 3 * i for i, n in enumerate(sizes)]
    bns = [torch.nn.BatchNorm1d(F) for _ in sizes]
    for i, bn in enumerate(bns):
        with torch.no_grad():
            bn.weight.uniform_(0.5, 1.5, generator=gen)
            bn.bias.uniform_(-0.5, 0.5, generator=gen)
            bn.running_mean.uniform_(-1, 1, generator=gen)
            bn.running_var.uniform_(0.5, 2, generator=gen)
        bn.train(training)
    import copy
    states = [copy.deepcopy(bn.state_dict()) for bn in bns]
    masks = [(torch.rand(n, F, generator=gen) > 0.4).float() / 0.6 for n in sizes]
    gys = [torch.randn(n, F, generator=gen) for n in sizes]
    gas = [torch.randn(n, F, generator=gen) for n in sizes]
    # reference
    refs = []
    for x, bn, mk, gy, ga in zip(xs, bns, masks, gys, gas):
        xr = x.clone().requires_grad_(True)
        y = bn(xr)
        a = torch.relu(y) * mk
        ((y * gy).sum() + (a * ga).sum()).backward()
        refs.append((y.detach(), a.detach(), xr.grad, bn.weight.grad.clone(), bn.bias.grad.clone(),
                     bn.running_mean.clone(), bn.running_var.clone()))
    # product, from the same initial parameters / running statistics
    bnd = []
    for st in states:
        b_dev = torch.nn.BatchNorm1d(F)
        b_dev.load_state_dict(st)
        bnd.append(b_dev.to(DEV).train(training))
    xd = [x.to(DEV).requires_grad_(True) for x in xs]
    spec = AF.BNSpec(n=len(sizes), F=F, training=training, momentum=0.1, eps=1e-5,
                     running=[(b.running_mean, b.running_var) for b in bnd], with_act=True,
                     dmasks=[m.to(DEV) for m in masks])
    res = AF.batch_norm_act(spec, xd, [b.weight for b in bnd], [b.bias for b in bnd])
    ys, acts = res[:len(sizes)], res[len(sizes):]
    loss = sum((y * gy.to(DEV)).sum() + (a * ga.to(DEV)).sum()
               for y, a, gy, ga in zip(ys, acts, gys, gas))
    loss.backward()
    for i, r in enumerate(refs):
        assert rel_err(ys[i], r[0]) <= RTOL_F32, i
        assert rel_err(acts[i], r[1]) <= RTOL_F32, i
        assert rel_err(xd[i].grad, r[2]) <= 5 * RTOL_F32, i
        assert rel_err(bnd[i].weight.grad, r[3]) <= 5 * RTOL_F32, i
        assert rel_err(bnd[i].bias.grad, r[4]) <= 5 * RTOL_F32, i
        assert rel_err(bnd[i].running_mean, r[5]) <= RTOL_F32, i
        assert rel_err(bnd[i].running_var, r[6]) <= RTOL_F32, i


@pytest.mark.parametrize('F', [128, 30])
def test_batch_norm_dropout_without_mask_tensor(F):
    """Dropout generated inside the normalising pass (agx_bn_desc_t.drop_seed / drop_offset) keeps
    exactly the elements agx_dropout_mask keeps for the same Philox state, y itself is not written
    (need_y=False), and the backward pass -- relu gate read from y_act, gradient scaled by
    1 / (1 - p), no mask tensor -- equals the path with the materialised mask bit for bit."""
    gen = torch.Generator().manual_seed(40 + F)
    sizes = [515, 32, 1201]
    p = 0.4
    xs = [torch.randn(n, F, generator=gen) * (1 + i) + i for i, n in enumerate(sizes)]
    gas = [torch.randn(n, F, generator=gen).to(DEV) for n in sizes]
    seed = torch.tensor([0x1234567, 977], dtype=torch.int64, device=DEV)
    padded = [((n * F + 3) // 4) * 4 for n in sizes]
    flat = ops.dropout_mask((sum(padded),), p, seed)
    offs = [sum(padded[:i]) for i in range(len(sizes))]
    masks = [flat[o:o + n * F].view(n, F) for o, n in zip(offs, sizes)]
    assert 0.5 < float(flat.count_nonzero()) / flat.numel() < 0.7

    def run(virtual):
        bns = [torch.nn.BatchNorm1d(F).to(DEV).train() for _ in sizes]
        for i, bn in enumerate(bns):
            with torch.no_grad():
                bn.weight.fill_(0.5 + 0.25 * i)
                bn.bias.fill_(0.1 * i - 0.1)
        xd = [x.to(DEV).requires_grad_(True) for x in xs]
        spec = AF.BNSpec(n=len(sizes), F=F, training=True, momentum=0.1, eps=1e-5,
                         running=[(b.running_mean, b.running_var) for b in bns], with_act=True,
                         dmasks=None if virtual else masks,
                         drop=[(seed, o) for o in offs] if virtual else None, drop_p=p,
                         need_y=not virtual)
        res = AF.batch_norm_act(spec, xd, [b.weight for b in bns], [b.bias for b in bns])
        acts = res[-len(sizes):]
        assert len(res) == (len(sizes) if virtual else 2 * len(sizes))
        sum((a * g).sum() for a, g in zip(acts, gas)).backward()
        return acts, [x.grad for x in xd], [b.weight.grad for b in bns], [b.bias.grad for b in bns]

    a0, dx0, dw0, db0 = run(False)
    a1, dx1, dw1, db1 = run(True)
    for i in range(len(sizes)):
        assert torch.equal(a0[i], a1[i]), i
        assert torch.equal(dx0[i], dx1[i]), i
        assert torch.equal(dw0[i], dw1[i]) and torch.equal(db0[i], db1[i]), i


def test_bn_training_rejects_single_row():
    x = torch.randn(1, 128, device=DEV)
    bn = torch.nn.BatchNorm1d(128).to(DEV)
    spec = AF.BNSpec(n=1, F=128, training=True, momentum=0.1, eps=1e-5,
                     running=[(bn.running_mean, bn.running_var)], with_act=False)
    with pytest.raises(agx.AgxError):
        AF.batch_norm_act(spec, [x], [bn.weight], [bn.bias])


def test_gather_pack_unpack():
    gen = torch.Generator().manual_seed(13)
    table = torch.randn(50, 128, generator=gen)
    idx = torch.randint(0, 50, (200,), generator=gen)
    out = agx.select_embeddings(table.to(DEV), idx)
    assert torch.equal(out.cpu(), table[idx])
    uniq = torch.randperm(50, generator=gen)[:20].to(torch.int32)
    packed = ops.pack_rows(table.to(DEV), uniq.to(DEV))
    assert torch.equal(packed.cpu(), table[uniq.long()])
    acc = table.to(DEV).clone()
    ops.unpack_rows_add_(acc, uniq.to(DEV), packed)
    ref = table.clone()
    ref[uniq.long()] *= 2
    assert torch.equal(acc.cpu(), ref)


# ---------------------------------------------------------------- GATConv kernels (agx_gat.cu)
def _cpu_csr(c):
    """Host copy of a device CSR (the torch restatements in cpu_shim.py work on CPU tensors)."""
    return ops.CSR(c.rowptr.cpu(), c.col.cpu(), c.eid.cpu(), c.cnt.cpu(), c.n_rows, c.n_cols,
                   c.n_edges, c.max_degree)


def _skewed_rel(gen, n_src, n_dst, e, hubs):
    """Random edges; ``hubs``: {row: edges} rows that receive that many extra edges; the last two
    rows stay empty."""
    src = [torch.randint(0, n_src, (e,), generator=gen)]
    dst = [torch.randint(0, max(n_dst - 2, 1), (e,), generator=gen)]
    for row, k in hubs.items():
        src.append(torch.randint(0, n_src, (k,), generator=gen))
        dst.append(torch.full((k,), row, dtype=torch.int64))
    ei = torch.stack([torch.cat(src), torch.cat(dst)])
    return ei[:, torch.randperm(ei.shape[1], generator=gen)]


def test_gat_edge_softmax_and_backward_hub_rows():
    """Rows of 0, a few, ~1000 and > AGX_GAT_LONG_ROW edges (warp and whole-CTA paths), two
    relations in one launch; forward and backward against the torch restatement."""
    import cpu_shim
    gen = torch.Generator().manual_seed(3)
    rels = [_skewed_rel(gen, 500, 40, 600, {1: 1000, 7: 1500, 8: 5000, 30: 1025}),
            _skewed_rel(gen, 64, 9, 50, {0: 3000}),
            # mostly rows of 0..8 edges (8 lanes per row), some of 9..32 and 33..100 (one warp)
            _skewed_rel(gen, 300, 1003, 2500, {5: 9, 6: 31, 7: 32, 8: 33, 500: 100, 1000: 8})]
    sizes = [(500, 40), (64, 9), (300, 1003)]
    csrs = ops.csr_build([(ei[1].to(DEV), ei[0].to(DEV), nd, ns) for ei, (ns, nd) in zip(rels, sizes)])
    dev_args, cpu_args = [], []
    for c, (ns, nd) in zip(csrs, sizes):
        a_l, a_r = torch.randn(ns, generator=gen) * 2, torch.randn(nd, generator=gen) * 2
        dal = torch.randn(c.n_edges, generator=gen)
        mk = lambda n: torch.empty(n, dtype=torch.float32, device=DEV)            # noqa: E731
        # relation 1 is launched WITHOUT its hub list (a warp then walks the hub row itself)
        dev_args.append(ops.GatArg(c, a_l.to(DEV), a_r.to(DEV), mk(c.n_edges), dal.to(DEV),
                                   mk(c.n_edges), mk(nd),
                                   long_rows=ops.gat_long_rows(c) if len(dev_args) != 1 else None))
        z = lambda n: torch.zeros(n)                                              # noqa: E731
        cpu_args.append(ops.GatArg(_cpu_csr(c), a_l, a_r, z(c.n_edges), dal, z(c.n_edges), z(nd)))
    assert sorted(dev_args[0].long_rows.cpu().tolist()) == [7, 8, 30]     # > 1024 edges
    assert dev_args[2].long_rows is None
    ops.gat_edge_softmax(dev_args, 0.2)
    cpu_shim._gat_edge_softmax(cpu_args, 0.2)
    for d, h in zip(dev_args, cpu_args):
        assert rel_err(d.alpha, h.alpha) <= RTOL_F32
        rows = cpu_shim._rows_of(h.csr)
        sums = torch.zeros(h.csr.n_rows, dtype=torch.float64).index_add_(0, rows, d.alpha.cpu().double())
        deg = (h.csr.rowptr[1:] - h.csr.rowptr[:-1])
        assert float((sums[deg > 0] - 1).abs().max()) <= 1e-5         # every row sums to one
        h.alpha.copy_(d.alpha.cpu())                 # same coefficients into the backward check
    ops.gat_edge_softmax(dev_args, 0.2, backward=True)
    cpu_shim._gat_edge_softmax(cpu_args, 0.2, backward=True)
    for d, h in zip(dev_args, cpu_args):
        assert rel_err(d.de, h.de) <= 2e-5
        assert float((d.da_r.cpu() - h.da_r).abs().max()) <= 2e-5 * float(h.de.abs().max()) * 10
    # reproducible: a second launch gives the same bits
    a0 = dev_args[0].alpha.clone()
    ops.gat_edge_softmax(dev_args, 0.2)
    assert torch.equal(a0, dev_args[0].alpha)


@pytest.mark.parametrize('F', [128, 32, 18, 64, 200])
def test_sddmm_matches_torch(F):
    gen = torch.Generator().manual_seed(F)
    segs_d, refs = [], []
    for (na, nb, e) in ((30, 500, 4000), (7, 9, 33), (5, 5, 0)):
        row = torch.randint(0, na, (e,), generator=gen).to(torch.int32)
        col = torch.randint(0, nb, (e,), generator=gen).to(torch.int32)
        a, b = torch.randn(na, F, generator=gen), torch.randn(nb, F, generator=gen)
        out = torch.full((max(e, 1),), 7.0, device=DEV)
        segs_d.append((row.to(DEV), col.to(DEV), a.to(DEV), b.to(DEV), out))
        refs.append((a.double()[row.long()] * b.double()[col.long()]).sum(1))
    ops.sddmm([s for s in segs_d if s[0].numel() > 0], F)
    for s, ref in zip(segs_d, refs):
        if ref.numel():
            assert rel_err(s[4][:ref.numel()], ref) <= RTOL_F32


@pytest.mark.parametrize('F', [128, 32, 18, 1])
def test_weighted_aggregation_rows_and_chunks(F):
    """Per-edge weights (directly and through an index array), the row-group bias, and the
    one-column aggregation used for the attention logit gradients."""
    import cpu_shim
    gen = torch.Generator().manual_seed(40 + F)
    n_src, n_dst = 300, 50
    ei = _skewed_rel(gen, n_src, n_dst, 400, {3: 700, 9: 2000})
    (csr,) = ops.csr_build([(ei[1].to(DEV), ei[0].to(DEV), n_dst, n_src)])
    E = csr.n_edges
    x = torch.randn(n_src, F, generator=gen)
    w = torch.rand(E, generator=gen)
    perm = torch.randperm(E, generator=gen).to(torch.int32)
    bias = torch.randn(F, generator=gen)
    h = _cpu_csr(csr)
    for idx in (None, perm):
        ref = cpu_shim._rel_sum(ops.RelArg(h, x, edge_w=w, edge_w_idx=idx), n_dst, F)
        arg = ops.RelArg(csr, x.to(DEV), edge_w=w.to(DEV),
                         edge_w_idx=None if idx is None else idx.to(DEV))
        out = torch.full((n_dst, F), 5.0, device=DEV)
        ops.aggregate_chunks([(out, arg)], F)
        assert rel_err(out, ref) <= RTOL_F32
        out2 = torch.full((n_dst, F), 5.0, device=DEV)
        ops.aggregate_rows([(out2, [arg], False, bias.to(DEV))], F)
        assert rel_err(out2, ref + bias.double()) <= RTOL_F32
        out3 = torch.ones(n_dst, F, device=DEV)
        ops.aggregate_rows([(out3, [arg, arg], True)], F)
        assert rel_err(out3, 2 * ref + 1) <= RTOL_F32


def test_sum_arrays_with_row_bias():
    a, b = torch.randn(37, 24), torch.randn(37, 24)
    bias = torch.randn(24)
    out = torch.empty(37, 24, device=DEV)
    ops.sum_arrays([(out, [a.to(DEV), b.to(DEV)], bias.to(DEV)), ])
    assert torch.equal(out.cpu(), (a + b) + bias)
    out2 = torch.empty(37, 24, device=DEV)
    ops.sum_arrays([(out2, [a.to(DEV)], bias.to(DEV)), (out, [a.to(DEV), b.to(DEV)])])
    assert torch.equal(out2.cpu(), a + bias) and torch.equal(out.cpu(), a + b)
