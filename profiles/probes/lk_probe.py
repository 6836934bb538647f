import sys, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
import mmac_b200 as agx
from mmac_b200 import ops
dev = 'cuda:0'
M = N = 128
for K, kind in ((2048, 'ones'), (2048, 'rowid'), (2048, 'colid'), (4096, 'rand')):
    if kind == 'ones':
        dout = torch.ones(K, M); x = torch.ones(K, N)
    elif kind == 'rowid':     # dout[k, m] = m ; x = 1  -> dW[m, n] = K*m
        dout = torch.arange(M).float().repeat(K, 1); x = torch.ones(K, N)
    elif kind == 'colid':     # x[k, n] = n -> dW[m,n] = K*n
        dout = torch.ones(K, M); x = torch.arange(N).float().repeat(K, 1)
    else:
        g = torch.Generator().manual_seed(1); dout = torch.randn(K, M, generator=g); x = torch.randn(K, N, generator=g)
    dd, xd = dout.to(dev), x.to(dev)
    dw = torch.full((M, N), float('nan'), device=dev)
    gb = ops.GemmBatch()
    sk = ops.split_k_for(K)
    gb.add(dw, [(dd.t(), xd)], split_k=sk)
    part = gb._keep[0]
    part.fill_(float('nan'))
    gb.run()
    torch.cuda.synchronize()
    ref = dout.double().t() @ x.double()
    p = part.view(sk, M, N)
    print(kind, K, 'split', sk, 'dw[0,:4]', dw[0, :4].tolist(), 'dw[5,:4]', dw[5, :4].tolist(), 'ref[5,:4]', ref[5, :4].tolist())
    print('   partial slab0 [0,:4]', p[0, 0, :4].tolist(), 'nan slabs', int(torch.isnan(p[:, 0, 0]).sum()), 'of', sk,
          'err', float((dw.double().cpu() - ref).abs().max() / ref.abs().max()))
