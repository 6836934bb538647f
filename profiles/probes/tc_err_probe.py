import sys, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
import mmac_b200 as agx
from mmac_b200 import ops
dev='cuda:0'
gen = torch.Generator().manual_seed(0)
M, N = 2048, 128
for K in [32, 128, 256, 512, 1024, 2048, 4096]:
    A = torch.randn(M, K, generator=gen); W = torch.randn(N, K, generator=gen)
    ref = A.double() @ W.double().t()
    C = torch.empty(M, N, device=dev)
    gb = ops.GemmBatch(); gb.add(C, [(A.to(dev), W.to(dev).t())]); gb.run()
    e_tc = float((C.cpu().double()-ref).abs().max())/float(ref.abs().max())
    c32 = (A @ W.t())
    e_32 = float((c32.double()-ref).abs().max())/float(ref.abs().max())
    # positive data (no cancellation): systematic truncation shows up
    A2 = A.abs(); W2 = W.abs(); ref2 = A2.double() @ W2.double().t()
    gb = ops.GemmBatch(); gb.add(C, [(A2.to(dev), W2.to(dev).t())]); gb.run()
    e_tc2 = float((C.cpu().double()-ref2).abs().max())/float(ref2.abs().max())
    e_322 = float(((A2 @ W2.t()).double()-ref2).abs().max())/float(ref2.abs().max())
    print(f'K={K:5d} randn: tc {e_tc:.2e} cpu32 {e_32:.2e} | positive: tc {e_tc2:.2e} cpu32 {e_322:.2e}')
