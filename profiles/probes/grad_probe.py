import sys, copy, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import util
import test_gpu_model as T
from oracle import graph_oracle as go
import mmac_b200 as agx
g, ei, orc, prod = T._build_pair('SAGEConv', 32, 'small', features='dense', dropout=0.4)
gen = torch.Generator().manual_seed(77)
masks = {t: (torch.rand(n, 128, generator=gen) >= 0.4).float() / 0.6 for t, n in g.num_nodes_dict.items()}
orc.gnn.dropout_masks = masks
prod.gnn.dropout_masks = {t: m.to('cuda') for t, m in masks.items()}
orc.train(); prod.train()
y = g['artwork'].y_style
e_o, o_o = orc(g.x_dict, ei); l_o = go.nll_loss_artwork(o_o[0], y); l_o.backward()
e_p, o_p = prod(T._to_dev(g.x_dict), T._to_dev(ei)); l_p = agx.functional.nll_loss(o_p[0]['artwork'], y.to('cuda')); l_p.backward()
g64 = T._oracle_grads_fp64(orc, g, ei, y, masks)
print('--- embeddings (h2) err vs fp64: ref32 / prod')
for t in e_o:
    r64 = g64['__emb__'][t]; s = float(r64.abs().max())
    print(t, f"{float((e_o[t].detach().double()-r64).abs().max())/s:.2e} {float((e_p[t].detach().cpu().double()-r64).abs().max())/s:.2e}")
og = {n: p.grad for n, p in orc.named_parameters() if p.grad is not None}
pg = {n: p.grad for n, p in prod.named_parameters() if p.grad is not None}
rows = []
for n, gref in og.items():
    if n not in pg: continue
    r64 = g64[n]; s = float(r64.abs().max()) + 1e-30
    re = float((gref.double()-r64).abs().max())/s; pe = float((pg[n].cpu().double()-r64).abs().max())/s
    rows.append((pe/max(re,1e-12), n, re, pe, s))
rows.sort(reverse=True)
for ratio, n, re, pe, s in rows[:25]:
    print(f'{ratio:8.1f}x ref {re:.2e} prod {pe:.2e} scale {s:.2e} {n}')
