"""Shared test helpers (deterministic weights, error metrics, oracle/product model builders)."""
from __future__ import annotations

import math
import os
import sys
from collections import OrderedDict

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')

# north_star tolerances: fp32 paths agree to rel 1e-5, bf16 to rel 2e-2
RTOL_F32 = 1e-5
RTOL_BF16 = 2e-2


def parity_log(kind: str, what, value: float, **extra):
    """Every error a parity test measures is appended to the file ``AGX_PARITY_LOG`` names (one JSON
    object per line: test id, what was compared, the error) -- and, for comparisons that needed the
    float64 fallback of test_gpu_model._compare_grads / _assert_close, always to
    ``gpurun_out/parity_fallbacks.jsonl`` -- so that the tolerance actually reached is on record,
    not only pass / fail."""
    import json
    rec = {'test': os.environ.get('PYTEST_CURRENT_TEST', '').split(' ')[0], 'kind': kind,
           'what': str(what), 'err': float(value), **extra}
    paths = []
    if os.environ.get('AGX_PARITY_LOG'):
        paths.append(os.environ['AGX_PARITY_LOG'])
    if kind == 'fallback':
        paths.append(os.path.join(ROOT, 'gpurun_out', 'parity_fallbacks.jsonl'))
        print(f"[parity fallback] {rec['what']}: rel err {value:.3e} {extra}")
    for path in paths:
        try:
            os.makedirs(os.path.dirname(path) or '.', exist_ok=True)
            with open(path, 'a') as fh:
                fh.write(json.dumps(rec) + '\n')
        except OSError:
            pass


def rel_err(a, b) -> float:
    """max |a - b| / max |b| over the WHOLE tensor (scale-relative: the error of every element
    against the largest reference entry -- the metric every 1e-5 / 2e-2 bound in tests/ is stated
    in; it is not an element-wise relative error)."""
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    if b.numel() == 0:
        return 0.0
    scale = max(float(b.abs().max()), 1e-30)
    e = float((a - b).abs().max()) / scale
    if os.environ.get('AGX_PARITY_LOG'):
        import inspect
        fr = inspect.stack()[1]
        parity_log('rel_err', f'{os.path.basename(fr.filename)}:{fr.lineno}', e)
    return e


def fill_params_deterministic(model: torch.nn.Module, only=None):
    """Closed-form weights so fixtures need not store them: p[i] = s * sin(0.37*i + 1.3*k) where
    k = crc32(parameter name) % 1000 and s = 1/sqrt(fan_in) (1-D: 1 +- 0.1 for '.weight' of norms,
    +-0.1 otherwise).  Running stats of BatchNorm are left untouched."""
    import zlib
    names = sorted(n for n, p in model.named_parameters()
                   if not isinstance(p, torch.nn.parameter.UninitializedParameter))
    with torch.no_grad():
        params = dict(model.named_parameters())
        for name in names:
            k = zlib.crc32(name.encode()) % 1000
            if only is not None and not any(name.startswith(o) or ('.' + o + '.') in name or
                                            name.startswith(o + '.') for o in only):
                continue
            p = params[name]
            idx = torch.arange(p.numel(), dtype=torch.float64)
            base = torch.sin(0.37 * idx + 1.3 * k)
            if p.dim() >= 2:
                v = base / math.sqrt(p.shape[1])
            elif name.endswith('.weight'):          # BatchNorm gamma
                v = 1.0 + 0.1 * base
            else:
                v = 0.1 * base
            p.copy_(v.view(p.shape).to(p.dtype))


def copy_state(src: torch.nn.Module, dst: torch.nn.Module):
    """Load src's state into dst (lazy weights of either side handled by the Linear modules)."""
    sd = OrderedDict()
    for k, v in src.state_dict().items():
        sd[k] = v if isinstance(v, torch.nn.parameter.UninitializedParameter) else v.detach().cpu()
    missing, unexpected = dst.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    return missing


def load_golden(name: str):
    return dict(np.load(os.path.join(GOLDEN, name)))


def undirected_graph(size: str, features: str = 'one-hot', seed=None):
    """(graph, undirected edge_index dict, metadata) on CPU via the oracle's ToUndirected."""
    import mmac_b200  # noqa: F401
    from mmac_b200 import synth
    from oracle import graph_oracle as go
    g = synth.make_artgraph(size, features=features, seed=seed)
    ei = go.to_undirected(g.edge_index_dict)
    return g, ei, (g.node_types, list(ei.keys()))


def reset_bn(model: torch.nn.Module):
    """Forget the running statistics gathered by the lazy-initialisation forward (which runs with
    the random initial weights), so fixtures depend on the deterministic weights only."""
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            m.reset_running_stats()
