"""QLSTM (reference quaternion_neural_networks.py:21-172, cfg/DIRHA_baselines/DIRHA_QLSTM_MFCC.cfg) on the native LSTM
kernels, against the reference's own QLSTM (baseline/_ref, CPU fp32, both linear-layer variants), pytest -m gpu."""
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

import golden_util as gu

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref", "quaternion_neural_networks.py")


def _opts(autograd, use_cuda):
    return {"lstm_lay": "24,24", "lstm_drop": "0.2,0.2", "lstm_bidir": "True", "lstm_act": "tanh,tanh",
            "quaternion_init": "quaternion", "autograd": autograd, "use_cuda": use_cuda, "to_do": "train"}


@pytest.mark.skipif(not os.path.exists(REF), reason="baseline/_ref missing (python -c 'import __graft_entry__ as g; g.build()')")
@pytest.mark.parametrize("autograd", ["True", "False"])
def test_qlstm_matches_reference(autograd):
    import quaternion_neural_networks as pkq
    spec = importlib.util.spec_from_file_location("ref_qnn", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    T, B, D = 11, 4, 20
    np.random.seed(4)
    torch.manual_seed(4)
    r = ref.QLSTM(_opts(autograd, "False"), D)
    np.random.seed(4)
    torch.manual_seed(4)
    p = pkq.QLSTM(_opts(autograd, "True"), D)
    assert r.out_dim == p.out_dim == 48
    for (k, a), (k2, b) in zip(r.state_dict().items(), p.state_dict().items()):
        assert k == k2 and torch.equal(a, b), k          # same draws from numpy / scipy, same parameter names
    with torch.no_grad():
        for m_r, m_p in zip(r.wfx, p.wfx):                 # biases that matter
            m_r.bias.normal_(0, 0.3)
            m_p.bias.copy_(m_r.bias)
    p.cuda()
    g = torch.Generator().manual_seed(9)
    x = torch.randn(T, B, D, generator=g)
    w = torch.randn(T, B, 48, generator=g)
    torch.manual_seed(33)
    yr = r(x)
    (yr * w).sum().backward()
    torch.manual_seed(33)            # the dropout masks come from the CPU generator on both sides
    yp = p(x.cuda())
    (yp * w.cuda()).sum().backward()
    assert gu.relerr(yp.detach().cpu().numpy(), yr.detach().numpy()) < 2e-3
    worst = 0.0
    for (k, a), (_, b) in zip(r.named_parameters(), p.named_parameters()):
        assert b.grad is not None, k
        ga, gb = a.grad.double().numpy(), b.grad.double().cpu().numpy()
        l2 = float(np.linalg.norm(ga - gb) / max(np.linalg.norm(ga), 1e-30))
        worst = max(worst, l2)
        assert l2 < 2e-2, (k, l2)
    print(f"QLSTM (autograd={autograd}) vs reference: worst gradient rel-L2 {worst:.3e}")
