"""Composite recipe of BASELINE.json configs[4] in miniature (pytest -m gpu): raw waveform chunk -> SincNet
(non-sequential, [T*B, L]) -> view [T, B, feat] (utils.forward_model's 2-D -> 3-D reshape, utils.py:2336-2337) ->
bidirectional liGRU -> softmax head -> NLLLoss; one backward through all three architectures.

Oracle: the reference's own modules (baseline/_ref/neural_networks.py, CPU fp32) with the same state_dicts and the
same CPU-drawn liGRU dropout masks (SincNet dropout is 0 here: nn.Dropout draws from the device generator)."""
import importlib.util
import os

import numpy as np
import pytest
import torch

import golden_util as gu

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref", "neural_networks.py")


def sinc_opts(use_cuda):
    return {"sinc_N_filt": "16,12", "sinc_len_filt": "33,5", "sinc_max_pool_len": "3,3", "sinc_use_laynorm_inp": "True",
            "sinc_use_batchnorm_inp": "False", "sinc_use_laynorm": "True,True", "sinc_use_batchnorm": "False,False",
            "sinc_act": "relu,relu", "sinc_drop": "0.0,0.0", "sinc_sample_rate": "16000", "sinc_min_low_hz": "50",
            "sinc_min_band_hz": "50", "use_cuda": use_cuda, "to_do": "train"}


def ligru_opts(use_cuda):
    return {"ligru_lay": "32,32", "ligru_drop": "0.2,0.2", "ligru_use_laynorm_inp": "False", "ligru_use_batchnorm_inp": "False",
            "ligru_use_laynorm": "False,False", "ligru_use_batchnorm": "True,True", "ligru_bidir": "True",
            "ligru_act": "relu,relu", "ligru_orthinit": "True", "use_cuda": use_cuda, "to_do": "train"}


def head_opts(S, use_cuda):
    return {"dnn_lay": str(S), "dnn_drop": "0.0", "dnn_use_laynorm_inp": "False", "dnn_use_batchnorm_inp": "False",
            "dnn_use_batchnorm": "False", "dnn_use_laynorm": "False", "dnn_act": "softmax", "use_cuda": use_cuda,
            "to_do": "train"}


def build(lib, use_cuda, L0, S):
    torch.manual_seed(21)
    sn = lib.SincNet(sinc_opts(use_cuda), L0)
    net = lib.liGRU(ligru_opts(use_cuda), sn.out_dim)
    head = lib.MLP(head_opts(S, use_cuda), net.out_dim)
    with torch.no_grad():
        head.wx[0].weight.mul_(20.0)
    return sn, net, head


def forward(mods, x, lab, T, B):
    sn, net, head = mods
    f = sn(x.view(T * B, -1)).view(T, B, -1)        # utils.py:2322-2337
    h = net(f)
    logp = head(h.view(T * B, -1))
    return f, h, logp, torch.nn.functional.nll_loss(logp, lab)


@pytest.mark.skipif(not os.path.exists(REF), reason="baseline/_ref missing (python -c 'import __graft_entry__ as g; g.build()')")
def test_sincnet_into_ligru_matches_reference():
    import neural_networks as pknn
    spec = importlib.util.spec_from_file_location("ref_nn_composite", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    T, B, L0, S = 9, 4, 400, 19
    g = torch.Generator().manual_seed(5)
    x = torch.randn(T, B, L0, generator=g) * (1.0 + torch.rand(T, B, 1, generator=g))
    lab = torch.randint(0, S, (T * B,), generator=g)
    rmods = build(ref, "False", L0, S)
    pmods = build(pknn, "True", L0, S)
    for r, p in zip(rmods, pmods):  # identical constructors -> identical parameters
        for (k, a), (k2, b) in zip(r.state_dict().items(), p.state_dict().items()):
            assert k == k2 and torch.equal(a, b), k
        r.train()
        p.cuda().train()
    torch.manual_seed(33)
    _, h_r, logp_r, loss_r = forward(rmods, x, lab, T, B)
    loss_r.backward()
    torch.manual_seed(33)   # the liGRU masks come from the CPU generator on both sides
    _, h_p, logp_p, loss_p = forward(pmods, x.cuda(), lab.cuda(), T, B)
    loss_p.backward()
    assert gu.relerr(h_p.detach().cpu().numpy(), h_r.detach().numpy()) < 2e-3
    assert gu.relerr(logp_p.detach().cpu().numpy(), logp_r.detach().numpy()) < 1e-3
    assert abs(loss_p.item() - loss_r.item()) / abs(loss_r.item()) < 1e-3
    # gradients reach all three architectures (ReLU + max-pool re-routing: relative L2, tests/test_gpu_parity.py)
    worst = 0.0
    for r, p, name in zip(rmods, pmods, ("sincnet", "ligru", "head")):
        for (k, a), (_, b) in zip(r.named_parameters(), p.named_parameters()):
            if a.grad is None:
                assert b.grad is None or float(b.grad.abs().max()) == 0.0, (name, k)
                continue
            assert b.grad is not None, (name, k)
            ga, gb = a.grad.numpy().astype(np.float64), b.grad.cpu().numpy().astype(np.float64)
            if np.linalg.norm(ga) < 1e-7:
                continue
            l2 = float(np.linalg.norm(ga - gb) / np.linalg.norm(ga))
            worst = max(worst, l2)
            tol = 0.2 if k in ("conv.0.low_hz_", "conv.0.band_hz_") else 0.1   # band edges: cancelling sums
            assert l2 < tol, (name, k, l2)
    print(f"composite SincNet -> liGRU -> head: loss {loss_p.item():.6f} vs {loss_r.item():.6f}, worst gradient rel-L2 {worst:.3e}")
