"""FusionRNN family (pytest -m gpu): FusionLinearConv, liGRU_layer and fusionRNN_jit (reference neural_networks.py
:2057-2099, :795-995, :719-793; cfg/DIRHA_baselines/DIRHA_fusionRNN_MFCC_6ch.cfg) against the reference's OWN classes
(baseline/_ref/neural_networks.py, git-ignored copy made by __graft_entry__.build()).

FusionLinearConv is a plain nn.Module: its oracle runs on the CPU in fp32.  liGRU_layer / fusionRNN_jit hard-code
device="cuda" inside TorchScript methods, so their oracle is the unmodified reference on the same GPU with stock PyTorch
(TF32 off) — none of this repository's kernels on that side."""
import importlib.util
import os

import numpy as np
import pytest
import torch

import golden_util as gu

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref", "neural_networks.py")
needs_ref = pytest.mark.skipif(not os.path.exists(REF), reason="baseline/_ref missing (python -c 'import __graft_entry__ as g; g.build()')")


def _ref():
    spec = importlib.util.spec_from_file_location("ref_nn_fusion", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    return ref


def _l2(a, b):
    a, b = a.detach().double().cpu().numpy(), b.detach().double().cpu().numpy()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(a), 1e-30))


@needs_ref
@pytest.mark.parametrize("act,reduce", [("prelu", "sum"), ("relu", "mean"), ("leaky_relu", "sum"), ("tanh", "mean")])
def test_fusion_linear_conv_matches_reference(act, reduce):
    import neural_networks as pknn
    ref = _ref()
    T, B, M, d, H = 7, 5, 3, 24, 40
    torch.manual_seed(3)
    r = ref.FusionLinearConv(M * d, H, number_of_mic=M, act=act, reduce=reduce)
    torch.manual_seed(3)
    p = pknn.FusionLinearConv(M * d, H, number_of_mic=M, act=act, reduce=reduce)
    for (k, a), (k2, b) in zip(r.state_dict().items(), p.state_dict().items()):
        assert k == k2 and torch.equal(a, b), k
    with torch.no_grad():   # a bias and a slope that matter
        r.conv.bias.normal_(0, 0.3)
        p.conv.bias.copy_(r.conv.bias)
    p.cuda()
    g = torch.Generator().manual_seed(11)
    x = torch.randn(T, B, M * d, generator=g)
    w = torch.randn(T, B, H, generator=g)
    xr = x.clone().requires_grad_(True)
    yr = r(xr)
    (yr * w).sum().backward()
    xp = x.cuda().requires_grad_(True)
    yp = p(xp)
    (yp * w.cuda()).sum().backward()
    assert yp.shape == yr.shape
    assert gu.relerr(yp.detach().cpu().numpy(), yr.detach().numpy()) < 2e-3
    assert _l2(xr.grad, xp.grad) < 5e-3, ("dx", _l2(xr.grad, xp.grad))
    for (k, a), (_, b) in zip(r.named_parameters(), p.named_parameters()):
        assert b.grad is not None, k
        assert _l2(a.grad, b.grad) < 5e-3, (k, _l2(a.grad, b.grad))


def _opts(to_do, drop="0.0,0.0,0.0"):
    return {"fusionRNN_lay": "48,48,48", "fusionRNN_drop": drop, "batches": "6", "fusionRNN_do_fusion": "True",
            "fusionRNN_fusion_act": "prelu", "fusionRNN_fusion_reduce": "sum", "fusionRNN_fusion_layer_size": "144",
            "fusionRNN_number_of_mic": "3", "fusionRNN_bidir": "True", "fusionRNN_act": "prelu,prelu,prelu",
            "use_cuda": "True", "to_do": to_do}


@needs_ref
def test_fusion_rnn_matches_the_reference_on_the_same_gpu():
    """Unmodified reference fusionRNN_jit (TorchScript, stock PyTorch kernels) vs the drop-in, same state_dict: training
    step (dropout 0 -> all-ones masks on both sides) with every gradient, BatchNorm running statistics, then eval."""
    import neural_networks as pknn
    ref = _ref()
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    T, B, M, d = 21, 6, 3, 13
    D = M * d
    torch.manual_seed(17)
    r = ref.fusionRNN_jit(_opts("train"), D)
    torch.manual_seed(17)
    p = pknn.fusionRNN_jit(_opts("train"), D)
    assert r.out_dim == p.out_dim == 96
    for (k, a), (k2, b) in zip(r.state_dict().items(), p.state_dict().items()):
        # same keys / shapes; the VALUES cannot match: the reference moves `u` to the GPU before orthogonal_ (:866-871),
        # i.e. it draws from the CUDA generator, so the state_dict is copied over instead
        assert k == k2 and a.shape == b.shape, (k, k2)
    p.load_state_dict({k: v.cpu() for k, v in r.state_dict().items()})
    with torch.no_grad():   # biases in front of BatchNorm that are not zero (they must cancel / fold correctly)
        for lr_, lp_ in zip(list(r.model)[1:], list(p.model)[1:]):
            for nm in ("wz", "wh"):
                getattr(lr_, nm).bias.normal_(0, 0.2)
                getattr(lp_, nm).bias.copy_(getattr(lr_, nm).bias.cpu())
    r.cuda().train()
    p.cuda().train()
    g = torch.Generator().manual_seed(5)
    x = torch.randn(T, B, D, generator=g).cuda()
    w = torch.randn(T, B, 96, generator=g).cuda()
    yr = r(x)
    (yr * w).sum().backward()
    yp = p(x)
    (yp * w).sum().backward()
    assert yp.shape == yr.shape == (T, B, 96)
    assert gu.relerr(yp.detach().cpu().numpy(), yr.detach().cpu().numpy()) < 2e-3
    worst = 0.0
    for (k, a), (k2, b) in zip(r.named_parameters(), p.named_parameters()):
        assert k == k2
        if k.endswith(("wz.bias", "wh.bias")) and ".0." not in k:
            continue   # Linear bias in front of BatchNorm: mathematically zero gradient (reference: rounding noise)
        assert b.grad is not None, k
        l2 = _l2(a.grad, b.grad)
        worst = max(worst, l2)
        assert l2 < 0.1, (k, l2)   # ReLU recurrences: relative L2 (tests/test_gpu_parity.py)
    for (k, a), (_, b) in zip(r.state_dict().items(), p.state_dict().items()):
        if "running_" in k:
            assert torch.allclose(a.float().cpu(), b.float().cpu(), rtol=2e-3, atol=2e-4), k
        if "num_batches_tracked" in k:
            assert int(a) == int(b) == 1, k
    # eval: running statistics, biases folded into the shift, mask = 1
    r.eval()
    p.eval()
    with torch.no_grad():
        p.load_state_dict({k: v.cpu() for k, v in r.state_dict().items()})
        er, ep = r(x), p(x)
    assert gu.relerr(ep.cpu().numpy(), er.cpu().numpy()) < 2e-3
    print(f"fusionRNN_jit vs reference on the GPU: worst gradient rel-L2 {worst:.3e}")


def test_fusion_rnn_dropout_mask_is_inverted_dropout():
    """Training masks are nn.Dropout masks (kept entries scaled by 1/(1-p), :935-945), constant over time: with a forced
    mask the output must equal the run with that mask, and a fresh draw only holds the values {0, 1/(1-p)}."""
    import neural_networks as pknn
    torch.manual_seed(1)
    lay = pknn.liGRU_layer(20, 32, 1, 4, dropout=0.25, bidirectional=True).cuda().train()
    m, s = lay._mask(8, torch.device("cuda"))
    vals = torch.unique(m).cpu().tolist()
    assert s == 1.0 and all(abs(v) < 1e-6 or abs(v - 1.0 / 0.75) < 1e-5 for v in vals), vals
    x = torch.randn(9, 4, 20, device="cuda")
    lay._mask_override = m
    y1 = lay(x)
    y2 = lay(x)
    assert torch.allclose(y1, y2, rtol=1e-6, atol=1e-7) and y1.shape == (9, 4, 64)
    lay.eval()
    assert lay._mask(8, torch.device("cuda")) == (None, 1.0)
