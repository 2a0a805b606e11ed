"""LSTM_cudnn / RNN_cudnn (reference neural_networks.py:153-297) on the native kernels (pytest -m gpu).

The reference classes are thin wrappers over torch.nn.LSTM / nn.RNN; their arithmetic is torch's, so the oracle is
torch.nn.LSTM / nn.RNN itself on the CPU in fp32 (no cuDNN involved) with the SAME parameters (state_dict copied
through the reference-compatible keys).  Checked: outputs within 1e-3 relative, every parameter gradient
(tanh recurrences: 5e-3 of the tensor's max; ReLU nn.RNN: relative L2, see tests/test_gpu_parity.py on kinks).
"""
import numpy as np
import pytest
import torch

import golden_util as gu

pytestmark = pytest.mark.gpu

OPTS = dict(hidden_size="40", num_layers="2", bias="True", batch_first="True", dropout="0.0", bidirectional="True",
            nonlinearity="tanh", use_cuda="True", to_do="train")


@pytest.mark.parametrize("name,attr,nonlin", [("LSTM_cudnn", "lstm", "tanh"), ("RNN_cudnn", "rnn", "tanh"),
                                               ("RNN_cudnn", "rnn", "relu")])
def test_cudnn_layout_matches_torch_rnn(name, attr, nonlin):
    import neural_networks as pknn
    T, B, D = 23, 5, 11
    torch.manual_seed(3)
    opts = dict(OPTS, nonlinearity=nonlin)
    m = getattr(pknn, name)(opts, D)
    with torch.no_grad():  # the reference zero-initialises the biases: make them matter
        for k, p in m.named_parameters():
            if "bias" in k:
                p.normal_(0, 0.2)
    kw = dict(bias=True, bidirectional=True)
    if name == "RNN_cudnn":
        kw["nonlinearity"] = nonlin
    oracle = (torch.nn.LSTM if name == "LSTM_cudnn" else torch.nn.RNN)(D, 40, 2, **kw)
    oracle.load_state_dict({k[len(attr) + 3:]: v for k, v in m.state_dict().items()})
    g = torch.Generator().manual_seed(4)
    x = torch.randn(T, B, D, generator=g)
    wgt = torch.randn(T, B, 80, generator=g)
    y_ref, _ = oracle(x)
    (y_ref * wgt).sum().backward()
    m.cuda().train()
    xg = x.cuda().requires_grad_(True)
    y = m(xg)
    (y * wgt.cuda()).sum().backward()
    assert gu.relerr(y.detach().cpu().numpy(), y_ref.detach().numpy()) < 1e-3
    for (k, p), (kr, pr) in zip(m.named_parameters(), oracle.named_parameters()):
        assert k.endswith(kr)
        got, ref = p.grad.cpu().numpy(), pr.grad.numpy()
        if nonlin == "relu":
            l2 = float(np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-30))
            assert l2 < 0.05, (k, l2)
        else:
            assert gu.relerr(got, ref, floor=1e-6) < 5e-3, k
    # the gradient w.r.t. the input flows too (a module in front of the RNN trains through it)
    xr = x.clone().requires_grad_(True)
    (oracle(xr)[0] * wgt).sum().backward()
    if nonlin == "relu":  # kinks: relative L2 (tests/test_gpu_parity.py TOL_GRAD_KINK_L2)
        got, ref = xg.grad.cpu().numpy(), xr.grad.numpy()
        assert float(np.linalg.norm(got - ref) / np.linalg.norm(ref)) < 0.05
    else:
        assert gu.relerr(xg.grad.cpu().numpy(), xr.grad.numpy()) < 5e-3
