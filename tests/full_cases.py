"""Full-size parity cases (BASELINE.json configs 2, 3, 4): ONE construction recipe shared by

  * tests/golden/make_golden_full.py — runs it against the UNMODIFIED reference classes in the build container and
    commits reduced outputs (loss, err, sampled log-posterior rows, arg-max + margins, hidden-state samples,
    sampled gradients, parameters after one RMSprop step) as tests/golden/full_*.npz;
  * tests/test_gpu_full_parity.py and bench.py's step-0 self check — run the SAME recipe against the drop-in
    classes on the GPU.

Nothing big is committed: weights come from the seeded constructors (constructor RNG parity is pinned by
tests/test_dropin_cpu.py against structure.json; the fixture additionally stores per-tensor checksums), inputs
from a seeded generator (checksummed), dropout masks from the CPU generator exactly as the reference draws them
(neural_networks.py:1103-1105; stored bit-packed as a cross-check).
"""
import numpy as np
import torch

CELLS = {"ligru": ("liGRU", ("wh", "wz")), "lstm": ("LSTM", ("wfx", "wix", "wox", "wcx")),
         "gru": ("GRU", ("wh", "wz", "wr")), "minimalgru": ("minimalGRU", ("wh", "wz")), "rnn": ("RNN", ("wh",))}

# name -> recipe.  `backward`: also run NLLLoss backward + one RMSprop(4e-4, .95, 1e-8) step (utils.py:2121-2131)
CASES = {
    # config 2 — the headline: TIMIT liGRU 5x550 bidir, BN, ReLU, dropout .2, 1936 senones, 500x32x40
    "full_ligru5x550": dict(cell="ligru", T=500, B=32, D=40, lay=[550] * 5, S=1936, drop=0.2, act="relu", seed=2001,
                            backward=True),
    # config 3 — TIMIT LSTM 4x550 bidir, BN, tanh (cfg/TIMIT_baselines/TIMIT_LSTM_fbank.cfg)
    "full_lstm4x550": dict(cell="lstm", T=500, B=32, D=40, lay=[550] * 4, S=1936, drop=0.2, act="tanh", seed=2002,
                           backward=True),
    # config 4 — Librispeech stress shape liGRU 5x1024 / 3440 senones: forward only
    "full_ligru5x1024": dict(cell="ligru", T=500, B=32, D=40, lay=[1024] * 5, S=3440, drop=0.2, act="relu", seed=2003,
                             backward=False),
}

ROW_STRIDE = 97  # every 97th frame's full log-posterior row is stored


def rec_opts(cell, lay, drop, act, to_do="train", use_cuda="False"):
    n = len(lay)
    o = {
        "_lay": ",".join(map(str, lay)), "_drop": ",".join([str(drop)] * n), "_use_laynorm_inp": "False",
        "_use_batchnorm_inp": "False", "_use_laynorm": ",".join(["False"] * n),
        "_use_batchnorm": ",".join(["True"] * n), "_bidir": "True", "_act": ",".join([act] * n), "_orthinit": "True",
    }
    o = {cell + k: v for k, v in o.items()}
    o.update(use_cuda=use_cuda, to_do=to_do)
    return o


def head_opts(S, to_do="train", use_cuda="False"):
    return {"dnn_lay": str(S), "dnn_drop": "0.0", "dnn_use_laynorm_inp": "False", "dnn_use_batchnorm_inp": "False",
            "dnn_use_batchnorm": "False", "dnn_use_laynorm": "False", "dnn_act": "softmax", "use_cuda": use_cuda,
            "to_do": to_do}


def build(nn_lib, case, use_cuda="False"):
    """Seeded constructors + the deterministic margin tweak.  Returns (net, head); everything on the CPU."""
    c = CASES[case]
    torch.manual_seed(c["seed"])
    cls_name, gates = CELLS[c["cell"]]
    net = getattr(nn_lib, cls_name)(rec_opts(c["cell"], c["lay"], c["drop"], c["act"], use_cuda=use_cuda), c["D"])
    head = nn_lib.MLP(head_opts(c["S"], use_cuda=use_cuda), net.out_dim)
    with torch.no_grad():  # untrained heads give near-uniform posteriors: give them real margins
        head.wx[0].weight.mul_(30.0)
        head.wx[0].bias.normal_(0, 0.1)
        for i in range(len(c["lay"])):
            for g in gates:
                getattr(net, "bn_" + g)[i].weight.uniform_(0.5, 1.5)
                getattr(net, "bn_" + g)[i].bias.normal_(0, 0.2)
    return net, head


def inputs(case):
    c = CASES[case]
    g = torch.Generator().manual_seed(c["seed"] + 1)
    x = torch.randn(c["T"], c["B"], c["D"], generator=g)
    lab = torch.randint(0, c["S"], (c["T"] * c["B"],), generator=g)
    return x, lab


def forward_seed(case):
    """torch.manual_seed value set right before the training-mode forward (dropout masks, :1103-1105)."""
    return CASES[case]["seed"] + 2


def checksum(t):
    a = np.asarray(t.detach().cpu().numpy() if torch.is_tensor(t) else t, dtype=np.float64)
    return np.array([a.sum(), (a * a).sum()])


def sample_idx(n, k=16384, seed=0, keep=None):
    """Deterministic sample positions (never stored): tensors up to k elements are kept whole, larger ones are
    sampled at the first `keep` of k seeded draws."""
    if n <= k:
        return np.arange(n)
    idx = np.random.default_rng(seed).integers(0, n, k)
    return idx if keep is None else idx[:keep]


GRAD_KEEP = 8192


def state_pairs(net, head):
    """(name, tensor) over both modules' parameters in registration order."""
    for pfx, m in (("net.", net), ("head.", head)):
        for k, p in m.named_parameters():
            yield pfx + k, p
