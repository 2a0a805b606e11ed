"""Drop-in `quaternion_neural_networks` module (cfg/DIRHA_baselines/DIRHA_QLSTM_MFCC.cfg: `arch_library =
quaternion_neural_networks`, `arch_class = QLSTM`), B200-native compute.

A quaternion linear layer is a real linear layer whose weight is assembled from four component matrices by the Hamilton
product pattern (reference quaternion_neural_networks.py:375-395); the QLSTM recurrence itself (:143-156) is the LSTM
algebra of neural_networks.LSTM with biases instead of BatchNorm.  So the classes keep the reference's constructors
(parameter names, shapes, registration order, numpy / scipy-based initialisation with the same draws) and the forward runs
this library's LSTM kernels (cluster-persistent for H <= 560) on the assembled matrices: assembling them is a handful of
`torch.cat` on parameters (autograd routes the dense weight gradients of the kernels back to r / i / j / k), everything
that touches activations is libpk_b200.so.  There is no eager-PyTorch or CPU path for the forward."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
from numpy.random import RandomState
from torch.nn.parameter import Parameter

import pk_functions as pkf
import pk_native as pk
from neural_networks import _floats, _ints, _require_cuda, flip, strtobool  # noqa: F401  (flip: reference API)


def act_fun(act_type):
    """Reference quaternion_neural_networks.py:280-308 (module objects kept for API / state parity)."""
    table = {"relu": nn.ReLU, "prelu": nn.PReLU, "tanh": nn.Tanh, "sigmoid": nn.Sigmoid, "hardtanh": nn.Hardtanh,
             "leaky_relu": lambda: nn.LeakyReLU(0.2), "elu": nn.ELU, "softmax": lambda: nn.LogSoftmax(dim=1),
             "linear": lambda: nn.LeakyReLU(1)}
    return table[act_type]() if act_type in table else None


# ---------------------------------------------------------------------------------------------
# initialisation (reference :486-647): same draws from the same generators in the same order
# ---------------------------------------------------------------------------------------------


def _scale(fan_in, fan_out, criterion):
    if criterion == "glorot":
        return 1.0 / np.sqrt(2 * (fan_in + fan_out))
    if criterion == "he":
        return 1.0 / np.sqrt(2 * fan_in)
    raise ValueError("Invalid criterion: " + criterion)


def _shape_and_scale(in_features, out_features, kernel_size, criterion):
    if kernel_size is None:
        return (in_features, out_features), _scale(in_features, out_features, criterion)
    rf = int(np.prod(kernel_size))
    ks = (kernel_size,) if isinstance(kernel_size, int) else tuple(kernel_size)
    return (out_features, in_features) + ks, _scale(in_features * rf, out_features * rf, criterion)


def quaternion_init(in_features, out_features, rng, kernel_size=None, criterion="glorot"):
    """:578-627 — chi(4)-distributed modulus, a random unit purely-imaginary axis, a uniform phase."""
    from scipy.stats import chi
    shape, s = _shape_and_scale(in_features, out_features, kernel_size, criterion)
    rng = RandomState(np.random.randint(1, 1234))      # the argument is replaced, as in the reference (:594)
    modulus = chi.rvs(4, loc=0, scale=s, size=shape)
    n = int(np.prod(shape))
    v = [np.random.normal(0, 1.0, n) for _ in range(3)]
    norm = np.sqrt(v[0] ** 2 + v[1] ** 2 + v[2] ** 2 + 0.0001)
    v = [(c / norm).reshape(shape) for c in v]
    phase = rng.uniform(low=-np.pi, high=np.pi, size=shape)
    return (modulus * np.cos(phase),) + tuple(modulus * c * np.sin(phase) for c in v)


def unitary_init(in_features, out_features, rng, kernel_size=None, criterion="he"):
    """:490-537 — uniform components normalised to unit quaternions."""
    shape, s = _shape_and_scale(in_features, out_features, kernel_size, criterion)
    s = np.sqrt(3.0) * s
    n = int(np.prod(shape))
    v = [np.random.uniform(-s, s, n) for _ in range(4)]
    norm = np.sqrt(v[0] ** 2 + v[1] ** 2 + v[2] ** 2 + v[3] ** 2) + 0.0001
    return tuple((c / norm).reshape(shape) for c in v)


def random_init(in_features, out_features, rng, kernel_size=None, criterion="glorot"):
    """:539-575 — uniform [0, 1) components times the fan-in/out scale."""
    shape, s = _shape_and_scale(in_features, out_features, kernel_size, criterion)
    n = int(np.prod(shape))
    return tuple(np.random.uniform(0.0, 1.0, n).reshape(shape) * s for _ in range(4))


def affect_init(r_weight, i_weight, j_weight, k_weight, init_func, rng, init_criterion):
    """:629-647"""
    if not (r_weight.size() == i_weight.size() == j_weight.size() == k_weight.size()):
        raise ValueError("The real and imaginary weights should have the same size")
    if r_weight.dim() != 2:
        raise Exception("affect_init accepts only matrices. Found dimension = " + str(r_weight.dim()))
    comps = init_func(r_weight.size(0), r_weight.size(1), rng, None, init_criterion)
    for w, c in zip((r_weight, i_weight, j_weight, k_weight), comps):
        w.data = torch.from_numpy(c).type_as(w.data)


def hamilton_matrix(r, i, j, k):
    """The real [4*in, 4*out] matrix K with x @ K = W (*) x (Hamilton product), :375-385."""
    return torch.cat([torch.cat([r, -i, -j, -k], dim=0), torch.cat([i, r, -k, j], dim=0),
                      torch.cat([j, k, r, -i], dim=0), torch.cat([k, -j, i, r], dim=0)], dim=1)


# ---------------------------------------------------------------------------------------------
# layers
# ---------------------------------------------------------------------------------------------


class _QuaternionLinearBase(nn.Module):
    _INITS = {"quaternion": quaternion_init, "unitary": unitary_init, "random": random_init}

    def __init__(self, in_features, out_features, bias=True, init_criterion="glorot", weight_init="quaternion", seed=None):
        super().__init__()
        self.in_features = in_features // 4
        self.out_features = out_features // 4
        for name in ("r_weight", "i_weight", "j_weight", "k_weight"):
            setattr(self, name, Parameter(torch.Tensor(self.in_features, self.out_features)))
        self._make_bias(bias)
        self.init_criterion = init_criterion
        self.weight_init = weight_init
        self.seed = seed if seed is not None else np.random.randint(0, 1234)
        self.rng = RandomState(self.seed)
        self.reset_parameters()

    def reset_parameters(self):
        if self.bias is not None:
            self.bias.data.fill_(0)
        affect_init(self.r_weight, self.i_weight, self.j_weight, self.k_weight, self._INITS[self.weight_init], self.rng,
                    self.init_criterion)

    def dense_weight(self):
        """nn.Linear-style [out, in] matrix of this layer (what the kernels consume)."""
        return hamilton_matrix(self.r_weight, self.i_weight, self.j_weight, self.k_weight).t().contiguous()

    def dense_bias(self, device):
        b = self.bias
        if b is None:
            return torch.zeros(self.out_features * 4, device=device)
        return b if b.device == device else b.to(device)

    def forward(self, input):
        """x @ K + b through the native dense layer (tcgen05 GEMM), 2-D or 3-D input."""
        _require_cuda(input, type(self).__name__)
        if input.dim() not in (2, 3):
            raise RuntimeError("quaternion linear accepts only input of dimension 2 or 3. input.dim = " + str(input.dim()))
        if input.size(-1) % 4 != 0:
            raise RuntimeError("Quaternion Tensors must be divisible by 4. input.size()[1] = " + str(input.size(-1)))
        x2 = input.reshape(-1, input.size(-1))
        O = self.out_features * 4
        cfg = pkf.DenseLayerCfg(O=O, act="linear", use_bn=False, bn_training=False, grad_enabled=torch.is_grad_enabled())
        with torch.cuda.device(input.device):
            y = pkf.MLPStackFn.apply(x2, [cfg], self.dense_weight(), self.dense_bias(input.device))
        return y.view(*input.shape[:-1], O)

    def __repr__(self):
        return (f"{type(self).__name__}(in_features={self.in_features}, out_features={self.out_features}, "
                f"bias={self.bias is not None}, init_criterion={self.init_criterion}, weight_init={self.weight_init}, "
                f"seed={self.seed})")


class QuaternionLinearAutograd(_QuaternionLinearBase):
    """Reference :175-220.  Without bias the reference keeps a plain zero tensor attribute (not a parameter)."""

    def _make_bias(self, bias):
        if bias:
            self.bias = Parameter(torch.Tensor(self.out_features * 4))
        else:
            self.bias = torch.zeros(self.out_features * 4)


class QuaternionLinear(_QuaternionLinearBase):
    """Reference :222-276 (custom-backward variant there; the same layer here)."""
    _INITS = {"quaternion": quaternion_init, "unitary": unitary_init}

    def _make_bias(self, bias):
        if bias:
            self.bias = Parameter(torch.Tensor(self.out_features * 4))
        else:
            self.register_parameter("bias", None)


class QLSTM(nn.Module):
    """Reference :21-172 ("Quaternion Recurrent Neural Networks", Parcollet et al., ICLR 2019): per layer four input
    and four recurrent quaternion linear layers (gates f, i, o, c), biases on the input side, one Bernoulli(1 - p) mask
    per layer applied to the candidate, shared weights for the two directions.  Constructor as in the reference; the
    forward hands the assembled dense matrices to this library's LSTM kernels (one call for the whole stack)."""

    _GATES = (("wfx", "ufh"), ("wix", "uih"), ("wox", "uoh"), ("wcx", "uch"))

    def __init__(self, options, inp_dim):
        super().__init__()
        self.input_dim = inp_dim
        self.lstm_lay = _ints(options["lstm_lay"])
        self.lstm_drop = _floats(options["lstm_drop"])
        self.lstm_act = str(options["lstm_act"]).split(",")
        self.bidir = strtobool(options["lstm_bidir"])
        self.use_cuda = strtobool(options["use_cuda"])
        self.autograd = strtobool(options["autograd"])
        self.to_do = options["to_do"]
        self.test_flag = self.to_do != "train"
        for w, u in self._GATES:
            setattr(self, w, nn.ModuleList([]))
            setattr(self, u, nn.ModuleList([]))
        self.act = nn.ModuleList([])
        self.N_lstm_lay = len(self.lstm_lay)
        lin = QuaternionLinearAutograd if self.autograd else QuaternionLinear
        cur = self.input_dim
        for i, H in enumerate(self.lstm_lay):
            self.act.append(act_fun(self.lstm_act[i]))
            for w, _ in self._GATES:
                getattr(self, w).append(lin(cur, H, bias=True))
            for _, u in self._GATES:
                getattr(self, u).append(lin(H, H, bias=False))
            cur = 2 * H if self.bidir else H
        self.out_dim = self.lstm_lay[-1] + self.bidir * self.lstm_lay[-1]
        self.fast_dropout = False
        self.cell_flags = 0

    def _mask(self, i, rows, H, device):
        if self.test_flag:
            return None, 1.0 - self.lstm_drop[i]
        if self.fast_dropout:
            return torch.empty(rows, H, device=device).bernoulli_(1.0 - self.lstm_drop[i]), 1.0
        m = torch.bernoulli(torch.Tensor(rows, H).fill_(1 - self.lstm_drop[i]))   # CPU generator, like the reference (:133)
        return m.to(device, non_blocking=True), 1.0

    def forward(self, x):
        _require_cuda(x, "QLSTM")
        for a in self.lstm_act:
            if a not in pk.ACT_IDS:
                raise NotImplementedError(f"pytorch-kaldi_b200.QLSTM: activation {a!r} is not implemented inside a recurrent layer")
        T, B, _ = x.shape
        rows = (2 if self.bidir else 1) * B
        cfg = pkf.RecStackCfg(bidir=bool(self.bidir), cell=pk.CELL_LSTM, cell_flags=self.cell_flags,
                              grad_enabled=torch.is_grad_enabled())
        params = []
        for i, H in enumerate(self.lstm_lay):
            mask, mscal = self._mask(i, rows, H, x.device)
            cfg.layers.append(pkf.RecLayerCfg(H=H, act=pk.ACT_IDS[self.lstm_act[i]], use_bn=False, bn_training=False, bns=[],
                                              mask=mask, mask_scalar=mscal))
            ws = [getattr(self, w)[i] for w, _ in self._GATES]
            us = [getattr(self, u)[i] for _, u in self._GATES]
            params += [m.dense_weight() for m in ws] + [m.dense_weight() for m in us] + [m.dense_bias(x.device) for m in ws]
        with torch.cuda.device(x.device):
            return pkf.LiGRUStackFn.apply(x, cfg, *params)
