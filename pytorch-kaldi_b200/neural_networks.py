"""Drop-in `neural_networks` module for pytorch-kaldi recipes, B200-native compute.

Put this directory first on sys.path and every cfg that says `arch_library = neural_networks`
resolves to these classes (utils.model_init: importlib.import_module + getattr + Class(options,
inp_dim), reference utils.py:2047-2057).  The plug-in surface is kept bit-for-bit:

  * constructor signature `Class(options, inp_dim)` with the reference's option names
    (proto/*.proto) parsed from strings the same way (comma lists, strtobool);
  * the same sub-modules registered under the same names in the same order and created in the
    same sequence -> identical state_dict keys/shapes, optimizer parameter indices and
    CPU-generator consumption (reference neural_networks.py constructors, e.g. liGRU :998-1080);
  * `.out_dim`, train()/eval(), `to_do`/`use_cuda` options.

Only `forward` differs: it runs the hand-written sm_100a kernels of libpk_b200.so through
pk_functions (torch.autograd.Function over a C ABI).  There is NO eager-PyTorch or CPU
fallback: unsupported option combinations raise NotImplementedError, CPU tensors raise
RuntimeError.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn

import pk_functions as pkf
import pk_native as pk


def strtobool(val) -> int:
    """distutils.util.strtobool semantics (the reference parses cfg strings with it)."""
    v = str(val).lower()
    if v in ("y", "yes", "t", "true", "on", "1"):
        return 1
    if v in ("n", "no", "f", "false", "off", "0"):
        return 0
    raise ValueError(f"invalid truth value {val!r}")


class LayerNorm(nn.Module):
    """Reference neural_networks.py:23-33 — unbiased std, eps added to the std."""

    def __init__(self, features, eps=1e-6):
        super().__init__()
        self.gamma = nn.Parameter(torch.ones(features))
        self.beta = nn.Parameter(torch.zeros(features))
        self.eps = eps

    def forward(self, x):
        mean = x.mean(-1, keepdim=True)
        std = x.std(-1, keepdim=True)
        return self.gamma * (x - mean) / (std + self.eps) + self.beta


def act_fun(act_type):
    """Reference neural_networks.py:36-57 (module objects are kept for state/API parity; the
    kernels take the activation id)."""
    table = {
        "relu": nn.ReLU, "tanh": nn.Tanh, "sigmoid": nn.Sigmoid, "elu": nn.ELU,
        "leaky_relu": lambda: nn.LeakyReLU(0.2), "softmax": lambda: nn.LogSoftmax(dim=1),
        "linear": lambda: nn.LeakyReLU(1),
    }
    return table[act_type]() if act_type in table else None


def flip(x, dim):
    """Reference neural_networks.py:1962-1970 (API parity only: the kernels index time instead)."""
    return torch.flip(x, dims=[dim if dim >= 0 else x.dim() + dim])


def _floats(s):
    return list(map(float, str(s).split(",")))


def _ints(s):
    return list(map(int, str(s).split(",")))


def _bools(s):
    return list(map(strtobool, str(s).split(",")))


def _require_cuda(x, who):
    if not x.is_cuda:
        raise RuntimeError(f"pytorch-kaldi_b200.{who}: input is not a CUDA tensor — this library has no CPU path "
                           "(the CPU oracle lives under oracle/ and is test infrastructure only)")


# ---------------------------------------------------------------------------------------------
# MLP (reference neural_networks.py:60-150)
# ---------------------------------------------------------------------------------------------


class MLP(nn.Module):
    def __init__(self, options, inp_dim):
        super().__init__()
        self.input_dim = inp_dim
        self.dnn_lay = _ints(options["dnn_lay"])
        self.dnn_drop = _floats(options["dnn_drop"])
        self.dnn_use_batchnorm = _bools(options["dnn_use_batchnorm"])
        self.dnn_use_laynorm = _bools(options["dnn_use_laynorm"])
        self.dnn_use_laynorm_inp = strtobool(options["dnn_use_laynorm_inp"])
        self.dnn_use_batchnorm_inp = strtobool(options["dnn_use_batchnorm_inp"])
        self.dnn_act = str(options["dnn_act"]).split(",")

        self.wx = nn.ModuleList([])
        self.bn = nn.ModuleList([])
        self.ln = nn.ModuleList([])
        self.act = nn.ModuleList([])
        self.drop = nn.ModuleList([])
        if self.dnn_use_laynorm_inp:
            self.ln0 = LayerNorm(self.input_dim)
        if self.dnn_use_batchnorm_inp:
            self.bn0 = nn.BatchNorm1d(self.input_dim, momentum=0.05)
        self.N_dnn_lay = len(self.dnn_lay)
        cur = self.input_dim
        for i, width in enumerate(self.dnn_lay):
            self.drop.append(nn.Dropout(p=self.dnn_drop[i]))
            self.act.append(act_fun(self.dnn_act[i]))
            self.ln.append(LayerNorm(width))
            self.bn.append(nn.BatchNorm1d(width, momentum=0.05))
            add_bias = not (self.dnn_use_laynorm[i] or self.dnn_use_batchnorm[i])
            lin = nn.Linear(cur, width, bias=add_bias)  # draws from the generator like the reference (:111)
            bound = np.sqrt(0.01 / (cur + width))       # :114-119
            lin.weight = nn.Parameter(torch.Tensor(width, cur).uniform_(-bound, bound))
            lin.bias = nn.Parameter(torch.zeros(width))  # always re-created (:120)
            self.wx.append(lin)
            cur = width
        self.out_dim = cur

    def _is_plain_head(self):
        return (self.N_dnn_lay == 1 and self.dnn_act[0] == "softmax" and not self.dnn_use_batchnorm[0]
                and not self.dnn_use_laynorm[0] and not self.dnn_use_laynorm_inp and not self.dnn_use_batchnorm_inp
                and (self.dnn_drop[0] == 0.0 or not self.training))

    def forward(self, x):
        _require_cuda(x, "MLP")
        with torch.cuda.device(x.device):  # kernels launch on the tensor's device, not the thread's current one
            return self._forward_on_device(x)

    def _forward_on_device(self, x):
        if self._is_plain_head():
            return pkf.LinearLogSoftmaxFn.apply(x, self.wx[0].weight, self.wx[0].bias, torch.is_grad_enabled())
        return pkf.mlp_forward(self, x)


# ---------------------------------------------------------------------------------------------
# recurrent zoo: liGRU :997, GRU :486, minimalGRU :1158, RNN :1319, LSTM :300
# ---------------------------------------------------------------------------------------------


class _Recurrent(nn.Module):
    """Table-driven constructor shared by the five reference classes.  `_GATES` lists, in the
    reference's registration order, (input-projection list name, recurrent list name); the
    BatchNorm lists are `bn_<w-name>`."""

    _PREFIX = ""
    _GATES = ()
    _CELL = -1

    def __init__(self, options, inp_dim):
        super().__init__()
        p = self._PREFIX
        self.input_dim = inp_dim
        self.lay = _ints(options[p + "_lay"])
        self.drop = _floats(options[p + "_drop"])
        self.use_batchnorm = _bools(options[p + "_use_batchnorm"])
        self.use_laynorm = _bools(options[p + "_use_laynorm"])
        self.use_laynorm_inp = strtobool(options[p + "_use_laynorm_inp"])
        self.use_batchnorm_inp = strtobool(options[p + "_use_batchnorm_inp"])
        self.orthinit = strtobool(options[p + "_orthinit"])
        self.act_names = str(options[p + "_act"]).split(",")
        self.bidir = strtobool(options[p + "_bidir"])
        self.use_cuda = strtobool(options["use_cuda"])
        self.to_do = options["to_do"]
        self.test_flag = self.to_do != "train"
        # reference attribute names (kept for user code that pokes at them)
        for k in ("lay", "drop", "use_batchnorm", "use_laynorm", "use_laynorm_inp", "use_batchnorm_inp", "orthinit"):
            setattr(self, f"{p}_{k}", getattr(self, k))
        setattr(self, f"{p}_act", self.act_names)

        for w, u in self._GATES:  # registration order: (w, u) per gate ...
            setattr(self, w, nn.ModuleList([]))
            setattr(self, u, nn.ModuleList([]))
        self.ln = nn.ModuleList([])  # ... then ln, the per-gate BatchNorm lists, act
        for w, _ in self._GATES:
            setattr(self, "bn_" + w, nn.ModuleList([]))
        self.act = nn.ModuleList([])
        if self.use_laynorm_inp:
            self.ln0 = LayerNorm(self.input_dim)
        if self.use_batchnorm_inp:
            self.bn0 = nn.BatchNorm1d(self.input_dim, momentum=0.05)
        self.N_lay = len(self.lay)
        setattr(self, f"N_{p}_lay", self.N_lay)
        cur = self.input_dim
        for i, H in enumerate(self.lay):  # creation order inside a layer = generator consumption order
            self.act.append(act_fun(self.act_names[i]))
            add_bias = not (self.use_laynorm[i] or self.use_batchnorm[i])
            for w, _ in self._GATES:
                getattr(self, w).append(nn.Linear(cur, H, bias=add_bias))
            for _, u in self._GATES:
                getattr(self, u).append(nn.Linear(H, H, bias=False))
            if self.orthinit:
                for _, u in self._GATES:
                    nn.init.orthogonal_(getattr(self, u)[i].weight)
            for w, _ in self._GATES:
                getattr(self, "bn_" + w).append(nn.BatchNorm1d(H, momentum=0.05))
            self.ln.append(LayerNorm(H))
            cur = 2 * H if self.bidir else H
        self.out_dim = self.lay[-1] + self.bidir * self.lay[-1]
        self.fast_dropout = False  # True: draw the masks with the device generator (no H2D copy)
        self.cell_flags = 0        # tuning flags for the persistent kernel (pk_native.REC_*)

    # -- dropout masks: reference draws Bernoulli(1-p) on the CPU generator once per layer per
    #    forward and moves it to the device (:1102-1111); eval multiplies by the scalar 1-p
    def _mask(self, i, rows, H, device):
        if self.test_flag:
            return None, 1.0 - self.drop[i]
        if self.fast_dropout:
            return torch.empty(rows, H, device=device).bernoulli_(1.0 - self.drop[i]), 1.0
        m = torch.bernoulli(torch.Tensor(rows, H).fill_(1 - self.drop[i]))
        return m.to(device, non_blocking=True), 1.0

    def _check_supported(self):
        if self.use_laynorm_inp or self.use_batchnorm_inp or any(self.use_laynorm):
            raise NotImplementedError(
                f"pytorch-kaldi_b200.{type(self).__name__}: *_use_laynorm / *_use_laynorm_inp / *_use_batchnorm_inp "
                "are not implemented natively yet (no shipped recurrent recipe enables them); there is no fallback")
        for a in self.act_names:
            if a not in pk.ACT_IDS:
                raise NotImplementedError(f"activation {a!r} is not valid inside a recurrent layer")
        if self._CELL == pk.CELL_RNN and max(self.lay) > pkf.PERSISTENT_MAX_H:
            raise NotImplementedError(f"pytorch-kaldi_b200.RNN: rnn_lay > {pkf.PERSISTENT_MAX_H} is not supported (the plain RNN cell "
                                      "runs on the persistent kernels only); there is no fallback")

    def forward(self, x):
        _require_cuda(x, type(self).__name__)
        self._check_supported()
        T, B, _ = x.shape
        rows = (2 if self.bidir else 1) * B
        cfg = pkf.RecStackCfg(bidir=bool(self.bidir), cell=self._CELL, cell_flags=self.cell_flags,
                              grad_enabled=torch.is_grad_enabled())
        params = []
        for i, H in enumerate(self.lay):
            mask, mscal = self._mask(i, rows, H, x.device)
            ws = [getattr(self, w)[i] for w, _ in self._GATES]
            us = [getattr(self, u)[i] for _, u in self._GATES]
            bns = [getattr(self, "bn_" + w)[i] for w, _ in self._GATES]
            use_bn = bool(self.use_batchnorm[i])
            cfg.layers.append(pkf.RecLayerCfg(H=H, act=pk.ACT_IDS[self.act_names[i]], use_bn=use_bn,
                                              bn_training=self.training, bns=bns, mask=mask, mask_scalar=mscal))
            params += [m.weight for m in ws] + [m.weight for m in us]
            if use_bn:
                for bn in bns:
                    params += [bn.weight, bn.bias]
            else:
                params += [m.bias for m in ws]
        with torch.cuda.device(x.device):
            return pkf.LiGRUStackFn.apply(x, cfg, *params)


class liGRU(_Recurrent):
    _PREFIX, _CELL = "ligru", pk.CELL_LIGRU
    _GATES = (("wh", "uh"), ("wz", "uz"))


class GRU(_Recurrent):
    _PREFIX, _CELL = "gru", pk.CELL_GRU
    _GATES = (("wh", "uh"), ("wz", "uz"), ("wr", "ur"))


class minimalGRU(_Recurrent):
    _PREFIX, _CELL = "minimalgru", pk.CELL_MGRU
    _GATES = (("wh", "uh"), ("wz", "uz"))


class RNN(_Recurrent):
    _PREFIX, _CELL = "rnn", pk.CELL_RNN
    _GATES = (("wh", "uh"),)


class LSTM(_Recurrent):
    _PREFIX, _CELL = "lstm", pk.CELL_LSTM
    _GATES = (("wfx", "ufh"), ("wix", "uih"), ("wox", "uoh"), ("wcx", "uch"))


class _CudnnLayout(nn.Module):
    """LSTM_cudnn / GRU_cudnn / RNN_cudnn (reference neural_networks.py:153-297): thin wrappers over torch.nn.LSTM /
    GRU / RNN (per-direction weights, two biases, no BatchNorm, zero initial state, inter-layer nn.Dropout).  The
    constructor is the reference's (same sub-module, same init calls -> identical state_dict keys / shapes /
    generator consumption, so `cfg/TIMIT_baselines/TIMIT_LSTM_fmllr_cudnn.cfg` and its checkpoints drop in), but the
    forward never calls cuDNN: every (layer, direction) runs on this library's own recurrent kernels as a
    unidirectional, bias-only layer (pk_functions.LiGRUStackFn), the reverse direction on the time-flipped input."""

    _ATTR = ""      # ModuleList attribute name of the reference class
    _TORCH = None   # torch.nn class
    _CELL = -1
    _ORDER = ()     # torch's row-block order -> this library's gate order (blocks of H rows)

    def __init__(self, options, inp_dim):
        super().__init__()
        self.input_dim = inp_dim
        self.hidden_size = int(options["hidden_size"])
        self.num_layers = int(options["num_layers"])
        if self._TORCH is nn.RNN:
            self.nonlinearity = options["nonlinearity"]
        self.bias = bool(strtobool(options["bias"]))
        self.batch_first = bool(strtobool(options["batch_first"]))
        self.dropout = float(options["dropout"])
        self.bidirectional = bool(strtobool(options["bidirectional"]))
        kw = dict(bias=self.bias, dropout=self.dropout, bidirectional=self.bidirectional)
        if self._TORCH is nn.RNN:
            kw["nonlinearity"] = self.nonlinearity
        setattr(self, self._ATTR, nn.ModuleList([self._TORCH(self.input_dim, self.hidden_size, self.num_layers, **kw)]))
        self._init_like_reference()
        self.out_dim = self.hidden_size + self.bidirectional * self.hidden_size
        self.cell_flags = 0

    def _init_like_reference(self):
        pass

    def _act(self):
        return pk.ACT_IDS["tanh"]

    def forward(self, x):
        _require_cuda(x, type(self).__name__)
        rnn = getattr(self, self._ATTR)[0]
        H, G = self.hidden_size, len(self._ORDER)
        out = x
        with torch.cuda.device(x.device):
            for layer in range(self.num_layers):
                ys = []
                for d in range(2 if self.bidirectional else 1):
                    sfx = f"_l{layer}" + ("_reverse" if d else "")
                    w_ih, w_hh = getattr(rnn, "weight_ih" + sfx), getattr(rnn, "weight_hh" + sfx)
                    ws = [w_ih[g * H:(g + 1) * H] for g in self._ORDER]
                    us = [w_hh[g * H:(g + 1) * H] for g in self._ORDER]
                    if self.bias:
                        b = getattr(rnn, "bias_ih" + sfx) + getattr(rnn, "bias_hh" + sfx)
                        bs = [b[g * H:(g + 1) * H] for g in self._ORDER]
                    else:
                        bs = [torch.zeros(H, device=x.device) for _ in range(G)]
                    cfg = pkf.RecStackCfg(bidir=False, cell=self._CELL, cell_flags=self.cell_flags,
                                          grad_enabled=torch.is_grad_enabled())
                    cfg.layers.append(pkf.RecLayerCfg(H=H, act=self._act(), use_bn=False, bn_training=False, bns=[],
                                                      mask=None, mask_scalar=1.0))
                    xin = torch.flip(out, dims=[0]) if d else out       # reverse direction: time-flipped input ...
                    y = pkf.LiGRUStackFn.apply(xin, cfg, *ws, *us, *bs)
                    ys.append(torch.flip(y, dims=[0]) if d else y)      # ... and output
                out = torch.cat(ys, dim=2) if len(ys) > 1 else ys[0]
                if self.dropout > 0.0 and self.training and layer < self.num_layers - 1:
                    out = torch.nn.functional.dropout(out, self.dropout, True)  # nn.LSTM's inter-layer dropout
        return out


class LSTM_cudnn(_CudnnLayout):
    """nn.LSTM gate algebra (i, f, g, o row blocks; c = f c + i tanh(g); h = o tanh(c)) = this library's LSTM cell with
    act = tanh and no mask; gate order of the kernels: f, i, o, c."""
    _ATTR, _TORCH, _CELL, _ORDER = "lstm", nn.LSTM, pk.CELL_LSTM, (1, 0, 3, 2)

    def _init_like_reference(self):
        for name, param in self.lstm[0].named_parameters():  # reference :178-184
            if "weight_hh" in name:
                if self.batch_first:
                    nn.init.orthogonal_(param)
            elif "bias" in name:
                nn.init.zeros_(param)


class RNN_cudnn(_CudnnLayout):
    """nn.RNN: h = act(W x + b_ih + U h + b_hh), act in {tanh, relu}."""
    _ATTR, _TORCH, _CELL, _ORDER = "rnn", nn.RNN, pk.CELL_RNN, (0,)

    def _act(self):
        return pk.ACT_IDS[self.nonlinearity]


class GRU_cudnn(_CudnnLayout):
    """nn.GRU computes the candidate as tanh(W x + b + r * (U h + b_hn)); the reference's own GRU class — and this
    library's GRU kernels — contract U with (r * h) instead (neural_networks.py:634).  The constructor (state_dict
    layout, init) is provided; the forward is refused rather than silently computing different math."""
    _ATTR, _TORCH, _CELL, _ORDER = "gru", nn.GRU, pk.CELL_GRU, (2, 1, 0)

    def _init_like_reference(self):
        for name, param in self.gru[0].named_parameters():  # reference :229-235
            if "weight_hh" in name:
                nn.init.orthogonal_(param)
            elif "weight_ih" in name:
                nn.init.xavier_uniform_(param)
            elif "bias" in name:
                nn.init.zeros_(param)

    def forward(self, x):  # noqa: D401
        raise NotImplementedError("pytorch-kaldi_b200.GRU_cudnn: nn.GRU's candidate gate r * (U h + b) is not what the native "
                                  "GRU kernels compute (U (r * h), the reference's own GRU class); use arch_class = GRU")


# ---------------------------------------------------------------------------------------------
# FusionRNN: FusionLinearConv :2057-2099, liGRU_layer :795-995, fusionRNN_jit :719-793
# ---------------------------------------------------------------------------------------------


class FusionLinearConv(nn.Module):
    """Reference :2057-2099 — one affine map shared by the `number_of_mic` channels that are concatenated along the last
    dimension (Conv1d(1, out, kernel = in/mic, stride = in/mic)), an activation, and a sum / mean over the channels.
    Same constructor, parameters and init; the forward is the native GEMM + reduction (pk_functions.FusionProjFn)."""

    def __init__(self, in_features, out_features, number_of_mic=1, bias=True, seed=None, act="leaky", reduce="sum"):
        super().__init__()
        self.in_features = in_features // number_of_mic
        self.out_features = out_features
        self.number_of_mic = number_of_mic
        self.reduce = reduce
        if act == "leaky_relu":
            self.act_function = nn.LeakyReLU()
        elif act == "prelu":
            self.act_function = nn.PReLU()
        elif act == "relu":
            self.act_function = nn.ReLU()
        else:
            self.act_function = nn.Tanh()
        self.conv = nn.Conv1d(1, self.out_features, kernel_size=self.in_features, stride=self.in_features, bias=True, padding=0)
        self.conv.bias.data.fill_(0)
        torch.nn.init.xavier_normal_(self.conv.weight.data)

    def _fusion_cfg(self):
        a = self.act_function
        if isinstance(a, nn.PReLU):
            mode, prelu, slope = 0, True, 0.0
        elif isinstance(a, nn.LeakyReLU):
            mode, prelu, slope = 0, False, float(a.negative_slope)
        elif isinstance(a, nn.ReLU):
            mode, prelu, slope = 0, False, 0.0
        else:
            mode, prelu, slope = 1, False, 0.0
        red = 1.0 / self.number_of_mic if self.reduce == "mean" else 1.0
        return pkf.FusionCfg(M=self.number_of_mic, mode=mode, red=red, prelu=prelu, slope=slope,
                             grad_enabled=torch.is_grad_enabled())

    def _fusion_params(self):
        p = [self.conv.weight, self.conv.bias]
        if isinstance(self.act_function, nn.PReLU):
            p.append(self.act_function.weight)
        return p

    def forward(self, input):
        _require_cuda(input, "FusionLinearConv")
        with torch.cuda.device(input.device):
            return pkf.FusionProjFn.apply(input, self._fusion_cfg(), *self._fusion_params())


class liGRU_layer(nn.Module):
    """Reference :795-995 (a torch.jit.ScriptModule there): one bidirectional liGRU layer with the two input
    projections `wz`, `wh` (nn.Linear with bias, or FusionLinearConv when `do_fusion`), BatchNorm on both, ONE stacked
    recurrent matrix `u` = [Uz; Uh] (:866-868, chunked as (uz, uh) at :975), ReLU candidate and an nn.Dropout-style
    mask (kept entries scaled by 1/(1-p)) that is constant over time (:935-970).  Same constructor (sub-module names,
    creation order, init) -> identical state_dict and generator consumption; the forward runs the persistent liGRU
    kernels: gate blocks re-ordered to this library's (h, z), the Linear biases cancel against BatchNorm."""

    def __init__(self, input_size, hidden_size, num_layers, batch_size, dropout=0.0, nonlinearity="relu", bidirectional=True,
                 device="cuda", do_fusion=False, fusion_layer_size=64, number_of_mic=1, act="relu", reduce="mean"):
        super().__init__()
        self.hidden_size = int(hidden_size)
        self.input_size = int(input_size)
        self.batch_size = batch_size
        self.bidirectional = bidirectional
        self.dropout = dropout
        self.device = device
        self.do_fusion = bool(do_fusion)
        self.fusion_layer_size = fusion_layer_size
        self.number_of_mic = number_of_mic
        self.reduce = reduce
        if self.do_fusion:
            self.hidden_size = self.fusion_layer_size // self.number_of_mic
            self.wz = FusionLinearConv(self.input_size, self.hidden_size, bias=True, number_of_mic=self.number_of_mic, act=act,
                                       reduce=self.reduce)
            self.wh = FusionLinearConv(self.input_size, self.hidden_size, bias=True, number_of_mic=self.number_of_mic, act=act,
                                       reduce=self.reduce)
        else:
            self.wz = nn.Linear(self.input_size, self.hidden_size, bias=True)
            self.wh = nn.Linear(self.input_size, self.hidden_size, bias=True)
            self.wz.bias.data.fill_(0)
            torch.nn.init.xavier_normal_(self.wz.weight.data)
            self.wh.bias.data.fill_(0)
            torch.nn.init.xavier_normal_(self.wh.weight.data)
        self.u = nn.Linear(self.hidden_size, 2 * self.hidden_size, bias=False)
        nn.init.orthogonal_(self.u.weight)
        self.bn_wh = nn.BatchNorm1d(self.hidden_size, momentum=0.05)
        self.bn_wz = nn.BatchNorm1d(self.hidden_size, momentum=0.05)
        self.drop = torch.nn.Dropout(p=self.dropout, inplace=False)
        self.N_drop_masks = 100
        self.drop_mask_cnt = 0
        self.act = torch.nn.ReLU()
        self.cell_flags = 0
        self._mask_override = None  # tests: a [ndir*B, H] mask (already scaled) instead of a fresh draw

    def _mask(self, rows, device):
        if not self.training:
            return None, 1.0          # drop_mask_te = 1.0 (:875)
        if self._mask_override is not None:
            return self._mask_override.to(device), 1.0
        if self.dropout <= 0.0:
            return None, 1.0
        keep = 1.0 - self.dropout
        return torch.empty(rows, self.hidden_size, device=device).bernoulli_(keep).mul_(1.0 / keep), 1.0

    def forward(self, x):
        _require_cuda(x, "liGRU_layer")
        T, B, _ = x.shape
        H = self.hidden_size
        rows = (2 if self.bidirectional else 1) * B
        with torch.cuda.device(x.device):
            if self.do_fusion:
                # both fused projections in one GEMM: [T, B, 2H] = [P_h | P_z]; the recurrent stack then sees them
                # through identity "weights" (exact in the fp32-accumulating GEMM, one fp16 rounding of P)
                cfgf = self.wh._fusion_cfg()
                x = pkf.FusionProjFn.apply(x, cfgf, *self.wh._fusion_params(), *self.wz._fusion_params())
                eye = torch.eye(2 * H, device=x.device)
                ws, bias = [eye[:H], eye[H:]], None
            else:
                ws, bias = [self.wh.weight, self.wz.weight], [self.wh.bias, self.wz.bias]
            mask, mscal = self._mask(rows, x.device)
            cfg = pkf.RecStackCfg(bidir=bool(self.bidirectional), cell=pk.CELL_LIGRU, cell_flags=self.cell_flags,
                                  grad_enabled=torch.is_grad_enabled())
            cfg.layers.append(pkf.RecLayerCfg(H=H, act=pk.ACT_IDS["relu"], use_bn=True, bn_training=self.training,
                                              bns=[self.bn_wh, self.bn_wz], mask=mask, mask_scalar=mscal, proj_bias=bias))
            us = [self.u.weight[H:], self.u.weight[:H]]   # (uz, uh) = u(h).chunk(2, 1)  ->  this library's (h, z)
            return pkf.LiGRUStackFn.apply(x, cfg, *ws, *us, self.bn_wh.weight, self.bn_wh.bias, self.bn_wz.weight,
                                          self.bn_wz.bias)


class fusionRNN_jit(nn.Module):
    """Reference :719-793 (cfg/DIRHA_baselines/DIRHA_fusionRNN_MFCC_6ch.cfg): a stack of liGRU_layer, the first one
    with FusionLinearConv input projections over the microphone channels.  The reference stores
    `map(strtobool, options["fusionRNN_do_fusion"])` — a map object, always truthy — so the FIRST layer always fuses and
    the others never do; kept.  `options["batches"]` (the reference's fixed batch size, :727) is optional here: the
    kernels take the batch size from the input."""

    def __init__(self, options, inp_dim):
        super().__init__()
        input_size = inp_dim
        lay = _ints(options["fusionRNN_lay"])
        hidden_size = lay[0]
        dropout = _floats(options["fusionRNN_drop"])[0]
        num_layers = len(lay)
        batch_size = int(options["batches"]) if "batches" in options else 0
        self.do_fusion = True
        self.act = str(options["fusionRNN_fusion_act"])
        self.reduce = str(options["fusionRNN_fusion_reduce"])
        self.fusion_layer_size = int(options["fusionRNN_fusion_layer_size"])
        self.to_do = options["to_do"]
        self.number_of_mic = int(options["fusionRNN_number_of_mic"])
        self.save_mic = self.number_of_mic
        bidirectional = True
        self.out_dim = 2 * hidden_size
        current_dim = int(input_size)
        self.model = torch.nn.ModuleList([])
        self.training = self.to_do == "train"
        for i in range(num_layers):
            rnn_lay = liGRU_layer(current_dim, hidden_size, num_layers, batch_size, dropout=dropout, bidirectional=bidirectional,
                                  device="cuda", do_fusion=self.do_fusion, fusion_layer_size=self.fusion_layer_size,
                                  number_of_mic=self.number_of_mic, act=self.act, reduce=self.reduce)
            if i == 0:
                current_dim = (self.fusion_layer_size // self.save_mic) * 2
                self.number_of_mic = 1
                self.do_fusion = False
            else:
                current_dim = hidden_size * 2
            self.model.append(rnn_lay)

    def forward(self, x):
        for ligru_lay in self.model:
            x = ligru_lay(x)
        return x


# ---------------------------------------------------------------------------------------------
# convolutional front-ends: CNN :1464-1556, SincNet :1559-1665, SincConv :1668-1813
# ---------------------------------------------------------------------------------------------


class SincConv(nn.Module):
    """Band-pass filterbank layer parametrised by (low_hz_, band_hz_) — reference :1668-1813.  Same constructor,
    parameters and derived constants; the filters are synthesised and applied by the native path."""

    @staticmethod
    def to_mel(hz):
        return 2595 * np.log10(1 + hz / 700)

    @staticmethod
    def to_hz(mel):
        return 700 * (10 ** (mel / 2595) - 1)

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, bias=False, groups=1,
                 sample_rate=16000, min_low_hz=50, min_band_hz=50):
        super().__init__()
        if in_channels != 1:
            raise ValueError("SincConv only support one input channel (here, in_channels = {%i})" % (in_channels))
        if bias:
            raise ValueError("SincConv does not support bias.")
        if groups > 1:
            raise ValueError("SincConv does not support groups.")
        self.out_channels = out_channels
        self.kernel_size = kernel_size + (1 - kernel_size % 2)  # odd length -> symmetric filters (:1722-1724)
        self.stride, self.padding, self.dilation = stride, padding, dilation
        self.sample_rate, self.min_low_hz, self.min_band_hz = sample_rate, min_low_hz, min_band_hz
        # mel-spaced initial band edges, normalised by the sample rate (:1740-1750)
        hi = self.sample_rate / 2 - (self.min_low_hz + self.min_band_hz)
        hz = self.to_hz(np.linspace(self.to_mel(30), self.to_mel(hi), self.out_channels + 1)) / self.sample_rate
        self.low_hz_ = nn.Parameter(torch.Tensor(hz[:-1]).view(-1, 1))
        self.band_hz_ = nn.Parameter(torch.Tensor(np.diff(hz)).view(-1, 1))
        n_lin = torch.linspace(0, self.kernel_size, steps=self.kernel_size)
        self.window_ = 0.54 - 0.46 * torch.cos(2 * math.pi * n_lin / self.kernel_size)
        n = (self.kernel_size - 1) / 2
        self.n_ = torch.arange(-n, n + 1).view(1, -1) / self.sample_rate

    def forward(self, waveforms):
        """waveforms [N, 1, n_samples] -> [N, out_channels, n_samples - kernel_size + 1]"""
        _require_cuda(waveforms, "SincConv")
        if (self.stride, self.padding, self.dilation) != (1, 0, 1) or waveforms.shape[1] != 1:
            raise NotImplementedError("pytorch-kaldi_b200.SincConv: only stride 1 / no padding / no dilation (what "
                                      "SincNet uses) is implemented natively")
        cfg = pkf.ConvStackCfg(flat_output=False, grad_enabled=torch.is_grad_enabled())
        cfg.layers.append(pkf.ConvLayerCfg(kind="sinc", C=self.out_channels, k=self.kernel_size, pool=1,
                                           act=pk.ACT_IDS["linear"], use_ln=False, sample_rate=self.sample_rate,
                                           min_low_hz=self.min_low_hz, min_band_hz=self.min_band_hz))
        with torch.cuda.device(waveforms.device):
            return pkf.ConvStackFn.apply(waveforms[:, 0, :], cfg, self.low_hz_, self.band_hz_)


class SincConv_fast(nn.Module):
    """Reference :1816-1959 (Hz-domain parametrisation, half-window evaluation).  No module of the zoo
    instantiates it (SincNet uses SincConv, :1619-1628); constructor parity only."""

    to_mel = SincConv.to_mel
    to_hz = SincConv.to_hz

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, bias=False, groups=1,
                 sample_rate=16000, min_low_hz=50, min_band_hz=50):
        super().__init__()
        if in_channels != 1:
            raise ValueError("SincConv only support one input channel (here, in_channels = {%i})" % (in_channels))
        if bias:
            raise ValueError("SincConv does not support bias.")
        if groups > 1:
            raise ValueError("SincConv does not support groups.")
        self.out_channels = out_channels
        self.kernel_size = kernel_size + (1 - kernel_size % 2)
        self.stride, self.padding, self.dilation = stride, padding, dilation
        self.sample_rate, self.min_low_hz, self.min_band_hz = sample_rate, min_low_hz, min_band_hz
        hi = self.sample_rate / 2 - (self.min_low_hz + self.min_band_hz)
        hz = self.to_hz(np.linspace(self.to_mel(30), self.to_mel(hi), self.out_channels + 1))
        self.low_hz_ = nn.Parameter(torch.Tensor(hz[:-1]).view(-1, 1))
        self.band_hz_ = nn.Parameter(torch.Tensor(np.diff(hz)).view(-1, 1))
        n_lin = torch.linspace(0, (self.kernel_size / 2) - 1, steps=int((self.kernel_size / 2)))
        self.window_ = 0.54 - 0.46 * torch.cos(2 * math.pi * n_lin / self.kernel_size)
        n = (self.kernel_size - 1) / 2.0
        self.n_ = 2 * math.pi * torch.arange(-n, 0).view(1, -1) / self.sample_rate

    def forward(self, waveforms):
        raise NotImplementedError("pytorch-kaldi_b200.SincConv_fast: not used by any reference module; use SincConv")


class _ConvFrontEnd(nn.Module):
    """Shared constructor of CNN (:1465-1528) and SincNet (:1560-1636): per layer Dropout, activation, LayerNorm
    over [N_filt, L_pooled], BatchNorm1d (constructed exactly like the reference does — its second positional
    argument lands in `eps`), and the convolution (a SincConv for SincNet's first layer)."""

    _PREFIX = ""
    _SINC = False

    def __init__(self, options, inp_dim):
        super().__init__()
        p = self._PREFIX
        self.input_dim = inp_dim
        vals = dict(N_filt=_ints(options[p + "_N_filt"]), len_filt=_ints(options[p + "_len_filt"]),
                    max_pool_len=_ints(options[p + "_max_pool_len"]), act=str(options[p + "_act"]).split(","),
                    drop=_floats(options[p + "_drop"]), use_laynorm=_bools(options[p + "_use_laynorm"]),
                    use_batchnorm=_bools(options[p + "_use_batchnorm"]),
                    use_laynorm_inp=strtobool(options[p + "_use_laynorm_inp"]),
                    use_batchnorm_inp=strtobool(options[p + "_use_batchnorm_inp"]))
        for k, v in vals.items():
            setattr(self, f"{p}_{k}", v)
        n_lay = len(vals["N_filt"])
        setattr(self, f"N_{p}_lay", n_lay)
        if self._SINC:
            self.sinc_sample_rate = int(options["sinc_sample_rate"])
            self.sinc_min_low_hz = int(options["sinc_min_low_hz"])
            self.sinc_min_band_hz = int(options["sinc_min_band_hz"])
        self.conv = nn.ModuleList([])
        self.bn = nn.ModuleList([])
        self.ln = nn.ModuleList([])
        self.act = nn.ModuleList([])
        self.drop = nn.ModuleList([])
        if vals["use_laynorm_inp"]:
            self.ln0 = LayerNorm(self.input_dim)
        if vals["use_batchnorm_inp"]:
            self.bn0 = nn.BatchNorm1d([self.input_dim], momentum=0.05)
        cur = self.input_dim
        for i in range(n_lay):
            n_filt, len_filt, pool = vals["N_filt"][i], vals["len_filt"][i], vals["max_pool_len"][i]
            pooled = int((cur - len_filt + 1) / pool)
            self.drop.append(nn.Dropout(p=vals["drop"][i]))
            self.act.append(act_fun(vals["act"][i]))
            self.ln.append(LayerNorm([n_filt, pooled]))
            self.bn.append(nn.BatchNorm1d(n_filt, pooled, momentum=0.05))
            if i == 0 and self._SINC:
                self.conv.append(SincConv(1, n_filt, len_filt, sample_rate=self.sinc_sample_rate,
                                          min_low_hz=self.sinc_min_low_hz, min_band_hz=self.sinc_min_band_hz))
            else:
                self.conv.append(nn.Conv1d(1 if i == 0 else vals["N_filt"][i - 1], n_filt, len_filt))
            cur = pooled
        self.out_dim = cur * n_filt

    def forward(self, x):
        _require_cuda(x, type(self).__name__)
        with torch.cuda.device(x.device):
            return pkf.conv_forward(self, x, self._PREFIX)


class CNN(_ConvFrontEnd):
    _PREFIX, _SINC = "cnn", False


class SincNet(_ConvFrontEnd):
    _PREFIX, _SINC = "sinc", True
