"""CPU oracle for the acoustic-model hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain-numpy restatement of the reference algorithm (mravanelli/pytorch-kaldi) for the path
BASELINE.json names: module zoo forward/backward (neural_networks.py) + the cost ops of
utils.forward_model + the optimizer step of core.run_nn.  Every function cites the reference
file:line it follows.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this file; the product (pytorch-kaldi_b200/) never does.

Pinning: the reference ships no golden vectors or tests for this path, so this
restatement is pinned against outputs of the reference itself: tests/golden/make_golden.py
imports /root/reference/neural_networks.py in the build container, runs it on seeded inputs and
commits the vectors under tests/golden/*.npz; tests/test_oracle.py checks this file against
them (forward, loss, error rate, every parameter gradient, one optimizer step).

The restatement is deliberately literal (it concatenates x and flip(x) on the batch axis and
normalises the duplicated rows exactly like the reference does) so that the de-duplicated,
fused CUDA path is checked against the reference's own formulation, not against itself.
Backward passes are derived by hand (no autograd) -> an independent check of the kernels' math.
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------------------
# activations  (neural_networks.py:36-57 act_fun)
# --------------------------------------------------------------------------------------


def act_fwd(name: str, x: np.ndarray) -> np.ndarray:
    if name == "relu":
        return np.maximum(x, 0)
    if name == "tanh":
        return np.tanh(x)
    if name == "sigmoid":
        return 1.0 / (1.0 + np.exp(-x))
    if name == "leaky_relu":  # nn.LeakyReLU(0.2), neural_networks.py:47-48
        return np.where(x > 0, x, 0.2 * x)
    if name == "elu":
        return np.where(x > 0, x, np.expm1(np.minimum(x, 0)))
    if name == "linear":  # nn.LeakyReLU(1) == identity, neural_networks.py:56-57
        return x
    raise ValueError(f"unknown activation {name}")


def act_bwd(name: str, x: np.ndarray, y: np.ndarray) -> np.ndarray:
    """d act / d x given pre-activation x and output y."""
    if name == "relu":
        return (x > 0).astype(x.dtype)
    if name == "tanh":
        return 1.0 - y * y
    if name == "sigmoid":
        return y * (1.0 - y)
    if name == "leaky_relu":
        return np.where(x > 0, 1.0, 0.2).astype(x.dtype)
    if name == "elu":
        return np.where(x > 0, 1.0, y + 1.0).astype(x.dtype)
    if name == "linear":
        return np.ones_like(x)
    raise ValueError(f"unknown activation {name}")


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def q16(a: np.ndarray) -> np.ndarray:
    """Round to fp16 and back: emulates the tensor-core OPERAND rounding of the CUDA path (fp32
    accumulation is kept).  Only used when `quant=True` (see ligru_forward)."""
    return a.astype(np.float16).astype(a.dtype)


def flip0(x: np.ndarray) -> np.ndarray:
    """neural_networks.py:1962-1970 flip(x, 0): time reversal."""
    return x[::-1].copy()


# --------------------------------------------------------------------------------------
# normalisation layers
# --------------------------------------------------------------------------------------


def layernorm_fwd(x, gamma, beta, eps=1e-6):
    """neural_networks.py:23-33 custom LayerNorm: UNBIASED std, eps added to the std."""
    n = x.shape[-1]
    mean = x.mean(-1, keepdims=True)
    xc = x - mean
    std = np.sqrt((xc * xc).sum(-1, keepdims=True) / (n - 1))
    y = gamma * xc / (std + eps) + beta
    return y, (xc, std, gamma, eps)


def layernorm_bwd(dy, cache):
    xc, std, gamma, eps = cache
    n = xc.shape[-1]
    dgamma = (dy * xc / (std + eps)).reshape(-1, n).sum(0)
    dbeta = dy.reshape(-1, n).sum(0)
    g = dy * gamma
    # y = g * xc / (std+eps); std = sqrt(sum xc^2/(n-1)); xc = x - mean
    dxc = g / (std + eps)
    dstd = -(g * xc).sum(-1, keepdims=True) / (std + eps) ** 2
    dxc = dxc + dstd * xc / ((n - 1) * np.maximum(std, 1e-30))
    dx = dxc - dxc.mean(-1, keepdims=True)
    return dx, dgamma, dbeta


def batchnorm_fwd(x2d, bn, training):
    """nn.BatchNorm1d(C, momentum=0.05) on [N, C] rows (neural_networks.py:1070-1071, :1118-1124).
    `bn` is a dict(weight, bias, running_mean, running_var, num_batches_tracked, eps, momentum);
    running stats are updated in place when training (biased var for normalisation, unbiased for
    the running estimate — torch semantics)."""
    eps = bn.get("eps", 1e-5)
    if training:
        n = x2d.shape[0]
        mean = x2d.mean(0)
        var = ((x2d - mean) ** 2).mean(0)
        m = bn.get("momentum", 0.05)
        bn["running_mean"] = (1 - m) * bn["running_mean"] + m * mean
        bn["running_var"] = (1 - m) * bn["running_var"] + m * var * n / max(n - 1, 1)
        bn["num_batches_tracked"] = bn.get("num_batches_tracked", 0) + 1
    else:
        mean, var = bn["running_mean"], bn["running_var"]
    rstd = 1.0 / np.sqrt(var + eps)
    xhat = (x2d - mean) * rstd
    y = xhat * bn["weight"] + bn["bias"]
    return y, (xhat, rstd, bn["weight"], training)


def batchnorm_bwd(dy, cache):
    xhat, rstd, gamma, training = cache
    dgamma = (dy * xhat).sum(0)
    dbeta = dy.sum(0)
    if training:
        n = dy.shape[0]
        dx = gamma * rstd * (dy - dbeta / n - xhat * dgamma / n)
    else:
        dx = gamma * rstd * dy
    return dx, dgamma, dbeta


# --------------------------------------------------------------------------------------
# liGRU  (neural_networks.py:997-1155)
# --------------------------------------------------------------------------------------


def ligru_forward(x, layers, *, bidir=True, training=True, masks=None, quant=False, cell="ligru"):
    """x [T,B,D] -> [T,B,(2)H_last].

    cell="ligru": reference liGRU (:997-1155).  cell="rnn": reference RNN (:1319-1461), i.e. the same loop with
    no update gate: h_t = act(wh_k + Uh h) * mask (:1438-1447); layer dicts then carry wh/uh/bn_wh/bh only.

    quant=False is the reference's fp32 algorithm.  quant=True rounds every GEMM operand (inputs,
    weights, recurrent state) to fp16 exactly where the CUDA path does, so that non-smooth points
    (ReLU kinks) fall on the same side in both and gradients can be compared tightly.

    layers: list of dicts with keys
        wh, wz   [H,D]   (nn.Linear weights, :1054-1059)      bh, bz [H] or None (bias only when no BN/LN)
        uh, uz   [H,H]   (:1061-1063)
        bn_wh, bn_wz     BatchNorm dicts or None (:1070-1071)
        act      activation name, drop  dropout probability
    masks: per-layer [rows,H] Bernoulli(1-p) masks (training; reference draws them on the CPU
           generator, :1102-1105) or None -> eval scalar (1-p) (:1107).
    Returns (out, caches).
    """
    caches = []
    for li, L in enumerate(layers):
        T, B, D = x.shape
        H = L["uh"].shape[0]
        xin = x
        if bidir:  # :1095-1097
            x2 = np.concatenate([x, flip0(x)], axis=1)
        else:
            x2 = x
        R = x2.shape[1]
        if training and masks is not None:
            mask = masks[li].astype(x.dtype)
        else:
            mask = np.asarray(1.0 - L["drop"], dtype=x.dtype)  # :1107
        Q = q16 if quant else (lambda a: a)
        x2q = Q(x2)
        rnn = cell == "rnn"
        wh_q, uh_q = Q(L["wh"]), Q(L["uh"])
        wz_q = np.zeros_like(wh_q) if rnn else Q(L["wz"])
        uz_q = np.zeros_like(uh_q) if rnn else Q(L["uz"])
        wh_out = x2q @ wh_q.T  # :1114
        wz_out = x2q @ wz_q.T  # :1115
        if L.get("bh") is not None:
            wh_out = wh_out + L["bh"]
            if not rnn:
                wz_out = wz_out + L["bz"]
        bn_cache_h = bn_cache_z = None
        if L.get("bn_wh") is not None:  # :1118-1124
            yh, bn_cache_h = batchnorm_fwd(wh_out.reshape(T * R, H), L["bn_wh"], training)
            wh_out = yh.reshape(T, R, H)
            if not rnn:
                yz, bn_cache_z = batchnorm_fwd(wz_out.reshape(T * R, H), L["bn_wz"], training)
                wz_out = yz.reshape(T, R, H)
        ht = np.zeros((R, H), dtype=x.dtype)  # :1096
        hs = np.zeros((T, R, H), dtype=x.dtype)
        zs = np.zeros_like(hs)
        ats = np.zeros_like(hs)
        for k in range(T):  # :1130-1141
            htq = Q(ht)
            zt = np.zeros_like(ht) if rnn else sigmoid(wz_out[k] + htq @ uz_q.T)
            at = wh_out[k] + htq @ uh_q.T
            hcand = act_fwd(L["act"], at) * mask
            ht = zt * ht + (1 - zt) * hcand
            hs[k], zs[k], ats[k] = ht, zt, at
        if bidir:  # :1147-1150
            h_f = hs[:, :B]
            h_b = flip0(hs[:, B:])
            out = np.concatenate([h_f, h_b], axis=2)
        else:
            out = hs
        caches.append(dict(x2=x2q, hs=hs, zs=zs, ats=ats, mask=mask, bn_h=bn_cache_h, bn_z=bn_cache_z, B=B,
                           xin_shape=xin.shape, wq=(wh_q, wz_q, uh_q, uz_q), Q=Q, rnn=rnn))
        x = out
    return x, caches


def ligru_backward(dout, layers, caches, *, bidir=True):
    """Hand-derived BPTT of ligru_forward.  Returns (dx, grads) with grads[i] = dict(wh, wz, uh,
    uz, bh, bz, bn_wh_weight, bn_wh_bias, bn_wz_weight, bn_wz_bias)."""
    grads = [None] * len(layers)
    for li in reversed(range(len(layers))):
        L, c = layers[li], caches[li]
        x2, hs, zs, ats, mask, B = c["x2"], c["hs"], c["zs"], c["ats"], c["mask"], c["B"]
        T, R, H = hs.shape
        wh_q, wz_q, uh_q, uz_q = c["wq"]
        Q = c["Q"]
        if bidir:
            dH = np.concatenate([dout[:, :, :H], flip0(dout[:, :, H:])], axis=1)
        else:
            dH = dout
        carry = np.zeros((R, H), dtype=dout.dtype)
        da_all = np.zeros_like(hs)
        dz_all = np.zeros_like(hs)
        duh = np.zeros_like(L["uh"])
        duz = np.zeros_like(L["uh"])
        for k in reversed(range(T)):
            hprev = hs[k - 1] if k > 0 else np.zeros((R, H), dtype=dout.dtype)
            zt, at = zs[k], ats[k]
            y = act_fwd(L["act"], at)
            hcand = y * mask
            dh = dH[k] + carry
            dzt = dh * (hprev - hcand)
            dhc = dh * (1 - zt)
            da = dhc * mask * act_bwd(L["act"], at, y)
            dzp = dzt * zt * (1 - zt)
            carry = dh * zt + da @ uh_q + dzp @ uz_q
            duh += da.T @ Q(hprev)
            duz += dzp.T @ Q(hprev)
            da_all[k], dz_all[k] = da, dzp
        g = dict(uh=duh) if c["rnn"] else dict(uh=duh, uz=duz)
        dwh_pre = da_all.reshape(T * R, H)
        dwz_pre = dz_all.reshape(T * R, H)
        if c["bn_h"] is not None:
            dwh_pre, g["bn_wh_weight"], g["bn_wh_bias"] = batchnorm_bwd(dwh_pre, c["bn_h"])
            if not c["rnn"]:
                dwz_pre, g["bn_wz_weight"], g["bn_wz_bias"] = batchnorm_bwd(dwz_pre, c["bn_z"])
        if L.get("bh") is not None:
            g["bh"] = dwh_pre.sum(0)
            if not c["rnn"]:
                g["bz"] = dwz_pre.sum(0)
        x2f = x2.reshape(T * R, -1)
        g["wh"] = dwh_pre.T @ x2f
        if not c["rnn"]:
            g["wz"] = dwz_pre.T @ x2f
        dx2 = (dwh_pre @ wh_q + dwz_pre @ wz_q).reshape(T, R, -1)
        if bidir:
            dout = dx2[:, :B] + flip0(dx2[:, B:])
        else:
            dout = dx2
        grads[li] = g
    return dout, grads


# --------------------------------------------------------------------------------------
# LSTM :300-483, GRU :486-654, minimalGRU :1158-1316 — same scaffolding as the liGRU, other step equations
# --------------------------------------------------------------------------------------

GATES = {"lstm": ("f", "i", "o", "c"), "gru": ("h", "z", "r"), "minimalgru": ("h", "z")}


def cell_forward(x, layers, *, cell, bidir=True, training=True, masks=None, quant=False):
    """x [T,B,D] -> [T,B,(2)H_last] for cell in {"lstm", "gru", "minimalgru"}.

    layers: list of dicts with w / u (lists of [H,D] / [H,H] in the reference's registration order GATES[cell]),
    b (list of biases or None), bn (list of BatchNorm dicts or None), act, drop.
    Step equations: LSTM :457-469, GRU :631-637, minimalGRU :1293-1297.  quant as in ligru_forward."""
    names = GATES[cell]
    caches = []
    for li, L in enumerate(layers):
        T, B, D = x.shape
        H = L["u"][0].shape[0]
        x2 = np.concatenate([x, flip0(x)], axis=1) if bidir else x
        R = x2.shape[1]
        if training and masks is not None:
            mask = masks[li].astype(x.dtype)
        else:
            mask = np.asarray(1.0 - L["drop"], dtype=x.dtype)
        Q = q16 if quant else (lambda a: a)
        x2q = Q(x2)
        wq = [Q(w) for w in L["w"]]
        uq = dict(zip(names, [Q(u) for u in L["u"]]))
        pre, bnc = {}, {}
        for gi, n in enumerate(names):
            p = x2q @ wq[gi].T
            if L.get("b") is not None:
                p = p + L["b"][gi]
            bnc[n] = None
            if L.get("bn") is not None:
                y, bnc[n] = batchnorm_fwd(p.reshape(T * R, H), L["bn"][gi], training)
                p = y.reshape(T, R, H)
            pre[n] = p
        ht = np.zeros((R, H), dtype=x.dtype)
        ct = np.zeros((R, H), dtype=x.dtype)
        st = {k: np.zeros((T, R, H), dtype=x.dtype) for k in ("h", "c", "g0", "g1", "g2", "a")}
        for k in range(T):
            hq = Q(ht)
            if cell == "lstm":
                ft = sigmoid(pre["f"][k] + hq @ uq["f"].T)
                it = sigmoid(pre["i"][k] + hq @ uq["i"].T)
                ot = sigmoid(pre["o"][k] + hq @ uq["o"].T)
                at = pre["c"][k] + hq @ uq["c"].T
                ct = it * act_fwd(L["act"], at) * mask + ft * ct
                ht = ot * act_fwd(L["act"], ct)
                st["g0"][k], st["g1"][k], st["g2"][k], st["c"][k] = ft, it, ot, ct
            else:
                zt = sigmoid(pre["z"][k] + hq @ uq["z"].T)
                if cell == "gru":
                    rt = sigmoid(pre["r"][k] + hq @ uq["r"].T)
                    st["g1"][k] = rt
                    at = pre["h"][k] + Q(rt * ht) @ uq["h"].T
                else:
                    at = pre["h"][k] + Q(zt * ht) @ uq["h"].T
                hcand = act_fwd(L["act"], at) * mask
                ht = zt * ht + (1 - zt) * hcand
                st["g0"][k] = zt
            st["h"][k], st["a"][k] = ht, at
        hs = st["h"]
        out = np.concatenate([hs[:, :B], flip0(hs[:, B:])], axis=2) if bidir else hs
        caches.append(dict(x2=x2q, st=st, mask=mask, bn=bnc, B=B, wq=wq, uq=uq, Q=Q))
        x = out
    return x, caches


def cell_backward(dout, layers, caches, *, cell, bidir=True):
    """Hand-derived BPTT of cell_forward.  grads[i] = dict(w=[..], u=[..], b=[..] or None, bn_weight=[..],
    bn_bias=[..]) in GATES[cell] order."""
    names = GATES[cell]
    grads = [None] * len(layers)
    for li in reversed(range(len(layers))):
        L, c = layers[li], caches[li]
        st, mask, B, Q, uq = c["st"], c["mask"], c["B"], c["Q"], c["uq"]
        hs = st["h"]
        T, R, H = hs.shape
        dH = np.concatenate([dout[:, :, :H], flip0(dout[:, :, H:])], axis=1) if bidir else dout
        zero = np.zeros((R, H), dtype=dout.dtype)
        carry, ccarry = zero.copy(), zero.copy()
        dpre = {n: np.zeros_like(hs) for n in names}
        du = {n: np.zeros_like(L["u"][0]) for n in names}
        for k in reversed(range(T)):
            hp = hs[k - 1] if k > 0 else zero
            at = st["a"][k]
            dh = dH[k] + carry
            if cell == "lstm":
                ft, it, ot, ct = st["g0"][k], st["g1"][k], st["g2"][k], st["c"][k]
                cp = st["c"][k - 1] if k > 0 else zero
                ac = act_fwd(L["act"], ct)
                ya = act_fwd(L["act"], at)
                dc = dh * ot * act_bwd(L["act"], ct, ac) + ccarry
                d = dict(f=dc * cp * ft * (1 - ft), i=dc * ya * mask * it * (1 - it), o=dh * ac * ot * (1 - ot),
                         c=dc * it * mask * act_bwd(L["act"], at, ya))
                ccarry = dc * ft
                carry = sum(d[n] @ uq[n] for n in names)
                for n in names:
                    du[n] += d[n].T @ Q(hp)
            else:
                zt = st["g0"][k]
                y = act_fwd(L["act"], at)
                hcand = y * mask
                da = dh * (1 - zt) * mask * act_bwd(L["act"], at, y)
                v = da @ uq["h"]
                dzt = dh * (hp - hcand)
                if cell == "gru":
                    rt = st["g1"][k]
                    drp = v * hp * rt * (1 - rt)
                    dzp = dzt * zt * (1 - zt)
                    carry = dh * zt + v * rt + dzp @ uq["z"] + drp @ uq["r"]
                    d = dict(h=da, z=dzp, r=drp)
                    du["h"] += da.T @ Q(rt * hp)
                    du["r"] += drp.T @ Q(hp)
                else:
                    dzp = (dzt + v * hp) * zt * (1 - zt)
                    carry = dh * zt + v * zt + dzp @ uq["z"]
                    d = dict(h=da, z=dzp)
                    du["h"] += da.T @ Q(zt * hp)
                du["z"] += dzp.T @ Q(hp)
            for n in names:
                dpre[n][k] = d[n]
        g = dict(w=[], u=[du[n] for n in names], b=None, bn_weight=None, bn_bias=None)
        if L.get("bn") is not None:
            g["bn_weight"], g["bn_bias"] = [], []
        if L.get("b") is not None:
            g["b"] = []
        x2f = c["x2"].reshape(T * R, -1)
        dx2 = 0.0
        for gi, n in enumerate(names):
            dp = dpre[n].reshape(T * R, H)
            if c["bn"][n] is not None:
                dp, gw, gb = batchnorm_bwd(dp, c["bn"][n])
                g["bn_weight"].append(gw)
                g["bn_bias"].append(gb)
            if L.get("b") is not None:
                g["b"].append(dp.sum(0))
            g["w"].append(dp.T @ x2f)
            dx2 = dx2 + dp @ c["wq"][gi]
        dx2 = dx2.reshape(T, R, -1)
        dout = dx2[:, :B] + flip0(dx2[:, B:]) if bidir else dx2
        grads[li] = g
    return dout, grads


# --------------------------------------------------------------------------------------
# MLP  (neural_networks.py:60-150) — also the senone head (dnn_act = softmax)
# --------------------------------------------------------------------------------------


def log_softmax(x):
    """act_fun("softmax") = nn.LogSoftmax(dim=1), neural_networks.py:53-54."""
    m = x.max(axis=1, keepdims=True)
    e = np.exp(x - m)
    return x - m - np.log(e.sum(axis=1, keepdims=True))


def mlp_forward(x, layers, *, training=True, drop_masks=None, quant=False):
    """x [N,D].  layers: dicts(w [O,I], b [O], bn (dict|None), ln (dict(gamma,beta)|None), act, drop).
    Order per layer: drop(act(bn(ln(w x + b))))  (neural_networks.py:138-148).
    drop_masks: per-layer keep masks (nn.Dropout inverted scaling applied here) or None."""
    caches = []
    for li, L in enumerate(layers):
        Q = q16 if quant else (lambda a: a)
        xq, wq = Q(x), Q(L["w"])
        lin = xq @ wq.T + L["b"]
        pre = lin
        ln_cache = bn_cache = None
        if L.get("ln") is not None:
            pre, ln_cache = layernorm_fwd(pre, L["ln"]["gamma"], L["ln"]["beta"])
        if L.get("bn") is not None:
            pre, bn_cache = batchnorm_fwd(pre, L["bn"], training)
        if L["act"] == "softmax":
            y = log_softmax(pre)
        else:
            y = act_fwd(L["act"], pre)
        keep = None
        if training and L.get("drop", 0.0) > 0 and drop_masks is not None and drop_masks[li] is not None:
            keep = drop_masks[li].astype(x.dtype) / (1.0 - L["drop"])
            out = y * keep
        else:
            out = y
        caches.append(dict(x=xq, wq=wq, pre=pre, y=y, keep=keep, ln=ln_cache, bn=bn_cache))
        x = out
    return x, caches


def mlp_backward(dout, layers, caches):
    grads = [None] * len(layers)
    for li in reversed(range(len(layers))):
        L, c = layers[li], caches[li]
        d = dout if c["keep"] is None else dout * c["keep"]
        if L["act"] == "softmax":
            d = d - np.exp(c["y"]) * d.sum(axis=1, keepdims=True)
        else:
            d = d * act_bwd(L["act"], c["pre"], c["y"])
        g = {}
        if c["bn"] is not None:
            d, g["bn_weight"], g["bn_bias"] = batchnorm_bwd(d, c["bn"])
        if c["ln"] is not None:
            d, g["ln_gamma"], g["ln_beta"] = layernorm_bwd(d, c["ln"])
        g["w"] = d.T @ c["x"]
        g["b"] = d.sum(0)
        dout = d @ c["wq"]
        grads[li] = g
    return dout, grads


# --------------------------------------------------------------------------------------
# CNN :1464-1556 / SincNet :1559-1665 / SincConv :1668-1813 — conv front-ends (non-sequential modules)
# --------------------------------------------------------------------------------------


def sinc_filters(low_hz_, band_hz_, k, sample_rate=16000, min_low_hz=50, min_band_hz=50):
    """SincConv.forward filter synthesis (:1777-1803).  low_hz_/band_hz_ [C,1] (normalised by the sample rate,
    :1743-1750).  Returns (filters [C,k], cache)."""
    n = (k - 1) / 2
    n_ = (np.arange(-n, n + 1) / sample_rate).reshape(1, -1)                      # :1759-1760
    n_lin = np.linspace(0, k, k)
    window = 0.54 - 0.46 * np.cos(2 * np.pi * n_lin / k)                          # :1755-1756
    low = min_low_hz / sample_rate + np.abs(low_hz_)                              # :1789
    high = low + min_band_hz / sample_rate + np.abs(band_hz_)                     # :1790
    a = 2 * np.pi * n_ * sample_rate                                              # sinc argument = f * a
    half = int((k - 1) / 2)

    def lowpass(f):  # 2 f sinc(f a) with the reference's mirrored evaluation (:1762-1770)
        xl = (f * a)[:, :half]
        yl = np.sin(xl) / xl
        s = np.concatenate([yl, np.ones((f.shape[0], 1)), yl[:, ::-1]], axis=1)
        return 2 * f * s

    bp = lowpass(high) - lowpass(low)                                             # :1798
    m = bp.max(axis=1, keepdims=True)                                             # :1799
    filt = bp / m * window                                                        # :1800-1803
    return filt, dict(low=low, high=high, a=a, bp=bp, m=m, window=window, low_hz_=low_hz_, band_hz_=band_hz_,
                      half=half)


def sinc_filters_bwd(dfilt, c):
    """d(filters [C,k]) -> (d low_hz_, d band_hz_) [C,1]."""
    bp, m, win, a, half = c["bp"], c["m"], c["window"], c["a"], c["half"]
    k = bp.shape[1]
    g = dfilt * win
    dbp = g / m
    jstar = bp.argmax(axis=1)
    corr = (g * bp).sum(axis=1) / (m[:, 0] ** 2)
    dbp[np.arange(bp.shape[0]), jstar] -= corr

    def dlowpass(f):  # d/df of 2 f sinc(f a): 2 cos(f a) off-centre (mirrored like the forward), 2 at the centre
        dl = 2 * np.cos((f * a)[:, :half])
        return np.concatenate([dl, 2 * np.ones((f.shape[0], 1)), dl[:, ::-1]], axis=1)

    dhigh = (dbp * dlowpass(c["high"])).sum(axis=1, keepdims=True)
    dlow = -(dbp * dlowpass(c["low"])).sum(axis=1, keepdims=True) + dhigh
    return np.sign(c["low_hz_"]) * dlow, np.sign(c["band_hz_"]) * dhigh


def conv1d_valid(x, w, b=None):
    """F.conv1d, stride 1, no padding.  x [N,Ci,L], w [Co,Ci,k] -> [N,Co,L-k+1]."""
    k = w.shape[2]
    win = np.lib.stride_tricks.sliding_window_view(x, k, axis=2)  # [N,Ci,Lo,k]
    y = np.einsum("nilk,oik->nol", win, w, optimize=True)
    return y if b is None else y + b[None, :, None]


def convnet_forward(x, layers, *, ln0=None, training=True, keeps=None, quant=False):
    """CNN.forward :1530-1556 / SincNet.forward :1638-1665 on x [N, L0].

    layers: dicts with  kind "sinc" (low_hz_, band_hz_, k, sample_rate, min_low_hz, min_band_hz) or "conv" (w [Co,Ci,k],
    b [Co]);  pool (max_pool1d length);  ln = dict(gamma [C,Lp], beta) or None (the reference's LayerNorm over the
    LAST axis with a [C,Lp] affine, :1505-1507);  act;  drop (nn.Dropout p, inverted scaling).
    ln0: dict(gamma [L0], beta) input LayerNorm (:1541-1542) or None.  keeps: per-layer 0/1 keep masks [N,C,Lp].
    quant: round conv operands to fp16 where the CUDA path does.  Returns ([N, C*Lp], caches)."""
    Q = q16 if quant else (lambda a_: a_)
    c0 = None
    if ln0 is not None:
        x, c0 = layernorm_fwd(x, ln0["gamma"], ln0["beta"])
    h = x[:, None, :]
    caches = []
    for li, L in enumerate(layers):
        fc = None
        if L["kind"] == "sinc":
            filt, fc = sinc_filters(L["low_hz_"], L["band_hz_"], L["k"], L.get("sample_rate", 16000),
                                    L.get("min_low_hz", 50), L.get("min_band_hz", 50))
            w, b = filt[:, None, :], None
        else:
            w, b = L["w"], L["b"]
        hq, wq = Q(h), Q(w)
        y = conv1d_valid(hq, wq, b)
        p = L["pool"]
        Lp = y.shape[2] // p
        yw = y[:, :, :Lp * p].reshape(y.shape[0], y.shape[1], Lp, p)
        arg = yw.argmax(axis=3)
        pooled = np.take_along_axis(yw, arg[..., None], axis=3)[..., 0]
        lnc = None
        pre = pooled
        if L.get("ln") is not None:
            pre, lnc = layernorm_fwd(pooled, L["ln"]["gamma"], L["ln"]["beta"])
        act = act_fwd(L["act"], pre)
        keep = None
        out = act
        if training and L.get("drop", 0.0) > 0 and keeps is not None and keeps[li] is not None:
            keep = keeps[li].astype(x.dtype) / (1.0 - L["drop"])
            out = act * keep
        caches.append(dict(hq=hq, wq=wq, arg=arg, ylen=y.shape[2], pre=pre, act=act, keep=keep, ln=lnc, fc=fc, p=p))
        h = out
    return h.reshape(h.shape[0], -1), dict(layers=caches, ln0=c0)


def convnet_backward(dout, layers, caches):
    """Returns (dx [N,L0] w.r.t. the (normalised) input, grads list, ln0 grads or None).  grads[i]: dict(w, b) or
    dict(low_hz_, band_hz_), plus ln_gamma / ln_beta [C,Lp]."""
    cs = caches["layers"]
    last = cs[-1]
    N = dout.shape[0]
    d = dout.reshape(N, last["act"].shape[1], last["act"].shape[2])
    grads = [None] * len(layers)
    for li in reversed(range(len(layers))):
        L, c = layers[li], cs[li]
        g = {}
        if c["keep"] is not None:
            d = d * c["keep"]
        d = d * act_bwd(L["act"], c["pre"], c["act"])
        if c["ln"] is not None:
            xc, std, gamma, eps = c["ln"]
            g["ln_gamma"] = (d * xc / (std + eps)).sum(0)
            g["ln_beta"] = d.sum(0)
            d, _, _ = layernorm_bwd(d, c["ln"])
        # max_pool1d backward: route to the arg-max position of each window
        Co, Lp, p = d.shape[1], d.shape[2], c["p"]
        dy = np.zeros((N, Co, c["ylen"]), dtype=d.dtype)
        dyw = np.zeros((N, Co, Lp, p), dtype=d.dtype)
        np.put_along_axis(dyw, c["arg"][..., None], d[..., None], axis=3)
        dy[:, :, :Lp * p] = dyw.reshape(N, Co, Lp * p)
        k = c["wq"].shape[2]
        win = np.lib.stride_tricks.sliding_window_view(c["hq"], k, axis=2)  # [N,Ci,Lo,k]
        dw = np.einsum("nol,nilk->oik", dy, win, optimize=True)
        if L["kind"] == "sinc":
            g["low_hz_"], g["band_hz_"] = sinc_filters_bwd(dw[:, 0, :], c["fc"])
        else:
            g["w"], g["b"] = dw, dy.sum(axis=(0, 2))
        # dx[n,i,l+j] += dy[n,o,l] w[o,i,j]
        dh = np.zeros_like(c["hq"])
        for j in range(k):
            dh[:, :, j:j + dy.shape[2]] += np.einsum("nol,oi->nil", dy, c["wq"][:, :, j], optimize=True)
        d = dh
        grads[li] = g
    dx = d[:, 0, :]
    g0 = None
    if caches["ln0"] is not None:
        xc, std, gamma, eps = caches["ln0"]
        g0 = dict(gamma=(dx * xc / (std + eps)).sum(0), beta=dx.sum(0))
        dx, _, _ = layernorm_bwd(dx, caches["ln0"])
    return dx, grads, g0


# --------------------------------------------------------------------------------------
# cost ops of utils.forward_model
# --------------------------------------------------------------------------------------


def nll_loss(logp, labels):
    """nn.NLLLoss() mean over ALL rows incl. padding (utils.py:2087, :2344-2361)."""
    n = logp.shape[0]
    return -logp[np.arange(n), labels].mean()


def nll_loss_bwd(logp, labels, gout=1.0):
    d = np.zeros_like(logp)
    n = logp.shape[0]
    d[np.arange(n), labels] = -gout / n
    return d


def cost_err(logp, labels):
    """utils.py:2379-2380: mean(argmax != lab); argmax = first maximum."""
    return float((np.argmax(logp, axis=1) != labels).mean())


# --------------------------------------------------------------------------------------
# optimizers (torch.optim semantics as configured by utils.optimizer_init, utils.py:2106-2164)
# --------------------------------------------------------------------------------------


def rmsprop_step(p, g, v, lr=0.0004, alpha=0.95, eps=1e-8):
    v = alpha * v + (1 - alpha) * g * g
    p = p - lr * g / (np.sqrt(v) + eps)
    return p, v


def adam_step(p, g, m, v, step, lr=0.001, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
    """torch.optim.Adam as built by utils.optimizer_init (utils.py:2131-2145), amsgrad off.  step counts from 1."""
    g = g + weight_decay * p
    m = betas[0] * m + (1 - betas[0]) * g
    v = betas[1] * v + (1 - betas[1]) * g * g
    bc1, bc2 = 1 - betas[0] ** step, 1 - betas[1] ** step
    p = p - (lr / bc1) * m / (np.sqrt(v) / np.sqrt(bc2) + eps)
    return p, m, v


def sgd_step(p, g, lr=0.08):
    return p - lr * g


# --------------------------------------------------------------------------------------
# whole training step of the headline recipe: liGRU stack -> softmax head(s) -> NLL (+ err)
# (cfg/TIMIT_baselines/TIMIT_liGRU_fmllr.cfg [model]; core.py:616-642)
# --------------------------------------------------------------------------------------


def ligru_model_step(x, labels, ligru_layers, heads, *, masks, bidir=True, loss_weights=None, quant=False,
                     cell="ligru"):
    """x [T,B,D], labels: list of [T*B] int arrays (one per head, t-major rows utils.py:2323).
    heads: list of single-layer softmax MLP layer dicts.  Returns dict(loss, losses, err, logp, grads)."""
    generic = cell in GATES
    if generic:
        out, caches = cell_forward(x, ligru_layers, cell=cell, bidir=bidir, training=True, masks=masks, quant=quant)
    else:
        out, caches = ligru_forward(x, ligru_layers, bidir=bidir, training=True, masks=masks, quant=quant, cell=cell)
    T, B, F = out.shape
    flat = out.reshape(T * B, F)
    dflat = np.zeros_like(flat)
    losses, logps, hgrads = [], [], []
    lw = loss_weights or [1.0] * len(heads)
    for hd, lab, w in zip(heads, labels, lw):
        logp, hc = mlp_forward(flat, [hd], training=True, quant=quant)
        losses.append(nll_loss(logp, lab))
        logps.append(logp)
        dlogp = nll_loss_bwd(logp, lab, w)
        dx, hg = mlp_backward(dlogp, [hd], hc)
        dflat += dx
        hgrads.append(hg[0])
    if generic:
        _, lgrads = cell_backward(dflat.reshape(T, B, F), ligru_layers, caches, cell=cell, bidir=bidir)
    else:
        _, lgrads = ligru_backward(dflat.reshape(T, B, F), ligru_layers, caches, bidir=bidir)
    loss = sum(w * l for w, l in zip(lw, losses))
    return dict(loss=loss, losses=losses, err=cost_err(logps[0], labels[0]), logp=logps, out=out, ligru_grads=lgrads,
                head_grads=hgrads)


# --------------------------------------------------------------------------------------
# input side of the path: data_io.load_chunk array work and core.run_nn batch assembly
# --------------------------------------------------------------------------------------


def context_window(fea, left, right):
    """data_io.py:228-241 (np.roll per lag, then drop the wrapped rows)."""
    n, f = fea.shape
    out = np.empty((n, f * (left + right + 1)))
    for j, lag in enumerate(range(-left, right + 1)):
        out[:, j * f:(j + 1) * f] = np.roll(fea, -lag, axis=0)
    return out[left:n - right]


def prepare_chunk(fea, lab, left, right):
    """data_io.py:255-272: context window, (x - mean) / std per column (float64, population std), label column."""
    # literal dtype behaviour: context_window allocates float64 (np.empty default); without a context window the
    # float32 features are normalised in float32
    data = context_window(fea, left, right) if (left or right) else fea
    data = (data - np.mean(data, axis=0)) / np.std(data, axis=0)
    if lab is None:
        return data
    lab = lab - lab.min()
    lab = lab[left:-right] if right > 0 else lab[left:]
    return np.column_stack((data, lab))


def assemble_batch(data_set, data_end_index, snt_index, beg_snt, batch_size, randint):
    """core.py:581-598, literally: zero tensor [max_len, B, D], every sentence copied behind a random number of
    leading zero frames drawn with randint(0, N_zeros)."""
    arr_len = [int(data_end_index[0])] + [int(data_end_index[i] - data_end_index[i - 1]) for i in range(1, len(data_end_index))]
    max_len = int(max(arr_len[snt_index:snt_index + batch_size]))
    inp = np.zeros((max_len, batch_size, data_set.shape[1]), dtype=data_set.dtype)
    for k in range(batch_size):
        snt_len = int(data_end_index[snt_index]) - beg_snt
        n_left = randint(0, max_len - snt_len)
        inp[n_left:n_left + snt_len, k, :] = data_set[beg_snt:beg_snt + snt_len, :]
        beg_snt = int(data_end_index[snt_index])
        snt_index += 1
    return inp, snt_index, beg_snt


def posterior_ark_bytes(key, logp, counts=None):
    """core.py:660-671 + data_io.write_mat (:1200-1239): optional `out - log(counts/sum(counts))`, then the binary
    Kaldi matrix entry (key, "\0B", "FM "/"DM ", \4 rows, \4 cols, payload)."""
    import struct
    out = logp
    if counts is not None:
        out = out - np.log(counts / np.sum(counts))
    tag = {"float32": b"FM ", "float64": b"DM "}[str(out.dtype)]
    head = ((key + " ").encode("latin1") if key != "" else b"") + b"\0B" + tag
    return head + b"\x04" + struct.pack("<I", out.shape[0]) + b"\x04" + struct.pack("<I", out.shape[1]) + out.tobytes()


def cm_decode(col_headers, data, globmin, globrange):
    """Kaldi CompressedMatrix payload -> float32 [rows, cols] (data_io.py:1150-1196): col_headers [cols,4] uint16
    percentiles, data [cols,rows] uint8 column-major; three linear segments per column (0..64, 65..192, 193..255)."""
    globmin, globrange = np.float32(globmin), np.float32(globrange)
    p = (col_headers.astype(np.uint16) * globrange * 1.52590218966964e-05 + globmin).astype(np.float32)
    p0, p25, p75, p100 = (p[:, k].reshape(-1, 1) for k in range(4))
    lo, hi = data <= 64, data > 192
    mid = ~(lo | hi)
    mat = np.zeros(data.shape, dtype=np.float32)
    mat += (p0 + (p25 - p0) / 64.0 * data) * lo.astype(np.float32)
    mat += (p25 + (p75 - p25) / 128.0 * (data - 64)) * mid.astype(np.float32)
    mat += (p75 + (p100 - p75) / 63.0 * (data - 192)) * hi.astype(np.float32)
    return mat.T
