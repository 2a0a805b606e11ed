"""Helpers shared by the parity tests: load a golden fixture (written by tests/golden/make_golden.py
from the unmodified reference) and turn it into oracle-style layer dicts."""
import ast
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    d = {k: z[k] for k in z.files}
    d["meta"] = ast.literal_eval(str(d["meta"]))
    return d


def bn_dict(d, prefix, stage="init.", dtype=np.float64):
    return dict(
        weight=d[f"{stage}{prefix}.weight"].astype(dtype),
        bias=d[f"{stage}{prefix}.bias"].astype(dtype),
        running_mean=d[f"{stage}{prefix}.running_mean"].astype(dtype),
        running_var=d[f"{stage}{prefix}.running_var"].astype(dtype),
        num_batches_tracked=int(d[f"{stage}{prefix}.num_batches_tracked"]),
        eps=1e-5,
        momentum=0.05,
    )


def ligru_layers(d, dtype=np.float64, stage="init."):
    m = d["meta"]
    layers = []
    rnn = m.get("cell", "ligru") == "rnn"  # single-gate cell: no wz / uz / bn_wz
    for i in range(len(m["lay"])):
        L = dict(
            wh=d[f"{stage}ligru.wh.{i}.weight"].astype(dtype),
            uh=d[f"{stage}ligru.uh.{i}.weight"].astype(dtype),
            act=m["act"],
            drop=m["drop"],
        )
        if not rnn:
            L["wz"] = d[f"{stage}ligru.wz.{i}.weight"].astype(dtype)
            L["uz"] = d[f"{stage}ligru.uz.{i}.weight"].astype(dtype)
        if m["bn"]:
            L["bn_wh"] = bn_dict(d, f"ligru.bn_wh.{i}", "init.", dtype)
            L["bn_wz"] = None if rnn else bn_dict(d, f"ligru.bn_wz.{i}", "init.", dtype)
            L["bh"] = L["bz"] = None
        else:
            L["bn_wh"] = L["bn_wz"] = None
            L["bh"] = d[f"{stage}ligru.wh.{i}.bias"].astype(dtype)
            L["bz"] = None if rnn else d[f"{stage}ligru.wz.{i}.bias"].astype(dtype)
        layers.append(L)
    return layers


# reference ModuleList names per cell, registration order (= oracle GATES order)
CELL_LISTS = {"lstm": (("wfx", "wix", "wox", "wcx"), ("ufh", "uih", "uoh", "uch")),
              "gru": (("wh", "wz", "wr"), ("uh", "uz", "ur")),
              "minimalgru": (("wh", "wz"), ("uh", "uz"))}


def cell_layers(d, dtype=np.float64, stage="init."):
    """Layer dicts for oracle.cell_forward (LSTM / GRU / minimalGRU fixtures, module prefix `net.`)."""
    m = d["meta"]
    wn, un = CELL_LISTS[m["cell"]]
    layers = []
    for i in range(len(m["lay"])):
        L = dict(w=[d[f"{stage}net.{w}.{i}.weight"].astype(dtype) for w in wn],
                 u=[d[f"{stage}net.{u}.{i}.weight"].astype(dtype) for u in un], act=m["act"], drop=m["drop"],
                 b=None, bn=None)
        if m["bn"]:
            L["bn"] = [bn_dict(d, f"net.bn_{w}.{i}", "init.", dtype) for w in wn]
        else:
            L["b"] = [d[f"{stage}net.{w}.{i}.bias"].astype(dtype) for w in wn]
        layers.append(L)
    return layers


def conv_layers(d, dtype=np.float64, stage="init."):
    """(layers, ln0) for oracle.convnet_forward from a SincNet / CNN fixture (module prefix `net.`)."""
    m = d["meta"]
    layers = []
    for i in range(len(m["n_filt"])):
        if m["kind"] == "SincNet" and i == 0:
            k = m["len_filt"][0] + (1 - m["len_filt"][0] % 2)  # SincConv forces odd lengths (:1722-1724)
            L = dict(kind="sinc", low_hz_=d[f"{stage}net.conv.0.low_hz_"].astype(dtype),
                     band_hz_=d[f"{stage}net.conv.0.band_hz_"].astype(dtype), k=k)
        else:
            L = dict(kind="conv", w=d[f"{stage}net.conv.{i}.weight"].astype(dtype),
                     b=d[f"{stage}net.conv.{i}.bias"].astype(dtype))
        L.update(pool=m["pool"][i], act=m["act"], drop=m["drop"], ln=None)
        if m["ln"]:
            L["ln"] = dict(gamma=d[f"{stage}net.ln.{i}.gamma"].astype(dtype), beta=d[f"{stage}net.ln.{i}.beta"].astype(dtype))
        layers.append(L)
    ln0 = None
    if m["ln_inp"]:
        ln0 = dict(gamma=d[f"{stage}net.ln0.gamma"].astype(dtype), beta=d[f"{stage}net.ln0.beta"].astype(dtype))
    return layers, ln0


def head_layer(d, prefix="head", dtype=np.float64, stage="init."):
    return dict(w=d[f"{stage}{prefix}.wx.0.weight"].astype(dtype), b=d[f"{stage}{prefix}.wx.0.bias"].astype(dtype),
                bn=None, ln=None, act="softmax", drop=0.0)


def masks(d):
    n = len(d["meta"]["lay"])
    return [d[f"mask{i}"] for i in range(n)]


def mlp_layers(d, dtype=np.float64, stage="init."):
    m = d["meta"]
    layers = []
    for i in range(len(m["lay"])):
        L = dict(w=d[f"{stage}mlp.wx.{i}.weight"].astype(dtype), b=d[f"{stage}mlp.wx.{i}.bias"].astype(dtype),
                 act=m["act"][i], drop=m["drop"][i], bn=None, ln=None)
        if m["bn"][i]:
            L["bn"] = bn_dict(d, f"mlp.bn.{i}", "init.", dtype)
        if m["ln"][i]:
            L["ln"] = dict(gamma=d[f"init.mlp.ln.{i}.gamma"].astype(dtype), beta=d[f"init.mlp.ln.{i}.beta"].astype(dtype))
        layers.append(L)
    return layers


def relerr(a, b, floor=1e-12):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), floor))


def grad_entry(d, key):
    """Returns (kind, payload): full tensor, or (idx, val, sum, sumsq) samples for big ones."""
    if key in d:
        return "full", d[key]
    return "sampled", (d[key + ".idx"], d[key + ".val"], float(d[key + ".sum"]), float(d[key + ".sumsq"]))


def check_tensor(d, key, got, tol):
    kind, ref = grad_entry(d, key)
    got = np.asarray(got, dtype=np.float64)
    if kind == "full":
        assert got.shape == ref.shape, (key, got.shape, ref.shape)
        # floor: gradients that are mathematically zero (a bias in front of BatchNorm) are
        # pure rounding noise in the fp32 reference
        e = relerr(got, ref, floor=1e-5)
        assert e <= tol, f"{key}: rel err {e:.3e} > {tol}"
    else:
        idx, val, s, ss = ref
        scale = max(float(np.max(np.abs(val))), 1e-12)
        e = float(np.max(np.abs(got.reshape(-1)[idx] - val))) / scale
        assert e <= tol, f"{key}: sampled rel err {e:.3e} > {tol}"
        n2 = float((got ** 2).sum())
        assert abs(n2 - ss) <= 10 * tol * max(ss, 1e-30), f"{key}: sumsq {n2} vs {ss}"
