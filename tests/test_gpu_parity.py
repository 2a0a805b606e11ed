"""GPU parity tests (run on the B200 box: pytest -m gpu).

The CUDA path (drop-in modules -> autograd Functions -> C ABI -> sm_100a kernels) is compared with
  * the golden vectors produced by the unmodified reference (tests/golden/*.npz), and
  * the numpy oracle (oracle/pk_oracle.py) on seeded inputs at sizes it finishes in seconds,
and, at BASELINE.json's full size, through size-independent properties.

Tolerances (north star: "within 1e-3 relative in fp32"): tensor-core operands are fp16 with fp32
accumulation, so per-frame log-posteriors / loss are checked at 1e-3 relative; gradients at 5e-3
relative to the tensor's max-abs (they pass through two fp16-operand GEMM chains); the argmax
error rate is an integer path and must match exactly given margins (fixtures use scaled heads).
"""
import numpy as np
import pytest
import torch

import golden_util as gu

pytestmark = pytest.mark.gpu

TOL_FWD = 1e-3
TOL_GRAD = 5e-3
# Gradients through a ReLU-family recurrence are only piecewise smooth: the fp16-operand recurrent GEMM
# perturbs pre-activations by ~1e-3 relative, which flips act'(a) for the ~0.1 % of (unit, step) entries with
# a ~ 0, and each flip changes that entry's gradient by O(1).  Against the fp32 reference such fixtures are
# therefore held to an L2 bound; the exact backward math is pinned by the quantisation-matched oracle tests
# below (same operand rounding -> same side of every kink -> TOL_GRAD).
TOL_GRAD_KINK_L2 = 0.10
KINK_ACTS = ("relu", "leaky_relu")


def rel_l2(got, ref):
    got = np.asarray(got, dtype=np.float64).reshape(-1)
    ref = np.asarray(ref, dtype=np.float64).reshape(-1)
    return float(np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-30))


def _mods():
    import neural_networks as pknn
    return pknn


def ligru_opts(m):
    n = len(m["lay"])
    return {
        "ligru_lay": ",".join(map(str, m["lay"])), "ligru_drop": ",".join([str(m["drop"])] * n),
        "ligru_use_laynorm_inp": "False", "ligru_use_batchnorm_inp": "False",
        "ligru_use_laynorm": ",".join(["False"] * n), "ligru_use_batchnorm": ",".join([str(m["bn"])] * n),
        "ligru_bidir": str(m["bidir"]), "ligru_act": ",".join([m["act"]] * n), "ligru_orthinit": "True",
        "use_cuda": "True", "to_do": "train",
    }


def head_opts(S):
    return {"dnn_lay": str(S), "dnn_drop": "0.0", "dnn_use_laynorm_inp": "False", "dnn_use_batchnorm_inp": "False",
            "dnn_use_batchnorm": "False", "dnn_use_laynorm": "False", "dnn_act": "softmax", "use_cuda": "True",
            "to_do": "train"}


def net_prefix(m):
    return "ligru" if m.get("cell", "ligru") in ("ligru", "rnn") else "net"


FIXTURES = ["ligru_small", "ligru_uni_tanh_nobn", "ligru_ragged", "ligru_h550", "rnn_bidir_bn", "rnn_uni_tanh",
            "lstm_bidir_bn", "lstm_uni_nobn", "gru_bidir_bn", "gru_uni_nobn", "minimalgru_bidir_bn", "minimalgru_uni_nobn"]


def build_from_fixture(d, stage="init."):
    pknn = _mods()
    m = d["meta"]
    cell = m.get("cell", "ligru")
    cls = {"ligru": "liGRU", "rnn": "RNN", "lstm": "LSTM", "gru": "GRU", "minimalgru": "minimalGRU"}[cell]
    net = getattr(pknn, cls)({k.replace("ligru_", cell + "_"): v for k, v in ligru_opts(m).items()}, m["D"])
    head = pknn.MLP(head_opts(m["S"]), net.out_dim)
    head2 = pknn.MLP(head_opts(m["S2"]), net.out_dim) if m["S2"] else None

    def load(mod, prefix):
        sd = {}
        for k in mod.state_dict().keys():
            key = f"{stage}{prefix}.{k}"
            if key not in d:
                key = f"init.{prefix}.{k}"
            sd[k] = torch.from_numpy(np.asarray(d[key]))
        mod.load_state_dict(sd)

    load(net, net_prefix(m))
    load(head, "head")
    if head2 is not None:
        load(head2, "head2")
    mods = [x for x in (net, head, head2) if x is not None]
    for x in mods:
        x.cuda().train()
    masks = gu.masks(d) if stage == "init." else None
    if masks is not None:
        net._mask = lambda i, rows, H, dev: (torch.from_numpy(masks[i]).to(dev), 1.0)
    return net, head, head2


def run_step(d):
    m = d["meta"]
    net, head, head2 = build_from_fixture(d)
    x = torch.from_numpy(d["x"]).cuda()
    lab = torch.from_numpy(d["lab"]).cuda().long()
    h = net(x)
    h.retain_grad()
    flat = h.view(m["T"] * m["B"], -1)
    logp = head(flat)
    loss_cd = torch.nn.functional.nll_loss(logp, lab)
    loss = loss_cd
    logp2 = None
    if head2 is not None:
        logp2 = head2(flat)
        loss = loss_cd + 1.0 * torch.nn.functional.nll_loss(logp2, torch.from_numpy(d["lab2"]).cuda().long())
    err = (logp.max(dim=1)[1] != lab).float().mean()
    loss.backward()
    return net, head, head2, h, logp, logp2, loss, loss_cd, err


@pytest.mark.parametrize("name", FIXTURES)
def test_forward_matches_reference(name):
    d = gu.load(name)
    net, head, head2, h, logp, logp2, loss, loss_cd, err = run_step(d)
    assert gu.relerr(h.detach().cpu().numpy(), d["out"]) < 2 * TOL_FWD
    # per-frame senone log-posteriors within 1e-3 relative
    ref = d["logp"].astype(np.float64)
    got = logp.detach().cpu().numpy().astype(np.float64)
    assert np.max(np.abs(got - ref)) / np.max(np.abs(ref)) < TOL_FWD
    assert abs(loss.item() - float(d["loss"])) / abs(float(d["loss"])) < TOL_FWD
    assert abs(loss_cd.item() - float(d["loss_cd"])) / abs(float(d["loss_cd"])) < TOL_FWD
    if logp2 is not None:
        assert gu.relerr(logp2.detach().cpu().numpy(), d["logp2"]) < TOL_FWD
    # integer path: identical argmax wherever the reference's top-2 margin exceeds the tolerance
    top2 = np.sort(ref, axis=1)[:, -2:]
    safe = (top2[:, 1] - top2[:, 0]) > 2 * TOL_FWD * np.max(np.abs(ref))
    assert safe.mean() > 0.8
    assert np.array_equal(got.argmax(1)[safe], ref.argmax(1)[safe])
    if safe.all():
        assert err.item() == pytest.approx(float(d["err"]), abs=1e-7)


@pytest.mark.parametrize("name", FIXTURES)
def test_gradients_match_reference(name):
    d = gu.load(name)
    net, head, head2, h, *_ = run_step(d)
    assert gu.relerr(h.grad.cpu().numpy(), d["dout"]) < TOL_GRAD
    npfx = net_prefix(d["meta"])
    for prefix, mod in ((npfx, net), ("head", head), ("head2", head2)):
        if mod is None:
            continue
        for k, p in mod.named_parameters():
            key = f"grad.{prefix}.{k}"
            if p.grad is None:
                assert key not in d and key + ".idx" not in d, f"{key}: reference has a gradient, we do not"
                continue
            g = p.grad.detach().cpu().numpy()
            if prefix == npfx and d["meta"]["act"] in KINK_ACTS and max(d["meta"]["lay"]) > 32:
                kind, ref = gu.grad_entry(d, key)
                if kind == "full":
                    assert rel_l2(g, ref) < TOL_GRAD_KINK_L2, key
                else:
                    assert rel_l2(g.reshape(-1)[ref[0]], ref[1]) < TOL_GRAD_KINK_L2, key
            else:
                gu.check_tensor(d, key, g, TOL_GRAD)
    # BatchNorm running statistics (unbiased variance over the T*2B rows the reference normalised)
    if d["meta"]["bn"]:
        sd = net.state_dict()
        for k in sd:
            if "running" in k:
                assert gu.relerr(sd[k].cpu().numpy(), d[f"bnstat.{npfx}.{k}"]) < TOL_FWD, k
            if "num_batches" in k:
                assert int(sd[k]) == int(d[f"bnstat.{npfx}.{k}"])


def test_eval_mode_matches_reference():
    """to_do=valid/forward: scalar (1-p) dropout, BatchNorm running statistics (utils.py:2062-2069)."""
    d = gu.load("ligru_small")
    m = d["meta"]
    net, head, _ = build_from_fixture(d, stage="step1.")
    sd = net.state_dict()
    for k in list(sd):
        if "running" in k or "num_batches" in k:
            sd[k] = torch.from_numpy(np.asarray(d[f"bnstat.ligru.{k}"]))
    net.load_state_dict(sd)
    net.cuda().eval()
    head.eval()
    net.test_flag = True
    with torch.no_grad():
        logp = head(net(torch.from_numpy(d["x"]).cuda()).view(m["T"] * m["B"], -1))
    assert gu.relerr(logp.cpu().numpy(), d["eval_logp"]) < TOL_FWD


def test_fused_head_nll_matches_reference():
    import pk_functions as pkf
    d = gu.load("ligru_small")
    x = torch.from_numpy(d["out"]).cuda().view(-1, d["out"].shape[-1]).requires_grad_(True)
    W = torch.from_numpy(d["init.head.wx.0.weight"]).cuda().requires_grad_(True)
    b = torch.from_numpy(d["init.head.wx.0.bias"]).cuda().requires_grad_(True)
    lab = torch.from_numpy(d["lab"]).cuda().long()
    loss, err, logp = pkf.HeadNLLFn.apply(x, W, b, lab)
    assert abs(loss.item() - float(d["loss_cd"])) / float(d["loss_cd"]) < TOL_FWD
    assert err.item() == pytest.approx(float(d["err"]), abs=1e-7)
    assert gu.relerr(logp.cpu().numpy(), d["logp"]) < TOL_FWD
    loss.backward()
    # reference head gradients come from loss_cd + loss_mono; compare with a torch fp32 head on the same input
    x2 = x.detach().clone().requires_grad_(True)
    W2 = W.detach().clone().requires_grad_(True)
    b2 = b.detach().clone().requires_grad_(True)
    torch.nn.functional.nll_loss(torch.log_softmax(x2 @ W2.t() + b2, dim=1), lab).backward()
    for g, r in ((x.grad, x2.grad), (W.grad, W2.grad), (b.grad, b2.grad)):
        assert gu.relerr(g.cpu().numpy(), r.cpu().numpy()) < TOL_GRAD


@pytest.mark.parametrize("act,quant", [("relu", True), ("tanh", False), ("tanh", True)])
def test_against_oracle_medium(act, quant):
    """Seeded random problem at a size the numpy oracle finishes in seconds: 2 x 550 bidirectional.
    relu: against the oracle with the SAME fp16 operand rounding (pins the backward math exactly);
    tanh (smooth): also against the plain fp32 algorithm."""
    import pk_oracle as orc
    pknn = _mods()
    T, B, D, H, S = 24, 8, 40, 550, 200
    meta = dict(lay=[H, H], drop=0.2, bn=True, bidir=True, act=act, D=D)
    torch.manual_seed(5)
    net = pknn.liGRU(ligru_opts(meta), D)
    head = pknn.MLP(head_opts(S), net.out_dim)
    with torch.no_grad():
        head.wx[0].weight.mul_(20.0)
    g = torch.Generator().manual_seed(6)
    x = torch.randn(T, B, D, generator=g)
    lab = torch.randint(0, S, (T * B,), generator=g)
    masks = [torch.bernoulli(torch.full((2 * B, H), 0.8), generator=g) for _ in range(2)]
    sd = {k: v.detach().numpy().astype(np.float64) for k, v in net.state_dict().items()}
    layers = []
    for i in range(2):
        layers.append(dict(
            wh=sd[f"wh.{i}.weight"], wz=sd[f"wz.{i}.weight"], uh=sd[f"uh.{i}.weight"], uz=sd[f"uz.{i}.weight"],
            bh=None, bz=None, act=act, drop=0.2,
            bn_wh=dict(weight=sd[f"bn_wh.{i}.weight"], bias=sd[f"bn_wh.{i}.bias"], running_mean=np.zeros(H),
                       running_var=np.ones(H), eps=1e-5, momentum=0.05),
            bn_wz=dict(weight=sd[f"bn_wz.{i}.weight"], bias=sd[f"bn_wz.{i}.bias"], running_mean=np.zeros(H),
                       running_var=np.ones(H), eps=1e-5, momentum=0.05)))
    hd = dict(w=head.wx[0].weight.detach().numpy().astype(np.float64),
              b=head.wx[0].bias.detach().numpy().astype(np.float64), bn=None, ln=None, act="softmax", drop=0.0)
    ref = orc.ligru_model_step(x.numpy().astype(np.float64), [lab.numpy()], layers, [hd],
                               masks=[mk.numpy() for mk in masks], bidir=True, quant=quant)
    net.cuda().train()
    head.cuda().train()
    net._mask = lambda i, rows, Hh, dev: (masks[i].to(dev), 1.0)
    out = net(x.cuda())
    logp = head(out.view(T * B, -1))
    loss = torch.nn.functional.nll_loss(logp, lab.cuda())
    loss.backward()
    assert gu.relerr(logp.detach().cpu().numpy(), ref["logp"][0]) < TOL_FWD
    assert abs(loss.item() - ref["loss"]) / ref["loss"] < TOL_FWD
    # relu: accumulation-order differences (~1e-6) can still flip an isolated kink, which moves ONE row of a
    # weight gradient by percents; the L2 metric is robust to that, the max metric gets a wider bound
    tol_max = 0.2 if act in KINK_ACTS else TOL_GRAD  # an isolated flipped kink moves one row by up to ~10 %
    tol_l2 = 2 * TOL_GRAD if act in KINK_ACTS else TOL_GRAD

    def close(got, want, what):
        assert rel_l2(got, want) < tol_l2, what
        assert gu.relerr(got, want) < tol_max, what

    for i in range(2):
        for k in ("wh", "wz", "uh", "uz"):
            close(getattr(net, k)[i].weight.grad.cpu().numpy(), ref["ligru_grads"][i][k], (i, k))
        for gate in ("wh", "wz"):
            bn = getattr(net, "bn_" + gate)[i]
            close(bn.weight.grad.cpu().numpy(), ref["ligru_grads"][i][f"bn_{gate}_weight"], (i, gate, "gamma"))
            close(bn.bias.grad.cpu().numpy(), ref["ligru_grads"][i][f"bn_{gate}_bias"], (i, gate, "beta"))
    close(head.wx[0].weight.grad.cpu().numpy(), ref["head_grads"][0]["w"], "head")


def _bn_dict(sd, key, H):
    return dict(weight=sd[key + ".weight"], bias=sd[key + ".bias"], running_mean=np.zeros(H), running_var=np.ones(H),
                eps=1e-5, momentum=0.05)


@pytest.mark.parametrize("cell,H,act", [("lstm", 550, "tanh"), ("lstm", 200, "relu"), ("ligru", 1024, "relu"),
                                        ("gru", 550, "tanh"), ("gru", 200, "relu"), ("minimalgru", 550, "tanh")])
def test_stepwise_cells_against_oracle_medium(cell, H, act):
    """Step-wise kernels (pk_cell_step.cu): LSTM / GRU / minimalGRU at the recipes' hidden size and liGRU beyond
    the persistent kernels' 560-unit limit (the 5x1024 stress shape), against the oracle with the SAME fp16
    operand rounding."""
    import pk_oracle as orc
    pknn = _mods()
    T, B, D, S = 12, 8, 40, 150
    meta = dict(lay=[H, H], drop=0.2, bn=True, bidir=True, act=act, D=D)
    torch.manual_seed(7)
    opts = {k.replace("ligru_", cell + "_"): v for k, v in ligru_opts(meta).items()}
    cls = {"ligru": "liGRU", "lstm": "LSTM", "gru": "GRU", "minimalgru": "minimalGRU"}[cell]
    net = getattr(pknn, cls)(opts, D)
    head = pknn.MLP(head_opts(S), net.out_dim)
    with torch.no_grad():
        head.wx[0].weight.mul_(20.0)
    g = torch.Generator().manual_seed(8)
    x = torch.randn(T, B, D, generator=g)
    lab = torch.randint(0, S, (T * B,), generator=g)
    masks = [torch.bernoulli(torch.full((2 * B, H), 0.8), generator=g) for _ in range(2)]
    sd = {k: v.detach().numpy().astype(np.float64) for k, v in net.state_dict().items()}
    layers = []
    generic = cell in gu.CELL_LISTS
    if generic:
        wn, un = gu.CELL_LISTS[cell]
        for i in range(2):
            layers.append(dict(w=[sd[f"{w}.{i}.weight"] for w in wn], u=[sd[f"{u}.{i}.weight"] for u in un], b=None,
                               bn=[_bn_dict(sd, f"bn_{w}.{i}", H) for w in wn], act=act, drop=0.2))
    else:
        wn, un = ("wh", "wz"), ("uh", "uz")
        for i in range(2):
            layers.append(dict(wh=sd[f"wh.{i}.weight"], wz=sd[f"wz.{i}.weight"], uh=sd[f"uh.{i}.weight"],
                               uz=sd[f"uz.{i}.weight"], bh=None, bz=None, act=act, drop=0.2,
                               bn_wh=_bn_dict(sd, f"bn_wh.{i}", H), bn_wz=_bn_dict(sd, f"bn_wz.{i}", H)))
    hd = dict(w=head.wx[0].weight.detach().numpy().astype(np.float64),
              b=head.wx[0].bias.detach().numpy().astype(np.float64), bn=None, ln=None, act="softmax", drop=0.0)
    ref = orc.ligru_model_step(x.numpy().astype(np.float64), [lab.numpy()], layers, [hd],
                               masks=[mk.numpy() for mk in masks], bidir=True, quant=True, cell=cell)
    net.cuda().train()
    head.cuda().train()
    net._mask = lambda i, rows, Hh, dev: (masks[i].to(dev), 1.0)
    out = net(x.cuda())
    logp = head(out.view(T * B, -1))
    loss = torch.nn.functional.nll_loss(logp, lab.cuda())
    loss.backward()
    assert gu.relerr(out.detach().cpu().numpy(), ref["out"]) < 2 * TOL_FWD
    assert gu.relerr(logp.detach().cpu().numpy(), ref["logp"][0]) < TOL_FWD
    assert abs(loss.item() - ref["loss"]) / ref["loss"] < TOL_FWD
    # ReLU: isolated kink flips from accumulation-order differences (see test_against_oracle_medium); wider layers
    # and the two-product cells (GRU: fp16 r*h operand) have more of them -> 2 % L2 bound; tanh cases stay at 0.5 %
    tol_max = 0.2 if act in KINK_ACTS else TOL_GRAD
    tol_l2 = 4 * TOL_GRAD if act in KINK_ACTS else TOL_GRAD

    def close(got, want, what):
        assert rel_l2(got, want) < tol_l2, what
        assert gu.relerr(got, want) < tol_max, what

    for i in range(2):
        gr = ref["ligru_grads"][i]
        for gi, (w, u) in enumerate(zip(wn, un)):
            bn = getattr(net, "bn_" + w)[i]
            if generic:
                want = (gr["w"][gi], gr["u"][gi], gr["bn_weight"][gi], gr["bn_bias"][gi])
            else:
                want = (gr[w], gr[u], gr[f"bn_{w}_weight"], gr[f"bn_{w}_bias"])
            close(getattr(net, w)[i].weight.grad.cpu().numpy(), want[0], (i, w))
            close(getattr(net, u)[i].weight.grad.cpu().numpy(), want[1], (i, u))
            close(bn.weight.grad.cpu().numpy(), want[2], (i, w, "gamma"))
            close(bn.bias.grad.cpu().numpy(), want[3], (i, w, "beta"))
    close(head.wx[0].weight.grad.cpu().numpy(), ref["head_grads"][0]["w"], "head")


def test_full_size_time_reversal_property():
    """BASELINE config 2 shape (500 x 32 x 40 -> 550 bidirectional).  Both directions share the weights
    (reference :1095-1097), so in eval mode a time-flipped chunk must give the time-flipped output with the
    direction halves swapped.  One layer: bit for bit (and the forward pass is deterministic).  Five layers: the
    same holds once the next layers' input columns for the two halves are swapped as well (then only the fp32
    accumulation order differs)."""
    import copy
    pknn = _mods()
    T, B, D, H = 500, 32, 40, 550
    x = torch.randn(T, B, D, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
    for nlay in (1, 5):
        meta = dict(lay=[H] * nlay, drop=0.2, bn=True, bidir=True, act="relu", D=D)
        torch.manual_seed(1)
        opts = ligru_opts(meta)
        opts["to_do"] = "valid"
        net = pknn.liGRU(opts, D).cuda().eval()
        net2 = copy.deepcopy(net)
        with torch.no_grad():
            for i in range(1, nlay):
                for w in (net2.wh[i].weight, net2.wz[i].weight):
                    w.copy_(torch.cat([w[:, H:], w[:, :H]], dim=1))
            y1 = net(x)
            y1b = net(x)
            y2 = net2(torch.flip(x, dims=[0]))
        assert y1.shape == (T, B, 2 * H)
        assert torch.isfinite(y1).all()
        assert torch.equal(y1, y1b)
        want = torch.flip(torch.cat([y1[:, :, H:], y1[:, :, :H]], dim=2), dims=[0])
        if nlay == 1:
            assert torch.equal(y2, want)
        else:
            assert (y2 - want).abs().max().item() <= 1e-3 * want.abs().max().item()


def test_full_size_training_step_runs_and_learns():
    """Full-size training steps: loss finite and decreasing on a fixed batch (the reference's own
    'overfit a tiny dataset' advice, README.md:669)."""
    pknn = _mods()
    T, B, D, H, S = 500, 32, 40, 550, 1936
    meta = dict(lay=[H] * 5, drop=0.2, bn=True, bidir=True, act="relu", D=D)
    torch.manual_seed(2)
    net = pknn.liGRU(ligru_opts(meta), D).cuda().train()
    head = pknn.MLP(head_opts(S), net.out_dim).cuda().train()
    net.fast_dropout = True
    opt = [torch.optim.RMSprop(m.parameters(), lr=0.0004, alpha=0.95, eps=1e-8) for m in (net, head)]
    x = torch.randn(T, B, D, device="cuda")
    lab = torch.randint(0, S, (T * B,), device="cuda")
    losses = []
    for _ in range(4):
        for o in opt:
            o.zero_grad()
        loss = torch.nn.functional.nll_loss(head(net(x).view(T * B, -1)), lab)
        loss.backward()
        for o in opt:
            o.step()
        losses.append(loss.item())
    assert all(np.isfinite(losses))
    assert abs(losses[0] - np.log(S)) < 0.5
    assert losses[-1] < losses[0]


def mlp_opts(m):
    return {"dnn_lay": ",".join(map(str, m["lay"])), "dnn_drop": ",".join(map(str, m["drop"])),
            "dnn_use_laynorm_inp": "False", "dnn_use_batchnorm_inp": "False",
            "dnn_use_batchnorm": ",".join(map(str, m["bn"])), "dnn_use_laynorm": ",".join(map(str, m["ln"])),
            "dnn_act": ",".join(m["act"]), "use_cuda": "True", "to_do": "train"}


@pytest.mark.parametrize("name", ["mlp_bn_relu", "mlp_ln_tanh"])
def test_mlp_stack_matches_reference(name):
    """BASELINE configs[0] family (neural_networks.py:60-150), SGD.  mlp_bn_relu: BatchNorm + ReLU + dropout + softmax;
    mlp_ln_tanh: the reference's custom LayerNorm (:23-33) alone (tanh) and under BatchNorm (sigmoid, dropout)."""
    pknn = _mods()
    d = gu.load(name)
    m = d["meta"]
    net = pknn.MLP(mlp_opts(m), m["D"])
    net.load_state_dict({k: torch.from_numpy(np.asarray(d["init.mlp." + k])) for k in net.state_dict()})
    net.cuda().train()
    net._keep_override = [torch.from_numpy(d[f"keep{i}"]) if m["drop"][i] > 0 else None for i in range(len(m["lay"]))]
    x = torch.from_numpy(d["x"]).cuda()
    lab = torch.from_numpy(d["lab"]).cuda().long()
    logp = net(x)
    loss = torch.nn.functional.nll_loss(logp, lab)
    loss.backward()
    assert gu.relerr(logp.detach().cpu().numpy(), d["logp"]) < TOL_FWD
    assert abs(loss.item() - float(d["loss"])) / float(d["loss"]) < TOL_FWD
    for k, p in net.named_parameters():
        key = "grad.mlp." + k
        if p.grad is None:
            assert key not in d
            continue
        g = p.grad.cpu().numpy()
        li = int(k.split(".")[1])
        if k.startswith("wx.") and k.endswith("bias") and m["bn"][li] and not m["ln"][li]:
            assert np.abs(g).max() < 1e-6  # bias directly in front of BatchNorm: zero gradient
        elif k.startswith("ln.") and m["bn"][li]:
            # a per-channel affine (LayerNorm's gamma / beta) directly in front of BatchNorm is removed by it: the
            # gradient is mathematically zero, both sides hold rounding noise
            assert np.abs(g).max() < 1e-4 * max(np.abs(d["grad.mlp.wx.%d.weight" % li]).max(), 1e-6) + 1e-5, k
        else:
            assert rel_l2(g, d[key]) < 2 * TOL_GRAD, k  # ReLU kinks: L2 metric (see TOL_GRAD_KINK_L2)
    sd = net.state_dict()
    for k in sd:
        if "running" in k and ("bnstat.mlp." + k) in d:
            assert gu.relerr(sd[k].cpu().numpy(), d["bnstat.mlp." + k]) < TOL_FWD, k


def test_mlp_unsupported_options_raise_on_gpu():
    """Input normalisation of the MLP (dnn_use_laynorm_inp / dnn_use_batchnorm_inp) is not built: loud, no fallback."""
    pknn = _mods()
    d = gu.load("mlp_ln_tanh")
    for opt in ("dnn_use_laynorm_inp", "dnn_use_batchnorm_inp"):
        o = mlp_opts(d["meta"])
        o[opt] = "True"
        net = pknn.MLP(o, d["meta"]["D"]).cuda()
        with pytest.raises(NotImplementedError):
            net(torch.from_numpy(d["x"]).cuda())


def conv_opts(m):
    prefix = "sinc" if m["kind"] == "SincNet" else "cnn"
    n = len(m["n_filt"])
    j = lambda v: ",".join(map(str, v))
    o = {f"{prefix}_N_filt": j(m["n_filt"]), f"{prefix}_len_filt": j(m["len_filt"]), f"{prefix}_max_pool_len": j(m["pool"]),
         f"{prefix}_use_laynorm_inp": str(m["ln_inp"]), f"{prefix}_use_batchnorm_inp": "False",
         f"{prefix}_use_laynorm": j([m["ln"]] * n), f"{prefix}_use_batchnorm": j([False] * n),
         f"{prefix}_act": j([m["act"]] * n), f"{prefix}_drop": j([m["drop"]] * n), "use_cuda": "True", "to_do": "train"}
    if prefix == "sinc":
        o.update(sinc_sample_rate="16000", sinc_min_low_hz="50", sinc_min_band_hz="50")
    return o


@pytest.mark.parametrize("name", ["sincnet_ln_relu", "sincnet_tanh_noln", "cnn_ln_relu"])
def test_conv_frontends_match_reference(name):
    """SincNet (:1559-1813) / CNN (:1464-1556) + softmax head against the unmodified reference: module output,
    log-posteriors, loss, and every parameter gradient (sinc band edges, conv weights, LayerNorm affines, ln0)."""
    pknn = _mods()
    d = gu.load(name)
    m = d["meta"]
    net = getattr(pknn, m["kind"])(conv_opts(m), m["L0"])
    net.load_state_dict({k: torch.from_numpy(np.asarray(d["init.net." + k])) for k in net.state_dict()})
    head = pknn.MLP(head_opts(m["S"]), net.out_dim)
    head.load_state_dict({k: torch.from_numpy(np.asarray(d["init.head." + k])) for k in head.state_dict()})
    net.cuda().train()
    head.cuda().train()
    if m["drop"] > 0:
        net._keep_override = [torch.from_numpy(d[f"keep{i}"]) for i in range(len(m["n_filt"]))]
    x = torch.from_numpy(d["x"]).cuda()
    lab = torch.from_numpy(d["lab"]).cuda().long()
    h = net(x)
    h.retain_grad()
    logp = head(h)
    loss = torch.nn.functional.nll_loss(logp, lab)
    loss.backward()
    assert gu.relerr(h.detach().cpu().numpy(), d["out"]) < 2 * TOL_FWD
    assert gu.relerr(logp.detach().cpu().numpy(), d["logp"]) < 2 * TOL_FWD
    assert abs(loss.item() - float(d["loss"])) / float(d["loss"]) < TOL_FWD
    assert gu.relerr(h.grad.cpu().numpy(), d["dout"]) < TOL_GRAD
    # max-pool arg-max ties / ReLU kinks can route single entries differently under fp16 operand rounding: L2 bound
    errs = {}
    for k, p in net.named_parameters():
        key = "grad.net." + k
        if p.grad is None:
            assert key not in d, key
            continue
        g = p.grad.cpu().numpy()
        ref = d[key]
        if k.startswith("conv.") and k.endswith("bias") and m["ln"]:
            assert np.abs(g).max() < 1e-4 * max(np.abs(d[key.replace("bias", "weight")]).max(), 1e-30)  # cancels
        elif k.endswith("_hz_"):
            # band edges: d/df of an oscillating 2 cos(f a_j)-weighted sum over the taps of the filter gradient, i.e.
            # a strongly cancelling sum that amplifies the fp16-operand rounding of dF (the kernel's own math is
            # pinned to 1e-4 in test_sinc_filter_kernels_against_oracle)
            assert rel_l2(g, ref) < 0.15, (k, rel_l2(g, ref))
        else:
            errs[k] = rel_l2(g, ref)
    print(name, "gradient rel-L2:", {k: round(v, 5) for k, v in errs.items()})
    # tiny fixtures (5-7 frames) vs the fp32 reference: a single re-routed max-pool / ReLU decision moves a whole
    # filter's gradient, hence the L2 bound; the medium-size oracle test below holds the same math tighter
    bad = {k: v for k, v in errs.items() if v > 0.05}
    assert not bad, bad
    # eval mode after loading the reference's post-step parameters
    net.load_state_dict({k: torch.from_numpy(np.asarray(d["step1.net." + k] if "step1.net." + k in d else d["init.net." + k]))
                         for k in net.state_dict()})
    head.load_state_dict({k: torch.from_numpy(np.asarray(d["step1.head." + k] if "step1.head." + k in d else d["init.head." + k]))
                          for k in head.state_dict()})
    net.cuda().eval()
    head.cuda().eval()
    with torch.no_grad():
        logp_e = head(net(x))
    assert gu.relerr(logp_e.cpu().numpy(), d["eval_logp"]) < 2 * TOL_FWD


def test_sincnet_against_oracle_medium():
    """SincNet at a size closer to the recipe (4 layers, 48/24/24/24 filters, 129-tap sinc layer, 4 frames of 900
    samples — sized so the numpy oracle finishes in seconds) against the oracle with the same fp16 operand rounding."""
    import pk_oracle as orc
    pknn = _mods()
    meta = dict(kind="SincNet", n_filt=[48, 24, 24, 24], len_filt=[129, 5, 5, 3], pool=[3, 3, 3, 2], ln=True, ln_inp=True,
                act="leaky_relu", drop=0.0)
    N, L0, S = 4, 900, 40
    torch.manual_seed(9)
    net = pknn.SincNet(conv_opts(meta), L0)
    head = pknn.MLP(head_opts(S), net.out_dim)
    with torch.no_grad():
        head.wx[0].weight.mul_(10.0)
        for i in range(4):
            net.ln[i].gamma.uniform_(0.5, 1.5)
            net.ln[i].beta.normal_(0, 0.2)
    g = torch.Generator().manual_seed(10)
    x = torch.randn(N, L0, generator=g)
    lab = torch.randint(0, S, (N,), generator=g)
    sd = {k: v.detach().numpy().astype(np.float64) for k, v in net.state_dict().items()}
    layers = [dict(kind="sinc", low_hz_=sd["conv.0.low_hz_"], band_hz_=sd["conv.0.band_hz_"], k=129)]
    for i in range(1, 4):
        layers.append(dict(kind="conv", w=sd[f"conv.{i}.weight"], b=sd[f"conv.{i}.bias"]))
    for i, L in enumerate(layers):
        L.update(pool=meta["pool"][i], act=meta["act"], drop=0.0, ln=dict(gamma=sd[f"ln.{i}.gamma"], beta=sd[f"ln.{i}.beta"]))
    ln0 = dict(gamma=sd["ln0.gamma"], beta=sd["ln0.beta"])
    out_ref, caches = orc.convnet_forward(x.numpy().astype(np.float64), layers, ln0=ln0, training=True, quant=True)
    hd = dict(w=head.wx[0].weight.detach().numpy().astype(np.float64),
              b=head.wx[0].bias.detach().numpy().astype(np.float64), bn=None, ln=None, act="softmax", drop=0.0)
    logp_ref, hc = orc.mlp_forward(out_ref, [hd], training=True, quant=True)
    dx, hg = orc.mlp_backward(orc.nll_loss_bwd(logp_ref, lab.numpy()), [hd], hc)
    _, grads, g0 = orc.convnet_backward(dx, layers, caches)
    net.cuda().train()
    head.cuda().train()
    h = net(x.cuda())
    logp = head(h)
    loss = torch.nn.functional.nll_loss(logp, lab.cuda())
    loss.backward()
    assert gu.relerr(h.detach().cpu().numpy(), out_ref) < 2 * TOL_FWD
    assert gu.relerr(logp.detach().cpu().numpy(), logp_ref) < 2 * TOL_FWD
    errs = {"low_hz_": rel_l2(net.conv[0].low_hz_.grad.cpu().numpy(), grads[0]["low_hz_"]),
            "band_hz_": rel_l2(net.conv[0].band_hz_.grad.cpu().numpy(), grads[0]["band_hz_"]),
            "ln0.gamma": rel_l2(net.ln0.gamma.grad.cpu().numpy(), g0["gamma"]),
            "ln0.beta": rel_l2(net.ln0.beta.grad.cpu().numpy(), g0["beta"])}
    for i in range(1, 4):
        errs[f"conv.{i}.weight"] = rel_l2(net.conv[i].weight.grad.cpu().numpy(), grads[i]["w"])
    for i in range(4):
        errs[f"ln.{i}.gamma"] = rel_l2(net.ln[i].gamma.grad.cpu().numpy(), grads[i]["ln_gamma"])
        errs[f"ln.{i}.beta"] = rel_l2(net.ln[i].beta.grad.cpu().numpy(), grads[i]["ln_beta"])
    print("sincnet medium gradient rel-L2:", {k: round(v, 5) for k, v in errs.items()})
    # the top layer (no routing decisions below it) agrees to ~1e-4; below it single max-pool / leaky-ReLU decisions
    # that flip under accumulation-order differences move whole filters with only 4 frames in the batch: 3 % L2;
    # band edges: cancelling sum over taps (see the fixture test)
    assert errs["ln.3.gamma"] < TOL_GRAD and errs["ln.3.beta"] < TOL_GRAD
    bad = {k: v for k, v in errs.items() if v > (0.05 if k.endswith("_hz_") else 0.03)}
    assert not bad, bad


def test_sinc_filter_kernels_against_oracle():
    """pk_sinc_filters_fwd / _bwd (SincConv :1777-1803 and its chain rule) alone, fp32 against the float64 oracle."""
    import pk_native as pk
    import pk_oracle as orc
    pknn = _mods()
    C, k = 80, 129
    conv = pknn.SincConv(1, C, k)
    low = conv.low_hz_.detach().numpy().astype(np.float64)
    band = conv.band_hz_.detach().numpy().astype(np.float64)
    filt_ref, cache = orc.sinc_filters(low, band, k)
    rng = np.random.default_rng(3)
    dfilt = rng.standard_normal((C, k))
    dlow_ref, dband_ref = orc.sinc_filters_bwd(dfilt, cache)
    lo_d, ba_d = conv.low_hz_.detach().cuda(), conv.band_hz_.detach().cuda()
    filt = torch.empty(C, k, device="cuda")
    pk.sinc_filters_fwd(lo_d, ba_d, C, k, 16000, 50, 50, filt)
    assert gu.relerr(filt.cpu().numpy(), filt_ref) < 2e-5
    dlow = torch.empty(C, 1, device="cuda")
    dband = torch.empty(C, 1, device="cuda")
    pk.sinc_filters_bwd(lo_d, ba_d, C, k, 16000, 50, 50, torch.from_numpy(dfilt).float().cuda(), dlow, dband)
    assert gu.relerr(dlow.cpu().numpy(), dlow_ref) < 1e-4
    assert gu.relerr(dband.cpu().numpy(), dband_ref) < 1e-4
    # the bare module: [N,1,L] -> [N,C,L-k+1] like F.conv1d with the synthesised filters (:1805-1813)
    x = torch.randn(3, 1, 400, generator=torch.Generator().manual_seed(4))
    y = conv.cuda()(x.cuda())
    want = orc.conv1d_valid(orc.q16(x.numpy().astype(np.float64)), orc.q16(filt_ref)[:, None, :])
    assert y.shape == (3, C, 400 - k + 1)
    assert gu.relerr(y.detach().cpu().numpy(), want) < TOL_FWD


def test_fused_optimizer_kernels_against_oracle():
    """pk_rmsprop_step / pk_sgd_step / pk_adam_step on a flat buffer vs the oracle's update rules (which are pinned
    to torch.optim in tests/test_oracle.py), three steps, with the 1/world gradient scale."""
    import pk_native as pk
    import pk_oracle as orc
    rng = np.random.default_rng(1)
    n = 100_003
    p0 = rng.standard_normal(n).astype(np.float32)
    grads = [(rng.standard_normal(n) * s).astype(np.float32) for s in (1.0, 0.05, 2.0)]
    gs = 0.5
    # Adam
    p = torch.from_numpy(p0.copy()).cuda()
    m = torch.zeros(n, device="cuda")
    v = torch.zeros(n, device="cuda")
    pr, mr, vr = p0.astype(np.float64), np.zeros(n), np.zeros(n)
    for k, g in enumerate(grads, 1):
        pk.adam_step(p, torch.from_numpy(g).cuda(), m, v, 0.002, 0.9, 0.98, 1e-7, 0.01, k, gs)
        pr, mr, vr = orc.adam_step(pr, g.astype(np.float64) * gs, mr, vr, k, lr=0.002, betas=(0.9, 0.98), eps=1e-7,
                                   weight_decay=0.01)
    assert np.max(np.abs(p.cpu().numpy() - pr)) < 5e-6   # fp32 parameters (|p| up to ~4.5: ulp 4.8e-7) over 3 steps
    # RMSprop / SGD
    p = torch.from_numpy(p0.copy()).cuda()
    v = torch.zeros(n, device="cuda")
    q = torch.from_numpy(p0.copy()).cuda()
    pr, vr, qr = p0.astype(np.float64), np.zeros(n), p0.astype(np.float64)
    for g in grads:
        gd = torch.from_numpy(g).cuda()
        pk.rmsprop_step(p, gd, v, 0.0004, 0.95, 1e-8, gs)
        pk.sgd_step(q, gd, 0.08, gs)
        pr, vr = orc.rmsprop_step(pr, g.astype(np.float64) * gs, vr, lr=0.0004, alpha=0.95, eps=1e-8)
        qr = orc.sgd_step(qr, g.astype(np.float64) * gs, lr=0.08)
    assert np.max(np.abs(p.cpu().numpy() - pr)) < 5e-6
    assert np.max(np.abs(q.cpu().numpy() - qr)) < 5e-6


@pytest.mark.parametrize("name", ["input_cw", "input_nocw"])
def test_input_side_kernels_match_reference(name):
    """pk_chunk_prepare / pk_batch_assemble against the reference-generated fixtures."""
    import random
    import pk_train
    d = gu.load(name)
    m = d["meta"]
    fea = torch.from_numpy(d["fea"]).cuda()
    lab = torch.from_numpy(d["lab"]).cuda()
    ds = pk_train.prepare_chunk(fea, lab, m["left"], m["right"])
    ref = d["data_set"]
    assert tuple(ds.shape) == ref.shape
    assert np.max(np.abs(ds.cpu().numpy().astype(np.float64) - ref)) < 2e-6 * max(1.0, np.abs(ref).max())
    assert np.array_equal(ds[:, -1].cpu().numpy(), ref[:, -1].astype(np.float32))   # labels: exact
    # assembly from the reference's own (float32) chunk: a pure gather, bit exact
    chunk = torch.from_numpy(ref.astype(np.float32)).cuda()
    rng = random.Random(m["seed"])
    snt, beg = 0, 0
    for i in range(m["n_snt"] // m["batch"]):
        desc, max_len, snt, beg = pk_train.batch_descriptors(d["data_end_index"], snt, beg, m["batch"], rng)
        inp = pk_train.assemble_batch(chunk, desc, max_len)
        assert torch.equal(inp.cpu(), torch.from_numpy(d[f"inp{i}"]))


@pytest.mark.parametrize("M,N,K", [(300, 1500, 200), (1100, 2048, 1100), (129, 1025, 72), (260, 700, 136)])
def test_gemm_tn_tiles_against_fp32(M, N, K):
    """pk_gemm_tn (tcgen05): the 128x256 tile (N >= 1024) and the 128x128 one, ragged edges, bias along N / M,
    BatchNorm row statistics and the amax side output, against fp32 matmul of the same fp16-rounded operands."""
    import pk_native as pk
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    ldk = pk.pad8(K)
    A = torch.zeros(M, ldk, device="cuda", dtype=torch.float16)
    B = torch.zeros(N, ldk, device="cuda", dtype=torch.float16)
    A[:, :K] = torch.randn(M, K, device="cuda", generator=g).half()
    B[:, :K] = torch.randn(N, K, device="cuda", generator=g).half()
    ref = A[:, :K].float() @ B[:, :K].float().t()
    for mode in (0, 1, 2):
        bias = None if mode == 0 else torch.randn(N if mode == 1 else M, device="cuda", generator=g)
        C = torch.full((M, N), float("nan"), device="cuda")
        stats = torch.zeros(M, 2, device="cuda", dtype=torch.float64)
        amax = torch.zeros(1, device="cuda", dtype=torch.int32)
        pk.gemm_tn(A, B, C, M, N, K, lda=ldk, ldb=ldk, ldc=N, bias=bias, bias_mode=mode, rowstats=stats, alpha=0.5,
                   amax_bits=amax)
        want = 0.5 * ref
        if mode == 1:
            want = want + bias[None, :]
        if mode == 2:
            want = want + bias[:, None]
        scale = want.abs().max().item()
        assert torch.isfinite(C).all()
        assert (C - want).abs().max().item() < 2e-5 * scale + 1e-4
        assert (stats[:, 0].float() - want.sum(1)).abs().max().item() < 1e-3 * scale
        assert abs(amax.view(torch.float32).item() - want.abs().max().item()) < 1e-4 * scale


def test_posterior_writer_matches_reference_bytes():
    """pk_train.write_posteriors (device-side prior subtraction + Kaldi ark entry) reproduces the archive the
    reference's core.py:660-671 / data_io.write_mat wrote, byte for byte."""
    import io
    import pk_train
    d = gu.load("post_ark")
    logp = torch.from_numpy(d["logp"]).cuda()
    buf = io.BytesIO()
    pk_train.write_posteriors(buf, "utt_0001", logp, d["counts"])
    pk_train.write_posteriors(buf, "utt_0002", logp[:3])
    assert buf.getvalue() == d["ark"].tobytes()


def test_kaldi_archive_reader_with_device_cm_decode():
    """pk_train.read_mat_ark with device=cuda: FM / DM entries and the CompressedMatrix entry (pk_cm_decode) equal the
    reference reader's output bit for bit."""
    import io
    import pk_train
    d = gu.load("ark_read")
    got = dict(pk_train.read_mat_ark(io.BytesIO(d["feats"].tobytes()), device="cuda"))
    assert list(got) == ["utt_fm", "utt_dm", "utt_cm"]
    for k, m in got.items():
        assert m.is_cuda
        assert np.array_equal(m.cpu().numpy(), d["mat." + k]), k
