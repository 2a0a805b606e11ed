"""Generate golden vectors from the UNMODIFIED reference (mravanelli/pytorch-kaldi).

Run in the build container only (needs /root/reference; the GPU box does not have it):

    python tests/golden/make_golden.py

It imports /root/reference/neural_networks.py, builds the reference modules from cfg-style
option dicts (the same strings utils.model_init hands to the constructors, utils.py:2047-2057),
runs forward / NLLLoss / backward / one optimizer step on seeded synthetic chunks on the CPU in
fp32, and writes small .npz fixtures next to this file.  The oracle (oracle/pk_oracle.py) and the
CUDA path are both checked against these files.
"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("PK_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
import neural_networks as ref_nn  # noqa: E402  (the reference's module zoo)

HERE = os.path.dirname(os.path.abspath(__file__))


class MaskRecorder:
    """Records the CPU-generator dropout masks the reference draws inside forward()
    (torch.bernoulli(...), neural_networks.py:1103-1105)."""

    def __init__(self):
        self.masks = []
        self._orig = torch.bernoulli

    def __enter__(self):
        def rec(*a, **k):
            m = self._orig(*a, **k)
            self.masks.append(m.clone().numpy())
            return m

        torch.bernoulli = rec
        return self

    def __exit__(self, *exc):
        torch.bernoulli = self._orig


def ligru_opts(lay, drop, bn, act, bidir, orth=True):
    n = len(lay)
    return {
        "ligru_lay": ",".join(map(str, lay)),
        "ligru_drop": ",".join([str(drop)] * n),
        "ligru_use_laynorm_inp": "False",
        "ligru_use_batchnorm_inp": "False",
        "ligru_use_laynorm": ",".join(["False"] * n),
        "ligru_use_batchnorm": ",".join([str(bn)] * n),
        "ligru_bidir": str(bidir),
        "ligru_act": ",".join([act] * n),
        "ligru_orthinit": str(orth),
        "use_cuda": "False",
        "to_do": "train",
    }


def mlp_opts(lay, drop, bn, ln, act):
    n = len(lay)
    as_list = lambda v: ",".join(map(str, v if isinstance(v, (list, tuple)) else [v] * n))
    return {
        "dnn_lay": ",".join(map(str, lay)),
        "dnn_drop": as_list(drop),
        "dnn_use_laynorm_inp": "False",
        "dnn_use_batchnorm_inp": "False",
        "dnn_use_batchnorm": as_list(bn),
        "dnn_use_laynorm": as_list(ln),
        "dnn_act": as_list(act),
        "use_cuda": "False",
        "to_do": "train",
    }


def sd_np(module, prefix):
    return {prefix + k: v.detach().numpy().copy() for k, v in module.state_dict().items()}


def grads_np(module, prefix):
    out = {}
    for k, p in module.named_parameters():
        if p.grad is not None:
            out[prefix + k] = p.grad.detach().numpy().copy()
    return out


CELLS = {"ligru": ("liGRU", ("wh", "wz"), "ligru."), "rnn": ("RNN", ("wh",), "ligru."),
         "lstm": ("LSTM", ("wfx", "wix", "wox", "wcx"), "net."), "gru": ("GRU", ("wh", "wz", "wr"), "net."),
         "minimalgru": ("minimalGRU", ("wh", "wz"), "net.")}


def ligru_case(name, *, T, B, D, lay, S, S2, drop, bn, act, bidir, seed, head_scale=20.0, full=True, cell="ligru"):
    torch.manual_seed(seed)
    opts = ligru_opts(lay, drop, bn, act, bidir)
    # same option names with the cell's own prefix (proto/*.proto); W-list names in registration order
    cls_name, gates, net_pfx = CELLS[cell]
    opts = {k.replace("ligru_", cell + "_"): v for k, v in opts.items()}
    net = getattr(ref_nn, cls_name)(opts, D)
    head = ref_nn.MLP(mlp_opts([S], 0.0, False, False, "softmax"), net.out_dim)
    head2 = ref_nn.MLP(mlp_opts([S2], 0.0, False, False, "softmax"), net.out_dim) if S2 else None
    with torch.no_grad():  # give the posteriors real margins
        head.wx[0].weight.mul_(head_scale)
        head.wx[0].bias.normal_(0, 0.1)
        if head2 is not None:
            head2.wx[0].weight.mul_(head_scale)
        for i in range(len(lay)):
            for gname in gates:
                if bn:
                    getattr(net, "bn_" + gname)[i].weight.uniform_(0.5, 1.5)
                    getattr(net, "bn_" + gname)[i].bias.normal_(0, 0.2)
                else:
                    getattr(net, gname)[i].bias.normal_(0, 0.2)
    mods = [(net_pfx, net), ("head.", head)] + ([("head2.", head2)] if head2 is not None else [])
    out = {}
    for pfx, m in mods:
        out.update({"init." + k: v for k, v in sd_np(m, pfx).items()})
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(T, B, D, generator=g)
    lab = torch.randint(0, S, (T * B,), generator=g)
    lab2 = torch.randint(0, S2, (T * B,), generator=g) if S2 else None
    for m in (net, head, head2):
        if m is not None:
            m.train()
    opts = [torch.optim.RMSprop(m.parameters(), lr=0.0004, alpha=0.95, eps=1e-8) for _, m in mods]
    torch.manual_seed(seed + 2)
    with MaskRecorder() as rec:
        h = net(x)
    flat = h.view(T * B, -1)
    logp = head(flat)
    loss_cd = torch.nn.NLLLoss()(logp, lab)
    loss = loss_cd
    if head2 is not None:
        logp2 = head2(flat)
        loss_mono = torch.nn.NLLLoss()(logp2, lab2)
        loss = loss_cd + loss_mono * 1.0  # loss_final=sum(loss_cd, mult_constant(loss_mono,1.0))
    pred = torch.max(logp, dim=1)[1]
    err = torch.mean((pred != lab).float())
    h.retain_grad()
    for o in opts:
        o.zero_grad()
    loss.backward()
    out.update(x=x.numpy(), lab=lab.numpy(), out=h.detach().numpy(), logp=logp.detach().numpy(),
               loss=np.float64(loss.item()), loss_cd=np.float64(loss_cd.item()), err=np.float64(err.item()),
               dout=h.grad.numpy())
    if head2 is not None:
        out.update(lab2=lab2.numpy(), logp2=logp2.detach().numpy())
    for i, m in enumerate(rec.masks):
        out[f"mask{i}"] = m
    for pfx, m in mods:
        out.update({"grad." + k: v for k, v in grads_np(m, pfx).items()})
        out.update({"bnstat." + k: v for k, v in sd_np(m, pfx).items() if "running" in k or "num_batches" in k})
    for o in opts:
        o.step()
    for pfx, m in mods:
        out.update({"step1." + k: v for k, v in sd_np(m, pfx).items() if "running" not in k and "num_batches" not in k})
    # eval-mode forward (to_do=valid: test_flag True -> scalar (1-p) mask, BN running stats)
    net.eval(); head.eval()
    net.test_flag = True
    with torch.no_grad():
        out["eval_logp"] = head(net(x).view(T * B, -1)).numpy()
    meta = dict(T=T, B=B, D=D, lay=lay, S=S, S2=S2 or 0, drop=drop, bn=bn, act=act, bidir=bidir, cell=cell)
    out["meta"] = np.array(repr(meta))
    if not full:  # keep the fixture small: drop the big square matrices' gradients down to samples
        rng = np.random.default_rng(0)
        for k in list(out.keys()):
            v = out[k]
            if isinstance(v, np.ndarray) and v.size > 200_000 and (k.startswith("grad.") or k.startswith("step1.")):
                idx = rng.integers(0, v.size, 4096)
                out[k + ".idx"] = idx
                out[k + ".val"] = v.reshape(-1)[idx]
                out[k + ".sum"] = np.float64(v.astype(np.float64).sum())
                out[k + ".sumsq"] = np.float64((v.astype(np.float64) ** 2).sum())
                del out[k]
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: loss={loss.item():.6f} err={err.item():.4f} -> {path} ({os.path.getsize(path)/1e6:.2f} MB)")


def mlp_case(name, *, N, D, lay, drop, bn, ln, act, seed):
    torch.manual_seed(seed)
    net = ref_nn.MLP(mlp_opts(lay, drop, bn, ln, act), D)
    with torch.no_grad():
        net.wx[-1].weight.mul_(20.0)
        for i in range(len(lay)):
            net.bn[i].weight.uniform_(0.5, 1.5)
            net.bn[i].bias.normal_(0, 0.2)
            net.ln[i].gamma.uniform_(0.5, 1.5)
            net.ln[i].beta.normal_(0, 0.2)
    out = {"init." + k: v for k, v in sd_np(net, "mlp.").items()}
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(N, D, generator=g)
    lab = torch.randint(0, lay[-1], (N,), generator=g)
    net.train()
    keeps = []
    hooks = [m.register_forward_hook(lambda mod, inp, o: keeps.append(((o != 0) | (inp[0] == 0)).numpy().copy()))
             for m in net.drop]
    opt = torch.optim.SGD(net.parameters(), lr=0.08)
    torch.manual_seed(seed + 2)
    logp = net(x)
    for h in hooks:
        h.remove()
    loss = torch.nn.NLLLoss()(logp, lab)
    err = torch.mean((torch.max(logp, dim=1)[1] != lab).float())
    opt.zero_grad()
    loss.backward()
    out.update(x=x.numpy(), lab=lab.numpy(), logp=logp.detach().numpy(), loss=np.float64(loss.item()),
               err=np.float64(err.item()))
    for i, k in enumerate(keeps):
        out[f"keep{i}"] = k
    out.update({"grad." + k: v for k, v in grads_np(net, "mlp.").items()})
    out.update({"bnstat." + k: v for k, v in sd_np(net, "mlp.").items() if "running" in k or "num_batches" in k})
    opt.step()
    out.update({"step1." + k: v for k, v in sd_np(net, "mlp.").items() if "running" not in k and "num_batches" not in k})
    meta = dict(N=N, D=D, lay=lay, drop=drop, bn=bn, ln=ln, act=act)
    out["meta"] = np.array(repr(meta))
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: loss={loss.item():.6f} err={err.item():.4f} -> {path} ({os.path.getsize(path)/1e6:.2f} MB)")


def conv_opts(prefix, n_filt, len_filt, pool, ln, ln_inp, act, drop):
    n = len(n_filt)
    j = lambda v: ",".join(map(str, v))
    o = {f"{prefix}_N_filt": j(n_filt), f"{prefix}_len_filt": j(len_filt), f"{prefix}_max_pool_len": j(pool),
         f"{prefix}_use_laynorm_inp": str(ln_inp), f"{prefix}_use_batchnorm_inp": "False",
         f"{prefix}_use_laynorm": j([ln] * n), f"{prefix}_use_batchnorm": j([False] * n),
         f"{prefix}_act": j([act] * n), f"{prefix}_drop": j([drop] * n), "use_cuda": "False", "to_do": "train"}
    if prefix == "sinc":
        o.update(sinc_sample_rate="16000", sinc_min_low_hz="50", sinc_min_band_hz="50")
    return o


def conv_case(name, *, kind, N, L0, n_filt, len_filt, pool, ln, ln_inp, act, drop, S, seed):
    """CNN :1464-1556 / SincNet :1559-1665 followed by a softmax MLP head, one SGD step."""
    torch.manual_seed(seed)
    prefix = "sinc" if kind == "SincNet" else "cnn"
    net = getattr(ref_nn, kind)(conv_opts(prefix, n_filt, len_filt, pool, ln, ln_inp, act, drop), L0)
    head = ref_nn.MLP(mlp_opts([S], 0.0, False, False, "softmax"), net.out_dim)
    with torch.no_grad():
        head.wx[0].weight.mul_(10.0)
        for i in range(len(n_filt)):
            net.ln[i].gamma.uniform_(0.5, 1.5)
            net.ln[i].beta.normal_(0, 0.2)
        if ln_inp:
            net.ln0.gamma.uniform_(0.5, 1.5)
            net.ln0.beta.normal_(0, 0.2)
    mods = [("net.", net), ("head.", head)]
    out = {}
    for pfx, m in mods:
        out.update({"init." + k: v for k, v in sd_np(m, pfx).items()})
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(N, L0, generator=g) * (1.0 + torch.rand(N, 1, generator=g))
    lab = torch.randint(0, S, (N,), generator=g)
    net.train(); head.train()
    keeps = []
    hooks = [m.register_forward_hook(lambda mod, inp, o: keeps.append(((o != 0) | (inp[0] == 0)).numpy().copy()))
             for m in net.drop]
    opts = [torch.optim.SGD(m.parameters(), lr=0.08) for _, m in mods]
    torch.manual_seed(seed + 2)
    h = net(x)
    for hk in hooks:
        hk.remove()
    h.retain_grad()
    logp = head(h)
    loss = torch.nn.NLLLoss()(logp, lab)
    err = torch.mean((torch.max(logp, dim=1)[1] != lab).float())
    for o in opts:
        o.zero_grad()
    loss.backward()
    out.update(x=x.numpy(), lab=lab.numpy(), out=h.detach().numpy(), logp=logp.detach().numpy(),
               loss=np.float64(loss.item()), err=np.float64(err.item()), dout=h.grad.numpy())
    if kind == "SincNet":
        out["filters"] = net.conv[0].filters.detach().numpy().reshape(n_filt[0], -1)
    for i, k in enumerate(keeps):
        out[f"keep{i}"] = k
    for pfx, m in mods:
        out.update({"grad." + k: v for k, v in grads_np(m, pfx).items()})
    for o in opts:
        o.step()
    for pfx, m in mods:
        out.update({"step1." + k: v for k, v in sd_np(m, pfx).items() if "running" not in k and "num_batches" not in k})
    net.eval(); head.eval()
    with torch.no_grad():
        out["eval_logp"] = head(net(x)).numpy()
    meta = dict(kind=kind, N=N, L0=L0, n_filt=n_filt, len_filt=len_filt, pool=pool, ln=ln, ln_inp=ln_inp, act=act,
                drop=drop, S=S)
    out["meta"] = np.array(repr(meta))
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: loss={loss.item():.6f} err={err.item():.4f} -> {path} ({os.path.getsize(path)/1e6:.2f} MB)")


def input_case(name, *, N, F, left, right, n_snt, batch, seed):
    """Input side: data_io.load_chunk's array work (data_io.py:255-272, using the reference's own
    context_window) and the minibatch assembly loop of core.run_nn (core.py:577-598, restated here line by line
    because it is inline code of a 150-line function) with Python's `random` seeded."""
    import random
    import data_io as ref_io  # noqa: E402
    rng = np.random.default_rng(seed)
    fea = rng.standard_normal((N, F)).astype(np.float32)
    lab = rng.integers(3, 40, N)
    # sentence boundaries (end indices in the un-windowed frame numbering), as load_dataset returns them
    cuts = np.sort(rng.choice(np.arange(10, N - 10), n_snt - 1, replace=False))
    end_index = np.concatenate([cuts, [N]]).astype(np.int64)
    # --- data_io.py:255-272
    data_set = ref_io.context_window(fea, left, right) if (left != 0 or right != 0) else fea
    end_index_fea = end_index - left
    end_index_fea[-1] = end_index_fea[-1] - right
    data_set = (data_set - np.mean(data_set, axis=0)) / np.std(data_set, axis=0)
    data_lab = lab - lab.min()
    data_lab = data_lab[left:-right] if right > 0 else data_lab[left:]
    data_set = np.column_stack((data_set, data_lab))
    # --- core.py:560-598 (seq_model branch)
    random.seed(seed)
    data_end_index = end_index_fea
    t = torch.from_numpy(data_set).float()
    snt_index, beg_snt = 0, 0
    out = dict(fea=fea, lab=lab, data_set=data_set, data_end_index=data_end_index)
    for i in range(n_snt // batch):
        arr_snt_len = np.diff(np.concatenate([[0], data_end_index]))
        max_len = int(max(arr_snt_len[snt_index:snt_index + batch]))
        inp = torch.zeros(max_len, batch, t.shape[1]).contiguous()
        for k in range(batch):
            snt_len = data_end_index[snt_index] - beg_snt
            N_zeros = max_len - snt_len
            N_zeros_left = random.randint(0, N_zeros)
            inp[N_zeros_left:N_zeros_left + snt_len, k, :] = t[beg_snt:beg_snt + snt_len, :]
            beg_snt = data_end_index[snt_index]
            snt_index = snt_index + 1
        out[f"inp{i}"] = inp.numpy()
    out["meta"] = np.array(repr(dict(N=N, F=F, left=left, right=right, n_snt=n_snt, batch=batch, seed=seed)))
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: data_set {data_set.shape} -> {path} ({os.path.getsize(path)/1e6:.2f} MB)")


def ark_case(name, *, rows, cols, seed):
    """Output side: prior normalisation (core.py:664-667) + the reference's own data_io.write_mat."""
    import tempfile
    import data_io as ref_io  # noqa: E402
    rng = np.random.default_rng(seed)
    logp = np.log(rng.dirichlet(np.ones(cols), rows)).astype(np.float32)
    counts = rng.integers(5, 5000, cols).astype(np.float32)
    with tempfile.TemporaryDirectory() as td:
        cf = os.path.join(td, "counts")
        with open(cf, "w") as f:
            f.write("[ " + " ".join(str(int(c)) for c in counts) + " ]\n")
        c2 = ref_io.load_counts(cf)
        out_save = logp - np.log(c2 / np.sum(c2))                     # core.py:666-667
        path = os.path.join(td, "post.ark")
        with open(path, "wb") as fd:
            ref_io.write_mat(td, fd, out_save, "utt_0001")           # core.py:670
            ref_io.write_mat(td, fd, logp[:3], "utt_0002")
        blob = np.frombuffer(open(path, "rb").read(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), logp=logp, counts=counts, ark=blob,
                        meta=np.array(repr(dict(rows=rows, cols=cols))))
    print(f"{name}: {blob.size} ark bytes")


def ark_read_case(name, *, seed):
    """Input side: a binary feature archive (FM, DM and CM entries; the CM bytes are laid out per
    kaldi/src/matrix/compressed-matrix.h) and an alignment archive, decoded by the reference's own readers."""
    import io
    import struct
    import data_io as ref_io  # noqa: E402
    rng = np.random.default_rng(seed)
    fm = rng.standard_normal((9, 6)).astype(np.float32)
    dm = rng.standard_normal((4, 5))
    rows, cols = 37, 13
    gmin, grange = np.float32(-7.5), np.float32(19.25)
    pct = np.sort(rng.integers(0, 65536, (cols, 4)), axis=1).astype(np.uint16)
    data = rng.integers(0, 256, (cols, rows)).astype(np.uint8)
    data[:, :6] = [0, 64, 65, 192, 193, 255]  # segment boundaries
    import tempfile
    alis = {"utt_a": rng.integers(0, 1936, 23).astype(np.int32), "utt_b": rng.integers(0, 1936, 1).astype(np.int32)}
    with tempfile.TemporaryDirectory() as td:
        fp, ap = os.path.join(td, "feats.ark"), os.path.join(td, "ali.ark")
        with open(fp, "wb") as buf:
            ref_io.write_mat(td, buf, fm, "utt_fm")
            ref_io.write_mat(td, buf, dm, "utt_dm")
            buf.write(b"utt_cm \0BCM " + struct.pack("<ffii", gmin, grange, rows, cols) + pct.tobytes() + data.tobytes())
        with open(ap, "wb") as abuf:
            for k, v in alis.items():
                ref_io.write_vec_int(abuf, td, v, key=k)
        feats, ali_bytes = open(fp, "rb").read(), open(ap, "rb").read()
        dec = {k: np.array(m) for k, m in ref_io.read_mat_ark(fp, td)}
        dec_ali = {k: np.array(v) for k, v in ref_io.read_vec_int_ark(ap, td)}
    out = dict(feats=np.frombuffer(feats, dtype=np.uint8), alis=np.frombuffer(ali_bytes, dtype=np.uint8),
               cm_pct=pct, cm_data=data, cm_min=gmin, cm_range=grange, meta=np.array(repr(dict(rows=rows, cols=cols))))
    for k, m in dec.items():
        out["mat." + k] = m
    for k, v in dec_ali.items():
        out["ali." + k] = v
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: {len(feats)} feature bytes, {len(ali_bytes)} alignment bytes, entries {list(dec)} {list(dec_ali)}")


if __name__ == "__main__":
    torch.set_num_threads(4)
    only = sys.argv[1:]  # optional: names of the fixtures to (re)generate
    # A: the headline recipe in miniature (2 heads like cfg/TIMIT_baselines/TIMIT_liGRU_fmllr.cfg)
    if not only or "ligru_small" in only:
      ligru_case("ligru_small", T=12, B=4, D=10, lay=[24, 24], S=30, S2=7, drop=0.2, bn=True, act="relu", bidir=True,
               seed=11)
    # B: no BatchNorm (biases), tanh, unidirectional, odd sizes
    if not only or "ligru_uni_tanh_nobn" in only:
      ligru_case("ligru_uni_tanh_nobn", T=9, B=3, D=7, lay=[20], S=13, S2=0, drop=0.3, bn=False, act="tanh",
               bidir=False, seed=12)
    # C: ragged sizes that do not divide any tile: B not multiple of 8, H not multiple of 8
    if not only or "ligru_ragged" in only:
      ligru_case("ligru_ragged", T=17, B=5, D=13, lay=[37, 29, 37], S=41, S2=0, drop=0.2, bn=True, act="leaky_relu",
               bidir=True, seed=13)
    # D: the real hidden size of config 2 (one layer, short chunk) — sampled gradients only
    if not only or "ligru_h550" in only:
      ligru_case("ligru_h550", T=10, B=8, D=40, lay=[550], S=100, S2=0, drop=0.2, bn=True, act="relu", bidir=True,
               seed=14, full=False)
    # E: config-1 family: MLP with BatchNorm + ReLU + dropout + softmax output
    if not only or "mlp_bn_relu" in only:
      mlp_case("mlp_bn_relu", N=64, D=23, lay=[48, 48, 19], drop=[0.15, 0.15, 0.0], bn=[True, True, False],
             ln=[False, False, False], act=["relu", "relu", "softmax"], seed=21)
    if not only or "mlp_ln_tanh" in only:
      mlp_case("mlp_ln_tanh", N=33, D=11, lay=[20, 16, 9], drop=[0.0, 0.1, 0.0], bn=[False, True, False],
             ln=[True, True, False], act=["tanh", "sigmoid", "softmax"], seed=22)
    # F: plain RNN cell (neural_networks.py:1319-1461): bidirectional BN+ReLU stack and a unidirectional tanh/bias one
    if not only or "rnn_bidir_bn" in only:
      ligru_case("rnn_bidir_bn", T=14, B=6, D=11, lay=[40, 40], S=23, S2=0, drop=0.2, bn=True, act="relu", bidir=True,
                 seed=31, cell="rnn")
    if not only or "rnn_uni_tanh" in only:
      ligru_case("rnn_uni_tanh", T=10, B=3, D=9, lay=[33], S=12, S2=0, drop=0.1, bn=False, act="tanh", bidir=False,
                 seed=32, cell="rnn")
    # G: LSTM :300-483, GRU :486-654, minimalGRU :1158-1316 — bidirectional BN+ReLU/tanh stacks and bias variants
    for cell, seed in (("lstm", 41), ("gru", 51), ("minimalgru", 61)):
        if not only or f"{cell}_bidir_bn" in only:
            ligru_case(f"{cell}_bidir_bn", T=13, B=5, D=11, lay=[36, 28], S=19, S2=0, drop=0.2, bn=True,
                       act="tanh" if cell == "lstm" else "relu", bidir=True, seed=seed, cell=cell)
        if not only or f"{cell}_uni_nobn" in only:
            ligru_case(f"{cell}_uni_nobn", T=9, B=3, D=8, lay=[21], S=11, S2=0, drop=0.1, bn=False,
                       act="tanh", bidir=False, seed=seed + 1, cell=cell)
    # H: conv front-ends: SincNet :1559-1813 (TIMIT_SincNet_raw.cfg in miniature) and CNN :1464-1556
    if not only or "sincnet_ln_relu" in only:
        conv_case("sincnet_ln_relu", kind="SincNet", N=6, L0=400, n_filt=[16, 12, 12], len_filt=[33, 5, 3],
                  pool=[3, 3, 2], ln=True, ln_inp=True, act="relu", drop=0.15, S=9, seed=71)
    if not only or "sincnet_tanh_noln" in only:
        conv_case("sincnet_tanh_noln", kind="SincNet", N=5, L0=300, n_filt=[10, 8], len_filt=[21, 4],
                  pool=[2, 3], ln=False, ln_inp=False, act="tanh", drop=0.0, S=7, seed=72)
    if not only or "cnn_ln_relu" in only:
        conv_case("cnn_ln_relu", kind="CNN", N=7, L0=120, n_filt=[20, 12, 12], len_filt=[10, 3, 3], pool=[3, 2, 1],
                  ln=True, ln_inp=False, act="leaky_relu", drop=0.15, S=8, seed=73)
    # I: input side (context window + chunk normalisation + minibatch assembly)
    if not only or "input_cw" in only:
        input_case("input_cw", N=400, F=7, left=3, right=2, n_snt=8, batch=4, seed=81)
    if not only or "input_nocw" in only:
        input_case("input_nocw", N=300, F=5, left=0, right=0, n_snt=6, batch=3, seed=82)
    # J: output side (prior-normalised posteriors -> Kaldi ark)
    if not only or "post_ark" in only:
        ark_case("post_ark", rows=11, cols=23, seed=91)
    # K: input side (Kaldi archive reader incl. compressed matrices)
    if not only or "ark_read" in only:
        ark_read_case("ark_read", seed=95)
