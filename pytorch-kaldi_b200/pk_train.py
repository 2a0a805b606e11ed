"""Training-step glue for the drop-in modules: flat parameter / gradient buffers, ONE gradient
allreduce per step (NCCL over NVLink when world_size > 1) and ONE fused optimizer kernel.

Mirrors what core.run_nn does per minibatch (reference core.py:616-642): forward_model ->
zero_grad -> loss.backward() -> optimizer.step(), and replaces torch.nn.DataParallel
(core.py:537-538) by one process per GPU with utterance columns sharded across ranks
(SURVEY.md 8e).  The optimizer math is torch.optim.RMSprop / SGD as configured by
utils.optimizer_init (utils.py:2106-2164).
"""
from __future__ import annotations

import struct
from typing import Iterable, List

import numpy as np

import torch
import torch.distributed as dist

import pk_native as pk


class FlatTrainer:
    """Re-homes the parameters of `modules` into one contiguous fp32 buffer (and their gradients
    into another), so that a step is: backward -> (allreduce of the flat gradient) -> one
    rmsprop/sgd/adam kernel over the flat buffers.  state_dict()/load_state_dict() of the modules keep
    working (parameters stay nn.Parameters, only their storage moves).

    Optimizer semantics are torch.optim's as configured by utils.optimizer_init (utils.py:2106-2164); options the
    fused kernels do not implement are REFUSED (never silently ignored).  optimizer_state_dicts() /
    load_optimizer_state_dicts() emit / accept the torch.optim state_dict layout per architecture, i.e. what the
    reference stores as `optimizer_par` and reloads for every chunk (core.py:523-535, :713-722)."""

    def __init__(self, modules: Iterable[torch.nn.Module], opt: str = "rmsprop", lr: float = 0.0004,
                 alpha: float = 0.95, eps: float = 1e-8, betas=(0.9, 0.999), weight_decay: float = 0.0,
                 momentum: float = 0.0, centered: bool = False, nesterov: bool = False, dampening: float = 0.0,
                 amsgrad: bool = False, direct_grad: bool = True):
        if opt not in ("rmsprop", "sgd", "adam"):
            raise NotImplementedError(f"FlatTrainer: optimizer {opt!r} (reference offers sgd / adam / rmsprop)")
        unsupported = []
        if opt == "rmsprop":
            unsupported = [n for n, v in (("momentum", momentum), ("centered", centered), ("weight_decay", weight_decay)) if v]
        elif opt == "sgd":
            unsupported = [n for n, v in (("momentum", momentum), ("weight_decay", weight_decay), ("nesterov", nesterov),
                                          ("dampening", dampening)) if v]
        elif amsgrad:
            unsupported = ["amsgrad"]
        if unsupported:
            raise NotImplementedError(f"FlatTrainer({opt}): option(s) {unsupported} are not implemented by the fused optimizer "
                                      "kernel; refusing instead of silently ignoring them")
        self.modules: List[torch.nn.Module] = list(modules)
        self.params = [p for m in self.modules for p in m.parameters()]   # registration order = optimizer index order
        if not self.params:
            raise RuntimeError("FlatTrainer: no parameters")
        self._require_device(self.params[0])
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.flat_p = torch.empty(n, device=dev, dtype=torch.float32)
        self.flat_g = torch.zeros(n, device=dev, dtype=torch.float32)
        self.flat_v = torch.zeros(n, device=dev, dtype=torch.float32) if opt in ("rmsprop", "adam") else None
        self.flat_m = torch.zeros(n, device=dev, dtype=torch.float32) if opt == "adam" else None
        self.betas, self.weight_decay, self.steps = betas, weight_decay, 0
        # Placement inside the flat buffer: the gate matrices of one recurrent layer go back to back ([wh_i; wz_i],
        # [uh_i; uz_i], ...) so that the stacked operands the kernels want are zero-copy views and their gradients are
        # written in place (pk_functions._stacked / DIRECT_GRAD); everything else follows in registration order.  The
        # optimizer-state indices (self.params order) are not affected.
        placed, order = set(), []
        for m in self.modules:
            for sub in m.modules():
                gates = getattr(sub, "_GATES", None)
                if not gates or not hasattr(sub, "lay"):
                    continue
                for i in range(len(sub.lay)):
                    for col in (0, 1):
                        for pair in gates:
                            p = getattr(sub, pair[col])[i].weight
                            if id(p) not in placed:
                                placed.add(id(p))
                                order.append(p)
        order += [p for p in self.params if id(p) not in placed]
        where = {}
        off = 0
        for p in order:
            k = p.numel()
            self.flat_p[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat_p[off:off + k].view_as(p)
            p.grad = self.flat_g[off:off + k].view_as(p)
            where[id(p)] = (off, k)
            off += k
        self.offsets = [where[id(p)] for p in self.params]
        if direct_grad:
            # one backward per step(): weight gradients are WRITTEN into the flat buffer by the GEMM epilogues instead of
            # being accumulated by autograd (no AccumulateGrad adds); pass direct_grad=False to accumulate several
            # backward passes before a step
            import pk_functions
            pk_functions.register_direct_grad_buffer(self.flat_g)
        self.n = n
        self.opt, self.lr, self.alpha, self.eps = opt, lr, alpha, eps
        self.used = None  # per parameter: does it ever receive a gradient? (decided at the first step)
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1

    @classmethod
    def from_options(cls, modules, options, lr):
        """Build from an [architecture] cfg section the way utils.optimizer_init does (utils.py:2106-2164): arch_opt
        and the opt_* keys (strings)."""
        sb = lambda v: str(v).lower() in ("true", "1", "yes")
        kind = options["arch_opt"]
        if kind == "sgd":
            return cls(modules, "sgd", lr, momentum=float(options["opt_momentum"]), weight_decay=float(options["opt_weight_decay"]),
                       nesterov=sb(options["opt_nesterov"]), dampening=float(options["opt_dampening"]))
        if kind == "adam":
            return cls(modules, "adam", lr, betas=tuple(map(float, str(options["opt_betas"]).split(","))),
                       eps=float(options["opt_eps"]), weight_decay=float(options["opt_weight_decay"]),
                       amsgrad=sb(options["opt_amsgrad"]))
        if kind == "rmsprop":
            return cls(modules, "rmsprop", lr, alpha=float(options["opt_alpha"]), eps=float(options["opt_eps"]),
                       momentum=float(options["opt_momentum"]), centered=sb(options["opt_centered"]),
                       weight_decay=float(options["opt_weight_decay"]))
        raise NotImplementedError(f"arch_opt = {kind!r}")

    def _require_device(self, p):
        if not p.is_cuda:
            raise RuntimeError("FlatTrainer needs CUDA modules (no CPU path)")

    def zero_grad(self):
        self.flat_g.zero_()

    def _mark_used(self):
        """torch.optim skips parameters whose .grad is None (the reference's unused ln.* / disabled bn.* modules).  The
        flat buffer always holds a tensor, so 'never written by backward' is detected once, after the first backward:
        an all-zero gradient range of a parameter means autograd never touched it."""
        nz = torch.stack([self.flat_g[o:o + k].abs().max() if k else self.flat_g.new_zeros(()) for o, k in self.offsets])
        self.used = (nz > 0).cpu().tolist()

    def _ranges(self):
        """Contiguous [offset, length) ranges of the parameters that receive gradients."""
        out = []
        for (o, k), u in zip(self.offsets, self.used):
            if not u or k == 0:
                continue
            if out and out[-1][0] + out[-1][1] == o:
                out[-1][1] += k
            else:
                out.append([o, k])
        return out

    def _apply_update(self, gscale):
        """The fused optimizer kernel(s) over the flat buffers."""
        if self.opt == "rmsprop":
            pk.rmsprop_step(self.flat_p, self.flat_g, self.flat_v, self.lr, self.alpha, self.eps, gscale)
        elif self.opt == "sgd":
            pk.sgd_step(self.flat_p, self.flat_g, self.lr, gscale)
        else:
            self.steps += 1
            if self.weight_decay == 0.0:
                pk.adam_step(self.flat_p, self.flat_g, self.flat_m, self.flat_v, self.lr, self.betas[0], self.betas[1], self.eps,
                             0.0, self.steps, gscale)
            else:
                # weight decay moves parameters even where the gradient is zero: parameters that never receive a
                # gradient must be skipped like torch.optim does -> one launch per contiguous used range
                for o, k in self._ranges():
                    pk.adam_step(self.flat_p[o:o + k], self.flat_g[o:o + k], self.flat_m[o:o + k], self.flat_v[o:o + k], self.lr,
                                 self.betas[0], self.betas[1], self.eps, self.weight_decay, self.steps, gscale)

    def step(self):
        """(allreduce) + optimizer.  Gradients are summed across ranks and scaled by 1/world inside the
        optimizer kernel, i.e. the loss is the mean over the global batch (equal shard sizes)."""
        if self.world > 1:
            dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM)
        if self.used is None and (self.opt == "adam" and self.weight_decay != 0.0):
            self._mark_used()
        if self.opt != "adam":
            self.steps += 1
        self._apply_update(1.0 / self.world)

    # ---- optimizer state in the torch.optim layout, one dict per architecture (module) -------------------------
    def _group(self):
        if self.opt == "rmsprop":
            return dict(lr=self.lr, momentum=0, alpha=self.alpha, eps=self.eps, centered=False, weight_decay=0)
        if self.opt == "sgd":
            return dict(lr=self.lr, momentum=0, dampening=0, weight_decay=0, nesterov=False)
        return dict(lr=self.lr, betas=tuple(self.betas), eps=self.eps, weight_decay=self.weight_decay, amsgrad=False)

    def optimizer_state_dicts(self):
        """[{'state': {i: {...}}, 'param_groups': [{..., 'params': [0..n-1]}]} per module] — the layout of
        torch.optim.{RMSprop,SGD,Adam}.state_dict() (RMSprop: step, square_avg; Adam: step, exp_avg, exp_avg_sq; SGD
        without momentum: empty), parameters indexed in registration order like the reference's per-architecture
        optimizers.  Parameters that never received a gradient carry no state, as in torch.optim."""
        if self.used is None and self.steps > 0:
            self._mark_used()
        out, pi = [], 0
        for m in self.modules:
            ps = list(m.parameters())
            state = {}
            for j, p in enumerate(ps):
                o, k = self.offsets[pi + j]
                if self.steps == 0 or self.opt == "sgd" or (self.used is not None and not self.used[pi + j]):
                    continue
                st = {"step": torch.tensor(float(self.steps))}
                if self.opt == "rmsprop":
                    st["square_avg"] = self.flat_v[o:o + k].view_as(p).clone()
                else:
                    st["exp_avg"] = self.flat_m[o:o + k].view_as(p).clone()
                    st["exp_avg_sq"] = self.flat_v[o:o + k].view_as(p).clone()
                state[j] = st
            g = self._group()
            g["params"] = list(range(len(ps)))
            out.append({"state": state, "param_groups": [g]})
            pi += len(ps)
        return out

    def load_optimizer_state_dicts(self, dicts):
        """Inverse of optimizer_state_dicts(); accepts dictionaries written by the reference's own torch.optim
        optimizers (core.py:531-535: `optimizer_par`)."""
        pi, steps = 0, 0
        for m, d in zip(self.modules, dicts):
            ps = list(m.parameters())
            for j, st in d.get("state", {}).items():
                o, k = self.offsets[pi + int(j)]
                steps = max(steps, int(float(st.get("step", 0))))
                if self.opt == "rmsprop" and "square_avg" in st:
                    self.flat_v[o:o + k].copy_(st["square_avg"].reshape(-1))
                if self.opt == "adam":
                    self.flat_m[o:o + k].copy_(st["exp_avg"].reshape(-1))
                    self.flat_v[o:o + k].copy_(st["exp_avg_sq"].reshape(-1))
            g = d.get("param_groups", [{}])[0]
            self.lr = float(g.get("lr", self.lr))
            pi += len(ps)
        self.steps = steps


class GraphedStep:
    """A whole training step (forward_model -> zero_grad -> backward -> optimizer) captured ONCE into a CUDA graph and
    replayed per minibatch: for launch-bound recipes (the MLP of cfg/TIMIT_baselines/TIMIT_MLP_mfcc_basic.cfg trains on
    128-frame minibatches: ~45 kernels of a few microseconds each) the host launch cost disappears.  `step_fn(inp)`
    must be free of host synchronisation (device-side dropout masks: `fast_dropout` / nn.Dropout-style device RNG) and
    return device tensors; inputs are copied into a static buffer, outputs are read from static tensors.  Shapes are
    fixed at capture (one graph per minibatch shape)."""

    def __init__(self, step_fn, example_inp: torch.Tensor, warmup: int = 3):
        if not example_inp.is_cuda:
            raise RuntimeError("GraphedStep needs CUDA tensors (no CPU path)")
        self.static_inp = example_inp.clone()
        s = torch.cuda.Stream(device=example_inp.device)
        s.wait_stream(torch.cuda.current_stream(example_inp.device))
        with torch.cuda.stream(s):   # warm-up outside the capture: one-time attribute set-up, allocator pools
            for _ in range(warmup):
                step_fn(self.static_inp)
        torch.cuda.current_stream(example_inp.device).wait_stream(s)
        torch.cuda.synchronize(example_inp.device)
        self.graph = torch.cuda.CUDAGraph()
        n0 = pk.launch_count
        with torch.cuda.graph(self.graph):
            self.static_out = step_fn(self.static_inp)
        self.launches_per_replay = pk.launch_count - n0   # library kernels recorded into the graph (one step)

    def __call__(self, inp: torch.Tensor):
        self.static_inp.copy_(inp, non_blocking=True)
        self.graph.replay()
        return self.static_out


def chunk_step(net, head, trainer: FlatTrainer, inp: torch.Tensor, n_fea: int, fused_head: bool = True):
    """One minibatch as core.run_nn + utils.forward_model run it: `inp` is the reference's chunk
    layout [T, B, n_fea + 1] with the label in the last column stored as float (data_io.py:272,
    utils.py:2305-2352).  Returns (loss, err) as device scalars."""
    T, B, _ = inp.shape
    lab = inp[:, :, n_fea].reshape(-1).long()          # utils.py:2348-2352
    x = inp[:, :, :n_fea]                              # utils.py:2321 (a strided view; no copy)
    trainer.zero_grad()
    out = net(x)                                       # out_dnn1 = compute(liGRU_layers, fea)
    if fused_head and head._is_plain_head():
        # out_dnn2 = compute(MLP_layers, out_dnn1); loss = cost_nll(out_dnn2, lab); err = cost_err(out_dnn2, lab)
        # as ONE fused op (linear + log-softmax + NLL + argmax error; utils.py:2339-2381)
        import pk_functions as pkf
        loss, err, _ = pkf.HeadNLLFn.apply(out.view(T * B, -1), head.wx[0].weight, head.wx[0].bias, lab)
    else:
        logp = head(out.view(T * B, -1))                   # out_dnn2 = compute(MLP_layers, out_dnn1)
        loss = torch.nn.functional.nll_loss(logp, lab)     # loss_final = cost_nll(out_dnn2, lab_cd)
        err = (logp.detach().max(dim=1)[1] != lab).float().mean()  # err_final = cost_err(out_dnn2, lab_cd)
    loss.backward()
    trainer.step()
    return loss.detach(), err


def prepare_chunk(fea: torch.Tensor, lab, left: int, right: int) -> torch.Tensor:
    """data_io.load_chunk's array work (data_io.py:255-272) on the device: context window, per-column mean / std
    normalisation, label column.  fea [N, F] fp32 CUDA, lab [N] integer CUDA tensor (or None) -> data_set
    [N-left-right, F*(left+right+1) (+1)] fp32, the layout core.run_nn indexes."""
    if not fea.is_cuda:
        raise RuntimeError("pytorch-kaldi_b200: prepare_chunk needs CUDA tensors (there is no CPU fallback)")
    fea = fea.float().contiguous()
    n_out = fea.shape[0] - left - right
    cols = fea.shape[1] * (left + right + 1) + (1 if lab is not None else 0)
    out = torch.empty(n_out, cols, device=fea.device, dtype=torch.float32)
    lab64, lab_min = None, 0
    if lab is not None:
        lab64 = lab.long().contiguous()
        lab_min = int(lab64.min().item())       # data_io.py:266 (one scalar per chunk)
    pk.chunk_prepare(fea, lab64, lab_min, left, right, out)
    return out


def batch_descriptors(data_end_index, snt_index: int, beg_snt: int, batch_size: int, rng):
    """The bookkeeping of core.py:581-598 for one minibatch: returns (desc [3][B] int64 = first frame / length / left
    zeros, max_len, next snt_index, next beg_snt).  `rng.randint` is consumed exactly like the reference does (one
    draw per sentence, in order), so a seeded run pads identically."""
    lens, begs = [], []
    b, s = beg_snt, snt_index
    for _ in range(batch_size):
        end = int(data_end_index[s])
        begs.append(b)
        lens.append(end - b)
        b, s = end, s + 1
    max_len = max(lens)
    lefts = [rng.randint(0, max_len - L) for L in lens]   # core.py:592
    return torch.tensor([begs, lens, lefts], dtype=torch.int64), max_len, s, b


def assemble_batch(data_set: torch.Tensor, desc: torch.Tensor, max_len: int) -> torch.Tensor:
    """inp [max_len, B, D] of core.py:584-595 built by one gather kernel from the device-resident chunk."""
    if not data_set.is_cuda:
        raise RuntimeError("pytorch-kaldi_b200: assemble_batch needs the chunk on the device (there is no CPU fallback)")
    B = desc.shape[1]
    inp = torch.empty(max_len, B, data_set.shape[1], device=data_set.device, dtype=torch.float32)
    pk.batch_assemble(data_set, desc.to(data_set.device, non_blocking=True), B, max_len, inp)
    return inp


# ---- output side: forward-phase posterior writer -------------------------------------------------


def load_counts(class_counts_file: str) -> np.ndarray:
    """The senone counts file Kaldi's analyze-counts writes: one line `[ c0 c1 ... ]` (data_io.py:277-281)."""
    with open(class_counts_file) as f:
        row = next(f).strip().strip("[]").strip()
    return np.array([np.float32(v) for v in row.split()])


def write_kaldi_matrix(fd, key: str, m: np.ndarray) -> None:
    """One entry of a binary Kaldi matrix archive: `key ` + "\0B" + "FM "/"DM " + \4 rows + \4 cols + row-major
    payload — the byte stream data_io.write_mat produces (data_io.py:1200-1239); fd is a binary stream."""
    if m.dtype == np.float32:
        tag = b"FM "
    elif m.dtype == np.float64:
        tag = b"DM "
    else:
        raise TypeError(f"'{m.dtype}', please use 'float32' or 'float64'")
    if m.ndim != 2:
        raise ValueError("a Kaldi matrix has two axes")
    if key != "":
        fd.write((key + " ").encode("latin1"))
    fd.write(b"\0B" + tag)
    fd.write(b"\x04" + struct.pack("<I", m.shape[0]))
    fd.write(b"\x04" + struct.pack("<I", m.shape[1]))
    fd.write(np.ascontiguousarray(m).tobytes())


def write_posteriors(fd, key: str, logp: torch.Tensor, counts: np.ndarray = None) -> None:
    """core.py:660-671 for one sentence: optional prior normalisation `out - log(counts / sum(counts))` (on the
    device, in place on a copy), device -> host, Kaldi ark entry."""
    if not logp.is_cuda:
        raise RuntimeError("pytorch-kaldi_b200: write_posteriors needs the posteriors on the device (no CPU path)")
    out = logp.detach().float().contiguous()
    if counts is not None:
        out = out.clone()
        log_prior = np.log(counts / np.sum(counts)).astype(np.float32)   # core.py:666-667, float32 like the reference
        pk.sub_log_prior(out, torch.from_numpy(log_prior).to(out.device))
    write_kaldi_matrix(fd, key, out.cpu().numpy())


# ---- input side: Kaldi archive reader -------------------------------------------------------------


def _read_key(fd):
    """Utterance key in front of every archive entry (data_io.py:762-778); None at end of file."""
    key = b""
    while True:
        ch = fd.read(1)
        if ch == b"" or ch == b" ":
            break
        key += ch
    key = key.decode("latin1").strip()
    return key or None


def open_rx(rxspec, mode="rb"):
    """Kaldi rxfilename / rspecifier -> (binary stream, close?) the way data_io.open_or_fd does (data_io.py:685-718):
    optional `ark:` / `scp:` prefix with modifiers, optional `:offset` suffix, `cmd |` input pipes (run through the
    shell, as the reference does — e.g. `copy-feats ... |` when the Kaldi binaries are installed), `.gz` files, plain
    files; an already opened stream is passed through."""
    import gzip
    import re
    import subprocess
    if not isinstance(rxspec, str):
        return rxspec, False
    f, offset = rxspec, None
    if re.search(r"^(ark|scp)(,scp|,b|,t|,n?f|,n?p|,b?o|,n?s|,n?cs)*:", f):
        f = f.split(":", 1)[1]
    if re.search(r":[0-9]+$", f):
        f, off = f.rsplit(":", 1)
        offset = int(off)
    f = f.strip()
    if f.endswith("|"):
        proc = subprocess.Popen(f[:-1], shell=True, stdout=subprocess.PIPE)
        fd = proc.stdout
    elif f.split(".")[-1] == "gz":
        fd = gzip.open(f, mode)
    else:
        fd = open(f, mode)
    if offset is not None:
        fd.seek(offset)
    return fd, True


def _read_ascii_matrix(fd):
    """Text-mode Kaldi matrix after the opening " [" (data_io.py:1133-1147): rows of numbers, the last one ends in "]"."""
    rows = []
    while True:
        line = fd.readline().decode()
        if len(line) == 0:
            raise ValueError("unexpected end of file inside an ascii matrix")
        arr = line.strip().split()
        if not arr:
            continue
        if arr[-1] != "]":
            rows.append(np.array(arr, dtype="float32"))
        else:
            rows.append(np.array(arr[:-1], dtype="float32"))
            return np.vstack(rows)


def _read_binary_matrix(fd, device=None):
    """One Kaldi matrix at the stream position (after the key): "\0B" + "FM " / "DM " / "CM " (data_io.py:1087-1196), or
    the text form " [ ... ]" (data_io.py:1133-1147)."""
    flag = fd.read(2)
    if flag == b" [":
        mat = _read_ascii_matrix(fd)
        return torch.from_numpy(mat).to(device) if device is not None else mat
    if flag != b"\0B":
        raise ValueError("neither a binary nor a text Kaldi matrix")
    header = fd.read(3).decode()
    if header == "CM ":
        gmin, grange, rows, cols = np.frombuffer(fd.read(16), dtype="float32,float32,int32,int32", count=1)[0]
        hdr = np.frombuffer(fd.read(int(cols) * 8), dtype=np.uint16).copy()
        data = np.frombuffer(fd.read(int(cols) * int(rows)), dtype=np.uint8).copy()
        if device is None:
            raise RuntimeError("pytorch-kaldi_b200: compressed matrices are decoded on the GPU (pass device=; no CPU path)")
        out = torch.empty(int(rows), int(cols), device=device, dtype=torch.float32)
        pk.cm_decode(torch.from_numpy(hdr.view(np.int16)).to(device), torch.from_numpy(data).to(device), gmin, grange,
                     int(rows), int(cols), out)
        return out
    if header == "FM ":
        dt = np.float32
    elif header == "DM ":
        dt = np.float64
    else:
        raise ValueError(f"The header contained '{header}'")
    _, rows, _, cols = np.frombuffer(fd.read(10), dtype="int8,int32,int8,int32", count=1)[0]
    mat = np.frombuffer(fd.read(int(rows) * int(cols) * np.dtype(dt).itemsize), dtype=dt).reshape(int(rows), int(cols))
    return torch.from_numpy(mat.copy()).to(device) if device is not None else mat


def read_mat_ark(fd, device=None):
    """Generator of (key, matrix) over a binary Kaldi matrix archive (data_io.py:1062-1131): "FM " / "DM " payloads are
    returned as numpy arrays (or moved to `device`); "CM " (CompressedMatrix) payloads are uploaded as bytes and decoded
    on the GPU (`pk_cm_decode`), which needs `device`."""
    while True:
        key = _read_key(fd)
        if key is None:
            return
        yield key, _read_binary_matrix(fd, device)


def read_mat_scp(scp_path, device=None):
    """Generator of (key, matrix) over a Kaldi script file, one `key path[:offset]` per line (data_io.py:1039-1059,
    offsets as open_or_fd handles them, :696-716)."""
    with open(scp_path) as scp:
        for line in scp:
            if not line.strip():
                continue
            key, rx = line.strip().split(None, 1)
            fd, close = open_rx(rx)   # plain / gzipped file, `cmd |` pipe, optional :offset
            try:
                yield key, _read_binary_matrix(fd, device)
            finally:
                if close:
                    fd.close()


def read_vec_flt(fd):
    """One Kaldi float vector at the stream position (data_io.py:922-990): binary "\0B" + "FV " / "DV " + \4 + int32
    length + payload, or the text form "[ a b c ]"."""
    flag = fd.read(2)
    if flag == b"\0B":
        header = fd.read(3).decode()
        if header not in ("FV ", "DV "):
            raise ValueError(f"The header contained '{header}'")
        dt = np.float32 if header == "FV " else np.float64
        if fd.read(1) != b"\x04":
            raise ValueError("bad float-vector header")
        n = int(np.frombuffer(fd.read(4), dtype="int32", count=1)[0])
        if n == 0:
            return np.array([], dtype="float32")
        return np.frombuffer(fd.read(n * np.dtype(dt).itemsize), dtype=dt).copy()
    arr = (flag + fd.readline()).decode().strip().split()
    arr = [a for a in arr if a not in ("[", "]")]
    return np.array(arr, dtype=float)


def read_vec_flt_ark(fd):
    """Generator of (key, float vector) over a Kaldi float-vector archive (data_io.py:901-920), binary or text."""
    while True:
        key = _read_key(fd)
        if key is None:
            return
        yield key, read_vec_flt(fd)


def read_vec_int_ark(fd):
    """Generator of (key, int32 vector) over a binary Kaldi integer-vector archive — alignments (data_io.py:790-836):
    "\0B", \4, int32 length, then (int8 size = 4, int32 value) pairs."""
    while True:
        key = _read_key(fd)
        if key is None:
            return
        if fd.read(2) != b"\0B":
            raise ValueError("only binary Kaldi archives are supported")
        if fd.read(1) != b"\x04":
            raise ValueError("bad integer-vector header")
        n = int(np.frombuffer(fd.read(4), dtype="int32", count=1)[0])
        if n == 0:
            yield key, np.array([], dtype="int32")
            continue
        vec = np.frombuffer(fd.read(n * 5), dtype=[("size", "int8"), ("value", "int32")], count=n)
        if vec[0]["size"] != 4:
            raise ValueError("integer vector with a non-int32 element size")
        yield key, vec["value"].copy()


def read_post(fd):
    """One Kaldi `Posterior` (vector<vector<pair<int32, float>>>: frames x (index, value) records) at the stream position
    (data_io.py:1314-1348): "\0B", then \4 + int32 frame count, per frame \4 + int32 record count and 10-byte records
    (int8 4, int32 index, int8 4, float32 value).  Returns a list (frames) of lists of (index, value) tuples like the
    reference; an empty frame is an empty list (the reference asserts on it)."""
    if fd.read(2) != b"\0B":
        raise ValueError("only binary Kaldi posteriors are supported")
    if fd.read(1) != b"\x04":
        raise ValueError("bad posterior header")
    n_frames = int(np.frombuffer(fd.read(4), dtype="int32", count=1)[0])
    rec = np.dtype([("size_idx", "int8"), ("idx", "int32"), ("size_post", "int8"), ("post", "float32")])
    out = []
    for _ in range(n_frames):
        if fd.read(1) != b"\x04":
            raise ValueError("bad posterior frame header")
        n = int(np.frombuffer(fd.read(4), dtype="int32", count=1)[0])
        data = np.frombuffer(fd.read(n * 10), dtype=rec, count=n)
        if n and (data[0]["size_idx"] != 4 or data[0]["size_post"] != 4):
            raise ValueError("posterior record with a non-4-byte field")
        out.append(data[["idx", "post"]].tolist())
    return out


def read_post_ark(fd):
    """Generator of (key, posterior) over a binary Kaldi posterior archive (data_io.py:1291-1311)."""
    while True:
        key = _read_key(fd)
        if key is None:
            return
        yield key, read_post(fd)


def read_post_scp(scp_path):
    """Generator of (key, posterior) over a Kaldi script file of posteriors (data_io.py:1270-1288)."""
    with open(scp_path) as scp:
        for line in scp:
            if not line.strip():
                continue
            key, rx = line.strip().split(None, 1)
            fd, close = open_rx(rx)
            try:
                yield key, read_post(fd)
            finally:
                if close:
                    fd.close()


def read_post_rxspec(spec):
    """`ark:...` / `scp:...` adaptor (data_io.py:1256-1267)."""
    if spec.startswith("ark:"):
        fd, close = open_rx(spec)
        try:
            yield from read_post_ark(fd)
        finally:
            if close:
                fd.close()
    elif spec.startswith("scp:"):
        yield from read_post_scp(spec.split(":", 1)[1])
    else:
        raise ValueError(f"unsupported input type {spec!r}: it should begin with 'ark:' or 'scp:'")


def read_cntime(fd):
    """Kaldi confusion-network bin times, vector<pair<float, float>> (data_io.py:1384-1416): "\0B", \4 + int32 count, then
    10-byte records (int8 4, float32 begin, int8 4, float32 end).  Returns a list of (begin, end) tuples."""
    if fd.read(2) != b"\0B":
        raise ValueError("only binary Kaldi confusion-network times are supported")
    if fd.read(1) != b"\x04":
        raise ValueError("bad cntime header")
    n = int(np.frombuffer(fd.read(4), dtype="int32", count=1)[0])
    rec = np.dtype([("size_beg", "int8"), ("t_beg", "float32"), ("size_end", "int8"), ("t_end", "float32")])
    data = np.frombuffer(fd.read(n * 10), dtype=rec, count=n)
    if n and (data[0]["size_beg"] != 4 or data[0]["size_end"] != 4):
        raise ValueError("cntime record with a non-4-byte field")
    return data[["t_beg", "t_end"]].tolist()


def read_cntime_ark(fd):
    """Generator of (key, bin times) over a Kaldi confusion-network time archive (data_io.py:1360-1381)."""
    while True:
        key = _read_key(fd)
        if key is None:
            return
        yield key, read_cntime(fd)
