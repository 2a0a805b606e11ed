"""torch.autograd.Function shims over the C ABI (include/pk_b200.h).

These are the only places where the drop-in modules of neural_networks.py touch the device:
every forward / backward below is a fixed sequence of libpk_b200.so calls (tcgen05 GEMMs, the
persistent cluster recurrent kernels, the streaming companions).  torch is used for device
memory, streams and autograd bookkeeping only.  There is no eager-PyTorch or CPU fallback.

Data layout ("Data layout in HBM"):
  row-major     [T*B, C]    frames as rows, n = t*B + b   (the reference's own 2-D view, utils.py:2323)
  channel-major [C, ldt]    one row per unit/feature, ldt = pad8(T*B)
fp16 copies are the tensor-core operands (fp32 accumulate); gradients are multiplied by a
power-of-two loss scale before the fp16 conversion and the GEMM epilogues undo it.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
from typing import List, Optional

import torch

import pk_native as pk

pad8 = pk.pad8


@dataclass
class RecLayerCfg:
    """Static description of one recurrent layer (built by neural_networks.liGRU.forward)."""
    H: int
    act: int
    use_bn: bool
    bn_training: bool            # BatchNorm uses batch statistics (module.training)
    bns: List[torch.nn.Module] = field(default_factory=list)  # per-gate nn.BatchNorm1d (running stats updated in place)
    mask: Optional[torch.Tensor] = None      # [ndir*B, H] device tensor (training) or None
    mask_scalar: float = 1.0                 # eval: 1 - p
    # liGRU_layer (:861-870) keeps a bias on the Linear IN FRONT of BatchNorm: it cancels against the batch statistics
    # (training) and is folded into the shift (eval); per real gate, or None
    proj_bias: Optional[List[torch.Tensor]] = None


@dataclass
class RecStackCfg:
    bidir: bool
    layers: List[RecLayerCfg] = field(default_factory=list)
    cell: int = pk.CELL_LIGRU
    cell_flags: int = 0  # tuning flags OR-ed into the cell id (cluster size, ...)
    grad_enabled: bool = True  # torch.is_grad_enabled() at call time (inside Function.forward it is always off)


def _rows_view(x2d_src: torch.Tensor, T: int, B: int):
    """[T,B,D] (possibly a column slice of the chunk tensor, utils.py:2321) -> (tensor, row pitch)."""
    x = x2d_src
    if x.dtype != torch.float32:
        x = x.float()
    if x.stride(2) != 1 or x.stride(0) != B * x.stride(1):
        x = x.contiguous()
    return x, x.stride(1)


_SIDE_STREAMS = {}


def _side_stream(dev):
    """One extra stream per device for work that is off the critical path of the backward sweep (weight-gradient
    GEMMs run next to the next layer's recurrent kernel, which occupies 80 of the 148 SMs)."""
    key = torch.device(dev).index if torch.device(dev).index is not None else torch.cuda.current_device()
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=key)
    return _SIDE_STREAMS[key]


OVERLAP_WGRAD = os.environ.get("PK_OVERLAP", "1") != "0"

# Set by pk_train.FlatTrainer: parameters live in ONE flat buffer whose gradient buffer is zeroed / consumed once per
# step, every parameter is used once per step -> backward may write a weight gradient straight into `param.grad`
# (a view of the flat gradient buffer) and hand autograd `None`, which removes the AccumulateGrad read-modify-write
# (one eager `add` per parameter per step) from the hot path.
DIRECT_GRAD = False
_DIRECT_STORAGES = set()   # data_ptr of the flat gradient storages that opted in (FlatTrainer(direct_grad=True))


def register_direct_grad_buffer(flat_g):
    global DIRECT_GRAD
    _DIRECT_STORAGES.add(flat_g.untyped_storage().data_ptr())
    DIRECT_GRAD = True


def _stacked(ts):
    """[t0; t1; ...] as ONE matrix without a copy when the tensors are adjacent rows of the same storage (FlatTrainer
    lays the gates of a layer out back to back), else None."""
    t0 = ts[0]
    if t0.dim() != 2 or not all(t.is_contiguous() and t.dtype == t0.dtype and t.shape[1] == t0.shape[1] for t in ts):
        return None
    off = t0.storage_offset()
    for t in ts:
        if t.untyped_storage().data_ptr() != t0.untyped_storage().data_ptr() or t.storage_offset() != off:
            return None
        off += t.numel()
    rows = sum(t.shape[0] for t in ts)
    return t0.as_strided((rows, t0.shape[1]), (t0.shape[1], 1))


def _direct_grad_target(params):
    """The stacked `.grad` view of `params` if gradients may be written in place (see DIRECT_GRAD), else None."""
    if not DIRECT_GRAD or any(not p.is_leaf for p in params):   # views (e.g. slices of a stacked matrix) have no .grad
        return None
    gs = [p.grad for p in params]
    if any(g is None or g.untyped_storage().data_ptr() not in _DIRECT_STORAGES for g in gs):
        return None
    return _stacked(gs)

PERSISTENT_MAX_H = 1024  # largest hidden size the persistent tcgen05 kernels hold (16 CTAs x 64 units; csrc/pk_rnn_tc.cu)


def _kernel_gates(cell: int) -> int:
    """Gate blocks the recurrent kernels see (single-gate RNN is padded with a zero update gate)."""
    return {pk.CELL_LSTM: 4, pk.CELL_GRU: 3}.get(cell, 2)


def _real_gates(cell: int) -> int:
    return {pk.CELL_RNN: 1, pk.CELL_LSTM: 4, pk.CELL_GRU: 3}.get(cell, 2)


_TWO_PHASE = (pk.CELL_GRU, pk.CELL_MGRU)  # candidate contracts (gate * h): two dependent products per step


def _stepwise(cell: int, H: int) -> bool:
    """LSTM, GRU, minimalGRU and large-H liGRU run on the step-wise kernels (pk_cell_step.cu)."""
    return cell in (pk.CELL_LSTM, pk.CELL_GRU, pk.CELL_MGRU) or (cell == pk.CELL_LIGRU and H > PERSISTENT_MAX_H)


class LiGRUStackFn(torch.autograd.Function):
    """Whole recurrent stack (liGRU :1082-1155, RNN :1398-1461, LSTM :403-483 of the reference
    neural_networks.py) and its autograd, as a fixed sequence of C-ABI calls per layer."""

    @staticmethod
    def forward(ctx, x, cfg: RecStackCfg, *params):
        # params per layer (ngr real gates, reference registration order): w_g..., u_g..., then per gate
        # (bn.weight, bn.bias) if use_bn else per gate (w.bias).
        if not x.is_cuda:
            raise RuntimeError("pytorch-kaldi_b200: recurrent layers need CUDA tensors (there is no CPU fallback)")
        dev = x.device
        T, B, D0 = x.shape
        TB = T * B
        ldt = pad8(TB)
        ndir = 2 if cfg.bidir else 1
        # needs_input_grad reflects requires_grad of the inputs, not the grad mode: under torch.no_grad() (the valid /
        # forward phases) nothing is saved for backward and none of the training-only buffers is allocated or written
        need_grad = cfg.grad_enabled and any(ctx.needs_input_grad)
        f32 = dict(device=dev, dtype=torch.float32)
        f16 = dict(device=dev, dtype=torch.float16)
        ngr, ngk = _real_gates(cfg.cell), _kernel_gates(cfg.cell)

        saved = []  # per layer dict of tensors needed by backward
        xsrc, ldx = _rows_view(x, T, B)
        main = torch.cuda.current_stream(dev)
        side = _side_stream(dev)

        # ---- weight-only work of ALL layers up front: fp16 / transposed copies of the stacked projection weights
        # and the stacked recurrent weights depend on parameters only, so layers 1.. are prepared on the side stream
        # while layer 0's recurrence (80 of 148 SMs) runs; each layer waits on its own event.
        packs = []
        pi = 0
        D = D0
        if OVERLAP_WGRAD:
            side.wait_stream(main)   # parameters were last written (optimizer) on the main stream
        for li, L in enumerate(cfg.layers):
            H = L.H
            ws_, us_ = list(params[pi:pi + ngr]), list(params[pi + ngr:pi + 2 * ngr])
            pi += 2 * ngr
            gammas = betas = biases = None
            if L.use_bn:
                bnp = params[pi:pi + 2 * ngr]
                pi += 2 * ngr
                gammas = [bnp[2 * g] for g in range(ngr)]
                betas = [bnp[2 * g + 1] for g in range(ngr)]
            else:
                biases = list(params[pi:pi + ngr])
                pi += ngr
            CG = ngk * H
            ldD, ldG = pad8(D), pad8(CG)
            on_side = OVERLAP_WGRAD and li > 0
            with torch.cuda.stream(side if on_side else main):
                wparams, uparams = list(ws_), list(us_)   # the nn.Parameters themselves (targets of direct gradients)
                for _ in range(ngk - ngr):  # RNN: the update-gate block is all zeros (and pinned to 0 in the kernel)
                    ws_.append(torch.zeros_like(ws_[0]))
                    us_.append(torch.zeros_like(us_[0]))
                Wcat = _stacked(ws_)       # zero-copy when the gates are adjacent in the flat parameter buffer
                if Wcat is None:
                    Wcat = torch.cat(ws_, 0).contiguous()
                W16 = torch.empty(CG, ldD, **f16)
                WT16 = torch.empty(D, ldG, **f16) if need_grad else None
                pk.transpose_f32(Wcat, D, CG, D, outT16=WT16, ldo16=ldG, in16=W16, ldi16=ldD)
                U = _stacked(us_)
                if U is None:
                    U = torch.cat(us_, 0).contiguous()
                bias_cat = None
                if biases is not None:
                    bias_cat = torch.cat(biases + [torch.zeros_like(biases[0]) for _ in range(ngk - ngr)]).contiguous()
                ev = None
                if on_side:
                    ev = torch.cuda.Event()
                    ev.record(side)
                    for t in (Wcat, W16, WT16, U, bias_cat):
                        if t is not None:
                            t.record_stream(main)   # allocated under the side stream, consumed on the main one
            packs.append(dict(W16=W16, WT16=WT16, U=U, gammas=gammas, betas=betas, bias_cat=bias_cat, ev=ev,
                              wparams=wparams, uparams=uparams))
            D = ndir * H

        D = D0
        X16 = XT16 = None
        y32 = None
        for li, L in enumerate(cfg.layers):
            H = L.H
            pkd = packs[li]
            W16, WT16, gammas, betas = pkd["W16"], pkd["WT16"], pkd["gammas"], pkd["betas"]
            if pkd["ev"] is not None:
                main.wait_event(pkd["ev"])
            CG = ngk * H
            ldD = pad8(D)
            ldG = pad8(CG)
            # ---- operands: fp16 copy of the layer input (layer 0; deeper layers read the previous layer's fp16 outputs)
            if li == 0:
                X16 = torch.empty(TB, ldD, **f16)
                XT16 = torch.empty(D, ldt, **f16) if need_grad else None
                pk.transpose_f32(xsrc, ldx, TB, D, outT16=XT16, ldo16=ldt, in16=X16, ldi16=ldD)
            # ---- projection PT = [W_g] X^T (channel-major) with BatchNorm statistics in the epilogue
            PT = torch.empty(CG, ldt, **f32)
            bn_train = L.use_bn and L.bn_training
            stats = torch.zeros(CG, 2, device=dev, dtype=torch.float64) if bn_train else None
            pk.gemm_tn(W16, X16, PT, CG, TB, D, lda=ldD, ldb=ldD, ldc=ldt, rowstats=stats)
            scale = torch.empty(CG, **f32)
            shift = torch.empty(CG, **f32)
            mean = torch.empty(CG, **f32) if L.use_bn else None
            rstd = torch.empty(CG, **f32) if L.use_bn else None
            gamma = None
            if L.use_bn:
                if ngk > ngr:  # padded gate: identity "normalisation" (its projections are exactly zero)
                    scale[ngr * H:].fill_(1.0); shift[ngr * H:].zero_(); mean[ngr * H:].zero_(); rstd[ngr * H:].fill_(1.0)
                for gi in range(ngr):
                    bn = L.bns[gi]
                    sl = slice(gi * H, (gi + 1) * H)
                    pk.bn_finalize(stats[sl] if bn_train else None, H, TB, TB * ndir, gammas[gi], betas[gi], bn.eps,
                                   bn.momentum if bn.momentum is not None else 0.1, bn_train,
                                   bn.running_mean, bn.running_var, bn.num_batches_tracked if bn_train else None,
                                   scale[sl], shift[sl], mean[sl], rstd[sl])
                    if L.proj_bias is not None:
                        if bn_train:   # the statistics came from W x; the module normalises W x + b
                            bn.running_mean.add_(L.proj_bias[gi].detach(), alpha=bn.momentum if bn.momentum is not None else 0.1)
                        else:
                            shift[sl].addcmul_(scale[sl], L.proj_bias[gi].detach())
                gamma = torch.cat(gammas + [torch.ones_like(gammas[0]) for _ in range(ngk - ngr)]).contiguous()
            else:
                pk.fill_scale_shift(pkd["bias_cat"], CG, scale, shift)
            # ---- the recurrence
            U = pkd["U"]
            F = ndir * H
            last = li == len(cfg.layers) - 1
            ldF = pad8(F)
            Y16 = None if last else torch.empty(TB, ldF, **f16)
            if last:
                y32 = torch.empty(T, B, F, **f32)
            HT = torch.empty(F, ldt, **f32) if need_grad else None
            HT16 = torch.empty(F, ldt, **f16) if (need_grad and not last) else None
            HP16 = torch.empty(F, ldt, **f16) if need_grad else None
            stepwise = _stepwise(cfg.cell, H)
            nsv = {pk.CELL_LSTM: 5, pk.CELL_GRU: 3}.get(cfg.cell, 2)
            HX16 = torch.empty(F, ldt, **f16) if (need_grad and cfg.cell in _TWO_PHASE) else None
            SV = [torch.empty(F, ldt, **f32) if need_grad else None for _ in range(nsv)]
            if stepwise:  # one fused kernel per time step, T launches inside this call
                wsb = pk.rnn_step_workspace_bytes(cfg.cell, T, B, H, ndir, False)
                wsp = torch.empty(wsb, device=dev, dtype=torch.uint8)
                pk.rnn_step_fwd(cfg.cell, T, B, H, ndir, L.act, PT, ldt, scale, shift, U, L.mask, L.mask_scalar,
                                y32 if last else None, F, Y16, ldF, HT, HT16, HP16, HX16, SV, ldt, wsp)
            else:  # one persistent cluster kernel for all T steps
                pk.rnn_layer_fwd(cfg.cell | cfg.cell_flags, T, B, H, ndir, L.act, PT, ldt, scale, shift, U, L.mask,
                                 L.mask_scalar, y32 if last else None, F, Y16, ldF, HT, HT16, HP16, SV[0], SV[1], ldt)
            if need_grad:
                saved.append(dict(D=D, H=H, XT16=XT16, WT16=WT16, PT=PT if bn_train else None, mean=mean, rstd=rstd,
                                  gamma=gamma, HT=HT, SV=SV, HP16=HP16, HX16=HX16, U=U, mask=L.mask, mask_scalar=L.mask_scalar,
                                  act=L.act, use_bn=L.use_bn, bn_train=bn_train, stepwise=stepwise,
                                  wparams=pkd["wparams"], uparams=pkd["uparams"]))
            # next layer reads this layer's fp16 outputs directly
            X16, XT16, D = Y16, HT16, F
        ctx.cfg = cfg
        ctx.saved = saved
        ctx.dims = (T, B, D0, ldt, ndir)
        ctx.x_needs_grad = ctx.needs_input_grad[0]
        return y32

    @staticmethod
    def backward(ctx, dY):
        cfg, saved = ctx.cfg, ctx.saved
        T, B, D0, ldt, ndir = ctx.dims
        TB = T * B
        dev = dY.device
        f32 = dict(device=dev, dtype=torch.float32)
        f16 = dict(device=dev, dtype=torch.float16)
        ngr, ngk = _real_gates(cfg.cell), _kernel_gates(cfg.cell)
        grads = []
        dYT = None
        dx = None
        amax_acc = torch.zeros(1, device=dev, dtype=torch.int32)  # |grad| max accumulated by the producers
        main = torch.cuda.current_stream(dev)
        side = _side_stream(dev)
        for li in reversed(range(len(saved))):
            S = saved[li]
            H, D = S["H"], S["D"]
            F, CG = ndir * H, ngk * H
            ldG = pad8(CG)
            if dYT is None:  # top layer: autograd hands us the row-major gradient of the module output
                dy2 = dY.reshape(TB, F)
                if dy2.dtype != torch.float32 or not dy2.is_contiguous():
                    dy2 = dy2.float().contiguous()
                dYT = torch.empty(F, ldt, **f32)
                pk.transpose_f32(dy2, F, TB, F, outT=dYT, ldo=ldt, amax_bits=amax_acc)
            # power-of-two loss scale for this layer's fp16 gradient operands (amax came with the producer)
            sc = torch.empty(2, **f32)
            pk.amax_finalize(amax_acc, 8.0, sc)
            GT = None
            GT16 = torch.empty(ndir, CG, ldt, **f16)
            if S["stepwise"]:
                wsb = pk.rnn_step_workspace_bytes(cfg.cell, T, B, H, ndir, True)
                wsp = torch.empty(wsb, device=dev, dtype=torch.uint8)
                pk.rnn_step_bwd(cfg.cell, T, B, H, ndir, S["act"], dYT, S["HT"], S["SV"], ldt, S["U"], S["mask"],
                                S["mask_scalar"], sc, GT16, wsp)
            else:
                pk.rnn_layer_bwd(cfg.cell | cfg.cell_flags, T, B, H, ndir, S["act"], dYT, S["HT"], S["SV"][0],
                                 S["SV"][1], ldt, S["U"], S["mask"], S["mask_scalar"], sc, GT, GT16)
            inv = sc[1:2]
            # BatchNorm backward on the de-duplicated projection (both directions folded) — on the critical path
            dgamma = torch.empty(CG, **f32)
            dbeta = torch.empty(CG, **f32)
            need_dx = li > 0 or ctx.x_needs_grad
            dPT16 = torch.empty(CG, ldt, **f16)
            dP16 = torch.empty(TB, ldG, **f16) if need_dx else None
            sums = torch.empty(2 * CG, device=dev, dtype=torch.float64)
            pk.bn_bwd(CG, ndir, TB, GT, GT16, ldt, S["PT"], ldt, S["use_bn"], S["bn_train"], S["mean"], S["rstd"],
                      S["gamma"], sc, dgamma, dbeta, dPT16, ldt, dP16, ldG, sums)
            # weight gradients: off the critical path -> side stream, next to the following layer's recurrence
            dU_direct = _direct_grad_target(S["uparams"]) if ngk == ngr else None
            dW_direct = _direct_grad_target(S["wparams"]) if ngk == ngr else None
            dU = dU_direct if dU_direct is not None else torch.empty(CG, H, **f32)
            dW = dW_direct if dW_direct is not None else torch.empty(CG, D, **f32)
            if OVERLAP_WGRAD:
                side.wait_stream(main)
                for t in (GT16, S["HP16"], S["HX16"], dPT16, S["XT16"], dU, dW, sc):
                    if t is not None:
                        t.record_stream(side)
            with torch.cuda.stream(side if OVERLAP_WGRAD else main):
                # dU = sum_t G_t^T h_{t-1}  (both directions accumulate into the shared weights)
                for d in range(ndir):
                    hp = S["HP16"][d * H:(d + 1) * H]
                    if S["HX16"] is None:
                        pk.gemm_tn(GT16[d], hp, dU, CG, H, TB, lda=ldt, ldb=ldt, ldc=H, alpha_dev=inv, accumulate=(d > 0),
                                   split_k=16)
                    else:  # GRU / minimalGRU: the candidate block contracted (gate * h_{t-1}) (:634, :1295)
                        pk.gemm_tn(GT16[d][:H], S["HX16"][d * H:(d + 1) * H], dU[:H], H, H, TB, lda=ldt, ldb=ldt, ldc=H,
                                   alpha_dev=inv, accumulate=(d > 0), split_k=16)
                        pk.gemm_tn(GT16[d][H:], hp, dU[H:], CG - H, H, TB, lda=ldt, ldb=ldt, ldc=H, alpha_dev=inv,
                                   accumulate=(d > 0), split_k=16)
                # dW = dP^T X
                pk.gemm_tn(dPT16, S["XT16"], dW, CG, D, TB, lda=ldt, ldb=ldt, ldc=D, alpha_dev=inv, split_k=8)
            sl = [slice(g * H, (g + 1) * H) for g in range(ngr)]
            # gradients written in place into param.grad are reported to autograd as None (nothing left to accumulate)
            lg = ([None] * ngr if dW_direct is not None else [dW[s] for s in sl]) + \
                 ([None] * ngr if dU_direct is not None else [dU[s] for s in sl])
            if S["use_bn"]:
                for s in sl:
                    lg += [dgamma[s], dbeta[s]]
            else:
                lg += [dbeta[s] for s in sl]
            grads = lg + grads
            # gradient w.r.t. the layer input
            if li > 0:
                dXT = torch.empty(D, ldt, **f32)
                pk.gemm_tn(S["WT16"], dP16, dXT, D, TB, CG, lda=ldG, ldb=ldG, ldc=ldt, alpha_dev=inv,
                           amax_bits=amax_acc)
                dYT = dXT
            elif ctx.x_needs_grad:
                dx = torch.empty(T, B, D, **f32)
                pk.gemm_tn(dP16, S["WT16"], dx, TB, D, CG, lda=ldG, ldb=ldG, ldc=D, alpha_dev=inv)
        if OVERLAP_WGRAD:
            main.wait_stream(side)
        ctx.saved = None
        return (dx, None, *grads)


class LinearLogSoftmaxFn(torch.autograd.Function):
    """MLP layer `LogSoftmax(W x + b)` — the senone head (neural_networks.py:138-148 with
    dnn_act=softmax, :53-54).  Input [N,F] row-major fp32, output log-posteriors [N,S]."""

    @staticmethod
    def forward(ctx, x, W, b, grad_enabled=True):
        if not x.is_cuda:
            raise RuntimeError("pytorch-kaldi_b200: MLP needs CUDA tensors (there is no CPU fallback)")
        dev = x.device
        N, F = x.shape
        S = W.shape[0]
        f32 = dict(device=dev, dtype=torch.float32)
        f16 = dict(device=dev, dtype=torch.float16)
        need_grad = grad_enabled and any(ctx.needs_input_grad)
        ldF, ldn, ldS = pad8(F), pad8(N), pad8(S)
        x2 = x if (x.dtype == torch.float32 and x.stride(1) == 1) else x.float().contiguous()
        X16 = torch.empty(N, ldF, **f16)
        XT16 = torch.empty(F, ldn, **f16) if (need_grad and ctx.needs_input_grad[1]) else None
        pk.transpose_f32(x2, x2.stride(0), N, F, outT16=XT16, ldo16=ldn, in16=X16, ldi16=ldF)
        Wc = W.contiguous()
        W16 = torch.empty(S, ldF, **f16)
        WT16 = torch.empty(F, ldS, **f16) if (need_grad and ctx.needs_input_grad[0]) else None
        pk.transpose_f32(Wc, F, S, F, outT16=WT16, ldo16=ldS, in16=W16, ldi16=ldF)
        logp = torch.empty(N, S, **f32)
        pk.gemm_tn(X16, W16, logp, N, S, F, lda=ldF, ldb=ldF, ldc=S, bias=b.contiguous() if b is not None else None,
                   bias_mode=1)
        pk.logsoftmax_nll(N, S, logp, S, None, None)
        if need_grad:
            ctx.save_for_backward(logp)
            ctx.aux = (XT16, WT16, N, F, S, b is not None)
        return logp

    @staticmethod
    def backward(ctx, dlogp):
        (logp,) = ctx.saved_tensors
        XT16, WT16, N, F, S, has_bias = ctx.aux
        dev = logp.device
        f32 = dict(device=dev, dtype=torch.float32)
        f16 = dict(device=dev, dtype=torch.float16)
        ldn, ldS = pad8(N), pad8(S)
        dl = dlogp if (dlogp.dtype == torch.float32 and dlogp.is_contiguous()) else dlogp.float().contiguous()
        sc = torch.empty(2, **f32)
        pk.amax_scale(dl, S, N, S, 8.0, torch.empty(1, **f32), sc)
        d16 = torch.empty(N, ldS, **f16) if ctx.needs_input_grad[0] else None
        dT16 = torch.empty(S, ldn, **f16) if ctx.needs_input_grad[1] else None
        db = torch.empty(S, **f32) if (has_bias and ctx.needs_input_grad[2]) else None
        rowsum = torch.empty(N, **f32)
        pk.logsoftmax_bwd(N, S, logp, S, None, dl, S, 1.0, 1.0, sc, d16, ldS, dT16, ldn, db, rowsum)
        inv = sc[1:2]
        dx = dW = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(N, F, **f32)
            pk.gemm_tn(d16, WT16, dx, N, F, S, lda=ldS, ldb=ldS, ldc=F, alpha_dev=inv)
        if ctx.needs_input_grad[1]:
            dW = torch.empty(S, F, **f32)
            pk.gemm_tn(dT16, XT16, dW, S, F, N, lda=ldn, ldb=ldn, ldc=F, alpha_dev=inv, split_k=8)
        return dx, dW, db, None


class HeadNLLFn(torch.autograd.Function):
    """Fused senone head + cost ops: LogSoftmax(W x + b) -> NLLLoss (mean over all rows) and the
    frame error count (utils.py:2344-2381) without a separate pass over the [N,S] posteriors in
    the backward.  Returns (loss, err, logp); `logp` is returned for inspection only (no grad)."""

    @staticmethod
    def forward(ctx, x, W, b, labels):
        dev = x.device
        N, F = x.shape
        S = W.shape[0]
        f32 = dict(device=dev, dtype=torch.float32)
        f16 = dict(device=dev, dtype=torch.float16)
        ldF, ldn, ldS = pad8(F), pad8(N), pad8(S)
        x2 = x if (x.dtype == torch.float32 and x.stride(1) == 1) else x.float().contiguous()
        X16 = torch.empty(N, ldF, **f16)
        XT16 = torch.empty(F, ldn, **f16)
        pk.transpose_f32(x2, x2.stride(0), N, F, outT16=XT16, ldo16=ldn, in16=X16, ldi16=ldF)
        Wc = W.contiguous()
        W16 = torch.empty(S, ldF, **f16)
        WT16 = torch.empty(F, ldS, **f16)
        pk.transpose_f32(Wc, F, S, F, outT16=WT16, ldo16=ldS, in16=W16, ldi16=ldF)
        logp = torch.empty(N, S, **f32)
        pk.gemm_tn(X16, W16, logp, N, S, F, lda=ldF, ldb=ldF, ldc=S, bias=b.contiguous(), bias_mode=1)
        acc = torch.empty(2, device=dev, dtype=torch.float64)
        lab = labels if labels.dtype == torch.int64 else labels.long()
        lab = lab.contiguous()
        pk.logsoftmax_nll(N, S, logp, S, lab, acc)
        loss = (acc[0] / N).float()
        err = (acc[1] / N).float()
        ctx.save_for_backward(logp, lab)
        ctx.aux = (XT16, WT16, N, F, S)
        ctx.wparam = W
        ctx.mark_non_differentiable(err, logp)
        return loss, err, logp

    @staticmethod
    def backward(ctx, dloss, _derr, _dlogp):
        logp, lab = ctx.saved_tensors
        XT16, WT16, N, F, S = ctx.aux
        dev = logp.device
        f32 = dict(device=dev, dtype=torch.float32)
        f16 = dict(device=dev, dtype=torch.float16)
        ldn, ldS = pad8(N), pad8(S)
        # |dlogits| <= |dloss| / N; dloss stays on the device (no host sync): it rides along as the
        # kernel's device-side scale factor, a host-side power of two centres 1/N near 2^8
        gcoef = 1.0 / N
        out_scale = 2.0 ** (8 - math.ceil(math.log2(gcoef)))
        dl = dloss.reshape(1).float().contiguous()
        d16 = torch.empty(N, ldS, **f16)
        dT16 = torch.empty(S, ldn, **f16)
        db = torch.empty(S, **f32)
        pk.logsoftmax_bwd(N, S, logp, S, lab, None, 0, gcoef, out_scale, dl, d16, ldS, dT16, ldn, db, None)
        db = db * dl  # the kernel's bias gradient excludes the device-side factor
        dx = torch.empty(N, F, **f32)
        pk.gemm_tn(d16, WT16, dx, N, F, S, lda=ldS, ldb=ldS, ldc=F, alpha=1.0 / out_scale)
        dW_direct = _direct_grad_target([ctx.wparam])
        dW = dW_direct if dW_direct is not None else torch.empty(S, F, **f32)
        pk.gemm_tn(dT16, XT16, dW, S, F, N, lda=ldn, ldb=ldn, ldc=F, alpha=1.0 / out_scale, split_k=8)
        ctx.wparam = None
        return dx, (None if dW_direct is not None else dW), db, None


@dataclass
class FusionCfg:
    """Static description of a FusionLinearConv group (reference neural_networks.py:2057-2099)."""
    M: int                    # microphone channels concatenated along the feature axis
    mode: int                 # 0: act(x) = x > 0 ? x : slope * x (relu / leaky_relu / prelu), 1: tanh
    red: float                # 1 ("sum") or 1 / M ("mean")
    prelu: bool               # the slope is the nn.PReLU parameter of each gate (trained)
    slope: float = 0.0        # fixed slope otherwise (relu 0, nn.LeakyReLU() 0.01)
    grad_enabled: bool = True


class FusionProjFn(torch.autograd.Function):
    """`reduce_m act(Conv1d(1, H, kernel = d, stride = d)(x))` for G FusionLinearConv modules that read the same input
    (wh and wz of a liGRU_layer, :817-825): ONE tcgen05 GEMM over the zero-copy view [N*M, d] of the input against the
    stacked filters [G*H, d] (+ bias), then the activation + channel reduction kernel; backward = its pointwise
    backward (dO, bias and PReLU-slope gradients) -> the weight-gradient GEMM (and dX when the input needs one).
    Returns [..., G*H] fp32 (gate blocks side by side)."""

    @staticmethod
    def forward(ctx, x, cfg: FusionCfg, *params):
        if not x.is_cuda:
            raise RuntimeError("pytorch-kaldi_b200: FusionLinearConv needs CUDA tensors (there is no CPU fallback)")
        per = 3 if cfg.prelu else 2          # per gate: conv.weight [H,1,d], conv.bias [H] (, prelu.weight [1])
        G = len(params) // per
        ws = [params[per * g] for g in range(G)]
        bs = [params[per * g + 1] for g in range(G)]
        H, d = ws[0].shape[0], ws[0].shape[-1]
        M = cfg.M
        if x.shape[-1] != M * d:
            raise ValueError(f"FusionLinearConv: last dimension {x.shape[-1]} != number_of_mic * in_features = {M} * {d}")
        dev = x.device
        f32 = dict(device=dev, dtype=torch.float32)
        f16 = dict(device=dev, dtype=torch.float16)
        need_grad = cfg.grad_enabled and any(ctx.needs_input_grad)
        lead = x.shape[:-1]
        N = int(math.prod(lead))
        rows, C = N * M, G * H
        xr = x.reshape(rows, d)
        if xr.dtype != torch.float32 or not xr.is_contiguous():
            xr = xr.float().contiguous()
        ldd, ldr, ldC = pad8(d), pad8(rows), pad8(C)
        X16 = torch.empty(rows, ldd, **f16)
        XT16 = torch.empty(d, ldr, **f16) if need_grad else None
        pk.transpose_f32(xr, d, rows, d, outT16=XT16, ldo16=ldr, in16=X16, ldi16=ldd)
        Wcat = torch.cat([w.reshape(H, d) for w in ws]).contiguous()
        W16 = torch.empty(C, ldd, **f16)
        WT16 = torch.empty(d, ldC, **f16) if (need_grad and ctx.needs_input_grad[0]) else None
        pk.transpose_f32(Wcat, d, C, d, outT16=WT16, ldo16=ldC, in16=W16, ldi16=ldd)
        if cfg.prelu:
            slopes = torch.cat([params[per * g + 2].reshape(1) for g in range(G)]).float().contiguous()
        else:
            slopes = torch.full((G,), float(cfg.slope), **f32)
        O = torch.empty(rows, C, **f32)
        pk.gemm_tn(X16, W16, O, rows, C, d, lda=ldd, ldb=ldd, ldc=C, bias=torch.cat(bs).contiguous(), bias_mode=1)
        P = torch.empty(N, C, **f32)
        pk.fusion_reduce_fwd(O, C, N, M, C, H, cfg.mode, slopes, cfg.red, P, C)
        if need_grad:
            ctx.saved = (O, XT16, WT16, slopes)
            ctx.dims = (N, M, d, H, G, per, tuple(x.shape), [tuple(w.shape) for w in ws])
            ctx.cfg = cfg
        return P.view(*lead, C)

    @staticmethod
    def backward(ctx, dP):
        O, XT16, WT16, slopes = ctx.saved
        N, M, d, H, G, per, xshape, wshapes = ctx.dims
        cfg = ctx.cfg
        dev = dP.device
        f32 = dict(device=dev, dtype=torch.float32)
        f16 = dict(device=dev, dtype=torch.float16)
        rows, C = N * M, G * H
        ldr, ldC = pad8(rows), pad8(C)
        dp2 = dP.reshape(N, C)
        if dp2.dtype != torch.float32 or not dp2.is_contiguous():
            dp2 = dp2.float().contiguous()
        dO = torch.empty(rows, C, **f32)
        dbias = torch.zeros(C, **f32)
        dsl = torch.zeros(G, **f32)
        pk.fusion_reduce_bwd(dp2, C, O, C, N, M, C, H, cfg.mode, slopes, cfg.red, dO, C, dbias, dsl)
        sc = torch.empty(2, **f32)
        pk.amax_scale(dO, C, rows, C, 8.0, torch.empty(1, **f32), sc)
        inv = sc[1:2]
        need_dx = ctx.needs_input_grad[0]
        dOT16 = torch.empty(C, ldr, **f16)
        d16 = torch.empty(rows, ldC, **f16) if need_dx else None
        pk.transpose_f32(dO, C, rows, C, outT16=dOT16, ldo16=ldr, in16=d16, ldi16=ldC, scale_dev=sc)
        dW = torch.empty(C, d, **f32)
        pk.gemm_tn(dOT16, XT16, dW, C, d, rows, lda=ldr, ldb=ldr, ldc=d, alpha_dev=inv, split_k=8 if rows >= 4096 else 1)
        dx = None
        if need_dx:
            dxr = torch.empty(rows, d, **f32)
            pk.gemm_tn(d16, WT16, dxr, rows, d, C, lda=ldC, ldb=ldC, ldc=d, alpha_dev=inv)
            dx = dxr.view(xshape)
        grads = []
        for g in range(G):
            grads += [dW[g * H:(g + 1) * H].reshape(wshapes[g]), dbias[g * H:(g + 1) * H]]
            if cfg.prelu:
                grads.append(dsl[g:g + 1])
        ctx.saved = None
        return (dx, None, *grads)


@dataclass
class DenseLayerCfg:
    O: int
    act: str                      # activation name; "softmax" only on the last layer
    use_bn: bool
    bn_training: bool
    bn: Optional[torch.nn.Module] = None
    keepT: Optional[torch.Tensor] = None   # fp16 [O, pad8(N)] with 0 or 1/(1-p) (training dropout) or None
    grad_enabled: bool = True              # torch.is_grad_enabled() at call time
    use_ln: bool = False                   # reference LayerNorm between the Linear and BatchNorm / act (:129-145)
    ln_eps: float = 1e-6


class MLPStackFn(torch.autograd.Function):
    """MLP stack `drop(act(bn(W x + b)))` per layer (reference neural_networks.py:130-150); a final softmax
    layer is the fused linear + LogSoftmax.  Hidden layers run channel-major like the recurrent projections:
    GEMM (+bias, +BatchNorm statistics in the epilogue) -> folded scale/shift -> fused act/dropout epilogue that
    emits the fp16 operands of the next layer; backward = act/dropout backward -> the same BatchNorm-backward
    kernel -> dW / dX GEMMs."""

    @staticmethod
    def forward(ctx, x, layers, *params):
        if not x.is_cuda:
            raise RuntimeError("pytorch-kaldi_b200: MLP needs CUDA tensors (there is no CPU fallback)")
        dev = x.device
        N, I0 = x.shape
        ldn = pad8(N)
        f32 = dict(device=dev, dtype=torch.float32)
        f16 = dict(device=dev, dtype=torch.float16)
        need_grad = layers[0].grad_enabled and any(ctx.needs_input_grad)
        x2 = x if (x.dtype == torch.float32 and x.stride(1) == 1) else x.float().contiguous()
        I = I0
        X16 = torch.empty(N, pad8(I), **f16)
        XT16 = torch.empty(I, ldn, **f16) if need_grad else None
        pk.transpose_f32(x2, x2.stride(0), N, I, outT16=XT16, ldo16=ldn, in16=X16, ldi16=pad8(I))
        saved, pi, out = [], 0, None
        for li, L in enumerate(layers):
            O = L.O
            W, b = params[pi], params[pi + 1]
            pi += 2
            if L.use_bn:
                gam, bet = params[pi], params[pi + 1]
                pi += 2
            if L.use_ln:
                lgam, lbet = params[pi], params[pi + 1]
                pi += 2
            ldI, ldO = pad8(I), pad8(O)
            W16 = torch.empty(O, ldI, **f16)
            WT16 = torch.empty(I, ldO, **f16) if need_grad else None
            pk.transpose_f32(W.contiguous(), I, O, I, outT16=WT16, ldo16=ldO, in16=W16, ldi16=ldI)
            last = li == len(layers) - 1
            if L.act == "softmax":
                if not last or L.use_bn or L.use_ln or L.keepT is not None:
                    raise NotImplementedError("softmax is only supported as the plain last MLP layer")
                logp = torch.empty(N, O, **f32)
                pk.gemm_tn(X16, W16, logp, N, O, I, lda=ldI, ldb=ldI, ldc=O, bias=b.contiguous(), bias_mode=1)
                pk.logsoftmax_nll(N, O, logp, O, None, None)
                out = logp
                if need_grad:
                    saved.append(dict(kind="softmax", I=I, O=O, XT16=XT16, WT16=WT16, logp=logp))
                break
            bn_train = L.use_bn and L.bn_training
            PT = torch.empty(O, ldn, **f32)
            stats = torch.zeros(O, 2, device=dev, dtype=torch.float64) if bn_train else None
            pk.gemm_tn(W16, X16, PT, O, N, I, lda=ldI, ldb=ldI, ldc=ldn, bias=b.contiguous(), bias_mode=2,
                       rowstats=None if L.use_ln else stats)
            XH = lnstats = None
            if L.use_ln:  # y = gamma (x - mean_features) / (std + eps) + beta per frame, in place (:23-33)
                XH = torch.empty(O, ldn, **f32) if need_grad else None
                lnstats = torch.empty(N, 2, **f32)
                pk.ln_cm_fwd(PT, O, N, ldn, lgam.contiguous(), lbet.contiguous(), L.ln_eps, XH, lnstats)
                if bn_train:   # BatchNorm over the LayerNorm output: its batch statistics need their own pass
                    pk.row_stats(PT, O, N, ldn, stats)
            scale, shift = torch.empty(O, **f32), torch.empty(O, **f32)
            mean = rstd = None
            if L.use_bn:
                mean, rstd = torch.empty(O, **f32), torch.empty(O, **f32)
                bn = L.bn
                pk.bn_finalize(stats, O, N, N, gam, bet, bn.eps, bn.momentum if bn.momentum is not None else 0.1,
                               bn_train, bn.running_mean, bn.running_var,
                               bn.num_batches_tracked if bn_train else None, scale, shift, mean, rstd)
            else:
                pk.fill_scale_shift(None, O, scale, shift)
            YT16 = torch.empty(O, ldn, **f16) if need_grad else None
            Y16 = None if last else torch.empty(N, ldO, **f16)
            if last:
                out = torch.empty(N, O, **f32)
            pk.dense_act_fwd(O, N, pk.ACT_IDS[L.act], PT, ldn, scale, shift, L.keepT, ldn, YT16, ldn, Y16, ldO,
                             out if last else None, O)
            if need_grad:
                saved.append(dict(kind="dense", I=I, O=O, XT16=XT16, WT16=WT16, PT=PT if bn_train else None, mean=mean,
                                  rstd=rstd, gamma=gam if L.use_bn else None, YT16=YT16, keepT=L.keepT,
                                  act=pk.ACT_IDS[L.act], use_bn=L.use_bn, bn_train=bn_train, use_ln=L.use_ln,
                                  XH=XH, lnstats=lnstats, lgam=lgam.contiguous() if L.use_ln else None, ln_eps=L.ln_eps))
            X16, XT16, I = Y16, YT16, O
        ctx.saved = saved
        ctx.dims = (N, I0, ldn)
        ctx.x_needs_grad = ctx.needs_input_grad[0]
        return out

    @staticmethod
    def backward(ctx, dOut):
        saved = ctx.saved
        N, I0, ldn = ctx.dims
        dev = dOut.device
        f32 = dict(device=dev, dtype=torch.float32)
        f16 = dict(device=dev, dtype=torch.float16)
        amax_acc = torch.zeros(1, device=dev, dtype=torch.int32)
        grads = []
        dYT = None
        dx = None
        for li in reversed(range(len(saved))):
            S = saved[li]
            I, O = S["I"], S["O"]
            ldO = pad8(O)
            need_dx = li > 0 or ctx.x_needs_grad
            sc = torch.empty(2, **f32)
            if S["kind"] == "softmax":
                dl = dOut if (dOut.dtype == torch.float32 and dOut.is_contiguous()) else dOut.float().contiguous()
                pk.amax_scale(dl, O, N, O, 8.0, torch.empty(1, **f32), sc)
                d16 = torch.empty(N, ldO, **f16)
                dT16 = torch.empty(O, ldn, **f16)
                db = torch.empty(O, **f32)
                pk.logsoftmax_bwd(N, O, S["logp"], O, None, dl, O, 1.0, 1.0, sc, d16, ldO, dT16, ldn, db,
                                  torch.empty(N, **f32))
                dPT16, dP16 = dT16, d16
                lg = []
            else:
                if dYT is None:  # this dense layer is the module output: autograd gives the row-major gradient
                    dy2 = dOut if (dOut.dtype == torch.float32 and dOut.is_contiguous()) else dOut.float().contiguous()
                    dYT = torch.empty(O, ldn, **f32)
                    pk.transpose_f32(dy2, O, N, O, outT=dYT, ldo=ldn, amax_bits=amax_acc)
                pk.amax_finalize(amax_acc, 8.0, sc)
                GT16 = torch.empty(O, ldn, **f16)
                pk.dense_act_bwd(O, N, S["act"], dYT, ldn, S["YT16"], ldn, S["keepT"], ldn, sc, GT16, ldn)
                dgamma, dbeta = torch.empty(O, **f32), torch.empty(O, **f32)
                dPT16 = torch.empty(O, ldn, **f16)
                dP16 = torch.empty(N, ldO, **f16) if need_dx else None
                pk.bn_bwd(O, 1, N, None, GT16, ldn, S["PT"], ldn, S["use_bn"], S["bn_train"], S["mean"], S["rstd"],
                          S["gamma"], sc, dgamma, dbeta, dPT16, ldn, dP16, ldO,
                          torch.empty(2 * O, device=dev, dtype=torch.float64))
                # a bias in front of BatchNorm has a mathematically zero gradient (reference: rounding noise)
                db = torch.zeros(O, **f32) if S["use_bn"] else dbeta
                lg = [dgamma, dbeta] if S["use_bn"] else []
                if S["use_ln"]:  # LayerNorm backward turns the gradient w.r.t. its output into that w.r.t. the Linear output
                    dlg, dlb, db = torch.empty(O, **f32), torch.empty(O, **f32), torch.empty(O, **f32)
                    pk.ln_cm_bwd(dPT16, ldn, dP16, ldO, S["XH"], ldn, O, N, S["lgam"], S["lnstats"], S["ln_eps"], sc,
                                 dlg, dlb, db)
                    lg = lg + [dlg, dlb]
            inv = sc[1:2]
            dW = torch.empty(O, I, **f32)
            pk.gemm_tn(dPT16, S["XT16"], dW, O, I, N, lda=ldn, ldb=ldn, ldc=I, alpha_dev=inv,
                       split_k=8 if N >= 4096 else 1)
            grads = [dW, db] + lg + grads
            if li > 0:
                dXT = torch.empty(I, ldn, **f32)
                pk.gemm_tn(S["WT16"], dP16, dXT, I, N, O, lda=ldO, ldb=ldO, ldc=ldn, alpha_dev=inv, amax_bits=amax_acc)
                dYT = dXT
            elif ctx.x_needs_grad:
                dx = torch.empty(N, I, **f32)
                pk.gemm_tn(dP16, S["WT16"], dx, N, I, O, lda=ldO, ldb=ldO, ldc=I, alpha_dev=inv)
        ctx.saved = None
        return (dx, None, *grads)


def mlp_forward(module, x):
    """neural_networks.MLP.forward for general stacks: builds the static layer description and calls MLPStackFn."""
    if module.dnn_use_laynorm_inp or module.dnn_use_batchnorm_inp:
        raise NotImplementedError(
            "pytorch-kaldi_b200.MLP: dnn_use_laynorm_inp / dnn_use_batchnorm_inp (normalisation of the INPUT) are not "
            "implemented natively yet (no shipped recipe enables them); there is no eager fallback")
    N = x.shape[0]
    ldn = pad8(N)
    layers, params = [], []
    for i, O in enumerate(module.dnn_lay):
        p = module.dnn_drop[i]
        keepT = None
        if module.training and p > 0.0:
            override = getattr(module, "_keep_override", None)
            if override is not None and override[i] is not None:
                keepT = torch.zeros(O, ldn, device=x.device, dtype=torch.float16)
                keepT[:, :N] = (override[i].to(x.device).t().float() / (1.0 - p)).half()
            else:
                keepT = ((torch.rand(O, ldn, device=x.device) >= p).half() / (1.0 - p)).contiguous()
        use_bn = bool(module.dnn_use_batchnorm[i])
        use_ln = bool(module.dnn_use_laynorm[i])
        layers.append(DenseLayerCfg(O=O, act=module.dnn_act[i], use_bn=use_bn, bn_training=module.training,
                                    bn=module.bn[i] if use_bn else None, keepT=keepT,
                                    grad_enabled=torch.is_grad_enabled(), use_ln=use_ln, ln_eps=module.ln[i].eps))
        params += [module.wx[i].weight, module.wx[i].bias]
        if use_bn:
            params += [module.bn[i].weight, module.bn[i].bias]
        if use_ln:
            params += [module.ln[i].gamma, module.ln[i].beta]
    return MLPStackFn.apply(x, layers, *params)


# ---------------------------------------------------------------------------------------------
# conv front-ends: CNN (neural_networks.py:1464-1556), SincNet (:1559-1665), SincConv (:1668-1813)
# ---------------------------------------------------------------------------------------------


@dataclass
class ConvLayerCfg:
    kind: str                  # "sinc" (band-pass filters synthesised from low_hz_/band_hz_) or "conv"
    C: int                     # output channels
    k: int                     # taps
    pool: int                  # max_pool1d length
    act: int
    use_ln: bool               # LayerNorm over the length axis with a [C, Lp] affine (:1505-1507)
    ln_eps: float = 1e-6
    keep16: Optional[torch.Tensor] = None   # [N, Lp, C] fp16: 0 or 1/(1-p) (training dropout) or None
    sample_rate: float = 16000.0
    min_low_hz: float = 50.0
    min_band_hz: float = 50.0


@dataclass
class ConvStackCfg:
    layers: List[ConvLayerCfg] = field(default_factory=list)
    ln0: bool = False
    ln0_eps: float = 1e-6
    flat_output: bool = True   # CNN/SincNet return x.view(batch, -1); a bare SincConv returns [N, C, Lout]
    grad_enabled: bool = True  # torch.is_grad_enabled() at call time


class ConvStackFn(torch.autograd.Function):
    """`drop(act(ln(max_pool1d(conv(x)))))` stacks.  Activations are position-major fp16 `[n][l][c]`; every
    convolution (and its input gradient) is ONE tcgen05 GEMM over an overlapping-row view of the activation
    buffer (csrc/pk_conv.cu), the first layer (one input channel) uses an explicit fp16 im2col."""

    @staticmethod
    def forward(ctx, x, cfg: ConvStackCfg, *params):
        if not x.is_cuda:
            raise RuntimeError("pytorch-kaldi_b200: CNN/SincNet need CUDA tensors (there is no CPU fallback)")
        dev = x.device
        N, L0 = x.shape
        f32 = dict(device=dev, dtype=torch.float32)
        f16 = dict(device=dev, dtype=torch.float16)
        need_grad = cfg.grad_enabled and any(ctx.needs_input_grad)
        if need_grad and ctx.needs_input_grad[0]:
            # the gradient w.r.t. the waveform (a trainable module IN FRONT of the CNN / SincNet / SincConv, e.g. a
            # joint-training enhancement front-end) is not implemented: fail loudly instead of handing autograd a
            # silent zero
            raise NotImplementedError("pytorch-kaldi_b200: CNN / SincNet / SincConv do not produce the gradient w.r.t. their "
                                      "input yet; detach the input or keep trainable modules behind the front-end")
        xr = x if (x.dtype == torch.float32 and x.stride(1) == 1) else x.float().contiguous()
        pi = 0
        ln0_saved = None
        X0, ldx0 = xr, xr.stride(0)
        if cfg.ln0:
            g0, b0 = params[0], params[1]
            pi = 2
            X0 = torch.empty(N, L0, **f32)
            st0 = torch.empty(N, 2, **f32)
            pk.rowln_fwd(xr, xr.stride(0), N, L0, g0, b0, cfg.ln0_eps, X0, st0)
            ldx0 = L0
            ln0_saved = (xr, st0)
        saved = []
        L, Ci, Cp = L0, 1, 1
        A16 = None
        y32 = None
        nl = len(cfg.layers)
        for li, Lc in enumerate(cfg.layers):
            C, k, p = Lc.C, Lc.k, Lc.pool
            Lout = L - k + 1
            Lp = Lout // p
            if Lout < 1 or Lp < 1:
                raise ValueError(f"conv layer {li}: input length {L} too short for kernel {k} / pool {p}")
            rows = N * L
            if Lc.kind == "sinc":
                low, band = params[pi], params[pi + 1]
                pi += 2
                w = torch.empty(C, 1, k, **f32)
                pk.sinc_filters_fwd(low, band, C, k, Lc.sample_rate, Lc.min_low_hz, Lc.min_band_hz, w)
                bias = None
            else:
                w, bias = params[pi].contiguous(), params[pi + 1]
                pi += 2
            gamma = beta = None
            if Lc.use_ln:
                gamma, beta = params[pi].contiguous(), params[pi + 1].contiguous()
                pi += 2
            O = torch.empty(rows, C, **f32)
            XcolT = None
            if li == 0:
                Kp = pad8(k)
                Xcol = torch.empty(rows, Kp, **f16)
                ldp = pad8(rows)
                XcolT = torch.empty(k, ldp, **f16) if need_grad else None
                pk.conv_im2col0(X0, ldx0, N, L, k, Lout, Xcol, Kp, XcolT, ldp)
                W16 = torch.zeros(C, Kp, **f16)
                pk.conv_pack_weights(w, C, 1, k, W16, 1, Kp, None, 0, 0)
                pk.gemm_tn(Xcol, W16, O, rows, C, Kp, lda=Kp, ldb=Kp, ldc=C, bias=bias, bias_mode=1)
            else:
                K = k * Cp
                W16 = torch.zeros(C, K, **f16)
                pk.conv_pack_weights(w, C, Ci, k, W16, Cp, K, None, 0, 0)
                pk.gemm_tn(A16, W16, O, rows, C, K, lda=Cp, ldb=K, ldc=C, bias=bias, bias_mode=1)
            last = li == nl - 1
            P = torch.empty(N, Lp, C, **f32)
            arg = torch.empty(N, Lp, C, device=dev, dtype=torch.uint8)
            stats = torch.empty(N, C, 2, **f32) if Lc.use_ln else None
            A16n, Cpn = None, 0
            if not last:
                Cpn = pad8(C)
                # tail rows: the next layer's overlapping im2col rows of the last frame read past the end
                A16n = torch.zeros(N * Lp + cfg.layers[li + 1].k, Cpn, **f16)
            else:
                y32 = torch.empty(N, C, Lp, **f32)
            pk.conv_post_fwd(O, C, N, L, Lout, p, Lp, C, Lc.act, gamma, beta, Lc.ln_eps, Lc.keep16, P, arg, stats, A16n,
                             Cpn, y32)
            if need_grad:
                saved.append(dict(cfg=Lc, L=L, Lout=Lout, Lp=Lp, Ci=Ci, Cp=Cp, A16=A16, XcolT=XcolT, w=w, gamma=gamma,
                                  beta=beta, P=P, arg=arg, stats=stats,
                                  low=low if Lc.kind == "sinc" else None, band=band if Lc.kind == "sinc" else None))
            A16, L, Ci, Cp = A16n, Lp, C, Cpn
        ctx.cfg, ctx.saved, ctx.ln0_saved, ctx.N, ctx.L0 = cfg, saved, ln0_saved, N, L0
        return y32.view(N, -1) if cfg.flat_output else y32

    @staticmethod
    def backward(ctx, dY):
        cfg, saved, N, L0 = ctx.cfg, ctx.saved, ctx.N, ctx.L0
        dev = dY.device
        f32 = dict(device=dev, dtype=torch.float32)
        f16 = dict(device=dev, dtype=torch.float16)
        amax_acc = torch.zeros(1, device=dev, dtype=torch.int32)
        grads_rev = []
        g0 = []
        dA = None
        for li in reversed(range(len(saved))):
            S = saved[li]
            Lc = S["cfg"]
            C, k, p, L, Lout, Lp, Ci = Lc.C, Lc.k, Lc.pool, S["L"], S["Lout"], S["Lp"], S["Ci"]
            rows = N * L
            if dA is None:  # module output gradient, [N, C, Lp] order
                dy = dY.reshape(N, C, Lp)
                if dy.dtype != torch.float32 or not dy.is_contiguous():
                    dy = dy.float().contiguous()
                sn, sl, sc_ = C * Lp, 1, Lp
            else:           # dX of the next layer: position-major [N*Lp, C]
                dy, sn, sl, sc_ = dA, Lp * C, C, 1
            dO = torch.empty(rows, C, **f32)
            dgamma = torch.empty(C, Lp, **f32) if Lc.use_ln else None
            dbeta = torch.empty(C, Lp, **f32) if Lc.use_ln else None
            dbias = torch.empty(C, **f32) if Lc.kind == "conv" else None
            pk.conv_post_bwd(dy, sn, sl, sc_, N, L, Lout, p, Lp, C, Lc.act, S["gamma"], S["beta"], Lc.ln_eps, Lc.keep16,
                             S["P"], S["arg"], S["stats"], dgamma, dbeta, dbias, dO, amax_acc)
            sc = torch.empty(2, **f32)
            pk.amax_finalize(amax_acc, 8.0, sc)
            inv = sc[1:2]
            Cop = pad8(C)
            ldp = pad8(rows)
            dO16 = torch.zeros(k - 1 + rows, Cop, **f16)   # k-1 zero rows in front: the dX GEMM reads positions r-k+1..r
            dOT16 = torch.empty(C, ldp, **f16)
            pk.transpose_f32(dO, C, rows, C, outT16=dOT16, ldo16=ldp, in16=dO16[k - 1:], ldi16=Cop, scale_dev=sc)
            need_dx = li > 0
            if li == 0:
                dF = torch.empty(C, k, **f32)
                pk.gemm_tn(dOT16, S["XcolT"], dF, C, k, rows, lda=ldp, ldb=ldp, ldc=k, alpha_dev=inv, split_k=8)
                if Lc.kind == "sinc":
                    dlow = torch.empty(C, 1, **f32)
                    dband = torch.empty(C, 1, **f32)
                    pk.sinc_filters_bwd(S["low"], S["band"], C, k, Lc.sample_rate, Lc.min_low_hz, Lc.min_band_hz, dF, dlow,
                                        dband)
                    lg = [dlow, dband]
                else:
                    lg = [dF.view(C, 1, k), dbias]
                if cfg.ln0:
                    Wf = torch.zeros(1, k * Cop, **f16)
                    pk.conv_pack_weights(S["w"], C, 1, k, None, 0, 0, Wf, Cop, k * Cop)
                    G = torch.empty(rows, k, **f32)
                    pk.gemm_tn(dO16[k - 1:], Wf, G, rows, k, Cop, lda=Cop, ldb=Cop, ldc=k, alpha_dev=inv)
                    xraw, st0 = ctx.ln0_saved
                    dg0 = torch.empty(L0, **f32)
                    db0 = torch.empty(L0, **f32)
                    pk.conv_ln0_bwd(G, k, N, L0, Lout, k, xraw, xraw.stride(0), st0, dg0, db0)
                    g0 = [dg0, db0]
            else:
                Cp = S["Cp"]
                XT = torch.empty(k * Ci, ldp, **f16)
                pk.conv_im2col_t(S["A16"], rows, Cp, Ci, k, XT, ldp)
                dWk = torch.empty(C, k * Ci, **f32)
                pk.gemm_tn(dOT16, XT, dWk, C, k * Ci, rows, lda=ldp, ldb=ldp, ldc=k * Ci, alpha_dev=inv, split_k=8)
                lg = [dWk.view(C, k, Ci).permute(0, 2, 1).contiguous(), dbias]
            if Lc.use_ln:
                lg += [dgamma, dbeta]
            grads_rev.append(lg)
            if need_dx:
                Wf = torch.zeros(Ci, k * Cop, **f16)
                pk.conv_pack_weights(S["w"], C, Ci, k, None, 0, 0, Wf, Cop, k * Cop)
                dA = torch.empty(rows, Ci, **f32)
                pk.gemm_tn(dO16, Wf, dA, rows, Ci, k * Cop, lda=Cop, ldb=k * Cop, ldc=Ci, alpha_dev=inv)
        grads = list(g0)
        for lg in reversed(grads_rev):
            grads += lg
        ctx.saved = None
        return (None, None, *grads)


def conv_forward(module, x, prefix):
    """CNN.forward / SincNet.forward (neural_networks.py:1530-1556, :1638-1665): builds the static description
    and calls ConvStackFn.  `prefix` = "cnn" or "sinc" (option-name prefix of the reference class)."""
    g = lambda name: getattr(module, f"{prefix}_{name}")
    if g("use_batchnorm_inp") or any(g("use_batchnorm")):
        raise NotImplementedError(
            f"pytorch-kaldi_b200.{type(module).__name__}: {prefix}_use_batchnorm / {prefix}_use_batchnorm_inp are not "
            "implemented natively yet (the shipped CNN / SincNet recipes use LayerNorm); there is no eager fallback")
    N, L0 = x.shape
    cfg = ConvStackCfg(ln0=bool(g("use_laynorm_inp")), grad_enabled=torch.is_grad_enabled())
    params = []
    if cfg.ln0:
        cfg.ln0_eps = module.ln0.eps
        params += [module.ln0.gamma, module.ln0.beta]
    L = L0
    n_lay = len(g("N_filt"))
    for i in range(n_lay):
        conv = module.conv[i]
        sinc = hasattr(conv, "low_hz_")
        k = conv.kernel_size if sinc else conv.kernel_size[0]
        C = conv.out_channels
        pool = g("max_pool_len")[i]
        Lp = (L - k + 1) // pool
        pdrop = g("drop")[i]
        keep16 = None
        if module.training and pdrop > 0.0:
            override = getattr(module, "_keep_override", None)
            if override is not None and override[i] is not None:   # tests: the reference's own mask, [N, C, Lp]
                keep16 = (override[i].to(x.device).permute(0, 2, 1).float() / (1.0 - pdrop)).half().contiguous()
            else:
                keep16 = torch.empty(N, Lp, C, device=x.device, dtype=torch.float16).bernoulli_(1.0 - pdrop)
                keep16.mul_(1.0 / (1.0 - pdrop))
        use_ln = bool(g("use_laynorm")[i])
        act_name = g("act")[i]
        if act_name not in pk.ACT_IDS:
            raise NotImplementedError(f"activation {act_name!r} is not valid inside a conv layer")
        lc = ConvLayerCfg(kind="sinc" if sinc else "conv", C=C, k=k, pool=pool, act=pk.ACT_IDS[act_name], use_ln=use_ln,
                          ln_eps=module.ln[i].eps, keep16=keep16)
        if sinc:
            lc.sample_rate, lc.min_low_hz, lc.min_band_hz = conv.sample_rate, conv.min_low_hz, conv.min_band_hz
            params += [conv.low_hz_, conv.band_hz_]
        else:
            params += [conv.weight, conv.bias]
        if use_ln:
            params += [module.ln[i].gamma, module.ln[i].beta]
        cfg.layers.append(lc)
        L = Lp
    return ConvStackFn.apply(x, cfg, *params)
