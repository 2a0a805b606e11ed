"""ctypes binding of libpk_b200.so (the C ABI declared in include/pk_b200.h).

PyTorch is used only for device memory and streams: every wrapper takes torch tensors, checks
device / dtype / contiguity, and passes raw device pointers + the current CUDA stream to the
library.  There is NO fallback: if the shared library is missing or a call fails, a
RuntimeError is raised (the product path must fail loudly, never silently run on the CPU).
"""
from __future__ import annotations

import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpk_b200.so")

F16, TF32 = 0, 2
ACT_IDS = {"relu": 0, "tanh": 1, "sigmoid": 2, "leaky_relu": 3, "elu": 4, "linear": 5}
CELL_LIGRU, CELL_RNN, CELL_GRU, CELL_MGRU, CELL_LSTM = 0, 1, 2, 3, 4
REC_BWD_ALLGATHER, REC_BWD_KSPLIT = 0x800000, 0x1000000  # backward mma.sync kernel: exchange formulation
REC_WS, REC_TC = 0x8000, 0x400000  # force the mma.sync / tcgen05 persistent kernels (default: faster one for this H)

_c_int, _c_i64, _c_f, _c_p = ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p

# name -> argtypes (mirrors include/pk_b200.h; tests check that every symbol resolves)
SIGNATURES = {
    "pk_gemm_tn": [_c_int, _c_int, _c_int, _c_int, _c_p, _c_i64, _c_i64, _c_i64, _c_p, _c_i64, _c_i64, _c_i64, _c_p,
                   _c_i64, _c_p, _c_int, _c_p, _c_f, _c_p, _c_int, _c_int, _c_p, _c_p],
    "pk_transpose_f32": [_c_p, _c_i64, _c_int, _c_int, _c_p, _c_i64, _c_p, _c_i64, _c_p, _c_i64, _c_p, _c_p, _c_p],
    "pk_amax_finalize": [_c_p, _c_f, _c_p, _c_p],
    "pk_convert_f16": [_c_p, _c_i64, _c_int, _c_int, _c_p, _c_i64, _c_p, _c_p],
    "pk_amax_scale": [_c_p, _c_i64, _c_int, _c_int, _c_f, _c_p, _c_p, _c_p],
    "pk_bn_finalize": [_c_p, _c_int, _c_i64, _c_i64, _c_p, _c_p, _c_f, _c_f, _c_int, _c_p, _c_p, _c_p, _c_p, _c_p,
                       _c_p, _c_p, _c_p],
    "pk_fill_scale_shift": [_c_p, _c_int, _c_p, _c_p, _c_p],
    "pk_bn_bwd": [_c_int, _c_int, _c_i64, _c_p, _c_p, _c_i64, _c_p, _c_i64, _c_int, _c_int, _c_p, _c_p, _c_p, _c_p, _c_p,
                  _c_p, _c_p, _c_i64, _c_p, _c_i64, _c_p, _c_p],
    "pk_rnn_layer_fwd": [_c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_p, _c_i64, _c_p, _c_p, _c_p, _c_p, _c_f,
                         _c_p, _c_i64, _c_p, _c_i64, _c_p, _c_p, _c_p, _c_p, _c_p, _c_i64, _c_p],
    "pk_rnn_layer_bwd": [_c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_p, _c_p, _c_p, _c_p, _c_i64, _c_p, _c_p,
                         _c_f, _c_p, _c_p, _c_p, _c_p],
    "pk_rnn_step_workspace_bytes": [_c_int, _c_int, _c_int, _c_int, _c_int, _c_int],
    "pk_rnn_step_launches": [_c_int, _c_int, _c_int, _c_int, _c_int, _c_int],
    "pk_rnn_step_is_cluster": [_c_int, _c_int],
    "pk_rnn_step_fwd": [_c_int] * 6 + [_c_p, _c_i64, _c_p, _c_p, _c_p, _c_p, _c_f, _c_p, _c_i64, _c_p, _c_i64, _c_p, _c_p,
                                       _c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _c_i64, _c_p, _c_i64, _c_p],
    "pk_rnn_step_bwd": [_c_int] * 6 + [_c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _c_i64, _c_p, _c_p, _c_f, _c_p, _c_p,
                                       _c_p, _c_i64, _c_p],
    "pk_fusion_reduce_fwd": [_c_p, _c_i64, _c_i64, _c_int, _c_int, _c_int, _c_int, _c_p, _c_f, _c_p, _c_i64, _c_p],
    "pk_fusion_reduce_bwd": [_c_p, _c_i64, _c_p, _c_i64, _c_i64, _c_int, _c_int, _c_int, _c_int, _c_p, _c_f, _c_p, _c_i64,
                             _c_p, _c_p, _c_p],
    "pk_rowln_fwd": [_c_p, _c_i64, _c_int, _c_int, _c_p, _c_p, _c_f, _c_p, _c_p, _c_p],
    "pk_conv_ln0_bwd": [_c_p, _c_i64, _c_int, _c_int, _c_int, _c_int, _c_p, _c_i64, _c_p, _c_p, _c_p, _c_p],
    "pk_sinc_filters_fwd": [_c_p, _c_p, _c_int, _c_int, _c_f, _c_f, _c_f, _c_p, _c_p],
    "pk_sinc_filters_bwd": [_c_p, _c_p, _c_int, _c_int, _c_f, _c_f, _c_f, _c_p, _c_p, _c_p, _c_p],
    "pk_conv_pack_weights": [_c_p, _c_int, _c_int, _c_int, _c_p, _c_int, _c_i64, _c_p, _c_int, _c_i64, _c_p],
    "pk_conv_im2col0": [_c_p, _c_i64, _c_int, _c_int, _c_int, _c_int, _c_p, _c_int, _c_p, _c_i64, _c_p],
    "pk_conv_im2col_t": [_c_p, _c_i64, _c_int, _c_int, _c_int, _c_p, _c_i64, _c_p],
    "pk_conv_post_fwd": [_c_p, _c_i64] + [_c_int] * 7 + [_c_p, _c_p, _c_f, _c_p, _c_p, _c_p, _c_p, _c_p, _c_int, _c_p, _c_p],
    "pk_conv_post_bwd": [_c_p, _c_i64, _c_i64, _c_i64] + [_c_int] * 7 + [_c_p, _c_p, _c_f] + [_c_p] * 10,
    "pk_logsoftmax_nll": [_c_int, _c_int, _c_p, _c_i64, _c_p, _c_p, _c_p],
    "pk_logsoftmax_bwd": [_c_int, _c_int, _c_p, _c_i64, _c_p, _c_p, _c_i64, _c_f, _c_f, _c_p, _c_p, _c_i64, _c_p,
                          _c_i64, _c_p, _c_p, _c_p],
    "pk_dense_act_fwd": [_c_int, _c_i64, _c_int, _c_p, _c_i64, _c_p, _c_p, _c_p, _c_i64, _c_p, _c_i64, _c_p, _c_i64,
                         _c_p, _c_i64, _c_p],
    "pk_dense_act_bwd": [_c_int, _c_i64, _c_int, _c_p, _c_i64, _c_p, _c_i64, _c_p, _c_i64, _c_p, _c_p, _c_i64, _c_p],
    "pk_rmsprop_step": [_c_p, _c_p, _c_p, _c_i64, _c_f, _c_f, _c_f, _c_f, _c_p],
    "pk_chunk_prepare": [_c_p, _c_i64, _c_p, _c_i64, _c_i64, _c_int, _c_int, _c_int, _c_p, _c_p, _c_i64, _c_p],
    "pk_batch_assemble": [_c_p, _c_i64, _c_int, _c_p, _c_int, _c_int, _c_p, _c_p],
    "pk_cm_decode": [_c_p, _c_p, _c_f, _c_f, _c_int, _c_int, _c_p, _c_i64, _c_p],
    "pk_sub_log_prior": [_c_p, _c_i64, _c_i64, _c_int, _c_p, _c_p],
    "pk_adam_step": [_c_p, _c_p, _c_p, _c_p, _c_i64, _c_f, _c_f, _c_f, _c_f, _c_f, _c_i64, _c_f, _c_p],
    "pk_sgd_step": [_c_p, _c_p, _c_i64, _c_f, _c_f, _c_p],
    "pk_ln_cm_fwd": [_c_p, _c_int, _c_i64, _c_i64, _c_p, _c_p, _c_f, _c_p, _c_p, _c_p],
    "pk_ln_cm_bwd": [_c_p, _c_i64, _c_p, _c_i64, _c_p, _c_i64, _c_int, _c_i64, _c_p, _c_p, _c_f, _c_p, _c_p, _c_p, _c_p, _c_p],
    "pk_row_stats": [_c_p, _c_int, _c_i64, _c_i64, _c_p, _c_p],
}

_lib = None


def lib():
    """Load the shared library once.  Raises if it has not been built (python __graft_entry__.py)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback for the B200 path)")
        L = ctypes.CDLL(LIB_PATH)
        L.pk_last_error.restype = ctypes.c_char_p
        L.pk_last_error.argtypes = []
        L.pk_version.restype = _c_int
        L.pk_version.argtypes = []
        for name, argtypes in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = _c_i64 if name.endswith("_bytes") else _c_int
            fn.argtypes = argtypes
        _lib = L
    return _lib


# kernels launched by this library since import (bench.py reports it as gpu_launches); the
# numbers are the __global__ launches each C entry point performs (memsets are not counted)
launch_count = 0
KERNELS_PER_CALL = {"pk_dense_act_fwd": 1, "pk_dense_act_bwd": 1, "pk_amax_finalize": 1, "pk_gemm_tn": 1, "pk_transpose_f32": 1, "pk_convert_f16": 1, "pk_amax_scale": 2,
                    "pk_bn_finalize": 1, "pk_fill_scale_shift": 1, "pk_bn_bwd": 2, "pk_rnn_layer_fwd": 1,
                    "pk_rnn_layer_bwd": 1, "pk_rnn_step_fwd": 0, "pk_rnn_step_bwd": 0, "pk_rowln_fwd": 1, "pk_conv_ln0_bwd": 1, "pk_fusion_reduce_fwd": 1, "pk_fusion_reduce_bwd": 1,
                    "pk_sinc_filters_fwd": 1, "pk_sinc_filters_bwd": 1, "pk_conv_pack_weights": 1, "pk_conv_im2col0": 1,
                    "pk_conv_im2col_t": 1, "pk_conv_post_fwd": 1, "pk_conv_post_bwd": 1, "pk_logsoftmax_nll": 1, "pk_logsoftmax_bwd": 1, "pk_rmsprop_step": 1, "pk_adam_step": 1, "pk_chunk_prepare": 2, "pk_batch_assemble": 1, "pk_sub_log_prior": 1, "pk_cm_decode": 1,
                    "pk_sgd_step": 1, "pk_ln_cm_fwd": 1, "pk_ln_cm_bwd": 1, "pk_row_stats": 1}


def _check(rc, what, extra_kernels=0):
    global launch_count
    if rc != 0:
        raise RuntimeError(f"{what} failed (rc={rc}): {lib().pk_last_error().decode()}")
    launch_count += KERNELS_PER_CALL[what] + extra_kernels


def _ptr(t):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("pk_native: tensor is not on a CUDA device (no CPU fallback)")
    return t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def pad8(n: int) -> int:
    return (int(n) + 7) // 8 * 8


# ---------------------------------------------------------------------------------------------
# thin wrappers (shapes/strides spelled out by the caller; see include/pk_b200.h)
# ---------------------------------------------------------------------------------------------


def gemm_tn(A, B, C, M, N, K, *, lda, ldb, ldc, dtype=F16, a_k0=0, a_kext=0, b_k0=0, b_kext=0, bias=None,
            bias_mode=0, rowstats=None, alpha=1.0, alpha_dev=None, accumulate=False, split_k=1, amax_bits=None):
    _check(lib().pk_gemm_tn(dtype, M, N, K, _ptr(A), lda, a_k0, a_kext, _ptr(B), ldb, b_k0, b_kext, _ptr(C), ldc,
                            _ptr(bias), bias_mode if bias is not None else 0, _ptr(rowstats), float(alpha),
                            _ptr(alpha_dev), int(accumulate), int(split_k), _ptr(amax_bits), _stream()), "pk_gemm_tn")


def transpose_f32(inp, ldi, R, C, *, outT=None, ldo=0, outT16=None, ldo16=0, in16=None, ldi16=0, scale_dev=None,
                  amax_bits=None):
    _check(lib().pk_transpose_f32(_ptr(inp), ldi, R, C, _ptr(outT), ldo, _ptr(outT16), ldo16, _ptr(in16), ldi16,
                                  _ptr(scale_dev), _ptr(amax_bits), _stream()), "pk_transpose_f32")


def amax_finalize(amax_bits, target_log2, scale_out):
    _check(lib().pk_amax_finalize(_ptr(amax_bits), float(target_log2), _ptr(scale_out), _stream()), "pk_amax_finalize")


def convert_f16(inp, ldi, R, C, out, ldo, scale_dev=None):
    _check(lib().pk_convert_f16(_ptr(inp), ldi, R, C, _ptr(out), ldo, _ptr(scale_dev), _stream()), "pk_convert_f16")


def amax_scale(x, ld, R, C, target_log2, scratch, scale_out):
    _check(lib().pk_amax_scale(_ptr(x), ld, R, C, float(target_log2), _ptr(scratch), _ptr(scale_out), _stream()),
           "pk_amax_scale")


def bn_finalize(stats, C, n_unique, n_ref, gamma, beta, eps, momentum, training, running_mean, running_var,
                num_batches, scale, shift, mean_out, rstd_out):
    _check(lib().pk_bn_finalize(_ptr(stats), C, n_unique, n_ref, _ptr(gamma), _ptr(beta), float(eps),
                                float(momentum), int(training), _ptr(running_mean), _ptr(running_var),
                                _ptr(num_batches), _ptr(scale), _ptr(shift), _ptr(mean_out), _ptr(rstd_out),
                                _stream()), "pk_bn_finalize")


def fill_scale_shift(bias, C, scale, shift):
    _check(lib().pk_fill_scale_shift(_ptr(bias), C, _ptr(scale), _ptr(shift), _stream()), "pk_fill_scale_shift")


def bn_bwd(C, ndir, n, GT, GT16, ldt, PT, ldp, use_bn, training, mean, rstd, gamma, gscale, dgamma, dbeta, dPT16, ld16t,
           dP16, ld16r, sums_scratch):
    _check(lib().pk_bn_bwd(C, ndir, n, _ptr(GT), _ptr(GT16), ldt, _ptr(PT), ldp, int(use_bn), int(training), _ptr(mean),
                           _ptr(rstd), _ptr(gamma), _ptr(gscale), _ptr(dgamma), _ptr(dbeta), _ptr(dPT16), ld16t,
                           _ptr(dP16), ld16r, _ptr(sums_scratch), _stream()), "pk_bn_bwd")


def rnn_layer_fwd(cell, T, B, H, ndir, act, PT, ldp, scale, shift, U, mask, mask_scalar, Y32, ldy32, Y16, ldy16, HT,
                  HT16, HP16, ZT, HCT, ldt):
    _check(lib().pk_rnn_layer_fwd(cell, T, B, H, ndir, act, _ptr(PT), ldp, _ptr(scale), _ptr(shift), _ptr(U),
                                  _ptr(mask), float(mask_scalar), _ptr(Y32), ldy32, _ptr(Y16), ldy16, _ptr(HT),
                                  _ptr(HT16), _ptr(HP16), _ptr(ZT), _ptr(HCT), ldt, _stream()), "pk_rnn_layer_fwd")


def rnn_layer_bwd(cell, T, B, H, ndir, act, dYT, HT, ZT, HCT, ldt, U, mask, mask_scalar, gscale, GT, GT16):
    _check(lib().pk_rnn_layer_bwd(cell, T, B, H, ndir, act, _ptr(dYT), _ptr(HT), _ptr(ZT), _ptr(HCT), ldt, _ptr(U),
                                  _ptr(mask), float(mask_scalar), _ptr(gscale), _ptr(GT), _ptr(GT16), _stream()),
           "pk_rnn_layer_bwd")


def rnn_step_workspace_bytes(cell, T, B, H, ndir, backward):
    return int(lib().pk_rnn_step_workspace_bytes(cell, T, B, H, ndir, 1 if backward else 0))


def rnn_step_launches(cell, T, B, H, ndir, backward):
    """__global__ launches of one rnn_step_fwd / rnn_step_bwd call (depends on which kernel family takes the shape)."""
    return int(lib().pk_rnn_step_launches(cell, T, B, H, ndir, 1 if backward else 0))


def rnn_step_is_cluster(cell, H):
    """True when the step entry points run this (cell, H) on the cluster-persistent kernels (csrc/pk_cell_cluster.cu)."""
    return bool(lib().pk_rnn_step_is_cluster(cell, H))


def rnn_step_fwd(cell, T, B, H, ndir, act, PT, ldp, scale, shift, U, mask, mask_scalar, Y32, ldy32, Y16, ldy16, HT, HT16,
                 HP16, HX16, SV, ldt, workspace):
    sv = list(SV) + [None] * (5 - len(SV))
    _check(lib().pk_rnn_step_fwd(cell, T, B, H, ndir, act, _ptr(PT), ldp, _ptr(scale), _ptr(shift), _ptr(U), _ptr(mask),
                                 float(mask_scalar), _ptr(Y32), ldy32, _ptr(Y16), ldy16, _ptr(HT), _ptr(HT16),
                                 _ptr(HP16), _ptr(HX16), *[_ptr(s) for s in sv], ldt, _ptr(workspace), workspace.numel(), _stream()),
           "pk_rnn_step_fwd", rnn_step_launches(cell, T, B, H, ndir, False))


def rnn_step_bwd(cell, T, B, H, ndir, act, dYT, HT, SV, ldt, U, mask, mask_scalar, gscale, GT16, workspace):
    sv = list(SV) + [None] * (5 - len(SV))
    _check(lib().pk_rnn_step_bwd(cell, T, B, H, ndir, act, _ptr(dYT), _ptr(HT), *[_ptr(s) for s in sv], ldt, _ptr(U),
                                 _ptr(mask), float(mask_scalar), _ptr(gscale), _ptr(GT16), _ptr(workspace),
                                 workspace.numel(), _stream()), "pk_rnn_step_bwd",
           rnn_step_launches(cell, T, B, H, ndir, True))


def fusion_reduce_fwd(O, ldo, N, M, C, Hh, mode, slopes, red, P, ldp):
    _check(lib().pk_fusion_reduce_fwd(_ptr(O), ldo, N, M, C, Hh, mode, _ptr(slopes), float(red), _ptr(P), ldp, _stream()),
           "pk_fusion_reduce_fwd")


def fusion_reduce_bwd(dP, lddp, O, ldo, N, M, C, Hh, mode, slopes, red, dO, lddo, dbias, dslope):
    _check(lib().pk_fusion_reduce_bwd(_ptr(dP), lddp, _ptr(O), ldo, N, M, C, Hh, mode, _ptr(slopes), float(red), _ptr(dO),
                                      lddo, _ptr(dbias), _ptr(dslope), _stream()), "pk_fusion_reduce_bwd")


def rowln_fwd(x, ldx, N, L, gamma, beta, eps, y, stats):
    _check(lib().pk_rowln_fwd(_ptr(x), ldx, N, L, _ptr(gamma), _ptr(beta), float(eps), _ptr(y), _ptr(stats), _stream()),
           "pk_rowln_fwd")


def conv_ln0_bwd(G, ldg, N, L, Lout, k, x, ldx, stats, dgamma, dbeta):
    _check(lib().pk_conv_ln0_bwd(_ptr(G), ldg, N, L, Lout, k, _ptr(x), ldx, _ptr(stats), _ptr(dgamma), _ptr(dbeta),
                                 _stream()), "pk_conv_ln0_bwd")


def sinc_filters_fwd(low, band, C, k, sr, min_low, min_band, filt):
    _check(lib().pk_sinc_filters_fwd(_ptr(low), _ptr(band), C, k, float(sr), float(min_low), float(min_band), _ptr(filt),
                                     _stream()), "pk_sinc_filters_fwd")


def sinc_filters_bwd(low, band, C, k, sr, min_low, min_band, dfilt, dlow, dband):
    _check(lib().pk_sinc_filters_bwd(_ptr(low), _ptr(band), C, k, float(sr), float(min_low), float(min_band), _ptr(dfilt),
                                     _ptr(dlow), _ptr(dband), _stream()), "pk_sinc_filters_bwd")


def conv_pack_weights(w, Co, Ci, k, W16, Cip, ldw, Wflip16, Cop, ldf):
    _check(lib().pk_conv_pack_weights(_ptr(w), Co, Ci, k, _ptr(W16), Cip, ldw, _ptr(Wflip16), Cop, ldf, _stream()),
           "pk_conv_pack_weights")


def conv_im2col0(x, ldx, N, L, k, Lout, Xcol, Kp, XcolT, ldp):
    _check(lib().pk_conv_im2col0(_ptr(x), ldx, N, L, k, Lout, _ptr(Xcol), Kp, _ptr(XcolT), ldp, _stream()),
           "pk_conv_im2col0")


def conv_im2col_t(A16, rows, Cp, Ci, k, XT, ldp):
    _check(lib().pk_conv_im2col_t(_ptr(A16), rows, Cp, Ci, k, _ptr(XT), ldp, _stream()), "pk_conv_im2col_t")


def conv_post_fwd(O, ldo, N, L, Lout, p, Lp, C, act, gamma, beta, eps, keep16, P, arg, stats, A16n, Cpn, Y32):
    _check(lib().pk_conv_post_fwd(_ptr(O), ldo, N, L, Lout, p, Lp, C, act, _ptr(gamma), _ptr(beta), float(eps),
                                  _ptr(keep16), _ptr(P), _ptr(arg), _ptr(stats), _ptr(A16n), Cpn, _ptr(Y32), _stream()),
           "pk_conv_post_fwd")


def conv_post_bwd(dY, sn, sl, sc, N, L, Lout, p, Lp, C, act, gamma, beta, eps, keep16, P, arg, stats, dgamma, dbeta, dbias,
                  dO, amax_bits):
    _check(lib().pk_conv_post_bwd(_ptr(dY), sn, sl, sc, N, L, Lout, p, Lp, C, act, _ptr(gamma), _ptr(beta), float(eps),
                                  _ptr(keep16), _ptr(P), _ptr(arg), _ptr(stats), _ptr(dgamma), _ptr(dbeta), _ptr(dbias),
                                  _ptr(dO), _ptr(amax_bits), _stream()), "pk_conv_post_bwd")


def logsoftmax_nll(N, S, logits, ld, labels, acc):
    _check(lib().pk_logsoftmax_nll(N, S, _ptr(logits), ld, _ptr(labels), _ptr(acc), _stream()), "pk_logsoftmax_nll")


def logsoftmax_bwd(N, S, logp, ld, labels, dlogp, lddl, gcoef, out_scale, scale_dev, d16, ld16, dT16, ld16t, dbias,
                   rowsum_scratch):
    _check(lib().pk_logsoftmax_bwd(N, S, _ptr(logp), ld, _ptr(labels), _ptr(dlogp), lddl, float(gcoef),
                                   float(out_scale), _ptr(scale_dev), _ptr(d16), ld16, _ptr(dT16), ld16t,
                                   _ptr(dbias), _ptr(rowsum_scratch), _stream()), "pk_logsoftmax_bwd",
           1 if dlogp is not None else 0)


def dense_act_fwd(C, n, act, PT, ldp, scale, shift, keepT, ldk, YT16, ld16t, Y16, ld16r, Y32, ld32):
    _check(lib().pk_dense_act_fwd(C, n, act, _ptr(PT), ldp, _ptr(scale), _ptr(shift), _ptr(keepT), ldk, _ptr(YT16),
                                  ld16t, _ptr(Y16), ld16r, _ptr(Y32), ld32, _stream()), "pk_dense_act_fwd")


def dense_act_bwd(C, n, act, dYT, ldy, YT16, ld16t, keepT, ldk, gscale, GT16, ldg):
    _check(lib().pk_dense_act_bwd(C, n, act, _ptr(dYT), ldy, _ptr(YT16), ld16t, _ptr(keepT), ldk, _ptr(gscale),
                                  _ptr(GT16), ldg, _stream()), "pk_dense_act_bwd")


def rmsprop_step(p, g, v, lr, alpha, eps, gscale=1.0):
    _check(lib().pk_rmsprop_step(_ptr(p), _ptr(g), _ptr(v), p.numel(), float(lr), float(alpha), float(eps),
                                 float(gscale), _stream()), "pk_rmsprop_step")


def chunk_prepare(fea, lab, lab_min, left, right, out):
    n_in, F = fea.shape
    stats = torch.empty(2 * (left + right + 1) * F, device=fea.device, dtype=torch.float64)
    _check(lib().pk_chunk_prepare(_ptr(fea), fea.stride(0), _ptr(lab), int(lab_min), n_in, F, left, right, _ptr(stats),
                                  _ptr(out), out.stride(0), _stream()), "pk_chunk_prepare")


def batch_assemble(data_set, desc, batch_size, max_len, inp):
    _check(lib().pk_batch_assemble(_ptr(data_set), data_set.stride(0), data_set.shape[1], _ptr(desc), batch_size, max_len,
                                   _ptr(inp), _stream()), "pk_batch_assemble")


def cm_decode(col_headers, data, min_value, rng, rows, cols, out):
    _check(lib().pk_cm_decode(_ptr(col_headers), _ptr(data), float(min_value), float(rng), rows, cols, _ptr(out), out.stride(0),
                              _stream()), "pk_cm_decode")


def sub_log_prior(logp, log_prior):
    _check(lib().pk_sub_log_prior(_ptr(logp), logp.stride(0), logp.shape[0], logp.shape[1], _ptr(log_prior), _stream()),
           "pk_sub_log_prior")


def adam_step(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step, gscale=1.0):
    _check(lib().pk_adam_step(_ptr(p), _ptr(g), _ptr(m), _ptr(v), p.numel(), float(lr), float(beta1), float(beta2),
                              float(eps), float(weight_decay), int(step), float(gscale), _stream()), "pk_adam_step")


def sgd_step(p, g, lr, gscale=1.0):
    _check(lib().pk_sgd_step(_ptr(p), _ptr(g), p.numel(), float(lr), float(gscale), _stream()), "pk_sgd_step")


def ln_cm_fwd(PT, C, n, ld, gamma, beta, eps, XH, stats):
    _check(lib().pk_ln_cm_fwd(_ptr(PT), C, n, ld, _ptr(gamma), _ptr(beta), eps, _ptr(XH), _ptr(stats), _stream()), "pk_ln_cm_fwd")


def ln_cm_bwd(dT16, ld16t, dR16, ld16r, XH, ld, C, n, gamma, stats, eps, scale, dgamma, dbeta, dbias=None):
    _check(lib().pk_ln_cm_bwd(_ptr(dT16), ld16t, _ptr(dR16), ld16r, _ptr(XH), ld, C, n, _ptr(gamma), _ptr(stats), eps,
                              _ptr(scale), _ptr(dgamma), _ptr(dbeta), _ptr(dbias), _stream()), "pk_ln_cm_bwd", 2)


def row_stats(PT, C, n, ld, stats):
    _check(lib().pk_row_stats(_ptr(PT), C, n, ld, _ptr(stats), _stream()), "pk_row_stats")
