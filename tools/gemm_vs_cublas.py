"""pk_gemm_tn (hand-written tcgen05 / TMA / TMEM GEMM, csrc/pk_gemm.cu) next to cuBLAS (torch.matmul, fp16 operands,
fp32 accumulate inside the library) on the GEMM shapes of the config-2 training step, same B200, CUDA events.

    python tools/gemm_vs_cublas.py

cuBLAS returns fp16 here (its fastest path); pk_gemm_tn writes fp32 (what the recurrent kernels consume), i.e. it
moves 2x the output bytes — the comparison is conservative for the hand-written kernel."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-kaldi_b200"))
import pk_native as pk  # noqa: E402


def time_ms(fn, iters=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    torch.cuda.set_device(0)
    T, B, H, S = 500, 32, 550, 1936
    TB = T * B
    shapes = [  # name, M, N, K, split_k   (C[M,N] = A[M,K] . B[N,K]^T)
        ("projection  [2H x TB x 2H]", 2 * H, TB, 2 * H, 1),
        ("logits      [TB x S x 2H]", TB, S, 2 * H, 1),
        ("dX          [2H x TB x 2H]", 2 * H, TB, 2 * H, 1),
        ("dW          [2H x 2H x TB] split-K 8", 2 * H, 2 * H, TB, 8),
        ("dU          [2H x H x TB] split-K 16", 2 * H, H, TB, 16),
        ("head dW     [S x 2H x TB] split-K 8", S, 2 * H, TB, 8),
    ]
    print(f"{'shape':44s} {'pk_gemm_tn':>22s} {'cuBLAS fp16 (torch.matmul)':>30s} {'cuBLAS, K padded to 8':>26s}")
    for name, M, N, K, sk in shapes:
        ldk = pk.pad8(K)
        A = torch.randn(M, ldk, device="cuda").half()
        Bm = torch.randn(N, ldk, device="cuda").half()
        C = torch.empty(M, N, device="cuda")
        t_pk = time_ms(lambda: pk.gemm_tn(A, Bm, C, M, N, K, lda=ldk, ldb=ldk, ldc=N, split_k=sk))
        At, Bt = A[:, :K].contiguous(), Bm[:, :K].contiguous()
        t_cb = time_ms(lambda: torch.matmul(At, Bt.t()))
        # K = 1100 gives cuBLAS rows that are not 16-byte aligned; the same product on zero-padded operands (K = pad8(K),
        # what pk_gemm_tn's TMA descriptors see) is the fair bar
        Ap, Bp = A.clone(), Bm.clone()
        Ap[:, K:] = 0
        Bp[:, K:] = 0
        t_cp = time_ms(lambda: torch.matmul(Ap, Bp.t()))
        fl = 2.0 * M * N * K
        ref = torch.matmul(At.float(), Bt.float().t())
        err = ((C - ref).abs().max() / ref.abs().max()).item()
        print(f"{name:44s} {t_pk * 1e3:8.1f} us {fl / t_pk / 1e9:7.0f} TF/s   {t_cb * 1e3:8.1f} us {fl / t_cb / 1e9:7.0f} TF/s   "
              f"padded K: {t_cp * 1e3:8.1f} us {fl / t_cp / 1e9:7.0f} TF/s   (max rel err {err:.1e})")


if __name__ == "__main__":
    main()
