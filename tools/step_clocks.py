"""Bring-up: per-phase cycle breakdown of the persistent step-wise kernels (CTA 0 / thread 0).
    python tools/step_clocks.py lstm 550"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-kaldi_b200"))
import neural_networks as pknn  # noqa: E402
import pk_native as pk  # noqa: E402

cell = sys.argv[1] if len(sys.argv) > 1 else "lstm"
H = int(sys.argv[2]) if len(sys.argv) > 2 else 550
T, B, F = 500, 32, 40
o = {f"{cell}_lay": str(H), f"{cell}_drop": "0.2", f"{cell}_use_laynorm_inp": "False", f"{cell}_use_batchnorm_inp": "False",
     f"{cell}_use_laynorm": "False", f"{cell}_use_batchnorm": "True", f"{cell}_bidir": "True", f"{cell}_act": "tanh",
     f"{cell}_orthinit": "True", "use_cuda": "True", "to_do": "train"}
cls = {"ligru": "liGRU", "lstm": "LSTM", "gru": "GRU", "minimalgru": "minimalGRU"}[cell]
net = getattr(pknn, cls)(o, F).cuda().train()
net.fast_dropout = True
x = torch.randn(T, B, F, device="cuda")
L = pk.lib()
out = (ctypes.c_longlong * 16)()
for it in range(3):
    L.pk_debug_step_clocks(1, out)
    y = net(x)
    y.sum().backward()
    torch.cuda.synchronize()
    L.pk_debug_step_clocks(0, out)
v = list(out)
nb = T * (2 if cell in ("gru", "minimalgru") else 1)
print(f"{cell} H={H}: cycles per body call (CTA 0): fwd wait {v[0]/nb:.0f} | copy+mma {v[1]/nb:.0f} | epilogue {v[2]/nb:.0f} | arrive {v[3]/nb:.0f}"
      f"  ||  bwd wait {v[8]/nb:.0f} | gemm {v[9]/nb:.0f} | epilogue {v[10]/nb:.0f} | arrive {v[11]/nb:.0f}")
