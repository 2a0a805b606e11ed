#!/usr/bin/env python
"""Second comparison bar: the UNMODIFIED reference modules on the same B200 with stock
PyTorch ops (use_cuda=True, TF32 off), and cuDNN's nn.LSTM through the reference's LSTM_cudnn class as the speed
bar of config 3.  None of this repository's kernels run here.  GPU box only; needs baseline/_ref (shipped by
__graft_entry__.build()).

    python tools/ref_on_gpu.py > gpurun_out/ref_on_gpu.json
"""
import importlib.util
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (option builders only)


def load_ref():
    p = os.path.join(ROOT, "baseline", "_ref", "neural_networks.py")
    spec = importlib.util.spec_from_file_location("ref_neural_networks", p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def time_steps(step, warm, n):
    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def make_step(mods, x, lab, lr):
    opts = [torch.optim.RMSprop(m.parameters(), lr=lr, alpha=0.95, eps=1e-8) for m in mods]
    lossf = torch.nn.NLLLoss()

    def step():
        h = mods[0](x)
        logp = mods[1](h.view(h.shape[0] * h.shape[1], -1))
        loss = lossf(logp, lab)
        for o in opts:
            o.zero_grad()
        loss.backward()
        for o in opts:
            o.step()
        return loss

    return step


def main():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    ref = load_ref()
    dev = torch.device("cuda", 0)
    out = {"device": torch.cuda.get_device_name(0), "torch": torch.__version__, "tf32": False, "rows": []}
    g = torch.Generator().manual_seed(1234)
    T, B, F, S = 500, 32, 40, 1936
    x = torch.randn(T, B, F, generator=g).to(dev)
    lab = torch.randint(0, S, (T * B,), generator=g).to(dev)

    def row(name, mods, warm, n, lr=0.0004):
        torch.manual_seed(1234)
        for m in mods:
            m.to(dev).train()
        ms = time_steps(make_step(mods, x, lab, lr), warm, n)
        out["rows"].append({"what": name, "ms_per_step": ms, "frames_per_s": T * B / (ms * 1e-3), "steps_timed": n})
        print(name, f"{ms:.1f} ms/step", file=sys.stderr, flush=True)

    # config 2: the reference's own liGRU (python time loop, stock PyTorch kernels) + MLP head
    c = bench.CONFIGS["ligru5x550"]
    net = ref.liGRU(bench.rec_opts("ligru", c["lay"], c["act"], "True"), F)
    head = ref.MLP(bench.head_opts(S, "True"), net.out_dim)
    row("reference liGRU 5x550 bidir + 1936 head on B200, stock PyTorch fp32 (configs[1])", [net, head], 1, 3)
    # config 3: the reference's own LSTM class, and cuDNN through LSTM_cudnn
    c = bench.CONFIGS["lstm4x550"]
    net = ref.LSTM(bench.rec_opts("lstm", c["lay"], c["act"], "True"), F)
    head = ref.MLP(bench.head_opts(S, "True"), net.out_dim)
    row("reference LSTM 4x550 bidir (BN, shared-direction weights) + head on B200, stock PyTorch fp32 (configs[2])",
        [net, head], 1, 3, 0.0016)
    copts = dict(hidden_size="550", num_layers="4", bias="True", batch_first="True", dropout="0.2", bidirectional="True",
                 use_cuda="True", to_do="train")
    net = ref.LSTM_cudnn(copts, F)
    head = ref.MLP(bench.head_opts(S, "True"), net.out_dim)
    row("reference LSTM_cudnn (nn.LSTM 4x550 bidir, cuDNN fp32, no BN, per-direction weights) + head on B200 — "
        "config 3's speed bar, not a parity oracle", [net, head], 3, 10, 0.0016)
    # the same with PyTorch's DEFAULT flags (torch.backends.cudnn.allow_tf32 = True: what the unmodified reference runs)
    torch.backends.cudnn.allow_tf32 = True
    net = ref.LSTM_cudnn(copts, F)
    head = ref.MLP(bench.head_opts(S, "True"), net.out_dim)
    row("reference LSTM_cudnn as above with PyTorch's default torch.backends.cudnn.allow_tf32=True (TF32 recurrent GEMMs)",
        [net, head], 3, 10, 0.0016)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
