"""Time one full training step (forward + backward) of a recurrent stack on cuda:0.

    python tools/time_cell.py lstm 4 550      # cell, layers, hidden   (T=500, B=32, F=40, S=1936)

Bring-up / documentation helper (DESIGN.md quotes its numbers); bench.py stays the contract benchmark."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-kaldi_b200"))
import neural_networks as pknn  # noqa: E402


def main():
    cell = sys.argv[1] if len(sys.argv) > 1 else "lstm"
    nlay = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    H = int(sys.argv[3]) if len(sys.argv) > 3 else 550
    T, B, F, S = int(os.environ.get("PK_T", "500")), 32, 40, 1936
    n = nlay
    opts = {f"{cell}_lay": ",".join([str(H)] * n), f"{cell}_drop": ",".join(["0.2"] * n),
            f"{cell}_use_laynorm_inp": "False", f"{cell}_use_batchnorm_inp": "False",
            f"{cell}_use_laynorm": ",".join(["False"] * n), f"{cell}_use_batchnorm": ",".join(["True"] * n),
            f"{cell}_bidir": "True", f"{cell}_act": ",".join(["tanh" if cell == "lstm" else "relu"] * n),
            f"{cell}_orthinit": "True", "use_cuda": "True", "to_do": "train"}
    cls = {"ligru": "liGRU", "rnn": "RNN", "lstm": "LSTM", "gru": "GRU", "minimalgru": "minimalGRU"}[cell]
    net = getattr(pknn, cls)(opts, F).cuda().train()
    net.fast_dropout = True
    head = pknn.MLP({"dnn_lay": str(S), "dnn_drop": "0.0", "dnn_use_laynorm_inp": "False",
                     "dnn_use_batchnorm_inp": "False", "dnn_use_batchnorm": "False", "dnn_use_laynorm": "False",
                     "dnn_act": "softmax", "use_cuda": "True", "to_do": "train"}, net.out_dim).cuda().train()
    x = torch.randn(T, B, F, device="cuda")
    lab = torch.randint(0, S, (T * B,), device="cuda")

    def step():
        for p in list(net.parameters()) + list(head.parameters()):
            p.grad = None
        loss = torch.nn.functional.nll_loss(head(net(x).view(T * B, -1)), lab)
        loss.backward()
        return loss

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    K = 5
    for _ in range(K):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print(f"{cell} {nlay}x{H} bidir: {ms:.2f} ms/step (wall {1e3 * (time.time() - t0) / K:.2f}), "
          f"{T * B / ms * 1e3:.0f} frames/s, loss {loss.item():.4f}")


if __name__ == "__main__":
    main()
