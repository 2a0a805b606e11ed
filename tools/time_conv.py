"""Time one training step (forward + backward) of the SincNet front-end + softmax head on cuda:0 at the shapes of
cfg/TIMIT_baselines/TIMIT_SincNet_raw.cfg (128 filters x 129 taps, 60 x 5, 60 x 5, 60 x 3; pools 3,3,3,2; LayerNorm;
3200-sample frames).   python tools/time_conv.py [N_frames=128]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-kaldi_b200"))
import neural_networks as pknn  # noqa: E402


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    L0, S = 3200, 1936
    o = {"sinc_N_filt": "128,60,60,60", "sinc_len_filt": "129,5,5,3", "sinc_max_pool_len": "3,3,3,2",
         "sinc_use_laynorm_inp": "True", "sinc_use_batchnorm_inp": "False", "sinc_use_laynorm": "True,True,True,True",
         "sinc_use_batchnorm": "False,False,False,False", "sinc_act": "relu,relu,relu,relu",
         "sinc_drop": "0.15,0.15,0.15,0.15", "sinc_sample_rate": "16000", "sinc_min_low_hz": "50",
         "sinc_min_band_hz": "50", "use_cuda": "True", "to_do": "train"}
    net = pknn.SincNet(o, L0).cuda().train()
    head = pknn.MLP({"dnn_lay": str(S), "dnn_drop": "0.0", "dnn_use_laynorm_inp": "False",
                     "dnn_use_batchnorm_inp": "False", "dnn_use_batchnorm": "False", "dnn_use_laynorm": "False",
                     "dnn_act": "softmax", "use_cuda": "True", "to_do": "train"}, net.out_dim).cuda().train()
    x = torch.randn(N, L0, device="cuda")
    lab = torch.randint(0, S, (N,), device="cuda")

    def step():
        for p in list(net.parameters()) + list(head.parameters()):
            p.grad = None
        loss = torch.nn.functional.nll_loss(head(net(x)), lab)
        loss.backward()
        return loss

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 5
    t0 = time.time()
    e0.record()
    for _ in range(K):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    flops = 3 * 2 * N * (3072 * 128 * 129 + 1020 * 60 * 128 * 5 + 336 * 60 * 60 * 5 + 110 * 60 * 60 * 3)
    print(f"SincNet N={N}: {ms:.2f} ms/step (wall {1e3 * (time.time() - t0) / K:.2f}), {N / ms * 1e3:.0f} frames/s, "
          f"{flops / ms / 1e9:.1f} TFLOP/s (conv FLOPs, train = 3x fwd), loss {loss.item():.4f}")


if __name__ == "__main__":
    main()
