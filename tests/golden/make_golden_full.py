"""Full-size golden vectors (BASELINE.json configs 2, 3, 4) from the UNMODIFIED reference.

Build container only (needs /root/reference):

    python tests/golden/make_golden_full.py [case ...]

Runs the recipe of tests/full_cases.py against /root/reference/neural_networks.py on the CPU in fp32 at the
size the headline metric is quoted on (500x32x40 chunks, 5x550 bidirectional liGRU + 1936-way head, ...), i.e.
the reference path neural_networks.py:997-1155 (+ :300-483 for the LSTM) -> MLP head :60-150 -> NLLLoss /
cost_err (utils.py:2344-2381) -> backward -> RMSprop (utils.py:2121-2131), and stores REDUCED outputs
(a few MB per case): loss, err, every 97th log-posterior row, arg-max + top-2 margin of every row, hidden
state samples, sampled gradients with per-tensor norms, sampled parameters after the optimizer step.
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
REF = os.environ.get("PK_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
import neural_networks as ref_nn  # noqa: E402  (the reference's module zoo)
import full_cases as fc  # noqa: E402
from make_golden import MaskRecorder  # noqa: E402


def run(case):
    c = fc.CASES[case]
    T, B, S = c["T"], c["B"], c["S"]
    t0 = time.time()
    net, head = fc.build(ref_nn, case)
    out = {}
    for k, p in fc.state_pairs(net, head):
        out["csum." + k] = fc.checksum(p)
    x, lab = fc.inputs(case)
    out["csum.x"] = fc.checksum(x)
    out["csum.lab"] = fc.checksum(lab.double())
    net.train(); head.train()
    torch.manual_seed(fc.forward_seed(case))
    with MaskRecorder() as rec:
        h = net(x)
    flat = h.view(T * B, -1)
    logp = head(flat)
    loss = torch.nn.NLLLoss()(logp, lab)
    top2 = torch.topk(logp, 2, dim=1)
    pred = top2.indices[:, 0]
    err = torch.mean((pred != lab).float())
    print(f"{case}: forward {time.time()-t0:.1f}s loss={loss.item():.6f} err={err.item():.4f}", flush=True)
    out.update(loss=np.float64(loss.item()), err=np.float64(err.item()),
               logp_rows=logp[::fc.ROW_STRIDE].detach().numpy().copy(),
               pred=pred.numpy().astype(np.int16), margin=(top2.values[:, 0] - top2.values[:, 1]).detach().numpy(),
               logp_lab=logp.detach().gather(1, lab.view(-1, 1)).view(-1).numpy().copy(),
               logp_rowmax=top2.values[:, 0].detach().numpy().copy())
    # hidden state of the last layer at a few (t, b) and a strided sample over everything
    tb = [(0, 0), (1, 5), (17, 31), (123, 7), (249, 16), (250, 16), (377, 2), (498, 30), (499, 0), (499, 31)]
    out["h_tb"] = np.array(tb)
    out["h_rows"] = np.stack([h[t, b].detach().numpy() for t, b in tb])
    hidx = fc.sample_idx(h.numel(), 65536, seed=1)
    out["h_val"] = h.detach().reshape(-1).numpy()[hidx]
    out["h_sumsq"] = np.float64((h.detach().double() ** 2).sum().item())
    for i, m in enumerate(rec.masks):
        out[f"maskbits{i}"] = np.packbits(m.astype(np.uint8))
        out[f"maskshape{i}"] = np.array(m.shape)
    if c["backward"]:
        mods = [net, head]
        opts = [torch.optim.RMSprop(m.parameters(), lr=0.0004, alpha=0.95, eps=1e-8) for m in mods]
        for o in opts:
            o.zero_grad()
        t1 = time.time()
        loss.backward()
        print(f"{case}: backward {time.time()-t1:.1f}s", flush=True)
        for k, p in fc.state_pairs(net, head):
            if p.grad is None:
                continue
            g = p.grad.detach().reshape(-1).numpy()
            idx = fc.sample_idx(g.size, keep=fc.GRAD_KEEP)
            out["grad." + k + ".val"] = g[idx].copy()
            out["grad." + k + ".sumsq"] = np.float64((g.astype(np.float64) ** 2).sum())
        for o in opts:
            o.step()
        for k, p in fc.state_pairs(net, head):
            if p.grad is None:
                continue
            v = p.detach().reshape(-1).numpy()
            out["step1." + k + ".val"] = v[fc.sample_idx(v.size, keep=fc.GRAD_KEEP)].copy()
        for pfx, m in (("net.", net), ("head.", head)):
            for k, v in m.state_dict().items():
                if "running" in k:
                    out["bnstat." + pfx + k] = v.numpy().copy()
    out["meta"] = np.array(repr(dict(c, case=case, torch=torch.__version__)))
    path = os.path.join(HERE, case + ".npz")
    np.savez_compressed(path, **out)
    print(f"{case}: total {time.time()-t0:.1f}s -> {path} ({os.path.getsize(path)/1e6:.2f} MB)", flush=True)


if __name__ == "__main__":
    torch.set_num_threads(int(os.environ.get("PK_THREADS", "8")))
    for name in (sys.argv[1:] or list(fc.CASES)):
        run(name)
