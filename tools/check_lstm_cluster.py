"""A/B check + timing of the cluster-persistent LSTM kernels (csrc/pk_cell_cluster.cu) against the step-wise path
(csrc/pk_cell_step.cu) through the same C-ABI entry points (pk_rnn_step_fwd / pk_rnn_step_bwd).

    python tools/check_lstm_cluster.py [lstm] [gru] [minimalgru]   # small shapes + T=500, B=32, H=550, bidirectional

Forward: both paths run the same fp16 operands through mma.sync in the same k order -> outputs must agree to fp32
rounding.  Backward: the cluster kernel sums K-split partial products in a different order -> fp16 outputs agree to
an fp16 ulp.  Exit code 1 on any mismatch.  Bring-up helper; tests/test_gpu_parity.py holds the oracle tests."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-kaldi_b200"))
import pk_native as pk  # noqa: E402

pad8 = pk.pad8


CELLS = {"lstm": (pk.CELL_LSTM, 4, "PK_LSTM_CLUSTER"), "gru": (pk.CELL_GRU, 3, "PK_GRU_CLUSTER"),
         "minimalgru": (pk.CELL_MGRU, 2, "PK_GRU_CLUSTER")}


def make(T, B, H, ndir, seed=0, cell="lstm"):
    g = torch.Generator(device="cuda").manual_seed(seed)
    dev = "cuda"
    TB, ldt, F = T * B, pad8(T * B), ndir * H
    cid, ng, env = CELLS[cell]
    d = dict(T=T, B=B, H=H, ndir=ndir, ldt=ldt, F=F, cell=cid, ng=ng, env=env)
    d["PT"] = torch.randn(ng * H, ldt, device=dev, generator=g)
    d["scale"] = torch.rand(ng * H, device=dev, generator=g) + 0.5
    d["shift"] = torch.randn(ng * H, device=dev, generator=g) * 0.1
    d["U"] = torch.randn(ng * H, H, device=dev, generator=g) / H ** 0.5
    d["mask"] = (torch.rand(ndir * B, H, device=dev, generator=g) < 0.8).float()
    d["dYT"] = torch.randn(F, ldt, device=dev, generator=g) * 1e-3
    return d


def run_fwd(d, act, cluster):
    os.environ[d["env"]] = "1" if cluster else "0"
    T, B, H, ndir, ldt, F, cell = d["T"], d["B"], d["H"], d["ndir"], d["ldt"], d["F"], d["cell"]
    dev = "cuda"
    nsv = {pk.CELL_LSTM: 5, pk.CELL_GRU: 3}.get(cell, 2)
    o = dict(Y32=torch.zeros(T, B, F, device=dev), Y16=torch.zeros(T * B, pad8(F), device=dev, dtype=torch.float16),
             HT=torch.zeros(F, ldt, device=dev), HT16=torch.zeros(F, ldt, device=dev, dtype=torch.float16),
             HP16=torch.zeros(F, ldt, device=dev, dtype=torch.float16),
             HX16=torch.zeros(F, ldt, device=dev, dtype=torch.float16) if cell != pk.CELL_LSTM else None,
             SV=[torch.zeros(F, ldt, device=dev) for _ in range(nsv)])
    ws = torch.empty(pk.rnn_step_workspace_bytes(cell, T, B, H, ndir, False), device=dev, dtype=torch.uint8)
    n = pk.rnn_step_launches(cell, T, B, H, ndir, False)

    def call():
        pk.rnn_step_fwd(cell, T, B, H, ndir, act, d["PT"], ldt, d["scale"], d["shift"], d["U"], d["mask"], 1.0,
                        o["Y32"], F, o["Y16"], pad8(F), o["HT"], o["HT16"], o["HP16"], o["HX16"], o["SV"], ldt, ws)
    call()
    torch.cuda.synchronize()
    return o, call, n


def run_bwd(d, act, saved, cluster):
    os.environ[d["env"]] = "1" if cluster else "0"
    T, B, H, ndir, ldt, cell = d["T"], d["B"], d["H"], d["ndir"], d["ldt"], d["cell"]
    dev = "cuda"
    GT16 = torch.zeros(ndir, d["ng"] * H, ldt, device=dev, dtype=torch.float16)
    sc = torch.tensor([2.0 ** 14, 2.0 ** -14], device=dev)
    ws = torch.empty(pk.rnn_step_workspace_bytes(cell, T, B, H, ndir, True), device=dev, dtype=torch.uint8)

    def call():
        pk.rnn_step_bwd(cell, T, B, H, ndir, act, d["dYT"], saved["HT"], saved["SV"], ldt, d["U"], d["mask"], 1.0,
                        sc, GT16, ws)
    call()
    torch.cuda.synchronize()
    return GT16, call


def time_ms(call, iters):
    for _ in range(2):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        call()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def phase_clocks(call, T, names):
    """cycles per step and phase of CTA 0 / thread 0 (bring-up instrumentation of pk_cell_cluster.cu)"""
    import ctypes
    L = pk.lib()
    buf = (ctypes.c_longlong * 16)()
    L.pk_debug_cluster_clocks(1, None)
    call()
    torch.cuda.synchronize()
    L.pk_debug_cluster_clocks(0, buf)
    return {n: round(buf[i] / T) for i, n in names.items() if buf[i]}


FWD_PHASES = {0: "prefetch", 1: "wait", 2: "mma", 3: "gates", 4: "push", 5: "stores"}
BWD_PHASES = {8: "prefetch", 9: "pointwise", 10: "mma+push", 11: "wait", 12: "sum"}


def main():
    torch.cuda.set_device(0)
    fails = 0
    cells = sys.argv[1:] or ["lstm"]
    shapes = [(20, 8, 96, 2, "tanh"), (12, 5, 200, 2, "relu"), (7, 3, 24, 1, "tanh"), (9, 16, 330, 1, "sigmoid"),
              (500, 32, 550, 2, "tanh")]
    for cell, (T, B, H, ndir, actn) in [(c, sh) for c in cells for sh in shapes]:
        if cell != "lstm" and (T, actn) == (500, "tanh"):
            actn = "relu"     # the shipped GRU recipes
        act = pk.ACT_IDS[actn]
        d = make(T, B, H, ndir, cell=cell)
        env = d["env"]
        old, call_old, n_old = run_fwd(d, act, False)
        new, call_new, n_new = run_fwd(d, act, True)
        errs = {}
        for k in ("Y32", "Y16", "HT", "HT16", "HP16") + (("HX16",) if cell != "lstm" else ()):
            errs[k] = (old[k].float() - new[k].float()).abs().max().item()
        for i in range(len(old["SV"])):
            errs[f"SV{i}"] = (old["SV"][i] - new["SV"][i]).abs().max().item()
        worst = max(errs.values())
        ok = worst <= 2e-3 and torch.isfinite(new["Y32"]).all().item()
        big = T * B * H >= 1e6
        line = f"{cell} fwd T={T} B={B} H={H} ndir={ndir} {actn}: max|old-new| = {worst:.3e} ({'OK' if ok else 'FAIL'}; launches {n_old} -> {n_new})"
        if big:
            os.environ[env] = "0"
            t_old = time_ms(call_old, 5)
            os.environ[env] = "1"
            t_new = time_ms(call_new, 5)
            line += f"  old {t_old:.3f} ms  new {t_new:.3f} ms ({1e3 * t_new / T:.2f} us/step)  cycles/step {phase_clocks(call_new, T, FWD_PHASES)}"
        print(line, flush=True)
        if not ok:
            fails += 1
            print("   per-tensor:", {k: f"{v:.2e}" for k, v in errs.items()})
        g_old, bcall_old = run_bwd(d, act, old, False)
        g_new, bcall_new = run_bwd(d, act, old, True)
        ref = g_old.float().abs().max().item()
        err = (g_old.float() - g_new.float()).abs().max().item()
        rel = ((g_old.float() - g_new.float()).norm() / g_old.float().norm().clamp_min(1e-30)).item()
        okb = rel <= 5e-3 and err <= 2e-2 * max(ref, 1e-6) and torch.isfinite(g_new.float()).all().item()
        line = (f"{cell} bwd T={T} B={B} H={H} ndir={ndir} {actn}: rel-L2 {rel:.2e}, max|old-new| = {err:.3e} of max {ref:.3e} "
                f"({'OK' if okb else 'FAIL'})")
        if big:
            os.environ[env] = "0"
            t_old = time_ms(bcall_old, 5)
            os.environ[env] = "1"
            t_new = time_ms(bcall_new, 5)
            line += f"  old {t_old:.3f} ms  new {t_new:.3f} ms ({1e3 * t_new / T:.2f} us/step)  cycles/step {phase_clocks(bcall_new, T, BWD_PHASES)}"
        print(line, flush=True)
        if not okb:
            fails += 1
    os.environ.pop("PK_LSTM_CLUSTER", None)
    os.environ.pop("PK_GRU_CLUSTER", None)
    print("RESULT", "FAIL" if fails else "PASS")
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
