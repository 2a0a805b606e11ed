#!/usr/bin/env python
"""Benchmark of the hot path BASELINE.json names: training frames/sec of the neural_networks.py module zoo driven
the way core.run_nn drives it, headline = 5 x 550 bidirectional liGRU + 1936-way softmax head on synthetic
500 x 32 x 40 chunks (cfg/TIMIT_baselines/TIMIT_liGRU_fmllr.cfg shape, BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W                  # this repository's CUDA path, config 2
    python bench.py --config lstm4x550 ...                         # configs[2]; also mlp4x1024 (configs[0]),
                                                                   # ligru5x1024_3440 (configs[3]), sincnet_ligru (configs[4])
    python bench.py --impl reference --gpus N --steps K ...        # the reference's own CPU path on the host cores

One "step" = one minibatch exactly as core.run_nn runs it (reference core.py:616-642): forward_model (module
stack -> softmax head -> NLLLoss + frame error), zero_grad, backward, optimizer step (+ ONE NCCL gradient allreduce
when N > 1; every rank trains on its own chunk: weak scaling).

Timing protocol: W >= 3 warm-up steps; then `--repeats` windows of exactly K steps, each window bracketed by a
barrier + torch.cuda.synchronize() and timed with CUDA events, MAX over ranks per window, MEDIAN window reported
(`windows_ms` lists all).  Nothing else runs inside a window: clocks are sampled by ONE nvidia-smi process started
by rank 0 before the first window; the per-launch timing of the dominant kernel is a separate pass afterwards.
Between steps a 256 MiB buffer is written (L2 flush) and an 8-chunk input ring is rotated.

Prints ONE JSON line (rank 0).  `value` = frames/s with the chunk already resident in HBM; `e2e` = the same through
the drop-in module API from pinned HOST buffers (H2D of the chunk and D2H of the loss inside the timed region).
`roofline` describes the dominant kernel; `parity` is the step-0 self check against the committed full-size
reference fixture (tests/golden/full_*.npz); `cpu_baseline` = the reference's own modules (baseline/_ref, shipped
git-ignored by __graft_entry__.build()) on this box's host cores on a bounded sample.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "pytorch-kaldi_b200"))
REF_DIR = os.path.join(ROOT, "baseline", "_ref")

METRIC = "frames/sec liGRU-5x550 on 500x32x40 chunks @1/2/4/8 B200; loss match <=1e-3"

# name -> workload.  kind: "rec" = [T,B,F] chunk -> recurrent stack -> head (chunk_step); "mlp" = [N,F] -> MLP;
# "sinc" = raw [T*B, 3200] -> SincNet -> view [T,B,3300] -> liGRU -> head (utils.py:2336-2337)
CONFIGS = {
    "ligru5x550": dict(kind="rec", cell="ligru", T=500, B=32, F=40, lay=[550] * 5, act="relu", S=1936, opt="rmsprop", lr=0.0004,
                       fixture="full_ligru5x550", metric=METRIC,
                       workload="TIMIT liGRU 5x550 bidir, fMLLR-40, synthetic 500x32x40 chunks, 1936 senones (configs[1])"),
    "lstm4x550": dict(kind="rec", cell="lstm", T=500, B=32, F=40, lay=[550] * 4, act="tanh", S=1936, opt="rmsprop", lr=0.0016,
                      fixture="full_lstm4x550", metric="frames/sec LSTM-4x550 bidir on 500x32x40 chunks",
                      workload="TIMIT LSTM 4x550 bidir, fbank-40, synthetic 500x32x40 chunks, 1936 senones (configs[2])"),
    "ligru5x1024_3440": dict(kind="rec", cell="ligru", T=500, B=32, F=40, lay=[1024] * 5, act="relu", S=3440, opt="rmsprop",
                             lr=0.0004, fixture="full_ligru5x1024", metric="frames/sec liGRU-5x1024 / 3440 on 500x32x40 chunks",
                             workload="Librispeech liGRU 5x1024 bidir, 3440 senones, synthetic 500x32x40 chunks (configs[3])"),
    "mlp4x1024": dict(kind="mlp", N=128, F=429, lay=[1024] * 4, S=1936, opt="sgd", lr=0.08, fixture=None,
                      metric="frames/sec MLP-4x1024 on 128x429 minibatches",
                      workload="TIMIT MLP 4x1024 (BN, ReLU, dropout .15) on 39-dim MFCC x 11 frames, batches of 128 (configs[0])"),
    "sincnet_ligru": dict(kind="sinc", cell="ligru", T=500, B=8, F=3200, lay=[550] * 5, act="relu", S=1936, opt="rmsprop",
                          lr=0.0004, fixture=None, metric="frames/sec SincNet + liGRU-5x550 on 500x8x3200 raw chunks",
                          workload="SincNet (128/60/60/60; 129/5/5/3; pool 3/3/3/2; LN) -> liGRU 5x550 bidir -> 1936 senones, "
                                   "synthetic 500x8x3200 raw-waveform chunks (configs[4])"),
}


def rec_opts(cell, lay, act, use_cuda="True", drop=0.2):
    n = len(lay)
    o = {"_lay": ",".join(map(str, lay)), "_drop": ",".join([str(drop)] * n), "_use_laynorm_inp": "False",
         "_use_batchnorm_inp": "False", "_use_laynorm": ",".join(["False"] * n), "_use_batchnorm": ",".join(["True"] * n),
         "_bidir": "True", "_act": ",".join([act] * n), "_orthinit": "True"}
    o = {cell + k: v for k, v in o.items()}
    o.update(use_cuda=use_cuda, to_do="train")
    return o


def mlp_opts(lay, act, bn, drop, use_cuda="True"):
    j = lambda v: ",".join(map(str, v))
    return {"dnn_lay": j(lay), "dnn_drop": j(drop), "dnn_use_laynorm_inp": "False", "dnn_use_batchnorm_inp": "False",
            "dnn_use_batchnorm": j(bn), "dnn_use_laynorm": j([False] * len(lay)), "dnn_act": j(act), "use_cuda": use_cuda,
            "to_do": "train"}


def head_opts(S, use_cuda="True"):
    return mlp_opts([S], ["softmax"], [False], [0.0], use_cuda)


def sinc_opts(use_cuda="True"):
    return {"sinc_N_filt": "128,60,60,60", "sinc_len_filt": "129,5,5,3", "sinc_max_pool_len": "3,3,3,2",
            "sinc_use_laynorm_inp": "True", "sinc_use_batchnorm_inp": "False", "sinc_use_laynorm": "True,True,True,True",
            "sinc_use_batchnorm": "False,False,False,False", "sinc_act": "relu,relu,relu,relu", "sinc_drop": "0.0,0.0,0.0,0.0",
            "sinc_sample_rate": "16000", "sinc_min_low_hz": "50", "sinc_min_band_hz": "50", "use_cuda": use_cuda,
            "to_do": "train"}


CELL_CLASS = {"ligru": "liGRU", "lstm": "LSTM", "gru": "GRU", "minimalgru": "minimalGRU", "rnn": "RNN"}


def build_modules(nn_lib, c, use_cuda):
    """The architectures of one config from a `neural_networks`-compatible library (drop-in or the reference)."""
    if c["kind"] == "mlp":
        n = len(c["lay"])
        mlp = nn_lib.MLP(mlp_opts(c["lay"] + [c["S"]], ["relu"] * n + ["softmax"], [True] * n + [False], [0.15] * n + [0.0],
                                  use_cuda), c["F"])
        return [mlp]
    mods = []
    D = c["F"]
    if c["kind"] == "sinc":
        sn = nn_lib.SincNet(sinc_opts(use_cuda), c["F"])
        mods.append(sn)
        D = sn.out_dim
    net = getattr(nn_lib, CELL_CLASS[c["cell"]])(rec_opts(c["cell"], c["lay"], c["act"], use_cuda), D)
    head = nn_lib.MLP(head_opts(c["S"], use_cuda), net.out_dim)
    return mods + [net, head]


def frames_per_step(c):
    return c["N"] if c["kind"] == "mlp" else c["T"] * c["B"]


def train_flops_per_step(c):
    """Algorithmic (non-redundant) FLOPs of one training step = 3 x forward GEMM FLOPs."""
    if c["kind"] == "mlp":
        dims = [c["F"]] + c["lay"] + [c["S"]]
        return 3 * sum(2.0 * c["N"] * a * b for a, b in zip(dims[:-1], dims[1:]))
    n = c["T"] * c["B"]
    ng = {"ligru": 2, "lstm": 4, "gru": 3, "minimalgru": 2, "rnn": 1}[c["cell"]]
    fwd, D = 0.0, c["F"]
    if c["kind"] == "sinc":
        fwd += n * 1.94e8
        D = 3300
    for H in c["lay"]:
        fwd += 2.0 * n * D * ng * H + 2.0 * (2 * n) * H * ng * H
        D = 2 * H
    fwd += 2.0 * n * D * c["S"]
    return 3 * fwd


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe): ONE process, started
    by rank 0 only, before the first window."""

    def __init__(self, index, enabled=True):
        self.rows, self.proc, self.index, self.enabled = [], None, index, enabled

    def start(self):
        if not self.enabled:
            return self
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                          str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
            time.sleep(0.5)  # the process start-up happens outside every timed window
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t_begin=None, t_end=None):
        sm, mx, reasons = [], [], set()
        for ts, r in self.rows:
            if t_begin is not None and not (t_begin - 0.05 <= ts <= t_end + 0.15):
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's OWN modules on the host cores
# ---------------------------------------------------------------------------------------------


def load_reference_lib():
    """baseline/_ref/neural_networks.py = the unmodified reference file, copied there (git-ignored, never committed)
    by __graft_entry__.build() in the build container; it ships to the GPU box with the snapshot."""
    p = os.path.join(REF_DIR, "neural_networks.py")
    if not os.path.exists(p):
        return None
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_neural_networks", p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def pick_threads(c):
    """Thread count for the CPU arm.  The reference's per-step GEMMs are small ([64 x 550] x [550 x 1100] inside a Python
    time loop), so intra-op threading stops paying early: measured on this pool's 128-thread hosts (gpurun_out /
    profiles/r2_bench_reference.json) 16 threads is the fastest of {8, 16, 32, 64, 128} and 128 threads is 20x SLOWER
    (8 frames/s vs 182) — the arm therefore uses min(host threads, 16) and reports that as `cores`.  PK_REF_THREADS
    overrides.  Set regardless of OMP_NUM_THREADS (torchrun exports OMP_NUM_THREADS=1)."""
    ncpu = os.cpu_count() or 1
    forced = os.environ.get("PK_REF_THREADS")
    return max(1, min(int(forced), ncpu)) if forced else min(ncpu, 16)


def reference_step_factory(c, Ts=None):
    """(step_fn, frames_per_step, threads, description, kind) for config `c` on the CPU; Ts bounds the chunk length."""
    import torch
    threads = pick_threads(c)
    torch.set_num_threads(threads)
    step, frames, desc, kind = _reference_step(c, Ts)
    return step, frames, threads, desc.replace("@THREADS@", str(threads)), kind


def _reference_step(c, Ts):
    import torch
    ref = load_reference_lib()
    if ref is None:
        step, frames, threads, desc = port_step_factory(c, Ts)
        return step, frames, desc, "port"
    torch.manual_seed(1234)
    mods = build_modules(ref, c, "False")
    for m in mods:
        m.train()
    if c["opt"] == "rmsprop":  # utils.optimizer_init (utils.py:2121-2131): one optimizer per architecture
        opts = [torch.optim.RMSprop(m.parameters(), lr=c["lr"], alpha=0.95, eps=1e-8) for m in mods]
    else:
        opts = [torch.optim.SGD(m.parameters(), lr=c["lr"]) for m in mods]
    lossf = torch.nn.NLLLoss()
    g = torch.Generator().manual_seed(1234)
    if c["kind"] == "mlp":
        x = torch.randn(c["N"], c["F"], generator=g)
        lab = torch.randint(0, c["S"], (c["N"],), generator=g)
        frames = c["N"]
    else:
        T = Ts or c["T"]
        x = torch.randn(T, c["B"], c["F"], generator=g)
        lab = torch.randint(0, c["S"], (T * c["B"],), generator=g)
        frames = T * c["B"]

    def step():
        out = x
        if c["kind"] == "sinc":
            T_, B_ = out.shape[0], out.shape[1]
            out = mods[0](out.view(T_ * B_, -1)).view(T_, B_, -1)   # utils.py:2322-2337
            rest = mods[1:]
        else:
            rest = mods
        if c["kind"] == "mlp":
            logp = rest[0](out)
        else:
            h = rest[0](out)
            logp = rest[1](h.view(h.shape[0] * h.shape[1], -1))
        loss = lossf(logp, lab)
        err = torch.mean((torch.max(logp, dim=1)[1] != lab).float())
        for o in opts:
            o.zero_grad()
        loss.backward()
        for o in opts:
            o.step()
        return float(loss.item()), float(err.item())

    desc = (f"the reference's own modules (baseline/_ref/neural_networks.py, torch {torch.__version__} CPU fp32, "
            f"@THREADS@ of {os.cpu_count()} host threads, see pick_threads): fwd + NLLLoss + cost_err + backward + {c['opt']}")
    return step, frames, desc, "reference"


def port_step_factory(c, Ts):
    """Fallback when baseline/_ref is absent: the numpy oracle port (liGRU configs only)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import pk_oracle as orc
    if c["kind"] != "rec" or c["cell"] != "ligru":
        raise SystemExit(json.dumps({"impl": "reference", "unavailable": "baseline/_ref missing and the oracle port only covers liGRU"}))
    try:
        from threadpoolctl import threadpool_info
        threads = max([p.get("num_threads", 1) for p in threadpool_info()] or [os.cpu_count() or 1])
    except Exception:
        threads = os.cpu_count() or 1
    T, B, F, S = Ts or c["T"], c["B"], c["F"], c["S"]
    rng = np.random.default_rng(1234)
    f32 = np.float32
    layers, D = [], F
    for H in c["lay"]:
        def bn():
            return dict(weight=np.ones(H, f32), bias=np.zeros(H, f32), running_mean=np.zeros(H, f32),
                        running_var=np.ones(H, f32), eps=1e-5, momentum=0.05)
        k = 1.0 / np.sqrt(D)
        layers.append(dict(wh=rng.uniform(-k, k, (H, D)).astype(f32), wz=rng.uniform(-k, k, (H, D)).astype(f32),
                           uh=np.linalg.qr(rng.standard_normal((H, H)))[0].astype(f32),
                           uz=np.linalg.qr(rng.standard_normal((H, H)))[0].astype(f32), bh=None, bz=None,
                           bn_wh=bn(), bn_wz=bn(), act="relu", drop=0.2))
        D = 2 * H
    kh = np.sqrt(0.01 / (D + S))
    head = dict(w=rng.uniform(-kh, kh, (S, D)).astype(f32), b=np.zeros(S, f32), bn=None, ln=None, act="softmax", drop=0.0)
    x = rng.standard_normal((T, B, F)).astype(f32)
    lab = rng.integers(0, S, T * B)
    state = {"v": None}

    def step():
        masks = [(rng.random((2 * B, L["uh"].shape[0])) < 0.8).astype(f32) for L in layers]
        res = orc.ligru_model_step(x, [lab], layers, [head], masks=masks, bidir=True)
        if state["v"] is None:
            state["v"] = [dict((k, np.zeros_like(g)) for k, g in lg.items()) for lg in res["ligru_grads"]]
        for Ld, g, v in zip(layers, res["ligru_grads"], state["v"]):
            for k in ("wh", "wz", "uh", "uz"):
                Ld[k], v[k] = orc.rmsprop_step(Ld[k], g[k].astype(f32), v[k])
                Ld[k] = Ld[k].astype(f32)
        return float(res["loss"]), 0.0

    return step, T * B, threads, "numpy fp32 oracle port of the reference algorithm (oracle/pk_oracle.py): fwd+bwd+RMSprop"


def run_reference_arm(args, c, rank):
    """The reference's own CPU implementation of the path on the FULL workload of the config; rank 0 only.  Steps are
    bounded by a time budget (a 500x32x40 liGRU step takes tens of seconds of CPU), the line says how many ran."""
    if rank != 0:
        return
    step, frames, threads, desc, kind = reference_step_factory(c)
    budget = float(os.environ.get("PK_REF_BUDGET_S", "120"))
    t_start = time.perf_counter()
    # one untimed step first (allocator / thread-pool warm-up) unless a single step already eats half the budget
    t0 = time.perf_counter()
    loss, err = step()
    first = time.perf_counter() - t0
    nwarm = 1
    n, dt = 0, 0.0
    if first > 0.5 * budget:
        nwarm, n, dt = 0, 1, first   # the only affordable sample is the first step itself
    while n < args.steps and (time.perf_counter() - t_start) + (dt / n if n else first) < budget:
        t0 = time.perf_counter()
        loss, err = step()
        dt += time.perf_counter() - t0
        n += 1
    if n == 0:
        nwarm, n, dt = 0, 1, first
    val = frames * n / dt
    sample = f"{n} timed full steps ({frames} frames each) after {nwarm} warm-up; {desc}"
    out = {"impl": "reference", "metric": c["metric"], "value": val, "unit": "frames/s", "n_gpus": args.gpus,
           "steps": n, "requested_steps": args.steps, "warmup": nwarm, "ms_per_step": 1e3 * dt / n,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": c["workload"], "sample": sample},
           "cpu_baseline": {"value": val, "unit": "frames/s", "cores": threads, "kind": kind, "sample": sample},
           "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "last_loss": loss}
    print(json.dumps(out), flush=True)
    sys.stdout.flush()
    os._exit(0)   # do not wait for intra-op thread pools / allocator teardown of a 100-second CPU job


# ---------------------------------------------------------------------------------------------
# this repository's arm
# ---------------------------------------------------------------------------------------------


def parity_check(c, dev):
    """Step-0 self check: the committed full-size fixture recipe (tests/full_cases.py) run through the drop-in modules
    on this GPU; reports the loss error against the reference's own value (tests/golden/full_*.npz)."""
    if not c.get("fixture"):
        return None
    try:
        import numpy as np
        import torch
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import full_cases as fc
        import neural_networks as pknn
        z = np.load(os.path.join(ROOT, "tests", "golden", c["fixture"] + ".npz"))
        net, head = fc.build(pknn, c["fixture"], use_cuda="True")
        x, lab = fc.inputs(c["fixture"])
        net.to(dev).train()
        head.to(dev).train()
        torch.manual_seed(fc.forward_seed(c["fixture"]))
        with torch.no_grad():
            h = net(x.to(dev))
            logp = head(h.view(h.shape[0] * h.shape[1], -1))
            loss = torch.nn.functional.nll_loss(logp, lab.to(dev)).item()
            rows = logp[::fc.ROW_STRIDE].cpu().numpy()
        ref = float(z["loss"])
        e_rows = float(np.max(np.abs(rows - z["logp_rows"])) / np.max(np.abs(z["logp_rows"])))
        del net, head, h, logp
        torch.cuda.empty_cache()
        return {"fixture": f"tests/golden/{c['fixture']}.npz (reference neural_networks.py, CPU fp32, same seeds / masks)",
                "loss": loss, "loss_ref": ref, "loss_rel_err": abs(loss - ref) / abs(ref), "logp_max_rel_err": e_rows,
                "bar": 1e-3}
    except Exception as e:  # the self check must never take the benchmark down
        return {"error": repr(e)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--repeats", type=int, default=5)
    ap.add_argument("--config", default="ligru5x550", choices=list(CONFIGS))
    ap.add_argument("--impl", default="pk", choices=["pk", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"],
                    help="replay the step from a CUDA graph (pk_train.GraphedStep); auto = on for single-GPU runs")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    c = CONFIGS[args.config]

    if args.impl == "reference":
        run_reference_arm(args, c, rank)
        return

    import torch
    import torch.distributed as dist

    import neural_networks as pknn
    import pk_functions as pkf
    import pk_native as pk
    import pk_train

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    args.warmup = max(args.warmup, 3)

    parity = None
    if rank == 0 and not args.no_parity:
        parity = parity_check(c, dev)

    torch.manual_seed(1234)
    mods = [m.cuda().train() for m in build_modules(pknn, c, "True")]
    for m in mods:
        if hasattr(m, "fast_dropout"):
            m.fast_dropout = True  # masks drawn with the device generator (same Bernoulli(1-p) per layer per step)
    if world > 1:  # identical replicas
        for m in mods:
            for p in m.parameters():
                dist.broadcast(p.data, 0)
    trainer = pk_train.FlatTrainer(mods, opt=c["opt"], lr=c["lr"], alpha=0.95, eps=1e-8)

    # synthetic inputs in the reference's chunk layout (label = last column, stored as float, data_io.py:272).
    # A ring of chunks is rotated and 256 MiB are written between steps, so no step finds its input in L2.
    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    nring = 8
    if c["kind"] == "mlp":
        shape, nfea = (c["N"], c["F"]), c["F"]
        host = [torch.cat([torch.randn(*shape, generator=g), torch.randint(0, c["S"], (c["N"], 1), generator=g).float()], 1)
                .pin_memory() for _ in range(nring)]
    else:
        shape, nfea = (c["T"], c["B"], c["F"]), c["F"]
        host = [torch.cat([torch.randn(*shape, generator=g), torch.randint(0, c["S"], (c["T"], c["B"], 1), generator=g).float()], 2)
                .pin_memory() for _ in range(nring)]
    devchunks = [h.to(dev) for h in host]
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)  # 256 MiB > 126 MB L2

    def one_step(inp):
        if c["kind"] == "rec":
            return pk_train.chunk_step(mods[0], mods[1], trainer, inp, nfea)
        if c["kind"] == "mlp":
            lab = inp[:, nfea].long()
            trainer.zero_grad()
            logp = mods[0](inp[:, :nfea])
            loss = torch.nn.functional.nll_loss(logp, lab)
            loss.backward()
            trainer.step()
            return loss.detach(), None
        # sinc: raw chunk -> SincNet on [T*B, 3200] -> [T, B, 3300] -> liGRU -> head (utils.py:2322-2337)
        T_, B_ = inp.shape[0], inp.shape[1]
        lab = inp[:, :, nfea].reshape(-1).long()
        trainer.zero_grad()
        f = mods[0](inp[:, :, :nfea].reshape(T_ * B_, nfea)).view(T_, B_, -1)
        out = mods[1](f)
        loss, err, _ = pkf.HeadNLLFn.apply(out.view(T_ * B_, -1), mods[2].wx[0].weight, mods[2].wx[0].bias, lab)
        loss.backward()
        trainer.step()
        return loss.detach(), err

    # launch-bound recipes replay the whole step from a CUDA graph (public API: pk_train.GraphedStep)
    use_graph, graph_note = False, "off"
    eager_step = one_step
    # auto: single-GPU runs replay the step from a graph (the MLP recipe is launch-bound: 2.2x; the recurrent recipes gain
    # 2 % on `value` and 7 % end to end); multi-GPU runs stay eager unless --graph on (the NCCL allreduce would be captured)
    if args.graph == "on" or (args.graph == "auto" and world == 1):
        try:
            graphed = pk_train.GraphedStep(lambda inp: eager_step(inp), devchunks[0])
            one_step = lambda inp: graphed(inp)   # noqa: E731
            use_graph, graph_note = True, "whole step replayed from one CUDA graph (pk_train.GraphedStep)"
        except Exception as e:  # capture is an optimisation, never a requirement
            graph_note = f"capture failed, eager launches: {e!r}"[:200]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(nsteps, from_host):
        losses = []
        for i in range(nsteps):
            flush.zero_()  # evict the previous step's working set from L2
            inp = host[i % nring].to(dev, non_blocking=True) if from_host else devchunks[i % nring]
            loss, _ = one_step(inp)
            losses.append(loss.item() if from_host else loss)  # e2e: D2H read of the step's result (core.py:689)
        return losses

    def window(nsteps, from_host):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = pk.launch_count
        e0.record()
        losses = run(nsteps, from_host)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item(), pk.launch_count - n0, losses

    clk = ClockSampler(local_rank, enabled=(rank == 0)).start()
    run(args.warmup, False)
    t_begin = time.time()
    dev_windows = [window(args.steps, False) for _ in range(args.repeats)]
    t_end = time.time()
    run(2, True)
    e2e_windows = [window(args.steps, True) for _ in range(args.repeats)]
    clk.stop()
    clocks = clk.summary(t_begin, t_end) if rank == 0 else None

    ms_dev = statistics.median(w[0] for w in dev_windows)
    ms_e2e = statistics.median(w[0] for w in e2e_windows)
    launches = dev_windows[0][1]
    if use_graph:   # replays do not pass through the Python entry points: kernels recorded in the graph x steps replayed
        launches = graphed.launches_per_replay * args.steps
    losses = [float(x) for x in dev_windows[0][2]] if not use_graph else [float(e2e_windows[0][2][0]), float(e2e_windows[-1][2][-1])]
    frames = frames_per_step(c) * args.steps * world
    value = frames / (ms_dev * 1e-3)
    e2e_value = frames / (ms_e2e * 1e-3)

    # ---- separate pass: per-launch CUDA-event timing of the dominant kernel (outside every reported window)
    dom = {"rec": "rnn_layer_bwd", "sinc": "rnn_layer_bwd", "mlp": "gemm_tn"}[c["kind"]]
    stepwise = c["kind"] != "mlp" and (c["cell"] != "ligru" or max(c["lay"]) > pkf.PERSISTENT_MAX_H)
    if stepwise:
        dom = "rnn_step_bwd"
    events = []
    orig = getattr(pk, dom)

    def timed_call(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = orig(*a, **k)
        e1.record()
        events.append((e0, e1))
        return r

    setattr(pk, dom, timed_call)
    one_step = eager_step   # a replayed graph does not pass through the Python entry points: time the eager launches
    run(4, False)
    setattr(pk, dom, orig)
    torch.cuda.synchronize()
    k_ms = [a.elapsed_time(b) for a, b in events]
    k_avg = sum(k_ms) / max(len(k_ms), 1)
    per_step = len(k_ms) / 4.0
    pk_peaks, how = peaks()
    peak = pk_peaks.get("bf16_tflops_sustained", 1400.0)
    step_ms = ms_dev / args.steps
    if c["kind"] == "mlp":
        flops_launch = train_flops_per_step(c) / max(per_step, 1)
        kname = "gemm_tn_kernel / gemm_tn_persist_kernel (tcgen05 GEMMs of the MLP layers; average launch)"
        extra = {}
    else:
        H = c["lay"][0]
        ng = {"ligru": 2, "lstm": 4, "gru": 3, "minimalgru": 2, "rnn": 1}[c["cell"]]
        flops_launch = 2.0 * c["T"] * (2 * c["B"]) * (ng * H) * H  # U^T [gate gradients] for every row of the direction-stacked batch
        cluster = stepwise and c["cell"] == "lstm" and pk.rnn_step_is_cluster(pk.CELL_LSTM, H)
        kname = ("lstm_cluster_bwd_kernel (cluster-persistent reverse-time recurrence: weights stationary in shared memory, "
                 "K-split reduce-scatter over distributed shared memory) + weight packing" if cluster else
                 "cell_bwd_persist_kernel (step-wise reverse-time recurrence)" if stepwise else
                 "ligru_bwd_ws_kernel (persistent reverse-time recurrence, register-stationary mma.sync; the faster of the "
                 "two kernel families at this H)" if H <= 560 else
                 "ligru_bwd_tc_kernel (persistent reverse-time recurrence on tcgen05, weights stationary in TMEM)")
        extra = {"us_per_recurrent_step": 1e3 * k_avg / c["T"]}
    achieved = flops_launch / (k_avg * 1e-3) / 1e12 if k_avg > 0 else 0.0
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r2_ncu_traffic.json"))).get(args.config)
    except Exception:
        pass
    roofline = {"kernel": kname, "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak, "peak_source": f"{how} bf16 sustained (fp16 operands, fp32 accumulate)",
                "traffic": traffic, "traffic_unit": "bytes/launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, profiles/)",
                "avg_launch_ms": k_avg, "launches_timed": len(k_ms), "launches_per_step": per_step,
                "share_of_step": (k_avg * per_step) / step_ms if step_ms > 0 else None,
                "whole_step_tflops": train_flops_per_step(c) / (step_ms * 1e-3) / 1e12,
                "whole_step_frac": train_flops_per_step(c) / (step_ms * 1e-3) / 1e12 / peak, **extra}

    out = {"metric": c["metric"], "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "repeats": args.repeats, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f16 operands / f32 accumulate+state", "data": "synthetic",
           "config": {"workload": c["workload"], "name": args.config, "global_batch": (c.get("B") or c["N"]) * world,
                      "seq_len": c.get("T"), "parallelism": f"dp{world}",
                      "optimizer": f"{c['opt']} lr {c['lr']} (utils.py:2106-2164)",
                      "l2": "256 MiB flush between steps + 8-chunk input ring",
                      "dropout": "device-drawn Bernoulli masks (fast_dropout)", "cuda_graph": graph_note,
                      "timing": f"median of {args.repeats} windows x {args.steps} steps, CUDA events, max over ranks"},
           "windows_ms": [w[0] for w in dev_windows],
           "clocks": clocks, "gpu_launches": launches,
           "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": host[0].numel() * 4,
                   "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps, "windows_ms": [w[0] for w in e2e_windows]},
           "roofline": roofline, "parity": parity,
           "loss_first_last": [float(losses[0]), float(losses[-1])]}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # bounded sample: a quarter-length chunk (all utterance columns) through the reference's own modules
        Ts = None if c["kind"] == "mlp" else max(c["T"] // 4, 1)
        step, fr, threads, desc, kind = reference_step_factory(c, Ts)
        t0 = time.perf_counter()
        step()
        first = time.perf_counter() - t0
        n, dt = 0, 0.0
        while (n < 1 and first < 20.0) or (n >= 1 and dt + dt / n < 12.0 and n < 200):
            t0 = time.perf_counter()
            step()
            dt += time.perf_counter() - t0
            n += 1
        warm = 1
        if n == 0:   # one step already exceeds the sample budget: it IS the sample
            n, dt, warm = 1, first, 0
        out["cpu_baseline"] = {"value": fr * n / dt, "unit": "frames/s", "cores": threads, "kind": kind,
                               "sample": f"{n} steps of {fr} frames ({'minibatch' if Ts is None else f'{Ts}-frame sub-chunk, all columns'}) "
                                         f"after {warm} warm-up; {desc}"}
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
