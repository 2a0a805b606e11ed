#!/usr/bin/env python
"""Benchmark of the hot path BASELINE.json names: training frames/sec of a 5 x 550 bidirectional liGRU
+ 1936-way softmax head on synthetic 500 x 32 x 40 chunks (cfg/TIMIT_baselines/TIMIT_liGRU_fmllr.cfg shape).

    python bench.py --gpus N --steps K --warmup W            # this repository's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU

One "step" = one minibatch exactly as core.run_nn runs it (reference core.py:616-642): forward_model
(liGRU stack -> softmax head -> NLLLoss + frame error), zero_grad, backward, RMSprop step (+ one NCCL
gradient allreduce when N > 1; every rank trains on its own 500 x 32 chunk: weak scaling).

Prints ONE JSON line (rank 0).  `value` = frames/s with the chunk already resident in HBM;
`e2e` = the same through the drop-in module API from pinned HOST buffers (H2D of the chunk and D2H of
the loss inside the timed region).  `roofline` describes the dominant kernel (the reverse-time
persistent recurrent kernel), timed live with CUDA events on the launching stream.
`cpu_baseline` = the oracle port of the reference algorithm timed on this box's host cores on a bounded
sample (the reference itself is pure Python + PyTorch under /root/reference, which does not exist on
the GPU box).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "pytorch-kaldi_b200"))

T, B, F, H, L, S = 500, 32, 40, 550, 5, 1936
METRIC = "frames/sec liGRU-5x550 on 500x32x40 chunks @1/2/4/8 B200; loss match <=1e-3"
WORKLOAD = "TIMIT liGRU 5x550 bidir, fMLLR-40, synthetic 500x32x40 chunks, 1936 senones (configs[1])"


def ligru_opts():
    return {"ligru_lay": ",".join([str(H)] * L), "ligru_drop": ",".join(["0.2"] * L),
            "ligru_use_laynorm_inp": "False", "ligru_use_batchnorm_inp": "False",
            "ligru_use_laynorm": ",".join(["False"] * L), "ligru_use_batchnorm": ",".join(["True"] * L),
            "ligru_bidir": "True", "ligru_act": ",".join(["relu"] * L), "ligru_orthinit": "True",
            "use_cuda": "True", "to_do": "train"}


def head_opts():
    return {"dnn_lay": str(S), "dnn_drop": "0.0", "dnn_use_laynorm_inp": "False", "dnn_use_batchnorm_inp": "False",
            "dnn_use_batchnorm": "False", "dnn_use_laynorm": "False", "dnn_act": "softmax", "use_cuda": "True",
            "to_do": "train"}


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def __enter__(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                          str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
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
# reference arm / cpu_baseline: the oracle port of the reference algorithm on the host cores
# ---------------------------------------------------------------------------------------------


def cpu_reference_step_factory(Ts, Bs):
    """Builds the reference algorithm (oracle/pk_oracle.py, fp32 like the reference) for a bounded sample
    [Ts, Bs, 40] of the workload and returns (step_fn, frames_per_step, threads)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import pk_oracle as orc
    try:
        from threadpoolctl import threadpool_info
        threads = max([p.get("num_threads", 1) for p in threadpool_info()] or [os.cpu_count() or 1])
    except Exception:
        threads = os.cpu_count() or 1
    rng = np.random.default_rng(1234)
    f32 = np.float32
    layers, D = [], F
    for _ in range(L):
        def bn():
            return dict(weight=np.ones(H, f32), bias=np.zeros(H, f32), running_mean=np.zeros(H, f32),
                        running_var=np.ones(H, f32), eps=1e-5, momentum=0.05)
        k = 1.0 / np.sqrt(D)
        layers.append(dict(wh=rng.uniform(-k, k, (H, D)).astype(f32), wz=rng.uniform(-k, k, (H, D)).astype(f32),
                           uh=np.linalg.qr(rng.standard_normal((H, H)))[0].astype(f32),
                           uz=np.linalg.qr(rng.standard_normal((H, H)))[0].astype(f32), bh=None, bz=None,
                           bn_wh=bn(), bn_wz=bn(), act="relu", drop=0.2))
        D = 2 * H
    kh = np.sqrt(0.01 / (2 * H + S))
    head = dict(w=rng.uniform(-kh, kh, (S, 2 * H)).astype(f32), b=np.zeros(S, f32), bn=None, ln=None, act="softmax",
                drop=0.0)
    x = rng.standard_normal((Ts, Bs, F)).astype(f32)
    lab = rng.integers(0, S, Ts * Bs)
    state = {"v": None}

    def step():
        masks = [(rng.random((2 * Bs, H)) < 0.8).astype(f32) for _ in range(L)]
        res = orc.ligru_model_step(x, [lab], layers, [head], masks=masks, bidir=True)
        # RMSprop step on every parameter (utils.py:2121-2128: lr 4e-4, alpha .95, eps 1e-8)
        if state["v"] is None:
            state["v"] = [dict((k, np.zeros_like(g)) for k, g in lg.items()) for lg in res["ligru_grads"]]
        for Ld, g, v in zip(layers, res["ligru_grads"], state["v"]):
            for k in ("wh", "wz", "uh", "uz"):
                Ld[k], v[k] = orc.rmsprop_step(Ld[k], g[k].astype(f32), v[k])
                Ld[k] = Ld[k].astype(f32)
        return float(res["loss"])

    return step, Ts * Bs, threads


def run_reference_arm(args, rank):
    if rank != 0:
        return
    Ts, Bs = 100, B  # bounded sample: the first 100 frames of every utterance of one 500 x 32 chunk
    step, frames, threads = cpu_reference_step_factory(Ts, Bs)
    for _ in range(min(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = frames * args.steps / dt
    sample = f"{Ts}x{Bs}x{F} sub-chunk per step (full model 5x550 bidir + {S} head, fwd+bwd+RMSprop), numpy fp32"
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": "frames/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * dt / args.steps,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": WORKLOAD, "sample": sample},
           "cpu_baseline": {"value": val, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample},
           "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


# ---------------------------------------------------------------------------------------------
# this repository's arm
# ---------------------------------------------------------------------------------------------


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="pk", choices=["pk", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank)
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

    torch.manual_seed(1234)
    net = pknn.liGRU(ligru_opts(), F).cuda().train()
    head = pknn.MLP(head_opts(), net.out_dim).cuda().train()
    net.fast_dropout = True  # masks drawn with the device generator (same Bernoulli(1-p) per layer per step)
    if world > 1:  # identical replicas
        for p in list(net.parameters()) + list(head.parameters()):
            dist.broadcast(p.data, 0)
    trainer = pk_train.FlatTrainer([net, head], opt="rmsprop", lr=0.0004, alpha=0.95, eps=1e-8)

    # synthetic chunks in the reference's layout [T, B, F + 1] (label = last column, stored as float).
    # A ring of chunks larger than L2 is rotated through so no step finds its input cached.
    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    nring = 8
    host = [torch.cat([torch.randn(T, B, F, generator=g), torch.randint(0, S, (T, B, 1), generator=g).float()], 2)
            .pin_memory() for _ in range(nring)]
    devchunks = [h.to(dev) for h in host]
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)  # 256 MiB > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # per-launch timing of the dominant kernel (reverse-time recurrent kernel) with CUDA events
    rec_events = []
    orig_bwd = pk.rnn_layer_bwd

    def timed_bwd(*a, **k):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        orig_bwd(*a, **k)
        e1.record()
        rec_events.append((e0, e1))

    def run(nsteps, from_host, time_kernel=False):
        losses = []
        if time_kernel:
            pkf.pk.rnn_layer_bwd = timed_bwd
        for i in range(nsteps):
            flush.zero_()  # evict the previous step's working set from L2
            if from_host:
                inp = host[i % nring].to(dev, non_blocking=True)
            else:
                inp = devchunks[i % nring]
            loss, err = pk_train.chunk_step(net, head, trainer, inp, F)
            if from_host:
                losses.append(loss.item())  # D2H read of the step's result (as core.py:689 does every batch)
            else:
                losses.append(loss)
        if time_kernel:
            pkf.pk.rnn_layer_bwd = orig_bwd
        return losses

    def timed(nsteps, from_host, time_kernel=False):
        barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        n0 = pk.launch_count
        e0.record()
        losses = run(nsteps, from_host, time_kernel)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item(), pk.launch_count - n0, losses

    run(args.warmup, False)
    with ClockSampler(local_rank) as clk:
        ms_dev, launches, losses = timed(args.steps, False, time_kernel=True)
    clocks = clk.summary()
    run(2, True)
    ms_e2e, _, losses_e2e = timed(args.steps, True)

    frames = T * B * args.steps * world
    value = frames / (ms_dev * 1e-3)
    e2e_value = frames / (ms_e2e * 1e-3)

    # roofline of the dominant kernel: algorithmic FLOPs per launch = T steps x (2B rows) x [2H x H] x 2
    # (U^T [da; dpz] for every row of the direction-stacked batch), DESIGN.md "Kernels"
    torch.cuda.synchronize()
    k_ms = [a.elapsed_time(b) for a, b in rec_events]
    k_avg = sum(k_ms) / max(len(k_ms), 1)
    flops_launch = 2.0 * T * (2 * B) * (2 * H) * H
    pk_peaks, how = peaks()
    peak = pk_peaks.get("bf16_tflops_sustained", 1400.0)
    achieved = flops_launch / (k_avg * 1e-3) / 1e12 if k_avg > 0 else 0.0
    roofline = {"kernel": "ligru_bwd_ws_kernel (persistent reverse-time recurrence, 5 launches/step)", "bound": "tensor",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "peak_source": f"{how} bf16 sustained (fp16 operands, fp32 accumulate)", "traffic": 348441088, "traffic_unit": "bytes/launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, profiles/r1_ncu_ligru_ws_kernels.txt)",
                "avg_launch_ms": k_avg, "launches_timed": len(k_ms),
                "share_of_step": (k_avg * L) / (ms_dev / args.steps) if ms_dev > 0 else None,
                "us_per_recurrent_step": 1e3 * k_avg / T}

    out = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f16 operands / f32 accumulate+state", "data": "synthetic",
           "config": {"workload": WORKLOAD, "global_batch": B * world, "seq_len": T, "parallelism": f"dp{world}",
                      "optimizer": "rmsprop lr 4e-4 alpha .95 (utils.py:2121)", "l2": "256 MiB flush between steps + "
                      "8-chunk input ring", "dropout": "device-drawn Bernoulli masks (fast_dropout)"},
           "clocks": clocks, "gpu_launches": launches,
           "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": host[0].numel() * 4,
                   "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
           "roofline": roofline,
           "loss_first_last": [float(losses[0]), float(losses[-1])]}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        step, fr, threads = cpu_reference_step_factory(50, B)
        step()
        t0 = time.perf_counter()
        n = 0
        while time.perf_counter() - t0 < 12.0 or n < 2:
            step()
            n += 1
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": fr * n / dt, "unit": "frames/s", "cores": threads, "kind": "port",
                               "sample": f"{n} steps of a 50x{B}x{F} sub-chunk (full 5x550 bidir model + {S} head, "
                                         "fwd+bwd+RMSprop) with the numpy fp32 oracle port"}
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
