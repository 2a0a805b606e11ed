#!/bin/bash
# round-2 end-of-round validation on the GPU box (one gpurun call): full GPU suite, smoke, bench lines of configs 2 and 3,
# GEMM vs cuBLAS, launch list of the LSTM step.  Outputs land in gpurun_out/.
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu > gpurun_out/final_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/final_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/final_bench_ligru5x550.json 2> gpurun_out/final_bench_ligru5x550.err; echo "bench2 rc=$?"
timeout 240 python bench.py --config lstm4x550 --steps 20 --warmup 3 > gpurun_out/final_bench_lstm4x550.json 2> gpurun_out/final_bench_lstm4x550.err; echo "bench3 rc=$?"
timeout 60 python tools/gemm_vs_cublas.py > gpurun_out/final_gemm_vs_cublas.log 2>&1; echo "gemm rc=$?"
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/final_launches_lstm.csv \
    python bench.py --config lstm4x550 --steps 2 --warmup 3 --repeats 1 --no-cpu-baseline --no-parity > gpurun_out/final_lstm_under_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import json
for n in ("ligru5x550", "lstm4x550"):
    try:
        d = json.loads(open(f"gpurun_out/final_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, round(d["ms_per_step"], 3), round(d["value"]), round(d["e2e"]["value"]), d["parity"].get("logp_max_rel_err"), d["gpu_launches"], d["clocks"])
    except Exception as e:
        print(n, "ERR", e)
PY
cat gpurun_out/final_gemm_vs_cublas.log
