timeout 300 python -m pytest tests -q -m gpu -x -k "fused_head or matches_reference or mlp" 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1d.json 2> gpurun_out/bench_r1d.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r1d.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["us_per_recurrent_step"], d["gpu_launches"])
PY
