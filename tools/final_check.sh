#!/bin/bash
# round-end sanity: full GPU test suite, smoke, 1-GPU bench
timeout 500 python -m pytest tests -q -m gpu 2>&1 | tail -3
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1e.json 2> gpurun_out/bench_r1e.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r1e.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["us_per_recurrent_step"], d["gpu_launches"], d["clocks"])
PY
