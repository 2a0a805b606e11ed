#!/bin/bash
# Round-2 profile captures (GPU box, under gpurun): launch list of the benchmark step and one full ncu capture of each
# recurrent kernel (tcgen05 and mma.sync variants, forward and reverse time).  Outputs land in gpurun_out/.
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 450 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --repeats 1 --no-cpu-baseline --no-parity > gpurun_out/r2_bench_under_ncu.log 2>&1
for k in ligru_fwd_ws ligru_bwd_ws ligru_fwd_tc ligru_bwd_tc; do
  ncu --set full --clock-control none --import-source on -k regex:${k}_kernel -s 1 -c 1 -f -o gpurun_out/r2_prof_${k} \
      ./pytorch-kaldi_b200/pk_selftest bench > gpurun_out/r2_ncu_${k}.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
