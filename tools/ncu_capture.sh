#!/bin/bash
# Run on the GPU box (via gpurun): plain run, launch list, and one --set full capture of K1 + K2.
set -u
CMD="python bench.py --steps 1 --warmup 3 --m 113664 --no-cpu-baseline"
mkdir -p gpurun_out
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"trmm_sqnorm|kcross_mean" -s 6 -c 2 -f -o gpurun_out/prof_k1k2 $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"
tail -3 gpurun_out/plain.log
