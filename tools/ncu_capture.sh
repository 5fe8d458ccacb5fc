#!/bin/bash
# Run on the GPU box (via gpurun): plain run, launch list, and one --set full capture of the hot kernels.
#   value path   : bench.py at one scratch chunk (113,664 candidates)  -> K1, K2, top-k
#   gradient path: tools/bench_grad_once.py (n=2048, d=20, 37,888 candidates) -> K1, K2d, K3g
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --m 113664 --no-cpu-baseline --no-ask"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"trmm_kernel|kcross_mean" -s 6 -c 2 -f -o gpurun_out/prof_k1k2 $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture (value path) rc=$?"
CMD2="python tools/bench_grad_once.py"
$CMD2 > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"trmm_kernel|grad_contract" -s 4 -c 2 -f -o gpurun_out/prof_grad $CMD2 > gpurun_out/ncu_full_grad.log 2>&1
echo "full capture (gradient path) rc=$?"
tail -2 gpurun_out/plain.log | cut -c1-400
tail -2 gpurun_out/plain3.log
