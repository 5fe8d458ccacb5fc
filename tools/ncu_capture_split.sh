#!/bin/bash
# ncu --set full capture of the split-integer path kernels (K1 digits + K2o) at n_train=2048, d=20.
set -u
mkdir -p gpurun_out
CMD="python tools/bench_split_once.py"
$CMD > gpurun_out/plain_split.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"ozaki_trmm|kcross_digits" -s 4 -c 2 -f -o gpurun_out/prof_split $CMD > gpurun_out/ncu_split.log 2>&1
echo "full capture (split path) rc=$?"
tail -1 gpurun_out/plain_split.log
