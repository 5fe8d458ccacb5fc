#!/bin/bash
# Round-2 profiling pass on the bench's own command (B200_PROFILING.md): plain run first, then the launch list
# (gpu__time_duration per launch), then ONE --set full capture of a full-size K1 and K2o launch of the headline path.
set -u
mkdir -p gpurun_out
CMD="python bench.py --legs none --no-ask --no-cpu-baseline --steps 2 --warmup 3"
$CMD > gpurun_out/r02_plain_bench.json 2> gpurun_out/r02_plain_bench.err || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench.csv $CMD > gpurun_out/r02_ncu_launches.log 2>&1
echo "launch list rc=$?"
# full-size launches only: skip the self-check's small ones (the first 12 matching launches)
ncu --set full --clock-control none --import-source on -k regex:"ozaki_trmm|kcross_digits" -s 12 -c 2 -f -o gpurun_out/r02_prof_split $CMD > gpurun_out/r02_ncu_full.log 2>&1
echo "full capture rc=$?"
ncu -i gpurun_out/r02_prof_split.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_split_raw.csv 2>/dev/null
ncu -i gpurun_out/r02_prof_split.ncu-rep --page details --csv > gpurun_out/r02_ncu_full_split_details.csv 2>/dev/null
ls -la gpurun_out/r02_*
