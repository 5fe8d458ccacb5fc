#!/bin/bash
# One gpurun call that re-establishes the measured state of the repository on a fresh B200 box:
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_round_check.sh'
# 1. the GPU suite, 1b. the host-level tests against the real library instead of the ABI model, 2. smoke(), 3. the default
# bench line, 3b. BASELINE.json configs[0] / configs[1] as whole gp_minimize runs (host and GPU fit), 4. the launch list of the same bench command (ncu, after the plain run has exited 0).  Everything lands in
# gpurun_out/; copy what should be judged into profiles/.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/t_all.log 2>&1; echo "pytest -m gpu rc=$?"; tail -3 gpurun_out/t_all.log
SKB_TESTS_ON_DEVICE=1 timeout 600 python -m pytest tests/test_host_optimizer.py -q > gpurun_out/t_host_on_device.log 2>&1; echo "host-level tests on the device rc=$?"; tail -2 gpurun_out/t_host_on_device.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py > gpurun_out/bench_head.json 2> gpurun_out/bench_head.err; rc=$?; echo "bench rc=$rc"
timeout 300 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "reference arm rc=$?"
for cfg in branin hart6; do for fit in host gpu; do
    timeout 400 python tools/bench_gp_minimize.py $cfg --runs 1 --fit-device $fit >> gpurun_out/gp_minimize_runs.jsonl 2>> gpurun_out/gp_minimize_runs.err
done; done; echo "whole-run configs[0]/[1]:"; cut -c1-300 gpurun_out/gp_minimize_runs.jsonl
if [ "$rc" = 0 ]; then
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
        python bench.py --steps 2 --warmup 3 --no-ask --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
    echo "launch list rc=$?"
fi
cut -c1-400 gpurun_out/bench_head.json
