#!/bin/bash
# Round-2 closing session on one GPU: smoke, both bench arms, then the ncu launch list of the bench and a full capture of
# the restructured Dopri5 kernel (each ncu run follows the same command run plain).
set -x
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_n1.json 2> gpurun_out/bench_ref.err; tail -c 600 gpurun_out/r02_bench_reference_n1.json
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/bench.err; tail -c 400 gpurun_out/r02_bench_n1.json; tail -3 gpurun_out/bench.err
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv $B > gpurun_out/ncu_l.log 2>&1
python scripts/gpu_time_dopri5.py > gpurun_out/plain_dp5.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sv_dopri5_kernel -s 1 -c 1 -o gpurun_out/r02_prof_dopri5 -f python scripts/gpu_time_dopri5.py > gpurun_out/ncu_dp5.log 2>&1
cat gpurun_out/plain_dp5.log
