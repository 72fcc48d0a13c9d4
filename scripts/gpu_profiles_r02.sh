#!/bin/bash
# Round-2 profile session (one ncu-bearing gpurun call): launch list of the bench, full captures of the headline RK4 kernel,
# the Lindblad chain stage kernel (TMA), the general state-vector steppers (d = 9).  Every ncu run follows the same command
# run plain.
set -x
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs"
$B > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_bench.csv $B > gpurun_out/ncu_l.log 2>&1
$B > gpurun_out/plain_bench2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sv_rk4_small2 -s 3 -c 1 -o gpurun_out/r02_prof_rk4 -f $B > gpurun_out/ncu_rk4.log 2>&1
CASQ_LB_PARTS=1 python scripts/gpu_time_lindblad.py 128 > gpurun_out/plain_lb.log 2>&1 &&
CASQ_LB_PARTS=1 ncu --set full --clock-control none --import-source on -k regex:lindblad_chain_tma -s 240 -c 4 -o gpurun_out/r02_prof_lindblad_tma -f python scripts/gpu_time_lindblad.py 128 > gpurun_out/ncu_lb.log 2>&1
python scripts/gpu_time_general.py 4096 > gpurun_out/plain_gen.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"sv_rk4_general|sv_dopri5_general" -c 2 -o gpurun_out/r02_prof_general -f python scripts/gpu_time_general.py 4096 > gpurun_out/ncu_gen.log 2>&1
cat gpurun_out/plain_lb.log gpurun_out/plain_gen.log
tail -2 gpurun_out/plain_bench.log
