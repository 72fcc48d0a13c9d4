#!/bin/bash
# One GPU session: tests, smoke, bench (both arms), config timings, then ncu launch list + full captures.
set -x
mkdir -p gpurun_out
python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
python scripts/bench_configs.py --out gpurun_out/configs.json > gpurun_out/configs.log 2>&1; tail -5 gpurun_out/configs.log
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sv_rk4_small -s 3 -c 1 -o gpurun_out/prof_rk4 -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
python scripts/gpu_time_lindblad.py 64 > gpurun_out/plain3.log 2>&1 &&
CASQ_LB_PARTS=1 python scripts/gpu_time_lindblad.py 64 > gpurun_out/plain3b.log 2>&1 &&
CASQ_LB_PARTS=1 ncu --set full --clock-control none --import-source on -k regex:"lindblad_stage|lindblad_coef" -s 210 -c 5 -o gpurun_out/prof_lindblad -f \
    python scripts/gpu_time_lindblad.py 64 > gpurun_out/ncu_lb.log 2>&1
python scripts/gpu_time_expm.py > gpurun_out/plain4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:propagator_expm -c 1 -o gpurun_out/prof_expm -f \
    python scripts/gpu_time_expm.py > gpurun_out/ncu_expm.log 2>&1
python scripts/gpu_time_dopri5.py > gpurun_out/plain5.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sv_dopri5 -c 1 -o gpurun_out/prof_dopri5 -f \
    python scripts/gpu_time_dopri5.py > gpurun_out/ncu_dp5.log 2>&1
cat gpurun_out/plain3.log gpurun_out/plain4.log gpurun_out/plain5.log
ls -la gpurun_out | head -30
