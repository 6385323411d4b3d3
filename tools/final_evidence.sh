#!/bin/bash
# One gpurun call: GPU tests, the driver-style bench (both arms), the ncu launch list of the same bench command and one
# `--set full` capture of the three kernels of the headline step.  usage: tools/final_evidence.sh <session tag>
s=$1
set -x
python -m pytest tests -x -q -m gpu > gpurun_out/r2_gputests_$s.log 2>&1; tail -3 gpurun_out/r2_gputests_$s.log
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r02_bench_ref_$s.json 2> gpurun_out/r02_bench_ref_$s.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_$s.json 2> gpurun_out/r02_bench_$s.err
python bench.py --steps 20 --warmup 5 --no-extra-configs --no-cpu-baseline > gpurun_out/r02_bench_plain_$s.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launches_$s.csv python bench.py --steps 20 --warmup 5 --no-extra-configs --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
PB_CASE=cfg2 ncu --set full --clock-control none --import-source on -k regex:pb_ -s 45 -c 6 -o gpurun_out/prof_r2_cfg2_$s -f python tools/ncu_target.py > gpurun_out/ncu_c2.log 2>&1; tail -2 gpurun_out/ncu_c2.log
ls -la gpurun_out/*_$s*
