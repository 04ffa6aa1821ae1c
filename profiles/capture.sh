#!/bin/bash
# Round-1 profile capture (run on the GPU box through gpurun from the repo root):
#   1. bench line of the default path                      -> gpurun_out/bench_final.json
#   2. ncu launch list of a short bench run (3 frames)     -> gpurun_out/launches_final.csv
#   3. ncu --set full of the persistent grouped kernel     -> gpurun_out/prof_phase.ncu-rep
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_final.log 2>&1
tail -1 gpurun_out/bench_final.log > gpurun_out/bench_final.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final.csv \
    python tests/gpu_phase_time.py 2 1024 6 > gpurun_out/launches_final.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rc_tc_phase -s 9 -c 3 -o gpurun_out/prof_phase \
    python tests/gpu_phase_time.py 2 1024 6 > gpurun_out/prof_phase.log 2>&1
ls -la gpurun_out
