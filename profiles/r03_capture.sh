#!/bin/bash
# Round-2, second half: capture for the 256 x 256 CTA-pair variant of the grouped kernel (RC_PH_PAIR=2, opt-in) next to the default
# 128 x 128 kernel.  Run on the GPU box through gpurun from the repo root; outputs under gpurun_out/, summary in profiles/r03_pair256.md.
#   1. tile trace of the three phases of one frame (B = 1024, mixed)                                  -> r03_trace_pair256.log
#   2. same-box A/B: default kernel / pair256 (256-column tiles, updater phase in 128-column tiles) / pair256 all 128-column
#      tiles / CTA-pair kernel of round 1 (256 x 128 tiles), 40 frames each, pose difference vs the default             -> r03_sweep.log
#   3. ncu --set full of two launches of the pair256 kernel                                                           -> r03_pair256.ncu-rep
set -x
mkdir -p gpurun_out
RC_PH_PAIR=2 python tests/gpu_phase_trace.py 1024 6 > gpurun_out/r03_trace_pair256.log 2>&1
python tests/gpu_pair_sweep.py 1024 40 "RC_PH_PAIR=0;RC_PH_PAIR=2;RC_PH_PAIR=2,RC_PH_TW=128;RC_PH_PAIR=1" > gpurun_out/r03_sweep.log 2>&1
RC_PH_PAIR=2 ncu --set full --clock-control none --import-source on -k regex:pair256 -s 9 -c 2 -o gpurun_out/r03_pair256 \
    python tests/gpu_phase_time.py 2 1024 6 > gpurun_out/r03_ncu.log 2>&1
ls -la gpurun_out | grep r03_
