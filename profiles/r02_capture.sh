#!/bin/bash
# Round-2 profile capture (run on the GPU box through gpurun from the repo root; outputs under gpurun_out/, summaries are
# extracted in the build container by profiles/r02_extract.py and committed under profiles/):
#   1. ncu launch list (device time of every launch) of a short default bench pass, B = 1024      -> r02_launches_b1024.csv
#   2. the same for one 128-sequence shard (persistent sequence kernel)                              -> r02_launches_b128.csv
#   3. ncu --set full: persistent sequence kernel (B = 128), single-stream TMA-staged kernel, device-resident L-BFGS,
#      per-frame row kernels (prep / lists / split / kin) of the multi-launch path                    -> r02_*.ncu-rep
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_b1024.csv \
    python tests/gpu_phase_time.py 2 1024 8 > gpurun_out/r02_l1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_b128.csv \
    python tests/gpu_phase_time.py 3 128 64 > gpurun_out/r02_l2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rc_seq_kernel -c 1 -o gpurun_out/r02_seq \
    python tests/gpu_phase_time.py 3 128 40 > gpurun_out/r02_p1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rc_stream2_kernel -s 30 -c 2 -o gpurun_out/r02_stream2 \
    python tests/gpu_phase_time.py 2 1 64 high > gpurun_out/r02_p2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rc_smplify_lbfgs -c 1 -o gpurun_out/r02_smplify \
    python tests/gpu_smplify_profile.py > gpurun_out/r02_p3.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:rc_kin_warp|rc_prep_warp|rc_split_multi|rc_lists' -s 12 -c 6 -o gpurun_out/r02_rows \
    python tests/gpu_phase_time.py 2 1024 6 > gpurun_out/r02_p4.log 2>&1
ls -la gpurun_out | grep r02_
