#!/bin/bash
# Run on the GPU box (through gpurun): ncu evidence for the round.  Outputs land in gpurun_out/.
set -x
mkdir -p gpurun_out
# 1. every launch of the default bench command with its device time (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err
# 2. full captures of the dominant kernels (one launch each)
ncu --set full --clock-control none --import-source on -k regex:"sws_vscale_rgb24_fast" -s 2 -c 1 -o gpurun_out/r01_sws_fate python scripts/profile_target.py sws 32 > gpurun_out/p1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"sws_unscaled_kernel" -s 2 -c 1 -o gpurun_out/r01_sws_lut python scripts/profile_target.py lut 32 > gpurun_out/p2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"idct_mb420" -s 2 -c 1 -o gpurun_out/r01_idct_put python scripts/profile_target.py idct 32 > gpurun_out/p3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"idct_mb420" -s 7 -c 1 -o gpurun_out/r01_idct_add python scripts/profile_target.py idct 32 > gpurun_out/p4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"tx_fft" -s 1 -c 1 -o gpurun_out/r01_tx_fft python scripts/profile_target.py tx 8 > gpurun_out/p5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"qpel_" -s 1 -c 1 -o gpurun_out/r01_qpel python scripts/profile_target.py qpel 8 > gpurun_out/p6.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"chroma_" -s 1 -c 1 -o gpurun_out/r01_chroma python scripts/profile_target.py chroma 8 > gpurun_out/p8.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"esa_" -s 1 -c 1 -o gpurun_out/r01_esa python scripts/profile_target.py esa 8 > gpurun_out/p7.log 2>&1
ls -la gpurun_out
