#!/bin/bash
# Run on the GPU box (through gpurun): ncu evidence for the round.  Outputs land in gpurun_out/.
R=${1:-r02}
mkdir -p gpurun_out
# 1. every launch of the default bench command with its device time (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${R}_launches.csv \
    python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err
# 2. full captures of the dominant kernels (one launch each)
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:"sws_vscale_rgb24_pair" -s 1 -c 1 -o gpurun_out/${R}_sws_pair python scripts/profile_target.py sws 32 > gpurun_out/p1.log 2>&1
$NCU -k regex:"sws_unscaled_kernel" -s 2 -c 1 -o gpurun_out/${R}_sws_lut python scripts/profile_target.py lut 32 > gpurun_out/p2.log 2>&1
$NCU -k regex:"sws_mma_plane" -s 0 -c 1 -o gpurun_out/${R}_mma_plane_y python scripts/profile_target.py scalep 8 > gpurun_out/p3.log 2>&1
$NCU -k regex:"sws_mma_rgb" -s 1 -c 1 -o gpurun_out/${R}_mma_rgb_final python scripts/profile_target.py scale 8 > gpurun_out/p4.log 2>&1
$NCU -k regex:"idct_mb420" -s 1 -c 1 -o gpurun_out/${R}_idct_put_final python scripts/profile_target.py idct 32 > gpurun_out/p5.log 2>&1
$NCU -k regex:"idct_mb420" -s 7 -c 1 -o gpurun_out/${R}_idct_add_final python scripts/profile_target.py idct 32 > gpurun_out/p6.log 2>&1
$NCU -k regex:"tx_r16_kernel" -s 1 -c 1 -o gpurun_out/${R}_tx_fft1024 python scripts/profile_target.py tx 8 > gpurun_out/p7.log 2>&1
$NCU -k regex:"tx_r16_kernel" -s 1 -c 1 -o gpurun_out/${R}_tx_fft2048_final python scripts/profile_target.py tx2048 8 > gpurun_out/p7b.log 2>&1
$NCU -k regex:"tx_r16_kernel" -s 4 -c 1 -o gpurun_out/${R}_tx_imdct2048_final python scripts/profile_target.py tx2048 8 > gpurun_out/p7c.log 2>&1
$NCU -k regex:"qpel_kernel" -s 1 -c 1 -o gpurun_out/${R}_qpel_ldg python scripts/profile_target.py qpel 16 > gpurun_out/p8.log 2>&1
$NCU -k regex:"esa" -s 1 -c 1 -o gpurun_out/${R}_esa python scripts/profile_target.py esa 8 > gpurun_out/p9.log 2>&1
$NCU -k regex:"h264_idct_kernel" -s 0 -c 1 -o gpurun_out/${R}_h264_idct4_final python scripts/profile_target.py h264 16 > gpurun_out/p10.log 2>&1
ls -la gpurun_out | tail -20
