#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"chroma_kernel_p" -s 1 -c 1 -o gpurun_out/r02_chroma_p python scripts/profile_target.py chroma 32 > gpurun_out/pc.log 2>&1
ls -la gpurun_out/r02_chroma_p.ncu-rep; tail -2 gpurun_out/pc.log
