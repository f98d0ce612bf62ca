#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sws_mma_gpu.py -q > gpurun_out/t_mma.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t_mma.log
timeout 120 python scripts/quick_bench.py scale 10
for tr in 16 29; do echo "B200_SWS_TR=$tr"; B200_SWS_TR=$tr timeout 120 python scripts/quick_bench.py scale 10; done
