#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "h264 or idct" > gpurun_out/t_h264.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t_h264.log
timeout 120 python scripts/quick_bench.py h264 10
