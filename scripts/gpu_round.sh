#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "qpel or pel or host_entries or dropin" > gpurun_out/t_pel.log 2>&1; echo "pel tests rc=$?"; tail -4 gpurun_out/t_pel.log
for m in 0 1; do echo "QPEL_MMA=$m"; B200_QPEL_MMA=$m timeout 120 python scripts/quick_bench.py qpel 10; done
