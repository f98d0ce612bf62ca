#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "round2 or fused or mecmp or satd or unquant" > gpurun_out/t_new.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/t_new.log
