#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "tx" > gpurun_out/t_tx.log 2>&1; tail -3 gpurun_out/t_tx.log
python scripts/quick_bench.py tx > gpurun_out/q_tx.log 2>&1; cat gpurun_out/q_tx.log
