#!/bin/bash
mkdir -p gpurun_out
export QB_TX_SIZES=512,4096,8192,960,1920
for i in 1 2; do B200_TX_LEAF=0 python scripts/quick_bench.py tx; B200_TX_LEAF=1 python scripts/quick_bench.py tx; done > gpurun_out/q_tx.log 2>&1; cat gpurun_out/q_tx.log
