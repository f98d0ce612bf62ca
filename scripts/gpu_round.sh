#!/bin/bash
mkdir -p gpurun_out
for v in 7 8 9 10; do
  echo "== var $v"
  B200_CHROMA_VAR=$v python -m pytest tests -m gpu -x -q -k "chroma" 2>&1 | tail -1
  B200_CHROMA_VAR=$v python scripts/quick_bench.py chroma 2>&1 | tail -1
  B200_CHROMA_VAR=$v python scripts/quick_bench.py chroma 2>&1 | tail -1
done > gpurun_out/chroma_vars.log 2>&1
cat gpurun_out/chroma_vars.log
