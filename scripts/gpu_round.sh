#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02_final_tests_gpu.log 2>&1; tail -6 gpurun_out/r02_final_tests_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_final_smoke.log 2>&1; tail -1 gpurun_out/r02_final_smoke.log
