#!/bin/bash
mkdir -p gpurun_out
timeout 1300 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 0 python -m pytest tests -m gpu -q -x > gpurun_out/sanitizer_all.log 2>&1
echo "exit $?"; grep -c "Invalid\|out of bounds" gpurun_out/sanitizer_all.log; tail -8 gpurun_out/sanitizer_all.log
