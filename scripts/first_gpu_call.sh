#!/bin/bash
# First GPU call of a round (through gpurun): run the tests whose kernels have only been executed under the host emulation as
# ordinary tests, one pytest process per test so that a fault in one cannot hide the others, and keep every log.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/first_gpu_call.sh'
# Outputs: gpurun_out/unverified/<test>.log, gpurun_out/unverified/summary.txt
set -u
mkdir -p gpurun_out/unverified
export B200_RUN_UNVERIFIED=1 B200_ISOLATED_CHILD=1
: > gpurun_out/unverified/summary.txt
# the verified tier first: if this fails nothing below means anything
timeout 600 python -m pytest tests -x -q -m gpu --deselect tests/test_zz_next_gpu.py -p no:cacheprovider > gpurun_out/unverified/verified_tier.log 2>&1
echo "verified tier: exit $?" >> gpurun_out/unverified/summary.txt
for t in $(python -m pytest tests/test_zz_next_gpu.py --collect-only -q -m gpu 2>/dev/null | grep '::'); do
    name=${t##*::}
    timeout 400 python -m pytest "$t" -m gpu -q -x -p no:cacheprovider > "gpurun_out/unverified/$name.log" 2>&1
    echo "$name: exit $?" >> gpurun_out/unverified/summary.txt
done
cat gpurun_out/unverified/summary.txt
