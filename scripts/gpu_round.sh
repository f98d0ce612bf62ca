#!/bin/bash
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_final_bench_n1.json 2> gpurun_out/bench.err; tail -c 200 gpurun_out/r02_final_bench_n1.json; tail -3 gpurun_out/bench.err
