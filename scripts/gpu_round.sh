#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "sws or rgb or scale or planar or range" > gpurun_out/t_a.log 2>&1; tail -8 gpurun_out/t_a.log
