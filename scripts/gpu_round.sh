#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "hbd or chroma or edge or pel" > gpurun_out/t_a.log 2>&1; tail -8 gpurun_out/t_a.log
