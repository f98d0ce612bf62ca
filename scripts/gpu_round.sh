#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "hbd or tx_fft_pfa or qpel" > gpurun_out/t_a.log 2>&1; tail -5 gpurun_out/t_a.log
