#!/usr/bin/env bash
# GPU-box helper: decomposition timings + one ncu --set full capture of K1 (source-level)
mkdir -p gpurun_out
python tools/perf_forward_decomp.py > gpurun_out/k1_decomp.txt 2>&1; cat gpurun_out/k1_decomp.txt
ncu --set full --clock-control none --import-source on -k regex:forward_kernel -s 8 -c 1 -f -o gpurun_out/k1_full \
    python tools/perf_forward.py > gpurun_out/k1_ncu.log 2>&1; tail -3 gpurun_out/k1_ncu.log
ls -la gpurun_out
