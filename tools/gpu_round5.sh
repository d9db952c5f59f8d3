#!/usr/bin/env bash
mkdir -p gpurun_out
python tools/perf_forward.py --check > gpurun_out/k1_default.log 2>&1; tail -4 gpurun_out/k1_default.log
: > gpurun_out/k1_variants.txt
for v in "$@"; do
  echo "== variant [$v]" >> gpurun_out/k1_variants.txt
  VQCUDA_LIB=variants/$v.so timeout 120 python tools/perf_forward.py 2>&1 | grep -E "forward 4K prepared|rror" >> gpurun_out/k1_variants.txt
done
cat gpurun_out/k1_variants.txt
python tools/perf_forward_decomp.py > gpurun_out/k1_decomp.txt 2>&1; cat gpurun_out/k1_decomp.txt
