#!/usr/bin/env bash
mkdir -p gpurun_out
python -m pytest tests/test_forward_gpu.py tests/test_host_gpu.py -x -q -m gpu > gpurun_out/k1_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/k1_tests.log
tail -3 gpurun_out/k1_tests.log
python tools/perf_forward.py --check > gpurun_out/k1_default.log 2>&1; tail -4 gpurun_out/k1_default.log
: > gpurun_out/k1_variants.txt
for v in "$@"; do
  echo "== variant [$v]" >> gpurun_out/k1_variants.txt
  VQCUDA_LIB=variants/$v.so timeout 120 python tools/perf_forward.py 2>&1 | grep -E "forward 4K prepared|rror" >> gpurun_out/k1_variants.txt
done
cat gpurun_out/k1_variants.txt
python tools/perf_forward_decomp.py > gpurun_out/k1_decomp.txt 2>&1; cat gpurun_out/k1_decomp.txt
