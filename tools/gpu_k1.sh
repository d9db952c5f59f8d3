#!/usr/bin/env bash
# GPU-box helper: K1 A/B over the builds in variants/ (usage: bash tools/gpu_k1.sh v1 v2 ...), the forward parity tests
# (tests/test_forward_gpu.py and the full-size C3/C5 forward tests) on the in-tree library, then one full ncu capture of it.
mkdir -p gpurun_out
: > gpurun_out/k1_variants.txt
for v in "$@"; do
  echo "== variant [$v]" >> gpurun_out/k1_variants.txt
  VQCUDA_LIB=variants/$v.so timeout 120 python tools/perf_forward.py 2>&1 | grep -E "forward 4K prepared|rror" >> gpurun_out/k1_variants.txt
done
cat gpurun_out/k1_variants.txt
timeout 900 python -m pytest tests/test_forward_gpu.py tests/test_fullsize_gpu.py tests/test_shadow_gpu.py tests/test_host_gpu.py -q -m gpu -k "not c4 and not c2 and not specular" -s > gpurun_out/k1_tests.log 2>&1
grep -E "passed|failed|Error" gpurun_out/k1_tests.log | tail -15
timeout 600 ncu --set full --clock-control none --import-source on -k regex:forward_kernel -s 3 -c 2 -f -o gpurun_out/k1_full python tools/perf_forward.py > gpurun_out/k1_ncu.log 2>&1; tail -3 gpurun_out/k1_ncu.log
