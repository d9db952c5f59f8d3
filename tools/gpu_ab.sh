#!/usr/bin/env bash
# GPU-box helper: A/B of the builds in variants/ — K1 at 4K (tools/perf_forward.py) and the 2x EASU (tools/perf_post.py easu) —
# then the parity tests that cover what changed, on the in-tree library.   usage: bash tools/gpu_ab.sh name1 name2 ...
mkdir -p gpurun_out
: > gpurun_out/ab_variants.txt
for v in "$@"; do
  echo "== variant [$v]" >> gpurun_out/ab_variants.txt
  VQCUDA_LIB=variants/$v.so timeout 120 python tools/perf_forward.py 2>&1 | grep -E "forward 4K prepared|rror" >> gpurun_out/ab_variants.txt
  VQCUDA_LIB=variants/$v.so timeout 120 python tools/perf_post.py easu 2>&1 | grep -E "easu|rror" >> gpurun_out/ab_variants.txt
done
cat gpurun_out/ab_variants.txt
timeout 1200 python -m pytest tests/test_forward_gpu.py tests/test_fullsize_gpu.py tests/test_post_gpu.py tests/test_ibl_gpu.py tests/test_shadow_gpu.py tests/test_host_gpu.py -q -m gpu > gpurun_out/ab_tests.log 2>&1
grep -E "passed|failed|^E  " gpurun_out/ab_tests.log | tail -12 | cut -c1-400
