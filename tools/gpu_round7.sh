#!/usr/bin/env bash
mkdir -p gpurun_out
python -m pytest tests/test_frame_gpu.py tests/test_ibl_gpu.py tests/test_host_gpu.py -q -m gpu > gpurun_out/r7_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r7_tests.log
tail -15 gpurun_out/r7_tests.log
python tools/perf_frame.py > gpurun_out/r7_frame.json 2>&1; cat gpurun_out/r7_frame.json | head -40
: > gpurun_out/k1_variants.txt
for v in "$@"; do
  echo "== variant [$v]" >> gpurun_out/k1_variants.txt
  VQCUDA_LIB=variants/$v.so timeout 120 python tools/perf_forward.py --check 2>&1 | grep -E "forward 4K prepared|rror|parity" >> gpurun_out/k1_variants.txt
done
cat gpurun_out/k1_variants.txt
