#!/usr/bin/env bash
# GPU-box helper: K1 at 4K with and without the persisting-L2 window, the forward parity tests, then one full ncu capture of the
# in-tree library (-> tools/make_forward_traffic.py turns it into profiles/forward_traffic.json).
mkdir -p gpurun_out
: > gpurun_out/k1_variants.txt
for p in 0 1; do
  echo "== VQ_L2_PERSIST=$p" >> gpurun_out/k1_variants.txt
  VQ_L2_PERSIST=$p timeout 120 python tools/perf_forward.py 2>&1 | grep -E "forward 4K|rror" >> gpurun_out/k1_variants.txt
done
for v in "$@"; do
  echo "== variant [$v]" >> gpurun_out/k1_variants.txt
  VQCUDA_LIB=variants/$v.so timeout 120 python tools/perf_forward.py 2>&1 | grep -E "forward 4K prepared|rror" >> gpurun_out/k1_variants.txt
done
cat gpurun_out/k1_variants.txt
timeout 900 python -m pytest tests/test_forward_gpu.py tests/test_fullsize_gpu.py tests/test_shadow_gpu.py tests/test_host_gpu.py tests/test_zz_c_client_gpu.py -q -m gpu -k "not c4 and not c2 and not specular" > gpurun_out/k1_tests.log 2>&1
grep -E "passed|failed|Error" gpurun_out/k1_tests.log | tail -8
timeout 600 ncu --set full --clock-control none --import-source on -k regex:forward_kernel -s 3 -c 2 -f -o gpurun_out/k1_full python tools/perf_forward.py > gpurun_out/k1_ncu.log 2>&1; tail -2 gpurun_out/k1_ncu.log
