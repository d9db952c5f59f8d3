#!/usr/bin/env bash
# GPU-box helper: the warp-specialised K1 (VQ_FWD_WS=1) against the default kernel: timing, parity tests, decomposition.
mkdir -p gpurun_out
: > gpurun_out/ws.txt
echo "== default" >> gpurun_out/ws.txt; timeout 120 python tools/perf_forward.py 2>&1 | grep -E "forward 4K|rror" >> gpurun_out/ws.txt
echo "== VQ_FWD_WS=1" >> gpurun_out/ws.txt; VQ_FWD_WS=1 timeout 120 python tools/perf_forward.py 2>&1 | grep -E "forward 4K|rror|rap" >> gpurun_out/ws.txt
for v in "$@"; do echo "== VQ_FWD_WS=1 variant [$v]" >> gpurun_out/ws.txt; VQ_FWD_WS=1 VQCUDA_LIB=variants/$v.so timeout 120 python tools/perf_forward.py 2>&1 | grep -E "forward 4K prepared|rror|rap" >> gpurun_out/ws.txt; done
cat gpurun_out/ws.txt
VQ_FWD_WS=1 timeout 900 python -m pytest tests/test_forward_gpu.py tests/test_fullsize_gpu.py tests/test_shadow_gpu.py tests/test_host_gpu.py tests/test_zz_c_client_gpu.py -q -m gpu -k "not c4 and not c2 and not specular" -x > gpurun_out/ws_tests.log 2>&1
grep -E "passed|failed|Error" gpurun_out/ws_tests.log | tail -8
VQ_FWD_WS=1 timeout 200 python tools/perf_forward_decomp.py 2>&1 | grep "us$" | tee gpurun_out/ws_decomp.txt
