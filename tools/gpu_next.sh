#!/usr/bin/env bash
# GPU-box helper, FIRST job of the next round: the kernels written after this round's GPU budget ran out (vq_shadow.cu).
# 1. their parity tests (marker gpu_next; promote to gpu once green)  2. timings  3. memcheck  4. one ncu capture each
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu_next -x > gpurun_out/gpu_next_tests.log 2>&1; echo "gpu_next tests rc=$?" | tee -a gpurun_out/gpu_next_tests.log
tail -15 gpurun_out/gpu_next_tests.log
timeout 300 python tools/perf_shadow.py > gpurun_out/perf_shadow.json 2> gpurun_out/perf_shadow.err; echo "perf rc=$?"; cat gpurun_out/perf_shadow.json
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests -q -m gpu_next -k "33 or 5-3 or 65" > gpurun_out/sanitize_shadow.txt 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/sanitize_shadow.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'shadow_casters|depth_min_level' -s 4 -c 3 -f -o gpurun_out/shadow_full python tools/perf_shadow.py > gpurun_out/shadow_ncu.log 2>&1
ls -la gpurun_out
