#!/usr/bin/env bash
# GPU-box helper: parity of the frame passes (.hdr decode walk/expand kernel, fused Mitchell downsize, skydome) and their timings.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_frame_gpu.py tests/test_host_gpu.py tests/test_zz_c_client_gpu.py -q -m gpu > gpurun_out/r2g_tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|^FAILED|^ERROR|^E  " gpurun_out/r2g_tests.log | head -30 | cut -c1-300
timeout 300 python tools/perf_frame.py > gpurun_out/r2g_perf.txt 2>&1; grep -E '"ms|hbm_frac|e2e|rror|_4k|_4096|resize' gpurun_out/r2g_perf.txt | head -40

timeout 400 ncu --set full --clock-control none --import-source on -k regex:"hdr_decode_rle" -s 4 -c 2 -f -o gpurun_out/frame_full python tools/perf_frame.py > gpurun_out/frame_ncu.log 2>&1; tail -1 gpurun_out/frame_ncu.log
