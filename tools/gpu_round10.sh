#!/usr/bin/env bash
mkdir -p gpurun_out
python -m pytest tests/test_frame_gpu.py -q -m gpu > gpurun_out/r10_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r10_tests.log
tail -4 gpurun_out/r10_tests.log
python tools/perf_frame.py > gpurun_out/r10_frame.json 2>&1; grep -A4 "image_resize\|hdr_decode" gpurun_out/r10_frame.json | head -30
