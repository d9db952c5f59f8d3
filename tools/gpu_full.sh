#!/usr/bin/env bash
# GPU-box helper: the whole -m gpu suite, the official bench line, the ncu evidence kept under profiles/, memcheck
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/full_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/full_tests.log
tail -6 gpurun_out/full_tests.log
python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench rc=$?"; tail -c 1200 gpurun_out/bench_1gpu.json
ncu --set full --clock-control none --import-source on -k regex:forward_kernel -s 8 -c 1 -f -o gpurun_out/k1_full python tools/perf_forward.py > gpurun_out/k1_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'hdr_|skydome|apply_refl|resize_' -s 6 -c 7 -f -o gpurun_out/frame_full python tools/run_pass.py frame 2 > gpurun_out/frame_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
timeout 240 compute-sanitizer --tool memcheck python tools/sanitize_small.py > gpurun_out/sanitize_memcheck.txt 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/sanitize_memcheck.txt; tail -4 gpurun_out/sanitize_memcheck.txt
./vqengine_b200/host/vq_headless_test -Test -TestFrames=100 > gpurun_out/headless_100frames.txt 2>&1; echo "headless rc=$?" | tee -a gpurun_out/headless_100frames.txt; tail -3 gpurun_out/headless_100frames.txt
ls -la gpurun_out
