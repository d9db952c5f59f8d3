#!/usr/bin/env bash
# GPU-box helper: the whole -m gpu suite, the official bench lines, and the ncu evidence kept under profiles/
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/full_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/full_tests.log
tail -6 gpurun_out/full_tests.log
python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_1gpu.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; tail -c 600 gpurun_out/bench_ref.json
ncu --set full --clock-control none --import-source on -k regex:forward_kernel -s 8 -c 1 -f -o gpurun_out/k1_full python tools/perf_forward.py > gpurun_out/k1_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'hdr_|skydome|apply_refl' -s 4 -c 5 -f -o gpurun_out/frame_full python tools/run_pass.py frame 2 > gpurun_out/frame_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
ls -la gpurun_out
