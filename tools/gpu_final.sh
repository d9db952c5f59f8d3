#!/usr/bin/env bash
# GPU-box helper: what the driver runs at round end (the -m gpu suite, smoke(), the 1-GPU bench line), the sanitizer passes over one
# small launch of every kernel and the ncu launch list of the bench command.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --durations=5 > gpurun_out/full_tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|^FAILED|^ERROR|^E  " gpurun_out/full_tests.log | head -20 | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench rc=$?"; head -c 400 gpurun_out/bench_1gpu.json; echo; tail -2 gpurun_out/bench_1gpu.err
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_small.py > gpurun_out/sanitize_memcheck.txt 2>&1; tail -2 gpurun_out/sanitize_memcheck.txt
timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_small.py > gpurun_out/sanitize_racecheck.txt 2>&1; tail -2 gpurun_out/sanitize_racecheck.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1; echo "launch list rc=$?"
