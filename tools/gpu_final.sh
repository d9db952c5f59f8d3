#!/usr/bin/env bash
# GPU-box helper: what the driver runs at round end — the -m gpu suite, smoke(), the bench line
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/full_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/full_tests.log
tail -5 gpurun_out/full_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_1gpu.json
