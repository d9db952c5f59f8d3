#!/usr/bin/env bash
# GPU-box helper (round 2, second session): the whole -m gpu suite, then timings of what changed (surface producer with texel
# records, shadowed pass = PCF records + K1, depth pyramid, K1 itself) and one ncu capture each of the two new kernels.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --durations=5 > gpurun_out/gpu_tests.log 2>&1; echo "gpu tests rc=$?"; grep -E "passed|failed|^FAILED|^ERROR|^E  " gpurun_out/gpu_tests.log | head -30 | cut -c1-300
{
echo "== surface (in-tree)"; timeout 200 python tools/perf_surface.py 2>&1 | grep -E '"ms|rror|hbm_frac'
for v in "$@"; do echo "== surface variant [$v]"; VQCUDA_LIB=variants/$v.so timeout 200 python tools/perf_surface.py 2>&1 | grep -E '"ms|rror'; done
echo "== shadow"; timeout 200 python tools/perf_shadow.py 2>&1 | grep -E '"ms|rror|hbm_frac|shadow'
echo "== forward"; timeout 200 python tools/perf_forward.py 2>&1 | grep -E "forward 4K|rror"
} > gpurun_out/r2b_perf.txt 2>&1
cat gpurun_out/r2b_perf.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:surface_kernel -s 2 -c 1 -f -o gpurun_out/surf_full python tools/perf_surface.py > gpurun_out/surf_ncu.log 2>&1; tail -1 gpurun_out/surf_ncu.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:shadow_pcf_kernel -s 2 -c 1 -f -o gpurun_out/pcf_full python tools/perf_shadow.py > gpurun_out/pcf_ncu.log 2>&1; tail -1 gpurun_out/pcf_ncu.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:forward_kernel -s 3 -c 2 -f -o gpurun_out/k1_full python tools/perf_forward.py > gpurun_out/k1_ncu.log 2>&1; tail -1 gpurun_out/k1_ncu.log
