#!/usr/bin/env bash
# GPU-box helper: parity of the surface producer / shadowed pass on the in-tree library, then A/B timings over variants/
# (surface: args starting with surf_, PCF: args starting with pcf_), and one ncu capture of each of the two kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_surface_gpu.py tests/test_shadow_gpu.py tests/test_host_gpu.py -q -m gpu > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|^FAILED|^ERROR|^E  " gpurun_out/r2c_tests.log | head -30 | cut -c1-300
{
echo "== surface (in-tree)"; timeout 200 python tools/perf_surface.py 2>&1 | grep -E '"ms|rror|hbm_frac' | head -3
echo "== shadow (in-tree)"; timeout 200 python tools/perf_shadow.py 2>&1 | grep -E '"ms|rror' | head -2
for v in "$@"; do
  case $v in
    surf_*) echo "== surface variant [$v]"; VQCUDA_LIB=variants/$v.so timeout 200 python tools/perf_surface.py 2>&1 | grep -E '"ms|rror' | head -2;;
    pcf_*) echo "== shadow variant [$v]"; VQCUDA_LIB=variants/$v.so timeout 200 python tools/perf_shadow.py 2>&1 | grep -E '"ms|rror' | head -2;;
  esac
done
} > gpurun_out/r2c_perf.txt 2>&1
cat gpurun_out/r2c_perf.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:surface_kernel -s 2 -c 1 -f -o gpurun_out/surf_full2 python tools/perf_surface.py > gpurun_out/surf_ncu.log 2>&1; tail -1 gpurun_out/surf_ncu.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:shadow_pcf_kernel -s 2 -c 1 -f -o gpurun_out/pcf_full2 python tools/perf_shadow.py > gpurun_out/pcf_ncu.log 2>&1; tail -1 gpurun_out/pcf_ncu.log
