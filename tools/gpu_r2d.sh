#!/usr/bin/env bash
# GPU-box helper: parity of the surface producer (minority queue) and the frame passes (skydome), A/B timings over variants/, ncu of the surface kernel.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_surface_gpu.py tests/test_frame_gpu.py tests/test_host_gpu.py -q -m gpu -x > gpurun_out/r2d_tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|^FAILED|^ERROR|^E  " gpurun_out/r2d_tests.log | head -30 | cut -c1-300
{
echo "== surface (in-tree)"; timeout 200 python tools/perf_surface.py 2>&1 | grep -E '"ms|rror|hbm_frac' | head -4
for v in "$@"; do echo "== surface variant [$v]"; VQCUDA_LIB=variants/$v.so timeout 200 python tools/perf_surface.py 2>&1 | grep -E '"ms|rror' | head -3; done
echo "== frame passes"; timeout 300 python tools/perf_frame.py 2>&1 | grep -E 'skydome|"ms|rror' | head -20
} > gpurun_out/r2d_perf.txt 2>&1
cat gpurun_out/r2d_perf.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:surface_ -s 4 -c 2 -f -o gpurun_out/surf_full3 python tools/perf_surface.py > gpurun_out/surf_ncu.log 2>&1; tail -1 gpurun_out/surf_ncu.log
