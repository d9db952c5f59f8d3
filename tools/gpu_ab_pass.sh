#!/usr/bin/env bash
# GPU-box helper: parity tests of one pass family on the in-tree library, then its timings for the in-tree library and for every build
# in variants/ named on the command line, then (NCU=regex) one full ncu capture of the kernels matching the regex.
#   usage: [NCU='surface_kernel'] bash tools/gpu_ab_pass.sh surface|shadow|frame [variant ...]     (variants/<variant>.so)
pass=${1:?surface|shadow|frame}; shift
case $pass in
  surface) tests="tests/test_surface_gpu.py tests/test_host_gpu.py"; perf=tools/perf_surface.py;;
  shadow)  tests="tests/test_shadow_gpu.py"; perf=tools/perf_shadow.py;;
  frame)   tests="tests/test_frame_gpu.py tests/test_host_gpu.py tests/test_zz_c_client_gpu.py"; perf=tools/perf_frame.py;;
  *) echo "unknown pass $pass"; exit 2;;
esac
mkdir -p gpurun_out
timeout 1200 python -m pytest $tests -q -m gpu > gpurun_out/ab_${pass}_tests.log 2>&1; echo "tests rc=$?"
grep -E "passed|failed|^FAILED|^ERROR|^E  " gpurun_out/ab_${pass}_tests.log | head -30 | cut -c1-300
{
echo "== $pass (in-tree)"; timeout 300 python $perf 2>&1 | grep -E '"ms|hbm_frac|rror|_4k"|_4096' | head -24
for v in "$@"; do echo "== $pass variant [$v]"; VQCUDA_LIB=variants/$v.so timeout 300 python $perf 2>&1 | grep -E '"ms|rror' | head -12; done
} > gpurun_out/ab_${pass}_perf.txt 2>&1
cat gpurun_out/ab_${pass}_perf.txt
if [ -n "$NCU" ]; then
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:"$NCU" -s 2 -c 1 -f -o gpurun_out/ab_${pass}_full python $perf > gpurun_out/ab_${pass}_ncu.log 2>&1; tail -1 gpurun_out/ab_${pass}_ncu.log
fi
