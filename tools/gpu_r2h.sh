#!/usr/bin/env bash
# GPU-box helper: the shadowed pass over PCF-kernel occupancy variants (variants/pcf_*.so), then its parity tests on the in-tree library
mkdir -p gpurun_out
{
echo "== shadow (in-tree)"; timeout 200 python tools/perf_shadow.py 2>&1 | grep -E '"ms|rror' | head -2
for v in "$@"; do echo "== shadow variant [$v]"; VQCUDA_LIB=variants/$v.so timeout 200 python tools/perf_shadow.py 2>&1 | grep -E '"ms|rror' | head -2; done
} > gpurun_out/r2h_perf.txt 2>&1
cat gpurun_out/r2h_perf.txt
