#!/usr/bin/env bash
# GPU-box helper: A/B timings of K1 builds kept in variants/<name>.so (built here with
#   VQ_OUT=variants/<name>.so bash vqengine_b200/csrc/build.sh -DFWD_CTAS_PER_SM=5 ...), 4K frame; parity of each build through the forward tests
# usage: bash tools/gpu_variants.sh name1 name2 ...      -> gpurun_out/k1_variants.txt
mkdir -p gpurun_out
: > gpurun_out/k1_variants.txt
for v in "$@"; do
  echo "== variant [$v]" >> gpurun_out/k1_variants.txt
  VQCUDA_LIB=variants/$v.so timeout 120 python tools/perf_forward.py 2>&1 | grep -E "forward 4K prepared|rror" >> gpurun_out/k1_variants.txt
  VQCUDA_LIB=variants/$v.so timeout 300 python -m pytest tests/test_forward_gpu.py -q -m gpu -k "full_size or config3" 2>&1 | tail -1 >> gpurun_out/k1_variants.txt
done
cat gpurun_out/k1_variants.txt
python tools/perf_forward_decomp.py > gpurun_out/k1_decomp.txt 2>&1; cat gpurun_out/k1_decomp.txt
