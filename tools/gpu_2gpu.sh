#!/usr/bin/env bash
# GPU-box helper (N GPUs, default 2): bench.py under torchrun — the strong-scaled 8K forward step with the fused peer-store
# assembly + in-kernel rendezvous, and the single-launch specular prefilter strong scaling.
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_${N}gpu.json; tail -5 gpurun_out/bench_${N}gpu.err
