"""GPU-box helper: timings of the SURVEY 8(f).2/(f).3 kernels (same code as bench.py's frame_format_passes)."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import vqengine_b200 as vq
import bench
ctx = vq.Context(0)
envk = bench.build_env_maps_gpu(ctx, vq, torch)
peak, _ = bench.hbm_peak() if hasattr(bench, "hbm_peak") else (6588.0, "")
print(json.dumps(bench.frame_format_passes(ctx, vq, torch, envk, peak), indent=1))
