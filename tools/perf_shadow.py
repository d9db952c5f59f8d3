"""GPU-box helper: time the §8(f).4 passes (shadowed forward pass = PCF kernel + K1, MIN depth pyramid) at 4K with warm clocks."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import vqengine_b200 as vq
import bench
ctx = vq.Context(0)
peak, _ = bench.hbm_peak()
envk = bench.build_env_maps_gpu(ctx, vq, torch)
print(json.dumps(bench.shadow_passes(ctx, vq, torch, envk, peak), indent=1))
