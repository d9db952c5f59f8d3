"""GPU-box helper: time the §8(f).1 surface producer + RGBA8 mip chain at 4K with warm clocks."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import vqengine_b200 as vq
import bench
ctx = vq.Context(0)
peak, _ = bench.hbm_peak()
print(json.dumps(bench.surface_producer_pass(ctx, vq, torch, peak), indent=1))
