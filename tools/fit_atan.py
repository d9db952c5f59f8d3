"""Fits the degree-8 odd minimax polynomial used by atan01() in vqengine_b200/csrc/vq_ibl.cu and checks it in fp32."""
import numpy as np
x = np.unique(np.clip(np.cos(np.linspace(0, np.pi, 4001)) * 0.5 + 0.5, 1e-9, 1)); z = x * x
A = np.stack([x * z ** k for k in range(9)], axis=1); y = np.arctan(x); w = np.ones_like(x)
for _ in range(200):
    c, *_ = np.linalg.lstsq(A * w[:, None], y * w, rcond=None); e = np.abs(A @ c - y); w = w * (1 + 4 * e / e.max()); w /= w.mean()
q = np.linspace(0, 1, 2000001, dtype=np.float32); zz = (q * q).astype(np.float32); p = np.float32(c[-1]) * np.ones_like(zz)
for k in range(7, -1, -1): p = (p * zz + np.float32(c[k])).astype(np.float32)
print("coeffs (low to high):", [float(np.float32(v)) for v in c])
print("max |err| in fp32:", np.abs((p * q).astype(np.float64) - np.arctan(q.astype(np.float64))).max())
