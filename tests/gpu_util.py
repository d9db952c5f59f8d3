"""Helpers shared by the -m gpu parity tests."""
import numpy as np
import torch

TOL = 1e-4   # BASELINE.json north_star: |delta| <= 1e-4 per channel (fp32)


def dev(a: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def host(t: torch.Tensor) -> np.ndarray:
    torch.cuda.synchronize()
    return t.detach().cpu().numpy()


def report(name, got, ref):
    d = np.abs(got.astype(np.float64) - ref.astype(np.float64))
    scale = np.maximum(1.0, np.abs(ref.astype(np.float64)))
    return {"name": name, "max_abs": float(d.max()), "max_scaled": float((d / scale).max()),
            "frac_abs_le_tol": float((d <= TOL).mean()), "ref_max": float(np.abs(ref).max())}


def assert_abs(name, got, ref, tol=TOL):
    """strict: |delta| <= tol per channel"""
    r = report(name, got, ref)
    assert np.isfinite(got).all(), f"{name}: non-finite output"
    assert r["max_abs"] <= tol, f"{name}: max |delta| = {r['max_abs']:.3e} > {tol} ({r})"
    return r


def assert_scaled(name, got, ref, tol=TOL, min_frac_strict=0.99):
    """HDR outputs: |delta| <= tol * max(1, |ref|)  (fp32 cannot hold an absolute 1e-4 above ~1e3: 1 ulp at 1024 is 1.2e-4),
    AND, explicitly, the north star's strict absolute bound |delta| <= tol on every value with |ref| <= 1, AND at least
    `min_frac_strict` of ALL values inside the strict absolute bound (reported as frac_abs_le_tol)."""
    r = report(name, got, ref)
    assert np.isfinite(got).all(), f"{name}: non-finite output"
    assert r["max_scaled"] <= tol, f"{name}: max scaled |delta| = {r['max_scaled']:.3e} > {tol} ({r})"
    d = np.abs(got.astype(np.float64) - ref.astype(np.float64))
    ldr = np.abs(ref) <= 1.0
    r["max_abs_where_ref_le_1"] = float(d[ldr].max()) if ldr.any() else 0.0
    r["frac_ref_le_1"] = float(ldr.mean())
    assert r["max_abs_where_ref_le_1"] <= tol, f"{name}: strict |delta| = {r['max_abs_where_ref_le_1']:.3e} > {tol} where |ref| <= 1 ({r})"
    assert r["frac_abs_le_tol"] >= min_frac_strict, f"{name}: only {r['frac_abs_le_tol']:.6f} of the values within the strict {tol} ({r})"
    return r
