"""ctypes/numpy front end of the scalar oracle (oracle/liboracle.so) and of the reference-built
FidelityFX setup functions (oracle/_ref/libffxref.so). TEST INFRASTRUCTURE ONLY: imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs — never by the product.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "liboracle.so")
REF_LIB = os.path.join(ORACLE_DIR, "_ref", "libffxref.so")

f32 = C.c_float
u32 = C.c_uint32
_fp = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")


def build(force: bool = False) -> None:
    if force or not os.path.exists(LIB) or any(
            os.path.getmtime(os.path.join(ORACLE_DIR, f)) > os.path.getmtime(LIB)
            for f in os.listdir(ORACLE_DIR) if f.endswith((".cpp", ".h"))):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s", "all"])


_lib = None
_ref = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.orc_ndf_ggx.restype = f32
        _lib.orc_geometry_smith.restype = f32
        _lib.orc_spotlight_intensity.restype = f32
        _lib.orc_aprx.restype = f32
        _lib.orc_cubemap_texel_count.restype = C.c_uint64
        _lib.orc_pyramid_texel_count.restype = C.c_uint64
    return _lib


def ref():
    """The reference's own A_CPU FidelityFX setup functions, or None if oracle/_ref was not built."""
    global _ref
    if _ref is None and os.path.exists(REF_LIB):
        _ref = C.CDLL(REF_LIB)
    return _ref


def _f(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(f32))


def cpu_threads() -> int:
    return max(1, os.cpu_count() or 1)


# ---- whole passes ------------------------------------------------------------------------------
def forward_lighting(pf, pv, planes, diff_cube, diff_res, spec_cube, spec_res, spec_mips, lut,
                     row_begin=0, row_end=None, threads=None) -> np.ndarray:
    pos, nrm, alb = (_f(p) for p in planes[:3])
    em = _f(planes[3]) if len(planes) > 3 else None
    h, w = pos.shape[:2]
    out = np.zeros((h, w, 4), np.float32)
    lutc = _f(lut)
    lib().orc_forward_lighting(C.byref(pf), C.byref(pv), _p(pos), _p(nrm), _p(alb),
                               _p(em) if em is not None else None, w, h,
                               _p(_f(diff_cube)), diff_res, _p(_f(spec_cube)), spec_res, spec_mips,
                               _p(lutc), lutc.shape[1], lutc.shape[0], _p(out),
                               row_begin, h if row_end is None else row_end, threads or cpu_threads())
    return out


def hdri_build_mips(level0: np.ndarray, levels: int) -> np.ndarray:
    h, w = level0.shape[:2]
    n = int(lib().orc_pyramid_texel_count(w, h, levels))
    pyr = np.zeros((n, 4), np.float32)
    pyr[: w * h] = _f(level0).reshape(-1, 4)
    lib().orc_hdri_build_mips(_p(pyr), w, h, levels)
    return pyr


def diffuse_irradiance(pyr, w, h, levels, res, step=0.0, n_phi=64, n_theta=16, src_mip=3,
                       row_begin=0, row_end=None, threads=None, f64_accum=False) -> np.ndarray:
    out = np.zeros((6 * res * res, 4), np.float32)
    lib().orc_diffuse_irradiance(_p(_f(pyr)), w, h, levels, f32(step), n_phi, n_theta, src_mip, _p(out), res,
                                 row_begin, 6 * res if row_end is None else row_end, threads or cpu_threads(), int(f64_accum))
    return out


def specular_prefilter(pyr, w, h, levels, res, mips, num_samples=512, row_begin=0, row_end=None, threads=None):
    n = int(lib().orc_cubemap_texel_count(res, mips))
    out = np.zeros((n, 4), np.float32)
    total = lib().orc_cubemap_row_count(res, mips)
    lib().orc_specular_prefilter(_p(_f(pyr)), w, h, levels, _p(out), res, mips, num_samples,
                                 row_begin, total if row_end is None else row_end, threads or cpu_threads())
    return out


def brdf_integration_lut(w, h, samples=2048, row_begin=0, row_end=None, threads=None) -> np.ndarray:
    out = np.zeros((h, w, 2), np.float32)
    lib().orc_brdf_integration_lut(_p(out), w, h, samples, row_begin, h if row_end is None else row_end,
                                   threads or cpu_threads())
    return out


def gaussian_blur(img, vertical: bool, threads=None) -> np.ndarray:
    img = _f(img)
    out = np.zeros_like(img)
    lib().orc_gaussian_blur(_p(img), _p(out), img.shape[1], img.shape[0], int(vertical), threads or cpu_threads())
    return out


def tonemap(params, img, threads=None) -> np.ndarray:
    img = _f(img)
    out = np.zeros_like(img)
    lib().orc_tonemap(C.byref(params), _p(img), _p(out), img.shape[1], img.shape[0], threads or cpu_threads())
    return out


def cas(con, img, threads=None) -> np.ndarray:
    img = _f(img)
    out = np.zeros_like(img)
    lib().orc_cas(con, _p(img), _p(out), img.shape[1], img.shape[0], threads or cpu_threads())
    return out


def fsr_easu(con, img, out_w, out_h, address_mode=0, threads=None) -> np.ndarray:
    img = _f(img)
    out = np.zeros((out_h, out_w, 4), np.float32)
    lib().orc_fsr_easu(con, address_mode, _p(img), img.shape[1], img.shape[0], _p(out), out_w, out_h,
                       threads or cpu_threads())
    return out


def fsr_rcas(con, img, threads=None) -> np.ndarray:
    img = _f(img)
    out = np.zeros_like(img)
    lib().orc_fsr_rcas(con, _p(img), _p(out), img.shape[1], img.shape[0], threads or cpu_threads())
    return out


def spd_downsample(img, mips: int):
    img = _f(img)
    h, w = img.shape[:2]
    sizes = [(w >> l, h >> l) for l in range(1, mips + 1) if (w >> l) >= 1 and (h >> l) >= 1]
    packed = np.zeros((sum(a * b for a, b in sizes), 4), np.float32)
    lib().orc_spd_downsample(_p(img), w, h, mips, _p(packed))
    out, o = [], 0
    for (lw, lh) in sizes:
        out.append(packed[o:o + lw * lh].reshape(lh, lw, 4))
        o += lw * lh
    return out


# ---- setup functions ---------------------------------------------------------------------------
def cas_setup(sharp, iw, ih, ow, oh, which="oracle"):
    con = (u32 * 8)()
    fn = lib().orc_cas_setup if which == "oracle" else ref().ref_cas_setup
    fn(con, f32(sharp), f32(iw), f32(ih), f32(ow), f32(oh))
    return con


def fsr_easu_con(vw, vh, iw, ih, ow, oh, which="oracle"):
    con = (u32 * 16)()
    fn = lib().orc_fsr_easu_con if which == "oracle" else ref().ref_fsr_easu_con
    fn(con, f32(vw), f32(vh), f32(iw), f32(ih), f32(ow), f32(oh))
    return con


def fsr_rcas_con(stops, which="oracle"):
    con = (u32 * 4)()
    fn = lib().orc_fsr_rcas_con if which == "oracle" else ref().ref_fsr_rcas_con
    fn(con, f32(stops))
    return con


def spd_setup(rect, mips=-1, which="oracle"):
    d, o, n = (u32 * 2)(), (u32 * 2)(), (u32 * 2)()
    r = (u32 * 4)(*rect)
    fn = lib().orc_spd_setup if which == "oracle" else ref().ref_spd_setup
    fn(d, o, n, r, mips)
    return list(d), list(o), list(n)
