"""Seeded synthetic inputs for the shading passes; numpy only, no GPU.

The reference ships no HDRI assets and its forward pass consumes rasterised attributes, so every
input is synthetic: the same bytes are fed to the CUDA kernels and to the oracle.
"""
from __future__ import annotations

import ctypes as C
import numpy as np

from . import (PerFrameData, PerViewLightingData, TonemapperParams, COLOR_SPACE_REC_709, DISPLAY_CURVE_SRGB)

SEED_BASE = 0x5EED0000


def hdri(width: int, height: int, seed: int = SEED_BASE + 2, peak: float = 16.0) -> np.ndarray:
    """Equirect RGBA32F [H,W,4]: sky gradient 0.05..2.0 + 8 Gaussian 'sun' lobes (peak <= `peak`) + 5 % noise."""
    rng = np.random.default_rng(seed)
    v = (np.arange(height, dtype=np.float32) + 0.5) / height
    u = (np.arange(width, dtype=np.float32) + 0.5) / width
    U, V = np.meshgrid(u, v)
    sky = 0.05 + 1.95 * (1.0 - V) ** 2
    img = np.stack([sky * 0.6, sky * 0.8, sky * 1.0], axis=-1).astype(np.float32)
    img += (0.25 * (1.0 + np.sin(2 * np.pi * 3 * U))[..., None] * np.array([0.9, 0.7, 0.4], np.float32)) * V[..., None]
    for _ in range(8):
        cu, cv = rng.uniform(0, 1), rng.uniform(0.1, 0.6)
        sig = rng.uniform(0.01, 0.04)
        amp = rng.uniform(0.25, 1.0) * peak
        col = rng.uniform(0.6, 1.0, 3).astype(np.float32)
        du = np.minimum(np.abs(U - cu), 1.0 - np.abs(U - cu))
        lobe = amp * np.exp(-(du ** 2 + (V - cv) ** 2) / (2 * sig * sig))
        img += lobe[..., None].astype(np.float32) * col
    img *= (1.0 + rng.uniform(-0.05, 0.05, img.shape[:2]).astype(np.float32))[..., None]
    img = np.minimum(img, peak).astype(np.float32)
    out = np.ones((height, width, 4), np.float32)
    out[..., :3] = img
    return out


def gbuffer(width: int, height: int, seed: int = SEED_BASE + 3, emissive: bool = False):
    """G-buffer planes [H,W,4] x3 (+emissive): P on a height field over x,z in [-20,20], y in [-5,5]."""
    rng = np.random.default_rng(seed)
    xs = np.linspace(-20.0, 20.0, width, dtype=np.float32)
    zs = np.linspace(20.0, -20.0, height, dtype=np.float32)
    X, Z = np.meshgrid(xs, zs)
    fx, fz = 0.35, 0.27
    Y = (2.5 * np.sin(fx * X) * np.cos(fz * Z) + 2.5 * np.sin(0.11 * X + 0.19 * Z)).astype(np.float32)
    dYdx = 2.5 * fx * np.cos(fx * X) * np.cos(fz * Z) + 2.5 * 0.11 * np.cos(0.11 * X + 0.19 * Z)
    dYdz = -2.5 * fz * np.sin(fx * X) * np.sin(fz * Z) + 2.5 * 0.19 * np.cos(0.11 * X + 0.19 * Z)
    N = np.stack([-dYdx, np.ones_like(dYdx), -dYdz], axis=-1).astype(np.float32)
    N += rng.normal(0.0, 0.15, N.shape).astype(np.float32)
    N /= np.sqrt((N * N).sum(-1, keepdims=True, dtype=np.float32))
    pos_ao = np.empty((height, width, 4), np.float32)
    pos_ao[..., 0], pos_ao[..., 1], pos_ao[..., 2] = X, Y, Z
    pos_ao[..., 3] = rng.uniform(0.0, 0.3, (height, width))
    nrm_rough = np.empty((height, width, 4), np.float32)
    nrm_rough[..., :3] = N
    nrm_rough[..., 3] = rng.uniform(0.04, 1.0, (height, width))
    alb_metal = np.empty((height, width, 4), np.float32)
    alb_metal[..., :3] = rng.uniform(0.02, 0.9, (height, width, 3))
    sel = rng.uniform(0, 1, (height, width))
    metal = np.where(sel < 0.2, 1.0, np.where(sel < 0.3, rng.uniform(0, 1, (height, width)), 0.0))
    alb_metal[..., 3] = metal
    planes = [pos_ao, nrm_rough, alb_metal]
    if emissive:
        em = np.zeros((height, width, 4), np.float32)
        em[..., :3] = rng.uniform(0, 1, (height, width, 3))
        em[..., 3] = np.where(rng.uniform(0, 1, (height, width)) < 0.05, rng.uniform(0, 4, (height, width)), 0.0)
        planes.append(em)
    return [np.ascontiguousarray(p) for p in planes]


def scene_constants(width: int, height: int, spec_mips: int, seed: int = SEED_BASE + 3, n_point: int = 4,
                    n_spot: int = 0, directional: bool = True, hdri_offset: float = 0.0, casters: bool = False,
                    point_range: float = 50.0):
    """PerFrameData + PerViewLightingData: n_point point lights 3..15 units above the field, brightness
    300..1500, range 50 (Data/Levels/Default.xml:223-264 spans 20-200 / 35-1500), one directional light
    normalize(-0.6,-1,0.3) x 0.9 (Default.xml:202-221), camera (0,5,-17) (EnvironmentMapUnitTest.xml)."""
    rng = np.random.default_rng(seed + 101)
    pf = PerFrameData()
    C.memset(C.byref(pf), 0, C.sizeof(pf))
    L = pf.Lights
    L.numPointLights = n_point
    for i in range(n_point):
        l = L.point_lights[i]
        l.position.x, l.position.z = rng.uniform(-18, 18), rng.uniform(-18, 18)
        l.position.y = rng.uniform(3, 15) + 5.0
        l.range = point_range
        l.color.x, l.color.y, l.color.z = rng.uniform(0.4, 1.0, 3)
        l.brightness = rng.uniform(300, 1500)
    L.numSpotLights = n_spot
    for i in range(n_spot):
        l = L.spot_lights[i]
        l.position.x, l.position.z = rng.uniform(-15, 15), rng.uniform(-15, 15)
        l.position.y = rng.uniform(8, 14)
        d = np.array([rng.uniform(-0.4, 0.4), -1.0, rng.uniform(-0.4, 0.4)])
        l.spotDir.x, l.spotDir.y, l.spotDir.z = d          # deliberately not unit: the shader normalises
        l.innerConeAngle = rng.uniform(0.3, 0.5)
        l.outerConeAngle = l.innerConeAngle + rng.uniform(0.15, 0.35)
        l.color.x, l.color.y, l.color.z = rng.uniform(0.4, 1.0, 3)
        l.brightness = rng.uniform(300, 1500)
        l.range = 100.0
    if casters:
        L.numPointCasters = 1
        l = L.point_casters[0]
        l.position.x, l.position.y, l.position.z = 3.0, 12.0, -2.0
        l.range = 30.0
        l.color.x, l.color.y, l.color.z = 1.0, 0.8, 0.6
        l.brightness = 700.0
        L.numSpotCasters = 1
        s = L.spot_casters[0]
        s.position.x, s.position.y, s.position.z = -6.0, 11.0, 4.0
        s.spotDir.x, s.spotDir.y, s.spotDir.z = 0.1, -1.0, -0.2
        s.innerConeAngle, s.outerConeAngle = 0.35, 0.6
        s.color.x, s.color.y, s.color.z = 0.7, 0.9, 1.0
        s.brightness = 900.0
    if directional:
        d = L.directional
        v = np.array([-0.6, -1.0, 0.3])
        v /= np.linalg.norm(v)
        d.lightDirection.x, d.lightDirection.y, d.lightDirection.z = v
        d.brightness = 0.9
        d.color.x = d.color.y = d.color.z = 1.0
        d.enabled = 1
    pf.fAmbientLightingFactor = 0.055
    pf.fHDRIOffsetInRadians = hdri_offset
    pv = PerViewLightingData()
    C.memset(C.byref(pv), 0, C.sizeof(pv))
    pv.CameraPosition.x, pv.CameraPosition.y, pv.CameraPosition.z = 0.0, 5.0, -17.0
    pv.MaxEnvMapLODLevels = float(spec_mips)
    pv.ScreenDimensions.x, pv.ScreenDimensions.y = float(width), float(height)
    pv.EnvironmentMapDiffuseOnlyIllumination = 0
    return pf, pv


def hdr_image(width: int, height: int, seed: int = SEED_BASE + 4, max_value: float = 8.0) -> np.ndarray:
    """Scene-referred RGBA32F [H,W,4] in 0..max_value: gradient + rectangles + 1-px lines + diagonal stripes + noise."""
    rng = np.random.default_rng(seed)
    y = (np.arange(height, dtype=np.float32) + 0.5) / height
    x = (np.arange(width, dtype=np.float32) + 0.5) / width
    X, Y = np.meshgrid(x, y)
    img = np.stack([0.2 + 1.5 * X, 0.1 + 1.2 * Y, 0.3 + 0.8 * (1 - X) * Y], axis=-1).astype(np.float32)
    for _ in range(24):
        x0, y0 = rng.integers(0, max(width - 8, 1)), rng.integers(0, max(height - 8, 1))
        w, h = rng.integers(4, max(width // 6, 5)), rng.integers(4, max(height // 6, 5))
        img[y0:y0 + h, x0:x0 + w, :] = rng.uniform(0.0, max_value, 3).astype(np.float32)
    for _ in range(16):
        if rng.uniform() < 0.5:
            img[rng.integers(0, height), :, :] = rng.uniform(0, max_value, 3)
        else:
            img[:, rng.integers(0, width), :] = rng.uniform(0, max_value, 3)
    xi = np.arange(width)[None, :]
    yi = np.arange(height)[:, None]
    stripes = (((xi + yi) // 3) % 2).astype(np.float32)
    band = (Y > 0.6) & (Y < 0.8)
    img[band] = img[band] * 0.25 + stripes[band][:, None] * np.array([2.0, 1.5, 1.0], np.float32)
    img *= (1.0 + rng.uniform(-0.08, 0.08, (height, width)).astype(np.float32))[..., None]
    img = np.clip(img, 0.0, max_value).astype(np.float32)
    out = np.ones((height, width, 4), np.float32)
    out[..., :3] = img
    out[..., 3] = rng.uniform(0, 1, (height, width)).astype(np.float32)   # alpha must pass through the tonemapper
    return out


def default_tonemapper() -> TonemapperParams:
    """PostProcess.h:84-91 defaults: REC_709 content, sRGB curve, 200 nits, gamma on."""
    return TonemapperParams(COLOR_SPACE_REC_709, DISPLAY_CURVE_SRGB, 200.0, 1, 1.0)
