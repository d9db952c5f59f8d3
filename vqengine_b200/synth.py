"""Seeded synthetic inputs for the shading passes; numpy only, no GPU.

The reference ships no HDRI assets and its forward pass consumes rasterised attributes, so every
input is synthetic: the same bytes are fed to the CUDA kernels and to the oracle.
"""
from __future__ import annotations

import ctypes as C
import numpy as np

from .shader_data import (PerFrameData, PerViewLightingData, TonemapperParams, COLOR_SPACE_REC_709, DISPLAY_CURVE_SRGB)

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


# ---- material textures + rasteriser interpolants ---------------------------------
def _value_noise(rng, res_w: int, res_h: int, cells: int, channels: int) -> np.ndarray:
    """smooth tileable value noise in [0,1], [res_h,res_w,channels] (bilinear upsample of a cells x cells lattice)"""
    lat = rng.uniform(0.0, 1.0, (cells, cells, channels)).astype(np.float32)
    ys = (np.arange(res_h, dtype=np.float32) + 0.5) * cells / res_h
    xs = (np.arange(res_w, dtype=np.float32) + 0.5) * cells / res_w
    y0, x0 = np.floor(ys).astype(int), np.floor(xs).astype(int)
    fy, fx = (ys - y0)[:, None, None], (xs - x0)[None, :, None]
    y1, x1 = (y0 + 1) % cells, (x0 + 1) % cells
    y0 %= cells; x0 %= cells
    top = lat[y0][:, x0] * (1 - fx) + lat[y0][:, x1] * fx
    bot = lat[y1][:, x0] * (1 - fx) + lat[y1][:, x1] * fx
    return top * (1 - fy) + bot * fy


def material_texture(kind: str, width: int, height: int, seed: int) -> np.ndarray:
    """level 0 of a synthetic RGBA8 material map, [H,W,4] uint8. kinds: albedo (alpha has cut-outs), normal (tangent-space,
    z-dominant), emissive (sparse), scalar (R = smooth noise + grain), orm"""
    rng = np.random.default_rng(seed)
    grain = rng.uniform(-0.06, 0.06, (height, width, 4)).astype(np.float32)
    if kind == "albedo":
        img = np.empty((height, width, 4), np.float32)
        img[..., :3] = 0.15 + 0.8 * _value_noise(rng, width, height, 8, 3)
        img[..., 3] = np.where(_value_noise(rng, width, height, 6, 1)[..., 0] < 0.28, 0.0, 1.0)
        img[..., :3] += grain[..., :3]
    elif kind == "normal":
        bump = _value_noise(rng, width, height, 16, 2) - 0.5
        n = np.stack([bump[..., 0] * 1.2, bump[..., 1] * 1.2, np.ones((height, width), np.float32)], -1)
        n /= np.sqrt((n * n).sum(-1, keepdims=True))
        img = np.concatenate([n * 0.5 + 0.5, np.ones((height, width, 1), np.float32)], -1)
    elif kind == "emissive":
        img = np.zeros((height, width, 4), np.float32)
        mask = _value_noise(rng, width, height, 10, 1)[..., 0] > 0.7
        img[..., :3] = np.where(mask[..., None], 0.3 + 0.7 * _value_noise(rng, width, height, 4, 3), 0.0)
        img[..., 3] = 1.0
    elif kind in ("scalar", "orm"):
        img = 0.1 + 0.85 * _value_noise(rng, width, height, 12, 4) + grain
    else:
        raise ValueError(kind)
    return np.ascontiguousarray(np.clip(np.rint(img * 255.0), 0, 255).astype(np.uint8))


def materials(n: int = 4, tex_res: int = 256, seed: int = SEED_BASE + 11, uniform: bool = False):
    """n materials cycling through the texture configurations the engine's material sets use (Data/Materials/*.xml):
    0 fully textured (separate roughness/metal/AO maps, emissive), 1 glTF style (albedo+normal+ORM),
    2 constants only (all SRVs null), 3 albedo+normal with uv tiling and a non-square, non-pow2 albedo.
    uniform=True gives every map of material 0 the same tex_res^2 size (how shipped material sets look); the default mixes
    sizes and aspect ratios to exercise the sampler. Returns (list[MaterialData], list[dict slot -> level-0 uint8 array or None])."""
    from .shader_data import MaterialData, TEXCFG_DIFFUSE, TEXCFG_NORMAL, TEXCFG_AO, TEXCFG_ROUGHNESS, TEXCFG_METALLIC, \
        TEXCFG_EMISSIVE, TEXCFG_ORM, MATERIAL_TEXTURE_SLOTS
    rng = np.random.default_rng(seed)
    mats, texs = [], []
    for i in range(n):
        m = MaterialData()
        C.memset(C.byref(m), 0, C.sizeof(m))
        m.diffuse.x, m.diffuse.y, m.diffuse.z = rng.uniform(0.5, 1.0, 3)
        m.alpha = 1.0
        m.emissiveColor.x, m.emissiveColor.y, m.emissiveColor.z = rng.uniform(0.2, 1.0, 3)
        m.specular.x = m.specular.y = m.specular.z = 1.0
        m.uvScaleOffset.x, m.uvScaleOffset.y, m.uvScaleOffset.z, m.uvScaleOffset.w = 1.0, 1.0, 0.0, 0.0
        m.roughness, m.metalness = rng.uniform(0.3, 1.0), rng.uniform(0.0, 1.0)
        t = {k: None for k in MATERIAL_TEXTURE_SLOTS}
        s = seed * 131 + i * 17
        kind = i % 4
        if kind == 0:
            t["diffuse"] = material_texture("albedo", tex_res, tex_res, s + 1)
            t["normals"] = material_texture("normal", tex_res, tex_res, s + 2)
            hr = tex_res if uniform else tex_res // 2
            t["emissive"] = material_texture("emissive", hr, hr, s + 3)
            t["metalness"] = material_texture("scalar", hr, hr, s + 4)
            t["roughness"] = material_texture("scalar", tex_res, hr, s + 5)
            t["local_ao"] = material_texture("scalar", hr, tex_res, s + 6)
            cfg = TEXCFG_DIFFUSE | TEXCFG_NORMAL | TEXCFG_EMISSIVE | TEXCFG_METALLIC | TEXCFG_ROUGHNESS | TEXCFG_AO
            m.emissiveIntensity = 2.5
            m.normalMapMipBias = 0.0 if uniform else -0.5
        elif kind == 1:
            t["diffuse"] = material_texture("albedo", tex_res, tex_res, s + 1)
            t["normals"] = material_texture("normal", tex_res, tex_res, s + 2)
            t["occl_rough_metal"] = material_texture("orm", tex_res, tex_res, s + 3)
            cfg = TEXCFG_DIFFUSE | TEXCFG_NORMAL | TEXCFG_ORM
            m.uvScaleOffset.x, m.uvScaleOffset.y, m.uvScaleOffset.z, m.uvScaleOffset.w = 3.0, 2.0, 0.25, 0.4
        elif kind == 2:
            cfg = 0
            m.emissiveIntensity = 0.4
        else:
            w3 = max(12, (tex_res * 3) // 4 - 4)            # neither square nor a power of two
            t["diffuse"] = material_texture("albedo", w3, tex_res // 2 + 6, s + 1)
            t["normals"] = material_texture("normal", tex_res, tex_res, s + 2)
            cfg = TEXCFG_DIFFUSE | TEXCFG_NORMAL
            m.uvScaleOffset.x, m.uvScaleOffset.y, m.uvScaleOffset.z, m.uvScaleOffset.w = 7.5, 7.5, -0.3, 0.1
            m.normalMapMipBias = 0.75
        m.textureConfig = float(cfg)
        mats.append(m)
        texs.append(t)
    return mats, texs


def surface_inputs(width: int, height: int, n_materials: int, seed: int = SEED_BASE + 12, ssao: bool = True,
                   uv_scale: float = 0.1):
    """PSInput planes of a height field seen from above: position_u, normal_v, tangent_m [H,W,4] (+ssao [H,W,1]).
    uv = world xz * uv_scale (so texel footprints grow with the pixel pitch: every mip level is exercised between 96x54 and 4K),
    materials in 64-pixel screen blocks with ragged borders, interpolated (non-unit) normals/tangents."""
    rng = np.random.default_rng(seed)
    g = gbuffer(width, height, seed)
    P, N = g[0][..., :3], g[1][..., :3]
    pos_u = np.empty((height, width, 4), np.float32)
    nrm_v = np.empty((height, width, 4), np.float32)
    tan_m = np.empty((height, width, 4), np.float32)
    pos_u[..., :3] = P
    pos_u[..., 3] = P[..., 0] * np.float32(uv_scale)
    nrm_v[..., :3] = N * rng.uniform(0.6, 1.0, (height, width, 1)).astype(np.float32)
    nrm_v[..., 3] = P[..., 2] * np.float32(uv_scale)
    T = np.stack([np.ones((height, width), np.float32), 0.3 * N[..., 0], 0.1 * np.ones((height, width), np.float32)], -1)
    tan_m[..., :3] = T * rng.uniform(0.7, 1.3, (height, width, 1)).astype(np.float32)
    by, bx = np.meshgrid(np.arange(height) // 64, np.arange(width) // 64, indexing="ij")
    jitter = rng.integers(0, 40, (height, width)) == 0
    mid = (by * 3 + bx + jitter) % n_materials
    tan_m[..., 3] = mid.astype(np.float32)
    planes = [np.ascontiguousarray(pos_u), np.ascontiguousarray(nrm_v), np.ascontiguousarray(tan_m)]
    if ssao:
        planes.append(np.ascontiguousarray(rng.uniform(0.2, 1.0, (height, width, 1)).astype(np.float32)))
    return planes


def sky_view_proj(yaw: float, pitch: float, fov_y: float, aspect: float, near: float = 0.1, far: float = 1000.0):
    """SceneView.EnvironmentMapViewProj (Scene.cpp:573-584): a camera at the origin rotated by yaw (about +Y) and pitch
    (about +X), times a left-handed D3D perspective projection (XMMatrixPerspectiveFovLH), row-vector convention
    (v' = v * M). Returns (view_proj, inverse) as float64 4x4 row-major arrays; the caller rounds to fp32."""
    cy, sy, cp, sp = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch)
    rot_y = np.array([[cy, 0, -sy, 0], [0, 1, 0, 0], [sy, 0, cy, 0], [0, 0, 0, 1]], dtype=np.float64)
    rot_x = np.array([[1, 0, 0, 0], [0, cp, sp, 0], [0, -sp, cp, 0], [0, 0, 0, 1]], dtype=np.float64)
    world = rot_x @ rot_y                      # camera orientation (row vectors: pitch first, then yaw)
    view = np.linalg.inv(world)
    h = 1.0 / np.tan(0.5 * fov_y)
    w = h / aspect
    q = far / (far - near)
    proj = np.array([[w, 0, 0, 0], [0, h, 0, 0], [0, 0, q, 1], [0, 0, -near * q, 0]], dtype=np.float64)
    vp = view @ proj
    return vp, np.linalg.inv(vp)


def smooth_hdri(width: int, height: int, seed: int = SEED_BASE + 6, peak: float = 4.0) -> np.ndarray:
    """Band-limited equirect radiance (no per-texel noise): sums of low-order sinusoids in (phi, theta), >= 0, alpha 1.
    Used where a single bilinear sample is compared at 1e-4: texel-to-texel contrast stays ~1e-2."""
    rng = np.random.default_rng(seed)
    v, u = np.meshgrid((np.arange(height) + 0.5) / height, (np.arange(width) + 0.5) / width, indexing="ij")
    img = np.zeros((height, width, 4), dtype=np.float64)
    for c in range(3):
        acc = np.full((height, width), 1.0)
        for _ in range(4):
            ku, kv = rng.integers(1, 4), rng.integers(1, 3)
            acc += 0.2 * np.sin(2 * np.pi * ku * u + rng.uniform(0, 6.28)) * np.cos(np.pi * kv * v + rng.uniform(0, 6.28))
        img[..., c] = acc * (peak / 2.0)
    img[..., 3] = 1.0
    return np.ascontiguousarray(np.clip(img, 0.0, None).astype(np.float32))
