"""CPU (-m "not gpu"): byte compatibility of include/vq_shader_data.h with the REFERENCE'S OWN shared CPU/GPU struct header.
Shaders/LightingConstantBufferData.h is compiled unmodified as its CPU side (VQ_CPU; <DirectXMath.h> is a stand-in with the
public XMFLOATn / XMMATRIX storage layouts, oracle/ref_shim/dxmath_shim) next to our header, and sizeof / offsetof of every
member the hot path reads are compared. Skipped where /root/reference is absent (the static_asserts in our header and
tests/test_abi.py::test_struct_layouts still hold the numbers)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("VQ_REFERENCE", "/root/reference")

PAIRS = {   # reference struct -> (ours, members)
    "PointLight": ("VqPointLight", ["position", "range", "color", "brightness", "attenuation", "depthBias"]),
    "SpotLight": ("VqSpotLight", ["position", "outerConeAngle", "color", "brightness", "spotDir", "depthBias", "innerConeAngle", "range"]),
    "DirectionalLight": ("VqDirectionalLight", ["lightDirection", "brightness", "color", "depthBias", "shadowing", "enabled"]),
    "SceneLighting": ("VqSceneLighting", ["numPointLights", "numSpotLights", "numPointCasters", "numSpotCasters", "directional",
                                          "shadowViewDirectional", "point_lights", "point_casters", "spot_lights", "spot_casters", "shadowViews"]),
    "PerFrameData": ("VqPerFrameData", ["Lights", "f2PointLightShadowMapDimensions", "f2SpotLightShadowMapDimensions",
                                        "f2DirectionalLightShadowMapDimensions", "fAmbientLightingFactor", "fHDRIOffsetInRadians"]),
    "PerViewLightingData": ("VqPerViewLightingData", ["matView", "matViewToWorld", "matProjInverse", "WorldFrustumPlanes", "CameraPosition",
                                                      "MaxEnvMapLODLevels", "ScreenDimensions", "EnvironmentMapDiffuseOnlyIllumination"]),
    "MaterialData": ("VqMaterialData", ["diffuse", "alpha", "emissiveColor", "emissiveIntensity", "specular", "normalMapMipBias",
                                        "uvScaleOffset", "roughness", "metalness", "displacement", "textureConfig"]),
}


def test_layouts_match_the_reference_header(tmp_path):
    hdr = os.path.join(REF, "Shaders", "LightingConstantBufferData.h")
    if not os.path.exists(hdr) or not shutil.which("g++"):
        pytest.skip("reference header / g++ not available")
    lines = ['#include <cstddef>', '#include <cstdio>', '#include "LightingConstantBufferData.h"', '#include "vq_shader_data.h"',
             'int main() { int bad = 0;']
    for rs, (ours, members) in PAIRS.items():
        lines.append(f'  if (sizeof(VQ_SHADER_DATA::{rs}) != sizeof({ours})) {{ std::printf("sizeof {rs}: %zu vs %zu\\n", sizeof(VQ_SHADER_DATA::{rs}), sizeof({ours})); ++bad; }}')
        for m in members:
            lines.append(f'  if (offsetof(VQ_SHADER_DATA::{rs}, {m}) != offsetof({ours}, {m})) {{ std::printf("offsetof {rs}.{m}: %zu vs %zu\\n", '
                         f'offsetof(VQ_SHADER_DATA::{rs}, {m}), offsetof({ours}, {m})); ++bad; }}')
    # the texture-configuration bit field: our VQ_TEXCFG_* constants against the reference's own Has*Map() decoders, and the
    # array extents
    for ours, fn in [("VQ_TEXCFG_DIFFUSE", "HasDiffuseMap"), ("VQ_TEXCFG_NORMAL", "HasNormalMap"), ("VQ_TEXCFG_AO", "HasAmbientOcclusionMap"),
                     ("VQ_TEXCFG_ALPHA_MASK", "HasAlphaMask"), ("VQ_TEXCFG_ROUGHNESS", "HasRoughnessMap"), ("VQ_TEXCFG_METALLIC", "HasMetallicMap"),
                     ("VQ_TEXCFG_HEIGHT", "HasHeightMap"), ("VQ_TEXCFG_EMISSIVE", "HasEmissiveMap"), ("VQ_TEXCFG_ORM", "HasOcclusionRoughnessMetalnessMap")]:
        lines.append(f'  if (VQ_SHADER_DATA::{fn}({ours}) != 1 || VQ_SHADER_DATA::{fn}(0x1ff & ~{ours}) != 0) {{ std::printf("{ours} vs {fn}\\n"); ++bad; }}')
    for ours, theirs in [("VQ_NUM_LIGHTS_POINT", "NUM_LIGHTS__POINT"), ("VQ_NUM_LIGHTS_SPOT", "NUM_LIGHTS__SPOT"),
                         ("VQ_NUM_SHADOWING_LIGHTS_POINT", "NUM_SHADOWING_LIGHTS__POINT"), ("VQ_NUM_SHADOWING_LIGHTS_SPOT", "NUM_SHADOWING_LIGHTS__SPOT")]:
        lines.append(f'  if ({ours} != {theirs}) {{ std::printf("{ours} != {theirs}\\n"); ++bad; }}')
    lines += ['  std::printf("checked, %d mismatches, sizeof PerFrameData %zu\\n", bad, sizeof(VQ_SHADER_DATA::PerFrameData));', '  return bad; }']
    src = tmp_path / "layout.cpp"
    src.write_text("\n".join(lines))
    exe = str(tmp_path / "layout")
    subprocess.check_call(["g++", "-std=c++17", "-w", str(src), "-I", os.path.join(ROOT, "oracle", "ref_shim", "dxmath_shim"),
                           "-I", os.path.join(REF, "Shaders"), "-I", os.path.join(ROOT, "include"), "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    print(r.stdout)
    assert r.returncode == 0, r.stdout
    assert "0 mismatches" in r.stdout and "7120" in r.stdout
