"""ctypes mirror of include/vq_shader_data.h and of the descriptor structs of include/vqcuda.h — plain data, no library.

Kept apart from the package's __init__ (which loads libvqcuda.so) so that code which only needs the struct layouts and the
seeded synthetic inputs — bench.py's CPU arms, the oracle-side tests — can import them WITHOUT loading the CUDA library:
    import vqengine_b200.shader_data   # still runs the package __init__; bench.py's reference arm therefore imports this
                                       # file and synth.py under a library-free alias package (bench.py: _nolib()).
"""
from __future__ import annotations

import ctypes as C

VQ_OK = 0
VQ_ERR_INVALID_ARG = -1
VQ_ERR_CUDA = -2
VQ_ERR_UNSUPPORTED = -3
VQ_ERR_NO_DEVICE = -4
VQ_ERR_OUT_OF_MEMORY = -5

VQ_ADDRESS_WRAP, VQ_ADDRESS_CLAMP = 0, 1
COLOR_SPACE_REC_709, COLOR_SPACE_REC_2020 = 0, 1
DISPLAY_CURVE_SRGB, DISPLAY_CURVE_ST2084, DISPLAY_CURVE_LINEAR = 0, 1, 2

f32 = C.c_float
i32 = C.c_int32
u32 = C.c_uint32


# ---- include/vq_shader_data.h ---------------------------------------------------------------
class Float2(C.Structure):
    _fields_ = [("x", f32), ("y", f32)]


class Float3(C.Structure):
    _fields_ = [("x", f32), ("y", f32), ("z", f32)]


class Float4(C.Structure):
    _fields_ = [("x", f32), ("y", f32), ("z", f32), ("w", f32)]


class Matrix(C.Structure):
    _fields_ = [("m", f32 * 16)]


class PointLight(C.Structure):
    _fields_ = [("position", Float3), ("range", f32), ("color", Float3), ("brightness", f32),
                ("attenuation", Float3), ("depthBias", f32)]


class SpotLight(C.Structure):
    _fields_ = [("position", Float3), ("outerConeAngle", f32), ("color", Float3), ("brightness", f32),
                ("spotDir", Float3), ("depthBias", f32), ("innerConeAngle", f32), ("range", f32),
                ("dummy1", f32), ("dummy2", f32)]


class DirectionalLight(C.Structure):
    _fields_ = [("lightDirection", Float3), ("brightness", f32), ("color", Float3), ("depthBias", f32),
                ("shadowing", i32), ("enabled", i32)]


class SceneLighting(C.Structure):
    _fields_ = [("numPointLights", i32), ("numSpotLights", i32), ("numPointCasters", i32), ("numSpotCasters", i32),
                ("directional", DirectionalLight), ("_pad_matrix_align", u32 * 2),
                ("shadowViewDirectional", Matrix),
                ("point_lights", PointLight * 100), ("point_casters", PointLight * 5),
                ("spot_lights", SpotLight * 20), ("spot_casters", SpotLight * 5),
                ("shadowViews", Matrix * 5)]


class PerFrameData(C.Structure):
    _fields_ = [("Lights", SceneLighting),
                ("f2PointLightShadowMapDimensions", Float2), ("f2SpotLightShadowMapDimensions", Float2),
                ("f2DirectionalLightShadowMapDimensions", Float2),
                ("fAmbientLightingFactor", f32), ("fHDRIOffsetInRadians", f32)]


class PerViewLightingData(C.Structure):
    _fields_ = [("matView", Matrix), ("matViewToWorld", Matrix), ("matProjInverse", Matrix),
                ("WorldFrustumPlanes", Float4 * 6), ("CameraPosition", Float3), ("MaxEnvMapLODLevels", f32),
                ("ScreenDimensions", Float2), ("EnvironmentMapDiffuseOnlyIllumination", i32), ("pad1", f32)]


class TonemapperParams(C.Structure):
    _fields_ = [("ContentColorSpace", i32), ("OutputDisplayCurve", i32),
                ("DisplayReferenceBrightnessLevel", f32), ("ToggleGammaCorrection", i32), ("UIHDRBrightness", f32)]


class BlurParams(C.Structure):
    _fields_ = [("iImageSizeX", i32), ("iImageSizeY", i32)]


class SpdConstants(C.Structure):
    _fields_ = [("mips", u32), ("numWorkGroups", u32), ("workGroupOffset", u32 * 2)]


class DiffuseIrradianceParams(C.Structure):
    _fields_ = [("step", f32), ("n_phi", i32), ("n_theta", i32), ("src_mip", i32)]


assert C.sizeof(PointLight) == 48 and C.sizeof(SpotLight) == 64 and C.sizeof(DirectionalLight) == 40
assert C.sizeof(SceneLighting) == 7088 and C.sizeof(PerFrameData) == 7120 and C.sizeof(PerViewLightingData) == 320


# ---- include/vqcuda.h -----------------------------------------------------------------------
class Image(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("width", i32), ("height", i32), ("pitch_bytes", C.c_size_t)]


class Cubemap(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("res", i32), ("mips", i32)]


class Pyramid(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("width", i32), ("height", i32), ("levels", i32)]


class GBuffer(C.Structure):
    _fields_ = [("position_ao", Image), ("normal_roughness", Image), ("albedo_metalness", Image), ("emissive", Image)]


class EnvironmentMaps(C.Structure):
    _fields_ = [("irradiance_diffuse", Cubemap), ("irradiance_specular", Cubemap), ("brdf_lut", Image)]


class ShadowMaps(C.Structure):            # include/vqcuda.h VqShadowMaps
    _fields_ = [("point_cubes", C.c_void_p), ("point_res", C.c_int32),
                ("spot_maps", C.c_void_p), ("spot_width", C.c_int32), ("spot_height", C.c_int32),
                ("directional_map", C.c_void_p), ("directional_width", C.c_int32), ("directional_height", C.c_int32)]


class MaterialData(C.Structure):          # include/vq_shader_data.h VqMaterialData (LightingConstantBufferData.h:126-143)
    _fields_ = [("diffuse", Float3), ("alpha", f32), ("emissiveColor", Float3), ("emissiveIntensity", f32),
                ("specular", Float3), ("normalMapMipBias", f32), ("uvScaleOffset", Float4),
                ("roughness", f32), ("metalness", f32), ("displacement", f32), ("textureConfig", f32)]


assert C.sizeof(MaterialData) == 80

TEXCFG_DIFFUSE, TEXCFG_NORMAL, TEXCFG_AO, TEXCFG_ALPHA_MASK, TEXCFG_ROUGHNESS, TEXCFG_METALLIC, TEXCFG_HEIGHT, \
    TEXCFG_EMISSIVE, TEXCFG_ORM = (1 << b for b in range(9))
MATERIAL_TEXTURE_SLOTS = ("diffuse", "normals", "emissive", "metalness", "roughness", "occl_rough_metal", "local_ao")


class Texture2D(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("width", i32), ("height", i32), ("levels", i32)]


class MaterialTextures(C.Structure):
    _fields_ = [(k, Texture2D) for k in MATERIAL_TEXTURE_SLOTS]


class PeerSignal(C.Structure):            # include/vqcuda.h VqPeerSignal: end-of-pass rendezvous run by the kernel's last CTA
    _fields_ = [("flags", C.c_void_p * 8), ("n_ranks", C.c_int32), ("my_index", C.c_int32), ("epoch", C.c_uint32)]


class HdrInfo(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("data_offset", C.c_uint64), ("flat", C.c_int32),
                ("reserved", C.c_int32)]


class SurfaceInputs(C.Structure):
    _fields_ = [("position_u", Image), ("normal_v", Image), ("tangent_m", Image), ("ssao", Image)]
