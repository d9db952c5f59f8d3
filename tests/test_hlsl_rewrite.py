"""CPU: the token-level rewrite that lets g++ compile the reference's shader text (oracle/ref_shim/hlsl_to_cpp.py). It must respell
only what C++ cannot parse and leave every expression alone; each rule is pinned on a snippet here."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("hlsl_to_cpp", os.path.join(ROOT, "oracle", "ref_shim", "hlsl_to_cpp.py"))
h2c = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(h2c)


def rw(text, **kw):
    return h2c.rewrite(text, **kw)


def test_float_literals_get_a_suffix_and_nothing_else_changes():
    assert rw("x = 1.0 - a * 0.999 + 5.0;") == "x = 1.0f - a * 0.999f + 5.0f;"
    assert rw("return float(bits) * 2.3283064365386963e-10;") == "return float(bits) * 2.3283064365386963e-10f;"
    assert rw("a = 1.0f + 2.5f; b = 1e-6f; i = 3; u = 16u; h = 0x7ef07ebb;") == "a = 1.0f + 2.5f; b = 1e-6f; i = 3; u = 16u; h = 0x7ef07ebb;"
    assert rw("v.x1 = m2.0;") == "v.x1 = m2.0;"                      # digits inside identifiers / member chains are not literals
    assert rw("float m1 = 2610.0 / 4096.0 / 4;") == "float m1 = 2610.0f / 4096.0f / 4;"


def test_scalar_swizzles_become_splat_calls():
    assert rw("float3 a = 0.04f.xxx;") == "float3 a = splat3(0.04f);"
    assert rw("float3 a = 0.0.xxx;") == "float3 a = splat3(0.0f);"
    assert rw("float2 a = 0.5f.xx;") == "float2 a = splat2(0.5f);"
    assert rw("max((1.0f - roughness).xxx, F0)") == "max(splat3((1.0f - roughness)), F0)"
    assert rw("o.color = float4(I_total, Surface.roughness.r);") == "o.color = float4(I_total, Surface.roughness);"
    assert rw("return g_depth_buffer[index].xxxx;") == "return splat4(g_depth_buffer[index]);"
    assert rw("float f = v; return f.xxxx;", scalars=("f",)) == "float f = v; return splat4(f);"
    # vector swizzles are left to the emulation header
    assert rw("c = a.xyz * b.rgb + t.xxx;") == "c = a.xyz * b.rgb + t.xxx;"


def test_annotations_and_qualifiers_are_dropped():
    assert rw("Texture2D texDiffuse : register(t0);").split() == "Texture2D texDiffuse ;".split()
    assert rw("float4 color : SV_TARGET0;").split() == "float4 color ;".split()
    assert rw("float3 position : POSITION;\nfloat2 uv : TEXCOORD0;").split() == "float3 position ; float2 uv ;".split()
    assert rw("[numthreads(8, 8, 1)]\nvoid CSMain(uint3 id : SV_DispatchThreadID) {}").split() == "void CSMain(uint3 id ) {}".split()
    assert rw("float f(in const float3 c, const in float d, in float e)") == "float f(const float3 c, const float d, float e)"
    assert rw("void g(inout float r, out float3 v)") == "void g(float& r, float3& v)"
    assert rw("groupshared float s[16][16];\ngloballycoherent RWTexture2DArray<float4> d;").split() == "float s[16][16]; RWTexture2DArray<float4> d;".split()
    # a ternary's colon is not a semantic
    assert rw("x = c ? a : b;") == "x = c ? a : b;"


def test_cbuffers_casts_discard_textures():
    assert rw("cbuffer CB : register(b0) { PerFrameData cbPerFrame; }").split() == "PerFrameData cbPerFrame;".split()
    assert rw("cbuffer P\n{\n int a;\n float b;\n}\nint z;").split() == "int a; float b; int z;".split()
    assert rw("PSOutput o = (PSOutput)0;") == "PSOutput o = PSOutput{};"
    assert rw("BRDF_Surface s = (BRDF_Surface) 0;") == "BRDF_Surface s = BRDF_Surface{};"
    assert rw("if (a < 0.01f)\n  discard;") == "if (a < 0.01f)\n  throw hl::Discard();"
    assert rw("Texture2D<float4> t; Texture2D<float> d; RWTexture2D<float> o;") == "Texture2D t; Texture2DF d; RWTexture2DF o;"
    assert rw("int2 p = DispatchThreadID.xy;") == "int2 p = int2(DispatchThreadID.xy);"


def test_expressions_are_untouched():
    src = "const float3 specular = D * F * G / denom;\nreturn F0 + (float3(1, 1, 1) - F0) * pow(1.0f - max(0.0f, dot(N, V)), 5.0f);"
    assert rw(src) == src


def test_swizzle_generator_covers_every_combination(tmp_path):
    out = tmp_path / "sw.inc"
    h2c.gen_swizzles(str(out))
    text = out.read_text()
    for n, count in ((2, 2 * (4 + 8 + 16)), (3, 2 * (9 + 27 + 81)), (4, 2 * (16 + 64 + 256))):
        line = [l for l in text.splitlines() if l.startswith("#define HLSL_SWIZZLES_%d(T)" % n)][0]
        assert line.count("swz<T,%d," % n) == count
    assert " xyz;" in text and " rgb;" in text and " xyww;" in text and " zw;" in text
