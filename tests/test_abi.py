"""CPU (-m "not gpu"): the C-ABI library loads, exports every symbol include/vqcuda.h declares, its POD
structs have the reference cbuffer layouts, and the host-side setup functions equal the oracle and the
reference's own A_CPU build bit for bit. No compute calls without a GPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_all_exported(vq):
    hdr = open(os.path.join(ROOT, "include", "vqcuda.h")).read()
    declared = sorted(set(re.findall(r"VQ_API\s+[\w\s\*]+?\b(vq_\w+)\s*\(", hdr)))
    assert declared, "no declarations parsed"
    assert sorted(vq.ABI_SYMBOLS) == declared, set(declared) ^ set(vq.ABI_SYMBOLS)
    for name in declared:
        assert hasattr(vq.lib, name), f"libvqcuda.so does not export {name}"


def test_struct_layouts(vq):
    # SURVEY.md A26-A29; static_asserts in include/vq_shader_data.h say the same for the C side
    assert C.sizeof(vq.PointLight) == 48 and C.sizeof(vq.SpotLight) == 64 and C.sizeof(vq.DirectionalLight) == 40
    assert vq.SceneLighting.shadowViewDirectional.offset == 64
    assert vq.SceneLighting.point_lights.offset == 128
    assert vq.SceneLighting.spot_lights.offset == 5168
    assert C.sizeof(vq.SceneLighting) == 7088
    assert C.sizeof(vq.PerFrameData) == 7120
    assert C.sizeof(vq.PerViewLightingData) == 320 and vq.PerViewLightingData.CameraPosition.offset == 288
    assert C.sizeof(vq.TonemapperParams) == 20


def test_library_contains_sm100a_code():
    import subprocess
    so = os.path.join(ROOT, "vqengine_b200", "libvqcuda.so")
    out = subprocess.run(["cuobjdump", "--list-elf", so], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in out.stdout, out.stdout


@pytest.mark.parametrize("args", [(1920, 1080, 1920, 1080, 3840, 2160), (1478, 831, 1920, 1080, 1920, 1080),
                                  (1286, 723, 1286, 723, 1920, 1080), (640, 360, 640, 360, 1280, 720)])
def test_easu_con_bitexact(vq, orc, args):
    a = list(vq.fsr_easu_con(*args))
    assert a == list(orc.fsr_easu_con(*args))
    if orc.ref():
        assert a == list(orc.fsr_easu_con(*args, which="ref"))


@pytest.mark.parametrize("stops", [0.0, 0.2, 1.0, 2.0, 6.64])
def test_rcas_con_bitexact(vq, orc, stops):
    a = list(vq.fsr_rcas_con(stops))
    assert a == list(orc.fsr_rcas_con(stops))
    if orc.ref():
        assert a == list(orc.fsr_rcas_con(stops, which="ref"))


@pytest.mark.parametrize("sharp", [0.0, 0.25, 0.8, 1.0, 1.7, -0.3])
def test_cas_setup_bitexact(vq, orc, sharp):
    a = list(vq.cas_setup(sharp, 3840, 2160, 3840, 2160))
    assert a == list(orc.cas_setup(sharp, 3840, 2160, 3840, 2160))
    if orc.ref():
        assert a == list(orc.cas_setup(sharp, 3840, 2160, 3840, 2160, which="ref"))


@pytest.mark.parametrize("rect", [(0, 0, 3840, 2160), (0, 0, 64, 64), (0, 0, 65, 1), (128, 64, 500, 300), (0, 0, 4096, 4096)])
def test_spd_setup_equal(vq, orc, rect):
    (dx, dy), c = vq.spd_setup(rect[2], rect[3], -1, rect[0], rect[1])
    d, o, n = orc.spd_setup(rect)
    assert [dx, dy] == d and list(c.workGroupOffset) == o and [c.numWorkGroups, c.mips] == n
    if orc.ref():
        assert (d, o, n) == orc.spd_setup(rect, which="ref")


def test_layout_helpers(vq, orc):
    for (w, h) in [(2048, 1024), (4096, 2048), (4096, 4096), (512, 512), (1, 1), (2, 1)]:
        assert vq.mip_level_count(w, h) == orc.lib().orc_mip_level_count(w, h)
    assert vq.mip_level_count(2048, 1024) == 11 and vq.mip_level_count(4096, 2048) == 12
    assert vq.mip_level_count(512, 512) - 1 == 9                                          # spec cube mips, F5
    assert vq.cubemap_texel_count(512, 9) == 2097144 and vq.cubemap_texel_count(512, 7) == 2097024   # A33
    assert vq.cubemap_row_count(512, 9) == 6 * (512 + 256 + 128 + 64 + 32 + 16 + 8 + 4 + 2)
    assert vq.cubemap_offset(8, 1, 2) == 6 * 64 + 2 * 16
    assert vq.pyramid_offset(8, 4, 2) == 32 + 8


def test_no_cpu_fallback(vq):
    """without a CUDA device the product must fail loudly, not compute on the CPU"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    rc = vq.lib.vq_ctx_create(0, C.byref(h))
    assert rc == vq.VQ_ERR_NO_DEVICE
    assert b"no CPU path" in vq.lib.vq_last_error()
    with pytest.raises(vq.VqError):
        vq.Context(0)


def test_product_does_not_touch_oracle():
    """the product path must not import / include / link / dlopen anything under oracle/"""
    import subprocess
    pkg = os.path.join(ROOT, "vqengine_b200")
    bad = re.compile(r"(import\s+oracle|from\s+oracle|oracle_lib|liboracle|libffxref|#\s*include\s*[\"<][^\">]*oracle)")
    # the package, the public headers, the C example and the GPU-box helper scripts (only tests/, smoke() and bench.py's
    # cpu_baseline / --impl reference legs may execute the checker)
    for top in (pkg, os.path.join(ROOT, "include"), os.path.join(ROOT, "examples"), os.path.join(ROOT, "tools")):
        for dirpath, _, files in os.walk(top):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".c", ".h", ".hpp", ".sh")):
                    src = open(os.path.join(dirpath, f), errors="ignore").read()
                    assert not bad.search(src), f"{f} reaches into oracle/"
    so = os.path.join(pkg, "libvqcuda.so")
    assert "oracle" not in subprocess.run(["ldd", so], capture_output=True, text=True).stdout
    strings = subprocess.run(["strings", so], capture_output=True, text=True).stdout
    assert "liboracle" not in strings and "orc_" not in strings


def test_header_is_plain_c_and_c_client_links(tmp_path):
    """include/vqcuda.h must be consumable from C (the drop-in boundary carries no C++ or torch types): the example client
    compiles as C99 with -Wall -Wextra -Werror and links against nothing but libvqcuda.so and the CUDA runtime."""
    import os
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    if not shutil.which("gcc") or not os.path.exists(os.path.join(cuda, "include", "cuda_runtime_api.h")):
        pytest.skip("gcc / CUDA headers not available")
    obj, exe = str(tmp_path / "c_client.o"), str(tmp_path / "c_client")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-c", os.path.join(root, "examples", "c_client.c"),
                           "-I", os.path.join(root, "include"), "-I", os.path.join(cuda, "include"), "-o", obj])
    subprocess.check_call(["gcc", obj, "-L", os.path.join(root, "vqengine_b200"), "-lvqcuda", "-L", os.path.join(cuda, "lib64"),
                           "-lcudart", "-lm", "-o", exe])
    # and a translation unit that only includes the headers, as strict C89-style C
    src = tmp_path / "hdr_only.c"
    src.write_text('#include "vqcuda.h"\nint main(void) { return (int)sizeof(VqPerFrameData) == 0; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-c", str(src), "-I", os.path.join(root, "include"),
                           "-o", str(tmp_path / "hdr_only.o")])
