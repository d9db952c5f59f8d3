"""-m gpu: the C++ host layer (vq::VQRenderer, mirror of the reference's renderer front end) in the shape of the
reference's own automated test — a smoke run of N frames that must exit 0 (Scripts/TestVQE.bat:90-105)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "vqengine_b200", "host", "vq_headless_test")


@pytest.mark.gpu
def test_headless_smoke_run():
    assert os.path.exists(EXE), "host layer not built (python -c 'import __graft_entry__ as g; g.build()')"
    r = subprocess.run([EXE, "-Test", "-TestFrames=20", "-W=320", "-H=180"], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, (r.returncode, r.stderr)
    assert "20 frames OK" in r.stdout and ".hdr file image" in r.stdout


def test_host_layer_exports_reference_shaped_api():
    so = os.path.join(ROOT, "vqengine_b200", "host", "libvqhost.so")
    if not os.path.exists(so):
        pytest.skip("libvqhost.so not built")
    syms = subprocess.run(["nm", "-DC", so], capture_output=True, text=True).stdout
    for name in ["vq::VQRenderer::RenderSceneColor", "vq::VQRenderer::RenderPostProcess", "vq::VQRenderer::PreFilterEnvironmentMap",
                 "vq::VQRenderer::LoadDefaultResources", "vq::FEnvironmentMapRenderingResources::CreateRenderingResources",
                 "vq::FPostProcessParameters::FFSR1_EASU::UpdateEASUConstantBlock", "vq::GaussianBlurPass::RecordCommands",
                 "vq::FEnvironmentMapRenderingResources::CreateRenderingResourcesFromHDRFile", "vq::VQRenderer::RenderEnvironmentMap",
                 "vq::VQRenderer::ApplyReflections", "vq::VQRenderer::SaveToHDRFileImage", "vq::ApplyReflectionsPass::RecordCommands", "vq::ParseEnvironmentMapsINI", "vq::CreateEnvironmentMapFileImageFromHiRes"]:
        assert name in syms, name


def test_environment_maps_ini_parser(tmp_path):
    """FileParser::ParseEnvironmentMapsFile (FileParser.cpp:264-321) mirrored in the host layer: sections, Path / MaxCLL keys,
    ';' comments, CRLF, and the engine's rule that a section header only closes the previous entry after an empty line."""
    import shutil
    host = os.path.join(ROOT, "vqengine_b200", "host")
    if not shutil.which("g++") or not os.path.exists(os.path.join(host, "libvqhost.so")):
        pytest.skip("g++ / libvqhost.so not available")
    src = tmp_path / "ini.cpp"
    src.write_text(r'''
#include "vq_renderer.hpp"
#include <cstdio>
int main() {
    const std::string ini = "; comment\n[VondelPark]\nPath=Data/Textures/HDRI/sunny_vondelpark_8k.hdr\nMaxCLL=11987\n\n"
                            "[Stadium01]\nPath=Data/Textures/HDRI/stadium_01_8k.hdr\nMaxCLL=1590\n\n[Last]\r\nPath=x.hdr\r\n"
                            "[NoBlankLineBefore]\nMaxCLL=7\n";
    for (auto& d : vq::ParseEnvironmentMapsINI(ini)) std::printf("%s|%s|%g\n", d.Name.c_str(), d.FilePath.c_str(), d.MaxContentLightLevel);
    return 0;
}''')
    exe = str(tmp_path / "ini")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    subprocess.check_call(["g++", "-std=c++17", str(src), "-I", host, "-I", os.path.join(cuda, "include"), "-L", host, "-lvqhost",
                           "-L", os.path.join(ROOT, "vqengine_b200"), "-lvqcuda", "-L", os.path.join(cuda, "lib64"), "-lcudart",
                           f"-Wl,-rpath,{host}", f"-Wl,-rpath,{os.path.join(ROOT, 'vqengine_b200')}", f"-Wl,-rpath,{os.path.join(cuda, 'lib64')}",
                           "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=60).stdout.splitlines()
    # the last header follows no empty line: as in the engine it renames the entry being read instead of starting a new one
    assert out == ["VondelPark|Data/Textures/HDRI/sunny_vondelpark_8k.hdr|11987", "Stadium01|Data/Textures/HDRI/stadium_01_8k.hdr|1590",
                   "NoBlankLineBefore|x.hdr|7"]
