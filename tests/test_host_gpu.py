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
                 "vq::VQRenderer::ApplyReflections", "vq::VQRenderer::SaveToHDRFileImage", "vq::ApplyReflectionsPass::RecordCommands"]:
        assert name in syms, name
