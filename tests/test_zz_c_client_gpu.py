"""-m gpu: the plain-C client of the boundary (examples/c_client.c) end to end: .hdr file in -> environment maps ->
forward + skydome + tonemap -> .hdr file out, using nothing but include/vqcuda.h and the CUDA runtime."""
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_c_client_end_to_end(tmp_path, orc):
    from vqengine_b200 import synth
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    if not shutil.which("gcc"):
        pytest.skip("gcc not available")
    exe = str(tmp_path / "c_client")
    subprocess.check_call(["gcc", "-std=c99", "-O1", os.path.join(ROOT, "examples", "c_client.c"), "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(cuda, "include"), "-L", os.path.join(ROOT, "vqengine_b200"), "-lvqcuda",
                           "-L", os.path.join(cuda, "lib64"), "-lcudart", "-lm", "-Wl,-rpath," + os.path.join(ROOT, "vqengine_b200"),
                           "-Wl,-rpath," + os.path.join(cuda, "lib64"), "-o", exe])
    env_file, out_file = str(tmp_path / "env.hdr"), str(tmp_path / "out.hdr")
    open(env_file, "wb").write(orc.hdr_encode(synth.hdri(512, 256)))
    r = subprocess.run([exe, env_file, out_file], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, (r.returncode, r.stderr)
    rc, img, lum = orc.hdr_decode(open(out_file, "rb").read())
    assert rc == 0 and img.shape == (360, 640, 4) and np.isfinite(img).all()
    centre, corner = img[180, 320, :3], img[5, 5, :3]
    assert centre.max() > 0.0 and corner.max() > 0.0            # the lit sphere in the middle, sky in the corner
    # the corner is pure sky: the skydome wrote the HDRI seen along that ray; RGBE keeps it within 1/128
    assert lum > 0.0
