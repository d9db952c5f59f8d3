"""CPU: `bench.py --impl reference` (the arm the driver runs next to ours) prints one well-formed JSON line without a GPU:
the reference's shader text compiled for the CPU when oracle/_ref/libhlslref.so is present (kind "reference"), else the port."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT, env=dict(os.environ, VQ_CPU_ARM_BUDGET_S="3", VQ_CPU_ENV_SMALL="1"))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "forward_pbr_4k_mpixels_per_s" and j["unit"] == "Mpixels/s"
    assert j["higher_is_better"] is True and j["value"] > 0 and j["gpu_launches"] == 0
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = j["cpu_baseline"]
    assert cb["value"] == j["value"] and cb["cores"] >= 1 and cb["kind"] in ("reference", "port") and cb["sample"]
    assert j["product_library_loaded"] is False      # the CPU arm gets its synthetic inputs without loading libvqcuda.so
    import oracle_lib as orc
    if orc.hlsl_ref() is not None:
        assert cb["kind"] == "reference" and "ForwardLighting.hlsl" in cb["sample"]


def test_reference_arm_other_ranks_do_no_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
