"""CPU (-m "not gpu"): the oracle against COMMITTED outputs of the reference's shader text.

tests/golden/shader_golden.npz was written by tests/golden/make_shader_golden.py from oracle/_ref/libhlslref.so — the reference's
Shaders/*.hlsl compiled as C++ in the build container (oracle/Makefile). Unlike tests/test_hlsl_ref.py this file needs neither
/root/reference nor oracle/_ref: the pin travels with the repository. Every case must match bit for bit (NaN marks a pixel the
ENABLE_ALPHA_MASK permutation discards)."""
import os

import numpy as np
import pytest

import shader_cases

GOLDEN = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "shader_golden.npz"))
CASES = shader_cases.cases()


def test_golden_file_and_cases_agree_on_names():
    assert sorted(GOLDEN.files) == sorted(CASES)


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_equals_reference_shader_output(name):
    want = GOLDEN[name]
    got = np.ascontiguousarray(CASES[name][1](), np.float32)
    assert got.shape == want.shape
    same = (got.view(np.uint32) == want.view(np.uint32)) | (np.isnan(got) & np.isnan(want))
    assert same.all(), (name, np.argwhere(~same)[:5])
    assert np.isfinite(want).any()
