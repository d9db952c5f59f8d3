"""Generates tests/golden/shader_golden.npz: the outputs of the REFERENCE's shader text (Shaders/*.hlsl of /root/reference compiled
as C++ into oracle/_ref/libhlslref.so, see oracle/Makefile) on the cases of tests/shader_cases.py. Needs oracle/_ref built (i.e.
/root/reference present at build time). The committed file lets tests/test_shader_golden.py pin the oracle to the reference's
shaders on machines that have neither the reference nor oracle/_ref.

    python tests/golden/make_shader_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(os.path.dirname(HERE)), os.path.dirname(HERE)]
import oracle_lib as orc
import shader_cases

if orc.hlsl_ref() is None:
    sys.exit("oracle/_ref/libhlslref.so is not built (needs /root/reference): run `make -C oracle`")
data = {name: np.ascontiguousarray(ref(), np.float32) for name, (ref, _) in shader_cases.cases().items()}
path = os.path.join(HERE, "shader_golden.npz")
np.savez_compressed(path, **data)
print(path, os.path.getsize(path), "bytes;", {k: v.shape for k, v in data.items()})
