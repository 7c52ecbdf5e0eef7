"""The boundary from plain C: examples/c_abi_demo.c includes include/smplk.h, links libsmplk.so with gcc (no CUDA
headers, no Python, host buffers only) and checks smplk_forward_host against the same arithmetic in double."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "3d-human-body-reconstruction_b200")


def _build(tmp_path):
    exe = str(tmp_path / "c_abi_demo")
    subprocess.check_call(["gcc", "-O2", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "c_abi_demo.c"), "-o", exe, "-L" + LIBDIR, "-lsmplk", "-lm",
                           "-Wl,-rpath," + LIBDIR])
    return exe


def test_c_demo_builds_against_the_header_and_fails_loudly_without_a_gpu(tmp_path):
    import smplk
    import torch
    smplk.load()                                     # makes sure libsmplk.so exists
    exe = _build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu-marked test")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_c_demo_matches_double_precision_arithmetic(tmp_path):
    import smplk
    smplk.load()
    r = subprocess.run([_build(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout, r.stderr)
    m = re.search(r"max \|error\| vs double = ([0-9.eE+-]+) m, (\d+) kernels", r.stdout)
    assert m and float(m.group(1)) <= 1e-5 and int(m.group(2)) >= 2
