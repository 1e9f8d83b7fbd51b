"""The C++ host mirror (include/parry_b200.hpp) end to end: tests/cpp/host_mirror_kats.cpp runs the reference's exact pins
(epa3.rs:8-23, ball_ball_toi.rs) and hand-checkable ray / pair / intersect_aabb cases through the C++ types, linked against the
in-tree C-ABI library with a plain host compiler. Without a CUDA device the program must fail loudly (no CPU fallback)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_kats(tmp_path):
    from parry_b200 import _ffi
    exe = str(tmp_path / "host_mirror_kats")
    libdir = os.path.dirname(_ffi.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "host_mirror_kats.cpp"),
                           "-o", exe, "-L", libdir, "-lparry_b200", "-Wl,-rpath," + libdir])
    return exe


def test_cpp_mirror_links_and_fails_loudly_without_a_gpu(tmp_path):
    import torch
    exe = build_kats(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present: covered by the gpu test")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 3 and "no usable CUDA device" in r.stdout


@pytest.mark.gpu
def test_cpp_mirror_known_answers(tmp_path):
    exe = build_kats(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "known answers: ok" in r.stdout
