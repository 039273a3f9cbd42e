"""The C++ mirror of Jalla's interface (include/jalla_b200.hpp) compiled with g++ against libjalla_b200.so:
header-level check on CPU, functional run on the GPU box."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "mirror_test.cpp")
LIBDIR = os.path.join(ROOT, "jalla_b200")


def _build(tmp_path):
    exe = str(tmp_path / "mirror_test")
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), SRC, "-L", LIBDIR, "-ljalla_b200",
                        "-Wl,-rpath," + LIBDIR, "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_cpp_mirror_builds_and_fails_loudly_without_gpu(tmp_path):
    from jalla_b200 import _lib
    exe = _build(tmp_path)
    if _lib.lib.jb_device_count() > 0:
        pytest.skip("GPU present: covered by the gpu-marked run")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 77 and "no GPU" in r.stdout


def _build_c(tmp_path):
    exe = str(tmp_path / "solve_sequence")
    r = subprocess.run(["gcc", "-std=c11", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "solve_sequence.c"),
                        "-L", LIBDIR, "-ljalla_b200", "-lm", "-Wl,-rpath," + LIBDIR, "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_c_solve_sequence_links_and_fails_loudly_without_gpu(tmp_path):
    """The header is valid C11 and the link-level symbols Jalla imports (cblas_dcopy, LAPACKE_dgesv) resolve."""
    from jalla_b200 import _lib
    exe = _build_c(tmp_path)
    if _lib.lib.jb_device_count() > 0:
        pytest.skip("GPU present: covered by the gpu-marked run")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 77 and "no GPU" in r.stdout


@pytest.mark.gpu
def test_c_solve_sequence_on_gpu(tmp_path):
    r = subprocess.run([_build_c(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "solve sequence ok" in r.stdout


@pytest.mark.gpu
def test_cpp_mirror_on_gpu(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "cpp mirror ok" in r.stdout
