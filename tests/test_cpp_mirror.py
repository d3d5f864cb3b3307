"""Builds and runs tests/cpp/test_bzip2.cpp: the reference's bzip2 encoder tests restated against the C++ mirror of
its API (bzb200.hpp over the C ABI). The build (g++ + link) is checked on CPU; running it needs the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_bzip2")


def _build():
    from oracle import orc
    orc.build()
    pkg = os.path.join(ROOT, "rust-compression_b200")
    cmd = ["g++", "-std=c++17", "-O2", os.path.join(ROOT, "tests", "cpp", "test_bzip2.cpp"), "-o", EXE,
           f"-L{pkg}", "-lbzb200", f"-Wl,-rpath,{pkg}", f"-L{ROOT}/oracle", "-lorc", f"-Wl,-rpath,{ROOT}/oracle"]
    subprocess.check_call(cmd)


def test_cpp_mirror_builds():
    _build()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_cpp_mirror_runs():
    _build()
    p = subprocess.run([EXE, os.path.join(ROOT, "tests", "golden", "data")], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "all C++ mirror tests passed" in p.stdout
