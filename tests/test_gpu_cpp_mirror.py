"""Builds and runs the C++ host-mirror test (tests/cpp/test_mirror.cpp over include/bellman_b200.hpp)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_cpp_mirror(tmp_path):
    exe = str(tmp_path / "test_mirror")
    libdir = os.path.join(ROOT, "bellman_mpc_b200")
    subprocess.run(["g++", "-O1", "-std=c++17", os.path.join(ROOT, "tests", "cpp", "test_mirror.cpp"), "-o", exe,
                    "-L" + libdir, "-lbellman_b200", "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "cpp mirror ok" in out.stdout
