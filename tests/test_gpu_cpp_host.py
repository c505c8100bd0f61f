"""GPU test of the C++ host mirror (include/bdeflate.hpp): a port of the reference's
tests/batch_test.rs compiled with g++ against libbdeflate.so."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_batch_tests_through_cpp_mirror(tmp_path):
    exe = str(tmp_path / "batch_test")
    libdir = os.path.join(ROOT, "libdeflate_rsx_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "batch_test.cpp"), "-o", exe,
                           "-L", libdir, "-lbdeflate", f"-Wl,-rpath,{libdir}"])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr + out.stdout
    assert "all reference batch tests passed" in out.stdout
