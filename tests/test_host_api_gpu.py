"""Builds tests/cpp/test_host_api.cpp against include/fcl_b200/fcl.h + libfclb200.so with g++
and runs it: the C++ mirror of the fcl API works end to end through the C ABI."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_host_api(tmp_path):
    exe = str(tmp_path / "test_host_api")
    lib_dir = os.path.join(ROOT, "mind-fcl_b200")
    cmd = ["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_host_api.cpp"),
           "-L", lib_dir, "-lfclb200", f"-Wl,-rpath,{lib_dir}", "-o", exe]
    subprocess.run(cmd, check=True)
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    print(r.stdout)
    assert r.returncode == 0 and "ALL OK" in r.stdout
