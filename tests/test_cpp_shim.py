"""C++ host shim (include/hpdg_b200.hh): compiles on CPU; the driver written like the reference's tests runs on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_shim")


def build_driver(orc, hp):
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "test_shim.cc"), "-o", EXE,
           hp.LIB_PATH, os.path.join(ROOT, "oracle", "libhpdg_oracle.so"),
           f"-Wl,-rpath,{os.path.dirname(hp.LIB_PATH)}", f"-Wl,-rpath,{os.path.join(ROOT, 'oracle')}"]
    subprocess.check_call(cmd)


def test_cpp_shim_compiles_and_links(orc, hp):
    build_driver(orc, hp)
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_cpp_shim_runs(orc, hp):
    build_driver(orc, hp)
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=600)
    assert "CPP_SHIM PASS" in out.stdout, out.stdout + out.stderr
