"""The C++ host mirror of the reference's class surface (include/emf_b200.hpp: emfb::TSDF / emfb::ObjTSDF over the C ABI).
CPU: the header and its test program compile and link (g++, CUDA runtime headers, libemf_b200.so, the C oracle).
-m gpu: the program runs -- integrate / raycast / gradients bit-exact against the C oracle, object raycast + association,
the device-resident tracker, resize."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "csrc", "_bin", "host_mirror_test")
CUDA = "/usr/local/cuda"


def build():
    from emfusion_b200 import _lib
    from tests import oracle_c
    oracle_c.load()
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    libdir, odir = os.path.dirname(_lib.LIB_PATH), os.path.join(ROOT, "oracle", "_build")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wno-unused-function", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(CUDA, "include"),
           os.path.join(ROOT, "tests", "csrc", "host_mirror_test.cpp"), "-o", BIN, "-L", libdir, "-lemf_b200", "-L", odir,
           "-lemf_oracle", "-L", os.path.join(CUDA, "lib64"), "-lcudart", f"-Wl,-rpath,{libdir}", f"-Wl,-rpath,{odir}",
           f"-Wl,-rpath,{os.path.join(CUDA, 'lib64')}"]
    return subprocess.run(cmd, capture_output=True, text=True)


@pytest.mark.skipif(shutil.which("g++") is None or not os.path.isdir(os.path.join(CUDA, "include")), reason="no g++ / CUDA headers")
def test_host_mirror_compiles_and_links():
    r = build()
    assert r.returncode == 0, r.stderr[-4000:]
    assert "warning" not in r.stderr.lower(), r.stderr[-2000:]
    assert os.path.exists(BIN)


@pytest.mark.gpu
def test_host_mirror_runs(cuda_dev):
    if not os.path.exists(BIN):
        r = build()
        assert r.returncode == 0, r.stderr[-4000:]
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "host mirror ok" in r.stdout, (r.stdout[-3000:], r.stderr[-2000:])
