"""CPU: the C++ facade + miniapp compile against the in-tree library with the reference's names, and the binary
refuses to run without a GPU (no CPU fallback)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_miniapp_compiles_and_refuses_without_gpu(tmp_path):
    exe = tmp_path / "conflux_miniapp"
    lib = os.path.join(ROOT, "conflux_b200")
    if not os.path.exists(os.path.join(lib, "libconflux_b200.so")):
        pytest.skip("library not built")
    subprocess.check_call(["g++", "-std=c++17", "-O1", f"-I{ROOT}/include", f"{ROOT}/examples/conflux_miniapp.cpp", "-o",
                           str(exe), f"-L{lib}", "-lconflux_b200", f"-Wl,-rpath,{lib}", "-lpthread"])
    out = subprocess.run([str(exe), "-h"], capture_output=True, text=True)
    assert out.returncode == 0 and "-N" in out.stdout
    import ctypes
    n = ctypes.c_int()
    ctypes.CDLL(os.path.join(lib, "libconflux_b200.so")).cflx_device_count(ctypes.byref(n))
    if n.value == 0:
        out = subprocess.run([str(exe), "-N", "256", "-b", "32", "-r", "1"], capture_output=True, text=True)
        assert out.returncode != 0 and "no CPU fallback" in out.stderr
