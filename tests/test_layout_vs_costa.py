"""CPU (build container only): conflux::conflux_layout of this repo's facade, converted to real COSTA descriptors,
is field-for-field identical to what the REFERENCE's conflux_layout / lu_params::matrix produce (layout.cpp:30-135),
for both overloads, both orderings, several grids.  Skipped where /root/reference is absent (the GPU box)."""
import os
import subprocess
import sysconfig

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("CONFLUX_REFERENCE", "/root/reference")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "libs", "costa", "src")), reason="reference tree absent")
def test_layout_descriptors_match_costa(tmp_path):
    costa = os.path.join(REF, "libs", "costa", "src")
    sp = sysconfig.get_paths()["purelib"]
    blas = [os.path.join(sp, "opencv_python_headless.libs", f) for f in os.listdir(os.path.join(sp, "opencv_python_headless.libs"))
            if f.startswith("libopenblasp")][0]
    exe = tmp_path / "layout_check"
    common = ["g++", "-O1", "-std=c++17", "-fopenmp", "-w", "-DNDEBUG", f"-I{ROOT}/oracle/mpi_stub", f"-I{costa}"]
    o1, o2 = tmp_path / "mine.o", tmp_path / "check.o"
    # this repo's facade (own include dir first; never sees the reference's conflux headers)
    subprocess.check_call(common + ["-DCONFLUX_B200_WITH_COSTA", f"-I{ROOT}/include", "-c", f"{ROOT}/tests/cpp/layout_mine.cpp", "-o", str(o1)])
    # the reference side
    subprocess.check_call(common + [f"-I{REF}/src", "-c", f"{ROOT}/tests/cpp/layout_check.cpp", "-o", str(o2)])
    srcs = [f"{REF}/src/conflux/lu/layout.cpp", f"{costa}/costa/layout.cpp", f"{ROOT}/oracle/mpi_stub/mpi_threads.cpp"] + \
           [f"{costa}/costa/grid2grid/{f}.cpp" for f in ("block", "grid2D", "interval", "scalapack_layout", "ranks_reordering")]
    subprocess.check_call(common + [f"-I{REF}/src", str(o1), str(o2)] + srcs +
                          [blas, f"-Wl,-rpath,{os.path.dirname(blas)}", f"-Wl,-rpath,{sp}/scipy.libs", "-lpthread", "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and "layouts identical" in out.stdout, out.stderr[-2000:]
