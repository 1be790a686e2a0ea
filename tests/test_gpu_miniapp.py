"""GPU: the C++ miniapp (examples/conflux_miniapp.cpp on include/conflux/lu/conflux_b200.hpp) runs on the device with
the reference's flags and prints the reference's `_result_` line (examples/conflux_miniapp.cpp:150-163 there) and, with
--validate, the reference's "Total Frobenius norm" line (conflux_miniapp.cpp:494-500)."""
import os
import re
import subprocess

import pytest

from tests._harness import n_gpus

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = tmp_path_factory.mktemp("miniapp") / "conflux_miniapp"
    lib = os.path.join(ROOT, "conflux_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", f"-I{ROOT}/include", f"{ROOT}/examples/conflux_miniapp.cpp", "-o", str(out),
                           f"-L{lib}", "-lconflux_b200", f"-Wl,-rpath,{lib}", "-lpthread"])
    return str(out)


def _run(exe, *args):
    out = subprocess.run([exe, *args], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout


def test_miniapp_result_lines_single_gpu(exe):
    txt = _run(exe, "-N", "2048", "-b", "128", "-r", "3", "--validate")
    assert "Rank: 0, M: 2048, N: 2048, P:1, v:128, Px:1, Py: 1, Pz: 1, Nt: 16, tA11x: 16, tA11y: 16" in txt
    res = re.findall(r"^_result_ lu,conflux,2048,2048,1,1x1x1,time,other,(\d+),128$", txt, flags=re.M)
    assert len(res) == 3, txt                                    # one line per non-warm-up repetition
    m = re.search(r"^Total Frobenius norm = ([0-9.]+)$", txt, flags=re.M)
    rel = re.search(r"^Relative residual \|\|PA-LU\|\|_F/\|\|A\|\|_F = ([0-9.e+-]+)$", txt, flags=re.M)
    assert m and rel and float(rel.group(1)) <= 1e-12 and float(m.group(1)) <= 1e-6


def test_miniapp_weak_type_and_padding(exe):
    txt = _run(exe, "-N", "1000", "-b", "256", "-r", "1", "-t", "weak")   # padded to 1024 (lu_params.hpp:67-71)
    assert re.search(r"^_result_ lu,conflux,1024,1024,1,1x1x1,time,weak,\d+,256$", txt, flags=re.M), txt


def test_miniapp_multi_gpu_grid(exe):
    if n_gpus() < 4:
        pytest.skip("needs 4 GPUs")
    txt = _run(exe, "-N", "2048", "-b", "128", "-r", "2", "-p", "2,2,1", "--validate")
    assert len(re.findall(r"^_result_ lu,conflux,2048,2048,4,2x2x1,time,other,\d+,128$", txt, flags=re.M)) == 2, txt
    rel = re.search(r"^Relative residual .* = ([0-9.e+-]+)$", txt, flags=re.M)
    assert rel and float(rel.group(1)) <= 1e-12
