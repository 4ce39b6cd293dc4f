"""The drop-in boundary from a plain C host: compile tests/c_abi/example.c against include/pycusdr_b200.h, link the shared
library and run it.  On the CPU-only box it must report "no CUDA device" (exit 3) -- the library has no fallback."""
import os
import shutil
import subprocess

import pytest

from tests.helpers import have_gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "pycusdr_b200")


def _build(tmp_path, source="example.c"):
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    exe = str(tmp_path / ("c_abi_" + source[:-2]))
    subprocess.run(["gcc", "-std=c99", "-D_GNU_SOURCE", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c_abi", source), "-L", LIBDIR, "-lpycusdr_b200", "-lm",
                    f"-Wl,-rpath,{LIBDIR}", "-o", exe], check=True)
    return exe


def test_c_host_of_the_entry_points_that_need_no_gpu(tmp_path):
    """Bit post-processing with the carry handed between stitchers, window mean, gap filling, sync search -- from plain C."""
    r = subprocess.run([_build(tmp_path, "host_only.c")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "c abi host-only ok" in r.stdout


@pytest.mark.skipif(have_gpu(), reason="CPU-box behaviour")
def test_c_host_links_and_fails_loudly_without_a_gpu(tmp_path):
    r = subprocess.run([_build(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 3, r.stdout + r.stderr
    assert "no CUDA device" in r.stdout


@pytest.mark.gpu
def test_c_host_runs_the_hot_path(tmp_path):
    r = subprocess.run([_build(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "c abi example ok" in r.stdout
