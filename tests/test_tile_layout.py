"""Row-tile layout contract between linearise_kernel (writes by pixel) and the IRLS kernels (read by position):
``staticfusion_b200/csrc/sf_device.cuh::tile_pos``.  A host program compiled with nvcc, run on the CPU."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not on PATH")
def test_tile_layout_contract(tmp_path):
    exe = str(tmp_path / "tile_layout_check")
    subprocess.check_call(["nvcc", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "staticfusion_b200", "csrc"), "-I", os.path.join(ROOT, "include"),
                           "-o", exe, os.path.join(ROOT, "tests", "host", "tile_layout_check.cu")])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "tile layout ok" in out.stdout, out.stdout + out.stderr
