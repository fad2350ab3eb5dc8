"""CPU check of the register-FFT building blocks: fb_fft.cuh compiled for the host."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not available")
def test_host_fft_blocks(tmp_path):
    exe = str(tmp_path / "host_fft_check")
    src = os.path.join(ROOT, "tests", "host", "host_fft_check.cu")
    subprocess.check_call(["nvcc", "-std=c++17", "-O1", "-o", exe, src])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "OK" in out.stdout
