"""The 1-d-pass forms of the constant tables used by the BBM and parametric strip kernels reproduce the tables
they replace: PSI<DG,3> (dynamics/src/include/codeGenerationDGinGauss.hpp:424-603), iMJwPSI / divS1 / divS2 on the unit square
(dynamics/src/ParametricMap.cpp:225-296) and the Q2 basis in the Gauss points (codeGenerationCGinGauss.hpp:14-407).  Host code
only: the functions are __host__ __device__, compiled here with nvcc for the host."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="nvcc not available")
def test_separable_forms_equal_the_tables(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = tmp_path / "separable_tables"
    subprocess.run([nvcc, "-std=c++20", "--expt-relaxed-constexpr", "-gencode", "arch=compute_100a,code=sm_100a", "-I",
                    os.path.join(ROOT, "nextsimdg_b200", "csrc"), "-I", os.path.join(ROOT, "include"), "-o", str(exe),
                    os.path.join(ROOT, "tests", "host", "separable_tables.cu")], check=True, capture_output=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
