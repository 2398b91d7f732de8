"""What the shipped sm_100a code of libnsdg_cuda.so contains (cuobjdump -sass; no GPU needed): the BBM and the parametric strip
kernels stage their plane rows with TMA tensor copies (UTMALDG) completed on mbarriers (SYNCS), the uniform mEVP kernel with
cp.async (LDGSTS); every strip kernel computes in FP64 (DFMA) without local-memory spills (no STL / LDL).  DESIGN 3.2, 3.8."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "nextsimdg_b200", "libnsdg_cuda.so")
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"


@pytest.fixture(scope="module")
def sass():
    if not os.path.exists(LIB) or not os.path.exists(CUOBJDUMP):
        pytest.skip("library or cuobjdump not available")
    out = subprocess.run([CUOBJDUMP, "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, name = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = []
        elif name and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            kernels[name].append(line)
    return kernels


def _ops(lines):
    return [re.sub(r"^@!?U?P\w+\s+", "", re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", l).group(1).strip()).split()[0].split(".")[0] for l in lines]


@pytest.mark.parametrize("kernel,tma", [("subcycle_strip_ubbmI", True), ("subcycle_strip_pbbmI", True), ("subcycle_strip_pmevpI", True),
                                        ("subcycle_strip_umevp1I", True), ("subcycle_strip_umevpI", False)])
def test_strip_kernels_staging_and_spills(sass, kernel, tma):
    found = [k for k in sass if kernel in k]  # mangled names: the template argument list starts with `I`
    assert found, f"{kernel} not in the library"
    for k in found:
        ops = _ops(sass[k])
        assert ops.count("DFMA") > (100 if "umevp1" in kernel else 500), (k, "FP64 arithmetic expected")
        assert "STL" not in ops and "LDL" not in ops, (k, "register spills")
        if tma:
            assert ops.count("UTMALDG") >= 4 and "SYNCS" in ops, (k, "TMA tensor copies + mbarrier expected")
        else:
            assert ops.count("LDGSTS") > 20 and "UTMALDG" not in ops, (k, "cp.async staging expected")
