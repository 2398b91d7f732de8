"""Every reference citation `File.ext:line[-line]` in the headers, kernels, host code, oracle and docs points at a file of
the reference tree that has at least that many lines.  Runs where /root/reference is mounted (the build container);
skipped elsewhere (the GPU box has no reference tree)."""
import glob
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("NSDG_REFERENCE_ROOT", "/root/reference")
CITE = re.compile(r"(?<![\w/.])((?:[\w.+-]+/)*[A-Za-z_][\w+-]*\.(?:cpp|hpp|py|cfg|cdl|txt|smesh|yaml)):(\d+)(?:[-–](\d+))?")
OWN = {"bench.py", "dynamics.py", "partition.py", "synthetic.py", "capi.py", "summarize.py", "quickbench.py"}  # this repo's files


def sources():
    pats = ["include/*.h", "nextsimdg_b200/csrc/*", "nextsimdg_b200/host/*.?pp", "nextsimdg_b200/*.py", "oracle/*.hpp", "oracle/*.cpp",
            "oracle/*.py", "DESIGN.md", "INTEGRATION.md", "README.md", "bench.py", "tests/*.py"]
    out = []
    for p in pats:
        out += sorted(glob.glob(os.path.join(ROOT, p)))
    return [f for f in out if os.path.isfile(f)]


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "dynamics", "src")), reason="reference tree not mounted")
def test_reference_citations_resolve():
    index = {}
    for dirpath, _, files in os.walk(REF):
        if "/.git" in dirpath:
            continue
        for f in files:
            index.setdefault(f, []).append(os.path.join(dirpath, f))
    nlines = {}

    def lines_of(path):
        if path not in nlines:
            with open(path, "rb") as fh:
                nlines[path] = fh.read().count(b"\n") + 1
        return nlines[path]

    bad, checked = [], 0
    for src in sources():
        text = open(src, encoding="utf-8", errors="replace").read()
        for m in CITE.finditer(text):
            path, a, b = m.group(1), int(m.group(2)), int(m.group(3) or m.group(2))
            base = os.path.basename(path)
            if base in OWN or os.path.exists(os.path.join(ROOT, path)):
                continue
            cands = [c for c in index.get(base, []) if c.endswith("/" + path) or "/" not in path]
            checked += 1
            if not cands:
                bad.append(f"{os.path.relpath(src, ROOT)}: {m.group(0)} (no such file in the reference)")
            elif max(lines_of(c) for c in cands) < max(a, b) or b < a:
                bad.append(f"{os.path.relpath(src, ROOT)}: {m.group(0)} (file has {max(lines_of(c) for c in cands)} lines)")
    assert checked > 200, checked
    assert not bad, "\n".join(bad[:40])
