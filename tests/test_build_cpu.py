"""Build-time resource checks (CPU): the register budgets the kernels' occupancy figures in DESIGN.md rest on, read from the
ptxas logs the in-tree build leaves under dedflow_b200/_obj/.  A register or two more can drop a kernel a whole CTA per SM
without any test noticing (k_spmv_fs: 82 registers are allocated as 88 -> 2 CTAs instead of 3, 13 % slower)."""
import re
from pathlib import Path

import pytest

OBJ = Path(__file__).resolve().parents[1] / "dedflow_b200" / "_obj"


def kernels(log):
    """{mangled entry name: (registers, spill bytes)} from a `-Xptxas -v` log"""
    out, name, spill = {}, None, 0
    for line in log.read_text().splitlines():
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            name, spill = m.group(1), 0
        m = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m:
            spill = int(m.group(1)) + int(m.group(2))
        m = re.search(r"Used (\d+) registers", line)
        if m and name:
            out[name] = (int(m.group(1)), spill)
    return out


@pytest.mark.parametrize("log,pattern,max_regs,max_spill", [
    ("solve.ptxas.log", "k_spmv_fs", 80, 0),      # 3 CTAs of 256 threads per SM
    ("solve.ptxas.log", "k_multidot", 64, 0),     # 4 CTAs of 256 threads per SM
    ("solve.ptxas.log", "k_update", 64, 8),       # (one double of the scalar tail)
    ("assemble.ptxas.log", "k_pairJILi1ELi96ELi4E", 168, 0),   # 4 CTAs of 96 threads per SM
    ("assemble.ptxas.log", "k_patchFILi2E", 255, 0),           # 2 CTAs of 128 threads per SM
])
def test_register_budget(log, pattern, max_regs, max_spill):
    from dedflow_b200 import _build
    _build.build()
    path = OBJ / log
    if not path.exists():
        pytest.skip(f"{path} not written by this build")
    ks = {k: v for k, v in kernels(path).items() if pattern in k}
    assert ks, pattern
    for name, (regs, spill) in ks.items():
        assert regs <= max_regs and spill <= max_spill, (name, regs, spill)
