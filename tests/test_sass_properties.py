"""Properties of the compiled sm_100a code that the measured performance rests on, read from the SASS
(`cuobjdump`, CPU only): the Blackwell instructions are there, and the MMA issue loops are the tight form —
`UTCHMMA`s issued back to back from uniform registers, not one elect / broadcast loop per instruction
(that form made the issuing thread the bound, 153 instead of 147 cycles per MMA)."""
import re
import shutil
import subprocess
from pathlib import Path

import pytest

BUILD = Path(__file__).resolve().parent.parent / "speechless_b200" / "csrc" / "build"


def sass(obj: str):
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    path = BUILD / obj
    if not path.exists() or not Path(cuobjdump).exists():
        pytest.skip("needs the built objects (python -c 'import __graft_entry__ as g; g.build()') and cuobjdump")
    text = subprocess.run([cuobjdump, "-sass", str(path)], capture_output=True, text=True, check=True).stdout
    functions, name = {}, None
    for line in text.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            functions[name] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if m and name is not None:
            functions[name].append(m.group(1).strip())
    return functions


def mma_runs(instructions):
    """Lengths of the gaps (in instructions) between consecutive UTCHMMAs."""
    where = [i for i, ins in enumerate(instructions) if "UTCHMMA" in ins]
    return where, [b - a for a, b in zip(where, where[1:])]


@pytest.mark.parametrize("obj,pattern", [
    ("conv_umma.o", r"conv_gemm_kernelILi256ELi0ELb0ELi1ELb0E"),   # forward
    ("conv_umma.o", r"conv_gemm_kernelILi256ELi0ELb1ELi1ELb0E"),   # input gradient
    ("conv_umma.o", r"conv_gemm_kernelILi256ELi2ELb1ELi1ELb0E"),   # split-K input gradient
    ("wgrad_umma.o", r"wgrad_kernelILi256E"),                       # weight gradient
])
def test_mma_issue_loops_are_back_to_back(obj, pattern):
    functions = {name: ins for name, ins in sass(obj).items() if re.search(pattern, name)}
    assert functions, "kernel not found in " + obj
    for name, instructions in functions.items():
        where, gaps = mma_runs(instructions)
        assert len(where) >= 4, name
        # every group of four MMAs of a K-step is issued within a handful of instructions of each other ...
        tight = [g for g in gaps if g <= 4]
        assert len(tight) >= 3 * (len(where) // 4) - 2, (name, gaps)
        # ... and no MMA sits in a per-lane elect / broadcast loop any more
        for i in where:
            window = instructions[max(0, i - 3):i + 3]
            assert not any("R2UR.BROADCAST" in w or "BRA.U.ANY" in w for w in window), (name, window)


def test_blackwell_instructions_are_present():
    conv = [i for ins in sass("conv_umma.o").values() for i in ins]
    wgrad = [i for ins in sass("wgrad_umma.o").values() for i in ins]
    for needle in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR", "SYNCS"):
        assert any(needle in i for i in conv), needle
    assert any("UTMAREDG" in i for i in conv) and any("UTMAREDG" in i for i in wgrad)  # TMA reduce-add epilogues
    assert any("UTCHMMA.2CTA" in i for i in conv)  # the CTA-pair variant
