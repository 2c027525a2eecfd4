"""The launch chain ("The launch chain"): kernels launched with programmatic stream
serialization must not touch global memory before `griddepcontrol.wait` (SASS: ACQBULK) unless the
design says so.  The compiler is allowed to move read-only-path loads across the wait in either
direction, so this is checked on the SASS of the built library, not on the source."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "jaxsplat_b200", "lib", "libjaxsplat_b200.so")

# kernels that load, ahead of the wait, what kernels OLDER than their predecessor left
EARLY_LOADS = ("rc_row_scatter_kernel", "rc_col_scatter_kernel")
GLOBAL_OP = re.compile(r"\b(LDG|STG|RED|ATOMG|ATOM|LD|ST)\b[.\w]*\s")


def _functions():
    text = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    for blk in re.split(r"\n\s*Function : ", text)[1:]:
        name, _, body = blk.partition("\n")
        yield name.strip(), [l for l in body.split("\n") if re.search(r"/\*[0-9a-f]{4}\*/", l)]


@pytest.mark.skipif(shutil.which("cuobjdump") is None or not os.path.exists(LIB),
                    reason="needs cuobjdump and the built library")
def test_no_global_access_ahead_of_the_wait_on_the_predecessor():
    chained = 0
    for name, lines in _functions():
        if not any("ACQBULK" in l for l in lines):
            continue
        chained += 1
        ahead = []
        for l in lines:
            if "ACQBULK" in l:
                break
            if GLOBAL_OP.search(l):
                ahead.append(l.strip())
        stores = [l for l in ahead if not re.search(r"\bLDG?\b", l)]
        assert not stores, (name, stores)
        if not any(k in name for k in EARLY_LOADS):
            assert not ahead, (name, ahead)
    # projection fwd/bwd, eight list kernels, both tuned rasterisers, the loss
    assert chained >= 13


@pytest.mark.skipif(shutil.which("cuobjdump") is None or not os.path.exists(LIB),
                    reason="needs cuobjdump and the built library")
def test_early_loads_are_issued_ahead_of_the_wait():
    """the point of the early loads: they must sit in front of the (last) wait in the SASS"""
    seen = set()
    for name, lines in _functions():
        for k in EARLY_LOADS + ("project_bwd_kernel",):
            if k in name:
                idx = [i for i, l in enumerate(lines) if "ACQBULK" in l]
                assert idx, name
                loads_before = sum(1 for l in lines[:idx[-1]] if re.search(r"\bLDG\b", l))
                assert loads_before >= 8, (name, loads_before)
                seen.add(k)
    assert seen == set(EARLY_LOADS + ("project_bwd_kernel",))
