"""Static check of the built library (no GPU needed): the store instructions of the packed diffusion kernels.

Why: the rows of the fast path go out with one 256-bit store per lane (inline PTX `st.global.v8.f32` -> STG.E.ENL2.256).
ptxas 12.9 was seen to turn that same inline-asm store into a single 32-bit STG inside the out-of-line exact-division
path (only the first float of each lane's row reached memory; found with PFS_DIFFUSE_FORCE_REPAIR=1).  The exact path
therefore uses two 16-byte stores, and this test pins what the compiler actually emitted for every instantiation:
  * every diffuse_packed_kernel<T, ..., LAST> has 256-bit stores (>= 4 per unrolled trip, >= 8 with the iterate n-1 store),
  * none of them contains a 32-bit global store."""
import collections
import os
import re
import shutil
import subprocess

import pytest

from probabilistic_fluid_simulation_b200 import _cabi


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")
def test_packed_diffusion_kernels_store_whole_rows():
    assert os.path.exists(_cabi.LIB_PATH), "libpfs_b200.so not built"
    out = subprocess.run(["cuobjdump", "-sass", "-fun", "diffuse_packed_kernel", _cabi.LIB_PATH], capture_output=True, text=True)
    text = out.stdout
    if "Function :" not in text:                  # older cuobjdump: no -fun filter by substring
        text = subprocess.run(["cuobjdump", "-sass", _cabi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    cur, counts = None, collections.defaultdict(collections.Counter)
    for line in text.splitlines():
        if "Function :" in line:
            name = line.split("Function :")[1].strip()
            cur = name if "diffuse_packed_kernel" in name else None
            continue
        if cur:
            m = re.search(r"\b(STG\.E[.\w]*|ST\.E[.\w]*)\b", line)
            if m:
                counts[cur][m.group(1)] += 1
    assert len(counts) == 20, sorted(counts)       # depths 2..6 x {2-, 3-instruction division} x {inner pass, last pass}
    for name, c in counts.items():
        last = re.search(r"diffuse_packed_kernelILi\dELi\dELb[01]ELb1E", name) is not None
        wide = sum(n for op, n in c.items() if op.endswith(".256"))
        narrow32 = sum(n for op, n in c.items() if op in ("STG.E", "ST.E"))
        assert wide >= (8 if last else 4), (name, dict(c))
        assert narrow32 == 0, (name, dict(c))
        assert any(op.endswith(".128") for op in c), (name, dict(c))     # the exact path's two 16-byte stores per row
