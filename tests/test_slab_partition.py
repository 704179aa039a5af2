"""Host logic of the multi-GPU path (no GPU needed): the slab partition, and -- with two gloo processes --
the rendezvous helpers bench.py uses (id broadcast, max-over-ranks timing)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from probabilistic_fluid_simulation_b200.slab import partition

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("gh,ih", [(64, 64), (100, 150), (256, 768), (512, 512), (16384, 16384), (96, 37), (72, 96)])
@pytest.mark.parametrize("nranks", [1, 2, 3, 4, 8])
def test_partition_covers_grid_and_image(gh, ih, nranks):
    if gh // nranks < 8:
        pytest.skip("slab too thin")
    vih = np.float32(gh) / np.float32(ih)
    next_row, covered = 0, np.zeros(ih, dtype=int)
    for r in range(nranks):
        row0, rows, irow0, irows = partition(r, nranks, gh, ih)
        assert row0 == next_row and rows >= gh // nranks
        next_row = row0 + rows
        covered[irow0:irow0 + irows] += 1
        # every image row of the band looks its velocity up inside the band (fluid.cpp:90)
        j = np.arange(irow0, irow0 + irows, dtype=np.float32)
        vj = (j * vih).astype(np.int32)
        assert ((vj >= row0) & (vj < row0 + rows)).all()
    assert next_row == gh
    assert (covered == 1).all()


def test_partition_rejects_bad_arguments():
    from probabilistic_fluid_simulation_b200 import PfsError
    with pytest.raises(PfsError):
        partition(2, 2, 64, 64)
    with pytest.raises(PfsError):
        partition(0, 0, 64, 64)


def test_gloo_two_process_rendezvous(tmp_path):
    """world_size 2 over gloo on CPU: the id broadcast and the max-over-ranks reduction of slab_bench."""
    script = os.path.join(ROOT, "tests", "gloo_rendezvous_check.py")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", script]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "rendezvous ok" in r.stdout
