"""torchrun worker (gloo, CPU) for tests/test_slab_partition.py: exercises the host-side helpers of the
multi-GPU bench without a GPU."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from probabilistic_fluid_simulation_b200 import slab_bench  # noqa: E402
from probabilistic_fluid_simulation_b200.slab import partition  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    # id broadcast: rank 0's 128 bytes arrive everywhere
    payload = bytes(range(128)) if rank == 0 else None
    got = slab_bench.broadcast_bytes(payload, 128, device="cpu")
    assert got == bytes(range(128)), got[:8]
    # timing reduction: the slowest rank decides
    t = slab_bench.max_over_ranks(10.0 + rank, device="cpu")
    assert t == 10.0 + (world - 1)
    # the bands of all ranks tile the grid
    rows = slab_bench.sum_over_ranks(float(partition(rank, world, 1000, 0)[1]), device="cpu")
    assert rows == 1000.0
    # band-wise synthetic inputs equal the corresponding rows of the whole-grid generator
    import bench
    import numpy as np
    full = bench.make_inputs(64, 32)
    r0, nr, _, _ = partition(rank, world, 64, 64)
    band = bench.make_inputs(64, 32, rows=(r0, r0 + nr))
    for a, b in zip(full, band):
        assert np.array_equal(a[r0:r0 + nr], b)
    dist.barrier()
    if rank == 0:
        print("rendezvous ok")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
