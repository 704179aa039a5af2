"""torchrun worker for tests/test_gpu_slabs.py::test_ring_of_processes: one process per GPU, NCCL halo
exchange, result gathered on rank 0 and compared with the CPU oracle bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from probabilistic_fluid_simulation_b200 import fixtures, vp_field  # noqa: E402
from probabilistic_fluid_simulation_b200.slab import SlabRank  # noqa: E402


# (grid h, w, image h, w, dt, steps): the first case runs long enough for the resident ring to capture its sweep graphs and
# replay them (the plane roles alternate with period 2, a key is captured the second time it turns up); the second case has bands of different heights and a 25-row gather halo, the
# third a gather deeper than the peer transport's fixed halos (it must fall back to send/recv for that exchange),
# the fourth a displacement larger than a band (whole-field path)
CASES = [(128, 256, 192, 256, 2.0, 6), (131, 64, 131, 64, 3000.0, 2), (400, 32, 400, 32, 32000.0, 2),
         (96, 64, 96, 64, 20000.0, 1)]


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl")
    for case in CASES:
        run_case(rank, world, *case)
    run_jump_case(rank, world)
    run_adaptive_case(rank, world)
    dist.barrier()
    dist.destroy_process_group()


def connect(rank, world, w, h, iw, ih):
    slab = SlabRank(rank, world, w, h, iw, ih)
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(SlabRank.unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    slab.connect(bytes(uid.cpu().numpy().tobytes()))
    return slab


def run_jump_case(rank, world):
    """The caller replaces the velocities by ones 100x larger between two steps: the guessed gather depth of the third
    step is too shallow, the step must rerun itself with the measured bound (tests/test_gpu_slabs.py has the
    single-process twin of this case)."""
    h, w = 256, 64
    vp, vtmp, _, _ = fixtures.make_state(fixtures.smooth_velocity_bytes(h, w), fixtures.random_image_bytes(8, 8, 1))
    big = vp[..., :2].copy()
    vp[..., :2] *= np.float32(0.01)
    dt, visc, nd, npr = 4000.0, 0.001, 4, 4
    slab = connect(rank, world, w, h, 0, 0)
    r0, rows = slab.row0, slab.rows
    fv, ft = vp_field(torch.from_numpy(vp[r0:r0 + rows].copy()).cuda()), vp_field(torch.from_numpy(vtmp[r0:r0 + rows].copy()).cuda())
    slab.simulate_fluid_step(fv, ft, dt, visc, nd, npr)
    slab.simulate_fluid_step(fv, ft, dt, visc, nd, npr)
    fv.data[..., :2] = torch.from_numpy(big[r0:r0 + rows].copy()).cuda()
    slab.simulate_fluid_step(fv, ft, dt, visc, nd, npr)
    slab.check()
    parts = [None] * world
    dist.all_gather_object(parts, (fv.data.cpu().numpy(), ft.data.cpu().numpy()))
    if rank == 0:
        orc = oracle.Oracle(nd, npr)
        wv, wt = orc.simulate_fluid_step(vp, vtmp, dt, visc)
        wv, wt = orc.simulate_fluid_step(wv, wt, dt, visc)
        wv = wv.copy()
        wv[..., :2] = big
        wv, wt = orc.simulate_fluid_step(wv, wt, dt, visc)
        got = [np.concatenate([p[k] for p in parts], axis=0) for k in range(2)]
        assert np.array_equal(got[0].view(np.uint32), wv.view(np.uint32)), "vp after the jump"
        assert np.array_equal(got[1].view(np.uint32), wt.view(np.uint32)), "vtmp after the jump"
        print("jump case matches oracle", flush=True)
    dist.barrier()
    slab.close()


def run_adaptive_case(rank, world):
    """pfs_slab_compute_pressure_adaptive across processes: every rank stops at the same count (the rms is all-reduced after
    every batch), and the bands hold the reference's computePressure at that count."""
    h, w, dt, every = 160, 96, 0.37, 8
    rng = np.random.default_rng(7)
    y, x = np.mgrid[0:h, 0:w]
    a = np.zeros((h, w, 4), np.float32)
    a[..., 0] = np.sin(2 * np.pi * x / w) * np.cos(2 * np.pi * y / h) + 0.05 * rng.standard_normal((h, w))
    a[..., 1] = np.cos(4 * np.pi * x / w) * np.sin(2 * np.pi * y / h) + 0.05 * rng.standard_normal((h, w))
    a[..., 2:] = rng.standard_normal((h, w, 2)).astype(np.float32)
    b = np.ascontiguousarray(a[::-1])

    def rms_at(n):
        ra, rb = oracle.Oracle().compute_pressure(a.copy(), b.copy(), dt, n)
        d = rb[..., 2].astype(np.float64) - ra[..., 2].astype(np.float64)
        return ra, rb, float(np.sqrt(np.mean(d * d)))
    tol = 0.5 * (rms_at(3 * every)[2] + rms_at(2 * every)[2])
    slab = connect(rank, world, w, h, 0, 0)
    r0, rows = slab.row0, slab.rows
    fa, fb = vp_field(torch.from_numpy(a[r0:r0 + rows].copy()).cuda()), vp_field(torch.from_numpy(b[r0:r0 + rows].copy()).cuda())
    n, rms = slab.compute_pressure_adaptive(fa, fb, dt, tol, 200, every)
    slab.check()
    parts = [None] * world
    dist.all_gather_object(parts, (n, rms, fa.data.cpu().numpy(), fb.data.cpu().numpy()))
    assert all(p[0] == n and p[1] == rms for p in parts), [(p[0], p[1]) for p in parts]
    if rank == 0:
        assert n == 3 * every, n
        ra, rb, want = rms_at(n)
        assert abs(rms - want) <= 1e-12 * max(1.0, want)
        got = [np.concatenate([p[k] for p in parts], axis=0) for k in (2, 3)]
        assert np.array_equal(got[0].view(np.uint32), ra.view(np.uint32)), "adaptive vp"
        assert np.array_equal(got[1].view(np.uint32), rb.view(np.uint32)), "adaptive vp_out"
        print("adaptive case matches oracle", flush=True)
    dist.barrier()
    slab.close()


def run_case(rank, world, h, w, ih, iw, dt, steps):
    vel = fixtures.smooth_velocity_bytes(h, w)
    img = fixtures.random_image_bytes(ih, iw, 11)
    vp, vtmp, image, itmp = fixtures.make_state(vel, img)
    slab = SlabRank(rank, world, w, h, iw, ih)
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(SlabRank.unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    slab.connect(bytes(uid.cpu().numpy().tobytes()))
    if rank == 0:
        print("transport", slab.transport, flush=True)
    r0, rows, i0, irows = slab.row0, slab.rows, slab.irow0, slab.irows
    fv, ft = vp_field(torch.from_numpy(vp[r0:r0 + rows].copy()).cuda()), vp_field(torch.from_numpy(vtmp[r0:r0 + rows].copy()).cuda())
    fi, fm = vp_field(torch.from_numpy(image[i0:i0 + irows].copy()).cuda()), vp_field(torch.from_numpy(itmp[i0:i0 + irows].copy()).cuda())
    visc, nd, npr = 0.002, 30, 30
    for _ in range(steps):
        slab.simulate_fluid_step(fv, ft, dt, visc, nd, npr)
        slab.advect_color_step(fi, fm, fv, dt)
    slab.check()
    # the same run with the state resident in the slab's planes (pfs_slab_upload / _step / _download)
    rv, rt = torch.from_numpy(vp[r0:r0 + rows].copy()).cuda(), torch.from_numpy(vtmp[r0:r0 + rows].copy()).cuda()
    ri = torch.from_numpy(image[i0:i0 + irows].copy()).cuda()
    slab.upload(rv, rt, ri)
    slab.step(steps, dt, visc, nd, npr)
    slab.download(rv, rt, ri)
    slab.check()
    assert torch.equal(rv.view(torch.int32), fv.data.view(torch.int32)), "resident vp != stateless vp"
    assert torch.equal(rt.view(torch.int32), ft.data.view(torch.int32)), "resident vtmp != stateless vtmp"
    assert torch.equal(ri.view(torch.int32), fi.data.view(torch.int32)), "resident image != stateless image"
    norms = slab.step_norms(fv, ft)          # all-reduced over the ring: identical on every rank
    parts = [None] * world
    dist.all_gather_object(parts, (fv.data.cpu().numpy(), ft.data.cpu().numpy(), fi.data.cpu().numpy()))
    if rank == 0:
        got = [np.concatenate([p[k] for p in parts], axis=0) for k in range(3)]
        want = oracle.Oracle(nd, npr).run_steps(vp, vtmp, image, itmp, dt, visc, steps)
        for name, g, wv in zip(("vp", "vtmp", "image"), got, want):
            assert np.array_equal(g.view(np.uint32), wv.view(np.uint32)), name
        d = want[0][..., 3].astype(np.float64)
        r = want[1][..., 2].astype(np.float64) - want[0][..., 2].astype(np.float64)
        assert abs(norms["div_l2"] - np.sqrt((d * d).sum())) <= 1e-11 * np.sqrt((d * d).sum())
        assert abs(norms["pressure_update_l2"] - np.sqrt((r * r).sum())) <= 1e-11 * np.sqrt((r * r).sum())
        assert norms["speed_max"] == float(max(np.abs(want[0][..., 0]).max(), np.abs(want[0][..., 1]).max()))
        print(f"ring matches oracle {h}x{w}", flush=True)
    all_norms = [None] * world
    dist.all_gather_object(all_norms, norms)
    assert all(n == all_norms[0] for n in all_norms), all_norms
    dist.barrier()
    slab.close()


if __name__ == "__main__":
    main()
