"""bench.py's multi-GPU arm: one process per GPU (torchrun), a ring of row slabs connected by NCCL.

Workload (BASELINE.json configs[3] family): a 16384-column grid of 2048 rows PER GPU (weak scaling; 8 GPUs
is the 16384 x 16384 grid), image of the same size, 100 + 100 sweeps.  Every rank builds only its own
band of the synthetic inputs.  Timing: barrier + synchronize on both sides, CUDA events on every rank,
the MAX over ranks is the step time; value = (cells of the WHOLE grid) * n_pressure / time.
"""
from __future__ import annotations

import json
import os
import time


# ---- rendezvous helpers (backend-agnostic; exercised over gloo by tests/test_slab_partition.py) -----
def _tensor(data, dtype, device):
    import torch
    return torch.tensor(data, dtype=dtype, device=device)


def broadcast_bytes(payload: bytes | None, n: int, device="cuda") -> bytes:
    """Rank 0 passes `payload` (n bytes); every rank returns it."""
    import torch
    import torch.distributed as dist
    buf = torch.zeros(n, dtype=torch.uint8, device=device)
    if dist.get_rank() == 0:
        buf.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(buf, 0)
    return bytes(buf.cpu().numpy().tobytes())


def max_over_ranks(x: float, device="cuda") -> float:
    import torch
    import torch.distributed as dist
    t = _tensor([x], torch.float64, device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x: float, device="cuda") -> float:
    import torch
    import torch.distributed as dist
    t = _tensor([x], torch.float64, device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


# ---- host placement --------------------------------------------------------------------------------
def bind_to_gpu_numa_node(local: int):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, BEFORE any pinned host memory is allocated: the
    default memory policy places pages on the allocating thread's node, so the rank's staging buffers end up next to its
    own PCIe root instead of all eight ranks' buffers on one node (which is what made the 8-GPU end-to-end run host-bound).
    -> {"node": n, "cpus": count} or None if the topology cannot be read (then nothing is changed)."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        pci = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{pci}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return {"node": node, "cpus": len(cpus)}
    except Exception:
        return None


# ---- the bench arm ---------------------------------------------------------------------------------
TRANSPORT_TEXT = {
    "p2p": "halo rows stored straight into the ring neighbours' memory (CUDA IPC mappings over NVLink, one kernel per "
           "exchange with flag handshakes); NCCL only for the all-reduce of the displacement bound (one process per GPU)",
    "nccl": "NCCL send/recv of halo rows between ring neighbours (one process per GPU)",
}


def run(args, bench) -> None:
    import numpy as np
    import torch
    import torch.distributed as dist

    import probabilistic_fluid_simulation_b200 as pfs
    from probabilistic_fluid_simulation_b200.slab import SlabRank

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if world != args.gpus:
        raise SystemExit(f"bench.py --gpus {args.gpus} needs WORLD_SIZE={args.gpus} (launch with torch.distributed.run); "
                         f"got WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local)
    w, h, n = bench.workload_shape(args)
    # NCCL announces its version on stdout when a communicator is created; stdout must carry exactly one
    # JSON line, so fd 1 points at stderr until both communicators (torch's and the library's) exist.
    import sys
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    def make_slab(gw, gh, iw, ih):
        sl = SlabRank(rank, world, gw, gh, iw, ih)
        sl.connect(broadcast_bytes(SlabRank.unique_id() if rank == 0 else None, 128))
        return sl

    parity = None
    try:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        if not getattr(args, "no_parity", False):
            # the ring is verified against the CPU oracle before anything is timed (bench.ring_parity_check)
            parity = bench.ring_parity_check(rank, world, make_slab)
        slab = make_slab(w, h, w, h)
        dist.barrier()
        torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)

    vp, vtmp, image, itmp = bench.make_inputs(h, w, rows=(slab.row0, slab.row0 + slab.rows))
    assert slab.irow0 == slab.row0 and slab.irows == slab.rows        # image res == grid res in this workload
    fv, ft, fi, fm = (pfs.vp_field(torch.from_numpy(x).cuda()) for x in (vp, vtmp, image, itmp))
    DT, VISC = bench.DT, bench.VISC

    # The timed steps run on resident state (pfs_slab_upload once, pfs_slab_step per timestep), as the N = 1 arm runs on a
    # pfs_ctx; --stateless times pfs_slab_simulate_fluid_step + pfs_slab_advect_color_step on the caller-owned bands instead.
    resident = not getattr(args, "stateless", False)
    if resident:
        slab.upload(fv.data, ft.data, fi.data)

    def step():
        if resident:
            slab.step(1, DT, VISC, n, n)
        else:
            slab.simulate_fluid_step(fv, ft, DT, VISC, n, n)
            slab.advect_color_step(fi, fm, fv, DT)

    # setup, as in the N = 1 arm: resident slabs capture their sweep segment once per plane-role assignment (period 2, the
    # second occurrence is captured); run through that before the warm-up so that warm-up and timed region are steady state
    for _ in range(bench.PRIME_STEPS):
        step()
    for _ in range(args.warmup):
        step()
    slab.check()
    torch.cuda.synchronize()
    dist.barrier()

    sampler = bench.ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.12)
    l0 = pfs.kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    dist.barrier()
    t_begin = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    t_end = time.perf_counter()
    ms_local = e0.elapsed_time(e1)
    launches = pfs.kernel_launch_count() - l0
    clocks = sampler.stop(t_begin, t_end) if sampler else None
    # per-phase breakdown: a second pass of the same steps with the library's phase events on (they serialise a little and
    # switch the captured sweep graphs off, so they stay out of the timed region above, as in the N = 1 arm)
    pfs.phase_timing(True)
    pfs.phase_times(reset=True)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    dist.barrier()
    p0.record()
    for _ in range(args.steps):
        step()
    p1.record()
    torch.cuda.synchronize()
    phase_region_ms = p0.elapsed_time(p1) / args.steps
    phase_ms, phase_launches = pfs.phase_times(reset=True)
    pfs.phase_timing(False)
    dist.barrier()
    ms_total = max_over_ranks(ms_local)
    ms_step = ms_total / args.steps
    cells_global = w * h
    cells_local = w * slab.rows
    value = cells_global * n / (ms_step * 1e-3)
    total_launches = int(sum_over_ranks(float(launches)))

    # ---- the weak-scaling unit: the SAME band shape (w x rows-per-GPU) as one periodic grid on the regular 1-GPU path, measured
    #      in this run on rank 0 while the other ranks wait, so that the N-GPU line carries its own denominator ----
    unit = None
    if not getattr(args, "no_unit", False):
        if rank == 0:
            uh = slab.rows
            uvp, uvt, uimg, _ = bench.make_inputs(uh, w)
            ctx = pfs.FluidContext(w, uh, w, uh)
            ctx.upload(torch.from_numpy(uvp).cuda(), torch.from_numpy(uvt).cuda(), torch.from_numpy(uimg).cuda())
            for _ in range(bench.PRIME_STEPS + 3):          # graph priming + warm-up, as in the N = 1 arm
                ctx.step(1, DT, VISC, n, n)
            u_steps = max(3, min(args.steps, 10))
            u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            u0.record()
            for _ in range(u_steps):
                ctx.step(1, DT, VISC, n, n)
            u1.record()
            torch.cuda.synchronize()
            unit_ms = u0.elapsed_time(u1) / u_steps
            ctx.close()
            del ctx
            torch.cuda.empty_cache()
            unit = {"grid": [w, uh], "ms_per_step": unit_ms, "value": w * uh * n / (unit_ms * 1e-3), "unit": bench.UNIT,
                    "steps": u_steps, "what": "one periodic grid of the per-GPU band shape on the regular single-GPU path "
                                              "(persistent context), same run, rank 0"}
        dist.barrier()

    # ---- end to end: every rank uploads its bands from pinned memory, steps, downloads them ----
    # Pinned buffers are allocated after the process was bound to its GPU's NUMA node.  Uploads run on one stream, the
    # kernels on the current one, downloads on a third, so the two PCIe directions overlap where the data dependencies
    # allow: the image goes up while the fluid step runs, vp/vtmp come down while advect_color runs.
    e2e = None
    if not args.no_e2e:
        hv, ht, hi = (torch.from_numpy(x).pin_memory() for x in (vp, vtmp, image))
        e2e_steps = max(2, min(args.steps, 5))
        s_up, s_down = torch.cuda.Stream(), torch.cuda.Stream()
        main = torch.cuda.current_stream()
        ev = [torch.cuda.Event() for _ in range(6)]

        def e2e_step():
            if not resident:
                fv.data.copy_(hv, non_blocking=True); ft.data.copy_(ht, non_blocking=True); fi.data.copy_(hi, non_blocking=True)
                step()
                hv.copy_(fv.data, non_blocking=True); ht.copy_(ft.data, non_blocking=True); hi.copy_(fi.data, non_blocking=True)
                torch.cuda.synchronize()
                return
            s_up.wait_stream(main)
            with torch.cuda.stream(s_up):
                fv.data.copy_(hv, non_blocking=True); ft.data.copy_(ht, non_blocking=True)
                ev[0].record()
                fi.data.copy_(hi, non_blocking=True)
                ev[1].record()
            main.wait_event(ev[0])
            slab.upload(fv.data, ft.data, None)
            slab.step_fluid(DT, VISC, n, n)
            slab.download(fv.data, ft.data, None)
            ev[2].record()
            with torch.cuda.stream(s_down):
                s_down.wait_event(ev[2])
                hv.copy_(fv.data, non_blocking=True); ht.copy_(ft.data, non_blocking=True)
            main.wait_event(ev[1])
            slab.upload(None, None, fi.data)
            slab.step_color(DT)
            slab.download(None, None, fi.data)
            ev[3].record()
            with torch.cuda.stream(s_down):
                s_down.wait_event(ev[3])
                hi.copy_(fi.data, non_blocking=True)
            torch.cuda.synchronize()
        e2e_step()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        dist.barrier()
        e2e_s = max_over_ranks((time.perf_counter() - t0) / e2e_steps)
        e2e = {"value": cells_global * n / e2e_s, "unit": bench.UNIT, "ms_per_step": e2e_s * 1e3,
               "h2d_bytes_per_step": 3 * cells_global * 16, "d2h_bytes_per_step": 3 * cells_global * 16,
               "api": "every rank: pinned host bands -> device (upload stream), pfs_slab_upload + pfs_slab_step_fluid / _color + "
                      "pfs_slab_download, device -> pinned host bands (download stream); image upload overlaps the fluid step, "
                      "velocity download overlaps advect_color", "steps": e2e_steps,
               "host_placement": numa if numa else "NUMA topology not readable: default placement"}

    if rank == 0:
        peak, peak_src = bench.measured_peak_gbs()
        per_phase = {k: v / args.steps for k, v in phase_ms.items()}
        kernels = {}
        for phase, key in (("diffuse", "diffuse_sweep"), ("pressure", "pressure_sweep")):
            nl = phase_launches[phase] / args.steps
            if nl > 0 and per_phase[phase] > 0:
                b = bench.BYTES[key] * cells_local * n
                gbs = b / (per_phase[phase] * 1e-3) / 1e9
                kernels[phase] = {"launches_per_step": nl, "phase_ms": per_phase[phase], "alg_bytes_per_phase": b,
                                  "achieved_gbs": gbs, "frac": gbs / peak, "share_of_step": per_phase[phase] / phase_region_ms,
                                  "note": "rank 0, includes the halo exchanges between passes"}
        dom = max(kernels, key=lambda k: kernels[k]["share_of_step"]) if kernels else None
        roofline = None
        if dom:
            roofline = {"bound": "hbm", "kernel": dom + " phase (fused passes + halo exchanges), per GPU",
                        "achieved": kernels[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": kernels[dom]["frac"],
                        "traffic": None, "peak_source": peak_src,
                        "frac_of_nominal_8000_gbs": kernels[dom]["achieved_gbs"] / 8000.0}
        step_bytes = (88 + 16 * n + 12 * n) * cells_local
        line = {"metric": bench.METRIC, "value": value, "unit": bench.UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": bench.workload_config(args),
                "roofline": roofline, "cpu_baseline": None, "e2e": e2e, "gpu_launches": total_launches, "clocks": clocks,
                "whole_step_roofline_per_gpu": {"alg_bytes": step_bytes,
                                                "achieved_gbs": step_bytes / (ms_step * 1e-3) / 1e9,
                                                "frac": step_bytes / (ms_step * 1e-3) / 1e9 / peak},
                "cell_steps_per_s": cells_global / (ms_step * 1e-3), "phases_ms_rank0": per_phase,
                "phase_region": {"steps": args.steps, "ms_per_step_with_phase_events": phase_region_ms}, "kernels": kernels,
                "api": ("pfs_slab_step on resident state (pfs_slab_upload once)" if resident else
                        "pfs_slab_simulate_fluid_step + pfs_slab_advect_color_step on caller-owned bands (--stateless)"),
                "transport": TRANSPORT_TEXT.get(slab.transport, slab.transport),
                "rows_per_gpu": slab.rows, "parity": parity,
                "weak_unit_1gpu": unit,
                "efficiency_vs_unit": (unit["ms_per_step"] / ms_step) if unit else None}
        print(json.dumps(line), flush=True)
    dist.barrier()
    slab.close()
    dist.destroy_process_group()
    if parity is not None and not parity["bit_identical"]:
        raise SystemExit(3)          # a ring that does not reproduce the oracle must not look like a result
