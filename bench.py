#!/usr/bin/env python
"""bench.py -- the driver-facing benchmark of the fluid-step hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A "step" is one timestep of the reference driver loop (main.cpp:236-239): simulate_fluid_step
(advect, n_diffuse smoothing sweeps, divergence, n_pressure Jacobi sweeps, gradient subtraction)
followed by advect_color_step.  Metric (BASELINE.json): pressure cell-updates per second,
    value = W * H * n_pressure / (seconds per WHOLE timestep)
i.e. the pressure-iteration rate the whole step sustains (every other phase counts against it).
The pressure solve alone and every phase are reported beside it (`phases`, `pressure_solve`).

Workloads:
  N = 1 : BASELINE.json configs[2] -- synthetic 4096x4096 grid and image, 100 + 100 sweeps.
  N > 1 : configs[3] family -- 16384 columns x (2048 * N) rows, row slabs of 16384 x 2048 per GPU
          (N = 8 is the 16384^2 grid); weak scaling, halo rows exchanged between neighbours.
Inputs are far larger than L2 (256 MiB per buffer vs 126 MB), so no explicit flush is needed.

Prints ONE JSON line (rank 0).  See the task contract for the keys; `roofline` describes the
dominant kernel (largest share of the step), `cpu_baseline` the reference's own fluid.cpp compiled
from /root/reference (oracle/_ref) timed on one host core on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "cell-updates/s per pressure iter"
UNIT = "cell-updates/s"
DT, VISC = 0.1, 0.001

# algorithmic bytes per cell (SURVEY.md 8d / DESIGN.md): compulsory planar traffic, no credit for
# temporal blocking
BYTES = {"advect": 16, "diffuse_sweep": 16, "divergence": 12, "pressure_sweep": 12, "project": 20, "advect_color": 40}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons of one GPU sampled DURING the timed region: NVML polled every 5 ms from a
    thread (time-stamped on this host's clock, so samples can be matched to the region exactly); if NVML cannot
    be loaded, `nvidia-smi -lms 50` through a pipe (coarser: its output arrives in bursts)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    # nvmlClocksThrottleReason* / nvmlClocksEventReason* bits (nvml.h)
    BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index: int, period_s: float = 0.005):
        self.index = index
        self.period = period_s
        self.rows = []          # (t, sm_mhz, sm_max_mhz, set(reasons))
        self.proc = None
        self.thread = None
        self.source = None
        self._stop = threading.Event()

    # ---- NVML ----
    def _nvml_handle(self):
        import pynvml
        import torch
        pynvml.nvmlInit()
        try:                                    # CUDA ordinal -> NVML device through the UUID (CUDA_VISIBLE_DEVICES-proof)
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            return pynvml, pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
        except Exception:
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.index)

    def _poll_nvml(self, nv, h):
        try:
            mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        except Exception:
            mx = None
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self._stop.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                mask = int(get_reasons(h))
                t = time.perf_counter()
                self.rows.append((t, sm, mx, {n for n, b in self.BITS.items() if mask & b}))
            except Exception:
                pass
            self._stop.wait(self.period)

    # ---- nvidia-smi fallback ----
    def _read_smi(self):
        for line in self.proc.stdout:
            r = [x.strip() for x in line.split(",")]
            try:
                reasons = {n for n, v in zip(self.NAMES, r[3:7]) if v.lower().startswith("active")}
                self.rows.append((time.perf_counter(), float(r[0]), float(r[1]), reasons))
            except Exception:
                continue

    def start(self):
        try:
            nv, h = self._nvml_handle()
            nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            self.thread = threading.Thread(target=self._poll_nvml, args=(nv, h), daemon=True)
            self.thread.start()
            self.source = f"nvml, every {self.period * 1e3:.0f} ms"
            return
        except Exception:
            pass
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read_smi, daemon=True)
            self.thread.start()
            self.source = "nvidia-smi -lms 50"
        except Exception:
            self.proc = None

    def stop(self, t0=None, t1=None):
        if self.source is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml and nvidia-smi unavailable"]}
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        else:
            self._stop.set()
            self.thread.join(timeout=2)
        slack = 0.06 if self.proc else 0.0
        inside = [r for r in self.rows if t0 is None or (t0 <= r[0] <= t1 + slack)]
        window = "timed region"
        if not inside:                       # region shorter than the sampling period
            inside, window = list(self.rows), "whole run (timed region shorter than one sample)"
        sm = sorted(r[1] for r in inside)
        mx = [r[2] for r in inside if r[2] is not None]
        reasons = set().union(*[r[3] for r in inside]) if inside else set()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window, "source": self.source}


def make_inputs(h: int, w: int, rows=None):
    """Synthetic inputs of the named shape: smooth coherent flow + position-hashed noise of +-6 byte
    levels, position-hashed image.  `rows = (r0, r1)` builds only that band of the h x w grid and image
    (multi-GPU ranks build their own band; it equals the same rows of the whole-grid inputs)."""
    import numpy as np
    from probabilistic_fluid_simulation_b200 import fixtures
    r0, r1 = rows if rows is not None else (0, h)
    vel = fixtures.smooth_velocity_bytes(h, w, rows=(r0, r1))
    noise = fixtures.hash_bytes(h, w, 2, 1234, rows=(r0, r1)).astype(np.int16) % 13 - 6
    vel[..., :2] = np.clip(vel[..., :2].astype(np.int16) + noise, 0, 255).astype(np.uint8)
    img = fixtures.hash_bytes(h, w, 4, 4321, rows=(r0, r1))
    img[..., 3] = 255
    return fixtures.make_state(vel, img)


# --------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own fluid.cpp (oracle/_ref) on one host core
# --------------------------------------------------------------------------------------------
def cpu_reference_rate(n_iters: int, w: int, sample_rows: int, steps: int, warmup: int):
    """Times `steps` timesteps of the unmodified reference on a w x sample_rows grid (same dt,
    viscosity, sweep counts as the GPU workload).  -> (cell-updates/s, ms per step, kind)."""
    import oracle
    kind = "reference"
    if oracle.Reference.available(n_iters):
        impl = oracle.Reference(n_iters)
    else:                                   # compiled reference did not travel: use the C port
        impl = oracle.Oracle(n_iters)
        kind = "port"
    vp, vtmp, image, itmp = make_inputs(sample_rows, w)
    for _ in range(warmup):
        vp, vtmp, image, itmp = impl.run_steps(vp, vtmp, image, itmp, DT, VISC, 1)
    t0 = time.perf_counter()
    for _ in range(steps):
        vp, vtmp, image, itmp = impl.run_steps(vp, vtmp, image, itmp, DT, VISC, 1)
    dt_s = (time.perf_counter() - t0) / steps
    return w * sample_rows * n_iters / dt_s, dt_s * 1e3, kind


def fluid_cu_baseline(n_iters: int, w: int, h: int, steps: int = 5, warmup: int = 2):
    """Times the reference's OWN CUDA backend (src/fluid.cu, written for sm_75, recompiled unmodified for
    sm_100a into oracle/_ref/libfluid_refcu_<N>.so) on the same workload -- a reported baseline only: it is
    not numerically equal to fluid.cpp.  -> dict or None if the library did not travel."""
    import ctypes
    import torch
    path = os.path.join(ROOT, "oracle", "_ref", f"libfluid_refcu_{n_iters}.so")
    if not os.path.exists(path):
        return None
    lib = ctypes.CDLL(path)
    pp = ctypes.POINTER(ctypes.c_void_p)
    lib.refcu_timestep.argtypes = [pp, pp, pp, pp, ctypes.c_float, ctypes.c_float] + [ctypes.c_int] * 4
    lib.refcu_timestep.restype = None
    vp, vtmp, image, itmp = make_inputs(h, w)
    bufs = [torch.from_numpy(x).cuda() for x in (vp, vtmp, image, itmp)]
    ptrs = [ctypes.c_void_p(t.data_ptr()) for t in bufs]

    def step():
        lib.refcu_timestep(*(ctypes.byref(p) for p in ptrs), DT, VISC, w, h, w, h)

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    torch.cuda.synchronize()       # the reference's own driver stops its clock before this (main.cpp:247)
    ms = (time.perf_counter() - t0) / steps * 1e3
    return {"value": w * h * n_iters / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "launches_per_step": 2 * n_iters + 4,
            "what": "reference src/fluid.cu compiled unmodified with nvcc -O3 for sm_100a, same grid and sweep counts; "
                    "timed baseline only (its results differ from fluid.cpp, SURVEY.md 2.2)"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w, h, n = workload_shape(args)
    sample_rows = max(64, (1 << 20) // w)          # ~1 Mcell per step: ~2-3 s of CPU work at N=100
    rate, ms, kind = cpu_reference_rate(n, w, sample_rows, args.steps, args.warmup)
    sample = (f"{w}x{sample_rows} rows of the {w}x{h} workload, {n}+{n} sweeps, dt={DT}, nu={VISC}; "
              f"one timestep per step; fluid.cpp is single-threaded (SURVEY.md 2.1)")
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args),
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": 1, "host_cores": os.cpu_count(), "kind": kind,
                             "sample": sample},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# multi-GPU arm: parity of the ring the timed steps run on (outside the timed region)
# --------------------------------------------------------------------------------------------
def ring_parity_check(rank: int, world: int, make_slab):
    """Runs small whole-grid cases through the SAME one-process-per-GPU ring as the timed workload (same SlabRank
    class, same transport) and compares every band, bit for bit, with the CPU oracle run on the whole grid by rank 0
    (reference semantics: fluid.cpp:298-320).  `make_slab(w, h, iw, ih)` returns a connected SlabRank.
    Grid 1024 x (64 * ranks), image 1536 x (96 * ranks) (image/grid ratio 1.5: the look-up factor of fluid.cpp:82-83 is
    inexact in binary32), 7 + 10 sweeps, 6 steps (so that the resident ring's captured sweep graphs are replayed, not only built); two time steps: one whose departure rows reach ~20 rows into the
    neighbouring bands (peer-written gather halos) and one whose departure rows lie beyond the neighbouring bands
    (whole-field gather).  -> the "parity" object of the JSON line (identical on every rank)."""
    import numpy as np
    import torch
    import torch.distributed as dist

    import probabilistic_fluid_simulation_b200 as pfs
    from probabilistic_fluid_simulation_b200 import fixtures

    w, h, iw, ih = 1024, 64 * world, 1536, 96 * world
    nd, npr, steps, visc = 7, 10, 6, 0.002
    vel = fixtures.smooth_velocity_bytes(h, w)
    noise = fixtures.hash_bytes(h, w, 2, 77).astype(np.int16) % 13 - 6
    vel[..., :2] = np.clip(vel[..., :2].astype(np.int16) + noise, 0, 255).astype(np.uint8)
    img = fixtures.hash_bytes(ih, iw, 4, 78)
    state = fixtures.make_state(vel, img)
    cases = []
    ok_all = True
    transport = None
    for label, dt in (("gather reaches into the neighbouring bands", 20.0 * h), ("gather beyond the neighbouring bands", 150.0 * h)):
        slab = make_slab(w, h, iw, ih)
        transport = slab.transport
        r0, rows, i0, irows = slab.row0, slab.rows, slab.irow0, slab.irows
        vp, vtmp, image, itmp = (x.copy() for x in state)
        fv, ft = (pfs.vp_field(torch.from_numpy(x[r0:r0 + rows].copy()).cuda()) for x in (vp, vtmp))
        fi, fm = (pfs.vp_field(torch.from_numpy(x[i0:i0 + irows].copy()).cuda()) for x in (image, itmp))
        for _ in range(steps):
            slab.simulate_fluid_step(fv, ft, dt, visc, nd, npr)
            slab.advect_color_step(fi, fm, fv, dt)
        slab.check()
        # ... and the same steps on resident state (pfs_slab_upload / _step / _download), the path the timed region uses
        rv, rt, ri = (torch.from_numpy(x[a:a + m].copy()).cuda() for x, a, m in ((vp, r0, rows), (vtmp, r0, rows), (image, i0, irows)))
        slab.upload(rv, rt, ri)
        slab.step(steps, dt, visc, nd, npr)
        slab.download(rv, rt, ri)
        slab.check()
        resident_equal = bool(torch.equal(rv.view(torch.int32), fv.data.view(torch.int32)) and
                              torch.equal(rt.view(torch.int32), ft.data.view(torch.int32)) and
                              torch.equal(ri.view(torch.int32), fi.data.view(torch.int32)))
        mine = {"rank": rank, "rows": (r0, rows), "irows": (i0, irows), "resident_equal": resident_equal,
                "vp": fv.data.cpu().numpy(), "vtmp": ft.data.cpu().numpy(), "image": fi.data.cpu().numpy()}
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        verdict = [True, {}]
        if rank == 0:
            import oracle          # the checker: never on the timed or shipped path
            want = oracle.Oracle(nd, npr).run_steps(*(x.copy() for x in state), np.float32(dt), np.float32(visc), steps)
            bad = [f"rank {part['rank']}: resident state != stateless bands" for part in parts if not part["resident_equal"]]
            for part in parts:
                a, n = part["rows"]
                b, m = part["irows"]
                for name, wfield, sl in (("vp", want[0], slice(a, a + n)), ("vtmp", want[1], slice(a, a + n)),
                                         ("image", want[2], slice(b, b + m))):
                    if not np.array_equal(part[name].view(np.uint32), np.ascontiguousarray(wfield[sl]).view(np.uint32)):
                        bad.append(f"rank {part['rank']} {name}")
            got_vp = np.concatenate([p_["vp"] for p_ in sorted(parts, key=lambda q: q["rank"])], axis=0)
            verdict = [not bad, {"mismatches": bad, "fnv1a_vp_whole_grid": oracle.field_hashes(got_vp),
                                 "fnv1a_vp_oracle": oracle.field_hashes(want[0])}]
        box = [verdict]
        dist.broadcast_object_list(box, src=0)
        ok, detail = box[0]
        ok_all = ok_all and ok
        cases.append({"case": label, "dt": dt, "bit_identical": ok, **detail})
        dist.barrier()
        slab.close()
    return {"ranks": world, "transport": transport, "bit_identical": ok_all, "checker": "oracle.Oracle (C restatement of "
            "fluid.cpp, pinned against the compiled reference) on the whole grid, rank 0; stateless bands and resident state "
            "(the timed path) both compared", "grid": [w, h], "image": [iw, ih],
            "sweeps": [nd, npr], "steps": steps, "cases": cases}


# --------------------------------------------------------------------------------------------
# workload description
# --------------------------------------------------------------------------------------------
PRIME_STEPS = 7          # setup steps before the warm-up: one period of the step-graph role assignments + the eager first step


def workload_shape(args):
    if args.width and args.height:
        return args.width, args.height, args.iters
    if args.gpus == 1:
        return 4096, 4096, args.iters
    return 16384, 2048 * args.gpus, args.iters


def workload_config(args):
    w, h, n = workload_shape(args)
    if args.gpus == 1:
        name = f"synthetic {w}x{h} grid + {w}x{h} image, {n} diffusion + {n} pressure sweeps/step (BASELINE configs[2])"
    else:
        name = (f"synthetic {w}x{h} grid + image, {n}+{n} sweeps/step, row slabs of {w}x{h // args.gpus} per GPU "
                f"(BASELINE configs[3] family; N=8 is 16384x16384)")
    return {"workload": name, "grid": [w, h], "image": [w, h], "n_diffuse": n, "n_pressure": n, "dt": DT,
            "viscosity": VISC, "parallelism": "single GPU" if args.gpus == 1 else f"row slabs x{args.gpus}",
            "l2": "inputs larger than L2 (no flush needed)"}


# --------------------------------------------------------------------------------------------
# B200 arm, one GPU
# --------------------------------------------------------------------------------------------
def run_single_gpu(args):
    import numpy as np
    import torch

    import probabilistic_fluid_simulation_b200 as pfs

    torch.cuda.set_device(0)
    w, h, n = workload_shape(args)
    cells = w * h
    vp, vtmp, image, itmp = make_inputs(h, w)
    fv, ft, fi, fm = (pfs.vp_field(torch.from_numpy(x).cuda()) for x in (vp, vtmp, image, itmp))

    # The timed workload runs on a persistent-state context (pfs_ctx_*): the fields are resident in HBM in the library's
    # planar layout, one pfs_ctx_step per timestep = simulate_fluid_step + advect_color_step (main.cpp:236-239).  The
    # stateless entry points (caller-owned interleaved buffers re-read and re-written every step) are timed beside it.
    ctx = None
    if args.graph:
        args.stateless = True          # an outer graph of two steps needs the caller-owned buffers of the stateless calls
    if not args.stateless:
        ctx = pfs.FluidContext(w, h, w, h)
        ctx.upload(fv.data, ft.data, fi.data)

    def step():
        if ctx is not None:
            ctx.step(1, DT, VISC, n, n)
        else:
            pfs.simulate_fluid_step(fv, ft, DT, VISC, n, n)
            pfs.advect_color_step(fi, fm, fv, DT)

    # One-time setup, like the upload above: the library captures one CUDA graph per plane-role assignment of a step (the three
    # pressure planes rotate with period 3, the image ping-pong with period 2: six assignments, each captured the first time it
    # recurs).  Running through one full period here keeps those captures -- host work that normally hides behind the GPU but
    # can stall it when the host hiccups -- out of the warm-up and the timed region, which then see only steady-state steps.
    for _ in range(PRIME_STEPS):
        step()
    torch.cuda.synchronize()
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()

    graph = None
    if args.graph:
        # Two timesteps per graph: the image/itmp exchange of advect_color_step returns to the captured
        # pointers after an even number of steps (vp/tmp are never exchanged for equal sweep parities).
        if args.steps % 2:
            args.steps += 1
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            with torch.cuda.graph(graph, stream=side):
                step()
                step()
        torch.cuda.synchronize()
        graph.replay()
        torch.cuda.synchronize()

    # ---- timed region: K steps, device-resident inputs, CUDA events on the launching stream ----
    # (the library replays its cached step graph here; per-phase events would force it to launch eagerly,
    #  so the per-phase / per-kernel figures come from a second region right after, see below)
    sampler = ClockSampler(0)
    sampler.start()
    time.sleep(0.05)          # the sampler is running before the (short) timed region starts
    l0 = pfs.kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t_begin = time.perf_counter()
    e0.record()
    if graph is None:
        for _ in range(args.steps):
            step()
    else:
        for _ in range(args.steps // 2):
            graph.replay()
    e1.record()
    torch.cuda.synchronize()
    t_end = time.perf_counter()
    ms_total = e0.elapsed_time(e1)
    launches = pfs.kernel_launch_count() - l0
    if graph is not None:       # an outer torch graph replays without going through the library's counter
        l1 = pfs.kernel_launch_count()
        step(); step()
        torch.cuda.synchronize()
        launches = (pfs.kernel_launch_count() - l1) * (args.steps // 2)
    clocks = sampler.stop(t_begin, t_end)

    # ---- second region: the same steps with per-phase CUDA events on the launching stream (eager launches) ----
    phase_steps = max(2, min(args.steps, 20))
    pfs.phase_timing(True)
    pfs.phase_times(reset=True)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    p0.record()
    for _ in range(phase_steps):
        step()
    p1.record()
    torch.cuda.synchronize()
    ms_step_eager = p0.elapsed_time(p1) / phase_steps
    phase_ms, phase_launches = pfs.phase_times(reset=True)
    pfs.phase_timing(False)
    phase_ms = {k: v * args.steps / phase_steps for k, v in phase_ms.items()}            # scaled to K steps
    phase_launches = {k: v * args.steps / phase_steps for k, v in phase_launches.items()}
    ms_step = ms_total / args.steps
    value = cells * n / (ms_step * 1e-3)

    # ---- roofline of the dominant kernel (largest share of the step) ----
    peak, peak_src = measured_peak_gbs()
    per_phase = {k: v / args.steps for k, v in phase_ms.items()}
    depth = pfs.get_fuse_depth()
    kernels = {}
    for phase, key, sweeps in (("diffuse", "diffuse_sweep", n), ("pressure", "pressure_sweep", n)):
        nl = phase_launches[phase] / args.steps
        if nl > 0 and per_phase[phase] > 0:
            bytes_per_launch = BYTES[key] * cells * sweeps / nl       # algorithmic bytes x sweeps fused per launch
            gbs = bytes_per_launch / (per_phase[phase] / nl * 1e-3) / 1e9
            kernels[phase] = {"launches_per_step": nl, "avg_launch_ms": per_phase[phase] / nl,
                              "sweeps_per_launch": sweeps / nl, "alg_bytes_per_launch": bytes_per_launch,
                              "achieved_gbs": gbs, "frac": gbs / peak, "share_of_step": per_phase[phase] / ms_step_eager}
    for phase, key in (("advect", "advect"), ("divergence", "divergence"), ("project", "project"),
                       ("advect_color", "advect_color")):
        if per_phase[phase] > 0:
            gbs = BYTES[key] * cells / (per_phase[phase] * 1e-3) / 1e9
            kernels[phase] = {"launches_per_step": phase_launches[phase] / args.steps, "avg_launch_ms": per_phase[phase],
                              "alg_bytes_per_launch": BYTES[key] * cells, "achieved_gbs": gbs, "frac": gbs / peak,
                              "share_of_step": per_phase[phase] / ms_step_eager}
    dom = max(kernels, key=lambda k: kernels[k]["share_of_step"])
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(dom, {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["achieved_gbs"], "peak": peak,
                "unit": "GB/s", "frac": kernels[dom]["frac"], "traffic": traffic, "peak_source": peak_src,
                "frac_of_nominal_8000_gbs": kernels[dom]["achieved_gbs"] / 8000.0,
                "alg_bytes_per_launch": kernels[dom]["alg_bytes_per_launch"],
                "avg_launch_ms": kernels[dom]["avg_launch_ms"],
                "note": "algorithmic bytes give no credit for temporal blocking, so frac may exceed 1; measured DRAM bytes per launch are in "
                        "`traffic`; after temporal blocking the fused sweeps are bound by FP32 instruction dispatch, not HBM "
                        "(DESIGN.md 3.4: ~20 cycles per (u,v) update per SM sub-partition is the floor, 22.7 measured)"}
    step_bytes = (88 + 16 * n + 12 * n) * cells
    whole_step = {"alg_bytes": step_bytes, "achieved_gbs": step_bytes / (ms_step * 1e-3) / 1e9,
                  "frac": step_bytes / (ms_step * 1e-3) / 1e9 / peak}

    # ---- the stateless entry points on the same workload (what a caller that keeps its own interleaved buffers gets) ----
    stateless = None
    if ctx is not None:
        def sl_step():
            pfs.simulate_fluid_step(fv, ft, DT, VISC, n, n)
            pfs.advect_color_step(fi, fm, fv, DT)
        for _ in range(3):
            sl_step()
        sl_steps = max(2, min(args.steps, 20))
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        s0.record()
        for _ in range(sl_steps):
            sl_step()
        s1.record()
        torch.cuda.synchronize()
        sl_ms = s0.elapsed_time(s1) / sl_steps
        stateless = {"ms_per_step": sl_ms, "value": cells * n / (sl_ms * 1e-3), "unit": UNIT, "steps": sl_steps,
                     "api": "pfs_simulate_fluid_step + pfs_advect_color_step on caller-owned interleaved device buffers"}

    # ---- end to end: host buffers in pinned memory, H2D + kernels + D2H inside the timed region ----
    e2e = None
    if not args.no_e2e:
        hv, ht, hi, hm = (pfs.pinned_empty(x.shape) for x in (vp, vtmp, image, itmp))
        hv[...] = vp; ht[...] = vtmp; hi[...] = image; hm[...] = itmp
        gv, gt, gi, gm = (pfs.vp_field(x) for x in (hv, ht, hi, hm))
        e2e_steps = max(3, min(args.steps, 10))
        for _ in range(2):
            pfs.timestep_host(gv, gt, gi, gm, DT, VISC, n, n)
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            pfs.timestep_host(gv, gt, gi, gm, DT, VISC, n, n)     # synchronises before returning
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        e2e = {"value": cells * n / e2e_s, "unit": UNIT, "ms_per_step": e2e_s * 1e3,
               "h2d_bytes_per_step": 3 * cells * 16, "d2h_bytes_per_step": 3 * cells * 16,
               "api": "pfs_timestep_host (vp_field structs with pinned host pointers; uploads vp, vtmp, image; "
                      "downloads vp, vtmp, image)", "steps": e2e_steps}

    # ---- CPU baseline: the reference's fluid.cpp on one host core, bounded sample ----
    cpu = None
    if not args.no_cpu:
        sample_rows = max(64, (1 << 20) // w)
        cpu_rate, cpu_ms, kind = cpu_reference_rate(n, w, sample_rows, 3, 1)
        cpu = {"value": cpu_rate, "unit": UNIT, "cores": 1, "host_cores": os.cpu_count(), "kind": kind,
               "sample": f"3 timesteps of a {w}x{sample_rows} slab of the workload ({n}+{n} sweeps), "
                         f"{cpu_ms:.0f} ms each; fluid.cpp is single-threaded"}

    refcu = None
    if not args.no_cpu:
        try:
            refcu = fluid_cu_baseline(n, w, h)
        except Exception as e:                      # a baseline must never take the bench down
            refcu = {"error": repr(e)}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks, "whole_step_roofline": whole_step,
            "pressure_solve": {"ms": per_phase["pressure"], "cell_updates_per_s": cells * n / (per_phase["pressure"] * 1e-3)},
            "cell_steps_per_s": cells / (ms_step * 1e-3), "phases_ms": per_phase, "kernels": kernels,
            "fuse_depth": depth, "fluid_cu_baseline": refcu,
            "api": ("pfs_ctx_step on a persistent-state context (fields resident in HBM in the planar layout)" if ctx is not None
                    else "pfs_simulate_fluid_step + pfs_advect_color_step (stateless, --stateless)"),
            "stateless_entry_points": stateless,
            "launch_mode": ("torch CUDA graph of two timesteps (--graph)" if graph is not None else
                            "library step graphs: a step whose plane roles were seen before replays a captured CUDA graph "
                            "(PFS_STEP_GRAPH=0 disables)"),
            "phase_region": {"steps": phase_steps, "ms_per_step_eager_with_phase_events": ms_step_eager},
            "setup": {"graph_priming_steps": PRIME_STEPS,
                      "what": "untimed steps before the warm-up, part of setup like the upload: one period of the library's step-graph "
                              "role assignments, so that no graph is captured during warm-up or the timed region"}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--fuse-depth", type=int, default=None)
    ap.add_argument("--graph", action="store_true",
                    help="replay the timed steps from a CUDA graph of two captured timesteps (small, launch-bound grids)")
    ap.add_argument("--stateless", action="store_true",
                    help="time the stateless entry points (caller-owned interleaved buffers) instead of a persistent context")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end leg (tuning runs)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg (tuning runs)")
    ap.add_argument("--no-unit", action="store_true",
                    help="multi-GPU arm: skip the 1-GPU measurement of the per-GPU band shape (the weak-scaling unit)")
    ap.add_argument("--no-parity", action="store_true",
                    help="multi-GPU arm: skip the oracle check of the ring that precedes the timed region (tuning runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference_arm(args)
        return
    args.warmup = max(args.warmup, 3)
    if args.fuse_depth is not None:
        import probabilistic_fluid_simulation_b200 as pfs
        pfs.set_fuse_depth(args.fuse_depth)
    if args.gpus == 1 and int(os.environ.get("WORLD_SIZE", "1")) == 1:
        run_single_gpu(args)
    else:
        from probabilistic_fluid_simulation_b200 import slab_bench
        slab_bench.run(args, sys.modules[__name__])


if __name__ == "__main__":
    main()
