"""Row-slab decomposition of one periodic grid over several GPUs (csrc/slab.cu).

Two ways to drive it, both through the C-ABI (``pfs_slab_*`` in include/pfs_b200.h):

* :class:`SlabRank` -- one process per GPU (torchrun): every process creates its rank's slab, rank 0
  creates an NCCL id that the caller broadcasts (torch.distributed, any backend), and each timestep is
  ``simulate_fluid_step`` / ``advect_color_step`` on the rank's band of the interleaved buffers.
* :class:`SlabRing` -- all ranks inside one process (on one or several devices); halo rows move by
  direct device copies.  This is what the single-GPU tests use to check the multi-rank logic bit for bit.

The partition itself is host arithmetic (:func:`partition`) and works without a GPU.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _cabi
from ._cabi import check
from .fluid import NUM_JACOBI_ITERS, vp_field


def partition(rank: int, nranks: int, gh: int, ih: int = 0):
    """-> (row0, rows, irow0, irows): the band of velocity rows of `rank` and the band of image rows
    whose velocity look-up (int)((float)j * (gh/ih)) (fluid.cpp:83,90) falls into it."""
    vals = [ctypes.c_int(0) for _ in range(4)]
    check(_cabi.lib().pfs_slab_partition(rank, nranks, gh, ih, *(ctypes.byref(v) for v in vals)))
    return tuple(v.value for v in vals)


def _ptr_array(values):
    arr = (ctypes.c_void_p * len(values))(*values)
    return arr


class _SlabBase:
    def __init__(self):
        self._handles = []

    def _create(self, rank, nranks, gw, gh, iw, ih):
        h = ctypes.c_void_p()
        check(_cabi.lib().pfs_slab_create(ctypes.byref(h), rank, nranks, gw, gh, iw, ih))
        self._handles.append(h)
        return h

    def close(self):
        L = _cabi.lib()
        for h in self._handles:
            L.pfs_slab_destroy(h)
        self._handles = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SlabRing(_SlabBase):
    """All `nranks` slabs of a gw x gh grid (+ iw x ih image) in this process."""

    def __init__(self, nranks: int, gw: int, gh: int, iw: int = 0, ih: int = 0, devices=None):
        import torch
        super().__init__()
        self.nranks, self.gw, self.gh, self.iw, self.ih = nranks, gw, gh, iw, ih
        self.devices = list(devices) if devices is not None else [torch.cuda.current_device()] * nranks
        assert len(self.devices) == nranks
        self.bands = [partition(r, nranks, gh, ih) for r in range(nranks)]
        for r in range(nranks):
            with torch.cuda.device(self.devices[r]):
                self._create(r, nranks, gw, gh, iw, ih)
        if nranks > 1:
            check(_cabi.lib().pfs_slab_connect_local(_ptr_array([h.value for h in self._handles]), nranks))

    # -- host <-> bands -------------------------------------------------------------------------
    def split(self, field: np.ndarray, image: bool = False):
        import torch
        out = []
        for r, (row0, rows, irow0, irows) in enumerate(self.bands):
            a, n = (irow0, irows) if image else (row0, rows)
            out.append(torch.from_numpy(np.ascontiguousarray(field[a:a + n])).to(f"cuda:{self.devices[r]}"))
        return out

    @staticmethod
    def gather(bands) -> np.ndarray:
        return np.concatenate([b.detach().cpu().numpy() for b in bands], axis=0)

    # -- steps ----------------------------------------------------------------------------------
    def _streams(self):
        import torch
        return _ptr_array([torch.cuda.current_stream(d).cuda_stream for d in self.devices])

    def simulate_fluid_step(self, vp: list, tmp: list, dt: float, viscosity: float,
                            n_diffuse: int = NUM_JACOBI_ITERS, n_pressure: int | None = None, forces: list | None = None) -> None:
        """vp / tmp: lists of the bands (CUDA tensors); entries are exchanged as the reference would.
        forces: list of the bands of the force field (addForces slot), or None."""
        n_pressure = n_diffuse if n_pressure is None else n_pressure
        by_ptr = {t.data_ptr(): t for t in vp + tmp}
        pv, pt = _ptr_array([t.data_ptr() for t in vp]), _ptr_array([t.data_ptr() for t in tmp])
        handles = _ptr_array([h.value for h in self._handles])
        if forces is not None:
            check(_cabi.lib().pfs_slab_simulate_fluid_step_forced(handles, self.nranks, pv, pt, dt, viscosity, n_diffuse,
                                                                  n_pressure, _ptr_array([t.data_ptr() for t in forces]),
                                                                  self._streams()))
        else:
            check(_cabi.lib().pfs_slab_simulate_fluid_step(handles, self.nranks, pv, pt, dt, viscosity, n_diffuse,
                                                           n_pressure, self._streams()))
        for k in range(self.nranks):
            vp[k], tmp[k] = by_ptr[pv[k]], by_ptr[pt[k]]

    def advect_color_step(self, image: list, itmp: list, vp: list, dt: float) -> None:
        by_ptr = {t.data_ptr(): t for t in image + itmp}
        pi, pm = _ptr_array([t.data_ptr() for t in image]), _ptr_array([t.data_ptr() for t in itmp])
        pv = _ptr_array([t.data_ptr() for t in vp])
        check(_cabi.lib().pfs_slab_advect_color_step(_ptr_array([h.value for h in self._handles]), self.nranks, pi, pm,
                                                     pv, dt, self._streams()))
        for k in range(self.nranks):
            image[k], itmp[k] = by_ptr[pi[k]], by_ptr[pm[k]]

    # -- resident state (pfs_slab_upload / pfs_slab_step / pfs_slab_download) ---------------------
    def upload(self, vp: list, tmp: list, image: list | None = None) -> None:
        """The bands (lists of CUDA tensors) become the slabs' resident state."""
        img = _ptr_array([t.data_ptr() for t in image]) if image is not None else None
        check(_cabi.lib().pfs_slab_upload(_ptr_array([h.value for h in self._handles]), self.nranks,
                                          _ptr_array([t.data_ptr() for t in vp]), _ptr_array([t.data_ptr() for t in tmp]), img,
                                          self._streams()))

    def step(self, n_steps: int, dt: float, viscosity: float, n_diffuse: int = NUM_JACOBI_ITERS, n_pressure: int | None = None) -> None:
        n_pressure = n_diffuse if n_pressure is None else n_pressure
        check(_cabi.lib().pfs_slab_step(_ptr_array([h.value for h in self._handles]), self.nranks, n_steps, dt, viscosity,
                                        n_diffuse, n_pressure, self._streams()))

    def download(self, vp: list, tmp: list, image: list | None = None) -> None:
        """Writes the resident state into the given bands (lists of CUDA tensors of the band shapes)."""
        img = _ptr_array([t.data_ptr() for t in image]) if image is not None else None
        check(_cabi.lib().pfs_slab_download(_ptr_array([h.value for h in self._handles]), self.nranks,
                                            _ptr_array([t.data_ptr() for t in vp]), _ptr_array([t.data_ptr() for t in tmp]), img,
                                            self._streams()))

    def compute_pressure_adaptive(self, vp: list, vp_out: list, dt: float, tol: float, max_sweeps: int, check_every: int = 16):
        """computePressure with the sweep count chosen at run time, on the bands (pfs_slab_compute_pressure_adaptive).
        -> (sweeps done, rms of the last update over the whole grid); entries of vp / vp_out are exchanged as the reference would."""
        by_ptr = {t.data_ptr(): t for t in vp + vp_out}
        pv, po = _ptr_array([t.data_ptr() for t in vp]), _ptr_array([t.data_ptr() for t in vp_out])
        n, rms = ctypes.c_int(0), ctypes.c_double(0.0)
        check(_cabi.lib().pfs_slab_compute_pressure_adaptive(_ptr_array([h.value for h in self._handles]), self.nranks, pv, po, dt, tol,
                                                             max_sweeps, check_every, ctypes.byref(n), ctypes.byref(rms), self._streams()))
        for k in range(self.nranks):
            vp[k], vp_out[k] = by_ptr[pv[k]], by_ptr[po[k]]
        return n.value, rms.value

    def check(self) -> None:
        check(_cabi.lib().pfs_slab_check(_ptr_array([h.value for h in self._handles]), self.nranks))

    def step_norms(self, vp: list, tmp: list) -> dict:
        out = (ctypes.c_double * 4)()
        check(_cabi.lib().pfs_slab_step_norms(_ptr_array([h.value for h in self._handles]), self.nranks,
                                              _ptr_array([t.data_ptr() for t in vp]), _ptr_array([t.data_ptr() for t in tmp]),
                                              out, self._streams()))
        return {"div_l2": out[0], "pressure_update_l2": out[1], "velocity_l2": out[2], "speed_max": out[3]}


class SlabRank(_SlabBase):
    """This process's slab of a ring of `nranks` processes (one GPU each), connected through NCCL."""

    def __init__(self, rank: int, nranks: int, gw: int, gh: int, iw: int = 0, ih: int = 0):
        super().__init__()
        self.rank, self.nranks, self.gw, self.gh, self.iw, self.ih = rank, nranks, gw, gh, iw, ih
        self.row0, self.rows, self.irow0, self.irows = partition(rank, nranks, gh, ih)
        self._h = self._create(rank, nranks, gw, gh, iw, ih)

    @staticmethod
    def unique_id() -> bytes:
        """Rank 0 only.  Ship the 128 bytes to every rank (e.g. torch.distributed.broadcast)."""
        buf = ctypes.create_string_buffer(128)
        check(_cabi.lib().pfs_slab_nccl_unique_id(buf))
        return buf.raw

    def connect(self, unique_id: bytes) -> None:
        """Collective: every rank calls it with rank 0's id."""
        assert len(unique_id) == 128
        if self.nranks > 1:
            check(_cabi.lib().pfs_slab_connect_nccl(self._h, ctypes.create_string_buffer(unique_id, 128)))

    @property
    def transport(self) -> str:
        """"p2p" (halo rows stored into the neighbours' memory over NVLink), "nccl" (send/recv) or "unconnected"."""
        return _cabi.lib().pfs_slab_transport(self._h).decode()

    def _stream(self, t):
        import torch
        return _ptr_array([torch.cuda.current_stream(t.device).cuda_stream])

    def simulate_fluid_step(self, vp: vp_field, tmp: vp_field, dt: float, viscosity: float,
                            n_diffuse: int = NUM_JACOBI_ITERS, n_pressure: int | None = None, forces=None) -> None:
        """forces: this rank's band of the force field (CUDA tensor shaped like vp.data), or None."""
        n_pressure = n_diffuse if n_pressure is None else n_pressure
        by_ptr = {vp.data.data_ptr(): vp.data, tmp.data.data_ptr(): tmp.data}
        pv, pt = _ptr_array([vp.data.data_ptr()]), _ptr_array([tmp.data.data_ptr()])
        if forces is not None:
            check(_cabi.lib().pfs_slab_simulate_fluid_step_forced(_ptr_array([self._h.value]), 1, pv, pt, dt, viscosity,
                                                                  n_diffuse, n_pressure, _ptr_array([forces.data_ptr()]),
                                                                  self._stream(vp.data)))
        else:
            check(_cabi.lib().pfs_slab_simulate_fluid_step(_ptr_array([self._h.value]), 1, pv, pt, dt, viscosity,
                                                           n_diffuse, n_pressure, self._stream(vp.data)))
        vp.data, tmp.data = by_ptr[pv[0]], by_ptr[pt[0]]

    def advect_color_step(self, image: vp_field, itmp: vp_field, vp: vp_field, dt: float) -> None:
        by_ptr = {image.data.data_ptr(): image.data, itmp.data.data_ptr(): itmp.data}
        pi, pm = _ptr_array([image.data.data_ptr()]), _ptr_array([itmp.data.data_ptr()])
        pv = _ptr_array([vp.data.data_ptr()])
        check(_cabi.lib().pfs_slab_advect_color_step(_ptr_array([self._h.value]), 1, pi, pm, pv, dt,
                                                     self._stream(vp.data)))
        image.data, itmp.data = by_ptr[pi[0]], by_ptr[pm[0]]

    # -- resident state (pfs_slab_upload / pfs_slab_step / pfs_slab_download) ---------------------
    def upload(self, vp=None, tmp=None, image=None, stream=None) -> None:
        """This rank's bands (CUDA tensors [rows, gw, 4] / [irows, iw, 4]) become the slab's resident state.  After the first
        upload the velocity pair (vp and tmp) or the image may be replaced on their own."""
        img = _ptr_array([image.data_ptr()]) if image is not None else None
        pv = _ptr_array([vp.data_ptr()]) if vp is not None else None
        pt = _ptr_array([tmp.data_ptr()]) if tmp is not None else None
        ref = vp if vp is not None else image
        check(_cabi.lib().pfs_slab_upload(_ptr_array([self._h.value]), 1, pv, pt, img,
                                          _ptr_array([stream]) if stream is not None else self._stream(ref)))

    def step_fluid(self, dt: float, viscosity: float, n_diffuse: int = NUM_JACOBI_ITERS, n_pressure: int | None = None, stream=None) -> None:
        import torch
        n_pressure = n_diffuse if n_pressure is None else n_pressure
        st = _ptr_array([stream if stream is not None else torch.cuda.current_stream().cuda_stream])
        check(_cabi.lib().pfs_slab_step_fluid(_ptr_array([self._h.value]), 1, dt, viscosity, n_diffuse, n_pressure, st))

    def step_color(self, dt: float, stream=None) -> None:
        import torch
        st = _ptr_array([stream if stream is not None else torch.cuda.current_stream().cuda_stream])
        check(_cabi.lib().pfs_slab_step_color(_ptr_array([self._h.value]), 1, dt, st))

    def step(self, n_steps: int, dt: float, viscosity: float, n_diffuse: int = NUM_JACOBI_ITERS, n_pressure: int | None = None,
             device=None) -> None:
        import torch
        n_pressure = n_diffuse if n_pressure is None else n_pressure
        stream = _ptr_array([torch.cuda.current_stream(device).cuda_stream])
        check(_cabi.lib().pfs_slab_step(_ptr_array([self._h.value]), 1, n_steps, dt, viscosity, n_diffuse, n_pressure, stream))

    def download(self, vp=None, tmp=None, image=None, stream=None) -> None:
        img = _ptr_array([image.data_ptr()]) if image is not None else None
        pv = _ptr_array([vp.data_ptr()]) if vp is not None else None
        pt = _ptr_array([tmp.data_ptr()]) if tmp is not None else None
        ref = vp if vp is not None else image
        check(_cabi.lib().pfs_slab_download(_ptr_array([self._h.value]), 1, pv, pt, img,
                                            _ptr_array([stream]) if stream is not None else self._stream(ref)))

    def check(self) -> None:
        check(_cabi.lib().pfs_slab_check(_ptr_array([self._h.value]), 1))

    def compute_pressure_adaptive(self, vp: vp_field, vp_out: vp_field, dt: float, tol: float, max_sweeps: int, check_every: int = 16):
        """This rank's part of pfs_slab_compute_pressure_adaptive (collective: the rms is all-reduced after every batch)."""
        by_ptr = {vp.data.data_ptr(): vp.data, vp_out.data.data_ptr(): vp_out.data}
        pv, po = _ptr_array([vp.data.data_ptr()]), _ptr_array([vp_out.data.data_ptr()])
        n, rms = ctypes.c_int(0), ctypes.c_double(0.0)
        check(_cabi.lib().pfs_slab_compute_pressure_adaptive(_ptr_array([self._h.value]), 1, pv, po, dt, tol, max_sweeps, check_every,
                                                             ctypes.byref(n), ctypes.byref(rms), self._stream(vp.data)))
        vp.data, vp_out.data = by_ptr[pv[0]], by_ptr[po[0]]
        return n.value, rms.value

    def step_norms(self, vp: vp_field, tmp: vp_field) -> dict:
        """Norms of the WHOLE grid (all-reduced over the ring); the same on every rank."""
        out = (ctypes.c_double * 4)()
        check(_cabi.lib().pfs_slab_step_norms(_ptr_array([self._h.value]), 1, _ptr_array([vp.data.data_ptr()]),
                                              _ptr_array([tmp.data.data_ptr()]), out, self._stream(vp.data)))
        return {"div_l2": out[0], "pressure_update_l2": out[1], "velocity_l2": out[2], "speed_max": out[3]}
