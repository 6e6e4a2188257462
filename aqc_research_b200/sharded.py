"""
One state vector over 2^g GPUs (global-qubit sharding) -- SURVEY.md section 8(e), item 2.

One process per GPU (``torchrun``); ``torch.distributed`` is the plumbing (barriers, the tiny
all-reduce of the partial gradient sums and of the gathered amplitudes).  The heavy step, the
layout switch between epochs, is a block transpose over the ranks.  It is FUSED into the last tile
pass of every epoch: instead of writing its tiles back in place, the pass stores each 256-byte run
straight into the HBM of the rank it belongs to after the transpose (peer stores over NVLink, the
peers' buffers mapped with CUDA IPC), so the transfer overlaps the DMMA work tile by tile and no
separate copy pass over the vector exists.  Fallbacks: ``aqc_sv_exchange`` (a separate pull kernel,
``AQC_SHARD_PUSH=0``) and, without IPC, ``isend / irecv`` pairs of the process group (NCCL, or gloo
in the CPU tests).

Layouts (see ``build_program_sharded`` in csrc/aqc_sv.cu), n qubits, g = log2(world), nl = n - g:
  A: qubits n-g..n-1 global;  qubits 0..g-1 on local bits nl-g..nl-1;  qubit q -> bit q - g otherwise
  B: qubits 0..g-1 global;    qubits n-g..n-1 on local bits nl-g..nl-1; the rest as in A
Data at rest (between calls) are always in layout A.

The numeric backend is pluggable: ``GpuShardBackend`` (the CUDA workspace) in production and a
NumPy replay of the compiled program in the CPU test-suite (tests/sharded_sim.py).
"""

import ctypes as ct
from typing import Dict, List, Optional, Sequence
import numpy as np
from . import _lib
from .engine import CircuitHandle, _dptr, _thetas_ptr
from .parametric_circuit import ParametricCircuit

MODE_GRAD, MODE_FWD, MODE_DAG = 0, 1, 2


def locate(index: int, n: int, g: int):
    """Logical amplitude index -> (rank, local offset) in layout A."""
    nl = n - g
    cb = nl - g
    rank = index >> nl
    low = index & ((1 << g) - 1)  # qubits 0..g-1 -> top local bits
    mid = (index >> g) & ((1 << cb) - 1)
    return rank, (low << cb) | mid


class GpuShardBackend:
    """Thin wrapper of a sharded ``aqc_sv`` workspace (one rank)."""

    def __init__(self, circ: ParametricCircuit, log2_world: int, rank: int, device: int, num_slots: int):
        self._lib = _lib.load()
        self.circuit = CircuitHandle(circ)
        self.num_thetas = circ.num_thetas
        self.num_slots = num_slots
        handle = ct.c_void_p()
        _lib.check(
            self._lib.aqc_sv_create_sharded(
                self.circuit.handle, device, log2_world, rank, num_slots, ct.byref(handle)
            )
        )
        self.handle = handle
        self.size = int(self._lib.aqc_sv_state_size(handle))
        self.device = device
        self.p2p_ready = False

    def num_epochs(self, mode):
        return int(self._lib.aqc_sv_num_epochs(self.handle, mode))

    def epoch_layout(self, mode, epoch):
        return int(self._lib.aqc_sv_epoch_layout(self.handle, mode, epoch))

    def begin(self, thetas, mode):
        _, ptr = _thetas_ptr(thetas, self.num_thetas)
        _lib.check(self._lib.aqc_sv_begin(self.handle, ptr, mode))

    def run_epoch(self, mode, epoch, src0, basis_local, src1, dst0, dst1, push0=-1, push1=-1):
        _lib.check(
            self._lib.aqc_sv_run_epoch(self.handle, mode, epoch, src0, int(basis_local), src1, dst0, dst1,
                                       push0, push1)
        )

    def can_push(self) -> bool:
        return bool(self._lib.aqc_sv_can_push(self.handle))

    def ipc_close(self):
        if getattr(self, "handle", None):
            _lib.check(self._lib.aqc_sv_ipc_close(self.handle))

    def grad_finish(self) -> np.ndarray:
        out = np.empty(self.num_thetas, dtype=np.complex128)
        _lib.check(self._lib.aqc_sv_grad_finish(self.handle, _dptr(out)))
        return out

    def gather(self, slot, local_indices) -> np.ndarray:
        idx = np.ascontiguousarray(local_indices, dtype=np.int64)
        out = np.empty(idx.size, dtype=np.complex128)
        _lib.check(
            self._lib.aqc_sv_gather(self.handle, slot, idx.ctypes.data_as(_lib.c_int64_p), idx.size, _dptr(out))
        )
        return out

    def set_basis(self, slot, local_index):
        _lib.check(self._lib.aqc_sv_set_basis(self.handle, slot, int(local_index)))

    def fill_random_logical(self, slot, seed) -> float:
        out = ct.c_double(0.0)
        _lib.check(self._lib.aqc_sv_fill_random_logical(self.handle, slot, int(seed), ct.byref(out)))
        return float(out.value)

    def scale(self, slot, factor):
        _lib.check(self._lib.aqc_sv_scale(self.handle, slot, float(factor)))

    def vdot(self, a, b) -> complex:
        out = np.empty(1, dtype=np.complex128)
        _lib.check(self._lib.aqc_sv_vdot(self.handle, a, b, _dptr(out)))
        return complex(out[0])

    def upload(self, slot, data):
        arr = np.ascontiguousarray(data, dtype=np.complex128).ravel()
        _lib.check(self._lib.aqc_sv_upload(self.handle, slot, -1, _dptr(arr), arr.size))

    def download(self, slot) -> np.ndarray:
        out = np.empty(self.size, dtype=np.complex128)
        _lib.check(self._lib.aqc_sv_download(self.handle, slot, 0, _dptr(out), out.size))
        return out

    # -- layout switch ---------------------------------------------------------------------
    def ipc_export(self, slot) -> bytes:
        buf = (ct.c_ubyte * 64)()
        _lib.check(self._lib.aqc_sv_ipc_export(self.handle, slot, buf))
        return bytes(buf)

    def ipc_import(self, peer_rank, slot, handle: bytes):
        buf = (ct.c_ubyte * 64).from_buffer_copy(handle)
        _lib.check(self._lib.aqc_sv_ipc_import(self.handle, peer_rank, slot, buf))

    def exchange_p2p(self, src, dst):
        _lib.check(self._lib.aqc_sv_exchange(self.handle, src, dst))

    def slot_tensor(self, slot):
        """torch view (float64, 2 * size) of a slot's device memory (zero copy)."""
        import torch  # pylint: disable=import-outside-toplevel

        ptr = int(self._lib.aqc_sv_slot_ptr(self.handle, slot))

        class _Iface:  # pylint: disable=too-few-public-methods
            __cuda_array_interface__ = {
                "shape": (2 * self.size,), "typestr": "<f8", "data": (ptr, False), "version": 2,
            }

        return torch.as_tensor(_Iface(), device=f"cuda:{self.device}")

    @property
    def last_kernel_ms(self):
        return float(self._lib.aqc_sv_last_kernel_ms(self.handle))

    def close(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            self._lib.aqc_sv_destroy(h)

    def __del__(self):
        self.close()


class DistComm:
    """torch.distributed process group (NCCL on GPUs, gloo on CPU)."""

    def __init__(self):
        import torch.distributed as dist  # pylint: disable=import-outside-toplevel

        self.dist = dist
        self.rank = dist.get_rank()
        self.world = dist.get_world_size()
        self.backend = dist.get_backend()

    def barrier(self):
        self.dist.barrier()

    def allreduce_sum(self, arr: np.ndarray, device: Optional[str] = None) -> np.ndarray:
        import torch  # pylint: disable=import-outside-toplevel

        flat = np.ascontiguousarray(arr).view(np.float64).copy()
        t = torch.from_numpy(flat)
        if self.backend == "nccl":
            t = t.to(device)
        self.dist.all_reduce(t)
        return t.cpu().numpy().view(arr.dtype).reshape(arr.shape)

    def allgather_bytes(self, blob: bytes) -> List[bytes]:
        out = [None] * self.world
        self.dist.all_gather_object(out, blob)
        return out

    def transpose_chunks(self, send, recv):
        """recv[chunk r] = (rank r).send[chunk my_rank]; send/recv: 1-D tensors of world equal chunks."""
        world, rank = self.world, self.rank
        cs = send.numel() // world
        ops = []
        for r in range(world):
            if r == rank:
                recv[r * cs : (r + 1) * cs].copy_(send[r * cs : (r + 1) * cs])
            else:
                ops.append(self.dist.P2POp(self.dist.irecv, recv[r * cs : (r + 1) * cs], r))
                ops.append(self.dist.P2POp(self.dist.isend, send[r * cs : (r + 1) * cs], r))
        if ops:
            for req in self.dist.batch_isend_irecv(ops):
                req.wait()


class ShardedStateVector:
    """
    Objective / gradient building blocks on a sharded state.  Vectors are addressed by NAME
    ("target", "z0", "w", "z"); the driver keeps the name -> slot map and a pool of free slots,
    because a layout switch delivers a vector into another slot.  Five slots are enough for the
    gradient sweep: the target, the two swept vectors and two landing slots (the sweep starts IN
    PLACE on the slot of z0 = V^H target, which every evaluation recomputes anyway).
    """

    NAMES = ("target", "z0", "w", "z")

    def __init__(self, circ: ParametricCircuit, comm, backend, use_p2p: bool = True):
        import os  # pylint: disable=import-outside-toplevel

        self.circ = circ
        self.comm = comm
        self.be = backend
        self.n = circ.num_qubits
        self.g = int(comm.world).bit_length() - 1
        assert 1 << self.g == comm.world, "world size must be a power of two"
        assert backend.num_slots >= 5, "a sharded state needs at least 5 slots"
        self.slot: Dict[str, int] = {"target": 0}
        self.free: List[int] = list(range(1, backend.num_slots))
        for name in ("z0", "w", "z"):  # every name owns a slot at rest (callers upload / download by name)
            self.slot[name] = self.free.pop(0)
        self.layout: Dict[str, int] = {name: 0 for name in self.NAMES}
        self.exchange_ms = 0.0
        self.compute_ms = 0.0
        self.p2p = False
        if use_p2p and hasattr(backend, "ipc_export"):
            self._setup_p2p()
        self.push = bool(self.p2p and getattr(backend, "can_push", lambda: False)()
                         and os.environ.get("AQC_SHARD_PUSH", "1") != "0")

    def _setup_p2p(self):
        """Maps every slot of every peer through CUDA IPC; falls back to send/recv on failure."""
        ok = 1
        try:
            mine = [self.be.ipc_export(s) for s in range(self.be.num_slots)]
            everyone = self.comm.allgather_bytes(b"".join(mine))
            for r, blob in enumerate(everyone):
                for s in range(self.be.num_slots):
                    self.be.ipc_import(r, s, blob[64 * s : 64 * (s + 1)])
        except Exception:  # noqa: BLE001  (IPC unavailable: use the process group instead)
            ok = 0
        total = self.comm.allreduce_sum(np.array([float(ok)]), device=self._dev())
        self.p2p = bool(total[0] == self.comm.world)

    def close(self):
        """Importers unmap the peers' buffers, all ranks meet, then the buffers are freed."""
        if self.p2p and hasattr(self.be, "ipc_close"):
            self.be.ipc_close()
        self.comm.barrier()
        if hasattr(self.be, "close"):
            self.be.close()

    def _dev(self):
        return f"cuda:{self.be.device}" if hasattr(self.be, "device") else None

    # -- layout switch (separate pass: pull kernel or send / recv) ---------------------------------
    def _switch(self, names: Sequence[str]):
        import time  # pylint: disable=import-outside-toplevel

        for name in names:
            src, dst = self.slot[name], self.free.pop(0)
            self.comm.barrier()  # every rank finished writing its source
            t0 = time.perf_counter()
            if self.p2p:
                self.be.exchange_p2p(src, dst)
            else:
                self.comm.transpose_chunks(self.be.slot_tensor(src), self.be.slot_tensor(dst))
            self.comm.barrier()  # every rank pulled: sources are free again
            self.exchange_ms += (time.perf_counter() - t0) * 1e3
            self.slot[name] = dst
            self.free.append(src)
            self.layout[name] ^= 1

    def _run(self, mode, first_sources, names, rest_in_layout_a=True):
        """
        Runs all epochs of ``mode`` on the named vectors (vec0[, vec1]); ``first_sources`` =
        (src0 slot or -1, local basis offset, src1 slot) are read by the first pass only.  The named
        vectors must own slots (their content is overwritten).  ``rest_in_layout_a=False`` leaves the
        results in the layout of the last epoch (work states nobody reads again).
        """
        import time  # pylint: disable=import-outside-toplevel

        be = self.be
        ne = be.num_epochs(mode)
        for e in range(ne):
            need = be.epoch_layout(mode, e)
            nxt = be.epoch_layout(mode, e + 1) if e + 1 < ne else (0 if rest_in_layout_a else need)
            dst = [self.slot[nm] for nm in names]
            if e == 0:
                if need != 0:
                    raise NotImplementedError("first epoch must run in layout A")
                src0, basis_local, src1 = first_sources
            else:
                assert self.layout[names[0]] == need
                src0, basis_local, src1 = dst[0], -1, dst[1] if len(dst) > 1 else 0
            d1 = dst[1] if len(dst) > 1 else 0
            if nxt != need and self.push:
                land = [self.free.pop(0) for _ in names]
                be.run_epoch(mode, e, src0, basis_local, src1, dst[0], d1, land[0], land[1] if len(land) > 1 else -1)
                self.compute_ms += getattr(be, "last_kernel_ms", 0.0)
                t0 = time.perf_counter()
                self.comm.barrier()  # every rank has delivered: the landing slots are complete
                self.exchange_ms += (time.perf_counter() - t0) * 1e3
                for nm, ls in zip(names, land):
                    self.free.append(self.slot[nm])
                    self.slot[nm] = ls
                    self.layout[nm] = nxt
            else:
                be.run_epoch(mode, e, src0, basis_local, src1, dst[0], d1)
                self.compute_ms += getattr(be, "last_kernel_ms", 0.0)
                for nm in names:
                    self.layout[nm] = need
                if nxt != need:
                    self._switch(names)

    # -- public operations -------------------------------------------------------------------
    def set_target_random(self, seed: int):
        local = self.be.fill_random_logical(self.slot["target"], seed)
        total = self.comm.allreduce_sum(np.array([local]), device=self._dev())[0]
        self.be.scale(self.slot["target"], 1.0 / np.sqrt(total))

    def set_basis(self, name: str, index: int):
        rank, off = locate(index, self.n, self.g)
        self.be.set_basis(self.slot[name], off if rank == self.comm.rank else -1)
        self.layout[name] = 0

    def apply(self, thetas, src: str, dst: str, dagger: bool = False):
        """dst = V src or V^H src (src in layout A; dst may equal src)."""
        mode = MODE_DAG if dagger else MODE_FWD
        self.be.begin(thetas, mode)
        self._run(mode, (self.slot[src], -1, 0), [dst])

    def amplitudes(self, name: str, indices) -> np.ndarray:
        """<e_idx | vector> for logical basis indices (sum of per-rank gathers)."""
        vals = np.zeros(len(indices), dtype=np.complex128)
        mine = [(i, locate(int(ix), self.n, self.g)) for i, ix in enumerate(indices)]
        pos = [i for i, (r, _) in mine if r == self.comm.rank]
        if pos:
            vals[pos] = self.be.gather(self.slot[name], [mine[i][1][1] for i in pos])
        return self.comm.allreduce_sum(vals, device=self._dev())

    def objective(self, thetas, indices) -> np.ndarray:
        """z0 = V^H target; returns hs_i = z0[indices] (objective_lhs_sur_max.py:98-106)."""
        self.apply(thetas, "target", "z0", dagger=True)
        return self.amplitudes("z0", indices)

    def grad(self, thetas, x_basis: int, keep_states: bool = False) -> np.ndarray:
        """
        Complex gradient of <V e_x | y> given z0 (grad_of_dot_product).  The sweep runs in place on
        the slot of z0, which is CONSUMED (every evaluation recomputes it, objective() first);
        ``keep_states`` brings the final w = V e_x and z = V z0 back to layout A.
        """
        rank, off = locate(int(x_basis), self.n, self.g)
        self.be.begin(thetas, MODE_GRAD)
        # z takes over the slot of z0 and sweeps it in place; the old slot of z joins the free slots, so that
        # two landing slots exist for the fused layout switches (5 slots in all: n = 34 on 8 GPUs)
        self.free.append(self.slot["z"])
        self.slot["z"] = self.slot.pop("z0")
        self._run(MODE_GRAD, (-1, off if rank == self.comm.rank else -1, self.slot["z"]), ["w", "z"],
                  rest_in_layout_a=keep_states)
        self.slot["z0"] = self.free.pop(0)  # content undefined until the next objective()
        return self.comm.allreduce_sum(self.be.grad_finish(), device=self._dev())

    def vdot(self, a: str, b: str) -> complex:
        v = self.be.vdot(self.slot[a], self.slot[b])
        return complex(self.comm.allreduce_sum(np.array([v]), device=self._dev())[0])
