"""
NumPy stand-in for one rank of a sharded workspace (CPU tests of the multi-GPU driver).
It replays the per-epoch tile-pass programs compiled by the real scheduler
(``aqc_debug_program_sharded``) with the oracle's gate primitives, so the epoch planner, the two
data layouts, the exchange protocol and the partial-sum reductions of
``aqc_research_b200/sharded.py`` are exercised end to end without a GPU (gloo process group).
"""

import ctypes as ct
import numpy as np

from aqc_research_b200 import _lib
from aqc_research_b200.engine import CircuitHandle
from program_sim import replay


def _parse_sharded(words):
    w = [int(x) for x in words]
    pos = 0
    nep = w[pos]
    pos += 1
    epochs = []
    for _ in range(nep):
        layout, npass = w[pos], w[pos + 1]
        pos += 2
        passes = []
        for _ in range(npass):
            tb, nstages, nouter = w[pos : pos + 3]
            pos += 3
            bitpos = w[pos : pos + 16]
            pos += 16
            outer = w[pos : pos + 48]
            pos += 48
            stages = []
            for _ in range(nstages):
                p, q, nunits = w[pos : pos + 3]
                pos += 3
                units = []
                for u in range(3):
                    kind, flags, theta = w[pos : pos + 3]
                    pos += 3
                    if u < nunits:
                        units.append((kind, flags, theta))
                stages.append((p, q, units))
            passes.append(dict(tb=tb, nouter=nouter, bitpos=bitpos[:tb], outer=outer[:nouter], stages=stages))
        epochs.append((layout, passes))
    assert pos == len(w)
    return epochs


def sharded_programs(circ, g, tile_bits=6, low_bits=2):
    lib = _lib.load()
    h = CircuitHandle(circ)
    out = {}
    for mode, rev in ((0, 0), (1, 0), (2, 1)):
        need = ct.c_int64(0)
        _lib.check(lib.aqc_debug_program_sharded(h.handle, g, tile_bits, low_bits, rev, None, 0, ct.byref(need)))
        buf = np.zeros(need.value, dtype=np.int32)
        _lib.check(
            lib.aqc_debug_program_sharded(
                h.handle, g, tile_bits, low_bits, rev, buf.ctypes.data_as(_lib.c_int32_p), buf.size, ct.byref(need)
            )
        )
        out[mode] = _parse_sharded(buf)
    return out


class SimShardBackend:
    """Same interface as ``GpuShardBackend`` (the subset the driver uses), NumPy arithmetic."""

    def __init__(self, circ, log2_world, rank, num_slots):
        self.circ = circ
        self.g, self.rank = log2_world, rank
        self.n = circ.num_qubits
        self.size = 1 << (self.n - log2_world)
        self.num_slots = num_slots
        self.num_thetas = circ.num_thetas
        self.slots = [np.zeros(self.size, dtype=np.complex128) for _ in range(num_slots)]
        self.progs = sharded_programs(circ, log2_world)
        self.thetas = None
        self.gacc = None

    def num_epochs(self, mode):
        return len(self.progs[mode])

    def epoch_layout(self, mode, epoch):
        return self.progs[mode][epoch][0]

    def begin(self, thetas, mode):
        self.thetas = np.array(thetas, dtype=np.float64)
        if mode == 0:
            self.gacc = np.zeros(self.num_thetas, dtype=np.complex128)

    # -- emulation of the peer-memory path (enabled by the worker's --sim-push): IPC mapping is a no-op and
    #    a fused push is the epoch followed by the block transpose into the landing slots, so the slot
    #    pool / landing-slot logic of the driver runs on the CPU exactly as it does with the CUDA backend
    transpose = None  # set to DistComm.transpose_chunks

    def ipc_export(self, slot):
        if self.transpose is None:
            raise RuntimeError("no peer memory in the NumPy stand-in")
        return bytes(64)

    def ipc_import(self, peer_rank, slot, handle):
        pass

    def can_push(self):
        return self.transpose is not None

    def exchange_p2p(self, src, dst):
        self.transpose(self.slot_tensor(src), self.slot_tensor(dst))

    def run_epoch(self, mode, epoch, src0, basis_local, src1, dst0, dst1, push0=-1, push1=-1):
        if push0 >= 0:
            assert self.transpose is not None and push0 not in (src0, dst0) and (mode != 0 or push1 not in (src1, dst1, push0))
            self.run_epoch(mode, epoch, src0, basis_local, src1, dst0, dst1)
            self.slots[push0][:] = np.nan  # a landing slot holds nothing before the delivery
            self.transpose(self.slot_tensor(dst0), self.slot_tensor(push0))
            if mode == 0:
                self.slots[push1][:] = np.nan
                self.transpose(self.slot_tensor(dst1), self.slot_tensor(push1))
            self.slots[dst0][:] = np.nan  # the in-place result is scratch once pushed
            if mode == 0:
                self.slots[dst1][:] = np.nan
            return
        passes = self.progs[mode][epoch][1]
        if src0 >= 0:
            v0 = self.slots[src0].copy()
        else:
            v0 = np.zeros(self.size, dtype=np.complex128)
            if basis_local >= 0:
                v0[basis_local] = 1
        if mode == 0:
            (w, z), g = replay(passes, self.circ.entangler, self.thetas, [v0, self.slots[src1]], dagger=False, grad=True)
            self.slots[dst0][:], self.slots[dst1][:] = w, z
            self.gacc += g
        else:
            (v,), _ = replay(passes, self.circ.entangler, self.thetas, [v0], dagger=(mode == 2), grad=False)
            self.slots[dst0][:] = v

    def grad_finish(self):
        return self.gacc.copy()

    def gather(self, slot, local_indices):
        return self.slots[slot][np.asarray(local_indices, dtype=np.int64)]

    def set_basis(self, slot, local_index):
        self.slots[slot][:] = 0
        if local_index >= 0:
            self.slots[slot][local_index] = 1

    def upload(self, slot, data):
        self.slots[slot][:] = data

    def download(self, slot):
        return self.slots[slot].copy()

    def vdot(self, a, b):
        return complex(np.vdot(self.slots[a], self.slots[b]))

    def slot_tensor(self, slot):
        import torch

        return torch.from_numpy(self.slots[slot].view(np.float64))


def shard_of(vec, n, g, rank):
    """Local part (layout A) of a full logical vector."""
    nl, cb = n - g, n - 2 * g
    off = np.arange(1 << nl)
    logical = (rank << nl) | ((off & ((1 << cb) - 1)) << g) | (off >> cb)
    return vec[logical]
