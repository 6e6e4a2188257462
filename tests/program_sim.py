"""
CPU replay of the tile-pass program compiled by the scheduler in csrc/aqc_sv.cu
(``aqc_debug_program``).  Test infrastructure: it interprets the serialised program unit by
unit with the oracle's primitive gates, so the SCHEDULER (pass/stage grouping, commutation
legality, local-bit mapping, Trotter flags, gradient slots) is validated without a GPU.
The CUDA arithmetic itself is validated by the ``-m gpu`` parity tests.
"""

import numpy as np
from oracle import sv_oracle as O

U_FRONT_LO, U_FRONT_HI, U_BLOCK_CHI, U_BLOCK_CLO = 1, 2, 3, 4
F_PRE, F_POST = 1, 2


def parse_program(words: np.ndarray):
    w = [int(x) for x in words]
    pos = 0
    npasses = w[pos]
    pos += 1
    passes = []
    for _ in range(npasses):
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
    assert pos == len(w)
    return passes


def check_structure(passes, nbits):
    """Every pass partitions the index bits into tile bits and outer bits."""
    for ps in passes:
        bits = sorted(list(ps["bitpos"]) + list(ps["outer"]))
        assert bits == list(range(nbits)), (bits, nbits)
        assert list(ps["bitpos"]) == sorted(ps["bitpos"])
        for p, q, units in ps["stages"]:
            assert 0 <= q < p < ps["tb"]
            assert 1 <= len(units) <= 3


def replay(passes, entangler: str, thetas: np.ndarray, vecs, dagger: bool, grad: bool):
    """
    Runs the program on ``vecs`` (list of flat arrays: [v] for apply, [w, z] for grad).
    Returns (vecs, complex grad or None).
    """
    vecs = [np.array(v, dtype=np.complex128).ravel().copy() for v in vecs]
    g = np.zeros(thetas.size, dtype=np.complex128) if grad else None
    make_rs, pauli_s = (O.rx, O.PAULI_X) if entangler == "cx" else (O.rz, O.PAULI_Z)
    sgn = -1.0 if dagger else 1.0

    def rot(bit, gate, pauli, slot):
        for i, v in enumerate(vecs):
            vecs[i] = O.op1(v, bit, gate)
        if grad:
            g[slot] += O.pauli_dot(vecs[0], vecs[1], bit, pauli)

    def all1(bit, gate):
        for i, v in enumerate(vecs):
            vecs[i] = O.op1(v, bit, gate)

    def ent(c, t, th):
        if entangler == "cx":
            e = O.PAULI_X
        elif entangler == "cz":
            e = O.PAULI_Z
        else:
            e = O.phase(sgn * th[4])
        for i, v in enumerate(vecs):
            vecs[i] = O.ctrl_op(v, c, t, e)

    for ps in passes:
        bp = ps["bitpos"]
        for p, q, units in ps["stages"]:
            hi, lo = bp[p], bp[q]
            for kind, flags, theta in units:
                if kind in (U_FRONT_LO, U_FRONT_HI):
                    b = hi if kind == U_FRONT_HI else lo
                    th = thetas[theta : theta + 3]
                    if not dagger:
                        rot(b, O.rz(th[2]), O.PAULI_Z, theta + 2)
                        rot(b, O.ry(th[1]), O.PAULI_Y, theta + 1)
                        rot(b, O.rz(th[0]), O.PAULI_Z, theta + 0)
                    else:
                        all1(b, O.rz(-th[0]))
                        all1(b, O.ry(-th[1]))
                        all1(b, O.rz(-th[2]))
                else:
                    c, t = (hi, lo) if kind == U_BLOCK_CHI else (lo, hi)
                    tpb = 5 if entangler == "cp" else 4
                    th = thetas[theta : theta + tpb]
                    if not dagger:
                        if flags & F_PRE:
                            all1(c, O.rz(-np.pi / 2))
                        if grad and entangler == "cp":
                            rows = np.arange(vecs[0].size)
                            both = ((rows >> c) & 1 == 1) & ((rows >> t) & 1 == 1)
                            g[theta + 4] += -1j * np.vdot(vecs[0][both], vecs[1][both])
                        ent(c, t, th)
                        rot(c, O.ry(th[0]), O.PAULI_Y, theta + 0)
                        rot(c, O.rz(th[1]), O.PAULI_Z, theta + 1)
                        rot(t, O.ry(th[2]), O.PAULI_Y, theta + 2)
                        rot(t, make_rs(th[3]), pauli_s, theta + 3)
                        if flags & F_POST:
                            all1(t, O.rz(np.pi / 2))
                    else:
                        if flags & F_POST:
                            all1(t, O.rz(-np.pi / 2))
                        all1(t, make_rs(-th[3]))
                        all1(t, O.ry(-th[2]))
                        all1(c, O.rz(-th[1]))
                        all1(c, O.ry(-th[0]))
                        ent(c, t, th)
                        if flags & F_PRE:
                            all1(c, O.rz(np.pi / 2))
    return vecs, g
