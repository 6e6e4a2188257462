"""
CPU replay of the tile-pass program compiled by the scheduler in csrc/aqc_sv.cu
(``aqc_debug_program``).  Test infrastructure: it interprets the serialised program unit by
unit with the oracle's primitive gates, so the SCHEDULER (pass/stage grouping, commutation
legality, local-bit mapping, Trotter flags, gradient slots) is validated without a GPU.
The CUDA arithmetic itself is validated by the ``-m gpu`` parity tests.
"""

import numpy as np
from oracle import sv_oracle as O

NWARPS = 8  # kDWarps of aqc_dense.cuh (warps per CTA of dense_pass_kernel)

U_FRONT_LO, U_FRONT_HI, U_BLOCK_CHI, U_BLOCK_CLO = 1, 2, 3, 4
F_PRE, F_POST = 1, 2


def parse_program(words: np.ndarray, dense: bool = False):
    """``dense``: format of aqc_debug_dense_program (5 units per stage + lane tables)."""
    w = [int(x) for x in words]
    per_stage = 5 if dense else 3
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
            for u in range(per_stage):
                kind, flags, theta = w[pos : pos + 3]
                pos += 3
                if u < nunits:
                    units.append((kind, flags, theta))
            if dense:
                rbits = w[pos : pos + 3]
                pos += 3
                lanes = np.array(w[pos : pos + NWARPS * 32 * 4], dtype=np.int64).reshape(NWARPS, 32, 4)
                pos += NWARPS * 32 * 4
                stages.append((p, q, units, rbits, lanes))
            else:
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
        for st in ps["stages"]:
            p, q, units = st[:3]
            assert 0 <= q < p < ps["tb"]
            assert 1 <= len(units) <= (5 if len(st) > 3 else 3)


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
        for st in ps["stages"]:
            p, q, units = st[:3]
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


# ------------------------------------------------------------------------------------------------
# Emulation of the dense-stage (DMMA) sweep kernel of csrc/aqc_dense.cuh: same lane tables, same
# fragment algebra (PTX mma.m8n8k4.f64: A[l>>2][l&3], B[l&3][l>>2], C[l>>2][2(l&3)+{0,1}]), same
# post-processing of the accumulated stage matrices.
# ------------------------------------------------------------------------------------------------
def _swz(i):
    return i ^ (((i >> 3) ^ (i >> 6) ^ (i >> 9)) & 7)


def _mini_pass(units):
    """The units of one stage as a program on bits (0 = lo, 1 = hi) of a small state."""
    return [dict(tb=2, nouter=0, bitpos=[0, 1, 2, 3], outer=[], stages=[(1, 0, units)])]


def stage_unitary(units, entangler, thetas, dagger):
    cols = []
    for k in range(4):
        e = np.zeros(4, dtype=np.complex128)
        e[k] = 1.0
        (v,), _ = replay(_mini_pass(units), entangler, thetas, [e], dagger=dagger, grad=False)
        cols.append(v)
    return np.stack(cols, axis=1)  # U[i][k]


PAIR_FIRST, PAIR_SWAP, SL_MASK = 0x8000, 0x4000, 0x0FFF  # flags in the `sl` word (aqc_dense.cuh)


def dense_steps(stages):
    """Splits a pass's stage list into steps: (first, second | None) stage indices in run order."""
    out, s = [], 0
    while s < len(stages):
        sl0 = int(stages[s][4][0, 0, 0])
        if sl0 & PAIR_FIRST:
            a = s + (1 if sl0 & PAIR_SWAP else 0)
            out.append((a, 2 * s + 1 - a, s))
            s += 2
        else:
            out.append((s, None, s))
            s += 1
    return out


def dense_check_tables(passes):
    """Every step's load and store tables address each tile element exactly once."""
    for ps in passes:
        tb = ps["tb"]
        nit = 1 << (tb - 5)
        for sa, sb_, st in dense_steps(ps["stages"]):
            p, q, units, rbits, lanes = ps["stages"][st]
            if sb_ is None:
                assert len({p, q, *rbits}) == 5 and all(0 <= r < tb for r in rbits)
            else:
                pa, qa = ps["stages"][sa][0], ps["stages"][sa][1]
                pb, qb = ps["stages"][sb_][0], ps["stages"][sb_][1]
                assert len({pa, qa, pb, qb}) == 4  # fused stages act on disjoint bit pairs
            slots, dslots = [], []
            for it in range(nit):
                w, j = it % NWARPS, it // NWARPS
                b = lanes[w, j, 3]
                slots += list(b ^ (lanes[w, :, 0] & SL_MASK))
                dslots += list((b << 1) ^ lanes[w, :, 1]) + list((b << 1) ^ lanes[w, :, 2])
            assert sorted(slots) == list(range(1 << tb))
            assert sorted(dslots) == list(range(2 << tb))


def dense_bank_conflicts(passes):
    """Worst-case shared-memory conflict degree over all stages: (loads, stores)."""
    worst_l = worst_s = 1
    for ps in passes:
        for p, q, units, rbits, lanes in ps["stages"]:
            sl = lanes[0, :, 0] & SL_MASK
            for qw in range(4):  # LDS.128: quarter warps, 8 bank groups of 16 bytes
                groups = [int(x) & 7 for x in sl[8 * qw : 8 * qw + 8]]
                worst_l = max(worst_l, max(groups.count(g) for g in set(groups)))
            for i in (1, 2):  # STS.64: half warps, 16 bank pairs of 8 bytes
                so = lanes[0, :, i]
                for hw in range(2):
                    banks = [int(x) & 15 for x in so[16 * hw : 16 * hw + 16]]
                    worst_s = max(worst_s, max(banks.count(g) for g in set(banks)))
    return worst_l, worst_s


def dense_emulate(passes, entangler, thetas, vecs, dagger, grad):
    """
    Runs the dense program the way dense_pass_kernel does.  Returns (vecs, complex grad | None);
    grad is the sum the post kernel (dense_grad_kernel) derives from the stage matrices.
    """
    vecs = [np.array(v, dtype=np.complex128).ravel().copy() for v in vecs]
    nv = len(vecs)
    g = np.zeros(thetas.size, dtype=np.complex128) if grad else None
    for ps in passes:
        tb, bp, outer = ps["tb"], ps["bitpos"], ps["outer"]
        tsize, nit = 1 << tb, 1 << (tb - 5)
        loc = np.zeros(tsize, dtype=np.int64)
        for k in range(tb):
            loc |= ((np.arange(tsize) >> k) & 1) << bp[k]
        swz = np.array([_swz(l) for l in range(tsize)])
        rmats = [np.zeros((8, 8)) for _ in ps["stages"]]
        for tile in range(1 << len(outer)):
            base = 0
            for k, b in enumerate(outer):
                base |= ((tile >> k) & 1) << b
            idx = base | loc
            sm = [np.zeros(2 * tsize) for _ in range(nv)]  # doubles, swizzled slots
            for v in range(nv):
                sm[v][2 * swz] = vecs[v][idx].real
                sm[v][2 * swz + 1] = vecs[v][idx].imag
            def real_form(units):
                U = stage_unitary(units, entangler, thetas, dagger)
                ua0, ua1 = np.zeros((8, 4)), np.zeros((8, 4))
                for i in range(4):
                    ua0[2 * i], ua1[2 * i] = U[i].real, -U[i].imag
                    ua0[2 * i + 1], ua1[2 * i + 1] = U[i].imag, U[i].real
                return ua0, ua1

            for sa, sb_, st in dense_steps(ps["stages"]):
                lanes = ps["stages"][st][4]
                ua0, ua1 = real_form(ps["stages"][sa][2])
                if sb_ is not None:
                    ub0, ub1 = real_form(ps["stages"][sb_][2])
                for it in range(nit):
                    w, j = it % NWARPS, it // NWARPS
                    b = int(lanes[w, j, 3])
                    lane = np.arange(32)
                    slot = b ^ (lanes[w, :, 0] & SL_MASK)
                    outs = []
                    for v in range(nv):
                        b0 = np.zeros((4, 8))
                        b1 = np.zeros((4, 8))
                        b0[lane & 3, lane >> 2] = sm[v][2 * slot]
                        b1[lane & 3, lane >> 2] = sm[v][2 * slot + 1]
                        outs.append(ua0 @ b0 + ua1 @ b1)  # D[c][g]
                    if grad:
                        rmats[sa] += outs[1] @ outs[0].T  # R[cz][cw] = sum_g Z[cz][g] W[cw][g]
                    if sb_ is not None:
                        # lane ^ 4 exchange: the lane keeps slot i = (lane >> 2) & 1 of its pair and
                        # receives the other real component of that amplitude from its partner
                        outs2 = []
                        for v in range(nv):
                            d = outs[v]
                            hi = (lane >> 2) & 1
                            cre, cim = (lane >> 2) & ~1, (lane >> 2) | 1
                            col = 2 * (lane & 3) + hi
                            xr, xi = d[cre, col], d[cim, col]
                            b0 = np.zeros((4, 8))
                            b1 = np.zeros((4, 8))
                            b0[lane & 3, lane >> 2] = xr
                            b1[lane & 3, lane >> 2] = xi
                            outs2.append(ub0 @ b0 + ub1 @ b1)
                        if grad:
                            rmats[sb_] += outs2[1] @ outs2[0].T
                        outs = outs2
                    for v in range(nv):
                        d = outs[v]
                        sm[v][(b << 1) ^ lanes[w, :, 1]] = d[lane >> 2, 2 * (lane & 3)]
                        sm[v][(b << 1) ^ lanes[w, :, 2]] = d[lane >> 2, 2 * (lane & 3) + 1]
            for v in range(nv):
                vecs[v][idx] = sm[v][2 * swz] + 1j * sm[v][2 * swz + 1]
        if grad:
            for si, (p, q, units, rbits, lanes) in enumerate(ps["stages"]):
                R = rmats[si]
                M = np.zeros((4, 4), dtype=np.complex128)
                for i in range(4):
                    for r in range(4):
                        M[i, r] = (R[2 * i, 2 * r] + R[2 * i + 1, 2 * r + 1]) + 1j * (
                            R[2 * i + 1, 2 * r] - R[2 * i, 2 * r + 1]
                        )
                # virtual quadruples on bits (0, 1); bits (2, 3) number them
                wv = np.zeros(16, dtype=np.complex128)
                zv = np.zeros(16, dtype=np.complex128)
                for r in range(4):
                    wv[4 * r + r] = 1.0
                    zv[4 * r : 4 * r + 4] = M[:, r]
                back = [(k, f, t) for (k, f, t) in reversed(units)]
                (wv, zv), _ = replay(_mini_pass(back), entangler, thetas, [wv, zv], dagger=True, grad=False)
                _, gs = replay(_mini_pass(units), entangler, thetas, [wv, zv], dagger=False, grad=True)
                g += gs
    return vecs, g
