/*
 * CPU ORACLE (test / baseline infrastructure only) -- plain C + OpenMP restatement of the
 * reference's state-vector objective/gradient algorithm, gate by gate, for sizes the NumPy
 * oracle (oracle/sv_oracle.py) is too slow for, and as the multi-threaded CPU baseline of
 * bench.py.  NOT part of the product: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.
 *
 * Parity status: PINNED -- tests/test_oracle_golden.py checks it against the golden vectors of
 * the unmodified reference (tests/golden/sv_cases.npz, mat_cases.npz) and against
 * oracle/sv_oracle.py.
 *
 * Reference (qiskit-community/aqc-research), paths relative to its root:
 *   rx/ry/rz_mul_vec            aqc_research/core_operations.py:164-264
 *   dot_x/dot_y/dot_z           aqc_research/core_operations.py:267-351
 *   cx/cz/cp_mul_vec            aqc_research/core_operations.py:422-558
 *   v_mul_vec, v_dagger_mul_vec aqc_research/core_operations.py:606-820
 *   grad_of_dot_product         aqc_research/core_operations.py:823-1019
 *   matrix variants             aqc_research/core_op_matrix.py:480-762 (qoff = log2 columns)
 * Qubit q <-> bit (q + qoff) of the flat index; complex numbers are (re, im) doubles.
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef double complex cplx;
enum { ROT_Y = 0, ROT_Z = 1, ROT_X = 2 };
enum { ENT_CX = 0, ENT_CZ = 1, ENT_CP = 2 };

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU baseline wants every host core */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

static inline int64_t insert0(int64_t j, int bit) {
  const int64_t lo = j & (((int64_t)1 << bit) - 1);
  return ((j - lo) << 1) | lo;
}

/* 2x2 gate {g00,g01,g10,g11} on bit `bit` of every vector in v[0..nv-1]; if dot != NULL also
 * returns sum conj((P w)_i) z_i for the Pauli of `rot` with w = v[0], z = v[1] AFTER the gate. */
static void gate1(int nbits, int bit, const cplx g[4], int rot, cplx** v, int nv, cplx* dot) {
  const int64_t half = (int64_t)1 << (nbits - 1);
  double dr = 0.0, di = 0.0;
#pragma omp parallel for reduction(+ : dr, di) schedule(static)
  for (int64_t j = 0; j < half; ++j) {
    const int64_t i0 = insert0(j, bit), i1 = i0 | ((int64_t)1 << bit);
    cplx w0 = 0, w1 = 0, z0 = 0, z1 = 0;
    for (int k = 0; k < nv; ++k) {
      const cplx a0 = v[k][i0], a1 = v[k][i1];
      const cplx b0 = g[0] * a0 + g[1] * a1, b1 = g[2] * a0 + g[3] * a1;
      v[k][i0] = b0;
      v[k][i1] = b1;
      if (k == 0) w0 = b0, w1 = b1;
      if (k == 1) z0 = b0, z1 = b1;
    }
    if (dot) {
      cplx d;
      if (rot == ROT_Y) /* (Y w)_0 = -i w1, (Y w)_1 = i w0 */
        d = conj(-I * w1) * z0 + conj(I * w0) * z1;
      else if (rot == ROT_Z)
        d = conj(w0) * z0 - conj(w1) * z1;
      else
        d = conj(w1) * z0 + conj(w0) * z1;
      dr += creal(d);
      di += cimag(d);
    }
  }
  if (dot) *dot = dr + I * di;
}

static void make_rot(int rot, double angle, cplx g[4]) {
  const double c = cos(0.5 * angle), s = sin(0.5 * angle);
  if (rot == ROT_Y) {
    g[0] = c, g[1] = -s, g[2] = s, g[3] = c;
  } else if (rot == ROT_Z) {
    g[0] = c - I * s, g[1] = 0, g[2] = 0, g[3] = c + I * s;
  } else {
    g[0] = c, g[1] = -I * s, g[2] = -I * s, g[3] = c;
  }
}

/* controlled gate |0><0|_c x I + |1><1|_c x G_t, G = X | Z | diag(1, e^{i phi});
 * if dot != NULL: sum over (c = 1, t = 1) of conj(w) z BEFORE the gate (cphase derivative). */
static void gate_ctrl(int nbits, int cb, int tb, int ent, double phi, cplx** v, int nv, cplx* dot) {
  const int64_t quarter = (int64_t)1 << (nbits - 2);
  const int lo = cb < tb ? cb : tb, hi = cb < tb ? tb : cb;
  const cplx ph = cos(phi) + I * sin(phi);
  double dr = 0.0, di = 0.0;
#pragma omp parallel for reduction(+ : dr, di) schedule(static)
  for (int64_t j = 0; j < quarter; ++j) {
    const int64_t base = insert0(insert0(j, lo), hi) | ((int64_t)1 << cb);
    const int64_t i0 = base, i1 = base | ((int64_t)1 << tb);
    if (dot) {
      const cplx d = conj(v[0][i1]) * v[1][i1];
      dr += creal(d);
      di += cimag(d);
    }
    for (int k = 0; k < nv; ++k) {
      if (ent == ENT_CX) {
        const cplx t = v[k][i0];
        v[k][i0] = v[k][i1];
        v[k][i1] = t;
      } else if (ent == ENT_CZ) {
        v[k][i1] = -v[k][i1];
      } else {
        v[k][i1] *= ph;
      }
    }
  }
  if (dot) *dot = dr + I * di;
}

static void rot_all(int nbits, int bit, int rot, double angle, cplx** v, int nv, cplx* dot) {
  cplx g[4];
  make_rot(rot, angle, g);
  gate1(nbits, bit, g, rot, v, nv, dot);
}

/* vec <- V vec (dagger = 0) or V^H vec (dagger = 1).  trotter: 0 generic, 1 first, 2 second order */
void orc_apply(int n, int qoff, int ent, int trotter, int nb, const int32_t* ctrl,
               const int32_t* targ, const double* th, double* vec_ri, int dagger) {
  cplx* vec = (cplx*)vec_ri;
  cplx* v[1] = {vec};
  const int nbits = n + qoff;
  const int tpb = ent == ENT_CP ? 5 : 4;
  const int half = (trotter == 2 && nb > 0) ? 3 * (n / 2) : 0;
  const int total = nb + half;
  const int rs = ent == ENT_CX ? ROT_X : ROT_Z;
  const double* th2 = th + 3 * n;
  if (!dagger) {
    for (int q = 0; q < n; ++q) {
      rot_all(nbits, q + qoff, ROT_Z, th[3 * q + 2], v, 1, 0);
      rot_all(nbits, q + qoff, ROT_Y, th[3 * q + 1], v, 1, 0);
      rot_all(nbits, q + qoff, ROT_Z, th[3 * q + 0], v, 1, 0);
    }
    for (int i = 0; i < total; ++i) {
      const int k = i % nb, c = ctrl[k] + qoff, t = targ[k] + qoff;
      const double* a = th2 + tpb * k;
      if (trotter && i % 3 == 0) rot_all(nbits, c, ROT_Z, -M_PI / 2, v, 1, 0);
      gate_ctrl(nbits, c, t, ent, ent == ENT_CP ? a[4] : 0.0, v, 1, 0);
      rot_all(nbits, c, ROT_Y, a[0], v, 1, 0);
      rot_all(nbits, c, ROT_Z, a[1], v, 1, 0);
      rot_all(nbits, t, ROT_Y, a[2], v, 1, 0);
      rot_all(nbits, t, rs, a[3], v, 1, 0);
      if (trotter && i % 3 == 2) rot_all(nbits, t, ROT_Z, M_PI / 2, v, 1, 0);
    }
  } else {
    for (int i = total - 1; i >= 0; --i) {
      const int k = i % nb, c = ctrl[k] + qoff, t = targ[k] + qoff;
      const double* a = th2 + tpb * k;
      if (trotter && i % 3 == 2) rot_all(nbits, t, ROT_Z, -M_PI / 2, v, 1, 0);
      rot_all(nbits, t, rs, -a[3], v, 1, 0);
      rot_all(nbits, t, ROT_Y, -a[2], v, 1, 0);
      rot_all(nbits, c, ROT_Z, -a[1], v, 1, 0);
      rot_all(nbits, c, ROT_Y, -a[0], v, 1, 0);
      gate_ctrl(nbits, c, t, ent, ent == ENT_CP ? -a[4] : 0.0, v, 1, 0);
      if (trotter && i % 3 == 0) rot_all(nbits, c, ROT_Z, M_PI / 2, v, 1, 0);
    }
    for (int q = 0; q < n; ++q) {
      rot_all(nbits, q + qoff, ROT_Z, -th[3 * q + 0], v, 1, 0);
      rot_all(nbits, q + qoff, ROT_Y, -th[3 * q + 1], v, 1, 0);
      rot_all(nbits, q + qoff, ROT_Z, -th[3 * q + 2], v, 1, 0);
    }
  }
}

/* Gradient sweep: w (initially x) and z (initially V^H y) are pushed through the circuit in
 * place; grad[k] (+)= 0.5j <P w|z> after each rotation (core_operations.py:919-1017). */
void orc_grad(int n, int qoff, int ent, int trotter, int nb, const int32_t* ctrl,
              const int32_t* targ, const double* th, double* w_ri, double* z_ri, double* grad_ri) {
  cplx* v[2] = {(cplx*)w_ri, (cplx*)z_ri};
  cplx* grad = (cplx*)grad_ri;
  const int nbits = n + qoff;
  const int tpb = ent == ENT_CP ? 5 : 4;
  const int half = (trotter == 2 && nb > 0) ? 3 * (n / 2) : 0;
  const int total = nb + half;
  const int rs = ent == ENT_CX ? ROT_X : ROT_Z;
  const double* th2 = th + 3 * n;
  const int T = 3 * n + tpb * nb;
  cplx d;
  for (int k = 0; k < T; ++k) grad[k] = 0;
  for (int q = 0; q < n; ++q) {
    rot_all(nbits, q + qoff, ROT_Z, th[3 * q + 2], v, 2, &d);
    grad[3 * q + 2] = 0.5 * I * d;
    rot_all(nbits, q + qoff, ROT_Y, th[3 * q + 1], v, 2, &d);
    grad[3 * q + 1] = 0.5 * I * d;
    rot_all(nbits, q + qoff, ROT_Z, th[3 * q + 0], v, 2, &d);
    grad[3 * q + 0] = 0.5 * I * d;
  }
  for (int i = 0; i < total; ++i) {
    const int k = i % nb, c = ctrl[k] + qoff, t = targ[k] + qoff;
    const double* a = th2 + tpb * k;
    cplx* g = grad + 3 * n + tpb * k;
    if (trotter && i % 3 == 0) rot_all(nbits, c, ROT_Z, -M_PI / 2, v, 2, 0);
    if (ent == ENT_CP) {
      gate_ctrl(nbits, c, t, ent, a[4], v, 2, &d);
      g[4] += -I * d; /* <(i e^{i phi} P11) w | E z> = -i sum conj(w11) z11 (:561-603,972-975) */
    } else {
      gate_ctrl(nbits, c, t, ent, 0.0, v, 2, 0);
    }
    rot_all(nbits, c, ROT_Y, a[0], v, 2, &d);
    g[0] += 0.5 * I * d;
    rot_all(nbits, c, ROT_Z, a[1], v, 2, &d);
    g[1] += 0.5 * I * d;
    rot_all(nbits, t, ROT_Y, a[2], v, 2, &d);
    g[2] += 0.5 * I * d;
    rot_all(nbits, t, rs, a[3], v, 2, &d);
    g[3] += 0.5 * I * d;
    if (trotter && i % 3 == 2) rot_all(nbits, t, ROT_Z, M_PI / 2, v, 2, 0);
  }
}
