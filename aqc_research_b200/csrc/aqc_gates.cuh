// aqc_gates.cuh -- register-level gate arithmetic shared by the state-vector engine (aqc_sv.cu)
// and the MPS engine (aqc_mps.cu).  An amplitude quadruple a[4] is indexed (hi << 1) | lo over the
// two register-resident bits of a stage.  Reference formulas: aqc_research/core_operations.py
// (rx/ry/rz_mul_vec :164-264, dot_x/y/z :267-351, cx/cz/cp_mul_vec :422-558, block order
// :686-708 / :787-809 / :956-1017); conventions aqc_research/elementary_operations.py:143-291.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "../../include/aqc_b200.h"

enum UnitKind : int32_t {
  U_NONE = 0,
  U_FRONT_LO = 1,  // Rz Ry Rz front gate on the low register bit of the quad
  U_FRONT_HI = 2,
  U_BLOCK_CHI = 3,  // unit block, control on the high register bit
  U_BLOCK_CLO = 4,
  U_SWAP = 5,  // MPS swap network only (aqc_mps.cu)
};
enum UnitFlags : int32_t {
  F_PRE = 1,   // Trotter Rz(-pi/2) on control before the block (i % 3 == 0)
  F_POST = 2,  // Trotter Rz(+pi/2) on target after the block  (i % 3 == 2)
};

struct cd {
  double x, y;
};

__device__ __forceinline__ void mul_mi(cd& a) {  // a *= -i
  const double t = a.x;
  a.x = a.y;
  a.y = -t;
}
__device__ __forceinline__ void mul_pi(cd& a) {  // a *= +i
  const double t = a.x;
  a.x = -a.y;
  a.y = t;
}
// a *= (c + i s)
__device__ __forceinline__ void mul_cs(cd& a, double c, double s) {
  const double x = a.x, y = a.y;
  a.x = fma(-s, y, c * x);
  a.y = fma(s, x, c * y);
}
__device__ __forceinline__ void ry_pair(cd& a0, cd& a1, double c, double s) {
  const cd b0 = a0, b1 = a1;
  a0.x = fma(-s, b1.x, c * b0.x);
  a0.y = fma(-s, b1.y, c * b0.y);
  a1.x = fma(s, b0.x, c * b1.x);
  a1.y = fma(s, b0.y, c * b1.y);
}
__device__ __forceinline__ void rx_pair(cd& a0, cd& a1, double c, double s) {
  const cd b0 = a0, b1 = a1;
  a0.x = fma(s, b1.y, c * b0.x);
  a0.y = fma(-s, b1.x, c * b0.y);
  a1.x = fma(s, b0.y, c * b1.x);
  a1.y = fma(-s, b0.x, c * b1.y);
}
__device__ __forceinline__ void rz_pair(cd& a0, cd& a1, double c, double s) {
  mul_cs(a0, c, -s);
  mul_cs(a1, c, s);
}
// acc += conj(w) * z
__device__ __forceinline__ void cdot_add(double* acc, const cd& w, const cd& z) {
  acc[0] = fma(w.x, z.x, acc[0]);
  acc[0] = fma(w.y, z.y, acc[0]);
  acc[1] = fma(w.x, z.y, acc[1]);
  acc[1] = fma(-w.y, z.x, acc[1]);
}
__device__ __forceinline__ void cdot_sub(double* acc, const cd& w, const cd& z) {
  acc[0] = fma(-w.x, z.x, acc[0]);
  acc[0] = fma(-w.y, z.y, acc[0]);
  acc[1] = fma(-w.x, z.y, acc[1]);
  acc[1] = fma(w.y, z.x, acc[1]);
}

enum { ROT_Y = 0, ROT_Z = 1, ROT_X = 2 };

// One-qubit rotation on register bit HI/LO of a quad for NVEC vectors, plus (NVEC == 2) the raw
// inner product <P w|z> / (i for Z, X) accumulated into acc[0..1]:
//   Y: sum conj(w0) z1 - conj(w1) z0      (dot_y core_operations.py:317-322, factor 0.5)
//   Z: sum conj(w0) z0 - conj(w1) z1      (dot_z :346-351, factor 0.5j)
//   X: sum conj(w1) z0 + conj(w0) z1      (dot_x :288-293, factor 0.5j)
template <int NVEC, bool HI, int ROT>
__device__ __forceinline__ void rot1q(cd (&a)[NVEC][4], double c, double s, double* acc) {
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int i0 = HI ? r : 2 * r;
    const int i1 = i0 + (HI ? 2 : 1);
#pragma unroll
    for (int v = 0; v < NVEC; ++v) {
      if (ROT == ROT_Y) ry_pair(a[v][i0], a[v][i1], c, s);
      if (ROT == ROT_Z) rz_pair(a[v][i0], a[v][i1], c, s);
      if (ROT == ROT_X) rx_pair(a[v][i0], a[v][i1], c, s);
    }
    if (NVEC == 2) {
      if (ROT == ROT_Y) {
        cdot_add(acc, a[0][i0], a[NVEC - 1][i1]);
        cdot_sub(acc, a[0][i1], a[NVEC - 1][i0]);
      }
      if (ROT == ROT_Z) {
        cdot_add(acc, a[0][i0], a[NVEC - 1][i0]);
        cdot_sub(acc, a[0][i1], a[NVEC - 1][i1]);
      }
      if (ROT == ROT_X) {
        cdot_add(acc, a[0][i1], a[NVEC - 1][i0]);
        cdot_add(acc, a[0][i0], a[NVEC - 1][i1]);
      }
    }
  }
}

// Front-layer gate of one qubit: forward Rz(t2) Ry(t1) Rz(t0) applied right-to-left
// (core_operations.py:921-935); dagger Rz(-t0) Ry(-t1) Rz(-t2) (:812-818).
template <int NVEC, bool HI, bool DAG>
__device__ __forceinline__ void front_unit(cd (&a)[NVEC][4], const double2* __restrict__ tr,
                                           double* acc) {
  const double2 t0 = tr[0], t1 = tr[1], t2 = tr[2];
  if (!DAG) {
    rot1q<NVEC, HI, ROT_Z>(a, t2.x, t2.y, acc + 4);
    rot1q<NVEC, HI, ROT_Y>(a, t1.x, t1.y, acc + 2);
    rot1q<NVEC, HI, ROT_Z>(a, t0.x, t0.y, acc + 0);
  } else {
    rot1q<NVEC, HI, ROT_Z>(a, t0.x, -t0.y, acc);
    rot1q<NVEC, HI, ROT_Y>(a, t1.x, -t1.y, acc);
    rot1q<NVEC, HI, ROT_Z>(a, t2.x, -t2.y, acc);
  }
}

// Unit block (core_operations.py:686-708 forward, :787-809 dagger, :956-1017 gradient sweep).
// Trotter Rz(-+pi/2) = e^{+-i pi/4} diag(1, -+i): the scalar phases of the two ends of a
// triplet cancel exactly, so only the free diag(1, -+i) parts are applied.
template <int NVEC, int ENT, bool CHI, bool DAG>
__device__ __forceinline__ void block_unit(cd (&a)[NVEC][4], const double2* __restrict__ tr,
                                           int flags, double* acc) {
  constexpr int C1A = CHI ? 2 : 1;  // amplitudes with control bit set: C1A, 3
  constexpr int T1A = CHI ? 1 : 2;  // amplitudes with target bit set:  T1A, 3
  constexpr int ROT_S = (ENT == AQC_ENT_CX) ? ROT_X : ROT_Z;
  const double2 t0 = tr[0], t1 = tr[1], t2 = tr[2], t3 = tr[3];
  if (!DAG) {
    if (flags & F_PRE) {
#pragma unroll
      for (int v = 0; v < NVEC; ++v) mul_mi(a[v][C1A]), mul_mi(a[v][3]);
    }
    if (ENT == AQC_ENT_CX) {
#pragma unroll
      for (int v = 0; v < NVEC; ++v) {
        const cd t = a[v][C1A];
        a[v][C1A] = a[v][3];
        a[v][3] = t;
      }
    } else if (ENT == AQC_ENT_CZ) {
#pragma unroll
      for (int v = 0; v < NVEC; ++v) a[v][3].x = -a[v][3].x, a[v][3].y = -a[v][3].y;
    } else {
      const double2 t4 = tr[4];  // (cos phi, sin phi)
      if (NVEC == 2) cdot_add(acc + 8, a[0][3], a[NVEC - 1][3]);  // factor -i, :972-975
#pragma unroll
      for (int v = 0; v < NVEC; ++v) mul_cs(a[v][3], t4.x, t4.y);
    }
    rot1q<NVEC, CHI, ROT_Y>(a, t0.x, t0.y, acc + 0);
    rot1q<NVEC, CHI, ROT_Z>(a, t1.x, t1.y, acc + 2);
    rot1q<NVEC, !CHI, ROT_Y>(a, t2.x, t2.y, acc + 4);
    rot1q<NVEC, !CHI, ROT_S>(a, t3.x, t3.y, acc + 6);
    if (flags & F_POST) {
#pragma unroll
      for (int v = 0; v < NVEC; ++v) mul_pi(a[v][T1A]), mul_pi(a[v][3]);
    }
  } else {
    if (flags & F_POST) {  // Rz_t(-pi/2) first (:793-794)
#pragma unroll
      for (int v = 0; v < NVEC; ++v) mul_mi(a[v][T1A]), mul_mi(a[v][3]);
    }
    rot1q<NVEC, !CHI, ROT_S>(a, t3.x, -t3.y, acc);
    rot1q<NVEC, !CHI, ROT_Y>(a, t2.x, -t2.y, acc);
    rot1q<NVEC, CHI, ROT_Z>(a, t1.x, -t1.y, acc);
    rot1q<NVEC, CHI, ROT_Y>(a, t0.x, -t0.y, acc);
    if (ENT == AQC_ENT_CX) {
#pragma unroll
      for (int v = 0; v < NVEC; ++v) {
        const cd t = a[v][C1A];
        a[v][C1A] = a[v][3];
        a[v][3] = t;
      }
    } else if (ENT == AQC_ENT_CZ) {
#pragma unroll
      for (int v = 0; v < NVEC; ++v) a[v][3].x = -a[v][3].x, a[v][3].y = -a[v][3].y;
    } else {
      const double2 t4 = tr[4];
#pragma unroll
      for (int v = 0; v < NVEC; ++v) mul_cs(a[v][3], t4.x, -t4.y);
    }
    if (flags & F_PRE) {  // Rz_c(+pi/2) last (:808-809)
#pragma unroll
      for (int v = 0; v < NVEC; ++v) mul_pi(a[v][C1A]), mul_pi(a[v][3]);
    }
  }
}
