// aqc_cd.cuh -- coordinate descent for unitary AQC on the device (included by aqc_sv.cu).
//
// Replaces coord_descent_single_sweep (aqc_research/core_op_matrix.py:765-917): w = I, z = V^H U;
// before every rotation R_P(theta_k) the reference takes prod = <w|z>_F and grad = 0.5j <P w|z>_F
// over the whole (2^n x 2^n) matrices, makes a Newton step (or a clipped gradient step) for theta_k
// (:833-850), rotates z with the OLD angle and w with the NEW one -- 4 full-matrix reductions and 8
// full-matrix rotations per unit block, ~20 NumPy passes each.
//
// Here one CTA owns one start (grid = batch).  All rotations of a gate unit act on one bit pair, so
// both inner products of every rotation of the unit follow from ONE 4x4 complex matrix
// M = sum_quads z w^H taken after the entangler:  prod = Tr M,  <P w|z> = sum_ij conj(P_ij) M_ij,
// and a rotation pair (z <- A z, w <- B w) maps M -> A M B^H.  Per unit the kernel therefore makes
// one reduction pass (M), lets 16 lanes of warp 0 run the unit's 3-4 sequential angle updates on
// the 4x4 matrices (M, U_z = product of the old rotations, U_w = product of the new ones) with
// shuffles, and applies U_z / U_w in one update pass: 2 passes per unit instead of ~80.
#pragma once

struct CdUnit {
  int32_t kind;   // 0: front gate (Rz Ry Rz on qa), 1: unit block (control qa, target qb)
  int32_t qa, qb; // qubits; for a front gate qb is a passive partner
  int32_t theta;  // offset of the unit's first angle
};

struct CdArgs {
  double2* w;            // [batch][2^(2n)]
  double2* z;
  long long vec_stride;
  double* thetas;        // [batch][T], updated in place
  const CdUnit* units;
  int nunits, n, T;
  double* fobj;          // [batch]
};

constexpr int kCdThreads = 256;

struct c2 {
  double x, y;
};
__device__ __forceinline__ c2 cmul(c2 a, c2 b) { return {fma(-a.y, b.y, a.x * b.x), fma(a.y, b.x, a.x * b.y)}; }
__device__ __forceinline__ c2 cmulc(c2 a, c2 b) {  // a * conj(b)
  return {fma(a.y, b.y, a.x * b.x), fma(a.y, b.x, -a.x * b.y)};
}
__device__ __forceinline__ c2 cadd(c2 a, c2 b) { return {a.x + b.x, a.y + b.y}; }
__device__ __forceinline__ c2 cshfl(c2 a, int src) {
  return {__shfl_sync(0xffffffffu, a.x, src), __shfl_sync(0xffffffffu, a.y, src)};
}

// 2x2 rotation matrix entries R[r][c]; kind: 0 Ry, 1 Rz, 2 Rx (elementary_operations.py:143-291)
__device__ __forceinline__ c2 rot_entry(int kind, double cs, double sn, int r, int c) {
  if (kind == 1) return (r != c) ? c2{0.0, 0.0} : (r == 0 ? c2{cs, -sn} : c2{cs, sn});
  if (kind == 0) return (r == c) ? c2{cs, 0.0} : (r == 0 ? c2{-sn, 0.0} : c2{sn, 0.0});
  return (r == c) ? c2{cs, 0.0} : c2{0.0, -sn};
}

// X <- (R on the factor `mask` of the quad index) * X for the 4x4 matrix held one entry per lane
// (lane = 4 i + j)
__device__ __forceinline__ c2 left_mul(c2 x, int lane, int mask, int kind, double cs, double sn) {
  const int i = (lane >> 2) & 3, j = lane & 3;
  const int b = (i & mask) ? 1 : 0;
  const c2 other = cshfl(x, 4 * (i ^ mask) + j);
  return cadd(cmul(rot_entry(kind, cs, sn, b, b), x), cmul(rot_entry(kind, cs, sn, b, 1 - b), other));
}
// X <- X * (R on the factor)^H
__device__ __forceinline__ c2 right_mul_h(c2 x, int lane, int mask, int kind, double cs, double sn) {
  const int i = (lane >> 2) & 3, j = lane & 3;
  const int b = (j & mask) ? 1 : 0;
  const c2 other = cshfl(x, 4 * i + (j ^ mask));
  // (X B^H)[i][j] = X[i][j] conj(B[j][j]) + X[i][j^m] conj(B[j][j^m])
  return cadd(cmulc(x, rot_entry(kind, cs, sn, b, b)), cmulc(other, rot_entry(kind, cs, sn, b, 1 - b)));
}

// angle increment (core_op_matrix.py:833-850)
__device__ __forceinline__ double cd_delta(c2 prod, c2 grad, double dim2) {
  const double tol = 1.4901161193847656e-08;  // sqrt(eps)
  double d1 = -2.0 * (prod.x * grad.x + prod.y * grad.y) / dim2;
  const double d2 = (-2.0 * (grad.x * grad.x + grad.y * grad.y) + 0.5 * (prod.x * prod.x + prod.y * prod.y)) / dim2;
  double dt;
  if (d2 < tol) {
    d1 /= fmax(fabs(d1), 1.0);
    dt = -(3.14159265358979323846 / 16.0) * d1;
  } else {
    dt = -d1 / d2;
  }
  const double a = fabs(dt / (3.14159265358979323846 / 4.0));
  return a <= 1.0 ? dt : dt / a;
}

template <int ENT>
__global__ void __launch_bounds__(kCdThreads) cd_sweep_kernel(const CdArgs A) {
  __shared__ double s_red[kCdThreads / 32][32];
  __shared__ double s_m[32];
  __shared__ c2 s_u[2][16];  // [0]: U_w, [1]: U_z
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = A.n;
  const long long nquads = 1ll << (2 * n - 2);
  double2* __restrict__ wv = A.w + (long long)blockIdx.x * A.vec_stride;
  double2* __restrict__ zv = A.z + (long long)blockIdx.x * A.vec_stride;
  double* th = A.thetas + (size_t)blockIdx.x * A.T;
  const double dim2 = (double)(1ll << n) * (double)(1ll << n);
  c2 last_prod = {0.0, 0.0};

  for (int u = 0; u < A.nunits; ++u) {
    const CdUnit un = A.units[u];
    const int ba = n + un.qa, bb = n + un.qb;  // flat index bits (row bits sit above the column bits)
    const int hi = max(ba, bb), lo = min(ba, bb);
    const int mask_a = (ba == hi) ? 2 : 1, mask_b = 3 - mask_a;
    const long long mlo = (1ll << lo) - 1, mhi = (1ll << hi) - 1;

    // ---- pass 1: M[i][j] = sum_quads z_i conj(w_j), after the entangler
    double acc[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) acc[k] = 0.0;
    for (long long q = tid; q < nquads; q += kCdThreads) {
      long long i0 = ((q & ~mlo) << 1) | (q & mlo);
      i0 = ((i0 & ~mhi) << 1) | (i0 & mhi);
      long long idx[4] = {i0, i0 | (1ll << lo), i0 | (1ll << hi), i0 | (1ll << lo) | (1ll << hi)};
      double2 wq[4], zq[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) wq[a] = wv[idx[a]], zq[a] = zv[idx[a]];
      if (un.kind == 1) {
        if (ENT == AQC_ENT_CX) {  // control set: swap the target bit
          if (mask_a == 2) {
            double2 t = wq[2];
            wq[2] = wq[3], wq[3] = t;
            t = zq[2];
            zq[2] = zq[3], zq[3] = t;
          } else {
            double2 t = wq[1];
            wq[1] = wq[3], wq[3] = t;
            t = zq[1];
            zq[1] = zq[3], zq[3] = t;
          }
        } else {  // cz
          wq[3].x = -wq[3].x, wq[3].y = -wq[3].y;
          zq[3].x = -zq[3].x, zq[3].y = -zq[3].y;
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          // z_i conj(w_j)
          acc[2 * (4 * i + j)] = fma(zq[i].x, wq[j].x, fma(zq[i].y, wq[j].y, acc[2 * (4 * i + j)]));
          acc[2 * (4 * i + j) + 1] = fma(zq[i].y, wq[j].x, fma(-zq[i].x, wq[j].y, acc[2 * (4 * i + j) + 1]));
        }
    }
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      int which;
      const double r = warp_reduce8(acc + 8 * h, lane, which);
      if ((lane & 3) == 0) s_red[warp][8 * h + which] = r;
    }
    __syncthreads();
    if (tid < 32) {
      double r = 0.0;
#pragma unroll
      for (int w = 0; w < kCdThreads / 32; ++w) r += s_red[w][tid];
      s_m[tid] = r;
    }
    __syncthreads();

    // ---- the unit's sequential angle updates on 4x4 matrices: warp 0, lane = 4 i + j
    if (warp == 0) {
      const int l16 = lane & 15;
      const int i = l16 >> 2, j = l16 & 3;
      c2 M = {s_m[2 * l16], s_m[2 * l16 + 1]};
      c2 Uw = {i == j ? 1.0 : 0.0, 0.0}, Uz = Uw;
      const int nrot = un.kind == 0 ? 3 : 4;
      for (int r = 0; r < nrot; ++r) {
        int kind, mask, k;
        if (un.kind == 0) {  // Rz(t2) Ry(t1) Rz(t0) on qa
          k = 2 - r;
          kind = (r == 1) ? 0 : 1;
          mask = mask_a;
        } else {  // Ry(t0) Rz(t1) on the control, Ry(t2) Rs(t3) on the target
          k = r;
          kind = (r == 0 || r == 2) ? 0 : (r == 1 ? 1 : (ENT == AQC_ENT_CX ? 2 : 1));
          mask = r < 2 ? mask_a : mask_b;
        }
        // prod = Tr M;  <P w|z> = sum_ij conj(P_ij) M_ij
        c2 pc = (i == j) ? M : c2{0.0, 0.0};
        c2 gc = {0.0, 0.0};
        const int bi = (i & mask) ? 1 : 0;
        if (kind == 1) {
          if (i == j) gc = bi ? c2{-M.x, -M.y} : M;
        } else if (j == (i ^ mask)) {
          if (kind == 0)
            gc = bi ? c2{M.y, -M.x} : c2{-M.y, M.x};  // conj(P_ij) = +i (bit 0), -i (bit 1)
          else
            gc = M;
        }
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) {
          pc = cadd(pc, cshfl(pc, lane ^ o));
          gc = cadd(gc, cshfl(gc, lane ^ o));
        }
        const c2 grad = {-0.5 * gc.y, 0.5 * gc.x};  // 0.5j * gc
        const double told = th[un.theta + k];
        const double tnew = told + cd_delta(pc, grad, dim2);
        double so, co, sn, cn;
        sincos(0.5 * told, &so, &co);
        sincos(0.5 * tnew, &sn, &cn);
        M = left_mul(M, lane, mask, kind, co, so);
        M = right_mul_h(M, lane, mask, kind, cn, sn);
        Uz = left_mul(Uz, lane, mask, kind, co, so);
        Uw = left_mul(Uw, lane, mask, kind, cn, sn);
        if (lane == 0) th[un.theta + k] = tnew;
      }
      if (lane < 16) s_u[0][lane] = Uw, s_u[1][lane] = Uz;
      c2 pc = (i == j) ? M : c2{0.0, 0.0};
#pragma unroll
      for (int o = 1; o < 16; o <<= 1) pc = cadd(pc, cshfl(pc, lane ^ o));
      last_prod = pc;
    }
    __syncthreads();

    // ---- pass 2: w <- U_w E w, z <- U_z E z
    for (long long q = tid; q < nquads; q += kCdThreads) {
      long long i0 = ((q & ~mlo) << 1) | (q & mlo);
      i0 = ((i0 & ~mhi) << 1) | (i0 & mhi);
      long long idx[4] = {i0, i0 | (1ll << lo), i0 | (1ll << hi), i0 | (1ll << lo) | (1ll << hi)};
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        double2* __restrict__ p = v == 0 ? wv : zv;
        double2 x[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) x[a] = p[idx[a]];
        if (un.kind == 1) {
          if (ENT == AQC_ENT_CX) {
            if (mask_a == 2) {
              const double2 t = x[2];
              x[2] = x[3], x[3] = t;
            } else {
              const double2 t = x[1];
              x[1] = x[3], x[3] = t;
            }
          } else {
            x[3].x = -x[3].x, x[3].y = -x[3].y;
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          double re = 0.0, im = 0.0;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const c2 uu = s_u[v][4 * i + j];
            re = fma(uu.x, x[j].x, fma(-uu.y, x[j].y, re));
            im = fma(uu.x, x[j].y, fma(uu.y, x[j].x, im));
          }
          p[idx[i]] = make_double2(re, im);
        }
      }
    }
    __syncthreads();
  }
  if (tid == 0) {
    const double dim = (double)(1ll << n);
    const double ax = last_prod.x / dim, ay = last_prod.y / dim;
    A.fobj[blockIdx.x] = 1.0 - (ax * ax + ay * ay);
  }
}
