// aqc_sketch.cuh -- device side of the sketching-vector generators (included by aqc_sv.cu).
//
// Replaces the NumPy work inside RandomSketchingVectors / AlternatingSketchingVectors /
// EigenSketchingVectors.generate (aqc_research/model_sketching/sk_core.py:329-464):
//   Y = U @ X, U^H @ Omega        dense complex GEMMs (d x d) . (d x m)   -> zgemm_kernel (FP64 DMMA)
//   X, _ = np.linalg.qr(A)        thin QR of a d x m matrix               -> shifted Cholesky-QR3
//   column gathers of U, X = unit vectors, A -= B                         -> small kernels
// The objective and its gradient depend on X only through its column SPACE
// (Tr(X^H V^H U X) is invariant under X -> X W, W unitary), so any orthonormal basis of
// span(A) reproduces the reference's values; Cholesky-QR maps the whole factorisation onto GEMMs.
//
// All matrices are row-major complex128 (interleaved re, im).
#pragma once

constexpr int kGemmThreads = 256;
constexpr int kGemmTileM = 64, kGemmTileN = 64, kGemmTileK = 16;
constexpr int kGemmLdA = kGemmTileK + 4;   // doubles per smem row of the A planes (conflict-free DMMA loads)
constexpr int kGemmLdB = kGemmTileN + 4;   // ... of the B planes

struct GemmArgs {
  const double2* A;  // TRANS == 0: M x K (lda);  TRANS == 1: K x M (lda), used as conj-transpose
  const double2* B;  // K x N (ldb)
  double2* C;        // M x N (ldc); SPLITK > 1: accumulated with atomics (C zeroed by the caller)
  int M, N, K;
  long long lda, ldb, ldc;
  int ksplit;        // number of K slices (gridDim.z)
};

// C = op(A) . B with mma.sync.m8n8k4.f64: real form  Cr = Ar Br - Ai Bi,  Ci = Ar Bi + Ai Br.
// CTA tile 64 x 64, 8 warps, warp w owns rows 8w..8w+7 and all eight 8-column tiles.
template <int TRANS>
__global__ void __launch_bounds__(kGemmThreads) zgemm_kernel(const GemmArgs G) {
  __shared__ double sAr[kGemmTileM * kGemmLdA], sAi[kGemmTileM * kGemmLdA];
  __shared__ double sBr[kGemmTileK * kGemmLdB], sBi[kGemmTileK * kGemmLdB];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m0 = blockIdx.x * kGemmTileM, n0 = blockIdx.y * kGemmTileN;
  const int kchunk = (G.K + G.ksplit - 1) / G.ksplit;
  const int kbeg = blockIdx.z * kchunk, kend = min(G.K, kbeg + kchunk);
  double cr[8][2], ci[8][2];
#pragma unroll
  for (int t = 0; t < 8; ++t) cr[t][0] = cr[t][1] = ci[t][0] = ci[t][1] = 0.0;

  for (int k0 = kbeg; k0 < kend; k0 += kGemmTileK) {
    // A tile -> planes sA[i][k] (i local row, k local depth), zero padded
    for (int e = tid; e < kGemmTileM * kGemmTileK; e += kGemmThreads) {
      int i, k;
      if (TRANS) {
        i = e % kGemmTileM, k = e / kGemmTileM;  // consecutive threads walk along i (contiguous)
      } else {
        k = e % kGemmTileK, i = e / kGemmTileK;
      }
      double2 v = make_double2(0.0, 0.0);
      if (m0 + i < G.M && k0 + k < kend)
        v = TRANS ? G.A[(long long)(k0 + k) * G.lda + (m0 + i)] : G.A[(long long)(m0 + i) * G.lda + (k0 + k)];
      sAr[i * kGemmLdA + k] = v.x;
      sAi[i * kGemmLdA + k] = TRANS ? -v.y : v.y;
    }
    for (int e = tid; e < kGemmTileK * kGemmTileN; e += kGemmThreads) {
      const int j = e % kGemmTileN, k = e / kGemmTileN;
      double2 v = make_double2(0.0, 0.0);
      if (n0 + j < G.N && k0 + k < kend) v = G.B[(long long)(k0 + k) * G.ldb + (n0 + j)];
      sBr[k * kGemmLdB + j] = v.x;
      sBi[k * kGemmLdB + j] = v.y;
    }
    __syncthreads();
#pragma unroll
    for (int ks = 0; ks < kGemmTileK; ks += 4) {
      const int ai = (8 * warp + (lane >> 2)) * kGemmLdA + ks + (lane & 3);
      const double ar = sAr[ai], aim = sAi[ai], nai = -aim;
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const int bi = (ks + (lane & 3)) * kGemmLdB + 8 * t + (lane >> 2);
        const double br = sBr[bi], bim = sBi[bi];
        dmma884(cr[t][0], cr[t][1], ar, br);
        dmma884(cr[t][0], cr[t][1], nai, bim);
        dmma884(ci[t][0], ci[t][1], ar, bim);
        dmma884(ci[t][0], ci[t][1], aim, br);
      }
    }
    __syncthreads();
  }
  const int row = m0 + 8 * warp + (lane >> 2);
  if (row < G.M) {
#pragma unroll
    for (int t = 0; t < 8; ++t)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int col = n0 + 8 * t + 2 * (lane & 3) + h;
        if (col >= G.N) continue;
        double2* c = G.C + (long long)row * G.ldc + col;
        if (G.ksplit > 1) {
          atomicAdd(&c->x, cr[t][h]);
          atomicAdd(&c->y, ci[t][h]);
        } else {
          *c = make_double2(cr[t][h], ci[t][h]);
        }
      }
  }
}

// One CTA: G (m x m Hermitian positive definite, + shift on the diagonal) = R^H R (upper R),
// then Rinv = R^-1 (upper triangular), both in global memory.  Right-looking column Cholesky and a
// column-by-column back substitution; m is small (the number of sketching vectors).
__global__ void __launch_bounds__(256) chol_inv_kernel(double2* __restrict__ Gm, double2* __restrict__ Rinv, int m,
                                                      double shift, int* __restrict__ info) {
  const int tid = threadIdx.x;
  __shared__ double s_piv;
  // L = R^H stored in the lower triangle of Gm (row-major): G = L L^H
  for (int j = 0; j < m; ++j) {
    if (tid == 0) {
      const double d = Gm[(long long)j * m + j].x + shift;
      if (!(d > 0.0)) *info = j + 1;
      s_piv = sqrt(d > 0.0 ? d : 1.0);
    }
    __syncthreads();
    const double piv = s_piv;
    for (int i = j + tid; i < m; i += 256) {
      double2 v = Gm[(long long)i * m + j];
      if (i == j)
        v = make_double2(piv, 0.0);
      else
        v = make_double2(v.x / piv, v.y / piv);
      Gm[(long long)i * m + j] = v;
    }
    __syncthreads();
    // trailing update: G[i][k] -= L[i][j] conj(L[k][j]) for j < k <= i
    const int rem = m - j - 1;
    for (int e = tid; e < rem * rem; e += 256) {
      const int i = j + 1 + e / rem, k = j + 1 + e % rem;
      if (k > i) continue;
      const double2 a = Gm[(long long)i * m + j], b = Gm[(long long)k * m + j];
      double2 g = Gm[(long long)i * m + k];
      g.x -= a.x * b.x + a.y * b.y;
      g.y -= a.y * b.x - a.x * b.y;
      Gm[(long long)i * m + k] = g;
    }
    __syncthreads();
  }
  // R = L^H (upper).  Solve R Rinv = I column by column: thread c owns column c of Rinv.
  for (int c = tid; c < m; c += 256) {
    for (int i = m - 1; i >= 0; --i) {
      if (i > c) {
        Rinv[(long long)i * m + c] = make_double2(0.0, 0.0);
        continue;
      }
      double sx = (i == c) ? 1.0 : 0.0, sy = 0.0;
      for (int k = i + 1; k <= c; ++k) {
        // R[i][k] = conj(L[k][i])
        const double2 l = Gm[(long long)k * m + i], x = Rinv[(long long)k * m + c];
        sx -= l.x * x.x + l.y * x.y;
        sy -= l.x * x.y - l.y * x.x;
      }
      const double d = Gm[(long long)i * m + i].x;
      Rinv[(long long)i * m + c] = make_double2(sx / d, sy / d);
    }
  }
}

// sum of |a_ij|^2 (for the Cholesky-QR shift)
__global__ void norm2_kernel(const double2* __restrict__ a, long long count, double* __restrict__ out) {
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    const double2 v = a[i];
    acc = fma(v.x, v.x, fma(v.y, v.y, acc));
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

// dst -= src
__global__ void sub_kernel(double2* __restrict__ dst, const double2* __restrict__ src, long long count) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    double2 a = dst[i];
    const double2 b = src[i];
    a.x -= b.x, a.y -= b.y;
    dst[i] = a;
  }
}

// X[:, i] = e_{idx[i]},  Y[:, i] = U[:, idx[i]]   (X, Y: d x m row-major; U: d x d row-major)
__global__ void gather_cols_kernel(const double2* __restrict__ U, const long long* __restrict__ idx, int d, int m,
                                   double2* __restrict__ X, double2* __restrict__ Y) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)d * m) return;
  const int i = (int)(e % m);
  const long long r = e / m, c = idx[i];
  X[e] = make_double2(r == c ? 1.0 : 0.0, 0.0);
  Y[e] = U[r * d + c];
}
