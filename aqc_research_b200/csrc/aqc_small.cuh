// aqc_small.cuh -- elementwise helper kernels of the state-vector workspace.  Included by aqc_sv.cu.
#pragma once
// (cos, sin) table: half angles for rotations, full angle for the CPhase parameter.
__global__ void trig_kernel(const double* __restrict__ thetas, double2* __restrict__ trig,
                            long long total, int nthetas, int n3, int tpb) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int k = (int)(i % nthetas);
  const bool full = (tpb == 5) && k >= n3 && ((k - n3) % 5 == 4);
  double s, c;
  sincos(full ? thetas[i] : 0.5 * thetas[i], &s, &c);
  trig[i] = make_double2(c, s);
}

__global__ void set_basis_kernel(double2* __restrict__ v, long long size, long long stride,
                                 long long index) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= size) return;
  v[(long long)blockIdx.y * stride + i] = make_double2(i == index ? 1.0 : 0.0, 0.0);
}

struct SparseInit {
  long long index[8];
  double2 amp[8];
  int count;
};

// amplitudes of a few basis states on top of a zeroed vector (later entries win on equal indices)
__global__ void set_sparse_kernel(double2* __restrict__ v, long long stride, SparseInit s) {
  if (threadIdx.x == 0)
    for (int k = 0; k < s.count; ++k) v[(long long)blockIdx.x * stride + s.index[k]] = s.amp[k];
}

__global__ void set_identity_kernel(double2* __restrict__ v, long long size, long long stride,
                                    int log2_cols) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= size) return;
  const long long row = i >> log2_cols, col = i & ((1ll << log2_cols) - 1);
  v[(long long)blockIdx.y * stride + i] = make_double2(row == col ? 1.0 : 0.0, 0.0);
}

__global__ void gather_kernel(const double2* __restrict__ v, long long stride,
                              const long long* __restrict__ idx, int count,
                              double2* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  out[(size_t)blockIdx.y * count + i] = v[(long long)blockIdx.y * stride + idx[i]];
}

// out[b] += <a|b> partial sums (out must be zeroed before launch)
__global__ void vdot_kernel(const double2* __restrict__ a, const double2* __restrict__ b,
                            long long size, long long stride, double* __restrict__ out) {
  const double2* pa = a + (long long)blockIdx.y * stride;
  const double2* pb = b + (long long)blockIdx.y * stride;
  double re = 0.0, im = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < size;
       i += (long long)gridDim.x * blockDim.x) {
    const double2 x = pa[i], y = pb[i];
    re = fma(x.x, y.x, re);
    re = fma(x.y, y.y, re);
    im = fma(x.x, y.y, im);
    im = fma(-x.y, y.x, im);
  }
  for (int o = 16; o > 0; o >>= 1) {
    re += __shfl_xor_sync(0xffffffffu, re, o);
    im += __shfl_xor_sync(0xffffffffu, im, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(out + 2 * blockIdx.y, re);
    atomicAdd(out + 2 * blockIdx.y + 1, im);
  }
}

// splitmix64-based counter RNG -> U[0,1)
__device__ __forceinline__ double u01(unsigned long long seed, unsigned long long ctr) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (ctr + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}
__global__ void fill_random_kernel(double2* __restrict__ v, long long size, long long stride,
                                   unsigned long long seed, double* __restrict__ norm2) {
  double acc = 0.0;
  double2* p = v + (long long)blockIdx.y * stride;
  const unsigned long long s = seed + 0x632BE59BD9B4E019ull * blockIdx.y;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < size;
       i += (long long)gridDim.x * blockDim.x) {
    const double re = u01(s, 2ull * i), im = u01(s, 2ull * i + 1);
    p[i] = make_double2(re, im);
    acc = fma(re, re, acc);
    acc = fma(im, im, acc);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(norm2 + blockIdx.y, acc);
}
__global__ void scale_kernel(double2* __restrict__ v, long long size, long long stride,
                             const double* __restrict__ norm2) {
  double2* p = v + (long long)blockIdx.y * stride;
  const double f = rsqrt(norm2[blockIdx.y]);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < size;
       i += (long long)gridDim.x * blockDim.x) {
    double2 x = p[i];
    x.x *= f;
    x.y *= f;
    p[i] = x;
  }
}

