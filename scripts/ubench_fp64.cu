// Micro-benchmark (development aid): FP64 issue rates on sm_100a.
//   DFMA with independent chains, mma.sync m8n8k4 / m16n8k4 / m16n8k8 / m16n8k16 .f64,
//   and a DFMA + DMMA mix (do they share the pipe?).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench_fp64.bin scripts/ubench_fp64.cu
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int CH>
__global__ void k_dfma(double* out, int iters, double a, double b) {
  double acc[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) acc[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1684(double (&c)[4], const double (&a)[2], double b) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int CH>
__global__ void k_dmma884(double* out, int iters, double a, double b) {
  double c[CH][2];
#pragma unroll
  for (int i = 0; i < CH; ++i) c[i][0] = threadIdx.x, c[i][1] = i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) dmma884(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// chained: output of one mma feeds the A operand of the next (as in a gate pipeline)
template <int CH>
__global__ void k_dmma884_chain(double* out, int iters, double b) {
  double c[CH][2];
#pragma unroll
  for (int i = 0; i < CH; ++i) c[i][0] = threadIdx.x, c[i][1] = i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      double d[2] = {0.0, 0.0};
      dmma884(d, c[i][0], b);
      dmma884(d, c[i][1], b);
      c[i][0] = d[0], c[i][1] = d[1];
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CH, int K>
__global__ void k_dmma16(double* out, int iters, double a, double b) {
  double c[CH][4];
#pragma unroll
  for (int i = 0; i < CH; ++i) c[i][0] = threadIdx.x, c[i][1] = i, c[i][2] = 1, c[i][3] = 2;
  double a2[2] = {a, a + 1}, a4[4] = {a, a + 1, a + 2, a + 3}, a8[8] = {a, a + 1, a + 2, a + 3, a, a, a, a};
  double b2[2] = {b, b + 1}, b4[4] = {b, b + 1, b + 2, b + 3};
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      if (K == 4) dmma1684(c[i], a2, b);
      if (K == 8) dmma1688(c[i], a4, b2);
      if (K == 16) dmma16816(c[i], a8, b4);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mix: CH dmma884 chains + CH2 dfma chains per iteration
template <int CH, int CF>
__global__ void k_mix(double* out, int iters, double a, double b) {
  double c[CH][2];
  double f[CF];
#pragma unroll
  for (int i = 0; i < CH; ++i) c[i][0] = threadIdx.x, c[i][1] = i;
#pragma unroll
  for (int i = 0; i < CF; ++i) f[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) dmma884(c[i], a, b);
#pragma unroll
    for (int i = 0; i < CF; ++i) f[i] = fma(f[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1];
#pragma unroll
  for (int i = 0; i < CF; ++i) s += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// shared-memory bandwidth: LDS.128 conflict-free, each thread reads 16 B per access
__global__ void k_lds(double* out, int iters) {
  __shared__ double2 sm[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_double2(i, 1.0);
  __syncthreads();
  double2 acc = make_double2(0, 0);
  int idx = threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const double2 v = sm[(idx + u * 256) & 2047];
      acc.x += v.x;
      acc.y += v.y;
    }
    idx = (idx + 32) & 2047;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y;
}

template <typename F>
static float time_it(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf("device %s, %d SMs, max clock %.0f MHz\n", prop.name, sms, clk_khz / 1e3);
  double* out;
  CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
  const int iters = 20000;
  for (int warps : {4, 8, 16, 32}) {
    const int threads = warps * 32;
    const int blocks = sms;
    float ms;
    ms = time_it([&] { k_dfma<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    printf("warps/SM %2d  DFMA x8 chains        : %7.2f TFLOP/s\n", warps,
           2.0 * 8 * iters * (double)threads * blocks / ms / 1e9);
    ms = time_it([&] { k_dmma884<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    printf("warps/SM %2d  DMMA m8n8k4 x8        : %7.2f TFLOP/s\n", warps,
           2.0 * 256 * 8 * iters * (double)warps * blocks / ms / 1e9);
    ms = time_it([&] { k_dmma884<2><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    printf("warps/SM %2d  DMMA m8n8k4 x2        : %7.2f TFLOP/s\n", warps,
           2.0 * 256 * 2 * iters * (double)warps * blocks / ms / 1e9);
    ms = time_it([&] { k_dmma884_chain<4><<<blocks, threads>>>(out, iters, 1e-3); });
    printf("warps/SM %2d  DMMA m8n8k4 chain x4  : %7.2f TFLOP/s\n", warps,
           2.0 * 256 * 2 * 4 * iters * (double)warps * blocks / ms / 1e9);
    ms = time_it([&] { k_dmma16<4, 4><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    printf("warps/SM %2d  DMMA m16n8k4 x4       : %7.2f TFLOP/s\n", warps,
           2.0 * 512 * 4 * iters * (double)warps * blocks / ms / 1e9);
    ms = time_it([&] { k_dmma16<4, 8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    printf("warps/SM %2d  DMMA m16n8k8 x4       : %7.2f TFLOP/s\n", warps,
           2.0 * 1024 * 4 * iters * (double)warps * blocks / ms / 1e9);
    ms = time_it([&] { k_dmma16<4, 16><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    printf("warps/SM %2d  DMMA m16n8k16 x4      : %7.2f TFLOP/s\n", warps,
           2.0 * 2048 * 4 * iters * (double)warps * blocks / ms / 1e9);
    ms = time_it([&] { k_mix<4, 8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    printf("warps/SM %2d  mix 4 DMMA + 8 DFMA   : %7.2f TFLOP/s (dmma %.2f + dfma %.2f)\n", warps,
           (2.0 * 256 * 4 * warps + 2.0 * 8 * threads) * iters * blocks / ms / 1e9,
           2.0 * 256 * 4 * warps * iters * (double)blocks / ms / 1e9,
           2.0 * 8 * threads * iters * (double)blocks / ms / 1e9);
    ms = time_it([&] { k_mix<2, 16><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    printf("warps/SM %2d  mix 2 DMMA + 16 DFMA  : %7.2f TFLOP/s (dmma %.2f + dfma %.2f)\n", warps,
           (2.0 * 256 * 2 * warps + 2.0 * 16 * threads) * iters * blocks / ms / 1e9,
           2.0 * 256 * 2 * warps * iters * (double)blocks / ms / 1e9,
           2.0 * 16 * threads * iters * (double)blocks / ms / 1e9);
    ms = time_it([&] { k_mix<4, 16><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    printf("warps/SM %2d  mix 4 DMMA + 16 DFMA  : %7.2f TFLOP/s (dmma %.2f + dfma %.2f)\n", warps,
           (2.0 * 256 * 4 * warps + 2.0 * 16 * threads) * iters * blocks / ms / 1e9,
           2.0 * 256 * 4 * warps * iters * (double)blocks / ms / 1e9,
           2.0 * 16 * threads * iters * (double)blocks / ms / 1e9);
    ms = time_it([&] { k_lds<<<blocks, threads>>>(out, iters); });
    printf("warps/SM %2d  LDS.128               : %7.2f B/clk/SM at max clock\n", warps,
           16.0 * 8 * iters * (double)threads / (ms * 1e-3 * clk_khz * 1e3));
  }
  cudaFree(out);
  return 0;
}
