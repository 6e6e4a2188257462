// Single-gate primitives of the reference's numeric core, one kernel per gate:
//   core_operations.py:46-603  (gate2x2_mul_vec, proj00/11, rx/ry/rz_mul_vec, dot_x/y/z, block_mul_vec,
//                               cx/cz/cp_mul_vec, derv_cphase_mul_vec)
//   core_op_matrix.py:32-477   (the same gates on (2^n, m) row-major matrices, x/y/z_dot_mat, derv_cphase)
// These are what the reference's own unit tests and tools call gate by gate.  The hot path does NOT go
// through here (it runs whole pair-runs per tile pass, aqc_dense.cuh); a call copies its host arrays
// over PCIe, runs a short list of gates on the device and copies the result back.
//
// A gate acts on an index "bit" given by its STRIDE: element i pairs with i + stride when
// (i / stride) is even.  Vectors: stride = 2^(n-1-pos) (core_operations.py:34-43 flips the bit order);
// matrices: stride = m 2^qubit with m columns, any m (the gate acts on the row index).

#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <mutex>
#include "../../include/aqc_b200.h"

int aqc_fail(int code, const char* fmt, ...);  // aqc_sv.cu
#define PCU(call)                                                                                   \
  do {                                                                                              \
    cudaError_t e_ = (call);                                                                        \
    if (e_ != cudaSuccess)                                                                          \
      return aqc_fail(AQC_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__,  \
                      __LINE__);                                                                    \
  } while (0)

namespace {

struct PrimGate {
  long long st;  // target stride
  long long sc;  // control stride (mode != 0)
  int mode;      // 0: plain gate; 1: controlled, identity where the control bit is 0; 2: controlled, ZERO there
  double2 g00, g01, g10, g11;
};

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }

// entries equal to exactly zero are skipped, as in gate2x2_mul_vec (core_operations.py:76-119): a
// projector must not turn an inf/nan of the discarded half into nan
__device__ __forceinline__ double2 lin2(double2 ga, double2 a, double2 gb, double2 b) {
  const bool za = ga.x == 0.0 && ga.y == 0.0, zb = gb.x == 0.0 && gb.y == 0.0;
  const double2 ta = za ? make_double2(0.0, 0.0) : cmul(ga, a);
  const double2 tb = zb ? make_double2(0.0, 0.0) : cmul(gb, b);
  return cadd(ta, tb);
}

__global__ void prim_gate_kernel(double2* __restrict__ v, long long npairs, PrimGate G) {
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npairs;
       p += (long long)gridDim.x * blockDim.x) {
    const long long i0 = (p / G.st) * 2 * G.st + (p % G.st), i1 = i0 + G.st;
    if (G.mode != 0 && ((i0 / G.sc) & 1) == 0) {
      if (G.mode == 2) v[i0] = v[i1] = make_double2(0.0, 0.0);
      continue;
    }
    const double2 a = v[i0], b = v[i1];
    v[i0] = lin2(G.g00, a, G.g01, b);
    v[i1] = lin2(G.g10, a, G.g11, b);
  }
}

// acc += sum_i conj((G w)_i) z_i
__global__ void prim_dot_kernel(const double2* __restrict__ w, const double2* __restrict__ z, long long npairs,
                                PrimGate G, double* __restrict__ acc) {
  double re = 0.0, im = 0.0;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npairs;
       p += (long long)gridDim.x * blockDim.x) {
    const long long i0 = (p / G.st) * 2 * G.st + (p % G.st), i1 = i0 + G.st;
    double2 a = w[i0], b = w[i1];
    if (G.mode != 0 && ((i0 / G.sc) & 1) == 0) {
      if (G.mode == 2) continue;
    } else {
      const double2 na = lin2(G.g00, a, G.g01, b), nb = lin2(G.g10, a, G.g11, b);
      a = na, b = nb;
    }
    const double2 za = z[i0], zb = z[i1];
    re += a.x * za.x + a.y * za.y + b.x * zb.x + b.y * zb.y;
    im += a.x * za.y - a.y * za.x + b.x * zb.y - b.y * zb.x;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    re += __shfl_xor_sync(0xffffffffu, re, o);
    im += __shfl_xor_sync(0xffffffffu, im, o);
  }
  __shared__ double s_re[8], s_im[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) s_re[warp] = re, s_im[warp] = im;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tr = 0.0, ti = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tr += s_re[i], ti += s_im[i];
    atomicAdd(acc, tr);
    atomicAdd(acc + 1, ti);
  }
}

struct Scratch {
  double2* buf[2] = {nullptr, nullptr};
  size_t cap[2] = {0, 0};
  double* acc = nullptr;
};
Scratch g_scratch[16];
std::mutex g_mutex;

int ensure(int device, int which, size_t count) {
  Scratch& s = g_scratch[device];
  if (s.cap[which] < count) {
    if (s.buf[which]) cudaFree(s.buf[which]);
    s.buf[which] = nullptr, s.cap[which] = 0;
    if (cudaMalloc((void**)&s.buf[which], count * sizeof(double2)) != cudaSuccess) {
      cudaGetLastError();
      return aqc_fail(AQC_ENOMEM, "cannot allocate %zu bytes of device scratch", count * sizeof(double2));
    }
    s.cap[which] = count;
  }
  if (!s.acc) PCU(cudaMalloc((void**)&s.acc, 2 * sizeof(double)));
  return AQC_OK;
}

int parse(const int64_t* strides, const int32_t* modes, const double* gates, int k, int64_t count, PrimGate* g) {
  g->st = strides[2 * k], g->sc = strides[2 * k + 1], g->mode = modes[k];
  if (g->st < 1 || count % (2 * g->st) != 0) return aqc_fail(AQC_EINVAL, "gate %d: bad target stride", k);
  if (g->mode < 0 || g->mode > 2) return aqc_fail(AQC_EINVAL, "gate %d: bad control mode", k);
  if (g->mode != 0 && (g->sc < 1 || count % (2 * g->sc) != 0 || g->sc == g->st))
    return aqc_fail(AQC_EINVAL, "gate %d: bad control stride", k);
  const double* m = gates + 8 * k;
  g->g00 = make_double2(m[0], m[1]), g->g01 = make_double2(m[2], m[3]);
  g->g10 = make_double2(m[4], m[5]), g->g11 = make_double2(m[6], m[7]);
  return AQC_OK;
}

int check_device(int device) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    cudaGetLastError();
    return aqc_fail(AQC_ENODEV, "no CUDA device visible: this library has no CPU path");
  }
  if (device < 0 || device >= ndev || device >= 16) return aqc_fail(AQC_EINVAL, "device %d out of range", device);
  return AQC_OK;
}

unsigned grid_for(long long npairs) {
  long long b = (npairs + 255) / 256;
  return (unsigned)(b < 1 ? 1 : (b > 148 * 16 ? 148 * 16 : b));
}

}  // namespace

extern "C" int aqc_prim_apply(int device, double* host, int64_t count, int num_gates, const int64_t* strides,
                              const int32_t* modes, const double* gates) {
  if (!host || !strides || !modes || !gates) return aqc_fail(AQC_EINVAL, "null pointer argument");
  if (count < 2 || num_gates < 1) return aqc_fail(AQC_EINVAL, "bad element or gate count");
  int rc = check_device(device);
  if (rc) return rc;
  std::lock_guard<std::mutex> lock(g_mutex);
  PCU(cudaSetDevice(device));
  rc = ensure(device, 0, (size_t)count);
  if (rc) return rc;
  double2* d = g_scratch[device].buf[0];
  PCU(cudaMemcpy(d, host, (size_t)count * sizeof(double2), cudaMemcpyHostToDevice));
  for (int k = 0; k < num_gates; ++k) {
    PrimGate g;
    rc = parse(strides, modes, gates, k, count, &g);
    if (rc) return rc;
    prim_gate_kernel<<<grid_for(count / 2), 256>>>(d, count / 2, g);
    PCU(cudaGetLastError());
  }
  PCU(cudaMemcpy(host, d, (size_t)count * sizeof(double2), cudaMemcpyDeviceToHost));
  return AQC_OK;
}

extern "C" int aqc_prim_dot(int device, const double* host_w, const double* host_z, int64_t count,
                            const int64_t* strides, const int32_t* modes, const double* gates, double* out) {
  if (!host_w || !host_z || !strides || !modes || !gates || !out) return aqc_fail(AQC_EINVAL, "null pointer argument");
  if (count < 2) return aqc_fail(AQC_EINVAL, "bad element count");
  int rc = check_device(device);
  if (rc) return rc;
  std::lock_guard<std::mutex> lock(g_mutex);
  PCU(cudaSetDevice(device));
  rc = ensure(device, 0, (size_t)count);
  if (!rc) rc = ensure(device, 1, (size_t)count);
  if (rc) return rc;
  Scratch& s = g_scratch[device];
  PrimGate g;
  rc = parse(strides, modes, gates, 0, count, &g);
  if (rc) return rc;
  PCU(cudaMemcpy(s.buf[0], host_w, (size_t)count * sizeof(double2), cudaMemcpyHostToDevice));
  PCU(cudaMemcpy(s.buf[1], host_z, (size_t)count * sizeof(double2), cudaMemcpyHostToDevice));
  PCU(cudaMemset(s.acc, 0, 2 * sizeof(double)));
  prim_dot_kernel<<<grid_for(count / 2), 256>>>(s.buf[0], s.buf[1], count / 2, g, s.acc);
  PCU(cudaGetLastError());
  PCU(cudaMemcpy(out, s.acc, 2 * sizeof(double), cudaMemcpyDeviceToHost));
  return AQC_OK;
}
