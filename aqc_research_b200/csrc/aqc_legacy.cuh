// aqc_legacy.cuh -- gate-by-gate register kernel (tile passes without the FP64 tensor pipe).  It is the
// only engine for states of fewer than 32 amplitudes and the cross-check of the dense engine
// (AQC_ENGINE=legacy).  Included by aqc_sv.cu after aqc_program.h.
#pragma once
// ------------------------------------------------------------------------------------------
// device code
// ------------------------------------------------------------------------------------------
// Sum 8 per-lane doubles over the warp with 7 + 2 shuffles: after the three halving steps lane L
// holds entry ((L>>4)&1)*4 + ((L>>3)&1)*2 + ((L>>2)&1) summed over lane bits 4,3,2.
__device__ __forceinline__ double warp_reduce8(const double* v, int lane, int& which) {
  double a[4], b[2], c;
  {
    const bool up = lane & 16;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const double send = up ? v[i] : v[i + 4];
      const double keep = up ? v[i + 4] : v[i];
      a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  {
    const bool up = lane & 8;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const double send = up ? a[i] : a[i + 2];
      const double keep = up ? a[i + 2] : a[i];
      b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
  {
    const bool up = lane & 4;
    const double send = up ? b[0] : b[1];
    const double keep = up ? b[1] : b[0];
    c = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  c += __shfl_xor_sync(0xffffffffu, c, 2);
  c += __shfl_xor_sync(0xffffffffu, c, 1);
  which = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
  return c;
}

struct PassArgs {
  const double2* src[2];  // [0] = w (NVEC == 2) or the single vector; [1] = z
  double2* dst[2];
  long long vec_stride;   // amplitudes between consecutive batch elements
  long long basis_index;  // >= 0: src[0] is the basis state |basis_index> (no load)
  const StageDesc* stages;
  const double2* trig;  // [batch][T]
  double* gacc;         // [batch][T] complex raw inner products
  int nthetas;
  PassDesc pd;
};

template <int NVEC, int ENT, bool DAG>
__global__ void __launch_bounds__(kThreads, (NVEC == 2 ? 3 : 4)) pass_kernel(const PassArgs A) {
  extern __shared__ double2 smem[];
  __shared__ long long s_hioff[32];
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int tb = A.pd.tb;
  const int tsize = 1 << tb;

  long long base = 0;
  {
    const unsigned long long tile = blockIdx.x;
    for (int k = 0; k < A.pd.nouter; ++k)
      base |= (long long)((tile >> k) & 1ull) << A.pd.outerpos[k];
  }
  // offsets: local index l = tid + 128*j  ->  global offset lo_off(tid) | hi_off(j)
  long long lo_off = 0;
  for (int k = 0; k < 7 && k < tb; ++k) lo_off |= (long long)((tid >> k) & 1) << A.pd.bitpos[k];
  if (tid < 32) {
    long long h = 0;
    for (int k = 7; k < tb; ++k) h |= (long long)((tid >> (k - 7)) & 1) << A.pd.bitpos[k];
    s_hioff[tid] = h;
  }
  __syncthreads();
  const long long boff = (long long)blockIdx.y * A.vec_stride + base;

#pragma unroll
  for (int v = 0; v < NVEC; ++v) {
    double2* sm = smem + (size_t)v * tsize;
    if (v == 0 && A.basis_index >= 0) {
      for (int l = tid; l < tsize; l += kThreads) {
        const long long g = base | lo_off | s_hioff[l >> 7];
        sm[l] = make_double2(g == A.basis_index ? 1.0 : 0.0, 0.0);
      }
    } else {
      const double2* __restrict__ src = A.src[v] + boff;
      for (int l = tid; l < tsize; l += kThreads) sm[l] = src[lo_off | s_hioff[l >> 7]];
    }
  }
  __syncthreads();

  const double2* __restrict__ trig = A.trig + (size_t)blockIdx.y * A.nthetas;
  double* gacc = A.gacc + (size_t)blockIdx.y * A.nthetas * 2;
  constexpr int NACC = (ENT == AQC_ENT_CP) ? 16 : 8;  // doubles per unit (padded to 8/16)
  const int nquads = tsize >> 2;

  for (int s = 0; s < A.pd.nstages; ++s) {
    const StageDesc* __restrict__ sd = A.stages + A.pd.stage0 + s;
    const int p = sd->p, q = sd->q, nunits = sd->nunits;
    const int mq = (1 << q) - 1, mp = (1 << p) - 1;
    double acc[kMaxUnits][NACC];
    if (NVEC == 2) {
#pragma unroll
      for (int u = 0; u < kMaxUnits; ++u)
#pragma unroll
        for (int k = 0; k < NACC; ++k) acc[u][k] = 0.0;
    }
    for (int j = tid; j < nquads; j += kThreads) {
      int i0 = ((j & ~mq) << 1) | (j & mq);
      i0 = ((i0 & ~mp) << 1) | (i0 & mp);
      const int i1 = i0 | (1 << q), i2 = i0 | (1 << p), i3 = i1 | (1 << p);
      cd a[NVEC][4];
#pragma unroll
      for (int v = 0; v < NVEC; ++v) {
        const double2* sm = smem + (size_t)v * tsize;
        const double2 x0 = sm[i0], x1 = sm[i1], x2 = sm[i2], x3 = sm[i3];
        a[v][0].x = x0.x, a[v][0].y = x0.y;
        a[v][1].x = x1.x, a[v][1].y = x1.y;
        a[v][2].x = x2.x, a[v][2].y = x2.y;
        a[v][3].x = x3.x, a[v][3].y = x3.y;
      }
#pragma unroll
      for (int u = 0; u < kMaxUnits; ++u) {
        if (u < nunits) {
          const int kind = sd->u[u].kind, flags = sd->u[u].flags;
          const double2* tr = trig + sd->u[u].theta;
          double* ac = (NVEC == 2) ? acc[u] : nullptr;
          switch (kind) {
            case U_FRONT_LO: front_unit<NVEC, false, DAG>(a, tr, ac); break;
            case U_FRONT_HI: front_unit<NVEC, true, DAG>(a, tr, ac); break;
            case U_BLOCK_CHI: block_unit<NVEC, ENT, true, DAG>(a, tr, flags, ac); break;
            case U_BLOCK_CLO: block_unit<NVEC, ENT, false, DAG>(a, tr, flags, ac); break;
            default: break;
          }
        }
      }
#pragma unroll
      for (int v = 0; v < NVEC; ++v) {
        double2* sm = smem + (size_t)v * tsize;
        sm[i0] = make_double2(a[v][0].x, a[v][0].y);
        sm[i1] = make_double2(a[v][1].x, a[v][1].y);
        sm[i2] = make_double2(a[v][2].x, a[v][2].y);
        sm[i3] = make_double2(a[v][3].x, a[v][3].y);
      }
    }
    if (NVEC == 2) {
#pragma unroll
      for (int u = 0; u < kMaxUnits; ++u) {
        if (u < nunits) {
          const int kind = sd->u[u].kind;
          const int nval = (kind == U_FRONT_LO || kind == U_FRONT_HI)
                               ? 6
                               : (ENT == AQC_ENT_CP ? 10 : 8);
          double* g = gacc + 2 * (size_t)sd->u[u].theta;
#pragma unroll
          for (int h = 0; h < NACC / 8; ++h) {
            int which;
            const double r = warp_reduce8(acc[u] + 8 * h, lane, which);
            which += 8 * h;
            if ((lane & 3) == 0 && which < nval) atomicAdd(g + which, r);
          }
        }
      }
    }
    __syncthreads();
  }

#pragma unroll
  for (int v = 0; v < NVEC; ++v) {
    const double2* sm = smem + (size_t)v * tsize;
    double2* __restrict__ dst = A.dst[v] + boff;
    for (int l = tid; l < tsize; l += kThreads) dst[lo_off | s_hioff[l >> 7]] = sm[l];
  }
}

