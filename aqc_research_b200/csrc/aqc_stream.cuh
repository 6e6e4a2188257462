// aqc_stream.cuh -- persistent, warp-specialised sweep kernel of the dense-stage engine.
//
// One CTA per SM runs a whole SEQUENCE of tile passes (one launch per sweep; the passes are separated
// by a grid-wide barrier, so the kernel is launched cooperatively).  Inside a CTA
//   * kSGroups compute groups of kDWarps warps each work on their own tile: the stage loop of
//     aqc_dense.cuh (DMMA stage matrices, fused two-stage steps, per-stage 4x4 cross products),
//     synchronised with a NAMED barrier per group, never with the other groups;
//   * one producer warp moves every tile: 16-byte cp.async (LDGSTS) straight into the XOR-swizzled
//     slots, completion signalled on an mbarrier (cp.async.mbarrier.arrive.noinc), and the write-back
//     of finished tiles (LDS.128 -> STG.128).  Each group owns two tile buffers, so while it computes
//     on one the producer stores the previous result of the other and loads the group's next tile:
//     the compute warps never wait for global memory after the first tile of a pass.
// Tiles are dealt to the CTAs round robin (tile t -> CTA t mod gridDim.x) and inside a CTA to
// whichever group asks first, which keeps the SMs balanced when tiles / SMs is a small non-integer
// (n = 20: 1024 tiles over 148 SMs = 6.9 per SM instead of 3 waves of 444 CTAs).
//
// Reference semantics of what a pass computes: aqc_dense.cuh (core_operations.py:606-1019).
#pragma once

constexpr int kSGroups = 3;
constexpr int kSWarps = kSGroups * kDWarps + 4;  // + one warpgroup of producers (one warp per group, one idle)
constexpr int kSThreads = kSWarps * 32;
constexpr int kSBufBytes = 32768;  // one tile buffer: (w, z) tile of 2^10 amplitudes or one tile of 2^11
constexpr int kSOffPart = kSGroups * 2 * kSBufBytes;               // per-group partial sums [2][2][warps][32]
constexpr int kSPartDoubles = 2 * 2 * kDWarps * 32;                // per group
constexpr int kSOffHoff = kSOffPart + kSGroups * kSPartDoubles * 8;  // per producer 64 x int64: high-bit offsets
constexpr int kSOffTile = kSOffHoff + kSGroups * 64 * 8;           // [group][buffer] tile id (-1: no more tiles)
constexpr int kSOffNext = kSOffTile + 32;                          // CTA-local tile counter
constexpr int kSOffBar = kSOffTile + 64;                           // mbarriers full[group][buffer], done[group][buffer]
// Per-group running sums of the stage matrices M = sum z w^H of the first kSAccStages stages of a pass:
// they collect every tile the group processes and reach global memory (red.global.add.f64) once per
// pass instead of once per tile (n = 28: 590 tiles per group and pass); later stages of very long
// passes (one-pass matrix programs) keep the per-tile global adds.
constexpr int kSAccStages = 12;
constexpr int kSOffAcc = kSOffBar + 4 * kSGroups * 8 + 16;
constexpr int kSOffXdst = kSOffAcc + kSGroups * kSAccStages * 32 * 8;  // [vector][rank] push destinations
constexpr int kSSmemBytes = kSOffXdst + 2 * 16 * 8;
static_assert(kSSmemBytes <= 227 * 1024, "stream kernel shared memory");
// Registers: 896 threads at 72 registers each.  The two roles live in separate code regions (their own
// pass loops), which is what lets ptxas fit each of them into 72 without spills; rebalancing with
// setmaxnreg (80 for the compute warpgroups, 24 for the producers -- the CTA's own pool has to balance
// exactly, or setmaxnreg.inc blocks forever) was measured at HALF the speed: the producers, spilling at
// 24 registers, become the critical path (profiles/r02_stream_kernel.md).

struct StreamArgs {
  const PassDesc* passes;  // device copy of the program's passes
  int pass_begin, pass_end;
  int batch;
  int nstages_total;
  const double2* src[2];  // read by pass `pass_begin` ([0] = w or the single vector; [1] = z)
  double2* dst[2];        // written by every pass, read by the later ones
  long long vec_stride;
  // xcount > 0: src[0] of the first pass is the sparse vector sum_k xamp[k] |xindex[k]> (no load)
  long long xindex[8];
  const double2* xamp;    // device memory
  int xcount;
  const DLane* lanes;
  const double* umat;
  double* gm;
  unsigned long long* grid_bar;  // monotonic arrival counter (all launches of a workspace use one grid size)
  // Layout switch of a sharded state fused into the LAST pass of the range (xchg_world > 0): instead of
  // writing its tile back in place, the pass stores every 256-byte run where it belongs after the block
  // transpose over the ranks -- element (chunk c, offset o) of this rank goes to (chunk xchg_rank,
  // offset o) of rank c -- straight into the peers' memory (NVLink peer stores; xdst[v][c] is the
  // destination vector of rank c, mapped with CUDA IPC; c = xchg_rank is a local buffer).
  int nbuf;  // tile buffers per compute group: 2 (32 KiB each, double buffered) or 1 (64 KiB)
  int xchg_world, xchg_rank, xchg_shift;  // chunk = local index >> xchg_shift
  double2* xdst[2][16];
};

__device__ __forceinline__ void mbar_init(unsigned addr, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cp_async(unsigned addr) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ bool mbar_test(unsigned addr, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(addr), "r"(parity)
      : "memory");
  return ok != 0;
}
// Waits are bounded: a wait that lasts longer than ~10 s of SM clocks is a protocol bug, and a trapped
// kernel (an error the host sees) is better than a hung GPU.
constexpr long long kSWatchdogClocks = 20000000000ll;
__device__ __forceinline__ bool mbar_try(unsigned addr, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(addr), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned addr, unsigned parity) {
  if (mbar_try(addr, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try(addr, parity))
    if (clock64() - t0 > kSWatchdogClocks) __trap();
}
__device__ __forceinline__ void group_bar(int id) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(kDThreads) : "memory");
}
// 16-byte cp.async with zero fill: copies `bytes` (0 or 16) from gsrc, the rest of the slot is zeroed
__device__ __forceinline__ void cp_async16_zfill(unsigned saddr, const void* gsrc, unsigned bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(saddr), "l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void stg128(double2* p, const double2 v) {
  asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}

// All CTAs of the (cooperative) grid meet here; global writes made before are visible afterwards.
__device__ __forceinline__ void stream_grid_barrier(unsigned long long* bar, volatile int* s_next) {
  __syncthreads();
  if (threadIdx.x == 0) {
    *s_next = 0;
    __threadfence();
    const unsigned long long n = gridDim.x;
    const unsigned long long old = atomicAdd(bar, 1ull);
    const unsigned long long target = (old / n + 1ull) * n;
    unsigned long long seen;
    const long long t0 = clock64();
    do {
      asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(bar) : "memory");
      if (clock64() - t0 > kSWatchdogClocks) __trap();
    } while (seen < target);
    __threadfence();
  }
  __syncthreads();
}

template <int NVEC>
__global__ void __launch_bounds__(kSThreads, 1) dense_stream_kernel(const StreamArgs A) {
  extern __shared__ __align__(128) unsigned char s_raw[];
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const unsigned sm0 = (unsigned)__cvta_generic_to_shared(s_raw);
  volatile int* s_tile = reinterpret_cast<volatile int*>(s_raw + kSOffTile);
  volatile int* s_next = reinterpret_cast<volatile int*>(s_raw + kSOffNext);
  const unsigned bar0 = sm0 + kSOffBar;  // full[i] at bar0 + 8 i, done[i] at bar0 + 8 (2 G + i), i = 2 g + b
  if (tid == 0) {
    for (int i = 0; i < 2 * kSGroups; ++i) {
      mbar_init(bar0 + 8 * i, 32);
      mbar_init(bar0 + 8 * (2 * kSGroups + i), 1);
    }
    *s_next = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  double2** s_xdst = reinterpret_cast<double2**>(s_raw + kSOffXdst);
  if (tid < 32) s_xdst[tid] = A.xdst[tid >> 4][tid & 15];
  __syncthreads();
  const bool producer = warp >= kSGroups * kDWarps;
  const int g = producer ? warp - kSGroups * kDWarps : warp / kDWarps;  // group served / compute group
  const int wg = warp % kDWarps;                                        // warp inside the compute group
  // The two roles run the same pass loop in separate code regions.
  if (producer) {
    unsigned pdone = 0;  // phase parity of done[g][b] (bit b)
  for (int p = A.pass_begin; p < A.pass_end; ++p) {
    const PassDesc* __restrict__ pd = A.passes + p;
    const int tb = pd->tb, nstages = pd->nstages, stage0 = pd->stage0, nouter = pd->nouter;
    const int tsize = 1 << tb;
    const unsigned vbytes = (unsigned)tsize * 16u;
    const long long tiles_per_state = 1ll << nouter;
    const long long ntiles = tiles_per_state * A.batch;
    const bool first = p == A.pass_begin;
    (void)nstages, (void)stage0, (void)ntiles, (void)first;
    {
      // ---------------------------------------------------------------- producer warp of group g
      if (g < kSGroups) {
        long long* s_hoff = reinterpret_cast<long long*>(s_raw + kSOffHoff) + 64 * g;
        long long lo_off = 0;
        for (int k = 0; k < 5 && k < tb; ++k) lo_off |= (long long)((lane >> k) & 1) << pd->bitpos[k];
        const int nj = tsize >> 5;  // 16-byte chunks per lane and vector
        for (int j = lane; j < nj && j < 64; j += 32) {
          long long h = 0;
          for (int k = 5; k < tb && k < 11; ++k) h |= (long long)((j >> (k - 5)) & 1) << pd->bitpos[k];
          s_hoff[j] = h;
        }
        __syncwarp();
        // high-bit offset of chunk j of a lane (the table holds 64 entries; a 2^12 tile has 128 chunks)
        const long long hi12 = tb > 11 ? (1ll << pd->bitpos[11]) : 0ll;
        auto hoff = [&](int j) { return s_hoff[j & 63] + ((j >> 6) ? hi12 : 0ll); };
        const bool synth = first && A.xcount > 0;
        unsigned valid = 0;
        bool ended = false;
        // buffer b of this group: (re)fill it with the CTA's next tile, or tell the group that there is none
        auto refill = [&](int b) {
          if (ended) return;
          const int i = 2 * g + b;
          int k = 0;
          if (lane == 0) k = atomicAdd(const_cast<int*>(s_next), 1);
          k = __shfl_sync(0xffffffffu, k, 0);
          const long long t = (long long)blockIdx.x + (long long)k * gridDim.x;
          if (t >= ntiles) {
            ended = true;
            if (lane == 0) s_tile[i] = -1;
            __threadfence_block();
            mbar_arrive(bar0 + 8 * i);
            return;
          }
          const long long y = t >> nouter, x = t & (tiles_per_state - 1);
          long long base = 0;
          for (int kk = 0; kk < nouter; ++kk) base |= ((x >> kk) & 1ll) << pd->outerpos[kk];
          const long long boff = y * A.vec_stride + base;
          const unsigned sbuf = sm0 + (unsigned)i * kSBufBytes;
#pragma unroll
          for (int v = 0; v < NVEC; ++v) {
            const unsigned sv_ = sbuf + (unsigned)v * vbytes;
            if (v == 0 && synth) {
              for (int j = 0; j < nj; ++j) {
                const unsigned l = (unsigned)lane + 32u * (unsigned)j;
                const long long gi = base | lo_off | hoff(j);
                const double2* srcp = A.xamp;
                unsigned bytes = 0;
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)
                  if (kk < A.xcount && A.xindex[kk] == gi) srcp = A.xamp + kk, bytes = 16;
                cp_async16_zfill(sv_ + (dense_swz(l) << 4), srcp, bytes);
              }
            } else {
              const double2* __restrict__ src = (first ? A.src[v] : A.dst[v]) + boff + lo_off;
#pragma unroll 4
              for (int j = 0; j < nj; ++j) {
                const unsigned l = (unsigned)lane + 32u * (unsigned)j;
                cp_async16_s(sv_ + (dense_swz(l) << 4), src + hoff(j));
              }
            }
          }
          if (lane == 0) s_tile[i] = (int)t;
          __threadfence_block();
          mbar_arrive_cp_async(bar0 + 8 * i);
          valid |= 1u << b;
        };
        refill(0);
        if (A.nbuf > 1) refill(1);
        for (int b = 0; valid; b ^= 1) {
          if (!((valid >> b) & 1)) continue;
          const int i = 2 * g + b;
          mbar_wait(bar0 + 8 * (2 * kSGroups + i), (pdone >> b) & 1);
          pdone ^= 1u << b;
          {  // write the finished tile back
            const long long t = (long long)s_tile[i];
            const long long y = t >> nouter, x = t & (tiles_per_state - 1);
            long long base = 0;
            for (int kk = 0; kk < nouter; ++kk) base |= ((x >> kk) & 1ll) << pd->outerpos[kk];
            const long long boff = y * A.vec_stride + base;
            const unsigned sbuf = sm0 + (unsigned)i * kSBufBytes;
            const bool push = A.xchg_world > 0 && p + 1 == A.pass_end;
#pragma unroll
            for (int v = 0; v < NVEC; ++v) {
              const unsigned sv_ = sbuf + (unsigned)v * vbytes;
              double2* __restrict__ dst = A.dst[v] + boff + lo_off;
              for (int j0 = 0; j0 < nj; j0 += 4) {
                double2 val[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  const unsigned l = (unsigned)lane + 32u * (unsigned)(j0 + u);
                  if (j0 + u < nj) val[u] = lds128(sv_ + (dense_swz(l) << 4));
                }
                if (!push) {
#pragma unroll
                  for (int u = 0; u < 4; ++u)
                    if (j0 + u < nj) stg128(dst + hoff(j0 + u), val[u]);
                } else {
#pragma unroll
                  for (int u = 0; u < 4; ++u)
                    if (j0 + u < nj) {
                      const long long li = base + lo_off + hoff(j0 + u);  // local index (batch is 1)
                      const long long c = li >> A.xchg_shift, o = li & ((1ll << A.xchg_shift) - 1);
                      stg128(s_xdst[v * 16 + c] + (((long long)A.xchg_rank << A.xchg_shift) | o), val[u]);
                    }
                }
              }
            }
          }
          valid &= ~(1u << b);
          refill(b);
        }
        __threadfence();
      }
    }
    if (p + 1 < A.pass_end) stream_grid_barrier(A.grid_bar, s_next);
  }
  } else {
    unsigned pfull = 0;  // phase parity of full[g][b] (bit b)
    int step = 0;        // steps so far (parity selects the partial-sum buffer, also across tiles)
    for (int k = wg * 32 + lane; k < kSAccStages * 32; k += kDThreads)
      reinterpret_cast<double*>(s_raw + kSOffAcc)[(size_t)g * kSAccStages * 32 + k] = 0.0;
    group_bar(g + 1);
  for (int p = A.pass_begin; p < A.pass_end; ++p) {
    const PassDesc* __restrict__ pd = A.passes + p;
    const int tb = pd->tb, nstages = pd->nstages, stage0 = pd->stage0, nouter = pd->nouter;
    const int tsize = 1 << tb;
    const unsigned vbytes = (unsigned)tsize * 16u;
    const long long tiles_per_state = 1ll << nouter;
    const long long ntiles = tiles_per_state * A.batch;
    const bool first = p == A.pass_begin;
    (void)nstages, (void)stage0, (void)ntiles, (void)first;
    {
      // ---------------------------------------------------------------- compute group
      const int nit = tsize >> 5;
      double* part_g = reinterpret_cast<double*>(s_raw + kSOffPart) + (size_t)g * kSPartDoubles;
      double* acc_g = reinterpret_cast<double*>(s_raw + kSOffAcc) + (size_t)g * kSAccStages * 32;
      long long y_acc = -1;  // batch element the running sums belong to
      // running sums -> global memory (all warps of the group; the sums are left at zero)
      auto flush_acc = [&]() {
        if (NVEC != 2 || y_acc < 0) return;
        group_bar(g + 1);  // the last reducing warp has finished
        double* gdst = A.gm + ((size_t)y_acc * A.nstages_total + stage0) * 64 + lane;
        for (int sidx = wg; sidx < nstages && sidx < kSAccStages; sidx += kDWarps) {
          const double v = acc_g[sidx * 32 + lane];
          acc_g[sidx * 32 + lane] = 0.0;
          atomicAdd(gdst + (size_t)sidx * 64, v);
        }
        group_bar(g + 1);
      };
      for (int b = 0;; b ^= (A.nbuf - 1)) {
        const int i = 2 * g + b;
        mbar_wait(bar0 + 8 * i, (pfull >> b) & 1);
        pfull ^= 1u << b;
        const int t = s_tile[i];
        if (t < 0) break;
        const long long y = (long long)t >> nouter;
        if (y != y_acc) {
          flush_acc();
          y_acc = y;
        }
        const unsigned sm_u32 = sm0 + (unsigned)i * kSBufBytes;
        const size_t sbase = (size_t)y * A.nstages_total + stage0;
        const double* __restrict__ um = A.umat + sbase * 64 + lane;
        const uint2* __restrict__ lt =
            reinterpret_cast<const uint2*>(A.lanes + ((size_t)stage0 * kDWarps + wg) * 32 + lane);
        double* __restrict__ gmp = A.gm + sbase * 64 + lane;
        double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
        uint2 dl = make_uint2(0u, 0u);
        if (nstages > 0) {
          dl = lt[0];
          c0 = um[0];
          c1 = um[32];
          if (nstages > 1) {
            e0 = um[64];
            e1 = um[96];
          }
        }
        for (int s = 0; s < nstages; ++step) {
          const bool paired = (dl.x & kPairFirst) != 0;
          const bool swapped = (dl.x & kPairSwap) != 0;
          const int sA = s + (swapped ? 1 : 0);
          const int sB = 2 * s + 1 - sA;
          const int snext = s + (paired ? 2 : 1);
          const double ua0 = swapped ? e0 : c0, ua1 = swapped ? e1 : c1;
          const double ub0 = swapped ? c0 : e0, ub1 = swapped ? c1 : e1;
          double nc0 = 0.0, nc1 = 0.0, ne0 = 0.0, ne1 = 0.0;
          uint2 nl = make_uint2(0u, 0u);
          if (snext < nstages) {
            nl = lt[(size_t)snext * kDWarps * 32];
            if (paired) {
              nc0 = um[(size_t)snext * 64];
              nc1 = um[(size_t)snext * 64 + 32];
            } else {
              nc0 = e0, nc1 = e1;
            }
            if (snext + 1 < nstages) {
              ne0 = um[(size_t)(snext + 1) * 64];
              ne1 = um[(size_t)(snext + 1) * 64 + 32];
            }
          }
          const unsigned sl16 = (dl.x & 0x0fffu) << 4, so0 = (dl.x >> 16) << 3, so1 = (dl.y & 0xffffu) << 3;
          const unsigned sb16 = (dl.y >> 16) << 4;
          const bool hi = (lane & 4) != 0;
          double m0 = 0.0, m1 = 0.0, n0 = 0.0, n1 = 0.0;
          int j = 0;
          if (!paired) {
#pragma unroll 2
            for (int it = wg; it < nit; it += kDWarps, ++j) {
              const unsigned b16 = __shfl_sync(0xffffffffu, sb16, j);
              const unsigned la = sm_u32 + (b16 ^ sl16);
              const unsigned d0 = sm_u32 + (b16 ^ so0), d1 = sm_u32 + (b16 ^ so1);
              if (NVEC == 2) {
                const double2 w = lds128(la);
                const double2 z = lds128(la + vbytes);
                double w0 = 0.0, w1 = 0.0, z0 = 0.0, z1 = 0.0;
                dmma884(w0, w1, ua0, w.x);
                dmma884(z0, z1, ua0, z.x);
                dmma884(w0, w1, ua1, w.y);
                dmma884(z0, z1, ua1, z.y);
                dmma884(m0, m1, z0, w0);
                dmma884(m0, m1, z1, w1);
                sts64(d0, w0);
                sts64(d1, w1);
                sts64(d0 + vbytes, z0);
                sts64(d1 + vbytes, z1);
              } else {
                const double2 x = lds128(la);
                double x0 = 0.0, x1 = 0.0;
                dmma884(x0, x1, ua0, x.x);
                dmma884(x0, x1, ua1, x.y);
                sts64(d0, x0);
                sts64(d1, x1);
              }
            }
          } else {
#pragma unroll 2
            for (int it = wg; it < nit; it += kDWarps, ++j) {
              const unsigned b16 = __shfl_sync(0xffffffffu, sb16, j);
              const unsigned la = sm_u32 + (b16 ^ sl16);
              const unsigned d0 = sm_u32 + (b16 ^ so0), d1 = sm_u32 + (b16 ^ so1);
              if (NVEC == 2) {
                const double2 w = lds128(la);
                const double2 z = lds128(la + vbytes);
                double w0 = 0.0, w1 = 0.0, z0 = 0.0, z1 = 0.0;
                dmma884(w0, w1, ua0, w.x);
                dmma884(z0, z1, ua0, z.x);
                dmma884(w0, w1, ua1, w.y);
                dmma884(z0, z1, ua1, z.y);
                dmma884(m0, m1, z0, w0);
                dmma884(m0, m1, z1, w1);
                const double wr = __shfl_xor_sync(0xffffffffu, hi ? w0 : w1, 4);
                const double zr = __shfl_xor_sync(0xffffffffu, hi ? z0 : z1, 4);
                const double wx = hi ? wr : w0, wy = hi ? w1 : wr;
                const double zx = hi ? zr : z0, zy = hi ? z1 : zr;
                double v0 = 0.0, v1 = 0.0, y0 = 0.0, y1 = 0.0;
                dmma884(v0, v1, ub0, wx);
                dmma884(y0, y1, ub0, zx);
                dmma884(v0, v1, ub1, wy);
                dmma884(y0, y1, ub1, zy);
                dmma884(n0, n1, y0, v0);
                dmma884(n0, n1, y1, v1);
                sts64(d0, v0);
                sts64(d1, v1);
                sts64(d0 + vbytes, y0);
                sts64(d1 + vbytes, y1);
              } else {
                const double2 x = lds128(la);
                double x0 = 0.0, x1 = 0.0;
                dmma884(x0, x1, ua0, x.x);
                dmma884(x0, x1, ua1, x.y);
                const double xr = __shfl_xor_sync(0xffffffffu, hi ? x0 : x1, 4);
                const double xx = hi ? xr : x0, xy = hi ? x1 : xr;
                double v0 = 0.0, v1 = 0.0;
                dmma884(v0, v1, ub0, xx);
                dmma884(v0, v1, ub1, xy);
                sts64(d0, v0);
                sts64(d1, v1);
              }
            }
          }
          if (NVEC == 2) {
            double* part = part_g + (size_t)(step & 1) * 2 * kDWarps * 32;  // [set][warp][32]
            const double recv = __shfl_xor_sync(0xffffffffu, m1, 4);
            part[wg * 32 + lane] = hi ? (m0 - recv) : (m0 + recv);
            if (paired) {
              const double recv2 = __shfl_xor_sync(0xffffffffu, n1, 4);
              part[(kDWarps + wg) * 32 + lane] = hi ? (n0 - recv2) : (n0 + recv2);
            }
            group_bar(g + 1);
            const int red0 = step & (kDWarps - 1), red1 = (step + kDWarps / 2) & (kDWarps - 1);
            if (wg == red0 || (paired && wg == red1)) {
              const int set = (wg == red0) ? 0 : 1;
              const double* pp = part + (size_t)set * kDWarps * 32;
              double r0 = 0.0;
#pragma unroll
              for (int w = 0; w < kDWarps; ++w) r0 += pp[w * 32 + lane];
              const int sidx = set == 0 ? sA : sB;
              if (sidx < kSAccStages)
                acc_g[sidx * 32 + lane] += r0;  // this warp owns the slot until the next group barrier
              else
                atomicAdd(gmp + (size_t)sidx * 64, r0);
            }
          } else {
            group_bar(g + 1);
          }
          c0 = nc0, c1 = nc1, e0 = ne0, e1 = ne1, dl = nl;
          s = snext;
        }
        if (nstages == 0) group_bar(g + 1);
        if (wg == 0 && lane == 0) mbar_arrive(bar0 + 8 * (2 * kSGroups + i));
      }
      flush_acc();
    }
    if (p + 1 < A.pass_end) stream_grid_barrier(A.grid_bar, s_next);
  }
  }
}

// ------------------------------------------------------------------------------------------------
// Prologue and epilogue of a sweep (one small launch each, no host copies in between):
//   sweep_prologue_kernel : stage matrices U_s(theta) of one program straight from the angles (the host
//                           writes them into pinned, device-mapped memory -- no H2D copy, no (cos, sin)
//                           table), and the zeroing of the gradient accumulators;
//   grad_epilogue_kernel  : per-rotation inner products from the accumulated stage matrices
//                           (see dense_grad_kernel), the 0.5 / 0.5j / -i factors of the reference
//                           (core_operations.py:317-351, 972-975) and the write of the finished complex
//                           gradient into pinned host memory by the last CTA that completes.
// ------------------------------------------------------------------------------------------------
template <int ENT, bool DAG, int NVEC>
__device__ __forceinline__ void unit_from_theta(const UnitDesc& u, const double* __restrict__ th, cd (&a)[NVEC][4],
                                                double* acc) {
  const bool front = u.kind == U_FRONT_LO || u.kind == U_FRONT_HI;
  const int np = front ? 3 : (ENT == AQC_ENT_CP ? 5 : 4);
  double2 tr[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    if (k < np) {
      const double t = th[u.theta + k];
      double s, c;
      sincos(k == 4 ? t : 0.5 * t, &s, &c);  // full angle for the CPhase parameter
      tr[k] = make_double2(c, s);
    }
  }
  switch (u.kind) {
    case U_FRONT_LO: front_unit<NVEC, false, DAG>(a, tr, acc); break;
    case U_FRONT_HI: front_unit<NVEC, true, DAG>(a, tr, acc); break;
    case U_BLOCK_CHI: block_unit<NVEC, ENT, true, DAG>(a, tr, u.flags, acc); break;
    case U_BLOCK_CLO: block_unit<NVEC, ENT, false, DAG>(a, tr, u.flags, acc); break;
    default: break;
  }
}

struct PrologueArgs {
  const StageDesc* stages;
  int nstages, nthetas, batch;
  const double* thetas;  // [batch][nthetas], pinned host memory mapped into the device address space
  double* umat;          // [batch][nstages][64]
  double* zero0;         // two arrays to clear (stage-matrix sums, per-angle sums); may be null
  long long nzero0;
  double* zero1;
  long long nzero1;
};

template <int ENT, bool DAG>
__global__ void __launch_bounds__(128) sweep_prologue_kernel(const PrologueArgs A) {
  const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x, gsz = (long long)gridDim.x * blockDim.x;
  for (long long i = gt; i < A.nzero0; i += gsz) A.zero0[i] = 0.0;
  for (long long i = gt; i < A.nzero1; i += gsz) A.zero1[i] = 0.0;
  for (long long t = gt; t < (long long)A.batch * A.nstages * 4; t += gsz) {
    const int k = (int)(t & 3), s = (int)((t >> 2) % A.nstages), b = (int)((t >> 2) / A.nstages);
    const StageDesc sd = A.stages[s];
    const double* th = A.thetas + (size_t)b * A.nthetas;
    cd a[1][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[0][i].x = (i == k) ? 1.0 : 0.0, a[0][i].y = 0.0;
    for (int u = 0; u < sd.nunits; ++u) unit_from_theta<ENT, DAG, 1>(sd.u[u], th, a, nullptr);
    double* um = A.umat + ((size_t)b * A.nstages + s) * 64;
#pragma unroll
    for (int i = 0; i < 4; ++i) {  // U[i][k] = a[0][i]; DMMA A-fragment order (see dense_umat_kernel)
      um[(2 * i) * 4 + k] = a[0][i].x;
      um[32 + (2 * i) * 4 + k] = -a[0][i].y;
      um[(2 * i + 1) * 4 + k] = a[0][i].y;
      um[32 + (2 * i + 1) * 4 + k] = a[0][i].x;
    }
  }
}

struct EpilogueArgs {
  const StageDesc* stages;
  int nstages, nthetas, batch, n3, tpb;
  const double* thetas;
  const double* gm;        // [batch][nstages][64] accumulated stage matrices
  double* gacc;            // [batch][nthetas] complex raw sums (zeroed by the prologue)
  double* out;             // [batch][nthetas] complex gradient 0.5j <P w|z>, pinned host memory
  unsigned* ticket;        // completion counter (left at zero)
};

template <int ENT>
__global__ void __launch_bounds__(128) grad_epilogue_kernel(const EpilogueArgs A) {
  __shared__ int s_last;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < A.batch * A.nstages * 4) {
    const int r = t & 3, s = (t >> 2) % A.nstages, b = (t >> 2) / A.nstages;
    const StageDesc sd = A.stages[s];
    const double* th = A.thetas + (size_t)b * A.nthetas;
    const double* Mq = A.gm + ((size_t)b * A.nstages + s) * 64;
    // virtual quadruple r: w' = e_r, z' = M_out[:, r]; pull both back through the stage, then run it
    // forward with the reference's gate-by-gate accumulation (dense_grad_kernel)
    cd a[2][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      a[0][i].x = (i == r) ? 1.0 : 0.0, a[0][i].y = 0.0;
      a[1][i].x = Mq[(i << 3) | r];
      a[1][i].y = Mq[(i << 3) | 4 | r];
    }
    constexpr int NACC = (ENT == AQC_ENT_CP) ? 16 : 8;
    double acc[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
    for (int u = sd.nunits - 1; u >= 0; --u) unit_from_theta<ENT, true, 2>(sd.u[u], th, a, acc);
    double* g = A.gacc + (size_t)b * A.nthetas * 2;
    for (int u = 0; u < sd.nunits; ++u) {
#pragma unroll
      for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
      unit_from_theta<ENT, false, 2>(sd.u[u], th, a, acc);
      const int kind = sd.u[u].kind;
      const int nval = (kind == U_FRONT_LO || kind == U_FRONT_HI) ? 6 : ((kind == U_NONE) ? 0 : (ENT == AQC_ENT_CP ? 10 : 8));
      double* gu = g + 2 * (size_t)sd.u[u].theta;
#pragma unroll
      for (int k = 0; k < NACC; ++k)
        if (k < nval) atomicAdd(gu + k, acc[k]);
    }
  }
  // the CTA that finishes last converts the raw sums and hands the gradient to the host
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(A.ticket, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const int total = A.batch * A.nthetas;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int k = i % A.nthetas;
    const double re = __ldcg(A.gacc + 2 * (size_t)i), im = __ldcg(A.gacc + 2 * (size_t)i + 1);
    int kind;  // 0: Ry (0.5), 1: Rz / Rx (0.5j), 2: CPhase (-i)
    if (k < A.n3)
      kind = (k % 3 == 1) ? 0 : 1;
    else {
      const int q = (k - A.n3) % A.tpb;
      kind = (q == 4) ? 2 : ((q == 0 || q == 2) ? 0 : 1);
    }
    double2 v;
    if (kind == 0)
      v = make_double2(0.5 * re, 0.5 * im);
    else if (kind == 1)
      v = make_double2(-0.5 * im, 0.5 * re);
    else
      v = make_double2(im, -re);
    reinterpret_cast<double2*>(A.out)[i] = v;
  }
  if (threadIdx.x == 0) *A.ticket = 0u;
}

// hs[b][i] = v[b][idx[i]] written straight into pinned host memory
__global__ void gather_out_kernel(const double2* __restrict__ v, long long stride, const long long* __restrict__ idx,
                                  int count, double2* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  out[(size_t)blockIdx.y * count + i] = v[(long long)blockIdx.y * stride + idx[i]];
}
