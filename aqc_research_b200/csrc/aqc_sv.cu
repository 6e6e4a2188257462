// aqc_sv.cu -- state-vector / column-batched objective-and-gradient engine for sm_100a.
//
// What this replaces (reference = qiskit-community/aqc-research, paths relative to its root):
//   v_mul_vec / v_dagger_mul_vec / grad_of_dot_product     aqc_research/core_operations.py:606,713,823
//   v_mul_mat / v_dagger_mul_mat / grad_of_matrix_dot_product   aqc_research/core_op_matrix.py:480,562,645
// The reference walks the circuit gate by gate and makes ~40-70 full NumPy passes over the
// 2^n vector per unit block.  Here the circuit STRUCTURE is compiled once (host side, below)
// into a short list of *tile passes*; each pass stages a tile of 2^tb amplitudes of w and z in
// shared memory, runs every gate whose qubits live inside the tile on register-resident
// amplitude quadruples (including the 0.5j<P w|z> inner products of the gradient sweep,
// reduced with warp shuffles) and writes the tile back once.  Angles enter only through a
// small (cos, sin) table built on the device, so a new theta costs one tiny launch.
//
// Design notes are in DESIGN.md; the C-ABI is declared in include/aqc_b200.h.

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/aqc_b200.h"
#include "aqc_gates.cuh"

// ------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CU(call)                                                                            \
  do {                                                                                      \
    cudaError_t e_ = (call);                                                                \
    if (e_ != cudaSuccess)                                                                  \
      return fail(AQC_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, \
                  __LINE__);                                                                \
  } while (0)

// shared with aqc_mps.cu
int aqc_fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

extern "C" const char* aqc_last_error(void) { return g_err.c_str(); }
extern "C" int aqc_version(void) { return 100; }
extern "C" int aqc_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

// ------------------------------------------------------------------------------------------
// compiled program: passes -> stages -> units
// ------------------------------------------------------------------------------------------
constexpr int kThreads = 128;  // threads per CTA of the pass kernel
constexpr int kMaxUnits = 3;   // units fused into one stage (a Trotter triplet)
constexpr int kStageUnits = 5; // dense-stage programs: two front gates + a triplet on one bit pair
constexpr int kMaxTileBits = 12;

struct UnitDesc {
  int32_t kind;
  int32_t flags;
  int32_t theta;  // index of the unit's first angle
  int32_t slot;   // first raw-gradient accumulator of this unit OCCURRENCE (5 per unit)
};
struct StageDesc {
  int32_t p, q;  // tile-local bit positions held in registers, p > q
  int32_t nunits;
  int32_t triplet;  // 1: Trotter triplet (ctrl hi / lo / hi, Rz(-pi/2) first, Rz(+pi/2) last)
  UnitDesc u[kStageUnits];  // legacy / scale-free programs use at most kMaxUnits of them
};
static_assert(sizeof(StageDesc) == 96, "StageDesc layout");

struct PassDesc {
  int32_t tb;       // tile bits
  int32_t nstages;  // stages in this pass
  int32_t stage0;   // first stage in the program's stage array
  int32_t nouter;   // number of index bits outside the tile
  uint8_t bitpos[16];    // global bit position of tile-local bit k
  uint8_t outerpos[48];  // global bit positions of the non-tile bits, ascending
};

struct Program {
  std::vector<PassDesc> passes;
  std::vector<StageDesc> stages;
  StageDesc* d_stages = nullptr;
  // sharded execution: passes [epoch_pass0[e], epoch_pass0[e+1]) need data layout epoch_layout[e]
  std::vector<int> epoch_pass0, epoch_layout;
  // gradient only: rotations in execution order (scaled-rotation bookkeeping)
  std::vector<int> sched_theta, sched_occ, sched_pass, pass_start, occ_theta;
  int *d_sched_theta = nullptr, *d_sched_occ = nullptr, *d_sched_pass = nullptr, *d_pass_start = nullptr,
      *d_occ_theta = nullptr;
};

struct HostUnit {
  int kind;  // 0 front, 1 block
  int qa;    // front qubit | control
  int qb;    // -1 | target
  int theta;
  int flags;
  int seq;  // index of the unit in forward circuit order (names its gradient accumulators)
};

struct aqc_circuit {
  int n = 0;
  int ent = 0;
  int trotter = 0;
  int nb = 0;  // blocks in full layers
  int half = 0;
  int tpb = 4;
  int nthetas = 0;
  std::vector<int> ctrl, targ;
};

// read-only accessors for aqc_mps.cu
int aqc_circ_n(const aqc_circuit* c) { return c->n; }
int aqc_circ_ent(const aqc_circuit* c) { return c->ent; }
int aqc_circ_trotter(const aqc_circuit* c) { return c->trotter; }
int aqc_circ_nb(const aqc_circuit* c) { return c->nb; }
int aqc_circ_half(const aqc_circuit* c) { return c->half; }
int aqc_circ_tpb(const aqc_circuit* c) { return c->tpb; }
int aqc_circ_nthetas(const aqc_circuit* c) { return c->nthetas; }
int aqc_circ_ctrl(const aqc_circuit* c, int i) { return c->ctrl[i]; }
int aqc_circ_targ(const aqc_circuit* c, int i) { return c->targ[i]; }

static void build_units(const aqc_circuit& c, bool reversed, std::vector<HostUnit>& out) {
  out.clear();
  for (int q = 0; q < c.n; ++q) out.push_back({0, q, -1, 3 * q, 0, q});
  const int total = c.nb + c.half;
  for (int i = 0; i < total; ++i) {
    const int im = c.nb > 0 ? i % c.nb : 0;
    int flags = 0;
    if (c.trotter != AQC_GENERIC) {
      if (i % 3 == 0) flags |= F_PRE;
      if (i % 3 == 2) flags |= F_POST;
    }
    out.push_back({1, c.ctrl[im], c.targ[im], 3 * c.n + c.tpb * im, flags, c.n + i});
  }
  if (reversed) std::reverse(out.begin(), out.end());
}

// Greedy tile-pass scheduler.  `units` is the gate-unit sequence in execution order; units on
// disjoint qubits commute, so a unit may run in the current pass iff all its qubits are inside
// the tile and none of them is touched by an earlier unit that had to be deferred.
// `units` carry PHYSICAL bit positions in qa / qb.  Passes are appended to `prog`.
// `max_units` caps the units of a stage; `merge_fronts` (dense-stage programs) lets a block unit
// join the front-gate stage that holds its qubits, so the front layer costs no stages of its own.
static void build_program_units(const std::vector<HostUnit>& units, int nbits, int tb_max,
                                int lowbits, Program& prog, int max_units = kMaxUnits,
                                bool merge_fronts = false) {
  const int qoff = 0;
  const size_t pass_begin = prog.passes.size();
  const int tb = std::min(nbits, tb_max);
  const int low = std::min(lowbits, tb);

  std::vector<char> done(units.size(), 0);
  size_t ndone = 0;
  while (ndone < units.size() || prog.passes.size() == pass_begin) {
    std::vector<char> intile(nbits, 0), blocked(nbits, 0);
    int ntile = 0;
    for (int b = 0; b < low; ++b) intile[b] = 1, ++ntile;
    std::vector<int> picked;
    for (size_t k = 0; k < units.size(); ++k) {
      if (done[k]) continue;
      const HostUnit& u = units[k];
      const int ba = u.qa + qoff, bb = u.kind ? u.qb + qoff : -1;
      const bool blk = blocked[ba] || (bb >= 0 && blocked[bb]);
      int need = (intile[ba] ? 0 : 1) + ((bb >= 0 && !intile[bb]) ? 1 : 0);
      if (!blk && ntile + need <= tb) {
        if (!intile[ba]) intile[ba] = 1, ++ntile;
        if (bb >= 0 && !intile[bb]) intile[bb] = 1, ++ntile;
        picked.push_back((int)k);
      } else {
        blocked[ba] = 1;
        if (bb >= 0) blocked[bb] = 1;
      }
    }
    // pad the tile with the lowest free bits (longer contiguous runs)
    for (int b = 0; b < nbits && ntile < tb; ++b)
      if (!intile[b]) intile[b] = 1, ++ntile;

    PassDesc pd;
    memset(&pd, 0, sizeof(pd));
    pd.tb = tb;
    pd.stage0 = (int)prog.stages.size();
    std::vector<int> local(nbits, -1);
    int kt = 0, ko = 0;
    for (int b = 0; b < nbits; ++b) {
      if (intile[b]) {
        local[b] = kt;
        pd.bitpos[kt++] = (uint8_t)b;
      } else {
        pd.outerpos[ko++] = (uint8_t)b;
      }
    }
    pd.nouter = ko;

    // group the picked units into stages (register-resident quads on one bit pair)
    struct Open {
      int a, b;  // tile-local bits (b == -1: partner still free, front-only stage)
      bool front;
      StageDesc sd;
    };
    std::vector<Open> open;
    std::vector<int> last(tb, -1);  // last stage that touched tile-local bit
    auto put = [&](Open& o, int kind, const HostUnit& u) {
      UnitDesc& d = o.sd.u[o.sd.nunits++];
      d.kind = kind;
      d.flags = u.flags;
      d.theta = u.theta;
      d.slot = 5 * u.seq;
    };
    // units are stored with *qubit roles*; the LO/HI kind is fixed up when the stage closes
    struct Pending {
      int stage;
      int slot;
      int la, lb;
      bool front;
    };
    std::vector<Pending> pend;
    for (int k : picked) {
      const HostUnit& u = units[k];
      const int la = local[u.qa + qoff];
      const int lb = u.kind ? local[u.qb + qoff] : -1;
      int s = -1;
      if (u.kind == 0) {
        // front gate: join an open front stage that has a free partner seat
        // (legal iff no stage created after it has touched this bit)
        if (merge_fronts && last[la] >= 0 && !open[last[la]].front &&
            open[last[la]].sd.nunits < max_units)
          s = last[la];  // reversed sweeps: the front gate follows the last stage on its qubit
        for (size_t i = 0; s < 0 && i < open.size(); ++i)
          if (open[i].front && open[i].b < 0 && open[i].a != la && open[i].sd.nunits < 2 &&
              (int)i > last[la]) {
            s = (int)i;
            open[i].b = la;
            break;
          }
        if (s < 0) {
          Open o;
          memset(&o.sd, 0, sizeof(o.sd));
          o.a = la;
          o.b = -1;
          o.front = true;
          open.push_back(o);
          s = (int)open.size() - 1;
        }
        last[la] = s;
      } else {
        const int sa = last[la], sb = last[lb];
        if (sa >= 0 && sa == sb && (merge_fronts || !open[sa].front) && open[sa].sd.nunits < max_units &&
            ((open[sa].a == la && open[sa].b == lb) || (open[sa].a == lb && open[sa].b == la))) {
          s = sa;
          open[sa].front = false;  // no further front gate may take a seat here
        } else if (merge_fronts && sa >= 0 && sa > sb && open[sa].front && open[sa].b < 0 &&
                   open[sa].a == la && open[sa].sd.nunits < max_units) {
          // single front gate on la with a free partner seat; everything on lb happened earlier
          s = sa;
          open[sa].b = lb;
          open[sa].front = false;
        } else if (merge_fronts && sb >= 0 && sb > sa && open[sb].front && open[sb].b < 0 &&
                   open[sb].a == lb && open[sb].sd.nunits < max_units) {
          s = sb;
          open[sb].b = la;
          open[sb].front = false;
        } else {
          Open o;
          memset(&o.sd, 0, sizeof(o.sd));
          o.a = la;
          o.b = lb;
          o.front = false;
          open.push_back(o);
          s = (int)open.size() - 1;
        }
        last[la] = last[lb] = s;
      }
      pend.push_back({s, open[s].sd.nunits, la, lb, u.kind == 0});
      put(open[s], U_NONE, u);
    }
    // A front stage whose partner seat stayed free gets any other tile bit as a passive partner.
    for (auto& o : open)
      if (o.b < 0) o.b = (o.a == 0) ? 1 : 0;
    for (auto& pe : pend) {
      Open& o = open[pe.stage];
      const int hi = std::max(o.a, o.b), lo = std::min(o.a, o.b);
      o.sd.p = hi;
      o.sd.q = lo;
      UnitDesc& d = o.sd.u[pe.slot];
      if (pe.front)
        d.kind = (pe.la == hi) ? U_FRONT_HI : U_FRONT_LO;
      else
        d.kind = (pe.la == hi) ? U_BLOCK_CHI : U_BLOCK_CLO;
      (void)lo;
    }
    for (auto& o : open) {
      const StageDesc& d = o.sd;
      o.sd.triplet = (d.nunits == 3 && d.u[0].kind == U_BLOCK_CHI && d.u[1].kind == U_BLOCK_CLO &&
                      d.u[2].kind == U_BLOCK_CHI && d.u[0].flags == F_PRE && d.u[1].flags == 0 &&
                      d.u[2].flags == F_POST)
                         ? 1
                         : 0;
      prog.stages.push_back(o.sd);
    }
    pd.nstages = (int)open.size();
    prog.passes.push_back(pd);
    for (int k : picked) done[k] = 1;
    ndone += picked.size();
    if (picked.empty() && ndone < units.size()) break;  // cannot happen (tb >= 2)
  }
}

static void build_program(const aqc_circuit& c, int qoff, int nbits, int tb_max, int lowbits,
                          bool reversed, Program& prog, int max_units = kMaxUnits,
                          bool merge_fronts = false) {
  std::vector<HostUnit> units;
  build_units(c, reversed, units);
  for (HostUnit& u : units) {
    u.qa += qoff;
    if (u.kind) u.qb += qoff;
  }
  prog.passes.clear();
  prog.stages.clear();
  prog.epoch_pass0.assign(1, 0);
  prog.epoch_layout.assign(1, 0);
  build_program_units(units, nbits, tb_max, lowbits, prog, max_units, merge_fronts);
}

// ---- global-qubit sharding (one state over 2^g GPUs) --------------------------------------------
// The top g index bits select the rank.  Two data layouts alternate:
//   layout A: qubits n-g..n-1 are global; qubits 0..g-1 sit on the TOP g local bits;
//   layout B: qubits 0..g-1 are global; qubits n-g..n-1 sit on the top g local bits;
// the other qubits occupy local bits 0..nl-g-1 (q -> q - g) in both.  Switching layouts is the
// block transpose new[rank c][chunk r] = old[rank r][chunk c] over chunks of 2^(nl-g) amplitudes
// (all-to-all over NVLink).  An epoch runs every gate unit that is executable without touching a
// global qubit (a light-cone trapezoid of the brick-wall circuit); then the layout is switched.
static int phys_bit(int q, int n, int g, int layout) {
  const int nl = n - g;
  if (q < g) return layout == 0 ? nl - g + q : -1;
  if (q >= n - g) return layout == 0 ? -1 : nl - g + (q - (n - g));
  return q - g;
}

static int build_program_sharded(const aqc_circuit& c, int g, int tb_max, int lowbits, bool reversed,
                                 Program& prog, std::string& err, int max_units = kMaxUnits,
                                 bool merge_fronts = false) {
  const int n = c.n, nl = n - g;
  if (nl - g < 2 || 2 * g > n - 2) {
    err = "too few qubits for this number of GPUs";
    return AQC_EINVAL;
  }
  std::vector<HostUnit> units;
  build_units(c, reversed, units);
  prog.passes.clear();
  prog.stages.clear();
  prog.epoch_pass0.clear();
  prog.epoch_layout.clear();
  std::vector<char> done(units.size(), 0);
  size_t ndone = 0;
  int layout = 0, idle = 0;
  while (ndone < units.size()) {
    std::vector<char> blocked(n, 0);
    std::vector<HostUnit> now;
    std::vector<size_t> ids;
    for (size_t k = 0; k < units.size(); ++k) {
      if (done[k]) continue;
      const HostUnit& u = units[k];
      const int pa = phys_bit(u.qa, n, g, layout);
      const int pb = u.kind ? phys_bit(u.qb, n, g, layout) : 0;
      const bool blk = blocked[u.qa] || (u.kind && blocked[u.qb]);
      if (!blk && pa >= 0 && pb >= 0) {
        HostUnit v = u;
        v.qa = pa;
        if (u.kind) v.qb = pb;
        now.push_back(v);
        ids.push_back(k);
      } else {
        blocked[u.qa] = 1;
        if (u.kind) blocked[u.qb] = 1;
      }
    }
    if (now.empty()) {
      if (++idle > 1) {
        err = "circuit cannot be scheduled over global qubits (a unit couples the lowest and highest qubits)";
        return AQC_EINVAL;
      }
      layout ^= 1;
      continue;
    }
    idle = 0;
    prog.epoch_pass0.push_back((int)prog.passes.size());
    prog.epoch_layout.push_back(layout);
    build_program_units(now, nl, tb_max, lowbits, prog, max_units, merge_fronts);
    for (size_t k : ids) done[k] = 1;
    ndone += ids.size();
    layout ^= 1;
  }
  if (prog.epoch_pass0.empty()) {  // circuit without units cannot happen (front layer), keep safe
    prog.epoch_pass0.push_back(0);
    prog.epoch_layout.push_back(0);
  }
  return AQC_OK;
}

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

#include "aqc_dense.cuh"
#include "aqc_cd.cuh"
#include "aqc_sketch.cuh"

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

// ------------------------------------------------------------------------------------------
// gradient sweep with scale-free rotations (see aqc_gates.cuh): prep, pass and finalize kernels
// ------------------------------------------------------------------------------------------
struct PrepArgs {
  const double* thetas;    // [batch][T]
  double2* par;            // [batch][T]  rotation parameters
  double* logf;            // [batch][T]  log2 |dropped scale| of each angle's rotation
  double* lbuf;            // [batch][J]  inclusive prefix of logf in execution order
  double* ebuf;            // [batch][npasses]  cumulative power-of-two renormalisation exponent
  double* rescale;         // [batch][npasses]  factor applied to the tile when a pass loads it
  double* dscale;          // [batch][nocc]  squared cumulative scale of each accumulator
  const int* sched_theta;  // [J] angle of the j-th rotation in execution order
  const int* sched_occ;    // [J] its accumulator
  const int* sched_pass;   // [J] its pass
  const int* pass_start;   // [npasses] first rotation of each pass
  int T, J, npasses, nocc, n3, tpb, cx;
};

// One CTA per angle set.  (1) rotation parameters; (2) prefix sums of log2|scale| along the
// execution order; (3) per-pass power-of-two renormalisation keeping stored ~ true magnitudes;
// (4) the squared scale each raw inner product has to be multiplied with.
__global__ void __launch_bounds__(256) prep_kernel(const PrepArgs A) {
  __shared__ double s_part[256];
  const int b = blockIdx.x, tid = threadIdx.x;
  const double* th = A.thetas + (size_t)b * A.T;
  double2* par = A.par + (size_t)b * A.T;
  double* logf = A.logf + (size_t)b * A.T;
  for (int k = tid; k < A.T; k += 256) {
    int kind;  // 0 Ry, 1 Rz, 2 Rx, 3 cphase
    if (k < A.n3)
      kind = (k % 3 == 1) ? 0 : 1;
    else {
      const int r = (k - A.n3) % A.tpb;
      kind = (r == 4) ? 3 : ((r == 0 || r == 2) ? 0 : (r == 1 ? 1 : (A.cx ? 2 : 1)));
    }
    double sn, cs;
    if (kind == 1 || kind == 3) {
      sincos(th[k], &sn, &cs);
      par[k] = make_double2(cs, sn);
      logf[k] = 0.0;
    } else {
      sincos(0.5 * th[k], &sn, &cs);
      // c-form unless cos is tiny: |t| <= 50 keeps the growth within a pass far from overflow and
      // makes the branch-free all-c-form stage path the common case
      if (fabs(cs) >= 0.02) {
        par[k] = make_double2(sn / cs, 0.0);
        logf[k] = log2(fabs(cs));
      } else {
        par[k] = make_double2(cs / sn, 1.0);
        logf[k] = log2(fabs(sn));
      }
    }
  }
  __syncthreads();
  // chunked inclusive scan over the execution order
  double* L = A.lbuf + (size_t)b * A.J;
  const int chunk = (A.J + 255) / 256;
  const int j0 = tid * chunk, j1 = min(A.J, j0 + chunk);
  double sum = 0.0;
  for (int j = j0; j < j1; ++j) sum += logf[A.sched_theta[j]];
  s_part[tid] = sum;
  __syncthreads();
  for (int o = 1; o < 256; o <<= 1) {
    const double add = (tid >= o) ? s_part[tid - o] : 0.0;
    __syncthreads();
    s_part[tid] += add;
    __syncthreads();
  }
  double run = (tid > 0) ? s_part[tid - 1] : 0.0;
  for (int j = j0; j < j1; ++j) {
    run += logf[A.sched_theta[j]];
    L[j] = run;
  }
  __syncthreads();
  double* E = A.ebuf + (size_t)b * A.npasses;
  double* rs = A.rescale + (size_t)b * A.npasses;
  for (int p = tid; p < A.npasses; p += 256) {
    const int js = A.pass_start[p];
    E[p] = (js > 0) ? rint(L[js - 1]) : 0.0;
  }
  __syncthreads();
  for (int p = tid; p < A.npasses; p += 256) rs[p] = exp2(E[p] - (p > 0 ? E[p - 1] : 0.0));
  double* ds = A.dscale + (size_t)b * A.nocc;
  for (int j = tid; j < A.J; j += 256) {
    const int occ = A.sched_occ[j];
    if (occ >= 0) ds[occ] = exp2(2.0 * (L[j] - E[A.sched_pass[j]]));
  }
}

// gacc[theta] += dscale[occ] * raw[occ]  (complex raw sums per accumulator -> per angle)
__global__ void finalize_kernel(const double* __restrict__ raw, const double* __restrict__ dscale,
                                const int* __restrict__ occ_theta, int nocc, int T,
                                double* __restrict__ gacc) {
  const int occ = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (occ >= nocc) return;
  const int k = occ_theta[occ];
  if (k < 0) return;
  const double f = dscale[(size_t)b * nocc + occ];
  atomicAdd(gacc + ((size_t)b * T + k) * 2, f * raw[((size_t)b * nocc + occ) * 2]);
  atomicAdd(gacc + ((size_t)b * T + k) * 2 + 1, f * raw[((size_t)b * nocc + occ) * 2 + 1]);
}

struct GradPassArgs {
  const double2* src[2];  // w, z
  double2* dst[2];
  long long vec_stride;
  long long basis_index;
  const StageDesc* stages;
  const double2* par;     // [batch][T]
  const double* rescale;  // [batch][npasses]
  double* gocc;           // [batch][nocc] complex raw sums
  int nthetas, nocc, npasses, pass_index;
  PassDesc pd;
};

constexpr int kParStages = 32;  // stages whose parameters are staged in shared memory

template <int ENT>
__global__ void __launch_bounds__(kThreads, 3) grad_pass_kernel(const GradPassArgs A) {
  extern __shared__ double2 smem[];
  __shared__ long long s_hioff[32];
  __shared__ double2 s_par[kParStages * kMaxUnits * 5];
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
  long long lo_off = 0;
  for (int k = 0; k < 7 && k < tb; ++k) lo_off |= (long long)((tid >> k) & 1) << A.pd.bitpos[k];
  if (tid < 32) {
    long long h = 0;
    for (int k = 7; k < tb; ++k) h |= (long long)((tid >> (k - 7)) & 1) << A.pd.bitpos[k];
    s_hioff[tid] = h;
  }
  const double2* __restrict__ par = A.par + (size_t)blockIdx.y * A.nthetas;
  const StageDesc* __restrict__ stages = A.stages + A.pd.stage0;
  const int nstages = A.pd.nstages;
  __syncthreads();
  const long long boff = (long long)blockIdx.y * A.vec_stride + base;
  const double rs = A.rescale[(size_t)blockIdx.y * A.npasses + A.pass_index];

#pragma unroll
  for (int v = 0; v < 2; ++v) {
    double2* sm = smem + (size_t)v * tsize;
    if (v == 0 && A.basis_index >= 0) {
      for (int l = tid; l < tsize; l += kThreads) {
        const long long g = base | lo_off | s_hioff[l >> 7];
        sm[l] = make_double2(g == A.basis_index ? rs : 0.0, 0.0);
      }
    } else {
      const double2* __restrict__ src = A.src[v] + boff;
      for (int l = tid; l < tsize; l += kThreads) {
        double2 x = src[lo_off | s_hioff[l >> 7]];
        x.x *= rs;
        x.y *= rs;
        sm[l] = x;
      }
    }
  }
  __syncthreads();

  double* gocc = A.gocc + (size_t)blockIdx.y * A.nocc * 2;
  constexpr int NACC = (ENT == AQC_ENT_CP) ? 16 : 8;
  const int nquads = tsize >> 2;

  for (int s = 0; s < nstages; ++s) {
    const StageDesc* __restrict__ sd = stages + s;
    if ((s % kParStages) == 0) {
      // rotation parameters of the next kParStages stages -> shared memory (the previous chunk is
      // no longer read: every stage ends with a barrier)
      const int cnt = min(kParStages, nstages - s);
      for (int i = tid; i < cnt * kMaxUnits * 5; i += kThreads) {
        const int ss = s + i / (kMaxUnits * 5), u = (i / 5) % kMaxUnits, k = i % 5;
        const int kind = stages[ss].u[u].kind;
        const int np = (kind == U_NONE || u >= stages[ss].nunits)
                           ? 0
                           : ((kind == U_FRONT_LO || kind == U_FRONT_HI) ? 3 : (ENT == AQC_ENT_CP ? 5 : 4));
        // an absent rotation counts as c-form (flag 0) for the all-c-form test below
        s_par[i] = (k < np) ? par[stages[ss].u[u].theta + k] : make_double2(0.0, 0.0);
      }
      __syncthreads();
    }
    const int p = sd->p, q = sd->q, nunits = sd->nunits;
    const int mq = (1 << q) - 1, mp = (1 << p) - 1;
    const bool triplet = sd->triplet != 0;
    double acc[kMaxUnits][NACC];
#pragma unroll
    for (int u = 0; u < kMaxUnits; ++u)
#pragma unroll
      for (int k = 0; k < NACC; ++k) acc[u][k] = 0.0;
    const double2* pu[kMaxUnits];
#pragma unroll
    for (int u = 0; u < kMaxUnits; ++u) pu[u] = s_par + ((s % kParStages) * kMaxUnits + u) * 5;
    // all scaled rotations of a triplet in c-form (the common case) -> branch-free code
    bool allc = false;
    if (triplet) {
      allc = pu[0][0].y == 0.0 && pu[0][2].y == 0.0 && pu[0][3].y == 0.0 && pu[1][0].y == 0.0 &&
             pu[1][2].y == 0.0 && pu[1][3].y == 0.0 && pu[2][0].y == 0.0 && pu[2][2].y == 0.0 &&
             pu[2][3].y == 0.0;
    }
    for (int j = tid; j < nquads; j += kThreads) {
      int i0 = ((j & ~mq) << 1) | (j & mq);
      i0 = ((i0 & ~mp) << 1) | (i0 & mp);
      const int i1 = i0 | (1 << q), i2 = i0 | (1 << p), i3 = i1 | (1 << p);
      cd a[2][4];
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        const double2* sm = smem + (size_t)v * tsize;
        const double2 x0 = sm[i0], x1 = sm[i1], x2 = sm[i2], x3 = sm[i3];
        a[v][0].x = x0.x, a[v][0].y = x0.y;
        a[v][1].x = x1.x, a[v][1].y = x1.y;
        a[v][2].x = x2.x, a[v][2].y = x2.y;
        a[v][3].x = x3.x, a[v][3].y = x3.y;
      }
      if (allc) {
        // straight-line Trotter triplet (cx): ctrl hi + Rz(-pi/2) | ctrl lo | ctrl hi + Rz(+pi/2)
        sblock_unit<AQC_ENT_CX, true, 1, 0, true>(a, pu[0], 0, acc[0]);
        sblock_unit<AQC_ENT_CX, false, 0, 0, true>(a, pu[1], 0, acc[1]);
        sblock_unit<AQC_ENT_CX, true, 0, 1, true>(a, pu[2], 0, acc[2]);
      } else if (triplet) {
        sblock_unit<AQC_ENT_CX, true, 1, 0>(a, pu[0], 0, acc[0]);
        sblock_unit<AQC_ENT_CX, false, 0, 0>(a, pu[1], 0, acc[1]);
        sblock_unit<AQC_ENT_CX, true, 0, 1>(a, pu[2], 0, acc[2]);
      } else {
#pragma unroll
        for (int u = 0; u < kMaxUnits; ++u) {
          if (u < nunits) {
            const int kind = sd->u[u].kind, flags = sd->u[u].flags;
            switch (kind) {
              case U_FRONT_LO: sfront_unit<false>(a, pu[u], acc[u]); break;
              case U_FRONT_HI: sfront_unit<true>(a, pu[u], acc[u]); break;
              case U_BLOCK_CHI: sblock_unit<ENT, true, -1, -1>(a, pu[u], flags, acc[u]); break;
              case U_BLOCK_CLO: sblock_unit<ENT, false, -1, -1>(a, pu[u], flags, acc[u]); break;
              default: break;
            }
          }
        }
      }
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        double2* sm = smem + (size_t)v * tsize;
        sm[i0] = make_double2(a[v][0].x, a[v][0].y);
        sm[i1] = make_double2(a[v][1].x, a[v][1].y);
        sm[i2] = make_double2(a[v][2].x, a[v][2].y);
        sm[i3] = make_double2(a[v][3].x, a[v][3].y);
      }
    }
#pragma unroll
    for (int u = 0; u < kMaxUnits; ++u) {
      if (u < nunits) {
        const int kind = sd->u[u].kind;
        const int nval = (kind == U_FRONT_LO || kind == U_FRONT_HI) ? 6 : (ENT == AQC_ENT_CP ? 10 : 8);
        double* g = gocc + 2 * (size_t)sd->u[u].slot;
#pragma unroll
        for (int h = 0; h < NACC / 8; ++h) {
          int which;
          const double r = warp_reduce8(acc[u] + 8 * h, lane, which);
          which += 8 * h;
          if ((lane & 3) == 0 && which < nval) atomicAdd(g + which, r);
        }
      }
    }
    __syncthreads();
  }

#pragma unroll
  for (int v = 0; v < 2; ++v) {
    const double2* sm = smem + (size_t)v * tsize;
    double2* __restrict__ dst = A.dst[v] + boff;
    for (int l = tid; l < tsize; l += kThreads) dst[lo_off | s_hioff[l >> 7]] = sm[l];
  }
}

// ------------------------------------------------------------------------------------------
// single-vector sweeps (V x, V^H x) with scale-free rotations
// ------------------------------------------------------------------------------------------
struct PrepApplyArgs {
  const double* thetas;    // [batch][T]
  double2* par;            // [batch][T]
  double* logf;            // [batch][T]
  double2* uph;            // [batch][T] unit-modulus part of each rotation's dropped scalar
  double* lbuf;            // [batch][J]
  double* ebuf;            // [batch][npasses]
  double* rescale;         // [batch][npasses]
  double2* kappa;          // [batch] scalar restoring the true amplitudes at the last store
  const int* sched_theta;  // [J]
  const int* pass_start;   // [npasses]
  int T, J, npasses, n3, tpb, cx, dagger;
};

__global__ void __launch_bounds__(256) prep_apply_kernel(const PrepApplyArgs A) {
  __shared__ double s_part[256];
  __shared__ double2 s_ph[256];
  const int b = blockIdx.x, tid = threadIdx.x;
  const double* th = A.thetas + (size_t)b * A.T;
  double2* par = A.par + (size_t)b * A.T;
  double* logf = A.logf + (size_t)b * A.T;
  double2* uph = A.uph + (size_t)b * A.T;
  const double sgn = A.dagger ? -1.0 : 1.0;
  for (int k = tid; k < A.T; k += 256) {
    int kind;  // 0 Ry, 1 Rz, 2 Rx, 3 cphase
    if (k < A.n3)
      kind = (k % 3 == 1) ? 0 : 1;
    else {
      const int r = (k - A.n3) % A.tpb;
      kind = (r == 4) ? 3 : ((r == 0 || r == 2) ? 0 : (r == 1 ? 1 : (A.cx ? 2 : 1)));
    }
    const double phi = sgn * th[k];
    double sn, cs;
    if (kind == 3) {
      sincos(phi, &sn, &cs);
      par[k] = make_double2(cs, sn);
      logf[k] = 0.0;
      uph[k] = make_double2(1.0, 0.0);
    } else if (kind == 1) {
      double sh, ch;
      sincos(0.5 * phi, &sh, &ch);
      // full-angle phase from the half angle: (ch + i sh)^2; dropped scalar e^{-i phi/2}
      par[k] = make_double2(fma(ch, ch, -sh * sh), 2.0 * ch * sh);
      logf[k] = 0.0;
      uph[k] = make_double2(ch, -sh);
    } else {
      sincos(0.5 * phi, &sn, &cs);
      if (fabs(cs) >= 0.02) {
        par[k] = make_double2(sn / cs, 0.0);
        logf[k] = log2(fabs(cs));
        uph[k] = make_double2(cs < 0.0 ? -1.0 : 1.0, 0.0);
      } else {
        par[k] = make_double2(cs / sn, 1.0);
        logf[k] = log2(fabs(sn));
        uph[k] = make_double2(sn < 0.0 ? -1.0 : 1.0, 0.0);
      }
    }
  }
  __syncthreads();
  double* L = A.lbuf + (size_t)b * A.J;
  const int chunk = (A.J + 255) / 256;
  const int j0 = tid * chunk, j1 = min(A.J, j0 + chunk);
  double sum = 0.0;
  double2 ph = make_double2(1.0, 0.0);
  for (int j = j0; j < j1; ++j) {
    const int k = A.sched_theta[j];
    sum += logf[k];
    const double2 u = uph[k];
    ph = make_double2(fma(-ph.y, u.y, ph.x * u.x), fma(ph.y, u.x, ph.x * u.y));
  }
  s_part[tid] = sum;
  s_ph[tid] = ph;
  __syncthreads();
  for (int o = 1; o < 256; o <<= 1) {
    const double add = (tid >= o) ? s_part[tid - o] : 0.0;
    __syncthreads();
    s_part[tid] += add;
    __syncthreads();
  }
  double run = (tid > 0) ? s_part[tid - 1] : 0.0;
  for (int j = j0; j < j1; ++j) {
    run += logf[A.sched_theta[j]];
    L[j] = run;
  }
  // product of the unit-modulus parts (tree)
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) {
      const double2 x = s_ph[tid], y = s_ph[tid + o];
      s_ph[tid] = make_double2(fma(-x.y, y.y, x.x * y.x), fma(x.y, y.x, x.x * y.y));
    }
    __syncthreads();
  }
  double* E = A.ebuf + (size_t)b * A.npasses;
  double* rs = A.rescale + (size_t)b * A.npasses;
  for (int p = tid; p < A.npasses; p += 256) {
    const int js = A.pass_start[p];
    E[p] = (js > 0) ? rint(L[js - 1]) : 0.0;
  }
  __syncthreads();
  for (int p = tid; p < A.npasses; p += 256) rs[p] = exp2(E[p] - (p > 0 ? E[p - 1] : 0.0));
  if (tid == 0) {
    const double ltot = (A.J > 0) ? L[A.J - 1] : 0.0;
    const double mag = exp2(ltot - (A.npasses > 0 ? E[A.npasses - 1] : 0.0));
    A.kappa[b] = make_double2(mag * s_ph[0].x, mag * s_ph[0].y);
  }
}

struct ApplyPassArgs {
  const double2* src;
  double2* dst;
  long long vec_stride;
  const StageDesc* stages;
  const double2* par;      // [batch][T]
  const double* rescale;   // [batch][npasses]
  const double2* kappa;    // [batch]; applied at the store of the LAST pass only
  int nthetas, npasses, pass_index, last;
  PassDesc pd;
};

template <int ENT, bool DAG>
__global__ void __launch_bounds__(kThreads, 4) apply_pass_kernel(const ApplyPassArgs A) {
  extern __shared__ double2 smem[];
  __shared__ long long s_hioff[32];
  __shared__ double2 s_par[kParStages * kMaxUnits * 5];
  const int tid = threadIdx.x;
  const int tb = A.pd.tb;
  const int tsize = 1 << tb;
  long long base = 0;
  {
    const unsigned long long tile = blockIdx.x;
    for (int k = 0; k < A.pd.nouter; ++k)
      base |= (long long)((tile >> k) & 1ull) << A.pd.outerpos[k];
  }
  long long lo_off = 0;
  for (int k = 0; k < 7 && k < tb; ++k) lo_off |= (long long)((tid >> k) & 1) << A.pd.bitpos[k];
  if (tid < 32) {
    long long h = 0;
    for (int k = 7; k < tb; ++k) h |= (long long)((tid >> (k - 7)) & 1) << A.pd.bitpos[k];
    s_hioff[tid] = h;
  }
  const double2* __restrict__ par = A.par + (size_t)blockIdx.y * A.nthetas;
  const StageDesc* __restrict__ stages = A.stages + A.pd.stage0;
  const int nstages = A.pd.nstages;
  __syncthreads();
  const long long boff = (long long)blockIdx.y * A.vec_stride + base;
  const double rs = A.rescale[(size_t)blockIdx.y * A.npasses + A.pass_index];
  {
    const double2* __restrict__ src = A.src + boff;
    for (int l = tid; l < tsize; l += kThreads) {
      double2 x = src[lo_off | s_hioff[l >> 7]];
      x.x *= rs;
      x.y *= rs;
      smem[l] = x;
    }
  }
  __syncthreads();
  const int nquads = tsize >> 2;
  for (int s = 0; s < nstages; ++s) {
    const StageDesc* __restrict__ sd = stages + s;
    if ((s % kParStages) == 0) {
      const int cnt = min(kParStages, nstages - s);
      for (int i = tid; i < cnt * kMaxUnits * 5; i += kThreads) {
        const int ss = s + i / (kMaxUnits * 5), u = (i / 5) % kMaxUnits, k = i % 5;
        const int kind = stages[ss].u[u].kind;
        const int np = (kind == U_NONE || u >= stages[ss].nunits)
                           ? 0
                           : ((kind == U_FRONT_LO || kind == U_FRONT_HI) ? 3 : (ENT == AQC_ENT_CP ? 5 : 4));
        s_par[i] = (k < np) ? par[stages[ss].u[u].theta + k] : make_double2(0.0, 0.0);
      }
      __syncthreads();
    }
    const int p = sd->p, q = sd->q, nunits = sd->nunits;
    const int mq = (1 << q) - 1, mp = (1 << p) - 1;
    for (int j = tid; j < nquads; j += kThreads) {
      int i0 = ((j & ~mq) << 1) | (j & mq);
      i0 = ((i0 & ~mp) << 1) | (i0 & mp);
      const int i1 = i0 | (1 << q), i2 = i0 | (1 << p), i3 = i1 | (1 << p);
      cd a[4];
      {
        const double2 x0 = smem[i0], x1 = smem[i1], x2 = smem[i2], x3 = smem[i3];
        a[0].x = x0.x, a[0].y = x0.y;
        a[1].x = x1.x, a[1].y = x1.y;
        a[2].x = x2.x, a[2].y = x2.y;
        a[3].x = x3.x, a[3].y = x3.y;
      }
#pragma unroll
      for (int u = 0; u < kMaxUnits; ++u) {
        if (u < nunits) {
          const int kind = sd->u[u].kind, flags = sd->u[u].flags;
          const double2* pu = s_par + ((s % kParStages) * kMaxUnits + u) * 5;
          switch (kind) {
            case U_FRONT_LO: afront_unit<false, DAG>(a, pu); break;
            case U_FRONT_HI: afront_unit<true, DAG>(a, pu); break;
            case U_BLOCK_CHI: ablock_unit<ENT, true, DAG>(a, pu, flags); break;
            case U_BLOCK_CLO: ablock_unit<ENT, false, DAG>(a, pu, flags); break;
            default: break;
          }
        }
      }
      smem[i0] = make_double2(a[0].x, a[0].y);
      smem[i1] = make_double2(a[1].x, a[1].y);
      smem[i2] = make_double2(a[2].x, a[2].y);
      smem[i3] = make_double2(a[3].x, a[3].y);
    }
    __syncthreads();
  }
  double2* __restrict__ dst = A.dst + boff;
  if (A.last) {
    const double2 kp = A.kappa[blockIdx.y];
    for (int l = tid; l < tsize; l += kThreads) {
      const double2 x = smem[l];
      dst[lo_off | s_hioff[l >> 7]] = make_double2(fma(-kp.y, x.y, kp.x * x.x), fma(kp.y, x.x, kp.x * x.y));
    }
  } else {
    for (int l = tid; l < tsize; l += kThreads) dst[lo_off | s_hioff[l >> 7]] = smem[l];
  }
}

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

// ------------------------------------------------------------------------------------------
// workspace
// ------------------------------------------------------------------------------------------
struct aqc_sv {
  aqc_circuit circ;
  int device = 0;
  int log2_cols = 0;
  int nbits = 0;
  int batch = 1;
  int nslots = 0;
  long long size = 0;  // amplitudes per state
  std::vector<double2*> slots;
  double* d_thetas = nullptr;
  double2* d_trig = nullptr;
  double* d_gacc = nullptr;
  double* d_scratch = nullptr;     // small outputs (gather / vdot)
  long long* d_idx = nullptr;
  std::vector<int64_t> idx_cached;  // host copy of what d_idx holds (gather_async)
  size_t idx_cap = 0, scratch_cap = 0;
  double* h_pinned = nullptr;  // pinned staging for thetas and small results
  size_t pinned_cap = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t tm0 = nullptr, tm1 = nullptr;  // user timer (aqc_sv_timer_*)
  float last_ms = 0.f;
  int last_launches = 0;
  Program prog_grad, prog_fwd, prog_dag;
  // scale-free gradient sweep (prep_kernel / grad_pass_kernel / finalize_kernel)
  bool legacy_grad = false;
  int nocc = 0, nsched = 0;
  double2* d_par = nullptr;
  double *d_logf = nullptr, *d_lbuf = nullptr, *d_ebuf = nullptr, *d_rescale = nullptr,
         *d_dscale = nullptr, *d_gocc = nullptr;
  // scale-free single-vector sweeps
  double2 *d_apar = nullptr, *d_uph = nullptr, *d_kappa = nullptr;
  double *d_albuf = nullptr, *d_aebuf = nullptr, *d_arescale = nullptr;
  // dense-stage engine (aqc_dense.cuh): DMMA sweeps; the default whenever the tile has >= 5 bits
  bool dense = false;
  bool grad_pending = false;  // aqc_sv_grad_begin enqueued, results not collected yet
  int num_sms = 148;
  DenseTables dt_grad, dt_fwd, dt_dag;
  double *d_umat = nullptr, *d_gm = nullptr;
  // coordinate descent (aqc_cd.cuh)
  CdUnit* d_cd_units = nullptr;
  int cd_nunits = 0;
  double* d_cd_fobj = nullptr;
  // sketching generators (aqc_sketch.cuh): dense target U (d x d) and m x m factorisation scratch
  double2* d_target = nullptr;
  double2 *d_gram = nullptr, *d_rinv = nullptr;
  int* d_info = nullptr;
  // global-qubit sharding (0 = single GPU)
  int g = 0, rank = 0;
  const double2* peer[64][16];  // peer[slot][rank]: IPC-mapped base pointers of the other ranks
};

static int ensure_pinned(aqc_sv* sv, size_t doubles) {
  if (sv->grad_pending) {  // an uncollected gradient sweep still owns the staging buffer: drop it
    CU(cudaStreamSynchronize(sv->stream));
    sv->grad_pending = false;
  }
  if (doubles <= sv->pinned_cap) return AQC_OK;
  if (sv->h_pinned) cudaFreeHost(sv->h_pinned);
  sv->h_pinned = nullptr;
  sv->pinned_cap = 0;
  CU(cudaMallocHost(&sv->h_pinned, doubles * sizeof(double)));
  sv->pinned_cap = doubles;
  return AQC_OK;
}
static int ensure_scratch(aqc_sv* sv, size_t doubles) {
  if (doubles <= sv->scratch_cap) return AQC_OK;
  if (sv->d_scratch) cudaFree(sv->d_scratch);
  sv->d_scratch = nullptr;
  sv->scratch_cap = 0;
  CU(cudaMalloc(&sv->d_scratch, doubles * sizeof(double)));
  sv->scratch_cap = doubles;
  return AQC_OK;
}
static int ensure_idx(aqc_sv* sv, size_t count) {
  if (count <= sv->idx_cap) return AQC_OK;
  if (sv->d_idx) cudaFree(sv->d_idx);
  sv->d_idx = nullptr;
  sv->idx_cached.clear();
  sv->idx_cap = 0;
  CU(cudaMalloc(&sv->d_idx, count * sizeof(long long)));
  sv->idx_cap = count;
  return AQC_OK;
}

template <int NVEC, int ENT, bool DAG>
static int launch_pass_t(aqc_sv* sv, const PassArgs& args) {
  const size_t smem = (size_t)NVEC * sizeof(double2) << args.pd.tb;
  static bool configured[8] = {false};  // per device
  if (!configured[sv->device & 7]) {
    CU(cudaFuncSetAttribute(pass_kernel<NVEC, ENT, DAG>,
                            cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)((size_t)NVEC * sizeof(double2) << kMaxTileBits)));
    configured[sv->device & 7] = true;
  }
  dim3 grid((unsigned)(1ull << args.pd.nouter), (unsigned)sv->batch);
  pass_kernel<NVEC, ENT, DAG><<<grid, kThreads, smem, sv->stream>>>(args);
  CU(cudaGetLastError());
  return AQC_OK;
}

template <int NVEC, bool DAG>
static int launch_pass_e(aqc_sv* sv, const PassArgs& args) {
  switch (sv->circ.ent) {
    case AQC_ENT_CX: return launch_pass_t<NVEC, AQC_ENT_CX, DAG>(sv, args);
    case AQC_ENT_CZ: return launch_pass_t<NVEC, AQC_ENT_CZ, DAG>(sv, args);
    default: return launch_pass_t<NVEC, AQC_ENT_CP, DAG>(sv, args);
  }
}

// NOTE: callers size the pinned staging buffer (ensure_pinned) BEFORE calling this, so that it
// is never re-allocated while the asynchronous copy below is in flight.
static int upload_thetas(aqc_sv* sv, const double* thetas) {
  const size_t tot = (size_t)sv->batch * sv->circ.nthetas;
  if (sv->pinned_cap < tot) return fail(AQC_EINVAL, "internal: pinned buffer too small");
  memcpy(sv->h_pinned, thetas, tot * sizeof(double));
  CU(cudaMemcpyAsync(sv->d_thetas, sv->h_pinned, tot * sizeof(double), cudaMemcpyHostToDevice,
                     sv->stream));
  const int thr = 128;
  trig_kernel<<<(unsigned)((tot + thr - 1) / thr), thr, 0, sv->stream>>>(
      sv->d_thetas, sv->d_trig, (long long)tot, sv->circ.nthetas, 3 * sv->circ.n, sv->circ.tpb);
  CU(cudaGetLastError());
  sv->last_launches += 1;
  return AQC_OK;
}

static int check_slot(const aqc_sv* sv, int slot) {
  if (!sv) return fail(AQC_EINVAL, "null workspace");
  if (slot < 0 || slot >= sv->nslots) return fail(AQC_EINVAL, "slot %d out of range", slot);
  return AQC_OK;
}

template <int ENT>
static int launch_grad_pass_t(aqc_sv* sv, const GradPassArgs& args) {
  const size_t smem = (size_t)2 * sizeof(double2) << args.pd.tb;
  static bool configured[8] = {false};
  if (!configured[sv->device & 7]) {
    CU(cudaFuncSetAttribute(grad_pass_kernel<ENT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)((size_t)2 * sizeof(double2) << (kMaxTileBits - 1))));
    configured[sv->device & 7] = true;
  }
  dim3 grid((unsigned)(1ull << args.pd.nouter), (unsigned)sv->batch);
  grad_pass_kernel<ENT><<<grid, kThreads, smem, sv->stream>>>(args);
  CU(cudaGetLastError());
  return AQC_OK;
}

// angle-dependent tables of the scale-free gradient sweep (thetas already uploaded) + zeroed sums
static int grad_prepare(aqc_sv* sv) {
  const Program& p = sv->prog_grad;
  PrepArgs a;
  a.thetas = sv->d_thetas;
  a.par = sv->d_par;
  a.logf = sv->d_logf;
  a.lbuf = sv->d_lbuf;
  a.ebuf = sv->d_ebuf;
  a.rescale = sv->d_rescale;
  a.dscale = sv->d_dscale;
  a.sched_theta = p.d_sched_theta;
  a.sched_occ = p.d_sched_occ;
  a.sched_pass = p.d_sched_pass;
  a.pass_start = p.d_pass_start;
  a.T = sv->circ.nthetas;
  a.J = sv->nsched;
  a.npasses = (int)p.passes.size();
  a.nocc = sv->nocc;
  a.n3 = 3 * sv->circ.n;
  a.tpb = sv->circ.tpb;
  a.cx = sv->circ.ent == AQC_ENT_CX;
  CU(cudaMemsetAsync(sv->d_gocc, 0, (size_t)sv->batch * sv->nocc * 2 * sizeof(double), sv->stream));
  prep_kernel<<<sv->batch, 256, 0, sv->stream>>>(a);
  CU(cudaGetLastError());
  sv->last_launches += 1;
  return AQC_OK;
}

// per-accumulator raw sums -> per-angle sums in d_gacc
static int grad_collect(aqc_sv* sv) {
  const size_t tot = (size_t)sv->batch * sv->circ.nthetas;
  CU(cudaMemsetAsync(sv->d_gacc, 0, tot * 2 * sizeof(double), sv->stream));
  finalize_kernel<<<dim3((sv->nocc + 127) / 128, sv->batch), 128, 0, sv->stream>>>(
      sv->d_gocc, sv->d_dscale, sv->prog_grad.d_occ_theta, sv->nocc, sv->circ.nthetas, sv->d_gacc);
  CU(cudaGetLastError());
  sv->last_launches += 1;
  return AQC_OK;
}

static int run_grad_program(aqc_sv* sv, const double2* src0, long long basis, const double2* src1,
                            double2* dst0, double2* dst1, int pass_begin, int pass_end) {
  const Program& prog = sv->prog_grad;
  GradPassArgs a;
  memset(&a, 0, sizeof(a));
  a.vec_stride = sv->size;
  a.stages = prog.d_stages;
  a.par = sv->d_par;
  a.rescale = sv->d_rescale;
  a.gocc = sv->d_gocc;
  a.nthetas = sv->circ.nthetas;
  a.nocc = sv->nocc;
  a.npasses = (int)prog.passes.size();
  if (pass_end < 0) pass_end = (int)prog.passes.size();
  for (int i = pass_begin; i < pass_end; ++i) {
    a.pd = prog.passes[i];
    a.pass_index = i;
    a.src[0] = (i == pass_begin) ? src0 : dst0;
    a.src[1] = (i == pass_begin) ? src1 : dst1;
    a.dst[0] = dst0;
    a.dst[1] = dst1;
    a.basis_index = (i == pass_begin) ? basis : -1;
    int rc;
    switch (sv->circ.ent) {
      case AQC_ENT_CX: rc = launch_grad_pass_t<AQC_ENT_CX>(sv, a); break;
      case AQC_ENT_CZ: rc = launch_grad_pass_t<AQC_ENT_CZ>(sv, a); break;
      default: rc = launch_grad_pass_t<AQC_ENT_CP>(sv, a);
    }
    if (rc) return rc;
    sv->last_launches += 1;
  }
  return AQC_OK;
}

template <int ENT, bool DAG>
static int launch_apply_pass_t(aqc_sv* sv, const ApplyPassArgs& args) {
  const size_t smem = sizeof(double2) << args.pd.tb;
  static bool configured[8] = {false};
  if (!configured[sv->device & 7]) {
    CU(cudaFuncSetAttribute(apply_pass_kernel<ENT, DAG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)(sizeof(double2) << kMaxTileBits)));
    configured[sv->device & 7] = true;
  }
  dim3 grid((unsigned)(1ull << args.pd.nouter), (unsigned)sv->batch);
  apply_pass_kernel<ENT, DAG><<<grid, kThreads, smem, sv->stream>>>(args);
  CU(cudaGetLastError());
  return AQC_OK;
}

// angle-dependent tables of a single-vector sweep (thetas already uploaded)
static int apply_prepare(aqc_sv* sv, bool dagger) {
  const Program& p = dagger ? sv->prog_dag : sv->prog_fwd;
  PrepApplyArgs a;
  a.thetas = sv->d_thetas;
  a.par = sv->d_apar;
  a.logf = sv->d_logf;
  a.uph = sv->d_uph;
  a.lbuf = sv->d_albuf;
  a.ebuf = sv->d_aebuf;
  a.rescale = sv->d_arescale;
  a.kappa = sv->d_kappa;
  a.sched_theta = p.d_sched_theta;
  a.pass_start = p.d_pass_start;
  a.T = sv->circ.nthetas;
  a.J = (int)p.sched_theta.size();
  a.npasses = (int)p.passes.size();
  a.n3 = 3 * sv->circ.n;
  a.tpb = sv->circ.tpb;
  a.cx = sv->circ.ent == AQC_ENT_CX;
  a.dagger = dagger ? 1 : 0;
  prep_apply_kernel<<<sv->batch, 256, 0, sv->stream>>>(a);
  CU(cudaGetLastError());
  sv->last_launches += 1;
  return AQC_OK;
}

// passes [pass_begin, pass_end) of a single-vector program; `final` marks the end of the whole
// sweep (the stored scalar is multiplied back by the last pass)
static int run_apply_program(aqc_sv* sv, bool dagger, const double2* src, double2* dst,
                             int pass_begin, int pass_end) {
  const Program& prog = dagger ? sv->prog_dag : sv->prog_fwd;
  ApplyPassArgs a;
  memset(&a, 0, sizeof(a));
  a.vec_stride = sv->size;
  a.stages = prog.d_stages;
  a.par = sv->d_apar;
  a.rescale = sv->d_arescale;
  a.kappa = sv->d_kappa;
  a.nthetas = sv->circ.nthetas;
  a.npasses = (int)prog.passes.size();
  if (pass_end < 0) pass_end = (int)prog.passes.size();
  for (int i = pass_begin; i < pass_end; ++i) {
    a.pd = prog.passes[i];
    a.pass_index = i;
    a.last = (i + 1 == (int)prog.passes.size()) ? 1 : 0;
    a.src = (i == pass_begin) ? src : dst;
    a.dst = dst;
    int rc;
    if (dagger) {
      switch (sv->circ.ent) {
        case AQC_ENT_CX: rc = launch_apply_pass_t<AQC_ENT_CX, true>(sv, a); break;
        case AQC_ENT_CZ: rc = launch_apply_pass_t<AQC_ENT_CZ, true>(sv, a); break;
        default: rc = launch_apply_pass_t<AQC_ENT_CP, true>(sv, a);
      }
    } else {
      switch (sv->circ.ent) {
        case AQC_ENT_CX: rc = launch_apply_pass_t<AQC_ENT_CX, false>(sv, a); break;
        case AQC_ENT_CZ: rc = launch_apply_pass_t<AQC_ENT_CZ, false>(sv, a); break;
        default: rc = launch_apply_pass_t<AQC_ENT_CP, false>(sv, a);
      }
    }
    if (rc) return rc;
    sv->last_launches += 1;
  }
  return AQC_OK;
}

// runs one compiled program; NVEC == 1: src0 -> dst0; NVEC == 2: (w, z)
static int run_program(aqc_sv* sv, const Program& prog, bool grad, bool dag, const double2* src0,
                       long long basis, const double2* src1, double2* dst0, double2* dst1,
                       int pass_begin = 0, int pass_end = -1) {
  PassArgs a;
  memset(&a, 0, sizeof(a));
  a.vec_stride = sv->size;
  a.stages = prog.d_stages;
  a.trig = sv->d_trig;
  a.gacc = sv->d_gacc;
  a.nthetas = sv->circ.nthetas;
  if (pass_end < 0) pass_end = (int)prog.passes.size();
  for (int i = pass_begin; i < pass_end; ++i) {
    a.pd = prog.passes[i];
    a.src[0] = (i == pass_begin) ? src0 : dst0;
    a.src[1] = (i == pass_begin) ? src1 : dst1;
    a.dst[0] = dst0;
    a.dst[1] = dst1;
    a.basis_index = (i == pass_begin) ? basis : -1;
    int rc;
    if (grad)
      rc = launch_pass_e<2, false>(sv, a);
    else if (dag)
      rc = launch_pass_e<1, true>(sv, a);
    else
      rc = launch_pass_e<1, false>(sv, a);
    if (rc) return rc;
    sv->last_launches += 1;
  }
  return AQC_OK;
}

static int env_int(const char* name, int dflt);
// ---- dense-stage engine: host side ----------------------------------------------------------------
template <int NVEC>
static int launch_dense_pass(aqc_sv* sv, const DensePassArgs& args) {
  const size_t smem = (size_t)NVEC * sizeof(double2) << args.pd.tb;
  static bool configured[8] = {false};
  if (!configured[sv->device & 7]) {
    CU(cudaFuncSetAttribute(dense_pass_kernel<NVEC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)((size_t)NVEC * sizeof(double2) << (NVEC == 2 ? kMaxTileBits - 1 : kMaxTileBits))));
    configured[sv->device & 7] = true;
  }
  dim3 grid((unsigned)(1ull << args.pd.nouter), (unsigned)sv->batch);
  dense_pass_kernel<NVEC><<<grid, kDThreads, smem, sv->stream>>>(args);
  CU(cudaGetLastError());
  return AQC_OK;
}

// stage matrices of one program for the uploaded angles (mode 0 gradient, 1 V, 2 V^H)
static int dense_prepare(aqc_sv* sv, int mode) {
  const Program& p = mode == 0 ? sv->prog_grad : (mode == 1 ? sv->prog_fwd : sv->prog_dag);
  const int ns = (int)p.stages.size();
  if (ns == 0) return AQC_OK;
  const dim3 grid((unsigned)((ns * 4 + 127) / 128), (unsigned)sv->batch);
  const bool dag = mode == 2;
#define AQC_UMAT(E)                                                                                  \
  do {                                                                                               \
    if (dag)                                                                                         \
      dense_umat_kernel<E, true><<<grid, 128, 0, sv->stream>>>(p.d_stages, ns, sv->d_trig,          \
                                                               sv->circ.nthetas, sv->d_umat);       \
    else                                                                                             \
      dense_umat_kernel<E, false><<<grid, 128, 0, sv->stream>>>(p.d_stages, ns, sv->d_trig,         \
                                                                sv->circ.nthetas, sv->d_umat);      \
  } while (0)
  switch (sv->circ.ent) {
    case AQC_ENT_CX: AQC_UMAT(AQC_ENT_CX); break;
    case AQC_ENT_CZ: AQC_UMAT(AQC_ENT_CZ); break;
    default: AQC_UMAT(AQC_ENT_CP);
  }
#undef AQC_UMAT
  CU(cudaGetLastError());
  sv->last_launches += 1;
  if (mode == 0)
    CU(cudaMemsetAsync(sv->d_gm, 0, (size_t)sv->batch * ns * 64 * sizeof(double), sv->stream));
  return AQC_OK;
}

// raw per-rotation sums (the format of pass_kernel) from the accumulated stage matrices
static int dense_collect(aqc_sv* sv) {
  const Program& p = sv->prog_grad;
  const int ns = (int)p.stages.size();
  const size_t tot = (size_t)sv->batch * sv->circ.nthetas;
  CU(cudaMemsetAsync(sv->d_gacc, 0, tot * 2 * sizeof(double), sv->stream));
  if (ns == 0) return AQC_OK;
  const dim3 grid((unsigned)((ns * 4 + 127) / 128), (unsigned)sv->batch);
  switch (sv->circ.ent) {
    case AQC_ENT_CX:
      dense_grad_kernel<AQC_ENT_CX><<<grid, 128, 0, sv->stream>>>(p.d_stages, ns, sv->d_trig,
                                                                  sv->circ.nthetas, sv->d_gm, sv->d_gacc);
      break;
    case AQC_ENT_CZ:
      dense_grad_kernel<AQC_ENT_CZ><<<grid, 128, 0, sv->stream>>>(p.d_stages, ns, sv->d_trig,
                                                                  sv->circ.nthetas, sv->d_gm, sv->d_gacc);
      break;
    default:
      dense_grad_kernel<AQC_ENT_CP><<<grid, 128, 0, sv->stream>>>(p.d_stages, ns, sv->d_trig,
                                                                  sv->circ.nthetas, sv->d_gm, sv->d_gacc);
  }
  CU(cudaGetLastError());
  sv->last_launches += 1;
  return AQC_OK;
}

// passes [pass_begin, pass_end) of a program on the dense engine; mode 0: (w, z), else one vector
static int run_dense_program(aqc_sv* sv, int mode, const double2* src0, long long basis,
                             const double2* src1, double2* dst0, double2* dst1, int pass_begin,
                             int pass_end) {
  const Program& prog = mode == 0 ? sv->prog_grad : (mode == 1 ? sv->prog_fwd : sv->prog_dag);
  const DenseTables& dt = mode == 0 ? sv->dt_grad : (mode == 1 ? sv->dt_fwd : sv->dt_dag);
  DensePassArgs a;
  memset(&a, 0, sizeof(a));
  a.vec_stride = sv->size;
  a.lanes = dt.d_lanes;
  a.umat = sv->d_umat;
  a.gm = sv->d_gm;
  a.nstages_total = (int)prog.stages.size();
  if (pass_end < 0) pass_end = (int)prog.passes.size();
  for (int i = pass_begin; i < pass_end; ++i) {
    a.pd = prog.passes[i];
    a.src[0] = (i == pass_begin) ? src0 : dst0;
    a.src[1] = (i == pass_begin) ? src1 : dst1;
    a.dst[0] = dst0;
    a.dst[1] = dst1;
    a.basis_index = (i == pass_begin) ? basis : -1;
    const int rc = mode == 0 ? launch_dense_pass<2>(sv, a) : launch_dense_pass<1>(sv, a);
    if (rc) return rc;
    sv->last_launches += 1;
  }
  return AQC_OK;
}

// ------------------------------------------------------------------------------------------
// C-ABI
// ------------------------------------------------------------------------------------------
extern "C" int aqc_circuit_create(int num_qubits, int entangler, const int32_t* blocks,
                                  int num_blocks, int trotter, aqc_circuit** out) {
  if (!out) return fail(AQC_EINVAL, "out is null");
  *out = nullptr;
  if (num_qubits < 2 || num_qubits > 62) return fail(AQC_EINVAL, "num_qubits must be in [2, 62]");
  if (entangler < 0 || entangler > 2) return fail(AQC_EINVAL, "unknown entangler %d", entangler);
  if (num_blocks < 0 || (num_blocks > 0 && !blocks)) return fail(AQC_EINVAL, "bad blocks");
  if (trotter < 0 || trotter > 2) return fail(AQC_EINVAL, "bad trotter flag");
  if (trotter != AQC_GENERIC && entangler != AQC_ENT_CX)
    return fail(AQC_EINVAL, "Trotter ansatz requires the cx entangler");
  aqc_circuit* c = new aqc_circuit();
  c->n = num_qubits;
  c->ent = entangler;
  c->trotter = trotter;
  c->nb = num_blocks;
  c->tpb = entangler == AQC_ENT_CP ? 5 : 4;
  c->nthetas = 3 * num_qubits + c->tpb * num_blocks;
  c->ctrl.assign(blocks, blocks + num_blocks);
  c->targ.assign(blocks + num_blocks, blocks + 2 * num_blocks);
  for (int i = 0; i < num_blocks; ++i) {
    const int a = c->ctrl[i], b = c->targ[i];
    if (a < 0 || a >= num_qubits || b < 0 || b >= num_qubits || a == b) {
      delete c;
      return fail(AQC_EINVAL, "not a valid structure of unit-blocks (block %d)", i);
    }
  }
  if (trotter != AQC_GENERIC) {
    // parametric_circuit.py:391-423: layers of triplets on adjacent qubits, middle one flipped
    bool ok = num_blocks % (3 * (num_qubits - 1)) == 0;
    for (int t = 0; ok && t < num_blocks / 3; ++t) {
      const int i = 3 * t;
      ok = c->ctrl[i] == c->ctrl[i + 2] && c->targ[i] == c->targ[i + 2] &&
           c->ctrl[i] == c->targ[i + 1] && c->targ[i] == c->ctrl[i + 1] &&
           c->ctrl[i] == c->targ[i] + 1;
    }
    if (ok && trotter == AQC_TROTTER_2ND && num_blocks > 0)
      for (int i = 0; ok && i < num_qubits / 2; ++i)
        ok = c->ctrl[3 * i + 1] == 2 * i && c->targ[3 * i + 1] == 2 * i + 1;
    if (!ok) {
      delete c;
      return fail(AQC_EINVAL, "not a valid Trotterized block layout");
    }
    c->half = (trotter == AQC_TROTTER_2ND && num_blocks > 0) ? 3 * (num_qubits / 2) : 0;
  }
  *out = c;
  return AQC_OK;
}

extern "C" void aqc_circuit_destroy(aqc_circuit* c) { delete c; }
extern "C" int aqc_circuit_num_thetas(const aqc_circuit* c) { return c ? c->nthetas : AQC_EINVAL; }

// Rotations of the gradient program in execution order (pass -> stage -> unit -> rotation), with the
// accumulator each one feeds: input of prep_kernel.
static void build_schedule(const aqc_circuit& c, Program& p) {
  const int units_total = c.n + c.nb + c.half;
  p.sched_theta.clear();
  p.sched_occ.clear();
  p.sched_pass.clear();
  p.pass_start.clear();
  p.occ_theta.assign((size_t)units_total * 5, -1);
  for (size_t ip = 0; ip < p.passes.size(); ++ip) {
    const PassDesc& pd = p.passes[ip];
    p.pass_start.push_back((int)p.sched_theta.size());
    for (int s = 0; s < pd.nstages; ++s) {
      const StageDesc& sd = p.stages[pd.stage0 + s];
      for (int u = 0; u < sd.nunits; ++u) {
        const UnitDesc& ud = sd.u[u];
        const bool front = ud.kind == U_FRONT_LO || ud.kind == U_FRONT_HI;
        // execution order inside a unit: front Rz(t2) Ry(t1) Rz(t0); block [cphase] t0 t1 t2 t3
        const int order_front[3] = {2, 1, 0};
        const int order_block[5] = {4, 0, 1, 2, 3};
        const int cnt = front ? 3 : 5;
        for (int i = 0; i < cnt; ++i) {
          const int k = front ? order_front[i] : order_block[i];
          if (!front && k == 4 && c.tpb != 5) continue;
          p.sched_theta.push_back(ud.theta + k);
          p.sched_occ.push_back(ud.slot + k);
          p.sched_pass.push_back((int)ip);
          p.occ_theta[ud.slot + k] = ud.theta + k;
        }
      }
    }
  }
}

static int upload_ints(const std::vector<int>& v, int** d) {
  if (v.empty()) return AQC_OK;
  CU(cudaMalloc(d, v.size() * sizeof(int)));
  CU(cudaMemcpy(*d, v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice));
  return AQC_OK;
}

static int upload_program(Program& p) {
  if (p.stages.empty()) return AQC_OK;
  CU(cudaMalloc(&p.d_stages, p.stages.size() * sizeof(StageDesc)));
  CU(cudaMemcpy(p.d_stages, p.stages.data(), p.stages.size() * sizeof(StageDesc),
                cudaMemcpyHostToDevice));
  return AQC_OK;
}

extern "C" void aqc_sv_destroy(aqc_sv* sv) {
  if (!sv) return;
  cudaSetDevice(sv->device);
  for (auto p : sv->slots)
    if (p) cudaFree(p);
  if (sv->d_thetas) cudaFree(sv->d_thetas);
  if (sv->d_trig) cudaFree(sv->d_trig);
  if (sv->d_gacc) cudaFree(sv->d_gacc);
  if (sv->d_scratch) cudaFree(sv->d_scratch);
  if (sv->d_idx) cudaFree(sv->d_idx);
  for (void* q : {(void*)sv->d_par, (void*)sv->d_logf, (void*)sv->d_lbuf, (void*)sv->d_ebuf, (void*)sv->d_rescale,
                  (void*)sv->d_dscale, (void*)sv->d_gocc, (void*)sv->prog_grad.d_sched_theta,
                  (void*)sv->prog_grad.d_sched_occ, (void*)sv->prog_grad.d_sched_pass,
                  (void*)sv->prog_grad.d_pass_start, (void*)sv->prog_grad.d_occ_theta,
                  (void*)sv->prog_fwd.d_sched_theta, (void*)sv->prog_fwd.d_pass_start,
                  (void*)sv->prog_dag.d_sched_theta, (void*)sv->prog_dag.d_pass_start, (void*)sv->d_apar,
                  (void*)sv->d_uph, (void*)sv->d_kappa, (void*)sv->d_albuf, (void*)sv->d_aebuf,
                  (void*)sv->d_arescale})
    if (q) cudaFree(q);
  for (void* q : {(void*)sv->d_cd_units, (void*)sv->d_cd_fobj, (void*)sv->d_target, (void*)sv->d_gram,
                  (void*)sv->d_rinv, (void*)sv->d_info})
    if (q) cudaFree(q);
  for (void* q : {(void*)sv->d_umat, (void*)sv->d_gm, (void*)sv->dt_grad.d_lanes, (void*)sv->dt_fwd.d_lanes,
                  (void*)sv->dt_dag.d_lanes})
    if (q) cudaFree(q);
  if (sv->h_pinned) cudaFreeHost(sv->h_pinned);
  for (Program* p : {&sv->prog_grad, &sv->prog_fwd, &sv->prog_dag})
    if (p->d_stages) cudaFree(p->d_stages);
  if (sv->ev0) cudaEventDestroy(sv->ev0);
  if (sv->ev1) cudaEventDestroy(sv->ev1);
  if (sv->tm0) cudaEventDestroy(sv->tm0);
  if (sv->tm1) cudaEventDestroy(sv->tm1);
  if (sv->stream) cudaStreamDestroy(sv->stream);
  delete sv;
}

static int env_int(const char* name, int dflt) {
  const char* s = getenv(name);
  return (s && *s) ? atoi(s) : dflt;
}

static int sv_create_impl(const aqc_circuit* circ, int device, int log2_cols, int batch,
                          int num_slots, int g, int rank, aqc_sv** out) {
  if (!out) return fail(AQC_EINVAL, "out is null");
  *out = nullptr;
  if (!circ) return fail(AQC_EINVAL, "circuit is null");
  if (log2_cols < 0 || log2_cols > circ->n)
    return fail(AQC_EINVAL, "log2_cols must be in [0, num_qubits]");
  if (batch < 1 || batch > 65535) return fail(AQC_EINVAL, "batch must be in [1, 65535]");
  if (num_slots < 1 || num_slots > 64) return fail(AQC_EINVAL, "num_slots must be in [1, 64]");
  const int ndev = aqc_device_count();
  if (ndev <= 0) return fail(AQC_ENODEV, "no CUDA device visible: this library has no CPU path");
  if (device < 0 || device >= ndev) return fail(AQC_EINVAL, "device %d out of range", device);
  if (circ->n + log2_cols - g > 34) return fail(AQC_EINVAL, "state too large for one GPU");
  if (g < 0 || g > 4 || (g > 0 && (log2_cols != 0 || batch != 1)))
    return fail(AQC_EINVAL, "sharding needs 1 <= log2_world <= 4, a vector state and batch 1");
  if (rank < 0 || rank >= (1 << g)) return fail(AQC_EINVAL, "rank out of range");
  CU(cudaSetDevice(device));
  aqc_sv* sv = new aqc_sv();
  sv->circ = *circ;
  sv->device = device;
  sv->log2_cols = log2_cols;
  sv->g = g;
  sv->rank = rank;
  memset(sv->peer, 0, sizeof(sv->peer));
  sv->nbits = circ->n + log2_cols - g;
  sv->batch = batch;
  sv->nslots = num_slots;
  sv->size = 1ll << sv->nbits;
  sv->slots.assign(num_slots, nullptr);
  auto bail = [&](int rc) {
    std::string keep = g_err;
    aqc_sv_destroy(sv);
    g_err = keep;
    return rc;
  };
#define CUB(call)                                                                        \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      fail(e_ == cudaErrorMemoryAllocation ? AQC_ENOMEM : AQC_ECUDA, "%s failed: %s", #call, \
           cudaGetErrorString(e_));                                                      \
      return bail(e_ == cudaErrorMemoryAllocation ? AQC_ENOMEM : AQC_ECUDA);             \
    }                                                                                    \
  } while (0)
  CUB(cudaDeviceGetAttribute(&sv->num_sms, cudaDevAttrMultiProcessorCount, device));
  CUB(cudaStreamCreateWithFlags(&sv->stream, cudaStreamNonBlocking));
  CUB(cudaEventCreate(&sv->ev0));
  CUB(cudaEventCreate(&sv->ev1));
  CUB(cudaEventCreate(&sv->tm0));
  CUB(cudaEventCreate(&sv->tm1));
  const size_t bytes = (size_t)sv->size * batch * sizeof(double2);
  for (int s = 0; s < num_slots; ++s) CUB(cudaMalloc(&sv->slots[s], bytes));
  const size_t tot = (size_t)batch * circ->nthetas;
  CUB(cudaMalloc(&sv->d_thetas, tot * sizeof(double)));
  CUB(cudaMalloc(&sv->d_trig, tot * sizeof(double2)));
  CUB(cudaMalloc(&sv->d_gacc, tot * 2 * sizeof(double)));
#undef CUB
  // Tile shape.  States beyond the L2 (> 64 MiB) want 256-byte contiguous runs (4 low bits) and the
  // largest tile; L2-resident states (nbits <= 22) have too few tiles to fill 3 CTAs on each of the
  // SMs, so they use smaller gradient tiles and spend the low bits on gate qubits instead
  // (measured at n = 20: 0.42 -> 0.36 ms per evaluation).
  const bool l2_resident = sv->nbits <= 22;
  const int tb_grad = std::min(env_int("AQC_TILE_BITS_GRAD", l2_resident ? 10 : 11), kMaxTileBits - 1);
  // single-vector sweeps of large states: 2^12-amplitude tiles (64 KiB) with 128-byte runs need fewer
  // passes (n = 28: 16 -> 13, 47.2 -> 45.5 ms)
  const int tb_apply = std::min(env_int("AQC_TILE_BITS_APPLY", l2_resident ? 11 : 12), kMaxTileBits);
  const int low = env_int("AQC_TILE_LOW_BITS", l2_resident ? 2 : 4);
  const int low_apply = env_int("AQC_TILE_LOW_BITS_APPLY", env_int("AQC_TILE_LOW_BITS", l2_resident ? 1 : 3));
  // engine: dense-stage DMMA sweeps (default), "scaled" (scale-free rotations) or "legacy"
  {
    const char* eng = getenv("AQC_ENGINE");
    const std::string e = eng ? eng : "dense";
    sv->legacy_grad = env_int("AQC_GRAD_LEGACY", 0) != 0 || e == "legacy";
    sv->dense = !sv->legacy_grad && e != "scaled";
    if (sv->nbits < kDMinTileBits || tb_grad < kDMinTileBits || tb_apply < kDMinTileBits) {
      // fewer than 8 amplitude quadruples: nothing for a DMMA to do
      sv->dense = false;
      sv->legacy_grad = true;
    }
  }
  const int max_units = sv->dense ? kStageUnits : kMaxUnits;
  if (g == 0) {
    build_program(sv->circ, log2_cols, sv->nbits, tb_grad, low, false, sv->prog_grad, max_units, sv->dense);
    build_program(sv->circ, log2_cols, sv->nbits, tb_apply, low_apply, false, sv->prog_fwd, max_units, sv->dense);
    build_program(sv->circ, log2_cols, sv->nbits, tb_apply, low_apply, true, sv->prog_dag, max_units, sv->dense);
  } else {
    std::string err;
    if (build_program_sharded(sv->circ, g, tb_grad, low, false, sv->prog_grad, err, max_units, sv->dense) ||
        build_program_sharded(sv->circ, g, tb_apply, low_apply, false, sv->prog_fwd, err, max_units, sv->dense) ||
        build_program_sharded(sv->circ, g, tb_apply, low_apply, true, sv->prog_dag, err, max_units, sv->dense)) {
      fail(AQC_EINVAL, "%s", err.c_str());
      return bail(AQC_EINVAL);
    }
  }
  for (Program* p : {&sv->prog_grad, &sv->prog_fwd, &sv->prog_dag}) {
    int rc = upload_program(*p);
    if (rc) return bail(rc);
  }
  if (sv->dense) {
    size_t smax = 1;
    Program* progs[3] = {&sv->prog_grad, &sv->prog_fwd, &sv->prog_dag};
    DenseTables* tabs[3] = {&sv->dt_grad, &sv->dt_fwd, &sv->dt_dag};
    for (int i = 0; i < 3; ++i) {
      if (!build_dense_tables(*progs[i], *tabs[i], env_int("AQC_DENSE_PAIRS", 1))) {
        fail(AQC_EINVAL, "internal: dense engine needs tiles of >= %d bits", kDMinTileBits);
        return bail(AQC_EINVAL);
      }
      smax = std::max(smax, progs[i]->stages.size());
      if (tabs[i]->lanes.empty()) continue;
      cudaError_t e = cudaMalloc(&tabs[i]->d_lanes, tabs[i]->lanes.size() * sizeof(DLane));
      if (e == cudaSuccess)
        e = cudaMemcpy(tabs[i]->d_lanes, tabs[i]->lanes.data(), tabs[i]->lanes.size() * sizeof(DLane),
                       cudaMemcpyHostToDevice);
      if (e != cudaSuccess) {
        fail(AQC_ENOMEM, "dense table upload failed: %s", cudaGetErrorString(e));
        return bail(AQC_ENOMEM);
      }
    }
    const size_t B = batch;
    cudaError_t e = cudaMalloc(&sv->d_umat, B * smax * 64 * sizeof(double));
    if (e == cudaSuccess)
      e = cudaMalloc(&sv->d_gm, B * std::max<size_t>(1, sv->prog_grad.stages.size()) * 64 * sizeof(double));
    if (e != cudaSuccess) {
      fail(AQC_ENOMEM, "dense scratch allocation failed: %s", cudaGetErrorString(e));
      return bail(AQC_ENOMEM);
    }
  } else {
  {
    Program& p = sv->prog_grad;
    build_schedule(sv->circ, p);
    sv->nocc = (int)p.occ_theta.size();
    sv->nsched = (int)p.sched_theta.size();
    int rc = upload_ints(p.sched_theta, &p.d_sched_theta);
    if (!rc) rc = upload_ints(p.sched_occ, &p.d_sched_occ);
    if (!rc) rc = upload_ints(p.sched_pass, &p.d_sched_pass);
    if (!rc) rc = upload_ints(p.pass_start, &p.d_pass_start);
    if (!rc) rc = upload_ints(p.occ_theta, &p.d_occ_theta);
    if (rc) return bail(rc);
    const size_t B = batch, np = p.passes.size();
    cudaError_t e = cudaMalloc(&sv->d_par, B * circ->nthetas * sizeof(double2));
    if (e == cudaSuccess) e = cudaMalloc(&sv->d_logf, B * circ->nthetas * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&sv->d_lbuf, B * std::max(1, sv->nsched) * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&sv->d_ebuf, B * np * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&sv->d_rescale, B * np * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&sv->d_dscale, B * sv->nocc * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&sv->d_gocc, B * sv->nocc * 2 * sizeof(double));
    if (e != cudaSuccess) {
      fail(AQC_ENOMEM, "gradient scratch allocation failed: %s", cudaGetErrorString(e));
      return bail(AQC_ENOMEM);
    }
  }
  {
    size_t jmax = 1, pmax = 1;
    for (Program* p : {&sv->prog_fwd, &sv->prog_dag}) {
      build_schedule(sv->circ, *p);
      int rc = upload_ints(p->sched_theta, &p->d_sched_theta);
      if (!rc) rc = upload_ints(p->pass_start, &p->d_pass_start);
      if (rc) return bail(rc);
      jmax = std::max(jmax, p->sched_theta.size());
      pmax = std::max(pmax, p->passes.size());
    }
    const size_t B = batch;
    cudaError_t e = cudaMalloc(&sv->d_apar, B * circ->nthetas * sizeof(double2));
    if (e == cudaSuccess) e = cudaMalloc(&sv->d_uph, B * circ->nthetas * sizeof(double2));
    if (e == cudaSuccess) e = cudaMalloc(&sv->d_kappa, B * sizeof(double2));
    if (e == cudaSuccess) e = cudaMalloc(&sv->d_albuf, B * jmax * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&sv->d_aebuf, B * pmax * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&sv->d_arescale, B * pmax * sizeof(double));
    if (e != cudaSuccess) {
      fail(AQC_ENOMEM, "apply scratch allocation failed: %s", cudaGetErrorString(e));
      return bail(AQC_ENOMEM);
    }
  }
  }
  *out = sv;
  return AQC_OK;
}

extern "C" int aqc_sv_create(const aqc_circuit* circ, int device, int log2_cols, int batch,
                             int num_slots, aqc_sv** out) {
  return sv_create_impl(circ, device, log2_cols, batch, num_slots, 0, 0, out);
}

extern "C" int aqc_sv_create_sharded(const aqc_circuit* circ, int device, int log2_world, int rank,
                                     int num_slots, aqc_sv** out) {
  if (log2_world < 1) return fail(AQC_EINVAL, "log2_world must be >= 1");
  return sv_create_impl(circ, device, 0, 1, num_slots, log2_world, rank, out);
}

extern "C" int64_t aqc_sv_state_size(const aqc_sv* sv) { return sv ? sv->size : 0; }
extern "C" float aqc_sv_last_kernel_ms(const aqc_sv* sv) { return sv ? sv->last_ms : 0.f; }
extern "C" int aqc_sv_last_num_launches(const aqc_sv* sv) { return sv ? sv->last_launches : 0; }
extern "C" void* aqc_sv_slot_ptr(aqc_sv* sv, int slot) {
  return (sv && slot >= 0 && slot < sv->nslots) ? (void*)sv->slots[slot] : nullptr;
}
extern "C" void* aqc_sv_stream(aqc_sv* sv) { return sv ? (void*)sv->stream : nullptr; }
extern "C" int aqc_sv_num_passes(const aqc_sv* sv, int mode) {
  if (!sv) return AQC_EINVAL;
  const Program& p = mode == 0 ? sv->prog_grad : (mode == 1 ? sv->prog_fwd : sv->prog_dag);
  return (int)p.passes.size();
}

extern "C" int aqc_sv_num_stages(const aqc_sv* sv, int mode) {
  if (!sv) return AQC_EINVAL;
  const Program& p = mode == 0 ? sv->prog_grad : (mode == 1 ? sv->prog_fwd : sv->prog_dag);
  return (int)p.stages.size();
}

extern "C" int aqc_sv_upload(aqc_sv* sv, int slot, int batch_index, const double* host,
                             int64_t count) {
  int rc = check_slot(sv, slot);
  if (rc) return rc;
  if (!host || count < 0 || count > sv->size) return fail(AQC_EINVAL, "bad host buffer/count");
  if (batch_index < -1 || batch_index >= sv->batch) return fail(AQC_EINVAL, "bad batch index");
  CU(cudaSetDevice(sv->device));
  const int b0 = batch_index < 0 ? 0 : batch_index, b1 = batch_index < 0 ? sv->batch : b0 + 1;
  for (int b = b0; b < b1; ++b)
    CU(cudaMemcpyAsync(sv->slots[slot] + (size_t)b * sv->size, host, (size_t)count * 16,
                       cudaMemcpyHostToDevice, sv->stream));
  CU(cudaStreamSynchronize(sv->stream));
  return AQC_OK;
}

extern "C" int aqc_sv_download(aqc_sv* sv, int slot, int batch_index, double* host,
                               int64_t count) {
  int rc = check_slot(sv, slot);
  if (rc) return rc;
  if (!host || count < 0 || count > sv->size) return fail(AQC_EINVAL, "bad host buffer/count");
  if (batch_index < 0 || batch_index >= sv->batch) return fail(AQC_EINVAL, "bad batch index");
  CU(cudaSetDevice(sv->device));
  CU(cudaMemcpyAsync(host, sv->slots[slot] + (size_t)batch_index * sv->size, (size_t)count * 16,
                     cudaMemcpyDeviceToHost, sv->stream));
  CU(cudaStreamSynchronize(sv->stream));
  return AQC_OK;
}

static dim3 grid1d(long long size, int batch, int thr) {
  return dim3((unsigned)((size + thr - 1) / thr), (unsigned)batch);
}

extern "C" int aqc_sv_set_basis(aqc_sv* sv, int slot, int64_t index) {
  int rc = check_slot(sv, slot);
  if (rc) return rc;
  if (index < -1 || index >= sv->size) return fail(AQC_EINVAL, "basis index out of range");
  CU(cudaSetDevice(sv->device));
  set_basis_kernel<<<grid1d(sv->size, sv->batch, 256), 256, 0, sv->stream>>>(
      sv->slots[slot], sv->size, sv->size, index);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(sv->stream));
  return AQC_OK;
}

extern "C" int aqc_sv_set_sparse(aqc_sv* sv, int slot, const int64_t* indices, const double* amps,
                                 int count) {
  int rc = check_slot(sv, slot);
  if (rc) return rc;
  if (!indices || !amps) return fail(AQC_EINVAL, "null pointer argument");
  if (count < 1 || count > 8) return fail(AQC_EINVAL, "a sparse state holds 1 to 8 basis amplitudes");
  SparseInit init;
  init.count = count;
  for (int k = 0; k < count; ++k) {
    if (indices[k] < 0 || indices[k] >= sv->size) return fail(AQC_EINVAL, "basis index out of range");
    init.index[k] = indices[k];
    init.amp[k] = make_double2(amps[2 * k], amps[2 * k + 1]);
  }
  CU(cudaSetDevice(sv->device));
  CU(cudaMemsetAsync(sv->slots[slot], 0, (size_t)sv->batch * sv->size * sizeof(double2), sv->stream));
  set_sparse_kernel<<<sv->batch, 32, 0, sv->stream>>>(sv->slots[slot], sv->size, init);
  CU(cudaGetLastError());
  return AQC_OK;
}

extern "C" int aqc_sv_set_identity(aqc_sv* sv, int slot) {
  int rc = check_slot(sv, slot);
  if (rc) return rc;
  if (sv->log2_cols != sv->circ.n) return fail(AQC_EINVAL, "identity needs log2_cols == n");
  CU(cudaSetDevice(sv->device));
  set_identity_kernel<<<grid1d(sv->size, sv->batch, 256), 256, 0, sv->stream>>>(
      sv->slots[slot], sv->size, sv->size, sv->log2_cols);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(sv->stream));
  return AQC_OK;
}

extern "C" int aqc_sv_fill_random(aqc_sv* sv, int slot, uint64_t seed) {
  int rc = check_slot(sv, slot);
  if (rc) return rc;
  CU(cudaSetDevice(sv->device));
  rc = ensure_scratch(sv, (size_t)sv->batch);
  if (rc) return rc;
  CU(cudaMemsetAsync(sv->d_scratch, 0, sv->batch * sizeof(double), sv->stream));
  const unsigned gx = (unsigned)std::min<long long>((sv->size + 255) / 256, 148 * 16);
  fill_random_kernel<<<dim3(gx, sv->batch), 256, 0, sv->stream>>>(sv->slots[slot], sv->size,
                                                                   sv->size, seed, sv->d_scratch);
  scale_kernel<<<dim3(gx, sv->batch), 256, 0, sv->stream>>>(sv->slots[slot], sv->size, sv->size,
                                                             sv->d_scratch);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(sv->stream));
  return AQC_OK;
}

static int gather_async(aqc_sv* sv, int slot, const int64_t* idx, int count) {
  int rc = ensure_idx(sv, (size_t)count);
  if (rc) return rc;
  rc = ensure_scratch(sv, (size_t)2 * count * sv->batch);
  if (rc) return rc;
  for (int i = 0; i < count; ++i)
    if (idx[i] < 0 || idx[i] >= sv->size) return fail(AQC_EINVAL, "gather index out of range");
  // the index list is the same on every objective call: upload it only when it changes (a copy from
  // pageable host memory stalls the submitting thread)
  if (sv->idx_cached.size() != (size_t)count || memcmp(sv->idx_cached.data(), idx, (size_t)count * sizeof(int64_t))) {
    sv->idx_cached.assign(idx, idx + count);
    CU(cudaStreamSynchronize(sv->stream));  // a previous gather may still read d_idx
    CU(cudaMemcpy(sv->d_idx, sv->idx_cached.data(), (size_t)count * sizeof(long long), cudaMemcpyHostToDevice));
  }
  gather_kernel<<<dim3((count + 127) / 128, sv->batch), 128, 0, sv->stream>>>(
      sv->slots[slot], sv->size, sv->d_idx, count, (double2*)sv->d_scratch);
  CU(cudaGetLastError());
  return AQC_OK;
}

extern "C" int aqc_sv_gather(aqc_sv* sv, int slot, const int64_t* idx, int count, double* out) {
  int rc = check_slot(sv, slot);
  if (rc) return rc;
  if (!idx || !out || count <= 0) return fail(AQC_EINVAL, "bad gather arguments");
  CU(cudaSetDevice(sv->device));
  rc = gather_async(sv, slot, idx, count);
  if (rc) return rc;
  CU(cudaMemcpyAsync(out, sv->d_scratch, (size_t)2 * count * sv->batch * sizeof(double),
                     cudaMemcpyDeviceToHost, sv->stream));
  CU(cudaStreamSynchronize(sv->stream));
  return AQC_OK;
}

extern "C" int aqc_sv_vdot(aqc_sv* sv, int slot_a, int slot_b, double* out) {
  int rc = check_slot(sv, slot_a);
  if (rc) return rc;
  rc = check_slot(sv, slot_b);
  if (rc) return rc;
  if (!out) return fail(AQC_EINVAL, "out is null");
  CU(cudaSetDevice(sv->device));
  rc = ensure_scratch(sv, (size_t)2 * sv->batch);
  if (rc) return rc;
  CU(cudaMemsetAsync(sv->d_scratch, 0, 2 * sv->batch * sizeof(double), sv->stream));
  const unsigned gx = (unsigned)std::min<long long>((sv->size + 255) / 256, 148 * 8);
  vdot_kernel<<<dim3(gx, sv->batch), 256, 0, sv->stream>>>(sv->slots[slot_a], sv->slots[slot_b],
                                                            sv->size, sv->size, sv->d_scratch);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(out, sv->d_scratch, 2 * sv->batch * sizeof(double), cudaMemcpyDeviceToHost,
                     sv->stream));
  CU(cudaStreamSynchronize(sv->stream));
  return AQC_OK;
}

static int apply_async(aqc_sv* sv, const double* thetas, int dagger, int src_slot, int dst_slot) {
  int rc = ensure_pinned(sv, (size_t)sv->batch * sv->circ.nthetas * 2 + 64);
  if (rc) return rc;
  rc = upload_thetas(sv, thetas);
  if (rc) return rc;
  CU(cudaEventRecord(sv->ev0, sv->stream));
  if (sv->dense) {
    rc = dense_prepare(sv, dagger ? 2 : 1);
    if (!rc)
      rc = run_dense_program(sv, dagger ? 2 : 1, sv->slots[src_slot], -1, nullptr, sv->slots[dst_slot],
                             nullptr, 0, -1);
  } else if (sv->legacy_grad) {
    const Program& prog = dagger ? sv->prog_dag : sv->prog_fwd;
    rc = run_program(sv, prog, false, dagger != 0, sv->slots[src_slot], -1, nullptr,
                     sv->slots[dst_slot], nullptr);
  } else {
    rc = apply_prepare(sv, dagger != 0);
    if (!rc) rc = run_apply_program(sv, dagger != 0, sv->slots[src_slot], sv->slots[dst_slot], 0, -1);
  }
  if (rc) return rc;
  CU(cudaEventRecord(sv->ev1, sv->stream));
  return AQC_OK;
}

extern "C" int aqc_sv_apply(aqc_sv* sv, const double* thetas, int dagger, int src_slot,
                            int dst_slot) {
  int rc = check_slot(sv, src_slot);
  if (rc) return rc;
  rc = check_slot(sv, dst_slot);
  if (rc) return rc;
  if (!thetas) return fail(AQC_EINVAL, "thetas is null");
  CU(cudaSetDevice(sv->device));
  sv->last_launches = 0;
  rc = apply_async(sv, thetas, dagger, src_slot, dst_slot);
  if (rc) return rc;
  CU(cudaStreamSynchronize(sv->stream));
  CU(cudaEventElapsedTime(&sv->last_ms, sv->ev0, sv->ev1));
  return AQC_OK;
}

// ------------------------------------------------------------------------------------------
// coordinate descent (unitary AQC)
// ------------------------------------------------------------------------------------------
extern "C" int aqc_sv_coord_descent(aqc_sv* sv, double* thetas, int target_slot, int w_slot, int z_slot,
                                    int num_sweeps, double* fobj_out) {
  int rc = check_slot(sv, target_slot);
  if (!rc) rc = check_slot(sv, w_slot);
  if (!rc) rc = check_slot(sv, z_slot);
  if (rc) return rc;
  if (!thetas || !fobj_out || num_sweeps < 1) return fail(AQC_EINVAL, "bad arguments");
  if (w_slot == z_slot || w_slot == target_slot || z_slot == target_slot)
    return fail(AQC_EINVAL, "target, w and z must be three different slots");
  if (sv->log2_cols != sv->circ.n) return fail(AQC_EINVAL, "coordinate descent needs a square (2^n x 2^n) target");
  if (sv->g != 0) return fail(AQC_EINVAL, "coordinate descent is not sharded");
  if (sv->circ.ent == AQC_ENT_CP) return fail(AQC_EINVAL, "CPhase entangler is not supported yet");
  if (sv->circ.trotter != AQC_GENERIC) return fail(AQC_EINVAL, "coordinate descent needs a ParametricCircuit");
  CU(cudaSetDevice(sv->device));
  const int n = sv->circ.n, T = sv->circ.nthetas;
  if (!sv->d_cd_units) {
    std::vector<CdUnit> units;
    for (int q = 0; q < n; ++q) units.push_back({0, q, (q + 1) % n, 3 * q});
    for (int i = 0; i < sv->circ.nb; ++i)
      units.push_back({1, sv->circ.ctrl[i], sv->circ.targ[i], 3 * n + sv->circ.tpb * i});
    sv->cd_nunits = (int)units.size();
    CU(cudaMalloc(&sv->d_cd_units, units.size() * sizeof(CdUnit)));
    CU(cudaMemcpy(sv->d_cd_units, units.data(), units.size() * sizeof(CdUnit), cudaMemcpyHostToDevice));
    CU(cudaMalloc(&sv->d_cd_fobj, sv->batch * sizeof(double)));
  }
  const size_t tot = (size_t)sv->batch * T;
  rc = ensure_pinned(sv, tot * 2 + 64 + sv->batch);
  if (rc) return rc;
  sv->last_launches = 0;
  float total_ms = 0.f;
  CdArgs a;
  a.w = sv->slots[w_slot];
  a.z = sv->slots[z_slot];
  a.vec_stride = sv->size;
  a.thetas = sv->d_thetas;
  a.units = sv->d_cd_units;
  a.nunits = sv->cd_nunits;
  a.n = n;
  a.T = T;
  a.fobj = sv->d_cd_fobj;
  for (int sweep = 0; sweep < num_sweeps; ++sweep) {
    // z = V(thetas)^H target (thetas uploaded to d_thetas on the way), w = I
    rc = apply_async(sv, thetas, 1, target_slot, z_slot);
    if (rc) return rc;
    set_identity_kernel<<<grid1d(sv->size, sv->batch, 256), 256, 0, sv->stream>>>(
        sv->slots[w_slot], sv->size, sv->size, sv->log2_cols);
    if (sv->circ.ent == AQC_ENT_CX)
      cd_sweep_kernel<AQC_ENT_CX><<<sv->batch, kCdThreads, 0, sv->stream>>>(a);
    else
      cd_sweep_kernel<AQC_ENT_CZ><<<sv->batch, kCdThreads, 0, sv->stream>>>(a);
    CU(cudaGetLastError());
    sv->last_launches += 2;
    CU(cudaEventRecord(sv->ev1, sv->stream));
    // the pinned buffer holds the angles the apply above has consumed already (stream order)
    CU(cudaMemcpyAsync(sv->h_pinned, sv->d_thetas, tot * sizeof(double), cudaMemcpyDeviceToHost, sv->stream));
    CU(cudaMemcpyAsync(sv->h_pinned + tot, sv->d_cd_fobj, sv->batch * sizeof(double), cudaMemcpyDeviceToHost,
                       sv->stream));
    CU(cudaStreamSynchronize(sv->stream));
    memcpy(thetas, sv->h_pinned, tot * sizeof(double));
    memcpy(fobj_out + (size_t)sweep * sv->batch, sv->h_pinned + tot, sv->batch * sizeof(double));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, sv->ev0, sv->ev1));
    total_ms += ms;
  }
  sv->last_ms = total_ms;
  return AQC_OK;
}

// ------------------------------------------------------------------------------------------
// sketching-vector generators (dense target, GEMMs, thin QR)
// ------------------------------------------------------------------------------------------
static int sketch_check(aqc_sv* sv) {
  if (!sv) return fail(AQC_EINVAL, "null workspace");
  if (sv->g != 0 || sv->batch != 1) return fail(AQC_EINVAL, "sketching needs an unsharded workspace with batch 1");
  return AQC_OK;
}

static int launch_zgemm(aqc_sv* sv, bool trans, const double2* A, long long lda, const double2* B, long long ldb,
                        double2* C, long long ldc, int M, int N, int K) {
  GemmArgs g;
  g.A = A, g.B = B, g.C = C;
  g.M = M, g.N = N, g.K = K;
  g.lda = lda, g.ldb = ldb, g.ldc = ldc;
  const int tiles = ((M + kGemmTileM - 1) / kGemmTileM) * ((N + kGemmTileN - 1) / kGemmTileN);
  int ksplit = 1;
  if (tiles < sv->num_sms) ksplit = std::max(1, std::min((K + 255) / 256, (2 * sv->num_sms) / std::max(1, tiles)));
  g.ksplit = ksplit;
  if (ksplit > 1) CU(cudaMemsetAsync(C, 0, (size_t)M * ldc * sizeof(double2), sv->stream));
  const dim3 grid((M + kGemmTileM - 1) / kGemmTileM, (N + kGemmTileN - 1) / kGemmTileN, ksplit);
  if (trans)
    zgemm_kernel<1><<<grid, kGemmThreads, 0, sv->stream>>>(g);
  else
    zgemm_kernel<0><<<grid, kGemmThreads, 0, sv->stream>>>(g);
  CU(cudaGetLastError());
  sv->last_launches += 1;
  return AQC_OK;
}

extern "C" int aqc_sv_set_dense_target(aqc_sv* sv, const double* target) {
  int rc = sketch_check(sv);
  if (rc) return rc;
  if (!target) return fail(AQC_EINVAL, "target is null");
  CU(cudaSetDevice(sv->device));
  const size_t d = (size_t)1 << sv->circ.n;
  if (!sv->d_target) {
    cudaError_t e = cudaMalloc(&sv->d_target, d * d * sizeof(double2));
    if (e != cudaSuccess) return fail(AQC_ENOMEM, "dense target allocation failed: %s", cudaGetErrorString(e));
  }
  CU(cudaMemcpyAsync(sv->d_target, target, d * d * sizeof(double2), cudaMemcpyHostToDevice, sv->stream));
  CU(cudaStreamSynchronize(sv->stream));
  return AQC_OK;
}

extern "C" int aqc_sv_target_matmul(aqc_sv* sv, int conj_transpose, int src_slot, int dst_slot) {
  int rc = sketch_check(sv);
  if (!rc) rc = check_slot(sv, src_slot);
  if (!rc) rc = check_slot(sv, dst_slot);
  if (rc) return rc;
  if (src_slot == dst_slot) return fail(AQC_EINVAL, "matmul is out of place");
  if (!sv->d_target) return fail(AQC_EINVAL, "no dense target (aqc_sv_set_dense_target)");
  CU(cudaSetDevice(sv->device));
  sv->last_launches = 0;
  const int d = 1 << sv->circ.n, m = 1 << sv->log2_cols;
  CU(cudaEventRecord(sv->ev0, sv->stream));
  rc = launch_zgemm(sv, conj_transpose != 0, sv->d_target, d, sv->slots[src_slot], m, sv->slots[dst_slot], m, d, m, d);
  if (rc) return rc;
  CU(cudaEventRecord(sv->ev1, sv->stream));
  CU(cudaStreamSynchronize(sv->stream));
  CU(cudaEventElapsedTime(&sv->last_ms, sv->ev0, sv->ev1));
  return AQC_OK;
}

extern "C" int aqc_sv_orthonormalize(aqc_sv* sv, int slot, int tmp_slot) {
  int rc = sketch_check(sv);
  if (!rc) rc = check_slot(sv, slot);
  if (!rc) rc = check_slot(sv, tmp_slot);
  if (rc) return rc;
  if (slot == tmp_slot) return fail(AQC_EINVAL, "tmp_slot must differ from slot");
  CU(cudaSetDevice(sv->device));
  sv->last_launches = 0;
  const int d = 1 << sv->circ.n, m = 1 << sv->log2_cols;
  if (!sv->d_gram) {
    CU(cudaMalloc(&sv->d_gram, (size_t)m * m * sizeof(double2)));
    CU(cudaMalloc(&sv->d_rinv, (size_t)m * m * sizeof(double2)));
    CU(cudaMalloc(&sv->d_info, sizeof(int)));
  }
  rc = ensure_scratch(sv, 4);
  if (rc) return rc;
  rc = ensure_pinned(sv, 64);
  if (rc) return rc;
  CU(cudaEventRecord(sv->ev0, sv->stream));
  CU(cudaMemsetAsync(sv->d_info, 0, sizeof(int), sv->stream));
  // shift of the first pass: 11 (d m + m (m + 1)) u ||A||_F^2 (shifted Cholesky-QR3)
  CU(cudaMemsetAsync(sv->d_scratch, 0, sizeof(double), sv->stream));
  norm2_kernel<<<std::min(1024, (int)(((long long)d * m + 255) / 256)), 256, 0, sv->stream>>>(
      sv->slots[slot], (long long)d * m, sv->d_scratch);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(sv->h_pinned, sv->d_scratch, sizeof(double), cudaMemcpyDeviceToHost, sv->stream));
  CU(cudaStreamSynchronize(sv->stream));
  const double norm2 = sv->h_pinned[0];
  if (!(norm2 > 0.0)) return fail(AQC_EINVAL, "cannot orthonormalise a zero matrix");
  const double shift0 = 11.0 * ((double)d * m + (double)m * (m + 1)) * 1.1102230246251565e-16 * norm2;
  double2 *cur = sv->slots[slot], *nxt = sv->slots[tmp_slot];
  for (int pass = 0; pass < 3; ++pass) {
    rc = launch_zgemm(sv, true, cur, m, cur, m, sv->d_gram, m, m, m, d);  // G = A^H A
    if (rc) return rc;
    chol_inv_kernel<<<1, 256, 0, sv->stream>>>(sv->d_gram, sv->d_rinv, m, pass == 0 ? shift0 : 0.0, sv->d_info);
    CU(cudaGetLastError());
    rc = launch_zgemm(sv, false, cur, m, sv->d_rinv, m, nxt, m, d, m, m);  // A <- A R^-1
    if (rc) return rc;
    sv->last_launches += 1;
    std::swap(cur, nxt);
  }
  // three passes: the result sits in tmp_slot
  CU(cudaMemcpyAsync(sv->slots[slot], cur, (size_t)d * m * sizeof(double2), cudaMemcpyDeviceToDevice, sv->stream));
  CU(cudaEventRecord(sv->ev1, sv->stream));
  int* h_info = reinterpret_cast<int*>(sv->h_pinned);
  CU(cudaMemcpyAsync(h_info, sv->d_info, sizeof(int), cudaMemcpyDeviceToHost, sv->stream));
  CU(cudaStreamSynchronize(sv->stream));
  CU(cudaEventElapsedTime(&sv->last_ms, sv->ev0, sv->ev1));
  if (*h_info != 0) return fail(AQC_EINVAL, "sketching matrix is numerically rank deficient (column %d)", *h_info - 1);
  return AQC_OK;
}

extern "C" int aqc_sv_sub(aqc_sv* sv, int dst_slot, int src_slot) {
  int rc = check_slot(sv, dst_slot);
  if (!rc) rc = check_slot(sv, src_slot);
  if (rc) return rc;
  CU(cudaSetDevice(sv->device));
  const long long count = sv->size * sv->batch;
  sub_kernel<<<(unsigned)std::min<long long>((count + 255) / 256, 148 * 16), 256, 0, sv->stream>>>(
      sv->slots[dst_slot], sv->slots[src_slot], count);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(sv->stream));
  return AQC_OK;
}

extern "C" int aqc_sv_gather_target_columns(aqc_sv* sv, const int64_t* idx, int count, int x_slot, int y_slot) {
  int rc = sketch_check(sv);
  if (!rc) rc = check_slot(sv, x_slot);
  if (!rc) rc = check_slot(sv, y_slot);
  if (rc) return rc;
  const int d = 1 << sv->circ.n, m = 1 << sv->log2_cols;
  if (!idx || count != m) return fail(AQC_EINVAL, "expects exactly %d column indices", m);
  if (x_slot == y_slot) return fail(AQC_EINVAL, "x and y must be different slots");
  if (!sv->d_target) return fail(AQC_EINVAL, "no dense target (aqc_sv_set_dense_target)");
  for (int i = 0; i < count; ++i)
    if (idx[i] < 0 || idx[i] >= d) return fail(AQC_EINVAL, "column index out of range");
  CU(cudaSetDevice(sv->device));
  rc = ensure_idx(sv, (size_t)count);
  if (rc) return rc;
  sv->idx_cached.clear();  // d_idx is about to hold something else
  CU(cudaMemcpyAsync(sv->d_idx, idx, (size_t)count * sizeof(long long), cudaMemcpyHostToDevice, sv->stream));
  const long long tot = (long long)d * m;
  gather_cols_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, sv->stream>>>(sv->d_target, sv->d_idx, d, m,
                                                                           sv->slots[x_slot], sv->slots[y_slot]);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(sv->stream));
  return AQC_OK;
}

extern "C" int aqc_sv_objective(aqc_sv* sv, const double* thetas, int target_slot, int z0_slot,
                                const int64_t* idx, int count, double* hs_out) {
  int rc = check_slot(sv, target_slot);
  if (rc) return rc;
  rc = check_slot(sv, z0_slot);
  if (rc) return rc;
  if (!thetas || !idx || !hs_out || count <= 0) return fail(AQC_EINVAL, "bad arguments");
  CU(cudaSetDevice(sv->device));
  sv->last_launches = 0;
  const size_t nout = (size_t)2 * count * sv->batch;
  rc = ensure_pinned(sv, nout + (size_t)sv->batch * sv->circ.nthetas * 2 + 64);
  if (rc) return rc;
  rc = apply_async(sv, thetas, 1, target_slot, z0_slot);
  if (rc) return rc;
  rc = gather_async(sv, z0_slot, idx, count);
  if (rc) return rc;
  sv->last_launches += 1;
  CU(cudaMemcpyAsync(sv->h_pinned, sv->d_scratch, nout * sizeof(double), cudaMemcpyDeviceToHost,
                     sv->stream));
  CU(cudaStreamSynchronize(sv->stream));
  memcpy(hs_out, sv->h_pinned, nout * sizeof(double));
  CU(cudaEventElapsedTime(&sv->last_ms, sv->ev0, sv->ev1));
  return AQC_OK;
}

extern "C" int aqc_sv_grad_begin(aqc_sv* sv, const double* thetas, int x_slot, int64_t x_basis,
                                 int z0_slot, int w_slot, int z_slot) {
  int rc = check_slot(sv, z0_slot);
  if (rc) return rc;
  rc = check_slot(sv, w_slot);
  if (rc) return rc;
  rc = check_slot(sv, z_slot);
  if (rc) return rc;
  if (x_slot >= 0) {
    rc = check_slot(sv, x_slot);
    if (rc) return rc;
  } else if (x_basis < 0 || x_basis >= sv->size) {
    return fail(AQC_EINVAL, "basis index out of range");
  }
  if (w_slot == z_slot || w_slot == z0_slot || (x_slot >= 0 && x_slot == z_slot))
    return fail(AQC_EINVAL, "slot aliasing: w must differ from z/z0 and x from z");
  if (!thetas) return fail(AQC_EINVAL, "null pointer argument");
  CU(cudaSetDevice(sv->device));
  sv->last_launches = 0;
  const size_t tot = (size_t)sv->batch * sv->circ.nthetas;
  rc = ensure_pinned(sv, tot * 2 + 64);
  if (rc) return rc;
  rc = upload_thetas(sv, thetas);
  if (rc) return rc;
  CU(cudaMemsetAsync(sv->d_gacc, 0, tot * 2 * sizeof(double), sv->stream));
  CU(cudaEventRecord(sv->ev0, sv->stream));
  if (sv->dense) {
    rc = dense_prepare(sv, 0);
    if (!rc)
      rc = run_dense_program(sv, 0, x_slot >= 0 ? sv->slots[x_slot] : nullptr, x_slot >= 0 ? -1 : x_basis,
                             sv->slots[z0_slot], sv->slots[w_slot], sv->slots[z_slot], 0, -1);
    if (!rc) rc = dense_collect(sv);
  } else if (sv->legacy_grad) {
    rc = run_program(sv, sv->prog_grad, true, false, x_slot >= 0 ? sv->slots[x_slot] : nullptr,
                     x_slot >= 0 ? -1 : x_basis, sv->slots[z0_slot], sv->slots[w_slot],
                     sv->slots[z_slot]);
  } else {
    rc = grad_prepare(sv);
    if (!rc)
      rc = run_grad_program(sv, x_slot >= 0 ? sv->slots[x_slot] : nullptr, x_slot >= 0 ? -1 : x_basis,
                            sv->slots[z0_slot], sv->slots[w_slot], sv->slots[z_slot], 0, -1);
    if (!rc) rc = grad_collect(sv);
  }
  if (rc) return rc;
  CU(cudaEventRecord(sv->ev1, sv->stream));
  CU(cudaMemcpyAsync(sv->h_pinned, sv->d_gacc, tot * 2 * sizeof(double), cudaMemcpyDeviceToHost,
                     sv->stream));
  sv->grad_pending = true;
  return AQC_OK;
}

extern "C" int aqc_sv_grad_end(aqc_sv* sv, double* grad_out) {
  if (!sv || !grad_out) return fail(AQC_EINVAL, "null pointer argument");
  if (!sv->grad_pending) return fail(AQC_EINVAL, "no gradient sweep in flight (aqc_sv_grad_begin)");
  CU(cudaSetDevice(sv->device));
  sv->grad_pending = false;
  const size_t tot = (size_t)sv->batch * sv->circ.nthetas;
  CU(cudaStreamSynchronize(sv->stream));
  CU(cudaEventElapsedTime(&sv->last_ms, sv->ev0, sv->ev1));
  // raw sums -> 0.5j <P w|z>: Ry 0.5, Rz/Rx 0.5j, CPhase -i
  const int n3 = 3 * sv->circ.n, tpb = sv->circ.tpb, T = sv->circ.nthetas;
  const bool cx = sv->circ.ent == AQC_ENT_CX;
  (void)cx;
  for (int b = 0; b < sv->batch; ++b) {
    const double* raw = sv->h_pinned + (size_t)b * T * 2;
    double* g = grad_out + (size_t)b * T * 2;
    for (int k = 0; k < T; ++k) {
      const double re = raw[2 * k], im = raw[2 * k + 1];
      int kind;  // 0: Ry (0.5), 1: Rz/Rx (0.5j), 2: cphase (-i)
      if (k < n3)
        kind = (k % 3 == 1) ? 0 : 1;
      else {
        const int r = (k - n3) % tpb;
        kind = (r == 4) ? 2 : ((r == 0 || r == 2) ? 0 : 1);
      }
      if (kind == 0) {
        g[2 * k] = 0.5 * re;
        g[2 * k + 1] = 0.5 * im;
      } else if (kind == 1) {
        g[2 * k] = -0.5 * im;
        g[2 * k + 1] = 0.5 * re;
      } else {
        g[2 * k] = im;
        g[2 * k + 1] = -re;
      }
    }
  }
  return AQC_OK;
}

extern "C" int aqc_sv_grad(aqc_sv* sv, const double* thetas, int x_slot, int64_t x_basis,
                           int z0_slot, int w_slot, int z_slot, double* grad_out) {
  if (!grad_out) return fail(AQC_EINVAL, "null pointer argument");
  int rc = aqc_sv_grad_begin(sv, thetas, x_slot, x_basis, z0_slot, w_slot, z_slot);
  if (rc) return rc;
  return aqc_sv_grad_end(sv, grad_out);
}


// Host-only scheduler introspection (no device needed): serialises the compiled program as
// int32 words so that the CPU test-suite can replay it gate by gate against the oracle.
// Layout: npasses, then per pass {tb, nstages, nouter, bitpos[16], outerpos[48],
// per stage {p, q, nunits, (kind, flags, theta) x 3}}.
extern "C" int aqc_debug_program(const aqc_circuit* circ, int log2_cols, int tile_bits,
                                 int low_bits, int reversed, int32_t* out, int64_t cap,
                                 int64_t* needed) {
  if (!circ || !needed) return fail(AQC_EINVAL, "null argument");
  if (tile_bits < 2 || tile_bits > kMaxTileBits) return fail(AQC_EINVAL, "bad tile_bits");
  Program p;
  build_program(*circ, log2_cols, circ->n + log2_cols, tile_bits, low_bits, reversed != 0, p);
  std::vector<int32_t> w;
  w.push_back((int32_t)p.passes.size());
  for (const PassDesc& pd : p.passes) {
    w.push_back(pd.tb);
    w.push_back(pd.nstages);
    w.push_back(pd.nouter);
    for (int k = 0; k < 16; ++k) w.push_back(pd.bitpos[k]);
    for (int k = 0; k < 48; ++k) w.push_back(pd.outerpos[k]);
    for (int s = 0; s < pd.nstages; ++s) {
      const StageDesc& sd = p.stages[pd.stage0 + s];
      w.push_back(sd.p);
      w.push_back(sd.q);
      w.push_back(sd.nunits);
      for (int u = 0; u < kMaxUnits; ++u) {
        w.push_back(sd.u[u].kind);
        w.push_back(sd.u[u].flags);
        w.push_back(sd.u[u].theta);
      }
    }
  }
  *needed = (int64_t)w.size();
  if (out && cap >= (int64_t)w.size()) memcpy(out, w.data(), w.size() * sizeof(int32_t));
  return AQC_OK;
}


// Same for the dense-stage engine (up to kStageUnits units per stage, front gates merged), plus the
// shared-memory lane tables, so that the CPU test-suite can emulate the DMMA data flow.
// Layout: npasses, then per pass {tb, nstages, nouter, bitpos[16], outerpos[48], per stage
// {p, q, nunits, (kind, flags, theta) x kStageUnits, r0, r1, r2, (sl, so0, so1, sb) x 8 warps x 32}}.
static int debug_dense_program_impl(const aqc_circuit* circ, int log2_cols, int tile_bits, int low_bits,
                                    int reversed, int fuse_pairs, int32_t* out, int64_t cap, int64_t* needed);

extern "C" int aqc_debug_dense_program(const aqc_circuit* circ, int log2_cols, int tile_bits,
                                       int low_bits, int reversed, int32_t* out, int64_t cap,
                                       int64_t* needed) {
  return debug_dense_program_impl(circ, log2_cols, tile_bits, low_bits, reversed, 0, out, cap, needed);
}

// Same with the fused steps the production tables use (flags kPairFirst = 0x8000 / kPairSwap = 0x4000
// in the `sl` word of a step's first stage; the partner stage's own table is then unused).
extern "C" int aqc_debug_dense_program_fused(const aqc_circuit* circ, int log2_cols, int tile_bits,
                                             int low_bits, int reversed, int32_t* out, int64_t cap,
                                             int64_t* needed) {
  return debug_dense_program_impl(circ, log2_cols, tile_bits, low_bits, reversed, 1, out, cap, needed);
}

static int debug_dense_program_impl(const aqc_circuit* circ, int log2_cols, int tile_bits, int low_bits,
                                    int reversed, int fuse_pairs, int32_t* out, int64_t cap, int64_t* needed) {
  if (!circ || !needed) return fail(AQC_EINVAL, "null argument");
  if (tile_bits < kDMinTileBits || tile_bits > kMaxTileBits) return fail(AQC_EINVAL, "bad tile_bits");
  if (circ->n + log2_cols < kDMinTileBits) return fail(AQC_EINVAL, "state too small for the dense engine");
  Program p;
  build_program(*circ, log2_cols, circ->n + log2_cols, tile_bits, low_bits, reversed != 0, p,
                kStageUnits, true);
  DenseTables dt;
  if (!build_dense_tables(p, dt, fuse_pairs)) return fail(AQC_EINVAL, "tile too small for the dense engine");
  std::vector<int32_t> w;
  w.push_back((int32_t)p.passes.size());
  for (const PassDesc& pd : p.passes) {
    w.push_back(pd.tb);
    w.push_back(pd.nstages);
    w.push_back(pd.nouter);
    for (int k = 0; k < 16; ++k) w.push_back(pd.bitpos[k]);
    for (int k = 0; k < 48; ++k) w.push_back(pd.outerpos[k]);
    for (int s = 0; s < pd.nstages; ++s) {
      const StageDesc& sd = p.stages[pd.stage0 + s];
      w.push_back(sd.p);
      w.push_back(sd.q);
      w.push_back(sd.nunits);
      for (int u = 0; u < kStageUnits; ++u) {
        w.push_back(sd.u[u].kind);
        w.push_back(sd.u[u].flags);
        w.push_back(sd.u[u].theta);
      }
      for (int k = 0; k < 3; ++k) w.push_back(dt.rbits[(pd.stage0 + s) * 3 + k]);
      for (int i = 0; i < kDWarps * 32; ++i) {
        const DLane& d = dt.lanes[(size_t)(pd.stage0 + s) * kDWarps * 32 + i];
        w.push_back(d.sl);
        w.push_back(d.so0);
        w.push_back(d.so1);
        w.push_back(d.sb);
      }
    }
  }
  *needed = (int64_t)w.size();
  if (out && cap >= (int64_t)w.size()) memcpy(out, w.data(), w.size() * sizeof(int32_t));
  return AQC_OK;
}

// CUDA-event stopwatch on the workspace stream: brackets any sequence of calls on this
// workspace (bench.py times one objective + gradient step with it).
extern "C" int aqc_sv_timer_start(aqc_sv* sv) {
  if (!sv) return fail(AQC_EINVAL, "null workspace");
  CU(cudaSetDevice(sv->device));
  CU(cudaEventRecord(sv->tm0, sv->stream));
  return AQC_OK;
}
extern "C" int aqc_sv_timer_stop(aqc_sv* sv, float* ms) {
  if (!sv || !ms) return fail(AQC_EINVAL, "null argument");
  CU(cudaSetDevice(sv->device));
  CU(cudaEventRecord(sv->tm1, sv->stream));
  CU(cudaEventSynchronize(sv->tm1));
  CU(cudaEventElapsedTime(ms, sv->tm0, sv->tm1));
  return AQC_OK;
}


// ------------------------------------------------------------------------------------------
// epoch-wise execution (global-qubit sharding; a single-GPU workspace has exactly one epoch)
// ------------------------------------------------------------------------------------------
static const Program* prog_of(const aqc_sv* sv, int mode) {
  return mode == 0 ? &sv->prog_grad : (mode == 1 ? &sv->prog_fwd : &sv->prog_dag);
}

extern "C" int aqc_sv_num_epochs(const aqc_sv* sv, int mode) {
  if (!sv || mode < 0 || mode > 2) return AQC_EINVAL;
  return (int)prog_of(sv, mode)->epoch_pass0.size();
}

extern "C" int aqc_sv_epoch_layout(const aqc_sv* sv, int mode, int epoch) {
  if (!sv || mode < 0 || mode > 2) return AQC_EINVAL;
  const Program* p = prog_of(sv, mode);
  if (epoch < 0 || epoch >= (int)p->epoch_layout.size()) return AQC_EINVAL;
  return p->epoch_layout[epoch];
}

// Uploads thetas (cos/sin table) for a following sequence of aqc_sv_run_epoch calls; mode 0
// (gradient) also clears the raw inner-product accumulators.
extern "C" int aqc_sv_begin(aqc_sv* sv, const double* thetas, int mode) {
  if (!sv || !thetas || mode < 0 || mode > 2) return fail(AQC_EINVAL, "bad arguments");
  CU(cudaSetDevice(sv->device));
  sv->last_launches = 0;
  const size_t tot = (size_t)sv->batch * sv->circ.nthetas;
  int rc = ensure_pinned(sv, tot * 2 + 64);
  if (rc) return rc;
  rc = upload_thetas(sv, thetas);
  if (rc) return rc;
  if (sv->dense) {
    if (mode == 0) CU(cudaMemsetAsync(sv->d_gacc, 0, tot * 2 * sizeof(double), sv->stream));
    if ((rc = dense_prepare(sv, mode))) return rc;
  } else if (mode == 0) {
    CU(cudaMemsetAsync(sv->d_gacc, 0, tot * 2 * sizeof(double), sv->stream));
    if (!sv->legacy_grad && (rc = grad_prepare(sv))) return rc;
  } else if (!sv->legacy_grad) {
    if ((rc = apply_prepare(sv, mode == 2))) return rc;
  }
  CU(cudaStreamSynchronize(sv->stream));
  return AQC_OK;
}

// Runs the tile passes of one epoch.  mode 0: gradient on (vec0, vec1) = (w, z); mode 1 / 2:
// V / V^H on vec0.  src slots are read by the first pass only (src0 < 0: vec0 is the local part
// of a basis state: offset `basis_local`, or all zeros if basis_local < 0); dst slots receive the
// result and are updated in place by the remaining passes.
extern "C" int aqc_sv_run_epoch(aqc_sv* sv, int mode, int epoch, int src0, int64_t basis_local,
                                int src1, int dst0, int dst1) {
  if (!sv || mode < 0 || mode > 2) return fail(AQC_EINVAL, "bad arguments");
  const Program* p = prog_of(sv, mode);
  if (epoch < 0 || epoch >= (int)p->epoch_pass0.size()) return fail(AQC_EINVAL, "bad epoch");
  int rc = check_slot(sv, dst0);
  if (rc) return rc;
  if (src0 >= 0 && (rc = check_slot(sv, src0))) return rc;
  if (mode == 0) {
    if ((rc = check_slot(sv, dst1)) || (rc = check_slot(sv, src1))) return rc;
    if (dst0 == dst1) return fail(AQC_EINVAL, "w and z must be different slots");
  }
  if (src0 < 0 && mode != 0) return fail(AQC_EINVAL, "basis source is only valid for the gradient");
  CU(cudaSetDevice(sv->device));
  const int p0 = p->epoch_pass0[epoch];
  const int p1 = epoch + 1 < (int)p->epoch_pass0.size() ? p->epoch_pass0[epoch + 1] : (int)p->passes.size();
  const long long basis = src0 >= 0 ? -1 : (basis_local >= 0 ? (long long)basis_local : (1ll << 62));
  CU(cudaEventRecord(sv->ev0, sv->stream));
  if (sv->dense)
    rc = run_dense_program(sv, mode, src0 >= 0 ? sv->slots[src0] : nullptr, basis,
                           mode == 0 ? sv->slots[src1] : nullptr, sv->slots[dst0],
                           mode == 0 ? sv->slots[dst1] : nullptr, p0, p1);
  else if (mode == 0 && !sv->legacy_grad)
    rc = run_grad_program(sv, src0 >= 0 ? sv->slots[src0] : nullptr, basis, sv->slots[src1],
                          sv->slots[dst0], sv->slots[dst1], p0, p1);
  else if (!sv->legacy_grad)
    rc = run_apply_program(sv, mode == 2, sv->slots[src0], sv->slots[dst0], p0, p1);
  else
    rc = run_program(sv, *p, mode == 0, mode == 2, src0 >= 0 ? sv->slots[src0] : nullptr, basis,
                     mode == 0 ? sv->slots[src1] : nullptr, sv->slots[dst0],
                     mode == 0 ? sv->slots[dst1] : nullptr, p0, p1);
  if (rc) return rc;
  CU(cudaEventRecord(sv->ev1, sv->stream));
  CU(cudaStreamSynchronize(sv->stream));
  CU(cudaEventElapsedTime(&sv->last_ms, sv->ev0, sv->ev1));
  return AQC_OK;
}

// Downloads this workspace's (partial) raw inner products and converts them to 0.5j <P w|z>
// (linear, so partial sums of several ranks may be added afterwards).
extern "C" int aqc_sv_grad_finish(aqc_sv* sv, double* grad_out) {
  if (!sv || !grad_out) return fail(AQC_EINVAL, "bad arguments");
  CU(cudaSetDevice(sv->device));
  const size_t tot = (size_t)sv->batch * sv->circ.nthetas;
  int rc = ensure_pinned(sv, tot * 2 + 64);
  if (rc) return rc;
  if (sv->dense) {
    if ((rc = dense_collect(sv))) return rc;
  } else if (!sv->legacy_grad && (rc = grad_collect(sv))) {
    return rc;
  }
  CU(cudaMemcpyAsync(sv->h_pinned, sv->d_gacc, tot * 2 * sizeof(double), cudaMemcpyDeviceToHost,
                     sv->stream));
  CU(cudaStreamSynchronize(sv->stream));
  const int n3 = 3 * sv->circ.n, tpb = sv->circ.tpb, T = sv->circ.nthetas;
  for (int b = 0; b < sv->batch; ++b) {
    const double* raw = sv->h_pinned + (size_t)b * T * 2;
    double* g = grad_out + (size_t)b * T * 2;
    for (int k = 0; k < T; ++k) {
      const double re = raw[2 * k], im = raw[2 * k + 1];
      int kind;
      if (k < n3)
        kind = (k % 3 == 1) ? 0 : 1;
      else {
        const int r = (k - n3) % tpb;
        kind = (r == 4) ? 2 : ((r == 0 || r == 2) ? 0 : 1);
      }
      if (kind == 0)
        g[2 * k] = 0.5 * re, g[2 * k + 1] = 0.5 * im;
      else if (kind == 1)
        g[2 * k] = -0.5 * im, g[2 * k + 1] = 0.5 * re;
      else
        g[2 * k] = im, g[2 * k + 1] = -re;
    }
  }
  return AQC_OK;
}

// ---- layout switch: block transpose over the ranks through peer memory (NVLink P2P) ----------
extern "C" int aqc_sv_ipc_export(aqc_sv* sv, int slot, unsigned char* handle64) {
  int rc = check_slot(sv, slot);
  if (rc) return rc;
  if (!handle64) return fail(AQC_EINVAL, "null handle");
  CU(cudaSetDevice(sv->device));
  cudaIpcMemHandle_t h;
  CU(cudaIpcGetMemHandle(&h, sv->slots[slot]));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t size");
  memcpy(handle64, &h, 64);
  return AQC_OK;
}

extern "C" int aqc_sv_ipc_import(aqc_sv* sv, int peer_rank, int slot, const unsigned char* handle64) {
  int rc = check_slot(sv, slot);
  if (rc) return rc;
  if (!handle64 || peer_rank < 0 || peer_rank >= (1 << sv->g) || peer_rank >= 16)
    return fail(AQC_EINVAL, "bad peer rank");
  CU(cudaSetDevice(sv->device));
  if (peer_rank == sv->rank) {
    sv->peer[slot][peer_rank] = sv->slots[slot];
    return AQC_OK;
  }
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  sv->peer[slot][peer_rank] = (const double2*)p;
  return AQC_OK;
}

extern "C" int aqc_sv_peer_attach(aqc_sv* sv, int peer_rank, int slot, aqc_sv* peer) {
  int rc = check_slot(sv, slot);
  if (rc) return rc;
  if (!peer || peer_rank < 0 || peer_rank >= (1 << sv->g) || peer_rank >= 16 || slot >= peer->nslots)
    return fail(AQC_EINVAL, "bad peer");
  CU(cudaSetDevice(sv->device));
  if (peer->device != sv->device) {
    int can = 0;
    CU(cudaDeviceCanAccessPeer(&can, sv->device, peer->device));
    if (!can) return fail(AQC_ECUDA, "device %d cannot access device %d", sv->device, peer->device);
    cudaError_t e = cudaDeviceEnablePeerAccess(peer->device, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
      return fail(AQC_ECUDA, "cudaDeviceEnablePeerAccess failed: %s", cudaGetErrorString(e));
    cudaGetLastError();
  }
  sv->peer[slot][peer_rank] = peer->slots[slot];
  return AQC_OK;
}

struct ExchangeArgs {
  const double2* src[16];  // src[r] = rank r's source slot
  double2* dst;
  long long chunk;  // amplitudes per chunk
  int world, rank;
};

// dst[chunk r] = (rank r).src[chunk my_rank]: every rank pulls its column of the block matrix
// over NVLink with plain peer loads (coalesced 16-byte accesses) and stores locally.
__global__ void exchange_kernel(const ExchangeArgs A) {
  const int r = blockIdx.y;
  const double2* __restrict__ s = A.src[r] + (long long)A.rank * A.chunk;
  double2* __restrict__ d = A.dst + (long long)r * A.chunk;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < A.chunk; i += 4 * stride) {  // four peer loads in flight per thread
    const double2 v0 = s[i], v1 = s[i + stride], v2 = s[i + 2 * stride], v3 = s[i + 3 * stride];
    d[i] = v0, d[i + stride] = v1, d[i + 2 * stride] = v2, d[i + 3 * stride] = v3;
  }
  for (; i < A.chunk; i += stride) d[i] = s[i];
}

struct PushArgs {
  const double2* src;  // this rank's source slot
  double2* dst[16];    // dst[r] = rank r's destination slot
  long long chunk;
  int world, rank;
};

// The same block transpose as remote STORES: (rank r).dst[chunk my_rank] = src[chunk r].  Stores over
// NVLink are fire-and-forget, so the link is not throttled by outstanding read requests.
__global__ void exchange_push_kernel(const PushArgs A) {
  const int r = blockIdx.y;
  const double2* __restrict__ s = A.src + (long long)r * A.chunk;
  double2* __restrict__ d = A.dst[r] + (long long)A.rank * A.chunk;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < A.chunk; i += 4 * stride) {
    const double2 v0 = s[i], v1 = s[i + stride], v2 = s[i + 2 * stride], v3 = s[i + 3 * stride];
    d[i] = v0, d[i + stride] = v1, d[i + 2 * stride] = v2, d[i + 3 * stride] = v3;
  }
  for (; i < A.chunk; i += stride) d[i] = s[i];
}

extern "C" int aqc_sv_exchange(aqc_sv* sv, int src_slot, int dst_slot) {
  int rc = check_slot(sv, src_slot);
  if (rc) return rc;
  rc = check_slot(sv, dst_slot);
  if (rc) return rc;
  if (sv->g <= 0) return fail(AQC_EINVAL, "workspace is not sharded");
  if (src_slot == dst_slot) return fail(AQC_EINVAL, "exchange is out of place");
  CU(cudaSetDevice(sv->device));
  ExchangeArgs a;
  memset(&a, 0, sizeof(a));
  a.world = 1 << sv->g;
  a.rank = sv->rank;
  a.chunk = sv->size >> sv->g;
  a.dst = sv->slots[dst_slot];
  for (int r = 0; r < a.world; ++r) {
    a.src[r] = (r == sv->rank) ? sv->slots[src_slot] : sv->peer[src_slot][r];
    if (!a.src[r]) return fail(AQC_EINVAL, "peer %d slot %d was not imported", r, src_slot);
  }
  CU(cudaEventRecord(sv->ev0, sv->stream));
  // AQC_EXCHANGE = kernel (SM peer loads, default) | memcpy (one copy-engine transfer per peer chunk)
  static const int mode = [] {
    const char* e = getenv("AQC_EXCHANGE");
    return (e && std::string(e) == "memcpy") ? 1 : ((e && std::string(e) == "push") ? 2 : 0);
  }();
  if (mode == 2) {
    PushArgs pa;
    memset(&pa, 0, sizeof(pa));
    pa.src = sv->slots[src_slot];
    pa.chunk = a.chunk;
    pa.world = a.world;
    pa.rank = a.rank;
    for (int r = 0; r < a.world; ++r) {
      pa.dst[r] = (r == sv->rank) ? sv->slots[dst_slot] : const_cast<double2*>(sv->peer[dst_slot][r]);
      if (!pa.dst[r]) return fail(AQC_EINVAL, "peer %d slot %d was not imported", r, dst_slot);
    }
    const unsigned gx = (unsigned)std::min<long long>((a.chunk + 255) / 256, 148 * 4);
    exchange_push_kernel<<<dim3(gx, a.world), 256, 0, sv->stream>>>(pa);
    CU(cudaGetLastError());
  } else
  if (mode == 1) {
    for (int r = 0; r < a.world; ++r)
      CU(cudaMemcpyAsync(a.dst + (long long)r * a.chunk, a.src[r] + (long long)a.rank * a.chunk,
                         (size_t)a.chunk * sizeof(double2), cudaMemcpyDefault, sv->stream));
  } else {
    const unsigned gx = (unsigned)std::min<long long>((a.chunk + 255) / 256, 148 * 4);
    exchange_kernel<<<dim3(gx, a.world), 256, 0, sv->stream>>>(a);
    CU(cudaGetLastError());
  }
  CU(cudaEventRecord(sv->ev1, sv->stream));
  CU(cudaStreamSynchronize(sv->stream));
  CU(cudaEventElapsedTime(&sv->last_ms, sv->ev0, sv->ev1));
  sv->last_launches = 1;
  return AQC_OK;
}

// Sharded synthetic target: re, im ~ U[0,1) keyed on (seed, LOGICAL amplitude index) in layout A,
// identical for any number of ranks.  Not normalised: *norm2_out receives the local sum of squares
// (all-reduce it and call aqc_sv_scale).
__global__ void fill_random_logical_kernel(double2* __restrict__ v, long long size, int n, int g,
                                           int rank, unsigned long long seed, double* __restrict__ norm2) {
  const int nl = n - g, cb = nl - g;
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < size;
       i += (long long)gridDim.x * blockDim.x) {
    const unsigned long long lo = (unsigned long long)i & ((1ull << cb) - 1);
    const unsigned long long top = (unsigned long long)i >> cb;  // qubits 0..g-1
    const unsigned long long logical = ((unsigned long long)rank << nl) | (lo << g) | top;
    const double re = u01(seed, 2ull * logical), im = u01(seed, 2ull * logical + 1);
    v[i] = make_double2(re, im);
    acc = fma(re, re, acc);
    acc = fma(im, im, acc);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(norm2, acc);
}

extern "C" int aqc_sv_fill_random_logical(aqc_sv* sv, int slot, uint64_t seed, double* norm2_out) {
  int rc = check_slot(sv, slot);
  if (rc) return rc;
  if (!norm2_out || sv->batch != 1 || sv->log2_cols != 0) return fail(AQC_EINVAL, "bad arguments");
  CU(cudaSetDevice(sv->device));
  rc = ensure_scratch(sv, 8);
  if (rc) return rc;
  CU(cudaMemsetAsync(sv->d_scratch, 0, sizeof(double), sv->stream));
  const unsigned gx = (unsigned)std::min<long long>((sv->size + 255) / 256, 148 * 16);
  fill_random_logical_kernel<<<gx, 256, 0, sv->stream>>>(sv->slots[slot], sv->size, sv->circ.n, sv->g,
                                                        sv->rank, seed, sv->d_scratch);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(norm2_out, sv->d_scratch, sizeof(double), cudaMemcpyDeviceToHost, sv->stream));
  CU(cudaStreamSynchronize(sv->stream));
  return AQC_OK;
}

__global__ void scale_const_kernel(double2* __restrict__ v, long long total, double f) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    double2 x = v[i];
    x.x *= f;
    x.y *= f;
    v[i] = x;
  }
}

extern "C" int aqc_sv_scale(aqc_sv* sv, int slot, double factor) {
  int rc = check_slot(sv, slot);
  if (rc) return rc;
  CU(cudaSetDevice(sv->device));
  const long long total = sv->size * sv->batch;
  const unsigned gx = (unsigned)std::min<long long>((total + 255) / 256, 148 * 16);
  scale_const_kernel<<<gx, 256, 0, sv->stream>>>(sv->slots[slot], total, factor);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(sv->stream));
  return AQC_OK;
}

// Host-only: epoch plan of the sharded scheduler (no device needed).  Serialises, per epoch, the
// layout id and the tile-pass program (same word layout as aqc_debug_program) so that the CPU
// test-suite can replay a sharded run rank by rank.  mode: 0 / 1 forward, 2 reversed (V^H).
extern "C" int aqc_debug_program_sharded(const aqc_circuit* circ, int log2_world, int tile_bits,
                                         int low_bits, int reversed, int32_t* out, int64_t cap,
                                         int64_t* needed) {
  if (!circ || !needed) return fail(AQC_EINVAL, "null argument");
  if (tile_bits < 2 || tile_bits > kMaxTileBits) return fail(AQC_EINVAL, "bad tile_bits");
  Program p;
  std::string err;
  if (build_program_sharded(*circ, log2_world, tile_bits, low_bits, reversed != 0, p, err))
    return fail(AQC_EINVAL, "%s", err.c_str());
  std::vector<int32_t> w;
  const int ne = (int)p.epoch_pass0.size();
  w.push_back(ne);
  for (int e = 0; e < ne; ++e) {
    const int p0 = p.epoch_pass0[e], p1 = e + 1 < ne ? p.epoch_pass0[e + 1] : (int)p.passes.size();
    w.push_back(p.epoch_layout[e]);
    w.push_back(p1 - p0);
    for (int i = p0; i < p1; ++i) {
      const PassDesc& pd = p.passes[i];
      w.push_back(pd.tb);
      w.push_back(pd.nstages);
      w.push_back(pd.nouter);
      for (int k = 0; k < 16; ++k) w.push_back(pd.bitpos[k]);
      for (int k = 0; k < 48; ++k) w.push_back(pd.outerpos[k]);
      for (int s = 0; s < pd.nstages; ++s) {
        const StageDesc& sd = p.stages[pd.stage0 + s];
        w.push_back(sd.p);
        w.push_back(sd.q);
        w.push_back(sd.nunits);
        for (int u = 0; u < kMaxUnits; ++u) {
          w.push_back(sd.u[u].kind);
          w.push_back(sd.u[u].flags);
          w.push_back(sd.u[u].theta);
        }
      }
    }
  }
  *needed = (int64_t)w.size();
  if (out && cap >= (int64_t)w.size()) memcpy(out, w.data(), w.size() * sizeof(int32_t));
  return AQC_OK;
}
