// aqc_mps.cu -- matrix-product-state objective-and-gradient engine for sm_100a.
//
// What this replaces (reference = qiskit-community/aqc-research, paths relative to its root):
//   mps_dot                         aqc_research/mps_operations.py:192-213
//   v_mul_mps / v_dagger_mul_mps    aqc_research/mps_operations.py:326-371 (gate arithmetic inside
//                                   qiskit-aer's C++ MPS simulator, one simulator run per call)
//   fast_dot_gradient               aqc_research/mps_dot_objective.py:41-242 (one qiskit-aer run per
//                                   gate and one FULL-chain mps_dot per parameter)
//   MpsStateHandler.state_dot_vector aqc_research/model_sp_lhs/objective_base.py:400-403
//
// Design.  States live on the device in Vidal form (Gamma tensors padded to
// the bond capacity C, lambda vectors, bond dimensions -- all device resident, so a whole sweep is
// enqueued without a single host synchronisation).  Gates of a pair-run (a Trotter triplet) act on
// the physical indices of one two-site tensor only, therefore
//   * the two-site tensor is formed ONCE per pair-run, the 4x4 product of the run's gates is
//     applied, and ONE Jacobi SVD splits it again (the reference does 3 SVDs per triplet);
//   * all 12 derivatives of the run follow from ONE 4x4 reduced overlap matrix
//     rho[b', b] = <w-part b' | E_L . E_R | z-part b> built with cached left / right environments
//     (O(chi^3)), instead of 12 full-chain contractions (O(n chi^3) each);
//   * pair-runs of a half-layer touch disjoint sites, so every kernel is batched over them, and
//     the environments at the bonds a half-layer modifies are refreshed by ONE batched single-site
//     transfer step (unitaries applied to both states leave environments across untouched bonds
//     invariant).
// Truncation follows the published qiskit-aer rule (see oracle/mps_oracle.py); truncated results
// are "parity unpinned" against qiskit-aer, untruncated ones equal the state-vector path.

#include <cooperative_groups.h>
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

int aqc_fail(int code, const char* fmt, ...);  // aqc_sv.cu
#define MCU(call)                                                                                \
  do {                                                                                           \
    cudaError_t e_ = (call);                                                                     \
    if (e_ != cudaSuccess)                                                                       \
      return aqc_fail(AQC_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, \
                      __LINE__);                                                                 \
  } while (0)

constexpr int kMaxChi = 64;
constexpr double kChop = 1e-16;  // singular values <= this are exact zeros (qiskit-aer CHOP)

// One pair-run (or one front gate) of the compiled MPS program.
struct MpsTask {
  int32_t site;    // left site of the pair (or the site of a front gate)
  int32_t nunits;  // 1..3 units
  int32_t kind[3];
  int32_t flags[3];
  int32_t theta[3];
  int32_t pad;
};

struct MpsStep {
  int task0, ntasks;  // range in the task array
};

struct EnvTask {
  int32_t site;  // site being absorbed
  int32_t dir;   // 0: L (E at bond site -> bond site+1), 1: R (E at bond site+1 -> bond site)
};

struct MpsProgram {
  std::vector<MpsTask> tasks;  // [front tasks | step tasks ...]
  std::vector<MpsStep> steps;  // two-qubit steps in execution order
  int nfront = 0;
  MpsTask* d_tasks = nullptr;
  // environment refresh lists: [L sweep 0..n-1 | R sweep n-1..0 | per step: (k, L), (k+1, R) ...]
  std::vector<EnvTask> env;
  std::vector<int> env_step0;  // offset of every step's list
  EnvTask* d_env = nullptr;
};

struct MpsState {
  double2* gam = nullptr;  // [n][2][C][C]
  double* lam = nullptr;   // [n+1][C]  bond j sits left of site j; bonds 0 and n have dimension 1
  int* dims = nullptr;     // [n+1]
};

struct aqc_circuit;  // defined in aqc_sv.cu
// accessors implemented in aqc_sv.cu (keeps struct aqc_circuit private to that file)
int aqc_circ_n(const aqc_circuit* c);
int aqc_circ_ent(const aqc_circuit* c);
int aqc_circ_trotter(const aqc_circuit* c);
int aqc_circ_nb(const aqc_circuit* c);
int aqc_circ_half(const aqc_circuit* c);
int aqc_circ_tpb(const aqc_circuit* c);
int aqc_circ_nthetas(const aqc_circuit* c);
int aqc_circ_ctrl(const aqc_circuit* c, int i);
int aqc_circ_targ(const aqc_circuit* c, int i);

struct aqc_mps {
  int device = 0;
  int n = 0, C = 0, ent = 0, tpb = 4, nthetas = 0;
  int chi_max = 0;
  double trunc_thr = 1e-16;
  int nslots = 0;
  int maxtasks = 0;  // max tasks in one step (>= n for the front layer)
  std::vector<MpsState> st;
  MpsProgram fwd, dag;
  // scratch
  double* d_thetas = nullptr;
  double2* d_gate = nullptr;    // [maxtasks][16] 4x4 gate of each task of the current step
  double2* d_theta0 = nullptr;  // [2][maxtasks][4][C][C]  two-site tensors before the gate
  double2* d_work = nullptr;    // [2][maxtasks][2C][2C]   SVD working matrices (column-major)
  double2* d_vmat = nullptr;    // [2][maxtasks][2C][2C]   right singular vectors of the kept columns
  double2* d_work0 = nullptr;   // [2][maxtasks][2C][2C]   working matrices before the rotations
  double2* d_envL = nullptr;    // [n+1][C][C]
  double2* d_envR = nullptr;    // [n+1][C][C]
  double2* d_rho = nullptr;     // [maxtasks][16]
  double* d_gacc = nullptr;     // [nthetas] complex raw sums
  double2* d_small = nullptr;   // small outputs
  long long* d_idx = nullptr;
  double* d_trunc = nullptr;   // [4] truncation record of the current call (see SvdArgs)
  int* d_sweeps = nullptr;     // [2][maxtasks]
  int* d_conv = nullptr;       // [2][maxtasks][32]
  int num_sms = 148;
  bool svd_precond = true;    // AQC_MPS_SVD=plain: Jacobi directly on the working matrix
  bool svd_fence = false;     // AQC_MPS_FENCE=1
  int svd_cluster = 0;        // AQC_MPS_CLUSTER=1|2|4: CTAs per SVD instead of "as many as fit one wave"
  bool svd_relaxed = true;    // AQC_MPS_SVD_TOL=strict: rounding-level Jacobi convergence also when truncating
  bool theta_scalar = false;  // AQC_MPS_THETA=scalar: thread-per-column contraction instead of the DMMA GEMM
  double* h_pinned = nullptr;
  size_t pinned_cap = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  float last_ms = 0.f;
  int last_launches = 0;
};

// ------------------------------------------------------------------------------------------
// host: program construction
// ------------------------------------------------------------------------------------------
static int build_mps_program(const aqc_circuit* c, bool reversed, MpsProgram& prog, std::string& err) {
  const int n = aqc_circ_n(c), nb = aqc_circ_nb(c), half = aqc_circ_half(c), tpb = aqc_circ_tpb(c);
  const bool trot = aqc_circ_trotter(c) != AQC_GENERIC;
  prog.tasks.clear();
  prog.steps.clear();
  for (int q = 0; q < n; ++q) {
    MpsTask t;
    memset(&t, 0, sizeof(t));
    t.site = q;
    t.nunits = 1;
    t.kind[0] = U_FRONT_LO;
    t.theta[0] = 3 * q;
    prog.tasks.push_back(t);
  }
  prog.nfront = n;
  // blocks in circuit order -> steps of pair-runs on disjoint pairs
  struct Blk {
    int lo, kind, flags, theta;
  };
  std::vector<Blk> blks;
  // A unit-block on non-adjacent qubits (the reference accepts any (ctrl, targ):
  // mps_dot_objective.py:380-468, test_mps_fast_dot_gradient.py:119-153) runs through a swap network:
  // adjacent SWAPs bring the upper qubit down next to the lower one, the block acts on the
  // neighbouring sites, and the SWAPs are undone.  A SWAP that would directly follow the same SWAP
  // (back-to-back blocks on one distant pair) cancels against it.
  auto push_swap = [&](int lo) {
    if (!blks.empty() && blks.back().kind == U_SWAP && blks.back().lo == lo)
      blks.pop_back();
    else
      blks.push_back({lo, U_SWAP, 0, 0});
  };
  for (int i = 0; i < nb + half; ++i) {
    const int k = nb > 0 ? i % nb : 0;
    const int cq = aqc_circ_ctrl(c, k), tq = aqc_circ_targ(c, k);
    const int lo = std::min(cq, tq), hi = std::max(cq, tq);
    if (hi - lo != 1 && trot) {
      err = "a Trotterized ansatz has unit-blocks on adjacent qubits only";
      return AQC_EINVAL;
    }
    int flags = 0;
    if (trot && i % 3 == 0) flags |= F_PRE;
    if (trot && i % 3 == 2) flags |= F_POST;
    for (int s = hi - 1; s > lo; --s) push_swap(s);  // qubit `hi` travels down to site lo + 1
    blks.push_back({lo, cq > tq ? U_BLOCK_CHI : U_BLOCK_CLO, flags, 3 * n + tpb * k});
    for (int s = lo + 1; s < hi; ++s) push_swap(s);  // ... and back
  }
  // level[s] = index of the last step that touched site s; open task of that site (if the last
  // thing that touched both sites of a pair is the same task, a block can be chained onto it)
  std::vector<int> last_step(n, -1), last_task(n, -1);
  std::vector<std::vector<MpsTask>> steps;
  for (const Blk& b : blks) {
    const int s0 = b.lo, s1 = b.lo + 1;
    int step = -1, task = -1;
    if (last_step[s0] >= 0 && last_step[s0] == last_step[s1] && last_task[s0] == last_task[s1] &&
        last_task[s0] >= 0) {
      MpsTask& t = steps[last_step[s0]][last_task[s0]];
      if (t.site == b.lo && t.nunits < 3) {
        step = last_step[s0];
        task = last_task[s0];
      }
    }
    if (task < 0) {
      step = std::max(last_step[s0], last_step[s1]) + 1;
      if ((int)steps.size() <= step) steps.resize(step + 1);
      MpsTask t;
      memset(&t, 0, sizeof(t));
      t.site = b.lo;
      steps[step].push_back(t);
      task = (int)steps[step].size() - 1;
    }
    MpsTask& t = steps[step][task];
    t.kind[t.nunits] = b.kind;
    t.flags[t.nunits] = b.flags;
    t.theta[t.nunits] = b.theta;
    t.nunits++;
    last_step[s0] = last_step[s1] = step;
    last_task[s0] = last_task[s1] = task;
  }
  if (reversed) {
    std::reverse(steps.begin(), steps.end());
    for (auto& s : steps)
      for (auto& t : s) {
        std::reverse(t.kind, t.kind + t.nunits);
        std::reverse(t.flags, t.flags + t.nunits);
        std::reverse(t.theta, t.theta + t.nunits);
      }
  }
  for (auto& s : steps) {
    prog.steps.push_back({(int)prog.tasks.size(), (int)s.size()});
    for (auto& t : s) prog.tasks.push_back(t);
  }
  prog.env.clear();
  prog.env_step0.clear();
  for (int k = 0; k < n; ++k) prog.env.push_back({k, 0});
  for (int k = n - 1; k >= 0; --k) prog.env.push_back({k, 1});
  for (auto& st : prog.steps) {
    prog.env_step0.push_back((int)prog.env.size());
    for (int i = 0; i < st.ntasks; ++i) {
      const int k = prog.tasks[st.task0 + i].site;
      prog.env.push_back({k, 0});      // E_L(k+1) from E_L(k) and the new site k
      prog.env.push_back({k + 1, 1});  // E_R(k+1) from E_R(k+2) and the new site k+1
    }
  }
  return AQC_OK;
}

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double2 cmul2(double2 a, double2 b) {
  return make_double2(fma(-a.y, b.y, a.x * b.x), fma(a.y, b.x, a.x * b.y));
}
__device__ __forceinline__ void cfma(double2& acc, double2 a, double2 b) {  // acc += a * b
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.y = fma(a.y, b.x, acc.y);
}
__device__ __forceinline__ void cfma_conj(double2& acc, double2 a, double2 b) {  // acc += conj(a) * b
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.y = fma(-a.y, b.x, acc.y);
}

struct StateView {
  const double2* gam;
  const double* lam;
  const int* dims;
};
struct StateMut {
  double2* gam;
  double* lam;
  int* dims;
};

// ------------------------------------------------------------------------------------------
// kernel: 4x4 algebra of one task -- gate product, and (gradient) the derivatives from rho
// ------------------------------------------------------------------------------------------
// mode 0: write the task's 4x4 gate (forward product)         -> gate[task][16]
// mode 1: same for the daggered product (units already reversed by the host)
// mode 2: gradient: derivatives of all units from rho[task] (raw sums, same convention as the
//         state-vector kernel) are ADDED to gacc; one thread per (task, basis vector j)
template <int ENT>
__global__ void mps_algebra_kernel(const MpsTask* __restrict__ tasks, int ntasks,
                                   const double* __restrict__ thetas, int mode,
                                   double2* __restrict__ gate, const double2* __restrict__ rho,
                                   double* __restrict__ gacc) {
  const int t = blockIdx.x * (blockDim.x / 4) + threadIdx.x / 4;
  const int j = threadIdx.x & 3;
  if (t >= ntasks) return;  // the 4 threads of a task leave together
  const unsigned gmask = 0xFu << ((threadIdx.x & 31) & ~3);
  const MpsTask tk = tasks[t];
  constexpr int NP = (ENT == AQC_ENT_CP) ? 5 : 4;
  if (mode < 2) {
    cd a[1][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[0][i].x = (i == j) ? 1.0 : 0.0, a[0][i].y = 0.0;
    for (int u = 0; u < tk.nunits; ++u) {
      double2 tr[5];
      const int kind = tk.kind[u];
      if (kind == U_SWAP) {  // parameter-free, self-inverse
        const cd t = a[0][1];
        a[0][1] = a[0][2];
        a[0][2] = t;
        continue;
      }
      const int np = (kind == U_FRONT_LO || kind == U_FRONT_HI) ? 3 : NP;
      for (int k = 0; k < np; ++k) {
        const double th = thetas[tk.theta[u] + k];
        double s, c;
        sincos((k == 4) ? th : 0.5 * th, &s, &c);
        tr[k] = make_double2(c, s);
      }
      if (mode == 0) {
        if (kind == U_FRONT_LO) front_unit<1, false, false>(a, tr, nullptr);
        if (kind == U_BLOCK_CHI) block_unit<1, ENT, true, false>(a, tr, tk.flags[u], nullptr);
        if (kind == U_BLOCK_CLO) block_unit<1, ENT, false, false>(a, tr, tk.flags[u], nullptr);
      } else {
        if (kind == U_FRONT_LO) front_unit<1, false, true>(a, tr, nullptr);
        if (kind == U_BLOCK_CHI) block_unit<1, ENT, true, true>(a, tr, tk.flags[u], nullptr);
        if (kind == U_BLOCK_CLO) block_unit<1, ENT, false, true>(a, tr, tk.flags[u], nullptr);
      }
    }
    // column j of the gate: gate[row i][col j]
#pragma unroll
    for (int i = 0; i < 4; ++i) gate[(size_t)t * 16 + i * 4 + j] = make_double2(a[0][i].x, a[0][i].y);
  } else {
    // rho[b'][b] = <w-part b'|z-part b>  =  sum_j conj(e_j)[b'] (rho row j)[b]:
    // run the two-vector gradient code on (w, z) = (e_j, rho[j, :]) and add the four partial sums
    cd a[2][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      a[0][i].x = (i == j) ? 1.0 : 0.0;
      a[0][i].y = 0.0;
      const double2 r = rho[(size_t)t * 16 + j * 4 + i];
      a[1][i].x = r.x;
      a[1][i].y = r.y;
    }
    for (int u = 0; u < tk.nunits; ++u) {
      double2 tr[5];
      double acc[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) acc[k] = 0.0;
      const int kind = tk.kind[u];
      if (kind == U_SWAP) {  // both swept vectors go through the SWAP; nothing to differentiate
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          const cd t = a[v][1];
          a[v][1] = a[v][2];
          a[v][2] = t;
        }
        continue;
      }
      const bool front = (kind == U_FRONT_LO || kind == U_FRONT_HI);
      const int np = front ? 3 : NP;
      for (int k = 0; k < np; ++k) {
        const double th = thetas[tk.theta[u] + k];
        double s, c;
        sincos((k == 4) ? th : 0.5 * th, &s, &c);
        tr[k] = make_double2(c, s);
      }
      if (kind == U_FRONT_LO) front_unit<2, false, false>(a, tr, acc);
      if (kind == U_BLOCK_CHI) block_unit<2, ENT, true, false>(a, tr, tk.flags[u], acc);
      if (kind == U_BLOCK_CLO) block_unit<2, ENT, false, false>(a, tr, tk.flags[u], acc);
      const int nval = 2 * np;
      for (int k = 0; k < nval; ++k) {
        double v = acc[k];
        v += __shfl_xor_sync(gmask, v, 1);
        v += __shfl_xor_sync(gmask, v, 2);
        if (j == 0) atomicAdd(gacc + 2 * (size_t)tk.theta[u] + k, v);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// kernel: two-site tensor of a task (+ gate) -> SVD working matrix; optional copy before the gate
// ------------------------------------------------------------------------------------------
// grid (ntasks, nstates), 512 threads.  thread: column gamma = tid % 64, rows alpha = 8 per thread.
struct ThetaArgs {
  StateView st[2];
  const MpsTask* tasks;
  const double2* gate;  // [task][16] or nullptr (identity)
  double2* theta0;      // [state][maxtasks][4][C][C] or nullptr
  double2* work;        // [state][maxtasks][2C*2C] or nullptr
  double2* work0;       // same layout: untouched copy of the working matrix (V is recovered from it)
  int C, maxtasks, single_site;
};

__global__ void __launch_bounds__(512) mps_theta_kernel(const ThetaArgs A) {
  const int t = blockIdx.x, s = blockIdx.y, C = A.C;
  const MpsTask tk = A.tasks[t];
  const StateView S = A.st[s];
  const int k = tk.site;
  const int cl = S.dims[k], cm = S.dims[k + 1];
  const int cr = A.single_site ? cm : S.dims[k + 2];
  const int g = threadIdx.x & 63, grp = threadIdx.x >> 6;
  const double2* Ga = S.gam + (size_t)k * 2 * C * C;
  if (A.single_site) {
    // theta0[b][alpha][gamma] = Gamma_k[b][alpha, gamma] * lambda_{k+1}[gamma]   (b = 0, 1)
    double2* out = A.theta0 + ((size_t)s * A.maxtasks + t) * 4 * C * C;
    const double lr = (g < cr) ? S.lam[(size_t)(k + 1) * C + g] : 0.0;
    for (int r = 0; r < 8; ++r) {
      const int al = grp * 8 + r;
      if (al < cl && g < cr) {
        for (int b = 0; b < 2; ++b) {
          const double2 v = Ga[((size_t)b * C + al) * C + g];
          out[((size_t)b * C + al) * C + g] = make_double2(v.x * lr, v.y * lr);
        }
      }
    }
    return;
  }
  const double2* Gb = S.gam + (size_t)(k + 1) * 2 * C * C;
  const double* lamL = S.lam + (size_t)k * C;
  const double* lamM = S.lam + (size_t)(k + 1) * C;
  const double* lamR = S.lam + (size_t)(k + 2) * C;
  double2 acc[8][4];
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = make_double2(0.0, 0.0);
  const bool colok = g < cr;
  for (int be = 0; be < cm; ++be) {
    const double lm = lamM[be];
    double2 b0 = make_double2(0.0, 0.0), b1 = b0;
    if (colok) {
      b0 = Gb[((size_t)0 * C + be) * C + g];
      b1 = Gb[((size_t)1 * C + be) * C + g];
      b0.x *= lm, b0.y *= lm, b1.x *= lm, b1.y *= lm;
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int al = grp * 8 + r;
      if (al < cl) {
        const double2 a0 = Ga[((size_t)0 * C + al) * C + be];
        const double2 a1 = Ga[((size_t)1 * C + al) * C + be];
        // quad index = (b2 << 1) | b1   (hi = site k+1, lo = site k)
        cfma(acc[r][0], a0, b0);
        cfma(acc[r][1], a1, b0);
        cfma(acc[r][2], a0, b1);
        cfma(acc[r][3], a1, b1);
      }
    }
  }
  if (!colok) return;
  const double lr = lamR[g];
  double2 G4[16];
  if (A.gate) {
#pragma unroll
    for (int i = 0; i < 16; ++i) G4[i] = A.gate[(size_t)t * 16 + i];
  }
  const int M = 2 * cl, N = 2 * cr, LD = 2 * C;
  const bool transposed = M < N;
  double2* T0 = A.theta0 ? A.theta0 + ((size_t)s * A.maxtasks + t) * 4 * C * C : nullptr;
  double2* W = A.work ? A.work + ((size_t)s * A.maxtasks + t) * (size_t)LD * LD : nullptr;
  double2* W0 = A.work ? A.work0 + ((size_t)s * A.maxtasks + t) * (size_t)LD * LD : nullptr;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int al = grp * 8 + r;
    if (al >= cl) continue;
    const double ll = lamL[al];
    double2 v[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) v[c] = make_double2(acc[r][c].x * lr, acc[r][c].y * lr);
    if (T0) {
#pragma unroll
      for (int c = 0; c < 4; ++c) T0[((size_t)c * C + al) * C + g] = v[c];
    }
    if (W) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {  // output quad index c = (b2' << 1) | b1'
        double2 o = make_double2(0.0, 0.0);
        if (A.gate) {
#pragma unroll
          for (int d = 0; d < 4; ++d) cfma(o, G4[c * 4 + d], v[d]);
        } else {
          o = v[c];
        }
        o.x *= ll, o.y *= ll;
        const int row = (c & 1) * cl + al, col = (c >> 1) * cr + g;
        if (!transposed) {
          W[row + (size_t)col * LD] = o;
          W0[row + (size_t)col * LD] = o;
        } else {
          W[col + (size_t)row * LD] = make_double2(o.x, -o.y);
          W0[col + (size_t)row * LD] = make_double2(o.x, -o.y);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// kernel: the same two-site tensor as a tensor-core GEMM (chi >= 32 makes it a genuine dense
// complex contraction):  theta[(b1, alpha), (b2, gamma)] = sum_beta GammaA[b1][alpha, beta]
// lambda_m[beta] GammaB[b2][beta, gamma], real form on mma.sync.m8n8k4.f64 (DMMA).
// CTA tile: 32 alpha x 32 gamma x the four (b1, b2) combinations = 64 x 64 outputs; warp w owns
// alpha0 + 4w .. 4w + 3 for BOTH b1 (fragment rows 0-3 / 4-7), so the 4x4 gate on (b1, b2) needs one
// shuffle (lane ^ 16) and no staging.  grid (ntasks * 4, nstates), 256 threads.
// ------------------------------------------------------------------------------------------
constexpr int kThLdA = 20, kThLdB = 68;

__device__ __forceinline__ void mps_dmma884(double& d0, double& d1, const double a, const double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256) mps_theta_dmma_kernel(const ThetaArgs A) {
  __shared__ double sAr[64 * kThLdA], sAi[64 * kThLdA];
  __shared__ double sBr[16 * kThLdB], sBi[16 * kThLdB];
  const int t = blockIdx.x >> 2, tile = blockIdx.x & 3, s = blockIdx.y, C = A.C;
  const MpsTask tk = A.tasks[t];
  const StateView S = A.st[s];
  const int k = tk.site;
  const int cl = S.dims[k], cm = S.dims[k + 1], cr = S.dims[k + 2];
  const int al0 = (tile & 1) * 32, ga0 = (tile >> 1) * 32;
  if (al0 >= cl || ga0 >= cr) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double2* Ga = S.gam + (size_t)k * 2 * C * C;
  const double2* Gb = S.gam + (size_t)(k + 1) * 2 * C * C;
  const double* lamL = S.lam + (size_t)k * C;
  const double* lamM = S.lam + (size_t)(k + 1) * C;
  const double* lamR = S.lam + (size_t)(k + 2) * C;
  double cre[8][2], cim[8][2];
#pragma unroll
  for (int q = 0; q < 8; ++q) cre[q][0] = cre[q][1] = cim[q][0] = cim[q][1] = 0.0;

  for (int be0 = 0; be0 < cm; be0 += 16) {
    // A planes: smem row i = 8 w + ri  <->  b1 = ri >> 2, alpha = al0 + 4 w + (ri & 3)
    for (int e = tid; e < 64 * 16; e += 256) {
      const int kk = e & 15, i = e >> 4;
      const int b1 = (i & 7) >> 2, al = al0 + 4 * (i >> 3) + (i & 3);
      double2 v = make_double2(0.0, 0.0);
      if (al < cl && be0 + kk < cm) v = Ga[((size_t)b1 * C + al) * C + be0 + kk];
      sAr[i * kThLdA + kk] = v.x;
      sAi[i * kThLdA + kk] = v.y;
    }
    // B planes: column j  <->  b2 = j >> 5, gamma = ga0 + (j & 31); lambda_m folded in
    for (int e = tid; e < 16 * 64; e += 256) {
      const int j = e & 63, kk = e >> 6;
      const int b2 = j >> 5, ga = ga0 + (j & 31);
      double2 v = make_double2(0.0, 0.0);
      if (ga < cr && be0 + kk < cm) {
        v = Gb[((size_t)b2 * C + be0 + kk) * C + ga];
        const double lm = lamM[be0 + kk];
        v.x *= lm, v.y *= lm;
      }
      sBr[kk * kThLdB + j] = v.x;
      sBi[kk * kThLdB + j] = v.y;
    }
    __syncthreads();
#pragma unroll
    for (int ks = 0; ks < 16; ks += 4) {
      const int ai = (8 * warp + (lane >> 2)) * kThLdA + ks + (lane & 3);
      const double ar = sAr[ai], aim = sAi[ai], nai = -aim;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int bi = (ks + (lane & 3)) * kThLdB + 8 * q + (lane >> 2);
        const double br = sBr[bi], bim = sBi[bi];
        mps_dmma884(cre[q][0], cre[q][1], ar, br);
        mps_dmma884(cre[q][0], cre[q][1], nai, bim);
        mps_dmma884(cim[q][0], cim[q][1], ar, bim);
        mps_dmma884(cim[q][0], cim[q][1], aim, br);
      }
    }
    __syncthreads();
  }

  // epilogue: lane holds theta for (b1 = ri >> 2, alpha) and, per n-tile q, (b2 = q >> 2, two gammas)
  const int ri = lane >> 2, b1 = ri >> 2, al = al0 + 4 * warp + (ri & 3);
  const int M = 2 * cl, N = 2 * cr, LD = 2 * C;
  const bool transposed = M < N;
  double2* T0 = A.theta0 ? A.theta0 + ((size_t)s * A.maxtasks + t) * 4 * C * C : nullptr;
  double2* W = A.work ? A.work + ((size_t)s * A.maxtasks + t) * (size_t)LD * LD : nullptr;
  double2* W0 = A.work ? A.work0 + ((size_t)s * A.maxtasks + t) * (size_t)LD * LD : nullptr;
  // gate rows of the two output combinations this lane produces: c' = (b2' << 1) | b1
  double2 Grow[2][4];
#pragma unroll
  for (int b2 = 0; b2 < 2; ++b2)
#pragma unroll
    for (int d = 0; d < 4; ++d)
      Grow[b2][d] = A.gate ? A.gate[(size_t)t * 16 + ((b2 << 1) | b1) * 4 + d]
                           : make_double2((d == ((b2 << 1) | b1)) ? 1.0 : 0.0, 0.0);
  const double ll = (al < cl) ? lamL[al] : 0.0;
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int ga = ga0 + 8 * q + 2 * (lane & 3) + h;
      const double lr = (ga < cr) ? lamR[ga] : 0.0;
      double2 mine[2], other[2];  // index: b2
      mine[0] = make_double2(cre[q][h] * lr, cim[q][h] * lr);
      mine[1] = make_double2(cre[q + 4][h] * lr, cim[q + 4][h] * lr);
#pragma unroll
      for (int b2 = 0; b2 < 2; ++b2) {
        other[b2].x = __shfl_xor_sync(0xffffffffu, mine[b2].x, 16);
        other[b2].y = __shfl_xor_sync(0xffffffffu, mine[b2].y, 16);
      }
      if (al >= cl || ga >= cr) continue;
      double2 v[4];  // quad index c = (b2 << 1) | b1
#pragma unroll
      for (int b2 = 0; b2 < 2; ++b2) {
        v[(b2 << 1) | 0] = b1 ? other[b2] : mine[b2];
        v[(b2 << 1) | 1] = b1 ? mine[b2] : other[b2];
      }
      if (T0) {
#pragma unroll
        for (int b2 = 0; b2 < 2; ++b2) T0[((size_t)((b2 << 1) | b1) * C + al) * C + ga] = mine[b2];
      }
      if (W) {
#pragma unroll
        for (int b2 = 0; b2 < 2; ++b2) {
          const int c = (b2 << 1) | b1;  // the two output combinations with b1' = this lane's b1
          double2 o = make_double2(0.0, 0.0);
#pragma unroll
          for (int d = 0; d < 4; ++d) cfma(o, Grow[b2][d], v[d]);
          o.x *= ll, o.y *= ll;
          const int row = (c & 1) * cl + al, col = (c >> 1) * cr + ga;
          if (!transposed) {
            W[row + (size_t)col * LD] = o;
            W0[row + (size_t)col * LD] = o;
          } else {
            W[col + (size_t)row * LD] = make_double2(o.x, -o.y);
            W0[col + (size_t)row * LD] = make_double2(o.x, -o.y);
          }
        }
      }
    }
}

// One inner round of the register-resident block Jacobi: 4 disjoint column pairs of the 8 columns a
// warp holds (4 rows per lane) are rotated side by side.  CROSS = false: round IR of the 7-round
// tournament over all 28 pairs; CROSS = true: pairs (i, 4 + (i + IR) % 4), i.e. only pairs across
// the two column groups (4 rounds cover the 16 of them).
template <int IR, bool CROSS>
__device__ __forceinline__ void jacobi_pair(int ip, int& pa, int& pb) {
  if (CROSS) {
    pa = ip;
    pb = 4 + (ip + IR) % 4;
  } else {
    const int a_ = (ip == 0) ? 7 : (IR + ip) % 7;
    const int b_ = (ip == 0) ? IR : (IR + 7 - ip) % 7;
    pa = a_ < b_ ? a_ : b_;
    pb = a_ < b_ ? b_ : a_;
  }
}

template <int IR, bool CROSS, int NE>
__device__ __forceinline__ void jacobi_inner(double2 (&x)[8][NE], double (&nrm)[8], int lane, double tol2,
                                             bool& any) {
  // the 4 disjoint pairs of this inner round: partial cross products of all four ...
  double pv[8];
  double al = 0.0, be = 0.0;
#pragma unroll
  for (int ip = 0; ip < 4; ++ip) {
    int pa, pb;
    jacobi_pair<IR, CROSS>(ip, pa, pb);
    double gr = 0.0, gi = 0.0;
#pragma unroll
    for (int e = 0; e < NE; ++e) {
      const double2 u = x[pa][e], v = x[pb][e];
      gr = fma(u.x, v.x, fma(u.y, v.y, gr));   // Re conj(p) q
      gi = fma(u.x, v.y, fma(-u.y, v.x, gi));  // Im conj(p) q
    }
    pv[ip * 2 + 0] = gr, pv[ip * 2 + 1] = gi;
    if ((lane >> 3) == ip) al = nrm[pa], be = nrm[pb];
  }
  // ... reduced over the warp with a transposing butterfly (7 + 2 shuffles): afterwards
  // lane L holds entry (L >> 2) & 7, i.e. the 8 lanes of group ip = L >> 3 hold pair ip
  double red;
  {
    double w4[4], w2[2];
    const bool u16 = lane & 16, u8 = lane & 8, u4 = lane & 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const double send = u16 ? pv[i] : pv[i + 4];
      const double keep = u16 ? pv[i + 4] : pv[i];
      w4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const double send = u8 ? w4[i] : w4[i + 2];
      const double keep = u8 ? w4[i + 2] : w4[i];
      w2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    const double send = u4 ? w2[0] : w2[1];
    const double keep = u4 ? w2[1] : w2[0];
    red = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    red += __shfl_xor_sync(0xffffffffu, red, 2);
    red += __shfl_xor_sync(0xffffffffu, red, 1);
  }
  const int gbase = lane & 0x18;
  const double gr = __shfl_sync(0xffffffffu, red, gbase | 0);
  const double gi = __shfl_sync(0xffffffffu, red, gbase | 4);
  // every 8-lane group computes the rotation of ITS pair (4 pairs in parallel):
  // tan t = 2|g| sign(d) / (|d| + sqrt(d^2 + 4|g|^2)), d = |q|^2 - |p|^2; the phase
  // e^{i phi} = g / |g| only enters as sn e^{i phi} = cs kappa g
  const double g2 = gr * gr + gi * gi;
  const bool doit = (g2 > tol2 * al * be) && g2 > 0.0;
  // with hh = d^2 + 4|g|^2, h = sqrt(hh), q = |d| + h:  1 + kappa^2 |g|^2 = 2h / q, hence with
  // y = q / 2h (in [1/2, 1]):  cs = sqrt(y),  cs kappa = sign(d) / (h sqrt(y)),  kappa = sign(d) / (h y)
  // -- two special functions in all, rsqrt(hh) and rsqrt(y), instead of the sqrt -> divide -> rsqrt
  // chain: the inner rounds are bound by FP64 issue and by this dependent chain
  const double d = be - al;
  const double hh = fma(d, d, 4.0 * g2);
  const double rh = rsqrt(hh);                     // 1 / h
  const double y = fma(0.5 * fabs(d), rh, 0.5);    // (|d| + h) / 2h
  const double ry = rsqrt(y);
  const double sgn_rh = (d >= 0.0) ? rh : -rh;
  const double csl = y * ry;
  const double sf = sgn_rh * ry;
  const double my_dn = doit ? (sgn_rh * g2) * (ry * ry) : 0.0;
  const double my_cs = doit ? csl : 1.0;
  const double my_sr = doit ? sf * gr : 0.0;
  const double my_si = doit ? sf * gi : 0.0;
#pragma unroll
  for (int ip = 0; ip < 4; ++ip) {
    int pa, pb;
    jacobi_pair<IR, CROSS>(ip, pa, pb);
    const double cs = __shfl_sync(0xffffffffu, my_cs, ip << 3);
    const double sr = __shfl_sync(0xffffffffu, my_sr, ip << 3);
    const double si = __shfl_sync(0xffffffffu, my_si, ip << 3);
    const double dn = __shfl_sync(0xffffffffu, my_dn, ip << 3);
    nrm[pa] -= dn;
    nrm[pb] += dn;
    any = any || (sr != 0.0) || (si != 0.0);
    // p' = cs p - sn e^{-i phi} q ;  q' = sn e^{i phi} p + cs q   (identity if not rotated)
    const double2 fm = make_double2(-sr, si), fp = make_double2(sr, si);
#pragma unroll
    for (int e = 0; e < NE; ++e) {
      const double2 u = x[pa][e], v = x[pb][e];
      double2 nu = make_double2(cs * u.x, cs * u.y);
      cfma(nu, fm, v);
      double2 nv = make_double2(cs * v.x, cs * v.y);
      cfma(nv, fp, u);
      x[pa][e] = nu;
      x[pb][e] = nv;
    }
  }
}

// The Jacobi sweeps of one SVD on the matrix Bw (Rj x Cc, leading dimension ldw; global or shared
// memory), shared by `csize` CTAs (crank = this CTA).  NE = rows per lane: 32 NE >= Rj, so a small
// matrix does not pay for the zero rows of the 128-row layout (a warp with an SM sub-partition to
// itself is bound by FP64 issue: 2 cycles per instruction).
template <int NE>
__device__ __forceinline__ void jacobi_sweeps(double2* Bw, int ldw, int Rj, int Cc, bool solo, int crank,
                                              int csize, int* conv, int* sweeps_out, bool fence, double tol_floor) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  auto team_sync = [&]() {
    if (solo)
      __syncthreads();
    else
      cluster.sync();
  };
  const int ng = (Cc + 3) / 4;       // column groups
  const int ne = (ng + 1) & ~1;      // even number of players (one phantom group if ng is odd)
  const int npairs = ne / 2;
  // convergence: |<p, q>| <= tol |p| |q| with tol ~ 2 sqrt(R) eps (LAPACK xGESVJ uses sqrt(m) eps);
  // a tighter value sits below the rounding noise of the inner product and never converges
  // When the split that follows discards weight up to trunc_thr anyway, orthogonality far below that is
  // wasted sweeps: tol_floor (0 for untruncated runs) relaxes the test (section 6 of DESIGN.md).
  const double tol = fmax(2.0 * sqrt((double)Rj) * 2.220446049250313e-16, tol_floor);
  const double tol2 = tol * tol;
  for (int sweep = 0; sweep < 30; ++sweep) {
    int rotated = 0;
    const int nrounds = (ne > 1) ? ne - 1 : 1;
    for (int round = 0; round < nrounds; ++round) {
      for (int pi = warp * csize + crank; pi < npairs; pi += nwarps * csize) {
        int gI, gJ;
        if (ne == 1 || ng == 1) {
          gI = 0;
          gJ = 1;  // single group: the second group is empty
        } else if (pi == 0) {
          gI = ne - 1;
          gJ = round;
        } else {
          gI = (round + pi) % (ne - 1);
          gJ = (round + ne - 1 - pi) % (ne - 1);
        }
        if (gI > gJ) {
          const int tmp = gI;
          gI = gJ;
          gJ = tmp;
        }
        if (gI >= ng) continue;
        int col[8];
#pragma unroll
        for (int c = 0; c < 4; ++c) col[c] = 4 * gI + c, col[4 + c] = 4 * gJ + c;
        double2 x[8][NE];
#pragma unroll
        for (int c = 0; c < 8; ++c)
#pragma unroll
          for (int e = 0; e < NE; ++e) {
            const int r = lane + 32 * e;
            x[c][e] = (col[c] < Cc && r < Rj) ? Bw[r + (size_t)col[c] * ldw] : make_double2(0.0, 0.0);
          }
        // squared column norms of the block: computed once per visit, then tracked through the
        // rotations (|p'|^2 = |p|^2 - t|g|, |q'|^2 = |q|^2 + t|g| with t|g| = kappa |g|^2), so the inner
        // rounds only need the cross products <p|q>
        double nrm[8];
        {
          double pn[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            double acc = 0.0;
#pragma unroll
            for (int e = 0; e < NE; ++e) acc = fma(x[c][e].x, x[c][e].x, fma(x[c][e].y, x[c][e].y, acc));
            pn[c] = acc;
          }
          double w4[4], w2[2];
          const bool u16 = lane & 16, u8 = lane & 8, u4 = lane & 4;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const double send = u16 ? pn[i] : pn[i + 4];
            const double keep = u16 ? pn[i + 4] : pn[i];
            w4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
          }
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const double send = u8 ? w4[i] : w4[i + 2];
            const double keep = u8 ? w4[i + 2] : w4[i];
            w2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
          }
          const double send = u4 ? w2[0] : w2[1];
          const double keep = u4 ? w2[1] : w2[0];
          double red = keep + __shfl_xor_sync(0xffffffffu, send, 4);
          red += __shfl_xor_sync(0xffffffffu, red, 2);
          red += __shfl_xor_sync(0xffffffffu, red, 1);
          // lane L holds column ((L >> 4) & 1) * 4 + ((L >> 3) & 1) * 2 + ((L >> 2) & 1)
#pragma unroll
          for (int c = 0; c < 8; ++c)
            nrm[c] = __shfl_sync(0xffffffffu, red, ((c >> 2) & 1) * 16 + ((c >> 1) & 1) * 8 + (c & 1) * 4);
        }
        // Round 0 of a sweep orthogonalises all 28 pairs of the 8 columns (7 inner rounds); in it every
        // column group meets exactly one partner, so the 6 pairs INSIDE each group are covered once
        // per sweep there.  The other rounds only take the 16 pairs ACROSS the two groups (4 inner
        // rounds): repeating the inside pairs in all ne-1 rounds cost 12 of every 28 rotations.
        bool any = false;
        if (round == 0) {
          jacobi_inner<0, false, NE>(x, nrm, lane, tol2, any);
          jacobi_inner<1, false, NE>(x, nrm, lane, tol2, any);
          jacobi_inner<2, false, NE>(x, nrm, lane, tol2, any);
          jacobi_inner<3, false, NE>(x, nrm, lane, tol2, any);
          jacobi_inner<4, false, NE>(x, nrm, lane, tol2, any);
          jacobi_inner<5, false, NE>(x, nrm, lane, tol2, any);
          jacobi_inner<6, false, NE>(x, nrm, lane, tol2, any);
        } else {
          jacobi_inner<0, true, NE>(x, nrm, lane, tol2, any);
          jacobi_inner<1, true, NE>(x, nrm, lane, tol2, any);
          jacobi_inner<2, true, NE>(x, nrm, lane, tol2, any);
          jacobi_inner<3, true, NE>(x, nrm, lane, tol2, any);
        }
        if (!any) continue;  // warp-uniform
        rotated = 1;
#pragma unroll
        for (int c = 0; c < 8; ++c)
#pragma unroll
          for (int e = 0; e < NE; ++e) {
            const int r = lane + 32 * e;
            if (col[c] < Cc && r < Rj) Bw[r + (size_t)col[c] * ldw] = x[c][e];
          }
        __syncwarp();
      }
      // cluster.sync() is barrier.cluster.arrive.release + wait.acquire: it orders the global-memory
      // columns between the CTAs of the cluster by itself; the device-scope fence in front of it was
      // 9 % of the stall samples (ERRBAR).  AQC_MPS_FENCE=1 puts it back.
      if (fence) __threadfence();
      team_sync();
    }
    if (tid == 0 && sweeps_out && crank == 0) *sweeps_out = sweep + 1;
    const int mine = __syncthreads_or(rotated);
    if (csize == 1) {
      if (mine == 0) break;
    } else {
      if (tid == 0 && mine) atomicAdd(conv + sweep, 1);
      if (fence) __threadfence();
      team_sync();
      if (*(volatile int*)(conv + sweep) == 0) break;
    }
  }
}

// ------------------------------------------------------------------------------------------
// kernel: one-sided Jacobi SVD of the working matrix + truncation + split into Gamma, lambda
// ------------------------------------------------------------------------------------------
struct SvdArgs {
  StateMut st[2];
  const MpsTask* tasks;
  double2* work;
  const double2* work0;  // the working matrix before the rotations
  double2* vmat;         // right singular vectors of the KEPT columns (compacted by rank)
  int C, maxtasks, chi_max;
  double trunc_thr;
  int* sweeps;  // [state][maxtasks] Jacobi sweeps used (diagnostics)
  int* conv;    // [state][maxtasks][32] rotations counted per sweep (cluster-wide convergence)
  int precond;  // 1: Householder QR first, Jacobi on R^H (Drmac-Veselic preconditioning)
  int fence;    // 1: device-scope fence in front of every cluster barrier (AQC_MPS_FENCE=1)
  // truncation record of the current public call: [0] sum of discarded weights (sum of squared Schmidt
  // values relative to the split's total), [1] largest single discard, [2] part of [0] that only the
  // chi_max cap removed (beyond the trunc_thr rule), [3] number of splits the cap cut
  double* trunc_stats;
  // relaxed Jacobi convergence for truncated runs: columns count as orthogonal once |<p, q>| <= tol_floor
  // |p| |q| (1e-3 trunc_thr, at most 1e-9; 0 keeps the rounding-level test of untruncated runs)
  double tol_floor;
};

__global__ void __launch_bounds__(256) mps_svd_kernel(const SvdArgs A) {
  __shared__ double2 s_ga[8 * 2 * kMaxChi], s_gb[8 * kMaxChi];  // k-slices of B0 and of the kept columns of B
  __shared__ double s_sig[2 * kMaxChi];
  __shared__ int s_order[2 * kMaxChi];
  __shared__ int s_keep, s_total;
  __shared__ double s_scale;
  __shared__ double s_beta;
  __shared__ double2 s_alpha;
  // One SVD is shared by the CTAs of a thread-block cluster: each CTA rotates its share of the
  // block pairs of a round (the matrix lives in global memory / L2), rounds are separated by
  // cluster barriers.  Rank 0 finishes (sort, truncate, split).
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int crank_hw = (int)cluster.block_rank(), csize_hw = (int)cluster.num_blocks();
  const int t = blockIdx.x / csize_hw, s = blockIdx.y, C = A.C;
  const MpsTask tk = A.tasks[t];
  const StateMut S = A.st[s];
  const int k = tk.site;
  const int cl = S.dims[k], cr = S.dims[k + 2];
  const int M = 2 * cl, N = 2 * cr, LD = 2 * C;
  const bool transposed = M < N;
  const int R = transposed ? N : M;   // rows of the working matrix
  const int Cc = transposed ? M : N;  // columns (<= rows)
  double2* B = A.work + ((size_t)s * A.maxtasks + t) * (size_t)LD * LD;
  double2* V = A.vmat + ((size_t)s * A.maxtasks + t) * (size_t)LD * LD;
  const double2* B0 = A.work0 + ((size_t)s * A.maxtasks + t) * (size_t)LD * LD;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  // Small matrices (<= 32 columns: at most 4 block pairs per round) are done by ONE CTA with CTA
  // barriers; the other CTAs of the cluster leave before any cluster barrier is used.  Production
  // runs (trunc_thr = 1e-6, physical targets) live in this regime: a cluster barrier per round costs
  // more than the round itself.
  const bool solo = Cc <= 32;
  if (solo && crank_hw != 0) return;
  const int crank = solo ? 0 : crank_hw, csize = solo ? 1 : csize_hw;
  auto team_sync = [&]() {
    if (solo)
      __syncthreads();
    else
      cluster.sync();
  };

  int* conv = A.conv + ((size_t)s * A.maxtasks + t) * 32;
  if (crank == 0 && tid < 32) conv[tid] = 0;

  // Preconditioning (Drmac-Veselic): W = Q R by Householder reflections, then the one-sided Jacobi
  // runs on X = R^H (Cc x Cc, lower triangular), whose columns are far closer to orthogonal than
  // those of W: 6-8 sweeps instead of 10-20 on graded spectra.  X V' = U' Sigma gives the RIGHT
  // singular vectors of W directly (U' = normalised columns of the converged X), the left ones are
  // recovered from W U' Sigma^-1 below; Q is never formed.
  const bool pre = A.precond != 0 && Cc >= 8;
  const int Rj = pre ? Cc : R;  // rows of the matrix the Jacobi sweeps work on
  if (pre && crank == 0) {
    double2* s_v = s_ga;  // Householder vector of the current column
    for (int j = 0; j < Cc; ++j) {
      if (warp == 0) {
        double acc = 0.0;
        for (int r = j + lane; r < R; r += 32) {
          const double2 b = B[r + (size_t)j * LD];
          acc = fma(b.x, b.x, fma(b.y, b.y, acc));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        const double nx = sqrt(acc);
        const double2 x0 = B[j + (size_t)j * LD];
        const double ax0 = hypot(x0.x, x0.y);
        const double2 sph = ax0 > 0.0 ? make_double2(x0.x / ax0, x0.y / ax0) : make_double2(1.0, 0.0);
        // H x = alpha e_0, alpha = -sph |x|;  v = x - alpha e_0;  H = I - beta v v^H, beta = 2 / v^H v
        for (int r = j + lane; r < R; r += 32)
          s_v[r - j] = (r == j) ? make_double2(sph.x * (ax0 + nx), sph.y * (ax0 + nx)) : B[r + (size_t)j * LD];
        if (lane == 0) {
          s_beta = nx > 0.0 ? 1.0 / (nx * (nx + ax0)) : 0.0;
          s_alpha = make_double2(-sph.x * nx, -sph.y * nx);
        }
      }
      __syncthreads();
      const double beta = s_beta;
      if (beta != 0.0) {
        // four columns per warp in flight: the update is a chain of L2 round trips otherwise
        double2 hv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int r = j + lane + 32 * e;
          hv[e] = (r < R) ? s_v[r - j] : make_double2(0.0, 0.0);
        }
        for (int cb = j + 1 + warp; cb < Cc; cb += 4 * nwarps) {
          double2 a[4][4], dot[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int c = cb + u * nwarps;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int r = j + lane + 32 * e;
              a[u][e] = (c < Cc && r < R) ? B[r + (size_t)c * LD] : make_double2(0.0, 0.0);
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            dot[u] = make_double2(0.0, 0.0);
#pragma unroll
            for (int e = 0; e < 4; ++e) cfma_conj(dot[u], hv[e], a[u][e]);
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              dot[u].x += __shfl_xor_sync(0xffffffffu, dot[u].x, o);
              dot[u].y += __shfl_xor_sync(0xffffffffu, dot[u].y, o);
            }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int c = cb + u * nwarps;
            const double2 f = make_double2(-beta * dot[u].x, -beta * dot[u].y);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int r = j + lane + 32 * e;
              if (c < Cc && r < R) {
                cfma(a[u][e], hv[e], f);
                B[r + (size_t)c * LD] = a[u][e];
              }
            }
          }
        }
      }
      if (tid == 0) B[j + (size_t)j * LD] = s_alpha;
      __syncthreads();
    }
    // X = R^H in place (the strict lower triangle held reflector leftovers)
    for (int e = tid; e < Cc * Cc; e += blockDim.x) {
      const int i = e % Cc, jc = e / Cc;
      if (i > jc) {
        const double2 r = B[jc + (size_t)i * LD];
        B[i + (size_t)jc * LD] = make_double2(r.x, -r.y);
        B[jc + (size_t)i * LD] = make_double2(0.0, 0.0);
      } else if (i == jc) {
        const double2 r = B[i + (size_t)i * LD];
        B[i + (size_t)i * LD] = make_double2(r.x, -r.y);
      }
    }
    __syncthreads();
  }
  if (A.fence) __threadfence();
  team_sync();

  // Block one-sided Jacobi.  Columns are grouped in fours; a warp takes a PAIR of groups (8 columns,
  // 4 rows per lane => 32 complex numbers in registers) and orthogonalises all 28 column pairs of
  // the block in registers (7 inner rounds of 4 independent rotations), so every column is loaded
  // and stored once per 7 rotations instead of once per rotation: the plain cyclic scheme is bound
  // by per-SM L2 bandwidth (16 KiB moved per rotation).  V is not accumulated: the other set of
  // singular vectors is recovered from the untouched working matrix after convergence (below).
  // solo mode: the working matrix (Rj x Cc <= 1024 elements) moves into shared memory for the sweeps,
  // so a block visit costs shared-memory instead of L2 latency (the sweeps of small matrices are a
  // chain of dependent rounds, not throughput)
  double2* Bw = B;
  int ldw = LD;
  if (solo && Rj * Cc <= 8 * 2 * kMaxChi) {
    Bw = s_ga;
    ldw = Rj;
    for (int e = tid; e < Rj * Cc; e += blockDim.x) Bw[e] = B[(e % Rj) + (size_t)(e / Rj) * LD];
    __syncthreads();
  }
  int* sweeps_out = A.sweeps ? A.sweeps + s * A.maxtasks + t : nullptr;
  if (Rj <= 32)
    jacobi_sweeps<1>(Bw, ldw, Rj, Cc, solo, crank, csize, conv, sweeps_out, A.fence != 0, A.tol_floor);
  else if (Rj <= 64)
    jacobi_sweeps<2>(Bw, ldw, Rj, Cc, solo, crank, csize, conv, sweeps_out, A.fence != 0, A.tol_floor);
  else
    jacobi_sweeps<4>(Bw, ldw, Rj, Cc, solo, crank, csize, conv, sweeps_out, A.fence != 0, A.tol_floor);
  if (crank != 0) return;
  if (Bw != B) {  // back to global memory for the split below
    __syncthreads();
    for (int e = tid; e < Rj * Cc; e += blockDim.x) B[(e % Rj) + (size_t)(e / Rj) * LD] = Bw[e];
    __syncthreads();
  }

  // singular values = column norms
  for (int c = warp; c < Cc; c += nwarps) {
    double acc = 0.0;
    const double2* bc = B + (size_t)c * LD;
    for (int r = lane; r < Rj; r += 32) acc = fma(bc[r].x, bc[r].x, fma(bc[r].y, bc[r].y, acc));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) s_sig[c] = sqrt(acc);
  }
  __syncthreads();
  if (tid < Cc) {  // descending rank
    const double me = s_sig[tid];
    int rank = 0;
    for (int i = 0; i < Cc; ++i) {
      const double o = s_sig[i];
      rank += (o > me || (o == me && i < tid)) ? 1 : 0;
    }
    s_order[rank] = tid;
  }
  __syncthreads();
  if (tid == 0) {
    int total = 0;
    for (int i = 0; i < Cc; ++i) total += (s_sig[s_order[i]] > kChop) ? 1 : 0;
    // the reference's rule (qiskit-aer run WITHOUT a bond cap, mps_operations.py:248-265): drop the smallest
    // Schmidt values while the sum of their squares stays below trunc_thr ...
    int keep = total < 1 ? 1 : total;
    double dropped = 0.0;
    while (keep > 1) {
      const double v = s_sig[s_order[keep - 1]];
      if (dropped + v * v < A.trunc_thr) {
        dropped += v * v;
        --keep;
      } else {
        break;
      }
    }
    // ... and only then this engine's bond capacity: the cap cuts a split only if the rule alone would have
    // kept more than chi_max values
    const int keep_rule = keep;
    if (keep > A.chi_max) keep = A.chi_max;
    double scale = 1.0;
    if (keep < total) {
      double nrm = 0.0;
      for (int i = 0; i < keep; ++i) nrm += s_sig[s_order[i]] * s_sig[s_order[i]];
      scale = 1.0 / sqrt(nrm);
      double cut = 0.0, capcut = 0.0;
      for (int i = keep; i < total; ++i) cut += s_sig[s_order[i]] * s_sig[s_order[i]];
      for (int i = keep; i < keep_rule; ++i) capcut += s_sig[s_order[i]] * s_sig[s_order[i]];
      const double all = nrm + cut;
      if (A.trunc_stats && all > 0.0) {
        atomicAdd(A.trunc_stats + 0, cut / all);
        // (non-negative doubles order like their bit patterns)
        atomicMax(reinterpret_cast<unsigned long long*>(A.trunc_stats + 1),
                  (unsigned long long)__double_as_longlong(cut / all));
        if (keep_rule > keep) {
          atomicAdd(A.trunc_stats + 2, capcut / all);
          atomicAdd(A.trunc_stats + 3, 1.0);
        }
      }
    }
    s_keep = keep;
    s_total = total;
    s_scale = scale;
  }
  __syncthreads();
  const int keep = s_keep;
  const double scale = s_scale;
  // The other set of singular vectors for the KEPT columns, compacted by rank, from one product with
  // the untouched working matrix B0 (instead of replaying every rotation on an accumulated V):
  //   plain:          B = U Sigma  ->  V[:, r] = B0^H B[:, j_r] / sigma^2   (Cc x R) . (R x keep)
  //   preconditioned: B = V Sigma  ->  U[:, r] = B0   B[:, j_r] / sigma^2   (R x Cc) . (Cc x keep)
  {
    const int tr = tid & 31, tc = tid >> 5;  // rows 4 tr .. 4 tr + 3, columns 8 tc .. 8 tc + 7
    const int orows = pre ? R : Cc, kdim = pre ? Cc : R;
    double2 acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[i][c] = make_double2(0.0, 0.0);
    for (int k0 = 0; k0 < kdim; k0 += 8) {
      for (int e = tid; e < 8 * 2 * kMaxChi; e += 256) {
        double2 v = make_double2(0.0, 0.0);
        if (!pre) {
          const int kk = e & 7, r = e >> 3;  // consecutive threads walk along k (contiguous)
          if (r < orows && k0 + kk < kdim) v = B0[(k0 + kk) + (size_t)r * LD];
          s_ga[kk * 2 * kMaxChi + r] = make_double2(v.x, -v.y);
        } else {
          const int r = e & (2 * kMaxChi - 1), kk = e >> 7;  // ... along the rows
          if (r < orows && k0 + kk < kdim) v = B0[r + (size_t)(k0 + kk) * LD];
          s_ga[kk * 2 * kMaxChi + r] = v;
        }
      }
      for (int e = tid; e < 8 * kMaxChi; e += 256) {
        const int kk = e & 7, jj = e >> 3;
        s_gb[kk * kMaxChi + jj] =
            (jj < keep && k0 + kk < kdim) ? B[(k0 + kk) + (size_t)s_order[jj] * LD] : make_double2(0.0, 0.0);
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        double2 a[4], b[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = s_ga[kk * 2 * kMaxChi + 4 * tr + i];
#pragma unroll
        for (int c = 0; c < 8; ++c) b[c] = s_gb[kk * kMaxChi + 8 * tc + c];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int c = 0; c < 8; ++c) cfma(acc[i][c], a[i], b[c]);
      }
      __syncthreads();
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int jj = 8 * tc + c;
      if (jj >= keep) continue;
      const double sg = s_sig[s_order[jj]];
      const double inv = sg > 0.0 ? 1.0 / (sg * sg) : 0.0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = 4 * tr + i;
        if (r < orows) V[r + (size_t)jj * LD] = make_double2(acc[i][c].x * inv, acc[i][c].y * inv);
      }
    }
  }
  __syncthreads();
  // write lambda (bond k+1), dims, Gamma_k, Gamma_{k+1}
  double* lamM = S.lam + (size_t)(k + 1) * C;
  const double* lamL = S.lam + (size_t)k * C;
  const double* lamR = S.lam + (size_t)(k + 2) * C;
  for (int r = tid; r < C; r += blockDim.x) lamM[r] = (r < keep) ? s_sig[s_order[r]] * scale : 0.0;
  if (tid == 0) S.dims[k + 1] = keep;
  double2* Ga = S.gam + (size_t)k * 2 * C * C;
  double2* Gb = S.gam + (size_t)(k + 1) * 2 * C * C;
  // Gamma_k[b1][alpha, r] = Uleft[(b1 cl + alpha), j_r] / lamL[alpha]
  for (int i = tid; i < M * keep; i += blockDim.x) {
    const int row = i % M, r = i / M;
    const int j = s_order[r];
    const double sg = s_sig[j];
    double2 u;
    if (transposed == pre) {  // the vectors held by the converged B (scaled by sigma)
      const double2 b = B[row + (size_t)j * LD];
      const double inv = sg > 0.0 ? 1.0 / sg : 0.0;
      u = make_double2(b.x * inv, b.y * inv);
    } else {  // the recovered set
      u = V[row + (size_t)r * LD];
    }
    const int b1 = row / cl, al = row % cl;
    const double il = 1.0 / lamL[al];
    Ga[((size_t)b1 * C + al) * C + r] = make_double2(u.x * il, u.y * il);
  }
  // Gamma_{k+1}[b2][r, gamma] = conj(Vright[(b2 cr + gamma), j_r]) / lamR[gamma]
  for (int i = tid; i < N * keep; i += blockDim.x) {
    const int col = i % N, r = i / N;
    const int j = s_order[r];
    const double sg = s_sig[j];
    double2 v;
    if (transposed == pre) {
      v = V[col + (size_t)r * LD];
    } else {
      const double2 b = B[col + (size_t)j * LD];
      const double inv = sg > 0.0 ? 1.0 / sg : 0.0;
      v = make_double2(b.x * inv, b.y * inv);
    }
    const int b2 = col / cr, ga = col % cr;
    const double ir = 1.0 / lamR[ga];
    Gb[((size_t)b2 * C + r) * C + ga] = make_double2(v.x * ir, -v.y * ir);
  }
}

// ------------------------------------------------------------------------------------------
// kernel: single-site transfer step of an environment (left-to-right or right-to-left)
// ------------------------------------------------------------------------------------------
//   L: E'[i', j'] = sum_b sum_{i,j} conj(Aw[b][i, i']) E[i, j] Az[b][j, j']   (i = left bond)
//   R: E'[i', j'] = sum_b sum_{i,j} conj(Aw[b][i', i]) E[i, j] Az[b][j', j]   (i = right bond)
// with A[b] = Gamma[b] diag(lambda_right).   grid (ntasks), 512 threads, 64 KiB dynamic smem (X).
struct EnvArgs {
  StateView w, z;
  const EnvTask* tasks;
  double2* envL;
  double2* envR;
  int C;
};

template <bool RIGHT>
__device__ __forceinline__ double2 site_elem(const StateView& S, int k, int b, int i, int ip, int C) {
  // element A[b](i -> i'): L: row i (left bond), col i' (right bond);  R: row i' (left), col i (right)
  const int row = RIGHT ? ip : i, col = RIGHT ? i : ip;
  const double2 v = S.gam[(((size_t)k * 2 + b) * C + row) * C + col];
  const double l = S.lam[(size_t)(k + 1) * C + col];
  return make_double2(v.x * l, v.y * l);
}

template <bool RIGHT>
__device__ void env_step(const EnvArgs& A, int k, double2* X) {
  const int C = A.C;
  const int wl = A.w.dims[k], wr = A.w.dims[k + 1], zl = A.z.dims[k], zr = A.z.dims[k + 1];
  const int wi = RIGHT ? wr : wl, wo = RIGHT ? wl : wr;  // contracted / new bond of w
  const int zi = RIGHT ? zr : zl, zo = RIGHT ? zl : zr;
  const double2* E = RIGHT ? A.envR + (size_t)(k + 1) * C * C : A.envL + (size_t)k * C * C;
  double2* Eo = RIGHT ? A.envR + (size_t)k * C * C : A.envL + (size_t)(k + 1) * C * C;
  const int g = threadIdx.x & 63, grp = threadIdx.x >> 6;
  double2 out[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) out[r] = make_double2(0.0, 0.0);
  for (int b = 0; b < 2; ++b) {
    // X[i, j'] = sum_j E[i, j] Az[b](j -> j')
    double2 acc[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) acc[r] = make_double2(0.0, 0.0);
    if (g < zo) {
      for (int j = 0; j < zi; ++j) {
        const double2 az = site_elem<RIGHT>(A.z, k, b, j, g, C);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const int i = grp * 8 + r;
          if (i < wi) cfma(acc[r], E[(size_t)i * C + j], az);
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 8; ++r) X[(size_t)(grp * 8 + r) * 64 + g] = acc[r];
    __syncthreads();
    // E'[i', j'] += sum_i conj(Aw[b](i -> i')) X[i, j']
    if (g < zo) {
      for (int i = 0; i < wi; ++i) {
        const double2 x = X[(size_t)i * 64 + g];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const int ip = grp * 8 + r;
          if (ip < wo) cfma_conj(out[r], site_elem<RIGHT>(A.w, k, b, i, ip, C), x);
        }
      }
    }
  }
  if (g < zo) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int ip = grp * 8 + r;
      if (ip < wo) Eo[(size_t)ip * C + g] = out[r];
    }
  }
}

// The same transfer step as two chained tensor-core GEMMs per physical index b (bond capacity >= 32):
//   X = E . Az[b]  (wi x zi . zi x zo),   E' += conj(Aw[b])^T . X  (wo x wi . wi x zo)
// real form on mma.sync.m8n8k4.f64; X stays in shared memory as split re/im planes and is the B
// operand of the second product.  grid (ntasks), 256 threads, warp w owns output rows 8w .. 8w + 7.
constexpr int kEnvSmemDoubles = 2 * 64 * kThLdA + 2 * 16 * kThLdB + 2 * 64 * kThLdB;

template <bool RIGHT>
__device__ void env_step_dmma(const EnvArgs& A, int k, double* dsm) {
  double* sAr = dsm;
  double* sAi = sAr + 64 * kThLdA;
  double* sBr = sAi + 64 * kThLdA;
  double* sBi = sBr + 16 * kThLdB;
  double* Xr = sBi + 16 * kThLdB;
  double* Xi = Xr + 64 * kThLdB;
  const int C = A.C;
  const int wl = A.w.dims[k], wr = A.w.dims[k + 1], zl = A.z.dims[k], zr = A.z.dims[k + 1];
  const int wi = RIGHT ? wr : wl, wo = RIGHT ? wl : wr;  // contracted / new bond of w
  const int zi = RIGHT ? zr : zl, zo = RIGHT ? zl : zr;
  const double2* E = RIGHT ? A.envR + (size_t)(k + 1) * C * C : A.envL + (size_t)k * C * C;
  double2* Eo = RIGHT ? A.envR + (size_t)k * C * C : A.envL + (size_t)(k + 1) * C * C;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double ore[8][2], oim[8][2];
#pragma unroll
  for (int q = 0; q < 8; ++q) ore[q][0] = ore[q][1] = oim[q][0] = oim[q][1] = 0.0;

  for (int b = 0; b < 2; ++b) {
    double xre[8][2], xim[8][2];
#pragma unroll
    for (int q = 0; q < 8; ++q) xre[q][0] = xre[q][1] = xim[q][0] = xim[q][1] = 0.0;
    for (int j0 = 0; j0 < zi; j0 += 16) {
      for (int e = tid; e < 64 * 16; e += 256) {
        const int kk = e & 15, i = e >> 4;
        double2 v = make_double2(0.0, 0.0);
        if (i < wi && j0 + kk < zi) v = E[(size_t)i * C + j0 + kk];
        sAr[i * kThLdA + kk] = v.x;
        sAi[i * kThLdA + kk] = v.y;
      }
      for (int e = tid; e < 16 * 64; e += 256) {
        const int j = e & 63, kk = e >> 6;
        double2 v = make_double2(0.0, 0.0);
        if (j < zo && j0 + kk < zi) v = site_elem<RIGHT>(A.z, k, b, j0 + kk, j, C);
        sBr[kk * kThLdB + j] = v.x;
        sBi[kk * kThLdB + j] = v.y;
      }
      __syncthreads();
#pragma unroll
      for (int ks = 0; ks < 16; ks += 4) {
        const int ai = (8 * warp + (lane >> 2)) * kThLdA + ks + (lane & 3);
        const double ar = sAr[ai], aim = sAi[ai], nai = -aim;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int bi = (ks + (lane & 3)) * kThLdB + 8 * q + (lane >> 2);
          const double br = sBr[bi], bim = sBi[bi];
          mps_dmma884(xre[q][0], xre[q][1], ar, br);
          mps_dmma884(xre[q][0], xre[q][1], nai, bim);
          mps_dmma884(xim[q][0], xim[q][1], ar, bim);
          mps_dmma884(xim[q][0], xim[q][1], aim, br);
        }
      }
      __syncthreads();
    }
    // X -> shared memory planes (row = 8 warp + lane / 4, columns 8 q + 2 (lane & 3) + h)
#pragma unroll
    for (int q = 0; q < 8; ++q)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int idx = (8 * warp + (lane >> 2)) * kThLdB + 8 * q + 2 * (lane & 3) + h;
        Xr[idx] = xre[q][h];
        Xi[idx] = xim[q][h];
      }
    __syncthreads();
    for (int i0 = 0; i0 < wi; i0 += 16) {
      // A planes: row i' (new bond of w), depth i: conj(Aw[b](i -> i'))
      for (int e = tid; e < 64 * 16; e += 256) {
        const int kk = e & 15, ip = e >> 4;
        double2 v = make_double2(0.0, 0.0);
        if (ip < wo && i0 + kk < wi) v = site_elem<RIGHT>(A.w, k, b, i0 + kk, ip, C);
        sAr[ip * kThLdA + kk] = v.x;
        sAi[ip * kThLdA + kk] = -v.y;
      }
      __syncthreads();
#pragma unroll
      for (int ks = 0; ks < 16; ks += 4) {
        const int ai = (8 * warp + (lane >> 2)) * kThLdA + ks + (lane & 3);
        const double ar = sAr[ai], aim = sAi[ai], nai = -aim;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int bi = (i0 + ks + (lane & 3)) * kThLdB + 8 * q + (lane >> 2);
          const double br = Xr[bi], bim = Xi[bi];
          mps_dmma884(ore[q][0], ore[q][1], ar, br);
          mps_dmma884(ore[q][0], ore[q][1], nai, bim);
          mps_dmma884(oim[q][0], oim[q][1], ar, bim);
          mps_dmma884(oim[q][0], oim[q][1], aim, br);
        }
      }
      __syncthreads();
    }
  }
  const int ip = 8 * warp + (lane >> 2);
  if (ip < wo) {
#pragma unroll
    for (int q = 0; q < 8; ++q)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int j = 8 * q + 2 * (lane & 3) + h;
        if (j < zo) Eo[(size_t)ip * C + j] = make_double2(ore[q][h], oim[q][h]);
      }
  }
}

__global__ void __launch_bounds__(256) mps_env_dmma_kernel(const EnvArgs A) {
  extern __shared__ double env_dsm[];
  const EnvTask tk = A.tasks[blockIdx.x];
  if (tk.dir == 0)
    env_step_dmma<false>(A, tk.site, env_dsm);
  else
    env_step_dmma<true>(A, tk.site, env_dsm);
}

__global__ void __launch_bounds__(512) mps_env_kernel(const EnvArgs A) {
  extern __shared__ double2 smem_x[];
  const EnvTask tk = A.tasks[blockIdx.x];
  if (tk.dir == 0)
    env_step<false>(A, tk.site, smem_x);
  else
    env_step<true>(A, tk.site, smem_x);
}

__global__ void mps_env_init_kernel(double2* envL, double2* envR, int n, int C) {
  // bond 0 and bond n have dimension 1: E = [[1]]
  if (threadIdx.x == 0) {
    envL[0] = make_double2(1.0, 0.0);
    envR[(size_t)n * C * C] = make_double2(1.0, 0.0);
  }
}

// ------------------------------------------------------------------------------------------
// kernel: reduced overlap matrix rho[b'][b] = sum conj(Tw[b'][a, c]) EL[a, a'] Tz[b][a', c'] ER[c, c']
// ------------------------------------------------------------------------------------------
// grid (ntasks, nphys), 512 threads, 128 KiB dynamic smem (X and ER^T)
struct RhoArgs {
  StateView w, z;
  const MpsTask* tasks;
  const double2* theta0;  // [2][maxtasks][4][C][C]  (state 0 = w, 1 = z)
  const double2* envL;
  const double2* envR;
  double2* rho;  // [task][16]
  int C, maxtasks, nphys, span;  // span = sites covered by a task (2 pairs, 1 front)
};

__global__ void __launch_bounds__(512) mps_rho_kernel(const RhoArgs A) {
  extern __shared__ double2 smem_r[];
  __shared__ double s_red[16][2 * 4];
  double2* X = smem_r;             // [64][64]
  double2* ERt = smem_r + 64 * 64;  // [c'][c]
  const int t = blockIdx.x, b = blockIdx.y, C = A.C;
  const MpsTask tk = A.tasks[t];
  const int k = tk.site;
  const int wl = A.w.dims[k], wr = A.w.dims[k + A.span];
  const int zl = A.z.dims[k], zr = A.z.dims[k + A.span];
  const double2* EL = A.envL + (size_t)k * C * C;
  const double2* ER = A.envR + (size_t)(k + A.span) * C * C;
  const double2* Tz = A.theta0 + (((size_t)1 * A.maxtasks + t) * 4 + b) * C * C;
  const double2* Tw = A.theta0 + (((size_t)0 * A.maxtasks + t) * 4) * C * C;
  const int g = threadIdx.x & 63, grp = threadIdx.x >> 6;
  for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
    const int c = i / 64, cp = i % 64;  // ER[c][cp] -> ERt[cp][c]
    ERt[(size_t)cp * 64 + c] = (c < wr && cp < zr) ? ER[(size_t)c * C + cp] : make_double2(0.0, 0.0);
  }
  // X[a, c'] = sum_a' EL[a, a'] Tz[b][a', c']
  double2 acc[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) acc[r] = make_double2(0.0, 0.0);
  if (g < zr) {
    for (int ap = 0; ap < zl; ++ap) {
      const double2 tz = Tz[(size_t)ap * C + g];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int a = grp * 8 + r;
        if (a < wl) cfma(acc[r], EL[(size_t)a * C + ap], tz);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) X[(size_t)(grp * 8 + r) * 64 + g] = acc[r];
  __syncthreads();
  // Mb[a, c] = sum_c' X[a, c'] ER[c, c'] ; rho[b', b] += conj(Tw[b'][a, c]) Mb[a, c]
  double2 part[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) part[q] = make_double2(0.0, 0.0);
  if (g < wr) {
#pragma unroll
    for (int r = 0; r < 8; ++r) acc[r] = make_double2(0.0, 0.0);
    for (int cp = 0; cp < zr; ++cp) {
      const double2 er = ERt[(size_t)cp * 64 + g];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int a = grp * 8 + r;
        if (a < wl) cfma(acc[r], X[(size_t)a * 64 + cp], er);
      }
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int a = grp * 8 + r;
      if (a < wl) {
        for (int bp = 0; bp < A.nphys; ++bp)
          cfma_conj(part[bp], Tw[((size_t)bp * C + a) * C + g], acc[r]);
      }
    }
  }
  // block reduction of 4 complex numbers
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      part[q].x += __shfl_xor_sync(0xffffffffu, part[q].x, o);
      part[q].y += __shfl_xor_sync(0xffffffffu, part[q].y, o);
    }
    if (lane == 0) {
      s_red[warp][2 * q] = part[q].x;
      s_red[warp][2 * q + 1] = part[q].y;
    }
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    double v = 0.0;
    for (int w = 0; w < 16; ++w) v += s_red[w][threadIdx.x];
    const int bp = threadIdx.x >> 1;
    if (bp < A.nphys) {
      // rho layout [task][b' * 4 + b]; single-site tasks use the lo bit only (b', b in {0, 1})
      double* out = (double*)(A.rho + (size_t)t * 16 + bp * 4 + b);
      out[threadIdx.x & 1] = v;
    }
  }
}

__global__ void mps_zero_rho_kernel(double2* rho, int count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) rho[i] = make_double2(0.0, 0.0);
}

// ------------------------------------------------------------------------------------------
// kernel: apply the 2x2 gate of single-site tasks (front layer) to Gamma
// ------------------------------------------------------------------------------------------
__global__ void mps_apply1q_kernel(StateMut S, const MpsTask* __restrict__ tasks,
                                   const double2* __restrict__ gate, int C) {
  const MpsTask tk = tasks[blockIdx.x];
  const int k = tk.site;
  const int cl = S.dims[k], cr = S.dims[k + 1];
  // 2x2 block of the 4x4 gate acting on the lo bit: rows/cols 0, 1
  const double2 g00 = gate[(size_t)blockIdx.x * 16 + 0], g01 = gate[(size_t)blockIdx.x * 16 + 1];
  const double2 g10 = gate[(size_t)blockIdx.x * 16 + 4], g11 = gate[(size_t)blockIdx.x * 16 + 5];
  double2* G0 = S.gam + ((size_t)k * 2 + 0) * C * C;
  double2* G1 = S.gam + ((size_t)k * 2 + 1) * C * C;
  for (int i = threadIdx.x; i < cl * cr; i += blockDim.x) {
    const int a = i / cr, c = i % cr;
    const double2 x0 = G0[(size_t)a * C + c], x1 = G1[(size_t)a * C + c];
    double2 y0 = cmul2(g00, x0), y1 = cmul2(g10, x0);
    cfma(y0, g01, x1);
    cfma(y1, g11, x1);
    G0[(size_t)a * C + c] = y0;
    G1[(size_t)a * C + c] = y1;
  }
}

// ------------------------------------------------------------------------------------------
// kernel: amplitudes of basis states  out[i] = <idx_i | state>  (one CTA per index)
// ------------------------------------------------------------------------------------------
__global__ void mps_amplitude_kernel(StateView S, const long long* __restrict__ idx, int n, int C,
                                     double2* __restrict__ out) {
  __shared__ double2 v[2][kMaxChi];
  const long long index = idx[blockIdx.x];
  const int tid = threadIdx.x;
  if (tid < kMaxChi) v[0][tid] = make_double2(tid == 0 ? 1.0 : 0.0, 0.0);
  __syncthreads();
  int cur = 0;
  for (int k = 0; k < n; ++k) {
    const int b = (int)((index >> k) & 1);
    const int cl = S.dims[k], cr = S.dims[k + 1];
    const double2* G = S.gam + ((size_t)k * 2 + b) * C * C;
    if (tid < cr) {
      double2 acc = make_double2(0.0, 0.0);
      for (int a = 0; a < cl; ++a) cfma(acc, v[cur][a], G[(size_t)a * C + tid]);
      const double l = S.lam[(size_t)(k + 1) * C + tid];
      v[cur ^ 1][tid] = make_double2(acc.x * l, acc.y * l);
    }
    __syncthreads();
    cur ^= 1;
  }
  if (tid == 0) out[blockIdx.x] = v[cur][0];
}

// |index> as a bond-1 MPS; site `site` (if >= 0) holds a0 |0> + a1 |1> instead of a basis vector
__global__ void mps_set_product_kernel(StateMut S, long long index, int n, int C, int site, double2 a0,
                                       double2 a1) {
  const int k = blockIdx.x;
  double2* G = S.gam + (size_t)k * 2 * C * C;
  for (int i = threadIdx.x; i < 2 * C * C; i += blockDim.x) G[i] = make_double2(0.0, 0.0);
  __syncthreads();
  if (threadIdx.x == 0) {
    const int b = (int)((index >> k) & 1);
    if (k == site) {
      G[0] = a0;
      G[(size_t)C * C] = a1;
    } else {
      G[(size_t)b * C * C] = make_double2(1.0, 0.0);
    }
    S.dims[k] = 1;
    if (k == n - 1) S.dims[n] = 1;
  }
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    S.lam[(size_t)k * C + i] = (i == 0) ? 1.0 : 0.0;
    if (k == n - 1) S.lam[(size_t)n * C + i] = (i == 0) ? 1.0 : 0.0;
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static size_t gam_elems(const aqc_mps* m) { return (size_t)m->n * 2 * m->C * m->C; }
static size_t lam_elems(const aqc_mps* m) { return (size_t)(m->n + 1) * m->C; }

static StateView view(const aqc_mps* m, int slot) {
  return {m->st[slot].gam, m->st[slot].lam, m->st[slot].dims};
}
static StateMut mut(const aqc_mps* m, int slot) {
  return {m->st[slot].gam, m->st[slot].lam, m->st[slot].dims};
}

static int mps_check_slot(const aqc_mps* m, int slot) {
  if (!m) return aqc_fail(AQC_EINVAL, "null MPS workspace");
  if (slot < 0 || slot >= m->nslots) return aqc_fail(AQC_EINVAL, "MPS slot %d out of range", slot);
  return AQC_OK;
}

static int mps_ensure_pinned(aqc_mps* m, size_t doubles) {
  if (doubles <= m->pinned_cap) return AQC_OK;
  if (m->h_pinned) cudaFreeHost(m->h_pinned);
  m->h_pinned = nullptr;
  m->pinned_cap = 0;
  MCU(cudaMallocHost(&m->h_pinned, doubles * sizeof(double)));
  m->pinned_cap = doubles;
  return AQC_OK;
}

extern "C" void aqc_mps_destroy(aqc_mps* m) {
  if (!m) return;
  cudaSetDevice(m->device);
  for (auto& s : m->st) {
    if (s.gam) cudaFree(s.gam);
    if (s.lam) cudaFree(s.lam);
    if (s.dims) cudaFree(s.dims);
  }
  for (void* p : {(void*)m->d_thetas, (void*)m->d_gate, (void*)m->d_theta0, (void*)m->d_work,
                  (void*)m->d_vmat, (void*)m->d_work0, (void*)m->d_envL, (void*)m->d_envR, (void*)m->d_rho,
                  (void*)m->d_gacc, (void*)m->d_small, (void*)m->d_idx, (void*)m->fwd.d_tasks,
                  (void*)m->dag.d_tasks, (void*)m->fwd.d_env, (void*)m->dag.d_env, (void*)m->d_sweeps, (void*)m->d_conv,
                  (void*)m->d_trunc})
    if (p) cudaFree(p);
  if (m->h_pinned) cudaFreeHost(m->h_pinned);
  if (m->ev0) cudaEventDestroy(m->ev0);
  if (m->ev1) cudaEventDestroy(m->ev1);
  if (m->stream) cudaStreamDestroy(m->stream);
  delete m;
}

extern "C" int aqc_mps_create(const aqc_circuit* circ, int device, int chi_max, double trunc_thr,
                              int num_slots, aqc_mps** out) {
  if (!out) return aqc_fail(AQC_EINVAL, "out is null");
  *out = nullptr;
  if (!circ) return aqc_fail(AQC_EINVAL, "circuit is null");
  if (chi_max < 1 || chi_max > kMaxChi)
    return aqc_fail(AQC_EINVAL, "chi_max must be in [1, %d]", kMaxChi);
  if (!(trunc_thr >= 0.0 && trunc_thr <= 0.1)) return aqc_fail(AQC_EINVAL, "bad trunc_thr");
  if (num_slots < 1 || num_slots > 64) return aqc_fail(AQC_EINVAL, "num_slots must be in [1, 64]");
  const int ndev = aqc_device_count();
  if (ndev <= 0) return aqc_fail(AQC_ENODEV, "no CUDA device visible: this library has no CPU path");
  if (device < 0 || device >= ndev) return aqc_fail(AQC_EINVAL, "device %d out of range", device);
  aqc_mps* m = new aqc_mps();
  m->device = device;
  m->n = aqc_circ_n(circ);
  m->C = kMaxChi;
  m->ent = aqc_circ_ent(circ);
  m->tpb = aqc_circ_tpb(circ);
  m->nthetas = aqc_circ_nthetas(circ);
  m->chi_max = chi_max;
  m->trunc_thr = trunc_thr;
  m->nslots = num_slots;
  std::string err;
  if (build_mps_program(circ, false, m->fwd, err) || build_mps_program(circ, true, m->dag, err)) {
    delete m;
    return aqc_fail(AQC_EINVAL, "%s", err.c_str());
  }
  m->maxtasks = m->n;
  for (auto& s : m->fwd.steps) m->maxtasks = std::max(m->maxtasks, s.ntasks);
  cudaError_t e = cudaSetDevice(device);
  auto alloc = [&](void** p, size_t bytes) {
    if (e == cudaSuccess) e = cudaMalloc(p, bytes);
    if (e == cudaSuccess) e = cudaMemsetAsync(*p, 0, bytes, 0);
  };
  m->st.resize(num_slots);
  for (auto& s : m->st) {
    alloc((void**)&s.gam, gam_elems(m) * sizeof(double2));
    alloc((void**)&s.lam, lam_elems(m) * sizeof(double));
    alloc((void**)&s.dims, (m->n + 1) * sizeof(int));
  }
  const size_t C = m->C, mt = m->maxtasks;
  alloc((void**)&m->d_thetas, (size_t)m->nthetas * sizeof(double));
  alloc((void**)&m->d_gate, mt * 16 * sizeof(double2));
  alloc((void**)&m->d_theta0, 2 * mt * 4 * C * C * sizeof(double2));
  alloc((void**)&m->d_work, 2 * mt * 4 * C * C * sizeof(double2));
  alloc((void**)&m->d_vmat, 2 * mt * 4 * C * C * sizeof(double2));
  alloc((void**)&m->d_work0, 2 * mt * 4 * C * C * sizeof(double2));
  alloc((void**)&m->d_envL, (size_t)(m->n + 1) * C * C * sizeof(double2));
  alloc((void**)&m->d_envR, (size_t)(m->n + 1) * C * C * sizeof(double2));
  alloc((void**)&m->d_rho, mt * 16 * sizeof(double2));
  alloc((void**)&m->d_gacc, (size_t)m->nthetas * 2 * sizeof(double));
  alloc((void**)&m->d_small, 4096 * sizeof(double2));
  alloc((void**)&m->d_idx, 4096 * sizeof(long long));
  alloc((void**)&m->d_sweeps, 2 * mt * sizeof(int));
  alloc((void**)&m->d_trunc, 4 * sizeof(double));
  if (e == cudaSuccess) e = cudaMemset(m->d_trunc, 0, 4 * sizeof(double));
  alloc((void**)&m->d_conv, 2 * mt * 32 * sizeof(int));
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&m->num_sms, cudaDevAttrMultiProcessorCount, device);
  {
    const char* th = getenv("AQC_MPS_THETA");
    m->theta_scalar = th && std::string(th) == "scalar";
    const char* sv = getenv("AQC_MPS_SVD");
    m->svd_precond = !(sv && std::string(sv) == "plain");
    const char* fe = getenv("AQC_MPS_FENCE");
    m->svd_fence = fe && std::string(fe) == "1";
    const char* cl = getenv("AQC_MPS_CLUSTER");
    m->svd_cluster = (cl && (*cl == '1' || *cl == '2' || *cl == '4') && !cl[1]) ? (*cl - '0') : 0;
    const char* st = getenv("AQC_MPS_SVD_TOL");
    m->svd_relaxed = !(st && std::string(st) == "strict");
  }
  for (MpsProgram* p : {&m->fwd, &m->dag}) {
    alloc((void**)&p->d_tasks, p->tasks.size() * sizeof(MpsTask));
    if (e == cudaSuccess)
      e = cudaMemcpy(p->d_tasks, p->tasks.data(), p->tasks.size() * sizeof(MpsTask),
                     cudaMemcpyHostToDevice);
    alloc((void**)&p->d_env, p->env.size() * sizeof(EnvTask));
    if (e == cudaSuccess)
      e = cudaMemcpy(p->d_env, p->env.data(), p->env.size() * sizeof(EnvTask), cudaMemcpyHostToDevice);
  }
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreate(&m->ev0);
  if (e == cudaSuccess) e = cudaEventCreate(&m->ev1);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(mps_env_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(mps_env_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)(kEnvSmemDoubles * sizeof(double)));
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(mps_rho_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    const int code = (e == cudaErrorMemoryAllocation) ? AQC_ENOMEM : AQC_ECUDA;
    aqc_fail(code, "MPS workspace setup failed: %s", cudaGetErrorString(e));
    std::string keep = aqc_last_error();
    aqc_mps_destroy(m);
    aqc_fail(code, "%s", keep.c_str());
    return code;
  }
  *out = m;
  return AQC_OK;
}

extern "C" int aqc_mps_bond_capacity(const aqc_mps* m) { return m ? m->C : AQC_EINVAL; }

// Host layout of a state (all sites padded to the capacity C):
//   gam: complex128 [n][2][C][C], lam: float64 [n+1][C] (bond j left of site j; bonds 0, n = [1]),
//   dims: int32 [n+1].
extern "C" int aqc_mps_upload(aqc_mps* m, int slot, const double* gam, const double* lam,
                              const int32_t* dims) {
  int rc = mps_check_slot(m, slot);
  if (rc) return rc;
  if (!gam || !lam || !dims) return aqc_fail(AQC_EINVAL, "null argument");
  for (int j = 0; j <= m->n; ++j)
    if (dims[j] < 1 || dims[j] > m->C) return aqc_fail(AQC_EINVAL, "bond dimension out of range");
  if (dims[0] != 1 || dims[m->n] != 1) return aqc_fail(AQC_EINVAL, "outer bonds must have dimension 1");
  MCU(cudaSetDevice(m->device));
  MCU(cudaMemcpyAsync(m->st[slot].gam, gam, gam_elems(m) * sizeof(double2), cudaMemcpyHostToDevice, m->stream));
  MCU(cudaMemcpyAsync(m->st[slot].lam, lam, lam_elems(m) * sizeof(double), cudaMemcpyHostToDevice, m->stream));
  MCU(cudaMemcpyAsync(m->st[slot].dims, dims, (m->n + 1) * sizeof(int), cudaMemcpyHostToDevice, m->stream));
  MCU(cudaStreamSynchronize(m->stream));
  return AQC_OK;
}

extern "C" int aqc_mps_download(aqc_mps* m, int slot, double* gam, double* lam, int32_t* dims) {
  int rc = mps_check_slot(m, slot);
  if (rc) return rc;
  if (!gam || !lam || !dims) return aqc_fail(AQC_EINVAL, "null argument");
  MCU(cudaSetDevice(m->device));
  MCU(cudaMemcpyAsync(gam, m->st[slot].gam, gam_elems(m) * sizeof(double2), cudaMemcpyDeviceToHost, m->stream));
  MCU(cudaMemcpyAsync(lam, m->st[slot].lam, lam_elems(m) * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
  MCU(cudaMemcpyAsync(dims, m->st[slot].dims, (m->n + 1) * sizeof(int), cudaMemcpyDeviceToHost, m->stream));
  MCU(cudaStreamSynchronize(m->stream));
  return AQC_OK;
}

static int copy_state_async(aqc_mps* m, int src, int dst) {
  if (src == dst) return AQC_OK;
  MCU(cudaMemcpyAsync(m->st[dst].gam, m->st[src].gam, gam_elems(m) * sizeof(double2), cudaMemcpyDeviceToDevice, m->stream));
  MCU(cudaMemcpyAsync(m->st[dst].lam, m->st[src].lam, lam_elems(m) * sizeof(double), cudaMemcpyDeviceToDevice, m->stream));
  MCU(cudaMemcpyAsync(m->st[dst].dims, m->st[src].dims, (m->n + 1) * sizeof(int), cudaMemcpyDeviceToDevice, m->stream));
  return AQC_OK;
}

static int set_product_async(aqc_mps* m, int slot, long long index, int site = -1,
                             double2 a0 = make_double2(0.0, 0.0), double2 a1 = make_double2(0.0, 0.0)) {
  mps_set_product_kernel<<<m->n, 256, 0, m->stream>>>(mut(m, slot), index, m->n, m->C, site, a0, a1);
  MCU(cudaGetLastError());
  m->last_launches++;
  return AQC_OK;
}

extern "C" int aqc_mps_set_product(aqc_mps* m, int slot, int64_t index) {
  int rc = mps_check_slot(m, slot);
  if (rc) return rc;
  if (index < 0 || (m->n < 63 && index >= (1ll << m->n))) return aqc_fail(AQC_EINVAL, "basis index out of range");
  MCU(cudaSetDevice(m->device));
  rc = set_product_async(m, slot, index);
  if (rc) return rc;
  MCU(cudaStreamSynchronize(m->stream));
  return AQC_OK;
}

extern "C" int aqc_mps_set_product_site(aqc_mps* m, int slot, int64_t index, int site, const double* amps) {
  int rc = mps_check_slot(m, slot);
  if (rc) return rc;
  if (!amps) return aqc_fail(AQC_EINVAL, "null argument");
  if (index < 0 || (m->n < 63 && index >= (1ll << m->n))) return aqc_fail(AQC_EINVAL, "basis index out of range");
  if (site < 0 || site >= m->n) return aqc_fail(AQC_EINVAL, "site out of range");
  MCU(cudaSetDevice(m->device));
  rc = set_product_async(m, slot, index, site, make_double2(amps[0], amps[1]), make_double2(amps[2], amps[3]));
  if (rc) return rc;
  MCU(cudaStreamSynchronize(m->stream));
  return AQC_OK;
}

static int upload_thetas_async(aqc_mps* m, const double* thetas) {
  int rc = mps_ensure_pinned(m, (size_t)m->nthetas * 2 + 8192);
  if (rc) return rc;
  memcpy(m->h_pinned, thetas, (size_t)m->nthetas * sizeof(double));
  MCU(cudaMemcpyAsync(m->d_thetas, m->h_pinned, (size_t)m->nthetas * sizeof(double), cudaMemcpyHostToDevice, m->stream));
  return AQC_OK;
}

static int launch_algebra(aqc_mps* m, const MpsTask* tasks, int ntasks, int mode) {
  const int per = 32;  // tasks per CTA (4 threads each)
  dim3 grid((ntasks + per - 1) / per);
  switch (m->ent) {
    case AQC_ENT_CX:
      mps_algebra_kernel<AQC_ENT_CX><<<grid, per * 4, 0, m->stream>>>(tasks, ntasks, m->d_thetas, mode, m->d_gate, m->d_rho, m->d_gacc);
      break;
    case AQC_ENT_CZ:
      mps_algebra_kernel<AQC_ENT_CZ><<<grid, per * 4, 0, m->stream>>>(tasks, ntasks, m->d_thetas, mode, m->d_gate, m->d_rho, m->d_gacc);
      break;
    default:
      mps_algebra_kernel<AQC_ENT_CP><<<grid, per * 4, 0, m->stream>>>(tasks, ntasks, m->d_thetas, mode, m->d_gate, m->d_rho, m->d_gacc);
  }
  MCU(cudaGetLastError());
  m->last_launches++;
  return AQC_OK;
}

// applies the front layer (single-site tasks) of `prog` to the given states
static int run_front(aqc_mps* m, const MpsProgram& prog, bool dagger, const int* slots, int nstates) {
  int rc = launch_algebra(m, prog.d_tasks, prog.nfront, dagger ? 1 : 0);
  if (rc) return rc;
  for (int s = 0; s < nstates; ++s) {
    mps_apply1q_kernel<<<prog.nfront, 256, 0, m->stream>>>(mut(m, slots[s]), prog.d_tasks, m->d_gate, m->C);
    MCU(cudaGetLastError());
    m->last_launches++;
  }
  return AQC_OK;
}

// gate + SVD of one two-qubit step for the given states (theta0 saved when keep0)
static int run_step_theta(aqc_mps* m, const MpsProgram& prog, const MpsStep& st, const int* slots,
                          int nstates, bool with_gate, bool keep0, bool with_work) {
  ThetaArgs ta;
  memset(&ta, 0, sizeof(ta));
  for (int s = 0; s < nstates; ++s) ta.st[s] = view(m, slots[s]);
  ta.tasks = prog.d_tasks + st.task0;
  ta.gate = with_gate ? m->d_gate : nullptr;
  ta.theta0 = keep0 ? m->d_theta0 : nullptr;
  ta.work = with_work ? m->d_work : nullptr;
  ta.work0 = with_work ? m->d_work0 : nullptr;
  ta.C = m->C;
  ta.maxtasks = m->maxtasks;
  ta.single_site = 0;
  if (m->C >= 32 && !m->theta_scalar)
    mps_theta_dmma_kernel<<<dim3(st.ntasks * 4, nstates), 256, 0, m->stream>>>(ta);
  else
    mps_theta_kernel<<<dim3(st.ntasks, nstates), 512, 0, m->stream>>>(ta);
  MCU(cudaGetLastError());
  m->last_launches++;
  return AQC_OK;
}

static int run_step_svd(aqc_mps* m, const MpsProgram& prog, const MpsStep& st, const int* slots, int nstates) {
  SvdArgs sa;
  memset(&sa, 0, sizeof(sa));
  for (int s = 0; s < nstates; ++s) sa.st[s] = mut(m, slots[s]);
  sa.tasks = prog.d_tasks + st.task0;
  sa.work = m->d_work;
  sa.work0 = m->d_work0;
  sa.vmat = m->d_vmat;
  sa.C = m->C;
  sa.maxtasks = m->maxtasks;
  sa.chi_max = m->chi_max;
  sa.trunc_thr = m->trunc_thr;
  sa.trunc_stats = m->d_trunc;
  sa.tol_floor = m->svd_relaxed && m->trunc_thr >= 1e-10 ? fmin(1e-9, 1e-3 * m->trunc_thr) : 0.0;
  sa.sweeps = m->d_sweeps;
  sa.conv = m->d_conv;
  sa.precond = m->svd_precond ? 1 : 0;
  sa.fence = m->svd_fence ? 1 : 0;
  // as many CTAs per SVD as fit in one wave (cluster of 1, 2 or 4)
  int csize = 1;
  const int nsvd = st.ntasks * nstates;
  if (nsvd * 4 <= m->num_sms) csize = 4;
  else if (nsvd * 2 <= m->num_sms) csize = 2;
  if (m->svd_cluster > 0) csize = m->svd_cluster;  // AQC_MPS_CLUSTER (development switch)
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(st.ntasks * csize, nstates);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = m->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = csize;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  MCU(cudaLaunchKernelEx(&cfg, mps_svd_kernel, sa));
  MCU(cudaGetLastError());
  m->last_launches++;
  return AQC_OK;
}

static int apply_async(aqc_mps* m, const double* thetas, int dagger, int src, int dst) {
  int rc = upload_thetas_async(m, thetas);
  if (rc) return rc;
  rc = copy_state_async(m, src, dst);
  if (rc) return rc;
  const MpsProgram& prog = dagger ? m->dag : m->fwd;
  const int slots[1] = {dst};
  if (!dagger) {
    rc = run_front(m, prog, false, slots, 1);
    if (rc) return rc;
  }
  for (const MpsStep& st : prog.steps) {
    rc = launch_algebra(m, prog.d_tasks + st.task0, st.ntasks, dagger ? 1 : 0);
    if (rc) return rc;
    rc = run_step_theta(m, prog, st, slots, 1, true, false, true);
    if (rc) return rc;
    rc = run_step_svd(m, prog, st, slots, 1);
    if (rc) return rc;
  }
  if (dagger) {
    rc = run_front(m, prog, true, slots, 1);
    if (rc) return rc;
  }
  return AQC_OK;
}

extern "C" int aqc_mps_apply(aqc_mps* m, const double* thetas, int dagger, int src_slot, int dst_slot) {
  int rc = mps_check_slot(m, src_slot);
  if (rc) return rc;
  rc = mps_check_slot(m, dst_slot);
  if (rc) return rc;
  if (!thetas) return aqc_fail(AQC_EINVAL, "thetas is null");
  MCU(cudaSetDevice(m->device));
  m->last_launches = 0;
  cudaMemsetAsync(m->d_trunc, 0, 4 * sizeof(double), m->stream);
  MCU(cudaEventRecord(m->ev0, m->stream));
  rc = apply_async(m, thetas, dagger, src_slot, dst_slot);
  if (rc) return rc;
  MCU(cudaEventRecord(m->ev1, m->stream));
  MCU(cudaStreamSynchronize(m->stream));
  MCU(cudaEventElapsedTime(&m->last_ms, m->ev0, m->ev1));
  return AQC_OK;
}

static int amplitudes_async(aqc_mps* m, int slot, const int64_t* idx, int count) {
  if (count > 4096) return aqc_fail(AQC_EINVAL, "too many indices");
  MCU(cudaMemcpyAsync(m->d_idx, idx, (size_t)count * sizeof(long long), cudaMemcpyHostToDevice, m->stream));
  mps_amplitude_kernel<<<count, 64, 0, m->stream>>>(view(m, slot), m->d_idx, m->n, m->C, m->d_small);
  MCU(cudaGetLastError());
  m->last_launches++;
  return AQC_OK;
}

// out[i] = <idx_i | slot>   (MpsStateHandler.state_dot_vector for basis-state handlers)
extern "C" int aqc_mps_amplitudes(aqc_mps* m, int slot, const int64_t* idx, int count, double* out) {
  int rc = mps_check_slot(m, slot);
  if (rc) return rc;
  if (!idx || !out || count <= 0) return aqc_fail(AQC_EINVAL, "bad arguments");
  MCU(cudaSetDevice(m->device));
  rc = amplitudes_async(m, slot, idx, count);
  if (rc) return rc;
  MCU(cudaMemcpyAsync(out, m->d_small, (size_t)count * sizeof(double2), cudaMemcpyDeviceToHost, m->stream));
  MCU(cudaStreamSynchronize(m->stream));
  return AQC_OK;
}

extern "C" int aqc_mps_objective(aqc_mps* m, const double* thetas, int target_slot, int z0_slot,
                                 const int64_t* idx, int count, double* hs_out) {
  int rc = mps_check_slot(m, target_slot);
  if (rc) return rc;
  rc = mps_check_slot(m, z0_slot);
  if (rc) return rc;
  if (!thetas || !idx || !hs_out || count <= 0) return aqc_fail(AQC_EINVAL, "bad arguments");
  MCU(cudaSetDevice(m->device));
  m->last_launches = 0;
  cudaMemsetAsync(m->d_trunc, 0, 4 * sizeof(double), m->stream);
  MCU(cudaEventRecord(m->ev0, m->stream));
  rc = apply_async(m, thetas, 1, target_slot, z0_slot);
  if (rc) return rc;
  rc = amplitudes_async(m, z0_slot, idx, count);
  if (rc) return rc;
  MCU(cudaEventRecord(m->ev1, m->stream));
  MCU(cudaMemcpyAsync(hs_out, m->d_small, (size_t)count * sizeof(double2), cudaMemcpyDeviceToHost, m->stream));
  MCU(cudaStreamSynchronize(m->stream));
  MCU(cudaEventElapsedTime(&m->last_ms, m->ev0, m->ev1));
  return AQC_OK;
}

static int env_launch(aqc_mps* m, int wslot, int zslot, int count, const EnvTask* d_tasks) {
  if (count <= 0) return AQC_OK;
  EnvArgs ea;
  ea.w = view(m, wslot);
  ea.z = view(m, zslot);
  ea.tasks = d_tasks;
  ea.envL = m->d_envL;
  ea.envR = m->d_envR;
  ea.C = m->C;
  if (m->C >= 32 && m->C <= 64 && !m->theta_scalar)
    mps_env_dmma_kernel<<<(unsigned)count, 256, kEnvSmemDoubles * sizeof(double), m->stream>>>(ea);
  else
    mps_env_kernel<<<(unsigned)count, 512, 64 * 1024, m->stream>>>(ea);
  MCU(cudaGetLastError());
  m->last_launches++;
  return AQC_OK;
}

// full left-to-right and right-to-left environment sweeps of <w|z>
static int env_full_sweeps(aqc_mps* m, int wslot, int zslot, bool left_only) {
  mps_env_init_kernel<<<1, 32, 0, m->stream>>>(m->d_envL, m->d_envR, m->n, m->C);
  MCU(cudaGetLastError());
  m->last_launches++;
  const EnvTask* d = m->fwd.d_env;  // [L 0..n-1 | R n-1..0]
  for (int i = 0; i < m->n; ++i) {
    int rc = env_launch(m, wslot, zslot, 1, d + i);
    if (rc) return rc;
    if (!left_only) {
      rc = env_launch(m, wslot, zslot, 1, d + m->n + i);
      if (rc) return rc;
    }
  }
  return AQC_OK;
}

// <slot_a | slot_b>  (mps_dot, mps_operations.py:192-213)
extern "C" int aqc_mps_dot(aqc_mps* m, int slot_a, int slot_b, double* out) {
  int rc = mps_check_slot(m, slot_a);
  if (rc) return rc;
  rc = mps_check_slot(m, slot_b);
  if (rc) return rc;
  if (!out) return aqc_fail(AQC_EINVAL, "out is null");
  MCU(cudaSetDevice(m->device));
  m->last_launches = 0;
  cudaMemsetAsync(m->d_trunc, 0, 4 * sizeof(double), m->stream);
  rc = env_full_sweeps(m, slot_a, slot_b, true);
  if (rc) return rc;
  MCU(cudaMemcpyAsync((void*)out, m->d_envL + (size_t)m->n * m->C * m->C,
                      sizeof(double2), cudaMemcpyDeviceToHost, m->stream));
  MCU(cudaStreamSynchronize(m->stream));
  return AQC_OK;
}

// Complex gradient of <V x | y> given z0 = V^H y in MPS form (fast_dot_gradient).
// x = basis state |x_basis> (x_slot < 0) or the state in x_slot.  w_slot / z_slot receive V x, V z0.
extern "C" int aqc_mps_grad(aqc_mps* m, const double* thetas, int x_slot, int64_t x_basis, int z0_slot,
                            int w_slot, int z_slot, double* grad_out) {
  int rc = mps_check_slot(m, z0_slot);
  if (rc) return rc;
  rc = mps_check_slot(m, w_slot);
  if (rc) return rc;
  rc = mps_check_slot(m, z_slot);
  if (rc) return rc;
  if (x_slot >= 0 && (rc = mps_check_slot(m, x_slot))) return rc;
  if (w_slot == z_slot || w_slot == z0_slot) return aqc_fail(AQC_EINVAL, "slot aliasing");
  if (!thetas || !grad_out) return aqc_fail(AQC_EINVAL, "null argument");
  MCU(cudaSetDevice(m->device));
  m->last_launches = 0;
  cudaMemsetAsync(m->d_trunc, 0, 4 * sizeof(double), m->stream);
  const int n = m->n;
  auto done = [&](int code) {
    cudaStreamSynchronize(m->stream);
    return code;
  };
  MCU(cudaEventRecord(m->ev0, m->stream));
  rc = upload_thetas_async(m, thetas);
  if (rc) return done(rc);
  cudaMemsetAsync(m->d_gacc, 0, (size_t)m->nthetas * 2 * sizeof(double), m->stream);
  if (x_slot >= 0)
    rc = copy_state_async(m, x_slot, w_slot);
  else
    rc = set_product_async(m, w_slot, x_basis);
  if (rc) return done(rc);
  rc = copy_state_async(m, z0_slot, z_slot);
  if (rc) return done(rc);
  const int slots[2] = {w_slot, z_slot};
  const MpsProgram& prog = m->fwd;

  // environments of <w|z> at every bond
  rc = env_full_sweeps(m, w_slot, z_slot, false);
  if (rc) return done(rc);

  // ---- front layer: rho of every site, derivatives, then the gates
  {
    ThetaArgs ta;
    memset(&ta, 0, sizeof(ta));
    ta.st[0] = view(m, w_slot);
    ta.st[1] = view(m, z_slot);
    ta.tasks = prog.d_tasks;
    ta.theta0 = m->d_theta0;
    ta.C = m->C;
    ta.maxtasks = m->maxtasks;
    ta.single_site = 1;
    mps_theta_kernel<<<dim3(n, 2), 512, 0, m->stream>>>(ta);
    mps_zero_rho_kernel<<<(m->maxtasks * 16 + 255) / 256, 256, 0, m->stream>>>(m->d_rho, m->maxtasks * 16);
    RhoArgs ra;
    ra.w = view(m, w_slot);
    ra.z = view(m, z_slot);
    ra.tasks = prog.d_tasks;
    ra.theta0 = m->d_theta0;
    ra.envL = m->d_envL;
    ra.envR = m->d_envR;
    ra.rho = m->d_rho;
    ra.C = m->C;
    ra.maxtasks = m->maxtasks;
    ra.nphys = 2;
    ra.span = 1;
    mps_rho_kernel<<<dim3(n, 2), 512, 128 * 1024, m->stream>>>(ra);
    MCU(cudaGetLastError());
    m->last_launches += 3;
    rc = launch_algebra(m, prog.d_tasks, n, 2);
    if (rc) return done(rc);
    rc = run_front(m, prog, false, slots, 2);
    if (rc) return done(rc);
  }

  // ---- two-qubit steps
  for (size_t si = 0; si < prog.steps.size(); ++si) {
    const MpsStep& st = prog.steps[si];
    const MpsTask* tasks = prog.d_tasks + st.task0;
    rc = launch_algebra(m, tasks, st.ntasks, 0);  // 4x4 gates of the step
    if (rc) return done(rc);
    rc = run_step_theta(m, prog, st, slots, 2, true, true, true);
    if (rc) return done(rc);
    RhoArgs ra;
    ra.w = view(m, w_slot);
    ra.z = view(m, z_slot);
    ra.tasks = tasks;
    ra.theta0 = m->d_theta0;
    ra.envL = m->d_envL;
    ra.envR = m->d_envR;
    ra.rho = m->d_rho;
    ra.C = m->C;
    ra.maxtasks = m->maxtasks;
    ra.nphys = 4;
    ra.span = 2;
    mps_rho_kernel<<<dim3(st.ntasks, 4), 512, 128 * 1024, m->stream>>>(ra);
    MCU(cudaGetLastError());
    m->last_launches++;
    rc = launch_algebra(m, tasks, st.ntasks, 2);  // derivatives
    if (rc) return done(rc);
    rc = run_step_svd(m, prog, st, slots, 2);
    if (rc) return done(rc);
    // refresh the environments at the bonds this step modified: bond k+1 of every pair (k, k+1)
    rc = env_launch(m, w_slot, z_slot, 2 * st.ntasks, prog.d_env + prog.env_step0[si]);
    if (rc) return done(rc);
  }
  MCU(cudaEventRecord(m->ev1, m->stream));
  rc = mps_ensure_pinned(m, (size_t)m->nthetas * 2 + 8192);
  if (rc) return done(rc);
  MCU(cudaMemcpyAsync(m->h_pinned, m->d_gacc, (size_t)m->nthetas * 2 * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
  MCU(cudaStreamSynchronize(m->stream));
  MCU(cudaEventElapsedTime(&m->last_ms, m->ev0, m->ev1));
  // raw sums -> 0.5j <P w|z>: same factor map as the state-vector path
  const int n3 = 3 * n, tpb = m->tpb, T = m->nthetas;
  for (int k = 0; k < T; ++k) {
    const double re = m->h_pinned[2 * k], im = m->h_pinned[2 * k + 1];
    int kind;
    if (k < n3)
      kind = (k % 3 == 1) ? 0 : 1;
    else {
      const int r = (k - n3) % tpb;
      kind = (r == 4) ? 2 : ((r == 0 || r == 2) ? 0 : 1);
    }
    if (kind == 0)
      grad_out[2 * k] = 0.5 * re, grad_out[2 * k + 1] = 0.5 * im;
    else if (kind == 1)
      grad_out[2 * k] = -0.5 * im, grad_out[2 * k + 1] = 0.5 * re;
    else
      grad_out[2 * k] = im, grad_out[2 * k + 1] = -re;
  }
  return AQC_OK;
}

extern "C" float aqc_mps_last_kernel_ms(const aqc_mps* m) { return m ? m->last_ms : 0.f; }

// Truncation record of the most recent apply / objective / grad call on this workspace:
// out[0] sum over all splits of the discarded weight (squared Schmidt values, relative to the split),
// out[1] largest single discard, out[2] the part of out[0] removed by the chi_max cap alone (beyond the
// trunc_thr rule of mps_operations.py:248-265, which has no cap), out[3] number of splits the cap cut.
extern "C" int aqc_mps_truncation_stats(aqc_mps* m, double* out4) {
  if (!m || !out4) return aqc_fail(AQC_EINVAL, "null argument");
  MCU(cudaSetDevice(m->device));
  MCU(cudaMemcpyAsync(out4, m->d_trunc, 4 * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
  MCU(cudaStreamSynchronize(m->stream));
  return AQC_OK;
}
extern "C" int aqc_mps_last_num_launches(const aqc_mps* m) { return m ? m->last_launches : 0; }

// Diagnostics: Jacobi sweeps used by the SVDs of the most recent two-qubit step
// (out[state * maxtasks + task]); returns the number of ints written.
extern "C" int aqc_mps_debug_sweeps(aqc_mps* m, int32_t* out, int cap) {
  if (!m || !out) return aqc_fail(AQC_EINVAL, "null argument");
  const int cnt = std::min(cap, 2 * m->maxtasks);
  MCU(cudaSetDevice(m->device));
  MCU(cudaMemcpy(out, m->d_sweeps, (size_t)cnt * sizeof(int), cudaMemcpyDeviceToHost));
  return cnt;
}
