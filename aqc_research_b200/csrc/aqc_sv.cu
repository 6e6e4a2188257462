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
#include <set>
#include <string>
#include <thread>
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

#include "aqc_program.h"
#include "aqc_legacy.cuh"
#include "aqc_dense.cuh"
#include "aqc_cd.cuh"
#include "aqc_sketch.cuh"
#include "aqc_small.cuh"

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
  double* h_pinned = nullptr;  // pinned staging for small results (the device writes them directly)
  double* h_thetas = nullptr;  // pinned, device-readable copy of the angles of the call in flight
  bool trig_in_global = false;   // > 12 800 angles: the (cos, sin) table does not fit shared memory
  size_t pinned_cap = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t ev_hs = nullptr;   // aqc_sv_eval_begin: the gathered amplitudes have reached pinned memory
  cudaEvent_t ev_obj = nullptr;  // ... and the end of its V^H sweep (kernel-time split for bench.py)
  float last_obj_ms = 0.f, last_grad_ms = 0.f;
  size_t eval_hs_off = 0, eval_hs_count = 0;  // where aqc_sv_eval_hs finds its result in h_pinned
  cudaEvent_t tm0 = nullptr, tm1 = nullptr;  // user timer (aqc_sv_timer_*)
  float last_ms = 0.f;
  int last_launches = 0;
  Program prog_grad, prog_fwd, prog_dag;
  // dense-stage engine (aqc_dense.cuh): DMMA sweeps; the default whenever the tile has >= 5 bits
  bool dense = false;
  bool grad_pending = false;  // aqc_sv_grad_begin enqueued, results not collected yet
  int num_sms = 148;
  DenseTables dt_grad, dt_fwd, dt_dag;
  double *d_umat = nullptr, *d_gm = nullptr;
  double* d_umat_grad = nullptr;        // stage matrices of the gradient program (own buffer: its prologue may
                                        // run on stream_aux while the V^H sweep still uses d_umat)
  cudaStream_t stream_aux = nullptr;    // aqc_sv_eval_begin: gradient prologue next to the V^H sweep
  cudaEvent_t ev_aux1 = nullptr;
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
  std::vector<void*> ipc_opened;  // what cudaIpcOpenMemHandle returned (closed on destroy)
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
  static bool configured[64] = {false};  // per device; setting the attribute twice is harmless  // per device
  if (!configured[sv->device & 63]) {
    CU(cudaFuncSetAttribute(pass_kernel<NVEC, ENT, DAG>,
                            cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)((size_t)NVEC * sizeof(double2) << kMaxTileBits)));
    configured[sv->device & 63] = true;
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

// Stages the angles of a call in pinned host memory, where the prologue / epilogue kernels of the dense
// engine read them directly (no H2D copy, no (cos, sin) table).  `to_device` additionally copies them
// to d_thetas and builds the (cos, sin) table: the legacy engine and coordinate descent need that.
static int upload_thetas(aqc_sv* sv, const double* thetas, bool to_device) {
  const size_t tot = (size_t)sv->batch * sv->circ.nthetas;
  if (sv->grad_pending) {  // an uncollected gradient sweep still reads the staged angles: drop it
    CU(cudaStreamSynchronize(sv->stream));
    sv->grad_pending = false;
  }
  memcpy(sv->h_thetas, thetas, tot * sizeof(double));
  if (!to_device && !sv->trig_in_global) return AQC_OK;
  CU(cudaMemcpyAsync(sv->d_thetas, sv->h_thetas, tot * sizeof(double), cudaMemcpyHostToDevice, sv->stream));
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
// Launch with the programmatic-stream-serialization attribute (see pdl_wait in aqc_dense.cuh): the kernel
// may be scheduled while its predecessor in the stream drains.  AQC_PDL=0 launches plainly.
static bool g_pdl = true;  // AQC_PDL, read whenever a workspace is created
template <typename... KArgs, typename... Args>
static cudaError_t launch_chained(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                  Args&&... args) {
  const bool pdl = g_pdl;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

template <int NVEC, int MINB>
static int launch_dense_pass_inst(aqc_sv* sv, const DensePassArgs& args) {
  const size_t smem = (size_t)NVEC * sizeof(double2) << args.pd.tb;
  static bool configured[64] = {false};  // per device; setting the attribute twice is harmless
  if (!configured[sv->device & 63]) {
    CU(cudaFuncSetAttribute(dense_pass_kernel<NVEC, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)((size_t)NVEC * sizeof(double2) << kMaxTileBits)));
    configured[sv->device & 63] = true;
  }
  dim3 grid((unsigned)(1ull << args.pd.nouter), (unsigned)sv->batch);
  CU(launch_chained(dense_pass_kernel<NVEC, MINB>, grid, dim3(kDThreads), smem, sv->stream, args));
  return AQC_OK;
}
template <int NVEC>
static int launch_dense_pass(aqc_sv* sv, const DensePassArgs& args) {
  if (NVEC == 1) {
    if (args.pd.tb == 12) return launch_dense_pass_inst<1, 3>(sv, args);
    return launch_dense_pass_inst<1, 4>(sv, args);
  }
  // (a two-CTA, four-iterations-in-flight instantiation of the gradient pass, 110 registers, was measured:
  // n = 20 gradient sweep 0.223 -> 0.264 ms, n = 24 6.13 -> 6.31 ms -- the third resident CTA is worth more)
  return launch_dense_pass_inst<2, 3>(sv, args);
}

// Prologue of a sweep: stage matrices of one program from the staged angles (mode 0 gradient, 1 V,
// 2 V^H); the gradient prologue also clears the stage-matrix sums.
static int dense_prepare(aqc_sv* sv, int mode, cudaStream_t stream = nullptr) {
  if (!stream) stream = sv->stream;
  const Program& p = mode == 0 ? sv->prog_grad : (mode == 1 ? sv->prog_fwd : sv->prog_dag);
  PrologueArgs a;
  memset(&a, 0, sizeof(a));
  a.stages = p.d_stages;
  a.nstages = (int)p.stages.size();
  a.nthetas = sv->circ.nthetas;
  a.batch = sv->batch;
  a.thetas = sv->h_thetas;
  a.umat = mode == 0 ? sv->d_umat_grad : sv->d_umat;
  if (mode == 0) {
    a.zero0 = sv->d_gm;
    a.nzero0 = (long long)sv->batch * a.nstages * 64;
  }
  if (a.nstages == 0 && a.nzero0 == 0 && a.nzero1 == 0) return AQC_OK;
  a.n3 = 3 * sv->circ.n;
  a.tpb = sv->circ.tpb;
  a.gtrig = sv->trig_in_global ? sv->d_trig : nullptr;
  a.trig_out = (mode == 0 && !sv->trig_in_global) ? sv->d_trig : nullptr;
  // (cos, sin) table of one batch element in shared memory (or in d_trig for very long circuits)
  const size_t dyn = sv->trig_in_global ? 0 : (size_t)sv->circ.nthetas * sizeof(double2);
  const dim3 grid((unsigned)std::max(1, (a.nstages * 4 + 127) / 128), (unsigned)sv->batch);
  const bool dag = mode == 2;
#define AQC_PRO(E)                                                                                        \
  do {                                                                                                    \
    if (dag) {                                                                                            \
      if (dyn > 48 * 1024)                                                                                \
        CU(cudaFuncSetAttribute(sweep_prologue_kernel<E, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn)); \
      CU(launch_chained(sweep_prologue_kernel<E, true>, grid, dim3(128), dyn, stream, a));            \
    } else {                                                                                              \
      if (dyn > 48 * 1024)                                                                                \
        CU(cudaFuncSetAttribute(sweep_prologue_kernel<E, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn)); \
      CU(launch_chained(sweep_prologue_kernel<E, false>, grid, dim3(128), dyn, stream, a));           \
    }                                                                                                     \
  } while (0)
  switch (sv->circ.ent) {
    case AQC_ENT_CX: AQC_PRO(AQC_ENT_CX); break;
    case AQC_ENT_CZ: AQC_PRO(AQC_ENT_CZ); break;
    default: AQC_PRO(AQC_ENT_CP);
  }
#undef AQC_PRO
  CU(cudaGetLastError());
  sv->last_launches += 1;
  return AQC_OK;
}

// Epilogue of the gradient sweep: complex gradient 0.5j <P w|z> from the accumulated stage matrices,
// written by the device into the pinned result buffer (h_pinned[0 .. 2 batch T)).
static int dense_collect(aqc_sv* sv) {
  const Program& p = sv->prog_grad;
  EpilogueArgs a;
  memset(&a, 0, sizeof(a));
  a.stages = p.d_stages;
  a.nstages = (int)p.stages.size();
  a.nthetas = sv->circ.nthetas;
  a.batch = sv->batch;
  a.n3 = 3 * sv->circ.n;
  a.tpb = sv->circ.tpb;
  a.gm = sv->d_gm;
  a.out = sv->h_pinned;
  a.out2 = sv->h_pinned + 2 * (size_t)sv->batch * sv->circ.nthetas;
  a.extra_seq = sv->circ.n + sv->circ.nb;
  a.gtrig = sv->d_trig;
  // the (cos, sin) table of one batch element is staged in shared memory when it fits the default window
  const size_t table = (size_t)sv->circ.nthetas * sizeof(double2);
  a.trig_smem = table <= 40 * 1024 ? 1 : 0;
  const size_t dyn = a.trig_smem ? table : 0;
  const dim3 grid((unsigned)std::max(1, (a.nstages * kStageUnits * 4 + 127) / 128), (unsigned)sv->batch);
#define AQC_EPI(E)                                                                                                  \
  do {                                                                                                              \
    if (dyn > 48 * 1024)                                                                                            \
      CU(cudaFuncSetAttribute(grad_epilogue_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));     \
    CU(launch_chained(grad_epilogue_kernel<E>, grid, dim3(128), dyn, sv->stream, a));                               \
  } while (0)
  switch (sv->circ.ent) {
    case AQC_ENT_CX: AQC_EPI(AQC_ENT_CX); break;
    case AQC_ENT_CZ: AQC_EPI(AQC_ENT_CZ); break;
    default: AQC_EPI(AQC_ENT_CP);
  }
#undef AQC_EPI
  CU(cudaGetLastError());
  sv->last_launches += 1;
  return AQC_OK;
}

// The finished gradient of the dense engine from the pinned buffer the epilogue wrote (after a synchronize):
// the first array plus, for the angles a trailing half layer shares with the first layer, the second one.
static void dense_gradient_from_pinned(const aqc_sv* sv, double* grad_out) {
  const size_t T = (size_t)sv->circ.nthetas, tot = (size_t)sv->batch * T;
  memcpy(grad_out, sv->h_pinned, tot * 2 * sizeof(double));
  if (sv->circ.half <= 0) return;
  const size_t k0 = 2 * (size_t)(3 * sv->circ.n), k1 = k0 + 2 * (size_t)sv->circ.half * sv->circ.tpb;
  for (int b = 0; b < sv->batch; ++b) {
    const double* second = sv->h_pinned + 2 * tot + (size_t)b * T * 2;
    double* g = grad_out + (size_t)b * T * 2;
    for (size_t k = k0; k < k1; ++k) g[k] += second[k];
  }
}

// passes [pass_begin, pass_end) of a program on the dense engine; mode 0: (w, z), else one vector
static int run_dense_program(aqc_sv* sv, int mode, const double2* src0, long long basis,
                             const double2* src1, double2* dst0, double2* dst1, int pass_begin,
                             int pass_end, int push0 = -1, int push1 = -1) {
  const Program& prog = mode == 0 ? sv->prog_grad : (mode == 1 ? sv->prog_fwd : sv->prog_dag);
  const DenseTables& dt = mode == 0 ? sv->dt_grad : (mode == 1 ? sv->dt_fwd : sv->dt_dag);
  DensePassArgs a;
  memset(&a, 0, sizeof(a));
  a.vec_stride = sv->size;
  a.lanes = dt.d_lanes;
  a.umat = mode == 0 ? sv->d_umat_grad : sv->d_umat;
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
    a.xchg_world = 0;
    a.tile_xor = 0;
    if (push0 >= 0 && i + 1 == pass_end) {  // the last pass of the range delivers its result to the peers
      a.xchg_world = 1 << sv->g;
      a.xchg_rank = sv->rank;
      a.xchg_shift = sv->nbits - sv->g;
      for (int b = 0; b < sv->g; ++b)  // chunk bit b of the local index, if it numbers tiles (outer bit k)
        for (int k = 0; k < a.pd.nouter; ++k)
          if (a.pd.outerpos[k] == a.xchg_shift + b && ((sv->rank >> b) & 1)) a.tile_xor |= 1u << k;
      for (int r = 0; r < a.xchg_world; ++r) {
        a.xdst[0][r] = r == sv->rank ? sv->slots[push0] : const_cast<double2*>(sv->peer[push0][r]);
        a.xdst[1][r] = mode != 0 ? nullptr
                                 : (r == sv->rank ? sv->slots[push1] : const_cast<double2*>(sv->peer[push1][r]));
        if (!a.xdst[0][r] || (mode == 0 && !a.xdst[1][r]))
          return fail(AQC_EINVAL, "peer %d: destination slot was not imported", r);
      }
    }
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

static int upload_program(Program& p) {
  if (p.stages.empty()) return AQC_OK;
  CU(cudaMalloc(&p.d_stages, p.stages.size() * sizeof(StageDesc)));
  CU(cudaMemcpy(p.d_stages, p.stages.data(), p.stages.size() * sizeof(StageDesc),
                cudaMemcpyHostToDevice));
  CU(cudaMalloc(&p.d_passes, p.passes.size() * sizeof(PassDesc)));
  CU(cudaMemcpy(p.d_passes, p.passes.data(), p.passes.size() * sizeof(PassDesc), cudaMemcpyHostToDevice));
  return AQC_OK;
}

extern "C" void aqc_sv_destroy(aqc_sv* sv) {
  if (!sv) return;
  cudaSetDevice(sv->device);
  if (sv->stream) cudaStreamSynchronize(sv->stream);
  for (void* p : sv->ipc_opened) cudaIpcCloseMemHandle(p);
  sv->ipc_opened.clear();
  for (auto p : sv->slots)
    if (p) cudaFree(p);
  if (sv->d_thetas) cudaFree(sv->d_thetas);
  if (sv->d_trig) cudaFree(sv->d_trig);
  if (sv->d_gacc) cudaFree(sv->d_gacc);
  if (sv->d_scratch) cudaFree(sv->d_scratch);
  if (sv->d_idx) cudaFree(sv->d_idx);
  for (void* q : {(void*)sv->d_cd_units, (void*)sv->d_cd_fobj, (void*)sv->d_target, (void*)sv->d_gram,
                  (void*)sv->d_rinv, (void*)sv->d_info})
    if (q) cudaFree(q);
  for (void* q : {(void*)sv->d_umat, (void*)sv->d_umat_grad, (void*)sv->d_gm, (void*)sv->dt_grad.d_lanes, (void*)sv->dt_fwd.d_lanes,
                  (void*)sv->dt_dag.d_lanes})
    if (q) cudaFree(q);
  if (sv->h_pinned) cudaFreeHost(sv->h_pinned);
  if (sv->h_thetas) cudaFreeHost(sv->h_thetas);
  for (Program* p : {&sv->prog_grad, &sv->prog_fwd, &sv->prog_dag})
    if (p->d_stages) cudaFree(p->d_stages), cudaFree(p->d_passes);
  if (sv->ev_aux1) cudaEventDestroy(sv->ev_aux1);
  if (sv->stream_aux) cudaStreamDestroy(sv->stream_aux);
  if (sv->ev_hs) cudaEventDestroy(sv->ev_hs);
  if (sv->ev_obj) cudaEventDestroy(sv->ev_obj);
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
  CUB(cudaStreamCreateWithFlags(&sv->stream_aux, cudaStreamNonBlocking));
  CUB(cudaEventCreateWithFlags(&sv->ev_aux1, cudaEventDisableTiming));
  CUB(cudaEventCreate(&sv->ev_hs));
  CUB(cudaEventCreate(&sv->ev_obj));
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
  CUB(cudaMallocHost(&sv->h_thetas, std::max<size_t>(tot, 1) * sizeof(double)));
  sv->trig_in_global = (size_t)circ->nthetas * sizeof(double2) > 200 * 1024;
#undef CUB
  // Tile shape.  States beyond the L2 (> 64 MiB) want long contiguous runs and the largest tile (with the
  // tile planner 128-byte runs, 3 low bits, beat 256-byte ones: one more free tile bit saves two of eleven
  // passes at n = 28, 114.3 -> 109.7 ms per gradient sweep); L2-resident states (nbits <= 21) have too few tiles to fill 3 CTAs on each of the
  // SMs, so they use smaller gradient tiles and spend the low bits on gate qubits instead
  // (measured at n = 20: 0.42 -> 0.36 ms per evaluation).
  g_pdl = env_int("AQC_PDL", 1) != 0;
  // (n = 22: w and z together are 128 MiB, more than the L2 holds; with the planner the large-state tiles
  // win there too: 836 -> 869 evals/s)
  const bool l2_resident = sv->nbits <= 21;
  // tiny states (one vector <= 128 KiB) are latency bound: smaller tiles spread the few amplitudes over
  // more SMs (n = 12: 5 797 -> 6 328 evals/s with 2^8 / 2^9 tiles)
  const bool tiny = sv->nbits <= 13 && batch == 1;
  const bool small = sv->nbits <= 17 && batch == 1;  // n = 16: 6 046 -> 6 326 evals/s with 2^9 / 2^10 tiles
  const int tb_grad =
      std::min(env_int("AQC_TILE_BITS_GRAD", tiny ? 8 : (small ? 9 : (l2_resident ? 10 : 11))), kMaxTileBits);
  // single-vector sweeps of large states: 2^12-amplitude tiles (64 KiB) with 128-byte runs need fewer
  // passes (n = 28: 16 -> 13, 47.2 -> 45.5 ms)
  const int tb_apply =
      std::min(env_int("AQC_TILE_BITS_APPLY", tiny ? 9 : (small ? 10 : (l2_resident ? 11 : 12))), kMaxTileBits);
  const int low = env_int("AQC_TILE_LOW_BITS", l2_resident ? 2 : 3);
  const int low_apply =
      env_int("AQC_TILE_LOW_BITS_APPLY", env_int("AQC_TILE_LOW_BITS", l2_resident ? 1 : (sv->nbits <= 22 ? 2 : 3)));
  // engine: dense-stage DMMA sweeps (default) or "legacy" (gate-by-gate register kernel)
  {
    const char* eng = getenv("AQC_ENGINE");
    sv->dense = !(eng && std::string(eng) == "legacy");
    // fewer than 8 amplitude quadruples: nothing for a DMMA to do
    if (sv->nbits < kDMinTileBits || tb_grad < kDMinTileBits || tb_apply < kDMinTileBits) sv->dense = false;
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
    // grad_epilogue_kernel stores every angle's inner product from the ONE unit that owns it: each angle
    // must be covered exactly once by the gradient program
    // (the trailing half layer of a second-order circuit repeats the first one's angles: those occurrences
    // go to a second array, see EpilogueArgs::out2)
    std::vector<int> cover((size_t)sv->circ.nthetas, 0), cover2((size_t)sv->circ.nthetas, 0);
    const int extra_seq = sv->circ.n + sv->circ.nb;
    for (const StageDesc& sd : sv->prog_grad.stages)
      for (int u = 0; u < sd.nunits; ++u) {
        const int k = sd.u[u].kind;
        const int na = (k == U_FRONT_LO || k == U_FRONT_HI) ? 3 : (k == U_NONE ? 0 : sv->circ.tpb);
        for (int j = 0; j < na; ++j) {
          const int th = sd.u[u].theta + j;
          if (th < 0 || th >= sv->circ.nthetas) {
            fail(AQC_EINVAL, "internal: unit angle %d outside the circuit's %d angles", th, sv->circ.nthetas);
            return bail(AQC_EINVAL);
          }
          (sd.u[u].slot >= 5 * extra_seq ? cover2 : cover)[(size_t)th] += 1;
        }
      }
    const int n3 = 3 * sv->circ.n, shared_end = n3 + sv->circ.half * sv->circ.tpb;
    for (int th = 0; th < sv->circ.nthetas; ++th) {
      const int want2 = (th >= n3 && th < shared_end) ? 1 : 0;
      if (cover[(size_t)th] != 1 || cover2[(size_t)th] != want2) {
        fail(AQC_EINVAL, "internal: angle %d is covered by %d + %d units of the gradient program", th,
             cover[(size_t)th], cover2[(size_t)th]);
        return bail(AQC_EINVAL);
      }
    }
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
      e = cudaMalloc(&sv->d_umat_grad, B * std::max<size_t>(1, sv->prog_grad.stages.size()) * 64 * sizeof(double));
    if (e == cudaSuccess)
      e = cudaMalloc(&sv->d_gm, B * std::max<size_t>(1, sv->prog_grad.stages.size()) * 64 * sizeof(double));
    if (e != cudaSuccess) {
      fail(AQC_ENOMEM, "dense scratch allocation failed: %s", cudaGetErrorString(e));
      return bail(AQC_ENOMEM);
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

// The index list of a gather is the same on every objective call: it is uploaded only when it changes
// (a copy from pageable host memory stalls the submitting thread).
static int gather_indices(aqc_sv* sv, const int64_t* idx, int count) {
  int rc = ensure_idx(sv, (size_t)count);
  if (rc) return rc;
  for (int i = 0; i < count; ++i)
    if (idx[i] < 0 || idx[i] >= sv->size) return fail(AQC_EINVAL, "gather index out of range");
  if (sv->idx_cached.size() != (size_t)count || memcmp(sv->idx_cached.data(), idx, (size_t)count * sizeof(int64_t))) {
    sv->idx_cached.assign(idx, idx + count);
    CU(cudaStreamSynchronize(sv->stream));  // a previous gather may still read d_idx
    CU(cudaMemcpy(sv->d_idx, sv->idx_cached.data(), (size_t)count * sizeof(long long), cudaMemcpyHostToDevice));
  }
  return AQC_OK;
}

static int gather_async(aqc_sv* sv, int slot, const int64_t* idx, int count) {
  int rc = gather_indices(sv, idx, count);
  if (rc) return rc;
  rc = ensure_scratch(sv, (size_t)2 * count * sv->batch);
  if (rc) return rc;
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

static int apply_async(aqc_sv* sv, const double* thetas, int dagger, int src_slot, int dst_slot,
                       bool thetas_to_device = false) {
  int rc = ensure_pinned(sv, (size_t)sv->batch * sv->circ.nthetas * 2 + 64);
  if (rc) return rc;
  rc = upload_thetas(sv, thetas, thetas_to_device || !sv->dense);
  if (rc) return rc;
  CU(cudaEventRecord(sv->ev0, sv->stream));
  if (sv->dense) {
    rc = dense_prepare(sv, dagger ? 2 : 1);
    if (!rc)
      rc = run_dense_program(sv, dagger ? 2 : 1, sv->slots[src_slot], -1, nullptr, sv->slots[dst_slot],
                             nullptr, 0, -1);
  } else {
    const Program& prog = dagger ? sv->prog_dag : sv->prog_fwd;
    rc = run_program(sv, prog, false, dagger != 0, sv->slots[src_slot], -1, nullptr,
                     sv->slots[dst_slot], nullptr);
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
    rc = apply_async(sv, thetas, 1, target_slot, z_slot, true);  // the sweep kernel updates d_thetas
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
  rc = gather_indices(sv, idx, count);
  if (rc) return rc;
  rc = apply_async(sv, thetas, 1, target_slot, z0_slot);
  if (rc) return rc;
  // the gathered amplitudes go straight into the pinned result buffer
  CU(launch_chained(gather_out_kernel, dim3((count + 127) / 128, sv->batch), dim3(128), 0, sv->stream,
                    (const double2*)sv->slots[z0_slot], (long long)sv->size, (const long long*)sv->d_idx, count,
                    reinterpret_cast<double2*>(sv->h_pinned)));
  sv->last_launches += 1;
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
  rc = ensure_pinned(sv, tot * 4 + 64);  // (two arrays: see EpilogueArgs::out2)
  if (rc) return rc;
  rc = upload_thetas(sv, thetas, !sv->dense);
  if (rc) return rc;
  if (!sv->dense) CU(cudaMemsetAsync(sv->d_gacc, 0, tot * 2 * sizeof(double), sv->stream));
  CU(cudaEventRecord(sv->ev0, sv->stream));
  if (sv->dense) {
    rc = dense_prepare(sv, 0);
    if (!rc)
      rc = run_dense_program(sv, 0, x_slot >= 0 ? sv->slots[x_slot] : nullptr, x_slot >= 0 ? -1 : x_basis,
                             sv->slots[z0_slot], sv->slots[w_slot], sv->slots[z_slot], 0, -1);
    if (!rc) rc = dense_collect(sv);
  } else {
    rc = run_program(sv, sv->prog_grad, true, false, x_slot >= 0 ? sv->slots[x_slot] : nullptr,
                     x_slot >= 0 ? -1 : x_basis, sv->slots[z0_slot], sv->slots[w_slot],
                     sv->slots[z_slot]);
  }
  if (rc) return rc;
  CU(cudaEventRecord(sv->ev1, sv->stream));
  if (!sv->dense)  // (the dense epilogue has written the finished gradient into h_pinned itself)
    CU(cudaMemcpyAsync(sv->h_pinned, sv->d_gacc, tot * 2 * sizeof(double), cudaMemcpyDeviceToHost, sv->stream));
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
  if (sv->dense) {
    dense_gradient_from_pinned(sv, grad_out);
    return AQC_OK;
  }
  // legacy engine: raw sums -> 0.5j <P w|z>: Ry 0.5, Rz/Rx 0.5j, CPhase -i
  const int n3 = 3 * sv->circ.n, tpb = sv->circ.tpb, T = sv->circ.nthetas;
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


// ------------------------------------------------------------------------------------------
// One evaluation = objective + gradient at the same angles, enqueued in ONE go (dense engine).
// scipy's L-BFGS-B asks for fun(theta) and then jac(theta) (optimizer.py:585-590): instead of
// sweeping, waiting, returning to Python and only then enqueueing the gradient sweep, everything --
// V^H sweep, gather of hs, gradient sweep from the basis state |x_basis>, epilogue -- goes to the
// stream at once.  aqc_sv_eval_hs waits for an EVENT recorded behind the gather (the gradient sweep
// keeps running on the device while the host forms the objective value); aqc_sv_grad_end collects the
// gradient.  A gradient the caller turns out not to need (another leading state, other angles) is
// simply dropped by the next call.
// ------------------------------------------------------------------------------------------
extern "C" int aqc_sv_eval_begin(aqc_sv* sv, const double* thetas, int target_slot, int z0_slot,
                                 const int64_t* idx, int count, int64_t x_basis, int w_slot, int z_slot) {
  int rc = check_slot(sv, target_slot);
  if (!rc) rc = check_slot(sv, z0_slot);
  if (!rc) rc = check_slot(sv, w_slot);
  if (!rc) rc = check_slot(sv, z_slot);
  if (rc) return rc;
  if (!thetas || !idx || count <= 0) return fail(AQC_EINVAL, "bad arguments");
  if (!sv->dense) return fail(AQC_EINVAL, "aqc_sv_eval_begin needs the dense engine");
  if (sv->g != 0) return fail(AQC_EINVAL, "aqc_sv_eval_begin is for unsharded workspaces (sharded states run epoch by epoch)");
  if (x_basis < 0 || x_basis >= sv->size) return fail(AQC_EINVAL, "basis index out of range");
  if (w_slot == z_slot || w_slot == z0_slot || w_slot == target_slot || z_slot == target_slot || z0_slot == target_slot)
    return fail(AQC_EINVAL, "slot aliasing");
  CU(cudaSetDevice(sv->device));
  sv->last_launches = 0;
  const size_t tot = (size_t)sv->batch * sv->circ.nthetas;
  const size_t nout = (size_t)2 * count * sv->batch;
  rc = ensure_pinned(sv, 4 * tot + nout + 64);
  if (rc) return rc;
  rc = gather_indices(sv, idx, count);
  if (rc) return rc;
  rc = upload_thetas(sv, thetas, false);
  if (rc) return rc;
  CU(cudaEventRecord(sv->ev0, sv->stream));
  // The V^H sweep is enqueued first (it is the critical path); the gradient program's prologue (stage
  // matrices, cleared sums, trig table) does not depend on it and runs next to it on the auxiliary stream,
  // behind everything that was in the main stream when the evaluation began (ev0).
  rc = dense_prepare(sv, 2);
  if (!rc) rc = run_dense_program(sv, 2, sv->slots[target_slot], -1, nullptr, sv->slots[z0_slot], nullptr, 0, -1);
  if (rc) return rc;
  CU(cudaEventRecord(sv->ev_obj, sv->stream));
  CU(cudaStreamWaitEvent(sv->stream_aux, sv->ev0, 0));
  rc = dense_prepare(sv, 0, sv->stream_aux);
  if (rc) return rc;
  CU(cudaEventRecord(sv->ev_aux1, sv->stream_aux));
  sv->eval_hs_off = 4 * tot;  // behind the gradient's two arrays: the results never share bytes
  sv->eval_hs_count = nout;
  // hs leaves on the auxiliary stream, next to the gradient sweep (which only reads z0): the gather is
  // not on the critical path of the evaluation
  CU(cudaStreamWaitEvent(sv->stream_aux, sv->ev_obj, 0));
  gather_out_kernel<<<dim3((count + 127) / 128, sv->batch), 128, 0, sv->stream_aux>>>(
      sv->slots[z0_slot], sv->size, sv->d_idx, count, reinterpret_cast<double2*>(sv->h_pinned + sv->eval_hs_off));
  CU(cudaGetLastError());
  sv->last_launches += 1;
  CU(cudaEventRecord(sv->ev_hs, sv->stream_aux));
  CU(cudaStreamWaitEvent(sv->stream, sv->ev_aux1, 0));
  if (!rc)
    rc = run_dense_program(sv, 0, nullptr, x_basis, sv->slots[z0_slot], sv->slots[w_slot], sv->slots[z_slot], 0, -1);
  if (!rc) rc = dense_collect(sv);
  if (rc) return rc;
  // (the main stream ends behind the gather as well: a synchronize on it covers the whole evaluation)
  CU(cudaStreamWaitEvent(sv->stream, sv->ev_hs, 0));
  CU(cudaEventRecord(sv->ev1, sv->stream));
  sv->grad_pending = true;
  return AQC_OK;
}

// 1 if aqc_sv_eval_begin is available on this workspace (dense engine, unsharded), else 0.
extern "C" int aqc_sv_can_eval(const aqc_sv* sv) { return (sv && sv->dense && sv->g == 0) ? 1 : 0; }

// hs of the evaluation in flight (the gradient sweep may still be running).
extern "C" int aqc_sv_eval_hs(aqc_sv* sv, double* hs_out) {
  if (!sv || !hs_out) return fail(AQC_EINVAL, "null pointer argument");
  if (!sv->grad_pending || sv->eval_hs_count == 0) return fail(AQC_EINVAL, "no evaluation in flight (aqc_sv_eval_begin)");
  CU(cudaSetDevice(sv->device));
  CU(cudaEventSynchronize(sv->ev_hs));
  memcpy(hs_out, sv->h_pinned + sv->eval_hs_off, sv->eval_hs_count * sizeof(double));
  sv->eval_hs_count = 0;
  return AQC_OK;
}

// Kernel-time split of the last COMPLETED aqc_sv_eval_begin .. aqc_sv_grad_end pair (ms): V^H sweep
// (prologue + passes) and gradient sweep (prologue + passes + epilogue).
extern "C" int aqc_sv_eval_times(aqc_sv* sv, float* obj_ms, float* grad_ms) {
  if (!sv || !obj_ms || !grad_ms) return fail(AQC_EINVAL, "null pointer argument");
  CU(cudaSetDevice(sv->device));
  CU(cudaEventElapsedTime(obj_ms, sv->ev0, sv->ev_obj));
  CU(cudaEventElapsedTime(grad_ms, sv->ev_obj, sv->ev1));
  return AQC_OK;
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



#include "aqc_shard.cuh"
