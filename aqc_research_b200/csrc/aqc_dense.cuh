// aqc_dense.cuh -- "dense-stage" sweeps on the FP64 tensor pipe (included by aqc_sv.cu after the
// program structures).
//
// A *stage* is every consecutive gate of the circuit that acts on one pair of tile bits (front
// gates of the two qubits, the unit blocks of a Trotter triplet, their Rz(-+pi/2)); its action on
// an amplitude quadruple is ONE 4x4 complex matrix U_s(theta).  The gradient sweep of the reference
// (core_operations.py:823-1019) records 0.5j <P w|z> after every rotation; with w_k = G_k w_in and
// z_k = G_k z_in inside the stage (G_k = partial product up to rotation k),
//     sum_quads <P w_k | z_k> = Tr(G_k^H P G_k  M),      M = sum_quads z_in w_in^H   (4x4 complex),
// so the sweep over the 2^n amplitudes only has to (1) apply U_s to w and z and (2) accumulate the
// 4x4 matrix sum_quads z w^H; all per-rotation inner products follow from that matrix in a tiny
// post-processing kernel (dense_grad_kernel) that re-runs the gate-by-gate recipe of the reference
// on four "virtual quadruples".  Per quadruple and stage this is 128 + 64 real FMAs instead of
// ~130 per UNIT, and it is pure small-matrix work, which mma.sync.m8n8k4.f64 (DMMA) executes at
// the full FP64 rate (37.1 TFLOP/s measured, scripts/ubench_fp64.cu) with ~1/16 of the issue slots
// and registers of scalar DFMA code.
//
// Fragment algebra (PTX m8n8k4.f64: A[l>>2][l&3], B[l&3][l>>2], C[l>>2][2(l&3)+{0,1}]):
//   apply:  D[c][g] = sum_k UA[c][k] X[k][g],  c = out component (re/im | amp<<1), g = quad slot,
//           two k-steps: k = re of amplitude (l&3), then im of amplitude (l&3), of quad slot (l>>2)
//           -> the B fragments are exactly the lane's own (re, im) of one LDS.128;
//           lane holds D = component c = l>>2 of quad slots 2(l&3), 2(l&3)+1 of the OUTPUT;
//   M:      R[cz][cw] += sum_g Zout[cz][g] Wout[cw][g]: A = the lane's Zout values, B = its Wout
//           values, k-steps = the two quad slots a lane holds -> no shuffles, no re-reads.
//   (the post kernel pulls M_out = U M_in U^H back through the stage).
//
// Two consecutive stages on disjoint bit pairs are fused into one step (see build_dense_tables and
// dense_pass_kernel): the first stage's output fragments become the second stage's B fragments with
// one lane ^ 4 exchange, so a shared-memory round trip and a CTA barrier serve two stages.
//
// Shared memory holds the tile with an XOR swizzle, slot(i) = i ^ fold3(i >> 3), so that the LDS.128
// of a quarter warp and the STS.64 of a half warp are bank-conflict free whenever the stage's bits
// (q, p) and the three quad-slot bits (r0, r1, r2) chosen by the host have suitable residues mod 3
// (always possible for adjacent-qubit pairs in an 11-bit tile); other pairs only lose bandwidth.
#pragma once

constexpr int kDThreads = 256;
constexpr int kDWarps = kDThreads / 32;
constexpr int kDMinTileBits = 5;  // 8 quads of 4 amplitudes per DMMA

__host__ __device__ __forceinline__ unsigned dense_swz(unsigned i) {
  return i ^ (((i >> 3) ^ (i >> 6) ^ (i >> 9)) & 7u);
}

// Per stage, per (warp, lane): everything the inner loop needs, packed into 8 bytes.
//   sl : swizzled slot offset (16-byte units) of the lane's amplitude in the load layout
//   so0, so1 : swizzled offsets (8-byte units) of the lane's two stores
//   sb : swizzled tile-local base of iteration `warp + kDWarps * lane` (lanes < iterations / warp)
struct alignas(8) DLane {
  uint16_t sl, so0, so1, sb;
};
static_assert(sizeof(DLane) == 8, "DLane layout");

struct DenseTables {
  std::vector<DLane> lanes;      // [stage][warp][lane]
  std::vector<int32_t> rbits;    // [stage][3]  quad-slot bits r0, r1, r2 (introspection)
  DLane* d_lanes = nullptr;
};

// rank of three 3-bit vectors over GF(2)
static int gf2_rank3(unsigned a, unsigned b, unsigned c) {
  unsigned v[3] = {a, b, c};
  int rank = 0;
  for (int bit = 0; bit < 3; ++bit) {
    int piv = -1;
    for (int i = rank; i < 3; ++i)
      if (v[i] >> bit & 1u) {
        piv = i;
        break;
      }
    if (piv < 0) continue;
    std::swap(v[rank], v[piv]);
    for (int i = 0; i < 3; ++i)
      if (i != rank && (v[i] >> bit & 1u)) v[i] ^= v[rank];
    ++rank;
  }
  return rank;
}

// Builds the per-stage lane tables of one program (all passes).  Returns false if a pass has a
// tile of fewer than kDMinTileBits bits.
// Flags packed into DLane::sl (slot offsets need 12 bits):
//   kPairFirst: this stage and the next one act on DISJOINT bit pairs and run as one fused step --
//               one LDS.128 and two STS.64 per vector for both stages (see dense_pass_kernel);
//   kPairSwap : the fused step runs the NEXT stage first (the two commute; chosen for bank conflicts).
constexpr uint16_t kPairFirst = 0x8000, kPairSwap = 0x4000;

static bool build_dense_tables(const Program& prog, DenseTables& T, int fuse_pairs = 0) {
  T.lanes.assign(prog.stages.size() * kDWarps * 32, DLane{0, 0, 0, 0});
  T.rbits.assign(prog.stages.size() * 3, 0);
  for (const PassDesc& pd : prog.passes) {
    const int tb = pd.tb;
    if (tb < kDMinTileBits) return false;
    const int nit = 1 << (tb - 5);
    for (int s = 0; s < pd.nstages; ++s) {
      const StageDesc& sd = prog.stages[pd.stage0 + s];
      const int p = sd.p, q = sd.q;
      auto h = [](int b) { return 1u << (b % 3); };
      // quad-slot bits: maximise the GF(2) ranks that make the loads (q, p, r0) and the stores
      // (q, r1, r2) conflict free; ties -> lowest bits
      int best = -1, r0 = -1, r1 = -1, r2 = -1;
      for (int a = 0; a < tb; ++a) {
        if (a == p || a == q) continue;
        for (int b = 0; b < tb; ++b) {
          if (b == p || b == q || b == a) continue;
          for (int c = b + 1; c < tb; ++c) {
            if (c == p || c == q || c == a) continue;
            const int score = 4 * gf2_rank3(h(q), h(p), h(a)) + 4 * gf2_rank3(h(q), h(b), h(c));
            if (score > best) best = score, r0 = a, r1 = b, r2 = c;
          }
        }
      }
      T.rbits[(pd.stage0 + s) * 3 + 0] = r0;
      T.rbits[(pd.stage0 + s) * 3 + 1] = r1;
      T.rbits[(pd.stage0 + s) * 3 + 2] = r2;
      std::vector<int> outer;
      for (int b = 0; b < tb; ++b)
        if (b != p && b != q && b != r0 && b != r1 && b != r2) outer.push_back(b);
      auto base_of = [&](int it) {
        unsigned idx = 0;
        for (size_t k = 0; k < outer.size(); ++k) idx |= (unsigned)((it >> k) & 1) << outer[k];
        return idx;
      };
      for (int w = 0; w < kDWarps; ++w)
        for (int l = 0; l < 32; ++l) {
          DLane& d = T.lanes[((size_t)(pd.stage0 + s) * kDWarps + w) * 32 + l];
          {  // load layout: quad slot g = l >> 2, amplitude a = l & 3
            const int g = l >> 2, a = l & 3;
            const unsigned idx = (unsigned)(a & 1) << q | (unsigned)(a >> 1) << p |
                                 (unsigned)(g & 1) << r0 | (unsigned)((g >> 1) & 1) << r1 |
                                 (unsigned)(g >> 2) << r2;
            d.sl = (uint16_t)dense_swz(idx);
          }
          for (int i = 0; i < 2; ++i) {  // store layout: component c = l >> 2, quad slots 2t + i
            const int c = l >> 2, t = l & 3;
            const int reim = c & 1, a0 = (c >> 1) & 1, a1 = c >> 2;
            const unsigned idx = (unsigned)a0 << q | (unsigned)a1 << p | (unsigned)i << r0 |
                                 (unsigned)(t & 1) << r1 | (unsigned)(t >> 1) << r2;
            const uint16_t v = (uint16_t)(2u * dense_swz(idx) + (unsigned)reim);
            if (i == 0)
              d.so0 = v;
            else
              d.so1 = v;
          }
          const int it = w + kDWarps * l;
          d.sb = (uint16_t)(it < nit ? dense_swz(base_of(it)) : 0);
        }
    }
    if (!fuse_pairs) continue;
    // fused steps: consecutive stages on disjoint bit pairs (p1, q1), (p2, q2).  The first stage
    // loads with quad-slot bits (r0, q2, p2); after its DMMAs a lane holds one real component of
    // the slots (i = r0 bit, t = (q2, p2) bits).  One exchange with lane ^ 4 trades the re/im lane
    // bit for the r0 register bit, which turns the data into the B fragment of the second stage
    // (k = lane & 3 <-> (q2, p2), n = lane >> 2 <-> (r0, q1, p1)) without touching shared memory.
    for (int s = 0; s + 1 < pd.nstages; ++s) {
      const StageDesc& s1 = prog.stages[pd.stage0 + s];
      const StageDesc& s2 = prog.stages[pd.stage0 + s + 1];
      if (s1.p == s2.p || s1.p == s2.q || s1.q == s2.p || s1.q == s2.q) continue;
      auto h = [](int b) { return 1u << (b % 3); };
      // order: stores of the fused step are indexed by (second.q, first.q, first.p)
      const int sc_fwd = gf2_rank3(h(s2.q), h(s1.q), h(s1.p)), sc_rev = gf2_rank3(h(s1.q), h(s2.q), h(s2.p));
      const bool swap = sc_rev > sc_fwd;
      if (fuse_pairs == 2 && std::max(sc_fwd, sc_rev) < 3) continue;  // only bank-conflict-free fusions
      const StageDesc& fa = swap ? s2 : s1;  // runs first
      const StageDesc& fb = swap ? s1 : s2;
      const int q1 = fa.q, p1 = fa.p, q2 = fb.q, p2 = fb.p;
      int best = -1, r0 = -1;
      for (int a = 0; a < tb; ++a) {
        if (a == p1 || a == q1 || a == p2 || a == q2) continue;
        const int score = gf2_rank3(h(q1), h(p1), h(a));
        if (score > best) best = score, r0 = a;
      }
      if (r0 < 0) continue;  // tb == 4 cannot happen (kDMinTileBits = 5)
      std::vector<int> outer;
      for (int b = 0; b < tb; ++b)
        if (b != p1 && b != q1 && b != p2 && b != q2 && b != r0) outer.push_back(b);
      auto base_of = [&](int it) {
        unsigned idx = 0;
        for (size_t k = 0; k < outer.size(); ++k) idx |= (unsigned)((it >> k) & 1) << outer[k];
        return idx;
      };
      for (int w = 0; w < kDWarps; ++w)
        for (int l = 0; l < 32; ++l) {
          DLane& d = T.lanes[((size_t)(pd.stage0 + s) * kDWarps + w) * 32 + l];
          {
            const int g = l >> 2, a = l & 3;
            const unsigned idx = (unsigned)(a & 1) << q1 | (unsigned)(a >> 1) << p1 | (unsigned)(g & 1) << r0 |
                                 (unsigned)((g >> 1) & 1) << q2 | (unsigned)(g >> 2) << p2;
            d.sl = (uint16_t)(dense_swz(idx) | kPairFirst | (swap ? kPairSwap : 0));
          }
          for (int i = 0; i < 2; ++i) {  // output of the second stage: component c = l >> 2, slots n = 2t + i
            const int c = l >> 2, t = l & 3;
            const int reim = c & 1, a0 = (c >> 1) & 1, a1 = c >> 2;
            const unsigned idx = (unsigned)a0 << q2 | (unsigned)a1 << p2 | (unsigned)i << r0 |
                                 (unsigned)(t & 1) << q1 | (unsigned)(t >> 1) << p1;
            const uint16_t v = (uint16_t)(2u * dense_swz(idx) + (unsigned)reim);
            if (i == 0)
              d.so0 = v;
            else
              d.so1 = v;
          }
          const int it = w + kDWarps * l;
          d.sb = (uint16_t)(it < nit ? dense_swz(base_of(it)) : 0);
        }
      ++s;  // the partner stage is consumed by this step
    }
  }
  return true;
}

// ------------------------------------------------------------------------------------------------
// Prologue and epilogue of a sweep (one small launch each, no host copies in between):
//   sweep_prologue_kernel : stage matrices U_s(theta) of one program straight from the angles (the host
//                           writes them into pinned, device-mapped memory -- no H2D copy, no (cos, sin)
//                           table), and the zeroing of the gradient accumulators;
//   grad_epilogue_kernel  : per-rotation inner products from the accumulated stage matrices, the
//                           0.5 / 0.5j / -i factors of the reference (core_operations.py:317-351, 972-975)
//                           and the write of the finished complex gradient into pinned host memory.
// ------------------------------------------------------------------------------------------------
// One gate unit applied to NVEC amplitude quadruples with the (cos, sin) pairs of its angles taken from
// `tr` (indexed like thetas; half angles, full angle for the CPhase parameter).
template <int ENT, bool DAG, int NVEC>
__device__ __forceinline__ void unit_from_trig(const UnitDesc& u, const double2* __restrict__ tr, cd (&a)[NVEC][4],
                                               double* acc) {
  const double2* t = tr + u.theta;
  switch (u.kind) {
    case U_FRONT_LO: front_unit<NVEC, false, DAG>(a, t, acc); break;
    case U_FRONT_HI: front_unit<NVEC, true, DAG>(a, t, acc); break;
    case U_BLOCK_CHI: block_unit<NVEC, ENT, true, DAG>(a, t, u.flags, acc); break;
    case U_BLOCK_CLO: block_unit<NVEC, ENT, false, DAG>(a, t, u.flags, acc); break;
    default: break;
  }
}

// Programmatic dependent launch (the kernels of one evaluation are a chain in one stream): a kernel lets its
// successor be scheduled at once, and waits for its predecessor's memory before it touches anything a
// kernel wrote or still reads -- the successor's launch latency and index arithmetic hide behind the
// predecessor's tail.  Both are no-ops in a launch without the attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

struct PrologueArgs {
  const StageDesc* stages;
  int nstages, nthetas, batch, n3, tpb;
  const double2* gtrig;  // != nullptr: (cos, sin) table in device memory (circuits with > 12 800 angles)
  double2* trig_out;     // != nullptr: the table is also written here (gradient sweep: read by the epilogue)
  const double* thetas;  // [batch][nthetas], pinned host memory mapped into the device address space
  double* umat;          // [batch][nstages][64]
  double* zero0;         // two arrays to clear (stage-matrix sums, per-angle sums); may be null
  long long nzero0;
  double* zero1;
  long long nzero1;
};

// (cos, sin) of every angle of the block's batch element, built cooperatively in shared memory: the angles
// live in pinned HOST memory (every access is a PCIe round trip) and a sincos costs far more than the gate
// arithmetic it feeds, so neither is left to the serial per-unit recipes (a first version that did took
// 14 us per prologue and 28 us per epilogue launch).
__device__ __forceinline__ void build_trig_smem(const double* __restrict__ host_thetas, int nthetas, int n3, int tpb,
                                                double2* s_trig, double2* gtrig_out) {
  // eight loads per thread in flight before the first sincos: one PCIe latency per 1024 angles, not one
  // per angle (a plain loop cost ~9 us for 468 angles)
  for (int k0 = 0; k0 < nthetas; k0 += 8 * blockDim.x) {
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int k = k0 + u * blockDim.x + threadIdx.x;
      v[u] = k < nthetas ? host_thetas[k] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int k = k0 + u * blockDim.x + threadIdx.x;
      if (k < nthetas) {
        const bool full = (tpb == 5) && k >= n3 && ((k - n3) % 5 == 4);  // CPhase parameter
        double sn, cs;
        sincos(full ? v[u] : 0.5 * v[u], &sn, &cs);
        const double2 t = make_double2(cs, sn);
        s_trig[k] = t;
        if (gtrig_out) gtrig_out[k] = t;  // kept for the epilogue of the same sweep
      }
    }
  }
  __syncthreads();
}

template <int ENT, bool DAG>
__device__ __forceinline__ void sweep_prologue_body(const PrologueArgs& A, const double* thetas) {
  extern __shared__ double2 s_trig_buf[];
  const int b = blockIdx.y;
  const double2* s_trig = A.gtrig ? A.gtrig + (size_t)b * A.nthetas : s_trig_buf;
  pdl_launch_dependents();
  // (the angles come from the host -- pinned memory -- so the table is built BEFORE the wait; everything
  // below writes memory the previous evaluation's kernels read.  Angles inside the launch parameters were
  // measured too: a 4 KiB parameter block costs more at launch than the PCIe reads it saves.)
  if (!A.gtrig) build_trig_smem(thetas + (size_t)b * A.nthetas, A.nthetas, A.n3, A.tpb, s_trig_buf, nullptr);
  pdl_wait();
  if (!A.gtrig && A.trig_out && blockIdx.x == 0)  // kept for the epilogue of the same sweep
    for (int k = threadIdx.x; k < A.nthetas; k += blockDim.x) A.trig_out[(size_t)b * A.nthetas + k] = s_trig_buf[k];
  const long long gt = ((long long)blockIdx.y * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
  const long long gsz = (long long)gridDim.x * gridDim.y * blockDim.x;
  for (long long i = gt; i < A.nzero0; i += gsz) A.zero0[i] = 0.0;
  for (long long i = gt; i < A.nzero1; i += gsz) A.zero1[i] = 0.0;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < A.nstages * 4; t += gridDim.x * blockDim.x) {
    const int k = t & 3, s = t >> 2;
    const StageDesc sd = A.stages[s];
    cd a[1][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[0][i].x = (i == k) ? 1.0 : 0.0, a[0][i].y = 0.0;
    for (int u = 0; u < sd.nunits; ++u) unit_from_trig<ENT, DAG, 1>(sd.u[u], s_trig, a, nullptr);
    double* um = A.umat + ((size_t)b * A.nstages + s) * 64;
#pragma unroll
    for (int i = 0; i < 4; ++i) {  // U[i][k] = a[0][i]; DMMA A-fragment order: lane = c * 4 + k, c = reim | amp << 1
      um[(2 * i) * 4 + k] = a[0][i].x;
      um[32 + (2 * i) * 4 + k] = -a[0][i].y;
      um[(2 * i + 1) * 4 + k] = a[0][i].y;
      um[32 + (2 * i + 1) * 4 + k] = a[0][i].x;
    }
  }
}

template <int ENT, bool DAG>
__global__ void __launch_bounds__(128) sweep_prologue_kernel(const PrologueArgs A) {
  sweep_prologue_body<ENT, DAG>(A, A.thetas);
}

struct EpilogueArgs {
  const StageDesc* stages;
  int nstages, nthetas, batch, n3, tpb;
  int trig_smem;           // 1: the (cos, sin) table is staged in shared memory first (it fits the launch's window)
  const double2* gtrig;    // [batch][nthetas] (cos, sin) table in device memory, left by the sweep's prologue
  const double* gm;        // [batch][nstages][64] accumulated stage matrices
  double* out;             // [batch][nthetas] complex gradient 0.5j <P w|z>, pinned host memory
  // Second-order Trotter circuits end with a half layer that REUSES the angles of the first one
  // (parametric_circuit.py:326-336): units of that trailing layer (occurrence number >= extra_seq) store
  // into out2, a second [batch][nthetas] array, and the host adds the two
  double* out2;
  int extra_seq;
};

template <int ENT>
__global__ void __launch_bounds__(128) grad_epilogue_kernel(const EpilogueArgs A) {
  extern __shared__ double2 s_trig_e[];
  const int b = blockIdx.y;
  const double2* gt = A.gtrig + (size_t)b * A.nthetas;  // written by the prologue of this sweep
  // one thread per (stage, unit, virtual quadruple r): w' = e_r, z' = M_out[:, r]; pull both back through
  // the units behind unit u0 and through u0 itself, then run u0 forward with the reference's gate-by-gate
  // accumulation (a chain of at most nunits + 1 recipes per thread instead of 2 nunits)
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = t & 3, s = (t >> 2) / kStageUnits, u0 = (t >> 2) % kStageUnits;
  pdl_launch_dependents();
  const bool live = s < A.nstages && u0 < A.stages[s].nunits;
  pdl_wait();
  // the stage matrix and the table are requested together: one L2 round trip in front of the chain
  cd a[2][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) a[0][i].x = (i == r) ? 1.0 : 0.0, a[0][i].y = 0.0, a[1][i].x = a[1][i].y = 0.0;
  if (live) {
    const double* Mq = A.gm + ((size_t)b * A.nstages + s) * 64;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      a[1][i].x = __ldcg(Mq + ((i << 3) | r));
      a[1][i].y = __ldcg(Mq + ((i << 3) | 4 | r));
    }
  }
  const double2* s_trig = gt;
  if (A.trig_smem) {
    for (int k = threadIdx.x; k < A.nthetas; k += blockDim.x) s_trig_e[k] = gt[k];
    __syncthreads();
    s_trig = s_trig_e;
  }
  constexpr int NACC = (ENT == AQC_ENT_CP) ? 16 : 8;
  double acc[NACC];
#pragma unroll
  for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
  int theta0 = 0, nval = 0;
  double* out = A.out;
  if (live) {
    const StageDesc& sd = A.stages[s];
    for (int u = sd.nunits - 1; u >= u0; --u) unit_from_trig<ENT, true, 2>(sd.u[u], s_trig, a, acc);
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
    const UnitDesc ud = sd.u[u0];
    unit_from_trig<ENT, false, 2>(ud, s_trig, a, acc);
    const int kind = ud.kind;
    nval = (kind == U_FRONT_LO || kind == U_FRONT_HI) ? 6 : ((kind == U_NONE) ? 0 : (ENT == AQC_ENT_CP ? 10 : 8));
    theta0 = ud.theta;
    if (ud.slot >= 5 * A.extra_seq) out = A.out2;
  }
  // Every angle belongs to exactly ONE unit occurrence per output array (checked when the program is built),
  // so its inner product is the sum over the unit's four r lanes -- adjacent lanes of one warp: two
  // shuffles, no atomics, no second phase.  Lane r converts and stores angles r and r + 4 of the unit with the
  // reference's 0.5 / 0.5j / -i factors (core_operations.py:317-351, 972-975), straight into pinned memory.
#pragma unroll
  for (int k = 0; k < NACC; ++k) {
    acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 1);
    acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 2);
  }
#pragma unroll
  for (int j = 0; j < NACC / 2; ++j) {
    if ((j & 3) != r || 2 * j >= nval) continue;
    const int k = theta0 + j;
    const double re = acc[2 * j], im = acc[2 * j + 1];
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
    reinterpret_cast<double2*>(out)[(size_t)b * A.nthetas + k] = v;
  }
}

// hs[b][i] = v[b][idx[i]] written straight into pinned host memory
__global__ void gather_out_kernel(const double2* __restrict__ v, long long stride, const long long* __restrict__ idx,
                                  int count, double2* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  out[(size_t)blockIdx.y * count + i] = v[(long long)blockIdx.y * stride + idx[i]];
}

// ------------------------------------------------------------------------------------------------
// the sweep kernel
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double& d0, double& d1, const double a, const double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

struct DensePassArgs {
  const double2* src[2];  // [0] = w (NVEC == 2) or the single vector; [1] = z
  double2* dst[2];
  long long vec_stride;
  long long basis_index;  // >= 0: src[0] is the basis state |basis_index> (no load)
  const DLane* lanes;     // program-wide [stage][warp][lane]
  const double* umat;     // [batch][nstages_total][64]
  double* gm;             // [batch][nstages_total][64]: 32 used, (az << 3 | reim << 2 | aw)  (NVEC == 2)
  int nstages_total;
  PassDesc pd;
  // Layout switch of a sharded state FUSED into this pass (xchg_world > 0; the last pass of an epoch):
  // instead of writing its tile back in place, the pass stores every 256-byte run where it belongs
  // after the block transpose over the ranks -- element (chunk c, offset o) of this rank goes to
  // (chunk xchg_rank, offset o) of rank c -- straight into the peers' HBM (NVLink peer stores;
  // xdst[v][c] = destination vector v of rank c, mapped with CUDA IPC; c = xchg_rank is local).
  int xchg_world, xchg_rank, xchg_shift;  // chunk = local index >> xchg_shift
  // XORed into the tile number: rank r visits the tiles of chunk (t ^ r) while the grid works through
  // tile group t, so at any moment every rank stores to a DIFFERENT peer (a permutation instead of
  // all ranks hitting the same destination's NVLink ingress at once)
  unsigned tile_xor;
  double2* xdst[2][16];
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// shared-memory accesses with explicit 32-bit shared-window addresses (no generic-address
// arithmetic inside the stage loop)
__device__ __forceinline__ double2 lds128(unsigned addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts64(unsigned addr, double v) {
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ void cp_async16_s(unsigned saddr, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(gsrc) : "memory");
}

// Tile <-> global memory with the addresses of element l = tid + 256 j built from per-thread
// constants: swizzled slot = ((tid ^ m(tid)) ^ c_j) + 256 j, c_j = ((j & 1) << 2) ^ (j >> 1), and
// global offset = lo_off + sum of the strides of the set bits of j (compile-time j).
template <int NJ, bool LOAD>
__device__ __forceinline__ void tile_copy(unsigned sbase, unsigned s0x16, const double2* g0,
                                          double2* gd0, const long long (&hs)[4]) {
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const unsigned cj = (unsigned)((((j & 1) << 2) ^ (j >> 1)) & 7);
    const unsigned sa = sbase + (s0x16 ^ (cj << 4)) + (unsigned)j * 4096u;
    long long off = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if ((j >> k) & 1) off += hs[k];
    if (LOAD) {
      cp_async16_s(sa, g0 + off);
    } else {
      gd0[off] = lds128(sa);
    }
  }
}

// (A 64-register instantiation with 4 CTAs per SM for tiles of <= 2^10 amplitudes was measured in r02: no
// difference at n = 12 ... 22 -- occupancy is not what limits the small-tile passes.)
// MINB = CTAs per SM the register budget is sized for.  The single-vector passes keep four iterations in
// flight per warp (their top stall is the LDS -> DMMA latency, short_scoreboard): 2^12 tiles are limited to
// three CTAs per SM by shared memory anyway (MINB = 3), smaller tiles run with four (64 registers; a fifth
// resident CTA with two iterations in flight was slower: n = 20 V^H sweep 0.115 -> 0.104 ms).
template <int NVEC, int MINB>
__global__ void __launch_bounds__(kDThreads, MINB) dense_pass_kernel(const DensePassArgs A) {
  constexpr int UNR = (NVEC == 1) ? 4 : 2;
  extern __shared__ double2 smem[];
  __shared__ long long s_hioff[16];
  __shared__ double s_mpart[(NVEC == 2) ? 2 * 2 * kDWarps * 32 : 2];  // [parity][set][warp][32]
  __shared__ double2* s_xdst[32];                                       // [vector][rank] push destinations
  const int tid = threadIdx.x;
  pdl_launch_dependents();
  if (A.xchg_world > 0 && tid < 32) s_xdst[tid] = A.xdst[tid >> 4][tid & 15];  // (read after the stage barriers)
  const int lane = tid & 31, warp = tid >> 5;
  const int tb = A.pd.tb;
  const int tsize = 1 << tb;
  const unsigned sm_u32 = (unsigned)__cvta_generic_to_shared(smem);
  const unsigned vbytes = (unsigned)tsize * 16u;  // bytes of one vector's tile

  long long base = 0;
  {
    const unsigned long long tile = blockIdx.x ^ A.tile_xor;
    for (int k = 0; k < A.pd.nouter; ++k)
      base |= (long long)((tile >> k) & 1ull) << A.pd.outerpos[k];
  }
  // local index l = tid + 256 * j  ->  global offset lo_off(tid) | hi_off(j)
  long long lo_off = 0;
  for (int k = 0; k < 8 && k < tb; ++k) lo_off |= (long long)((tid >> k) & 1) << A.pd.bitpos[k];
  long long hs[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) hs[k] = (8 + k < tb) ? (1ll << A.pd.bitpos[8 + k]) : 0ll;
  const bool fast = (tb == 10) || (tb == 11) || (NVEC == 1 && tb == 12);
  if (!fast) {
    if (tid < 16) {
      long long h = 0;
      for (int k = 8; k < tb; ++k) h |= (long long)((tid >> (k - 8)) & 1) << A.pd.bitpos[k];
      s_hioff[tid] = h;
    }
    __syncthreads();
  }
  // (the first step's constants are requested before the tile load so that their latency hides behind it)
  const int nstages = A.pd.nstages;
  const int nit = tsize >> 5;
  const size_t sbase = (size_t)blockIdx.y * A.nstages_total + A.pd.stage0;
  const double* __restrict__ um = A.umat + sbase * 64 + lane;
  const uint2* __restrict__ lt =
      reinterpret_cast<const uint2*>(A.lanes + ((size_t)A.pd.stage0 * kDWarps + warp) * 32 + lane);
  double* __restrict__ gmp = A.gm + sbase * 64 + lane;

  // one-step-ahead prefetch of the per-stage constants: the lane-table entry of the next step and the
  // stage matrices of its (up to) two stages; which of the two runs first is decided by the entry's
  // kPairSwap flag once it has arrived, so no load waits for another one
  double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;  // stage matrices of stages s and s + 1
  uint2 dl = make_uint2(0u, 0u);
  pdl_wait();  // everything above is index arithmetic; from here on the pass reads what its predecessors wrote
  if (nstages > 0) {
    dl = lt[0];
    c0 = um[0];
    c1 = um[32];
    if (nstages > 1) {
      e0 = um[64];
      e1 = um[96];
    }
  }
  const long long boff = (long long)blockIdx.y * A.vec_stride + base;
  const unsigned s0x16 = (unsigned)(tid ^ (((tid >> 3) ^ (tid >> 6)) & 7)) << 4;

#pragma unroll
  for (int v = 0; v < NVEC; ++v) {
    double2* sm = smem + (size_t)v * tsize;
    if (v == 0 && A.basis_index >= 0) {
      for (int l = tid; l < tsize; l += kDThreads) {
        long long h = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (((l >> 8) >> k) & 1) h |= hs[k];
        const long long g = base | lo_off | h;
        sm[dense_swz(l)] = make_double2(g == A.basis_index ? 1.0 : 0.0, 0.0);
      }
    } else {
      // the whole tile goes in flight at once (16-byte LDGSTS straight into the swizzled slots):
      // the load phase costs one memory round trip instead of one per unrolled batch of LDGs
      const double2* __restrict__ src = A.src[v] + boff;
      if (tb == 11) {
        tile_copy<8, true>(sm_u32 + v * vbytes, s0x16, src + lo_off, nullptr, hs);
      } else if (tb == 10) {
        tile_copy<4, true>(sm_u32 + v * vbytes, s0x16, src + lo_off, nullptr, hs);
      } else if (NVEC == 1 && tb == 12) {
        tile_copy<16, true>(sm_u32 + v * vbytes, s0x16, src + lo_off, nullptr, hs);
      } else {
        for (int l = tid; l < tsize; l += kDThreads)
          cp_async16(sm + dense_swz(l), src + (lo_off | s_hioff[l >> 8]));
      }
    }
  }
  cp_async_commit();
  cp_async_wait_all();
  __syncthreads();

  int step = 0;
  for (int s = 0; s < nstages; ++step) {
    const bool paired = (dl.x & kPairFirst) != 0;
    const bool swapped = (dl.x & kPairSwap) != 0;
    const int sA = s + (swapped ? 1 : 0);  // stage that runs first
    const int sB = 2 * s + 1 - sA;         // its partner (paired steps only)
    const int snext = s + (paired ? 2 : 1);
    const double ua0 = swapped ? e0 : c0, ua1 = swapped ? e1 : c1;
    const double ub0 = swapped ? c0 : e0, ub1 = swapped ? c1 : e1;
    double nc0 = 0.0, nc1 = 0.0, ne0 = 0.0, ne1 = 0.0;
    uint2 nl = make_uint2(0u, 0u);
    if (snext < nstages) {
      nl = lt[(size_t)snext * kDWarps * 32];
      if (paired) {  // an unpaired step already holds the next stage's matrix in (e0, e1)
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
    // byte offsets inside one vector's tile: load slot (16-byte units), store slots (8-byte units)
    const unsigned sl16 = (dl.x & 0x0fffu) << 4, so0 = (dl.x >> 16) << 3, so1 = (dl.y & 0xffffu) << 3;
    const unsigned sb16 = (dl.y >> 16) << 4;
    const bool hi = (lane & 4) != 0;  // lane holds Im (first-stage output) / r0 = 1 (second-stage input)
    double m0 = 0.0, m1 = 0.0, n0 = 0.0, n1 = 0.0;
    int j = 0;
    if (!paired) {
#pragma unroll UNR
      for (int it = warp; it < nit; it += kDWarps, ++j) {
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
      // fused step: stage A, lane ^ 4 exchange (re/im lane bit <-> r0 register bit), stage B
#pragma unroll UNR
      for (int it = warp; it < nit; it += kDWarps, ++j) {
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
      // The warp's partial R[cz = lane >> 2][cw = 2 (lane & 3) + {0, 1}] (8x8 real cross products)
      // folds into the 4x4 complex M = sum z w^H with one exchange between the lanes holding Re z
      // and Im z (lane ^ 4): lane (az, 0, aw) ends with Re M[az][aw], lane (az, 1, aw) with Im.
      double* part = s_mpart + (size_t)(step & 1) * 2 * kDWarps * 32;  // [set][warp][32]
      const double recv = __shfl_xor_sync(0xffffffffu, m1, 4);
      part[warp * 32 + lane] = hi ? (m0 - recv) : (m0 + recv);
      if (paired) {
        const double recv2 = __shfl_xor_sync(0xffffffffu, n1, 4);
        part[(kDWarps + warp) * 32 + lane] = hi ? (n0 - recv2) : (n0 + recv2);
      }
      __syncthreads();
      const int red0 = step & (kDWarps - 1), red1 = (step + kDWarps / 2) & (kDWarps - 1);
      if (warp == red0 || (paired && warp == red1)) {
        const int set = (warp == red0) ? 0 : 1;
        const double* pp = part + (size_t)set * kDWarps * 32;
        double r0 = 0.0;
#pragma unroll
        for (int w = 0; w < kDWarps; ++w) r0 += pp[w * 32 + lane];
        atomicAdd(gmp + (size_t)(set == 0 ? sA : sB) * 64, r0);
      }
    } else {
      __syncthreads();
    }
    c0 = nc0, c1 = nc1, e0 = ne0, e1 = ne1, dl = nl;
    s = snext;
  }

  if (A.xchg_world > 0) {
    // fused layout switch: the tile leaves for the ranks it belongs to (runs of 2^low >= 16 amplitudes stay
    // contiguous: the low tile bits are low index bits, the chunk bits are the top ones)
    if (nstages == 0) __syncthreads();
    const long long omask = (1ll << A.xchg_shift) - 1;
    const long long rbase = (long long)A.xchg_rank << A.xchg_shift;
#pragma unroll
    for (int v = 0; v < NVEC; ++v) {
      const double2* sm = smem + (size_t)v * tsize;
      for (int l = tid; l < tsize; l += kDThreads) {
        long long li = base | lo_off;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (((l >> 8) >> k) & 1) li |= hs[k];
        double2* d = s_xdst[v * 16 + (int)(li >> A.xchg_shift)] + (rbase | (li & omask));
        *d = sm[dense_swz(l)];
      }
    }
    return;
  }
#pragma unroll
  for (int v = 0; v < NVEC; ++v) {
    const double2* sm = smem + (size_t)v * tsize;
    double2* __restrict__ dst = A.dst[v] + boff;
    if (tb == 11) {
      tile_copy<8, false>(sm_u32 + v * vbytes, s0x16, nullptr, dst + lo_off, hs);
    } else if (tb == 10) {
      tile_copy<4, false>(sm_u32 + v * vbytes, s0x16, nullptr, dst + lo_off, hs);
    } else if (NVEC == 1 && tb == 12) {
      tile_copy<16, false>(sm_u32 + v * vbytes, s0x16, nullptr, dst + lo_off, hs);
    } else {
      for (int l = tid; l < tsize; l += kDThreads) dst[lo_off | s_hioff[l >> 8]] = sm[dense_swz(l)];
    }
  }
}
