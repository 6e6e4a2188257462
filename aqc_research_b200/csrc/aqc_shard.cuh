// aqc_shard.cuh -- global-qubit sharding: epoch-wise execution, peer mapping and the layout switch
// (block transpose over the ranks through NVLink peer memory).  Included at the end of aqc_sv.cu.
#pragma once
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
  rc = upload_thetas(sv, thetas, !sv->dense);
  if (rc) return rc;
  if (sv->dense) {
    if ((rc = dense_prepare(sv, mode))) return rc;  // (clears the gradient sums too)
  } else if (mode == 0) {
    CU(cudaMemsetAsync(sv->d_gacc, 0, tot * 2 * sizeof(double), sv->stream));
  }
  CU(cudaStreamSynchronize(sv->stream));
  return AQC_OK;
}

// Runs the tile passes of one epoch.  mode 0: gradient on (vec0, vec1) = (w, z); mode 1 / 2:
// V / V^H on vec0.  src slots are read by the first pass only (src0 < 0: vec0 is the local part
// of a basis state: offset `basis_local`, or all zeros if basis_local < 0); dst slots receive the
// result and are updated in place by the remaining passes.
// push0 (and push1 for the gradient) >= 0: the layout switch that follows the epoch is FUSED into its
// last pass -- the pass stores its tiles straight into slot push0 / push1 of the rank each 256-byte
// run belongs to after the block transpose (peer stores over NVLink, buffers mapped with
// aqc_sv_ipc_import); dst then only holds intermediate results.  The caller makes all ranks meet
// (host barrier) before anyone reads the pushed slots.
extern "C" int aqc_sv_run_epoch(aqc_sv* sv, int mode, int epoch, int src0, int64_t basis_local,
                                int src1, int dst0, int dst1, int push0, int push1) {
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
  if (push0 >= 0) {
    if (sv->g <= 0 || !sv->dense)
      return fail(AQC_EINVAL, "the fused layout switch needs a sharded workspace on the dense engine");
    if ((rc = check_slot(sv, push0))) return rc;
    if (mode == 0 && (rc = check_slot(sv, push1))) return rc;
    if (push0 == dst0 || push0 == src0 || (mode == 0 && (push1 == dst1 || push1 == dst0 || push0 == dst1 ||
                                                         push1 == src1 || push0 == src1 || push0 == push1)))
      return fail(AQC_EINVAL, "pushed slots must differ from the slots the epoch works on");
  }
  CU(cudaSetDevice(sv->device));
  const int p0 = p->epoch_pass0[epoch];
  const int p1 = epoch + 1 < (int)p->epoch_pass0.size() ? p->epoch_pass0[epoch + 1] : (int)p->passes.size();
  const long long basis = src0 >= 0 ? -1 : (basis_local >= 0 ? (long long)basis_local : (1ll << 62));
  CU(cudaEventRecord(sv->ev0, sv->stream));
  if (sv->dense)
    rc = run_dense_program(sv, mode, src0 >= 0 ? sv->slots[src0] : nullptr, basis,
                           mode == 0 ? sv->slots[src1] : nullptr, sv->slots[dst0],
                           mode == 0 ? sv->slots[dst1] : nullptr, p0, p1, push0, push1);
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

// 1 if this workspace can fuse the layout switch into the last pass of an epoch (aqc_sv_run_epoch).
extern "C" int aqc_sv_can_push(const aqc_sv* sv) { return (sv && sv->g > 0 && sv->dense) ? 1 : 0; }

// Downloads this workspace's (partial) raw inner products and converts them to 0.5j <P w|z>
// (linear, so partial sums of several ranks may be added afterwards).
extern "C" int aqc_sv_grad_finish(aqc_sv* sv, double* grad_out) {
  if (!sv || !grad_out) return fail(AQC_EINVAL, "bad arguments");
  CU(cudaSetDevice(sv->device));
  const size_t tot = (size_t)sv->batch * sv->circ.nthetas;
  int rc = ensure_pinned(sv, tot * 4 + 64);
  if (rc) return rc;
  if (sv->dense) {  // the epilogue kernel converts and writes the (partial) gradient into h_pinned
    if ((rc = dense_collect(sv))) return rc;
    CU(cudaStreamSynchronize(sv->stream));
    dense_gradient_from_pinned(sv, grad_out);
    return AQC_OK;
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
  sv->ipc_opened.push_back(p);  // closed by aqc_sv_ipc_close / aqc_sv_destroy
  return AQC_OK;
}

// Unmaps every peer buffer imported with aqc_sv_ipc_import.  Importers call this (and all ranks meet)
// BEFORE the exporting workspaces are destroyed.
extern "C" int aqc_sv_ipc_close(aqc_sv* sv) {
  if (!sv) return fail(AQC_EINVAL, "null workspace");
  CU(cudaSetDevice(sv->device));
  CU(cudaStreamSynchronize(sv->stream));
  for (void* p : sv->ipc_opened) cudaIpcCloseMemHandle(p);
  sv->ipc_opened.clear();
  for (auto& row : sv->peer)
    for (auto& q : row) q = nullptr;
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
  const int r = (blockIdx.y + A.rank) % A.world;  // every rank starts with another peer (no hot spot)
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
  const int r = (blockIdx.y + A.rank) % A.world;
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
