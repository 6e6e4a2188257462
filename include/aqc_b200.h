/*
 * aqc_b200.h -- C-ABI of the B200-native objective-and-gradient engine.
 *
 * The reference (qiskit-community/aqc-research) is pure Python and has no FFI of
 * its own; its hot path is entered through duck-typed Python calls
 *     objv.objective(thetas) / objv.gradient(thetas)       (optimizer.py:561-590)
 * which bottom out in the NumPy kernels of core_operations.py, core_op_matrix.py
 * and mps_operations.py / mps_dot_objective.py.  Every entry point below names the
 * reference function (file:line, relative to the reference root) it replaces.  A
 * ctypes binding that a reference maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *  - all functions return 0 on success, a negative AQC_E* code otherwise;
 *    aqc_last_error() gives a thread-local human-readable message;
 *  - handles are opaque; host pointers are borrowed for the duration of the call;
 *  - complex numbers are interleaved (re, im) float64 pairs, i.e. numpy complex128;
 *  - qubit q <-> bit q (LSB = qubit 0) of the flat amplitude index
 *    (core_operations.py:34-43,77);
 *  - all calls are synchronous: outputs are valid on return;
 *  - there is NO CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef AQC_B200_H
#define AQC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AQC_OK 0
#define AQC_EINVAL (-1)  /* bad argument */
#define AQC_ECUDA (-2)   /* CUDA runtime error */
#define AQC_ENODEV (-3)  /* no CUDA device */
#define AQC_ENOMEM (-4)  /* allocation failed */

#define AQC_ENT_CX 0
#define AQC_ENT_CZ 1
#define AQC_ENT_CP 2

#define AQC_GENERIC 0        /* ParametricCircuit                       */
#define AQC_TROTTER_1ST 1    /* TrotterAnsatz(second_order=False)       */
#define AQC_TROTTER_2ND 2    /* TrotterAnsatz(second_order=True)        */

typedef struct aqc_circuit aqc_circuit;
typedef struct aqc_sv aqc_sv;
typedef struct aqc_mps aqc_mps;

const char* aqc_last_error(void);
int aqc_version(void);
/* Number of visible CUDA devices (0 if none / no driver). */
int aqc_device_count(void);

/* ---------------------------------------------------------------------------
 * Circuit structure (host only, no angles).
 * Replaces: ParametricCircuit / TrotterAnsatz (parametric_circuit.py:37-70,300-320).
 * blocks: int32[2*num_blocks], row 0 = control qubits, row 1 = target qubits.
 * theta layout: [3n front | num_blocks x tpb], tpb = 5 for cp else 4 (:66,108-112).
 * For AQC_TROTTER_2ND the trailing half-layer of 3*(n/2) blocks re-using the
 * angles of the first half-layer is implied (core_operations.py:638-641,686-690).
 */
int aqc_circuit_create(int num_qubits, int entangler, const int32_t* blocks, int num_blocks,
                       int trotter, aqc_circuit** out);
void aqc_circuit_destroy(aqc_circuit* circ);
int aqc_circuit_num_thetas(const aqc_circuit* circ);

/* ---------------------------------------------------------------------------
 * State-vector / column-batched workspace on one GPU.
 *
 * A workspace owns `num_slots` device arrays ("slots"), each holding `batch`
 * independent states of 2^(n + log2_cols) complex128 amplitudes.  log2_cols = 0 is
 * the vector path (core_operations.py); log2_cols = k is the matrix path
 * (core_op_matrix.py): a row-major (2^n, 2^k) matrix whose gates act on the ROW
 * index bits (core_op_matrix.py:56), i.e. on bits k..k+n-1 of the flat index.
 * `batch` independent angle sets (multistart) are evaluated by one launch sequence.
 */
int aqc_sv_create(const aqc_circuit* circ, int device, int log2_cols, int batch, int num_slots,
                  aqc_sv** out);
void aqc_sv_destroy(aqc_sv* sv);
/* Amplitudes per state = 2^(n + log2_cols). */
int64_t aqc_sv_state_size(const aqc_sv* sv);

/* Host -> slot.  `count` complex numbers are copied into state `batch_index`
 * (batch_index = -1: the same host data is replicated into every batch element). */
int aqc_sv_upload(aqc_sv* sv, int slot, int batch_index, const double* host, int64_t count);
/* Slot -> host. */
int aqc_sv_download(aqc_sv* sv, int slot, int batch_index, double* host, int64_t count);
/* slot[b] = |index>  for all b (ThinStateHandler states, objective_base.py:42-173). */
int aqc_sv_set_basis(aqc_sv* sv, int slot, int64_t index);
/* slot[b] = sum_k (amps[2k] + i amps[2k+1]) |indices[k]>  for all b, 1 <= count <= 8.  Stream
 * ordered, returns without waiting.  The weighted two-term gradient of the surrogate objective
 * (objective_lhs_sur_max.py:150-186) is antilinear in the flip states, so both terms are taken in
 * ONE sweep started from -2(1-w) hs_0 |s_0> - 2 w hs_max |s_max>. */
int aqc_sv_set_sparse(aqc_sv* sv, int slot, const int64_t* indices, const double* amps, int count);
/* slot[b] = identity matrix (requires log2_cols == n), FullRangeSketchingVectors
 * (sk_core.py:317-326). */
int aqc_sv_set_identity(aqc_sv* sv, int slot);
/* Fills a slot with the distribution of utils.rand_state (utils.py:71-79): re, im ~ U[0,1)
 * from a counter-based generator keyed on (seed, amplitude index), then normalised. */
int aqc_sv_fill_random(aqc_sv* sv, int slot, uint64_t seed);
/* out[b*count + i] = slot[b][idx[i]]   (ThinStateHandler.state_dot_vector,
 * objective_base.py:164-173: <e_k|v> = v[k]). */
int aqc_sv_gather(aqc_sv* sv, int slot, const int64_t* idx, int count, double* out);
/* out[b] = <slot_a[b] | slot_b[b]>  (np.vdot; GenericStateHandler.state_dot_vector,
 * objective_base.py:313-316; sk_core.py:192). */
int aqc_sv_vdot(aqc_sv* sv, int slot_a, int slot_b, double* out);

/* dst = V(thetas) src  (dagger = 0; v_mul_vec core_operations.py:606-710,
 *                       v_mul_mat core_op_matrix.py:480-559)
 * dst = V(thetas)^H src (dagger = 1; v_dagger_mul_vec core_operations.py:713-820,
 *                       v_dagger_mul_mat core_op_matrix.py:562-642).
 * thetas: float64[batch * num_thetas].  src_slot may equal dst_slot. */
int aqc_sv_apply(aqc_sv* sv, const double* thetas, int dagger, int src_slot, int dst_slot);

/* Full complex gradient of <V x | y> by all thetas, given z0 = V^H y
 * (grad_of_dot_product core_operations.py:823-1019; grad_of_matrix_dot_product
 * core_op_matrix.py:645-762).  x is slot `x_slot`, or the basis state |x_basis>
 * when x_slot < 0.  w_slot / z_slot are scratch slots for the swept states (their contents
 * afterwards are unspecified: the sweep keeps V x and V z0 only up to a common, tracked scale);
 * they may equal x_slot / z0_slot, which are then destroyed like in the matrix version of the
 * reference.  Second-order Trotter half-layer derivatives are
 * accumulated into the entries of the leading half-layer (core_operations.py:966-994).
 * grad_out: complex128[batch * num_thetas]. */
int aqc_sv_grad(aqc_sv* sv, const double* thetas, int x_slot, int64_t x_basis, int z0_slot,
                int w_slot, int z_slot, double* grad_out);
/* The same sweep split in two: _begin enqueues everything (angles from a pinned staging buffer,
 * kernels, the copy of the raw sums back to the host) and returns at once; _end waits for it and
 * writes grad_out.  Lets a caller that knows the optimiser asks for the gradient right after the
 * objective at the same angles (scipy's L-BFGS-B does, optimizer.py:585-590) start the sweep while
 * the host is still busy.  Any other compute call on the workspace drops an uncollected sweep. */
int aqc_sv_grad_begin(aqc_sv* sv, const double* thetas, int x_slot, int64_t x_basis, int z0_slot,
                      int w_slot, int z_slot);
int aqc_sv_grad_end(aqc_sv* sv, double* grad_out);

/* Fused objective step of the state-preparation objectives
 * (objective_lhs_sur_max.py:98-106): z0_slot = V^H target_slot; hs_out[b*count+i] =
 * z0_slot[b][idx[i]]. */
int aqc_sv_objective(aqc_sv* sv, const double* thetas, int target_slot, int z0_slot,
                     const int64_t* idx, int count, double* hs_out);

/* One evaluation (objective + gradient at the same angles) enqueued in one go -- scipy asks for
 * fun(theta) and then jac(theta) (optimizer.py:585-590): V^H sweep of target_slot into z0_slot, gather of
 * hs (objective_lhs_sur_max.py:98-106), gradient sweep of <V e_x | target> from the basis state x_basis
 * (core_operations.py:823) and its epilogue.  aqc_sv_eval_hs returns hs as soon as the gather has
 * finished (the gradient sweep keeps running); aqc_sv_grad_end collects the gradient.  A gradient that
 * is not collected is dropped by the next call on the workspace. */
int aqc_sv_eval_begin(aqc_sv* sv, const double* thetas, int target_slot, int z0_slot,
                      const int64_t* idx, int count, int64_t x_basis, int w_slot, int z_slot);
int aqc_sv_eval_hs(aqc_sv* sv, double* hs_out);
/* 1 if aqc_sv_eval_begin is available (dense engine, unsharded workspace; states of fewer than 32
 * amplitudes run on the legacy engine), else 0: callers then use aqc_sv_objective + aqc_sv_grad. */
int aqc_sv_can_eval(const aqc_sv* sv);
/* kernel-time split (ms) of the last completed aqc_sv_eval_begin .. aqc_sv_grad_end pair */
int aqc_sv_eval_times(aqc_sv* sv, float* obj_ms, float* grad_ms);

/* Device time (ms, CUDA events on the workspace stream) of the kernels launched by the
 * most recent aqc_sv_apply / aqc_sv_grad / aqc_sv_objective call, and how many
 * kernels that call launched.  Used by bench.py for the roofline figure. */
float aqc_sv_last_kernel_ms(const aqc_sv* sv);
int aqc_sv_last_num_launches(const aqc_sv* sv);
/* CUDA-event stopwatch on the workspace stream; brackets any sequence of calls made on this
 * workspace between start and stop (ms = device time between the two events). */
int aqc_sv_timer_start(aqc_sv* sv);
int aqc_sv_timer_stop(aqc_sv* sv, float* ms);
/* Sketching-vector generators (sk_core.py:329-464) on a matrix workspace (d = 2^n rows,
 * m = 2^log2_cols columns per slot, batch 1).  The dense target U (d x d, row-major complex128)
 * lives in its own device buffer. */
int aqc_sv_set_dense_target(aqc_sv* sv, const double* target);
/* dst = U @ src (conj_transpose == 0) or U^H @ src: replaces np.dot(target, x_vecs) (sk_core.py:358,
 * 461) and the U^H Omega product (:446-448); FP64 tensor-core (DMMA) GEMM. */
int aqc_sv_target_matmul(aqc_sv* sv, int conj_transpose, int src_slot, int dst_slot);
/* slot <- an orthonormal basis of its column space: replaces `x_vecs, _ = np.linalg.qr(...)`
 * (sk_core.py:355-357, 458).  Shifted Cholesky-QR3 (three Gram / Cholesky / triangular-solve
 * rounds, all GEMM shaped); the basis differs from LAPACK's by a unitary m x m factor, which the
 * sketched objective and gradient are invariant to.  Fails for numerically rank-deficient input. */
int aqc_sv_orthonormalize(aqc_sv* sv, int slot, int tmp_slot);
/* dst -= src (sk_core.py:453). */
int aqc_sv_sub(aqc_sv* sv, int dst_slot, int src_slot);
/* X[:, i] = e_idx[i], Y[:, i] = U[:, idx[i]] (AlternatingSketchingVectors, sk_core.py:398-404). */
int aqc_sv_gather_target_columns(aqc_sv* sv, const int64_t* idx, int count, int x_slot, int y_slot);
/* Coordinate descent for unitary AQC.
 * Replaces: coord_descent_single_sweep (core_op_matrix.py:765-917), called num_sweeps times.
 * The workspace must be a matrix workspace with log2_cols == num_qubits (target 2^n x 2^n in
 * target_slot, row-major); cx and cz ParametricCircuits only (the reference raises
 * NotImplementedError for cp).  Every sweep: z = V(thetas)^H target, w = I, then every angle
 * gets one Newton / clipped-gradient update in circuit order (z rotated with the old angle, w with
 * the new one).  thetas: float64[batch * T], updated IN PLACE; fobj_out: float64[num_sweeps * batch],
 * fobj = 1 - |<w|z>_F / 2^n|^2 at the end of each sweep (as returned by the reference). */
int aqc_sv_coord_descent(aqc_sv* sv, double* thetas, int target_slot, int w_slot, int z_slot,
                         int num_sweeps, double* fobj_out);
/* Scheduler introspection: number of tile passes over the state for one gradient
 * sweep (mode 0), one V apply (1) or one V^H apply (2). */
int aqc_sv_num_passes(const aqc_sv* sv, int mode);
/* ... and the number of stages (one 4x4 stage matrix applied to every amplitude quadruple of the
 * state) summed over those passes: the FP64 work of a sweep is proportional to it. */
int aqc_sv_num_stages(const aqc_sv* sv, int mode);
/* Host-only (no device needed): serialises the tile-pass program the scheduler compiles for
 * `circ` as int32 words (layout documented in csrc/aqc_sv.cu); reversed = 1 gives the V^H
 * program.  *needed receives the word count; data are written iff cap is large enough. */
int aqc_debug_program(const aqc_circuit* circ, int log2_cols, int tile_bits, int low_bits,
                      int reversed, int32_t* out, int64_t cap, int64_t* needed);
/* Same for the dense-stage (DMMA) engine: stages hold up to 5 units (front gates merged into
 * the first stage on their qubit pair) and carry the shared-memory lane tables of the sweep
 * kernel (layout documented in csrc/aqc_sv.cu), so CPU tests can emulate its data flow. */
int aqc_debug_dense_program(const aqc_circuit* circ, int log2_cols, int tile_bits, int low_bits,
                            int reversed, int32_t* out, int64_t cap, int64_t* needed);
/* ... with the fused two-stage steps of the production tables (flags in the `sl` word). */
int aqc_debug_dense_program_fused(const aqc_circuit* circ, int log2_cols, int tile_bits, int low_bits,
                            int reversed, int32_t* out, int64_t cap, int64_t* needed);
/* Device-pointer access for zero-copy callers (torch tensors): address of slot. */
void* aqc_sv_slot_ptr(aqc_sv* sv, int slot);
/* CUDA stream handle (cudaStream_t) the workspace launches on. */
void* aqc_sv_stream(aqc_sv* sv);

/* ---------------------------------------------------------------------------
 * One state vector over 2^log2_world GPUs (global-qubit sharding, n >= 32 qubits; one process or
 * workspace per GPU).  The reference has no distributed code; this is new.
 * The top log2_world index bits select the rank; two layouts alternate (A: highest g qubits
 * global, B: lowest g qubits global); an "epoch" runs every gate unit that does not touch a
 * global qubit, then the layout is switched by a block transpose over the ranks
 * (aqc_sv_exchange: NVLink peer loads; or any all-to-all of the caller).  Inner products are
 * partial per rank: add the outputs of aqc_sv_grad_finish / aqc_sv_gather over the ranks.
 * A single-GPU workspace behaves as one epoch in layout A.
 */
int aqc_sv_create_sharded(const aqc_circuit* circ, int device, int log2_world, int rank,
                          int num_slots, aqc_sv** out);
/* mode: 0 gradient sweep, 1 V apply, 2 V^H apply. */
int aqc_sv_num_epochs(const aqc_sv* sv, int mode);
int aqc_sv_epoch_layout(const aqc_sv* sv, int mode, int epoch); /* 0 = A, 1 = B */
int aqc_sv_begin(aqc_sv* sv, const double* thetas, int mode);
/* Tile passes of one epoch (src slots are read by the first pass only; src0 < 0: vec0 = local part
 * of a basis state at offset basis_local, or zeros if basis_local < 0).
 * push0 (and push1 for the gradient) >= 0 FUSES the layout switch that follows the epoch into its
 * last tile pass: the pass stores every 256-byte run of its tiles straight into slot push0 / push1
 * of the rank the run belongs to after the block transpose (peer stores over NVLink; all slots
 * imported with aqc_sv_ipc_import / aqc_sv_peer_attach).  All ranks must meet before anyone reads
 * the pushed slots.  -1: the result stays in dst0 / dst1 in the epoch's own layout. */
int aqc_sv_run_epoch(aqc_sv* sv, int mode, int epoch, int src0, int64_t basis_local, int src1,
                     int dst0, int dst1, int push0, int push1);
/* 1 if aqc_sv_run_epoch can fuse the layout switch (sharded workspace on the dense engine). */
int aqc_sv_can_push(const aqc_sv* sv);
/* This rank's partial complex gradient (already scaled like grad_of_dot_product). */
int aqc_sv_grad_finish(aqc_sv* sv, double* grad_out);
/* Peer mapping of the other ranks' slots: CUDA IPC between processes ... */
int aqc_sv_ipc_export(aqc_sv* sv, int slot, unsigned char* handle64);
int aqc_sv_ipc_import(aqc_sv* sv, int peer_rank, int slot, const unsigned char* handle64);
/* Unmaps everything aqc_sv_ipc_import mapped; importers call it (and all ranks meet) before the
 * exporting workspaces are destroyed. */
int aqc_sv_ipc_close(aqc_sv* sv);
/* ... or direct peer access when all workspaces live in one process. */
int aqc_sv_peer_attach(aqc_sv* sv, int peer_rank, int slot, aqc_sv* peer);
/* dst_slot[chunk r] = (rank r).src_slot[chunk my_rank] for all r: layout switch A <-> B.
 * Callers must barrier before (sources complete) and after (sources free again). */
int aqc_sv_exchange(aqc_sv* sv, int src_slot, int dst_slot);
/* Synthetic target keyed on the LOGICAL amplitude index (layout A), not normalised; returns the
 * local sum of squares. */
int aqc_sv_fill_random_logical(aqc_sv* sv, int slot, uint64_t seed, double* norm2_out);
int aqc_sv_scale(aqc_sv* sv, int slot, double factor);
/* Host-only: per-epoch tile-pass programs of the sharded scheduler (for CPU replay tests). */
int aqc_debug_program_sharded(const aqc_circuit* circ, int log2_world, int tile_bits,
                              int low_bits, int reversed, int32_t* out, int64_t cap,
                              int64_t* needed);

/* ---------------------------------------------------------------------------
 * MPS workspace on one GPU (Vidal form, bond capacity C = aqc_mps_bond_capacity()).
 * Unit-blocks may act on any (ctrl, targ) pair (mps_dot_objective.py:380-468); non-adjacent pairs run
 * through a swap network of adjacent two-site updates.  The production case is
 * SpSurrogateObjectiveFastMpsTrotter (objective_lhs_sur_fast_mps_trotter.py:57-99).  Gate arithmetic that the reference delegates to
 * qiskit-aer (mps_operations.py:248-265) runs here: two-site contraction, one-sided Jacobi SVD,
 * truncation rule "drop the smallest Schmidt values while the sum of their squares < trunc_thr,
 * cap at chi_max, renormalise".
 * Host layout of a state: gam complex128[n][2][C][C] (Gamma_k[b], rows = left bond), lam
 * float64[n+1][C] (bond j sits LEFT of site j; bonds 0 and n hold [1]), dims int32[n+1].
 */
int aqc_mps_create(const aqc_circuit* circ, int device, int chi_max, double trunc_thr,
                   int num_slots, aqc_mps** out);
void aqc_mps_destroy(aqc_mps* mps);
int aqc_mps_bond_capacity(const aqc_mps* mps);
/* QiskitMPS (mps_operations.py:33) -> slot and back. */
int aqc_mps_upload(aqc_mps* mps, int slot, const double* gam, const double* lam,
                   const int32_t* dims);
int aqc_mps_download(aqc_mps* mps, int slot, double* gam, double* lam, int32_t* dims);
/* slot = |index> as a bond-dimension-1 MPS (MpsStateHandler states for X-type preparations,
 * objective_base.py:345-398). */
int aqc_mps_set_product(aqc_mps* mps, int slot, int64_t index);
/* The same product state except that qubit `site` holds (amps[0] + i amps[1]) |0> +
 * (amps[2] + i amps[3]) |1>: a weighted pair of states that differ by one flip is still a
 * product state (the caller keeps the pair at unit norm), so both terms of the surrogate gradient
 * (objective_lhs_sur_fast_mps_trotter.py:190-227) come from one sweep. */
int aqc_mps_set_product_site(aqc_mps* mps, int slot, int64_t index, int site, const double* amps);
/* dst = V src (dagger 0, v_mul_mps mps_operations.py:326-346) or V^H src (dagger 1,
 * v_dagger_mul_mps :349-371), with truncation. */
int aqc_mps_apply(aqc_mps* mps, const double* thetas, int dagger, int src_slot, int dst_slot);
/* out[i] = <idx_i | slot>: MpsStateHandler.state_dot_vector (objective_base.py:400-403) for
 * basis states. */
int aqc_mps_amplitudes(aqc_mps* mps, int slot, const int64_t* idx, int count, double* out);
/* z0_slot = V^H target_slot; hs_out[i] = <idx_i | z0>
 * (objective_lhs_sur_fast_mps_trotter.py:130-140). */
int aqc_mps_objective(aqc_mps* mps, const double* thetas, int target_slot, int z0_slot,
                      const int64_t* idx, int count, double* hs_out);
/* out = <slot_a | slot_b>  (mps_dot, mps_operations.py:192-213). */
int aqc_mps_dot(aqc_mps* mps, int slot_a, int slot_b, double* out);
/* Complex gradient of <V x | y> given z0 = V^H y (fast_dot_gradient,
 * mps_dot_objective.py:41-242).  x = |x_basis> if x_slot < 0. */
int aqc_mps_grad(aqc_mps* mps, const double* thetas, int x_slot, int64_t x_basis, int z0_slot,
                 int w_slot, int z_slot, double* grad_out);
/* Diagnostics: Jacobi sweeps used by the SVDs of the most recent two-qubit step. */
int aqc_mps_debug_sweeps(aqc_mps* mps, int32_t* out, int cap);
float aqc_mps_last_kernel_ms(const aqc_mps* mps);
int aqc_mps_last_num_launches(const aqc_mps* mps);
/* Truncation record of the most recent apply / objective / grad call: out4[0] = sum over all splits
 * of the discarded weight (squared Schmidt values relative to the split), [1] = largest single
 * discard, [2] = the part of [0] removed ONLY by the chi_max cap (qiskit-aer, which the reference
 * calls at mps_operations.py:248-265, has no cap), [3] = number of splits the cap cut. */
int aqc_mps_truncation_stats(aqc_mps* mps, double* out4);

/* ---------------------------------------------------------------------------------------------
 * Single-gate primitives (csrc/aqc_prim.cu): the gate-by-gate functions the reference's unit tests
 * and tools call directly -- core_operations.py:46-603 (gate2x2_mul_vec, proj00/11_mul_vec,
 * rx/ry/rz_mul_vec, cx/cz/cp_mul_vec, block_mul_vec, derv_cphase_mul_vec, dot_x/y/z) and
 * core_op_matrix.py:32-477 (the same on (2^n, m) matrices, x/y/z_dot_mat, derv_cphase).  Host arrays
 * in and out (complex128, interleaved); NOT the hot path, which fuses whole pair-runs per pass.
 *
 * Gate k acts on the index bit of stride strides[2k] (element i pairs with i + stride where
 * (i / stride) is even) with the row-major complex 2x2 matrix gates[8k .. 8k+7]; modes[k]: 0 plain,
 * 1 controlled by the bit of stride strides[2k+1] (identity where it is 0), 2 controlled with ZERO
 * where the control bit is 0 (|1><1| (x) G).  Vectors: stride = 2^(n-1-pos); matrices with m columns:
 * stride = m 2^qubit. */
int aqc_prim_apply(int device, double* host, int64_t count, int num_gates, const int64_t* strides,
                   const int32_t* modes, const double* gates);
/* out[0..1] = sum_i conj((G w)_i) z_i for ONE gate G described as above (np.vdot(G w, z)). */
int aqc_prim_dot(int device, const double* host_w, const double* host_z, int64_t count,
                 const int64_t* strides, const int32_t* modes, const double* gates, double* out);

#ifdef __cplusplus
}
#endif
#endif /* AQC_B200_H */
