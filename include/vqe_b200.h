/*
 * vqe_b200.h -- C ABI of the B200-native VQE energy-and-gradient engine.
 *
 * Drop-in boundary for the hot path of OpenVQE (`openvqe.ucc_family`, `openvqe.adapt`).
 * The reference has no FFI layer of its own: its hot path is Python calling myQLM's
 * state-vector simulator and scipy.sparse.  Each entry point below names the reference
 * code whose arithmetic it replaces (paths relative to the reference root).  The Python
 * mirror of the reference interface (`openvqe_b200/ucc_family`, `openvqe_b200/adapt`)
 * binds these symbols with ctypes; INTEGRATION.md shows the stub a reference maintainer
 * would add.
 *
 * Conventions
 *   - State: complex128, interleaved (re, im), 16 bytes per amplitude, 2^n amplitudes,
 *     resident in HBM and owned by the library.  Caller owns every host array.
 *   - Basis index: reference qubit q is index bit (n-1-q)  (myQLM: qubit 0 = MSB).
 *   - A Pauli string is (xmask, zmask, ny) IN INDEX-BIT SPACE: bit b of xmask is set when
 *     the letter on qubit n-1-b is X or Y, bit b of zmask when it is Y or Z, ny = number
 *     of Y letters.   P|i> = i^ny (-1)^popcount(i & zmask) |i ^ xmask>.
 *   - Every function returns VQE_OK (0) or a negative error code; vqe_last_error() returns
 *     a thread-local message.  One host thread per context; one CUDA stream per context.
 *   - There is no CPU fallback: without a CUDA device vqe_create fails.
 */
#ifndef VQE_B200_H
#define VQE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VQE_OK 0
#define VQE_ERR_INVALID (-1)
#define VQE_ERR_CUDA (-2)
#define VQE_ERR_NOMEM (-3)

typedef struct vqe_ctx vqe_ctx;
typedef struct vqe_paulisum vqe_paulisum; /* device-resident, X-mask-grouped Pauli sum */

/* gate kinds for vqe_apply_gates (myQLM conventions: RX(t)=exp(-itX/2), RY(t)=exp(-itY/2),
 * RZ(t)=diag(e^{-it/2}, e^{it/2}), CNOT(control=q0, target=q1)) */
enum { VQE_GATE_X = 0, VQE_GATE_H = 1, VQE_GATE_RX = 2, VQE_GATE_RY = 3, VQE_GATE_RZ = 4, VQE_GATE_CNOT = 5 };

/* state buffers inside a context */
/* VQE_BUF_AUX is a fourth, rank-local vector (never an operand of a peer pass of a sharded state): it keeps the
 * Lanczos ground state that replaces the dense eigh of fermionic_adapt_vqe.py:474 for the fidelity */
enum { VQE_BUF_PSI = 0, VQE_BUF_SIGMA = 1, VQE_BUF_WORK = 2, VQE_BUF_AUX = 3 };

const char* vqe_last_error(void);
int vqe_version(void);
int vqe_device_count(void);

/* Context = one state vector of n_qubits on one device.
 * Replaces: the myQLM `Program`/`qalloc`/`get_default_qpu()` set-up at
 * openvqe/ucc_family/get_energy_ucc.py:38-40. */
int vqe_create(vqe_ctx** out, int n_qubits, int device);
void vqe_destroy(vqe_ctx* ctx);
int vqe_n_qubits(const vqe_ctx* ctx);
/* 1 while the state buffer is kept in the REAL LAYOUT (a purely real state -- |HF> followed by UCC / QUCCSD rotations -- stored
 * as 2^n contiguous doubles; every pass then moves half the bytes), 0 = interleaved complex128.  Measurement only: the layout is
 * internal, every entry point converts as needed. */
int vqe_state_layout(const vqe_ctx* ctx);
/* number of kernels this context has launched so far (bench.py `gpu_launches`) */
uint64_t vqe_launch_count(const vqe_ctx* ctx);
/* cumulative device time (ms) of the named kernel class since the last reset, measured with CUDA
 * events recorded around every launch on the context's stream while profiling is enabled (no extra
 * synchronisation; bench.py roofline leg).
 * which: 0 = state-preparation tile kernel, 1 = expectation kernel, 2 = pauli-sum apply, 3 = pool sweep,
 *        4 = state-preparation peer pass (sharded), 5 = expectation peer pass (sharded) */
int vqe_profile_enable(vqe_ctx* ctx, int on);
int vqe_profile_read(vqe_ctx* ctx, int which, double* ms_total, uint64_t* launches, int reset);

/* Device-side step timer: CUDA events recorded on the context's stream (bench.py times steps with it). */
int vqe_timer_begin(vqe_ctx* ctx);
int vqe_timer_end(vqe_ctx* ctx, double* ms);
/* bytes this context has copied host->device / device->host so far (bench.py e2e accounting) */
int vqe_transfer_bytes(vqe_ctx* ctx, uint64_t* h2d, uint64_t* d2h, int reset);

/* bytes this rank has read from partner shards in gather-form peer passes so far (sharded states; bench.py NVLink
 * accounting: the dependency closure of a pass decides how many partner amplitudes are fetched) */
int vqe_peer_bytes(vqe_ctx* ctx, uint64_t* gathered, int reset);

/* |psi> = |index>.  Replaces the X-gate Hartree-Fock preparation
 * (openvqe/adapt/fermionic_adapt_vqe.py:183-213, get_energy_qucc.py:40-45). */
int vqe_set_basis_state(vqe_ctx* ctx, uint64_t index);
/* host <-> device copies of a whole buffer (interleaved re,im; 2*2^n doubles).
 * get replaces get_statevector (fermionic_adapt_vqe.py:309-328). */
int vqe_set_state(vqe_ctx* ctx, int buf, const double* re_im);
int vqe_get_state(vqe_ctx* ctx, int buf, double* re_im);
int vqe_copy_buffer(vqe_ctx* ctx, int dst_buf, int src_buf);

/* Ordered product  psi <- prod_k exp(-i angle_k P_k) psi  (k = 0 applied first).
 * Replaces build_ucc_ansatz + CLinalg gate-by-gate simulation
 * (get_energy_ucc.py:42-48; fermionic_adapt_vqe.py:126-162; qubit_adapt_vqe.py:271-307) and,
 * for single-string generators, the sparse expm at qubit_adapt_vqe.py:44-55. */
int vqe_apply_pauli_rotations(vqe_ctx* ctx, int n_rot, const uint64_t* xmask, const uint64_t* zmask,
                              const int32_t* ny, const double* angle);

/* The same ordered product on any buffer of the context (VQE_BUF_*).  Used by the opt-in adjoint gradient
 * (openvqe_b200/_hotpath.py: ucc_energy_and_gradient), whose co-state H|psi> lives in VQE_BUF_SIGMA; the reference
 * differentiates by finite differences only (get_energy_ucc.py:158-166). */
int vqe_apply_pauli_rotations_buf(vqe_ctx* ctx, int buf, int n_rot, const uint64_t* xmask, const uint64_t* zmask,
                                  const int32_t* ny, const double* angle);

/* Gate-level circuit (reference qubit numbering, qubit 0 = MSB).  q1 is used by CNOT only.
 * Replaces the QUCCSD excitation circuits of openvqe/common_files/circuit.py:13-106 as
 * executed at get_energy_qucc.py:50-55. */
int vqe_apply_gates(vqe_ctx* ctx, int n_gates, const int32_t* kind, const int32_t* q0, const int32_t* q1,
                    const double* angle);

/* Tabulated plane rotations.  Operation k couples every pair of basis states (l, l ^ xmask[k]); for each listed
 * a-side pattern p (bits inside xmask[k], its highest bit clear; entries tab_offsets[k] .. tab_offsets[k+1]-1)
 * the pairs with (l & xmask[k]) == p rotate as  a' = cos a - sin b,  b' = sin a + cos b;  unlisted patterns are
 * left untouched.  This is the exact unitary of a gate template that only couples states differing by a fixed
 * X-mask -- the QUCCSD excitation circuits of openvqe/common_files/circuit.py:13-93 (run at
 * get_energy_qucc.py:50-55) are of this form -- applied in ONE sweep over 2^-k of the pairs instead of gate by
 * gate.  Not available for X-masks that touch a global qubit of a sharded state (use vqe_apply_gates there). */
int vqe_apply_plane_rotations(vqe_ctx* ctx, int n_ops, const uint64_t* xmask, const int32_t* tab_offsets,
                              const uint64_t* pattern, const double* cosv, const double* sinv);
/* buf <- (re + i im) * buf  (global phase or scale) */
int vqe_scale_state(vqe_ctx* ctx, int buf, double re, double im);
/* dst <- alpha * x + beta * dst on two buffers of the context (the vector updates of the Lanczos iteration that
 * replaces np.linalg.eigh(hamiltonian_sp.get_matrix()), fermionic_adapt_vqe.py:474, and of the Taylor form of
 * expm_multiply, fermionic_adapt_vqe.py:35-38, when it is driven from the host on a sharded state) */
int vqe_axpby(vqe_ctx* ctx, int dst_buf, int x_buf, double alpha_re, double alpha_im, double beta_re, double beta_im);

/* Device-resident Pauli sum  O = sum_k (cre_k + i cim_k) P_k, grouped by X-mask at creation.
 * Replaces the `observable=hamiltonian_sp` argument of the OBS job (get_energy_ucc.py:47) and the
 * 2^n x 2^n scipy matrices hamiltonian_sparse / cluster_ops_sparse (fermionic_adapt_vqe.py:77-122). */
int vqe_paulisum_create(vqe_ctx* ctx, vqe_paulisum** out, int n_terms, const uint64_t* xmask,
                        const uint64_t* zmask, const int32_t* ny, const double* cre, const double* cim);
void vqe_paulisum_destroy(vqe_paulisum* ps);
int vqe_paulisum_groups(const vqe_paulisum* ps); /* number of distinct X-masks */
int vqe_paulisum_passes(const vqe_paulisum* ps); /* number of state sweeps one evaluation makes */

/* out[0] + i out[1] = <buf| O |buf>.  Replaces qpu.submit(OBS job).value (get_energy_ucc.py:48). */
int vqe_expectation(vqe_ctx* ctx, int buf, const vqe_paulisum* ps, double* out_re_im);

/* dst <- O src   (dst != src).  Replaces sig = hamiltonian_sparse.dot(curr_state)
 * (fermionic_adapt_vqe.py:114). */
int vqe_apply_paulisum(vqe_ctx* ctx, int dst_buf, int src_buf, const vqe_paulisum* ps);

/* Pool sweep: for every operator k of the pool (terms op_offsets[k] .. op_offsets[k+1]-1)
 *   out[2k] + i out[2k+1] = <bra_buf| A_k |ket_buf>.
 * The caller forms 2*Re (fermionic ADAPT, fermionic_adapt_vqe.py:67-73) or 2*|.| (qubit ADAPT,
 * qubit_adapt_vqe.py:145-149).  The whole pool is evaluated in one batched sweep. */
int vqe_pool_overlaps(vqe_ctx* ctx, int bra_buf, int ket_buf, int n_ops, const int32_t* op_offsets,
                      const uint64_t* xmask, const uint64_t* zmask, const int32_t* ny, const double* cre,
                      const double* cim, double* out);

/* psi <- exp(theta * A) psi for A = sum_k c_k P_k anti-Hermitian (exact exponential of the whole
 * generator).  Replaces scipy.sparse.linalg.expm_multiply (fermionic_adapt_vqe.py:35-38).
 * Commuting strings collapse to rotations; otherwise a scaled Taylor series on the device. */
int vqe_apply_exp_paulisum(vqe_ctx* ctx, int n_terms, const uint64_t* xmask, const uint64_t* zmask,
                           const int32_t* ny, const double* cre, const double* cim, double theta);

/* out[0] + i out[1] = <vec|buf> with vec a host vector (fun_fidelity, fermionic_adapt_vqe.py:331-361) */
int vqe_overlap_host(vqe_ctx* ctx, int buf, const double* vec_re_im, double* out_re_im);
/* out = <buf|buf> */
int vqe_norm2(vqe_ctx* ctx, int buf, double* out);
/* out[0] + i out[1] = <a|b> for two device buffers */
int vqe_inner(vqe_ctx* ctx, int a_buf, int b_buf, double* out_re_im);

/* ------------------------------------------------------------------------------------------------
 * Sharded state (nothing comparable exists in the reference, which stops at ~24
 * qubits on one host).  The top n_global index bits -- reference qubits 0 .. n_global-1 -- are the rank;
 * every rank owns one context holding 2^(n_qubits - n_global) amplitudes.  All entry points above keep
 * their meaning on a sharded context with FULL-WIDTH masks: rotations / gates / Pauli strings whose
 * X-mask flips global bits run as "peer passes" -- ONE kernel that stages the same tile of ranks r and
 * r ^ m in shared memory (the partner's half read and written through peer memory over NVLink), applies
 * every fusible operation and writes both halves back; the two ranks of a pair split the tiles.  Z
 * letters on global qubits are per-rank signs and move no data.  Reductions (vqe_expectation,
 * vqe_pool_overlaps, vqe_norm2, vqe_inner, vqe_overlap_host) return THIS RANK'S PARTIAL SUM; the caller
 * adds the partials in rank order (torch.distributed all_gather in openvqe_b200/sharded.py).
 *
 * Two ways to connect the ranks:
 *   - one process per GPU: vqe_shard_export -> exchange the 64-byte handles -> vqe_shard_attach_ipc.
 *     Cross-rank ordering is a device-side flag barrier enqueued on the stream around every peer pass
 *     (no host round trip, no collective library on the data path).  All ranks must issue the same calls.
 *   - one process driving all ranks (tests on one GPU; single-process multi-GPU): vqe_shard_attach_local
 *     and the vqe_group_* entry points, which launch every pass on all ranks and order the ranks'
 *     streams with CUDA events. */
#define VQE_IPC_HANDLE_BYTES 64
#define VQE_SHARD_FLAGS 3 /* `what` id of the barrier flag array; 0..2 are the state buffers */
int vqe_create_shard(vqe_ctx** out, int n_qubits, int n_global, int rank, int device);
int vqe_shard_info(const vqe_ctx* ctx, int* n_global, int* rank, int* n_local);
int vqe_shard_export(vqe_ctx* ctx, int what, void* handle_out /* VQE_IPC_HANDLE_BYTES */);
int vqe_shard_attach_ipc(vqe_ctx* ctx, int peer_rank, int what, const void* handle);
int vqe_shard_attach_local(vqe_ctx* ctx, vqe_ctx* peer);
/* Qubit relabelling of a sharded state (no counterpart in the reference): vqe_apply_pauli_rotations moves a qubit that its
 * rotations flip out of a global (rank) index bit into a local one -- one exchange of half a shard over NVLink per move,
 * chosen by Belady's rule over the program -- instead of a peer pass per rotation; the relabelling is internal (masks are
 * always given in the caller's labelling).  Statistics: moves executed and bytes this rank read from partners in them. */
int vqe_relabel_stats(vqe_ctx* ctx, uint64_t* swaps, uint64_t* swap_bytes, int reset);
int vqe_shard_barrier(vqe_ctx* ctx); /* device-side barrier over all ranks, enqueued on the stream */
int vqe_shard_status(vqe_ctx* ctx);  /* VQE_ERR_CUDA after a barrier timed out (a peer died) */

/* in-process groups: ranks[k] must be rank k, n_ranks = 2^n_global */
int vqe_group_apply_pauli_rotations(vqe_ctx* const* ranks, int n_ranks, int n_rot, const uint64_t* xmask,
                                    const uint64_t* zmask, const int32_t* ny, const double* angle);
int vqe_group_apply_gates(vqe_ctx* const* ranks, int n_ranks, int n_gates, const int32_t* kind,
                          const int32_t* q0, const int32_t* q1, const double* angle);
/* ps[k] = the Pauli sum created on ranks[k]; out = sum of the rank partials in rank order */
int vqe_group_expectation(vqe_ctx* const* ranks, int n_ranks, int buf, const vqe_paulisum* const* ps, double* out_re_im);
int vqe_group_apply_paulisum(vqe_ctx* const* ranks, int n_ranks, int dst_buf, int src_buf, const vqe_paulisum* const* ps);
int vqe_group_pool_overlaps(vqe_ctx* const* ranks, int n_ranks, int bra_buf, int ket_buf, int n_ops,
                            const int32_t* op_offsets, const uint64_t* xmask, const uint64_t* zmask,
                            const int32_t* ny, const double* cre, const double* cim, double* out);

/* Host-only view of the pass planner (no CUDA call): cuts an ordered rotation list into tile passes for a
 * state with n_global rank bits.  pass_kind[p]: 0 = local pass, 1 / 2 = peer pass between ranks r and
 * r ^ pass_pattern[p] in exchange form (half-tiles read and written through peer memory) / in gather form (only the
 * partner amplitudes a rank's results depend on are fetched; all writes local).  At most `cap` passes are written;
 * *n_passes is the full count. */
int vqe_plan_rotations(int n_qubits, int n_global, int tile_bits, int low_bits, int n_rot, const uint64_t* xmask,
                       const uint64_t* zmask, const int32_t* ny, const double* angle, int cap, int32_t* n_passes,
                       int32_t* pass_kind, uint64_t* pass_pattern, int32_t* pass_n_ops, uint64_t* pass_tile_mask);

/* Host-only interpreter of the item-table rotation kernels (no CUDA call; test support): plans the program for the real
 * layout of an unsharded n-qubit state as vqe_apply_pauli_rotations does, builds the per-item tables the kernels read
 * (form 1: the 16-bit segments of k_col_stab, the default GPU path; form 0: the 32-bit words of k_col_tab) and applies
 * them tile by tile, with the kernels' per-item arithmetic, to the HOST state psi_re (2^n doubles, in place).
 * VQE_ERR_INVALID when the program has no real-layout collapsed-run plan. */
int vqe_debug_coltab_host(int n_qubits, int tile_bits, int low_bits, int form, int n_rot, const uint64_t* xmask,
                          const uint64_t* zmask, const int32_t* ny, const double* angle, double* psi_re,
                          int32_t* n_passes, int32_t* n_words);

/* Host-only census of the shared-memory bank pairs the item tables of k_col_stab address (no CUDA call; test support): a
 * 64-bit access is served per half-warp, conflict-free when its 16 elements sit in 16 different 8-byte bank pairs.
 * *accesses = half-warp accesses over all segments of the plan, *conflicting = those with two lanes in one bank pair,
 * *unavoidable = those of runs whose free tile positions cannot reach all 16 bank pairs whatever the item order. */
int vqe_debug_coltab_banks(int n_qubits, int tile_bits, int low_bits, int n_rot, const uint64_t* xmask,
                           const uint64_t* zmask, const int32_t* ny, const double* angle, int64_t* accesses,
                           int64_t* conflicting, int64_t* unavoidable);

/* Host-only view of the Pauli-sum planner (no CUDA call): X-mask grouping and packing of the groups into tile
 * passes (what vqe_paulisum_create builds).  *n_groups = distinct X-masks, *n_passes = state sweeps per evaluation;
 * per pass (at most `cap` written): number of groups, number of terms, tile-bit mask. */
int vqe_plan_paulisum(int n_qubits, int n_global, int tile_bits, int low_bits, int n_terms, const uint64_t* xmask,
                      const uint64_t* zmask, const int32_t* ny, const double* cre, const double* cim, int32_t* n_groups,
                      int32_t* n_passes, int cap, int32_t* pass_groups, int32_t* pass_terms, uint64_t* pass_tile_mask);

/* Host-only interpreter of the "lean" passes of a Pauli sum (no CUDA call; test support): walks the entry tables
 * vqe_paulisum_create would upload, with the decode routine the kernels use, on a HOST state of rank `rank`.
 * *out_re = sum over the lean X-mask groups of <psi|O_g|psi>; sigma (optional, may be NULL) += O_lean psi;
 * n_lean_terms / n_fat_terms = Pauli strings handled by lean passes / left to the general passes. */
int vqe_debug_lean_host(int n_qubits, int n_global, int rank, int tile_bits, int low_bits, int n_terms, const uint64_t* xmask,
                        const uint64_t* zmask, const int32_t* ny, const double* cre, const double* cim,
                        const double* psi_re_im, double* out_re, int32_t* n_lean_terms, int32_t* n_fat_terms,
                        double* sigma_re_im);

/* Host-only interpreter of the diagonal-part kernel k_expect_diag2_rl (no CUDA call; test support): the X-mask-0 strings of
 * the sum as a quadratic form in the Z letters, built as vqe_paulisum_create builds it, evaluated on the HOST state psi_re
 * (2^(n_qubits - n_global) doubles of rank `rank`) with the kernel's chunk decomposition.  *is_form = 0 when the diagonal
 * part is not a quadratic form (the GPU path then takes the general pass). */
int vqe_debug_diag2_host(int n_qubits, int n_global, int rank, int n_terms, const uint64_t* xmask, const uint64_t* zmask,
                         const int32_t* ny, const double* cre, const double* cim, const double* psi_re, double* out_re,
                         int32_t* is_form);

/* Host-only check (no CUDA call; test support) of the tensor-map (TMA) form of the tile plan that covers the local
 * index bits `need_mask`: emulates the box traversal of every request and compares with the gather addresses of the
 * tile.  Returns the number of mismatching elements (0 = consistent), -1 when the plan keeps per-segment copies. */
int vqe_debug_tma_check(int n_local, uint64_t need_mask, int tile_bits, int low_bits, int max_tiles, int32_t* n_req,
                        int32_t* dims_used);
/* the same for the real layout of the state buffer (one double per amplitude, tile elements of 8 bytes) */
int vqe_debug_tma_check_rl(int n_local, uint64_t need_mask, int tile_bits, int low_bits, int max_tiles, int32_t* n_req,
                           int32_t* dims_used);

/* Raw device pointer / stream of a buffer. */
int vqe_buffer_ptr(vqe_ctx* ctx, int buf, void** dev_ptr, uint64_t* n_amplitudes);
int vqe_synchronize(vqe_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* VQE_B200_H */
