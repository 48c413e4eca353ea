/*
 * tcc_b200.h -- C ABI of the B200-native `civector` UCC engine.
 *
 * Drop-in boundary for ONE path of TenCirChem (paths relative to the reference
 * tree): the functions that `tencirchem/static/engine_ucc.py` dispatches to for
 * engine="civector" --
 *     GETVECTOR_MAP["civector"]        engine_ucc.py:33-39  -> evolve_civector.py:180  get_civector
 *     ENERGY_AND_GRAD_MAP["civector"]  engine_ucc.py:110-116 -> evolve_civector.py:201 get_energy_and_grad_civector
 *     APPLY_EXCITATION_MAP["civector"] engine_ucc.py:129-136 -> evolve_civector.py:311 apply_excitation_civector
 * and the Hamiltonian action they call, hamiltonian.py:146-183 (`fci_func`).
 *
 * Conventions
 *   - plain pointers and sizes only; every entry point returns an int status
 *     (0 = ok, negative = error class below) and never throws;
 *     tcc_last_error() returns the message of the last failure on this thread.
 *   - `*_dev` pointers are device pointers on the plan's device, `*_host` are host
 *     pointers.  A CI vector is `tcc_civector_size()` doubles, row-major [na, nb]
 *     with the alpha string as the row (ci_utils.py:22-25); hcb: [1, nstr].
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *     Work is enqueued on it; calls that return host results synchronise it.
 *   - a plan is not thread-safe; use one plan per host thread / stream.
 */
#ifndef TCC_B200_H
#define TCC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tcc_plan tcc_plan;

/* status codes; the Python host layer maps them back to the reference's exceptions */
#define TCC_OK 0
#define TCC_ERR_TOO_MANY_QUBITS (-1)   /* ValueError("Too many qubits ...")       ci_utils.py:19-20 */
#define TCC_ERR_BAD_EXCITATION (-2)    /* ValueError("Excitation ... not supported") evolve_civector.py:92-93,
                                          also spin-non-conserving tuples which the reference silently
                                          mis-computes */
#define TCC_ERR_BAD_ARGUMENT (-3)      /* shapes / null pointers / parameter ids  ucc.py:423-424 */
#define TCC_ERR_NO_HAMILTONIAN (-4)    /* energy requested before tcc_set_hamiltonian */
#define TCC_ERR_CUDA (-5)              /* CUDA runtime failure (message in tcc_last_error) */
#define TCC_ERR_NO_DEVICE (-6)         /* no CUDA device: there is no CPU fallback */

const char* tcc_last_error(void);
const char* tcc_version(void);
/* number of CUDA devices visible, or a negative status */
int tcc_device_count(void);

/* ---- plan: the cached excitation tables ------------------------------------------------
 * Replaces get_ci_strings + get_operator_tensors (ci_utils.py:16-37, evolve_civector.py:113-150):
 * instead of [M, N] tables it keeps, per distinct spin-string hop, a compact list of
 * (source address, partner address, sign) -- SURVEY.md F3.
 *   ex_ops_flat / ex_op_len : the excitation tuples, concatenated, and their lengths (2 or 4)
 *   param_ids               : length n_ops, values >= 0 (engine_ucc.py:43-44 default is range)
 *   device                  : CUDA device ordinal
 */
int tcc_plan_create(int n_qubits, int n_elec, int hcb, int n_ops, const int32_t* ex_ops_flat,
                    const int32_t* ex_op_len, const int32_t* param_ids, int device, tcc_plan** out);
void tcc_plan_destroy(tcc_plan* plan);

int64_t tcc_civector_size(const tcc_plan* plan);  /* ucc.py:1282-1290 */
int64_t tcc_num_strings(const tcc_plan* plan);    /* per-spin strings (hcb: all strings) */
int tcc_num_ops(const tcc_plan* plan);
int tcc_num_params(const tcc_plan* plan);         /* max(param_ids)+1, ucc.py:1266-1271 */
/* the batch schedule: runs of excitations (after provably commuting swaps) applied by one launch each */
int tcc_num_batches(const tcc_plan* plan);
int tcc_batch_info(const tcc_plan* plan, int b, int32_t* n_ops, int32_t* active_alpha, int32_t* active_beta,
                   int32_t* n_tiles, int32_t* max_tile_rows, int32_t* max_tile_cols,
                   int32_t* op_indices /* [n_ops of the batch] or NULL */);
/* The pair list of the excitation at position `pos` of batch `b` for tiles whose classes hold (e_alpha, e_beta) active
 * electrons, as the kernels read it (host-side introspection for the tests of the list ORDER: the kernels may apply the
 * pairs of one excitation in any order and with either end first).  which = 0: the plain list of batch_kernel
 * (entries i1 | i2 << 15 | sign << 30, tile-local element indices); 1 / 2: the per-thread streams of the lean kernels
 * for 8-byte / 16-byte elements (the same entries | valid << 31, rounds of one entry per thread, idle slots = 0).
 * n_out = number of entries (0 for a batch without streams), round_len (may be NULL) = entries per round of a stream
 * (thread j of a sub-block applies entry j of every round); out_host may be NULL to query the size. */
int tcc_batch_pairs(const tcc_plan* plan, int b, int pos, int e_alpha, int e_beta, int which, uint32_t* out_host,
                    int capacity, int32_t* n_out, int32_t* round_len);
/* bytes of device memory held by the plan's tables / workspaces */
int64_t tcc_plan_device_bytes(const tcc_plan* plan);

/* ---- introspection for bit-exact parity tests -------------------------------------------
 * per-spin CI strings in address order (non-hcb: the beta strings of ci_utils.py:22; the full
 * string of determinant (a, b) is (s[a] << n_qubits/2) + s[b]); hcb: ci_utils.py:31 */
int tcc_ci_strings_spin(const tcc_plan* plan, uint64_t* out_host /* [num_strings] */);
/* the [N] tables of evolve_civector.py:91-106 for excitation j, expanded on the host from the
 * compact per-spin lists: partner address, G matrix element (fket_phase), diag of G^2 */
int tcc_excitation_tables(const tcc_plan* plan, int j, uint64_t* perm_host, int8_t* phase_host,
                          int8_t* f2_host);
/* the engine's internal string order: ref2int[r] = storage address of the string whose reference
 * address (ci_utils.py:22 order) is r.  Vectors crossing this ABI are ALWAYS in reference order,
 * except for the storage-order building blocks below. */
int tcc_storage_order(const tcc_plan* plan, uint32_t* ref2int_host /* [num_strings] */);

/* ---- Hamiltonian -------------------------------------------------------------------------
 * Replaces get_h_fcifunc_from_integral (hamiltonian.py:146-155; pyscf absorb_h1e with fac 0.5)
 * or, for an hcb plan, get_h_fcifunc_hcb_from_integral (hamiltonian.py:158-183).
 * int1e [n, n], int2e [n, n, n, n] row-major doubles on the host, n = spatial orbitals. */
int tcc_set_hamiltonian(tcc_plan* plan, const double* int1e_host, const double* int2e_host);
/* sigma = H c   (apply_op, hamiltonian.py:240-252).  c and sigma must not alias. */
int tcc_apply_h(tcc_plan* plan, const double* c_dev, double* sigma_dev, void* stream);

/* ---- state preparation: get_civector, evolve_civector.py:180-198 ------------------------
 * params_host [num_params]; init_dev NULL = HF determinant (ci_utils.py:83-87), else a CI vector
 * on the device (copied, never modified); out_dev receives the evolved vector. */
int tcc_civector(tcc_plan* plan, const double* params_host, const double* init_dev, double* out_dev,
                 void* stream);

/* ---- energy / energy and gradient: engine_ucc.py:79-89, evolve_civector.py:201-243 --------
 * energy excludes e_core (added by the caller, ucc.py:655/:707).  grad_host [num_params] already
 * holds 2 * sum over excitations sharing a parameter (evolve_civector.py:213-218). */
int tcc_energy(tcc_plan* plan, const double* params_host, const double* init_dev, double* energy_host,
               void* stream);
int tcc_energy_and_grad(tcc_plan* plan, const double* params_host, const double* init_dev,
                        double* energy_host, double* grad_host, void* stream);

/* ---- batched parameter sets (additive API; the reference accepts one theta per call and loops in Python:
 * kupccgsd.py:98-125 multi-start, h2o_pes.ipynb cell 7).  params_host [n_sets, num_params] row-major;
 * energies_host [n_sets]; grads_host [n_sets, num_params] or NULL (energies only).  Starts from the HF
 * determinant.  All sets of a chunk run in the same launches (one grid dimension per set). */
int tcc_energy_and_grad_batch(tcc_plan* plan, int n_sets, const double* params_host, double* energies_host,
                              double* grads_host, void* stream);

/* ---- G|c> for one excitation: apply_excitation_civector, evolve_civector.py:311-313 ------ */
int tcc_apply_excitation(tcc_plan* plan, const double* c_dev, const int32_t* ex_op, int ex_op_len,
                         double* out_dev, void* stream);

/* ---- host-buffer conveniences (what a ctypes/cffi binding of the reference would call):
 * inputs and outputs live in host memory; copies are part of the call. ---------------------- */
int tcc_civector_host(tcc_plan* plan, const double* params_host, const double* init_host,
                      double* out_host);
int tcc_apply_h_host(tcc_plan* plan, const double* c_host, double* sigma_host);
int tcc_energy_and_grad_host(tcc_plan* plan, const double* params_host, const double* init_host,
                             double* energy_host, double* grad_host);
int tcc_apply_excitation_host(tcc_plan* plan, const double* c_host, const int32_t* ex_op,
                              int ex_op_len, double* out_host);

/* ---- building blocks (used by the sharded multi-GPU driver and the benchmarks).  These operate
 * on vectors in the engine's STORAGE order; convert with the two calls below. --------------- */
int tcc_to_storage_order(tcc_plan* plan, const double* ref_dev, double* storage_dev, void* stream);
int tcc_from_storage_order(tcc_plan* plan, const double* storage_dev, double* ref_dev, void* stream);
/* one UCC factor exp(theta G_j) applied in place to a device vector */
int tcc_rotate(tcc_plan* plan, int j, double theta, double* c_dev, void* stream);
/* reverse-sweep step for excitation j: ket,bra <- exp(-theta G_j); *grad_dev = <bra|G_j|ket> (overwritten) */
int tcc_reverse_step(tcc_plan* plan, int j, double theta, double* ket_dev, double* bra_dev,
                     double* grad_dev /* one double on the device */, void* stream);
/* algorithmic CI-vector bytes one rotation of excitation j must move: 16 * |support_j| */
int64_t tcc_op_support_bytes(const tcc_plan* plan, int j);
/* kernel launches issued by this plan since creation (for bench.py's gpu_launches) */
int64_t tcc_launch_count(const tcc_plan* plan);

/* ---- operator-pool gradients (ADAPT-VQE screening) ----------------------------------------
 * docs/source/tutorial_jupyter/adapt_vqe.ipynb cell 9 loops over the pool calling apply_excitation and a dot product
 * per operator.  Here the excitation list of `plan` IS the pool: out_host[k] = <sigma| G_k |c> for every excitation k
 * of the plan in one batched launch set (the reverse-sweep kernels at zero angles).  c_dev / sigma_dev: reference-order
 * CI vectors on the device (sigma = H c from tcc_apply_h), left unchanged.  The ADAPT gradient of a pool entry is twice
 * the sum over its excitations. */
int tcc_pool_gradients(tcc_plan* plan, const double* c_dev, const double* sigma_dev, double* out_host /* [num_ops] */,
                       void* stream);

/* ---- reduced density matrices (the immediate consumer of the CI vector) ------------------
 * Replaces UCC.make_rdm1 / make_rdm2 in the MO basis (ucc.py:755-852), which copy the CI vector to the host and call
 * pyscf direct_spin1.make_rdm1 / make_rdm12.  c_dev: reference-order CI vector on the device.  Spin-traced:
 * rdm1[p,q] = <q+ p> ([n, n]); rdm2[p,q,r,s] = <p+ r+ s q> ([n, n, n, n], chemists' order, so that
 * E = sum h1[p,q] rdm1[p,q] + 1/2 sum (pq|rs) rdm2[p,q,r,s]).  Either output may be NULL.  Not for hcb plans. */
int tcc_make_rdm12(tcc_plan* plan, const double* c_dev, double* rdm1_host, double* rdm2_host, void* stream);

/* pUCCD density-matrix elements on the hard-core-boson CI vector.  Replaces the numpy mask loops of
 * PUCCD.make_rdm1 / make_rdm2 (tencirchem/static/puccd.py:148-203).  c_dev: the hcb CI vector in reference order on
 * the device.  occ[p] = <n_p> (one spin), hop[p,q] = sum_I c_I c_{I with the pair moved q -> p} (symmetric, zero
 * diagonal), both[p,q] = <n_p n_q> (symmetric, both[p,p] = occ[p]); all [n] / [n, n] host arrays.  hcb plans only. */
int tcc_make_rdm12_hcb(tcc_plan* plan, const double* c_dev, double* occ_host, double* hop_host, double* both_host,
                       void* stream);

/* ---- multi-GPU: alpha-row sharded plans -----------------------------
 * The [na, nb] CI matrix is split by alpha rows over `world` ranks (one process per GPU, world a power of
 * two).  A rank owns the rows whose occupation pattern on `t = log2(world)` alpha orbitals T equals its rank;
 * T is chosen per batch among the orbitals that are NOT active on the alpha side of the batch, so every tile
 * of the batch lies on one rank and the batch kernels run on the local [rows, nb] matrix without any
 * communication.  Consecutive batches with the same T form a segment; between segments the caller moves rows
 * between ranks (all-to-all; tencirchem_b200/sharded.py does it with torch.distributed over NCCL).
 * Replaces the single-process loops of evolve_civector.py:165-177 / :221-243 for vectors that exceed one GPU. */
int tcc_plan_create_sharded(int n_qubits, int n_elec, int n_ops, const int32_t* ex_ops_flat,
                            const int32_t* ex_op_len, const int32_t* param_ids, int device, int rank, int world,
                            tcc_plan** out);
int tcc_shard_info(const tcc_plan* plan, int32_t* rank, int32_t* world, int32_t* n_layouts, int32_t* n_segments,
                   int64_t* max_local_rows);
int tcc_shard_segment(const tcc_plan* plan, int seg, int32_t* layout, int32_t* batch_lo, int32_t* batch_hi);
/* owner_host[a] = rank that owns storage row a in this layout (local rows ascend with a); orbital_mask = T */
int tcc_shard_layout(const tcc_plan* plan, int layout, int32_t* owner_host /* [num_strings] */,
                     uint32_t* orbital_mask);
/* uploads cos / sin of every factor and clears the gradient accumulators: once per evaluation */
int tcc_shard_set_params(tcc_plan* plan, const double* params_host, void* stream);
/* all batches of segment `seg` on the local rows (storage order, layout of the segment) */
int tcc_shard_forward(tcc_plan* plan, int seg, double* vec_local_dev, void* stream);
int tcc_shard_reverse(tcc_plan* plan, int seg, double* ket_local_dev, double* bra_local_dev, void* stream);
/* the same with a row stride (in doubles): 2 * num_strings when ket and bra rows are interleaved in one
 * [rows, 2, nb] block, so that a single row move between segments serves both vectors */
int tcc_shard_reverse_strided(tcc_plan* plan, int seg, double* ket_local_dev, double* bra_local_dev,
                              int64_t row_stride, void* stream);
/* the same on ONE block of interleaved elements kb_local = [local rows][num_strings][2], (ket, bra) of a determinant
 * side by side: a single row move between segments serves both vectors and the kernel gathers / rotates 16-byte
 * elements.  tcc_shard_can_interleave() is 1 when every batch of the plan supports it. */
int tcc_shard_reverse_interleaved(tcc_plan* plan, int seg, double* kb_local_dev, void* stream);
int tcc_shard_can_interleave(const tcc_plan* plan);
/* Row move between two row layouts over PEER MEMORY (NVLink / NVSwitch): the i-th moved local row src[src_row[i], :]
 * is written straight into row dst_row[i] of the destination block of rank dst_rank[i], whose base address as mapped
 * into this process is peer_base[dst_rank[i]] (a symmetric allocation; the own rank's entry is a local address).  One
 * pass of 16-byte stores: no pack buffer, no collective, no unpack.  The caller orders it against the peers' use of the
 * destination block (a stream barrier after the push).  All pointers are device pointers of `device`. */
int tcc_push_rows(int device, const double* src_dev, int64_t row_doubles, int64_t n_rows, const int64_t* src_row_dev,
                  const int64_t* dst_row_dev, const int32_t* dst_rank_dev, const uint64_t* peer_base_dev,
                  int64_t dst_offset_doubles, void* stream);
/* this rank's share of 2 * <bra|G_j|ket> summed by parameter; the caller adds the shares (all-reduce) */
int tcc_shard_gradient(tcc_plan* plan, double* grad_host /* [num_params] */, void* stream);
/* Contributions to sigma = H c of the storage rows [row_lo, row_lo + row_cnt): same-spin terms of these rows and
 * opposite-spin terms scattered from them, ADDED into sigma_full_dev.  Full [na, nb] storage-order matrices; the
 * sum over any partition of the rows is H c (hamiltonian.py:146-155), so ranks split the work by row blocks and
 * all-reduce. */
int tcc_apply_h_rows(tcc_plan* plan, const double* c_full_dev, double* sigma_full_dev, int64_t row_lo,
                     int64_t row_cnt, void* stream);

/* Sharded sigma = H c (hamiltonian.py:146-155) with the vector kept distributed.  layouts3: the three row layouts
 * (pairwise disjoint T) of the alpha-beta passes; layouts3[0] is the layout the forward sweep ends in. */
int tcc_shard_sigma_layouts(const tcc_plan* plan, int32_t* layouts3);
/* sigma_local += (beta-beta terms) c_local for `rows` local rows (any layout: these terms act inside a row) */
int tcc_shard_sigma_bb(tcc_plan* plan, const double* c_local_dev, double* sigma_local_dev, int64_t rows, void* stream);
/* sigma_cols = (alpha-alpha terms) c_cols on a block of `ncols` columns of ALL alpha rows, [num_strings, ncols] */
int tcc_shard_sigma_aa(tcc_plan* plan, const double* c_cols_dev, double* sigma_cols_dev, int64_t ncols, void* stream);
/* sigma_local += pass `pass` (0..2) of the alpha-beta terms: the alpha replacements E_pq that do not touch the
 * orbitals T of layouts3[pass] (the first such layout per pair) keep every row on its rank */
int tcc_shard_sigma_ab(tcc_plan* plan, int pass, const double* c_local_dev, double* sigma_local_dev, void* stream);

/* ---- per-phase device timing (CUDA events on the call's stream, recorded only where the phase
 * changes).  Slots: forward rotations by kind, reverse-sweep steps by kind, the same-spin and
 * opposite-spin parts of H.c, the energy dot product.  `bytes` are ALGORITHMIC CI-vector bytes
 * (DESIGN.md): 32 per rotated pair, 64 per reverse-sweep pair, vector passes for the rest. */
#define TCC_PROF_FWD_BETA 0
#define TCC_PROF_FWD_ALPHA 1
#define TCC_PROF_FWD_ALPHABETA 2
#define TCC_PROF_REV_BETA 3
#define TCC_PROF_REV_ALPHA 4
#define TCC_PROF_REV_ALPHABETA 5
#define TCC_PROF_SIGMA_SS 6
#define TCC_PROF_SIGMA_AB 7
#define TCC_PROF_DOT 8
#define TCC_PROF_FWD_BATCH 9  /* batched forward sweep: 16 N bytes per launch (whole vector read + written) */
#define TCC_PROF_REV_BATCH 10 /* batched reverse sweep: 32 N bytes per launch pair (ket and bra) */
#define TCC_PROF_SLOTS 11
int tcc_profile_enable(tcc_plan* plan, int on); /* also resets the accumulators */
int tcc_profile_read(tcc_plan* plan, double* ms /*[TCC_PROF_SLOTS]*/, int64_t* launches, int64_t* bytes);
/* per-batch device time of the last tcc_energy_and_grad call (forward launch, reverse launch pair), in ms;
 * `on` (de)activates the recording for the following calls; fwd_ms / rev_ms [tcc_num_batches] may be NULL */
int tcc_profile_batches(tcc_plan* plan, int on, float* fwd_ms, float* rev_ms);

#ifdef __cplusplus
}
#endif
#endif /* TCC_B200_H */
