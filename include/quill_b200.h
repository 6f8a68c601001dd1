/* quill_b200.h -- C ABI of the B200-native proving hot path of Quill (gio54321/quill-zkvm).
 *
 * The reference has no FFI layer; its seam is the Rust API below (paths relative to the reference root).  Each entry
 * point here is what a Rust shim binds to replace that call with the sm_100a CUDA path:
 *
 *   KZG::commit / msm_unchecked        pcs/src/kzg.rs:61-73          -> qz_kzg_commit, qz_msm
 *   KZG::open                          pcs/src/kzg.rs:75-96          -> qz_kzg_open
 *   KZG::trusted_setup (G1 powers)     pcs/src/kzg.rs:35-59          -> qz_srs_upload, qz_srs_generate
 *   MultilinearPCS::commit             pcs/src/mlpcs.rs:188-190      -> qz_kzg_commit
 *   MultilinearPCS::open               pcs/src/mlpcs.rs:83-124,191-198 -> qz_mlpcs_open
 *   compute_s_polynomial               pcs/src/ipa.rs:122-157        -> qz_compute_s_polynomial
 *   SumcheckProof::prove               hyperplonk/src/piops/sumcheck.rs:28-114   -> qz_sumcheck_prove
 *   ZeroCheckProof::prove              hyperplonk/src/piops/zerocheck.rs:14-49   -> qz_zerocheck_prove
 *   (sharded over the GPUs of one box: qz_msm_sharded, qz_sumcheck_prove_sharded, qz_zerocheck_prove_sharded)
 *   fast_eq_eval_hypercube             hyperplonk/src/utils/eq_eval.rs:6-31      -> qz_eq_table
 *   logup denominators                 hyperplonk/src/piops/multiset_check.rs:43-95 -> qz_logup_denominators
 *   Transcript                         transcript/src/transcript.rs:14-75        -> qz_transcript_*
 *
 * Data conventions (all little-endian):
 *   Fr / Fq element   32 bytes = 4 x u64 Montgomery limbs (R = 2^256): the in-memory layout of ark_bn254::Fr / Fq,
 *                     so `&[Fr]` crosses zero-copy.
 *   G1 affine point   64 bytes = x ‖ y (Montgomery Fq).  All-zero stands for the point at infinity ((0,0) is not on
 *                     y^2 = x^3 + 3).  ark's G1Affine is not repr(C); the shim packs it once at SRS upload.
 *   transcript state  32 bytes (`Transcript.state`, a pub Vec<u8>, transcript/src/transcript.rs:6-9), in/out.
 *   expression tree   VirtualPolyExpr (hyperplonk/src/utils/virtual_polynomial.rs:9-18) flattened to an array of
 *                     qz_expr_node, children before parents, the root last.
 *
 * Ownership: the caller owns every host buffer; the library copies in and writes results to caller buffers.
 * Errors: every function returns a qz_status; nothing aborts.  The reference's prover panics on the conditions
 * mapped to QZ_ERR_DEGREE / QZ_ERR_INVALID_ARG; a shim turns non-zero into panic!.
 * Threading: one in-flight call per context; contexts are independent.
 * There is NO CPU fallback: without a CUDA device qz_ctx_create fails with QZ_ERR_NO_DEVICE.
 */
#ifndef QUILL_B200_H
#define QUILL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  QZ_OK = 0,
  QZ_ERR_INVALID_ARG = 1, /* null pointer, inconsistent sizes, table length != 2^num_vars (virtual_polynomial.rs:162-166) */
  QZ_ERR_DEGREE = 2,      /* polynomial longer than the SRS: "Polynomial degree exceeds max degree" (kzg.rs:62-65) */
  QZ_ERR_CUDA = 3,        /* a CUDA runtime call failed; see qz_last_error */
  QZ_ERR_NCCL = 4,        /* an NCCL call failed */
  QZ_ERR_EXPR = 5,        /* malformed expression tree, or degree / size beyond the compiled limits */
  QZ_ERR_NO_DEVICE = 6,   /* no usable CUDA device (there is no CPU path) */
  QZ_ERR_ALLOC = 7        /* device or host allocation failed */
} qz_status;

typedef struct qz_ctx qz_ctx; /* one device, one stream, scratch memory */
typedef struct qz_srs qz_srs; /* device-resident KZG G1 powers */

/* VirtualPolyExpr node.  op: 0 Input(a = polynomial index), 1 Const(a = index into consts),
 * 2 Add(a, b = node indices), 3 Mul(a, b = node indices). */
typedef struct {
  uint32_t op, a, b;
} qz_expr_node;
enum { QZ_EX_INPUT = 0, QZ_EX_CONST = 1, QZ_EX_ADD = 2, QZ_EX_MUL = 3 };

#define QZ_MAX_ROUND_COEFFS 33 /* round polynomials of degree <= 32 */

/* ---- context ------------------------------------------------------------------------------------------------- */
/* `stream` is a cudaStream_t to launch on (e.g. the caller's current stream) or NULL to create a private one. */
int qz_ctx_create(int device, void* stream, qz_ctx** out);
void qz_ctx_destroy(qz_ctx* ctx);
const char* qz_status_str(int status);
const char* qz_last_error(const qz_ctx* ctx); /* detail of the last non-OK status on this context */
int qz_ctx_sync(qz_ctx* ctx);                 /* cudaStreamSynchronize on the context's stream */
/* number of kernels this library launched on the context since creation (bench accounting) */
uint64_t qz_kernel_launches(const qz_ctx* ctx);

/* device buffers for callers that keep inputs resident (the kernel-only bench leg; a shim may ignore these).
 * qz_dev_free parks the block in the context's pool for the next qz_dev_alloc of (about) that size instead of calling
 * cudaFree -- a device-wide synchronisation a prover would otherwise pay a dozen times per proof; qz_dev_trim returns
 * the parked blocks to the driver (qz_ctx_destroy does too). */
int qz_dev_alloc(qz_ctx* ctx, size_t bytes, void** out_dev);
int qz_dev_free(qz_ctx* ctx, void* dev);
int qz_dev_trim(qz_ctx* ctx);
int qz_dev_upload(qz_ctx* ctx, void* dev, const void* host, size_t bytes);   /* H2D, synchronous on return */
int qz_dev_download(qz_ctx* ctx, void* host, const void* dev, size_t bytes); /* D2H, synchronous on return */
/* fill dev with n pseudo-random Fr elements (Montgomery form of uniformly-spread values < r), seeded */
int qz_dev_random_fr(qz_ctx* ctx, void* dev, size_t n, uint64_t seed);

/* ---- transcript (host side; blake3) -- transcript/src/transcript.rs ------------------------------------------ */
void qz_transcript_new(const uint8_t* domain, size_t len, uint8_t state[32]);              /* :15-23 */
void qz_transcript_append_bytes(uint8_t state[32], const uint8_t* msg, size_t len);        /* :26-32 */
void qz_transcript_draw_challenge(uint8_t state[32], uint8_t* out, size_t n);              /* :49-63 */
/* draw_field_element::<Fr>() (:71-75): 48 challenge bytes reduced mod r, returned in Montgomery form */
int qz_transcript_draw_fr(qz_ctx* ctx, uint8_t state[32], uint8_t out_fr[32]);
/* append_serializable(&Fr) / (&G1): canonical encodings of ark-serialize (32 B; 64 B with the y-sign / infinity flags) */
int qz_transcript_append_fr(qz_ctx* ctx, uint8_t state[32], const uint8_t fr[32]);
int qz_transcript_append_g1(qz_ctx* ctx, uint8_t state[32], const uint8_t xy[64]);
/* ark-serialize uncompressed bytes of a G1 point (what a Commitment contributes to the transcript) */
int qz_g1_serialize(qz_ctx* ctx, const uint8_t xy[64], uint8_t out[64]);

/* ---- KZG / MSM -- pcs/src/kzg.rs ------------------------------------------------------------------------------ */
/* Upload n affine G1 points (the normalised `g1_points`).  The reference re-normalises the SRS on every commit
 * (kzg.rs:67-71); here that happens once, in the shim, before this call. */
int qz_srs_upload(qz_ctx* ctx, const uint8_t* xy, size_t n, qz_srs** out);
/* Build g * tau^i, i < n, on the device (kzg.rs:44-47) from an affine generator and tau (Montgomery Fr). */
int qz_srs_generate(qz_ctx* ctx, const uint8_t g_xy[64], const uint8_t tau[32], size_t n, qz_srs** out);
/* Optional, once per SRS: store 2^(c w) * P_i for every window w (W x the SRS in HBM; 12 GiB at 2^24 with c = 22) so that all
 * windows of a scalar share one bucket set -- fewer additions per point, one bucket reduction, no window-combine doublings.
 * window_bits = 0 picks c from the SRS size.  Later MSMs use the table whenever the cost model says it is cheaper. */
int qz_srs_precompute(qz_ctx* ctx, qz_srs* srs, int window_bits);
void qz_srs_free(qz_srs* srs);
size_t qz_srs_len(const qz_srs* srs);
int qz_srs_download(qz_ctx* ctx, const qz_srs* srs, size_t first, size_t count, uint8_t* out_xy);

/* msm_unchecked(bases, scalars) (kzg.rs:72): sum_i scalars[i] * bases[i] over i < min(n_scalars, srs len).
 * `scalars` is a host pointer unless scalars_on_device != 0.  Result: affine x ‖ y, all-zero for the identity. */
int qz_msm(qz_ctx* ctx, const qz_srs* srs, const void* scalars, size_t n_scalars, int scalars_on_device,
           uint8_t out_xy[64]);
/* KZG::commit (kzg.rs:61-73): QZ_ERR_DEGREE when n_coeffs > srs len (the reference panics), else the MSM. */
int qz_kzg_commit(qz_ctx* ctx, const qz_srs* srs, const void* coeffs, size_t n_coeffs, int coeffs_on_device,
                  uint8_t out_xy[64]);
/* KZG::open (kzg.rs:75-96): y = p(x), proof = commit((p - y) / (X - x)).  The reference's multiply-back assert
 * (:85) has no output and is not reproduced. */
int qz_kzg_open(qz_ctx* ctx, const qz_srs* srs, const void* coeffs, size_t n_coeffs, int coeffs_on_device,
                const uint8_t x[32], uint8_t out_y[32], uint8_t out_proof_xy[64]);

/* ---- multilinear PCS opening -- pcs/src/mlpcs.rs, pcs/src/ipa.rs ------------------------------------------------- */
/* MLEvalProof::prove(poly, eval_point, kzg, transcript) (mlpcs.rs:83-124): P_r (= eq table of the point), evaluation
 * <poly, P_r>, S polynomial (NTT), commit(S), transcript (absorb point, evaluation, S commitment; squeeze r), and the
 * four KZG openings of poly and S at r and 1/r -- five MSMs, everything resident on the device.
 *   out_openings   4 x [x (32) ‖ y (32) ‖ proof (64)] in the order poly_opening, poly_opening_inv, s_opening,
 *                  s_opening_inv (mlpcs.rs:109-113)
 * QZ_ERR_DEGREE where the reference's commit would panic (S or a quotient longer than the SRS). */
int qz_mlpcs_open(qz_ctx* ctx, const qz_srs* srs, const void* poly, size_t n, int poly_on_device, const uint8_t* point,
                  size_t n_point, uint8_t state[32], uint8_t out_evaluation[32], uint8_t out_s_comm[64],
                  uint8_t out_openings[512]);
/* The same opening in the two halves its Fiat-Shamir challenge separates, so that independent openings (HyperPlonk issues
 * num_cols + num_public + 5 per trace, proof.rs:202-226, multiset_check.rs:167-170) can be spread over GPUs:
 *   begin   P_r, evaluation, S, commit(S) (mlpcs.rs:86-97); S stays on the device in *out_s_dev (release it with
 *           qz_dev_free), *out_s_len is the length its commitment and openings use
 *   (caller: absorb point, evaluation, S commitment; squeeze r -- mlpcs.rs:100-105, the qz_transcript_* calls)
 *   finish  the four KZG openings of poly and S at r and 1/r (mlpcs.rs:107-113), layout as in qz_mlpcs_open */
int qz_mlpcs_open_begin(qz_ctx* ctx, const qz_srs* srs, const void* poly, size_t n, int poly_on_device,
                        const uint8_t* point, size_t n_point, uint8_t out_evaluation[32], uint8_t out_s_comm[64],
                        void** out_s_dev, size_t* out_s_len);
int qz_mlpcs_open_finish(qz_ctx* ctx, const qz_srs* srs, const void* poly, size_t n, int poly_on_device,
                         const void* s_dev, size_t s_len, const uint8_t r[32], uint8_t out_openings[512]);
/* InnerProductProof::compute_s_polynomial (ipa.rs:122-157): writes max(n1, n2) - 1 coefficients to `out` (host), NOT
 * trimmed (the reference's DensePolynomial drops trailing zeros; they do not change any commitment or opening). */
int qz_compute_s_polynomial(qz_ctx* ctx, const uint8_t* p1, size_t n1, const uint8_t* p2, size_t n2, uint8_t* out);

/* ---- sumcheck / zero-check -- hyperplonk/src/piops/{sumcheck,zerocheck}.rs ------------------------------------- */
/* SumcheckProof::prove(num_vars, store, h, claimed_sum, transcript) (sumcheck.rs:28-114).
 *   tables[k]       the store's polynomials, each 2^num_vars Fr; host pointers unless tables_on_device != 0
 *   nodes/consts    h as a flattened VirtualPolyExpr; consts are Montgomery Fr
 *   state           transcript state, updated exactly as the reference's `&mut Transcript`
 *   out_coeffs      num_vars x max_coeffs Fr (Montgomery), row j = r_polys[j].coeffs zero padded
 *   out_lens        num_vars: r_polys[j].coeffs.len() (DensePolynomial trims trailing zeros)
 *   out_point       num_vars Fr: the challenges r_0 .. r_{n-1};   out_eval: EvaluationClaim.evaluation
 * Returns QZ_ERR_EXPR if deg(h)+1 > max_coeffs. */
int qz_sumcheck_prove(qz_ctx* ctx, size_t num_vars, size_t k, const void* const* tables, int tables_on_device,
                      const qz_expr_node* nodes, size_t n_nodes, const uint8_t* consts, size_t n_consts,
                      const uint8_t claimed_sum[32], uint8_t state[32], size_t max_coeffs, uint8_t* out_coeffs,
                      uint32_t* out_lens, uint8_t* out_point, uint8_t out_eval[32]);
/* ZeroCheckProof::prove(store, h, transcript) (zerocheck.rs:14-49): draws z, builds eq(., z) on the device as one more
 * table, proves sum h*eq = 0, divides the final claim by eq(z, r).  out_z: the num_vars challenges z. */
int qz_zerocheck_prove(qz_ctx* ctx, size_t num_vars, size_t k, const void* const* tables, int tables_on_device,
                       const qz_expr_node* nodes, size_t n_nodes, const uint8_t* consts, size_t n_consts,
                       uint8_t state[32], size_t max_coeffs, uint8_t* out_coeffs, uint32_t* out_lens,
                       uint8_t* out_point, uint8_t out_eval[32], uint8_t* out_z);
/* Logup denominators (hyperplonk/src/piops/multiset_check.rs:43-95): out[i] = m(row_i) / (gamma + h(row_i)) for every row of
 * the store, with h and the optional multiplicities expression m (n_nodes_m = 0 means m = 1, "Equality" mode) given as
 * flattened VirtualPolyExpr over the same tables and the same consts array.  Batch inversion on the device.
 * QZ_ERR_INVALID_ARG if some gamma + h(row) is zero (the reference panics on `.inverse().unwrap()`). */
int qz_logup_denominators(qz_ctx* ctx, size_t num_vars, size_t k, const void* const* tables, int tables_on_device,
                          const qz_expr_node* nodes_h, size_t n_nodes_h, const qz_expr_node* nodes_m, size_t n_nodes_m,
                          const uint8_t* consts, size_t n_consts, const uint8_t gamma[32], void* out, int out_on_device);
/* fast_eq_eval_hypercube(n, point) (eq_eval.rs:6-31) -> 2^n Fr written to `out` (host, or device if out_on_device). */
int qz_eq_table(qz_ctx* ctx, size_t n, const uint8_t* point, void* out, int out_on_device);

/* ---- multi-GPU (one process per GPU) ------------------------------------------------------------------ */
/* Join an NCCL communicator: `unique_id` is the 128-byte ncclUniqueId produced by qz_comm_unique_id on rank 0 and
 * broadcast by the caller (torch.distributed, MPI, a file...). */
int qz_comm_unique_id(uint8_t out_id[128]);
int qz_comm_init(qz_ctx* ctx, const uint8_t unique_id[128], int rank, int nranks);
/* 1 when qz_comm_init mapped every peer's mailbox into this process (CUDA IPC over NVLink / NVSwitch): the per-round
 * sumcheck exchange and the MSM partial-sum exchange are then stores into the peers' HBM issued by the kernel that
 * produced the value, with no collective launch in between.  0: the exchanges are NCCL all-gathers (peer access
 * unavailable, or QZ_NO_P2P=1 in the environment at qz_comm_init).  The results are identical either way. */
int qz_comm_peer_memory(const qz_ctx* ctx);
/* Sharded MSM: this rank holds SRS points [first, first+len) and the matching scalars; every rank receives the full
 * commitment (partial sums are gathered as raw limbs and added with the group law, never ncclSum). */
int qz_msm_sharded(qz_ctx* ctx, const qz_srs* srs_shard, const void* scalars_shard, size_t n_scalars,
                   int scalars_on_device, uint8_t out_xy[64]);
/* The same product when every rank holds the WHOLE SRS and the WHOLE scalar vector (the witness and logup-denominator
 * commits of HyperPlonk::prove, proof.rs:270-276 / multiset_check.rs:98-99, at N > 1): rank g multiplies the index
 * range [n g / G, n (g + 1) / G) and every rank receives KZG::commit(scalars).  Without a communicator it is
 * qz_kzg_commit.  QZ_ERR_DEGREE when n_scalars exceeds the SRS (kzg.rs:65). */
int qz_msm_split(qz_ctx* ctx, const qz_srs* srs, const void* scalars, size_t n_scalars, int scalars_on_device,
                 uint8_t out_xy[64]);
/* Sharded sumcheck: rank g holds elements [g*2^m, (g+1)*2^m) of every table, m = num_vars - log2(nranks), i.e. the
 * tables are split by the top variables so every (2p, 2p+1) pair is local (sumcheck.rs:56-57).  Per round the ranks
 * exchange (deg+1) partial sums; all ranks run the same transcript and return the same proof. */
int qz_sumcheck_prove_sharded(qz_ctx* ctx, size_t num_vars, size_t k, const void* const* table_shards,
                              int tables_on_device, const qz_expr_node* nodes, size_t n_nodes, const uint8_t* consts,
                              size_t n_consts, const uint8_t claimed_sum[32], uint8_t state[32], size_t max_coeffs,
                              uint8_t* out_coeffs, uint32_t* out_lens, uint8_t* out_point, uint8_t out_eval[32]);

/* Sharded zero-check: qz_zerocheck_prove over the same table shards; z is drawn identically on every rank, rank g
 * builds elements [g*2^m, (g+1)*2^m) of the eq table (or of its weight tables), all ranks return the same proof. */
int qz_zerocheck_prove_sharded(qz_ctx* ctx, size_t num_vars, size_t k, const void* const* table_shards,
                               int tables_on_device, const qz_expr_node* nodes, size_t n_nodes, const uint8_t* consts,
                               size_t n_consts, uint8_t state[32], size_t max_coeffs, uint8_t* out_coeffs,
                               uint32_t* out_lens, uint8_t* out_point, uint8_t out_eval[32], uint8_t* out_z);

/* Collective (every rank calls it): re-agree on the peer-mailbox sequence numbers after a sharded call returned an error
 * on some rank -- a rank that left a call early has consumed fewer exchange numbers than the others and every later
 * wait would time out (QZ_ERR_NCCL after 20 s).  Also clears the timed-out mark.  A no-op without a communicator. */
int qz_comm_resync(qz_ctx* ctx);
/* All-gather of `bytes` host bytes per rank into `recv` (nranks * bytes, rank-major) over the library's communicator:
 * the exchange of (evaluation, S commitment) and of finished openings between the halves above. */
int qz_comm_allgather_host(qz_ctx* ctx, const void* send, void* recv, size_t bytes);

/* ---- measurement hooks ---------------------------------------------------------------------------------------------- */
/* CUDA-event timing of the most recent call on this context, on the context's stream (milliseconds):
 *   which = 0 whole call (device side), 1 dominant kernel(s) only (sumcheck round kernels / MSM bucket accumulation) */
float qz_last_elapsed_ms(qz_ctx* ctx, int which);
/* shape of the most recent MSM on this context: which = 0 window bits c, 1 digits (mixed additions) per scalar,
 * 2 whether the precomputed shared bucket set was used (0/1), 3 total mixed additions */
double qz_last_stat(const qz_ctx* ctx, int which);
/* msm_accumulate (the MSM's dominant kernel) over MANY calls: reset = 1 starts collecting -- every accumulate launch is
 * then bracketed by its own CUDA event pair on the context's stream --, reset = 0 reads the totals so far (summed
 * kernel time in milliseconds, mixed additions executed, launches; synchronises the stream first), reset = -1 reads
 * and stops.  Lets a caller that runs several MSMs per step (MultilinearPCS::open: 5, HyperPlonk::prove: ~140) report
 * the kernel's share of the step and its integer-pipe fraction.  Not collecting is the default and costs nothing. */
int qz_msm_accumulate_stats(qz_ctx* ctx, int reset, double* out_ms, double* out_mixed_additions, uint64_t* out_launches);
/* integer-pipe micro-benchmark: returns 32x32->64 multiply-accumulates per second sustained by all SMs (the MSM
 * roofline denominator).  variant 0 = IMAD.WIDE.U32 carry chains as used by the field multiplier, 1 = 32-bit IMAD. */
int qz_bench_imad(qz_ctx* ctx, int variant, double* out_ops_per_s);
/* Montgomery multiplications per second (Fr if field == 0, Fq if 1) on `n` independent chains */
int qz_bench_fp_mul(qz_ctx* ctx, int field, double* out_muls_per_s);

/* test hooks: element-wise field ops on host buffers through the device (op: 0 add, 1 sub, 2 mul, 3 inverse,
 * 4 to_mont, 5 from_mont, 6 a*b + a*a + b*b through the deferred-reduction accumulator, 7 4096 * (a*b) through it,
 * 8 a*b + (a+b)*(a-b) by the fused two-product multiplier, 9 a*a by the dedicated squaring,
 * 10 inverse by the binary extended Euclidean algorithm (the single-thread critical-path inverse);
 * field: 0 Fr, 1 Fq); G1 add / scalar-mul */
int qz_test_field_op(qz_ctx* ctx, int field, int op, const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n);
/* out[i] = a0[i] + r (a1[i] - a0[i]) by the fixed-challenge fold of the large sumcheck passes (table of shifted multiples
 * of r in constant memory, 88 instead of 136 multiply-adds per product); a1 - a0 is taken without reduction, so any
 * 256-bit a0 < p, a1 < p are valid */
int qz_test_fold(qz_ctx* ctx, const uint8_t r[32], const uint8_t* a0, const uint8_t* a1, uint8_t* out, size_t n);
/* test hook, host logic only (no device needed): the round plan of the persistent short-round kernel for tables of
 * `size` entries per rank (pending != 0: a challenge still to be folded in), k tables, degree d, at most `cap`
 * co-resident blocks, G ranks.  Arrays of QZ_TEST_PLAN_ROUNDS entries: blocks working on round j, the largest block
 * count any round >= j needs, pairs per block (0: whole pairs per thread).  Returns the grid size, < 0 on bad input. */
#define QZ_TEST_PLAN_ROUNDS 40
int qz_test_mid_plan(uint64_t size, int pending, int k, int d, unsigned int cap, int G, uint32_t* out_nblk,
                     uint32_t* out_future, uint32_t* out_chunk, uint32_t* out_tile);
int qz_test_g1_add(qz_ctx* ctx, const uint8_t* a_xy, const uint8_t* b_xy, uint8_t* out_xy, size_t n);
int qz_test_g1_mul(qz_ctx* ctx, const uint8_t* a_xy, const uint8_t* scalars, uint8_t* out_xy, size_t n);

#ifdef __cplusplus
}
#endif
#endif /* QUILL_B200_H */
