//! Raw FFI declarations for `include/quill_b200.h`, one for every exported `qz_*` symbol, in the header's order.
//! Conventions (see the header): Fr/Fq = 32 bytes of little-endian Montgomery limbs (the in-memory layout of
//! `ark_bn254::Fr`), G1 affine = 64 bytes x ‖ y with all-zero = infinity, transcript state = 32 bytes in/out,
//! every function returns a `qz_status` (0 = ok).
#![allow(non_camel_case_types)]
use core::ffi::{c_char, c_void};

#[repr(C)]
pub struct qz_ctx {
    _private: [u8; 0],
}
#[repr(C)]
pub struct qz_srs {
    _private: [u8; 0],
}
/// VirtualPolyExpr node: op 0 Input(a), 1 Const(a = index into consts), 2 Add(a, b), 3 Mul(a, b); children first, root last
#[repr(C)]
#[derive(Clone, Copy, Debug, Default, PartialEq, Eq)]
pub struct qz_expr_node {
    pub op: u32,
    pub a: u32,
    pub b: u32,
}
pub const QZ_EX_INPUT: u32 = 0;
pub const QZ_EX_CONST: u32 = 1;
pub const QZ_EX_ADD: u32 = 2;
pub const QZ_EX_MUL: u32 = 3;

pub const QZ_OK: i32 = 0;
pub const QZ_ERR_INVALID_ARG: i32 = 1;
pub const QZ_ERR_DEGREE: i32 = 2;
pub const QZ_ERR_CUDA: i32 = 3;
pub const QZ_ERR_NCCL: i32 = 4;
pub const QZ_ERR_EXPR: i32 = 5;
pub const QZ_ERR_NO_DEVICE: i32 = 6;
pub const QZ_ERR_ALLOC: i32 = 7;
pub const QZ_MAX_ROUND_COEFFS: usize = 33;

extern "C" {
    // ---- context
    pub fn qz_ctx_create(device: i32, stream: *mut c_void, out: *mut *mut qz_ctx) -> i32;
    pub fn qz_ctx_destroy(ctx: *mut qz_ctx);
    pub fn qz_status_str(status: i32) -> *const c_char;
    pub fn qz_last_error(ctx: *const qz_ctx) -> *const c_char;
    pub fn qz_ctx_sync(ctx: *mut qz_ctx) -> i32;
    pub fn qz_kernel_launches(ctx: *const qz_ctx) -> u64;
    pub fn qz_dev_alloc(ctx: *mut qz_ctx, bytes: usize, out_dev: *mut *mut c_void) -> i32;
    pub fn qz_dev_free(ctx: *mut qz_ctx, dev: *mut c_void) -> i32;
    pub fn qz_dev_trim(ctx: *mut qz_ctx) -> i32;
    pub fn qz_dev_upload(ctx: *mut qz_ctx, dev: *mut c_void, host: *const c_void, bytes: usize) -> i32;
    pub fn qz_dev_download(ctx: *mut qz_ctx, host: *mut c_void, dev: *const c_void, bytes: usize) -> i32;
    pub fn qz_dev_random_fr(ctx: *mut qz_ctx, dev: *mut c_void, n: usize, seed: u64) -> i32;
    // ---- transcript (transcript/src/transcript.rs)
    pub fn qz_transcript_new(domain: *const u8, len: usize, state: *mut u8);
    pub fn qz_transcript_append_bytes(state: *mut u8, msg: *const u8, len: usize);
    pub fn qz_transcript_draw_challenge(state: *mut u8, out: *mut u8, n: usize);
    pub fn qz_transcript_draw_fr(ctx: *mut qz_ctx, state: *mut u8, out_fr: *mut u8) -> i32;
    pub fn qz_transcript_append_fr(ctx: *mut qz_ctx, state: *mut u8, fr: *const u8) -> i32;
    pub fn qz_transcript_append_g1(ctx: *mut qz_ctx, state: *mut u8, xy: *const u8) -> i32;
    pub fn qz_g1_serialize(ctx: *mut qz_ctx, xy: *const u8, out: *mut u8) -> i32;
    // ---- KZG / MSM (pcs/src/kzg.rs)
    pub fn qz_srs_upload(ctx: *mut qz_ctx, xy: *const u8, n: usize, out: *mut *mut qz_srs) -> i32;
    pub fn qz_srs_generate(ctx: *mut qz_ctx, g_xy: *const u8, tau: *const u8, n: usize, out: *mut *mut qz_srs) -> i32;
    pub fn qz_srs_precompute(ctx: *mut qz_ctx, srs: *mut qz_srs, window_bits: i32) -> i32;
    pub fn qz_srs_free(srs: *mut qz_srs);
    pub fn qz_srs_len(srs: *const qz_srs) -> usize;
    pub fn qz_srs_download(ctx: *mut qz_ctx, srs: *const qz_srs, first: usize, count: usize, out_xy: *mut u8) -> i32;
    pub fn qz_msm(ctx: *mut qz_ctx, srs: *const qz_srs, scalars: *const c_void, n_scalars: usize, scalars_on_device: i32,
                  out_xy: *mut u8) -> i32;
    pub fn qz_kzg_commit(ctx: *mut qz_ctx, srs: *const qz_srs, coeffs: *const c_void, n_coeffs: usize,
                         coeffs_on_device: i32, out_xy: *mut u8) -> i32;
    pub fn qz_kzg_open(ctx: *mut qz_ctx, srs: *const qz_srs, coeffs: *const c_void, n_coeffs: usize, coeffs_on_device: i32,
                       x: *const u8, out_y: *mut u8, out_proof_xy: *mut u8) -> i32;
    // ---- multilinear PCS opening (pcs/src/mlpcs.rs, pcs/src/ipa.rs)
    pub fn qz_mlpcs_open(ctx: *mut qz_ctx, srs: *const qz_srs, poly: *const c_void, n: usize, poly_on_device: i32,
                         point: *const u8, n_point: usize, state: *mut u8, out_evaluation: *mut u8, out_s_comm: *mut u8,
                         out_openings: *mut u8) -> i32;
    pub fn qz_mlpcs_open_begin(ctx: *mut qz_ctx, srs: *const qz_srs, poly: *const c_void, n: usize, poly_on_device: i32,
                               point: *const u8, n_point: usize, out_evaluation: *mut u8, out_s_comm: *mut u8,
                               out_s_dev: *mut *mut c_void, out_s_len: *mut usize) -> i32;
    pub fn qz_mlpcs_open_finish(ctx: *mut qz_ctx, srs: *const qz_srs, poly: *const c_void, n: usize, poly_on_device: i32,
                                s_dev: *const c_void, s_len: usize, r: *const u8, out_openings: *mut u8) -> i32;
    pub fn qz_compute_s_polynomial(ctx: *mut qz_ctx, p1: *const u8, n1: usize, p2: *const u8, n2: usize, out: *mut u8) -> i32;
    // ---- sumcheck / zero-check (hyperplonk/src/piops/{sumcheck,zerocheck}.rs)
    pub fn qz_sumcheck_prove(ctx: *mut qz_ctx, num_vars: usize, k: usize, tables: *const *const c_void,
                             tables_on_device: i32, nodes: *const qz_expr_node, n_nodes: usize, consts: *const u8,
                             n_consts: usize, claimed_sum: *const u8, state: *mut u8, max_coeffs: usize,
                             out_coeffs: *mut u8, out_lens: *mut u32, out_point: *mut u8, out_eval: *mut u8) -> i32;
    pub fn qz_zerocheck_prove(ctx: *mut qz_ctx, num_vars: usize, k: usize, tables: *const *const c_void,
                              tables_on_device: i32, nodes: *const qz_expr_node, n_nodes: usize, consts: *const u8,
                              n_consts: usize, state: *mut u8, max_coeffs: usize, out_coeffs: *mut u8,
                              out_lens: *mut u32, out_point: *mut u8, out_eval: *mut u8, out_z: *mut u8) -> i32;
    pub fn qz_logup_denominators(ctx: *mut qz_ctx, num_vars: usize, k: usize, tables: *const *const c_void,
                                 tables_on_device: i32, nodes_h: *const qz_expr_node, n_nodes_h: usize,
                                 nodes_m: *const qz_expr_node, n_nodes_m: usize, consts: *const u8, n_consts: usize,
                                 gamma: *const u8, out: *mut c_void, out_on_device: i32) -> i32;
    pub fn qz_eq_table(ctx: *mut qz_ctx, n: usize, point: *const u8, out: *mut c_void, out_on_device: i32) -> i32;
    // ---- multi-GPU (one process per GPU)
    pub fn qz_comm_unique_id(out_id: *mut u8) -> i32;
    pub fn qz_comm_init(ctx: *mut qz_ctx, unique_id: *const u8, rank: i32, nranks: i32) -> i32;
    pub fn qz_comm_peer_memory(ctx: *const qz_ctx) -> i32;
    pub fn qz_msm_sharded(ctx: *mut qz_ctx, srs_shard: *const qz_srs, scalars_shard: *const c_void, n_scalars: usize,
                          scalars_on_device: i32, out_xy: *mut u8) -> i32;
    pub fn qz_msm_split(ctx: *mut qz_ctx, srs: *const qz_srs, scalars: *const c_void, n_scalars: usize, scalars_on_device: i32, out_xy: *mut u8) -> i32;
    pub fn qz_sumcheck_prove_sharded(ctx: *mut qz_ctx, num_vars: usize, k: usize, table_shards: *const *const c_void,
                                     tables_on_device: i32, nodes: *const qz_expr_node, n_nodes: usize,
                                     consts: *const u8, n_consts: usize, claimed_sum: *const u8, state: *mut u8,
                                     max_coeffs: usize, out_coeffs: *mut u8, out_lens: *mut u32, out_point: *mut u8,
                                     out_eval: *mut u8) -> i32;
    pub fn qz_zerocheck_prove_sharded(ctx: *mut qz_ctx, num_vars: usize, k: usize, table_shards: *const *const c_void,
                                      tables_on_device: i32, nodes: *const qz_expr_node, n_nodes: usize,
                                      consts: *const u8, n_consts: usize, state: *mut u8, max_coeffs: usize,
                                      out_coeffs: *mut u8, out_lens: *mut u32, out_point: *mut u8, out_eval: *mut u8,
                                      out_z: *mut u8) -> i32;
    pub fn qz_comm_resync(ctx: *mut qz_ctx) -> i32;
    pub fn qz_comm_allgather_host(ctx: *mut qz_ctx, send: *const c_void, recv: *mut c_void, bytes: usize) -> i32;
    // ---- measurement and test hooks
    pub fn qz_last_elapsed_ms(ctx: *mut qz_ctx, which: i32) -> f32;
    pub fn qz_last_stat(ctx: *const qz_ctx, which: i32) -> f64;
    pub fn qz_msm_accumulate_stats(ctx: *mut qz_ctx, reset: i32, out_ms: *mut f64, out_mixed_additions: *mut f64, out_launches: *mut u64) -> i32;
    pub fn qz_bench_imad(ctx: *mut qz_ctx, variant: i32, out_ops_per_s: *mut f64) -> i32;
    pub fn qz_bench_fp_mul(ctx: *mut qz_ctx, field: i32, out_muls_per_s: *mut f64) -> i32;
    pub fn qz_test_field_op(ctx: *mut qz_ctx, field: i32, op: i32, a: *const u8, b: *const u8, out: *mut u8, n: usize) -> i32;
    pub fn qz_test_mid_plan(size: u64, pending: i32, k: i32, d: i32, cap: u32, g: i32, out_nblk: *mut u32, out_future: *mut u32, out_chunk: *mut u32, out_tile: *mut u32) -> i32;
    pub fn qz_test_fold(ctx: *mut qz_ctx, r: *const u8, a0: *const u8, a1: *const u8, out: *mut u8, n: usize) -> i32;
    pub fn qz_test_g1_add(ctx: *mut qz_ctx, a_xy: *const u8, b_xy: *const u8, out_xy: *mut u8, n: usize) -> i32;
    pub fn qz_test_g1_mul(ctx: *mut qz_ctx, a_xy: *const u8, scalars: *const u8, out_xy: *mut u8, n: usize) -> i32;
}
