/* The C ABI exercised from plain C, no Python in between (tests/test_gpu_c_abi.py compiles and runs this on the GPU box).
 * Shapes are the reference's own tests: test_sumcheck_proof (hyperplonk/src/piops/sumcheck.rs:159-230: 3 variables,
 * h = g1 * g2) and test_kzg (pcs/src/kzg.rs:119-151: p = 2 + x + 3x^2 on a degree-4 SRS); expected values come from
 * tests/golden/golden.json through the generated header golden_c.h.  Test infrastructure, not product code. */
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "golden_c.h" /* GOLDEN_* byte arrays (canonical little-endian field elements, raw bytes) */
#include "quill_b200.h"

static int failures = 0;
#define CHECK(cond, what)                                  \
  do {                                                     \
    if (!(cond)) {                                         \
      fprintf(stderr, "FAIL: %s (line %d)\n", what, __LINE__); \
      failures++;                                          \
    }                                                      \
  } while (0)
#define OK(call)                                                                      \
  do {                                                                                \
    int rc_ = (call);                                                                 \
    if (rc_ != QZ_OK) {                                                               \
      fprintf(stderr, "%s -> %s: %s\n", #call, qz_status_str(rc_), qz_last_error(ctx)); \
      return 2;                                                                       \
    }                                                                                 \
  } while (0)

static void small(uint8_t out[32], uint64_t v) { /* canonical little-endian limbs of a small integer */
  memset(out, 0, 32);
  for (int i = 0; i < 8; i++) out[i] = (uint8_t)(v >> (8 * i));
}

int main(void) {
  qz_ctx* ctx = NULL;
  int rc = qz_ctx_create(0, NULL, &ctx);
  if (rc != QZ_OK) {
    fprintf(stderr, "qz_ctx_create: %s\n", qz_status_str(rc));
    return 2;
  }
  /* ---- SumcheckProof::prove on g1 = x1 + 2 x2 + 3 x3, g2 = 2 x1 x2 + 3 x1 x3 (sumcheck.rs:161-201) ---- */
  uint8_t can[2][8][32], tab[2][8][32], cs_can[32], cs[32];
  uint64_t sum = 0;
  for (uint64_t i = 0; i < 8; i++) {
    uint64_t a = (i & 1) + 2 * ((i >> 1) & 1) + 3 * ((i >> 2) & 1);
    uint64_t b = (i & 1) * 2 * ((i >> 1) & 1) + 3 * (i & 1) * ((i >> 2) & 1);
    small(can[0][i], a);
    small(can[1][i], b);
    sum += a * b;
  }
  small(cs_can, sum);
  CHECK(sum == 48, "claimed sum of the reference's test is 48");
  OK(qz_test_field_op(ctx, 0, 4, &can[0][0][0], NULL, &tab[0][0][0], 16)); /* to Montgomery form */
  OK(qz_test_field_op(ctx, 0, 4, cs_can, NULL, cs, 1));
  const void* tables[2] = {tab[0], tab[1]};
  const qz_expr_node nodes[3] = {{QZ_EX_INPUT, 0, 0}, {QZ_EX_INPUT, 1, 0}, {QZ_EX_MUL, 0, 1}};
  uint8_t state[32], coeffs[3 * QZ_MAX_ROUND_COEFFS * 32], point[3 * 32], eval[32];
  uint32_t lens[3];
  qz_transcript_new((const uint8_t*)"sumcheck_test", 13, state);
  OK(qz_sumcheck_prove(ctx, 3, 2, tables, 0, nodes, 3, NULL, 0, cs, state, QZ_MAX_ROUND_COEFFS, coeffs, lens, point, eval));
  CHECK(memcmp(state, GOLDEN_SUMCHECK_STATE_END, 32) == 0, "sumcheck: final transcript state");
  uint8_t got[QZ_MAX_ROUND_COEFFS * 32];
  for (int j = 0; j < 3; j++) {
    CHECK(lens[j] == GOLDEN_SUMCHECK_LENS[j], "sumcheck: round polynomial length");
    OK(qz_test_field_op(ctx, 0, 5, coeffs + (size_t)j * QZ_MAX_ROUND_COEFFS * 32, NULL, got, lens[j])); /* canonical */
    CHECK(memcmp(got, GOLDEN_SUMCHECK_RPOLYS[j], 32 * lens[j]) == 0, "sumcheck: round polynomial coefficients");
  }
  OK(qz_test_field_op(ctx, 0, 5, point, NULL, got, 3));
  CHECK(memcmp(got, GOLDEN_SUMCHECK_POINT, 96) == 0, "sumcheck: challenge point");
  OK(qz_test_field_op(ctx, 0, 5, eval, NULL, got, 1));
  CHECK(memcmp(got, GOLDEN_SUMCHECK_EVALUATION, 32) == 0, "sumcheck: final evaluation");

  /* ---- KZG::commit / ::open of p = 2 + x + 3 x^2 on g * tau^i, i <= 4 (kzg.rs:119-151) ---- */
  uint8_t g_mont[64], tau_mont[32], poly_can[3][32], poly[3][32], x_can[32], x[32];
  OK(qz_test_field_op(ctx, 1, 4, GOLDEN_KZG_G, NULL, g_mont, 2)); /* Fq: x, y */
  OK(qz_test_field_op(ctx, 0, 4, GOLDEN_KZG_TAU, NULL, tau_mont, 1));
  qz_srs* srs = NULL;
  OK(qz_srs_generate(ctx, g_mont, tau_mont, 5, &srs));
  CHECK(qz_srs_len(srs) == 5, "srs length");
  small(poly_can[0], 2);
  small(poly_can[1], 1);
  small(poly_can[2], 3);
  small(x_can, 5);
  OK(qz_test_field_op(ctx, 0, 4, &poly_can[0][0], NULL, &poly[0][0], 3));
  OK(qz_test_field_op(ctx, 0, 4, x_can, NULL, x, 1));
  uint8_t com[64], ser[64], y[32], proof[64];
  OK(qz_kzg_commit(ctx, srs, poly, 3, 0, com));
  OK(qz_g1_serialize(ctx, com, ser));
  CHECK(memcmp(ser, GOLDEN_KZG_COMMITMENT_BYTES, 64) == 0, "kzg: serialized commitment");
  OK(qz_kzg_open(ctx, srs, poly, 3, 0, x, y, proof));
  OK(qz_test_field_op(ctx, 0, 5, y, NULL, got, 1));
  CHECK(memcmp(got, GOLDEN_KZG_Y, 32) == 0, "kzg: y = p(5) = 82");
  OK(qz_test_field_op(ctx, 1, 5, proof, NULL, got, 2));
  CHECK(memcmp(got, GOLDEN_KZG_PROOF, 64) == 0, "kzg: opening proof");
  /* the reference panics when the polynomial is longer than the SRS (kzg.rs:62-65): a status here, never an abort */
  uint8_t longpoly[6][32];
  memset(longpoly, 0, sizeof longpoly);
  CHECK(qz_kzg_commit(ctx, srs, longpoly, 6, 0, com) == QZ_ERR_DEGREE, "kzg: degree check");
  /* commit(&[]) is the identity (all-zero encoding) */
  OK(qz_kzg_commit(ctx, srs, poly, 0, 0, com));
  for (int i = 0; i < 64; i++) CHECK(com[i] == 0, "kzg: empty commitment is the identity");
  qz_srs_free(srs);
  printf("launches=%llu\n", (unsigned long long)qz_kernel_launches(ctx));
  qz_ctx_destroy(ctx);
  if (failures) return 1;
  printf("C ABI golden check ok\n");
  return 0;
}
