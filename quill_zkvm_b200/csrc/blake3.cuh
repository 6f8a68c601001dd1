// BLAKE3 (default hash mode + XOF) as one __host__ __device__ routine over a contiguous message.
//
// The reference's Fiat-Shamir transcript (transcript/src/transcript.rs:14-75) hashes `state ‖ message` with the
// blake3 crate (Cargo.lock:167).  The sumcheck prover squeezes one challenge per round, so the hash sits on the
// serial critical path of every round; running it on the device (one thread, the four column / diagonal G calls of
// a round give ILP 4) lets a whole proof be enqueued without a host round trip per round.
#pragma once
#include <cstddef>
#include <cstdint>

namespace qz {

#if defined(__CUDACC__)
#define QZ_HD __host__ __device__
#else
#define QZ_HD
#endif

struct Blake3 {
  static constexpr uint32_t F_CHUNK_START = 1, F_CHUNK_END = 2, F_PARENT = 4, F_ROOT = 8;

  QZ_HD static uint32_t iv(int i) {
    constexpr uint32_t t[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au,
                               0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
    return t[i];
  }
  QZ_HD static uint32_t ror(uint32_t x, int n) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(x, x, n);
#else
    return (x >> n) | (x << (32 - n));
#endif
  }

#define QZ_B3_G(a, b, c, d, mx, my) \
  a = a + b + (mx);                 \
  d = ror(d ^ a, 16);               \
  c = c + d;                        \
  b = ror(b ^ c, 12);               \
  a = a + b + (my);                 \
  d = ror(d ^ a, 8);                \
  c = c + d;                        \
  b = ror(b ^ c, 7);

#define QZ_B3_ROUND(m0, m1, m2, m3, m4, m5, m6, m7, m8, m9, m10, m11, m12, m13, m14, m15) \
  QZ_B3_G(s0, s4, s8, s12, m[m0], m[m1])                                                  \
  QZ_B3_G(s1, s5, s9, s13, m[m2], m[m3])                                                  \
  QZ_B3_G(s2, s6, s10, s14, m[m4], m[m5])                                                 \
  QZ_B3_G(s3, s7, s11, s15, m[m6], m[m7])                                                 \
  QZ_B3_G(s0, s5, s10, s15, m[m8], m[m9])                                                 \
  QZ_B3_G(s1, s6, s11, s12, m[m10], m[m11])                                               \
  QZ_B3_G(s2, s7, s8, s13, m[m12], m[m13])                                                \
  QZ_B3_G(s3, s4, s9, s14, m[m14], m[m15])

  // The seven rounds use the message schedule obtained by iterating the fixed permutation; spelled out so the
  // message words stay in registers under full unrolling.
  QZ_HD static void compress(const uint32_t cv[8], const uint32_t m[16], uint64_t counter, uint32_t block_len,
                             uint32_t flags, uint32_t out[16]) {
    uint32_t s0 = cv[0], s1 = cv[1], s2 = cv[2], s3 = cv[3], s4 = cv[4], s5 = cv[5], s6 = cv[6], s7 = cv[7];
    uint32_t s8 = iv(0), s9 = iv(1), s10 = iv(2), s11 = iv(3);
    uint32_t s12 = (uint32_t)counter, s13 = (uint32_t)(counter >> 32), s14 = block_len, s15 = flags;
    QZ_B3_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15)
    QZ_B3_ROUND(2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8)
    QZ_B3_ROUND(3, 4, 10, 12, 13, 2, 7, 14, 6, 5, 9, 0, 11, 15, 8, 1)
    QZ_B3_ROUND(10, 7, 12, 9, 14, 3, 13, 15, 4, 0, 11, 2, 5, 8, 1, 6)
    QZ_B3_ROUND(12, 13, 9, 11, 15, 10, 14, 8, 7, 2, 5, 3, 0, 1, 6, 4)
    QZ_B3_ROUND(9, 14, 11, 5, 8, 12, 15, 1, 13, 3, 0, 10, 2, 6, 4, 7)
    QZ_B3_ROUND(11, 15, 5, 0, 1, 9, 8, 6, 14, 10, 2, 12, 3, 4, 7, 13)
    out[0] = s0 ^ s8;   out[1] = s1 ^ s9;   out[2] = s2 ^ s10;  out[3] = s3 ^ s11;
    out[4] = s4 ^ s12;  out[5] = s5 ^ s13;  out[6] = s6 ^ s14;  out[7] = s7 ^ s15;
    out[8] = s8 ^ cv[0];   out[9] = s9 ^ cv[1];   out[10] = s10 ^ cv[2];  out[11] = s11 ^ cv[3];
    out[12] = s12 ^ cv[4]; out[13] = s13 ^ cv[5]; out[14] = s14 ^ cv[6];  out[15] = s15 ^ cv[7];
  }
#undef QZ_B3_ROUND
#undef QZ_B3_G

  // little-endian load of up to 64 message bytes, zero padded
  QZ_HD static void load_block(const uint8_t* p, size_t n, uint32_t m[16]) {
    for (int i = 0; i < 16; i++) m[i] = 0;
    for (size_t i = 0; i < n; i++) m[i >> 2] |= (uint32_t)p[i] << (8 * (i & 3));
  }

  // Description of the last compression of a subtree, so the caller can finish it with or without the ROOT flag.
  struct Node {
    uint32_t cv[8], m[16];
    uint64_t counter;
    uint32_t block_len, flags;
  };

  // one chunk (<= 1024 bytes): all blocks but the last are compressed, the last is returned pending
  QZ_HD static void chunk_node(const uint8_t* p, size_t len, uint64_t chunk_index, Node& nd) {
    for (int i = 0; i < 8; i++) nd.cv[i] = iv(i);
    size_t nblocks = len == 0 ? 1 : (len + 63) / 64;
    uint32_t out[16];
    for (size_t b = 0; b + 1 < nblocks; b++) {
      load_block(p + 64 * b, 64, nd.m);
      compress(nd.cv, nd.m, chunk_index, 64, b == 0 ? F_CHUNK_START : 0, out);
      for (int i = 0; i < 8; i++) nd.cv[i] = out[i];
    }
    size_t last = nblocks - 1, rem = len - 64 * last;
    load_block(p + 64 * last, rem, nd.m);
    nd.counter = chunk_index;
    nd.block_len = (uint32_t)rem;
    nd.flags = (last == 0 ? F_CHUNK_START : 0) | F_CHUNK_END;
  }
  QZ_HD static void node_cv(const Node& nd, uint32_t cv[8]) {
    uint32_t out[16];
    compress(nd.cv, nd.m, nd.counter, nd.block_len, nd.flags, out);
    for (int i = 0; i < 8; i++) cv[i] = out[i];
  }

  // hash `len` bytes at `in`, write `out_len` bytes of (extended) output
  QZ_HD static void hash(const uint8_t* in, size_t len, uint8_t* out, size_t out_len) {
    const size_t nchunks = len == 0 ? 1 : (len + 1023) / 1024;
    uint32_t stack[40][8];  // chaining values of completed left subtrees
    int sp = 0;
    Node cur;
    for (size_t c = 0; c < nchunks; c++) {
      size_t off = c * 1024, clen = (c + 1 == nchunks) ? len - off : 1024;
      chunk_node(in + off, clen, c, cur);
      if (c + 1 == nchunks) break;
      // a completed, non-final chunk: push its CV, merging equal-height subtrees (binary-counter carry)
      uint32_t cv[8];
      node_cv(cur, cv);
      size_t total = c + 1;
      while ((total & 1) == 0) {
        uint32_t pm[16], po[16], ivs[8];
        for (int i = 0; i < 8; i++) {
          pm[i] = stack[sp - 1][i];
          pm[8 + i] = cv[i];
          ivs[i] = iv(i);
        }
        sp--;
        compress(ivs, pm, 0, 64, F_PARENT, po);
        for (int i = 0; i < 8; i++) cv[i] = po[i];
        total >>= 1;
      }
      for (int i = 0; i < 8; i++) stack[sp][i] = cv[i];
      sp++;
    }
    // fold the pending node with the stacked left siblings, right to left
    while (sp > 0) {
      uint32_t cv[8];
      node_cv(cur, cv);
      sp--;
      for (int i = 0; i < 8; i++) {
        cur.m[i] = stack[sp][i];
        cur.m[8 + i] = cv[i];
        cur.cv[i] = iv(i);
      }
      cur.counter = 0;
      cur.block_len = 64;
      cur.flags = F_PARENT;
    }
    uint64_t block = 0;
    while (out_len) {
      uint32_t w[16];
      compress(cur.cv, cur.m, block++, cur.block_len, cur.flags | F_ROOT, w);
      size_t take = out_len < 64 ? out_len : 64;
      for (size_t i = 0; i < take; i++) out[i] = (uint8_t)(w[i >> 2] >> (8 * (i & 3)));
      out += take;
      out_len -= take;
    }
  }
};

}  // namespace qz
