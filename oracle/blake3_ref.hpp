// CPU ORACLE (test infrastructure, NOT product code).
// BLAKE3 (hash mode, arbitrary input length, arbitrary XOF output length) restated from the published
// algorithm; the reference uses the `blake3` crate 1.8.2 (Cargo.lock:167) at transcript/src/transcript.rs:15-31
// and :49-56.  Checked in tests/ against the `blake3` Python wheel, which wraps that same crate.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace orc {

class Blake3 {
 public:
  static constexpr uint32_t IV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au,
                                     0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
  enum : uint32_t { CHUNK_START = 1, CHUNK_END = 2, PARENT = 4, ROOT = 8 };

  Blake3() { start_chunk(0); }
  void update(const uint8_t* p, size_t n) {
    while (n) {
      if (chunk_len_ == 1024) {  // chunk full and more input follows: finish it as a non-root chunk
        uint32_t cv[8];
        chunk_output().chaining_value(cv);
        push_cv(cv, ++total_chunks_);
        start_chunk(total_chunks_);
      }
      if (buf_len_ == 64) {  // flush a full block (it is not the last block of the chunk)
        uint32_t w[16], out[16];
        words(buf_, w);
        compress(chunk_cv_, w, chunk_counter_, 64, block_flags(), out);
        memcpy(chunk_cv_, out, 32);
        blocks_done_++;
        buf_len_ = 0;
        memset(buf_, 0, 64);
      }
      size_t take = std::min<size_t>(64 - buf_len_, n);
      memcpy(buf_ + buf_len_, p, take);
      buf_len_ += take;
      chunk_len_ += take;
      p += take;
      n -= take;
    }
  }
  void finalize(uint8_t* out, size_t out_len) {
    Output o = chunk_output();
    for (size_t i = stack_.size(); i-- > 0;) {
      uint32_t cv[8];
      o.chaining_value(cv);
      Output parent;
      memcpy(parent.cv, IV, 32);
      memcpy(parent.block, stack_[i].w, 32);
      memcpy(parent.block + 8, cv, 32);
      parent.counter = 0;
      parent.block_len = 64;
      parent.flags = PARENT;
      o = parent;
    }
    uint64_t ctr = 0;
    while (out_len) {
      uint32_t w[16];
      compress(o.cv, o.block, ctr++, o.block_len, o.flags | ROOT, w);
      size_t take = std::min<size_t>(64, out_len);
      memcpy(out, w, take);  // little-endian host
      out += take;
      out_len -= take;
    }
  }
  static void hash(const uint8_t* p, size_t n, uint8_t out[32]) {
    Blake3 h;
    h.update(p, n);
    h.finalize(out, 32);
  }

 private:
  struct CV {
    uint32_t w[8];
  };
  struct Output {
    uint32_t cv[8], block[16];
    uint64_t counter;
    uint32_t block_len, flags;
    void chaining_value(uint32_t out8[8]) const {
      uint32_t o[16];
      compress(cv, block, counter, block_len, flags, o);
      memcpy(out8, o, 32);
    }
  };
  std::vector<CV> stack_;
  uint32_t chunk_cv_[8];
  uint8_t buf_[64];
  size_t buf_len_, chunk_len_;
  uint64_t chunk_counter_, total_chunks_ = 0;
  int blocks_done_;

  static uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
  static void g(uint32_t* s, int a, int b, int c, int d, uint32_t mx, uint32_t my) {
    s[a] = s[a] + s[b] + mx;
    s[d] = rotr(s[d] ^ s[a], 16);
    s[c] = s[c] + s[d];
    s[b] = rotr(s[b] ^ s[c], 12);
    s[a] = s[a] + s[b] + my;
    s[d] = rotr(s[d] ^ s[a], 8);
    s[c] = s[c] + s[d];
    s[b] = rotr(s[b] ^ s[c], 7);
  }
  static void compress(const uint32_t cv[8], const uint32_t block[16], uint64_t counter, uint32_t block_len,
                       uint32_t flags, uint32_t out[16]) {
    static const int PERM[16] = {2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8};
    uint32_t s[16] = {cv[0], cv[1], cv[2], cv[3], cv[4], cv[5], cv[6], cv[7], IV[0], IV[1], IV[2], IV[3],
                      (uint32_t)counter, (uint32_t)(counter >> 32), block_len, flags};
    uint32_t m[16];
    memcpy(m, block, 64);
    for (int r = 0; r < 7; r++) {
      g(s, 0, 4, 8, 12, m[0], m[1]);
      g(s, 1, 5, 9, 13, m[2], m[3]);
      g(s, 2, 6, 10, 14, m[4], m[5]);
      g(s, 3, 7, 11, 15, m[6], m[7]);
      g(s, 0, 5, 10, 15, m[8], m[9]);
      g(s, 1, 6, 11, 12, m[10], m[11]);
      g(s, 2, 7, 8, 13, m[12], m[13]);
      g(s, 3, 4, 9, 14, m[14], m[15]);
      uint32_t t[16];
      for (int i = 0; i < 16; i++) t[i] = m[PERM[i]];
      memcpy(m, t, 64);
    }
    for (int i = 0; i < 8; i++) {
      out[i] = s[i] ^ s[i + 8];
      out[i + 8] = s[i + 8] ^ cv[i];
    }
  }
  static void words(const uint8_t b[64], uint32_t w[16]) { memcpy(w, b, 64); }
  void start_chunk(uint64_t counter) {
    memcpy(chunk_cv_, IV, 32);
    memset(buf_, 0, 64);
    buf_len_ = 0;
    chunk_len_ = 0;
    blocks_done_ = 0;
    chunk_counter_ = counter;
  }
  uint32_t block_flags() const { return blocks_done_ == 0 ? (uint32_t)CHUNK_START : 0u; }
  Output chunk_output() const {
    Output o;
    memcpy(o.cv, chunk_cv_, 32);
    words(buf_, o.block);
    o.counter = chunk_counter_;
    o.block_len = (uint32_t)buf_len_;
    o.flags = block_flags() | CHUNK_END;
    return o;
  }
  void push_cv(const uint32_t cv_in[8], uint64_t total_chunks) {
    uint32_t cv[8];
    memcpy(cv, cv_in, 32);
    while ((total_chunks & 1) == 0) {  // merge completed subtrees
      uint32_t blk[16], out[16];
      memcpy(blk, stack_.back().w, 32);
      memcpy(blk + 8, cv, 32);
      stack_.pop_back();
      compress(IV, blk, 0, 64, PARENT, out);
      memcpy(cv, out, 32);
      total_chunks >>= 1;
    }
    CV e;
    memcpy(e.w, cv, 32);
    stack_.push_back(e);
  }
};

}  // namespace orc
