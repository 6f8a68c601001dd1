// Peer mailboxes: the multi-GPU exchanges of the proving path are a few hundred bytes per rank, so their
// cost is launch and protocol latency, not bandwidth.  Every rank owns a small mailbox in its HBM, maps every peer's
// mailbox into its address space through CUDA IPC once (qz_comm_init), and the kernel that produced a value stores it
// straight into all the peers' mailboxes over NVLink / NVSwitch and then raises a flag; the consumer kernel on each
// rank spins on its local flags.  No collective launch sits between the producer and the consumer.
#pragma once
#include <cstdint>
#include "ctx.cuh"
#include "sumcheck.cuh"

namespace qz {

constexpr int QZ_MAX_PEERS = 16;
constexpr int MBOX_SLOTS = 4;  // a sender can be at most one exchange ahead of a receiver; four slots leave slack

struct PeerSlot {
  Fr data[QZ_MAX_PEERS][SC_MAX_COEFFS];  // data[src]: the vector rank `src` sent
  uint32_t flag[QZ_MAX_PEERS][8];        // flag[src][0] = sequence number of the exchange whose data[src] is complete
};
struct PeerMailbox {
  PeerSlot slot[MBOX_SLOTS];
  uint32_t timed_out;  // set by a consumer that gave up waiting (a peer died): the host turns it into QZ_ERR_NCCL
  uint32_t pad_[7];
  // hand-over of a sharded sumcheck to its last rounds (sumcheck.cu sc_mid): every rank stores its shard of every table
  // here, in every rank's mailbox, rank order = index order.  Two copies, picked by the parity of the exchange number: a
  // rank can be at most one proof ahead of the slowest one (it cannot finish proof n + 1 without that rank's shards).
  Fr gather[2][SC_MAX_K][1 << SC_TAIL_LOG];
};

// ---- device side -------------------------------------------------------------------------------------------------------
QZ_DEV uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
QZ_DEV void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
QZ_DEV unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
QZ_DEV Fr ld_fresh(const Fr* p) {  // bypass L1: the line may hold what an earlier exchange left in this slot
  const uint4 a = __ldcv(reinterpret_cast<const uint4*>(p)), b = __ldcv(reinterpret_cast<const uint4*>(p) + 1);
  Fr r;
  r.v[0] = a.x, r.v[1] = a.y, r.v[2] = a.z, r.v[3] = a.w, r.v[4] = b.x, r.v[5] = b.y, r.v[6] = b.z, r.v[7] = b.w;
  return r;
}
constexpr unsigned long long PEER_WAIT_NS = 20ull * 1000 * 1000 * 1000;  // a peer that is 20 s late is gone

// Exchange `n` field elements per rank (n <= SC_MAX_COEFFS), called by every thread of a block of >= max(G, n)
// threads: vals[x] (shared memory, x < n) go to data[rank][x] of slot seq % MBOX_SLOTS in EVERY rank's mailbox, then
// this rank's flag is raised there; returns once all G flags of the local mailbox carry `seq`.  On return
// mine->slot[seq % MBOX_SLOTS].data[g][x] holds rank g's values (read them with ld_fresh).  Ends with a barrier.
QZ_DEV PeerSlot* peer_exchange(PeerMailbox* const* peers, int rank, int G, uint32_t seq, const Fr* vals, int n) {
  const int slot = (int)(seq % MBOX_SLOTS);
  for (int i = threadIdx.x; i < G * n; i += blockDim.x) {
    const int g = i / n, x = i % n;
    fp_store<FrParams>(reinterpret_cast<uint4*>(&peers[g]->slot[slot].data[rank][x]), vals[x]);
  }
  __threadfence_system();
  __syncthreads();
  PeerMailbox* mine = peers[rank];
  if ((int)threadIdx.x < G) {
    st_release_sys(&peers[threadIdx.x]->slot[slot].flag[rank][0], seq);
    const uint32_t* f = &mine->slot[slot].flag[threadIdx.x][0];
    volatile uint32_t* dead = &mine->timed_out;
    if (ld_acquire_sys(f) != seq && !*dead) {
      const unsigned long long t0 = global_timer_ns();
      for (unsigned int polls = 1; ld_acquire_sys(f) != seq; polls++) {
        if ((polls & 255u) == 0 && global_timer_ns() - t0 > PEER_WAIT_NS) {  // the clock costs more than a poll
          *dead = 1;
          break;
        }
      }
    }
  }
  __syncthreads();
  return &mine->slot[slot];
}

// all-gather `bytes` from every rank into recv (rank-major) on the context's stream (NCCL)
int comm_allgather(qz_ctx* ctx, const void* send, void* recv, size_t bytes);
// true when the peers' mailboxes are mapped (qz_comm_init succeeded in opening them and QZ_NO_P2P is unset)
inline bool comm_has_peers(const qz_ctx* ctx) { return ctx->nranks > 1 && ctx->peer_mbox_dev != nullptr; }

}  // namespace qz
