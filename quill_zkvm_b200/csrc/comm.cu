// Multi-GPU layer: one process per GPU, NCCL over NVLink for the two tiny exchanges the path has:
//   * MSM: bases/scalars are split by index range; each rank runs the whole single-GPU pipeline to one XYZZ partial
//     sum, the partial sums are all-gathered as raw limbs (128 B per rank) and added with the group law on every rank;
//   * sumcheck: tables are split by the top variables; per round the ranks all-gather (deg+1) partial sums (sumcheck.cu).
// NCCL is bound at run time with dlopen (the image carries it inside the torch wheel), so building this library needs
// no NCCL headers and a single-GPU user never loads it.  Group elements are never reduced with ncclSum.
#include <dlfcn.h>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include "ctx.cuh"
#include "msm.cuh"

namespace {

// minimal NCCL surface (stable ABI since 2.x)
typedef struct {
  char internal[128];
} ncclUniqueIdT;
typedef int ncclResultT;
typedef ncclResultT (*fn_GetUniqueId)(ncclUniqueIdT*);
typedef ncclResultT (*fn_CommInitRank)(ncclComm**, int, ncclUniqueIdT, int);
typedef ncclResultT (*fn_CommDestroy)(ncclComm*);
typedef ncclResultT (*fn_AllGather)(const void*, void*, size_t, int /*ncclDataType_t*/, ncclComm*, cudaStream_t);
typedef const char* (*fn_GetErrorString)(ncclResultT);
constexpr int NCCL_UINT8 = 1;  // ncclUint8

struct NcclApi {
  void* handle = nullptr;
  fn_GetUniqueId GetUniqueId = nullptr;
  fn_CommInitRank CommInitRank = nullptr;
  fn_CommDestroy CommDestroy = nullptr;
  fn_AllGather AllGather = nullptr;
  fn_GetErrorString GetErrorString = nullptr;
  bool ok() const { return handle && GetUniqueId && CommInitRank && AllGather; }
};

NcclApi& nccl() {
  static NcclApi api;
  if (api.handle) return api;
  const char* env = getenv("QZ_NCCL_LIB");
  const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    if (!n || !*n) continue;
    api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
  }
  if (!api.handle) return api;
  api.GetUniqueId = (fn_GetUniqueId)dlsym(api.handle, "ncclGetUniqueId");
  api.CommInitRank = (fn_CommInitRank)dlsym(api.handle, "ncclCommInitRank");
  api.CommDestroy = (fn_CommDestroy)dlsym(api.handle, "ncclCommDestroy");
  api.AllGather = (fn_AllGather)dlsym(api.handle, "ncclAllGather");
  api.GetErrorString = (fn_GetErrorString)dlsym(api.handle, "ncclGetErrorString");
  return api;
}

}  // namespace

namespace qz {

// all-gather `bytes` from every rank into recv (rank-major) on the context's stream
int comm_allgather(qz_ctx* ctx, const void* send, void* recv, size_t bytes) {
  if (ctx->nranks == 1) {
    if (send != recv) QZ_CUDA(ctx, cudaMemcpyAsync(recv, send, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return QZ_OK;
  }
  NcclApi& api = nccl();
  if (!api.ok() || !ctx->comm) return ctx->fail(QZ_ERR_NCCL, "communicator not initialised");
  ncclResultT r = api.AllGather(send, recv, bytes, NCCL_UINT8, ctx->comm, ctx->stream);
  if (r != 0) return ctx->fail(QZ_ERR_NCCL, api.GetErrorString ? api.GetErrorString(r) : "ncclAllGather failed");
  return QZ_OK;
}

void comm_destroy(qz_ctx* ctx) {
  if (ctx->comm && nccl().CommDestroy) nccl().CommDestroy(ctx->comm);
  ctx->comm = nullptr;
}

}  // namespace qz

using namespace qz;

extern "C" {

int qz_comm_unique_id(uint8_t out_id[128]) {
  if (!out_id) return QZ_ERR_INVALID_ARG;
  NcclApi& api = nccl();
  if (!api.ok()) return QZ_ERR_NCCL;
  ncclUniqueIdT id;
  if (api.GetUniqueId(&id) != 0) return QZ_ERR_NCCL;
  memcpy(out_id, id.internal, 128);
  return QZ_OK;
}

int qz_comm_init(qz_ctx* ctx, const uint8_t unique_id[128], int rank, int nranks) {
  if (!ctx || !unique_id || nranks < 1 || rank < 0 || rank >= nranks) return QZ_ERR_INVALID_ARG;
  if (nranks & (nranks - 1)) return ctx->fail(QZ_ERR_INVALID_ARG, "rank count must be a power of two");
  QZ_CUDA(ctx, cudaSetDevice(ctx->device));
  ctx->rank = rank;
  ctx->nranks = nranks;
  if (nranks == 1) return QZ_OK;
  NcclApi& api = nccl();
  if (!api.ok()) return ctx->fail(QZ_ERR_NCCL, "libnccl.so.2 not found (set QZ_NCCL_LIB)");
  ncclUniqueIdT id;
  memcpy(id.internal, unique_id, 128);
  ncclResultT r = api.CommInitRank(&ctx->comm, nranks, id, rank);
  if (r != 0) return ctx->fail(QZ_ERR_NCCL, api.GetErrorString ? api.GetErrorString(r) : "ncclCommInitRank failed");
  return QZ_OK;
}

int qz_msm_sharded(qz_ctx* ctx, const qz_srs* srs, const void* scalars, size_t n_scalars, int on_device,
                   uint8_t out_xy[64]) {
  if (!ctx || !srs || !out_xy || (n_scalars && !scalars)) return QZ_ERR_INVALID_ARG;
  QZ_CUDA(ctx, cudaSetDevice(ctx->device));
  ctx->arena_reset();
  const size_t n = std::min(n_scalars, srs->n);
  cudaStream_t st = ctx->stream;
  QZ_CUDA(ctx, cudaEventRecord(ctx->ev_call0, st));
  QZ_CUDA(ctx, cudaEventRecord(ctx->ev_k0, st));
  QZ_CUDA(ctx, cudaEventRecord(ctx->ev_k1, st));
  const uint4* sdev = (const uint4*)scalars;
  if (!on_device && n) {
    void* p = ctx->arena_alloc(32 * n);
    if (!p) return ctx->fail(QZ_ERR_ALLOC, "scalars");
    QZ_CUDA(ctx, cudaMemcpyAsync(p, scalars, 32 * n, cudaMemcpyHostToDevice, st));
    sdev = (const uint4*)p;
  }
  uint8_t* mine = (uint8_t*)ctx->arena_alloc(128);
  uint8_t* all = (uint8_t*)ctx->arena_alloc((size_t)128 * ctx->nranks);
  uint8_t* out_dev = (uint8_t*)ctx->arena_alloc(64);
  if (!mine || !all || !out_dev) return ctx->fail(QZ_ERR_ALLOC, "result");
  int rc = msm_device(ctx, srs, sdev, n, mine, nullptr);
  if (rc) return rc;
  rc = comm_allgather(ctx, mine, all, 128);
  if (rc) return rc;
  rc = msm_sum_points_launch(ctx, all, ctx->nranks, out_dev);
  if (rc) return rc;
  QZ_CUDA(ctx, cudaMemcpyAsync(out_xy, out_dev, 64, cudaMemcpyDeviceToHost, st));
  QZ_CUDA(ctx, cudaEventRecord(ctx->ev_call1, st));
  QZ_CUDA(ctx, cudaStreamSynchronize(st));
  cudaEventElapsedTime(&ctx->last_ms[0], ctx->ev_call0, ctx->ev_call1);
  cudaEventElapsedTime(&ctx->last_ms[1], ctx->ev_k0, ctx->ev_k1);
  return QZ_OK;
}

int qz_comm_allgather_host(qz_ctx* ctx, const void* send, void* recv, size_t bytes) {
  if (!ctx || !send || !recv) return QZ_ERR_INVALID_ARG;
  if (bytes == 0) return QZ_OK;
  QZ_CUDA(ctx, cudaSetDevice(ctx->device));
  ctx->arena_reset();
  cudaStream_t st = ctx->stream;
  uint8_t* mine = (uint8_t*)ctx->arena_alloc(bytes);
  uint8_t* all = (uint8_t*)ctx->arena_alloc(bytes * ctx->nranks);
  if (!mine || !all) return ctx->fail(QZ_ERR_ALLOC, "all-gather buffers");
  QZ_CUDA(ctx, cudaMemcpyAsync(mine, send, bytes, cudaMemcpyHostToDevice, st));
  int rc = comm_allgather(ctx, mine, all, bytes);
  if (rc) return rc;
  QZ_CUDA(ctx, cudaMemcpyAsync(recv, all, bytes * ctx->nranks, cudaMemcpyDeviceToHost, st));
  QZ_CUDA(ctx, cudaStreamSynchronize(st));
  return QZ_OK;
}

}  // extern "C"
