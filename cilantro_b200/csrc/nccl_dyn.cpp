// NCCL binding, resolved at run time with dlopen (product code).
//
// The multi-GPU path is one process per GPU; each rank owns a shard of the source points and the
// only exchange per iteration is ONE ncclAllReduce(sum, f64) of <= 28 values (ICP) or K x 4 values
// (k-means) over NVLink/NVSwitch — latency-bound, issued on the context stream
// right behind the accumulation kernel. libnccl is loaded lazily so that single-GPU users need no
// NCCL at all; inside a torch process dlopen("libnccl.so.2") resolves to the copy torch already
// loaded.
#include "cb_internal.hpp"
#include <dlfcn.h>
#include <cstring>

namespace cb {

namespace {

struct NcclUniqueId {
  char internal[128];
};
typedef void* NcclComm;
typedef int (*fn_get_unique_id)(NcclUniqueId*);
typedef int (*fn_comm_init_rank)(NcclComm*, int, NcclUniqueId, int);
typedef int (*fn_all_reduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t);
typedef int (*fn_comm_destroy)(NcclComm);
typedef const char* (*fn_get_error_string)(int);

constexpr int kNcclFloat64 = 8;  // ncclDataType_t::ncclFloat64 / ncclDouble
constexpr int kNcclUint32 = 3;   // ncclDataType_t::ncclUint32
constexpr int kNcclSum = 0;      // ncclRedOp_t::ncclSum

struct Api {
  void* handle = nullptr;
  fn_get_unique_id get_unique_id = nullptr;
  fn_comm_init_rank comm_init_rank = nullptr;
  fn_all_reduce all_reduce = nullptr;
  fn_comm_destroy comm_destroy = nullptr;
  fn_get_error_string get_error_string = nullptr;
};

Api* api() {
  static Api a;
  static bool tried = false;
  if (!tried) {
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      a.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (a.handle) break;
    }
    if (a.handle) {
      a.get_unique_id = (fn_get_unique_id)dlsym(a.handle, "ncclGetUniqueId");
      a.comm_init_rank = (fn_comm_init_rank)dlsym(a.handle, "ncclCommInitRank");
      a.all_reduce = (fn_all_reduce)dlsym(a.handle, "ncclAllReduce");
      a.comm_destroy = (fn_comm_destroy)dlsym(a.handle, "ncclCommDestroy");
      a.get_error_string = (fn_get_error_string)dlsym(a.handle, "ncclGetErrorString");
    }
  }
  if (!a.handle || !a.get_unique_id || !a.comm_init_rank || !a.all_reduce || !a.comm_destroy) return nullptr;
  return &a;
}

int fail(const char* what, int rc) {
  Api* a = api();
  set_error("NCCL %s failed: %s", what, (a && a->get_error_string) ? a->get_error_string(rc) : "unknown");
  return CB_ERR_NCCL;
}

}  // namespace

int nccl_unique_id(void* out128) {
  if (!out128) return CB_ERR_INVALID;
  Api* a = api();
  if (!a) {
    set_error("libnccl.so.2 could not be loaded");
    return CB_ERR_NCCL;
  }
  NcclUniqueId id;
  int rc = a->get_unique_id(&id);
  if (rc != 0) return fail("ncclGetUniqueId", rc);
  std::memcpy(out128, &id, sizeof(id));
  return CB_OK;
}

int nccl_init(cb_context* ctx, const void* id128, int rank, int world) {
  if (world == 1) {
    ctx->rank = 0;
    ctx->world = 1;
    return CB_OK;
  }
  Api* a = api();
  if (!a) {
    set_error("libnccl.so.2 could not be loaded");
    return CB_ERR_NCCL;
  }
  NcclUniqueId id;
  std::memcpy(&id, id128, sizeof(id));
  NcclComm comm = nullptr;
  int rc = a->comm_init_rank(&comm, world, id, rank);
  if (rc != 0) return fail("ncclCommInitRank", rc);
  ctx->nccl_comm = comm;
  ctx->rank = rank;
  ctx->world = world;
  return CB_OK;
}

int nccl_allreduce_sum_f64(cb_context* ctx, double* d_buf, size_t count) {
  if (ctx->world <= 1) return CB_OK;
  Api* a = api();
  if (!a || !ctx->nccl_comm) {
    set_error("communicator not initialised");
    return CB_ERR_NCCL;
  }
  int rc = a->all_reduce(d_buf, d_buf, count, kNcclFloat64, kNcclSum, (NcclComm)ctx->nccl_comm, ctx->stream);
  if (rc != 0) return fail("ncclAllReduce", rc);
  return CB_OK;
}

// Integer sum of 32-bit words: with every word non-zero on at most one rank this is an exact all-gather of bit patterns
// (used to replicate a cloud whose blocks were uploaded by different ranks: cb_cloud_create_replicated).
int nccl_allreduce_sum_u32(cb_context* ctx, uint32_t* d_buf, size_t count) {
  if (ctx->world <= 1) return CB_OK;
  Api* a = api();
  if (!a || !ctx->nccl_comm) {
    set_error("communicator not initialised");
    return CB_ERR_NCCL;
  }
  int rc = a->all_reduce(d_buf, d_buf, count, kNcclUint32, kNcclSum, (NcclComm)ctx->nccl_comm, ctx->stream);
  if (rc != 0) return fail("ncclAllReduce", rc);
  return CB_OK;
}

void nccl_destroy(cb_context* ctx) {
  if (!ctx->nccl_comm) return;
  Api* a = api();
  if (a) a->comm_destroy((NcclComm)ctx->nccl_comm);
  ctx->nccl_comm = nullptr;
  ctx->world = 1;
  ctx->rank = 0;
}

}  // namespace cb
