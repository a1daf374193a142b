// The job's single exchange step: sum of the per-position accumulators of all GPUs over
// NCCL (NVLink / NVSwitch).  Reference equivalent: the dict accumulation over ALL reads in sum_handler
// (bin/DeepMod_scripts/myDetect.py:1089-1100) and the offline merge of per-run summaries in
// DeepMod_tools/sum_chr_mod.py:36-63.  Integer sums of packed cells: bit-exact in any order.
//
// NCCL is loaded at first use with dlopen("libnccl.so.2") -- the copy a host process already
// holds (e.g. the one PyTorch bundles) if there is one -- so the library itself loads on any
// box and only these entry points need NCCL.
#include "dm_common.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <mutex>

namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string why;
};

NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      api.handle = dlopen(n, RTLD_NOW | RTLD_NOLOAD);          // already in the process (torch's copy)?
      if (api.handle) break;
    }
    for (const char* n : names) {
      if (api.handle) break;
      api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    }
    if (!api.handle) { api.why = std::string("cannot load libnccl.so.2: ") + dlerror(); return; }
    auto sym = [&](const char* name) {
      void* p = dlsym(api.handle, name);
      if (!p && api.why.empty()) api.why = std::string("libnccl lacks ") + name;
      return p;
    };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(sym("ncclCommInitAll"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
  });
  return &api;
}

int nccl_fail(dm_ctx* ctx, const char* what, const NcclApi* api, ncclResult_t r) {
  dm_set_error(ctx, std::string(what) + ": " + (api->GetErrorString ? api->GetErrorString(r) : "NCCL error"));
  return DM_ERR_NCCL;
}

#define DM_NCCL(ctx, api, call)                                      \
  do {                                                               \
    ncclResult_t r__ = (call);                                       \
    if (r__ != ncclSuccess) return nccl_fail(ctx, #call, api, r__);  \
  } while (0)

}  // namespace

void dm_reduce_release(dm_ctx* ctx) {
  if (ctx->nccl_comm) {
    NcclApi* api = nccl_api();
    if (api->CommDestroy) api->CommDestroy(static_cast<ncclComm_t>(ctx->nccl_comm));
    ctx->nccl_comm = nullptr;
    ctx->nccl_ranks = 0;
  }
}

extern "C" {

int dm_reduce_unique_id(uint8_t id_out[128]) {
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  if (!id_out) return DM_ERR_ARG;
  NcclApi* api = nccl_api();
  if (!api->why.empty()) { dm_set_error(nullptr, "dm_reduce_unique_id: " + api->why); return DM_ERR_NCCL; }
  ncclUniqueId id;
  DM_NCCL(nullptr, api, api->GetUniqueId(&id));
  memcpy(id_out, &id, sizeof(id));
  return DM_OK;
}

int dm_reduce_comm(dm_ctx* ctx, const uint8_t* id, int rank, int n_ranks) {
  if (!ctx || n_ranks < 1 || rank < 0 || rank >= n_ranks) return DM_ERR_ARG;
  if (!ctx->cells) { dm_set_error(ctx, "dm_reduce_comm: dm_set_genome not called"); return DM_ERR_STATE; }
  if (n_ranks > (int)DM_CELL_DEL_MASK) { dm_set_error(ctx, "dm_reduce_comm: at most 255 ranks"); return DM_ERR_ARG; }
  DM_CUDA(ctx, cudaSetDevice(ctx->device));
  ctx->reduce_ms = 0.f;
  if (n_ranks == 1) return DM_OK;
  NcclApi* api = nccl_api();
  if (!api->why.empty()) { dm_set_error(ctx, "dm_reduce_comm: " + api->why); return DM_ERR_NCCL; }
  if (ctx->nccl_comm && (ctx->nccl_ranks != n_ranks || ctx->nccl_rank != rank || id != nullptr)) dm_reduce_release(ctx);
  if (!ctx->nccl_comm) {
    if (!id) { dm_set_error(ctx, "dm_reduce_comm: no communicator yet, an id is required"); return DM_ERR_ARG; }
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    ncclComm_t comm = nullptr;
    DM_NCCL(ctx, api, api->CommInitRank(&comm, n_ranks, uid, rank));
    ctx->nccl_comm = comm;
    ctx->nccl_rank = rank;
    ctx->nccl_ranks = n_ranks;
  }
  ncclComm_t comm = static_cast<ncclComm_t>(ctx->nccl_comm);
  cudaStream_t s = ctx->stream;
  // every rank must take the same decision: the largest coverage anywhere decides whether the sum can overflow
  unsigned long long mx = 0;
  int rc = dm_hist_max_cov(ctx, &mx);
  if (rc != DM_OK) return rc;
  unsigned long long* mx_d = nullptr;
  DM_CUDA(ctx, cudaMalloc(&mx_d, sizeof(unsigned long long)));
  cudaMemcpyAsync(mx_d, &mx, sizeof(mx), cudaMemcpyHostToDevice, s);
  ncclResult_t r = api->AllReduce(mx_d, mx_d, 1, ncclUint64, ncclMax, comm, s);
  if (r == ncclSuccess) cudaMemcpyAsync(&mx, mx_d, sizeof(mx), cudaMemcpyDeviceToHost, s);
  cudaError_t e = cudaStreamSynchronize(s);
  cudaFree(mx_d);
  if (r != ncclSuccess) return nccl_fail(ctx, "ncclAllReduce(max coverage)", api, r);
  if (e != cudaSuccess) { dm_set_error(ctx, std::string("dm_reduce_comm: ") + cudaGetErrorString(e)); return DM_ERR_CUDA; }
  if (mx * (unsigned long long)n_ranks > DM_CELL_MASK) {
    dm_set_error(ctx, "dm_reduce_comm: merged coverage could exceed the per-position counter (2^28 - 1); nothing was summed");
    return DM_ERR_OVERFLOW;
  }
  DM_CUDA(ctx, cudaEventRecord(ctx->ev0, s));
  DM_NCCL(ctx, api, api->AllReduce(ctx->cells, ctx->cells, (size_t)ctx->n_cells, ncclUint64, ncclSum, comm, s));
  rc = dm_hist_normalise_flags(ctx);
  if (rc != DM_OK) return rc;
  DM_CUDA(ctx, cudaEventRecord(ctx->ev3, s));
  DM_CUDA(ctx, cudaStreamSynchronize(s));
  DM_CUDA(ctx, cudaEventElapsedTime(&ctx->reduce_ms, ctx->ev0, ctx->ev3));
  return DM_OK;
}

int dm_reduce(dm_ctx** ctxs, int n) {
  if (!ctxs || n < 1) return DM_ERR_ARG;
  for (int i = 0; i < n; ++i) {
    if (!ctxs[i]) return DM_ERR_ARG;
    if (!ctxs[i]->cells) { dm_set_error(ctxs[i], "dm_reduce: dm_set_genome not called"); return DM_ERR_STATE; }
    if (ctxs[i]->n_cells != ctxs[0]->n_cells) { dm_set_error(ctxs[0], "dm_reduce: contexts hold different genomes"); return DM_ERR_ARG; }
    for (int j = 0; j < i; ++j)
      if (ctxs[j]->device == ctxs[i]->device) { dm_set_error(ctxs[0], "dm_reduce: two contexts on one device"); return DM_ERR_ARG; }
  }
  dm_ctx* c0 = ctxs[0];
  c0->reduce_ms = 0.f;
  if (n == 1) return DM_OK;
  if (n > (int)DM_CELL_DEL_MASK) { dm_set_error(c0, "dm_reduce: at most 255 contexts"); return DM_ERR_ARG; }
  NcclApi* api = nccl_api();
  if (!api->why.empty()) { dm_set_error(c0, "dm_reduce: " + api->why); return DM_ERR_NCCL; }
  unsigned long long mx = 0;
  for (int i = 0; i < n; ++i) {
    unsigned long long m = 0;
    DM_CUDA(ctxs[i], cudaSetDevice(ctxs[i]->device));
    int rc = dm_hist_max_cov(ctxs[i], &m);
    if (rc != DM_OK) return rc;
    mx = m > mx ? m : mx;
  }
  if (mx * (unsigned long long)n > DM_CELL_MASK) {
    dm_set_error(c0, "dm_reduce: merged coverage could exceed the per-position counter (2^28 - 1); nothing was summed");
    return DM_ERR_OVERFLOW;
  }
  std::vector<int> devs((size_t)n);
  std::vector<ncclComm_t> comms((size_t)n, nullptr);
  for (int i = 0; i < n; ++i) devs[i] = ctxs[i]->device;
  DM_NCCL(c0, api, api->CommInitAll(comms.data(), n, devs.data()));
  int rc = DM_OK;
  for (int i = 0; i < n; ++i) { cudaSetDevice(devs[i]); cudaEventRecord(ctxs[i]->ev0, ctxs[i]->stream); }
  ncclResult_t r = api->GroupStart();
  for (int i = 0; i < n && r == ncclSuccess; ++i)
    r = api->AllReduce(ctxs[i]->cells, ctxs[i]->cells, (size_t)ctxs[i]->n_cells, ncclUint64, ncclSum, comms[i], ctxs[i]->stream);
  ncclResult_t r2 = api->GroupEnd();
  if (r == ncclSuccess) r = r2;
  if (r != ncclSuccess) rc = nccl_fail(c0, "dm_reduce: grouped ncclAllReduce", api, r);
  for (int i = 0; i < n; ++i) {
    cudaSetDevice(devs[i]);
    if (rc == DM_OK) rc = dm_hist_normalise_flags(ctxs[i]);
    cudaEventRecord(ctxs[i]->ev3, ctxs[i]->stream);
  }
  for (int i = 0; i < n; ++i) {
    cudaSetDevice(devs[i]);
    cudaError_t e = cudaStreamSynchronize(ctxs[i]->stream);
    if (e != cudaSuccess && rc == DM_OK) { dm_set_error(ctxs[i], std::string("dm_reduce: ") + cudaGetErrorString(e)); rc = DM_ERR_CUDA; }
    if (e == cudaSuccess) cudaEventElapsedTime(&ctxs[i]->reduce_ms, ctxs[i]->ev0, ctxs[i]->ev3);
  }
  for (int i = 0; i < n; ++i) api->CommDestroy(comms[i]);
  return rc;
}

int dm_reduce_finalize(dm_ctx* ctx) {
  if (!ctx) return DM_ERR_ARG;
  if (ctx->stream) DM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  dm_reduce_release(ctx);
  return DM_OK;
}

int dm_last_reduce_ms(const dm_ctx* ctx, float* ms) {
  if (!ctx || !ms) return DM_ERR_ARG;
  *ms = ctx->reduce_ms;
  return DM_OK;
}

}  // extern "C"
