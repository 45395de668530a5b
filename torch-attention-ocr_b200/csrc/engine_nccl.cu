// engine_nccl.cu — the native data-parallel exchange: NCCL all-reduces (NVLink 5 / NVSwitch) issued by the step
// scheduler itself, on its own streams, with no host runtime in the loop — so a data-parallel step is one fixed
// launch sequence like the single-device step and is captured into the same whole-step CUDA graph.
//
// NCCL is bound at run time (dlopen of libnccl.so.2, already mapped in a torch process): libaocr.so has no link-time
// dependency on it and single-device use never touches it.  Bootstrap: rank 0 calls aocr_dp_unique_id, the host
// runtime ships the 128 bytes to every rank (aocr/dist.py uses torch.distributed for that), every rank calls
// aocr_dp_init.  The exchange points are those of the hook (include/aocr.h: aocr_set_allreduce):
//   kind 0  batch-norm statistics, ordered on the engine stream
//   kind 1  gradient bucket [first group, last group], on the communication stream, overlapping later backward work
//   kind 2  join before clip + SGD
#include <dlfcn.h>

#include "engine.h"

namespace aocr {

namespace {

// the few NCCL entry points used, by their public C signatures (nccl.h); handles are opaque here
struct NcclUniqueId { char internal[128]; };
typedef int (*fn_get_unique_id)(NcclUniqueId*);
typedef int (*fn_comm_init_rank)(void** comm, int nranks, NcclUniqueId id, int rank);
typedef int (*fn_comm_destroy)(void* comm);
typedef int (*fn_comm_split)(void* comm, int color, int key, void** newcomm, void* config);
typedef int (*fn_all_reduce)(const void* send, void* recv, size_t count, int dtype, int op, void* comm, cudaStream_t st);
typedef const char* (*fn_get_error_string)(int);
constexpr int kNcclFloat32 = 7, kNcclSum = 0;   // ncclDataType_t / ncclRedOp_t values of nccl.h (stable since 2.0)

struct NcclApi {
  void* lib = nullptr;
  fn_get_unique_id get_unique_id = nullptr;
  fn_comm_init_rank comm_init_rank = nullptr;
  fn_comm_destroy comm_destroy = nullptr;
  fn_comm_split comm_split = nullptr;       // optional (NCCL >= 2.18)
  fn_all_reduce all_reduce = nullptr;
  fn_get_error_string error_string = nullptr;
};

NcclApi& nccl_api() {
  static NcclApi api;
  if (api.lib) return api;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.lib) break;
  }
  AOCR_CHECK(api.lib != nullptr, "data parallelism needs NCCL: libnccl.so.2 not found (import torch first, or set LD_LIBRARY_PATH)");
  api.get_unique_id = (fn_get_unique_id)dlsym(api.lib, "ncclGetUniqueId");
  api.comm_init_rank = (fn_comm_init_rank)dlsym(api.lib, "ncclCommInitRank");
  api.comm_destroy = (fn_comm_destroy)dlsym(api.lib, "ncclCommDestroy");
  api.comm_split = (fn_comm_split)dlsym(api.lib, "ncclCommSplit");
  api.all_reduce = (fn_all_reduce)dlsym(api.lib, "ncclAllReduce");
  api.error_string = (fn_get_error_string)dlsym(api.lib, "ncclGetErrorString");
  AOCR_CHECK(api.get_unique_id && api.comm_init_rank && api.comm_destroy && api.all_reduce, "libnccl lacks a required symbol");
  return api;
}

void nccl_check(int rc, const char* what) {
  if (rc == 0) return;
  NcclApi& a = nccl_api();
  std::string msg = std::string("NCCL failure in ") + what + ": " + (a.error_string ? a.error_string(rc) : "?");
  throw CudaError(msg);
}

}  // namespace

void dp_unique_id(void* out128) {
  NcclUniqueId id;
  nccl_check(nccl_api().get_unique_id(&id), "ncclGetUniqueId");
  memcpy(out128, &id, sizeof(id));
}

void Engine::dp_init(const void* id128) {
  AOCR_CHECK(cfg.dp_world > 1, "aocr_dp_init: the handle was created with dp_world <= 1");
  AOCR_CHECK(nccl_comm_ == nullptr, "aocr_dp_init: already initialised");
  AOCR_CUDA(cudaSetDevice(device_));
  NcclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  nccl_check(nccl_api().comm_init_rank(&nccl_comm_, cfg.dp_world, id, cfg.dp_rank), "ncclCommInitRank");
  // A second communicator for the batch-norm statistics: operations on ONE communicator are serialised in issue
  // order even across streams, so a tiny statistics all-reduce in the CNN backward would wait for the 80 MB decoder
  // gradient bucket that is in flight on the communication stream.
  nccl_comm_stat_ = nccl_comm_;
  if (nccl_api().comm_split && !getenv("AOCR_DP_ONE_COMM"))
    nccl_check(nccl_api().comm_split(nccl_comm_, 0, cfg.dp_rank, &nccl_comm_stat_, nullptr), "ncclCommSplit");
  AOCR_CUDA(cudaStreamCreateWithFlags(&comm_st_, cudaStreamNonBlocking));
  for (int i = 0; i < 4; i++) AOCR_CUDA(cudaEventCreateWithFlags(&comm_ev_[i], cudaEventDisableTiming));
  // one warm-up collective outside any capture: NCCL sets up its channels / buffers on first use
  nccl_check(nccl_api().all_reduce(d_sumsq, d_sumsq, 1, kNcclFloat32, kNcclSum, nccl_comm_, ctx_.st), "ncclAllReduce");
  nccl_check(nccl_api().all_reduce(d_sumsq, d_sumsq, 1, kNcclFloat32, kNcclSum, nccl_comm_stat_, ctx_.st), "ncclAllReduce");
  AOCR_CUDA(cudaStreamSynchronize(ctx_.st));
}

void Engine::dp_shutdown() {
  if (!nccl_comm_) return;
  if (comm_st_) { cudaStreamSynchronize(comm_st_); cudaStreamDestroy(comm_st_); comm_st_ = nullptr; }
  for (int i = 0; i < 4; i++) if (comm_ev_[i]) { cudaEventDestroy(comm_ev_[i]); comm_ev_[i] = nullptr; }
  if (nccl_comm_stat_ && nccl_comm_stat_ != nccl_comm_) nccl_api().comm_destroy(nccl_comm_stat_);
  nccl_api().comm_destroy(nccl_comm_);
  nccl_comm_ = nullptr; nccl_comm_stat_ = nullptr;
}

// the three exchange kinds, native flavour (the hook flavour lives in engine_dec.cu)
void Engine::dp_allreduce(float* buf, int64_t n, int kind) {
  NcclApi& a = nccl_api();
  if (kind == 0) {
    nccl_check(a.all_reduce(buf, buf, (size_t)n, kNcclFloat32, kNcclSum, nccl_comm_stat_, ctx_.st), "ncclAllReduce(stat)");
    ctx_.launches++;
  } else if (kind == 1) {
    cudaEvent_t ev = comm_ev_[comm_ev_next_++ % 3];
    AOCR_CUDA(cudaEventRecord(ev, ctx_.st));            // the bucket is complete on the current lane's stream
    AOCR_CUDA(cudaStreamWaitEvent(comm_st_, ev, 0));
    nccl_check(a.all_reduce(buf, buf, (size_t)n, kNcclFloat32, kNcclSum, nccl_comm_, comm_st_), "ncclAllReduce(bucket)");
    ctx_.launches++;
    comm_pending_ = true;
  } else {
    if (comm_pending_) {
      AOCR_CUDA(cudaEventRecord(comm_ev_[3], comm_st_));
      AOCR_CUDA(cudaStreamWaitEvent(ctx_.st, comm_ev_[3], 0));
    }
    comm_pending_ = false;
  }
}

}  // namespace aocr
