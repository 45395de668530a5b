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
typedef int (*fn_all_gather)(const void* send, void* recv, size_t sendcount, int dtype, void* comm, cudaStream_t st);
typedef const char* (*fn_get_error_string)(int);
constexpr int kNcclFloat32 = 7, kNcclSum = 0, kNcclInt8 = 0;   // ncclDataType_t / ncclRedOp_t values of nccl.h (stable since 2.0)

struct NcclApi {
  void* lib = nullptr;
  fn_get_unique_id get_unique_id = nullptr;
  fn_comm_init_rank comm_init_rank = nullptr;
  fn_comm_destroy comm_destroy = nullptr;
  fn_comm_split comm_split = nullptr;       // optional (NCCL >= 2.18)
  fn_all_reduce all_reduce = nullptr;
  fn_all_gather all_gather = nullptr;       // optional: bootstrap of the peer-memory statistics exchange
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
  api.all_gather = (fn_all_gather)dlsym(api.lib, "ncclAllGather");
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


// ---- one-shot all-reduce of a short vector through NVLink peer memory -------------------------------------------------
// The six batch-norm statistics reductions of a data-parallel step are <= 4 KB each and sit on the critical path of the
// CNN forward / backward: an NCCL all-reduce costs ~40 us there (launch + protocol latency), the payload nothing.  Here
// every rank stores its vector straight into a mailbox slot of every peer (st over NVLink), then a flag; each rank polls
// the flags in its OWN memory and sums the `world` vectors in rank order (identical result on every rank).  One kernel,
// one CTA, graph-capturable; mailbox halves alternate with the parity of the slot's use count, and a slot is reused
// only after every rank has passed through the other exchanges of a step, so a writer never overtakes a reader.
constexpr int kMboxSlots = 16;         // distinct exchanges per step (6 are used)
constexpr int kMboxMaxN = 1024;        // floats per exchange (2 x 512 channels)
constexpr int kMboxMaxWorld = 16;
__host__ __device__ inline size_t mbox_data_off(int par, int slot, int rank) {
  return ((size_t)(par * kMboxSlots + slot) * kMboxMaxWorld + rank) * kMboxMaxN;
}
__host__ __device__ inline size_t mbox_flag_off(int par, int slot, int rank) {      // in 4-byte words, after the data area
  return (size_t)2 * kMboxSlots * kMboxMaxWorld * kMboxMaxN + (size_t)(par * kMboxSlots + slot) * kMboxMaxWorld + rank;
}
constexpr size_t kMboxWords = (size_t)2 * kMboxSlots * kMboxMaxWorld * kMboxMaxN + (size_t)2 * kMboxSlots * kMboxMaxWorld;

__global__ void __launch_bounds__(1024) peer_allreduce_kernel(float* __restrict__ buf, int n, int rank, int world,
                                                              float* const* __restrict__ peers, unsigned* __restrict__ seqctr,
                                                              int slot, unsigned long long timeout_ns) {
  __shared__ unsigned seq_s;
  const int tid = threadIdx.x;
  if (tid == 0) seq_s = ++seqctr[slot];
  __syncthreads();
  const unsigned seq = seq_s;
  const int par = (int)(seq & 1u);
  if (tid < n) {
    const float v = buf[tid];
    for (int r = 0; r < world; r++) {
      volatile float* dst = peers[r] + mbox_data_off(par, slot, rank);
      dst[tid] = v;
    }
  }
  __threadfence_system();
  __syncthreads();
  if (tid < world) {       // my flag in rank `tid`'s mailbox: my data there is complete
    unsigned* f = reinterpret_cast<unsigned*>(peers[tid]) + mbox_flag_off(par, slot, rank);
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(seq) : "memory");
    // ... and wait for rank `tid`'s flag in MY mailbox
    const unsigned* g = reinterpret_cast<const unsigned*>(peers[rank]) + mbox_flag_off(par, slot, tid);
    unsigned v = 0;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(g) : "memory");
      if (v != seq) {
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > timeout_ns) {         // a peer is gone: fail the launch instead of hanging the device for ever
          printf("[aocr] peer statistics exchange timed out (rank %d waits for rank %d, slot %d, seq %u, saw %u)\n", rank, tid, slot, seq, v);
          __trap();
        }
      }
    } while (v != seq);
  }
  __syncthreads();
  if (tid < n) {
    float s = 0.f;
    for (int r = 0; r < world; r++) {
      const volatile float* src = peers[rank] + mbox_data_off(par, slot, r);
      s += src[tid];
    }
    buf[tid] = s;
  }
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
  dp_peer_setup();
}

// Peer mailboxes: allocate, exchange the IPC handles with one ncclAllGather, map every peer.  Any failure leaves the
// NCCL all-reduce in place (mbox_ stays nullptr on EVERY rank: the decision is made on gathered data only).
void Engine::dp_peer_setup() {
  NcclApi& a = nccl_api();
  if (getenv("AOCR_DP_PEER") && atoi(getenv("AOCR_DP_PEER")) == 0) return;
  if (!a.all_gather || cfg.dp_world > kMboxMaxWorld) return;
  const int world = cfg.dp_world, rank = cfg.dp_rank;
  struct Rec { cudaIpcMemHandle_t h; int ok; int pad[3]; };
  static_assert(sizeof(Rec) == 80, "record layout");
  float* mb = nullptr;
  Rec mine{};
  if (cudaMalloc(&mb, kMboxWords * sizeof(float)) == cudaSuccess && cudaMemset(mb, 0, kMboxWords * sizeof(float)) == cudaSuccess &&
      cudaIpcGetMemHandle(&mine.h, mb) == cudaSuccess)
    mine.ok = 1;
  cudaGetLastError();
  Rec *d_send = nullptr, *d_recv = nullptr;
  AOCR_CUDA(cudaMalloc(&d_send, sizeof(Rec)));
  AOCR_CUDA(cudaMalloc(&d_recv, sizeof(Rec) * world));
  AOCR_CUDA(cudaMemcpy(d_send, &mine, sizeof(Rec), cudaMemcpyHostToDevice));
  nccl_check(a.all_gather(d_send, d_recv, sizeof(Rec), kNcclInt8, nccl_comm_, ctx_.st), "ncclAllGather(ipc handles)");
  AOCR_CUDA(cudaStreamSynchronize(ctx_.st));
  std::vector<Rec> all(world);
  AOCR_CUDA(cudaMemcpy(all.data(), d_recv, sizeof(Rec) * world, cudaMemcpyDeviceToHost));
  cudaFree(d_send); cudaFree(d_recv);
  bool ok = true;
  for (int r = 0; r < world; r++) ok = ok && all[r].ok;
  std::vector<float*> ptrs(world, nullptr);
  if (ok) {
    for (int r = 0; r < world && ok; r++) {
      if (r == rank) { ptrs[r] = mb; continue; }
      void* p = nullptr;
      if (cudaIpcOpenMemHandle(&p, all[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
      peer_mapped_.push_back(p);
      ptrs[r] = (float*)p;
    }
  }
  // second round: every rank must have mapped every peer, or nobody uses the mailboxes
  int good = ok ? 1 : 0, *d_good = nullptr;
  AOCR_CUDA(cudaMalloc(&d_good, sizeof(float)));
  float gf = (float)good;
  AOCR_CUDA(cudaMemcpy(d_good, &gf, sizeof(float), cudaMemcpyHostToDevice));
  nccl_check(a.all_reduce(d_good, d_good, 1, kNcclFloat32, kNcclSum, nccl_comm_, ctx_.st), "ncclAllReduce(peer setup)");
  AOCR_CUDA(cudaStreamSynchronize(ctx_.st));
  AOCR_CUDA(cudaMemcpy(&gf, d_good, sizeof(float), cudaMemcpyDeviceToHost));
  cudaFree(d_good);
  if ((int)(gf + 0.5f) != world) {
    for (void* p : peer_mapped_) cudaIpcCloseMemHandle(p);
    peer_mapped_.clear();
    if (mb) cudaFree(mb);
    return;
  }
  mbox_ = mb;
  AOCR_CUDA(cudaMalloc(&d_peer_mbox_, sizeof(float*) * world));
  AOCR_CUDA(cudaMemcpy(d_peer_mbox_, ptrs.data(), sizeof(float*) * world, cudaMemcpyHostToDevice));
  AOCR_CUDA(cudaMalloc(&d_mbox_seq_, sizeof(unsigned) * kMboxSlots));
  AOCR_CUDA(cudaMemset(d_mbox_seq_, 0, sizeof(unsigned) * kMboxSlots));
  if (getenv("AOCR_DP_VERBOSE")) fprintf(stderr, "[aocr] rank %d: batch-norm statistics go through peer mailboxes (%d ranks)\n", rank, world);
}

void Engine::dp_shutdown() {
  if (!nccl_comm_) return;
  cudaDeviceSynchronize();
  for (void* p : peer_mapped_) cudaIpcCloseMemHandle(p);
  peer_mapped_.clear();
  if (mbox_) { cudaFree(mbox_); mbox_ = nullptr; }
  if (d_peer_mbox_) { cudaFree(d_peer_mbox_); d_peer_mbox_ = nullptr; }
  if (d_mbox_seq_) { cudaFree(d_mbox_seq_); d_mbox_seq_ = nullptr; }
  if (comm_st_) { cudaStreamSynchronize(comm_st_); cudaStreamDestroy(comm_st_); comm_st_ = nullptr; }
  for (int i = 0; i < 4; i++) if (comm_ev_[i]) { cudaEventDestroy(comm_ev_[i]); comm_ev_[i] = nullptr; }
  if (nccl_comm_stat_ && nccl_comm_stat_ != nccl_comm_) nccl_api().comm_destroy(nccl_comm_stat_);
  nccl_api().comm_destroy(nccl_comm_);
  nccl_comm_ = nullptr; nccl_comm_stat_ = nullptr;
}

bool Engine::dp_peer_allreduce(float* buf, int64_t n) {
  if (!mbox_ || n > kMboxMaxN || mbox_slot_next_ >= kMboxSlots) return false;
  // ranks may reach an exchange far apart (one of them saving a checkpoint, loading data): wait as long as a collective
  // library's watchdog would (10 minutes; AOCR_DP_PEER_TIMEOUT_S overrides), then fail the launch
  static const unsigned long long timeout_ns =
      (unsigned long long)(getenv("AOCR_DP_PEER_TIMEOUT_S") ? atof(getenv("AOCR_DP_PEER_TIMEOUT_S")) : 600.0) * 1000000000ull;
  peer_allreduce_kernel<<<1, 1024, 0, ctx_.st>>>(buf, (int)n, cfg.dp_rank, cfg.dp_world, d_peer_mbox_, d_mbox_seq_, mbox_slot_next_++,
                                                 timeout_ns);
  AOCR_LAUNCH_CHECK(ctx_);
  return true;
}

// the three exchange kinds, native flavour (the hook flavour lives in engine_dec.cu)
void Engine::dp_allreduce(float* buf, int64_t n, int kind) {
  NcclApi& a = nccl_api();
  if (kind == 0) {
    if (dp_peer_allreduce(buf, n)) return;
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
