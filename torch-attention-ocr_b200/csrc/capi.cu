// capi.cu — the extern "C" surface of libaocr.so (include/aocr.h).  Exceptions are converted to status
// codes here; nothing C++ crosses the ABI.
#include <string.h>

#include <new>

#include "engine.h"

using aocr::Engine;

struct aocr_handle {
  Engine* eng = nullptr;
  std::string err;
};

static thread_local std::string g_create_error;

#define AOCR_API_BEGIN(h)                          \
  if (!(h) || !(h)->eng) return AOCR_ERR_INVALID; \
  try {
#define AOCR_API_END(h)                                                \
  }                                                                    \
  catch (const aocr::InvalidError& e) { (h)->err = e.what(); return AOCR_ERR_INVALID; } \
  catch (const aocr::CudaError& e) { (h)->err = e.what(); return AOCR_ERR_CUDA; }       \
  catch (const std::exception& e) { (h)->err = e.what(); return AOCR_ERR_STATE; }       \
  return AOCR_OK;

extern "C" {

int aocr_create(const aocr_config* cfg, int device, aocr_handle** out) {
  if (!cfg || !out) { g_create_error = "null argument"; return AOCR_ERR_INVALID; }
  *out = nullptr;
  try {
    aocr_handle* h = new aocr_handle();
    try {
      h->eng = new Engine(*cfg, device);
    } catch (...) {
      delete h;
      throw;
    }
    *out = h;
  } catch (const aocr::InvalidError& e) { g_create_error = e.what(); return AOCR_ERR_INVALID; }
  catch (const aocr::CudaError& e) { g_create_error = e.what(); return AOCR_ERR_CUDA; }
  catch (const std::exception& e) { g_create_error = e.what(); return AOCR_ERR_STATE; }
  return AOCR_OK;
}

void aocr_destroy(aocr_handle* h) {
  if (!h) return;
  delete h->eng;
  delete h;
}

const char* aocr_last_error(const aocr_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int aocr_param_groups(const aocr_handle* h, int32_t* n_groups, int64_t sizes[AOCR_NUM_GROUPS]) {
  if (!h || !h->eng) return AOCR_ERR_INVALID;
  if (n_groups) *n_groups = AOCR_NUM_GROUPS;
  if (sizes) for (int g = 0; g < AOCR_NUM_GROUPS; g++) sizes[g] = h->eng->L.gsize[g];
  return AOCR_OK;
}

int aocr_init_params(aocr_handle* h, uint64_t seed) {
  AOCR_API_BEGIN(h) h->eng->init_params(seed); AOCR_API_END(h)
}
int aocr_set_global_batch(aocr_handle* h, int32_t global_batch) {
  AOCR_API_BEGIN(h) h->eng->set_global_batch(global_batch); AOCR_API_END(h)
}
const char* aocr_last_global_error(void) { return g_create_error.c_str(); }

int aocr_set_params(aocr_handle* h, int group, const float* host, int64_t n) {
  AOCR_API_BEGIN(h) h->eng->set_params(group, host, n); AOCR_API_END(h)
}
int aocr_get_params(aocr_handle* h, int group, float* host, int64_t n) {
  AOCR_API_BEGIN(h) h->eng->get_flat(false, group, host, n); AOCR_API_END(h)
}
int aocr_get_grads(aocr_handle* h, int group, float* host, int64_t n) {
  AOCR_API_BEGIN(h) h->eng->get_flat(true, group, host, n); AOCR_API_END(h)
}
int aocr_set_bn_stats(aocr_handle* h, int layer, const float* mean, const float* var, int64_t n) {
  AOCR_API_BEGIN(h) h->eng->set_bn(layer, mean, var, n); AOCR_API_END(h)
}
int aocr_get_bn_stats(aocr_handle* h, int layer, float* mean, float* var, int64_t n) {
  AOCR_API_BEGIN(h) h->eng->get_bn(layer, mean, var, n); AOCR_API_END(h)
}

int aocr_forward_backward(aocr_handle* h, const float* images, int b, int W, const int32_t* targets,
                          const int32_t* targets_eval, int T, double* loss_sum) {
  AOCR_API_BEGIN(h)
  h->eng->stage_batch(images, b, W, targets, targets_eval, T);
  h->eng->forward_backward_enqueue();
  double l = h->eng->read_loss();
  if (loss_sum) *loss_sum = l;
  AOCR_API_END(h)
}

int aocr_group_norms(aocr_handle* h, double pnorm[AOCR_NUM_GROUPS], double gnorm[AOCR_NUM_GROUPS]) {
  AOCR_API_BEGIN(h) h->eng->group_norms(pnorm, gnorm); AOCR_API_END(h)
}

int aocr_sgd_update(aocr_handle* h, double lr, double clip) {
  AOCR_API_BEGIN(h)
  h->eng->sgd_enqueue(lr, clip);
  h->eng->sync();
  AOCR_API_END(h)
}

int aocr_train_step(aocr_handle* h, const float* images, int b, int W, const int32_t* targets,
                    const int32_t* targets_eval, int T, double lr, double* loss_sum) {
  AOCR_API_BEGIN(h)
  h->eng->stage_batch(images, b, W, targets, targets_eval, T);
  h->eng->train_step_enqueue(lr, 5.0);
  double l = h->eng->read_loss();
  if (loss_sum) *loss_sum = l;
  AOCR_API_END(h)
}

int aocr_decode_greedy(aocr_handle* h, const float* images, int b, int W, const int32_t* targets,
                       const int32_t* targets_eval, int T, int32_t* labels, double* pred_scores, double* gold_scores,
                       double* loss_sum, int32_t* num_correct) {
  AOCR_API_BEGIN(h)
  h->eng->stage_batch(images, b, W, targets, targets_eval, T);
  h->eng->decode_step_enqueue();
  h->eng->decode_collect(labels, pred_scores, gold_scores, loss_sum, num_correct);
  AOCR_API_END(h)
}

int aocr_decode_beam(aocr_handle* h, const float* images, int b, int W, const int32_t* targets, const int32_t* targets_eval,
                     int T, int beam_size, const int32_t* trie_table, int32_t trie_nodes, int32_t* labels,
                     double* pred_scores, double* gold_scores, double* loss_sum, int32_t* num_correct) {
  AOCR_API_BEGIN(h)
  h->eng->stage_batch(images, b, W, targets, targets_eval, T);
  h->eng->decode_beam_enqueue(beam_size, trie_table, trie_nodes);
  h->eng->decode_collect(labels, pred_scores, gold_scores, loss_sum, num_correct);
  AOCR_API_END(h)
}

int aocr_get_logprobs(aocr_handle* h, int which, float* out, int64_t n) {
  AOCR_API_BEGIN(h) h->eng->get_logprobs(which, out, n); AOCR_API_END(h)
}
int aocr_debug_read(aocr_handle* h, const char* name, float* out, int64_t n) {
  AOCR_API_BEGIN(h) h->eng->debug_read(name, out, n); AOCR_API_END(h)
}

int aocr_stage_batch(aocr_handle* h, const float* images, int b, int W, const int32_t* targets,
                     const int32_t* targets_eval, int T) {
  AOCR_API_BEGIN(h)
  h->eng->stage_batch(images, b, W, targets, targets_eval, T);
  h->eng->sync();
  AOCR_API_END(h)
}
int aocr_train_step_staged(aocr_handle* h, double lr, int sync, double* loss_sum) {
  AOCR_API_BEGIN(h)
  h->eng->train_step_enqueue(lr, 5.0);
  if (sync) {
    double l = h->eng->read_loss();
    if (loss_sum) *loss_sum = l;
  }
  AOCR_API_END(h)
}
int aocr_decode_greedy_staged(aocr_handle* h, int sync) {
  AOCR_API_BEGIN(h)
  h->eng->decode_step_enqueue();
  if (sync) h->eng->sync();
  AOCR_API_END(h)
}
int aocr_grad_buffer(aocr_handle* h, void** dev_ptr, int64_t* n_floats) {
  AOCR_API_BEGIN(h)
  if (dev_ptr) *dev_ptr = h->eng->d_grads;
  if (n_floats) *n_floats = h->eng->L.total;
  AOCR_API_END(h)
}
int aocr_group_extent(aocr_handle* h, int group, int64_t* offset_floats, int64_t* n_floats) {
  AOCR_API_BEGIN(h)
  AOCR_CHECK(group >= 0 && group < 5, "group must be in [0,5)");
  if (offset_floats) *offset_floats = h->eng->L.goff[group];
  if (n_floats) *n_floats = h->eng->L.gphys[group];
  AOCR_API_END(h)
}
int aocr_forward_backward_staged(aocr_handle* h) {
  AOCR_API_BEGIN(h) h->eng->forward_backward_enqueue(); AOCR_API_END(h)
}
int aocr_sgd_update_async(aocr_handle* h, double lr, double clip) {
  AOCR_API_BEGIN(h) h->eng->sgd_enqueue(lr, clip); AOCR_API_END(h)
}
int aocr_read_loss(aocr_handle* h, double* loss_sum) {
  AOCR_API_BEGIN(h)
  double l = h->eng->read_loss();
  if (loss_sum) *loss_sum = l;
  AOCR_API_END(h)
}
int aocr_set_allreduce(aocr_handle* h, aocr_allreduce_fn fn, void* user) {
  AOCR_API_BEGIN(h)
  h->eng->ar_fn = fn; h->eng->ar_user = user;
  AOCR_API_END(h)
}
int aocr_host_alloc(void** ptr, int64_t bytes) {
  if (!ptr || bytes <= 0) { g_create_error = "aocr_host_alloc: null pointer or non-positive size"; return AOCR_ERR_INVALID; }
  *ptr = nullptr;
  cudaError_t e = cudaHostAlloc(ptr, (size_t)bytes, cudaHostAllocPortable);
  if (e != cudaSuccess) {
    cudaGetLastError();
    *ptr = nullptr;
    g_create_error = std::string("aocr_host_alloc: ") + cudaGetErrorString(e);
    return AOCR_ERR_CUDA;
  }
  return AOCR_OK;
}
void aocr_host_free(void* ptr) {
  if (ptr) cudaFreeHost(ptr);
}
int aocr_dp_unique_id(void* out128) {
  try {
    if (!out128) return AOCR_ERR_INVALID;
    aocr::dp_unique_id(out128);
    return AOCR_OK;
  } catch (const std::exception& e) {
    g_create_error = e.what();        // aocr_last_global_error() / aocr_last_error(NULL)
    return AOCR_ERR_CUDA;
  }
}
int aocr_dp_init(aocr_handle* h, const void* id128) {
  AOCR_API_BEGIN(h)
  AOCR_CHECK(id128 != nullptr, "null unique id");
  h->eng->dp_init(id128);
  AOCR_API_END(h)
}
int aocr_stream(aocr_handle* h, void** cuda_stream) {
  AOCR_API_BEGIN(h)
  if (cuda_stream) *cuda_stream = (void*)h->eng->ctx_.st;
  AOCR_API_END(h)
}
int aocr_synchronize(aocr_handle* h) {
  AOCR_API_BEGIN(h) h->eng->sync(); AOCR_API_END(h)
}
int64_t aocr_launch_count(const aocr_handle* h) { return (h && h->eng) ? h->eng->ctx_.launches : -1; }
int aocr_prof_enable(aocr_handle* h, int on) {
  AOCR_API_BEGIN(h)
  h->eng->prof_collect();
  h->eng->prof_on = on != 0;
  for (int i = 0; i < 4; i++) { h->eng->prof_ms[i] = 0; h->eng->prof_launches[i] = 0; h->eng->prof_work[i] = 0; }
  AOCR_API_END(h)
}
int aocr_prof_read(aocr_handle* h, int cls, double* ms, int64_t* launches, double* work) {
  AOCR_API_BEGIN(h)
  AOCR_CHECK(cls >= 0 && cls < 4, "cls must be in [0,4)");
  h->eng->prof_collect();
  if (ms) *ms = h->eng->prof_ms[cls];
  if (launches) *launches = h->eng->prof_launches[cls];
  if (work) *work = h->eng->prof_work[cls];
  AOCR_API_END(h)
}

}  // extern "C"

extern "C" {
int aocr_grad_scale(aocr_handle* h, int group, double s) {
  AOCR_API_BEGIN(h)
  AOCR_CHECK(group >= 0 && group < 5, "group must be in [0,5)");
  aocr::scale_vec(h->eng->ctx_, h->eng->d_grads + h->eng->L.goff[group], h->eng->L.gphys[group], (float)s);
  AOCR_API_END(h)
}
int aocr_param_axpy(aocr_handle* h, int group, double a) {
  AOCR_API_BEGIN(h)
  AOCR_CHECK(group >= 0 && group < 5, "group must be in [0,5)");
  aocr::axpy_vec(h->eng->ctx_, h->eng->d_params + h->eng->L.goff[group], h->eng->d_grads + h->eng->L.goff[group],
                 h->eng->L.gphys[group], (float)a);
  h->eng->mark_weights_dirty();
  AOCR_API_END(h)
}
}
