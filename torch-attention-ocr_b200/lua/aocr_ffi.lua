--[[ aocr_ffi.lua — LuaJIT FFI binding of include/aocr.h (libaocr.so).
     This is the Lua twin of aocr/capi.py; keep the two mechanically identical.  It cannot be executed in the build
     container (no LuaJIT/Torch7 there): every entry point is exercised through the Python twin instead. ]]
local ffi = require 'ffi'

ffi.cdef[[
typedef struct aocr_handle aocr_handle;
typedef struct aocr_config {
  int32_t batch_size, max_encoder_l, max_decoder_l, encoder_num_hidden, encoder_num_layers, decoder_num_layers,
          target_vocab_size, target_embedding_size, input_feed;
  float dropout, learning_rate;
  int32_t dp_rank, dp_world, global_batch, gemm_mode;
} aocr_config;
int aocr_create(const aocr_config* cfg, int device, aocr_handle** out);
void aocr_destroy(aocr_handle* h);
const char* aocr_last_error(const aocr_handle* h);
int aocr_param_groups(const aocr_handle* h, int32_t* n_groups, int64_t sizes[5]);
int aocr_init_params(aocr_handle* h, uint64_t seed);
int aocr_set_params(aocr_handle* h, int group, const float* host, int64_t n);
int aocr_get_params(aocr_handle* h, int group, float* host, int64_t n);
int aocr_get_grads(aocr_handle* h, int group, float* host, int64_t n);
int aocr_set_bn_stats(aocr_handle* h, int layer, const float* mean, const float* var, int64_t n);
int aocr_get_bn_stats(aocr_handle* h, int layer, float* mean, float* var, int64_t n);
int aocr_forward_backward(aocr_handle* h, const float* images, int b, int W, const int32_t* targets,
                          const int32_t* targets_eval, int T, double* loss_sum);
int aocr_group_norms(aocr_handle* h, double pnorm[5], double gnorm[5]);
int aocr_sgd_update(aocr_handle* h, double lr, double clip);
int aocr_grad_scale(aocr_handle* h, int group, double s);
int aocr_param_axpy(aocr_handle* h, int group, double a);
int aocr_train_step(aocr_handle* h, const float* images, int b, int W, const int32_t* targets,
                    const int32_t* targets_eval, int T, double lr, double* loss_sum);
int aocr_decode_greedy(aocr_handle* h, const float* images, int b, int W, const int32_t* targets,
                       const int32_t* targets_eval, int T, int32_t* labels, double* pred_scores,
                       double* gold_scores, double* loss_sum, int32_t* num_correct);
int aocr_decode_beam(aocr_handle* h, const float* images, int b, int W, const int32_t* targets,
                     const int32_t* targets_eval, int T, int beam_size, const int32_t* trie_table, int32_t trie_nodes,
                     int32_t* labels, double* pred_scores, double* gold_scores, double* loss_sum, int32_t* num_correct);
int aocr_trie_load(const char* path, int allow_digit_prefix, int32_t** table, int32_t* num_nodes);
int aocr_trie_from_words(const char* words, int allow_digit_prefix, int32_t** table, int32_t* num_nodes);
void aocr_trie_free(int32_t* table);
int aocr_host_alloc(void** ptr, int64_t bytes);
void aocr_host_free(void* ptr);
int aocr_get_logprobs(aocr_handle* h, int which, float* out, int64_t n);
int aocr_debug_read(aocr_handle* h, const char* name, float* out, int64_t n);
/* device-resident entry points and the data-parallel plumbing (no reference counterpart: train.lua is single-device) */
int aocr_stage_batch(aocr_handle* h, const float* images, int b, int W, const int32_t* targets,
                     const int32_t* targets_eval, int T);
int aocr_train_step_staged(aocr_handle* h, double lr, int sync, double* loss_sum);
int aocr_decode_greedy_staged(aocr_handle* h, int sync);
int aocr_grad_buffer(aocr_handle* h, void** dev_ptr, int64_t* n_floats);
int aocr_group_extent(aocr_handle* h, int group, int64_t* offset_floats, int64_t* n_floats);
int aocr_forward_backward_staged(aocr_handle* h);
int aocr_sgd_update_async(aocr_handle* h, double lr, double clip);
int aocr_read_loss(aocr_handle* h, double* loss_sum);
int aocr_stream(aocr_handle* h, void** cuda_stream);
typedef void (*aocr_allreduce_fn)(void* user, void* dev_ptr, int64_t n_floats, int kind);
int aocr_set_allreduce(aocr_handle* h, aocr_allreduce_fn fn, void* user);
int aocr_dp_unique_id(void* out128);
int aocr_dp_init(aocr_handle* h, const void* id128);
const char* aocr_last_global_error(void);
int aocr_set_global_batch(aocr_handle* h, int32_t global_batch);
int aocr_synchronize(aocr_handle* h);
int64_t aocr_launch_count(const aocr_handle* h);
int aocr_prof_enable(aocr_handle* h, int on);
int aocr_prof_read(aocr_handle* h, int cls, double* ms, int64_t* launches, double* work);
]]

local lib = ffi.load(os.getenv('AOCR_LIB') or 'torch-attention-ocr_b200/lib/libaocr.so')
local M = { lib = lib, ffi = ffi }

-- loadDictionary(dictionary_path, allow_digit_prefix) of src/utils/utils.lua:177-218, as the flat child table the
-- library consumes: returns {table = int32_t*, nodes = n} (pass it as the `trie` argument of model:step)
function M.loadDictionary(path, allow_digit_prefix)
  local t, n = ffi.new('int32_t*[1]'), ffi.new('int32_t[1]')
  if lib.aocr_trie_load(path, allow_digit_prefix and 1 or 0, t, n) ~= 0 then
    error(string.format('Error: Data file %s not found ', path))
  end
  return { table = ffi.gc(t[0], lib.aocr_trie_free), nodes = n[0] }
end

function M.check(h, rc)
  if rc ~= 0 then error(ffi.string(lib.aocr_last_error(h)), 2) end   -- same texts as the reference's asserts
end
return M
