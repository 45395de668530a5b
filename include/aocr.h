/* aocr.h — C ABI of libaocr.so: the B200-native recognition engine behind torch-Attention-OCR's
 * `Model` interface (CNN -> row-wise BiLSTM encoder -> input-fed attention LSTM decoder ->
 * log-softmax generator -> padding-masked NLL, forward + backward, greedy decode, clip+SGD).
 *
 * Every entry point names the reference interface it replaces (file:line under the reference
 * tree).  The reference binds nothing natively today (it is pure Lua over Torch7 modules), so
 * these are the symbols a LuaJIT `ffi.cdef` / Python `ctypes` binding loads.
 *
 * Conventions: plain C, opaque handle, `int` status (0 = OK, <0 = error; text via
 * aocr_last_error).  No exceptions cross the ABI.  One host thread per handle; a handle is
 * bound to one CUDA device.  The caller owns every host buffer; the library owns all device
 * memory.  Host buffers may be pageable or pinned.  Token ids are the reference's 1-based ids
 * (1=PAD 2=GO 3=EOS 4..13='0'..'9' 14..39='a'..'z', src/utils/utils.lua:104-118).
 * Parameter vectors use the external ("Torch") layout documented in DESIGN.md §3:
 * 5 groups in the order of src/model/model.lua:150 {cnn, enc_fw, enc_bw, decoder, proj}.
 */
#ifndef AOCR_H_
#define AOCR_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define AOCR_OK 0
#define AOCR_ERR_INVALID (-1)   /* bad argument / contract violation (reference: Lua assert) */
#define AOCR_ERR_CUDA (-2)      /* CUDA runtime / driver failure */
#define AOCR_ERR_STATE (-3)     /* call order violation (e.g. get_grads before a backward) */
#define AOCR_NUM_GROUPS 5

typedef struct aocr_handle aocr_handle;

/* The fields Model:create reads from `opt` (src/model/model.lua:83-110; defaults src/train.lua:41-63). */
typedef struct aocr_config {
  int32_t batch_size;            /* -batch_size: largest b a step may carry */
  int32_t max_encoder_l;         /* -max_encoder_l (80) */
  int32_t max_decoder_l;         /* -max_decoder_l (50) */
  int32_t encoder_num_hidden;    /* -encoder_num_hidden (512); decoder hidden = 2x (model.lua:88) */
  int32_t encoder_num_layers;    /* must be 1 (reference default) */
  int32_t decoder_num_layers;    /* must be 2 (reference default) */
  int32_t target_vocab_size;     /* 39 */
  int32_t target_embedding_size; /* 20 */
  int32_t input_feed;            /* -input_feed (README enables it) */
  float dropout;                 /* must be 0 (reference default; Dropout(0) is the identity) */
  float learning_rate;           /* initial optim_state.learningRate (model.lua:110) */
  /* data-parallel extension (no reference counterpart): */
  int32_t dp_rank, dp_world;     /* 0,1 for single device */
  int32_t global_batch;          /* 0 = use the step's own b (reference behaviour, model.lua:645-647) */
  int32_t gemm_mode;             /* 0 = tcgen05 bf16x3 split (fp32-grade, default), 1 = tcgen05 bf16x1,
                                    2 = fp32 SIMT (bring-up / on-device cross-check) */
} aocr_config;

/* Model() + model:create(config)  — src/model/model.lua:36,83-112,115-223 */
int aocr_create(const aocr_config* cfg, int device, aocr_handle** out);
/* model:shutdown() + GC — src/model/model.lua:727 */
void aocr_destroy(aocr_handle* h);
const char* aocr_last_error(const aocr_handle* h);   /* h may be NULL: error of a failed aocr_create */

/* Fresh parameters for a created handle: what the module constructors behind model:create draw
 * (src/model/model.lua:83-112 -> cnn.lua:9-45, LSTM.lua:18-162, output_projector.lua:3-8; nn.LinearNoBias:reset,
 * src/utils/model_utils.lua:68-85).  Torch7 reset() distributions: Linear / convolution weight and bias
 * U(+-1/sqrt(fan_in)), batch-norm gamma U(0,1), beta 0, LookupTable N(0,1); running statistics (0,1).
 * A handle refuses to step until aocr_init_params or aocr_set_params has been called (an all-zero model trains nothing). */
int aocr_init_params(aocr_handle* h, uint64_t seed);

/* self.params[i] / self.grad_params[i] — src/model/model.lua:161-168 (getParameters per layer) */
int aocr_param_groups(const aocr_handle* h, int32_t* n_groups, int64_t sizes[AOCR_NUM_GROUPS]);
int aocr_set_params(aocr_handle* h, int group, const float* host, int64_t n);
int aocr_get_params(aocr_handle* h, int group, float* host, int64_t n);
int aocr_get_grads(aocr_handle* h, int group, float* host, int64_t n);
/* BN running statistics (module state saved by model:save, src/model/model.lua:720-725).
 * layer in {0,1,2} = BN after conv3/conv5/conv7; mean/var have 256/512/512 entries. */
int aocr_set_bn_stats(aocr_handle* h, int layer, const float* mean, const float* var, int64_t n);
int aocr_get_bn_stats(aocr_handle* h, int layer, float* mean, float* var, int64_t n);

/* feval, train branch — src/model/model.lua:284-316,537-569,634-695.  images (b,1,32,W) raw gray
 * 0..255; targets/targets_eval (b,T) int32 (src/data/data_gen.lua:106-117).  loss_sum = loss*b
 * (model.lua:701).  Gradients stay on the device (aocr_get_grads / aocr_sgd_update). */
int aocr_forward_backward(aocr_handle* h, const float* images, int b, int W,
                          const int32_t* targets, const int32_t* targets_eval, int T, double* loss_sum);
/* y:norm(), dfdy:norm() per group — src/optim/optim_sgd.lua:49-50 */
int aocr_group_norms(aocr_handle* h, double pnorm[AOCR_NUM_GROUPS], double gnorm[AOCR_NUM_GROUPS]);
/* clip (>clip -> scale to clip) + p -= lr*g per group — src/optim/optim_sgd.lua:50-52,90 */
int aocr_sgd_update(aocr_handle* h, double lr, double clip);
/* the two tensor methods an UNMODIFIED optim.sgd_list applies to the group vectors:
 * dfdy:mul(s) (optim_sgd.lua:51) and y:add(-clr, dfdy) (optim_sgd.lua:90) */
int aocr_grad_scale(aocr_handle* h, int group, double s);
int aocr_param_axpy(aocr_handle* h, int group, double a);   /* params[group] += a * grads[group] */
/* model:step(batch, false): feval + optim.sgd_list — src/model/model.lua:698-701 */
int aocr_train_step(aocr_handle* h, const float* images, int b, int W,
                    const int32_t* targets, const int32_t* targets_eval, int T,
                    double lr, double* loss_sum);
/* model:step(batch, true, 1, nil): greedy decode + gold pass — src/model/model.lua:360-404,446-459,
 * 516-536,570-627,703-704.  labels (b,max_decoder_l); pred/gold scores (b).  Any output pointer
 * may be NULL.  Returns loss_sum and the exact-match count (stats[2]). */
int aocr_decode_greedy(aocr_handle* h, const float* images, int b, int W,
                       const int32_t* targets, const int32_t* targets_eval, int T,
                       int32_t* labels, double* pred_scores, double* gold_scores,
                       double* loss_sum, int32_t* num_correct);
/* model:step(batch, true, beam_size, trie) with beam_size > 1 and / or a dictionary — src/model/model.lua:226-251,
 * 321-536 (beam search: first step on the un-replicated rows, sticky PAD, top-k over beam x vocabulary totals or the
 * sorted walk over trie-valid continuations, parent re-gather of the state), :573-585 (backtrack from the best final
 * beam), :589-627 (gold pass).  trie_table: (trie_nodes, target_vocab_size + 1) child table from aocr_trie_load /
 * aocr_trie_from_words (column = 1-based vocabulary id, -1 = no child, node 0 = the start symbol), or NULL for an
 * unconstrained search.  beam_size is clipped to the vocabulary (model.lua:229); beam_size 1 without a trie equals
 * aocr_decode_greedy.  Outputs as aocr_decode_greedy (pred_scores = score of the best beam). */
int aocr_decode_beam(aocr_handle* h, const float* images, int b, int W,
                     const int32_t* targets, const int32_t* targets_eval, int T,
                     int beam_size, const int32_t* trie_table, int32_t trie_nodes,
                     int32_t* labels, double* pred_scores, double* gold_scores,
                     double* loss_sum, int32_t* num_correct);
/* loadDictionary(dictionary_path, allow_digit_prefix) — src/utils/utils.lua:177-218: one word per line over [0-9a-z];
 * every word ends in an EOS child; allow_digit_prefix lets the root loop on EOS and digits.  The table is malloc'ed by
 * the library and released with aocr_trie_free.  aocr_trie_from_words takes the words '\n'-separated in memory. */
int aocr_trie_load(const char* path, int allow_digit_prefix, int32_t** table, int32_t* num_nodes);
int aocr_trie_from_words(const char* words, int allow_digit_prefix, int32_t** table, int32_t* num_nodes);
void aocr_trie_free(int32_t* table);

/* parity tap: log-probs of the last call. which=0: train (T,b,V); 1: greedy pass (L,b,V) after the
 * sticky-PAD edit; 2: gold pass (L,b,V) - computed for the batch's own target length T; rows t >= T may be zero (their
 * targets are padding: no loss, no score, src/model/criterion.lua:5, src/model/model.lua:614-618). */
int aocr_get_logprobs(aocr_handle* h, int which, float* out, int64_t n);
/* parity tap for intermediate tensors ("cnn_out" (S,b,512), "context" (b,S,1024), ...) */
int aocr_debug_read(aocr_handle* h, const char* name, float* out, int64_t n);

/* ---- page-locked host buffers for the batch tuple ---- */
/* The data layer builds every batch into fresh host tensors (`torch.Tensor(batch_size, 1, imgH, imgW)`,
 * src/data/data_gen.lua:97,105-109).  Allocated here instead, they are page-locked: the copies aocr_train_step /
 * aocr_decode_* / aocr_stage_batch issue from them are asynchronous DMA transfers, so a prefetching data layer overlaps
 * input transfer with the previous step.  Needs a CUDA device (AOCR_ERR_CUDA otherwise, message through
 * aocr_last_global_error); any host pointer remains valid input to every entry point. */
int aocr_host_alloc(void** ptr, int64_t bytes);
void aocr_host_free(void* ptr);

/* ---- device-resident entry points (bench `value`, data-parallel plumbing) ---- */
/* stage a batch in HBM once; *_staged calls then run with no host<->device copies of inputs */
int aocr_stage_batch(aocr_handle* h, const float* images, int b, int W,
                     const int32_t* targets, const int32_t* targets_eval, int T);
int aocr_train_step_staged(aocr_handle* h, double lr, int sync, double* loss_sum /* may be NULL if !sync */);
int aocr_decode_greedy_staged(aocr_handle* h, int sync);
/* flat gradient buffer (device pointer, fp32, physical order [proj|decoder|enc_fw|enc_bw|cnn]) for the
 * NCCL allreduce done by the host runtime; and the backward split used to overlap it. */
int aocr_grad_buffer(aocr_handle* h, void** dev_ptr, int64_t* n_floats);
int aocr_group_extent(aocr_handle* h, int group, int64_t* offset_floats, int64_t* n_floats);
int aocr_forward_backward_staged(aocr_handle* h);          /* enqueue only (no sync, no update) */
int aocr_sgd_update_async(aocr_handle* h, double lr, double clip);
int aocr_read_loss(aocr_handle* h, double* loss_sum);      /* syncs */
int aocr_stream(aocr_handle* h, void** cuda_stream);       /* cudaStream_t the engine enqueues on */
/* Data-parallel exchange hook (dp_world > 1).  The engine calls fn at its exchange points with a device buffer
 * that must be SUM-all-reduced over the ranks:
 *   kind 0: batch-norm statistics (a few KB) - must be ordered on the engine stream (the next kernel reads it);
 *   kind 1: a finished bucket of the flat gradient buffer - may run on another stream concurrently with the rest
 *           of backward; the engine does not touch it again before the join;
 *   kind 2: join - dev_ptr NULL; make the engine stream wait for every outstanding kind-1 reduction.
 * The host runtime implements it with NCCL (aocr/dist.py: torch.distributed). */
typedef void (*aocr_allreduce_fn)(void* user, void* dev_ptr, int64_t n_floats, int kind);
int aocr_set_allreduce(aocr_handle* h, aocr_allreduce_fn fn, void* user);
/* Native exchange: the library issues the same three kinds as NCCL all-reduces itself (libnccl.so.2 bound at run
 * time), on its own streams, so a data-parallel step is captured into the whole-step CUDA graph.  Rank 0 calls
 * aocr_dp_unique_id (128 bytes = ncclUniqueId), the host ships it to every rank, every rank of the handle's
 * dp_world calls aocr_dp_init (collective).  Takes precedence over a hook.  No reference counterpart. */
int aocr_dp_unique_id(void* out128);
/* text of the last failure of a handle-less call (aocr_create, aocr_dp_unique_id) on this thread */
const char* aocr_last_global_error(void);
/* data parallelism: the GLOBAL batch of the next steps (loss scale 1/B and batch-norm row count).  Needed
 * when a step carries fewer images than aocr_config.global_batch (the reference's final bucket flush,
 * src/data/data_gen.lua:125-153); shards of a step need not be equal. */
int aocr_set_global_batch(aocr_handle* h, int32_t global_batch);
int aocr_dp_init(aocr_handle* h, const void* id128);
int aocr_synchronize(aocr_handle* h);
/* number of kernel launches the library has issued on this handle (bench `gpu_launches`) */
int64_t aocr_launch_count(const aocr_handle* h);
/* timing of the dominant kernel class, for bench.py's roofline: accumulates CUDA-event time (ms) of
 * kernel class `cls` (0=tensor GEMM/conv [flops], 1=attention kernels [bytes], 2=recurrence executor [flops of its
 * GEMM commands], 3=recurrence executor [operand bytes its GEMM commands stream]) between reset and read. */
int aocr_prof_enable(aocr_handle* h, int on);
int aocr_prof_read(aocr_handle* h, int cls, double* ms, int64_t* launches, double* work /* flops or bytes */);

#ifdef __cplusplus
}
#endif
#endif /* AOCR_H_ */
