// engine.h — host-side step scheduler of libaocr: owns device memory, the parameter/gradient flat
// buffers and runs the reference's `feval` schedule (src/model/model.lua:284-695) as kernel launches.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "../../include/aocr.h"
#include "kernels.h"
#include "kernels_dec.h"
#include "persist.h"
#include "gemm_tc.cuh"
#include <tuple>
#include <string.h>

namespace aocr {

// API group ids follow src/model/model.lua:150; the physical order in the flat buffers is the order in
// which backward finishes them ([proj | decoder | enc_fw | enc_bw | cnn]) so a bucketed allreduce can
// start on the head of the buffer while the tail is still being produced.
enum Group { G_CNN = 0, G_ENC_FW = 1, G_ENC_BW = 2, G_DEC = 3, G_PROJ = 4 };

struct ConvSpec { int cin, cout, k, pad, bn /* -1 or bn index */, pool_kw /* 0 none, 1: 2x1, 2: 2x2 */; };
extern const ConvSpec kConv[7];

// one parameter tensor: where it sits in the caller-visible ("Torch") group vector and in the device buffer.
// Device offsets are padded to 64 floats so every tensor is 256-byte aligned (float4 / TMA requirements);
// the padding holds zeros in both params and grads, so group norms and axpys are unaffected.
// init: how Torch7's reset() draws the tensor: 0 = U(+-1/sqrt(fan_in)) weight, 1 = bias of the preceding
// weight (same bound), 2 = batch-norm gamma U(0,1), 3 = batch-norm beta (0), 4 = LookupTable N(0,1)
struct TensorEntry { int group; int64_t ext_off, phys_off, n; int conv_cout, conv_cin, conv_kk; int init = 0; int fan_in = 0; };

struct ParamLayout {
  int64_t goff[5], gsize[5] /* caller-visible sizes */, gphys[5] /* device extent incl. padding */, total;
  std::vector<TensorEntry> tensors;
  int64_t conv_w[7], conv_b[7], bn_g[3], bn_b[3];
  int64_t enc_wi[2], enc_bi[2], enc_wh[2], enc_bh[2];
  int64_t emb, l1_wi, l1_bi, l1_wh, l1_bh, l2_wi, l2_bi, l2_wh, l2_bh, wa, wc;
  int64_t wo, bo;
};

struct Tap { const void* ptr; int64_t n; bool bytes = false; };   // bytes: uint8 on the device, widened to float on read

class Engine {
 public:
  Engine(const aocr_config& cfg, int device);
  ~Engine();

  void set_params(int group, const float* host, int64_t n);
  void init_params(uint64_t seed);   // Model:create's module construction (model.lua:83-112): Torch7 reset() distributions
  void get_flat(bool grads, int group, float* host, int64_t n);
  void set_bn(int layer, const float* mean, const float* var, int64_t n);
  void get_bn(int layer, float* mean, float* var, int64_t n);

  void stage_batch(const float* images, int b, int W, const int32_t* tgt, const int32_t* tev, int T);
  void forward_backward_enqueue();
  double read_loss();
  void group_norms(double* pn, double* gn);
  void sgd_enqueue(double lr, double clip);
  void set_lr_clip(double lr, double clip);
  void sgd_enqueue_kernels();
  void set_global_batch(int n) { AOCR_CHECK(n >= 0, "global batch must be >= 0"); cfg.global_batch = n; }
  int global_b() const { return cfg.global_batch > 0 ? cfg.global_batch : b_; }
  void decode_enqueue();
  void decode_beam_enqueue(int beam, const int32_t* trie_host, int32_t trie_nodes);   // engine_beam.cu
  void decode_collect(int32_t* labels, double* pred, double* gold, double* loss_sum, int32_t* num_correct);
  void get_logprobs(int which, float* out, int64_t n);
  void debug_read(const char* name, float* out, int64_t n);
  void sync() { AOCR_CUDA(cudaStreamSynchronize(ctx_.st)); }
  // whole-step CUDA graphs: the fixed launch sequence of a (b, W, T, lr) train step / (b, W, T) decode is captured
  // once (after one eager warm-up that fills every cache) and replayed; AOCR_GRAPHS=0 disables
  void train_step_enqueue(double lr, double clip);
  void decode_step_enqueue();
  // data-parallel hook (aocr_set_allreduce): kind 0 = small sum ordered on the engine stream (BN statistics),
  // 1 = gradient bucket that may run concurrently with later kernels, 2 = join all outstanding buckets
  aocr_allreduce_fn ar_fn = nullptr;
  void* ar_user = nullptr;
  // native flavour of the same exchange (engine_nccl.cu): NCCL bound at run time, collectives on the engine's own streams
  void dp_init(const void* id128);
  void dp_shutdown();
  bool dp_native() const { return nccl_comm_ != nullptr; }
  StatSync stat_sync();
  void grad_bucket(int first_group, int last_group);
  void grad_range(int64_t off, int64_t end);
  void grad_join();
  void mark_weights_dirty() { weights_dirty_ = true; weights_version_++; }

  aocr_config cfg;
  ParamLayout L;
  Ctx ctx_;
  float* d_params = nullptr;
  float* d_grads = nullptr;
  std::string last_error;
  // profiling of kernel classes (bench roofline): 0 tensor GEMM/conv (flops), 1 attention kernels (bytes),
  // 2 executor (flops of its GEMM commands), 3 executor again (operand bytes its GEMM commands stream)
  bool prof_on = false;
  double prof_ms[4] = {0, 0, 0, 0};
  int64_t prof_launches[4] = {0, 0, 0, 0};
  double prof_work[4] = {0, 0, 0, 0};

 private:
  template <typename T> T* alloc(int64_t n);
  void layout_params();
  void gemm(const Gemm& g, int cls = 0);
  bool prep_weights();
  void cnn_forward(bool train);
  void cnn_backward();
  void encoder_forward();
  void encoder_backward();
  void decoder_init(int reps = 1);
  void attention_precompute();
  void decoder_step(int t, const int32_t* tok);
  void decoder_backward();
  void decoder_step_simt(int t, const int32_t* tok);
  void decoder_step_tc(int t, const int32_t* tok);
  void decoder_backward_steps_simt();
  void decoder_backward_steps_tc();
  void build_decoder_packs();
  bool fused_rec() const { return rec_ && fuse_on_ && rec_->cluster > 1; }
  // the decoder layers use the fused GEMM -> cell commands (and therefore the gate-interleaved weight packs)
  bool dec_fused_ok() const {
    return persist_on_ && fuse_on_ && cluster_ > 1 && b_ <= 256 && cfg.gemm_mode != 2 && Hd % 32 == 0 &&
           pad64(K1) / 64 >= cluster_ && pad64(2 * Hd) / 64 >= cluster_;
  }
  // emitters: launch a piece of a recurrence as its own kernel, or record it into a persistent program
  TcOut emit_gemm(const Pack& W, int M, const Pack& X, int64_t row0, int64_t k0, int K, float* ws);
  template <typename CellT>
  void emit_gemm_fused(int type, const Pack& W, int M, const Pack& X, int64_t row0, int64_t k0, int K, const CellT& cell) {
    const int terms = cfg.gemm_mode == 1 ? 1 : 3;
    PGemmPlan pl = persist_plan_gemm_fused(M, b_, K, rec_->cluster);
    PGemm g{};
    g.map_a = rec_->add_map_pair(tc_map_2d(W.hi, W.rows, W.kp, 128), tc_map_2d(W.lo, W.rows, W.kp, 128));
    g.map_b = rec_->add_map_pair(tc_map_2d(X.hi, X.rows, X.kp, rec_->bn), tc_map_2d(X.lo, X.rows, X.kp, rec_->bn));
    g.m_tiles = pl.m_tiles; g.splits = pl.splits; g.kb_per = pl.kb_per; g.num_kb = pl.num_kb;
    g.b_row0 = (int)row0; g.b_k0 = (int)k0; g.M = M; g.N = b_; g.terms = terms;
    g.ws = nullptr; g.part_stride = 0; g.ldc = M;
    rec_->add2(type, g, cell);
    if (pl.m_tiles * pl.splits > rec_->grid) rec_->grid = pl.m_tiles * pl.splits;
  }
  void emit(const CellFwdTc& p); void emit(const CellBwdTc& p);
  void emit(const EncCellFwdTc& p); void emit(const EncCellBwdTc& p); void emit(const AttnOutTc& p); void emit(const AttnDuTc& p);
  void emit_to_dense(const PartIn& in, float* dst, int64_t ld, int B, int cols);
  void encoder_dir_forward(int d);
  void encoder_dir_backward(int d);
  enum ProgKind { PK_DEC_FWD = 0, PK_DEC_BWD = 1, PK_ENC_FWD0 = 2, PK_ENC_BWD0 = 4, PK_DEC_GREEDY = 6, PK_DEC_DUAL = 7, PK_DEC_DUAL_TAIL = 8 };
  struct StepTail { GenTc gen; GreedyTc sel; };
  const StepTail* tail_ = nullptr;   // set while recording a dual decode step: generator + selection ride on the attention command
  int dual_rows_ = 0;   // > 0 while the dual decode pass is recorded / initialised: the number of real batch rows
  int slot_rows_ = 0;   // > 0: rows per time slot of the decoder state when it differs from the active rows b_ (the
                        // greedy-only tail of the dual decode pass continues on the dual pass's 2B-row layout)
  bool short_gold_on_ = true;   // AOCR_SHORT_GOLD=0: the gold rows ride along for all max_decoder_l steps
  int ctx_row0_ = 0;    // first context row of the state rows (beam search works on chunks of the batch)
  // beam search (engine_beam.cu)
  uint8_t* beam_tmp_ = nullptr; float* beam_logp_ = nullptr; double* beam_scores_[2] = {}; int32_t* beam_tok_[2] = {};
  int32_t *beam_parent_ = nullptr, *beam_loc_[2] = {}, *beam_hist_tok_ = nullptr, *beam_hist_par_ = nullptr;
  int32_t* d_trie_ = nullptr; const int32_t* trie_src_ = nullptr; int32_t trie_nodes_ = 0;
  struct ProgKey { int kind, b, S, nsteps, variant; bool operator<(const ProgKey& o) const {
    return std::tie(kind, b, S, nsteps, variant) < std::tie(o.kind, o.b, o.S, o.nsteps, o.variant); } };
  std::map<ProgKey, PersistProgram> programs_;
  PersistProgram* rec_ = nullptr;
  int rec_max_ctas_ = 128;
  bool dual_on_ = true;         // AOCR_DUAL=0: greedy and gold decode passes one after the other
  bool persist_on_ = true;      // AOCR_PERSIST=0: per-kernel chains instead of the persistent executor
  void run_program(int kind, int nsteps, int variant);
  void encoder_forward_steps_tc();
  void ensure_enc_packs();
  void ensure_wicat_pack();
  cudaEvent_t conv_packs_ev_ = nullptr;    // lane 1 has written the forward convolutions' weight planes (prep_weights)
  bool conv_packs_pending_ = false;
  void encoder_backward_steps_tc();
  void conv_wgrad_tc(const float* dz, const float* x, int N, int H, int W, int Cin, int k, int pad, int Ho, int Wo,
                     int Cout, float* dW, const Pack* zpack = nullptr, const Pack* xpack = nullptr);
  void conv_tc(const float* x, int N, int H, int W, int C, int k, int pad, int Ho, int Wo, const float* Wk, int Cout,
               float* out, const float* bias, const Pack* xpack = nullptr);
  void conv_dims(int l, int& Hin, int& Win, int& Hout, int& Wout) const;
  Pack alloc_pack(int64_t rows, int64_t kp);
  bool is_param(const float* p) const;
  Pack operand_pack(const float* ptr, int64_t rows, int64_t K, int64_t srs, int64_t sks, int slot);
  void prof_begin(int cls);
  void prof_end(int cls, double work);
 public:
  void prof_collect();
 private:
  // AOCR_PHASES=1: per-phase device time of a training step (diagnostic, prints to stderr)
  // lanes: independent work runs on side streams (encoder backward direction on lane 2; time-batched weight
  // gradients on lane 1) while the latency-bound per-timestep chain keeps lane 0 busy.  use_lane() swaps the
  // stream and the stream-private scratch into ctx_/scratch_, so the code that runs on a lane is unchanged.
  struct Lane { cudaStream_t st = nullptr; float* tc_ws = nullptr; Pack scratch[2]; float* partial = nullptr; float* tmpvec = nullptr; };
  Lane lanes_[3];
  int cur_lane_ = 0;
  bool lanes_on_ = true;
  cudaEvent_t lane_ev_[8] = {};
  int lane_ev_next_ = 0;
  void use_lane(int i);
  void fork_to(int i);
  void join_from(int i);
  bool graphs_on_ = true;
  int64_t graph_launches_train_ = 0, graph_launches_decode_ = 0;
  struct GraphKey { int kind, b, W, T, gb; bool operator<(const GraphKey& o) const {
    return std::tie(kind, b, W, T, gb) < std::tie(o.kind, o.b, o.W, o.T, o.gb); } };
  static constexpr size_t kMaxGraphs = 64, kMaxPrograms = 256;   // caches keyed by batch shape: dropped wholesale when full
  void drop_graphs();
  struct GraphEntry { int seen = 0; cudaGraphExec_t exec = nullptr; };
  std::map<GraphKey, GraphEntry> graphs_;
  bool phases_on_ = false;
  std::vector<std::pair<std::string, cudaEvent_t>> phase_marks_;
  void phase_mark(const char* name);
  void phase_report();
  std::vector<cudaEvent_t> prof_pool_;
  std::vector<std::pair<int, double>> prof_recs_;
  size_t prof_used_ = 0;
  bool prof_dump_ = getenv("AOCR_PROF_DUMP") != nullptr;   // print every profiled call (class, time, work)
  std::vector<size_t> prof_open_;

  int device_;
  void* nccl_comm_ = nullptr;        // gradient buckets (communication stream)
  void* nccl_comm_stat_ = nullptr;   // batch-norm statistics (engine stream)
  cudaStream_t comm_st_ = nullptr;
  cudaEvent_t comm_ev_[4] = {};
  int comm_ev_next_ = 0;
  bool comm_pending_ = false;
  // one-shot exchange of the batch-norm statistics through NVLink peer memory (engine_nccl.cu): every rank owns a
  // mailbox that its peers write into; nullptr = not set up (NCCL all-reduce is used instead)
  float* mbox_ = nullptr;            // this rank's mailbox (cudaMalloc, exported with cudaIpcGetMemHandle)
  float** d_peer_mbox_ = nullptr;    // device array [world]: every rank's mailbox as mapped into this process
  std::vector<void*> peer_mapped_;   // cudaIpcOpenMemHandle mappings to close
  unsigned* d_mbox_seq_ = nullptr;   // device counters [kMboxSlots]: use count of a slot (its parity picks the buffer half)
  int mbox_slot_next_ = 0;           // host: slot of the next exchange of the current step (reset at step start)
  void dp_peer_setup();
  bool dp_peer_allreduce(float* buf, int64_t n);   // false: mailboxes unavailable / vector too long (caller falls back to NCCL)
  int64_t cnn_bucket_split_ = -1;    // >= 0: cnn_backward issues [split, end of cnn group) as soon as conv5 is done
  void dp_allreduce(float* buf, int64_t n, int kind);
  void grad_small(int64_t off, int64_t end);
  // optional: per-group clip + SGD as soon as a group's gradients are final (a training step only: aocr_train_step): the
  // update of the decoder / projector / encoder groups (97 MB of the 119 MB) runs on its own stream under the CNN
  // backward, so only the CNN group is left for after the last kernel
  cudaStream_t upd_st_ = nullptr;
  cudaEvent_t upd_ev_[4] = {};
  int upd_ev_next_ = 0;
  unsigned updated_mask_ = 0;        // groups already updated in this step
  bool fused_update_ = false, upd_pending_ = false;
  bool early_update_on_ = false;     // AOCR_EARLY_UPDATE=1 enables.  Measured at config 2 / 4: no gain (3.99 vs 3.98 ms, 9.27 vs 9.28): the
                                     // update's 0.4 GB of HBM traffic slows the CNN backward it overlaps by what the tail saves
  void early_update(int g_first, int g_last, cudaStream_t after);
  void sgd_subset(Ctx& c, unsigned mask);
 public:
  void exchange(float* buf, int64_t n, int kind);
 private:
  struct WeightPack { Pack pack; int64_t version; };
  std::map<std::tuple<const float*, int64_t, int64_t, int64_t, int64_t>, WeightPack> wcache_;
  int64_t weights_version_ = 0;
  Pack scratch_[2];
  // tensor-core decoder path: concatenated weight packs, per-step operand packs, split-K partial regions
  Pack Wcat1p, Wcat2p, W3p, Wcat1Tp, Wcat2Tp, W3Tp;   // W3 = [W_a ; W_c[:, H:]] (rows), W3T = its transpose
  Pack X1p, X2p, H2p, dUQp, dG2p, dG1p;
  float* dzf_[8] = {};   // dz of conv layer l in fp32, per layer: its bias-gradient column sum runs on the side lane too
  Pack dzp_[8];    // dz of conv layer l as bf16 planes (per layer: the weight-gradient GEMM runs on a side lane)
  Pack actp_[8];   // act[l] (input of conv l+1) as bf16 planes, written by the producing kernel; reused by the weight gradient
  Pack Whp[2], WhTp[2], HencP[2], dGeP[2];
  Pack Wcat1pG, Wcat2pG;   // decoder [W_i | W_h] with gate-interleaved rows (fused commands)
  bool dec_packs_inter_ = false;   // which row order the forward packs currently hold
  Pack WiCatP, srcP_;   // [W_i fw ; W_i bw] stacked (8He x 512) and the CNN output as operand planes (one input-projection GEMM)
  int64_t wicat_version_ = -1;
  Pack WhpG[2];     // W_h with gate-interleaved rows (fused GEMM -> cell commands of the executor)
  bool fuse_on_ = true;   // AOCR_FUSE=0: separate GEMM and cell commands
  int cluster_ = 4;       // thread-block cluster size of the executor launches
  float* dec_ws[4] = {nullptr, nullptr, nullptr, nullptr};
  int64_t dec_ws_floats = 0;
  int64_t dec_packs_version_ = -1;
  int64_t enc_packs_version_ = -1;
  int64_t scratch_elems_ = 0;
  int He, Hd, E, V, K1, h1off, Bmax, Smax, Tmax, Wmax;
  std::vector<void*> allocs_;
  std::map<std::string, Tap> taps_;

  // current batch
  int b_ = 0, W_ = 0, T_ = 0, W1_ = 0, W2_ = 0, S_ = 0;
  bool have_batch_ = false, have_grads_ = false, weights_dirty_ = true;
  bool params_set_ = false;   // a step on a handle whose parameters were never drawn / imported is refused (all-zero model)
  int dec_steps_ = 0;
  std::vector<int32_t> h_tev_;
  int last_logp_rows_[3] = {0, 0, 0};
  float* x0 = nullptr;
  int32_t *tgt_bt = nullptr, *tev_bt = nullptr, *tgt_tb = nullptr, *tev_tb = nullptr;
  // CNN
  float *act[8] = {}, *zb[8] = {};
  uint8_t* pidx[8] = {};
  float *bn_mean[3], *bn_var[3], *bn_rmean[3], *bn_rvar[3];
  float *col = nullptr, *gA = nullptr, *gB = nullptr, *partial = nullptr, *tmpvec = nullptr, *wt[8] = {};
  bool cnn_train_ = true;
  // encoder
  float *src = nullptr, *xg = nullptr, *Henc = nullptr, *Cenc = nullptr, *acts_enc = nullptr, *ctx = nullptr;
  float *encb = nullptr, *Dctx = nullptr, *dGe = nullptr, *enc_dh = nullptr, *enc_dc = nullptr, *dsrc = nullptr;
  // decoder
  float *X1 = nullptr, *C1 = nullptr, *ACT1 = nullptr, *X2 = nullptr, *C2 = nullptr, *ACT2 = nullptr, *CAT = nullptr,
        *Q = nullptr, *ALPHA = nullptr, *A_all = nullptr, *Ptab = nullptr, *bsum1 = nullptr, *bsum2 = nullptr,
        *Gs = nullptr;
  float *logp[3] = {}, *dZ = nullptr, *rowloss = nullptr, *dAgen = nullptr, *dU = nullptr, *dCAT = nullptr,
        *DE = nullptr, *dQ = nullptr, *dH2q = nullptr, *dG2 = nullptr, *dG1 = nullptr, *dX2 = nullptr, *dX1 = nullptr,
        *dc1 = nullptr, *dc2 = nullptr, *dP = nullptr, *CtxWc = nullptr, *dCtxWc = nullptr;
  int32_t *tok = nullptr, *labels = nullptr, *tokseq = nullptr;
  double *score = nullptr, *d_loss = nullptr, *d_sumsq = nullptr, *d_sq_partial = nullptr, *d_lrclip = nullptr;
  cudaEvent_t ev0_ = nullptr, ev1_ = nullptr;
};

void dp_unique_id(void* out128);

}  // namespace aocr
