// engine.cu — step scheduler: the reference's feval (src/model/model.lua:284-695) as a fixed launch
// sequence on one CUDA stream.  No host synchronisation inside a step; the loss scalar, greedy labels
// and scores are read back once at the end (the reference syncs every decoder step, model.lua:403,612).
#include "engine.h"

#include <math.h>
#include <string.h>

#include <random>

namespace aocr {

const ConvSpec kConv[7] = {
    // cin cout k pad bn pool_kw          (src/model/cnn.lua:12-42)
    {1, 64, 3, 1, -1, 2},   {64, 128, 3, 1, -1, 2}, {128, 256, 3, 1, 0, 0}, {256, 256, 3, 1, -1, 1},
    {256, 512, 3, 1, 1, 0}, {512, 512, 3, 1, -1, 1}, {512, 512, 2, 0, 2, 0},
};

namespace {
// Wt[ci][k*k-1-tap][co] = W[co][tap][ci]  (flipped taps, in/out swapped): the data-gradient weight.
__global__ void conv_wt_kernel(const float* __restrict__ W, float* __restrict__ Wt, int cout, int kk, int cin) {
  const int64_t total = (int64_t)cout * kk * cin;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int co = (int)(e % cout);
    int64_t r = e / cout;
    int tp = (int)(r % kk);
    int ci = (int)(r / kk);
    Wt[e] = W[((int64_t)co * kk + (kk - 1 - tp)) * cin + ci];
  }
}
}  // namespace

template <typename T>
T* Engine::alloc(int64_t n) {
  void* p = nullptr;
  size_t bytes = (size_t)(n > 0 ? n : 1) * sizeof(T);
  bytes = (bytes + 255) & ~(size_t)255;
  AOCR_CUDA(cudaMalloc(&p, bytes));
  AOCR_CUDA(cudaMemsetAsync(p, 0, bytes, ctx_.st));
  allocs_.push_back(p);
  return reinterpret_cast<T*>(p);
}

Pack Engine::alloc_pack(int64_t rows, int64_t kp) {
  Pack p;
  p.rows = rows; p.kp = kp;
  p.hi = alloc<__nv_bfloat16>(rows * kp);
  p.lo = alloc<__nv_bfloat16>(rows * kp);
  return p;
}

void Engine::layout_params() {
  int64_t off = 0;
  int cur_group = 0;
  int64_t ext = 0;
  // init / fan_in: see TensorEntry (a bias takes the bound of the weight in front of it)
  int last_fan = 1;
  auto take = [&](int64_t n, int cout = 0, int cin = 0, int kk = 0, int init = 1, int fan_in = 0) {
    off = (off + 63) & ~(int64_t)63;
    TensorEntry e{cur_group, ext, off, n, cout, cin, kk};
    if (cout) { init = 0; fan_in = cin * kk; }
    if (init == 0) last_fan = fan_in;
    e.init = init; e.fan_in = init <= 1 ? last_fan : 0;
    L.tensors.push_back(e);
    int64_t o = off;
    off += n; ext += n;
    return o;
  };
  auto begin_group = [&](int g) { off = (off + 63) & ~(int64_t)63; cur_group = g; ext = 0; L.goff[g] = off; };
  auto end_group = [&](int g) { off = (off + 63) & ~(int64_t)63; L.gsize[g] = ext; L.gphys[g] = off - L.goff[g]; };
  const int in1 = E + (cfg.input_feed ? Hd : 0);
  // physical order: proj | decoder | enc_fw | enc_bw | cnn ; within a group the caller-visible tensor order
  begin_group(G_PROJ);
  L.wo = take((int64_t)V * Hd, 0, 0, 0, 0, Hd); L.bo = take(V);
  end_group(G_PROJ);
  begin_group(G_DEC);
  L.emb = take((int64_t)V * E, 0, 0, 0, 4);
  L.l1_wi = take((int64_t)4 * Hd * in1, 0, 0, 0, 0, in1); L.l1_bi = take(4 * Hd);
  L.l1_wh = take((int64_t)4 * Hd * Hd, 0, 0, 0, 0, Hd);   L.l1_bh = take(4 * Hd);
  L.l2_wi = take((int64_t)4 * Hd * Hd, 0, 0, 0, 0, Hd);   L.l2_bi = take(4 * Hd);
  L.l2_wh = take((int64_t)4 * Hd * Hd, 0, 0, 0, 0, Hd);   L.l2_bh = take(4 * Hd);
  L.wa = take((int64_t)Hd * Hd, 0, 0, 0, 0, Hd); L.wc = take((int64_t)Hd * 2 * Hd, 0, 0, 0, 0, 2 * Hd);
  end_group(G_DEC);
  for (int d = 0; d < 2; d++) {
    int g = d == 0 ? G_ENC_FW : G_ENC_BW;
    begin_group(g);
    L.enc_wi[d] = take((int64_t)4 * He * 512, 0, 0, 0, 0, 512); L.enc_bi[d] = take(4 * He);
    L.enc_wh[d] = take((int64_t)4 * He * He, 0, 0, 0, 0, He);   L.enc_bh[d] = take(4 * He);
    end_group(g);
  }
  begin_group(G_CNN);
  for (int l = 0; l < 7; l++) {
    const ConvSpec& c = kConv[l];
    L.conv_w[l] = take((int64_t)c.cout * c.cin * c.k * c.k, c.cout, c.cin, c.k * c.k);
    L.conv_b[l] = take(c.cout);
    if (c.bn >= 0) { L.bn_g[c.bn] = take(c.cout, 0, 0, 0, 2); L.bn_b[c.bn] = take(c.cout, 0, 0, 0, 3); }
  }
  end_group(G_CNN);
  L.total = off;
}

Engine::Engine(const aocr_config& c, int device) : cfg(c), device_(device) {
  AOCR_CHECK(c.batch_size >= 1 && c.max_encoder_l >= 1 && c.max_decoder_l >= 1, "batch_size/max_*_l must be >= 1");
  AOCR_CHECK(c.encoder_num_layers == 1 && c.decoder_num_layers == 2,
             "only encoder_num_layers=1, decoder_num_layers=2 (the reference defaults) are supported");
  AOCR_CHECK(c.encoder_num_hidden >= 64 && c.encoder_num_hidden % 64 == 0 && c.encoder_num_hidden <= 512,
             "encoder_num_hidden must be a multiple of 64 in [64,512]");
  AOCR_CHECK(c.target_vocab_size >= 4 && c.target_vocab_size <= 64, "target_vocab_size must be in [4,64]");
  AOCR_CHECK(c.target_embedding_size >= 1 && c.target_embedding_size % 4 == 0, "target_embedding_size must be a multiple of 4");
  AOCR_CHECK(c.dropout == 0.0f, "dropout must be 0 (reference default)");
  AOCR_CHECK(c.dp_world >= 0 && c.dp_rank >= 0, "bad dp_rank/dp_world");
  AOCR_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  AOCR_CUDA(cudaGetDeviceProperties(&prop, device));
  AOCR_CHECK(prop.major == 10, "libaocr is built for sm_100a (B200) only");
  ctx_.num_sms = prop.multiProcessorCount;
  if (const char* e = getenv("AOCR_PDL")) ctx_.pdl = atoi(e) != 0;
  if (const char* e = getenv("AOCR_PHASES")) phases_on_ = atoi(e) != 0;
  if (const char* e = getenv("AOCR_GRAPHS")) graphs_on_ = atoi(e) != 0;
  if (const char* e = getenv("AOCR_PERSIST")) persist_on_ = atoi(e) != 0;
  if (const char* e = getenv("AOCR_DUAL")) dual_on_ = atoi(e) != 0;
  if (const char* e = getenv("AOCR_FUSE")) fuse_on_ = atoi(e) != 0;
  if (const char* e = getenv("AOCR_CLUSTER")) cluster_ = atoi(e) > 0 ? atoi(e) : 1;
  // Clusters must fit inside a GPC: if this part cannot keep the 128 CTAs the programs need (the decoder, or the two
  // encoder directions side by side) co-resident as clusters, run the executor without clusters and fused commands.
  // Decided here, before any weight plane is built (their row order depends on it).
  if (cluster_ > 1 && c.gemm_mode != 2) {
    const int cap = persist_max_cluster_ctas(64, cluster_);
    if (cap < 128) {
      fprintf(stderr, "[aocr] %d CTAs can be co-resident in clusters of %d (< 128): executor runs without clusters\n", cap, cluster_);
      cluster_ = 1;
    }
  }
  if (c.batch_size > 256 || c.gemm_mode == 2) persist_on_ = false;   // one UMMA N tile (<= 256 rows) per command
  AOCR_CUDA(cudaStreamCreateWithFlags(&ctx_.st, cudaStreamNonBlocking));
  // Launch mode of the executor.  With thread-block clusters the launch does NOT carry the cooperative attribute:
  // Nsight Compute rejects cooperative launches whose kernels use cluster barriers / distributed shared memory
  // (LaunchFailed; measured: cooperative + cluster attribute alone is accepted, the fused commands are not), and the
  // attribute only makes the driver check what is checked above - that the grid (<= 128 CTAs, one per SM) is
  // co-resident as clusters.  Kernels of the other lanes that occupy SMs at launch time always finish, so the spinning
  // CTAs of a partially resident grid cannot deadlock.  Without clusters the launch stays cooperative.
  // AOCR_COOP=1 forces the attribute.  Decided here, before any weight plane is built (their row order depends on it).
  ctx_.persist_coop = cluster_ <= 1;
  if (const char* e = getenv("AOCR_COOP")) ctx_.persist_coop = atoi(e) != 0;
  if (persist_on_ && cluster_ > 1 && !getenv("AOCR_NO_PROBE")) {
    if (persist_probe(ctx_.st, 128, cluster_, ctx_.persist_coop) != cudaSuccess) {
      cluster_ = 1;
      ctx_.persist_coop = true;
      fprintf(stderr, "[aocr] cluster launches refused: executor runs without clusters\n");
      AOCR_CHECK(persist_probe(ctx_.st, 128, 1, true) == cudaSuccess, "the persistent executor cannot be launched on this device");
    }
  }
  AOCR_CUDA(cudaEventCreate(&ev0_));
  AOCR_CUDA(cudaEventCreate(&ev1_));
  He = c.encoder_num_hidden; Hd = 2 * He; E = c.target_embedding_size; V = c.target_vocab_size;
  K1 = cfg.input_feed ? 2 * Hd : Hd; h1off = cfg.input_feed ? Hd : 0;
  Bmax = c.batch_size; Smax = c.max_encoder_l; Tmax = c.max_decoder_l;
  Wmax = 4 * (Smax + 1) + 3;
  layout_params();
  d_params = alloc<float>(L.total);
  d_grads = alloc<float>(L.total);

  const int64_t B = Bmax, W1 = Wmax / 2, W2 = W1 / 2, S = Smax, T = Tmax;
  x0 = alloc<float>(B * 32 * Wmax);
  tgt_bt = alloc<int32_t>(B * T); tev_bt = alloc<int32_t>(B * T);
  tgt_tb = alloc<int32_t>(B * T); tev_tb = alloc<int32_t>(B * T);
  act[1] = alloc<float>(B * 16 * W1 * 64);   pidx[1] = alloc<uint8_t>(B * 16 * W1 * 64);
  zb[2] = alloc<float>(B * 16 * W1 * 128);   act[2] = alloc<float>(B * 8 * W2 * 128); pidx[2] = alloc<uint8_t>(B * 8 * W2 * 128);
  zb[3] = alloc<float>(B * 8 * W2 * 256);    act[3] = alloc<float>(B * 8 * W2 * 256);
  zb[4] = alloc<float>(B * 8 * W2 * 256);    act[4] = alloc<float>(B * 4 * W2 * 256); pidx[4] = alloc<uint8_t>(B * 4 * W2 * 256);
  zb[5] = alloc<float>(B * 4 * W2 * 512);    act[5] = alloc<float>(B * 4 * W2 * 512);
  zb[6] = alloc<float>(B * 4 * W2 * 512);    act[6] = alloc<float>(B * 2 * W2 * 512); pidx[6] = alloc<uint8_t>(B * 2 * W2 * 512);
  zb[7] = alloc<float>(B * S * 512);
  src = alloc<float>(S * B * 512);
  act[7] = src;
  if (cfg.gemm_mode != 2) {
    const int64_t n_act[7] = {0, B * 16 * W1 * 64, B * 8 * W2 * 128, B * 8 * W2 * 256, B * 4 * W2 * 256, B * 4 * W2 * 512, B * 2 * W2 * 512};
    for (int l = 1; l <= 6; l++) actp_[l] = alloc_pack(1, n_act[l]);
    const int64_t n_z[7] = {0, B * 16 * W1 * 128, B * 8 * W2 * 256, B * 8 * W2 * 256, B * 4 * W2 * 512, B * 4 * W2 * 512, B * S * 512};
    for (int l = 1; l <= 6; l++) { dzp_[l] = alloc_pack(1, n_z[l]); dzf_[l] = alloc<float>(n_z[l]); }
    srcP_ = alloc_pack(S * B, 512);
    WiCatP = alloc_pack(8 * c.encoder_num_hidden, 512);   // dz of conv_{l+1} has the shape of zb[l+1]
  }
  const int bnc[3] = {256, 512, 512};
  for (int i = 0; i < 3; i++) {
    bn_mean[i] = alloc<float>(bnc[i]); bn_var[i] = alloc<float>(bnc[i]);
    bn_rmean[i] = alloc<float>(bnc[i]); bn_rvar[i] = alloc<float>(bnc[i]);
    std::vector<float> ones(bnc[i], 1.0f);
    AOCR_CUDA(cudaMemcpyAsync(bn_rvar[i], ones.data(), bnc[i] * sizeof(float), cudaMemcpyHostToDevice, ctx_.st));
    AOCR_CUDA(cudaStreamSynchronize(ctx_.st));
  }
  col = alloc<float>(B * W1 * 18432);
  gA = alloc<float>(B * 16 * W1 * 128);
  gB = alloc<float>(B * 16 * W1 * 128);
  partial = alloc<float>(kPartialFloats);
  AOCR_CUDA(cudaMemset(partial, 0, (size_t)kPartialFloats * sizeof(float)));      // arrival counters in its tail start at 0
  tmpvec = alloc<float>(8192);
  if (cfg.gemm_mode != 2) {
    ctx_.tc_ws_floats = (int64_t)32 * 148 * 128 * 128;
    ctx_.tc_ws = alloc<float>(ctx_.tc_ws_floats);
    ctx_.tc_counters_n = 4096;
    ctx_.tc_counters = alloc<int>(ctx_.tc_counters_n);
    scratch_elems_ = B * W1 * 18432 + (int64_t)4608 * 64 * 8;
    for (int i = 0; i < 2; i++) scratch_[i] = alloc_pack(1, scratch_elems_);
    // tensor-core decoder path (engine_dec_tc.cu)
    const int64_t Hd_ = 2 * c.encoder_num_hidden, K1_ = c.input_feed ? 2 * Hd_ : Hd_, Tm = c.max_decoder_l;
    Wcat1p = alloc_pack(4 * Hd_, K1_); Wcat2p = alloc_pack(4 * Hd_, 2 * Hd_);
    Wcat1pG = alloc_pack(4 * Hd_, K1_); Wcat2pG = alloc_pack(4 * Hd_, 2 * Hd_);
    W3p = alloc_pack(2 * Hd_, Hd_);
    Wcat1Tp = alloc_pack(K1_, 4 * Hd_); Wcat2Tp = alloc_pack(2 * Hd_, 4 * Hd_);
    W3Tp = alloc_pack(Hd_, 2 * Hd_);
    // decoder forward state holds 2B rows: greedy decode runs its greedy and its gold pass as ONE batch of 2B
    X1p = alloc_pack(Tm * 2 * B, K1_); X2p = alloc_pack(Tm * 2 * B, 2 * Hd_); H2p = alloc_pack(Tm * 2 * B, Hd_);
    // backward operands are kept for ALL timesteps: the time-batched weight gradients read them as they are
    dUQp = alloc_pack(Tm * B, 2 * Hd_); dG2p = alloc_pack(Tm * B, 4 * Hd_); dG1p = alloc_pack(Tm * B, 4 * Hd_);
    dec_ws_floats = (int64_t)16 * 2 * B * 4 * Hd_ + 1024;
    for (int i = 0; i < 4; i++) dec_ws[i] = alloc<float>(dec_ws_floats);
    // tensor-core encoder recurrence (engine_enc_tc.cu)
    const int64_t He_ = c.encoder_num_hidden;
    for (int d = 0; d < 2; d++) {
      Whp[d] = alloc_pack(4 * He_, He_); WhTp[d] = alloc_pack(He_, 4 * He_); WhpG[d] = alloc_pack(4 * He_, He_);
      HencP[d] = alloc_pack((S + 1) * B, He_); dGeP[d] = alloc_pack(S * B, 4 * He_);   // all timesteps (weight gradients)
    }
  }
  for (int l = 1; l < 7; l++) wt[l] = alloc<float>((int64_t)kConv[l].cout * kConv[l].cin * kConv[l].k * kConv[l].k);
  // side lanes (lane 0 = the members above)
  if (const char* e = getenv("AOCR_LANES")) lanes_on_ = atoi(e) != 0;
  for (int i = 0; i < 8; i++) AOCR_CUDA(cudaEventCreateWithFlags(&lane_ev_[i], cudaEventDisableTiming));
  if (const char* e = getenv("AOCR_EARLY_UPDATE")) early_update_on_ = atoi(e) != 0;
  if (const char* e = getenv("AOCR_SHORT_GOLD")) short_gold_on_ = atoi(e) != 0;
  AOCR_CUDA(cudaEventCreateWithFlags(&conv_packs_ev_, cudaEventDisableTiming));
  if (lanes_on_) {
    AOCR_CUDA(cudaStreamCreateWithFlags(&upd_st_, cudaStreamNonBlocking));
    for (int i = 0; i < 4; i++) AOCR_CUDA(cudaEventCreateWithFlags(&upd_ev_[i], cudaEventDisableTiming));
  }
  lanes_[0].st = ctx_.st; lanes_[0].tc_ws = ctx_.tc_ws; lanes_[0].scratch[0] = scratch_[0]; lanes_[0].scratch[1] = scratch_[1];
  lanes_[0].partial = partial; lanes_[0].tmpvec = tmpvec;
  for (int i = 1; i < 3; i++) {
    if (!lanes_on_) { lanes_[i] = lanes_[0]; continue; }
    AOCR_CUDA(cudaStreamCreateWithFlags(&lanes_[i].st, cudaStreamNonBlocking));
    lanes_[i].partial = alloc<float>(kPartialFloats);
    AOCR_CUDA(cudaMemset(lanes_[i].partial, 0, (size_t)kPartialFloats * sizeof(float)));
    lanes_[i].tmpvec = alloc<float>(8192);
    if (cfg.gemm_mode != 2) {
      lanes_[i].tc_ws = alloc<float>(ctx_.tc_ws_floats);
      if (i == 1) for (int k = 0; k < 2; k++) lanes_[i].scratch[k] = alloc_pack(1, scratch_elems_);
    }
  }

  xg = alloc<float>(S * B * 8 * He);
  Henc = alloc<float>(2 * (S + 1) * B * He); Cenc = alloc<float>(2 * (S + 1) * B * He);
  acts_enc = alloc<float>(2 * S * B * 4 * He);
  ctx = alloc<float>(B * S * 2 * He);
  encb = alloc<float>(8 * He);
  Dctx = alloc<float>(B * S * 2 * He); dGe = alloc<float>(S * B * 8 * He);
  enc_dh = alloc<float>(2 * B * He); enc_dc = alloc<float>(2 * B * He);
  dsrc = alloc<float>(S * B * 512);

  const int64_t B2 = 2 * B;   // see the dual decode pass (engine_dec.cu: decode_enqueue)
  X1 = alloc<float>(T * B2 * K1);  C1 = alloc<float>((T + 1) * B2 * Hd); ACT1 = alloc<float>(T * B2 * 4 * Hd);
  X2 = alloc<float>(T * B2 * 2 * Hd); C2 = alloc<float>((T + 1) * B2 * Hd); ACT2 = alloc<float>(T * B2 * 4 * Hd);
  CAT = alloc<float>(T * B2 * 2 * Hd); Q = alloc<float>(T * B2 * Hd); ALPHA = alloc<float>(T * B2 * S);
  A_all = alloc<float>(T * B2 * Hd); tokseq = alloc<int32_t>((T + 1) * B2);
  Ptab = alloc<float>((int64_t)V * 4 * Hd); bsum1 = alloc<float>(4 * Hd); bsum2 = alloc<float>(4 * Hd);
  Gs = alloc<float>(B * 4 * Hd);
  for (int i = 0; i < 3; i++) logp[i] = alloc<float>(T * B * V);
  dZ = alloc<float>(T * B * V); rowloss = alloc<float>(T * B); dAgen = alloc<float>(T * B * Hd);
  dU = alloc<float>(T * B * Hd); dCAT = alloc<float>(T * B * 2 * Hd); DE = alloc<float>(T * B * S);
  dQ = alloc<float>(T * B * Hd); dH2q = alloc<float>(B * Hd);
  CtxWc = alloc<float>(B * S * Hd); dCtxWc = alloc<float>(B * S * Hd);
  dG2 = alloc<float>(T * B * 4 * Hd); dG1 = alloc<float>(T * B * 4 * Hd);
  dX2 = alloc<float>(B * 2 * Hd); dX1 = alloc<float>(B * K1);
  dc1 = alloc<float>(B * Hd); dc2 = alloc<float>(B * Hd); dP = alloc<float>((int64_t)V * 4 * Hd);
  tok = alloc<int32_t>(B); labels = alloc<int32_t>(B * T);
  score = alloc<double>(B); d_loss = alloc<double>(1); d_sumsq = alloc<double>(16); d_sq_partial = alloc<double>(5 * 1024);
  d_lrclip = alloc<double>(2);
  AOCR_CUDA(cudaStreamSynchronize(ctx_.st));
}

Engine::~Engine() {
  cudaSetDevice(device_);
  use_lane(0);
  if (ctx_.st) cudaStreamSynchronize(ctx_.st);
  // captured graphs hold references on the NCCL communicator (persistent plans): they must go first, or
  // ncclCommDestroy waits for them forever
  for (auto& kv : graphs_) if (kv.second.exec) { cudaGraphExecDestroy(kv.second.exec); kv.second.exec = nullptr; }
  dp_shutdown();
  for (int i = 1; i < 3; i++)
    if (lanes_on_ && lanes_[i].st) { cudaStreamSynchronize(lanes_[i].st); cudaStreamDestroy(lanes_[i].st); }
  for (int i = 0; i < 8; i++) if (lane_ev_[i]) cudaEventDestroy(lane_ev_[i]);
  if (conv_packs_ev_) cudaEventDestroy(conv_packs_ev_);
  if (upd_st_) { cudaStreamSynchronize(upd_st_); cudaStreamDestroy(upd_st_); }
  for (int i = 0; i < 4; i++) if (upd_ev_[i]) cudaEventDestroy(upd_ev_[i]);
  for (void* p : allocs_) cudaFree(p);
  if (d_trie_) cudaFree(d_trie_);
  if (ev0_) cudaEventDestroy(ev0_);
  if (ev1_) cudaEventDestroy(ev1_);
  for (cudaEvent_t e : prof_pool_) cudaEventDestroy(e);
  for (auto& kv : programs_) persist_free(kv.second);
  if (ctx_.st) cudaStreamDestroy(ctx_.st);
}

// ---------------------------------------------------------------------------------------------
// parameter I/O.  External layout = Torch layout (conv weights (Cout,Cin,kH,kW)); native layout stores conv
// weights as (Cout,kH,kW,Cin) so the K dimension of the implicit GEMM is channel-contiguous.
static void permute_conv(const float* in, float* out, int cout, int cin, int kk, bool to_native) {
  for (int co = 0; co < cout; co++)
    for (int ci = 0; ci < cin; ci++)
      for (int t = 0; t < kk; t++) {
        int64_t ext = ((int64_t)co * cin + ci) * kk + t, nat = ((int64_t)co * kk + t) * cin + ci;
        if (to_native) out[nat] = in[ext]; else out[ext] = in[nat];
      }
}

void Engine::set_params(int group, const float* host, int64_t n) {
  AOCR_CHECK(group >= 0 && group < 5, "group must be in [0,5)");
  AOCR_CHECK(n == L.gsize[group], "parameter vector length does not match the group size");
  AOCR_CUDA(cudaSetDevice(device_));
  std::vector<float> tmp(L.gphys[group], 0.0f);
  for (const TensorEntry& e : L.tensors) {
    if (e.group != group) continue;
    float* dst = tmp.data() + (e.phys_off - L.goff[group]);
    if (e.conv_kk) permute_conv(host + e.ext_off, dst, e.conv_cout, e.conv_cin, e.conv_kk, true);
    else memcpy(dst, host + e.ext_off, e.n * sizeof(float));
  }
  AOCR_CUDA(cudaStreamSynchronize(ctx_.st));
  AOCR_CUDA(cudaMemcpy(d_params + L.goff[group], tmp.data(), tmp.size() * sizeof(float), cudaMemcpyHostToDevice));
  mark_weights_dirty();
  params_set_ = true;
}

// Fresh parameters, as the module constructors behind Model:create draw them (model.lua:83-112 -> cnn.lua, LSTM.lua,
// output_projector.lua; reset() semantics [T7]; nn.LinearNoBias: model_utils.lua:68-85): Linear / conv
// weight and bias ~ U(+-1/sqrt(fan_in)), batch-norm gamma ~ U(0,1), beta = 0, LookupTable ~ N(0,1); running statistics
// (0, 1).  Generator: mt19937 with Torch7's transforms (uniform = a + (b-a) * u32 / 2^32, normal = Box-Muller); the draw
// ORDER is this library's tensor order (group by group in the order of model.lua:150), not nngraph's, so a seed does not
// reproduce Torch7's exact weights - distributions only (no Torch7 exists here to compare against).
void Engine::init_params(uint64_t seed) {
  std::mt19937 gen((uint32_t)(seed ^ (seed >> 32)));
  auto uni = [&](double a, double b) { return a + (b - a) * ((double)gen() * (1.0 / 4294967296.0)); };
  auto normal = [&]() {
    const double u1 = uni(0.0, 1.0), u2 = uni(0.0, 1.0);
    return sqrt(-2.0 * log(1.0 - u2)) * cos(2.0 * M_PI * u1);
  };
  for (int g = 0; g < 5; g++) {
    std::vector<float> v((size_t)L.gsize[g]);
    for (const TensorEntry& e : L.tensors) {
      if (e.group != g) continue;
      float* d = v.data() + e.ext_off;
      const double bound = e.fan_in > 0 ? 1.0 / sqrt((double)e.fan_in) : 0.0;
      for (int64_t i = 0; i < e.n; i++) {
        switch (e.init) {
          case 0: case 1: d[i] = (float)uni(-bound, bound); break;
          case 2: d[i] = (float)uni(0.0, 1.0); break;
          case 3: d[i] = 0.f; break;
          default: d[i] = (float)normal(); break;
        }
      }
    }
    set_params(g, v.data(), (int64_t)v.size());
  }
  const int bnc[3] = {256, 512, 512};
  for (int i = 0; i < 3; i++) {
    std::vector<float> z(bnc[i], 0.f), o(bnc[i], 1.f);
    set_bn(i, z.data(), o.data(), bnc[i]);
  }
  params_set_ = true;
}

void Engine::get_flat(bool grads, int group, float* host, int64_t n) {
  AOCR_CHECK(group >= 0 && group < 5, "group must be in [0,5)");
  AOCR_CHECK(n == L.gsize[group], "vector length does not match the group size");
  if (grads) AOCR_CHECK(have_grads_, "no gradients yet: call aocr_forward_backward first");
  AOCR_CUDA(cudaSetDevice(device_));
  AOCR_CUDA(cudaStreamSynchronize(ctx_.st));
  const float* base = grads ? d_grads : d_params;
  std::vector<float> tmp(L.gphys[group]);
  AOCR_CUDA(cudaMemcpy(tmp.data(), base + L.goff[group], tmp.size() * sizeof(float), cudaMemcpyDeviceToHost));
  for (const TensorEntry& e : L.tensors) {
    if (e.group != group) continue;
    const float* srcp = tmp.data() + (e.phys_off - L.goff[group]);
    if (e.conv_kk) permute_conv(srcp, host + e.ext_off, e.conv_cout, e.conv_cin, e.conv_kk, false);
    else memcpy(host + e.ext_off, srcp, e.n * sizeof(float));
  }
}

void Engine::set_bn(int layer, const float* mean, const float* var, int64_t n) {
  AOCR_CHECK(layer >= 0 && layer < 3, "bn layer must be in [0,3)");
  AOCR_CHECK(n == (layer == 0 ? 256 : 512), "bn stat length mismatch");
  AOCR_CUDA(cudaStreamSynchronize(ctx_.st));
  AOCR_CUDA(cudaMemcpy(bn_rmean[layer], mean, n * sizeof(float), cudaMemcpyHostToDevice));
  AOCR_CUDA(cudaMemcpy(bn_rvar[layer], var, n * sizeof(float), cudaMemcpyHostToDevice));
}
void Engine::get_bn(int layer, float* mean, float* var, int64_t n) {
  AOCR_CHECK(layer >= 0 && layer < 3, "bn layer must be in [0,3)");
  AOCR_CHECK(n == (layer == 0 ? 256 : 512), "bn stat length mismatch");
  AOCR_CUDA(cudaStreamSynchronize(ctx_.st));
  AOCR_CUDA(cudaMemcpy(mean, bn_rmean[layer], n * sizeof(float), cudaMemcpyDeviceToHost));
  AOCR_CUDA(cudaMemcpy(var, bn_rvar[layer], n * sizeof(float), cudaMemcpyDeviceToHost));
}

// ---------------------------------------------------------------------------------------------
void Engine::stage_batch(const float* images, int b, int W, const int32_t* tgt, const int32_t* tev, int T) {
  AOCR_CHECK(b >= 1 && b <= Bmax, "batch larger than config.batch_size");
  char msg[128];
  if (T > Tmax) {   // model.lua:264
    snprintf(msg, sizeof(msg), "max_decoder_l (%d) < target_l (%d)!", Tmax, T);
    throw InvalidError(msg);
  }
  AOCR_CHECK(T >= 1, "target_l must be >= 1");
  AOCR_CHECK(W >= 8, "image too narrow");
  int S = (W / 2) / 2 - 1;
  if (S > Smax) {   // model.lua:287
    snprintf(msg, sizeof(msg), "max_encoder_l (%d) < source_l (%d)!", Smax, S);
    throw InvalidError(msg);
  }
  AOCR_CHECK(W <= Wmax, "image width out of range for max_encoder_l");
  AOCR_CHECK(S >= 1, "image too narrow: source_l < 1");
  for (int64_t i = 0; i < (int64_t)b * T; i++)
    AOCR_CHECK(tgt[i] >= 1 && tgt[i] <= V && tev[i] >= 1 && tev[i] <= V, "token id outside [1, target_vocab_size]");
  AOCR_CUDA(cudaSetDevice(device_));
  b_ = b; W_ = W; T_ = T; W1_ = W / 2; W2_ = W1_ / 2; S_ = S;
  AOCR_CUDA(cudaMemcpyAsync(x0, images, (size_t)b * 32 * W * sizeof(float), cudaMemcpyHostToDevice, ctx_.st));
  AOCR_CUDA(cudaMemcpyAsync(tgt_bt, tgt, (size_t)b * T * sizeof(int32_t), cudaMemcpyHostToDevice, ctx_.st));
  AOCR_CUDA(cudaMemcpyAsync(tev_bt, tev, (size_t)b * T * sizeof(int32_t), cudaMemcpyHostToDevice, ctx_.st));
  h_tev_.assign(tev, tev + (size_t)b * T);
  have_batch_ = true;
}

// Per-class device timing for bench.py's roofline: CUDA events recorded on the engine stream around every call
// of a class, no host synchronisation (the elapsed times are read once, after the step).
void Engine::prof_begin(int cls) {
  if (!prof_on) return;
  if (prof_used_ * 2 + 2 > prof_pool_.size()) {
    for (int i = 0; i < 512; i++) {
      cudaEvent_t e;
      AOCR_CUDA(cudaEventCreate(&e));
      prof_pool_.push_back(e);
    }
  }
  AOCR_CUDA(cudaEventRecord(prof_pool_[prof_used_ * 2], ctx_.st));
  prof_open_.push_back(prof_used_);
  prof_recs_.push_back({cls, -1.0});
  prof_used_++;
}
void Engine::prof_end(int cls, double work) {
  if (!prof_on) return;
  if (prof_open_.empty()) return;
  const size_t slot = prof_open_.back();
  prof_open_.pop_back();
  AOCR_CUDA(cudaEventRecord(prof_pool_[slot * 2 + 1], ctx_.st));
  prof_recs_[slot].second = work;
}
void Engine::use_lane(int i) {
  if (i == cur_lane_) return;
  Lane& o = lanes_[cur_lane_];
  o.st = ctx_.st; o.tc_ws = ctx_.tc_ws; o.scratch[0] = scratch_[0]; o.scratch[1] = scratch_[1]; o.partial = partial; o.tmpvec = tmpvec;
  const Lane& n = lanes_[i];
  ctx_.st = n.st; ctx_.tc_ws = n.tc_ws; scratch_[0] = n.scratch[0]; scratch_[1] = n.scratch[1]; partial = n.partial; tmpvec = n.tmpvec;
  cur_lane_ = i;
}
// lane i starts after everything enqueued so far on lane 0
void Engine::fork_to(int i) {
  if (!lanes_on_) return;
  cudaEvent_t ev = lane_ev_[lane_ev_next_++ % 8];
  AOCR_CUDA(cudaEventRecord(ev, lanes_[0].st));
  AOCR_CUDA(cudaStreamWaitEvent(lanes_[i].st, ev, 0));
}
// lane 0 continues after everything enqueued so far on lane i
void Engine::join_from(int i) {
  if (!lanes_on_) return;
  cudaEvent_t ev = lane_ev_[lane_ev_next_++ % 8];
  AOCR_CUDA(cudaEventRecord(ev, lanes_[i].st));
  AOCR_CUDA(cudaStreamWaitEvent(lanes_[0].st, ev, 0));
}
void Engine::phase_mark(const char* name) {
  if (!phases_on_) return;
  cudaEvent_t e;
  AOCR_CUDA(cudaEventCreate(&e));
  AOCR_CUDA(cudaEventRecord(e, ctx_.st));
  phase_marks_.push_back({name, e});
}
void Engine::phase_report() {
  if (!phases_on_ || phase_marks_.size() < 2) return;
  AOCR_CUDA(cudaStreamSynchronize(ctx_.st));
  fprintf(stderr, "[aocr phases]");
  for (size_t i = 1; i < phase_marks_.size(); i++) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, phase_marks_[i - 1].second, phase_marks_[i].second);
    fprintf(stderr, " %s=%.3fms", phase_marks_[i].first.c_str(), ms);
  }
  float tot = 0.f;
  cudaEventElapsedTime(&tot, phase_marks_.front().second, phase_marks_.back().second);
  fprintf(stderr, " total=%.3fms\n", tot);
  for (auto& m : phase_marks_) cudaEventDestroy(m.second);
  phase_marks_.clear();
}
void Engine::prof_collect() {
  AOCR_CUDA(cudaStreamSynchronize(ctx_.st));
  for (int l = 1; l < 3; l++) if (lanes_on_ && lanes_[l].st) AOCR_CUDA(cudaStreamSynchronize(lanes_[l].st));
  for (size_t i = 0; i < prof_recs_.size(); i++) {
    if (prof_recs_[i].second < 0) continue;
    float ms = 0.f;
    AOCR_CUDA(cudaEventElapsedTime(&ms, prof_pool_[2 * i], prof_pool_[2 * i + 1]));
    if (prof_dump_)
      fprintf(stderr, "[aocr prof] #%zu class %d  %.1f us  %.3f GFLOP(or GB)  -> %.1f T/s\n", i, prof_recs_[i].first, ms * 1e3,
              prof_recs_[i].second * 1e-9, prof_recs_[i].second / (ms * 1e-3) * 1e-12);
    prof_ms[prof_recs_[i].first] += ms;
    prof_launches[prof_recs_[i].first] += 1;
    prof_work[prof_recs_[i].first] += prof_recs_[i].second;
  }
  prof_recs_.clear();
  prof_open_.clear();
  prof_used_ = 0;
}

void Engine::conv_dims(int l, int& Hin, int& Win, int& Hout, int& Wout) const {
  // l = 1..6 (0-based index into kConv): input of conv_{l+1}
  static const int hin[7] = {32, 16, 8, 8, 4, 4, 2};
  Hin = hin[l];
  Win = l == 0 ? W_ : (l == 1 ? W1_ : W2_);
  Hout = Hin + 2 * kConv[l].pad - kConv[l].k + 1;
  Wout = Win + 2 * kConv[l].pad - kConv[l].k + 1;
}

// Everything derived from the parameters after an update (flipped convolution weights, fused biases, the embedding
// table of the decoder's first layer, the decoder's concatenated / transposed weight planes).  None of it is needed by
// the CNN forward, so it is built on lane 1 while the CNN forward runs; the caller joins before the encoder.
bool Engine::prep_weights() {
  if (!weights_dirty_) return false;
  fork_to(1);
  use_lane(1);
  static const bool prewarm = !(getenv("AOCR_PREWARM") && atoi(getenv("AOCR_PREWARM")) == 0);
  const bool tcm = cfg.gemm_mode != 2 && prewarm;
  if (tcm && lanes_on_) {
    // the forward convolutions need their weight planes first (conv2 right after conv1): converted here, ahead of
    // everything else on this lane, and lane 0 waits for exactly this point before its first convolution GEMM - not
    // lazily in front of every GEMM on the critical path
    for (int l = 1; l < 7; l++) {
      const ConvSpec& c = kConv[l];
      const int Kc = c.k * c.k * c.cin;
      operand_pack(d_params + L.conv_w[l], c.cout, Kc, Kc, 1, 1);
    }
    AOCR_CUDA(cudaEventRecord(conv_packs_ev_, ctx_.st));
    conv_packs_pending_ = true;
  }
  for (int l = 1; l < 7; l++) {
    const ConvSpec& c = kConv[l];
    int64_t total = (int64_t)c.cout * c.k * c.k * c.cin;
    conv_wt_kernel<<<cdiv(total, 256) < 1184 ? cdiv(total, 256) : 1184, 256, 0, ctx_.st>>>(d_params + L.conv_w[l], wt[l],
                                                                                       c.cout, c.k * c.k, c.cin);
    AOCR_LAUNCH_CHECK(ctx_);
  }
  for (int d = 0; d < 2; d++)
    add_vec(ctx_, encb + d * 4 * He, d_params + L.enc_bi[d], d_params + L.enc_bh[d], 4 * He);
  add_vec(ctx_, bsum1, d_params + L.l1_bi, d_params + L.l1_bh, 4 * Hd);
  add_vec(ctx_, bsum2, d_params + L.l2_bi, d_params + L.l2_bh, 4 * Hd);
  // P[v] = E[v] W_i1[:, :E]^T + b_i1 + b_h1 : the embedding half of the layer-1 input projection (K12)
  const int in1 = E + (cfg.input_feed ? Hd : 0);
  Gemm g;
  g.M = V; g.N = 4 * Hd; g.K = E;
  g.A = d_params + L.emb; g.sam = E; g.sak = 1;
  g.B = d_params + L.l1_wi; g.sbk = 1; g.sbn = in1;
  g.C = Ptab; g.ldc = 4 * Hd; g.bias_n = bsum1;
  gemm_simt(ctx_, g);
  if (cfg.gemm_mode != 2) build_decoder_packs();
  if (tcm && lanes_on_) {
    // every other parameter-derived operand of the step, so that no conversion launch sits in front of a GEMM on the
    // critical path: encoder input projection and recurrent weights, W_c1 of the attention precompute, the flipped
    // convolution weights of the data gradients (all joined before the encoder starts)
    ensure_wicat_pack();
    ensure_enc_packs();
    operand_pack(d_params + L.wc, Hd, Hd, 2 * Hd, 1, 1);
    for (int l = 1; l < 7; l++) {
      const ConvSpec& c = kConv[l];
      const int Kd = c.k * c.k * c.cout;
      operand_pack(wt[l], c.cin, Kd, Kd, 1, 1);
    }
  }
  use_lane(0);
  weights_dirty_ = false;
  return true;      // lane 1 carries the work: join before the encoder
}

// ---------------------------------------------------------------------------------------------
// CNN forward (src/model/cnn.lua:9-45; called model.lua:285)
void Engine::cnn_forward(bool train) {
  cnn_train_ = train;
  mbox_slot_next_ = 0;        // the statistics exchanges of a step use mailbox slots 0, 1, ... in issue order
  const int B = b_;
  const bool tc = cfg.gemm_mode != 2;
  conv1_fwd(ctx_, x0, d_params + L.conv_w[0], d_params + L.conv_b[0], act[1], pidx[1], B, W_, tc ? actp_[1].hi : nullptr,
            tc ? actp_[1].lo : nullptr);
  if (conv_packs_pending_) {      // the weight planes of the forward convolutions are being written on lane 1 (prep_weights)
    AOCR_CUDA(cudaStreamWaitEvent(ctx_.st, conv_packs_ev_, 0));
    conv_packs_pending_ = false;
  }
  for (int l = 1; l < 7; l++) {
    const ConvSpec& c = kConv[l];
    int Hin, Win, Hout, Wout;
    conv_dims(l, Hin, Win, Hout, Wout);
    const int64_t rows = (int64_t)B * Hout * Wout;
    const int Kc = c.k * c.k * c.cin;
    if (cfg.gemm_mode != 2) {
      // implicit GEMM: the NHWC activation is the A operand (4-D tensor map, halo by TMA zero fill); no im2col
      conv_tc(act[l], B, Hin, Win, c.cin, c.k, c.pad, Hout, Wout, d_params + L.conv_w[l], c.cout, zb[l + 1],
              d_params + L.conv_b[l], &actp_[l]);
    } else {
      im2col(ctx_, act[l], col, B, Hin, Win, c.cin, c.k, c.pad);
      Gemm g;
      g.M = (int)rows; g.N = c.cout; g.K = Kc;
      g.A = col; g.sam = Kc; g.sak = 1;
      g.B = d_params + L.conv_w[l]; g.sbk = 1; g.sbn = Kc;
      g.C = zb[l + 1]; g.ldc = c.cout; g.bias_n = d_params + L.conv_b[l];
      gemm(g);
    }
    if (c.bn >= 0) {
      const float *mean, *var;
      if (train) {
        bn_stats(ctx_, zb[l + 1], rows, c.cout, bn_mean[c.bn], bn_var[c.bn], bn_rmean[c.bn], bn_rvar[c.bn], partial, tmpvec,
                 stat_sync());
        mean = bn_mean[c.bn]; var = bn_var[c.bn];
      } else {
        mean = bn_rmean[c.bn]; var = bn_rvar[c.bn];
      }
      const bool last = (l == 6);
      // the next contraction's operand planes come straight from this kernel (last layer: the encoder input projection)
      __nv_bfloat16* ph = tc ? (last ? srcP_.hi : actp_[l + 1].hi) : nullptr;
      __nv_bfloat16* pl = tc ? (last ? srcP_.lo : actp_[l + 1].lo) : nullptr;
      bn_relu_fwd(ctx_, zb[l + 1], mean, var, d_params + L.bn_g[c.bn], d_params + L.bn_b[c.bn], act[l + 1], rows, c.cout,
                  last ? S_ : 0, last ? B : 0, ph, pl);
    } else {
      relu_pool_fwd(ctx_, zb[l + 1], act[l + 1], pidx[l + 1], B, Hout, Wout, c.cout, c.pool_kw, tc ? actp_[l + 1].hi : nullptr,
                    tc ? actp_[l + 1].lo : nullptr);
    }
  }
  taps_["cnn_out"] = {src, (int64_t)S_ * B * 512};
  // per-layer taps: activations (NHWC; act7 = cnn_out, time-major) and the pooling choices (0..3 = dy*2+dx)
  {
    static const char* an[8] = {"", "act1", "act2", "act3", "act4", "act5", "act6", "act7"};
    static const char* pn[8] = {"", "pidx1", "pidx2", "", "pidx4", "", "pidx6", ""};
    const int64_t n_act[8] = {0, (int64_t)B * 16 * W1_ * 64, (int64_t)B * 8 * W2_ * 128, (int64_t)B * 8 * W2_ * 256,
                              (int64_t)B * 4 * W2_ * 256, (int64_t)B * 4 * W2_ * 512, (int64_t)B * 2 * W2_ * 512,
                              (int64_t)S_ * B * 512};
    for (int l = 1; l <= 7; l++) {
      taps_[an[l]] = {act[l], n_act[l]};
      if (pidx[l]) taps_[pn[l]] = {pidx[l], n_act[l], true};
    }
  }
}

// CNN backward (model.lua:692)
void Engine::cnn_backward() {
  const int B = b_;
  const float* dcur = dsrc;   // gradient wrt act[l+1]
  for (int l = 6; l >= 1; l--) {
    const ConvSpec& c = kConv[l];
    int Hin, Win, Hout, Wout;
    conv_dims(l, Hin, Win, Hout, Wout);
    const int64_t rows = (int64_t)B * Hout * Wout;
    const int Kc = c.k * c.k * c.cin;
    // dz as bf16 planes, written by the kernel that produces dz: shared by the weight- and the data-gradient GEMM
    const bool tc = cfg.gemm_mode != 2;
    float* dz = tc ? dzf_[l] : gA;      // per layer in tensor-core mode: read by the side lane after lane 0 has moved on
    Pack dzp;
    dzp.rows = rows; dzp.kp = c.cout; dzp.hi = tc ? dzp_[l].hi : nullptr; dzp.lo = tc ? dzp_[l].lo : nullptr;
    if (c.bn >= 0) {
      const bool last = (l == 6);
      const float* mean = cnn_train_ ? bn_mean[c.bn] : bn_rmean[c.bn];
      const float* var = cnn_train_ ? bn_var[c.bn] : bn_rvar[c.bn];
      bn_relu_bwd(ctx_, dcur, act[l + 1], zb[l + 1], mean, var, d_params + L.bn_g[c.bn], dz, d_grads + L.bn_g[c.bn],
                  d_grads + L.bn_b[c.bn], partial, rows, c.cout, last ? S_ : 0, last ? B : 0, cnn_train_ ? 1 : 0,
                  cnn_train_ ? stat_sync() : StatSync(), dzp.hi, dzp.lo);
    } else {
      relu_pool_bwd(ctx_, dcur, act[l + 1], pidx[l + 1], dz, B, Hout, Wout, c.cout, c.pool_kw, dzp.hi, dzp.lo);
    }
    // weight grad: dW[co][tap,ci] = sum_rows dz[row][co] * col[row][tap,ci]
    if (cfg.gemm_mode != 2) {
      // the weight and bias gradients feed nothing downstream: lane 2 (idle since the encoder backward), so they fill the
      // partial waves of the data-gradient chain that continues on lane 0; operands are per-layer buffers (no reuse race)
      fork_to(2);
      use_lane(2);
      col_sum(ctx_, dz, rows, c.cout, d_grads + L.conv_b[l], partial, 0);     // bias grad (this lane's partial buffer)
      conv_wgrad_tc(dz, act[l], B, Hin, Win, c.cin, c.k, c.pad, Hout, Wout, c.cout, d_grads + L.conv_w[l], &dzp, &actp_[l]);
      if (l == 4 && cnn_bucket_split_ >= 0) grad_range(cnn_bucket_split_, L.goff[G_CNN] + L.gphys[G_CNN]);
      // conv2..conv4 (+bn3) are complete once this lane has run conv2's weight gradient: their bucket leaves now, so
      // that only conv1's 640 gradients remain for after the last kernel of the backward
      if (l == 1 && cnn_bucket_split_ >= 0) grad_range(L.conv_w[1], cnn_bucket_split_);
      use_lane(0);
    } else {
      col_sum(ctx_, dz, rows, c.cout, d_grads + L.conv_b[l], partial, 0);     // bias grad
      im2col(ctx_, act[l], col, B, Hin, Win, c.cin, c.k, c.pad);
      Gemm gw;
      gw.M = c.cout; gw.N = Kc; gw.K = (int)rows;
      gw.A = dz; gw.sam = 1; gw.sak = c.cout;
      gw.B = col; gw.sbk = Kc; gw.sbn = 1;
      gw.C = d_grads + L.conv_w[l]; gw.ldc = Kc;
      gemm(gw);
    }
    // data grad: correlation of dz with flipped, in/out-swapped weights, padding k-1-pad
    const int padd = c.k - 1 - c.pad;
    if (cfg.gemm_mode != 2) {
      conv_tc(dz, B, Hout, Wout, c.cout, c.k, padd, Hin, Win, wt[l], c.cin, gB, nullptr, &dzp);
    } else {
      im2col(ctx_, dz, col, B, Hout, Wout, c.cout, c.k, padd);
      const int Kd = c.k * c.k * c.cout;
      const int64_t rows_in = (int64_t)B * Hin * Win;
      Gemm gd;
      gd.M = (int)rows_in; gd.N = c.cin; gd.K = Kd;
      gd.A = col; gd.sam = Kd; gd.sak = 1;
      gd.B = wt[l]; gd.sbk = 1; gd.sbn = Kd;
      gd.C = gB; gd.ldc = c.cin;
      gemm(gd);
    }
    dcur = gB;
    // next iteration writes dz into gA again and reads dcur=gB: fine (distinct buffers)
  }
  const int nblk = 8 * 148;      // ~43 pooled pixels per block at batch 64: the per-thread loop is a chain of dependent loads
  conv1_bwd(ctx_, x0, act[1], pidx[1], dcur, d_grads + L.conv_w[0], d_grads + L.conv_b[0], partial, nblk, B, W_);
  if (cnn_bucket_split_ >= 0) grad_small(L.goff[G_CNN], L.conv_w[1]);     // conv1: the tail of the gradient exchange
  join_from(2);
}

// operand planes of [W_i fw ; W_i bw] (the encoder's time-batched input projection); refreshed once per weight update
void Engine::ensure_wicat_pack() {
  if (wicat_version_ == weights_version_) return;
  for (int d = 0; d < 2; d++) {
    Pack half = WiCatP;
    half.hi += (int64_t)d * 4 * He * 512; half.lo += (int64_t)d * 4 * He * 512; half.rows = 4 * He;
    split_to_pack(ctx_, d_params + L.enc_wi[d], 4 * He, 512, 512, 1, half);
  }
  wicat_version_ = weights_version_;
}

// ---------------------------------------------------------------------------------------------
// encoder (model.lua:293-316).  Slot convention: fw slot t+1 = state after column t (slot 0 = zeros);
// bw slot t = state after column t (slot S = zeros).
void Engine::encoder_forward() {
  const int B = b_, S = S_;
  if (cfg.gemm_mode != 2) {
    // time-batched input projection of BOTH directions as one GEMM: [W_i fw ; W_i bw] stacked on the UMMA M side,
    // the CNN output (operand planes written by the last batch-norm kernel) on the N side
    ensure_wicat_pack();
    prof_begin(0);
    TcGemm t;
    Pack sp = srcP_; sp.rows = (int64_t)S * B;
    t.A = WiCatP; t.B = sp; t.M = 8 * He; t.N = S * B; t.K = 512;
    t.C = xg; t.ldc = 8 * He; t.transpose_out = true; t.bias_m = encb;
    t.terms = cfg.gemm_mode == 1 ? 1 : 3;
    gemm_tc(ctx_, t);
    prof_end(0, 2.0 * S * B * 8.0 * He * 512);
  }
  for (int d = 0; d < 2 && cfg.gemm_mode == 2; d++) {   // time-batched input projection for all columns (K10)
    Gemm g;
    g.M = S * B; g.N = 4 * He; g.K = 512;
    g.A = src; g.sam = 512; g.sak = 1;
    g.B = d_params + L.enc_wi[d]; g.sbk = 1; g.sbn = 512;
    g.C = xg + d * 4 * He; g.ldc = 8 * He; g.bias_n = encb + d * 4 * He;
    gemm(g);
  }
  const int64_t slot = (int64_t)B * He;
  fill_zero(ctx_, Henc, slot * sizeof(float));                                   // fw slot 0
  fill_zero(ctx_, Cenc, slot * sizeof(float));
  fill_zero(ctx_, Henc + ((int64_t)(S + 1) + S) * slot, slot * sizeof(float));   // bw slot S
  fill_zero(ctx_, Cenc + ((int64_t)(S + 1) + S) * slot, slot * sizeof(float));
  if (cfg.gemm_mode != 2) {
    encoder_forward_steps_tc();
  } else {
    EncStep p;
    p.xg = xg; p.Wh[0] = d_params + L.enc_wh[0]; p.Wh[1] = d_params + L.enc_wh[1];
    p.H = Henc; p.Cst = Cenc; p.acts = acts_enc; p.ctx = ctx; p.B = B; p.S = S; p.He = He;
    prof_begin(2);
    for (int i = 0; i < S; i++) {
      p.step = i;
      enc_step_fwd(ctx_, p);
    }
    prof_end(2, 2.0 * 2 * S * (double)B * He * 4 * He);
  }
  taps_["context"] = {ctx, (int64_t)B * S * 2 * He};
}

void Engine::encoder_backward() {
  const int B = b_, S = S_;
  const int64_t slot = (int64_t)B * He;
  // seeds: halves of d c1(0), d h1(0) (model.lua:666-667,680-681).  dc1 / dX1[:, h1off:] hold them.
  for (int d = 0; d < 2; d++) {
    copy_strided(ctx_, enc_dc + d * slot, He, dc1 + d * He, Hd, B, He);
    copy_strided(ctx_, enc_dh + d * slot, He, dX1 + h1off + d * He, K1, B, He);
  }
  if (cfg.gemm_mode != 2) {
    encoder_backward_steps_tc();
  } else {
    EncStepBwd p;
    p.Wh[0] = d_params + L.enc_wh[0]; p.Wh[1] = d_params + L.enc_wh[1];
    p.Cst = Cenc; p.acts = acts_enc; p.Dctx = Dctx; p.dh = enc_dh; p.dc = enc_dc; p.dG = dGe;
    p.B = B; p.S = S; p.He = He;
    for (int i = 0; i < S; i++) {
      p.step = i;
      enc_cell_bwd(ctx_, p);
      // dh_prev[d] = dG_t[d] W_h[d]   (B x 4He) x (4He x He)
      for (int d = 0; d < 2; d++) {
        int t = d == 0 ? S - 1 - i : i;
        Gemm g;
        g.M = B; g.N = He; g.K = 4 * He;
        g.A = dGe + (int64_t)t * B * 8 * He + d * 4 * He; g.sam = 8 * He; g.sak = 1;
        g.B = d_params + L.enc_wh[d]; g.sbk = He; g.sbn = 1;
        g.C = enc_dh + d * slot; g.ldc = He;
        gemm(g, 2);
      }
    }
  }
  // d src = dG_fw W_i_fw + dG_bw W_i_bw  (model.lua:675,689): needed next by the CNN backward, stays on lane 0
  for (int d = 0; d < 2; d++) {
    Gemm g;
    g.M = S * B; g.N = 512; g.K = 4 * He;
    g.A = dGe + d * 4 * He; g.sam = 8 * He; g.sak = 1;
    g.B = d_params + L.enc_wi[d]; g.sbk = 512; g.sbn = 1;
    g.C = dsrc; g.ldc = 512; g.accumulate = d;
    if (cfg.gemm_mode != 2) g.pa = &dGeP[d];       // written by the cell-backward bodies, rows = time*batch
    gemm(g);
  }
  // time-batched parameter gradients: independent of the CNN backward -> lane 1
  fork_to(1);
  use_lane(1);
  for (int d = 0; d < 2; d++) {
    const float* dG = dGe + d * 4 * He;
    Gemm gi;   // dW_i = dG^T src
    gi.M = 4 * He; gi.N = 512; gi.K = S * B;
    gi.A = dG; gi.sam = 1; gi.sak = 8 * He;
    gi.B = src; gi.sbk = 512; gi.sbn = 1;
    gi.C = d_grads + L.enc_wi[d]; gi.ldc = 512;
    Pack hprev = HencP[d];                          // h_prev(t): slot t (fw) / slot t+1 (bw)
    if (d == 1) { hprev.hi += (int64_t)B * He; hprev.lo += (int64_t)B * He; }
    if (cfg.gemm_mode != 2) { gi.pa = &dGeP[d]; gi.pb = &srcP_; }
    gemm(gi);
    // dW_h = sum_t dG_t^T h_prev(t).  fw: h_prev(t) = slot t -> rows [0, S*B) of Henc[0]; bw: h_prev(t) = slot t+1.
    Gemm gh;
    gh.M = 4 * He; gh.N = He; gh.K = S * B;
    gh.A = dG; gh.sam = 1; gh.sak = 8 * He;
    gh.B = Henc + (int64_t)d * (S + 1) * slot + (d == 0 ? 0 : slot); gh.sbk = He; gh.sbn = 1;
    gh.C = d_grads + L.enc_wh[d]; gh.ldc = He;
    if (cfg.gemm_mode != 2) { gh.pa = &dGeP[d]; gh.pb = &hprev; }
    gemm(gh);
  }
  // bias grads: b_i and b_h both receive the column sums of dG (two biases per cell, LSTM.lua:79-87)
  col_sum(ctx_, dGe, (int64_t)S * B, 8 * He, tmpvec, partial, 0);
  for (int d = 0; d < 2; d++) {
    AOCR_CUDA(cudaMemcpyAsync(d_grads + L.enc_bi[d], tmpvec + d * 4 * He, (size_t)4 * He * sizeof(float),
                              cudaMemcpyDeviceToDevice, ctx_.st));
    AOCR_CUDA(cudaMemcpyAsync(d_grads + L.enc_bh[d], tmpvec + d * 4 * He, (size_t)4 * He * sizeof(float),
                              cudaMemcpyDeviceToDevice, ctx_.st));
  }
  use_lane(0);
}

}  // namespace aocr
