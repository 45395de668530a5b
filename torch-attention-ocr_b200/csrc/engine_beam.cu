// engine_beam.cu — forward_only step with beam_size > 1 and / or a dictionary trie
// (src/model/model.lua:226-251,321-536,570-633; trie: src/utils/utils.lua:177-218).
// CNN, encoder and the teacher-forced gold pass run exactly as in the greedy step; the beam pass runs the decoder on
// beam*Bc replicated rows (beam-major, see kernels_beam.cu) in chunks of Bc images such that beam*Bc fits the decoder
// state (2 x batch_size rows), one per-kernel decoder step + generator + selection + parent re-gather per timestep.
#include "engine.h"

#include <stdio.h>

#include <string>
#include <unordered_map>

namespace aocr {

void Engine::decode_beam_enqueue(int beam, const int32_t* trie_host, int32_t trie_nodes) {
  AOCR_CHECK(have_batch_, "no batch staged");
  AOCR_CHECK(params_set_, "the model has no parameters yet: call aocr_init_params or aocr_set_params first");
  AOCR_CHECK(cfg.gemm_mode != 2, "beam search runs on the tensor-core path (gemm_mode 0 or 1)");
  AOCR_CHECK(beam >= 1, "beam_size must be >= 1");
  AOCR_CUDA(cudaSetDevice(device_));
  const int B = b_, T = T_, Ld = Tmax;
  const int K = beam < V ? beam : V;                                   // model.lua:229
  AOCR_CHECK(K <= 2 * Bmax, "beam_size exceeds the decoder state capacity (2 x batch_size rows)");
  // ---- dictionary: flat child table, uploaded once per (pointer, size)
  if (trie_host) {
    AOCR_CHECK(trie_nodes >= 1, "empty trie");
    if (trie_host != trie_src_ || trie_nodes != trie_nodes_) {
      AOCR_CUDA(cudaStreamSynchronize(ctx_.st));
      if (d_trie_) cudaFree(d_trie_);
      const size_t bytes = (size_t)trie_nodes * (V + 1) * sizeof(int32_t);
      AOCR_CUDA(cudaMalloc(&d_trie_, bytes));
      AOCR_CUDA(cudaMemcpy(d_trie_, trie_host, bytes, cudaMemcpyHostToDevice));
      trie_src_ = trie_host; trie_nodes_ = trie_nodes;
    }
  }
  const int Rmax = 2 * Bmax;
  if (!beam_tmp_) {   // first beam call: scratch for the selection / re-gather (sized for 2 x batch_size rows)
    const int64_t row_bytes = (int64_t)K1 * 4 + (int64_t)K1 * 2 * 2 + (int64_t)Hd * 4 + (int64_t)Hd * 2 * 2 + 2 * (int64_t)Hd * 4;
    beam_tmp_ = alloc<uint8_t>(Rmax * row_bytes);
    beam_logp_ = alloc<float>((int64_t)Rmax * V);
    beam_scores_[0] = alloc<double>(Rmax); beam_scores_[1] = alloc<double>(Rmax);
    beam_tok_[0] = alloc<int32_t>(Rmax); beam_tok_[1] = alloc<int32_t>(Rmax);
    beam_parent_ = alloc<int32_t>(Rmax);
    beam_loc_[0] = alloc<int32_t>(Rmax); beam_loc_[1] = alloc<int32_t>(Rmax);
    beam_hist_tok_ = alloc<int32_t>((int64_t)Tmax * Rmax); beam_hist_par_ = alloc<int32_t>((int64_t)Tmax * Rmax);
  }
  prep_weights();
  gather_tokens(ctx_, tgt_bt, tgt_tb, B, T, Ld, 1);     // model.lua:266-274: pad to max_decoder_l with PAD
  gather_tokens(ctx_, tev_bt, tev_tb, B, T, Ld, 1);
  cnn_forward(false);
  join_from(1);
  encoder_forward();
  attention_precompute();
  dec_steps_ = Ld;
  // ---- gold pass first (model.lua:589-627): its outputs (row losses, log-probs) are final before the beam pass reuses
  // the decoder state
  decoder_init();
  if (persist_on_) run_program(PK_DEC_FWD, Ld, 0);
  else for (int t = 0; t < Ld; t++) decoder_step(t, tgt_tb + (int64_t)t * B);
  generator_fwd(ctx_, A_all, d_params + L.wo, d_params + L.bo, tev_tb, logp[2], nullptr, rowloss, (int64_t)Ld * B, Hd, V, 1.0f);
  reduce_sum_double(ctx_, rowloss, (int64_t)Ld * B, d_loss);
  last_logp_rows_[2] = Ld * B;
  last_logp_rows_[1] = 0;
  // ---- beam pass: per-kernel decoder steps on the plain (not gate-interleaved) weight planes
  const bool persist_save = persist_on_;
  persist_on_ = false;
  try {
    build_decoder_packs();
    const int Bc_max = Rmax / K;
    const int64_t slot = (int64_t)B * He;
    for (int b0 = 0; b0 < B; b0 += Bc_max) {
      const int Bc = B - b0 < Bc_max ? B - b0 : Bc_max;
      const int R = K * Bc;
      // initial state of every replica (model.lua:360-372): c1(0) = [c_fw(S); c_bw(1)], h1(0) = 0 (Q14) or the finals
      b_ = R; dual_rows_ = Bc; ctx_row0_ = b0;
      fill_zero(ctx_, X1, (size_t)R * K1 * sizeof(float));
      for (int k = 0; k < K; k++) {
        const float* cfw = Cenc + ((int64_t)0 * (S_ + 1) + S_) * slot + (int64_t)b0 * He;
        const float* cbw = Cenc + ((int64_t)1 * (S_ + 1) + 0) * slot + (int64_t)b0 * He;
        concat_enc_finals(ctx_, cfw, cbw, C1 + (int64_t)k * Bc * Hd, Hd, Bc, He);
        if (!cfg.input_feed) {
          const float* hfw = Henc + ((int64_t)0 * (S_ + 1) + S_) * slot + (int64_t)b0 * He;
          const float* hbw = Henc + ((int64_t)1 * (S_ + 1) + 0) * slot + (int64_t)b0 * He;
          concat_enc_finals(ctx_, hfw, hbw, X1 + (int64_t)k * Bc * K1, K1, Bc, He);
        }
      }
      fill_zero(ctx_, C2, (size_t)R * Hd * sizeof(float));
      fill_zero(ctx_, X2, (size_t)R * 2 * Hd * sizeof(float));
      Pack x1_0 = X1p; x1_0.rows = R;
      split_to_pack(ctx_, X1, R, K1, K1, 1, x1_0);
      fill_zero(ctx_, X2p.hi, (size_t)R * 2 * Hd * sizeof(__nv_bfloat16));
      fill_zero(ctx_, X2p.lo, (size_t)R * 2 * Hd * sizeof(__nv_bfloat16));
      // GO token for every replica (model.lua:388)
      for (int k = 0; k < K; k++)
        AOCR_CUDA(cudaMemcpyAsync(beam_tok_[0] + (int64_t)k * Bc, tgt_tb + b0, (size_t)Bc * sizeof(int32_t),
                                  cudaMemcpyDeviceToDevice, ctx_.st));
      int cur = 0;
      for (int t = 0; t < Ld; t++) {
        decoder_step(t, beam_tok_[cur]);
        generator_fwd(ctx_, A_all + (int64_t)t * R * Hd, d_params + L.wo, d_params + L.bo, nullptr, beam_logp_, nullptr, nullptr, R,
                      Hd, V, 1.0f);
        BeamSelect s;
        s.logp = beam_logp_; s.tok = beam_tok_[cur]; s.scores = beam_scores_[cur]; s.new_scores = beam_scores_[cur ^ 1];
        s.tok_out = beam_tok_[cur ^ 1]; s.parent_row = beam_parent_;
        s.hist_tok = beam_hist_tok_; s.hist_par = beam_hist_par_;
        s.trie = trie_host ? d_trie_ : nullptr; s.loc = beam_loc_[cur]; s.new_loc = beam_loc_[cur ^ 1];
        s.t = t; s.Bc = Bc; s.K = K; s.V = V;
        beam_select(ctx_, s);
        if (t + 1 < Ld) {
          // every state tensor of slot t+1 follows its parent (model.lua:516-534): [a_t | h1_t] (fp32 + operand planes),
          // h2_t (its half of X2, fp32 + planes), c1, c2
          const int64_t r1 = (int64_t)(t + 1) * R;
          BeamGather g;
          int n = 0;
          int64_t off = 0;
          auto add = [&](void* ptr, int64_t pitch, int64_t bytes) {
            g.t[n].ptr = reinterpret_cast<uint8_t*>(ptr); g.t[n].pitch_bytes = pitch; g.t[n].row_bytes = bytes; g.t[n].tmp_off = off;
            off += (int64_t)R * bytes; n++;
          };
          add(X1 + r1 * K1, (int64_t)K1 * 4, (int64_t)K1 * 4);
          add(X1p.hi + r1 * X1p.kp, X1p.kp * 2, (int64_t)K1 * 2);
          add(X1p.lo + r1 * X1p.kp, X1p.kp * 2, (int64_t)K1 * 2);
          add(X2 + r1 * 2 * Hd + Hd, (int64_t)2 * Hd * 4, (int64_t)Hd * 4);
          add(X2p.hi + r1 * X2p.kp + Hd, X2p.kp * 2, (int64_t)Hd * 2);
          add(X2p.lo + r1 * X2p.kp + Hd, X2p.kp * 2, (int64_t)Hd * 2);
          add(C1 + r1 * Hd, (int64_t)Hd * 4, (int64_t)Hd * 4);
          add(C2 + r1 * Hd, (int64_t)Hd * 4, (int64_t)Hd * 4);
          g.n = n; g.rows = R; g.parent_row = beam_parent_; g.tmp = beam_tmp_;
          beam_gather(ctx_, g);
        }
        cur ^= 1;
      }
      BeamBacktrack bt;
      bt.scores = beam_scores_[cur]; bt.hist_tok = beam_hist_tok_; bt.hist_par = beam_hist_par_;
      bt.labels = labels + (int64_t)b0 * Ld; bt.ldl = Ld; bt.score_out = score + b0; bt.Bc = Bc; bt.K = K; bt.L = Ld;
      beam_backtrack(ctx_, bt);
    }
  } catch (...) {
    b_ = B; dual_rows_ = 0; ctx_row0_ = 0; persist_on_ = persist_save;
    throw;
  }
  b_ = B; dual_rows_ = 0; ctx_row0_ = 0; persist_on_ = persist_save;
  // back to the weight planes the executor's programs were recorded against (gate-interleaved when fused): a program
  // recorded against the plain planes would go stale at the next update, which only rebuilds the kind in use
  build_decoder_packs();
}

// ---- loadDictionary (src/utils/utils.lua:177-218) as a flat child table: node 0 = trie[2] (the start symbol), entry
// [node][vocab_id] = child node or -1; every word ends in an EOS (3) child; allow_digit_prefix loops the root to itself
// on EOS and on the ten digits.
static void trie_build(const std::vector<std::string>& words, bool allow_digit_prefix, int V, std::vector<int32_t>& table) {
  table.assign((size_t)(V + 1), -1);
  auto new_node = [&]() { table.resize(table.size() + (size_t)(V + 1), -1); return (int32_t)(table.size() / (V + 1) - 1); };
  for (const std::string& raw : words) {
    size_t a = 0, b = raw.size();
    while (a < b && isspace((unsigned char)raw[a])) a++;
    while (b > a && isspace((unsigned char)raw[b - 1])) b--;
    int32_t node = 0;
    if (allow_digit_prefix) {                                   // utils.lua:195-201
      table[3] = 0;
      for (int l = 48; l <= 57; l++) table[(size_t)(l - 48 + 3 + 1)] = 0;
    }
    for (size_t i = a; i < b; i++) {                            // utils.lua:202-214
      const int l = (unsigned char)raw[i];
      const int vid = l > 96 ? l - 97 + 13 + 1 : l - 48 + 3 + 1;
      if (vid < 1 || vid > V) throw InvalidError("dictionary word with a character outside [0-9a-z]: " + raw);
      int32_t child = table[(size_t)node * (V + 1) + vid];
      if (child < 0) {
        child = new_node();
        table[(size_t)node * (V + 1) + vid] = child;
      }
      node = child;
    }
    if (table[(size_t)node * (V + 1) + 3] < 0) {                // utils.lua:215-217
      const int32_t leaf = new_node();
      table[(size_t)node * (V + 1) + 3] = leaf;
    }
  }
}

}  // namespace aocr

extern "C" {
int aocr_trie_from_words(const char* words, int allow_digit_prefix, int32_t** table, int32_t* num_nodes) {
  if (!words || !table || !num_nodes) return AOCR_ERR_INVALID;
  try {
    std::vector<std::string> w;
    const char* p = words;
    while (*p) {
      const char* e = strchr(p, '\n');
      if (!e) e = p + strlen(p);
      w.emplace_back(p, e - p);
      p = *e ? e + 1 : e;
    }
    std::vector<int32_t> t;
    aocr::trie_build(w, allow_digit_prefix != 0, 39, t);
    *table = (int32_t*)malloc(t.size() * sizeof(int32_t));
    if (!*table) return AOCR_ERR_STATE;
    memcpy(*table, t.data(), t.size() * sizeof(int32_t));
    *num_nodes = (int32_t)(t.size() / 40);
    return AOCR_OK;
  } catch (...) {
    return AOCR_ERR_INVALID;
  }
}
int aocr_trie_load(const char* path, int allow_digit_prefix, int32_t** table, int32_t* num_nodes) {
  if (!path || !table || !num_nodes) return AOCR_ERR_INVALID;
  FILE* f = fopen(path, "r");
  if (!f) return AOCR_ERR_INVALID;                               // utils.lua:179-182
  std::string all;
  char buf[65536];
  size_t n;
  while ((n = fread(buf, 1, sizeof(buf), f)) > 0) all.append(buf, n);
  fclose(f);
  return aocr_trie_from_words(all.c_str(), allow_digit_prefix, table, num_nodes);
}
void aocr_trie_free(int32_t* table) { free(table); }
}
