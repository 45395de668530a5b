// engine_dec.cu — decoder schedule, criterion, optimiser and the greedy path of the step scheduler.
// Reference: src/model/model.lua:360-404,446-459,516-536 (greedy), :539-568 (train forward),
// :589-627 (gold pass), :643-661 (decoder backward), src/optim/optim_sgd.lua:40-95.
#include "engine.h"

#include <math.h>
#include <string.h>

namespace aocr {

// model.lua:539-552 / 360-372 / 589-602.  Quirk Q14: with -input_feed the reference's
// "zero layers >= 2" loop indexes the state list without the input-feed offset and so zeroes h1(0); the
// decoder starts from c1(0) = [c_fw(S); c_bw(1)], h1(0) = 0.  Without input feed h1(0) = [h_fw(S); h_bw(1)].
void Engine::decoder_init(int reps) {
  const int B = b_, S = S_;
  const int64_t slot = (int64_t)B * He;
  const float* cfw = Cenc + ((int64_t)0 * (S + 1) + S) * slot;
  const float* cbw = Cenc + ((int64_t)1 * (S + 1) + 0) * slot;
  const int64_t BR = (int64_t)B * reps;            // reps = 2: the dual decode pass, rows [B, 2B) start from the same state
  fill_zero(ctx_, X1, (size_t)BR * K1 * sizeof(float));
  for (int r = 0; r < reps; r++) {
    concat_enc_finals(ctx_, cfw, cbw, C1 + (int64_t)r * B * Hd, Hd, B, He);
    if (!cfg.input_feed) {
      const float* hfw = Henc + ((int64_t)0 * (S + 1) + S) * slot;
      const float* hbw = Henc + ((int64_t)1 * (S + 1) + 0) * slot;
      concat_enc_finals(ctx_, hfw, hbw, X1 + (int64_t)r * B * K1, K1, B, He);
    }
  }
  fill_zero(ctx_, C2, (size_t)BR * Hd * sizeof(float));
  fill_zero(ctx_, X2, (size_t)BR * 2 * Hd * sizeof(float));
  if (cfg.gemm_mode != 2) {   // the same initial state as bf16 operand planes for the first step's GEMMs
    Pack x1_0 = X1p; x1_0.rows = BR;
    split_to_pack(ctx_, X1, BR, K1, K1, 1, x1_0);
    fill_zero(ctx_, X2p.hi, (size_t)BR * 2 * Hd * sizeof(__nv_bfloat16));
    fill_zero(ctx_, X2p.lo, (size_t)BR * 2 * Hd * sizeof(__nv_bfloat16));
  }
}

// One decoder step (src/model/LSTM.lua:18-162).  Saved-state layout, per step t:
//   X1[t] = [a_{t-1} | h1_{t-1}]   X2[t] = [h1_t | h2_{t-1}]   CAT[t] = [cv_t | h2_t]   A_all[t] = a_t
void Engine::decoder_step(int t, const int32_t* tokens) {
  if (cfg.gemm_mode != 2) decoder_step_tc(t, tokens); else decoder_step_simt(t, tokens);
}

void Engine::decoder_step_simt(int t, const int32_t* tokens) {
  const int B = b_, S = S_;
  const int in1 = E + (cfg.input_feed ? Hd : 0);
  const int nsteps = dec_steps_;
  float* x1 = X1 + (int64_t)t * B * K1;
  float* x2 = X2 + (int64_t)t * B * 2 * Hd;
  float* cat = CAT + (int64_t)t * B * 2 * Hd;
  const bool has_next = (t + 1 < nsteps);
  Gemm g;
  // ---- layer 1 gates: a_{t-1} W_i1[:,E:]^T + h1_{t-1} W_h1^T   (embedding part + biases come from Ptab)
  if (cfg.input_feed) {
    g = Gemm();
    g.M = B; g.N = 4 * Hd; g.K = Hd;
    g.A = x1; g.sam = K1; g.sak = 1;
    g.B = d_params + L.l1_wi + E; g.sbk = 1; g.sbn = in1;
    g.C = Gs; g.ldc = 4 * Hd;
    gemm(g);
  }
  g = Gemm();
  g.M = B; g.N = 4 * Hd; g.K = Hd;
  g.A = x1 + h1off; g.sam = K1; g.sak = 1;
  g.B = d_params + L.l1_wh; g.sbk = 1; g.sbn = Hd;
  g.C = Gs; g.ldc = 4 * Hd; g.accumulate = cfg.input_feed ? 1 : 0;
  gemm(g);
  DecCell c1;
  c1.G = Gs; c1.addrows = Ptab; c1.rowsel = tokens; c1.addld = 4 * Hd;
  c1.c_prev = C1 + (int64_t)t * B * Hd; c1.c_new = C1 + (int64_t)(t + 1) * B * Hd;
  c1.acts = ACT1 + (int64_t)t * B * 4 * Hd;
  c1.h_out0 = x2; c1.ld0 = 2 * Hd;
  c1.h_out1 = has_next ? x1 + (int64_t)B * K1 + h1off : nullptr; c1.ld1 = K1;
  c1.B = B; c1.H = Hd;
  dec_cell_fwd(ctx_, c1);
  // ---- layer 2 gates: h1_t W_i2^T + h2_{t-1} W_h2^T + (b_i2 + b_h2)
  g = Gemm();
  g.M = B; g.N = 4 * Hd; g.K = Hd;
  g.A = x2; g.sam = 2 * Hd; g.sak = 1;
  g.B = d_params + L.l2_wi; g.sbk = 1; g.sbn = Hd;
  g.C = Gs; g.ldc = 4 * Hd;
  gemm(g);
  g.A = x2 + Hd; g.B = d_params + L.l2_wh; g.accumulate = 1;
  gemm(g);
  DecCell c2;
  c2.G = Gs; c2.addrows = bsum2; c2.rowsel = nullptr; c2.addld = 0;
  c2.c_prev = C2 + (int64_t)t * B * Hd; c2.c_new = C2 + (int64_t)(t + 1) * B * Hd;
  c2.acts = ACT2 + (int64_t)t * B * 4 * Hd;
  c2.h_out0 = cat + Hd; c2.ld0 = 2 * Hd;
  c2.h_out1 = has_next ? x2 + (int64_t)B * 2 * Hd + Hd : nullptr; c2.ld1 = 2 * Hd;
  c2.B = B; c2.H = Hd;
  dec_cell_fwd(ctx_, c2);
  // ---- attention: q = W_a h2 ; scores, softmax, context vector (fused kernel)
  float* q = Q + (int64_t)t * B * Hd;
  g = Gemm();
  g.M = B; g.N = Hd; g.K = Hd;
  g.A = cat + Hd; g.sam = 2 * Hd; g.sak = 1;
  g.B = d_params + L.wa; g.sbk = 1; g.sbn = Hd;
  g.C = q; g.ldc = Hd;
  gemm(g);
  prof_begin(1);
  attn_fwd(ctx_, ctx, q, ALPHA + (int64_t)t * B * S, cat, 2 * Hd, B, S, Hd);
  prof_end(1, (double)B * S * Hd * 4 + (double)B * (2.0 * Hd + S) * 4);
  // ---- a_t = tanh(W_c [cv ; h2])
  float* a = A_all + (int64_t)t * B * Hd;
  g = Gemm();
  g.M = B; g.N = Hd; g.K = 2 * Hd;
  g.A = cat; g.sam = 2 * Hd; g.sak = 1;
  g.B = d_params + L.wc; g.sbk = 1; g.sbn = 2 * Hd;
  g.C = a; g.ldc = Hd; g.act = ACT_TANH;
  gemm(g);
  if (cfg.input_feed && has_next) copy_strided(ctx_, x1 + (int64_t)B * K1, K1, a, Hd, B, Hd);
}

// per-timestep part of the decoder backward, fp32 SIMT flavour (gemm_mode 2: on-device cross-check path)
void Engine::decoder_backward_steps_simt() {
  const int B = b_, S = S_, T = T_;
  const int in1 = E + (cfg.input_feed ? Hd : 0);
  Gemm g;
  fill_zero(ctx_, dc1, (size_t)B * Hd * sizeof(float));
  fill_zero(ctx_, dc2, (size_t)B * Hd * sizeof(float));
  for (int t = T - 1; t >= 0; t--) {
    const bool last = (t == T - 1);
    float* du = dU + (int64_t)t * B * Hd;
    float* dcat = dCAT + (int64_t)t * B * 2 * Hd;
    float* dq = dQ + (int64_t)t * B * Hd;
    float* dg2 = dG2 + (int64_t)t * B * 4 * Hd;
    float* dg1 = dG1 + (int64_t)t * B * 4 * Hd;
    du_from_da(ctx_, (!last && cfg.input_feed) ? dX1 : nullptr, K1, dAgen + (int64_t)t * B * Hd,
               A_all + (int64_t)t * B * Hd, du, (int64_t)B * Hd, Hd);
    g = Gemm();                                           // d[cv;h2] = du W_c
    g.M = B; g.N = 2 * Hd; g.K = Hd;
    g.A = du; g.sam = Hd; g.sak = 1;
    g.B = d_params + L.wc; g.sbk = 2 * Hd; g.sbn = 1;
    g.C = dcat; g.ldc = 2 * Hd;
    gemm(g);
    prof_begin(1);
    attn_bwd(ctx_, ctx, ALPHA + (int64_t)t * B * S, dcat, 2 * Hd, DE + (int64_t)t * B * S, dq, B, S, Hd);
    prof_end(1, 2.0 * B * S * Hd * 4);
    g = Gemm();                                           // dh2 += dq W_a
    g.M = B; g.N = Hd; g.K = Hd;
    g.A = dq; g.sam = Hd; g.sak = 1;
    g.B = d_params + L.wa; g.sbk = Hd; g.sbn = 1;
    g.C = dH2q; g.ldc = Hd;
    gemm(g);
    DecCellBwd b2;
    b2.dh_a = dcat + Hd; b2.lda = 2 * Hd;
    b2.dh_b = dH2q; b2.ldb = Hd;
    b2.dh_c = last ? nullptr : dX2 + Hd; b2.ldc = 2 * Hd;
    b2.dc = dc2; b2.c_prev = C2 + (int64_t)t * B * Hd; b2.c_new = C2 + (int64_t)(t + 1) * B * Hd;
    b2.acts = ACT2 + (int64_t)t * B * 4 * Hd; b2.dG = dg2; b2.B = B; b2.H = Hd;
    dec_cell_bwd(ctx_, b2);
    g = Gemm();                                           // [dh1 | dh2_prev] = dg2 [W_i2 | W_h2]
    g.M = B; g.N = Hd; g.K = 4 * Hd;
    g.A = dg2; g.sam = 4 * Hd; g.sak = 1;
    g.B = d_params + L.l2_wi; g.sbk = Hd; g.sbn = 1;
    g.C = dX2; g.ldc = 2 * Hd;
    gemm(g);
    g.B = d_params + L.l2_wh; g.C = dX2 + Hd;
    gemm(g);
    DecCellBwd b1;
    b1.dh_a = dX2; b1.lda = 2 * Hd;
    b1.dh_b = last ? nullptr : dX1 + h1off; b1.ldb = K1;
    b1.dh_c = nullptr; b1.ldc = 0;
    b1.dc = dc1; b1.c_prev = C1 + (int64_t)t * B * Hd; b1.c_new = C1 + (int64_t)(t + 1) * B * Hd;
    b1.acts = ACT1 + (int64_t)t * B * 4 * Hd; b1.dG = dg1; b1.B = B; b1.H = Hd;
    dec_cell_bwd(ctx_, b1);
    if (cfg.input_feed) {                                 // da_prev = dg1 W_i1[:, E:]
      g = Gemm();
      g.M = B; g.N = Hd; g.K = 4 * Hd;
      g.A = dg1; g.sam = 4 * Hd; g.sak = 1;
      g.B = d_params + L.l1_wi + E; g.sbk = in1; g.sbn = 1;
      g.C = dX1; g.ldc = K1;
      gemm(g);
    }
    g = Gemm();                                           // dh1_prev = dg1 W_h1
    g.M = B; g.N = Hd; g.K = 4 * Hd;
    g.A = dg1; g.sam = 4 * Hd; g.sak = 1;
    g.B = d_params + L.l1_wh; g.sbk = Hd; g.sbn = 1;
    g.C = dX1 + h1off; g.ldc = K1;
    gemm(g);
  }
  // the carries the encoder backward starts from are now in place: d c1(0) in dc1, d h1(0) in dX1[:, h1off:]
}

// decoder backward through time (model.lua:643-661) with every parameter gradient time-batched afterwards
void Engine::decoder_backward() {
  const int B = b_, S = S_, T = T_;
  const int in1 = E + (cfg.input_feed ? Hd : 0);
  const int64_t R = (int64_t)T * B;
  const float inv_bn = 1.0f / (float)global_b();    // model.lua:645-647 (Q7: the GLOBAL batch under data parallelism)
  // generator + criterion for all steps at once (a_t are all known): model.lua:644-648
  generator_fwd(ctx_, A_all, d_params + L.wo, d_params + L.bo, tev_tb, logp[0], dZ, rowloss, R, Hd, V, inv_bn);
  reduce_sum_double(ctx_, rowloss, R, d_loss);
  Gemm g;
  g.M = (int)R; g.N = Hd; g.K = V;                       // dA_gen = dZ W_o
  g.A = dZ; g.sam = V; g.sak = 1;
  g.B = d_params + L.wo; g.sbk = Hd; g.sbn = 1;
  g.C = dAgen; g.ldc = Hd;
  gemm(g);

  if (cfg.gemm_mode != 2) {
    if (persist_on_) {
      fill_zero(ctx_, dc1, (size_t)B * Hd * sizeof(float));
      fill_zero(ctx_, dc2, (size_t)B * Hd * sizeof(float));
      run_program(PK_DEC_BWD, T, 0);
    } else {
      decoder_backward_steps_tc();
    }
  } else {
    decoder_backward_steps_simt();
  }
  const bool tc = cfg.gemm_mode != 2;
  if (tc) {   // dctxwc[b,s,:] = sum_t alpha_t[b,s] du_t[b,:]: feeds both D_ctx (lane 0) and dW_c[:, :H] (lane 1)
    g = Gemm();
    g.M = S; g.N = Hd; g.K = T; g.batch = B;
    g.A = ALPHA; g.sam = 1; g.sak = (int64_t)B * S; g.bsa = S;
    g.B = dU; g.sbk = (int64_t)B * Hd; g.sbn = 1; g.bsb = Hd;
    g.C = dCtxWc; g.ldc = Hd; g.bsc = (int64_t)S * Hd;
    gemm(g);
  }
  // the time-batched parameter gradients below depend only on the saved per-step tensors: lane 1, concurrently
  // with D_ctx + the encoder/CNN backward that continue on lane 0
  fork_to(1);
  use_lane(1);
  // ---- time-batched parameter gradients (weights are tied across t: clone_many_times, model_utils.lua:3-50)
  g = Gemm();
  g.M = V; g.N = Hd; g.K = (int)R;                       // dW_o = dZ^T A   (feeds nothing in the chain: lane 1)
  g.A = dZ; g.sam = 1; g.sak = V;
  g.B = A_all; g.sbk = Hd; g.sbn = 1;
  g.C = d_grads + L.wo; g.ldc = Hd;
  gemm(g);
  col_sum(ctx_, dZ, R, V, d_grads + L.bo, partial, 0);
  // operand planes of a column range of a saved per-timestep pack (rows = time*batch, as the fp32 tensors)
  auto cols_of = [&](const Pack& p, int64_t col0) {
    Pack v = p;
    v.hi += col0; v.lo += col0;
    return v;
  };
  auto wgrad = [&](const float* dY, int M, const float* X, int64_t ldx, int N, float* dW, int64_t ldw,
                   const Pack* pY = nullptr, const Pack* pX = nullptr) {
    Gemm w;
    w.M = M; w.N = N; w.K = (int)R;
    w.A = dY; w.sam = 1; w.sak = M;
    w.B = X; w.sbk = ldx; w.sbn = 1;
    w.C = dW; w.ldc = ldw;
    if (tc) { w.pa = pY; w.pb = pX; }     // written by the recurrence bodies: no conversion pass
    gemm(w);
  };
  const Pack pdU = tc ? cols_of(dUQp, 0) : Pack(), pdQ = tc ? cols_of(dUQp, Hd) : Pack();
  const Pack pX2a = tc ? cols_of(X2p, 0) : Pack(), pX2b = tc ? cols_of(X2p, Hd) : Pack();
  const Pack pX1a = tc ? cols_of(X1p, 0) : Pack(), pX1b = tc ? cols_of(X1p, h1off) : Pack();
  if (!tc) {
    wgrad(dU, Hd, CAT, 2 * Hd, 2 * Hd, d_grads + L.wc, 2 * Hd);
  } else {
    // the tensor-core path never forms cv_t (engine_dec_tc.cu): dW_c[:, H:] = dU^T H2, and dW_c[:, :H] from dctxwc
    // (dW_c1 = sum_t du_t^T cv_t re-associated over source positions)
    wgrad(dU, Hd, CAT + Hd, 2 * Hd, Hd, d_grads + L.wc + Hd, 2 * Hd, &pdU, &H2p);
    g = Gemm();                                           // dW_c[:, :H] = dctxwc^T ctx   (K = B*S)
    g.M = Hd; g.N = Hd; g.K = B * S;
    g.A = dCtxWc; g.sam = 1; g.sak = Hd;
    g.B = ctx; g.sbk = Hd; g.sbn = 1;
    g.C = d_grads + L.wc; g.ldc = 2 * Hd;
    gemm(g);
  }
  wgrad(dQ, Hd, CAT + Hd, 2 * Hd, Hd, d_grads + L.wa, Hd, tc ? &pdQ : nullptr, tc ? &H2p : nullptr);
  wgrad(dG2, 4 * Hd, X2, 2 * Hd, Hd, d_grads + L.l2_wi, Hd, tc ? &dG2p : nullptr, tc ? &pX2a : nullptr);
  wgrad(dG2, 4 * Hd, X2 + Hd, 2 * Hd, Hd, d_grads + L.l2_wh, Hd, tc ? &dG2p : nullptr, tc ? &pX2b : nullptr);
  col_sum(ctx_, dG2, R, 4 * Hd, d_grads + L.l2_bi, partial, 0);
  AOCR_CUDA(cudaMemcpyAsync(d_grads + L.l2_bh, d_grads + L.l2_bi, (size_t)4 * Hd * sizeof(float), cudaMemcpyDeviceToDevice,
                            ctx_.st));
  if (cfg.input_feed) wgrad(dG1, 4 * Hd, X1, K1, Hd, d_grads + L.l1_wi + E, in1, tc ? &dG1p : nullptr, tc ? &pX1a : nullptr);
  wgrad(dG1, 4 * Hd, X1 + h1off, K1, Hd, d_grads + L.l1_wh, Hd, tc ? &dG1p : nullptr, tc ? &pX1b : nullptr);
  col_sum(ctx_, dG1, R, 4 * Hd, d_grads + L.l1_bi, partial, 0);
  AOCR_CUDA(cudaMemcpyAsync(d_grads + L.l1_bh, d_grads + L.l1_bi, (size_t)4 * Hd * sizeof(float), cudaMemcpyDeviceToDevice,
                            ctx_.st));
  // embedding path: dP[v] = sum of dg1 rows whose input token is v (LookupTable has no paddingValue: PAD rows count)
  if (cfg.gemm_mode != 2) {
    onehot(ctx_, tgt_tb, dZ /* reused as (R x V) one-hot scratch: dZ is dead by now */, R, V);
    g = Gemm();                                           // dP = onehot^T dG1   (MN-major tensor-core GEMM)
    g.M = V; g.N = 4 * Hd; g.K = (int)R;
    g.A = dZ; g.sam = 1; g.sak = V;
    g.B = dG1; g.sbk = 4 * Hd; g.sbn = 1;
    g.C = dP; g.ldc = 4 * Hd;
    g.pb = &dG1p;
    gemm(g);
  } else {
    token_segment_sum(ctx_, dG1, tgt_tb, dP, R, 4 * Hd, V);
  }
  // dEmb = dP W_i1[:, :E]
  thin_n_gemm(ctx_, dP, 4 * Hd, d_params + L.l1_wi, in1, d_grads + L.emb, E, V, 4 * Hd, E);
  g = Gemm();                                             // dW_i1[:, :E] = dP^T Emb
  g.M = 4 * Hd; g.N = E; g.K = V;
  g.A = dP; g.sam = 1; g.sak = 4 * Hd;
  g.B = d_params + L.emb; g.sbk = E; g.sbn = 1;
  g.C = d_grads + L.l1_wi; g.ldc = in1;
  gemm(g);
  use_lane(0);
  // D_ctx[b] = sum_t alpha_t[b]^T dcv_t[b] + de_t[b]^T q_t[b]   (replaces the per-step RMW of model.lua:652-653)
  g = Gemm();
  g.M = S; g.N = Hd; g.K = T; g.batch = B;
  g.A = DE; g.sam = 1; g.sak = (int64_t)B * S; g.bsa = S;
  g.B = Q; g.sbk = (int64_t)B * Hd; g.sbn = 1; g.bsb = Hd;
  g.C = Dctx; g.ldc = Hd; g.bsc = (int64_t)S * Hd;
  gemm(g);
  if (!tc) {
    g.A = ALPHA;
    g.B = dCAT; g.sbk = (int64_t)B * 2 * Hd; g.bsb = 2 * Hd;
    g.accumulate = 1;
    gemm(g);
  } else {
    // sum_t alpha^T dcv_t = (sum_t alpha^T du_t) W_c[:, :H] = dctxwc W_c1
    g = Gemm();
    g.M = B * S; g.N = Hd; g.K = Hd;
    g.A = dCtxWc; g.sam = Hd; g.sak = 1;
    g.B = d_params + L.wc; g.sbk = 2 * Hd; g.sbn = 1;
    g.C = Dctx; g.ldc = Hd; g.accumulate = 1;
    gemm(g);
  }
  taps_["dctx"] = {Dctx, (int64_t)B * S * Hd};
}

// ---- data-parallel hooks (no reference counterpart) ------------------------------------------
static void stat_sync_tramp(void* user, float* buf, int64_t n) {
  static_cast<Engine*>(user)->exchange(buf, n, 0);
}
// one exchange point: the native NCCL path when aocr_dp_init was called, else the host hook
void Engine::exchange(float* buf, int64_t n, int kind) {
  if (nccl_comm_) { dp_allreduce(buf, n, kind); return; }
  AOCR_CHECK(ar_fn != nullptr, "dp_world > 1 needs aocr_dp_init or aocr_set_allreduce before a training step");
  ar_fn(ar_user, buf, n, kind);
}
StatSync Engine::stat_sync() {
  StatSync s;
  if (cfg.dp_world > 1) {
    s.fn = stat_sync_tramp; s.user = this; s.world = cfg.dp_world;
    s.grows = (double)global_b() / (double)b_;     // all ranks share H x W: global rows = rows * global batch / local batch
  }
  return s;
}
// groups are laid out [proj | decoder | enc_fw | enc_bw | cnn] = the order in which backward completes them
void Engine::grad_bucket(int first_group, int last_group) {
  if (cfg.dp_world <= 1) return;
  grad_range(L.goff[first_group], L.goff[last_group] + L.gphys[last_group]);
}
void Engine::grad_range(int64_t off, int64_t end) {
  if (cfg.dp_world > 1 && end > off) exchange(d_grads + off, end - off, 1);
}
// a short gradient range at the very end of backward: through the peer mailboxes when they exist (one small kernel on the
// current stream, ~10 us) instead of an NCCL all-reduce queued behind the earlier buckets
void Engine::grad_small(int64_t off, int64_t end) {
  if (cfg.dp_world <= 1 || end <= off) return;
  if (!dp_peer_allreduce(d_grads + off, end - off)) exchange(d_grads + off, end - off, 1);
}
void Engine::grad_join() {
  if (cfg.dp_world > 1) exchange(nullptr, 0, 2);
}

// feval, train branch (model.lua:284-316,537-569,634-695)
void Engine::forward_backward_enqueue() {
  AOCR_CHECK(have_batch_, "no batch staged");
  AOCR_CHECK(params_set_, "the model has no parameters yet: call aocr_init_params or aocr_set_params first");
  // data parallelism: the loss scale and the batch-norm row count need the step's GLOBAL batch; a SUM all-reduce with
  // the local 1/b would silently yield gradients `world` times too large
  AOCR_CHECK(cfg.dp_world <= 1 || cfg.global_batch >= b_,
             "dp_world > 1 needs global_batch (aocr_config.global_batch or aocr_set_global_batch) >= the local batch");
  AOCR_CUDA(cudaSetDevice(device_));
  const int B = b_, T = T_;
  const bool prepped = prep_weights();
  gather_tokens(ctx_, tgt_bt, tgt_tb, B, T, T, 1);
  gather_tokens(ctx_, tev_bt, tev_tb, B, T, T, 1);
  phase_mark("start");
  fork_to(1);                      // model.lua:637-639: gradients are not touched before the decoder backward
  use_lane(1);
  fill_zero(ctx_, d_grads, (size_t)L.total * sizeof(float));
  use_lane(0);
  cnn_forward(true);
  (void)prepped;
  join_from(1);                    // prep_weights and the gradient clear (lane 1) are complete
  phase_mark("cnn_fwd");
  encoder_forward();
  attention_precompute();
  phase_mark("enc_fwd");
  decoder_init();
  dec_steps_ = T;
  if (persist_on_ && cfg.gemm_mode != 2) run_program(PK_DEC_FWD, T, 0);
  else for (int t = 0; t < T; t++) decoder_step(t, tgt_tb + (int64_t)t * B);
  phase_mark("dec_fwd");
  taps_["a_all"] = {A_all, (int64_t)T * B * Hd};
  taps_["alpha"] = {ALPHA, (int64_t)T * B * S_};
  decoder_backward();
  phase_mark("dec_bwd");
  encoder_backward();
  phase_mark("enc_bwd");
  taps_["dsrc"] = {dsrc, (int64_t)S_ * B * 512};
  // Gradient buckets in the order backward completes them (the flat buffer is laid out in that order).  Native
  // exchange: a bucket is issued from the lane that finishes it (the communication stream waits on that lane only),
  // so lane 0 never waits for the time-batched weight gradients of lane 1; the CNN bucket is split so that only the
  // small conv1-4 tail is exposed after the last kernel.  Hook flavour: the host runtime orders on lane 0's stream.
  const bool dp = cfg.dp_world > 1, native = dp && dp_native();
  // [proj | decoder] are adjacent in the flat buffer but not as group numbers (decoder 3, projector 4): two masks
  if (dp) {
    if (native) {
      use_lane(1); grad_bucket(G_PROJ, G_DEC); use_lane(0);
      early_update(G_DEC, G_PROJ, comm_st_);
    } else { join_from(1); grad_bucket(G_PROJ, G_DEC); }
  } else {
    early_update(G_DEC, G_PROJ, lanes_[1].st);
  }
  cnn_bucket_split_ = native ? L.conv_w[4] : -1;      // conv5..conv7 (+BN) complete after the conv5 iteration
  if (native) {                                       // all encoder gradients are lane-1 work
    use_lane(1); grad_bucket(G_ENC_FW, G_ENC_BW); use_lane(0);
    early_update(G_ENC_FW, G_ENC_BW, comm_st_);
  } else if (!dp) {
    early_update(G_ENC_FW, G_ENC_BW, lanes_[1].st);
  }
  cnn_backward();
  phase_mark("cnn_bwd");
  join_from(1);                    // encoder (and decoder) weight gradients
  phase_mark("lane_join");
  if (dp && !native) { grad_bucket(G_ENC_FW, G_ENC_BW); grad_bucket(G_CNN, G_CNN); }
  grad_join();
  phase_report();
  have_grads_ = true;
  last_logp_rows_[0] = T * B;
}

double Engine::read_loss() {
  double v = 0.0;
  AOCR_CUDA(cudaMemcpyAsync(&v, d_loss, sizeof(double), cudaMemcpyDeviceToHost, ctx_.st));
  AOCR_CUDA(cudaStreamSynchronize(ctx_.st));
  return v;
}

static int sq_blocks(int64_t n) { return (int)(n / 4096 + 1 < 1024 ? n / 4096 + 1 : 1024); }

void Engine::group_norms(double* pn, double* gn) {
  AOCR_CHECK(have_grads_, "no gradients yet: call aocr_forward_backward first");
  for (int g = 0; g < 5; g++) {
    int nb = sq_blocks(L.gphys[g]);
    sumsq_partial(ctx_, d_grads + L.goff[g], L.gphys[g], d_sq_partial + g * 1024, nb);
    sumsq_final(ctx_, d_sq_partial + g * 1024, nb, d_sumsq + g);
  }
  for (int g = 0; g < 5; g++) {
    int nb = sq_blocks(L.gphys[g]);
    sumsq_partial(ctx_, d_params + L.goff[g], L.gphys[g], d_sq_partial + g * 1024, nb);
    sumsq_final(ctx_, d_sq_partial + g * 1024, nb, d_sumsq + 5 + g);
  }
  double h[10];
  AOCR_CUDA(cudaMemcpyAsync(h, d_sumsq, sizeof(h), cudaMemcpyDeviceToHost, ctx_.st));
  AOCR_CUDA(cudaStreamSynchronize(ctx_.st));
  for (int g = 0; g < 5; g++) { gn[g] = sqrt(h[g]); pn[g] = sqrt(h[5 + g]); }
}

// optim.sgd_list, default branch (optim_sgd.lua:49-52,90): per group clip to `clip`, p -= lr*g.  No host sync.
void Engine::set_lr_clip(double lr, double clip) {
  // pageable source: the runtime stages the 16 bytes before cudaMemcpyAsync returns, so back-to-back enqueues with
  // different learning rates cannot race on a host buffer.  Never captured: the step size is data of a replayed graph.
  const double v[2] = {lr, clip};
  AOCR_CUDA(cudaMemcpyAsync(d_lrclip, v, sizeof(v), cudaMemcpyHostToDevice, ctx_.st));
}
void Engine::sgd_enqueue(double lr, double clip) {
  AOCR_CHECK(have_grads_, "no gradients yet: call aocr_forward_backward first");
  set_lr_clip(lr, clip);
  sgd_enqueue_kernels();
}
void Engine::sgd_subset(Ctx& c, unsigned mask) {
  SgdGroups G;
  for (int g = 0; g < 5; g++) {
    G.off[g] = L.goff[g]; G.n[g] = L.gphys[g];
    int64_t nb = L.gphys[g] / 16384 + 1;          // ~16k floats per block and pass
    G.nb[g] = (mask >> g) & 1u ? (int)(nb < 1024 ? nb : 1024) : 0;      // 0 blocks: the group is not touched
  }
  sgd_groups(c, d_params, d_grads, G, d_sq_partial, d_sumsq, d_lrclip);
}
// Groups [g_first, g_last] are final once everything enqueued so far on stream `after` has run (the lane that computed
// their weight gradients, or the communication stream behind their all-reduce): clip + SGD of those groups on the update
// stream, concurrently with the rest of backward.  Nothing downstream reads their fp32 master parameters any more (the
// recurrences and their weight gradients run on operand planes built at the start of the step).
void Engine::early_update(int g_first, int g_last, cudaStream_t after) {
  if (!fused_update_ || !early_update_on_ || !lanes_on_ || !upd_st_) return;
  cudaEvent_t ev = upd_ev_[upd_ev_next_++ % 3];
  AOCR_CUDA(cudaEventRecord(ev, after));
  AOCR_CUDA(cudaStreamWaitEvent(upd_st_, ev, 0));
  unsigned mask = 0;
  for (int g = g_first; g <= g_last; g++) mask |= 1u << g;
  Ctx c = ctx_;
  c.st = upd_st_;
  sgd_subset(c, mask);
  ctx_.launches = c.launches;
  updated_mask_ |= mask;
  upd_pending_ = true;
}
void Engine::sgd_enqueue_kernels() {
  if (upd_pending_) {            // the early updates join the main stream
    AOCR_CUDA(cudaEventRecord(upd_ev_[3], upd_st_));
    AOCR_CUDA(cudaStreamWaitEvent(ctx_.st, upd_ev_[3], 0));
    upd_pending_ = false;
  }
  const unsigned rest = 0x1fu & ~updated_mask_;
  updated_mask_ = 0;
  if (rest) sgd_subset(ctx_, rest);
  mark_weights_dirty();
}

// ---- whole-step CUDA graphs ------------------------------------------------------------------------------
// A step is a fixed sequence of ~700 dependent launches whose per-launch latency (4-5 us on this part), not their
// work, bounds the step at batch 64.  The second time a (shape, lr) key is seen the sequence is stream-captured and
// instantiated; from then on one cudaGraphLaunch replays it.  Eager first (fills the weight-pack / tensor-map
// caches: no allocation may happen during capture).  Data-parallel steps are captured too when the exchange is the
// native NCCL one (engine_nccl.cu); with a host hook (aocr_set_allreduce) they stay eager.  Never while profiling.
void Engine::train_step_enqueue(double lr, double clip) {
  const bool eligible = graphs_on_ && (cfg.dp_world <= 1 || dp_native()) && !prof_on && !phases_on_;
  set_lr_clip(lr, clip);           // outside any capture: one graph per shape serves every learning rate
  struct FusedScope { bool& f; FusedScope(bool& x) : f(x) { f = true; } ~FusedScope() { f = false; } } fused_scope(fused_update_);
  if (!eligible) {
    forward_backward_enqueue();
    sgd_enqueue_kernels();
    return;
  }
  GraphKey key{0, b_, W_, T_, cfg.global_batch};
  if (graphs_.size() >= kMaxGraphs && graphs_.find(key) == graphs_.end()) drop_graphs();   // variable T / W: bounded
  GraphEntry& e = graphs_[key];
  if (e.exec == nullptr && e.seen >= 1) {
    mark_weights_dirty();          // the captured sequence must contain the weight-pack refresh
    cudaGraph_t graph = nullptr;
    AOCR_CUDA(cudaStreamBeginCapture(ctx_.st, cudaStreamCaptureModeThreadLocal));
    try {
      forward_backward_enqueue();
      sgd_enqueue_kernels();
    } catch (...) {
      cudaStreamEndCapture(ctx_.st, &graph);
      if (graph) cudaGraphDestroy(graph);
      throw;
    }
    AOCR_CUDA(cudaStreamEndCapture(ctx_.st, &graph));
    AOCR_CUDA(cudaGraphInstantiate(&e.exec, graph, 0));
    AOCR_CUDA(cudaGraphDestroy(graph));
  }
  e.seen++;
  if (e.exec) {
    AOCR_CUDA(cudaGraphLaunch(e.exec, ctx_.st));
    ctx_.launches += graph_launches_train_;
    have_grads_ = true;
    mark_weights_dirty();
  } else {
    const int64_t l0 = ctx_.launches;
    forward_backward_enqueue();
    sgd_enqueue_kernels();
    graph_launches_train_ = ctx_.launches - l0;
  }
}

void Engine::drop_graphs() {
  AOCR_CUDA(cudaStreamSynchronize(ctx_.st));
  for (auto& kv : graphs_) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  graphs_.clear();
}

void Engine::decode_step_enqueue() {
  const bool eligible = graphs_on_ && !prof_on && !phases_on_;
  if (!eligible) { decode_enqueue(); return; }
  GraphKey key{1, b_, W_, T_, 0};
  if (graphs_.size() >= kMaxGraphs && graphs_.find(key) == graphs_.end()) drop_graphs();
  GraphEntry& e = graphs_[key];
  if (e.exec == nullptr && e.seen >= 1 && !weights_dirty_) {
    cudaGraph_t graph = nullptr;
    AOCR_CUDA(cudaStreamBeginCapture(ctx_.st, cudaStreamCaptureModeThreadLocal));
    try {
      decode_enqueue();
    } catch (...) {
      cudaStreamEndCapture(ctx_.st, &graph);
      if (graph) cudaGraphDestroy(graph);
      throw;
    }
    AOCR_CUDA(cudaStreamEndCapture(ctx_.st, &graph));
    AOCR_CUDA(cudaGraphInstantiate(&e.exec, graph, 0));
    AOCR_CUDA(cudaGraphDestroy(graph));
  }
  e.seen++;
  if (e.exec && !weights_dirty_) {
    AOCR_CUDA(cudaGraphLaunch(e.exec, ctx_.st));
    ctx_.launches += graph_launches_decode_;
    last_logp_rows_[1] = last_logp_rows_[2] = Tmax * b_;
  } else {
    const int64_t l0 = ctx_.launches;
    decode_enqueue();
    if (!weights_dirty_) graph_launches_decode_ = ctx_.launches - l0;
  }
}

// forward_only branch, beam 1, no trie (model.lua:360-404,446-459,516-536,570-627)
void Engine::decode_enqueue() {
  AOCR_CHECK(have_batch_, "no batch staged");
  AOCR_CHECK(params_set_, "the model has no parameters yet: call aocr_init_params or aocr_set_params first");
  AOCR_CUDA(cudaSetDevice(device_));
  const int B = b_, T = T_, Ld = Tmax;
  const bool prepped = prep_weights();
  gather_tokens(ctx_, tgt_bt, tgt_tb, B, T, Ld, 1);   // model.lua:266-274: pad to max_decoder_l with PAD
  gather_tokens(ctx_, tev_bt, tev_tb, B, T, Ld, 1);
  cnn_forward(false);
  if (prepped) join_from(1);       // prep_weights (lane 1), when the weights changed since the last call
  encoder_forward();
  attention_precompute();
  dec_steps_ = Ld;
  if (persist_on_ && cfg.gemm_mode != 2 && 2 * B <= 256 && dual_on_) {
    // Dual pass: the greedy pass and the teacher-forced gold pass are independent recurrences over the same encoder
    // state, and a decoder step is bound by streaming the weights, not by the batch: run them as ONE batch of 2B rows
    // (rows [0,B) greedy, rows [B,2B) gold) and halve the number of sequential steps (model.lua:360-404 + 589-627).
    const int B2 = 2 * B;
    AOCR_CUDA(cudaMemcpyAsync(tokseq, tgt_tb, (size_t)B * sizeof(int32_t), cudaMemcpyDeviceToDevice, ctx_.st));  // GO row
    AOCR_CUDA(cudaMemcpy2DAsync(tokseq + B, (size_t)B2 * sizeof(int32_t), tgt_tb, (size_t)B * sizeof(int32_t),
                                (size_t)B * sizeof(int32_t), (size_t)Ld, cudaMemcpyDeviceToDevice, ctx_.st));
    decoder_init(2);
    // The gold rows are needed for the batch's own target length only: past it the padded targets carry no loss and no
    // score (criterion.lua:5; model.lua:614-618), so nothing the step returns depends on those steps.  The dual pass
    // runs T steps, then the greedy rows go on alone - half the rows per command - on the same state layout.
    const int Tg = (short_gold_on_ && T + 4 <= Ld) ? T : Ld;
    if (Tg < Ld) {
      AOCR_CUDA(cudaMemsetAsync(rowloss + (int64_t)Tg * B, 0, (size_t)(Ld - Tg) * B * sizeof(float), ctx_.st));
      AOCR_CUDA(cudaMemsetAsync(logp[2] + (int64_t)Tg * B * V, 0, (size_t)(Ld - Tg) * B * V * sizeof(float), ctx_.st));
    }
    dual_rows_ = B;
    b_ = B2;
    try {
      run_program(PK_DEC_DUAL, Tg, Tg < Ld ? Ld : 0);
      if (Tg < Ld) {
        b_ = B; dual_rows_ = 0; slot_rows_ = B2;
        run_program(PK_DEC_DUAL_TAIL, Ld, Tg);
      }
    } catch (...) {
      b_ = B; dual_rows_ = 0; slot_rows_ = 0;
      throw;
    }
    b_ = B; dual_rows_ = 0; slot_rows_ = 0;
    reduce_sum_double(ctx_, rowloss, (int64_t)Ld * B, d_loss);
    last_logp_rows_[1] = last_logp_rows_[2] = Ld * B;
    return;
  }
  // greedy pass
  AOCR_CUDA(cudaMemcpyAsync(tok, tgt_tb, (size_t)B * sizeof(int32_t), cudaMemcpyDeviceToDevice, ctx_.st));  // GO row
  decoder_init();
  if (persist_on_ && cfg.gemm_mode != 2) {
    run_program(PK_DEC_GREEDY, Ld, 0);
  } else {
    for (int t = 0; t < Ld; t++) {
      decoder_step(t, tok);
      float* lp = logp[1] + (int64_t)t * B * V;
      generator_fwd(ctx_, A_all + (int64_t)t * B * Hd, d_params + L.wo, d_params + L.bo, nullptr, lp, nullptr, nullptr, B,
                    Hd, V, 1.0f);
      greedy_select(ctx_, lp, tok, score, labels, Ld, t, B, V);
    }
  }
  // gold pass: teacher forced with the padded targets (model.lua:589-627); only the batch's own target length carries
  // loss / score (see the dual branch above), so the pass stops there
  decoder_init();
  const int Tg = (short_gold_on_ && T + 4 <= Ld) ? T : Ld;
  if (Tg < Ld) {
    AOCR_CUDA(cudaMemsetAsync(rowloss + (int64_t)Tg * B, 0, (size_t)(Ld - Tg) * B * sizeof(float), ctx_.st));
    AOCR_CUDA(cudaMemsetAsync(logp[2] + (int64_t)Tg * B * V, 0, (size_t)(Ld - Tg) * B * V * sizeof(float), ctx_.st));
  }
  dec_steps_ = Tg;
  if (persist_on_ && cfg.gemm_mode != 2) run_program(PK_DEC_FWD, Tg, 0);
  else for (int t = 0; t < Tg; t++) decoder_step(t, tgt_tb + (int64_t)t * B);
  dec_steps_ = Ld;
  generator_fwd(ctx_, A_all, d_params + L.wo, d_params + L.bo, tev_tb, logp[2], nullptr, rowloss, (int64_t)Tg * B, Hd, V,
                1.0f);
  reduce_sum_double(ctx_, rowloss, (int64_t)Ld * B, d_loss);
  last_logp_rows_[1] = last_logp_rows_[2] = Ld * B;
}

void Engine::decode_collect(int32_t* out_labels, double* pred, double* gold, double* loss_sum, int32_t* num_correct) {
  const int B = b_, Ld = Tmax, T = T_;
  std::vector<int32_t> hl((size_t)B * Ld);
  std::vector<double> hs(B);
  std::vector<float> hr((size_t)Ld * B);
  double hloss = 0.0;
  AOCR_CUDA(cudaMemcpyAsync(hl.data(), labels, hl.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx_.st));
  AOCR_CUDA(cudaMemcpyAsync(hs.data(), score, hs.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx_.st));
  AOCR_CUDA(cudaMemcpyAsync(hr.data(), rowloss, hr.size() * sizeof(float), cudaMemcpyDeviceToHost, ctx_.st));
  AOCR_CUDA(cudaMemcpyAsync(&hloss, d_loss, sizeof(double), cudaMemcpyDeviceToHost, ctx_.st));
  AOCR_CUDA(cudaStreamSynchronize(ctx_.st));
  if (out_labels) memcpy(out_labels, hl.data(), hl.size() * sizeof(int32_t));
  if (pred) memcpy(pred, hs.data(), hs.size() * sizeof(double));
  if (gold) {
    for (int b = 0; b < B; b++) {
      double s = 0.0;
      for (int t = 0; t < Ld; t++) s -= (double)hr[(size_t)t * B + b];
      gold[b] = s;
    }
  }
  if (loss_sum) *loss_sum = hloss;
  if (num_correct) {
    // evalWordErrRate (utils.lua:136-175): compare id lists up to (not including) the first EOS (3)
    int nc = 0;
    for (int b = 0; b < B; b++) {
      int lp = 0, lt = 0;
      while (lp < Ld && hl[(size_t)b * Ld + lp] != 3) lp++;
      auto tv = [&](int t) { return t < T ? h_tev_[(size_t)b * T + t] : 1; };
      while (lt < Ld && tv(lt) != 3) lt++;
      bool same = (lp == lt);
      for (int t = 0; same && t < lp; t++) same = (hl[(size_t)b * Ld + t] == tv(t));
      nc += same ? 1 : 0;
    }
    *num_correct = nc;
  }
}

void Engine::get_logprobs(int which, float* out, int64_t n) {
  AOCR_CHECK(which >= 0 && which < 3, "which must be 0 (train), 1 (greedy) or 2 (gold)");
  AOCR_CHECK(last_logp_rows_[which] > 0, "no log-probs of that kind have been produced yet");
  AOCR_CHECK(n == (int64_t)last_logp_rows_[which] * V, "log-prob buffer length mismatch");
  AOCR_CUDA(cudaStreamSynchronize(ctx_.st));
  AOCR_CUDA(cudaMemcpy(out, logp[which], (size_t)n * sizeof(float), cudaMemcpyDeviceToHost));
}

void Engine::debug_read(const char* name, float* out, int64_t n) {
  auto it = taps_.find(name);
  AOCR_CHECK(it != taps_.end(), std::string("unknown debug tap: ") + name);
  AOCR_CHECK(n == it->second.n, "debug tap length mismatch");
  AOCR_CUDA(cudaStreamSynchronize(ctx_.st));
  if (it->second.bytes) {
    std::vector<uint8_t> tmp((size_t)n);
    AOCR_CUDA(cudaMemcpy(tmp.data(), it->second.ptr, (size_t)n, cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < n; i++) out[i] = (float)tmp[i];
    return;
  }
  AOCR_CUDA(cudaMemcpy(out, it->second.ptr, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost));
}

}  // namespace aocr
