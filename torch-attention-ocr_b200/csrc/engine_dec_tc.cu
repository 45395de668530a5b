// engine_dec_tc.cu — the decoder schedule on tensor cores (gemm_mode 0/1).  Per timestep: three tcgen05 GEMMs
// (layer-1 gates, layer-2 gates, attention query + h2 half of the output projection), each followed by ONE
// memory-bound body that sums the GEMM's split-K partials, applies the cell / attention+output math and writes
// the next GEMM's operand as bf16 planes.  The context half of the output projection is time-independent per
// source position and is precomputed once per batch (attention_precompute).  Weights are concatenated along K ([W_i | W_h]) so a cell needs one GEMM,
// and the batch rides the UMMA N dimension (swap-AB) so the 128-row M tile is filled by weight rows.
// The same step functions either LAUNCH each piece as its own kernel or RECORD it into the command list of the
// persistent recurrence executor (persist.cu), which then runs the whole recurrence as one cooperative kernel.
// Reference schedule: src/model/model.lua:553-568 (forward), :643-661 (backward).
#include "engine.h"

namespace aocr {

namespace {
Pack sub_rows(const Pack& p, int64_t row0, int64_t rows) {
  Pack s;
  s.hi = p.hi + row0 * p.kp; s.lo = p.lo + row0 * p.kp; s.rows = rows; s.kp = p.kp;
  return s;
}
Pack sub_cols(const Pack& p, int64_t col0) {   // same pitch, shifted origin (for split_to_pack with kwrite)
  Pack s = p;
  s.hi = p.hi + col0; s.lo = p.lo + col0;
  return s;
}
PackOut pack_out(const Pack& p, int64_t row0, int64_t col0) {
  PackOut o;
  o.hi = p.hi + row0 * p.kp + col0; o.lo = p.lo + row0 * p.kp + col0; o.ld = p.kp;
  return o;
}
PartIn part_in(const TcOut& o, int64_t ld, int64_t col0 = 0) {
  PartIn a;
  a.p = o.base + col0; a.nz = o.nz; a.stride = o.stride; a.ld = ld;
  return a;
}
}  // namespace

// ---- emitters: launch now, or append to the program being recorded -----------------------------------------
// weights (M rows) x activation rows [row0, row0+b) of the pack X, K columns starting at k0 -> (b x M) partials in ws
TcOut Engine::emit_gemm(const Pack& W, int M, const Pack& X, int64_t row0, int64_t k0, int K, float* ws) {
  const int B = b_;
  const int terms = cfg.gemm_mode == 1 ? 1 : 3;
  if (rec_) {
    PGemmPlan pl = persist_plan_gemm(M, B, K, rec_max_ctas_, dec_ws_floats);
    PGemm g{};
    g.map_a = rec_->add_map_pair(tc_map_2d(W.hi, W.rows, W.kp, 128), tc_map_2d(W.lo, W.rows, W.kp, 128));
    g.map_b = rec_->add_map_pair(tc_map_2d(X.hi, X.rows, X.kp, rec_->bn), tc_map_2d(X.lo, X.rows, X.kp, rec_->bn));
    g.m_tiles = pl.m_tiles; g.splits = pl.splits; g.kb_per = pl.kb_per; g.num_kb = pl.num_kb;
    g.b_row0 = (int)row0; g.b_k0 = (int)k0; g.M = M; g.N = B; g.terms = terms;
    g.ws = ws; g.part_stride = pl.part_stride; g.ldc = M;
    rec_->add(P_GEMM, g);
    if (pl.m_tiles * pl.splits > rec_->grid) rec_->grid = pl.m_tiles * pl.splits;
    TcOut o;
    o.base = ws; o.nz = pl.splits; o.stride = pl.part_stride;
    return o;
  }
  Pack xs = sub_rows(X, row0, B);
  xs.hi += k0; xs.lo += k0;
  TcGemm g;
  g.A = W; g.B = xs; g.M = M; g.N = B; g.K = K; g.ldc = M; g.transpose_out = true;   // C[b][m]
  g.terms = terms; g.defer_reduce = true; g.ws = ws; g.ws_floats = dec_ws_floats;
  prof_begin(0);
  TcOut o = gemm_tc(ctx_, g);
  prof_end(0, 2.0 * M * (double)B * K);
  return o;
}
void Engine::emit(const CellFwdTc& p) { if (rec_) rec_->add(P_CELL_FWD, p); else cell_fwd_tc(ctx_, p); }
void Engine::emit(const CellBwdTc& p) { if (rec_) rec_->add(P_CELL_BWD, p); else cell_bwd_tc(ctx_, p); }
void Engine::emit(const EncCellFwdTc& p) { if (rec_) rec_->add(P_ENC_CELL_FWD, p); else enc_cell_fwd_tc(ctx_, p); }
void Engine::emit(const EncCellBwdTc& p) { if (rec_) rec_->add(P_ENC_CELL_BWD, p); else enc_cell_bwd_tc(ctx_, p); }
void Engine::emit(const AttnOutTc& p) {
  if (rec_) { rec_->add(P_ATTN_OUT, p); return; }
  prof_begin(1);
  attn_out_tc(ctx_, p);
  prof_end(1, 2.0 * p.B * p.S * p.H * 4 + (double)p.B * (3.0 * p.H + p.S) * 4);
}
void Engine::emit(const AttnDuTc& p) {
  if (rec_) { rec_->add(P_ATTN_DU, p); return; }
  prof_begin(1);
  attn_du_tc(ctx_, p);
  prof_end(1, 2.0 * p.B * p.S * p.H * 4);
}
void Engine::emit_to_dense(const PartIn& in, float* dst, int64_t ld, int B, int cols) {
  if (rec_) {
    PToDense p;
    p.in = in; p.dst = dst; p.ld = ld; p.B = B; p.cols = cols;
    rec_->add(P_TO_DENSE, p);
  } else {
    part_to_dense(ctx_, in, dst, ld, B, cols);
  }
}

// (re)build the decoder weight packs after a parameter change
void Engine::build_decoder_packs() {
  const bool want_inter = dec_fused_ok();
  if (dec_packs_version_ == weights_version_ && dec_packs_inter_ == want_inter) return;
  const int in1 = E + (cfg.input_feed ? Hd : 0);
  const float* P = d_params;
  // forward packs: rows = output units (the UMMA M side), K = concatenated inputs.  The executor's fused GEMM -> cell
  // commands want the gate-interleaved row order, the per-kernel path the plain [gate][unit] order: build what is used
  const bool inter = dec_fused_ok();
  const int64_t gh = inter ? Hd : 0;
  const Pack& W1 = inter ? Wcat1pG : Wcat1p;
  const Pack& W2 = inter ? Wcat2pG : Wcat2p;
  if (cfg.input_feed) split_to_pack(ctx_, P + L.l1_wi + E, 4 * Hd, Hd, in1, 1, sub_cols(W1, 0), Hd, gh);
  split_to_pack(ctx_, P + L.l1_wh, 4 * Hd, Hd, Hd, 1, sub_cols(W1, h1off), Hd, gh);
  split_to_pack(ctx_, P + L.l2_wi, 4 * Hd, Hd, Hd, 1, sub_cols(W2, 0), Hd, gh);
  split_to_pack(ctx_, P + L.l2_wh, 4 * Hd, Hd, Hd, 1, sub_cols(W2, Hd), Hd, gh);
  dec_packs_inter_ = inter;
  split_to_pack(ctx_, P + L.wa, Hd, Hd, Hd, 1, sub_rows(W3p, 0, Hd));                 // rows 0..H-1   : W_a
  split_to_pack(ctx_, P + L.wc + Hd, Hd, Hd, 2 * Hd, 1, sub_rows(W3p, Hd, Hd));        // rows H..2H-1  : W_c[:, H:]
  // backward packs: rows = input units, K = output units (transposes)
  if (cfg.input_feed) split_to_pack(ctx_, P + L.l1_wi + E, Hd, 4 * Hd, 1, in1, sub_rows(Wcat1Tp, 0, Hd));
  split_to_pack(ctx_, P + L.l1_wh, Hd, 4 * Hd, 1, Hd, sub_rows(Wcat1Tp, h1off, Hd));
  split_to_pack(ctx_, P + L.l2_wi, Hd, 4 * Hd, 1, Hd, sub_rows(Wcat2Tp, 0, Hd));
  split_to_pack(ctx_, P + L.l2_wh, Hd, 4 * Hd, 1, Hd, sub_rows(Wcat2Tp, Hd, Hd));
  split_to_pack(ctx_, P + L.wc + Hd, Hd, Hd, 1, 2 * Hd, sub_cols(W3Tp, 0), Hd);        // K 0..H-1  : W_c[:, H:]^T (du)
  split_to_pack(ctx_, P + L.wa, Hd, Hd, 1, Hd, sub_cols(W3Tp, Hd), Hd);                // K H..2H-1 : W_a^T        (dq)
  dec_packs_version_ = weights_version_;
}

// Time-independent half of the output projection: ctxwc[b,s,:] = W_c[:, :H] ctx[b,s,:]  (see AttnOutTc)
void Engine::attention_precompute() {
  if (cfg.gemm_mode == 2) return;
  Gemm g;
  g.M = b_ * S_; g.N = Hd; g.K = Hd;
  g.A = ctx; g.sam = Hd; g.sak = 1;
  g.B = d_params + L.wc; g.sbk = 1; g.sbn = 2 * Hd;
  g.C = CtxWc; g.ldc = Hd;
  gemm(g);
}

// One decoder step.  Saved fp32 state (for backward) has the layout of the SIMT path, minus the context vector:
//   X1[t] = [a_{t-1} | h1_{t-1}]   X2[t] = [h1_t | h2_{t-1}]   CAT[t] = [ - | h2_t]   A_all[t] = a_t   Q[t] = W_a h2_t
// and X1p / X2p / H2p mirror them as bf16 planes (written by the producing bodies, never by a conversion pass).
// Three GEMMs + three bodies per step: gates1 -> cell1 -> gates2 -> cell2 -> [q ; v] = [W_a ; W_c2] h2 -> attention+output.
void Engine::decoder_step_tc(int t, const int32_t* tokens) {
  const int B = b_, S = S_;
  const int64_t RS = slot_rows_ > 0 ? slot_rows_ : b_;     // rows per time slot (>= the active rows B)
  const int nsteps = dec_steps_;
  const bool has_next = (t + 1 < nsteps);
  float* x1 = X1 + (int64_t)t * RS * K1;
  float* x2 = X2 + (int64_t)t * RS * 2 * Hd;
  float* cat = CAT + (int64_t)t * RS * 2 * Hd;
  const int64_t r0 = (int64_t)t * RS, r1 = (int64_t)(t + 1) * RS;
  // ---- layer 1
  const bool fuse = fused_rec() && dec_packs_inter_ && rec_->cluster == cluster_;
  CellFwdTc c1;
  c1.addrows = Ptab; c1.rowsel = tokens; c1.addld = 4 * Hd;
  c1.c_prev = C1 + (int64_t)t * RS * Hd; c1.c_new = C1 + (int64_t)(t + 1) * RS * Hd;
  c1.acts = ACT1 + (int64_t)t * RS * 4 * Hd;
  c1.h_out0 = x2; c1.ld0 = 2 * Hd;
  c1.h_out1 = has_next ? x1 + RS * K1 + h1off : nullptr; c1.ld1 = K1;
  c1.pk0 = pack_out(X2p, r0, 0);
  c1.pk1 = has_next ? pack_out(X1p, r1, h1off) : PackOut();
  c1.B = B; c1.H = Hd;
  if (fuse) {
    emit_gemm_fused(P_GEMM_CELL_FWD, Wcat1pG, 4 * Hd, X1p, r0, 0, K1, c1);
  } else {
    TcOut g1 = emit_gemm(Wcat1p, 4 * Hd, X1p, r0, 0, K1, dec_ws[0]);
    c1.G = part_in(g1, 4 * Hd);
    emit(c1);
  }
  // ---- layer 2
  CellFwdTc c2;
  c2.addrows = bsum2; c2.rowsel = nullptr; c2.addld = 0;
  c2.c_prev = C2 + (int64_t)t * RS * Hd; c2.c_new = C2 + (int64_t)(t + 1) * RS * Hd;
  c2.acts = ACT2 + (int64_t)t * RS * 4 * Hd;
  c2.h_out0 = cat + Hd; c2.ld0 = 2 * Hd;
  c2.h_out1 = has_next ? x2 + RS * 2 * Hd + Hd : nullptr; c2.ld1 = 2 * Hd;
  c2.pk0 = pack_out(H2p, r0, 0);
  c2.pk1 = has_next ? pack_out(X2p, r1, Hd) : PackOut();
  c2.B = B; c2.H = Hd;
  if (fuse) {
    emit_gemm_fused(P_GEMM_CELL_FWD, Wcat2pG, 4 * Hd, X2p, r0, 0, 2 * Hd, c2);
  } else {
    TcOut g2 = emit_gemm(Wcat2p, 4 * Hd, X2p, r0, 0, 2 * Hd, dec_ws[1]);
    c2.G = part_in(g2, 4 * Hd);
    emit(c2);
  }
  // ---- [q ; v] = [W_a ; W_c2] h2, then scores / softmax / alpha-weighted ctxwc + v / tanh in one body
  TcOut g3 = emit_gemm(W3p, 2 * Hd, H2p, r0, 0, Hd, dec_ws[2]);
  AttnOutTc ao;
  ao.ctx = ctx + (int64_t)ctx_row0_ * S * Hd; ao.ctxwc = CtxWc + (int64_t)ctx_row0_ * S * Hd; ao.g3 = part_in(g3, 2 * Hd);
  ao.alpha = ALPHA + (int64_t)t * RS * S; ao.q_out = Q + (int64_t)t * RS * Hd; ao.a_out = A_all + (int64_t)t * RS * Hd;
  ao.x_next = (cfg.input_feed && has_next) ? x1 + RS * K1 : nullptr; ao.ld_next = K1;
  ao.pk_next = (cfg.input_feed && has_next) ? pack_out(X1p, r1, 0) : PackOut();
  ao.B = B; ao.S = S; ao.H = Hd; ao.ctx_rows = dual_rows_;
  if (rec_ && tail_) rec_->add3(P_ATTN_OUT_GEN, ao, tail_->gen, tail_->sel);
  else emit(ao);
}

// per-timestep part of the decoder backward (model.lua:643-661) on tensor cores: three bodies + three GEMMs per step
void Engine::decoder_backward_steps_tc() {
  const int B = b_, S = S_, T = T_;
  if (!rec_) {
    fill_zero(ctx_, dc1, (size_t)B * Hd * sizeof(float));
    fill_zero(ctx_, dc2, (size_t)B * Hd * sizeof(float));
  }
  TcOut dx1, dx2;   // carries of the previous (t+1) iteration; valid when !last
  for (int t = T - 1; t >= 0; t--) {
    const bool last = (t == T - 1);
    // du = (da_prev + W_o^T dz) (1 - a^2); attention backward -> [du | dq]
    AttnDuTc ad;
    ad.ctx = ctx; ad.ctxwc = CtxWc; ad.alpha = ALPHA + (int64_t)t * B * S;
    ad.da_carry = (!last && cfg.input_feed) ? part_in(dx1, K1, 0) : PartIn();
    ad.da_gen = dAgen + (int64_t)t * B * Hd; ad.a = A_all + (int64_t)t * B * Hd;
    ad.du_out = dU + (int64_t)t * B * Hd; ad.de = DE + (int64_t)t * B * S; ad.dq = dQ + (int64_t)t * B * Hd;
    const int64_t rt = (int64_t)t * B;     // operand rows of this step
    ad.pk = pack_out(dUQp, rt, 0); ad.B = B; ad.S = S; ad.H = Hd;
    emit(ad);
    // dh2 (through the output projection and the query) = [du | dq] [W_c2 ; W_a]
    TcOut dh2 = emit_gemm(W3Tp, Hd, dUQp, rt, 0, 2 * Hd, dec_ws[1]);
    CellBwdTc b2;
    b2.dh_a = part_in(dh2, Hd, 0);
    b2.dh_b = last ? PartIn() : part_in(dx2, 2 * Hd, Hd);
    b2.dh_c = PartIn();
    b2.dc = dc2; b2.c_prev = C2 + (int64_t)t * B * Hd; b2.c_new = C2 + (int64_t)(t + 1) * B * Hd;
    b2.acts = ACT2 + (int64_t)t * B * 4 * Hd; b2.dG = dG2 + (int64_t)t * B * 4 * Hd; b2.pk = pack_out(dG2p, rt, 0);
    b2.B = B; b2.H = Hd;
    emit(b2);
    // [dh1 | dh2_prev] = dg2 [W_i2 | W_h2]      (slot 2 must outlive this iteration: read again at t-1)
    dx2 = emit_gemm(Wcat2Tp, 2 * Hd, dG2p, rt, 0, 4 * Hd, dec_ws[2]);
    CellBwdTc b1;
    b1.dh_a = part_in(dx2, 2 * Hd, 0);
    b1.dh_b = last ? PartIn() : part_in(dx1, K1, h1off);
    b1.dh_c = PartIn();
    b1.dc = dc1; b1.c_prev = C1 + (int64_t)t * B * Hd; b1.c_new = C1 + (int64_t)(t + 1) * B * Hd;
    b1.acts = ACT1 + (int64_t)t * B * 4 * Hd; b1.dG = dG1 + (int64_t)t * B * 4 * Hd; b1.pk = pack_out(dG1p, rt, 0);
    b1.B = B; b1.H = Hd;
    emit(b1);
    // [da_prev | dh1_prev] = dg1 [W_i1[:,E:] | W_h1]
    dx1 = emit_gemm(Wcat1Tp, K1, dG1p, rt, 0, 4 * Hd, dec_ws[3]);
  }
  // hand d h1(0) to the encoder backward through the dense dX1 buffer (model.lua:666-667,680-681; quirk Q14)
  emit_to_dense(part_in(dx1, K1, 0), dX1, K1, B, K1);
}

}  // namespace aocr
