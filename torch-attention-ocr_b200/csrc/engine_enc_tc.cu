// engine_enc_tc.cu — the encoder recurrence (src/model/model.lua:293-316, backward :664-690) on tensor cores.
// Per timestep and direction: one tcgen05 GEMM (h_prev W_h^T, batch on the UMMA N side, split-K partials) and one
// cell body that sums the partials, adds the time-batched input projection, updates (c, h), writes the context
// column and emits h_t as the next step's bf16 operand planes.  The two directions are independent chains and run
// concurrently (lane 0 / lane 2), either as per-step kernels or as one persistent executor kernel each.
#include "engine.h"

namespace aocr {

namespace {
PackOut out_of(const Pack& p, int64_t row0) {
  PackOut o;
  o.hi = p.hi + row0 * p.kp; o.lo = p.lo + row0 * p.kp; o.ld = p.kp;
  return o;
}
}  // namespace

void Engine::encoder_dir_forward(int d) {
  const int B = b_, S = S_;
  for (int i = 0; i < S; i++) {
    const int t = d == 0 ? i : S - 1 - i;
    const int prev_slot = d == 0 ? t : t + 1, out_slot = d == 0 ? t + 1 : t;
    EncCellFwdTc c;
    c.hp[d] = out_of(HencP[d], (int64_t)out_slot * B);
    c.xg = xg; c.H = Henc; c.Cst = Cenc; c.acts = acts_enc; c.ctx = ctx; c.B = B; c.S = S; c.He = He; c.step = i;
    c.d_only = d;
    if (rec_ && fuse_on_ && rec_->cluster > 1 && He % 32 == 0 && pad64(He) / 64 >= rec_->cluster) {
      // one command: GEMM tile per cluster (rows gate-interleaved), split-K over the cluster ranks, DSMEM reduce + cell
      emit_gemm_fused(P_GEMM_ENC_FWD, WhpG[d], 4 * He, HencP[d], (int64_t)prev_slot * B, 0, He, c);
      continue;
    }
    TcOut o = emit_gemm(Whp[d], 4 * He, HencP[d], (int64_t)prev_slot * B, 0, He, dec_ws[d]);
    c.G[d].p = o.base; c.G[d].nz = o.nz; c.G[d].stride = o.stride; c.G[d].ld = 4 * He;
    emit(c);
  }
}

void Engine::encoder_dir_backward(int d) {
  const int B = b_, S = S_;
  const int64_t slot = (int64_t)B * He;
  PartIn dh;
  dh.p = enc_dh + d * slot; dh.nz = 1; dh.stride = 0; dh.ld = He;    // decoder seeds
  auto cell_of = [&](int i) {
    const int t = d == 0 ? S - 1 - i : i;            // the time index this backward step works on
    EncCellBwdTc c;
    c.Cst = Cenc; c.acts = acts_enc; c.Dctx = Dctx; c.dc = enc_dc; c.dG = dGe;
    c.dgp[d] = out_of(dGeP[d], (int64_t)t * B);      // kept for all timesteps: the weight gradients read the planes
    c.B = B; c.S = S; c.He = He; c.step = i; c.d_only = d;
    return c;
  };
  // (a fused GEMM -> cell-backward command was measured and lost: M = He gives only 4 tiles = 16 CTAs per direction)
  for (int i = 0; i < S; i++) {
    EncCellBwdTc c = cell_of(i);
    c.dh[d] = dh;
    emit(c);
    // dh_prev = dG_t W_h : rows of W_h^T on the M side, K = 4He
    const int t = d == 0 ? S - 1 - i : i;
    TcOut o = emit_gemm(WhTp[d], He, dGeP[d], (int64_t)t * B, 0, 4 * He, dec_ws[2 + d]);
    dh.p = o.base; dh.nz = o.nz; dh.stride = o.stride; dh.ld = He;
  }
}

// operand planes of the encoder's recurrent weights (forward, gate-interleaved forward, transposed for the backward)
void Engine::ensure_enc_packs() {
  if (enc_packs_version_ == weights_version_) return;
  for (int d = 0; d < 2; d++) {
    split_to_pack(ctx_, d_params + L.enc_wh[d], 4 * He, He, He, 1, Whp[d]);       // rows = gate units
    if (He % 32 == 0) split_to_pack(ctx_, d_params + L.enc_wh[d], 4 * He, He, He, 1, WhpG[d], -1, He);   // gate-interleaved
    split_to_pack(ctx_, d_params + L.enc_wh[d], He, 4 * He, 1, He, WhTp[d]);      // rows = input units (transpose)
  }
  enc_packs_version_ = weights_version_;
}

void Engine::encoder_forward_steps_tc() {
  const int B = b_, S = S_;
  ensure_enc_packs();
  // zero initial states as operand planes: fw slot 0, bw slot S
  const size_t slot_bytes = (size_t)B * He * sizeof(__nv_bfloat16);
  fill_zero(ctx_, HencP[0].hi, slot_bytes); fill_zero(ctx_, HencP[0].lo, slot_bytes);
  fill_zero(ctx_, HencP[1].hi + (int64_t)S * B * He, slot_bytes); fill_zero(ctx_, HencP[1].lo + (int64_t)S * B * He, slot_bytes);
  fork_to(2);
  for (int d = 0; d < 2; d++) {
    use_lane(d == 0 ? 0 : 2);
    if (persist_on_) run_program(PK_ENC_FWD0 + d, S, 0);
    else encoder_dir_forward(d);
  }
  use_lane(0);
  join_from(2);
}

void Engine::encoder_backward_steps_tc() {
  fork_to(2);
  for (int d = 0; d < 2; d++) {
    use_lane(d == 0 ? 0 : 2);
    if (persist_on_) run_program(PK_ENC_BWD0 + d, S_, 0);
    else encoder_dir_backward(d);
  }
  use_lane(0);
  join_from(2);
}

}  // namespace aocr
