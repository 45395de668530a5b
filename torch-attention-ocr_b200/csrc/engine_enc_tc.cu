// engine_enc_tc.cu — the encoder recurrence (src/model/model.lua:293-316, backward :664-690) on tensor cores.
// Per timestep: one tcgen05 GEMM per direction (h_prev W_h^T, batch on the UMMA N side, split-K partials) and ONE
// cell kernel for both directions that sums the partials, adds the time-batched input projection, updates
// (c, h), writes the context column and emits h_t as the next step's bf16 operand planes.
#include "engine.h"

namespace aocr {

namespace {
Pack rows_of(const Pack& p, int64_t row0, int64_t rows) {
  Pack s;
  s.hi = p.hi + row0 * p.kp; s.lo = p.lo + row0 * p.kp; s.rows = rows; s.kp = p.kp;
  return s;
}
PackOut out_of(const Pack& p, int64_t row0) {
  PackOut o;
  o.hi = p.hi + row0 * p.kp; o.lo = p.lo + row0 * p.kp; o.ld = p.kp;
  return o;
}
}  // namespace

void Engine::encoder_forward_steps_tc() {
  const int B = b_, S = S_;
  const int terms = cfg.gemm_mode == 1 ? 1 : 3;
  if (enc_packs_version_ != weights_version_) {
    for (int d = 0; d < 2; d++) {
      split_to_pack(ctx_, d_params + L.enc_wh[d], 4 * He, He, He, 1, Whp[d]);       // rows = gate units
      split_to_pack(ctx_, d_params + L.enc_wh[d], He, 4 * He, 1, He, WhTp[d]);      // rows = input units (transpose)
    }
    enc_packs_version_ = weights_version_;
  }
  // zero initial states as operand planes: fw slot 0, bw slot S
  const size_t slot_bytes = (size_t)B * He * sizeof(__nv_bfloat16);
  fill_zero(ctx_, HencP[0].hi, slot_bytes); fill_zero(ctx_, HencP[0].lo, slot_bytes);
  fill_zero(ctx_, HencP[1].hi + (int64_t)S * B * He, slot_bytes); fill_zero(ctx_, HencP[1].lo + (int64_t)S * B * He, slot_bytes);
  // the two directions are independent chains: forward direction on lane 0, backward direction on lane 2
  prof_begin(2);
  fork_to(2);
  for (int d = 0; d < 2; d++) {
    use_lane(d == 0 ? 0 : 2);
    for (int i = 0; i < S; i++) {
      const int t = d == 0 ? i : S - 1 - i;
      const int prev_slot = d == 0 ? t : t + 1, out_slot = d == 0 ? t + 1 : t;
      TcGemm g;
      g.A = Whp[d]; g.B = rows_of(HencP[d], (int64_t)prev_slot * B, B);
      g.M = 4 * He; g.N = B; g.K = He; g.ldc = 4 * He; g.transpose_out = true;
      g.terms = terms; g.defer_reduce = true; g.ws = dec_ws[d]; g.ws_floats = dec_ws_floats;
      TcOut o = gemm_tc(ctx_, g);
      EncCellFwdTc c;
      c.G[d].p = o.base; c.G[d].nz = o.nz; c.G[d].stride = o.stride; c.G[d].ld = 4 * He;
      c.hp[d] = out_of(HencP[d], (int64_t)out_slot * B);
      c.xg = xg; c.H = Henc; c.Cst = Cenc; c.acts = acts_enc; c.ctx = ctx; c.B = B; c.S = S; c.He = He; c.step = i;
      c.d_only = d;
      enc_cell_fwd_tc(ctx_, c);
    }
  }
  use_lane(0);
  join_from(2);
  prof_end(2, 2.0 * 2 * S * (double)B * He * 4 * He);
}

void Engine::encoder_backward_steps_tc() {
  const int B = b_, S = S_;
  const int terms = cfg.gemm_mode == 1 ? 1 : 3;
  const int64_t slot = (int64_t)B * He;
  PartIn dh[2];
  for (int d = 0; d < 2; d++) { dh[d].p = enc_dh + d * slot; dh[d].nz = 1; dh[d].stride = 0; dh[d].ld = He; }   // decoder seeds
  fork_to(2);
  for (int d = 0; d < 2; d++) {
    use_lane(d == 0 ? 0 : 2);
    for (int i = 0; i < S; i++) {
      EncCellBwdTc c;
      c.dh[d] = dh[d];
      c.Cst = Cenc; c.acts = acts_enc; c.Dctx = Dctx; c.dc = enc_dc; c.dG = dGe;
      c.dgp[d] = out_of(dGeP[d], 0);
      c.B = B; c.S = S; c.He = He; c.step = i; c.d_only = d;
      enc_cell_bwd_tc(ctx_, c);
      TcGemm g;                      // dh_prev = dG_t W_h : rows of W_h^T on the M side, K = 4He
      g.A = WhTp[d]; g.B = dGeP[d];
      g.M = He; g.N = B; g.K = 4 * He; g.ldc = He; g.transpose_out = true;
      g.terms = terms; g.defer_reduce = true; g.ws = dec_ws[2 + d]; g.ws_floats = dec_ws_floats;
      TcOut o = gemm_tc(ctx_, g);
      dh[d].p = o.base; dh[d].nz = o.nz; dh[d].stride = o.stride; dh[d].ld = He;
    }
  }
  use_lane(0);
  join_from(2);
}

}  // namespace aocr
