// kernels_dec.h — kernels of the tensor-core decoder path (definitions in kernels_dec.cu).
#pragma once
#include "common.cuh"

namespace aocr {

// value(b, j) = sum_{z < nz} p[z*stride + b*ld + j] : a GEMM result still split along K (gemm_tc TcOut)
struct PartIn { const float* p = nullptr; int nz = 1; int64_t stride = 0; int64_t ld = 0; };
// destination of a bf16 (hi, lo) operand plane pair; element (b, j) at [b*ld + j]; hi == nullptr: skip
struct PackOut { __nv_bfloat16* hi = nullptr; __nv_bfloat16* lo = nullptr; int64_t ld = 0; };

struct CellFwdTc {
  PartIn G;                                   // gate pre-activations (B, 4H), without bias
  const float* addrows; const int32_t* rowsel; int64_t addld;   // + addrows[(rowsel[b]-1)*addld] (embedding table / bias)
  const float* c_prev; float* c_new; float* acts;
  float* h_out0; int64_t ld0; float* h_out1; int64_t ld1;       // fp32 sinks (h_out1 may be null)
  PackOut pk0, pk1;                                             // bf16 sinks = operands of the next GEMMs
  int B, H;
};
void cell_fwd_tc(Ctx&, const CellFwdTc&);

struct CellBwdTc {
  PartIn dh_a, dh_b, dh_c;                    // summed; p == nullptr: absent
  float* dc; const float* c_prev; const float* c_new; const float* acts;
  float* dG; PackOut pk;                      // (B, 4H)
  int B, H;
};
void cell_bwd_tc(Ctx&, const CellBwdTc&);

// dst[b*ld + j] = value(b, j): materialise a split result (used once per step for the encoder seeds)
void part_to_dense(Ctx&, const PartIn& in, float* dst, int64_t ld, int B, int cols);

// Attention + decoder output of one step in ONE body.  The output projection W_c [cv ; h2] is split as
//   W_c1 cv + W_c2 h2 = sum_s alpha_s (W_c1 ctx_s) + W_c2 h2,
// with ctxwc = ctx W_c1^T precomputed once per batch (time-independent) and [q ; v] = [W_a ; W_c2] h2 from one GEMM:
// scores, softmax, the alpha-weighted sum of ctxwc rows, + v, tanh.  (model.lua:553-568 / LSTM.lua:124-162 math,
// re-associated; removes one GEMM and one pointwise pass per step.)
struct AttnOutTc {
  const float* ctx; const float* ctxwc;       // (B, S, H) each
  PartIn g3;                                  // (B, 2H): [q | v] split-K partials
  float* alpha; float* q_out;                 // (B, S), (B, H) saved for backward
  float* a_out;                               // (B, H) a_t
  float* x_next; int64_t ld_next; PackOut pk_next;   // input feed of step t+1 (may be null)
  int B, S, H;
  int ctx_rows = 0;                           // > 0: batch row b attends over source b % ctx_rows (dual pass)
};
void attn_out_tc(Ctx&, const AttnOutTc&);
// Backward of the same: du = (da_gen + da_carry) (1 - a^2);  d alpha_s = ctxwc_s . du;  de = softmax backward;
// dq = sum_s de_s ctx_s.  Writes [du | dq] as the operand of the [W_c2 ; W_a]^T GEMM.
struct AttnDuTc {
  const float* ctx; const float* ctxwc; const float* alpha;
  PartIn da_carry; const float* da_gen; const float* a;
  float* du_out; float* de; float* dq;
  PackOut pk;                                 // (B, 2H): du at column 0, dq at column H
  int B, S, H;
};
void attn_du_tc(Ctx&, const AttnDuTc&);
// generator + greedy selection as executor commands (greedy decode, model.lua:393-404,446-459)
struct GenTc {
  const float* a; const float* W; const float* bias; const int32_t* y; float* logp; float* dz; float* rowloss;
  int R, H, V; float inv_bn;
  // dual pass (greedy rows [0, split) + teacher-forced rows [split, R) in one batch): rows >= split write their
  // log-probs to logp2 and take y / rowloss at index r - split; rows < split have no target.  split = 0: off.
  int split = 0; float* logp2 = nullptr;
};
// tok: token chosen at step t-1 (read for the sticky-PAD rule); tok_out: where the token of step t goes (may alias tok)
struct GreedyTc { float* logp; const int32_t* tok; int32_t* tok_out; double* score; int32_t* labels; long long ldl; int t, B, V; };
inline size_t attn_smem_bytes(int S, int H) { return (size_t)(((S + 3) & ~3) + 16 + 9 * H) * sizeof(float); }

}  // namespace aocr

namespace aocr {
// ---- encoder recurrence on tensor cores: one cell kernel per step handles both directions
struct EncCellFwdTc {
  PartIn G[2];          // h_prev W_h^T per direction, (B, 4He), split-K partials
  const float* xg;      // (S*B, 8He) input projections incl. biases, [fw | bw]
  float* H; float* Cst; float* acts; float* ctx;     // layouts of kernels.h:EncStep
  PackOut hp[2];        // h_t as bf16 planes = next step's GEMM operand, per direction (row 0 = batch row 0)
  int B, S, He, step;
  int d_only = -1;      // -1: both directions in one launch; 0/1: that direction only (directions on separate streams)
};
void enc_cell_fwd_tc(Ctx&, const EncCellFwdTc&);
struct EncCellBwdTc {
  PartIn dh[2];         // recurrent part of d h_t per direction (B, He)
  const float* Cst; const float* acts; const float* Dctx;
  float* dc;            // (2,B,He) carry
  float* dG;            // (S*B, 8He)
  PackOut dgp[2];       // d gates of this step as bf16 planes (B, 4He) per direction
  int B, S, He, step;
  int d_only = -1;
};
void enc_cell_bwd_tc(Ctx&, const EncCellBwdTc&);
}  // namespace aocr
