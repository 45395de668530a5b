// kernels.h — host launchers of the non-GEMM kernels of libaocr (definitions in kernels_*.cu).
#pragma once
#include "common.cuh"

namespace aocr {

// ---------------- CNN (reference: src/model/cnn.lua:9-45) -------------------------------------
// K1+K2: (x-128)/128 -> conv1 3x3 p1 (1->64) + bias + ReLU + maxpool 2x2.  x (B,32,W) -> a1 (B,16,W/2,64) NHWC
// (phi, plo), where given: the output also as bf16 (hi, lo) operand planes for the next tensor-core contraction
void conv1_fwd(Ctx&, const float* x, const float* w /*[64][9]*/, const float* bias, float* a1, uint8_t* idx,
               int B, int W, __nv_bfloat16* phi = nullptr, __nv_bfloat16* plo = nullptr);
// dW1 (64x9), db1 from d a1 routed through the pool argmax / ReLU mask.  Deterministic two-stage reduction.
void conv1_bwd(Ctx&, const float* x, const float* a1, const uint8_t* idx, const float* da1, float* dw, float* db,
               float* partial /* [nblk][640] scratch */, int nblk, int B, int W);
// ReLU + maxpool (KH x KW in {2x2, 2x1}), floor mode.  z (B,H,Wi,C) -> a (B,H/2,Wi/KW,C) + argmax index.
void relu_pool_fwd(Ctx&, const float* z, float* a, uint8_t* idx, int B, int H, int Wi, int C, int kw,
                   __nv_bfloat16* phi = nullptr, __nv_bfloat16* plo = nullptr);
// dz (full, zero-filled) from da through argmax + (a>0)
void relu_pool_bwd(Ctx&, const float* da, const float* a, const uint8_t* idx, float* dz, int B, int H, int Wi, int C,
                   int kw, __nv_bfloat16* phi = nullptr, __nv_bfloat16* plo = nullptr);
// im2col of an NHWC activation: a (B,H,Wi,C) -> col (B*Ho*Wo, k*k*C), tap-major / channel-fastest.
// flip=1 reads taps mirrored (used for the data gradient, where `a` is dz and the GEMM weight is W^T).
void im2col(Ctx&, const float* a, float* col, int B, int H, int Wi, int C, int k, int pad);
// column statistics of z (R,C): sum and (two-pass) centred sum of squares; deterministic.
void col_sum(Ctx&, const float* z, int64_t R, int C, float* out /*[C]*/, float* partial, int accumulate);
// Scratch of the column reductions: `partial` buffers hold kPartialFloats floats; their last kSlabCounters words are the
// arrival counters of the "last slab block finalises" reductions (one per block of 32 columns; zero at allocation,
// self-resetting), one buffer per lane.
constexpr int64_t kPartialFloats = (int64_t)256 * 8192;
constexpr int kSlabCounters = 1024;
// cross-rank summation hook for batch-norm statistics (data parallelism); world == 1: unused
// grows: global rows / local rows of a batch-norm reduction (= global batch / this rank's batch; the shards need not be equal)
struct StatSync { void (*fn)(void* user, float* buf, int64_t n) = nullptr; void* user = nullptr; int world = 1; double grows = 1.0; };
// one pass over z: batch mean / biased variance (global over the data-parallel ranks: ONE all-reduce of [sum | sum sq])
// and the running-statistics update (momentum 0.1, unbiased variance) [T7 nn.SpatialBatchNormalization]
void bn_stats(Ctx&, const float* z, int64_t R, int C, float* mean, float* var /*biased*/, float* rmean, float* rvar,
              float* partial, float* sums /* [2C] scratch */, const StatSync& sync);
// running stats update (momentum 0.1, unbiased var) [T7 nn.SpatialBatchNormalization]
void bn_update_running(Ctx&, const float* mean, const float* var, float* rmean, float* rvar, int C, int64_t R);
// y = relu(gamma*(z-mean)/sqrt(var+eps)+beta).  If tm_S>0 rows (n,s) are written time-major to row s*tm_B+n.
void bn_relu_fwd(Ctx&, const float* z, const float* mean, const float* var, const float* gamma, const float* beta,
                 float* a, int64_t R, int C, int tm_S, int tm_B, __nv_bfloat16* phi = nullptr, __nv_bfloat16* plo = nullptr);
// BN backward.  step 1: dy = da*(a>0) written to dz; sums s1=sum(dy), s2=sum(dy*xhat) via col reductions;
// step 2: dz = gamma*inv*(dy - s1/R - xhat*s2/R) (train) or gamma*inv*dy (eval stats).  dgamma=s2, dbeta=s1.
void bn_relu_bwd(Ctx&, const float* da, const float* a, const float* z, const float* mean, const float* var,
                 const float* gamma, float* dz, float* dgamma, float* dbeta, float* partial, int64_t R, int C,
                 int tm_S, int tm_B, int train, const StatSync& sync, __nv_bfloat16* phi = nullptr,
                 __nv_bfloat16* plo = nullptr);

// ---------------- LSTM cells (reference: src/model/LSTM.lua:79-105) ---------------------------
// One encoder step for both directions: g = xg[t] + h_prev W_h^T ; cell.  Layouts in DESIGN.md §4.
struct EncStep {
  const float* xg;      // (S*B, 8He): [fw 4He | bw 4He], biases folded
  const float* Wh[2];   // (4He, He) per direction
  float* H;             // (2, S+1, B, He)   slot convention in engine.cu
  float* Cst;           // (2, S+1, B, He)
  float* acts;          // (2, S, B, 4, He)
  float* ctx;           // (B, S, 2He)
  int B, S, He, step;   // step i: fw processes t=i, bw processes t=S-1-i
};
void enc_step_fwd(Ctx&, const EncStep&);
struct EncStepBwd {
  const float* Wh[2];
  const float* Cst; const float* acts; const float* Dctx;   // Dctx (B,S,2He)
  float* dh;      // (2,B,He) carry: gradient wrt h output of the current step (excluding Dctx)
  float* dc;      // (2,B,He) carry
  float* dG;      // (S*B, 8He) gate pre-activation grads, [fw|bw]
  int B, S, He, step;
};
void enc_cell_bwd(Ctx&, const EncStepBwd&);   // writes dG rows of this step, updates dc; dh_prev done by GEMM

// Decoder cell: gates = G (B,4H) + addrow[rowsel ? rowsel[b]-1 : 0] ; writes c_new, acts, h to two sinks.
struct DecCell {
  const float* G; const float* addrows; const int32_t* rowsel; int64_t addld;
  const float* c_prev; float* c_new; float* acts;      // acts (B,4,H)
  float* h_out0; int64_t ld0; float* h_out1; int64_t ld1;   // h_out1 may be null
  int B, H;
};
void dec_cell_fwd(Ctx&, const DecCell&);
struct DecCellBwd {
  const float* dh_a; int64_t lda; const float* dh_b; int64_t ldb; const float* dh_c; int64_t ldc;  // summed (null ok)
  float* dc;             // in/out carry (B,H)
  const float* c_prev; const float* c_new; const float* acts; float* dG;   // dG (B,4H)
  int B, H;
};
void dec_cell_bwd(Ctx&, const DecCellBwd&);

// ---------------- attention (reference: src/model/LSTM.lua:124-162) ---------------------------
// one CTA per batch row: e_s = ctx_s . q ; alpha = softmax_s(e) (mask all-ones) ; cv = sum alpha_s ctx_s
// single pass over ctx (online softmax).  cv written with row stride ldcv.
void attn_fwd(Ctx&, const float* ctx, const float* q, float* alpha, float* cv, int64_t ldcv, int B, int S, int H);
// dalpha_s = dcv.ctx_s ; de = alpha*(dalpha - sum alpha*dalpha) ; dq = sum_s de_s ctx_s
void attn_bwd(Ctx&, const float* ctx, const float* alpha, const float* dcv, int64_t lddcv, float* de, float* dq,
              int B, int S, int H);

// ---------------- generator + criterion (output_projector.lua:5-6, criterion.lua:4-7) ---------
// rows r = (t,b).  logp = log_softmax(W a + b).  y = targets_eval (1-based; PAD=1 has weight 0).
// Writes logp (R,V); if dz: dz = (softmax - onehot) * w/Bn ; rowloss[r] = -w*logp[y].
void generator_fwd(Ctx&, const float* a, const float* W, const float* bias, const int32_t* y, float* logp, float* dz,
                   float* rowloss, int64_t R, int H, int V, float inv_bn);
void reduce_sum_double(Ctx&, const float* v, int64_t n, double* out);   // deterministic single-block tree
// greedy selection (model.lua:402,448-458): sticky-PAD edit, argmax, score accumulate, next token
void greedy_select(Ctx&, float* logp /*(B,V) edited in place*/, int32_t* tok /*in: prev, out: new*/, double* score,
                   int32_t* labels, int64_t ldl, int t, int B, int V);

// ---------------- small helpers ---------------------------------------------------------------
void fill_zero(Ctx&, void* p, size_t bytes);
void add_vec(Ctx&, float* out, const float* a, const float* b, int64_t n);                       // out = a + b
void copy_strided(Ctx&, float* dst, int64_t ldd, const float* src, int64_t lds, int rows, int cols);
void concat_enc_finals(Ctx&, const float* src_fw, const float* src_bw, float* dst, int64_t ldd, int B, int He);
void du_from_da(Ctx&, const float* da_carry, int64_t ldc, const float* da_gen, const float* a, float* du, int64_t n,
                int H);
void gather_tokens(Ctx&, const int32_t* tgt_bt /*(B,T)*/, int32_t* out_tb /*(T,B)*/, int B, int T, int Tpad,
                   int32_t padval);
// segment-sum of dG1 rows by token id: dP[v] = sum_{r: y[r]==v} dG[r]  (V x N); deterministic per column
void token_segment_sum(Ctx&, const float* dG, const int32_t* y_tb, float* dP, int64_t R, int N, int V);
void onehot(Ctx&, const int32_t* y, float* oh, int64_t R, int V);
// out (M x E) = A (M x K, lda) * W (K x E, ldw) for E <= 32 (embedding-sized right-hand sides)
void thin_n_gemm(Ctx&, const float* A, int64_t lda, const float* W, int64_t ldw, float* out, int64_t ldo, int M, int K,
                 int E);

// ---------------- optimiser (src/optim/optim_sgd.lua:49-52,90) --------------------------------
void sumsq_partial(Ctx&, const float* v, int64_t n, double* partial, int nblk);   // partial[nblk]
void sumsq_final(Ctx&, const double* partial, int nblk, double* out);
// p -= lr * scale * g with scale = (norm>clip ? clip/norm : 1), norm = sqrt(*sumsq)
void sgd_apply(Ctx&, float* p, float* g, int64_t n, const double* sumsq, double lr, double clip);
// the three steps above for all 5 parameter groups in 3 launches; nb[g] <= 1024 blocks work on group g (fixed
// assignment: deterministic), partial is [5][1024] doubles, sumsq [5]
struct SgdGroups { int64_t off[5]; int64_t n[5]; int nb[5]; };
// lrclip: device pointer to {lr, clip} (the step size is data, not a launch parameter: a captured step is replayed with any lr)
void sgd_groups(Ctx&, float* params, float* grads, const SgdGroups& G, double* partial, double* sumsq, const double* lrclip);
void scale_vec(Ctx&, float* v, int64_t n, float s);

// ---- beam search / dictionary-constrained decode (kernels_beam.cu; model.lua:380-387,405-536,573-585)
// state rows are beam-major: row = k*Bc + b (k < K beams, b < Bc images of the chunk)
struct BeamSelect {
  const float* logp;          // (K*Bc, V) log-probs of this step (first step: only the rows of beam 0 are read)
  const int32_t* tok;         // (K*Bc) input tokens of this step (sticky PAD test)
  const double* scores;       // (Bc, K) beam scores so far
  double* new_scores;         // (Bc, K)
  int32_t* tok_out;           // (K*Bc) tokens chosen = inputs of the next step
  int32_t* parent_row;        // (K*Bc) state row each new beam continues
  int32_t* hist_tok; int32_t* hist_par;   // (L, Bc, K) history for the backtrack
  const int32_t* trie;        // (nodes, V+1) child table or nullptr; column = 1-based vocabulary id, root = node 0
  const int32_t* loc; int32_t* new_loc;   // (Bc, K) trie node per beam
  int t, Bc, K, V;
};
void beam_select(Ctx&, const BeamSelect&);
struct BeamGatherTensor { uint8_t* ptr; int64_t pitch_bytes, row_bytes, tmp_off; };
struct BeamGather { BeamGatherTensor t[10]; int n; int rows; const int32_t* parent_row; uint8_t* tmp; };
void beam_gather(Ctx&, const BeamGather&);
struct BeamBacktrack { const double* scores; const int32_t* hist_tok; const int32_t* hist_par; int32_t* labels; int64_t ldl;
                       double* score_out; int Bc, K, L; };
void beam_backtrack(Ctx&, const BeamBacktrack&);
void axpy_vec(Ctx&, float* y, const float* x, int64_t n, float a);

}  // namespace aocr
