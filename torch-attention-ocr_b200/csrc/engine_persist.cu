// engine_persist.cu — builds, caches and launches the command lists of the persistent recurrence executor.
// A program is recorded by running the ordinary step functions with `rec_` set (every emit() appends a command
// instead of launching a kernel), so the per-kernel path and the persistent path share one description of the math.
#include "engine.h"

namespace aocr {

void Engine::run_program(int kind, int nsteps, int variant) {
  ProgKey key{kind, b_, S_, nsteps, variant};
  auto it = programs_.find(key);
  if (it == programs_.end() && programs_.size() >= kMaxPrograms) {
    // variable S / T per batch: the cache is bounded.  Captured graphs reference the programs' device buffers.
    drop_graphs();
    for (int l = 1; l < 3; l++) if (lanes_on_ && lanes_[l].st) AOCR_CUDA(cudaStreamSynchronize(lanes_[l].st));
    for (auto& kv : programs_) persist_free(kv.second);
    programs_.clear();
  }
  if (it == programs_.end()) {
    PersistProgram prog;
    int bn = b_ > 128 ? 256 : (b_ > 64 ? 128 : (b_ > 32 ? 64 : (b_ > 16 ? 32 : 16)));
    AOCR_CHECK(b_ <= 256, "persistent executor: at most 256 rows per command (UMMA N <= 256)");
    prog.bn = bn;
    const bool enc = (kind == PK_ENC_FWD0 || kind == PK_ENC_FWD0 + 1 || kind == PK_ENC_BWD0 || kind == PK_ENC_BWD0 + 1);
    int cap = persist_max_ctas(bn);
    rec_max_ctas_ = enc ? (cap / 2 < 64 ? cap / 2 : 64) : (cap < 128 ? cap : 128);   // both encoder directions co-resident
    prog.grid = 1;
    prog.cluster = cluster_;
    rec_ = &prog;
    try {
      switch (kind) {
        case PK_DEC_FWD: {
          const int save = dec_steps_;
          dec_steps_ = nsteps;
          for (int t = 0; t < nsteps; t++) decoder_step_tc(t, tgt_tb + (int64_t)t * b_);
          dec_steps_ = save;
          break;
        }
        case PK_DEC_GREEDY: {   // greedy pass: the argmax of step t feeds the embedding lookup of step t+1 (device-side)
          const int save = dec_steps_;
          dec_steps_ = nsteps;
          for (int t = 0; t < nsteps; t++) {
            decoder_step_tc(t, tok);
            GenTc gp;
            gp.a = A_all + (int64_t)t * b_ * Hd; gp.W = d_params + L.wo; gp.bias = d_params + L.bo; gp.y = nullptr;
            gp.logp = logp[1] + (int64_t)t * b_ * V; gp.dz = nullptr; gp.rowloss = nullptr;
            gp.R = b_; gp.H = Hd; gp.V = V; gp.inv_bn = 1.0f;
            prog.add(P_GENERATOR, gp);
            GreedyTc gs;
            gs.logp = gp.logp; gs.tok = tok; gs.tok_out = tok; gs.score = score; gs.labels = labels; gs.ldl = Tmax; gs.t = t; gs.B = b_; gs.V = V;
            prog.add(P_GREEDY, gs);
          }
          dec_steps_ = save;
          break;
        }
        case PK_DEC_DUAL_TAIL: {   // greedy rows only, steps [variant, nsteps), on the dual pass's layout (slot_rows_ = 2B)
          const int save = dec_steps_;
          const int64_t RS = slot_rows_;
          dec_steps_ = nsteps;
          for (int t = variant; t < nsteps; t++) {
            StepTail tl;
            GenTc& gp = tl.gen;
            gp.a = A_all + (int64_t)t * RS * Hd; gp.W = d_params + L.wo; gp.bias = d_params + L.bo;
            gp.y = nullptr; gp.logp = logp[1] + (int64_t)t * b_ * V; gp.logp2 = nullptr;
            gp.dz = nullptr; gp.rowloss = nullptr; gp.split = b_;          // every row is a greedy row
            gp.R = b_; gp.H = Hd; gp.V = V; gp.inv_bn = 1.0f;
            GreedyTc& gs = tl.sel;
            gs.logp = gp.logp; gs.tok = tokseq + (int64_t)t * RS; gs.tok_out = tokseq + (int64_t)(t + 1) * RS;
            gs.score = score; gs.labels = labels; gs.ldl = Tmax; gs.t = t; gs.B = b_; gs.V = V;
            tail_ = &tl;
            try { decoder_step_tc(t, tokseq + (int64_t)t * RS); } catch (...) { tail_ = nullptr; throw; }
            tail_ = nullptr;
          }
          dec_steps_ = save;
          break;
        }
        case PK_DEC_DUAL: {     // b_ = 2B here: rows [0,B) greedy (argmax feedback), rows [B,2B) teacher forced
          // variant > 0: the recurrence goes on after this program (greedy-only tail) for `variant` steps in total, so the
          // last step still writes the next step's inputs
          const int save = dec_steps_, Bh = dual_rows_;
          dec_steps_ = variant > 0 ? variant : nsteps;
          for (int t = 0; t < nsteps; t++) {
            StepTail tl;
            GenTc& gp = tl.gen;
            gp.a = A_all + (int64_t)t * b_ * Hd; gp.W = d_params + L.wo; gp.bias = d_params + L.bo;
            gp.y = tev_tb + (int64_t)t * Bh; gp.logp = logp[1] + (int64_t)t * Bh * V; gp.logp2 = logp[2] + (int64_t)t * Bh * V;
            gp.dz = nullptr; gp.rowloss = rowloss + (int64_t)t * Bh; gp.split = Bh;
            gp.R = b_; gp.H = Hd; gp.V = V; gp.inv_bn = 1.0f;
            GreedyTc& gs = tl.sel;
            gs.logp = gp.logp; gs.tok = tokseq + (int64_t)t * b_; gs.tok_out = tokseq + (int64_t)(t + 1) * b_;
            gs.score = score; gs.labels = labels; gs.ldl = Tmax; gs.t = t; gs.B = Bh; gs.V = V;
            tail_ = &tl;                // generator + selection ride on the step's attention+output command
            try { decoder_step_tc(t, tokseq + (int64_t)t * b_); } catch (...) { tail_ = nullptr; throw; }
            tail_ = nullptr;
          }
          dec_steps_ = save;
          break;
        }
        case PK_DEC_BWD: decoder_backward_steps_tc(); break;
        case PK_ENC_FWD0: encoder_dir_forward(0); break;
        case PK_ENC_FWD0 + 1: encoder_dir_forward(1); break;
        case PK_ENC_BWD0: encoder_dir_backward(0); break;
        case PK_ENC_BWD0 + 1: encoder_dir_backward(1); break;
        default: throw InvalidError("unknown persistent program kind");
      }
    } catch (...) {
      rec_ = nullptr;
      throw;
    }
    rec_ = nullptr;
    // grid = the largest number of GEMM tiles of any command.  (Giving the encoder-backward bodies 64 CTAs instead of
    // 32 halves their time but changes nothing end to end: the executor's SMs are taken from the weight-gradient lanes.)
    if (prog.grid < 16) prog.grid = 16;
    if (prog.grid > rec_max_ctas_) prog.grid = rec_max_ctas_;
    prog.grid = ((prog.grid + prog.cluster - 1) / prog.cluster) * prog.cluster;
    it = programs_.emplace(key, std::move(prog)).first;
  }
  PersistProgram& prog = it->second;
  if (!prog.uploaded) persist_upload(ctx_, prog);
  double flops = 0.0, bytes = 0.0;
  if (prof_on)
    for (const PCmd& c : prog.cmds)
      if (c.type == P_GEMM || c.type == P_GEMM_ENC_FWD || c.type == P_GEMM_CELL_FWD) {
        const PGemm* g = reinterpret_cast<const PGemm*>(c.payload);
        const double kp = (double)g->num_kb * 64.0, planes = g->terms == 3 ? 2.0 : 1.0;
        flops += 2.0 * g->M * (double)g->N * kp;
        bytes += ((double)g->M + (double)g->N) * kp * 2.0 * planes;   // every weight / activation plane element once
      }
  prof_begin(3);
  prof_begin(2);
  persist_launch(ctx_, prog);
  prof_end(2, flops);
  prof_end(3, bytes);
}

}  // namespace aocr
