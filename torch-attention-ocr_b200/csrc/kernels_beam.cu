// kernels_beam.cu — device side of beam search (beam > 1) and dictionary-constrained decode
// (src/model/model.lua:380-387,405-445,460-536,573-585).  The replicated state holds beam*Bc rows, BEAM-major
// (row = k*Bc + b): the context row of a state row is row % Bc, so the attention kernels of the greedy path serve
// unchanged.  Per step: beam_select (sticky PAD, totals, top-k or the sorted walk over trie-valid continuations, new
// trie nodes, history), beam_gather (parent re-gather of every state tensor); after the last step beam_backtrack.
#include "kernels.h"

namespace aocr {

namespace {

constexpr int kMaxCand = 39 * 39 + 64;   // beam <= V <= 64 would need 4096; engine checks beam*V <= kBeamCandMax
constexpr int kSelThreads = 128;

// one CTA per image.  Candidates j = k*V + v (beam k of the previous step, vocabulary id v+1).
__global__ void __launch_bounds__(kSelThreads) beam_select_kernel(BeamSelect p) {
  extern __shared__ float total[];                       // nb*V candidate totals, -inf = not admissible
  __shared__ float red_v[kSelThreads / 32];
  __shared__ int red_i[kSelThreads / 32];
  __shared__ int picked[64];
  const int b = blockIdx.x, V = p.V, K = p.K, Bc = p.Bc;
  const int nb = p.t == 0 ? 1 : K;                       // first step: every replica of an image holds the same state
  const int n = nb * V;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const int k = j / V, v = j % V;
    const int r = k * Bc + b;
    float lp = p.logp[(int64_t)r * V + v];
    if (p.t > 0 && v == 0) {                             // sticky PAD: log-prob[PAD] <- 0 after PAD / EOS (model.lua:448-449)
      const int prev = p.tok[r];
      if (prev == 1 || prev == 3) lp = 0.f;
    }
    float tot = lp + (p.t > 0 ? (float)p.scores[(int64_t)b * K + k] : 0.f);
    if (p.trie) {                                        // model.lua:417 (first step: children of the root), :472
      const int node = p.t == 0 ? 0 : p.loc[(int64_t)b * K + k];
      const bool ok = (p.t > 0 && v == 0) || p.trie[(int64_t)node * (V + 1) + v + 1] >= 0;
      if (!ok) tot = -INFINITY;
    }
    total[j] = tot;
  }
  __syncthreads();
  // K rounds of arg-max (ties: lowest candidate index; Torch's topk leaves the order of ties unspecified)
  for (int round = 0; round < K; round++) {
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
      const float x = total[j];
      if (x > best || (x == best && j < bi && x != -INFINITY)) { best = x; bi = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { red_v[threadIdx.x >> 5] = best; red_i[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < kSelThreads / 32; w++)
        if (red_v[w] > best || (red_v[w] == best && red_i[w] < bi)) { best = red_v[w]; bi = red_i[w]; }
      // fewer admissible continuations than beams (dictionary, first step): pad with the best one (model.lua:424-436)
      if (best == -INFINITY) bi = picked[0];
      picked[round] = bi;
      p.new_scores[(int64_t)b * K + round] = best == -INFINITY ? p.new_scores[(int64_t)b * K] : (double)best;
      if (best != -INFINITY) total[bi] = -INFINITY;      // taken
    }
    __syncthreads();
  }
  if (threadIdx.x < K) {
    const int kk = threadIdx.x, j = picked[kk];
    const int pk = j / V, v = j % V;
    const int rnew = kk * Bc + b;
    p.tok_out[rnew] = v + 1;
    p.parent_row[rnew] = pk * Bc + b;                    // source row of the re-gather (model.lua:516,522-531)
    p.hist_tok[((int64_t)p.t * Bc + b) * K + kk] = v + 1;
    p.hist_par[((int64_t)p.t * Bc + b) * K + kk] = pk;
    if (p.trie) {                                        // model.lua:437-442,498-511
      const int node = p.t == 0 ? 0 : p.loc[(int64_t)b * K + pk];
      p.new_loc[(int64_t)b * K + kk] = (p.t > 0 && v == 0) ? node : p.trie[(int64_t)node * (V + 1) + v + 1];
    }
  }
}

// dst row r <- src row parent_row[r] of every listed tensor, 16 bytes per thread and step, through `tmp` (the
// permutation is not in place); second pass copies back
__global__ void __launch_bounds__(256) beam_gather_kernel(BeamGather g, int pass) {
  for (int ti = 0; ti < g.n; ti++) {
    const BeamGatherTensor t = g.t[ti];
    const int64_t chunks = t.row_bytes / 16;
    const int64_t total = (int64_t)g.rows * chunks;
    uint8_t* tmp = g.tmp + t.tmp_off;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
      const int64_t r = e / chunks, c = e % chunks;
      if (pass == 0) {
        const int64_t sr = g.parent_row[r];
        *reinterpret_cast<uint4*>(tmp + r * t.row_bytes + c * 16) =
            *reinterpret_cast<const uint4*>(t.ptr + sr * t.pitch_bytes + c * 16);
      } else {
        *reinterpret_cast<uint4*>(t.ptr + r * t.pitch_bytes + c * 16) =
            *reinterpret_cast<const uint4*>(tmp + r * t.row_bytes + c * 16);
      }
    }
  }
}

// best final beam (first maximum, model.lua:574-576) and the walk back through the parent history (:577-585)
__global__ void beam_backtrack_kernel(BeamBacktrack p) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= p.Bc) return;
  const int K = p.K;
  int idx = 0;
  double best = p.scores[(int64_t)b * K];
  for (int k = 1; k < K; k++)
    if (p.scores[(int64_t)b * K + k] > best) { best = p.scores[(int64_t)b * K + k]; idx = k; }
  p.score_out[b] = best;
  for (int t = p.L - 1; t >= 0; t--) {
    p.labels[(int64_t)b * p.ldl + t] = p.hist_tok[((int64_t)t * p.Bc + b) * K + idx];
    idx = p.hist_par[((int64_t)t * p.Bc + b) * K + idx];
  }
}

}  // namespace

void beam_select(Ctx& ctx, const BeamSelect& p) {
  AOCR_CHECK(p.K <= 64 && p.K * p.V <= 4096, "beam_select: beam * vocabulary too large");
  const size_t smem = (size_t)(p.t == 0 ? 1 : p.K) * p.V * sizeof(float);
  beam_select_kernel<<<p.Bc, kSelThreads, smem, ctx.st>>>(p);
  AOCR_LAUNCH_CHECK(ctx);
}
void beam_gather(Ctx& ctx, const BeamGather& g) {
  int64_t most = 0;
  for (int i = 0; i < g.n; i++) most = std::max<int64_t>(most, (int64_t)g.rows * (g.t[i].row_bytes / 16));
  const int grid = (int)std::min<int64_t>((most + 255) / 256, (int64_t)ctx.num_sms * 4);
  for (int pass = 0; pass < 2; pass++) {
    beam_gather_kernel<<<grid > 0 ? grid : 1, 256, 0, ctx.st>>>(g, pass);
    AOCR_LAUNCH_CHECK(ctx);
  }
}
void beam_backtrack(Ctx& ctx, const BeamBacktrack& p) {
  beam_backtrack_kernel<<<cdiv(p.Bc, 128), 128, 0, ctx.st>>>(p);
  AOCR_LAUNCH_CHECK(ctx);
}

}  // namespace aocr
