// engine_gemm.cu — routes every contraction of the step to the tcgen05 GEMM core (gemm_tc.cu) or, for the
// thin shapes tensor cores cannot help (K < 64: generator K=39, embedding K=20, attention-gradient K=T),
// to the fp32 FFMA kernel.  Operands are converted fp32 -> (hi, lo) bf16 planes on the way in; parameter
// operands are converted once per weight update and cached.
#include "engine.h"

namespace aocr {

// parameters, and the flipped/transposed conv weights derived from them in prep_weights()
bool Engine::is_param(const float* p) const {
  if (p >= d_params && p < d_params + L.total) return true;
  for (int l = 1; l < 7; l++) {
    const int64_t n = (int64_t)kConv[l].cout * kConv[l].cin * kConv[l].k * kConv[l].k;
    if (wt[l] && p >= wt[l] && p < wt[l] + n) return true;
  }
  return false;
}

// operand (rows x K, strides srs/sks) as a K-major pack: cached for parameters, scratch slot otherwise
Pack Engine::operand_pack(const float* ptr, int64_t rows, int64_t K, int64_t srs, int64_t sks, int slot) {
  const int64_t kp = pad64(K);
  if (is_param(ptr)) {
    auto key = std::make_tuple(ptr, rows, K, srs, sks);
    auto it = wcache_.find(key);
    if (it == wcache_.end()) {
      WeightPack wp;
      wp.pack = alloc_pack(rows, kp);
      wp.version = -1;
      it = wcache_.emplace(key, wp).first;
    }
    if (it->second.version != weights_version_) {
      split_to_pack(ctx_, ptr, rows, K, srs, sks, it->second.pack);
      it->second.version = weights_version_;
    }
    return it->second.pack;
  }
  AOCR_CHECK(rows * kp <= scratch_elems_, "operand larger than the scratch pack (internal sizing error)");
  Pack p;
  p.rows = rows; p.kp = kp; p.hi = scratch_[slot].hi; p.lo = scratch_[slot].lo;
  split_to_pack(ctx_, ptr, rows, K, srs, sks, p);
  return p;
}

void Engine::gemm(const Gemm& g, int cls) {
  prof_begin(cls);
  const bool tc = cfg.gemm_mode != 2 && g.batch == 1 && g.K >= 64 && (g.M >= 64 || g.N >= 64);
  if (!tc) {
    gemm_simt(ctx_, g);
  } else if (g.sam == 1 && g.sbn == 1 && g.sak >= g.M && g.sbk >= g.N) {
    // dW = dY^T X with dY (K x M) and X (K x N) stored row-major: both operands are MN-major as they lie in
    // memory - no transposing conversion, the contraction index K (time*batch) is the TMA row coordinate
    TcGemm t;
    Pack pa, pb;
    if (g.pa) {
      pa = *g.pa; pa.rows = g.K;
    } else {
      pa.rows = g.K; pa.kp = pad64(g.M); pa.hi = scratch_[0].hi; pa.lo = scratch_[0].lo;
      AOCR_CHECK(pa.rows * pa.kp <= scratch_elems_, "wgrad operand exceeds scratch");
      split_to_pack(ctx_, g.A, g.K, g.M, g.sak, 1, pa);
    }
    if (g.pb) {
      pb = *g.pb; pb.rows = g.K;
    } else {
      pb.rows = g.K; pb.kp = pad64(g.N); pb.hi = scratch_[1].hi; pb.lo = scratch_[1].lo;
      AOCR_CHECK(pb.rows * pb.kp <= scratch_elems_, "wgrad operand exceeds scratch");
      split_to_pack(ctx_, g.B, g.K, g.N, g.sbk, 1, pb);
    }
    t.mn = 1; t.K = g.K; t.C = g.C; t.ldc = g.ldc; t.act = g.act; t.accumulate = g.accumulate;
    t.terms = cfg.gemm_mode == 1 ? 1 : 3;
    if (g.M >= g.N) {
      t.A = pa; t.B = pb; t.M = g.M; t.N = g.N; t.transpose_out = false; t.bias_m = g.bias_m; t.bias_n = g.bias_n;
    } else {
      t.A = pb; t.B = pa; t.M = g.N; t.N = g.M; t.transpose_out = true; t.bias_m = g.bias_n; t.bias_n = g.bias_m;
    }
    gemm_tc(ctx_, t);
  } else {
    TcGemm t;
    const bool swap = g.M < g.N;   // the larger side becomes the 128-row M side of the UMMA tile
    Pack pa = g.pa ? *g.pa : operand_pack(g.A, g.M, g.K, g.sam, g.sak, 0);  // rows m (K-major planes may be given)
    if (g.pa) pa.rows = g.M;
    Pack pb = operand_pack(g.B, g.N, g.K, g.sbn, g.sbk, 1);                 // rows n
    t.K = g.K; t.C = g.C; t.ldc = g.ldc; t.act = g.act; t.accumulate = g.accumulate;
    t.terms = cfg.gemm_mode == 1 ? 1 : 3;
    if (!swap) {
      t.A = pa; t.B = pb; t.M = g.M; t.N = g.N; t.transpose_out = false;
      t.bias_m = g.bias_m; t.bias_n = g.bias_n;
    } else {
      t.A = pb; t.B = pa; t.M = g.N; t.N = g.M; t.transpose_out = true;
      t.bias_m = g.bias_n; t.bias_n = g.bias_m;
    }
    gemm_tc(ctx_, t);
  }
  prof_end(cls, 2.0 * g.M * g.N * (double)g.K * g.batch);
}

// Convolution as implicit GEMM on the tcgen05 core: out (N,Ho,Wo,Cout) = conv(x (N,H,W,C), Wk [Cout][k*k*C]) + bias.
// Used for the forward convolutions (x = activation, Wk = parameters) and for the data gradient (x = dz,
// Wk = flipped / in-out-swapped weights, pad = k-1-pad).  Only the small NHWC activation is converted to bf16 planes.
void Engine::conv_tc(const float* x, int N, int H, int W, int C, int k, int pad, int Ho, int Wo, const float* Wk, int Cout,
                     float* out, const float* bias, const Pack* xpack) {
  const int64_t rows_in = (int64_t)N * H * W;
  const int Kc = k * k * C;
  prof_begin(0);
  Pack xp;
  AOCR_CHECK(C % 64 == 0, "conv_tc: channel count must be a multiple of 64");
  if (xpack) {          // the producer of x already wrote it as bf16 planes
    xp = *xpack; xp.rows = rows_in; xp.kp = C;
  } else {
    xp.rows = rows_in; xp.kp = C; xp.hi = scratch_[0].hi; xp.lo = scratch_[0].lo;
    AOCR_CHECK(rows_in * C <= scratch_elems_, "conv_tc: activation larger than the scratch pack");
    split_to_pack(ctx_, x, rows_in, C, C, 1, xp);
  }
  Pack wp = operand_pack(Wk, Cout, Kc, Kc, 1, 1);
  ConvView v;
  v.N = N; v.H = H; v.W = W; v.C = C; v.k = k; v.pad = pad; v.Ho = Ho; v.Wo = Wo;
  TcGemm t;
  t.A = xp; t.B = wp; t.conv = &v;
  t.M = N * Ho * Wo; t.N = Cout; t.K = Kc;
  t.C = out; t.ldc = Cout; t.bias_n = bias;
  t.terms = cfg.gemm_mode == 1 ? 1 : 3;
  gemm_tc(ctx_, t);
  prof_end(0, 2.0 * N * Ho * Wo * (double)Cout * Kc);
}

// Convolution weight gradient as an implicit GEMM with MN-major operands: dW[co][tap*Cin+ci] =
// sum over output pixels of dz[pix][co] * x[pix + tap][ci].  Both operands are the NHWC tensors as stored; the
// filter tap is a shift of the TMA box origin of x (zero fill = padding).  No im2col, no transposes.
void Engine::conv_wgrad_tc(const float* dz, const float* x, int N, int H, int W, int Cin, int k, int pad, int Ho, int Wo,
                           int Cout, float* dW, const Pack* zpack, const Pack* xpack) {
  const int64_t rows_out = (int64_t)N * Ho * Wo, rows_in = (int64_t)N * H * W;
  prof_begin(0);
  Pack zp, xp;
  zp.rows = rows_out; zp.kp = Cout; zp.hi = scratch_[0].hi; zp.lo = scratch_[0].lo;
  xp.rows = rows_in; xp.kp = Cin; xp.hi = scratch_[1].hi; xp.lo = scratch_[1].lo;
  AOCR_CHECK(rows_out * Cout <= scratch_elems_ && rows_in * Cin <= scratch_elems_, "conv wgrad operand exceeds scratch");
  if (zpack) { zp.hi = zpack->hi; zp.lo = zpack->lo; } else split_to_pack(ctx_, dz, rows_out, Cout, Cout, 1, zp);
  if (xpack) { xp.hi = xpack->hi; xp.lo = xpack->lo; } else split_to_pack(ctx_, x, rows_in, Cin, Cin, 1, xp);
  ConvView v;
  v.N = N; v.H = H; v.W = W; v.C = Cin; v.k = k; v.pad = pad; v.Ho = Ho; v.Wo = Wo;
  TcGemm t;
  t.mn = 2; t.conv = &v; t.A = zp; t.B = xp;
  t.M = Cout; t.N = k * k * Cin; t.K = (int)rows_out;
  t.C = dW; t.ldc = k * k * Cin;
  t.terms = cfg.gemm_mode == 1 ? 1 : 3;
  gemm_tc(ctx_, t);
  prof_end(0, 2.0 * rows_out * (double)Cout * k * k * Cin);
}

}  // namespace aocr

// ---------------------------------------------------------------------------------------------
// stand-alone GEMM self test (no handle): C = A(MxK) * B(KxN), row-major fp32 host buffers.
// mode 0: tcgen05 bf16x3, 1: tcgen05 bf16, 2: fp32 SIMT.  ta/tb: operand stored transposed.
extern "C" int aocr_selftest_gemm(int M, int N, int K, int ta, int tb, int mode, int swap, const float* A,
                                  const float* B, float* C, char* err, int errlen, int splits) {
  using namespace aocr;
  float *dA = nullptr, *dB = nullptr, *dC = nullptr;
  __nv_bfloat16* planes[4] = {nullptr, nullptr, nullptr, nullptr};
  Ctx ctx;
  int rc = 0;
  try {
    AOCR_CUDA(cudaStreamCreate(&ctx.st));
    ctx.tc_ws_floats = (int64_t)32 * 148 * 128 * 128;
    AOCR_CUDA(cudaMalloc(&ctx.tc_ws, ctx.tc_ws_floats * 4));
    ctx.tc_counters_n = 4096;
    AOCR_CUDA(cudaMalloc(&ctx.tc_counters, 4096 * 4));
    AOCR_CUDA(cudaMemset(ctx.tc_counters, 0, 4096 * 4));
    AOCR_CUDA(cudaMalloc(&dA, (size_t)M * K * 4));
    AOCR_CUDA(cudaMalloc(&dB, (size_t)K * N * 4));
    AOCR_CUDA(cudaMalloc(&dC, (size_t)M * N * 4));
    AOCR_CUDA(cudaMemcpy(dA, A, (size_t)M * K * 4, cudaMemcpyHostToDevice));
    AOCR_CUDA(cudaMemcpy(dB, B, (size_t)K * N * 4, cudaMemcpyHostToDevice));
    AOCR_CUDA(cudaMemset(dC, 0, (size_t)M * N * 4));
    // element strides of the logical operands inside the stored buffers
    int64_t sam = ta ? 1 : K, sak = ta ? M : 1;       // A(m,k)
    int64_t sbk = tb ? 1 : N, sbn = tb ? K : 1;       // B(k,n)
    if (mode == 2) {
      Gemm g;
      g.M = M; g.N = N; g.K = K; g.A = dA; g.sam = sam; g.sak = sak; g.B = dB; g.sbk = sbk; g.sbn = sbn;
      g.C = dC; g.ldc = N;
      gemm_simt(ctx, g);
    } else {
      int64_t kp = pad64(K);
      AOCR_CUDA(cudaMalloc(&planes[0], (size_t)M * kp * 2));
      AOCR_CUDA(cudaMalloc(&planes[1], (size_t)M * kp * 2));
      AOCR_CUDA(cudaMalloc(&planes[2], (size_t)N * kp * 2));
      AOCR_CUDA(cudaMalloc(&planes[3], (size_t)N * kp * 2));
      Pack pa, pb;
      pa.hi = planes[0]; pa.lo = planes[1]; pa.rows = M; pa.kp = kp;
      pb.hi = planes[2]; pb.lo = planes[3]; pb.rows = N; pb.kp = kp;
      TcGemm t;
      t.K = K; t.C = dC; t.ldc = N; t.terms = mode == 1 ? 1 : 3; t.force_splits = splits;
      if (swap == 2) {
        // MN-major test: operands as [K][M] and [K][N] planes (pitch padded to 64), no transposition
        int64_t mp = pad64(M), np_ = pad64(N);
        for (auto& q : planes) { cudaFree(q); q = nullptr; }
        AOCR_CUDA(cudaMalloc(&planes[0], (size_t)K * mp * 2)); AOCR_CUDA(cudaMalloc(&planes[1], (size_t)K * mp * 2));
        AOCR_CUDA(cudaMalloc(&planes[2], (size_t)K * np_ * 2)); AOCR_CUDA(cudaMalloc(&planes[3], (size_t)K * np_ * 2));
        Pack ka, kb;
        ka.hi = planes[0]; ka.lo = planes[1]; ka.rows = K; ka.kp = mp;
        kb.hi = planes[2]; kb.lo = planes[3]; kb.rows = K; kb.kp = np_;
        split_to_pack(ctx, dA, K, M, sak, sam, ka);     // rows = k, cols = m
        split_to_pack(ctx, dB, K, N, sbk, sbn, kb);
        t.A = ka; t.B = kb; t.M = M; t.N = N; t.mn = 1; t.transpose_out = false;
      } else {
        split_to_pack(ctx, dA, M, K, sam, sak, pa);
        split_to_pack(ctx, dB, N, K, sbn, sbk, pb);
        if (!swap) { t.A = pa; t.B = pb; t.M = M; t.N = N; t.transpose_out = false; }
        else { t.A = pb; t.B = pa; t.M = N; t.N = M; t.transpose_out = true; }
      }
      gemm_tc(ctx, t);
    }
    AOCR_CUDA(cudaStreamSynchronize(ctx.st));
    AOCR_CUDA(cudaMemcpy(C, dC, (size_t)M * N * 4, cudaMemcpyDeviceToHost));
  } catch (const std::exception& e) {
    if (err && errlen > 0) { strncpy(err, e.what(), errlen - 1); err[errlen - 1] = 0; }
    rc = -2;
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(ctx.tc_ws); cudaFree(ctx.tc_counters);
  for (auto p : planes) cudaFree(p);
  if (ctx.st) cudaStreamDestroy(ctx.st);
  return rc;
}

// micro-benchmark of the tcgen05 GEMM core on device-resident random operands (warm L2, CUDA events):
// returns the average microseconds per call.  layout: 0 = K-major (N side = batch when swap), 1 = MN-major.
extern "C" int aocr_bench_gemm(int M, int N, int K, int terms, int mn, int splits, int reps, float* us_out) {
  using namespace aocr;
  Ctx ctx;
  __nv_bfloat16* pl[4] = {nullptr, nullptr, nullptr, nullptr};
  float* dC = nullptr;
  int rc = 0;
  try {
    AOCR_CUDA(cudaStreamCreate(&ctx.st));
    cudaDeviceProp prop; AOCR_CUDA(cudaGetDeviceProperties(&prop, 0)); ctx.num_sms = prop.multiProcessorCount;
    ctx.tc_ws_floats = (int64_t)32 * 148 * 128 * 128;
    AOCR_CUDA(cudaMalloc(&ctx.tc_ws, ctx.tc_ws_floats * 4));
    const int64_t kp = pad64(K), mp = pad64(M), np_ = pad64(N);
    const size_t ea = mn ? (size_t)K * mp : (size_t)M * kp, eb = mn ? (size_t)K * np_ : (size_t)N * kp;
    AOCR_CUDA(cudaMalloc(&pl[0], ea * 2)); AOCR_CUDA(cudaMalloc(&pl[1], ea * 2));
    AOCR_CUDA(cudaMalloc(&pl[2], eb * 2)); AOCR_CUDA(cudaMalloc(&pl[3], eb * 2));
    for (int i = 0; i < 4; i++) AOCR_CUDA(cudaMemset(pl[i], 0x11, (i < 2 ? ea : eb) * 2));
    AOCR_CUDA(cudaMalloc(&dC, (size_t)M * N * 4));
    TcGemm t;
    t.A.hi = pl[0]; t.A.lo = pl[1]; t.B.hi = pl[2]; t.B.lo = pl[3];
    if (mn) { t.A.rows = K; t.A.kp = mp; t.B.rows = K; t.B.kp = np_; t.mn = 1; }
    else { t.A.rows = M; t.A.kp = kp; t.B.rows = N; t.B.kp = kp; }
    t.M = M; t.N = N; t.K = K; t.C = dC; t.ldc = N; t.terms = terms; t.force_splits = splits;
    t.defer_reduce = true; t.ws = ctx.tc_ws; t.ws_floats = ctx.tc_ws_floats;
    if (const char* e = getenv("AOCR_TC_DBG")) t.dbg = atoi(e);
    t.transpose_out = getenv("AOCR_TC_TRANSPOSE") != nullptr;
    if (t.transpose_out) t.ldc = M;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; i++) gemm_tc(ctx, t);
    AOCR_CUDA(cudaEventRecord(e0, ctx.st));
    for (int i = 0; i < reps; i++) gemm_tc(ctx, t);
    AOCR_CUDA(cudaEventRecord(e1, ctx.st));
    AOCR_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    *us_out = ms * 1000.f / reps;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
  } catch (const std::exception& e) {
    fprintf(stderr, "aocr_bench_gemm: %s\n", e.what());
    rc = -2;
  }
  for (auto q : pl) cudaFree(q);
  cudaFree(dC); cudaFree(ctx.tc_ws);
  if (ctx.st) cudaStreamDestroy(ctx.st);
  return rc;
}
