// gemm_simt.cu — fp32 FFMA GEMM with arbitrary operand strides.
// Role: (1) the thin contractions tensor cores cannot help (generator N=39, embedding K=20,
// per-row attention-gradient products), (2) on-device cross-check of the tcgen05 path
// (aocr_config.gemm_mode = 2).  64x64x16 CTA tile, 256 threads, 4x4 register tile.
#include "common.cuh"

namespace aocr {

namespace {
constexpr int TM = 64, TN = 64, TK = 16;

__global__ void __launch_bounds__(256) gemm_simt_kernel(Gemm g) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const float* A = g.A + (int64_t)blockIdx.z * g.bsa;
  const float* B = g.B + (int64_t)blockIdx.z * g.bsb;
  float* C = g.C + (int64_t)blockIdx.z * g.bsc;
  const int tx = tid % 16, ty = tid / 16;   // 16x16 threads, each a 4x4 micro tile
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

  const bool a_kfast = (g.sak == 1);
  const bool b_nfast = (g.sbn == 1);
  for (int k0 = 0; k0 < g.K; k0 += TK) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      int e = tid + i * 256;
      int m, k;
      if (a_kfast) { k = e % TK; m = e / TK; } else { m = e % TM; k = e / TM; }
      int gm = m0 + m, gk = k0 + k;
      As[k][m] = (gm < g.M && gk < g.K) ? A[(int64_t)gm * g.sam + (int64_t)gk * g.sak] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
      int e = tid + i * 256;
      int n, k;
      if (b_nfast) { n = e % TN; k = e / TN; } else { k = e % TK; n = e / TK; }
      int gn = n0 + n, gk = k0 + k;
      Bs[k][n] = (gn < g.N && gk < g.K) ? B[(int64_t)gk * g.sbk + (int64_t)gn * g.sbn] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; k++) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; i++) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; j++) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    int m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      int n = n0 + tx * 4 + j;
      if (n >= g.N) continue;
      float v = acc[i][j];
      if (g.bias_n) v += g.bias_n[n];
      if (g.bias_m) v += g.bias_m[m];
      if (g.act == ACT_TANH) v = tanhf(v);
      float* p = C + (int64_t)m * g.ldc + n;
      if (g.accumulate) v += *p;
      *p = v;
    }
  }
}
}  // namespace

void gemm_simt(Ctx& ctx, const Gemm& g) {
  if (g.M <= 0 || g.N <= 0) return;
  dim3 grid(cdiv(g.N, TN), cdiv(g.M, TM), g.batch);
  launch_pdl(ctx, gemm_simt_kernel, dim3(grid), dim3(256), 0, g);
  AOCR_CUDA(cudaGetLastError());
}

}  // namespace aocr
