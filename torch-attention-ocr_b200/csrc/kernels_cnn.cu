// kernels_cnn.cu — memory-bound pieces of the CNN (reference: src/model/cnn.lua:9-45):
// conv1 (K=9, too thin for tensor cores), ReLU+max-pool, batch-norm, im2col, column reductions.
// All activations are NHWC fp32 (channel fastest) so every warp access is a contiguous run of channels.
#include "kernels.h"

namespace aocr {

namespace {

constexpr float BN_EPS = 1e-5f;

// ------------------------------------------------------------------ conv1
// thread = one pooled output element (n, ph, pw, c); the 4x4 input patch is shared by the 64 channel
// threads (L1 broadcast), weights live in smem.
// optional bf16 (hi, lo) operand planes written next to an fp32 activation (same linear index: kp == C)
__device__ __forceinline__ void split_store1(__nv_bfloat16* phi, __nv_bfloat16* plo, int64_t i, float v) {
  if (!phi) return;
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  phi[i] = h;
  plo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
}
__device__ __forceinline__ void split_store4(__nv_bfloat16* phi, __nv_bfloat16* plo, int64_t i, float4 v) {
  if (!phi) return;
  __nv_bfloat16 h[4], l[4];
  const float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int k = 0; k < 4; k++) {
    h[k] = __float2bfloat16_rn(x[k]);
    l[k] = __float2bfloat16_rn(x[k] - __bfloat162float(h[k]));
  }
  *reinterpret_cast<uint2*>(phi + i) = *reinterpret_cast<uint2*>(h);
  *reinterpret_cast<uint2*>(plo + i) = *reinterpret_cast<uint2*>(l);
}

// One thread = one pooled pixel x 16 output channels: the 4 x 4 input patch is loaded once (not once per channel), the
// four conv outputs of the pooling window are formed per channel from weights in shared memory, and the 16 results leave
// as 16-byte stores (fp32, the two bf16 planes, the pooling choice).  The four threads of a pixel are adjacent.
__global__ void __launch_bounds__(256) conv1_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ bias, float* __restrict__ a1,
                                                        uint8_t* __restrict__ idx, int B, int W, __nv_bfloat16* __restrict__ phi,
                                                        __nv_bfloat16* __restrict__ plo) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float ws[9][68];          // [tap][channel group * 17 + k]: the four groups of a warp hit distinct banks
  __shared__ float bs[64];
  for (int i = threadIdx.x; i < 576; i += blockDim.x) ws[i % 9][((i / 9) >> 4) * 17 + ((i / 9) & 15)] = w[i];
  if (threadIdx.x < 64) bs[threadIdx.x] = bias[threadIdx.x];
  __syncthreads();
  const int W1 = W / 2;
  const int64_t total = (int64_t)B * 16 * W1 * 4;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int cg = (int)(t & 3);
    int64_t r = t >> 2;
    const int pw = (int)(r % W1); r /= W1;
    const int ph = (int)(r % 16);
    const int n = (int)(r / 16);
    const float* xi = x + (int64_t)n * 32 * W;
    float p[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int h = 2 * ph - 1 + i;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int ww = 2 * pw - 1 + j;
        float v = 0.f;
        if (h >= 0 && h < 32 && ww >= 0 && ww < W) v = (__ldg(xi + h * W + ww) - 128.0f) * (1.0f / 128.0f);
        p[i][j] = v;
      }
    }
    const int64_t e0 = (t >> 2) * 64 + cg * 16;            // first of this thread's 16 consecutive output elements
    float outv[16];
    uint32_t packed_idx[4] = {0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 16; k++) {
      const int c = cg * 16 + k;
      float best = -1.f;
      int bi = 0;
#pragma unroll
      for (int dy = 0; dy < 2; dy++)
#pragma unroll
        for (int dx = 0; dx < 2; dx++) {
          float sacc = bs[c];
#pragma unroll
          for (int kh = 0; kh < 3; kh++)
#pragma unroll
            for (int kw = 0; kw < 3; kw++) sacc = fmaf(ws[kh * 3 + kw][cg * 17 + k], p[dy + kh][dx + kw], sacc);
          sacc = fmaxf(sacc, 0.f);
          if (sacc > best) { best = sacc; bi = dy * 2 + dx; }
        }
      outv[k] = best;
      packed_idx[k >> 2] |= (uint32_t)bi << (8 * (k & 3));
    }
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const float4 v = make_float4(outv[4 * q], outv[4 * q + 1], outv[4 * q + 2], outv[4 * q + 3]);
      *reinterpret_cast<float4*>(a1 + e0 + 4 * q) = v;
      split_store4(phi, plo, e0 + 4 * q, v);
    }
    *reinterpret_cast<uint4*>(idx + e0) = make_uint4(packed_idx[0], packed_idx[1], packed_idx[2], packed_idx[3]);
  }
}

// dW1[c][tap] = sum over pooled elements with a1>0 of da1 * xnorm(argmax position + tap); db1[c] = sum da1*(a1>0).
// block = 64 channel lanes x 4 pixel lanes; each block reduces a contiguous slab of pooled pixels, writes a
// (10 x 64) partial; a second kernel folds the partials in fixed order (deterministic).
__global__ void __launch_bounds__(256) conv1_bwd_partial_kernel(const float* __restrict__ x, const float* __restrict__ a1,
                                                                const uint8_t* __restrict__ idx,
                                                                const float* __restrict__ da1, float* __restrict__ partial,
                                                                int B, int W, int64_t pix_per_blk) {
  pdl_launch_dependents();
  pdl_wait();
  const int c = threadIdx.x % 64, lane = threadIdx.x / 64;   // 4 pixel lanes
  const int W1 = W / 2;
  const int64_t npix = (int64_t)B * 16 * W1;
  int64_t p0 = (int64_t)blockIdx.x * pix_per_blk;
  int64_t p1 = p0 + pix_per_blk < npix ? p0 + pix_per_blk : npix;
  float acc[10];
#pragma unroll
  for (int i = 0; i < 10; i++) acc[i] = 0.f;
  for (int64_t p = p0 + lane; p < p1; p += 4) {
    float g = da1[p * 64 + c];
    float a = a1[p * 64 + c];
    if (!(a > 0.f) || g == 0.f) continue;
    int bi = idx[p * 64 + c];
    int64_t r = p;
    int pw = (int)(r % W1); r /= W1;
    int ph = (int)(r % 16);
    int n = (int)(r / 16);
    int h0 = 2 * ph + (bi >> 1) - 1, w0 = 2 * pw + (bi & 1) - 1;
    const float* xi = x + (int64_t)n * 32 * W;
#pragma unroll
    for (int kh = 0; kh < 3; kh++)
#pragma unroll
      for (int kw = 0; kw < 3; kw++) {
        int h = h0 + kh, ww = w0 + kw;
        float v = 0.f;
        if (h >= 0 && h < 32 && ww >= 0 && ww < W) v = (xi[h * W + ww] - 128.0f) * (1.0f / 128.0f);
        acc[kh * 3 + kw] = fmaf(g, v, acc[kh * 3 + kw]);
      }
    acc[9] += g;
  }
  __shared__ float red[4][10][64];
#pragma unroll
  for (int i = 0; i < 10; i++) red[lane][i][c] = acc[i];
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 10; i++)
      partial[((int64_t)blockIdx.x * 10 + i) * 64 + c] = red[0][i][c] + red[1][i][c] + red[2][i][c] + red[3][i][c];
  }
}
// folds the per-block partials in a fixed order: block i (0..9: nine taps + bias), 64 channel lanes x 16 block lanes,
// each thread sums every 16th partial (coalesced over channels), then the 16 lane sums are added in order
__global__ void __launch_bounds__(1024) conv1_bwd_final_kernel(const float* __restrict__ partial, int nblk, float* __restrict__ dw,
                                                               float* __restrict__ db) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[16][64];
  const int i = blockIdx.x, c = threadIdx.x % 64, lane = threadIdx.x / 64;
  float s = 0.f;
  for (int b = lane; b < nblk; b += 16) s += partial[((int64_t)b * 10 + i) * 64 + c];
  red[lane][c] = s;
  __syncthreads();
  if (lane == 0) {
    float t = 0.f;
#pragma unroll
    for (int l = 0; l < 16; l++) t += red[l][c];
    if (i < 9) dw[c * 9 + i] += t; else db[c] += t;
  }
}

// ------------------------------------------------------------------ ReLU + max-pool
template <int KW>
__global__ void __launch_bounds__(256) relu_pool_fwd_kernel(const float* __restrict__ z, float* __restrict__ a,
                                                            uint8_t* __restrict__ idx, int B, int H, int Wi, int C,
                                                            __nv_bfloat16* __restrict__ phi, __nv_bfloat16* __restrict__ plo) {
  pdl_launch_dependents();
  pdl_wait();
  const int Ho = H / 2, Wo = Wi / KW;
  const int64_t total = (int64_t)B * Ho * Wo * C;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(e % C);
    int64_t r = e / C;
    int pw = (int)(r % Wo); r /= Wo;
    int ph = (int)(r % Ho);
    int n = (int)(r / Ho);
    float best = -1.f;
    int bi = 0;
#pragma unroll
    for (int dy = 0; dy < 2; dy++)
#pragma unroll
      for (int dx = 0; dx < KW; dx++) {
        float v = z[(((int64_t)n * H + 2 * ph + dy) * Wi + KW * pw + dx) * C + c];
        v = fmaxf(v, 0.f);
        if (v > best) { best = v; bi = dy * 2 + dx; }
      }
    a[e] = best;
    idx[e] = (uint8_t)bi;
    split_store1(phi, plo, e, best);
  }
}

// one thread per *input* (pre-pool) element so dz is written exactly once, coalesced, incl. zeros and
// the floor-mode leftover column.
template <int KW>
__global__ void __launch_bounds__(256) relu_pool_bwd_kernel(const float* __restrict__ da, const float* __restrict__ a,
                                                            const uint8_t* __restrict__ idx, float* __restrict__ dz,
                                                            int B, int H, int Wi, int C, __nv_bfloat16* __restrict__ phi,
                                                            __nv_bfloat16* __restrict__ plo) {
  pdl_launch_dependents();
  pdl_wait();
  const int Ho = H / 2, Wo = Wi / KW;
  const int64_t total = (int64_t)B * H * Wi * C;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(e % C);
    int64_t r = e / C;
    int w = (int)(r % Wi); r /= Wi;
    int h = (int)(r % H);
    int n = (int)(r / H);
    int ph = h / 2, pw = w / KW;
    float v = 0.f;
    if (ph < Ho && pw < Wo) {
      int64_t o = (((int64_t)n * Ho + ph) * Wo + pw) * C + c;
      int bi = (h & 1) * 2 + (KW == 2 ? (w & 1) : 0);
      if (idx[o] == bi && a[o] > 0.f) v = da[o];
    }
    dz[e] = v;
    split_store1(phi, plo, e, v);
  }
}

// ------------------------------------------------------------------ im2col
__global__ void __launch_bounds__(256) im2col_kernel(const float* __restrict__ a, float* __restrict__ col, int B, int H,
                                                     int Wi, int C, int k, int pad, int Ho, int Wo) {
  pdl_launch_dependents();
  pdl_wait();
  const int C4 = C / 4;
  const int64_t total = (int64_t)B * Ho * Wo * k * k * C4;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int c4 = (int)(e % C4);
    int64_t r = e / C4;
    int tap = (int)(r % (k * k)); r /= (k * k);
    int wo = (int)(r % Wo); r /= Wo;
    int ho = (int)(r % Ho);
    int n = (int)(r / Ho);
    int kh = tap / k, kw = tap % k;
    int h = ho + kh - pad, w = wo + kw - pad;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (h >= 0 && h < H && w >= 0 && w < Wi)
      v = *reinterpret_cast<const float4*>(a + (((int64_t)n * H + h) * Wi + w) * C + c4 * 4);
    *reinterpret_cast<float4*>(col + e * 4) = v;
  }
}

// ------------------------------------------------------------------ column reductions over (R, C)
// Two-stage reductions in ONE launch: every slab block writes its partial sums, the block that arrives LAST at the
// counter of its column block (blockIdx.x) folds the partials in slab order 0..n-1 - the order, and therefore the bits,
// of the former second kernel - and resets the counter for the next launch.
__device__ __forceinline__ unsigned* slab_counters(float* partial) {
  return reinterpret_cast<unsigned*>(partial + kPartialFloats - kSlabCounters);
}
__device__ __forceinline__ bool last_slab_block(float* partial) {
  __shared__ int s_last;
  __threadfence();                 // this block's partials are visible device-wide before it is counted
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned* ctr = slab_counters(partial) + blockIdx.x;
    const unsigned t = atomicAdd(ctr, 1u);
    s_last = (t == gridDim.y - 1);
    if (s_last) *ctr = 0;          // everyone has passed: ready for the next launch
  }
  __syncthreads();
  if (s_last) __threadfence();
  return s_last != 0;
}
// mode 0: sum z ; mode 1: sum (z-mean)^2 ; mode 2: sum z*xhat where xhat=(z2-mean)*inv  (z = dy)
// grid (C/32, nslab); block 32 x 8: each thread walks rows r = slab*rows_per + ty, += 8
__global__ void __launch_bounds__(256) col_reduce_kernel(const float* __restrict__ z, const float* __restrict__ z2,
                                                         const float* __restrict__ mean, const float* __restrict__ var,
                                                         int64_t R, int C, int64_t rows_per, int mode,
                                                         float* __restrict__ partial, float scale, int accumulate,
                                                         float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
  const int c = blockIdx.x * 32 + tx;
  int64_t r0 = (int64_t)blockIdx.y * rows_per;
  int64_t r1 = r0 + rows_per < R ? r0 + rows_per : R;
  float acc = 0.f;
  if (c < C) {
    float mu = (mode >= 1) ? mean[c] : 0.f;
    float inv = (mode == 2) ? rsqrtf(var[c] + BN_EPS) : 0.f;
    for (int64_t r = r0 + ty; r < r1; r += 8) {
      float v = z[r * C + c];
      if (mode == 0) acc += v;
      else if (mode == 1) { float d = v - mu; acc = fmaf(d, d, acc); }
      else acc = fmaf(v, (z2[r * C + c] - mu) * inv, acc);
    }
  }
  __shared__ float red[8][33];
  red[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && c < C) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) s += red[i][tx];
    partial[(int64_t)blockIdx.y * C + c] = s;
  }
  if (!last_slab_block(partial)) return;
  if (ty == 0 && c < C) {          // out[c] (op)= scale * sum over slabs, fixed order
    float s = 0.f;
    for (int i = 0; i < (int)gridDim.y; i++) s += __ldcg(partial + (int64_t)i * C + c);
    s *= scale;
    out[c] = accumulate ? out[c] + s : s;
  }
}
// batch-norm forward statistics in ONE pass: s1 = sum (z - k), s2 = sum (z - k)^2 with the shift k = running mean
// (identical on every data-parallel rank, and close to the batch mean after a few steps: no cancellation in
// s2/R - (s1/R)^2).  partial: [2][nslab][C]
// single device (fin != 0): the last block also forms mean / biased variance and updates the running statistics for its
// columns (the work of bn_finalize_kernel); under data parallelism it writes the local sums for the exchange.
__global__ void __launch_bounds__(256) col_reduce_bn_fwd_kernel(const float* __restrict__ z, const float* __restrict__ shift,
                                                                int64_t R, int C, int64_t rows_per, float* __restrict__ partial,
                                                                float* __restrict__ sums, int fin, float invR, float unbias,
                                                                float* __restrict__ mean, float* __restrict__ var,
                                                                float* __restrict__ rmean, float* __restrict__ rvar) {
  pdl_launch_dependents();
  pdl_wait();
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
  const int c = blockIdx.x * 32 + tx;
  int64_t r0 = (int64_t)blockIdx.y * rows_per;
  int64_t r1 = r0 + rows_per < R ? r0 + rows_per : R;
  float a1 = 0.f, a2 = 0.f;
  if (c < C) {
    const float k = shift[c];
    for (int64_t r = r0 + ty; r < r1; r += 8) {
      const float d = z[r * C + c] - k;
      a1 += d;
      a2 = fmaf(d, d, a2);
    }
  }
  __shared__ float red[2][8][33];
  red[0][ty][tx] = a1; red[1][ty][tx] = a2;
  __syncthreads();
  if (ty < 2 && c < C) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) s += red[ty][i][tx];
    partial[((int64_t)ty * gridDim.y + blockIdx.y) * C + c] = s;
  }
  if (!last_slab_block(partial)) return;
  if (ty < 2 && c < C) {
    float s = 0.f;
    for (int i = 0; i < (int)gridDim.y; i++) s += __ldcg(partial + ((int64_t)ty * gridDim.y + i) * C + c);
    sums[ty * C + c] = s;
    red[ty][0][tx] = s;
  }
  if (!fin) return;
  __syncthreads();
  if (ty == 0 && c < C) {          // bn_finalize_kernel, for this block's 32 columns
    const float k = rmean[c];
    const float m1 = red[0][0][tx] * invR, m2 = red[1][0][tx] * invR;
    const float mu = k + m1;
    const float v = fmaxf(m2 - m1 * m1, 0.f);
    mean[c] = mu; var[c] = v;
    rmean[c] = 0.9f * k + 0.1f * mu;
    rvar[c] = 0.9f * rvar[c] + 0.1f * v * unbias;
  }
}
// mean / biased variance from the (globally summed) shifted sums, and the running-statistics update (momentum 0.1,
// unbiased variance) in the same kernel
__global__ void bn_finalize_kernel(const float* __restrict__ sums, int C, float invR, float unbias, float* __restrict__ mean,
                                   float* __restrict__ var, float* __restrict__ rmean, float* __restrict__ rvar) {
  pdl_launch_dependents();
  pdl_wait();
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float k = rmean[c];
  const float m1 = sums[c] * invR, m2 = sums[C + c] * invR;
  const float mu = k + m1;
  const float v = fmaxf(m2 - m1 * m1, 0.f);
  mean[c] = mu; var[c] = v;
  rmean[c] = 0.9f * k + 0.1f * mu;
  rvar[c] = 0.9f * rvar[c] + 0.1f * v * unbias;
}
// both batch-norm backward sums in one pass over dy and z: s1 = sum dy, s2 = sum dy * xhat.  partial: [2][nslab][C]
__global__ void __launch_bounds__(256) col_reduce_bn_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ z,
                                                                const float* __restrict__ mean, const float* __restrict__ var,
                                                                int64_t R, int C, int64_t rows_per, float* __restrict__ partial,
                                                                float* __restrict__ out1, float* __restrict__ out2) {
  pdl_launch_dependents();
  pdl_wait();
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
  const int c = blockIdx.x * 32 + tx;
  int64_t r0 = (int64_t)blockIdx.y * rows_per;
  int64_t r1 = r0 + rows_per < R ? r0 + rows_per : R;
  float a1 = 0.f, a2 = 0.f;
  if (c < C) {
    const float mu = mean[c], inv = rsqrtf(var[c] + BN_EPS);
    for (int64_t r = r0 + ty; r < r1; r += 8) {
      const float v = dy[r * C + c];
      a1 += v;
      a2 = fmaf(v, (z[r * C + c] - mu) * inv, a2);
    }
  }
  __shared__ float red[2][8][33];
  red[0][ty][tx] = a1; red[1][ty][tx] = a2;
  __syncthreads();
  if (ty < 2 && c < C) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) s += red[ty][i][tx];
    partial[((int64_t)ty * gridDim.y + blockIdx.y) * C + c] = s;
  }
  if (!last_slab_block(partial)) return;
  if (ty < 2 && c < C) {
    float s = 0.f;
    for (int i = 0; i < (int)gridDim.y; i++) s += __ldcg(partial + ((int64_t)ty * gridDim.y + i) * C + c);
    (ty ? out2 : out1)[c] = s;
  }
}
__global__ void bn_update_running_kernel(const float* __restrict__ mean, const float* __restrict__ var,
                                         float* __restrict__ rmean, float* __restrict__ rvar, int C, float unbias) {
  pdl_launch_dependents();
  pdl_wait();
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  rmean[c] = 0.9f * rmean[c] + 0.1f * mean[c];
  rvar[c] = 0.9f * rvar[c] + 0.1f * var[c] * unbias;
}

__global__ void __launch_bounds__(256) bn_relu_fwd_kernel(const float* __restrict__ z, const float* __restrict__ mean,
                                                          const float* __restrict__ var, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, float* __restrict__ a, int64_t R,
                                                          int C, int tm_S, int tm_B, __nv_bfloat16* __restrict__ phi,
                                                          __nv_bfloat16* __restrict__ plo) {
  pdl_launch_dependents();
  pdl_wait();
  const int C4 = C / 4;
  const int64_t total = R * C4;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(e % C4) * 4;
    int64_t r = e / C4;
    float4 v = *reinterpret_cast<const float4*>(z + r * C + c);
    float4 mu = *reinterpret_cast<const float4*>(mean + c);
    float4 va = *reinterpret_cast<const float4*>(var + c);
    float4 ga = *reinterpret_cast<const float4*>(gamma + c);
    float4 be = *reinterpret_cast<const float4*>(beta + c);
    float4 o;
    o.x = fmaxf(fmaf((v.x - mu.x) * rsqrtf(va.x + BN_EPS), ga.x, be.x), 0.f);
    o.y = fmaxf(fmaf((v.y - mu.y) * rsqrtf(va.y + BN_EPS), ga.y, be.y), 0.f);
    o.z = fmaxf(fmaf((v.z - mu.z) * rsqrtf(va.z + BN_EPS), ga.z, be.z), 0.f);
    o.w = fmaxf(fmaf((v.w - mu.w) * rsqrtf(va.w + BN_EPS), ga.w, be.w), 0.f);
    int64_t ro = r;
    if (tm_S > 0) ro = (r % tm_S) * tm_B + r / tm_S;
    *reinterpret_cast<float4*>(a + ro * C + c) = o;
    split_store4(phi, plo, ro * C + c, o);
  }
}

// dy = da * (a > 0), rows optionally read time-major
__global__ void __launch_bounds__(256) relu_mask_kernel(const float* __restrict__ da, const float* __restrict__ a,
                                                        float* __restrict__ dy, int64_t R, int C, int tm_S, int tm_B) {
  pdl_launch_dependents();
  pdl_wait();
  const int C4 = C / 4;
  const int64_t total = R * C4;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(e % C4) * 4;
    int64_t r = e / C4;
    int64_t ri = r;
    if (tm_S > 0) ri = (r % tm_S) * tm_B + r / tm_S;
    float4 g = *reinterpret_cast<const float4*>(da + ri * C + c);
    float4 v = *reinterpret_cast<const float4*>(a + ri * C + c);
    float4 o;
    o.x = v.x > 0.f ? g.x : 0.f;
    o.y = v.y > 0.f ? g.y : 0.f;
    o.z = v.z > 0.f ? g.z : 0.f;
    o.w = v.w > 0.f ? g.w : 0.f;
    *reinterpret_cast<float4*>(dy + r * C + c) = o;
  }
}

// dz = gamma*inv*(dy - s1/R - xhat*s2/R) in place on dz (=dy)
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(float* __restrict__ dz, const float* __restrict__ z,
                                                           const float* __restrict__ mean, const float* __restrict__ var,
                                                           const float* __restrict__ gamma, const float* __restrict__ s1,
                                                           const float* __restrict__ s2, int64_t R, int64_t Rglobal, int C,
                                                           int train, __nv_bfloat16* __restrict__ phi,
                                                           __nv_bfloat16* __restrict__ plo) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t total = R * C;                 // local rows
  const float invR = 1.0f / (float)Rglobal;    // statistics are over the global batch (data parallelism)
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(e % C);
    float inv = rsqrtf(var[c] + BN_EPS);
    float dy = dz[e];
    float o;
    if (train) {
      float xh = (z[e] - mean[c]) * inv;
      o = gamma[c] * inv * (dy - s1[c] * invR - xh * s2[c] * invR);
    } else {
      o = gamma[c] * inv * dy;
    }
    dz[e] = o;
    split_store1(phi, plo, e, o);
  }
}

inline int grid_for(int64_t total, int threads, int num_sms) {
  int64_t g = (total + threads - 1) / threads;
  int64_t cap = (int64_t)num_sms * 8;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

inline int nslabs_for(int64_t R, int num_sms, int C) {
  int cb = (C + 31) / 32;
  int want = (2 * num_sms + cb - 1) / cb;
  int64_t maxs = (R + 63) / 64;
  if (want > maxs) want = (int)maxs;
  if (want < 1) want = 1;
  if (want > 256) want = 256;
  return want;
}

void col_reduce(Ctx& ctx, const float* z, const float* z2, const float* mean, const float* var, int64_t R, int C,
                int mode, float scale, int accumulate, float* out, float* partial) {
  int ns = nslabs_for(R, ctx.num_sms, C);
  int64_t rows_per = (R + ns - 1) / ns;
  dim3 grid(cdiv(C, 32), ns);
  AOCR_CHECK(cdiv(C, 32) <= kSlabCounters && (int64_t)ns * C <= kPartialFloats - kSlabCounters, "col_reduce: scratch too small");
  launch_pdl(ctx, col_reduce_kernel, dim3(grid), dim3(256), 0, z, z2, mean, var, R, C, rows_per, mode, partial, scale, accumulate, out);
  AOCR_CUDA(cudaGetLastError());
}

}  // namespace

void conv1_fwd(Ctx& ctx, const float* x, const float* w, const float* bias, float* a1, uint8_t* idx, int B, int W,
               __nv_bfloat16* phi, __nv_bfloat16* plo) {
  int64_t total = (int64_t)B * 16 * (W / 2) * 4;          // one thread per pooled pixel and 16 channels
  launch_pdl(ctx, conv1_fwd_kernel, dim3(grid_for(total, 256, ctx.num_sms)), dim3(256), 0, x, w, bias, a1, idx, B, W, phi, plo);
  AOCR_CUDA(cudaGetLastError());
}

void conv1_bwd(Ctx& ctx, const float* x, const float* a1, const uint8_t* idx, const float* da1, float* dw, float* db,
               float* partial, int nblk, int B, int W) {
  int64_t npix = (int64_t)B * 16 * (W / 2);
  int64_t per = (npix + nblk - 1) / nblk;
  launch_pdl(ctx, conv1_bwd_partial_kernel, dim3(nblk), dim3(256), 0, x, a1, idx, da1, partial, B, W, per);
  AOCR_CUDA(cudaGetLastError());
  launch_pdl(ctx, conv1_bwd_final_kernel, dim3(10), dim3(1024), 0, partial, nblk, dw, db);
  AOCR_CUDA(cudaGetLastError());
}

void relu_pool_fwd(Ctx& ctx, const float* z, float* a, uint8_t* idx, int B, int H, int Wi, int C, int kw,
                   __nv_bfloat16* phi, __nv_bfloat16* plo) {
  int64_t total = (int64_t)B * (H / 2) * (Wi / kw) * C;
  int g = grid_for(total, 256, ctx.num_sms);
  if (kw == 2) relu_pool_fwd_kernel<2><<<g, 256, 0, ctx.st>>>(z, a, idx, B, H, Wi, C, phi, plo);
  else relu_pool_fwd_kernel<1><<<g, 256, 0, ctx.st>>>(z, a, idx, B, H, Wi, C, phi, plo);
  AOCR_LAUNCH_CHECK(ctx);
}

void relu_pool_bwd(Ctx& ctx, const float* da, const float* a, const uint8_t* idx, float* dz, int B, int H, int Wi, int C,
                   int kw, __nv_bfloat16* phi, __nv_bfloat16* plo) {
  int64_t total = (int64_t)B * H * Wi * C;
  int g = grid_for(total, 256, ctx.num_sms);
  if (kw == 2) relu_pool_bwd_kernel<2><<<g, 256, 0, ctx.st>>>(da, a, idx, dz, B, H, Wi, C, phi, plo);
  else relu_pool_bwd_kernel<1><<<g, 256, 0, ctx.st>>>(da, a, idx, dz, B, H, Wi, C, phi, plo);
  AOCR_LAUNCH_CHECK(ctx);
}

void im2col(Ctx& ctx, const float* a, float* col, int B, int H, int Wi, int C, int k, int pad) {
  int Ho = H + 2 * pad - k + 1, Wo = Wi + 2 * pad - k + 1;
  int64_t total = (int64_t)B * Ho * Wo * k * k * (C / 4);
  launch_pdl(ctx, im2col_kernel, dim3(grid_for(total, 256, ctx.num_sms)), dim3(256), 0, a, col, B, H, Wi, C, k, pad, Ho, Wo);
  AOCR_CUDA(cudaGetLastError());
}

void col_sum(Ctx& ctx, const float* z, int64_t R, int C, float* out, float* partial, int accumulate) {
  col_reduce(ctx, z, nullptr, nullptr, nullptr, R, C, 0, 1.0f, accumulate, out, partial);
}

// Batch statistics.  Under data parallelism (sync.world > 1) the local column sums are summed across ranks
// through sync.fn (an all-reduce ordered on the engine stream) before they are normalised by the GLOBAL row
// count, so every rank normalises with the statistics of the whole batch, as the single-device reference does.
void bn_stats(Ctx& ctx, const float* z, int64_t R, int C, float* mean, float* var, float* rmean, float* rvar, float* partial,
              float* sums, const StatSync& sync) {
  int ns = nslabs_for(R, ctx.num_sms, C);
  if (ns > 128) ns = 128;
  const int64_t rows_per = (R + ns - 1) / ns;
  const double Rg = (double)R * sync.grows;
  const float invR = (float)(1.0 / Rg), unbias = Rg > 1 ? (float)(Rg / (Rg - 1)) : 1.f;
  const int fin = sync.world > 1 ? 0 : 1;          // single device: one launch does sums, statistics and running update
  AOCR_CHECK(cdiv(C, 32) <= kSlabCounters && (int64_t)2 * ns * C <= kPartialFloats - kSlabCounters, "bn_stats: scratch too small");
  launch_pdl(ctx, col_reduce_bn_fwd_kernel, dim3(cdiv(C, 32), ns), dim3(256), 0, z, (const float*)rmean, R, C, rows_per, partial, sums,
             fin, invR, unbias, mean, var, rmean, rvar);
  AOCR_CUDA(cudaGetLastError());
  if (fin) return;
  sync.fn(sync.user, sums, 2 * (int64_t)C);        // ONE exchange: [sum | sum of squares]
  launch_pdl(ctx, bn_finalize_kernel, dim3(cdiv(C, 128)), dim3(128), 0, (const float*)sums, C, invR, unbias, mean, var, rmean, rvar);
  AOCR_CUDA(cudaGetLastError());
}

void bn_update_running(Ctx& ctx, const float* mean, const float* var, float* rmean, float* rvar, int C, int64_t R) {
  float unbias = R > 1 ? (float)((double)R / (double)(R - 1)) : 1.f;
  launch_pdl(ctx, bn_update_running_kernel, dim3(cdiv(C, 128)), dim3(128), 0, mean, var, rmean, rvar, C, unbias);
  AOCR_CUDA(cudaGetLastError());
}

void bn_relu_fwd(Ctx& ctx, const float* z, const float* mean, const float* var, const float* gamma, const float* beta,
                 float* a, int64_t R, int C, int tm_S, int tm_B, __nv_bfloat16* phi, __nv_bfloat16* plo) {
  launch_pdl(ctx, bn_relu_fwd_kernel, dim3(grid_for(R * (C / 4), 256, ctx.num_sms)), dim3(256), 0, z, mean, var, gamma, beta, a, R, C,
                                                                                 tm_S, tm_B, phi, plo);
  AOCR_CUDA(cudaGetLastError());
}

void bn_relu_bwd(Ctx& ctx, const float* da, const float* a, const float* z, const float* mean, const float* var,
                 const float* gamma, float* dz, float* dgamma, float* dbeta, float* partial, int64_t R, int C, int tm_S,
                 int tm_B, int train, const StatSync& sync, __nv_bfloat16* phi, __nv_bfloat16* plo) {
  launch_pdl(ctx, relu_mask_kernel, dim3(grid_for(R * (C / 4), 256, ctx.num_sms)), dim3(256), 0, da, a, dz, R, C, tm_S, tm_B);
  AOCR_CUDA(cudaGetLastError());
  // dbeta = s1 ; dgamma = s2 (each BN parameter receives gradient exactly once per step)
  {   // one pass, the summation order of the two separate reductions
    int ns = nslabs_for(R, ctx.num_sms, C);
    if (ns > 128) ns = 128;                       // partial holds [2][ns][C]
    int64_t rows_per = (R + ns - 1) / ns;
    launch_pdl(ctx, col_reduce_bn_bwd_kernel, dim3(cdiv(C, 32), ns), dim3(256), 0, (const float*)dz, z, mean, var, R, C, rows_per,
               partial, dbeta, dgamma);
    AOCR_CUDA(cudaGetLastError());
  }
  if (sync.world > 1) {   // global sums of dy and dy*xhat (dgamma, dbeta are adjacent: one all-reduce when contiguous)
    if (dbeta == dgamma + C) sync.fn(sync.user, dgamma, 2 * (int64_t)C);
    else { sync.fn(sync.user, dgamma, C); sync.fn(sync.user, dbeta, C); }
  }
  launch_pdl(ctx, bn_bwd_apply_kernel, dim3(grid_for(R * C, 256, ctx.num_sms)), dim3(256), 0, dz, z, mean, var, gamma, dbeta, dgamma,
                                                                           R, (int64_t)llround((double)R * sync.grows), C, train, phi, plo);
  AOCR_CUDA(cudaGetLastError());
  if (sync.world > 1) {   // the gradient all-reduce will sum these again over ranks: pre-divide
    scale_vec(ctx, dgamma, C, 1.0f / (float)sync.world);
    scale_vec(ctx, dbeta, C, 1.0f / (float)sync.world);
  }
}

}  // namespace aocr
