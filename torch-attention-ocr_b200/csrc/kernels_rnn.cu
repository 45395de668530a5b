// kernels_rnn.cu — LSTM cells, attention, generator/criterion, greedy selection, optimiser.
// Reference math: src/model/LSTM.lua:79-105 (cell, gate order [in|forget|out|candidate]),
// :124-162 (attention), src/model/output_projector.lua:5-6, src/model/criterion.lua:4-7,
// src/model/model.lua:402,448-458 (greedy), src/optim/optim_sgd.lua:49-52,90 (clip + SGD).
#include "kernels.h"

namespace aocr {

namespace {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

inline int grid_for(int64_t total, int threads, int num_sms) {
  int64_t g = (total + threads - 1) / threads;
  int64_t cap = (int64_t)num_sms * 8;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

// ------------------------------------------------------------------ encoder step (both directions)
// grid (He/8, 2, ceil(B/32)); CTA tile = 32 batch rows x (8 hidden units x 4 gates); K = He in chunks of 32.
__global__ void __launch_bounds__(256) enc_step_fwd_kernel(EncStep p) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float hs[32][33];
  __shared__ float ws[32][33];
  __shared__ float gs[32][33];
  const int tid = threadIdx.x, tx = tid % 32, ty = tid / 32;
  const int d = blockIdx.y, u0 = blockIdx.x * 8, b0 = blockIdx.z * 32;
  const int He = p.He, B = p.B, S = p.S;
  const int t = d == 0 ? p.step : S - 1 - p.step;
  const int prev_slot = d == 0 ? t : t + 1, out_slot = d == 0 ? t + 1 : t;
  const float* hprev = p.H + ((int64_t)(d * (S + 1) + prev_slot) * B) * He;
  const float* W = p.Wh[d];
  // column j of the tile <-> gate j/8, unit u0 + j%8
  const int wrow_l = tid / 8;           // 0..31 : tile column loaded by this thread
  const int wk_l = (tid % 8) * 4;       // 4 consecutive k
  const int64_t wrow_g = (int64_t)((wrow_l / 8) * He + u0 + (wrow_l % 8)) * He;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k0 = 0; k0 < He; k0 += 32) {
    {  // h tile: 32 rows x 32 k ; thread loads 4 consecutive k of one row
      int r = tid / 8, kk = (tid % 8) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (b0 + r < B) v = *reinterpret_cast<const float4*>(hprev + (int64_t)(b0 + r) * He + k0 + kk);
      hs[r][kk] = v.x; hs[r][kk + 1] = v.y; hs[r][kk + 2] = v.z; hs[r][kk + 3] = v.w;
      float4 w = *reinterpret_cast<const float4*>(W + wrow_g + k0 + wk_l);
      ws[wrow_l][wk_l] = w.x; ws[wrow_l][wk_l + 1] = w.y; ws[wrow_l][wk_l + 2] = w.z; ws[wrow_l][wk_l + 3] = w.w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 32; k++) {
      float wv = ws[tx][k];
#pragma unroll
      for (int i = 0; i < 4; i++) acc[i] = fmaf(hs[ty * 4 + i][k], wv, acc[i]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; i++) gs[ty * 4 + i][tx] = acc[i];
  __syncthreads();
  // cell update: thread <-> (row r, unit u)
  const int r = tid / 8, u = tid % 8;
  const int b = b0 + r;
  if (b < B) {
    const int unit = u0 + u;
    const float* xg = p.xg + ((int64_t)t * B + b) * (8 * He) + (int64_t)d * 4 * He + unit;
    float gi = gs[r][0 * 8 + u] + xg[0 * He];
    float gf = gs[r][1 * 8 + u] + xg[1 * He];
    float go = gs[r][2 * 8 + u] + xg[2 * He];
    float gg = gs[r][3 * 8 + u] + xg[3 * He];
    float i_ = sigmoidf_(gi), f_ = sigmoidf_(gf), o_ = sigmoidf_(go), g_ = tanhf(gg);
    float cp = p.Cst[((int64_t)(d * (S + 1) + prev_slot) * B + b) * He + unit];
    float c = f_ * cp + i_ * g_;
    float h = o_ * tanhf(c);
    p.Cst[((int64_t)(d * (S + 1) + out_slot) * B + b) * He + unit] = c;
    p.H[((int64_t)(d * (S + 1) + out_slot) * B + b) * He + unit] = h;
    float* a = p.acts + (((int64_t)(d * S + t) * B + b) * 4) * He + unit;
    a[0] = i_; a[He] = f_; a[2 * He] = o_; a[3 * He] = g_;
    p.ctx[((int64_t)b * S + t) * (2 * He) + d * He + unit] = h;
  }
}

// elementwise over (dir, b, unit)
__global__ void __launch_bounds__(256) enc_cell_bwd_kernel(EncStepBwd p) {
  pdl_launch_dependents();
  pdl_wait();
  const int He = p.He, B = p.B, S = p.S;
  const int64_t total = (int64_t)2 * B * He;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int unit = (int)(e % He);
    int b = (int)((e / He) % B);
    int d = (int)(e / ((int64_t)He * B));
    // backward order: fw walks t = S-1..0, bw walks t = 0..S-1 (model.lua:668,682)
    int t = d == 0 ? S - 1 - p.step : p.step;
    int prev_slot = d == 0 ? t : t + 1, out_slot = d == 0 ? t + 1 : t;
    const float* a = p.acts + (((int64_t)(d * S + t) * B + b) * 4) * He + unit;
    float i_ = a[0], f_ = a[He], o_ = a[2 * He], g_ = a[3 * He];
    float cp = p.Cst[((int64_t)(d * (S + 1) + prev_slot) * B + b) * He + unit];
    float tc = tanhf(p.Cst[((int64_t)(d * (S + 1) + out_slot) * B + b) * He + unit]);
    float dh = p.dh[e] + p.Dctx[((int64_t)b * S + t) * (2 * He) + d * He + unit];
    float dc = p.dc[e] + dh * o_ * (1.f - tc * tc);
    float* dg = p.dG + ((int64_t)t * B + b) * (8 * He) + (int64_t)d * 4 * He + unit;
    dg[0] = dc * g_ * i_ * (1.f - i_);
    dg[He] = dc * cp * f_ * (1.f - f_);
    dg[2 * He] = dh * tc * o_ * (1.f - o_);
    dg[3 * He] = dc * i_ * (1.f - g_ * g_);
    p.dc[e] = dc * f_;
  }
}

// ------------------------------------------------------------------ decoder cells
__global__ void __launch_bounds__(256) dec_cell_fwd_kernel(DecCell p) {
  pdl_launch_dependents();
  pdl_wait();
  const int H = p.H;
  const int64_t total = (int64_t)p.B * H;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int u = (int)(e % H);
    int b = (int)(e / H);
    const float* g = p.G + (int64_t)b * 4 * H + u;
    const float* ar = p.addrows + (p.rowsel ? (int64_t)(p.rowsel[b] - 1) * p.addld : 0) + u;
    float i_ = sigmoidf_(g[0] + ar[0]);
    float f_ = sigmoidf_(g[H] + ar[H]);
    float o_ = sigmoidf_(g[2 * H] + ar[2 * H]);
    float g_ = tanhf(g[3 * H] + ar[3 * H]);
    float c = f_ * p.c_prev[e] + i_ * g_;
    float h = o_ * tanhf(c);
    p.c_new[e] = c;
    float* a = p.acts + (int64_t)b * 4 * H + u;
    a[0] = i_; a[H] = f_; a[2 * H] = o_; a[3 * H] = g_;
    p.h_out0[(int64_t)b * p.ld0 + u] = h;
    if (p.h_out1) p.h_out1[(int64_t)b * p.ld1 + u] = h;
  }
}

__global__ void __launch_bounds__(256) dec_cell_bwd_kernel(DecCellBwd p) {
  pdl_launch_dependents();
  pdl_wait();
  const int H = p.H;
  const int64_t total = (int64_t)p.B * H;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int u = (int)(e % H);
    int b = (int)(e / H);
    float dh = 0.f;
    if (p.dh_a) dh += p.dh_a[(int64_t)b * p.lda + u];
    if (p.dh_b) dh += p.dh_b[(int64_t)b * p.ldb + u];
    if (p.dh_c) dh += p.dh_c[(int64_t)b * p.ldc + u];
    const float* a = p.acts + (int64_t)b * 4 * H + u;
    float i_ = a[0], f_ = a[H], o_ = a[2 * H], g_ = a[3 * H];
    float tc = tanhf(p.c_new[e]);
    float dc = p.dc[e] + dh * o_ * (1.f - tc * tc);
    float* dg = p.dG + (int64_t)b * 4 * H + u;
    dg[0] = dc * g_ * i_ * (1.f - i_);
    dg[H] = dc * p.c_prev[e] * f_ * (1.f - f_);
    dg[2 * H] = dh * tc * o_ * (1.f - o_);
    dg[3 * H] = dc * i_ * (1.f - g_ * g_);
    p.dc[e] = dc * f_;
  }
}

// ------------------------------------------------------------------ attention
// One CTA (8 warps) per batch row.  Each lane owns H/32 channels as float4 groups at lane*4 + 128*i, so a
// warp reads one context row as fully coalesced 512-byte requests.  ctx[b] is read ONCE: the row that gave
// the score is still in registers when it is folded into the running (online-softmax) context vector.
constexpr int ATT_WARPS = 8;
constexpr int ATT_MAXV = 8;   // H <= 1024

__global__ void __launch_bounds__(256) attn_fwd_kernel(const float* __restrict__ ctx, const float* __restrict__ q,
                                                       float* __restrict__ alpha, float* __restrict__ cv, int64_t ldcv,
                                                       int S, int H) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float sm[];
  float* es = sm;                       // S scores
  float* wm = es + ((S + 3) & ~3);      // ATT_WARPS maxima
  float* wl = wm + ATT_WARPS;           // ATT_WARPS partial sums
  float* accs = wl + ATT_WARPS;         // ATT_WARPS x H partial context vectors
  const int b = blockIdx.x, warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int nv = H / 128;
  const float* cb = ctx + (int64_t)b * S * H;
  float4 qv[ATT_MAXV], acc[ATT_MAXV];
#pragma unroll
  for (int i = 0; i < ATT_MAXV; i++) {
    acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    qv[i] = (i < nv) ? *reinterpret_cast<const float4*>(q + (int64_t)b * H + lane * 4 + 128 * i)
                     : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float m = -INFINITY, l = 0.f;
  for (int s = warp; s < S; s += ATT_WARPS) {
    float4 row[ATT_MAXV];
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < ATT_MAXV; i++) {
      if (i < nv) {
        row[i] = *reinterpret_cast<const float4*>(cb + (int64_t)s * H + lane * 4 + 128 * i);
        dot += row[i].x * qv[i].x + row[i].y * qv[i].y + row[i].z * qv[i].z + row[i].w * qv[i].w;
      }
    }
    dot = warp_sum(dot);
    if (lane == 0) es[s] = dot;
    float mn = fmaxf(m, dot);
    float sc = expf(m - mn);     // m = -inf on the first row -> 0
    float pe = expf(dot - mn);
    l = l * sc + pe;
#pragma unroll
    for (int i = 0; i < ATT_MAXV; i++) {
      if (i < nv) {
        acc[i].x = acc[i].x * sc + pe * row[i].x;
        acc[i].y = acc[i].y * sc + pe * row[i].y;
        acc[i].z = acc[i].z * sc + pe * row[i].z;
        acc[i].w = acc[i].w * sc + pe * row[i].w;
      }
    }
    m = mn;
  }
  if (lane == 0) { wm[warp] = m; wl[warp] = l; }
  __syncthreads();
  float M = -INFINITY;
#pragma unroll
  for (int w = 0; w < ATT_WARPS; w++) M = fmaxf(M, wm[w]);
  float L = 0.f;
#pragma unroll
  for (int w = 0; w < ATT_WARPS; w++) L += (wm[w] == -INFINITY) ? 0.f : wl[w] * expf(wm[w] - M);
  const float myscale = (m == -INFINITY) ? 0.f : expf(m - M) / L;
#pragma unroll
  for (int i = 0; i < ATT_MAXV; i++) {
    if (i < nv) {
      float4 v = acc[i];
      v.x *= myscale; v.y *= myscale; v.z *= myscale; v.w *= myscale;
      *reinterpret_cast<float4*>(accs + warp * H + lane * 4 + 128 * i) = v;
    }
  }
  __syncthreads();
  for (int h = threadIdx.x; h < H; h += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < ATT_WARPS; w++) s += accs[w * H + h];
    cv[(int64_t)b * ldcv + h] = s;
  }
  for (int s = threadIdx.x; s < S; s += blockDim.x) alpha[(int64_t)b * S + s] = expf(es[s] - M) / L;
}

__global__ void __launch_bounds__(256) attn_bwd_kernel(const float* __restrict__ ctx, const float* __restrict__ alpha,
                                                       const float* __restrict__ dcv, int64_t lddcv,
                                                       float* __restrict__ de, float* __restrict__ dq, int S, int H) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float sm[];
  float* das = sm;                        // S: dalpha then de
  float* red = das + ((S + 3) & ~3);      // 1
  float* accs = red + 4;                  // ATT_WARPS x H
  const int b = blockIdx.x, warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int nv = H / 128;
  const float* cb = ctx + (int64_t)b * S * H;
  float4 gv[ATT_MAXV];
#pragma unroll
  for (int i = 0; i < ATT_MAXV; i++)
    gv[i] = (i < nv) ? *reinterpret_cast<const float4*>(dcv + (int64_t)b * lddcv + lane * 4 + 128 * i)
                     : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s = warp; s < S; s += ATT_WARPS) {
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < ATT_MAXV; i++) {
      if (i < nv) {
        float4 r = *reinterpret_cast<const float4*>(cb + (int64_t)s * H + lane * 4 + 128 * i);
        dot += r.x * gv[i].x + r.y * gv[i].y + r.z * gv[i].z + r.w * gv[i].w;
      }
    }
    dot = warp_sum(dot);
    if (lane == 0) das[s] = dot;
  }
  __syncthreads();
  if (warp == 0) {
    float s_ = 0.f;
    for (int s = lane; s < S; s += 32) s_ += alpha[(int64_t)b * S + s] * das[s];
    s_ = warp_sum(s_);
    if (lane == 0) red[0] = s_;
  }
  __syncthreads();
  const float tot = red[0];
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    float v = alpha[(int64_t)b * S + s] * (das[s] - tot);
    de[(int64_t)b * S + s] = v;
  }
  __syncthreads();   // das[] still holds dalpha; recompute de on the fly below
  float4 acc[ATT_MAXV];
#pragma unroll
  for (int i = 0; i < ATT_MAXV; i++) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s = warp; s < S; s += ATT_WARPS) {
    float w = alpha[(int64_t)b * S + s] * (das[s] - tot);
#pragma unroll
    for (int i = 0; i < ATT_MAXV; i++) {
      if (i < nv) {
        float4 r = *reinterpret_cast<const float4*>(cb + (int64_t)s * H + lane * 4 + 128 * i);
        acc[i].x = fmaf(w, r.x, acc[i].x); acc[i].y = fmaf(w, r.y, acc[i].y);
        acc[i].z = fmaf(w, r.z, acc[i].z); acc[i].w = fmaf(w, r.w, acc[i].w);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < ATT_MAXV; i++)
    if (i < nv) *reinterpret_cast<float4*>(accs + warp * H + lane * 4 + 128 * i) = acc[i];
  __syncthreads();
  for (int h = threadIdx.x; h < H; h += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < ATT_WARPS; w++) s += accs[w * H + h];
    dq[(int64_t)b * H + h] = s;
  }
}

// ------------------------------------------------------------------ generator + criterion
// one CTA (4 warps) per row; the activation row stays in registers while the 39 weight rows stream by.
constexpr int GEN_MAXV = 64;
__global__ void __launch_bounds__(128) generator_kernel(const float* __restrict__ a, const float* __restrict__ W,
                                                        const float* __restrict__ bias, const int32_t* __restrict__ y,
                                                        float* __restrict__ logp, float* __restrict__ dz,
                                                        float* __restrict__ rowloss, int H, int V, float inv_bn) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float zs[GEN_MAXV];
  const int64_t r = blockIdx.x;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int nv = H / 128;
  float4 av[ATT_MAXV];
#pragma unroll
  for (int i = 0; i < ATT_MAXV; i++)
    av[i] = (i < nv) ? *reinterpret_cast<const float4*>(a + r * H + lane * 4 + 128 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int v = warp; v < V; v += 4) {
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < ATT_MAXV; i++) {
      if (i < nv) {
        float4 w = *reinterpret_cast<const float4*>(W + (int64_t)v * H + lane * 4 + 128 * i);
        dot += w.x * av[i].x + w.y * av[i].y + w.z * av[i].z + w.w * av[i].w;
      }
    }
    dot = warp_sum(dot);
    if (lane == 0) zs[v] = dot + bias[v];
  }
  __syncthreads();
  if (warp == 0) {
    float z0 = lane < V ? zs[lane] : -INFINITY;
    float z1 = lane + 32 < V ? zs[lane + 32] : -INFINITY;
    float mx = warp_max(fmaxf(z0, z1));
    float se = (lane < V ? expf(z0 - mx) : 0.f) + (lane + 32 < V ? expf(z1 - mx) : 0.f);
    se = warp_sum(se);
    float lse = mx + logf(se);
    int yy = y ? y[r] - 1 : -1;
    float w = (y && yy != 0) ? 1.f : 0.f;   // PAD (id 1 -> index 0) has weight 0
#pragma unroll
    for (int j = 0; j < 2; j++) {
      int v = lane + 32 * j;
      if (v < V) {
        float lp = (j == 0 ? z0 : z1) - lse;
        logp[r * V + v] = lp;
        if (dz) dz[r * V + v] = (expf(lp) - (v == yy ? 1.f : 0.f)) * w * inv_bn;
        if (rowloss && v == yy) rowloss[r] = -w * lp;
      }
    }
  }
}

__global__ void __launch_bounds__(1024) reduce_sum_double_kernel(const float* __restrict__ v, int64_t n,
                                                                 double* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ double red[1024];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += 1024) s += (double)v[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = red[0];
}

__global__ void greedy_select_kernel(float* __restrict__ logp, int32_t* __restrict__ tok, double* __restrict__ score,
                                     int32_t* __restrict__ labels, int64_t ldl, int t, int B, int V) {
  pdl_launch_dependents();
  pdl_wait();
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float* lp = logp + (int64_t)b * V;
  if (t > 0) {
    int prev = tok[b];
    if (prev == 1 || prev == 3) lp[0] = 0.f;   // model.lua:448-449
  }
  float best = lp[0];
  int bi = 0;
  for (int v = 1; v < V; v++) {
    float x = lp[v];
    if (x > best) { best = x; bi = v; }
  }
  score[b] = (t == 0 ? 0.0 : score[b]) + (double)best;
  tok[b] = bi + 1;
  labels[(int64_t)b * ldl + t] = bi + 1;
}

// ------------------------------------------------------------------ helpers
__global__ void add_vec_kernel(float* out, const float* a, const float* b, int64_t n) {
  pdl_launch_dependents();
  pdl_wait();
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
    out[e] = a[e] + b[e];
}
__global__ void copy_strided_kernel(float* dst, int64_t ldd, const float* src, int64_t lds, int rows, int cols) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t total = (int64_t)rows * cols;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(e % cols);
    int64_t r = e / cols;
    dst[r * ldd + c] = src[r * lds + c];
  }
}
__global__ void concat2_kernel(const float* s0, const float* s1, float* dst, int64_t ldd, int B, int He) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t total = (int64_t)B * 2 * He;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(e % (2 * He));
    int b = (int)(e / (2 * He));
    dst[(int64_t)b * ldd + c] = c < He ? s0[(int64_t)b * He + c] : s1[(int64_t)b * He + c - He];
  }
}
// du = (da_carry + da_gen) * (1 - a^2)
__global__ void du_from_da_kernel(const float* da_carry, int64_t ldc, const float* da_gen, const float* a, float* du,
                                  int64_t n, int H) {
  pdl_launch_dependents();
  pdl_wait();
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    int u = (int)(e % H);
    int64_t b = e / H;
    float g = da_gen[e] + (da_carry ? da_carry[b * ldc + u] : 0.f);
    float av = a[e];
    du[e] = g * (1.f - av * av);
  }
}
__global__ void gather_tokens_kernel(const int32_t* tgt_bt, int32_t* out_tb, int B, int T, int Tpad, int32_t padval) {
  pdl_launch_dependents();
  pdl_wait();
  const int total = B * Tpad;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    int b = e % B, t = e / B;
    out_tb[e] = t < T ? tgt_bt[b * T + t] : padval;
  }
}
// grid (N/256, V): column n of token v
__global__ void __launch_bounds__(256) token_segment_sum_kernel(const float* __restrict__ dG,
                                                                const int32_t* __restrict__ y, float* __restrict__ dP,
                                                                int64_t R, int N) {
  pdl_launch_dependents();
  pdl_wait();
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  int v = blockIdx.y;
  if (n >= N) return;
  float s = 0.f;
  for (int64_t r = 0; r < R; r++)
    if (y[r] - 1 == v) s += dG[r * N + n];
  dP[(int64_t)v * N + n] = s;
}

// oh[r][v] = 1 if token y[r] == v+1 : turns the embedding-row segment sum into a GEMM
__global__ void onehot_kernel(const int32_t* __restrict__ y, float* __restrict__ oh, int64_t R, int V) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t total = R * V;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x)
    oh[e] = (y[e / V] - 1 == (int)(e % V)) ? 1.f : 0.f;
}
// out[m][e] = sum_k A[m*lda + k] * W[k*ldw + e], e < E <= 32 : one CTA per row m, K split over the threads
__global__ void __launch_bounds__(256) thin_n_gemm_kernel(const float* __restrict__ A, int64_t lda,
                                                          const float* __restrict__ W, int64_t ldw,
                                                          float* __restrict__ out, int64_t ldo, int K, int E) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[8][32];
  const int m = blockIdx.x, lane = threadIdx.x % 32, warp = threadIdx.x / 32;
  float acc[32];
#pragma unroll
  for (int e = 0; e < 32; e++) acc[e] = 0.f;
  for (int k = threadIdx.x; k < K; k += 256) {          // coalesced over k; each thread owns whole rows of W
    const float a = A[(int64_t)m * lda + k];
    const float* w = W + (int64_t)k * ldw;
#pragma unroll
    for (int e = 0; e < 32; e++)
      if (e < E) acc[e] = fmaf(a, w[e], acc[e]);
  }
#pragma unroll
  for (int e = 0; e < 32; e++) {
    float v = acc[e];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][e] = v;
  }
  __syncthreads();
  if (warp == 0 && lane < E) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; w++) s += red[w][lane];
    out[(int64_t)m * ldo + lane] = s;
  }
}

// ------------------------------------------------------------------ optimiser
__global__ void __launch_bounds__(256) sumsq_partial_kernel(const float* __restrict__ v, int64_t n,
                                                            double* __restrict__ partial) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ double red[256];
  double s = 0.0;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    double x = (double)v[e];
    s += x * x;
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}
__global__ void __launch_bounds__(256) sumsq_final_kernel(const double* partial, int nblk, double* out) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ double red[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < nblk; i += 256) s += partial[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = red[0];
}
__global__ void __launch_bounds__(256) sgd_apply_kernel(float* __restrict__ p, float* __restrict__ g, int64_t n,
                                                        const double* __restrict__ sumsq, double lr, double clip) {
  pdl_launch_dependents();
  pdl_wait();
  const double norm = sqrt(*sumsq);
  const float scale = norm > clip ? (float)(clip / norm) : 1.0f;   // optim_sgd.lua:50-52
  const float flr = (float)lr;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    float gv = g[e] * scale;
    g[e] = gv;                       // the reference clips dfdy in place
    p[e] = p[e] - flr * gv;          // optim_sgd.lua:90
  }
}

// ---- the same for all parameter groups at once (3 launches per update instead of 15) ----------------------------
__device__ __forceinline__ int group_of_block(const SgdGroups& G, int blk, int& local) {
  int g = 0, first = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    if (g == i && blk >= first + G.nb[i]) { first += G.nb[i]; g = i + 1; }
  }
  local = blk - first;
  return g;
}
__global__ void __launch_bounds__(256) sumsq_groups_kernel(const float* __restrict__ base, SgdGroups G,
                                                           double* __restrict__ partial) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ double red[8];
  int local;
  const int g = group_of_block(G, blockIdx.x, local);
  const float4* v = reinterpret_cast<const float4*>(base + G.off[g]);
  const int64_t n4 = G.n[g] / 4;           // group extents are padded to 64 floats
  double s = 0.0;
  for (int64_t e = (int64_t)local * blockDim.x + threadIdx.x; e < n4; e += (int64_t)G.nb[g] * blockDim.x) {
    const float4 x = v[e];
    s += (double)x.x * x.x + (double)x.y * x.y + (double)x.z * x.z + (double)x.w * x.w;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 8; w++) t += red[w];
    partial[g * 1024 + local] = t;
  }
}
__global__ void __launch_bounds__(256) sumsq_final_groups_kernel(const double* __restrict__ partial, SgdGroups G,
                                                                 double* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ double red[256];
  const int g = blockIdx.x;
  if (G.nb[g] == 0) return;          // group not part of this update: leave its sum alone
  double s = 0.0;
  for (int i = threadIdx.x; i < G.nb[g]; i += 256) s += partial[g * 1024 + i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[g] = red[0];
}
__global__ void __launch_bounds__(256) sgd_apply_groups_kernel(float* __restrict__ pbase, float* __restrict__ gbase, SgdGroups G,
                                                               const double* __restrict__ sumsq,
                                                               const double* __restrict__ lrclip) {
  pdl_launch_dependents();
  pdl_wait();
  const double lr = lrclip[0], clip = lrclip[1];
  int local;
  const int g = group_of_block(G, blockIdx.x, local);
  const double norm = sqrt(sumsq[g]);
  const float scale = norm > clip ? (float)(clip / norm) : 1.0f;   // optim_sgd.lua:50-52
  const float flr = (float)lr;
  float4* p = reinterpret_cast<float4*>(pbase + G.off[g]);
  float4* gr = reinterpret_cast<float4*>(gbase + G.off[g]);
  const int64_t n4 = G.n[g] / 4;
  for (int64_t e = (int64_t)local * blockDim.x + threadIdx.x; e < n4; e += (int64_t)G.nb[g] * blockDim.x) {
    float4 gv = gr[e];
    float4 pv = p[e];
    gv.x *= scale; gv.y *= scale; gv.z *= scale; gv.w *= scale;
    gr[e] = gv;                      // the reference clips dfdy in place
    pv.x -= flr * gv.x; pv.y -= flr * gv.y; pv.z -= flr * gv.z; pv.w -= flr * gv.w;   // optim_sgd.lua:90
    p[e] = pv;
  }
}

}  // namespace

void sgd_groups(Ctx& ctx, float* params, float* grads, const SgdGroups& G, double* partial, double* sumsq,
                const double* lrclip) {
  int total = 0;
  for (int g = 0; g < 5; g++) total += G.nb[g];
  launch_pdl(ctx, sumsq_groups_kernel, dim3(total), dim3(256), 0, (const float*)grads, G, partial);
  AOCR_CUDA(cudaGetLastError());
  launch_pdl(ctx, sumsq_final_groups_kernel, dim3(5), dim3(256), 0, (const double*)partial, G, sumsq);
  AOCR_CUDA(cudaGetLastError());
  launch_pdl(ctx, sgd_apply_groups_kernel, dim3(total), dim3(256), 0, params, grads, G, (const double*)sumsq, lrclip);
  AOCR_CUDA(cudaGetLastError());
}

void enc_step_fwd(Ctx& ctx, const EncStep& p) {
  dim3 grid(p.He / 8, 2, cdiv(p.B, 32));
  launch_pdl(ctx, enc_step_fwd_kernel, dim3(grid), dim3(256), 0, p);
  AOCR_CUDA(cudaGetLastError());
}
void enc_cell_bwd(Ctx& ctx, const EncStepBwd& p) {
  launch_pdl(ctx, enc_cell_bwd_kernel, dim3(grid_for((int64_t)2 * p.B * p.He, 256, ctx.num_sms)), dim3(256), 0, p);
  AOCR_CUDA(cudaGetLastError());
}
void dec_cell_fwd(Ctx& ctx, const DecCell& p) {
  launch_pdl(ctx, dec_cell_fwd_kernel, dim3(grid_for((int64_t)p.B * p.H, 256, ctx.num_sms)), dim3(256), 0, p);
  AOCR_CUDA(cudaGetLastError());
}
void dec_cell_bwd(Ctx& ctx, const DecCellBwd& p) {
  launch_pdl(ctx, dec_cell_bwd_kernel, dim3(grid_for((int64_t)p.B * p.H, 256, ctx.num_sms)), dim3(256), 0, p);
  AOCR_CUDA(cudaGetLastError());
}
void attn_fwd(Ctx& ctx, const float* c, const float* q, float* alpha, float* cv, int64_t ldcv, int B, int S, int H) {
  AOCR_CHECK(H % 128 == 0 && H <= 128 * ATT_MAXV, "attention kernel needs decoder hidden size in {128,...,1024}");
  size_t smem = (size_t)(((S + 3) & ~3) + 2 * ATT_WARPS + ATT_WARPS * H) * sizeof(float);
  launch_pdl(ctx, attn_fwd_kernel, dim3(B), dim3(256), smem, c, q, alpha, cv, ldcv, S, H);
  AOCR_CUDA(cudaGetLastError());
}
void attn_bwd(Ctx& ctx, const float* c, const float* alpha, const float* dcv, int64_t lddcv, float* de, float* dq, int B,
              int S, int H) {
  size_t smem = (size_t)(((S + 3) & ~3) + 4 + ATT_WARPS * H) * sizeof(float);
  launch_pdl(ctx, attn_bwd_kernel, dim3(B), dim3(256), smem, c, alpha, dcv, lddcv, de, dq, S, H);
  AOCR_CUDA(cudaGetLastError());
}
void generator_fwd(Ctx& ctx, const float* a, const float* W, const float* bias, const int32_t* y, float* logp, float* dz,
                   float* rowloss, int64_t R, int H, int V, float inv_bn) {
  AOCR_CHECK(V <= GEN_MAXV && H % 128 == 0 && H <= 128 * ATT_MAXV, "generator kernel: V<=64, H in {128..1024}");
  launch_pdl(ctx, generator_kernel, dim3((unsigned)R), dim3(128), 0, a, W, bias, y, logp, dz, rowloss, H, V, inv_bn);
  AOCR_CUDA(cudaGetLastError());
}
void reduce_sum_double(Ctx& ctx, const float* v, int64_t n, double* out) {
  launch_pdl(ctx, reduce_sum_double_kernel, dim3(1), dim3(1024), 0, v, n, out);
  AOCR_CUDA(cudaGetLastError());
}
void greedy_select(Ctx& ctx, float* logp, int32_t* tok, double* score, int32_t* labels, int64_t ldl, int t, int B, int V) {
  launch_pdl(ctx, greedy_select_kernel, dim3(cdiv(B, 128)), dim3(128), 0, logp, tok, score, labels, ldl, t, B, V);
  AOCR_CUDA(cudaGetLastError());
}
void fill_zero(Ctx& ctx, void* p, size_t bytes) {
  if (bytes) AOCR_CUDA(cudaMemsetAsync(p, 0, bytes, ctx.st));
}
void add_vec(Ctx& ctx, float* out, const float* a, const float* b, int64_t n) {
  launch_pdl(ctx, add_vec_kernel, dim3(grid_for(n, 256, ctx.num_sms)), dim3(256), 0, out, a, b, n);
  AOCR_CUDA(cudaGetLastError());
}
void copy_strided(Ctx& ctx, float* dst, int64_t ldd, const float* src, int64_t lds, int rows, int cols) {
  launch_pdl(ctx, copy_strided_kernel, dim3(grid_for((int64_t)rows * cols, 256, ctx.num_sms)), dim3(256), 0, dst, ldd, src, lds, rows, cols);
  AOCR_CUDA(cudaGetLastError());
}
void concat_enc_finals(Ctx& ctx, const float* s0, const float* s1, float* dst, int64_t ldd, int B, int He) {
  launch_pdl(ctx, concat2_kernel, dim3(grid_for((int64_t)B * 2 * He, 256, ctx.num_sms)), dim3(256), 0, s0, s1, dst, ldd, B, He);
  AOCR_CUDA(cudaGetLastError());
}
void du_from_da(Ctx& ctx, const float* da_carry, int64_t ldc, const float* da_gen, const float* a, float* du, int64_t n,
                int H) {
  launch_pdl(ctx, du_from_da_kernel, dim3(grid_for(n, 256, ctx.num_sms)), dim3(256), 0, da_carry, ldc, da_gen, a, du, n, H);
  AOCR_CUDA(cudaGetLastError());
}
void gather_tokens(Ctx& ctx, const int32_t* tgt_bt, int32_t* out_tb, int B, int T, int Tpad, int32_t padval) {
  launch_pdl(ctx, gather_tokens_kernel, dim3(cdiv((int64_t)B * Tpad, 256)), dim3(256), 0, tgt_bt, out_tb, B, T, Tpad, padval);
  AOCR_CUDA(cudaGetLastError());
}
void token_segment_sum(Ctx& ctx, const float* dG, const int32_t* y, float* dP, int64_t R, int N, int V) {
  dim3 grid(cdiv(N, 256), V);
  launch_pdl(ctx, token_segment_sum_kernel, dim3(grid), dim3(256), 0, dG, y, dP, R, N);
  AOCR_CUDA(cudaGetLastError());
}
void onehot(Ctx& ctx, const int32_t* y, float* oh, int64_t R, int V) {
  launch_pdl(ctx, onehot_kernel, dim3(grid_for(R * V, 256, ctx.num_sms)), dim3(256), 0, y, oh, R, V);
  AOCR_CUDA(cudaGetLastError());
}
void thin_n_gemm(Ctx& ctx, const float* A, int64_t lda, const float* W, int64_t ldw, float* out, int64_t ldo, int M, int K,
                 int E) {
  AOCR_CHECK(E <= 32, "thin_n_gemm: N must be <= 32");
  launch_pdl(ctx, thin_n_gemm_kernel, dim3(M), dim3(256), 0, A, lda, W, ldw, out, ldo, K, E);
  AOCR_CUDA(cudaGetLastError());
}
void sumsq_partial(Ctx& ctx, const float* v, int64_t n, double* partial, int nblk) {
  launch_pdl(ctx, sumsq_partial_kernel, dim3(nblk), dim3(256), 0, v, n, partial);
  AOCR_CUDA(cudaGetLastError());
}
void sumsq_final(Ctx& ctx, const double* partial, int nblk, double* out) {
  launch_pdl(ctx, sumsq_final_kernel, dim3(1), dim3(256), 0, partial, nblk, out);
  AOCR_CUDA(cudaGetLastError());
}
void sgd_apply(Ctx& ctx, float* p, float* g, int64_t n, const double* sumsq, double lr, double clip) {
  launch_pdl(ctx, sgd_apply_kernel, dim3(grid_for(n, 256, ctx.num_sms)), dim3(256), 0, p, g, n, sumsq, lr, clip);
  AOCR_CUDA(cudaGetLastError());
}

}  // namespace aocr

namespace aocr {
namespace {
__global__ void scale_vec_kernel(float* v, int64_t n, float s) {
  pdl_launch_dependents();
  pdl_wait();
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) v[e] *= s;
}
__global__ void axpy_vec_kernel(float* y, const float* x, int64_t n, float a) {
  pdl_launch_dependents();
  pdl_wait();
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
    y[e] = fmaf(a, x[e], y[e]);
}
}  // namespace
void scale_vec(Ctx& ctx, float* v, int64_t n, float s) {
  launch_pdl(ctx, scale_vec_kernel, dim3(1184), dim3(256), 0, v, n, s);
  AOCR_CUDA(cudaGetLastError());
}
void axpy_vec(Ctx& ctx, float* y, const float* x, int64_t n, float a) {
  launch_pdl(ctx, axpy_vec_kernel, dim3(1184), dim3(256), 0, y, x, n, a);
  AOCR_CUDA(cudaGetLastError());
}
}  // namespace aocr
