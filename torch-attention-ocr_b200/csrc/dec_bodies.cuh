// dec_bodies.cuh — device bodies of the fused recurrence kernels (cell / attention / output), shared by the
// stand-alone kernels (kernels_dec.cu) and by the persistent recurrence executor (persist.cu).
// Each body is written for a 256-thread CTA; `bid` / `nblk` are the CTA's index and the number of CTAs sharing the
// op, `sm` is a scratch region of shared memory (attention: ((S+3)&~3) + 16 + 9*H floats).
#pragma once
#include "kernels_dec.h"

// AOCR_BT(i): optional phase time stamps inside the bodies (persist.cu defines it under AOCR_PERSIST_TRACE builds)
#ifndef AOCR_VARIANT
#define AOCR_VARIANT 0
#endif
#ifndef AOCR_BT
#define AOCR_BT(i)
#endif

namespace aocr {
namespace decb {


__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// Split-K partial sums.  All (<= kMaxSplits) loads are issued before the first add: a `for (z < nz) s += load` loop
// serialises one L2 round trip per partial (measured: 8 us for a cell body), this form costs one.
constexpr int kMaxSplits = 8;
__device__ __forceinline__ float part_load(const PartIn& a, int64_t b, int64_t j) {
  const float* p = a.p + b * a.ld + j;
  float v[kMaxSplits];
#pragma unroll
  for (int z = 0; z < kMaxSplits; z++) v[z] = (z < a.nz) ? __ldcg(p + (int64_t)z * a.stride) : 0.f;   // L2: other SMs wrote it
  return ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
}
__device__ __forceinline__ float4 part_load4(const PartIn& a, int64_t b, int64_t j) {
  const float* p = a.p + b * a.ld + j;
  float4 v[kMaxSplits];
#pragma unroll
  for (int z = 0; z < kMaxSplits; z++)
    v[z] = (z < a.nz) ? __ldcg(reinterpret_cast<const float4*>(p + (int64_t)z * a.stride)) : make_float4(0.f, 0.f, 0.f, 0.f);
  float4 s;
  s.x = ((v[0].x + v[1].x) + (v[2].x + v[3].x)) + ((v[4].x + v[5].x) + (v[6].x + v[7].x));
  s.y = ((v[0].y + v[1].y) + (v[2].y + v[3].y)) + ((v[4].y + v[5].y) + (v[6].y + v[7].y));
  s.z = ((v[0].z + v[1].z) + (v[2].z + v[3].z)) + ((v[4].z + v[5].z) + (v[6].z + v[7].z));
  s.w = ((v[0].w + v[1].w) + (v[2].w + v[3].w)) + ((v[4].w + v[5].w) + (v[6].w + v[7].w));
  return s;
}
__device__ __forceinline__ void pack_store(const PackOut& o, int64_t b, int64_t j, float v) {
  if (!o.hi) return;
  __nv_bfloat16 h = __float2bfloat16_rn(v);
  o.hi[b * o.ld + j] = h;
  o.lo[b * o.ld + j] = __float2bfloat16_rn(v - __bfloat162float(h));
}
__device__ __forceinline__ void pack_store4(const PackOut& o, int64_t b, int64_t j, float4 v) {
  if (!o.hi) return;
  __nv_bfloat16 h[4], l[4];
  const float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; i++) {
    h[i] = __float2bfloat16_rn(x[i]);
    l[i] = __float2bfloat16_rn(x[i] - __bfloat162float(h[i]));
  }
  *reinterpret_cast<uint2*>(o.hi + b * o.ld + j) = *reinterpret_cast<uint2*>(h);
  *reinterpret_cast<uint2*>(o.lo + b * o.ld + j) = *reinterpret_cast<uint2*>(l);
}


__device__ __forceinline__ void cell_fwd_tc_body(CellFwdTc p, int bid, int nblk, float* sm) {
  const int H = p.H;
  const int64_t total = (int64_t)p.B * H;
  for (int64_t e = (int64_t)bid * blockDim.x + threadIdx.x; e < total; e += (int64_t)nblk * blockDim.x) {
    const int u = (int)(e % H);
    const int64_t b = e / H;
    const float* ar = p.addrows + (p.rowsel ? (int64_t)(__ldcg(p.rowsel + b) - 1) * p.addld : 0) + u;
    const float i_ = sigmoidf_(part_load(p.G, b, u) + ar[0]);
    const float f_ = sigmoidf_(part_load(p.G, b, H + u) + ar[H]);
    const float o_ = sigmoidf_(part_load(p.G, b, 2 * H + u) + ar[2 * H]);
    const float g_ = tanhf(part_load(p.G, b, 3 * H + u) + ar[3 * H]);
    const float c = f_ * p.c_prev[e] + i_ * g_;
    const float h = o_ * tanhf(c);
    p.c_new[e] = c;
    float* a = p.acts + b * 4 * H + u;
    a[0] = i_; a[H] = f_; a[2 * H] = o_; a[3 * H] = g_;
    p.h_out0[b * p.ld0 + u] = h;
    if (p.h_out1) p.h_out1[b * p.ld1 + u] = h;
    pack_store(p.pk0, b, u, h);
    pack_store(p.pk1, b, u, h);
  }
}

__device__ __forceinline__ void cell_bwd_tc_scalar(const CellBwdTc& p, int bid, int nblk) {
  const int H = p.H;
  const int64_t total = (int64_t)p.B * H;
  const int64_t stride = (int64_t)nblk * blockDim.x;
  constexpr int U = 2;      // elements per thread and pass: their loads are issued together (one L2 round trip, not U)
  for (int64_t e0 = (int64_t)bid * blockDim.x + threadIdx.x; e0 < total; e0 += stride * U) {
    float dh[U], ai[U], af[U], ao[U], ag[U], cn[U], cp[U], dcv[U];
#pragma unroll
    for (int k = 0; k < U; k++) {
      const int64_t e = e0 + k * stride;
      if (e >= total) continue;
      const int u = (int)(e % H);
      const int64_t b = e / H;
      float d = 0.f;
      if (p.dh_a.p) d += part_load(p.dh_a, b, u);
      if (p.dh_b.p) d += part_load(p.dh_b, b, u);
      if (p.dh_c.p) d += part_load(p.dh_c, b, u);
      dh[k] = d;
      const float* a = p.acts + b * 4 * H + u;
      ai[k] = __ldcg(a); af[k] = __ldcg(a + H); ao[k] = __ldcg(a + 2 * H); ag[k] = __ldcg(a + 3 * H);
      cn[k] = __ldcg(p.c_new + e); cp[k] = __ldcg(p.c_prev + e); dcv[k] = __ldcg(p.dc + e);
    }
#pragma unroll
    for (int k = 0; k < U; k++) {
      const int64_t e = e0 + k * stride;
      if (e >= total) continue;
      const int u = (int)(e % H);
      const int64_t b = e / H;
      const float i_ = ai[k], f_ = af[k], o_ = ao[k], g_ = ag[k];
      const float tc = tanhf(cn[k]);
      const float dc = dcv[k] + dh[k] * o_ * (1.f - tc * tc);
      const float d0 = dc * g_ * i_ * (1.f - i_);
      const float d1 = dc * cp[k] * f_ * (1.f - f_);
      const float d2 = dh[k] * tc * o_ * (1.f - o_);
      const float d3 = dc * i_ * (1.f - g_ * g_);
      float* dg = p.dG + b * 4 * H + u;
      dg[0] = d0; dg[H] = d1; dg[2 * H] = d2; dg[3 * H] = d3;
      pack_store(p.pk, b, u, d0);
      pack_store(p.pk, b, H + u, d1);
      pack_store(p.pk, b, 2 * H + u, d2);
      pack_store(p.pk, b, 3 * H + u, d3);
      p.dc[e] = dc * f_;
    }
  }
}

// ------------------------------------------------------------------ attention (see kernels_rnn.cu for the layout notes)
constexpr int ATT_WARPS = 8;
constexpr int ATT_MAXV = 8;
constexpr int ATT_RIF = 2;      // source rows in flight per warp and pass (3 spills registers in the executor)

// a 16-byte load of a time-invariant operand that is re-read every timestep (ctx, ctx W_c1^T): L2 evict_last, so the
// weight stream of the GEMM commands does not push it out to DRAM between two steps
__device__ __forceinline__ float4 ld_keep4(const float* p, uint64_t pol) {
  float4 v;
  asm("ld.global.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ uint64_t l2_keep_policy(bool keep = true) {
  uint64_t pol;
  if (keep) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}

// attention + output projection of batch row b (see AttnOutTc)
__device__ __forceinline__ void attn_out_tc_body(const AttnOutTc& p, int b, float* sm, bool keep_a = false) {
  const int S = p.S, H = p.H;
  float* es = sm;
  float* wm = es + ((S + 3) & ~3);
  float* wl = wm + ATT_WARPS;
  float* accs = wl + ATT_WARPS;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int nv = H / 128;
  const int bsrc = p.ctx_rows > 0 ? b % p.ctx_rows : b;
  const float* cb = p.ctx + (int64_t)bsrc * S * H;
  const float* wb = p.ctxwc + (int64_t)bsrc * S * H;
  float* qs = accs + ATT_WARPS * H;               // the summed query, shared by the 8 warps
  // the first source row of this warp does not depend on the query: its loads are issued before the partial sums and
  // the barrier below, so that L2 round trip overlaps them
  float4 row[ATT_MAXV], vrow[ATT_MAXV];
  if (warp < S) {
#pragma unroll
    for (int i = 0; i < ATT_MAXV; i++)
      if (i < nv) {
        row[i] = __ldcg(reinterpret_cast<const float4*>(cb + (int64_t)warp * H + lane * 4 + 128 * i));
        vrow[i] = __ldcg(reinterpret_cast<const float4*>(wb + (int64_t)warp * H + lane * 4 + 128 * i));
      }
  }
  // split-K partials of [q | v] are summed ONCE per CTA (4 consecutive columns per thread), not once per warp
  const int e4 = threadIdx.x * 4;
  float4 vpre = make_float4(0.f, 0.f, 0.f, 0.f);  // W_c2 h2 for the 4 outputs this thread finishes
  if (e4 < H) {
    const float4 q4 = part_load4(p.g3, b, e4);
    vpre = part_load4(p.g3, b, H + e4);
    *reinterpret_cast<float4*>(qs + e4) = q4;
    if (p.q_out) *reinterpret_cast<float4*>(p.q_out + (int64_t)b * H + e4) = q4;   // needed again by the backward
  }
  __syncthreads();
  float4 qv[ATT_MAXV], acc[ATT_MAXV];
#pragma unroll
  for (int i = 0; i < ATT_MAXV; i++) {
    acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    qv[i] = (i < nv) ? *reinterpret_cast<const float4*>(qs + lane * 4 + 128 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float m = -INFINITY, l = 0.f;
  for (int s = warp; s < S; s += ATT_WARPS) {
    if (s != warp) {
#pragma unroll
      for (int i = 0; i < ATT_MAXV; i++)
        if (i < nv) {
          row[i] = __ldcg(reinterpret_cast<const float4*>(cb + (int64_t)s * H + lane * 4 + 128 * i));
          vrow[i] = __ldcg(reinterpret_cast<const float4*>(wb + (int64_t)s * H + lane * 4 + 128 * i));
        }
    }
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < ATT_MAXV; i++)
      if (i < nv) dot += row[i].x * qv[i].x + row[i].y * qv[i].y + row[i].z * qv[i].z + row[i].w * qv[i].w;
    dot = warp_sum(dot);
    if (lane == 0) es[s] = dot;
    const float mn = fmaxf(m, dot);
    const float sc = expf(m - mn);
    const float pe = expf(dot - mn);
    l = l * sc + pe;
#pragma unroll
    for (int i = 0; i < ATT_MAXV; i++) {
      if (i < nv) {
        acc[i].x = acc[i].x * sc + pe * vrow[i].x;
        acc[i].y = acc[i].y * sc + pe * vrow[i].y;
        acc[i].z = acc[i].z * sc + pe * vrow[i].z;
        acc[i].w = acc[i].w * sc + pe * vrow[i].w;
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
  if (e4 < H) {
    float4 u = vpre;
#pragma unroll
    for (int w = 0; w < ATT_WARPS; w++) {
      const float4 t = *reinterpret_cast<const float4*>(accs + w * H + e4);
      u.x += t.x; u.y += t.y; u.z += t.z; u.w += t.w;
    }
    const float4 a = make_float4(tanhf(u.x), tanhf(u.y), tanhf(u.z), tanhf(u.w));
    *reinterpret_cast<float4*>(p.a_out + (int64_t)b * H + e4) = a;
    if (p.x_next) *reinterpret_cast<float4*>(p.x_next + (int64_t)b * p.ld_next + e4) = a;
    pack_store4(p.pk_next, b, e4, a);
    if (keep_a) *reinterpret_cast<float4*>(qs + e4) = a;      // q is dead: a_t stays in shared memory for the generator tail
  }
  for (int s = threadIdx.x; s < S; s += blockDim.x) p.alpha[(int64_t)b * S + s] = expf(es[s] - M) / L;
}

// backward of attention + output projection for batch row b (see AttnDuTc)
__device__ __forceinline__ void attn_du_tc_body(const AttnDuTc& p, int b, float* sm) {
  const int S = p.S, H = p.H;
  float* das = sm;
  float* red = das + ((S + 3) & ~3);
  float* accs = red + 16;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int nv = H / 128;
  const float* cb = p.ctx + (int64_t)b * S * H;
  const float* wb = p.ctxwc + (int64_t)b * S * H;
  const float* alpha = p.alpha + (int64_t)b * S;
  float* gs = accs + ATT_WARPS * H;               // du, shared by the 8 warps
  const int var = AOCR_VARIANT;
  const uint64_t pol = l2_keep_policy(!(var & 4));
  AOCR_BT(0);
  const int e4 = threadIdx.x * 4;
  if (e4 < H) {   // du once per CTA: kept for the time-batched weight gradients + first half of the next GEMM's operand
    const int64_t e = (int64_t)b * H + e4;
    float4 g = __ldcg(reinterpret_cast<const float4*>(p.da_gen + e));
    if (p.da_carry.p) {
      const float4 c = part_load4(p.da_carry, b, e4);
      g.x += c.x; g.y += c.y; g.z += c.z; g.w += c.w;
    }
    const float4 av = __ldcg(reinterpret_cast<const float4*>(p.a + e));
    g.x *= 1.f - av.x * av.x; g.y *= 1.f - av.y * av.y; g.z *= 1.f - av.z * av.z; g.w *= 1.f - av.w * av.w;
    *reinterpret_cast<float4*>(gs + e4) = g;
    *reinterpret_cast<float4*>(p.du_out + e) = g;
    pack_store4(p.pk, b, e4, g);
  }
  __syncthreads();
  AOCR_BT(1);
  float4 gv[ATT_MAXV];
#pragma unroll
  for (int i = 0; i < ATT_MAXV; i++)
    gv[i] = (i < nv) ? *reinterpret_cast<const float4*>(gs + lane * 4 + 128 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
  // each warp takes ATT_RIF consecutive source rows per pass, all of their loads in flight before the first use
  for (int s0 = warp * ATT_RIF; s0 < S; s0 += ATT_WARPS * ATT_RIF) {
    float4 r[ATT_RIF][ATT_MAXV];
#pragma unroll
    for (int k = 0; k < ATT_RIF; k++) {
      const int s = min(s0 + k, S - 1);
#pragma unroll
      for (int i = 0; i < ATT_MAXV; i++)
        if (i < nv) {
          const float* src = wb + (int64_t)s * H + lane * 4 + 128 * i;
          r[k][i] = ld_keep4(src, pol);
        }
    }
#pragma unroll
    for (int k = 0; k < ATT_RIF; k++) {
      float dot = 0.f;
#pragma unroll
      for (int i = 0; i < ATT_MAXV; i++)
        if (i < nv) dot += r[k][i].x * gv[i].x + r[k][i].y * gv[i].y + r[k][i].z * gv[i].z + r[k][i].w * gv[i].w;
      dot = warp_sum(dot);
      if (lane == 0 && s0 + k < S) das[s0 + k] = dot;
    }
  }
  __syncthreads();
  AOCR_BT(2);
  if (warp == 0) {
    float s_ = 0.f;
    for (int s = lane; s < S; s += 32) s_ += alpha[s] * das[s];
    s_ = warp_sum(s_);
    if (lane == 0) red[0] = s_;
  }
  __syncthreads();
  AOCR_BT(3);
  const float tot = red[0];
  for (int s = threadIdx.x; s < S; s += blockDim.x) p.de[(int64_t)b * S + s] = alpha[s] * (das[s] - tot);
  float4 acc[ATT_MAXV];
#pragma unroll
  for (int i = 0; i < ATT_MAXV; i++) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s0 = warp * ATT_RIF; s0 < S; s0 += ATT_WARPS * ATT_RIF) {
    float4 r[ATT_RIF][ATT_MAXV];
    float w[ATT_RIF];
#pragma unroll
    for (int k = 0; k < ATT_RIF; k++) {
      const int s = min(s0 + k, S - 1);
      w[k] = s0 + k < S ? alpha[s] * (das[s] - tot) : 0.f;
#pragma unroll
      for (int i = 0; i < ATT_MAXV; i++)
        if (i < nv) {
          const float* src = cb + (int64_t)s * H + lane * 4 + 128 * i;
          r[k][i] = ld_keep4(src, pol);
        }
    }
#pragma unroll
    for (int k = 0; k < ATT_RIF; k++)
#pragma unroll
      for (int i = 0; i < ATT_MAXV; i++)
        if (i < nv) {
          acc[i].x = fmaf(w[k], r[k][i].x, acc[i].x); acc[i].y = fmaf(w[k], r[k][i].y, acc[i].y);
          acc[i].z = fmaf(w[k], r[k][i].z, acc[i].z); acc[i].w = fmaf(w[k], r[k][i].w, acc[i].w);
        }
  }
#pragma unroll
  for (int i = 0; i < ATT_MAXV; i++)
    if (i < nv) *reinterpret_cast<float4*>(accs + warp * H + lane * 4 + 128 * i) = acc[i];
  __syncthreads();
  AOCR_BT(4);
  if (e4 < H) {
    float4 u = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int w = 0; w < ATT_WARPS; w++) {
      const float4 t = *reinterpret_cast<const float4*>(accs + w * H + e4);
      u.x += t.x; u.y += t.y; u.z += t.z; u.w += t.w;
    }
    *reinterpret_cast<float4*>(p.dq + (int64_t)b * H + e4) = u;
    pack_store4(p.pk, b, H + e4, u);
  }
  AOCR_BT(5);
}

// encoder cell, both directions (grid-stride over dir x batch x unit); slot convention of engine.cu
__device__ __forceinline__ void enc_cell_fwd_tc_body(EncCellFwdTc p, int bid, int nblk, float* sm) {
  const int He = p.He, B = p.B, S = p.S;
  const int64_t total = (int64_t)(p.d_only < 0 ? 2 : 1) * B * He;
  for (int64_t e = (int64_t)bid * blockDim.x + threadIdx.x; e < total; e += (int64_t)nblk * blockDim.x) {
    const int unit = (int)(e % He);
    const int64_t b = (e / He) % B;
    const int d = p.d_only < 0 ? (int)(e / ((int64_t)He * B)) : p.d_only;
    const int t = d == 0 ? p.step : S - 1 - p.step;
    const int prev_slot = d == 0 ? t : t + 1, out_slot = d == 0 ? t + 1 : t;
    const float* xg = p.xg + ((int64_t)t * B + b) * (8 * He) + (int64_t)d * 4 * He + unit;
    const PartIn& G = p.G[d];
    const float i_ = sigmoidf_(part_load(G, b, unit) + xg[0]);
    const float f_ = sigmoidf_(part_load(G, b, He + unit) + xg[He]);
    const float o_ = sigmoidf_(part_load(G, b, 2 * He + unit) + xg[2 * He]);
    const float g_ = tanhf(part_load(G, b, 3 * He + unit) + xg[3 * He]);
    const float cp = p.Cst[((int64_t)(d * (S + 1) + prev_slot) * B + b) * He + unit];
    const float c = f_ * cp + i_ * g_;
    const float h = o_ * tanhf(c);
    p.Cst[((int64_t)(d * (S + 1) + out_slot) * B + b) * He + unit] = c;
    p.H[((int64_t)(d * (S + 1) + out_slot) * B + b) * He + unit] = h;
    float* a = p.acts + (((int64_t)(d * S + t) * B + b) * 4) * He + unit;
    a[0] = i_; a[He] = f_; a[2 * He] = o_; a[3 * He] = g_;
    p.ctx[((int64_t)b * S + t) * (2 * He) + d * He + unit] = h;
    pack_store(p.hp[d], b, unit, h);
  }
}

__device__ __forceinline__ void enc_cell_bwd_tc_scalar(const EncCellBwdTc& p, int bid, int nblk) {
  const int He = p.He, B = p.B, S = p.S;
  const int64_t total = (int64_t)(p.d_only < 0 ? 2 : 1) * B * He;
  const int64_t stride = (int64_t)nblk * blockDim.x;
  constexpr int U = 4;      // elements per thread and pass, loads of all U in front (32 CTAs serve 64 x 512 elements)
  for (int64_t e0 = (int64_t)bid * blockDim.x + threadIdx.x; e0 < total; e0 += stride * U) {
    float dhv[U], ai[U], af[U], ao[U], ag[U], cpv[U], cnv[U], dcv[U];
#pragma unroll
    for (int k = 0; k < U; k++) {
      const int64_t e = e0 + k * stride;
      if (e >= total) continue;
      const int unit = (int)(e % He);
      const int64_t b = (e / He) % B;
      const int d = p.d_only < 0 ? (int)(e / ((int64_t)He * B)) : p.d_only;
      const int t = d == 0 ? S - 1 - p.step : p.step;
      const int prev_slot = d == 0 ? t : t + 1, out_slot = d == 0 ? t + 1 : t;
      const float* a = p.acts + (((int64_t)(d * S + t) * B + b) * 4) * He + unit;
      ai[k] = __ldcg(a); af[k] = __ldcg(a + He); ao[k] = __ldcg(a + 2 * He); ag[k] = __ldcg(a + 3 * He);
      cpv[k] = __ldcg(p.Cst + ((int64_t)(d * (S + 1) + prev_slot) * B + b) * He + unit);
      cnv[k] = __ldcg(p.Cst + ((int64_t)(d * (S + 1) + out_slot) * B + b) * He + unit);
      dhv[k] = part_load(p.dh[d], b, unit) + __ldcg(p.Dctx + ((int64_t)b * S + t) * (2 * He) + d * He + unit);
      dcv[k] = __ldcg(p.dc + ((int64_t)d * B + b) * He + unit);
    }
#pragma unroll
    for (int k = 0; k < U; k++) {
      const int64_t e = e0 + k * stride;
      if (e >= total) continue;
      const int unit = (int)(e % He);
      const int64_t b = (e / He) % B;
      const int d = p.d_only < 0 ? (int)(e / ((int64_t)He * B)) : p.d_only;
      const int t = d == 0 ? S - 1 - p.step : p.step;
      const float i_ = ai[k], f_ = af[k], o_ = ao[k], g_ = ag[k];
      const float tc = tanhf(cnv[k]);
      const float dh = dhv[k];
      const int64_t ce = ((int64_t)d * B + b) * He + unit;
      const float dc = dcv[k] + dh * o_ * (1.f - tc * tc);
      const float d0 = dc * g_ * i_ * (1.f - i_);
      const float d1 = dc * cpv[k] * f_ * (1.f - f_);
      const float d2 = dh * tc * o_ * (1.f - o_);
      const float d3 = dc * i_ * (1.f - g_ * g_);
      float* dg = p.dG + ((int64_t)t * B + b) * (8 * He) + (int64_t)d * 4 * He + unit;
      dg[0] = d0; dg[He] = d1; dg[2 * He] = d2; dg[3 * He] = d3;
      pack_store(p.dgp[d], b, unit, d0);
      pack_store(p.dgp[d], b, He + unit, d1);
      pack_store(p.dgp[d], b, 2 * He + unit, d2);
      pack_store(p.dgp[d], b, 3 * He + unit, d3);
      p.dc[ce] = dc * f_;
    }
  }
}

// ---- vector forms of the two cell-backward bodies: one thread = 4 consecutive hidden units of one batch row, every
// operand as one 16-byte access (a quarter of the memory instructions of the scalar form; these bodies are bound by the
// number of outstanding requests, not by bytes).  Used when every base pointer / pitch keeps 16-byte alignment.
__device__ __forceinline__ bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
__device__ __forceinline__ bool part_al16(const PartIn& a) { return !a.p || (al16(a.p) && (a.ld & 3) == 0 && (a.stride & 3) == 0); }
__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
struct GateGrads { float4 d0, d1, d2, d3, dcn; };
__device__ __forceinline__ GateGrads lstm_cell_bwd4(float4 i_, float4 f_, float4 o_, float4 g_, float4 cn, float4 cp, float4 dcin,
                                                    float4 dh) {
  GateGrads r;
  const float iv[4] = {i_.x, i_.y, i_.z, i_.w}, fv[4] = {f_.x, f_.y, f_.z, f_.w}, ov[4] = {o_.x, o_.y, o_.z, o_.w};
  const float gv[4] = {g_.x, g_.y, g_.z, g_.w}, cnv[4] = {cn.x, cn.y, cn.z, cn.w}, cpv[4] = {cp.x, cp.y, cp.z, cp.w};
  const float dcv[4] = {dcin.x, dcin.y, dcin.z, dcin.w}, dhv[4] = {dh.x, dh.y, dh.z, dh.w};
  float d0[4], d1[4], d2[4], d3[4], dn[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const float tc = tanhf(cnv[k]);
    const float dc = dcv[k] + dhv[k] * ov[k] * (1.f - tc * tc);
    d0[k] = dc * gv[k] * iv[k] * (1.f - iv[k]);
    d1[k] = dc * cpv[k] * fv[k] * (1.f - fv[k]);
    d2[k] = dhv[k] * tc * ov[k] * (1.f - ov[k]);
    d3[k] = dc * iv[k] * (1.f - gv[k] * gv[k]);
    dn[k] = dc * fv[k];
  }
  r.d0 = make_float4(d0[0], d0[1], d0[2], d0[3]); r.d1 = make_float4(d1[0], d1[1], d1[2], d1[3]);
  r.d2 = make_float4(d2[0], d2[1], d2[2], d2[3]); r.d3 = make_float4(d3[0], d3[1], d3[2], d3[3]);
  r.dcn = make_float4(dn[0], dn[1], dn[2], dn[3]);
  return r;
}
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

__device__ __forceinline__ void cell_bwd_tc_body(CellBwdTc p, int bid, int nblk, float* sm) {
  const int H = p.H;
  const bool vec = (H & 3) == 0 && part_al16(p.dh_a) && part_al16(p.dh_b) && part_al16(p.dh_c) && al16(p.dc) && al16(p.c_prev) &&
                   al16(p.c_new) && al16(p.acts) && al16(p.dG) && (!p.pk.hi || ((p.pk.ld & 3) == 0));
  if (!vec) { cell_bwd_tc_scalar(p, bid, nblk); return; }
  const int H4 = H / 4;
  const int64_t total = (int64_t)p.B * H4;
  for (int64_t q = (int64_t)bid * blockDim.x + threadIdx.x; q < total; q += (int64_t)nblk * blockDim.x) {
    const int u = (int)(q % H4) * 4;
    const int64_t b = q / H4, e = b * H + u;
    const float* a = p.acts + b * 4 * H + u;
    const float4 ai = ldcg4(a), af = ldcg4(a + H), ao = ldcg4(a + 2 * H), ag = ldcg4(a + 3 * H);
    const float4 cn = ldcg4(p.c_new + e), cp = ldcg4(p.c_prev + e), dcin = ldcg4(p.dc + e);
    float4 dh = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.dh_a.p) dh = add4(dh, part_load4(p.dh_a, b, u));
    if (p.dh_b.p) dh = add4(dh, part_load4(p.dh_b, b, u));
    if (p.dh_c.p) dh = add4(dh, part_load4(p.dh_c, b, u));
    const GateGrads g = lstm_cell_bwd4(ai, af, ao, ag, cn, cp, dcin, dh);
    float* dg = p.dG + b * 4 * H + u;
    *reinterpret_cast<float4*>(dg) = g.d0; *reinterpret_cast<float4*>(dg + H) = g.d1;
    *reinterpret_cast<float4*>(dg + 2 * H) = g.d2; *reinterpret_cast<float4*>(dg + 3 * H) = g.d3;
    pack_store4(p.pk, b, u, g.d0); pack_store4(p.pk, b, H + u, g.d1);
    pack_store4(p.pk, b, 2 * H + u, g.d2); pack_store4(p.pk, b, 3 * H + u, g.d3);
    *reinterpret_cast<float4*>(p.dc + e) = g.dcn;
  }
}

__device__ __forceinline__ void enc_cell_bwd_tc_body(EncCellBwdTc p, int bid, int nblk, float* sm) {
  const int He = p.He, B = p.B, S = p.S;
  const bool vec = (He & 3) == 0 && part_al16(p.dh[0]) && part_al16(p.dh[1]) && al16(p.Cst) && al16(p.acts) && al16(p.Dctx) &&
                   al16(p.dc) && al16(p.dG) && (!p.dgp[0].hi || (p.dgp[0].ld & 3) == 0) && (!p.dgp[1].hi || (p.dgp[1].ld & 3) == 0);
  if (!vec) { enc_cell_bwd_tc_scalar(p, bid, nblk); return; }
  const int H4 = He / 4;
  const int64_t total = (int64_t)(p.d_only < 0 ? 2 : 1) * B * H4;
  AOCR_BT(8);
  for (int64_t q = (int64_t)bid * blockDim.x + threadIdx.x; q < total; q += (int64_t)nblk * blockDim.x) {
    const int unit = (int)(q % H4) * 4;
    const int64_t b = (q / H4) % B;
    const int d = p.d_only < 0 ? (int)(q / ((int64_t)H4 * B)) : p.d_only;
    const int t = d == 0 ? S - 1 - p.step : p.step;
    const int prev_slot = d == 0 ? t : t + 1, out_slot = d == 0 ? t + 1 : t;
    const float* a = p.acts + (((int64_t)(d * S + t) * B + b) * 4) * He + unit;
    const float4 ai = ldcg4(a), af = ldcg4(a + He), ao = ldcg4(a + 2 * He), ag = ldcg4(a + 3 * He);
    const float4 cp = ldcg4(p.Cst + ((int64_t)(d * (S + 1) + prev_slot) * B + b) * He + unit);
    const float4 cn = ldcg4(p.Cst + ((int64_t)(d * (S + 1) + out_slot) * B + b) * He + unit);
    const int64_t ce = ((int64_t)d * B + b) * He + unit;
    const float4 dcin = ldcg4(p.dc + ce);
    const float4 dh = add4(part_load4(p.dh[d], b, unit), ldcg4(p.Dctx + ((int64_t)b * S + t) * (2 * He) + d * He + unit));
    AOCR_BT(9);
    const GateGrads g = lstm_cell_bwd4(ai, af, ao, ag, cn, cp, dcin, dh);
    float* dg = p.dG + ((int64_t)t * B + b) * (8 * He) + (int64_t)d * 4 * He + unit;
    *reinterpret_cast<float4*>(dg) = g.d0; *reinterpret_cast<float4*>(dg + He) = g.d1;
    *reinterpret_cast<float4*>(dg + 2 * He) = g.d2; *reinterpret_cast<float4*>(dg + 3 * He) = g.d3;
    pack_store4(p.dgp[d], b, unit, g.d0); pack_store4(p.dgp[d], b, He + unit, g.d1);
    pack_store4(p.dgp[d], b, 2 * He + unit, g.d2); pack_store4(p.dgp[d], b, 3 * He + unit, g.d3);
    *reinterpret_cast<float4*>(p.dc + ce) = g.dcn;
    AOCR_BT(10);
  }
}

__device__ __forceinline__ void part_to_dense_body(PartIn in, float* dst, int64_t ld, int B, int cols, int bid, int nblk, float* sm) {
  const int64_t total = (int64_t)B * cols;
  for (int64_t e = (int64_t)bid * blockDim.x + threadIdx.x; e < total; e += (int64_t)nblk * blockDim.x) {
    const int j = (int)(e % cols);
    const int64_t b = e / cols;
    dst[b * ld + j] = part_load(in, b, j);
  }
}


// ---- generator (Linear(H,V) + log-softmax [+ NLL / dlogits]) for the rows r = bid, bid+nblk, ... ; 8 warps per row
__device__ __forceinline__ float warp_max_(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ void generator_body(const GenTc& p, int bid, int nblk, float* sm) {
  float* zs = sm;                                   // V <= 64 logits
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32, nw = blockDim.x / 32;
  const int nv = p.H / 128, V = p.V, H = p.H;
  for (int64_t r = bid; r < p.R; r += nblk) {
    float4 av[ATT_MAXV];
#pragma unroll
    for (int i = 0; i < ATT_MAXV; i++)
      av[i] = (i < nv) ? __ldcg(reinterpret_cast<const float4*>(p.a + r * H + lane * 4 + 128 * i)) : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int v = warp; v < V; v += nw) {
      float dot = 0.f;
#pragma unroll
      for (int i = 0; i < ATT_MAXV; i++) {
        if (i < nv) {
          float4 w = *reinterpret_cast<const float4*>(p.W + (int64_t)v * H + lane * 4 + 128 * i);
          dot += w.x * av[i].x + w.y * av[i].y + w.z * av[i].z + w.w * av[i].w;
        }
      }
      dot = warp_sum(dot);
      if (lane == 0) zs[v] = dot + p.bias[v];
    }
    __syncthreads();
    if (warp == 0) {
      const float z0 = lane < V ? zs[lane] : -INFINITY;
      const float z1 = lane + 32 < V ? zs[lane + 32] : -INFINITY;
      const float mx = warp_max_(fmaxf(z0, z1));
      float se = (lane < V ? expf(z0 - mx) : 0.f) + (lane + 32 < V ? expf(z1 - mx) : 0.f);
      se = warp_sum(se);
      const float lse = mx + logf(se);
      const bool second = p.split > 0 && r >= p.split;           // teacher-forced half of a dual pass
      const int64_t ro = second ? r - p.split : r;
      const bool has_y = p.y && (p.split == 0 || second);
      const int yy = has_y ? p.y[ro] - 1 : -1;
      const float w = (has_y && yy != 0) ? 1.f : 0.f;
      float* lpo = second ? p.logp2 : p.logp;
#pragma unroll
      for (int j = 0; j < 2; j++) {
        const int v = lane + 32 * j;
        if (v < V) {
          const float lp = (j == 0 ? z0 : z1) - lse;
          lpo[ro * V + v] = lp;
          if (p.dz) p.dz[r * V + v] = (expf(lp) - (v == yy ? 1.f : 0.f)) * w * p.inv_bn;
          if (p.rowloss && has_y && v == yy) p.rowloss[ro] = -w * lp;
        }
      }
    }
    __syncthreads();
  }
}
// Generator + selection for ONE batch row r whose a_t is in shared memory (`as`): logits, log-softmax, then either the
// greedy selection (rows < g.split: sticky PAD, argmax with lowest-index ties, score, next token; model.lua:393-404,
// 446-459) or the teacher-forced row loss (rows >= g.split).  Tail of the dual decode pass: no grid barrier between
// attention+output, generator and selection, all three are local to the CTA that owns the row.
__device__ __forceinline__ void gen_select_tail(const GenTc& g, const GreedyTc& s, int r, const float* as, float* zs) {
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32, nw = blockDim.x / 32;
  const int nv = g.H / 128, V = g.V, H = g.H;
  __syncthreads();                                   // a_t complete in shared memory
  float4 av[ATT_MAXV];
#pragma unroll
  for (int i = 0; i < ATT_MAXV; i++)
    av[i] = (i < nv) ? *reinterpret_cast<const float4*>(as + lane * 4 + 128 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int v = warp; v < V; v += nw) {
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < ATT_MAXV; i++) {
      if (i < nv) {
        const float4 w = *reinterpret_cast<const float4*>(g.W + (int64_t)v * H + lane * 4 + 128 * i);
        dot += w.x * av[i].x + w.y * av[i].y + w.z * av[i].z + w.w * av[i].w;
      }
    }
    dot = warp_sum(dot);
    if (lane == 0) zs[v] = dot + g.bias[v];
  }
  __syncthreads();
  if (warp == 0) {
    const float z0 = lane < V ? zs[lane] : -INFINITY;
    const float z1 = lane + 32 < V ? zs[lane + 32] : -INFINITY;
    const float mx = warp_max_(fmaxf(z0, z1));
    float se = (lane < V ? expf(z0 - mx) : 0.f) + (lane + 32 < V ? expf(z1 - mx) : 0.f);
    se = warp_sum(se);
    const float lse = mx + logf(se);
    float lp0 = z0 - lse, lp1 = z1 - lse;            // log-probs of v = lane, lane + 32
    const bool second = r >= g.split;
    const int64_t ro = second ? r - g.split : r;
    float* lpo = second ? g.logp2 : g.logp;
    if (second) {
      const int yy = g.y[ro] - 1;
      const float w = yy != 0 ? 1.f : 0.f;
      if (lane < V) lpo[ro * V + lane] = lp0;
      if (lane + 32 < V) lpo[ro * V + lane + 32] = lp1;
      if (lane == yy) g.rowloss[ro] = -w * lp0;
      if (lane + 32 == yy) g.rowloss[ro] = -w * lp1;
    } else {
      if (s.t > 0 && lane == 0) {
        const int prev = __ldcg(s.tok + r);
        if (prev == 1 || prev == 3) lp0 = 0.f;       // sticky PAD: log-prob[PAD] <- 0 (model.lua:448-449)
      }
      if (lane < V) lpo[ro * V + lane] = lp0;
      if (lane + 32 < V) lpo[ro * V + lane + 32] = lp1;
      // argmax, ties to the lowest index (the reference's max returns the first maximum)
      float best = lane < V ? lp0 : -INFINITY;
      int bi = lane;
      if (lane + 32 < V && lp1 > best) { best = lp1; bi = lane + 32; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      if (lane == 0) {
        s.score[r] = (s.t == 0 ? 0.0 : __ldcg(s.score + r)) + (double)best;
        s.tok_out[r] = bi + 1;
        s.labels[(int64_t)r * s.ldl + s.t] = bi + 1;
      }
    }
  }
}
// sticky-PAD edit, argmax, score accumulate, next token (model.lua:402,448-458); one thread per batch row
__device__ __forceinline__ void greedy_select_body(const GreedyTc& p, int bid, int nblk, float* sm) {
  for (int b = bid * blockDim.x + threadIdx.x; b < p.B; b += nblk * blockDim.x) {
    float* lp = p.logp + (int64_t)b * p.V;
    float first = __ldcg(lp);
    if (p.t > 0) {
      const int prev = __ldcg(p.tok + b);
      if (prev == 1 || prev == 3) { first = 0.f; lp[0] = 0.f; }
    }
    float best = first;
    int bi = 0;
    for (int v = 1; v < p.V; v++) {
      const float x = __ldcg(lp + v);
      if (x > best) { best = x; bi = v; }
    }
    p.score[b] = (p.t == 0 ? 0.0 : __ldcg(p.score + b)) + (double)best;
    p.tok_out[b] = bi + 1;
    p.labels[(int64_t)b * p.ldl + p.t] = bi + 1;
  }
}

}  // namespace decb
}  // namespace aocr
