// Micro-benchmark (development tool): how fast can one CTA per batch row ingest its (S x H) fp32 ctx block?
// Variants of the attention bodies' access pattern, timed per CTA with %globaltimer (max over CTAs), one 256-thread CTA
// per SM (large dynamic shared memory forces 1 CTA/SM, as in the executor).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o attn_pattern attn_pattern.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int H = 1024;
constexpr int NV = H / 128;

// V0: the current body: warp per row, rows s = warp, warp+8, ...; one row's 8 float4 in flight per lane
template <int ROWS_IN_FLIGHT>
__global__ void __launch_bounds__(256, 1) k_ldg(const float* __restrict__ ctx, const float* __restrict__ g, float* __restrict__ out,
                                               unsigned long long* __restrict__ tms, int S, int rows_per_cta_div) {
  extern __shared__ float sm[];
  const int b = blockIdx.x / rows_per_cta_div, part = blockIdx.x % rows_per_cta_div;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* cb = ctx + (int64_t)b * S * H;
  float4 gv[NV];
#pragma unroll
  for (int i = 0; i < NV; i++) gv[i] = *reinterpret_cast<const float4*>(g + b * H + lane * 4 + 128 * i);
  __syncthreads();
  const unsigned long long t0 = gtime();
  // rows of this CTA: s = part + rows_per_cta_div * j
  const int nrows = (S - part + rows_per_cta_div - 1) / rows_per_cta_div;
  for (int j0 = warp * ROWS_IN_FLIGHT; j0 < nrows; j0 += 8 * ROWS_IN_FLIGHT) {
    float4 r[ROWS_IN_FLIGHT][NV];
#pragma unroll
    for (int k = 0; k < ROWS_IN_FLIGHT; k++) {
      const int j = j0 + k;
      const int s = part + rows_per_cta_div * (j < nrows ? j : nrows - 1);
#pragma unroll
      for (int i = 0; i < NV; i++) r[k][i] = __ldcg(reinterpret_cast<const float4*>(cb + (int64_t)s * H + lane * 4 + 128 * i));
    }
#pragma unroll
    for (int k = 0; k < ROWS_IN_FLIGHT; k++) {
      float dot = 0.f;
#pragma unroll
      for (int i = 0; i < NV; i++) dot += r[k][i].x * gv[i].x + r[k][i].y * gv[i].y + r[k][i].z * gv[i].z + r[k][i].w * gv[i].w;
      dot = warp_sum(dot);
      if (lane == 0 && j0 + k < nrows) sm[j0 + k] = dot;
    }
  }
  __syncthreads();
  const unsigned long long t1 = gtime();
  if (threadIdx.x < nrows) out[(int64_t)blockIdx.x * 256 + threadIdx.x] = sm[threadIdx.x];
  if (threadIdx.x == 0) tms[blockIdx.x] = t1 - t0;
}

// V2: bulk async copies (cp.async.bulk.shared::cluster.global.mbarrier) of the CTA's rows into shared memory in chunks,
// warps consume chunk by chunk
__global__ void __launch_bounds__(256, 1) k_bulk(const float* __restrict__ ctx, const float* __restrict__ g, float* __restrict__ out,
                                                unsigned long long* __restrict__ tms, int S, int chunk_rows) {
  extern __shared__ __align__(128) uint8_t smraw[];
  float* stage = reinterpret_cast<float*>(smraw + 1024);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smraw);
  float* res = reinterpret_cast<float*>(smraw + 512);
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* cb = ctx + (int64_t)b * S * H;
  const int nchunks = (S + chunk_rows - 1) / chunk_rows;
  if (threadIdx.x == 0) {
    for (int c = 0; c < nchunks; c++)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bars + c)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  float4 gv[NV];
#pragma unroll
  for (int i = 0; i < NV; i++) gv[i] = *reinterpret_cast<const float4*>(g + b * H + lane * 4 + 128 * i);
  __syncthreads();
  const unsigned long long t0 = gtime();
  if (threadIdx.x == 0) {
    for (int c = 0; c < nchunks; c++) {
      const int r0 = c * chunk_rows, nr = min(chunk_rows, S - r0);
      const uint32_t bytes = (uint32_t)nr * H * 4;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bars + c)), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_u32(stage + (int64_t)r0 * H)),
                   "l"(cb + (int64_t)r0 * H), "r"(bytes), "r"(smem_u32(bars + c))
                   : "memory");
    }
  }
  for (int c = 0; c < nchunks; c++) {
    // wait for chunk c
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                   : "=r"(ok) : "r"(smem_u32(bars + c)), "r"(0u) : "memory");
    }
    const int r0 = c * chunk_rows, nr = min(chunk_rows, S - r0);
    for (int s = r0 + warp; s < r0 + nr; s += 8) {
      float dot = 0.f;
#pragma unroll
      for (int i = 0; i < NV; i++) {
        const float4 r = *reinterpret_cast<const float4*>(stage + (int64_t)s * H + lane * 4 + 128 * i);
        dot += r.x * gv[i].x + r.y * gv[i].y + r.z * gv[i].z + r.w * gv[i].w;
      }
      dot = warp_sum(dot);
      if (lane == 0) res[s] = dot;
    }
  }
  __syncthreads();
  const unsigned long long t1 = gtime();
  if (threadIdx.x < S) out[(int64_t)blockIdx.x * 256 + threadIdx.x] = res[threadIdx.x];
  if (threadIdx.x == 0) tms[blockIdx.x] = t1 - t0;
}

int main(int argc, char** argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 64;
  const int S = argc > 2 ? atoi(argv[2]) : 24;
  float *ctx, *g, *out, *flush;
  unsigned long long* tms;
  const size_t nctx = (size_t)B * S * H;
  CK(cudaMalloc(&ctx, nctx * 4)); CK(cudaMalloc(&g, (size_t)B * H * 4)); CK(cudaMalloc(&out, (size_t)2 * B * 256 * 4));
  CK(cudaMalloc(&tms, 2 * B * 8)); CK(cudaMalloc(&flush, 256u << 20));
  std::vector<float> h(nctx);
  for (size_t i = 0; i < nctx; i++) h[i] = (float)((i * 2654435761u) % 1000) * 1e-3f;
  CK(cudaMemcpy(ctx, h.data(), nctx * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(g, h.data(), (size_t)B * H * 4, cudaMemcpyHostToDevice));
  const int smem = 200 * 1024;
  CK(cudaFuncSetAttribute(k_ldg<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(k_ldg<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(k_ldg<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  auto report = [&](const char* name, int nblk) {
    CK(cudaDeviceSynchronize());
    std::vector<unsigned long long> t(nblk);
    CK(cudaMemcpy(t.data(), tms, nblk * 8, cudaMemcpyDeviceToHost));
    std::sort(t.begin(), t.end());
    printf("%-44s CTAs %3d: median %.2f us  max %.2f us\n", name, nblk, t[nblk / 2] / 1e3, t[nblk - 1] / 1e3);
  };
  for (int cold = 0; cold < 2; cold++) {
    printf("---- B=%d S=%d H=%d  %s\n", B, S, H, cold ? "cold (L2 flushed before each launch)" : "warm (second launch)");
    auto pre = [&]() { if (cold) CK(cudaMemset(flush, 1, 256u << 20)); };
    for (int rep = 0; rep < 2; rep++) { pre(); k_ldg<1><<<B, 256, smem>>>(ctx, g, out, tms, S, 1); }
    report("ldg, 1 row in flight per warp (current)", B);
    for (int rep = 0; rep < 2; rep++) { pre(); k_ldg<3><<<B, 256, smem>>>(ctx, g, out, tms, S, 1); }
    report("ldg, 3 rows in flight per warp", B);
    for (int rep = 0; rep < 2; rep++) { pre(); k_ldg<1><<<2 * B, 256, smem>>>(ctx, g, out, tms, S, 2); }
    report("ldg, 1 row in flight, 2 CTAs per batch row", 2 * B);
    for (int rep = 0; rep < 2; rep++) { pre(); k_ldg<2><<<2 * B, 256, smem>>>(ctx, g, out, tms, S, 2); }
    report("ldg, 2 rows in flight, 2 CTAs per batch row", 2 * B);
    for (int cr : {4, 8, 24}) {
      for (int rep = 0; rep < 2; rep++) { pre(); k_bulk<<<B, 256, smem>>>(ctx, g, out, tms, S, cr); }
      char nm[64]; snprintf(nm, 64, "bulk copy -> smem, chunks of %d rows", cr);
      report(nm, B);
    }
  }
  return 0;
}
