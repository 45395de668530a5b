// persist.cu — persistent recurrence executor (see persist.h).  256 threads per CTA, one CTA per SM (197 KB of shared
// memory for the TMA ring): warp 0 = TMA producer, warp 1 = tcgen05 issuer and TMEM owner, warps 2-5 = epilogue (TMEM ->
// split-K partials in global memory, or -> this CTA's shared memory for the fused GEMM -> cell commands), all 8 warps =
// the cell / attention bodies and the cluster-side cell of the fused commands.
// mbarrier phases and the TMEM allocation persist across commands; a grid barrier (one global counter, acquire /
// release, preceded by fence.proxy.async so the generic-proxy stores of a command are visible to the TMA loads of
// the next one on every SM) replaces the kernel boundary between consecutive commands.
#include <algorithm>
#include <type_traits>

#include <cooperative_groups.h>

#include "persist.h"

#include "dec_bodies.cuh"
#include "tc_ptx.cuh"

namespace aocr {

namespace {
using namespace tcp;
namespace cg = cooperative_groups;

template <int BN> struct PCfg {
  static constexpr int kStages = (BN == 128) ? 3 : 4;
  static constexpr int kBPlane = BN * BK * 2;
  static constexpr int kStageBytes = 2 * A_PLANE_BYTES + 2 * kBPlane;
  static constexpr int kTmemCols = BN < 32 ? 32 : BN;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
};

__device__ __forceinline__ void grid_sync(unsigned* ctr, unsigned nblk, unsigned& epoch) {
  asm volatile("fence.proxy.async;" ::: "memory");      // this thread's stores -> visible to the async proxy (TMA)
  __syncthreads();
  epoch++;
  if (threadIdx.x == 0) {
    // arrive: a release reduction (no return value, so no round trip before the polling starts); the release is
    // cumulative over the CTA's writes ordered before it by the bar.sync above
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(ctr), "r"(1u) : "memory");
    const unsigned target = epoch * nblk;
    unsigned v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    } while (v < target);
  }
  __syncthreads();
}

// ---- fused GEMM -> cell: the cluster owns 32 hidden units x 4 gates (gate-interleaved weight rows: tile row =
// gate*32 + unit) for the whole batch; rank z holds the split-K partial z in shared memory as stage[column][row].
// Rank z finishes batch columns [z*BN/nz, (z+1)*BN/nz): sums the nz partials of the 4 gates through DSMEM, applies the
// LSTM cell and writes what the stand-alone cell body writes.
template <int BN>
__device__ __forceinline__ void fused_enc_cell_fwd(const EncCellFwdTc& p, cg::cluster_group& cluster, float* stage, int mt,
                                                   int z, int nz) {
  const int He = p.He, B = p.B, S = p.S;
  const int d = p.d_only;
  const int ul = threadIdx.x & 31, unit = mt * 32 + ul;
  const int cols = BN / nz;
  const int t = d == 0 ? p.step : S - 1 - p.step;
  const int prev_slot = d == 0 ? t : t + 1, out_slot = d == 0 ? t + 1 : t;
  const float* rs[8];
#pragma unroll
  for (int q = 0; q < 8; q++) rs[q] = q < nz ? cluster.map_shared_rank(stage, q) : stage;
  for (int bq = threadIdx.x >> 5; bq < cols; bq += 8) {
    const int b = z * cols + bq;
    if (b >= B || unit >= He) continue;
    float gsum[4];
#pragma unroll
    for (int g = 0; g < 4; g++) {
      float v[8];
#pragma unroll
      for (int q = 0; q < 8; q++) v[q] = q < nz ? rs[q][b * BM + g * 32 + ul] : 0.f;
      gsum[g] = ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
    }
    const float* xg = p.xg + ((int64_t)t * B + b) * (8 * He) + (int64_t)d * 4 * He + unit;
    const float i_ = decb::sigmoidf_(gsum[0] + xg[0]);
    const float f_ = decb::sigmoidf_(gsum[1] + xg[He]);
    const float o_ = decb::sigmoidf_(gsum[2] + xg[2 * He]);
    const float g_ = tanhf(gsum[3] + xg[3 * He]);
    const float cp = p.Cst[((int64_t)(d * (S + 1) + prev_slot) * B + b) * He + unit];
    const float c = f_ * cp + i_ * g_;
    const float h = o_ * tanhf(c);
    p.Cst[((int64_t)(d * (S + 1) + out_slot) * B + b) * He + unit] = c;
    p.H[((int64_t)(d * (S + 1) + out_slot) * B + b) * He + unit] = h;
    float* a = p.acts + (((int64_t)(d * S + t) * B + b) * 4) * He + unit;
    a[0] = i_; a[He] = f_; a[2 * He] = o_; a[3 * He] = g_;
    p.ctx[((int64_t)b * S + t) * (2 * He) + d * He + unit] = h;
    decb::pack_store(p.hp[d], b, unit, h);
  }
}

// decoder layer: same structure, CellFwdTc semantics (embedding / bias rows added per batch row)
template <int BN>
__device__ __forceinline__ void fused_cell_fwd(const CellFwdTc& p, cg::cluster_group& cluster, float* stage, int mt, int z,
                                               int nz) {
  const int H = p.H;
  const int ul = threadIdx.x & 31, u = mt * 32 + ul;
  const int cols = BN / nz;
  const float* rs[8];
#pragma unroll
  for (int q = 0; q < 8; q++) rs[q] = q < nz ? cluster.map_shared_rank(stage, q) : stage;
  for (int bq = threadIdx.x >> 5; bq < cols; bq += 8) {
    const int b = z * cols + bq;
    if (b >= p.B || u >= H) continue;
    float gsum[4];
#pragma unroll
    for (int g = 0; g < 4; g++) {
      float v[8];
#pragma unroll
      for (int q = 0; q < 8; q++) v[q] = q < nz ? rs[q][b * BM + g * 32 + ul] : 0.f;
      gsum[g] = ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
    }
    const float* ar = p.addrows + (p.rowsel ? (int64_t)(__ldcg(p.rowsel + b) - 1) * p.addld : 0) + u;
    const float i_ = decb::sigmoidf_(gsum[0] + ar[0]);
    const float f_ = decb::sigmoidf_(gsum[1] + ar[H]);
    const float o_ = decb::sigmoidf_(gsum[2] + ar[2 * H]);
    const float g_ = tanhf(gsum[3] + ar[3 * H]);
    const int64_t e = (int64_t)b * H + u;
    const float c = f_ * p.c_prev[e] + i_ * g_;
    const float h = o_ * tanhf(c);
    p.c_new[e] = c;
    float* a = p.acts + (int64_t)b * 4 * H + u;
    a[0] = i_; a[H] = f_; a[2 * H] = o_; a[3 * H] = g_;
    p.h_out0[(int64_t)b * p.ld0 + u] = h;
    if (p.h_out1) p.h_out1[(int64_t)b * p.ld1 + u] = h;
    decb::pack_store(p.pk0, b, u, h);
    decb::pack_store(p.pk1, b, u, h);
  }
}

template <typename T>
__device__ __forceinline__ const T& payload(const PCmd& c) { return *reinterpret_cast<const T*>(c.payload); }

// kSet: bit mask of the command types compiled into this instance.  Every program family (encoder forward / backward,
// decoder forward / backward, decode) runs the instance that holds just its commands: measured, an interpreter that
// carries the decode tail costs the training programs ~3 % through register allocation and code layout alone.
__host__ __device__ constexpr unsigned bit(int t) { return 1u << t; }
constexpr unsigned kSetEncFwd = bit(P_GEMM) | bit(P_ENC_CELL_FWD) | bit(P_GEMM_ENC_FWD);
constexpr unsigned kSetEncBwd = bit(P_GEMM) | bit(P_ENC_CELL_BWD);
constexpr unsigned kSetDecFwd = bit(P_GEMM) | bit(P_CELL_FWD) | bit(P_GEMM_CELL_FWD) | bit(P_ATTN_OUT);
constexpr unsigned kSetDecBwd = bit(P_GEMM) | bit(P_CELL_BWD) | bit(P_ATTN_DU) | bit(P_TO_DENSE);
constexpr unsigned kSetDecode = kSetDecFwd | bit(P_GENERATOR) | bit(P_GREEDY) | bit(P_ATTN_OUT_GEN);
constexpr unsigned kSetAll = 0xffffffffu;
template <int BN, unsigned kSet>
__global__ void __launch_bounds__(256, 1)
persist_kernel(const PCmd* __restrict__ cmds, int ncmds, const CUtensorMap* __restrict__ maps, unsigned* barrier,
               unsigned long long* trace) {
  using C_ = PCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t bars = base + C_::kStages * C_::kStageBytes;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (C_::kStages + s); };
  const uint32_t tmem_full_bar = bars + 8u * (2 * C_::kStages);
  const uint32_t tmem_slot = bars + 8u * (2 * C_::kStages + 1);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
  float* scratch = reinterpret_cast<float*>(smem_raw + (base - raw));   // the (idle) TMA ring doubles as body scratch

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bid = blockIdx.x, nblk = gridDim.x;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < C_::kStages; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "n"(C_::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  unsigned epoch = 0;
  uint32_t it = 0;        // k-blocks this CTA has pushed through the ring so far (producer and issuer count alike)
  uint32_t tiles = 0;     // GEMM tiles this CTA has finished (parity of the TMEM-full barrier)

  for (int c = 0; c < ncmds; c++) {
    const PCmd& cmd = cmds[c];
    const int type = cmd.type;
    unsigned long long t_begin = 0;
    if (trace && threadIdx.x == 0) {
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_begin));
      if (bid == 0) trace[2 * c] = t_begin;
    }
    auto is = [&](int t) { return (kSet & bit(t)) != 0 && type == t; };
    if (is(P_GEMM) || is(P_GEMM_ENC_FWD) || is(P_GEMM_CELL_FWD)) {
      // P_GEMM: tile (mt, z) = (bid % m_tiles, bid / m_tiles), raw split-K partial -> global workspace.
      // Fused commands: the cluster is the M tile and the rank in the cluster is the split; the partial goes to this
      // CTA's shared memory (the idle TMA ring), the cluster reduces through DSMEM and applies the cell right away.
      const bool fused = (kSet & (bit(P_GEMM_ENC_FWD) | bit(P_GEMM_CELL_FWD))) != 0 && (type != P_GEMM);
      const PGemm g = payload<PGemm>(cmd);
      const int ntiles = g.m_tiles * g.splits;
      const int z = fused ? (int)(bid % g.splits) : bid / g.m_tiles;
      const int mt = fused ? (int)(bid / g.splits) : bid % g.m_tiles;
      float* stage = reinterpret_cast<float*>(smem_raw + (base - raw));   // [BN columns][128 rows] fp32
      if (bid < ntiles) {
        const int kb_begin = z * g.kb_per;
        const int kb_end = min(g.num_kb, kb_begin + g.kb_per);
        const int nkb = kb_end - kb_begin;
        const int m0 = mt * BM;
        if (warp == 0) {
          if (lane == 0) {
            const uint32_t tx = (uint32_t)(g.terms == 3 ? 2 : 1) * (A_PLANE_BYTES + C_::kBPlane);
            const CUtensorMap* tAh = maps + g.map_a;
            const CUtensorMap* tBh = maps + g.map_b;
            const uint64_t pol_w = l2_policy_evict_last_frac();   // weights: re-read every timestep
            for (int i = 0; i < nkb; i++) {
              const uint32_t n = it + i;
              const int s = n % C_::kStages;
              const uint32_t ph = (n / C_::kStages) & 1u;
              mbar_wait(empty_bar(s), ph ^ 1u);
              const uint32_t sa = base + s * C_::kStageBytes;
              const uint32_t sb = sa + 2 * A_PLANE_BYTES;
              mbar_expect_tx(full_bar(s), tx);
              const int kc = (kb_begin + i) * BK;
              tma_load_2d_hint(sa, tAh, full_bar(s), kc, m0, pol_w);
              tma_load_2d(sb, tBh, full_bar(s), g.b_k0 + kc, g.b_row0);
              if (g.terms == 3) {
                tma_load_2d_hint(sa + A_PLANE_BYTES, tAh + 1, full_bar(s), kc, m0, pol_w);
                tma_load_2d(sb + C_::kBPlane, tBh + 1, full_bar(s), g.b_k0 + kc, g.b_row0);
              }
            }
          }
        } else if (warp == 1) {
          const uint32_t idesc = make_idesc(BN, 0);
          for (int i = 0; i < nkb; i++) {
            const uint32_t n = it + i;
            const int s = n % C_::kStages;
            const uint32_t ph = (n / C_::kStages) & 1u;
            mbar_wait(full_bar(s), ph);
            tc_fence_after();
            if (lane == 0) {
              const uint32_t sa = base + s * C_::kStageBytes;
              const uint32_t sb = sa + 2 * A_PLANE_BYTES;
              const uint64_t dah = make_desc_kmajor_sw128(sa), dal = make_desc_kmajor_sw128(sa + A_PLANE_BYTES);
              const uint64_t dbh = make_desc_kmajor_sw128(sb), dbl = make_desc_kmajor_sw128(sb + C_::kBPlane);
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; k++) {
                const uint64_t adv = (uint64_t)((k * UMMA_K * 2) >> 4);
                tc_mma(tmem_base, dah + adv, dbh + adv, idesc, (i > 0 || k > 0) ? 1u : 0u);
                if (g.terms == 3) {
                  tc_mma(tmem_base, dah + adv, dbl + adv, idesc, 1u);
                  tc_mma(tmem_base, dal + adv, dbh + adv, idesc, 1u);
                }
              }
              tc_commit(empty_bar(s));
              if (i == nkb - 1) tc_commit(tmem_full_bar);
            }
            __syncwarp();
          }
        } else if (warp < 6) {
          const int q = warp & 3;
          mbar_wait(tmem_full_bar, tiles & 1u);
          tc_fence_after();
          const int row = m0 + q * 32 + lane;                 // weight row = output column of the (batch x M) result
          float* outp = g.ws + (long long)z * g.part_stride;
#pragma unroll 1
          for (int c0 = 0; c0 < BN; c0 += 16) {
            uint32_t r[16];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (fused) {
              // all of this CTA's MMAs have retired (tmem_full), so the ring is free: stage[column][tile row]
#pragma unroll
              for (int j = 0; j < 16; j++) stage[(c0 + j) * BM + q * 32 + lane] = __uint_as_float(r[j]);
            } else if (row < g.M) {
#pragma unroll
              for (int j = 0; j < 16; j++) {
                const int n = c0 + j;
                if (n < g.N) outp[(long long)n * g.ldc + row] = __uint_as_float(r[j]);   // coalesced over rows
              }
            }
          }
          tc_fence_before();
        }
        it += (uint32_t)nkb;
        tiles += 1;
      }
      if (fused) {
        cg::cluster_group cluster = cg::this_cluster();
        cluster.sync();                                       // every rank's partial tile is in its shared memory
        if (bid < ntiles) {
          if constexpr ((kSet & bit(P_GEMM_ENC_FWD)) != 0) {
            if (type == P_GEMM_ENC_FWD)
              fused_enc_cell_fwd<BN>(*reinterpret_cast<const EncCellFwdTc*>(cmd.payload + sizeof(PGemm)), cluster, stage, mt, z,
                                     g.splits);
          }
          if constexpr ((kSet & bit(P_GEMM_CELL_FWD)) != 0) {
            if (type == P_GEMM_CELL_FWD)
              fused_cell_fwd<BN>(*reinterpret_cast<const CellFwdTc*>(cmd.payload + sizeof(PGemm)), cluster, stage, mt, z, g.splits);
          }
        }
        // no second cluster barrier: the grid barrier that ends the command orders the peers' reads of this CTA's
        // partial before anything reuses the ring
      }
    } else if (is(P_CELL_FWD)) {
      if constexpr ((kSet & bit(P_CELL_FWD)) != 0) decb::cell_fwd_tc_body(payload<CellFwdTc>(cmd), bid, nblk, scratch);
    } else if (is(P_CELL_BWD)) {
      if constexpr ((kSet & bit(P_CELL_BWD)) != 0) decb::cell_bwd_tc_body(payload<CellBwdTc>(cmd), bid, nblk, scratch);
    } else if (is(P_ENC_CELL_FWD)) {
      if constexpr ((kSet & bit(P_ENC_CELL_FWD)) != 0) decb::enc_cell_fwd_tc_body(payload<EncCellFwdTc>(cmd), bid, nblk, scratch);
    } else if (is(P_ENC_CELL_BWD)) {
      if constexpr ((kSet & bit(P_ENC_CELL_BWD)) != 0) decb::enc_cell_bwd_tc_body(payload<EncCellBwdTc>(cmd), bid, nblk, scratch);
    } else if (is(P_TO_DENSE)) {
      if constexpr ((kSet & bit(P_TO_DENSE)) != 0) {
        const PToDense p = payload<PToDense>(cmd);
        decb::part_to_dense_body(p.in, p.dst, p.ld, p.B, p.cols, bid, nblk, scratch);
      }
    } else if (is(P_GENERATOR)) {
      if constexpr ((kSet & bit(P_GENERATOR)) != 0) decb::generator_body(payload<GenTc>(cmd), bid, nblk, scratch);
    } else if (is(P_GREEDY)) {
      if constexpr ((kSet & bit(P_GREEDY)) != 0) decb::greedy_select_body(payload<GreedyTc>(cmd), bid, nblk, scratch);
    } else if (is(P_ATTN_OUT)) {
      if constexpr ((kSet & bit(P_ATTN_OUT)) != 0) {
        const AttnOutTc p = payload<AttnOutTc>(cmd);
        for (int b = bid; b < p.B; b += nblk) {
          decb::attn_out_tc_body(p, b, scratch);
          __syncthreads();
        }
      }
    } else if (is(P_ATTN_OUT_GEN)) {
      if constexpr ((kSet & bit(P_ATTN_OUT_GEN)) != 0) {
      const AttnOutTc p = payload<AttnOutTc>(cmd);
      const GenTc& gp = *reinterpret_cast<const GenTc*>(cmd.payload + sizeof(AttnOutTc));
      const GreedyTc& gs = *reinterpret_cast<const GreedyTc*>(cmd.payload + sizeof(AttnOutTc) + sizeof(GenTc));
      float* as = scratch + ((p.S + 3) & ~3) + 16 + 8 * p.H;     // = the `qs` region of the attention body
      float* zs = as + p.H;
      for (int b = bid; b < p.B; b += nblk) {
        decb::attn_out_tc_body(p, b, scratch, true);
        decb::gen_select_tail(gp, gs, b, as, zs);
        __syncthreads();
      }
      }
    } else if (is(P_ATTN_DU)) {
      if constexpr ((kSet & bit(P_ATTN_DU)) != 0) {
        const AttnDuTc p = payload<AttnDuTc>(cmd);
        for (int b = bid; b < p.B; b += nblk) {
          decb::attn_du_tc_body(p, b, scratch);
          __syncthreads();
        }
      }
    }
    if (trace) {
      __syncthreads();           // (tracing only) the whole CTA is done with the command
      if (threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        if (bid == 0) trace[2 * c + 1] = t;      // work of CTA 0 done; the barrier wait follows
        trace[2 * (ncmds + 1) + (size_t)c * nblk + bid] = t - t_begin;
      }
    }
    grid_sync(barrier, (unsigned)nblk, epoch);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C_::kTmemCols) : "memory");
  }
}

template <int BN, unsigned kSet>
void launch_bn(Ctx& ctx, PersistProgram& prog) {
  static bool attr_set = false;
  if (!attr_set) {
    AOCR_CUDA(cudaFuncSetAttribute(persist_kernel<BN, kSet>, cudaFuncAttributeMaxDynamicSharedMemorySize, PCfg<BN>::kSmemBytes));
    attr_set = true;
  }
  const PCmd* cmds = prog.d_cmds;
  int ncmds = (int)prog.cmds.size();
  const CUtensorMap* maps = prog.d_maps;
  unsigned* bar = prog.d_barrier;
  unsigned long long* trace = prog.d_trace;
  void* args[] = {(void*)&cmds, (void*)&ncmds, (void*)&maps, (void*)&bar, (void*)&trace};
  AOCR_CUDA(cudaMemsetAsync(prog.d_barrier, 0, sizeof(unsigned), ctx.st));
  // cooperative (all CTAs co-resident: the grid barrier spins) + thread-block clusters of kCluster CTAs (the fused
  // GEMM -> cell commands reduce their split-K partials through distributed shared memory inside a cluster)
  cudaLaunchConfig_t cfgl = {};
  cfgl.gridDim = dim3(prog.grid); cfgl.blockDim = dim3(256);
  cfgl.dynamicSmemBytes = (size_t)PCfg<BN>::kSmemBytes; cfgl.stream = ctx.st;
  cudaLaunchAttribute attrs[2];
  attrs[0].id = cudaLaunchAttributeCooperative; attrs[0].val.cooperative = 1;
  attrs[1].id = cudaLaunchAttributeClusterDimension;
  attrs[1].val.clusterDim.x = prog.cluster; attrs[1].val.clusterDim.y = 1; attrs[1].val.clusterDim.z = 1;
  cfgl.attrs = attrs; cfgl.numAttrs = prog.cluster > 1 ? 2 : 1;
  AOCR_CUDA(cudaLaunchKernelExC(&cfgl, (const void*)persist_kernel<BN, kSet>, args));
  ctx.launches++;
}

template <int BN>
int max_ctas_bn() {
  int dev = 0, sms = 0, per_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaFuncSetAttribute(persist_kernel<BN, kSetAll>, cudaFuncAttributeMaxDynamicSharedMemorySize, PCfg<BN>::kSmemBytes);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, persist_kernel<BN, kSetAll>, 256, (size_t)PCfg<BN>::kSmemBytes);
  return sms * per_sm;
}

}  // namespace

PGemmPlan persist_plan_gemm(int M, int N, int K, int max_ctas, long long ws_floats) {
  PGemmPlan p;
  p.m_tiles = (M + BM - 1) / BM;
  p.num_kb = (int)(pad64(K) / BK);
  p.part_stride = (((long long)(N - 1) * M + M) + 63) & ~63LL;
  int splits = max_ctas / p.m_tiles;
  if (splits > p.num_kb / 2) splits = p.num_kb / 2;
  if (splits > 8) splits = 8;       // = decb::kMaxSplits (the consumers sum at most 8 partials)
  if (splits < 1) splits = 1;
  while (splits > 1 && (long long)splits * p.part_stride > ws_floats) splits--;
  p.kb_per = (p.num_kb + splits - 1) / splits;
  p.splits = (p.num_kb + p.kb_per - 1) / p.kb_per;
  return p;
}

PGemmPlan persist_plan_gemm_fused(int M, int N, int K, int cluster) {
  PGemmPlan p;
  p.m_tiles = (M + BM - 1) / BM;
  p.num_kb = (int)(pad64(K) / BK);
  p.part_stride = 0;
  p.kb_per = (p.num_kb + cluster - 1) / cluster;
  p.splits = cluster;
  AOCR_CHECK((p.num_kb + p.kb_per - 1) / p.kb_per == cluster, "fused GEMM: K too short for one k-block range per cluster rank");
  return p;
}

// co-resident CTAs when the executor is launched with clusters of `cluster` CTAs: clusters must fit inside a GPC, so this
// can be less than the plain occupancy (148 SMs in unequal GPCs)
template <int BN>
int max_cluster_ctas_bn(int cluster) {
  cudaFuncSetAttribute(persist_kernel<BN, kSetAll>, cudaFuncAttributeMaxDynamicSharedMemorySize, PCfg<BN>::kSmemBytes);
  cudaFuncSetAttribute(persist_kernel<BN, kSetAll>, cudaFuncAttributeNonPortableClusterSizeAllowed, 0);
  int sms = 0, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaLaunchConfig_t cfgl = {};
  cfgl.gridDim = dim3((sms / cluster) * cluster); cfgl.blockDim = dim3(256);
  cfgl.dynamicSmemBytes = (size_t)PCfg<BN>::kSmemBytes;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = cluster; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
  cfgl.attrs = &attr; cfgl.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, (const void*)persist_kernel<BN, kSetAll>, &cfgl) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n * cluster;
}
int persist_max_cluster_ctas(int bn, int cluster) {
  if (cluster <= 1) return persist_max_ctas(bn);
  switch (bn) {
    case 128: return max_cluster_ctas_bn<128>(cluster);
    case 64: return max_cluster_ctas_bn<64>(cluster);
    case 32: return max_cluster_ctas_bn<32>(cluster);
    default: return max_cluster_ctas_bn<16>(cluster);
  }
}

int persist_max_ctas(int bn) {
  switch (bn) {
    case 128: return max_ctas_bn<128>();
    case 64: return max_ctas_bn<64>();
    case 32: return max_ctas_bn<32>();
    default: return max_ctas_bn<16>();
  }
}

void persist_upload(Ctx& ctx, PersistProgram& prog) {
  if (prog.uploaded) return;
  AOCR_CHECK(!prog.cmds.empty(), "empty persistent program");
  AOCR_CUDA(cudaMalloc(&prog.d_cmds, prog.cmds.size() * sizeof(PCmd)));
  AOCR_CUDA(cudaMalloc(&prog.d_maps, (prog.maps.size() + 1) * sizeof(CUtensorMap)));
  AOCR_CUDA(cudaMalloc(&prog.d_barrier, 256));
  if (getenv("AOCR_PERSIST_TRACE")) {
    const size_t nt = (prog.cmds.size() + 1) * 2 + prog.cmds.size() * (size_t)prog.grid;
    AOCR_CUDA(cudaMalloc(&prog.d_trace, nt * sizeof(unsigned long long)));
    AOCR_CUDA(cudaMemset(prog.d_trace, 0, nt * sizeof(unsigned long long)));
  }
  AOCR_CUDA(cudaMemcpy(prog.d_cmds, prog.cmds.data(), prog.cmds.size() * sizeof(PCmd), cudaMemcpyHostToDevice));
  if (!prog.maps.empty())
    AOCR_CUDA(cudaMemcpy(prog.d_maps, prog.maps.data(), prog.maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
  prog.uploaded = true;
}

void persist_launch(Ctx& ctx, PersistProgram& prog) {
  AOCR_CHECK(prog.uploaded, "persistent program not uploaded");
  unsigned used = 0;
  for (const PCmd& c : prog.cmds) used |= bit(c.type);
  // the smallest predefined command set that covers the program
  auto go = [&](auto set) {
    constexpr unsigned S = decltype(set)::value;
    switch (prog.bn) {
      case 128: launch_bn<128, S>(ctx, prog); break;
      case 64: launch_bn<64, S>(ctx, prog); break;
      case 32: launch_bn<32, S>(ctx, prog); break;
      default: launch_bn<16, S>(ctx, prog); break;
    }
  };
  if ((used & ~kSetEncFwd) == 0) go(std::integral_constant<unsigned, kSetEncFwd>());
  else if ((used & ~kSetEncBwd) == 0) go(std::integral_constant<unsigned, kSetEncBwd>());
  else if ((used & ~kSetDecFwd) == 0) go(std::integral_constant<unsigned, kSetDecFwd>());
  else if ((used & ~kSetDecBwd) == 0) go(std::integral_constant<unsigned, kSetDecBwd>());
  else if ((used & ~kSetDecode) == 0) go(std::integral_constant<unsigned, kSetDecode>());
  else go(std::integral_constant<unsigned, kSetAll>());
}

void persist_free(PersistProgram& prog) {
  if (prog.d_cmds) cudaFree(prog.d_cmds);
  if (prog.d_maps) cudaFree(prog.d_maps);
  if (prog.d_barrier) cudaFree(prog.d_barrier);
  if (prog.d_trace) {   // AOCR_PERSIST_TRACE: per-command time of CTA 0 (work, then barrier wait), averaged by type
    const size_t nc = prog.cmds.size(), ng = (size_t)prog.grid;
    std::vector<unsigned long long> t((nc + 1) * 2 + nc * ng);
    cudaMemcpy(t.data(), prog.d_trace, t.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    double work[16] = {0}, wait[16] = {0}, wmax[16] = {0}, wmed[16] = {0}; int cnt[16] = {0};
    std::vector<int> slow(16 * ng, 0);
    for (size_t c = 0; c + 1 < nc; c++) {
      int ty = prog.cmds[c].type & 15;
      work[ty] += (double)(t[2 * c + 1] - t[2 * c]); wait[ty] += (double)(t[2 * c + 2] - t[2 * c + 1]); cnt[ty]++;
      std::vector<unsigned long long> w(t.begin() + (nc + 1) * 2 + c * ng, t.begin() + (nc + 1) * 2 + (c + 1) * ng);
      size_t am = 0;
      for (size_t i = 0; i < ng; i++) if (w[i] > w[am]) am = i;
      slow[ty * ng + am]++;
      wmax[ty] += (double)w[am];
      std::sort(w.begin(), w.end());
      wmed[ty] += (double)w[ng / 2];
    }
    fprintf(stderr, "[persist trace] %zu cmds grid %d:", nc, prog.grid);
    for (int ty = 0; ty < 16; ty++) if (cnt[ty]) {
      size_t top = 0;
      for (size_t i = 0; i < ng; i++) if (slow[ty * ng + i] > slow[ty * ng + top]) top = i;
      fprintf(stderr, " type%d n=%d cta0 work=%.2fus wait=%.2fus | all CTAs: median %.2fus max %.2fus (slowest most often: CTA %zu, %d times);",
              ty, cnt[ty], work[ty] / cnt[ty] / 1e3, wait[ty] / cnt[ty] / 1e3, wmed[ty] / cnt[ty] / 1e3, wmax[ty] / cnt[ty] / 1e3, top, slow[ty * ng + top]);
    }
    fprintf(stderr, "\n");
    cudaFree(prog.d_trace); prog.d_trace = nullptr;
  }
  prog.d_cmds = nullptr; prog.d_maps = nullptr; prog.d_barrier = nullptr; prog.uploaded = false;
}

}  // namespace aocr
