// persist.cu — persistent recurrence executor (see persist.h).  256 threads per CTA, one CTA per SM (197 KB of shared
// memory for the TMA ring): warp 0 = TMA producer, warp 1 = tcgen05 issuer and TMEM owner, warps 2-5 = epilogue (TMEM ->
// split-K partials in global memory, or -> this CTA's shared memory for the fused GEMM -> cell commands), all 8 warps =
// the cell / attention bodies and the cluster-side cell of the fused commands.
// mbarrier phases and the TMEM allocation persist across commands; a grid barrier (one global counter, acquire /
// release, preceded by fence.proxy.async so the generic-proxy stores of a command are visible to the TMA loads of
// the next one on every SM) replaces the kernel boundary between consecutive commands.
#include <algorithm>
#include <mutex>
#include <set>
#include <type_traits>

#include <cooperative_groups.h>

#include "persist.h"

// phase time stamps inside the bodies (diagnostic; CTA 0, thread 0; printed by persist_free under AOCR_PERSIST_TRACE)
namespace aocr { __device__ unsigned long long g_bt[16]; __device__ int g_bt_on = 0; __device__ int g_variant = 0; }
#define AOCR_VARIANT (aocr::g_variant)
#define AOCR_BT(i)                                                                          \
  do {                                                                                      \
    if (aocr::g_bt_on && blockIdx.x == 0 && threadIdx.x == 0) {                             \
      unsigned long long _t;                                                                \
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(_t));                                 \
      aocr::g_bt[i] = _t;                                                                   \
    }                                                                                       \
  } while (0)

#define AOCR_BT_T(i, t)                                                                     \
  do {                                                                                      \
    if (aocr::g_bt_on && blockIdx.x == 0 && threadIdx.x == (t)) {                           \
      unsigned long long _t;                                                                \
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(_t));                                 \
      aocr::g_bt[i] = _t;                                                                   \
    }                                                                                       \
  } while (0)

#include "dec_bodies.cuh"
#include "tc_ptx.cuh"

namespace aocr {

namespace {
using namespace tcp;
namespace cg = cooperative_groups;

// Shared-memory layout (per CTA, one CTA per SM):
//   [A ring: kStages x (hi 16 KB | lo 16 KB)]  weight tiles only.  Never overlaid by anything, so (i) the A tiles of the
//        NEXT command are loaded before the grid barrier that ends the current one (weights do not depend on it) and
//        (ii) a tile that is still in its slot from an earlier timestep is not loaded again (slot tags) - the encoder's
//        recurrent weights (64 KB per CTA) become resident after two steps.
//   [B ring: kStages x (hi | lo) of BN activation rows]  = the scratch region: dead once a command's MMAs have retired,
//        it doubles as the fused commands' split-K staging tile ([BN columns][128 rows] fp32, read by the cluster
//        peers through DSMEM until the grid barrier) and as the bodies' scratch; B loads are issued after the barrier.
//   [mbarriers | slot tags | TMEM slot | 2 command buffers (the next command is fetched while the current one runs)]
template <int BN> struct PCfg {
  static constexpr int kStages = BN == 256 ? 2 : (BN == 128 ? 3 : 4);
  static constexpr int kBPlane = BN * BK * 2;
  static constexpr int kABytes = 2 * A_PLANE_BYTES;
  static constexpr int kBBytes = 2 * kBPlane;
  static constexpr int kARegion = kStages * kABytes;
  static constexpr int kBRegion = kStages * kBBytes;
  static constexpr int kScratch = kBRegion > 49152 ? kBRegion : 49152;
  static constexpr int kTmemCols = BN < 32 ? 32 : BN;
  static constexpr int kCtlBytes = 256;                       // barriers (2*kStages + 1) * 8, TMEM slot, slot tags
  static constexpr int kSmemBytes = kARegion + kScratch + kCtlBytes + 2 * (int)sizeof(PCmd) + 1024;
  static_assert(BN * BM * 4 <= kScratch, "fused staging tile must fit in the scratch region");
};

__device__ __forceinline__ void mbar_expect_tx_only(uint32_t bar, uint32_t bytes) {   // no arrival: a prefetch posts its bytes early
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// identity of an A (weight) tile: tensor-map pair, row origin, k origin
__device__ __forceinline__ unsigned long long a_tag(int map_a, int m0, int kc) {
  return ((unsigned long long)(unsigned)map_a << 44) | ((unsigned long long)(unsigned)m0 << 22) | (unsigned long long)(unsigned)kc;
}

// tile assignment of a GEMM-type command: (mt, z) of this CTA, or has == false
struct TileOf { int mt, z, kb_begin, nkb; bool has; };
__device__ __forceinline__ TileOf tile_of(const PGemm& g, bool fused, int bid) {
  TileOf t;
  t.has = bid < g.m_tiles * g.splits;
  t.z = fused ? (int)(bid % g.splits) : bid / g.m_tiles;
  t.mt = fused ? (int)(bid / g.splits) : bid % g.m_tiles;
  t.kb_begin = t.z * g.kb_per;
  t.nkb = min(g.num_kb, t.kb_begin + g.kb_per) - t.kb_begin;
  return t;
}

// ---- fused GEMM -> cell: the cluster owns 32 hidden units x 4 gates (gate-interleaved weight rows: tile row =
// gate*32 + unit) for the whole batch; rank z holds the split-K partial z in shared memory as stage[column][row].
// Rank z finishes batch columns [z*BN/nz, (z+1)*BN/nz): sums the nz partials of the 4 gates through DSMEM, applies the
// LSTM cell and writes what the stand-alone cell body writes.
// Everything the cell needs besides the gate sums (input projection / embedding row, previous cell state) is loaded
// into registers at the START of the command (CellPre), so those L2 round trips overlap the GEMM instead of following
// the cluster barrier.
constexpr int kCellPre = 4;      // batch columns per thread whose inputs are preloaded (BN / cluster / 8 warps <= 4 up to BN = 128)
struct CellPre { float a[kCellPre][4]; float c[kCellPre]; };

template <int BN>
__device__ __forceinline__ void enc_cell_preload(const EncCellFwdTc& p, int mt, int z, int nz, CellPre& r) {
  const int He = p.He, B = p.B, S = p.S, d = p.d_only;
  const int unit = mt * 32 + (threadIdx.x & 31), cols = BN / nz;
  const int t = d == 0 ? p.step : S - 1 - p.step;
  const int prev_slot = d == 0 ? t : t + 1;
#pragma unroll
  for (int j = 0; j < kCellPre; j++) {
    const int bq = (threadIdx.x >> 5) + 8 * j, b = z * cols + bq;
    if (bq >= cols || b >= B || unit >= He) continue;
    const float* xg = p.xg + ((int64_t)t * B + b) * (8 * He) + (int64_t)d * 4 * He + unit;
#pragma unroll
    for (int g = 0; g < 4; g++) r.a[j][g] = __ldcg(xg + g * He);
    r.c[j] = __ldcg(p.Cst + ((int64_t)(d * (S + 1) + prev_slot) * B + b) * He + unit);
  }
}
template <int BN>
__device__ __forceinline__ void fused_enc_cell_fwd(const EncCellFwdTc& p, cg::cluster_group& cluster, float* stage, int mt,
                                                   int z, int nz, const CellPre& r) {
  const int He = p.He, B = p.B, S = p.S;
  const int d = p.d_only;
  const int ul = threadIdx.x & 31, unit = mt * 32 + ul;
  const int cols = BN / nz;
  const int t = d == 0 ? p.step : S - 1 - p.step;
  const int prev_slot = d == 0 ? t : t + 1, out_slot = d == 0 ? t + 1 : t;
  const float* rs[8];
#pragma unroll
  for (int q = 0; q < 8; q++) rs[q] = q < nz ? cluster.map_shared_rank(stage, q) : stage;
#pragma unroll 1
  for (int j = 0; j * 8 < cols; j++) {
    const int bq = (threadIdx.x >> 5) + 8 * j, b = z * cols + bq;
    if (bq >= cols || b >= B || unit >= He) continue;
    float gsum[4];
#pragma unroll
    for (int g = 0; g < 4; g++) {
      float v[8];
#pragma unroll
      for (int q = 0; q < 8; q++) v[q] = q < nz ? rs[q][b * BM + g * 32 + ul] : 0.f;
      gsum[g] = ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
    }
    float a4[4], cp;
    if (j < kCellPre) {
#pragma unroll
      for (int jj = 0; jj < kCellPre; jj++)
        if (jj == j) { a4[0] = r.a[jj][0]; a4[1] = r.a[jj][1]; a4[2] = r.a[jj][2]; a4[3] = r.a[jj][3]; cp = r.c[jj]; }
    } else {
      const float* xg = p.xg + ((int64_t)t * B + b) * (8 * He) + (int64_t)d * 4 * He + unit;
      a4[0] = __ldcg(xg); a4[1] = __ldcg(xg + He); a4[2] = __ldcg(xg + 2 * He); a4[3] = __ldcg(xg + 3 * He);
      cp = __ldcg(p.Cst + ((int64_t)(d * (S + 1) + prev_slot) * B + b) * He + unit);
    }
    const float i_ = decb::sigmoidf_(gsum[0] + a4[0]);
    const float f_ = decb::sigmoidf_(gsum[1] + a4[1]);
    const float o_ = decb::sigmoidf_(gsum[2] + a4[2]);
    const float g_ = tanhf(gsum[3] + a4[3]);
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
__device__ __forceinline__ void cell_preload(const CellFwdTc& p, int mt, int z, int nz, CellPre& r) {
  const int H = p.H;
  const int u = mt * 32 + (threadIdx.x & 31), cols = BN / nz;
#pragma unroll
  for (int j = 0; j < kCellPre; j++) {
    const int bq = (threadIdx.x >> 5) + 8 * j, b = z * cols + bq;
    if (bq >= cols || b >= p.B || u >= H) continue;
    const float* ar = p.addrows + (p.rowsel ? (int64_t)(__ldcg(p.rowsel + b) - 1) * p.addld : 0) + u;
#pragma unroll
    for (int g = 0; g < 4; g++) r.a[j][g] = __ldcg(ar + g * H);
    r.c[j] = __ldcg(p.c_prev + (int64_t)b * H + u);
  }
}
template <int BN>
__device__ __forceinline__ void fused_cell_fwd(const CellFwdTc& p, cg::cluster_group& cluster, float* stage, int mt, int z,
                                               int nz, const CellPre& r) {
  const int H = p.H;
  const int ul = threadIdx.x & 31, u = mt * 32 + ul;
  const int cols = BN / nz;
  const float* rs[8];
#pragma unroll
  for (int q = 0; q < 8; q++) rs[q] = q < nz ? cluster.map_shared_rank(stage, q) : stage;
#pragma unroll 1
  for (int j = 0; j * 8 < cols; j++) {
    const int bq = (threadIdx.x >> 5) + 8 * j, b = z * cols + bq;
    if (bq >= cols || b >= p.B || u >= H) continue;
    float gsum[4];
#pragma unroll
    for (int g = 0; g < 4; g++) {
      float v[8];
#pragma unroll
      for (int q = 0; q < 8; q++) v[q] = q < nz ? rs[q][b * BM + g * 32 + ul] : 0.f;
      gsum[g] = ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
    }
    const int64_t e = (int64_t)b * H + u;
    float a4[4], cp;
    if (j < kCellPre) {
#pragma unroll
      for (int jj = 0; jj < kCellPre; jj++)
        if (jj == j) { a4[0] = r.a[jj][0]; a4[1] = r.a[jj][1]; a4[2] = r.a[jj][2]; a4[3] = r.a[jj][3]; cp = r.c[jj]; }
    } else {
      const float* ar = p.addrows + (p.rowsel ? (int64_t)(__ldcg(p.rowsel + b) - 1) * p.addld : 0) + u;
      a4[0] = __ldcg(ar); a4[1] = __ldcg(ar + H); a4[2] = __ldcg(ar + 2 * H); a4[3] = __ldcg(ar + 3 * H);
      cp = __ldcg(p.c_prev + e);
    }
    const float i_ = decb::sigmoidf_(gsum[0] + a4[0]);
    const float f_ = decb::sigmoidf_(gsum[1] + a4[1]);
    const float o_ = decb::sigmoidf_(gsum[2] + a4[2]);
    const float g_ = tanhf(gsum[3] + a4[3]);
    const float c = f_ * cp + i_ * g_;
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
constexpr unsigned kGemmTypes = bit(P_GEMM) | bit(P_GEMM_ENC_FWD) | bit(P_GEMM_CELL_FWD);

template <int BN, unsigned kSet>
__global__ void __launch_bounds__(256, 1)
persist_kernel(const PCmd* __restrict__ cmds, int ncmds, const CUtensorMap* __restrict__ maps, unsigned* barrier,
               unsigned long long* trace, int flags) {
  using C_ = PCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t baseB = base + C_::kARegion;
  const uint32_t bars = baseB + C_::kScratch;
  auto slot_a = [&](int s) { return base + (uint32_t)s * C_::kABytes; };
  auto slot_b = [&](int s) { return baseB + (uint32_t)s * C_::kBBytes; };
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (C_::kStages + s); };
  const uint32_t tmem_full_bar = bars + 8u * (2 * C_::kStages);
  const uint32_t tmem_slot = bars + 8u * (2 * C_::kStages + 1);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
  unsigned long long* tags = reinterpret_cast<unsigned long long*>(smem_raw + (bars + 128u - raw));   // [kStages], producer thread only
  const uint32_t cbuf0 = bars + C_::kCtlBytes;
  auto cbuf = [&](int i) { return reinterpret_cast<const PCmd*>(smem_raw + (cbuf0 + (uint32_t)(i & 1) * (uint32_t)sizeof(PCmd) - raw)); };
  float* scratch = reinterpret_cast<float*>(smem_raw + (baseB - raw));   // the (idle) B ring doubles as body scratch / fused stage

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bid = blockIdx.x, nblk = gridDim.x;
  constexpr int kCmdChunks = (int)sizeof(PCmd) / 16;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < C_::kStages; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); tags[s] = ~0ull; }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "n"(C_::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp == 7 && lane < kCmdChunks && ncmds > 0) {
    cp_async16(cbuf0 + lane * 16, reinterpret_cast<const uint8_t*>(cmds) + lane * 16);
    cp_async_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  unsigned epoch = 0;
  uint32_t it = 0;        // k-blocks this CTA has pushed through the ring so far (producer and issuer count alike)
  uint32_t tiles = 0;     // GEMM tiles this CTA has finished (parity of the TMEM-full barrier)
  const uint32_t txA = (uint32_t)C_::kABytes, txB = (uint32_t)C_::kBBytes;

  // A (weight) tile of k-block `kb` of GEMM g into slot s, unless the slot still holds it.  `arrive_with_b`: the call
  // that follows the grid barrier also posts the arrival and the B bytes; the prefetch posts the A bytes only.
  auto load_a = [&](const PGemm& g, int m0, int kb, int s, bool two, uint64_t pol_w) {
    const unsigned long long tag = a_tag(g.map_a, m0, kb);
    if (tags[s] == tag && !(flags & 2)) return false;
    const CUtensorMap* tAh = maps + g.map_a;
    tma_load_2d_hint(slot_a(s), tAh, full_bar(s), kb * BK, m0, pol_w);
    if (two) tma_load_2d_hint(slot_a(s) + A_PLANE_BYTES, tAh + 1, full_bar(s), kb * BK, m0, pol_w);
    tags[s] = tag;
    return true;
  };

  for (int c = 0; c < ncmds; c++) {
    // GEMM-type commands are decoded from the shared-memory copy (no L2 round trip after the barrier); the bodies read
    // their payload from global memory: a payload in shared memory aliases the bodies' scratch stores as far as the
    // compiler can tell, which serialises their loads (measured: attention backward 11.6 -> 17.5 us)
    const PCmd& scmd = *cbuf(c);
    const int type = scmd.type;
    const PCmd& gcmd = cmds[c];
    // fetch the next command into the other buffer while this one runs (its last reader finished before the barrier)
    if (warp == 7 && lane < kCmdChunks && c + 1 < ncmds)
      cp_async16(cbuf0 + (uint32_t)((c + 1) & 1) * (uint32_t)sizeof(PCmd) + lane * 16,
                 reinterpret_cast<const uint8_t*>(cmds + (c + 1)) + lane * 16);
    unsigned long long t_begin = 0;
    if (trace && threadIdx.x == 0) {
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_begin));
      if (bid == 0) trace[2 * c] = t_begin;
    }
    auto is = [&](int t) { return (kSet & bit(t)) != 0 && type == t; };
    if (is(P_GEMM) || is(P_GEMM_ENC_FWD) || is(P_GEMM_CELL_FWD)) {
      // P_GEMM: tile (mt, z) = (bid % m_tiles, bid / m_tiles), raw split-K partial -> global workspace.
      // Fused commands: the cluster is the M tile and the rank in the cluster is the split; the partial goes to this
      // CTA's shared memory (the idle B ring), the cluster reduces through DSMEM and applies the cell right away.
      const bool fused = (kSet & (bit(P_GEMM_ENC_FWD) | bit(P_GEMM_CELL_FWD))) != 0 && (type != P_GEMM);
      const PGemm g = payload<PGemm>(scmd);
      const TileOf tl = tile_of(g, fused, bid);
      const int z = tl.z, mt = tl.mt;
      float* stage = scratch;                                  // [BN columns][128 rows] fp32
      CellPre pre;
      if (fused && tl.has) {
        if constexpr ((kSet & bit(P_GEMM_ENC_FWD)) != 0) {
          if (type == P_GEMM_ENC_FWD)
            enc_cell_preload<BN>(*reinterpret_cast<const EncCellFwdTc*>(scmd.payload + sizeof(PGemm)), mt, z, g.splits, pre);
        }
        if constexpr ((kSet & bit(P_GEMM_CELL_FWD)) != 0) {
          if (type == P_GEMM_CELL_FWD)
            cell_preload<BN>(*reinterpret_cast<const CellFwdTc*>(scmd.payload + sizeof(PGemm)), mt, z, g.splits, pre);
        }
      }
      if (fused) AOCR_BT(11);
      if (tl.has) {
        const int kb_begin = tl.kb_begin, nkb = tl.nkb;
        const int m0 = mt * BM;
        const bool two = g.terms == 3;
        if (warp == 0) {
          if (lane == 0) {
            const CUtensorMap* tBh = maps + g.map_b;
            const uint64_t pol_w = l2_policy_evict_last_frac();   // weights: re-read every timestep
            for (int i = 0; i < nkb; i++) {
              const uint32_t n = it + i;
              const int s = n % C_::kStages;
              const uint32_t ph = (n / C_::kStages) & 1u;
              mbar_wait(empty_bar(s), ph ^ 1u);
              const bool need_a = (flags & 2) || tags[s] != a_tag(g.map_a, m0, kb_begin + i);
              mbar_expect_tx(full_bar(s), (need_a ? (two ? txA : txA / 2) : 0u) + (two ? txB : txB / 2));
              if (need_a) load_a(g, m0, kb_begin + i, s, two, pol_w);
              const int kc = (kb_begin + i) * BK;
              tma_load_2d(slot_b(s), tBh, full_bar(s), g.b_k0 + kc, g.b_row0);
              if (two) tma_load_2d(slot_b(s) + C_::kBPlane, tBh + 1, full_bar(s), g.b_k0 + kc, g.b_row0);
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
              const uint32_t sa = slot_a(s), sb = slot_b(s);
              const uint64_t dah = make_desc_kmajor_sw128(sa), dal = make_desc_kmajor_sw128(sa + A_PLANE_BYTES);
              const uint64_t dbh = make_desc_kmajor_sw128(sb), dbl = make_desc_kmajor_sw128(sb + C_::kBPlane);
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; k++) {
                const uint64_t adv = (uint64_t)((k * UMMA_K * 2) >> 4);
                tc_mma(tmem_base, dah + adv, dbh + adv, idesc, (i > 0 || k > 0) ? 1u : 0u);
                if (two) {
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
          if (fused) AOCR_BT_T(12, 64);
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
              // all of this CTA's MMAs have retired (tmem_full), so the B ring is free: stage[column][tile row]
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
          if (fused) AOCR_BT_T(13, 64);
        }
        it += (uint32_t)nkb;
        tiles += 1;
      }
      if (fused) {
        cg::cluster_group cluster = cg::this_cluster();
        cluster.sync();                                       // every rank's partial tile is in its shared memory
        AOCR_BT(14);
        if (tl.has) {
          if constexpr ((kSet & bit(P_GEMM_ENC_FWD)) != 0) {
            if (type == P_GEMM_ENC_FWD)
              fused_enc_cell_fwd<BN>(*reinterpret_cast<const EncCellFwdTc*>(scmd.payload + sizeof(PGemm)), cluster, stage, mt, z,
                                     g.splits, pre);
          }
          if constexpr ((kSet & bit(P_GEMM_CELL_FWD)) != 0) {
            if (type == P_GEMM_CELL_FWD)
              fused_cell_fwd<BN>(*reinterpret_cast<const CellFwdTc*>(scmd.payload + sizeof(PGemm)), cluster, stage, mt, z, g.splits,
                                 pre);
          }
        }
        AOCR_BT(15);
        // no second cluster barrier: the grid barrier that ends the command orders the peers' reads of this CTA's
        // partial before anything reuses the B ring
      }
    } else if (is(P_CELL_FWD)) {
      if constexpr ((kSet & bit(P_CELL_FWD)) != 0) decb::cell_fwd_tc_body(payload<CellFwdTc>(gcmd), bid, nblk, scratch);
    } else if (is(P_CELL_BWD)) {
      if constexpr ((kSet & bit(P_CELL_BWD)) != 0) decb::cell_bwd_tc_body(payload<CellBwdTc>(gcmd), bid, nblk, scratch);
    } else if (is(P_ENC_CELL_FWD)) {
      if constexpr ((kSet & bit(P_ENC_CELL_FWD)) != 0) decb::enc_cell_fwd_tc_body(payload<EncCellFwdTc>(gcmd), bid, nblk, scratch);
    } else if (is(P_ENC_CELL_BWD)) {
      if constexpr ((kSet & bit(P_ENC_CELL_BWD)) != 0) decb::enc_cell_bwd_tc_body(payload<EncCellBwdTc>(gcmd), bid, nblk, scratch);
    } else if (is(P_TO_DENSE)) {
      if constexpr ((kSet & bit(P_TO_DENSE)) != 0) {
        const PToDense p = payload<PToDense>(gcmd);
        decb::part_to_dense_body(p.in, p.dst, p.ld, p.B, p.cols, bid, nblk, scratch);
      }
    } else if (is(P_GENERATOR)) {
      if constexpr ((kSet & bit(P_GENERATOR)) != 0) decb::generator_body(payload<GenTc>(gcmd), bid, nblk, scratch);
    } else if (is(P_GREEDY)) {
      if constexpr ((kSet & bit(P_GREEDY)) != 0) decb::greedy_select_body(payload<GreedyTc>(gcmd), bid, nblk, scratch);
    } else if (is(P_ATTN_OUT)) {
      if constexpr ((kSet & bit(P_ATTN_OUT)) != 0) {
        const AttnOutTc p = payload<AttnOutTc>(gcmd);
        for (int b = bid; b < p.B; b += nblk) {
          decb::attn_out_tc_body(p, b, scratch);
          __syncthreads();
        }
      }
    } else if (is(P_ATTN_OUT_GEN)) {
      if constexpr ((kSet & bit(P_ATTN_OUT_GEN)) != 0) {
      const AttnOutTc p = payload<AttnOutTc>(gcmd);
      const GenTc& gp = *reinterpret_cast<const GenTc*>(gcmd.payload + sizeof(AttnOutTc));
      const GreedyTc& gs = *reinterpret_cast<const GreedyTc*>(gcmd.payload + sizeof(AttnOutTc) + sizeof(GenTc));
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
        const AttnDuTc p = payload<AttnDuTc>(gcmd);
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
    // ---- grid barrier (one global counter, release arrive / acquire poll) with the weight prefetch of the next command
    // folded in: between this CTA's arrival and the last CTA's, thread 0 issues the A-tile loads of command c+1
    if (warp == 7) cp_async_wait_all();                        // the next command is in its buffer
    asm volatile("fence.proxy.async;" ::: "memory");          // this thread's stores -> visible to the async proxy (TMA)
    __syncthreads();
    epoch++;
    if (threadIdx.x == 0) {
      // arrive: a release reduction (no return value, so no round trip before the polling starts); the release is
      // cumulative over the CTA's writes ordered before it by the bar.sync above
      asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(barrier), "r"(1u) : "memory");
      if (c + 1 < ncmds && !(flags & 1)) {
        const PCmd& nx = *cbuf(c + 1);
        if ((kGemmTypes >> nx.type) & 1u) {
          const PGemm& g = payload<PGemm>(nx);
          const TileOf tl = tile_of(g, nx.type != P_GEMM, bid);
          if (tl.has) {
            const bool two = g.terms == 3;
            const uint64_t pol_w = l2_policy_evict_last_frac();
            const int npre = tl.nkb < C_::kStages ? tl.nkb : C_::kStages;
            for (int i = 0; i < npre; i++) {
              const uint32_t n = it + i;
              const int s = n % C_::kStages;
              const uint32_t ph = (n / C_::kStages) & 1u;
              if (tags[s] == a_tag(g.map_a, tl.mt * BM, tl.kb_begin + i)) continue;     // resident
              mbar_wait(empty_bar(s), ph ^ 1u);             // the slot's last readers have retired (this CTA is idle)
              mbar_expect_tx_only(full_bar(s), two ? txA : txA / 2);
              load_a(g, tl.mt * BM, tl.kb_begin + i, s, two, pol_w);
            }
          }
        }
      }
      const unsigned target = epoch * (unsigned)nblk;
      unsigned v;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(barrier) : "memory");
      } while (v < target);
    }
    __syncthreads();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C_::kTmemCols) : "memory");
  }
}

// function attributes are per device: set once for every device a handle of this process launches on
template <int BN, unsigned kSet>
void ensure_attrs() {
  static std::mutex mu;
  static std::set<int> done;
  int dev = 0;
  AOCR_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(mu);
  if (done.count(dev)) return;
  AOCR_CUDA(cudaFuncSetAttribute(persist_kernel<BN, kSet>, cudaFuncAttributeMaxDynamicSharedMemorySize, PCfg<BN>::kSmemBytes));
  done.insert(dev);
}

template <int BN, unsigned kSet>
void launch_bn(Ctx& ctx, PersistProgram& prog) {
  ensure_attrs<BN, kSet>();
  const PCmd* cmds = prog.d_cmds;
  int ncmds = (int)prog.cmds.size();
  const CUtensorMap* maps = prog.d_maps;
  unsigned* bar = prog.d_barrier;
  unsigned long long* trace = prog.d_trace;
  static const int flags = getenv("AOCR_PERSIST_FLAGS") ? atoi(getenv("AOCR_PERSIST_FLAGS")) : 0;   // A/B: 1 no weight prefetch, 2 no resident tiles
  void* args[] = {(void*)&cmds, (void*)&ncmds, (void*)&maps, (void*)&bar, (void*)&trace, (void*)&flags};
  AOCR_CUDA(cudaMemsetAsync(prog.d_barrier, 0, sizeof(unsigned), ctx.st));
  // cooperative (all CTAs co-resident: the grid barrier spins) + thread-block clusters of kCluster CTAs (the fused
  // GEMM -> cell commands reduce their split-K partials through distributed shared memory inside a cluster)
  cudaLaunchConfig_t cfgl = {};
  cfgl.gridDim = dim3(prog.grid); cfgl.blockDim = dim3(256);
  cfgl.dynamicSmemBytes = (size_t)PCfg<BN>::kSmemBytes; cfgl.stream = ctx.st;
  cudaLaunchAttribute attrs[2];
  int na = 0;
  if (prog.cluster > 1) {
    attrs[na].id = cudaLaunchAttributeClusterDimension;
    attrs[na].val.clusterDim.x = prog.cluster; attrs[na].val.clusterDim.y = 1; attrs[na].val.clusterDim.z = 1;
    na++;
  }
  if (ctx.persist_coop) { attrs[na].id = cudaLaunchAttributeCooperative; attrs[na].val.cooperative = 1; na++; }
  cfgl.attrs = attrs; cfgl.numAttrs = na;
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
    case 256: return max_cluster_ctas_bn<256>(cluster);
    case 128: return max_cluster_ctas_bn<128>(cluster);
    case 64: return max_cluster_ctas_bn<64>(cluster);
    case 32: return max_cluster_ctas_bn<32>(cluster);
    default: return max_cluster_ctas_bn<16>(cluster);
  }
}

// Launch-mode probe: an EMPTY program (no commands: TMEM alloc / dealloc only) launched the way the executor will be.
// Returns cudaSuccess or the launch / execution error, with the error state cleared.  Tools that intercept launches
// (Nsight Compute) reject the cooperative + thread-block-cluster combination; the engine then picks another mode.
cudaError_t persist_probe(cudaStream_t st, int grid, int cluster, bool coop) {
  constexpr int BN = 64;
  cudaFuncSetAttribute(persist_kernel<BN, kSetEncFwd>, cudaFuncAttributeMaxDynamicSharedMemorySize, PCfg<BN>::kSmemBytes);
  const PCmd* cmds = nullptr; int ncmds = 0; const CUtensorMap* maps = nullptr; unsigned* bar = nullptr;
  unsigned long long* trace = nullptr;
  int flags = 0;
  void* args[] = {(void*)&cmds, (void*)&ncmds, (void*)&maps, (void*)&bar, (void*)&trace, (void*)&flags};
  cudaLaunchConfig_t cfgl = {};
  cfgl.gridDim = dim3(grid); cfgl.blockDim = dim3(256);
  cfgl.dynamicSmemBytes = (size_t)PCfg<BN>::kSmemBytes; cfgl.stream = st;
  cudaLaunchAttribute attrs[2];
  int na = 0;
  if (cluster > 1) {
    attrs[na].id = cudaLaunchAttributeClusterDimension;
    attrs[na].val.clusterDim.x = cluster; attrs[na].val.clusterDim.y = 1; attrs[na].val.clusterDim.z = 1;
    na++;
  }
  if (coop) { attrs[na].id = cudaLaunchAttributeCooperative; attrs[na].val.cooperative = 1; na++; }
  cfgl.attrs = attrs; cfgl.numAttrs = na;
  cudaError_t e = cudaLaunchKernelExC(&cfgl, (const void*)persist_kernel<BN, kSetEncFwd>, args);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaGetLastError();
  return e;
}

int persist_max_ctas(int bn) {
  switch (bn) {
    case 256: return max_ctas_bn<256>();
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
  if (getenv("AOCR_VARIANT")) {          // development A/B switch read by the bodies (dec_bodies.cuh)
    int v = atoi(getenv("AOCR_VARIANT"));
    cudaMemcpyToSymbol(g_variant, &v, sizeof(int));
  }
  if (getenv("AOCR_PERSIST_TRACE")) {
    int on = 1;
    cudaMemcpyToSymbol(g_bt_on, &on, sizeof(int));
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
      case 256: launch_bn<256, S>(ctx, prog); break;
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
    unsigned long long bt[16];
    if (cudaMemcpyFromSymbol(bt, g_bt, sizeof(bt)) == cudaSuccess) {
      fprintf(stderr, "[body stamps, last call, CTA 0] attn_du:");
      for (int i = 1; i <= 5; i++) fprintf(stderr, " %d->%d %.2fus", i - 1, i, (double)(bt[i] - bt[i - 1]) / 1e3);
      fprintf(stderr, " | enc_cell_bwd: loads %.2fus compute+stores %.2fus", (double)(bt[9] - bt[8]) / 1e3, (double)(bt[10] - bt[9]) / 1e3);
      fprintf(stderr, " | fused GEMM->cell: start->accumulator ready %.2fus, ->staged %.2fus, ->cluster barrier %.2fus, ->cell done %.2fus\n",
              (double)(bt[12] - bt[11]) / 1e3, (double)(bt[13] - bt[12]) / 1e3, (double)(bt[14] - bt[13]) / 1e3, (double)(bt[15] - bt[14]) / 1e3);
    }
    cudaFree(prog.d_trace); prog.d_trace = nullptr;
  }
  prog.d_cmds = nullptr; prog.d_maps = nullptr; prog.d_barrier = nullptr; prog.uploaded = false;
}

}  // namespace aocr
