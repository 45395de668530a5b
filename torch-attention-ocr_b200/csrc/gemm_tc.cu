// gemm_tc.cu — the tensor-core contraction of libaocr: tcgen05.mma (UMMA 128 x BN x 16, kind::f16, bf16
// operands, fp32 accumulators in TMEM) fed by TMA (cp.async.bulk.tensor, 128B swizzle) through a 3-4 stage
// mbarrier pipeline.  One CTA per 128 x BN output tile; warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM
// alloc), warps 2-5 = epilogue (tcgen05.ld -> bias/activation -> global).
//
// fp32-grade mode ("bf16x3"): every operand is stored as two bf16 planes x ~= hi + lo; each k-step issues
// hi*hi + hi*lo + lo*hi into the same TMEM accumulator, which keeps ~16 mantissa bits per operand (needed for
// the 1e-3 logit bar and tie-exact greedy decode against the float64 oracle; plain bf16 cannot meet them).
//
// Convolution = implicit GEMM: the A operand is the NHWC activation itself, addressed by a 4-D tensor map
// (C, W, H, N); for every filter tap the producer shifts the box origin by (kw-pad, kh-pad) and TMA's
// out-of-bounds zero fill supplies the padding halo.  No im2col buffer exists.
#include <cuda.h>

#include <map>
#include <mutex>
#include <set>
#include <tuple>

#include "gemm_tc.cuh"
#include "tc_ptx.cuh"

namespace aocr {

namespace {

using namespace tcp;

template <int BN> struct Cfg {
  static constexpr int kStages = (BN == 128) ? 3 : 4;
  static constexpr int kBPlane = BN * BK * 2;
  static constexpr int kStageBytes = 2 * A_PLANE_BYTES + 2 * kBPlane;
  static constexpr int kTmemCols = BN < 32 ? 32 : BN;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
};

struct TcParams {
  int M, N, K;
  int num_kb;
  int terms;
  // conv mode
  int conv;
  int mn;          // 0: both operands K-major; 1: both MN-major (K = rows), 2-D maps; 2: MN-major conv weight gradient (4-D maps)
  int wg_cin;      // mn == 2: input channels (N index = tap*Cin + ci)
  int cin_blocks, ksz, pad;
  int bw, bh, bn, tiles_w, tiles_h;
  int a_rows;      // conv mode: pixels per A box = bw*bh*bn (<= 128; the remaining tile rows are never written nor stored)
  int Nimg, Ho, Wo;
  // epilogue
  float* C;
  long long ldc;
  int transpose_out;
  const float* bias_m;
  const float* bias_n;
  int act;
  int accumulate;
  // split-K: blockIdx.z owns k-blocks [z*kb_per, (z+1)*kb_per) and writes its raw partial to ws + z*part_stride
  int splits, kb_per;
  float* ws;
  long long part_stride;
  int dbg;   // bench-only ablations: 1 = no TMA/MMA, 2 = no epilogue stores, 4 = no TMEM alloc wait path
};

template <int BN>
__global__ void __launch_bounds__(192, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
               const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
               const TcParams p) {
  using C_ = Cfg<BN>;
  pdl_launch_dependents();   // let the next kernel's prologue start; it waits on this grid before touching memory
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;              // SWIZZLE_128B tiles need 1024-byte alignment
  const uint32_t bars = base + C_::kStages * C_::kStageBytes;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (C_::kStages + s); };
  const uint32_t tmem_full_bar = bars + 8u * (2 * C_::kStages);
  const uint32_t tmem_slot = bars + 8u * (2 * C_::kStages + 1);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
  const int kb_begin = blockIdx.z * p.kb_per;
  const int kb_end = min(p.num_kb, kb_begin + p.kb_per);
  const int nkb = kb_end - kb_begin;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  // M-tile origin
  int m0 = blockIdx.y * BM;
  int cn0 = 0, ch0 = 0, cw0 = 0;
  if (p.conv) {
    int id = blockIdx.y;
    int wb = id % p.tiles_w; id /= p.tiles_w;
    int hb = id % p.tiles_h; id /= p.tiles_h;
    cn0 = id * p.bn; ch0 = hb * p.bh; cw0 = wb * p.bw;
  }

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < C_::kStages; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmAh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBh) : "memory");
  }
  if (warp == 1 && !(p.dbg & 4)) {   // TMEM allocation (whole warp, .sync.aligned)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "n"(C_::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();                // barrier init / TMEM alloc / tensor-map prefetch above overlap the previous kernel's tail

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0 && !(p.dbg & 1)) {
      const uint32_t a_bytes = p.conv ? (uint32_t)p.a_rows * (BK * 2) : (uint32_t)A_PLANE_BYTES;
      const uint32_t tx = (uint32_t)(p.terms == 3 ? 2 : 1) * (a_bytes + C_::kBPlane);
      for (int i = 0; i < nkb; i++) {
        const int kb = kb_begin + i;
        const int s = i % C_::kStages;
        const uint32_t ph = (uint32_t)(i / C_::kStages) & 1u;
        mbar_wait(empty_bar(s), ph ^ 1u);
        const uint32_t sa = base + s * C_::kStageBytes;
        const uint32_t sb = sa + 2 * A_PLANE_BYTES;
        mbar_expect_tx(full_bar(s), tx);
        if (p.mn) {
          // MN-major: k-block kb = 64 contraction rows; each operand tile = BM/64 (BN/64) boxes of [64 rows][64 elems]
          int kc1 = kb * BK, kc2 = 0, kc3 = 0;          // contraction coordinates (2-D: row ; 4-D: w, h, n)
          int bs1 = 0, bs2 = 0, bco = n0;               // B shift (conv wgrad) and B inner coordinate
          if (p.mn == 2) {
            int id = kb;
            const int wb = id % p.tiles_w; id /= p.tiles_w;
            const int hb = id % p.tiles_h; id /= p.tiles_h;
            kc1 = wb * p.bw; kc2 = hb * p.bh; kc3 = id * p.bn;
            const int tap = n0 / p.wg_cin;
            bco = n0 % p.wg_cin;
            bs1 = tap % p.ksz - p.pad; bs2 = tap / p.ksz - p.pad;
          }
          for (int pl = 0; pl < (p.terms == 3 ? 2 : 1); pl++) {
            const CUtensorMap* ta = pl ? &tmAl : &tmAh;
            const CUtensorMap* tb = pl ? &tmBl : &tmBh;
#pragma unroll
            for (int j = 0; j < BM / 64; j++) {
              const uint32_t dst = sa + pl * A_PLANE_BYTES + j * 8192;
              if (p.mn == 2) tma_load_4d(dst, ta, full_bar(s), m0 + 64 * j, kc1, kc2, kc3);
              else tma_load_2d(dst, ta, full_bar(s), m0 + 64 * j, kc1);
            }
#pragma unroll
            for (int j = 0; j < (BN >= 64 ? BN / 64 : 1); j++) {
              const uint32_t dst = sb + pl * C_::kBPlane + j * 8192;
              if (p.mn == 2) tma_load_4d(dst, tb, full_bar(s), bco + 64 * j, kc1 + bs1, kc2 + bs2, kc3);
              else tma_load_2d(dst, tb, full_bar(s), n0 + 64 * j, kc1);
            }
          }
        } else {
        if (p.conv) {
          const int tap = kb / p.cin_blocks, cb = kb % p.cin_blocks;
          const int kh = tap / p.ksz, kw = tap % p.ksz;
          tma_load_4d(sa, &tmAh, full_bar(s), cb * BK, cw0 + kw - p.pad, ch0 + kh - p.pad, cn0);
          if (p.terms == 3)
            tma_load_4d(sa + A_PLANE_BYTES, &tmAl, full_bar(s), cb * BK, cw0 + kw - p.pad, ch0 + kh - p.pad, cn0);
        } else {
          tma_load_2d(sa, &tmAh, full_bar(s), kb * BK, m0);
          if (p.terms == 3) tma_load_2d(sa + A_PLANE_BYTES, &tmAl, full_bar(s), kb * BK, m0);
        }
        tma_load_2d(sb, &tmBh, full_bar(s), kb * BK, n0);
        if (p.terms == 3) tma_load_2d(sb + C_::kBPlane, &tmBl, full_bar(s), kb * BK, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one elected lane) =====================
    const uint32_t idesc = make_idesc(BN, p.mn);
    for (int i = 0; i < ((p.dbg & 1) ? 0 : nkb); i++) {
      const int s = i % C_::kStages;
      const uint32_t ph = (uint32_t)(i / C_::kStages) & 1u;
      mbar_wait(full_bar(s), ph);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t sa = base + s * C_::kStageBytes;
        const uint32_t sb = sa + 2 * A_PLANE_BYTES;
        const uint64_t dah = p.mn ? make_desc_mnmajor_sw128(sa) : make_desc_kmajor_sw128(sa);
        const uint64_t dal = p.mn ? make_desc_mnmajor_sw128(sa + A_PLANE_BYTES) : make_desc_kmajor_sw128(sa + A_PLANE_BYTES);
        const uint64_t dbh = p.mn ? make_desc_mnmajor_sw128(sb) : make_desc_kmajor_sw128(sb);
        const uint64_t dbl = p.mn ? make_desc_mnmajor_sw128(sb + C_::kBPlane) : make_desc_kmajor_sw128(sb + C_::kBPlane);
        // per k-step (16 contraction elements): K-major +32 B inside the swizzle row; MN-major +16 rows = 2048 B
        const uint64_t kstep = p.mn ? (uint64_t)(2048 >> 4) : (uint64_t)((UMMA_K * 2) >> 4);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; k++) {
          const uint64_t adv = kstep * k;
          tc_mma(tmem_base, dah + adv, dbh + adv, idesc, (i > 0 || k > 0) ? 1u : 0u);
          if (p.terms == 3) {
            tc_mma(tmem_base, dah + adv, dbl + adv, idesc, 1u);
            tc_mma(tmem_base, dal + adv, dbh + adv, idesc, 1u);
          }
        }
        tc_commit(empty_bar(s));                       // frees the smem slot when these MMAs retire
        if (i == nkb - 1) tc_commit(tmem_full_bar);
      }
      __syncwarp();
    }
  } else {
    // ===================== epilogue: TMEM -> registers -> global =====================
    const int q = warp & 3;                            // TMEM lane quadrant this warp may access
    if (!(p.dbg & 1)) mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int ml = q * 32 + lane;                      // row inside the tile
    long long row = (long long)m0 + ml;
    bool row_ok = row < p.M;
    if (p.conv) {
      const int wl = ml % p.bw, hl = (ml / p.bw) % p.bh, nl = ml / (p.bw * p.bh);
      const int n = cn0 + nl, h = ch0 + hl, w = cw0 + wl;
      row_ok = (nl < p.bn) && (n < p.Nimg) && (h < p.Ho) && (w < p.Wo);
      row = ((long long)n * p.Ho + h) * p.Wo + w;
    }
    const float bm = (p.bias_m && row_ok) ? p.bias_m[row] : 0.f;
    // split-K: raw partial sums go to ws + z*part_stride with the indexing of C; a reduce kernel (or the consumer
    // kernel) sums the `splits` partials in fixed order, so the result is deterministic
    float* __restrict__ outp = p.splits > 1 ? p.ws + (long long)blockIdx.z * p.part_stride : p.C;
    const bool plain = (p.splits > 1);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 16) {
      uint32_t r[16];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
      if (p.dbg & 4) {
#pragma unroll
        for (int j = 0; j < 16; j++) r[j] = 0;
      } else {
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      }
      if (!row_ok) continue;
      const int nb = n0 + c0;
      float bias[16], old[16];
#pragma unroll
      for (int j = 0; j < 16; j++) { bias[j] = 0.f; old[j] = 0.f; }
      if (!plain) {        // gather every read-modify-write / bias load of this chunk first (one latency, not 16)
        if (p.accumulate) {
#pragma unroll
          for (int j = 0; j < 16; j++)
            if (nb + j < p.N)
              old[j] = p.transpose_out ? p.C[(long long)(nb + j) * p.ldc + row] : p.C[row * p.ldc + nb + j];
        }
        if (p.bias_n) {
#pragma unroll
          for (int j = 0; j < 16; j++)
            if (nb + j < p.N) bias[j] = p.bias_n[nb + j];
        }
      }
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; j++) {
        v[j] = __uint_as_float(r[j]);
        if (!plain) {
          v[j] += bm + bias[j];
          if (p.act == ACT_TANH) v[j] = tanhf(v[j]);
          v[j] += old[j];
        }
      }
      if (p.dbg & 2) continue;
      if (!p.transpose_out && nb + 16 <= p.N && (p.ldc & 3) == 0 &&
          (reinterpret_cast<uintptr_t>(outp + row * p.ldc + nb) & 15) == 0) {
        // this thread owns 16 consecutive columns of one output row: four 16-byte stores
        float4* dst = reinterpret_cast<float4*>(outp + row * p.ldc + nb);
#pragma unroll
        for (int j = 0; j < 4; j++) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 16; j++) {
          const int n = nb + j;
          if (n < p.N) {
            float* dst = p.transpose_out ? outp + (long long)n * p.ldc + row : outp + row * p.ldc + n;
            *dst = v[j];
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1 && !(p.dbg & 4)) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C_::kTmemCols) : "memory");
  }
}

// ------------------------------------------------------------------ CTA-pair kernel (cta_group::2)
// One cluster of two CTAs per 256 x BN output tile (two consecutive 128-row M tiles).  Each CTA loads its own A tile and
// HALF of the B tile (BN/2 rows), so the operand bytes per MMA flop are 2/3 (BN = 128) or 1/2 (BN = 256) of the
// single-CTA kernel's: under bf16x3 (two planes per operand, three MMAs per k-step) the single-CTA 128 x 128 tile needs
// 85 B/clk of shared-memory fill per SM, about twice what L2 delivers per SM; the pair's 256 x 256 tile needs 42.
template <int BN> struct Cfg2 {
  static constexpr int BNH = BN / 2;                       // B rows this CTA loads
  static constexpr int kBPlane = BNH * BK * 2;
  static constexpr int kStageBytes = 2 * A_PLANE_BYTES + 2 * kBPlane;
  static constexpr int kStages = (BN == 256) ? 3 : 4;
  static constexpr int kTmemCols = BN;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
};

template <int BN>
__global__ void __launch_bounds__(192, 1)
tc_gemm2_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                const TcParams p) {
  using C_ = Cfg2<BN>;
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t bars = base + C_::kStages * C_::kStageBytes;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (C_::kStages + s); };
  const uint32_t tmem_full_bar = bars + 8u * (2 * C_::kStages);
  const uint32_t tmem_slot = bars + 8u * (2 * C_::kStages + 1);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
  const int kb_begin = blockIdx.z * p.kb_per;
  const int kb_end = min(p.num_kb, kb_begin + p.kb_per);
  const int nkb = kb_end - kb_begin;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();                 // 0 = leader (issues the MMAs)
  const int n0 = blockIdx.y * BN;                          // first output column of the pair's tile
  const int nb0 = n0 + (int)rank * C_::BNH;                // first B row this CTA loads
  int m0 = blockIdx.x * BM;                                // this CTA's 128 rows (cluster = a blockIdx.x pair: cta_group::2 pairs lie along x)
  int cn0 = 0, ch0 = 0, cw0 = 0;
  if (p.conv) {
    int id = blockIdx.x;
    int wb = id % p.tiles_w; id /= p.tiles_w;
    int hb = id % p.tiles_h; id /= p.tiles_h;
    cn0 = id * p.bn; ch0 = hb * p.bh; cw0 = wb * p.bw;
  }

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < C_::kStages; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmAh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBh) : "memory");
  }
  if (warp == 1) {   // TMEM allocation of the pair (the MMA warp of each CTA)
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(C_::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();        // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer (both CTAs; completion bytes go to the leader's full barrier) ==========
    if (lane == 0) {
      const int planes = p.terms == 3 ? 2 : 1;
      const uint32_t a_bytes = p.conv ? (uint32_t)p.a_rows * (BK * 2) : (uint32_t)A_PLANE_BYTES;
      const uint32_t tx_pair = 2u * (uint32_t)planes * (a_bytes + C_::kBPlane);
      for (int i = 0; i < nkb; i++) {
        const int kb = kb_begin + i;
        const int s = i % C_::kStages;
        const uint32_t ph = (uint32_t)(i / C_::kStages) & 1u;
        mbar_wait(empty_bar(s), ph ^ 1u);
        const uint32_t sa = base + s * C_::kStageBytes;
        const uint32_t sb = sa + 2 * A_PLANE_BYTES;
        const uint32_t fb = mapa_rank(full_bar(s), 0);
        if (rank == 0) mbar_expect_tx(full_bar(s), tx_pair);
        if (p.mn) {
          int kc1 = kb * BK, kc2 = 0, kc3 = 0;
          int bs1 = 0, bs2 = 0, bco = nb0;
          if (p.mn == 2) {
            int id = kb;
            const int wb = id % p.tiles_w; id /= p.tiles_w;
            const int hb = id % p.tiles_h; id /= p.tiles_h;
            kc1 = wb * p.bw; kc2 = hb * p.bh; kc3 = id * p.bn;
            const int tap = n0 / p.wg_cin;
            bco = nb0 % p.wg_cin;
            bs1 = tap % p.ksz - p.pad; bs2 = tap / p.ksz - p.pad;
          }
          for (int pl = 0; pl < planes; pl++) {
            const CUtensorMap* ta = pl ? &tmAl : &tmAh;
            const CUtensorMap* tb = pl ? &tmBl : &tmBh;
#pragma unroll
            for (int j = 0; j < BM / 64; j++) {
              const uint32_t dst = sa + pl * A_PLANE_BYTES + j * 8192;
              if (p.mn == 2) tma2_load_4d(dst, ta, fb, m0 + 64 * j, kc1, kc2, kc3);
              else tma2_load_2d(dst, ta, fb, m0 + 64 * j, kc1);
            }
#pragma unroll
            for (int j = 0; j < C_::BNH / 64; j++) {
              const uint32_t dst = sb + pl * C_::kBPlane + j * 8192;
              if (p.mn == 2) tma2_load_4d(dst, tb, fb, bco + 64 * j, kc1 + bs1, kc2 + bs2, kc3);
              else tma2_load_2d(dst, tb, fb, nb0 + 64 * j, kc1);
            }
          }
        } else {
          if (p.conv) {
            const int tap = kb / p.cin_blocks, cb = kb % p.cin_blocks;
            const int kh = tap / p.ksz, kw = tap % p.ksz;
            tma2_load_4d(sa, &tmAh, fb, cb * BK, cw0 + kw - p.pad, ch0 + kh - p.pad, cn0);
            if (planes == 2) tma2_load_4d(sa + A_PLANE_BYTES, &tmAl, fb, cb * BK, cw0 + kw - p.pad, ch0 + kh - p.pad, cn0);
          } else {
            tma2_load_2d(sa, &tmAh, fb, kb * BK, m0);
            if (planes == 2) tma2_load_2d(sa + A_PLANE_BYTES, &tmAl, fb, kb * BK, m0);
          }
          tma2_load_2d(sb, &tmBh, fb, kb * BK, nb0);
          if (planes == 2) tma2_load_2d(sb + C_::kBPlane, &tmBl, fb, kb * BK, nb0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: the leader CTA only =====================
    if (rank == 0) {
      const uint32_t idesc = make_idesc2(BN, p.mn);
      for (int i = 0; i < nkb; i++) {
        const int s = i % C_::kStages;
        const uint32_t ph = (uint32_t)(i / C_::kStages) & 1u;
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = base + s * C_::kStageBytes;
          const uint32_t sb = sa + 2 * A_PLANE_BYTES;
          const uint64_t dah = p.mn ? make_desc_mnmajor_sw128(sa) : make_desc_kmajor_sw128(sa);
          const uint64_t dal = p.mn ? make_desc_mnmajor_sw128(sa + A_PLANE_BYTES) : make_desc_kmajor_sw128(sa + A_PLANE_BYTES);
          const uint64_t dbh = p.mn ? make_desc_mnmajor_sw128(sb) : make_desc_kmajor_sw128(sb);
          const uint64_t dbl = p.mn ? make_desc_mnmajor_sw128(sb + C_::kBPlane) : make_desc_kmajor_sw128(sb + C_::kBPlane);
          const uint64_t kstep = p.mn ? (uint64_t)(2048 >> 4) : (uint64_t)((UMMA_K * 2) >> 4);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; k++) {
            const uint64_t adv = kstep * k;
            tc2_mma(tmem_base, dah + adv, dbh + adv, idesc, (i > 0 || k > 0) ? 1u : 0u);
            if (p.terms == 3) {
              tc2_mma(tmem_base, dah + adv, dbl + adv, idesc, 1u);
              tc2_mma(tmem_base, dal + adv, dbh + adv, idesc, 1u);
            }
          }
          tc2_commit(empty_bar(s));                      // frees the slot in BOTH CTAs when these MMAs retire
          if (i == nkb - 1) tc2_commit(tmem_full_bar);   // both epilogues
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue: this CTA's 128 accumulator rows, TMEM -> registers -> global ===============
    const int q = warp & 3;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int ml = q * 32 + lane;
    long long row = (long long)m0 + ml;
    bool row_ok = row < p.M;
    if (p.conv) {
      const int wl = ml % p.bw, hl = (ml / p.bw) % p.bh, nl = ml / (p.bw * p.bh);
      const int n = cn0 + nl, h = ch0 + hl, w = cw0 + wl;
      row_ok = (nl < p.bn) && (n < p.Nimg) && (h < p.Ho) && (w < p.Wo);
      row = ((long long)n * p.Ho + h) * p.Wo + w;
    }
    const float bm = (p.bias_m && row_ok) ? p.bias_m[row] : 0.f;
    float* __restrict__ outp = p.splits > 1 ? p.ws + (long long)blockIdx.z * p.part_stride : p.C;
    const bool plain = (p.splits > 1);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 16) {
      uint32_t r[16];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (!row_ok) continue;
      const int nb = n0 + c0;
      if (nb >= p.N) continue;
      float bias[16], old[16];
#pragma unroll
      for (int j = 0; j < 16; j++) { bias[j] = 0.f; old[j] = 0.f; }
      if (!plain) {
        if (p.accumulate) {
#pragma unroll
          for (int j = 0; j < 16; j++)
            if (nb + j < p.N)
              old[j] = p.transpose_out ? p.C[(long long)(nb + j) * p.ldc + row] : p.C[row * p.ldc + nb + j];
        }
        if (p.bias_n) {
#pragma unroll
          for (int j = 0; j < 16; j++)
            if (nb + j < p.N) bias[j] = p.bias_n[nb + j];
        }
      }
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; j++) {
        v[j] = __uint_as_float(r[j]);
        if (!plain) {
          v[j] += bm + bias[j];
          if (p.act == ACT_TANH) v[j] = tanhf(v[j]);
          v[j] += old[j];
        }
      }
      if (!p.transpose_out && nb + 16 <= p.N && (p.ldc & 3) == 0 &&
          (reinterpret_cast<uintptr_t>(outp + row * p.ldc + nb) & 15) == 0) {
        float4* dst = reinterpret_cast<float4*>(outp + row * p.ldc + nb);
#pragma unroll
        for (int j = 0; j < 4; j++) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 16; j++) {
          const int n = nb + j;
          if (n < p.N) {
            float* dst = p.transpose_out ? outp + (long long)n * p.ldc + row : outp + row * p.ldc + n;
            *dst = v[j];
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  cluster_sync_all();        // both CTAs are done with the pair's shared memory and TMEM
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C_::kTmemCols) : "memory");
  }
}

// ------------------------------------------------------------------ persistent CTA-pair kernel
// The pair kernel above with a tile loop: pair j of P resident pairs works through tiles j, j + P, j + 2P, ... of the
// (M pair, N tile, split) list.  TMEM holds TWO accumulators (2 x BN columns): while the epilogue warps of both CTAs
// drain accumulator a of tile i, the producer and the MMA warp are already on tile i + 1 in accumulator a ^ 1, so the
// per-tile prologue (TMEM allocation, barrier set-up, cluster rendezvous), the TMA fill latency and the epilogue leave
// the critical path.  Used when a launch has at least four tiles per pair slot (the large convolutions at batch 256);
// measured with the loop on every launch above 74 pair tiles: conv2 forward at batch 64 (200 pair tiles of nine k-blocks)
// 54 -> 42 us, config 4 9.19 -> 8.85 ms per step.
//   tmem_full[a]  : MMA -> epilogue (tcgen05.commit multicast into both CTAs)
//   tmem_empty[a] : epilogue -> MMA, on the LEADER's barrier: 2 CTAs x 4 epilogue warps arrive (the peer's remotely)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

template <int BN>
__global__ void __launch_bounds__(192, 1)
tc_gemm2p_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                 const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                 const TcParams p, const int total_tiles, const int m_pairs, const int n_tiles) {
  using C_ = Cfg2<BN>;
  // No programmatic dependent launch around this kernel (neither the early trigger here nor the launch attribute in
  // launch2p): it holds every SM for its whole duration, and with PDL a soak of the step hung at about one in 2-3
  // thousand launches (with AOCR_PDL=0 the same soak ran clean).  The kernel after it starts when this
  // one has completed; the overlap given up is one prologue per launch.
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t bars = base + C_::kStages * C_::kStageBytes;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (C_::kStages + s); };
  auto tmem_full_bar = [&](int a) { return bars + 8u * (2 * C_::kStages + a); };
  auto tmem_empty_bar = [&](int a) { return bars + 8u * (2 * C_::kStages + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * C_::kStages + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < C_::kStages; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; a++) { mbar_init(tmem_full_bar(a), 1); mbar_init(tmem_empty_bar(a), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmAh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBh) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(2 * BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();

  // tile -> (split z, N tile, M pair); consecutive pairs take consecutive M pairs of one N tile (they share the B tile in L2)
  auto decode = [&](int tile, int& z, int& n0, int& mt) {
    const int per_z = m_pairs * n_tiles;
    z = tile / per_z;
    const int rem = tile - z * per_z;
    const int nt = rem / m_pairs;
    mt = 2 * (rem - nt * m_pairs) + (int)rank;
    n0 = nt * BN;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const int planes = p.terms == 3 ? 2 : 1;
      const uint32_t a_bytes = p.conv ? (uint32_t)p.a_rows * (BK * 2) : (uint32_t)A_PLANE_BYTES;
      const uint32_t tx_pair = 2u * (uint32_t)planes * (a_bytes + C_::kBPlane);
      uint32_t it = 0;
      for (int tile = pair; tile < total_tiles; tile += npairs) {
        int z, n0, mt;
        decode(tile, z, n0, mt);
        const int m0 = mt * BM, nb0 = n0 + (int)rank * C_::BNH;
        int cn0 = 0, ch0 = 0, cw0 = 0;
        if (p.conv) {
          int id = mt;
          const int wb = id % p.tiles_w; id /= p.tiles_w;
          const int hb = id % p.tiles_h; id /= p.tiles_h;
          cn0 = id * p.bn; ch0 = hb * p.bh; cw0 = wb * p.bw;
        }
        const int kb_begin = z * p.kb_per;
        const int nkb = min(p.num_kb, kb_begin + p.kb_per) - kb_begin;
        for (int i = 0; i < nkb; i++, it++) {
          const int kb = kb_begin + i;
          const int s = (int)(it % C_::kStages);
          const uint32_t ph = (it / C_::kStages) & 1u;
          mbar_wait(empty_bar(s), ph ^ 1u);
          const uint32_t sa = base + s * C_::kStageBytes;
          const uint32_t sb = sa + 2 * A_PLANE_BYTES;
          const uint32_t fb = mapa_rank(full_bar(s), 0);
          if (rank == 0) mbar_expect_tx(full_bar(s), tx_pair);
          if (p.mn) {
            int kc1 = kb * BK, kc2 = 0, kc3 = 0;
            int bs1 = 0, bs2 = 0, bco = nb0;
            if (p.mn == 2) {
              int id = kb;
              const int wb = id % p.tiles_w; id /= p.tiles_w;
              const int hb = id % p.tiles_h; id /= p.tiles_h;
              kc1 = wb * p.bw; kc2 = hb * p.bh; kc3 = id * p.bn;
              const int tap = n0 / p.wg_cin;
              bco = nb0 % p.wg_cin;
              bs1 = tap % p.ksz - p.pad; bs2 = tap / p.ksz - p.pad;
            }
            for (int pl = 0; pl < planes; pl++) {
              const CUtensorMap* ta = pl ? &tmAl : &tmAh;
              const CUtensorMap* tb = pl ? &tmBl : &tmBh;
#pragma unroll
              for (int j = 0; j < BM / 64; j++) {
                const uint32_t dst = sa + pl * A_PLANE_BYTES + j * 8192;
                if (p.mn == 2) tma2_load_4d(dst, ta, fb, m0 + 64 * j, kc1, kc2, kc3);
                else tma2_load_2d(dst, ta, fb, m0 + 64 * j, kc1);
              }
#pragma unroll
              for (int j = 0; j < C_::BNH / 64; j++) {
                const uint32_t dst = sb + pl * C_::kBPlane + j * 8192;
                if (p.mn == 2) tma2_load_4d(dst, tb, fb, bco + 64 * j, kc1 + bs1, kc2 + bs2, kc3);
                else tma2_load_2d(dst, tb, fb, nb0 + 64 * j, kc1);
              }
            }
          } else {
            if (p.conv) {
              const int tap = kb / p.cin_blocks, cb = kb % p.cin_blocks;
              const int kh = tap / p.ksz, kw = tap % p.ksz;
              tma2_load_4d(sa, &tmAh, fb, cb * BK, cw0 + kw - p.pad, ch0 + kh - p.pad, cn0);
              if (planes == 2) tma2_load_4d(sa + A_PLANE_BYTES, &tmAl, fb, cb * BK, cw0 + kw - p.pad, ch0 + kh - p.pad, cn0);
            } else {
              tma2_load_2d(sa, &tmAh, fb, kb * BK, m0);
              if (planes == 2) tma2_load_2d(sa + A_PLANE_BYTES, &tmAl, fb, kb * BK, m0);
            }
            tma2_load_2d(sb, &tmBh, fb, kb * BK, nb0);
            if (planes == 2) tma2_load_2d(sb + C_::kBPlane, &tmBl, fb, kb * BK, nb0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: the leader CTA only =====================
    if (rank == 0) {
      const uint32_t idesc = make_idesc2(BN, p.mn);
      uint32_t it = 0, ti = 0;
      for (int tile = pair; tile < total_tiles; tile += npairs, ti++) {
        int z, n0, mt;
        decode(tile, z, n0, mt);
        const int kb_begin = z * p.kb_per;
        const int nkb = min(p.num_kb, kb_begin + p.kb_per) - kb_begin;
        const int acc = (int)(ti & 1u);
        mbar_wait(tmem_empty_bar(acc), ((ti >> 1) & 1u) ^ 1u);      // both CTAs have drained this accumulator's previous tile
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(acc * BN);
        for (int i = 0; i < nkb; i++, it++) {
          const int s = (int)(it % C_::kStages);
          const uint32_t ph = (it / C_::kStages) & 1u;
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t sa = base + s * C_::kStageBytes;
            const uint32_t sb = sa + 2 * A_PLANE_BYTES;
            const uint64_t dah = p.mn ? make_desc_mnmajor_sw128(sa) : make_desc_kmajor_sw128(sa);
            const uint64_t dal = p.mn ? make_desc_mnmajor_sw128(sa + A_PLANE_BYTES) : make_desc_kmajor_sw128(sa + A_PLANE_BYTES);
            const uint64_t dbh = p.mn ? make_desc_mnmajor_sw128(sb) : make_desc_kmajor_sw128(sb);
            const uint64_t dbl = p.mn ? make_desc_mnmajor_sw128(sb + C_::kBPlane) : make_desc_kmajor_sw128(sb + C_::kBPlane);
            const uint64_t kstep = p.mn ? (uint64_t)(2048 >> 4) : (uint64_t)((UMMA_K * 2) >> 4);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; k++) {
              const uint64_t adv = kstep * k;
              tc2_mma(tacc, dah + adv, dbh + adv, idesc, (i > 0 || k > 0) ? 1u : 0u);
              if (p.terms == 3) {
                tc2_mma(tacc, dah + adv, dbl + adv, idesc, 1u);
                tc2_mma(tacc, dal + adv, dbh + adv, idesc, 1u);
              }
            }
            tc2_commit(empty_bar(s));
            if (i == nkb - 1) tc2_commit(tmem_full_bar(acc));
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ===================== epilogue: this CTA's 128 rows of every tile of the pair =====================
    const int q = warp & 3;
    const uint32_t te_leader0 = mapa_rank(tmem_empty_bar(0), 0), te_leader1 = mapa_rank(tmem_empty_bar(1), 0);
    uint32_t ti = 0;
    for (int tile = pair; tile < total_tiles; tile += npairs, ti++) {
      int z, n0, mt;
      decode(tile, z, n0, mt);
      const int m0 = mt * BM;
      const int acc = (int)(ti & 1u);
      mbar_wait(tmem_full_bar(acc), (ti >> 1) & 1u);
      tc_fence_after();
      const int ml = q * 32 + lane;
      long long row = (long long)m0 + ml;
      bool row_ok = row < p.M;
      if (p.conv) {
        int id = mt;
        const int wb = id % p.tiles_w; id /= p.tiles_w;
        const int hb = id % p.tiles_h; id /= p.tiles_h;
        const int cn0 = id * p.bn, ch0 = hb * p.bh, cw0 = wb * p.bw;
        const int wl = ml % p.bw, hl = (ml / p.bw) % p.bh, nl = ml / (p.bw * p.bh);
        const int n = cn0 + nl, h = ch0 + hl, w = cw0 + wl;
        row_ok = (nl < p.bn) && (n < p.Nimg) && (h < p.Ho) && (w < p.Wo);
        row = ((long long)n * p.Ho + h) * p.Wo + w;
      }
      const float bm = (p.bias_m && row_ok) ? p.bias_m[row] : 0.f;
      float* __restrict__ outp = p.splits > 1 ? p.ws + (long long)z * p.part_stride : p.C;
      const bool plain = (p.splits > 1);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 16) {
        uint32_t r[16];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c0);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (!row_ok) continue;
        const int nb = n0 + c0;
        if (nb >= p.N) continue;
        float bias[16], old[16];
#pragma unroll
        for (int j = 0; j < 16; j++) { bias[j] = 0.f; old[j] = 0.f; }
        if (!plain) {
          if (p.accumulate) {
#pragma unroll
            for (int j = 0; j < 16; j++)
              if (nb + j < p.N)
                old[j] = p.transpose_out ? p.C[(long long)(nb + j) * p.ldc + row] : p.C[row * p.ldc + nb + j];
          }
          if (p.bias_n) {
#pragma unroll
            for (int j = 0; j < 16; j++)
              if (nb + j < p.N) bias[j] = p.bias_n[nb + j];
          }
        }
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; j++) {
          v[j] = __uint_as_float(r[j]);
          if (!plain) {
            v[j] += bm + bias[j];
            if (p.act == ACT_TANH) v[j] = tanhf(v[j]);
            v[j] += old[j];
          }
        }
        if (!p.transpose_out && nb + 16 <= p.N && (p.ldc & 3) == 0 &&
            (reinterpret_cast<uintptr_t>(outp + row * p.ldc + nb) & 15) == 0) {
          float4* dst = reinterpret_cast<float4*>(outp + row * p.ldc + nb);
#pragma unroll
          for (int j = 0; j < 4; j++) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 16; j++) {
            const int n = nb + j;
            if (n < p.N) {
              float* dst = p.transpose_out ? outp + (long long)n * p.ldc + row : outp + row * p.ldc + n;
              *dst = v[j];
            }
          }
        }
      }
      // this warp's quarter of the accumulator is drained: tell the leader's MMA warp (its own or the peer's barrier)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(acc ? te_leader1 : te_leader0);
    }
  }
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * BN) : "memory");
  }
}

// ------------------------------------------------------------------ fp32 -> (hi, lo) bf16 planes
__device__ __forceinline__ void split2(float x, __nv_bfloat16& h, __nv_bfloat16& l) {
  h = __float2bfloat16_rn(x);
  l = __float2bfloat16_rn(x - __bfloat162float(h));
}
// k contiguous in the source (sks == 1): one thread per 4 consecutive k
__global__ void __launch_bounds__(256) split_kfast_kernel(const float* __restrict__ src, int64_t rows, int64_t K,
                                                          int64_t srs, int64_t kp, int64_t kw,
                                                          __nv_bfloat16* __restrict__ hi,
                                                          __nv_bfloat16* __restrict__ lo, int64_t gate_h) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t kq = kw / 4;   // kw = columns written (zero padded past K), kp = row pitch of the planes
  const int64_t total = rows * kq;
  const bool vec = ((srs & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / kq, k = (e % kq) * 4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (vec && k + 3 < K) {
      float4 t = *reinterpret_cast<const float4*>(src + r * srs + k);
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; j++)
        if (k + j < K) v[j] = src[r * srs + k + j];
    }
    __nv_bfloat16 h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; j++) split2(v[j], h[j], l[j]);
    // gate_h > 0: source rows are [gate][unit] (4 x gate_h); destination rows are gate-interleaved in blocks of 32
    // units, (unit/32)*128 + gate*32 + unit%32, so one 128-row UMMA tile holds all four gates of 32 hidden units
    const int64_t rd = gate_h > 0 ? ((r % gate_h) / 32) * 128 + (r / gate_h) * 32 + (r % gate_h) % 32 : r;
    *reinterpret_cast<uint2*>(hi + rd * kp + k) = *reinterpret_cast<uint2*>(h);
    *reinterpret_cast<uint2*>(lo + rd * kp + k) = *reinterpret_cast<uint2*>(l);
  }
}
// general strides (typically srs == 1: the transposing case): 32x32 tile through shared memory
__global__ void __launch_bounds__(256) split_tile_kernel(const float* __restrict__ src, int64_t rows, int64_t K,
                                                         int64_t srs, int64_t sks, int64_t kp, int64_t kw,
                                                         __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float tile[32][33];
  const int64_t r0 = (int64_t)blockIdx.y * 32, k0 = (int64_t)blockIdx.x * 32;
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;   // 32 x 8
  // read with the source-contiguous index on tx
  const bool rows_fast = (srs == 1);
  for (int i = ty; i < 32; i += 8) {
    int64_t r = rows_fast ? r0 + tx : r0 + i;
    int64_t k = rows_fast ? k0 + i : k0 + tx;
    float v = (r < rows && k < K) ? src[r * srs + k * sks] : 0.f;
    if (rows_fast) tile[tx][i] = v; else tile[i][tx] = v;     // tile[row][k]
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    int64_t r = r0 + i, k = k0 + tx;
    if (r < rows && k < kw) {
      __nv_bfloat16 h, l;
      split2(tile[i][tx], h, l);
      hi[r * kp + k] = h;
      lo[r * kp + k] = l;
    }
  }
}

// ------------------------------------------------------------------ split-K reduction + epilogue
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ ws, int nz, long long part_stride,
                                                            float* __restrict__ C, long long rows, long long cols,
                                                            long long ldc, const float* __restrict__ bias_r,
                                                            const float* __restrict__ bias_c, int act, int accumulate) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = rows * cols;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / cols, c = e % cols;
    const long long off = r * ldc + c;
    float v = 0.f;
    for (int z0 = 0; z0 < nz; z0 += 8) {     // 8 independent loads in flight, then add (fixed order)
      float t[8];
#pragma unroll
      for (int j = 0; j < 8; j++) t[j] = (z0 + j < nz) ? __ldcg(ws + (long long)(z0 + j) * part_stride + off) : 0.f;
#pragma unroll
      for (int j = 0; j < 8; j++) v += t[j];
    }
    if (bias_r) v += bias_r[r];
    if (bias_c) v += bias_c[c];
    if (act == ACT_TANH) v = tanhf(v);
    if (accumulate) v += C[off];
    C[off] = v;
  }
}

// ------------------------------------------------------------------ tensor map cache
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
std::once_flag g_encode_once;

void resolve_encode() {
  std::call_once(g_encode_once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  });
}

typedef std::tuple<const void*, int64_t, int64_t, int64_t, int64_t, int, int, int, int> MapKey;
std::map<MapKey, CUtensorMap> g_maps;
std::mutex g_maps_mu;

// 2-D K-major plane [rows][kp], box = 64 x box_rows
const CUtensorMap& map_2d(const __nv_bfloat16* ptr, int64_t rows, int64_t kp, int box_rows) {
  MapKey key(ptr, rows, kp, 0, 0, box_rows, 0, 0, 2);
  std::lock_guard<std::mutex> lk(g_maps_mu);
  auto it = g_maps.find(key);
  if (it != g_maps.end()) return it->second;
  CUtensorMap tm;
  cuuint64_t dims[2] = {(cuuint64_t)kp, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)kp * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = g_encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(ptr), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw CudaError("cuTensorMapEncodeTiled(2d) failed: " + std::to_string((int)r));
  return g_maps.emplace(key, tm).first->second;
}
// 4-D NHWC activation (C, W, H, N), box = 64 x bw x bh x bn
const CUtensorMap& map_4d(const __nv_bfloat16* ptr, int N, int H, int W, int C, int bw, int bh, int bn) {
  MapKey key(ptr, N, H, W, C, bw, bh, bn, 4);
  std::lock_guard<std::mutex> lk(g_maps_mu);
  auto it = g_maps.find(key);
  if (it != g_maps.end()) return it->second;
  CUtensorMap tm;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = g_encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(ptr), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw CudaError("cuTensorMapEncodeTiled(4d) failed: " + std::to_string((int)r));
  return g_maps.emplace(key, tm).first->second;
}

template <int BN>
void launch(Ctx& ctx, const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& bh, const CUtensorMap& bl,
            const TcParams& p, dim3 grid) {
  {   // function attributes are per device: once for every device a handle of this process launches on
    static std::mutex mu;
    static std::set<int> done;
    int dev = 0;
    AOCR_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(mu);
    if (!done.count(dev)) {
      AOCR_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN>::kSmemBytes));
      done.insert(dev);
    }
  }
  launch_pdl(ctx, tc_gemm_kernel<BN>, grid, dim3(192), (p.dbg & 8) ? (size_t)4096 : (size_t)Cfg<BN>::kSmemBytes, ah, al, bh, bl, p);
  AOCR_CUDA(cudaGetLastError());
}

template <int BN>
void launch2(Ctx& ctx, const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& bh, const CUtensorMap& bl,
             const TcParams& p, dim3 grid) {
  {
    static std::mutex mu;
    static std::set<int> done;
    int dev = 0;
    AOCR_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(mu);
    if (!done.count(dev)) {
      AOCR_CUDA(cudaFuncSetAttribute(tc_gemm2_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg2<BN>::kSmemBytes));
      done.insert(dev);
    }
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid.y, grid.x, grid.z);      // M tiles along x: the CTA pair is a pair of consecutive blockIdx.x
  cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = (size_t)Cfg2<BN>::kSmemBytes; cfg.stream = ctx.st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = ctx.pdl ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, tc_gemm2_kernel<BN>, ah, al, bh, bl, p);
  if (e != cudaSuccess) {
    int ncl = -1;
    cudaGetLastError();
    cudaError_t e2 = cudaOccupancyMaxActiveClusters(&ncl, (const void*)tc_gemm2_kernel<BN>, &cfg);
    fprintf(stderr, "[aocr] pair-kernel launch failed: %s; grid (%u,%u,%u) smem %zu; max active clusters %d (%s)\n",
            cudaGetErrorString(e), grid.x, grid.y, grid.z, cfg.dynamicSmemBytes, ncl, cudaGetErrorString(e2));
    AOCR_CUDA(e);
  }
  ctx.launches++;
}

template <int BN>
void launch2p(Ctx& ctx, const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& bh, const CUtensorMap& bl,
              const TcParams& p, dim3 grid, int npairs) {
  {
    static std::mutex mu;
    static std::set<int> done;
    int dev = 0;
    AOCR_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(mu);
    if (!done.count(dev)) {
      AOCR_CUDA(cudaFuncSetAttribute(tc_gemm2p_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg2<BN>::kSmemBytes));
      done.insert(dev);
    }
  }
  const int m_pairs = (int)grid.y / 2, n_tiles = (int)grid.x, total = m_pairs * n_tiles * (int)grid.z;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * (unsigned)npairs, 1, 1);
  cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = (size_t)Cfg2<BN>::kSmemBytes; cfg.stream = ctx.st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;          // cluster dimension only: no PDL for the persistent kernel (see the kernel)
  AOCR_CUDA(cudaLaunchKernelEx(&cfg, tc_gemm2p_kernel<BN>, ah, al, bh, bl, p, total, m_pairs, n_tiles));
  ctx.launches++;
}

}  // namespace

const CUtensorMap& tc_map_2d(const __nv_bfloat16* ptr, int64_t rows, int64_t kp, int box_rows) {
  resolve_encode();
  if (!g_encode) throw CudaError("cuTensorMapEncodeTiled entry point not found (driver too old?)");
  return map_2d(ptr, rows, kp, box_rows);
}

bool gemm_tc_available() {
  resolve_encode();
  return g_encode != nullptr;
}

void split_to_pack(Ctx& ctx, const float* src, int64_t rows, int64_t K, int64_t srs, int64_t sks, const Pack& dst,
                   int64_t kwrite, int64_t gate_h) {
  AOCR_CHECK(gate_h == 0 || (sks == 1 && rows == 4 * gate_h && gate_h % 32 == 0), "split_to_pack: bad gate interleave request");
  int64_t kw = kwrite;
  if (kw < 0) {
    AOCR_CHECK(dst.kp == pad64(K), "split_to_pack: destination pitch must be K rounded up to 64");
    kw = dst.kp;
  }
  AOCR_CHECK(dst.rows >= rows && kw >= K && kw % 4 == 0 && kw <= dst.kp, "split_to_pack: bad destination extent");
  if (sks == 1) {
    int64_t total = rows * (kw / 4);
    int64_t g = (total + 255) / 256;
    int64_t cap = (int64_t)ctx.num_sms * 8;
    launch_pdl(ctx, split_kfast_kernel, dim3((unsigned)(g < cap ? (g > 0 ? g : 1) : cap)), dim3(256), 0, src, rows, K, srs,
               dst.kp, kw, dst.hi, dst.lo, gate_h);
  } else {
    dim3 grid((unsigned)((kw + 31) / 32), (unsigned)((rows + 31) / 32));
    launch_pdl(ctx, split_tile_kernel, grid, dim3(256), 0, src, rows, K, srs, sks, dst.kp, kw, dst.hi, dst.lo);
  }
  AOCR_CUDA(cudaGetLastError());
}

TcOut gemm_tc(Ctx& ctx, const TcGemm& g) {
  resolve_encode();
  if (!g_encode) throw CudaError("cuTensorMapEncodeTiled entry point not found (driver too old?)");
  AOCR_CHECK(g.M > 0 && g.N > 0 && g.K > 0, "gemm_tc: empty problem");
  int BN = g.N > 64 ? 128 : (g.N > 32 ? 64 : (g.N > 16 ? 32 : 16));
  if (g.mn && BN < 64) BN = 64;                     // MN-major boxes are 64 elements wide
  if (g.mn == 2 && g.conv && g.conv->C < 128) BN = 64;   // an N tile must not straddle two filter taps
  // CTA-pair kernel (256 x BN tiles, cta_group::2): whenever there are at least two M tiles and N fills a 128-wide tile
  const bool pair_on = !(getenv("AOCR_CG2") && atoi(getenv("AOCR_CG2")) == 0);      // read per call: an A/B switch
  const int pair_maxbn = getenv("AOCR_CG2_BN") ? atoi(getenv("AOCR_CG2_BN")) : 256;
  // (BN = 64 as well when the operands are K-major - conv2's data gradient has N = Cin = 64 and 400+ M tiles; the
  // MN-major modes need whole 64-column boxes per CTA)
  bool pair = pair_on && (BN == 128 || (BN == 64 && g.mn == 0 && g.N > 32 && g.M > 8 * BM)) && g.M > BM && !(g.dbg);
  if (pair && pair_maxbn >= 256 && g.N >= 256 && (g.mn != 2 || g.conv->C % 256 == 0)) BN = 256;
  const int BNB = pair ? BN / 2 : BN;               // B rows one CTA loads
  TcParams p{};
  p.M = g.M; p.N = g.N; p.K = g.K; p.terms = g.terms; p.mn = g.mn;
  p.C = g.C; p.ldc = g.ldc; p.transpose_out = g.transpose_out ? 1 : 0;
  p.bias_m = g.bias_m; p.bias_n = g.bias_n; p.act = g.act; p.accumulate = g.accumulate;
  dim3 grid;
  grid.x = (g.N + BN - 1) / BN;
  const CUtensorMap *ah, *al, *bhp, *blp;
  if (g.mn == 2) {
    // convolution weight gradient, implicit: dW[co][tap*Cin+ci] = sum_pixels dz[pix][co] * x[pix+tap][ci]
    const ConvView& c = *g.conv;
    AOCR_CHECK(c.C % 64 == 0 && g.B.kp == c.C && g.A.kp >= g.M && g.M % 64 == 0, "conv wgrad: channels must be multiples of 64");
    AOCR_CHECK(g.N == c.k * c.k * c.C && c.C % BN == 0, "conv wgrad: N must be k*k*Cin");
    AOCR_CHECK(!pair || BNB % 64 == 0, "pair kernel: MN-major B half must be whole 64-column boxes");
    // contraction block = 64 output pixels as a (bw x bh x bn) box.  Rows of the block that the box does not cover would
    // have to be zero, so only EXACT covers (bw*bh*bn == 64, every dimension divided evenly) replace the power-of-two
    // box: a 25-wide map takes 1 x 8 x 8 boxes (no padded pixels) instead of 32 x 2 x 1 (22 % zero-filled pixels).
    int bw = 8;
    while (bw < c.Wo && bw < 64) bw *= 2;
    int bh = 1;
    while (bh * 2 * bw <= 64 && bh < c.Ho) bh *= 2;
    int bn = 64 / (bw * bh);
    if (!(getenv("AOCR_BOX_POW2") && atoi(getenv("AOCR_BOX_POW2")))) {
      long long best = (long long)((c.Wo + bw - 1) / bw) * ((c.Ho + bh - 1) / bh) * ((c.N + bn - 1) / bn);
      for (int w = 1; w <= 64 && w <= c.Wo; w++) {
        if (c.Wo % w) continue;
        for (int h = 1; w * h <= 64 && h <= c.Ho; h++) {
          if (c.Ho % h || 64 % (w * h)) continue;
          const int n = 64 / (w * h);
          if (c.N % n) continue;
          const long long blocks = (long long)(c.Wo / w) * (c.Ho / h) * (c.N / n);
          if (blocks < best || (blocks == best && w > bw)) { best = blocks; bw = w; bh = h; bn = n; }
        }
      }
    }
    p.bw = bw; p.bh = bh; p.bn = bn; p.ksz = c.k; p.pad = c.pad; p.wg_cin = c.C;
    p.tiles_w = (c.Wo + bw - 1) / bw; p.tiles_h = (c.Ho + bh - 1) / bh;
    p.num_kb = p.tiles_w * p.tiles_h * ((c.N + bn - 1) / bn);
    grid.y = (g.M + BM - 1) / BM;
    ah = &map_4d(g.A.hi, c.N, c.Ho, c.Wo, (int)g.A.kp, bw, bh, bn);
    al = &map_4d(g.A.lo, c.N, c.Ho, c.Wo, (int)g.A.kp, bw, bh, bn);
    bhp = &map_4d(g.B.hi, c.N, c.H, c.W, c.C, bw, bh, bn);
    blp = &map_4d(g.B.lo, c.N, c.H, c.W, c.C, bw, bh, bn);
  } else if (g.mn == 1) {
    // both operands stored [K rows][M or N columns]
    AOCR_CHECK(g.A.rows >= g.K && g.B.rows >= g.K && g.A.kp >= g.M && g.B.kp >= g.N, "gemm_tc(mn): pack too small");
    p.num_kb = (g.K + BK - 1) / BK;
    grid.y = (g.M + BM - 1) / BM;
    ah = &map_2d(g.A.hi, g.K, g.A.kp, 64);
    al = &map_2d(g.A.lo, g.K, g.A.kp, 64);
    bhp = &map_2d(g.B.hi, g.K, g.B.kp, 64);
    blp = &map_2d(g.B.lo, g.K, g.B.kp, 64);
  } else {
  if (g.conv) {
    const ConvView& c = *g.conv;
    AOCR_CHECK(c.C % BK == 0 && g.A.kp == c.C, "conv A pack must be NHWC with C a multiple of 64");
    AOCR_CHECK(g.K == c.k * c.k * c.C && g.B.kp == pad64(g.K), "conv weight pack must be [Cout][k*k*C]");
    // pixel box (bw x bh x bn <= 128 rows of the M tile): the one that covers the output with the fewest tiles.  TMA
    // boxes need not be powers of two: a 25-wide map takes 25 x 1 x 5 boxes (125 of 128 rows live) instead of 32 x 4 x 1
    // (100 of 128); rows past the box are never written by TMA and never stored by the epilogue.
    int bw = 1, bh = 1, bn = 1;
    {
      const bool pow2 = getenv("AOCR_BOX_POW2") && atoi(getenv("AOCR_BOX_POW2"));
      long long best = -1;
      for (int w = 1; w <= (c.Wo < 128 ? c.Wo : 128); w++) {
        if (pow2 && (w & (w - 1)) && w != c.Wo) continue;
        for (int h = 1; h <= c.Ho && w * h <= 128; h++) {
          int n = 128 / (w * h);
          if (n > c.N) n = c.N;
          const long long tiles = (long long)((c.Wo + w - 1) / w) * ((c.Ho + h - 1) / h) * ((c.N + n - 1) / n);
          if (best < 0 || tiles < best || (tiles == best && w * h * n > bw * bh * bn)) { best = tiles; bw = w; bh = h; bn = n; }
        }
      }
      if (pow2) {
        bw = 8;
        while (bw < c.Wo && bw < 128) bw *= 2;
        bh = 1;
        while (bh * 2 * bw <= 128 && bh < c.Ho) bh *= 2;
        bn = 128 / (bw * bh);
      }
    }
    p.conv = 1; p.cin_blocks = c.C / BK; p.ksz = c.k; p.pad = c.pad;
    p.bw = bw; p.bh = bh; p.bn = bn; p.a_rows = bw * bh * bn;
    p.tiles_w = (c.Wo + bw - 1) / bw; p.tiles_h = (c.Ho + bh - 1) / bh;
    p.Nimg = c.N; p.Ho = c.Ho; p.Wo = c.Wo;
    p.num_kb = c.k * c.k * p.cin_blocks;
    grid.y = p.tiles_w * p.tiles_h * ((c.N + bn - 1) / bn);
    ah = &map_4d(g.A.hi, c.N, c.H, c.W, c.C, bw, bh, bn);
    al = &map_4d(g.A.lo, c.N, c.H, c.W, c.C, bw, bh, bn);
  } else {
    AOCR_CHECK(g.A.kp >= pad64(g.K) && g.B.kp >= pad64(g.K), "gemm_tc: operand pack narrower than K");
    p.num_kb = (int)(pad64(g.K) / BK);
    grid.y = (g.M + BM - 1) / BM;
    ah = &map_2d(g.A.hi, g.A.rows, g.A.kp, BM);
    al = &map_2d(g.A.lo, g.A.rows, g.A.kp, BM);
  }
  bhp = &map_2d(g.B.hi, g.B.rows, g.B.kp, BNB);
  blp = &map_2d(g.B.lo, g.B.rows, g.B.kp, BNB);
  }
  if (pair) grid.y = (grid.y + 1) & ~1u;            // whole pairs: a padding CTA loads zero-filled tiles and stores nothing
  const CUtensorMap& bh_ = *bhp;
  const CUtensorMap& bl_ = *blp;
  // split-K when the output tiles alone cannot fill the machine (per-timestep decoder GEMMs, weight gradients)
  const long long tiles = (long long)grid.x * grid.y;
  const long long crows = g.transpose_out ? g.N : g.M, ccols = g.transpose_out ? g.M : g.N;
  const long long part_stride = ((crows - 1) * g.ldc + ccols + 63) & ~63LL;
  int splits = 1;
  if (ctx.tc_ws && tiles * 2 <= ctx.num_sms && p.num_kb >= 4) {
    splits = (int)(ctx.num_sms / tiles);
    if (splits > p.num_kb / 2) splits = p.num_kb / 2;
    // consumers of deferred partials sum at most 8 (decb::kMaxSplits); the stand-alone reduction takes any number
    const int cap = g.defer_reduce ? 8 : 16;
    if (splits > cap) splits = cap;
  }
  if (g.force_splits > 0 && ctx.tc_ws) splits = g.force_splits < p.num_kb ? g.force_splits : p.num_kb;
  float* wsbase = g.ws ? g.ws : ctx.tc_ws;
  const long long wscap = g.ws ? g.ws_floats : ctx.tc_ws_floats;
  while (splits > 1 && (long long)splits * part_stride > wscap) splits--;
  if (splits < 1) splits = 1;
  p.kb_per = (p.num_kb + splits - 1) / splits;
  p.splits = (p.num_kb + p.kb_per - 1) / p.kb_per;
  p.ws = wsbase; p.part_stride = part_stride; p.dbg = g.dbg;
  if (g.defer_reduce) {   // raw sums only: the consumer kernel adds the partials (and any bias / activation)
    AOCR_CHECK(!g.bias_m && !g.bias_n && g.act == ACT_NONE && !g.accumulate && g.ws, "deferred GEMM takes no epilogue");
    AOCR_CHECK(part_stride <= wscap, "deferred GEMM workspace too small");
    p.C = wsbase;        // splits == 1 writes the single partial straight into the workspace
  }
  grid.z = p.splits;
  // many more tiles than resident pairs (>= 4 per pair): the persistent variant (tile loop, double-buffered accumulator).
  // At batch 64 no launch of the step qualifies (conv2: 200 pair tiles on 74 slots gains 0.4 % of the step); at batch
  // 256 the large convolutions do.
  const int pair_slots = ctx.num_sms / 2;
  const long long pair_tiles = pair ? (long long)(grid.y / 2) * grid.x * grid.z : 0;
  // AOCR_PERSIST_GEMM = minimum number of tiles per pair slot (default 4; 0 = never; 2 = the setting the batch-64
  // measurement above was taken with, also the stress setting: it puts conv2 of every batch-64 step on the tile loop)
  const int per_slot = getenv("AOCR_PERSIST_GEMM") ? atoi(getenv("AOCR_PERSIST_GEMM")) : 4;
  const bool persistent = pair && per_slot > 0 && pair_tiles >= (long long)(per_slot == 1 ? 4 : per_slot) * pair_slots;
  if (persistent) {
    if (BN == 256) launch2p<256>(ctx, *ah, *al, bh_, bl_, p, grid, pair_slots);
    else if (BN == 128) launch2p<128>(ctx, *ah, *al, bh_, bl_, p, grid, pair_slots);
    else launch2p<64>(ctx, *ah, *al, bh_, bl_, p, grid, pair_slots);
  } else if (pair) {
    if (BN == 256) launch2<256>(ctx, *ah, *al, bh_, bl_, p, grid);
    else if (BN == 128) launch2<128>(ctx, *ah, *al, bh_, bl_, p, grid);
    else launch2<64>(ctx, *ah, *al, bh_, bl_, p, grid);
  } else
  switch (BN) {
    case 128: launch<128>(ctx, *ah, *al, bh_, bl_, p, grid); break;
    case 64: launch<64>(ctx, *ah, *al, bh_, bl_, p, grid); break;
    case 32: launch<32>(ctx, *ah, *al, bh_, bl_, p, grid); break;
    default: launch<16>(ctx, *ah, *al, bh_, bl_, p, grid); break;
  }
  TcOut out;
  out.base = (p.splits > 1 || g.defer_reduce) ? wsbase : g.C;
  out.nz = p.splits; out.stride = part_stride;
  if (p.splits > 1 && !g.defer_reduce) {
    const float* bias_r = g.transpose_out ? g.bias_n : g.bias_m;
    const float* bias_c = g.transpose_out ? g.bias_m : g.bias_n;
    const long long total = crows * ccols;
    long long nb = (total + 255) / 256;
    if (nb > (long long)ctx.num_sms * 8) nb = (long long)ctx.num_sms * 8;
    launch_pdl(ctx, splitk_reduce_kernel, dim3((unsigned)nb), dim3(256), 0, (const float*)wsbase, p.splits, part_stride,
               g.C, crows, ccols, (long long)g.ldc, bias_r, bias_c, g.act, g.accumulate);
    AOCR_CUDA(cudaGetLastError());
  }
  return out;
}

}  // namespace aocr
