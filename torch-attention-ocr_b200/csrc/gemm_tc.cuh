// gemm_tc.cuh — host interface of the tcgen05/TMEM/TMA GEMM core (definitions in gemm_tc.cu).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace aocr {

// A GEMM operand as bf16 planes: x ~= hi + lo (hi = bf16(x), lo = bf16(x - hi)).  `rows` x `kp` row-major
// ("K-major"), kp = K rounded up to 64 and zero-padded, so every row is a whole number of 128-byte TMA boxes.
struct Pack {
  __nv_bfloat16* hi = nullptr;
  __nv_bfloat16* lo = nullptr;
  int64_t rows = 0, kp = 0;
};
inline int64_t pad64(int64_t k) { return (k + 63) & ~(int64_t)63; }

// fp32 (rows x K, element (r,k) at src[r*srs + k*sks]) -> Pack.  Handles either stride being 1 with
// coalesced access (the transposing case goes through a shared-memory tile).
// kwrite < 0: write the whole padded row (dst.kp == pad64(K)); otherwise write exactly `kwrite` (>= K, % 4 == 0)
// columns of a wider pack (used to assemble concatenated weight packs).
void split_to_pack(Ctx& ctx, const float* src, int64_t rows, int64_t K, int64_t srs, int64_t sks, const Pack& dst,
                   int64_t kwrite = -1, int64_t gate_h = 0);

// NHWC view of a Pack for the implicit-GEMM convolution: rows = n*H*W pixels, kp = C channels.
struct ConvView {
  int N = 0, H = 0, W = 0, C = 0;      // input activation
  int k = 0, pad = 0;                  // kernel size, padding
  int Ho = 0, Wo = 0;                  // output spatial dims
};

struct TcGemm {
  // D[m][n] = sum_k A[m][k] * B[n][k]; A is the M side (128-row tiles), B the N side.
  Pack A, B;
  int M = 0, N = 0, K = 0;             // logical sizes (K <= A.kp == B.kp), or K = k*k*C in conv mode
  const ConvView* conv = nullptr;      // if set: A is an NHWC activation pack, M = N*Ho*Wo output pixels
  // mn = 1: both operands are stored [K rows][M|N columns] ("MN-major": weight gradients dW = dY^T X need no
  // transposes).  mn = 2: implicit convolution weight gradient: A = dz NHWC pack (N*Ho*Wo x Cout), B = x NHWC pack
  // (N*H*W x Cin), conv = geometry of x; M = Cout, N = k*k*Cin, K = output pixels.
  int mn = 0;
  float* C = nullptr; int64_t ldc = 0;
  bool transpose_out = false;          // false: C[m*ldc+n] ; true: C[n*ldc+m]
  const float* bias_m = nullptr;       // indexed by m
  const float* bias_n = nullptr;       // indexed by n
  int act = ACT_NONE;
  int accumulate = 0;
  int terms = 3;                       // 3: hi*hi + hi*lo + lo*hi (fp32-grade) ; 1: hi*hi (plain bf16)
  int force_splits = 0;                // test hook: force a split-K factor
  int dbg = 0;                         // bench-only ablation bits (see TcParams::dbg)
  // deferred reduction: leave the split-K partial sums in `ws` (ws_floats capacity) for the consumer kernel
  bool defer_reduce = false;
  float* ws = nullptr;
  int64_t ws_floats = 0;
};
// where the result lives: value(i) = sum_{z < nz} base[z*stride + i], i indexed like C (ldc / transpose_out)
struct TcOut { const float* base = nullptr; int nz = 1; int64_t stride = 0; };
TcOut gemm_tc(Ctx& ctx, const TcGemm& g);
bool gemm_tc_available();
// cached 2-D tensor map of one bf16 plane [rows][kp] with a 64 x box_rows box, 128B swizzle (also used by persist.cu)
const CUtensorMap& tc_map_2d(const __nv_bfloat16* ptr, int64_t rows, int64_t kp, int box_rows);   // driver entry point for cuTensorMapEncodeTiled resolved

}  // namespace aocr
