// common.cuh — shared declarations of libaocr (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <stdexcept>

namespace aocr {

struct CudaError : std::runtime_error {
  explicit CudaError(const std::string& s) : std::runtime_error(s) {}
};
struct InvalidError : std::runtime_error {
  explicit InvalidError(const std::string& s) : std::runtime_error(s) {}
};

#define AOCR_CUDA(expr)                                                                     \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      char _b[512];                                                                         \
      snprintf(_b, sizeof(_b), "%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      throw ::aocr::CudaError(_b);                                                          \
    }                                                                                       \
  } while (0)

#define AOCR_CHECK(cond, msg)                                                               \
  do {                                                                                      \
    if (!(cond)) throw ::aocr::InvalidError(std::string(msg));                              \
  } while (0)

// Launch context: the engine's stream plus a launch counter (bench.py's `gpu_launches`).
struct Ctx {
  cudaStream_t st = nullptr;
  int64_t launches = 0;
  int num_sms = 148;
  // split-K workspace of the tcgen05 GEMM (partial tiles + per-tile arrival counters)
  float* tc_ws = nullptr;
  int64_t tc_ws_floats = 0;
  int* tc_counters = nullptr;
  int tc_counters_n = 0;
  bool pdl = true;   // AOCR_PDL=0 disables programmatic dependent launch
  bool persist_coop = true;   // executor launches carry the cooperative attribute (see Engine: launch-mode probe)
};

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- programmatic dependent launch (PDL): kernels of the per-timestep chains are launched with
// programmaticStreamSerialization so the next kernel's launch + prologue overlaps the tail of the current one.
// Contract: a kernel launched through launch_pdl() touches no global memory before pdl_wait().
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

#define AOCR_LAUNCH_CHECK(ctx)                \
  do {                                        \
    (ctx).launches++;                         \
    AOCR_CUDA(cudaGetLastError());            \
  } while (0)

// ---- GEMM description shared by the SIMT and tcgen05 back ends -------------------------------
// C[m][n] (ldc) = act( sum_k A(m,k) * B(k,n) + bias_n[n] + bias_m[m] ) (+ C if accumulate)
// A(m,k) = A[m*sam + k*sak], B(k,n) = B[k*sbk + n*sbn]; batched over blockIdx.z with element strides.
enum Act { ACT_NONE = 0, ACT_TANH = 1 };
struct Pack;   // gemm_tc.cuh
struct Gemm {
  int M = 0, N = 0, K = 0;
  const float* A = nullptr; int64_t sam = 0, sak = 0;
  const float* B = nullptr; int64_t sbk = 0, sbn = 0;
  float* C = nullptr; int64_t ldc = 0;
  const float* bias_n = nullptr;
  const float* bias_m = nullptr;
  int act = ACT_NONE;
  int accumulate = 0;
  int batch = 1; int64_t bsa = 0, bsb = 0, bsc = 0;
  // dW = dY^T X contractions (both operands stored [K rows][columns]): the operand already exists as bf16 planes
  // (pointer at its first column, kp = row pitch), written by the kernel that produced it - no conversion pass
  const Pack* pa = nullptr;
  const Pack* pb = nullptr;
};
void gemm_simt(Ctx& ctx, const Gemm& g);

#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
inline void launch_pdl(Ctx& ctx, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = ctx.st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = ctx.pdl ? 1 : 0;
  AOCR_CUDA(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
  ctx.launches++;
}
#endif

}  // namespace aocr
