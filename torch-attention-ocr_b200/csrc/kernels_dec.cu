// kernels_dec.cu — the memory-bound kernels of the tensor-core decoder path.  They sit between the per-timestep
// tcgen05 GEMMs: each one CONSUMES the split-K partial sums of the GEMM before it (summing them in fixed order,
// so no separate reduction pass and a deterministic result) and PRODUCES the next GEMM's operand directly as
// bf16 (hi, lo) planes next to the fp32 copy kept for backward.  Math: src/model/LSTM.lua:79-105,124-162.
#include "dec_bodies.cuh"

namespace aocr {

namespace {

inline int grid_for(int64_t total, int threads, int num_sms) {
  int64_t g = (total + threads - 1) / threads;
  int64_t cap = (int64_t)num_sms * 8;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

#define AOCR_WRAP(NAME, PTYPE)                                                     \
  __global__ void __launch_bounds__(256) NAME##_kernel(PTYPE p) {                  \
    pdl_launch_dependents();                                                       \
    pdl_wait();                                                                    \
    decb::NAME##_body(p, (int)blockIdx.x, (int)gridDim.x, nullptr);                \
  }
AOCR_WRAP(cell_fwd_tc, CellFwdTc)
AOCR_WRAP(cell_bwd_tc, CellBwdTc)
AOCR_WRAP(enc_cell_fwd_tc, EncCellFwdTc)
AOCR_WRAP(enc_cell_bwd_tc, EncCellBwdTc)

__global__ void __launch_bounds__(256) attn_out_tc_kernel(AttnOutTc p) {
  extern __shared__ float sm[];
  pdl_launch_dependents();
  pdl_wait();
  decb::attn_out_tc_body(p, (int)blockIdx.x, sm);
}
__global__ void __launch_bounds__(256) attn_du_tc_kernel(AttnDuTc p) {
  extern __shared__ float sm[];
  pdl_launch_dependents();
  pdl_wait();
  decb::attn_du_tc_body(p, (int)blockIdx.x, sm);
}
__global__ void part_to_dense_kernel(PartIn in, float* dst, int64_t ld, int B, int cols) {
  pdl_launch_dependents();
  pdl_wait();
  decb::part_to_dense_body(in, dst, ld, B, cols, (int)blockIdx.x, (int)gridDim.x, nullptr);
}

}  // namespace

void part_to_dense(Ctx& ctx, const PartIn& in, float* dst, int64_t ld, int B, int cols) {
  launch_pdl(ctx, part_to_dense_kernel, dim3(grid_for((int64_t)B * cols, 256, ctx.num_sms)), dim3(256), 0, in, dst, ld, B, cols);
  AOCR_CUDA(cudaGetLastError());
}
#define AOCR_LAUNCH_EW(NAME, PTYPE, TOTAL)                                                                   \
  void NAME(Ctx& ctx, const PTYPE& p) {                                                                      \
    launch_pdl(ctx, NAME##_kernel, dim3(grid_for((TOTAL), 256, ctx.num_sms)), dim3(256), 0, p);              \
    AOCR_CUDA(cudaGetLastError());                                                                           \
  }
AOCR_LAUNCH_EW(cell_fwd_tc, CellFwdTc, (int64_t)p.B * p.H)
AOCR_LAUNCH_EW(cell_bwd_tc, CellBwdTc, (int64_t)p.B * p.H)
AOCR_LAUNCH_EW(enc_cell_fwd_tc, EncCellFwdTc, (int64_t)(p.d_only < 0 ? 2 : 1) * p.B * p.He)
AOCR_LAUNCH_EW(enc_cell_bwd_tc, EncCellBwdTc, (int64_t)(p.d_only < 0 ? 2 : 1) * p.B * p.He)

void attn_out_tc(Ctx& ctx, const AttnOutTc& p) {
  AOCR_CHECK(p.H % 128 == 0 && p.H <= 1024, "attention kernel needs decoder hidden size in {128,...,1024}");
  launch_pdl(ctx, attn_out_tc_kernel, dim3(p.B), dim3(256), attn_smem_bytes(p.S, p.H), p);
  AOCR_CUDA(cudaGetLastError());
}
void attn_du_tc(Ctx& ctx, const AttnDuTc& p) {
  AOCR_CHECK(p.H % 128 == 0 && p.H <= 1024, "attention kernel needs decoder hidden size in {128,...,1024}");
  launch_pdl(ctx, attn_du_tc_kernel, dim3(p.B), dim3(256), attn_smem_bytes(p.S, p.H), p);
  AOCR_CUDA(cudaGetLastError());
}

}  // namespace aocr
