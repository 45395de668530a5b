// persist.h — the persistent recurrence executor: ONE kernel (cooperative launch, thread-block clusters) runs a whole
// recurrence (all timesteps of the decoder forward, the decoder backward, one encoder direction, or the dual greedy
// decode pass) as a command list.  A command is a swap-AB tcgen05 GEMM (the CTA's tile of `weights x batch`, split-K
// partials to global memory), a fused GEMM -> LSTM-cell (the cluster owns the M tile, split-K partials reduced through
// distributed shared memory), or one of the cell / attention+output / generator bodies of dec_bodies.cuh; a grid
// barrier separates consecutive commands.  Kernel boundaries (4-5 us each on this part) disappear; TMEM, the mbarrier
// ring and the tensor-map table live for the whole recurrence.  The command list is built once per (batch, S, T) shape
// on the host and cached on the device.
#pragma once
#include <vector>

#include "gemm_tc.cuh"
#include "kernels_dec.h"

namespace aocr {

enum PType {
  P_GEMM = 0, P_CELL_FWD, P_CELL_BWD, P_ATTN_OUT, P_ATTN_DU, P_ENC_CELL_FWD, P_ENC_CELL_BWD, P_TO_DENSE, P_GENERATOR, P_GREEDY,
  P_GEMM_ENC_FWD,     // fused: PGemm immediately followed by the EncCellFwdTc of the same step
  P_GEMM_CELL_FWD,    // fused: PGemm + CellFwdTc (decoder layer)
  P_ATTN_OUT_GEN      // AttnOutTc + GenTc + GreedyTc: attention+output, generator and selection of a dual decode step
};

// swap-AB GEMM with the batch on the UMMA N side: partial z of out(n, m) at ws[z*part_stride + n*ldc + m]
struct PGemm {
  int map_a, map_b;        // index of the hi plane's tensor map in the table (lo plane = +1)
  int m_tiles, splits, kb_per, num_kb;
  int b_row0;              // first pack row of the activation operand (timestep * batch)
  int b_k0;                // K offset (elements) inside the activation pack
  int M, N, terms;
  float* ws; long long part_stride; long long ldc;
};
struct PToDense { PartIn in; float* dst; long long ld; int B, cols; };

struct alignas(16) PCmd {
  int type;
  int pad[3];
  unsigned char payload[368];
};

struct PersistProgram {                 // host-side, then uploaded
  std::vector<PCmd> cmds;
  std::vector<CUtensorMap> maps;
  int grid = 1;                         // CTAs (>= the largest number of GEMM tiles of any command)
  int bn = 64;                          // UMMA N = batch rounded up to {16,32,64,128,256}
  int cluster = 1;                      // thread-block cluster size of the launch (grid is a multiple of it)
  // device copies
  PCmd* d_cmds = nullptr;
  CUtensorMap* d_maps = nullptr;
  unsigned* d_barrier = nullptr;
  unsigned long long* d_trace = nullptr;
  bool uploaded = false;

  int add_map_pair(const CUtensorMap& hi, const CUtensorMap& lo) {
    maps.push_back(hi); maps.push_back(lo);
    return (int)maps.size() - 2;
  }
  template <typename A, typename B, typename C> void add3(int type, const A& a, const B& b, const C& c3) {
    static_assert(sizeof(A) + sizeof(B) + sizeof(C) <= sizeof(PCmd::payload), "fused command payload too large");
    static_assert(sizeof(A) % 8 == 0 && sizeof(B) % 8 == 0, "payload parts must keep 8-byte alignment");
    PCmd c{};
    c.type = type;
    memcpy(c.payload, &a, sizeof(A));
    memcpy(c.payload + sizeof(A), &b, sizeof(B));
    memcpy(c.payload + sizeof(A) + sizeof(B), &c3, sizeof(C));
    cmds.push_back(c);
  }
  template <typename A, typename B> void add2(int type, const A& a, const B& b) {   // fused command: two payloads back to back
    static_assert(sizeof(A) + sizeof(B) <= sizeof(PCmd::payload), "fused command payload too large");
    PCmd c{};
    c.type = type;
    memcpy(c.payload, &a, sizeof(A));
    memcpy(c.payload + sizeof(A), &b, sizeof(B));
    cmds.push_back(c);
  }
  template <typename T> void add(int type, const T& p) {
    static_assert(sizeof(T) <= sizeof(PCmd::payload), "command payload too large");
    PCmd c{};
    c.type = type;
    memcpy(c.payload, &p, sizeof(T));
    cmds.push_back(c);
  }
};

// plan of one swap-AB GEMM inside a program: split factor such that m_tiles * splits <= max_ctas
struct PGemmPlan { int m_tiles, splits, kb_per, num_kb; long long part_stride; };
PGemmPlan persist_plan_gemm(int M, int N, int K, int max_ctas, long long ws_floats);
// fused GEMM -> cell: exactly `cluster` splits (one per CTA of the cluster that owns the M tile)
PGemmPlan persist_plan_gemm_fused(int M, int N, int K, int cluster);

void persist_upload(Ctx& ctx, PersistProgram& prog);     // allocates + copies (once)
void persist_launch(Ctx& ctx, PersistProgram& prog);     // cooperative launch on ctx.st
void persist_free(PersistProgram& prog);
cudaError_t persist_probe(cudaStream_t st, int grid, int cluster, bool coop);   // launch-mode probe (empty program)
int persist_max_ctas(int bn);                            // co-resident CTAs of the executor on this device
int persist_max_cluster_ctas(int bn, int cluster);       // the same when launched with thread-block clusters

}  // namespace aocr
