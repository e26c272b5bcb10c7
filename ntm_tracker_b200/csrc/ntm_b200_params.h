// ntm_b200_params.h -- launch parameters shared by the host side (ntm_b200.cu) and the two
// compilations of the persistent kernel (512 threads x 1 CTA/SM, 256 threads x 2 CTAs/SM).
#pragma once
#include <cuda_runtime.h>

#include <mutex>

#include "ntm_b200.h"

namespace ntm_b200 {

constexpr int RB = 4;            // memory rows per warp-group in pass 1
constexpr int TBMAX = 16;        // max sequences per warp tile in the SIMT skinny GEMMs
constexpr int MAXL = NTM_B200_MAX_LAYERS;
constexpr int SMAX = 2 * NTM_B200_MAX_SHIFT_RANGE + 1;
constexpr int B200_SMS = 148;
constexpr int B200_SMEM_OPTIN = 232448;   // 227 KiB per CTA
constexpr int B200_SMEM_SM = 233472;      // 228 KiB per SM (1 KiB reserved per resident CTA)
constexpr int PROF_SLOTS = 16;

struct GemmPlan {
  int K, NC, NCs, ldw, lda;            // NCs: row stride of the partial slabs
  int KS, KW, JW, NBT, TB, Gpad, njg, units;
  int tc;      // 1: tcgen05 path (128-column weight tiles resident in TMEM), 0: SIMT path
  int tcol;    // tc: first TMEM column of this GEMM's weight tile (KW/2 "hi" columns, then KW/2 "lo")
};

// One launch = `nteams` decoupled teams of `team_ctas` CTAs.  A team owns its own resident
// sequences, workspace slices and barrier counter and runs the four-phase timestep loop on its
// own; with two teams co-resident on every SM one team's barrier / latency bubbles are filled by
// the other's work.
struct KParams {
  int D, O, N, M, M4, MC, S, C, L, H, P, PO, PO4, write_first, shift0;
  int B, T, CS, NR, G, Npad;           // G: resident sequences (clusters) PER TEAM
  int nteams, team_ctas, tmem_cols;
  GemmPlan gA[MAXL];
  GemmPlan gC;
  const float* wA[MAXL];
  const float* bA[MAXL];
  const float* wC;
  const float* bC;
  const float* xw;
  const float *sM, *sw, *sread, *sctrl;
  long long ssM, ssw, ssread, ssctrl;
  float *dM, *dw, *dread, *dctrl;
  long long dsM, dsw, dsread, dsctrl;
  float* logits;
  float* outputs;
  float* dbg;
  long long dbgStride;
  // training-mode history (all optional, null = not recorded); see ntm_b200_history in ntm_b200.h
  float *hM, *hW, *hP, *hZ, *hC, *hH, *hRead, *hSim, *hCn;
  float* act[MAXL];                    // team 0's slice; team t at + t * act_ts[l]
  long long act_ts[MAXL];
  int actK[MAXL];
  float* cst;
  long long cst_ts;
  float* partA;
  long long partA_ts;
  float* partC;
  long long partC_ts;
  unsigned* ctr;                       // team t's barrier counter at ctr + 32 * t (128 B apart)
  int* err;
  long long* prof;                     // [grid CTAs][PROF_SLOTS] phase-cycle counters, or null
  // shared-memory carve-up, offsets in floats
  int oMs, oW0, oW1, oCn, oScr;
  int oSim, oSl, oWg, oK, oE, oA, oSm;
  int oTc;       // 4 floats: mbarrier (8 B) + TMEM base address (4 B); then 32 floats of phase counters
  int use_tc;    // any GEMM on the tensor path -> allocate TMEM
};

// Entry points each kernel compilation exports (namespaces k512 / k256).
struct KernelVariant {
  int threads;          // CTA size
  int ctas_per_sm;      // co-resident CTAs per SM the variant is built for
  bool cooperative_ok;  // the driver accepts a cooperative launch of a full grid of this build
  cudaError_t (*set_smem)(int R, int W, int smem_bytes);
  cudaError_t (*max_clusters)(int R, int W, int cluster_size, int grid_ctas, int smem_bytes, int* out);
  cudaError_t (*launch)(int R, int W, const KParams& p, int grid_ctas, int cluster_size, int smem_bytes,
                        bool cooperative, cudaStream_t stream);
};
// cudaFuncSetAttribute / occupancy results are PER DEVICE: launchers remember what they configured per
// (kernel instantiation, device) in a function-local `static int cfg[MAX_DEVICES]`, under this mutex.
constexpr int MAX_DEVICES = 64;
std::mutex& config_mutex();
inline int current_device_slot() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess) { cudaGetLastError(); d = 0; }
  return (d < 0 || d >= MAX_DEVICES) ? 0 : d;
}

// Environment switches (experiments / debugging), read ONCE per C-ABI call by read_env() and handed
// down, never inside the per-timestep loops.
struct EnvSwitches {
  int mode;                  // NTM_B200_MODE: -1 automatic, 0 resident, 1 stream
  long long stream_min_batch;// NTM_B200_STREAM_MIN_BATCH (-1 = default threshold)
  bool disable_tc;           // NTM_B200_DISABLE_TC
  bool dual_team;            // NTM_B200_DUAL_TEAM
  bool no_coop;              // NTM_B200_NO_COOP (needed under Nsight Compute kernel replay)
  bool no_tma_ring;          // NTM_B200_NO_TMA_RING
  bool old_gemm;             // NTM_B200_OLD_GEMM
  int mem_ctas_per_sm;       // NTM_B200_MEM_CTAS_PER_SM (0 = occupancy)
  int mem_grid;              // NTM_B200_MEM_GRID: cap on the persistent memory kernel's CTAs (0 = none; experiments)
  int exp;                   // NTM_B200_EXP: bit mask of experiment switches (development only; 0 in production):
                             //    1  gemm_ws: two bulk-copy issuing lanes per record instead of four
                             //    2  memory kernel: no pass-1 stages retained in the ring for pass 2
                             //    4  scalar gate kernel (lstm_stream_kernel) instead of the vectorised one
                             //    8  gemm_ws in-kernel phase timers, printed to stderr per call (synchronises)
                             //   16  no programmatic dependent launch (forward chain and reverse-time loop)
                             //   32  input projection on the tile kernel (ntm_b200_xproj_tc.cuh) in streaming mode too
                             //   64 / 128  resident mode: helper clusters never / always
                             //  256  memory kernel: evict-last instead of the normal L2 policy on re-read pass-1 stages
                             // 2048  memory kernel: no compact shared-memory plan for large N (one CTA per SM at C4)
                             // 4096  memory kernel pass 2: one stage pair per CTA barrier instead of two
};
EnvSwitches read_env();

// Programmatic dependent launch (the per-timestep kernel chain of the streaming mode): a kernel launched with
// the attribute may become resident while its predecessor in the stream is still running; everything it does
// before pdl_wait() (barrier init, TMEM allocation, weight staging -- nothing the predecessor produces) overlaps
// the predecessor's tail, and pdl_trigger() lets ITS successor do the same.  Both are no-ops in a plain launch.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_chain(void (*kern)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t stream,
                                bool pdl, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#endif

bool profiling_enabled();   // ntm_b200_set_profiling state (per-kernel CUDA events on the launching stream)
void count_launch();   // bumps the library-wide kernel-launch counter (ntm_b200_launch_count)
namespace k512 { const KernelVariant& variant(); }
namespace k256 { const KernelVariant& variant(); }

}  // namespace ntm_b200
