// ntm_b200.cu -- B200 (sm_100a) implementation of the NTM-cell hot path behind
// the C ABI of include/ntm_b200.h.
//
// What it replaces (paths relative to the reference root):
//   NTMCell.__call__ .......... ntm_cell.py:53-253
//   batched_smooth_cosine_similarity / batched_circular_convolution ... ops.py:135-242
//   LoopNTMTracker.__call__ ... ntm_tracker_new.py:13-64
//
// This file is the host side (planner, workspace layout, C ABI); the persistent kernel lives in
// ntm_b200_seq_kernel.cuh and is compiled twice (ntm_b200_k512.cu / ntm_b200_k256.cu).
//
// Design: ONE persistent kernel per call.  A thread-block
// cluster of CS CTAs owns one sequence; its N x M memory is split by rows over
// the cluster's shared memory and stays resident for all T steps, as do the
// head weightings and the per-column norms.  Each timestep has four phases
// separated by a device-wide barrier:
//   A  recurrent controller projection [read | h] @ W_rh, batched over all
//      resident sequences and split (columns x K-slices) over the whole grid;
//   B  split-K reduction + LSTM gates;
//   C  head-parameter / output projection  h @ [W_addr | W_out], same scheme;
//   D  per-cluster fused addressing: activations, column-normalised cosine
//      similarity (pass 1 over M), softmax, gate, circular shift, sharpening
//      (replicated per CTA after a DSMEM all-gather of the similarities), then
//      erase/add write + weighted read + next-step column norms (pass 2 over M),
//      cluster-reduced through DSMEM.
// The x_t @ W_x part of the controller projection has no state dependence and
// is hoisted out of the recurrence into one GEMM over all B*T rows.
//
// No CPU fallback, no dispatch: every entry point fails loudly without an
// sm_100 device.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "ntm_b200.h"
#include "ntm_b200_params.h"
#include "ntm_b200_xproj.cuh"
#include "ntm_b200_xproj_tc.cuh"
#include "ntm_b200_stream.h"

using namespace ntm_b200;

namespace {
std::atomic<long long> g_launches{0};
}
void ntm_b200::count_launch() { g_launches++; }
std::mutex& ntm_b200::config_mutex() {
  static std::mutex m;
  return m;
}
ntm_b200::EnvSwitches ntm_b200::read_env() {
  EnvSwitches e{};
  const char* m = getenv("NTM_B200_MODE");
  e.mode = (m == nullptr) ? -1 : (m[0] == 'r' ? 0 : (m[0] == 's' ? 1 : -1));
  const char* t = getenv("NTM_B200_STREAM_MIN_BATCH");
  e.stream_min_batch = t != nullptr ? atoll(t) : -1;
  e.disable_tc = getenv("NTM_B200_DISABLE_TC") != nullptr;
  e.dual_team = getenv("NTM_B200_DUAL_TEAM") != nullptr;
  e.no_coop = getenv("NTM_B200_NO_COOP") != nullptr;
  e.no_tma_ring = getenv("NTM_B200_NO_TMA_RING") != nullptr;
  e.old_gemm = getenv("NTM_B200_OLD_GEMM") != nullptr;
  const char* c = getenv("NTM_B200_MEM_CTAS_PER_SM");
  e.mem_ctas_per_sm = c != nullptr ? atoi(c) : 0;
  const char* gcap = getenv("NTM_B200_MEM_GRID");
  e.mem_grid = gcap != nullptr ? atoi(gcap) : 0;
  const char* x = getenv("NTM_B200_EXP");
  e.exp = x != nullptr ? atoi(x) : 0;
  return e;
}

namespace {
thread_local char g_cuda_err[256] = "";
}
int ntm_b200::set_cuda_error_ext(cudaError_t e, const char* where) {
  snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", where, cudaGetErrorString(e));
  return NTM_B200_ERR_CUDA;
}
int ntm_b200::gemm_tc(const float* x, int ldx, const float* w, int ldw, const float* bias, float* out, int ldo,
                      long long slab, long long rows, int K, int ncols, int kslices, int nsm, cudaStream_t stream) {
  return launch_gemm_tc(x, ldx, w, ldw, bias, out, ldo, slab, rows, K, ncols, kslices, nsm, stream);
}
int ntm_b200::gemm_tc_slices(int K) { return gemm_tc_kslices(K); }

namespace {

// ------------------------------------------------------------------ packing --
__global__ void pack_ao_kernel(const float* __restrict__ aw, const float* __restrict__ ab,
                               const float* __restrict__ ow, const float* __restrict__ ob,
                               float* wC, float* bC, int C, int P, int O, int PO4) {
  const int total = (C + 1) * PO4;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int r = i / PO4, q = i - r * PO4;
    float v = 0.0f;
    if (r < C) {
      if (q < P) v = aw[(size_t)r * P + q];
      else if (q < P + O) v = ow[(size_t)r * O + (q - P)];
      wC[i] = v;
    } else {
      if (q < P) v = ab[q];
      else if (q < P + O) v = ob[q - P];
      bC[q] = v;
    }
  }
}

// --------------------------------------------------------------- host side --
std::atomic<int> g_profiling{0};
}  // namespace
bool ntm_b200::profiling_enabled() { return g_profiling.load() != 0; }
namespace {
thread_local cudaEvent_t g_ev[3] = {nullptr, nullptr, nullptr};
thread_local bool g_ev_valid = false;
thread_local int g_last_info[16] = {0};
int set_cuda_error(cudaError_t e, const char* where) { return set_cuda_error_ext(e, where); }

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int round_up(int a, int b) { return ceil_div(a, b) * b; }
inline long long align_up_ll(long long a, long long b) { return (a + b - 1) / b * b; }

struct HostPlan {
  int H, S, P, PO, PO4, M4, MC, Npad;
  int CS, NR;
  int variant;          // 0: k512 (512 threads, 1 CTA/SM), 1: k256 (256 threads, 2 CTAs/SM)
  int nwarp, ctas_per_sm, tmem_cols;
  int max_teams;        // 2 when two CTAs share an SM (two decoupled teams), else 1
  int Gteam_max;        // resident sequences per team (upper bound, before the occupancy query)
  bool resident_ok;     // false: the per-sequence state does not fit an 8-CTA cluster -- streaming mode only
  int smem_floats;
  int oMs, oW0, oW1, oCn, oScr, oSim, oSl, oWg, oK, oE, oA, oSm, oTc;
  int scr_floats;
  int actK[MAXL];
  long long packed_bytes, debug_floats;
};

// The 256-thread two-team build (ntm_b200_k256.cu) is a measured negative result: it is compiled only
// with -DNTM_B200_WITH_K256 (NTM_B200_WITH_K256=1 for __graft_entry__.build()); without it NTM_B200_DUAL_TEAM is ignored.
#ifdef NTM_B200_WITH_K256
const KernelVariant& variant_of(const HostPlan& hp) { return hp.variant ? k256::variant() : k512::variant(); }
#else
const KernelVariant& variant_of(const HostPlan&) { return k512::variant(); }
#endif

int validate_shape(const ntm_b200_shape* s) {
  if (!s) return NTM_B200_ERR_NULL_POINTER;
  if (s->input_dim < 1 || s->output_dim < 1 || s->mem_size < 1 || s->mem_dim < 1 ||
      s->controller_hidden_size < 1 || s->controller_num_layers < 1 ||
      s->controller_num_layers > MAXL)
    return NTM_B200_ERR_BAD_SHAPE;
  if (s->shift_range < 0 || s->shift_range > NTM_B200_MAX_SHIFT_RANGE) return NTM_B200_ERR_BAD_SHIFT;
  // circular_shift asserts 0 <= splitting point < N for every tap (ops.py:229-231)
  const int S = 2 * s->shift_range + 1;
  const int start = -((S + 1) / 2);   // floor(-S / 2) for odd S
  if (-start >= s->mem_size || (S + start - 1) >= s->mem_size) return NTM_B200_ERR_BAD_SHIFT;
  if (s->read_head_size < 1 || s->read_head_size > NTM_B200_MAX_READ_HEADS ||
      s->write_head_size < 1 || s->write_head_size > NTM_B200_MAX_WRITE_HEADS)
    return NTM_B200_ERR_UNSUPPORTED_HEADS;
  return NTM_B200_OK;
}

// Shared-memory carve-up for one CTA of a CS-cluster with `nwarp` warps, inside `budget` bytes.
// Returns false if it does not fit.
bool layout_for(const ntm_b200_shape* s, int CS, int nwarp, int budget, HostPlan* hp) {
  const int H = s->read_head_size + s->write_head_size, W = s->write_head_size;
  hp->H = H;
  hp->S = 2 * s->shift_range + 1;
  hp->M4 = round_up(s->mem_dim, 4);
  hp->MC = hp->M4 / 4;
  hp->Npad = round_up(s->mem_size, 4);
  hp->P = H * s->mem_dim + 3 * H + hp->S * H + 2 * s->mem_dim * W;
  hp->PO = hp->P + s->output_dim;
  hp->PO4 = round_up(hp->PO, 4);
  hp->CS = CS;
  hp->NR = ceil_div(s->mem_size, CS);
  int o = 0;
  auto take = [&](int n) { int r = o; o += round_up(n, 4); return r; };
  hp->oMs = take(hp->NR * hp->M4);
  hp->oW0 = take(H * hp->Npad);
  hp->oW1 = take(H * hp->Npad);
  hp->oCn = take(hp->M4);
  hp->oTc = take(4 + 2 * PROF_SLOTS);   // mbarrier + TMEM base, then int64 phase-cycle accumulators
  hp->oScr = o;
  // phase-D temporaries inside the scratch union
  int d = o;
  auto taked = [&](int n) { int r = d; d += round_up(n, 4); return r; };
  hp->oSim = taked(H * hp->Npad);                      // private full-length similarities
  hp->oSl = taked(H * hp->NR);                         // this CTA's slice, pulled by the peers over DSMEM
  hp->oWg = taked(std::max(H * hp->Npad, hp->PO4));    // also holds the raw head-parameter vector
  hp->oK = taked(H * hp->M4);      // kS, eS, aS contiguous; kS doubles as the DSMEM exchange buffer
  hp->oE = taked(W * hp->M4);
  hp->oA = taked(W * hp->M4);
  hp->oSm = taked(4 * H + H * SMAX + nwarp * H + 3 * nwarp);
  const int dfl = d - o;
  const int min_stage = 16 * 1024 / 4, max_stage = 44 * 1024 / 4;
  int scr = std::max(dfl, min_stage);
  if (4ll * (o + scr) > budget) return false;
  scr = std::max(scr, std::min(max_stage, budget / 4 - o));   // staging gets what is left, up to 44 KiB
  hp->scr_floats = scr;
  hp->smem_floats = o + scr;
  return true;
}

int make_host_plan(const ntm_b200_shape* s, int nsm, int smem_optin, const EnvSwitches& env, HostPlan* hp) {
  int st = validate_shape(s);
  if (st) return st;
  // Experimental (NTM_B200_DUAL_TEAM=1, tensor path only): 256-thread CTAs, two per SM, as two
  // decoupled teams.  Measured on B200 it does NOT pay off -- the step is a chain of fixed latencies,
  // shared memory caps the resident sequences at the same ~74 either way, so two half-size teams
  // each run the same ~35 us step (C2: 35.0 us vs 32.3 us single team) -- hence off by default.
  const int half_budget = std::min(smem_optin, B200_SMEM_SM / 2 - 1024 - 2048);   // slack for allocation granularity
  bool ok = false;
  // (tensor path only: with the SIMT GEMMs the two-team build is known to time out at 35 sequences per team)
#ifdef NTM_B200_WITH_K256
  const bool dual_team = env.dual_team;
#else
  const bool dual_team = false;
#endif
  if (dual_team && !env.disable_tc) {
    for (int CS = 1; CS <= 8 && !ok; CS *= 2) {
      if (layout_for(s, CS, 8, half_budget, hp)) {
        ok = true;
        hp->variant = 1; hp->nwarp = 8; hp->ctas_per_sm = 2; hp->tmem_cols = 256; hp->max_teams = 2;
      }
    }
  }
  for (int CS = 1; CS <= 8 && !ok; CS *= 2) {
    if (layout_for(s, CS, 16, smem_optin, hp)) {
      ok = true;
      hp->variant = 0; hp->nwarp = 16; hp->ctas_per_sm = 1; hp->tmem_cols = 512; hp->max_teams = 1;
    }
  }
  hp->resident_ok = ok;
  if (!ok) {
    // The persistent kernel cannot hold a sequence (N*M*4 beyond what 8 CTAs' shared memory takes).  The streaming
    // kernels do not need that: such shapes run in streaming mode whatever the batch size (no debug taps there).
    if (!stream_supported(s, nsm)) return NTM_B200_ERR_TOO_LARGE;
    layout_for(s, 8, 16, 1 << 30, hp);      // shape-derived fields only (P, PO4, M4, ...); the carve-up is unused
    hp->variant = 0; hp->nwarp = 16; hp->ctas_per_sm = 1; hp->tmem_cols = 512; hp->max_teams = 1;
  }
  hp->Gteam_max = std::max(1, (nsm * hp->ctas_per_sm / hp->CS) / hp->max_teams);
  const int C = s->controller_hidden_size;
  for (int l = 0; l < s->controller_num_layers; ++l)
    hp->actK[l] = (l == 0) ? s->read_head_size * s->mem_dim + C : 2 * C;
  hp->packed_bytes = 4ll * (long long)(C + 1) * hp->PO4;
  hp->debug_floats = (long long)hp->P + 5ll * hp->H * s->mem_size;
  return NTM_B200_OK;
}

GemmPlan plan_gemm(int K, int NC, int NCs, int ldw, int lda, int G, int ncta, int nwarp, int scr_bytes) {
  GemmPlan g{};
  g.K = K; g.NC = NC; g.NCs = NCs; g.ldw = ldw; g.lda = lda;
  const int G4 = round_up(std::max(G, 1), 4);
  int best_nbt = nwarp, best_pad = 1 << 30;
  for (int nbt = ceil_div(G4, TBMAX); nbt <= nwarp; ++nbt) {
    const int tb = round_up(ceil_div(G4, nbt), 4);
    if (tb > TBMAX) continue;
    const int pad = nbt * tb;
    if (pad < best_pad) { best_pad = pad; best_nbt = nbt; }
  }
  g.NBT = best_nbt;
  g.TB = round_up(ceil_div(G4, g.NBT), 4);
  g.Gpad = g.NBT * g.TB;
  g.JW = std::max(1, nwarp / g.NBT);
  const int nj64 = ceil_div(NC, 64);
  g.JW = std::min(g.JW, nj64);
  g.njg = ceil_div(nj64, g.JW);
  int KS = std::max(1, std::min(ncta / std::max(1, g.njg), ceil_div(K, 16)));
  int KW = round_up(ceil_div(K, KS), 4);
  const int kw_cap = std::max(4, (scr_bytes / 4 / g.Gpad) / 4 * 4);
  KW = std::min(KW, kw_cap);
  g.KW = KW;
  g.KS = ceil_div(K, KW);
  g.units = g.KS * g.njg;
  return g;
}

// Tensor-path plan: one unit (128-column tile x K-slice) per CTA, weights resident in TMEM.
bool plan_gemm_tc(int K, int NC, int NCs, int ldw, int lda, int G, int ncta, int tcol, int scr_bytes,
                  GemmPlan* out) {
  GemmPlan g{};
  g.K = K; g.NC = NC; g.NCs = NCs; g.ldw = ldw; g.lda = lda;
  g.tc = 1;
  g.tcol = tcol;
  g.Gpad = round_up(std::max(G, 1), 16);
  if (g.Gpad > 256) return false;
  const int tiles = ceil_div(NC, 128);
  if (tiles > ncta) return false;
  int KS = std::max(1, std::min(ncta / tiles, ceil_div(K, 16)));
  const int KW = round_up(ceil_div(K, KS), 16);
  KS = ceil_div(K, KW);
  const int katoms = ceil_div(KW, 64);
  if (2 * g.Gpad * katoms * 128 + 1024 > scr_bytes) return false;
  g.KW = KW; g.KS = KS; g.njg = tiles; g.units = KS * tiles;
  *out = g;
  return true;
}

// All GEMM plans of one team.  The tensor path is used when every GEMM's weight tile fits the
// CTA's TMEM columns next to the accumulator (else the SIMT path, e.g. for small grids).
void choose_plans(const ntm_b200_shape* s, const HostPlan& hp, int G, int ncta, bool allow_tc,
                  GemmPlan* gA, GemmPlan* gC, int* use_tc) {   // allow_tc already folds in NTM_B200_DISABLE_TC
  const int C = s->controller_hidden_size, L = s->controller_num_layers;
  const int scr_bytes = 4 * hp.scr_floats;
  bool tc = allow_tc;
  if (tc) {
    int col = round_up(round_up(std::max(G, 1), 16), 32);     // accumulator columns first
    for (int l = 0; l < L && tc; ++l) {
      tc = plan_gemm_tc(hp.actK[l], 4 * C, 4 * C, 4 * C, hp.actK[l], G, ncta, col, scr_bytes, &gA[l]);
      if (tc) col += gA[l].KW;
    }
    if (tc) tc = plan_gemm_tc(C, hp.PO, hp.PO4, hp.PO4, hp.actK[L - 1], G, ncta, col, scr_bytes, gC);
    if (tc) col += gC->KW;
    if (col > hp.tmem_cols) tc = false;
  }
  if (!tc) {
    for (int l = 0; l < L; ++l)
      gA[l] = plan_gemm(hp.actK[l], 4 * C, 4 * C, 4 * C, hp.actK[l], G, ncta, hp.nwarp, scr_bytes);
    *gC = plan_gemm(C, hp.PO, hp.PO4, hp.PO4, hp.actK[L - 1], G, ncta, hp.nwarp, scr_bytes);
  }
  *use_tc = tc ? 1 : 0;
}

struct Workspace {
  long long off_ctr, off_err, off_prof, off_act[MAXL], off_cst, off_partA, off_partC, off_xw, total;
  long long act_ts[MAXL], cst_ts, partA_ts, partC_ts;   // per-team strides in floats
};

// Workspace sized for the planner's upper bounds (max_teams teams of Gteam_max resident sequences),
// so it does not depend on what the occupancy query returns later.
void layout_workspace(const ntm_b200_shape* s, const HostPlan& hp, long long B, long long T,
                      Workspace* ws) {
  const int C = s->controller_hidden_size, L = s->controller_num_layers;
  const int Gm = hp.Gteam_max, NTm = hp.max_teams;
  if (!hp.resident_ok) {      // streaming-only shape: the persistent kernel's workspace does not exist
    *ws = Workspace{};
    return;
  }
  long long o = 0;
  auto take = [&](long long bytes) { long long r = o; o = align_up_ll(o + bytes, 256); return r; };
  ws->off_ctr = take(256);        // one counter per team, 128 B apart
  ws->off_err = take(256);
  ws->off_prof = take(8ll * PROF_SLOTS * 1024);   // directly after ctr/err: zeroed by the same memset
  for (int l = 0; l < L; ++l) {
    ws->act_ts[l] = align_up_ll((long long)Gm * hp.actK[l], 64);
    ws->off_act[l] = take(4ll * ws->act_ts[l] * NTm);
  }
  ws->cst_ts = align_up_ll((long long)Gm * L * C, 64);
  ws->off_cst = take(4ll * ws->cst_ts * NTm);
  // partial-slab sizes: maximum over every resident-sequence count the launch may end up with,
  // on either GEMM path
  long long pa = 0, pc = 0;
  for (int G = 1; G <= Gm; ++G) {
    const int ncta = G * hp.CS;
    for (int variant = 0; variant < 2; ++variant) {
      GemmPlan gA[MAXL], gC;
      int use_tc = 0;
      choose_plans(s, hp, G, ncta, variant == 1, gA, &gC, &use_tc);   // both paths: sized for either
      for (int l = 0; l < L; ++l) pa = std::max(pa, (long long)gA[l].KS * gA[l].Gpad * gA[l].NCs);
      pc = std::max(pc, (long long)gC.KS * gC.Gpad * gC.NCs);
    }
  }
  ws->partA_ts = align_up_ll(pa, 64);
  ws->partC_ts = align_up_ll(pc, 64);
  ws->off_partA = take(4ll * ws->partA_ts * NTm);
  ws->off_partC = take(4ll * ws->partC_ts * NTm);
  ws->off_xw = take(4ll * B * T * 4 * C);
  ws->total = o;
}

struct DeviceInfo { int ok; int nsm; int smem_optin; int cc_major; };

DeviceInfo device_info() {
  DeviceInfo d{0, B200_SMS, B200_SMEM_OPTIN, 0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return d; }
  int major = 0, nsm = 0, smem = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) {
    cudaGetLastError();
    return d;
  }
  d.ok = (major == 10);
  d.nsm = nsm; d.smem_optin = smem; d.cc_major = major;
  return d;
}

int check_state(const ntm_b200_state* st) {
  if (!st || !st->M || !st->w || !st->read || !st->controller_state) return NTM_B200_ERR_NULL_POINTER;
  return NTM_B200_OK;
}

// Execution mode for a call: 0 = persistent shared-memory-resident kernel, 1 = streaming (lockstep over
// the whole shard, memory streamed from HBM).  NTM_B200_MODE=resident|stream overrides the choice.
int choose_mode(const ntm_b200_shape* s, const HostPlan& hp, long long B, bool debug_taps, int nsm, const EnvSwitches& env) {
  if (!hp.resident_ok) return 1;      // (make_host_plan made sure the streaming kernels cover the shape)
  if (debug_taps || !stream_supported(s, nsm)) return 0;
  if (env.mode >= 0) return env.mode;
  // resident: ceil(B / G) waves of ~30 us steps; streaming pays ~4 launches + GEMM weight loads per step
  // but its step time grows only with the HBM traffic.  Measured crossover at tracker shapes, T = 16
  // (profiles/r2_ab_probes.txt): with the round-2 streaming path (programmatic dependent launch, faster GEMM
  // issue / prologue) streaming wins as soon as the batch needs a second wave -- B = 74: 1.72 vs 1.19 M seq-steps/s,
  // 111: 1.45 vs 1.79, 148: 1.91 vs 2.30, 222: 1.96 vs 3.31 (round 1: three waves).  Shapes that run the fallback
  // streaming kernels keep the older threshold.
  long long thr = (stream_ws_path(s) ? 1ll : 3ll) * hp.Gteam_max * hp.max_teams;
  if (env.stream_min_batch >= 0) thr = env.stream_min_batch;
  return B > thr ? 1 : 0;
}

}  // namespace

// ------------------------------------------------------------------- C ABI --
extern "C" {

int32_t ntm_b200_abi_version(void) { return NTM_B200_ABI_VERSION; }

const char* ntm_b200_status_string(int32_t status) {
  switch (status) {
    case NTM_B200_OK: return "ok";
    case NTM_B200_ERR_BAD_SHAPE: return "bad shape (non-positive dimension or too many controller layers)";
    case NTM_B200_ERR_BAD_SHIFT: return "shift_range out of range for mem_size";
    case NTM_B200_ERR_NULL_POINTER: return "null pointer argument";
    case NTM_B200_ERR_UNSUPPORTED_HEADS: return "head count outside 1..4 read / 1..3 write";
    case NTM_B200_ERR_TOO_LARGE: return "per-sequence state does not fit an 8-CTA cluster's shared memory";
    case NTM_B200_ERR_WORKSPACE: return "workspace or packed-weight buffer too small";
    case NTM_B200_ERR_NO_DEVICE: return "no sm_100 CUDA device (there is no CPU fallback)";
    case NTM_B200_ERR_CUDA: return "CUDA runtime error";
    case NTM_B200_ERR_DEVICE_TIMEOUT: return "device-side grid barrier timed out";
    default: return "unknown status";
  }
}

const char* ntm_b200_last_cuda_error(void) { return g_cuda_err; }

int64_t ntm_b200_launch_count(void) { return g_launches.load(); }

int32_t ntm_b200_query(const ntm_b200_shape* shape, int64_t batch, int64_t steps,
                       ntm_b200_plan* plan_out) {
  if (!shape || !plan_out) return NTM_B200_ERR_NULL_POINTER;
  if (batch < 1 || steps < 1) return NTM_B200_ERR_BAD_SHAPE;
  DeviceInfo di = device_info();
  const EnvSwitches env = read_env();
  HostPlan hp{};
  int st = make_host_plan(shape, di.nsm, di.smem_optin, env, &hp);
  if (st) return st;
  Workspace ws{};
  layout_workspace(shape, hp, batch, steps, &ws);
  const int teams = (batch >= 2) ? hp.max_teams : 1;
  if (hp.resident_ok) {
    plan_out->cluster_size = hp.CS;
    plan_out->rows_per_cta = hp.NR;
    plan_out->sequences_resident = (int32_t)std::min<long long>((long long)hp.Gteam_max * teams, batch);
    plan_out->smem_bytes_per_cta = 4ll * hp.smem_floats;
  } else {      // streaming-only shape: there is no persistent-kernel geometry to report
    plan_out->cluster_size = 0;
    plan_out->rows_per_cta = 0;
    plan_out->sequences_resident = 0;
    plan_out->smem_bytes_per_cta = 0;
  }
  plan_out->threads_per_cta = 32 * hp.nwarp;
  plan_out->ctas_per_sm = hp.ctas_per_sm;
  plan_out->teams = teams;
  StreamWorkspace sws{};
  stream_layout(shape, batch, steps, &sws);
  plan_out->workspace_bytes = std::max(ws.total, sws.total + 1024);
  plan_out->packed_bytes = hp.packed_bytes;
  plan_out->debug_floats_per_sequence = hp.debug_floats;
  return NTM_B200_OK;
}

int32_t ntm_b200_query_mode(const ntm_b200_shape* shape, int64_t batch, int32_t* mode_out) {
  if (!shape || !mode_out) return NTM_B200_ERR_NULL_POINTER;
  if (batch < 1) return NTM_B200_ERR_BAD_SHAPE;
  DeviceInfo di = device_info();
  const EnvSwitches env = read_env();
  HostPlan hp{};
  int st = make_host_plan(shape, di.nsm, di.smem_optin, env, &hp);
  if (st) return st;
  *mode_out = choose_mode(shape, hp, batch, false, di.nsm, env);
  return NTM_B200_OK;
}

int32_t ntm_b200_pack_weights(const ntm_b200_shape* shape, const ntm_b200_weights* weights,
                              void* packed, int64_t packed_bytes, void* stream) {
  if (!shape || !weights || !packed) return NTM_B200_ERR_NULL_POINTER;
  DeviceInfo di = device_info();
  HostPlan hp{};
  int st = make_host_plan(shape, di.nsm, di.smem_optin, read_env(), &hp);
  if (st) return st;
  if (!di.ok) return NTM_B200_ERR_NO_DEVICE;
  if (packed_bytes < hp.packed_bytes) return NTM_B200_ERR_WORKSPACE;
  if (!weights->addr_w || !weights->addr_b || !weights->out_w || !weights->out_b)
    return NTM_B200_ERR_NULL_POINTER;
  const int C = shape->controller_hidden_size;
  float* wC = static_cast<float*>(packed);
  float* bC = wC + (size_t)C * hp.PO4;
  pack_ao_kernel<<<B200_SMS, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      weights->addr_w, weights->addr_b, weights->out_w, weights->out_b, wC, bC, C, hp.P,
      shape->output_dim, hp.PO4);
  g_launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "pack_ao_kernel");
  return NTM_B200_OK;
}

int32_t ntm_b200_forward_seq(const ntm_b200_shape* shape, const ntm_b200_weights* weights,
                             const void* packed, int64_t batch, int64_t steps,
                             const float* inputs, const ntm_b200_state* state_in,
                             const ntm_b200_state* state_out, float* logits, float* outputs,
                             float* debug_taps, void* workspace, int64_t workspace_bytes,
                             void* stream_v) {
  return ntm_b200_forward_seq_train(shape, weights, packed, batch, steps, inputs, state_in, state_out, logits,
                                    outputs, debug_taps, nullptr, workspace, workspace_bytes, stream_v);
}

// What the last streaming call of this thread left valid in its workspace (for ntm_b200_forward_seq_continue).
struct StreamResume { void* workspace; const float* M; long long batch; ntm_b200_shape shape; const void* packed; void* stream; };
static thread_local StreamResume g_resume = {nullptr, nullptr, 0, {}, nullptr, nullptr};
static thread_local StreamResume g_resume_prev = {nullptr, nullptr, 0, {}, nullptr, nullptr};
// Every call invalidates the record first (a resident-mode or debug call on the same workspace overwrites the
// streaming layout); only a streaming call without history re-arms it.
static void g_resume_reset() { g_resume_prev = g_resume; g_resume.workspace = nullptr; }

static int32_t forward_impl(const ntm_b200_shape* shape, const ntm_b200_weights* weights,
                            const void* packed, int64_t batch, int64_t steps,
                            const float* inputs, const ntm_b200_state* state_in,
                            const ntm_b200_state* state_out, float* logits, float* outputs,
                            float* debug_taps, const ntm_b200_history* history,
                            void* workspace, int64_t workspace_bytes, void* stream_v, bool want_cont, const ntm_b200::FeatureSource* fsrc = nullptr);

int32_t ntm_b200_forward_seq_train(const ntm_b200_shape* shape, const ntm_b200_weights* weights,
                                   const void* packed, int64_t batch, int64_t steps,
                                   const float* inputs, const ntm_b200_state* state_in,
                                   const ntm_b200_state* state_out, float* logits, float* outputs,
                                   float* debug_taps, const ntm_b200_history* history,
                                   void* workspace, int64_t workspace_bytes, void* stream_v) {
  return forward_impl(shape, weights, packed, batch, steps, inputs, state_in, state_out, logits, outputs, debug_taps,
                      history, workspace, workspace_bytes, stream_v, false);
}

int32_t ntm_b200_forward_seq_continue(const ntm_b200_shape* shape, const ntm_b200_weights* weights,
                                      const void* packed, int64_t batch, int64_t steps, const float* inputs,
                                      const ntm_b200_state* state, float* logits, float* outputs,
                                      void* workspace, int64_t workspace_bytes, void* stream_v) {
  return forward_impl(shape, weights, packed, batch, steps, inputs, state, state, logits, outputs, nullptr, nullptr,
                      workspace, workspace_bytes, stream_v, true);
}

static int32_t forward_impl(const ntm_b200_shape* shape, const ntm_b200_weights* weights,
                            const void* packed, int64_t batch, int64_t steps,
                            const float* inputs, const ntm_b200_state* state_in,
                            const ntm_b200_state* state_out, float* logits, float* outputs,
                            float* debug_taps, const ntm_b200_history* history,
                            void* workspace, int64_t workspace_bytes, void* stream_v, bool want_cont,
                            const ntm_b200::FeatureSource* fsrc) {
  if (!shape || !weights || !packed || (!inputs && !fsrc) || !logits || !workspace) return NTM_B200_ERR_NULL_POINTER;
  int st = check_state(state_in);
  if (st) return st;
  st = check_state(state_out);
  if (st) return st;
  if (batch < 1 || steps < 1 || batch > (1 << 24) || steps > (1 << 24)) return NTM_B200_ERR_BAD_SHAPE;
  DeviceInfo di = device_info();
  const EnvSwitches env = read_env();
  HostPlan hp{};
  st = make_host_plan(shape, di.nsm, di.smem_optin, env, &hp);
  if (st) return st;
  if (!di.ok) return NTM_B200_ERR_NO_DEVICE;
  g_resume_reset();
  const int L = shape->controller_num_layers, C = shape->controller_hidden_size;
  for (int l = 0; l < L; ++l)
    if (!weights->lstm_w[l] || !weights->lstm_b[l]) return NTM_B200_ERR_NULL_POINTER;
  Workspace ws{};
  layout_workspace(shape, hp, batch, steps, &ws);
  StreamWorkspace sws{};
  stream_layout(shape, batch, steps, &sws);
  const int mode = choose_mode(shape, hp, batch, debug_taps != nullptr, di.nsm, env);
  const long long ws_need = mode ? sws.total + 1024 : ws.total;
  if (workspace_bytes < ws_need) return NTM_B200_ERR_WORKSPACE;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  char* wsb = static_cast<char*>(workspace);
  cudaError_t e;
  g_last_info[13] = mode;

  if (!hp.resident_ok && debug_taps != nullptr) return NTM_B200_ERR_TOO_LARGE;   // taps come from the persistent kernel only
  const KernelVariant& kv = variant_of(hp);
  const int R = shape->read_head_size, W = shape->write_head_size;
  const int smem_bytes = hp.resident_ok ? 4 * hp.smem_floats : 0;
  const int nteams = (batch >= 2) ? hp.max_teams : 1;
  int max_clusters = nteams;
  if (mode == 0) {      // (the persistent kernel's launch geometry; the streaming mode configures its own kernels)
    e = kv.set_smem(R, W, smem_bytes);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(smem)");
    // how many clusters are co-resident, split evenly over the teams
    e = kv.max_clusters(R, W, hp.CS, hp.Gteam_max * hp.max_teams * hp.CS, smem_bytes, &max_clusters);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaOccupancyMaxActiveClusters");
    if (max_clusters < nteams) return NTM_B200_ERR_TOO_LARGE;
  }
  // Clusters per team.  Normally as many as there are sequences (up to what is co-resident).  When that would leave
  // each CTA more than 256 KiB of projection weights to stream per timestep -- batch 1, the serve path: one tracker,
  // 65 steps per frame, 10 MB of weights through two CTAs = 188 us per step -- every co-resident cluster is launched
  // instead: the ones without a sequence of their own ("helpers") take their share of the GEMM / gate phases and keep
  // their weight tiles in TMEM (33 us per step).  With enough sequences the extra CTAs only make the device-wide
  // barrier dearer (C1: 0.51 -> 0.59 ms per call with helpers, C2: 1.19 -> 1.24), hence the threshold.
  // NTM_B200_EXP bit 64: never, bit 128: always.
  const int Gfull = std::min(max_clusters / nteams, hp.Gteam_max);
  const int Gseq = (int)std::min<long long>(Gfull, (batch + nteams - 1) / nteams);
  long long wbytes = 4ll * C * hp.PO4;
  for (int l = 0; l < L; ++l) wbytes += 4ll * hp.actK[l] * 4 * C;
  const bool helpers = (env.exp & 128) || (!(env.exp & 64) && wbytes / ((long long)Gseq * hp.CS * nteams) > (256 << 10));
  const int G = helpers ? Gfull : Gseq;
  const int ncta = G * hp.CS;                                              // CTAs per team

  const bool prof = g_profiling.load() != 0;
  if (prof) {
    for (int i = 0; i < 3; ++i)
      if (!g_ev[i] && (e = cudaEventCreate(&g_ev[i])) != cudaSuccess) return set_cuda_error(e, "cudaEventCreate");
    cudaEventRecord(g_ev[0], stream);
  }
  // hoisted x-projection: xw[b,t,:] = x[b,t,:] @ W_lstm0[0:D,:] + b_lstm0
  // (streaming mode: the first 1 KiB of the workspace stays the error-flag block ntm_b200_finish reads)
  char* swsb = wsb + 1024;
  float* xw = mode ? reinterpret_cast<float*>(swsb + sws.off_xw) : reinterpret_cast<float*>(wsb + ws.off_xw);
  // tensor-core kernel (tcgen05, weight tile resident in TMEM + SMEM); fp32 SIMT kernel for shapes it
  // does not cover or when NTM_B200_DISABLE_TC is set
  st = -1;
  bool xw_partial = false;
  if (mode == 1 && !env.disable_tc) {
    // streaming mode: pack pass + the warp-specialised GEMM over the first 64*n input columns (the rest join in the
    // gate kernel); needs the same "continuation" decision stream_forward gets below
    const StreamResume& rs0 = g_resume_prev;
    const bool cont0 = want_cont && rs0.workspace == workspace && rs0.batch == batch && rs0.stream == stream_v &&
                       rs0.M == state_in->M && state_in->M == state_out->M && rs0.packed == packed &&
                       memcmp(&rs0.shape, shape, sizeof(*shape)) == 0;
    st = stream_xproj(shape, weights, batch, steps, inputs, xw, swsb, sws, di.nsm, stream, cont0, env, fsrc);
    if (st > 0) return st;
    xw_partial = (st == 0);
  }
  if (fsrc != nullptr && !xw_partial) {
    // feature-layout call on a path that wants serialised rows (resident mode, shapes the streaming projection does
    // not cover): materialise them behind the workspace proper (ntm_b200_features_workspace_bytes reserves the room)
    const long long xbytes = 4ll * batch * steps * shape->input_dim;
    const long long xoff = (ws_need + 255) / 256 * 256;
    if (workspace_bytes < xoff + xbytes) return NTM_B200_ERR_WORKSPACE;
    float* xmat = reinterpret_cast<float*>(wsb + xoff);
    const int sst = ntm_b200_serialize_tracker_inputs(fsrc->features, fsrc->target, xmat, batch, fsrc->L, fsrc->F, fsrc->Cch,
                                                      fsrc->delimiter_first, stream_v);
    if (sst) return sst;
    inputs = xmat;      // (st stays -1: the projection over the materialised rows runs below)
  }
  if (st < 0 && !env.disable_tc)
    st = ntm_b200::launch_xproj_tc(inputs, weights->lstm_w[0], weights->lstm_b[0], xw,
                                   (long long)batch * steps, shape->input_dim, 4 * C, di.nsm, stream);
  g_last_info[12] = (st == 0) ? 1 : 0;
  if (st < 0)
    st = ntm_b200::launch_xproj(inputs, weights->lstm_w[0], weights->lstm_b[0], xw,
                                (long long)batch * steps, shape->input_dim, 4 * C, stream);
  if (!xw_partial) g_launches++;
  if (st) return set_cuda_error(cudaGetLastError(), "xproj");

  if (mode == 1) {
    e = cudaMemsetAsync(wsb, 0, 1024, stream);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaMemsetAsync");
    if (prof) cudaEventRecord(g_ev[1], stream);
    const float* wCp = static_cast<const float*>(packed);
    // continuation: same workspace, batch, shape and packed weights as the last streaming call, and the state
    // it left (updated in place) -- otherwise this is an ordinary call
    const StreamResume& rs = g_resume_prev;   // what the call before this one left (g_resume itself is already reset)
    const bool cont = want_cont && rs.workspace == workspace && rs.batch == batch && rs.stream == stream_v &&
                      rs.M == state_in->M && state_in->M == state_out->M && rs.packed == packed &&
                      memcmp(&rs.shape, shape, sizeof(*shape)) == 0;
    st = stream_forward(shape, weights, wCp, wCp + (size_t)C * hp.PO4, batch, steps, xw, state_in, state_out,
                        logits, outputs, history, swsb, sws, di.nsm, stream, prof, cont, env, xw_partial);
    if (st) return st;
    if (history == nullptr) g_resume = StreamResume{workspace, state_out->M, (long long)batch, *shape, packed, stream_v};
    g_last_info[14] = cont ? 1 : 0;
    if (prof) {
      cudaEventRecord(g_ev[2], stream);
      g_ev_valid = true;
    }
    g_last_info[0] = 1; g_last_info[1] = (int)batch; g_last_info[2] = (int)batch; g_last_info[3] = 1;
    g_last_info[4] = sws.ksA[0]; g_last_info[6] = sws.ksC; g_last_info[8] = 1;
    g_last_info[9] = 256; g_last_info[10] = stream_mem_occupancy();
    return NTM_B200_OK;
  }

  KParams p{};
  p.D = shape->input_dim; p.O = shape->output_dim; p.N = shape->mem_size; p.M = shape->mem_dim;
  p.M4 = hp.M4; p.MC = hp.MC; p.S = hp.S; p.C = C; p.L = L; p.H = hp.H; p.P = hp.P; p.PO = hp.PO;
  p.PO4 = hp.PO4; p.write_first = shape->write_first ? 1 : 0;
  p.shift0 = -((hp.S + 1) / 2);   // Python-2 floor(-S/2), ops.py:204
  p.B = (int)batch; p.T = (int)steps; p.CS = hp.CS; p.NR = hp.NR; p.G = G; p.Npad = hp.Npad;
  p.nteams = nteams; p.team_ctas = ncta; p.tmem_cols = hp.tmem_cols;
  for (int l = 0; l < L; ++l) {
    p.actK[l] = hp.actK[l];
    p.act[l] = reinterpret_cast<float*>(wsb + ws.off_act[l]);
    p.act_ts[l] = ws.act_ts[l];
    p.wA[l] = weights->lstm_w[l] + (l == 0 ? (size_t)shape->input_dim * 4 * C : 0);
    p.bA[l] = weights->lstm_b[l];
  }
  choose_plans(shape, hp, G, ncta, !env.disable_tc, p.gA, &p.gC, &p.use_tc);
  g_last_info[0] = p.use_tc; g_last_info[1] = (int)std::min<long long>((long long)G * nteams, batch); g_last_info[2] = ncta * nteams; g_last_info[3] = hp.CS;
  g_last_info[4] = p.gA[0].KS; g_last_info[5] = p.gA[0].KW; g_last_info[6] = p.gC.KS; g_last_info[7] = p.gC.KW;
  g_last_info[8] = nteams; g_last_info[9] = kv.threads; g_last_info[10] = kv.ctas_per_sm; g_last_info[11] = smem_bytes;
  p.wC = static_cast<const float*>(packed);
  p.bC = p.wC + (size_t)C * hp.PO4;
  p.xw = xw;
  p.sM = state_in->M; p.sw = state_in->w; p.sread = state_in->read; p.sctrl = state_in->controller_state;
  p.ssM = state_in->stride_M; p.ssw = state_in->stride_w; p.ssread = state_in->stride_read;
  p.ssctrl = state_in->stride_controller_state;
  p.dM = state_out->M; p.dw = state_out->w; p.dread = state_out->read; p.dctrl = state_out->controller_state;
  p.dsM = state_out->stride_M; p.dsw = state_out->stride_w; p.dsread = state_out->stride_read;
  p.dsctrl = state_out->stride_controller_state;
  p.logits = logits; p.outputs = outputs; p.dbg = debug_taps; p.dbgStride = hp.debug_floats;
  if (history != nullptr) {
    p.hM = history->M_prev; p.hW = history->w_prev; p.hP = history->params; p.hZ = history->z;
    p.hC = history->c; p.hH = history->h; p.hRead = history->read; p.hSim = history->sim; p.hCn = history->cn;
  }
  p.cst = reinterpret_cast<float*>(wsb + ws.off_cst); p.cst_ts = ws.cst_ts;
  p.partA = reinterpret_cast<float*>(wsb + ws.off_partA); p.partA_ts = ws.partA_ts;
  p.partC = reinterpret_cast<float*>(wsb + ws.off_partC); p.partC_ts = ws.partC_ts;
  p.ctr = reinterpret_cast<unsigned*>(wsb + ws.off_ctr);
  p.err = reinterpret_cast<int*>(wsb + ws.off_err);
  p.prof = prof ? reinterpret_cast<long long*>(wsb + ws.off_prof) : nullptr;
  p.oMs = hp.oMs; p.oW0 = hp.oW0; p.oW1 = hp.oW1; p.oCn = hp.oCn;
  p.oScr = hp.oScr; p.oSim = hp.oSim; p.oSl = hp.oSl; p.oWg = hp.oWg; p.oK = hp.oK; p.oE = hp.oE; p.oA = hp.oA;
  p.oSm = hp.oSm; p.oTc = hp.oTc;

  e = cudaMemsetAsync(wsb + ws.off_ctr, 0, 512 + 8 * PROF_SLOTS * 1024, stream);   // barrier counter, error flag, phase counters
  if (e != cudaSuccess) return set_cuda_error(e, "cudaMemsetAsync");

  if (prof) cudaEventRecord(g_ev[1], stream);
  // cluster + cooperative (co-residency enforced by the driver).  NTM_B200_NO_COOP=1 drops the
  // cooperative attribute (the grid is sized from the occupancy query, so the CTAs are still
  // co-resident); needed under Nsight Compute, whose kernel replay rejects cooperative+cluster launches.
  // A rejected cooperative launch is an ERROR (the kernel contains a device-wide barrier; without the
  // driver's co-residency guarantee a concurrent kernel could starve it) -- there is no silent relaunch.  The
  // two-team experiment build (NTM_B200_DUAL_TEAM) cannot be launched cooperatively and stays opt-in.
  const bool coop = kv.cooperative_ok && !env.no_coop;
  e = kv.launch(R, W, p, ncta * nteams, hp.CS, smem_bytes, coop, stream);
  g_launches++;
  if (e != cudaSuccess) return set_cuda_error(e, "cudaLaunchKernelEx(ntm_seq_kernel)");
  if (prof) {
    cudaEventRecord(g_ev[2], stream);
    g_ev_valid = true;
  }
  return NTM_B200_OK;
}

int64_t ntm_b200_features_workspace_bytes(const ntm_b200_shape* shape, int64_t batch, int32_t frames, int32_t num_features) {
  if (!shape || batch < 1 || frames < 1 || num_features < 1) return -1;
  ntm_b200_plan plan{};
  const int64_t steps = (int64_t)frames * (num_features + 1);
  if (ntm_b200_query(shape, batch, steps, &plan) != NTM_B200_OK) return -1;
  return (plan.workspace_bytes + 255) / 256 * 256 + 4ll * batch * steps * shape->input_dim + 256;
}

int32_t ntm_b200_forward_seq_features(const ntm_b200_shape* shape, const ntm_b200_weights* weights, const void* packed,
                                      int64_t batch, int32_t frames, int32_t num_features, const float* features,
                                      const float* target, int32_t delimiter_first, const ntm_b200_state* state_in,
                                      const ntm_b200_state* state_out, float* logits, float* outputs, void* workspace,
                                      int64_t workspace_bytes, void* stream_v) {
  if (!shape || !features || !target) return NTM_B200_ERR_NULL_POINTER;
  if (frames < 1 || num_features < 1 || shape->input_dim < 3) return NTM_B200_ERR_BAD_SHAPE;
  ntm_b200::FeatureSource fs{features, target, frames, num_features, shape->input_dim - 2, delimiter_first ? 1 : 0};
  return forward_impl(shape, weights, packed, batch, (int64_t)frames * (num_features + 1), nullptr, state_in, state_out, logits,
                      outputs, nullptr, nullptr, workspace, workspace_bytes, stream_v, false, &fs);
}

int32_t ntm_b200_step(const ntm_b200_shape* shape, const ntm_b200_weights* weights,
                      const void* packed, int64_t batch, const float* inputs,
                      const ntm_b200_state* state_in, const ntm_b200_state* state_out,
                      float* logits, float* outputs, float* debug_taps, void* workspace,
                      int64_t workspace_bytes, void* stream) {
  return ntm_b200_forward_seq(shape, weights, packed, batch, 1, inputs, state_in, state_out, logits,
                              outputs, debug_taps, workspace, workspace_bytes, stream);
}

int32_t ntm_b200_copy_frames_h2d(float* frames_dev, const float* inputs_host, int64_t batch, int64_t steps,
                                 int64_t input_dim, int64_t t0, int64_t t1, void* stream) {
  if (!frames_dev || !inputs_host) return NTM_B200_ERR_NULL_POINTER;
  if (batch < 1 || input_dim < 1 || t0 < 0 || t1 <= t0 || t1 > steps) return NTM_B200_ERR_BAD_SHAPE;
  DeviceInfo di = device_info();
  if (!di.ok) return NTM_B200_ERR_NO_DEVICE;
  const size_t width = (size_t)(t1 - t0) * input_dim * sizeof(float);
  cudaError_t e = cudaMemcpy2DAsync(frames_dev, width, inputs_host + t0 * input_dim, (size_t)steps * input_dim * sizeof(float),
                                    width, (size_t)batch, cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return set_cuda_error(e, "cudaMemcpy2DAsync(frames)");
  return NTM_B200_OK;
}

int32_t ntm_b200_last_launch_info(int32_t* out16) {
  if (!out16) return NTM_B200_ERR_NULL_POINTER;
  for (int i = 0; i < 16; ++i) out16[i] = g_last_info[i];
  return NTM_B200_OK;
}

int32_t ntm_b200_set_profiling(int32_t enable) {
  g_profiling.store(enable ? 1 : 0);
  return NTM_B200_OK;
}

int32_t ntm_b200_last_kernel_ms(float* xproj_ms, float* seq_kernel_ms) {
  if (!xproj_ms || !seq_kernel_ms) return NTM_B200_ERR_NULL_POINTER;
  if (!g_ev_valid) return NTM_B200_ERR_BAD_SHAPE;
  cudaError_t e = cudaEventElapsedTime(xproj_ms, g_ev[0], g_ev[1]);
  if (e != cudaSuccess) return set_cuda_error(e, "cudaEventElapsedTime");
  e = cudaEventElapsedTime(seq_kernel_ms, g_ev[1], g_ev[2]);
  if (e != cudaSuccess) return set_cuda_error(e, "cudaEventElapsedTime");
  return NTM_B200_OK;
}

int32_t ntm_b200_last_stream_ms(float* out4, int32_t* steps) {
  if (!out4 || !steps) return NTM_B200_ERR_NULL_POINTER;
  *steps = stream_last_ms(out4);
  return NTM_B200_OK;
}

int32_t ntm_b200_stream_phase_ns(double* out9, int32_t* ctas) {
  if (!out9 || !ctas) return NTM_B200_ERR_NULL_POINTER;
  *ctas = stream_phase_ns(out9);
  return NTM_B200_OK;
}

int32_t ntm_b200_phase_cycles(const void* workspace, int64_t* out, int32_t max_ctas) {
  if (!workspace || !out) return NTM_B200_ERR_NULL_POINTER;
  if (max_ctas < 1 || max_ctas > 1024) return NTM_B200_ERR_BAD_SHAPE;
  cudaError_t e = cudaMemcpy(out, static_cast<const char*>(workspace) + 512, 8ll * PROF_SLOTS * max_ctas,
                             cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return set_cuda_error(e, "cudaMemcpy(phase cycles)");
  return NTM_B200_OK;
}

int32_t ntm_b200_finish(void* workspace, void* stream) {
  if (!workspace) return NTM_B200_ERR_NULL_POINTER;
  cudaError_t e = cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return set_cuda_error(e, "cudaStreamSynchronize");
  int flag = 0;
  e = cudaMemcpy(&flag, static_cast<char*>(workspace) + 256, sizeof(int), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return set_cuda_error(e, "cudaMemcpy(err flag)");
  return flag ? NTM_B200_ERR_DEVICE_TIMEOUT : NTM_B200_OK;
}

}  // extern "C"
