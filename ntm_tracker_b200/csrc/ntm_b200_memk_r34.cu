// ntm_b200_memk_r34.cu -- the TMA-ring memory kernel's variants for 3 and 4 read heads (ntm_b200_memk.cuh),
// a translation unit of their own so that they compile next to ntm_b200_stream.cu's.
#include "ntm_b200_memk.cuh"

namespace ntm_b200 {
namespace memk {

cudaError_t launch_tma_r34(int R, int W, int CPL, const MemArgs& a, long long B, int smem, cudaStream_t stream, TmaCtl& ctl) {
  switch (R) {
    case 3: return launch_tma_r<3>(W, CPL, a, B, smem, stream, ctl);
    case 4: return launch_tma_r<4>(W, CPL, a, B, smem, stream, ctl);
  }
  return cudaErrorInvalidValue;
}

}  // namespace memk
}  // namespace ntm_b200
