// 256 threads per CTA, two CTAs per SM (two decoupled teams overlap each other's barrier / latency bubbles).
#define NTM_NT 256
#define NTM_MIN_CTAS 2
// 2 x 256 threads x 128 registers fills the register file exactly, and the hardware then only
// grants one CTA per SM (allocation overhead); 120 leaves the headroom two CTAs need.
#define NTM_MAXREG 120
#define NTM_KNS k256
#include "ntm_b200_seq_kernel.cuh"
