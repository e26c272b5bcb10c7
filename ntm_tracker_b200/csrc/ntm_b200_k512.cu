// 512 threads per CTA, one CTA per SM (one team): the shape used when a CTA needs more than half an SM's shared memory.
#define NTM_NT 512
#define NTM_MIN_CTAS 1
#define NTM_KNS k512
#include "ntm_b200_seq_kernel.cuh"
