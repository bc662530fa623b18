// FP64 instantiations of the rollout kernels (K1, K1c, generic, logging) and the batched planar_model kernel.
#define B200MP_ROLLOUT_F64 1
#include "rollout_kernels.cuh"
