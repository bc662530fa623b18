// FP32 instantiations of the rollout kernels (K1f).
#define B200MP_ROLLOUT_F32 1
#include "rollout_kernels.cuh"
