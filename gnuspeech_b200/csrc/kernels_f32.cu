// kernels_f32.cu -- FP32 fast-mode instantiation (FMA contraction allowed).
#define TRM_KERNEL_NS trm_k32
#define TRM_STRICT 0
#include "launch.cuh"
TRM_DEFINE_LAUNCHERS(float, f32)
