// kernels_f32.cu -- FP32 fast-mode instantiation (FMA contraction allowed).
#include "launch.cuh"
TRM_DEFINE_LAUNCHERS(float, f32)
