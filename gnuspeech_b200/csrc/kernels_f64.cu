// kernels_f64.cu -- FP64 conformance instantiation.  MUST be compiled with -fmad=false: the reference
// build has no FMA contraction, and the 1e-9 conformance target is stated
// against that arithmetic.
#include "launch.cuh"
TRM_DEFINE_LAUNCHERS(double, f64)
