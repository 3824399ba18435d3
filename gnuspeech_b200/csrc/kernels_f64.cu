// kernels_f64.cu -- FP64 conformance instantiation (TRM_PRECISION_FP64): all state and arithmetic in double, within
// 1e-9 relative of the reference (BASELINE.json north_star).  FMA contraction allowed; the feed-forward phases use
// the cheaper forms described in tube_wide.cuh (geometric / rotation recurrences for the transcendentals of linearly
// interpolated parameters, fixed-point oscillator phase, Newton reciprocals).
#define TRM_KERNEL_NS trm_k64
#define TRM_STRICT 0
#include "launch.cuh"
TRM_DEFINE_LAUNCHERS(double, f64)
