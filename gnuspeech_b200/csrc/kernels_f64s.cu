// kernels_f64s.cu -- FP64 STRICT instantiation (TRM_PRECISION_FP64_STRICT): every operation of the reference in the
// reference's order.  MUST be compiled with -fmad=false: the reference build has no FMA contraction (SURVEY.md
// Appendix A.18).  This is the bit-faithful twin the FP64 conformance mode is checked against at sizes the CPU oracle
// cannot reach.
#define TRM_KERNEL_NS trm_k64s
#define TRM_STRICT 1
#include "launch.cuh"
TRM_DEFINE_LAUNCHERS(double, f64s)
