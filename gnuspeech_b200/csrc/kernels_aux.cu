// kernels_aux.cu -- kernels that are the same in both precision modes (compiled with -fmad=false: the frame
// generator's float / double arithmetic follows the reference operation by operation).
#include "framegen_kernel.cuh"

extern "C" int trm_k_framegen(const trm::FrameGenArgs *a, cudaStream_t s)
{
    if (a->n_utt <= 0) return 0;
    const int warps = 4;
    trm::framegen_kernel<<<(a->n_utt + warps - 1) / warps, warps * 32, 0, s>>>(*a);
    return (int)cudaGetLastError();
}
