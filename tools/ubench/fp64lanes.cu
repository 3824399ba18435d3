// Does a DFMA of a partially active warp occupy the FP64 pipe for less time?  One warp, ILP 8, lanes >= `active` exit
// before the loop; second table: the same with the active lanes spread (every other lane).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, long long *cyc, double b, double c, int iters, unsigned mask)
{
    if (!((mask >> (threadIdx.x & 31)) & 1u)) return;
    constexpr int ILP = 8;
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = 1.0 + i + threadIdx.x * 1e-3;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], b, c);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[threadIdx.x] = s;
    if ((threadIdx.x & 31) == __ffs(mask) - 1 && threadIdx.x < 32) cyc[0] = t1 - t0;
}
int main()
{
    double *o; long long *c;
    cudaMalloc(&o, 8 * 2048); cudaMalloc(&c, 8);
    const int iters = 256;
    struct { const char *name; unsigned mask; } cases[] = {
        {"32 lanes", 0xffffffffu}, {"lanes 0-15", 0x0000ffffu}, {"lanes 16-31", 0xffff0000u}, {"lanes 0-7", 0xffu},
        {"even lanes", 0x55555555u}, {"lanes 0-13", 0x3fffu}, {"lanes 0-27", 0x0fffffffu}, {"1 lane", 1u}};
    for (int warps : {1, 4, 5, 8}) for (auto &cs : cases) {
        k<<<1, 32 * warps>>>(o, c, 0.999999, 1e-7, iters, cs.mask);
        long long h;
        cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        const double n = (double)iters * 16 * 8;
        printf("warps %d %-12s: %.2f cycles per DFMA per warp\n", warps, cs.name, h / n);
    }
    return 0;
}
