// dependent-issue latency of a few instructions on the device (one warp, clock64 around an unrolled chain)
#include <cstdio>
#include <cuda_runtime.h>
template <int OP> __global__ void k(double *out, long long *cyc, double a, double b, float fa, float fb)
{
    double x = a; float y = fa;
    long long t0 = clock64();
#pragma unroll
    for (int i = 0; i < 512; ++i) {
        if (OP == 0) x = __dadd_rn(x, b);
        if (OP == 1) x = __dmul_rn(x, b);
        if (OP == 2) x = __fma_rn(x, b, a);
        if (OP == 3) y = __fadd_rn(y, fb);
        if (OP == 4) y = __fmaf_rn(y, fb, fa);
        if (OP == 5) { x = __dadd_rn(x, b); x = (x > 511.0) ? x - 512.0 : x; }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = x + y; cyc[0] = t1 - t0; }
}
int main()
{
    double *o; long long *c, h;
    cudaMalloc(&o, 8); cudaMalloc(&c, 8);
    const char *names[] = {"DADD", "DMUL", "DFMA", "FADD", "FFMA", "DADD+wrap"};
    for (int op = 0; op < 6; ++op) {
        for (int warps = 1; warps <= 4; warps *= 4) {
            if (op == 0) k<0><<<1, 32 * warps>>>(o, c, 1.0, 1e-9, 1.f, 1e-6f);
            if (op == 1) k<1><<<1, 32 * warps>>>(o, c, 1.0, 1.0000001, 1.f, 1e-6f);
            if (op == 2) k<2><<<1, 32 * warps>>>(o, c, 1.0, 0.999, 1.f, 1e-6f);
            if (op == 3) k<3><<<1, 32 * warps>>>(o, c, 1.0, 1e-9, 1.f, 1e-6f);
            if (op == 4) k<4><<<1, 32 * warps>>>(o, c, 1.0, 1e-9, 1.f, 0.999f);
            if (op == 5) k<5><<<1, 32 * warps>>>(o, c, 1.0, 1.7, 1.f, 1e-6f);
            cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
            printf("%-10s warps %d: %.1f cycles per dependent op\n", names[op], warps, h / 512.0);
        }
    }
    return 0;
}
