// FP64 issue rate of ONE warp (and of several warps on one SM sub-partition / SM): ILP independent DFMA chains.
// Answers: can a single warp keep its sub-partition's FP64 pipe busy?  What is the dependent-issue latency?
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP> __global__ void k(double *out, long long *cyc, double b, double c, int iters)
{
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
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int ILP> void run(int warps, double *o, long long *c)
{
    const int iters = 256;
    k<ILP><<<1, 32 * warps>>>(o, c, 0.999999, 1e-7, iters);
    long long h;
    cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    const double n = (double)iters * 16 * ILP;
    printf("ILP %2d warps %2d: %.2f cycles per DFMA per warp (%.2f warp-DFMA/clk/SM)\n", ILP, warps, h / n, warps * n / h);
}
int main()
{
    double *o; long long *c;
    cudaMalloc(&o, 8 * 2048); cudaMalloc(&c, 8);
    for (int w : {1, 2, 4, 8, 16}) { run<1>(w, o, c); run<2>(w, o, c); run<4>(w, o, c); run<8>(w, o, c); run<16>(w, o, c); }
    return 0;
}
