// FP64 tensor-core rate on this GPU: independent chains of mma.sync.m8n8k4.f64 per warp, 1..16 warps per SM.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int ILP> __global__ void k(double *out, long long *cyc, double a, double b, int iters)
{
    double c[ILP][2];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) dmma(c[i][0], c[i][1], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int ILP> void run(int warps, double *o, long long *c)
{
    const int iters = 256;
    k<ILP><<<1, 32 * warps>>>(o, c, 1e-3, 1e-3, iters);
    long long h;
    cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    const double n = (double)iters * 8 * ILP;
    printf("ILP %2d warps %2d: %.2f cycles per DMMA per warp; %.1f FP64 MAC/clk/SM (vector DFMA peak: 64)\n", ILP, warps, h / n, warps * n * 256 / h);
}
int main()
{
    double *o; long long *c;
    cudaMalloc(&o, 8 * 2048); cudaMalloc(&c, 8);
    for (int w : {1, 4, 8, 16}) { run<1>(w, o, c); run<4>(w, o, c); run<8>(w, o, c); }
    return 0;
}
