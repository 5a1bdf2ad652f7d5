#include <cstdio>
#include <cuda_runtime.h>
template <typename T>
__global__ void chain(T *out, int iters, T a, T b, long long *cyc)
{
    T v = (T)threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        v = v * a + b; v = v * a + b; v = v * a + b; v = v * a + b;
        v = v * a + b; v = v * a + b; v = v * a + b; v = v * a + b;
    }
    long long t1 = clock64();
    out[threadIdx.x] = v;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void chain_dadd_rn(double *out, int iters, double a, double b, long long *cyc)
{
    double v = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) v = __dadd_rn(__dmul_rn(v, a), b);
    }
    long long t1 = clock64();
    out[threadIdx.x] = v;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main()
{
    double *d; long long *c, h;
    cudaMalloc(&d, 4096); cudaMalloc(&c, 8);
    const int iters = 10000;
    for (int threads : {32, 128, 256, 512, 1024}) {
        chain<double><<<1, threads>>>(d, iters, 1.0000001, 1e-9, c);
        cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        printf("DFMA chain, %4d threads/SM: %.1f cycles per dependent op\n", threads, (double)h / (iters * 8));
    }
    chain<float><<<1, 32>>>((float *)d, iters, 1.0000001f, 1e-9f, c);
    cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("FFMA chain, 32 threads: %.1f cycles per dependent op\n", (double)h / (iters * 8));
    chain_dadd_rn<<<1, 32>>>(d, iters, 1.0000001, 1e-9, c);
    cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("DMUL+DADD chain, 32 threads: %.1f cycles per mul+add pair\n", (double)h / (iters * 8));
    return 0;
}
