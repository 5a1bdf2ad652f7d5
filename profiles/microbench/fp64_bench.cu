#include <cstdio>
#include <cuda_runtime.h>
template <typename T, int ILP>
__global__ void fma_kernel(T *out, int iters, T a, T b)
{
    T v[ILP];
    for (int i = 0; i < ILP; ++i) v[i] = (T)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < ILP; ++i) v[i] = v[i] * a + b;
    T s = 0;
    for (int i = 0; i < ILP; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename T, int ILP>
__global__ void muladd_kernel(T *out, int iters, T a, T b)   // separate mul and add (no fma)
{
    T v[ILP];
    for (int i = 0; i < ILP; ++i) v[i] = (T)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (sizeof(T) == 8) v[i] = __dadd_rn(__dmul_rn(v[i], a), b);
            else v[i] = __fadd_rn(__fmul_rn(v[i], a), b);
        }
    T s = 0;
    for (int i = 0; i < ILP; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename K>
float timeit(K k)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k(); cudaDeviceSynchronize();
    cudaEventRecord(e0); k(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    void *buf; cudaMalloc(&buf, sms * 8 * 1024 * 8);
    const int iters = 20000, blocks = sms * 8, threads = 256;
    double n = (double)blocks * threads * iters * 8;
    float ms;
    ms = timeit([&] { fma_kernel<double, 8><<<blocks, threads>>>((double *)buf, iters, 1.0000001, 1e-9); });
    printf("SMs %d  DFMA: %.2f ms  %.3f T-instr/s  (%.2f TFLOP/s)\n", sms, ms, n / ms / 1e9, 2 * n / ms / 1e9);
    ms = timeit([&] { muladd_kernel<double, 8><<<blocks, threads>>>((double *)buf, iters, 1.0000001, 1e-9); });
    printf("DMUL+DADD: %.2f ms  %.3f T-instr/s\n", ms, 2 * n / ms / 1e9);
    ms = timeit([&] { fma_kernel<float, 8><<<blocks, threads>>>((float *)buf, iters, 1.0000001f, 1e-9f); });
    printf("FFMA: %.2f ms  %.3f T-instr/s  (%.2f TFLOP/s)\n", ms, n / ms / 1e9, 2 * n / ms / 1e9);
    ms = timeit([&] { muladd_kernel<float, 8><<<blocks, threads>>>((float *)buf, iters, 1.0000001f, 1e-9f); });
    printf("FMUL+FADD: %.2f ms  %.3f T-instr/s\n", ms, 2 * n / ms / 1e9);
    return 0;
}
