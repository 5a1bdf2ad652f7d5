#include <cstdio>
#include <cuda_runtime.h>
template <int MODE> __global__ void wk(float4 *o, size_t n4)
{
    const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        if (MODE == 0) o[i] = v; else if (MODE == 1) __stcs(o + i, v); else __stwt(o + i, v);
    }
}
__global__ void rk(const float4 *o, size_t n4, float *out)
{
    float s = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) { float4 v = o[i]; s += v.x + v.y + v.z + v.w; }
    if (s == 12345.f) *out = s;
}
__global__ void ck(const float4 *a, float4 *b, size_t n4)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}
int main(){
    const size_t bytes = 8ull << 30; float4 *a, *b; float *o;
    cudaMalloc(&a, bytes); cudaMalloc(&b, bytes); cudaMalloc(&o, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto t = [&](const char *name, auto fn, double gb) { fn(); cudaDeviceSynchronize(); cudaEventRecord(e0); for (int r = 0; r < 3; ++r) fn(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); printf("%-28s %.3f ms  %.0f GB/s\n", name, ms / 3, gb / (ms / 3) * 1e3); };
    const size_t n4 = bytes / 16; const double gb = bytes / 1e9;
    for (int g : {148 * 8, 148 * 32}) {
        printf("grid %d\n", g);
        t("write default", [&] { wk<0><<<g, 256>>>(a, n4); }, gb);
        t("write stcs", [&] { wk<1><<<g, 256>>>(a, n4); }, gb);
        t("write stwt", [&] { wk<2><<<g, 256>>>(a, n4); }, gb);
        t("read", [&] { rk<<<g, 256>>>(a, n4, o); }, gb);
        t("copy (r+w bytes)", [&] { ck<<<g, 256>>>(a, b, n4); }, 2 * gb);
    }
    t("cudaMemset", [&] { cudaMemsetAsync(a, 0, bytes); }, gb);
    t("cudaMemcpy d2d (r+w)", [&] { cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice); }, 2 * gb);
    return 0;
}
