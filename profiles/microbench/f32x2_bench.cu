#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float2 mul2(float2 a, float2 b){ float2 r; asm volatile("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mul.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}" : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y)); return r; }
__device__ __forceinline__ float2 add2(float2 a, float2 b){ float2 r; asm volatile("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}" : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y)); return r; }
template <int MODE> __global__ void k(float *out, int iters, float s)
{
    float2 a[8]; float2 m = make_float2(s, s * 1.0001f);
    for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { a[i].x = __fmul_rn(a[i].x, m.x); a[i].y = __fmul_rn(a[i].y, m.y); a[i].x = __fadd_rn(a[i].x, m.y); a[i].y = __fadd_rn(a[i].y, m.x); }
            else { a[i] = mul2(a[i], m); a[i] = add2(a[i], m); }
        }
    }
    float r = 0; for (int i = 0; i < 8; ++i) r += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
int main(){
    float *o; cudaMalloc(&o, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 2; ++mode) for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<148 * 8, 256>>>(o, iters, 1.0000001f); else k<1><<<148 * 8, 256>>>(o, iters, 1.0000001f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double flop_pairs = 148.0 * 8 * 256 * iters * 8 * 2;   // (mul+add) on 2 floats each
        printf("mode %d: %.3f ms, %.2f T scalar-ops/s, warp-instr/s %.2f T\n", mode, ms, flop_pairs * 2 / ms / 1e9, (mode == 0 ? flop_pairs * 2 : flop_pairs) / 32 / ms / 1e9);
    }
    return 0;
}
