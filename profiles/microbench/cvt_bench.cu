// conversion / fp64 pipe throughput on sm_100a: thread-instructions per second
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE, int ILP>
__global__ void k(double *out, int iters, int seed)
{
    double acc[ILP];
    int iv[ILP];
    float fv[ILP];
    for (int i = 0; i < ILP; ++i) { acc[i] = threadIdx.x + i; iv[i] = seed + threadIdx.x * 7 + i; fv[i] = 1.0f + i + threadIdx.x; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (MODE == 0) {            // I2F.F64.S32 (+ int add to keep it alive)
                iv[i] += it;
                double d;
                asm volatile("cvt.rn.f64.s32 %0, %1;" : "=d"(d) : "r"(iv[i]));
                acc[i] = d;             // overwrite: only the conversion is on the fp64 side
                asm volatile("" : "+d"(acc[i]));
            } else if (MODE == 1) {     // F2F.F64.F32
                fv[i] = __int_as_float(__float_as_int(fv[i]) + 1);
                double d;
                asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d) : "f"(fv[i]));
                acc[i] = d;
                asm volatile("" : "+d"(acc[i]));
            } else if (MODE == 2) {     // F2F.F32.F64
                float f;
                asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f) : "d"(acc[i]));
                fv[i] = f;
                asm volatile("" : "+f"(fv[i]));
                long long b = __double_as_longlong(acc[i]) + 3;
                acc[i] = __longlong_as_double(b);
            } else if (MODE == 3) {     // DFMA reference
                acc[i] = fma(acc[i], 1.0000001, 1e-9);
            } else if (MODE == 4) {     // short -> double through the magic constant: PRMT/LOP + one DFMA
                iv[i] += it;
                const unsigned lo = (static_cast<unsigned>(iv[i]) & 0xffffu) ^ 0x8000u;
                const double d = __hiloint2double(0x43300000, lo);
                acc[i] = fma(d, 1.0 / 32768.0, -(137438953472.0 + 1.0));
                asm volatile("" : "+d"(acc[i]));
            } else if (MODE == 5) {     // float -> double by integer ops (normal numbers and zero)
                fv[i] = __int_as_float(__float_as_int(fv[i]) + 1);
                const unsigned b = __float_as_uint(fv[i]);
                const unsigned e = (b >> 23) & 0xffu;
                unsigned hi = (b & 0x80000000u) | (((b & 0x7fffffffu) >> 3) + 0x38000000u);
                if (e == 0) hi = b & 0x80000000u;
                acc[i] = __hiloint2double(hi, b << 29);
                asm volatile("" : "+d"(acc[i]));
            }
        }
    }
    double s = 0;
    for (int i = 0; i < ILP; ++i) s += acc[i] + fv[i] + iv[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename K>
float timeit(K kk)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    kk(); cudaDeviceSynchronize();
    cudaEventRecord(e0); kk(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount, iters = 4000, blocks = sms * 8, threads = 256;
    double *buf; cudaMalloc(&buf, sizeof(double) * blocks * threads);
    const double n = (double)blocks * threads * iters * 8;
    const char *names[6] = {"cvt.f64.s32", "cvt.f64.f32", "cvt.f32.f64", "DFMA", "s16->f64 magic+DFMA", "f32->f64 integer"};
    float ms;
    ms = timeit([&] { k<0, 8><<<blocks, threads>>>(buf, iters, 1); }); printf("%-22s %.3f ms  %.2f T/s\n", names[0], ms, n / ms / 1e9);
    ms = timeit([&] { k<1, 8><<<blocks, threads>>>(buf, iters, 1); }); printf("%-22s %.3f ms  %.2f T/s\n", names[1], ms, n / ms / 1e9);
    ms = timeit([&] { k<2, 8><<<blocks, threads>>>(buf, iters, 1); }); printf("%-22s %.3f ms  %.2f T/s\n", names[2], ms, n / ms / 1e9);
    ms = timeit([&] { k<3, 8><<<blocks, threads>>>(buf, iters, 1); }); printf("%-22s %.3f ms  %.2f T/s\n", names[3], ms, n / ms / 1e9);
    ms = timeit([&] { k<4, 8><<<blocks, threads>>>(buf, iters, 1); }); printf("%-22s %.3f ms  %.2f T/s\n", names[4], ms, n / ms / 1e9);
    ms = timeit([&] { k<5, 8><<<blocks, threads>>>(buf, iters, 1); }); printf("%-22s %.3f ms  %.2f T/s\n", names[5], ms, n / ms / 1e9);
    return 0;
}
