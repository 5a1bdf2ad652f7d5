// fft_core.cuh - register-resident radix-32 building blocks of the 1024-point STFT kernel.
//
// Replaces np.fft.rfft inside utils.pySTFT (reference utils.py:28).  Two real frames are packed
// as the real and imaginary parts of one complex 1024-point FFT computed by ONE warp as a
// 32 x 32 Cooley-Tukey decomposition: every lane owns 32 complex points in registers, does a
// fully unrolled 32-point DIF (constant twiddles become immediates), the 32x32 transpose goes
// through shared memory, and the two spectra are separated with one warp shuffle per value.
//
// Everything here is SSFE_HD (host+device) so tests/host_emu.cpp can run the same index math on
// the CPU lane by lane - the build container has no GPU.
#pragma once

#ifdef __CUDACC__
#define SSFE_HD __host__ __device__ __forceinline__
#else
#define SSFE_HD inline
#include <cmath>
struct float2 { float x, y; };
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
#endif

namespace ssfe {

// Complex add / subtract.  On sm_100a these are the packed FP32 instructions (add.rn.f32x2 ->
// SASS FADD2): one issue slot for both components.  The STFT kernel is issue bound, not FMA-pipe
// bound, so halving the instruction count of the butterfly adds is worth ~15 % of the kernel.
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
SSFE_HD float2 cadd(float2 a, float2 b)
{
    float2 r;
    asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
SSFE_HD float2 csub(float2 a, float2 b)
{
    float2 r;
    asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; sub.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
// (a.x*s, a.y*s) and a*b + c component-wise
SSFE_HD float2 cscale(float2 a, float s)
{
    float2 r;
    asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%4}; mul.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(s));
    return r;
}
SSFE_HD float2 cfma_s(float2 a, float s, float2 c)
{
    float2 r;
    asm("{.reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%4}; mov.b64 rc, {%5,%6}; "
        "fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(s), "f"(c.x), "f"(c.y));
    return r;
}
SSFE_HD float2 cmul2(float2 a, float2 b)   // component-wise product
{
    float2 r;
    asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mul.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
#else
SSFE_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
SSFE_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
SSFE_HD float2 cscale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }
SSFE_HD float2 cfma_s(float2 a, float s, float2 c) { return make_float2(a.x * s + c.x, a.y * s + c.y); }
SSFE_HD float2 cmul2(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
#endif

// cos(2*pi*p/32), sin(2*pi*p/32) for p = 0..8 (first quadrant incl. ends)
SSFE_HD constexpr float cos32q(int p)
{
    return p == 0 ? 1.0f
         : p == 1 ? 0.98078528040323044913f
         : p == 2 ? 0.92387953251128675613f
         : p == 3 ? 0.83146961230254523708f
         : p == 4 ? 0.70710678118654752440f
         : p == 5 ? 0.55557023301960222474f
         : p == 6 ? 0.38268343236508977173f
         : p == 7 ? 0.19509032201612826785f
         : 0.0f;
}
// W_32^p = exp(-2*pi*i*p/32) for p = 0..15
SSFE_HD constexpr float w32_re(int p) { return p <= 8 ? cos32q(p) : -cos32q(16 - p); }
SSFE_HD constexpr float w32_im(int p) { return p <= 8 ? -cos32q(8 - p) : -cos32q(p - 8); }

SSFE_HD constexpr int bitrev5(int k)
{
    return ((k & 1) << 4) | ((k & 2) << 2) | (k & 4) | ((k & 8) >> 2) | ((k & 16) >> 4);
}

// d * W_32^p with the trivial cases folded at compile time (p is a constant after unrolling)
SSFE_HD float2 mul_w32(float2 d, const int p)
{
    if (p == 0) return d;
    if (p == 8) return make_float2(d.y, -d.x);                 // * (-i)
    if (p == 4) {                                              // * (1 - i)/sqrt2
        const float c = 0.70710678118654752440f;
        return make_float2((d.x + d.y) * c, (d.y - d.x) * c);
    }
    if (p == 12) {                                             // * (-1 - i)/sqrt2
        const float c = 0.70710678118654752440f;
        return make_float2((d.y - d.x) * c, -(d.x + d.y) * c);
    }
    const float wr = w32_re(p), wi = w32_im(p);
    return make_float2(d.x * wr - d.y * wi, d.x * wi + d.y * wr);
}

// In-place forward 32-point DFT (sign -1), decimation in frequency.
// On return X[k] is stored in v[bitrev5(k)].
SSFE_HD void fft32_dif(float2 (&v)[32])
{
#pragma unroll
    for (int span = 32; span >= 2; span >>= 1) {
#pragma unroll
        for (int base = 0; base < 32; base += span) {
#pragma unroll
            for (int i = 0; i < span / 2; ++i) {
                const float2 a = v[base + i], b = v[base + i + span / 2];
                v[base + i] = cadd(a, b);
                v[base + i + span / 2] = mul_w32(csub(a, b), i * (32 / span));
            }
        }
    }
}

// Layout of the 1024-point transform over one warp (lane = 0..31):
//   input  n  = lane + 32*m      (m = register index of pass 1)
//   pass 1 : Y_lane[k1] = sum_m z[lane+32m] W_32^(m k1)            -> register bitrev5(k1)
//   twiddle: Y_lane[k1] *= W_1024^(lane*k1)                          (table tw[k1*32+lane])
//   transpose through shared memory, pass 2 runs in lane k1 over j:
//            X[k1 + 32*k2] = sum_j Y_j[k1] W_32^(j k2)              -> register bitrev5(k2)
// Separation of the two real frames A (real part) and B (imaginary part):
//   A[k] = (X[k] + conj(X[1024-k]))/2 ,  B[k] = (X[k] - conj(X[1024-k]))/(2i)
// Lane k1 handles k = k1 + 32*k2 for k2 = 0..15 (lane 0 also k2 = 16, i.e. k = 512); the partner
// value X[1024-k] lives in lane (32-k1)&31, register bitrev5(31-k2) (lane 0: bitrev5((32-k2)&31)).
constexpr int kTransStride = 33;   // float2 elements per row of the transpose buffer

}  // namespace ssfe
