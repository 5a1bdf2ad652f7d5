// mt_convert.cuh - MT19937 output stage shared by mt19937.cu and filtfilt.cu.
// numpy's RandomState.rand() (reference make_spect_f0.py:55): two successive 32-bit outputs a, b are
// tempered, then  ((a >> 5) * 2^26 + (b >> 6)) / 2^53.  The sequential generator kernel only twists
// the state and stores the RAW word pair per double; tempering and conversion are data parallel and
// run in the consumer (the last filtfilt kernel), off the generator's critical path.
#pragma once
#include <cstdint>

namespace ssfe {

__device__ __forceinline__ uint32_t mt_temper(uint32_t y)
{
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

__device__ __forceinline__ double mt_raw_to_double(uint2 raw)
{
    const uint32_t a = mt_temper(raw.x) >> 5, b = mt_temper(raw.y) >> 6;
    return (static_cast<double>(a) * 67108864.0 + static_cast<double>(b)) * (1.0 / 9007199254740992.0);   // exact
}

// The production path of ssfe_extract moves 4 bytes per sample through HBM instead of 8: of the word pair
// (a, b) that makes up one double only the raw word a travels (27 of the 53 random bits), and the consumer forms
// the dither term  (U - 0.5) * 1e-06  (make_spect_f0.py:55) in float from it.  U is then known to 2^-24, the
// term - at most 5e-7 in magnitude - to 3e-14: eight orders below one float ulp of the 0.1 ... 0.5 signal it is
// added to, and 1e7 times below the filter's own reproducibility (3e-7).  The validation paths (sequential
// filter mode, fp64 wav output, ssfe_rand) keep all 53 bits.
__device__ __forceinline__ float mt_a_to_dither_f32(uint32_t a_raw, float scale)
{
    const float u = __uint2float_rn(mt_temper(a_raw) >> 5);                 // 27 bits, rounded to float's 24
    return (fmaf(u, 1.0f / 134217728.0f, -0.5f)) * scale;
}

}  // namespace ssfe
