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

// The production path of ssfe_extract stores the dither term itself,  (U - 0.5) * 1e-06  (make_spect_f0.py:55),
// rounded once more to float: 4 bytes per sample instead of 8 through HBM.  The term is <= 5e-7 in magnitude, so
// the extra rounding is <= 3e-14 absolute - eight orders below one float ulp of the signal it is added to.
__device__ __forceinline__ float mt_raw_to_dither_f32(uint2 raw, double scale)
{
    return static_cast<float>(__dmul_rn(__dsub_rn(mt_raw_to_double(raw), 0.5), scale));
}

}  // namespace ssfe
