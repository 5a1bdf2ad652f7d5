// CPU emulation of the warp-level 1024-point FFT used by stft_mel.cu (same fft_core.cuh code,
// lanes executed one after another).  Checks the index math against a direct O(N^2) DFT.
//   g++ -O2 -std=c++17 -I speechsplit_b200/csrc tests/host_emu/fft_emu.cpp -o /tmp/fft_emu && /tmp/fft_emu
#include "fft_core.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <complex>
using namespace ssfe;

int main()
{
    const int N = 1024;
    const double pi = 3.14159265358979323846;
    std::vector<float> xa(N), xb(N), win(N);
    srand(1);
    for (int n = 0; n < N; ++n) {
        xa[n] = (rand() / (float)RAND_MAX - 0.5f);
        xb[n] = (rand() / (float)RAND_MAX - 0.5f) * 0.3f;
        win[n] = (float)(0.5 - 0.5 * cos(2 * pi * n / N));
    }
    std::vector<float2> tw(1024);
    for (int k1 = 0; k1 < 32; ++k1)
        for (int j = 0; j < 32; ++j) {
            double a = -2.0 * pi * (double)(j * k1) / 1024.0;
            tw[k1 * 32 + j] = make_float2((float)cos(a), (float)sin(a));
        }
    std::vector<float2> trans(32 * kTransStride);
    static float2 regs[32][32];
    for (int lane = 0; lane < 32; ++lane) {
        float2 v[32];
        for (int m = 0; m < 32; ++m) {
            int n = lane + 32 * m;
            v[m] = make_float2(xa[n] * win[n], xb[n] * win[n]);
        }
        fft32_dif(v);
        for (int r = 0; r < 32; ++r) {
            int k1 = bitrev5(r);
            float2 y = v[r];
            if (k1 != 0) {
                float2 t = tw[k1 * 32 + lane];
                y = make_float2(v[r].x * t.x - v[r].y * t.y, v[r].x * t.y + v[r].y * t.x);
            }
            trans[lane * kTransStride + k1] = y;
        }
    }
    for (int lane = 0; lane < 32; ++lane) {
        float2 v[32];
        for (int j = 0; j < 32; ++j) v[j] = trans[j * kTransStride + lane];
        fft32_dif(v);
        for (int r = 0; r < 32; ++r) regs[lane][r] = v[r];
    }
    std::vector<float> mA(520), mB(520);
    for (int lane = 0; lane < 32; ++lane) {
        int partner = (32 - lane) & 31;
        for (int k2 = 0; k2 < 16; ++k2) {
            float2 mine = regs[lane][bitrev5(k2)];
            float2 got = regs[partner][bitrev5(31 - k2)];
            if (lane == 0) got = regs[0][bitrev5((32 - k2) & 31)];
            float ar = mine.x + got.x, ai = mine.y - got.y, br = mine.x - got.x, bi = mine.y + got.y;
            mA[lane + 32 * k2] = 0.5f * sqrtf(ar * ar + ai * ai);
            mB[lane + 32 * k2] = 0.5f * sqrtf(br * br + bi * bi);
        }
        if (lane == 0) {
            float2 x = regs[0][bitrev5(16)];
            mA[512] = fabsf(x.x);
            mB[512] = fabsf(x.y);
        }
    }
    double maxerr = 0, maxref = 0;
    for (int k = 0; k <= 512; ++k) {
        std::complex<double> sa = 0, sb = 0;
        for (int n = 0; n < N; ++n) {
            std::complex<double> w = std::polar(1.0, -2 * pi * k * n / N);
            sa += (double)xa[n] * win[n] * w;
            sb += (double)xb[n] * win[n] * w;
        }
        maxerr = std::max(maxerr, std::abs(std::abs(sa) - mA[k]));
        maxerr = std::max(maxerr, std::abs(std::abs(sb) - mB[k]));
        maxref = std::max(maxref, std::abs(sa));
    }
    printf("max |err| = %.3e  (max |X| = %.3f)\n", maxerr, maxref);
    return maxerr < 1e-4 * maxref ? 0 : 1;
}
