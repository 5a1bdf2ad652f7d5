// filt_consts.cpp - host-side constants of the filtfilt scan (compiled by g++, not nvcc, because
// it uses __float128).
//
// The chunk-to-chunk carry of the order-5 DF2T recurrence (scipy.signal.lfilter as used by
// filtfilt, reference make_spect_f0.py:54) is  z' = A^C z + s  with A the 5x5 companion matrix
// of the denominator:  A[i][0] = -a[i+1],  A[i][i+1] = 1.  For butter(5, 30/8000, 'high') the
// poles sit at |p| = 0.988..0.996 and max |A^k| reaches ~1e8 near k = 256, so A^C is formed in
// 113-bit arithmetic here and handed to the GPU as double-double (hi, lo) pairs.
#include <cstring>

extern "C" void ssfe_filt_power_dd(const double *a6, int power, double *hi25, double *lo25)
{
    typedef __float128 q;
    q A[25], R[25], T[25];
    for (int i = 0; i < 25; ++i) A[i] = 0;
    for (int i = 0; i < 5; ++i) {
        A[i * 5 + 0] = -(q)a6[i + 1];
        if (i + 1 < 5) A[i * 5 + i + 1] = 1;
    }
    for (int i = 0; i < 25; ++i) R[i] = (i % 6 == 0) ? 1 : 0;
    // plain repeated multiplication: `power` is a few hundred and every step is exact to ~1e-34
    for (int s = 0; s < power; ++s) {
        for (int i = 0; i < 5; ++i)
            for (int j = 0; j < 5; ++j) {
                q acc = 0;
                for (int k = 0; k < 5; ++k) acc += A[i * 5 + k] * R[k * 5 + j];
                T[i * 5 + j] = acc;
            }
        std::memcpy(R, T, sizeof(R));
    }
    for (int i = 0; i < 25; ++i) {
        const double h = (double)R[i];
        hi25[i] = h;
        lo25[i] = (double)(R[i] - (q)h);
    }
}

// P[j] = A^(base * j) for j = 1..count, each as [hi 25][lo 25]; out holds (count + 1) * 50 doubles and
// entry 0 is left zero.  M = A^base by repeated multiplication, then P[j] = P[j-1] M, all in 113 bits.
extern "C" void ssfe_filt_power_table_dd(const double *a6, int base, int count, double *out)
{
    typedef __float128 q;
    q A[25], M[25], P[25], T[25];
    for (int i = 0; i < 25; ++i) A[i] = 0;
    for (int i = 0; i < 5; ++i) {
        A[i * 5 + 0] = -(q)a6[i + 1];
        if (i + 1 < 5) A[i * 5 + i + 1] = 1;
    }
    auto mul = [](const q *X, const q *Y, q *Z) {
        for (int i = 0; i < 5; ++i)
            for (int j = 0; j < 5; ++j) {
                q acc = 0;
                for (int k = 0; k < 5; ++k) acc += X[i * 5 + k] * Y[k * 5 + j];
                Z[i * 5 + j] = acc;
            }
    };
    for (int i = 0; i < 25; ++i) M[i] = (i % 6 == 0) ? 1 : 0;
    for (int s = 0; s < base; ++s) {
        mul(A, M, T);
        std::memcpy(M, T, sizeof(M));
    }
    std::memcpy(P, M, sizeof(P));
    for (int i = 0; i < 50; ++i) out[i] = 0.0;
    for (int j = 1; j <= count; ++j) {
        for (int i = 0; i < 25; ++i) {
            const double h = (double)P[i];
            out[j * 50 + i] = h;
            out[j * 50 + 25 + i] = (double)(P[i] - (q)h);
        }
        mul(P, M, T);
        std::memcpy(P, T, sizeof(P));
    }
}
