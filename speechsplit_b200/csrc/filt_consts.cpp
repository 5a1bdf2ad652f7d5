// filt_consts.cpp - host-side constants of the filtfilt scan (compiled by g++, not nvcc, because
// it uses __float128).
//
// scipy evaluates the order-5 filter of make_spect_f0.py:17,54 in ONE direct-form-II-transposed
// recurrence over (b, a).  For butter(5, 30/8000, 'high') that realisation is violently non-normal:
// the five poles sit in a cluster of radius 0.012 around z = 1 and the companion matrix A of `a` has
// max |A^k| ~ 1e8 near k = 256, so every fp64 rounding inside the recurrence is amplified by up to 1e8.
// scipy's own output is therefore reproducible only to ~3e-7 (perturbing its input by 1e-15 moves the
// result by 2.5e-7) - a smooth wobble below 30 Hz.  A CHUNKED evaluation of the same
// recurrence has the same error size, but the error restarts at every chunk boundary: a 62.5 Hz sawtooth
// of ~1e-7 whose harmonics land in the first two mel bands (94-150 Hz) and, in frames where those bands
// sit near the -100 dB floor, moved S by up to 1.8e-4 (full-corpus parity sweep of round 2, 36 of 656 M
// values above the 1e-4 gate).
//
// The scan therefore evaluates the SAME transfer function - the one the rounded fp64 coefficients
// (b, a) define, not the ideal Butterworth - as a cascade of one first-order and two second-order DF2T
// sections.  Each section pairs a pole group with a zero group next to it, its state matrix grows like
// k r^k <= 40 instead of 1e8, and plain fp64 reproduces the exact response of (b, a, zi) to ~1e-10
// (checked against an 80-bit evaluation of scipy's recurrence).  This file does the algebra in 113-bit
// arithmetic:
//   * roots of a(z) and of b(z) (Aberth iteration; the zeros of the ROUNDED b are a cluster of radius
//     1.5e-4 around 1, not a five-fold zero - they are what the reference's filter actually has);
//   * grouping into real sections, state-space matrices of the cascade (A_c, B_c, C_c, D_c);
//   * T with  O_c T = O  (observability matrices): z_c = T z maps scipy's DF2T state to the cascade
//     state with the same future output, so the initial condition  zi * x[0]  of scipy's filtfilt -
//     including whatever rounding lfilter_zi left in zi - is carried over exactly;
//   * A_c^chunk for the chunk-to-chunk carry (entries O(1): the carry is plain fp64 now).
#include <cmath>
#include <cstring>

namespace {

typedef __float128 q;

struct cq {
    q re, im;
};
inline cq operator+(cq a, cq b) { return {a.re + b.re, a.im + b.im}; }
inline cq operator-(cq a, cq b) { return {a.re - b.re, a.im - b.im}; }
inline cq operator*(cq a, cq b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
inline cq operator/(cq a, cq b)
{
    const q d = b.re * b.re + b.im * b.im;
    return {(a.re * b.re + a.im * b.im) / d, (a.im * b.re - a.re * b.im) / d};
}
inline q qabs(q x) { return x < 0 ? -x : x; }
inline q cabs2(cq a) { return a.re * a.re + a.im * a.im; }

// roots of c[0] z^n + c[1] z^(n-1) + ... + c[n] (real coefficients), Aberth-Ehrlich in 113 bits
bool poly_roots(const q *c, int n, cq *r)
{
    // start on a circle around the centroid of the roots (-c1 / (n c0)); a slight irrational twist
    // keeps the start points off the real axis and off each other's mirror images
    const q cen = -c[1] / (c[0] * n);
    for (int i = 0; i < n; ++i) {
        const double ang = 2.0 * M_PI * i / n + 0.7;
        r[i] = {cen + (q)(0.05 * std::cos(ang)), (q)(0.05 * std::sin(ang))};
    }
    for (int it = 0; it < 4000; ++it) {
        q worst = 0;
        for (int i = 0; i < n; ++i) {
            cq p = {c[0], 0}, dp = {0, 0};
            for (int k = 1; k <= n; ++k) {
                dp = dp * r[i] + p;
                p = p * r[i] + cq{c[k], 0};
            }
            if (cabs2(p) == 0) continue;
            const cq nt = p / dp;                 // Newton step
            cq s = {0, 0};
            for (int j = 0; j < n; ++j)
                if (j != i) s = s + cq{1, 0} / (r[i] - r[j]);
            const cq w = nt / (cq{1, 0} - nt * s);
            r[i] = r[i] - w;
            const q m = cabs2(w);
            if (m > worst) worst = m;
        }
        // Simple roots converge quadratically to ~1e-33.  The zeros of the rounded b are a cluster of radius
        // 1.5e-4, where 113-bit arithmetic leaves a noise floor of ~1e-18 in the iteration (root condition
        // 1 / |p'| ~ 1e15); 1e-17 is far below what the fp64 section coefficients resolve (1e-16 / 3e-4).
        if (worst < (q)1e-66) return true;
        if (it >= 200 && worst < (q)1e-34) return true;
    }
    return false;
}

struct Sec {          // (b0 + b1 z^-1 + b2 z^-2) / (1 + a1 z^-1 + a2 z^-2); order 1: b2 = a2 = 0
    int order;
    q b0, b1, b2, a1, a2;
};

// split n roots of a real polynomial into real ones and one representative (im > 0) of each conjugate pair
void split_roots(const cq *r, int n, q *re, int *n_re, cq *cx, int *n_cx)
{
    *n_re = *n_cx = 0;
    for (int i = 0; i < n; ++i) {
        if (qabs(r[i].im) < (q)1e-25) re[(*n_re)++] = r[i].re;
        else if (r[i].im > 0) cx[(*n_cx)++] = r[i];
    }
}

// quadratic (1, c1, c2) with the given root pair
struct Quad {
    int order;
    q c1, c2;
    q centre;          // where its roots sit (for pairing poles with zeros)
};

// group the roots into one first-order factor and two quadratics with real coefficients
bool group_roots(const cq *r, Quad *g /* [3], g[0] first order */)
{
    q re[5];
    cq cx[5];
    int nr, nc;
    split_roots(r, 5, re, &nr, cx, &nc);
    if (nr + 2 * nc != 5 || (nr & 1) == 0) return false;
    // sort the real roots; the middle one becomes the first-order factor, the others pair up from the outside
    for (int i = 0; i < nr; ++i)
        for (int j = i + 1; j < nr; ++j)
            if (re[j] < re[i]) { const q t = re[i]; re[i] = re[j]; re[j] = t; }
    const int mid = nr / 2;
    g[0] = {1, -re[mid], 0, re[mid]};
    int k = 1;
    for (int i = 0; i < mid; ++i) {
        const q x = re[i], y = re[nr - 1 - i];
        g[k++] = {2, -(x + y), x * y, (x + y) / 2};
    }
    for (int i = 0; i < nc; ++i) g[k++] = {2, -2 * cx[i].re, cabs2(cx[i]), cx[i].re};
    return k == 3;
}

// state space of one DF2T section:  z' = A z + B u,  y = C z + D u  (C = [1, 0])
void sec_ss(const Sec &s, q A[4], q B[2], q *D)
{
    A[0] = -s.a1; A[1] = 1; A[2] = -s.a2; A[3] = 0;
    B[0] = s.b1 - s.a1 * s.b0;
    B[1] = s.b2 - s.a2 * s.b0;
    *D = s.b0;
}

void mat5_mul(const q *X, const q *Y, q *Z)
{
    for (int i = 0; i < 5; ++i)
        for (int j = 0; j < 5; ++j) {
            q acc = 0;
            for (int k = 0; k < 5; ++k) acc += X[i * 5 + k] * Y[k * 5 + j];
            Z[i * 5 + j] = acc;
        }
}

// O[k][:] = C A^k, k = 0..4
void observability(const q *A, const q *C, q *O)
{
    q row[5], nx[5];
    for (int j = 0; j < 5; ++j) row[j] = C[j];
    for (int k = 0; k < 5; ++k) {
        for (int j = 0; j < 5; ++j) O[k * 5 + j] = row[j];
        for (int j = 0; j < 5; ++j) {
            q acc = 0;
            for (int i = 0; i < 5; ++i) acc += row[i] * A[i * 5 + j];
            nx[j] = acc;
        }
        std::memcpy(row, nx, sizeof(row));
    }
}

// solve X with  L X = R  (5 x 5, 5 right-hand sides), partial pivoting
bool solve5(q *L, q *R)
{
    for (int c = 0; c < 5; ++c) {
        int piv = c;
        for (int i = c + 1; i < 5; ++i)
            if (qabs(L[i * 5 + c]) > qabs(L[piv * 5 + c])) piv = i;
        if (L[piv * 5 + c] == 0) return false;
        if (piv != c)
            for (int j = 0; j < 5; ++j) {
                q t = L[c * 5 + j]; L[c * 5 + j] = L[piv * 5 + j]; L[piv * 5 + j] = t;
                t = R[c * 5 + j]; R[c * 5 + j] = R[piv * 5 + j]; R[piv * 5 + j] = t;
            }
        for (int i = c + 1; i < 5; ++i) {
            const q f = L[i * 5 + c] / L[c * 5 + c];
            for (int j = c; j < 5; ++j) L[i * 5 + j] -= f * L[c * 5 + j];
            for (int j = 0; j < 5; ++j) R[i * 5 + j] -= f * R[c * 5 + j];
        }
    }
    for (int c = 4; c >= 0; --c)
        for (int j = 0; j < 5; ++j) {
            q acc = R[c * 5 + j];
            for (int k = c + 1; k < 5; ++k) acc -= L[c * 5 + k] * R[k * 5 + j];
            R[c * 5 + j] = acc / L[c * 5 + c];
        }
    return true;
}

}  // namespace

// Cascade realisation of the order-5 filter (b6, a6) with scipy's initial-condition vector zi5.
//   sec   [3][5]  b0, b1, b2, a1, a2 of the sections in evaluation order (section 0 is first order: b2 = a2 = 0)
//   zic   [5]     cascade state equivalent to scipy's DF2T state zi (multiply by x[0] like scipy does)
//   m     [25]    A_c^chunk, row-major (state order: section 0, section 1 (2), section 2 (2))
// Returns 0, or a negative number when (b, a) does not factor the way the kernels need (root finding did not
// converge, an even number of real roots, a pole on or outside the unit circle).
extern "C" int ssfe_filt_cascade(const double *b6, const double *a6, const double *zi5, int chunk, double *sec,
                                 double *zic, double *m)
{
    q b[6], a[6];
    if (a6[0] == 0.0 || b6[0] == 0.0) return -1;
    for (int i = 0; i < 6; ++i) {
        b[i] = (q)b6[i] / (q)a6[0];
        a[i] = (q)a6[i] / (q)a6[0];
    }
    cq pz[5], zz[5];
    if (!poly_roots(a, 5, pz) || !poly_roots(b, 5, zz)) return -2;
    for (int i = 0; i < 5; ++i)
        if (cabs2(pz[i]) >= 1) return -3;
    Quad gp[3], gz[3];
    if (!group_roots(pz, gp) || !group_roots(zz, gz)) return -4;
    // pair each pole quadratic with the nearer zero quadratic (keeps every section's gain moderate)
    if (qabs(gp[1].centre - gz[1].centre) + qabs(gp[2].centre - gz[2].centre) >
        qabs(gp[1].centre - gz[2].centre) + qabs(gp[2].centre - gz[1].centre)) {
        const Quad t = gz[1];
        gz[1] = gz[2];
        gz[2] = t;
    }
    Sec s[3];
    s[0] = {1, b[0], b[0] * gz[0].c1, 0, gp[0].c1, 0};          // the overall gain b0 sits in the first section
    s[1] = {2, 1, gz[1].c1, gz[1].c2, gp[1].c1, gp[1].c2};
    s[2] = {2, 1, gz[2].c1, gz[2].c2, gp[2].c1, gp[2].c2};
    // the kernels run the sections with fp64 coefficients: round first, then derive everything else from the
    // rounded sections so that the state map and the carry matrix belong to exactly what the GPU evaluates
    for (int k = 0; k < 3; ++k) {
        s[k].b0 = (q)(double)s[k].b0; s[k].b1 = (q)(double)s[k].b1; s[k].b2 = (q)(double)s[k].b2;
        s[k].a1 = (q)(double)s[k].a1; s[k].a2 = (q)(double)s[k].a2;
        sec[k * 5 + 0] = (double)s[k].b0; sec[k * 5 + 1] = (double)s[k].b1; sec[k * 5 + 2] = (double)s[k].b2;
        sec[k * 5 + 3] = (double)s[k].a1; sec[k * 5 + 4] = (double)s[k].a2;
    }
    // cascade state space (5 states: [s0 | s1a s1b | s2a s2b]); u_k = output of section k-1
    q Ac[25], Cc[5];
    for (int i = 0; i < 25; ++i) Ac[i] = 0;
    for (int i = 0; i < 5; ++i) Cc[i] = 0;
    {
        q A0[4], B0[2], D0, A1[4], B1[2], D1, A2[4], B2[2], D2;
        sec_ss(s[0], A0, B0, &D0);
        sec_ss(s[1], A1, B1, &D1);
        sec_ss(s[2], A2, B2, &D2);
        // y0 = z[0] + D0 x;  y1 = z[1] + D1 y0;  y2 = z[3] + D2 y1
        Ac[0 * 5 + 0] = A0[0];                                      // section 0 (first order): z0' = -a1 z0 + B0 x
        // section 1 driven by y0 = z0 + D0 x
        Ac[1 * 5 + 1] = A1[0]; Ac[1 * 5 + 2] = A1[1]; Ac[1 * 5 + 0] = B1[0];
        Ac[2 * 5 + 1] = A1[2]; Ac[2 * 5 + 2] = A1[3]; Ac[2 * 5 + 0] = B1[1];
        // section 2 driven by y1 = z1 + D1 (z0 + D0 x)
        Ac[3 * 5 + 3] = A2[0]; Ac[3 * 5 + 4] = A2[1]; Ac[3 * 5 + 1] = B2[0]; Ac[3 * 5 + 0] = B2[0] * D1;
        Ac[4 * 5 + 3] = A2[2]; Ac[4 * 5 + 4] = A2[3]; Ac[4 * 5 + 1] = B2[1]; Ac[4 * 5 + 0] = B2[1] * D1;
        // y = y2 = z3 + D2 z1 + D2 D1 z0 + (D2 D1 D0) x
        Cc[3] = 1; Cc[1] = D2; Cc[0] = D2 * D1;
    }
    // scipy's DF2T realisation of (b, a)
    q Ad[25], Cd[5];
    for (int i = 0; i < 25; ++i) Ad[i] = 0;
    for (int i = 0; i < 5; ++i) {
        Ad[i * 5 + 0] = -a[i + 1];
        if (i + 1 < 5) Ad[i * 5 + i + 1] = 1;
        Cd[i] = (i == 0) ? 1 : 0;
    }
    q Oc[25], Od[25];
    observability(Ac, Cc, Oc);
    observability(Ad, Cd, Od);
    if (!solve5(Oc, Od)) return -5;                                 // Od now holds T
    for (int i = 0; i < 5; ++i) {
        q acc = 0;
        for (int j = 0; j < 5; ++j) acc += Od[i * 5 + j] * (q)zi5[j];
        zic[i] = (double)acc;
    }
    q P[25], Tm[25];
    for (int i = 0; i < 25; ++i) P[i] = (i % 6 == 0) ? 1 : 0;
    for (int k = 0; k < chunk; ++k) {
        mat5_mul(Ac, P, Tm);
        std::memcpy(P, Tm, sizeof(P));
    }
    for (int i = 0; i < 25; ++i) m[i] = (double)P[i];
    return 0;
}

// Taps of the zero-state chunk finals.  A chunk's final cascade state from a ZERO entry state is linear in its
// samples:  s = sum_i A_c^(chunk-1-i) B_c x[i].  The scan's local passes need nothing but s, so they evaluate these
// five dot products (filt_dot_kernel) instead of walking the recurrence.
//   sec15  the ROUNDED sections ssfe_filt_cascade returned (what Casc::step evaluates)
//   g      [5][chunk]: g[k * chunk + i] = (A_c^(chunk-1-i) B_c)[k], rounded from 113 bits
namespace {
// cascade state matrix A_c and input vector B_c of the rounded sections (what Casc::step evaluates)
void cascade_ab(const double *sec15, q *Ac, q *Bc)
{
    Sec s[3];
    for (int k = 0; k < 3; ++k)
        s[k] = {k == 0 ? 1 : 2, (q)sec15[k * 5 + 0], (q)sec15[k * 5 + 1], (q)sec15[k * 5 + 2], (q)sec15[k * 5 + 3],
                (q)sec15[k * 5 + 4]};
    q A0[4], B0[2], D0, A1[4], B1[2], D1, A2[4], B2[2], D2;
    sec_ss(s[0], A0, B0, &D0);
    sec_ss(s[1], A1, B1, &D1);
    sec_ss(s[2], A2, B2, &D2);
    for (int i = 0; i < 25; ++i) Ac[i] = 0;
    Ac[0 * 5 + 0] = A0[0];
    Ac[1 * 5 + 1] = A1[0]; Ac[1 * 5 + 2] = A1[1]; Ac[1 * 5 + 0] = B1[0];
    Ac[2 * 5 + 1] = A1[2]; Ac[2 * 5 + 2] = A1[3]; Ac[2 * 5 + 0] = B1[1];
    Ac[3 * 5 + 3] = A2[0]; Ac[3 * 5 + 4] = A2[1]; Ac[3 * 5 + 1] = B2[0]; Ac[3 * 5 + 0] = B2[0] * D1;
    Ac[4 * 5 + 3] = A2[2]; Ac[4 * 5 + 4] = A2[3]; Ac[4 * 5 + 1] = B2[1]; Ac[4 * 5 + 0] = B2[1] * D1;
    // x enters section 0 directly, section 1 through y0 = z0 + D0 x, section 2 through y1 = z1 + D1 y0
    Bc[0] = B0[0]; Bc[1] = B1[0] * D0; Bc[2] = B1[1] * D0; Bc[3] = B2[0] * D1 * D0; Bc[4] = B2[1] * D1 * D0;
}
}  // namespace

// Powers of the chunk-to-chunk carry matrix for the carry kernel's scan over runs of chunks:
//   out[k][25] = A_c^(chunk * run * 2^k),  k = 0 .. n_pow - 1   (row-major, rounded from 113 bits)
extern "C" int ssfe_filt_cascade_powers(const double *sec15, int chunk, int run, int n_pow, double *out)
{
    if (!sec15 || !out || chunk < 1 || run < 1 || n_pow < 1 || n_pow > 16) return -1;
    q Ac[25], Bc[5], P[25], T[25];
    cascade_ab(sec15, Ac, Bc);
    for (int i = 0; i < 25; ++i) P[i] = (i % 6 == 0) ? 1 : 0;
    for (long long k = 0; k < (long long)chunk * run; ++k) {
        mat5_mul(Ac, P, T);
        std::memcpy(P, T, sizeof(P));
    }
    for (int k = 0; k < n_pow; ++k) {
        for (int i = 0; i < 25; ++i) out[k * 25 + i] = (double)P[i];
        mat5_mul(P, P, T);
        std::memcpy(P, T, sizeof(P));
    }
    return 0;
}

extern "C" int ssfe_filt_cascade_taps(const double *sec15, int chunk, double *g)
{
    if (!sec15 || !g || chunk < 1) return -1;
    Sec s[3];
    for (int k = 0; k < 3; ++k)
        s[k] = {k == 0 ? 1 : 2, (q)sec15[k * 5 + 0], (q)sec15[k * 5 + 1], (q)sec15[k * 5 + 2], (q)sec15[k * 5 + 3],
                (q)sec15[k * 5 + 4]};
    q A0[4], B0[2], D0, A1[4], B1[2], D1, A2[4], B2[2], D2;
    sec_ss(s[0], A0, B0, &D0);
    sec_ss(s[1], A1, B1, &D1);
    sec_ss(s[2], A2, B2, &D2);
    q Ac[25];
    for (int i = 0; i < 25; ++i) Ac[i] = 0;
    Ac[0 * 5 + 0] = A0[0];
    Ac[1 * 5 + 1] = A1[0]; Ac[1 * 5 + 2] = A1[1]; Ac[1 * 5 + 0] = B1[0];
    Ac[2 * 5 + 1] = A1[2]; Ac[2 * 5 + 2] = A1[3]; Ac[2 * 5 + 0] = B1[1];
    Ac[3 * 5 + 3] = A2[0]; Ac[3 * 5 + 4] = A2[1]; Ac[3 * 5 + 1] = B2[0]; Ac[3 * 5 + 0] = B2[0] * D1;
    Ac[4 * 5 + 3] = A2[2]; Ac[4 * 5 + 4] = A2[3]; Ac[4 * 5 + 1] = B2[1]; Ac[4 * 5 + 0] = B2[1] * D1;
    // input vector: x enters section 0 directly, section 1 through y0 = z0 + D0 x, section 2 through y1 = z1 + D1 y0
    q v[5] = {B0[0], B1[0] * D0, B1[1] * D0, B2[0] * D1 * D0, B2[1] * D1 * D0};
    for (int m = 0; m < chunk; ++m) {                 // v = A_c^m B_c belongs to sample chunk-1-m
        for (int k = 0; k < 5; ++k) g[k * chunk + (chunk - 1 - m)] = (double)v[k];
        q nx[5];
        for (int i = 0; i < 5; ++i) {
            q acc = 0;
            for (int j = 0; j < 5; ++j) acc += Ac[i * 5 + j] * v[j];
            nx[i] = acc;
        }
        for (int i = 0; i < 5; ++i) v[i] = nx[i];
    }
    return 0;
}
