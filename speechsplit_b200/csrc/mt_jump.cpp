// mt_jump.cpp - host-side jump-ahead polynomials of MT19937 (compiled by g++).
//
// numpy.random.RandomState(seed).rand(L) (reference make_spect_f0.py:47,55) is one serial stream
// per speaker; the GPU generator cuts every stream into segments of kJumpBlocks state blocks and
// starts each segment from a state computed directly from the seed.  MT19937's state transition is
// linear over GF(2): with phi(x) its characteristic polynomial (degree 19937) and
//     g(x) = x^J mod phi(x) = sum_i g_i x^i,
// every bit sequence of the untempered word sequence w[n] obeys  w[n + J] = XOR_{g_i = 1} w[n + i]
// (n >= 1, and n >= 0 for the top bit, which is all the recurrence reads of w[n]).  The GPU kernel
// (mt19937.cu, mt_jump_kernel) evaluates that sum for the 624 words of the new state; this file
// supplies g as a list of tap positions.
//
//   phi   : Berlekamp-Massey on 2 * 19937 bits of the generator's own output (no constants to trust);
//           checked to have degree 19937
//   x^J   : square-and-multiply in GF(2)[x] / phi
//   g_a g_b mod phi : 4-bit comb multiplication + sparse top-down reduction
//
// Everything is plain integer arithmetic, deterministic, and cached per process.
#include <cstdint>
#include <cstring>
#include <mutex>
#include <vector>

namespace {

constexpr int kDeg = 19937;
constexpr int kW = 312;            // 64-bit words of a residue (19968 bits)
constexpr int kN = 624, kM = 397;

typedef std::vector<uint64_t> Bits;

inline int get_bit(const uint64_t *p, int i) { return static_cast<int>((p[i >> 6] >> (i & 63)) & 1u); }
inline void flip_bit(uint64_t *p, int i) { p[i >> 6] ^= 1ull << (i & 63); }

// dst ^= src << sh   (bit shift, src has n words, dst must hold the shifted range)
void xor_shifted(uint64_t *dst, const uint64_t *src, int n, int sh)
{
    const int ws = sh >> 6, bs = sh & 63;
    if (bs == 0) {
        for (int i = 0; i < n; ++i) dst[i + ws] ^= src[i];
    } else {
        for (int i = 0; i < n; ++i) {
            dst[i + ws] ^= src[i] << bs;
            dst[i + ws + 1] ^= src[i] >> (64 - bs);
        }
    }
}

struct Field {
    std::vector<int> terms;        // exponents e < kDeg with phi = x^kDeg + sum x^e
    bool wordwise = false;         // all terms at least 64 below the degree
    bool ok = false;
};

// LSB of the untempered word sequence w[1], w[2], ... of init_genrand(seed)
void lsb_sequence(uint32_t seed, int n_bits, uint64_t *out)
{
    std::vector<uint32_t> mt(kN);
    mt[0] = seed;
    for (int i = 1; i < kN; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + static_cast<uint32_t>(i);
    int produced = 0;
    auto put = [&](uint32_t v) {
        if (produced < n_bits) {
            if (v & 1u) flip_bit(out, produced);
            ++produced;
        }
    };
    for (int i = 1; i < kN; ++i) put(mt[i]);
    while (produced < n_bits) {
        for (int k = 0; k < kN; ++k) {
            const uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % kN] & 0x7fffffffu);
            mt[k] = mt[(k + kM) % kN] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        for (int k = 0; k < kN; ++k) put(mt[k]);
    }
}

Field build_field()
{
    Field f;
    const int n_bits = 2 * kDeg + 64;
    const int words = (n_bits + 63) / 64 + 2;
    Bits s(words, 0);
    lsb_sequence(5489u, n_bits, s.data());
    // Berlekamp-Massey over GF(2).  C, B are connection polynomials (bit i = coefficient of x^i),
    // R holds the sequence reversed: bit i = s[n - i].
    const int pw = kW + 2;
    Bits C(pw + 1, 0), B(pw + 1, 0), T(pw + 1, 0), R(pw + 1, 0);
    C[0] = B[0] = 1;
    int L = 0, m = 1, bw = 1;      // bw: words of B that can be non-zero
    for (int n = 0; n < n_bits; ++n) {
        // R = (R << 1) | s[n]
        for (int i = pw; i > 0; --i) R[i] = (R[i] << 1) | (R[i - 1] >> 63);
        R[0] = (R[0] << 1) | static_cast<uint64_t>(get_bit(s.data(), n));
        uint64_t acc = 0;
        const int lw = (L >> 6) + 1;
        for (int i = 0; i < lw; ++i) acc ^= C[i] & R[i];
        const int d = __builtin_parityll(acc);
        if (!d) {
            ++m;
            continue;
        }
        if ((m >> 6) + bw + 1 > pw + 1) return f;         // degenerate: not a 19937-degree recurrence
        if (2 * L <= n) {
            T = C;
            xor_shifted(C.data(), B.data(), bw, m);
            bw = (L >> 6) + 1;
            L = n + 1 - L;
            B = T;
            m = 1;
        } else {
            xor_shifted(C.data(), B.data(), bw, m);
            ++m;
        }
        if (L > kDeg) return f;
    }
    if (L != kDeg) return f;
    // s[n] = sum_{i=1..L} C_i s[n-i]  <=>  phi(x) = x^L C(1/x):  coefficient of x^(L-i) is C_i
    for (int i = 1; i <= kDeg; ++i)
        if (get_bit(C.data(), i)) f.terms.push_back(kDeg - i);
    int top = 0;
    for (int e : f.terms) top = e > top ? e : top;
    f.wordwise = (kDeg - top) >= 64;
    f.ok = !f.terms.empty();
    return f;
}

const Field &field()
{
    static Field f;
    static std::once_flag once;
    std::call_once(once, [] { f = build_field(); });
    return f;
}

// p: 2 * kW + 2 words, degree < 2 * kDeg; reduced in place to degree < kDeg
void reduce(uint64_t *p, const Field &f)
{
    const int top_word = 2 * kW + 1;
    if (f.wordwise) {
        for (int W = top_word; W >= kDeg / 64; --W) {
            uint64_t v = p[W];
            int base = 64 * W - kDeg;                 // x^(64W + b) = sum_e x^(base + b + e)
            if (W == kDeg / 64) v &= ~0ull << (kDeg & 63);
            if (!v) continue;
            p[W] ^= v;
            int sh = 0;
            if (base < 0) { sh = -base; v >>= sh; base = 0; }   // the dropped low bits were masked off above
            for (int e : f.terms) {
                const int pos = base + e, ws = pos >> 6, bs = pos & 63;
                p[ws] ^= v << bs;
                if (bs) p[ws + 1] ^= v >> (64 - bs);
            }
        }
    } else {
        for (int i = 64 * (top_word + 1) - 1; i >= kDeg; --i) {
            if (!get_bit(p, i)) continue;
            flip_bit(p, i);
            for (int e : f.terms) flip_bit(p, i - kDeg + e);
        }
    }
}

// r = a * b mod phi; a, b, r: kW words
void mulmod(const uint64_t *a, const uint64_t *b, uint64_t *r, const Field &f)
{
    // T[u] = u(x) * b(x) for the 16 polynomials u of degree < 4
    static thread_local std::vector<uint64_t> T, P;
    T.assign(16 * (kW + 1), 0);
    P.assign(2 * kW + 3, 0);
    for (int u = 1; u < 16; ++u)
        for (int k = 0; k < 4; ++k)
            if (u >> k & 1) xor_shifted(&T[u * (kW + 1)], b, kW, k);
    for (int nib = 15; nib >= 0; --nib) {
        if (nib != 15) {
            for (int i = 2 * kW + 1; i > 0; --i) P[i] = (P[i] << 4) | (P[i - 1] >> 60);
            P[0] <<= 4;
        }
        for (int wi = 0; wi < kW; ++wi) {
            const unsigned u = static_cast<unsigned>(a[wi] >> (4 * nib)) & 15u;
            if (!u) continue;
            const uint64_t *t = &T[u * (kW + 1)];
            uint64_t *d = &P[wi];
            for (int i = 0; i <= kW; ++i) d[i] ^= t[i];
        }
    }
    reduce(P.data(), f);
    std::memcpy(r, P.data(), kW * sizeof(uint64_t));
}

// r = x^e mod phi
void powx(uint64_t e, uint64_t *r, const Field &f)
{
    Bits acc(kW, 0), tmp(kW, 0), x1(kW, 0);
    acc[0] = 1;
    x1[0] = 2;
    int hi = 63;
    while (hi > 0 && !(e >> hi & 1)) --hi;
    for (int bit = hi; bit >= 0; --bit) {
        mulmod(acc.data(), acc.data(), tmp.data(), f);
        acc = tmp;
        if (e >> bit & 1) {
            mulmod(acc.data(), x1.data(), tmp.data(), f);
            acc = tmp;
        }
    }
    std::memcpy(r, acc.data(), kW * sizeof(uint64_t));
}

struct Level {
    std::vector<Bits> g;           // g[d] = x^(d * unit) mod phi, d = 1..; g[0] unused
};
struct Cache {
    std::mutex mu;
    uint64_t unit_words = 0;       // words per segment
    Level lv[3];
};
Cache &cache()
{
    static Cache c;
    return c;
}

}  // namespace

// Characteristic polynomial: returns the number of terms below x^19937 (0 = failed) and, if exps is
// not null, writes up to cap exponents.
extern "C" int ssfe_mt_charpoly_terms(int *exps, int cap)
{
    const Field &f = field();
    if (!f.ok) return 0;
    if (exps)
        for (int i = 0; i < static_cast<int>(f.terms.size()) && i < cap; ++i) exps[i] = f.terms[i];
    return static_cast<int>(f.terms.size());
}

// x^n_words mod phi as 312 little-endian 64-bit words (bit i = coefficient of x^i).  0 on success.
extern "C" int ssfe_mt_jump_poly(uint64_t n_words, uint64_t *poly312)
{
    const Field &f = field();
    if (!f.ok) return -1;
    powx(n_words, poly312, f);
    return 0;
}

// Tap list of x^(d * 256^level * unit_words) mod phi, d = 1..255, level = 0..2: the exponents with a
// non-zero coefficient, ascending.  Polynomials are built incrementally (g_d = g_(d-1) g_1) and cached
// for the life of the process.  Returns the number of taps (<= 19937), or -1.
extern "C" int ssfe_mt_jump_taps(uint64_t unit_words, int level, int d, uint16_t *taps, int cap)
{
    const Field &f = field();
    if (!f.ok || level < 0 || level > 2 || d < 1 || d > 255 || unit_words == 0) return -1;
    Cache &c = cache();
    std::lock_guard<std::mutex> lock(c.mu);
    if (c.unit_words != unit_words) {
        for (Level &l : c.lv) l.g.clear();
        c.unit_words = unit_words;
    }
    Level &l = c.lv[level];
    if (l.g.empty()) {
        l.g.resize(2, Bits(kW, 0));
        uint64_t e = unit_words;
        for (int i = 0; i < level; ++i) {
            if (e > (~0ull >> 8)) return -1;
            e <<= 8;
        }
        powx(e, l.g[1].data(), f);
    }
    while (static_cast<int>(l.g.size()) <= d) {
        Bits nx(kW, 0);
        mulmod(l.g.back().data(), l.g[1].data(), nx.data(), f);
        l.g.push_back(nx);
    }
    const Bits &g = l.g[d];
    int n = 0;
    for (int i = 0; i < kDeg; ++i)
        if (get_bit(g.data(), i)) {
            if (taps && n < cap) taps[n] = static_cast<uint16_t>(i);
            ++n;
        }
    return n;
}
