// filtfilt.cu - stages a0 + a1 (+ the deterministic part of a2).
//
// Replaces reference make_spect_f0.py:52-55:
//     if L % 256 == 0: x = concat(x, [1e-06])
//     y   = scipy.signal.filtfilt(b, a, x)           # b, a = butter_highpass(30, 16000, 5)
//     wav = y * 0.96 + (prng.rand(L) - 0.5) * 1e-06
// scipy's filtfilt (scipy/signal/_signaltools.py filtfilt/_validate_pad, _arraytools.odd_ext):
// odd-extend by 18 samples, run the direct-form-II-transposed recurrence forward from
// zi * ext[0], run it again over the reversed result from zi * y[-1], reverse, drop the pads.
//
// The order-5 recurrence is sequential in time.  It is evaluated as a chunked parallel scan:
//   1. local : one thread per 256-sample chunk runs the recurrence from a ZERO state and keeps
//              only the final state s_c (5 doubles);
//   2. carry : per utterance, z_in[c+1] = M z_in[c] + s_c with M = A^256;
//   3. final : one thread per chunk re-runs the recurrence from its true z_in and writes outputs.
// scipy's own realisation - one DF2T recurrence over (b, a) - cannot be chunked like this without leaving a
// trace: its state matrix has max |A^k| ~ 1e8, every rounding is amplified accordingly, and although the
// resulting error (~1e-7) is no larger than scipy's own, it RESTARTS at every chunk boundary - a 62.5 Hz
// sawtooth whose harmonics reach the first mel bands (filt_consts.cpp has the measurements).  The scan
// therefore evaluates the same transfer function and the same initial condition as a cascade of one
// first-order and two second-order sections (constants from filt_consts.cpp, 113-bit algebra): well
// conditioned, plain fp64 with FMAs, exact response of (b, a, zi) to ~1e-10; what remains against
// scipy.signal.filtfilt is scipy's own fp64 wobble (<= ~4e-7, below 30 Hz).
// The backward pass fuses the dither combine and scatters wav (f32) into the padded segment
// layout the STFT kernel reads.
// filtfilt_mode 1 runs one thread per utterance with scipy's recurrence in scipy's operation order
// ((z[i+1] + x*b) - y*a, no FMA contraction): bit-identical to scipy, a validation aid.
#include "common.cuh"
#include "mt_convert.cuh"
#include "tma.cuh"
#include <algorithm>

namespace ssfe {

extern "C" int ssfe_filt_cascade(const double *b6, const double *a6, const double *zi5, int chunk, double *sec,
                                 double *zic, double *m);     // filt_consts.cpp
extern "C" int ssfe_filt_cascade_taps(const double *sec15, int chunk, double *g);
extern "C" int ssfe_filt_cascade_powers(const double *sec15, int chunk, int run, int n_pow, double *out);

constexpr int kChunk = 256;
constexpr int kPadLen = 18;       // 3 * max(len(a), len(b))
constexpr int kFiltThreads = 128;

struct FiltConsts {
    double b[6], a[6], zi[5];     // scipy's realisation (sequential validation mode)
    double sec[15];               // cascade sections [3][b0, b1, b2, a1, a2] (section 0 is first order)
    double zic[5];                // cascade state equivalent to zi
    double mc[25];                // A_c^kChunk, row-major
    double mp[5][25];             // A_c^(kChunk * 8 * 2^k): the carry kernel's scan over runs of eight chunks
    double wav_scale, dither_scale;
};

struct FiltParams {
    const void *x;                // input samples (pass 1) or nullptr
    const double *y1;             // forward result over the extended signal (pass 2 input), sequential mode
    const float *y1f;             // the same for the tiled kernels: stored as float (see filt_tile_kernel)
    float *y1f_out;
    const int64_t *in_off;        // [n+1] raw input offsets
    const int64_t *fix_off;       // [n+1] fixed (post-append) offsets
    const int *chunk_off;         // [n+1] prefix of chunk counts
    int n, n_chunks, chunk_len;
    double *state;                // [n_chunks][5] zero-state finals of the local pass
    double *zin;                  // [n_chunks][5] true chunk-entry states (carry -> final pass)
    // outputs of the backward pass
    double *y;
    const double *dith;
    float *wavp;
    const int64_t *seg_off;
    float *wav;
    double *wav64;
    double *y1_out;               // forward pass output (extended)
    int dith_raw;                 // dith holds raw MT19937 word pairs
    int dith_f32;                 // dith holds ONE raw MT19937 word per sample (mt_convert.cuh: mt_a_to_dither_f32)
    int dith_pf;                  // prefetch of the dither words: 0 none, 1 into L2, 2 into L1
    int rows;                     // chunks per tile (= per warp): 32, or 8 for a batch too small to fill the GPU
};

__constant__ FiltConsts c_filt;

template <int DTYPE>
__device__ __forceinline__ double load_sample(const void *x, int64_t i)
{
    if (DTYPE == SSFE_F32) return static_cast<double>(static_cast<const float *>(x)[i]);
    if (DTYPE == SSFE_F64) return static_cast<const double *>(x)[i];
    return static_cast<double>(static_cast<const short *>(x)[i]) * (1.0 / 32768.0);
}

// x'[n] of the fixed-length signal: the appended sample is 1e-06 (make_spect_f0.py:53)
template <int DTYPE>
__device__ __forceinline__ double fixed_sample(const void *x, int64_t base, int64_t L, int64_t n)
{
    return (n < L) ? load_sample<DTYPE>(x, base + n) : 1e-06;
}

// ext[j], j in [0, Lf + 36): scipy odd_ext(x, 18)
template <int DTYPE>
__device__ __forceinline__ double ext_sample(const void *x, int64_t base, int64_t L, int64_t Lf, int64_t j)
{
    if (j < kPadLen)
        return __dsub_rn(2.0 * fixed_sample<DTYPE>(x, base, L, 0), fixed_sample<DTYPE>(x, base, L, kPadLen - j));
    if (j < kPadLen + Lf) return fixed_sample<DTYPE>(x, base, L, j - kPadLen);
    const int64_t k = j - kPadLen - Lf;
    return __dsub_rn(2.0 * fixed_sample<DTYPE>(x, base, L, Lf - 1), fixed_sample<DTYPE>(x, base, L, Lf - 2 - k));
}

struct Df2t {
    double z0, z1, z2, z3, z4;
    __device__ __forceinline__ double step(double x)
    {
        const FiltConsts &c = c_filt;
        const double y = __dadd_rn(z0, __dmul_rn(c.b[0], x));
        z0 = __dsub_rn(__dadd_rn(z1, __dmul_rn(x, c.b[1])), __dmul_rn(y, c.a[1]));
        z1 = __dsub_rn(__dadd_rn(z2, __dmul_rn(x, c.b[2])), __dmul_rn(y, c.a[2]));
        z2 = __dsub_rn(__dadd_rn(z3, __dmul_rn(x, c.b[3])), __dmul_rn(y, c.a[3]));
        z3 = __dsub_rn(__dadd_rn(z4, __dmul_rn(x, c.b[4])), __dmul_rn(y, c.a[4]));
        z4 = __dsub_rn(__dmul_rn(x, c.b[5]), __dmul_rn(y, c.a[5]));
        return y;
    }
};

// PASS: 0 = forward over ext (input x), 1 = backward over reversed y1.
// FINAL: false = zero-state local pass (writes final state), true = re-run from z_in, write outputs.
template <int DTYPE, int PASS, bool FINAL>
__global__ void __launch_bounds__(kFiltThreads) filt_chunk_kernel(const FiltParams p)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= p.n_chunks) return;
    const int u = find_segment(p.chunk_off, p.n, g);
    const int c = g - p.chunk_off[u];
    const int64_t L = p.in_off[u + 1] - p.in_off[u];
    const int64_t fbase = p.fix_off[u];
    const int64_t Lf = p.fix_off[u + 1] - fbase;
    const int64_t M = Lf + 2 * kPadLen;
    const int64_t ebase = fbase + static_cast<int64_t>(u) * 2 * kPadLen;   // extended-signal offset
    const int64_t j0 = static_cast<int64_t>(c) * p.chunk_len;
    const int64_t j1 = min(M, j0 + p.chunk_len);
    const bool last_chunk = (j1 == M);
    if (!FINAL && last_chunk) return;        // nobody consumes the carry out of the last chunk

    const int64_t xbase = p.in_off[u];
    const double *y1 = p.y1 + ebase;
    Df2t f;
    if (FINAL) {
        // sequential mode: the chunk is the whole utterance and starts from scipy's  zi * x[0]
        const double x0 = (PASS == 0) ? ext_sample<DTYPE>(p.x, xbase, L, Lf, 0) : y1[M - 1];
        f.z0 = __dmul_rn(c_filt.zi[0], x0); f.z1 = __dmul_rn(c_filt.zi[1], x0); f.z2 = __dmul_rn(c_filt.zi[2], x0);
        f.z3 = __dmul_rn(c_filt.zi[3], x0); f.z4 = __dmul_rn(c_filt.zi[4], x0);
    } else {
        f.z0 = f.z1 = f.z2 = f.z3 = f.z4 = 0.0;
    }
    for (int64_t j = j0; j < j1; ++j) {
        double xin;
        if (PASS == 0) xin = ext_sample<DTYPE>(p.x, xbase, L, Lf, j);
        else xin = y1[M - 1 - j];
        const double y = f.step(xin);
        if (FINAL) {
            if (PASS == 0) {
                p.y1_out[ebase + j] = y;
            } else {
                const int64_t nidx = M - 1 - kPadLen - j;      // output sample index
                if (nidx >= 0 && nidx < Lf) {
                    if (p.y) p.y[fbase + nidx] = y;
                    if (p.dith) {
                        double d;
                        if (p.dith_f32) {
                            d = static_cast<double>(mt_a_to_dither_f32(reinterpret_cast<const uint32_t *>(p.dith)[fbase + nidx],
                                                                       static_cast<float>(c_filt.dither_scale)));
                        } else {
                            const double uu = p.dith_raw ? mt_raw_to_double(reinterpret_cast<const uint2 *>(p.dith)[fbase + nidx])
                                                         : p.dith[fbase + nidx];
                            d = __dmul_rn(__dsub_rn(uu, 0.5), c_filt.dither_scale);
                        }
                        const double w = __dadd_rn(__dmul_rn(y, c_filt.wav_scale), d);
                        if (p.wav64) p.wav64[fbase + nidx] = w;
                        const float wf = static_cast<float>(w);
                        if (p.wav) p.wav[fbase + nidx] = wf;
                        if (p.wavp) p.wavp[p.seg_off[u] + kHalfPad + nidx] = wf;
                    }
                }
            }
        }
    }
    if (!FINAL) {
        double *s = p.state + static_cast<int64_t>(g) * 5;
        s[0] = f.z0; s[1] = f.z1; s[2] = f.z2; s[3] = f.z3; s[4] = f.z4;
    }
}

// ---- warp-tiled version of the chunk kernel (the production path) ---------------------------------
// A thread walking its own 256-sample chunk touches a different cache line than its 31 neighbours on
// every load and store (32 L1 wavefronts per instruction).  Here a warp owns 32 CONSECUTIVE chunks of
// one utterance (8192 contiguous samples) and moves them through a shared-memory tile 32 samples at
// a time: global traffic is row-wise and fully coalesced (PCM decode, odd extension, dither combine
// and the f32 scatter happen on that side), the recurrence reads and writes the tile column-wise.
constexpr int kTileW = 32;                   // samples per row and sub-step
constexpr int kTileStride = kTileW + 1;      // doubles per tile row (column reads are conflict-free)
constexpr int kFiltWarps = 4;
constexpr int kSmallRows = 8;                // chunks per tile for small batches (see filtfilt_run)
#ifndef SSFE_FILT_BWD_CTAS
#define SSFE_FILT_BWD_CTAS 4      // CTAs per SM the backward final pass is compiled for (5 spills 120 bytes; A/B in profiles/)
#endif

// The cascade the scan runs (see the file header): 13 fp64 operations per sample, FMAs allowed.
struct Casc {
    double s0, s1, s2, s3, s4;
    __device__ __forceinline__ double step(double x)
    {
        const double *c = c_filt.sec;
        const double y0 = fma(c[0], x, s0);                   // section 0: (b0 + b1 z^-1) / (1 + a1 z^-1)
        s0 = fma(-c[3], y0, c[1] * x);
        const double y1 = fma(c[5], y0, s1);                  // section 1
        s1 = fma(-c[8], y1, fma(c[6], y0, s2));
        s2 = fma(-c[9], y1, c[7] * y0);
        const double y2 = fma(c[10], y1, s3);                 // section 2
        s3 = fma(-c[13], y2, fma(c[11], y1, s4));
        s4 = fma(-c[14], y2, c[12] * y1);
        return y2;
    }
};

template <int DTYPE, int PASS> struct RawType { using T = float; };      // pass 1 reads y1, stored as float
template <> struct RawType<SSFE_I16, 0> { using T = short; };
template <> struct RawType<SSFE_F64, 0> { using T = double; };

template <int DTYPE, int PASS, bool FINAL, int ROWS = 32>
__global__ void __launch_bounds__(kFiltWarps * 32, ((PASS == 0 || !FINAL) ? 5 : SSFE_FILT_BWD_CTAS)) filt_tile_kernel(const FiltParams p, const int *__restrict__ tile_off,
                                                                     int n_tiles, const int *__restrict__ tile_map)
{
    __shared__ double s_tile[kFiltWarps][ROWS * kTileStride];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int tile = blockIdx.x * kFiltWarps + w;
    if (tile >= n_tiles) return;
    double *tl = s_tile[w];
    const int u = tile_map[tile];
    const int sc = tile - tile_off[u];                         // super-chunk (ROWS chunks) inside the utterance
    const int nch = p.chunk_off[u + 1] - p.chunk_off[u];
    const int c = sc * ROWS + lane;                            // this lane's chunk
    const int64_t xbase = p.in_off[u];
    const int64_t fbase = p.fix_off[u];
    // positions inside one utterance fit 32 bits (checked on the host); 64-bit only for the bases
    const int L = static_cast<int>(p.in_off[u + 1] - xbase);
    const int Lf = static_cast<int>(p.fix_off[u + 1] - fbase);
    const int M = Lf + 2 * kPadLen;
    const int64_t ebase = fbase + static_cast<int64_t>(u) * 2 * kPadLen;
    const int jt = sc * ROWS * kChunk;                         // first sample of the super-chunk
    const int rows = min(ROWS, nch - sc * ROWS);               // chunks present in this tile
    const bool have = lane < ROWS && c < nch;
    const int j0 = c * kChunk;
    const bool last_chunk = have && (j0 + kChunk >= M);
    const bool run = have && (FINAL || !last_chunk);           // nobody consumes the carry of the last chunk
    const int64_t g = static_cast<int64_t>(p.chunk_off[u]) + c;

    Casc f;
    f.s0 = f.s1 = f.s2 = f.s3 = f.s4 = 0.0;
    if (FINAL && have) {
        const double *s = p.zin + g * 5;
        f.s0 = s[0]; f.s1 = s[1]; f.s2 = s[2]; f.s3 = s[3]; f.s4 = s[4];
    }
    // The forward result travels to the backward pass as FLOAT (8.4 instead of 16.7 GB written once and read
    // twice for the VCTK-shaped corpus).  The backward pass then filters exactly that rounded signal - both its
    // local and its final pass read the same floats, so chunk-entry states stay consistent to the last bit,
    // which is what the ill-conditioned recurrence needs - and the rounding enters as input noise of
    // <= 1.5e-8 through a filter of gain <= 1.  Measured against scipy (CPU simulation, tests): the result moves
    // by <= 2.6e-7, the size of scipy's own fp64 round-off for this filter and inside the 1e-6 bar.
    const float *y1 = (PASS == 1) ? p.y1f + ebase : nullptr;
    float *y1o = (FINAL && PASS == 0) ? p.y1f_out + ebase : nullptr;
    const uint32_t *dith = (FINAL && PASS == 1 && p.dith && p.dith_f32) ? reinterpret_cast<const uint32_t *>(p.dith) + fbase : nullptr;
    const float dscale = static_cast<float>(c_filt.dither_scale);
    const double *dith_gen = (FINAL && PASS == 1 && p.dith && !p.dith_f32) ? p.dith + fbase : nullptr;
    float *wavp = (FINAL && PASS == 1 && p.wavp) ? p.wavp + p.seg_off[u] + kHalfPad : nullptr;
    // ---- load machinery -------------------------------------------------------------------------------
    // (For 16-bit input a tile is one contiguous 16 KB run; fetching it with a single TMA bulk copy into
    // shared memory and cutting the sub-tiles from there was built and measured: correct, but the 16.5 KB
    // of staging per warp leave 8 warps per SM instead of 16 and the forward passes took 5 ms LONGER.)
    // The kernel is bound by global-memory latency.  (1) On the common path (full tile, every row
    // segment inside the signal) the 32 raw values of a sub-tile are fetched branch-free into registers
    // and only later converted and stored - a conversion placed right behind its load makes the
    // in-order warp wait for every load in turn.  (2) The fetch for sub-tile s+1 is issued BEFORE the
    // recurrence of sub-tile s runs, so the ~1-2 us of DRAM latency hide behind ~2000 cycles of fp64.
    using RawT = typename RawType<DTYPE, PASS>::T;
    RawT raw[ROWS];
    // 2: every row present and every row segment inside the signal; 1: the same for the rows the tile has (an
    // utterance's last tile is ragged: 28 of 32 rows for 3 s - a sixth of all tiles, and taking them down the
    // element-wise path cost 2.5x the instructions of a plain sub-tile: 28 % of the forward final pass); 0: a sub-tile
    // that touches the odd extension, the appended sample or the end of the signal.
    auto is_fast = [&](int sub) -> int {
        const int js = jt + sub * kTileW;
        bool inside;
        if (PASS == 0) inside = (js >= kPadLen) && (js + (rows - 1) * kChunk + kTileW <= kPadLen + L);
        else inside = js + (rows - 1) * kChunk + kTileW <= M;
        // (the backward final pass keeps its ragged tiles on the element-wise path: with the partial branch it ran at
        // 6.29 instead of 5.84 ms - it sits at its 128-register limit - while the forward final pass went 4.14 -> 3.83)
        return !inside ? 0 : (rows == ROWS ? 2 : (PASS == 0 ? 1 : 0));
    };
    // Edge tiles (the first tile of an utterance touches the 18 reflected samples, the last one is
    // ragged) are a third of all tiles for 3 s utterances, so they get the same treatment: every
    // position is mapped to a clamped source index, the 32 loads go out together, and extension,
    // appended sample and validity are applied afterwards from the index arithmetic alone.
    auto src_index = [&](int j) -> int {                 // position in the extended signal -> sample of x'
        int nn = j - kPadLen;
        if (j < kPadLen) nn = kPadLen - j;
        else if (j >= kPadLen + Lf) nn = Lf - 2 - (j - kPadLen - Lf);
        return nn;
    };
    auto issue = [&](int sub, int fast_path) {
        const int js = jt + sub * kTileW;
        if (PASS == 0) {
            if (fast_path == 2) {
                const RawT *xs = static_cast<const RawT *>(p.x) + (xbase + (js - kPadLen) + lane);
#pragma unroll
                for (int r = 0; r < ROWS; ++r) raw[r] = xs[r * kChunk];
            } else if (fast_path == 1) {
                const RawT *xs = static_cast<const RawT *>(p.x) + (xbase + (js - kPadLen) + lane);
#pragma unroll
                for (int r = 0; r < ROWS; ++r) raw[r] = (r < rows) ? xs[r * kChunk] : RawT(0);
            } else {
                const RawT *xs = static_cast<const RawT *>(p.x) + xbase;
#pragma unroll
                for (int r = 0; r < ROWS; ++r) raw[r] = xs[min(max(src_index(js + r * kChunk + lane), 0), L - 1)];
            }
        } else {
            if (fast_path == 2) {
                const float *ys = y1 + (M - 1 - js - lane);          // reversed: row r is kChunk samples earlier
#pragma unroll
                for (int r = 0; r < ROWS; ++r) raw[r] = ys[-r * kChunk];
            } else {
#pragma unroll
                for (int r = 0; r < ROWS; ++r)
                    raw[r] = y1[min(max(M - 1 - (js + r * kChunk + lane), 0), M - 1)];
            }
        }
    };
    // the two samples the odd extension reflects about (x'[0], x'[Lf-1])
    double e0 = 0.0, e1 = 0.0;
    if (PASS == 0) {
        e0 = fixed_sample<DTYPE>(p.x, xbase, L, 0);
        e1 = fixed_sample<DTYPE>(p.x, xbase, L, Lf - 1);
    }
    // The dither words of a sub-tile's store phase are 32 row segments of 128 bytes (lane = row).  Fetched where they
    // are used, they were the backward final pass's largest stall (70 % of its stall samples: the words come straight
    // from HBM, mt_walk_kernel wrote them a pass ago); a prefetch per lane sends them on their way one recurrence earlier.
    auto prefetch_dither = [&](int js) {
        if (!(FINAL && PASS == 1) || !dith || !p.dith_pf || lane >= ROWS) return;
        const int nb = M - 1 - kPadLen - js - lane * kChunk;
        if (nb - 31 < 0 || nb >= Lf) return;
        if (p.dith_pf == 1) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(dith + nb));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(dith + nb - 31));
        } else {
            asm volatile("prefetch.global.L1 [%0];" ::"l"(dith + nb));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(dith + nb - 31));
        }
    };
    issue(0, is_fast(0));
    prefetch_dither(jt);
    for (int sub = 0; sub < kChunk / kTileW; ++sub) {
        const int jsub = jt + sub * kTileW;
        const int fast = is_fast(sub);
        if (fast) {                                  // (absent rows of a ragged tile were fetched as zeros)
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                double v = static_cast<double>(raw[r]);
                if (PASS == 0 && DTYPE == SSFE_I16) v *= (1.0 / 32768.0);
                tl[r * kTileStride + lane] = v;
            }
        } else {
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                const int j = jsub + r * kChunk + lane;
                double v = static_cast<double>(raw[r]);
                if (PASS == 0) {
                    if (DTYPE == SSFE_I16) v *= (1.0 / 32768.0);
                    if (src_index(j) >= L) v = 1e-06;                          // the appended sample
                    if (j < kPadLen) v = __dsub_rn(2.0 * e0, v);
                    else if (j >= kPadLen + Lf) v = __dsub_rn(2.0 * e1, v);
                }
                if (r >= rows || j >= M) v = 0.0;
                tl[r * kTileStride + lane] = v;
            }
        }
        if (sub + 1 < kChunk / kTileW) issue(sub + 1, is_fast(sub + 1));       // in flight during the recurrence
        if (sub + 1 < kChunk / kTileW) prefetch_dither(jsub + kTileW);          // for the NEXT sub-tile's store phase
        __syncwarp();
        // ---- recurrence: this lane's chunk is row `lane` -------------------------------------------
        if (run) {
            const int left = M - (j0 + sub * kTileW);                        // may be <= 0 past the end
            const int cnt = left < kTileW ? left : kTileW;
            double *row = tl + lane * kTileStride;
            if (FINAL) {
#pragma unroll 4
                for (int i = 0; i < cnt; ++i) row[i] = f.step(row[i]);
            } else {
#pragma unroll 4
                for (int i = 0; i < cnt; ++i) f.step(row[i]);
            }
        }
        __syncwarp();
        // ---- store (final passes): row-wise again ---------------------------------------------------
        if (FINAL) {
            int fast_out = fast;
            if (PASS == 1) {
                // all row segments map to output samples (none in the 18-sample pads)
                const int n_hi = M - 1 - kPadLen - jsub, n_lo = n_hi - (rows - 1) * kChunk - (kTileW - 1);
                if (!(n_lo >= 0 && n_hi < Lf && dith && wavp && !p.y && !p.wav && !p.wav64)) fast_out = 0;
            }
            if (fast_out == 2 && PASS == 0) {
#pragma unroll
                for (int r = 0; r < ROWS; ++r) y1o[jsub + r * kChunk + lane] = static_cast<float>(tl[r * kTileStride + lane]);
            } else if (fast_out == 1 && PASS == 0) {
#pragma unroll
                for (int r = 0; r < ROWS; ++r)
                    if (r < rows) y1o[jsub + r * kChunk + lane] = static_cast<float>(tl[r * kTileStride + lane]);
            } else if (fast_out) {
                // production path of the backward pass: wav = y * 0.96 + (U - 0.5) * 1e-06 -> f32 segment; the
                // dither arrives as one raw generator word per sample (mt_walk_kernel<true>), tempering and the
                // float conversion happen here, where issue slots are idle (the kernel waits for memory).  These
                // loads were the kernel's largest stall when they were 8-byte word pairs (55 % of the samples);
                // staging them with cp.async (8 KB of shared memory per warp) during the recurrence was measured
                // then: 3 CTAs per SM instead of 4, and the pass got 4.5 ms SLOWER on the full corpus.
                const int n0 = M - 1 - kPadLen - jsub - lane;
                const uint32_t *dr = dith + n0;
                uint32_t dv[ROWS];
#pragma unroll
                for (int r = 0; r < ROWS; ++r) dv[r] = dr[-r * kChunk];
#pragma unroll
                for (int r = 0; r < ROWS; ++r) {
                    const double y = tl[r * kTileStride + lane];
                    wavp[n0 - r * kChunk] = static_cast<float>(
                        __dadd_rn(__dmul_rn(y, c_filt.wav_scale), static_cast<double>(mt_a_to_dither_f32(dv[r], dscale))));
                }
            } else if (PASS == 1 && dith && wavp && !p.y && !p.wav && !p.wav64) {
                // edge tile of the production path: same combine, dither words fetched together
                uint32_t dv[ROWS];
#pragma unroll
                for (int r = 0; r < ROWS; ++r) {
                    const int nidx = M - 1 - kPadLen - (jsub + r * kChunk + lane);
                    dv[r] = dith[min(max(nidx, 0), Lf - 1)];
                }
#pragma unroll
                for (int r = 0; r < ROWS; ++r) {
                    const int j = jsub + r * kChunk + lane;
                    const int nidx = M - 1 - kPadLen - j;
                    if (r < rows && j < M && nidx >= 0 && nidx < Lf) {
                        const double y = tl[r * kTileStride + lane];
                        wavp[nidx] = static_cast<float>(
                            __dadd_rn(__dmul_rn(y, c_filt.wav_scale), static_cast<double>(mt_a_to_dither_f32(dv[r], dscale))));
                    }
                }
            } else {
                for (int r = 0; r < rows; ++r) {
                    const int j = jsub + r * kChunk + lane;
                    if (j >= M) continue;
                    const double y = tl[r * kTileStride + lane];
                    if (PASS == 0) {
                        y1o[j] = static_cast<float>(y);
                    } else {
                        const int nidx = M - 1 - kPadLen - j;
                        if (nidx >= 0 && nidx < Lf) {
                            if (p.y) p.y[fbase + nidx] = y;
                            if (dith || dith_gen) {
                                double d;
                                if (dith) {
                                    d = static_cast<double>(mt_a_to_dither_f32(dith[nidx], dscale));
                                } else {
                                    const double uu = p.dith_raw ? mt_raw_to_double(reinterpret_cast<const uint2 *>(dith_gen)[nidx]) : dith_gen[nidx];
                                    d = __dmul_rn(__dsub_rn(uu, 0.5), c_filt.dither_scale);
                                }
                                const double wv = __dadd_rn(__dmul_rn(y, c_filt.wav_scale), d);
                                if (p.wav64) p.wav64[fbase + nidx] = wv;
                                const float wf = static_cast<float>(wv);
                                if (p.wav) p.wav[fbase + nidx] = wf;
                                if (wavp) wavp[nidx] = wf;
                            }
                        }
                    }
                }
            }
            __syncwarp();
        }
    }
    if (!FINAL && run) {
        double *s = p.state + g * 5;
        s[0] = f.s0; s[1] = f.s1; s[2] = f.s2; s[3] = f.s3; s[4] = f.s4;
    }
}

// ---- local passes as dot products ----------------------------------------------------------------------
// The local passes need only the FINAL state of every chunk entered with a zero state, and that is linear in the
// chunk's samples:  s[k] = sum_i g[k][i] x[i]  with  g[k][i] = (A_c^(255-i) B_c)[k]  (filt_consts.cpp, 113-bit
// algebra, |g| <= ~1).  So instead of walking the recurrence through a shared-memory transpose (13 dependent
// fp64 operations per sample, the tile kernel above with FINAL = false: 3.3 + 2.3 ms on the full corpus) a warp
// forms five dot products per chunk - lane l owns samples l + 32 q, its 40 taps stay in registers for the whole
// kernel - 5 FMAs per sample, no dependent chain.  The partial sums of FOUR chunks (20 values per lane) are summed
// over the lanes by two select-free halving steps and three butterflies (30 shuffle rounds instead of 100).  The
// summation order is fixed by the lane layout alone, so a chunk's state does not depend on the batch it is in.
// What is left is a streaming read, and the samples are staged by TMA: a group of four chunks is ONE contiguous run
// (2 KB of PCM, 4 KB of y1), fetched by a single bulk copy from the 16-byte line its first sample lies in into one
// of the warp's three stage buffers; every warp is its own producer (lane 0, two groups ahead, across tile and
// utterance boundaries) and consumer, no block-level barrier.  (The first version took the samples with 32
// plain loads per lane and group: 2 KB in flight per warp only between the reductions - 2.4 / 2.1 ms, 1.7 TB/s.)
// Groups that touch the odd extension, the appended sample or the ragged end of an utterance (two of ~47 for a 3 s
// utterance) take their samples element by element.
// A bulk copy starts at the 16-byte line of the group's first sample and ends with the line of its last one: up to
// 14 bytes before and after the run.  Both lines hold samples of the same buffer (the caller's PCM array, whose
// first utterance's first group is never a TMA group; the y1 workspace, allocated with 16 bytes to spare), so the
// copy never leaves the allocation it reads.
__device__ double g_filt_taps[5 * kChunk];      // [k][i], see ssfe_filt_cascade_taps
constexpr int kDotWarps = 4;
constexpr int kDotGroup = 4;                    // chunks per group / reduction
constexpr int kDotStages = 3;
constexpr int kDotElems = kDotGroup * kChunk;   // samples per group

template <typename RawT> constexpr int dot_stage_bytes() { return (kDotElems * static_cast<int>(sizeof(RawT)) + 16 + 127) / 128 * 128; }

struct DotDesc {          // one per stage, written by the producer step, read by the consumer step of the same warp
    long long state_idx;  // first of the group's 20 state values
    int tile, r0, cnt, fast, off;
    int pad;
};

struct DotTile {          // warp-uniform geometry of a tile (32 chunks of one utterance)
    int u, sc, c_first, n_run, L, Lf, M;
    long long xbase, ebase;
    __device__ __forceinline__ void load(const FiltParams &p, const int *__restrict__ tile_off, const int *__restrict__ tile_map, int tile)
    {
        u = tile_map[tile];
        sc = tile - tile_off[u];
        c_first = p.chunk_off[u];
        const int nch = p.chunk_off[u + 1] - c_first;
        xbase = p.in_off[u];
        const long long fbase = p.fix_off[u];
        L = static_cast<int>(p.in_off[u + 1] - xbase);
        Lf = static_cast<int>(p.fix_off[u + 1] - fbase);
        M = Lf + 2 * kPadLen;
        ebase = fbase + static_cast<long long>(u) * 2 * kPadLen;
        n_run = min(p.rows, nch - 1 - sc * p.rows);   // nobody consumes the carry out of the utterance's last chunk
    }
};

template <int DTYPE, int PASS>
__global__ void __launch_bounds__(kDotWarps * 32) filt_dot_kernel(const FiltParams p, const int *__restrict__ tile_off,
                                                                  int n_tiles, const int *__restrict__ tile_map)
{
    using RawT = typename RawType<DTYPE, PASS>::T;
    constexpr bool kTma = sizeof(RawT) <= 4;          // fp64 input (a validation dtype) is read element by element
    constexpr int kStage = dot_stage_bytes<RawT>();
    extern __shared__ __align__(128) unsigned char s_dot[];
    __shared__ __align__(8) uint64_t s_bar[kDotWarps][kDotStages];
    __shared__ DotDesc s_desc[kDotWarps][kDotStages];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int n_warps = gridDim.x * kDotWarps;
    unsigned char *stage0 = s_dot + static_cast<size_t>(w) * kDotStages * kStage;
    uint64_t *bar = s_bar[w];
    DotDesc *desc = s_desc[w];
    if (kTma && lane == 0) {
#pragma unroll
        for (int i = 0; i < kDotStages; ++i) mbar_init(&bar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    double gl[5][8];
#pragma unroll
    for (int k = 0; k < 5; ++k)
#pragma unroll
        for (int q = 0; q < 8; ++q) gl[k][q] = g_filt_taps[k * kChunk + lane + 32 * q];
    // Accumulator slot i of a lane belongs to chunk i ^ cx of the group, cx = lane bits 4:3.  With that, the first two
    // reduction steps are halvings without a single select: a lane keeps slots 0-1 (then slot 0) and adds its partner's
    // slots 2-3 (slot 1), which are the same chunks seen from the other side; the lanes with bits 4:3 = c end up with
    // chunk c, five values summed over four lanes, and three butterfly steps over bits 2:0 finish them.
    const int cx = (lane >> 3) & 3;
    __syncwarp();

    // ---- producer side: the warp's groups in order, (tile, r0) with tile = warp, warp + n_warps, ... ----------------
    DotTile pt;
    int p_tile = blockIdx.x * kDotWarps + w, p_r0 = 0;
    auto settle = [&]() {                             // move to the next tile that has a group at p_r0
        while (p_tile < n_tiles) {
            pt.load(p, tile_off, tile_map, p_tile);
            if (p_r0 < pt.n_run) break;
            p_tile += n_warps;
            p_r0 = 0;
        }
    };
    settle();
    auto produce = [&](int s) {
        DotDesc d;
        d.pad = 0;
        if (p_tile >= n_tiles) {                      // the sentinel: no more groups
            d.state_idx = 0; d.tile = -1; d.r0 = 0; d.cnt = 0; d.fast = 0; d.off = 0;
        } else {
            const int j0 = (pt.sc * p.rows + p_r0) * kChunk;
            d.tile = p_tile;
            d.r0 = p_r0;
            d.cnt = min(kDotGroup, pt.n_run - p_r0);
            d.state_idx = (static_cast<long long>(pt.c_first) + pt.sc * p.rows + p_r0) * 5;
            d.fast = (d.cnt == kDotGroup) && (PASS == 1 || (j0 >= kPadLen && j0 + kDotElems <= kPadLen + pt.L));
            d.off = 0;
            if (kTma && d.fast) {
                // lowest address of the group's run: forward x[xbase + j0 - 18 ...], backward y1[ebase + M - j0 - 1024 ...]
                const RawT *src = (PASS == 0) ? static_cast<const RawT *>(p.x) + (pt.xbase + (j0 - kPadLen))
                                              : reinterpret_cast<const RawT *>(p.y1f) + (pt.ebase + (pt.M - j0 - kDotElems));
                const uintptr_t a = reinterpret_cast<uintptr_t>(src), a0 = a & ~static_cast<uintptr_t>(15);
                const uint32_t bytes = static_cast<uint32_t>(((a + kDotElems * sizeof(RawT) + 15) & ~static_cast<uintptr_t>(15)) - a0);
                d.off = static_cast<int>((a - a0) / sizeof(RawT));
                if (lane == 0) {
                    mbar_expect_tx(&bar[s], bytes);
                    tma_load_1d(stage0 + s * kStage, reinterpret_cast<const void *>(a0), bytes, &bar[s]);
                }
            }
            p_r0 += kDotGroup;
            if (p_r0 >= pt.n_run) {
                p_tile += n_warps;
                p_r0 = 0;
                settle();
            }
        }
        if (lane == 0) desc[s] = d;
    };
#pragma unroll
    for (int s = 0; s < kDotStages - 1; ++s) produce(s);
    __syncwarp();

    uint32_t parity = 0;                              // bit s: phase of stage s's barrier
    for (int it = 0;; ++it) {
        const int s = it % kDotStages;
        produce((it + kDotStages - 1) % kDotStages);  // the stage the previous iteration released
        __syncwarp();
        const DotDesc d = desc[s];
        if (d.tile < 0) break;
        double val[kDotGroup * 5];
#pragma unroll
        for (int i = 0; i < kDotGroup * 5; ++i) val[i] = 0.0;
        if (kTma && d.fast) {
            mbar_wait(&bar[s], (parity >> s) & 1u);
            parity ^= 1u << s;
            const RawT *buf = reinterpret_cast<const RawT *>(stage0 + s * kStage) + d.off;
#pragma unroll
            for (int i = 0; i < kDotGroup; ++i) {
                const int e0 = ((i ^ cx) << 8) + lane;              // kChunk == 256
                const RawT *bi = buf + (PASS == 0 ? e0 : kDotElems - 1 - e0);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const double v = static_cast<double>(bi[PASS == 0 ? 32 * q : -32 * q]);
#pragma unroll
                    for (int k = 0; k < 5; ++k) val[i * 5 + k] = fma(gl[k][q], v, val[i * 5 + k]);
                }
            }
        } else {
            // element-wise groups: every position is mapped to a clamped source index, the 32 loads go out together, and
            // extension, appended sample and validity are applied afterwards from the index arithmetic alone (as in the
            // tile kernel) - taken one dependent load at a time, these two groups of an utterance cost a fifth of the kernel
            DotTile ct;
            ct.load(p, tile_off, tile_map, d.tile);
            const int j0 = (ct.sc * p.rows + d.r0) * kChunk;
            const int L = ct.L, Lf = ct.Lf, M = ct.M;
            auto src_index = [&](int j) -> int {             // position in the extended signal -> sample of x'
                int nn = j - kPadLen;
                if (j < kPadLen) nn = kPadLen - j;
                else if (j >= kPadLen + Lf) nn = Lf - 2 - (j - kPadLen - Lf);
                return nn;
            };
            double e0 = 0.0, e1 = 0.0;
            if (PASS == 0) {
                e0 = fixed_sample<DTYPE>(p.x, ct.xbase, L, 0);
                e1 = fixed_sample<DTYPE>(p.x, ct.xbase, L, Lf - 1);
            }
            RawT raw[kDotGroup][8];
#pragma unroll
            for (int i = 0; i < kDotGroup; ++i)
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int j = j0 + (i ^ cx) * kChunk + lane + 32 * q;
                    if (PASS == 0) raw[i][q] = static_cast<const RawT *>(p.x)[ct.xbase + min(max(src_index(j), 0), L - 1)];
                    else raw[i][q] = reinterpret_cast<const RawT *>(p.y1f)[ct.ebase + min(max(M - 1 - j, 0), M - 1)];
                }
#pragma unroll
            for (int i = 0; i < kDotGroup; ++i)
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int j = j0 + (i ^ cx) * kChunk + lane + 32 * q;
                    double v = static_cast<double>(raw[i][q]);
                    if (PASS == 0) {
                        if (DTYPE == SSFE_I16) v *= (1.0 / 32768.0);
                        if (src_index(j) >= L) v = 1e-06;                          // the appended sample
                        if (j < kPadLen) v = __dsub_rn(2.0 * e0, v);
                        else if (j >= kPadLen + Lf) v = __dsub_rn(2.0 * e1, v);
                    }
                    if ((i ^ cx) >= d.cnt || j >= M) v = 0.0;
#pragma unroll
                    for (int k = 0; k < 5; ++k) val[i * 5 + k] = fma(gl[k][q], v, val[i * 5 + k]);
                }
        }
        __syncwarp();                                 // every lane has read the stage: the next produce may refill it
        static_assert(kDotGroup == 4 && kChunk == 256, "the reduction below is written for four chunks of 256");
#pragma unroll
        for (int i = 0; i < 10; ++i) val[i] += __shfl_xor_sync(0xffffffffu, val[i + 10], 16);
#pragma unroll
        for (int i = 0; i < 5; ++i) val[i] += __shfl_xor_sync(0xffffffffu, val[i + 5], 8);
#pragma unroll
        for (int m = 4; m > 0; m >>= 1)
#pragma unroll
            for (int i = 0; i < 5; ++i) val[i] += __shfl_xor_sync(0xffffffffu, val[i], m);
        if ((lane & 7) == 0 && cx < d.cnt) {
            // the PCM scale of the TMA path (raw int16 values), exact: a power of two; the element-wise path scales its samples
            const double sc = (PASS == 0 && DTYPE == SSFE_I16 && d.fast) ? (1.0 / 32768.0) : 1.0;
            double *dst = p.state + d.state_idx + cx * 5;
#pragma unroll
            for (int k = 0; k < 5; ++k) dst[k] = val[k] * sc;
        }
    }
}

template <int DTYPE, int PASS> constexpr size_t dot_smem()
{
    using RawT = typename RawType<DTYPE, PASS>::T;
    return sizeof(RawT) <= 4 ? static_cast<size_t>(kDotWarps) * kDotStages * dot_stage_bytes<RawT>() : 16;
}

// ---- carry ---------------------------------------------------------------------------------------------
// Turn the zero-state finals into true chunk-entry states: z' = M z + s per chunk, M = A_c^256 of the cascade.
// Its entries are O(10) and the recurrence is well conditioned, so this is plain fp64 (the DF2T realisation
// needed double-double here, five lanes per utterance and a warp-wide scan for long utterances).
// A WARP per utterance, 256 chunks per round, as a three-phase scan of its own: lane l owns the run of chunks
// 8 l .. 8 l + 7; (1) it walks its run from a zero state (8 dependent mat-vecs); (2) the 32 run totals are combined
// by a Kogge-Stone scan with the precomputed powers M^8, M^16, ... M^128 (five steps of five shuffled doubles and a
// mat-vec; the round's entry state rides in lane 0's total); (3) it walks its run again from the true entry state
// and keeps the state before every chunk.  21 dependent mat-vecs per 256 chunks instead of 256: a 3 s utterance's
// 188 chunks take ~3 us instead of 20 (a quarter-warp per utterance walking the chunks eight at a time was the
// previous version: 7 us warm, 20 us in the ncu launch list, 0.22 ms per pass on the full corpus).  Finals and entry
// states cross shared memory so that global traffic is coalesced (a lane's run is 320 contiguous bytes).
constexpr int kCarryWarps = 4, kCarryRun = 8, kCarryPow = 5;
constexpr int kCarryStride = kCarryRun * 5 + 1;       // doubles per lane row in shared memory (odd: conflict-free)

__device__ __forceinline__ void carry_matvec(const double *__restrict__ m, const double (&z)[5], const double (&add)[5], double (&out)[5])
{
#pragma unroll
    for (int r = 0; r < 5; ++r) {
        double acc = add[r];
#pragma unroll
        for (int k = 0; k < 5; ++k) acc = fma(m[r * 5 + k], z[k], acc);
        out[r] = acc;
    }
}

template <int DTYPE, int PASS>
__global__ void __launch_bounds__(kCarryWarps * 32) filt_carry_kernel(const FiltParams p)
{
    __shared__ double s_st[kCarryWarps][32 * kCarryStride];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int u = blockIdx.x * kCarryWarps + w;
    if (u >= p.n) return;
    double *st = s_st[w];
    const int c0 = p.chunk_off[u];
    const int nc = p.chunk_off[u + 1] - c0;
    double x0;
    {
        const int64_t L = p.in_off[u + 1] - p.in_off[u];
        const int64_t fbase = p.fix_off[u];
        const int64_t Lf = p.fix_off[u + 1] - fbase;
        const int64_t M = Lf + 2 * kPadLen;
        const int64_t ebase = fbase + static_cast<int64_t>(u) * 2 * kPadLen;
        if (PASS == 0) x0 = ext_sample<DTYPE>(p.x, p.in_off[u], L, Lf, 0);
        else x0 = static_cast<double>(p.y1f[ebase + M - 1]);
    }
    double ze[5];                                                 // entry state of the round's first chunk
#pragma unroll
    for (int i = 0; i < 5; ++i) ze[i] = c_filt.zic[i] * x0;       // scipy's zi * x[0], in cascade coordinates
    const double *__restrict__ s_in = p.state + static_cast<int64_t>(c0) * 5;
    double *__restrict__ zout = p.zin + static_cast<int64_t>(c0) * 5;
    constexpr int kRound = 32 * kCarryRun, kVals = kCarryRun * 5;
    for (int base = 0; base < nc; base += kRound) {
        // finals of the round's chunks (the utterance's last chunk has none), coalesced -> one row per lane
#pragma unroll 8
        for (int j = 0; j < kVals; ++j) {
            const int e = lane + 32 * j, ch = base + e / 5;
            const double v = (ch + 1 < nc) ? s_in[static_cast<int64_t>(base) * 5 + e] : 0.0;
            st[(e / kVals) * kCarryStride + (e % kVals)] = v;
        }
        __syncwarp();
        double sv[kCarryRun][5];
        const double *row = st + lane * kCarryStride;
#pragma unroll
        for (int i = 0; i < kCarryRun; ++i)
#pragma unroll
            for (int k = 0; k < 5; ++k) sv[i][k] = row[i * 5 + k];
        // (1) my run from a zero state
        double t[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) t[k] = sv[0][k];
#pragma unroll
        for (int i = 1; i < kCarryRun; ++i) {
            double o[5];
            carry_matvec(c_filt.mc, t, sv[i], o);
#pragma unroll
            for (int k = 0; k < 5; ++k) t[k] = o[k];
        }
        if (lane == 0) {                                          // the round's entry state enters through lane 0
            double o[5];
            carry_matvec(c_filt.mp[0], ze, t, o);
#pragma unroll
            for (int k = 0; k < 5; ++k) t[k] = o[k];
        }
        // (2) inclusive scan of the run totals: t_l <- M^(8 d) t_(l-d) + t_l for d = 1, 2, 4, 8, 16
#pragma unroll
        for (int lv = 0; lv < kCarryPow; ++lv) {
            const int d = 1 << lv;
            double up[5], o[5];
#pragma unroll
            for (int k = 0; k < 5; ++k) up[k] = __shfl_up_sync(0xffffffffu, t[k], d);
            carry_matvec(c_filt.mp[lv], up, t, o);
            if (lane >= d) {
#pragma unroll
                for (int k = 0; k < 5; ++k) t[k] = o[k];
            }
        }
        // t of lane l is now the entry state of run l + 1
        double z[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const double prev = __shfl_up_sync(0xffffffffu, t[k], 1);
            z[k] = (lane == 0) ? ze[k] : prev;
            ze[k] = __shfl_sync(0xffffffffu, t[k], 31);           // entry state of the next round
        }
        __syncwarp();                                             // every lane has taken its finals out of its row
        // (3) my run from its true entry state; the row now collects the entry states
        double *orow = st + lane * kCarryStride;
#pragma unroll
        for (int i = 0; i < kCarryRun; ++i) {
#pragma unroll
            for (int k = 0; k < 5; ++k) orow[i * 5 + k] = z[k];
            if (i + 1 < kCarryRun) {
                double o[5];
                carry_matvec(c_filt.mc, z, sv[i], o);
#pragma unroll
                for (int k = 0; k < 5; ++k) z[k] = o[k];
            }
        }
        __syncwarp();
#pragma unroll 8
        for (int j = 0; j < kVals; ++j) {
            const int e = lane + 32 * j, ch = base + e / 5;
            if (ch < nc) zout[static_cast<int64_t>(base) * 5 + e] = st[(e / kVals) * kCarryStride + (e % kVals)];
        }
        __syncwarp();                                             // the rows are refilled by the next round
    }
}

// reflect edges of the padded segments once the interior is written (np.pad 'reflect', utils.py:20)
__global__ void reflect_edges_kernel(float *__restrict__ wavp, const int64_t *__restrict__ seg_off,
                                     const int64_t *__restrict__ fix_off, int n)
{
    const int u = blockIdx.x;
    if (u >= n) return;
    const int64_t Lf = fix_off[u + 1] - fix_off[u];
    float *seg = wavp + seg_off[u];
    const int64_t period = (Lf > 1) ? 2 * (Lf - 1) : 1;
    for (int t = threadIdx.x; t < 2 * kHalfPad; t += blockDim.x) {
        const int64_t j = (t < kHalfPad) ? t : Lf + t;          // position in the padded segment
        int64_t m = j - kHalfPad;
        m %= period;
        if (m < 0) m += period;
        if (m >= Lf) m = period - m;
        seg[j] = seg[kHalfPad + m];
    }
}

int fill_reflect_edges(ssfe_ctx *ctx, float *wavp, const int64_t *seg_off_dev, const int64_t *fix_off_dev, int n)
{
    if (n == 0) return SSFE_OK;
    reflect_edges_kernel<<<n, 256, 0, ctx->stream>>>(wavp, seg_off_dev, fix_off_dev, n);
    SSFE_LAUNCHED(ctx);
    return SSFE_OK;
}

template <int DTYPE>
static int filtfilt_typed(ssfe_ctx *ctx, FiltParams p, bool sequential, cudaEvent_t dith_ready,
                          const int *tile_off, int n_tiles, const int *tile_map)
{
    const unsigned gc = (p.n_chunks + kFiltThreads - 1) / kFiltThreads;
    const unsigned gt = (n_tiles + kFiltWarps - 1) / kFiltWarps;
    const unsigned gu = (p.n + kCarryWarps - 1) / kCarryWarps;
    cudaStream_t st = ctx->stream;
    if (sequential) {
        filt_chunk_kernel<DTYPE, 0, true><<<gc, kFiltThreads, 0, st>>>(p);
        SSFE_LAUNCHED(ctx);
        p.y1 = p.y1_out;
        if (dith_ready) SSFE_CUDA(ctx, cudaStreamWaitEvent(st, dith_ready, 0));
        filt_chunk_kernel<DTYPE, 1, true><<<gc, kFiltThreads, 0, st>>>(p);
        SSFE_LAUNCHED(ctx);
        if (dith_ready) SSFE_CUDA(ctx, cudaEventRecord(ctx->ev_dith_free, st));
        return SSFE_OK;
    }
    // local passes: grid-stride over the tiles, a few tiles per warp so that the 40 tap loads amortise
    const unsigned gd = static_cast<unsigned>(std::max(1, std::min((n_tiles + kDotWarps - 1) / kDotWarps, ctx->num_sms * 3)));
    if (ctx->cfg.filtfilt_mode == 3)                                   // the recurrence-walking local pass (A/B, validation)
        filt_tile_kernel<DTYPE, 0, false><<<gt, kFiltWarps * 32, 0, st>>>(p, tile_off, n_tiles, tile_map);
    else
        filt_dot_kernel<DTYPE, 0><<<gd, kDotWarps * 32, dot_smem<DTYPE, 0>(), st>>>(p, tile_off, n_tiles, tile_map);
    SSFE_LAUNCHED(ctx);
    filt_carry_kernel<DTYPE, 0><<<gu, kCarryWarps * 32, 0, st>>>(p);
    SSFE_LAUNCHED(ctx);
    if (p.rows == kSmallRows)
        filt_tile_kernel<DTYPE, 0, true, kSmallRows><<<gt, kFiltWarps * 32, 0, st>>>(p, tile_off, n_tiles, tile_map);
    else
        filt_tile_kernel<DTYPE, 0, true><<<gt, kFiltWarps * 32, 0, st>>>(p, tile_off, n_tiles, tile_map);
    SSFE_LAUNCHED(ctx);
    p.y1f = p.y1f_out;
    if (ctx->cfg.filtfilt_mode == 3)
        filt_tile_kernel<DTYPE, 1, false><<<gt, kFiltWarps * 32, 0, st>>>(p, tile_off, n_tiles, tile_map);
    else
        filt_dot_kernel<DTYPE, 1><<<gd, kDotWarps * 32, dot_smem<DTYPE, 1>(), st>>>(p, tile_off, n_tiles, tile_map);
    SSFE_LAUNCHED(ctx);
    filt_carry_kernel<DTYPE, 1><<<gu, kCarryWarps * 32, 0, st>>>(p);
    SSFE_LAUNCHED(ctx);
    if (dith_ready) SSFE_CUDA(ctx, cudaStreamWaitEvent(st, dith_ready, 0));   // join the dither stream
    if (p.rows == kSmallRows)
        filt_tile_kernel<DTYPE, 1, true, kSmallRows><<<gt, kFiltWarps * 32, 0, st>>>(p, tile_off, n_tiles, tile_map);
    else
        filt_tile_kernel<DTYPE, 1, true><<<gt, kFiltWarps * 32, 0, st>>>(p, tile_off, n_tiles, tile_map);
    SSFE_LAUNCHED(ctx);
    if (dith_ready) SSFE_CUDA(ctx, cudaEventRecord(ctx->ev_dith_free, st));   // the dither buffer may be refilled
    return SSFE_OK;
}

int filtfilt_run(ssfe_ctx *ctx, const void *x_dev, int dtype, const int64_t *in_off_host,
                 const int64_t *fix_off_host, int n, const FiltOut &out)
{
    if (n == 0) return SSFE_OK;
    const bool sequential = ctx->cfg.filtfilt_mode == 1;
    std::vector<int> chunk_off(n + 1), tile_off(n + 1);
    // Chunks per tile.  A warp walks its tile's chunks side by side, one lane each, and a chunk is 8 sub-tiles of
    // load / recurrence / store in a row whatever the tile holds - for a batch that cannot fill the GPU (one 3 s
    // utterance is 6 tiles of 32) tiles of 8 chunks spread the rows over four times the warps: the load and store
    // phases shrink fourfold, the recurrence phase is a latency chain either way.
    int64_t est_tiles = 0;
    for (int i = 0; i < n && !sequential; ++i)
        est_tiles += ((fix_off_host[i + 1] - fix_off_host[i] + 2 * kPadLen + kChunk - 1) / kChunk + 31) / 32;
    const int rows = (!sequential && ctx->cfg.filtfilt_mode != 3 && est_tiles < static_cast<int64_t>(ctx->num_sms) * 4) ? kSmallRows : 32;
    int64_t chunks = 0, max_m = 0, tiles = 0;
    for (int i = 0; i < n; ++i) {
        const int64_t Lf = fix_off_host[i + 1] - fix_off_host[i];
        if (Lf <= kPadLen)
            return set_error(ctx, SSFE_ERR_TOO_SHORT,
                             "utterance %d: the length of the input vector x must be greater than padlen, which is 18", i);
        const int64_t M = Lf + 2 * kPadLen;
        if (M > 0x3fffffff) return set_error(ctx, SSFE_ERR_INVALID, "utterance %d is too long (%lld samples)", i, (long long)Lf);
        max_m = std::max(max_m, M);
        chunk_off[i] = static_cast<int>(chunks);
        tile_off[i] = static_cast<int>(tiles);
        const int64_t nc = sequential ? 1 : (M + kChunk - 1) / kChunk;
        chunks += nc;
        tiles += (nc + rows - 1) / rows;
        if (chunks > 0x7fffffff) return set_error(ctx, SSFE_ERR_INVALID, "batch too large (chunks)");
    }
    chunk_off[n] = static_cast<int>(chunks);
    tile_off[n] = static_cast<int>(tiles);
    const int64_t ext_total = fix_off_host[n] + static_cast<int64_t>(n) * 2 * kPadLen;
    // (+16: filt_dot_kernel's bulk copies run to the end of the 16-byte line their last sample lies in)
    int rc = ensure(ctx, ctx->ws.y1, ext_total * (sequential ? sizeof(double) : sizeof(float)) + 16);
    if (rc) return rc;
    rc = ensure(ctx, ctx->ws.carry, 2 * chunks * 5 * sizeof(double));
    if (rc) return rc;

    FiltParams p;
    memset(&p, 0, sizeof(p));
    p.x = x_dev;
    p.y1 = nullptr;
    p.in_off = upload(ctx, in_off_host, n + 1);
    p.fix_off = upload(ctx, fix_off_host, n + 1);
    p.chunk_off = upload(ctx, chunk_off.data(), n + 1);
    const int *d_tile_off = upload(ctx, tile_off.data(), n + 1);
    if (!p.in_off || !p.fix_off || !p.chunk_off || !d_tile_off) return SSFE_ERR_NOMEM;
    if (int rcf = flush_meta(ctx)) return rcf;
    const int n_tiles = static_cast<int>(tiles);
    {
        const int rc_map = ensure(ctx, ctx->ws.filt_map, (tiles + 1) * sizeof(int));
        if (rc_map) return rc_map;
    }
    int *d_tile_map = static_cast<int *>(ctx->ws.filt_map.p);
    if (!sequential) {
        segment_map_kernel<int><<<static_cast<unsigned>((n + 255) / 256), 256, 0, ctx->stream>>>(d_tile_off, n, 1, d_tile_map);
        SSFE_LAUNCHED(ctx);
    }
    p.n = n;
    p.n_chunks = static_cast<int>(chunks);
    p.chunk_len = sequential ? static_cast<int>(std::min<int64_t>(max_m, 0x7fffffff)) : kChunk;
    p.state = static_cast<double *>(ctx->ws.carry.p);
    p.zin = p.state + chunks * 5;
    p.y = out.y;
    p.dith = out.dith;
    p.dith_raw = out.dith_raw ? 1 : 0;
    p.dith_f32 = out.dith_f32 ? 1 : 0;
    p.rows = rows;
    p.dith_pf = 1;      // measured on the full corpus: backward final pass 6.00 ms without, 5.82 ms into L2, 5.82 ms into L1
    p.wavp = out.wavp;
    p.seg_off = out.seg_off_dev;
    p.wav = out.wav;
    p.wav64 = out.wav64;
    p.y1_out = sequential ? static_cast<double *>(ctx->ws.y1.p) : nullptr;
    p.y1f_out = sequential ? nullptr : static_cast<float *>(ctx->ws.y1.p);
    switch (dtype) {
    case SSFE_F32: return filtfilt_typed<SSFE_F32>(ctx, p, sequential, out.dith_ready, d_tile_off, n_tiles, d_tile_map);
    case SSFE_F64: return filtfilt_typed<SSFE_F64>(ctx, p, sequential, out.dith_ready, d_tile_off, n_tiles, d_tile_map);
    case SSFE_I16: return filtfilt_typed<SSFE_I16>(ctx, p, sequential, out.dith_ready, d_tile_off, n_tiles, d_tile_map);
    default: return set_error(ctx, SSFE_ERR_INVALID, "filtfilt: unknown dtype %d", dtype);
    }
}

int init_filtfilt(ssfe_ctx *ctx)
{
    FiltConsts c;
    memset(&c, 0, sizeof(c));
    for (int i = 0; i < 6; ++i) {
        c.b[i] = ctx->cfg.b[i] / ctx->cfg.a[0];      // scipy normalises by a[0] (== 1 here)
        c.a[i] = ctx->cfg.a[i] / ctx->cfg.a[0];
    }
    for (int i = 0; i < 5; ++i) c.zi[i] = ctx->cfg.zi[i];
    {
        const int rc = ssfe_filt_cascade(c.b, c.a, c.zi, kChunk, c.sec, c.zic, c.mc);
        if (rc != 0)
            return set_error(ctx, SSFE_ERR_INVALID,
                             "filtfilt: the filter (b, a) does not factor into one first-order and two second-order stable "
                             "sections (code %d); only filtfilt_mode = 1 could run it", rc);
    }
    c.wav_scale = ctx->cfg.wav_scale;
    c.dither_scale = ctx->cfg.dither_scale;
    static_assert(kCarryPow == 5 && kCarryRun == 8, "FiltConsts::mp holds five powers of the eight-chunk carry");
    if (ssfe_filt_cascade_powers(c.sec, kChunk, kCarryRun, kCarryPow, &c.mp[0][0]) != 0)
        return set_error(ctx, SSFE_ERR_INVALID, "filtfilt: powers of the carry matrix could not be formed");
    SSFE_CUDA(ctx, cudaMemcpyToSymbol(c_filt, &c, sizeof(c)));
    {
        std::vector<double> g(5 * kChunk);
        if (ssfe_filt_cascade_taps(c.sec, kChunk, g.data()) != 0)
            return set_error(ctx, SSFE_ERR_INVALID, "filtfilt: taps of the local passes could not be formed");
        SSFE_CUDA(ctx, cudaMemcpyToSymbol(g_filt_taps, g.data(), g.size() * sizeof(double)));
        SSFE_CUDA(ctx, cudaFuncSetAttribute(filt_dot_kernel<SSFE_F32, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(dot_smem<SSFE_F32, 0>())));
        SSFE_CUDA(ctx, cudaFuncSetAttribute(filt_dot_kernel<SSFE_F64, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(dot_smem<SSFE_F64, 0>())));
        SSFE_CUDA(ctx, cudaFuncSetAttribute(filt_dot_kernel<SSFE_I16, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(dot_smem<SSFE_I16, 0>())));
        SSFE_CUDA(ctx, cudaFuncSetAttribute(filt_dot_kernel<SSFE_F32, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(dot_smem<SSFE_F32, 1>())));
        SSFE_CUDA(ctx, cudaFuncSetAttribute(filt_dot_kernel<SSFE_F64, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(dot_smem<SSFE_F64, 1>())));
        SSFE_CUDA(ctx, cudaFuncSetAttribute(filt_dot_kernel<SSFE_I16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(dot_smem<SSFE_I16, 1>())));
    }
    return SSFE_OK;
}

void free_filtfilt(ssfe_ctx *ctx)
{
    (void)ctx;
}

}  // namespace ssfe
