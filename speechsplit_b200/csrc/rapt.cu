// rapt.cu - RAPT pitch tracking (stage a6) for ragged batches.
//
// Replaces reference make_spect_f0.py:64
//     f0_rapt = pysptk.sptk.rapt(wav.astype(np.float32)*32768, 16000, 256, min=lo, max=hi, otype=2)
// i.e. SPTK's Snack/ESPS get_f0 (D. Talkin's RAPT) with the parameters rapt() installs.  pysptk is
// a third-party dependency that is not vendored in the reference; the algorithm is restated from
// its published description (see oracle/rapt_ref.c, which is this file's CPU checker).
//
// The original is a streaming C program (0.2 s reads, float accumulators, static buffers).  Voicing
// decisions are discontinuous in the arithmetic, so this file keeps its evaluation ORDER - every
// running sum is formed by one thread, left to right, in the original's precision mix; the file is
// compiled with --fmad=false - and re-expresses the streaming structure in closed form per frame:
//   K1 rapt_decimate   thread per 2 kHz sample: 81-tap symmetric FIR, every 8th output        (A1)
//   K2 rapt_cand       warp per frame: coarse NCCF on the 2 kHz stream, peak picking, parabolic
//                      refinement, fine NCCF (+-3 lags around each coarse peak), peak picking,
//                      local costs and the refined F0 each candidate would emit          (A2-A5)
//   K3 rapt_stat       warp per frame: two 30 ms LPC analyses (order 18, Hanning, pre-emphasis),
//                      Itakura distance -> stationarity, rms ratio                          (A6)
//   K4 rapt_dp         warp per utterance: Viterbi over <= 20 candidates per frame with warp
//                      shuffles (lane = current candidate), the original's commit-on-convergence
//                      back-tracking at every 0.2 s read boundary, log(F0) output          (A7-A8)
// K1-K3 are frame parallel (the bulk of the work); only K4 is sequential in time.
#include "common.cuh"
#include "tma.cuh"
#include <algorithm>
#include <cfloat>
#include <cmath>

namespace ssfe {

constexpr int kCMax = 20;          // n_cands
constexpr int kDec = 8;            // (int)(fs / 2000)
constexpr int kWin = 120;          // round(0.0075 * fs)
constexpr int kNco = 81;           // decimator taps
constexpr int kStatW = 480, kStatGap = 320, kLpcOrd = 18;
constexpr int kRing = 128;         // back-pointer history per utterance (>= DP_LIMIT + one read)
constexpr int kCcMax = 288;        // nlags (<= 257) + slack

struct RaptCfg {
    int start, stop, nlags, ncomp, pad, F, buff_size, sdstep;
    int decstart, decnlags, decsize, n_el;
    float lagwt, lag_wt;
    int table_off, pad0;
};

struct RaptConsts {
    RaptCfg cfg[2];
    float co[kNco];
    float cand_thresh, tcost, tfact_a, tfact_s, vbias, ffact, preemp;
    float ln2, fdouble, freqwt;
    float one;                       // 1.0f the compiler cannot see (f2add below)
    int size_frame_hist, size_frame_out;
};

struct RaptUtt {
    long long wav_off, out_off, fr_off, ds_off;
    int L, cfg, n_out, n_fr, n_ds, R_last, nl, pad;
};

struct RaptTables {
    float *d_ferr = nullptr;       // voiced->voiced transition cost [cfg][lag1][lag2]
    double *d_log = nullptr;       // log(lag), lag < kMaxLag
    int use_log = 0;
    float *d_w479 = nullptr, *d_w480 = nullptr;
    RaptConsts host;
    // where the last rapt_run left its per-frame records (ssfe_rapt_dump)
    long long last_fr = 0;
    const unsigned char *last_ncand = nullptr;
    const short *last_loc = nullptr;
    const float *last_mp = nullptr, *last_f0c = nullptr, *last_sta = nullptr, *last_rr = nullptr;
};

__constant__ RaptConsts c_rapt;

struct RaptParams {
    const float *wav;
    const RaptUtt *utts;
    const long long *fr_offs;      // [n+1]
    const long long *ds_offs;      // [n+1]
    int n;
    long long total_fr, total_ds;
    float *ds;
    // per analysed frame
    unsigned char *ncand;          // incl. the unvoiced candidate
    short *loc;                    // [fr][20]
    float *mp;                     // [fr][20] local cost
    float *f0c;                    // [fr][20] F0 (Hz) the candidate would emit, 0 = unvoiced
    float *sta, *rr;
    const float *ferr;
    const double *log_lag;         // [kMaxLag] log(lag) in double
    int use_log;                   // the log-difference shortcut reproduces the table for every pair
    const float *w479, *w480;
    float *out;                    // log-F0
};

// Packed FP32 (sm_100a: mul / fma.rn.f32x2 -> FMUL2 / FFMA2, one issue slot for two lanes).  This
// file is compiled with --fmad=false because the original rounds every product before it is added;
// the packed forms keep exactly that rounding while halving the instruction count of a multiply-add.
// Pairs live in 64-bit registers (p2) so that they stay packed between uses.
typedef unsigned long long p2;
__device__ __forceinline__ p2 p2pack(float x, float y)
{
    p2 r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(x), "f"(y));
    return r;
}
__device__ __forceinline__ float p2lo(p2 a)
{
    float x;
    asm("{.reg .f32 t; mov.b64 {%0,t}, %1;}" : "=f"(x) : "l"(a));
    return x;
}
__device__ __forceinline__ float p2hi(p2 a)
{
    float y;
    asm("{.reg .f32 t; mov.b64 {t,%0}, %1;}" : "=f"(y) : "l"(a));
    return y;
}
__device__ __forceinline__ p2 p2mul(p2 a, p2 b)
{
    p2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
// a + b as fma(b, one, a) with `one` = 1.0f read from constant memory.  ptxas (12.9) contracts
// mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 even with -fmad=false and explicit rounding modifiers
// (and still does when the 1 is a literal), which would skip the rounding of the product; an fma
// whose multiplier it cannot see is never merged with the multiply that feeds it, and b * 1 + a
// rounds exactly like a + b.  SASS: FMUL2 + FFMA2 R, R.F32x2, UR.F32, R.F32x2.
__device__ __forceinline__ p2 p2add(p2 a, p2 b)
{
    p2 r;
    asm("{.reg .b64 ro; mov.b64 ro, {%3,%3}; fma.rn.f32x2 %0, %2, ro, %1;}" : "=l"(r) : "l"(a), "l"(b), "f"(c_rapt.one));
    return r;
}

// ---- K1 ------------------------------------------------------------------------------------
// 1024 consecutive 2 kHz outputs of one utterance per CTA (n_ds is padded to a multiple of 1024),
// FOUR ADJACENT OUTPUTS PER THREAD.  Output o needs samples 8 o + j (j = 0..80), so tap j of output
// o + 1 is tap j + 8 of output o: a thread walks its 105 samples once and feeds each to up to four
// running sums - 26 shared-memory reads per output instead of 81 (the first version was bound by
// exactly those reads).  The sums of outputs (o, o+1) and (o+2, o+3) advance as packed pairs with the
// coefficient pairs (co[j], co[j-8]) from constant memory; every sum still takes its taps in
// ascending order, product rounded before the add.
// The 8265 input samples are staged once, scaled by 32768, as 8 phase rows (sample i at row i % 8,
// column i / 8), the columns dealt round-robin to four sub-rows of odd pitch: thread t reads column
// 4 t + m, i.e. sub-row m % 4, word t + m / 4 - consecutive threads, consecutive words - and the
// coalesced staging stores hit 32 different banks as well.
constexpr int kDecTile = 1024;
constexpr int kDecThreads = kDecTile / 4;
constexpr int kDecIn = kDec * (kDecTile - 1) + kNco;          // 8265
constexpr int kDecSpan = kNco + 3 * kDec;                     // 105 samples per thread
constexpr int kDecP4 = 261;                                   // words per sub-row (>= 259, odd)
__constant__ float2 c_dec_pair[kDecSpan + 16];                // [i + 16] = (co[i], co[i - 8]), zero outside 0..80

// (tile -> utterance comes from segment_map_kernel: a binary search over the offsets at the start of each of
// these short CTAs was 46 % of the stall samples)
__global__ void __launch_bounds__(kDecThreads) rapt_decimate_kernel(const RaptParams p, const int *__restrict__ tile_map)
{
    __shared__ float s_x[kDec * 4 * kDecP4];
    const long long gid0 = blockIdx.x * static_cast<long long>(kDecTile);
    const int u = tile_map[blockIdx.x];
    const RaptUtt ut = p.utts[u];
    const int m0 = static_cast<int>(gid0 - ut.ds_off);
    const float *x = p.wav + ut.wav_off;
    const int base = kDec * m0 - (kNco / 2);
    // Staging.  Thread t owns samples i = t + 256 k: its phase row (i % 8) and sub-row ((i / 8) % 4) never
    // change, only the word advances by 8 per k.  Interior tiles need no clamping; edge tiles load through
    // clamped indices and select afterwards.  All loads of a round are in flight together.
    const int tid = threadIdx.x;
    float *srow = s_x + ((tid & 7) * 4 + ((tid >> 3) & 3)) * kDecP4 + (tid >> 5);
    constexpr int kRound = 11, kRounds = (kDecIn + kRound * kDecThreads - 1) / (kRound * kDecThreads);   // 3 x 11 x 256
    const bool interior = base >= 0 && base + kRounds * kRound * kDecThreads <= ut.L;
    if (interior) {
        const float *xs = x + base + tid;
#pragma unroll 1
        for (int rd = 0; rd < kRounds; ++rd) {
            float raw[kRound];
#pragma unroll
            for (int q = 0; q < kRound; ++q) raw[q] = xs[(rd * kRound + q) * kDecThreads];
#pragma unroll
            for (int q = 0; q < kRound; ++q)
                if ((rd * kRound + q) * kDecThreads + tid < kDecIn) srow[(rd * kRound + q) * 8] = raw[q] * 32768.0f;
        }
    } else {
#pragma unroll 1
        for (int rd = 0; rd < kRounds; ++rd) {
            float raw[kRound];
#pragma unroll
            for (int q = 0; q < kRound; ++q) {
                const int idx = base + (rd * kRound + q) * kDecThreads + tid;
                raw[q] = x[min(max(idx, 0), ut.L - 1)];
            }
#pragma unroll
            for (int q = 0; q < kRound; ++q) {
                const int i = (rd * kRound + q) * kDecThreads + tid, idx = base + i;
                if (i < kDecIn) srow[(rd * kRound + q) * 8] = (idx >= 0 && idx < ut.L) ? raw[q] * 32768.0f : 0.0f;
            }
        }
    }
    __syncthreads();
    const float *col0 = s_x + threadIdx.x;
    p2 s01 = p2pack(0.0f, 0.0f), s23 = p2pack(0.0f, 0.0f);
#pragma unroll
    for (int i = 0; i < kDecSpan; ++i) {
        const int m = i >> 3;
        const float xv = col0[((i & 7) * 4 + (m & 3)) * kDecP4 + (m >> 2)];
        const p2 xx = p2pack(xv, xv);
        if (i <= kNco - 1 + kDec)                             // outputs o, o+1: taps i, i-8
            s01 = p2add(s01, p2mul(xx, *reinterpret_cast<const p2 *>(&c_dec_pair[i + 16])));
        if (i >= 2 * kDec)                                    // outputs o+2, o+3: taps i-16, i-24
            s23 = p2add(s23, p2mul(xx, *reinterpret_cast<const p2 *>(&c_dec_pair[i])));
    }
    auto rnd = [](float sum) {
        return static_cast<float>((sum < 0.0) ? static_cast<double>(sum) - 0.5 : static_cast<double>(sum) + 0.5);
    };
    *reinterpret_cast<float4 *>(p.ds + gid0 + 4 * threadIdx.x) =
        make_float4(rnd(p2lo(s01)), rnd(p2hi(s01)), rnd(p2lo(s23)), rnd(p2hi(s23)));
}

// ---- warp helpers ----------------------------------------------------------------------------
__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// peaks of correl[0..nl) above cand_thresh*maxval, in ascending lag order (get_cand)
// [i_lo, i_hi): the index range that can hold non-zero correlations (everything outside is zero and,
// with clip >= 0, can never be a peak), so the scan may skip it
template <int CAP>
__device__ __forceinline__ int pick_candidates(const float *cc, int nl, int firstlag, float maxval,
                                               float *peaks, int *locs, int lane, int i_lo = 1, int i_hi = 1 << 30)
{
    const float clip = c_rapt.cand_thresh * maxval;
    const int lastl = min(nl - 2, i_hi);
    int count = 0;
    for (int base = max(1, i_lo); base < lastl; base += 32) {
        const int i = base + lane;
        bool ok = false;
        float q = 0.0f;
        if (i < lastl) {
            q = cc[i];
            ok = (q > clip) && (q >= cc[i + 1]) && (q >= cc[i - 1]);
        }
        const unsigned mask = __ballot_sync(0xffffffffu, ok);
        const int pos = count + __popc(mask & ((1u << lane) - 1u));
        if (ok && pos < CAP) {
            peaks[pos] = q;
            locs[pos] = i + firstlag;
        }
        count += __popc(mask);
    }
    __syncwarp();
    return min(count, CAP);
}

// keep the n_cands-1 largest, by the original's partial bubble pass (order matters downstream)
__device__ __forceinline__ int prune_candidates(float *peaks, int *locs, int ncand, int lane)
{
    if (ncand < kCMax) return ncand;
    if (lane == 0) {
        for (int outer = 0; outer < kCMax - 1; ++outer) {
            int idx = ncand - 1;
            for (int inner = ncand - 1 - outer; inner-- > 0; --idx) {
                const float sm = peaks[idx];
                if (sm > peaks[idx - 1]) {
                    const int lt = locs[idx];
                    peaks[idx] = peaks[idx - 1];
                    peaks[idx - 1] = sm;
                    locs[idx] = locs[idx - 1];
                    locs[idx - 1] = lt;
                }
            }
        }
    }
    __syncwarp();
    return kCMax - 1;
}

// ---- K2 ------------------------------------------------------------------------------------
// A CTA takes 16 consecutive frames of one utterance, a warp four of them (frames w, w + 4, w + 8, w + 12 of
// the tile).  Everything that is a long sequential chain - and RAPT's fine stage is nothing else: per frame a
// 120-term mean, a 120-term reference energy, and per coarse candidate a 120-term window energy plus SEVEN
// 120-term cross products - is advanced for the warp's four frames TOGETHER, one lane per (frame, candidate):
//   * the lane keeps the candidate's 7-lag window  x[st + j .. st + j + 6]  as a sliding window in registers:
//     one new shared-memory value per step feeds all seven cross products and the window energy (the first
//     version read one value per multiply, and ncu had the kernel at 91 % of the shared-memory pipe);
//   * the reference sample ref[j] arrives as a 128-bit read shared by the lanes of a frame;
//   * every chain is still one left-to-right sum by one thread, product rounded before the add (--fmad=false),
//     so the bits are those of the serial original;
//   * the reference energy of a frame is one more item (a pseudo-candidate at lag 0 whose cross products are
//     ignored), so it costs no pass of its own.
// With ~3 coarse candidates per frame a warp's four frames are ~16 items, i.e. one pass of 120 steps; a group of
// frames whose items exceed 32 is split (warp-uniformly) into several passes.
// Shared memory per warp: four mean-free 441-sample windows at a stride of 456 floats (the four frames start
// 8 banks apart, so the per-frame broadcast reads of one instruction never collide).  The coarse stage's scratch
// and, later, the per-frame correlation array and peak lists alias the window of frame 0 (dead by then).
constexpr int kCandWarps = 4;
constexpr int kCandTile = 16;      // frames per CTA (4 per warp)
constexpr int kXfStride = 456;     // >= ncomp (441), = 8 mod 32, multiple of 4
constexpr int kFinePk = 84;        // peaks of the fine correlation: it is non-zero on <= 20 x 7 lags, so <= 70

// float(num / sqrt(prod)) in double, out of line: the double-precision division and square root expand to ~50
// instructions each, and rapt_cand_kernel holds eleven copies of the pair when they are inlined into its unrolled
// normalisation loops (its 59 KB of SASS miss the instruction cache: 12 % of its stall samples are no_instructions)
__device__ __noinline__ float nccf_norm(float num, double prod)
{
    return static_cast<float>(static_cast<double>(num) / sqrt(prod));
}

__global__ void __launch_bounds__(kCandWarps * 32, 6) rapt_cand_kernel(const RaptParams p, const int *__restrict__ tile_off,
                                                                    const int *__restrict__ tile_map)
{
    __shared__ __align__(16) float s_xf[kCandWarps][4 * kXfStride];
    // candidates of the warp's four frames after the coarse stage: [4][20] peaks, [4][20] lags, [4] counts
    __shared__ __align__(16) float s_stash[kCandWarps][4 * kCMax + 4 * kCMax + 4];
    __shared__ int s_fr[kCandWarps][12];                              // per frame of the warp: see the per-frame phase

    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float *xf = s_xf[w];
    // per-frame phase: every frame's correlation array [kCcMax = 288] and peak lists [2 x kFinePk] alias its own window
    static_assert(kCcMax + 2 * kFinePk <= kXfStride, "per-frame buffers must fit the frame's window");

    const int u = tile_map[blockIdx.x];
    const RaptUtt ut = p.utts[u];
    const RaptCfg &cf = c_rapt.cfg[ut.cfg];
    const int g_tile = (static_cast<int>(blockIdx.x) - tile_off[u]) * kCandTile;
    const float *x = p.wav + ut.wav_off;
    float *st_pk = s_stash[w];
    int *st_lc = reinterpret_cast<int *>(s_stash[w] + 4 * kCMax), *st_n = reinterpret_cast<int *>(s_stash[w] + 8 * kCMax);

    // -- coarse stage on the 2 kHz stream, the warp's FOUR frames at once ----------------------------
    // A quarter-warp per frame (frame w + 4 q for lanes 8 q .. 8 q + 7).  The coarse window is only ~57
    // samples and its expensive parts are short sequential chains (mean, energies, the lagged-energy
    // update) that one lane has to walk: with a whole warp per frame they ran on 1 lane of 32, here four
    // frames' chains advance together.  Control flow is uniform across the warp (same configuration for
    // all frames of an utterance), validity is a predicate, so every __syncwarp / ballot is convergent.
    // The working buffers alias the fine-stage windows, which are idle until the staging below.
    {
        const int sub = lane >> 3, l8 = lane & 7;
        const int gq = g_tile + w + kCandWarps * sub;
        const int gc = min(gq, ut.n_fr - 1);                       // frames past the end: computed, not kept
        // (strides of 72 floats / 44 doubles, not 64 / 40: the four quarter-warps then start 8 banks apart
        // instead of on the same bank - the plain layout made every access of this stage a 4-way conflict,
        // a quarter of the kernel's shared-memory wavefronts in ncu)
        float *cdb = xf + 72 * sub;                                // [64] coarse window
        double *cec = reinterpret_cast<double *>(xf + 288) + 44 * sub;   // [40] lagged energy per lag
        float *ccc = xf + 288 + 352 + 40 * sub;                    // [40] coarse correlation
        float *cpk = xf + 288 + 352 + 160 + kCMax * sub;           // [20] peaks
        int *clc = reinterpret_cast<int *>(xf + 288 + 352 + 160 + 4 * kCMax) + kCMax * sub;   // [20] lags
        static_assert(288 + 352 + 160 + 8 * kCMax <= 4 * kXfStride, "coarse scratch must fit the window buffers");
        int r, i, nfr_r;
        const int full = ut.R_last * cf.F;
        if (gc < full) { r = gc / cf.F; i = gc - r * cf.F; nfr_r = cf.F; }
        else { r = ut.R_last; i = gc - full; nfr_r = ut.nl; }
        const bool flushed = (r == ut.R_last) && (r > 0);     // last read of several: the FIR was flushed
        const int samsds = ((nfr_r - 1) * kHop + cf.ncomp) / kDec;
        const float *ds = p.ds + ut.ds_off + static_cast<long long>(r) * (cf.sdstep / kDec);
        for (int t = l8; t < cf.n_el; t += 8) {
            const int q = (kHop / kDec) * i + t;
            float v = 0.0f;
            if (q < samsds) {
                v = ds[q];
            } else if (flushed) {
                // samples the original produced while flushing the filter with zeros: output e of the
                // flush sees its last 8e taps zeroed
                const int e = q - samsds;
                if (e == 0) {
                    v = ds[q];
                } else {
                    const int base = r * cf.sdstep + kDec * q - (kNco / 2);
                    float sum = 0.0f;
                    for (int j = 0; j < kNco - kDec * e; ++j) {
                        const int idx = base + j;
                        const float xv = (idx >= 0 && idx < ut.L) ? x[idx] * 32768.0f : 0.0f;
                        sum += c_rapt.co[j] * xv;
                    }
                    v = static_cast<float>((sum < 0.0) ? static_cast<double>(sum) - 0.5 : static_cast<double>(sum) + 0.5);
                }
            }
            cdb[t] = v;
        }
        __syncwarp();
        const int size = cf.decsize, start = cf.decstart, nlags = cf.decnlags;
        float engr = 0.0f;
        for (int j = 0; j < size; ++j) engr += cdb[j];
        engr /= size;
        __syncwarp();
        for (int t = l8; t < cf.n_el; t += 8) cdb[t] = cdb[t] - engr;
        __syncwarp();
        float sum = 0.0f;
        for (int j = 0; j < size; ++j) { const float st = cdb[j]; sum += st * st; }
        engr = sum;
        const bool pos_e = engr > 0.0f;
        sum = 0.0f;
        for (int j = 0; j < size; ++j) { const float st = cdb[start + j]; sum += st * st; }
        // cross products, one left-to-right chain per lag
        for (int lag = l8; lag < nlags; lag += 8) {
            float dot = 0.0f;
            for (int j = 0; j < size; ++j) dot += cdb[j] * cdb[lag + start + j];
            ccc[lag] = dot;
        }
        // the lagged energy is a running (sequential, double) update: one lane per frame walks it and
        // publishes the values, so that the square root and the division are done by the lane that owns a lag
        // (forming the squares it subtracts and adds beforehand, eight lanes per frame, was measured: 15.14 against
        // 15.00 ms - the extra shared-memory round trip costs more than the four instructions per lag it saves)
        if (l8 == 0) {
            double engc = sum;
            for (int k = 0; k < nlags; ++k) {
                cec[k] = engc;
                const float a0 = cdb[k + start], az = cdb[k + start + size];
                engc -= static_cast<double>(a0 * a0);
                if ((engc += static_cast<double>(az * az)) < 1.0) engc = 1.0;
            }
        }
        __syncwarp();
        float tmax = 0.0f;
        for (int lag = l8; lag < nlags; lag += 8) {
            const float t = pos_e ? nccf_norm(ccc[lag], cec[lag] * engr) : 0.0f;
            ccc[lag] = t;
            tmax = fmaxf(tmax, t);
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
        const float maxval_c = tmax;
        __syncwarp();
        // peaks above cand_thresh * maxval, in ascending lag order (get_cand), eight lags per step
        int ncand_c = 0;
        {
            const float clip = c_rapt.cand_thresh * maxval_c;
            const int lastl = nlags - 2;
            for (int base = 1; base < lastl; base += 8) {
                const int ii = base + l8;
                bool ok = false;
                float q = 0.0f;
                if (ii < lastl) {
                    q = ccc[ii];
                    ok = (q > clip) && (q >= ccc[ii + 1]) && (q >= ccc[ii - 1]);
                }
                const unsigned m8 = (__ballot_sync(0xffffffffu, ok) >> (8 * sub)) & 0xffu;
                const int pos = ncand_c + __popc(m8 & ((1u << l8) - 1u));
                if (ok && pos < kCMax) {
                    cpk[pos] = q;
                    clc[pos] = ii + start;
                }
                ncand_c += __popc(m8);
            }
            ncand_c = min(ncand_c, kCMax);
        }
        __syncwarp();
        // parabolic refinement to full-rate lags
        const float lag_wt = cf.lag_wt;
        for (int c = l8; c < ncand_c; c += 8) {
            const float *y = ccc + (clc[c] - start - 1);
            float xp, yp;
            const float a = static_cast<float>((y[2] - y[1]) + (.5 * (y[0] - y[2])));
            if (fabs(a) > .000001) {
                const float cq = static_cast<float>((y[0] - y[2]) / (4.0 * a));
                xp = cq;
                yp = y[1] - (a * cq * cq);
            } else {
                xp = 0.0f;
                yp = y[1];
            }
            const int l2 = (clc[c] * kDec) + static_cast<int>(0.5 + (xp * kDec));
            clc[c] = l2;
            cpk[c] = yp * (1.0 - (lag_wt * l2));
        }
        __syncwarp();
        // keep the n_cands - 1 largest (never needed for <= 40 coarse lags, kept for completeness)
        if (ncand_c >= kCMax) {
            if (l8 == 0) {
                for (int outer = 0; outer < kCMax - 1; ++outer) {
                    int idx = ncand_c - 1;
                    for (int inner = ncand_c - 1 - outer; inner-- > 0; --idx) {
                        const float sm = cpk[idx];
                        if (sm > cpk[idx - 1]) {
                            const int lt = clc[idx];
                            cpk[idx] = cpk[idx - 1];
                            cpk[idx - 1] = sm;
                            clc[idx] = clc[idx - 1];
                            clc[idx - 1] = lt;
                        }
                    }
                }
            }
            ncand_c = kCMax - 1;
        }
        __syncwarp();
        for (int c = l8; c < ncand_c; c += 8) {
            st_pk[kCMax * sub + c] = cpk[c];
            st_lc[kCMax * sub + c] = clc[c];
        }
        if (l8 == 0) st_n[sub] = (gq < ut.n_fr) ? ncand_c : -1;      // -1: no such frame
        __syncwarp();
    }

    // -- fine stage: stage the four windows, mean-free ------------------------------------------------
    const int start0 = cf.start, nlags0 = cf.nlags, total = cf.ncomp;   // size + nlags0 + start0
    int nfq = 0;                                                   // frames of this warp that exist (a prefix)
#pragma unroll
    for (int f = 0; f < 4; ++f) nfq += (g_tile + w + kCandWarps * f < ut.n_fr) ? 1 : 0;
    if (nfq == 0) return;
    for (int f = 0; f < nfq; ++f) {
        const float *fx = x + static_cast<long long>(g_tile + w + kCandWarps * f) * kHop;
        float raw[14];
#pragma unroll
        for (int q = 0; q < 14; ++q) raw[q] = fx[min(lane + 32 * q, total - 1)];
#pragma unroll
        for (int q = 0; q < 14; ++q)
            if (lane + 32 * q < total) xf[f * kXfStride + lane + 32 * q] = raw[q] * 32768.0f;
    }
    __syncwarp();
    {
        // mean of the reference window: one left-to-right chain per frame, lane f walks frame f
        float mean = 0.0f;
        if (lane < nfq) {
            const float *q = xf + lane * kXfStride;
#pragma unroll 6
            for (int j = 0; j < kWin; j += 4) {
                const float4 v4 = *reinterpret_cast<const float4 *>(q + j);
                mean += v4.x;
                mean += v4.y;
                mean += v4.z;
                mean += v4.w;
            }
            mean /= kWin;
        }
        for (int f = 0; f < nfq; ++f) {
            const float m = __shfl_sync(0xffffffffu, mean, f);
            float *q = xf + f * kXfStride;
            for (int t = lane; t < total; t += 32) q[t] = q[t] - m;
        }
    }
    __syncwarp();

    // -- groups of frames whose items (candidates + one reference item per frame) fit the 32 lanes -------
    for (int f_begin = 0; f_begin < nfq;) {
        int f_end = f_begin, n_items = 0;
        while (f_end < nfq && n_items + st_n[f_end] + 1 <= 32) {
            n_items += st_n[f_end] + 1;
            ++f_end;
        }
        // this lane's item: frame my_f, candidate my_c (== count of the frame: the reference item)
        int my_f = f_begin, my_c = lane, base_f = 0;
        while (my_f < f_end && my_c > st_n[my_f]) {
            my_c -= st_n[my_f] + 1;
            base_f += st_n[my_f] + 1;
            ++my_f;
        }
        const bool active = my_f < f_end;
        if (!active) { my_f = f_begin; my_c = st_n[f_begin]; base_f = 0; }   // idle lanes shadow a reference item
        const bool is_ref = my_c == st_n[my_f];
        int my_st = 0;
        if (!is_ref) {
            my_st = st_lc[kCMax * my_f + my_c] - 3;
            if (my_st < start0) my_st = start0;
        }
        const float *xr = xf + my_f * kXfStride;       // reference window of my frame
        const float *xs = xr + my_st;                  // my candidate's first lagged window
        // The chains are SKEWED so that they share their multiplicand: chain t (lag st + t) takes its j-th term at
        // time tau = j + t, and at time tau every chain multiplies the same sample x[st + tau] - the one value the lane
        // loads per step - by ref[tau - t]:   dot_t = sum_tau ref[tau - t] * x[st + tau],  tau = 0 .. 125, ref = 0
        // outside 0 .. 119 (a term 0 * x adds exactly nothing, and every chain still sums its own products in
        // ascending j, product rounded before the add).  The reference values are the same for all lanes of a frame
        // (broadcast reads), so the chains advance as PACKED pairs (t, t + 1) against (ref[tau - t], ref[tau - t - 1])
        // - FMUL2 + FFMA2 with the hidden 1.0 (p2add) - without any per-lane repacking: three packed pairs, chain 6
        // and the window energy (x^2, tau < 120) scalar.  14 instead of 18 instructions per step; the body is one
        // rotation of the six-deep ring of reference pairs.
        float dot[7], s2 = 0.0f;
        {
            p2 a01 = p2pack(0.0f, 0.0f), a23 = a01, a45 = a01;
            float a6 = 0.0f;
            p2 q[6];                                          // q[(tau - k) % 6] = (ref[tau - k], ref[tau - k - 1])
#pragma unroll
            for (int k = 0; k < 6; ++k) q[k] = a01;
#pragma unroll 1
            for (int tb = 0; tb < 120; tb += 6) {
#pragma unroll
                for (int ph = 0; ph < 6; ++ph) {
                    const int tau = tb + ph;
                    const float x1 = xs[tau];
                    q[ph] = p2pack(xr[tau], (tau > 0) ? xr[tau - 1] : 0.0f);
                    const p2 xx = p2pack(x1, x1);
                    a01 = p2add(a01, p2mul(xx, q[ph]));
                    a23 = p2add(a23, p2mul(xx, q[(ph + 4) % 6]));          // (ref[tau - 2], ref[tau - 3])
                    a45 = p2add(a45, p2mul(xx, q[(ph + 2) % 6]));          // (ref[tau - 4], ref[tau - 5])
                    a6 += x1 * p2hi(q[(ph + 1) % 6]);                      // ref[tau - 6] = second half of the pair of tau - 5
                    s2 += x1 * x1;
                }
            }
#pragma unroll
            for (int ph = 0; ph < 6; ++ph) {                  // tau = 120 .. 125: the reference has run out
                const int tau = 120 + ph;
                const float x1 = xs[tau];
                q[ph] = p2pack(0.0f, (ph == 0) ? xr[119] : 0.0f);
                const p2 xx = p2pack(x1, x1);
                a01 = p2add(a01, p2mul(xx, q[ph]));
                a23 = p2add(a23, p2mul(xx, q[(ph + 4) % 6]));
                a45 = p2add(a45, p2mul(xx, q[(ph + 2) % 6]));
                a6 += x1 * p2hi(q[(ph + 1) % 6]);
            }
            dot[0] = p2lo(a01); dot[1] = p2hi(a01); dot[2] = p2lo(a23); dot[3] = p2hi(a23);
            dot[4] = p2lo(a45); dot[5] = p2hi(a45); dot[6] = a6;
        }
        // reference energy of my frame: from the lane that holds its reference item
        const float engr = __shfl_sync(0xffffffffu, s2, base_f + st_n[my_f]);
        float vmax = 0.0f;
        if (engr > 0.0f && !is_ref) {
            // lagged energies: a short sequential double chain, then square root and division
            double engc = s2;
#pragma unroll
            for (int t = 0; t < 7; ++t) {
                if (engc < 1.0) engc = 1.0;
                const float v = nccf_norm(dot[t], 10000.0 + (engc * engr));
                dot[t] = v;
                vmax = fmaxf(vmax, v);
                const float a0 = xs[t], az = xs[t + kWin];
                engc -= static_cast<double>(a0 * a0);
                engc += static_cast<double>(az * az);
            }
        }
        __syncwarp();          // every lane is done with the windows of this group (frame 0's is overwritten below)

        // -- per frame: correlation array, peaks, pruning, local costs --------------------------------------
        // The group's frames are finished TOGETHER, a quarter-warp per frame (frame f_begin + q for lanes 8 q .. 8 q + 7),
        // each in its own window buffer (dead by now).  Taken one frame at a time by the whole warp, this phase was a
        // fifth of the kernel's instructions - the same ~330 instructions four times over, most lanes idle: a frame
        // has ~3 coarse candidates, ~25 lags to scan and 20 record slots to write.
        {
            const int q = lane >> 3, l8 = lane & 7;
            const int fq = f_begin + q;
            const bool has_f = fq < f_end;
            const int fqc = has_f ? fq : f_begin;
            int *fr = s_fr[w];                                       // [4][3]: min / max first lag, max correlation
            if (l8 == 0) { fr[3 * q] = 1 << 30; fr[3 * q + 1] = -(1 << 30); fr[3 * q + 2] = 0; }
            __syncwarp();
            const bool item = active && !is_ref;
            if (item) {
                const int k = my_f - f_begin;
                atomicMin(&fr[3 * k], my_st);
                atomicMax(&fr[3 * k + 1], my_st);
                atomicMax(&fr[3 * k + 2], static_cast<int>(__float_as_uint(vmax)));   // values are >= 0: integer order
            }
            int rb = 0, cmax = 0;                                    // my frame's first item lane; most candidates of a frame
            for (int f = f_begin; f < f_end; ++f) {
                if (f < fqc) rb += st_n[f] + 1;
                cmax = max(cmax, st_n[f]);
            }
            const int ncand_c = st_n[fqc];
            const float engr_q = __shfl_sync(0xffffffffu, s2, rb + ncand_c);
            __syncwarp();
            // The fine correlation is non-zero only inside the 7-lag windows, so only the span of those
            // windows (plus one zero on either side, which the peak test reads) is cleared and scanned.
            int z_lo = 0, z_hi = 0;
            if (ncand_c > 0) {
                z_lo = max(fr[3 * q] - start0 - 1, 0);
                z_hi = min(fr[3 * q + 1] - start0 + 7 + 1, nlags0);
            }
            const float maxval_r = __uint_as_float(static_cast<unsigned>(fr[3 * q + 2]));
            const bool e_ok = has_f && engr_q > 0.0f;
            float *ccq = xf + fqc * kXfStride;                       // [kCcMax] fine correlation of my frame
            float *pkq = ccq + kCcMax;                               // [kFinePk] peaks
            int *lcq = reinterpret_cast<int *>(ccq + kCcMax + kFinePk);   // [kFinePk] lags
            if (e_ok)
                for (int t = z_lo + l8; t < z_hi; t += 8) ccq[t] = 0.0f;
            __syncwarp();
            // windows are written in candidate order; later ones overwrite earlier ones
            for (int c = 0; c < cmax; ++c) {
                if (item && my_c == c && engr > 0.0f) {
                    float *ccm = xf + my_f * kXfStride;
                    const int o = my_st - start0;
#pragma unroll
                    for (int t = 0; t < 7; ++t)
                        if (o + t < kCcMax) ccm[o + t] = dot[t];
                }
                __syncwarp();
            }
            // peaks above cand_thresh * maxval, in ascending lag order (get_cand), eight lags per step
            const float maxval = e_ok ? maxval_r : 0.0f;
            int ncand = 0;
            {
                const float clip = c_rapt.cand_thresh * maxval;
                const int lastl = min(nlags0 - 2, z_hi - 1);
                for (int base = max(1, z_lo + 1); __any_sync(0xffffffffu, e_ok && base < lastl); base += 8) {
                    const int i = base + l8;
                    bool ok = false;
                    float qv = 0.0f;
                    if (e_ok && i < lastl) {
                        qv = ccq[i];
                        ok = (qv > clip) && (qv >= ccq[i + 1]) && (qv >= ccq[i - 1]);
                    }
                    const unsigned m8 = (__ballot_sync(0xffffffffu, ok) >> (8 * q)) & 0xffu;
                    const int pos = ncand + __popc(m8 & ((1u << l8) - 1u));
                    if (ok && pos < kFinePk) {
                        pkq[pos] = qv;
                        lcq[pos] = i + start0;
                    }
                    ncand += __popc(m8);
                }
                ncand = min(ncand, kFinePk);
            }
            __syncwarp();
            // keep the n_cands - 1 largest, by the original's partial bubble pass (order matters downstream)
            if (__any_sync(0xffffffffu, ncand >= kCMax)) {
                if (ncand >= kCMax && l8 == 0) {
                    for (int outer = 0; outer < kCMax - 1; ++outer) {
                        int idx = ncand - 1;
                        for (int inner = ncand - 1 - outer; inner-- > 0; --idx) {
                            const float sm = pkq[idx];
                            if (sm > pkq[idx - 1]) {
                                const int lt = lcq[idx];
                                pkq[idx] = pkq[idx - 1];
                                pkq[idx - 1] = sm;
                                lcq[idx] = lcq[idx - 1];
                                lcq[idx - 1] = lt;
                            }
                        }
                    }
                }
                __syncwarp();
                if (ncand >= kCMax) ncand = kCMax - 1;
            }
            // local costs (A5) and the value each candidate would emit (A8)
            if (has_f) {
                const long long gf = ut.fr_off + (g_tile + w + kCandWarps * fq);
                const float lagwt = cf.lagwt;
                short *oloc = p.loc + gf * kCMax;
                float *omp = p.mp + gf * kCMax, *of0 = p.f0c + gf * kCMax;
                for (int slot = l8; slot < kCMax; slot += 8) {
                    if (slot < ncand) {
                        const int loc1 = lcq[slot];
                        const float pv = pkq[slot];
                        float ftemp = 1.0 - (static_cast<float>(loc1) * lagwt);
                        omp[slot] = 1.0 - (pv * ftemp);
                        oloc[slot] = static_cast<short>(loc1);
                        ftemp = loc1;
                        if (loc1 > cf.start && loc1 < cf.stop) {
                            const int jj = loc1 - cf.start;
                            const float cormax = ccq[jj], cprev = ccq[jj + 1], cnext = ccq[jj - 1];
                            const float den = (2.0 * (cprev + cnext - (2.0 * cormax)));
                            if (fabs(den) > 0.000001)
                                ftemp += 2.0 - ((((5.0 * cprev) + (3.0 * cnext) - (8.0 * cormax)) / den));
                        }
                        of0[slot] = 16000.0 / ftemp;
                    } else if (slot == ncand) {
                        oloc[slot] = -1;
                        omp[slot] = c_rapt.vbias + maxval;
                        of0[slot] = 0.0f;
                    } else {
                        oloc[slot] = -1;
                        omp[slot] = 0.0f;
                        of0[slot] = 0.0f;
                    }
                }
                if (l8 == 0) p.ncand[gf] = static_cast<unsigned char>(ncand + 1);
            }
            __syncwarp();
        }
        f_begin = f_end;
    }
}

// ---- K3 ------------------------------------------------------------------------------------
// Levinson-Durbin on all lanes redundantly, fully unrolled so a[] and b[] stay in registers
__device__ __forceinline__ void durbin18(const float *r, float *a_out, float *err)
{
    float a[kLpcOrd], b[kLpcOrd], k;
    float e = r[0];
    k = -r[1] / e;
    a[0] = k;
    e *= (1. - k * k);
#pragma unroll
    for (int i = 1; i < kLpcOrd; ++i) {
        float s = 0;
#pragma unroll
        for (int j = 0; j < i; ++j) s -= a[j] * r[i - j];
        k = (s - r[i + 1]) / e;
        a[i] = k;
#pragma unroll
        for (int j = 0; j <= i; ++j) b[j] = a[j];
#pragma unroll
        for (int j = 0; j < i; ++j) a[j] += k * b[i - j - 1];
        e *= (1. - (k * k));
    }
#pragma unroll
    for (int i = 0; i < kLpcOrd; ++i) a_out[i] = a[i];
    *err = e;
}

// One THREAD per 30 ms window.  The 19 autocorrelation chains (lags 0..18) of a window all consume
// the same samples, so a thread keeps a sliding window of pre-emphasised, Hanning-weighted samples in
// registers and feeds the chains from it; every chain is still summed left to right by one thread,
// product rounded before the add, so the result is bit-identical to the serial original.
// The chains are held as PAIRS (lag 2t, lag 2t+1) and advanced with packed multiplies and adds: for
// sample j the pair needs (d[j+2t], d[j+2t+1]), i.e. even-aligned sample pairs E for even j and
// odd-aligned pairs O for odd j; both are rings of 11 register pairs, refilled by one packed
// produce step (window, pre-emphasis, energy) per two samples.  20 packed + ~3 move instructions
// per sample instead of 38 scalar ones.
// (Packing the previous and the current window of one frame into the two halves instead - all operands
// naturally paired, 29 instructions per window sample, no moves - was measured too: a thread per FRAME
// leaves room for only 5 warps per SM beside the 35 KB staged span, and the kernel took 21.9 ms.)
// A CTA covers 32 consecutive frames of one utterance: warp 0 takes the current windows
// (x + 256 g - 80), warp 1 the previous ones (x + 256 g - 400); the signal span is staged once
// in shared memory - for interior tiles by 35 TMA bulk copies of one 256-sample row each, issued by
// one thread and awaited once (the element-wise staging it replaces exposed nine load round trips,
// a fifth of the CTA's life); raw samples, the factor 32768 is folded exactly into the window tables.
// Four pad words per 256 samples keep the rows 16-byte aligned for the copies.
constexpr int kStatFrames = 32;
constexpr int kStatSpan = kHop * (kStatFrames - 1) + kStatGap + kStatW;      // 16928 samples
constexpr int kStatSpanPad = kStatSpan + 32;                                  // read-ahead of the last window
constexpr int kStatXWords = kStatSpanPad + 4 * (kStatSpanPad / 256) + 4;
constexpr int kStatWinPairs = 256;                                            // window pairs (zero past 479 / 480)
__device__ __forceinline__ int stat_skew(int i) { return i + 4 * (i >> 8); }

__global__ void __launch_bounds__(2 * kStatFrames) rapt_stat_kernel(const RaptParams p, const int *__restrict__ tile_off,
                                                                    const int *__restrict__ tile_map)
{
    extern __shared__ __align__(16) float s_stat[];
    float4 *s_w4 = reinterpret_cast<float4 *>(s_stat);      // [256] {w480[i], w480[i+1], w479[i], w479[i+1]}, i = 2q
    float *s_x = s_stat + 4 * kStatWinPairs;                // [kStatXWords], skewed
    float *s_ex = s_x + kStatXWords;                        // [frames][20]: rho1[1..18], err1, rms1 of the previous window

    const int tid = threadIdx.x;
    const int u = tile_map[blockIdx.x];
    const RaptUtt ut = p.utts[u];
    const int g0 = (static_cast<int>(blockIdx.x) - tile_off[u]) * kStatFrames;
    const float *x = p.wav + ut.wav_off;
    const int s_lo = kHop * g0 - (kStatGap + 80);
    __shared__ __align__(8) uint64_t s_bar;
    // every staged sample is a sample of the utterance, and the run is 16-byte aligned (always true in the
    // padded segment layout of ssfe_extract; the stage-level ssfe_rapt takes arbitrary offsets)
    const bool interior = s_lo >= 0 && s_lo + kStatSpanPad <= ut.L && (reinterpret_cast<uintptr_t>(x + s_lo) & 15) == 0;
    if (interior) {
        if (tid == 0) {
            mbar_init(&s_bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            mbar_expect_tx(&s_bar, kStatSpanPad * 4);
            constexpr int kRows = kStatSpanPad / 256, kRest = kStatSpanPad - 256 * kRows;
            for (int k = 0; k < kRows; ++k) tma_load_1d(s_x + 260 * k, x + s_lo + 256 * k, 1024, &s_bar);
            if (kRest) tma_load_1d(s_x + 260 * kRows, x + s_lo + 256 * kRows, kRest * 4, &s_bar);
        }
    } else {
        for (int i0 = 0; i0 < kStatSpanPad; i0 += 16 * 2 * kStatFrames) {   // 16 loads in flight per thread
            float raw[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const int idx = s_lo + i0 + tid + q * 2 * kStatFrames;
                raw[q] = x[min(max(idx, 0), ut.L - 1)];
            }
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const int i = i0 + tid + q * 2 * kStatFrames, idx = s_lo + i;
                if (i < kStatSpanPad) s_x[stat_skew(i)] = (idx >= 0 && idx < ut.L) ? raw[q] : 0.0f;
            }
        }
    }
    // window tables, times 32768: the original scales the samples, a power of two commutes with every
    // rounding on the way, so scaling the weights instead gives the same bits
    for (int q = tid; q < kStatWinPairs; q += 2 * kStatFrames) {
        const int i = 2 * q;
        float4 w;
        w.x = (i < kStatW) ? p.w480[i] * 32768.0f : 0.0f;
        w.y = (i + 1 < kStatW) ? p.w480[i + 1] * 32768.0f : 0.0f;
        w.z = (i < kStatW - 1) ? p.w479[i] * 32768.0f : 0.0f;
        w.w = (i + 1 < kStatW - 1) ? p.w479[i + 1] * 32768.0f : 0.0f;
        s_w4[q] = w;
    }
    __syncthreads();
    if (interior) mbar_wait(&s_bar, 0);

    const bool is_prev = tid >= kStatFrames;
    const int fl = is_prev ? tid - kStatFrames : tid;      // local frame
    const int g = g0 + fl;
    const bool live = (g < ut.n_fr) && (g >= 2);
    const int b = kHop * fl + (is_prev ? 0 : kStatGap);    // window start inside the staged span

    float acc[kLpcOrd + 1];
    float rms = 0.0f;
    if (live) {
        float en = 0.0f;
        const p2 npre = p2pack(-c_rapt.preemp, -c_rapt.preemp);
        auto ldx = [&](int i) -> p2 {                      // (x[b+i], x[b+i+1]), i even
            const int a = b + i;
            return *reinterpret_cast<const p2 *>(s_x + stat_skew(a));
        };
        p2 xc = ldx(0);
        // produce2(i): the windowed pre-emphasised samples (d[i], d[i+1]) - zero from 479 on - and the
        // energy terms of samples i, i+1 (zero from 480 on), in the original's order
        auto produce2 = [&](int i) -> p2 {
            const ulonglong2 w = *reinterpret_cast<const ulonglong2 *>(s_w4 + (i >> 1));   // (w480 pair, w479 pair)
            const p2 xn = ldx(i + 2);
            const p2 f = p2mul(w.x, xc);
            const p2 ff = p2mul(f, f);
            en += p2lo(ff);
            en += p2hi(ff);
            const p2 t = p2mul(npre, xc);                                      // -(preemp * x[i])
            const p2 d = p2mul(w.y, p2add(p2pack(p2hi(xc), p2lo(xn)), t));
            xc = xn;
            return d;
        };
        p2 E[11], O[11], A[10];
        // An odd-aligned pair is two halves of its even-aligned neighbours.  Left as a plain repacking, ptxas does not
        // keep the eleven O pairs in registers: it rebuilds one (two MOVs into a scratch pair) for every use, ten
        // uses per double step - 160 of the loop's 761 instructions were MOVs.  Passing the repacked pair through a
        // multiply by a 1.0 the compiler cannot see (exact) makes it a value worth keeping: one FMUL2 and two MOVs
        // per double step instead.
        const p2 one2 = p2pack(c_rapt.one, c_rapt.one);
        auto odd_pair = [&](p2 lo_src, p2 hi_src) -> p2 { return p2mul(p2pack(p2hi(lo_src), p2lo(hi_src)), one2); };
#pragma unroll
        for (int t = 0; t < 11; ++t) E[t] = produce2(2 * t);
#pragma unroll
        for (int t = 0; t < 10; ++t) {
            O[t] = odd_pair(E[t], E[t + 1]);
            A[t] = p2pack(0.0f, 0.0f);
        }
        O[10] = p2pack(0.0f, 0.0f);
        // 22 x 11 double steps = 484 samples; the ones past 478 are zeros and add nothing
        for (int jb = 0; jb < 22; ++jb) {
#pragma unroll
            for (int st = 0; st < 11; ++st) {
                const p2 cur = E[st];
                const p2 d0 = p2pack(p2lo(cur), p2lo(cur)), d1 = p2pack(p2hi(cur), p2hi(cur));
#pragma unroll
                for (int t = 0; t < 10; ++t) A[t] = p2add(A[t], p2mul(d0, E[(st + t) % 11]));
#pragma unroll
                for (int t = 0; t < 10; ++t) A[t] = p2add(A[t], p2mul(d1, O[(st + t) % 11]));
                const p2 nw = produce2(22 * jb + 2 * st + 22);
                O[(st + 10) % 11] = odd_pair(E[(st + 10) % 11], nw);
                E[st] = nw;
            }
        }
#pragma unroll
        for (int t = 0; t < 10; ++t) {
            acc[2 * t] = p2lo(A[t]);
            if (2 * t + 1 <= kLpcOrd) acc[2 * t + 1] = p2hi(A[t]);
        }
        rms = static_cast<float>(sqrt(static_cast<double>(en / kStatW)));
    }
    float a_lpc[kLpcOrd], err = 0.0f;
    float rho[kLpcOrd + 1];
    if (live) {
        const float sum0 = acc[0];
        rho[0] = 1.0f;
        if (sum0 == 0.0f) {
#pragma unroll
            for (int k = 1; k <= kLpcOrd; ++k) rho[k] = 0.0f;
        } else {
            const float inv = 1.0 / sum0;
#pragma unroll
            for (int k = 1; k <= kLpcOrd; ++k) rho[k] = c_rapt.ffact * (acc[k] * inv);
        }
        durbin18(rho, a_lpc, &err);
        if (is_prev) {
            float *e = s_ex + fl * 20;
#pragma unroll
            for (int k = 1; k <= kLpcOrd; ++k) e[k - 1] = rho[k];
            e[18] = err;
            e[19] = rms;
        }
    }
    __syncthreads();
    if (is_prev || g >= ut.n_fr) return;
    const long long gf = ut.fr_off + g;
    if (g < 2) {
        p.sta[gf] = (g == 0) ? (0.01f * 0.2f) : static_cast<float>(0.2 / 10.0f);
        p.rr[gf] = 1.0f;
        return;
    }
    // b = autocorrelation of the current inverse filter (a_to_aca), Itakura distance against the
    // previous window's autocorrelation
    const float *e = s_ex + fl * 20;
    float s = 1.;
#pragma unroll
    for (int i = 0; i < kLpcOrd; ++i) s += a_lpc[i] * a_lpc[i];
    float dist = s;
#pragma unroll
    for (int i = 1; i <= kLpcOrd; ++i) {
        float q = a_lpc[i - 1];
#pragma unroll
        for (int j = 0; j < kLpcOrd - i; ++j) q += (a_lpc[j] * a_lpc[j + i]);
        const float bi = 2. * q;
        dist += e[i - 1] * bi;
    }
    const float t = (dist / e[18]) - .8;
    p.rr[gf] = (0.001 + rms) / e[19];
    p.sta[gf] = static_cast<float>(0.2 / t);
}

constexpr size_t kStatSmem = (4 * kStatWinPairs + kStatXWords + kStatFrames * 20) * sizeof(float);

// ---- K4 ------------------------------------------------------------------------------------
// The kernel is one latency chain per utterance, so every round trip to global memory inside the
// frame loop is removed: the records of frame g+1 are fetched while frame g is processed, and the
// voiced->voiced jump cost  float(log(lag2 / lag1))  is formed from a 321-entry table of log(lag)
// held in shared memory (one value per candidate, the previous frame's arrive by shuffle) instead of
// being gathered from the 257 x 257 table in L2.  log(l2) - log(l1) rounded to float is verified on the
// host, for every lag pair, to equal the original's float(log(l2 / l1)); if it ever does not, the
// kernel falls back to the exact table (use_log == 0).
constexpr int kDpWarps = 4;
constexpr int kMaxLag = 336;
constexpr int kDpBlk = 16;         // frames per record block of the Viterbi pass (g >> 4 below)

__device__ __forceinline__ float jump_cost(float ftemp, float ln2, float fdouble, float freqwt)
{
    // The original mixes precisions: ttemp = (float) fabs((double) ftemp), ft1 = (float) ((double) fdouble +
    // fabs((double) (ftemp + ln2))).  Both are reproduced exactly in float: widening and fabs are exact, and the
    // double sum of the two floats (0.35 and a value below 8) is exact unless the smaller one is below 2^-29 of
    // the larger - where it cannot reach a rounding boundary of the float result either - so rounding it to float
    // is the correctly rounded float sum.  (Six fp64 conversions less on the Viterbi pass's critical path.)
    // (the original's  if (ttemp > ft1) ttemp = ft1;  twice: a minimum of finite non-negative values, one FMNMX each)
    float ttemp = fabsf(ftemp);
    ttemp = fminf(ttemp, fdouble + fabsf(ftemp + ln2));
    ttemp = fminf(ttemp, fdouble + fabsf(ftemp - ln2));
    return ttemp * freqwt;
}

__global__ void __launch_bounds__(kDpWarps * 32) rapt_dp_kernel(const RaptParams p)
{
    __shared__ unsigned char s_ring[kDpWarps][kRing][kCMax];
    __shared__ unsigned char s_path[kDpWarps][kRing];
    __shared__ double s_log[kMaxLag];
    // records of 16 consecutive frames of each warp's utterance (see load_block below)
    __shared__ __align__(16) float s_bmp[kDpWarps][kDpBlk * kCMax];
    __shared__ __align__(16) int s_bloc[kDpWarps][kDpBlk * kCMax / 2];
    __shared__ float s_bsr[kDpWarps][2 * kDpBlk];          // per frame of the block: voiced<-unvoiced, unvoiced<-voiced transition cost
    __shared__ unsigned char s_bnc[kDpWarps][kDpBlk];
    // log(lag) of every candidate of the block's frames (0 for unvoiced / absent), row 0 = the last frame of the
    // previous block: looked up once per block, sixteen frames side by side.  A frame takes its own row per lane and the
    // previous frame's values as broadcast reads - neither depends on the Viterbi chain, only the accumulated costs
    // (one shuffle per previous candidate) do.  (The first version shuffled lag, cost and the two halves of log(lag).)
    __shared__ __align__(16) double s_blg[kDpWarps][(kDpBlk + 1) * kCMax];
    for (int i = threadIdx.x; i < kMaxLag; i += blockDim.x) s_log[i] = p.log_lag[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int u = blockIdx.x * kDpWarps + w;
    if (u >= p.n) return;
    const RaptUtt ut = p.utts[u];
    const RaptCfg &cf = c_rapt.cfg[ut.cfg];
    const float *ferr_tab = p.ferr + cf.table_off;
    const bool use_log = p.use_log != 0;
    float *out = p.out + ut.out_off;
    unsigned char(*ring)[kCMax] = s_ring[w];
    unsigned char *path = s_path[w];
    const unsigned full_mask = 0xffffffffu;

    float d_prev = 0.0f;
    int loc_prev = -1, ncandp = 0;
    int head = -1, tail = 0, num_active = 0;
    int my_pre = 0;
    // Software pipeline in blocks of kDpBlk = 16 frames: while the frames of block b are processed out of shared
    // memory, the records of block b + 1 (320 local costs, 320 lags, 16 x stationarity / rms ratio / count) are in
    // flight into registers, all loads issued together; at the block boundary they are stored to shared memory and
    // the next block's loads go out.  The first version fetched ONE frame ahead, and a frame (~0.3 us of shuffles
    // and compares) is shorter than a trip to L2: ncu had a quarter of the stall samples on the first use of the
    // fetched record, and the pass cost ~1 us per frame - 4.8 ms for a 60 s utterance.
    // (prefetch.global.L1 eight frames ahead was measured first: no gain.)
    float *bmp = s_bmp[w];
    const short *bloc = reinterpret_cast<const short *>(s_bloc[w]);
    float *bsr = s_bsr[w];
    unsigned char *bnc = s_bnc[w];
    float r_mp[kDpBlk * kCMax / 32];
    int r_loc[kDpBlk * kCMax / 64];
    float r_sr = 0.0f;
    int r_nc = 0;
    auto load_block = [&](int blk) {            // global -> registers (the record arrays have kDpBlk frames of slack)
        const long long gf0 = ut.fr_off + static_cast<long long>(blk) * kDpBlk;
        const float *gm = p.mp + gf0 * kCMax;
        const int *gl = reinterpret_cast<const int *>(p.loc + gf0 * kCMax);
#pragma unroll
        for (int k = 0; k < kDpBlk * kCMax / 32; ++k) r_mp[k] = gm[lane + 32 * k];
#pragma unroll
        for (int k = 0; k < kDpBlk * kCMax / 64; ++k) r_loc[k] = gl[lane + 32 * k];
        r_sr = (lane < kDpBlk) ? p.sta[gf0 + lane] : p.rr[gf0 + lane - kDpBlk];
        r_nc = (lane < kDpBlk) ? p.ncand[gf0 + lane] : 0;
    };
    double *blg = s_blg[w];
    auto store_block = [&]() {                  // registers -> shared memory
        if (lane < kCMax) blg[lane] = blg[kDpBlk * kCMax + lane];
        __syncwarp();
#pragma unroll
        for (int k = 0; k < kDpBlk * kCMax / 64; ++k) {
            const int l0 = static_cast<short>(r_loc[k] & 0xffff), l1 = static_cast<short>(r_loc[k] >> 16);
            double2 v;                          // (slack frames past the utterance hold arbitrary lags: range-checked)
            v.x = (l0 > 0 && l0 < kMaxLag) ? s_log[l0] : 0.0;
            v.y = (l1 > 0 && l1 < kMaxLag) ? s_log[l1] : 0.0;
            *reinterpret_cast<double2 *>(blg + kCMax + 2 * (lane + 32 * k)) = v;
        }
#pragma unroll
        for (int k = 0; k < kDpBlk * kCMax / 32; ++k) bmp[lane + 32 * k] = r_mp[k];
#pragma unroll
        for (int k = 0; k < kDpBlk * kCMax / 64; ++k) s_bloc[w][lane + 32 * k] = r_loc[k];
        // the two voicing-transition costs of a frame depend on the frame alone: formed here for the block's 16
        // frames side by side (lane = frame) instead of by every lane in every frame - a division and four more
        // operations off the per-frame chain; same float expressions, same bits
        const float rr_f = __shfl_down_sync(full_mask, r_sr, kDpBlk);
        if (lane < kDpBlk) {
            const float sta = r_sr;
            bsr[lane] = c_rapt.tcost + (c_rapt.tfact_s * sta) + (c_rapt.tfact_a / rr_f);            // voiced from unvoiced
            bsr[kDpBlk + lane] = c_rapt.tcost + (c_rapt.tfact_s * sta) + (c_rapt.tfact_a * rr_f);   // unvoiced from voiced
            bnc[lane] = static_cast<unsigned char>(r_nc);
        }
        __syncwarp();
    };
    load_block(0);
    store_block();
    load_block(1);
    for (int r = 0; r <= ut.R_last; ++r) {
        const int nfr = (r < ut.R_last) ? cf.F : ut.nl;
        const bool last_time = (r == ut.R_last);
        num_active += nfr;
        for (int i = 0; i < nfr; ++i) {
            const int g = r * cf.F + i;
            const int fi = g & (kDpBlk - 1);
            if (fi == 0 && g > 0) {
                __syncwarp();                   // every lane is done with the previous block
                store_block();
                load_block((g >> 4) + 1);
            }
            // (the record is read unconditionally and masked afterwards: a load guarded by the candidate count waits
            // for the count's own trip to shared memory first - one more latency on every frame of the chain)
            const int lcl = min(lane, kCMax - 1);
            const int ncand = bnc[fi];
            const int loc_r = bloc[fi * kCMax + lcl];
            const float mp_r = bmp[fi * kCMax + lcl];
            const double lg_r = blg[(fi + 1) * kCMax + lcl];
            const float v_from_uv = bsr[fi], uv_from_v = bsr[kDpBlk + fi];
            const int loc = (lane < ncand) ? loc_r : -1;
            const float mp = (lane < ncand) ? mp_r : 0.0f;
            const double lg = (lane < ncand) ? lg_r : 0.0;
            const double *lg_row = blg + fi * kCMax;             // the previous frame's
            float errmin = FLT_MAX;
            int minloc = 0;
            // Only the previous frame's live candidates are visited (typically 3-5 of the 20 slots; the count is
            // warp-uniform).  They are taken four at a time, branch-free: the twelve shuffles of a group go out
            // together and the four costs are formed side by side - one frame is one dependent chain, and with a
            // candidate per iteration and divergent branches around the cost it was ~290 instructions long.
            if (use_log) {
                const bool v_cur = loc > 0;
                for (int j0 = 0; j0 < ncandp; j0 += 4) {
                    float errs[4];
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const int j = j0 + jj;
                        const double lg1 = lg_row[min(j, kCMax - 1)];             // same address in every lane
                        const float dp = __shfl_sync(full_mask, d_prev, j);
                        const float jc = jump_cost(static_cast<float>(lg - lg1), c_rapt.ln2, c_rapt.fdouble, c_rapt.freqwt);
                        const bool v_prev = lg1 != 0.0;                            // lags are >= fs / 600 = 26: log(lag) > 0
                        const float ferr = v_cur ? (v_prev ? jc : v_from_uv) : (v_prev ? uv_from_v : 0.0f);
                        errs[jj] = (j < ncandp) ? ferr + dp : FLT_MAX;
                    }
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj)                 // first minimum wins, in candidate order
                        if (errs[jj] < errmin) { errmin = errs[jj]; minloc = j0 + jj; }
                }
            } else {
                for (int j = 0; j < ncandp; ++j) {
                    const int loc1 = __shfl_sync(full_mask, loc_prev, j);
                    const float dp = __shfl_sync(full_mask, d_prev, j);
                    float ferr;
                    if (loc > 0) ferr = (loc1 > 0) ? ferr_tab[(loc1 - cf.start) * cf.nlags + (loc - cf.start)] : v_from_uv;
                    else ferr = (loc1 > 0) ? uv_from_v : 0.0f;
                    const float err = ferr + dp;
                    if (err < errmin) { errmin = err; minloc = j; }
                }
            }
            float dcur;
            if (g == 0) { dcur = mp; my_pre = 0; }
            else { dcur = errmin + mp; my_pre = minloc; }
            if (lane < kCMax) ring[g & (kRing - 1)][lane] = static_cast<unsigned char>(my_pre);
            d_prev = dcur;
            loc_prev = loc;
            ncandp = ncand;
            head = g;
        }
        __syncwarp();
        // ---- commit: back-track from a frame where all surviving paths agree ------------------
        if (head >= 0 && (num_active >= c_rapt.size_frame_hist || last_time)) {
            const int num_paths = ncandp;
            // best candidate of the newest frame: first minimum of the accumulated cost
            float bv = (lane < num_paths) ? d_prev : FLT_MAX;
            int bi = lane;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(full_mask, bv, o);
                const int oi = __shfl_xor_sync(full_mask, bi, o);
                if (ov < bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            int best_cand = bi;
            int pc = (lane < num_paths) ? my_pre : 0;
            int cmpth = -1;
            bool done = true;
            if (last_time) {
                cmpth = head;
            } else {
                int frm = head;
                while (true) {
                    frm -= 1;
                    const int v0 = __shfl_sync(full_mask, pc, 0);
                    done = __all_sync(full_mask, (lane >= num_paths) || (pc == v0));
                    if (!done) {
                        if (lane < num_paths) pc = ring[frm & (kRing - 1)][pc];
                    } else {
                        cmpth = frm;
                        best_cand = v0;
                        break;
                    }
                    if (frm <= tail) {
                        if (num_active < c_rapt.size_frame_out) { done = false; cmpth = -1; }
                        else { done = true; cmpth = head; }
                        break;
                    }
                }
            }
            if (done) {
                // lane 0 follows the back pointers (shared memory only), then all lanes emit
                if (lane == 0) {
                    int bc = best_cand;
                    for (int frm = cmpth; frm >= tail; --frm) {
                        path[frm & (kRing - 1)] = static_cast<unsigned char>(bc);
                        bc = ring[frm & (kRing - 1)][bc];
                    }
                }
                __syncwarp();
                for (int frm = tail + lane; frm <= cmpth; frm += 32) {
                    const float f0 = p.f0c[(ut.fr_off + frm) * kCMax + path[frm & (kRing - 1)]];
                    out[frm] = (f0 == 0.0f) ? kUnvoiced : static_cast<float>(log(static_cast<double>(f0)));
                }
                num_active -= (cmpth - tail + 1);
                tail = cmpth + 1;
            }
        }
        __syncwarp();
    }
    for (int t = ut.n_fr + lane; t < ut.n_out; t += 32) out[t] = kUnvoiced;
    if (head < 0)
        for (int t = lane; t < ut.n_fr; t += 32) out[t] = kUnvoiced;
}

// tile -> utterance lookups of the three frame-parallel kernels in one launch (blockIdx.y picks the table)
__global__ void rapt_maps_kernel(const long long *__restrict__ ds_offs, const int *__restrict__ cand_tiles,
                                 const int *__restrict__ stat_tiles, int n, int *__restrict__ dec_map,
                                 int *__restrict__ cand_map, int *__restrict__ stat_map)
{
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n) return;
    if (blockIdx.y == 0) {
        const long long t1 = ds_offs[u + 1] / kDecTile;
        for (long long t = ds_offs[u] / kDecTile; t < t1; ++t) dec_map[t] = u;
    } else {
        const int *off = (blockIdx.y == 1) ? cand_tiles : stat_tiles;
        int *map = (blockIdx.y == 1) ? cand_map : stat_map;
        const int t1 = off[u + 1];
        for (int t = off[u]; t < t1; ++t) map[t] = u;
    }
}

// ---- host side ---------------------------------------------------------------------------------
static int iround(double x) { return static_cast<int>(x + 0.5); }

int init_rapt(ssfe_ctx *ctx)
{
    RaptTables *T = new RaptTables();
    ctx->rapt = T;
    RaptConsts &h = T->host;
    memset(&h, 0, sizeof(h));
    const double freq = 16000.0;
    // ESPS defaults installed by SPTK's rapt(); single precision like F0_params
    const float cand_thresh = 0.3f, lag_weight = 0.3f, freq_weight = 0.02f, trans_cost = 0.005f, trans_amp = 0.5f,
                trans_spec = 0.5f, voice_bias = 0.0f, double_cost = 0.35f, wind_dur = 0.0075f;
    const float frame_step = static_cast<float>(256.0 / freq);
    const int step = iround(frame_step * freq), size = iround(wind_dur * freq);
    const float frame_int = static_cast<float>(static_cast<float>(step) / freq);
    if (step != kHop || size != kWin) return set_error(ctx, SSFE_ERR_INVALID, "RAPT geometry mismatch");
    h.cand_thresh = cand_thresh;
    h.tcost = trans_cost;
    h.tfact_a = trans_amp;
    h.tfact_s = trans_spec;
    h.vbias = voice_bias;
    h.preemp = 0.4f;
    h.ffact = static_cast<float>(1.0 / (1.0 + exp((-30.0f / 20.0) * log(10.0))));
    h.size_frame_hist = static_cast<int>(0.5 / frame_int);
    h.size_frame_out = static_cast<int>(1.0 / frame_int);
    const float ln2 = static_cast<float>(log(2.0));
    const float freqwt = freq_weight / frame_int;
    h.ln2 = ln2;
    h.one = 1.0f;
    h.fdouble = double_cost;
    h.freqwt = freqwt;
    std::vector<double> log_lag(336, 0.0);
    for (int l = 1; l < 336; ++l) log_lag[l] = log(static_cast<double>(l));
    bool log_ok = true;
    const float ranges[2][2] = {{50.0f, 250.0f}, {100.0f, 600.0f}};
    std::vector<float> ferr;
    for (int c = 0; c < 2; ++c) {
        RaptCfg &cf = h.cfg[c];
        cf.start = iround(freq / ranges[c][1]);
        cf.stop = iround(freq / ranges[c][0]);
        cf.nlags = cf.stop - cf.start + 1;
        cf.ncomp = size + cf.stop + 1;
        const int i = static_cast<int>(0.2 * freq);
        cf.F = (cf.ncomp >= step) ? ((i - cf.ncomp) / step) + 1 : i / step;
        const int downpatch = ((static_cast<int>(freq * 0.005)) + 1) / 2;
        const int stat_wsize = static_cast<int>(0.030 * freq), agap = static_cast<int>(0.020 * freq);
        const int ind = (agap - stat_wsize) / 2, i2 = stat_wsize + ind;
        cf.pad = downpatch + ((i2 > cf.ncomp) ? i2 : cf.ncomp);
        cf.buff_size = cf.F * step + cf.pad;
        cf.sdstep = cf.F * step;
        cf.decnlags = 1 + (cf.nlags / kDec);
        cf.decstart = std::max(1, cf.start / kDec);
        cf.decsize = 1 + (size / kDec);
        cf.n_el = cf.decsize + cf.decstart + cf.decnlags;
        cf.lagwt = lag_weight / cf.stop;
        cf.lag_wt = lag_weight / cf.nlags;
        cf.table_off = static_cast<int>(ferr.size());
        if (cf.ncomp > 448 || cf.nlags + 8 > kCcMax || cf.n_el > 64)
            return set_error(ctx, SSFE_ERR_INVALID, "RAPT range does not fit the kernel buffers");
        // voiced -> voiced frequency-jump cost, with the original's precision mix, on the host so
        // that log() is the same libm the CPU path uses
        for (int l1 = cf.start; l1 <= cf.stop; ++l1)
            for (int l2 = cf.start; l2 <= cf.stop; ++l2) {
                float ftemp = log((static_cast<double>(l2)) / l1);
                float ttemp = fabs(ftemp);
                float ft1 = double_cost + fabs(ftemp + ln2);
                if (ttemp > ft1) ttemp = ft1;
                ft1 = double_cost + fabs(ftemp - ln2);
                if (ttemp > ft1) ttemp = ft1;
                ferr.push_back(ttemp * freqwt);
                // the shortcut the DP kernel uses: log(l2) - log(l1), rounded to float
                if (l1 >= 336 || l2 >= 336) { log_ok = false; continue; }
                float f2 = static_cast<float>(log_lag[l2] - log_lag[l1]);
                float t2 = fabs(f2);
                float g1 = double_cost + fabs(f2 + ln2);
                if (t2 > g1) t2 = g1;
                g1 = double_cost + fabs(f2 - ln2);
                if (t2 > g1) t2 = g1;
                if (t2 * freqwt != ferr.back()) log_ok = false;
            }
    }
    // decimation filter: Hanning-windowed sinc, 81 taps, cut-off 0.5/8 cycles per sample
    {
        int nf = (static_cast<int>(freq * .005)) | 1;
        const float fc = .5 / kDec;
        if ((nf % 2) != 1) nf = nf + 1;
        const int n = (nf + 1) / 2;
        if (nf != kNco) return set_error(ctx, SSFE_ERR_INVALID, "RAPT decimator size mismatch");
        std::vector<float> coef(n);
        const double twopi = M_PI * 2.0;
        coef[0] = 2.0 * fc;
        const double c = M_PI;
        double fn = twopi * fc;
        for (int i = 1; i < n; ++i) coef[i] = sin(i * fn) / (c * i);
        fn = twopi / static_cast<double>(nf);
        for (int i = 0; i < n; ++i) coef[n - i - 1] *= (.5 - (.5 * cos(fn * (static_cast<double>(i) + 0.5))));
        for (int i = 0; i < n - 1; ++i) h.co[i] = h.co[nf - 1 - i] = coef[n - 1 - i];
        h.co[n - 1] = coef[0];
    }
    std::vector<float> w479(kStatW, 0.0f), w480(kStatW, 0.0f);
    for (int n = kStatW - 1; n <= kStatW; ++n) {
        std::vector<float> &wv = (n == kStatW) ? w480 : w479;
        const double arg = 3.1415927 * 2.0 / n, half = 0.5;
        for (int i = 0; i < n; ++i) wv[i] = (half - half * cos((half + static_cast<double>(i)) * arg));
    }
    T->use_log = log_ok ? 1 : 0;
    SSFE_CUDA(ctx, cudaMalloc(&T->d_log, log_lag.size() * sizeof(double)));
    SSFE_CUDA(ctx, cudaMemcpy(T->d_log, log_lag.data(), log_lag.size() * sizeof(double), cudaMemcpyHostToDevice));
    SSFE_CUDA(ctx, cudaMalloc(&T->d_ferr, ferr.size() * sizeof(float)));
    SSFE_CUDA(ctx, cudaMalloc(&T->d_w479, kStatW * sizeof(float)));
    SSFE_CUDA(ctx, cudaMalloc(&T->d_w480, kStatW * sizeof(float)));
    SSFE_CUDA(ctx, cudaMemcpy(T->d_ferr, ferr.data(), ferr.size() * sizeof(float), cudaMemcpyHostToDevice));
    SSFE_CUDA(ctx, cudaMemcpy(T->d_w479, w479.data(), kStatW * sizeof(float), cudaMemcpyHostToDevice));
    SSFE_CUDA(ctx, cudaMemcpy(T->d_w480, w480.data(), kStatW * sizeof(float), cudaMemcpyHostToDevice));
    SSFE_CUDA(ctx, cudaMemcpyToSymbol(c_rapt, &h, sizeof(h)));
    {   // coefficient pairs of the decimator: [i + 16] = (co[i], co[i - 8])
        std::vector<float2> cp(kDecSpan + 16);
        for (int k = 0; k < kDecSpan + 16; ++k) {
            const int i = k - 16;
            cp[k].x = (i >= 0 && i < kNco) ? h.co[i] : 0.0f;
            cp[k].y = (i - 8 >= 0 && i - 8 < kNco) ? h.co[i - 8] : 0.0f;
        }
        SSFE_CUDA(ctx, cudaMemcpyToSymbol(c_dec_pair, cp.data(), cp.size() * sizeof(float2)));
    }
    SSFE_CUDA(ctx, cudaFuncSetAttribute(rapt_stat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(kStatSmem)));
    return SSFE_OK;
}

void free_rapt(ssfe_ctx *ctx)
{
    if (!ctx->rapt) return;
    cudaFree(ctx->rapt->d_ferr);
    cudaFree(ctx->rapt->d_log);
    cudaFree(ctx->rapt->d_w479);
    cudaFree(ctx->rapt->d_w480);
    delete ctx->rapt;
    ctx->rapt = nullptr;
}

int rapt_run(ssfe_ctx *ctx, const float *wav_base, const int64_t *start_host, const int64_t *len_host,
             const int64_t *frame_off_host, int n, const float *f0_lo, const float *f0_hi, float *f0_dev)
{
    if (n == 0) return SSFE_OK;
    const RaptConsts &h = ctx->rapt->host;
    std::vector<RaptUtt> utts(n);
    std::vector<long long> fr_offs(n + 1), ds_offs(n + 1);
    long long fr = 0, dsn = 0;
    for (int i = 0; i < n; ++i) {
        RaptUtt &ut = utts[i];
        memset(&ut, 0, sizeof(ut));
        const int64_t L = len_host[i];
        // get_f0's minimum: total_samps < ((frame_step * 2.0) + wind_dur) * fs with float parameters
        // (0.016f and 0.0075f round up, so 632 samples are rejected and 633 accepted)
        if (static_cast<double>(L) < ((static_cast<float>(256.0 / 16000.0) * 2.0) + 0.0075f) * 16000.0)
            return set_error(ctx, SSFE_ERR_TOO_SHORT, "utterance %d: input range too small for analysis by get_f0", i);
        if (L > 0x7fffffff / 2) return set_error(ctx, SSFE_ERR_INVALID, "utterance %d too long for RAPT", i);
        ut.cfg = (f0_lo[i] == 50.0f && f0_hi[i] == 250.0f) ? 0 : 1;
        if (ut.cfg == 1 && !(f0_lo[i] == 100.0f && f0_hi[i] == 600.0f))
            return set_error(ctx, SSFE_ERR_GENDER, "utterance %d: unsupported F0 range (%g, %g)", i, f0_lo[i], f0_hi[i]);
        const RaptCfg &cf = h.cfg[ut.cfg];
        ut.L = static_cast<int>(L);
        ut.wav_off = start_host[i];
        ut.out_off = frame_off_host[i];
        ut.n_out = static_cast<int>(frame_off_host[i + 1] - frame_off_host[i]);
        ut.R_last = (L > cf.buff_size) ? static_cast<int>((L - cf.buff_size + cf.sdstep - 1) / cf.sdstep) : 0;
        const int64_t rem = L - static_cast<int64_t>(ut.R_last) * cf.sdstep;
        ut.nl = (rem < cf.pad) ? 0 : static_cast<int>((rem - cf.pad) / kHop);
        ut.n_fr = ut.R_last * cf.F + ut.nl;
        if (ut.n_fr > ut.n_out) return set_error(ctx, SSFE_ERR_INVALID, "utterance %d: frame count mismatch", i);
        ut.n_ds = (static_cast<int>(L / kDec) + 8 + kDecTile - 1) / kDecTile * kDecTile;
        ut.fr_off = fr;
        ut.ds_off = dsn;
        fr_offs[i] = fr;
        ds_offs[i] = dsn;
        fr += ut.n_fr;
        dsn += ut.n_ds;
    }
    fr_offs[n] = fr;
    ds_offs[n] = dsn;
    std::vector<int> stat_tile_off(n + 1), cand_tile_off(n + 1);
    long long stat_tiles = 0, cand_tiles = 0;
    for (int i = 0; i < n; ++i) {
        stat_tile_off[i] = static_cast<int>(stat_tiles);
        cand_tile_off[i] = static_cast<int>(cand_tiles);
        stat_tiles += (utts[i].n_fr + kStatFrames - 1) / kStatFrames;
        cand_tiles += (utts[i].n_fr + kCandTile - 1) / kCandTile;
    }
    stat_tile_off[n] = static_cast<int>(stat_tiles);
    cand_tile_off[n] = static_cast<int>(cand_tiles);

    int rc;
    if ((rc = ensure(ctx, ctx->ws.rapt_ds, dsn * sizeof(float)))) return rc;
    if ((rc = ensure(ctx, ctx->ws.dec_map, (dsn / kDecTile + 1) * sizeof(int)))) return rc;
    int *dec_map = static_cast<int *>(ctx->ws.dec_map.p);
    if ((rc = ensure(ctx, ctx->ws.cand_map, (cand_tiles + 1) * sizeof(int)))) return rc;
    if ((rc = ensure(ctx, ctx->ws.stat_map, (stat_tiles + 1) * sizeof(int)))) return rc;
    int *cand_map = static_cast<int *>(ctx->ws.cand_map.p), *stat_map = static_cast<int *>(ctx->ws.stat_map.p);
    const size_t per_fr = kCMax * (sizeof(short) + 2 * sizeof(float)) + 2 * sizeof(float) + 8;
    // (kDpBlk + frames of slack: the Viterbi pass loads whole 16-frame blocks, up to two past an utterance's end)
    if ((rc = ensure(ctx, ctx->ws.rapt_cand, (fr + 48) * per_fr))) return rc;

    RaptParams p;
    memset(&p, 0, sizeof(p));
    p.wav = wav_base;
    const int *d_stat_tiles = upload(ctx, stat_tile_off.data(), n + 1);
    const int *d_cand_tiles = upload(ctx, cand_tile_off.data(), n + 1);
    if (!d_stat_tiles || !d_cand_tiles) return SSFE_ERR_NOMEM;
    p.utts = upload(ctx, utts.data(), n);
    p.fr_offs = upload(ctx, fr_offs.data(), n + 1);
    p.ds_offs = upload(ctx, ds_offs.data(), n + 1);
    if (!p.utts || !p.fr_offs || !p.ds_offs) return SSFE_ERR_NOMEM;
    if ((rc = flush_meta(ctx))) return rc;
    p.n = n;
    p.total_fr = fr;
    p.total_ds = dsn;
    p.ds = static_cast<float *>(ctx->ws.rapt_ds.p);
    char *base = static_cast<char *>(ctx->ws.rapt_cand.p);
    const long long frp = fr + 48;
    p.mp = reinterpret_cast<float *>(base);
    p.f0c = p.mp + frp * kCMax;
    p.sta = p.f0c + frp * kCMax;
    p.rr = p.sta + frp;
    p.loc = reinterpret_cast<short *>(p.rr + frp);
    p.ncand = reinterpret_cast<unsigned char *>(p.loc + frp * kCMax);
    p.ferr = ctx->rapt->d_ferr;
    p.log_lag = ctx->rapt->d_log;
    p.use_log = ctx->rapt->use_log;
    p.w479 = ctx->rapt->d_w479;
    p.w480 = ctx->rapt->d_w480;
    p.out = f0_dev;

    RaptTables *T = ctx->rapt;
    T->last_fr = fr;
    T->last_ncand = p.ncand;
    T->last_loc = p.loc;
    T->last_mp = p.mp;
    T->last_f0c = p.f0c;
    T->last_sta = p.sta;
    T->last_rr = p.rr;

    cudaStream_t st = ctx->stream;
    rapt_maps_kernel<<<dim3(static_cast<unsigned>((n + 255) / 256), 3), 256, 0, st>>>(p.ds_offs, d_cand_tiles, d_stat_tiles, n, dec_map,
                                                                                       cand_map, stat_map);
    SSFE_LAUNCHED(ctx);
    rapt_decimate_kernel<<<static_cast<unsigned>(dsn / kDecTile), kDecThreads, 0, st>>>(p, dec_map);
    SSFE_LAUNCHED(ctx);
    mark(ctx, ST_RAPT_CAND);
    // the zero stream of the one-hot output starts here by default (extract_device): beside the candidate kernel,
    // which leaves HBM idle - beside the decimation kernel it cost that kernel what it saved at the end
    if ((rc = onehot_zero_fork(ctx, 2))) return rc;
    if (fr > 0) {
        rapt_cand_kernel<<<static_cast<unsigned>(cand_tiles), kCandWarps * 32, 0, st>>>(p, d_cand_tiles, cand_map);
        SSFE_LAUNCHED(ctx);
        mark(ctx, ST_RAPT_STAT);
        // the next call's dither generation (side stream, high priority) may start here, beside the stationarity
        // kernel (extract_device has the measurements)
        if (!ctx->mt_go_at_start) SSFE_CUDA(ctx, cudaEventRecord(ctx->ev_mt_go, st));
        if ((rc = onehot_zero_fork(ctx, 3))) return rc;
        rapt_stat_kernel<<<static_cast<unsigned>(stat_tiles), 2 * kStatFrames, kStatSmem, st>>>(p, d_stat_tiles, stat_map);
        SSFE_LAUNCHED(ctx);
    } else {
        mark(ctx, ST_RAPT_STAT);
        if (!ctx->mt_go_at_start) SSFE_CUDA(ctx, cudaEventRecord(ctx->ev_mt_go, st));
    }
    mark(ctx, ST_RAPT_DP);
    if ((rc = onehot_zero_fork(ctx, 3)) || (rc = onehot_zero_fork(ctx, 4))) return rc;
    rapt_dp_kernel<<<static_cast<unsigned>((n + kDpWarps - 1) / kDpWarps), kDpWarps * 32, 0, st>>>(p);
    SSFE_LAUNCHED(ctx);
    return SSFE_OK;
}

}  // namespace ssfe

// diagnostics: per-frame records of the most recent RAPT run (synchronous)
extern "C" int64_t ssfe_rapt_dump(ssfe_ctx *ctx, int64_t max_frames, uint8_t *ncand, int16_t *loc, float *mp,
                                  float *f0cand, float *stat, float *rms_ratio)
{
    if (!ctx || !ctx->rapt) return SSFE_ERR_INVALID;
    const ssfe::RaptTables *T = ctx->rapt;
    const long long n = std::min<long long>(T->last_fr, max_frames);
    if (n <= 0) return 0;
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return SSFE_ERR_CUDA;
    cudaError_t e = cudaSuccess;
    if (ncand && e == cudaSuccess) e = cudaMemcpy(ncand, T->last_ncand, n, cudaMemcpyDeviceToHost);
    if (loc && e == cudaSuccess) e = cudaMemcpy(loc, T->last_loc, n * ssfe::kCMax * sizeof(short), cudaMemcpyDeviceToHost);
    if (mp && e == cudaSuccess) e = cudaMemcpy(mp, T->last_mp, n * ssfe::kCMax * sizeof(float), cudaMemcpyDeviceToHost);
    if (f0cand && e == cudaSuccess) e = cudaMemcpy(f0cand, T->last_f0c, n * ssfe::kCMax * sizeof(float), cudaMemcpyDeviceToHost);
    if (stat && e == cudaSuccess) e = cudaMemcpy(stat, T->last_sta, n * sizeof(float), cudaMemcpyDeviceToHost);
    if (rms_ratio && e == cudaSuccess) e = cudaMemcpy(rms_ratio, T->last_rr, n * sizeof(float), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return ssfe::cuda_fail(ctx, e, "ssfe_rapt_dump");
    return n;
}
