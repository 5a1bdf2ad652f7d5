// stft_mel.cu - fused pySTFT -> mel -> dB -> normalise kernel (stages a3+a4+a5).
//
// Replaces, per frame, reference utils.py:18-31 (reflect pad, 1024-sample frames at hop 256,
// periodic Hann, rfft, abs) and make_spect_f0.py:59-61 (dot with the (513,80) mel basis,
// 20*log10(max(1e-5, .)) - 16, (.+100)/100).  One launch covers a whole ragged batch.
//
// Data layout in HBM: the dithered wav of utterance i lives in a *padded segment*
//     [512 reflected | L_i samples | 512 reflected | slack]          (float32, 256 B aligned)
// so frame t of utterance i is the contiguous run  seg_i + 256 t ... + 1024, and a PAIR of adjacent
// frames is one contiguous 5 KiB run.
//
// Work split (v2): one persistent CTA of 12 warps per SM; every warp is an independent pipeline over
// frame pairs - no block-level barrier anywhere in the main loop:
//   * its pair is staged with ONE TMA bulk copy (cp.async.bulk -> SASS UBLKCP) completing on the
//     warp's own mbarrier.  The staged samples are only needed until they sit in registers (the
//     first ~130 instructions of ~1300), so the copy for the warp's NEXT pair is issued right after
//     that and lands while the FFT runs: a single 5 KiB stage buffer per warp is enough;
//   * the two frames are the real and imaginary parts of one complex 1024-point FFT held in
//     registers (fft_core.cuh, 32 x 32 Cooley-Tukey, transpose through shared memory); butterfly
//     adds, window multiplies, the spectrum split and the mel FMAs use the packed FP32 instructions
//     of sm_100a (FADD2 / FMUL2 / FFMA2): the kernel is bound by instruction issue, not by HBM;
//   * |A[k]|, |B[k]| of the two frames are stored interleaved so one 128-bit shared-memory load
//     feeds two bins of both frames; the 941 non-zero mel weights are grouped in aligned runs of 4
//     bins (302 groups); every lane walks a LIST of ~10 groups that strings several bands together
//     (longest-processing-time assignment, then a local search that spreads the 8 lanes of each
//     128-bit wavefront over the 8 bank groups), flushing its accumulator pair where a band ends.
// Arithmetic is fp32 on the CUDA cores - no tensor cores (BASELINE.json north_star).
#include "common.cuh"
#include "tma.cuh"
#include "fft_core.cuh"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <numeric>

namespace ssfe {

constexpr int kStftWarps = 12;
constexpr int kStftThreads = kStftWarps * 32;
constexpr int kPairFloats = kNfft + kHop;                   // 1280 samples = 5120 B per frame pair
constexpr int kWarpTrans = 32 * kTransStride;               // float2 per warp
constexpr int kMaxGroups = 16;
constexpr int kTwBatch = 8;                                 // inter-pass twiddles fetched per batch
constexpr int kMelFlush = 1 << 15;                          // mel_k flag: the lane's band ends with this group

struct PairInfo {            // 16 bytes, built on the device once per launch
    long long src;           // float offset of the pair's first sample in wavp
    int out_frame;           // global frame index of frame A
    int has_b;               // frame B exists
};

struct StftParams {
    const float *wavp;
    const PairInfo *pairs;
    int n_pairs;
    float *out;
    const float *window;
    const float2 *tw;
    const float4 *mel_w;     // [n_groups][32] weights of the lane's g-th group (already * 0.5)
    const int *mel_k;        // [n_groups][32] swizzled address of the group | kMelFlush | band << 16
    int n_groups;
    float min_level, c1, c0;
};

__device__ __forceinline__ float fast_sqrt(float x)
{
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fast_log2(float x)
{
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// pair table: thread per pair
__global__ void stft_pairs_kernel(const int64_t *__restrict__ seg_off, const int *__restrict__ pair_off,
                                  const int64_t *__restrict__ frame_off, int n, int n_pairs,
                                  PairInfo *__restrict__ pairs)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pairs) return;
    const int u = find_segment(pair_off, n, i);
    const int k = i - pair_off[u];
    const int n_frames = static_cast<int>(frame_off[u + 1] - frame_off[u]);
    PairInfo pi;
    pi.src = seg_off[u] + static_cast<long long>(k) * 2 * kHop;
    pi.out_frame = static_cast<int>(frame_off[u]) + 2 * k;
    pi.has_b = (2 * k + 1 < n_frames) ? 1 : 0;
    pairs[i] = pi;
}

struct __align__(16) WarpSmem {
    float stage[kPairFloats];            // 5120 B, TMA destination
    float2 trans[kWarpTrans];            // 8448 B; later re-used as the interleaved magnitudes
    float2 outs[kMels];                  // 640 B: (frame A, frame B) mel sums of every band
    uint64_t bar;
    uint64_t pad;
};

template <int MODE>
__global__ void __launch_bounds__(kStftThreads, 1) stft_mel_kernel(const StftParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    WarpSmem *ws_all = reinterpret_cast<WarpSmem *>(smem_raw);
    float2 *s_tw = reinterpret_cast<float2 *>(ws_all + kStftWarps);        // [1024]
    float *s_win = reinterpret_cast<float *>(s_tw + 1024);                 // [1024]
    float4 *s_mw = reinterpret_cast<float4 *>(s_win + 1024);               // [kMaxGroups][32]
    int *s_mk = reinterpret_cast<int *>(s_mw + kMaxGroups * 32);           // [kMaxGroups][32]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    WarpSmem &ws = ws_all[warp];

    for (int i = tid; i < 1024; i += kStftThreads) {
        s_tw[i] = p.tw[i];
        s_win[i] = p.window[i];
    }
    if (MODE == 0) {
        for (int i = tid; i < p.n_groups * 32; i += kStftThreads) {
            s_mw[i] = p.mel_w[i];
            s_mk[i] = p.mel_k[i];
        }
    }
    if (lane == 0) {
        mbar_init(&ws.bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();   // the only block-level barrier: tables and barriers are ready

    const int stride = gridDim.x * kStftWarps;
    int pair = blockIdx.x * kStftWarps + warp;
    PairInfo cur;
    cur.src = 0; cur.out_frame = 0; cur.has_b = 0;
    if (pair < p.n_pairs) {
        cur = p.pairs[pair];
        if (lane == 0) {
            mbar_expect_tx(&ws.bar, kPairFloats * 4);
            tma_load_1d(ws.stage, p.wavp + cur.src, kPairFloats * 4, &ws.bar);
        }
    }
    float2 *trans = ws.trans;
    float2 *mag2 = ws.trans;               // [520] (|A[k]|, |B[k]|) without the common factor 1/2
    uint32_t parity = 0;

    for (; pair < p.n_pairs; pair += stride) {
        // prefetch the descriptor of this warp's next pair (all lanes, same address)
        const int next = pair + stride;
        PairInfo nxt = cur;
        if (next < p.n_pairs) nxt = p.pairs[next];

        mbar_wait(&ws.bar, parity);
        parity ^= 1;
        float2 v[32];
        // (Re-using frame A's reads for frame B - sample m of B is sample m+8 of A, 40 reads instead of
        // 64 - was measured 4 % slower: the longer live ranges cost more than the 24 wavefronts saved.)
        if (cur.has_b) {
#pragma unroll
            for (int m = 0; m < 32; ++m) {
                const int n = lane + 32 * m;
                v[m] = cscale(make_float2(ws.stage[n], ws.stage[n + kHop]), s_win[n]);
            }
        } else {   // the missing second frame must be exactly zero: it shares the transform with A
#pragma unroll
            for (int m = 0; m < 32; ++m) {
                const int n = lane + 32 * m;
                v[m] = make_float2(ws.stage[n] * s_win[n], 0.0f);
            }
        }
        __syncwarp();                      // every lane has consumed the stage buffer
        if (lane == 0 && next < p.n_pairs) {
            mbar_expect_tx(&ws.bar, kPairFloats * 4);
            tma_load_1d(ws.stage, p.wavp + nxt.src, kPairFloats * 4, &ws.bar);
        }

        fft32_dif(v);
        // inter-pass twiddles W_1024^(lane k1) from the [k1][lane] table.  (Building them from five
        // table reads by complex products was measured: -52 shared-memory wavefronts per pair but +100
        // FP32 instructions made the kernel 7 % slower - issue and shared memory are balanced here.)
        // (taken kTwBatch at a time, loads first, the compiler kept from mixing them: with one load right in front of its
        // multiply the warp waited for shared memory on every twiddle - 17 % of the kernel's stall samples sat on the
        // multiply.  9.20 -> 8.87 ms at 8 or 16 per batch, 9.10 at 4; the same treatment of the window loads changed
        // nothing and a one-group prefetch in the mel loop cost 0.1 ms)
#pragma unroll
        for (int r0 = 0; r0 < 32; r0 += kTwBatch) {
            float2 t[kTwBatch];
#pragma unroll
            for (int q = 0; q < kTwBatch; ++q) {
                const int k1 = bitrev5(r0 + q);
                t[q] = (k1 != 0) ? s_tw[k1 * 32 + lane] : make_float2(1.0f, 0.0f);
            }
            asm volatile("" ::: "memory");
#pragma unroll
            for (int q = 0; q < kTwBatch; ++q) {
                const int r = r0 + q, k1 = bitrev5(r);
                float2 y = v[r];
                if (k1 != 0) y = make_float2(v[r].x * t[q].x - v[r].y * t[q].y, v[r].x * t[q].y + v[r].y * t[q].x);
                trans[lane * kTransStride + k1] = y;
            }
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = trans[j * kTransStride + lane];
        __syncwarp();
        fft32_dif(v);

        // separate the two real spectra: with X = v, P = X[1024-k] from the partner lane,
        //   2A = X + conj(P) = (Xr+Pr, Xi-Pi),   2iB = X - conj(P) = (Xr-Pr, Xi+Pi)
        // Magnitudes are stored as (|A[k]|, |B[k]|) pairs.  A 4-bin group is 32 bytes = two 16-byte
        // halves; a plain layout would put every first half in an even 16-byte bank group and every
        // second half in an odd one, so a 128-bit load could only ever use half the banks.  Swapping
        // the halves of groups 4..7 (mod 8) spreads the same half of 8 consecutive groups over all 8
        // bank groups: element k lives at k ^ 2 when bit 4 of k is set (for k = lane + 32 k2 that is
        // lane >= 16).  The mel tables hold the swizzled address of the first half; the second half is
        // at that address ^ 2.
        const int partner = (32 - lane) & 31;
        const int swz = (lane & 16) >> 3;
#pragma unroll
        for (int k2 = 0; k2 < 16; ++k2) {
            const float2 mine = v[bitrev5(k2)];
            const float2 prov = v[bitrev5(31 - k2)];
            float2 got;
            got.x = __shfl_sync(0xffffffffu, prov.x, partner);
            got.y = __shfl_sync(0xffffffffu, prov.y, partner);
            if (lane == 0) got = v[bitrev5((32 - k2) & 31)];
            const float2 s = cadd(mine, got);      // (ar, bi)
            const float2 d = csub(mine, got);      // (br, ai)
            const float2 s2 = cmul2(s, s);         // (ar^2, bi^2)
            mag2[(lane + 32 * k2) ^ swz] = make_float2(fast_sqrt(fmaf(d.y, d.y, s2.x)), fast_sqrt(fmaf(d.x, d.x, s2.y)));
        }
        if (lane == 0) {   // k = 512 pairs with itself
            const float2 x = v[bitrev5(16)];
            mag2[512] = make_float2(2.0f * fabsf(x.x), 2.0f * fabsf(x.y));
        }
        __syncwarp();

        if (MODE == 0) {
            float2 acc = make_float2(0.0f, 0.0f);
            const int ng = p.n_groups;
            for (int g = 0; g < ng; ++g) {
                const int kw = s_mk[g * 32 + lane];
                const float4 w = s_mw[g * 32 + lane];
                const int k0 = kw & 0x3ff;
                const float4 m01 = *reinterpret_cast<const float4 *>(mag2 + k0);          // swizzled first half
                const float4 m23 = *reinterpret_cast<const float4 *>(mag2 + (k0 ^ 2));
                acc = cfma_s(make_float2(m01.x, m01.y), w.x, acc);
                acc = cfma_s(make_float2(m01.z, m01.w), w.y, acc);
                acc = cfma_s(make_float2(m23.x, m23.y), w.z, acc);
                acc = cfma_s(make_float2(m23.z, m23.w), w.w, acc);
                if (kw & kMelFlush) {          // the band ends here: hand over its two sums, start the next band
                    ws.outs[kw >> 16] = acc;
                    acc = make_float2(0.0f, 0.0f);
                }
            }
            __syncwarp();
            // dB + normalise and store: frame A's 80 values, then frame B's
            float *dst = p.out + static_cast<long long>(cur.out_frame) * kMels;
            const int n_out = cur.has_b ? 2 * kMels : kMels;
            const float *of = reinterpret_cast<const float *>(ws.outs);
            for (int i = lane; i < n_out; i += 32) {
                const int band = (i >= kMels) ? i - kMels : i;
                const float m = of[2 * band + (i >= kMels ? 1 : 0)];
                dst[i] = fmaf(p.c1, fast_log2(fmaxf(p.min_level, m)), p.c0);
            }
        } else {
            float *dst = p.out + static_cast<long long>(cur.out_frame) * kBins;
            for (int i = lane; i < kBins; i += 32) dst[i] = 0.5f * mag2[i ^ ((i & 16) >> 3)].x;
            if (cur.has_b)
                for (int i = lane; i < kBins; i += 32) dst[kBins + i] = 0.5f * mag2[i ^ ((i & 16) >> 3)].y;
        }
        __syncwarp();                      // mag2 / outs are free again
        cur = nxt;
    }
}

constexpr size_t kStftSmem = kStftWarps * sizeof(WarpSmem) + 1024 * 8 + 1024 * 4 + kMaxGroups * 32 * (16 + 4);

// ---- reflect padding into the segment layout (np.pad(x, 512, 'reflect'), utils.py:20) ---------
__global__ void pad_reflect_kernel(const float *__restrict__ wav, const int64_t *__restrict__ off,
                                   const int64_t *__restrict__ seg_off, int n, float *__restrict__ wavp)
{
    // one block row per utterance chunk: blockIdx.y = utterance (grid-stride), x covers samples
    for (int u = blockIdx.y; u < n; u += gridDim.y) {
        const int64_t L = off[u + 1] - off[u];
        const int64_t total = L + 2 * kHalfPad;
        const float *src = wav + off[u];
        float *dst = wavp + seg_off[u];
        const int64_t period = (L > 1) ? 2 * (L - 1) : 1;
        for (int64_t j = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; j < total;
             j += static_cast<int64_t>(gridDim.x) * blockDim.x) {
            int64_t m = j - kHalfPad;
            if (m < 0 || m >= L) {
                m %= period;
                if (m < 0) m += period;
                if (m >= L) m = period - m;
            }
            dst[j] = src[m];
        }
    }
}

int pad_reflect(ssfe_ctx *ctx, const float *wav_dev, const int64_t *offsets_host, int n,
                std::vector<int64_t> &seg_off_host)
{
    seg_off_host.resize(n + 1);
    int64_t pos = 0, maxL = 0;
    for (int i = 0; i < n; ++i) {
        const int64_t L = offsets_host[i + 1] - offsets_host[i];
        if (L < 1) return set_error(ctx, SSFE_ERR_TOO_SHORT, "utterance %d is empty", i);
        seg_off_host[i] = pos;
        pos += (L + 2 * kHalfPad + kSegAlign - 1) / kSegAlign * kSegAlign;
        maxL = std::max(maxL, L);
    }
    seg_off_host[n] = pos;
    int rc = ensure(ctx, ctx->ws.wavp, (pos + kSegSlack) * sizeof(float));
    if (rc) return rc;
    int64_t *d_off = upload(ctx, offsets_host, n + 1);
    int64_t *d_seg = upload(ctx, seg_off_host.data(), n + 1);
    if (!d_off || !d_seg) return SSFE_ERR_NOMEM;
    if ((rc = flush_meta(ctx))) return rc;
    dim3 grid(static_cast<unsigned>(std::min<int64_t>((maxL + 1024 + 255) / 256, 64)),
              static_cast<unsigned>(std::min(n, 32768)));
    pad_reflect_kernel<<<grid, 256, 0, ctx->stream>>>(wav_dev, d_off, d_seg, n,
                                                      static_cast<float *>(ctx->ws.wavp.p));
    SSFE_LAUNCHED(ctx);
    return SSFE_OK;
}

int stft_padded(ssfe_ctx *ctx, const float *wavp, const int64_t *seg_off_host, const int64_t *frames_host,
                int n, int mode, float *out)
{
    std::vector<int> pair_off(n + 1);
    std::vector<int64_t> frame_off(n + 1);
    int64_t pairs = 0, frames = 0;
    for (int i = 0; i < n; ++i) {
        pair_off[i] = static_cast<int>(pairs);
        frame_off[i] = frames;
        pairs += (frames_host[i] + 1) / 2;
        frames += frames_host[i];
        if (frames > 0x7fffffff) return set_error(ctx, SSFE_ERR_INVALID, "batch too large (frames)");
    }
    pair_off[n] = static_cast<int>(pairs);
    frame_off[n] = frames;
    if (pairs == 0) return SSFE_OK;
    int rc = ensure(ctx, ctx->ws.tiles, pairs * sizeof(PairInfo));
    if (rc) return rc;

    const int64_t *d_seg = upload(ctx, seg_off_host, n);
    const int *d_pair_off = upload(ctx, pair_off.data(), n + 1);
    const int64_t *d_frame_off = upload(ctx, frame_off.data(), n + 1);
    if (!d_seg || !d_pair_off || !d_frame_off) return SSFE_ERR_NOMEM;
    if ((rc = flush_meta(ctx))) return rc;
    PairInfo *d_pairs = static_cast<PairInfo *>(ctx->ws.tiles.p);
    stft_pairs_kernel<<<static_cast<unsigned>((pairs + 255) / 256), 256, 0, ctx->stream>>>(
        d_seg, d_pair_off, d_frame_off, n, static_cast<int>(pairs), d_pairs);
    SSFE_LAUNCHED(ctx);

    StftParams p;
    p.wavp = wavp;
    p.pairs = d_pairs;
    p.n_pairs = static_cast<int>(pairs);
    p.out = out;
    p.window = ctx->d_window;
    p.tw = ctx->d_tw;
    p.mel_w = ctx->mel.w4;
    p.mel_k = ctx->mel.k0;
    p.n_groups = ctx->mel.n_groups;
    p.min_level = static_cast<float>(ctx->cfg.min_level);
    p.c1 = static_cast<float>(0.2 * std::log10(2.0));
    p.c0 = static_cast<float>((100.0 - ctx->cfg.ref_db) / 100.0);
    const int grid = static_cast<int>(std::min<int64_t>((pairs + kStftWarps - 1) / kStftWarps, ctx->num_sms));
    if (mode == 0)
        stft_mel_kernel<0><<<grid, kStftThreads, kStftSmem, ctx->stream>>>(p);
    else
        stft_mel_kernel<1><<<grid, kStftThreads, kStftSmem, ctx->stream>>>(p);
    SSFE_LAUNCHED(ctx);
    return SSFE_OK;
}

// ---- tables ----------------------------------------------------------------------------------
int init_stft_tables(ssfe_ctx *ctx)
{
    const double pi = 3.14159265358979323846;
    std::vector<float> win(kNfft);
    for (int n = 0; n < kNfft; ++n) win[n] = static_cast<float>(0.5 - 0.5 * std::cos(2.0 * pi * n / kNfft));
    std::vector<float2> tw(1024);
    for (int k1 = 0; k1 < 32; ++k1)
        for (int j = 0; j < 32; ++j) {
            const double a = -2.0 * pi * (double)(j * k1) / 1024.0;
            tw[k1 * 32 + j] = make_float2(static_cast<float>(std::cos(a)), static_cast<float>(std::sin(a)));
        }
    SSFE_CUDA(ctx, cudaMalloc(&ctx->d_window, kNfft * sizeof(float)));
    SSFE_CUDA(ctx, cudaMalloc(&ctx->d_tw, 1024 * sizeof(float2)));
    SSFE_CUDA(ctx, cudaMemcpy(ctx->d_window, win.data(), kNfft * sizeof(float), cudaMemcpyHostToDevice));
    SSFE_CUDA(ctx, cudaMemcpy(ctx->d_tw, tw.data(), 1024 * sizeof(float2), cudaMemcpyHostToDevice));

    // mel: every band is a run of bins [kb, ke]; it is covered by aligned groups of 4 bins (302 groups for the
    // reference basis).  Every lane gets a LIST of groups that strings several bands together: bands are dealt to
    // the 32 lanes longest-first onto the least loaded lane (10 groups per lane instead of the 15 a
    // one-band-per-(slot, lane) layout needed), then a local search reorders bands inside lanes and swaps them
    // between lanes so that, at every step of the list, the 8 lanes that share a 128-bit shared-memory wavefront
    // read from different bank groups (or the same address).
    const float *mb = ctx->mel_basis.data();   // [bin][band]
    int kb[kMels], ke[kMels], ng[kMels];
    for (int m = 0; m < kMels; ++m) {
        kb[m] = -1;
        ke[m] = -1;
        for (int k = 0; k < kBins; ++k)
            if (mb[k * kMels + m] != 0.0f) {
                if (kb[m] < 0) kb[m] = k;
                ke[m] = k;
            }
        ng[m] = (kb[m] < 0) ? 0 : ((ke[m] | 3) + 1 - (kb[m] & ~3)) / 4;
    }
    std::vector<int> order(kMels);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return ng[a] > ng[b]; });
    // (the search below takes ~50 ms; a process creates several contexts - lanes, tests - with the same basis)
    static std::vector<float> cache_basis;
    static std::vector<std::vector<int>> cache_lists;
    static int cache_cost = 0;
    const bool cached = cache_basis == ctx->mel_basis;
    std::vector<std::vector<int>> lists(32);
    int load[32] = {};
    if (cached) {
        lists = cache_lists;
        for (int l = 0; l < 32; ++l)
            for (int m : lists[l]) load[l] += ng[m];
    } else {
        for (int m : order) {
            if (ng[m] == 0) continue;
            int l = 0;
            for (int i = 1; i < 32; ++i)
                if (load[i] < load[l]) l = i;
            lists[l].push_back(m);
            load[l] += ng[m];
        }
    }
    int total = *std::max_element(load, load + 32);
    if (total > kMaxGroups || total == 0)
        return set_error(ctx, SSFE_ERR_INVALID, "mel basis too dense for the fused kernel (%d groups per lane)", total);
    // bank group (16-byte unit mod 8) of the first half of group q = k / 4 in the swizzled magnitude array; the
    // second half sits in the neighbouring unit (^ 1), so both halves see the same conflicts
    auto bank_of = [](int q) { return 2 * (q & 3) + ((q >> 2) & 1); };
    auto lane_groups = [&](int l, int *qs) {          // group index of every step of lane l, -1 = padding
        int n = 0;
        for (int m : lists[l])
            for (int g = 0; g < ng[m]; ++g) qs[n++] = (kb[m] >> 2) + g;
        for (; n < total; ++n) qs[n] = -1;
    };
    auto cost = [&]() {
        int qs[32][kMaxGroups], c = 0;
        for (int l = 0; l < 32; ++l) lane_groups(l, qs[l]);
        for (int st = 0; st < total; ++st)
            for (int qw = 0; qw < 4; ++qw) {
                int worst = 1;
                for (int b = 0; b < 8; ++b) {
                    int distinct = 0, seen[8];
                    for (int l = 8 * qw; l < 8 * qw + 8; ++l) {
                        const int q = qs[l][st] < 0 ? 0 : qs[l][st];     // padding reads group 0
                        if (bank_of(q) != b) continue;
                        bool dup = false;
                        for (int i = 0; i < distinct; ++i) dup = dup || seen[i] == q;
                        if (!dup) seen[distinct++] = q;
                    }
                    worst = std::max(worst, distinct);
                }
                c += worst;
            }
        return c;
    };
    if (cached) {
        ctx->mel.conflict_cost = cache_cost;
    } else {
        uint32_t rng = 12345u;
        auto rnd = [&](int n) { rng = rng * 1664525u + 1013904223u; return static_cast<int>((rng >> 8) % static_cast<uint32_t>(n)); };
        int best = cost();
        for (int it = 0; it < 20000 && best > 4 * total; ++it) {
            const int a = rnd(32), b2 = rnd(32);
            const int kind = rnd(3);
            if (kind == 0) {                                  // swap two whole lanes (changes the quarter-warp mix)
                std::swap(lists[a], lists[b2]);
                std::swap(load[a], load[b2]);
                const int c = cost();
                if (c <= best) best = c;
                else { std::swap(lists[a], lists[b2]); std::swap(load[a], load[b2]); }
            } else if (kind == 1) {                           // reorder the bands inside a lane
                if (lists[a].size() < 2) continue;
                const int i = rnd(static_cast<int>(lists[a].size())), j = rnd(static_cast<int>(lists[a].size()));
                std::swap(lists[a][i], lists[a][j]);
                const int c = cost();
                if (c <= best) best = c;
                else std::swap(lists[a][i], lists[a][j]);
            } else {                                          // exchange one band between two lanes
                if (a == b2 || lists[a].empty() || lists[b2].empty()) continue;
                const int i = rnd(static_cast<int>(lists[a].size())), j = rnd(static_cast<int>(lists[b2].size()));
                const int la = load[a] - ng[lists[a][i]] + ng[lists[b2][j]], lb = load[b2] - ng[lists[b2][j]] + ng[lists[a][i]];
                if (la > total || lb > total) continue;
                std::swap(lists[a][i], lists[b2][j]);
                const int c = cost();
                if (c <= best) { best = c; load[a] = la; load[b2] = lb; }
                else std::swap(lists[a][i], lists[b2][j]);
            }
        }
        ctx->mel.conflict_cost = best;
        if (getenv("SSFE_TRACE_HOST"))
            fprintf(stderr, "[stft] mel layout: %d groups per lane, %d wavefronts per half-load pass (ideal %d)\n", total, best, 4 * total);
        cache_basis = ctx->mel_basis;
        cache_lists = lists;
        cache_cost = best;
    }
    std::vector<float4> w4(static_cast<size_t>(total) * 32, make_float4(0.f, 0.f, 0.f, 0.f));
    std::vector<int> k0(static_cast<size_t>(total) * 32, 0);
    for (int l = 0; l < 32; ++l) {
        int st = 0;
        for (int m : lists[l])
            for (int g = 0; g < ng[m]; ++g, ++st) {
                const int k = (kb[m] & ~3) + 4 * g;
                float wv[4];
                for (int q = 0; q < 4; ++q) wv[q] = (k + q < kBins) ? 0.5f * mb[(k + q) * kMels + m] : 0.0f;   // 1/2 of the frame split, exact
                w4[st * 32 + l] = make_float4(wv[0], wv[1], wv[2], wv[3]);
                int kw = k ^ ((k & 16) >> 3);                  // swizzled address of the first half
                if (g == ng[m] - 1) kw |= kMelFlush | (m << 16);
                k0[st * 32 + l] = kw;
            }
    }
    // a band without any non-zero weight never gets flushed: its output is 0.84 + 0.2 log10(min_level) like the reference's
    for (int m = 0; m < kMels; ++m)
        if (ng[m] == 0) return set_error(ctx, SSFE_ERR_INVALID, "mel band %d of the basis is empty", m);
    ctx->mel.n_groups = total;
    SSFE_CUDA(ctx, cudaMalloc(&ctx->mel.w4, w4.size() * sizeof(float4)));
    SSFE_CUDA(ctx, cudaMalloc(&ctx->mel.k0, k0.size() * sizeof(int)));
    SSFE_CUDA(ctx, cudaMemcpy(ctx->mel.w4, w4.data(), w4.size() * sizeof(float4), cudaMemcpyHostToDevice));
    SSFE_CUDA(ctx, cudaMemcpy(ctx->mel.k0, k0.data(), k0.size() * sizeof(int), cudaMemcpyHostToDevice));
    SSFE_CUDA(ctx, cudaFuncSetAttribute(stft_mel_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(kStftSmem)));
    SSFE_CUDA(ctx, cudaFuncSetAttribute(stft_mel_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(kStftSmem)));
    return SSFE_OK;
}

void free_stft_tables(ssfe_ctx *ctx)
{
    cudaFree(ctx->d_window);
    cudaFree(ctx->d_tw);
    cudaFree(ctx->mel.w4);
    cudaFree(ctx->mel.k0);
    ctx->d_window = nullptr;
    ctx->d_tw = nullptr;
    ctx->mel.w4 = nullptr;
    ctx->mel.k0 = nullptr;
}

}  // namespace ssfe
