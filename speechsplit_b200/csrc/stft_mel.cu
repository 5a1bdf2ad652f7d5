// stft_mel.cu - fused pySTFT -> mel -> dB -> normalise kernel (stages a3+a4+a5).
//
// Replaces, per frame, reference utils.py:18-31 (reflect pad, 1024-sample frames at hop 256,
// periodic Hann, rfft, abs) and make_spect_f0.py:59-61 (dot with the (513,80) mel basis,
// 20*log10(max(1e-5, .)) - 16, (.+100)/100).  One launch covers a whole ragged batch.
//
// Data layout in HBM: the dithered wav of utterance i lives in a *padded segment*
//     [512 reflected | L_i samples | 512 reflected | slack]          (float32, 256 B aligned)
// so frame t of utterance i is the contiguous run  seg_i + 256 t ... + 1024  and a tile of 8
// consecutive frames is one contiguous 11 KiB run: it is staged with a single TMA bulk copy
// (cp.async.bulk + mbarrier), double buffered so the copy of tile n+1 overlaps the FFTs of tile n.
//
// Work split: persistent CTAs of 4 warps; a warp transforms two adjacent frames with one complex
// 1024-point FFT held in registers (fft_core.cuh); |X| of the 513 bins of both frames goes to
// shared memory, and the 941 non-zero mel weights are applied from a per-lane entry list built at
// ssfe_create so that every lane carries ~30 FMAs per frame.  Arithmetic is fp32 FMA on the CUDA
// cores - no tensor cores (BASELINE.json north_star) - and the kernel is bound by FP32 issue, not
// by HBM: 1344 algorithmic bytes per frame against ~900 warp instructions.
#include "common.cuh"
#include "fft_core.cuh"
#include <algorithm>
#include <cmath>
#include <numeric>

namespace ssfe {

constexpr int kTileFrames = 8;                              // frames per CTA iteration
constexpr int kStageFloats = (kTileFrames + 3) * kHop;      // 2816 floats = 11264 B
constexpr int kStftThreads = 128;
constexpr int kWarpTrans = 32 * kTransStride;               // float2 per warp
constexpr int kMaxEnt = 64;
constexpr int kMaxSeg = 8;

struct StftParams {
    const float *wavp;
    const int64_t *seg_off;      // [n]
    const int *tile_off;         // [n+1]
    const int64_t *frame_off;    // [n+1]
    int n_utts, n_tiles;
    float *out;
    const float *window;
    const float2 *tw;
    const int2 *ent;
    const int2 *seg;
    int n_ent, n_seg;
    float min_level, c1, c0;
};

struct TileInfo {
    long long out_frame;     // global frame index of the tile's first frame
    int n_valid;             // frames of this tile that exist
    int pad;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile(
            "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
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

template <int MODE>
__global__ void __launch_bounds__(kStftThreads, 2) stft_mel_kernel(const StftParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *s_stage = reinterpret_cast<float *>(smem_raw);                       // [2][kStageFloats]
    float2 *s_trans = reinterpret_cast<float2 *>(s_stage + 2 * kStageFloats);   // [4][kWarpTrans]
    float2 *s_tw = s_trans + 4 * kWarpTrans;                                    // [1024]
    float *s_win = reinterpret_cast<float *>(s_tw + 1024);                      // [1024]
    int2 *s_ent = reinterpret_cast<int2 *>(s_win + 1024);                       // [n_ent][32]
    int2 *s_seg = s_ent + kMaxEnt * 32;                                         // [n_seg][32]
    float *s_out = reinterpret_cast<float *>(s_seg + kMaxSeg * 32);             // [4][160]
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_out + 4 * 160);            // [2]
    TileInfo *s_info = reinterpret_cast<TileInfo *>(s_bar + 2);                 // [2]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int i = tid; i < 1024; i += kStftThreads) {
        s_tw[i] = p.tw[i];
        s_win[i] = p.window[i];
    }
    if (MODE == 0) {
        for (int i = tid; i < p.n_ent * 32; i += kStftThreads) s_ent[i] = p.ent[i];
        for (int i = tid; i < p.n_seg * 32; i += kStftThreads) s_seg[i] = p.seg[i];
    }
    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issue = [&](int tile, int buf) {   // thread 0 only
        const int u = find_segment(p.tile_off, p.n_utts, tile);
        const int f0 = (tile - p.tile_off[u]) * kTileFrames;
        const int n_frames = static_cast<int>(p.frame_off[u + 1] - p.frame_off[u]);
        TileInfo ti;
        ti.out_frame = p.frame_off[u] + f0;
        ti.n_valid = min(kTileFrames, n_frames - f0);
        ti.pad = 0;
        s_info[buf] = ti;
        mbar_expect_tx(&s_bar[buf], kStageFloats * 4);
        tma_load_1d(s_stage + buf * kStageFloats, p.wavp + p.seg_off[u] + static_cast<int64_t>(f0) * kHop,
                    kStageFloats * 4, &s_bar[buf]);
    };

    if (tid == 0 && static_cast<int>(blockIdx.x) < p.n_tiles) issue(blockIdx.x, 0);

    float2 *trans = s_trans + warp * kWarpTrans;
    float *mA = reinterpret_cast<float *>(trans);      // [520] aliases the transpose buffer
    float *mB = mA + 520;
    float *outs = s_out + warp * 160;

    int it = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        if (tid == 0) {
            const int nt = tile + gridDim.x;
            if (nt < p.n_tiles) issue(nt, buf ^ 1);
        }
        mbar_wait(&s_bar[buf], (it >> 1) & 1);
        const TileInfo ti = s_info[buf];
        const int fa = 2 * warp;
        if (fa < ti.n_valid) {
            const float *xa = s_stage + buf * kStageFloats + fa * kHop;
            const bool has_b = (fa + 1) < ti.n_valid;
            // a missing second frame must be exactly zero: it shares the transform with frame A
            const float wb = has_b ? 1.0f : 0.0f;
            float2 v[32];
#pragma unroll
            for (int m = 0; m < 32; ++m) {
                const int n = lane + 32 * m;
                const float w = s_win[n];
                const float xb = has_b ? xa[n + kHop] : 0.0f;
                v[m] = make_float2(xa[n] * w, xb * (w * wb));
            }
            fft32_dif(v);
#pragma unroll
            for (int r = 0; r < 32; ++r) {
                const int k1 = bitrev5(r);
                float2 y = v[r];
                if (k1 != 0) {
                    const float2 t = s_tw[k1 * 32 + lane];
                    y = make_float2(v[r].x * t.x - v[r].y * t.y, v[r].x * t.y + v[r].y * t.x);
                }
                trans[lane * kTransStride + k1] = y;
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = trans[j * kTransStride + lane];
            __syncwarp();
            fft32_dif(v);

            // separate the two real spectra and take magnitudes (without the common factor 1/2)
            const int partner = (32 - lane) & 31;
#pragma unroll
            for (int k2 = 0; k2 < 16; ++k2) {
                const float2 mine = v[bitrev5(k2)];
                const float2 prov = v[bitrev5(31 - k2)];
                float2 got;
                got.x = __shfl_sync(0xffffffffu, prov.x, partner);
                got.y = __shfl_sync(0xffffffffu, prov.y, partner);
                if (lane == 0) got = v[bitrev5((32 - k2) & 31)];
                const float ar = mine.x + got.x, ai = mine.y - got.y;
                const float br = mine.x - got.x, bi = mine.y + got.y;
                mA[lane + 32 * k2] = fast_sqrt(ar * ar + ai * ai);
                mB[lane + 32 * k2] = fast_sqrt(br * br + bi * bi);
            }
            if (lane == 0) {   // k = 512 pairs with itself
                const float2 x = v[bitrev5(16)];
                mA[512] = 2.0f * fabsf(x.x);
                mB[512] = 2.0f * fabsf(x.y);
            }
            __syncwarp();

            if (MODE == 0) {
                int e = 0;
                for (int s = 0; s < p.n_seg; ++s) {
                    const int2 sg = s_seg[s * 32 + lane];
                    float accA = 0.0f, accB = 0.0f;
                    for (; e < sg.x; ++e) {
                        const int2 en = s_ent[e * 32 + lane];
                        const float w = __int_as_float(en.y);
                        accA = fmaf(w, mA[en.x], accA);
                        accB = fmaf(w, mB[en.x], accB);
                    }
                    if (sg.y >= 0) {
                        outs[sg.y] = fmaf(p.c1, fast_log2(fmaxf(p.min_level, accA)), p.c0);
                        outs[kMels + sg.y] = fmaf(p.c1, fast_log2(fmaxf(p.min_level, accB)), p.c0);
                    }
                }
                __syncwarp();
                float *dst = p.out + (ti.out_frame + fa) * kMels;
                const int n_out = has_b ? 2 * kMels : kMels;
                for (int i = lane; i < n_out; i += 32) dst[i] = outs[i];
            } else {
                float *dst = p.out + (ti.out_frame + fa) * kBins;
                for (int i = lane; i < kBins; i += 32) dst[i] = 0.5f * mA[i];
                if (has_b)
                    for (int i = lane; i < kBins; i += 32) dst[kBins + i] = 0.5f * mB[i];
            }
        }
        __syncthreads();   // stage[buf] and the per-warp buffers are free again
    }
}

constexpr size_t kStftSmem = 2 * kStageFloats * 4 + 4 * kWarpTrans * 8 + 1024 * 8 + 1024 * 4 +
                             kMaxEnt * 32 * 8 + kMaxSeg * 32 * 8 + 4 * 160 * 4 + 2 * 8 + 2 * 16;

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
    std::vector<int> tile_off(n + 1);
    std::vector<int64_t> frame_off(n + 1);
    int64_t tiles = 0, frames = 0;
    for (int i = 0; i < n; ++i) {
        tile_off[i] = static_cast<int>(tiles);
        frame_off[i] = frames;
        tiles += (frames_host[i] + kTileFrames - 1) / kTileFrames;
        frames += frames_host[i];
        if (tiles > 0x7fffffff) return set_error(ctx, SSFE_ERR_INVALID, "batch too large (tiles)");
    }
    tile_off[n] = static_cast<int>(tiles);
    frame_off[n] = frames;
    if (tiles == 0) return SSFE_OK;

    StftParams p;
    p.wavp = wavp;
    p.seg_off = upload(ctx, seg_off_host, n);
    p.tile_off = upload(ctx, tile_off.data(), n + 1);
    p.frame_off = upload(ctx, frame_off.data(), n + 1);
    if (!p.seg_off || !p.tile_off || !p.frame_off) return SSFE_ERR_NOMEM;
    p.n_utts = n;
    p.n_tiles = static_cast<int>(tiles);
    p.out = out;
    p.window = ctx->d_window;
    p.tw = ctx->d_tw;
    p.ent = ctx->mel.ent;
    p.seg = ctx->mel.seg;
    p.n_ent = ctx->mel.n_entries;
    p.n_seg = ctx->mel.n_seg;
    p.min_level = static_cast<float>(ctx->cfg.min_level);
    p.c1 = static_cast<float>(0.2 * std::log10(2.0));
    p.c0 = static_cast<float>((100.0 - ctx->cfg.ref_db) / 100.0);
    const int grid = static_cast<int>(std::min<int64_t>(tiles, 2LL * ctx->num_sms));
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

    // sparse mel: bands -> lanes by longest-processing-time-first on the non-zero count
    const float *mb = ctx->mel_basis.data();   // [bin][band]
    std::vector<std::vector<int>> bins(kMels);
    for (int k = 0; k < kBins; ++k)
        for (int m = 0; m < kMels; ++m)
            if (mb[k * kMels + m] != 0.0f) bins[m].push_back(k);
    std::vector<int> order(kMels);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(),
                     [&](int a, int b) { return bins[a].size() > bins[b].size(); });
    std::vector<std::vector<int>> lane_bands(32);
    std::vector<size_t> load(32, 0);
    for (int m : order) {
        int best = 0;
        for (int l = 1; l < 32; ++l)
            if (load[l] < load[best]) best = l;
        lane_bands[best].push_back(m);
        load[best] += bins[m].size();
    }
    size_t E = 0, S = 0;
    for (int l = 0; l < 32; ++l) {
        E = std::max(E, load[l]);
        S = std::max(S, lane_bands[l].size());
    }
    if (E > kMaxEnt || S > kMaxSeg || E == 0)
        return set_error(ctx, SSFE_ERR_INVALID, "mel basis too dense for the fused kernel (%zu entries/lane, %zu bands/lane)", E, S);
    std::vector<int2> ent(E * 32, make_int2(0, 0)), seg(S * 32, make_int2(0, -1));
    for (int l = 0; l < 32; ++l) {
        int e = 0;
        for (size_t s = 0; s < S; ++s) {
            if (s < lane_bands[l].size()) {
                const int m = lane_bands[l][s];
                for (int k : bins[m]) {
                    const float w = 0.5f * mb[k * kMels + m];   // exact: folds the 1/2 of the frame split
                    int wi;
                    std::memcpy(&wi, &w, 4);
                    ent[e * 32 + l] = make_int2(k, wi);
                    ++e;
                }
                seg[s * 32 + l] = make_int2(e, m);
            } else {
                seg[s * 32 + l] = make_int2(e, -1);
            }
        }
    }
    ctx->mel.n_entries = static_cast<int>(E);
    ctx->mel.n_seg = static_cast<int>(S);
    SSFE_CUDA(ctx, cudaMalloc(&ctx->mel.ent, ent.size() * sizeof(int2)));
    SSFE_CUDA(ctx, cudaMalloc(&ctx->mel.seg, seg.size() * sizeof(int2)));
    SSFE_CUDA(ctx, cudaMemcpy(ctx->mel.ent, ent.data(), ent.size() * sizeof(int2), cudaMemcpyHostToDevice));
    SSFE_CUDA(ctx, cudaMemcpy(ctx->mel.seg, seg.data(), seg.size() * sizeof(int2), cudaMemcpyHostToDevice));
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
    cudaFree(ctx->mel.ent);
    cudaFree(ctx->mel.seg);
    ctx->d_window = nullptr;
    ctx->d_tw = nullptr;
    ctx->mel.ent = nullptr;
    ctx->mel.seg = nullptr;
}

}  // namespace ssfe
