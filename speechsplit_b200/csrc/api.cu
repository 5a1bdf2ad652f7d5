// api.cu - context lifetime, workspace, and the pipeline entry points of libssfe.so.
#include "common.cuh"
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdlib>

static char g_create_error[512] = "";

namespace ssfe {

int set_error(ssfe_ctx *ctx, int code, const char *fmt, ...)
{
    char *dst = ctx ? ctx->err : g_create_error;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(dst, 512, fmt, ap);
    va_end(ap);
    return code;
}

int cuda_fail(ssfe_ctx *ctx, cudaError_t e, const char *what)
{
    return set_error(ctx, SSFE_ERR_CUDA, "CUDA error %d (%s) at %s", static_cast<int>(e), cudaGetErrorString(e), what);
}

int ensure(ssfe_ctx *ctx, DevBuf &b, size_t bytes)
{
    if (bytes <= b.cap) return SSFE_OK;
    if (b.p) {
        SSFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        SSFE_CUDA(ctx, cudaFree(b.p));
        b.p = nullptr;
        b.cap = 0;
    }
    const size_t want = bytes + bytes / 8 + 4096;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        b.p = nullptr;
        return set_error(ctx, SSFE_ERR_NOMEM, "cudaMalloc of %zu bytes failed: %s", want, cudaGetErrorString(e));
    }
    b.cap = want;
    return SSFE_OK;
}

// Metadata goes from the pinned arena to the device arena with a tiny KERNEL that reads the pinned
// pages over PCIe (UVA), not with cudaMemcpyAsync: a copy-engine transfer queues behind whatever the
// engine is already doing, and in ssfe_extract_host that is a 0.5 GB PCM upload - every stage of the
// sub-batch being computed then waited ~9 ms for its few hundred bytes of offsets (measured: 25.8 ms
// per sub-batch instead of 16.4).
__global__ void meta_copy_kernel(uint4 *__restrict__ dst, const uint4 *__restrict__ src, size_t n16)
{
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n16;
         i += static_cast<size_t>(gridDim.x) * blockDim.x)
        dst[i] = src[i];
}

int stage_copy(ssfe_ctx *ctx, void *dst_dev, const void *src_pinned, size_t bytes, cudaStream_t st)
{
    const size_t n16 = (bytes + 15) / 16;
    if (n16 == 0) return SSFE_OK;
    const unsigned grid = static_cast<unsigned>(std::min<size_t>((n16 + 255) / 256, 4 * static_cast<size_t>(ctx->num_sms)));
    meta_copy_kernel<<<grid, 256, 0, st>>>(static_cast<uint4 *>(dst_dev), static_cast<const uint4 *>(src_pinned), n16);
    ctx->launches++;
    return cudaGetLastError() == cudaSuccess ? SSFE_OK : SSFE_ERR_CUDA;
}

void *upload_meta(ssfe_ctx *ctx, const void *host, size_t bytes)
{
    const size_t need = (bytes + 255) / 256 * 256;
    if (need > ctx->meta_cap || ctx->meta_used + need > ctx->meta_cap) {
        if (flush_meta(ctx) != SSFE_OK) return nullptr;        // what is staged goes out before the arena moves or wraps
    }
    if (need > ctx->meta_cap) {
        // grow: the old arena may still hold metadata of kernels that are queued or not even launched yet
        // (earlier uploads of the same call), so it is retired, not freed, until the context goes away
        if (ctx->meta_host) ctx->retired_host.push_back(ctx->meta_host);
        if (ctx->meta_dev) ctx->retired_dev.push_back(ctx->meta_dev);
        ctx->meta_host = nullptr;
        ctx->meta_dev = nullptr;
        // several calls' worth: the arena wraps (with a stream sync) only once every few calls, and a call's
        // uploads (about 250 bytes per utterance in total, 56 in the largest single one) never lap themselves
        const size_t cap = std::max<size_t>(need * 32, 64u << 20);
        if (cudaMallocHost(reinterpret_cast<void **>(&ctx->meta_host), cap) != cudaSuccess ||
            cudaMalloc(reinterpret_cast<void **>(&ctx->meta_dev), cap) != cudaSuccess) {
            set_error(ctx, SSFE_ERR_NOMEM, "metadata arena allocation of %zu bytes failed", cap);
            ctx->meta_cap = 0;
            return nullptr;
        }
        ctx->meta_cap = cap;
        ctx->meta_used = 0;
    }
    if (ctx->meta_used + need > ctx->meta_cap) {
        cudaStreamSynchronize(ctx->stream);     // everything staged so far has been consumed
        ctx->meta_used = 0;
    }
    char *h = ctx->meta_host + ctx->meta_used;
    char *d = ctx->meta_dev + ctx->meta_used;
    memcpy(h, host, bytes);
    if (ctx->meta_pend_bytes == 0) ctx->meta_pend_off = ctx->meta_used;
    ctx->meta_used += need;
    ctx->meta_pend_bytes = ctx->meta_used - ctx->meta_pend_off;
    return d;
}

int flush_meta(ssfe_ctx *ctx)
{
    if (ctx->meta_pend_bytes == 0) return SSFE_OK;
    const size_t off = ctx->meta_pend_off, bytes = ctx->meta_pend_bytes;
    ctx->meta_pend_bytes = 0;
    if (stage_copy(ctx, ctx->meta_dev + off, ctx->meta_host + off, bytes, ctx->stream) != SSFE_OK)
        return set_error(ctx, SSFE_ERR_CUDA, "metadata upload failed");
    return SSFE_OK;
}

void mark(ssfe_ctx *ctx, int boundary)
{
    if (!ctx->timing || boundary < 0 || boundary > ST_COUNT) return;
    if (boundary == 0) {
        ctx->timed_calls++;
        ctx->slot_marks = 0;
    }
    if (ctx->timed_calls == 0) return;
    const int slot = static_cast<int>((ctx->timed_calls - 1) % ssfe_ctx::kTimeRing);
    cudaEvent_t &e = ctx->ev[slot][boundary];
    if (!e) cudaEventCreate(&e);
    cudaEventRecord(e, ctx->stream);
    ctx->slot_marks = boundary + 1;
}

void mark_aux(ssfe_ctx *ctx, int which, cudaStream_t st)
{
    if (!ctx->timing || ctx->timed_calls == 0) return;
    const int slot = static_cast<int>((ctx->timed_calls - 1) % ssfe_ctx::kTimeRing);
    cudaEvent_t &e = ctx->ev_auxr[slot][which];
    if (!e) cudaEventCreate(&e);
    cudaEventRecord(e, st);
}

}  // namespace ssfe

using namespace ssfe;

extern "C" int ssfe_enable_timing(ssfe_ctx *ctx, int on)
{
    if (!ctx) return SSFE_ERR_INVALID;
    ctx->timing = on != 0;
    if (on) ctx->timed_calls = 0;
    return SSFE_OK;
}

extern "C" int ssfe_stage_ms(ssfe_ctx *ctx, float *ms_out, int n)
{
    if (!ctx || !ms_out) return SSFE_ERR_INVALID;
    if (ctx->timed_calls == 0 || ctx->slot_marks < ST_COUNT + 1)
        return set_error(ctx, SSFE_ERR_INVALID, "no timed ssfe_extract call yet");
    const long long calls = std::min<long long>(ctx->timed_calls, ssfe_ctx::kTimeRing);
    const int last = static_cast<int>((ctx->timed_calls - 1) % ssfe_ctx::kTimeRing);
    SSFE_CUDA(ctx, cudaEventSynchronize(ctx->ev[last][ST_COUNT]));
    for (int i = 0; i < n && i < ST_COUNT; ++i) ms_out[i] = 0.0f;
    for (long long c = 0; c < calls; ++c) {
        const int slot = static_cast<int>((ctx->timed_calls - 1 - c) % ssfe_ctx::kTimeRing);
        for (int i = 0; i < n && i < ST_COUNT; ++i) {
            float ms = 0.0f;
            if (i == ST_RAND && ctx->ev_auxr[slot][0] && ctx->ev_auxr[slot][1] &&
                cudaEventSynchronize(ctx->ev_auxr[slot][1]) == cudaSuccess &&
                cudaEventElapsedTime(&ms, ctx->ev_auxr[slot][0], ctx->ev_auxr[slot][1]) == cudaSuccess) {
                ms_out[i] += ms;       // the dither kernel runs on the side stream: its own duration
                continue;
            }
            SSFE_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev[slot][i], ctx->ev[slot][i + 1]));
            ms_out[i] += ms;
        }
    }
    for (int i = 0; i < n && i < ST_COUNT; ++i) ms_out[i] /= static_cast<float>(calls);
    cudaGetLastError();
    return ST_COUNT;
}

extern "C" const char *ssfe_version(void) { return "ssfe 0.1 (sm_100a)"; }

extern "C" const char *ssfe_last_error(const ssfe_ctx *ctx) { return ctx ? ctx->err : g_create_error; }

extern "C" int64_t ssfe_launch_count(const ssfe_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" int64_t ssfe_fixed_length(int64_t n) { return (n % kHop == 0) ? n + 1 : n; }

extern "C" int64_t ssfe_num_frames(int64_t n)
{
    const int64_t lf = ssfe_fixed_length(n);
    return (lf + kNfft - (kNfft - kHop)) / kHop;     // (len + 1024 - 768) // 256, utils.py:22-23
}

extern "C" int ssfe_plan_offsets(const int64_t *sample_offsets, int n, int64_t *fixed, int64_t *frames)
{
    if (!sample_offsets || n < 0) return SSFE_ERR_INVALID;
    int64_t f = 0, t = 0;
    for (int i = 0; i < n; ++i) {
        const int64_t L = sample_offsets[i + 1] - sample_offsets[i];
        if (L < 0) return SSFE_ERR_INVALID;
        if (fixed) fixed[i] = f;
        if (frames) frames[i] = t;
        f += ssfe_fixed_length(L);
        t += ssfe_num_frames(L);
    }
    if (fixed) fixed[n] = f;
    if (frames) frames[n] = t;
    return SSFE_OK;
}

extern "C" int ssfe_create(ssfe_ctx **out, int device, const ssfe_config *cfg)
{
    if (!out || !cfg) return set_error(nullptr, SSFE_ERR_INVALID, "ssfe_create: null argument");
    *out = nullptr;
    if (cfg->sample_rate != kFs || cfg->n_fft != kNfft || cfg->hop != kHop || cfg->n_mels != kMels)
        return set_error(nullptr, SSFE_ERR_INVALID,
                         "ssfe_create: kernels are specialised for sr=16000 n_fft=1024 hop=256 n_mels=80");
    if (!cfg->mel_basis) return set_error(nullptr, SSFE_ERR_INVALID, "ssfe_create: mel_basis is null");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return set_error(nullptr, SSFE_ERR_CUDA, "ssfe_create: no CUDA device (%s); there is no CPU fallback",
                         cudaGetErrorString(e));
    if (device < 0 || device >= count)
        return set_error(nullptr, SSFE_ERR_INVALID, "ssfe_create: device %d out of range (%d devices)", device, count);
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return set_error(nullptr, SSFE_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    ssfe_ctx *ctx = new ssfe_ctx();
    ctx->device = device;
    ctx->cfg = *cfg;
    ctx->err[0] = 0;
    if (const char *e = getenv("SSFE_MT_GO")) ctx->mt_go_at_start = strcmp(e, "start") == 0;   // "stat" = old schedule (A/B hook)
    if (const char *e = getenv("SSFE_HOST_CHUNK_SAMPLES")) {     // test hook: force many small sub-batches
        const long long v = atoll(e);
        if (v > 0) {
            ctx->host_chunk_samples = v;
            ctx->host_chunk_forced = true;
        }
    }
    ctx->mel_basis.assign(cfg->mel_basis, cfg->mel_basis + kBins * kMels);
    ctx->cfg.mel_basis = ctx->mel_basis.data();
    cudaDeviceProp prop;
    int rc = SSFE_OK;
    do {
        if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) { rc = cuda_fail(ctx, e, "cudaGetDeviceProperties"); break; }
        ctx->num_sms = prop.multiProcessorCount;
        if (prop.major < 10) { rc = set_error(ctx, SSFE_ERR_INVALID, "ssfe requires compute capability 10.x (found %d.%d)", prop.major, prop.minor); break; }
        if ((e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess) { rc = cuda_fail(ctx, e, "cudaStreamCreate"); break; }
        if ((e = cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking)) != cudaSuccess) { rc = cuda_fail(ctx, e, "cudaStreamCreate"); break; }
        if ((e = cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking)) != cudaSuccess) { rc = cuda_fail(ctx, e, "cudaStreamCreate"); break; }
        {   // the dither generator is a handful of latency chains on a side stream: when it competes with the big
            // kernels for SM slots it has to win, or the backward filter pass waits for it (2 ms at 1/8 of the corpus)
            int lo_pri = 0, hi_pri = 0;
            cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri);
            if ((e = cudaStreamCreateWithPriority(&ctx->aux, cudaStreamNonBlocking, hi_pri)) != cudaSuccess) { rc = cuda_fail(ctx, e, "cudaStreamCreate"); break; }
        }
        if ((e = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming)) != cudaSuccess) { rc = cuda_fail(ctx, e, "cudaEventCreate"); break; }
        if ((e = cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming)) != cudaSuccess) { rc = cuda_fail(ctx, e, "cudaEventCreate"); break; }
        if ((e = cudaEventCreateWithFlags(&ctx->ev_dith_free, cudaEventDisableTiming)) != cudaSuccess) { rc = cuda_fail(ctx, e, "cudaEventCreate"); break; }
        if ((e = cudaEventCreateWithFlags(&ctx->ev_mt_go, cudaEventDisableTiming)) != cudaSuccess) { rc = cuda_fail(ctx, e, "cudaEventCreate"); break; }
        {
            int lo_pri = 0, hi_pri = 0;
            cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri);
            const char *v = getenv("SSFE_OH_PRIO");
            const int pri = (v && atoi(v) == 0) ? 0 : hi_pri;
            if ((e = cudaStreamCreateWithPriority(&ctx->aux2, cudaStreamNonBlocking, pri)) != cudaSuccess) { rc = cuda_fail(ctx, e, "cudaStreamCreate"); break; }
        }
        if ((e = cudaEventCreateWithFlags(&ctx->ev_oh_go, cudaEventDisableTiming)) != cudaSuccess) { rc = cuda_fail(ctx, e, "cudaEventCreate"); break; }
        if ((e = cudaEventCreateWithFlags(&ctx->ev_oh_done, cudaEventDisableTiming)) != cudaSuccess) { rc = cuda_fail(ctx, e, "cudaEventCreate"); break; }
        if (const char *v = getenv("SSFE_ONEHOT_EARLY")) ctx->onehot_early = atoi(v);
        ctx->stream = ctx->own_stream;
        if ((rc = init_stft_tables(ctx))) break;
        if ((rc = init_filtfilt(ctx))) break;
        if ((rc = init_rapt(ctx))) break;
    } while (0);
    if (rc) {
        snprintf(g_create_error, sizeof(g_create_error), "%s", ctx->err);
        ssfe_destroy(ctx);
        return rc;
    }
    *out = ctx;
    return SSFE_OK;
}

static void free_buf(DevBuf &b)
{
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
}

extern "C" void ssfe_destroy(ssfe_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (ssfe_ctx *&ln : ctx->lane) {
        ssfe_destroy(ln);
        ln = nullptr;
    }
    free_stft_tables(ctx);
    free_filtfilt(ctx);
    free_rapt(ctx);
    Workspace &w = ctx->ws;
    DevBuf *all[] = {&w.wavp, &w.y1, &w.dith, &w.meta_dev, &w.tiles, &w.misc, &w.rapt_ds, &w.rapt_cand,
                     &w.rapt_stat, &w.rapt_f0, &w.dec_map, &w.cand_map, &w.stat_map, &w.filt_map, &w.carry, &w.mt_state, &w.mt_state_aux, &ctx->h_x, &ctx->h_mel, &ctx->h_f0, &ctx->h_bins,
                     &ctx->h_dith};
    for (DevBuf *b : all) free_buf(*b);
    if (ctx->mt_taps) cudaFree(ctx->mt_taps);
    if (ctx->meta_host) cudaFreeHost(ctx->meta_host);
    if (ctx->meta_dev) cudaFree(ctx->meta_dev);
    for (char *q : ctx->retired_host) cudaFreeHost(q);
    for (char *q : ctx->retired_dev) cudaFree(q);
    if (ctx->pin_in) cudaFreeHost(ctx->pin_in);
    if (ctx->pin_out) cudaFreeHost(ctx->pin_out);
    for (auto &row : ctx->ev)
        for (cudaEvent_t e : row)
            if (e) cudaEventDestroy(e);
    if (ctx->aux) cudaStreamDestroy(ctx->aux);
    if (ctx->aux2) cudaStreamDestroy(ctx->aux2);
    for (cudaEvent_t e : {ctx->ev_fork, ctx->ev_join, ctx->ev_dith_free, ctx->ev_mt_go, ctx->ev_oh_go, ctx->ev_oh_done, ctx->aux_free[0], ctx->aux_free[1],
                          ctx->ev_hd_ready[0], ctx->ev_hd_ready[1]})
        if (e) cudaEventDestroy(e);
    for (int i = 0; i < ssfe_ctx::kHostSlots; ++i)
        for (cudaEvent_t e : {ctx->ev_h2d[i], ctx->ev_comp[i], ctx->ev_d2h[i]})
            if (e) cudaEventDestroy(e);
    for (auto &row : ctx->ev_auxr)
        for (cudaEvent_t e : row)
            if (e) cudaEventDestroy(e);
    for (int i = 0; i < 2; ++i) {
        if (ctx->aux_host[i]) cudaFreeHost(ctx->aux_host[i]);
        if (ctx->aux_dev[i]) cudaFree(ctx->aux_dev[i]);
    }
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    if (ctx->copy_in) cudaStreamDestroy(ctx->copy_in);
    if (ctx->copy_out) cudaStreamDestroy(ctx->copy_out);
    delete ctx;
}

extern "C" int ssfe_set_stream(ssfe_ctx *ctx, void *cuda_stream)
{
    if (!ctx) return SSFE_ERR_INVALID;
    SSFE_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
    return SSFE_OK;
}

extern "C" int ssfe_synchronize(ssfe_ctx *ctx)
{
    if (!ctx) return SSFE_ERR_INVALID;
    SSFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SSFE_OK;
}

// ---- stage entry points ---------------------------------------------------------------------
extern "C" int ssfe_filtfilt(ssfe_ctx *ctx, const void *x_dev, int dtype, const int64_t *sample_offsets, int n,
                             double *y_dev)
{
    if (!ctx) return SSFE_ERR_INVALID;
    if (n < 0 || (n > 0 && (!x_dev || !sample_offsets || !y_dev)))
        return set_error(ctx, SSFE_ERR_INVALID, "ssfe_filtfilt: null argument");
    SSFE_CUDA(ctx, cudaSetDevice(ctx->device));
    std::vector<int64_t> fix(n + 1);
    ssfe_plan_offsets(sample_offsets, n, fix.data(), nullptr);
    FiltOut o;
    o.y = y_dev;
    return filtfilt_run(ctx, x_dev, dtype, sample_offsets, fix.data(), n, o);
}

static int stft_common(ssfe_ctx *ctx, const float *wav_dev, const int64_t *offsets, int n, int mode, float *out)
{
    if (!ctx) return SSFE_ERR_INVALID;
    if (n < 0 || (n > 0 && (!wav_dev || !offsets || !out)))
        return set_error(ctx, SSFE_ERR_INVALID, "ssfe_stft: null argument");
    if (n == 0) return SSFE_OK;
    SSFE_CUDA(ctx, cudaSetDevice(ctx->device));
    std::vector<int64_t> seg;
    int rc = pad_reflect(ctx, wav_dev, offsets, n, seg);
    if (rc) return rc;
    std::vector<int64_t> frames(n);
    for (int i = 0; i < n; ++i) frames[i] = (offsets[i + 1] - offsets[i] + kHop) / kHop;   // utils.py:22-23
    return stft_padded(ctx, static_cast<const float *>(ctx->ws.wavp.p), seg.data(), frames.data(), n, mode, out);
}

extern "C" int ssfe_stft_mag(ssfe_ctx *ctx, const float *wav_dev, const int64_t *offsets, int n, float *mag_dev)
{
    return stft_common(ctx, wav_dev, offsets, n, 1, mag_dev);
}

extern "C" int ssfe_stft_mel_db(ssfe_ctx *ctx, const float *wav_dev, const int64_t *offsets, int n, float *mel_dev)
{
    return stft_common(ctx, wav_dev, offsets, n, 0, mel_dev);
}

static int check_ranges(ssfe_ctx *ctx, const float *lo, const float *hi, int n)
{
    for (int i = 0; i < n; ++i) {
        const bool male = (lo[i] == 50.0f && hi[i] == 250.0f), female = (lo[i] == 100.0f && hi[i] == 600.0f);
        if (!male && !female)
            return set_error(ctx, SSFE_ERR_GENDER,
                             "utterance %d: F0 range (%g, %g) is neither male (50, 250) nor female (100, 600)", i,
                             lo[i], hi[i]);
    }
    return SSFE_OK;
}

extern "C" int ssfe_rapt(ssfe_ctx *ctx, const float *wav_dev, const int64_t *offsets, int n, const float *f0_lo,
                         const float *f0_hi, float *f0_dev)
{
    if (!ctx) return SSFE_ERR_INVALID;
    if (n < 0 || (n > 0 && (!wav_dev || !offsets || !f0_lo || !f0_hi || !f0_dev)))
        return set_error(ctx, SSFE_ERR_INVALID, "ssfe_rapt: null argument");
    if (n == 0) return SSFE_OK;
    SSFE_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = check_ranges(ctx, f0_lo, f0_hi, n);
    if (rc) return rc;
    std::vector<int64_t> start(n), len(n), foff(n + 1);
    int64_t t = 0;
    for (int i = 0; i < n; ++i) {
        start[i] = offsets[i];
        len[i] = offsets[i + 1] - offsets[i];
        foff[i] = t;
        t += (len[i] + kHop - 1) / kHop;        // ceil(L / hop): pysptk's output length
    }
    foff[n] = t;
    return rapt_run(ctx, wav_dev, start.data(), len.data(), foff.data(), n, f0_lo, f0_hi, f0_dev);
}

// ---- the whole hot loop (make_spect_f0.py:50-74) ----------------------------------------------
namespace ssfe {
int extract_device(ssfe_ctx *ctx, const ssfe_batch *b, const void *x_dev, int dtype, const ssfe_outputs *o,
                   const std::vector<int64_t> &fix, const std::vector<int64_t> &foff)
{
    const int n = b->n_utts;
    // padded segment layout of the dithered wav
    std::vector<int64_t> seg(n + 1), frames(n), start(n), len(n);
    int64_t pos = 0;
    for (int i = 0; i < n; ++i) {
        const int64_t Lf = fix[i + 1] - fix[i];
        if (Lf < 633)   // get_f0 minimum ((2 * frame_step) + wind_dur) * fs = 632.00002 with float parameters
            return set_error(ctx, SSFE_ERR_TOO_SHORT, "utterance %d: input range too small for analysis by get_f0", i);
        seg[i] = pos;
        pos += (Lf + 2 * kHalfPad + kSegAlign - 1) / kSegAlign * kSegAlign;
        frames[i] = foff[i + 1] - foff[i];
        start[i] = seg[i] + kHalfPad;
        len[i] = Lf;
    }
    seg[n] = pos;
    int rc;
    if ((rc = check_ranges(ctx, b->f0_lo, b->f0_hi, n))) return rc;
    if ((rc = ensure(ctx, ctx->ws.wavp, (pos + kSegSlack) * sizeof(float)))) return rc;
    const bool ext_dith = ctx->ext_dith != nullptr;          // ssfe_extract_host generated it for the whole call
    if (!ext_dith && (rc = ensure(ctx, ctx->ws.dith, fix[n] * sizeof(double)))) return rc;
    float *wavp = static_cast<float *>(ctx->ws.wavp.p);
    double *dith = ext_dith ? static_cast<double *>(const_cast<void *>(ctx->ext_dith)) : static_cast<double *>(ctx->ws.dith.p);

    // the dither stream is independent of the signal until the very last filtfilt kernel: generate
    // it on a side stream while the forward / backward-local passes run
    // Production path: the generator leaves one raw word per sample (4 bytes; the float dither term is formed from
    // its 27 random bits in the consumer, mt_convert.cuh).  The validation paths - sequential filter mode, or a
    // caller asking for the fp64 wav - keep the raw word pairs and convert bit for bit to numpy's doubles.
    const bool dith_f32 = ext_dith || (ctx->cfg.filtfilt_mode != 1 && !o->wav64);
    // Where the side stream may start: by default beside the PREVIOUS call's stationarity kernel (rapt_run records
    // ev_mt_go there), i.e. as early as the dither buffer is free.  For a large batch it makes no difference where
    // the generator's ~2 G warp instructions run - measured on the full corpus: rapt_stat 18.9 + filtfilt 17.8 ms
    // against 13.4 + 23.3 ms when it waits for its own call to start, the same total - but for a small batch (one
    // GPU's share of eight) its 2.9 ms jump + walk latency chain is longer than the forward filter passes it could
    // hide behind, and the head start covers it.  SSFE_MT_GO=start forces the late start (A/B hook).
    if (ctx->mt_go_at_start) SSFE_CUDA(ctx, cudaEventRecord(ctx->ev_mt_go, ctx->stream));
    mark(ctx, ST_RAND);
    if (!ext_dith && (rc = rand_run(ctx, b->spk_seed, b->dither_skip, fix.data(), n, dith, ctx->aux, dith_f32))) return rc;
    mark(ctx, ST_FILTFILT);
    int64_t *d_seg = upload(ctx, seg.data(), n + 1);
    int64_t *d_fix = upload(ctx, fix.data(), n + 1);
    int64_t *d_foff = upload(ctx, foff.data(), n + 1);          // for the F0 post-processing at the end of the call
    if (!d_seg || !d_fix || !d_foff) return SSFE_ERR_NOMEM;
    FiltOut fo;
    fo.dith = dith;
    fo.wavp = wavp;
    fo.seg_off_dev = d_seg;
    fo.wav = o->wav;
    fo.wav64 = o->wav64;
    fo.dith_ready = ext_dith ? ctx->ext_dith_ready : ctx->ev_join;
    fo.dith_raw = !dith_f32;
    fo.dith_f32 = dith_f32;
    if ((rc = filtfilt_run(ctx, x_dev, dtype, b->sample_offsets, fix.data(), n, fo))) return rc;
    mark(ctx, ST_EDGES);
    if ((rc = fill_reflect_edges(ctx, wavp, d_seg, d_fix, n))) return rc;
    mark(ctx, ST_STFT);
    if ((rc = stft_padded(ctx, wavp, seg.data(), frames.data(), n, 0, o->mel))) return rc;
    mark(ctx, ST_RAPT_DEC);

    float *f0_raw = o->f0_raw;
    if (!f0_raw) {
        if ((rc = ensure(ctx, ctx->ws.rapt_f0, foff[n] * sizeof(float)))) return rc;
        f0_raw = static_cast<float *>(ctx->ws.rapt_f0.p);
    }
    // The one-hot output is 257 floats per frame of which 256 are zero whatever the F0 turns out to be (8.4 GB for the
    // bench corpus, 1.5 ms at HBM write speed at the very end of the call): the zeros go out on a side stream beside
    // rapt_cand (rapt_run forks it) and the normalisation kernel drops the ones in at the end.  The overlap is far from
    // free - a full-width zero stream costs the kernel it meets ~1.1 ms, a 32-CTA one at high priority 0.8 ms - so
    // the call gains ~0.5 ms of the 1.2 (measurements in DESIGN.md 4.5; SSFE_ONEHOT_EARLY=0 restores the single kernel).
    ctx->oh_started = false;
    ctx->oh_pending = (o->onehot && ctx->onehot_early > 0) ? o->onehot : nullptr;
    ctx->oh_rows = foff[n];
    rc = onehot_zero_fork(ctx, 1);
    if (!rc) rc = rapt_run(ctx, wavp, start.data(), len.data(), foff.data(), n, b->f0_lo, b->f0_hi, f0_raw);
    ctx->oh_pending = nullptr;
    const bool oh_zeroed = ctx->oh_started;
    mark(ctx, ST_POST);
    // (also when rapt_run failed: nothing of this call may still be writing the caller's buffer once its stream is idle)
    if (oh_zeroed) SSFE_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_oh_done, 0));
    if (rc) return rc;
    rc = f0_post_run(ctx, f0_raw, foff.data(), n, o->f0_norm, nullptr, o->onehot, o->bins, d_foff, oh_zeroed);
    mark(ctx, ST_COUNT);
    return rc;
}
}  // namespace ssfe

static int check_batch(ssfe_ctx *ctx, const ssfe_batch *b)
{
    if (!b || b->n_utts < 0) return set_error(ctx, SSFE_ERR_INVALID, "ssfe_extract: bad batch");
    if (b->n_utts > 0 && (!b->sample_offsets || !b->f0_lo || !b->f0_hi || !b->spk_seed || !b->dither_skip))
        return set_error(ctx, SSFE_ERR_INVALID, "ssfe_extract: batch has null arrays");
    return SSFE_OK;
}

extern "C" int ssfe_extract(ssfe_ctx *ctx, const ssfe_batch *b, const void *x_dev, int dtype, const ssfe_outputs *o)
{
    if (!ctx) return SSFE_ERR_INVALID;
    int rc = check_batch(ctx, b);
    if (rc) return rc;
    if (b->n_utts == 0) return SSFE_OK;
    if (!x_dev || !o || !o->mel || !o->f0_norm) return set_error(ctx, SSFE_ERR_INVALID, "ssfe_extract: null argument");
    SSFE_CUDA(ctx, cudaSetDevice(ctx->device));
    std::vector<int64_t> fix(b->n_utts + 1), foff(b->n_utts + 1);
    ssfe_plan_offsets(b->sample_offsets, b->n_utts, fix.data(), foff.data());
    return extract_device(ctx, b, x_dev, dtype, o, fix, foff);
}

// Host buffers in and out.  The batch is cut into sub-batches (host_chunk_samples, 256 M samples by
// default, ramping up from 1/8 at the start and down to 1/4 at the end) that flow through the PCM upload stream, two
// compute lanes and the download stream: H2D of later sub-batches, the kernels of two sub-batches and
// D2H of an earlier one overlap; input and output slots are four deep.  Every lane generates the
// dither of its own sub-batches (jump-ahead makes any stream position cheap, mt19937.cu).
static int extract_host_pipeline(ssfe_ctx *ctx, const ssfe_batch *b, const void *x_host, int dtype, float *mel_host,
                                 float *f0_norm_host, int64_t *bins_host)
{
    int rc = check_batch(ctx, b);
    if (rc) return rc;
    const int n = b->n_utts;
    if (n == 0) return SSFE_OK;
    if (!x_host || !mel_host || !f0_norm_host) return set_error(ctx, SSFE_ERR_INVALID, "ssfe_extract_host: null argument");
    if (dtype != SSFE_F64 && dtype != SSFE_F32 && dtype != SSFE_I16)
        return set_error(ctx, SSFE_ERR_INVALID, "ssfe_extract_host: unknown dtype %d", dtype);
    SSFE_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t esz = dtype == SSFE_F64 ? 8 : dtype == SSFE_F32 ? 4 : 2;
    std::vector<int64_t> fix(n + 1), foff(n + 1);
    ssfe_plan_offsets(b->sample_offsets, n, fix.data(), foff.data());

    // Sub-batches of ~kChunkSamples, cut preferably where the speaker changes.  The first upload and the
    // last download are not hidden behind anything, so the schedule ramps up (1/4, 1/2, 1, 1, ...) and
    // down again (..., 1/2, 1/4).
    // full-size sub-batch: a sixteenth of the call, between 32 M and 128 M samples.  Measured on the full corpus
    // (2.09 G samples; e2e relative to the box's bare-copy time in the same run): 256 M 1.07, 192 M ... 32 M 1.02 -
    // the kernels of a sub-batch finish one sub-batch time after its upload, so the tail after the LAST upload
    // shrinks with the sub-batch; on one GPU's share of eight (262 M samples) 32 M sub-batches took 15.6 ms, 16 M
    // 16.1, 8 M 19.3 (their fixed latencies add up).  (Round 1's constant 256 M cut that share into 32 / 32 / 134 /
    // 64 M - no pipeline to speak of.)  The test hook SSFE_HOST_CHUNK_SAMPLES overrides it.
    const int64_t total_samples = b->sample_offsets[n] - b->sample_offsets[0];
    const int64_t kChunkSamples = ctx->host_chunk_forced ? ctx->host_chunk_samples
                                                         : std::min<int64_t>(std::max<int64_t>(total_samples / 16, 32LL << 20), 128LL << 20);
    std::vector<int64_t> targets;
    {
        // ramp up (T/8, T/8, T/4, T/4, T/2, T/2), full sub-batches, ramp down (T/2, T/4, T/8, T/16): nothing hides
        // the kernels and the download of the LAST sub-batch, so it is the smallest (measured on the full corpus:
        // the tail after the last upload went from 12.7 ms with a T/4 ending to ~5); a remainder too small to stand
        // alone is merged into the first ramp-down sub-batch
        int64_t remaining = total_samples;
        const int64_t T = std::max<int64_t>(kChunkSamples, 16);
        const int64_t down[4] = {T / 2, T / 4, T / 8, T / 16};
        const int64_t down_sum = down[0] + down[1] + down[2] + down[3];
        if (remaining <= T / 4) {
            targets.push_back(remaining);
        } else if (remaining <= down_sum + T / 8) {
            // small call: halves down to a sixteenth
            int64_t part = remaining / 2;
            for (int k = 0; k < 3 && remaining - part > (16 << 10); ++k) {
                targets.push_back(part);
                remaining -= part;
                part = remaining / 2;
            }
            targets.push_back(remaining);
        } else {
            for (int64_t h : {T / 8, T / 8, T / 4, T / 4, T / 2, T / 2})
                if (remaining - h >= down_sum) {
                    targets.push_back(h);
                    remaining -= h;
                }
            while (remaining - T >= down_sum) {
                targets.push_back(T);
                remaining -= T;
            }
            const int64_t rest = remaining - down_sum;          // 0 <= rest < T
            if (rest >= T / 4) targets.push_back(rest);
            targets.push_back(down[0] + (rest < T / 4 ? rest : 0));
            targets.push_back(down[1]);
            targets.push_back(down[2]);
            targets.push_back(down[3]);
        }
    }
    std::vector<int> cuts{0};
    {
        int start = 0;
        size_t ti = 0;
        while (start < n) {
            const int64_t want = targets[std::min(ti, targets.size() - 1)];
            ++ti;
            int end = start;
            int64_t acc = 0;
            while (end < n && (acc < want || end == start)) {
                acc += b->sample_offsets[end + 1] - b->sample_offsets[end];
                ++end;
            }
            // extend to the end of the current speaker if that is close (keeps stream requests contiguous)
            int ext = end;
            while (ext < n && b->spk_seed[ext] == b->spk_seed[end - 1] && ext - end < 64) ++ext;
            if (ext == n || b->spk_seed[ext] != b->spk_seed[end - 1]) end = ext;
            cuts.push_back(end);
            start = end;
        }
    }
    const int n_chunks = static_cast<int>(cuts.size()) - 1;
    int64_t max_in = 0, max_fr = 0;
    for (int c = 0; c < n_chunks; ++c) {
        max_in = std::max(max_in, b->sample_offsets[cuts[c + 1]] - b->sample_offsets[cuts[c]]);
        max_fr = std::max(max_fr, foff[cuts[c + 1]] - foff[cuts[c]]);
    }
    // Two compute lanes (full contexts of their own: stream, side stream, workspace) take the
    // sub-batches alternately, so that the serial chains inside a sub-batch (filter carries, Viterbi,
    // the tails of 20 kernels) are covered by the other lane's kernels instead of idling the GPU.
    for (int i = 0; i < ssfe_ctx::kHostLanes; ++i)
        if (!ctx->lane[i]) {
            if ((rc = ssfe_create(&ctx->lane[i], ctx->device, &ctx->cfg)))
                return set_error(ctx, rc, "ssfe_extract_host: lane context: %s", ssfe_last_error(nullptr));
        }
    // device slots [x | mel | f0 | bins], kHostSlots deep so that uploads run ahead of both lanes
    constexpr int kS = ssfe_ctx::kHostSlots;
    const size_t in_b = (max_in * esz + 255) / 256 * 256, mel_b = (max_fr * kMels * 4 + 255) / 256 * 256,
                 f0_b = (max_fr * 4 + 255) / 256 * 256, bins_b = bins_host ? (max_fr * 8 + 255) / 256 * 256 : 0;
    const size_t slot_b = in_b + mel_b + f0_b + bins_b;
    const int n_slots = std::min(kS, n_chunks);
    if ((rc = ensure(ctx, ctx->h_x, n_slots * slot_b))) return rc;
    for (int i = 0; i < kS; ++i)
        for (cudaEvent_t *e : {&ctx->ev_h2d[i], &ctx->ev_comp[i], &ctx->ev_d2h[i]})
            if (!*e) SSFE_CUDA(ctx, cudaEventCreateWithFlags(e, cudaEventDisableTiming));

    // nothing here may run ahead of work already queued on the caller's stream
    SSFE_CUDA(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
    SSFE_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_in, ctx->ev_fork, 0));
    SSFE_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_out, ctx->ev_fork, 0));
    for (int i = 0; i < ssfe_ctx::kHostLanes; ++i) SSFE_CUDA(ctx, cudaStreamWaitEvent(ctx->lane[i]->stream, ctx->ev_fork, 0));

    // The dither of the whole call is generated here, once, on this context's side stream: one raw word per sample
    // (mt_convert.cuh) for every utterance, in two groups - the short ramp-up sub-batches first, so that the first
    // backward filter pass does not wait for the whole corpus.  Generating it per sub-batch in the lanes cost every
    // sub-batch its own jump kernel (~1 M warp instructions per stream segment, however short the sub-batch) and
    // made the kernels of a sub-batch wait for a serial segment walk; per call it is 2 + 3.5 ms beside the uploads.
    const bool hoist = ctx->cfg.filtfilt_mode != 1;
    const int n_first = std::min(4, n_chunks);                 // sub-batches of the first group
    if (hoist) {
        if ((rc = ensure(ctx, ctx->h_dith, static_cast<size_t>(fix[n]) * sizeof(uint32_t)))) return rc;
        SSFE_CUDA(ctx, cudaStreamWaitEvent(ctx->aux, ctx->ev_fork, 0));
        for (int g = 0; g < 2; ++g) {
            if (!ctx->ev_hd_ready[g]) SSFE_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_hd_ready[g], cudaEventDisableTiming));
            const int ua = (g == 0) ? 0 : cuts[n_first], ub = (g == 0) ? cuts[n_first] : n;
            if (ub > ua && (rc = rand_run(ctx, b->spk_seed + ua, b->dither_skip + ua, fix.data() + ua, ub - ua,
                                          static_cast<double *>(ctx->h_dith.p), ctx->aux, true)))
                return rc;
            SSFE_CUDA(ctx, cudaEventRecord(ctx->ev_hd_ready[g], ctx->aux));
        }
    }
    const bool trace = getenv("SSFE_TRACE_HOST") != nullptr;
    std::vector<cudaEvent_t> tev;
    auto tmark = [&](cudaStream_t st) {
        if (!trace) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
        tev.push_back(e);
    };
    tmark(ctx->stream);
    const char *src = static_cast<const char *>(x_host);
    for (int c = 0; c < n_chunks; ++c) {
        const int u0 = cuts[c], u1 = cuts[c + 1], m = u1 - u0, slot = c % n_slots;
        ssfe_ctx *ln = ctx->lane[c % ssfe_ctx::kHostLanes];
        char *base = static_cast<char *>(ctx->h_x.p) + slot * slot_b;
        void *d_x = base;
        float *d_mel = reinterpret_cast<float *>(base + in_b);
        float *d_f0 = reinterpret_cast<float *>(base + in_b + mel_b);
        int64_t *d_bins = bins_host ? reinterpret_cast<int64_t *>(base + in_b + mel_b + f0_b) : nullptr;
        const int64_t s0 = b->sample_offsets[u0], ns = b->sample_offsets[u1] - s0;
        const int64_t f0 = foff[u0], nf = foff[u1] - f0;

        // H2D (the slot's input is free once the kernels of sub-batch c - n_slots are done with it)
        if (c >= n_slots) SSFE_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_in, ctx->ev_comp[slot], 0));
        SSFE_CUDA(ctx, cudaMemcpyAsync(d_x, src + s0 * esz, ns * esz, cudaMemcpyHostToDevice, ctx->copy_in));
        SSFE_CUDA(ctx, cudaEventRecord(ctx->ev_h2d[slot], ctx->copy_in));
        tmark(ctx->copy_in);

        // kernels (the slot's outputs of sub-batch c - n_slots must have left for the host)
        SSFE_CUDA(ctx, cudaStreamWaitEvent(ln->stream, ctx->ev_h2d[slot], 0));
        if (c >= n_slots) SSFE_CUDA(ctx, cudaStreamWaitEvent(ln->stream, ctx->ev_d2h[slot], 0));
        std::vector<int64_t> rel(m + 1), cfix(m + 1), cfoff(m + 1);
        for (int i = 0; i <= m; ++i) {
            rel[i] = b->sample_offsets[u0 + i] - s0;
            cfix[i] = fix[u0 + i] - fix[u0];
            cfoff[i] = foff[u0 + i] - f0;
        }
        ssfe_batch cb;
        cb.n_utts = m;
        cb.sample_offsets = rel.data();
        cb.f0_lo = b->f0_lo + u0;
        cb.f0_hi = b->f0_hi + u0;
        cb.spk_seed = b->spk_seed + u0;
        cb.dither_skip = b->dither_skip + u0;
        ssfe_outputs o;
        memset(&o, 0, sizeof(o));
        o.mel = d_mel;
        o.f0_norm = d_f0;
        o.bins = d_bins;
        tmark(ln->stream);
        ln->ext_dith = hoist ? static_cast<const uint32_t *>(ctx->h_dith.p) + fix[u0] : nullptr;
        ln->ext_dith_ready = hoist ? ctx->ev_hd_ready[c < n_first ? 0 : 1] : nullptr;
        if ((rc = extract_device(ln, &cb, d_x, dtype, &o, cfix, cfoff))) return set_error(ctx, rc, "%s", ln->err);
        SSFE_CUDA(ctx, cudaEventRecord(ctx->ev_comp[slot], ln->stream));
        tmark(ln->stream);
        ctx->launches += ln->launches;
        ln->launches = 0;

        // D2H
        SSFE_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_out, ctx->ev_comp[slot], 0));
        SSFE_CUDA(ctx, cudaMemcpyAsync(mel_host + f0 * kMels, d_mel, nf * kMels * sizeof(float), cudaMemcpyDeviceToHost, ctx->copy_out));
        SSFE_CUDA(ctx, cudaMemcpyAsync(f0_norm_host + f0, d_f0, nf * sizeof(float), cudaMemcpyDeviceToHost, ctx->copy_out));
        if (bins_host)
            SSFE_CUDA(ctx, cudaMemcpyAsync(bins_host + f0, d_bins, nf * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->copy_out));
        SSFE_CUDA(ctx, cudaEventRecord(ctx->ev_d2h[slot], ctx->copy_out));
        tmark(ctx->copy_out);
    }
    SSFE_CUDA(ctx, cudaStreamSynchronize(ctx->copy_out));      // every sub-batch's results are on the host
    for (int i = 0; i < ssfe_ctx::kHostLanes; ++i) SSFE_CUDA(ctx, cudaStreamSynchronize(ctx->lane[i]->stream));
    if (trace) {
        for (int c = 0; c < n_chunks; ++c) {
            float a, b2, cc, d;
            cudaEventElapsedTime(&a, tev[0], tev[1 + 4 * c]);
            cudaEventElapsedTime(&b2, tev[0], tev[2 + 4 * c]);
            cudaEventElapsedTime(&cc, tev[0], tev[3 + 4 * c]);
            cudaEventElapsedTime(&d, tev[0], tev[4 + 4 * c]);
            fprintf(stderr, "[host] chunk %d utts %d: h2d done %.2f | comp %.2f -> %.2f (%.2f) | d2h done %.2f\n", c, cuts[c + 1] - cuts[c], a,
                    b2, cc, cc - b2, d);
        }
        for (cudaEvent_t e : tev) cudaEventDestroy(e);
    }
    return SSFE_OK;
}

extern "C" int ssfe_extract_host(ssfe_ctx *ctx, const ssfe_batch *b, const void *x_host, int dtype, float *mel_host,
                                 float *f0_norm_host, int64_t *bins_host)
{
    if (!ctx) return SSFE_ERR_INVALID;
    const int rc = extract_host_pipeline(ctx, b, x_host, dtype, mel_host, f0_norm_host, bins_host);
    if (rc != SSFE_OK) {
        // a sub-batch failed after earlier ones were queued: their uploads read and their downloads write the
        // caller's host buffers, so nothing may be in flight when the error reaches the caller
        cudaStreamSynchronize(ctx->copy_in);
        cudaStreamSynchronize(ctx->aux);
        for (ssfe_ctx *ln : ctx->lane)
            if (ln) cudaStreamSynchronize(ln->stream);
        cudaStreamSynchronize(ctx->copy_out);
    }
    return rc;
}
