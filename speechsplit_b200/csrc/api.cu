// api.cu - context lifetime, workspace, and the pipeline entry points of libssfe.so.
#include "common.cuh"
#include <algorithm>
#include <cmath>
#include <cstdarg>

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

void *upload_meta(ssfe_ctx *ctx, const void *host, size_t bytes)
{
    const size_t need = (bytes + 255) / 256 * 256;
    if (need > ctx->meta_cap) {
        // grow: drain, then reallocate both halves
        cudaStreamSynchronize(ctx->stream);
        if (ctx->meta_host) cudaFreeHost(ctx->meta_host);
        if (ctx->meta_dev) cudaFree(ctx->meta_dev);
        ctx->meta_host = nullptr;
        ctx->meta_dev = nullptr;
        const size_t cap = std::max<size_t>(need * 4, 8u << 20);
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
    if (cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) {
        set_error(ctx, SSFE_ERR_CUDA, "metadata upload failed");
        return nullptr;
    }
    ctx->meta_used += need;
    return d;
}

void mark(ssfe_ctx *ctx, int boundary)
{
    if (!ctx->timing || boundary < 0 || boundary > ST_COUNT) return;
    if (!ctx->ev[boundary]) cudaEventCreate(&ctx->ev[boundary]);
    cudaEventRecord(ctx->ev[boundary], ctx->stream);
    if (boundary + 1 > ctx->n_ev) ctx->n_ev = boundary + 1;
}

}  // namespace ssfe

using namespace ssfe;

extern "C" int ssfe_enable_timing(ssfe_ctx *ctx, int on)
{
    if (!ctx) return SSFE_ERR_INVALID;
    ctx->timing = on != 0;
    ctx->n_ev = 0;
    return SSFE_OK;
}

extern "C" int ssfe_stage_ms(ssfe_ctx *ctx, float *ms_out, int n)
{
    if (!ctx || !ms_out) return SSFE_ERR_INVALID;
    if (ctx->n_ev < ST_COUNT + 1) return set_error(ctx, SSFE_ERR_INVALID, "no timed ssfe_extract call yet");
    SSFE_CUDA(ctx, cudaEventSynchronize(ctx->ev[ST_COUNT]));
    for (int i = 0; i < n && i < ST_COUNT; ++i)
        SSFE_CUDA(ctx, cudaEventElapsedTime(&ms_out[i], ctx->ev[i], ctx->ev[i + 1]));
    // the dither kernel runs on the side stream, overlapped with filtfilt: report its own duration
    if (n > ST_RAND && cudaEventQuery(ctx->ev_aux1) == cudaSuccess)
        cudaEventElapsedTime(&ms_out[ST_RAND], ctx->ev_aux0, ctx->ev_aux1);
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
        if ((e = cudaStreamCreateWithFlags(&ctx->aux, cudaStreamNonBlocking)) != cudaSuccess) { rc = cuda_fail(ctx, e, "cudaStreamCreate"); break; }
        if ((e = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming)) != cudaSuccess) { rc = cuda_fail(ctx, e, "cudaEventCreate"); break; }
        if ((e = cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming)) != cudaSuccess) { rc = cuda_fail(ctx, e, "cudaEventCreate"); break; }
        if ((e = cudaEventCreate(&ctx->ev_aux0)) != cudaSuccess) { rc = cuda_fail(ctx, e, "cudaEventCreate"); break; }
        if ((e = cudaEventCreate(&ctx->ev_aux1)) != cudaSuccess) { rc = cuda_fail(ctx, e, "cudaEventCreate"); break; }
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
    free_stft_tables(ctx);
    free_filtfilt(ctx);
    free_rapt(ctx);
    Workspace &w = ctx->ws;
    DevBuf *all[] = {&w.wavp, &w.y1, &w.dith, &w.meta_dev, &w.tiles, &w.misc, &w.rapt_ds, &w.rapt_cand,
                     &w.rapt_stat, &w.rapt_f0, &w.carry, &ctx->h_x, &ctx->h_mel, &ctx->h_f0, &ctx->h_bins};
    for (DevBuf *b : all) free_buf(*b);
    if (ctx->meta_host) cudaFreeHost(ctx->meta_host);
    if (ctx->meta_dev) cudaFree(ctx->meta_dev);
    if (ctx->pin_in) cudaFreeHost(ctx->pin_in);
    if (ctx->pin_out) cudaFreeHost(ctx->pin_out);
    for (cudaEvent_t e : ctx->ev)
        if (e) cudaEventDestroy(e);
    if (ctx->aux) cudaStreamDestroy(ctx->aux);
    for (cudaEvent_t e : {ctx->ev_fork, ctx->ev_join, ctx->ev_aux0, ctx->ev_aux1})
        if (e) cudaEventDestroy(e);
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
    if ((rc = ensure(ctx, ctx->ws.dith, fix[n] * sizeof(double)))) return rc;
    float *wavp = static_cast<float *>(ctx->ws.wavp.p);
    double *dith = static_cast<double *>(ctx->ws.dith.p);

    // the dither stream is independent of the signal until the very last filtfilt kernel: generate
    // it on a side stream while the forward / backward-local passes run
    mark(ctx, ST_RAND);
    if ((rc = rand_run(ctx, b->spk_seed, b->dither_skip, fix.data(), n, dith, ctx->aux))) return rc;
    mark(ctx, ST_FILTFILT);
    int64_t *d_seg = upload(ctx, seg.data(), n + 1);
    int64_t *d_fix = upload(ctx, fix.data(), n + 1);
    if (!d_seg || !d_fix) return SSFE_ERR_NOMEM;
    FiltOut fo;
    fo.dith = dith;
    fo.wavp = wavp;
    fo.seg_off_dev = d_seg;
    fo.wav = o->wav;
    fo.wav64 = o->wav64;
    fo.dith_ready = ctx->ev_join;
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
    if ((rc = rapt_run(ctx, wavp, start.data(), len.data(), foff.data(), n, b->f0_lo, b->f0_hi, f0_raw))) return rc;
    mark(ctx, ST_POST);
    rc = f0_post_run(ctx, f0_raw, foff.data(), n, o->f0_norm, nullptr, o->onehot, o->bins);
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

extern "C" int ssfe_extract_host(ssfe_ctx *ctx, const ssfe_batch *b, const void *x_host, int dtype, float *mel_host,
                                 float *f0_norm_host, int64_t *bins_host)
{
    if (!ctx) return SSFE_ERR_INVALID;
    int rc = check_batch(ctx, b);
    if (rc) return rc;
    const int n = b->n_utts;
    if (n == 0) return SSFE_OK;
    if (!x_host || !mel_host || !f0_norm_host) return set_error(ctx, SSFE_ERR_INVALID, "ssfe_extract_host: null argument");
    SSFE_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t esz = dtype == SSFE_F64 ? 8 : dtype == SSFE_F32 ? 4 : 2;
    std::vector<int64_t> fix(n + 1), foff(n + 1);
    ssfe_plan_offsets(b->sample_offsets, n, fix.data(), foff.data());
    const int64_t total_in = b->sample_offsets[n] - b->sample_offsets[0];
    const int64_t total_fr = foff[n];
    if ((rc = ensure(ctx, ctx->h_x, total_in * esz))) return rc;
    if ((rc = ensure(ctx, ctx->h_mel, total_fr * kMels * sizeof(float)))) return rc;
    if ((rc = ensure(ctx, ctx->h_f0, total_fr * sizeof(float)))) return rc;
    if (bins_host && (rc = ensure(ctx, ctx->h_bins, total_fr * sizeof(int64_t)))) return rc;
    // offsets relative to the first sample of the batch
    std::vector<int64_t> rel(n + 1);
    for (int i = 0; i <= n; ++i) rel[i] = b->sample_offsets[i] - b->sample_offsets[0];
    ssfe_batch rb = *b;
    rb.sample_offsets = rel.data();
    const char *src = static_cast<const char *>(x_host) + b->sample_offsets[0] * esz;
    SSFE_CUDA(ctx, cudaMemcpyAsync(ctx->h_x.p, src, total_in * esz, cudaMemcpyHostToDevice, ctx->stream));
    ssfe_outputs o;
    memset(&o, 0, sizeof(o));
    o.mel = static_cast<float *>(ctx->h_mel.p);
    o.f0_norm = static_cast<float *>(ctx->h_f0.p);
    o.bins = bins_host ? static_cast<int64_t *>(ctx->h_bins.p) : nullptr;
    if ((rc = extract_device(ctx, &rb, ctx->h_x.p, dtype, &o, fix, foff))) return rc;
    SSFE_CUDA(ctx, cudaMemcpyAsync(mel_host, o.mel, total_fr * kMels * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    SSFE_CUDA(ctx, cudaMemcpyAsync(f0_norm_host, o.f0_norm, total_fr * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    if (bins_host)
        SSFE_CUDA(ctx, cudaMemcpyAsync(bins_host, o.bins, total_fr * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    SSFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SSFE_OK;
}
