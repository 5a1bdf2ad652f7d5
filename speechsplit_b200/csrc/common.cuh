// common.cuh - shared declarations of libssfe.so (context, workspace, helpers).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
#include "../../include/ssfe.h"

namespace ssfe {

constexpr int kFs = 16000;
constexpr int kNfft = 1024;
constexpr int kHop = 256;
constexpr int kBins = 513;
constexpr int kMels = 80;
constexpr int kHalfPad = 512;          // reflect pad of pySTFT (utils.py:20)
constexpr int kSegAlign = 64;          // padded-wav segments start on 256-byte boundaries
constexpr int kSegSlack = 2048;        // floats readable past a segment (tile over-read)
constexpr float kUnvoiced = -1.0e10f;  // make_spect_f0.py:65

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

// grow-only device scratch; growth happens during warm-up only
struct Workspace {
    DevBuf wavp, y1, dith, meta_dev, tiles, misc;
    DevBuf rapt_ds, rapt_cand, rapt_stat, rapt_f0, dec_map, cand_map, stat_map, filt_map;
    DevBuf carry;
    DevBuf mt_state, mt_state_aux;     // segment start states of the dither streams (mt19937.cu)
};

struct MelTables {
    int n_groups = 0;                  // aligned 4-bin groups per lane
    int conflict_cost = 0;             // shared-memory wavefronts per 128-bit load step after the layout search (>= 4 n_groups)
    float4 *w4 = nullptr;              // [n_groups][32] weights (already * 0.5)
    int *k0 = nullptr;                 // [n_groups][32] swizzled address | flush flag | band << 16
};

struct RaptTables;                     // rapt.cu

}  // namespace ssfe

struct ssfe_ctx {
    int device = 0;
    int num_sms = 148;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaStream_t copy_in = nullptr, copy_out = nullptr;      // ssfe_extract_host pipeline
    cudaStream_t aux = nullptr;                               // the dither stream runs beside filtfilt
    cudaStream_t aux2 = nullptr;                              // zero stream of the one-hot output, beside the RAPT kernels
    cudaEvent_t ev_oh_go = nullptr, ev_oh_done = nullptr;
    // where the zero stream starts: 0 = never (one-hot written in one piece at the end), 1 = before the RAPT kernels,
    // 2 = beside rapt_cand, 3 = beside rapt_stat, 4 = beside rapt_dp (SSFE_ONEHOT_EARLY, A/B hook)
    int onehot_early = 2;
    float *oh_pending = nullptr;                              // one-hot buffer of the current call while its zero stream has not started
    int64_t oh_rows = 0;
    bool oh_started = false;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    static constexpr int kHostSlots = 4;                      // device slots of ssfe_extract_host
    cudaEvent_t ev_h2d[kHostSlots] = {}, ev_comp[kHostSlots] = {}, ev_d2h[kHostSlots] = {};
    static constexpr int kHostLanes = 2;                      // compute lanes of ssfe_extract_host
    ssfe_ctx *lane[kHostLanes] = {};                          // (full contexts, created on first use)
    cudaEvent_t ev_dith_free = nullptr;                       // recorded after the kernel that reads `dith`
    cudaEvent_t ev_mt_go = nullptr;                           // recorded where the dither walk may start
    bool mt_go_at_start = false;                              // true: only once its own call has started (A/B hook, api.cu)
    char *aux_host[2] = {nullptr, nullptr}, *aux_dev[2] = {nullptr, nullptr};   // side-stream metadata staging
    size_t aux_cap[2] = {0, 0};
    cudaEvent_t aux_free[2] = {nullptr, nullptr};
    int aux_idx = 0;
    bool mt_attr_set = false;
    // MT19937 jump-ahead tap lists (mt19937.cu): slot + 1 per (level, digit), 0 = not built yet
    int mt_slot[8][3][256] = {};                              // [log2(segment blocks) - 5][level][digit]
    int mt_cnt[8][3][256] = {};
    uint16_t *mt_taps = nullptr;
    int mt_slots_used = 0, mt_slots_cap = 0;
    ssfe_config cfg;
    std::vector<float> mel_basis;      // host copy (513 x 80)
    char err[512];
    int64_t launches = 0;
    // pinned <-> device metadata arena (bump allocated, reset after a stream sync when full)
    char *meta_host = nullptr;
    char *meta_dev = nullptr;
    size_t meta_cap = 0, meta_used = 0;
    size_t meta_pend_off = 0, meta_pend_bytes = 0;            // staged in the pinned arena, not yet copied (flush_meta)
    std::vector<char *> retired_host, retired_dev;            // outgrown arenas, kept alive until destroy
    // tables
    float *d_window = nullptr;         // periodic Hann(1024)
    float2 *d_tw = nullptr;            // [k1][lane] = exp(-2 pi i lane k1 / 1024)
    ssfe::MelTables mel;
    ssfe::RaptTables *rapt = nullptr;
    ssfe::Workspace ws;
    // host staging for ssfe_extract_host
    // optional in-stream stage timing of ssfe_extract (ssfe_enable_timing / ssfe_stage_ms)
    bool timing = false;
    static constexpr int kTimeRing = 64;                      // timed calls kept
    cudaEvent_t ev[kTimeRing][12] = {};                       // stage boundaries of each timed call
    cudaEvent_t ev_auxr[kTimeRing][2] = {};                   // side-stream (dither) start / end
    long long timed_calls = 0;                                // since ssfe_enable_timing(1)
    int slot_marks = 0;
    void *pin_in = nullptr;  size_t pin_in_cap = 0;
    void *pin_out = nullptr; size_t pin_out_cap = 0;
    ssfe::DevBuf h_x, h_mel, h_f0, h_bins;
    // ssfe_extract_host generates the dither of the WHOLE call on the main context's side stream (two groups of
    // sub-batches) and hands every lane a pointer into it (api.cu)
    ssfe::DevBuf h_dith;
    cudaEvent_t ev_hd_ready[2] = {nullptr, nullptr};
    const void *ext_dith = nullptr;                           // lane contexts: dither words of the current sub-batch
    cudaEvent_t ext_dith_ready = nullptr;
    long long host_chunk_samples = 256LL << 20;               // sub-batch size of ssfe_extract_host when forced (test hook)
    bool host_chunk_forced = false;
};

namespace ssfe {

int set_error(ssfe_ctx *ctx, int code, const char *fmt, ...);
// stage boundaries of ssfe_extract, in launch order
enum Stage { ST_RAND = 0, ST_FILTFILT, ST_EDGES, ST_STFT, ST_RAPT_DEC, ST_RAPT_CAND, ST_RAPT_STAT, ST_RAPT_DP,
             ST_POST, ST_COUNT };
void mark(ssfe_ctx *ctx, int boundary);     // records event `boundary` (0..ST_COUNT) when timing is on
void mark_aux(ssfe_ctx *ctx, int which, cudaStream_t st);   // side-stream start (0) / end (1) of the current call
int cuda_fail(ssfe_ctx *ctx, cudaError_t e, const char *what);
int ensure(ssfe_ctx *ctx, DevBuf &b, size_t bytes);
// stages `bytes` of host metadata in the pinned arena and returns the device pointer they will have; the copy
// itself happens in flush_meta(), ONE copy kernel for everything staged since the last flush - a stage function
// uploads all its arrays, flushes once, then launches (SSFE_LAUNCHED refuses a launch with staged data pending)
void *upload_meta(ssfe_ctx *ctx, const void *host, size_t bytes);
int flush_meta(ssfe_ctx *ctx);
// pinned host -> device by a small kernel on `st` (keeps metadata off the copy engines, see api.cu)
int stage_copy(ssfe_ctx *ctx, void *dst_dev, const void *src_pinned, size_t bytes, cudaStream_t st);
template <typename T>
inline T *upload(ssfe_ctx *ctx, const T *host, size_t n)
{
    return static_cast<T *>(upload_meta(ctx, host, n * sizeof(T)));
}

#define SSFE_CUDA(ctx, expr)                                            \
    do {                                                                \
        cudaError_t e__ = (expr);                                       \
        if (e__ != cudaSuccess) return ::ssfe::cuda_fail(ctx, e__, #expr); \
    } while (0)

#define SSFE_LAUNCHED(ctx)                                              \
    do {                                                                \
        (ctx)->launches++;                                              \
        if ((ctx)->meta_pend_bytes)                                     \
            return ::ssfe::set_error(ctx, SSFE_ERR_INVALID, "internal error: kernel launched with unflushed metadata (%s:%d)", __FILE__, __LINE__); \
        cudaError_t e__ = cudaGetLastError();                           \
        if (e__ != cudaSuccess) return ::ssfe::cuda_fail(ctx, e__, "kernel launch"); \
    } while (0)

// map[t] = u for every tile t in [off[u] / unit, off[u+1] / unit): one thread per segment writes the
// lookup that the per-tile kernels would otherwise redo as a 14-step binary search of dependent loads
template <typename T>
__global__ void segment_map_kernel(const T *__restrict__ off, int n, long long unit, int *__restrict__ map)
{
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n) return;
    const long long t1 = static_cast<long long>(off[u + 1]) / unit;
    for (long long t = static_cast<long long>(off[u]) / unit; t < t1; ++t) map[t] = u;
}

// index of the segment containing pos: largest i with off[i] <= pos (off has n+1 entries)
template <typename T>
__device__ __forceinline__ int find_segment(const T *__restrict__ off, int n, T pos)
{
    int lo = 0, hi = n;     // invariant: off[lo] <= pos < off[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (off[mid] <= pos) lo = mid; else hi = mid;
    }
    return lo;
}

// ---- stage implementations (internal, called by api.cu) ------------------------------------
int init_stft_tables(ssfe_ctx *ctx);
void free_stft_tables(ssfe_ctx *ctx);
// wav (f32, unpadded, offsets dev/host) -> reflect-padded layout in ws.wavp; fills seg_off
int pad_reflect(ssfe_ctx *ctx, const float *wav_dev, const int64_t *offsets_host, int n,
                std::vector<int64_t> &seg_off_host);
// fused kernel over an already padded buffer.  mode 0: mel-dB (out [F,80]); 1: magnitude (out [F,513])
int stft_padded(ssfe_ctx *ctx, const float *wavp, const int64_t *seg_off_host,
                const int64_t *frames_host /* [n] frames per utterance */, int n, int mode, float *out);

int init_filtfilt(ssfe_ctx *ctx);
void free_filtfilt(ssfe_ctx *ctx);
// x (dtype) -> y f64 [fixed offsets].  If wavp != nullptr also fuses  wav = 0.96*y + dith  into the
// backward pass and writes it (f32) into the padded layout / wav_out / wav64_out.
struct FiltOut {
    double *y = nullptr;             // filtfilt output (may be null when fused outputs requested)
    const double *dith = nullptr;    // dither stream (fixed offsets); wav = y*0.96 + (U-0.5)*1e-6
    bool dith_raw = false;           // dith holds raw MT19937 word pairs (mt_convert.cuh), not doubles
    bool dith_f32 = false;           // dith holds one raw generator word per sample (4 bytes; mt_a_to_dither_f32)
    float *wavp = nullptr;           // padded layout base
    const int64_t *seg_off_dev = nullptr;
    float *wav = nullptr;            // flat f32 [fixed offsets]
    double *wav64 = nullptr;
    cudaEvent_t dith_ready = nullptr; // waited on right before the kernel that reads `dith`
};
int filtfilt_run(ssfe_ctx *ctx, const void *x_dev, int dtype, const int64_t *in_off_host,
                 const int64_t *fix_off_host, int n, const FiltOut &out);
int fill_reflect_edges(ssfe_ctx *ctx, float *wavp, const int64_t *seg_off_dev,
                       const int64_t *fix_off_dev, int n);

// launch_on: nullptr / ctx->stream = public path (doubles out); ctx->aux = side stream, raw word pairs
// out, ordering handled inside (waits for ev_dith_free and ev_mt_go: starts beside the previous call's
// stationarity kernel).
// dither_f32 (side-stream path only): one raw word per double out (4 bytes) instead of the raw word pair.
int rand_run(ssfe_ctx *ctx, const uint32_t *seeds, const uint64_t *skip, const int64_t *out_off,
             int n, double *u_dev, cudaStream_t launch_on = nullptr, bool dither_f32 = false);

int init_rapt(ssfe_ctx *ctx);
void free_rapt(ssfe_ctx *ctx);
// wav source: either flat f32 (wav_dev + offsets) or padded layout (wavp + seg_off, data at +512)
int rapt_run(ssfe_ctx *ctx, const float *wav_base, const int64_t *start_host /* [n] first sample */,
             const int64_t *len_host /* [n] */, const int64_t *frame_off_host /* [n+1] */, int n,
             const float *f0_lo, const float *f0_hi, float *f0_dev);

// a7+a8 (+a9 when bins/onehot given) over a ragged batch; frame offsets are HOST [n+1]
// onehot_zeroed: the one-hot buffer was already filled with zeros (onehot_zero_start, awaited by the caller):
// only the ones are dropped in, by the normalisation kernel itself
int f0_post_run(ssfe_ctx *ctx, const float *f0_dev, const int64_t *frame_off_host, int n,
                float *f0_norm_dev, float *stats_dev, float *onehot, int64_t *bins,
                const int64_t *frame_off_dev = nullptr, bool onehot_zeroed = false);
// rows x 257 zeros into `onehot` on the side stream aux2, from where ctx->stream stands now; ctx->ev_oh_done is
// recorded behind it.  Returns SSFE_OK with *started = false when the buffer is not 16-byte aligned.
int onehot_zero_start(ssfe_ctx *ctx, float *onehot, int64_t rows, bool *started);
// called by rapt_run at its fork points: starts the pending zero stream if `where` is the configured one
inline int onehot_zero_fork(ssfe_ctx *ctx, int where)
{
    if (!ctx->oh_pending || ctx->onehot_early != where) return SSFE_OK;
    float *buf = ctx->oh_pending;
    ctx->oh_pending = nullptr;
    return onehot_zero_start(ctx, buf, ctx->oh_rows, &ctx->oh_started);
}

}  // namespace ssfe
