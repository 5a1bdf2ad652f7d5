/*
 * ssfe.h - C ABI of libssfe.so: the B200-native (sm_100a) SpeechSplit feature front end.
 *
 * This is the drop-in boundary for ONE hot path of biggytruck/SpeechSplit: the body of the
 * per-utterance loop of make_spect_f0.py (reference make_spect_f0.py:48-74) plus
 * utils.quantize_f0_numpy / quantize_f0_torch (reference utils.py:46-74).  The reference has
 * no FFI of its own (it is pure Python calling scipy / numpy / librosa / pysptk); each entry
 * point below therefore names the reference *call site* it replaces.  The Python package
 * speechsplit_b200 binds these with ctypes and mirrors the reference's function names
 * (see INTEGRATION.md for the stub a maintainer would add to the reference).
 *
 * Conventions
 *   - plain C types only; no exceptions cross the boundary: every call returns SSFE_OK (0) or a
 *     negative ssfe_status, and ssfe_last_error(ctx) holds a human-readable message.
 *   - "dev" pointers are CUDA device pointers on the context's GPU, caller-allocated
 *     (e.g. a torch tensor's data_ptr()); "host" pointers are ordinary host memory.
 *     Small per-utterance metadata (offsets, ranges, seeds) is always passed as HOST arrays.
 *   - a ragged batch is the concatenation of its utterances plus an offsets array [n+1].
 *   - one context per GPU; a context is not thread-safe, different contexts are independent.
 *   - all work is enqueued on the context's stream (ssfe_set_stream); calls are asynchronous
 *     unless stated otherwise.  (Internal side streams - the dither generator, the zero stream of
 *     the one-hot output - are forked from that stream and joined back to it by events: whatever
 *     the caller orders against the context's stream is ordered against all of a call's work.)  There is NO CPU fallback: without a CUDA device ssfe_create fails.
 */
#ifndef SSFE_H_
#define SSFE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ssfe_ctx ssfe_ctx;

typedef enum {
    SSFE_OK = 0,
    SSFE_ERR_INVALID = -1,    /* bad argument / unsupported configuration                      */
    SSFE_ERR_CUDA = -2,       /* a CUDA runtime call failed (message has the CUDA error)       */
    SSFE_ERR_TOO_SHORT = -3,  /* utterance too short: filtfilt needs L > 18
                                 (scipy ValueError), RAPT needs L >= 633 (pysptk ValueError)   */
    SSFE_ERR_RANGE = -4,      /* quantize_f0: value outside [0,1] -> reference AssertionError
                                 (utils.py:52 / :68)                                           */
    SSFE_ERR_NOMEM = -5,
    SSFE_ERR_GENDER = -6      /* f0 range is not one of the reference's (make_spect_f0.py:45) */
} ssfe_status;

typedef enum { SSFE_F32 = 0, SSFE_F64 = 1, SSFE_I16 = 2 } ssfe_dtype;   /* I16 = PCM, x = v/32768 */

/* Hard-coded literals of make_spect_f0.py:15-17,55,60-61,64 gathered in one struct.  The kernels
 * are specialised for sample_rate 16000 / n_fft 1024 / hop 256 / n_mels 80; other values are
 * rejected with SSFE_ERR_INVALID.  b, a, zi and mel_basis are *inputs*: the host side computes
 * them exactly as the reference does (scipy.signal.butter / lfilter_zi; the (513,80) mel basis)
 * so that the GPU uses bit-identical constants. */
typedef struct {
    int32_t sample_rate;      /* 16000 */
    int32_t n_fft;            /* 1024  */
    int32_t hop;              /* 256   */
    int32_t n_mels;           /* 80    */
    double  b[6], a[6];       /* utils.butter_highpass(30, 16000, 5)   (utils.py:10-14)        */
    double  zi[5];            /* scipy.signal.lfilter_zi(b, a) as filtfilt uses it             */
    const float *mel_basis;   /* host, (n_fft/2+1) x n_mels row-major == make_spect_f0.py:15   */
    double  min_level;        /* make_spect_f0.py:16  exp(-100/20*ln10)                        */
    double  ref_db;           /* 16   (make_spect_f0.py:60)                                    */
    double  wav_scale;        /* 0.96 (make_spect_f0.py:55)                                    */
    double  dither_scale;     /* 1e-06 (make_spect_f0.py:55)                                   */
    int32_t filtfilt_mode;    /* 0 = chunked parallel scan (default), 1 = one thread per
                                 utterance, sequential (validation aid; same arithmetic order
                                 as scipy's C loop), 3 = the scan with its local passes walking
                                 the recurrence instead of taking dot products (A/B aid)       */
    int32_t reserved;
} ssfe_config;

/* ---- lifetime ---------------------------------------------------------------------------- */
int  ssfe_create(ssfe_ctx **out, int device, const ssfe_config *cfg);
void ssfe_destroy(ssfe_ctx *ctx);
const char *ssfe_last_error(const ssfe_ctx *ctx);   /* ctx may be NULL: last create() error    */
int  ssfe_set_stream(ssfe_ctx *ctx, void *cuda_stream);  /* NULL -> the context's own stream   */
int  ssfe_synchronize(ssfe_ctx *ctx);
const char *ssfe_version(void);
/* kernels launched by this context since creation (bench.py's gpu_launches) */
int64_t ssfe_launch_count(const ssfe_ctx *ctx);

/* In-stream CUDA-event timing of the stages of ssfe_extract (off by default; no host sync per
 * call).  ssfe_stage_ms waits for the most recent timed ssfe_extract and writes the milliseconds,
 * averaged over the (up to 64) calls since ssfe_enable_timing(ctx, 1), of, in order: rand (MT19937),
 * filtfilt (+ dither combine), reflect edges, fused STFT-mel kernel, RAPT decimate, RAPT
 * candidates, RAPT stationarity, RAPT Viterbi, F0 normalise/quantise.  Returns the stage count. */
int  ssfe_enable_timing(ssfe_ctx *ctx, int on);
int  ssfe_stage_ms(ssfe_ctx *ctx, float *ms_out, int n);

/* ---- geometry (pure host helpers) --------------------------------------------------------- */
/* make_spect_f0.py:52-53: a length that is a multiple of 256 grows by one sample (value 1e-06) */
int64_t ssfe_fixed_length(int64_t n_samples);
/* frames of mel and of F0 for an utterance of n_samples (before fix-up): len(S) == len(f0_rapt),
 * make_spect_f0.py:69 */
int64_t ssfe_num_frames(int64_t n_samples);
/* offsets_out[i] = sum_{j<i} fixed_length(L_j); frame_offsets_out[i] = sum_{j<i} num_frames(L_j) */
int  ssfe_plan_offsets(const int64_t *sample_offsets, int n_utts,
                       int64_t *fixed_offsets_out, int64_t *frame_offsets_out);

/* ---- stage entry points (device data, ragged batches) -------------------------------------- */

/* (a0+a1) make_spect_f0.py:52-54: length fix-up, then scipy.signal.filtfilt(b, a, x).
 * x_dev: concatenated samples of dtype `dtype`, sample_offsets host [n+1].
 * y_dev: float64 [fixed_offsets[n]] (fixed offsets from ssfe_plan_offsets). */
int ssfe_filtfilt(ssfe_ctx *ctx, const void *x_dev, int dtype, const int64_t *sample_offsets,
                  int n_utts, double *y_dev);

/* (a2, random part) numpy RandomState(seed).rand(): for utterance i writes counts[i] uniform
 * doubles taken from the MT19937 stream seeded with seeds[i], after skipping skip[i] doubles
 * (make_spect_f0.py:47,55: the stream of a speaker continues across its files).
 * u_dev: float64 [sum counts], utterance i at out_offsets[i]. */
int ssfe_rand(ssfe_ctx *ctx, const uint32_t *seeds, const uint64_t *skip, const int64_t *out_offsets,
              int n_utts, double *u_dev);

/* Host-only helpers behind ssfe_rand (no GPU needed; exported so that the algebra can be tested on
 * the CPU against numpy).  MT19937 is linear over GF(2): with phi its characteristic polynomial
 * (degree 19937) and g = x^J mod phi, the untempered word sequence obeys w[n+J] = XOR_{g_i=1} w[n+i],
 * which is how the GPU starts a speaker's stream (make_spect_f0.py:47,55) in many places at once.
 *   ssfe_mt_charpoly_terms: number of terms of phi below x^19937 (exponents into exps, up to cap)
 *   ssfe_mt_jump_poly:      x^n_words mod phi as 312 little-endian 64-bit words; 0 on success
 *   ssfe_mt_jump_taps:      exponents of x^(d * 256^level * unit_words) mod phi, ascending; returns
 *                           their number (d = 1..255, level = 0..2), or -1 */
int ssfe_mt_charpoly_terms(int *exps, int cap);
int ssfe_mt_jump_poly(uint64_t n_words, uint64_t *poly312);
int ssfe_mt_jump_taps(uint64_t unit_words, int level, int d, uint16_t *taps, int cap);

/* Host-only helper behind ssfe_filtfilt / ssfe_extract (no GPU needed; exported so that the algebra can be
 * tested on the CPU against scipy).  scipy.signal.filtfilt (make_spect_f0.py:54) runs ONE order-5 DF2T recurrence
 * over (b, a); the chunked scan evaluates the same transfer function, with the same initial condition zi * x[0],
 * as a cascade of one first-order and two second-order sections (csrc/filt_consts.cpp, 113-bit algebra).
 *   sec   [3][5]  b0, b1, b2, a1, a2 of the sections in evaluation order (section 0: b2 = a2 = 0)
 *   zic   [5]     cascade state equivalent to scipy's lfilter_zi vector zi5
 *   m     [25]    (cascade state matrix)^chunk, row-major
 * Returns 0, or a negative code when (b6, a6) does not factor into stable real sections. */
int ssfe_filt_cascade(const double *b6, const double *a6, const double *zi5, int chunk, double *sec, double *zic,
                      double *m);
/* Taps of the scan's local passes: the final cascade state of a chunk entered with a ZERO state is
 *   s[k] = sum_i g[k * chunk + i] * x[i],   g[k][i] = ((cascade state matrix)^(chunk-1-i) * input vector)[k],
 * so the two local passes are five dot products per chunk instead of a walk of the recurrence.  sec15 = the
 * sections ssfe_filt_cascade returned; g: [5][chunk].  Returns 0, or -1 for bad arguments. */
int ssfe_filt_cascade_taps(const double *sec15, int chunk, double *g);
/* Powers of the chunk-to-chunk carry matrix, for the carry kernel's scan over runs of `run` chunks:
 *   out[k][25] = (cascade state matrix)^(chunk * run * 2^k), k = 0 .. n_pow - 1, row-major.  Returns 0 or -1. */
int ssfe_filt_cascade_powers(const double *sec15, int chunk, int run, int n_pow, double *out);

/* (a3) utils.pySTFT(x) (utils.py:18-31) for 1-D inputs: reflect-pad 512, hop 256, periodic
 * Hann(1024), |rfft|.  wav_dev float32 concatenated, offsets host [n+1] (lengths as given, no
 * fix-up).  mag_dev: float32 [total_frames, 513] (frame-major, i.e. the transpose of pySTFT's
 * return value), frames of utterance i = (L_i + 256) / 256. */
int ssfe_stft_mag(ssfe_ctx *ctx, const float *wav_dev, const int64_t *offsets, int n_utts,
                  float *mag_dev);

/* (a3+a4+a5) make_spect_f0.py:58-61 fused: pySTFT -> dot(mel_basis) -> 20*log10(max(min_level,.))
 * - 16 -> (.+100)/100.  NOT clipped (the clip is data_loader.py:113).
 * mel_dev: float32 [total_frames, 80]. */
int ssfe_stft_mel_db(ssfe_ctx *ctx, const float *wav_dev, const int64_t *offsets, int n_utts,
                     float *mel_dev);

/* (a6) make_spect_f0.py:64: pysptk.sptk.rapt(wav.astype(f32)*32768, 16000, 256, min=lo, max=hi,
 * otype=2).  wav_dev float32 (NOT yet scaled by 32768), f0_lo/f0_hi host [n].
 * f0_dev: float32 [sum ceil(L_i/256)] log-F0, unvoiced = -1e10. */
int ssfe_rapt(ssfe_ctx *ctx, const float *wav_dev, const int64_t *offsets, int n_utts,
              const float *f0_lo, const float *f0_hi, float *f0_dev);

/* Diagnostics for the parity tests: the per-frame records the most recent ssfe_rapt / ssfe_extract
 * left behind (frames RAPT analysed, batch order): candidate count incl. the unvoiced one, lags
 * (-1 = unvoiced), local costs, the F0 each candidate would emit, stationarity, rms ratio.  Any
 * pointer may be NULL.  Synchronous; returns the number of frames copied or a negative status. */
int64_t ssfe_rapt_dump(ssfe_ctx *ctx, int64_t max_frames, uint8_t *ncand, int16_t *loc /* [f,20] */,
                       float *mp /* [f,20] */, float *f0cand /* [f,20] */, float *stat, float *rms_ratio);

/* (a7+a8) make_spect_f0.py:65-67 + utils.speaker_normalization (utils.py:35-42): per-utterance
 * float32 mean / std (ddof 0, numpy pairwise order) of the voiced log-F0, then
 * ((f0-mean)/std/4 clipped to [-1,1] + 1)/2 in float64, stored float32; unvoiced stays -1e10.
 * stats_dev: optional float32 [n,2] (mean, std). */
int ssfe_f0_normalize(ssfe_ctx *ctx, const float *f0_dev, const int64_t *frame_offsets, int n_utts,
                      float *f0_norm_dev, float *stats_dev);

/* utils.speaker_normalization with caller-supplied statistics (the function's own signature):
 * index_nonzero_dev uint8 [count]; result float64 [count]. */
int ssfe_speaker_normalization(ssfe_ctx *ctx, const void *f0_dev, int dtype,
                               const uint8_t *index_nonzero_dev, double mean_f0, double std_f0,
                               int64_t count, double *out_dev);

/* (a9) utils.quantize_f0_numpy / quantize_f0_torch (utils.py:46-74) over a flat array of `count`
 * values: uv = x<=0 -> bin 0, else round-half-even(x*num_bins-... see utils.py:53)+1; one-hot.
 * onehot_dev float32 [count, num_bins+1] (may be NULL), bins_dev int64 [count] (may be NULL).
 * check_range != 0 makes the call synchronous and returns SSFE_ERR_RANGE when the reference's
 * assert (utils.py:52) would fire. */
int ssfe_quantize_f0(ssfe_ctx *ctx, const void *x_dev, int dtype, int64_t count, int num_bins,
                     float *onehot_dev, int64_t *bins_dev, int check_range);

/* ---- the whole hot loop -------------------------------------------------------------------- */
typedef struct {
    int32_t n_utts;
    const int64_t  *sample_offsets;   /* host [n+1] into the concatenated input                */
    const float    *f0_lo, *f0_hi;    /* host [n]: (50,250) male / (100,600) female            */
    const uint32_t *spk_seed;         /* host [n]: int(speaker_dir[1:])  (make_spect_f0.py:47) */
    const uint64_t *dither_skip;      /* host [n]: doubles of the speaker's stream consumed by
                                         this speaker's earlier files (sum of their fixed lengths) */
} ssfe_batch;

typedef struct {
    float   *mel;        /* dev f32 [total_frames, 80]   == S.astype(f32)      (:71-72)        */
    float   *f0_norm;    /* dev f32 [total_frames]       == f0_norm.astype(f32) (:73-74)       */
    float   *f0_raw;     /* dev f32 [total_frames] RAPT log-F0, optional (NULL to skip)        */
    float   *onehot;     /* dev f32 [total_frames, 257] quantize_f0_numpy(f0_norm)[0], optional */
    int64_t *bins;       /* dev i64 [total_frames]      quantize_f0_numpy(f0_norm)[1], optional */
    float   *wav;        /* dev f32 [fixed total samples] the dithered wav (:55), optional      */
    double  *wav64;      /* dev f64 same, optional (parity tests)                               */
} ssfe_outputs;

/* make_spect_f0.py:50-74 for a ragged batch whose samples are already in HBM. */
int ssfe_extract(ssfe_ctx *ctx, const ssfe_batch *batch, const void *x_dev, int dtype,
                 const ssfe_outputs *out);

/* Same, HOST buffers in and out: x_host (pageable or pinned), outputs are host pointers
 * (onehot/wav/wav64 not offered here; bins optional).  The call stages through pinned memory,
 * overlaps H2D / kernels / D2H over sub-batches and returns when the results are in host memory.
 * On error nothing is left in flight on the caller's buffers either: the call drains its streams
 * before it returns the code (sub-batches queued before the failing one may have been written). */
int ssfe_extract_host(ssfe_ctx *ctx, const ssfe_batch *batch, const void *x_host, int dtype,
                      float *mel_host, float *f0_norm_host, int64_t *bins_host);

/* ---- "next" row: the on-GPU collator of data_loader.py:101-128 ------ */
/* For item i: crop frames [left[i], left[i]+len_crop[i]) of utterance utt[i], clip mel to [0,1],
 * zero-pad to max_len_pad frames, pad F0 with -1e10, and (optionally) quantize the padded F0 as
 * solver.py:162 does.  All index arrays are host; features are the extract outputs in HBM. */
int ssfe_collate(ssfe_ctx *ctx, const float *mel_dev, const float *f0_norm_dev,
                 const int64_t *frame_offsets, int n_items, const int32_t *utt, const int32_t *left,
                 const int32_t *len_crop, int max_len_pad,
                 float *melsp_dev /* [n,pad,80] */, float *pitch_dev /* [n,pad,1] */,
                 float *onehot_dev /* [n,pad,257] or NULL */, int64_t *bins_dev /* or NULL */);

/* ---- "next" row: InterpLnr.forward in training mode (model.py:380-436) ---- */
/* x_dev float32 [batch, T, C] (solver.py:160: mel and F0 concatenated, T = 192, C = 81), len_seq_dev
 * int64 [batch]; the random draws of model.py:392-393 and :401-404 are inputs, all on the device:
 * scales_dev float32 [batch * max_num_seg] in [0.5, 1.5), len_seg_dev int64 [batch * max_num_seg] in
 * [min_len_seg, max_len_seg).  out_dev float32 [batch, max_len_pad, C]: the surviving resampled frames
 * of every item in segment order, zero-padded / truncated at max_len_pad (pad_sequences, :366-377).
 * One kernel, asynchronous on the context's stream - no host sync (the reference has one at :432).
 * Position i0 + 1 of an item is only read where the reference's masks allow it (i0 < len_seq - 1 <= T - 1). */
int ssfe_interp_lnr(ssfe_ctx *ctx, const float *x_dev, int batch, int T, int C, const int64_t *len_seq_dev,
                    const float *scales_dev, const int64_t *len_seg_dev, int max_num_seg, int max_len_seg,
                    int max_len_pad, float *out_dev);

#ifdef __cplusplus
}
#endif
#endif /* SSFE_H_ */
