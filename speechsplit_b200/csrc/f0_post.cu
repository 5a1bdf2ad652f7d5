// f0_post.cu - stages a7, a8, a9 and the collator.
//
//   a7  make_spect_f0.py:65-66  index_nonzero = f0 != -1e10; np.mean / np.std of the voiced
//       log-F0 of THIS utterance.  The array is float32, so numpy reduces in float32 with its
//       pairwise summation (blocks of <=128 with 8 accumulators, halves rounded down to a
//       multiple of 8); that exact order is reproduced so the statistics are bit-identical.
//   a8  utils.speaker_normalization (utils.py:35-42) in float64, stored float32 (:73-74).
//   a9  utils.quantize_f0_numpy / quantize_f0_torch (utils.py:46-74): uv = x<=0 -> bin 0,
//       else round-half-even(x*num_bins-...)+1, one-hot float32.
//   collator: data_loader.py:101-128 (crop, clip [0,1], zero-pad to 192, F0 pad -1e10) feeding
//       solver.py:160-163.
#include "common.cuh"
#include <cstdlib>
#include <algorithm>

namespace ssfe {

// ---- numpy float32 pairwise sum (numpy/_core/src/umath/loops_utils.h.src semantics) -----------
__device__ float np_pairwise_sum_f32(const float *a, int n)
{
    if (n < 8) {
        float res = 0.0f;
        for (int i = 0; i < n; ++i) res = __fadd_rn(res, a[i]);
        return res;
    } else if (n <= 128) {
        float r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        int i;
        for (i = 8; i < n - (n % 8); i += 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], a[i + j]);
        }
        float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                              __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
        for (; i < n; ++i) res = __fadd_rn(res, a[i]);
        return res;
    } else {
        int n2 = n / 2;
        n2 -= n2 % 8;
        return __fadd_rn(np_pairwise_sum_f32(a, n2), np_pairwise_sum_f32(a + n2, n - n2));
    }
}

// One WARP per utterance: compact the voiced frames, then float32 mean and std (ddof = 0) with numpy's pairwise
// summation reproduced bit for bit - but not by one thread.  (One thread per utterance walked every frame three times
// through dependent global loads: 50 us for a 3 s utterance, 1.0 ms of the 12 ms long-form step for 60 s ones.)
// numpy's recursion is a fixed tree for a given n: halves (n/2 rounded down to a multiple of 8) until a block has
// <= 128 elements; a block is summed by 8 strided accumulators combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) plus a
// sequential tail.  Lane 0 unrolls the recursion once into a list of leaves and a postfix program; the leaves are
// summed four at a time, eight lanes per leaf with one accumulator each; lane 0 runs the program (one add per inner
// node).  The same tree serves the mean and the variance.
constexpr int kStatsWarps = 4;
constexpr int kMaxLeaves = 512;          // leaves hold 57 .. 128 elements: utterances up to ~29 k voiced frames (7.8 min)

struct PairwisePlan {
    int2 leaf[kMaxLeaves];               // (offset, length)
    float res[kMaxLeaves];
    unsigned char prog[2 * kMaxLeaves];  // postfix: 0 = next leaf, 1 = add
    float stack[40];
    int n_leaves, n_prog;
};

// lane 0: the recursion of np_pairwise_sum_f32 as leaves + postfix program; false when the plan does not fit
__device__ bool pairwise_plan(PairwisePlan &pl, int n)
{
    int so[40], sn[40], sp = 0, nl = 0, np = 0;      // explicit stack: sn < 0 marks "add the two results below"
    so[0] = 0; sn[0] = n; sp = 1;
    while (sp > 0) {
        --sp;
        const int off = so[sp], m = sn[sp];
        if (m < 0) {
            pl.prog[np++] = 1;
        } else if (m <= 128) {
            if (nl >= kMaxLeaves) return false;
            pl.leaf[nl++] = make_int2(off, m);
            pl.prog[np++] = 0;
        } else {
            int n2 = m / 2;
            n2 -= n2 % 8;
            if (sp + 3 > 40) return false;
            so[sp] = 0; sn[sp] = -1; ++sp;                   // after both halves: add
            so[sp] = off + n2; sn[sp] = m - n2; ++sp;        // right half (popped second)
            so[sp] = off; sn[sp] = n2; ++sp;                 // left half (popped first)
        }
    }
    pl.n_leaves = nl;
    pl.n_prog = np;
    return true;
}

// all lanes; a[0..n) must be visible to the warp.  Returns the sum in every lane.
__device__ float pairwise_sum_warp(PairwisePlan &pl, const float *a, int n, bool planned, int lane)
{
    float total = 0.0f;
    if (!planned) {                                   // a tree too large for the plan: the serial recursion
        if (lane == 0) total = np_pairwise_sum_f32(a, n);
        return __shfl_sync(0xffffffffu, total, 0);
    }
    const int grp = lane >> 3, j = lane & 7;
    const int nl = pl.n_leaves;
    for (int base = 0; base < nl; base += 4) {
        const int li = base + grp;
        const bool have = li < nl;
        const int2 lf = have ? pl.leaf[li] : make_int2(0, 0);
        const float *q = a + lf.x;
        const int m = lf.y - (lf.y % 8);
        float r = (have && lf.y >= 8) ? q[j] : 0.0f;
        for (int i = 8; i < 128; i += 8)
            if (i < m) r = __fadd_rn(r, q[i + j]);
        float t = __fadd_rn(r, __shfl_down_sync(0xffffffffu, r, 1, 8));
        t = __fadd_rn(t, __shfl_down_sync(0xffffffffu, t, 2, 8));
        float res = __fadd_rn(t, __shfl_down_sync(0xffffffffu, t, 4, 8));
        if (have && j == 0) {
            int i = m;
            if (lf.y < 8) { res = 0.0f; i = 0; }      // only a whole input of < 8 elements: plain left-to-right sum
            for (; i < lf.y; ++i) res = __fadd_rn(res, q[i]);
            pl.res[li] = res;
        }
    }
    __syncwarp();
    if (lane == 0) {
        int sp = 0, next = 0;
        for (int k = 0; k < pl.n_prog; ++k) {
            if (pl.prog[k] == 0) {
                pl.stack[sp++] = pl.res[next++];
            } else {
                --sp;
                pl.stack[sp - 1] = __fadd_rn(pl.stack[sp - 1], pl.stack[sp]);
            }
        }
        total = pl.stack[0];
    }
    __syncwarp();
    return __shfl_sync(0xffffffffu, total, 0);
}

__global__ void __launch_bounds__(kStatsWarps * 32) f0_stats_kernel(const float *__restrict__ f0, const int64_t *__restrict__ frame_off, int n,
                                float *__restrict__ scratch, float *__restrict__ stats /* [n][2] */)
{
    __shared__ PairwisePlan s_plan[kStatsWarps];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int u = blockIdx.x * kStatsWarps + w;
    if (u >= n) return;
    PairwisePlan &pl = s_plan[w];
    const int64_t beg = frame_off[u];
    const int T = static_cast<int>(frame_off[u + 1] - beg);
    float *buf = scratch + beg;
    int nv = 0;
    for (int t0 = 0; t0 < T; t0 += 32) {              // order-preserving compaction of the voiced frames
        const int t = t0 + lane;
        const float v = (t < T) ? f0[beg + t] : kUnvoiced;
        const bool voiced = (t < T) && (v != kUnvoiced);
        const unsigned mask = __ballot_sync(0xffffffffu, voiced);
        if (voiced) buf[nv + __popc(mask & ((1u << lane) - 1u))] = v;
        nv += __popc(mask);
    }
    __syncwarp();
    float mean, sd;
    if (nv == 0) {   // np.mean of an empty selection: nan (RuntimeWarning), make_spect_f0.py:66
        mean = __int_as_float(0x7fc00000);
        sd = mean;
    } else {
        int planned = 0;
        if (lane == 0) planned = pairwise_plan(pl, nv) ? 1 : 0;
        planned = __shfl_sync(0xffffffffu, planned, 0);
        __syncwarp();
        const float cnt = static_cast<float>(nv);
        mean = __fdiv_rn(__fadd_rn(0.0f, pairwise_sum_warp(pl, buf, nv, planned != 0, lane)), cnt);
        for (int i = lane; i < nv; i += 32) {
            const float d = __fsub_rn(buf[i], mean);
            buf[i] = __fmul_rn(d, d);
        }
        __syncwarp();
        const float var = __fdiv_rn(__fadd_rn(0.0f, pairwise_sum_warp(pl, buf, nv, planned != 0, lane)), cnt);
        sd = __fsqrt_rn(var);
    }
    if (lane == 0) {
        stats[2 * u] = mean;
        stats[2 * u + 1] = sd;
    }
}

__device__ __forceinline__ long long quantize_value(double x, int num_bins, bool *bad)
{
    // utils.py:50-55.  NaN: (x<=0) is false and the range assert fails in the reference.
    if (x <= 0.0) return 0;
    if (!(x <= 1.0)) {
        *bad = true;
        return 0;
    }
    return static_cast<long long>(rint(x * static_cast<double>(num_bins - 1))) + 1;
}

// per frame: speaker_normalization with the utterance's stats, then the bin of the stored f32 value
__global__ void f0_norm_quant_kernel(const float *__restrict__ f0, const int64_t *__restrict__ frame_off,
                                     int n, const float *__restrict__ stats, int64_t total,
                                     float *__restrict__ f0_norm, int64_t *__restrict__ bins,
                                     float *__restrict__ onehot_ones)
{
    const int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (t >= total) return;
    const int u = find_segment(frame_off, n, t);
    const float v = f0[t];
    float o = v;
    if (v != kUnvoiced) {
        const double mean = stats[2 * u], sd = stats[2 * u + 1];
        double z = __ddiv_rn(__ddiv_rn(__dsub_rn(static_cast<double>(v), mean), sd), 4.0);   // utils.py:38
        const bool isnan_z = (z != z);                                                        // np.clip keeps nan
        z = fmin(fmax(z, -1.0), 1.0);                                                         // :39
        if (isnan_z) z = __longlong_as_double(0x7ff8000000000000LL);
        o = static_cast<float>(__ddiv_rn(__dadd_rn(z, 1.0), 2.0));                            // :40, saved f32
    }
    if (f0_norm) f0_norm[t] = o;
    if (bins || onehot_ones) {
        bool bad = false;
        const long long b = quantize_value(static_cast<double>(o), 256, &bad);
        if (bins) bins[t] = b;
        if (onehot_ones) onehot_ones[t * 257 + b] = 1.0f;       // the row's zeros are already there (onehot_zero_start)
    }
}

// count float4 of zeros, grid-stride: the zero stream of the one-hot output.  A small persistent grid - it runs
// beside the RAPT kernels and should take the HBM bandwidth they leave idle, not their SM slots.
__global__ void __launch_bounds__(256) zero_stream_kernel(float4 *__restrict__ dst, int64_t count)
{
    const float4 z = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    for (; i + 3 * stride < count; i += 4 * stride) {
        dst[i] = z;
        dst[i + stride] = z;
        dst[i + 2 * stride] = z;
        dst[i + 3 * stride] = z;
    }
    for (; i < count; i += stride) dst[i] = z;
}

template <typename T>
__global__ void quantize_kernel(const T *__restrict__ x, int64_t count, int num_bins,
                                int64_t *__restrict__ bins, int *__restrict__ bins32, int *__restrict__ bad_flag)
{
    const int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (t >= count) return;
    bool bad = false;
    const long long b = quantize_value(static_cast<double>(x[t]), num_bins, &bad);
    if (bad) atomicExch(bad_flag, 1);
    if (bins) bins[t] = b;
    bins32[t] = static_cast<int>(b);
}

// one-hot rows: out[t][c] = (c == bin[t]).  The output is one flat stream of count * width floats
// (width = 257 is odd, so rows are not 16-byte aligned): every thread writes one aligned float4 of that
// stream, whose four elements may straddle two rows.
template <typename B>
__global__ void onehot_kernel(const B *__restrict__ bins, int64_t count, int width, float *__restrict__ out)
{
    const int64_t total = count * width;
    const int64_t n4 = (total + 3) / 4;
    const bool aligned = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    for (int64_t q = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; q < n4;
         q += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t i = 4 * q;
        int64_t t = i / width;
        int c = static_cast<int>(i - t * width);
        int b = static_cast<int>(bins[t]);
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            v[k] = (c == b) ? 1.0f : 0.0f;
            if (++c == width) {
                c = 0;
                ++t;
                b = (t < count) ? static_cast<int>(bins[t]) : -1;
            }
        }
        if (aligned && i + 3 < total) {
            *reinterpret_cast<float4 *>(out + i) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
            for (int k = 0; k < 4 && i + k < total; ++k) out[i + k] = v[k];
        }
    }
}

// The production width (257 = 256 bins + unvoiced): four rows are 1028 floats = 257 aligned float4.  A
// warp takes four rows at a time, streams 257 zero float4 and then - after a __syncwarp, which orders
// the two stores to the same line - lanes 0..3 drop the single 1.0f of their row into the line that is
// still in L2.  A handful of instructions per 4 KB, so the kernel runs at the write bandwidth of HBM
// (the element-wise version needed ~30 instructions per float4 and reached half of it).
template <typename B>
__global__ void onehot257_kernel(const B *__restrict__ bins, int64_t count, float *__restrict__ out)
{
    constexpr int W = 257;
    const int lane = threadIdx.x & 31;
    const int64_t groups = count / 4;
    const int64_t warp0 = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
    const int64_t n_warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    const float4 z = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    for (int64_t gq = warp0; gq < groups; gq += n_warps) {
        const int b = (lane < 4) ? static_cast<int>(bins[4 * gq + lane]) : 0;
        float *base = out + gq * (4 * W);
        float4 *dst = reinterpret_cast<float4 *>(base);
#pragma unroll
        for (int it = 0; it < 8; ++it) dst[it * 32 + lane] = z;
        if (lane == 0) dst[256] = z;
        __syncwarp();
        if (lane < 4) base[lane * W + b] = 1.0f;
    }
}

// rows left over by onehot257_kernel (count % 4) or any other width / alignment
template <typename B>
static void launch_onehot(cudaStream_t st, const B *bins, int64_t count, int width, float *out)
{
    int64_t done = 0;
    if (width == 257 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 && count >= 4) {
        const int64_t groups = count / 4;
        const unsigned grid = static_cast<unsigned>(std::min<int64_t>((groups + 7) / 8, 148LL * 32));
        onehot257_kernel<B><<<grid, 256, 0, st>>>(bins, count, out);
        done = groups * 4;
    }
    if (done < count) {
        const int64_t work = ((count - done) * width + 3) / 4;
        const unsigned grid = static_cast<unsigned>(std::min<int64_t>((work + 255) / 256, 148LL * 64));
        onehot_kernel<B><<<grid, 256, 0, st>>>(bins + done, count - done, width, out + done * width);
    }
}

template <typename T>
__global__ void speaker_norm_kernel(const T *__restrict__ f0, const uint8_t *__restrict__ nz, double mean,
                                    double sd, int64_t count, double *__restrict__ out)
{
    const int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (t >= count) return;
    double v = static_cast<double>(f0[t]);
    if (nz[t]) {
        double z = __ddiv_rn(__ddiv_rn(__dsub_rn(v, mean), sd), 4.0);
        const bool isnan_z = (z != z);
        z = fmin(fmax(z, -1.0), 1.0);
        if (isnan_z) z = __longlong_as_double(0x7ff8000000000000LL);
        v = __ddiv_rn(__dadd_rn(z, 1.0), 2.0);
    }
    out[t] = v;
}

}  // namespace ssfe

using namespace ssfe;

static int64_t grid_for(int64_t work, int threads) { return (work + threads - 1) / threads; }

namespace ssfe {
int f0_post_run(ssfe_ctx *ctx, const float *f0_dev, const int64_t *frame_off_host, int n, float *f0_norm_dev,
                float *stats_dev, float *onehot, int64_t *bins, const int64_t *frame_off_dev, bool onehot_zeroed)
{
    const int64_t total = frame_off_host[n];
    if (n == 0 || total == 0) return SSFE_OK;
    int rc = ensure(ctx, ctx->ws.misc, total * sizeof(float) + static_cast<size_t>(n) * 2 * sizeof(float) +
                                           total * sizeof(int64_t));
    if (rc) return rc;
    float *scratch = static_cast<float *>(ctx->ws.misc.p);
    float *stats = stats_dev ? stats_dev : scratch + total;
    int64_t *tmp_bins = reinterpret_cast<int64_t *>(
        reinterpret_cast<char *>(ctx->ws.misc.p) + (total * sizeof(float) + static_cast<size_t>(n) * 2 * sizeof(float) + 7) / 8 * 8);
    const int64_t *d_off = frame_off_dev;       // ssfe_extract uploaded it with the first stage's metadata
    if (!d_off) {
        d_off = upload(ctx, frame_off_host, n + 1);
        if (!d_off) return SSFE_ERR_NOMEM;
        if ((rc = flush_meta(ctx))) return rc;
    }
    f0_stats_kernel<<<static_cast<unsigned>(grid_for(n, kStatsWarps)), kStatsWarps * 32, 0, ctx->stream>>>(f0_dev, d_off, n, scratch, stats);
    SSFE_LAUNCHED(ctx);
    int64_t *use_bins = bins ? bins : ((onehot && !onehot_zeroed) ? tmp_bins : nullptr);
    f0_norm_quant_kernel<<<static_cast<unsigned>(grid_for(total, 256)), 256, 0, ctx->stream>>>(
        f0_dev, d_off, n, stats, total, f0_norm_dev, use_bins, (onehot && onehot_zeroed) ? onehot : nullptr);
    SSFE_LAUNCHED(ctx);
    if (onehot && !onehot_zeroed) {
        launch_onehot<int64_t>(ctx->stream, use_bins, total, 257, onehot);
        SSFE_LAUNCHED(ctx);
    }
    return SSFE_OK;
}

int onehot_zero_start(ssfe_ctx *ctx, float *onehot, int64_t rows, bool *started)
{
    *started = false;
    const int64_t total = rows * 257;
    if (total == 0 || (reinterpret_cast<uintptr_t>(onehot) & 15) != 0) return SSFE_OK;
    SSFE_CUDA(ctx, cudaEventRecord(ctx->ev_oh_go, ctx->stream));       // not ahead of anything queued before this call
    SSFE_CUDA(ctx, cudaStreamWaitEvent(ctx->aux2, ctx->ev_oh_go, 0));
    const int64_t n4 = total / 4;
    if (n4 > 0) {
        static const int ctas = getenv("SSFE_OH_GRID") ? atoi(getenv("SSFE_OH_GRID")) : 32;
        const unsigned grid = static_cast<unsigned>(std::min<int64_t>((n4 + 1023) / 1024, ctas));
        zero_stream_kernel<<<grid, 256, 0, ctx->aux2>>>(reinterpret_cast<float4 *>(onehot), n4);
        SSFE_LAUNCHED(ctx);
    }
    if (total > 4 * n4)                                                // the last one to three floats
        SSFE_CUDA(ctx, cudaMemsetAsync(onehot + 4 * n4, 0, (total - 4 * n4) * sizeof(float), ctx->aux2));
    SSFE_CUDA(ctx, cudaEventRecord(ctx->ev_oh_done, ctx->aux2));
    *started = true;
    return SSFE_OK;
}
}  // namespace ssfe

extern "C" int ssfe_f0_normalize(ssfe_ctx *ctx, const float *f0_dev, const int64_t *frame_offsets, int n_utts,
                                 float *f0_norm_dev, float *stats_dev)
{
    if (!ctx) return SSFE_ERR_INVALID;
    if (!f0_dev || !frame_offsets || !f0_norm_dev || n_utts < 0)
        return set_error(ctx, SSFE_ERR_INVALID, "ssfe_f0_normalize: null argument");
    return f0_post_run(ctx, f0_dev, frame_offsets, n_utts, f0_norm_dev, stats_dev, nullptr, nullptr);
}

extern "C" int ssfe_speaker_normalization(ssfe_ctx *ctx, const void *f0_dev, int dtype,
                                          const uint8_t *index_nonzero_dev, double mean_f0, double std_f0,
                                          int64_t count, double *out_dev)
{
    if (!ctx) return SSFE_ERR_INVALID;
    if (count == 0) return SSFE_OK;
    if (!f0_dev || !index_nonzero_dev || !out_dev)
        return set_error(ctx, SSFE_ERR_INVALID, "ssfe_speaker_normalization: null argument");
    const unsigned grid = static_cast<unsigned>(grid_for(count, 256));
    if (dtype == SSFE_F32)
        speaker_norm_kernel<float><<<grid, 256, 0, ctx->stream>>>(static_cast<const float *>(f0_dev),
                                                                  index_nonzero_dev, mean_f0, std_f0, count, out_dev);
    else if (dtype == SSFE_F64)
        speaker_norm_kernel<double><<<grid, 256, 0, ctx->stream>>>(static_cast<const double *>(f0_dev),
                                                                   index_nonzero_dev, mean_f0, std_f0, count, out_dev);
    else
        return set_error(ctx, SSFE_ERR_INVALID, "ssfe_speaker_normalization: dtype must be f32 or f64");
    SSFE_LAUNCHED(ctx);
    return SSFE_OK;
}

extern "C" int ssfe_quantize_f0(ssfe_ctx *ctx, const void *x_dev, int dtype, int64_t count, int num_bins,
                                float *onehot_dev, int64_t *bins_dev, int check_range)
{
    if (!ctx) return SSFE_ERR_INVALID;
    if (count == 0) return SSFE_OK;
    if (!x_dev || num_bins < 1) return set_error(ctx, SSFE_ERR_INVALID, "ssfe_quantize_f0: bad argument");
    int rc = ensure(ctx, ctx->ws.misc, count * sizeof(int) + 64);
    if (rc) return rc;
    int *flag = static_cast<int *>(ctx->ws.misc.p);
    int *bins32 = flag + 16;
    SSFE_CUDA(ctx, cudaMemsetAsync(flag, 0, sizeof(int), ctx->stream));
    const unsigned grid = static_cast<unsigned>(grid_for(count, 256));
    if (dtype == SSFE_F32)
        quantize_kernel<float><<<grid, 256, 0, ctx->stream>>>(static_cast<const float *>(x_dev), count, num_bins,
                                                              bins_dev, bins32, flag);
    else if (dtype == SSFE_F64)
        quantize_kernel<double><<<grid, 256, 0, ctx->stream>>>(static_cast<const double *>(x_dev), count, num_bins,
                                                               bins_dev, bins32, flag);
    else
        return set_error(ctx, SSFE_ERR_INVALID, "ssfe_quantize_f0: dtype must be f32 or f64");
    SSFE_LAUNCHED(ctx);
    if (check_range) {
        int h = 0;
        SSFE_CUDA(ctx, cudaMemcpyAsync(&h, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        SSFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (h) return set_error(ctx, SSFE_ERR_RANGE, "quantize_f0: value outside [0, 1] (utils.py:52)");
    }
    if (onehot_dev) {
        launch_onehot<int>(ctx->stream, bins32, count, num_bins + 1, onehot_dev);
        SSFE_LAUNCHED(ctx);
    }
    return SSFE_OK;
}

// ---- collator (data_loader.py:101-128) --------------------------------------------------------
__global__ void collate_kernel(const float *__restrict__ mel, const float *__restrict__ f0n,
                               const int64_t *__restrict__ frame_off, const int *__restrict__ utt,
                               const int *__restrict__ left, const int *__restrict__ len_crop, int max_len_pad,
                               float *__restrict__ melsp, float *__restrict__ pitch, float *__restrict__ onehot,
                               int64_t *__restrict__ bins)
{
    const int item = blockIdx.y;
    const int row = blockIdx.x;                 // padded frame index
    const int tid = threadIdx.x;
    const int u = utt[item];
    const int len = len_crop[item];
    const bool live = row < len;
    const int64_t src = frame_off[u] + left[item] + row;
    const int64_t dst = static_cast<int64_t>(item) * max_len_pad + row;
    if (tid < kMels) {
        float v = 0.0f;
        if (live) v = fminf(fmaxf(mel[src * kMels + tid], 0.0f), 1.0f);   // np.clip(a, 0, 1), :113
        melsp[dst * kMels + tid] = v;
    }
    const float p = live ? f0n[src] : kUnvoiced;                          // :116
    if (tid == 0) pitch[dst] = p;
    if (onehot || bins) {
        bool bad = false;
        const int b = static_cast<int>(quantize_value(static_cast<double>(p), 256, &bad));
        if (bins && tid == 0) bins[dst] = b;
        if (onehot)
            for (int c = tid; c < 257; c += blockDim.x) onehot[dst * 257 + c] = (c == b) ? 1.0f : 0.0f;
    }
}

extern "C" int ssfe_collate(ssfe_ctx *ctx, const float *mel_dev, const float *f0_norm_dev,
                            const int64_t *frame_offsets, int n_items, const int32_t *utt, const int32_t *left,
                            const int32_t *len_crop, int max_len_pad, float *melsp_dev, float *pitch_dev,
                            float *onehot_dev, int64_t *bins_dev)
{
    if (!ctx) return SSFE_ERR_INVALID;
    if (n_items == 0) return SSFE_OK;
    if (!mel_dev || !f0_norm_dev || !frame_offsets || !utt || !left || !len_crop || !melsp_dev || !pitch_dev ||
        max_len_pad < 1)
        return set_error(ctx, SSFE_ERR_INVALID, "ssfe_collate: bad argument");
    int max_u = 0;
    for (int i = 0; i < n_items; ++i) {
        if (utt[i] < 0 || left[i] < 0 || len_crop[i] < 0 || len_crop[i] > max_len_pad)
            return set_error(ctx, SSFE_ERR_INVALID, "ssfe_collate: item %d has a bad crop", i);
        max_u = std::max(max_u, utt[i]);
    }
    for (int i = 0; i < n_items; ++i) {
        const int64_t T = frame_offsets[utt[i] + 1] - frame_offsets[utt[i]];
        if (left[i] + len_crop[i] > T)
            return set_error(ctx, SSFE_ERR_INVALID, "ssfe_collate: crop of item %d exceeds utterance", i);
    }
    int64_t *d_off = upload(ctx, frame_offsets, max_u + 2);
    int *d_utt = upload(ctx, utt, n_items), *d_left = upload(ctx, left, n_items),
        *d_len = upload(ctx, len_crop, n_items);
    if (!d_off || !d_utt || !d_left || !d_len) return SSFE_ERR_NOMEM;
    if (int rcf = flush_meta(ctx)) return rcf;
    dim3 grid(max_len_pad, n_items);
    collate_kernel<<<grid, 96, 0, ctx->stream>>>(mel_dev, f0_norm_dev, d_off, d_utt, d_left, d_len, max_len_pad,
                                                 melsp_dev, pitch_dev, onehot_dev, bins_dev);
    SSFE_LAUNCHED(ctx);
    return SSFE_OK;
}
