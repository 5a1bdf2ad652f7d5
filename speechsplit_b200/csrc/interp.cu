// interp.cu - "next" row of SURVEY.md 8(f), rank 3: InterpLnr.forward in training mode
// (reference model.py:380-436), the random-resampling augmentation between the collator and
// quantize_f0_torch (solver.py:160-162).
//
// The reference builds it from ~20 eager ops, a boolean-mask gather, a host sync
// (counts.tolist(), model.py:432) and a Python loop over the batch (pad_sequences, :366-377).
// Here it is one kernel and no sync: a CTA per batch item decides, for its max_num_seg x
// 2 max_len_seg candidate positions, which survive the two masks (:405, :414), numbers the survivors
// in the reference's order with ballots, and then all threads write the output rows
// (1 - lambda) x[i0] + lambda x[i0 + 1], zero beyond the survivors, truncated at max_len_pad.
// The random draws (scales, segment lengths) are inputs: the Python mirror makes them with the same
// torch calls in the same order as the reference, so seeded runs consume the generator identically.
// Arithmetic follows torch's float32 evaluation order exactly (division, floor, subtraction, two
// products and a sum, no contraction): results are bit-identical.
#include "common.cuh"

namespace ssfe {

constexpr int kInterpThreads = 256;
constexpr int kInterpMaxItems = 1024;      // max_num_seg * 2 * max_len_seg
constexpr int kInterpMaxPad = 512;         // max_len_pad

__global__ void __launch_bounds__(kInterpThreads) interp_lnr_kernel(const float *__restrict__ x, int T, int C,
                                                                    const int64_t *__restrict__ len_seq,
                                                                    const float *__restrict__ scales,
                                                                    const int64_t *__restrict__ len_seg, int S, int seg2,
                                                                    int max_len_pad, float *__restrict__ out)
{
    __shared__ int s_i0[kInterpMaxPad];
    __shared__ float s_lam[kInterpMaxPad];
    __shared__ int s_warp[kInterpThreads / 32];
    __shared__ int s_base;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int n_items = S * seg2;
    const float lim_seq = static_cast<float>(len_seq[b] - 1);
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int it0 = 0; it0 < n_items; it0 += kInterpThreads) {
        const int it = it0 + tid;
        bool valid = false;
        int i0 = 0;
        float lam = 0.0f;
        if (it < n_items) {
            const int s = it / seg2, j = it - s * seg2;
            const float scale = scales[b * S + s];
            const float idx_scaled = __fdiv_rn(static_cast<float>(j), scale);            // :397
            const float fl = floorf(idx_scaled);                                         // :398
            lam = __fsub_rn(idx_scaled, fl);                                             // :399
            long long off = 0;                                                           // :408-410
            for (int q = 0; q < s; ++q) off += len_seg[b * S + q];
            const float org = __fadd_rn(fl, static_cast<float>(off));                    // :412
            valid = (fl < static_cast<float>(len_seg[b * S + s] - 1)) && (org < lim_seq);   // :405, :415-417
            i0 = static_cast<int>(org);
        }
        // number the survivors in (segment, position) order
        const unsigned m = __ballot_sync(0xffffffffu, valid);
        if (lane == 0) s_warp[w] = __popc(m);
        __syncthreads();
        int before = s_base;
        for (int q = 0; q < w; ++q) before += s_warp[q];
        const int pos = before + __popc(m & ((1u << lane) - 1u));
        if (valid && pos < max_len_pad) {
            s_i0[pos] = i0;
            s_lam[pos] = lam;
        }
        __syncthreads();
        if (tid == 0) {
            int tot = 0;
            for (int q = 0; q < kInterpThreads / 32; ++q) tot += s_warp[q];
            s_base += tot;
        }
        __syncthreads();
    }
    const int count = min(s_base, max_len_pad);
    const float *xb = x + static_cast<size_t>(b) * T * C;
    float *ob = out + static_cast<size_t>(b) * max_len_pad * C;
    for (int e = tid; e < max_len_pad * C; e += kInterpThreads) {
        const int pos = e / C, c = e - pos * C;
        float y = 0.0f;
        if (pos < count) {
            const float lam = s_lam[pos];
            const float *r0 = xb + static_cast<size_t>(s_i0[pos]) * C;
            // (1 - lambda) * y_fl + lambda * y_cl, each step rounded (:427)
            y = __fadd_rn(__fmul_rn(__fsub_rn(1.0f, lam), r0[c]), __fmul_rn(lam, r0[C + c]));
        }
        ob[e] = y;
    }
}

}  // namespace ssfe

extern "C" int ssfe_interp_lnr(ssfe_ctx *ctx, const float *x_dev, int batch, int T, int C, const int64_t *len_seq_dev,
                               const float *scales_dev, const int64_t *len_seg_dev, int max_num_seg, int max_len_seg,
                               int max_len_pad, float *out_dev)
{
    using namespace ssfe;
    if (!ctx) return SSFE_ERR_INVALID;
    if (batch == 0) return SSFE_OK;
    if (!x_dev || !len_seq_dev || !scales_dev || !len_seg_dev || !out_dev || batch < 0 || T < 2 || C < 1 ||
        max_num_seg < 1 || max_len_seg < 1)
        return set_error(ctx, SSFE_ERR_INVALID, "ssfe_interp_lnr: bad argument");
    if (max_num_seg * 2 * max_len_seg > kInterpMaxItems || max_len_pad < 1 || max_len_pad > kInterpMaxPad)
        return set_error(ctx, SSFE_ERR_RANGE, "ssfe_interp_lnr: at most %d candidate positions and %d output frames",
                         kInterpMaxItems, kInterpMaxPad);
    SSFE_CUDA(ctx, cudaSetDevice(ctx->device));
    interp_lnr_kernel<<<batch, kInterpThreads, 0, ctx->stream>>>(x_dev, T, C, len_seq_dev, scales_dev, len_seg_dev,
                                                                 max_num_seg, 2 * max_len_seg, max_len_pad, out_dev);
    SSFE_LAUNCHED(ctx);
    return SSFE_OK;
}
