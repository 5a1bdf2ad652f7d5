// mt19937.cu - the per-speaker dither stream (stage a2, random part).
//
// Replaces numpy.random.RandomState(int(subdir[1:])).rand(L) at reference make_spect_f0.py:47,55.
// RandomState(seed) is MT19937 seeded with init_genrand (Knuth LCG 1812433253); rand() returns
// ((a >> 5) * 2^26 + (b >> 6)) / 2^53 from two successive tempered 32-bit outputs.  A speaker's
// stream continues across its files in sorted order, so utterance k starts `skip` doubles in.
//
// One CTA per distinct seed walks the stream block by block (624 words).  Word k of the new block
// depends on new word k-227, so thread t produces words t, t+227, t+454 from the old block and its
// own results: one barrier per block.  The CTA is warp specialised (see mt19937_kernel): a twist
// group advances the state, an emit group trails through a ring of state blocks and writes the raw
// word pairs of every requested double; tempering and the conversion to double happen in the
// consumer (mt_convert.cuh).  Streams are independent, so seeds run on different SMs.
#include "common.cuh"
#include "mt_convert.cuh"
#include <algorithm>
#include <numeric>

namespace ssfe {

constexpr int kMtN = 624, kMtM = 397;

struct RandJob {
    uint32_t seed;
    int first_req, n_req;
    int pad;
};
struct RandReq {
    uint64_t skip;       // first double of the stream this request wants
    int64_t count;       // doubles
    int64_t out_off;     // into the output array
};

__device__ __forceinline__ uint32_t mt_mix(uint32_t a, uint32_t b)
{
    const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
    return (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}
// Named-barrier helpers (bar.sync / bar.arrive with an explicit thread count): barrier 0 is
// __syncthreads(), ids 1.. are ours.
__device__ __forceinline__ void bar_sync(int id, int count)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_arrive(int id, int count)
{
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

constexpr int kMtRing = 4;                       // state blocks in flight between the two warp groups
constexpr int kMtStreams = 1;                    // speaker streams advanced together by one CTA (see below)
constexpr int kMtProd = 256, kMtCons = 256;
constexpr int kBarFull = 1, kBarEmpty = 1 + kMtRing, kBarProd = 1 + 2 * kMtRing;

// Warp-specialised: warps 0-7 ("twist") only advance the state, one named barrier per block among
// themselves - that chain (shared-memory load, ~20 integer ops, store, barrier) is the serial
// bottleneck of a speaker's stream, so nothing else is on it.  Warps 8-15 ("emit") trail behind
// through a ring of kMtRing state blocks and copy the raw word pairs of every requested double to
// global memory.  full[slot] / empty[slot] named barriers hand blocks back and forth.
// Measured on B200: ~300 ns per 624-word block and stream whichever way the hand-off is built
// (one group + __syncthreads 340 ns, this version 306 ns, mbarrier hand-off 360-380 ns): the floor is
// the shared-memory round trip plus one CTA-level barrier per block, so a 400-utterance speaker
// (61.5 k blocks) costs ~18 ms.  Advancing several streams per CTA (kMtStreams > 1) was measured
// too: the time per round grows in proportion (1110 ns for 4 streams), i.e. the CTA is bound by its
// own instruction issue (~900 warp instructions per block), not by barrier latency - hence few warps
// with several words per thread here.
__global__ void __launch_bounds__(kMtProd + kMtCons) mt19937_kernel(const RandJob *__restrict__ jobs, int n_jobs,
                                                                    const RandReq *__restrict__ reqs,
                                                                    uint2 *__restrict__ out)
{
    extern __shared__ __align__(16) uint32_t s_mt_raw[];          // [kMtStreams][kMtRing][kMtN]
    auto mt = [&](int s, int slot) -> uint32_t * { return s_mt_raw + (s * kMtRing + slot) * kMtN; };
    const int tid = threadIdx.x;
    const int j0 = blockIdx.x * kMtStreams;
    const int ns = min(kMtStreams, n_jobs - j0);

    long long nb[kMtStreams];                                      // data blocks 1..nb[s] of stream s
    long long n_blocks = 0;
#pragma unroll
    for (int s = 0; s < kMtStreams; ++s) {
        nb[s] = 0;
        if (s < ns) {
            const RandJob job = jobs[j0 + s];
            const RandReq last = reqs[job.first_req + job.n_req - 1];
            nb[s] = static_cast<long long>((last.skip + static_cast<uint64_t>(last.count) + 311) / 312);
            n_blocks = max(n_blocks, nb[s]);
        }
    }
    if (tid < ns) {   // init_genrand(seed): a 624-step serial LCG, once per stream
        uint32_t v = jobs[j0 + tid].seed;
        uint32_t *m0 = mt(tid, 0);
        m0[0] = v;
        for (int i = 1; i < kMtN; ++i) {
            v = 1812433253u * (v ^ (v >> 30)) + static_cast<uint32_t>(i);
            m0[i] = v;
        }
    }
    __syncthreads();

    if (tid < kMtProd) {
        // ---- twist group --------------------------------------------------------------------------
        for (long long B = 1; B <= n_blocks; ++B) {
            const int slot = static_cast<int>(B & (kMtRing - 1)), prev = static_cast<int>((B - 1) & (kMtRing - 1));
            // slot held data block B - kMtRing: wait until the emit group is done with it
            if (B >= kMtRing + 1) bar_sync(kBarEmpty + slot, kMtProd + kMtCons);
            // Word k of the new block needs the NEW word k-227 and 3 * 227 > 624: thread t produces
            // words t, t+227, t+454 from the old block and its own results - no barrier in between.
#pragma unroll
            for (int s = 0; s < kMtStreams; ++s) {
                if (B > nb[s]) continue;
                const uint32_t *o = mt(s, prev);
                uint32_t *nw = mt(s, slot);
#pragma unroll
                for (int t = tid; t < 227; t += kMtProd) {
                    const uint32_t n0 = o[t + kMtM] ^ mt_mix(o[t], o[t + 1]);
                    const uint32_t n1 = n0 ^ mt_mix(o[t + 227], o[t + 228]);
                    nw[t] = n0;
                    nw[t + 227] = n1;
                    if (t < 169) {
                        nw[t + 454] = n1 ^ mt_mix(o[t + 454], o[t + 455]);
                    } else if (t == 169) {   // word 623 wraps around to the new word 0 (recomputed here)
                        const uint32_t new0 = o[kMtM] ^ mt_mix(o[0], o[1]);
                        nw[623] = n1 ^ mt_mix(o[623], new0);
                    }
                }
            }
            // block B complete for the twist group (the barrier also drains the shared-memory stores),
            // then published to the emit group
            bar_sync(kBarProd, kMtProd);
            bar_arrive(kBarFull + slot, kMtProd + kMtCons);
        }
    } else {
        // ---- emit group ---------------------------------------------------------------------------
        const int ct = tid - kMtProd;
        int r0[kMtStreams], nreq[kMtStreams];
        uint64_t r_beg[kMtStreams], r_end[kMtStreams];
        int64_t r_out[kMtStreams];
        const RandReq *rq[kMtStreams];
#pragma unroll
        for (int s = 0; s < kMtStreams; ++s) {
            r0[s] = 0; nreq[s] = 0; r_beg[s] = 0; r_end[s] = 0; r_out[s] = 0; rq[s] = reqs;
            if (s < ns) {
                const RandJob job = jobs[j0 + s];
                rq[s] = reqs + job.first_req;
                nreq[s] = job.n_req;
                r_beg[s] = rq[s][0].skip;
                r_end[s] = r_beg[s] + static_cast<uint64_t>(rq[s][0].count);
                r_out[s] = rq[s][0].out_off;
            }
        }
        for (long long B = 1; B <= n_blocks; ++B) {
            const int slot = static_cast<int>(B & (kMtRing - 1));
            bar_sync(kBarFull + slot, kMtProd + kMtCons);
            const uint64_t d0 = static_cast<uint64_t>(B - 1) * 312;   // first double of this block
#pragma unroll
            for (int s = 0; s < kMtStreams; ++s) {
                if (B > nb[s]) continue;
                const uint32_t *nw = mt(s, slot);
                while (r0[s] < nreq[s] && r_end[s] <= d0) {
                    ++r0[s];
                    if (r0[s] < nreq[s]) {
                        r_beg[s] = rq[s][r0[s]].skip;
                        r_end[s] = r_beg[s] + static_cast<uint64_t>(rq[s][r0[s]].count);
                        r_out[s] = rq[s][r0[s]].out_off;
                    }
                }
                if (r0[s] >= nreq[s]) continue;
                for (int t = ct; t < 312; t += kMtCons) {
                    if (r_beg[s] <= d0 && d0 + 312 <= r_end[s]) {
                        // common case: the whole block belongs to the current request
                        out[r_out[s] + static_cast<int64_t>(d0 - r_beg[s]) + t] = *reinterpret_cast<const uint2 *>(nw + 2 * t);
                    } else if (r_beg[s] < d0 + 312) {
                        const uint64_t d = d0 + t;
                        uint64_t b = r_beg[s], e = r_end[s];
                        int64_t oo = r_out[s];
                        int r = r0[s];
                        while (r < nreq[s] && e <= d) {      // the block spans a request boundary
                            ++r;
                            if (r < nreq[s]) {
                                b = rq[s][r].skip;
                                e = b + static_cast<uint64_t>(rq[s][r].count);
                                oo = rq[s][r].out_off;
                            }
                        }
                        if (r < nreq[s] && d >= b)
                            out[oo + static_cast<int64_t>(d - b)] = *reinterpret_cast<const uint2 *>(nw + 2 * t);
                    }
                }
            }
            // hand the slot back if the twist group will reuse it (data block B + kMtRing)
            if (B + kMtRing <= n_blocks) bar_arrive(kBarEmpty + slot, kMtProd + kMtCons);
        }
    }
}

constexpr size_t kMtSmem = static_cast<size_t>(kMtStreams) * kMtRing * kMtN * sizeof(uint32_t);

__global__ void mt_convert_kernel(double *__restrict__ u, int64_t count)
{
    const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i >= count) return;
    const uint2 raw = reinterpret_cast<const uint2 *>(u)[i];
    u[i] = mt_raw_to_double(raw);
}

int rand_run(ssfe_ctx *ctx, const uint32_t *seeds, const uint64_t *skip, const int64_t *out_off, int n,
             double *u_dev, cudaStream_t launch_on)
{
    if (n == 0) return SSFE_OK;
    // group requests by seed, each group ordered by stream position
    std::vector<int> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        if (seeds[a] != seeds[b]) return seeds[a] < seeds[b];
        return skip[a] < skip[b];
    });
    std::vector<RandJob> jobs;
    std::vector<RandReq> reqs;
    reqs.reserve(n);
    for (int k = 0; k < n; ++k) {
        const int i = order[k];
        const int64_t cnt = out_off[i + 1] - out_off[i];
        if (cnt <= 0) continue;
        if (jobs.empty() || jobs.back().seed != seeds[i]) {
            RandJob j;
            j.seed = seeds[i];
            j.first_req = static_cast<int>(reqs.size());
            j.n_req = 0;
            j.pad = 0;
            jobs.push_back(j);
        } else {
            const RandReq &prev = reqs.back();
            if (skip[i] < prev.skip + static_cast<uint64_t>(prev.count))
                return set_error(ctx, SSFE_ERR_INVALID, "ssfe_rand: overlapping stream ranges for seed %u", seeds[i]);
        }
        RandReq r;
        r.skip = skip[i];
        r.count = cnt;
        r.out_off = out_off[i];
        reqs.push_back(r);
        jobs.back().n_req++;
    }
    if (jobs.empty()) return SSFE_OK;
    // streams that share a CTA should have similar lengths: order the jobs by their last position
    std::stable_sort(jobs.begin(), jobs.end(), [&](const RandJob &a, const RandJob &b) {
        const RandReq &ra = reqs[a.first_req + a.n_req - 1], &rb = reqs[b.first_req + b.n_req - 1];
        return ra.skip + static_cast<uint64_t>(ra.count) > rb.skip + static_cast<uint64_t>(rb.count);
    });
    if (!ctx->mt_attr_set) {
        SSFE_CUDA(ctx, cudaFuncSetAttribute(mt19937_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            static_cast<int>(kMtSmem)));
        ctx->mt_attr_set = true;
    }
    if (launch_on == nullptr || launch_on == ctx->stream) {
        // public ssfe_rand path: everything on the caller's stream, doubles out
        RandJob *d_jobs = upload(ctx, jobs.data(), jobs.size());
        RandReq *d_reqs = upload(ctx, reqs.data(), reqs.size());
        if (!d_jobs || !d_reqs) return SSFE_ERR_NOMEM;
        mt19937_kernel<<<static_cast<unsigned>((jobs.size() + kMtStreams - 1) / kMtStreams), kMtProd + kMtCons, kMtSmem, ctx->stream>>>(
            d_jobs, static_cast<int>(jobs.size()), d_reqs, reinterpret_cast<uint2 *>(u_dev));
        SSFE_LAUNCHED(ctx);
        const int64_t total = out_off[n];
        mt_convert_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, ctx->stream>>>(u_dev, total);
        SSFE_LAUNCHED(ctx);
        return SSFE_OK;
    }
    // pipeline path: the generator runs on the side stream and leaves RAW word pairs.  It only has to
    // wait for the previous consumer of the dither buffer (ev_dith_free), not for the rest of the
    // previous ssfe_extract call, so in a stream of calls it overlaps the previous call's STFT / RAPT.
    // Its metadata therefore travels on the side stream too, through its own double-buffered staging.
    cudaStream_t st = launch_on;
    const size_t jb = jobs.size() * sizeof(RandJob), rb = reqs.size() * sizeof(RandReq);
    const size_t need = (jb + 255) / 256 * 256 + rb;
    const int slot = ctx->aux_idx ^= 1;
    if (ctx->aux_free[slot]) SSFE_CUDA(ctx, cudaEventSynchronize(ctx->aux_free[slot]));   // two calls back
    else SSFE_CUDA(ctx, cudaEventCreateWithFlags(&ctx->aux_free[slot], cudaEventDisableTiming));
    if (need > ctx->aux_cap[slot]) {
        if (ctx->aux_host[slot]) cudaFreeHost(ctx->aux_host[slot]);
        if (ctx->aux_dev[slot]) cudaFree(ctx->aux_dev[slot]);
        ctx->aux_cap[slot] = need + need / 4 + 4096;
        SSFE_CUDA(ctx, cudaMallocHost(reinterpret_cast<void **>(&ctx->aux_host[slot]), ctx->aux_cap[slot]));
        SSFE_CUDA(ctx, cudaMalloc(reinterpret_cast<void **>(&ctx->aux_dev[slot]), ctx->aux_cap[slot]));
    }
    char *h = ctx->aux_host[slot], *d = ctx->aux_dev[slot];
    memcpy(h, jobs.data(), jb);
    memcpy(h + (jb + 255) / 256 * 256, reqs.data(), rb);
    SSFE_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_dith_free, 0));
    SSFE_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_mt_go, 0));     // see rapt_run: start beside the Viterbi kernel
    SSFE_CUDA(ctx, cudaMemcpyAsync(d, h, need, cudaMemcpyHostToDevice, st));
    mark_aux(ctx, 0, st);
    mt19937_kernel<<<static_cast<unsigned>((jobs.size() + kMtStreams - 1) / kMtStreams), kMtProd + kMtCons, kMtSmem, st>>>(
        reinterpret_cast<const RandJob *>(d), static_cast<int>(jobs.size()),
        reinterpret_cast<const RandReq *>(d + (jb + 255) / 256 * 256), reinterpret_cast<uint2 *>(u_dev));
    SSFE_LAUNCHED(ctx);
    mark_aux(ctx, 1, st);
    SSFE_CUDA(ctx, cudaEventRecord(ctx->ev_join, st));
    SSFE_CUDA(ctx, cudaEventRecord(ctx->aux_free[slot], st));
    return SSFE_OK;
}

}  // namespace ssfe

extern "C" int ssfe_rand(ssfe_ctx *ctx, const uint32_t *seeds, const uint64_t *skip, const int64_t *out_offsets,
                         int n_utts, double *u_dev)
{
    if (!ctx) return SSFE_ERR_INVALID;
    if (n_utts < 0 || (n_utts > 0 && (!seeds || !skip || !out_offsets || !u_dev)))
        return ssfe::set_error(ctx, SSFE_ERR_INVALID, "ssfe_rand: null argument");
    return ssfe::rand_run(ctx, seeds, skip, out_offsets, n_utts, u_dev);
}
