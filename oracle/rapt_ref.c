/*
 * rapt_ref.c - CPU restatement of the RAPT pitch tracker as the reference invokes it.
 *
 * TEST INFRASTRUCTURE ONLY (oracle).  Never linked into, imported by or executed from the
 * product path; only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs use it.
 *
 * PARITY UNPINNED.  The reference calls a third-party routine that is neither vendored under
 * /root/reference nor installed in the build image:
 *     /root/reference/make_spect_f0.py:64
 *         f0_rapt = sptk.rapt(wav.astype(np.float32)*32768, fs, 256, min=lo, max=hi, otype=2)
 *     -> pysptk (unpinned, README.md:29) -> SPTK 3.x  bin/pitch/snack/{jkGetF0.c,sigproc.c}
 *        (the Snack/ESPS "get_f0" of D. Talkin; "A Robust Algorithm for Pitch Tracking", 1995).
 * This file restates that published algorithm with the ESPS default parameters that SPTK's
 * rapt() installs, including its block-streaming structure (0.2 s reads, commit-on-convergence
 * back-tracking) and its single/double precision mix, because voicing decisions are
 * discontinuous in the arithmetic.  It was written without access to the SPTK sources; where a
 * detail is uncertain it is marked [U].  The reference's own pins for this stage are only:
 * the output length  T == len(S) (make_spect_f0.py:69), the unvoiced sentinel -1e10 (:65) and
 * "f0 is log f0" (utils.py:36).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off: x86-64 SSE2 float semantics, no FMA).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#define RAPT_MAX_ORDER 100        /* BIGSORD */
#define RAPT_READ_SIZE 0.2        /* seconds per streaming read */
#define RAPT_DP_CIRCULAR 1.5
#define RAPT_DP_HIST 0.5          /* history needed before a commit is attempted */
#define RAPT_DP_LIMIT 1.0         /* latest commit */
#define RAPT_STAT_WSIZE 0.030
#define RAPT_STAT_AINT 0.020
#define RAPT_CMAX 20              /* n_cands */

/* per-frame debug record exported to the tests (all arrays caller-allocated, may be NULL) */
typedef struct {
    int   max_frames;
    int  *ncands;      /* [max_frames] including the unvoiced candidate */
    int  *locs;        /* [max_frames*RAPT_CMAX]  lag, -1 = unvoiced */
    float *pvals;      /* [max_frames*RAPT_CMAX]  fine NCCF peak (unvoiced: maxval) */
    float *mpvals;     /* [max_frames*RAPT_CMAX]  local cost */
    float *stat;       /* [max_frames] */
    float *rms_ratio;  /* [max_frames] */
    float *f0cand;     /* [max_frames*RAPT_CMAX] refined F0 (Hz) each candidate would emit */
    int  *prept;       /* [max_frames*RAPT_CMAX] back pointers */
    float *dpvals;     /* [max_frames*RAPT_CMAX] accumulated costs */
    float *ds;         /* [ds_cap] global 2 kHz stream as seen by the coarse stage */
    int   ds_cap;
    int   n_frames;    /* out: frames analysed */
    int   n_forced;    /* out: commits forced without path convergence */
} rapt_debug;

typedef struct {
    int   ncands;
    short locs[RAPT_CMAX];
    float pvals[RAPT_CMAX], mpvals[RAPT_CMAX], dpvals[RAPT_CMAX];
    short prept[RAPT_CMAX];
    float f0cand[RAPT_CMAX];
} frame_rec;

typedef struct {
    /* ESPS parameters (single precision, as in F0_params) */
    float cand_thresh, lag_weight, freq_weight, trans_cost, trans_amp, trans_spec;
    float voice_bias, double_cost, min_f0, max_f0, frame_step, wind_dur;
    int   n_cands;
    /* derived (init) */
    double freq;
    float tcost, tfact_a, tfact_s, frame_int, vbias, fdouble, ln2, freqwt, lagwt;
    int   step, size, nlags, start, stop, ncomp, pad;
    int   size_frame_hist, size_frame_out;
    int   decimate;
    /* decimator state */
    int   ncoeff, ncoefft;
    float half[2048];
    float co[4096], dsmem[4096], dsstate[1000];
    float *dsout; int dsout_cap;
    /* stationarity state */
    float *stmem; int stmemsize;
    float *hwin479, *hwin480;
    /* DP store */
    frame_rec *fr; int fr_cap, n_fr;
    int   head, tail, num_active, first_time;
    float *correl;      /* nlags + slack */
    float *dbdata;
} rapt_ctx;

static int iround(double x) { return (int)(x + 0.5); }

/* ---- A1: 2 kHz decimator ------------------------------------------------------------- */
static void design_lowpass(float fc, int *nf, float *coef)
{
    int i, n;
    double twopi, fn, c;
    if ((*nf % 2) != 1) *nf = *nf + 1;
    n = (*nf + 1) / 2;
    twopi = M_PI * 2.0;
    coef[0] = 2.0 * fc;
    c = M_PI;
    fn = twopi * fc;
    for (i = 1; i < n; i++) coef[i] = sin(i * fn) / (c * i);
    fn = twopi / (double)(*nf);
    for (i = 0; i < n; i++) coef[n - i - 1] *= (.5 - (.5 * cos(fn * ((double)i + 0.5))));
}

/* One streaming read through the symmetric FIR, keeping every `skip`-th output.
 * init&1: start of signal (zero history); init&2: end of signal (flush with zeros).
 * Output l of this read is centred on input sample skip*l of this read. */
static void decimate_read(rapt_ctx *c, const float *in, int in_samps, float *out, int *out_samps,
                          int state_idx, int init)
{
    const int nc = c->ncoefft, k = 2 * nc - 1, skip = c->decimate;
    float *mem = c->dsmem, *co = c->co;
    const float *buf = in;
    int i, j, l;
    float sum;

    for (i = 0; i < nc; i++) mem[nc - 1 + i] = *buf++;
    if (init & 1) {
        for (i = 0; i < nc - 1; i++) co[i] = co[k - 1 - i] = c->half[nc - 1 - i];
        co[nc - 1] = c->half[0];
        for (i = 0; i < nc - 1; i++) mem[i] = 0;
    } else {
        for (i = 0; i < nc - 1; i++) mem[i] = c->dsstate[i];
    }
    for (l = 0; l < *out_samps; l++) {
        sum = 0.0;
        for (j = 0; j < k - skip; j++) { sum += co[j] * mem[j]; mem[j] = mem[j + skip]; }
        for (; j < k; j++) { sum += co[j] * mem[j]; mem[j] = *buf++; }
        out[l] = (sum < 0.0) ? sum - 0.5 : sum + 0.5;
    }
    if (init & 2) {
        int resid = in_samps - *out_samps * skip;
        for (l = resid / skip; l-- > 0;) {
            sum = 0.0;
            for (j = 0; j < k - skip; j++) { sum += co[j] * mem[j]; mem[j] = mem[j + skip]; }
            for (; j < k; j++) { sum += co[j] * mem[j]; mem[j] = 0.0; }
            out[*out_samps] = (sum < 0.0) ? sum - 0.5 : sum + 0.5;
            (*out_samps)++;
        }
    } else {
        /* history for the next read: the nc-1 samples before state_idx.  (For a first read
         * that is also the only one this may index past in_samps; the values are never used.) */
        for (l = 0; l < nc - 1; l++) {
            int p = state_idx - nc + 1 + l;
            c->dsstate[l] = (p >= 0 && p < in_samps) ? in[p] : 0.0f;
        }
    }
}

static float *decimate_stream(rapt_ctx *c, const float *in, int samsin, int state_idx,
                              int *samsout, int first_time, int last_time)
{
    int init;
    if (first_time) {
        /* buffer sized (and zeroed) once, BEFORE ncoeff is updated from its initial 127 */
        int nbuff = (samsin / c->decimate) + (2 * 127);
        float beta;
        c->ncoeff = ((int)(c->freq * .005)) | 1;
        beta = .5 / c->decimate;
        free(c->dsout);
        c->dsout = (float *)calloc((size_t)nbuff + 64, sizeof(float));
        c->dsout_cap = nbuff;
        design_lowpass(beta, &c->ncoeff, c->half);
        c->ncoefft = (c->ncoeff / 2) + 1;
    }
    if (first_time) init = 1;
    else if (last_time) init = 2;
    else init = 0;
    decimate_read(c, in, samsin, c->dsout, samsout, state_idx, init);
    return c->dsout;
}

/* ---- windows, LPC, Itakura distance (A6) --------------------------------------------- */
static void make_hanning(float *w, int n)
{
    int i;
    double arg = 3.1415927 * 2.0 / n, half = 0.5;
    for (i = 0; i < n; i++) w[i] = (half - half * cos((half + (double)i) * arg));
}

/* dout[i] = w[i]*(din[i+1]-preemp*din[i])  (needs n+1 input samples when preemp != 0) */
static void window_preemph(const float *din, float *dout, const float *w, int n, float preemp)
{
    int i;
    if (preemp != 0.0) {
        for (i = 0; i < n; i++) dout[i] = w[i] * ((float)din[i + 1] - (preemp * din[i]));
    } else {
        for (i = 0; i < n; i++) dout[i] = w[i] * din[i];
    }
}

static void autoc(int wsize, const float *s, int p, float *r, float *e)
{
    int i, j;
    float sum, sum0;
    for (i = 0, sum0 = 0.0; i < wsize; i++) sum0 += s[i] * s[i];
    r[0] = 1.;
    if (sum0 == 0.0) {
        *e = 1.;
        for (i = 1; i <= p; i++) r[i] = 0.;
        return;
    }
    *e = sqrt((double)(sum0 / wsize));
    sum0 = 1.0 / sum0;
    for (i = 1; i <= p; i++) {
        for (sum = 0.0, j = 0; j < wsize - i; j++) sum += s[j] * s[j + i];
        r[i] = sum * sum0;
    }
}

static void durbin(const float *r, float *k, float *a, int p, float *ex)
{
    float bb[RAPT_MAX_ORDER];
    int i, j;
    float e, s, *b = bb;
    e = *r;
    *k = -r[1] / e;
    *a = *k;
    e *= (1. - (*k) * (*k));
    for (i = 1; i < p; i++) {
        s = 0;
        for (j = 0; j < i; j++) s -= a[j] * r[i - j];
        k[i] = (s - r[i + 1]) / e;
        a[i] = k[i];
        for (j = 0; j <= i; j++) b[j] = a[j];
        for (j = 0; j < i; j++) a[j] += k[i] * b[i - j - 1];
        e *= (1. - (k[i] * k[i]));
    }
    *ex = e;
}

/* LPC of one window: lpca[0..p], normalised autocorrelation ar[0..p], residual energy */
static void lpc_window(rapt_ctx *c, int order, float stabl, int wsize, const float *data,
                       float *lpca, float *ar, float *normerr, float preemp)
{
    float dwind[1024], rho[RAPT_MAX_ORDER + 1], k[RAPT_MAX_ORDER], en, er;
    float *r = ar;
    window_preemph(data, dwind, c->hwin479, wsize, preemp);
    autoc(wsize, dwind, order, r, &en);
    if (stabl > 1.0) {
        int i;
        float ffact;
        ffact = 1.0 / (1.0 + exp((-stabl / 20.0) * log(10.0)));
        for (i = 1; i <= order; i++) rho[i] = ffact * r[i];
        *rho = *r;
        r = rho;
        for (i = 0; i <= order; i++) ar[i] = r[i];
    }
    durbin(r, k, &lpca[1], order, &er);
    *lpca = 1.0;
    *normerr = er;
}

static void a_to_aca(const float *a, float *b, float *c, int p)
{
    float s;
    int i, j;
    for (s = 1., i = 0; i < p; i++) s += a[i] * a[i];
    *c = s;
    for (i = 1; i <= p; i++) {
        s = a[i - 1];
        for (j = 0; j < p - i; j++) s += (a[j] * a[j + i]);
        b[i - 1] = 2. * s;
    }
}

static float itakura(int p, const float *b, const float *c, const float *r, const float *gain)
{
    float s;
    int i;
    for (s = *c, i = 0; i < p; i++) s += r[i] * b[i];
    return (s / *gain);
}

static float wind_energy(rapt_ctx *c, const float *data, int size)
{
    float sum, f;
    int i;
    for (i = 0, sum = 0.0; i < size; i++) {
        f = c->hwin480[i] * (float)data[i];
        sum += f * f;
    }
    return (float)sqrt((double)(sum / size));
}

static float similarity(rapt_ctx *c, int order, int size, const float *pdata, const float *cdata,
                        float *rmsa, float *rms_ratio, float pre, float stab, int init)
{
    float rho3[RAPT_MAX_ORDER + 1], err3, rms3, b0, t, a2[RAPT_MAX_ORDER + 1];
    float rho1[RAPT_MAX_ORDER + 1], a1[RAPT_MAX_ORDER + 1], b[RAPT_MAX_ORDER + 1], err1, rms1;

    lpc_window(c, order, stab, size - 1, cdata, a2, rho3, &err3, pre);
    rms3 = wind_energy(c, cdata, size);
    if (!init) {
        lpc_window(c, order, stab, size - 1, pdata, a1, rho1, &err1, pre);
        a_to_aca(a2 + 1, b, &b0, order);
        t = itakura(order, b, &b0, rho1 + 1, &err1) - .8;
        rms1 = wind_energy(c, pdata, size);
        *rms_ratio = (0.001 + rms3) / rms1;
    } else {
        *rms_ratio = 1.0;
        t = 10.0;
    }
    *rmsa = rms3;
    return ((float)(0.2 / t));
}

/* per-read stationarity / rms-ratio track; stream position is carried in c->stmem */
static void stationarity_read(rapt_ctx *c, const float *fdata, int buff_size, int nframes,
                              int frame_step, int first_time, float *stat, float *rms, float *rms_ratio)
{
    float preemp = 0.4f, stab = 30.0f;
    const float *p, *q, *r, *datend;
    int ind, i, j, m, size, order, agap;
    int memsize;
    float *mem;

    agap = (int)(RAPT_STAT_AINT * c->freq);
    size = (int)(RAPT_STAT_WSIZE * c->freq);
    ind = (agap - size) / 2;
    memsize = (int)(RAPT_STAT_WSIZE * c->freq) + (int)(RAPT_STAT_AINT * c->freq);
    if (first_time || !c->stmem) {
        free(c->stmem);
        c->stmem = (float *)calloc((size_t)memsize, sizeof(float));
        c->stmemsize = memsize;
    }
    mem = c->stmem;
    if (nframes == 0) return;

    q = fdata + ind;
    datend = fdata + buff_size;
    if ((order = 2.0 + (c->freq / 1000.0)) > RAPT_MAX_ORDER) order = RAPT_MAX_ORDER;

    for (j = memsize / 2, i = 0; j < memsize; j++, i++) mem[j] = fdata[i];

    for (j = 0, p = q - agap; j < nframes; j++, p += frame_step, q += frame_step) {
        if ((p >= fdata) && (q >= fdata) && (q + size <= datend)) {
            stat[j] = similarity(c, order, size, p, q, &rms[j], &rms_ratio[j], preemp, stab, 0);
        } else {
            if (first_time) {
                if ((p < fdata) && (q >= fdata) && (q + size <= datend)) {
                    stat[j] = similarity(c, order, size, NULL, q, &rms[j], &rms_ratio[j], preemp, stab, 1);
                } else {
                    rms[j] = 0.0;
                    stat[j] = 0.01f * 0.2f;
                    rms_ratio[j] = 1.0;
                }
            } else {
                if ((p < fdata) && (q + size <= datend)) {
                    stat[j] = similarity(c, order, size, mem, mem + (memsize / 2) + ind,
                                         &rms[j], &rms_ratio[j], preemp, stab, 0);
                    if (p + frame_step < fdata) {
                        for (m = 0; m < (memsize - frame_step); m++) mem[m] = mem[m + frame_step];
                        r = q + size;
                        for (m = 0; m < frame_step; m++) mem[memsize - frame_step + m] = *r++;
                    }
                }
            }
        }
    }
    for (j = (memsize / 2) - 1, p = fdata + (nframes * frame_step) - 1; j >= 0 && p >= fdata; j--)
        mem[j] = *p--;
}

/* ---- A2/A4: normalised cross-correlation --------------------------------------------- */
static void ncc_all_lags(rapt_ctx *c, const float *data, int size, int start, int nlags,
                         float *engref, int *maxloc, float *maxval, float *correl)
{
    float *dbdata = c->dbdata;
    float sum, st, t, engr, amax;
    double engc;
    int i, j, iloc;

    for (engr = 0.0, j = 0; j < size; j++) engr += data[j];
    engr /= size;
    for (j = 0; j < size + nlags + start; j++) dbdata[j] = data[j] - engr;

    for (j = 0, sum = 0.0; j < size; j++) { st = dbdata[j]; sum += st * st; }
    *engref = engr = sum;
    if (engr > 0.0) {
        for (j = 0, sum = 0.0; j < size; j++) { st = dbdata[start + j]; sum += st * st; }
        engc = sum;
        for (i = 0, amax = 0.0, iloc = -1; i < nlags; i++) {
            const float *ds = dbdata + i + start;
            for (j = 0, sum = 0.0; j < size; j++) sum += dbdata[j] * ds[j];
            correl[i] = t = (sum / sqrt((double)(engc * engr)));
            engc -= (double)(ds[0] * ds[0]);
            if ((engc += (double)(ds[size] * ds[size])) < 1.0) engc = 1.0;
            if (t > amax) { amax = t; iloc = i + start; }
        }
        *maxloc = iloc;
        *maxval = amax;
    } else {
        *maxloc = 0;
        *maxval = 0.0;
        for (i = 0; i < nlags; i++) correl[i] = 0.0;
    }
}

static void ncc_near_lags(rapt_ctx *c, const float *data, int size, int start0, int nlags0, int nlags,
                          float *engref, int *maxloc, float *maxval, float *correl,
                          const int *locs, int nlocs)
{
    float *dbdata = c->dbdata;
    float sum, st, t, engr, amax;
    double engc;
    int i, j, iloc, start;

    for (engr = 0.0, j = 0; j < size; j++) engr += data[j];
    engr /= size;
    for (j = 0; j < size + nlags0 + start0; j++) dbdata[j] = data[j] - engr;
    for (i = 0; i < nlags0; i++) correl[i] = 0.0;

    for (j = 0, sum = 0.0; j < size; j++) { st = dbdata[j]; sum += st * st; }
    *engref = engr = sum;
    amax = 0.0;
    iloc = -1;
    if (engr > 0.0) {
        for (; nlocs > 0; nlocs--, locs++) {
            float *dq;
            start = *locs - (nlags >> 1);
            if (start < start0) start = start0;
            dq = correl + start - start0;
            for (j = 0, sum = 0.0; j < size; j++) { st = dbdata[start + j]; sum += st * st; }
            engc = sum;
            for (i = 0; i < nlags; i++) {
                const float *ds = dbdata + i + start;
                for (j = 0, sum = 0.0; j < size; j++) sum += dbdata[j] * ds[j];
                if (engc < 1.0) engc = 1.0;
                *dq++ = t = (float)(sum / sqrt((double)(10000.0 + (engc * engr))));
                engc -= (double)(ds[0] * ds[0]);
                engc += (double)(ds[size] * ds[size]);
                if (t > amax) { amax = t; iloc = i + start; }
            }
        }
        *maxloc = iloc;
        *maxval = amax;
    } else {
        *maxloc = 0;
        *maxval = 0.0;
    }
}

/* ---- A3: candidate picking ------------------------------------------------------------ */
static void pick_candidates(const float *correl, float maxval, int firstlag, float *peak, int *loc,
                            int nlags, int *ncand, float cand_thresh)
{
    int i, lastl, ncan;
    float o, p, q, clip;
    const float *r = correl;
    clip = cand_thresh * maxval;
    lastl = nlags - 2;
    o = *r++;
    q = *r++;
    p = *r++;
    ncan = 0;
    for (i = 1; i < lastl; i++, o = q, q = p, p = *r++) {
        if ((q > clip) && (q >= p) && (q >= o)) {
            peak[ncan] = q;
            loc[ncan] = i + firstlag;
            ncan++;
        }
    }
    *ncand = ncan;
}

static void parabola(const float *y, float *xp, float *yp)
{
    float a, c;
    a = (float)((y[2] - y[1]) + (.5 * (y[0] - y[2])));
    if (fabs(a) > .000001) {
        *xp = c = (float)((y[0] - y[2]) / (4.0 * a));
        *yp = y[1] - (a * c * c);
    } else {
        *xp = 0.0;
        *yp = y[1];
    }
}

/* keep the n_keep largest peaks at the front, by the original's partial bubble pass */
static void prune_candidates(float *peaks, int *locs, int *ncand, int n_cands)
{
    int outer, inner, lim = n_cands - 1;
    for (outer = 0; outer < lim; outer++) {
        int idx = *ncand - 1;
        for (inner = *ncand - 1 - outer; inner-- > 0; idx--) {
            float smaxval = peaks[idx];
            if (smaxval > peaks[idx - 1]) {
                int lt = locs[idx];
                peaks[idx] = peaks[idx - 1];
                peaks[idx - 1] = smaxval;
                locs[idx] = locs[idx - 1];
                locs[idx - 1] = lt;
            }
        }
    }
    *ncand = n_cands - 1;
}

static void fast_candidates(rapt_ctx *c, const float *fdata, const float *fdsdata, int ind,
                            float *engref, int *maxloc, float *maxval, float *corp,
                            float *peaks, int *locs, int *ncand)
{
    int decind, decstart, decnlags, decsize, i, j;
    float xp, yp, lag_wt;
    const int dec = c->decimate, step = c->step, size = c->size, start = c->start, nlags = c->nlags;

    lag_wt = c->lag_weight / nlags;
    decnlags = 1 + (nlags / dec);
    if ((decstart = start / dec) < 1) decstart = 1;
    decind = (ind * step) / dec;
    decsize = 1 + (size / dec);

    ncc_all_lags(c, fdsdata + decind, decsize, decstart, decnlags, engref, maxloc, maxval, corp);
    pick_candidates(corp, *maxval, decstart, peaks, locs, decnlags, ncand, c->cand_thresh);

    for (i = 0; i < *ncand; i++) {
        j = locs[i] - decstart - 1;
        parabola(&corp[j], &xp, &yp);
        locs[i] = (locs[i] * dec) + (int)(0.5 + (xp * dec));
        peaks[i] = yp * (1.0 - (lag_wt * locs[i]));
    }
    if (*ncand >= c->n_cands) prune_candidates(peaks, locs, ncand, c->n_cands);

    ncc_near_lags(c, fdata + (ind * step), size, start, nlags, 7, engref, maxloc, maxval, corp,
                  locs, *ncand);
    pick_candidates(corp, *maxval, start, peaks, locs, nlags, ncand, c->cand_thresh);
    if (*ncand >= c->n_cands) prune_candidates(peaks, locs, ncand, c->n_cands);
}

/* ---- init (init_dp_f0) ---------------------------------------------------------------- */
static int rapt_init(rapt_ctx *c, double freq, long *buffsize, long *sdstep)
{
    int nframes, i, stat_wsize, agap, ind, downpatch;
    c->freq = freq;
    c->tcost = c->trans_cost;
    c->tfact_a = c->trans_amp;
    c->tfact_s = c->trans_spec;
    c->vbias = c->voice_bias;
    c->fdouble = c->double_cost;
    c->frame_int = c->frame_step;

    c->step = iround(c->frame_int * freq);
    c->size = iround(c->wind_dur * freq);
    c->frame_int = ((float)c->step) / freq;
    c->start = iround(freq / c->max_f0);
    c->stop = iround(freq / c->min_f0);
    c->nlags = c->stop - c->start + 1;
    c->ncomp = c->size + c->stop + 1;
    c->ln2 = log(2.0);
    c->size_frame_hist = (int)(RAPT_DP_HIST / c->frame_int);
    c->size_frame_out = (int)(RAPT_DP_LIMIT / c->frame_int);
    c->lagwt = c->lag_weight / c->stop;
    c->freqwt = c->freq_weight / c->frame_int;

    i = (int)(RAPT_READ_SIZE * freq);
    if (c->ncomp >= c->step) nframes = ((i - c->ncomp) / c->step) + 1;
    else nframes = i / c->step;

    downpatch = (((int)(freq * 0.005)) + 1) / 2;
    stat_wsize = (int)(RAPT_STAT_WSIZE * freq);
    agap = (int)(RAPT_STAT_AINT * freq);
    ind = (agap - stat_wsize) / 2;
    i = stat_wsize + ind;
    c->pad = downpatch + ((i > c->ncomp) ? i : c->ncomp);
    *buffsize = nframes * c->step + c->pad;
    *sdstep = nframes * c->step;

    c->decimate = (int)(freq / 2000.0);
    c->correl = (float *)calloc((size_t)c->nlags + 64, sizeof(float));
    c->dbdata = (float *)calloc((size_t)c->ncomp + 64, sizeof(float));
    stat_wsize = (int)(RAPT_STAT_WSIZE * freq);
    c->hwin479 = (float *)malloc(sizeof(float) * stat_wsize);
    c->hwin480 = (float *)malloc(sizeof(float) * stat_wsize);
    make_hanning(c->hwin479, stat_wsize - 1);
    make_hanning(c->hwin480, stat_wsize);
    c->fr = NULL; c->fr_cap = 0; c->n_fr = 0;
    c->head = -1; c->tail = 0; c->num_active = 0; c->first_time = 1;
    return 0;
}

static frame_rec *frame_at(rapt_ctx *c, int idx)
{
    if (idx >= c->fr_cap) {
        int ncap = c->fr_cap ? c->fr_cap * 2 : 256;
        while (ncap <= idx) ncap *= 2;
        c->fr = (frame_rec *)realloc(c->fr, sizeof(frame_rec) * (size_t)ncap);
        memset(c->fr + c->fr_cap, 0, sizeof(frame_rec) * (size_t)(ncap - c->fr_cap));
        c->fr_cap = ncap;
    }
    return &c->fr[idx];
}

/* ---- one streaming read: candidates, DP, commit (dp_f0) ------------------------------- */
static int rapt_read(rapt_ctx *c, const float *fdata, int buff_size, int sdstep,
                     float *f0_out /* reversed order */, int *vecsize, int last_time,
                     rapt_debug *dbg, int ds_origin)
{
    float maxval, engref, *dsdata;
    float ttemp, ftemp, ft1, ferr, err, errmin;
    int i, j, k, loc1, loc2;
    int nframes, maxloc, ncand, ncandp, minloc, samsds;
    float peaks[512];
    int locs[512];
    float *sta, *rms, *rms_ratio;

    nframes = (buff_size < c->pad) ? 0 : (buff_size - c->pad) / c->step;
    c->num_active += nframes;

    if (c->decimate <= 1) return 1;        /* fs < 4 kHz is outside the reference's use */
    samsds = ((nframes - 1) * c->step + c->ncomp) / c->decimate;
    if (samsds < 1) return 1;
    dsdata = decimate_stream(c, fdata, buff_size, sdstep, &samsds, c->first_time, last_time);
    if (dbg && dbg->ds) {
        for (i = 0; i < samsds && ds_origin + i < dbg->ds_cap; i++) dbg->ds[ds_origin + i] = dsdata[i];
    }

    sta = (float *)calloc((size_t)nframes + 1, sizeof(float));
    rms = (float *)calloc((size_t)nframes + 1, sizeof(float));
    rms_ratio = (float *)calloc((size_t)nframes + 1, sizeof(float));
    stationarity_read(c, fdata, buff_size, nframes, c->step, c->first_time, sta, rms, rms_ratio);

    if (!c->first_time && nframes > 0) c->head++;
    if (c->first_time && nframes > 0) c->head = 0;

    for (i = 0; i < nframes; i++) {
        frame_rec *cur = frame_at(c, c->head), *prev;
        fast_candidates(c, fdata, dsdata, i, &engref, &maxloc, &maxval, c->correl, peaks, locs, &ncand);

        for (j = 0; j < ncand; j++) { cur->pvals[j] = peaks[j]; cur->locs[j] = (short)locs[j]; }
        cur->locs[ncand] = -1;
        cur->pvals[ncand] = maxval;
        cur->mpvals[ncand] = c->vbias + maxval;
        for (j = 0; j < ncand; j++) {
            ftemp = 1.0 - ((float)locs[j] * c->lagwt);
            cur->mpvals[j] = 1.0 - (peaks[j] * ftemp);
        }
        /* the value each voiced candidate would emit if selected (parabolic refinement on the
         * fine NCCF, done at back-track time in the original; it only reads this frame's data) */
        for (j = 0; j < ncand; j++) {
            loc1 = locs[j];
            ftemp = loc1;
            if (loc1 > c->start && loc1 < c->stop) {
                float cormax, cprev, cnext, den;
                int jj = loc1 - c->start;
                cormax = c->correl[jj];
                cprev = c->correl[jj + 1];
                cnext = c->correl[jj - 1];
                den = (2.0 * (cprev + cnext - (2.0 * cormax)));
                if (fabs(den) > 0.000001)
                    ftemp += 2.0 - ((((5.0 * cprev) + (3.0 * cnext) - (8.0 * cormax)) / den));
            }
            cur->f0cand[j] = c->freq / ftemp;
        }
        cur->f0cand[ncand] = 0;
        ncand++;
        cur->ncands = ncand;

        if (c->head > 0) { prev = &c->fr[c->head - 1]; ncandp = prev->ncands; }
        else { prev = NULL; ncandp = 0; }
        for (k = 0; k < ncand; k++) {
            minloc = 0;
            errmin = FLT_MAX;
            if ((loc2 = cur->locs[k]) > 0) {
                for (j = 0; j < ncandp; j++) {
                    loc1 = prev->locs[j];
                    if (loc1 > 0) {
                        ftemp = log(((double)loc2) / loc1);
                        ttemp = fabs(ftemp);
                        ft1 = c->fdouble + fabs(ftemp + c->ln2);
                        if (ttemp > ft1) ttemp = ft1;
                        ft1 = c->fdouble + fabs(ftemp - c->ln2);
                        if (ttemp > ft1) ttemp = ft1;
                        ferr = ttemp * c->freqwt;
                    } else {
                        ferr = c->tcost + (c->tfact_s * sta[i]) + (c->tfact_a / rms_ratio[i]);
                    }
                    err = ferr + prev->dpvals[j];
                    if (err < errmin) { errmin = err; minloc = j; }
                }
            } else {
                for (j = 0; j < ncandp; j++) {
                    if (prev->locs[j] > 0)
                        ferr = c->tcost + (c->tfact_s * sta[i]) + (c->tfact_a * rms_ratio[i]);
                    else
                        ferr = 0.0;
                    err = ferr + prev->dpvals[j];
                    if (err < errmin) { errmin = err; minloc = j; }
                }
            }
            if (c->first_time && i == 0) {
                cur->dpvals[k] = cur->mpvals[k];
                cur->prept[k] = 0;
            } else {
                cur->dpvals[k] = errmin + cur->mpvals[k];
                cur->prept[k] = (short)minloc;
            }
        }
        if (dbg && c->head < dbg->max_frames) {
            int g = c->head;
            if (dbg->ncands) dbg->ncands[g] = ncand;
            if (dbg->stat) dbg->stat[g] = sta[i];
            if (dbg->rms_ratio) dbg->rms_ratio[g] = rms_ratio[i];
            for (j = 0; j < ncand; j++) {
                if (dbg->locs) dbg->locs[g * RAPT_CMAX + j] = cur->locs[j];
                if (dbg->pvals) dbg->pvals[g * RAPT_CMAX + j] = cur->pvals[j];
                if (dbg->mpvals) dbg->mpvals[g * RAPT_CMAX + j] = cur->mpvals[j];
                if (dbg->f0cand) dbg->f0cand[g * RAPT_CMAX + j] = cur->f0cand[j];
                if (dbg->prept) dbg->prept[g * RAPT_CMAX + j] = cur->prept[j];
                if (dbg->dpvals) dbg->dpvals[g * RAPT_CMAX + j] = cur->dpvals[j];
            }
        }
        if (i < nframes - 1) c->head++;
    }
    free(sta); free(rms); free(rms_ratio);

    /* commit: find a frame where all surviving paths agree and back-track from there */
    *vecsize = 0;
    if (c->head >= 0 && (c->num_active >= c->size_frame_hist || last_time)) {
        int num_paths, best_cand, checkpath_done = 1, cmpth = -1, frm;
        float patherrmin;
        int pcands[RAPT_CMAX];
        frame_rec *h = &c->fr[c->head];

        patherrmin = FLT_MAX;
        best_cand = 0;
        num_paths = h->ncands;
        frm = c->head;
        for (k = 0; k < num_paths; k++) {
            if (patherrmin > h->dpvals[k]) { patherrmin = h->dpvals[k]; best_cand = k; }
            pcands[k] = h->prept[k];
        }
        if (last_time) {
            cmpth = c->head;
        } else {
            while (1) {
                frm = frm - 1;
                checkpath_done = 1;
                for (k = 1; k < num_paths; k++)
                    if (pcands[0] != pcands[k]) checkpath_done = 0;
                if (!checkpath_done) {
                    for (k = 0; k < num_paths; k++) pcands[k] = c->fr[frm].prept[pcands[k]];
                } else {
                    cmpth = frm;
                    best_cand = pcands[0];
                    break;
                }
                if (frm == c->tail) {
                    if (c->num_active < c->size_frame_out) {
                        checkpath_done = 0;
                        cmpth = -1;
                    } else {
                        checkpath_done = 1;
                        cmpth = c->head;
                        if (dbg) dbg->n_forced++;
                    }
                    break;
                }
            }
        }
        i = 0;
        frm = cmpth;
        while (checkpath_done && frm != c->tail - 1) {
            frame_rec *f = &c->fr[frm];
            loc1 = f->locs[best_cand];
            f0_out[i] = (loc1 > 0) ? f->f0cand[best_cand] : 0.0f;
            best_cand = f->prept[best_cand];
            frm--;
            i++;
        }
        if (checkpath_done) {
            *vecsize = i;
            c->tail = cmpth + 1;
            c->num_active -= *vecsize;
        }
    }
    if (c->first_time) c->first_time = 0;
    return 0;
}

/* ---- public entry: the call at make_spect_f0.py:64 ------------------------------------ */
/* returns 0 ok, 2 input too short, 3 bad parameters.  out has n_out = ceil(length/hop) entries;
 * frames RAPT cannot analyse at the tail are emitted unvoiced [U]. */
int rapt_ref(const float *x, int length, double sample_freq, int frame_shift, double min_f0,
             double max_f0, double voice_bias, int otype, float *out, int n_out, rapt_debug *dbg)
{
    rapt_ctx *c = (rapt_ctx *)calloc(1, sizeof(rapt_ctx));
    long buff_size = 0, sdstep = 0, total_samps, actsize;
    int ndone = 0, count = 0, done, vecsize, i, rc = 0;
    float *fdata, *f0p;
    const float unv = (otype == 2) ? -1.0e10f : 0.0f;

    c->cand_thresh = 0.3f; c->lag_weight = 0.3f; c->freq_weight = 0.02f; c->trans_cost = 0.005f;
    c->trans_amp = 0.5f; c->trans_spec = 0.5f; c->voice_bias = (float)voice_bias;
    c->double_cost = 0.35f; c->min_f0 = (float)min_f0; c->max_f0 = (float)max_f0;
    c->frame_step = (float)((double)frame_shift / sample_freq); c->wind_dur = 0.0075f;
    c->n_cands = RAPT_CMAX;

    for (i = 0; i < n_out; i++) out[i] = unv;
    if (dbg) { dbg->n_frames = 0; dbg->n_forced = 0; }
    if ((c->max_f0 <= c->min_f0) || (c->max_f0 >= (sample_freq / 2.0)) ||
        (c->min_f0 < (sample_freq / 10000.0))) { free(c); return 3; }
    total_samps = length;
    if (total_samps < ((c->frame_step * 2.0) + c->wind_dur) * sample_freq) { free(c); return 2; }
    rapt_init(c, sample_freq, &buff_size, &sdstep);
    if (buff_size > total_samps) buff_size = total_samps;
    actsize = (buff_size < length) ? buff_size : length;
    fdata = (float *)calloc((size_t)((buff_size > sdstep) ? buff_size : sdstep) + 64, sizeof(float));
    f0p = (float *)calloc((size_t)(length / frame_shift) + 64, sizeof(float));

    while (1) {
        done = (actsize < buff_size) || (total_samps == buff_size);
        for (i = 0; i < actsize; i++) fdata[i] = x[ndone + i];
        if (rapt_read(c, fdata, (int)actsize, (int)sdstep, f0p, &vecsize, done, dbg,
                      ndone / (c->decimate > 0 ? c->decimate : 1))) { rc = 3; break; }
        for (i = vecsize - 1; i >= 0; i--) {
            float v;
            switch (otype) {
            case 1: v = f0p[i]; break;
            case 2: v = (f0p[i] == 0.0) ? -1.0e10f : (float)log(f0p[i]); break;
            default: v = (f0p[i] == 0.0) ? 0.0f : (float)(sample_freq / f0p[i]); break;
            }
            if (count < n_out) out[count] = v;
            count++;
        }
        if (done) break;
        ndone += sdstep;
        actsize = (buff_size < length - ndone) ? buff_size : length - ndone;
        total_samps -= sdstep;
        if (actsize > total_samps) actsize = total_samps;
    }
    if (dbg) dbg->n_frames = count;
    free(fdata); free(f0p);
    free(c->fr); free(c->correl); free(c->dbdata); free(c->hwin479); free(c->hwin480);
    free(c->stmem); free(c->dsout); free(c);
    return rc;
}

int rapt_ref_cmax(void) { return RAPT_CMAX; }
