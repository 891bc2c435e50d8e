/*
 * oracle/plf_oracle.c -- CPU restatement of the reference PLF "newview" step.
 * TEST INFRASTRUCTURE ONLY (see plf_oracle.h).  Parity status: PINNED (see plf_oracle.h).
 *
 * Build with  gcc -O2 -ffp-contract=off  (never let the compiler fuse a*b+c: the
 * reference result is defined by separately rounded fp32 multiplies and adds).
 *
 * Citations are relative to /root/reference/.
 */
#include "plf_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define PLF_CATS 4
#define PLF_STATES 4
#define PLF_SITE (PLF_CATS * PLF_STATES)

/* 2^-32 and 2^32 are exact in fp32; the reference compares a float against the double
 * 1.0/4294967296.0 (plf.cpp:4-6,55), which is the same predicate.                       */
static const float kMinLikelihood = 0x1p-32f;
static const float kTwoToThe32 = 0x1p+32f;

/* ((((+0 + v0*m0) + v1*m1) + v2*m2) + v3*m3): the accumulation order of plf.cpp:32-39.
 * The leading "+0 +" is kept because it turns a -0.0 first product into +0.0.            */
static inline float dot4_seq(const float *v, const float *m)
{
    float acc = 0.0f;
    for (int l = 0; l < PLF_STATES; ++l) {
        float prod = v[l] * m[l];
        acc = acc + prod;
    }
    return acc;
}

/* One rate category of one site: two 4x4 mat-vecs, element-wise product (plf.cpp:41),
 * then the back-transform x3[l] = sum_k p[k]*EV[k][l], k ascending (plf.cpp:45-50).       */
static inline void category_newview(const float *c1, const float *c2, float *c3,
                                    const float *pl, const float *pr, const float *ev)
{
    float p[PLF_STATES];
    for (int k = 0; k < PLF_STATES; ++k) {
        float a = dot4_seq(c1, pl + 4 * k);
        float b = dot4_seq(c2, pr + 4 * k);
        p[k] = a * b;
    }
    for (int l = 0; l < PLF_STATES; ++l) {
        float acc = 0.0f;
        for (int k = 0; k < PLF_STATES; ++k) {
            float prod = p[k] * ev[4 * k + l];
            acc = acc + prod;
        }
        c3[l] = acc;
    }
}

/* Underflow test + rescale of one finished site.  plf.cpp:53-64; equivalently the
 * 16-bit mask of s2mm_memDNAwindowComb.cpp:71-85.  NaN compares false -> no rescale.      */
static inline int site_rescale(float *s3)
{
    for (int e = 0; e < PLF_SITE; ++e) {
        float mag = s3[e] < 0 ? -s3[e] : s3[e];
        if (!(mag < kMinLikelihood))
            return 0;
    }
    for (int e = 0; e < PLF_SITE; ++e)
        s3[e] = s3[e] * kTwoToThe32;
    return 1;
}

static int64_t newview_impl(const float *x1, const float *x2, float *x3,
                            const float *ev, size_t ev_stride, size_t n,
                            const float *left, const float *right,
                            const int *wgt, unsigned char *scaler)
{
    int64_t total = 0;
    for (size_t i = 0; i < n; ++i) {
        const float *s1 = x1 + i * PLF_SITE;
        const float *s2 = x2 + i * PLF_SITE;
        float *s3 = x3 + i * PLF_SITE;
        for (int j = 0; j < PLF_CATS; ++j)
            category_newview(s1 + 4 * j, s2 + 4 * j, s3 + 4 * j,
                             left + 16 * j, right + 16 * j, ev + ev_stride * j);
        int scaled = site_rescale(s3);
        if (scaler)
            scaler[i] = (unsigned char)scaled;
        if (scaled)
            total += wgt ? wgt[i] : 1;
    }
    return total;
}

int64_t plf_oracle_newview(const float *x1, const float *x2, float *x3,
                           const float *ev, size_t n,
                           const float *left, const float *right,
                           const int *wgt, unsigned char *scaler)
{
    return newview_impl(x1, x2, x3, ev, 0, n, left, right, wgt, scaler);
}

int64_t plf_oracle_newview_ev4(const float *x1, const float *x2, float *x3,
                               const float *ev4, size_t n,
                               const float *left, const float *right,
                               const int *wgt, unsigned char *scaler)
{
    return newview_impl(x1, x2, x3, ev4, 16, n, left, right, wgt, scaler);
}

int64_t plf_oracle_newview_packed(const float *left_buf, const float *right_buf,
                                  int layout, size_t n, float *out,
                                  const int *wgt, unsigned char *scaler)
{
    const float *ev = left_buf;                 /* mem[0]      mm2sleft_memDNAwindowComb.cpp:32 */
    const float *pl = left_buf + 16;            /* mem[1..4]   :39-42 */
    const float *x1 = left_buf + 80;            /* mem[5+i]    :86    */
    const float *pr = layout == 0 ? right_buf + 16 : right_buf;       /* host_mem.cpp:236,239 */
    const float *x2 = layout == 0 ? right_buf + 80 : right_buf + 64;  /* host_mem.cpp:237,240 */
    return newview_impl(x1, x2, out, ev, 0, n, pl, pr, wgt, scaler);
}

int64_t plf_oracle_scaler_increment(const unsigned char *scaler, const int *wgt, size_t n)
{
    int64_t total = 0;
    for (size_t i = 0; i < n; ++i)
        total += (int64_t)scaler[i] * (wgt ? wgt[i] : 1);
    return total;
}

void plf_oracle_transpose4(const float *in, float *out)
{
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c)
            out[4 * c + r] = in[4 * r + c];
}

/* ---- multi-threaded CPU-baseline driver ------------------------------------------- */

struct mt_job {
    const float *x1, *x2, *ev, *left, *right;
    float *x3;
    const int *wgt;
    unsigned char *scaler;
    size_t n;
    int64_t total;
};

static void *mt_worker(void *arg)
{
    struct mt_job *j = (struct mt_job *)arg;
    j->total = newview_impl(j->x1, j->x2, j->x3, j->ev, 0, j->n, j->left, j->right,
                            j->wgt, j->scaler);
    return NULL;
}

int64_t plf_oracle_newview_mt(const float *x1, const float *x2, float *x3,
                              const float *ev, size_t n,
                              const float *left, const float *right,
                              const int *wgt, unsigned char *scaler, int nthreads)
{
    if (nthreads < 1)
        nthreads = 1;
    if ((size_t)nthreads > n && n > 0)
        nthreads = (int)n;
    if (nthreads == 1 || n == 0)
        return newview_impl(x1, x2, x3, ev, 0, n, left, right, wgt, scaler);

    struct mt_job *jobs = (struct mt_job *)calloc((size_t)nthreads, sizeof(*jobs));
    pthread_t *tids = (pthread_t *)calloc((size_t)nthreads, sizeof(*tids));
    size_t chunk = (n + (size_t)nthreads - 1) / (size_t)nthreads;
    int launched = 0;
    for (int t = 0; t < nthreads; ++t) {
        size_t lo = (size_t)t * chunk;
        if (lo >= n)
            break;
        size_t cnt = n - lo < chunk ? n - lo : chunk;
        struct mt_job *j = &jobs[t];
        j->x1 = x1 + lo * PLF_SITE;
        j->x2 = x2 + lo * PLF_SITE;
        j->x3 = x3 + lo * PLF_SITE;
        j->ev = ev;
        j->left = left;
        j->right = right;
        j->wgt = wgt ? wgt + lo : NULL;
        j->scaler = scaler ? scaler + lo : NULL;
        j->n = cnt;
        pthread_create(&tids[t], NULL, mt_worker, j);
        ++launched;
    }
    int64_t total = 0;
    for (int t = 0; t < launched; ++t) {
        pthread_join(tids[t], NULL);
        total += jobs[t].total;
    }
    free(jobs);
    free(tids);
    return total;
}

/* ------------------------------------------------------------------------------------------
 * General state count (the reference's STATES knob, README.md:36,67; "Implement protein-based
 * PLF" is an open to-do there, README.md:202).  The loops of app/src/plf.cpp:19-65 with every
 * literal 4 that means "states" replaced by S (and 16 by S*S / 4*S): same accumulation order
 * (x3 zeroed, l-then-k for the branch sums, "x3[j*S+l] += p[k]*EV[S*k+l]" with k outer), same
 * threshold on all 4*S entries of the site.  Parity status: PINNED for S = 4 (tests assert it
 * is bit-identical to the reference's plf()); PARITY UNPINNED for S = 20 -- the same code path,
 * but there is no protein implementation, golden vector or test in the reference to pin against.
 * ------------------------------------------------------------------------------------------ */
int64_t plf_oracle_newview_states(int S, const float *x1, const float *x2, float *x3,
                                  const float *ev, size_t n,
                                  const float *left, const float *right,
                                  const int *wgt, unsigned char *scaler)
{
    if (S < 1 || S > 64)
        return -1;
    const size_t site = (size_t)PLF_CATS * (size_t)S;
    int64_t total = 0;
    float p[64];
    for (size_t i = 0; i < n; ++i) {
        const float *s1 = x1 + i * site, *s2 = x2 + i * site;
        float *s3 = x3 + i * site;
        for (size_t e = 0; e < site; ++e)
            s3[e] = 0.0f;
        for (int j = 0; j < PLF_CATS; ++j) {
            for (int k = 0; k < S; ++k) {
                float a = 0.0f, b = 0.0f;
                for (int l = 0; l < S; ++l) {
                    float pa = s1[j * S + l] * left[(j * S + k) * S + l];
                    float pb = s2[j * S + l] * right[(j * S + k) * S + l];
                    a = a + pa;
                    b = b + pb;
                }
                p[k] = a * b;
            }
            for (int k = 0; k < S; ++k)
                for (int l = 0; l < S; ++l) {
                    float prod = p[k] * ev[S * k + l];
                    s3[j * S + l] = s3[j * S + l] + prod;
                }
        }
        int scale = 1;
        for (size_t e = 0; scale && e < site; ++e) {
            float mag = s3[e] < 0 ? -s3[e] : s3[e];
            scale = mag < kMinLikelihood;
        }
        if (scale) {
            for (size_t e = 0; e < site; ++e)
                s3[e] = s3[e] * kTwoToThe32;
            total += wgt ? (int64_t)wgt[i] : 1;
        }
        if (scaler)
            scaler[i] = (unsigned char)scale;
    }
    return total;
}

struct states_job {
    int S;
    const float *x1, *x2, *ev, *left, *right;
    float *x3;
    const int *wgt;
    unsigned char *scaler;
    size_t n;
    int64_t result;
};

static void *states_worker(void *arg)
{
    struct states_job *j = (struct states_job *)arg;
    j->result = plf_oracle_newview_states(j->S, j->x1, j->x2, j->x3, j->ev, j->n, j->left, j->right, j->wgt, j->scaler);
    return NULL;
}

int64_t plf_oracle_newview_states_mt(int S, const float *x1, const float *x2, float *x3,
                                     const float *ev, size_t n, const float *left, const float *right,
                                     const int *wgt, unsigned char *scaler, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    if ((size_t)nthreads > n) nthreads = n ? (int)n : 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    struct states_job *jobs = (struct states_job *)malloc(sizeof(struct states_job) * (size_t)nthreads);
    const size_t per = (n + (size_t)nthreads - 1) / (size_t)nthreads, site = (size_t)PLF_CATS * (size_t)S;
    for (int t = 0; t < nthreads; ++t) {
        size_t lo = per * (size_t)t < n ? per * (size_t)t : n;
        size_t cnt = lo + per <= n ? per : n - lo;
        struct states_job job = {S, x1 + lo * site, x2 + lo * site, ev, left, right, x3 + lo * site,
                                 wgt ? wgt + lo : NULL, scaler ? scaler + lo : NULL, cnt, 0};
        jobs[t] = job;
        pthread_create(&th[t], NULL, states_worker, &jobs[t]);
    }
    int64_t total = 0;
    for (int t = 0; t < nthreads; ++t) {
        pthread_join(th[t], NULL);
        total += jobs[t].result;
    }
    free(th);
    free(jobs);
    return total;
}
