/*
 * oracle/fft_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Float32 real FFT (unnormalised r2c / c2r, FFTW sign conventions) for the
 * zita-convolver restatement.  Method: length-2m real transform through one
 * length-m complex Stockham autosort FFT (radix 4, one radix-2 pass when
 * log2(m) is odd) on split re/im arrays plus the usual even/odd unpacking.
 * Twiddles are computed in double precision and rounded once.
 *
 * This is our own code; zita-convolver calls FFTW3f here
 * (zita-convolver 4.0.3 zita-convolver.cc, Convlevel::process /
 * Convlevel::impdata_write -- third-party, not in /root/reference, absent from
 * this image).  FFTW's output differs from this one in the last bits; both
 * approximate the same DFT.  PARITY UNPINNED at the bit level for that reason.
 */
#include "fft_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

struct offt_plan {
    int nreal;
    int m;            /* complex length = nreal / 2 */
    int nstage;
    int stage_n[16];  /* remaining length at each radix-4 stage */
    float *tw[16];    /* per stage: 6 arrays of n/4 (w1r w1i w2r w2i w3r w3i), forward sign */
    float *pr, *pi;   /* exp(-i pi k / m), k = 0..m/2 */
    float *ar, *ai, *br, *bi;
};

static void *xalloc(size_t bytes) {
    void *p = NULL;
    if (posix_memalign(&p, 64, bytes ? bytes : 64)) return NULL;
    return p;
}

offt_plan *offt_plan_create(int nreal) {
    if (nreal < 8 || (nreal & (nreal - 1))) return NULL;
    offt_plan *p = (offt_plan *)calloc(1, sizeof(*p));
    if (!p) return NULL;
    const double PI = 3.14159265358979323846264338327950288;
    p->nreal = nreal;
    p->m = nreal / 2;
    int n = p->m, st = 0;
    while (n >= 4) {
        const int q = n / 4;
        float *t = (float *)xalloc(sizeof(float) * 6 * q);
        for (int k = 0; k < q; k++) {
            for (int r = 1; r <= 3; r++) {
                const double a = -2.0 * PI * (double)r * (double)k / (double)n;
                t[(2 * (r - 1)) * q + k] = (float)cos(a);
                t[(2 * (r - 1) + 1) * q + k] = (float)sin(a);
            }
        }
        p->stage_n[st] = n;
        p->tw[st] = t;
        st++;
        n /= 4;
    }
    p->nstage = st;
    const int h = p->m / 2;
    p->pr = (float *)xalloc(sizeof(float) * (h + 1));
    p->pi = (float *)xalloc(sizeof(float) * (h + 1));
    for (int k = 0; k <= h; k++) {
        const double a = -PI * (double)k / (double)p->m;
        p->pr[k] = (float)cos(a);
        p->pi[k] = (float)sin(a);
    }
    p->ar = (float *)xalloc(sizeof(float) * p->m);
    p->ai = (float *)xalloc(sizeof(float) * p->m);
    p->br = (float *)xalloc(sizeof(float) * p->m);
    p->bi = (float *)xalloc(sizeof(float) * p->m);
    return p;
}

void offt_plan_destroy(offt_plan *p) {
    if (!p) return;
    for (int i = 0; i < p->nstage; i++) free(p->tw[i]);
    free(p->pr); free(p->pi);
    free(p->ar); free(p->ai); free(p->br); free(p->bi);
    free(p);
}

/* One Stockham radix-4 pass.  sg = -1 forward, +1 inverse. */
static void pass4(int n, int s, const float *restrict tw, float sg,
                  const float *restrict xr, const float *restrict xi,
                  float *restrict yr, float *restrict yi) {
    const int q4 = n / 4;
    const float *w1r = tw, *w1i = tw + q4, *w2r = tw + 2 * q4, *w2i = tw + 3 * q4,
                *w3r = tw + 4 * q4, *w3i = tw + 5 * q4;
    if (s == 1) {
        /* first pass: contiguous reads over p, stride-4 writes */
        for (int p = 0; p < q4; p++) {
            const float ar = xr[p], ai = xi[p];
            const float br = xr[p + q4], bi = xi[p + q4];
            const float cr = xr[p + 2 * q4], ci = xi[p + 2 * q4];
            const float dr = xr[p + 3 * q4], di = xi[p + 3 * q4];
            const float apcr = ar + cr, apci = ai + ci, amcr = ar - cr, amci = ai - ci;
            const float bpdr = br + dr, bpdi = bi + di, bmdr = br - dr, bmdi = bi - di;
            const float jr = -sg * bmdi, ji = sg * bmdr; /* (sg*i) * (b-d) */
            const float t1r = amcr + jr, t1i = amci + ji;
            const float t2r = apcr - bpdr, t2i = apci - bpdi;
            const float t3r = amcr - jr, t3i = amci - ji;
            const float u1r = w1r[p], u1i = -sg * w1i[p];
            const float u2r = w2r[p], u2i = -sg * w2i[p];
            const float u3r = w3r[p], u3i = -sg * w3i[p];
            yr[4 * p] = apcr + bpdr;
            yi[4 * p] = apci + bpdi;
            yr[4 * p + 1] = t1r * u1r - t1i * u1i;
            yi[4 * p + 1] = t1r * u1i + t1i * u1r;
            yr[4 * p + 2] = t2r * u2r - t2i * u2i;
            yi[4 * p + 2] = t2r * u2i + t2i * u2r;
            yr[4 * p + 3] = t3r * u3r - t3i * u3i;
            yi[4 * p + 3] = t3r * u3i + t3i * u3r;
        }
        return;
    }
    for (int p = 0; p < q4; p++) {
        const float u1r = w1r[p], u1i = -sg * w1i[p];
        const float u2r = w2r[p], u2i = -sg * w2i[p];
        const float u3r = w3r[p], u3i = -sg * w3i[p];
        const float *x0r = xr + s * p, *x0i = xi + s * p;
        const float *x1r = x0r + s * q4, *x1i = x0i + s * q4;
        const float *x2r = x1r + s * q4, *x2i = x1i + s * q4;
        const float *x3r = x2r + s * q4, *x3i = x2i + s * q4;
        float *y0r = yr + s * 4 * p, *y0i = yi + s * 4 * p;
        float *y1r = y0r + s, *y1i = y0i + s;
        float *y2r = y1r + s, *y2i = y1i + s;
        float *y3r = y2r + s, *y3i = y2i + s;
        for (int q = 0; q < s; q++) {
            const float ar = x0r[q], ai = x0i[q];
            const float br = x1r[q], bi = x1i[q];
            const float cr = x2r[q], ci = x2i[q];
            const float dr = x3r[q], di = x3i[q];
            const float apcr = ar + cr, apci = ai + ci, amcr = ar - cr, amci = ai - ci;
            const float bpdr = br + dr, bpdi = bi + di, bmdr = br - dr, bmdi = bi - di;
            const float jr = -sg * bmdi, ji = sg * bmdr;
            const float t1r = amcr + jr, t1i = amci + ji;
            const float t2r = apcr - bpdr, t2i = apci - bpdi;
            const float t3r = amcr - jr, t3i = amci - ji;
            y0r[q] = apcr + bpdr;
            y0i[q] = apci + bpdi;
            y1r[q] = t1r * u1r - t1i * u1i;
            y1i[q] = t1r * u1i + t1i * u1r;
            y2r[q] = t2r * u2r - t2i * u2i;
            y2i[q] = t2r * u2i + t2i * u2r;
            y3r[q] = t3r * u3r - t3i * u3i;
            y3i[q] = t3r * u3i + t3i * u3r;
        }
    }
}

/* Final radix-2 pass (remaining n == 2, twiddle 1). */
static void pass2(int s, const float *restrict xr, const float *restrict xi,
                  float *restrict yr, float *restrict yi) {
    for (int q = 0; q < s; q++) {
        const float ar = xr[q], ai = xi[q], br = xr[q + s], bi = xi[q + s];
        yr[q] = ar + br; yi[q] = ai + bi;
        yr[q + s] = ar - br; yi[q + s] = ai - bi;
    }
}

/* In: (ar,ai).  Returns 0 if the result is in (ar,ai), 1 if in (br,bi). */
static int cfft(offt_plan *p, float sg) {
    float *xr = p->ar, *xi = p->ai, *yr = p->br, *yi = p->bi;
    int s = 1, where = 0;
    for (int st = 0; st < p->nstage; st++) {
        pass4(p->stage_n[st], s, p->tw[st], sg, xr, xi, yr, yi);
        float *t;
        t = xr; xr = yr; yr = t;
        t = xi; xi = yi; yi = t;
        where ^= 1;
        s *= 4;
    }
    if (s < p->m) { /* one radix-2 pass left */
        pass2(s, xr, xi, yr, yi);
        where ^= 1;
    }
    return where;
}

void offt_r2c(offt_plan *p, const float *in, float *out) {
    const int m = p->m;
    for (int n = 0; n < m; n++) {
        p->ar[n] = in[2 * n];
        p->ai[n] = in[2 * n + 1];
    }
    const int w = cfft(p, -1.0f);
    const float *zr = w ? p->br : p->ar, *zi = w ? p->bi : p->ai;
    out[0] = zr[0] + zi[0];
    out[1] = 0.0f;
    out[2 * m] = zr[0] - zi[0];
    out[2 * m + 1] = 0.0f;
    const int h = m / 2;
    for (int k = 1; k <= h; k++) {
        const int k2 = m - k;
        const float er = 0.5f * (zr[k] + zr[k2]), ei = 0.5f * (zi[k] - zi[k2]);
        const float dr = 0.5f * (zr[k] - zr[k2]), di = 0.5f * (zi[k] + zi[k2]);
        /* o = -i * w * d, w = exp(-i pi k / m) */
        const float wr = p->pr[k], wi = p->pi[k];
        const float tr = wr * dr - wi * di, ti = wr * di + wi * dr;
        const float orr = ti, oi = -tr;
        out[2 * k] = er + orr;
        out[2 * k + 1] = ei + oi;
        /* X[m-k] = conj(E[k]) - conj(O[k]) ... derived from the same pair */
        out[2 * k2] = er - orr;
        out[2 * k2 + 1] = -(ei - oi);
    }
}

void offt_c2r(offt_plan *p, const float *in, float *out) {
    const int m = p->m;
    const int h = m / 2;
    /* Z[k] = (Y[k] + conj Y[m-k]) + i w^-1 (Y[k] - conj Y[m-k]) */
    {
        const float y0 = in[0], ym = in[2 * m];
        p->ar[0] = y0 + ym;
        p->ai[0] = y0 - ym;
    }
    for (int k = 1; k <= h; k++) {
        const int k2 = m - k;
        const float yr = in[2 * k], yi = in[2 * k + 1];
        const float cr = in[2 * k2], ci = in[2 * k2 + 1];
        const float er = yr + cr, ei = yi - ci;
        const float dr = yr - cr, di = yi + ci;
        const float wr = p->pr[k], wi = -p->pi[k]; /* exp(+i pi k / m) */
        const float tr = wr * dr - wi * di, ti = wr * di + wi * dr;
        /* i * t */
        const float orr = -ti, oi = tr;
        p->ar[k] = er + orr;
        p->ai[k] = ei + oi;
        /* Z[m-k] = conj(E[k]) + i * conj(w^-1)... = conj(E) - conj(iO)... */
        p->ar[k2] = er - orr;
        p->ai[k2] = -(ei - oi);
    }
    const int w = cfft(p, +1.0f);
    const float *zr = w ? p->br : p->ar, *zi = w ? p->bi : p->ai;
    for (int n = 0; n < m; n++) {
        out[2 * n] = zr[n];
        out[2 * n + 1] = zi[n];
    }
}
