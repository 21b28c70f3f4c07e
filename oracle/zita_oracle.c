/*
 * oracle/zita_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * See zita_oracle.h for scope, provenance and the "parity unpinned" statement.
 *
 * Restates, for the single-level case folve uses, these pieces of
 * zita-convolver 4.0.3 (zita-convolver.cc; third-party, not in /root/reference):
 *   Convproc::configure / impdata_create / impdata_link / reset /
 *   start_process / process / stop_process / cleanup, and
 *   Convlevel::configure / impdata_write / impdata_link / reset / readout /
 *   process / findmacnode.
 * Structure (linked lists of input, output and MAC nodes, three rotating
 * output buffers, spectra ring indexed by _ptind, 0.5/parsize folded into the
 * filter spectra, lazily allocated filter partitions) follows that source.
 */
#include "zita_oracle.h"
#include "fft_oracle.h"

#include <stdlib.h>
#include <string.h>

typedef struct zo_inpnode {
    struct zo_inpnode *next;
    float **ffta; /* [npar] -> (parsize+1) complex */
    uint32_t npar;
    uint32_t inp;
} zo_inpnode;

typedef struct zo_macnode {
    struct zo_macnode *next;
    zo_inpnode *inpn;
    struct zo_macnode *link;
    float **fftb; /* [npar] -> (parsize+1) complex or NULL */
    uint32_t npar;
} zo_macnode;

typedef struct zo_outnode {
    struct zo_outnode *next;
    zo_macnode *list;
    float *buff[3];
    uint32_t out;
} zo_outnode;

typedef struct {
    /* Convlevel */
    uint32_t offs, npar, parsize, outsize, outoffs, inpsize, inpoffs, ptind, opind;
    zo_inpnode *inp_list;
    zo_outnode *out_list;
    offt_plan *plan;
    float *time_data; /* 2*parsize */
    float *prep_data; /* 2*parsize */
    float *freq_data; /* (parsize+1) complex */
    float **inpbuff;
    float **outbuff;
} zo_level;

struct zo_convproc {
    int state;
    int reset_is_fresh;
    uint32_t options;
    uint32_t ninp, nout, quantum, minpart, maxpart, nlevels, inpsize, inpoffs, outoffs, latecnt;
    float *inpbuff[ZO_MAXINP];
    float *outbuff[ZO_MAXOUT];
    zo_level *lev; /* single level (quantum == minpart == maxpart) */
};

static float *zalloc_f(size_t n) {
    void *p = NULL;
    if (posix_memalign(&p, 64, (n ? n : 1) * sizeof(float))) return NULL;
    memset(p, 0, (n ? n : 1) * sizeof(float));
    return (float *)p;
}

/* ---- Convlevel ---------------------------------------------------------- */

static zo_level *level_new(uint32_t offs, uint32_t npar, uint32_t parsize) {
    zo_level *L = (zo_level *)calloc(1, sizeof(*L));
    if (!L) return NULL;
    L->offs = offs;
    L->npar = npar;
    L->parsize = parsize;
    L->plan = offt_plan_create((int)(2 * parsize));
    L->time_data = zalloc_f(2 * parsize);
    L->prep_data = zalloc_f(2 * parsize);
    L->freq_data = zalloc_f(2 * (parsize + 1));
    if (!L->plan || !L->time_data || !L->prep_data || !L->freq_data) return NULL;
    return L;
}

static void macnode_free_fftb(zo_macnode *M) {
    if (!M->fftb) return;
    for (uint32_t k = 0; k < M->npar; k++) free(M->fftb[k]);
    free(M->fftb);
    M->fftb = NULL;
    M->npar = 0;
}

static void level_free(zo_level *L) {
    if (!L) return;
    zo_inpnode *X = L->inp_list;
    while (X) {
        zo_inpnode *n = X->next;
        for (uint32_t k = 0; k < X->npar; k++) free(X->ffta[k]);
        free(X->ffta);
        free(X);
        X = n;
    }
    zo_outnode *Y = L->out_list;
    while (Y) {
        zo_outnode *n = Y->next;
        zo_macnode *M = Y->list;
        while (M) {
            zo_macnode *mn = M->next;
            macnode_free_fftb(M);
            free(M);
            M = mn;
        }
        for (int i = 0; i < 3; i++) free(Y->buff[i]);
        free(Y);
        Y = n;
    }
    offt_plan_destroy(L->plan);
    free(L->time_data);
    free(L->prep_data);
    free(L->freq_data);
    free(L);
}

/* Convlevel::findmacnode: nodes are created on demand and PREPENDED. */
static zo_macnode *level_findmacnode(zo_level *L, uint32_t inp, uint32_t out, int create) {
    zo_inpnode *X;
    zo_outnode *Y;
    zo_macnode *M;

    for (X = L->inp_list; X && X->inp != inp; X = X->next) {}
    if (!X) {
        if (!create) return NULL;
        X = (zo_inpnode *)calloc(1, sizeof(*X));
        X->inp = inp;
        X->npar = L->npar;
        X->ffta = (float **)calloc(L->npar, sizeof(float *));
        for (uint32_t k = 0; k < L->npar; k++) X->ffta[k] = zalloc_f(2 * (L->parsize + 1));
        X->next = L->inp_list;
        L->inp_list = X;
    }
    for (Y = L->out_list; Y && Y->out != out; Y = Y->next) {}
    if (!Y) {
        if (!create) return NULL;
        Y = (zo_outnode *)calloc(1, sizeof(*Y));
        Y->out = out;
        for (int i = 0; i < 3; i++) Y->buff[i] = zalloc_f(L->parsize);
        Y->next = L->out_list;
        L->out_list = Y;
    }
    for (M = Y->list; M && M->inpn != X; M = M->next) {}
    if (!M) {
        if (!create) return NULL;
        M = (zo_macnode *)calloc(1, sizeof(*M));
        M->inpn = X;
        M->next = Y->list;
        Y->list = M;
    }
    return M;
}

/* Convlevel::impdata_write (create == true path is the one folve reaches). */
static void level_impdata_write(zo_level *L, uint32_t inp, uint32_t out, int32_t step,
                                const float *data, int32_t i0, int32_t i1, int create) {
    const int32_t n = i1 - i0;
    i0 = (int32_t)L->offs - i0;
    i1 = i0 + (int32_t)(L->npar * L->parsize);
    if ((i0 >= n) || (i1 <= 0)) return;

    zo_macnode *M;
    if (create) {
        M = level_findmacnode(L, inp, out, 1);
        if (M == NULL || M->link) return;
        if (M->fftb == NULL) {
            M->npar = L->npar;
            M->fftb = (float **)calloc(L->npar, sizeof(float *));
        }
    } else {
        M = level_findmacnode(L, inp, out, 0);
        if (M == NULL || M->link || M->fftb == NULL) return;
    }

    const float norm = 0.5f / (float)L->parsize;
    const int32_t ps = (int32_t)L->parsize;
    for (uint32_t k = 0; k < L->npar; k++) {
        i1 = i0 + ps;
        if ((i0 < n) && (i1 > 0)) {
            float *fftb = M->fftb[k];
            if (fftb == NULL && create) M->fftb[k] = fftb = zalloc_f(2 * (L->parsize + 1));
            if (fftb && data) {
                memset(L->prep_data, 0, 2 * L->parsize * sizeof(float));
                const int32_t j0 = (i0 < 0) ? 0 : i0;
                const int32_t j1 = (i1 > n) ? n : i1;
                for (int32_t j = j0; j < j1; j++) L->prep_data[j - i0] = norm * data[(long)j * step];
                offt_r2c(L->plan, L->prep_data, L->freq_data);
                for (int32_t j = 0; j <= ps; j++) {
                    fftb[2 * j] += L->freq_data[2 * j];
                    fftb[2 * j + 1] += L->freq_data[2 * j + 1];
                }
            }
        }
        i0 = i1;
    }
}

/* Convlevel::impdata_link */
static void level_impdata_link(zo_level *L, uint32_t inp1, uint32_t out1, uint32_t inp2, uint32_t out2) {
    zo_macnode *M1 = level_findmacnode(L, inp1, out1, 0);
    if (!M1) return;
    zo_macnode *M2 = level_findmacnode(L, inp2, out2, 1);
    macnode_free_fftb(M2);
    M2->link = M1;
}

/* Convlevel::reset */
static void level_reset(zo_level *L, uint32_t inpsize, uint32_t outsize, float **inpbuff, float **outbuff) {
    L->inpsize = inpsize;
    L->outsize = outsize;
    L->inpbuff = inpbuff;
    L->outbuff = outbuff;
    for (zo_inpnode *X = L->inp_list; X; X = X->next)
        for (uint32_t i = 0; i < L->npar; i++) memset(X->ffta[i], 0, 2 * (L->parsize + 1) * sizeof(float));
    for (zo_outnode *Y = L->out_list; Y; Y = Y->next)
        for (int i = 0; i < 3; i++) memset(Y->buff[i], 0, L->parsize * sizeof(float));
    /* _parsize == _outsize always holds for the single level */
    L->outoffs = 0;
    L->inpoffs = 0;
    L->ptind = 0;
    L->opind = 0;
}

/* The inner loop of Convlevel::process: freq_data[k] += ffta[k] * fftb[k]. */
static void mac_row(float *restrict fd, const float *restrict a, const float *restrict b, uint32_t n) {
    for (uint32_t k = 0; k < n; k++) {
        fd[2 * k] += a[2 * k] * b[2 * k] - a[2 * k + 1] * b[2 * k + 1];
        fd[2 * k + 1] += a[2 * k] * b[2 * k + 1] + a[2 * k + 1] * b[2 * k];
    }
}

/* Convlevel::process(skip = false) */
static void level_process(zo_level *L) {
    uint32_t i1 = L->inpoffs, n1 = L->parsize, n2 = 0;
    L->inpoffs = i1 + n1;
    if (L->inpoffs >= L->inpsize) {
        L->inpoffs -= L->inpsize;
        n2 = L->inpoffs;
        n1 -= n2;
    }
    const uint32_t opi1 = (L->opind + 1) % 3;
    const uint32_t opi2 = (L->opind + 2) % 3;
    const uint32_t ps = L->parsize;

    for (zo_inpnode *X = L->inp_list; X; X = X->next) {
        const float *inpd = L->inpbuff[X->inp];
        if (n1) memcpy(L->time_data, inpd + i1, n1 * sizeof(float));
        if (n2) memcpy(L->time_data + n1, inpd, n2 * sizeof(float));
        memset(L->time_data + ps, 0, ps * sizeof(float));
        offt_r2c(L->plan, L->time_data, X->ffta[L->ptind]);
    }

    for (zo_outnode *Y = L->out_list; Y; Y = Y->next) {
        float *fd = L->freq_data;
        memset(fd, 0, 2 * (ps + 1) * sizeof(float));
        for (zo_macnode *M = Y->list; M; M = M->next) {
            zo_inpnode *X = M->inpn;
            uint32_t i = L->ptind;
            for (uint32_t j = 0; j < L->npar; j++) {
                const float *ffta = X->ffta[i];
                const float *fftb = M->link ? (M->link->fftb ? M->link->fftb[j] : NULL)
                                            : (M->fftb ? M->fftb[j] : NULL);
                if (fftb) mac_row(fd, ffta, fftb, ps + 1);
                if (i == 0) i = L->npar;
                i--;
            }
        }
        offt_c2r(L->plan, fd, L->time_data);
        float *outd = Y->buff[opi1];
        for (uint32_t k = 0; k < ps; k++) outd[k] += L->time_data[k];
        outd = Y->buff[opi2];
        memcpy(outd, L->time_data + ps, ps * sizeof(float));
    }

    L->ptind++;
    if (L->ptind == L->npar) L->ptind = 0;
}

/* Convlevel::readout for an unthreaded level (_stat != ST_PROC). */
static int level_readout(zo_level *L) {
    L->outoffs += L->outsize;
    if (L->outoffs == L->parsize) {
        L->outoffs = 0;
        level_process(L);
        if (++L->opind == 3) L->opind = 0;
    }
    for (zo_outnode *Y = L->out_list; Y; Y = Y->next) {
        const float *p = Y->buff[L->opind] + L->outoffs;
        float *q = L->outbuff[Y->out];
        for (uint32_t i = 0; i < L->outsize; i++) q[i] += p[i];
    }
    return 0;
}

/* ---- Convproc ----------------------------------------------------------- */

zo_convproc *zo_new(void) {
    zo_convproc *p = (zo_convproc *)calloc(1, sizeof(*p));
    if (p) p->state = ZO_ST_IDLE;
    return p;
}

void zo_delete(zo_convproc *p) {
    if (!p) return;
    zo_stop_process(p);
    zo_cleanup(p);
    free(p);
}

void zo_set_reset_is_fresh(zo_convproc *p, int on) { p->reset_is_fresh = on; }
void zo_set_options(zo_convproc *p, uint32_t options) { p->options = options; }
int zo_state(const zo_convproc *p) { return p->state; }

int zo_configure(zo_convproc *p, uint32_t ninp, uint32_t nout, uint32_t maxsize,
                 uint32_t quantum, uint32_t minpart, uint32_t maxpart, float density) {
    (void)density; /* only steers the multi-level size sequence, unused with one level */
    if (p->state != ZO_ST_IDLE) return ZO_BAD_STATE;
    if ((ninp < 1) || (ninp > ZO_MAXINP) || (nout < 1) || (nout > ZO_MAXOUT) ||
        (quantum & (quantum - 1)) || (quantum < ZO_MINQUANT) || (quantum > ZO_MAXQUANT) ||
        (minpart & (minpart - 1)) || (minpart < ZO_MINPART) || (minpart < quantum) ||
        (minpart > ZO_MAXDIVIS * quantum) || (maxpart & (maxpart - 1)) ||
        (maxpart > ZO_MAXPART) || (maxpart < minpart))
        return ZO_BAD_PARAM;
    /* Restated only for folve's call: one level (zita-fconfig.cc:80-81). */
    if (quantum != minpart || minpart != maxpart) return ZO_BAD_PARAM;

    const uint32_t size = quantum;
    /* for (offs = pind = 0; offs < maxsize; pind++): with size == maxpart the
     * first level takes npar = ceil(maxsize / size) and covers everything. */
    if (maxsize == 0) {
        /* the loop body never runs: zero levels, buffers still allocated */
        p->lev = NULL;
        p->nlevels = 0;
    } else {
        const uint32_t npar = (maxsize + size - 1) / size;
        p->lev = level_new(0, npar, size);
        if (!p->lev) return ZO_MEM_ALLOC;
        p->nlevels = 1;
    }
    p->ninp = ninp;
    p->nout = nout;
    p->quantum = quantum;
    p->minpart = minpart;
    p->maxpart = size;
    p->latecnt = 0;
    p->inpsize = 2 * size;
    for (uint32_t i = 0; i < ninp; i++) p->inpbuff[i] = zalloc_f(p->inpsize);
    for (uint32_t i = 0; i < nout; i++) p->outbuff[i] = zalloc_f(p->minpart);
    p->state = ZO_ST_STOP;
    return 0;
}

int zo_impdata_create(zo_convproc *p, uint32_t inp, uint32_t out, int32_t step,
                      const float *data, int32_t ind0, int32_t ind1) {
    if (p->state != ZO_ST_STOP) return ZO_BAD_STATE;
    if ((inp >= p->ninp) || (out >= p->nout)) return ZO_BAD_PARAM;
    if (p->lev) level_impdata_write(p->lev, inp, out, step, data, ind0, ind1, 1);
    return 0;
}

int zo_impdata_link(zo_convproc *p, uint32_t inp1, uint32_t out1, uint32_t inp2, uint32_t out2) {
    if ((inp1 >= p->ninp) || (out1 >= p->nout)) return ZO_BAD_PARAM;
    if ((inp2 >= p->ninp) || (out2 >= p->nout)) return ZO_BAD_PARAM;
    if ((inp1 == inp2) && (out1 == out2)) return ZO_BAD_PARAM;
    if (p->state != ZO_ST_STOP) return ZO_BAD_STATE;
    if (p->lev) level_impdata_link(p->lev, inp1, out1, inp2, out2);
    return 0;
}

int zo_reset(zo_convproc *p) {
    if (p->state == ZO_ST_IDLE) return ZO_BAD_STATE;
    for (uint32_t k = 0; k < p->ninp; k++) memset(p->inpbuff[k], 0, p->inpsize * sizeof(float));
    for (uint32_t k = 0; k < p->nout; k++) memset(p->outbuff[k], 0, p->minpart * sizeof(float));
    if (p->lev) level_reset(p->lev, p->inpsize, p->minpart, p->inpbuff, p->outbuff);
    return 0;
}

int zo_start_process(zo_convproc *p, int abspri, int policy) {
    (void)abspri; (void)policy; /* level 0 runs in the caller's thread when minpart == quantum */
    if (p->state != ZO_ST_STOP) {
        if (!(p->reset_is_fresh && p->state == ZO_ST_PROC)) return ZO_BAD_STATE;
    }
    p->latecnt = 0;
    p->inpoffs = 0;
    p->outoffs = 0;
    zo_reset(p);
    p->state = ZO_ST_PROC;
    return 0;
}

int zo_process(zo_convproc *p) {
    if (p->state != ZO_ST_PROC) return 0;
    p->inpoffs += p->quantum;
    if (p->inpoffs == p->inpsize) p->inpoffs = 0;
    p->outoffs += p->quantum;
    if (p->outoffs == p->minpart) {
        p->outoffs = 0;
        for (uint32_t k = 0; k < p->nout; k++) memset(p->outbuff[k], 0, p->minpart * sizeof(float));
        if (p->lev) level_readout(p->lev);
    }
    return 0;
}

int zo_stop_process(zo_convproc *p) {
    if (p->state != ZO_ST_PROC) return ZO_BAD_STATE;
    p->state = ZO_ST_STOP; /* no worker threads to join for an unthreaded level */
    return 0;
}

int zo_cleanup(zo_convproc *p) {
    if (p->state == ZO_ST_IDLE) return 0;
    for (uint32_t k = 0; k < p->ninp; k++) { free(p->inpbuff[k]); p->inpbuff[k] = NULL; }
    for (uint32_t k = 0; k < p->nout; k++) { free(p->outbuff[k]); p->outbuff[k] = NULL; }
    level_free(p->lev);
    p->lev = NULL;
    p->state = ZO_ST_IDLE;
    p->options = 0;
    p->ninp = p->nout = p->quantum = p->minpart = p->maxpart = p->nlevels = 0;
    p->inpsize = p->inpoffs = p->outoffs = p->latecnt = 0;
    return 0;
}

float *zo_inpdata(const zo_convproc *p, uint32_t inp) {
    if (inp >= ZO_MAXINP || !p->inpbuff[inp]) return NULL;
    return p->inpbuff[inp] + p->inpoffs;
}

float *zo_outdata(const zo_convproc *p, uint32_t out) {
    if (out >= ZO_MAXOUT || !p->outbuff[out]) return NULL;
    return p->outbuff[out] + p->outoffs;
}

uint32_t zo_parsize(const zo_convproc *p) { return p->lev ? p->lev->parsize : 0; }
uint32_t zo_npar(const zo_convproc *p) { return p->lev ? p->lev->npar : 0; }
uint32_t zo_ptind(const zo_convproc *p) { return p->lev ? p->lev->ptind : 0; }

const float *zo_fftb(const zo_convproc *p, uint32_t inp, uint32_t out, uint32_t j) {
    if (!p->lev || j >= p->lev->npar) return NULL;
    zo_macnode *M = level_findmacnode(p->lev, inp, out, 0);
    if (!M) return NULL;
    if (M->link) return M->link->fftb ? M->link->fftb[j] : NULL;
    return M->fftb ? M->fftb[j] : NULL;
}

const float *zo_ffta(const zo_convproc *p, uint32_t inp, uint32_t slot) {
    if (!p->lev || slot >= p->lev->npar) return NULL;
    for (zo_inpnode *X = p->lev->inp_list; X; X = X->next)
        if (X->inp == inp) return X->ffta[slot];
    return NULL;
}
