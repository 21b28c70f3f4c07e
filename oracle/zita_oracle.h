/*
 * oracle/zita_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the one third-party algorithm on folve's hot path:
 * zita-convolver's `Convproc` (libzita-convolver 4.0.3, "also compatible with
 * 3.1.0": /root/reference/README.md:394, INSTALL.md:55-57) as folve configures
 * it -- quantum == minpart == maxpart == fragm, i.e. ONE level of uniform
 * partitions executed synchronously (/root/reference/zita-fconfig.cc:74-85).
 *
 * zita-convolver is NOT vendored in /root/reference and is NOT installed in
 * this image (no header, no library, no FFTW).  The algorithm below is restated
 * from the published 4.0.3 source as recalled, and anchored on the reference's
 * own call sites:
 *   configure / set_options        zita-fconfig.cc:78-93
 *   impdata_create                 zita-config.cc:163,203,252
 *   impdata_copy (== impdata_link) zita-config.cc:274
 *   inpdata / process / outdata    sound-processor.cc:107,113,117
 *   reset / start_process          sound-processor.cc:140,144
 *   stop_process / cleanup         sound-processor.cc:70-71
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for
 * this path, and the real library cannot be run here.  The
 * restatement is instead pinned against (a) a float64 direct convolution of
 * the same inputs (tests/test_oracle.py) and (b) the reference's OWN
 * sound-processor.cc / zita-config.cc compiled unmodified against this facade
 * (oracle/_ref, see oracle/Makefile).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may link, call or execute anything in oracle/.
 */
#ifndef FOLVE_ZITA_ORACLE_H
#define FOLVE_ZITA_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Convproc class constants (zita-convolver.h 4.0.3). */
enum {
    ZO_MAXINP = 64,
    ZO_MAXOUT = 64,
    ZO_MAXLEV = 8,
    ZO_MINPART = 64,
    ZO_MAXPART = 8192,
    ZO_MAXDIVIS = 16,
    ZO_MINQUANT = 16,
    ZO_MAXQUANT = 8192
};

/* Convproc::_state */
enum { ZO_ST_IDLE = 0, ZO_ST_STOP = 1, ZO_ST_WAIT = 2, ZO_ST_PROC = 3 };

/* Converror codes */
enum { ZO_BAD_STATE = -1, ZO_BAD_PARAM = -2, ZO_MEM_ALLOC = -3 };

typedef struct zo_convproc zo_convproc;

zo_convproc *zo_new(void);
void zo_delete(zo_convproc *p);

/* When non-zero (default 0), start_process() on a convolver that is already
 * processing is accepted and re-zeroes the input/output offsets, i.e. Reset()
 * behaves exactly like a fresh object.  With 0 the recalled library behaviour
 * is kept: start_process() returns BAD_STATE and Convproc::_inpoffs keeps its
 * value, so a pooled processor re-used after an odd number of blocks lags by
 * one block. */
void zo_set_reset_is_fresh(zo_convproc *p, int on);

void zo_set_options(zo_convproc *p, uint32_t options);
int zo_configure(zo_convproc *p, uint32_t ninp, uint32_t nout, uint32_t maxsize,
                 uint32_t quantum, uint32_t minpart, uint32_t maxpart, float density);
int zo_impdata_create(zo_convproc *p, uint32_t inp, uint32_t out, int32_t step,
                      const float *data, int32_t ind0, int32_t ind1);
int zo_impdata_link(zo_convproc *p, uint32_t inp1, uint32_t out1, uint32_t inp2, uint32_t out2);
int zo_reset(zo_convproc *p);
int zo_start_process(zo_convproc *p, int abspri, int policy);
int zo_process(zo_convproc *p);
int zo_stop_process(zo_convproc *p);
int zo_cleanup(zo_convproc *p);
int zo_state(const zo_convproc *p);
float *zo_inpdata(const zo_convproc *p, uint32_t inp);
float *zo_outdata(const zo_convproc *p, uint32_t out);

/* Introspection for per-kernel parity tests (not part of Convproc). */
uint32_t zo_parsize(const zo_convproc *p);
uint32_t zo_npar(const zo_convproc *p);
uint32_t zo_ptind(const zo_convproc *p); /* ring slot the NEXT process() writes */
/* filter spectrum of partition j for (inp,out), (parsize+1) interleaved complex, or NULL */
const float *zo_fftb(const zo_convproc *p, uint32_t inp, uint32_t out, uint32_t j);
/* input spectrum ring slot for input `inp`, or NULL if the input has no node */
const float *zo_ffta(const zo_convproc *p, uint32_t inp, uint32_t slot);

#ifdef __cplusplus
}
#endif
#endif
