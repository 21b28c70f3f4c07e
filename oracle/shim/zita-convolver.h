// oracle/shim/zita-convolver.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A `Convproc`-shaped C++ facade over the C restatement in ../zita_oracle.c, so
// that the reference's OWN hot-path sources (sound-processor.cc, zita-config.cc,
// zita-fconfig.cc, processor-pool.cc ...) can be compiled unmodified from
// /root/reference although libzita-convolver is not installed here.  Only the
// members the reference calls exist (call sites listed in ../zita_oracle.h).
// Class constants are those of zita-convolver 4.0.3.
#ifndef FOLVE_ORACLE_ZITA_CONVOLVER_SHIM_H
#define FOLVE_ORACLE_ZITA_CONVOLVER_SHIM_H

#include <stdint.h>

#include "../zita_oracle.h"

#define ZITA_CONVOLVER_MAJOR_VERSION 4
#define ZITA_CONVOLVER_MINOR_VERSION 0

// Optional observer of impdata_create / impdata_link calls (used by the parity
// tests to compare what the reference's parser feeds the convolver with what
// the product's loader feeds the CUDA engine).
extern "C" {
typedef void (*zo_shim_impdata_hook)(void *user, unsigned inp, unsigned out, int step, const float *data,
                                     int ind0, int ind1);
typedef void (*zo_shim_link_hook)(void *user, unsigned inp1, unsigned out1, unsigned inp2, unsigned out2);
extern zo_shim_impdata_hook zo_shim_on_impdata;
extern zo_shim_link_hook zo_shim_on_link;
extern void *zo_shim_hook_user;
extern int zo_shim_reset_is_fresh;  // applied to every Convproc created afterwards
}

class Converror {
public:
    enum { BAD_STATE = -1, BAD_PARAM = -2, MEM_ALLOC = -3 };
};

class Convproc {
public:
    enum { ST_IDLE, ST_STOP, ST_WAIT, ST_PROC };
    enum { FL_LATE = 0x0000FFFF, FL_LOAD = 0x01000000 };
    enum { OPT_FFTW_MEASURE = 1, OPT_VECTOR_MODE = 2, OPT_LATE_CONTIN = 4 };
    enum {
        MAXINP = ZO_MAXINP, MAXOUT = ZO_MAXOUT, MAXLEV = ZO_MAXLEV, MINPART = ZO_MINPART,
        MAXPART = ZO_MAXPART, MAXDIVIS = ZO_MAXDIVIS, MINQUANT = ZO_MINQUANT, MAXQUANT = ZO_MAXQUANT
    };

    Convproc() : p_(zo_new()) { zo_set_reset_is_fresh(p_, zo_shim_reset_is_fresh); }
    ~Convproc() { zo_delete(p_); }

    uint32_t state() const { return (uint32_t)zo_state(p_); }
    float *inpdata(uint32_t inp) const { return zo_inpdata(p_, inp); }
    float *outdata(uint32_t out) const { return zo_outdata(p_, out); }
    void set_options(uint32_t options) { zo_set_options(p_, options); }

    int configure(uint32_t ninp, uint32_t nout, uint32_t maxsize, uint32_t quantum, uint32_t minpart,
                  uint32_t maxpart, float density) {
        return zo_configure(p_, ninp, nout, maxsize, quantum, minpart, maxpart, density);
    }
    int impdata_create(uint32_t inp, uint32_t out, int32_t step, float *data, int32_t ind0, int32_t ind1) {
        if (zo_shim_on_impdata) zo_shim_on_impdata(zo_shim_hook_user, inp, out, step, data, ind0, ind1);
        return zo_impdata_create(p_, inp, out, step, data, ind0, ind1);
    }
    int impdata_link(uint32_t inp1, uint32_t out1, uint32_t inp2, uint32_t out2) {
        if (zo_shim_on_link) zo_shim_on_link(zo_shim_hook_user, inp1, out1, inp2, out2);
        return zo_impdata_link(p_, inp1, out1, inp2, out2);
    }
    // deprecated alias kept by zita-convolver 4 for the 3.x name folve uses
    int impdata_copy(uint32_t inp1, uint32_t out1, uint32_t inp2, uint32_t out2) {
        return impdata_link(inp1, out1, inp2, out2);
    }
    int reset() { return zo_reset(p_); }
    int start_process(int abspri, int policy) { return zo_start_process(p_, abspri, policy); }
    int process(bool sync = false) { (void)sync; return zo_process(p_); }
    int stop_process() { return zo_stop_process(p_); }
    int cleanup() { return zo_cleanup(p_); }

    zo_convproc *oracle_handle() const { return p_; }

private:
    Convproc(const Convproc &);
    Convproc &operator=(const Convproc &);
    zo_convproc *p_;
};

#endif
