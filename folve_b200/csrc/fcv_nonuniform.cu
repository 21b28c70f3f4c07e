// fcv_nonuniform.cu -- non-uniform partitioning: a head of small partitions for latency, a tail
// of large partitions for long impulse responses (north star).
//
// Role in the reference: Convproc::configure(ninp, nout, maxsize, quantum, minpart, maxpart) with
// quantum = minpart < maxpart -- zita-convolver's non-uniform mode.  folve itself always passes
// quantum = minpart = maxpart = fragm (/root/reference/zita-fconfig.cc:74-93), i.e. uniform
// partitions and a block of fragm frames; this file is what a caller with a smaller block would use.
//
// Two levels, built from the engine's own filters and streams (a composition over the C ABI: every
// transform, multiply-accumulate AND the sum of the two levels run in the CUDA kernels -- the tail
// level's output block stays on the device and the head level's inverse transform adds it in
// its epilogue, before PCM conversion and maximum; the host only moves PCM):
//   head : the first `maxpart` taps as maxpart / quantum partitions of `quantum` frames, evaluated for
//          every block of `quantum` frames                            -> latency of one small block
//   tail : the taps from `maxpart` on as partitions of `maxpart` frames, evaluated once per `maxpart`
//          input frames.  Its contribution to output frame n is (x * h_tail)[n - maxpart]: everything
//          it needs lies at least one large block in the past, so the large block that was completed
//          by the previous small block is always in time.
//   y = x * h_head + delay_maxpart(x * h_tail)  ==  x * h, sample for sample (in exact arithmetic).
// Truncation is that of the uniform engine with fragm = maxpart (taps at or beyond
// ceil(size / maxpart) * maxpart are dropped); links and additive impulses behave the same.
#include <cstring>
#include <new>
#include <vector>

#include "fcv_internal.h"

using namespace fcv;

struct fcv_nufilter {
    std::atomic<int> refs{1};
    int ninp = 0, nout = 0;
    unsigned size = 0;
    int quantum = 0, maxpart = 0;
    fcv_filter *head = nullptr, *tail = nullptr;   // tail == nullptr: the whole filter fits the head
    std::vector<char> exists;                       // pair has a MAC node (as in the uniform engine)
    bool committed = false;
};

struct fcv_nustream {
    fcv_nufilter *f = nullptr;
    fcv_stream *a = nullptr, *b = nullptr;
    float *abuf = nullptr, *bbuf = nullptr;
    int pos = 0;                   // frames of the current large block seen so far
    bool have_tail = false;        // the tail stream's device output block holds the tail's share of the current large block
};

static bool pow2_in(unsigned v, unsigned lo, unsigned hi) { return v >= lo && v <= hi && !(v & (v - 1)); }

extern "C" fcv_nufilter *fcv_nufilter_begin(int ninp, int nout, unsigned size, unsigned quantum, unsigned maxpart) {
    if (!pow2_in(quantum, FCV_MINPART, FCV_MAXQUANT) || !pow2_in(maxpart, FCV_MINPART, FCV_MAXQUANT) || quantum > maxpart) {
        fail(FCV_E_PARAM, "quantum %u / maxpart %u: powers of two in [%d, %d] with quantum <= maxpart", quantum, maxpart,
             FCV_MINPART, FCV_MAXQUANT);
        return nullptr;
    }
    fcv_nufilter *f = new (std::nothrow) fcv_nufilter();
    if (!f) { fail(FCV_E_ALLOC, "out of memory"); return nullptr; }
    f->ninp = ninp;
    f->nout = nout;
    f->size = size;
    f->quantum = (int)quantum;
    f->maxpart = (int)maxpart;
    // quantum == maxpart is the uniform engine: one level
    const unsigned head_size = (size < maxpart || quantum == maxpart) ? size : maxpart;
    f->head = fcv_filter_begin(ninp, nout, head_size, quantum);
    if (f->head && size > maxpart && quantum < maxpart) {
        f->tail = fcv_filter_begin(ninp, nout, size - maxpart, maxpart);
        if (!f->tail) { fcv_filter_unref(f->head); f->head = nullptr; }
    }
    if (!f->head) { delete f; return nullptr; }
    f->exists.assign((size_t)ninp * nout, 0);
    return f;
}

// a pair that has a node in one level has it in both (links made later must find it)
static int mark_exists(fcv_nufilter *f, int inp, int out) {
    f->exists[(size_t)inp * f->nout + out] = 1;
    int rc = fcv_filter_add(f->head, inp, out, 1, nullptr, 0, 1);
    if (!rc && f->tail) rc = fcv_filter_add(f->tail, inp, out, 1, nullptr, 0, 1);
    return rc;
}

extern "C" int fcv_nufilter_add(fcv_nufilter *f, int inp, int out, int step, const float *data, int ind0, int ind1) {
    if (!f) return fail(FCV_E_PARAM, "null filter");
    if (f->committed) return fail(FCV_E_STATE, "filter already committed");
    if (inp < 0 || inp >= f->ninp || out < 0 || out >= f->nout) return fail(FCV_E_PARAM, "bad input/output index");
    // range test of the uniform engine with fragm = maxpart (Convlevel::impdata_write)
    const long total = (long)((f->size + f->maxpart - 1) / f->maxpart) * f->maxpart;
    const long n = (long)ind1 - (long)ind0, i0 = -(long)ind0;
    if (i0 >= n || i0 + total <= 0) return 0;
    int rc = mark_exists(f, inp, out);
    if (!rc) rc = fcv_filter_add(f->head, inp, out, step, data, ind0, ind1);
    if (!rc && f->tail) rc = fcv_filter_add(f->tail, inp, out, step, data, ind0 - f->maxpart, ind1 - f->maxpart);
    return rc;
}

extern "C" int fcv_nufilter_link(fcv_nufilter *f, int inp1, int out1, int inp2, int out2) {
    if (!f) return fail(FCV_E_PARAM, "null filter");
    if (f->committed) return fail(FCV_E_STATE, "filter already committed");
    if (inp1 < 0 || inp1 >= f->ninp || out1 < 0 || out1 >= f->nout || inp2 < 0 || inp2 >= f->ninp || out2 < 0 ||
        out2 >= f->nout || (inp1 == inp2 && out1 == out2))
        return fail(FCV_E_PARAM, "bad link");
    if (!f->exists[(size_t)inp1 * f->nout + out1]) return 0;   // no source node: no-op, as in zita
    f->exists[(size_t)inp2 * f->nout + out2] = 1;
    int rc = fcv_filter_link(f->head, inp1, out1, inp2, out2);
    if (!rc && f->tail) rc = fcv_filter_link(f->tail, inp1, out1, inp2, out2);
    return rc;
}

extern "C" int fcv_nufilter_commit(fcv_nufilter *f, int device) {
    if (!f) return fail(FCV_E_PARAM, "null filter");
    if (f->committed) return fail(FCV_E_STATE, "filter already committed");
    int rc = fcv_filter_commit(f->head, device);
    if (!rc && f->tail) rc = fcv_filter_commit(f->tail, device);
    if (!rc) f->committed = true;
    return rc;
}

extern "C" void fcv_nufilter_ref(fcv_nufilter *f) { if (f) f->refs++; }
extern "C" void fcv_nufilter_unref(fcv_nufilter *f) {
    if (!f) return;
    if (--f->refs == 0) {
        fcv_filter_unref(f->head);
        if (f->tail) fcv_filter_unref(f->tail);
        delete f;
    }
}
extern "C" int fcv_nufilter_quantum(const fcv_nufilter *f) { return f ? f->quantum : 0; }
extern "C" int fcv_nufilter_head_partitions(const fcv_nufilter *f) { return f ? fcv_filter_partitions(f->head) : 0; }
extern "C" int fcv_nufilter_tail_partitions(const fcv_nufilter *f) { return f && f->tail ? fcv_filter_partitions(f->tail) : 0; }

extern "C" void fcv_nustream_destroy(fcv_nustream *s) {
    if (!s) return;
    if (s->a) fcv_stream_destroy(s->a);
    if (s->b) fcv_stream_destroy(s->b);
    fcv_nufilter_unref(s->f);
    delete s;
}

extern "C" fcv_nustream *fcv_nustream_create(fcv_nufilter *f) {
    if (!f || !f->committed) { fail(FCV_E_STATE, "filter not committed"); return nullptr; }
    fcv_nustream *s = new (std::nothrow) fcv_nustream();
    if (!s) { fail(FCV_E_ALLOC, "out of memory"); return nullptr; }
    s->f = f;
    fcv_nufilter_ref(f);
    s->a = fcv_stream_create(f->head);
    if (s->a && f->tail) s->b = fcv_stream_create(f->tail);
    if (!s->a || (f->tail && !s->b)) { fcv_nustream_destroy(s); return nullptr; }
    s->abuf = fcv_stream_buffer(s->a);
    if (s->b) s->bbuf = fcv_stream_buffer(s->b);
    return s;
}

// Pinned block of quantum * max(ninp, nout) floats: input frames in, processed frames out.
extern "C" float *fcv_nustream_buffer(fcv_nustream *s) { return s ? s->abuf : nullptr; }

extern "C" int fcv_nustream_reset(fcv_nustream *s) {
    if (!s) return fail(FCV_E_PARAM, "null stream");
    int rc = fcv_stream_reset(s->a);
    if (!rc && s->b) rc = fcv_stream_reset(s->b);
    s->pos = 0;
    s->have_tail = false;
    return rc;
}

// One block of up to `quantum` frames (a short block ends the stream: reset before the next file).
extern "C" int fcv_nustream_process(fcv_nustream *s, int frames_valid, float *max_inout) {
    if (!s) return fail(FCV_E_PARAM, "null stream");
    const fcv_nufilter *f = s->f;
    if (frames_valid < 0 || frames_valid > f->quantum) return fail(FCV_E_PARAM, "frames_valid out of range");
    if (s->b && frames_valid > 0)   // the tail level collects its large block from the same input
        memcpy(s->bbuf + (size_t)s->pos * f->ninp, s->abuf, (size_t)frames_valid * f->ninp * sizeof(float));
    // head level: this block, now -- plus, in the inverse transform's epilogue, the tail level's
    // output for these frames (computed one large block ago, still in the tail stream's device block)
    int rc = stream_set_mix(s->a, s->have_tail ? stream_device_out(s->b) + (size_t)s->pos * f->nout : nullptr);
    if (!rc) rc = fcv_stream_process(s->a, frames_valid, max_inout);
    if (rc) return rc;
    s->pos += frames_valid;
    if (s->b && s->pos == f->maxpart) {   // a large block is complete: the tail level convolves it for the NEXT one
        rc = fcv_stream_process(s->b, f->maxpart, nullptr);
        if (rc) return rc;
        s->have_tail = true;
        s->pos = 0;
    }
    return 0;
}
