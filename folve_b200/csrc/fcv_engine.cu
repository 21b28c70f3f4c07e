// fcv_engine.cu -- host side of the C ABI of the B200 convolution engine (include/folve_b200.h):
// filter spectra resident in HBM, per-stream device state, the launch sequence per step, pinned
// staging, the batched entry points and the coalescer that turns concurrent synchronous
// single-stream calls into one launch sequence.  The kernels live in fcv_k_*.cu.
//
// Replaces the Convproc object behind folve's SoundProcessor
// (/root/reference/sound-processor.cc:34-145) and the impulse-loading calls of
// the zita-config loader (/root/reference/zita-config.cc:163,203,252,274,
// /root/reference/zita-fconfig.cc:78-93).  No CPU fallback: without a usable
// sm_100 device every entry point fails with FCV_E_CUDA.
#include "fcv_internal.h"

#include <sched.h>

#include <cmath>
#include <thread>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>

using namespace fcv;

// ---------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------
static thread_local std::string g_err;
std::atomic<unsigned long long> fcv::g_launches{0};

int fcv::fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

// ---------------------------------------------------------------------------
// per-device context
// ---------------------------------------------------------------------------
struct DeviceCtx {
    std::mutex mu;
    bool checked = false;
    bool ok = false;
    std::string why;
};
static std::mutex g_ctx_mu;
static std::map<int, DeviceCtx *> g_ctx;

static DeviceCtx *get_ctx(int device) {
    std::lock_guard<std::mutex> l(g_ctx_mu);
    auto it = g_ctx.find(device);
    if (it != g_ctx.end()) return it->second;
    DeviceCtx *c = new DeviceCtx();
    g_ctx[device] = c;
    return c;
}

static int check_device(int device) {
    DeviceCtx *c = get_ctx(device);
    std::lock_guard<std::mutex> l(c->mu);
    if (!c->checked) {
        c->checked = true;
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess) {
            c->why = std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e);
        } else if (device < 0 || device >= n) {
            c->why = "no such CUDA device";
        } else {
            cudaDeviceProp p;
            e = cudaGetDeviceProperties(&p, device);
            if (e != cudaSuccess) c->why = cudaGetErrorString(e);
            else if (p.major != 10) {
                char b[128];
                snprintf(b, sizeof(b), "device %d is sm_%d%d; this library is built for sm_100a only", device,
                         p.major, p.minor);
                c->why = b;
            } else c->ok = true;
        }
    }
    if (!c->ok) return fail(FCV_E_CUDA, "device %d unusable: %s", device, c->why.c_str());
    return 0;
}

// fragm = 8192 runs the transforms of fcv_fft13.cuh (their spectrum layout differs from the
// generic kernels', so the choice is process-wide); FCV_GENERIC_FFT=1 keeps the generic ones.
static bool use_f13(int log2n) {
    static const bool generic = [] {
        const char *v = getenv("FCV_GENERIC_FFT");
        return v && *v && *v != '0';
    }();
    return log2n == 13 && !generic;
}

static FcvCombiner *combiner_create(fcv_filter *f);
static void combiner_destroy(FcvCombiner *c);

// ---------------------------------------------------------------------------
// filter
// ---------------------------------------------------------------------------
extern "C" int fcv_abi_version(void) { return FCV_ABI_VERSION; }
extern "C" const char *fcv_last_error(void) { return g_err.c_str(); }
extern "C" unsigned long long fcv_kernel_launches(void) { return g_launches.load(); }

extern "C" int fcv_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return fail(FCV_E_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    int ok = 0;
    for (int d = 0; d < n; d++) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, d) == cudaSuccess && p.major == 10) ok++;
    }
    return ok;
}

extern "C" fcv_filter *fcv_filter_begin(int ninp, int nout, unsigned size, unsigned fragm) {
    if (ninp < 1 || ninp > FCV_MAXINP || nout < 1 || nout > FCV_MAXOUT) {
        fail(FCV_E_PARAM, "inputs/outputs out of range (%d, %d)", ninp, nout);
        return nullptr;
    }
    if (fragm < FCV_MINPART || fragm > FCV_MAXQUANT || (fragm & (fragm - 1))) {
        fail(FCV_E_PARAM, "fragm %u is not a power of two in [%d, %d]", fragm, FCV_MINPART, FCV_MAXQUANT);
        return nullptr;
    }
    if (size > FCV_MAXSIZE) {
        fail(FCV_E_PARAM, "size %u exceeds %u", size, (unsigned)FCV_MAXSIZE);
        return nullptr;
    }
    fcv_filter *f = new (std::nothrow) fcv_filter();
    if (!f) { fail(FCV_E_ALLOC, "out of memory"); return nullptr; }
    f->ninp = ninp;
    f->nout = nout;
    f->size = size;
    f->fragm = (int)fragm;
    f->log2n = 0;
    while ((1u << f->log2n) < fragm) f->log2n++;
    f->npar = (int)((size + fragm - 1) / fragm);
    f->pairs.resize((size_t)ninp * nout);
    return f;
}

extern "C" int fcv_filter_add(fcv_filter *f, int inp, int out, int step, const float *data, int ind0, int ind1) {
    if (!f) return fail(FCV_E_PARAM, "null filter");
    if (f->committed) return fail(FCV_E_STATE, "filter already committed");
    if (inp < 0 || inp >= f->ninp || out < 0 || out >= f->nout) return fail(FCV_E_PARAM, "bad input/output index");
    // Convlevel::impdata_write range test: nothing to do if the data lies outside
    // [0, npar * fragm) -- no MAC node is created in that case either.
    const long n = (long)ind1 - (long)ind0;
    const long total = (long)f->npar * f->fragm;
    const long i0 = -(long)ind0;
    if (i0 >= n || i0 + total <= 0) return 0;
    FcvPair &p = f->pairs[(size_t)inp * f->nout + out];
    p.exists = true;
    if (p.link >= 0) return 0;  // linked pairs ignore new data
    if (!data) return 0;
    try {
        if (p.h.empty()) p.h.assign((size_t)total, 0.0f);
    } catch (...) {
        return fail(FCV_E_ALLOC, "out of memory");
    }
    const float norm = 0.5f / (float)f->fragm;
    const long j0 = i0 < 0 ? 0 : i0;
    const long j1 = (i0 + total > n) ? n : i0 + total;
    for (long j = j0; j < j1; j++) p.h[(size_t)(j - i0)] += norm * data[j * step];
    return 0;
}

extern "C" int fcv_filter_link(fcv_filter *f, int inp1, int out1, int inp2, int out2) {
    if (!f) return fail(FCV_E_PARAM, "null filter");
    if (inp1 < 0 || inp1 >= f->ninp || out1 < 0 || out1 >= f->nout || inp2 < 0 || inp2 >= f->ninp ||
        out2 < 0 || out2 >= f->nout || (inp1 == inp2 && out1 == out2))
        return fail(FCV_E_PARAM, "bad link");
    if (f->committed) return fail(FCV_E_STATE, "filter already committed");
    const int src = inp1 * f->nout + out1, dst = inp2 * f->nout + out2;
    if (!f->pairs[src].exists) return 0;  // Convlevel::impdata_link: no source node, no-op
    FcvPair &d = f->pairs[dst];
    d.exists = true;
    d.h.clear();
    d.h.shrink_to_fit();
    d.link = src;
    return 0;
}

static void filter_free_device(fcv_filter *f) {
    if (f->device >= 0) cudaSetDevice(f->device);
    if (f->dH) cudaFree(f->dH);
    if (f->dsteps) cudaFree(f->dsteps);
    if (f->dgroup_off) cudaFree(f->dgroup_off);
    if (f->dpairs) cudaFree(f->dpairs);
    if (f->dpair_off) cudaFree(f->dpair_off);
    if (f->dtt_rows) cudaFree(f->dtt_rows);
    f->dpairs = nullptr;
    f->dpair_off = nullptr;
    f->dtt_rows = nullptr;
    f->dH = nullptr;
    f->dsteps = nullptr;
    f->dgroup_off = nullptr;
}

extern "C" int fcv_filter_commit(fcv_filter *f, int device) {
    if (!f) return fail(FCV_E_PARAM, "null filter");
    if (f->committed) return fail(FCV_E_STATE, "filter already committed");
    int rc = check_device(device);
    if (rc) return rc;
    CU_TRY(cudaSetDevice(device));
    rc = fft_tables(device, f->log2n, &f->tb);
    if (rc) return rc;
    f->k13 = use_f13(f->log2n);
    if (f->k13) {
        rc = fft13_tables(device, &f->tb13);
        if (rc) return rc;
    }
    f->device = device;
    const int N = f->fragm;

    // Rows: every non-zero partition of every pair that owns data.
    std::vector<float> rows;
    int nrows = 0, last_part = -1;
    for (auto &p : f->pairs) {
        p.row.assign((size_t)(f->npar > 0 ? f->npar : 1), -1);  // at least one slot: ring depth is >= 1
        if (!p.exists || p.link >= 0 || p.h.empty()) continue;
        for (int j = 0; j < f->npar; j++) {
            const float *src = p.h.data() + (size_t)j * N;
            bool nz = false;
            for (int k = 0; k < N; k++)
                if (src[k] != 0.0f) { nz = true; break; }
            if (!nz) continue;
            p.row[j] = nrows++;
            rows.insert(rows.end(), src, src + N);
        }
    }
    // Resolve links (one level, as zita does: a link to a link sees no data).
    for (auto &p : f->pairs) {
        if (p.exists && p.link >= 0) {
            const FcvPair &s = f->pairs[(size_t)p.link];
            if (s.link < 0 && !s.row.empty()) p.row = s.row;
        }
    }
    f->active_pairs = 0;
    for (auto &p : f->pairs) {
        bool any = false;
        for (int j = 0; j < f->npar; j++)
            if (p.row[j] >= 0) { any = true; if (j > last_part) last_part = j; }
        if (any) f->active_pairs++;
    }
    f->nrows = nrows;
    f->ring = last_part + 1 > 0 ? last_part + 1 : 1;

    // MAC step table: outputs in groups of group_no.
    f->group_no = f->nout >= 8 ? 8 : (f->nout > 4 ? 8 : (f->nout > 2 ? 4 : f->nout));
    f->ngroups = (f->nout + f->group_no - 1) / f->group_no;
    f->hsteps.clear();
    f->hgroup_off.assign(1, 0);
    for (int g = 0; g < f->ngroups; g++) {
        for (int i = 0; i < f->ninp; i++)
            for (int j = 0; j < f->ring; j++) {
                MacStep sp;
                sp.inp = i;
                sp.part = j;
                bool any = false;
                for (int o = 0; o < MAC_NO_MAX; o++) {
                    const int oo = g * f->group_no + o;
                    sp.row[o] = (o < f->group_no && oo < f->nout) ? f->pairs[(size_t)i * f->nout + oo].row[j] : -1;
                    any |= sp.row[o] >= 0;
                }
                if (any) f->hsteps.push_back(sp);
            }
        f->hgroup_off.push_back((int)f->hsteps.size());
    }
    f->nsteps = (int)f->hsteps.size();

    // Per-output pair lists: for every output the inputs that feed it and, per
    // partition j < ring, the filter row (or -1).
    std::vector<TTPair> hpairs;
    std::vector<int> hpair_off(1, 0), htt_rows;
    for (int o = 0; o < f->nout; o++) {
        for (int i = 0; i < f->ninp; i++) {
            const FcvPair &p = f->pairs[(size_t)i * f->nout + o];
            bool any = false;
            for (int j = 0; j < f->ring && j < (int)p.row.size(); j++) any |= p.row[j] >= 0;
            if (!any) continue;
            TTPair tp;
            tp.inp = i;
            tp.rowbase = (int)htt_rows.size();
            for (int j = 0; j < f->ring; j++) htt_rows.push_back(j < (int)p.row.size() ? p.row[j] : -1);
            hpairs.push_back(tp);
        }
        hpair_off.push_back((int)hpairs.size());
    }
    CU_TRY(cudaMalloc(&f->dpairs, (hpairs.size() + 1) * sizeof(TTPair)));
    CU_TRY(cudaMalloc(&f->dpair_off, hpair_off.size() * sizeof(int)));
    CU_TRY(cudaMalloc(&f->dtt_rows, (htt_rows.size() + 1) * sizeof(int)));
    if (!hpairs.empty())
        CU_TRY(cudaMemcpy(f->dpairs, hpairs.data(), hpairs.size() * sizeof(TTPair), cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(f->dpair_off, hpair_off.data(), hpair_off.size() * sizeof(int), cudaMemcpyHostToDevice));
    if (!htt_rows.empty())
        CU_TRY(cudaMemcpy(f->dtt_rows, htt_rows.data(), htt_rows.size() * sizeof(int), cudaMemcpyHostToDevice));

    const size_t M = (size_t)N;
    // one extra, all-zero row after the filter rows: what the TMA-staged MAC loads for a
    // partition in which a pair has no data
    CU_TRY(cudaMalloc(&f->dH, (size_t)(nrows + 1) * M * sizeof(float2)));
    CU_TRY(cudaMemset(f->dH + (size_t)nrows * M, 0, M * sizeof(float2)));
    CU_TRY(cudaMalloc(&f->dsteps, (size_t)(f->nsteps > 0 ? f->nsteps : 1) * sizeof(MacStep)));
    CU_TRY(cudaMalloc(&f->dgroup_off, f->hgroup_off.size() * sizeof(int)));
    if (f->nsteps)
        CU_TRY(cudaMemcpy(f->dsteps, f->hsteps.data(), (size_t)f->nsteps * sizeof(MacStep), cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(f->dgroup_off, f->hgroup_off.data(), f->hgroup_off.size() * sizeof(int), cudaMemcpyHostToDevice));
    if (nrows) {
        float *dsrc = nullptr;
        CU_TRY(cudaMalloc(&dsrc, rows.size() * sizeof(float)));
        CU_TRY(cudaMemcpy(dsrc, rows.data(), rows.size() * sizeof(float), cudaMemcpyHostToDevice));
        rc = launch_filter_fft(f, dsrc, f->dH, nrows);
        cudaFree(dsrc);
        if (rc) return rc;
    }
    // time-domain copies are no longer needed
    for (auto &p : f->pairs) { p.h.clear(); p.h.shrink_to_fit(); }
    f->combiner = combiner_create(f);
    f->committed = true;
    return 0;
}

extern "C" void fcv_filter_ref(fcv_filter *f) { if (f) f->refs++; }
extern "C" void fcv_filter_unref(fcv_filter *f) {
    if (!f) return;
    if (--f->refs == 0) {
        filter_free_device(f);
        combiner_destroy(f->combiner);
        delete f;
    }
}
extern "C" int fcv_filter_ninp(const fcv_filter *f) { return f ? f->ninp : 0; }
extern "C" int fcv_filter_nout(const fcv_filter *f) { return f ? f->nout : 0; }
extern "C" int fcv_filter_fragm(const fcv_filter *f) { return f ? f->fragm : 0; }
extern "C" int fcv_filter_partitions(const fcv_filter *f) { return f ? f->npar : 0; }
extern "C" int fcv_filter_ring_depth(const fcv_filter *f) { return f ? f->ring : 0; }
extern "C" int fcv_filter_active_rows(const fcv_filter *f) { return f ? f->nrows : 0; }
extern "C" int fcv_filter_active_pairs(const fcv_filter *f) { return f ? f->active_pairs : 0; }
extern "C" int fcv_filter_device(const fcv_filter *f) { return f ? f->device : -1; }

// packed-permuted device row -> natural order (N+1 interleaved complex)
static void unpermute_row(int log2n, const float2 *row, float *dst) {
    const int q = log2n - 1, Q = 1 << q, M = 2 * Q;
    dst[0] = row[0].x; dst[1] = 0.f;
    dst[2 * M] = row[0].y; dst[2 * M + 1] = 0.f;
    for (int k = 1; k < M; k++) {
        // fragm = 8192: split-parity natural layout (fcv_fft13.cuh); else digit-reversed halves
        const int e = use_f13(log2n) ? ((k & 1) << q) + (k >> 1) : fft_entry_of_bin(log2n, k);
        dst[2 * k] = row[e].x;
        dst[2 * k + 1] = row[e].y;
    }
}

extern "C" int fcv_filter_get_spectrum(fcv_filter *f, int inp, int out, int j, float *dst) {
    if (!f || !f->committed) return fail(FCV_E_STATE, "filter not committed");
    if (inp < 0 || inp >= f->ninp || out < 0 || out >= f->nout || j < 0 || j >= f->npar)
        return fail(FCV_E_PARAM, "bad index");
    const int row = f->pairs[(size_t)inp * f->nout + out].row[j];
    if (row < 0) return 0;
    CU_TRY(cudaSetDevice(f->device));
    std::vector<float2> h((size_t)f->fragm);
    CU_TRY(cudaMemcpy(h.data(), f->dH + (size_t)row * f->fragm, h.size() * sizeof(float2), cudaMemcpyDeviceToHost));
    unpermute_row(f->log2n, h.data(), dst);
    return 1;
}

extern "C" int fcv_filter_get_impulse(fcv_filter *f, int inp, int out, float *dst, int capacity) {
    if (!f || !dst || capacity < 0) return fail(FCV_E_PARAM, "null argument");
    if (f->committed) return fail(FCV_E_STATE, "impulses are dropped at commit");
    if (inp < 0 || inp >= f->ninp || out < 0 || out >= f->nout) return fail(FCV_E_PARAM, "bad index");
    const FcvPair &p = f->pairs[(size_t)inp * f->nout + out];
    if (!p.exists) return 0;
    const FcvPair *src = &p;
    if (p.link >= 0) src = &f->pairs[(size_t)p.link];
    const size_t total = (size_t)f->npar * f->fragm;
    const size_t n = total < (size_t)capacity ? total : (size_t)capacity;
    memset(dst, 0, (size_t)capacity * sizeof(float));
    if (src->link < 0 && !src->h.empty()) memcpy(dst, src->h.data(), n * sizeof(float));
    return p.link >= 0 ? 2 : 1;
}

// ---------------------------------------------------------------------------
// batch
// ---------------------------------------------------------------------------
static size_t pcm_bytes(int fmt) { return fmt == FCV_PCM_S16 ? 2 : 4; }
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Pinned host staging of a batch: anonymous memory on transparent huge pages (2 MB aligned,
// madvise(MADV_HUGEPAGE), faulted in, then pinned with cudaHostRegister) -- 512 x fewer IOMMU /
// page-table entries under the DMA than 4 KB pages.  Measured on the end-to-end loop
// (profiles/r02_experiments.md): +1.5 % at one GPU (three alternating pairs, inside the run-to-run
// spread), +4 % at two, +1 % at eight, where the host fabric saturates: never worse, so it is the
// default.  FCV_HUGEPAGES=0, or any failure on the way: plain cudaHostAlloc.
#include <sys/mman.h>
static bool use_hugepages() {
    static const bool on = !(getenv("FCV_HUGEPAGES") && atoi(getenv("FCV_HUGEPAGES")) == 0);
    return on;
}
namespace {
struct StagingMap { void *base; size_t len; };
std::mutex g_staging_mu;
std::map<void *, StagingMap> g_staging;   // registered pointer -> its mapping
}  // namespace
static const size_t kHuge = 2u << 20;
static cudaError_t staging_alloc(void **p, size_t bytes) {
    if (use_hugepages() && bytes >= kHuge) {
        const size_t len = (bytes + kHuge - 1) / kHuge * kHuge;
        void *base = mmap(nullptr, len + kHuge, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (base != MAP_FAILED) {
            void *m = (void *)(((uintptr_t)base + kHuge - 1) / kHuge * kHuge);
            madvise(m, len, MADV_HUGEPAGE);
            memset(m, 0, len);   // fault the pages in before they are pinned
            if (cudaHostRegister(m, len, cudaHostRegisterDefault) == cudaSuccess) {
                std::lock_guard<std::mutex> l(g_staging_mu);
                g_staging[m] = StagingMap{base, len + kHuge};
                *p = m;
                return cudaSuccess;
            }
            cudaGetLastError();   // clear, fall back
            munmap(base, len + kHuge);
        }
    }
    return cudaHostAlloc(p, bytes, cudaHostAllocDefault);
}
static void staging_free(void *p) {
    if (!p) return;
    StagingMap m{nullptr, 0};
    {
        std::lock_guard<std::mutex> l(g_staging_mu);
        auto it = g_staging.find(p);
        if (it != g_staging.end()) {
            m = it->second;
            g_staging.erase(it);
        }
    }
    if (!m.base) { cudaFreeHost(p); return; }
    cudaHostUnregister(p);
    munmap(m.base, m.len);
}

static void batch_free(fcv_batch *b) {
    if (!b) return;
    if (b->f && b->f->device >= 0) cudaSetDevice(b->f->device);
    for (auto e : b->ev) cudaEventDestroy(e);
    for (int i = 0; i < 16; i++) if (b->sw[i]) cudaEventDestroy(b->sw[i]);
    for (int i = 0; i < 9; i++) if (b->fj[i]) cudaEventDestroy(b->fj[i]);
    for (int i = 0; i < fcv_batch::NQ; i++)
        if (b->q[i]) { cudaStreamSynchronize(b->q[i]); cudaStreamDestroy(b->q[i]); }
    if (b->dmem) cudaFree(b->dmem);
    if (b->hin) { if (b->per_block_max) cudaFreeHost(b->hin); else staging_free(b->hin); }
    if (b->hout) staging_free(b->hout);
    if (b->hfv) cudaFreeHost(b->hfv);
    if (b->hin1) staging_free(b->hin1);
    if (b->hout1) staging_free(b->hout1);
    if (b->hfv1) cudaFreeHost(b->hfv1);
    if (b->dfv1) cudaFree(b->dfv1);
    for (int k = 0; k < 2; k++) if (b->hbmax[k]) cudaFreeHost(b->hbmax[k]);
    for (int k = 0; k < 2; k++)
        for (int i = 0; i < 8; i++)
            if (b->slot_done[k][i]) cudaEventDestroy(b->slot_done[k][i]);
    if (b->f) fcv_filter_unref(b->f);
    delete b;
}

static fcv_batch *batch_create(fcv_filter *f, int nstreams, int in_fmt, int out_fmt, bool shared_host_buffer, int T) {
    if (!f || !f->committed) { fail(FCV_E_STATE, "filter not committed"); return nullptr; }
    if (nstreams < 1) { fail(FCV_E_PARAM, "nstreams < 1"); return nullptr; }
    if (T != 1 && T != 2 && T != 4 && T != 8) { fail(FCV_E_PARAM, "blocks per step must be 1, 2, 4 or 8"); return nullptr; }
    if (in_fmt < 0 || in_fmt > FCV_PCM_S24 || out_fmt < 0 || out_fmt > FCV_PCM_S24) {
        fail(FCV_E_PARAM, "bad PCM format");
        return nullptr;
    }
    if (cudaSetDevice(f->device) != cudaSuccess) { fail(FCV_E_CUDA, "cudaSetDevice failed"); return nullptr; }
    fcv_batch *b = new (std::nothrow) fcv_batch();
    if (!b) { fail(FCV_E_ALLOC, "out of memory"); return nullptr; }
    b->f = f;
    fcv_filter_ref(f);
    b->B = nstreams;
    b->in_fmt = in_fmt;
    b->out_fmt = out_fmt;
    const size_t N = (size_t)f->fragm, B = (size_t)nstreams;
    b->T = T;
    b->R = f->ring + T - 1;
    b->in_block = (size_t)T * N * f->ninp * pcm_bytes(in_fmt);
    b->out_block = (size_t)T * N * f->nout * pcm_bytes(out_fmt);
    b->out_pad = 256;
    b->per_block_max = shared_host_buffer;
    if (cudaDeviceGetAttribute(&b->num_sms, cudaDevAttrMultiProcessorCount, f->device) != cudaSuccess || b->num_sms < 1)
        b->num_sms = 148;

    const size_t xring_b = align_up(B * f->ninp * b->R * N * sizeof(float2), 256);
    const size_t tail_b = align_up(B * f->nout * N * sizeof(float), 256);
    const size_t din_b = align_up(B * b->in_block, 256);
    const size_t dout_b = align_up(B * b->out_block + b->out_pad, 256);
    const size_t y_b = align_up(B * f->nout * T * N * sizeof(float2), 256);
    const size_t max_b = align_up(B * sizeof(float), 256);
    const size_t bmax_b = align_up(B * (size_t)T * sizeof(float), 256);
    const size_t st_b = align_up(B * sizeof(StreamDev), 256);
    const size_t fv_b = align_up(B * sizeof(int), 256);
    const size_t zc_b = align_up(B * f->nout * (size_t)T * sizeof(float2), 256);
    const size_t arr_b = 256;   // single streams: arrival counter of the output-channel CTAs
    const size_t total = xring_b + tail_b + din_b + dout_b + y_b + max_b + bmax_b + st_b + fv_b + zc_b + arr_b;
    cudaError_t e = cudaMalloc(&b->dmem, total);
    if (e != cudaSuccess) {
        fail(FCV_E_ALLOC, "cudaMalloc(%zu bytes) failed: %s", total, cudaGetErrorString(e));
        batch_free(b);
        return nullptr;
    }
    unsigned char *p = b->dmem;
    b->xring = (float2 *)p; p += xring_b;
    b->tail = (float *)p; p += tail_b;
    b->din = p; p += din_b;
    b->dout = p; p += dout_b;
    b->Y = (float2 *)p; p += y_b;
    // single-stream mode keeps the running maximum right behind the output block
    // so that one device->host copy can bring back both
    b->maxv = shared_host_buffer ? (float *)(b->dout + B * b->out_block) : (float *)p;
    p += max_b;
    b->bmax = (float *)p; p += bmax_b;
    b->dst = (StreamDev *)p; p += st_b;
    b->dfv = (int *)p; p += fv_b;
    b->zc0 = (float2 *)p; p += zc_b;
    unsigned *arrive = (unsigned *)p; p += arr_b;
    b->state_bytes_per_stream = (size_t)f->ninp * b->R * N * sizeof(float2);

    bool ok = cudaMemset(b->dmem, 0, total) == cudaSuccess;
    if (shared_host_buffer) {
        // SoundProcessor::buffer_: fragm * max(ninp, nout) samples (sound-processor.cc:62-63), here in
        // the wire formats of the stream; the running maximum is mirrored right behind it
        b->host_block = b->in_block > b->out_block ? b->in_block : b->out_block;
        const size_t bytes = b->host_block + b->out_pad;
        ok = ok && cudaHostAlloc((void **)&b->hin, bytes, cudaHostAllocMapped) == cudaSuccess;
        if (ok) memset(b->hin, 0, bytes);
        // The forward kernel of a single stream reads its block straight from this pinned host
        // buffer (64 KB of coalesced reads over the link): one copy operation and one driver call
        // less per block than staging it in device memory first.  FCV_STREAM_ZEROCOPY=0: stage.
        static const bool zc = !(getenv("FCV_STREAM_ZEROCOPY") && atoi(getenv("FCV_STREAM_ZEROCOPY")) == 0);
        void *dp = nullptr;
        b->in_zero_copy = ok && zc && cudaHostGetDevicePointer(&dp, b->hin, 0) == cudaSuccess && dp;
        if (b->in_zero_copy) b->hin_dev = dp;
    } else {
        ok = ok && staging_alloc((void **)&b->hin, B * b->in_block) == cudaSuccess;
        ok = ok && staging_alloc((void **)&b->hout, B * b->out_block) == cudaSuccess;
        if (ok) { memset(b->hin, 0, B * b->in_block); memset(b->hout, 0, B * b->out_block); }
    }
    std::vector<StreamDev> hs(B);
    for (size_t s = 0; s < B; s++) {
        hs[s].xring = b->xring + s * f->ninp * b->R * N;
        hs[s].tail = b->tail + s * f->nout * N;
        hs[s].din = b->in_zero_copy ? b->hin_dev : (const void *)(b->din + s * b->in_block);
        hs[s].dout = b->dout + s * b->out_block;
        hs[s].maxv = b->maxv + s;
        hs[s].bmax = b->bmax + s * (size_t)T;
        hs[s].Y = b->Y + s * f->nout * T * N;
        hs[s].zc0 = b->zc0 + s * f->nout * T;
        // single streams whose host block the device can address: the inverse kernel writes the
        // output and the maximum there itself (host_copy_out)
        hs[s].hout = b->in_zero_copy ? const_cast<void *>(b->hin_dev) : nullptr;
        hs[s].hmax = b->in_zero_copy ? (float *)((unsigned char *)const_cast<void *>(b->hin_dev) + b->host_block) : nullptr;
        hs[s].hdone = b->in_zero_copy ? (unsigned *)((unsigned char *)const_cast<void *>(b->hin_dev) + b->host_block + 64) : nullptr;
        hs[s].arrive = arrive;
    }
    ok = ok && cudaMemcpy(b->dst, hs.data(), B * sizeof(StreamDev), cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && cudaHostAlloc((void **)&b->hfv, B * sizeof(int), cudaHostAllocDefault) == cudaSuccess;
    const int nq = shared_host_buffer ? 1 : fcv_batch::NQ;
    for (int i = 0; ok && i < nq; i++) ok = cudaStreamCreateWithFlags(&b->q[i], cudaStreamNonBlocking) == cudaSuccess;
    if (!ok) {
        fail(FCV_E_ALLOC, "batch allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        batch_free(b);
        return nullptr;
    }
    return b;
}

// The launch sequence of one step on CUDA stream q (three launches at T = 1, four otherwise).
// diagnostic: FCV_ONLY=1|2|4 (bit mask fwd|mac|inv) launches only those kernels
static int launch_step(const StepArgs &a, cudaStream_t q, cudaEvent_t *ev) {
    static const int only = getenv("FCV_ONLY") ? atoi(getenv("FCV_ONLY")) : 7;
    const fcv_filter *f = a.f;
    if (ev) cudaEventRecord(ev[0], q);
    if (only & 1) {
        if (f->k13) launch_fwd13(a, q);
        else launch_fwd(a, q);
    }
    if (ev) cudaEventRecord(ev[1], q);
    if (only & 2) {
        if (a.T == 1) {
            launch_mac_t1(a, q);
        } else {
            const int newest = (a.bsel.pt + a.T - 1) % a.R;
            if (a.T == 2 || !launch_mac_tma(a, newest, q)) launch_mac_tt(a, newest, q);
        }
    }
    if (ev) cudaEventRecord(ev[2], q);
    if ((only & 4) && a.T > 1) launch_dcny(a, q);   // T == 1: done inside mac_kernel
    if (only & 4) {
        if (f->k13) launch_inv13(a, q);
        else launch_inv(a, q);
    }
    if (ev) cudaEventRecord(ev[3], q);
    g_launches += a.T > 1 ? 4 : 3;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(FCV_E_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
    return 0;
}

// The launches for streams [off, off+cnt) of the batch on CUDA stream q.
// fv_base: per-stream valid-frame counts on the device, or nullptr = the whole step for every stream
static int run_kernels(fcv_batch *b, int off, int cnt, const int *fv_base, cudaStream_t q, cudaEvent_t *ev) {
    fcv_filter *f = b->f;
    StepArgs a;
    a.f = f;
    a.T = b->T;
    a.R = b->R;
    a.in_fmt = b->in_fmt;
    a.out_fmt = b->out_fmt;
    a.num_sms = b->num_sms;
    a.cnt = cnt;
    a.bsel.st = b->dst + off;
    a.bsel.fv = fv_base ? fv_base + off : nullptr;
    a.bsel.fv_all = b->T * f->fragm;
    a.bsel.xring_stride = (size_t)f->ninp * b->R * f->fragm;
    a.bsel.xring0 = b->xring + (size_t)off * a.bsel.xring_stride;
    a.bsel.din_stride = b->in_zero_copy ? 0 : b->in_block;
    a.bsel.din0 = b->in_zero_copy ? (const char *)b->hin_dev : (const char *)b->din + (size_t)off * b->in_block;
    // the step's first block goes to ring slot (step * T) mod R
    a.bsel.pt = (int)((b->step * (unsigned long long)b->T) % (unsigned long long)b->R);
    a.Y = b->Y + (size_t)off * f->nout * b->T * f->fragm;
    a.zc0 = b->zc0 + (size_t)off * f->nout * b->T;
    return launch_step(a, q, ev);
}

static cudaEvent_t *prof_events(fcv_batch *b) {
    if (!b->profiling) return nullptr;
    if (b->ev_used + 4 > b->ev.size()) {
        for (int i = 0; i < 4; i++) {
            cudaEvent_t e;
            if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
            b->ev.push_back(e);
        }
    }
    cudaEvent_t *p = &b->ev[b->ev_used];
    b->ev_used += 4;
    b->prof_steps++;
    return p;
}

extern "C" fcv_batch *fcv_batch_create(fcv_filter *f, int nstreams, int in_format, int out_format) {
    return batch_create(f, nstreams, in_format, out_format, false, 1);
}
extern "C" fcv_batch *fcv_batch_create_tiled(fcv_filter *f, int nstreams, int in_format, int out_format,
                                             int blocks_per_step) {
    return batch_create(f, nstreams, in_format, out_format, false, blocks_per_step);
}
extern "C" int fcv_batch_blocks_per_step(const fcv_batch *b) { return b ? b->T : 0; }
extern "C" void fcv_batch_destroy(fcv_batch *b) { batch_free(b); }
extern "C" int fcv_batch_nstreams(const fcv_batch *b) { return b ? b->B : 0; }
extern "C" void *fcv_batch_host_in(fcv_batch *b) { return b ? b->hin : nullptr; }
extern "C" void *fcv_batch_host_out(fcv_batch *b) { return b ? b->hout : nullptr; }
extern "C" size_t fcv_batch_host_in_bytes(const fcv_batch *b) { return b ? (size_t)b->B * b->in_block : 0; }
extern "C" size_t fcv_batch_host_out_bytes(const fcv_batch *b) { return b ? (size_t)b->B * b->out_block : 0; }
extern "C" void *fcv_batch_device_in(fcv_batch *b) { return b ? b->din : nullptr; }
extern "C" void *fcv_batch_device_out(fcv_batch *b) { return b ? b->dout : nullptr; }
extern "C" void *fcv_batch_cuda_stream(fcv_batch *b) { return b ? (void *)b->q[0] : nullptr; }

static int stage_fv(fcv_batch *b, const int *frames_valid, cudaStream_t q) {
    for (int s = 0; s < b->B; s++) {
        const int v = frames_valid[s];
        if (v < 0 || v > b->T * b->f->fragm) return fail(FCV_E_PARAM, "frames_valid[%d] = %d out of range", s, v);
        b->hfv[s] = v;
    }
    CU_TRY(cudaMemcpyAsync(b->dfv, b->hfv, (size_t)b->B * sizeof(int), cudaMemcpyHostToDevice, q));
    return 0;
}

extern "C" int fcv_batch_sync(fcv_batch *b) {
    if (!b) return fail(FCV_E_PARAM, "null batch");
    CU_TRY(cudaSetDevice(b->f->device));
    for (int i = 0; i < fcv_batch::NQ; i++)
        if (b->q[i]) CU_TRY(cudaStreamSynchronize(b->q[i]));
    return 0;
}

// The submit path spreads chunks of streams over the batch's CUDA streams, the device path runs
// on q[0] (or its own chunking): steps of the two paths are only ordered against each other
// through a full synchronisation, done here whenever the path (or the chunking) changes.
enum { PATH_NONE = 0, PATH_DEVICE = 1, PATH_SUBMIT = 2 };
static int enter_path(fcv_batch *b, int path) {
    if (b->last_path != path) {
        if (b->last_path != PATH_NONE) {
            int rc = fcv_batch_sync(b);
            if (rc) return rc;
        }
        b->last_path = path;
    }
    return 0;
}

extern "C" int fcv_batch_process_device(fcv_batch *b, const int *frames_valid) {
    if (!b) return fail(FCV_E_PARAM, "null batch");
    CU_TRY(cudaSetDevice(b->f->device));
    int rc = enter_path(b, PATH_DEVICE);
    if (rc) return rc;
    if (frames_valid) {
        // the staging array is reused: the previous block must be done with it
        CU_TRY(cudaStreamSynchronize(b->q[0]));
        rc = stage_fv(b, frames_valid, b->q[0]);
        if (rc) return rc;
    }
    // Optional fork/join over the batch's CUDA streams: the FFT kernels (issue and
    // shared-memory bound) of one part of the batch can then overlap the MAC kernel
    // (HBM bound) of another.  q[0] stays the stream everything is ordered on.
    static const int env_chunks = getenv("FCV_DEVICE_CHUNKS") ? atoi(getenv("FCV_DEVICE_CHUNKS")) : 0;
    int nchunk = (b->profiling || env_chunks < 2) ? 1 : env_chunks;
    if (nchunk > b->B / 8) nchunk = b->B / 8 > 0 ? b->B / 8 : 1;
    if (nchunk == 1) {
        rc = run_kernels(b, 0, b->B, frames_valid ? b->dfv : nullptr, b->q[0], prof_events(b));
        if (rc) return rc;
        b->step++;
        return 0;
    }
    for (int i = 0; i < fcv_batch::NQ + 1; i++)
        if (!b->fj[i]) CU_TRY(cudaEventCreateWithFlags(&b->fj[i], cudaEventDisableTiming));
    CU_TRY(cudaEventRecord(b->fj[fcv_batch::NQ], b->q[0]));
    for (int i = 1; i < fcv_batch::NQ; i++) CU_TRY(cudaStreamWaitEvent(b->q[i], b->fj[fcv_batch::NQ], 0));
    const int per = ((b->B + nchunk - 1) / nchunk + 1) & ~1;
    for (int c = 0, off = 0; off < b->B; c++, off += per) {
        const int cnt = (b->B - off) < per ? (b->B - off) : per;
        rc = run_kernels(b, off, cnt, frames_valid ? b->dfv : nullptr, b->q[c % fcv_batch::NQ], nullptr);
        if (rc) return rc;
    }
    for (int i = 1; i < fcv_batch::NQ; i++) {
        CU_TRY(cudaEventRecord(b->fj[i], b->q[i]));
        CU_TRY(cudaStreamWaitEvent(b->q[0], b->fj[i], 0));
    }
    b->step++;
    return 0;
}

static int ensure_slot1(fcv_batch *b) {
    if (b->hin1) return 0;
    const size_t B = (size_t)b->B;
    CU_TRY(staging_alloc((void **)&b->hin1, B * b->in_block));
    CU_TRY(staging_alloc((void **)&b->hout1, B * b->out_block));
    CU_TRY(cudaHostAlloc((void **)&b->hfv1, B * sizeof(int), cudaHostAllocDefault));
    CU_TRY(cudaMalloc((void **)&b->dfv1, B * sizeof(int)));
    memset(b->hin1, 0, B * b->in_block);
    memset(b->hout1, 0, B * b->out_block);
    return 0;
}

extern "C" void *fcv_batch_host_in_slot(fcv_batch *b, int slot) {
    if (!b || !b->hout || slot < 0 || slot > 1) return nullptr;
    if (slot == 1 && (cudaSetDevice(b->f->device) != cudaSuccess || ensure_slot1(b))) return nullptr;
    return slot ? b->hin1 : b->hin;
}
extern "C" void *fcv_batch_host_out_slot(fcv_batch *b, int slot) {
    if (!b || !b->hout || slot < 0 || slot > 1) return nullptr;
    if (slot == 1 && (cudaSetDevice(b->f->device) != cudaSuccess || ensure_slot1(b))) return nullptr;
    return slot ? b->hout1 : b->hout;
}

extern "C" int fcv_batch_wait(fcv_batch *b, int slot) {
    if (!b || slot < 0 || slot > 1) return fail(FCV_E_PARAM, "bad slot");
    if (!b->slot_busy[slot]) return 0;
    CU_TRY(cudaSetDevice(b->f->device));
    for (int i = 0; i < fcv_batch::NQ; i++)
        if (b->slot_done[slot][i]) CU_TRY(cudaEventSynchronize(b->slot_done[slot][i]));
    b->slot_busy[slot] = false;
    return 0;
}

// Chunks of streams a submit is split into (each chunk: copy in, kernels, copy out on one of NQ
// CUDA streams, round robin -- a given stream of the batch always lands on the same CUDA stream).
// Fixed when the batch first needs it: profiling (which wants one chunk, so that the per-kernel
// events bracket whole launches) re-decides it behind a full synchronisation.
static int submit_chunks(const fcv_batch *b) {
    if (b->profiling) return 1;
    static const int env_chunks = getenv("FCV_CHUNKS") ? atoi(getenv("FCV_CHUNKS")) : 0;  // tuning knob
    // 128 streams per chunk for big batches (8 chunks at 1024 streams: measured best, 16 are
    // slower); small batches -- an album library sharded over many GPUs leaves 16 .. 128 chains
    // per device -- still get up to 4 chunks, because a single chunk puts the copy in, the kernels and
    // the copy out of consecutive steps on ONE CUDA stream, where nothing overlaps
    int nchunk = b->B / 128;
    const int small = b->B / 8 < 4 ? b->B / 8 : 4;
    if (nchunk < small) nchunk = small;
    if (env_chunks > 0) nchunk = env_chunks;
    if (nchunk > 16) nchunk = 16;
    if (nchunk > b->B) nchunk = b->B;
    if (nchunk < 1) nchunk = 1;
    return nchunk;
}
static cudaStream_t submit_stream_of(const fcv_batch *b, int stream_index) {
    const int nchunk = submit_chunks(b);
    const int per = (b->B + nchunk - 1) / nchunk;
    return b->q[nchunk == 1 ? 0 : (stream_index / per) % fcv_batch::NQ];
}

// Enqueue one block for every stream from host staging slot `slot`:
// per chunk of streams host->device copy, the three kernels, device->host copy,
// round-robin over NQ CUDA streams.  Chunks of consecutive submits run in order
// on their stream, so the device state needs no double buffering; only the host
// staging has two slots.
extern "C" int fcv_batch_submit(fcv_batch *b, int slot, const int *frames_valid) {
    if (!b || slot < 0 || slot > 1) return fail(FCV_E_PARAM, "bad slot");
    if (!b->hout) return fail(FCV_E_STATE, "batch has no host staging");
    CU_TRY(cudaSetDevice(b->f->device));
    int rc = enter_path(b, PATH_SUBMIT);
    if (rc) return rc;
    if (slot == 1) { rc = ensure_slot1(b); if (rc) return rc; }
    rc = fcv_batch_wait(b, slot);  // the slot's previous submit must have drained
    if (rc) return rc;
    unsigned char *hin = slot ? b->hin1 : b->hin, *hout = slot ? b->hout1 : b->hout;
    int *hfv = slot ? b->hfv1 : b->hfv, *dfv = slot ? b->dfv1 : b->dfv;
    if (frames_valid) {
        for (int s = 0; s < b->B; s++) {
            const int v = frames_valid[s];
            if (v < 0 || v > b->T * b->f->fragm) return fail(FCV_E_PARAM, "frames_valid[%d] = %d out of range", s, v);
            hfv[s] = v;
        }
    }
    const int nchunk = submit_chunks(b);
    if (!b->hbmax[slot]) {
        CU_TRY(cudaHostAlloc((void **)&b->hbmax[slot], (size_t)b->B * b->T * sizeof(float), cudaHostAllocDefault));
        memset(b->hbmax[slot], 0, (size_t)b->B * b->T * sizeof(float));
    }
    // diagnostic: copies without kernels (the link ceiling of the end-to-end loop)
    static const bool env_nokernels = getenv("FCV_COPY_ONLY") != nullptr;
    const bool nokernels = env_nokernels || b->copy_only;
    const int per = (b->B + nchunk - 1) / nchunk;
    for (int c = 0, off = 0; off < b->B; c++, off += per) {
        const int cnt = (b->B - off) < per ? (b->B - off) : per;
        cudaStream_t q = b->q[nchunk == 1 ? 0 : c % fcv_batch::NQ];
        if (frames_valid)
            CU_TRY(cudaMemcpyAsync(dfv + off, hfv + off, (size_t)cnt * sizeof(int), cudaMemcpyHostToDevice, q));
        CU_TRY(cudaMemcpyAsync(b->din + (size_t)off * b->in_block, hin + (size_t)off * b->in_block,
                               (size_t)cnt * b->in_block, cudaMemcpyHostToDevice, q));
        if (!nokernels) rc = run_kernels(b, off, cnt, frames_valid ? dfv : nullptr, q, nchunk == 1 ? prof_events(b) : nullptr);
        if (rc) return rc;
        CU_TRY(cudaMemcpyAsync(hout + (size_t)off * b->out_block, b->dout + (size_t)off * b->out_block,
                               (size_t)cnt * b->out_block, cudaMemcpyDeviceToHost, q));
        CU_TRY(cudaMemcpyAsync(b->hbmax[slot] + (size_t)off * b->T, b->bmax + (size_t)off * b->T,
                               (size_t)cnt * b->T * sizeof(float), cudaMemcpyDeviceToHost, q));
    }
    for (int i = 0; i < fcv_batch::NQ; i++) {
        if (!b->slot_done[slot][i]) CU_TRY(cudaEventCreateWithFlags(&b->slot_done[slot][i], cudaEventDisableTiming));
        CU_TRY(cudaEventRecord(b->slot_done[slot][i], b->q[i]));
    }
    b->slot_busy[slot] = true;
    b->step++;
    return 0;
}

extern "C" int fcv_batch_process(fcv_batch *b, const int *frames_valid) {
    int rc = fcv_batch_submit(b, 0, frames_valid);
    if (rc) return rc;
    return fcv_batch_wait(b, 0);
}

extern "C" int fcv_batch_set_copy_only(fcv_batch *b, int on) {
    if (!b) return fail(FCV_E_PARAM, "null batch");
    int rc = fcv_batch_sync(b);
    if (rc) return rc;
    b->copy_only = on != 0;
    return 0;
}

extern "C" int fcv_batch_reset_slot(fcv_batch *b, int slot) {
    if (!b || slot < 0 || slot >= b->B) return fail(FCV_E_PARAM, "bad slot");
    CU_TRY(cudaSetDevice(b->f->device));
    const fcv_filter *f = b->f;
    const size_t N = (size_t)f->fragm;
    // chunks of a batch run on several CUDA streams: order the reset after everything
    // already enqueued and before anything enqueued later
    int rc = fcv_batch_sync(b);
    if (rc) return rc;
    CU_TRY(cudaMemsetAsync(b->xring + (size_t)slot * f->ninp * b->R * N, 0, b->state_bytes_per_stream, b->q[0]));
    CU_TRY(cudaMemsetAsync(b->tail + (size_t)slot * f->nout * N, 0, (size_t)f->nout * N * sizeof(float), b->q[0]));
    CU_TRY(cudaMemsetAsync(b->maxv + slot, 0, sizeof(float), b->q[0]));
    CU_TRY(cudaStreamSynchronize(b->q[0]));
    return 0;
}

extern "C" int fcv_batch_reset_slot_async(fcv_batch *b, int slot) {
    if (!b || slot < 0 || slot >= b->B) return fail(FCV_E_PARAM, "bad slot");
    CU_TRY(cudaSetDevice(b->f->device));
    int rc = enter_path(b, PATH_SUBMIT);
    if (rc) return rc;
    const fcv_filter *f = b->f;
    const size_t N = (size_t)f->fragm;
    // the CUDA stream every submit processes this stream of the batch on: the reset lands behind
    // the steps already submitted and ahead of the next one
    cudaStream_t q = submit_stream_of(b, slot);
    CU_TRY(cudaMemsetAsync(b->xring + (size_t)slot * f->ninp * b->R * N, 0, b->state_bytes_per_stream, q));
    CU_TRY(cudaMemsetAsync(b->tail + (size_t)slot * f->nout * N, 0, (size_t)f->nout * N * sizeof(float), q));
    CU_TRY(cudaMemsetAsync(b->maxv + slot, 0, sizeof(float), q));
    return 0;
}

extern "C" const float *fcv_batch_host_block_max_slot(fcv_batch *b, int slot) {
    if (!b || slot < 0 || slot > 1) return nullptr;
    return b->hbmax[slot];
}

extern "C" int fcv_batch_get_max(fcv_batch *b, float *max_out) {
    if (!b || !max_out) return fail(FCV_E_PARAM, "null argument");
    CU_TRY(cudaSetDevice(b->f->device));
    int rc = fcv_batch_sync(b);
    if (rc) return rc;
    CU_TRY(cudaMemcpy(max_out, b->maxv, (size_t)b->B * sizeof(float), cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int fcv_batch_get_block_max(fcv_batch *b, float *max_out) {
    if (!b || !max_out) return fail(FCV_E_PARAM, "null argument");
    CU_TRY(cudaSetDevice(b->f->device));
    int rc = fcv_batch_sync(b);
    if (rc) return rc;
    CU_TRY(cudaMemcpy(max_out, b->bmax, (size_t)b->B * b->T * sizeof(float), cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int fcv_batch_event_record(fcv_batch *b, int slot) {
    if (!b || slot < 0 || slot >= 16) return fail(FCV_E_PARAM, "bad event slot");
    CU_TRY(cudaSetDevice(b->f->device));
    if (!b->sw[slot]) CU_TRY(cudaEventCreate(&b->sw[slot]));
    CU_TRY(cudaEventRecord(b->sw[slot], b->q[0]));
    return 0;
}

extern "C" int fcv_batch_event_elapsed_ms(fcv_batch *b, int slot0, int slot1, float *ms) {
    if (!b || !ms || slot0 < 0 || slot0 >= 16 || slot1 < 0 || slot1 >= 16 || !b->sw[slot0] || !b->sw[slot1])
        return fail(FCV_E_PARAM, "bad event slot");
    CU_TRY(cudaSetDevice(b->f->device));
    CU_TRY(cudaEventSynchronize(b->sw[slot1]));
    CU_TRY(cudaEventElapsedTime(ms, b->sw[slot0], b->sw[slot1]));
    return 0;
}

extern "C" int fcv_batch_set_profiling(fcv_batch *b, int on) {
    if (!b) return fail(FCV_E_PARAM, "null batch");
    // profiling changes how a submit is chunked over the CUDA streams: nothing may be in flight
    int rc = fcv_batch_sync(b);
    if (rc) return rc;
    b->profiling = on != 0;
    b->ev_used = 0;
    b->prof_steps = 0;
    return 0;
}

extern "C" int fcv_batch_profile(fcv_batch *b, float ms[3], int *steps) {
    if (!b || !ms) return fail(FCV_E_PARAM, "null argument");
    int rc = fcv_batch_sync(b);
    if (rc) return rc;
    ms[0] = ms[1] = ms[2] = 0.f;
    for (size_t i = 0; i + 3 < b->ev_used; i += 4) {
        for (int k = 0; k < 3; k++) {
            float t = 0.f;
            CU_TRY(cudaEventElapsedTime(&t, b->ev[i + k], b->ev[i + k + 1]));
            ms[k] += t;
        }
    }
    if (steps) *steps = b->prof_steps;
    b->ev_used = 0;
    b->prof_steps = 0;
    return 0;
}

// ---------------------------------------------------------------------------
// single stream == batch of one with a shared in/out host block, driven through the
// filter's coalescer
// ---------------------------------------------------------------------------
struct fcv_stream {
    fcv_batch *b = nullptr;
    // request state: written by the caller (submit / await) and by the filter's dispatcher thread
    enum State { IDLE = 0, QUEUED = 1, LAUNCHED = 2, FAILED = 3 };
    std::atomic<int> state{IDLE};
    int frames_valid = 0;
    int pt = 0;                       // ring slot of the block in flight
    unsigned seq = 0;                 // sequence number of the block in flight (what the GPU will publish)
    cudaStream_t launched_on = nullptr;
    int rc = 0;
    std::string err;
    double t_submit = 0;              // tracing only
    const float *mix = nullptr;       // device frames added to the next block's output (fcv_nonuniform.cu), or null
    // Sequence number of the last block whose request fields the dispatcher has finished reading.  The real
    // ordering "dispatcher reads -> launch -> GPU -> completion word -> caller writes the next request" runs
    // through the device; this release / acquire pair states it in terms the C++ memory model (and
    // ThreadSanitizer) can see.
    std::atomic<unsigned> read_done{0};
};

// Dispatcher ("group commit") of the synchronous per-file path.  folve convolves every open file
// on its own host thread, one block per call (SoundProcessor::Process, sound-processor.cc:98-127);
// one launch sequence per call makes the process launch-rate bound (a launch costs ~8 us of host
// time in the VMs this was measured on, and launches from different threads serialise in the
// driver).  Here
//   * a call only queues its block and then waits for ONE WORD in its own pinned block, which the
//     GPU writes (host_copy_out) after the block's output and maximum -- no CUDA call on the
//     caller's side, neither to launch nor to wait;
//   * one dispatcher thread per (filter, device) does every launch: whatever is queued when it comes
//     round -- up to GROUP_MAX streams of the same wire formats -- travels as one launch group
//     (one cooperative launch where the shape is covered, else forward / MAC / inverse).  Groups
//     form by themselves: while one is being launched the next requests queue up;
//   * results are those of one launch sequence per stream, bit for bit: the kernels are the same
//     and a stream never meets another stream's data (tests/test_coalesce_gpu.py).
// The dispatcher polls its queue while calls keep coming and goes to sleep on a condition variable
// after ~200 us without work.
struct FcvCombiner {
    fcv_filter *f = nullptr;
    std::mutex mu;
    std::condition_variable cv;
    std::vector<fcv_stream *> pending;      // guarded by mu
    std::atomic<int> npending{0};
    bool quit = false, started = false;     // guarded by mu
    int sleeping = 0;                       // dispatcher threads waiting on cv (guarded by mu)
    std::vector<std::thread> ths;           // FCV_DISPATCHERS of them (default 1)
    int nthreads = 1;
    static const int NQ = 8;
    cudaStream_t q[NQ] = {};
    int next_q = 0;
    std::atomic<int> active{0};             // callers between submit and the return of await
    int ncpu = 1;
    int group_max = GROUP_MAX;
    // FCV_COMBINE_TRACE=1: where a block's time goes, printed when the filter is released / at exit
    bool trace = false;
    std::atomic<unsigned long long> tr_groups{0}, tr_streams{0}, tr_fused{0}, tr_fused_failed{0};
    double tr_launch_us = 0, tr_queue_us = 0;           // dispatcher thread only
    std::atomic<unsigned long long> tr_total_ns{0};     // callers
};

static double now_us() {
    timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return 1e6 * (double)t.tv_sec + 1e-3 * (double)t.tv_nsec;
}
static inline void cpu_relax() {
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
}

static void combiner_report(FcvCombiner *c) {
    const unsigned long long g = c->tr_groups.load(), n = c->tr_streams.load();
    if (c->trace && g)
        fprintf(stderr, "fcv dispatcher: %llu blocks in %llu groups (%.2f per group); per group: host launch %.1f us; "
                        "per block: queued %.1f us, submit->done %.1f us; one-launch groups %llu (failed to launch: %llu)\n",
                n, g, (double)n / g, c->tr_launch_us / g, c->tr_queue_us / n, 1e-3 * (double)c->tr_total_ns.load() / n,
                c->tr_fused.load(), c->tr_fused_failed.load());
    c->tr_groups = 0;
    c->tr_streams = 0;
    c->tr_total_ns = 0;
    c->tr_launch_us = c->tr_queue_us = 0;
}

static std::mutex g_trace_mu;
static std::vector<FcvCombiner *> g_traced;
static void combiner_report_all() {
    std::lock_guard<std::mutex> l(g_trace_mu);
    for (FcvCombiner *c : g_traced) combiner_report(c);
}

static FcvCombiner *combiner_create(fcv_filter *f) {
    FcvCombiner *c = new (std::nothrow) FcvCombiner();
    if (!c) return nullptr;
    c->f = f;
    // FCV_COMBINE_MAX: streams per group (default and maximum 32; 1 = one launch group per block)
    if (const char *v = getenv("FCV_COMBINE_MAX")) c->group_max = atoi(v) > 0 && atoi(v) <= GROUP_MAX ? atoi(v) : GROUP_MAX;
    // FCV_DISPATCHERS: launching threads per (filter, device); whichever is free takes what has queued up meanwhile
    if (const char *v = getenv("FCV_DISPATCHERS")) c->nthreads = atoi(v) >= 1 && atoi(v) <= 8 ? atoi(v) : 1;
    {
        cpu_set_t set;
        CPU_ZERO(&set);
        c->ncpu = sched_getaffinity(0, sizeof(set), &set) == 0 ? CPU_COUNT(&set) : 1;
        if (c->ncpu < 1) c->ncpu = 1;
    }
    c->trace = getenv("FCV_COMBINE_TRACE") != nullptr;
    if (c->trace) {
        std::lock_guard<std::mutex> l(g_trace_mu);
        if (g_traced.empty()) atexit(combiner_report_all);
        g_traced.push_back(c);
    }
    return c;  // the dispatcher thread and its CUDA streams are made on first use
}

static void combiner_destroy(FcvCombiner *c) {
    if (!c) return;
    {
        std::unique_lock<std::mutex> lk(c->mu);
        c->quit = true;
        c->cv.notify_all();
    }
    for (std::thread &t : c->ths)
        if (t.joinable()) t.join();
    if (c->trace) {
        combiner_report(c);
        std::lock_guard<std::mutex> l(g_trace_mu);
        for (auto it = g_traced.begin(); it != g_traced.end(); ++it)
            if (*it == c) { g_traced.erase(it); break; }
    }
    for (int i = 0; i < FcvCombiner::NQ; i++)
        if (c->q[i]) { cudaStreamSynchronize(c->q[i]); cudaStreamDestroy(c->q[i]); }
    delete c;
}

// One cooperative launch per group (fcv_k_fused13.cu) instead of three launches: bit-identical
// and measured SLOWER where it was meant to help (16 callers: 20-22 k against 24 k x realtime --
// concurrent cooperative grids overlap badly on the device, and with a single launching thread a
// launch costs ~3 us, not the ~8 us of contended launches), so it is off unless asked for:
// FCV_FUSED=1, or fcv_debug_set_fused(1) from the tests that prove the equivalence.
static std::atomic<bool> g_fused_enabled{getenv("FCV_FUSED") && atoi(getenv("FCV_FUSED")) != 0};
static std::atomic<unsigned long long> g_fused_launches{0};
extern "C" void fcv_debug_set_fused(int on) { g_fused_enabled.store(on != 0); }
extern "C" unsigned long long fcv_debug_fused_launches(void) { return g_fused_launches.load(); }

static bool use_pdl() {
    static const bool on = !(getenv("FCV_PDL") && atoi(getenv("FCV_PDL")) == 0);
    return on;
}

static unsigned *stream_done_word(const fcv_batch *b) {   // host address of the completion word
    return reinterpret_cast<unsigned *>(b->hin + b->host_block + 64);
}

// Enqueue one group on CUDA stream q: (staged input copies,) one cooperative launch or three
// launches, (staged output copies).  Dispatcher thread only.
static int launch_group(FcvCombiner *c, fcv_stream *const *m, int n, cudaStream_t q) {
    const fcv_filter *f = c->f;
    const fcv_batch *b0 = m[0]->b;
    GroupSel sel;
    bool staged[GROUP_MAX];
    for (int i = 0; i < n; i++) {
        fcv_stream *s = m[i];
        fcv_batch *b = s->b;
        staged[i] = !b->in_zero_copy;
        sel.st[i] = b->dst;
        sel.fv[i] = s->frames_valid;
        sel.pt[i] = s->pt;
        sel.sq[i] = s->seq;
        sel.mx[i] = s->mix;
        if (s->frames_valid > 0 && !b->in_zero_copy)
            CU_TRY(cudaMemcpyAsync(b->din, b->hin, (size_t)s->frames_valid * f->ninp * pcm_bytes(b->in_fmt),
                                   cudaMemcpyHostToDevice, q));
    }
    for (int i = 0; i < n; i++)   // zero-copy streams are not touched after this (staged ones: see below)
        if (!staged[i]) m[i]->read_done.store(sel.sq[i], std::memory_order_release);
    StepArgs a;
    a.f = f;
    a.T = 1;
    a.R = b0->R;
    a.in_fmt = b0->in_fmt;
    a.out_fmt = b0->out_fmt;
    a.pdl = use_pdl();
    a.per_block_max = true;
    a.num_sms = b0->num_sms;
    a.cnt = n;
    a.grp = &sel;
    // forward / MAC / inverse as three launches chained by programmatic dependent launch -- or,
    // when asked for and the shape is covered (fragm 8192, stereo), one cooperative launch
    bool fused = false;
    if (g_fused_enabled.load(std::memory_order_relaxed) && fused13_available(f, a.in_fmt, a.out_fmt)) {
        fused = launch_fused13(a, q);
        if (fused) { c->tr_fused++; g_fused_launches++; }
        else c->tr_fused_failed++;
    }
    if (!fused) {
        int rc = launch_step(a, q, nullptr);
        if (rc) return rc;
    }
    // From here on a zero-copy stream must not be touched any more: its block may complete -- and its
    // caller return, resubmit or destroy the stream -- at any moment.
    for (int i = 0; i < n; i++) {
        if (!staged[i]) continue;   // the inverse kernel writes output, maximum and completion word itself
        fcv_stream *s = m[i];
        fcv_batch *b = s->b;
        // Staged variant (FCV_STREAM_ZEROCOPY=0): the first frames_valid output frames go back into
        // the block (sound-processor.cc:116-125), then the maximum, then -- in stream order -- the
        // completion word (until that copy has run the caller is still waiting: safe to touch).
        const size_t out_bytes = (size_t)s->frames_valid * f->nout * pcm_bytes(b->out_fmt);
        if (out_bytes) CU_TRY(cudaMemcpyAsync(b->hin, b->dout, out_bytes, cudaMemcpyDeviceToHost, q));
        CU_TRY(cudaMemcpyAsync(b->hin + b->host_block, b->maxv, sizeof(float), cudaMemcpyDeviceToHost, q));
        b->seq_src = s->seq;
        s->read_done.store(s->seq, std::memory_order_release);   // last touch of the stream
        CU_TRY(cudaMemcpyAsync(stream_done_word(b), &b->seq_src, sizeof(unsigned), cudaMemcpyHostToHost, q));
    }
    return 0;
}

static void dispatcher_main(FcvCombiner *c) {
    cudaSetDevice(c->f->device);
    std::vector<fcv_stream *> take, group;
    int idle_polls = 0;
    for (;;) {
        // poll the queue while calls keep coming; sleep after a while without work
        if (c->npending.load(std::memory_order_acquire) == 0) {
            if (++idle_polls < 4000) {
                cpu_relax();
                if ((idle_polls & 63) == 0) {
                    std::unique_lock<std::mutex> lk(c->mu);
                    if (c->quit) return;
                }
                continue;
            }
            std::unique_lock<std::mutex> lk(c->mu);
            if (c->quit && c->pending.empty()) return;
            if (c->pending.empty()) {
                c->sleeping++;
                c->cv.wait(lk, [c] { return c->quit || !c->pending.empty(); });
                c->sleeping--;
                if (c->quit && c->pending.empty()) return;
            }
        }
        idle_polls = 0;
        {
            std::unique_lock<std::mutex> lk(c->mu);
            if (c->nthreads == 1 || (int)c->pending.size() <= c->group_max) {
                take.swap(c->pending);
                c->npending.store(0, std::memory_order_release);
            } else {   // several dispatchers: one group's worth, the rest is for the others
                take.assign(c->pending.begin(), c->pending.begin() + c->group_max);
                c->pending.erase(c->pending.begin(), c->pending.begin() + c->group_max);
                c->npending.store((int)c->pending.size(), std::memory_order_release);
            }
        }
        if (take.empty()) continue;
        const double t_take = c->trace ? now_us() : 0;
        // everything queued, oldest first, in groups of equal wire formats
        while (!take.empty()) {
            const fcv_batch *b0 = take.front()->b;
            group.clear();
            for (auto it = take.begin(); it != take.end() && (int)group.size() < c->group_max;) {
                if ((*it)->b->in_fmt == b0->in_fmt && (*it)->b->out_fmt == b0->out_fmt) {
                    group.push_back(*it);
                    it = take.erase(it);
                } else {
                    ++it;
                }
            }
            cudaStream_t q = nullptr;
            {
                std::unique_lock<std::mutex> lk(c->mu);
                if (!c->q[c->next_q]) cudaStreamCreateWithFlags(&c->q[c->next_q], cudaStreamNonBlocking);
                q = c->q[c->next_q];
                if (q) c->next_q = (c->next_q + 1) % FcvCombiner::NQ;
            }
            if (!q) {
                cudaGetLastError();
                for (fcv_stream *s : group) {
                    s->rc = FCV_E_CUDA;
                    s->err = "cannot create a CUDA stream for the launch group";
                    s->state.store(fcv_stream::FAILED, std::memory_order_release);
                }
                continue;
            }
            // Everything that touches the streams happens BEFORE the launch: once the kernels are
            // enqueued a block may complete and its caller return, resubmit or destroy the stream while
            // this thread is still on its way out of the launch call.
            double queued_us = 0;
            for (fcv_stream *s : group) {
                if (c->trace) queued_us += t_take - s->t_submit;
                s->launched_on = q;
                s->state.store(fcv_stream::LAUNCHED, std::memory_order_release);
            }
            const size_t gn = group.size();
            const double h0 = c->trace ? now_us() : 0;
            const int rc = launch_group(c, group.data(), (int)gn, q);
            if (c->trace) {
                std::unique_lock<std::mutex> lk(c->mu);
                c->tr_launch_us += now_us() - h0;
                c->tr_groups++;
                c->tr_streams += gn;
                c->tr_queue_us += queued_us;
            }
            if (rc) {   // nothing was (completely) enqueued: the callers are still waiting, tell them
                const std::string err = fcv_last_error();
                for (fcv_stream *s : group) {
                    s->rc = rc;
                    s->err = err;
                    s->state.store(fcv_stream::FAILED, std::memory_order_release);
                }
            }
        }
    }
}

extern "C" fcv_stream *fcv_stream_create_fmt(fcv_filter *f, int in_format, int out_format) {
    fcv_batch *b = batch_create(f, 1, in_format, out_format, true, 1);
    if (!b) return nullptr;
    if (!f->combiner) { batch_free(b); fail(FCV_E_ALLOC, "out of memory"); return nullptr; }
    fcv_stream *s = new (std::nothrow) fcv_stream();
    if (!s) { batch_free(b); fail(FCV_E_ALLOC, "out of memory"); return nullptr; }
    s->b = b;
    return s;
}

extern "C" fcv_stream *fcv_stream_create(fcv_filter *f) { return fcv_stream_create_fmt(f, FCV_PCM_F32, FCV_PCM_F32); }

extern "C" int fcv_stream_submit(fcv_stream *s, int frames_valid) {
    if (!s) return fail(FCV_E_PARAM, "null stream");
    fcv_batch *b = s->b;
    const fcv_filter *f = b->f;
    if (frames_valid < 0 || frames_valid > f->fragm) return fail(FCV_E_PARAM, "frames_valid out of range");
    if (s->state.load(std::memory_order_acquire) != fcv_stream::IDLE)
        return fail(FCV_E_STATE, "stream has a block in flight: call fcv_stream_await first");
    FcvCombiner *c = f->combiner;
    // the unread rest of the input block is silence (sound-processor.cc:99-103)
    const size_t in_bytes = (size_t)frames_valid * f->ninp * pcm_bytes(b->in_fmt);
    if (in_bytes < b->in_block) memset(b->hin + in_bytes, 0, b->in_block - in_bytes);
    s->frames_valid = frames_valid;
    s->rc = 0;
    // ring slot and sequence number are the caller's (one block per stream in flight): the
    // sequence number is never the value the completion word holds now
    s->pt = (int)(b->step % (unsigned long long)b->R);
    b->step++;
    s->seq = (unsigned)b->step;
    if (c->trace) s->t_submit = now_us();
    if (c->active.fetch_add(1, std::memory_order_acq_rel) == 0 && c->group_max > 0) {
        // Nobody else has a block in flight: this caller launches its own group of one, right here,
        // on the stream's own CUDA stream -- no hand-over to the dispatcher thread (3 us less on the
        // lone-stream block latency).
        if (cudaSetDevice(f->device) != cudaSuccess) {
            c->active.fetch_sub(1, std::memory_order_relaxed);
            return fail(FCV_E_CUDA, "cudaSetDevice failed");
        }
        const double h0 = c->trace ? now_us() : 0;
        const int rc = launch_group(c, &s, 1, b->q[0]);
        if (rc) {
            c->active.fetch_sub(1, std::memory_order_relaxed);
            return rc;
        }
        if (c->trace) {
            std::unique_lock<std::mutex> lk(c->mu);
            c->tr_launch_us += now_us() - h0;
            c->tr_groups++;
            c->tr_streams++;
        }
        s->launched_on = b->q[0];
        s->state.store(fcv_stream::LAUNCHED, std::memory_order_release);
        return 0;
    }
    s->state.store(fcv_stream::QUEUED, std::memory_order_release);
    {
        std::unique_lock<std::mutex> lk(c->mu);
        if (!c->started) {
            c->started = true;
            try {
                for (int i = 0; i < c->nthreads; i++) c->ths.emplace_back(dispatcher_main, c);
            } catch (...) {
                if (!c->ths.empty()) goto started_some;   // fewer dispatchers than asked for will do
                c->started = false;
                s->state.store(fcv_stream::IDLE, std::memory_order_release);
                c->active.fetch_sub(1, std::memory_order_relaxed);
                return fail(FCV_E_ALLOC, "cannot start the dispatcher thread");
            }
        started_some:;
        }
        c->pending.push_back(s);
        c->npending.fetch_add(1, std::memory_order_release);
        if (c->sleeping > 0) c->cv.notify_one();
    }
    return 0;
}

extern "C" int fcv_stream_await(fcv_stream *s, float *max_inout) {
    if (!s) return fail(FCV_E_PARAM, "null stream");
    fcv_batch *b = s->b;
    FcvCombiner *c = b->f->combiner;
    if (s->state.load(std::memory_order_acquire) == fcv_stream::IDLE) return fail(FCV_E_STATE, "no block was submitted");
    volatile unsigned *done = stream_done_word(b);
    // Wait for the GPU to publish this block's sequence number in the pinned block.  With more
    // callers than CPUs the wait yields its time slice, otherwise it just polls the cache line.
    int rc = 0;
    const bool crowded = c->active.load(std::memory_order_relaxed) + 1 > c->ncpu;   // + the dispatcher
    for (unsigned long spins = 1;; spins++) {
        if (*done == s->seq) break;
        const int st = s->state.load(std::memory_order_acquire);
        if (st == fcv_stream::FAILED) {
            rc = s->rc ? s->rc : FCV_E_CUDA;
            break;
        }
        if (crowded) sched_yield();
        else cpu_relax();
        if ((spins & 0xfffff) == 0 && st == fcv_stream::LAUNCHED) {
            // a long wait: has the launch itself failed on the device?
            cudaError_t e = cudaSetDevice(b->f->device);
            if (e == cudaSuccess) e = cudaStreamQuery(s->launched_on);
            if (e != cudaSuccess && e != cudaErrorNotReady) {
                s->err = std::string("convolution failed: ") + cudaGetErrorString(e);
                rc = FCV_E_CUDA;
                break;
            }
        }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    // the dispatcher is done with this request's fields (always true by now: it read them before it launched)
    if (!rc)
        while (s->read_done.load(std::memory_order_acquire) != s->seq) cpu_relax();
    if (c->trace) c->tr_total_ns += (unsigned long long)(1e3 * (now_us() - s->t_submit));
    s->state.store(fcv_stream::IDLE, std::memory_order_release);
    c->active.fetch_sub(1, std::memory_order_relaxed);
    if (rc) return fail(rc, "%s", s->err.c_str());
    if (max_inout) {
        float m;
        memcpy(&m, b->hin + b->host_block, sizeof(float));
        if (m > *max_inout) *max_inout = m;
    }
    return 0;
}

extern "C" int fcv_stream_process(fcv_stream *s, int frames_valid, float *max_inout) {
    int rc = fcv_stream_submit(s, frames_valid);
    if (rc) return rc;
    return fcv_stream_await(s, max_inout);
}

static int stream_idle(fcv_stream *s) { return s->state.load(std::memory_order_acquire) == fcv_stream::IDLE; }

extern "C" void fcv_stream_destroy(fcv_stream *s) {
    if (!s) return;
    if (!stream_idle(s)) fcv_stream_await(s, nullptr);
    // the block's last kernel may still be finishing OTHER streams of its group; cudaFree below waits for it
    batch_free(s->b);
    delete s;
}

extern "C" int fcv_stream_reset(fcv_stream *s) {
    if (!s) return fail(FCV_E_PARAM, "null stream");
    if (!stream_idle(s)) return fail(FCV_E_STATE, "stream has a block in flight: call fcv_stream_await first");
    return fcv_batch_reset_slot(s->b, 0);
}

// internal (fcv_nonuniform.cu): device frames [fragm][nout] float added to the output of the blocks
// submitted from now on (any-size kernels only: fragm < 8192), and the stream's device output block
int fcv::stream_set_mix(fcv_stream *s, const float *device_frames) {
    if (device_frames && s->b->f->k13) return fail(FCV_E_STATE, "mixing is not available for fragm = 8192 streams");
    s->mix = device_frames;
    return 0;
}
const float *fcv::stream_device_out(const fcv_stream *s) { return reinterpret_cast<const float *>(s->b->dout); }

extern "C" float *fcv_stream_buffer(fcv_stream *s) { return s ? (float *)s->b->hin : nullptr; }
extern "C" size_t fcv_stream_buffer_bytes(const fcv_stream *s) { return s ? s->b->host_block : 0; }
extern "C" fcv_filter *fcv_stream_filter(fcv_stream *s) { return s ? s->b->f : nullptr; }

extern "C" int fcv_stream_get_input_spectrum(fcv_stream *s, int inp, int age, float *dst) {
    if (!s || !dst) return fail(FCV_E_PARAM, "null argument");
    fcv_batch *b = s->b;
    const fcv_filter *f = b->f;
    if (inp < 0 || inp >= f->ninp || age < 0 || age >= f->ring || (unsigned long long)age >= b->step)
        return fail(FCV_E_PARAM, "bad index");
    CU_TRY(cudaSetDevice(f->device));
    const int slot = (int)((b->step - 1 - age) % (unsigned long long)b->R);  // single streams have T == 1
    std::vector<float2> h((size_t)f->fragm);
    CU_TRY(cudaMemcpy(h.data(), b->xring + (size_t)(inp * b->R + slot) * f->fragm, h.size() * sizeof(float2),
                      cudaMemcpyDeviceToHost));
    unpermute_row(f->log2n, h.data(), dst);
    return 0;
}
