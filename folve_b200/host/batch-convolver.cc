// batch-convolver.cc -- see batch-convolver.h.
#include "batch-convolver.h"

#include <string.h>
#include <syslog.h>

#include <atomic>
#include <functional>
#include <thread>

#include "../../include/folve_b200.h"
#include "filter-config.h"

namespace folve_b200 {

// One chain in flight.  The block under construction may contain the tail of
// file `k` and the head of file `k + 1`; `share[]` says how the processed block
// is split between them (SoundProcessor::WriteProcessed(out, r) and the
// successor's pending_writes(), convolve-file-handler.cc:373-376,408).
struct BatchConvolver::Slot {
    Chain *chain = nullptr;
    size_t k = 0;          // file being read
    long left = 0;         // frames of file k not yet read
    int fill = 0;          // frames in the block being assembled (frames_valid)
    struct { size_t file; int frames; } share[2];
    int nshare = 0;
    bool reset_before_next = false;  // next block starts from the fresh state
    size_t finished[2];    // files completed by this block (receive the running max)
    int nfinished = 0;
    bool active() const { return chain != nullptr; }
};

BatchConvolver *BatchConvolver::Create(const std::string &config_file, int samplerate, int channels, int slots,
                                       bool gapless, int device) {
    FilterConfig cfg;
    cfg.fsamp = samplerate;
    cfg.ninp = channels;
    cfg.nout = channels;
    if (LoadFilterConfig(&cfg, config_file.c_str()) != 0 || !cfg.filter) return nullptr;
    if (fcv_filter_commit(cfg.filter, device) != 0) {
        syslog(LOG_ERR, "folve-b200: %s: %s", config_file.c_str(), fcv_last_error());
        fcv_filter_unref(cfg.filter);
        return nullptr;
    }
    fcv_batch *batch = fcv_batch_create(cfg.filter, slots, FCV_PCM_F32, FCV_PCM_F32);
    if (!batch) {
        syslog(LOG_ERR, "folve-b200: %s: %s", config_file.c_str(), fcv_last_error());
        fcv_filter_unref(cfg.filter);
        return nullptr;
    }
    BatchConvolver *bc = new BatchConvolver();
    bc->filter_ = cfg.filter;
    bc->batch_ = batch;
    bc->fragm_ = cfg.fragm;
    bc->ninp_ = cfg.ninp;
    bc->nout_ = cfg.nout;
    bc->slots_ = slots;
    bc->gapless_ = gapless;
    return bc;
}

BatchConvolver::~BatchConvolver() {
    if (batch_) fcv_batch_destroy(batch_);
    if (filter_) fcv_filter_unref(filter_);
}

// Assemble the next block of a chain: FillBuffer on file k, and -- if that file
// ends inside the block and gapless joining is on -- one top-up from file k+1
// (PassoverProcessor, convolve-file-handler.cc:345-348).
void BatchConvolver::FillSlot(Slot &s, float *in_block) {
    Chain &c = *s.chain;
    s.fill = 0;
    s.nshare = 0;
    s.nfinished = 0;
    s.reset_before_next = false;
    // skip empty files: AddMoreSoundData returns false at once for them
    while (s.k < c.size() && s.left == 0) {
        if (++s.k < c.size()) s.left = c[s.k].frames;
    }
    if (s.k >= c.size()) return;
    ChainFile &a = c[s.k];
    int r = (int)(s.left < fragm_ ? s.left : fragm_);
    r = (int)sf_readf_float(a.in, in_block, r);
    if (r == 0) {  // premature EOF: the file is over, nothing is written for it
        s.left = 0;
        s.reset_before_next = true;
        return;
    }
    s.left -= r;
    s.fill = r;
    s.share[s.nshare].file = s.k;
    s.share[s.nshare++].frames = r;
    if (s.left > 0) return;  // a full block from the middle of the file
    // file k ends with this block
    s.finished[s.nfinished++] = s.k;
    if (s.fill == fragm_ || !gapless_ || s.k + 1 >= c.size()) {
        // block complete (no hand-off, quirk 3), or nobody to hand over to
        s.k++;
        s.left = s.k < c.size() ? c[s.k].frames : 0;
        s.reset_before_next = true;
        return;
    }
    // hand the half-filled block over to the alphabetically next file
    ChainFile &b = c[s.k + 1];
    int r2 = (int)((long)(fragm_ - s.fill) < b.frames ? (fragm_ - s.fill) : b.frames);
    r2 = (int)sf_readf_float(b.in, in_block + (size_t)s.fill * ninp_, r2);
    a.out_gapless = true;
    b.in_gapless = true;
    s.fill += r2;
    s.k++;
    s.left = b.frames - r2;
    if (s.left == 0) {
        // the successor was swallowed by the top-up (quirk 4): it writes nothing,
        // and whoever comes next starts fresh
        s.finished[s.nfinished++] = s.k;
        s.k++;
        s.left = s.k < c.size() ? c[s.k].frames : 0;
        s.reset_before_next = true;
    } else {
        // the successor owns the rest of the block (its pending_writes())
        s.share[s.nshare].file = s.k;
        s.share[s.nshare++].frames = fragm_ - r;
    }
}

void BatchConvolver::DrainSlot(Slot &s, const float *out_block, float running_max) {
    Chain &c = *s.chain;
    int pos = 0;
    for (int i = 0; i < s.nshare; i++) {
        ChainFile &f = c[s.share[i].file];
        sf_writef_float(f.out, out_block + (size_t)pos * nout_, s.share[i].frames);
        f.written += s.share[i].frames;
        pos += s.share[i].frames;
    }
    for (int i = 0; i < s.nfinished; i++) c[s.finished[i]].max_value = running_max;
}

bool BatchConvolver::Run(const std::vector<Chain *> &chains, int threads) {
    if (threads < 1) threads = 1;
    std::vector<Slot> slots((size_t)slots_);
    std::vector<int> fv((size_t)slots_, 0);
    std::vector<float> maxv((size_t)slots_, 0.0f);
    float *hin = (float *)fcv_batch_host_in(batch_);
    const float *hout = (const float *)fcv_batch_host_out(batch_);
    const size_t in_stride = (size_t)fragm_ * ninp_, out_stride = (size_t)fragm_ * nout_;
    size_t next_chain = 0;
    bool ok = true;

    auto parallel = [&](const std::function<void(int)> &fn) {
        if (threads == 1) {
            for (int i = 0; i < slots_; i++) fn(i);
            return;
        }
        std::atomic<int> next(0);
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; t++)
            pool.emplace_back([&] {
                for (int i = next.fetch_add(1); i < slots_; i = next.fetch_add(1)) fn(i);
            });
        for (auto &th : pool) th.join();
    };

    for (;;) {
        // hand idle slots a new chain (from a reset state), reset where a hand-off chain ended
        for (int i = 0; i < slots_; i++) {
            Slot &s = slots[(size_t)i];
            if (s.active() && s.k >= s.chain->size()) s.chain = nullptr;
            if (!s.active() && next_chain < chains.size()) {
                s = Slot();
                s.chain = chains[next_chain++];
                s.k = 0;
                s.left = s.chain->empty() ? 0 : (*s.chain)[0].frames;
                s.reset_before_next = true;
            }
            if (s.active() && s.reset_before_next) {
                if (fcv_batch_reset_slot(batch_, i) != 0) ok = false;
                s.reset_before_next = false;
            }
        }
        bool any = false;
        parallel([&](int i) {
            Slot &s = slots[(size_t)i];
            fv[(size_t)i] = 0;
            if (!s.active()) return;
            FillSlot(s, hin + (size_t)i * in_stride);
            fv[(size_t)i] = s.fill;
        });
        for (int i = 0; i < slots_; i++) any |= fv[(size_t)i] > 0;
        if (!any) {
            bool more = next_chain < chains.size();
            for (auto &s : slots) more |= s.active() && s.k < s.chain->size();
            if (!more) break;
            continue;  // only empty files / premature EOFs this round
        }
        if (fcv_batch_process(batch_, fv.data()) != 0) {
            syslog(LOG_ERR, "folve-b200: batch step failed: %s", fcv_last_error());
            return false;
        }
        bool need_max = false;
        for (auto &s : slots) need_max |= s.active() && s.nfinished > 0;
        if (need_max && fcv_batch_get_max(batch_, maxv.data()) != 0) ok = false;
        parallel([&](int i) {
            Slot &s = slots[(size_t)i];
            if (!s.active() || fv[(size_t)i] == 0) return;
            DrainSlot(s, hout + (size_t)i * out_stride, maxv[(size_t)i]);
        });
        steps_++;
        for (int i = 0; i < slots_; i++) blocks_ += fv[(size_t)i] > 0;
    }
    return ok;
}

}  // namespace folve_b200
