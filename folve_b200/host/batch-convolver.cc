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

// One block of a chain.  It may contain the tail of file `k` and the head of file
// `k + 1`; `share[]` says how the processed block is split between them
// (SoundProcessor::WriteProcessed(out, r) and the successor's pending_writes(),
// convolve-file-handler.cc:373-376,408).
struct BatchConvolver::BlockPlan {
    int fill = 0;          // frames in the block (frames_valid)
    struct { size_t file; int frames; } share[2];
    int nshare = 0;
    size_t finished[2];    // files completed by this block (receive the running max)
    int nfinished = 0;
};

// One chain in flight.
struct BatchConvolver::Slot {
    Chain *chain = nullptr;
    size_t k = 0;          // file being read
    long left = 0;         // frames of file k not yet read
    BlockPlan block[8];    // the blocks of the step being assembled
    int nblocks = 0;
    int fill = 0;          // frames of the whole step (frames_valid of the batch call)
    bool reset_before_next = false;  // next block starts from the fresh state
    float running_max = 0.0f;        // SoundProcessor::max_output_value() of the chain's processor
    bool active() const { return chain != nullptr; }
};

BatchConvolver *BatchConvolver::Create(const std::string &config_file, int samplerate, int channels, int slots,
                                       bool gapless, int device, int blocks_per_step) {
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
    fcv_batch *batch = fcv_batch_create_tiled(cfg.filter, slots, FCV_PCM_F32, FCV_PCM_F32, blocks_per_step);
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
    bc->tblocks_ = blocks_per_step;
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
void BatchConvolver::FillBlock(Slot &s, BlockPlan &p, float *in_block) {
    Chain &c = *s.chain;
    p = BlockPlan();
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
    p.fill = r;
    p.share[p.nshare].file = s.k;
    p.share[p.nshare++].frames = r;
    if (s.left > 0) return;  // a full block from the middle of the file
    // file k ends with this block
    p.finished[p.nfinished++] = s.k;
    if (p.fill == fragm_ || !gapless_ || s.k + 1 >= c.size()) {
        // block complete (no hand-off, quirk 3), or nobody to hand over to
        s.k++;
        s.left = s.k < c.size() ? c[s.k].frames : 0;
        s.reset_before_next = true;
        return;
    }
    // hand the half-filled block over to the alphabetically next file
    ChainFile &b = c[s.k + 1];
    int r2 = (int)((long)(fragm_ - p.fill) < b.frames ? (fragm_ - p.fill) : b.frames);
    r2 = (int)sf_readf_float(b.in, in_block + (size_t)p.fill * ninp_, r2);
    a.out_gapless = true;
    b.in_gapless = true;
    p.fill += r2;
    s.k++;
    s.left = b.frames - r2;
    if (s.left == 0) {
        // the successor was swallowed by the top-up (quirk 4): it writes nothing,
        // and whoever comes next starts fresh
        p.finished[p.nfinished++] = s.k;
        s.k++;
        s.left = s.k < c.size() ? c[s.k].frames : 0;
        s.reset_before_next = true;
    } else {
        // the successor owns the rest of the block (its pending_writes())
        p.share[p.nshare].file = s.k;
        p.share[p.nshare++].frames = fragm_ - r;
    }
}

// Assemble the next step of a chain: up to blocks_per_step consecutive blocks.  The step
// ends early where the per-file path would go on from a reset processor (end of a chain,
// no hand-off, swallowed successor, premature EOF): a time-tiled step cannot reset a
// stream between two of its blocks, and only the last block of a step may be short.
void BatchConvolver::FillSlot(Slot &s, float *in_step) {
    s.nblocks = 0;
    s.fill = 0;
    s.reset_before_next = false;
    const size_t in_stride = (size_t)fragm_ * ninp_;
    while (s.nblocks < tblocks_) {
        BlockPlan &p = s.block[s.nblocks];
        FillBlock(s, p, in_step + (size_t)s.nblocks * in_stride);
        if (p.fill > 0) {
            s.nblocks++;
            s.fill += p.fill;
        }
        if (s.reset_before_next || s.k >= s.chain->size() || p.fill < fragm_) break;
    }
}

void BatchConvolver::DrainSlot(Slot &s, const float *out_step, const float *block_max) {
    Chain &c = *s.chain;
    const size_t out_stride = (size_t)fragm_ * nout_;
    for (int t = 0; t < s.nblocks; t++) {
        const BlockPlan &p = s.block[t];
        const float *out_block = out_step + (size_t)t * out_stride;
        int pos = 0;
        for (int i = 0; i < p.nshare; i++) {
            ChainFile &f = c[p.share[i].file];
            sf_writef_float(f.out, out_block + (size_t)pos * nout_, p.share[i].frames);
            f.written += p.share[i].frames;
            pos += p.share[i].frames;
        }
        // max_out_value_observed_ after this block (sound-processor.cc:120-123)
        if (block_max[t] > s.running_max) s.running_max = block_max[t];
        for (int i = 0; i < p.nfinished; i++) c[p.finished[i]].max_value = s.running_max;
    }
}

bool BatchConvolver::Run(const std::vector<Chain *> &chains, int threads) {
    if (threads < 1) threads = 1;
    std::vector<Slot> slots((size_t)slots_);
    std::vector<int> fv((size_t)slots_, 0);
    std::vector<float> maxv((size_t)slots_ * tblocks_, 0.0f);
    float *hin = (float *)fcv_batch_host_in(batch_);
    const float *hout = (const float *)fcv_batch_host_out(batch_);
    const size_t in_stride = (size_t)tblocks_ * fragm_ * ninp_, out_stride = (size_t)tblocks_ * fragm_ * nout_;
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
                s.running_max = 0.0f;   // SoundProcessor::Reset() (sound-processor.cc:139-145)
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
        if (fcv_batch_get_block_max(batch_, maxv.data()) != 0) ok = false;
        parallel([&](int i) {
            Slot &s = slots[(size_t)i];
            if (!s.active() || fv[(size_t)i] == 0) return;
            DrainSlot(s, hout + (size_t)i * out_stride, maxv.data() + (size_t)i * tblocks_);
        });
        steps_++;
        for (int i = 0; i < slots_; i++) blocks_ += slots[(size_t)i].active() && fv[(size_t)i] > 0 ? slots[(size_t)i].nblocks : 0;
    }
    return ok;
}

}  // namespace folve_b200
