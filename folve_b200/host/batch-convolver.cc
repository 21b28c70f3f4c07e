// batch-convolver.cc -- see batch-convolver.h.
#include "batch-convolver.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <syslog.h>
#include <time.h>

#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

#include "../../include/folve_b200.h"
#include "filter-config.h"
#include "sound-processor.h"

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

// One step of a chain as it was assembled: kept until its output has come back, while the
// next step of the same slot is already being assembled (two steps are in flight).
struct BatchConvolver::StepPlan {
    Chain *chain = nullptr;  // whose files the shares refer to
    BlockPlan block[8];
    int nblocks = 0;
    int fill = 0;            // frames of the whole step (frames_valid of the batch call)
    bool fresh = false;      // the step started from a reset processor: the running maximum restarts
};

// One chain in flight.
struct BatchConvolver::Slot {
    Chain *chain = nullptr;
    size_t k = 0;          // file being read
    long left = 0;         // frames of file k not yet read
    BlockPlan *block = nullptr;  // the blocks of the step being assembled (in one of plan[])
    int nblocks = 0;
    int fill = 0;
    bool reset_before_next = false;  // next block starts from the fresh state
    StepPlan plan[2];                // per host staging slot
    float running_max = 0.0f;        // SoundProcessor::max_output_value() of the chain's processor (drain side)
    bool active() const { return chain != nullptr; }
};

// A fixed set of host threads that run `fn(i)` for i in [0, n) on request.
class BatchConvolver::Workers {
public:
    explicit Workers(int n) {
        for (int t = 0; t < n; t++) pool_.emplace_back([this] { Loop(); });
    }
    ~Workers() {
        {
            std::lock_guard<std::mutex> l(mu_);
            quit_ = true;
        }
        cv_.notify_all();
        for (auto &th : pool_) th.join();
    }
    void Run(int n, const std::function<void(int)> &fn) {
        if (pool_.empty()) {
            for (int i = 0; i < n; i++) fn(i);
            return;
        }
        std::unique_lock<std::mutex> l(mu_);
        fn_ = &fn;
        n_ = n;
        next_ = 0;
        busy_ = (int)pool_.size();
        gen_++;
        cv_.notify_all();
        done_.wait(l, [this] { return busy_ == 0; });
        fn_ = nullptr;
    }

private:
    void Loop() {
        unsigned long seen = 0;
        std::unique_lock<std::mutex> l(mu_);
        for (;;) {
            cv_.wait(l, [&] { return quit_ || gen_ != seen; });
            if (quit_) return;
            seen = gen_;
            const std::function<void(int)> *fn = fn_;
            const int n = n_;
            l.unlock();
            for (int i = next_.fetch_add(1); i < n; i = next_.fetch_add(1)) (*fn)(i);
            l.lock();
            if (--busy_ == 0) done_.notify_one();
        }
    }
    std::vector<std::thread> pool_;
    std::mutex mu_;
    std::condition_variable cv_, done_;
    const std::function<void(int)> *fn_ = nullptr;
    int n_ = 0, busy_ = 0;
    std::atomic<int> next_{0};
    unsigned long gen_ = 0;
    bool quit_ = false;
};

BatchConvolver *BatchConvolver::Create(const std::string &config_file, int samplerate, int channels, int slots,
                                       bool gapless, int device, int blocks_per_step, bool pcm16) {
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
    const int wire = pcm16 ? FCV_PCM_S16 : FCV_PCM_F32;
    fcv_batch *batch = fcv_batch_create_tiled(cfg.filter, slots, wire, wire, blocks_per_step);
    if (!batch) {
        syslog(LOG_ERR, "folve-b200: %s: %s", config_file.c_str(), fcv_last_error());
        fcv_filter_unref(cfg.filter);
        return nullptr;
    }
    // both host staging slots up front: pinning ~1 GB takes a few hundred milliseconds
    if (!fcv_batch_host_in_slot(batch, 1) || !fcv_batch_host_out_slot(batch, 1)) {
        syslog(LOG_ERR, "folve-b200: %s: %s", config_file.c_str(), fcv_last_error());
        fcv_batch_destroy(batch);
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
    bc->pcm16_ = pcm16;
    bc->sample_bytes_ = pcm16 ? sizeof(short) : sizeof(float);
    return bc;
}

BatchConvolver::~BatchConvolver() {
    if (batch_) fcv_batch_destroy(batch_);
    if (filter_) fcv_filter_unref(filter_);
}

// Assemble the next block of a chain: FillBuffer on file k, and -- if that file
// ends inside the block and gapless joining is on -- one top-up from file k+1
// (PassoverProcessor, convolve-file-handler.cc:345-348).
// sf_readf_* of `frames` frames into the block at `dst`, `frame_offset` frames in
long BatchConvolver::ReadFrames(SNDFILE *in, void *dst, long frame_offset, long frames) const {
    char *p = (char *)dst + (size_t)frame_offset * ninp_ * sample_bytes_;
    return pcm16_ ? (long)sf_readf_short(in, (short *)p, frames) : (long)sf_readf_float(in, (float *)p, frames);
}

void BatchConvolver::FillBlock(Slot &s, BlockPlan &p, void *in_block) {
    Chain &c = *s.chain;
    p = BlockPlan();
    // skip empty files: AddMoreSoundData returns false at once for them
    while (s.k < c.size() && s.left == 0) {
        if (++s.k < c.size()) s.left = c[s.k].frames;
    }
    if (s.k >= c.size()) return;
    ChainFile &a = c[s.k];
    const int want = (int)(s.left < fragm_ ? s.left : fragm_);
    const int r = (int)ReadFrames(a.in, in_block, 0, want);
    if (r == 0) {  // premature EOF: the file is over, nothing is written for it
        s.left = 0;
        s.reset_before_next = true;
        return;
    }
    s.left -= r;
    p.fill = r;
    p.share[p.nshare].file = s.k;
    p.share[p.nshare++].frames = r;
    if (r < want) {
        // A short read in the middle of a file (truncated / corrupt input): the file ends here.
        // It gets the frames that were read, hands nothing over, and whoever comes next starts
        // from a reset processor -- zero padding is never spliced into a running convolution.
        s.left = 0;
        p.finished[p.nfinished++] = s.k;
        s.k++;
        s.left = s.k < c.size() ? c[s.k].frames : 0;
        s.reset_before_next = true;
        return;
    }
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
    const int want2 = (int)((long)(fragm_ - p.fill) < b.frames ? (fragm_ - p.fill) : b.frames);
    const int r2 = (int)ReadFrames(b.in, in_block, p.fill, want2);
    a.out_gapless = true;
    b.in_gapless = true;
    p.fill += r2;
    s.k++;
    s.left = r2 < want2 ? 0 : b.frames - r2;   // a successor that is shorter than it claims ends with the top-up
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
void BatchConvolver::FillSlot(Slot &s, void *in_step) {
    s.nblocks = 0;
    s.fill = 0;
    s.reset_before_next = false;
    const size_t in_stride = (size_t)fragm_ * ninp_ * sample_bytes_;
    while (s.nblocks < tblocks_) {
        BlockPlan &p = s.block[s.nblocks];
        FillBlock(s, p, (char *)in_step + (size_t)s.nblocks * in_stride);
        if (p.fill > 0) {
            s.nblocks++;
            s.fill += p.fill;
        }
        if (s.reset_before_next || s.k >= s.chain->size() || p.fill < fragm_) break;
    }
}

void BatchConvolver::DrainSlot(Slot &s, const StepPlan &sp, const void *out_step, const float *block_max) {
    Chain &c = *sp.chain;
    const size_t out_stride = (size_t)fragm_ * nout_ * sample_bytes_;
    if (sp.fresh) s.running_max = 0.0f;   // SoundProcessor::Reset() (sound-processor.cc:139-145)
    for (int t = 0; t < sp.nblocks; t++) {
        const BlockPlan &p = sp.block[t];
        const char *out_block = (const char *)out_step + (size_t)t * out_stride;
        int pos = 0;
        for (int i = 0; i < p.nshare; i++) {
            ChainFile &f = c[p.share[i].file];
            const char *src = out_block + (size_t)pos * nout_ * sample_bytes_;
            if (pcm16_) sf_writef_short(f.out, (const short *)src, p.share[i].frames);
            else sf_writef_float(f.out, (const float *)src, p.share[i].frames);
            f.written += p.share[i].frames;
            pos += p.share[i].frames;
        }
        // max_out_value_observed_ after this block (sound-processor.cc:120-123)
        if (block_max[t] > s.running_max) s.running_max = block_max[t];
        for (int i = 0; i < p.nfinished; i++) c[p.finished[i]].max_value = s.running_max;
    }
}

// Two steps are in flight: while the GPU works on step k (host staging slot k & 1) the host
// threads write the output of step k - 1 to the files and read the input of step k + 1.
// Resets between steps are enqueued on the device (fcv_batch_reset_slot_async), the per-block
// maxima come back with each step's output: the host never waits for anything but a finished step.
bool BatchConvolver::Run(const std::vector<Chain *> &chains, int threads) {
    if (threads < 1) threads = 1;
    std::vector<Slot> slots((size_t)slots_);
    std::vector<int> fv[2] = {std::vector<int>((size_t)slots_, 0), std::vector<int>((size_t)slots_, 0)};
    char *hin[2] = {(char *)fcv_batch_host_in_slot(batch_, 0), (char *)fcv_batch_host_in_slot(batch_, 1)};
    const char *hout[2] = {(const char *)fcv_batch_host_out_slot(batch_, 0),
                           (const char *)fcv_batch_host_out_slot(batch_, 1)};
    if (!hin[0] || !hin[1] || !hout[0] || !hout[1]) return false;
    const size_t in_stride = (size_t)tblocks_ * fragm_ * ninp_ * sample_bytes_;
    const size_t out_stride = (size_t)tblocks_ * fragm_ * nout_ * sample_bytes_;
    size_t next_chain = 0;
    bool ok = true;
    Workers workers(threads > 1 ? threads : 0);

    // FOLVE_B200_TRACE=1: where the host spends a step (seconds, summed over the run) on stderr
    static const bool trace = getenv("FOLVE_B200_TRACE") != nullptr;
    double t_wait = 0, t_drain = 0, t_assign = 0, t_fill = 0, t_submit = 0;
    auto now = [] {
        timespec t;
        clock_gettime(CLOCK_MONOTONIC, &t);
        return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
    };
    // output of the step submitted from staging slot `h` -> files
    auto drain = [&](int h) {
        const double t0 = now();
        if (fcv_batch_wait(batch_, h) != 0) {
            syslog(LOG_ERR, "folve-b200: batch step failed: %s", fcv_last_error());
            ok = false;
            return;
        }
        const double t1 = now();
        t_wait += t1 - t0;
        const float *bmax = fcv_batch_host_block_max_slot(batch_, h);
        workers.Run(slots_, [&](int i) {
            Slot &s = slots[(size_t)i];
            const StepPlan &sp = s.plan[h];
            if (!sp.chain || fv[h][(size_t)i] == 0) return;
            DrainSlot(s, sp, hout[h] + (size_t)i * out_stride, bmax + (size_t)i * tblocks_);
        });
        for (int i = 0; i < slots_; i++) {
            blocks_ += fv[h][(size_t)i] > 0 ? slots[(size_t)i].plan[h].nblocks : 0;
            fv[h][(size_t)i] = 0;
        }
        t_drain += now() - t1;
    };

    bool pending[2] = {false, false};
    for (long step = 0;; step++) {
        const int h = (int)(step & 1);
        if (pending[h]) {  // the staging slot is about to be refilled: its previous step has to be out
            drain(h);
            pending[h] = false;
        }
        // hand idle slots a new chain (from a reset state), reset where a hand-off chain ended
        const double ta = now();
        for (int i = 0; i < slots_; i++) {
            Slot &s = slots[(size_t)i];
            if (s.active() && s.k >= s.chain->size()) s.chain = nullptr;
            if (!s.active() && next_chain < chains.size()) {
                s.chain = chains[next_chain++];
                s.k = 0;
                s.left = s.chain->empty() ? 0 : (*s.chain)[0].frames;
                s.reset_before_next = true;
            }
            s.plan[h].chain = s.chain;
            s.plan[h].fresh = false;
            s.plan[h].nblocks = 0;
            if (s.active() && s.reset_before_next) {
                if (fcv_batch_reset_slot_async(batch_, i) != 0) ok = false;
                s.reset_before_next = false;
                s.plan[h].fresh = true;
            }
        }
        const double tf = now();
        t_assign += tf - ta;
        workers.Run(slots_, [&](int i) {
            Slot &s = slots[(size_t)i];
            fv[h][(size_t)i] = 0;
            if (!s.active()) return;
            s.block = s.plan[h].block;
            FillSlot(s, hin[h] + (size_t)i * in_stride);
            s.plan[h].nblocks = s.nblocks;
            s.plan[h].fill = s.fill;
            fv[h][(size_t)i] = s.fill;
        });
        const double ts = now();
        t_fill += ts - tf;
        bool any = false;
        for (int i = 0; i < slots_; i++) any |= fv[h][(size_t)i] > 0;
        if (!any) {
            bool more = next_chain < chains.size();
            for (auto &s : slots) more |= s.active() && s.k < s.chain->size();
            if (!more) break;
            step--;        // only empty files / premature EOFs this round: same staging slot again
            continue;
        }
        if (fcv_batch_submit(batch_, h, fv[h].data()) != 0) {
            syslog(LOG_ERR, "folve-b200: batch step failed: %s", fcv_last_error());
            return false;
        }
        t_submit += now() - ts;
        pending[h] = true;
        steps_++;
    }
    // drain what is still in flight, oldest first
    const int last = (int)(steps_ & 1);   // staging slot the next step would have used = the older one
    if (pending[last]) drain(last);
    if (pending[last ^ 1]) drain(last ^ 1);
    if (trace)
        fprintf(stderr, "BatchConvolver::Run: %ld steps; host seconds: wait %.3f drain %.3f assign+reset %.3f fill %.3f submit %.3f\n",
                steps_, t_wait, t_drain, t_assign, t_fill, t_submit);
    return ok;
}

// ---- every GPU of the box from one process ---------------------------------------------
MultiDeviceConvolver *MultiDeviceConvolver::Create(const std::string &config_file, int samplerate, int channels,
                                                   int slots_per_device, bool gapless,
                                                   const std::vector<int> &devices, int blocks_per_step,
                                                   bool pcm16, int instances_per_device) {
    std::vector<int> ids = devices;
    if (ids.empty()) {
        const int n = fcv_device_count();
        for (int d = 0; d < n; d++) ids.push_back(d);
    }
    if (ids.empty()) return nullptr;
    MultiDeviceConvolver *m = new MultiDeviceConvolver();
    m->inst_ = instances_per_device > 1 ? instances_per_device : 1;
    const int slots = (slots_per_device + m->inst_ - 1) / m->inst_;
    for (int d : ids) {
        for (int k = 0; k < m->inst_; k++) {
            BatchConvolver *bc = BatchConvolver::Create(config_file, samplerate, channels, slots, gapless, d,
                                                        blocks_per_step, pcm16);
            if (!bc) {
                delete m;
                return nullptr;
            }
            m->parts_.push_back(bc);
        }
        m->device_ids_.push_back(d);
    }
    return m;
}

MultiDeviceConvolver::~MultiDeviceConvolver() {
    for (BatchConvolver *bc : parts_) delete bc;
}

int MultiDeviceConvolver::fragment_size() const { return parts_.empty() ? 0 : parts_[0]->fragment_size(); }
int MultiDeviceConvolver::output_channels() const { return parts_.empty() ? 0 : parts_[0]->output_channels(); }

int MultiDeviceConvolver::PlacementOf(const std::string &key) const {
    return SoundProcessor::DeviceForKey(key, (int)device_ids_.size());
}

long MultiDeviceConvolver::blocks_processed() const {
    long n = 0;
    for (const BatchConvolver *bc : parts_) n += bc->blocks_processed();
    return n;
}

bool MultiDeviceConvolver::Run(const std::vector<Chain *> &chains, const std::vector<std::string> &keys,
                               int threads_per_device, std::vector<int> *assignment) {
    const size_t nd = device_ids_.size(), np = parts_.size();
    std::vector<std::vector<Chain *> > shard(np);
    std::vector<size_t> next(nd, 0);   // the chains of a device go round its instances
    if (assignment) assignment->assign(chains.size(), 0);
    for (size_t i = 0; i < chains.size(); i++) {
        const size_t d = keys.empty() ? i % nd : (size_t)PlacementOf(keys[i]);
        shard[d * (size_t)inst_ + next[d]++ % (size_t)inst_].push_back(chains[i]);
        if (assignment) (*assignment)[i] = (int)d;
    }
    // every instance gets the device's full thread count: the instances of a device take turns -- one fills and
    // drains while the others wait for their steps (measured: dividing the threads loses all of the gain)
    const int threads = threads_per_device;
    std::vector<char> ok(np, 1);
    std::vector<std::thread> th;
    for (size_t p = 0; p < np; p++)
        th.emplace_back([&, p] { ok[p] = parts_[p]->Run(shard[p], threads) ? 1 : 0; });
    for (auto &t : th) t.join();
    bool all = true;
    for (char o : ok) all = all && o;
    return all;
}

}  // namespace folve_b200
