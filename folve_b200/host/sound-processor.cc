// sound-processor.cc -- see sound-processor.h.  Semantics follow
// /root/reference/sound-processor.cc:34-145 line by line; the arithmetic runs
// in the CUDA engine (include/folve_b200.h).
#include "sound-processor.h"

#include <assert.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <syslog.h>

#include <map>
#include <mutex>

#include "../../include/folve_b200.h"
#include "filter-config.h"

namespace {

std::mutex g_mutex;  // guards the device choice and the filter cache
int g_device = -1;

// Filter spectra stay resident in HBM and are shared by every processor made
// from the same configuration file: one entry per (path, mtime, device).
struct CachedFilter {
    fcv_filter *filter;
    int fragm, ninp, nout;
};
struct CacheKey {
    std::string path;
    time_t mtime;
    int device;
    bool operator<(const CacheKey &o) const {
        if (path != o.path) return path < o.path;
        if (mtime != o.mtime) return mtime < o.mtime;
        return device < o.device;
    }
};
std::map<CacheKey, CachedFilter> g_filters;

time_t ModificationTime(const std::string &filename) {
    struct stat st;
    if (stat(filename.c_str(), &st) != 0) return 0;
    return st.st_mtime;
}

int CurrentDevice() {
    if (g_device < 0) {
        const char *env = getenv("FOLVE_B200_DEVICE");
        g_device = env ? atoi(env) : 0;
    }
    return g_device;
}

}  // namespace

void SoundProcessor::SetDevice(int device) {
    std::lock_guard<std::mutex> l(g_mutex);
    g_device = device;
}

int SoundProcessor::Device() {
    std::lock_guard<std::mutex> l(g_mutex);
    return CurrentDevice();
}

void SoundProcessor::PurgeFilterCache() {
    std::lock_guard<std::mutex> l(g_mutex);
    for (auto &kv : g_filters) fcv_filter_unref(kv.second.filter);
    g_filters.clear();
}

SoundProcessor *SoundProcessor::Create(const std::string &config_file,
                                       int samplerate, int channels) {
    CachedFilter cf;
    const time_t mtime = ModificationTime(config_file);
    {
        // The reference serialises creation too (fftw_mutex, sound-processor.cc:29-43).
        std::lock_guard<std::mutex> l(g_mutex);
        const CacheKey key{config_file, mtime, CurrentDevice()};
        auto it = g_filters.find(key);
        if (it == g_filters.end()) {
            // stale versions of the same file are no longer wanted
            for (auto old = g_filters.begin(); old != g_filters.end();) {
                if (old->first.path == config_file && old->first.device == key.device) {
                    fcv_filter_unref(old->second.filter);
                    old = g_filters.erase(old);
                } else {
                    ++old;
                }
            }
            folve_b200::FilterConfig cfg;
            cfg.fsamp = samplerate;
            cfg.ninp = channels;
            cfg.nout = channels;
            if (folve_b200::LoadFilterConfig(&cfg, config_file.c_str()) != 0 || cfg.filter == NULL) {
                if (cfg.filter) fcv_filter_unref(cfg.filter);
                return NULL;  // parse error, or no /convolver/new (sound-processor.cc:44-48)
            }
            if (fcv_filter_commit(cfg.filter, key.device) != 0) {
                syslog(LOG_ERR, "folve-b200: %s: %s", config_file.c_str(), fcv_last_error());
                fcv_filter_unref(cfg.filter);
                return NULL;
            }
            cf.filter = cfg.filter;
            cf.fragm = cfg.fragm;
            cf.ninp = cfg.ninp;
            cf.nout = cfg.nout;
            it = g_filters.insert(std::make_pair(key, cf)).first;
        }
        cf = it->second;
        fcv_filter_ref(cf.filter);  // the processor's own reference
    }
    fcv_stream *stream = fcv_stream_create(cf.filter);
    if (!stream) {
        syslog(LOG_ERR, "folve-b200: %s: %s", config_file.c_str(), fcv_last_error());
        fcv_filter_unref(cf.filter);
        return NULL;
    }
    return new SoundProcessor(cf.filter, stream, cf.fragm, cf.ninp, cf.nout, config_file, mtime);
}

SoundProcessor::SoundProcessor(fcv_filter *filter, fcv_stream *stream, int fragm,
                               int ninp, int nout, const std::string &cfg,
                               time_t cfg_mtime)
    : filter_(filter), stream_(stream), fragm_(fragm), ninp_(ninp), nout_(nout),
      config_file_(cfg), config_file_timestamp_(cfg_mtime),
      buffer_(fcv_stream_buffer(stream)),
      filled_(0), drained_(-1), peak_seen_(0.0) {
    // a fresh stream is already in the reset state
}

SoundProcessor::~SoundProcessor() {
    fcv_stream_destroy(stream_);
    fcv_filter_unref(filter_);
}

int SoundProcessor::FillBuffer(SNDFILE *in) {
    const int samples_needed = fragm_ - filled_;
    assert(samples_needed);  // Otherwise, call WriteProcessed() first.
    drained_ = -1;
    const int r = sf_readf_float(in, buffer_ + filled_ * ninp_, samples_needed);
    filled_ += r;
    return r;
}

void SoundProcessor::WriteProcessed(SNDFILE *out, int sample_count) {
    if (drained_ < 0) Process();
    assert(sample_count <= fragm_ - drained_);
    sf_writef_float(out, buffer_ + drained_ * nout_, sample_count);
    drained_ += sample_count;
    if (drained_ == fragm_) filled_ = 0;
}

void SoundProcessor::Process() {
    // One synchronous block: the engine takes the first filled_ interleaved
    // frames of buffer_, treats the rest of the block as silence, and writes the
    // first filled_ interleaved output frames back, raising the signed maximum.
    if (fcv_stream_process(stream_, filled_, &peak_seen_) != 0) {
        syslog(LOG_ERR, "folve-b200: processing failed: %s", fcv_last_error());
        memset(buffer_, 0, sizeof(float) * fragm_ * nout_);  // silence, never stale input
    }
    drained_ = 0;
}

bool SoundProcessor::ConfigStillUpToDate() const {
    return config_file_timestamp_ == ModificationTime(config_file_);
}

void SoundProcessor::ResetMaxValues() { peak_seen_ = 0.0; }

void SoundProcessor::Reset() {
    fcv_stream_reset(stream_);
    filled_ = 0;
    drained_ = -1;
    ResetMaxValues();
}
