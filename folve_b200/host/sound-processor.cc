// sound-processor.cc -- see sound-processor.h.  Semantics follow
// /root/reference/sound-processor.cc:34-145 line by line; the arithmetic runs
// in the CUDA engine (include/folve_b200.h).
//
// The block protocol (FillBuffer / WriteProcessed / Process) is that of folve's
// SoundProcessor, Copyright (C) 2012 Henner Zeller <h.zeller@acm.org>, GPL v3 or later
// (see COPYING); this file is distributed under the same terms.
#include "sound-processor.h"

#include <assert.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <syslog.h>

#include <map>
#include <mutex>

#include "../../include/folve_b200.h"
#include "filter-config.h"

using folve_b200::FileStamp;

namespace {

std::mutex g_mutex;  // guards the device choice, the live counts and the filter cache
int g_device = -2;   // -2: not decided yet; kAnyDevice; or a fixed device
int g_ndevices = -1;
std::vector<int> g_live;              // live processors per device
thread_local std::string t_placement_key;

// Filter spectra stay resident in HBM and are shared by every processor made from the same
// configuration: one entry per (config path, device), valid as long as the config file AND every
// impulse file it read keep their nanosecond mtime and size.
struct CachedFilter {
    fcv_filter *filter;
    int fragm, ninp, nout;
    std::shared_ptr<const std::vector<FileStamp> > stamps;  // [0] = the config file itself
};
typedef std::pair<std::string, int> CacheKey;
std::map<CacheKey, CachedFilter> g_filters;

int DeviceCountLocked() {
    if (g_ndevices < 0) {
        const int n = fcv_device_count();
        g_ndevices = n > 0 ? n : 0;
        g_live.assign((size_t)(g_ndevices > 0 ? g_ndevices : 1), 0);
    }
    return g_ndevices;
}

int FixedDeviceLocked() {
    if (g_device == -2) {
        const char *env = getenv("FOLVE_B200_DEVICE");
        g_device = (env && *env) ? atoi(env) : SoundProcessor::kAnyDevice;
    }
    return g_device;
}

void CountLive(int device, int delta) {
    std::lock_guard<std::mutex> l(g_mutex);
    DeviceCountLocked();
    if (device >= 0 && (size_t)device < g_live.size()) g_live[(size_t)device] += delta;
}

uint32_t Crc32(const std::string &s) {  // reflected 0xEDB88320, as zlib.crc32
    uint32_t c = 0xffffffffu;
    for (unsigned char ch : s) {
        c ^= ch;
        for (int k = 0; k < 8; k++) c = (c >> 1) ^ (0xedb88320u & (0u - (c & 1u)));
    }
    return c ^ 0xffffffffu;
}

}  // namespace

void SoundProcessor::SetDevice(int device) {
    std::lock_guard<std::mutex> l(g_mutex);
    g_device = device < 0 ? kAnyDevice : device;
}

int SoundProcessor::Device() {
    std::lock_guard<std::mutex> l(g_mutex);
    return FixedDeviceLocked();
}

int SoundProcessor::DeviceCount() {
    std::lock_guard<std::mutex> l(g_mutex);
    return DeviceCountLocked();
}

int SoundProcessor::DeviceForKey(const std::string &key, int ndevices) {
    return ndevices > 1 ? (int)(Crc32(key) % (uint32_t)ndevices) : 0;
}

void SoundProcessor::SetPlacementKey(const std::string &key) { t_placement_key = key; }

int SoundProcessor::LiveProcessors(int device) {
    std::lock_guard<std::mutex> l(g_mutex);
    DeviceCountLocked();
    return device >= 0 && (size_t)device < g_live.size() ? g_live[(size_t)device] : 0;
}

void SoundProcessor::PurgeFilterCache() {
    std::lock_guard<std::mutex> l(g_mutex);
    for (auto &kv : g_filters) fcv_filter_unref(kv.second.filter);
    g_filters.clear();
}

SoundProcessor *SoundProcessor::Create(const std::string &config_file,
                                       int samplerate, int channels) {
    int device;
    {
        std::lock_guard<std::mutex> l(g_mutex);
        device = FixedDeviceLocked();
        if (device == kAnyDevice) {
            const int n = DeviceCountLocked();
            if (!t_placement_key.empty()) {
                device = DeviceForKey(t_placement_key, n);
            } else {
                device = 0;
                for (int d = 1; d < n; d++)
                    if (g_live[(size_t)d] < g_live[(size_t)device]) device = d;
            }
        }
    }
    return CreateOnDevice(config_file, samplerate, channels, device);
}

SoundProcessor *SoundProcessor::CreateOnDevice(const std::string &config_file,
                                               int samplerate, int channels, int device) {
    CachedFilter cf;
    {
        // The reference serialises creation too (fftw_mutex, sound-processor.cc:29-43).
        std::lock_guard<std::mutex> l(g_mutex);
        const CacheKey key(config_file, device);
        auto it = g_filters.find(key);
        if (it != g_filters.end() && !folve_b200::StampsCurrent(*it->second.stamps)) {
            // the config or one of its impulse files changed: that version is no longer wanted
            fcv_filter_unref(it->second.filter);
            g_filters.erase(it);
            it = g_filters.end();
        }
        if (it == g_filters.end()) {
            auto stamps = std::make_shared<std::vector<FileStamp> >();
            stamps->push_back(folve_b200::StampFile(config_file));  // before reading it
            folve_b200::FilterConfig cfg;
            cfg.fsamp = samplerate;
            cfg.ninp = channels;
            cfg.nout = channels;
            if (folve_b200::LoadFilterConfig(&cfg, config_file.c_str()) != 0 || cfg.filter == NULL) {
                if (cfg.filter) fcv_filter_unref(cfg.filter);
                return NULL;  // parse error, or no /convolver/new (sound-processor.cc:44-48)
            }
            if (fcv_filter_commit(cfg.filter, device) != 0) {
                syslog(LOG_ERR, "folve-b200: %s: %s", config_file.c_str(), fcv_last_error());
                fcv_filter_unref(cfg.filter);
                return NULL;
            }
            stamps->insert(stamps->end(), cfg.impulse_files.begin(), cfg.impulse_files.end());
            cf.filter = cfg.filter;
            cf.fragm = cfg.fragm;
            cf.ninp = cfg.ninp;
            cf.nout = cfg.nout;
            cf.stamps = stamps;
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
    return new SoundProcessor(cf.filter, stream, cf.fragm, cf.ninp, cf.nout, config_file,
                              (*cf.stamps)[0].sec, device, cf.stamps);
}

SoundProcessor::SoundProcessor(fcv_filter *filter, fcv_stream *stream, int fragm,
                               int ninp, int nout, const std::string &cfg,
                               time_t cfg_mtime, int device, const Stamps &stamps)
    : filter_(filter), stream_(stream), fragm_(fragm), ninp_(ninp), nout_(nout),
      config_file_(cfg), config_file_timestamp_(cfg_mtime), device_(device), stamps_(stamps),
      buffer_(fcv_stream_buffer(stream)),
      filled_(0), drained_(-1), peak_seen_(0.0) {
    // a fresh stream is already in the reset state
    CountLive(device_, +1);
}

SoundProcessor::~SoundProcessor() {
    fcv_stream_destroy(stream_);
    fcv_filter_unref(filter_);
    CountLive(device_, -1);
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
    // The reference compares the config file's mtime only and leaves the impulse files as a
    // TODO (sound-processor.cc:129-133); both are checked here, to the nanosecond and the byte.
    return folve_b200::StampsCurrent(*stamps_);
}

void SoundProcessor::ResetMaxValues() { peak_seen_ = 0.0; }

void SoundProcessor::Reset() {
    fcv_stream_reset(stream_);
    filled_ = 0;
    drained_ = -1;
    ResetMaxValues();
}
