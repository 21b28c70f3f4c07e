// processor-pool.cc -- see processor-pool.h; behaviour follows
// /root/reference/processor-pool.cc:48-131 (folve, Copyright (C) 2012 Henner Zeller
// <h.zeller@acm.org>, GPL v3 or later -- see COPYING; same terms here).
#include "processor-pool.h"

#include <stdio.h>
#include <string.h>
#include <syslog.h>
#include <unistd.h>

#include "sound-processor.h"

ProcessorPool::ProcessorPool(int max_available)
    : keep_per_config_(max_available) {}

ProcessorPool::~ProcessorPool() {
    for (auto &kv : idle_)
        for (SoundProcessor *p : kv.second) delete p;
}

SoundProcessor *ProcessorPool::GetOrCreate(const std::string &base_dir,
                                           int sampling_rate, int channels,
                                           int bits, std::string *errmsg) {
    return GetOrCreate(base_dir, sampling_rate, channels, bits, errmsg, std::string());
}

SoundProcessor *ProcessorPool::GetOrCreate(const std::string &base_dir,
                                           int sampling_rate, int channels,
                                           int bits, std::string *errmsg,
                                           const std::string &placement_key) {
    // From specific to non-specific (processor-pool.cc:53-61).
    char name[3][96];
    snprintf(name[0], sizeof(name[0]), "/filter-%d-%d-%d.conf", sampling_rate, channels, bits);
    snprintf(name[1], sizeof(name[1]), "/filter-%d-%d.conf", sampling_rate, channels);
    snprintf(name[2], sizeof(name[2]), "/filter-%d.conf", sampling_rate);
    std::string config_path;
    for (int i = 0; i < 3 && config_path.empty(); ++i) {
        const std::string candidate = base_dir + name[i];
        if (access(candidate.c_str(), R_OK) == 0) config_path = candidate;
    }
    if (config_path.empty()) {
        const size_t slash = base_dir.find_last_of('/');
        const std::string short_dir = slash == std::string::npos ? base_dir : base_dir.substr(slash + 1);
        char msg[256];
        snprintf(msg, sizeof(msg), "No filter in %s for %.1fkHz/%d ch/%d bits",
                 short_dir.c_str(), sampling_rate / 1000.0, channels, bits);
        if (errmsg) *errmsg = msg;
        return NULL;
    }

    // a keyed request wants its album's GPU; an unkeyed one takes whatever is idle
    int device = SoundProcessor::Device();
    if (device < 0 && !placement_key.empty())
        device = SoundProcessor::DeviceForKey(placement_key, SoundProcessor::DeviceCount());
    SoundProcessor *result;
    while ((result = TakeIdle(config_path, device)) != NULL) {
        if (result->ConfigStillUpToDate()) return result;
        delete result;  // configuration or impulse file was touched since
    }
    result = device >= 0 ? SoundProcessor::CreateOnDevice(config_path, sampling_rate, channels, device)
                         : SoundProcessor::Create(config_path, sampling_rate, channels);
    if (result == NULL) {
        if (errmsg) *errmsg = "Problem parsing " + config_path;
        syslog(LOG_ERR, "filter-config %s is broken.", config_path.c_str());
    }
    return result;
}

void ProcessorPool::Return(SoundProcessor *processor) {
    if (processor == NULL) return;
    if (!processor->ConfigStillUpToDate()) {
        delete processor;
        return;
    }
    {
        std::lock_guard<std::mutex> l(pool_mutex_);
        IdleList &list = idle_[processor->config_file()];
        if (list.size() < keep_per_config_) {
            processor->Reset();
            list.push_back(processor);
            return;
        }
    }
    delete processor;  // enough idle processors for this configuration
}

SoundProcessor *ProcessorPool::TakeIdle(const std::string &config_path, int device) {
    std::lock_guard<std::mutex> l(pool_mutex_);
    IdleMap::iterator found = idle_.find(config_path);
    if (found == idle_.end()) return NULL;
    IdleList &list = found->second;
    for (IdleList::iterator it = list.begin(); it != list.end(); ++it) {
        if (device >= 0 && (*it)->device() != device) continue;
        SoundProcessor *result = *it;
        list.erase(it);
        return result;
    }
    return NULL;
}
