// -*- c++ -*-
// processor-pool.h -- pool of SoundProcessors per configuration file; same
// interface as /root/reference/processor-pool.h:30-55 (folve, Copyright (C) 2012 Henner
// Zeller <h.zeller@acm.org>, GPL v3 or later -- see COPYING; same terms here).
//
// Processors are worth pooling for a different reason than in the reference:
// creating one no longer re-parses the filter or recomputes its spectra (those
// are cached in HBM per (config path, mtime), see sound-processor.cc), but a
// pooled processor keeps its device state and pinned block allocated.
#ifndef FOLVE_B200_PROCESSOR_POOL_H
#define FOLVE_B200_PROCESSOR_POOL_H

#include <deque>
#include <map>
#include <mutex>
#include <string>

class SoundProcessor;

class ProcessorPool {
public:
  // Stores at most "max_per_config" idle processors per configuration file.
  explicit ProcessorPool(int max_per_config);
  ~ProcessorPool();

  // Resolve base_dir/filter-<rate>-<channels>-<bits>.conf, then
  // filter-<rate>-<channels>.conf, then filter-<rate>.conf and hand out a
  // processor for the first one that is readable.  NULL + *errmsg on failure.
  SoundProcessor *GetOrCreate(const std::string &base_dir,
                              int sampling_rate, int channels, int bits,
                              std::string *errmsg);

  // Addition for boxes with several GPUs: the same, for a file of album directory
  // `placement_key` -- the processor lives on SoundProcessor::DeviceForKey(key), so every file of
  // an album meets its gapless neighbours on one GPU.  (The reference signature above balances
  // new processors over the GPUs by load instead.)
  SoundProcessor *GetOrCreate(const std::string &base_dir,
                              int sampling_rate, int channels, int bits,
                              std::string *errmsg, const std::string &placement_key);

  // Give a processor back; it is reset and kept, or deleted if its
  // configuration changed or the pool is full.
  void Return(SoundProcessor *processor);

private:
  typedef std::deque<SoundProcessor*> IdleList;
  typedef std::map<std::string, IdleList> IdleMap;

  // device < 0: any
  SoundProcessor *TakeIdle(const std::string &config_path, int device);

  const size_t keep_per_config_;
  std::mutex pool_mutex_;
  IdleMap idle_;
};

#endif  // FOLVE_B200_PROCESSOR_POOL_H
