// -*- c++ -*-
// sound-processor.h -- drop-in replacement for folve's SoundProcessor
// (/root/reference/sound-processor.h:28-85) on top of the B200 engine.
//
// The public interface and the FillBuffer / WriteProcessed / Process block protocol are those
// of folve's SoundProcessor, Copyright (C) 2012 Henner Zeller <h.zeller@acm.org>, GPL v3 or
// later (see COPYING); this file is distributed under the same terms.
//
// The public interface is the reference's, member for member, so that
// ConvolveFileHandler (convolve-file-handler.cc:78-80,335-348,373-377,408,418)
// and ProcessorPool (processor-pool.cc:72,83,95,109) compile and behave
// unchanged.  What differs is behind it: instead of a Convproc the object
// owns one fcv_stream (device-resident input-spectra ring + overlap tails) and
// a reference on a shared, HBM-resident fcv_filter; `buffer_` is the stream's
// pinned host block.
#ifndef FOLVE_B200_SOUND_PROCESSOR_H
#define FOLVE_B200_SOUND_PROCESSOR_H

#include <sndfile.h>
#include <time.h>

#include <memory>
#include <string>
#include <vector>

namespace folve_b200 { struct FileStamp; }
struct fcv_filter;
struct fcv_stream;

class SoundProcessor {
public:
  // NULL if the configuration cannot be parsed, defines no convolver, or no
  // usable GPU is present (there is no CPU fallback).
  static SoundProcessor *Create(const std::string &config_file,
                                int samplerate, int channels);
  ~SoundProcessor();

  // Reads up to the free part of the block from `in`; returns the frames read (sound-processor.h:38-39).
  int FillBuffer(SNDFILE *in);

  inline int input_channels() const { return ninp_; }
  inline int output_channels() const { return nout_; }

  // True once a whole block of `fragm` frames is buffered.
  bool is_input_buffer_complete() const { return fragm_ == filled_; }

  // Processed frames not yet written (non-zero after a gapless hand-over).
  int pending_writes() const {
    return drained_ >= 0 ? fragm_ - drained_ : 0;
  }

  // Write `sample_count` processed frames, processing the block first if needed.
  void WriteProcessed(SNDFILE *out, int sample_count);

  // Reset processor for re-use: state identical to a freshly created one.
  void Reset();
  // The name BASELINE.json's north star uses for the same call (this checkout of folve spells it Reset()).
  void ResetBuffer() { Reset(); }

  // Largest (signed) output sample observed (>= 0.0).
  float max_output_value() const { return peak_seen_; }
  void ResetMaxValues();

  const std::string &config_file() const { return config_file_; }
  time_t config_file_timestamp() const { return config_file_timestamp_; }
  bool ConfigStillUpToDate() const;

  // --- additions (not in the reference) ---
  // Placement of new processors on the GPUs of the box.  folve is ONE process; independent
  // files / gapless album chains are spread over every usable GPU with no exchange between
  // them (a processor, once made, carries its device with it through every gapless hand-off).
  //   SetDevice(d >= 0) or $FOLVE_B200_DEVICE=d : every new processor on device d;
  //   SetDevice(kAnyDevice), the default        : Create() picks the device that has the
  //        fewest live processors, or -- when the calling thread has announced a placement
  //        key (the album directory) -- DeviceForKey(key), a pure function of the key.
  enum { kAnyDevice = -1 };
  static void SetDevice(int device);
  static int Device();          // the fixed device, or kAnyDevice
  static int DeviceCount();     // usable sm_100 devices (0 if none)
  // CRC-32 of the key modulo the device count: the same album always lands on the same GPU
  // (folve_b200/sharding.py computes the same function for the multi-process benchmark).
  static int DeviceForKey(const std::string &key, int ndevices);
  // Placement key of the processors this THREAD creates next ("" = none).
  static void SetPlacementKey(const std::string &key);
  static SoundProcessor *CreateOnDevice(const std::string &config_file, int samplerate, int channels,
                                        int device);
  int device() const { return device_; }
  static int LiveProcessors(int device);
  // Block size and the engine handles, for the batched submit layer.
  int fragment_size() const { return fragm_; }
  fcv_stream *stream() const { return stream_; }
  // Drops every cached filter (spectra in HBM) that no processor uses any more.
  static void PurgeFilterCache();

private:
  typedef std::shared_ptr<const std::vector<folve_b200::FileStamp> > Stamps;
  SoundProcessor(fcv_filter *filter, fcv_stream *stream, int fragm, int ninp,
                 int nout, const std::string &cfg_file, time_t cfg_mtime, int device,
                 const Stamps &stamps);
  void Process();

  fcv_filter *const filter_;
  fcv_stream *const stream_;
  const int fragm_, ninp_, nout_;
  const std::string config_file_;
  const time_t config_file_timestamp_;
  const int device_;
  // config file first, then every /impulse/read file: what ConfigStillUpToDate() re-checks
  const Stamps stamps_;

  float *const buffer_;  // pinned, owned by stream_; fragm * max(ninp, nout) floats
  int filled_;
  int drained_;  // frames of the processed block handed out so far; -1: block not processed yet
  float peak_seen_;
};

#endif  // FOLVE_B200_SOUND_PROCESSOR_H
