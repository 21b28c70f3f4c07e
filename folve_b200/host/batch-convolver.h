// -*- c++ -*-
// batch-convolver.h -- the batched submit layer: many files / gapless album chains
// convolved at once on one GPU.
//
// Role in the reference: what BufferThread + ConversionBuffer::FillUntil +
// ConvolveFileHandler::AddMoreSoundData do one file and ~one block per scheduler
// turn (buffer-thread.cc:73-105, conversion-buffer.cc:151-163,
// convolve-file-handler.cc:370-424) -- here every active chain contributes one
// block per step and all blocks go through ONE fcv_batch launch sequence.
//
// With blocks_per_step = T > 1 every chain contributes up to T consecutive blocks per step
// (the time-tiled MAC then reads each input spectrum once for all T outputs that need it);
// a step of a chain ends early wherever the per-file path would continue from a reset
// processor.
//
// Results are those of the per-file SoundProcessor path, including the gapless
// hand-off rules of PassoverProcessor (convolve-file-handler.cc:328-351) and
// their corner cases:
//   * a file that ends inside a block is topped up ONCE from the alphabetically
//     next file; the block's output is split between the two;
//   * a file whose length is a multiple of the block size hands nothing over:
//     its successor starts from a reset state;
//   * a successor that is swallowed whole by the top-up produces no output and
//     ends the hand-off; the file after it starts from a reset state.
#ifndef FOLVE_B200_BATCH_CONVOLVER_H
#define FOLVE_B200_BATCH_CONVOLVER_H

#include <sndfile.h>

#include <string>
#include <vector>

struct fcv_filter;
struct fcv_batch;

namespace folve_b200 {

struct ChainFile {
    SNDFILE *in = nullptr;   // borrowed; `frames` frames of `channels` channels
    SNDFILE *out = nullptr;  // borrowed; receives exactly the frames this file would get from folve
    long frames = 0;
    // results
    long written = 0;
    bool in_gapless = false, out_gapless = false;
    float max_value = 0.0f;  // SoundProcessor::max_output_value() at the moment the file was finished
};

typedef std::vector<ChainFile> Chain;  // files of one directory, alphabetical order

class BatchConvolver {
public:
    // `slots` chains are in flight at once.  NULL on configuration or GPU failure.
    // blocks_per_step: 1, 2, 4 or 8 consecutive blocks of every chain per GPU step.
    // pcm16: every file of every chain is 16-bit PCM in and out (what folve serves for 16-bit FLAC):
    // samples cross the link as int16 (sf_readf_short / sf_writef_short) with libsndfile's
    // int <-> float conversions done by the FFT kernels -- half the bytes of the float path, the same
    // samples in the files.
    static BatchConvolver *Create(const std::string &config_file, int samplerate, int channels, int slots,
                                  bool gapless, int device, int blocks_per_step = 1, bool pcm16 = false);
    ~BatchConvolver();

    int fragment_size() const { return fragm_; }
    int input_channels() const { return ninp_; }
    int output_channels() const { return nout_; }

    // Convolves every chain (pointers must stay valid until the call returns).
    // `threads` host threads move PCM between the SNDFILEs and the pinned staging.
    bool Run(const std::vector<Chain *> &chains, int threads);

    long blocks_processed() const { return blocks_; }
    long steps() const { return steps_; }

private:
    BatchConvolver() {}
    struct Slot;
    struct BlockPlan;
    struct StepPlan;
    class Workers;
    void FillBlock(Slot &s, BlockPlan &b, void *in_block);
    void FillSlot(Slot &s, void *in_step);
    void DrainSlot(Slot &s, const StepPlan &sp, const void *out_step, const float *block_max);
    long ReadFrames(SNDFILE *in, void *dst, long frame_offset, long frames) const;

    fcv_filter *filter_ = nullptr;
    fcv_batch *batch_ = nullptr;
    int fragm_ = 0, ninp_ = 0, nout_ = 0, slots_ = 0, tblocks_ = 1;
    bool gapless_ = true;
    bool pcm16_ = false;
    size_t sample_bytes_ = sizeof(float);   // of the wire format
    long blocks_ = 0, steps_ = 0;
};

// One process, every GPU of the box (folve is a single process): one
// BatchConvolver per device, chains placed by their album key -- SoundProcessor::DeviceForKey, the
// function folve_b200/sharding.py computes for the multi-process benchmark -- or round robin when
// no keys are given, every device driven by its own host thread and its own pool of file
// threads.  No data ever crosses GPUs: a chain lives and dies on one device, the filter spectra
// are replicated (a few MB).  Results are those of a single BatchConvolver, chain for chain.
class MultiDeviceConvolver {
public:
    // devices: CUDA device indices to use (empty: all usable ones).  NULL on failure.
    // instances_per_device: BatchConvolvers per GPU, the chains of a device dealt out among them.  Each keeps two
    // steps in flight, so k instances keep 2k: worth it for SMALL batches (a library of ~100 albums on one GPU),
    // whose steps are bound by the latency of copy in -> kernels -> copy out rather than by the link (measured
    // on the 128-album library of BASELINE config 5: +30 % with two instances).  slots_per_device is the total.
    static MultiDeviceConvolver *Create(const std::string &config_file, int samplerate, int channels,
                                        int slots_per_device, bool gapless, const std::vector<int> &devices,
                                        int blocks_per_step = 1, bool pcm16 = false, int instances_per_device = 1);
    ~MultiDeviceConvolver();

    int devices() const { return (int)device_ids_.size(); }
    int instances_per_device() const { return inst_; }
    int fragment_size() const;
    int output_channels() const;
    // Position (0 .. devices()-1) of the device a chain with this key is placed on.
    int PlacementOf(const std::string &key) const;

    // keys: one placement key per chain (the album directory), or empty for round robin.
    // assignment (optional): receives, per chain, the position of the device it ran on.
    bool Run(const std::vector<Chain *> &chains, const std::vector<std::string> &keys, int threads_per_device,
             std::vector<int> *assignment = nullptr);

    long blocks_processed() const;

private:
    MultiDeviceConvolver() {}
    std::vector<BatchConvolver *> parts_;   // [device position][instance]
    std::vector<int> device_ids_;
    int inst_ = 1;
};

}  // namespace folve_b200

#endif
