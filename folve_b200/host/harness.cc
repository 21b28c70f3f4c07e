// harness.cc -- C entry points that drive SoundProcessor / ProcessorPool the way
// folve's ConvolveFileHandler does, over in-memory PCM instead of FLAC files.
//
// The same source is compiled twice:
//   * against this directory's sound-processor.h / processor-pool.h (the B200
//     product)                                   -> folve_b200/libfolve_host.so
//   * with -DFOLVE_HARNESS_REFERENCE=1 against the reference's own headers and
//     unmodified sources in /root/reference      -> oracle/_ref/libfolve_ref.so
// so the parity tests and the benchmark run literally the same caller code on
// both.  The caller protocol restated here is
// ConvolveFileHandler::AddMoreSoundData (convolve-file-handler.cc:370-424) and
// ConvolveFileHandler::PassoverProcessor (convolve-file-handler.cc:328-351).
#include <math.h>
#include <sndfile.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <pthread.h>
#include <time.h>

#include <string>
#include <vector>

// Angle brackets on purpose: these must come from the -I path of the build
// (the reference's directory or this one), not from the directory of this file.
#include <processor-pool.h>
#include <sound-processor.h>

#if FOLVE_HARNESS_REFERENCE
#include <zita-config.h>
#else
#include <batch-convolver.h>
#include <filter-config.h>
#include <folve_b200.h>
#endif

#if FOLVE_HARNESS_REFERENCE
extern "C" {
zo_shim_impdata_hook zo_shim_on_impdata = 0;
zo_shim_link_hook zo_shim_on_link = 0;
void *zo_shim_hook_user = 0;
int zo_shim_reset_is_fresh = 0;
}
#endif

namespace {

struct FileJob {
    SNDFILE *in = nullptr;
    SNDFILE *out = nullptr;
    long frames_total = 0;
    long frames_left = 0;
    SoundProcessor *processor = nullptr;
    bool opened = false, closed = false;
    bool in_gapless = false, out_gapless = false;
    float max_value = 0.0f;
    bool HasStarted() const { return frames_total != frames_left; }
};

struct Chain {
    ProcessorPool *pool;
    std::string filter_dir;
    int samplerate, channels, bits;
    bool gapless;
    std::vector<FileJob> files;
    std::string error;

    bool Open(FileJob &f) {
        if (f.opened) return f.processor != nullptr;
        f.opened = true;
        std::string err;
        f.processor = pool->GetOrCreate(filter_dir, samplerate, channels, bits, &err);
        if (!f.processor) error = err;
        return f.processor != nullptr;
    }

    void Close(FileJob &f) {
        if (f.closed) return;
        f.closed = true;
        if (f.processor) {
            f.max_value = f.processor->max_output_value();
            pool->Return(f.processor);
            f.processor = nullptr;
        }
    }

    // convolve-file-handler.cc:328-351
    bool Passover(FileJob &b, SoundProcessor *passover) {
        if (b.HasStarted()) return false;
        if (!b.processor) return false;
        if (passover->config_file() != b.processor->config_file() ||
            passover->config_file_timestamp() != b.processor->config_file_timestamp())
            return false;
        pool->Return(b.processor);
        b.processor = passover;
        if (!b.processor->is_input_buffer_complete()) {
            // fill with our beginning so that the donor can finish its block
            b.frames_left -= b.processor->FillBuffer(b.in);
        }
        b.in_gapless = true;
        return true;
    }

    // convolve-file-handler.cc:370-424
    bool AddMoreSoundData(size_t k) {
        FileJob &a = files[k];
        if (!a.frames_left) return false;
        if (a.processor->pending_writes() > 0) {
            a.processor->WriteProcessed(a.out, a.processor->pending_writes());
            return a.frames_left;
        }
        const int r = a.processor->FillBuffer(a.in);
        if (r == 0) {  // premature EOF
            a.frames_left = 0;
            Close(a);
            return false;
        }
        a.frames_left -= r;
        if (!a.frames_left && !a.processor->is_input_buffer_complete() && gapless) {
            FileJob *next = (k + 1 < files.size()) ? &files[k + 1] : nullptr;
            const bool passed = next && Open(*next) && Passover(*next, a.processor);
            a.processor->WriteProcessed(a.out, r);
            if (passed) {
                a.out_gapless = true;
                a.max_value = a.processor->max_output_value();
                a.processor = nullptr;  // ownership handed over
                Close(a);
            }
        } else {
            a.processor->WriteProcessed(a.out, r);
        }
        if (a.frames_left == 0) Close(a);
        return a.frames_left;
    }
};

// largest block size any configuration can have (Convproc::MAXQUANT)
int Convproc_MAXQUANT_for_probe() { return 8192; }

size_t SampleBytes(int subformat) {
    return subformat == SF_FORMAT_PCM_16 ? 2 : 4;
}

ProcessorPool *g_pool = nullptr;

}  // namespace

extern "C" {

const char *fh_version(void) {
#if FOLVE_HARNESS_REFERENCE
    return "reference";
#else
    return "b200";
#endif
}

// Reset()-equals-fresh switch of the restated Convproc (reference build only;
// see oracle/zita_oracle.h).  No effect on the product, whose Reset is always fresh.
void fh_set_reset_is_fresh(int on) {
#if FOLVE_HARNESS_REFERENCE
    zo_shim_reset_is_fresh = on;
#else
    (void)on;
#endif
}

// Drop the process-wide processor pool (idle processors are deleted).
void fh_drop_pool(void) {
#if !FOLVE_HARNESS_REFERENCE
    delete g_pool;  // the reference's ProcessorPool has no destructor; it just leaks
#endif
    g_pool = nullptr;
}

// Runs `nfiles` in-memory files (alphabetical order == array order) through the
// caller protocol.  pcm[i]: interleaved `channels`-channel samples in
// `in_subformat` (SF_FORMAT_PCM_16: int16, PCM_24/PCM_32: int32, FLOAT: float).
// out_pcm[i] receives frames[i] * nout samples in `out_subformat`.
// Returns the number of output channels, or <0 with a message in errbuf.
int fh_run_chain(const char *filter_dir, int samplerate, int channels, int bits, int gapless, int nfiles,
                 const void *const *pcm, const long *frames, int in_subformat, int out_subformat,
                 void *const *out_pcm, long *out_frames, float *max_values, int *gapless_flags, char *errbuf,
                 int errcap) {
    if (!g_pool) g_pool = new ProcessorPool(3);  // folve-filesystem.cc:50
    Chain chain;
    chain.pool = g_pool;
    chain.filter_dir = filter_dir;
    chain.samplerate = samplerate;
    chain.channels = channels;
    chain.bits = bits;
    chain.gapless = gapless != 0;
    chain.files.resize((size_t)nfiles);
    int nout = -1;
    int rc = 0;
    for (int i = 0; i < nfiles; i++) {
        FileJob &f = chain.files[(size_t)i];
        f.in = sf_shim_open_memory_read(pcm[i], frames[i], channels, samplerate, in_subformat);
        f.frames_total = f.frames_left = frames[i];
    }
    for (int i = 0; i < nfiles && rc == 0; i++) {
        FileJob &f = chain.files[(size_t)i];
        if (!chain.Open(f)) {
            if (errbuf && errcap > 0) snprintf(errbuf, (size_t)errcap, "%s", chain.error.c_str());
            rc = -1;
            break;
        }
        if (f.processor) {
            if (nout < 0) nout = f.processor->output_channels();
        }
        // the output file is created lazily, once the channel count is known
        for (int j = i; j < nfiles && j <= i + 1; j++) {
            FileJob &g = chain.files[(size_t)j];
            if (!g.out) g.out = sf_shim_open_memory_write(nout, samplerate, out_subformat);
        }
        while (chain.AddMoreSoundData((size_t)i)) {}
        chain.Close(f);
    }
    for (int i = 0; i < nfiles; i++) {
        FileJob &f = chain.files[(size_t)i];
        chain.Close(f);
        const long n = f.out ? (long)sf_shim_memory_frames(f.out) : 0;
        if (out_frames) out_frames[i] = n;
        if (f.out && out_pcm && out_pcm[i] && n > 0)
            memcpy(out_pcm[i], sf_shim_memory_data(f.out), (size_t)n * (size_t)nout * SampleBytes(out_subformat));
        if (max_values) max_values[i] = f.max_value;
        if (gapless_flags) gapless_flags[i] = (f.in_gapless ? 1 : 0) | (f.out_gapless ? 2 : 0);
        if (f.in) sf_close(f.in);
        if (f.out) sf_close(f.out);
    }
    return rc ? rc : nout;
}

#if !FOLVE_HARNESS_REFERENCE
// ---- the batched submit layer ----------------------------------------------------
// Runs a whole library -- file i belongs to chain chain_of_file[i] (non-decreasing),
// files of a chain in alphabetical order -- through folve_b200::BatchConvolver with
// `slots` chains in flight.  PCM in/out is float, [frames][channels].
// Returns the number of output channels, <0 on failure.
// pcm16 != 0: PCM in/out is int16 (16-bit files, int16 on the wire) instead of float.
// Test hooks consumed by the next RunLibrary call: the length every file CLAIMS to have (a file
// that is shorter than its header says: truncated input), and the number of GPUs to spread the
// chains over from this one process (0 = the single BatchConvolver on the default device).
static std::vector<long> g_claimed_frames;
static int g_library_devices = 0;
static int g_library_instances = 1;
static std::vector<int> g_last_assignment;

static int RunLibrary(const char *config_file, int samplerate, int channels, int gapless, int slots, int threads,
                      int nfiles, const int *chain_of_file, const void *const *pcm, const long *frames,
                      void *const *out_pcm, long *out_frames, float *max_values, int *gapless_flags,
                      long *steps_out, int blocks_per_step, int pcm16) {
    const std::vector<long> claimed = g_claimed_frames;
    g_claimed_frames.clear();
    const int ndev = g_library_devices, inst = g_library_instances;
    g_library_devices = 0;
    g_library_instances = 1;
    folve_b200::BatchConvolver *bc = nullptr;
    folve_b200::MultiDeviceConvolver *md = nullptr;
    if (ndev > 0 || inst > 1) {
        std::vector<int> ids;
        for (int d = 0; d < (ndev > 0 ? ndev : 1); d++)
            ids.push_back(ndev > 0 ? d : (SoundProcessor::Device() < 0 ? 0 : SoundProcessor::Device()));
        md = folve_b200::MultiDeviceConvolver::Create(config_file, samplerate, channels, slots + inst, gapless != 0, ids,
                                                      blocks_per_step, pcm16 != 0, inst);
        if (!md) return -1;
    } else {
        bc = folve_b200::BatchConvolver::Create(config_file, samplerate, channels, slots, gapless != 0,
                                                (SoundProcessor::Device() < 0 ? 0 : SoundProcessor::Device()),
                                                blocks_per_step, pcm16 != 0);
        if (!bc) return -1;
    }
    const int nout = bc ? bc->output_channels() : md->output_channels();
    std::vector<folve_b200::Chain> chains;
    for (int i = 0; i < nfiles; i++) {
        if (chains.empty() || (i > 0 && chain_of_file[i] != chain_of_file[i - 1])) chains.emplace_back();
        folve_b200::ChainFile f;
        const int fmt = pcm16 ? SF_FORMAT_PCM_16 : SF_FORMAT_FLOAT;
        f.in = sf_shim_open_memory_read(pcm[i], frames[i], channels, samplerate, fmt);
        f.out = sf_shim_open_memory_write(nout, samplerate, fmt);
        f.frames = (size_t)i < claimed.size() ? claimed[(size_t)i] : frames[i];
        chains.back().push_back(f);
    }
    std::vector<folve_b200::Chain *> ptrs;
    for (auto &c : chains) ptrs.push_back(&c);
    bool ok;
    if (md) {
        // placement key of chain c: "album<c>" -- the directory name a folve mount would see
        std::vector<std::string> keys;
        for (size_t c = 0; c < chains.size(); c++) keys.push_back("album" + std::to_string(c));
        ok = md->Run(ptrs, keys, threads, &g_last_assignment);
    } else {
        ok = bc->Run(ptrs, threads);
    }
    int i = 0;
    for (auto &c : chains)
        for (auto &f : c) {
            const long n = (long)sf_shim_memory_frames(f.out);
            out_frames[i] = n;
            if (n > 0) memcpy(out_pcm[i], sf_shim_memory_data(f.out), (size_t)n * (size_t)nout * (pcm16 ? sizeof(short) : sizeof(float)));
            max_values[i] = f.max_value;
            gapless_flags[i] = (f.in_gapless ? 1 : 0) | (f.out_gapless ? 2 : 0);
            sf_close(f.in);
            sf_close(f.out);
            i++;
        }
    if (steps_out) *steps_out = bc ? bc->steps() : 0;
    delete bc;
    delete md;
    return ok ? nout : -2;
}

void fh_set_claimed_frames(const long *claimed, int n) { g_claimed_frames.assign(claimed, claimed + n); }
void fh_set_library_devices(int ndevices) { g_library_devices = ndevices; }
void fh_set_library_instances(int instances) { g_library_instances = instances > 1 ? instances : 1; }
// device position every chain of the last multi-device run was placed on; returns the chain count
int fh_last_assignment(int *out, int capacity) {
    for (int i = 0; i < capacity && (size_t)i < g_last_assignment.size(); i++) out[i] = g_last_assignment[(size_t)i];
    return (int)g_last_assignment.size();
}
int fh_device_for_key(const char *key, int ndevices) { return SoundProcessor::DeviceForKey(key, ndevices); }
// Creates `n` processors for one config through SoundProcessor::Create (device chosen by load)
// and reports where they live; all are deleted again.  Returns n, <0 on failure.
int fh_processor_devices(const char *config_file, int samplerate, int channels, int *devices, int n) {
    std::vector<SoundProcessor *> ps;
    int rc = n;
    for (int i = 0; i < n; i++) {
        SoundProcessor *p = SoundProcessor::Create(config_file, samplerate, channels);
        if (!p) { rc = -1; break; }
        devices[i] = p->device();
        ps.push_back(p);
    }
    for (SoundProcessor *p : ps) delete p;
    return rc;
}

int fh_run_library_tiled(const char *config_file, int samplerate, int channels, int gapless, int slots, int threads,
                         int nfiles, const int *chain_of_file, const float *const *pcm, const long *frames,
                         float *const *out_pcm, long *out_frames, float *max_values, int *gapless_flags,
                         long *steps_out, int blocks_per_step) {
    return RunLibrary(config_file, samplerate, channels, gapless, slots, threads, nfiles, chain_of_file,
                      (const void *const *)pcm, frames, (void *const *)out_pcm, out_frames, max_values, gapless_flags,
                      steps_out, blocks_per_step, 0);
}

int fh_run_library_pcm16(const char *config_file, int samplerate, int channels, int gapless, int slots, int threads,
                         int nfiles, const int *chain_of_file, const short *const *pcm, const long *frames,
                         short *const *out_pcm, long *out_frames, float *max_values, int *gapless_flags,
                         long *steps_out, int blocks_per_step) {
    return RunLibrary(config_file, samplerate, channels, gapless, slots, threads, nfiles, chain_of_file,
                      (const void *const *)pcm, frames, (void *const *)out_pcm, out_frames, max_values, gapless_flags,
                      steps_out, blocks_per_step, 1);
}

int fh_run_library(const char *config_file, int samplerate, int channels, int gapless, int slots, int threads,
                   int nfiles, const int *chain_of_file, const float *const *pcm, const long *frames,
                   float *const *out_pcm, long *out_frames, float *max_values, int *gapless_flags, long *steps_out) {
    return fh_run_library_tiled(config_file, samplerate, channels, gapless, slots, threads, nfiles, chain_of_file, pcm,
                                frames, out_pcm, out_frames, max_values, gapless_flags, steps_out, 1);
}
static double NowSeconds() {
    timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

// ---- throughput of the batched submit layer --------------------------------------
// `nchains` album chains of `files_per_chain` in-memory float files each (lengths around
// `frames_per_file`, never a multiple of the block size, all reading the same white-noise
// buffer), every chain in flight at once, output discarded.  Returns the wall seconds of
// BatchConvolver::Run, <0 on error; *audio_seconds = what was convolved.
double fh_bench_library(const char *config_file, int samplerate, int channels, int gapless, int nchains,
                        int files_per_chain, long frames_per_file, int blocks_per_step, int threads, int pcm16,
                        double *audio_seconds) {
    // FOLVE_B200_LIBRARY_INSTANCES=k: k BatchConvolvers on the GPU (MultiDeviceConvolver's instances_per_device)
    const int inst = getenv("FOLVE_B200_LIBRARY_INSTANCES") ? atoi(getenv("FOLVE_B200_LIBRARY_INSTANCES")) : 1;
    const int dev = SoundProcessor::Device() < 0 ? 0 : SoundProcessor::Device();
    folve_b200::BatchConvolver *bc = nullptr;
    folve_b200::MultiDeviceConvolver *md = nullptr;
    if (inst > 1)
        md = folve_b200::MultiDeviceConvolver::Create(config_file, samplerate, channels, nchains + inst, gapless != 0,
                                                      std::vector<int>(1, dev), blocks_per_step, pcm16 != 0, inst);
    else
        bc = folve_b200::BatchConvolver::Create(config_file, samplerate, channels, nchains, gapless != 0, dev,
                                                blocks_per_step, pcm16 != 0);
    if (!bc && !md) return -1.0;
    const int nout = bc ? bc->output_channels() : md->output_channels();
    const long longest = frames_per_file + 4099;
    std::vector<float> pcm((size_t)longest * (size_t)channels);
    std::vector<short> pcm_s16(pcm16 ? pcm.size() : 0);
    uint32_t s = 12345u;
    for (size_t i = 0; i < pcm.size(); i++) {
        s = s * 1664525u + 1013904223u;
        pcm[i] = 0.03f * ((float)(s >> 8) * (1.0f / 8388608.0f) - 1.0f);
        if (pcm16) pcm_s16[i] = (short)lrintf(pcm[i] * 32768.0f);
    }
    const int fmt = pcm16 ? SF_FORMAT_PCM_16 : SF_FORMAT_FLOAT;
    const void *src = pcm16 ? (const void *)pcm_s16.data() : (const void *)pcm.data();
    std::vector<folve_b200::Chain> chains((size_t)nchains);
    double frames_total = 0.0;
    for (int c = 0; c < nchains; c++)
        for (int k = 0; k < files_per_chain; k++) {
            folve_b200::ChainFile f;
            f.frames = frames_per_file + 1 + (long)((c * 131 + k * 977) % 4097);
            f.in = sf_shim_open_memory_read(src, f.frames, channels, samplerate, fmt);
            f.out = sf_shim_open_null_write(nout, samplerate, fmt);
            frames_total += (double)f.frames;
            chains[(size_t)c].push_back(f);
        }
    std::vector<folve_b200::Chain *> ptrs;
    for (auto &c : chains) ptrs.push_back(&c);
    const double t0 = NowSeconds();
    const bool ok = bc ? bc->Run(ptrs, threads) : md->Run(ptrs, std::vector<std::string>(), threads);
    const double wall = NowSeconds() - t0;
    for (auto &c : chains)
        for (auto &f : c) {
            sf_close(f.in);
            sf_close(f.out);
        }
    delete bc;
    delete md;
    if (audio_seconds) *audio_seconds = frames_total / (double)samplerate;
    return ok ? wall : -2.0;
}

// ---- BASELINE config 5: a library of gapless albums, prebuffered all at once ------------
// `nalbums` albums of `tracks` tracks; track lengths uniform in [120 s, 360 s) from an LCG seeded
// with 100 + album (never a multiple of the block size); every album is ONE gapless chain through
// one processor (convolve-file-handler.cc:390-415).  This process takes the albums
// a = rank, rank + world, ... (folve_b200/sharding.py balanced_albums) and, when ndevices > 1,
// spreads them over that many GPUs itself (MultiDeviceConvolver).  All chains of the shard are
// in flight at once ("library prebuffer"); PCM comes from one shared white-noise buffer, output is
// discarded.  Returns the wall seconds of Run, <0 on error; *audio_seconds = what was convolved.
double fh_bench_albums(const char *config_file, int samplerate, int channels, int nalbums, int tracks, int rank,
                       int world, int ndevices, int blocks_per_step, int threads, int pcm16, double *audio_seconds,
                       int *chains_out) {
    std::vector<int> mine;
    for (int a = rank; a < nalbums; a += world) mine.push_back(a);
    if (mine.empty()) return -1.0;
    // FOLVE_B200_LIBRARY_INSTANCES=k: k BatchConvolvers per GPU (MultiDeviceConvolver's instances_per_device)
    const int inst = getenv("FOLVE_B200_LIBRARY_INSTANCES") ? atoi(getenv("FOLVE_B200_LIBRARY_INSTANCES")) : 1;
    const int nd = ndevices > 1 ? ndevices : 1;
    const int slots = ((int)mine.size() + nd - 1) / nd + (nd > 1 ? 2 : 0) + (inst > 1 ? inst : 0);   // room for an uneven placement
    folve_b200::BatchConvolver *bc = nullptr;
    folve_b200::MultiDeviceConvolver *md = nullptr;
    if (nd > 1 || inst > 1) {
        std::vector<int> ids;
        const int dev0 = SoundProcessor::Device() < 0 ? 0 : SoundProcessor::Device();
        for (int d = 0; d < nd; d++) ids.push_back(nd > 1 ? d : dev0);
        md = folve_b200::MultiDeviceConvolver::Create(config_file, samplerate, channels, slots, true, ids,
                                                      blocks_per_step, pcm16 != 0, inst);
    } else {
        bc = folve_b200::BatchConvolver::Create(config_file, samplerate, channels, slots, true,
                                                (SoundProcessor::Device() < 0 ? 0 : SoundProcessor::Device()),
                                                blocks_per_step, pcm16 != 0);
    }
    if (!bc && !md) return -1.0;
    const int nout = bc ? bc->output_channels() : md->output_channels();
    const int fragm = bc ? bc->fragment_size() : md->fragment_size();
    const long longest = 360L * samplerate + 2;
    std::vector<float> pcm(pcm16 ? 0 : (size_t)longest * (size_t)channels);
    std::vector<short> pcm_s16(pcm16 ? (size_t)longest * (size_t)channels : 0);
    uint32_t s = 12345u;
    for (size_t i = 0; i < (size_t)longest * (size_t)channels; i++) {
        s = s * 1664525u + 1013904223u;
        const float v = 0.03f * ((float)(s >> 8) * (1.0f / 8388608.0f) - 1.0f);
        if (pcm16) pcm_s16[i] = (short)lrintf(v * 32768.0f);
        else pcm[i] = v;
    }
    const int fmt = pcm16 ? SF_FORMAT_PCM_16 : SF_FORMAT_FLOAT;
    const void *src = pcm16 ? (const void *)pcm_s16.data() : (const void *)pcm.data();
    std::vector<folve_b200::Chain> chains(mine.size());
    std::vector<std::string> keys;
    double frames_total = 0.0;
    for (size_t c = 0; c < mine.size(); c++) {
        uint32_t r = 100u + (uint32_t)mine[c];
        for (int k = 0; k < tracks; k++) {
            r = r * 1664525u + 1013904223u;
            folve_b200::ChainFile f;
            f.frames = (long)(120.0 * samplerate + (double)(r >> 8) * (1.0 / 16777216.0) * 240.0 * samplerate);
            if (f.frames % fragm == 0) f.frames++;
            f.in = sf_shim_open_memory_read(src, f.frames, channels, samplerate, fmt);
            f.out = sf_shim_open_null_write(nout, samplerate, fmt);
            frames_total += (double)f.frames;
            chains[c].push_back(f);
        }
        keys.push_back("album" + std::to_string(mine[c]));
    }
    std::vector<folve_b200::Chain *> ptrs;
    for (auto &c : chains) ptrs.push_back(&c);
    const double t0 = NowSeconds();
    const bool ok = md ? md->Run(ptrs, std::vector<std::string>(), threads) : bc->Run(ptrs, threads);
    const double wall = NowSeconds() - t0;
    for (auto &c : chains)
        for (auto &f : c) {
            sf_close(f.in);
            sf_close(f.out);
        }
    delete bc;
    delete md;
    if (audio_seconds) *audio_seconds = frames_total / (double)samplerate;
    if (chains_out) *chains_out = (int)mine.size();
    return ok ? wall : -2.0;
}

#endif

// ---- throughput of the synchronous drop-in API ---------------------------------
// `nthreads` independent files, one SoundProcessor each (from the pool, i.e. via
// SoundProcessor::Create and the filter-config parser), one file per thread --
// the way folve uses its cores (README.md:361-362).  Every thread runs the
// FillBuffer / WriteProcessed block loop of AddMoreSoundData for `nblocks` full
// blocks on in-memory float PCM (a 16-block loop of white noise, re-read by
// seeking back) into a discarding sink.  Returns the wall seconds of the block
// loops (processor creation excluded), <0 on error.
struct BenchJob {
    ProcessorPool *pool;
    const char *dir;
    int rate, channels, bits, nblocks, index;
    pthread_barrier_t *start, *stop;
    int ok;
    int fragm;
    float max_value;
};

static void *BenchWorker(void *arg) {
    BenchJob *j = (BenchJob *)arg;
    std::string err;
    SoundProcessor *p = j->pool->GetOrCreate(j->dir, j->rate, j->channels, j->bits, &err);
    j->ok = p != nullptr;
    std::vector<float> pcm;
    SNDFILE *in = nullptr, *out = nullptr;
    const int loop_blocks = 16;
    int fragm = 0;
    if (p) {
        // a processor that just came out of Create has an empty block: its size is
        // what one FillBuffer on a long enough file returns
        std::vector<float> probe((size_t)Convproc_MAXQUANT_for_probe() * (size_t)j->channels, 0.0f);
        SNDFILE *ps = sf_shim_open_memory_read(probe.data(), (sf_count_t)(probe.size() / (size_t)j->channels),
                                               j->channels, j->rate, SF_FORMAT_FLOAT);
        fragm = p->FillBuffer(ps);
        sf_close(ps);
        p->Reset();
        j->fragm = fragm;
        pcm.resize((size_t)fragm * (size_t)loop_blocks * (size_t)j->channels);
        uint32_t s = (uint32_t)(j->index + 1) * 2654435761u + 12345u;
        for (size_t i = 0; i < pcm.size(); i++) {
            s = s * 1664525u + 1013904223u;
            pcm[i] = 0.03f * ((float)(s >> 8) * (1.0f / 8388608.0f) - 1.0f);
        }
        in = sf_shim_open_memory_read(pcm.data(), (sf_count_t)fragm * loop_blocks, j->channels, j->rate, SF_FORMAT_FLOAT);
        out = sf_shim_open_null_write(p->output_channels(), j->rate, SF_FORMAT_FLOAT);
    }
    pthread_barrier_wait(j->start);
    if (p) {
        for (int b = 0; b < j->nblocks; b++) {
            if (b % loop_blocks == 0) sf_seek(in, 0, SEEK_SET);
            const int r = p->FillBuffer(in);
            p->WriteProcessed(out, r);
        }
        j->max_value = p->max_output_value();
    }
    pthread_barrier_wait(j->stop);
    if (in) sf_close(in);
    if (out) sf_close(out);
    if (p) j->pool->Return(p);
    return nullptr;
}

double fh_bench_threads(const char *filter_dir, int samplerate, int channels, int bits, int nthreads, int nblocks,
                        int *fragm_out) {
    if (nthreads < 1) return -1.0;
    ProcessorPool pool(0);  // nothing is kept: every thread creates and finally deletes its processor
    std::vector<pthread_t> th((size_t)nthreads);
    std::vector<BenchJob> jobs((size_t)nthreads);
    pthread_barrier_t start, stop;
    pthread_barrier_init(&start, nullptr, (unsigned)nthreads + 1);
    pthread_barrier_init(&stop, nullptr, (unsigned)nthreads + 1);
    for (int t = 0; t < nthreads; t++) {
        jobs[(size_t)t] = BenchJob{&pool, filter_dir, samplerate, channels, bits, nblocks, t, &start, &stop, 0, 0, 0.f};
        pthread_create(&th[(size_t)t], nullptr, BenchWorker, &jobs[(size_t)t]);
    }
    pthread_barrier_wait(&start);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    pthread_barrier_wait(&stop);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    int ok = 1;
    for (int t = 0; t < nthreads; t++) {
        pthread_join(th[(size_t)t], nullptr);
        ok &= jobs[(size_t)t].ok;
    }
    pthread_barrier_destroy(&start);
    pthread_barrier_destroy(&stop);
    if (fragm_out) *fragm_out = jobs[0].fragm;
    if (!ok) return -1.0;
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

// ---- what a configuration file loads ------------------------------------------
// Parses `config_file` with the library's own loader and reports, per (in,out)
// pair, the accumulated time-domain impulse as it is handed to the convolver
// (scaled by 0.5/fragm, the scaling zita applies in impdata_create).
struct fh_config {
    int rc = 0;
    int created = 0;  // would SoundProcessor::Create succeed
    int ninp = 0, nout = 0, size = 0, fragm = 0, npar = 0;
#if FOLVE_HARNESS_REFERENCE
    struct PairData { bool exists = false; int link = -1; std::vector<float> h; };
    std::vector<PairData> pairs;
#else
    folve_b200::FilterConfig cfg;
#endif
};

#if FOLVE_HARNESS_REFERENCE
static void RefOnImpdata(void *user, unsigned inp, unsigned out, int step, const float *data, int ind0, int ind1) {
    // mirror of fcv_filter_add's accumulation (fcv_engine.cu) on the calls the reference's parser makes
    fh_config *c = (fh_config *)user;
    if ((int)inp >= c->ninp || (int)out >= c->nout) return;
    const long n = (long)ind1 - ind0, total = (long)c->npar * c->fragm, i0 = -(long)ind0;
    if (i0 >= n || i0 + total <= 0) return;
    fh_config::PairData &p = c->pairs[(size_t)inp * c->nout + out];
    p.exists = true;
    if (p.link >= 0 || !data) return;
    if (p.h.empty()) p.h.assign((size_t)total, 0.0f);
    const float norm = 0.5f / (float)c->fragm;
    const long j0 = i0 < 0 ? 0 : i0, j1 = (i0 + total > n) ? n : i0 + total;
    for (long j = j0; j < j1; j++) p.h[(size_t)(j - i0)] += norm * data[j * step];
}
static void RefOnLink(void *user, unsigned inp1, unsigned out1, unsigned inp2, unsigned out2) {
    fh_config *c = (fh_config *)user;
    if ((int)inp1 >= c->ninp || (int)out1 >= c->nout || (int)inp2 >= c->ninp || (int)out2 >= c->nout) return;
    if (inp1 == inp2 && out1 == out2) return;
    const int src = (int)(inp1 * c->nout + out1), dst = (int)(inp2 * c->nout + out2);
    if (!c->pairs[(size_t)src].exists) return;
    c->pairs[(size_t)dst].exists = true;
    c->pairs[(size_t)dst].h.clear();
    c->pairs[(size_t)dst].link = src;
}
#endif

fh_config *fh_config_open(const char *config_file, int samplerate, int channels) {
    fh_config *c = new fh_config();
#if FOLVE_HARNESS_REFERENCE
    // what SoundProcessor::Create does (sound-processor.cc:36-48), with observers
    // on the impdata calls; the pair table is sized once /convolver/new is seen,
    // i.e. lazily inside the first callback.
    ZitaConfig zita;
    memset(&zita, 0, sizeof(zita));
    zita.fsamp = samplerate;
    zita.ninp = channels;
    zita.nout = channels;
    zita.convproc = new Convproc();
    struct Ctx { fh_config *c; ZitaConfig *z; } ctx = {c, &zita};
    zo_shim_hook_user = &ctx;
    zo_shim_on_impdata = [](void *u, unsigned inp, unsigned out, int step, const float *data, int i0, int i1) {
        Ctx *x = (Ctx *)u;
        if (x->c->pairs.empty()) {
            x->c->ninp = x->z->ninp; x->c->nout = x->z->nout; x->c->fragm = x->z->fragm;
            x->c->npar = (x->z->size + x->z->fragm - 1) / x->z->fragm;
            x->c->pairs.resize((size_t)x->c->ninp * x->c->nout);
        }
        RefOnImpdata(x->c, inp, out, step, data, i0, i1);
    };
    zo_shim_on_link = [](void *u, unsigned a, unsigned b, unsigned d, unsigned e) {
        Ctx *x = (Ctx *)u;
        if (x->c->pairs.empty()) return;
        RefOnLink(x->c, a, b, d, e);
    };
    c->rc = config(&zita, config_file);
    zo_shim_on_impdata = 0;
    zo_shim_on_link = 0;
    zo_shim_hook_user = 0;
    c->created = c->rc == 0 && zita.convproc->state() != Convproc::ST_IDLE &&
                 zita.convproc->inpdata(zita.ninp - 1) != NULL && zita.convproc->outdata(zita.nout - 1) != NULL;
    c->ninp = zita.ninp; c->nout = zita.nout; c->size = zita.size; c->fragm = zita.fragm;
    c->npar = zita.fragm ? (zita.size + zita.fragm - 1) / zita.fragm : 0;
    if (c->pairs.empty() && c->created) c->pairs.resize((size_t)c->ninp * c->nout);
    delete zita.convproc;
#else
    c->cfg.fsamp = samplerate;
    c->cfg.ninp = channels;
    c->cfg.nout = channels;
    c->rc = folve_b200::LoadFilterConfig(&c->cfg, config_file);
    c->created = c->rc == 0 && c->cfg.filter != nullptr;
    c->ninp = c->cfg.ninp; c->nout = c->cfg.nout; c->size = c->cfg.size; c->fragm = c->cfg.fragm;
    c->npar = c->cfg.filter ? fcv_filter_partitions(c->cfg.filter) : 0;
#endif
    return c;
}

// out[0..6] = rc, created, ninp, nout, size, fragm, partitions
void fh_config_info(const fh_config *c, int *out) {
    out[0] = c->rc; out[1] = c->created; out[2] = c->ninp; out[3] = c->nout;
    out[4] = c->size; out[5] = c->fragm; out[6] = c->npar;
}

// 0: pair absent, 1: owns data, 2: link (dst gets the source's data); capacity in floats
int fh_config_impulse(fh_config *c, int inp, int out, float *dst, int capacity) {
    if (!c->created || inp < 0 || inp >= c->ninp || out < 0 || out >= c->nout) return -1;
#if FOLVE_HARNESS_REFERENCE
    const fh_config::PairData &p = c->pairs[(size_t)inp * c->nout + out];
    if (!p.exists) return 0;
    const fh_config::PairData *src = p.link >= 0 ? &c->pairs[(size_t)p.link] : &p;
    memset(dst, 0, sizeof(float) * (size_t)capacity);
    if (src->link < 0 && !src->h.empty())
        memcpy(dst, src->h.data(), sizeof(float) * (src->h.size() < (size_t)capacity ? src->h.size() : (size_t)capacity));
    return p.link >= 0 ? 2 : 1;
#else
    return fcv_filter_get_impulse(c->cfg.filter, inp, out, dst, capacity);
#endif
}

// The files LoadFilterConfig stamped for the staleness check (every /impulse/read file that was reached):
// returns their number, *current = 1 while none of them changed (mtime in nanoseconds, size, existence).
// The reference has no such record (sound-processor.cc:129-133, TODO): -1 there.
int fh_config_impulse_stamps(const fh_config *c, int *current) {
#if FOLVE_HARNESS_REFERENCE
    (void)c;
    *current = 1;
    return -1;
#else
    *current = folve_b200::StampsCurrent(c->cfg.impulse_files) ? 1 : 0;
    return (int)c->cfg.impulse_files.size();
#endif
}

void fh_config_close(fh_config *c) {
#if !FOLVE_HARNESS_REFERENCE
    if (c && c->cfg.filter) fcv_filter_unref(c->cfg.filter);
#endif
    delete c;
}

}  // extern "C"
