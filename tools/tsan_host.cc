// tsan_host.cc -- ThreadSanitizer driver for the host-side concurrency of the product: the
// coalescer behind fcv_stream_process (mutex / epoch / launch slots, fcv_engine.cu), SoundProcessor
// + ProcessorPool used from many threads, and BatchConvolver's worker pool with two steps in
// flight.  Built by tools/sanitize.sh with -fsanitize=thread over the host sources AND the host
// side of fcv_engine.cu; run on a GPU box.  Exit code 0 and no "WARNING: ThreadSanitizer" = clean.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include <batch-convolver.h>
#include <processor-pool.h>
#include <sndfile.h>
#include <sound-processor.h>

static std::vector<float> Noise(size_t n, unsigned seed) {
    std::vector<float> v(n);
    unsigned s = seed * 2654435761u + 12345u;
    for (auto &x : v) {
        s = s * 1664525u + 1013904223u;
        x = 0.1f * ((float)(s >> 8) * (1.0f / 8388608.0f) - 1.0f);
    }
    return v;
}

int main() {
    char dir[] = "/tmp/fcv_tsan_XXXXXX";
    if (!mkdtemp(dir)) return 2;
    const std::string conf = std::string(dir) + "/filter-44100.conf";
    FILE *f = fopen(conf.c_str(), "w");
    fprintf(f, "/convolver/new 2 2 256 20000\n/impulse/hilbert 1 1 0.5 9000 4000\n/impulse/dirac 2 2 0.5 12345\n"
               "/impulse/dirac 1 2 0.1 300\n/impulse/hilbert 2 1 0.2 17000 2000\n");
    fclose(f);
    const int rate = 44100, ch = 2;
    std::atomic<int> rc{0};

    {   // 1. the drop-in API from 8 threads: pool, Create, coalesced Process() calls, Return
        ProcessorPool pool(3);
        std::vector<std::thread> th;
        std::vector<float> sums(8, 0.f);
        for (int t = 0; t < 8; t++)
            th.emplace_back([&, t] {
                for (int round = 0; round < 2; round++) {
                    std::string err;
                    SoundProcessor *p = pool.GetOrCreate(dir, rate, ch, 16, &err);
                    if (!p) { rc = 3; return; }
                    const int N = p->fragment_size();
                    std::vector<float> pcm = Noise((size_t)N * 6 * ch + 77 * ch, (unsigned)(t + 1));
                    SNDFILE *in = sf_shim_open_memory_read(pcm.data(), (sf_count_t)(pcm.size() / ch), ch, rate, SF_FORMAT_FLOAT);
                    SNDFILE *out = sf_shim_open_memory_write(p->output_channels(), rate, SF_FORMAT_FLOAT);
                    for (;;) {
                        const int r = p->FillBuffer(in);
                        if (r == 0) break;
                        p->WriteProcessed(out, r);
                        if (p->pending_writes()) p->WriteProcessed(out, 0);
                        if (r < N) break;
                    }
                    sums[t] += p->max_output_value();
                    sf_close(in);
                    sf_close(out);
                    pool.Return(p);
                }
            });
        for (auto &x : th) x.join();
        for (float s : sums) if (!(s > 0.f) && !rc) rc = 4;
    }
    {   // 2. the batched submit layer: 12 chains x 3 files, 6 slots, 4 file threads, 4 blocks per step
        folve_b200::BatchConvolver *bc = folve_b200::BatchConvolver::Create(conf, rate, ch, 6, true, 0, 4, false);
        if (!bc) return 5;
        const int N = bc->fragment_size();
        std::vector<std::vector<float> > pcm;
        std::vector<folve_b200::Chain> chains(12);
        for (int c = 0; c < 12; c++)
            for (int k = 0; k < 3; k++) {
                const long frames = (long)N * (1 + (c + k) % 5) + 17 * c + k;
                pcm.push_back(Noise((size_t)frames * ch, (unsigned)(100 + 3 * c + k)));
                folve_b200::ChainFile cf;
                cf.frames = frames;
                cf.in = sf_shim_open_memory_read(pcm.back().data(), frames, ch, rate, SF_FORMAT_FLOAT);
                cf.out = sf_shim_open_memory_write(bc->output_channels(), rate, SF_FORMAT_FLOAT);
                chains[(size_t)c].push_back(cf);
            }
        std::vector<folve_b200::Chain *> ptrs;
        for (auto &c : chains) ptrs.push_back(&c);
        if (!bc->Run(ptrs, 4) && !rc) rc = 6;
        for (auto &c : chains)
            for (auto &cf : c) {
                if (cf.written != cf.frames && !rc) rc = 7;
                sf_close(cf.in);
                sf_close(cf.out);
            }
        delete bc;
    }
    SoundProcessor::PurgeFilterCache();
    unlink(conf.c_str());
    rmdir(dir);
    printf("tsan_host: rc = %d\n", rc.load());
    return rc;
}
