// fcv_stream_dev.cuh -- how a kernel launch finds the streams it works on.
//
// Every stream (one SoundProcessor, or one slot of a batch) has a StreamDev descriptor in
// device memory.  A launch addresses its streams in one of two ways:
//   BatchSel : the streams of a batch -- consecutive descriptors, per-stream valid-frame counts
//              in a device array (or one count for all), one ring position for all of them
//              (a batch advances in lock-step);
//   GroupSel : up to GROUP_MAX unrelated single streams whose synchronous Process() calls
//              (sound-processor.cc:98-127, one per file and host thread in folve) arrived at the
//              same time and were coalesced into ONE launch sequence: descriptor pointers, frame
//              counts and ring positions travel BY VALUE in the kernel parameters, so a group
//              costs no host->device copy at all.
#pragma once
#include <cuda_runtime.h>

namespace fcv {

// Programmatic dependent launch: a kernel launched with the programmatic-stream-serialization
// attribute may start while its predecessor in the stream is still running; pdl_wait() returns
// once the predecessor has completed and its writes are visible, pdl_trigger() lets the successor
// be scheduled.  Both are no-ops for launches without the attribute (every batched launch).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }

// Per-stream device descriptor (array owned by a batch or a single stream).
struct StreamDev {
    float2 *xring;  // [ninp][R][M] input-spectra ring
    float *tail;    // [nout][N]    overlap tails
    const void *din;  // interleaved PCM in,  [T*N][ninp] wire format
    void *dout;       // interleaved PCM out, [T*N][nout] wire format
    float *maxv;    // running signed maximum
    float *bmax;    // [T] signed maximum of each block of the last step (valid frames only)
    float2 *Y;      // [nout][T][M] accumulated output spectra of the step
    float2 *zc0;    // [nout][T]    entry 0 of the sequences the inverse transform starts from
    // single streams only (else null): the caller's pinned block as the device sees it, the mirror
    // of the block maximum behind it, and the count of output-channel CTAs that have finished
    void *hout;
    float *hmax;
    unsigned *hdone;     // completion word in the pinned block: the block's sequence number once hout / hmax are complete
    unsigned *arrive;
};

struct BatchSel {
    static constexpr bool kSingle = false;
    const StreamDev *st;
    const int *fv;   // valid frames per stream, or nullptr: fv_all for every stream
    int fv_all;
    int pt;          // ring slot of the first block of the step
    // The streams of a batch live in slabs with a fixed stride: the forward kernel computes the two
    // addresses it needs from these instead of loading the descriptor first (a dependent load in
    // front of every CTA's PCM loads: 8 % of that kernel's stall samples, profiles/r02b_kernels.md).
    float2 *xring0;       // input-spectra ring of stream 0 of the launch
    size_t xring_stride;  // in float2 elements
    const char *din0;     // PCM in of stream 0 of the launch
    size_t din_stride;    // in bytes
    __device__ __forceinline__ float2 *xring(int b) const { return xring0 + (size_t)b * xring_stride; }
    __device__ __forceinline__ const void *din(int b) const { return din0 + (size_t)b * din_stride; }
    __device__ __forceinline__ StreamDev stream(int b) const { return st[b]; }
    __device__ __forceinline__ int frames(int b) const { return fv ? fv[b] : fv_all; }
    __device__ __forceinline__ int slot(int) const { return pt; }
    __device__ __forceinline__ unsigned seq(int) const { return 0u; }
    __device__ __forceinline__ const float *mix(int) const { return nullptr; }
};

// Single-stream path: the output-channel CTA of a stream that finishes LAST copies the stream's
// interleaved output block (just written to device memory by all of them, still in L2) and the
// block maximum into the caller's pinned host block -- coalesced 16-byte stores over the link
// instead of one device->host copy operation (a driver call and a copy-engine transaction) per
// stream and block.  Only the first out_bytes of the block are written, as the reference writes
// back only the frames it read (sound-processor.cc:116-125).  Last of all it publishes the block's
// sequence number in the pinned block: the caller waits for that word, not for a CUDA event, so
// a finished block costs its caller no driver call at all.  Call with all threads of the CTA,
// after the CTA's last store to s.dout and its update of s.maxv.
__device__ __forceinline__ void host_copy_out(const StreamDev &s, int nctas, size_t out_bytes, unsigned seq, int tid,
                                              int nt) {
    __shared__ int is_last;
    __syncthreads();
    if (tid == 0) {
        __threadfence();   // this CTA's output and maximum before its arrival
        is_last = atomicAdd(s.arrive, 1u) == (unsigned)(nctas - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();       // the other CTAs' output after their arrivals
    const int4 *src = reinterpret_cast<const int4 *>(s.dout);
    int4 *dst = reinterpret_cast<int4 *>(s.hout);
    const size_t n16 = out_bytes / 16;
    {   // eight loads in flight per thread before the first store: a plain copy loop is one L2 round trip per 16 bytes
        constexpr int K = 8;
        size_t i = tid;
        for (; i + (size_t)(K - 1) * nt < n16; i += (size_t)K * nt) {
            int4 v[K];
#pragma unroll
            for (int k = 0; k < K; k++) v[k] = __ldcg(src + i + (size_t)k * nt);
#pragma unroll
            for (int k = 0; k < K; k++) dst[i + (size_t)k * nt] = v[k];
        }
        for (; i < n16; i += nt) dst[i] = __ldcg(src + i);
    }
    for (size_t i = n16 * 16 + tid; i < out_bytes; i += nt)
        reinterpret_cast<unsigned char *>(s.hout)[i] = __ldcg(reinterpret_cast<const unsigned char *>(s.dout) + i);
    // One system-scope fence, by the thread that publishes the word, behind the CTA barrier: fences are
    // cumulative, so every thread's part of the block (ordered before the barrier) is visible to the host
    // before the word is -- the barrier / fence / flag idiom of NCCL's primitives.  A second fence by all
    // 256 threads in front of the barrier cost 2-4 us per block (tools/hostlink_probe.cu).
    __syncthreads();
    if (tid == 0) {
        *s.hmax = __ldcg(s.maxv);
        *s.arrive = 0u;    // ready for the stream's next block (the caller submits it only after this one)
        __threadfence_system();
        *reinterpret_cast<volatile unsigned *>(s.hdone) = seq;   // the word the caller is waiting for
    }
}

constexpr int GROUP_MAX = 32;
struct GroupSel {
    static constexpr bool kSingle = true;   // T == 1, streams may ask for the host copy-out
    const StreamDev *st[GROUP_MAX];
    int fv[GROUP_MAX];
    int pt[GROUP_MAX];
    unsigned sq[GROUP_MAX];   // sequence number of each stream's block (published by host_copy_out)
    // Optional second addend of the block's output, interleaved float frames on the device (or null):
    // the tail level of a non-uniformly partitioned filter (fcv_nonuniform.cu), added in the inverse
    // transform's epilogue before PCM conversion and maximum.  Any-size kernels (fcv_k_fft.cu) only.
    const float *mx[GROUP_MAX];
    __device__ __forceinline__ float2 *xring(int b) const { return st[b]->xring; }
    __device__ __forceinline__ const void *din(int b) const { return st[b]->din; }
    __device__ __forceinline__ StreamDev stream(int b) const { return *st[b]; }
    __device__ __forceinline__ int frames(int b) const { return fv[b]; }
    __device__ __forceinline__ int slot(int b) const { return pt[b]; }
    __device__ __forceinline__ unsigned seq(int b) const { return sq[b]; }
    __device__ __forceinline__ const float *mix(int b) const { return mx[b]; }
};

}  // namespace fcv
