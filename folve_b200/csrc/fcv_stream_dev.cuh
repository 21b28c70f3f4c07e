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
};

struct BatchSel {
    const StreamDev *st;
    const int *fv;   // valid frames per stream, or nullptr: fv_all for every stream
    int fv_all;
    int pt;          // ring slot of the first block of the step
    __device__ __forceinline__ StreamDev stream(int b) const { return st[b]; }
    __device__ __forceinline__ int frames(int b) const { return fv ? fv[b] : fv_all; }
    __device__ __forceinline__ int slot(int) const { return pt; }
};

constexpr int GROUP_MAX = 32;
struct GroupSel {
    const StreamDev *st[GROUP_MAX];
    int fv[GROUP_MAX];
    int pt[GROUP_MAX];
    __device__ __forceinline__ StreamDev stream(int b) const { return *st[b]; }
    __device__ __forceinline__ int frames(int b) const { return fv[b]; }
    __device__ __forceinline__ int slot(int b) const { return pt[b]; }
};

}  // namespace fcv
