/*
 * folve_b200.h -- C ABI of the B200 (sm_100a) convolution engine that sits
 * behind folve's SoundProcessor.
 *
 * This library replaces exactly one thing in the reference: the third-party
 * `Convproc` object (libzita-convolver + FFTW3f) that SoundProcessor and the
 * zita-config loader drive.  Every entry point below names the reference call
 * site (file:line under the folve tree) whose role it takes over.  Nothing
 * here takes or returns a C++ or torch type: plain pointers, ints and floats.
 *
 *   reference call                                   replaced by
 *   -----------------------------------------------  -------------------------
 *   new Convproc            sound-processor.cc:41    fcv_filter_begin + fcv_stream_create
 *   Convproc::configure     zita-fconfig.cc:80-81    fcv_filter_begin
 *   Convproc::impdata_create zita-config.cc:163,203,252  fcv_filter_add
 *   Convproc::impdata_copy  zita-config.cc:274       fcv_filter_link
 *   (end of config())       zita-config.cc:343       fcv_filter_commit
 *   inpdata/process/outdata sound-processor.cc:106-125  fcv_stream_process
 *                                                    (= fcv_stream_submit + fcv_stream_await)
 *   reset + start_process   sound-processor.cc:140,144  fcv_stream_reset
 *   stop_process/cleanup/delete sound-processor.cc:70-72 fcv_stream_destroy
 *
 * Threading (mirrors SURVEY section 8(b)): one handle is used by one thread at a
 * time but may migrate between threads; different handles may be used
 * concurrently.  Every call selects the handle's CUDA device itself.
 *
 * Errors: functions returning int give 0 on success and a negative FCV_E_*
 * code on failure; functions returning a pointer give NULL on failure.  In
 * both cases fcv_last_error() (thread local) describes the failure.  A missing
 * or unusable GPU is an error -- there is NO CPU fallback.
 */
#ifndef FOLVE_B200_H
#define FOLVE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FCV_ABI_VERSION 2

/* Limits: Convproc::MAXINP / MAXOUT / MINPART / MAXQUANT (zita-fconfig.cc:49,55,74-75)
 * and MAXSIZE (zita-config.h:61). */
#define FCV_MAXINP 64
#define FCV_MAXOUT 64
#define FCV_MINPART 64
#define FCV_MAXQUANT 8192
#define FCV_MAXSIZE 0x00100000

enum {
    FCV_OK = 0,
    FCV_E_PARAM = -1,   /* bad argument (Converror::BAD_PARAM) */
    FCV_E_STATE = -2,   /* call not valid in this state (Converror::BAD_STATE) */
    FCV_E_ALLOC = -3,   /* host or device allocation failed (Converror::MEM_ALLOC) */
    FCV_E_CUDA = -4,    /* CUDA runtime error, no device, or wrong architecture */
};

/* PCM wire formats of the interleaved blocks handed to / returned by the engine.
 * The scale factors are libsndfile's normalised-float conventions
 * (sf_readf_float sound-processor.cc:80, sf_writef_float sound-processor.cc:91). */
enum {
    FCV_PCM_F32 = 0, /* float32 in [-1,1): what sf_readf_float delivers */
    FCV_PCM_S16 = 1, /* int16; in: x/32768, out: lrintf(y*32767) */
    FCV_PCM_S24 = 2, /* int32 holding a 24-bit sample in its low 3 bytes, sign extended;
                        in: x/8388608, out: lrintf(y*8388607) */
};

typedef struct fcv_filter fcv_filter; /* immutable after commit; HBM-resident spectra; ref-counted */
typedef struct fcv_stream fcv_stream; /* one convolver state: input-spectra ring, overlap tails */
typedef struct fcv_batch fcv_batch;   /* a set of streams of one filter processed in lock-step */

/* ---- library ------------------------------------------------------------ */
int fcv_abi_version(void);
const char *fcv_last_error(void);
/* Number of usable sm_100 devices; <0 on CUDA failure. */
int fcv_device_count(void);

/* ---- filter (== the configured, impulse-loaded part of a Convproc) ------- */

/* Convproc::configure(ninp, nout, size, fragm, fragm, fragm[, dens])
 * zita-fconfig.cc:80-81.  `fragm` must be a power of two in [64, 8192];
 * partitions = ceil(size / fragm). */
fcv_filter *fcv_filter_begin(int ninp, int nout, unsigned size, unsigned fragm);

/* Convproc::impdata_create(inp, out, step, data, ind0, ind1) zita-config.cc:163:
 * ADDS data[k*step], k in [0, ind1-ind0), at taps ind0.. of pair (inp,out)
 * (0-based).  Taps at or beyond partitions*fragm are dropped, as zita does.
 * Ignored without error if the pair is currently a link target. */
int fcv_filter_add(fcv_filter *f, int inp, int out, int step, const float *data, int ind0, int ind1);

/* Convproc::impdata_copy(inp1, out1, inp2, out2) zita-config.cc:274: make pair
 * (inp2,out2) use the spectra of (inp1,out1), dropping its own data.  No-op
 * when (inp1,out1) has no impulse data yet, as in zita. */
int fcv_filter_link(fcv_filter *f, int inp1, int out1, int inp2, int out2);

/* Partition, scale by 1/(2*fragm), forward-transform on the device and keep
 * the spectra resident in HBM on CUDA device `device`.  After this the filter
 * is immutable and may be shared by any number of streams on that device. */
int fcv_filter_commit(fcv_filter *f, int device);

void fcv_filter_ref(fcv_filter *f);
void fcv_filter_unref(fcv_filter *f);

int fcv_filter_ninp(const fcv_filter *f);
int fcv_filter_nout(const fcv_filter *f);
int fcv_filter_fragm(const fcv_filter *f);
/* ceil(size/fragm): the partition count zita allocates room for. */
int fcv_filter_partitions(const fcv_filter *f);
/* Depth of the input-spectra ring the engine keeps (index of the last non-zero
 * partition + 1) and number of non-zero (pair, partition) spectra it streams. */
int fcv_filter_ring_depth(const fcv_filter *f);
int fcv_filter_active_rows(const fcv_filter *f);
/* Number of (inp,out) pairs with at least one non-zero partition (links included). */
int fcv_filter_active_pairs(const fcv_filter *f);
int fcv_filter_device(const fcv_filter *f);

/* ---- stream (== the running state of a Convproc) ------------------------- */

/* new Convproc + reset + start_process: all-zero state. */
fcv_stream *fcv_stream_create(fcv_filter *f);
/* The same with PCM wire formats other than float (FCV_PCM_*): the block buffer then holds
 * in_format samples on the way in and out_format samples on the way out, converted on the
 * device with libsndfile's scale factors (what sf_readf_short / sf_writef_short, or the 24-bit
 * paths, would have done on the host around sound-processor.cc:80,91). */
fcv_stream *fcv_stream_create_fmt(fcv_filter *f, int in_format, int out_format);
/* stop_process + cleanup + delete (sound-processor.cc:70-72). */
void fcv_stream_destroy(fcv_stream *s);
/* Convproc::reset + start_process (sound-processor.cc:140,144): state identical
 * to a freshly created stream; also zeroes the running maximum. */
int fcv_stream_reset(fcv_stream *s);

/* Pinned host block of fragm * max(ninp, nout) floats, the analogue of
 * SoundProcessor::buffer_ (sound-processor.cc:62-63).  Input frames are written
 * here interleaved; processed frames are read back from here interleaved. */
float *fcv_stream_buffer(fcv_stream *s);
/* Size of that block in bytes: fragm * max(ninp * input sample size, nout * output sample size). */
size_t fcv_stream_buffer_bytes(const fcv_stream *s);

/* SoundProcessor::Process() (sound-processor.cc:98-127) on the stream's own
 * buffer: the first `frames_valid` interleaved input frames are used, the rest
 * of the block is taken as zero; one block is convolved; the first
 * `frames_valid` interleaved output frames are written back to the buffer.
 * `*max_inout`, if given, is raised to the largest SIGNED output sample seen
 * (the reference compares without fabs: sound-processor.cc:120-123).
 * As in the reference, buffer content behind the first frames_valid input frames is zeroed
 * (sound-processor.cc:99-103) and only frames_valid output frames are written back.
 * Synchronous: returns when the output is in the buffer.
 *
 * Calls made at the same time from different threads on different streams of one filter
 * (folve: one open file per thread) are coalesced inside the library into ONE launch sequence
 * on the GPU; each caller still gets exactly the result of a call of its own. */
int fcv_stream_process(fcv_stream *s, int frames_valid, float *max_inout);

/* The two halves of fcv_stream_process, for callers that keep several files going from one
 * thread (north star: "one CUDA stream per open file"): submit queues the block that is in the
 * stream's buffer and returns without waiting for the GPU; await returns when the output is in the
 * buffer and raises *max_inout.  The buffer must not be touched in between.  One block per stream
 * can be in flight; blocks of different streams submitted before their awaits travel together. */
int fcv_stream_submit(fcv_stream *s, int frames_valid);
int fcv_stream_await(fcv_stream *s, float *max_inout);

fcv_filter *fcv_stream_filter(fcv_stream *s);

/* ---- non-uniform partitioning ------------------------------------------------
 * Convproc::configure(ninp, nout, maxsize, quantum, minpart, maxpart) with quantum = minpart <
 * maxpart: zita-convolver's non-uniform mode.  folve itself always configures uniform partitions
 * (quantum = minpart = maxpart = fragm, zita-fconfig.cc:74-93); a caller that wants a smaller block
 * than fragm -- lower latency -- uses these.  Two levels: the first `maxpart` taps as partitions of
 * `quantum` frames, evaluated for every block of `quantum` frames; the rest as partitions of
 * `maxpart` frames, evaluated once per `maxpart` frames and always one large block ahead of where its
 * output is needed.  Output == the uniform engine's with fragm = maxpart (same truncation, additive
 * impulses, links), up to float32 rounding.  float32 PCM only. */
typedef struct fcv_nufilter fcv_nufilter;
typedef struct fcv_nustream fcv_nustream;
fcv_nufilter *fcv_nufilter_begin(int ninp, int nout, unsigned size, unsigned quantum, unsigned maxpart);
int fcv_nufilter_add(fcv_nufilter *f, int inp, int out, int step, const float *data, int ind0, int ind1);
int fcv_nufilter_link(fcv_nufilter *f, int inp1, int out1, int inp2, int out2);
int fcv_nufilter_commit(fcv_nufilter *f, int device);
void fcv_nufilter_ref(fcv_nufilter *f);
void fcv_nufilter_unref(fcv_nufilter *f);
int fcv_nufilter_quantum(const fcv_nufilter *f);
int fcv_nufilter_head_partitions(const fcv_nufilter *f);
int fcv_nufilter_tail_partitions(const fcv_nufilter *f);
fcv_nustream *fcv_nustream_create(fcv_nufilter *f);
void fcv_nustream_destroy(fcv_nustream *s);
int fcv_nustream_reset(fcv_nustream *s);
/* Pinned block of quantum * max(ninp, nout) floats: interleaved input frames in, processed frames out. */
float *fcv_nustream_buffer(fcv_nustream *s);
/* One block of up to `quantum` frames, synchronous, like fcv_stream_process.  A block shorter than
 * `quantum` ends the stream (reset before the next file). */
int fcv_nustream_process(fcv_nustream *s, int frames_valid, float *max_inout);

/* ---- batch: many streams of one filter, one launch per stage ------------- */

/* Creates `nstreams` fresh streams that share `f` and are advanced together.
 * in_format / out_format: FCV_PCM_*. */
fcv_batch *fcv_batch_create(fcv_filter *f, int nstreams, int in_format, int out_format);
/* Same, but every step carries `blocks_per_step` (1, 2, 4 or 8) consecutive
 * blocks of every stream: staging areas are [nstreams][blocks_per_step*fragm][channels]
 * and frames_valid counts are in [0, blocks_per_step*fragm].  With more than one
 * block per step the complex MAC is time-tiled: each input spectrum is read from
 * HBM once for all the output blocks of the step that need it.  Results are
 * identical to feeding the blocks one at a time.  Blocks after a stream's
 * frames_valid are processed as silence (follow a short step by a slot reset). */
fcv_batch *fcv_batch_create_tiled(fcv_filter *f, int nstreams, int in_format, int out_format, int blocks_per_step);
int fcv_batch_blocks_per_step(const fcv_batch *b);
void fcv_batch_destroy(fcv_batch *b);
int fcv_batch_nstreams(const fcv_batch *b);

/* Pinned host staging areas, [nstreams][fragm][channels] in the wire format. */
void *fcv_batch_host_in(fcv_batch *b);
void *fcv_batch_host_out(fcv_batch *b);
size_t fcv_batch_host_in_bytes(const fcv_batch *b);
size_t fcv_batch_host_out_bytes(const fcv_batch *b);
/* Device staging areas with the same layout (for callers that produce or
 * consume PCM on the GPU, and for device-resident benchmarking). */
void *fcv_batch_device_in(fcv_batch *b);
void *fcv_batch_device_out(fcv_batch *b);

/* One block for every stream, end to end: host_in -> device, convolve,
 * device -> host_out; returns when host_out is complete.  frames_valid may be
 * NULL (all blocks full) or an array of nstreams counts in [0, fragm]. */
int fcv_batch_process(fcv_batch *b, const int *frames_valid);

/* Asynchronous form with two host staging slots (slot 0 is host_in/host_out
 * above), so that the copies of consecutive blocks overlap: submit enqueues
 * host_in_slot(slot) -> device, convolve, device -> host_out_slot(slot) and
 * returns; wait blocks until that slot's output is complete.  Blocks are
 * processed in submit order.  fcv_batch_process == submit(0) + wait(0). */
void *fcv_batch_host_in_slot(fcv_batch *b, int slot);
void *fcv_batch_host_out_slot(fcv_batch *b, int slot);
int fcv_batch_submit(fcv_batch *b, int slot, const int *frames_valid);
int fcv_batch_wait(fcv_batch *b, int slot);

/* Diagnostic for benchmarks: while on, fcv_batch_submit moves the PCM host->device and
 * device->host exactly as usual but launches no kernel -- the host link's ceiling for the
 * end-to-end loop, measured with the same code path.  Synchronises. */
int fcv_batch_set_copy_only(fcv_batch *b, int on);

/* One block for every stream on device-resident PCM (device_in -> device_out),
 * asynchronous on the batch's CUDA stream; no host copies.  Call
 * fcv_batch_sync() to wait. */
int fcv_batch_process_device(fcv_batch *b, const int *frames_valid);
int fcv_batch_sync(fcv_batch *b);

/* Per-slot control: reset one stream of the batch to the fresh state (end of a
 * gapless album chain), read its running signed maximum. */
int fcv_batch_reset_slot(fcv_batch *b, int slot);
int fcv_batch_get_max(fcv_batch *b, float *max_out /* [nstreams] */);
/* Signed maximum (>= 0, over all output channels, valid frames only) of every block of
 * the last step: what SoundProcessor::Process() adds to max_out_value_observed_ block by
 * block (sound-processor.cc:120-123), so that a caller stepping several blocks at once
 * can still tell the running maximum at the block where a file ended. */
int fcv_batch_get_block_max(fcv_batch *b, float *max_out /* [nstreams][blocks_per_step] */);
/* The same two operations for a caller that keeps two steps in flight with fcv_batch_submit /
 * fcv_batch_wait: the reset is enqueued behind the steps already submitted and ahead of the
 * next one (no host synchronisation); the block maxima of the step submitted from host slot
 * `slot` travel back with its output and are valid after fcv_batch_wait(b, slot). */
int fcv_batch_reset_slot_async(fcv_batch *b, int slot);
const float *fcv_batch_host_block_max_slot(fcv_batch *b, int slot /* -> [nstreams][blocks_per_step] */);

/* cudaStream_t the batch launches on (as void*), for CUDA-event timing by the caller. */
void *fcv_batch_cuda_stream(fcv_batch *b);

/* CUDA-event stopwatch on the batch's stream: record event `slot` (0..15) now,
 * and later read the device time between two recorded slots (synchronises). */
int fcv_batch_event_record(fcv_batch *b, int slot);
int fcv_batch_event_elapsed_ms(fcv_batch *b, int slot0, int slot1, float *ms);

/* Per-kernel timing: when enabled, every launch is bracketed by CUDA events on
 * the batch's stream.  fcv_batch_profile() synchronises and returns the summed
 * device milliseconds per kernel since the last call and the number of blocks
 * (steps) they cover: ms[0] forward FFT, ms[1] complex MAC, ms[2] inverse FFT. */
int fcv_batch_set_profiling(fcv_batch *b, int on);
int fcv_batch_profile(fcv_batch *b, float ms[3], int *steps);
/* Number of kernel launches issued by this library since load (all handles). */
unsigned long long fcv_kernel_launches(void);

/* ---- test hooks (kernel-level parity; not used by SoundProcessor) -------- */

/* Copy spectra to the host in NATURAL bin order, fragm+1 interleaved complex
 * values each.  fcv_filter_get_spectrum: partition j of pair (inp,out), returns
 * 1 if present, 0 if the partition is absent (all zero), <0 on error.
 * fcv_stream_get_input_spectrum: ring slot `age` blocks back (0 = newest). */
int fcv_filter_get_spectrum(fcv_filter *f, int inp, int out, int j, float *dst);
/* Before commit (no GPU needed): the accumulated time-domain impulse of pair
 * (inp,out) as it will be transformed, i.e. already scaled by 0.5/fragm;
 * `capacity` floats at most (partitions*fragm are available).  Returns 0 if
 * the pair has no MAC node, 1 if it owns data (possibly all zero), 2 if it is
 * a link (dst receives the source pair's data), <0 on error. */
int fcv_filter_get_impulse(fcv_filter *f, int inp, int out, float *dst, int capacity);
int fcv_stream_get_input_spectrum(fcv_stream *s, int inp, int age, float *dst);
/* Process-wide switch between the two implementations of a single-stream group: three chained
 * launches (default) or one cooperative launch where the shape is covered (fragm 8192, stereo;
 * also FCV_FUSED=1).  Results are identical; the tests use it to show that, and count the
 * cooperative launches that really happened. */
void fcv_debug_set_fused(int on);
unsigned long long fcv_debug_fused_launches(void);
/* Process-wide switch of the inverse transform for stereo blocks of fragm 8192: one CTA per output
 * channel storing its samples on its own (default), or the two channels' CTAs as a thread block cluster
 * that exchanges the converted samples through distributed shared memory and writes whole interleaved
 * frames.  mask bit 0: batches (also FCV_INV_PAIR=1), bit 1: the per-file path, where the frames then
 * go straight to the caller's pinned block (also FCV_INV_PAIR_SINGLE=1).  Results are identical; the
 * pair is not faster on either path. */
void fcv_debug_set_inv_pair(int mask);
/* Process-wide switch of the tensor-memory variants of the fragm 8192 transforms of a batch with several
 * blocks per step (a thread's twiddles -- and the inverse transform's overlap tail -- kept in the SM's
 * tensor memory across the blocks of a step instead of being re-read from global memory in every block).
 * mask bit 0: inverse transform (on by default; FCV_INV_TMEM=0 turns it off), bit 1: forward transform of
 * stereo blocks (off by default: measured slower; FCV_FWD_TMEM=1).  Results are bit-identical either way. */
void fcv_debug_set_tmem(int mask);
int fcv_debug_get_tmem(void);

#ifdef __cplusplus
}
#endif
#endif /* FOLVE_B200_H */
