/*
 * sndfile.h -- minimal stand-in for the libsndfile API subset that folve's hot
 * path touches, for environments (like this build image) where libsndfile is
 * not installed.
 *
 * What the hot path uses (and therefore what is here):
 *   sf_readf_float   /root/reference/sound-processor.cc:80, zita-audiofile.cc:181
 *   sf_writef_float  /root/reference/sound-processor.cc:91
 *   sf_open/sf_close/sf_seek/sf_command + SF_INFO and the format constants
 *                    /root/reference/zita-audiofile.cc:51-99,162-175
 *   sf_open_fd / sf_open_virtual / sf_get_string / sf_set_string
 *                    /root/reference/convolve-file-handler.cc:62,490-492,
 *                    /root/reference/conversion-buffer.cc:88-98 -- so that the reference's
 *                    own file handler and conversion buffer compile and run unmodified
 *                    on top of this shim (oracle/Makefile target `dropin`)
 * plus memory-backed files (sf_shim_*) that stand in for the FLAC decoder /
 * encoder around SoundProcessor in tests and benchmarks.
 *
 * Writing through sf_open_virtual produces, for SF_FORMAT_FLAC, an UNCOMPRESSED stand-in
 * with FLAC's header geometry: "fLaC", one STREAMINFO block (42 bytes in all, so that the
 * byte offsets folve patches -- convolve-file-handler.cc:289-310 -- mean what they mean in a
 * real FLAC file), then the interleaved little-endian PCM frames.  The codec itself is out of
 * scope.  SF_FORMAT_WAV gives a plain 44-byte RIFF header.
 *
 * Supported containers: RIFF/WAVE (PCM 16/24/32 and IEEE float32, also
 * WAVE_FORMAT_EXTENSIBLE) for reading; memory sinks for writing.  Sample
 * conversion follows libsndfile's normalised-float conventions:
 *   read : int16 -> x / 0x8000,  int24 -> x / 0x800000,  int32 -> x / 0x80000000
 *   write: lrintf(x * 0x7FFF) / lrintf(x * 0x7FFFFF) / lrintf(x * 0x7FFFFFFF), no clipping
 *          unless SFC_SET_CLIPPING was issued (folve never does).
 * In a real folve build this header is NOT used: link against libsndfile.
 */
#ifndef FOLVE_B200_SNDFILE_SHIM_H
#define FOLVE_B200_SNDFILE_SHIM_H

#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FOLVE_B200_SNDFILE_SHIM 1

typedef int64_t sf_count_t;
typedef struct SNDFILE_tag SNDFILE;

typedef struct SF_INFO {
    sf_count_t frames;
    int samplerate;
    int channels;
    int format;
    int sections;
    int seekable;
} SF_INFO;

enum {
    SF_FORMAT_WAV = 0x010000,
    SF_FORMAT_AIFF = 0x020000,
    SF_FORMAT_FLAC = 0x170000,
    SF_FORMAT_CAF = 0x180000,
    SF_FORMAT_WAVEX = 0x130000,
    SF_FORMAT_OGG = 0x200000,

    SF_FORMAT_PCM_S8 = 0x0001,
    SF_FORMAT_PCM_16 = 0x0002,
    SF_FORMAT_PCM_24 = 0x0003,
    SF_FORMAT_PCM_32 = 0x0004,
    SF_FORMAT_FLOAT = 0x0006,
    SF_FORMAT_DOUBLE = 0x0007,
    SF_FORMAT_VORBIS = 0x0060,

    SF_FORMAT_SUBMASK = 0x0000FFFF,
    SF_FORMAT_TYPEMASK = 0x0FFF0000,
    SF_FORMAT_ENDMASK = 0x30000000
};

enum { SFM_READ = 0x10, SFM_WRITE = 0x20, SFM_RDWR = 0x30 };

enum {
    SF_STR_TITLE = 0x01, SF_STR_COPYRIGHT = 0x02, SF_STR_SOFTWARE = 0x03, SF_STR_ARTIST = 0x04,
    SF_STR_COMMENT = 0x05, SF_STR_DATE = 0x06, SF_STR_ALBUM = 0x07, SF_STR_LICENSE = 0x08,
    SF_STR_TRACKNUMBER = 0x09, SF_STR_GENRE = 0x10
};
#define SF_STR_FIRST SF_STR_TITLE
#define SF_STR_LAST SF_STR_GENRE

typedef sf_count_t (*sf_vio_get_filelen)(void *user_data);
typedef sf_count_t (*sf_vio_seek)(sf_count_t offset, int whence, void *user_data);
typedef sf_count_t (*sf_vio_read)(void *ptr, sf_count_t count, void *user_data);
typedef sf_count_t (*sf_vio_write)(const void *ptr, sf_count_t count, void *user_data);
typedef sf_count_t (*sf_vio_tell)(void *user_data);
typedef struct SF_VIRTUAL_IO {
    sf_vio_get_filelen get_filelen;
    sf_vio_seek seek;
    sf_vio_read read;
    sf_vio_write write;
    sf_vio_tell tell;
} SF_VIRTUAL_IO;

enum {
    SFC_SET_CLIPPING = 0x10C0,
    SFC_GET_CLIPPING = 0x10C1,
    SFC_UPDATE_HEADER_NOW = 0x1060,
    SFC_WAVEX_SET_AMBISONIC = 0x1200,
    SFC_WAVEX_GET_AMBISONIC = 0x1201
};
enum { SF_AMBISONIC_NONE = 0x40, SF_AMBISONIC_B_FORMAT = 0x41 };
enum { SF_FALSE = 0, SF_TRUE = 1 };

SNDFILE *sf_open(const char *path, int mode, SF_INFO *sfinfo);
/* SFM_READ of a RIFF/WAVE file through a descriptor (dup'ed; close_desc closes the original). */
SNDFILE *sf_open_fd(int fd, int mode, SF_INFO *sfinfo, int close_desc);
/* SFM_WRITE only: header and PCM go out through sfvirtual->write. */
SNDFILE *sf_open_virtual(SF_VIRTUAL_IO *sfvirtual, int mode, SF_INFO *sfinfo, void *user_data);
/* string tags: none are stored */
const char *sf_get_string(SNDFILE *sndfile, int str_type);
int sf_set_string(SNDFILE *sndfile, int str_type, const char *str);
int sf_close(SNDFILE *sndfile);
sf_count_t sf_seek(SNDFILE *sndfile, sf_count_t frames, int whence);
sf_count_t sf_readf_float(SNDFILE *sndfile, float *ptr, sf_count_t frames);
sf_count_t sf_writef_float(SNDFILE *sndfile, const float *ptr, sf_count_t frames);
/* 16-bit access for the batched submit layer's int16 wire format (libsndfile has both; the shim
 * only implements them for 16-bit PCM files, where they are plain copies). */
sf_count_t sf_readf_short(SNDFILE *sndfile, short *ptr, sf_count_t frames);
sf_count_t sf_writef_short(SNDFILE *sndfile, const short *ptr, sf_count_t frames);
int sf_command(SNDFILE *sndfile, int command, void *data, int datasize);
const char *sf_strerror(SNDFILE *sndfile);
const char *sf_version_string(void);

/* ---- shim extensions: memory-backed PCM sources and sinks ------------------ */

/* Read-only file over caller-owned interleaved PCM.  `format` is
 * SF_FORMAT_PCM_16 (int16_t), SF_FORMAT_PCM_24 / SF_FORMAT_PCM_32 (int32_t,
 * 24-bit values sign-extended in the low bytes) or SF_FORMAT_FLOAT (float).
 * The memory must outlive the handle. */
SNDFILE *sf_shim_open_memory_read(const void *pcm, sf_count_t frames, int channels, int samplerate,
                                  int format);
/* Write-only sink that quantises to `format` and keeps the samples in memory. */
SNDFILE *sf_shim_open_memory_write(int channels, int samplerate, int format);
/* Write-only sink that converts like `format` would but discards the samples
 * (benchmarks: no memory growth). */
SNDFILE *sf_shim_open_null_write(int channels, int samplerate, int format);
/* Frames written so far and a pointer to the interleaved samples (int16_t for
 * PCM_16, int32_t for PCM_24/PCM_32, float for FLOAT); valid until sf_close. */
sf_count_t sf_shim_memory_frames(SNDFILE *sndfile);
const void *sf_shim_memory_data(SNDFILE *sndfile);

#ifdef __cplusplus
}
#endif
#endif
