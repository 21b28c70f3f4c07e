// sndfile_shim.cc -- see sndfile.h in this directory.  Own implementation of a
// small libsndfile API subset: RIFF/WAVE reader and memory-backed sources/sinks.
#include "sndfile.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <string>
#include <vector>

namespace {
thread_local std::string g_error = "No Error.";
}

struct SNDFILE_tag {
    int mode = 0;
    SF_INFO info{};
    int subformat = 0;
    bool clipping = false;
    sf_count_t pos = 0;  // in frames
    bool discard = false;
    // file-backed read
    FILE *fp = nullptr;
    long data_offset = 0;
    int bytes_per_sample = 0;
    // memory-backed read
    const void *mem = nullptr;
    // memory-backed write
    std::vector<int16_t> w16;
    std::vector<int32_t> w32;
    std::vector<float> wf;
    std::vector<unsigned char> scratch;
    // virtual write (sf_open_virtual)
    bool virt = false;
    SF_VIRTUAL_IO vio{};
    void *vuser = nullptr;
    bool header_written = false;
};

static uint32_t rd32(const unsigned char *p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }
static uint16_t rd16(const unsigned char *p) { return (uint16_t)(p[0] | (p[1] << 8)); }

static SNDFILE *open_wav_fp(FILE *fp, SF_INFO *sfinfo);

static SNDFILE *open_wav(const char *path, SF_INFO *sfinfo) {
    FILE *fp = fopen(path, "rb");
    if (!fp) {
        g_error = std::string("System error : could not open '") + path + "'.";
        return nullptr;
    }
    return open_wav_fp(fp, sfinfo);
}

static SNDFILE *open_wav_fp(FILE *fp, SF_INFO *sfinfo) {
    unsigned char hdr[12];
    if (fread(hdr, 1, 12, fp) != 12 || memcmp(hdr, "RIFF", 4) || memcmp(hdr + 8, "WAVE", 4)) {
        fclose(fp);
        g_error = "File contains data in an unknown format.";
        return nullptr;
    }
    int tag = 0, channels = 0, rate = 0, bits = 0, block_align = 0;
    bool have_fmt = false;
    long data_offset = -1;
    uint32_t data_bytes = 0;
    for (;;) {
        unsigned char ch[8];
        if (fread(ch, 1, 8, fp) != 8) break;
        const uint32_t sz = rd32(ch + 4);
        if (!memcmp(ch, "fmt ", 4)) {
            unsigned char f[40] = {0};
            const size_t n = sz < sizeof(f) ? sz : sizeof(f);
            if (fread(f, 1, n, fp) != n) break;
            tag = rd16(f);
            channels = rd16(f + 2);
            rate = (int)rd32(f + 4);
            block_align = rd16(f + 12);
            bits = rd16(f + 14);
            if (tag == 0xFFFE && n >= 26) tag = rd16(f + 24);  // WAVE_FORMAT_EXTENSIBLE sub-format
            have_fmt = true;
            fseek(fp, (long)(sz - n) + (sz & 1), SEEK_CUR);
        } else if (!memcmp(ch, "data", 4)) {
            data_offset = ftell(fp);
            data_bytes = sz;
            break;
        } else {
            fseek(fp, (long)sz + (sz & 1), SEEK_CUR);
        }
    }
    int sub = 0;
    if (have_fmt && tag == 1 && bits == 16) sub = SF_FORMAT_PCM_16;
    else if (have_fmt && tag == 1 && bits == 24) sub = SF_FORMAT_PCM_24;
    else if (have_fmt && tag == 1 && bits == 32) sub = SF_FORMAT_PCM_32;
    else if (have_fmt && tag == 3 && bits == 32) sub = SF_FORMAT_FLOAT;
    if (!sub || data_offset < 0 || channels < 1 || block_align != channels * bits / 8) {
        fclose(fp);
        g_error = "Format not recognised.";
        return nullptr;
    }
    // a data chunk running past the end of the file is clamped, as libsndfile does
    fseek(fp, 0, SEEK_END);
    const long file_end = ftell(fp);
    if ((long)data_bytes > file_end - data_offset) data_bytes = (uint32_t)(file_end - data_offset);
    fseek(fp, data_offset, SEEK_SET);

    SNDFILE *s = new SNDFILE_tag();
    s->mode = SFM_READ;
    s->fp = fp;
    s->data_offset = data_offset;
    s->bytes_per_sample = bits / 8;
    s->subformat = sub;
    s->info.frames = data_bytes / (uint32_t)block_align;
    s->info.samplerate = rate;
    s->info.channels = channels;
    s->info.format = SF_FORMAT_WAV | sub;
    s->info.sections = 1;
    s->info.seekable = 1;
    *sfinfo = s->info;
    return s;
}

extern "C" SNDFILE *sf_open(const char *path, int mode, SF_INFO *sfinfo) {
    if (!path || !sfinfo) { g_error = "Bad argument."; return nullptr; }
    if (mode == SFM_READ) return open_wav(path, sfinfo);
    g_error = "sndfile shim: only SFM_READ of WAV files and memory sinks are supported.";
    return nullptr;
}

extern "C" SNDFILE *sf_open_fd(int fd, int mode, SF_INFO *sfinfo, int close_desc) {
    if (fd < 0 || !sfinfo || mode != SFM_READ) { g_error = "Bad argument."; return nullptr; }
    const int own = dup(fd);
    if (close_desc) close(fd);
    FILE *fp = own >= 0 ? fdopen(own, "rb") : nullptr;
    if (!fp) {
        if (own >= 0) close(own);
        g_error = "System error : could not use the descriptor.";
        return nullptr;
    }
    fseek(fp, 0, SEEK_SET);
    return open_wav_fp(fp, sfinfo);
}

// ---- virtual write: header once, then interleaved little-endian PCM -------------------------
static int sub_bits(int sub) {
    return sub == SF_FORMAT_PCM_16 ? 16 : sub == SF_FORMAT_PCM_24 ? 24 : 32;
}

static void virt_header(SNDFILE *s) {
    if (!s->virt || s->header_written) return;
    s->header_written = true;
    const int ch = s->info.channels, rate = s->info.samplerate, bits = sub_bits(s->subformat);
    unsigned char h[44];
    memset(h, 0, sizeof(h));
    if ((s->info.format & SF_FORMAT_TYPEMASK) == SF_FORMAT_FLAC) {
        // "fLaC" | block header: last, STREAMINFO, 34 bytes | min/max blocksize 4096 | frame
        // sizes 0 | rate:20 channels-1:3 bps-1:5 samples:36 | md5 = 0
        memcpy(h, "fLaC", 4);
        h[4] = 0x80; h[7] = 34;
        h[8] = 0x10; h[10] = 0x10;
        h[18] = (unsigned char)(rate >> 12);
        h[19] = (unsigned char)(rate >> 4);
        h[20] = (unsigned char)(((rate & 0x0f) << 4) | ((ch - 1) << 1) | (((bits - 1) & 0x10) >> 4));
        h[21] = (unsigned char)(((bits - 1) & 0x0f) << 4);
        s->vio.write(h, 42, s->vuser);
        return;
    }
    const int block = ch * bits / 8;
    memcpy(h, "RIFF", 4);
    memcpy(h + 8, "WAVEfmt ", 8);
    h[16] = 16;
    h[20] = (unsigned char)(s->subformat == SF_FORMAT_FLOAT ? 3 : 1);
    h[22] = (unsigned char)ch;
    for (int i = 0; i < 4; i++) h[24 + i] = (unsigned char)((unsigned)rate >> (8 * i));
    for (int i = 0; i < 4; i++) h[28 + i] = (unsigned char)((unsigned)(rate * block) >> (8 * i));
    h[32] = (unsigned char)block;
    h[34] = (unsigned char)bits;
    memcpy(h + 36, "data", 4);
    s->vio.write(h, 44, s->vuser);
}

extern "C" SNDFILE *sf_open_virtual(SF_VIRTUAL_IO *v, int mode, SF_INFO *sfinfo, void *user_data) {
    if (!v || !v->write || !sfinfo || mode != SFM_WRITE) {
        g_error = "sndfile shim: sf_open_virtual supports SFM_WRITE only.";
        return nullptr;
    }
    const int type = sfinfo->format & SF_FORMAT_TYPEMASK, sub = sfinfo->format & SF_FORMAT_SUBMASK;
    const bool sub_ok = sub == SF_FORMAT_PCM_16 || sub == SF_FORMAT_PCM_24 || sub == SF_FORMAT_PCM_32 ||
                        (sub == SF_FORMAT_FLOAT && type == SF_FORMAT_WAV);
    if ((type != SF_FORMAT_FLAC && type != SF_FORMAT_WAV) || !sub_ok || sfinfo->channels < 1) {
        g_error = "Format not recognised.";
        return nullptr;
    }
    SNDFILE *s = new SNDFILE_tag();
    s->mode = SFM_WRITE;
    s->info = *sfinfo;
    s->info.frames = 0;
    s->subformat = sub;
    s->virt = true;
    s->vio = *v;
    s->vuser = user_data;
    return s;
}

extern "C" const char *sf_get_string(SNDFILE *, int) { return nullptr; }
extern "C" int sf_set_string(SNDFILE *, int, const char *) { return 0; }

extern "C" SNDFILE *sf_shim_open_memory_read(const void *pcm, sf_count_t frames, int channels,
                                             int samplerate, int format) {
    const int sub = format & SF_FORMAT_SUBMASK;
    if ((!pcm && frames) || frames < 0 || channels < 1 ||
        (sub != SF_FORMAT_PCM_16 && sub != SF_FORMAT_PCM_24 && sub != SF_FORMAT_PCM_32 && sub != SF_FORMAT_FLOAT)) {
        g_error = "Bad argument.";
        return nullptr;
    }
    SNDFILE *s = new SNDFILE_tag();
    s->mode = SFM_READ;
    s->mem = pcm;
    s->subformat = sub;
    s->info.frames = frames;
    s->info.samplerate = samplerate;
    s->info.channels = channels;
    s->info.format = (format & SF_FORMAT_TYPEMASK ? format & SF_FORMAT_TYPEMASK : SF_FORMAT_FLAC) | sub;
    s->info.sections = 1;
    s->info.seekable = 1;
    return s;
}

extern "C" SNDFILE *sf_shim_open_memory_write(int channels, int samplerate, int format) {
    const int sub = format & SF_FORMAT_SUBMASK;
    if (channels < 1 ||
        (sub != SF_FORMAT_PCM_16 && sub != SF_FORMAT_PCM_24 && sub != SF_FORMAT_PCM_32 && sub != SF_FORMAT_FLOAT)) {
        g_error = "Bad argument.";
        return nullptr;
    }
    SNDFILE *s = new SNDFILE_tag();
    s->mode = SFM_WRITE;
    s->subformat = sub;
    s->info.samplerate = samplerate;
    s->info.channels = channels;
    s->info.format = (format & SF_FORMAT_TYPEMASK ? format & SF_FORMAT_TYPEMASK : SF_FORMAT_FLAC) | sub;
    s->info.sections = 1;
    return s;
}

extern "C" SNDFILE *sf_shim_open_null_write(int channels, int samplerate, int format) {
    SNDFILE *s = sf_shim_open_memory_write(channels, samplerate, format);
    if (s) s->discard = true;
    return s;
}

extern "C" int sf_close(SNDFILE *s) {
    if (s && s->virt) virt_header(s);
    if (!s) return 1;
    if (s->fp) fclose(s->fp);
    delete s;
    return 0;
}

extern "C" sf_count_t sf_seek(SNDFILE *s, sf_count_t frames, int whence) {
    if (!s || s->mode != SFM_READ) return -1;
    sf_count_t target = frames;
    if (whence == SEEK_CUR) target = s->pos + frames;
    else if (whence == SEEK_END) target = s->info.frames + frames;
    if (target < 0 || target > s->info.frames) { g_error = "Attempt to seek beyond the end of the file."; return -1; }
    if (s->fp) fseek(s->fp, s->data_offset + (long)(target * s->info.channels * s->bytes_per_sample), SEEK_SET);
    s->pos = target;
    return target;
}

extern "C" sf_count_t sf_readf_float(SNDFILE *s, float *ptr, sf_count_t frames) {
    if (!s || s->mode != SFM_READ || frames <= 0) return 0;
    const sf_count_t left = s->info.frames - s->pos;
    if (frames > left) frames = left;
    if (frames <= 0) return 0;
    const size_t ch = (size_t)s->info.channels, n = (size_t)frames * ch;
    if (s->mem) {
        const size_t off = (size_t)s->pos * ch;
        switch (s->subformat) {
            case SF_FORMAT_PCM_16: {
                const int16_t *p = (const int16_t *)s->mem + off;
                for (size_t i = 0; i < n; i++) ptr[i] = (float)p[i] * (1.0f / 32768.0f);
                break;
            }
            case SF_FORMAT_PCM_24: {
                const int32_t *p = (const int32_t *)s->mem + off;
                for (size_t i = 0; i < n; i++) ptr[i] = (float)p[i] * (1.0f / 8388608.0f);
                break;
            }
            case SF_FORMAT_PCM_32: {
                const int32_t *p = (const int32_t *)s->mem + off;
                for (size_t i = 0; i < n; i++) ptr[i] = (float)p[i] * (1.0f / 2147483648.0f);
                break;
            }
            default:
                memcpy(ptr, (const float *)s->mem + off, n * sizeof(float));
        }
    } else {
        const size_t bps = (size_t)s->bytes_per_sample;
        s->scratch.resize(n * bps);
        const size_t got = fread(s->scratch.data(), bps * ch, (size_t)frames, s->fp);
        frames = (sf_count_t)got;
        const size_t m = got * ch;
        const unsigned char *b = s->scratch.data();
        switch (s->subformat) {
            case SF_FORMAT_PCM_16:
                for (size_t i = 0; i < m; i++) ptr[i] = (float)(int16_t)rd16(b + 2 * i) * (1.0f / 32768.0f);
                break;
            case SF_FORMAT_PCM_24:
                for (size_t i = 0; i < m; i++) {
                    // three bytes into the top of an int32, normalised by 2^31 like libsndfile's tribyte path
                    const int32_t v = (int32_t)(((uint32_t)b[3 * i] << 8) | ((uint32_t)b[3 * i + 1] << 16) |
                                                ((uint32_t)b[3 * i + 2] << 24));
                    ptr[i] = (float)v * (1.0f / 2147483648.0f);
                }
                break;
            case SF_FORMAT_PCM_32:
                for (size_t i = 0; i < m; i++) ptr[i] = (float)(int32_t)rd32(b + 4 * i) * (1.0f / 2147483648.0f);
                break;
            default:
                for (size_t i = 0; i < m; i++) {
                    const uint32_t u = rd32(b + 4 * i);
                    memcpy(&ptr[i], &u, 4);
                }
        }
    }
    s->pos += frames;
    return frames;
}

static inline long quant(float x, float scale, long lo, long hi, bool clip) {
    const float v = x * scale;
    if (clip) {
        if (v >= (float)hi) return hi;
        if (v <= (float)lo) return lo;
    }
    return lrintf(v);
}

extern "C" sf_count_t sf_writef_float(SNDFILE *s, const float *ptr, sf_count_t frames) {
    if (!s || s->mode != SFM_WRITE || frames <= 0) return 0;
    const size_t n = (size_t)frames * (size_t)s->info.channels;
    if (s->virt) {
        virt_header(s);
        const int bytes = sub_bits(s->subformat) / 8;
        s->scratch.resize(n * (size_t)bytes);
        unsigned char *o = s->scratch.data();
        for (size_t i = 0; i < n; i++, o += bytes) {
            if (s->subformat == SF_FORMAT_FLOAT) { memcpy(o, ptr + i, 4); continue; }
            long q;
            if (bytes == 2) q = quant(ptr[i], 32767.0f, -32768, 32767, s->clipping);
            else if (bytes == 3) q = quant(ptr[i], 8388607.0f, -8388608, 8388607, s->clipping);
            else q = quant(ptr[i], 2147483647.0f, INT32_MIN, INT32_MAX, s->clipping);
            for (int b = 0; b < bytes; b++) o[b] = (unsigned char)((unsigned long)q >> (8 * b));
        }
        s->vio.write(s->scratch.data(), (sf_count_t)s->scratch.size(), s->vuser);
        s->pos += frames;
        s->info.frames = s->pos;
        return frames;
    }
    if (s->discard) {
        // keep the conversion cost of the format, drop the result
        long acc = 0;
        if (s->subformat == SF_FORMAT_PCM_16) for (size_t i = 0; i < n; i++) acc += quant(ptr[i], 32767.0f, -32768, 32767, s->clipping);
        else if (s->subformat == SF_FORMAT_PCM_24) for (size_t i = 0; i < n; i++) acc += quant(ptr[i], 8388607.0f, -8388608, 8388607, s->clipping);
        s->scratch.resize(sizeof(long));
        memcpy(s->scratch.data(), &acc, sizeof(long));
        s->pos += frames;
        s->info.frames = s->pos;
        return frames;
    }
    switch (s->subformat) {
        case SF_FORMAT_PCM_16:
            for (size_t i = 0; i < n; i++) s->w16.push_back((int16_t)quant(ptr[i], 32767.0f, -32768, 32767, s->clipping));
            break;
        case SF_FORMAT_PCM_24:
            for (size_t i = 0; i < n; i++) s->w32.push_back((int32_t)quant(ptr[i], 8388607.0f, -8388608, 8388607, s->clipping));
            break;
        case SF_FORMAT_PCM_32:
            for (size_t i = 0; i < n; i++) s->w32.push_back((int32_t)quant(ptr[i], 2147483647.0f, INT32_MIN, INT32_MAX, s->clipping));
            break;
        default:
            s->wf.insert(s->wf.end(), ptr, ptr + n);
    }
    s->pos += frames;
    s->info.frames = s->pos;
    return frames;
}

extern "C" sf_count_t sf_readf_short(SNDFILE *s, short *ptr, sf_count_t frames) {
    if (!s || s->mode != SFM_READ || frames <= 0 || s->subformat != SF_FORMAT_PCM_16) return 0;
    const sf_count_t left = s->info.frames - s->pos;
    if (frames > left) frames = left;
    if (frames <= 0) return 0;
    const size_t ch = (size_t)s->info.channels;
    if (s->mem) {
        memcpy(ptr, (const int16_t *)s->mem + (size_t)s->pos * ch, (size_t)frames * ch * sizeof(int16_t));
    } else {
        const size_t got = fread(ptr, sizeof(int16_t) * ch, (size_t)frames, s->fp);   // little-endian host
        frames = (sf_count_t)got;
    }
    s->pos += frames;
    return frames;
}

extern "C" sf_count_t sf_writef_short(SNDFILE *s, const short *ptr, sf_count_t frames) {
    if (!s || s->mode != SFM_WRITE || frames <= 0 || s->subformat != SF_FORMAT_PCM_16) return 0;
    const size_t n = (size_t)frames * (size_t)s->info.channels;
    if (s->discard) {
        long acc = 0;
        for (size_t i = 0; i < n; i += 64) acc += ptr[i];   // touch the block like a consumer would
        s->scratch.resize(sizeof(long));
        memcpy(s->scratch.data(), &acc, sizeof(long));
    } else {
        s->w16.insert(s->w16.end(), ptr, ptr + n);
    }
    s->pos += frames;
    s->info.frames = s->pos;
    return frames;
}

extern "C" sf_count_t sf_shim_memory_frames(SNDFILE *s) { return s ? s->pos : 0; }

extern "C" const void *sf_shim_memory_data(SNDFILE *s) {
    if (!s || s->mode != SFM_WRITE) return nullptr;
    switch (s->subformat) {
        case SF_FORMAT_PCM_16: return s->w16.data();
        case SF_FORMAT_PCM_24:
        case SF_FORMAT_PCM_32: return s->w32.data();
        default: return s->wf.data();
    }
}

extern "C" int sf_command(SNDFILE *s, int command, void *data, int datasize) {
    (void)data;
    switch (command) {
        case SFC_SET_CLIPPING:
            if (s) s->clipping = datasize != 0;
            return s && s->clipping;
        case SFC_GET_CLIPPING: return s && s->clipping;
        case SFC_WAVEX_GET_AMBISONIC: return SF_AMBISONIC_NONE;
        case SFC_UPDATE_HEADER_NOW:
            if (s) virt_header(s);
            return 0;
        default: return 0;
    }
}

extern "C" const char *sf_strerror(SNDFILE *) { return g_error.c_str(); }
extern "C" const char *sf_version_string(void) { return "libsndfile-shim-folve_b200"; }
