// filter-config.cc -- see filter-config.h.  Behavioural mirror of
// /root/reference/zita-config.cc:55-378 and zita-fconfig.cc:38-109; every
// decision that is visible in the loaded filter cites the line it follows.
//
// The grammar, the error behaviour and the arithmetic order of the impulse commands follow
// jconvolver 0.9.2's config.cc as adapted in folve (zita-config.cc, zita-fconfig.cc, zita-sstring.cc):
// Copyright (C) 2006-2011 Fons Adriaensen <fons@linuxaudio.org>, Copyright (C) 2012 Henner Zeller
// <h.zeller@acm.org>; GNU General Public License, version 2 or later (zita-config) / 3 or later (folve).
// This file is distributed under the GPL, version 3 or later (COPYING).
#include "filter-config.h"

#include <ctype.h>
#include <math.h>
#include <sndfile.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <syslog.h>

#include <string>
#include <vector>

namespace folve_b200 {

namespace {

const unsigned kReadChunkFrames = 0x4000;  // BSIZE, zita-config.cc:43
const int kLineMax = 1024;                 // zita-config.cc:286

struct Parser {
    FilterConfig *cfg;
    const char *file;
    int line_no;
    std::string dir;  // where relative impulse files are looked up
};

// zita-config.cc:46-52
int CheckInOut(const Parser &ps, int ip, int op) {
    if (!ps.cfg->size) return CFG_ERR_NOCONV;
    if (ip < 1 || ip > ps.cfg->ninp) return CFG_ERR_IONUM;
    if (op < 1 || op > ps.cfg->nout) return CFG_ERR_IONUM;
    return CFG_NOERR;
}

// /convolver/new -- zita-fconfig.cc:38-97
int CmdConvolverNew(Parser &ps, const char *args) {
    FilterConfig *cfg = ps.cfg;
    unsigned ninp = 0, nout = 0, part = 0, size = 0;
    float dens = 0;
    // the reference scans straight into cfg->ninp / nout / size, so even a
    // failing line leaves them modified
    const int got = sscanf(args, "%u %u %u %u %f", &ninp, &nout, &part, &size, &dens);
    if (got >= 1) cfg->ninp = (int)ninp;
    if (got >= 2) cfg->nout = (int)nout;
    if (got >= 4) cfg->size = (int)size;
    if (got < 4) return CFG_ERR_PARAM;
    if (got < 5) dens = 0;
    (void)part;  // parsed and never used (zita-fconfig.cc:40,44)

    if (cfg->ninp == 0 || cfg->ninp > FCV_MAXINP) {
        syslog(LOG_ERR, "%s:%d: Number of inputs (%d) is out of range.\n", ps.file, ps.line_no, cfg->ninp);
        return CFG_ERR_OTHER;
    }
    if (cfg->nout == 0 || cfg->nout > FCV_MAXOUT) {
        syslog(LOG_ERR, "%s:%d: Number of outputs (%d) is out of range.\n", ps.file, ps.line_no, cfg->nout);
        return CFG_ERR_OTHER;
    }
    if (size > FCV_MAXSIZE) {
        syslog(LOG_ERR, "%s:%d: Convolver size (%d) is out of range.\n", ps.file, ps.line_no, cfg->size);
        return CFG_ERR_OTHER;
    }
    if (dens < 0.0f || dens > 1.0f) {
        syslog(LOG_ERR, "%s:%d: Density parameter is out of range.\n", ps.file, ps.line_no);
        return CFG_ERR_OTHER;
    }
    cfg->fragm = FragmForSize(size);
    if (cfg->filter) {
        // A second /convolver/new: Convproc::configure refuses (not idle) and the
        // reference reports "Can't initialise convolution engine" as ERR_OTHER, which
        // ends the parse "successfully" with the SECOND line's ninp / nout / size / fragm
        // in the config and the FIRST line's engine behind it.  SoundProcessor::Create
        // (sound-processor.cc:44-46) then returns NULL if the second line names more
        // inputs or outputs than the first, and otherwise builds a processor whose
        // block size and channel counts disagree with its Convproc.  Here such a file
        // never yields a processor: the filter is dropped.
        syslog(LOG_ERR, "Can't initialise convolution engine\n");
        fcv_filter_unref(cfg->filter);
        cfg->filter = nullptr;
        return CFG_ERR_OTHER;
    }
    cfg->filter = fcv_filter_begin(cfg->ninp, cfg->nout, size, (unsigned)cfg->fragm);
    if (!cfg->filter) {
        syslog(LOG_ERR, "Can't initialise convolution engine\n");
        return CFG_ERR_OTHER;
    }
    return CFG_NOERR;
}

// /impulse/read -- zita-config.cc:55-177
int CmdImpulseRead(Parser &ps, const char *args) {
    FilterConfig *cfg = ps.cfg;
    unsigned ip1, op1, delay, offset, length, ichan;
    float gain;
    int used = 0;
    char name[kLineMax];
    if (sscanf(args, "%u %u %f %u %u %u %u %n", &ip1, &op1, &gain, &delay, &offset, &length, &ichan, &used) != 7)
        return CFG_ERR_PARAM;
    if (!ScanString(args + used, name, kLineMax)) return CFG_ERR_PARAM;
    // cfg->latency is always 0 in folve (sound-processor.cc:37): no latency compensation.
    const int err = CheckInOut(ps, (int)ip1, (int)op1);
    if (err) return err;

    const std::string path = (name[0] == '/') ? std::string(name) : ps.dir + "/" + name;
    cfg->impulse_files.push_back(StampFile(path));
    SF_INFO info;
    memset(&info, 0, sizeof(info));
    SNDFILE *snd = sf_open(path.c_str(), SFM_READ, &info);
    if (!snd) {
        syslog(LOG_ERR, "%s:%d: Unable to open '%s' >%s<.\n", ps.file, ps.line_no, path.c_str(), ps.dir.c_str());
        return CFG_ERR_OTHER;
    }
    if (info.samplerate != cfg->fsamp)  // only a warning (zita-config.cc:108-112)
        syslog(LOG_ERR, "%s:%d: Sample rate (%d) of '%s' does not match.\n", ps.file, ps.line_no, info.samplerate,
               path.c_str());
    const unsigned nchan = (unsigned)info.channels;
    const unsigned nfram_file = (unsigned)info.frames;  // Audiofile::_size is uint32_t
    if (ichan < 1 || ichan > nchan) {
        syslog(LOG_ERR, "%s:%d: Channel not available.\n", ps.file, ps.line_no);
        sf_close(snd);
        return CFG_ERR_OTHER;
    }
    if (offset && sf_seek(snd, offset, SEEK_SET) != (sf_count_t)offset) {
        syslog(LOG_ERR, "%s:%d: Can't seek to offset.\n", ps.file, ps.line_no);
        sf_close(snd);
        return CFG_ERR_OTHER;
    }
    if (!length) length = nfram_file - offset;
    if (length > (unsigned)cfg->size - delay) {  // unsigned arithmetic as in the reference
        length = (unsigned)cfg->size - delay;
        syslog(LOG_ERR, "%s:%d: Data truncated.\n", ps.file, ps.line_no);
    }
    std::vector<float> buff;
    try {
        buff.resize((size_t)kReadChunkFrames * nchan);
    } catch (...) {
        sf_close(snd);
        return CFG_ERR_ALLOC;
    }
    while (length) {
        const unsigned want = length > kReadChunkFrames ? kReadChunkFrames : length;
        const int nfram = (int)sf_readf_float(snd, buff.data(), want);
        if (nfram < 0) {
            syslog(LOG_ERR, "%s:%d: Error reading file.\n", ps.file, ps.line_no);
            sf_close(snd);
            return CFG_ERR_OTHER;
        }
        if (nfram == 0) break;  // short file: the reference would spin here forever
        float *p = buff.data() + (ichan - 1);
        for (int i = 0; i < nfram; i++) p[(size_t)i * nchan] *= gain;  // float32 gain, zita-config.cc:161-162
        if (fcv_filter_add(cfg->filter, (int)ip1 - 1, (int)op1 - 1, (int)nchan, p, (int)delay, (int)delay + nfram)) {
            sf_close(snd);
            return CFG_ERR_ALLOC;
        }
        delay += (unsigned)nfram;
        length -= (unsigned)nfram;
    }
    sf_close(snd);
    return CFG_NOERR;
}

// /impulse/dirac -- zita-config.cc:180-209
int CmdImpulseDirac(Parser &ps, const char *args) {
    int ip1, op1, delay;
    float gain;
    if (sscanf(args, "%u %u %f %u", (unsigned *)&ip1, (unsigned *)&op1, &gain, (unsigned *)&delay) != 4)
        return CFG_ERR_PARAM;
    const int err = CheckInOut(ps, ip1, op1);
    if (err) return err;
    if (delay < 0) {  // delay < latency (== 0)
        syslog(LOG_ERR, "%s:%d: Dirac pulse removed: delay < latency.\n", ps.file, ps.line_no);
        return CFG_NOERR;
    }
    if (delay < ps.cfg->size) {
        if (fcv_filter_add(ps.cfg->filter, ip1 - 1, op1 - 1, 1, &gain, delay, delay + 1)) return CFG_ERR_ALLOC;
    }
    return CFG_NOERR;
}

// /impulse/hilbert -- zita-config.cc:212-259
int CmdImpulseHilbert(Parser &ps, const char *args) {
    unsigned ip1, op1, delay, length;
    float gain;
    if (sscanf(args, "%u %u %f %u %u", &ip1, &op1, &gain, &delay, &length) != 5) return CFG_ERR_PARAM;
    const int err = CheckInOut(ps, (int)ip1, (int)op1);
    if (err) return err;
    if (length < 64 || length > 65536) return CFG_ERR_PARAM;
    if (delay < length / 2) {
        syslog(LOG_ERR, "%s:%d: Hilbert impulse removed: delay < latency + lenght / 2.\n", ps.file, ps.line_no);
        return CFG_NOERR;
    }
    delay -= length / 2;
    std::vector<float> taps(length, 0.0f);
    // Same arithmetic types as the reference: gain scaled in double and rounded
    // to float once; window evaluated with cosf on the double argument rounded
    // to float; odd taps only, antisymmetric around h = length / 2.
    gain *= 2 / M_PI;
    const unsigned h = length / 2;
    for (unsigned i = 1; i < h; i += 2) {
        float v = gain / i;
        const float w = 0.43f + 0.57f * cosf(i * M_PI / h);
        v *= w;
        taps[h + i] = -v;
        taps[h - i] = v;
    }
    if (fcv_filter_add(ps.cfg->filter, (int)ip1 - 1, (int)op1 - 1, 1, taps.data(), (int)delay, (int)(delay + length)))
        return CFG_ERR_ALLOC;
    return CFG_NOERR;
}

// /impulse/copy -- zita-config.cc:262-279
int CmdImpulseCopy(Parser &ps, const char *args) {
    unsigned ip1, op1, ip2, op2;
    if (sscanf(args, "%u %u %u %u", &ip1, &op1, &ip2, &op2) != 4) return CFG_ERR_PARAM;
    const int err = CheckInOut(ps, (int)ip1, (int)op1) | CheckInOut(ps, (int)ip2, (int)op2);
    if (err) return err;
    if (ip1 == ip2 && op1 == op2) return CFG_ERR_PARAM;
    // destination (ip1,op1) uses the spectra of source (ip2,op2)
    if (fcv_filter_link(ps.cfg->filter, (int)ip2 - 1, (int)op2 - 1, (int)ip1 - 1, (int)op1 - 1)) return CFG_ERR_ALLOC;
    return CFG_NOERR;
}

void LogError(const Parser &ps, int stat) {
    const char *what = "Unknown error.";
    switch (stat) {
        case CFG_ERR_SYNTAX: what = "Syntax error."; break;
        case CFG_ERR_PARAM: what = "Bad or missing parameters."; break;
        case CFG_ERR_ALLOC: what = "Out of memory."; break;
        case CFG_ERR_CANTCD: what = "Can't change directory."; break;
        case CFG_ERR_COMMAND: what = "Unknown command."; break;
        case CFG_ERR_NOCONV: what = "No convolver yet defined."; break;
        case CFG_ERR_IONUM: what = "Bad input or output number."; break;
    }
    syslog(LOG_ERR, "%s:%d: %s\n", ps.file, ps.line_no, what);
}

}  // namespace

FileStamp StampFile(const std::string &path) {
    FileStamp s;
    s.path = path;
    struct stat st;
    if (stat(path.c_str(), &st) == 0) {
        s.sec = st.st_mtim.tv_sec;
        s.nsec = st.st_mtim.tv_nsec;
        s.size = (long long)st.st_size;
    }
    return s;
}

bool StampsCurrent(const std::vector<FileStamp> &stamps) {
    for (const FileStamp &s : stamps)
        if (!(StampFile(s.path) == s)) return false;
    return true;
}

int FragmForSize(unsigned size) {
    // zita-fconfig.cc:74-77: start at Convproc::MAXQUANT, halve while above
    // MINPART and at least twice the filter length.
    unsigned fragm = FCV_MAXQUANT;
    while (fragm > FCV_MINPART && fragm >= 2 * size) fragm /= 2;
    return (int)fragm;
}

int ScanString(const char *src, char *dest, int size) {
    // States: leading blanks, bare word, inside '...' or "...", after a backslash.
    if (size < 0) return 0;
    int in = 0, out = 0;
    char quote = 0;
    bool escaped = false;
    for (;;) {
        if (out == size) break;  // no room for the terminator: error
        unsigned char c = (unsigned char)src[in++];
        if (isblank(c)) c = ' ';
        if (iscntrl(c)) {  // includes NUL and newline: ends the input
            if (quote || escaped) break;
            dest[out] = 0;
            return in - 1;
        }
        if (escaped) {
            dest[out++] = (char)c;
            escaped = false;
        } else if (c == '\\') {
            if (quote == '\'') dest[out++] = (char)c;  // no escapes inside single quotes
            else escaped = true;
        } else if (c == '\'' || c == '"') {
            if (c == quote) {  // closing quote
                dest[out] = 0;
                return in;
            }
            if (quote || out) break;  // a different quote inside quotes, or a quote inside a word
            quote = (char)c;
        } else if (c == ' ') {
            if (quote) dest[out++] = ' ';
            else if (out) {  // end of a bare word
                dest[out] = 0;
                return in - 1;
            }
        } else {
            dest[out++] = (char)c;
        }
    }
    dest[0] = 0;
    return 0;
}

int LoadFilterConfig(FilterConfig *cfg, const char *config_file) {
    FILE *fp = fopen(config_file, "r");
    if (!fp) {
        syslog(LOG_ERR, "Can't open '%s' for reading\n", config_file);
        return -1;
    }
    Parser ps;
    ps.cfg = cfg;
    ps.file = config_file;
    ps.line_no = 0;
    {  // impulse files are relative to the directory of the config file (zita-config.cc:296-299)
        const std::string f(config_file);
        const size_t slash = f.find_last_of('/');
        if (slash == std::string::npos) ps.dir = ".";
        else if (slash == 0) ps.dir = "/";
        else ps.dir = f.substr(0, slash);
    }
    int stat = CFG_NOERR;
    char line[kLineMax];
    while (!stat && fgets(line, kLineMax, fp)) {
        ps.line_no++;
        char *p = line;
        if (*p != '/') {
            // only blank lines and '#' comments are allowed between commands
            while (isspace((unsigned char)*p)) p++;
            if (*p > ' ' && *p != '#') stat = CFG_ERR_SYNTAX;
            continue;
        }
        char *q = p;
        while (*q >= ' ' && !isspace((unsigned char)*q)) q++;  // end of the command word
        *q++ = 0;
        while (*q >= ' ' && isspace((unsigned char)*q)) q++;   // start of the arguments
        const std::string cmd(p);
        if (cmd == "/cd") {
            char tmp[kLineMax];
            if (ScanString(q, tmp, kLineMax) == 0) stat = CFG_ERR_PARAM;
            if (tmp[0] == '/') ps.dir = tmp;
            else ps.dir = ps.dir + "/" + tmp;
        } else if (cmd == "/convolver/new") stat = CmdConvolverNew(ps, q);
        else if (cmd == "/impulse/read") stat = CmdImpulseRead(ps, q);
        else if (cmd == "/impulse/dirac") stat = CmdImpulseDirac(ps, q);
        else if (cmd == "/impulse/hilbert") stat = CmdImpulseHilbert(ps, q);
        else if (cmd == "/impulse/copy") stat = CmdImpulseCopy(ps, q);
        else if (cmd == "/input/name" || cmd == "/output/name") stat = CFG_NOERR;  // zita-fconfig.cc:100-109
        else stat = CFG_ERR_COMMAND;
    }
    fclose(fp);
    if (stat == CFG_ERR_OTHER) stat = CFG_NOERR;  // zita-config.cc:345
    if (stat) {
        LogError(ps, stat);
        if (cfg->filter) fcv_filter_unref(cfg->filter);
        cfg->filter = nullptr;
    }
    return stat;
}

}  // namespace folve_b200
