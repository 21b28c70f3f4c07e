// filter-config.h -- loader for jconvolver / zita-config style filter files.
//
// Host-side mirror of /root/reference/zita-config.{h,cc} + zita-fconfig.cc: the
// same grammar, the same `fragm` rule and the same truncation / accumulation /
// link semantics, but the impulse data is fed to the B200 engine's C ABI
// (fcv_filter_add / fcv_filter_link, include/folve_b200.h) instead of
// Convproc::impdata_create / impdata_copy.
//
// Commands (README.CONFIG.txt:22-97):
//   /convolver/new  <in> <out> <partition> <size> [density]
//   /impulse/read   <in> <out> <gain> <delay> <offset> <length> <chan> <file>
//   /impulse/dirac  <in> <out> <gain> <delay>
//   /impulse/hilbert <in> <out> <gain> <delay> <length>
//   /impulse/copy   <in> <out> <from in> <from out>
//   /cd <path>      /input/name ...   /output/name ...   # comments
//
// The grammar, the error behaviour and the arithmetic order of the impulse commands follow
// jconvolver 0.9.2's config.cc as adapted in folve (zita-config.cc, zita-fconfig.cc, zita-sstring.cc):
// Copyright (C) 2006-2011 Fons Adriaensen <fons@linuxaudio.org>, Copyright (C) 2012 Henner Zeller
// <h.zeller@acm.org>; GNU General Public License, version 2 or later (zita-config) / 3 or later (folve).
// This file is distributed under the GPL, version 3 or later (COPYING).
#ifndef FOLVE_B200_FILTER_CONFIG_H
#define FOLVE_B200_FILTER_CONFIG_H

#include <time.h>

#include <string>
#include <vector>

#include "../../include/folve_b200.h"

namespace folve_b200 {

// Identity of a file at the moment it was read: nanosecond mtime and size (all zero if it
// could not be stat'ed).  The reference compares the config file's mtime in whole seconds and
// never looks at the impulse files (TODO at sound-processor.cc:130-131); here the config and every
// /impulse/read file are stamped, so an edited or replaced impulse response is noticed.
struct FileStamp {
    std::string path;
    time_t sec = 0;
    long nsec = 0;
    long long size = 0;
    bool operator==(const FileStamp &o) const {
        return path == o.path && sec == o.sec && nsec == o.nsec && size == o.size;
    }
};
FileStamp StampFile(const std::string &path);
// true if every file still has the stamp it was recorded with
bool StampsCurrent(const std::vector<FileStamp> &stamps);

// Same numeric values as the enum in /root/reference/zita-config.h:51.
enum ConfigError {
    CFG_NOERR, CFG_ERR_OTHER, CFG_ERR_SYNTAX, CFG_ERR_PARAM, CFG_ERR_ALLOC, CFG_ERR_CANTCD,
    CFG_ERR_COMMAND, CFG_ERR_NOCONV, CFG_ERR_IONUM
};

// What the reference keeps in struct ZitaConfig (zita-config.h:37-49).
struct FilterConfig {
    fcv_filter *filter = nullptr;  // built but NOT committed; NULL if no /convolver/new succeeded
    int fsamp = 0;
    int fragm = 0;
    int ninp = 0;
    int nout = 0;
    int size = 0;
    // every file named by an /impulse/read line that was reached (missing ones included), stamped
    // BEFORE it was opened: a replacement racing the load is seen as a change next time
    std::vector<FileStamp> impulse_files;
};

// Parses `config_file`.  Returns 0 on success (also when parsing stopped at an
// "other" error such as a missing impulse file -- zita-config.cc:345 swallows
// those), -1 if the file cannot be opened, or a ConfigError.  On a non-zero
// return cfg->filter is released and set to NULL.
int LoadFilterConfig(FilterConfig *cfg, const char *config_file);

// The block size rule of zita-fconfig.cc:74-77.
int FragmForSize(unsigned size);

// Quoted-string scanner with the semantics documented in zita-sstring.h:26-43.
// Returns the number of characters consumed, 0 on error.
int ScanString(const char *src, char *dest, int size);

}  // namespace folve_b200

#endif
