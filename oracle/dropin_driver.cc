// dropin_driver.cc -- TEST INFRASTRUCTURE.  C entry points over the reference's OWN filesystem
// layer: FolveFilesystem::GetOrCreateHandler, FileHandler::Read, ConversionBuffer::FillUntil,
// ConvolveFileHandler::AddMoreSoundData / PassoverProcessor, BufferThread, FileHandlerCache and
// ProcessorPool, all compiled UNMODIFIED from /root/reference (oracle/Makefile, targets
// `dropin` and `refstack`).  The same driver is linked twice:
//   _ref/libfolve_dropin.so    the reference's callers on top of THIS repository's
//                              sound-processor.h / sound-processor.cc / filter-config.cc and the
//                              CUDA engine -- the proof that the SoundProcessor replacement drops
//                              into folve unchanged;
//   _ref/libfolve_refstack.so  the same callers on the reference's own SoundProcessor over the
//                              restated zita-convolver -- what the first is compared with.
// A "mounted" file is read the way a media player reads it through FUSE: sequential read()
// calls of a fixed size until end of file.
#include <errno.h>
#include <string.h>
#include <sys/stat.h>

#include <string>
#include <vector>

#include "file-handler.h"
#include "folve-filesystem.h"

#if FOLVE_DROPIN_REFERENCE
// hooks the zita-convolver shim expects from whoever links it (see harness.cc)
extern "C" {
typedef void (*zo_shim_impdata_hook)(void *, unsigned, unsigned, int, const float *, int, int);
typedef void (*zo_shim_link_hook)(void *, unsigned, unsigned, unsigned, unsigned);
zo_shim_impdata_hook zo_shim_on_impdata = 0;
zo_shim_link_hook zo_shim_on_link = 0;
void *zo_shim_hook_user = 0;
int zo_shim_reset_is_fresh = 1;
}
#endif

extern "C" {

const char *dd_variant(void) {
#if FOLVE_DROPIN_REFERENCE
    return "refstack";
#else
    return "dropin";
#endif
}

// underlying_dir: the music; config_base_dir/<filter>/filter-*.conf: the filters.
void *dd_open(const char *underlying_dir, const char *config_base_dir, const char *filter, int gapless,
              int pre_buffer_bytes) {
    FolveFilesystem *fs = new FolveFilesystem();
    fs->set_underlying_dir(underlying_dir);
    fs->SetBaseConfigDir(config_base_dir);
    fs->set_gapless_processing(gapless != 0);
    fs->set_pre_buffer_size(pre_buffer_bytes);
    fs->set_initial_filter_config(filter);
    if (!fs->CheckInitialized()) {
        delete fs;
        return nullptr;
    }
    fs->SetupInitialConfig();
    return fs;
}

// Reads fs_path ("/album/01.wav") to its end in read_size-byte read() calls.  Returns the bytes
// delivered (the whole converted file, header included), or -1.  flags: bit 0 in_gapless, bit 1
// out_gapless, bit 2: the handler is a convolving one (has a filter).
long dd_read_file(void *h, const char *fs_path, unsigned char *dst, long cap, int read_size, int *flags,
                  float *max_output_value) {
    FolveFilesystem *fs = (FolveFilesystem *)h;
    FileHandler *fh = fs->GetOrCreateHandler(fs_path, false);
    if (!fh) return -1;
    long off = 0;
    std::vector<char> buf((size_t)read_size);
    for (;;) {
        const int r = fh->Read(buf.data(), (size_t)read_size, (off_t)off);
        if (r <= 0) break;
        if (off + r <= cap) memcpy(dst + off, buf.data(), (size_t)r);
        off += r;
    }
    HandlerStats st;
    fh->GetHandlerStatus(&st);
    if (flags) *flags = (st.in_gapless ? 1 : 0) | (st.out_gapless ? 2 : 0) | (st.filter_dir.empty() ? 0 : 4);
    if (max_output_value) *max_output_value = st.max_output_value;
    fs->Close(fs_path, fh);
    return off;
}

int dd_total_file_openings(void *h) { return ((FolveFilesystem *)h)->total_file_openings(); }

}  // extern "C"
