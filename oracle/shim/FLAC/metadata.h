/* FLAC/metadata.h -- TEST INFRASTRUCTURE.  Stand-in for libFLAC's header (not installed here)
 * with the three constants /root/reference/convolve-file-handler.cc:455,462,486 uses when it
 * copies a FLAC header: the metadata block type numbers of the FLAC format specification. */
#ifndef FOLVE_B200_SHIM_FLAC_METADATA_H
#define FOLVE_B200_SHIM_FLAC_METADATA_H
enum {
    FLAC__METADATA_TYPE_STREAMINFO = 0,
    FLAC__METADATA_TYPE_PADDING = 1,
    FLAC__METADATA_TYPE_APPLICATION = 2,
    FLAC__METADATA_TYPE_SEEKTABLE = 3,
    FLAC__METADATA_TYPE_VORBIS_COMMENT = 4
};
#endif
