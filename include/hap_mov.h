/*
 * include/hap_mov.h -- QuickTime (.mov) sample-table reader / writer for Hap streams.
 *
 * The reference codec works on single frames and leaves containers to its callers
 * (/root/reference/documentation/HapVideoDRAFT.md:14: "Hap frames are... stored in any container");
 * the FourCCs a container uses for the Hap flavours are listed in HapVideoDRAFT.md:132-142.  This is
 * the "next" row of SURVEY.md 8(f): the smallest container layer that lets the frame codec of hap.h /
 * hap_b200.h ingest and produce real files: one video track, frames stored verbatim as samples.
 * Host code only (no GPU work): plain C ABI, results are HapResult values (hap.h).
 */
#ifndef hap_mov_h
#define hap_mov_h
#include "hap.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct HapB200Mov HapB200Mov;

/* 'Hap1' 'Hap5' 'HapY' 'HapM' 'HapA' 'Hap7' 'HapH' as big-endian integers ('H' in the top byte) */
#define HAPB200_FOURCC(a, b, c, d) (((unsigned int)(a) << 24) | ((unsigned int)(b) << 16) | ((unsigned int)(c) << 8) | (unsigned int)(d))

/* The FourCC for a Hap frame, from the texture formats its header declares (HapVideoDRAFT.md:132-142).
 * Bad_Frame when the frame is not a Hap frame or the combination of textures has no name. */
unsigned int HapB200MovFourCCForFrame(const void *frame, unsigned long frameBytes, unsigned int *fourcc);

/* ---- reading ---- */
/* NULL when the file cannot be opened, is not a QuickTime / ISO-BMFF movie, has no video track whose sample
 * description is one of the Hap FourCCs, or its sample tables are inconsistent or point outside the file. */
HapB200Mov *HapB200MovOpen(const char *path);
unsigned int HapB200MovInfo(const HapB200Mov *mov, unsigned int *fourcc, unsigned int *width, unsigned int *height,
                            unsigned long *frameCount, unsigned int *timescale, unsigned long *duration);
/* size of frame `index` in bytes; 0 when out of range */
unsigned long HapB200MovFrameBytes(const HapB200Mov *mov, unsigned long index);
/* copies frame `index` into buffer: Bad_Arguments (index), Buffer_Too_Small, Internal_Error (I/O) */
unsigned int HapB200MovReadFrame(HapB200Mov *mov, unsigned long index, void *buffer, unsigned long bufferBytes,
                                 unsigned long *bytesUsed, unsigned int *durationTicks);

/* ---- writing ---- */
/* Creates `path` (truncating it): one video track of `fourcc`, width x height, time in 1/timescale seconds. */
HapB200Mov *HapB200MovCreate(const char *path, unsigned int fourcc, unsigned int width, unsigned int height,
                             unsigned int timescale);
/* appends one frame (a sample) that lasts durationTicks */
unsigned int HapB200MovWriteFrame(HapB200Mov *mov, const void *frame, unsigned long frameBytes, unsigned int durationTicks);

/* Closes either kind; a writer finishes the file here (sample tables, movie header).  NULL is allowed. */
unsigned int HapB200MovClose(HapB200Mov *mov);

#ifdef __cplusplus
}
#endif
#endif
