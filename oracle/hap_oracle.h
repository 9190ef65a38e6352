/*
 * oracle/hap_oracle.h -- TEST INFRASTRUCTURE ONLY (never linked into libhap_b200.so).
 *
 * CPU restatement of the Hap frame container algorithm of /root/reference/source/hap.c, one
 * function per public entry point of /root/reference/source/hap.h:76-152, with an orc_ prefix.
 * Argument meaning and HapResult codes follow hap.h:55-61.  Second-stage compression goes through
 * oracle/snappy_oracle.c.  Pinned by tests/test_oracle_container.py against the unmodified
 * reference built into oracle/_ref/ and against the known-answer frames in tests/golden/.
 */
#ifndef ORACLE_HAP_ORACLE_H
#define ORACLE_HAP_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef void (*orc_work_fn)(void *p, unsigned int index);
typedef void (*orc_decode_cb)(orc_work_fn function, void *p, unsigned int count, void *info);

unsigned long orc_HapMaxEncodedLength(unsigned int count, unsigned long *lengths,
                                      unsigned int *textureFormats, unsigned int *chunkCounts);

unsigned int orc_HapEncode(unsigned int count, const void **inputBuffers,
                           unsigned long *inputBuffersBytes, unsigned int *textureFormats,
                           unsigned int *compressors, unsigned int *chunkCounts, void *outputBuffer,
                           unsigned long outputBufferBytes, unsigned long *outputBufferBytesUsed);

unsigned int orc_HapDecode(const void *inputBuffer, unsigned long inputBufferBytes,
                           unsigned int index, orc_decode_cb callback, void *info,
                           void *outputBuffer, unsigned long outputBufferBytes,
                           unsigned long *outputBufferBytesUsed,
                           unsigned int *outputBufferTextureFormat);

unsigned int orc_HapGetFrameTextureCount(const void *inputBuffer, unsigned long inputBufferBytes,
                                         unsigned int *outputTextureCount);
unsigned int orc_HapGetFrameTextureFormat(const void *inputBuffer, unsigned long inputBufferBytes,
                                          unsigned int index,
                                          unsigned int *outputBufferTextureFormat);
unsigned int orc_HapGetFrameTextureChunkCount(const void *inputBuffer,
                                              unsigned long inputBufferBytes, unsigned int index,
                                              int *chunk_count);

/* hap.c:277-300, exposed for the chunk-limiter known-answer tests (SURVEY.md KAT-E) */
unsigned int orc_hap_limited_chunk_count(unsigned long bytes, unsigned int textureFormat,
                                         unsigned int chunkCount);

#ifdef __cplusplus
}
#endif
#endif
