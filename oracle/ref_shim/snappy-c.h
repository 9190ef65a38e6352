/*
 * oracle/ref_shim/snappy-c.h -- TEST INFRASTRUCTURE ONLY.
 * Declares exactly the four snappy-c functions the unmodified reference uses
 * (/root/reference/source/hap.c:32, call sites :313 :453 :612 :813 :890 :899) so that hap.c can be
 * compiled where it lies.  Two providers exist: snappy_shim_arrow.cc (genuine Google Snappy as
 * bundled in pyarrow's libarrow) and snappy_shim_orc.c (oracle/snappy_oracle.c).
 */
#ifndef HAP_ORACLE_SNAPPY_C_H
#define HAP_ORACLE_SNAPPY_C_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef enum { SNAPPY_OK = 0, SNAPPY_INVALID_INPUT = 1, SNAPPY_BUFFER_TOO_SMALL = 2 } snappy_status;
snappy_status snappy_compress(const char *input, size_t input_length, char *compressed, size_t *compressed_length);
snappy_status snappy_uncompress(const char *compressed, size_t compressed_length, char *uncompressed, size_t *uncompressed_length);
size_t snappy_max_compressed_length(size_t source_length);
snappy_status snappy_uncompressed_length(const char *compressed, size_t compressed_length, size_t *result);
#ifdef __cplusplus
}
#endif
#endif
