/* oracle/ref_shim/snappy_shim_orc.c -- TEST INFRASTRUCTURE ONLY.
 * snappy-c.h provider over oracle/snappy_oracle.c, used when libarrow cannot be linked. */
#include "snappy-c.h"
#include "../snappy_oracle.h"
#include <stdint.h>
size_t snappy_max_compressed_length(size_t n) { return orc_snappy_max_compressed_length(n); }
snappy_status snappy_uncompressed_length(const char *in, size_t n, size_t *result)
{ return (snappy_status)orc_snappy_uncompressed_length((const uint8_t *)in, n, result); }
snappy_status snappy_compress(const char *in, size_t n, char *out, size_t *out_len)
{ return (snappy_status)orc_snappy_compress((const uint8_t *)in, n, (uint8_t *)out, out_len); }
snappy_status snappy_uncompress(const char *in, size_t n, char *out, size_t *out_len)
{ return (snappy_status)orc_snappy_uncompress((const uint8_t *)in, n, (uint8_t *)out, out_len); }
