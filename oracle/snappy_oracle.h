/*
 * oracle/snappy_oracle.h -- TEST INFRASTRUCTURE ONLY (never linked into libhap_b200.so).
 *
 * CPU restatement of the raw Snappy block format that the reference reaches through
 * "snappy-c.h" (/root/reference/source/hap.c:32; call sites hap.c:313, 453, 612, 813, 890, 899).
 * Snappy itself is NOT under /root/reference (un-vendored, version unpinned), so this
 * file restates the published format description (HapVideoDRAFT.md:23,146 cites it) and is pinned
 * against genuine Google Snappy as bundled in pyarrow 24 (tests/test_oracle_snappy.py) and against
 * the reference's own output (tests/golden/).
 *
 * The four entry points mirror the snappy-c.h functions the reference calls, with an orc_ prefix.
 */
#ifndef ORACLE_SNAPPY_ORACLE_H
#define ORACLE_SNAPPY_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_SNAPPY_OK = 0, ORC_SNAPPY_INVALID_INPUT = 1, ORC_SNAPPY_BUFFER_TOO_SMALL = 2 };

/* 32 + n + n/6 : the bound hap.c:313 multiplies by the chunk count */
size_t orc_snappy_max_compressed_length(size_t n);

/* varint32 preamble; INVALID_INPUT when truncated or wider than 32 bits */
int orc_snappy_uncompressed_length(const uint8_t *in, size_t n, size_t *result);

/* *out_len: in = capacity, out = bytes produced */
int orc_snappy_compress(const uint8_t *in, size_t n, uint8_t *out, size_t *out_len);
int orc_snappy_uncompress(const uint8_t *in, size_t n, uint8_t *out, size_t *out_len);

/* statistics walker used by tests: counts elements of each kind without producing output */
typedef struct orc_snappy_stats {
    uint64_t literals, literal_bytes;
    uint64_t copy1, copy2, copy4, copy_bytes;
    uint64_t overlapping;    /* copies with offset < length */
    uint32_t max_offset;
} orc_snappy_stats;
int orc_snappy_scan(const uint8_t *in, size_t n, orc_snappy_stats *st);

#ifdef __cplusplus
}
#endif
#endif
