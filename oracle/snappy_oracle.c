/*
 * oracle/snappy_oracle.c -- TEST INFRASTRUCTURE ONLY. See snappy_oracle.h.
 *
 * Restates the raw Snappy block format:
 *   stream  = varint32(uncompressed_length) element*
 *   element = literal | copy1 | copy2 | copy4, selected by the low two bits of the tag byte.
 * The decoder is fully determined by the format.  The compressor is a plain greedy LZ77 with a
 * 4-byte hash over independent 64 KiB fragments, in the spirit of the public description of
 * Google's encoder; its exact bytes are NOT a parity target (the reference never pins them:
 * parity is "our frames decode in the reference / reference frames decode in ours").
 */
#include "snappy_oracle.h"
#include <string.h>

size_t orc_snappy_max_compressed_length(size_t n)
{
    return 32 + n + n / 6;
}

int orc_snappy_uncompressed_length(const uint8_t *in, size_t n, size_t *result)
{
    uint64_t v = 0;
    unsigned shift = 0;
    for (size_t i = 0; i < n && i < 5; i++) {
        uint8_t b = in[i];
        v |= (uint64_t)(b & 0x7F) << shift;
        if (!(b & 0x80)) {
            if (v > 0xFFFFFFFFull) return ORC_SNAPPY_INVALID_INPUT;
            *result = (size_t)v;
            return ORC_SNAPPY_OK;
        }
        shift += 7;
    }
    return ORC_SNAPPY_INVALID_INPUT;
}

static size_t varint_width(const uint8_t *in, size_t n)
{
    size_t i = 0;
    while (i < n && i < 5 && (in[i] & 0x80)) i++;
    return i + 1;
}

/* ---- element walker shared by decode and scan -------------------------------------------- */

typedef struct {
    int kind;          /* 0 literal, 1 copy1, 2 copy2, 3 copy4 */
    uint32_t length;   /* bytes produced */
    uint32_t offset;   /* copies only */
    size_t payload;    /* literals: index of first literal byte */
    size_t next;       /* index of the next tag */
} element;

/* returns 0 on success, nonzero when the element header runs past the input */
static int read_element(const uint8_t *in, size_t n, size_t pos, element *e)
{
    uint8_t tag = in[pos];
    e->kind = tag & 3;
    switch (e->kind) {
    case 0: {
        uint32_t m = tag >> 2;
        size_t p = pos + 1;
        if (m >= 60) {
            unsigned extra = m - 59;
            if (p + extra > n) return 1;
            m = 0;
            for (unsigned k = 0; k < extra; k++) m |= (uint32_t)in[p + k] << (8 * k);
            p += extra;
        }
        /* m + 1 may wrap for a hostile 4-byte length of 0xFFFFFFFF: treat as too long */
        if (m == 0xFFFFFFFFu) return 1;
        e->length = m + 1;
        e->payload = p;
        if ((uint64_t)p + e->length > n) return 1;
        e->next = p + e->length;
        return 0;
    }
    case 1:
        if (pos + 2 > n) return 1;
        e->length = 4 + ((tag >> 2) & 7);
        e->offset = ((uint32_t)(tag >> 5) << 8) | in[pos + 1];
        e->next = pos + 2;
        return 0;
    case 2:
        if (pos + 3 > n) return 1;
        e->length = 1 + (tag >> 2);
        e->offset = in[pos + 1] | ((uint32_t)in[pos + 2] << 8);
        e->next = pos + 3;
        return 0;
    default:
        if (pos + 5 > n) return 1;
        e->length = 1 + (tag >> 2);
        e->offset = in[pos + 1] | ((uint32_t)in[pos + 2] << 8) | ((uint32_t)in[pos + 3] << 16) |
                    ((uint32_t)in[pos + 4] << 24);
        e->next = pos + 5;
        return 0;
    }
}

int orc_snappy_uncompress(const uint8_t *in, size_t n, uint8_t *out, size_t *out_len)
{
    size_t want;
    if (orc_snappy_uncompressed_length(in, n, &want) != ORC_SNAPPY_OK) return ORC_SNAPPY_INVALID_INPUT;
    if (want > *out_len) return ORC_SNAPPY_BUFFER_TOO_SMALL;
    size_t pos = varint_width(in, n);
    size_t produced = 0;
    while (pos < n) {
        element e;
        if (read_element(in, n, pos, &e)) return ORC_SNAPPY_INVALID_INPUT;
        if (e.length > want - produced) return ORC_SNAPPY_INVALID_INPUT;
        if (e.kind == 0) {
            memcpy(out + produced, in + e.payload, e.length);
        } else {
            if (e.offset == 0 || e.offset > produced) return ORC_SNAPPY_INVALID_INPUT;
            /* byte-by-byte: overlapping copies repeat the last `offset` bytes */
            const uint8_t *src = out + produced - e.offset;
            uint8_t *dst = out + produced;
            for (uint32_t k = 0; k < e.length; k++) dst[k] = src[k];
        }
        produced += e.length;
        pos = e.next;
    }
    if (produced != want) return ORC_SNAPPY_INVALID_INPUT;
    *out_len = produced;
    return ORC_SNAPPY_OK;
}

int orc_snappy_scan(const uint8_t *in, size_t n, orc_snappy_stats *st)
{
    size_t want;
    memset(st, 0, sizeof *st);
    if (orc_snappy_uncompressed_length(in, n, &want) != ORC_SNAPPY_OK) return ORC_SNAPPY_INVALID_INPUT;
    size_t pos = varint_width(in, n);
    size_t produced = 0;
    while (pos < n) {
        element e;
        if (read_element(in, n, pos, &e)) return ORC_SNAPPY_INVALID_INPUT;
        if (e.kind == 0) {
            st->literals++;
            st->literal_bytes += e.length;
        } else {
            if (e.offset == 0 || e.offset > produced) return ORC_SNAPPY_INVALID_INPUT;
            if (e.kind == 1) st->copy1++; else if (e.kind == 2) st->copy2++; else st->copy4++;
            st->copy_bytes += e.length;
            if (e.offset < e.length) st->overlapping++;
            if (e.offset > st->max_offset) st->max_offset = e.offset;
        }
        produced += e.length;
        pos = e.next;
    }
    return produced == want ? ORC_SNAPPY_OK : ORC_SNAPPY_INVALID_INPUT;
}

/* ---- compressor ---------------------------------------------------------------------------- */

#define FRAGMENT 65536u
#define HASH_BITS 14

static uint32_t load32(const uint8_t *p)
{
    uint32_t v;
    memcpy(&v, p, 4);
    return v;
}

static uint8_t *put_literal(uint8_t *op, const uint8_t *src, uint32_t len)
{
    uint32_t m = len - 1;
    if (m < 60) {
        *op++ = (uint8_t)(m << 2);
    } else {
        unsigned extra = m < 0x100 ? 1 : m < 0x10000 ? 2 : m < 0x1000000 ? 3 : 4;
        *op++ = (uint8_t)((59 + extra) << 2);
        for (unsigned k = 0; k < extra; k++) *op++ = (uint8_t)(m >> (8 * k));
    }
    memcpy(op, src, len);
    return op + len;
}

static uint8_t *put_copy_piece(uint8_t *op, uint32_t offset, uint32_t len)
{
    if (len >= 4 && len <= 11 && offset < 2048) {
        *op++ = (uint8_t)(1 | ((len - 4) << 2) | ((offset >> 8) << 5));
        *op++ = (uint8_t)offset;
    } else {
        *op++ = (uint8_t)(2 | ((len - 1) << 2));
        *op++ = (uint8_t)offset;
        *op++ = (uint8_t)(offset >> 8);
    }
    return op;
}

static uint8_t *put_copy(uint8_t *op, uint32_t offset, uint32_t len)
{
    /* pieces of at most 64; keep the tail >= 4 so it can still be a copy element */
    while (len >= 68) {
        op = put_copy_piece(op, offset, 64);
        len -= 64;
    }
    if (len > 64) {
        op = put_copy_piece(op, offset, 60);
        len -= 60;
    }
    return put_copy_piece(op, offset, len);
}

static uint8_t *compress_fragment(const uint8_t *base, uint32_t n, uint8_t *op, uint16_t *table)
{
    memset(table, 0, sizeof(uint16_t) << HASH_BITS);
    const uint32_t shift = 32 - HASH_BITS;
    uint32_t anchor = 0; /* first byte not yet emitted */
    if (n >= 15) {
        const uint32_t limit = n - 4; /* last position where load32 is legal */
        uint32_t ip = 1;
        uint32_t miss = 32;
        while (ip <= limit) {
            uint32_t h = (load32(base + ip) * 0x1e35a7bdu) >> shift;
            uint32_t cand = table[h];
            table[h] = (uint16_t)ip;
            if (cand < ip && load32(base + cand) == load32(base + ip)) {
                if (ip > anchor) op = put_literal(op, base + anchor, ip - anchor);
                uint32_t len = 4;
                while (ip + len < n && base[cand + len] == base[ip + len]) len++;
                op = put_copy(op, ip - cand, len);
                ip += len;
                anchor = ip;
                miss = 32;
                if (ip <= limit && ip >= 1) {
                    /* seed the table with the position just before the new anchor */
                    uint32_t hp = (load32(base + ip - 1) * 0x1e35a7bdu) >> shift;
                    table[hp] = (uint16_t)(ip - 1);
                }
            } else {
                ip += miss++ >> 5; /* accelerate through incompressible data */
            }
        }
    }
    if (anchor < n) op = put_literal(op, base + anchor, n - anchor);
    return op;
}

int orc_snappy_compress(const uint8_t *in, size_t n, uint8_t *out, size_t *out_len)
{
    if (*out_len < orc_snappy_max_compressed_length(n)) return ORC_SNAPPY_BUFFER_TOO_SMALL;
    if (n > 0xFFFFFFFFull) return ORC_SNAPPY_INVALID_INPUT;
    uint8_t *op = out;
    uint32_t v = (uint32_t)n;
    while (v >= 0x80) {
        *op++ = (uint8_t)(v | 0x80);
        v >>= 7;
    }
    *op++ = (uint8_t)v;
    static __thread uint16_t table[1 << HASH_BITS];
    for (size_t done = 0; done < n; done += FRAGMENT) {
        uint32_t take = (uint32_t)((n - done) < FRAGMENT ? (n - done) : FRAGMENT);
        op = compress_fragment(in + done, take, op, table);
    }
    *out_len = (size_t)(op - out);
    return ORC_SNAPPY_OK;
}
