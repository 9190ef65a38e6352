/*
 * oracle/hap_oracle.c -- TEST INFRASTRUCTURE ONLY. See hap_oracle.h.
 *
 * Restates /root/reference/source/hap.c behaviour by behaviour (not line by line): every function
 * names the reference lines it follows.  Quirk numbers (Q1..Q11) refer to SURVEY.md section 8(a).
 */
#include "hap_oracle.h"
#include "snappy_oracle.h"
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { R_OK = 0, R_BAD_ARGS = 1, R_TOO_SMALL = 2, R_BAD_FRAME = 3, R_INTERNAL = 4 }; /* hap.h:55-61 */
enum { CMP_NONE = 0, CMP_SNAPPY = 1 };                                               /* hap.h:50-53 */
enum { ST_NONE = 0xA, ST_SNAPPY = 0xB, ST_COMPLEX = 0xC };                           /* hap.c:41-43 */
enum { SEC_MULTI = 0x0D, SEC_DI = 0x01, SEC_COMPRESSORS = 0x02, SEC_SIZES = 0x03, SEC_OFFSETS = 0x04 }; /* hap.c:84-88 */
#define U24_MAX 0xFFFFFFu

/* hap.c:215-261: wire nibble <-> API enum */
static const struct { unsigned nibble, api; } k_formats[] = {
    {0xB, 0x83F0}, {0xE, 0x83F3}, {0xF, 0x01}, {0x1, 0x8DBB}, {0xC, 0x8E8C}, {0x2, 0x8E8F}, {0x3, 0x8E8E},
};
static unsigned api_from_nibble(unsigned nib)
{
    for (unsigned i = 0; i < 7; i++) if (k_formats[i].nibble == nib) return k_formats[i].api;
    return 0;
}
static unsigned nibble_from_api(unsigned api)
{
    for (unsigned i = 0; i < 7; i++) if (k_formats[i].api == api) return k_formats[i].nibble;
    return 0;
}

static uint32_t le24(const uint8_t *p) { return p[0] | (p[1] << 8) | ((uint32_t)p[2] << 16); }
static uint32_t le32(const uint8_t *p) { return le24(p) | ((uint32_t)p[3] << 24); }
static void put_le32(uint8_t *p, uint32_t v) { p[0] = v; p[1] = v >> 8; p[2] = v >> 16; p[3] = v >> 24; }

typedef struct { uint32_t hdr, len; unsigned type; } section;

/* hap.c:137-187. avail is 32-bit in the reference and so is the hdr+len sum (hap.c:181). */
static int section_read(const uint8_t *p, uint32_t avail, section *s)
{
    if (avail < 4) return R_BAD_FRAME;
    s->len = le24(p);
    s->hdr = 4;
    if (s->len == 0) {
        if (avail < 8) return R_BAD_FRAME;
        s->len = le32(p + 4);
        s->hdr = 8;
    }
    s->type = p[3];
    if ((uint32_t)(s->hdr + s->len) > avail) return R_BAD_FRAME;
    return R_OK;
}

/* hap.c:189-212 */
static void section_write(uint8_t *p, size_t hdr, uint32_t len, unsigned type)
{
    if (hdr == 4) {
        p[0] = len; p[1] = len >> 8; p[2] = len >> 16;
    } else {
        p[0] = p[1] = p[2] = 0;
        put_le32(p + 4, len);
    }
    p[3] = (uint8_t)type;
}

static size_t di_length(unsigned k) { return 5u * (size_t)k + 8; } /* hap.c:265-275 */

/* hap.c:277-300 */
unsigned int orc_hap_limited_chunk_count(unsigned long bytes, unsigned int fmt, unsigned int k)
{
    if (k > 3355431u) k = 3355431u;
    unsigned long blocks = (fmt == 0x83F0 || fmt == 0x8DBB) ? bytes / 8 : bytes / 16;
    while (blocks % k) k--;
    return k;
}

/* hap.c:302-322 */
static size_t texture_bound(size_t bytes, unsigned fmt, unsigned compressor, unsigned k)
{
    k = orc_hap_limited_chunk_count(bytes, fmt, k);
    size_t payload = bytes;
    if (compressor == CMP_SNAPPY) payload = orc_snappy_max_compressed_length(bytes / k) * k;
    return payload + 8 + di_length(k) + 4;
}

/* hap.c:324-353 */
unsigned long orc_HapMaxEncodedLength(unsigned int count, unsigned long *lengths,
                                      unsigned int *fmts, unsigned int *chunks)
{
    if (count == 0 || count > 2 || !lengths || !fmts || !chunks) return 0;
    unsigned long total = 8;
    for (unsigned i = 0; i < count; i++) {
        if (chunks[i] == 0) return 0;
        total += texture_bound(lengths[i], fmts[i], CMP_SNAPPY, chunks[i]);
    }
    return total;
}

/* hap.c:355-504 */
static unsigned encode_texture(const uint8_t *in, unsigned long n, unsigned fmt, unsigned compressor,
                               unsigned k, uint8_t *out, unsigned long cap, unsigned long *used)
{
    if (!in || n == 0 || nibble_from_api(fmt) == 0 || (compressor != CMP_NONE && compressor != CMP_SNAPPY) ||
        !out || !used)
        return R_BAD_ARGS;
    if (cap < texture_bound(n, fmt, compressor, k)) return R_TOO_SMALL;

    size_t hdr = n > U24_MAX ? 8 : 4; /* hap.c:398-405 (Q2: chosen before compression) */
    size_t body = 0;
    unsigned stored = ST_NONE;

    if (compressor == CMP_SNAPPY) {
        k = orc_hap_limited_chunk_count(n, fmt, k);
        size_t di = di_length(k);
        if (n + di + 4 > U24_MAX) hdr = 8; /* hap.c:425-428 */
        uint8_t *p = out + hdr;
        section_write(p, 4, (uint32_t)di, SEC_DI);
        section_write(p + 4, 4, k, SEC_COMPRESSORS);
        uint8_t *ctab = p + 8;
        section_write(ctab + k, 4, 4 * k, SEC_SIZES);
        uint8_t *stab = ctab + k + 4;
        uint8_t *data = p + 4 + di;
        size_t room = cap - hdr - 4 - di;
        size_t chunk = n / k; /* Q3: truncating */
        body = 4 + di;
        for (unsigned i = 0; i < k; i++) {
            size_t packed = room;
            if (orc_snappy_compress(in + chunk * i, chunk, data, &packed) != ORC_SNAPPY_OK) return R_INTERNAL;
            if (packed >= chunk) { /* hap.c:460-466 */
                memcpy(data, in + chunk * i, chunk);
                packed = chunk;
                ctab[i] = ST_NONE;
            } else {
                ctab[i] = ST_SNAPPY;
            }
            put_le32(stab + 4 * i, (uint32_t)packed);
            data += packed;
            body += packed;
            room -= packed;
        }
        if (body < n + hdr) stored = ST_COMPLEX; /* hap.c:478-482; Q1: never 0xB */
        else compressor = CMP_NONE;              /* whole-frame fallback keeps hdr (Q2) */
    }
    if (compressor == CMP_NONE) {
        memcpy(out + hdr, in, n);
        body = n;
        stored = ST_NONE;
    }
    section_write(out, hdr, (uint32_t)body, (stored << 4) | (nibble_from_api(fmt) & 0xF));
    *used = body + hdr;
    return R_OK;
}

/* hap.c:506-604 */
unsigned int orc_HapEncode(unsigned int count, const void **ins, unsigned long *lens,
                           unsigned int *fmts, unsigned int *compressors, unsigned int *chunks,
                           void *outv, unsigned long cap, unsigned long *used)
{
    uint8_t *out = (uint8_t *)outv;
    if (count == 0 || count > 2 || !ins || !lens || !fmts || !compressors || !chunks || !out || cap == 0 || !used)
        return R_BAD_ARGS;
    for (unsigned i = 0; i < count; i++) if (chunks[i] == 0) return R_BAD_ARGS;
    if (count == 1)
        return encode_texture((const uint8_t *)ins[0], lens[0], fmts[0], compressors[0], chunks[0], out, cap, used);
    /* Q5: rejects only when neither texture is YCoCg and neither is RGTC1 */
    if (fmts[0] != 0x01 && fmts[1] != 0x01 && fmts[0] != 0x8DBB && fmts[1] != 0x8DBB) return R_BAD_ARGS;
    size_t worst = 0;
    for (unsigned i = 0; i < count; i++) worst += lens[i] + di_length(chunks[i]) + 4; /* Q6 */
    size_t hdr = worst > U24_MAX ? 8 : 4;
    size_t body = 0;
    for (unsigned i = 0; i < count; i++) {
        unsigned long sec = 0;
        unsigned r = encode_texture((const uint8_t *)ins[i], lens[i], fmts[i], compressors[i], chunks[i],
                                    out + hdr + body, cap - (hdr + body), &sec);
        if (r != R_OK) return r;
        body += sec;
    }
    section_write(out, hdr, (uint32_t)body, SEC_MULTI);
    *used = body + hdr;
    return R_OK;
}

/* ---- decode -------------------------------------------------------------------------------- */

typedef struct {
    unsigned result, compressor;
    const uint8_t *src;
    size_t src_bytes;
    uint8_t *dst;
    size_t dst_bytes;
} chunk_job; /* hap.c:93-100 */

/* hap.c:606-642 */
static void run_chunk(void *p, unsigned int i)
{
    chunk_job *jobs = (chunk_job *)p;
    if (!jobs) return;
    chunk_job *j = &jobs[i];
    if (j->compressor == ST_SNAPPY) {
        int r = orc_snappy_uncompress(j->src, j->src_bytes, j->dst, &j->dst_bytes);
        j->result = r == ORC_SNAPPY_OK ? R_OK : r == ORC_SNAPPY_INVALID_INPUT ? R_BAD_FRAME : R_INTERNAL;
    } else if (j->compressor == ST_NONE) {
        memcpy(j->dst, j->src, j->src_bytes);
        j->result = R_OK;
    } else {
        j->result = R_BAD_FRAME;
    }
}

typedef struct { const uint8_t *compressors, *sizes, *offsets, *data; int count; } di_tables;

/* hap.c:644-730; *count carries in/out exactly as the reference's int *chunk_count does */
static int parse_di(const uint8_t *sec, uint32_t sec_len, di_tables *t)
{
    section s;
    t->compressors = t->sizes = t->offsets = NULL;
    int r = section_read(sec, sec_len, &s);
    if (r == R_OK && s.type != SEC_DI) r = R_BAD_FRAME;
    if (r != R_OK) return r;
    t->data = sec + s.hdr + s.len;
    const uint8_t *p = sec + s.hdr;
    size_t left = s.len;
    while (left > 0) {
        section in;
        r = section_read(p, (uint32_t)left, &in);
        if (r != R_OK) return r;
        p += in.hdr;
        unsigned c = 0;
        if (in.type == SEC_COMPRESSORS) { t->compressors = p; c = in.len; }
        else if (in.type == SEC_SIZES) { t->sizes = p; c = in.len / 4; }
        else if (in.type == SEC_OFFSETS) { t->offsets = p; c = in.len / 4; }
        if (c != 0) {
            if (t->count != 0 && (int)c != t->count) return R_BAD_FRAME;
            t->count = (int)c;
        }
        p += in.len;
        left -= in.hdr + in.len;
    }
    if (!t->compressors || !t->sizes) return R_BAD_FRAME;
    return R_OK;
}

/* hap.c:732-930 */
static unsigned decode_texture(const uint8_t *sec, uint32_t sec_len, unsigned type, orc_decode_cb cb,
                               void *info, uint8_t *out, unsigned long cap, unsigned long *used, unsigned *fmt)
{
    unsigned compressor = (type >> 4) & 0xF;
    *fmt = api_from_nibble(type & 0xF);
    if (*fmt == 0) return R_BAD_FRAME;
    size_t produced = 0;
    if (compressor == ST_COMPLEX) {
        di_tables t;
        t.count = 0;
        int r = parse_di(sec, sec_len, &t);
        if (r != R_OK) return r;
        if (t.count > 0) {
            chunk_job *jobs = (chunk_job *)malloc(sizeof(chunk_job) * (size_t)t.count);
            if (!jobs) return R_INTERNAL;
            size_t in_run = 0, out_run = 0;
            for (int i = 0; i < t.count; i++) {
                jobs[i].compressor = t.compressors[i];
                jobs[i].src_bytes = le32(t.sizes + 4 * i);
                jobs[i].src = t.data + (t.offsets ? le32(t.offsets + 4 * i) : in_run);
                in_run += jobs[i].src_bytes;
                if (jobs[i].compressor == ST_SNAPPY) {
                    int sr = orc_snappy_uncompressed_length(jobs[i].src, jobs[i].src_bytes, &jobs[i].dst_bytes);
                    if (sr != ORC_SNAPPY_OK) { /* Q7 */
                        r = sr == ORC_SNAPPY_INVALID_INPUT ? R_BAD_FRAME : R_INTERNAL;
                        break;
                    }
                } else {
                    jobs[i].dst_bytes = jobs[i].src_bytes;
                }
                jobs[i].dst = out + out_run;
                out_run += jobs[i].dst_bytes;
            }
            if (r == R_OK && out_run > cap) r = R_TOO_SMALL;
            if (r == R_OK) {
                produced = out_run;
                if (t.count == 1) run_chunk(jobs, 0);          /* hap.c:852-858 */
                else cb(run_chunk, jobs, (unsigned)t.count, info); /* hap.c:861 */
                for (int i = 0; i < t.count; i++)
                    if (jobs[i].result != R_OK) { r = (int)jobs[i].result; break; }
            }
            free(jobs);
            if (r != R_OK) return (unsigned)r;
        }
    } else if (compressor == ST_SNAPPY) { /* hap.c:885-904; Q7: everything maps to Internal_Error */
        if (orc_snappy_uncompressed_length(sec, sec_len, &produced) != ORC_SNAPPY_OK) return R_INTERNAL;
        if (produced > cap) return R_TOO_SMALL;
        if (orc_snappy_uncompress(sec, sec_len, out, &produced) != ORC_SNAPPY_OK) return R_INTERNAL;
    } else if (compressor == ST_NONE) { /* hap.c:905-916 */
        produced = sec_len;
        if (sec_len > cap) return R_TOO_SMALL;
        memcpy(out, sec, sec_len);
    } else {
        return R_BAD_FRAME;
    }
    if (used) *used = produced;
    return R_OK;
}

/* hap.c:932-991 */
static int locate(const uint8_t *in, uint32_t n, unsigned index, const uint8_t **sec, uint32_t *len, unsigned *type)
{
    section top;
    int r = section_read(in, n, &top);
    *len = top.len; /* the reference writes through its out-pointers even when it fails */
    *type = top.type;
    if (r != R_OK) return r;
    if (top.type == SEC_MULTI) {
        const uint8_t *body = in + top.hdr;
        size_t off = 0;
        section cur = {0, 0, 0};
        for (unsigned i = 0; i <= index; i++) {
            off += cur.hdr + cur.len;
            if (off >= top.len) return R_BAD_ARGS; /* Q8 */
            r = section_read(body + off, (uint32_t)(top.len - off), &cur);
            if (r != R_OK) return r;
        }
        *sec = body + off + cur.hdr;
        *len = cur.len;
        *type = cur.type;
        return R_OK;
    }
    if (index == 0) {
        *sec = in + top.hdr;
        return R_OK;
    }
    *sec = NULL; *len = 0; *type = 0;
    return R_BAD_ARGS;
}

/* hap.c:993-1040 */
unsigned int orc_HapDecode(const void *in, unsigned long n, unsigned int index, orc_decode_cb cb, void *info,
                           void *out, unsigned long cap, unsigned long *used, unsigned int *fmt)
{
    if (!in || index > 1 || !cb || !out || !fmt) return R_BAD_ARGS;
    const uint8_t *sec; uint32_t len; unsigned type;
    int r = locate((const uint8_t *)in, (uint32_t)n, index, &sec, &len, &type);
    if (r != R_OK) return (unsigned)r;
    return decode_texture(sec, len, type, cb, info, (uint8_t *)out, cap, used, fmt);
}

/* hap.c:1042-1087 (Q10: no NULL checks) */
unsigned int orc_HapGetFrameTextureCount(const void *inv, unsigned long n, unsigned int *count)
{
    const uint8_t *in = (const uint8_t *)inv;
    section top;
    int r = section_read(in, (uint32_t)n, &top);
    if (r != R_OK) return (unsigned)r;
    if (top.type != SEC_MULTI) { *count = 1; return R_OK; }
    uint32_t off = top.hdr;
    *count = 0;
    while (off < top.len) { /* sic: compares an offset that includes the header with the body length */
        section s;
        r = section_read(in + off, (uint32_t)(n - off), &s);
        if (r != R_OK) return (unsigned)r;
        off += s.hdr + s.len;
        *count += 1;
    }
    return R_OK;
}

/* hap.c:1089-1126 */
unsigned int orc_HapGetFrameTextureFormat(const void *in, unsigned long n, unsigned int index, unsigned int *fmt)
{
    if (!in || index > 1 || !fmt) return R_BAD_ARGS;
    const uint8_t *sec; uint32_t len; unsigned type;
    int r = locate((const uint8_t *)in, (uint32_t)n, index, &sec, &len, &type);
    if (r != R_OK) return (unsigned)r;
    *fmt = api_from_nibble(type & 0xF);
    return *fmt ? R_OK : R_BAD_FRAME;
}

/* hap.c:1128-1188 */
unsigned int orc_HapGetFrameTextureChunkCount(const void *in, unsigned long n, unsigned int index, int *chunk_count)
{
    *chunk_count = 0; /* Q10 */
    if (!in || index > 1) return R_BAD_ARGS;
    const uint8_t *sec; uint32_t len; unsigned type;
    int r = locate((const uint8_t *)in, (uint32_t)n, index, &sec, &len, &type);
    if (r != R_OK) return (unsigned)r;
    unsigned compressor = (type >> 4) & 0xF;
    if (compressor == ST_COMPLEX) {
        di_tables t;
        t.count = *chunk_count;
        r = parse_di(sec, len, &t);
        *chunk_count = t.count;
        return (unsigned)r;
    }
    if (compressor == ST_SNAPPY || compressor == ST_NONE) { *chunk_count = 1; return R_OK; }
    return R_BAD_FRAME;
}
