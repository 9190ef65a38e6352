// hap_b200/csrc/hap_mov.h -- QuickTime sample-table reader / writer behind include/hap_mov.h.
//
// Host code only.  The reference has no container code (documentation/HapVideoDRAFT.md:14); what is
// implemented here is the QuickTime File Format's movie atom tree as far as one video track of Hap samples
// needs it: ftyp, mdat, moov { mvhd, trak { tkhd, mdia { mdhd, hdlr, minf { vmhd, hdlr, dinf { dref },
// stbl { stsd, stts, stsc, stsz, stco | co64 } } } } }.  All integers are big-endian.  The reader accepts what
// other writers produce (any atom order, unknown atoms skipped, 64-bit atom sizes, version-1 headers, several
// samples per chunk, stco or co64) and treats the file as hostile: every size and offset is checked against
// the enclosing atom and the file length before it is used.
#pragma once
#include "../../include/hap_mov.h"

#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

struct HapB200Mov {
    FILE *f = nullptr;
    bool writing = false;
    uint32_t fourcc = 0, width = 0, height = 0, timescale = 0;
    uint64_t duration = 0;                 // in timescale ticks
    std::vector<uint64_t> offset;          // per frame
    std::vector<uint32_t> size;
    std::vector<uint32_t> ticks;
    // writer state
    uint64_t mdat_header_at = 0;           // file position of the mdat atom
    uint64_t write_pos = 0;
    bool io_error = false;
};

namespace hapmov {

inline bool is_hap_fourcc(uint32_t c)
{
    return c == HAPB200_FOURCC('H', 'a', 'p', '1') || c == HAPB200_FOURCC('H', 'a', 'p', '5') || c == HAPB200_FOURCC('H', 'a', 'p', 'Y') ||
           c == HAPB200_FOURCC('H', 'a', 'p', 'M') || c == HAPB200_FOURCC('H', 'a', 'p', 'A') || c == HAPB200_FOURCC('H', 'a', 'p', '7') ||
           c == HAPB200_FOURCC('H', 'a', 'p', 'H');
}

inline uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
inline uint64_t be64(const uint8_t *p) { return ((uint64_t)be32(p) << 32) | be32(p + 4); }
inline uint16_t be16(const uint8_t *p) { return (uint16_t)((p[0] << 8) | p[1]); }

// ---- reading -----------------------------------------------------------------------------------------------

struct Atom {
    uint32_t type = 0;
    uint64_t body = 0, body_len = 0;   // file position and length of the payload
};

// Reads the atom header at `pos` inside [pos, end).  False when there is no well-formed atom there.
inline bool read_atom(FILE *f, uint64_t pos, uint64_t end, Atom &a)
{
    if (end < pos || end - pos < 8) return false;
    uint8_t h[16];
    if (fseeko(f, (off_t)pos, SEEK_SET) != 0 || fread(h, 1, 8, f) != 8) return false;
    uint64_t sz = be32(h);
    uint64_t hdr = 8;
    a.type = be32(h + 4);
    if (sz == 1) {
        if (end - pos < 16 || fread(h + 8, 1, 8, f) != 8) return false;
        sz = be64(h + 8);
        hdr = 16;
    } else if (sz == 0) {
        sz = end - pos;   // "extends to the end of the enclosing space"
    }
    if (sz < hdr || sz > end - pos) return false;
    a.body = pos + hdr;
    a.body_len = sz - hdr;
    return true;
}

// First child of type `type` inside the payload [body, body + len).
inline bool find_child(FILE *f, uint64_t body, uint64_t len, uint32_t type, Atom &out)
{
    uint64_t pos = body, end = body + len;
    while (end - pos >= 8) {
        Atom a;
        if (!read_atom(f, pos, end, a)) return false;
        if (a.type == type) { out = a; return true; }
        pos = a.body + a.body_len;
    }
    return false;
}

inline bool read_body(FILE *f, const Atom &a, uint64_t max_len, std::vector<uint8_t> &buf)
{
    if (a.body_len > max_len) return false;
    buf.resize((size_t)a.body_len);
    if (a.body_len == 0) return true;
    return fseeko(f, (off_t)a.body, SEEK_SET) == 0 && fread(buf.data(), 1, buf.size(), f) == buf.size();
}

constexpr uint64_t kMaxTableBytes = 1ull << 30;   // a sample table larger than this is not a movie we open
constexpr uint64_t kMaxFrames = 1ull << 27;

// Fills m from the first video track whose sample description is a Hap FourCC.
inline bool parse_track(FILE *f, const Atom &trak, uint64_t file_len, HapB200Mov &m)
{
    Atom mdia, hdlr, mdhd, minf, stbl, stsd, stts, stsc, stsz, stco;
    if (!find_child(f, trak.body, trak.body_len, HAPB200_FOURCC('m', 'd', 'i', 'a'), mdia)) return false;
    if (!find_child(f, mdia.body, mdia.body_len, HAPB200_FOURCC('h', 'd', 'l', 'r'), hdlr)) return false;
    std::vector<uint8_t> b;
    if (!read_body(f, hdlr, 4096, b) || b.size() < 12 || be32(&b[8]) != HAPB200_FOURCC('v', 'i', 'd', 'e')) return false;
    if (!find_child(f, mdia.body, mdia.body_len, HAPB200_FOURCC('m', 'd', 'h', 'd'), mdhd) || !read_body(f, mdhd, 4096, b) || b.size() < 24) return false;
    if (b[0] == 1) {
        if (b.size() < 36) return false;
        m.timescale = be32(&b[20]);
        m.duration = be64(&b[24]);
    } else {
        m.timescale = be32(&b[12]);
        m.duration = be32(&b[16]);
    }
    if (m.timescale == 0) return false;
    if (!find_child(f, mdia.body, mdia.body_len, HAPB200_FOURCC('m', 'i', 'n', 'f'), minf)) return false;
    if (!find_child(f, minf.body, minf.body_len, HAPB200_FOURCC('s', 't', 'b', 'l'), stbl)) return false;
    // sample description: first entry
    if (!find_child(f, stbl.body, stbl.body_len, HAPB200_FOURCC('s', 't', 's', 'd'), stsd) || !read_body(f, stsd, 1 << 20, b)) return false;
    if (b.size() < 8 + 86 || be32(&b[4]) < 1) return false;
    const uint8_t *d = &b[8];
    if (be32(d) < 86 || be32(d) > b.size() - 8) return false;
    m.fourcc = be32(d + 4);
    if (!is_hap_fourcc(m.fourcc)) return false;
    m.width = be16(d + 32);
    m.height = be16(d + 34);
    if (m.width == 0 || m.height == 0) return false;
    // sample sizes
    if (!find_child(f, stbl.body, stbl.body_len, HAPB200_FOURCC('s', 't', 's', 'z'), stsz) || !read_body(f, stsz, kMaxTableBytes, b) || b.size() < 12) return false;
    const uint32_t uniform = be32(&b[4]);
    const uint64_t n = be32(&b[8]);
    if (n > kMaxFrames) return false;
    if (uniform == 0 && b.size() < 12 + 4 * n) return false;
    m.size.resize((size_t)n);
    for (uint64_t i = 0; i < n; i++) m.size[i] = uniform ? uniform : be32(&b[12 + 4 * i]);
    // durations
    m.ticks.assign((size_t)n, 0);
    if (find_child(f, stbl.body, stbl.body_len, HAPB200_FOURCC('s', 't', 't', 's'), stts) && read_body(f, stts, kMaxTableBytes, b) && b.size() >= 8) {
        const uint64_t runs = be32(&b[4]);
        if (b.size() < 8 + 8 * runs) return false;
        uint64_t i = 0;
        for (uint64_t r = 0; r < runs && i < n; r++) {
            const uint64_t cnt = be32(&b[8 + 8 * r]);
            const uint32_t dt = be32(&b[12 + 8 * r]);
            for (uint64_t k = 0; k < cnt && i < n; k++) m.ticks[i++] = dt;
        }
    }
    // chunk offsets
    std::vector<uint64_t> chunk_off;
    if (find_child(f, stbl.body, stbl.body_len, HAPB200_FOURCC('s', 't', 'c', 'o'), stco)) {
        if (!read_body(f, stco, kMaxTableBytes, b) || b.size() < 8) return false;
        const uint64_t c = be32(&b[4]);
        if (b.size() < 8 + 4 * c) return false;
        chunk_off.resize((size_t)c);
        for (uint64_t i = 0; i < c; i++) chunk_off[i] = be32(&b[8 + 4 * i]);
    } else if (find_child(f, stbl.body, stbl.body_len, HAPB200_FOURCC('c', 'o', '6', '4'), stco)) {
        if (!read_body(f, stco, kMaxTableBytes, b) || b.size() < 8) return false;
        const uint64_t c = be32(&b[4]);
        if (b.size() < 8 + 8 * c) return false;
        chunk_off.resize((size_t)c);
        for (uint64_t i = 0; i < c; i++) chunk_off[i] = be64(&b[8 + 8 * i]);
    } else {
        return false;
    }
    // sample-to-chunk runs -> offset of every sample
    if (!find_child(f, stbl.body, stbl.body_len, HAPB200_FOURCC('s', 't', 's', 'c'), stsc) || !read_body(f, stsc, kMaxTableBytes, b) || b.size() < 8) return false;
    const uint64_t runs = be32(&b[4]);
    if (b.size() < 8 + 12 * runs) return false;
    m.offset.assign((size_t)n, 0);
    uint64_t sample = 0;
    for (uint64_t r = 0; r < runs && sample < n; r++) {
        const uint64_t first = be32(&b[8 + 12 * r]), per = be32(&b[12 + 12 * r]);
        const uint64_t next_first = r + 1 < runs ? be32(&b[8 + 12 * (r + 1)]) : (uint64_t)chunk_off.size() + 1;
        if (first < 1 || next_first < first || per == 0) return false;
        for (uint64_t c = first; c < next_first && sample < n; c++) {
            if (c > chunk_off.size()) return false;
            uint64_t at = chunk_off[(size_t)(c - 1)];
            for (uint64_t k = 0; k < per && sample < n; k++) {
                m.offset[(size_t)sample] = at;
                at += m.size[(size_t)sample];
                sample++;
            }
        }
    }
    if (sample != n) return false;
    for (uint64_t i = 0; i < n; i++)
        if (m.offset[(size_t)i] > file_len || m.size[(size_t)i] > file_len - m.offset[(size_t)i]) return false;
    return true;
}

inline HapB200Mov *open_read(const char *path)
{
    if (!path) return nullptr;
    FILE *f = fopen(path, "rb");
    if (!f) return nullptr;
    HapB200Mov *m = new HapB200Mov;
    m->f = f;
    bool ok = false;
    if (fseeko(f, 0, SEEK_END) == 0) {
        const uint64_t file_len = (uint64_t)ftello(f);
        uint64_t pos = 0;
        while (file_len - pos >= 8) {
            Atom a;
            if (!read_atom(f, pos, file_len, a)) break;
            if (a.type == HAPB200_FOURCC('m', 'o', 'o', 'v')) {
                uint64_t p2 = a.body;
                const uint64_t e2 = a.body + a.body_len;
                while (!ok && e2 - p2 >= 8) {
                    Atom t;
                    if (!read_atom(f, p2, e2, t)) break;
                    if (t.type == HAPB200_FOURCC('t', 'r', 'a', 'k')) {
                        HapB200Mov trial;
                        if (parse_track(f, t, file_len, trial)) {
                            m->fourcc = trial.fourcc; m->width = trial.width; m->height = trial.height;
                            m->timescale = trial.timescale; m->duration = trial.duration;
                            m->offset.swap(trial.offset); m->size.swap(trial.size); m->ticks.swap(trial.ticks);
                            ok = true;
                        }
                    }
                    p2 = t.body + t.body_len;
                }
                break;
            }
            pos = a.body + a.body_len;
        }
    }
    if (!ok) {
        fclose(f);
        delete m;
        return nullptr;
    }
    return m;
}

// ---- writing -----------------------------------------------------------------------------------------------

struct Buf {
    std::vector<uint8_t> b;
    void u8(uint32_t v) { b.push_back((uint8_t)v); }
    void u16(uint32_t v) { u8(v >> 8); u8(v); }
    void u32(uint32_t v) { u16(v >> 16); u16(v); }
    void u64(uint64_t v) { u32((uint32_t)(v >> 32)); u32((uint32_t)v); }
    void zeros(int n) { b.insert(b.end(), (size_t)n, 0); }
    void bytes(const void *p, size_t n) { b.insert(b.end(), (const uint8_t *)p, (const uint8_t *)p + n); }
    void pstr(const char *s, int field)   // Pascal string padded to `field` bytes
    {
        const size_t n = strlen(s);
        u8((uint32_t)n);
        bytes(s, n);
        zeros(field - 1 - (int)n);
    }
    size_t begin(uint32_t type) { const size_t at = b.size(); u32(0); u32(type); return at; }
    void end(size_t at)
    {
        const uint32_t n = (uint32_t)(b.size() - at);
        b[at] = (uint8_t)(n >> 24); b[at + 1] = (uint8_t)(n >> 16); b[at + 2] = (uint8_t)(n >> 8); b[at + 3] = (uint8_t)n;
    }
    void matrix() { u32(0x00010000); u32(0); u32(0); u32(0); u32(0x00010000); u32(0); u32(0); u32(0); u32(0x40000000); }
};

inline const char *codec_name(uint32_t fourcc)
{
    switch (fourcc & 0xFF) {
    case '1': return "Hap";
    case '5': return "Hap Alpha";
    case 'Y': return "Hap Q";
    case 'M': return "Hap Q Alpha";
    case 'A': return "Hap Alpha-Only";
    case '7': return "Hap R";
    default: return "Hap HDR";
    }
}

inline HapB200Mov *create(const char *path, uint32_t fourcc, uint32_t width, uint32_t height, uint32_t timescale)
{
    if (!path || !is_hap_fourcc(fourcc) || width == 0 || height == 0 || width > 0xFFFF || height > 0xFFFF || timescale == 0) return nullptr;
    FILE *f = fopen(path, "wb");
    if (!f) return nullptr;
    HapB200Mov *m = new HapB200Mov;
    m->f = f;
    m->writing = true;
    m->fourcc = fourcc; m->width = width; m->height = height; m->timescale = timescale;
    Buf h;
    size_t a = h.begin(HAPB200_FOURCC('f', 't', 'y', 'p'));
    h.u32(HAPB200_FOURCC('q', 't', ' ', ' ')); h.u32(0x00000200); h.u32(HAPB200_FOURCC('q', 't', ' ', ' '));
    h.end(a);
    m->mdat_header_at = h.b.size();
    h.u32(1); h.u32(HAPB200_FOURCC('m', 'd', 'a', 't')); h.u64(0);   // 64-bit size, patched when the file is closed
    if (fwrite(h.b.data(), 1, h.b.size(), f) != h.b.size()) { fclose(f); delete m; return nullptr; }
    m->write_pos = h.b.size();
    return m;
}

inline unsigned int write_frame(HapB200Mov *m, const void *frame, unsigned long n, unsigned int ticks)
{
    if (!m || !m->writing || !frame || n == 0 || n > 0xFFFFFFFFul) return HapResult_Bad_Arguments;
    if (m->io_error || m->size.size() >= kMaxFrames) return HapResult_Internal_Error;
    if (fwrite(frame, 1, n, m->f) != n) { m->io_error = true; return HapResult_Internal_Error; }
    m->offset.push_back(m->write_pos);
    m->size.push_back((uint32_t)n);
    m->ticks.push_back(ticks);
    m->write_pos += n;
    m->duration += ticks;
    return HapResult_No_Error;
}

inline unsigned int finish(HapB200Mov *m)
{
    const uint32_t n = (uint32_t)m->size.size();
    // movie duration in the movie timescale = the media timescale here
    const uint64_t dur64 = m->duration;
    const uint32_t dur = dur64 > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)dur64;
    const bool big = m->write_pos > 0xFFFFFFFFull;
    Buf q;
    const size_t moov = q.begin(HAPB200_FOURCC('m', 'o', 'o', 'v'));
    {
        const size_t a = q.begin(HAPB200_FOURCC('m', 'v', 'h', 'd'));
        q.u32(0); q.u32(0); q.u32(0); q.u32(m->timescale); q.u32(dur); q.u32(0x00010000); q.u16(0x0100); q.zeros(10);
        q.matrix(); q.zeros(24); q.u32(2);
        q.end(a);
    }
    const size_t trak = q.begin(HAPB200_FOURCC('t', 'r', 'a', 'k'));
    {
        const size_t a = q.begin(HAPB200_FOURCC('t', 'k', 'h', 'd'));
        q.u32(0x0000000F); q.u32(0); q.u32(0); q.u32(1); q.u32(0); q.u32(dur); q.zeros(8); q.u16(0); q.u16(0); q.u16(0); q.u16(0);
        q.matrix(); q.u32(m->width << 16); q.u32(m->height << 16);
        q.end(a);
    }
    const size_t mdia = q.begin(HAPB200_FOURCC('m', 'd', 'i', 'a'));
    {
        size_t a = q.begin(HAPB200_FOURCC('m', 'd', 'h', 'd'));
        q.u32(0); q.u32(0); q.u32(0); q.u32(m->timescale); q.u32(dur); q.u16(0); q.u16(0);
        q.end(a);
        a = q.begin(HAPB200_FOURCC('h', 'd', 'l', 'r'));
        q.u32(0); q.u32(HAPB200_FOURCC('m', 'h', 'l', 'r')); q.u32(HAPB200_FOURCC('v', 'i', 'd', 'e')); q.u32(0); q.u32(0); q.u32(0);
        q.pstr("VideoHandler", 13);
        q.end(a);
    }
    const size_t minf = q.begin(HAPB200_FOURCC('m', 'i', 'n', 'f'));
    {
        size_t a = q.begin(HAPB200_FOURCC('v', 'm', 'h', 'd'));
        q.u32(1); q.u16(0x0040); q.u16(0x8000); q.u16(0x8000); q.u16(0x8000);
        q.end(a);
        a = q.begin(HAPB200_FOURCC('h', 'd', 'l', 'r'));
        q.u32(0); q.u32(HAPB200_FOURCC('d', 'h', 'l', 'r')); q.u32(HAPB200_FOURCC('a', 'l', 'i', 's')); q.u32(0); q.u32(0); q.u32(0);
        q.pstr("DataHandler", 12);
        q.end(a);
        a = q.begin(HAPB200_FOURCC('d', 'i', 'n', 'f'));
        const size_t d = q.begin(HAPB200_FOURCC('d', 'r', 'e', 'f'));
        q.u32(0); q.u32(1); q.u32(12); q.u32(HAPB200_FOURCC('a', 'l', 'i', 's')); q.u32(1);   // one self-reference
        q.end(d);
        q.end(a);
    }
    const size_t stbl = q.begin(HAPB200_FOURCC('s', 't', 'b', 'l'));
    {
        size_t a = q.begin(HAPB200_FOURCC('s', 't', 's', 'd'));
        q.u32(0); q.u32(1);
        const size_t e = q.b.size();
        q.u32(86); q.u32(m->fourcc); q.zeros(6); q.u16(1);                 // image description, data reference 1
        q.u16(0); q.u16(0); q.u32(HAPB200_FOURCC('V', 'D', 'V', 'X'));     // version, revision, vendor
        q.u32(0); q.u32(0x00000200);                                       // temporal / spatial quality
        q.u16(m->width); q.u16(m->height); q.u32(0x00480000); q.u32(0x00480000);
        q.u32(0); q.u16(1);                                                // data size, frames per sample
        q.pstr(codec_name(m->fourcc), 32);
        const bool alpha = (m->fourcc & 0xFF) == '5' || (m->fourcc & 0xFF) == 'M' || (m->fourcc & 0xFF) == 'A' || (m->fourcc & 0xFF) == '7';
        q.u16(alpha ? 32 : 24); q.u16(0xFFFF);                             // depth, no colour table
        (void)e;
        q.end(a);
        a = q.begin(HAPB200_FOURCC('s', 't', 't', 's'));
        q.u32(0);
        const size_t cnt_at = q.b.size();
        q.u32(0);
        uint32_t runs = 0;
        for (uint32_t i = 0; i < n;) {
            uint32_t j = i;
            while (j < n && m->ticks[j] == m->ticks[i]) j++;
            q.u32(j - i); q.u32(m->ticks[i]);
            runs++;
            i = j;
        }
        q.b[cnt_at] = (uint8_t)(runs >> 24); q.b[cnt_at + 1] = (uint8_t)(runs >> 16); q.b[cnt_at + 2] = (uint8_t)(runs >> 8); q.b[cnt_at + 3] = (uint8_t)runs;
        q.end(a);
        a = q.begin(HAPB200_FOURCC('s', 't', 's', 'c'));
        q.u32(0);
        if (n) { q.u32(1); q.u32(1); q.u32(1); q.u32(1); } else q.u32(0);   // every sample is its own chunk
        q.end(a);
        a = q.begin(HAPB200_FOURCC('s', 't', 's', 'z'));
        q.u32(0); q.u32(0); q.u32(n);
        for (uint32_t i = 0; i < n; i++) q.u32(m->size[i]);
        q.end(a);
        if (big) {
            a = q.begin(HAPB200_FOURCC('c', 'o', '6', '4'));
            q.u32(0); q.u32(n);
            for (uint32_t i = 0; i < n; i++) q.u64(m->offset[i]);
        } else {
            a = q.begin(HAPB200_FOURCC('s', 't', 'c', 'o'));
            q.u32(0); q.u32(n);
            for (uint32_t i = 0; i < n; i++) q.u32((uint32_t)m->offset[i]);
        }
        q.end(a);
    }
    q.end(stbl); q.end(minf); q.end(mdia); q.end(trak); q.end(moov);
    if (q.b.size() > 0xFFFFFFFFull) return HapResult_Internal_Error;
    bool ok = !m->io_error && fwrite(q.b.data(), 1, q.b.size(), m->f) == q.b.size();
    // patch the 64-bit mdat size
    Buf sz;
    sz.u64(m->write_pos - m->mdat_header_at);
    ok = ok && fseeko(m->f, (off_t)(m->mdat_header_at + 8), SEEK_SET) == 0 && fwrite(sz.b.data(), 1, 8, m->f) == 8;
    ok = ok && fflush(m->f) == 0;
    return ok ? HapResult_No_Error : HapResult_Internal_Error;
}

}  // namespace hapmov
