// hap_b200/csrc/snappy_decode.cuh -- K7: per-chunk second-stage decompressor.
//
// Replaces hap_decode_chunk + snappy_uncompress (/root/reference/source/hap.c:606-642, call sites
// :612 and :899): one chunk of a Hap frame -> its decoded DXT bytes at a fixed destination.
// A chunk is one raw Snappy stream (compressor byte 0x0B) or a verbatim copy (0x0A).
//
// One CTA per chunk.  Snappy is serial inside a stream (an element's position depends on the
// length of every element before it, and copies read earlier output), so the kernel breaks both
// chains explicitly, window by window (16 KiB of compressed bytes each), and it splits the CTA
// into two warp groups that work on DIFFERENT windows at the same time:
//
//   PARSE group (2 warps), window k+1:
//   1. Stage the window into shared memory (realigned), then find the element boundaries without a
//      serial walk over elements: for EVERY offset o of every 64-byte sub-block, where does an
//      element chain entering at o leave the sub-block (one backward sweep per sub-block out of
//      registers, x[o] = x[o + length(o)], a byte table in shared memory).  One thread then follows
//      the true chain sub-block by sub-block with one table look-up per hop (long literals jump over
//      whole sub-blocks), and each entered sub-block is walked once from its true entry.
//      (A first version guessed entries and iterated to a fixpoint; on literal-heavy streams wrong
//      guesses do not re-synchronise and it needed ~one round per sub-block.)
//   2. Scan element counts / output bytes -> every element's descriptor slot and destination.
//   3. Write the descriptors; mark runs of adjacent copies with one offset (how every encoder emits
//      a long or overlapping match): they become independent periodic fills of the run's base period.
//
//   EXECUTE group (6 warps), window k:
//   4. FLATTEN.  DXT payloads are full of copy-of-copy chains ("same as the previous block except a
//      few bytes"); a copy whose source lies inside one earlier element takes over that element's
//      source (pointer jumping), which ends at input bytes or at earlier windows.
//   5. EXECUTE.  Literals are independent (source = the staged input).  What is left goes in
//      dependency rounds: a copy runs once every element overlapping its source range finished.
//
// The two groups hand windows over through two buffers (staged input + descriptors each) and four
// named barriers (full / empty per buffer): the parse group ARRIVES on "full" and goes on to the next
// window, the execute group SYNCS on it, and the other way round for "empty".  Before this split the
// serial chain hop alone kept 255 threads waiting for a fifth of the kernel (ncu: 37 % of all warp
// samples at barriers); now that wait overlaps the byte moves of the previous window.
// Every decision is taken on device; the host only reads one status word per chunk.
#pragma once
#include "block_primitives.cuh"
#include "hap_codes.h"
#include "simt.h"

namespace hapb200 {

struct ChunkJob {
    const uint8_t *src;   // compressed (0x0B) or raw (0x0A) chunk bytes
    uint8_t *dst;         // decoded bytes go here
    uint32_t src_bytes;
    uint32_t dst_bytes;   // decoded size the container arithmetic expects (hap.c:813, :833)
    uint32_t compressor;  // kHapCompressorNone 0x0A | kHapCompressorSnappy 0x0B (hap.c:41-42)
    uint32_t status;      // out: HapResult of this chunk (hap.c:617-640)
};

constexpr int kDecThreads = 256;
#ifndef HAPB200_DEC_PARSE_THREADS
#define HAPB200_DEC_PARSE_THREADS 64
#endif
constexpr int kDecParse = HAPB200_DEC_PARSE_THREADS;   // parse group: threads 0 .. kDecParse-1
constexpr int kDecExec = kDecThreads - kDecParse;      // execute group: the rest
constexpr int kDecExecWarps = kDecExec / 32;
constexpr int kDecSub = 64;                        // compressed bytes per sub-block
constexpr int kDecSubs = 256;                      // sub-blocks per window
constexpr int kDecSPT = kDecSubs / kDecParse;      // sub-blocks owned by one parse thread
constexpr int kDecWin = kDecSubs * kDecSub;        // 16 KiB of compressed input per window
constexpr int kDecMaxElems = 1024;                 // descriptors per window
static_assert(kDecParse % 32 == 0 && kDecExec % 32 == 0 && kDecSubs % kDecParse == 0, "whole warps; whole sub-blocks per thread");
constexpr uint32_t kSrcIn = 0u << 30, kSrcOut = 1u << 30, kSrcRun = 2u << 30, kSrcMask = 3u << 30, kPosMask = (1u << 30) - 1;
constexpr int kFlattenRounds = 2, kFlattenHops = 12;
constexpr uint32_t kLongLiteral = 1024;            // literals this long are copied by the whole execute group
constexpr int kMaxLong = 64;
constexpr int kMaxMid = 128;
constexpr uint32_t kExitMaxRel = 250;              // tbl value <= this: exit = sub-block end + value
constexpr uint32_t kExitFar = 253;                 // exit further away (a long literal): recomputed by walking
constexpr uint32_t kExitInvalid = 254;             // the chain runs into an invalid element header
constexpr uint32_t kNotKept = 0xFFFFFFFFu;
// named barriers (0 is __syncthreads)
constexpr int kBarParse = 1, kBarExec = 2, kBarFull = 3 /* +buffer */, kBarEmpty = 5 /* +buffer */;
constexpr uint32_t kWinData = 0, kWinEnd = 1, kWinFail = 2;

// One of the two hand-over buffers: the staged input of a window and its element descriptors.
struct alignas(16) DecodeWindow {
    // the exit tables are dead once every sub-block knows its entry, and the descriptors are born after that
    union {
        uint8_t tbl[kDecSub * kDecSubs];     // tbl[o][b]: where the chain entering sub-block b at offset o leaves it
        struct {
            uint32_t e_dst[kDecMaxElems];    // output offset inside the chunk
            uint32_t e_len[kDecMaxElems];
            uint32_t e_a[kDecMaxElems];      // packed source: kSrcIn|input position, kSrcOut|output position, kSrcRun|offset
            uint32_t e_b[kDecMaxElems];      // destination of the head of the element's same-offset run (its own, if alone)
        };
    };
    uint16_t e_done[kDecMaxElems];   // 0 = pending, r = finished in round r
    alignas(16) uint8_t cin[kDecWin + 64];   // staged window: aligned image of the input (+ alignment shift + header slack)
    uint32_t long_list[kMaxLong];    // literals of kLongLiteral bytes and more
    uint16_t mid_list[kMaxMid];      // literals of kThreadElem+1 .. kLongLiteral-1 bytes (one warp each)
    uint32_t n_long, n_mid;
    uint32_t total_e;                // descriptors in use
    uint32_t d0;                     // output bytes produced by earlier windows
    uint32_t wb;                     // input position of the window's first element
    uint32_t staged_end;             // input position up to which cin holds this window's bytes
    uint32_t status;                 // kWinData | kWinEnd | kWinFail
    uint32_t pad[3];
};
static_assert(sizeof(uint8_t[kDecSub * kDecSubs]) == 4 * sizeof(uint32_t[kDecMaxElems]), "tbl and the descriptors share storage");
static_assert(sizeof(DecodeWindow) % 16 == 0, "both buffers keep cin 16-byte aligned");

struct DecodeSmem {
    DecodeWindow win[2];
    // parse group's own state
    uint16_t entry[kDecSubs];        // true entry offset of each sub-block, 0xFFFF = jumped over
    uint32_t sub_out[kDecSubs];      // output bytes of the elements that start in the sub-block
    uint32_t sub_exit[kDecSubs];     // where its chain leaves it
    uint32_t sub_ebase[kDecSubs];    // first descriptor slot (kNotKept: beyond the descriptor arrays, next window)
    uint32_t sub_obase[kDecSubs];    // output offset of its first element inside the window
    uint8_t sub_cnt[kDecSubs];       // elements that start in it (<= 32)
    uint8_t sub_inv[kDecSubs];       // its chain runs into an invalid header
    uint32_t scratch[kDecParse / 32];
    uint32_t bcast[8];
    int fail;        // preamble / parse stage
    int fail_desc;   // descriptor stage
};

// Walk the element chain of one sub-block.  Positions are absolute inside the chunk input.
// Returns the first chain position >= blk_end (clamped to in_end when the chain is invalid).
struct WalkResult {
    uint32_t exit, count, out_bytes;
    int invalid;
};

__device__ __forceinline__ bool read_element_header(const uint8_t *cin, uint32_t wb, uint32_t pos, uint32_t in_end,
                                                    uint32_t &len, uint32_t &aux, uint32_t &hdr, uint32_t &kind)
{
    // cin[pos - wb .. +5) is always staged (zero beyond the input), so the loads below are in range
    const uint8_t *p = cin + (pos - wb);
    uint32_t tag = p[0];
    kind = tag & 3;
    if (kind == 0) {
        uint32_t m = tag >> 2;
        hdr = 1;
        if (m >= 60) {
            uint32_t extra = m - 59;
            if ((uint64_t)pos + 1 + extra > in_end) return false;
            uint32_t v = p[1] | (p[2] << 8) | (p[3] << 16) | ((uint32_t)p[4] << 24);
            m = extra == 4 ? v : (v & ((1u << (8 * extra)) - 1));
            hdr = 1 + extra;
        }
        if (m == 0xFFFFFFFFu) return false;
        len = m + 1;
        aux = pos + hdr;  // payload position
        return (uint64_t)pos + hdr + len <= in_end;
    }
    if (kind == 1) {
        hdr = 2;
        len = 4 + ((tag >> 2) & 7);
        aux = ((tag >> 5) << 8) | p[1];
    } else if (kind == 2) {
        hdr = 3;
        len = 1 + (tag >> 2);
        aux = p[1] | (p[2] << 8);
    } else {
        hdr = 5;
        len = 1 + (tag >> 2);
        aux = p[1] | (p[2] << 8) | (p[3] << 16) | ((uint32_t)p[4] << 24);
    }
    return (uint64_t)pos + hdr <= in_end;
}

__device__ __forceinline__ WalkResult walk_subblock(const uint8_t *cin, uint32_t wb, uint32_t entry, uint32_t blk_end,
                                                    uint32_t in_end)
{
    WalkResult r;
    r.count = 0;
    r.out_bytes = 0;
    r.invalid = 0;
    uint32_t pos = entry;
    while (pos < blk_end) {
        uint32_t len, aux, hdr, kind;
        if (!read_element_header(cin, wb, pos, in_end, len, aux, hdr, kind)) {
            r.invalid = 1;
            pos = in_end;
            break;
        }
        r.count++;
        r.out_bytes += len;
        pos += hdr + (kind == 0 ? len : 0);
    }
    r.exit = pos;
    return r;
}

// dst/src any alignment.  `lane` of `n_lanes` cooperating threads; 4 bytes per thread per step once the
// destination is word-aligned, the source word assembled from two aligned words when it is not.
template <int N_LANES>
__device__ __forceinline__ void lanes_copy(uint8_t *dst, const uint8_t *src, uint32_t len, uint32_t lane)
{
    uint32_t head = (uint32_t)((4 - ((uintptr_t)dst & 3)) & 3);
    if (head > len) head = len;
    if (lane < head) dst[lane] = src[lane];
    const uint32_t body = len - head;
    uint32_t nw = body >> 2;
    const uint8_t *s2 = src + head;
    uint32_t *d32 = reinterpret_cast<uint32_t *>(dst + head);
    const uint32_t mis = (uint32_t)((uintptr_t)s2 & 3);
    if (mis == 0) {
        const uint32_t *s32 = reinterpret_cast<const uint32_t *>(s2);
        for (uint32_t k = lane; k < nw; k += N_LANES) d32[k] = s32[k];
    } else {
        // word k needs aligned words k and k+1; the last one would read past the source: leave it to the tail
        const uint32_t *s32 = reinterpret_cast<const uint32_t *>(s2 - mis);
        if (nw) nw -= 1;
        for (uint32_t k = lane; k < nw; k += N_LANES) d32[k] = __funnelshift_r(s32[k], s32[k + 1], 8 * mis);
    }
    const uint32_t done = head + (nw << 2);
    for (uint32_t i = done + lane; i < len; i += N_LANES) dst[i] = src[i];
}
// Up to kSmallElem (64) bytes by one thread.  Word path: every load is issued before the first store, so the
// loads overlap instead of each waiting behind the store before it (the compiler must assume they alias).
constexpr uint32_t kSmallElem = 64;
constexpr uint32_t kThreadElem = 256;  // literals up to this long are also moved by one thread (their source is shared memory)
constexpr uint32_t kStageWords = 8;  // words held in registers at a time (two passes cover 64 bytes)
__device__ __forceinline__ void small_copy(uint8_t *d, const uint8_t *s, uint32_t len)
{
    if ((((uintptr_t)d | (uintptr_t)s | len) & 3) == 0) {
        const uint32_t *s32 = reinterpret_cast<const uint32_t *>(s);
        uint32_t *d32 = reinterpret_cast<uint32_t *>(d);
        const uint32_t nw = len >> 2;
#pragma unroll 1
        for (uint32_t b = 0; b < nw; b += kStageWords) {
            uint32_t v[kStageWords];
#pragma unroll
            for (uint32_t k = 0; k < kStageWords; k++)
                if (b + k < nw) v[k] = s32[b + k];
#pragma unroll
            for (uint32_t k = 0; k < kStageWords; k++)
                if (b + k < nw) d32[b + k] = v[k];
        }
    } else if ((((uintptr_t)d | len) & 3) == 0) {
        // destination aligned, source not: each word from two aligned source words
        const uint32_t mis = (uint32_t)((uintptr_t)s & 3);
        const uint32_t *s32 = reinterpret_cast<const uint32_t *>(s - mis);
        uint32_t *d32 = reinterpret_cast<uint32_t *>(d);
        const uint32_t nw = len >> 2;
        // the word after the last whole source word may hold only bytes before the element's end: bytewise
        uint32_t last = 0;
        for (uint32_t q = 0; q < mis; q++) last |= (uint32_t)s[len - mis + q] << (8 * q);
#pragma unroll 1
        for (uint32_t b = 0; b < nw; b += kStageWords) {
            uint32_t v[kStageWords + 1];
#pragma unroll
            for (uint32_t k = 0; k <= kStageWords; k++)
                v[k] = b + k < nw ? s32[b + k] : last;
#pragma unroll
            for (uint32_t k = 0; k < kStageWords; k++)
                if (b + k < nw) d32[b + k] = __funnelshift_r(v[k], v[k + 1], 8 * mis);
        }
    } else {
        // any alignment (every stream a byte-granular encoder such as Google's produces): 32 bytes per pass; all
        // source bytes are fetched first -- whole aligned words, plus single bytes for the two words that stick out
        // at the ends -- then written; a byte-by-byte copy would wait for each load behind the store before it
#pragma unroll 1
        for (uint32_t base = 0; base < len; base += 32) {
            const uint32_t n = len - base < 32u ? len - base : 32u;
            const uint8_t *sp = s + base;
            const uint32_t mis = (uint32_t)((uintptr_t)sp & 3);
            const uint32_t *s32 = reinterpret_cast<const uint32_t *>(sp - mis);
            const uint32_t nwords = (mis + n + 3) >> 2;           // aligned words touched, <= 9
            uint32_t v[10];
#pragma unroll
            for (uint32_t k = 0; k < 9; k++) {
                v[k] = 0;
                if (k < nwords) {
                    const bool inner = (k > 0 || mis == 0) && (4 * (k + 1) <= mis + n);  // every byte of the word is wanted
                    if (inner) {
                        v[k] = s32[k];
                    } else {
#pragma unroll
                        for (uint32_t q = 0; q < 4; q++) {
                            const uint32_t pos = 4 * k + q;  // byte position relative to the aligned base
                            if (pos >= mis && pos < mis + n) v[k] |= (uint32_t)sp[pos - mis] << (8 * q);
                        }
                    }
                }
            }
            v[9] = 0;
            uint32_t u[8];
#pragma unroll
            for (uint32_t k = 0; k < 8; k++) u[k] = __funnelshift_r(v[k], v[k + 1], 8 * mis);
            uint8_t *dp = d + base;
            if ((((uintptr_t)dp | n) & 3) == 0) {
                uint32_t *d32 = reinterpret_cast<uint32_t *>(dp);
#pragma unroll
                for (uint32_t k = 0; k < 8; k++)
                    if (4 * k < n) d32[k] = u[k];
            } else {
#pragma unroll
                for (uint32_t i = 0; i < 32; i++)
                    if (i < n) dp[i] = (uint8_t)(u[i >> 2] >> (8 * (i & 3)));
            }
        }
    }
}


// 16 aligned bytes at p, but only the bytes inside [lo, hi) are read (the others come back as zero)
__device__ __forceinline__ uint4 load16_inside(const uint4 *p, uintptr_t lo, uintptr_t hi)
{
    const uintptr_t a = (uintptr_t)p;
    if (a >= lo && a + 16 <= hi) return *p;
    uint32_t w[4] = {0, 0, 0, 0};
    const uint8_t *b = reinterpret_cast<const uint8_t *>(p);
    for (int k = 0; k < 16; k++)
        if (a + k >= lo && a + k < hi) w[k >> 2] |= (uint32_t)b[k] << (8 * (k & 3));
    return make_uint4(w[0], w[1], w[2], w[3]);
}

#ifdef HAPB200_EMU
__device__ __forceinline__ void hap_prefetch_l2(const void *) {}
#else
__device__ __forceinline__ void hap_prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#endif




#ifdef HAPB200_DECODE_PHASE_CYCLES
__device__ unsigned long long g_decode_phase_cycles[8];
__device__ unsigned long long g_decode_counts[8];  // windows, elements, execute rounds, -, -, parse-group wait cycles, execute-group wait cycles
#define COUNT_ADD(i, v) do { if (lead) atomicAdd(&g_decode_counts[i], (unsigned long long)(v)); } while (0)
#define PHASE_MARK(i) do { if (lead) { long long now_ = clock64(); atomicAdd(&g_decode_phase_cycles[i], (unsigned long long)(now_ - phase_t0_)); phase_t0_ = now_; } } while (0)
#define PHASE_WAIT(i) do { if (lead) { long long now_ = clock64(); atomicAdd(&g_decode_counts[i], (unsigned long long)(now_ - phase_t0_)); phase_t0_ = now_; } } while (0)
#define PHASE_INIT long long phase_t0_ = clock64()
#else
#define PHASE_MARK(i) do { } while (0)
#define PHASE_WAIT(i) do { } while (0)
#define PHASE_INIT do { } while (0)
#define COUNT_ADD(i, v) do { } while (0)
#endif

// Exclusive prefix maximum over the parse group (identity 0).  Ends with a group barrier.
__device__ __forceinline__ uint32_t parse_excl_max(uint32_t v, uint32_t *scratch)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(HAP_FULL_MASK, incl, d);
        if (lane >= d) incl = incl > o ? incl : o;
    }
    uint32_t prev = __shfl_up_sync(HAP_FULL_MASK, incl, 1);
    if (lane == 0) prev = 0;
    if (lane == 31) scratch[warp] = incl;
    hap_bar_sync(kBarParse, kDecParse);
    uint32_t base = 0;
#pragma unroll
    for (int w = 0; w < kDecParse / 32; w++) {
        uint32_t s = scratch[w];
        if (w < warp) base = base > s ? base : s;
    }
    hap_bar_sync(kBarParse, kDecParse);
    return base > prev ? base : prev;
}

// ---- PARSE group: stage, exit tables, chain hop, walk, scans, descriptors, same-offset runs ------------------
// gt = thread index inside the group.  Produces windows into S.win[0], S.win[1], S.win[0], ... and ends the
// sequence with a window whose status is kWinEnd or kWinFail.  Returns the number of data windows produced;
// *d0_out = output bytes they describe.
__device__ __forceinline__ uint32_t decode_parse_group(DecodeSmem &S, const uint32_t gt, const uint8_t *__restrict__ src,
                                                        const uint32_t in_end, const uint32_t expected, uint32_t wb,
                                                        uint32_t *d0_out)
{
    const bool lead = gt == 0;
    PHASE_INIT;
    uint32_t d0 = 0;           // output bytes described by earlier windows
    // Sub-blocks covered per window.  Dense streams (a few bytes per element, e.g. Google Snappy on DXT5) fill the
    // descriptor arrays long before 256 sub-blocks are used; the span then shrinks so that no exit table is computed
    // for bytes this window never reaches, and grows back when windows stop being cut short.
    uint32_t span = kDecSubs;
    uint32_t k = 0;
    bool failed = false;
    for (;; k++) {
        DecodeWindow &Wn = S.win[k & 1];
        if (k >= 2) hap_bar_sync(kBarEmpty + (int)(k & 1), kDecThreads);   // the execute group is done with this buffer
        PHASE_WAIT(5);
        if (wb >= in_end) {
            if (lead) Wn.status = kWinEnd;
            __threadfence_block();
            hap_bar_arrive(kBarFull + (int)(k & 1), kDecThreads);
            break;
        }
        // ---- stage the window.  Chunks are byte-packed in a frame, so the chunk is rarely aligned: read aligned
        //      16-byte words and shift them so that cin[0] is the byte at `wb` (word loads stay aligned later) ---
        if (lead) { Wn.n_long = 0; Wn.n_mid = 0; }
        uint32_t staged_end;  // input position up to which cin holds this window's bytes
        {
            const uintptr_t gaddr = (uintptr_t)(src + wb);
            const uint32_t shift = (uint32_t)(gaddr & 15);           // uniform over the group
            const uint32_t want = span * kDecSub + 16;
            const uint32_t avail = in_end - wb < want ? in_end - wb : want;
            staged_end = wb + avail;
            const uint32_t n16 = (avail + 15) >> 4;
            const uint4 *g4 = reinterpret_cast<const uint4 *>(gaddr - shift);
            uint4 *s4 = reinterpret_cast<uint4 *>(Wn.cin);
            const uint32_t ws = shift >> 2, bs = (shift & 3) * 8;
            // aligned words are only read whole when every byte of them belongs to the chunk; the (at most two) words
            // that stick out at the chunk's ends are gathered bytewise, so nothing outside [src, src + in_end) is touched
            const uintptr_t c_lo = (uintptr_t)src, c_hi = (uintptr_t)src + in_end;
            for (uint32_t i = gt; i < n16; i += kDecParse) {
                const uint4 a = load16_inside(g4 + i, c_lo, c_hi);
                uint4 b = make_uint4(0, 0, 0, 0);
                if (shift != 0 && 16 * (i + 1) < shift + avail) b = load16_inside(g4 + i + 1, c_lo, c_hi);
                uint32_t w0, w1, w2, w3, w4;
                if (ws == 0) { w0 = a.x; w1 = a.y; w2 = a.z; w3 = a.w; w4 = b.x; }
                else if (ws == 1) { w0 = a.y; w1 = a.z; w2 = a.w; w3 = b.x; w4 = b.y; }
                else if (ws == 2) { w0 = a.z; w1 = a.w; w2 = b.x; w3 = b.y; w4 = b.z; }
                else { w0 = a.w; w1 = b.x; w2 = b.y; w3 = b.z; w4 = b.w; }
                s4[i] = make_uint4(__funnelshift_r(w0, w1, bs), __funnelshift_r(w1, w2, bs), __funnelshift_r(w2, w3, bs),
                                   __funnelshift_r(w3, w4, bs));
            }
            // next window's lines on their way into L2 while this one is parsed
            for (uint32_t q = gt; q < (uint32_t)kDecWin / 128; q += kDecParse) {
                const uint64_t pf = (uint64_t)wb + kDecWin + (uint64_t)q * 128;
                if (pf < in_end) hap_prefetch_l2(src + pf);
            }
        }
        const uint8_t *cinp = Wn.cin;
        for (uint32_t b = gt; b < (uint32_t)kDecSubs; b += kDecParse) S.entry[b] = 0xFFFFu;
        hap_bar_sync(kBarParse, kDecParse);

        PHASE_MARK(0);
        // ---- 1. parse.  (a) for EACH of the 64 offsets of a sub-block, where does an element chain entering there
        //      leave the sub-block?  One backward sweep: x[o] = x[o + length(o)].  A thread sweeps kDecSPT sub-blocks.
        //      (b) one thread hops sub-block to sub-block along the true chain using that table.
        //      (c) every sub-block the chain enters is walked once from its true entry. ------------------
#pragma unroll 1
        for (uint32_t j = 0; j < (uint32_t)kDecSPT; j++) {
            const uint32_t b = j * kDecParse + gt;   // lanes take neighbouring sub-blocks: table columns are conflict-free
            // sub-blocks beyond the span are not part of this window
            const uint32_t blk_start = b < span && (uint64_t)wb + (uint64_t)b * kDecSub < in_end ? wb + b * kDecSub : in_end;
            const uint32_t blk_end = b < span && (uint64_t)wb + (uint64_t)(b + 1) * kDecSub < in_end ? wb + (b + 1) * kDecSub : in_end;
            if (blk_start >= in_end) continue;
            // the sub-block's 64 bytes (+ 4 bytes of header look-ahead) live in registers: the sweep is fully
            // unrolled, so every tag byte is a compile-time extraction and only the table access touches memory
            const uint32_t blk_len = blk_end - blk_start;
            const uint32_t limit = in_end - blk_start;  // a chain position may not pass this
            uint32_t w[18];
            const uint32_t *c32 = reinterpret_cast<const uint32_t *>(Wn.cin + (size_t)b * kDecSub);
#pragma unroll
            for (int q = 0; q < 18; q++) w[q] = c32[q];
#pragma unroll
            for (int o = kDecSub - 1; o >= 0; o--) {
                const uint32_t tag = (w[o >> 2] >> (8 * (o & 3))) & 0xFFu;
                const uint32_t kind = tag & 3u;
                uint64_t nxt;
                if (kind != 0) {
                    nxt = (uint32_t)o + ((0x5320u >> (4 * kind)) & 0xFu);  // copy headers: 2, 3 or 5 bytes
                } else {
                    const uint32_t m = tag >> 2;
                    if (m < 60) {
                        nxt = (uint32_t)o + m + 2;                           // tag + (m+1) literal bytes
                    } else {
                        const uint32_t extra = m - 59;
                        // the 4 bytes after the tag, assembled from registers
                        const uint32_t lo_w = w[(o + 1) >> 2], hi_w = w[((o + 1) >> 2) + 1];
                        const uint32_t v = __funnelshift_r(lo_w, hi_w, 8 * ((o + 1) & 3));
                        const uint32_t mm = extra == 4 ? v : (v & ((1u << (8 * extra)) - 1u));
                        nxt = mm == 0xFFFFFFFFu ? ~0ull : (uint64_t)o + 1 + extra + (uint64_t)mm + 1;
                    }
                }
                uint32_t x;
                if (nxt > limit) x = kExitInvalid;               // header or payload runs past the input
                else if (nxt < blk_len) x = Wn.tbl[(uint32_t)nxt * kDecSubs + b];
                else x = nxt - blk_len <= kExitMaxRel ? (uint32_t)(nxt - blk_len) : kExitFar;
                Wn.tbl[(uint32_t)o * kDecSubs + b] = (uint8_t)x;
            }
        }
        hap_bar_sync(kBarParse, kDecParse);
        PHASE_MARK(5);
        if (gt == 0) {
            // window-relative positions: the dependent chain per hop is one table load plus a few ALU ops
            const uint32_t rel_end = in_end - wb;
            uint32_t rel = 0;
            while (rel < rel_end && rel < span * kDecSub) {
                const uint32_t blk = rel >> 6, o = rel & 63;
                const uint32_t x = Wn.tbl[o * kDecSubs + blk];
                S.entry[blk] = (uint16_t)o;
                uint32_t bend = (blk + 1) << 6;
                bend = bend < rel_end ? bend : rel_end;
                if (x <= kExitMaxRel) {
                    rel = bend + x;
                } else if (x == kExitInvalid) {
                    S.fail = 1;
                    break;
                } else {
                    // a long literal leaves this sub-block by more than a byte can hold: walk to it
                    uint32_t p2 = wb + rel;
                    for (;;) {
                        uint32_t len, aux, hdr, kind;
                        if (!read_element_header(cinp, wb, p2, in_end, len, aux, hdr, kind)) { S.fail = 1; p2 = in_end; break; }
                        p2 += hdr + (kind == 0 ? len : 0);
                        if (p2 - wb >= bend) break;
                    }
                    rel = p2 - wb;
                }
            }
        }
        hap_bar_sync(kBarParse, kDecParse);
        PHASE_MARK(6);
#pragma unroll 1
        for (uint32_t j = 0; j < (uint32_t)kDecSPT; j++) {
            const uint32_t b = j * kDecParse + gt;
            const uint32_t blk_start = b < span && (uint64_t)wb + (uint64_t)b * kDecSub < in_end ? wb + b * kDecSub : in_end;
            const uint32_t blk_end = b < span && (uint64_t)wb + (uint64_t)(b + 1) * kDecSub < in_end ? wb + (b + 1) * kDecSub : in_end;
            WalkResult w;
            w.exit = 0; w.count = 0; w.out_bytes = 0; w.invalid = 0;
            if (S.entry[b] != 0xFFFFu && blk_start < in_end) w = walk_subblock(cinp, wb, blk_start + S.entry[b], blk_end, in_end);
            S.sub_cnt[b] = (uint8_t)w.count;
            S.sub_out[b] = w.out_bytes;
            S.sub_exit[b] = w.exit;
            S.sub_inv[b] = (uint8_t)w.invalid;
        }
        hap_bar_sync(kBarParse, kDecParse);

        PHASE_MARK(1);
        // ---- 2. scans over the 256 sub-blocks (one warp, 8 sub-blocks per lane): descriptor slots, output offsets,
        //         and where the window is cut when it has more elements than descriptor slots -------------------
        if (gt < 32) {
            const uint32_t lane = gt, b0 = lane * (kDecSubs / 32);
            uint32_t csum = 0;
#pragma unroll
            for (int i = 0; i < kDecSubs / 32; i++) csum += S.sub_cnt[b0 + i];
            uint32_t cincl = csum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t o = __shfl_up_sync(HAP_FULL_MASK, cincl, d);
                if (lane >= (uint32_t)d) cincl += o;
            }
            const uint32_t total_e_all = __shfl_sync(HAP_FULL_MASK, cincl, 31);
            uint32_t e = cincl - csum, kept_cnt = 0, kept_out = 0, kept_exit = 0, kept_inv = 0;
            uint32_t keepbits = 0;
#pragma unroll
            for (int i = 0; i < kDecSubs / 32; i++) {
                const uint32_t c = S.sub_cnt[b0 + i];
                const bool keep = e + c <= (uint32_t)kDecMaxElems;
                S.sub_ebase[b0 + i] = keep ? e : kNotKept;
                if (keep) {
                    keepbits |= 1u << i;
                    kept_cnt += c;
                    kept_out += S.sub_out[b0 + i];
                    const uint32_t x = S.sub_exit[b0 + i];
                    kept_exit = kept_exit > x ? kept_exit : x;
                    kept_inv |= S.sub_inv[b0 + i];
                }
                e += c;
            }
            uint32_t oincl = kept_out, tot_cnt = kept_cnt, mx_exit = kept_exit, any_inv = kept_inv;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t o = __shfl_up_sync(HAP_FULL_MASK, oincl, d);
                if (lane >= (uint32_t)d) oincl += o;
            }
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) {
                tot_cnt += __shfl_xor_sync(HAP_FULL_MASK, tot_cnt, d);
                const uint32_t ox = __shfl_xor_sync(HAP_FULL_MASK, mx_exit, d);
                mx_exit = mx_exit > ox ? mx_exit : ox;
                any_inv |= __shfl_xor_sync(HAP_FULL_MASK, any_inv, d);
            }
            const uint32_t total_o = __shfl_sync(HAP_FULL_MASK, oincl, 31);
            uint32_t ob = oincl - kept_out;
#pragma unroll
            for (int i = 0; i < kDecSubs / 32; i++) {
                S.sub_obase[b0 + i] = ob;
                if (keepbits & (1u << i)) ob += S.sub_out[b0 + i];
            }
            if (lane == 0) {
                S.bcast[0] = total_e_all;
                S.bcast[1] = tot_cnt;
                S.bcast[2] = total_o;
                S.bcast[3] = mx_exit;
                // a kept sub-block whose chain is invalid poisons the stream (it is the true chain now)
                if (any_inv) S.fail = 1;
                if ((uint64_t)d0 + total_o > expected) S.fail = 1;
            }
        }
        hap_bar_sync(kBarParse, kDecParse);
        const uint32_t total_e_all = S.bcast[0], total_e = S.bcast[1], total_o = S.bcast[2], next_wb = S.bcast[3];
        if (S.fail) {
            // the stream is bad: this buffer (still ours) carries the verdict to the execute group
            failed = true;
            if (lead) Wn.status = kWinFail;
            __threadfence_block();
            hap_bar_arrive(kBarFull + (int)(k & 1), kDecThreads);
            break;
        }
        {
            const uint32_t used_sub = (next_wb - wb + kDecSub - 1) / kDecSub;
            if (total_e_all > (uint32_t)kDecMaxElems) span = used_sub + used_sub / 4 < 16u ? 16u : (used_sub + used_sub / 4 > (uint32_t)kDecSubs ? (uint32_t)kDecSubs : used_sub + used_sub / 4);
            else if (total_e_all < (uint32_t)kDecMaxElems / 2) span = span * 2 > (uint32_t)kDecSubs ? (uint32_t)kDecSubs : span * 2;
        }

        PHASE_MARK(2);
        // ---- 3. descriptors.  e_a packs the SOURCE of an element as (kind << 30) | position:
        //      kSrcIn  : bytes of the compressed input at `position` (literals, and copies flattened onto them)
        //      kSrcOut : bytes of the output at `position` (plain copies; offset >= length)
        //      kSrcRun : periodic fill with period `position` (= the offset) of the e_b[e] - offset .. e_b[e] bytes
        //      (the exit tables of this window are dead from here on: the descriptors take their place)
#pragma unroll 1
        for (uint32_t j = 0; j < (uint32_t)kDecSPT; j++) {
            const uint32_t b = j * kDecParse + gt;
            const uint32_t eb = S.sub_ebase[b];
            if (eb == kNotKept || S.entry[b] == 0xFFFFu) continue;
            const uint32_t blk_end = (uint64_t)wb + (uint64_t)(b + 1) * kDecSub < in_end ? wb + (b + 1) * kDecSub : in_end;
            uint32_t pos = wb + b * kDecSub + S.entry[b], e = eb, o = d0 + S.sub_obase[b];
            while (pos < blk_end) {
                uint32_t len, aux, hdr, kind;
                read_element_header(cinp, wb, pos, in_end, len, aux, hdr, kind);
                Wn.e_dst[e] = o;
                Wn.e_len[e] = len;
                Wn.e_done[e] = 0;
                if (kind == 0) {
                    Wn.e_a[e] = kSrcIn | aux;
                    Wn.e_b[e] = 0;  // literals break same-offset runs (a copy's offset is never 0)
                    if (len >= kLongLiteral) {
                        uint32_t q = atomicAdd(&Wn.n_long, 1u);
                        if (q < (uint32_t)kMaxLong) Wn.long_list[q] = e;
                    } else if (len > kThreadElem) {
                        uint32_t q = atomicAdd(&Wn.n_mid, 1u);
                        if (q < (uint32_t)kMaxMid) Wn.mid_list[q] = (uint16_t)e;
                    }
                    pos += hdr + len;
                } else {
                    if (aux == 0 || aux > o) S.fail_desc = 1;  // offset 0 or before the start of the output
                    Wn.e_a[e] = aux >= len ? (kSrcOut | (o - aux)) : (kSrcRun | aux);
                    Wn.e_b[e] = aux;  // the offset, for run detection below; becomes the run base afterwards
                    pos += hdr;
                }
                o += len;
                e++;
            }
        }
        hap_bar_sync(kBarParse, kDecParse);
        if (S.fail_desc) {
            failed = true;
            if (lead) Wn.status = kWinFail;
            __threadfence_block();
            hap_bar_arrive(kBarFull + (int)(k & 1), kDecThreads);
            break;
        }

        // ---- same-offset runs: a copy with the offset of the copy right before it continues that copy's
        //      match, so it is a periodic fill of the run head's base period, independent of its neighbours --
        {
            const uint32_t strip = (total_e + kDecParse - 1) / kDecParse;
            const uint32_t lo = gt * strip < total_e ? gt * strip : total_e;
            const uint32_t hi = lo + strip < total_e ? lo + strip : total_e;
            // Pass 1: flag continuations in the spare top bit of e_len (copies are at most 64 long) and find
            // the last run head of the strip.
            uint32_t last_head = 0;  // index + 1
            for (uint32_t e = lo; e < hi; e++) {
                const uint32_t off = Wn.e_b[e];
                const bool cont = e > 0 && off != 0 && off == Wn.e_b[e - 1];
                if (cont) Wn.e_len[e] |= 0x80000000u;
                else last_head = e + 1;
            }
            uint32_t head = parse_excl_max(last_head, S.scratch);  // ends with a group barrier
            // Pass 2: e_b becomes the base (destination of the run head); continuations become periodic fills.
            for (uint32_t e = lo; e < hi; e++) {
                const uint32_t off = Wn.e_b[e];
                if (Wn.e_len[e] & 0x80000000u) {
                    Wn.e_len[e] &= 0x7FFFFFFFu;
                    Wn.e_a[e] = kSrcRun | off;
                    Wn.e_b[e] = Wn.e_dst[head - 1];
                } else {
                    head = e + 1;
                    Wn.e_b[e] = Wn.e_dst[e];
                }
            }
        }
        if (lead) {
            Wn.total_e = total_e;
            Wn.d0 = d0;
            Wn.wb = wb;
            Wn.staged_end = staged_end;
            Wn.status = kWinData;
        }
        PHASE_MARK(3);
        COUNT_ADD(0, 1);
        COUNT_ADD(1, total_e);
        __threadfence_block();
        hap_bar_arrive(kBarFull + (int)(k & 1), kDecThreads);    // hand the window over; do not wait
        d0 += total_o;
        wb = next_wb;
    }
    // k = index of the terminal window = number of data windows.  The execute group's last "empty" arrival has no
    // taker yet: take it, which also means that every byte has been written.
    if (k >= 1) hap_bar_sync(kBarEmpty + (int)((k - 1) & 1), kDecThreads);
    *d0_out = d0;
    return failed ? 0xFFFFFFFFu : k;
}

// ---- EXECUTE group: flatten copy chains, then move the bytes ------------------------------------------------
// xt = thread index inside the group.
__device__ __forceinline__ void decode_execute_group(DecodeSmem &S, const uint32_t xt, const uint8_t *__restrict__ src,
                                                     uint8_t *__restrict__ dst)
{
    const bool lead = xt == 0;
    const uint32_t xw = xt >> 5;
    PHASE_INIT;
    for (uint32_t k = 0;; k++) {
        DecodeWindow &Wn = S.win[k & 1];
        hap_bar_sync(kBarFull + (int)(k & 1), kDecThreads);
        PHASE_WAIT(6);
        if (Wn.status != kWinData) break;
        const uint32_t total_e = Wn.total_e, d0 = Wn.d0, wb = Wn.wb, staged_end = Wn.staged_end;
        const uint8_t *cinp = Wn.cin;
        // ---- flatten copy-of-copy chains.  DXT payloads are full of "same as the previous block except a few
        //      bytes": a copy whose source is itself a copy, hundreds deep.  A plain copy whose source bytes lie
        //      inside ONE earlier element of this window takes over that element's source (pointer jumping on
        //      the packed e_a words; a racing update only makes the hop longer, never wrong).  Chains end at
        //      literals (-> read the input instead) or at earlier windows (-> already written). --------------
#pragma unroll 1
        for (int fr = 0; fr < kFlattenRounds; fr++) {
            for (uint32_t e = xt; e < total_e; e += kDecExec) {
                uint32_t a = Wn.e_a[e];
                if ((a & kSrcMask) != kSrcOut) continue;
                const uint32_t len = Wn.e_len[e];
                bool changed = false;
#pragma unroll 1
                for (int hop = 0; hop < kFlattenHops; hop++) {
                    const uint32_t sp = a & kPosMask;
                    if (sp + len <= d0) break;                 // reads finished output of earlier windows
                    if (sp < d0) break;                        // straddles the window start: leave it
                    uint32_t lo2 = 0, hi2 = e;                 // last element with e_dst <= sp (it is before e)
                    while (hi2 - lo2 > 1) {
                        const uint32_t m = (lo2 + hi2) >> 1;
                        if (Wn.e_dst[m] <= sp) lo2 = m; else hi2 = m;
                    }
                    const uint32_t f = lo2, fd = Wn.e_dst[f];
                    if (sp + len > fd + Wn.e_len[f]) break;    // spans several producers
                    const uint32_t fa = Wn.e_a[f];
                    if ((fa & kSrcMask) == kSrcRun) break;     // periodic producer: stay dependent on it
                    a = (fa & kSrcMask) | ((fa & kPosMask) + (sp - fd));
                    changed = true;
                    if ((fa & kSrcMask) == kSrcIn) break;      // landed on input bytes: fully resolved
                }
                if (changed) Wn.e_a[e] = a;
            }
            hap_bar_sync(kBarExec, kDecExec);
        }

        PHASE_MARK(7);
        // ---- execute: round 1 = everything whose source is the input or earlier windows; later rounds =
        //      copies whose producers finished in an earlier round.  Elements of at most 256 bytes (every copy,
        //      most literals) are moved by ONE THREAD each, staged through registers so that all its loads are
        //      in flight together; longer literals by a warp each; the longest by the whole group.
        for (uint32_t round = 1;; round++) {
            int pending = 0;
            for (uint32_t e = xt; e < total_e; e += kDecExec) {
                if (Wn.e_done[e] != 0) continue;
                const uint32_t len = Wn.e_len[e];
                if (len > kThreadElem) continue;                 // (copies are at most 64 bytes)
                const uint32_t a = Wn.e_a[e], o = Wn.e_dst[e];
                const uint32_t kind = a & kSrcMask, ap = a & kPosMask;
                uint8_t *d = dst + o;
                if (kind == kSrcIn) {
                    const uint8_t *sl = (ap >= wb && (uint64_t)ap + len <= staged_end) ? cinp + (ap - wb) : src + ap;
                    small_copy(d, sl, len);
                    Wn.e_done[e] = (uint16_t)round;
                    continue;
                }
                if (kind == kSrcOut) {
                    // A plain copy.  Its source bytes either lie in earlier windows (final), or they are the output of
                    // producers of this window.  It does not have to wait for those producers to RUN: a literal's
                    // bytes are in the input, a resolved copy's bytes are wherever that copy reads them -- so the
                    // source range is walked producer by producer and each piece is pulled from where it really is.
                    // Only a piece whose producer is itself unresolved (or periodic) has to wait for a later round.
                    bool ok = true;
                    uint32_t x = ap;
                    const uint32_t x_end = ap + len;
                    if (x < d0) {
                        const uint32_t n0 = x_end <= d0 ? len : d0 - x;
                        small_copy(d, dst + x, n0);
                        x += n0;
                    }
                    if (x < x_end) {
                        uint32_t lo2 = 0, hi2 = e;  // last element with e_dst <= x; the producer is before e
                        while (hi2 - lo2 > 1) {
                            uint32_t m = (lo2 + hi2) >> 1;
                            if (Wn.e_dst[m] <= x) lo2 = m; else hi2 = m;
                        }
                        for (uint32_t f = lo2; x < x_end; f++) {
                            const uint32_t fd = Wn.e_dst[f], fl = Wn.e_len[f];
                            const uint32_t x1 = x_end < fd + fl ? x_end : fd + fl;
                            const uint32_t fa = Wn.e_a[f], fk = fa & kSrcMask, fp = (fa & kPosMask) + (x - fd);
                            const uint32_t df = Wn.e_done[f];
                            const uint8_t *from;
                            if (df != 0 && df < round) from = dst + x;                      // producer already ran
                            else if (fk == kSrcIn) from = src + fp;                          // literal bytes: the input
                            else if (fk == kSrcOut && fp + (x1 - x) <= d0) from = dst + fp;  // resolved copy: its source
                            else { ok = false; break; }
                            small_copy(d + (x - ap), from, x1 - x);
                            x = x1;
                        }
                    }
                    if (!ok) { pending = 1; continue; }
                    Wn.e_done[e] = (uint16_t)round;
                    continue;
                }
                const uint32_t base = Wn.e_b[e];
                const uint32_t rel = o - base;  // position of this element inside its same-offset run
                uint32_t need_lo, need_hi;      // bytes this element reads
                if (rel + len <= ap) { need_lo = o - ap; need_hi = need_lo + len; }
                else { need_lo = base - ap; need_hi = base; }
                if (need_hi > d0) {
                    bool ready = true;
                    uint32_t x = need_lo > d0 ? need_lo : d0;
                    uint32_t lo2 = 0, hi2 = e;  // last element with e_dst <= x; the producer is before e
                    while (hi2 - lo2 > 1) {
                        uint32_t m = (lo2 + hi2) >> 1;
                        if (Wn.e_dst[m] <= x) lo2 = m; else hi2 = m;
                    }
                    for (uint32_t f = lo2; f < e && Wn.e_dst[f] < need_hi; f++) {
                        uint32_t df = Wn.e_done[f];
                        if (df == 0 || df >= round) { ready = false; break; }
                    }
                    if (!ready) { pending = 1; continue; }
                }
                if (rel + len <= ap) {
                    small_copy(d, dst + (o - ap), len);
                } else {
                    const uint32_t off = ap;
                    const uint8_t *period = dst + (base - off);
                    if (((off | rel | len) & 3) == 0 && (((uintptr_t)d | (uintptr_t)period) & 3) == 0) {
                        const uint32_t *p32 = reinterpret_cast<const uint32_t *>(period);
                        uint32_t *d32 = reinterpret_cast<uint32_t *>(d);
                        const uint32_t pw = off >> 2, nw = len >> 2;
                        uint32_t idx = (rel >> 2) % pw;
#pragma unroll 1
                        for (uint32_t b = 0; b < nw; b += kStageWords) {
                            uint32_t v[kStageWords];
#pragma unroll
                            for (uint32_t q = 0; q < kStageWords; q++)
                                if (b + q < nw) { v[q] = p32[idx]; idx = idx + 1 == pw ? 0 : idx + 1; }
#pragma unroll
                            for (uint32_t q = 0; q < kStageWords; q++)
                                if (b + q < nw) d32[b + q] = v[q];
                        }
                    } else {
                        uint32_t idx = rel % off;
                        for (uint32_t i = 0; i < len; i++) { d[i] = period[idx]; idx = idx + 1 == off ? 0 : idx + 1; }
                    }
                }
                Wn.e_done[e] = (uint16_t)round;
            }
            if (round == 1) {
                // literals of kThreadElem+1 .. 1023 bytes: one warp each, from the list the descriptor pass made
                const uint32_t nmid = Wn.n_mid;
                if (nmid <= (uint32_t)kMaxMid) {
                    for (uint32_t q = xw; q < nmid; q += kDecExecWarps) {
                        const uint32_t e = Wn.mid_list[q];
                        const uint32_t len = Wn.e_len[e];
                        const uint32_t ap = Wn.e_a[e] & kPosMask;  // only literals are this long
                        const uint8_t *sl = (ap >= wb && (uint64_t)ap + len <= staged_end) ? cinp + (ap - wb) : src + ap;
                        lanes_copy<32>(dst + Wn.e_dst[e], sl, len, xt & 31);
                        if ((xt & 31) == 0) Wn.e_done[e] = 1;
                    }
                } else {
                    // more of them than the list holds (cannot happen with 16 KiB of input per window, kept for safety)
                    for (uint32_t e = xw; e < total_e; e += kDecExecWarps) {
                        const uint32_t len = Wn.e_len[e];
                        if (len <= kThreadElem || len >= kLongLiteral) continue;
                        const uint32_t ap = Wn.e_a[e] & kPosMask;
                        const uint8_t *sl = (ap >= wb && (uint64_t)ap + len <= staged_end) ? cinp + (ap - wb) : src + ap;
                        lanes_copy<32>(dst + Wn.e_dst[e], sl, len, xt & 31);
                        if ((xt & 31) == 0) Wn.e_done[e] = 1;
                    }
                }
                // long literals: the whole group moves each one
                const uint32_t nlong = Wn.n_long < (uint32_t)kMaxLong ? Wn.n_long : (uint32_t)kMaxLong;
                for (uint32_t q = 0; q < nlong; q++) {
                    const uint32_t e = Wn.long_list[q];
                    lanes_copy<kDecExec>(dst + Wn.e_dst[e], src + (Wn.e_a[e] & kPosMask), Wn.e_len[e], xt);
                    if (xt == 0) Wn.e_done[e] = 1;
                }
                if (Wn.n_long > (uint32_t)kMaxLong) {
                    // overflow of the list (pathological): sweep the descriptors instead
                    for (uint32_t e = 0; e < total_e; e++)
                        if ((Wn.e_a[e] & kSrcMask) == kSrcIn && Wn.e_len[e] >= kLongLiteral && Wn.e_done[e] == 0) {
                            lanes_copy<kDecExec>(dst + Wn.e_dst[e], src + (Wn.e_a[e] & kPosMask), Wn.e_len[e], xt);
                            hap_bar_sync(kBarExec, kDecExec);
                            if (xt == 0) Wn.e_done[e] = 1;
                        }
                }
            }
            COUNT_ADD(2, 1);
            if (!hap_bar_or(kBarExec, kDecExec, pending)) break;
        }
        PHASE_MARK(4);
        __threadfence_block();
        hap_bar_arrive(kBarEmpty + (int)(k & 1), kDecThreads);   // the buffer may be refilled
    }
}

__global__ void __launch_bounds__(kDecThreads, 3) snappy_decode_chunks_kernel(ChunkJob *jobs, int njobs)
{
    HAP_DYN_SMEM(smem_raw);
    DecodeSmem &S = *reinterpret_cast<DecodeSmem *>(smem_raw);
    const int t = threadIdx.x;
    if ((int)blockIdx.x >= njobs) return;
    ChunkJob &job = jobs[blockIdx.x];
    const uint8_t *__restrict__ src = job.src;
    uint8_t *__restrict__ dst = job.dst;
    const uint32_t in_end = job.src_bytes;
    const uint32_t expected = job.dst_bytes;

    if (job.compressor == 0) return;  // unused slot of a batched frame (hap_parse.cuh)
    if (job.compressor == kHapChunkRaw) {
        // hap.c:630-636: verbatim chunk
        if (in_end != expected) {
            if (t == 0) job.status = HapResult_Bad_Frame;
            return;
        }
        const bool aligned = (((uintptr_t)src | (uintptr_t)dst) & 15) == 0;
        if (aligned) {
            const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
            uint4 *d4 = reinterpret_cast<uint4 *>(dst);
            uint32_t n4 = in_end >> 4;
            for (uint32_t i = t; i < n4; i += kDecThreads) d4[i] = s4[i];
            for (uint32_t i = (n4 << 4) + t; i < in_end; i += kDecThreads) dst[i] = src[i];
        } else {
            for (uint32_t i = t; i < in_end; i += kDecThreads) dst[i] = src[i];
        }
        if (t == 0) job.status = HapResult_No_Error;
        return;
    }
    if (job.compressor != kHapChunkSnappy || in_end > kPosMask || expected > kPosMask) {
        // hap.c:637-640; also chunks of 1 GiB and more, whose positions do not fit the packed descriptors
        if (t == 0) job.status = HapResult_Bad_Frame;
        return;
    }

    // ---- preamble: varint32 uncompressed length -----------------------------------------------
    if (t == 0) {
        uint64_t v = 0;
        uint32_t i = 0;
        bool ok = false;
        for (; i < 5 && i < in_end; i++) {
            uint32_t b = src[i];
            v |= (uint64_t)(b & 0x7F) << (7 * i);
            if (!(b & 0x80)) { ok = true; i++; break; }
        }
        S.fail = (!ok || v != (uint64_t)expected) ? 1 : 0;
        S.fail_desc = 0;
        S.bcast[4] = i;
    }
    __syncthreads();
    if (S.fail) {
        if (t == 0) job.status = HapResult_Bad_Frame;
        return;
    }
    const uint32_t wb0 = S.bcast[4];  // first window base: a true element start
    __syncthreads();

    if (t < kDecParse) {
        uint32_t d0 = 0;
        const uint32_t n = decode_parse_group(S, (uint32_t)t, src, in_end, expected, wb0, &d0);
        if (t == 0) job.status = (n == 0xFFFFFFFFu || d0 != expected) ? HapResult_Bad_Frame : HapResult_No_Error;
    } else {
        decode_execute_group(S, (uint32_t)(t - kDecParse), src, dst);
    }
}

}  // namespace hapb200
