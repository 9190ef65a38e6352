// hap_b200/csrc/snappy_decode.cuh -- K7: per-chunk second-stage decompressor.
//
// Replaces hap_decode_chunk + snappy_uncompress (/root/reference/source/hap.c:606-642, call sites
// :612 and :899): one chunk of a Hap frame -> its decoded DXT bytes at a fixed destination.
// A chunk is one raw Snappy stream (compressor byte 0x0B) or a verbatim copy (0x0A).
//
// One CTA per chunk.  Snappy is serial inside a stream (an element's position depends on the
// length of every element before it, and copies read earlier output), so the kernel breaks both
// chains explicitly, window by window over the compressed bytes:
//   1. PARSE without a serial walk over elements.  Thread t owns 64 compressed bytes and computes, for
//      EVERY offset o inside them, where an element chain entering at o leaves the sub-block
//      (one backward sweep, x[o] = x[o + length(o)], kept as a byte table in shared memory).  One thread
//      then follows the true chain sub-block by sub-block with one table look-up per hop (long literals
//      jump over whole sub-blocks), and each entered sub-block is walked once from its true entry.
//      (A first version guessed entries and iterated to a fixpoint; on literal-heavy streams wrong
//      guesses do not re-synchronise and it needed ~one round per sub-block: 56 % of the kernel.)
//   2. SCAN element counts / output bytes -> every element's destination offset.
//   3. FLATTEN + EXECUTE.  Literals are independent (source = compressed bytes).  Runs of adjacent copies
//      with one offset (how every encoder emits a long or overlapping match) become independent
//      periodic fills of the run's base period.  DXT payloads are full of copy-of-copy chains ("same as
//      the previous block except a few bytes"); a copy whose source lies inside one earlier element takes
//      over that element's source (pointer jumping), which ends at input bytes or at earlier windows, so
//      almost everything runs in the first round.  What is left goes in dependency rounds: a copy runs
//      once every element overlapping its source range finished in an earlier round.
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
constexpr int kDecSub = 64;                        // compressed bytes owned by one thread per window
constexpr int kDecWin = kDecThreads * kDecSub;     // 16 KiB of compressed input per window
constexpr int kDecMaxElems = 2048;                 // descriptors held in shared memory per window
constexpr uint32_t kSrcIn = 0u << 30, kSrcOut = 1u << 30, kSrcRun = 2u << 30, kSrcMask = 3u << 30, kPosMask = (1u << 30) - 1;
constexpr int kFlattenRounds = 1, kFlattenHops = 12;  // (a second flatten round was measured: same execution rounds, 1.5-2 % slower; none: 3.8 -> 7.9 rounds on Google-Snappy streams)
constexpr uint32_t kLongLiteral = 16384;           // literals this long are copied by the whole CTA, one after the other; shorter ones by a
                                                   // warp each (measured: 1024 here cost 7 % of the kernel -- the CTA-wide copies serialise)
constexpr int kMaxLong = 64;
constexpr int kMaxMid = 256;
constexpr uint32_t kMidPiece = 1024;               // a warp moves a literal in pieces of this many bytes (256: +2 %, 512: +1.5 % kernel time)
static_assert(kDecMaxElems <= 2048 && kLongLiteral / kMidPiece <= 32, "mid_list packs element (11 bits) and piece (5 bits)");
constexpr int kGroups = kDecThreads / 8;           // 8-lane groups, one element each
constexpr uint32_t kExitMaxRel = 186;              // tbl value <= this: exit = sub-block end + value
constexpr uint32_t kExitFarBase = 187;             // kExitFarBase + o (o = 0..63): the chain leaves through a long literal whose
                                                   // header sits at offset o of the sub-block; its end is read from that header
constexpr uint32_t kExitInvalid = 254;             // the chain runs into an invalid element header

struct DecodeSmem {
    uint32_t e_dst[kDecMaxElems];    // output offset inside the chunk
    uint32_t e_len[kDecMaxElems];
    uint32_t e_a[kDecMaxElems];      // packed source: kSrcIn|input position, kSrcOut|output position, kSrcRun|offset
    uint32_t e_b[kDecMaxElems];      // destination of the head of the element's same-offset run (its own, if alone)
    uint16_t e_done[kDecMaxElems];   // 0 = pending, r = finished in round r
    uint8_t cin[kDecWin + 64];       // staged window: aligned image of the input (+ alignment shift + header slack)
    uint8_t tbl[kDecSub * kDecThreads];  // tbl[o][t]: where the chain entering sub-block t at offset o leaves it
    uint16_t entry[kDecThreads];     // true entry offset of each sub-block, 0xFFFF = jumped over
    uint32_t scratch[kDecThreads / 32];
    uint32_t bcast[4];
    unsigned long long saddr_box;    // see the chain hop
    uint32_t n_long;                 // long literals of the current window
    uint32_t long_list[kMaxLong];
    uint32_t n_mid;                  // pieces (<= kMidPiece bytes) of the literals of kThreadElem+1 .. kLongLiteral-1 bytes
    uint32_t mid_next;               // next piece to hand out (warps take pieces as they become free)
    uint16_t mid_list[kMaxMid];      // element | piece << 11
    int fail;        // preamble / parse stage
    int fail_desc;   // descriptor stage (separate word: it is written while slow threads may still read `fail`)
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

// Next chain position after a literal whose length sits in 1..4 bytes behind the tag (m = 60..63); `v` = those bytes.
// Rare, and not inlined into the 16 unrolled steps of the exit-table sweep.
__device__ __noinline__ uint64_t long_literal_next(uint32_t o, uint32_t m, uint32_t v)
{
    const uint32_t extra = m - 59;
    const uint32_t mm = extra == 4 ? v : (v & ((1u << (8 * extra)) - 1u));
    return mm == 0xFFFFFFFFu ? ~0ull : (uint64_t)o + 1 + extra + (uint64_t)mm + 1;
}

// dst/src any alignment.  `lane` of `n_lanes` cooperating threads; 4 bytes per thread per step once the
// destination is word-aligned, the source word assembled from two aligned words when it is not.
template <int N_LANES>
__device__ __noinline__ void lanes_copy(uint8_t *dst, const uint8_t *src, uint32_t len, uint32_t lane)
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
// (Not inlined, like lanes_copy: inlined at every call site the byte movers were a quarter of a 148 KB kernel whose
// instruction-cache hit rate was 84 %.  Halving the code -- this and the re-rolled exit-table sweep -- turned out not to
// change the kernel's speed, measured; it is kept for the smaller binary and because the kernel no longer spills.)
// Up to kSmallElem (64) bytes by one thread.  Word path: every load is issued before the first store, so the
// loads overlap instead of each waiting behind the store before it (the compiler must assume they alias).
constexpr uint32_t kSmallElem = 64;
constexpr uint32_t kThreadElem = 128;  // literals up to this long are also moved by one thread (their source is shared memory)
constexpr uint32_t kStageWords = 8;  // words held in registers at a time (two passes cover 64 bytes)
__device__ __noinline__ void small_copy(uint8_t *d, const uint8_t *s, uint32_t len)
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

__device__ __forceinline__ void group_copy(uint8_t *dst, const uint8_t *src, uint32_t len, uint32_t glane) { lanes_copy<8>(dst, src, len, glane); }
__device__ __forceinline__ void cta_copy(uint8_t *dst, const uint8_t *src, uint32_t len, uint32_t t) { lanes_copy<kDecThreads>(dst, src, len, t); }

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
__device__ unsigned long long g_decode_counts[8];  // windows, elements, execute rounds, pending-after-round-1, flatten changes
#define COUNT_ADD(i, v) do { if (t == 0) atomicAdd(&g_decode_counts[i], (unsigned long long)(v)); } while (0)
#define PHASE_MARK(i) do { if (t == 0) { long long now_ = clock64(); atomicAdd(&g_decode_phase_cycles[i], (unsigned long long)(now_ - phase_t0_)); phase_t0_ = now_; } } while (0)
#define PHASE_INIT long long phase_t0_ = clock64()
#else
#define PHASE_MARK(i) do { } while (0)
#define PHASE_INIT do { } while (0)
#define COUNT_ADD(i, v) do { } while (0)
#endif

__global__ void __launch_bounds__(kDecThreads, 3) snappy_decode_chunks_kernel(ChunkJob *jobs, int njobs)
{
    HAP_DYN_SMEM(smem_raw);
    DecodeSmem &S = *reinterpret_cast<DecodeSmem *>(smem_raw);
    const int t = threadIdx.x;
    const uint32_t wrp = t >> 5;
    if ((int)blockIdx.x >= njobs) return;
    PHASE_INIT;
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
        S.bcast[0] = i;
    }
    __syncthreads();
    if (S.fail) {
        if (t == 0) job.status = HapResult_Bad_Frame;
        return;
    }
    uint32_t wb = S.bcast[0];  // window base: a true element start
    uint32_t d0 = 0;           // output bytes produced by earlier windows
    __syncthreads();

    // Sub-blocks covered per window.  Dense streams (a few bytes per element, e.g. Google Snappy on DXT5) fill the
    // descriptor arrays long before 256 sub-blocks are used; the span then shrinks so that no exit table is computed
    // for bytes this window never reaches, and grows back when windows stop being cut short.
    uint32_t span = kDecThreads;
    while (wb < in_end) {
        // ---- stage the window.  Chunks are byte-packed in a frame, so the chunk is rarely aligned: read aligned
        //      16-byte words and shift them so that S.cin[0] is the byte at `wb` (word loads stay aligned later) ---
        if (t == 0) { S.n_long = 0; S.n_mid = 0; S.mid_next = 0; }
        uint32_t staged_end;  // input position up to which S.cin holds this window's bytes
        {
            const uintptr_t gaddr = (uintptr_t)(src + wb);
            const uint32_t shift = (uint32_t)(gaddr & 15);           // uniform over the CTA
            const uint32_t want = span * kDecSub + 16;
            const uint32_t avail = in_end - wb < want ? in_end - wb : want;
            staged_end = wb + avail;
            const uint32_t n16 = (avail + 15) >> 4;
            const uint4 *g4 = reinterpret_cast<const uint4 *>(gaddr - shift);
            uint4 *s4 = reinterpret_cast<uint4 *>(S.cin);
            const uint32_t ws = shift >> 2, bs = (shift & 3) * 8;
            // aligned words are only read whole when every byte of them belongs to the chunk; the (at most two) words
            // that stick out at the chunk's ends are gathered bytewise, so nothing outside [src, src + in_end) is touched
            const uintptr_t c_lo = (uintptr_t)src, c_hi = (uintptr_t)src + in_end;
            for (uint32_t i = t; i < n16; i += kDecThreads) {
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
            // next window's lines on their way into L2 while this one is parsed and executed
            const uint64_t pf = (uint64_t)wb + kDecWin + (uint64_t)t * 128;
            if (pf < in_end) hap_prefetch_l2(src + pf);
        }
        const uint8_t *cinp = S.cin;
        S.entry[t] = 0xFFFFu;
        __syncthreads();

        PHASE_MARK(0);
        // ---- 1. parse.  (a) every thread: for EACH of the 64 offsets of its sub-block, where does an element
        //      chain entering there leave the sub-block?  One backward sweep: x[o] = x[o + length(o)].
        //      (b) one thread hops sub-block to sub-block along the true chain using that table.
        //      (c) every sub-block the chain enters is walked once from its true entry. ------------------
        // threads beyond the span own nothing in this window
        const uint32_t blk_start = (uint32_t)t < span && (uint64_t)wb + (uint64_t)t * kDecSub < in_end ? wb + t * kDecSub : in_end;
        const uint32_t blk_end = (uint32_t)t < span && (uint64_t)wb + (uint64_t)(t + 1) * kDecSub < in_end ? wb + (t + 1) * kDecSub : in_end;
        if (blk_start < in_end) {
            // the sub-block's 64 bytes (+ 4 bytes of header look-ahead) live in registers: the sweep is fully
            // unrolled, so every tag byte is a compile-time extraction and only the table access touches memory
            const uint32_t blk_len = blk_end - blk_start;
            const uint32_t limit = in_end - blk_start;  // a chain position may not pass this
            // The sweep runs over four groups of 16 offsets, highest first; a group's 16 bytes (+ 5 bytes of header
            // look-ahead) are re-loaded from the staged window into registers, inside the group every tag byte is a
            // compile-time extraction.  (Fully unrolled over all 64 offsets this was a third of the kernel's code.)
            const uint32_t *c32 = reinterpret_cast<const uint32_t *>(S.cin + (size_t)t * kDecSub);
#pragma unroll 1
            for (int g = kDecSub / 16 - 1; g >= 0; g--) {
                uint32_t w[6];
#pragma unroll
                for (int k = 0; k < 6; k++) w[k] = c32[4 * g + k];
#pragma unroll
                for (int oo = 15; oo >= 0; oo--) {
                    const uint32_t o = (uint32_t)(16 * g + oo);
                    const uint32_t tag = (w[oo >> 2] >> (8 * (oo & 3))) & 0xFFu;
                    const uint32_t kind = tag & 3u;
                    uint64_t nxt;
                    if (kind != 0) {
                        nxt = o + ((0x5320u >> (4 * kind)) & 0xFu);  // copy headers: 2, 3 or 5 bytes
                    } else {
                        const uint32_t m = tag >> 2;
                        if (m < 60) {
                            nxt = o + m + 2;                           // tag + (m+1) literal bytes
                        } else {
                            // the 4 bytes after the tag, assembled from registers
                            const uint32_t lo_w = w[(oo + 1) >> 2], hi_w = w[((oo + 1) >> 2) + 1];
                            nxt = long_literal_next(o, m, __funnelshift_r(lo_w, hi_w, 8 * ((oo + 1) & 3)));
                        }
                    }
                    uint32_t x;
                    if (nxt > limit) x = kExitInvalid;               // header or payload runs past the input
                    else if (nxt < blk_len) x = S.tbl[(uint32_t)nxt * kDecThreads + t];
                    else x = nxt - blk_len <= kExitMaxRel ? (uint32_t)(nxt - blk_len) : kExitFarBase + o;
                    S.tbl[o * kDecThreads + t] = (uint8_t)x;
                }
            }
        }
        __syncthreads();
        PHASE_MARK(5);
        if (t == 0) {
            // window-relative positions: the dependent chain per hop is one table load plus a few ALU ops
            const uint32_t rel_end = in_end - wb;
            uint32_t rel = 0;
            // The table's shared-window address, read back through a volatile word: ptxas otherwise re-derives it from
            // the CTA-in-cluster id (an S2R, tens of cycles on this thread's critical path) in EVERY iteration of the hop.
            volatile hap_saddr_t *base_box = reinterpret_cast<volatile hap_saddr_t *>(&S.saddr_box);
            *base_box = hap_smem_addr(S.tbl);
            const hap_saddr_t tbl_s = *base_box, entry_s = tbl_s + (hap_saddr_t)((const uint8_t *)S.entry - (const uint8_t *)S.tbl);
            while (rel < rel_end && rel < span * kDecSub) {
                const uint32_t blk = rel >> 6, o = rel & 63;
                const uint32_t x = hap_lds_u8(tbl_s + (o * kDecThreads + blk));
                hap_sts_u16(entry_s + 2 * blk, o);
                uint32_t bend = (blk + 1) << 6;
                bend = bend < rel_end ? bend : rel_end;
                if (x <= kExitMaxRel) {
                    rel = bend + x;
                } else if (x == kExitInvalid) {
                    S.fail = 1;
                    break;
                } else {
                    // a long literal leaves this sub-block by more than a byte can hold: the table names its header
                    const uint32_t p2 = wb + (blk << 6) + (x - kExitFarBase);
                    uint32_t len, aux, hdr, kind;
                    if (!read_element_header(cinp, wb, p2, in_end, len, aux, hdr, kind) || kind != 0) { S.fail = 1; break; }
                    rel = p2 + hdr + len - wb;
                }
            }
        }
        __syncthreads();
        PHASE_MARK(6);
        uint32_t entry = blk_end;
        WalkResult w;
        w.exit = 0; w.count = 0; w.out_bytes = 0; w.invalid = 0;
        if (S.entry[t] != 0xFFFFu && blk_start < in_end) {
            entry = blk_start + S.entry[t];
            w = walk_subblock(cinp, wb, entry, blk_end, in_end);
        }

        PHASE_MARK(1);
        // ---- 2. scans: element slots and output offsets; window truncation -------------------
        uint32_t total_e, total_o;
        uint32_t ebase = block_excl_sum<kDecThreads>(w.count, &total_e, S.scratch);
        const uint32_t total_e_all = total_e;
        const bool keep = ebase + w.count <= (uint32_t)kDecMaxElems;
        uint32_t kept_cnt = keep ? w.count : 0;
        uint32_t kept_out = keep ? w.out_bytes : 0;
        uint32_t obase = block_excl_sum<kDecThreads>(kept_out, &total_o, S.scratch);
        if (total_e > (uint32_t)kDecMaxElems) block_excl_sum<kDecThreads>(kept_cnt, &total_e, S.scratch);
        uint32_t next_wb;
        block_excl_max<kDecThreads>(keep ? w.exit : 0, &next_wb, S.scratch);
        // a kept sub-block whose chain is invalid poisons the stream (it is the true chain now)
        if (keep && w.invalid) S.fail = 1;
        if (t == 0 && (uint64_t)d0 + total_o > expected) S.fail = 1;
        __syncthreads();
        if (S.fail) break;
        {
            const uint32_t used_sub = (next_wb - wb + kDecSub - 1) / kDecSub;   // uniform: both come from block-wide scans
            if (total_e_all > (uint32_t)kDecMaxElems) span = used_sub + used_sub / 4 < 16u ? 16u : (used_sub + used_sub / 4 > (uint32_t)kDecThreads ? (uint32_t)kDecThreads : used_sub + used_sub / 4);
            else if (total_e_all < (uint32_t)kDecMaxElems / 2) span = span * 2 > (uint32_t)kDecThreads ? (uint32_t)kDecThreads : span * 2;
        }

        PHASE_MARK(2);
        // ---- descriptors.  e_a packs the SOURCE of an element as (kind << 30) | position:
        //      kSrcIn  : bytes of the compressed input at `position` (literals, and copies flattened onto them)
        //      kSrcOut : bytes of the output at `position` (plain copies; offset >= length)
        //      kSrcRun : periodic fill with period `position` (= the offset) of the e_b[e] - offset .. e_b[e] bytes
        if (keep && entry < blk_end) {
            uint32_t pos = entry, e = ebase, o = d0 + obase;
            while (pos < blk_end) {
                uint32_t len, aux, hdr, kind;
                read_element_header(cinp, wb, pos, in_end, len, aux, hdr, kind);
                S.e_dst[e] = o;
                S.e_len[e] = len;
                S.e_done[e] = 0;
                if (kind == 0) {
                    S.e_a[e] = kSrcIn | aux;
                    S.e_b[e] = 0;  // literals break same-offset runs (a copy's offset is never 0)
                    if (len >= kLongLiteral) {
                        uint32_t q = atomicAdd(&S.n_long, 1u);
                        if (q < (uint32_t)kMaxLong) S.long_list[q] = e;
                    } else if (len > kThreadElem) {
                        const uint32_t np = (len + kMidPiece - 1) / kMidPiece;
                        const uint32_t q = atomicAdd(&S.n_mid, np);
                        for (uint32_t pc = 0; pc < np; pc++)
                            if (q + pc < (uint32_t)kMaxMid) S.mid_list[q + pc] = (uint16_t)(e | (pc << 11));
                    }
                    pos += hdr + len;
                } else {
                    if (aux == 0 || aux > o) S.fail_desc = 1;  // offset 0 or before the start of the output
                    S.e_a[e] = aux >= len ? (kSrcOut | (o - aux)) : (kSrcRun | aux);
                    S.e_b[e] = aux;  // the offset, for run detection below; becomes the run base afterwards
                    pos += hdr;
                }
                o += len;
                e++;
            }
        }
        __syncthreads();
        if (S.fail_desc) break;

        // ---- same-offset runs: a copy with the offset of the copy right before it continues that copy's
        //      match, so it is a periodic fill of the run head's base period, independent of its neighbours --
        {
            const uint32_t strip = (total_e + kDecThreads - 1) / kDecThreads;
            const uint32_t lo = t * strip < total_e ? t * strip : total_e;
            const uint32_t hi = lo + strip < total_e ? lo + strip : total_e;
            // Pass 1: flag continuations in the spare top bit of e_len (copies are at most 64 long) and find
            // the last run head of the strip.
            uint32_t last_head = 0;  // index + 1
            for (uint32_t e = lo; e < hi; e++) {
                const uint32_t off = S.e_b[e];
                const bool cont = e > 0 && off != 0 && off == S.e_b[e - 1];
                if (cont) S.e_len[e] |= 0x80000000u;
                else last_head = e + 1;
            }
            uint32_t unused;
            uint32_t head = block_excl_max<kDecThreads>(last_head, &unused, S.scratch);  // ends with a barrier
            // Pass 2: e_b becomes the base (destination of the run head); continuations become periodic fills.
            for (uint32_t e = lo; e < hi; e++) {
                const uint32_t off = S.e_b[e];
                if (S.e_len[e] & 0x80000000u) {
                    S.e_len[e] &= 0x7FFFFFFFu;
                    S.e_a[e] = kSrcRun | off;
                    S.e_b[e] = S.e_dst[head - 1];
                } else {
                    head = e + 1;
                    S.e_b[e] = S.e_dst[e];
                }
            }
        }
        __syncthreads();

        PHASE_MARK(3);
        // ---- flatten copy-of-copy chains.  DXT payloads are full of "same as the previous block except a few
        //      bytes": a copy whose source is itself a copy, hundreds deep.  A plain copy whose source bytes lie
        //      inside ONE earlier element of this window takes over that element's source (pointer jumping on
        //      the packed e_a words; a racing update only makes the hop longer, never wrong).  Chains end at
        //      literals (-> read the input instead) or at earlier windows (-> already written). --------------
#pragma unroll 1
        for (int fr = 0; fr < kFlattenRounds; fr++) {
            for (uint32_t e = t; e < total_e; e += kDecThreads) {
                uint32_t a = S.e_a[e];
                if ((a & kSrcMask) != kSrcOut) continue;
                const uint32_t len = S.e_len[e];
                bool changed = false;
#pragma unroll 1
                for (int hop = 0; hop < kFlattenHops; hop++) {
                    const uint32_t sp = a & kPosMask;
                    if (sp + len <= d0) break;                 // reads finished output of earlier windows
                    if (sp < d0) break;                        // straddles the window start: leave it
                    uint32_t lo2 = 0, hi2 = e;                 // last element with e_dst <= sp (it is before e)
                    while (hi2 - lo2 > 1) {
                        const uint32_t m = (lo2 + hi2) >> 1;
                        if (S.e_dst[m] <= sp) lo2 = m; else hi2 = m;
                    }
                    const uint32_t f = lo2, fd = S.e_dst[f];
                    if (sp + len > fd + S.e_len[f]) break;     // spans several producers
                    const uint32_t fa = S.e_a[f];
                    if ((fa & kSrcMask) == kSrcRun) break;     // periodic producer: stay dependent on it
                    a = (fa & kSrcMask) | ((fa & kPosMask) + (sp - fd));
                    changed = true;
                    if ((fa & kSrcMask) == kSrcIn) break;      // landed on input bytes: fully resolved
                }
                if (changed) S.e_a[e] = a;
            }
            __syncthreads();
        }

        PHASE_MARK(7);
        // ---- 3. execute: round 1 = everything whose source is the input or earlier windows; later rounds =
        //         copies whose producers finished in an earlier round.  Elements of at most 64 bytes (every copy,
        //         most literals) are moved by ONE THREAD each, staged through registers so that all its loads are
        //         in flight together; longer literals by a warp each; the longest by the whole CTA.
        for (uint32_t round = 1;; round++) {
            int pending = 0;
            for (uint32_t e = t; e < total_e; e += kDecThreads) {
                if (S.e_done[e] != 0) continue;
                const uint32_t len = S.e_len[e];
                if (len > kThreadElem) continue;                 // (copies are at most 64 bytes)
                const uint32_t a = S.e_a[e], o = S.e_dst[e];
                const uint32_t kind = a & kSrcMask, ap = a & kPosMask;
                uint8_t *d = dst + o;
                if (kind == kSrcIn) {
                    const uint8_t *sl = (ap >= wb && (uint64_t)ap + len <= staged_end) ? cinp + (ap - wb) : src + ap;
                    small_copy(d, sl, len);
                    S.e_done[e] = (uint16_t)round;
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
                            if (S.e_dst[m] <= x) lo2 = m; else hi2 = m;
                        }
                        for (uint32_t f = lo2; x < x_end; f++) {
                            const uint32_t fd = S.e_dst[f], fl = S.e_len[f];
                            const uint32_t x1 = x_end < fd + fl ? x_end : fd + fl;
                            const uint32_t fa = S.e_a[f], fk = fa & kSrcMask, fp = (fa & kPosMask) + (x - fd);
                            const uint32_t df = S.e_done[f];
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
                    S.e_done[e] = (uint16_t)round;
                    continue;
                }
                const uint32_t base = S.e_b[e];
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
                        if (S.e_dst[m] <= x) lo2 = m; else hi2 = m;
                    }
                    for (uint32_t f = lo2; f < e && S.e_dst[f] < need_hi; f++) {
                        uint32_t df = S.e_done[f];
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
                            for (uint32_t k = 0; k < kStageWords; k++)
                                if (b + k < nw) { v[k] = p32[idx]; idx = idx + 1 == pw ? 0 : idx + 1; }
#pragma unroll
                            for (uint32_t k = 0; k < kStageWords; k++)
                                if (b + k < nw) d32[b + k] = v[k];
                        }
                    } else {
                        uint32_t idx = rel % off;
                        for (uint32_t i = 0; i < len; i++) { d[i] = period[idx]; idx = idx + 1 == off ? 0 : idx + 1; }
                    }
                }
                S.e_done[e] = (uint16_t)round;
            }
            if (round == 1) {
                // literals of kThreadElem+1 .. kLongLiteral-1 bytes: one warp each, from the list the descriptor pass made
                const uint32_t nmid = S.n_mid;
                if (nmid <= (uint32_t)kMaxMid) {
                    // pieces are handed out one at a time, so a warp that drew short ones simply draws more
                    for (;;) {
                        uint32_t q = 0;
                        if ((t & 31) == 0) q = atomicAdd(&S.mid_next, 1u);
                        q = __shfl_sync(HAP_FULL_MASK, q, 0);
                        if (q >= nmid) break;
                        const uint32_t item = S.mid_list[q];
                        const uint32_t e = item & 2047u, off = (item >> 11) * kMidPiece;
                        const uint32_t len = S.e_len[e];
                        const uint32_t n = len - off < kMidPiece ? len - off : kMidPiece;
                        const uint32_t ap = S.e_a[e] & kPosMask;  // only literals are this long
                        const uint8_t *sl = (ap >= wb && (uint64_t)ap + len <= staged_end) ? cinp + (ap - wb) : src + ap;
                        lanes_copy<32>(dst + S.e_dst[e] + off, sl + off, n, t & 31);
                        if ((t & 31) == 0) S.e_done[e] = 1;   // honoured from round 2 on, when every piece is in place
                    }
                } else {
                    // more of them than the list holds (cannot happen with 16 KiB of input per window, kept for safety)
                    for (uint32_t e = wrp; e < total_e; e += kDecThreads / 32) {
                        const uint32_t len = S.e_len[e];
                        if (len <= kThreadElem || len >= kLongLiteral) continue;
                        const uint32_t ap = S.e_a[e] & kPosMask;
                        const uint8_t *sl = (ap >= wb && (uint64_t)ap + len <= staged_end) ? cinp + (ap - wb) : src + ap;
                        lanes_copy<32>(dst + S.e_dst[e], sl, len, t & 31);
                        if ((t & 31) == 0) S.e_done[e] = 1;
                    }
                }
            }
            if (round == 1) {
                // long literals: the whole CTA moves each one
                const uint32_t nlong = S.n_long < (uint32_t)kMaxLong ? S.n_long : (uint32_t)kMaxLong;
                for (uint32_t q = 0; q < nlong; q++) {
                    const uint32_t e = S.long_list[q];
                    cta_copy(dst + S.e_dst[e], src + (S.e_a[e] & kPosMask), S.e_len[e], t);
                    if (t == 0) S.e_done[e] = 1;
                }
                if (S.n_long > (uint32_t)kMaxLong) {
                    // overflow of the list (pathological): sweep the descriptors instead
                    for (uint32_t e = 0; e < total_e; e++)
                        if ((S.e_a[e] & kSrcMask) == kSrcIn && S.e_len[e] >= kLongLiteral && S.e_done[e] == 0) {
                            cta_copy(dst + S.e_dst[e], src + (S.e_a[e] & kPosMask), S.e_len[e], t);
                            __syncthreads();
                            if (t == 0) S.e_done[e] = 1;
                        }
                }
            }
            COUNT_ADD(2, 1);
            if (!__syncthreads_or(pending)) break;
        }
        COUNT_ADD(0, 1);
        COUNT_ADD(1, total_e);
        PHASE_MARK(4);

        d0 += total_o;
        wb = next_wb;
        __syncthreads();
    }

    __syncthreads();
    if (t == 0) job.status = (S.fail || S.fail_desc || d0 != expected) ? HapResult_Bad_Frame : HapResult_No_Error;
}

}  // namespace hapb200
