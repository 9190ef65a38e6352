// hap_b200/csrc/snappy_decode_sparse.cuh -- K7s: per-chunk decompressor for SPARSE element streams.
//
// Same job as snappy_decode.cuh (hap_decode_chunk + snappy_uncompress, /root/reference/source/hap.c:606-642),
// tried first on every chunk.  What it exploits: the streams this library's own encoder writes (and any stream of
// DXT data that does not compress much) are SPARSE -- tens of bytes per element, most compressed bytes inside
// long literals.  K7's exit tables price every compressed byte (64 table entries per 64-byte sub-block, a fifth
// of that kernel's time) although, measured on 4K Hap Q frames, only ~40 of a window's 256 sub-blocks are ever
// entered; and its serial chain hop keeps 255 threads waiting for another fifth.  Here ONE thread simply walks the
// element chain (~350 elements per 16 KiB window, a few thousand cycles) and writes the descriptors as it goes,
// while the other seven warps move the bytes of the PREVIOUS window:
//
//   warp 0 (walker), window k+1: stage 16 KiB of input in shared memory (32 lanes), then lane 0 decodes element
//       after element from it: header -> descriptor (destination, length, packed source, run base), same-offset
//       runs marked on the fly.
//   warps 1-7 (movers), window k: flatten copy-of-copy chains, then execute in dependency rounds -- the code of K7.
//
// The groups hand windows over through two buffers and four named barriers (full / empty per buffer; the producer
// ARRIVES and moves on, the consumer SYNCS).  A serial walk would be hopeless on a DENSE stream (Google's encoder on
// DXT5: ~4 bytes per element, thousands of elements per window), so the walker watches the density: when a
// window's descriptors fill up within kWalkDenseBytes of input it stops, records where (input position, output
// bytes so far) and marks the chunk kChunkNeedsTables; the host launches K7 right after this kernel and K7
// resumes exactly those chunks from that point.  Errors are decided here the same way K7 decides them.
#pragma once
#include "snappy_decode.cuh"

namespace hapb200 {

constexpr int kWalkParse = 32;                        // the walker warp
constexpr int kWalkExec = kDecThreads - kWalkParse;   // the movers
constexpr int kWalkExecWarps = kWalkExec / 32;
constexpr int kWalkMaxElems = 1024;                   // descriptors per window
constexpr uint32_t kWalkDenseBytes = 6144;            // descriptors full within this many input bytes: dense stream
constexpr int kWalkBarExec = 2, kWalkBarFull = 3 /* +buffer */, kWalkBarEmpty = 5 /* +buffer */;   // 0 is __syncthreads
constexpr uint32_t kWinData = 0, kWinEnd = 1, kWinFail = 2, kWinDense = 3;

// One of the two hand-over buffers: the staged input of a window and its element descriptors.
struct alignas(16) WalkWindow {
    uint32_t e_dst[kWalkMaxElems];   // output offset inside the chunk
    uint32_t e_len[kWalkMaxElems];
    uint32_t e_a[kWalkMaxElems];     // packed source: kSrcIn|input position, kSrcOut|output position, kSrcRun|offset
    uint32_t e_b[kWalkMaxElems];     // destination of the head of the element's same-offset run (its own, if alone)
    uint16_t e_done[kWalkMaxElems];  // 0 = pending, r = finished in round r
    alignas(16) uint8_t cin[kDecWin + 64];   // staged window: aligned image of the input (+ alignment shift + header slack)
    uint32_t long_list[kMaxLong];    // literals of kLongLiteral bytes and more (the whole mover group copies each)
    uint16_t mid_list[kMaxMid];      // literals of kThreadElem+1 .. kLongLiteral-1 bytes (one warp each)
    uint32_t n_long, n_mid;
    uint32_t total_e;                // descriptors in use
    uint32_t d0;                     // output bytes produced by earlier windows
    uint32_t wb;                     // input position of the window's first element
    uint32_t staged_end;             // input position up to which cin holds this window's bytes
    uint32_t status;                 // kWinData | kWinEnd | kWinFail | kWinDense
    uint32_t pad[5];
};
static_assert(sizeof(WalkWindow) % 16 == 0, "both buffers keep cin 16-byte aligned");

struct WalkSmem {
    WalkWindow win[2];
    int fail;
    uint32_t first_wb;
};

#ifdef HAPB200_DECODE_PHASE_CYCLES
#define WALK_COUNT(i, v) do { if (lead) atomicAdd(&g_decode_counts[i], (unsigned long long)(v)); } while (0)
#define WALK_MARK(i) do { if (lead) { long long now_ = clock64(); atomicAdd(&g_decode_phase_cycles[i], (unsigned long long)(now_ - phase_t0_)); phase_t0_ = now_; } } while (0)
#define WALK_WAIT(i) do { if (lead) { long long now_ = clock64(); atomicAdd(&g_decode_counts[i], (unsigned long long)(now_ - phase_t0_)); phase_t0_ = now_; } } while (0)
#define WALK_INIT long long phase_t0_ = clock64()
#else
#define WALK_COUNT(i, v) do { } while (0)
#define WALK_MARK(i) do { } while (0)
#define WALK_WAIT(i) do { } while (0)
#define WALK_INIT do { } while (0)
#endif

// ---- WALKER warp ------------------------------------------------------------------------------------------------
// Produces windows into S.win[0], S.win[1], S.win[0], ... and ends the sequence with a window whose status is
// kWinEnd, kWinFail or kWinDense.  Returns that status; *wb_out / *d0_out = input position and output bytes at the
// start of the terminal window (for kWinEnd: the end of the stream and everything it decodes to).
__device__ __forceinline__ uint32_t walk_parse_warp(WalkSmem &S, const uint32_t lane, const uint8_t *__restrict__ src,
                                                    const uint32_t in_end, const uint32_t expected, uint32_t wb,
                                                    uint32_t *wb_out, uint32_t *d0_out)
{
    const bool lead = lane == 0;
    WALK_INIT;
    uint32_t d0 = 0;           // output bytes described by earlier windows
    uint32_t k = 0, verdict = kWinEnd;
    for (;; k++) {
        WalkWindow &Wn = S.win[k & 1];
        if (k >= 2) hap_bar_sync(kWalkBarEmpty + (int)(k & 1), kDecThreads);   // the movers are done with this buffer
        WALK_WAIT(5);
        if (wb >= in_end) {
            if (lead) Wn.status = kWinEnd;
            __threadfence_block();
            hap_bar_arrive(kWalkBarFull + (int)(k & 1), kDecThreads);
            verdict = kWinEnd;
            break;
        }
        // ---- stage the window.  Chunks are byte-packed in a frame, so the chunk is rarely aligned: read aligned
        //      16-byte words and shift them so that cin[0] is the byte at `wb` (word loads stay aligned later) ---
        uint32_t staged_end;
        {
            const uintptr_t gaddr = (uintptr_t)(src + wb);
            const uint32_t shift = (uint32_t)(gaddr & 15);           // uniform over the warp
            const uint32_t want = (uint32_t)kDecWin + 16;
            const uint32_t avail = in_end - wb < want ? in_end - wb : want;
            staged_end = wb + avail;
            const uint32_t n16 = (avail + 15) >> 4;
            const uint4 *g4 = reinterpret_cast<const uint4 *>(gaddr - shift);
            uint4 *s4 = reinterpret_cast<uint4 *>(Wn.cin);
            const uint32_t ws = shift >> 2, bs = (shift & 3) * 8;
            // aligned words are only read whole when every byte of them belongs to the chunk; the (at most two) words
            // that stick out at the chunk's ends are gathered bytewise, so nothing outside [src, src + in_end) is touched
            const uintptr_t c_lo = (uintptr_t)src, c_hi = (uintptr_t)src + in_end;
            for (uint32_t i = lane; i < n16; i += kWalkParse) {
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
            // next window's lines on their way into L2 while this one is walked
            for (uint32_t q = lane; q < (uint32_t)kDecWin / 128; q += kWalkParse) {
                const uint64_t pf = (uint64_t)wb + kDecWin + (uint64_t)q * 128;
                if (pf < in_end) hap_prefetch_l2(src + pf);
            }
        }
        __syncwarp();
        WALK_MARK(0);

        // ---- lane 0 walks the element chain of the window and writes the descriptors.  e_a packs the SOURCE of an
        //      element as (kind << 30) | position:
        //      kSrcIn  : bytes of the compressed input at `position` (literals, and copies flattened onto them)
        //      kSrcOut : bytes of the output at `position` (plain copies; offset >= length)
        //      kSrcRun : periodic fill with period `position` (= the offset) of the e_b[e] - offset .. e_b[e] bytes
        //      A copy with the offset of the copy right before it continues that copy's match (how every encoder emits
        //      a long or overlapping match): it becomes a periodic fill of the run HEAD's base period, independent of
        //      its neighbours; e_b = destination of the run head.
        uint32_t total_e = 0, pos_end = wb, out_end = d0, st = kWinData;
        if (lane == 0) {
            const uint8_t *cin = Wn.cin;
            const uint32_t win_end = (uint64_t)wb + kDecWin < in_end ? wb + kDecWin : in_end;  // elements starting before it are ours
            uint32_t pos = wb, e = 0, o = d0, prev_off = 0, head_dst = 0, n_long = 0, n_mid = 0;
            while (pos < win_end && e < (uint32_t)kWalkMaxElems) {
                // the 8 bytes at pos, from three aligned words (cin[pos - wb .. +8) is always staged: header slack)
                const uint32_t r = pos - wb;
                const uint32_t *w32 = reinterpret_cast<const uint32_t *>(cin + (r & ~3u));
                const uint32_t x0 = w32[0], x1 = w32[1], x2 = w32[2];
                const uint32_t sh = (r & 3u) * 8u;
                const uint32_t b0 = __funnelshift_r(x0, x1, sh), b1 = __funnelshift_r(x1, x2, sh);
                const uint32_t tag = b0 & 0xFFu, kind = tag & 3u;
                uint32_t len, hdr;
                if (kind == 0) {
                    uint32_t m = tag >> 2;
                    hdr = 1;
                    if (m >= 60) {
                        const uint32_t extra = m - 59;
                        if ((uint64_t)pos + 1 + extra > in_end) { st = kWinFail; break; }
                        const uint32_t v = (b0 >> 8) | (b1 << 24);
                        m = extra == 4 ? v : (v & ((1u << (8 * extra)) - 1u));
                        hdr = 1 + extra;
                        if (m == 0xFFFFFFFFu) { st = kWinFail; break; }
                    }
                    len = m + 1;
                    if ((uint64_t)pos + hdr + len > in_end) { st = kWinFail; break; }
                    Wn.e_dst[e] = o;
                    Wn.e_len[e] = len;
                    Wn.e_a[e] = kSrcIn | (pos + hdr);
                    Wn.e_b[e] = o;
                    Wn.e_done[e] = 0;
                    if (len >= kLongLiteral) {
                        if (n_long < (uint32_t)kMaxLong) Wn.long_list[n_long] = e;
                        n_long++;
                    } else if (len > kThreadElem) {
                        if (n_mid < (uint32_t)kMaxMid) Wn.mid_list[n_mid] = (uint16_t)e;
                        n_mid++;
                    }
                    prev_off = 0;  // literals break same-offset runs (a copy's offset is never 0)
                    pos += hdr + len;
                } else {
                    uint32_t off;
                    if (kind == 1) { hdr = 2; len = 4 + ((tag >> 2) & 7u); off = ((tag >> 5) << 8) | ((b0 >> 8) & 0xFFu); }
                    else if (kind == 2) { hdr = 3; len = 1 + (tag >> 2); off = (b0 >> 8) & 0xFFFFu; }
                    else { hdr = 5; len = 1 + (tag >> 2); off = (b0 >> 8) | (b1 << 24); }
                    if ((uint64_t)pos + hdr > in_end) { st = kWinFail; break; }
                    if (off == 0 || off > o) { st = kWinFail; break; }  // offset 0 or before the start of the output
                    Wn.e_dst[e] = o;
                    Wn.e_len[e] = len;
                    Wn.e_done[e] = 0;
                    if (off == prev_off) {
                        Wn.e_a[e] = kSrcRun | off;
                        Wn.e_b[e] = head_dst;
                    } else {
                        Wn.e_a[e] = off >= len ? (kSrcOut | (o - off)) : (kSrcRun | off);
                        Wn.e_b[e] = o;
                        head_dst = o;
                    }
                    prev_off = off;
                    pos += hdr;
                }
                o += len;
                e++;
                if (o > expected) { st = kWinFail; break; }   // (o <= 2^30 + 2^30: no wrap)
            }
            if (st == kWinData && e == (uint32_t)kWalkMaxElems && pos - wb < kWalkDenseBytes) st = kWinDense;
            total_e = e; pos_end = pos; out_end = o;
            Wn.n_long = n_long;
            Wn.n_mid = n_mid;
            Wn.total_e = e;
            Wn.d0 = d0;
            Wn.wb = wb;
            Wn.staged_end = staged_end;
            Wn.status = st;
        }
        st = __shfl_sync(HAP_FULL_MASK, st, 0);
        total_e = __shfl_sync(HAP_FULL_MASK, total_e, 0);
        pos_end = __shfl_sync(HAP_FULL_MASK, pos_end, 0);
        out_end = __shfl_sync(HAP_FULL_MASK, out_end, 0);
        WALK_MARK(1);
        __threadfence_block();
        hap_bar_arrive(kWalkBarFull + (int)(k & 1), kDecThreads);    // hand the window over; do not wait
        if (st != kWinData) { verdict = st; break; }                  // a terminal window: nothing of it is executed
        WALK_COUNT(0, 1);
        WALK_COUNT(1, total_e);
        d0 = out_end;
        wb = pos_end;
    }
    // k = index of the terminal window = number of data windows.  The movers' last "empty" arrival has no taker
    // yet: take it, which also means that every byte of the data windows has been written.
    if (k >= 1) hap_bar_sync(kWalkBarEmpty + (int)((k - 1) & 1), kDecThreads);
    *wb_out = wb;
    *d0_out = d0;
    return verdict;
}

// ---- EXECUTE group: flatten copy chains, then move the bytes ------------------------------------------------
// xt = thread index inside the group.
__device__ __forceinline__ void walk_execute_group(WalkSmem &S, const uint32_t xt, const uint8_t *__restrict__ src,
                                                     uint8_t *__restrict__ dst)
{
    const bool lead = xt == 0;
    const uint32_t xw = xt >> 5;
    WALK_INIT;
    for (uint32_t k = 0;; k++) {
        WalkWindow &Wn = S.win[k & 1];
        hap_bar_sync(kWalkBarFull + (int)(k & 1), kDecThreads);
        WALK_WAIT(6);
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
            for (uint32_t e = xt; e < total_e; e += kWalkExec) {
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
            hap_bar_sync(kWalkBarExec, kWalkExec);
        }

        WALK_MARK(7);
        // ---- execute: round 1 = everything whose source is the input or earlier windows; later rounds =
        //      copies whose producers finished in an earlier round.  Elements of at most 256 bytes (every copy,
        //      most literals) are moved by ONE THREAD each, staged through registers so that all its loads are
        //      in flight together; longer literals by a warp each; the longest by the whole group.
        for (uint32_t round = 1;; round++) {
            int pending = 0;
            for (uint32_t e = xt; e < total_e; e += kWalkExec) {
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
                    for (uint32_t q = xw; q < nmid; q += kWalkExecWarps) {
                        const uint32_t e = Wn.mid_list[q];
                        const uint32_t len = Wn.e_len[e];
                        const uint32_t ap = Wn.e_a[e] & kPosMask;  // only literals are this long
                        const uint8_t *sl = (ap >= wb && (uint64_t)ap + len <= staged_end) ? cinp + (ap - wb) : src + ap;
                        lanes_copy<32>(dst + Wn.e_dst[e], sl, len, xt & 31);
                        if ((xt & 31) == 0) Wn.e_done[e] = 1;
                    }
                } else {
                    // more of them than the list holds (cannot happen with 16 KiB of input per window, kept for safety)
                    for (uint32_t e = xw; e < total_e; e += kWalkExecWarps) {
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
                    lanes_copy<kWalkExec>(dst + Wn.e_dst[e], src + (Wn.e_a[e] & kPosMask), Wn.e_len[e], xt);
                    if (xt == 0) Wn.e_done[e] = 1;
                }
                if (Wn.n_long > (uint32_t)kMaxLong) {
                    // overflow of the list (pathological): sweep the descriptors instead
                    for (uint32_t e = 0; e < total_e; e++)
                        if ((Wn.e_a[e] & kSrcMask) == kSrcIn && Wn.e_len[e] >= kLongLiteral && Wn.e_done[e] == 0) {
                            lanes_copy<kWalkExec>(dst + Wn.e_dst[e], src + (Wn.e_a[e] & kPosMask), Wn.e_len[e], xt);
                            hap_bar_sync(kWalkBarExec, kWalkExec);
                            if (xt == 0) Wn.e_done[e] = 1;
                        }
                }
            }
            WALK_COUNT(2, 1);
            if (!hap_bar_or(kWalkBarExec, kWalkExec, pending)) break;
        }
        WALK_MARK(4);
        __threadfence_block();
        hap_bar_arrive(kWalkBarEmpty + (int)(k & 1), kDecThreads);   // the buffer may be refilled
    }
}


__global__ void __launch_bounds__(kDecThreads, 3) snappy_decode_sparse_kernel(ChunkJob *jobs, int njobs)
{
    HAP_DYN_SMEM(smem_raw);
    WalkSmem &S = *reinterpret_cast<WalkSmem *>(smem_raw);
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
        S.first_wb = i;
    }
    __syncthreads();
    if (S.fail) {
        if (t == 0) job.status = HapResult_Bad_Frame;
        return;
    }
    const uint32_t wb0 = S.first_wb;  // first window base: a true element start

    if (t < kWalkParse) {
        uint32_t wb_end = 0, d0_end = 0;
        const uint32_t verdict = walk_parse_warp(S, (uint32_t)t, src, in_end, expected, wb0, &wb_end, &d0_end);
        if (t == 0) {
            if (verdict == kWinDense) {
                // everything before (wb_end, d0_end) is decoded; K7 takes the chunk from there
                job.resume_wb = wb_end;
                job.resume_d0 = d0_end;
                job.status = kChunkNeedsTables;
            } else {
                job.status = (verdict == kWinFail || d0_end != expected) ? HapResult_Bad_Frame : HapResult_No_Error;
            }
        }
    } else {
        walk_execute_group(S, (uint32_t)(t - kWalkParse), src, dst);
    }
}

}  // namespace hapb200
