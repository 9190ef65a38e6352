// hap_b200/csrc/snappy_decode.cuh -- K7: per-chunk second-stage decompressor.
//
// Replaces hap_decode_chunk + snappy_uncompress (/root/reference/source/hap.c:606-642, call sites
// :612 and :899): one chunk of a Hap frame -> its decoded DXT bytes at a fixed destination.
// A chunk is one raw Snappy stream (compressor byte 0x0B) or a verbatim copy (0x0A).
//
// One CTA per chunk.  Snappy is serial inside a stream (an element's position depends on the
// length of every element before it, and copies read earlier output), so the kernel breaks both
// chains explicitly, window by window over the compressed bytes:
//   1. PARSE, speculatively in parallel.  Thread t owns 64 compressed bytes and walks the element
//      chain from a guessed entry (its sub-block start).  The true entry of sub-block t is the
//      running maximum of the exits of the sub-blocks before it; threads whose entry moved re-walk,
//      and the loop ends at the fixpoint, which is the true chain (induction from sub-block 0).
//      Element chains re-synchronise after a few elements, so this is 2-3 rounds in practice.
//   2. SCAN element counts / output bytes -> every element's destination offset.
//   3. EXECUTE.  Literals are independent (source = compressed bytes).  Runs of adjacent copies
//      with one offset (how every encoder emits a long or overlapping match) become independent
//      periodic fills of the run's base period.  Remaining copies go in dependency rounds: a copy
//      runs once every element overlapping its source range finished in an earlier round.
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
constexpr int kDecMaxElems = 3072;                 // descriptors held in shared memory per window
constexpr uint32_t kLiteralMark = 0xFFFFFFFFu;

struct DecodeSmem {
    uint32_t e_dst[kDecMaxElems];    // output offset inside the chunk
    uint32_t e_len[kDecMaxElems];
    uint32_t e_src[kDecMaxElems];    // literal: payload position in the chunk input; copy: offset
    uint32_t e_base[kDecMaxElems];   // copy: destination of the head of its same-offset run; literal: mark
    uint16_t e_done[kDecMaxElems];   // 0 = pending, r = finished in round r (literals: 1)
    uint8_t cin[kDecWin + 16];       // staged window of compressed bytes (+ header slack)
    uint32_t scratch[kDecThreads / 32];
    uint32_t bcast[4];
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

__global__ void __launch_bounds__(kDecThreads) snappy_decode_chunks_kernel(ChunkJob *jobs, int njobs)
{
    HAP_DYN_SMEM(smem_raw);
    DecodeSmem &S = *reinterpret_cast<DecodeSmem *>(smem_raw);
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    constexpr int kWarps = kDecThreads / 32;
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
    if (job.compressor != kHapChunkSnappy) {
        if (t == 0) job.status = HapResult_Bad_Frame;  // hap.c:637-640
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

    while (wb < in_end) {
        // ---- stage the window ---------------------------------------------------------------
        for (uint32_t i = t; i < kDecWin + 16; i += kDecThreads) {
            uint64_t g = (uint64_t)wb + i;
            S.cin[i] = g < in_end ? src[g] : 0;
        }
        __syncthreads();

        // ---- 1. speculative parallel parse to the fixpoint -----------------------------------
        const uint32_t blk_start = (uint64_t)wb + (uint64_t)t * kDecSub < in_end ? wb + t * kDecSub : in_end;
        const uint32_t blk_end = (uint64_t)wb + (uint64_t)(t + 1) * kDecSub < in_end ? wb + (t + 1) * kDecSub : in_end;
        uint32_t entry = blk_start;
        WalkResult w;
        w.exit = 0; w.count = 0; w.out_bytes = 0; w.invalid = 0;
        bool need = true;
        for (;;) {
            if (need) {
                if (entry < blk_end) {
                    w = walk_subblock(S.cin, wb, entry, blk_end, in_end);
                } else {
                    w.exit = 0; w.count = 0; w.out_bytes = 0; w.invalid = 0;  // jumped over by a long literal
                }
            }
            uint32_t all_max;
            uint32_t before = block_excl_max<kDecThreads>(w.exit, &all_max, S.scratch);
            uint32_t new_entry = before > blk_start ? before : blk_start;
            need = new_entry != entry;
            entry = new_entry;
            if (!__syncthreads_or(need ? 1 : 0)) break;
        }

        // ---- 2. scans: element slots and output offsets; window truncation -------------------
        uint32_t total_e, total_o;
        uint32_t ebase = block_excl_sum<kDecThreads>(w.count, &total_e, S.scratch);
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

        // ---- descriptors ------------------------------------------------------------------------
        if (keep && entry < blk_end) {
            uint32_t pos = entry, e = ebase, o = d0 + obase;
            while (pos < blk_end) {
                uint32_t len, aux, hdr, kind;
                read_element_header(S.cin, wb, pos, in_end, len, aux, hdr, kind);
                S.e_dst[e] = o;
                S.e_len[e] = len;
                S.e_src[e] = aux;
                if (kind == 0) {
                    S.e_base[e] = kLiteralMark;
                    S.e_done[e] = 1;
                    pos += hdr + len;
                } else {
                    if (aux == 0 || aux > o) S.fail_desc = 1;  // offset 0 or before the start of the output
                    S.e_base[e] = o;
                    S.e_done[e] = 0;
                    pos += hdr;
                }
                o += len;
                e++;
            }
        }
        __syncthreads();
        if (S.fail_desc) break;

        // ---- same-offset runs: base of a continuation copy = destination of the run head ---------
        {
            const uint32_t strip = (total_e + kDecThreads - 1) / kDecThreads;
            const uint32_t lo = t * strip < total_e ? t * strip : total_e;
            const uint32_t hi = lo + strip < total_e ? lo + strip : total_e;
            // Pass 1: flag continuation copies (same offset as the copy right before them) in the
            // spare top bit of e_len (copies are at most 64 long) and find the last run head of the strip.
            uint32_t last_head = 0;  // index + 1; literals count as heads (they end every run)
            for (uint32_t e = lo; e < hi; e++) {
                bool cont = e > 0 && S.e_base[e] != kLiteralMark && S.e_base[e - 1] != kLiteralMark &&
                            S.e_src[e] == S.e_src[e - 1];
                if (cont) S.e_len[e] |= 0x80000000u;
                else last_head = e + 1;
            }
            uint32_t unused;
            uint32_t head = block_excl_max<kDecThreads>(last_head, &unused, S.scratch);  // ends with a barrier
            // Pass 2: a continuation copy reads the base period of its run head.
            for (uint32_t e = lo; e < hi; e++) {
                if (S.e_len[e] & 0x80000000u) {
                    S.e_len[e] &= 0x7FFFFFFFu;
                    S.e_base[e] = S.e_dst[head - 1];
                } else {
                    head = e + 1;
                }
            }
        }
        __syncthreads();

        // ---- 3a. literals: independent, source is the compressed stream --------------------------
        for (uint32_t e = warp; e < total_e; e += kWarps) {
            if (S.e_base[e] != kLiteralMark) continue;
            const uint32_t len = S.e_len[e];
            const uint8_t *s = src + S.e_src[e];
            uint8_t *d = dst + S.e_dst[e];
            for (uint32_t i = lane; i < len; i += 32) d[i] = s[i];
        }
        __syncthreads();

        // ---- 3b. copies in dependency rounds -------------------------------------------------------
        for (uint32_t round = 2;; round++) {
            int pending = 0;
            for (uint32_t e = warp; e < total_e; e += kWarps) {
                if (S.e_done[e]) continue;
                const uint32_t len = S.e_len[e], off = S.e_src[e], base = S.e_base[e], o = S.e_dst[e];
                const uint32_t rel = o - base;  // position of this element inside its run
                // bytes this element reads: the run's base period, or just its own window of it
                uint32_t need_lo, need_hi;
                if (rel + len <= off) {
                    need_lo = o - off;
                    need_hi = need_lo + len;
                } else {
                    need_lo = base - off;
                    need_hi = base;
                }
                bool ready = true;
                if (need_hi > d0) {
                    uint32_t x = need_lo > d0 ? need_lo : d0;
                    // last element with e_dst <= x
                    uint32_t a = 0, b = e;  // the producer is before e
                    while (b - a > 1) {
                        uint32_t m = (a + b) >> 1;
                        if (S.e_dst[m] <= x) a = m; else b = m;
                    }
                    for (uint32_t f = a; f < e && S.e_dst[f] < need_hi; f++) {
                        uint32_t dn = S.e_done[f];
                        if (dn == 0 || dn >= round) { ready = false; break; }
                    }
                }
                if (!ready) {
                    pending = 1;
                    continue;
                }
                const uint8_t *period = dst + (base - off);
                uint8_t *d = dst + o;
                for (uint32_t i = lane; i < len; i += 32) {
                    uint32_t idx = rel + i;
                    if (idx >= off) idx %= off;
                    d[i] = period[idx];
                }
                __syncwarp();  // every lane has read e_done[e] above before it changes
                if (lane == 0) S.e_done[e] = (uint16_t)round;
            }
            if (!__syncthreads_or(pending)) break;
        }

        d0 += total_o;
        wb = next_wb;
        __syncthreads();
    }

    __syncthreads();
    if (t == 0) job.status = (S.fail || S.fail_desc || d0 != expected) ? HapResult_Bad_Frame : HapResult_No_Error;
}

}  // namespace hapb200
