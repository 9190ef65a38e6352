// hap_b200/csrc/snappy_encode.cuh -- K5: second-stage compressor for one fragment of a Hap chunk.
//
// Replaces snappy_compress as the reference calls it from the chunk loop of hap_encode_texture
// (/root/reference/source/hap.c:448-476, call site :453).  The reference's loop is serial over chunks
// and Snappy is serial inside a chunk; here a chunk is cut into independent 32 KiB FRAGMENTS (a raw
// Snappy stream is varint(length) followed by elements, so fragment element streams concatenate
// into one legal chunk stream as long as copies never reach outside their fragment), one CTA per
// fragment, and inside the fragment every 4-byte word is handled by its own thread:
//   1. first-occurrence hash of every aligned word (shared-memory table, atomicMin => deterministic);
//   2. each word picks a source: the word one DXT block back when its whole block repeats the
//      previous block (block RLE, decodes as one periodic fill), else the first earlier occurrence
//      of the same word.  Choosing FIRST occurrences makes copies point at literal data, so the
//      decoder's dependency depth stays ~1 (see snappy_decode.cuh);
//   3. runs of words with one distance become copy elements (<= 64 bytes each), everything else
//      literal runs; element sizes are prefix-summed and every word writes its own bytes.
// The output differs from Google's encoder byte-for-byte (the reference pins no Snappy bytes,
// SURVEY.md 8c); parity is: the reference's HapDecode reproduces the input exactly.
#pragma once
#include "hap_codes.h"
#include "simt.h"

namespace hapb200 {

constexpr int kEncThreads = 1024;
constexpr int kEncWarps = kEncThreads / 32;
constexpr int kFragBytes = 32768;
constexpr int kFragWords = kFragBytes / 4;
constexpr int kFragCap = kFragBytes + 32;   // element stream of a fragment never exceeds n + 3
constexpr int kEncHashBits = 13;
constexpr uint32_t kFragStoredRaw = 0xFFFFFFFFu;  // fragment size marker: chunk must be stored raw

struct EncodeSmem {
    uint32_t data[kFragWords];
    union {
        uint32_t table[1 << kEncHashBits];          // steps 1-2
        struct {
            uint16_t dist2[kFragWords];             // step 3+: distance in words after demotion, 0 = literal
            uint16_t runend[kFragWords];            // indexed by run start: last word of the run
        } r;
    } u;
    uint16_t dist[kFragWords];                      // step 2: raw candidate; step 4+: run start of each word
    uint16_t pos[kFragWords];                       // output offset of each word's contribution
    uint8_t out[kFragCap];
    uint32_t warp_tot[kEncWarps];
    uint32_t total;
};

__device__ __forceinline__ uint32_t enc_hash(uint32_t w) { return (w * 0x9E3779B1u) >> (32 - kEncHashBits); }

__device__ __forceinline__ uint32_t literal_header_bytes(uint32_t len) { return len <= 60 ? 1u : len <= 256 ? 2u : 3u; }

// One fragment: `n` input bytes at `in` (n % 8 == 0, n <= kFragBytes) -> element stream in S.out,
// returns its size (all threads).  period_words = DXT block size in words (2 or 4).
__device__ __forceinline__ uint32_t compress_fragment(EncodeSmem &S, uint32_t W, uint32_t period_words)
{
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    constexpr int kPerWarp = kFragWords / kEncWarps;  // 256 words per warp, 8 lane-strided iterations
    constexpr int kIters = kPerWarp / 32;

    for (int i = t; i < (1 << kEncHashBits); i += kEncThreads) S.u.table[i] = 0xFFFFFFFFu;
    __syncthreads();
    // 1. first occurrence of every word
    for (uint32_t i = t; i < W; i += kEncThreads) atomicMin(&S.u.table[enc_hash(S.data[i])], i);
    __syncthreads();
    // 2. candidate distance per word
    for (uint32_t i = t; i < W; i += kEncThreads) {
        uint32_t w = S.data[i];
        uint32_t c = S.u.table[enc_hash(w)];
        uint32_t d = (c < i && S.data[c] == w) ? i - c : 0;
        uint32_t b0 = i - (i % period_words);
        if (b0 >= period_words && b0 + period_words <= W) {
            bool rep = true;
            for (uint32_t k = 0; k < period_words; k++) rep = rep && (S.data[b0 + k] == S.data[b0 + k - period_words]);
            if (rep) d = period_words;
        }
        S.dist[i] = (uint16_t)d;
    }
    __syncthreads();
    // 2b. propagation: a word that is not yet part of a run adopts its left (else right) neighbour's distance
    //     when its own data also matches there.  First occurrences of neighbouring words often point at
    //     different earlier blocks; two rounds of this re-align them and recover most of what a greedy
    //     match extension would find (0.92 -> 0.73 of the input on Hap Q picture content).
    {
        uint16_t *cur = S.dist, *nxt = S.u.r.dist2;  // the hash table is dead from here on
#pragma unroll 1
        for (int round = 0; round < 2; round++) {
            for (uint32_t i = t; i < W; i += kEncThreads) {
                const uint32_t d = cur[i];
                const uint32_t l = i > 0 ? cur[i - 1] : 0u, r = i + 1 < W ? cur[i + 1] : 0u;
                uint32_t nd = d;
                if (!(d != 0 && (d == l || d == r))) {
                    const uint32_t w = S.data[i];
                    if (l != 0 && i >= l && S.data[i - l] == w) nd = l;
                    else if (r != 0 && i >= r && S.data[i - r] == w) nd = r;
                }
                nxt[i] = (uint16_t)nd;
            }
            __syncthreads();
            uint16_t *tmp = cur; cur = nxt; nxt = tmp;
        }
    }
    // 3. demote matches that do not continue for at least two words (a 4-byte copy saves nothing)
    for (uint32_t i = t; i < W; i += kEncThreads) {
        uint32_t d = S.dist[i];
        bool keep = d != 0 && ((i > 0 && S.dist[i - 1] == d) || (i + 1 < W && S.dist[i + 1] == d));
        S.u.r.dist2[i] = keep ? (uint16_t)d : 0;
    }
    __syncthreads();
    // 4. run start of every word: inclusive max-scan of (i+1 where a run starts)
    {
        uint32_t carry = 0;
        const uint32_t wbase = warp * kPerWarp;
#pragma unroll 1
        for (int k = 0; k < kIters; k++) {
            uint32_t i = wbase + k * 32 + lane;
            uint32_t v = 0;
            if (i < W && (i == 0 || S.u.r.dist2[i] != S.u.r.dist2[i - 1])) v = i + 1;
#pragma unroll
            for (int dlt = 1; dlt < 32; dlt <<= 1) {
                uint32_t o = __shfl_up_sync(HAP_FULL_MASK, v, dlt);
                if (lane >= dlt) v = v > o ? v : o;
            }
            v = v > carry ? v : carry;
            if (i < W) S.dist[i] = (uint16_t)v;  // 0 = no start seen inside this warp's span yet
            carry = __shfl_sync(HAP_FULL_MASK, v, 31);
        }
        if (lane == 0) S.warp_tot[warp] = carry;
        __syncthreads();
        uint32_t base = 0;
        for (int w2 = 0; w2 < warp; w2++) base = base > S.warp_tot[w2] ? base : S.warp_tot[w2];
#pragma unroll 1
        for (int k = 0; k < kIters; k++) {
            uint32_t i = wbase + k * 32 + lane;
            if (i < W) {
                uint32_t v = S.dist[i];
                S.dist[i] = (uint16_t)((v ? v : base) - 1);
            }
        }
    }
    __syncthreads();
    // run ends, stored at the run's start index
    for (uint32_t i = t; i < W; i += kEncThreads)
        if (i + 1 == W || S.u.r.dist2[i + 1] != S.u.r.dist2[i]) S.u.r.runend[S.dist[i]] = (uint16_t)i;
    __syncthreads();
    // 5. bytes each word contributes, exclusive sum-scan -> S.pos
    {
        uint32_t carry = 0;
        const uint32_t wbase = warp * kPerWarp;
#pragma unroll 1
        for (int k = 0; k < kIters; k++) {
            uint32_t i = wbase + k * 32 + lane;
            uint32_t c = 0;
            if (i < W) {
                uint32_t rs = S.dist[i];
                if (S.u.r.dist2[i] == 0) {
                    c = 4;
                    if (i == rs) c += literal_header_bytes(4u * (S.u.r.runend[rs] - rs + 1));
                } else if (((i - rs) & 15) == 0) {
                    c = 3;
                }
            }
            uint32_t v = c;
#pragma unroll
            for (int dlt = 1; dlt < 32; dlt <<= 1) {
                uint32_t o = __shfl_up_sync(HAP_FULL_MASK, v, dlt);
                if (lane >= dlt) v += o;
            }
            if (i < W) S.pos[i] = (uint16_t)(carry + v - c);
            carry += __shfl_sync(HAP_FULL_MASK, v, 31);
        }
        if (lane == 0) S.warp_tot[warp] = carry;
        __syncthreads();
        uint32_t base = 0, tot = 0;
        for (int w2 = 0; w2 < kEncWarps; w2++) {
            uint32_t s = S.warp_tot[w2];
            if (w2 < warp) base += s;
            tot += s;
        }
#pragma unroll 1
        for (int k = 0; k < kIters; k++) {
            uint32_t i = wbase + k * 32 + lane;
            if (i < W) S.pos[i] = (uint16_t)(S.pos[i] + base);
        }
        if (t == 0) S.total = tot;
    }
    __syncthreads();
    // 6. every word writes its own bytes
    for (uint32_t i = t; i < W; i += kEncThreads) {
        uint32_t rs = S.dist[i], d = S.u.r.dist2[i], p = S.pos[i];
        if (d == 0) {
            if (i == rs) {
                uint32_t len = 4u * (S.u.r.runend[rs] - rs + 1);
                if (len <= 60) {
                    S.out[p++] = (uint8_t)((len - 1) << 2);
                } else if (len <= 256) {
                    S.out[p++] = (uint8_t)(60 << 2);
                    S.out[p++] = (uint8_t)(len - 1);
                } else {
                    S.out[p++] = (uint8_t)(61 << 2);
                    S.out[p++] = (uint8_t)(len - 1);
                    S.out[p++] = (uint8_t)((len - 1) >> 8);
                }
            }
            uint32_t w = S.data[i];
            S.out[p] = (uint8_t)w;
            S.out[p + 1] = (uint8_t)(w >> 8);
            S.out[p + 2] = (uint8_t)(w >> 16);
            S.out[p + 3] = (uint8_t)(w >> 24);
        } else if (((i - rs) & 15) == 0) {
            uint32_t left = S.u.r.runend[rs] - i + 1;       // words left in the run
            uint32_t len = 4u * (left < 16 ? left : 16);    // 4..64 bytes
            uint32_t off = 4u * d;
            S.out[p] = (uint8_t)(2u | ((len - 1) << 2));     // copy with 2-byte offset
            S.out[p + 1] = (uint8_t)off;
            S.out[p + 2] = (uint8_t)(off >> 8);
        }
    }
    __syncthreads();
    return S.total;
}

// Geometry of one texture section of a batch of identical frames (host fills it once per call).
struct SectionGeom {
    uint32_t bytes;            // texture bytes (hap.c: inputBufferBytes)
    uint32_t chunks;           // limited chunk count (hap.c:277-300)
    uint32_t chunk_bytes;      // bytes / chunks (hap.c:433)
    uint32_t frags_per_chunk;  // ceil(chunk_bytes / kFragBytes)
    uint32_t period_words;     // DXT block size in 4-byte words (2 or 4)
    uint32_t compress;         // 1 = HapCompressorSnappy requested and chunk_bytes % 8 == 0
    uint32_t frag_base;        // index of this section's first fragment inside a frame
    uint32_t fmt_nibble;       // wire format id (hap.c:45-51)
    uint32_t top_hdr;          // 4 or 8 (hap.c:398-405, :425-428)
    uint32_t want_snappy;      // caller passed HapCompressorSnappy
    uint64_t in_offset;        // this texture's bytes of frame f start at dxt + in_offset + f * in_stride
    uint64_t in_stride;
};

struct FrameGeom {
    uint32_t sections;         // 1 or 2
    uint32_t outer_hdr;        // 0 (single texture), 4 or 8 (hap.c:563-576)
    uint32_t frags_per_frame;
    uint32_t pad;
    SectionGeom s[2];
};

// grid.x = frames * frags_per_frame.  dxt: base pointer of the texture bytes; scratch: [grid.x][kFragCap];
// frag_size: [grid.x] (kFragStoredRaw when the fragment was not compressed).
__global__ void __launch_bounds__(kEncThreads) snappy_encode_fragments_kernel(
    const uint8_t *__restrict__ dxt, FrameGeom G, uint8_t *__restrict__ scratch, uint32_t *__restrict__ frag_size)
{
    HAP_DYN_SMEM(smem_raw);
    EncodeSmem &S = *reinterpret_cast<EncodeSmem *>(smem_raw);
    const int t = threadIdx.x;
    const uint32_t gfrag = blockIdx.x;
    const uint32_t frame = gfrag / G.frags_per_frame;
    const uint32_t f = gfrag % G.frags_per_frame;
    const SectionGeom &sec = (G.sections == 2 && f >= G.s[1].frag_base) ? G.s[1] : G.s[0];
    const uint32_t fl = f - sec.frag_base;
    const uint32_t chunk = fl / sec.frags_per_chunk, j = fl % sec.frags_per_chunk;
    if (!sec.compress) {
        if (t == 0) frag_size[gfrag] = kFragStoredRaw;
        return;
    }
    const uint64_t in_off = (uint64_t)frame * sec.in_stride + sec.in_offset + (uint64_t)chunk * sec.chunk_bytes +
                            (uint64_t)j * kFragBytes;
    const uint32_t left = sec.chunk_bytes - j * kFragBytes;
    const uint32_t n = left < (uint32_t)kFragBytes ? left : (uint32_t)kFragBytes;
    const uint32_t W = n >> 2;
    const uint8_t *in = dxt + in_off;
    if ((((uintptr_t)in) & 3) == 0) {
        const uint32_t *in32 = reinterpret_cast<const uint32_t *>(in);
        for (uint32_t i = t; i < W; i += kEncThreads) S.data[i] = in32[i];
    } else {
        for (uint32_t i = t; i < W; i += kEncThreads)
            S.data[i] = in[4 * i] | (in[4 * i + 1] << 8) | (in[4 * i + 2] << 16) | ((uint32_t)in[4 * i + 3] << 24);
    }
    __syncthreads();
    const uint32_t total = compress_fragment(S, W, sec.period_words);
    uint8_t *o = scratch + (uint64_t)gfrag * kFragCap;
    const uint32_t *o32s = reinterpret_cast<const uint32_t *>(S.out);
    uint32_t *o32 = reinterpret_cast<uint32_t *>(o);
    for (uint32_t i = t; i < (total + 3) / 4; i += kEncThreads) o32[i] = o32s[i];
    if (t == 0) frag_size[gfrag] = total;
}

}  // namespace hapb200
