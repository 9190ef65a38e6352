// hap_b200/csrc/hap_assemble.cuh -- K6: frame assembly on device.
//
// Everything hap_encode_texture / HapEncode (/root/reference/source/hap.c:355-604) decides AFTER
// compression -- chunk sizes, the per-chunk raw fallback (:460-466), the whole-texture fallback
// (:478-487), section lengths, the two tables of the Decode Instructions container (:430-442) and the
// 4/8-byte section headers (:189-212, :497-499, :598) -- is decided here from the fragment sizes K5
// produced, without a host round trip:
//   hap_plan_frames_kernel   one CTA per frame: sizes -> layout, writes headers + tables + varints and
//                            the destination of every fragment;
//   hap_place_fragments_kernel  one CTA per fragment: copies its element stream (or its raw DXT bytes
//                            on a fallback) to the final position.  Chunks end up back to back with no
//                            padding, exactly as hap.c:473 lays them out.
#pragma once
#include "block_primitives.cuh"
#include "snappy_encode.cuh"

namespace hapb200 {

constexpr int kPlanThreads = 256;
constexpr uint32_t kPlaceRawFlag = 0x80000000u;

__device__ __forceinline__ void put_section_header(uint8_t *p, uint32_t hdr, uint32_t len, uint32_t type)
{
    // hap.c:189-212
    if (hdr == 4) {
        p[0] = (uint8_t)len; p[1] = (uint8_t)(len >> 8); p[2] = (uint8_t)(len >> 16);
    } else {
        p[0] = p[1] = p[2] = 0;
        p[4] = (uint8_t)len; p[5] = (uint8_t)(len >> 8); p[6] = (uint8_t)(len >> 16); p[7] = (uint8_t)(len >> 24);
    }
    p[3] = (uint8_t)type;
}

__device__ __forceinline__ uint32_t varint_bytes(uint32_t v) { return v < (1u << 7) ? 1 : v < (1u << 14) ? 2 : v < (1u << 21) ? 3 : v < (1u << 28) ? 4 : 5; }

// Sum of the fragment sizes of one chunk; returns false when any fragment was not compressed.
__device__ __forceinline__ bool chunk_packed_size(const uint32_t *fs, uint32_t fpc, uint32_t chunk_bytes, uint32_t &packed)
{
    uint64_t sum = varint_bytes(chunk_bytes);
    bool ok = true;
    for (uint32_t j = 0; j < fpc; j++) {
        uint32_t s = fs[j];
        if (s == kFragStoredRaw) ok = false;
        sum += s;
    }
    packed = ok && sum < chunk_bytes ? (uint32_t)sum : chunk_bytes;  // hap.c:460: packed >= chunk -> raw
    return ok && sum < chunk_bytes;
}

// out: [frames][out_stride]; frag_size / frag_dst: [frames][G.frags_per_frame]; out_used: [frames].
__global__ void __launch_bounds__(kPlanThreads) hap_plan_frames_kernel(
    FrameGeom G, const uint8_t *__restrict__ dxt, const uint32_t *__restrict__ frag_size,
    uint32_t *__restrict__ frag_dst, uint8_t *__restrict__ out, uint64_t out_stride,
    unsigned long long *__restrict__ out_used)
{
    __shared__ uint32_t scratch[kPlanThreads / 32];
    __shared__ uint32_t sec_len_sh[2];
    const int t = threadIdx.x;
    const uint32_t frame = blockIdx.x;
    uint8_t *fo = out + (uint64_t)frame * out_stride;
    const uint32_t *fs_frame = frag_size + (uint64_t)frame * G.frags_per_frame;
    uint32_t *fd_frame = frag_dst + (uint64_t)frame * G.frags_per_frame;

    uint32_t sec_off = G.outer_hdr;
    for (uint32_t si = 0; si < G.sections; si++) {
        const SectionGeom &sec = G.s[si];
        const uint32_t k = sec.chunks, fpc = sec.frags_per_chunk, hdr = sec.top_hdr;
        const uint32_t di = 5u * k + 8u;  // hap.c:265-275
        const uint32_t *fs = fs_frame + sec.frag_base;
        uint32_t *fd = fd_frame + sec.frag_base;
        // pass 1: would-be complex body length (hap.c:446-476)
        bool complex_storage = false;
        uint64_t body = 0;
        if (sec.want_snappy) {
            for (uint32_t c0 = 0; c0 < k; c0 += kPlanThreads) {
                uint32_t c = c0 + t, sz = 0;
                if (c < k) chunk_packed_size(fs + (uint64_t)c * fpc, fpc, sec.chunk_bytes, sz);
                // chunk sizes can exceed 16 bits: sum the two halves separately (each fits 32 bits)
                uint32_t tot_lo, tot_hi;
                block_excl_sum<kPlanThreads>(sz & 0xFFFFu, &tot_lo, scratch);
                block_excl_sum<kPlanThreads>(sz >> 16, &tot_hi, scratch);
                body += (uint64_t)tot_lo + ((uint64_t)tot_hi << 16);
            }
            body += 4u + di;
            complex_storage = body < (uint64_t)sec.bytes + hdr;  // hap.c:478
        }
        uint8_t *so = fo + sec_off;
        uint32_t section_len;
        if (complex_storage) {
            section_len = (uint32_t)body;
            uint8_t *p = so + hdr;
            uint8_t *ctab = p + 8, *stab = p + 8 + k + 4;
            if (t == 0) {
                put_section_header(so, hdr, section_len, (kHapComplex << 4) | sec.fmt_nibble);
                put_section_header(p, 4, di, kSecDecodeInstructions);        // hap.c:436
                put_section_header(p + 4, 4, k, kSecCompressorTable);        // hap.c:438
                put_section_header(ctab + k, 4, 4u * k, kSecSizeTable);      // hap.c:440
            }
            uint32_t running = sec_off + hdr + 4u + di;  // offset of chunk 0 inside the frame
            for (uint32_t c0 = 0; c0 < k; c0 += kPlanThreads) {
                uint32_t c = c0 + t, sz = 0;
                bool snappy = false;
                if (c < k) snappy = chunk_packed_size(fs + (uint64_t)c * fpc, fpc, sec.chunk_bytes, sz);
                // chunk sizes can exceed 16 bits: scan the two halves separately
                uint32_t tl, th;
                uint32_t el = block_excl_sum<kPlanThreads>(sz & 0xFFFFu, &tl, scratch);
                uint32_t eh = block_excl_sum<kPlanThreads>(sz >> 16, &th, scratch);
                uint32_t start = running + el + (eh << 16);
                if (c < k) {
                    ctab[c] = snappy ? kHapChunkSnappy : kHapChunkRaw;
                    stab[4 * c] = (uint8_t)sz; stab[4 * c + 1] = (uint8_t)(sz >> 8);
                    stab[4 * c + 2] = (uint8_t)(sz >> 16); stab[4 * c + 3] = (uint8_t)(sz >> 24);
                    if (snappy) {
                        uint32_t v = sec.chunk_bytes, o = start;
                        while (v >= 0x80) { fo[o++] = (uint8_t)(v | 0x80); v >>= 7; }
                        fo[o++] = (uint8_t)v;
                        for (uint32_t j = 0; j < fpc; j++) {
                            fd[(uint64_t)c * fpc + j] = o;
                            o += fs[(uint64_t)c * fpc + j];
                        }
                    } else {
                        for (uint32_t j = 0; j < fpc; j++) fd[(uint64_t)c * fpc + j] = (start + j * kFragBytes) | kPlaceRawFlag;
                    }
                }
                running += tl + (th << 16);
            }
        } else {
            // hap.c:490-495: the whole texture verbatim
            section_len = sec.bytes;
            if (t == 0) put_section_header(so, hdr, section_len, (kHapChunkRaw << 4) | sec.fmt_nibble);
            const uint32_t data0 = sec_off + hdr;
            for (uint64_t i = t; i < (uint64_t)k * fpc; i += kPlanThreads) {
                uint32_t c = (uint32_t)(i / fpc), j = (uint32_t)(i % fpc);
                fd[i] = (data0 + c * sec.chunk_bytes + j * kFragBytes) | kPlaceRawFlag;
            }
            // bytes / chunks truncates (hap.c:433); the verbatim path still copies every byte (hap.c:492)
            const uint32_t covered = k * sec.chunk_bytes;
            const uint8_t *in = dxt + (uint64_t)frame * sec.in_stride + sec.in_offset;
            for (uint32_t i = covered + t; i < sec.bytes; i += kPlanThreads) fo[data0 + i] = in[i];
        }
        if (t == 0) sec_len_sh[si] = section_len + hdr;
        __syncthreads();
        sec_off += sec_len_sh[si];
    }
    if (t == 0) {
        if (G.sections == 2) put_section_header(fo, G.outer_hdr, sec_off - G.outer_hdr, kSecMultipleImages);  // hap.c:598
        out_used[frame] = sec_off;
    }
}

constexpr int kPlaceThreads = 256;

// grid.x = frames * frags_per_frame
__global__ void __launch_bounds__(kPlaceThreads) hap_place_fragments_kernel(
    FrameGeom G, const uint8_t *__restrict__ dxt, const uint8_t *__restrict__ scratch,
    const uint32_t *__restrict__ frag_size, const uint32_t *__restrict__ frag_dst, uint8_t *__restrict__ out,
    uint64_t out_stride)
{
    const int t = threadIdx.x;
    const uint32_t gfrag = blockIdx.x;
    const uint32_t frame = gfrag / G.frags_per_frame;
    const uint32_t f = gfrag % G.frags_per_frame;
    const SectionGeom &sec = (G.sections == 2 && f >= G.s[1].frag_base) ? G.s[1] : G.s[0];
    const uint32_t fl = f - sec.frag_base;
    const uint32_t chunk = fl / sec.frags_per_chunk, j = fl % sec.frags_per_chunk;
    const uint32_t d = frag_dst[gfrag];
    const uint8_t *src;
    uint32_t n;
    if (d & kPlaceRawFlag) {
        src = dxt + (uint64_t)frame * sec.in_stride + sec.in_offset + (uint64_t)chunk * sec.chunk_bytes + (uint64_t)j * kFragBytes;
        const uint32_t left = sec.chunk_bytes - j * kFragBytes;
        n = left < (uint32_t)kFragBytes ? left : (uint32_t)kFragBytes;
    } else {
        src = scratch + (uint64_t)gfrag * kFragCap;
        n = frag_size[gfrag];
    }
    uint8_t *dst = out + (uint64_t)frame * out_stride + (d & ~kPlaceRawFlag);
    // destination alignment is arbitrary (headers, tables and earlier chunks are byte-sized):
    // byte-copy to the first 4-byte boundary, then aligned words assembled from two source words
    uint32_t head = (uint32_t)((4 - ((uintptr_t)dst & 3)) & 3);
    if (head > n) head = n;
    if ((uint32_t)t < head) dst[t] = src[t];
    const uint32_t nw = (n - head) >> 2;
    const uint8_t *s2 = src + head;
    uint32_t *d32 = reinterpret_cast<uint32_t *>(dst + head);
    const uint32_t mis = (uint32_t)((uintptr_t)s2 & 3);
    const uint32_t *s32 = reinterpret_cast<const uint32_t *>(s2 - mis);
    uint32_t nw_done = nw;
    if (mis == 0) {
        for (uint32_t i = t; i < nw; i += kPlaceThreads) d32[i] = s32[i];
    } else {
        // the funnel reads word i+1, which for the last word would reach past the source bytes:
        // stop one word early and leave the rest to the byte tail
        nw_done = nw ? nw - 1 : 0;
        for (uint32_t i = t; i < nw_done; i += kPlaceThreads) d32[i] = __funnelshift_r(s32[i], s32[i + 1], 8 * mis);
    }
    const uint32_t done = head + (nw_done << 2);
    if ((uint32_t)t < n - done) dst[done + t] = src[done + t];
}

}  // namespace hapb200
