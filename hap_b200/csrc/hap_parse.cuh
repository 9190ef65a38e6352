// hap_b200/csrc/hap_parse.cuh -- device-side frame parser for the batched decode path.
//
// hap_decode_single_texture's serial prepass (/root/reference/source/hap.c:748-843: section type ->
// compressor/format, Decode-Instructions tables -> one HapChunkDecodeInfo per chunk, the running
// source/destination offsets and the output-size check) run by one thread per frame, so a batch of
// device-resident frames is decoded without the host ever reading a header.
#pragma once
#include "hap_wire.h"
#include "snappy_decode.cuh"

namespace hapb200 {

constexpr uint32_t kHapChunkSkip = 0;  // ChunkJob.compressor of an unused slot

// frame f: in + f*in_stride, in_bytes[f] bytes.  jobs: [frames][max_chunks].
// Frame f's texture goes to out + f*out_stride and may use out_capacity bytes there (the slot of a batch is often
// wider than what one texture may fill: two textures of a Hap Q Alpha frame share a slot).
// results[f] = HapResult of the prepass; whole_section[f] = 1 when the texture is one 0xB? stream
// (its errors map to Internal_Error, hap.c:891-903).
__global__ void hap_parse_frames_kernel(const uint8_t *__restrict__ in, uint64_t in_stride,
                                        const unsigned long long *__restrict__ in_bytes, uint32_t frames,
                                        uint32_t index, uint32_t max_chunks, uint8_t *__restrict__ out,
                                        uint64_t out_stride, uint64_t out_capacity, ChunkJob *__restrict__ jobs,
                                        unsigned long long *__restrict__ used, uint32_t *__restrict__ formats,
                                        uint32_t *__restrict__ results, uint32_t *__restrict__ whole_section)
{
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= frames) return;
    const uint8_t *frame = in + (uint64_t)f * in_stride;
    uint8_t *dst = out + (uint64_t)f * out_stride;
    ChunkJob *fj = jobs + (uint64_t)f * max_chunks;
    for (uint32_t c = 0; c < max_chunks; c++) {
        fj[c].compressor = kHapChunkSkip;
        fj[c].status = HapResult_No_Error;
    }
    uint32_t fmt = 0, whole = 0;
    unsigned long long produced = 0;
    uint32_t r = HapResult_No_Error;
    const unsigned long long nb = in_bytes[f];
    Located loc;
    if (index > 1) r = HapResult_Bad_Arguments;
    if (r == HapResult_No_Error) r = locate_texture(frame, (uint32_t)nb, index, loc);
    if (r == HapResult_No_Error) {
        const uint8_t *sec = frame + loc.offset;
        const uint32_t compressor = (loc.type >> 4) & 0xF;
        fmt = format_from_nibble(loc.type & 0xF);
        if (fmt == 0) {
            r = HapResult_Bad_Frame;
        } else if (compressor == kHapComplex) {
            ChunkTables t;
            t.count = 0;
            r = parse_decode_instructions(sec, loc.len, t);
            FragmentIndex ix;
            const bool have_ix = r == HapResult_No_Error && locate_fragment_index(frame, (uint32_t)nb, ix);
            if (r == HapResult_No_Error && t.count > 0) {
                if ((uint32_t)t.count > max_chunks) {
                    r = HapResult_Bad_Arguments;
                } else {
                    uint64_t in_run = 0, out_run = 0;
                    for (int i = 0; i < t.count && r == HapResult_No_Error; i++) {
                        const uint32_t cc = sec[t.compressors + i];
                        const uint32_t sz = rd_le32(sec + t.sizes + 4 * i);
                        const uint64_t start = t.data + (t.offsets != 0xFFFFFFFFu ? (uint64_t)rd_le32(sec + t.offsets + 4 * i) : in_run);
                        in_run += sz;
                        if (start + sz > loc.len) { r = HapResult_Bad_Frame; break; }  // Q9 tightening
                        uint32_t usz = sz;
                        if (cc == kHapChunkSnappy && !snappy_preamble(sec + start, sz, usz)) { r = HapResult_Bad_Frame; break; }
                        fj[i].src = sec + start;
                        fj[i].src_bytes = sz;
                        fj[i].dst = dst + out_run;
                        fj[i].dst_bytes = usz;
                        fj[i].compressor = (cc == kHapChunkSnappy || cc == kHapChunkRaw) ? cc : 0xFFu;  // 0xFF -> Bad_Frame in K7
                        fj[i].index = nullptr;
                        fj[i].index_bytes = 0;
                        uint32_t ioff, ibytes;
                        if (cc == kHapChunkSnappy && have_ix && fragment_index_record(frame, ix, index, (uint32_t)t.count, (uint32_t)i, ioff, ibytes)) {
                            fj[i].index = frame + ioff;
                            fj[i].index_bytes = ibytes;
                        }
                        out_run += usz;
                    }
                    if (r == HapResult_No_Error && out_run > out_capacity) r = HapResult_Buffer_Too_Small;
                    if (r == HapResult_No_Error) produced = out_run;
                    else for (int i = 0; i < t.count; i++) fj[i].compressor = kHapChunkSkip;
                }
            }
        } else if (compressor == kHapChunkSnappy) {
            uint32_t usz = 0;
            whole = 1;
            if (max_chunks < 1) r = HapResult_Bad_Arguments;
            else if (!snappy_preamble(sec, loc.len, usz)) r = HapResult_Internal_Error;  // hap.c:891-894
            else if (usz > out_capacity) r = HapResult_Buffer_Too_Small;
            else {
                fj[0].src = sec; fj[0].src_bytes = loc.len; fj[0].dst = dst; fj[0].dst_bytes = usz;
                fj[0].compressor = kHapChunkSnappy; fj[0].index = nullptr; fj[0].index_bytes = 0;
                produced = usz;
            }
        } else if (compressor == kHapChunkRaw) {
            if (max_chunks < 1) r = HapResult_Bad_Arguments;
            else if (loc.len > out_capacity) r = HapResult_Buffer_Too_Small;
            else {
                fj[0].src = sec; fj[0].src_bytes = loc.len; fj[0].dst = dst; fj[0].dst_bytes = loc.len;
                fj[0].compressor = kHapChunkRaw; fj[0].index = nullptr; fj[0].index_bytes = 0;
                produced = loc.len;
            }
        } else {
            r = HapResult_Bad_Frame;
        }
    }
    results[f] = r;
    formats[f] = fmt;
    used[f] = r == HapResult_No_Error ? produced : 0;
    whole_section[f] = whole;
}

// hap.c:867-875: the first chunk error wins
__global__ void hap_collect_status_kernel(const ChunkJob *__restrict__ jobs, uint32_t frames, uint32_t max_chunks,
                                          const uint32_t *__restrict__ whole_section, uint32_t *__restrict__ results,
                                          unsigned long long *__restrict__ used)
{
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= frames) return;
    if (results[f] != HapResult_No_Error) return;
    const ChunkJob *fj = jobs + (uint64_t)f * max_chunks;
    for (uint32_t c = 0; c < max_chunks; c++) {
        if (fj[c].compressor != kHapChunkSkip && fj[c].status != HapResult_No_Error) {
            results[f] = whole_section[f] ? (uint32_t)HapResult_Internal_Error : fj[c].status;
            used[f] = 0;
            return;
        }
    }
}

// RGBA decode path: a texture must have the format and size the caller's codec implies; the first
// failing texture of a frame decides that frame's result (merged into `merged`, which may alias `res`).
__global__ void hap_check_texture_kernel(uint32_t frames, const unsigned long long *__restrict__ used,
                                         const uint32_t *__restrict__ formats, const uint32_t *res,
                                         unsigned long long want_bytes, uint32_t want_format, uint32_t *merged)
{
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= frames) return;
    uint32_t r = res[f];
    if (r == HapResult_No_Error && (formats[f] != want_format || used[f] != want_bytes)) r = HapResult_Bad_Frame;
    if (res == merged) merged[f] = r;
    else if (merged[f] == HapResult_No_Error) merged[f] = r;
}

}  // namespace hapb200
