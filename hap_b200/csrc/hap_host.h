// hap_b200/csrc/hap_host.h -- host-only frame geometry: everything hap_encode_texture / HapEncode
// (/root/reference/source/hap.c:355-604) fix BEFORE compression, computed once per call and handed to
// the kernels as a FrameGeom.
#pragma once
#include "hap_wire.h"
#include "snappy_encode.cuh"

namespace hapb200 {

struct TextureArgs {
    uint64_t bytes;
    uint32_t format, compressor, chunks;
};

// hap.c:367-385 (argument checks of one texture, pointers excluded)
inline uint32_t validate_texture_args(const TextureArgs &a)
{
    if (a.bytes == 0 || nibble_from_format(a.format) == 0 ||
        (a.compressor != HapCompressorNone && a.compressor != HapCompressorSnappy))
        return HapResult_Bad_Arguments;
    return HapResult_No_Error;
}

// Fills G for `count` textures.  in_offset / in_stride are left to the caller.
// Returns Bad_Arguments for sizes this implementation cannot address (sections must stay < 2 GiB).
inline uint32_t build_frame_geom(uint32_t count, const TextureArgs *tex, FrameGeom &G)
{
    G.sections = count;
    G.outer_hdr = 0;
    uint32_t frag_base = 0;
    for (uint32_t i = 0; i < count; i++) {
        const TextureArgs &a = tex[i];
        SectionGeom &s = G.s[i];
        if (a.bytes >= (1ull << 31)) return HapResult_Bad_Arguments;
        s.bytes = (uint32_t)a.bytes;
        s.want_snappy = a.compressor == HapCompressorSnappy;
        s.fmt_nibble = nibble_from_format(a.format);
        s.top_hdr = a.bytes > kU24Max ? 8 : 4;  // hap.c:398-405
        // The reference limits the chunk count only on the Snappy path (hap.c:421); a verbatim
        // texture has no chunks on the wire, so any partition into copy units is equivalent.
        s.chunks = limited_chunk_count(a.bytes, a.format, a.chunks);
        if (s.want_snappy && a.bytes + decode_instructions_length(s.chunks) + 4 > kU24Max) s.top_hdr = 8;  // hap.c:425-428
        s.chunk_bytes = (uint32_t)(a.bytes / s.chunks);  // hap.c:433 (truncating)
        if (s.chunk_bytes == 0) {
            // fewer bytes than one block: nothing to chunk; one verbatim unit
            s.chunks = 1;
            s.chunk_bytes = (uint32_t)a.bytes;
        }
        s.frags_per_chunk = (s.chunk_bytes + kFragBytes - 1) / kFragBytes;
        s.inv_frags_per_chunk = s.frags_per_chunk <= 1 ? 0xFFFFFFFFu : (uint32_t)(0x100000000ull / s.frags_per_chunk);
        s.period_words = (a.format == HapFmt_RGB_DXT1 || a.format == HapFmt_A_RGTC1) ? 2 : 4;
        s.compress = s.want_snappy && (s.chunk_bytes % 8 == 0);
        s.frag_base = frag_base;
        s.in_offset = 0;
        s.in_stride = 0;
        uint64_t frags = (uint64_t)s.chunks * s.frags_per_chunk;
        if (frags + frag_base >= (1ull << 31)) return HapResult_Bad_Arguments;
        frag_base += (uint32_t)frags;
    }
    G.frags_per_frame = frag_base;
    G.inv_frags_per_frame = frag_base <= 1 ? 0xFFFFFFFFu : (uint32_t)(0x100000000ull / frag_base);
    if (count == 2) {
        // hap.c:563-576 (uses the REQUESTED chunk counts)
        uint64_t worst = 0;
        for (uint32_t i = 0; i < 2; i++) worst += tex[i].bytes + decode_instructions_length(tex[i].chunks) + 4;
        G.outer_hdr = worst > kU24Max ? 8 : 4;
    }
    return HapResult_No_Error;
}

}  // namespace hapb200
