// hap_b200/csrc/bc_decode.cuh -- K8: DXT1 / DXT5 / scaled-YCoCg-DXT5 / RGTC1 blocks -> RGBA8.
//
// The step AFTER HapDecode in every player: in the reference's world it is the GPU texture unit plus
// a YCoCg fragment shader (/root/reference/README.md:4, documentation/HapVideoDRAFT.md:8,81).  Here it
// is an optional tail kernel so a decoded frame can be handed on as plain RGBA.  Integer arithmetic
// throughout (truncating S3TC interpolation, exact quarter-unit YCoCg), bit-exact against
// oracle/bc_oracle.c, which is itself cross-checked against Pillow's bcn decoder.
// One thread per block: reads 8/16 contiguous bytes, writes four 16-byte row pieces; a warp therefore
// writes 512 contiguous bytes per pixel row.
#pragma once
#include "bc_encode.cuh"

namespace hapb200 {

HAP_HD void bc1_palette(uint32_t c0, uint32_t c1, bool force4, uint32_t pal[4])
{
    // pal entries packed r | g<<8 | b<<16 | a<<24
    uint32_t r0 = expand5(c0 >> 11), g0 = expand6((c0 >> 5) & 63), b0 = expand5(c0 & 31);
    uint32_t r1 = expand5(c1 >> 11), g1 = expand6((c1 >> 5) & 63), b1 = expand5(c1 & 31);
    pal[0] = r0 | (g0 << 8) | (b0 << 16) | 0xFF000000u;
    pal[1] = r1 | (g1 << 8) | (b1 << 16) | 0xFF000000u;
    if (force4 || c0 > c1) {
        pal[2] = ((2 * r0 + r1) / 3) | (((2 * g0 + g1) / 3) << 8) | (((2 * b0 + b1) / 3) << 16) | 0xFF000000u;
        pal[3] = ((r0 + 2 * r1) / 3) | (((g0 + 2 * g1) / 3) << 8) | (((b0 + 2 * b1) / 3) << 16) | 0xFF000000u;
    } else {
        pal[2] = ((r0 + r1) / 2) | (((g0 + g1) / 2) << 8) | (((b0 + b1) / 2) << 16) | 0xFF000000u;
        pal[3] = 0;  // transparent black
    }
}

HAP_HD void bc4_palette(uint32_t a0, uint32_t a1, uint32_t pal[8])
{
    pal[0] = a0;
    pal[1] = a1;
    if (a0 > a1) {
#pragma unroll
        for (uint32_t i = 2; i < 8; i++) pal[i] = ((8 - i) * a0 + (i - 1) * a1) / 7;
    } else {
#pragma unroll
        for (uint32_t i = 2; i < 6; i++) pal[i] = ((6 - i) * a0 + (i - 1) * a1) / 5;
        pal[6] = 0;
        pal[7] = 255;
    }
}

// texel t of a decoded block; out[16] packed RGBA
HAP_HD void decode_block(int kind, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t out[16])
{
    if (kind == kBcDxt1) {
        uint32_t pal[4];
        bc1_palette(w0 & 0xFFFF, w0 >> 16, false, pal);
#pragma unroll
        for (int t = 0; t < 16; t++) {
            uint32_t i = (w1 >> (2 * t)) & 3;
            out[t] = i == 0 ? pal[0] : i == 1 ? pal[1] : i == 2 ? pal[2] : pal[3];
        }
        return;
    }
    uint32_t ap[8];
    bc4_palette(w0 & 0xFF, (w0 >> 8) & 0xFF, ap);
    const uint64_t abits = ((uint64_t)w1 << 16) | (w0 >> 16);
    if (kind == kBcRgtc1) {
#pragma unroll
        for (int t = 0; t < 16; t++) {
            uint32_t i = (uint32_t)(abits >> (3 * t)) & 7, a = ap[0];
#pragma unroll
            for (uint32_t k = 1; k < 8; k++) a = i == k ? ap[k] : a;
            out[t] = a;  // one byte per texel, caller packs
        }
        return;
    }
    uint32_t pal[4];
    bc1_palette(w2 & 0xFFFF, w2 >> 16, true, pal);
#pragma unroll
    for (int t = 0; t < 16; t++) {
        uint32_t i = (uint32_t)(abits >> (3 * t)) & 7, a = ap[0];
#pragma unroll
        for (uint32_t k = 1; k < 8; k++) a = i == k ? ap[k] : a;
        uint32_t ci = (w3 >> (2 * t)) & 3;
        uint32_t c = ci == 0 ? pal[0] : ci == 1 ? pal[1] : ci == 2 ? pal[2] : pal[3];
        if (kind == kBcDxt5) {
            out[t] = (c & 0x00FFFFFFu) | (a << 24);
        } else {
            // scaled YCoCg: (R',G',B',A) = (Co', Cg', scale bits, Y); exact quarter units, round half up
            int sc = (int)((c >> 16) & 0xFF) >> 3;         // 0,1,3 -> scale 1,2,4
            int q = sc >= 3 ? 1 : sc >= 1 ? (sc == 2 ? 1 : 2) : 4;
            int co4 = ((int)(c & 0xFF) - 128) * q, cg4 = ((int)((c >> 8) & 0xFF) - 128) * q, y4 = 4 * (int)a;
            uint32_t R = (uint32_t)hap_clampi((y4 + co4 - cg4 + 2) >> 2, 0, 255);
            uint32_t G = (uint32_t)hap_clampi((y4 + cg4 + 2) >> 2, 0, 255);
            uint32_t B = (uint32_t)hap_clampi((y4 - co4 - cg4 + 2) >> 2, 0, 255);
            out[t] = R | (G << 8) | (B << 16) | 0xFF000000u;
        }
    }
}

struct BcDecodeGeom {
    uint32_t blocks_x, blocks_y;
    uint32_t row_bytes;      // RGBA row stride
    uint32_t merge_alpha;    // 1: `alpha_blocks` holds an RGTC1 plane to merge into A (Hap Q Alpha)
    uint64_t in_stride;      // DXT bytes per frame
    uint64_t alpha_stride;
    uint64_t frame_bytes;    // RGBA frame stride
};

// grid = (ceil(blocks/kBcThreads), frames).  kBcRgtc1 alone writes the value to all of R,G,B with A=255.
template <int KIND>
__global__ void __launch_bounds__(kBcThreads) bc_decode_kernel(const uint8_t *__restrict__ blocks,
                                                                const uint8_t *__restrict__ alpha_blocks,
                                                                BcDecodeGeom G, uint8_t *__restrict__ rgba)
{
    const uint32_t nblocks = G.blocks_x * G.blocks_y;
    const uint32_t bi = blockIdx.x * kBcThreads + threadIdx.x;
    if (bi >= nblocks) return;
    const uint32_t by = bi / G.blocks_x, bx = bi - by * G.blocks_x;
    const uint8_t *in = blocks + (uint64_t)blockIdx.y * G.in_stride;
    uint32_t px[16];
    if (KIND == kBcDxt1 || KIND == kBcRgtc1) {
        uint2 v = reinterpret_cast<const uint2 *>(in)[bi];
        decode_block(KIND, v.x, v.y, 0, 0, px);
        if (KIND == kBcRgtc1) {
#pragma unroll
            for (int t = 0; t < 16; t++) px[t] = px[t] * 0x010101u | 0xFF000000u;
        }
    } else {
        uint4 v = reinterpret_cast<const uint4 *>(in)[bi];
        decode_block(KIND, v.x, v.y, v.z, v.w, px);
        if (G.merge_alpha) {
            uint2 a = reinterpret_cast<const uint2 *>(alpha_blocks + (uint64_t)blockIdx.y * G.alpha_stride)[bi];
            uint32_t al[16];
            decode_block(kBcRgtc1, a.x, a.y, 0, 0, al);
#pragma unroll
            for (int t = 0; t < 16; t++) px[t] = (px[t] & 0x00FFFFFFu) | (al[t] << 24);
        }
    }
    uint8_t *dst = rgba + (uint64_t)blockIdx.y * G.frame_bytes + (uint64_t)(4 * by) * G.row_bytes + 16u * bx;
#pragma unroll
    for (int row = 0; row < 4; row++)
        *reinterpret_cast<uint4 *>(dst + (uint64_t)row * G.row_bytes) =
            make_uint4(px[4 * row], px[4 * row + 1], px[4 * row + 2], px[4 * row + 3]);
}

}  // namespace hapb200
