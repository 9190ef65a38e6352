// tests/emu/bc_twin.cc -- TEST INFRASTRUCTURE ONLY.
// Host build of the per-block encoder math in hap_b200/csrc/bc_block.cuh (the very same source the
// CUDA kernels compile), exported as plain C image-level functions.  Used by the CPU tests to hold
// the algorithm to the oracle's quality bar without a GPU, and by the GPU tests to check that the
// device result is bit-identical to the host result of the same source.
#define HAPB200_EMU
#include "bc_decode.cuh"

using namespace hapb200;

static void gather(const uint8_t *rgba, int w, int bx, int by, uint32_t px[16])
{
    for (int t = 0; t < 16; t++) memcpy(&px[t], rgba + 4 * ((size_t)(4 * by + t / 4) * w + 4 * bx + t % 4), 4);
}

extern "C" void twin_encode(const uint8_t *rgba, int w, int h, int kind, uint8_t *out)
{
    // kind: 0 dxt1, 1 dxt5, 2 ycocg-dxt5, 3 rgtc1 (alpha channel), 4 ycocg-dxt5 with the chroma refinement option
    for (int by = 0; by < h / 4; by++)
        for (int bx = 0; bx < w / 4; bx++) {
            uint32_t px[16];
            gather(rgba, w, bx, by, px);
            size_t bi = (size_t)by * (w / 4) + bx;
            if (kind == 0) { Block8 c = encode_dxt1(px); memcpy(out + 8 * bi, &c, 8); }
            else if (kind == 3) { Block8 a = encode_rgtc1_alpha(px); memcpy(out + 8 * bi, &a, 8); }
            else {
                Block8 a, c;
                if (kind == 1) encode_dxt5(px, a, c); else if (kind == 4) encode_ycocg_dxt5<true>(px, a, c); else encode_ycocg_dxt5<false>(px, a, c);
                memcpy(out + 16 * bi, &a, 8);
                memcpy(out + 16 * bi + 8, &c, 8);
            }
        }
}

// Host build of the block DECODER math (bc_decode.cuh): blocks -> RGBA8, same block order and kinds as above.
extern "C" void twin_decode(const uint8_t *blocks, int w, int h, int kind, uint8_t *rgba)
{
    const int kinds[4] = {kBcDxt1, kBcDxt5, kBcYCoCg, kBcRgtc1};
    for (int by = 0; by < h / 4; by++)
        for (int bx = 0; bx < w / 4; bx++) {
            size_t bi = (size_t)by * (w / 4) + bx;
            uint32_t v[4] = {0, 0, 0, 0}, px[16];
            memcpy(v, blocks + (kind == 0 || kind == 3 ? 8 : 16) * bi, kind == 0 || kind == 3 ? 8 : 16);
            decode_block(kinds[kind], v[0], v[1], v[2], v[3], px);
            for (int t = 0; t < 16; t++) memcpy(rgba + 4 * ((size_t)(4 * by + t / 4) * w + 4 * bx + t % 4), &px[t], 4);
        }
}
