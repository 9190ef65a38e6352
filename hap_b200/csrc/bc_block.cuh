// hap_b200/csrc/bc_block.cuh -- K1-K4 block math: one 4x4 RGBA block -> one 8/16-byte S3TC/RGTC block.
//
// The block compressors sit UPSTREAM of the reference (which takes already-compressed DXT bytes,
// /root/reference/source/hap.h:82-104; formats per documentation/HapVideoDRAFT.md:22-27); the
// north star makes them the encode hot kernel.  Quality bar: PSNR within 0.1 dB of a squish-HIGH
// (iterative cluster fit) encode, which oracle/bc_oracle.c restates.
//
// Work decomposition: ONE THREAD PER BLOCK.  SURVEY.md H1: at the bandwidth target a 4x4 block has
// ~13 warp-instructions if a warp owns it but ~420 thread-instructions if a thread owns it; all 16
// texels live in registers, every reduction is a private serial sum (no shuffles, no idle lanes),
// and coalescing is restored by staging tiles through shared memory in the kernels (bc_encode.cuh).
//
// Colour fit: principal axis of the block's covariance (power iteration), then REFINE rounds of
// [project texels onto the current segment -> 4 clusters] + [2x2 least squares for the endpoints
// given those clusters] -- the same normal equations cluster fit solves, for the partition the
// current endpoints imply instead of all 969 -- then 5:6:5 rounding and a final nearest-index pass
// against the decoder's integer palette.
//
// Floating point is written with explicit fused multiply-adds (hap_fma) and the translation unit is
// built with -fmad=false, so the CPU twin used by the tests (tests/emu/bc_twin.cc) reproduces the
// GPU result bit for bit: IEEE add/mul/div/sqrt and fma are exact on both sides.
#pragma once
#include "simt.h"
#include <math.h>

namespace hapb200 {

#ifdef HAPB200_EMU
#define HAP_HD inline
static inline float hap_fma(float a, float b, float c) { return fmaf(a, b, c); }
#else
#define HAP_HD __host__ __device__ __forceinline__
HAP_HD float hap_fma(float a, float b, float c) { return fmaf(a, b, c); }
#endif


struct Block8 { uint32_t lo, hi; };

// sum over the four bytes of a (unsigned) times the four bytes of b (signed), plus c: one DP4A instruction
#if defined(HAPB200_EMU) || !defined(__CUDA_ARCH__)
HAP_HD int hap_dp4a_us(uint32_t a, uint32_t b, int c)
{
    for (int k = 0; k < 4; k++) c += (int)((a >> (8 * k)) & 0xFF) * (int)(int8_t)((b >> (8 * k)) & 0xFF);
    return c;
}
#else
__device__ __forceinline__ int hap_dp4a_us(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
#endif

HAP_HD int hap_clampi(int v, int lo, int hi) { return v < lo ? lo : v > hi ? hi : v; }

// Three-input min/max, clamp(a+b, 0, c) and saturate: single instructions on sm_100a (VIMNMX3, FMNMX3,
// VIADDMNMX.RELU, FFMA.SAT); plain C++ on the host so that the CPU twin computes the same values.
#if defined(HAPB200_EMU) || !defined(__CUDA_ARCH__)
HAP_HD int hap_max3(int a, int b, int c) { int m = a > b ? a : b; return m > c ? m : c; }
HAP_HD int hap_min3(int a, int b, int c) { int m = a < b ? a : b; return m < c ? m : c; }
HAP_HD int hap_addmin_relu(int a, int b, int c) { int s = a + b; s = s < c ? s : c; return s < 0 ? 0 : s; }
#else
__device__ __forceinline__ int hap_max3(int a, int b, int c) { return __vimax3_s32(a, b, c); }
__device__ __forceinline__ int hap_min3(int a, int b, int c) { return __vimin3_s32(a, b, c); }
__device__ __forceinline__ int hap_addmin_relu(int a, int b, int c) { return __viaddmin_s32_relu(a, b, c); }
#endif
HAP_HD float hap_fmax3(float a, float b, float c) { return fmaxf(a, fmaxf(b, c)); }
HAP_HD float hap_fmin3(float a, float b, float c) { return fminf(a, fminf(b, c)); }
HAP_HD float hap_sat(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }
#if defined(HAPB200_EMU) || !defined(__CUDA_ARCH__)
HAP_HD uint32_t hap_float_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
#else
__device__ __forceinline__ uint32_t hap_float_bits(float f) { return __float_as_uint(f); }
#endif
constexpr float kRoundMagic = 12582912.0f;      // 1.5 * 2^23: x + magic rounds x to the nearest integer (ties to even) ...
constexpr uint32_t kRoundMagicBits = 0x4B400000u;  // ... which then sits in the low mantissa bits

// ---- BC4 / RGTC1: 16 values -> 8 bytes ------------------------------------------------------------
// 8-value mode (a0 > a1): palette a0, a1, then 6 interpolants (decoder: ((8-i)a0 + (i-1)a1)/7).
// Endpoints are the block's max/min (a least-squares refinement of them was measured to change nothing);
// each texel takes the nearest of the decoder's truncating palette values.
// decoder palette, truncating: ((7-L)*a1 + L*a0)/7 = a1 + floor(L*(a0-a1)/7); L*(a0-a1) <= 1785, where
// floor(q/7) == (q*9363)>>16 exactly (checked exhaustively)
HAP_HD int bc4_level_value(int L, int a0, int a1) { return a1 + (int)(((uint32_t)(L * (a0 - a1)) * 9363u) >> 16); }

// ---- BC4 ------------------------------------------------------------------------------------------
// v[t] = the value in 1/UNIT steps, UNIT a multiple of 7 (UNIT = 7: 7 * alpha; UNIT = 28: 7 * (R + 2G + B), luma
// with the two bits kept that rounding Y to 8 bits throws away -- the factors ride for free in the DP4A weights);
// vmax / vmin = max / min of v.  Endpoints = the rounded max and min of the block (a least-squares refinement of
// them was measured to change nothing).
// A texel's level L (0 = the min end ... 7 = the max end) is the nearest of the 8 palette values.  The decoder's
// palette truncates: value(L) = min + floor(L d / 7), i.e. the ideal ramp minus 0/7 .. 6/7 of a grey level, 3/7 on
// average -- so the texel is moved UP by 3/7 (exactly 3 UNIT/7 here) before it is rounded onto the ideal ramp.
// That is one VIADDMNMX (clamp(v - UNIT min + 3 UNIT/7, 0, range)), one fixed-point multiply and one shift per
// texel instead of seven threshold compares (measured against the exhaustive choice: see test_block_quality_cpu).
// Levels are packed 8 to a word and turned into DXT index numbering on all eight 3-bit fields at once.
HAP_HD uint32_t bc4_levels_to_indices8(uint32_t levels)
{
    const uint32_t x = levels ^ 0xFFFFFFu;                            // M = 7 - L: 0 = the max end (a0)
    const uint32_t m0 = 0x249249u;                                     // bit 0 of each of the eight fields
    const uint32_t x1 = x >> 1, x2 = x >> 2;
    const uint32_t sevens = x & x1 & x2 & m0, zeros = ~(x | x1 | x2) & m0;
    return (x & ~(sevens * 7u)) + m0 - zeros;                         // M: 0 -> 0, 7 -> 1, else M + 1
}

template <int UNIT>
HAP_HD Block8 encode_bc4_scaled(const int v[16], int vmax, int vmin)
{
    static_assert(UNIT % 7 == 0, "UNIT carries the factor 7");
    const int mx = (int)((uint32_t)(vmax + UNIT / 2) / (uint32_t)UNIT), mn = (int)((uint32_t)(vmin + UNIT / 2) / (uint32_t)UNIT);
    Block8 out;
    out.lo = (uint32_t)mx | ((uint32_t)mn << 8);
    out.hi = 0;
    if (mx == mn) return out;  // a0 == a1 selects the 6-value mode; index 0 decodes to a0 in both modes
    const int d = mx - mn, range = UNIT * d;
    const int r7 = d - 7 * (int)(((uint32_t)d * 9363u) >> 16);                       // d mod 7 (d <= 255)
    // integer inputs (UNIT = 7): 3/7 is exact for every d.  Quarter-step inputs: the best constant per d mod 7,
    // 0, 8, 9, 11, 11, 9, 8 twenty-eighths (0 when d is a multiple of 7: nothing is truncated then)
    const int bias = UNIT % 28 == 0 ? (int)((0x89BB980u >> (4 * r7)) & 15u) * (UNIT / 28) : 3 * (UNIT / 7);
    const int c0 = bias - UNIT * mn;
    // 7 * 2^20 / range, rounded: (t * mul + 2^19) >> 20 is round(7 t / range) for every t <= range (<= 7140)
    const int mul = (int)(7340032.0f * (1.0f / (float)range) + 0.5f);
    uint32_t a = 0, b = 0;
#pragma unroll
    for (int t = 0; t < 8; t++) {
        const uint32_t La = (uint32_t)(hap_addmin_relu(v[t], c0, range) * mul + (1 << 19)) >> 20;
        const uint32_t Lb = (uint32_t)(hap_addmin_relu(v[t + 8], c0, range) * mul + (1 << 19)) >> 20;
        a += La << (3 * t);
        b += Lb << (3 * t);
    }
    a = bc4_levels_to_indices8(a);
    b = bc4_levels_to_indices8(b);
    out.lo |= a << 16;
    out.hi = (a >> 16) | (b << 8);
    return out;
}

// ---- BC1 colour block: 16 (r,g,b) in 0..255 -> 8 bytes, always 4-colour mode ---------------------
HAP_HD uint32_t expand5(uint32_t c) { return (c << 3) | (c >> 2); }
HAP_HD uint32_t expand6(uint32_t c) { return (c << 2) | (c >> 4); }

// Picks, for one channel, the pair of grid values around the float endpoints (a,b) that minimises the
// cluster-weighted squared error.  levels = 31 or 63; values stay in 0..255 units.
HAP_HD float snap_channel(float &a, float &b, float a2, float b2, float ab, float ax, float bx, float levels)
{
    const float to_grid = levels * (1.0f / 255.0f), from_grid = 255.0f / levels;
    float a_lo = floorf(a * to_grid), b_lo = floorf(b * to_grid);
    float best = 1e30f, best_a = a, best_b = b;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        float ga = fminf(a_lo + (float)(k & 1), levels), gb = fminf(b_lo + (float)(k >> 1), levels);
        // the decoder expands by bit replication, which equals round(g*255/levels) for 5 and 6 bits
        float ca = floorf(hap_fma(ga, from_grid, 0.5f)), cb = floorf(hap_fma(gb, from_grid, 0.5f));
        float e = hap_fma(ca * ca, a2, hap_fma(cb * cb, b2, 2.0f * (ca * cb * ab - ca * ax - cb * bx)));
        if (e < best) { best = e; best_a = ca; best_b = cb; }
    }
    a = best_a;
    b = best_b;
    return best;  // error of the chosen pair, up to the constant sum of squares of the channel
}

// One round of [project texels onto the segment a-b -> 4 clusters] + accumulate the normal-equation sums.
struct FitSums {
    float a2, b2, ab, axr, axg, axb, bxr, bxg, bxb;
};
template <bool HAS_B>
HAP_HD bool cluster_sums(const float r[16], const float g[16], const float b[16], float ar, float ag, float ab_, float br,
                         float bg, float bb, FitSums &S)
{
    float dr = br - ar, dg = bg - ag, db = HAS_B ? bb - ab_ : 0.0f;
    float dd = HAS_B ? hap_fma(dr, dr, hap_fma(dg, dg, db * db)) : hap_fma(dr, dr, dg * dg);
    if (dd < 1e-6f) return false;
    float scale = 3.0f / dd;
    S.a2 = S.b2 = S.ab = S.axr = S.axg = S.axb = S.bxr = S.bxg = S.bxb = 0.f;
#pragma unroll
    for (int t = 0; t < 16; t++) {
        float s = (HAS_B ? hap_fma(r[t] - ar, dr, hap_fma(g[t] - ag, dg, (b[t] - ab_) * db))
                         : hap_fma(r[t] - ar, dr, (g[t] - ag) * dg)) * scale;
        float q = fminf(fmaxf(floorf(s + 0.5f), 0.0f), 3.0f);
        float be = q * (1.0f / 3.0f), al = 1.0f - be;
        S.a2 = hap_fma(al, al, S.a2); S.b2 = hap_fma(be, be, S.b2); S.ab = hap_fma(al, be, S.ab);
        S.axr = hap_fma(al, r[t], S.axr); S.axg = hap_fma(al, g[t], S.axg);
        S.bxr = hap_fma(be, r[t], S.bxr); S.bxg = hap_fma(be, g[t], S.bxg);
        if (HAS_B) { S.axb = hap_fma(al, b[t], S.axb); S.bxb = hap_fma(be, b[t], S.bxb); }
    }
    return S.a2 * S.b2 - S.ab * S.ab >= 1e-4f;
}

// REFINE: least-squares rounds; RESNAP: extra rounds of Lloyd on the snapped endpoints; EXACT: final
// indices by true nearest palette colour (else by projection onto the palette segment).
// r,g,b arrive PRE-MULTIPLIED by the metric (sr,sg,sb) so that plain Euclidean distance in that space
// is the error to minimise (RGB: 1,1,1; scaled YCoCg: sqrt2, sqrt3 -- an error (dCo,dCg) costs
// 2 dCo^2 + 3 dCg^2 in RGB); endpoints go back to storage units before they meet the 5:6:5 grid.
// HAS_B = false: the third channel is constant over the block (scaled YCoCg carries its scale code there) and
// drops out of every sum.
template <int REFINE, int RESNAP, bool EXACT, bool HAS_B = true, int STARTS = 1>
HAP_HD Block8 encode_colour_block(const float r[16], const float g[16], const float b[16], int fixed_blue5 = -1,
                                  float sr = 1.0f, float sg = 1.0f, float sb = 1.0f)
{
    const float isr = 1.0f / sr, isg = 1.0f / sg, isb = 1.0f / sb;
    // mean and covariance
    float mr = 0.f, mg = 0.f, mb = 0.f;
#pragma unroll
    for (int t = 0; t < 16; t++) { mr += r[t]; mg += g[t]; if (HAS_B) mb += b[t]; }
    mr *= 0.0625f; mg *= 0.0625f; mb = HAS_B ? mb * 0.0625f : b[0];
    float crr = 0.f, crg = 0.f, crb = 0.f, cgg = 0.f, cgb = 0.f, cbb = 0.f;
#pragma unroll
    for (int t = 0; t < 16; t++) {
        float dr = r[t] - mr, dg = g[t] - mg;
        crr = hap_fma(dr, dr, crr); crg = hap_fma(dr, dg, crg); cgg = hap_fma(dg, dg, cgg);
        if (HAS_B) {
            float db = b[t] - mb;
            crb = hap_fma(dr, db, crb); cgb = hap_fma(dg, db, cgb); cbb = hap_fma(db, db, cbb);
        }
    }
    float ar, ag, ab_, br, bg, bb;  // endpoints a (index 0 side) and b, in STORAGE units from here on
    const float var = crr + cgg + cbb;
    if (var < 0.5f) {
        // flat block: bracket the colour with its 5:6:5 grid neighbours so the 4 palette entries straddle
        // it; the final index pass picks the closest
        const float fr = mr * isr, fg = mg * isg, fb = mb * isb;
        ar = floorf(fr * (31.0f / 255.0f)) * (255.0f / 31.0f); br = ceilf(fr * (31.0f / 255.0f)) * (255.0f / 31.0f);
        ag = floorf(fg * (63.0f / 255.0f)) * (255.0f / 63.0f); bg = ceilf(fg * (63.0f / 255.0f)) * (255.0f / 63.0f);
        ab_ = floorf(fb * (31.0f / 255.0f)) * (255.0f / 31.0f); bb = ceilf(fb * (31.0f / 255.0f)) * (255.0f / 31.0f);
    } else {
        // principal axis: power iteration from the covariance row with the largest diagonal
        float vr, vg, vb;
        if (crr >= cgg && crr >= cbb) { vr = crr; vg = crg; vb = crb; }
        else if (cgg >= cbb) { vr = crg; vg = cgg; vb = cgb; }
        else { vr = crb; vg = cgb; vb = cbb; }
#pragma unroll
        for (int it = 0; it < 4; it++) {
            float nr = HAS_B ? hap_fma(crr, vr, hap_fma(crg, vg, crb * vb)) : hap_fma(crr, vr, crg * vg);
            float ng = HAS_B ? hap_fma(crg, vr, hap_fma(cgg, vg, cgb * vb)) : hap_fma(crg, vr, cgg * vg);
            float nb = HAS_B ? hap_fma(crb, vr, hap_fma(cgb, vg, cbb * vb)) : 0.0f;
            float m = fmaxf(fabsf(nr), fmaxf(fabsf(ng), fabsf(nb)));
            float inv = 1.0f / m;
            vr = nr * inv; vg = ng * inv; vb = nb * inv;
        }
        // extent along the axis -> first endpoints (metric space)
        float tmin = 1e30f, tmax = -1e30f;
        const float vv = HAS_B ? hap_fma(vr, vr, hap_fma(vg, vg, vb * vb)) : hap_fma(vr, vr, vg * vg);
        const float ivv = 1.0f / vv;
#pragma unroll
        for (int t = 0; t < 16; t++) {
            float d = (HAS_B ? hap_fma(r[t] - mr, vr, hap_fma(g[t] - mg, vg, (b[t] - mb) * vb)) : hap_fma(r[t] - mr, vr, (g[t] - mg) * vg)) * ivv;
            tmin = fminf(tmin, d);
            tmax = fmaxf(tmax, d);
        }
        // STARTS = 9 (the RGB formats): the Lloyd iteration only finds the local optimum next to its start, so it is
        // run from 3 x 3 placements of the two ends and the start with the lowest error AFTER grid snapping wins --
        // a cheap stand-in for cluster fit's search over all 969 ordered partitions (measured: +0.03 dB against it,
        // where a single start was -0.39 dB).
        FitSums S;
        bool have = false;
        float mar = 0.f, mag = 0.f, mab = 0.f, mbr = 0.f, mbg = 0.f, mbb = 0.f;
        float best_e = 1e30f;
#pragma unroll 1
        for (int st = 0; st < STARTS; st++) {
            // starts: both ends of the extent moved independently by -1/5, +1/5 or +3/5 of its length (3 x 3)
            const float step = 0.2f * (tmax - tmin);
            const float lo_t = STARTS == 1 ? tmin : tmin + ((float)(st % 3) * 2.0f - 1.0f) * step;
            const float hi_t = STARTS == 1 ? tmax : tmax - ((float)(st / 3) * 2.0f - 1.0f) * step;
            float car = hap_fma(vr, lo_t, mr), cag = hap_fma(vg, lo_t, mg), cab = hap_fma(vb, lo_t, mb);
            float cbr = hap_fma(vr, hi_t, mr), cbg = hap_fma(vg, hi_t, mg), cbb = hap_fma(vb, hi_t, mb);
            FitSums C;
            bool chave = false;
#pragma unroll 1
            for (int it = 0; it < REFINE; it++) {
                FitSums N;
                if (!cluster_sums<HAS_B>(r, g, b, car, cag, cab, cbr, cbg, cbb, N)) break;
                C = N;
                chave = true;
                float idet = 1.0f / (C.a2 * C.b2 - C.ab * C.ab);
                car = (C.axr * C.b2 - C.bxr * C.ab) * idet; cbr = (C.bxr * C.a2 - C.axr * C.ab) * idet;
                cag = (C.axg * C.b2 - C.bxg * C.ab) * idet; cbg = (C.bxg * C.a2 - C.axg * C.ab) * idet;
                if (HAS_B) { cab = (C.axb * C.b2 - C.bxb * C.ab) * idet; cbb = (C.bxb * C.a2 - C.axb * C.ab) * idet; }
            }
            if (STARTS == 1) {
                S = C; have = chave;
                mar = car; mag = cag; mab = cab; mbr = cbr; mbg = cbg; mbb = cbb;
            } else if (chave) {
                // error of this start after snapping (storage units; the metric is 1,1,1 for the RGB formats)
                float tr = fminf(fmaxf(car * isr, 0.f), 255.f), ur = fminf(fmaxf(cbr * isr, 0.f), 255.f);
                float tg = fminf(fmaxf(cag * isg, 0.f), 255.f), ug = fminf(fmaxf(cbg * isg, 0.f), 255.f);
                float tb = fminf(fmaxf(cab * isb, 0.f), 255.f), ub = fminf(fmaxf(cbb * isb, 0.f), 255.f);
                float e = snap_channel(tr, ur, C.a2, C.b2, C.ab, C.axr * isr, C.bxr * isr, 31.0f) * (sr * sr);
                e += snap_channel(tg, ug, C.a2, C.b2, C.ab, C.axg * isg, C.bxg * isg, 63.0f) * (sg * sg);
                if (HAS_B) e += snap_channel(tb, ub, C.a2, C.b2, C.ab, C.axb * isb, C.bxb * isb, 31.0f) * (sb * sb);
                if (e < best_e) {
                    best_e = e; S = C; have = true;
                    mar = car; mag = cag; mab = cab; mbr = cbr; mbg = cbg; mbb = cbb;
                }
            }
        }
        ar = fminf(fmaxf(mar * isr, 0.f), 255.f); br = fminf(fmaxf(mbr * isr, 0.f), 255.f);
        ag = fminf(fmaxf(mag * isg, 0.f), 255.f); bg = fminf(fmaxf(mbg * isg, 0.f), 255.f);
        ab_ = fminf(fmaxf(mab * isb, 0.f), 255.f); bb = fminf(fmaxf(mbb * isb, 0.f), 255.f);
        // Grid snapping: for fixed clusters the squared error is separable per channel,
        //   E(a,b) = a^2 A2 + b^2 B2 + 2ab AB - 2a AX - 2b BX,
        // so each channel tries floor/ceil of both endpoints on its 5- or 6-bit grid (4 candidates).
        if (have) {
            snap_channel(ar, br, S.a2, S.b2, S.ab, S.axr * isr, S.bxr * isr, 31.0f);
            snap_channel(ag, bg, S.a2, S.b2, S.ab, S.axg * isg, S.bxg * isg, 63.0f);
            if (HAS_B) snap_channel(ab_, bb, S.a2, S.b2, S.ab, S.axb * isb, S.bxb * isb, 31.0f);
#pragma unroll 1
            for (int it = 0; it < RESNAP; it++) {
                // Lloyd on the quantised problem: re-cluster against the snapped segment, re-solve, re-snap
                FitSums N;
                if (!cluster_sums<HAS_B>(r, g, b, ar * sr, ag * sg, ab_ * sb, br * sr, bg * sg, bb * sb, N)) break;
                float idet = 1.0f / (N.a2 * N.b2 - N.ab * N.ab);
                float car = fminf(fmaxf((N.axr * N.b2 - N.bxr * N.ab) * idet * isr, 0.f), 255.f);
                float cbr = fminf(fmaxf((N.bxr * N.a2 - N.axr * N.ab) * idet * isr, 0.f), 255.f);
                float cag = fminf(fmaxf((N.axg * N.b2 - N.bxg * N.ab) * idet * isg, 0.f), 255.f);
                float cbg = fminf(fmaxf((N.bxg * N.a2 - N.axg * N.ab) * idet * isg, 0.f), 255.f);
                float cab = fminf(fmaxf((N.axb * N.b2 - N.bxb * N.ab) * idet * isb, 0.f), 255.f);
                float cbb2 = fminf(fmaxf((N.bxb * N.a2 - N.axb * N.ab) * idet * isb, 0.f), 255.f);
                snap_channel(car, cbr, N.a2, N.b2, N.ab, N.axr * isr, N.bxr * isr, 31.0f);
                snap_channel(cag, cbg, N.a2, N.b2, N.ab, N.axg * isg, N.bxg * isg, 63.0f);
                snap_channel(cab, cbb2, N.a2, N.b2, N.ab, N.axb * isb, N.bxb * isb, 31.0f);
                ar = car; br = cbr; ag = cag; bg = cbg; ab_ = cab; bb = cbb2;
            }
        }
    }
    // 5:6:5
    uint32_t a5r = (uint32_t)hap_clampi((int)floorf(hap_fma(ar, 31.0f / 255.0f, 0.5f)), 0, 31);
    uint32_t a6g = (uint32_t)hap_clampi((int)floorf(hap_fma(ag, 63.0f / 255.0f, 0.5f)), 0, 63);
    uint32_t a5b = (uint32_t)hap_clampi((int)floorf(hap_fma(ab_, 31.0f / 255.0f, 0.5f)), 0, 31);
    uint32_t b5r = (uint32_t)hap_clampi((int)floorf(hap_fma(br, 31.0f / 255.0f, 0.5f)), 0, 31);
    uint32_t b6g = (uint32_t)hap_clampi((int)floorf(hap_fma(bg, 63.0f / 255.0f, 0.5f)), 0, 63);
    uint32_t b5b = (uint32_t)hap_clampi((int)floorf(hap_fma(bb, 31.0f / 255.0f, 0.5f)), 0, 31);
    if (fixed_blue5 >= 0) a5b = b5b = (uint32_t)fixed_blue5;  // scaled YCoCg: the scale code must survive exactly
    uint32_t c0 = (a5r << 11) | (a6g << 5) | a5b, c1 = (b5r << 11) | (b6g << 5) | b5b;
    Block8 out;
    if (c0 == c1) {
        out.lo = c0 | (c1 << 16);
        out.hi = 0;  // index 0 = c0 in either mode
        return out;
    }
    if (c0 < c1) {
        uint32_t tmp;
        tmp = c0; c0 = c1; c1 = tmp;
        tmp = a5r; a5r = b5r; b5r = tmp;
        tmp = a6g; a6g = b6g; b6g = tmp;
        tmp = a5b; a5b = b5b; b5b = tmp;
    }
    // decoder palette ends (c0 = first endpoint) in storage units, then in metric space for the comparisons
    const float e0r = (float)expand5(a5r), e0g = (float)expand6(a6g), e0b = (float)expand5(a5b);
    const float e1r = (float)expand5(b5r), e1g = (float)expand6(b6g), e1b = (float)expand5(b5b);
    const float p0r = e0r * sr, p0g = e0g * sg, p0b = e0b * sb, p1r = e1r * sr, p1g = e1g * sg, p1b = e1b * sb;
    uint32_t bits = 0;
    if (EXACT) {
        // decoder palette (truncating thirds), DXT numbering 0 = c0, 1 = c1, 2, 3
        const float q2r = floorf((2.0f * e0r + e1r) * (1.0f / 3.0f) + 0.01f) * sr, q3r = floorf((e0r + 2.0f * e1r) * (1.0f / 3.0f) + 0.01f) * sr;
        const float q2g = floorf((2.0f * e0g + e1g) * (1.0f / 3.0f) + 0.01f) * sg, q3g = floorf((e0g + 2.0f * e1g) * (1.0f / 3.0f) + 0.01f) * sg;
        const float q2b = floorf((2.0f * e0b + e1b) * (1.0f / 3.0f) + 0.01f) * sb, q3b = floorf((e0b + 2.0f * e1b) * (1.0f / 3.0f) + 0.01f) * sb;
#pragma unroll
        for (int t = 0; t < 16; t++) {
            float d0 = hap_fma(r[t] - p0r, r[t] - p0r, (g[t] - p0g) * (g[t] - p0g));
            float d1 = hap_fma(r[t] - p1r, r[t] - p1r, (g[t] - p1g) * (g[t] - p1g));
            float d2 = hap_fma(r[t] - q2r, r[t] - q2r, (g[t] - q2g) * (g[t] - q2g));
            float d3 = hap_fma(r[t] - q3r, r[t] - q3r, (g[t] - q3g) * (g[t] - q3g));
            if (HAS_B) {
                d0 = hap_fma(b[t] - p0b, b[t] - p0b, d0); d1 = hap_fma(b[t] - p1b, b[t] - p1b, d1);
                d2 = hap_fma(b[t] - q2b, b[t] - q2b, d2); d3 = hap_fma(b[t] - q3b, b[t] - q3b, d3);
            }
            uint32_t i01 = d1 < d0 ? 1u : 0u, i23 = d3 < d2 ? 3u : 2u;
            uint32_t idx = fminf(d2, d3) < fminf(d0, d1) ? i23 : i01;
            bits |= idx << (2 * t);
        }
    } else {
        const float er = p1r - p0r, eg = p1g - p0g, eb = p1b - p0b;
        const float ee = hap_fma(er, er, hap_fma(eg, eg, eb * eb));
        const float sc = 3.0f / ee;
#pragma unroll
        for (int t = 0; t < 16; t++) {
            float s = hap_fma(r[t] - p0r, er, hap_fma(g[t] - p0g, eg, (b[t] - p0b) * eb)) * sc;
            int q = hap_clampi((int)floorf(s + 0.5f), 0, 3);  // 0 = c0 ... 3 = c1 along the segment
            // DXT numbering: 0 = c0, 1 = c1, 2 = (2c0+c1)/3, 3 = (c0+2c1)/3
            uint32_t idx = q == 0 ? 0u : q == 3 ? 1u : (uint32_t)(q + 1);
            bits |= idx << (2 * t);
        }
    }
    out.lo = c0 | (c1 << 16);
    out.hi = bits;
    return out;
}

// ---- scaled YCoCg (van Waveren & Castano 2007) ----------------------------------------------------
// Per block: co = (R-B)/2, cg = (-R+2G-B)/4 kept as exact half/quarter integers; scale = largest of
// {4,2,1} with |co*scale|,|cg*scale| <= 127; stored texel (Co', Cg', (scale-1)*8, Y).
// RGB error of a chroma error (dCo, dCg): (dCo-dCg)^2 + dCg^2 + (dCo+dCg)^2 = 2 dCo^2 + 3 dCg^2
#ifndef HAP_YCOCG_FIT
#define HAP_YCOCG_FIT 1, 0, true
#endif
constexpr float kYCoCgMetricCo = 1.41421356f, kYCoCgMetricCg = 1.73205081f;

// One channel of the endpoint pair onto its 5- or 6-bit grid.  For fixed clusters the squared error separates per
// channel, E(a,b) = a^2 A2 + b^2 B2 + 2ab AB - 2a AX - 2b BX, so floor/ceil of both ends are tried (4 candidates)
// with the values the DECODER expands them to.  a, b: storage units 0..255; the sums are taken about `off`.
HAP_HD void snap_pair(float a, float b, float off, float A2, float B2, float AB, float AX, float BX, float levels,
                      uint32_t &ga_out, uint32_t &gb_out)
{
    const float to_grid = levels * (1.0f / 255.0f), from_grid = 255.0f / levels;
    const float ga0 = floorf(a * to_grid), gb0 = floorf(b * to_grid);
    const float ga1 = fminf(ga0 + 1.0f, levels), gb1 = fminf(gb0 + 1.0f, levels);
    // the decoder expands by bit replication, which equals round(g * 255 / levels) for 5 and 6 bits
    const float ca0 = floorf(hap_fma(ga0, from_grid, 0.5f)) - off, ca1 = floorf(hap_fma(ga1, from_grid, 0.5f)) - off;
    const float cb0 = floorf(hap_fma(gb0, from_grid, 0.5f)) - off, cb1 = floorf(hap_fma(gb1, from_grid, 0.5f)) - off;
    const float m2AX = -2.0f * AX, m2BX = -2.0f * BX, AB2 = 2.0f * AB;
    const float ua0 = ca0 * hap_fma(ca0, A2, m2AX), ua1 = ca1 * hap_fma(ca1, A2, m2AX);
    const float ub0 = cb0 * hap_fma(cb0, B2, m2BX), ub1 = cb1 * hap_fma(cb1, B2, m2BX);
    const float e00 = hap_fma(AB2 * ca0, cb0, ua0 + ub0), e10 = hap_fma(AB2 * ca1, cb0, ua1 + ub0);
    const float e01 = hap_fma(AB2 * ca0, cb1, ua0 + ub1), e11 = hap_fma(AB2 * ca1, cb1, ua1 + ub1);
    float best = e00, ga = ga0, gb = gb0;
    if (e10 < best) { best = e10; ga = ga1; gb = gb0; }
    if (e01 < best) { best = e01; ga = ga0; gb = gb1; }
    if (e11 < best) { ga = ga1; gb = gb1; }
    ga_out = (uint32_t)(int)ga;
    gb_out = (uint32_t)(int)gb;
}

// The chroma half of a scaled-YCoCg block: 16 (Co, Cg) pairs -> BC1 colour block (R' = Co', G' = Cg', B' = scale
// code).  co2 = R - B (half units), cg4 = -R + 2G - B (quarter units), with their max / min over the block.
//
// Everything is a fit in the 2-D plane of (Co', Cg') weighted by the metric above, done on the UNROUNDED chroma
// (the stored texel would be its rounding; the fit is against what the decoder should reproduce) and about the
// centre of the block's bounding box, which keeps the fp32 moment sums exact enough for a one-pass covariance:
//   moments -> 2x2 covariance -> principal axis in closed form -> extent along it -> 4 clusters by projection
//   -> the 2x2 least-squares system for the two endpoints given those clusters (the normal equations cluster
//   fit solves, for the partition the axis implies) -> per-channel grid search (snap_pair) -> indices by
//   projection onto the decoder's palette segment.
// Cluster sums are kept in terms of q = 0..3 (the cluster number): with beta = q/3, alpha = 1 - beta all nine
// sums of the normal equations follow from sum(q), sum(q^2), sum(q x), sum(q y) and the plain moments.
HAP_HD Block8 encode_ycocg_chroma(const int co2[16], const int cg4[16], int co_hi, int co_lo, int cg_hi, int cg_lo)
{
    const int m2 = co_hi > -co_lo ? co_hi : -co_lo, m4 = cg_hi > -cg_lo ? cg_hi : -cg_lo;
    int scale = 1;
    if (m2 * 4 <= 254 && m4 * 4 <= 508) scale = 4;
    else if (m2 * 2 <= 254 && m4 * 2 <= 508) scale = 2;
    const uint32_t blue5 = (uint32_t)(scale - 1);  // 5-bit code 0, 1 or 3: expands to B' = 0, 8, 24
    const float fs = (float)scale;
    // x = (Co' - offx) * metric, Co' = co2 * scale / 2 + 128
    const float cx = 0.5f * (float)(co_hi + co_lo), cy = 0.5f * (float)(cg_hi + cg_lo);
    const float kx = fs * (0.5f * kYCoCgMetricCo), ky = fs * (0.25f * kYCoCgMetricCg);
    const float bx0 = -cx * kx, by0 = -cy * ky;
    const float offx = hap_fma(cx, 0.5f * fs, 128.0f), offy = hap_fma(cy, 0.25f * fs, 128.0f);
    const float imx = 1.0f / kYCoCgMetricCo, imy = 1.0f / kYCoCgMetricCg;
    float x[16], y[16];
    float Sx = 0.f, Sy = 0.f, Sxx = 0.f, Sxy = 0.f, Syy = 0.f;
#pragma unroll
    for (int t = 0; t < 16; t++) {
        x[t] = hap_fma((float)co2[t], kx, bx0);
        y[t] = hap_fma((float)cg4[t], ky, by0);
        Sx += x[t]; Sy += y[t];
        Sxx = hap_fma(x[t], x[t], Sxx); Sxy = hap_fma(x[t], y[t], Sxy); Syy = hap_fma(y[t], y[t], Syy);
    }
    const float cxx = hap_fma(-0.0625f * Sx, Sx, Sxx), cxy = hap_fma(-0.0625f * Sx, Sy, Sxy), cyy = hap_fma(-0.0625f * Sy, Sy, Syy);

    uint32_t a5r, a6g, b5r, b6g;
    bool fitted = false;
    if (cxx + cyy >= 0.5f) {
        // principal axis of [[cxx, cxy], [cxy, cyy]]: eigenvector of the larger eigenvalue, closed form
        const float hd = 0.5f * (cxx - cyy);
        const float rad = sqrtf(hap_fma(hd, hd, cxy * cxy));
        float vx = hd >= 0.f ? hd + rad : cxy, vy = hd >= 0.f ? cxy : rad - hd;
        const float vv0 = hap_fma(vx, vx, vy * vy);
        if (vv0 < 1e-12f) { vx = 1.0f; vy = 0.0f; }  // isotropic spread: any axis
        // extent along the axis
        float d[16];
#pragma unroll
        for (int t = 0; t < 16; t++) d[t] = hap_fma(x[t], vx, y[t] * vy);
        float tmin = d[0], tmax = d[0];
#pragma unroll
        for (int t = 1; t < 15; t += 2) { tmin = hap_fmin3(tmin, d[t], d[t + 1]); tmax = hap_fmax3(tmax, d[t], d[t + 1]); }
        tmin = fminf(tmin, d[15]); tmax = fmaxf(tmax, d[15]);
        const float ext = tmax - tmin;
        if (ext > 1e-6f) {
            // clusters: q = round(3 (d - tmin) / ext) is 0..3 by construction, no clamp
            const float sc = 3.0f * (1.0f / ext);
            const float c0 = -tmin * sc;
            float Sq = 0.f, Sqq = 0.f, Sqx = 0.f, Sqy = 0.f;
#pragma unroll
            for (int t = 0; t < 16; t++) {
                const float q = rintf(hap_fma(d[t], sc, c0));
                Sq += q; Sqq = hap_fma(q, q, Sqq); Sqx = hap_fma(q, x[t], Sqx); Sqy = hap_fma(q, y[t], Sqy);
            }
            const float B2 = Sqq * (1.0f / 9.0f), AB = hap_fma(Sq, 1.0f / 3.0f, -B2), A2 = 16.0f - hap_fma(Sq, 2.0f / 3.0f, -B2);
            const float BXx = Sqx * (1.0f / 3.0f), BXy = Sqy * (1.0f / 3.0f), AXx = Sx - BXx, AXy = Sy - BXy;
            const float det = hap_fma(A2, B2, -(AB * AB));
            float eax, eay, ebx, eby;  // endpoints, metric space about the centre
            if (det >= 1e-4f) {
                const float idet = 1.0f / det;
                eax = hap_fma(AXx, B2, -(BXx * AB)) * idet; ebx = hap_fma(BXx, A2, -(AXx * AB)) * idet;
                eay = hap_fma(AXy, B2, -(BXy * AB)) * idet; eby = hap_fma(BXy, A2, -(AXy * AB)) * idet;
            } else {
                const float ivv = 1.0f / hap_fma(vx, vx, vy * vy);
                eax = vx * (tmin * ivv); eay = vy * (tmin * ivv); ebx = vx * (tmax * ivv); eby = vy * (tmax * ivv);
            }
            // storage units, then the grid
            const float ar = fminf(fmaxf(hap_fma(eax, imx, offx), 0.f), 255.f), br = fminf(fmaxf(hap_fma(ebx, imx, offx), 0.f), 255.f);
            const float ag = fminf(fmaxf(hap_fma(eay, imy, offy), 0.f), 255.f), bg = fminf(fmaxf(hap_fma(eby, imy, offy), 0.f), 255.f);
            if (det >= 1e-4f) {
                snap_pair(ar, br, offx, A2, B2, AB, AXx * imx, BXx * imx, 31.0f, a5r, b5r);
                a6g = (uint32_t)(int)floorf(hap_fma(ag, 63.0f / 255.0f, 0.5f)); b6g = (uint32_t)(int)floorf(hap_fma(bg, 63.0f / 255.0f, 0.5f));
            } else {
                a5r = (uint32_t)(int)floorf(hap_fma(ar, 31.0f / 255.0f, 0.5f)); b5r = (uint32_t)(int)floorf(hap_fma(br, 31.0f / 255.0f, 0.5f));
                a6g = (uint32_t)(int)floorf(hap_fma(ag, 63.0f / 255.0f, 0.5f)); b6g = (uint32_t)(int)floorf(hap_fma(bg, 63.0f / 255.0f, 0.5f));
            }
            fitted = true;
        }
    }
    if (!fitted) {
        // (nearly) one chroma: bracket it with its grid neighbours so the 4 palette entries straddle it
        const float fr = fminf(fmaxf(hap_fma(Sx * 0.0625f, imx, offx), 0.f), 255.f), fg = fminf(fmaxf(hap_fma(Sy * 0.0625f, imy, offy), 0.f), 255.f);
        a5r = (uint32_t)(int)floorf(fr * (31.0f / 255.0f)); b5r = (uint32_t)(int)ceilf(fr * (31.0f / 255.0f));
        a6g = (uint32_t)(int)floorf(fg * (63.0f / 255.0f)); b6g = (uint32_t)(int)ceilf(fg * (63.0f / 255.0f));
    }
    uint32_t c0 = (a5r << 11) | (a6g << 5) | blue5, c1 = (b5r << 11) | (b6g << 5) | blue5;
    Block8 out;
    out.lo = c0 | (c1 << 16);
    out.hi = 0;  // index 0 = c0 in either mode
    if (c0 == c1) return out;
    if (c0 < c1) {
        uint32_t tmp;
        tmp = c0; c0 = c1; c1 = tmp;
        tmp = a5r; a5r = b5r; b5r = tmp;
        tmp = a6g; a6g = b6g; b6g = tmp;
        out.lo = c0 | (c1 << 16);
    }
    // indices: position along the decoder's palette segment P0 -> P1, saturated to its ends, rounded to thirds.
    // The rounded level k = 0..3 sits in the low mantissa bits of (3w + magic); all sixteen are summed into one word
    // as k << 2t (the magic's own bits add up to a constant that is taken off once).
    const float p0x = ((float)expand5(a5r) - offx) * kYCoCgMetricCo, p0y = ((float)expand6(a6g) - offy) * kYCoCgMetricCg;
    const float p1x = ((float)expand5(b5r) - offx) * kYCoCgMetricCo, p1y = ((float)expand6(b6g) - offy) * kYCoCgMetricCg;
    const float ex = p1x - p0x, ey = p1y - p0y;
    const float iee = 1.0f / hap_fma(ex, ex, ey * ey);
    const float wx = ex * iee, wy = ey * iee, w0 = -hap_fma(p0x, wx, p0y * wy);
    uint32_t acc = 0;
#pragma unroll
    for (int t = 0; t < 16; t++) {
        const float w = hap_sat(hap_fma(x[t], wx, hap_fma(y[t], wy, w0)));
        acc += hap_float_bits(hap_fma(w, 3.0f, kRoundMagic)) << (2 * t);
    }
    const uint32_t k = acc - kRoundMagicBits * 0x55555555u;
    // level along the segment -> DXT numbering (0 = c0, 1 = c1, 2, 3 in between): 0,1,2,3 -> 0,2,3,1
    out.hi = (((k ^ (k >> 1)) & 0x55555555u) << 1) | ((k >> 1) & 0x55555555u);
    return out;
}

// A block whose 16 texels are one colour (letterbox bars, graphics, clipped highlights -- common in real
// footage): no statistics needed.  Endpoints bracket the colour on the 5:6:5 grid, the single index is the
// nearest of the four decoder palette entries under the metric (wr,wg,wb = squared channel weights).
HAP_HD Block8 encode_flat_colour(int R, int G, int B, int fixed_blue5, float wr, float wg, float wb)
{
    uint32_t a5r = (uint32_t)(R * 31) / 255u, a6g = (uint32_t)(G * 63) / 255u, a5b = (uint32_t)(B * 31) / 255u;
    uint32_t b5r = a5r + (expand5(a5r) != (uint32_t)R && a5r < 31 ? 1u : 0u);
    uint32_t b6g = a6g + (expand6(a6g) != (uint32_t)G && a6g < 63 ? 1u : 0u);
    uint32_t b5b = a5b + (expand5(a5b) != (uint32_t)B && a5b < 31 ? 1u : 0u);
    if (fixed_blue5 >= 0) a5b = b5b = (uint32_t)fixed_blue5;
    // c0 > c1 required for 4-colour mode: the "ceil" triple is the larger 565 word unless they are equal
    const uint32_t c0 = (b5r << 11) | (b6g << 5) | b5b, c1 = (a5r << 11) | (a6g << 5) | a5b;
    Block8 out;
    out.lo = c0 | (c1 << 16);
    out.hi = 0;
    if (c0 == c1) return out;
    if (c0 < c1) {
        // cannot happen (each field of c0 is >= the field of c1), kept for safety: index 0 still decodes to c0
        return out;
    }
    const int p0[3] = {(int)expand5(b5r), (int)expand6(b6g), (int)expand5(b5b)};
    const int p1[3] = {(int)expand5(a5r), (int)expand6(a6g), (int)expand5(a5b)};
    const int px3[3] = {R, G, B};
    const float wv[3] = {wr, wg, wb};
    float best = 1e30f;
    uint32_t bi = 0;
#pragma unroll
    for (uint32_t k = 0; k < 4; k++) {
        float e = 0.f;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            int v = k == 0 ? p0[c] : k == 1 ? p1[c] : k == 2 ? (2 * p0[c] + p1[c]) / 3 : (p0[c] + 2 * p1[c]) / 3;
            float dlt = (float)(px3[c] - v);
            e = hap_fma(dlt * wv[c], dlt, e);
        }
        if (e < best) { best = e; bi = k; }
    }
    out.hi = bi * 0x55555555u;
    return out;
}

// ---- whole-block encoders: px = 16 RGBA8 texels, row-major inside the block, little-endian ------
HAP_HD bool block_is_flat_rgb(const uint32_t px[16])
{
    uint32_t diff = 0;
#pragma unroll
    for (int t = 1; t < 16; t++) diff |= px[t] ^ px[0];
    return (diff & 0x00FFFFFFu) == 0;
}

HAP_HD Block8 encode_dxt1(const uint32_t px[16])
{
    if (block_is_flat_rgb(px))
        return encode_flat_colour((int)(px[0] & 0xFF), (int)((px[0] >> 8) & 0xFF), (int)((px[0] >> 16) & 0xFF), -1, 1.f, 1.f, 1.f);
    float r[16], g[16], b[16];
#pragma unroll
    for (int t = 0; t < 16; t++) {
        r[t] = (float)(px[t] & 0xFF); g[t] = (float)((px[t] >> 8) & 0xFF); b[t] = (float)((px[t] >> 16) & 0xFF);
    }
#ifndef HAP_RGB_FIT
#define HAP_RGB_FIT 1, 1, true, true, 9
#endif
    return encode_colour_block<HAP_RGB_FIT>(r, g, b);
}

HAP_HD Block8 encode_rgtc1_alpha(const uint32_t px[16])
{
    int a7[16];
#pragma unroll
    for (int t = 0; t < 16; t++) a7[t] = hap_dp4a_us(px[t], 0x07000000u, 0);  // 7 * alpha
    int hi = a7[0], lo = a7[0];
#pragma unroll
    for (int t = 1; t < 15; t += 2) { hi = hap_max3(hi, a7[t], a7[t + 1]); lo = hap_min3(lo, a7[t], a7[t + 1]); }
    hi = hap_max3(hi, a7[15], a7[15]); lo = hap_min3(lo, a7[15], a7[15]);
    return encode_bc4_scaled<7>(a7, hi, lo);
}

HAP_HD void encode_dxt5(const uint32_t px[16], Block8 &alpha, Block8 &colour)
{
    alpha = encode_rgtc1_alpha(px);
    colour = encode_dxt1(px);
}

HAP_HD void encode_ycocg_dxt5(const uint32_t px[16], Block8 &alpha, Block8 &colour)
{
    if (block_is_flat_rgb(px)) {
        const int R = (int)(px[0] & 0xFF), G = (int)((px[0] >> 8) & 0xFF), B = (int)((px[0] >> 16) & 0xFF);
        const int co2 = R - B, cg4 = -R + 2 * G - B;
        const int a2 = co2 < 0 ? -co2 : co2, a4 = cg4 < 0 ? -cg4 : cg4;
        const int scale = (a2 * 4 <= 254 && a4 * 4 <= 508) ? 4 : (a2 * 2 <= 254 && a4 * 2 <= 508) ? 2 : 1;
        const int co = hap_clampi(((co2 * scale + 1) >> 1) + 128, 0, 255), cg = hap_clampi(((cg4 * scale + 2) >> 2) + 128, 0, 255);
        const uint32_t Y = (uint32_t)((R + 2 * G + B + 2) >> 2);
        alpha.lo = Y | (Y << 8);
        alpha.hi = 0;
        colour = encode_flat_colour(co, cg, (scale - 1) << 3, scale - 1, 2.0f, 3.0f, 1.0f);
        return;
    }
    // texel bytes are R,G,B,A: weights (1,0,-1,0), (-1,2,-1,0) and 7 * (1,2,1,0) as signed bytes, one DP4A each
    int co2[16], cg4[16], y28[16];
#pragma unroll
    for (int t = 0; t < 16; t++) {
        co2[t] = hap_dp4a_us(px[t], 0x00FF0001u, 0);
        cg4[t] = hap_dp4a_us(px[t], 0x00FF02FFu, 0);
        y28[t] = hap_dp4a_us(px[t], 0x00070E07u, 0);  // 7 (R + 2G + B) = 28 * luma
    }
    int co_hi = co2[0], co_lo = co2[0], cg_hi = cg4[0], cg_lo = cg4[0], y_hi = y28[0], y_lo = y28[0];
#pragma unroll
    for (int t = 1; t < 15; t += 2) {
        co_hi = hap_max3(co_hi, co2[t], co2[t + 1]); co_lo = hap_min3(co_lo, co2[t], co2[t + 1]);
        cg_hi = hap_max3(cg_hi, cg4[t], cg4[t + 1]); cg_lo = hap_min3(cg_lo, cg4[t], cg4[t + 1]);
        y_hi = hap_max3(y_hi, y28[t], y28[t + 1]); y_lo = hap_min3(y_lo, y28[t], y28[t + 1]);
    }
    co_hi = hap_max3(co_hi, co2[15], co2[15]); co_lo = hap_min3(co_lo, co2[15], co2[15]);
    cg_hi = hap_max3(cg_hi, cg4[15], cg4[15]); cg_lo = hap_min3(cg_lo, cg4[15], cg4[15]);
    y_hi = hap_max3(y_hi, y28[15], y28[15]); y_lo = hap_min3(y_lo, y28[15], y28[15]);
    alpha = encode_bc4_scaled<28>(y28, y_hi, y_lo);
    colour = encode_ycocg_chroma(co2, cg4, co_hi, co_lo, cg_hi, cg_lo);
}

}  // namespace hapb200
