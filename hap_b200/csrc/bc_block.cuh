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
// and the four 16-byte row pieces of neighbouring threads are contiguous, so the loads are coalesced as
// they are (bc_encode.cuh).
//
// The fits, all on the same pattern -- principal axis of the block's covariance, 4 clusters by projection,
// the 2x2 least-squares system for the two endpoints given those clusters (the normal equations cluster
// fit solves, for the partition the axis implies instead of all 969), endpoints onto the 5:6:5 grid by a
// small search with the decoder's expanded values, indices against the decoder's palette:
//   * scaled YCoCg chroma (encode_ycocg_chroma): 2-D, closed-form axis, one start, ~600 instructions;
//   * RGB (encode_rgb_block): 3-D, power iteration, 2 x 2 starts scored after grid snapping, one Lloyd
//     round on the winner, exact nearest-palette indices, ~2 600 instructions;
//   * BC4 (encode_bc4_scaled): endpoints = min / max, indices by one biased fixed-point rounding per texel.
// Instruction counts are the currency here: the kernels are bound by instruction issue, and every step was
// held to the PSNR bar on the host build of this very source before it went to the GPU.
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

// ---- endpoint refinement where the grid is coarse against the data ---------------------------------------------
// snap_pair scores grid candidates with the clusters of the unquantised fit held fixed.  That is right while the
// grid step is small against the block's extent; when the Co' endpoints lie less than one 5-bit cell (8.2 storage
// units) apart, putting them on the grid stretches or collapses the segment so much that texels change cluster, and the
// fixed-partition score picks the wrong candidate about half the time (-0.47 dB on slow colour ramps against the
// oracle's search over all partitions).  There, the four floor/ceil candidates of the Co' pair are scored by their TRUE
// error instead: every texel re-assigned to the nearest of the candidate's four palette points.
//   sum |x - p0 - r e|^2 = sum |x - p0|^2 + |e|^2 sum r (r - 2u),  u = (x - p0).e / |e|^2,  r = round(3 sat(u)) / 3:
// the first sum follows from the block moments, the second costs 7 instructions per texel.
// It doubles the cost of the blocks it touches (half of the blocks of a smooth 4K picture), so it is an encoder OPTION
// (HAPB200_OPTION_CHROMA_REFINE, include/hap_b200.h) and a template parameter here: off costs nothing.
HAP_HD float chroma_true_error(const float x[16], const float y[16], float Sx, float Sy, float Sxx, float Syy, float p0x, float p0y, float p1x, float p1y)
{
    const float ex = p1x - p0x, ey = p1y - p0y;
    const float ee = hap_fma(ex, ex, ey * ey);
    const float iee = ee > 1e-9f ? 1.0f / ee : 0.0f;      // both endpoints on one grid point: every texel decodes to p0
    const float wx = ex * iee, wy = ey * iee, w0 = -hap_fma(p0x, wx, p0y * wy);
    float acc = 0.f;
#pragma unroll
    for (int t = 0; t < 16; t++) {
        const float u = hap_fma(x[t], wx, hap_fma(y[t], wy, w0));
        const float r3 = hap_fma(hap_sat(u), 3.0f, kRoundMagic) - kRoundMagic;     // 3 r = 0, 1, 2, 3
        acc = hap_fma(r3, hap_fma(u, -6.0f, r3), acc);                            // 9 r (r - 2u)
    }
    const float P = hap_fma(16.0f, hap_fma(p0x, p0x, p0y * p0y), hap_fma(-2.0f * p0x, Sx, hap_fma(-2.0f * p0y, Sy, Sxx + Syy)));
    return hap_fma(ee * (1.0f / 9.0f), acc, P);
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
template <bool REFINE>
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
                a6g = (uint32_t)(int)floorf(hap_fma(ag, 63.0f / 255.0f, 0.5f)); b6g = (uint32_t)(int)floorf(hap_fma(bg, 63.0f / 255.0f, 0.5f));
                if (REFINE && fabsf(ar - br) < 255.0f / 31.0f) {
                    const float p0y = ((float)expand6(a6g) - offy) * kYCoCgMetricCg, p1y = ((float)expand6(b6g) - offy) * kYCoCgMetricCg;
                    const uint32_t fa = (uint32_t)(int)floorf(ar * (31.0f / 255.0f)), fb = (uint32_t)(int)floorf(br * (31.0f / 255.0f));
                    float best = 3.0e38f;
                    a5r = fa; b5r = fb;
#pragma unroll 1
                    for (uint32_t c = 0; c < 4; c++) {
                        const uint32_t ca = fa + (c & 1u) < 31u ? fa + (c & 1u) : 31u, cb = fb + (c >> 1) < 31u ? fb + (c >> 1) : 31u;
                        const float e = chroma_true_error(x, y, Sx, Sy, Sxx, Syy, ((float)expand5(ca) - offx) * kYCoCgMetricCo, p0y,
                                                          ((float)expand5(cb) - offx) * kYCoCgMetricCo, p1y);
                        if (e < best) { best = e; a5r = ca; b5r = cb; }
                    }
                } else {
                    snap_pair(ar, br, offx, A2, B2, AB, AXx * imx, BXx * imx, 31.0f, a5r, b5r);
                }
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

// ---- RGB colour block (DXT1 / the colour half of DXT5), the short way ----------------------------------------
// Principal axis of the covariance, then a small multi-start search along it that stands in for cluster fit's
// search over all 969 ordered partitions: N starts = placements of the two ends of the extent; per start 4 clusters
// by projection, the 2x2 least-squares system for the endpoints, endpoints onto the 5:6:5 grid (4 candidates per
// channel) and the error there; the best start gets one more Lloyd round on the quantised problem, then exact
// nearest-palette indices.  (A single start was measured at -0.39 dB against the cluster-fit oracle, this at
// -0.09 ... +0.03 dB.)  Arranged so that a start costs ~330 instructions:
//  * everything is done about the block mean (sum of the centred texels = 0, so alpha.x = -beta.x for every channel);
//  * texels are projected on the axis ONCE; a start only rescales the projection;
//  * cluster sums are over the cluster NUMBER q (sum q, sum q^2, sum q x): with beta = q/3 all normal-equation
//    terms follow;
//  * the grid search returns its error, which is the start's score.
#ifndef HAP_RGB_STARTS
#define HAP_RGB_STARTS 4
#endif
#ifndef HAP_RGB_RESNAP
#define HAP_RGB_RESNAP 1
#endif
struct RgbFit {
    float A2, B2, AB, BXr, BXg, BXb;     // normal-equation sums about the mean (AX = -BX)
};

// grid value (0..levels) nearest to storage value v (0..255), and what the decoder expands it to
HAP_HD float grid_round(float v, float levels) { return floorf(hap_fma(v, levels * (1.0f / 255.0f), 0.5f)); }
HAP_HD float grid_expand(float gq, float levels) { return floorf(hap_fma(gq, 255.0f / levels, 0.5f)); }

// error (up to the constant sum of squares) of endpoints ca, cb (about the mean) for one channel
HAP_HD float pair_error(float ca, float cb, float A2, float B2, float AB, float BX)
{
    // E = ca^2 A2 + cb^2 B2 + 2 ca cb AB - 2 ca AX - 2 cb BX with AX = -BX
    return hap_fma(ca, hap_fma(ca, A2, 2.0f * BX), hap_fma(cb, hap_fma(cb, B2, -2.0f * BX), (2.0f * AB) * ca * cb));
}

// 4-candidate search of one channel: floor/ceil of both ends on the grid; returns grid values
HAP_HD float snap_pair_centred(float a, float b, float mean, float A2, float B2, float AB, float BX, float levels, float &ga_out, float &gb_out)
{
    const float to_grid = levels * (1.0f / 255.0f);
    const float ga0 = floorf(a * to_grid), gb0 = floorf(b * to_grid);
    const float ga1 = fminf(ga0 + 1.0f, levels), gb1 = fminf(gb0 + 1.0f, levels);
    const float ca0 = grid_expand(ga0, levels) - mean, ca1 = grid_expand(ga1, levels) - mean;
    const float cb0 = grid_expand(gb0, levels) - mean, cb1 = grid_expand(gb1, levels) - mean;
    // E = ua(ca) + ub(cb) + 2 AB ca cb with ua(c) = c (c A2 + 2 BX), ub(c) = c (c B2 - 2 BX)  (AX = -BX)
    const float BX2 = 2.0f * BX, AB2 = 2.0f * AB;
    const float ua0 = ca0 * hap_fma(ca0, A2, BX2), ua1 = ca1 * hap_fma(ca1, A2, BX2);
    const float ub0 = cb0 * hap_fma(cb0, B2, -BX2), ub1 = cb1 * hap_fma(cb1, B2, -BX2);
    const float e00 = hap_fma(AB2 * ca0, cb0, ua0 + ub0), e10 = hap_fma(AB2 * ca1, cb0, ua1 + ub0);
    const float e01 = hap_fma(AB2 * ca0, cb1, ua0 + ub1), e11 = hap_fma(AB2 * ca1, cb1, ua1 + ub1);
    float best = e00, ga = ga0, gb = gb0;
    if (e10 < best) { best = e10; ga = ga1; gb = gb0; }
    if (e01 < best) { best = e01; ga = ga0; gb = gb1; }
    if (e11 < best) { best = e11; ga = ga1; gb = gb1; }
    ga_out = ga;
    gb_out = gb;
    return best;
}

// clusters along the segment [lo, hi] of the axis projection d[] -> normal-equation sums; false when degenerate
HAP_HD bool rgb_cluster_sums(const float d[16], const float x[16], const float y[16], const float z[16], float lo, float hi, RgbFit &F)
{
    const float ext = hi - lo;
    if (!(ext > 1e-6f)) return false;
    const float sc = 3.0f * (1.0f / ext), c0 = -lo * sc;
    float Sq = 0.f, Sqq = 0.f, Sqx = 0.f, Sqy = 0.f, Sqz = 0.f;
#pragma unroll
    for (int t = 0; t < 16; t++) {
        const float q = fminf(fmaxf(rintf(hap_fma(d[t], sc, c0)), 0.0f), 3.0f);
        Sq += q; Sqq = hap_fma(q, q, Sqq);
        Sqx = hap_fma(q, x[t], Sqx); Sqy = hap_fma(q, y[t], Sqy); Sqz = hap_fma(q, z[t], Sqz);
    }
    F.B2 = Sqq * (1.0f / 9.0f);
    F.AB = hap_fma(Sq, 1.0f / 3.0f, -F.B2);
    F.A2 = 16.0f - hap_fma(Sq, 2.0f / 3.0f, -F.B2);
    F.BXr = Sqx * (1.0f / 3.0f); F.BXg = Sqy * (1.0f / 3.0f); F.BXb = Sqz * (1.0f / 3.0f);
    return hap_fma(F.A2, F.B2, -(F.AB * F.AB)) >= 1e-4f;
}

// least-squares endpoints (about the mean) of one channel: AX = -BX
HAP_HD void rgb_solve(const RgbFit &F, float idet, float BX, float &a, float &b)
{
    a = -BX * (F.B2 + F.AB) * idet;
    b = BX * (F.A2 + F.AB) * idet;
}

HAP_HD Block8 encode_rgb_block(const uint32_t px[16])
{
    float x[16], y[16], z[16];
    float mr = 0.f, mg = 0.f, mb = 0.f;
#pragma unroll
    for (int t = 0; t < 16; t++) {
        x[t] = (float)(px[t] & 0xFF); y[t] = (float)((px[t] >> 8) & 0xFF); z[t] = (float)((px[t] >> 16) & 0xFF);
        mr += x[t]; mg += y[t]; mb += z[t];
    }
    mr *= 0.0625f; mg *= 0.0625f; mb *= 0.0625f;
    float crr = 0.f, crg = 0.f, crb = 0.f, cgg = 0.f, cgb = 0.f, cbb = 0.f;
#pragma unroll
    for (int t = 0; t < 16; t++) {
        x[t] -= mr; y[t] -= mg; z[t] -= mb;
        crr = hap_fma(x[t], x[t], crr); crg = hap_fma(x[t], y[t], crg); crb = hap_fma(x[t], z[t], crb);
        cgg = hap_fma(y[t], y[t], cgg); cgb = hap_fma(y[t], z[t], cgb); cbb = hap_fma(z[t], z[t], cbb);
    }
    float gar, gag, gab, gbr, gbg, gbb;   // endpoints on the 5:6:5 grid
    if (crr + cgg + cbb < 0.5f) {
        // flat block: bracket the colour with its grid neighbours so the 4 palette entries straddle it
        gar = floorf(mr * (31.0f / 255.0f)); gbr = ceilf(mr * (31.0f / 255.0f));
        gag = floorf(mg * (63.0f / 255.0f)); gbg = ceilf(mg * (63.0f / 255.0f));
        gab = floorf(mb * (31.0f / 255.0f)); gbb = ceilf(mb * (31.0f / 255.0f));
    } else {
        // principal axis: power iteration from the covariance row with the largest diagonal
        float vr, vg, vb;
        if (crr >= cgg && crr >= cbb) { vr = crr; vg = crg; vb = crb; }
        else if (cgg >= cbb) { vr = crg; vg = cgg; vb = cgb; }
        else { vr = crb; vg = cgb; vb = cbb; }
#pragma unroll
        for (int it = 0; it < 4; it++) {
            const float nr = hap_fma(crr, vr, hap_fma(crg, vg, crb * vb));
            const float ng = hap_fma(crg, vr, hap_fma(cgg, vg, cgb * vb));
            const float nb = hap_fma(crb, vr, hap_fma(cgb, vg, cbb * vb));
            const float inv = 1.0f / fmaxf(fabsf(nr), fmaxf(fabsf(ng), fabsf(nb)));
            vr = nr * inv; vg = ng * inv; vb = nb * inv;
        }
        // projection on the axis (in units where the endpoints are mean + v * t), once
        const float ivv = 1.0f / hap_fma(vr, vr, hap_fma(vg, vg, vb * vb));
        float d[16];
        float tmin = 1e30f, tmax = -1e30f;
#pragma unroll
        for (int t = 0; t < 16; t++) {
            d[t] = hap_fma(x[t], vr, hap_fma(y[t], vg, z[t] * vb)) * ivv;
            tmin = fminf(tmin, d[t]);
            tmax = fmaxf(tmax, d[t]);
        }
        // starts: both ends of the extent moved independently outwards or inwards by 1/5 of its length (2 x 2).
        // (A 3 x 3 grid that also tried 3/5 inwards was measured: its five extra starts win almost never,
        // +0.005 dB for more than twice the work.)
        const float step = 0.2f * (tmax - tmin);
        bool have = false;
        float best_e = 1e30f;
#pragma unroll 1
        for (int st = 0; st < HAP_RGB_STARTS; st++) {
            const float lo = HAP_RGB_STARTS == 1 ? tmin : tmin + ((float)(st & 1) * 2.0f - 1.0f) * step;
            const float hi = HAP_RGB_STARTS == 1 ? tmax : tmax - ((float)(st >> 1) * 2.0f - 1.0f) * step;
            RgbFit F;
            if (!rgb_cluster_sums(d, x, y, z, lo, hi, F)) continue;
            const float idet = 1.0f / hap_fma(F.A2, F.B2, -(F.AB * F.AB));
            float ar, br, ag, bg, ab, bb;
            rgb_solve(F, idet, F.BXr, ar, br); rgb_solve(F, idet, F.BXg, ag, bg); rgb_solve(F, idet, F.BXb, ab, bb);
            // storage units, clamped; scored with plain rounding to the grid
            ar = fminf(fmaxf(ar + mr, 0.f), 255.f); br = fminf(fmaxf(br + mr, 0.f), 255.f);
            ag = fminf(fmaxf(ag + mg, 0.f), 255.f); bg = fminf(fmaxf(bg + mg, 0.f), 255.f);
            ab = fminf(fmaxf(ab + mb, 0.f), 255.f); bb = fminf(fmaxf(bb + mb, 0.f), 255.f);
            // scored AFTER the endpoints are put on the grid (4 candidates per channel): scoring the unquantised or the
            // plainly rounded endpoints picks the wrong start often enough to cost 0.01 ... 0.06 dB
            float s_ar, s_br, s_ag, s_bg, s_ab, s_bb;
            const float e = snap_pair_centred(ar, br, mr, F.A2, F.B2, F.AB, F.BXr, 31.f, s_ar, s_br) +
                            snap_pair_centred(ag, bg, mg, F.A2, F.B2, F.AB, F.BXg, 63.f, s_ag, s_bg) +
                            snap_pair_centred(ab, bb, mb, F.A2, F.B2, F.AB, F.BXb, 31.f, s_ab, s_bb);
            if (e < best_e) {
                best_e = e; have = true;
                gar = s_ar; gag = s_ag; gab = s_ab; gbr = s_br; gbg = s_bg; gbb = s_bb;
            }
        }
        if (!have) {
            // no start had two usable clusters: the ends of the extent, rounded
            gar = grid_round(fminf(fmaxf(hap_fma(vr, tmin, mr), 0.f), 255.f), 31.f); gbr = grid_round(fminf(fmaxf(hap_fma(vr, tmax, mr), 0.f), 255.f), 31.f);
            gag = grid_round(fminf(fmaxf(hap_fma(vg, tmin, mg), 0.f), 255.f), 63.f); gbg = grid_round(fminf(fmaxf(hap_fma(vg, tmax, mg), 0.f), 255.f), 63.f);
            gab = grid_round(fminf(fmaxf(hap_fma(vb, tmin, mb), 0.f), 255.f), 31.f); gbb = grid_round(fminf(fmaxf(hap_fma(vb, tmax, mb), 0.f), 255.f), 31.f);
        } else {
            // Lloyd rounds on the quantised problem: re-cluster against the snapped segment, re-solve, re-snap
#pragma unroll 1
            for (int rs = 0; rs < HAP_RGB_RESNAP; rs++) {
                const float ear = grid_expand(gar, 31.f) - mr, eag = grid_expand(gag, 63.f) - mg, eab = grid_expand(gab, 31.f) - mb;
                const float ebr = grid_expand(gbr, 31.f) - mr, ebg = grid_expand(gbg, 63.f) - mg, ebb = grid_expand(gbb, 31.f) - mb;
                const float sr_ = ebr - ear, sg_ = ebg - eag, sb_ = ebb - eab;
                const float ss = hap_fma(sr_, sr_, hap_fma(sg_, sg_, sb_ * sb_));
                if (ss > 1e-6f) {
                    // projection of the texels on the snapped segment: p = ((x - ea) . s) / (s . s), in 0..1
                    const float iss = 1.0f / ss;
                    float p[16];
#pragma unroll
                    for (int t = 0; t < 16; t++) p[t] = hap_fma(x[t] - ear, sr_, hap_fma(y[t] - eag, sg_, (z[t] - eab) * sb_)) * iss;
                    RgbFit N;
                    if (rgb_cluster_sums(p, x, y, z, 0.0f, 1.0f, N)) {
                        const float idet = 1.0f / hap_fma(N.A2, N.B2, -(N.AB * N.AB));
                        float ar, br, ag, bg, ab, bb;
                        rgb_solve(N, idet, N.BXr, ar, br); rgb_solve(N, idet, N.BXg, ag, bg); rgb_solve(N, idet, N.BXb, ab, bb);
                        ar = fminf(fmaxf(ar + mr, 0.f), 255.f); br = fminf(fmaxf(br + mr, 0.f), 255.f);
                        ag = fminf(fmaxf(ag + mg, 0.f), 255.f); bg = fminf(fmaxf(bg + mg, 0.f), 255.f);
                        ab = fminf(fmaxf(ab + mb, 0.f), 255.f); bb = fminf(fmaxf(bb + mb, 0.f), 255.f);
                        snap_pair_centred(ar, br, mr, N.A2, N.B2, N.AB, N.BXr, 31.f, gar, gbr);
                        snap_pair_centred(ag, bg, mg, N.A2, N.B2, N.AB, N.BXg, 63.f, gag, gbg);
                        snap_pair_centred(ab, bb, mb, N.A2, N.B2, N.AB, N.BXb, 31.f, gab, gbb);
                    }
                }
            }
        }
    }
    uint32_t a5r = (uint32_t)(int)gar, a6g = (uint32_t)(int)gag, a5b = (uint32_t)(int)gab;
    uint32_t b5r = (uint32_t)(int)gbr, b6g = (uint32_t)(int)gbg, b5b = (uint32_t)(int)gbb;
    uint32_t c0 = (a5r << 11) | (a6g << 5) | a5b, c1 = (b5r << 11) | (b6g << 5) | b5b;
    Block8 out;
    out.lo = c0 | (c1 << 16);
    out.hi = 0;  // index 0 = c0 in either mode
    if (c0 == c1) return out;
    if (c0 < c1) {
        uint32_t tmp;
        tmp = c0; c0 = c1; c1 = tmp;
        tmp = a5r; a5r = b5r; b5r = tmp;
        tmp = a6g; a6g = b6g; b6g = tmp;
        tmp = a5b; a5b = b5b; b5b = tmp;
        out.lo = c0 | (c1 << 16);
    }
    // exact indices: nearest of the decoder's four palette colours (truncating thirds), DXT numbering 0 = c0, 1 = c1, 2, 3
    const float e0r = (float)expand5(a5r), e0g = (float)expand6(a6g), e0b = (float)expand5(a5b);
    const float e1r = (float)expand5(b5r), e1g = (float)expand6(b6g), e1b = (float)expand5(b5b);
    const float p0r = e0r - mr, p0g = e0g - mg, p0b = e0b - mb, p1r = e1r - mr, p1g = e1g - mg, p1b = e1b - mb;
    const float q2r = floorf((2.0f * e0r + e1r) * (1.0f / 3.0f) + 0.01f) - mr, q3r = floorf((e0r + 2.0f * e1r) * (1.0f / 3.0f) + 0.01f) - mr;
    const float q2g = floorf((2.0f * e0g + e1g) * (1.0f / 3.0f) + 0.01f) - mg, q3g = floorf((e0g + 2.0f * e1g) * (1.0f / 3.0f) + 0.01f) - mg;
    const float q2b = floorf((2.0f * e0b + e1b) * (1.0f / 3.0f) + 0.01f) - mb, q3b = floorf((e0b + 2.0f * e1b) * (1.0f / 3.0f) + 0.01f) - mb;
    // |x - p|^2 = |x|^2 - 2 x.p + |p|^2: the |x|^2 term is common, so compare  |p|^2 - 2 x.p
    const float n0 = hap_fma(p0r, p0r, hap_fma(p0g, p0g, p0b * p0b)), n1 = hap_fma(p1r, p1r, hap_fma(p1g, p1g, p1b * p1b));
    const float n2 = hap_fma(q2r, q2r, hap_fma(q2g, q2g, q2b * q2b)), n3 = hap_fma(q3r, q3r, hap_fma(q3g, q3g, q3b * q3b));
    uint32_t bits = 0;
#pragma unroll
    for (int t = 0; t < 16; t++) {
        const float d0 = hap_fma(-2.0f * x[t], p0r, hap_fma(-2.0f * y[t], p0g, hap_fma(-2.0f * z[t], p0b, n0)));
        const float d1 = hap_fma(-2.0f * x[t], p1r, hap_fma(-2.0f * y[t], p1g, hap_fma(-2.0f * z[t], p1b, n1)));
        const float d2 = hap_fma(-2.0f * x[t], q2r, hap_fma(-2.0f * y[t], q2g, hap_fma(-2.0f * z[t], q2b, n2)));
        const float d3 = hap_fma(-2.0f * x[t], q3r, hap_fma(-2.0f * y[t], q3g, hap_fma(-2.0f * z[t], q3b, n3)));
        const uint32_t i01 = d1 < d0 ? 1u : 0u, i23 = d3 < d2 ? 3u : 2u;
        const uint32_t idx = fminf(d2, d3) < fminf(d0, d1) ? i23 : i01;
        bits |= idx << (2 * t);
    }
    out.hi = bits;
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
    return encode_rgb_block(px);
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

template <bool REFINE = false>
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
    colour = encode_ycocg_chroma<REFINE>(co2, cg4, co_hi, co_lo, cg_hi, cg_lo);
}

}  // namespace hapb200
