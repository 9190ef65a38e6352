/*
 * oracle/bc_oracle.c -- TEST INFRASTRUCTURE ONLY. See bc_oracle.h (PARITY UNPINNED).
 */
#include "bc_oracle.h"
#include <math.h>
#include <string.h>

/* ---- decoders (S3TC spec, truncating integer interpolation as in Pillow's bcn decoder) --------- */

static void expand565(unsigned c, int *rgb)
{
    int r = (c >> 11) & 31, g = (c >> 5) & 63, b = c & 31;
    rgb[0] = (r << 3) | (r >> 2);
    rgb[1] = (g << 2) | (g >> 4);
    rgb[2] = (b << 3) | (b >> 2);
}

/* palette[4][4] RGBA; force4 = BC3 colour block (always 4-colour mode) */
static void bc1_palette(const uint8_t *blk, int force4, int pal[4][4])
{
    unsigned c0 = blk[0] | (blk[1] << 8), c1 = blk[2] | (blk[3] << 8);
    expand565(c0, pal[0]);
    expand565(c1, pal[1]);
    pal[0][3] = pal[1][3] = pal[2][3] = pal[3][3] = 255;
    if (force4 || c0 > c1) {
        for (int k = 0; k < 3; k++) {
            pal[2][k] = (2 * pal[0][k] + pal[1][k]) / 3;
            pal[3][k] = (pal[0][k] + 2 * pal[1][k]) / 3;
        }
    } else {
        for (int k = 0; k < 3; k++) {
            pal[2][k] = (pal[0][k] + pal[1][k]) / 2;
            pal[3][k] = 0;
        }
        pal[3][3] = 0;
    }
}

static void bc4_palette(const uint8_t *blk, int pal[8])
{
    int a0 = blk[0], a1 = blk[1];
    pal[0] = a0;
    pal[1] = a1;
    if (a0 > a1) {
        for (int i = 2; i < 8; i++) pal[i] = ((8 - i) * a0 + (i - 1) * a1) / 7;
    } else {
        for (int i = 2; i < 6; i++) pal[i] = ((6 - i) * a0 + (i - 1) * a1) / 5;
        pal[6] = 0;
        pal[7] = 255;
    }
}

static void bc4_block_values(const uint8_t *blk, int out[16])
{
    int pal[8];
    bc4_palette(blk, pal);
    uint64_t bits = 0;
    for (int i = 0; i < 6; i++) bits |= (uint64_t)blk[2 + i] << (8 * i);
    for (int t = 0; t < 16; t++) out[t] = pal[(bits >> (3 * t)) & 7];
}

static void bc1_block_texels(const uint8_t *blk, int force4, int out[16][4])
{
    int pal[4][4];
    bc1_palette(blk, force4, pal);
    uint32_t bits = blk[4] | (blk[5] << 8) | (blk[6] << 16) | ((uint32_t)blk[7] << 24);
    for (int t = 0; t < 16; t++) {
        int idx = (bits >> (2 * t)) & 3;
        for (int k = 0; k < 4; k++) out[t][k] = pal[idx][k];
    }
}

void orc_bc1_decode(const uint8_t *blocks, int w, int h, uint8_t *rgba)
{
    int bw = w / 4, bh = h / 4;
    for (int by = 0; by < bh; by++)
        for (int bx = 0; bx < bw; bx++) {
            int tex[16][4];
            bc1_block_texels(blocks + 8 * ((size_t)by * bw + bx), 0, tex);
            for (int t = 0; t < 16; t++) {
                uint8_t *p = rgba + 4 * ((size_t)(4 * by + t / 4) * w + 4 * bx + t % 4);
                for (int k = 0; k < 4; k++) p[k] = (uint8_t)tex[t][k];
            }
        }
}

void orc_bc3_decode(const uint8_t *blocks, int w, int h, uint8_t *rgba)
{
    int bw = w / 4, bh = h / 4;
    for (int by = 0; by < bh; by++)
        for (int bx = 0; bx < bw; bx++) {
            const uint8_t *blk = blocks + 16 * ((size_t)by * bw + bx);
            int tex[16][4], al[16];
            bc4_block_values(blk, al);
            bc1_block_texels(blk + 8, 1, tex);
            for (int t = 0; t < 16; t++) {
                uint8_t *p = rgba + 4 * ((size_t)(4 * by + t / 4) * w + 4 * bx + t % 4);
                p[0] = (uint8_t)tex[t][0];
                p[1] = (uint8_t)tex[t][1];
                p[2] = (uint8_t)tex[t][2];
                p[3] = (uint8_t)al[t];
            }
        }
}

void orc_bc4_decode(const uint8_t *blocks, int w, int h, uint8_t *gray)
{
    int bw = w / 4, bh = h / 4;
    for (int by = 0; by < bh; by++)
        for (int bx = 0; bx < bw; bx++) {
            int v[16];
            bc4_block_values(blocks + 8 * ((size_t)by * bw + bx), v);
            for (int t = 0; t < 16; t++) gray[(size_t)(4 * by + t / 4) * w + 4 * bx + t % 4] = (uint8_t)v[t];
        }
}

static int clamp255(int v) { return v < 0 ? 0 : v > 255 ? 255 : v; }

/* Scaled YCoCg (van Waveren & Castano 2007): texel (R',G',B',A) = (Co', Cg', scale bits, Y).
 * scale = (B' >> 3) + 1 in {1,2,4}; Co = (R'-128)/scale, Cg = (G'-128)/scale;
 * R = Y + Co - Cg, G = Y + Cg, B = Y - Co - Cg.  Done in exact quarter units, round half up. */
void orc_ycocg_dxt5_decode(const uint8_t *blocks, int w, int h, uint8_t *rgba)
{
    int bw = w / 4, bh = h / 4;
    for (int by = 0; by < bh; by++)
        for (int bx = 0; bx < bw; bx++) {
            const uint8_t *blk = blocks + 16 * ((size_t)by * bw + bx);
            int tex[16][4], yv[16];
            bc4_block_values(blk, yv);
            bc1_block_texels(blk + 8, 1, tex);
            for (int t = 0; t < 16; t++) {
                int scale = (tex[t][2] >> 3) + 1;
                int q = scale >= 4 ? 1 : scale >= 2 ? 2 : 4; /* quarter units per stored step */
                if (scale == 3) q = 1;                        /* B'=16 is never written; treat as 4 */
                int co4 = (tex[t][0] - 128) * q, cg4 = (tex[t][1] - 128) * q, y4 = 4 * yv[t];
                uint8_t *p = rgba + 4 * ((size_t)(4 * by + t / 4) * w + 4 * bx + t % 4);
                p[0] = (uint8_t)clamp255((y4 + co4 - cg4 + 2) >> 2);
                p[1] = (uint8_t)clamp255((y4 + cg4 + 2) >> 2);
                p[2] = (uint8_t)clamp255((y4 - co4 - cg4 + 2) >> 2);
                p[3] = 255;
            }
        }
}

/* ---- forward scaled YCoCg ----------------------------------------------------------------------- */

/* per block: Y = round((R+2G+B)/4); co2 = R-B (half units), cg4 = -R+2G-B (quarter units);
 * scale = largest of {4,2,1} keeping |Co*scale|,|Cg*scale| <= 127; stored with the fraction kept. */
static void ycocg_block_texels(const uint8_t px[16][4], uint8_t tex[16][4])
{
    int co2[16], cg4[16], m2 = 0, m4 = 0;
    for (int t = 0; t < 16; t++) {
        int r = px[t][0], g = px[t][1], b = px[t][2];
        co2[t] = r - b;
        cg4[t] = -r + 2 * g - b;
        int a2 = co2[t] < 0 ? -co2[t] : co2[t], a4 = cg4[t] < 0 ? -cg4[t] : cg4[t];
        if (a2 > m2) m2 = a2;
        if (a4 > m4) m4 = a4;
    }
    /* |Co| = m2/2, |Cg| = m4/4 */
    int scale = 1;
    if (m2 * 4 <= 127 * 2 && m4 * 4 <= 127 * 4) scale = 4;
    else if (m2 * 2 <= 127 * 2 && m4 * 2 <= 127 * 4) scale = 2;
    for (int t = 0; t < 16; t++) {
        int r = px[t][0], g = px[t][1], b = px[t][2];
        int co = (int)floor(co2[t] * scale / 2.0 + 0.5);
        int cg = (int)floor(cg4[t] * scale / 4.0 + 0.5);
        tex[t][0] = (uint8_t)clamp255(co + 128);
        tex[t][1] = (uint8_t)clamp255(cg + 128);
        tex[t][2] = (uint8_t)((scale - 1) << 3);
        tex[t][3] = (uint8_t)((r + 2 * g + b + 2) >> 2);
    }
}

static void gather_block(const uint8_t *rgba, int w, int bx, int by, uint8_t px[16][4])
{
    for (int t = 0; t < 16; t++)
        memcpy(px[t], rgba + 4 * ((size_t)(4 * by + t / 4) * w + 4 * bx + t % 4), 4);
}

void orc_ycocg_scaled_texels(const uint8_t *rgba, int w, int h, uint8_t *texels)
{
    for (int by = 0; by < h / 4; by++)
        for (int bx = 0; bx < w / 4; bx++) {
            uint8_t px[16][4], tex[16][4];
            gather_block(rgba, w, bx, by, px);
            ycocg_block_texels(px, tex);
            for (int t = 0; t < 16; t++)
                memcpy(texels + 4 * ((size_t)(4 * by + t / 4) * w + 4 * bx + t % 4), tex[t], 4);
        }
}

/* ---- squish-style alpha (BC4) fit --------------------------------------------------------------- */

static void fix_range(int *mn, int *mx, int steps)
{
    if (*mx - *mn < steps) *mx = *mn + steps < 255 ? *mn + steps : 255;
    if (*mx - *mn < steps) *mn = *mx - steps > 0 ? *mx - steps : 0;
}

static int fit_codes(const int v[16], const int codes[8], int idx[16])
{
    int err = 0;
    for (int t = 0; t < 16; t++) {
        int best = 1 << 30, bi = 0;
        for (int c = 0; c < 8; c++) {
            int d = (v[t] - codes[c]) * (v[t] - codes[c]);
            if (d < best) { best = d; bi = c; }
        }
        idx[t] = bi;
        err += best;
    }
    return err;
}

static void write_bc4(uint8_t *blk, int a0, int a1, const int idx[16])
{
    blk[0] = (uint8_t)a0;
    blk[1] = (uint8_t)a1;
    uint64_t bits = 0;
    for (int t = 0; t < 16; t++) bits |= (uint64_t)idx[t] << (3 * t);
    for (int i = 0; i < 6; i++) blk[2 + i] = (uint8_t)(bits >> (8 * i));
}

static void bc4_block_squish(const int v[16], uint8_t *blk)
{
    int min5 = 255, max5 = 0, min7 = 255, max7 = 0;
    for (int t = 0; t < 16; t++) {
        if (v[t] < min7) min7 = v[t];
        if (v[t] > max7) max7 = v[t];
        if (v[t] != 0 && v[t] < min5) min5 = v[t];
        if (v[t] != 255 && v[t] > max5) max5 = v[t];
    }
    if (min5 > max5) min5 = max5;
    fix_range(&min5, &max5, 5);
    fix_range(&min7, &max7, 7);
    int c5[8], c7[8], i5[16], i7[16];
    c5[0] = min5; c5[1] = max5;
    for (int i = 1; i < 5; i++) c5[1 + i] = ((5 - i) * min5 + i * max5) / 5;
    c5[6] = 0; c5[7] = 255;
    c7[0] = min7; c7[1] = max7;
    for (int i = 1; i < 7; i++) c7[1 + i] = ((7 - i) * min7 + i * max7) / 7;
    int e5 = fit_codes(v, c5, i5), e7 = fit_codes(v, c7, i7);
    if (e5 <= e7) {
        /* 6-value mode needs a0 <= a1: min5 <= max5 already holds */
        write_bc4(blk, min5, max5, i5);
    } else {
        /* 8-value mode needs a0 > a1: swap endpoints and mirror the indices */
        if (min7 < max7) {
            for (int t = 0; t < 16; t++) i7[t] = i7[t] == 0 ? 1 : i7[t] == 1 ? 0 : 9 - i7[t];
            write_bc4(blk, max7, min7, i7);
        } else {
            /* equal endpoints: 6-value mode with every index 0 decodes to the same value */
            for (int t = 0; t < 16; t++) i7[t] = 0;
            write_bc4(blk, min7, max7, i7);
        }
    }
}

void orc_bc4_encode_squish(const uint8_t *rgba, int w, int h, int channel, uint8_t *blocks)
{
    for (int by = 0; by < h / 4; by++)
        for (int bx = 0; bx < w / 4; bx++) {
            uint8_t px[16][4];
            int v[16];
            gather_block(rgba, w, bx, by, px);
            for (int t = 0; t < 16; t++) v[t] = px[t][channel];
            bc4_block_squish(v, blocks + 8 * ((size_t)by * (w / 4) + bx));
        }
}

/* ---- cluster fit (colour) ----------------------------------------------------------------------- */

typedef struct { double x, y, z, w; } v4;
static v4 v4add(v4 a, v4 b) { v4 r = {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; return r; }
static v4 v4sub(v4 a, v4 b) { v4 r = {a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; return r; }
static v4 v4scale(v4 a, double s, double sw) { v4 r = {a.x * s, a.y * s, a.z * s, a.w * sw}; return r; }

typedef struct {
    int n;               /* unique colours */
    double p[16][3];     /* in [0,1] */
    double wgt[16];
    int remap[16];       /* texel -> point */
} colour_set;

static void build_set(const uint8_t px[16][4], colour_set *s)
{
    s->n = 0;
    for (int t = 0; t < 16; t++) {
        int found = -1;
        for (int u = 0; u < t; u++)
            if (px[u][0] == px[t][0] && px[u][1] == px[t][1] && px[u][2] == px[t][2]) { found = s->remap[u]; break; }
        if (found >= 0) {
            s->remap[t] = found;
            s->wgt[found] += 1.0;
        } else {
            s->remap[t] = s->n;
            for (int k = 0; k < 3; k++) s->p[s->n][k] = px[t][k] / 255.0;
            s->wgt[s->n] = 1.0;
            s->n++;
        }
    }
}

static void principal_axis(const colour_set *s, double axis[3])
{
    double c[3] = {0, 0, 0}, tw = 0;
    for (int i = 0; i < s->n; i++) {
        tw += s->wgt[i];
        for (int k = 0; k < 3; k++) c[k] += s->wgt[i] * s->p[i][k];
    }
    for (int k = 0; k < 3; k++) c[k] /= tw;
    double m[3][3] = {{0}};
    for (int i = 0; i < s->n; i++) {
        double d[3];
        for (int k = 0; k < 3; k++) d[k] = s->p[i][k] - c[k];
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) m[a][b] += s->wgt[i] * d[a] * d[b];
    }
    /* power iteration from the row of largest diagonal */
    int big = 0;
    if (m[1][1] > m[big][big]) big = 1;
    if (m[2][2] > m[big][big]) big = 2;
    double v[3] = {m[big][0], m[big][1], m[big][2]};
    if (v[0] == 0 && v[1] == 0 && v[2] == 0) { axis[0] = axis[1] = axis[2] = 1; return; }
    for (int it = 0; it < 32; it++) {
        double nv[3];
        for (int a = 0; a < 3; a++) nv[a] = m[a][0] * v[0] + m[a][1] * v[1] + m[a][2] * v[2];
        double mx = fabs(nv[0]);
        if (fabs(nv[1]) > mx) mx = fabs(nv[1]);
        if (fabs(nv[2]) > mx) mx = fabs(nv[2]);
        if (mx == 0) break;
        for (int a = 0; a < 3; a++) v[a] = nv[a] / mx;
    }
    for (int a = 0; a < 3; a++) axis[a] = v[a];
}

typedef struct {
    double best_err;
    double a[3], b[3];       /* endpoints on the 5:6:5 grid, in [0,1] */
    int idx[16];             /* per point (unique colour) palette index, DXT numbering */
    int three;               /* 1 = 3-colour mode solution */
} fit_result;

static const double k_grid[3] = {31.0, 63.0, 31.0};

static double snap(double v, int k)
{
    if (v < 0) v = 0;
    if (v > 1) v = 1;
    return floor(k_grid[k] * v + 0.5) / k_grid[k];
}

/* order[] = points sorted along axis; returns 0 when this ordering was already tried */
static int make_ordering(const colour_set *s, const double axis[3], int order[16], int tried[8][16], int ntried)
{
    double dps[16];
    for (int i = 0; i < s->n; i++) {
        dps[i] = s->p[i][0] * axis[0] + s->p[i][1] * axis[1] + s->p[i][2] * axis[2];
        order[i] = i;
    }
    for (int i = 1; i < s->n; i++)
        for (int j = i; j > 0 && dps[order[j]] < dps[order[j - 1]]; j--) {
            int t = order[j]; order[j] = order[j - 1]; order[j - 1] = t;
        }
    for (int it = 0; it < ntried; it++) {
        int same = 1;
        for (int i = 0; i < s->n; i++) if (tried[it][i] != order[i]) { same = 0; break; }
        if (same) return 0;
    }
    return 1;
}

/* evaluate one partition; alpha weights wa/wb are already folded into the sums */
static void try_endpoints(v4 ax, double a2, v4 bx, double b2, double ab, double *err, double a[3], double b[3])
{
    double det = a2 * b2 - ab * ab;
    double av[3], bv[3];
    double axv[3] = {ax.x, ax.y, ax.z}, bxv[3] = {bx.x, bx.y, bx.z};
    for (int k = 0; k < 3; k++) {
        if (det != 0.0) {
            av[k] = (axv[k] * b2 - bxv[k] * ab) / det;
            bv[k] = (bxv[k] * a2 - axv[k] * ab) / det;
        } else if (a2 > 0) {
            av[k] = axv[k] / a2; bv[k] = av[k];
        } else {
            bv[k] = b2 > 0 ? bxv[k] / b2 : 0; av[k] = bv[k];
        }
        av[k] = snap(av[k], k);
        bv[k] = snap(bv[k], k);
    }
    double e = 0;
    for (int k = 0; k < 3; k++)
        e += av[k] * av[k] * a2 + bv[k] * bv[k] * b2 + 2.0 * (av[k] * bv[k] * ab - av[k] * axv[k] - bv[k] * bxv[k]);
    *err = e;
    memcpy(a, av, sizeof av);
    memcpy(b, bv, sizeof bv);
}

static void cluster_fit(const colour_set *s, int iterations, int allow_three, fit_result *res)
{
    double axis[3];
    principal_axis(s, axis);
    int tried[8][16], ntried = 0;
    res->best_err = 1e300;
    res->three = 0;
    int best_i = 0, best_j = 0, best_k = 0, best_iter = -1, best_order[16];
    for (int iter = 0; iter < iterations && iter < 8; iter++) {
        int order[16];
        if (!make_ordering(s, axis, order, tried, ntried)) break;
        memcpy(tried[ntried++], order, sizeof order);
        v4 pw[16], total = {0, 0, 0, 0};
        for (int i = 0; i < s->n; i++) {
            int j = order[i];
            v4 q = {s->wgt[j] * s->p[j][0], s->wgt[j] * s->p[j][1], s->wgt[j] * s->p[j][2], s->wgt[j]};
            pw[i] = q;
            total = v4add(total, q);
        }
        int improved = 0;
        int n = s->n;
        /* 4-colour: clusters [0,i) w=1, [i,j) w=2/3, [j,k) w=1/3, [k,n) w=0 */
        v4 p0 = {0, 0, 0, 0};
        for (int i = 0; i <= n; i++) {
            v4 p1 = {0, 0, 0, 0};
            for (int j = i; j <= n; j++) {
                v4 p2 = {0, 0, 0, 0};
                for (int k = j; k <= n; k++) {
                    v4 p3 = v4sub(v4sub(v4sub(total, p2), p1), p0);
                    v4 ax = v4add(v4add(p0, v4scale(p1, 2.0 / 3.0, 4.0 / 9.0)), v4scale(p2, 1.0 / 3.0, 1.0 / 9.0));
                    v4 bx = v4add(v4add(p3, v4scale(p2, 2.0 / 3.0, 4.0 / 9.0)), v4scale(p1, 1.0 / 3.0, 1.0 / 9.0));
                    double ab = (2.0 / 9.0) * (p1.w + p2.w);
                    double e, a[3], b[3];
                    try_endpoints(ax, ax.w, bx, bx.w, ab, &e, a, b);
                    if (e < res->best_err) {
                        res->best_err = e;
                        memcpy(res->a, a, sizeof a);
                        memcpy(res->b, b, sizeof b);
                        best_i = i; best_j = j; best_k = k; best_iter = iter; res->three = 0;
                        memcpy(best_order, order, sizeof order);
                        improved = 1;
                    }
                    if (k < n) p2 = v4add(p2, pw[k]);
                }
                if (j < n) p1 = v4add(p1, pw[j]);
            }
            if (i < n) p0 = v4add(p0, pw[i]);
        }
        if (allow_three) {
            /* 3-colour: clusters [0,i) w=1, [i,j) w=1/2, [j,n) w=0 */
            v4 q0 = {0, 0, 0, 0};
            for (int i = 0; i <= n; i++) {
                v4 q1 = {0, 0, 0, 0};
                for (int j = i; j <= n; j++) {
                    v4 q2 = v4sub(v4sub(total, q1), q0);
                    v4 ax = v4add(q0, v4scale(q1, 0.5, 0.25));
                    v4 bx = v4add(q2, v4scale(q1, 0.5, 0.25));
                    double ab = 0.25 * q1.w;
                    double e, a[3], b[3];
                    try_endpoints(ax, ax.w, bx, bx.w, ab, &e, a, b);
                    if (e < res->best_err) {
                        res->best_err = e;
                        memcpy(res->a, a, sizeof a);
                        memcpy(res->b, b, sizeof b);
                        best_i = i; best_j = j; best_k = n; best_iter = iter; res->three = 1;
                        memcpy(best_order, order, sizeof order);
                        improved = 1;
                    }
                    if (j < n) q1 = v4add(q1, pw[j]);
                }
                if (i < n) q0 = v4add(q0, pw[i]);
            }
        }
        if (!improved || best_iter != iter) break;
        for (int k = 0; k < 3; k++) axis[k] = res->b[k] - res->a[k];
    }
    /* DXT index numbering: 0 = a (c0), 1 = b (c1), 2 = 2/3a+1/3b (or midpoint), 3 = 1/3a+2/3b */
    for (int pos = 0; pos < s->n; pos++) {
        int idx;
        if (res->three) idx = pos < best_i ? 0 : pos < best_j ? 2 : 1;
        else idx = pos < best_i ? 0 : pos < best_j ? 2 : pos < best_k ? 3 : 1;
        res->idx[best_order[pos]] = idx;
    }
}

static unsigned pack565(const double c[3])
{
    unsigned r = (unsigned)floor(31.0 * c[0] + 0.5), g = (unsigned)floor(63.0 * c[1] + 0.5), b = (unsigned)floor(31.0 * c[2] + 0.5);
    return (r << 11) | (g << 5) | b;
}

/* writes an 8-byte colour block; four_only = BC3 colour block (must not rely on c0<=c1 semantics) */
static void colour_block(const uint8_t px[16][4], int iterations, int allow_three, uint8_t *blk)
{
    colour_set s;
    build_set(px, &s);
    fit_result r;
    cluster_fit(&s, iterations, allow_three, &r);
    unsigned c0 = pack565(r.a), c1 = pack565(r.b);
    int idx[16];
    for (int t = 0; t < 16; t++) idx[t] = r.idx[s.remap[t]];
    if (!r.three) {
        if (c0 < c1) {
            unsigned tmp = c0; c0 = c1; c1 = tmp;
            for (int t = 0; t < 16; t++) idx[t] ^= 1; /* 0<->1, 2<->3 */
        } else if (c0 == c1) {
            for (int t = 0; t < 16; t++) idx[t] = 0;
            /* c0 == c1 selects 3-colour mode in BC1; index 0 decodes identically in both modes */
        }
    } else {
        if (c0 > c1) {
            unsigned tmp = c0; c0 = c1; c1 = tmp;
            for (int t = 0; t < 16; t++) idx[t] = idx[t] == 0 ? 1 : idx[t] == 1 ? 0 : idx[t];
        }
    }
    blk[0] = (uint8_t)c0; blk[1] = (uint8_t)(c0 >> 8);
    blk[2] = (uint8_t)c1; blk[3] = (uint8_t)(c1 >> 8);
    uint32_t bits = 0;
    for (int t = 0; t < 16; t++) bits |= (uint32_t)idx[t] << (2 * t);
    blk[4] = (uint8_t)bits; blk[5] = (uint8_t)(bits >> 8); blk[6] = (uint8_t)(bits >> 16); blk[7] = (uint8_t)(bits >> 24);
}

void orc_bc1_encode_clusterfit(const uint8_t *rgba, int w, int h, uint8_t *blocks, int iterations)
{
    for (int by = 0; by < h / 4; by++)
        for (int bx = 0; bx < w / 4; bx++) {
            uint8_t px[16][4];
            gather_block(rgba, w, bx, by, px);
            colour_block(px, iterations, 1, blocks + 8 * ((size_t)by * (w / 4) + bx));
        }
}

void orc_bc3_encode_clusterfit(const uint8_t *rgba, int w, int h, uint8_t *blocks, int iterations)
{
    for (int by = 0; by < h / 4; by++)
        for (int bx = 0; bx < w / 4; bx++) {
            uint8_t px[16][4];
            int al[16];
            gather_block(rgba, w, bx, by, px);
            uint8_t *blk = blocks + 16 * ((size_t)by * (w / 4) + bx);
            for (int t = 0; t < 16; t++) al[t] = px[t][3];
            bc4_block_squish(al, blk);
            colour_block(px, iterations, 0, blk + 8);
        }
}

void orc_ycocg_dxt5_encode_clusterfit(const uint8_t *rgba, int w, int h, uint8_t *blocks, int iterations)
{
    for (int by = 0; by < h / 4; by++)
        for (int bx = 0; bx < w / 4; bx++) {
            uint8_t px[16][4], tex[16][4];
            int yv[16];
            gather_block(rgba, w, bx, by, px);
            ycocg_block_texels(px, tex);
            uint8_t *blk = blocks + 16 * ((size_t)by * (w / 4) + bx);
            for (int t = 0; t < 16; t++) yv[t] = tex[t][3];
            bc4_block_squish(yv, blk);
            colour_block(tex, iterations, 0, blk + 8);
        }
}

double orc_mse_rgba(const uint8_t *a, const uint8_t *b, int w, int h, unsigned mask)
{
    double acc = 0;
    size_t cnt = 0;
    for (size_t i = 0; i < (size_t)w * h; i++)
        for (int k = 0; k < 4; k++)
            if (mask & (1u << k)) {
                int d = (int)a[4 * i + k] - (int)b[4 * i + k];
                acc += (double)d * d;
                cnt++;
            }
    return cnt ? acc / (double)cnt : 0.0;
}
