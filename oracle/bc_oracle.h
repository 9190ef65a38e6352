/*
 * oracle/bc_oracle.h -- TEST INFRASTRUCTURE ONLY (never linked into libhap_b200.so).
 *
 * The S3TC / RGTC / scaled-YCoCg block codecs that sit upstream and downstream of the reference's
 * HapEncode / HapDecode.  NOTHING of this is under /root/reference: the reference takes
 * already-compressed DXT bytes (hap.h:82-104) and only links to the format specs
 * (documentation/HapVideoDRAFT.md:22-27).  PARITY UNPINNED: there is no reference test, fixture or
 * source to pin these against.  What this file provides instead:
 *   - block DECODERS restated from the public S3TC / RGTC specs, cross-checked in
 *     tests/test_oracle_bc.py against Pillow's independent "bcn" decoder;
 *   - a "squish-HIGH"-class ENCODER (weighted cluster fit over all ordered 4/3-partitions along the
 *     principal axis, iterated up to 8x; squish-style 5/7-interpolant alpha fit) restated from the
 *     published description of libsquish's kColourIterativeClusterFit.  It is the QUALITY BAR the
 *     CUDA encoders are held to (PSNR within 0.1 dB, BASELINE.json north_star), not a byte oracle.
 * Layouts: RGBA8 row-major, stride = 4*w; blocks row-major over (w/4) x (h/4); w,h multiples of 4.
 */
#ifndef ORACLE_BC_ORACLE_H
#define ORACLE_BC_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* decoders -> RGBA8 (BC4 -> one byte per texel) */
void orc_bc1_decode(const uint8_t *blocks, int w, int h, uint8_t *rgba);
void orc_bc3_decode(const uint8_t *blocks, int w, int h, uint8_t *rgba);
void orc_bc4_decode(const uint8_t *blocks, int w, int h, uint8_t *gray);
/* scaled YCoCg DXT5 -> RGBA8 with alpha = 255 */
void orc_ycocg_dxt5_decode(const uint8_t *blocks, int w, int h, uint8_t *rgba);

/* forward colour transform used by the YCoCg path: writes the BC3-ready texels
 * (R'=Co*scale+128, G'=Cg*scale+128, B'=(scale-1)*8, A=Y) for every pixel, block by block */
void orc_ycocg_scaled_texels(const uint8_t *rgba, int w, int h, uint8_t *texels);

/* cluster-fit encoders; iterations = 1 (squish NORMAL) or 8 (squish HIGH) */
void orc_bc1_encode_clusterfit(const uint8_t *rgba, int w, int h, uint8_t *blocks, int iterations);
void orc_bc3_encode_clusterfit(const uint8_t *rgba, int w, int h, uint8_t *blocks, int iterations);
void orc_ycocg_dxt5_encode_clusterfit(const uint8_t *rgba, int w, int h, uint8_t *blocks, int iterations);
/* channel = 0..3 selects which RGBA byte is compressed (Hap Alpha-Only / Hap Q Alpha use 3) */
void orc_bc4_encode_squish(const uint8_t *rgba, int w, int h, int channel, uint8_t *blocks);

/* mean squared error over the given channel mask (bit c = channel c of RGBA) */
double orc_mse_rgba(const uint8_t *a, const uint8_t *b, int w, int h, unsigned channel_mask);

#ifdef __cplusplus
}
#endif
#endif
