// hap_b200/csrc/hap_codes.h -- numeric constants of the Hap wire format and API, shared by host and
// device code.  Values per /root/reference/source/hap.h:40-61 and hap.c:34-88.
#pragma once
#include <stdint.h>

#include "../../include/hap.h"

namespace hapb200 {

// HapResult (hap.h:55-61), HapCompressor (hap.h:50-53) and the HapTextureFormat enumerators come from
// include/hap.h itself (plain C enums, usable in device code as compile-time constants).

// HapTextureFormat (hap.h:40-48) under short names
enum : uint32_t {
    HapFmt_RGB_DXT1 = 0x83F0,
    HapFmt_RGBA_DXT5 = 0x83F3,
    HapFmt_YCoCg_DXT5 = 0x01,
    HapFmt_A_RGTC1 = 0x8DBB,
    HapFmt_RGBA_BPTC_UNORM = 0x8E8C,
    HapFmt_RGB_BPTC_UFLOAT = 0x8E8F,
    HapFmt_RGB_BPTC_SFLOAT = 0x8E8E,
};

// stored compressor nibbles / chunk compressor bytes (hap.c:41-43)
enum : uint32_t { kHapChunkRaw = 0xA, kHapChunkSnappy = 0xB, kHapComplex = 0xC };

// section types (hap.c:84-88)
enum : uint32_t {
    kSecMultipleImages = 0x0D,
    kSecDecodeInstructions = 0x01,
    kSecCompressorTable = 0x02,
    kSecSizeTable = 0x03,
    kSecOffsetTable = 0x04,
};

constexpr uint32_t kU24Max = 0x00FFFFFFu;       // hap.c:34
constexpr uint32_t kMaxChunkCount = 3355431u;   // hap.c:281

}  // namespace hapb200
